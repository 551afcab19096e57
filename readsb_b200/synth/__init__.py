"""ctypes front end of the synthetic capture generator (modes_synth.c)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

DF17, DF11, AP, DF18, DF11_IID, MODEAC = 0x01, 0x02, 0x04, 0x08, 0x10, 0x20
_LIB = None


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("frames_per_sec", C.c_double), ("df_mask", C.c_uint32),
                ("n_icao", C.c_uint32), ("amp_min", C.c_double), ("amp_max", C.c_double),
                ("noise_sigma", C.c_double), ("p_bit_error", C.c_double), ("p_two_bit_error", C.c_double)]


class SynthTruth(C.Structure):
    _fields_ = [("start_tick", C.c_int64), ("msg", C.c_uint8 * 14), ("nbits", C.c_uint8), ("errors", C.c_uint8)]


def lib():
    global _LIB
    if _LIB is None:
        from ..build import build_synth
        path = build_synth()
        _LIB = C.CDLL(str(path))
        _LIB.synth_generate_uc8.restype = C.c_long
        _LIB.synth_generate_uc8.argtypes = [C.POINTER(SynthParams), C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
        _LIB.synth_crc24.restype = C.c_uint32
        _LIB.synth_crc24.argtypes = [C.c_void_p, C.c_int]
    return _LIB


def generate(nsamples: int, seed: int = 1, frames_per_sec: float = 100.0, df_mask: int = DF17,
             n_icao: int = 64, amp=(0.1, 0.8), noise_sigma: float = 1.75, p_bit_error: float = 0.0,
             p_two_bit_error: float = 0.0, out: np.ndarray | None = None, want_truth: bool = False):
    """Returns a uint8 array of 2*nsamples interleaved I,Q bytes (and the truth list if asked)."""
    p = SynthParams(seed, frames_per_sec, df_mask, n_icao, amp[0], amp[1], noise_sigma, p_bit_error, p_two_bit_error)
    if out is None:
        out = np.empty(2 * nsamples, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= 2 * nsamples and out.flags.c_contiguous
    cap = int(frames_per_sec * nsamples / 2.4e6 + 2) if want_truth else 0
    truth = (SynthTruth * cap)() if cap else None
    n = lib().synth_generate_uc8(C.byref(p), nsamples, out.ctypes.data, C.cast(truth, C.c_void_p) if cap else None, cap)
    if n < 0:
        raise MemoryError("synth_generate_uc8 failed")
    if want_truth:
        return out, [(t.start_tick, bytes(t.msg[: t.nbits // 8]), t.errors) for t in truth[:n]]
    return out


# the named workloads of BASELINE.json / SURVEY.md section 8d
def config2_stream(seed: int, nsamples: int, out=None):
    """single stream, DF17 frames injected at 100/s (BASELINE configs[1])."""
    return generate(nsamples, seed=seed, frames_per_sec=100.0, df_mask=DF17, n_icao=64, out=out)


def config5_stream(seed: int, nsamples: int, out=None):
    """dense-preamble stress: 10k DF11+DF17 per second with overlaps (BASELINE configs[4])."""
    return generate(nsamples, seed=seed, frames_per_sec=10000.0, df_mask=DF17 | DF11, n_icao=64, out=out)


def modeac_stream(seed: int, nsamples: int, frames_per_sec: float = 3000.0, out=None):
    """Mode A/C replies mixed with Mode S traffic (for the --modeac path, demod_2400.c:575-761)."""
    return generate(nsamples, seed=seed, frames_per_sec=frames_per_sec, df_mask=MODEAC | DF17 | DF11, n_icao=16,
                    amp=(0.15, 0.9), out=out)


def mixed_stream(seed: int, nsamples: int, frames_per_sec: float = 2000.0, out=None):
    """every DF kind the demodulator understands plus injected 1- and 2-bit errors."""
    return generate(nsamples, seed=seed, frames_per_sec=frames_per_sec,
                    df_mask=DF17 | DF11 | AP | DF18 | DF11_IID, n_icao=48,
                    p_bit_error=0.35, p_two_bit_error=0.08, out=out)
