"""ctypes front ends of the parity oracle (oracle/libmodes_oracle.so) and of the reference built as a
library (oracle/_ref/libreadsb_ref.so).  TEST INFRASTRUCTURE — never imported by readsb_b200."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from readsb_b200.abi import BUFRES_DTYPE, FRAME_DTYPE, MODEAC_DTYPE, BufferResult, Stats

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "libmodes_oracle.so"
REF_SO = ROOT / "oracle" / "_ref" / "libreadsb_ref.so"
_BUILT = False


def _ensure_built():
    global _BUILT
    if not _BUILT:   # make is a no-op when everything is up to date
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "CC=gcc"], check=True, capture_output=True)
        _BUILT = True


def have_ref() -> bool:
    _ensure_built()
    return REF_SO.exists()


class Oracle:
    """One receiver's worth of the CPU restatement."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            _ensure_built()
            L = C.CDLL(str(ORACLE_SO))
            L.oracle_create.restype = C.c_void_p
            L.oracle_create.argtypes = [C.c_int] * 4
            L.oracle_destroy.argtypes = [C.c_void_p]
            L.oracle_uc8_lut.argtypes = [C.c_void_p]
            L.oracle_convert_uc8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
            L.oracle_crc24.restype = C.c_uint32
            L.oracle_crc24.argtypes = [C.c_void_p, C.c_int]
            L.oracle_crc_diagnose1.argtypes = [C.c_uint32, C.c_int]
            L.oracle_demodulate2400.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int64, C.c_uint64, C.c_uint64,
                                                C.c_void_p, C.c_uint, C.POINTER(C.c_uint), C.POINTER(BufferResult)]
            L.oracle_run_stream_uc8.restype = C.c_long
            L.oracle_run_stream_uc8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_int64, C.c_void_p,
                                                C.c_uint, C.c_void_p, C.c_uint, C.POINTER(C.c_uint)]
            L.oracle_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
            L.oracle_demodulate2400AC.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint, C.POINTER(C.c_uint)]
            L.oracle_icao_add.argtypes = [C.c_void_p, C.c_uint32]
            L.oracle_icao_test.argtypes = [C.c_void_p, C.c_uint32]
            L.oracle_icao_expire.argtypes = [C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, preamble_threshold=58, nfix_crc=1, fix_df=1, icao_ttl_ms=60000):
        self.L = self.lib()
        self.h = self.L.oracle_create(preamble_threshold, nfix_crc, fix_df, icao_ttl_ms)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_destroy(self.h)
            self.h = None

    @classmethod
    def lut(cls) -> np.ndarray:
        out = np.empty(65536, dtype=np.uint16)
        cls.lib().oracle_uc8_lut(out.ctypes.data)
        return out

    @classmethod
    def convert(cls, iq: np.ndarray):
        n = iq.size // 2
        mag = np.empty(n, dtype=np.uint16)
        sl, sp = C.c_uint64(), C.c_uint64()
        cls.lib().oracle_convert_uc8(iq.ctypes.data, mag.ctypes.data, n, C.byref(sl), C.byref(sp))
        return mag, sl.value, sp.value

    @classmethod
    def crc24(cls, msg: bytes) -> int:
        buf = (C.c_uint8 * len(msg)).from_buffer_copy(msg)
        return cls.lib().oracle_crc24(buf, len(msg) * 8)

    @classmethod
    def diagnose1(cls, syndrome: int, bits: int) -> int:
        return cls.lib().oracle_crc_diagnose1(syndrome, bits)

    def demodulate(self, data: np.ndarray, length: int, sample_ts: int, sum_level=0, sum_power=0, cap=8192):
        assert data.dtype == np.uint16 and data.size >= length + 326
        frames = np.zeros(cap, dtype=FRAME_DTYPE)
        n = C.c_uint(0)
        res = BufferResult()
        rc = self.L.oracle_demodulate2400(self.h, data.ctypes.data, length, sample_ts, sum_level, sum_power,
                                          frames.ctypes.data, cap, C.byref(n), C.byref(res))
        assert rc == 0
        return frames[: n.value].copy(), res

    def run_stream(self, iq: np.ndarray, buf_samples: int, first_ts: int = 0, cap: int | None = None):
        nsamples = iq.size // 2
        cap = cap or max(1024, nsamples // 100)
        nbuf_cap = nsamples // buf_samples + 2
        frames = np.zeros(cap, dtype=FRAME_DTYPE)
        bufres = np.zeros(nbuf_cap, dtype=BUFRES_DTYPE)
        nb = C.c_uint(0)
        n = self.L.oracle_run_stream_uc8(self.h, iq.ctypes.data, nsamples, buf_samples, first_ts, frames.ctypes.data,
                                         cap, bufres.ctypes.data, nbuf_cap, C.byref(nb))
        assert n >= 0, "oracle frame capacity exceeded"
        return frames[:n].copy(), bufres[: nb.value].copy()

    def set_preamble_threshold(self, thr: int):
        self.L.oracle_set_preamble_threshold.argtypes = [C.c_void_p, C.c_int]
        self.L.oracle_set_preamble_threshold(self.h, thr)

    def restart_stream(self):
        """The next run_stream call starts with a zero halo (a receiver that was reopened); filter and statistics stay."""
        self.L.oracle_stream_restart.argtypes = [C.c_void_p]
        self.L.oracle_stream_restart(self.h)

    def demodulate_ac(self, data: np.ndarray, length: int, sample_ts: int, sum_level: int, sum_power: int, cap=4096):
        out = np.zeros(cap, dtype=MODEAC_DTYPE)
        n = C.c_uint(0)
        rc = self.L.oracle_demodulate2400AC(self.h, data.ctypes.data, length, sample_ts, sum_level, sum_power, out.ctypes.data, cap, C.byref(n))
        assert rc == 0
        return out[: n.value].copy()

    def demodulate_ac_levels(self, data: np.ndarray, length: int, sample_ts: int, mean_level: float, mean_power: float, cap=4096):
        """demodulate2400AC with the mag_buf's own mean_level / mean_power (whatever converter filled it)."""
        self.L.oracle_demodulate2400AC_levels.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_uint, C.POINTER(C.c_uint)]
        out = np.zeros(cap, dtype=MODEAC_DTYPE)
        n = C.c_uint(0)
        rc = self.L.oracle_demodulate2400AC_levels(self.h, data.ctypes.data, length, sample_ts, mean_level, mean_power, out.ctypes.data, cap, C.byref(n))
        assert rc == 0
        return out[: n.value].copy()

    def run_stream_ac_sc16(self, iq16: np.ndarray, buf_samples: int, q11: bool = False, first_ts: int = 0):
        """Mode A/C over an sc16 capture: the converter's float-accumulated means feed the noise floor (convert.c:243-249)."""
        n = iq16.size // 2
        halo = np.zeros(326, np.uint16)
        outs, off, b = [], 0, 0
        while off < n:
            m = min(buf_samples, n - off)
            mag, sl, sp = self.convert_sc16(iq16[2 * off: 2 * (off + m)], q11)
            data = np.concatenate([halo, mag])
            # `sum_level / nsamples` is a float division whose result is widened to double
            a = self.demodulate_ac_levels(data, m, first_ts + off * 5, float(np.float32(sl) / np.float32(m)), float(np.float32(sp) / np.float32(m)))
            a["buffer_idx"] = b
            outs.append(a)
            halo = data[m: m + 326].copy() if m >= 326 else np.zeros(326, np.uint16)
            off += m; b += 1
        return np.concatenate(outs) if outs else np.zeros(0, MODEAC_DTYPE)

    def run_stream_ac(self, iq: np.ndarray, buf_samples: int, first_ts: int = 0):
        """Mode A/C over a capture, buffer by buffer like the ifile loop (halo carried)."""
        nsamples = iq.size // 2
        halo = np.zeros(326, dtype=np.uint16)
        outs = []
        for b, off in enumerate(range(0, nsamples, buf_samples)):
            ln = min(buf_samples, nsamples - off)
            mag, sl, sp = Oracle.convert(iq[2 * off: 2 * (off + ln)])
            data = np.concatenate([halo, mag]).astype(np.uint16)
            a = self.demodulate_ac(data, ln, first_ts + off * 5, sl, sp)
            a["buffer_idx"] = b
            outs.append(a)
            halo = data[ln: ln + 326].copy() if ln >= 326 else np.zeros(326, dtype=np.uint16)
        return np.concatenate(outs) if outs else np.zeros(0, MODEAC_DTYPE)

    @classmethod
    def convert_sc16(cls, iq16: np.ndarray, q11: bool = False):
        """convert_sc16_nodc / convert_sc16q11_nodc: (magnitudes, float32 sum_level, float32 sum_power)."""
        L = cls.lib()
        L.oracle_convert_sc16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        iq16 = np.ascontiguousarray(iq16, dtype=np.int16)
        n = iq16.size // 2
        mag = np.zeros(n, np.uint16)
        sl, sp = C.c_float(), C.c_float()
        L.oracle_convert_sc16(iq16.ctypes.data, mag.ctypes.data, n, 1 if q11 else 0, C.byref(sl), C.byref(sp))
        return mag, np.float32(sl.value), np.float32(sp.value)

    def run_stream_sc16(self, iq16: np.ndarray, buf_samples: int, q11: bool = False, first_ts: int = 0):
        """The ifile loop over an sc16 capture: converter, halo carry, demodulate2400 per buffer.
        Returns (frames, [(length, float32 sum_level, float32 sum_power) per buffer])."""
        n = iq16.size // 2
        halo = np.zeros(326, np.uint16)
        frames, sums, off = [], [], 0
        while off < n:
            m = min(buf_samples, n - off)
            mag, sl, sp = self.convert_sc16(iq16[2 * off: 2 * (off + m)], q11)
            data = np.concatenate([halo, mag])
            f, _ = self.demodulate(data, m, first_ts + off * 5)
            frames.append(f); sums.append((m, sl, sp))
            halo = data[m: m + 326].copy() if m >= 326 else np.zeros(326, np.uint16)
            off += m
        return (np.concatenate(frames) if frames else np.zeros(0, FRAME_DTYPE)), sums

    @classmethod
    def beast(cls, frames: np.ndarray, modeac: np.ndarray = None, verbatim: bool = False) -> bytes:
        """Beast records (net_io.c:1655-1714) in the reference's output order: per buffer the Mode S frames, then the
        Mode A/C replies.  frames["buffer_seq"] and modeac["buffer_idx"] must count buffers from the same origin."""
        L = cls.lib()
        L.oracle_beast_frame.restype = C.c_uint
        L.oracle_beast_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_beast_modeac.restype = C.c_uint
        L.oracle_beast_modeac.argtypes = [C.c_void_p, C.c_void_p]
        out = bytearray()
        tmp = (C.c_uint8 * 48)()
        frames = np.ascontiguousarray(frames)
        na = 0 if modeac is None else len(modeac)
        if na:
            modeac = np.ascontiguousarray(modeac)
        fi = ai = 0
        while fi < len(frames) or ai < na:
            # next buffer that still has records
            b = int(frames["buffer_seq"][fi]) if fi < len(frames) else 1 << 62
            if ai < na:
                b = min(b, int(modeac["buffer_idx"][ai]))
            while fi < len(frames) and int(frames["buffer_seq"][fi]) == b:
                n = L.oracle_beast_frame(frames[fi:fi + 1].ctypes.data, 1 if verbatim else 0, tmp)
                out += bytes(tmp[:n]); fi += 1
            while ai < na and int(modeac["buffer_idx"][ai]) == b:
                n = L.oracle_beast_modeac(modeac[ai:ai + 1].ctypes.data, tmp)
                out += bytes(tmp[:n]); ai += 1
        return bytes(out)

    def stats(self) -> dict:
        s = Stats()
        self.L.oracle_get_stats(self.h, C.byref(s))
        return s.as_dict()

    def icao_add(self, a):
        self.L.oracle_icao_add(self.h, a)

    def icao_test(self, a) -> bool:
        return bool(self.L.oracle_icao_test(self.h, a))

    def icao_expire(self):
        self.L.oracle_icao_expire(self.h)


class Reference:
    """The reference's own demodulator (one receiver per loaded copy: readsb keeps its state in globals,
    so every instance dlopens a private copy of the library)."""

    def __init__(self, preamble_threshold=58, nfix_crc=1, fix_df=1, icao_ttl_ms=60000):
        assert have_ref(), "reference library not built (needs /root/reference)"
        self._tmp = tempfile.NamedTemporaryFile(prefix="readsb_ref_", suffix=".so", delete=False)
        self._tmp.close()
        shutil.copyfile(REF_SO, self._tmp.name)
        L = self.L = C.CDLL(self._tmp.name)
        os.unlink(self._tmp.name)
        L.ref_init.argtypes = [C.c_int] * 4
        L.ref_uc8_lut.argtypes = [C.c_void_p]
        L.ref_convert_uc8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_crc24.restype = C.c_uint32
        L.ref_crc24.argtypes = [C.c_void_p, C.c_int]
        L.ref_crc_diagnose1.argtypes = [C.c_uint32, C.c_int]
        L.ref_score.argtypes = [C.c_void_p, C.c_int]
        L.ref_icao_add.argtypes = [C.c_uint32]
        L.ref_icao_test.argtypes = [C.c_uint32]
        L.ref_demodulate2400.argtypes = [C.c_void_p, C.c_uint, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_uint, C.POINTER(C.c_uint), C.POINTER(BufferResult)]
        L.ref_run_stream_uc8.restype = C.c_long
        L.ref_run_stream_uc8.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_int64, C.c_void_p, C.c_void_p, C.c_uint,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(C.c_uint)]
        L.ref_get_stats.argtypes = [C.POINTER(Stats), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_demodulate2400AC.argtypes = [C.c_void_p, C.c_uint, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_uint, C.POINTER(C.c_uint)]
        L.ref_modeac_count.restype = C.c_uint64
        L.ref_time_stream_uc8.restype = C.c_double
        L.ref_time_stream_uc8.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.POINTER(C.c_uint)]
        assert L.ref_init(preamble_threshold, nfix_crc, fix_df, icao_ttl_ms) == 0

    def lut(self) -> np.ndarray:
        out = np.empty(65536, dtype=np.uint16)
        self.L.ref_uc8_lut(out.ctypes.data)
        return out

    def convert(self, iq: np.ndarray):
        n = iq.size // 2
        mag = np.empty(n, dtype=np.uint16)
        ml, mp = C.c_double(), C.c_double()
        self.L.ref_convert_uc8(iq.ctypes.data, mag.ctypes.data, n, C.byref(ml), C.byref(mp))
        return mag, ml.value, mp.value

    def crc24(self, msg: bytes) -> int:
        buf = (C.c_uint8 * len(msg)).from_buffer_copy(msg)
        return self.L.ref_crc24(buf, len(msg) * 8)

    def diagnose1(self, syndrome: int, bits: int) -> int:
        return self.L.ref_crc_diagnose1(syndrome, bits)

    def score(self, msg: bytes, validbits: int) -> int:
        buf = (C.c_uint8 * 14).from_buffer_copy(msg.ljust(14, b"\0"))
        return self.L.ref_score(buf, validbits)

    def demodulate(self, data: np.ndarray, length: int, sample_ts: int, mean_level=0.0, mean_power=0.0, cap=8192):
        frames = np.zeros(cap, dtype=FRAME_DTYPE)
        levels = np.zeros(cap, dtype=np.float64)
        n = C.c_uint(0)
        res = BufferResult()
        rc = self.L.ref_demodulate2400(data.ctypes.data, length, sample_ts, mean_level, mean_power, frames.ctypes.data,
                                       levels.ctypes.data, cap, C.byref(n), C.byref(res))
        assert rc == 0
        return frames[: n.value].copy(), levels[: n.value].copy(), res

    def run_stream(self, iq: np.ndarray, buf_samples: int, first_ts: int = 0, cap: int | None = None):
        nsamples = iq.size // 2
        cap = cap or max(1024, nsamples // 100)
        nbuf_cap = nsamples // buf_samples + 2
        frames = np.zeros(cap, dtype=FRAME_DTYPE)
        levels = np.zeros(cap, dtype=np.float64)
        bufres = np.zeros(nbuf_cap, dtype=BUFRES_DTYPE)
        ml = np.zeros(nbuf_cap)
        mp = np.zeros(nbuf_cap)
        nb = C.c_uint(0)
        n = self.L.ref_run_stream_uc8(iq.ctypes.data, nsamples, buf_samples, first_ts, frames.ctypes.data,
                                      levels.ctypes.data, cap, bufres.ctypes.data, ml.ctypes.data, mp.ctypes.data,
                                      nbuf_cap, C.byref(nb))
        assert n >= 0
        k = nb.value
        return frames[:n].copy(), levels[:n].copy(), bufres[:k].copy(), ml[:k].copy(), mp[:k].copy()

    def convert_sc16(self, iq16: np.ndarray, q11: bool = False):
        iq16 = np.ascontiguousarray(iq16, dtype=np.int16)
        n = iq16.size // 2
        mag = np.zeros(n, np.uint16)
        ml, mp = C.c_double(), C.c_double()
        self.L.ref_convert_sc16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        assert self.L.ref_convert_sc16(iq16.ctypes.data, mag.ctypes.data, n, 1 if q11 else 0, C.byref(ml), C.byref(mp)) == 0
        return mag, ml.value, mp.value

    def run_stream_ac(self, iq: np.ndarray, buf_samples: int, first_ts: int = 0):
        nsamples = iq.size // 2
        halo = np.zeros(326, dtype=np.uint16)
        outs = []
        for b, off in enumerate(range(0, nsamples, buf_samples)):
            ln = min(buf_samples, nsamples - off)
            mag, ml, mp = self.convert(iq[2 * off: 2 * (off + ln)])
            data = np.concatenate([halo, mag]).astype(np.uint16)
            out = np.zeros(4096, dtype=MODEAC_DTYPE)
            n = C.c_uint(0)
            assert self.L.ref_demodulate2400AC(data.ctypes.data, ln, first_ts + off * 5, ml, mp, out.ctypes.data, 4096, C.byref(n)) == 0
            a = out[: n.value].copy(); a["buffer_idx"] = b
            outs.append(a)
            halo = data[ln: ln + 326].copy() if ln >= 326 else np.zeros(326, dtype=np.uint16)
        return np.concatenate(outs) if outs else np.zeros(0, MODEAC_DTYPE)

    def run_stream_ac_sc16(self, iq16: np.ndarray, buf_samples: int, q11: bool = False, first_ts: int = 0):
        """The reference's own sc16 converter feeding the reference's demodulate2400AC, buffer by buffer."""
        n = iq16.size // 2
        halo = np.zeros(326, np.uint16)
        outs, off, b = [], 0, 0
        while off < n:
            m = min(buf_samples, n - off)
            mag, ml, mp = self.convert_sc16(iq16[2 * off: 2 * (off + m)], q11)
            data = np.concatenate([halo, mag]).astype(np.uint16)
            out = np.zeros(4096, dtype=MODEAC_DTYPE)
            cnt = C.c_uint(0)
            assert self.L.ref_demodulate2400AC(data.ctypes.data, m, first_ts + off * 5, C.c_double(ml), C.c_double(mp), out.ctypes.data, 4096, C.byref(cnt)) == 0
            a = out[: cnt.value].copy(); a["buffer_idx"] = b
            outs.append(a)
            halo = data[m: m + 326].copy() if m >= 326 else np.zeros(326, np.uint16)
            off += m; b += 1
        return np.concatenate(outs) if outs else np.zeros(0, MODEAC_DTYPE)

    def set_samples_dropped(self, n: int):
        self.L.ref_set_samples_dropped.argtypes = [C.c_uint]
        self.L.ref_set_samples_dropped(n)

    def modeac_count(self) -> int:
        return int(self.L.ref_modeac_count())

    def stats(self):
        s = Stats()
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.L.ref_get_stats(C.byref(s), C.byref(a), C.byref(b), C.byref(c))
        return s.as_dict(), {"signal_power_sum": a.value, "noise_power_sum": b.value, "peak_signal_power": c.value}

    def icao_add(self, a):
        self.L.ref_icao_add(a)

    def icao_test(self, a) -> bool:
        return bool(self.L.ref_icao_test(a))

    def icao_expire(self):
        self.L.ref_icao_expire()

    def time_stream(self, iq: np.ndarray, buf_samples: int):
        nf = C.c_uint(0)
        secs = self.L.ref_time_stream_uc8(iq.ctypes.data, iq.size // 2, buf_samples, C.byref(nf))
        return secs, nf.value
