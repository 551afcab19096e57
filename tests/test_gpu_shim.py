"""the UNMODIFIED reference program with its demodulator swapped for the library.

`make -C oracle readsb-pair` (run by __graft_entry__.build() where /root/reference exists) leaves two programs in
oracle/_ref/: readsb_cpu, the stock build of the reference's own sources, and readsb_b200, the very same objects
linked with integration/readsb_shim.c and `-Wl,--wrap=demodulate2400 ...` so that readsb.c:871-874 calls into
libb200demod.so.  Both replay the same capture (`--device-type ifile`, the reference's own replay path,
sdr_ifile.c:169-259); every frame line of `--mlat --raw` (12 MHz timestamp + frame bytes, in order) and every
demodulator counter of `--stats` (stats.c:82-122) must be identical.
"""
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from readsb_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
CPU = ROOT / "oracle" / "_ref" / "readsb_cpu"
GPU = ROOT / "oracle" / "_ref" / "readsb_b200"
GPU_IQ = ROOT / "oracle" / "_ref" / "readsb_b200_iq"      # + integration/readsb_shim_iq.c: the converter call site keeps the raw IQ, the GPU converts

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (CPU.exists() and GPU.exists()), reason="oracle/_ref/readsb_{cpu,b200} not built")]


_EMU_DIR = None


def _env_for(exe):
    """B200_EMU=1 (tests/emu/): the swapped-demodulator program resolves libb200demod.so to the emulated library — LD_LIBRARY_PATH
    is searched before the binary's RUNPATH — so the whole unmodified reference program runs against the kernels' source on the CPU."""
    global _EMU_DIR
    env = dict(os.environ)
    if exe in (GPU, GPU_IQ) and os.environ.get("B200_EMU") == "1":
        if _EMU_DIR is None:
            import tempfile
            _EMU_DIR = tempfile.mkdtemp(prefix="b200emu_")
            os.symlink(os.environ["B200_DEMOD_LIB"], os.path.join(_EMU_DIR, "libb200demod.so"))
        env["LD_LIBRARY_PATH"] = _EMU_DIR + (":" + env["LD_LIBRARY_PATH"] if env.get("LD_LIBRARY_PATH") else "")
    return env


def _run_pair(args, timeout=180, exes=(CPU, GPU)):
    procs = [subprocess.Popen([str(exe)] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=_env_for(exe)) for exe in exes]
    outs = []
    for p in procs:
        out, _ = p.communicate(timeout=timeout)
        assert p.returncode == 0, out[-2000:]
        outs.append(out)
    return outs


def _stats_block(text):
    """The receiver section of --stats: from 'Local receiver:' to the CPR counters (wall-clock lines excluded)."""
    lines = text.splitlines()
    i = next(k for k, l in enumerate(lines) if l.startswith("Local receiver:"))
    j = next(k for k, l in enumerate(lines) if "total usable messages" in l)
    block = [l for l in lines[i:j + 1] if "samples lost" not in l]       # the ring-buffer overrun counter depends on host timing
    return block


@pytest.mark.parametrize("name,extra,gen", [
    ("default_modeac", ["--modeac"], dict(frames_per_sec=2500.0, df_mask=synth.MODEAC | synth.DF17 | synth.DF11 | synth.AP, n_icao=24,
                                          amp=(0.3, 0.9), p_bit_error=0.2)),
    ("buf128_thr40", ["--sdr-buffer-size=128", "--preamble-threshold=40"],
     dict(frames_per_sec=6000.0, df_mask=synth.DF17 | synth.DF11 | synth.AP | synth.DF18 | synth.DF11_IID, n_icao=48,
          p_bit_error=0.35, p_two_bit_error=0.08)),
])
def test_reference_program_with_swapped_demodulator(cuda, tmp_path, name, extra, gen):
    nsamples = 2_400_000 * 3 + 12345                      # 3 s and a partial last buffer
    cap = tmp_path / f"{name}.bin"
    synth.generate(nsamples, seed=2024, **gen).tofile(cap)
    base = ["--device-type", "ifile", "--ifile", str(cap)] + extra

    raw_cpu, raw_gpu = _run_pair(base + ["--mlat", "--raw"])
    frames_cpu = [l for l in raw_cpu.splitlines() if l.startswith("@")]
    frames_gpu = [l for l in raw_gpu.splitlines() if l.startswith("@")]
    assert len(frames_cpu) > 1000
    assert "b200 demodulator" not in raw_gpu
    assert frames_gpu == frames_cpu, f"first difference at line {next(i for i, (a, b) in enumerate(zip(frames_cpu, frames_gpu)) if a != b) if len(frames_cpu) == len(frames_gpu) else (len(frames_cpu), len(frames_gpu))}"

    st_cpu, st_gpu = _run_pair(base + ["--quiet", "--stats"])
    block_cpu, block_gpu = _stats_block(st_cpu), _stats_block(st_gpu)
    assert any(re.search(r"[1-9]\d* accepted with correct CRC", l) for l in block_cpu)
    if "--modeac" in extra:
        assert any(re.search(r"[1-9]\d* Mode A/C messages received", l) for l in block_cpu)
    assert block_gpu == block_cpu, "\n".join(f"{a!r} | {b!r}" for a, b in zip(block_cpu, block_gpu) if a != b)


@pytest.mark.skipif(not GPU_IQ.exists(), reason="oracle/_ref/readsb_b200_iq not built")
@pytest.mark.parametrize("extra", [["--modeac"], ["--sdr-buffer-size=128"]])
def test_reference_program_with_converter_hook(cuda, tmp_path, extra):
    """SURVEY.md 8(b), upstream hand-off #1 / north_star "drops in behind the existing sdr.c callback": init_converter is wrapped
    too (integration/readsb_shim_iq.c), the frontend's converter call only keeps the raw uc8 IQ and convert_uc8_nodc
    (convert.c:64-108) runs on the GPU, fused into the scan kernel.  Frames, every demodulator counter and the noise / signal power
    lines of --stats (which need the converter's exact mean_power) equal the stock program's."""
    nsamples = 2_400_000 * 2 + 4321
    cap = tmp_path / "iqhook.bin"
    synth.generate(nsamples, seed=77, frames_per_sec=3000.0, df_mask=synth.MODEAC | synth.DF17 | synth.DF11 | synth.AP, n_icao=24,
                   amp=(0.3, 0.9), p_bit_error=0.2).tofile(cap)
    base = ["--device-type", "ifile", "--ifile", str(cap)] + extra
    raw_cpu, raw_gpu = _run_pair(base + ["--mlat", "--raw"], exes=(CPU, GPU_IQ))
    assert "UC8 conversion moved to the GPU" in raw_gpu and "b200 demodulator" not in raw_gpu
    frames_cpu = [l for l in raw_cpu.splitlines() if l.startswith("@")]
    frames_gpu = [l for l in raw_gpu.splitlines() if l.startswith("@")]
    assert len(frames_cpu) > 500 and frames_gpu == frames_cpu
    st_cpu, st_gpu = _run_pair(base + ["--quiet", "--stats"], exes=(CPU, GPU_IQ))
    block_cpu, block_gpu = _stats_block(st_cpu), _stats_block(st_gpu)
    assert any("noise power" in l for l in block_cpu)
    assert block_gpu == block_cpu, "\n".join(f"{a!r} | {b!r}" for a, b in zip(block_cpu, block_gpu) if a != b)
