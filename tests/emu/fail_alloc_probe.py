"""TEST INFRASTRUCTURE: every allocation the host code makes is made to fail once (the emulated runtime's n-th cudaMalloc /
cudaHostAlloc returns cudaErrorMemoryAllocation): create / run / fetch must report an error — never crash, never leak — and a
context created after the fault is gone must work.  Run by tests/test_emu_kernels.py in a subprocess."""
import ctypes
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
os.environ["B200_EMU"] = "1"
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tests" / "emu"))
import build_emu  # noqa: E402

lib = build_emu.build(sanitize=os.environ.get("B200_EMU_SANITIZE") or None)
os.environ["B200_DEMOD_LIB"] = str(lib)
from oraclelib import Oracle  # noqa: E402
from paritylib import diff_frames  # noqa: E402
from readsb_b200 import synth  # noqa: E402
from readsb_b200.demod import DemodError, Demodulator  # noqa: E402

E = ctypes.CDLL(str(lib))
E.emu_fail_alloc_after.argtypes = [ctypes.c_long]
iq = synth.modeac_stream(3, 2 * 20000)
fo, _ = Oracle().run_stream(iq, 20000)
failed_create = failed_later = 0
for n in range(1, 400):
    E.emu_fail_alloc_after(n)
    try:
        d = Demodulator(n_streams=2, buf_samples=20000, max_buffers_per_run=2, mode_ac=True)
    except DemodError:
        failed_create += 1
        continue
    try:        # the fault may still be pending: pipeline slots, Beast buffers are allocated on first use
        fg, _, _ = d.replay(iq, want_modeac=True)
        d.beast(0)
        ok = True
    except DemodError:
        failed_later += 1
        ok = False
    E.emu_fail_alloc_after(0)
    d.close()
    if ok:
        assert not diff_frames(fg, fo)
        print(f"allocations that failed: {failed_create} in create, {failed_later} later; n = {n}: no fault left, results equal the oracle's")
        sys.exit(0)
print("the fault never stopped hitting")
sys.exit(1)
