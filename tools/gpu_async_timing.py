"""Per-step timings of the pipelined (async) device-resident path vs the blocking one."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from readsb_b200 import synth
from readsb_b200.demod import Demodulator
S, B, BUF, R = 256, 8, 65536, 2
base = [synth.config2_stream(900 + i, R * B * BUF) for i in range(16)]
host = np.stack([np.roll(base[s % 16], 2 * 1013 * (s // 16)) for s in range(S)])
pad = 4096
dev = torch.zeros(pad + host.size + 256, dtype=torch.uint8, device="cuda")
dev[pad:pad + host.size] = torch.from_numpy(host.reshape(-1)).cuda()
d = Demodulator(n_streams=S, buf_samples=BUF, max_buffers_per_run=B)
def args(k): return (dev.data_ptr() + pad + (k % R) * B * BUF * 2, host.shape[1], B, BUF, (k % R) > 0, k * B * BUF * 5)
for k in range(4): d.run_device(*args(k))
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(4, 24): d.run_device(*args(k))
torch.cuda.synchronize(); print("blocking  ms/step", (time.perf_counter() - t0) / 20 * 1e3, d.timing())
for mode in ("async",):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ts = []
    for k in range(24, 44):
        d.run_device_async(*args(k))
        if k > 24:
            d.wait(); ts.append((time.perf_counter() - t0, d.timing()))
    d.wait(); ts.append((time.perf_counter() - t0, d.timing()))
    torch.cuda.synchronize(); print("pipelined ms/step", (time.perf_counter() - t0) / 20 * 1e3)
    prev = 0
    for t, tm in ts[:8]:
        print("  +%.3f ms  scan %.3f  ev1->ev2 %.3f  total-in-flight %.3f" % ((t - prev) * 1e3, tm["scan_ms"], tm["resolve_ms"], tm["run_ms"])); prev = t
