"""Where the scan kernel's TIME goes (not its instructions): the same bench-shaped launch on inputs that switch parts of the work off.
  const     every byte 128: no position passes the pre-check -> staging + conversion + window pass only
  noise200  receiver noise, preamble threshold 200: pre-check passers are compacted and thresholded, nothing passes -> no DF gates, no slices
  noise     receiver noise, threshold 58 (no frames)
  cfg2      the bench workload
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from readsb_b200 import synth
from readsb_b200.demod import Demodulator
S, B, BUF = 256, 8, 65536
n = 2 * B * BUF


def run(name, host, thr):
    pad = 4096
    dev = torch.zeros(pad + host.size + 256, dtype=torch.uint8, device="cuda")
    dev[pad:pad + host.size] = torch.from_numpy(host.reshape(-1)).cuda()
    d = Demodulator(n_streams=S, buf_samples=BUF, max_buffers_per_run=B, preamble_threshold=thr)
    ts = []
    for k in range(8):
        d.run_device(dev.data_ptr() + pad + (k % 2) * B * BUF * 2, host.shape[1], B, BUF, continues=(k % 2) > 0, first_sample_timestamp=k * B * BUF * 5)
        ts.append(d.timing()["scan_ms"])
    c = d.debug_counters()
    print(f"{name:10s} scan_ms {np.median(ts[2:]):.4f}  positions {c['positions']} records {c['records']} frames {c['frames']}", flush=True)
    d.close()


base = [synth.config2_stream(900 + i, n) for i in range(16)]
cfg2 = np.stack([np.roll(base[s % 16], 2 * 1013 * (s // 16)) for s in range(S)])
noise_base = [synth.generate(n, seed=700 + i, frames_per_sec=0.0) for i in range(16)]
noise = np.stack([np.roll(noise_base[s % 16], 2 * 1013 * (s // 16)) for s in range(S)])
const = np.full_like(cfg2, 128)
run("const", const, 58)
run("noise200", noise, 200)
run("noise", noise, 58)
run("cfg2", cfg2, 58)
