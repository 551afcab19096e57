"""Prints the library's per-kernel CUDA-event timings for a bench-shaped device-resident step."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from readsb_b200 import synth
from readsb_b200.demod import Demodulator
S, B, BUF = 256, 8, 65536
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
gen = synth.config5_stream if wl == "cfg5" else synth.config2_stream
base = [gen(900 + i, 2 * B * BUF) for i in range(16)]
host = np.stack([np.roll(base[s % 16], 2 * 1013 * (s // 16)) for s in range(S)])
pad = 4096
dev = torch.zeros(pad + host.size + 256, dtype=torch.uint8, device="cuda")
dev[pad:pad + host.size] = torch.from_numpy(host.reshape(-1)).cuda()
d = Demodulator(n_streams=S, buf_samples=BUF, max_buffers_per_run=B)
for k in range(6):
    d.run_device(dev.data_ptr() + pad + (k % 2) * B * BUF * 2, host.shape[1], B, BUF, continues=(k % 2) > 0, first_sample_timestamp=k * B * BUF * 5)
    print(wl, k, d.timing(), d.debug_counters())
