"""Times the Mode A/C leg (modeac_ms of b200_demod_last_timing) next to the Mode S leg on a bench-shaped device step."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from readsb_b200 import synth
from readsb_b200.demod import Demodulator
S, B, BUF = 256, 8, 65536
base = [synth.modeac_stream(900 + i, B * BUF) for i in range(16)]
host = np.stack([np.roll(base[s % 16], 2 * 1013 * (s // 16)) for s in range(S)])
pad = 4096
dev = torch.zeros(pad + host.size + 4096, dtype=torch.uint8, device="cuda")
dev[pad:pad + host.size] = torch.from_numpy(host.reshape(-1)).cuda()
for mode_ac in (False, True):
    d = Demodulator(n_streams=S, buf_samples=BUF, max_buffers_per_run=B, mode_ac=mode_ac)
    for k in range(5):
        d.run_device(dev.data_ptr() + pad, host.shape[1], B, BUF, continues=False, first_sample_timestamp=0)
        print("mode_ac", mode_ac, k, d.timing(), "frames", d.total_frames(),
              "modeac", sum(len(d.modeac(s)) for s in range(S)) if mode_ac else 0)
    d.close()
