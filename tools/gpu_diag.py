"""First-contact diagnostics on a GPU box: CUDA path vs the oracle on a few captures, verbose on mismatch."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from oraclelib import Oracle
from readsb_b200 import synth
from readsb_b200.abi import FRAME_PARITY_FIELDS, frame_hex
from readsb_b200.demod import Demodulator

def compare(name, fg, fo, bg, bo, sg, so):
    ok = True
    if len(fg) != len(fo):
        print(f"[{name}] frame count gpu {len(fg)} oracle {len(fo)}"); ok = False
    n = min(len(fg), len(fo))
    for f in FRAME_PARITY_FIELDS:
        a, b = fg[f][:n], fo[f][:n]
        if f == "addr": a, b = a & 0xffffff, b & 0xffffff
        ne = np.nonzero((a != b).reshape(n, -1).any(axis=1))[0] if n else []
        if len(ne):
            ok = False
            i = ne[0]
            print(f"[{name}] field {f}: {len(ne)} mismatches, first at {i}: gpu {fg[f][i]} oracle {fo[f][i]} (gpu ts {fg['timestamp'][i]} j {fg['j'][i]} seq {fg['buffer_seq'][i]} {frame_hex(fg[i])} | oracle ts {fo['timestamp'][i]} j {fo['j'][i]} seq {fo['buffer_seq'][i]} {frame_hex(fo[i])})")
    for f in ("sum_level", "sum_power", "sum_signal_power", "length", "n_frames", "buffer_seq", "icao_flipped", "sample_timestamp"):
        if len(bg) != len(bo) or not np.array_equal(bg[f], bo[f]):
            ok = False
            print(f"[{name}] bufres {f} differs: gpu {bg[f][:6]} oracle {bo[f][:6]} (n {len(bg)} vs {len(bo)})")
    for k in so:
        if sg[k] != so[k]:
            ok = False
            print(f"[{name}] stats {k}: gpu {sg[k]} oracle {so[k]}")
    print(f"[{name}] {'OK' if ok else 'MISMATCH'} frames={len(fg)}")
    return ok

def main():
    allok = True
    for name, gen, ns in [("cfg2", synth.config2_stream, 600000), ("cfg5", synth.config5_stream, 600000), ("mixed", synth.mixed_stream, 600000)]:
        iq = gen(7, ns)
        for buf, K in ((65536, 1), (65536, 4)):
            o = Oracle(); fo, bo = o.run_stream(iq, buf)
            d = Demodulator(n_streams=1, buf_samples=buf, max_buffers_per_run=K)
            t = time.time(); fg, bg = d.replay(iq); dt = time.time() - t
            print("   debug", d.debug_counters(), "gpu stats", {k: v for k, v in d.stats(0).items() if "Phase" not in k})
            allok &= compare(f"{name}/buf{buf}/K{K}", fg, fo, bg, bo, d.stats(0), o.stats())
            print("   timing", d.timing(), "replay %.3fs" % dt)
            d.close()
    print("ALL OK" if allok else "SOME MISMATCH")
    return 0 if allok else 1

if __name__ == "__main__":
    sys.exit(main())
