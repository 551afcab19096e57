"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (oracle/_ref/libreadsb_ref.so, built from
/root/reference by oracle/Makefile) on small seeded synthetic captures.  Only runs where /root/reference exists;
the fixtures it writes are committed so the pin holds on boxes without the reference tree.

    python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

from oraclelib import Reference  # noqa: E402
from readsb_b200 import synth  # noqa: E402

CASES = {
    # name: (generator kwargs, nsamples, buf_samples, demod options)
    "df17_sparse": (dict(seed=41, frames_per_sec=1500.0, df_mask=synth.DF17, n_icao=8), 70000, 32768, dict()),
    "dense_df11_df17": (dict(seed=42, frames_per_sec=10000.0, df_mask=synth.DF17 | synth.DF11, n_icao=16), 70000, 32768, dict()),
    "mixed_biterrors": (dict(seed=43, frames_per_sec=6000.0, df_mask=synth.DF17 | synth.DF11 | synth.AP | synth.DF18 | synth.DF11_IID,
                             n_icao=6, p_bit_error=0.4, p_two_bit_error=0.1), 70000, 32768, dict()),
    "modeac_mix": (dict(seed=46, frames_per_sec=3000.0, df_mask=synth.MODEAC | synth.DF17, n_icao=8, amp=(0.4, 0.9)),
                   120000, 32768, dict()),
    "mixed_nofix": (dict(seed=44, frames_per_sec=6000.0, df_mask=synth.DF17 | synth.DF11 | synth.AP, n_icao=6, p_bit_error=0.4),
                    50000, 65536, dict(nfix_crc=0)),
    "mixed_thr40_nofixdf": (dict(seed=45, frames_per_sec=6000.0, df_mask=synth.DF17 | synth.DF11 | synth.AP, n_icao=6, p_bit_error=0.4),
                            50000, 65536, dict(preamble_threshold=40, fix_df=0)),
}


def main():
    for name, (gen_kw, ns, buf, opts) in CASES.items():
        iq = synth.generate(ns, **gen_kw)
        ref = Reference(**opts)
        frames, levels, bufres, ml, mp = ref.run_stream(iq, buf)
        stats, dstats = ref.stats()
        modeac = Reference(**opts).run_stream_ac(iq, buf)      # demodulate2400AC of the same reference build, same buffers
        np.savez_compressed(HERE / f"{name}.npz", iq=iq, modeac=modeac, frames=frames, signal_level=levels, bufres=bufres, mean_level=ml,
                            mean_power=mp, meta=np.frombuffer(json.dumps(dict(buf_samples=buf, options=opts, stats=stats, dstats=dstats,
                                                                             generator=gen_kw)).encode(), dtype=np.uint8))
        print(name, len(frames), "frames", stats["demod_preambles"], "preambles")


if __name__ == "__main__":
    main()
