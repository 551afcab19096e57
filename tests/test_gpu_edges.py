"""-m gpu: edge cases of the boundary — ragged / tiny / empty buffers, timestamp discontinuities, receivers with different
amounts of data in one run, random buffer sizes (seeded fuzz) — always against the oracle, bit for bit."""
from pathlib import Path

import numpy as np
import pytest

from oraclelib import Oracle
from paritylib import diff_bufres, diff_frames, diff_stats
from readsb_b200 import synth
from readsb_b200.abi import BUFRES_DTYPE, FRAME_DTYPE

pytestmark = pytest.mark.gpu


def _oracle_buffers(o, iq, cuts, ts_of):
    """Feed the oracle buffer by buffer exactly like sdr_ifile.c:209-241 (halo carried iff the previous buffer had >= 326 samples)."""
    halo = np.zeros(326, dtype=np.uint16)
    frames, bufres = [], []
    for b, (lo, hi) in enumerate(cuts):
        mag, sl, sp = Oracle.convert(iq[2 * lo: 2 * hi]) if hi > lo else (np.zeros(0, np.uint16), 0, 0)
        data = np.concatenate([halo, mag]).astype(np.uint16)
        f, r = o.demodulate(data, hi - lo, ts_of(b, lo), sl, sp)
        frames.append(f)
        bufres.append(np.frombuffer(bytes(r), dtype=BUFRES_DTYPE).copy())
        halo = data[hi - lo: hi - lo + 326].copy() if hi - lo >= 326 else np.zeros(326, dtype=np.uint16)
    return (np.concatenate(frames) if frames else np.zeros(0, FRAME_DTYPE)), (np.concatenate(bufres) if bufres else np.zeros(0, BUFRES_DTYPE))


def _gpu_buffers(d, iq, cuts, ts_of, per_run, stream=0):
    frames, bufres = [], []
    for i in range(0, len(cuts), per_run):
        for b in range(i, min(i + per_run, len(cuts))):
            lo, hi = cuts[b]
            d.submit_iq(stream, iq[2 * lo: 2 * hi], ts_of(b, lo))
        d.run()
        frames.append(d.frames(stream)); bufres.append(d.buffer_results(stream))
    return np.concatenate(frames), np.concatenate(bufres)


def test_ragged_and_tiny_buffers(cuda):
    from readsb_b200.demod import Demodulator
    iq = synth.mixed_stream(31, 260_000, frames_per_sec=4000)
    # full, partial, shorter than the halo, empty, full again, odd sizes (also not multiples of 8)
    sizes = [65536, 1000, 200, 0, 65536, 325, 326, 327, 12345, 65536, 7, 33333]
    cuts, pos = [], 0
    for n in sizes:
        cuts.append((pos, pos + n)); pos += n
    assert pos <= 260_000
    ts_of = lambda b, lo: lo * 5
    o = Oracle(); fo, bo = _oracle_buffers(o, iq, cuts, ts_of)
    for per_run in (1, 4):
        d = Demodulator(n_streams=1, buf_samples=65536, max_buffers_per_run=per_run)
        fg, bg = _gpu_buffers(d, iq, cuts, ts_of, per_run)
        problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + diff_stats(d.stats(0), o.stats())
        assert not problems, f"per_run={per_run}\n" + "\n".join(problems)
        d.close()
    assert len(fo) > 50


def test_timestamp_discontinuity_starts_a_new_segment(cuda):
    """Buffers whose sampleTimestamps are not contiguous (dropped samples upstream) must not be glued into one segment:
    frame timestamps come from each buffer's own sampleTimestamp (demod_2400.c:406)."""
    from readsb_b200.demod import Demodulator
    iq = synth.mixed_stream(32, 4 * 32768, frames_per_sec=5000)
    cuts = [(i * 32768, (i + 1) * 32768) for i in range(4)]
    ts_of = lambda b, lo: lo * 5 + (0, 0, 777_000, 777_000)[b]
    o = Oracle(); fo, bo = _oracle_buffers(o, iq, cuts, ts_of)
    d = Demodulator(n_streams=1, buf_samples=32768, max_buffers_per_run=4)
    fg, bg = _gpu_buffers(d, iq, cuts, ts_of, 4)
    problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + diff_stats(d.stats(0), o.stats())
    assert not problems, "\n".join(problems)
    d.close()


def test_one_segment_per_buffer_keeps_its_frames_inside_the_receivers_slots(cuda):
    """Every buffer of a run a segment of its own (a clock jump before each), eight per receiver: stage B cuts each of them
    into sub-ranges rounded up to whole scan tiles, and the frame regions the sub-ranges speculate into must still add up to
    no more than the receiver's slots — the last receiver's last regions lie at the end of the allocation (found by the
    emulator's fuzzer under AddressSanitizer; on the GPU the neighbouring receiver's frames would be overwritten)."""
    from readsb_b200.demod import Demodulator
    BUF, K, S = 37016, 8, 2
    ts_of = lambda b, lo: lo * 5 + b * 1_000_000
    cuts = [(i * BUF, (i + 1) * BUF) for i in range(K)]
    d = Demodulator(n_streams=S, buf_samples=BUF, max_buffers_per_run=K)
    want = []
    for s in range(S):
        iq = synth.mixed_stream(90 + s, K * BUF, frames_per_sec=9000)
        o = Oracle(); want.append(_oracle_buffers(o, iq, cuts, ts_of) + (o.stats(),))
        for b, (lo, hi) in enumerate(cuts):
            d.submit_iq(s, iq[2 * lo: 2 * hi], ts_of(b, lo))
    d.run()
    for s in range(S):
        fo, bo, so = want[s]
        problems = diff_frames(d.frames(s), fo) + diff_bufres(d.buffer_results(s), bo) + diff_stats(d.stats(s), so)
        assert not problems, f"receiver {s}\n" + "\n".join(problems)
        assert len(fo) > 150 and fo["j"][-1] > 33000            # the last sub-range of the last buffer holds frames
    d.close()


def test_receivers_with_unequal_work_in_one_run(cuda):
    """Some receivers idle, some with one buffer, some with several, some ending in a partial buffer."""
    from readsb_b200.demod import Demodulator
    S, BUF, K = 7, 32768, 3
    plan = [[], [BUF], [BUF, BUF, BUF], [BUF, 500], [100], [BUF, BUF], [BUF, BUF, 9999]]
    d = Demodulator(n_streams=S, buf_samples=BUF, max_buffers_per_run=K)
    iqs = [synth.mixed_stream(40 + s, K * BUF, frames_per_sec=5000) for s in range(S)]
    for rnd in range(2):                                   # two runs: halo + filter state carry per receiver
        for s in range(S):
            pos = rnd * K * BUF // 2
            for n in plan[s]:
                d.submit_iq(s, iqs[s][2 * pos: 2 * (pos + n)], pos * 5); pos += n
        d.run()
        for s in range(S):
            fr = d.frames(s)
            assert len(d.buffer_results(s)) == len(plan[s])
            if not plan[s]:
                assert len(fr) == 0
        if rnd == 0:
            first = [(d.frames(s).copy(), d.buffer_results(s).copy()) for s in range(S)]
    # compare run 0 of every receiver with a fresh oracle fed the same buffers
    for s in range(S):
        o = Oracle()
        cuts, pos = [], 0
        for n in plan[s]:
            cuts.append((pos, pos + n)); pos += n
        fo, bo = _oracle_buffers(o, iqs[s], cuts, lambda b, lo: lo * 5)
        problems = diff_frames(first[s][0], fo) + diff_bufres(first[s][1], bo)
        assert not problems, f"receiver {s}\n" + "\n".join(problems)
    d.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_buffer_sizes_fuzz(cuda, seed):
    from readsb_b200.demod import Demodulator
    rng = np.random.default_rng(seed)
    BUF = int(rng.choice([4096, 8192, 20000, 65536]))
    total = 150_000
    iq = synth.generate(total, seed=100 + seed, frames_per_sec=float(rng.choice([300, 3000, 12000])),
                        df_mask=synth.DF17 | synth.DF11 | synth.AP | synth.DF18 | synth.DF11_IID, n_icao=int(rng.integers(1, 40)),
                        amp=(0.05, 0.95), noise_sigma=float(rng.uniform(0.5, 6.0)), p_bit_error=0.3, p_two_bit_error=0.1)
    cuts, pos = [], 0
    while pos < total:
        n = int(min(total - pos, rng.integers(0, BUF + 1) if rng.random() < 0.4 else BUF))
        cuts.append((pos, pos + n)); pos += n
        if len(cuts) > 200:
            break
    thr = int(rng.choice([40, 58, 58, 90]))
    o = Oracle(thr); fo, bo = _oracle_buffers(o, iq, cuts, lambda b, lo: lo * 5)
    K = int(rng.integers(1, 6))
    d = Demodulator(n_streams=1, buf_samples=BUF, max_buffers_per_run=K, preamble_threshold=thr)
    fg, bg = _gpu_buffers(d, iq, cuts, lambda b, lo: lo * 5, K)
    problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + diff_stats(d.stats(0), o.stats())
    assert not problems, f"BUF={BUF} K={K} thr={thr} buffers={len(cuts)}\n" + "\n".join(problems)
    d.close()


@pytest.mark.parametrize("thr", [120, 200])
def test_few_survivors_in_long_runs(cuda, thr):
    """A high preamble threshold leaves so few DF-gate survivors that the pooled slices (32 at a time) wait for most of a
    run, some until its end: the ticks they are cut from must still be the run's (scan_kernel.cu, the per-warp tick copy).
    Two receivers with eight full buffers each: runs of many tiles per warp."""
    from readsb_b200.demod import Demodulator
    BUF, K = 65536, 8
    d = Demodulator(n_streams=2, buf_samples=BUF, max_buffers_per_run=K, preamble_threshold=thr)
    want = []
    for s in range(2):
        iq = synth.mixed_stream(5 + s, K * BUF, frames_per_sec=800)
        o = Oracle(thr); want.append(o.run_stream(iq, BUF) + (o.stats(),))
        for b in range(K):
            d.submit_iq(s, iq[2 * b * BUF: 2 * (b + 1) * BUF], b * BUF * 5)
    d.run()
    for s in range(2):
        fo, bo, so = want[s]
        problems = diff_frames(d.frames(s), fo) + diff_bufres(d.buffer_results(s), bo) + diff_stats(d.stats(s), so)
        assert not problems, f"receiver {s}\n" + "\n".join(problems)
        assert len(fo) > 50
    d.close()


def test_filter_flip_on_the_stream_clock(cuda):
    """Two-generation ICAO filter flipped by stream time (readsb.c:1227-1231), with a short TTL so several flips happen."""
    from readsb_b200.demod import Demodulator
    ttl = 40   # ms of stream time
    iq = synth.generate(1_200_000, seed=77, frames_per_sec=400, df_mask=synth.DF17 | synth.AP | synth.DF11, n_icao=3)
    o = Oracle(icao_ttl_ms=ttl); fo, bo = o.run_stream(iq, 32768)
    d = Demodulator(n_streams=1, buf_samples=32768, max_buffers_per_run=5, icao_ttl_ms=ttl)
    fg, bg = d.replay(iq)
    assert bo["icao_flipped"].sum() >= 8
    problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + diff_stats(d.stats(0), o.stats())
    assert not problems, "\n".join(problems)
    d.close()


@pytest.mark.parametrize("n_icao,fps,K,ttl", [(120, 3000.0, 1, 400), (400, 6000.0, 4, 400), (400, 6000.0, 8, 0), (1000, 9000.0, 3, 150)])
def test_busy_airspace_follows_the_reference_table_resize(cuda, n_icao, fps, K, ttl):
    """icao_filter.c:126-128 + :66-92: with the 86th, 171st, 342nd ... new address of a generation the reference doubles its tables
    and keeps the active generation only - address/parity replies of aircraft that only the older generation knew are rejected from
    that frame on (the oracle is pinned against the reference for this, tests/test_oracle.py).  Buffers resolved speculatively in
    parallel, in one run, across the resize points; with and without flips (ttl 0 = the reference's 60 s: none in this capture)."""
    from readsb_b200.demod import Demodulator
    iq = synth.generate(3_000_000, seed=5, frames_per_sec=fps, df_mask=synth.DF17 | synth.DF11 | synth.AP | synth.DF11_IID, n_icao=n_icao)
    kw = {"icao_ttl_ms": ttl} if ttl else {}
    o = Oracle(**kw); fo, bo = o.run_stream(iq, 65536)
    d = Demodulator(n_streams=1, buf_samples=65536, max_buffers_per_run=K, **kw)
    fg, bg = d.replay(iq)
    assert len(fo) > 1000
    problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + diff_stats(d.stats(0), o.stats())
    assert not problems, "\n".join(problems)
    d.close()


def test_filter_api_follows_the_reference_table_resize(cuda):
    """The same rule through b200_demod_icao_add / _expire (the shim forwards network-input adds and the 60 s flip): an address of the
    older generation dies with the 86th new address of the next one; a small generation shrinks the tables at the flip (:97-99)."""
    from readsb_b200.demod import Demodulator
    d = Demodulator(n_streams=1, buf_samples=4096, max_buffers_per_run=1)
    o = Oracle()
    for f in (lambda a: d.icao_add(0, a), o.icao_add):
        f(0xABCDEF)
    d.icao_expire(0); o.icao_expire()
    for a in range(1, 86):
        d.icao_add(0, a); o.icao_add(a)
    assert d.icao_test(0, 0xABCDEF) and o.icao_test(0xABCDEF)
    d.icao_add(0, 86); o.icao_add(86)
    assert not d.icao_test(0, 0xABCDEF) and not o.icao_test(0xABCDEF) and d.icao_test(0, 1) and d.icao_test(0, 86)
    rng = np.random.default_rng(3)
    pool = rng.choice(1 << 24, size=1500, replace=False)
    for step in range(6):
        for a in rng.choice(pool, size=int(rng.integers(20, 260))):
            d.icao_add(0, int(a)); o.icao_add(int(a))
        probe = rng.choice(pool, size=150)
        assert [d.icao_test(0, int(a)) for a in probe] == [o.icao_test(int(a)) for a in probe], step
        d.icao_expire(0); o.icao_expire()
    d.close()


def test_filter_grows_beyond_the_default_tables(cuda):
    """The default tables hold 2048 addresses per generation (stage B keeps them in shared memory); the reference's grow to 2^20
    buckets.  Receiver 0: 10 000 addresses forwarded through the API (an aggregator's network input), then a capture.  Receiver 1:
    its own traffic (thousands of aircraft) pushes it past the default size in the middle of a replay: the capacity check ahead
    of stage B asks for larger tables, the library grows them (icao_rehash_kernel) and repeats the step.  Receiver 2 stays small.
    Results are the oracle's for all three."""
    from readsb_b200.demod import Demodulator
    S = 3
    d = Demodulator(n_streams=S, buf_samples=65536, max_buffers_per_run=4)
    os_ = [Oracle() for _ in range(S)]
    rng = np.random.default_rng(17)
    fwd = rng.choice(1 << 24, size=10_000, replace=False)
    for a in fwd:
        d.icao_add(0, int(a)); os_[0].icao_add(int(a))
    probe = list(fwd[::97]) + list(rng.choice(1 << 24, size=100))
    assert [d.icao_test(0, int(a)) for a in probe] == [os_[0].icao_test(int(a)) for a in probe]
    iqs = [synth.generate(900_000, seed=40, frames_per_sec=2000.0, df_mask=synth.DF17 | synth.DF11 | synth.AP, n_icao=30),
           synth.generate(6_000_000, seed=41, frames_per_sec=4000.0, df_mask=synth.DF17 | synth.DF11 | synth.AP, n_icao=9000, amp=(0.5, 0.9)),
           synth.generate(600_000, seed=42, frames_per_sec=500.0, df_mask=synth.DF17 | synth.AP, n_icao=10)]
    for s in (1, 0, 2):
        fo, bo = os_[s].run_stream(iqs[s], 65536)
        fg, bg = d.replay(iqs[s], stream=s)
        problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + diff_stats(d.stats(s), os_[s].stats())
        assert not problems, f"receiver {s}\n" + "\n".join(problems)
        if s == 1:
            taught = {int(a) for a, t, c in zip(fo["addr"], fo["msgtype"], fo["correctedbits"]) if t in (11, 17) and c == 0}
            assert len(taught) > 2600, len(taught)        # more than a default table takes
            assert all(d.icao_test(1, a) == os_[1].icao_test(a) for a in list(taught)[::40])
    d.close()


def test_syndrome_all_ones_is_not_a_correctable_error(cuda):
    """A DF17 candidate whose syndrome is exactly 0xFFFFFF (tests/golden/regress/syndrome_ffffff.npz: 2200 samples cut out of a
    tools/emu_fuzz.py case, loud traffic at --preamble-threshold=33).  The single-bit-error lookup is a perfect hash whose
    empty slots once held 0xffffffff: this syndrome matched an empty slot, came back as 'bit 255' and the frame was accepted
    with score 700 (and finalize flipped msg[-1]).  The reference finds no such error (crc.c:383-406) and rejects the phase."""
    from readsb_b200.demod import Demodulator
    z = np.load(Path(__file__).resolve().parent / "golden" / "regress" / "syndrome_ffffff.npz")
    iq, thr = z["iq"], int(z["preamble_threshold"])
    o = Oracle(preamble_threshold=thr)
    fo, bo = o.run_stream(iq, 65536)
    d = Demodulator(n_streams=1, buf_samples=65536, preamble_threshold=thr)
    fg, bg = d.replay(iq)
    problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + diff_stats(d.stats(0), o.stats())
    assert not problems, "\n".join(problems)
    assert o.stats()["demod_preambles"] > 10
    d.close()


def test_modeac_with_empty_and_tiny_buffers(cuda):
    """Mode A/C next to buffers a frontend delivered empty or shorter than the halo.  An empty buffer has no tiles and no
    bit-map words: the walk once computed its last word from `x_end - 1` (wrapping for an empty range) and read the bit map out
    of bounds — found by tools/emu_fuzz.py under AddressSanitizer (seed 71, case 47)."""
    from readsb_b200.demod import Demodulator
    iq = synth.modeac_stream(45, 150_000)
    sizes = [20000, 0, 1, 7, 20000, 0, 325, 326, 327, 20000, 0, 13000, 20000]
    cuts, pos = [], 0
    for n in sizes:
        cuts.append((pos, pos + n)); pos += n
    assert pos <= 150_000
    for per_run in (1, 2, 5):
        d = Demodulator(n_streams=1, buf_samples=20000, max_buffers_per_run=per_run, mode_ac=True)
        o, oa = Oracle(), Oracle()
        fo, bo = _oracle_buffers(o, iq, cuts, lambda b, lo: lo * 5)
        fg, bg, ag, nb0 = [], [], [], 0
        for i in range(0, len(cuts), per_run):
            part = cuts[i:i + per_run]
            for lo, hi in part:
                d.submit_iq(0, iq[2 * lo: 2 * hi], lo * 5)
            d.run()
            fg.append(d.frames(0)); bg.append(d.buffer_results(0))
            a = d.modeac(0); a["buffer_idx"] += nb0; ag.append(a); nb0 += len(part)
        halo, want = np.zeros(326, np.uint16), []
        for b, (lo, hi) in enumerate(cuts):
            mag, sl, sp = Oracle.convert(iq[2 * lo: 2 * hi]) if hi > lo else (np.zeros(0, np.uint16), 0, 0)
            data = np.concatenate([halo, mag]).astype(np.uint16)
            a = oa.demodulate_ac(data, hi - lo, lo * 5, sl, sp); a["buffer_idx"] = b; want.append(a)
            halo = data[hi - lo: hi - lo + 326].copy() if hi - lo >= 326 else np.zeros(326, np.uint16)
        ao, agc = np.concatenate(want), np.concatenate(ag)
        assert len(ao) > 8 and len(agc) == len(ao), f"per_run={per_run}: {len(agc)} vs {len(ao)} replies"
        for f in ("timestamp", "f1_sample", "modeac", "buffer_idx"):
            assert np.array_equal(agc[f], ao[f]), f"per_run={per_run}: {f}"
        problems = diff_frames(np.concatenate(fg), fo) + diff_bufres(np.concatenate(bg), bo) + diff_stats(d.stats(0), o.stats())
        assert not problems, f"per_run={per_run}\n" + "\n".join(problems)
        assert d.stats(0)["demod_modeac"] == len(ao)
        d.close()


@pytest.mark.parametrize("mode_ac", [False, True])
def test_runs_with_nothing_submitted(cuda, mode_ac):
    """b200_demod_run with no buffer queued, before and after a run that had data for one of two receivers: no results, and the
    previous run's frames / replies / Beast bytes do not reappear."""
    from readsb_b200.demod import Demodulator
    iq = synth.modeac_stream(5, 40000)
    d = Demodulator(n_streams=2, buf_samples=20000, max_buffers_per_run=2, mode_ac=mode_ac)
    d.run()
    assert len(d.frames(0)) == 0 and len(d.buffer_results(1)) == 0 and d.beast(0) == b"" and (not mode_ac or len(d.modeac(0)) == 0)
    d.submit_iq(1, iq[:40000], 0)
    d.submit_iq(1, iq[40000:], 20000 * 5)
    d.run()
    fo, _ = Oracle().run_stream(iq, 20000)
    assert len(fo) > 5 and not diff_frames(d.frames(1), fo) and len(d.frames(0)) == 0 and len(d.beast(1)) > 0
    d.run()
    assert len(d.frames(1)) == 0 and d.beast(1) == b"" and (not mode_ac or len(d.modeac(1)) == 0)
    assert d.stats(1)["demod_accepted"][0] + d.stats(1)["demod_accepted"][1] == len(fo)
    d.close()
