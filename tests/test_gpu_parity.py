"""-m gpu: the CUDA path, called through the C ABI, must equal the CPU oracle bit for bit.

Sizes here are chosen so the oracle finishes in seconds; BASELINE-size runs are covered through
size-independent properties in test_gpu_fullsize.py.
"""
import numpy as np
import pytest

from oraclelib import Oracle
from paritylib import diff_bufres, diff_frames, diff_stats
from readsb_b200 import synth

pytestmark = pytest.mark.gpu

GENS = {"cfg2": synth.config2_stream, "cfg5": synth.config5_stream, "mixed": synth.mixed_stream}


def _check(d, o, fg, bg, fo, bo, stream=0):
    problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + diff_stats(d.stats(stream), o.stats())
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("kind", ["cfg2", "cfg5", "mixed"])
@pytest.mark.parametrize("buf,K", [(65536, 1), (65536, 4), (131072, 2)])
def test_replay_matches_oracle(cuda, kind, buf, K):
    from readsb_b200.demod import Demodulator
    iq = GENS[kind](11, 1_000_000)          # ends in a partial buffer
    o = Oracle()
    fo, bo = o.run_stream(iq, buf)
    d = Demodulator(n_streams=1, buf_samples=buf, max_buffers_per_run=K)
    fg, bg = d.replay(iq)
    assert len(fo) > 10
    _check(d, o, fg, bg, fo, bo)
    d.close()


@pytest.mark.parametrize("thr,nfix,fixdf", [(58, 0, 1), (58, 1, 0), (40, 1, 1), (120, 1, 1), (58, 0, 0)])
def test_option_variants(cuda, thr, nfix, fixdf):
    """--preamble-threshold, --no-fix, --no-fix-df (readsb.c:1460-1473)."""
    from readsb_b200.demod import Demodulator
    iq = synth.mixed_stream(5, 700_000)
    o = Oracle(thr, nfix, fixdf)
    fo, bo = o.run_stream(iq, 65536)
    d = Demodulator(n_streams=1, buf_samples=65536, max_buffers_per_run=3, preamble_threshold=thr, nfix_crc=nfix, fix_df=fixdf)
    fg, bg = d.replay(iq)
    _check(d, o, fg, bg, fo, bo)
    d.close()


def test_many_streams_independent(cuda):
    """Receivers are independent: own halo, own ICAO filter, own skip state."""
    from readsb_b200.demod import Demodulator
    S, buf, nb = 6, 65536, 3
    kinds = ["cfg2", "cfg5", "mixed"]
    iqs = [GENS[kinds[s % 3]](100 + s, buf * nb) for s in range(S)]
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=nb)
    for b in range(nb):
        for s in range(S):
            d.submit_iq(s, iqs[s][2 * b * buf: 2 * (b + 1) * buf], b * buf * 5)
    d.run()
    for s in range(S):
        o = Oracle()
        fo, bo = o.run_stream(iqs[s], buf)
        _check(d, o, d.frames(s), d.buffer_results(s), fo, bo, stream=s)
    d.close()


def test_magnitude_handoff(cuda):
    """demodulate2400(struct mag_buf*) boundary: uint16 magnitudes with their 326-sample halo."""
    from readsb_b200.demod import Demodulator
    buf = 65536
    iq = synth.mixed_stream(21, 3 * buf + 1234)
    mag, _, _ = Oracle.convert(iq)
    o = Oracle()
    d = Demodulator(n_streams=1, buf_samples=buf, max_buffers_per_run=2)
    halo = np.zeros(326, dtype=np.uint16)
    fg_all, fo_all = [], []
    off = 0
    n = len(mag)
    pending = []
    while off < n:
        ln = min(buf, n - off)
        data = np.concatenate([halo, mag[off:off + ln]])
        sl, sp = int(mag[off:off + ln].astype(np.uint64).sum()), int((mag[off:off + ln].astype(np.uint64) ** 2).sum())
        fo, ro = o.demodulate(data, ln, off * 5, sl, sp)
        fo_all.append(fo)
        d.submit_mag(0, data, ln, off * 5)
        pending.append((sl, sp, ln))
        if len(pending) == 2 or off + ln >= n:
            d.run()
            fg_all.append(d.frames(0))
            br = d.buffer_results(0)
            assert [int(x) for x in br["sum_level"]] == [p[0] for p in pending]
            assert [int(x) for x in br["sum_power"]] == [p[1] for p in pending]
            pending = []
        if ln >= 326:
            halo = data[ln:ln + 326].copy()
        else:
            halo = np.zeros(326, dtype=np.uint16)
        off += ln
    fg, fo = np.concatenate(fg_all), np.concatenate(fo_all)
    problems = diff_frames(fg, fo) + diff_stats(d.stats(0), o.stats())
    assert not problems, "\n".join(problems)
    d.close()


def test_device_resident_batches(cuda):
    """Inputs already in HBM: streams contiguous in device memory, several buffers per call, continuation."""
    import devbuf
    from readsb_b200.demod import Demodulator
    S, buf, nb, calls = 5, 65536, 2, 3
    total = buf * nb * calls
    iqs = [GENS[["cfg5", "mixed", "cfg2"][s % 3]](300 + s, total) for s in range(S)]
    pad = 1024  # room for the 326-sample halo in front of stream 0
    stride = total * 2 + 4096
    dev = devbuf.zeros(pad + S * stride)
    for s in range(S):
        dev[pad + s * stride: pad + s * stride + 2 * total] = devbuf.to_dev(iqs[s])
    devbuf.sync()
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=nb)
    got = [[] for _ in range(S)]
    gotb = [[] for _ in range(S)]
    for c in range(calls):
        base = dev.data_ptr() + pad + c * nb * buf * 2
        d.run_device(base, stride, nb, buf, continues=c > 0, first_sample_timestamp=c * nb * buf * 5)
        for s in range(S):
            got[s].append(d.frames(s))
            gotb[s].append(d.buffer_results(s))
    for s in range(S):
        o = Oracle()
        fo, bo = o.run_stream(iqs[s], buf)
        _check(d, o, np.concatenate(got[s]), np.concatenate(gotb[s]), fo, bo, stream=s)
    d.close()


def test_icao_filter_api(cuda):
    """icaoFilterAdd/Test/Expire semantics: an address lives until the second flip after its add."""
    from readsb_b200.demod import Demodulator
    d = Demodulator(n_streams=2, buf_samples=4096, max_buffers_per_run=1)
    assert not d.icao_test(0, 0x4840D6)
    d.icao_add(0, 0x4840D6)
    assert d.icao_test(0, 0x4840D6) and not d.icao_test(1, 0x4840D6)
    d.icao_expire(0)
    assert d.icao_test(0, 0x4840D6)
    d.icao_expire(0)
    assert not d.icao_test(0, 0x4840D6)
    d.close()


def test_seeded_filter_accepts_address_parity_replies(cuda):
    """DF0/4/5/16/20/21 are only accepted from known aircraft (mode_s.c:343-360): seeding the filter through the
    API must have the same effect as the oracle's."""
    from readsb_b200.demod import Demodulator
    iq = synth.generate(400_000, seed=9, frames_per_sec=3000, df_mask=synth.AP, n_icao=4)
    _, truth = synth.generate(400_000, seed=9, frames_per_sec=3000, df_mask=synth.AP, n_icao=4, want_truth=True)
    addrs = {Oracle.crc24(m) for _, m, _ in truth}
    assert 1 <= len(addrs) <= 4
    o = Oracle(); d = Demodulator(n_streams=1, buf_samples=131072, max_buffers_per_run=1)
    for a in addrs:
        o.icao_add(a); d.icao_add(0, a)
    fo, bo = o.run_stream(iq, 131072)
    fg, bg = d.replay(iq)
    assert len(fo) > 100
    _check(d, o, fg, bg, fo, bo)
    d.close()


def _device_streams(iqs, total, pad=1024):
    import devbuf
    stride = total * 2 + 4096
    dev = devbuf.zeros(pad + len(iqs) * stride)
    for s, iq in enumerate(iqs):
        dev[pad + s * stride: pad + s * stride + 2 * total] = devbuf.to_dev(iq)
    devbuf.sync()
    return dev, stride, pad


@pytest.mark.parametrize("depth", [2, 3])
def test_async_pipeline_matches_oracle(cuda, depth):
    """run_device_uc8_async / wait: two or three steps in flight (scans back to back, stage B of each step behind its scan);
    a fourth step is refused until the oldest one has been waited for."""
    from readsb_b200.demod import DemodError, Demodulator
    S, buf, nb, calls = 4, 65536, 2, 6
    total = buf * nb * calls
    iqs = [GENS[["cfg5", "mixed", "cfg2", "mixed"][s]](700 + s, total) for s in range(S)]
    dev, stride, pad = _device_streams(iqs, total)
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=nb)
    got = [[] for _ in range(S)]; gotb = [[] for _ in range(S)]

    def collect():
        d.wait()
        for s in range(S):
            got[s].append(d.frames(s)); gotb[s].append(d.buffer_results(s))
    flying = 0
    for c in range(calls):
        d.run_device_async(dev.data_ptr() + pad + c * nb * buf * 2, stride, nb, buf, continues=c > 0, first_sample_timestamp=c * nb * buf * 5)
        flying += 1
        if depth == 3 and flying == 3:
            with pytest.raises(DemodError):
                d.run_device_async(dev.data_ptr() + pad, stride, nb, buf, continues=False, first_sample_timestamp=0)
        if flying == depth:
            collect(); flying -= 1
    while flying:
        collect(); flying -= 1
    with pytest.raises(DemodError):
        d.wait()
    for s in range(S):
        o = Oracle()
        fo, bo = o.run_stream(iqs[s], buf)
        _check(d, o, np.concatenate(got[s]), np.concatenate(gotb[s]), fo, bo, stream=s)
    d.close()


def test_long_pipelined_session_stays_exact_while_the_sm_partition_is_measured(cuda):
    """A pipelined session measures how many SMs the scan should take (whole chip / 85 % / 82 %, demod_api.cu tune_partition):
    the scan grid changes between steps during the first ~60 of them, and once more when the size of the runs changes.  Grid
    size is a launch parameter only - every step's frames must be the oracle's.  90 steps of 8192 samples, then 20 of two buffers."""
    from readsb_b200.demod import Demodulator
    S, buf = 3, 8192
    calls_a, calls_b = 90, 20
    total = buf * (calls_a + 2 * calls_b)
    iqs = [synth.mixed_stream(4100 + s, total, frames_per_sec=3000) for s in range(S)]
    dev, stride, pad = _device_streams(iqs, total)
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=2)
    got = [[] for _ in range(S)]; gotb = [[] for _ in range(S)]

    def collect():
        d.wait()
        for s in range(S):
            got[s].append(d.frames(s)); gotb[s].append(d.buffer_results(s))
    flying, off = 0, 0
    for c in range(calls_a + calls_b):
        nb = 1 if c < calls_a else 2
        d.run_device_async(dev.data_ptr() + pad + off * 2, stride, nb, buf, continues=c > 0, first_sample_timestamp=off * 5)
        off += nb * buf
        flying += 1
        if flying == 3:
            collect(); flying -= 1
    while flying:
        collect(); flying -= 1
    for s in range(S):
        o = Oracle()
        fo, bo = o.run_stream(iqs[s], buf)
        _check(d, o, np.concatenate(got[s]), np.concatenate(gotb[s]), fo, bo, stream=s)
    d.close()


def test_host_async_pipeline_matches_oracle(cuda):
    """run_host_uc8_async / wait: pinned host slabs, the copy of step n+1 overlapping the kernels of step n; halo carried
    between the two library-owned input buffers; unpinned memory and a restart (continues=0) in the middle."""
    from readsb_b200.demod import DemodError, Demodulator, PinnedBuffer
    S, buf, nb, calls = 3, 32768, 2, 6
    total = buf * nb * calls
    iqs = [GENS[["cfg5", "mixed", "cfg2"][s]](800 + s, total) for s in range(S)]
    pin = PinnedBuffer(S * 2 * total)
    host = pin.array.reshape(S, 2 * total)
    for s in range(S):
        host[s] = iqs[s]
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=nb)
    with pytest.raises(DemodError):      # nothing to continue from yet
        d.run_host_async(pin.ptr, 2 * total, nb, buf, True, 0)
    got = [[] for _ in range(S)]; gotb = [[] for _ in range(S)]

    def collect():
        d.wait()
        for s in range(S):
            got[s].append(d.frames(s)); gotb[s].append(d.buffer_results(s))
    for c in range(calls):
        d.run_host_async(pin.ptr + c * nb * buf * 2, 2 * total, nb, buf, c > 0, c * nb * buf * 5)
        if c >= 1:
            collect()
    collect()
    for s in range(S):
        o = Oracle()
        fo, bo = o.run_stream(iqs[s], buf)
        _check(d, o, np.concatenate(got[s]), np.concatenate(gotb[s]), fo, bo, stream=s)
    d.close()
    # ordinary (pageable) memory, a partial last step, and a receiver restart: three steps, the third starts over
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=nb)
    plain = np.ascontiguousarray(host.copy())
    got = [[] for _ in range(S)]; gotb = [[] for _ in range(S)]
    d.run_host_async(plain.ctypes.data, 2 * total, nb, buf, False, 0)
    d.run_host_async(plain.ctypes.data + nb * buf * 2, 2 * total, 1, buf - 4096, True, nb * buf * 5)
    collect(); collect()
    restart_ts = (nb * buf + buf - 4096) * 5 + 12_000_000     # the receiver comes back a second later
    d.run_host_async(plain.ctypes.data, 2 * total, nb, buf, False, restart_ts)
    collect()
    for s in range(S):
        o = Oracle()
        fo1, bo1 = o.run_stream(iqs[s][: 2 * (nb * buf + buf - 4096)], buf)
        o.restart_stream()
        fo2, bo2 = o.run_stream(iqs[s][: 2 * nb * buf], buf, first_ts=restart_ts)
        _check(d, o, np.concatenate(got[s]), np.concatenate(gotb[s]), np.concatenate([fo1, fo2]), np.concatenate([bo1, bo2]), stream=s)
    d.close()
    pin.free()


def test_async_pipeline_repeats_steps_exactly_after_a_pool_failure(cuda):
    """A dense capture makes the FIRST pipelined step ask for the scratch arena while the second is already in flight:
    both must be repeated in order and stay bit-exact (stage B of the second must not have run on stale state)."""
    from readsb_b200.demod import Demodulator
    buf, nb, calls = 32768, 1, 5
    total = buf * nb * calls
    storm = synth.generate(total, seed=8, frames_per_sec=40000, df_mask=synth.DF17 | synth.DF11 | synth.AP, n_icao=4,
                           amp=(0.3, 0.9), p_bit_error=0.3)
    dev, stride, pad = _device_streams([storm], total)
    o = Oracle(40); fo, bo = o.run_stream(storm, buf)
    d = Demodulator(n_streams=1, buf_samples=buf, max_buffers_per_run=nb, preamble_threshold=40, mode_ac=True)
    got, gotb, gota = [], [], []
    for c in range(calls):
        d.run_device_async(dev.data_ptr() + pad + c * nb * buf * 2, stride, nb, buf, continues=c > 0, first_sample_timestamp=c * nb * buf * 5)
        if c >= 2:       # three steps in flight when the first one reports its failure: all three are repeated in order
            d.wait(); got.append(d.frames(0)); gotb.append(d.buffer_results(0)); gota.append(d.modeac(0))
    for _ in range(2):
        d.wait(); got.append(d.frames(0)); gotb.append(d.buffer_results(0)); gota.append(d.modeac(0))
    _check(d, o, np.concatenate(got), np.concatenate(gotb), fo, bo)
    ao = Oracle(40).run_stream_ac(storm, buf)          # the repeated steps must not count their Mode A/C replies twice
    ag = np.concatenate(gota)
    assert len(ag) == len(ao) and np.array_equal(ag["timestamp"], ao["timestamp"]) and d.stats(0)["demod_modeac"] == len(ao)
    d.close()


def _diff_modeac(a, b):
    if len(a) != len(b):
        return [f"modeac count {len(a)} vs {len(b)}"]
    return [f"modeac {f} differs at {np.nonzero(a[f] != b[f])[0][:3]}" for f in ("timestamp", "f1_sample", "modeac", "buffer_idx") if not np.array_equal(a[f], b[f])]


@pytest.mark.parametrize("buf,K", [(65536, 1), (65536, 3), (20000, 4), (1000, 16), (8191, 5)])
def test_modeac_matches_oracle(cuda, buf, K):
    """--modeac: demodulate2400AC (demod_2400.c:575-761) on the GPU next to the Mode S path, same buffers."""
    from readsb_b200.demod import Demodulator
    iq = synth.modeac_stream(17, 900_000)
    o = Oracle(); fo, bo = o.run_stream(iq, buf); ao = o.run_stream_ac(iq, buf)
    d = Demodulator(n_streams=1, buf_samples=buf, max_buffers_per_run=K, mode_ac=True)
    fg, bg, ag = d.replay(iq, want_modeac=True)
    assert len(ao) > 50
    problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + _diff_modeac(ag, ao)
    st = d.stats(0)
    assert st["demod_modeac"] == len(ao)
    assert not problems, "\n".join(problems)
    d.close()


def test_modeac_on_magnitude_handoff_and_many_streams(cuda):
    from readsb_b200.demod import Demodulator
    S, buf = 3, 32768
    iqs = [synth.generate(3 * buf, seed=60 + s, frames_per_sec=3000.0, df_mask=synth.MODEAC, amp=(0.5, 0.95)) for s in range(S)]
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=3, mode_ac=True)
    halos = [np.zeros(326, np.uint16) for _ in range(S)]
    for b in range(3):
        for s in range(S):
            mag, _, _ = Oracle.convert(iqs[s][2 * b * buf: 2 * (b + 1) * buf])
            data = np.concatenate([halos[s], mag]).astype(np.uint16)
            d.submit_mag(s, data, buf, b * buf * 5)
            halos[s] = data[buf: buf + 326].copy()
    d.run()
    for s in range(S):
        ao = Oracle().run_stream_ac(iqs[s], buf)
        assert len(ao) > 5
        assert not _diff_modeac(d.modeac(s), ao), f"stream {s}"
    d.close()


def test_modeac_device_path_async_pipeline(cuda):
    """Mode A/C through run_device_uc8_async: two steps in flight, replies and demod_modeac identical to the oracle."""
    import devbuf
    from readsb_b200.demod import Demodulator
    S, B, BUF, steps = 4, 2, 32768, 3
    n = B * BUF * steps
    iqs = [synth.generate(n, seed=80 + s, frames_per_sec=2500.0, df_mask=synth.MODEAC | synth.DF17, n_icao=4, amp=(0.4, 0.95)) for s in range(S)]
    pad, stride = 1024, 2 * n + 4096
    dev = devbuf.zeros(pad + S * stride)
    for s in range(S):
        dev[pad + s * stride: pad + s * stride + 2 * n] = devbuf.to_dev(iqs[s])
    devbuf.sync()
    d = Demodulator(n_streams=S, buf_samples=BUF, max_buffers_per_run=B, mode_ac=True)
    got_f = [[] for _ in range(S)]; got_a = [[] for _ in range(S)]

    def harvest(k):
        d.wait()
        for s in range(S):
            got_f[s].append(d.frames(s))
            a = d.modeac(s); a["buffer_idx"] += k * B; got_a[s].append(a)

    for k in range(steps):
        d.run_device_async(dev.data_ptr() + pad + k * B * BUF * 2, stride, B, BUF, continues=k > 0, first_sample_timestamp=k * B * BUF * 5)
        if k >= 1:
            harvest(k - 1)
    harvest(steps - 1)
    for s in range(S):
        o = Oracle()
        fo, _ = o.run_stream(iqs[s], BUF)
        ao = o.run_stream_ac(iqs[s], BUF) if False else Oracle().run_stream_ac(iqs[s], BUF)
        assert len(ao) > 10
        problems = diff_frames(np.concatenate(got_f[s]), fo) + _diff_modeac(np.concatenate(got_a[s]), ao)
        assert not problems, f"stream {s}: " + "\n".join(problems)
        assert d.stats(s)["demod_modeac"] == len(ao)
    d.close()


@pytest.mark.parametrize("verbatim", [False, True])
def test_beast_output_matches_oracle(cuda, verbatim):
    """b200_demod_fetch_beast (device-side modesSendBeastOutput, net_io.c:1655-1714): Mode S + Mode A/C records in the
    reference's order, byte for byte; several streams, several buffers per run, runs concatenated."""
    from readsb_b200.demod import Demodulator
    S, buf, K, runs = 3, 32768, 3, 2
    n = buf * K * runs
    iqs = [synth.generate(n, seed=90 + s, frames_per_sec=[3000.0, 9000.0, 500.0][s], df_mask=synth.MODEAC | synth.DF17 | synth.DF11 | synth.AP | synth.DF18,
                          n_icao=5, amp=(0.35, 0.95), p_bit_error=0.4) for s in range(S)]
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=K, mode_ac=True)
    got = [b"" for _ in range(S)]
    for r in range(runs):
        for b in range(K):
            k = r * K + b
            for s in range(S):
                d.submit_iq(s, iqs[s][2 * k * buf: 2 * (k + 1) * buf], k * buf * 5)
        d.run()
        for s in range(S):
            got[s] += d.beast(s, verbatim=verbatim)
            assert d.beast(s, verbatim=verbatim) == d.beast(s, verbatim=verbatim)       # cached second call
    for s in range(S):
        o = Oracle(); fo, _ = o.run_stream(iqs[s], buf); ao = Oracle().run_stream_ac(iqs[s], buf)
        want = Oracle.beast(fo, ao, verbatim=verbatim)
        assert len(fo) > 10 and (s != 0 or len(ao) > 3)
        assert got[s] == want, f"stream {s}: {len(got[s])} vs {len(want)} bytes"
    d.close()


def test_beast_golden_stream_from_reference_program(cuda):
    """The bytes the reference program sent to a beast_out client (tests/golden/beast_stream.npz) == the library's."""
    import json
    from pathlib import Path
    from readsb_b200.demod import Demodulator
    z = np.load(Path(__file__).parent / "golden" / "beast_stream.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    iq = np.concatenate([np.full(2 * meta["silence_samples"], 127, np.uint8), z["traffic"]])
    d = Demodulator(n_streams=1, buf_samples=131072, max_buffers_per_run=4, mode_ac=True)
    got, off, nsamples = b"", 0, iq.size // 2
    while off < nsamples:
        for _ in range(4):
            if off >= nsamples:
                break
            m = min(131072, nsamples - off)
            d.submit_iq(0, iq[2 * off: 2 * (off + m)], off * 5); off += m
        d.run(); got += d.beast(0, verbatim=True)
    assert got == bytes(z["beast"])
    d.close()


def test_beast_without_modeac_and_async(cuda):
    from readsb_b200.demod import Demodulator
    S, buf, nb, calls = 2, 32768, 2, 3
    total = buf * nb * calls
    iqs = [GENS["mixed"](40 + s, total) for s in range(S)]
    dev, stride, pad = _device_streams(iqs, total)
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=nb)
    got = [b"" for _ in range(S)]
    for c in range(calls):
        d.run_device_async(dev.data_ptr() + pad + c * nb * buf * 2, stride, nb, buf, continues=c > 0, first_sample_timestamp=c * nb * buf * 5)
        if c >= 1:
            d.wait()
            for s in range(S): got[s] += d.beast(s)        # the next step is in flight in the other slot
    d.wait()
    for s in range(S): got[s] += d.beast(s)
    for s in range(S):
        fo, _ = Oracle().run_stream(iqs[s], buf)
        assert len(fo) > 20 and got[s] == Oracle.beast(fo)
    d.close()


def _to_sc16(iq8: np.ndarray, q11: bool, seed: int) -> np.ndarray:
    """The uc8 synthetic capture as a 16-bit frontend would deliver it (plus a little sub-LSB dither)."""
    rng = np.random.default_rng(seed)
    scale = 16 if q11 else 256
    v = (iq8.astype(np.int32) - 128) * scale + rng.integers(0, scale, size=iq8.size)
    return np.clip(v, -2048 if q11 else -32768, 2047 if q11 else 32767).astype(np.int16)


@pytest.mark.parametrize("q11", [False, True])
def test_sc16_input_matches_oracle(cuda, q11):
    """b200_demod_submit_iq_sc16: the float-path converters on the GPU (magnitudes bit for bit -> identical frames) and
    the reference's sequential float accumulators (bit patterns in the buffer results)."""
    from readsb_b200.demod import Demodulator
    S, buf, K = 2, 32768, 2
    n = 3 * buf + 12345
    iq16 = [_to_sc16(GENS["mixed"](120 + s, n), q11, s) for s in range(S)]
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=K)
    got_f = [[] for _ in range(S)]; got_b = [[] for _ in range(S)]
    off = 0
    while off < n:
        for _ in range(K):
            if off >= n:
                break
            m = min(buf, n - off)
            for s in range(S):
                d.submit_iq_sc16(s, iq16[s][2 * off: 2 * (off + m)], off * 5, q11)
            off += m
        d.run()
        for s in range(S):
            got_f[s].append(d.frames(s)); got_b[s].append(d.buffer_results(s))
    for s in range(S):
        fo, sums = Oracle().run_stream_sc16(iq16[s], buf, q11)
        fg = np.concatenate(got_f[s]); bg = np.concatenate(got_b[s])
        assert len(fo) > 30
        problems = diff_frames(fg, fo)
        assert not problems, "\n".join(problems)
        assert len(bg) == len(sums)
        for r, (m, sl, sp) in zip(bg, sums):
            assert r["length"] == m
            assert np.uint32(r["sum_level"]).view(np.float32) == sl and np.uint32(r["sum_power"]).view(np.float32) == sp
    d.close()


@pytest.mark.parametrize("q11", [False, True])
def test_sc16_input_with_modeac(cuda, q11):
    """An sc16 frontend with --modeac: the Mode A/C noise floor comes from the converter's float-accumulated means
    (convert.c:243-249 -> demod_2400.c:580-581), the Mode S frames from the same magnitudes."""
    from readsb_b200.demod import Demodulator
    buf, K = 32768, 3
    n = 7 * buf + 4321
    iq8 = synth.generate(n, seed=91 + q11, frames_per_sec=2500.0, df_mask=synth.MODEAC | synth.DF17 | synth.DF11, n_icao=6, amp=(0.35, 0.95))
    iq16 = _to_sc16(iq8, q11, 5)
    d = Demodulator(n_streams=1, buf_samples=buf, max_buffers_per_run=K, mode_ac=True)
    got_f, got_a, off, b0 = [], [], 0, 0
    while off < n:
        k = 0
        while k < K and off < n:
            m = min(buf, n - off)
            d.submit_iq_sc16(0, iq16[2 * off: 2 * (off + m)], off * 5, q11)
            off += m; k += 1
        d.run()
        got_f.append(d.frames(0))
        a = d.modeac(0); a["buffer_idx"] += b0; got_a.append(a); b0 += k
    o = Oracle()
    fo, _ = o.run_stream_sc16(iq16, buf, q11)
    ao = Oracle().run_stream_ac_sc16(iq16, buf, q11)
    assert len(fo) > 30 and len(ao) > 15
    problems = diff_frames(np.concatenate(got_f), fo) + _diff_modeac(np.concatenate(got_a), ao)
    assert not problems, "\n".join(problems)
    assert d.stats(0)["demod_modeac"] == len(ao)
    d.close()


def test_modeac_noise_floor_from_the_mag_bufs_own_levels(cuda):
    """b200_demod_submit_mag_u16_levels: demodulate2400AC reads mag_buf.mean_level / mean_power (demod_2400.c:580-581).  Stream 0
    gets the means its (sc16) converter returned, stream 1 deliberately wrong ones (a noise floor far too high: replies vanish),
    stream 2 none (the library's own exact sums)."""
    from readsb_b200.demod import Demodulator
    buf, nb = 20000, 5
    iq8 = synth.generate(nb * buf, seed=77, frames_per_sec=3000.0, df_mask=synth.MODEAC, amp=(0.3, 0.9))
    iq16 = _to_sc16(iq8, False, 3)
    d = Demodulator(n_streams=3, buf_samples=buf, max_buffers_per_run=nb, mode_ac=True)
    o = [Oracle(), Oracle(), Oracle()]
    want = [[], [], []]
    halo = np.zeros(326, np.uint16)
    for b in range(nb):
        mag, sl, sp = Oracle.convert_sc16(iq16[2 * b * buf: 2 * (b + 1) * buf])
        data = np.concatenate([halo, mag]).astype(np.uint16)
        ml, mp = float(np.float32(sl) / np.float32(buf)), float(np.float32(sp) / np.float32(buf))
        levels = [(ml, mp), (0.3, 0.2), None]
        for s in range(3):
            if levels[s] is None:
                d.submit_mag(s, data, buf, b * buf * 5)
                a = o[s].demodulate_ac(data, buf, b * buf * 5, int(mag.astype(np.uint64).sum()), int((mag.astype(np.uint64) ** 2).sum()))
            else:
                d.submit_mag(s, data, buf, b * buf * 5, *levels[s])
                a = o[s].demodulate_ac_levels(data, buf, b * buf * 5, *levels[s])
            a["buffer_idx"] = b
            want[s].append(a)
        halo = data[buf: buf + 326].copy()
    d.run()
    n_found = []
    for s in range(3):
        w = np.concatenate(want[s])
        assert not _diff_modeac(d.modeac(s), w), f"stream {s}"
        n_found.append(len(w))
    assert n_found[0] > 40 and n_found[2] > 40 and n_found[1] < n_found[0] // 4
    d.close()


def test_preamble_threshold_changes_between_runs(cuda):
    """b200_demod_set_preamble_threshold: the reference re-reads Modes.preambleThreshold for every buffer and uses at least 75
    while samples were dropped recently (demod_2400.c:334-338)."""
    from readsb_b200.demod import Demodulator
    buf = 65536
    iq = GENS["mixed"](33, 6 * buf)
    d = Demodulator(n_streams=1, buf_samples=buf, max_buffers_per_run=2)
    o = Oracle()
    halo = np.zeros(326, np.uint16)
    fg, fo = [], []
    for step, thr in enumerate([58, 75, 58]):
        d.set_preamble_threshold(thr)
        o.set_preamble_threshold(thr)
        for b in (2 * step, 2 * step + 1):
            d.submit_iq(0, iq[2 * b * buf: 2 * (b + 1) * buf], b * buf * 5)
        d.run()
        fg.append(d.frames(0))
        fo.append(o.run_stream(iq[2 * 2 * step * buf: 2 * 2 * (step + 1) * buf], buf, first_ts=2 * step * buf * 5)[0])
    problems = diff_frames(np.concatenate(fg), np.concatenate(fo)) + diff_stats(d.stats(0), o.stats())
    assert not problems, "\n".join(problems)
    d.close()


def test_modeac_when_the_run_has_to_be_repeated(cuda):
    """A run whose stage A outgrows the record pool is repeated after regrowth; the Mode A/C walk skips the failed attempt, and
    the pack step must not touch that attempt's (absent) reply counts.  Found by tools/emu_fuzz.py (seed 7, case 12) under
    AddressSanitizer: loud overlapping traffic at --preamble-threshold=33, five receivers, device-resident path."""
    import devbuf
    from readsb_b200.demod import Demodulator
    S, buf, K = 5, 65536, 4
    total = K * buf
    iqs = [synth.generate(total, seed=12000 + s, frames_per_sec=8000.0, df_mask=synth.DF17 | synth.DF11 | synth.AP | synth.MODEAC, n_icao=8,
                          amp=(0.7, 1.0), p_bit_error=0.3) for s in range(S)]
    pad, stride = 1024, 2 * total + 4096
    dev = devbuf.zeros(pad + S * stride)
    for s in range(S):
        dev[pad + s * stride: pad + s * stride + 2 * total] = devbuf.to_dev(iqs[s])
    devbuf.sync()
    d = Demodulator(n_streams=S, buf_samples=buf, max_buffers_per_run=K, preamble_threshold=33, mode_ac=True)
    d.run_device(dev.data_ptr() + pad, stride, K, buf, continues=False, first_sample_timestamp=0)
    for s in range(S):
        o = Oracle(preamble_threshold=33)
        fo, bo = o.run_stream(iqs[s], buf)
        ao = Oracle().run_stream_ac(iqs[s], buf)
        problems = diff_frames(d.frames(s), fo) + diff_bufres(d.buffer_results(s), bo) + _diff_modeac(d.modeac(s), ao) + diff_stats(d.stats(s), o.stats())
        assert not problems, f"stream {s}: " + "\n".join(problems)
        assert d.stats(s)["demod_modeac"] == len(ao)
    d.close()
