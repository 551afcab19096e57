"""-m gpu, needs two devices: ONE host process drives several GPUs, one context per device (b200_demod_config.device), each from
its own thread — north_star's "independent sample buffers shard across the 8 B200s (one cudaStream per receiver batch, no NCCL)".
Kernel attributes (dynamic shared memory opt-in) are per device, and every context sets them for its own device at create."""
import threading

import numpy as np
import pytest

from oraclelib import Oracle
from paritylib import diff_bufres, diff_frames, diff_stats
from readsb_b200 import synth

pytestmark = pytest.mark.gpu


def _n_devices() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_devices() < 2, reason="needs two CUDA devices in one process")
@pytest.mark.parametrize("streams", [1, 3])
def test_two_devices_from_one_process(cuda, streams):
    from readsb_b200.demod import Demodulator
    ndev = min(_n_devices(), 4)
    n = 5 * 65536 + 777
    iqs = [[synth.mixed_stream(300 + 10 * g + s, n) for s in range(streams)] for g in range(ndev)]
    ds = [Demodulator(n_streams=streams, buf_samples=65536, max_buffers_per_run=2, device=g) for g in range(ndev)]
    got = [[None] * streams for _ in range(ndev)]
    errors = []

    def work(g):
        try:
            for rep in range(2):                      # the second pass interleaves the devices' runs some more
                for s in range(streams):
                    if rep == 0:
                        got[g][s] = ds[g].replay(iqs[g][s], stream=s)
        except Exception as e:                       # noqa: BLE001 - reported below
            errors.append((g, repr(e)))
    threads = [threading.Thread(target=work, args=(g,)) for g in range(ndev)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for g in range(ndev):
        for s in range(streams):
            o = Oracle()
            fo, bo = o.run_stream(iqs[g][s], 65536)
            fg, bg = got[g][s]
            problems = diff_frames(fg, fo) + diff_bufres(bg, bo) + diff_stats(ds[g].stats(s), o.stats())
            assert len(fo) > 100 and not problems, f"device {g} receiver {s}\n" + "\n".join(problems)
    for d in ds:
        d.close()


@pytest.mark.skipif(_n_devices() < 2, reason="needs two CUDA devices in one process")
def test_device_resident_batches_on_two_devices(cuda):
    """The bench's shape (device-resident batches, pipelined) on two devices from one process, interleaved step by step."""
    import torch
    from readsb_b200.demod import Demodulator
    S, B, BUF = 8, 2, 65536
    ctx = []
    for g in range(2):
        iqs = [synth.config2_stream(500 + 20 * g + s, 2 * B * BUF) for s in range(S)]
        stride = 2 * B * BUF * 2
        with torch.cuda.device(g):
            dev = torch.zeros(4096 + S * stride + 256, dtype=torch.uint8, device=f"cuda:{g}")
            for s in range(S):
                dev[4096 + s * stride: 4096 + (s + 1) * stride] = torch.from_numpy(iqs[s]).to(f"cuda:{g}")
            torch.cuda.synchronize(g)
        ctx.append((Demodulator(n_streams=S, buf_samples=BUF, max_buffers_per_run=B, device=g), dev, stride, iqs))
    frames = [[[] for _ in range(S)] for _ in range(2)]
    for k in range(2):
        for g, (d, dev, stride, _) in enumerate(ctx):
            d.run_device_async(dev.data_ptr() + 4096 + k * B * BUF * 2, stride, B, BUF, continues=k > 0, first_sample_timestamp=k * B * BUF * 5)
        for g, (d, _, _, _) in enumerate(ctx):
            d.wait()
            for s in range(S):
                frames[g][s].append(d.frames(s))
    for g, (d, _, _, iqs) in enumerate(ctx):
        for s in range(S):
            fo, _ = Oracle().run_stream(iqs[s], BUF)
            assert not diff_frames(np.concatenate(frames[g][s]), fo), (g, s)
        d.close()
