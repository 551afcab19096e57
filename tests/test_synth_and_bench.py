"""-m "not gpu": generator determinism, bench.py's reference arm, and the N>1 launch path on gloo."""
import json
import os
import subprocess
import sys
import zlib
from pathlib import Path

import numpy as np

from readsb_b200 import synth

ROOT = Path(__file__).resolve().parent.parent


def test_generator_is_deterministic_and_seeded():
    a = synth.config2_stream(3, 100000)
    b = synth.config2_stream(3, 100000)
    c = synth.config2_stream(4, 100000)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(float(a.mean()) - 127.5) < 0.2
    dense = synth.config5_stream(3, 100000)
    assert dense.std() > a.std()


def test_generator_injects_all_phases():
    _, truth = synth.generate(2_400_000, seed=1, frames_per_sec=100, df_mask=synth.DF17, want_truth=True)
    assert len(truth) == 100
    assert {t[0] % 5 for t in truth} == {0, 1, 2, 3, 4}


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--streams", "8", "--buffers", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_two_rank_launch_path_gloo(tmp_path):
    """bench.py under torchrun at world_size 2: rank 0 alone runs the reference arm, rank 1 exits 0 without work;
    the stream sharding helper gives disjoint seeds per rank."""
    script = tmp_path / "w2.py"
    script.write_text(
        "import os, sys, json\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "import torch, torch.distributed as dist\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "S = 4\n"
        "seeds = torch.tensor([1 + r * S + s for s in range(S)])\n"
        "allseeds = [torch.zeros(S, dtype=torch.long) for _ in range(w)]\n"
        "dist.all_gather(allseeds, seeds)\n"
        "flat = torch.cat(allseeds).tolist()\n"
        "assert len(set(flat)) == w * S, flat\n"
        "t = torch.tensor([10.0 + r]); dist.all_reduce(t, op=dist.ReduceOp.MAX); assert t.item() == 10.0 + w - 1\n"
        "dist.barrier(); dist.destroy_process_group()\n"
        "import bench\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1', '--streams', '4', '--buffers', '1']\n"
        "sys.exit(bench.main())\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", str(script)], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference" and json.loads(lines[0])["n_gpus"] == 2
