"""N > 1 host logic on CPU: world_size 2 over gloo (the GPU runs use the same code over NCCL)."""
import json
import os
import subprocess
import sys

import numpy as np

import common
from common import wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bilevel-gait-gen_b200"))
import sharding  # noqa: E402


def test_shard_ranges_partition_the_batch():
    for total in (1, 5, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            cuts = [sharding.shard_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_gloo_ranks_reproduce_the_single_process_result(tmp_path):
    total = 5   # odd on purpose: ragged slices
    out = tmp_path / "gathered.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "dist_worker.py"), str(total), str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    got = json.load(open(out))
    assert got["world"] == 2 and got["tmax"] == [2.0, 2.0]
    cfg_name = "a1_configuration"
    states, t0, ee = wl.batched_trot_inputs(wl.CONFIGS[cfg_name], total, seed=9)
    for b in range(total):
        o = common.make_oracle(cfg_name, states[b])
        st = o.solve(states[b], 0.0, ee[b], real_time=True)
        assert got["status"][b] == st
        assert got["cost"][b] == o.cost()   # same code, same inputs: bit-identical whichever rank solved it
