"""Worker of tests/test_sharding_cpu.py: one of WORLD_SIZE gloo ranks.  Every rank solves its slice of a small batch
with the CPU oracle (the CUDA path needs a GPU), then the results travel exactly as bench.py moves them."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "bilevel-gait-gen_b200")]
import common  # noqa: E402
import sharding  # noqa: E402
from common import wl  # noqa: E402


def main():
    total, out_path = int(sys.argv[1]), sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg_name = "a1_configuration"
    states, t0, ee = wl.batched_trot_inputs(wl.CONFIGS[cfg_name], total, seed=9)
    lo, hi = sharding.shard_range(total, rank, world)
    cost, status = np.zeros(hi - lo), np.zeros(hi - lo, np.int32)
    for i, b in enumerate(range(lo, hi)):
        o = common.make_oracle(cfg_name, states[b])
        status[i] = o.solve(states[b], 0.0, ee[b], real_time=True)
        cost[i] = o.cost()
    dist.barrier()
    tmax = sharding.max_over_ranks([float(rank + 1), float(world - rank)], dist)
    g_cost = sharding.gather_to_root(cost, total, dist)
    g_status = sharding.gather_to_root(status, total, dist)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"cost": g_cost.tolist(), "status": g_status.tolist(), "tmax": tmax, "world": world}, f)
    else:
        assert g_cost is None and g_status is None
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
