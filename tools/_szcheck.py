import sys; sys.path.insert(0, "tests")
import numpy as np, common
from common import wl
cfg_name="a1_configuration"; cfg=wl.CONFIGS[cfg_name]
states,t0,ee=wl.batched_trot_inputs(cfg,4,seed=0)
gpu=common.make_gpu(cfg_name,4,states)
for k in range(2):
    out=gpu.GetRealTimeUpdate(states,t0,ee)
    print(out["iters"], [gpu.sizes(b)["refined_iters"] for b in range(4)], [gpu.sizes(b)["no_iterate"] for b in range(4)], gpu.sizes(0)["qp_cost"])
