"""torchrun --nproc-per-node N profiles/multi_gpu_check.py
Row-sharded neighbourhood / quantile_fast with an NCCL halo exchange (config 4 shape: 8000 x 8000, halfwidth 15,
20 thresholds) checked against the single-GPU whole-field result, plus device-timed throughput."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridpp_b200 as gpp
from gridpp_b200 import device as gd, distributed as gdist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
gpp.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, hw = 8000, 15
thr = np.linspace(0, 1, 20).astype(np.float32)
g = torch.Generator(device="cuda").manual_seed(1000)
field = torch.rand((n, n), device="cuda", generator=g)          # same seed on every rank -> same field
r0, r1 = gdist.row_block(n, world, rank)
tile = field[r0:r1].contiguous()
res = {}
for name, fn_sharded, fn_whole in (
        ("mean_hw15", lambda: gdist.neighbourhood(tile, hw, gpp.Mean), lambda: gd.neighbourhood(field, hw, gpp.Mean)),
        ("quantile_fast_hw15_T20", lambda: gdist.neighbourhood_quantile_fast(tile, 0.5, hw, thr),
         lambda: gd.neighbourhood_quantile_fast(field, 0.5, hw, thr))):
    mine = fn_sharded()
    whole = fn_whole()
    same = torch.equal(torch.nan_to_num(mine, nan=-7.0), torch.nan_to_num(whole[r0:r1], nan=-7.0))
    flag = torch.tensor([int(same)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for _ in range(2):
        fn_sharded()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn_sharded()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = {"sharded_equals_whole": bool(flag.item()), "ms_incl_halo_exchange": float(t.item()),
                 "GB/s_aggregate": 8.0 * n * n / (float(t.item()) * 1e-3) / 1e9}
if rank == 0:
    print(json.dumps({"n_gpus": world, "shape": [n, n], "results": res}))
dist.destroy_process_group()
