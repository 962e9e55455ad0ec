"""BASELINE config C5 under torchrun: fit_plane on a 4M-point cloud, 100k hypotheses sharded over the
ranks (strong scaling: the hypothesis batch is fixed), one NCCL all-gather of the counts per wave.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port P tools/bench_c5.py
Prints one JSON line on rank 0; every rank checks that it got rank 0's result."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from misc3d_b200 import capi, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.current_stream()
ctx = capi.Context(local, stream=stream.cuda_stream)
if world > 1:
    ids = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.init_nccl(ids[0], rank, world)
N, H = 4_000_000, 100_000
xyz = synth.make_c5(N)
cloud = ctx.upload(xyz)
res = None
for w in range(2):
    res = ctx.ransac_fit_cloud(capi.PLANE, cloud, 0.01, H, 1.0, seed=1, want_inliers=False)
steps = 3
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for s in range(steps):
    res = ctx.ransac_fit_cloud(capi.PLANE, cloud, 0.01, H, 1.0, seed=1, want_inliers=False)
e1.record(stream)
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
sig = torch.tensor([res[3]["best_index"], res[3]["best_count"]], device=dev, dtype=torch.int64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ref = sig.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, sig), "ranks disagree on the best hypothesis"
if rank == 0:
    t = float(ms.item()) * 1e-3
    print(json.dumps({"config": "C5 fit_plane 4M pts x 100k hypotheses", "n_gpus": world, "ms_per_fit": 1e3 * t,
                      "hypotheses_per_s": H / t, "point_hypotheses_per_s": N * H / t,
                      "best_index": res[3]["best_index"], "best_count": res[3]["best_count"],
                      "score_ms_rank0": res[3]["score_ms"]}), flush=True)
if world > 1:
    dist.destroy_process_group()
