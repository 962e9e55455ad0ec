"""torchrun target: hypothesis-sharded fits over R ranks (NCCL all-gather of the counts inside the
library) must return exactly what a single-rank fit returns -- R-invariance incl. early exit.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from misc3d_b200 import capi, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
single = capi.Context(local)
sharded = capi.Context(local)
ids = [capi.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
sharded.init_nccl(ids[0], rank, world)

xyz, nrm = synth.make_c2(n=200000, seed=9)
ok = True
for kind in (0, 1, 2):
    # 5000 rows: one launch per rank; 20000 / 60000 rows: shards issued in parts (pipelined table draw) and,
    # at 60000, pre-sorted into culled / dense hypotheses
    for prob, H in ((0.9999, 3000), (1.0, 5000), (1.0, 20000)) + (((1.0, 60000),) if kind == 0 else ()):
        a = single.ransac_fit(kind, xyz, nrm if kind == 2 else None, 0.01, H, prob, seed=11)
        b = sharded.ransac_fit(kind, xyz, nrm if kind == 2 else None, 0.01, H, prob, seed=11)
        same = (a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and
                all(a[3][k] == b[3][k] for k in ("best_index", "best_count", "iterations_run", "stop_index")))
        ok = ok and same
        if rank == 0:
            print(f"kind {kind} prob {prob}: sharded == single: {same}; best {b[3]['best_index']} count {b[3]['best_count']} "
                  f"stop {b[3]['stop_index']} evaluated {b[3]['evaluated']} score_ms single {a[3]['score_ms']:.3f} "
                  f"sharded {b[3]['score_ms']:.3f}", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
