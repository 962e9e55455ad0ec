"""torchrun target: hypothesis-sharded fits over R ranks must return exactly what a single-rank fit returns --
R-invariance incl. early exit (probability < 1: all-gather of the counts + host replay; probability == 1: one
64-byte best record per rank, device-side arg-best, ties / fitness-1 replayed on the host).
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
# resident clouds (device normals: the cylinder's sample table is drawn on the device too), count ties, a perfect fit
cs, cm = single.upload(xyz, nrm), sharded.upload(xyz, nrm)
rng = np.random.default_rng(3)
lattice = np.round(rng.uniform(-1, 1, (60, 3)), 0)
flat = np.c_[rng.uniform(-1, 1, (5000, 2)), np.zeros(5000)]
cases = [(kind, lambda c, kind=kind: c.ransac_fit_cloud(kind, cs if c is single else cm, 0.01, 12000, 1.0, seed=5), f"resident kind {kind}")
         for kind in (0, 1, 2)]
odd_xyz, odd_nrm = synth.make_c2(n=70001, seed=10)  # 3 n not divisible by the rank count: ragged upload slices
cases += [(kind, lambda c, kind=kind: c.ransac_fit(kind, odd_xyz, odd_nrm if kind == 2 else None, 0.01, 4000, 1.0, seed=6),
           f"host buffers, 70001 points, kind {kind} (upload sharded over the ranks + all-gather)") for kind in (0, 1, 2)]
cases += [(0, lambda c: c.ransac_fit(0, lattice, None, 0.3, 3000, 1.0, seed=2), "ties (lattice cloud)"),
          (0, lambda c: c.ransac_fit(0, flat, None, 0.01, 2000, 1.0, seed=4), "fitness 1 stops the loop")]
for kind, run, name in cases:
    a, b = run(single), run(sharded)
    same = (a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and
            all(a[3][k] == b[3][k] for k in ("best_index", "best_count", "iterations_run", "stop_index")))
    ok = ok and same
    if rank == 0:
        print(f"{name}: sharded == single: {same}; best {b[3]['best_index']} count {b[3]['best_count']} stop {b[3]['stop_index']}", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
