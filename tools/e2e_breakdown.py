"""Where the time of the host-buffer entry point (m3d_ransac_fit) goes on the C2 cloud: wall clock per
call beside the device time of the fit itself, and the pinned H2D / D2H bandwidth of the box."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misc3d_b200 import capi, synth  # noqa: E402

xyz, nrm = synth.make_c2()
h_xyz = torch.from_numpy(xyz).pin_memory()
h_nrm = torch.from_numpy(nrm).pin_memory()
ctx = capi.Context(0, stream=torch.cuda.current_stream().cuda_stream)
inl_buf = torch.empty(len(xyz), dtype=torch.int64).pin_memory().numpy().view(np.uint64)
d = torch.empty_like(h_xyz, device="cuda")
out = {}
for name, src, dst in (("h2d_pinned", h_xyz, d), ("d2h_pinned", d, h_xyz)):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    out[name + "_GBps"] = xyz.nbytes / min(ts) / 1e9
pag = torch.from_numpy(xyz)
ts = []
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d.copy_(pag)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
out["h2d_pageable_GBps"] = xyz.nbytes / min(ts) / 1e9
cloud = ctx.upload(h_xyz.numpy(), h_nrm.numpy())
for kind, name in ((0, "plane"), (1, "sphere"), (2, "cylinder")):
    host, res, dev, sc = [], [], [], []
    for rep in range(5):
        t0 = time.perf_counter()
        rc, m, inl, st = ctx.ransac_fit(kind, h_xyz.numpy(), h_nrm.numpy() if kind == 2 else None, 0.01, 10000, 1.0,
                                        seed=rep, inl_buf=inl_buf)
        host.append(time.perf_counter() - t0)
        dev.append(st["device_ms"])
        sc.append(st["score_ms"])
        t0 = time.perf_counter()
        rc, m, inl, st = ctx.ransac_fit_cloud(kind, cloud, 0.01, 10000, 1.0, seed=rep, inl_buf=inl_buf)
        res.append(time.perf_counter() - t0)
    out[name] = {"host_call_ms": 1e3 * min(host[1:]), "resident_call_ms": 1e3 * min(res[1:]),
                 "device_fit_ms": min(dev[1:]), "score_ms": min(sc[1:]), "n_inl": int(len(inl))}
print(json.dumps(out))
