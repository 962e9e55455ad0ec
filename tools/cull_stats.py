"""Survival statistics of the culling score kernel on the C2 cloud (needs the stats build:
make -C misc3d_b200/csrc VARIANT=stats EXTRA=-DM3D_CULL_STATS; run with M3D_LIB=.../libm3d_stats.so)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from misc3d_b200 import capi, synth  # noqa: E402

xyz, nrm = synth.make_c2()
ctx = capi.Context(0)
cloud = ctx.upload(xyz, nrm)
L = capi.lib()
out = {}
for kind, name in ((0, "plane"), (1, "sphere"), (2, "cylinder")):
    tab = capi.sample_table(1, len(xyz), capi.KSAMPLE[kind], 10000)
    ctx.score_samples(kind, cloud, tab, 0.01, want_models=False)
    z = (C.c_ulonglong * 8)()
    L.m3d_debug_cull_stats(z)
    t0 = time.perf_counter()
    ctx.score_samples(kind, cloud, tab, 0.01, want_models=False)
    dt = time.perf_counter() - t0
    L.m3d_debug_cull_stats(z)
    tests, live, cells = z[0], z[1], z[2]
    out[name] = {"ms_with_stats": 1e3 * dt, "hyp_tile_tests": tests, "tile_survival": live / max(tests, 1),
                 "cells_per_live_tile": cells / max(live, 1), "pair_survival": cells / max(tests, 1) / 32}
print(json.dumps(out))
