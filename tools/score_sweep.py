"""Tuning sweep of the scoring kernel on the C2 cloud (run on the GPU box).

  python tools/score_sweep.py build      # compile the library variants (CPU box, nvcc)
  python tools/score_sweep.py run        # time every (library variant x launch variant x primitive)
  python tools/score_sweep.py one        # (internal) one configuration, JSON line on stdout
  python tools/score_sweep.py raw        # wall clock of m3d_score_samples per primitive (10k hypotheses, C2 cloud);
                                         # M3D_SCORE_PATH=dense / M3D_SCORE_VARIANT=256x4 / M3D_LIB=... select variants
"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIBS = {  # name -> extra nvcc flags (tuning builds land in misc3d_b200/variants/, select with M3D_LIB=...)
    "scalar": "-DM3D_PACKED=0",        # dense kernel without the fp32x2 inner loop
    "stats": "-DM3D_CULL_STATS",       # culling kernel with survival counters (tools/cull_stats.py)
    "st3": "-DM3D_CULL_STAGES=3",      # ring depth of the culling kernel
}
LAUNCH = ["256x2", "128x2", "256x1", "128x4", "256x4", "128x1"]


def build():
    for name, extra in LIBS.items():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "misc3d_b200", "csrc"), "-s", "-j4",
                               f"VARIANT={name}", f"EXTRA={extra}"])


def one():
    import numpy as np
    from misc3d_b200 import capi, synth
    xyz, nrm = synth.make_c2()
    ctx = capi.Context(0)
    cloud = ctx.upload(xyz, nrm)
    out = {}
    for kind, name in ((0, "plane"), (1, "sphere"), (2, "cylinder")):
        ms = []
        for rep in range(4):
            rc, model, inl, st = ctx.ransac_fit_cloud(kind, cloud, 0.01, 10000, 1.0, seed=rep, want_inliers=False)
            ms.append(st["score_ms"])
        out[name] = {"score_ms": float(np.min(ms[1:])), "resolves": st["exact_resolves"],
                     "Gpairs_per_s": 1e10 / (float(np.min(ms[1:])) * 1e-3) / 1e9}
    print(json.dumps(out))


def run():
    for launch in LAUNCH:
        env = dict(os.environ, M3D_SCORE_VARIANT=launch)
        r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            d = {"error": (r.stderr or r.stdout)[-400:]}
        print(json.dumps({"launch": launch, **d}), flush=True)


def raw():
    """wall-clock of m3d_score_samples only (no self-check)"""
    import time
    import numpy as np
    from misc3d_b200 import capi, synth
    xyz, nrm = synth.make_c2()
    ctx = capi.Context(0)
    cloud = ctx.upload(xyz, nrm)
    out = {}
    for kind, name in ((0, "plane"), (1, "sphere"), (2, "cylinder")):
        tab = capi.sample_table(1, len(xyz), capi.KSAMPLE[kind], 10000)
        ts = []
        for rep in range(5):
            t0 = time.perf_counter()
            ctx.score_samples(kind, cloud, tab, 0.01, want_models=False)
            ts.append(time.perf_counter() - t0)
        out[name] = round(1e3 * min(ts[1:]), 3)
    print(json.dumps(out))


if __name__ == "__main__":
    {"build": build, "run": run, "one": one, "raw": raw}[sys.argv[1]]()
