"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol the
header declares, the host-only helpers agree with the oracle, and (world_size 2, gloo) the
hypothesis-sharded path reproduces the sequential result.  No compute entry point is called:
those need a GPU and refuse to run without one."""
import os
import re
import socket

import numpy as np
import pytest

from misc3d_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "m3d_capi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(m3d_[a-z0-9_]+)\s*\(", hdr)) - {"m3d_allgather_fn"})
    assert len(declared) >= 20
    L = capi.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == declared
    assert L.m3d_abi_version() == 2


def test_no_cpu_fallback(capi):
    if capi.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(capi.M3DError) as e:
        capi.Context(0)
    assert e.value.code == capi.ERR_CUDA


def test_product_does_not_touch_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "misc3d_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle", "").lower() or f == "synth.py", (dirpath, f)
    for d in ("include", "python"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "m3d_oracle" not in txt and "import orc" not in txt, (dirpath, f)


@pytest.mark.parametrize("k,n", [(3, 50000), (4, 1000), (2, 7), (3, 3)])
def test_sample_table_equals_oracle(capi, orc, k, n):
    a = capi.sample_table(99, n, k, 500)
    b = orc.sample_table(99, n, k, 500)
    np.testing.assert_array_equal(a, b)
    assert all(len(set(r)) == k for r in a.tolist())


def _oracle_counts(orc, kind, xyz, nrm, table, thr):
    """per-hypothesis (valid, count, err) with the oracle's building blocks"""
    H = len(table)
    valid = np.zeros(H, np.uint8)
    counts = np.zeros(H, np.uint64)
    err = np.zeros(H)
    for i, row in enumerate(table):
        idx = np.sort(row)
        ok, m = orc.minimal_fit(kind, xyz[idx], None if nrm is None else nrm[idx])
        if ok:
            valid[i] = 1
            counts[i], err[i] = orc.evaluate(kind, xyz, m, thr)
    return valid, counts, err


@pytest.mark.parametrize("prob", [0.9999, 0.99, 1.0])
def test_ordered_scan_replays_the_sequential_loop(capi, orc, prob):
    xyz = synth.make_c1(n=3000, seed=21)
    H, thr, seed = 300, 0.01, 5
    table = orc.sample_table(seed, len(xyz), 3, H)
    valid, counts, err = _oracle_counts(orc, orc.PLANE, xyz, None, table, thr)
    st = capi.ordered_scan(counts, valid, err, len(xyz), 3, prob, H)
    rc, model, inl, ost = orc.ransac_fit(orc.PLANE, xyz, thr=thr, max_it=H, prob=prob, seed=seed)
    for key in ("best_index", "best_count", "iterations_run", "stop_index", "found"):
        assert st[key] == ost[key], key
    if prob < 1.0:
        assert ost["stop_index"] < H  # the adaptive exit really fired in this case


def test_ordered_scan_tie_break_uses_rmse(capi):
    counts = np.array([10, 10, 10, 7], np.uint64)
    valid = np.ones(4, np.uint8)
    err = np.array([3.0, 2.0, 2.0, 0.1])
    st = capi.ordered_scan(counts, valid, err, 100, 3, 1.0, 4)
    assert st["best_index"] == 1  # strictly smaller rmse wins, an equal one does not (ransac.h:595-596)
    st = capi.ordered_scan(np.array([0, 0], np.uint64), np.ones(2, np.uint8), None, 100, 3, 1.0, 2)
    assert st["found"] == 0 and st["iterations_run"] == 2
    st = capi.ordered_scan(np.array([5, 9], np.uint64), np.array([0, 1], np.uint8), None, 100, 3, 1.0, 2)
    assert st["best_index"] == 1 and st["iterations_run"] == 1  # failed MinimalFit is not counted


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import orc
    from misc3d_b200 import capi, synth as sy
    from misc3d_b200.sharding import shard_rows, gather_counts, gathered_to_wave_order
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyz = sy.make_c1(n=2000, seed=8)
    H, thr, seed = 701, 0.01, 3   # three blocks of 256 rows: ranks own 2 and 1 (partial) blocks
    table = capi.sample_table(seed, len(xyz), 3, H)  # every rank draws the same global table
    mine, S = shard_rows(H, rank, world)
    c_mine, c_S = capi.shard_rows(H, rank, world)    # the library's own partition (m3d_shard_rows)
    assert np.array_equal(mine, c_mine) and S == c_S
    valid, counts, _ = _oracle_counts(orc, orc.PLANE, xyz, None, table[mine], thr)
    packed = np.zeros(S, np.uint32)
    packed[: len(mine)] = counts.astype(np.uint32) | ((1 - valid).astype(np.uint32) << 31)
    allc = gathered_to_wave_order(gather_counts(packed, world, lambda t, outs: dist.all_gather(outs, t)), H, world)
    st = capi.ordered_scan(allc & 0x7FFFFFFF, ((allc >> 31) == 0).astype(np.uint8), None, len(xyz), 3, 0.9999, H)
    q.put((rank, st["best_index"], st["best_count"], st["iterations_run"], st["stop_index"]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_counts_over_gloo_match_single_process(capi, orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    xyz = synth.make_c1(n=2000, seed=8)
    rc, model, inl, ost = orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=701, prob=0.9999, seed=3)
    for r in res:
        assert r[1:] == (ost["best_index"], ost["best_count"], ost["iterations_run"], ost["stop_index"])


def _record_worker(rank, world, port, q):
    """probability == 1: every rank reduces its shard to one 64-byte record (sharding.best_record = the device's
    wave_best_kernel), the records are all-gathered (gloo) and merged (best_merge_kernel)"""
    import torch
    import torch.distributed as dist
    import orc
    from misc3d_b200 import capi, synth as sy
    from misc3d_b200.sharding import shard_rows, best_record, merge_records, REC_WORDS
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyz = sy.make_c1(n=2000, seed=8)
    H, thr, seed = 701, 0.01, 3
    table = capi.sample_table(seed, len(xyz), 3, H)
    mine, S = shard_rows(H, rank, world)
    valid, counts, _ = _oracle_counts(orc, orc.PLANE, xyz, None, table[mine], thr)
    packed = counts.astype(np.uint32) | ((1 - valid).astype(np.uint32) << 31)
    rec = torch.from_numpy(best_record(packed, mine, len(xyz)).astype(np.int32))
    outs = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(outs, rec)
    m = merge_records(torch.stack(outs).numpy().astype(np.uint32).reshape(world, REC_WORDS))
    q.put((rank, int(m[0]), int(m[1]), int(m[2]), int(m[3]), int(m[4])))
    dist.barrier()
    dist.destroy_process_group()


def test_best_record_exchange_over_gloo_matches_sequential_loop(capi, orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_record_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    xyz = synth.make_c1(n=2000, seed=8)
    rc, model, inl, ost = orc.ransac_fit(orc.PLANE, xyz, thr=0.01, max_it=701, prob=1.0, seed=3)
    for r in res:
        max_count, first_row, n_tied, n_valid, first_full = r[1:]
        assert max_count == ost["best_count"] and n_valid == ost["iterations_run"] and first_full == 0xFFFFFFFF
        if n_tied == 1:
            assert first_row == ost["best_index"]


def test_best_record_restatement_properties():
    """the record keeps exactly what the sequential best-update can end on: rows at the maximum count"""
    from misc3d_b200.sharding import shard_rows, best_record, merge_records, NO_ROW, TIED_CAP
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8):
        rows = 3000
        counts = rng.integers(0, 50, rows).astype(np.uint32)
        counts[rng.integers(0, rows, 40)] = 49            # ties at the maximum
        invalid = rng.random(rows) < 0.1
        packed = counts | (invalid.astype(np.uint32) << 31)
        recs = []
        for r in range(world):
            mine, S = shard_rows(rows, r, world)
            recs.append(best_record(packed[mine], mine, 10**6))
        m = merge_records(np.stack(recs))
        ok = ~invalid
        best = counts[ok].max()
        tied = np.flatnonzero(ok & (counts == best))
        assert m[0] == best and m[1] == tied[0] and m[3] == ok.sum() and m[4] == NO_ROW
        assert m[2] == len(tied) or (m[2] > TIED_CAP and len(tied) > TIED_CAP)
        if len(tied) <= TIED_CAP:
            assert list(m[8: 8 + len(tied)]) == list(tied)


@pytest.mark.parametrize("rows,world", [(1, 1), (101, 2), (256, 2), (257, 3), (10000, 8), (80000, 8), (65536, 4),
                                        (12345, 5), (511, 8)])
def test_shard_map_is_a_partition(capi, rows, world):
    """block-cyclic hypothesis sharding (csrc/scan.h ShardMap, m3d_shard_rows): the ranks' row lists partition
    the wave, the python restatement agrees, shards are balanced to within one block, and the rank-major
    all-gathered buffer maps back to wave order"""
    from misc3d_b200.sharding import shard_rows, gathered_to_wave_order, SHARD_BLOCK
    seen = np.zeros(rows, np.int32)
    sizes, S0 = [], None
    gathered = None
    for r in range(world):
        mine, S = capi.shard_rows(rows, r, world)
        py_mine, py_S = shard_rows(rows, r, world)
        assert np.array_equal(mine, py_mine) and S == py_S
        S0 = S if S0 is None else S0
        assert S == S0 and len(mine) <= S
        seen[mine] += 1
        sizes.append(len(mine))
        if gathered is None:
            gathered = np.full(world * S, -1, np.int64)
        gathered[r * S: r * S + len(mine)] = mine          # each rank contributes "its rows" as payload
    assert np.all(seen == 1)
    assert max(sizes) - min(sizes) <= SHARD_BLOCK
    assert np.array_equal(gathered_to_wave_order(gathered, rows, world), np.arange(rows))


def test_sampler_blocks_avx2_equals_generic_and_std_mt19937(capi, orc):
    """the block generator behind the sample tables (csrc/sampler.cpp): the run-time selected AVX2 body equals
    the generic one, and the tables equal std::mt19937 + `%` (the oracle draws with <random>) including
    duplicate rejection on tiny clouds and block boundaries"""
    import ctypes as C
    L = capi.lib()
    for seed, n in ((1, 1000000), (2, 7), (3, 2**31 - 1), (4, 3), (5, 50000), (6, 1), (7, 2), (8, 4000000)):
        assert L.m3d_sampler_selfcheck(C.c_uint32(seed), C.c_size_t(n), 100) == 1
    for seed, n, k, rows in ((1, 1000000, 3, 30000), (5, 7, 4, 3000), (9, 3, 3, 500), (3, 50000, 2, 30000),
                             (11, 4, 4, 2000), (13, 5, 2, 5000), (0, 4000000, 3, 100000)):
        assert np.array_equal(capi.sample_table(seed, n, k, rows), orc.sample_table(seed, n, k, rows)), (seed, n, k)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the
    contract's keys, timed on the reference's own compiled sources when oracle/_ref is there"""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1")          # as under torchrun: the arm must ask for the cores itself
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # a non-zero rank of a multi-rank launch prints nothing and exits 0
    env2 = dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                        capture_output=True, text=True, timeout=120, env=env2)
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_mt_jump_table_matches_generator():
    """csrc/mt_jump_table.h (jump-ahead polynomials of mt19937 used by the device-side sample draw) is what
    tools/gen_mt_jump.py produces, and every polynomial maps the first 19937+623 raw words of a stream onto the
    624-word window at its segment start (checked against the sequentially generated stream)"""
    import importlib.util
    import re
    spec = importlib.util.spec_from_file_location("gen_mt_jump", os.path.join(ROOT, "tools", "gen_mt_jump.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = open(os.path.join(ROOT, "misc3d_b200", "csrc", "mt_jump_table.h")).read()
    assert f"kMtSegBlocks = {gen.SEG_BLOCKS};" in text and f"kMtMaxSegments = {gen.MAX_SEGMENTS};" in text
    words = np.array([int(w, 16) for w in re.findall(r"0x([0-9a-f]{8})u", text)], dtype=np.uint32)
    assert words.size == (gen.MAX_SEGMENTS - 1) * gen.WORDS
    table = words.reshape(gen.MAX_SEGMENTS - 1, gen.WORDS)
    y = gen.raw_stream(20240607, gen.SEG_BLOCKS * (gen.MAX_SEGMENTS - 1) + 1)
    # the raw stream itself against numpy's mt19937 (tempered): temper(y) must equal random_raw()
    t = y.copy()
    t ^= t >> np.uint32(11)
    t ^= (t << np.uint32(7)) & np.uint32(0x9D2C5680)
    t ^= (t << np.uint32(15)) & np.uint32(0xEFC60000)
    t ^= t >> np.uint32(18)
    bg = np.random.MT19937()
    bg._legacy_seeding(20240607)
    np.testing.assert_array_equal(t[:5000], bg.random_raw(5000).astype(np.uint32))
    for p in (1, 2, 7, gen.MAX_SEGMENTS - 1):
        g = int.from_bytes(table[p - 1].astype("<u4").tobytes(), "little")
        j = p * gen.SEG_BLOCKS * 624
        np.testing.assert_array_equal(gen.apply_poly(g, y), y[j:j + 624])
    polys = gen.build()
    for p, g in enumerate(polys, 1):
        assert int.from_bytes(table[p - 1].astype("<u4").tobytes(), "little") == g, p


def test_cooperative_upload_slices_tile_the_cloud():
    """multi-rank host-buffer fit: the slices the ranks upload cover the cloud exactly once, whatever the divisibility,
    and fit the all-gather buffer of 3 n + world doubles"""
    from misc3d_b200 import sharding
    for n in (65536, 70001, 1_000_000, 999_983):
        for world in (2, 3, 4, 8):
            covered = 0
            for r in range(world):
                off, length, s = sharding.upload_slice(n, r, world)
                assert off == covered or length == 0
                covered += length
                assert off + length <= 3 * n
                assert s * world <= 3 * n + world - 1   # the receive buffer holds world x S doubles (3 n + world reserved)
            assert covered == 3 * n
