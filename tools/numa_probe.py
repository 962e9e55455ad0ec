"""Host-link probe for the multi-GPU bench (run under torchrun): every rank copies 2.7 MB device->pinned host
(the inlier list of one fit) and 24 MB pinned host->device (one cloud) at the same moment as all other ranks,
with and without the rank's CPUs restricted to its GPU's NUMA node.  Prints the topology it saw.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/numa_probe.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def gpu_node(local_rank):
    try:
        import pynvml
        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        return bdf, int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
    except Exception as e:  # noqa: BLE001
        return repr(e), None


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e3 * ts[len(ts) // 2]


def main():
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    bdf, node = gpu_node(lr)
    cpu_now = os.sched_getcpu() if hasattr(os, "sched_getcpu") else -1
    aff0 = len(os.sched_getaffinity(0))
    d_small = torch.empty(2_700_000 // 8, dtype=torch.int64, device="cuda")
    d_big = torch.empty(3_000_000, dtype=torch.float64, device="cuda")
    out = {}
    for mode in ("unbound", "bound"):
        if mode == "bound":
            bench.pin_to_gpu_numa_node(lr)
        h_small = torch.empty(2_700_000 // 8, dtype=torch.int64).pin_memory()
        h_big = torch.empty(3_000_000, dtype=torch.float64).pin_memory()
        h_big.zero_()
        h_small.zero_()
        for _ in range(3):
            h_small.copy_(d_small, non_blocking=True)
            d_big.copy_(h_big, non_blocking=True)
        out[mode] = (timed(lambda: h_small.copy_(d_small, non_blocking=True), 40),
                     timed(lambda: d_big.copy_(h_big, non_blocking=True), 20))
    aff1 = len(os.sched_getaffinity(0))
    print(f"rank {rank} gpu {bdf} numa_node {node} started on cpu {cpu_now} affinity {aff0} -> {aff1} cpus | "
          f"d2h 2.7MB ms unbound {out['unbound'][0]:.3f} bound {out['bound'][0]:.3f} | "
          f"h2d 24MB ms unbound {out['unbound'][1]:.3f} bound {out['bound'][1]:.3f}", flush=True)
    dist.barrier()
    if rank == 0:
        os.system("lscpu | grep -i 'numa\\|socket\\|^CPU(s)'; nvidia-smi topo -m 2>&1 | head -24; "
                  "cat /sys/devices/system/node/node*/cpulist 2>&1 | head")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
