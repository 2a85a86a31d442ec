"""Bare pinned host->device copy probe: what does the host fabric give each rank, alone and with
all ranks copying at once, for the allocation / binding variants the input pipeline could use?

    python profiles/h2d_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/h2d_probe.py

Variants: pinned allocation {default (torch pin_memory), write-combined (cudaHostAlloc WC), numa-spread (rank bound to
node rank % nodes before first touch), interleaved (MPOL_INTERLEAVE over all nodes)} x copy streams {1, 2} x piece size
{32 MB, 256 MB}.  Each rank copies 1 GiB per measurement; `alone` = ranks take turns, `together` = all at once
(barrier before, max over ranks after).  Prints one JSON object on rank 0."""
import ctypes
import glob
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import _parse_cpulist, host_topology      # noqa: E402

TOTAL = 1 << 30


def set_interleave(nodes):
    """set_mempolicy(MPOL_INTERLEAVE, mask) via the raw syscall (no libnuma in the image)."""
    mask = 0
    for n in nodes:
        mask |= 1 << n
    arr = (ctypes.c_ulong * 1)(mask)
    libc = ctypes.CDLL(None, use_errno=True)
    rc = libc.syscall(238, 3 if nodes else 0, arr if nodes else None, 65 if nodes else 0)   # 3 = MPOL_INTERLEAVE, 0 = DEFAULT
    return rc


def alloc(kind, nbytes, rank, nodes, allowed):
    """Returns (torch uint8 CPU tensor viewing pinned memory, keep-alive)."""
    if kind == "wc":
        cudart = ctypes.CDLL("libcudart.so.12")
        ptr = ctypes.c_void_p()
        rc = cudart.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))   # cudaHostAllocWriteCombined
        if rc != 0:
            raise RuntimeError("cudaHostAlloc(WC) -> %d" % rc)
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=np.uint8)
        arr[::4096] = 1                      # touch
        return torch.from_numpy(arr), (cudart, ptr, buf)
    if kind == "numa_spread" and len(nodes) > 1:
        order = sorted(nodes)
        os.sched_setaffinity(0, nodes[order[rank % len(order)]] & allowed or allowed)
    if kind == "interleave" and len(nodes) > 1:
        set_interleave(sorted(nodes))
    t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    t[::4096] = 1
    if kind == "interleave":
        set_interleave([])
    if kind == "numa_spread":
        os.sched_setaffinity(0, allowed)
    return t, None


def copy_once(src, dst, streams, piece):
    n = src.numel()
    evs = []
    for i, st in enumerate(streams):
        with torch.cuda.stream(st):
            for lo in range(i * piece, n, piece * len(streams)):
                hi = min(n, lo + piece)
                dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
    for st in streams:
        st.synchronize()


def measure(src, dst, n_streams, piece, world, together):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    copy_once(src, dst, streams, piece)        # warm
    if world > 1:
        dist.barrier()
    if together or world == 1:
        t0 = time.perf_counter()
        copy_once(src, dst, streams, piece)
        dt = time.perf_counter() - t0
    else:
        dt = None
        for r in range(world):
            if dist.get_rank() == r:
                t0 = time.perf_counter()
                copy_once(src, dst, streams, piece)
                dt = time.perf_counter() - t0
            dist.barrier()
    return src.numel() / dt / 1e9


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    topo = host_topology(torch)
    allowed = os.sched_getaffinity(0)
    nodes = {k: v & allowed for k, v in topo["nodes"].items() if v & allowed}
    dst = torch.empty(TOTAL, dtype=torch.uint8, device="cuda")
    out = {}
    kinds = ["default", "wc"] + (["numa_spread", "interleave"] if len(nodes) > 1 else [])
    for kind in kinds:
        try:
            src, keep = alloc(kind, TOTAL, rank, nodes, allowed)
        except Exception as e:          # noqa: BLE001
            out[kind] = {"error": str(e)}
            continue
        for n_streams in (1, 2):
            for piece in (32 << 20, 256 << 20):
                for together in ((False, True) if world > 1 else (True,)):
                    gbs = measure(src, dst, n_streams, piece, world, together)
                    t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
                    lo, hi = t.clone(), t.clone()
                    if world > 1:
                        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                    key = "%s/streams%d/piece%dMB/%s" % (kind, n_streams, piece >> 20, "together" if together else "alone")
                    out[key] = {"GBps_per_rank_min": float(lo.item()), "GBps_per_rank_max": float(hi.item()),
                                "aggregate_GBps_at_min": float(lo.item()) * (world if together else 1)}
        del src, keep
    if rank == 0:
        info = {"world": world, "numa_nodes": {str(k): len(v) for k, v in nodes.items()},
                "gpu_local_cpu_lists_distinct": len({frozenset(c) for c in topo["gpu_cpus"] if c}),
                "cpus_allowed": len(allowed), "results": out}
        print(json.dumps(info))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
