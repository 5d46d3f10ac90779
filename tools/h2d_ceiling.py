#!/usr/bin/env python
"""Host-to-device ceiling of the end-to-end path: N concurrent ranks, nothing but pinned `cudaMemcpyAsync`.

    python tools/h2d_ceiling.py                          # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py

bench.py's e2e arm streams 73.5 MB of u8 frames per 1024 pairs per GPU from pinned host memory (uahn_submit_sequence).
When 8 ranks do that at once the shared PCIe uplinks / host memory bound the job, not the GPUs (round 1: 2.41 M pairs/s
end to end against 3.13 M resident).  This tool measures that bound in isolation, for each of: plain pinned buffers
(what torch.pin_memory / cudaHostAlloc default gives), write-combined pinned buffers, one or two copies in flight per
GPU, with and without binding the rank to the GPU's NUMA node.  Rank 0 prints one JSON line.
"""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=73.5, help="MB per copy (one 1024-pair sequence submission)")
ap.add_argument("--copies", type=int, default=40)
ap.add_argument("--no-bind", action="store_true")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
from bench import bind_to_gpu_numa_node  # noqa: E402
numa = {"why": "--no-bind"} if a.no_bind else bind_to_gpu_numa_node(local)
nbytes = int(a.mb * 1e6) // 4096 * 4096
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
if rt is None:
    for cand in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = ctypes.CDLL(cand)
            break
        except OSError:
            pass


def host_alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc flags={flags} rc={rc}")
    ctypes.memset(p, 1, nbytes)       # first touch on this (possibly NUMA-bound) thread
    return p


def reduce_max(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measure(host_ptr, in_flight):
    dsts = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(in_flight)]
    streams = [torch.cuda.Stream(dev) for _ in range(in_flight)]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    def go(n):
        for i in range(n):
            k = i % in_flight
            rt.cudaMemcpyAsync(ctypes.c_void_p(dsts[k].data_ptr()), host_ptr, nbytes, 1, ctypes.c_void_p(streams[k].cuda_stream))
        for s in streams:
            s.synchronize()
    go(4)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    go(a.copies)
    dt = reduce_max(time.perf_counter() - t0)
    return world * nbytes * a.copies / dt / 1e9


out = {"ranks": world, "mb_per_copy": nbytes / 1e6, "copies": a.copies, "numa_binding_rank0": numa, "gbs_all_ranks": {}}
if rt is not None:
    for name, flags in (("pinned", 0), ("pinned_write_combined", 4)):
        hp = host_alloc(flags)
        for fl in (1, 2):
            out["gbs_all_ranks"][f"{name}_{fl}_in_flight"] = measure(hp, fl)
        rt.cudaFreeHost(hp)
best = max(out["gbs_all_ranks"].values()) if out["gbs_all_ranks"] else None
out["best_gbs_all_ranks"] = best
out["pairs_per_s_ceiling_sequence_mode"] = best * 1e9 / 71750 if best else None     # 71 680 B per frame + prior
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
