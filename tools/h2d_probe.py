#!/usr/bin/env python
"""Host-to-device supply ceiling of one box: every rank (one process per GPU, torchrun) copies its own pinned host
buffer to its GPU in a loop; aggregate GB/s = what any end-to-end path whose input starts in host memory can
reach at that GPU count.  Variants show where a shortfall comes from:

  default   cudaHostAlloc from an unbound process (what bench.py's e2e leg did in round 1)
  bound     the process is pinned to the CPUs NVML reports as local to its GPU BEFORE the allocation, so the
            first-touch policy places the pinned pages on the GPU's NUMA node (what aa_host_alloc does now)
  wc        bound + cudaHostAllocWriteCombined
  d2h       the way back (records are ~2 % of the input bytes, so this one hardly matters)

Rank 0 prints one JSON line.  usage: torchrun --nproc-per-node N tools/h2d_probe.py [GiB per rank = 2] [iters = 6]"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist


def gpu_cpus(index):
    """CPUs local to GPU `index` according to NVML (empty list if unknown)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        return cpus
    except Exception:
        return []


def numa_nodes():
    try:
        return sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except Exception:
        return []


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rt = None
    import glob

    for cand in ["libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12"] + glob.glob(
            os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
        try:
            rt = C.CDLL(cand)
            break
        except OSError:
            continue
    if rt is None:
        raise SystemExit("libcudart not found")
    rt.cudaSetDevice(local)
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    nbytes = int(gib * (1 << 30))
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    all_cpus = sorted(os.sched_getaffinity(0))
    local_cpus = [c for c in gpu_cpus(local) if c in all_cpus]

    def alloc(flags):
        p = C.c_void_p()
        e = rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags))
        if e != 0:
            raise RuntimeError(f"cudaHostAlloc failed: {e}")
        C.memset(p, 1, nbytes)                     # touch every page from this thread
        return p

    def copy_rate(p, to_device=True):
        kind = 1 if to_device else 2
        fn = rt.cudaMemcpyAsync
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        args = (C.c_void_p(d.data_ptr()), p) if to_device else (p, C.c_void_p(d.data_ptr()))
        fn(args[0], args[1], nbytes, kind, C.c_void_p(stream.cuda_stream))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn(args[0], args[1], nbytes, kind, C.c_void_p(stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        mine = torch.tensor([iters * nbytes / (ms.item() / 1e3) / 1e9], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            per = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(per, mine)
            per = [float(x.item()) for x in per]
        else:
            per = [float(mine.item())]
        return {"aggregate_gbs": world * iters * nbytes / (ms.item() / 1e3) / 1e9, "per_rank_gbs": [round(x, 1) for x in per]}

    out = {"n_gpus": world, "gib_per_rank": gib, "iters": iters, "host_cpus": len(all_cpus), "numa_nodes": numa_nodes(),
           "gpu_local_cpus": f"{local_cpus[0]}-{local_cpus[-1]} ({len(local_cpus)})" if local_cpus else None}
    p = alloc(0)
    out["default"] = copy_rate(p)
    out["d2h_default"] = copy_rate(p, to_device=False)
    rt.cudaFreeHost(p)
    if local_cpus:
        os.sched_setaffinity(0, local_cpus)
    p = alloc(0)
    out["bound"] = copy_rate(p)
    rt.cudaFreeHost(p)
    p = alloc(4)                                   # cudaHostAllocWriteCombined
    out["wc"] = copy_rate(p)
    rt.cudaFreeHost(p)
    os.sched_setaffinity(0, all_cpus)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
