"""Probe (GPU box, torchrun -N ranks): bare pinned-host -> device copy rate per GPU with ALL ranks copying at once -- the
platform ceiling under bench.py's `e2e` leg (101.7 MB of bf16 features per step and rank).  One cudaMemcpyAsync per
buffer (text 50.3 MB, image 51.4 MB), 30 rounds, CUDA events; run with and without NUMA binding of the host buffers."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
bind = os.environ.get("H2D_BIND", "1") == "1"
cpus = None
if bind:
    import bench
    cpus = bench.bind_to_gpu_numa_node(lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
B, L = 256, 128
host = [torch.empty(B, L, 768, dtype=torch.bfloat16).pin_memory(), torch.empty(B, 49, 2048, dtype=torch.bfloat16).pin_memory()]
for h in host:
    h.zero_()
devb = [torch.empty_like(h, device=dev) for h in host]
nbytes = sum(h.numel() * 2 for h in host)
res = {}
for mode in ("one stream", "two streams"):
    s2 = torch.cuda.Stream()
    def round_():
        devb[0].copy_(host[0], non_blocking=True)
        if mode == "two streams":
            s2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s2):
                devb[1].copy_(host[1], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s2)
        else:
            devb[1].copy_(host[1], non_blocking=True)
    for _ in range(3): round_()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): round_()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    else:
        tmax = tmin = t
    res[mode] = {"gbs_per_gpu_slowest": nbytes / (float(tmax) * 1e-3) / 1e9, "gbs_per_gpu_fastest": nbytes / (float(tmin) * 1e-3) / 1e9,
                 "ms_per_step_slowest": float(tmax)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "numa_bound": bind, "bound_cpus": len(cpus) if cpus else None, "bytes_per_step": nbytes,
                      "samples_per_s_ceiling_whole_job": B * world / (res["one stream"]["ms_per_step_slowest"] * 1e-3), **res}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
