"""Probe (torchrun, N GPUs): device time of the step's collectives alone (CUDA events, max over ranks)."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
flat = torch.randn(10_180_000, device=dev)
e2 = torch.randn(256, 768, device=dev); lab = torch.zeros(256, dtype=torch.int64, device=dev)
e2_all = torch.empty(256 * world, 768, device=dev); lab_all = torch.empty(256 * world, dtype=torch.int64, device=dev)
dx = torch.empty(256, 768, device=dev)
def t(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    v = torch.tensor([a.elapsed_time(b) / n], device=dev); dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v) * 1e3
r = {"all_reduce 40.7MB AVG": t(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG)),
     "all_gather 786KB": t(lambda: dist.all_gather_into_tensor(e2_all, e2)),
     "all_gather labels": t(lambda: dist.all_gather_into_tensor(lab_all, lab)),
     "reduce_scatter 786KB": t(lambda: dist.reduce_scatter_tensor(dx, e2_all))}
def graph_time(fn, reps=10):
    """device time per call of `fn` when `reps` calls are captured in one CUDA graph (as TrainStep captures the step)"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    torch.cuda.synchronize()
    return t(lambda: g.replay(), n=10) / reps
r["graph: all_reduce 40.7MB AVG"] = graph_time(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))
r["graph: all_gather 786KB"] = graph_time(lambda: dist.all_gather_into_tensor(e2_all, e2))
r["graph: reduce_scatter 786KB"] = graph_time(lambda: dist.reduce_scatter_tensor(dx, e2_all))
scratch = torch.randn(64 << 20, device=dev)
def mixed():
    scratch.mul_(1.0001)                      # ~80 us of independent device work between collectives
    dist.all_gather_into_tensor(e2_all, e2)
r["graph: [256 MB mul_ + all_gather]"] = graph_time(mixed)
r["graph: [256 MB mul_] alone"] = graph_time(lambda: scratch.mul_(1.0001))
if rank == 0:
    for k, v in r.items(): print(f"{k:40s} {v:8.1f} us")
dist.destroy_process_group()
