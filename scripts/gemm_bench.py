"""GEMM microbenchmark (GPU box): every Linear product shape of the configs[1] step through the C ABI,
timed with CUDA events, rotating over buffer sets larger than L2.  Prints TFLOP/s per shape."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))
import torch
from mmsa import kernels as K

dev = torch.device("cuda:0")
B, L, R, E = 256, 128, 49, 768
Mt, Mv = B * L, B * R
shapes = {
    "fwd":   [(Mt, 768, 768), (Mv, 768, 2048), (Mv, 1536, 768), (Mt, 1536, 768), (Mv, 768, 768), (Mt, 768, 1536), (Mv, 768, 1536),
              (256, 256, 2304), (256, 128, 256)],
    "dgrad": [(Mt, 768, 768), (Mv, 768, 768), (Mt, 1536, 768), (Mv, 1536, 768)],
    "wgrad": [(Mt, 768, 768), (Mv, 768, 2048), (Mt, 1536, 768), (Mv, 1536, 768), (Mv, 768, 768), (256, 256, 2304)],
}
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(os.environ.get("REPS", "20"))
res = []
for kind, lst in shapes.items():
    if only and kind != only:
        continue
    for (M, N, Kd) in lst:
        nset = max(2, int(200e6 // ((M * Kd + M * N) * 2)) + 1)
        xs = [torch.randn(M, Kd, device=dev).bfloat16() for _ in range(nset)]
        w = (torch.randn(N, Kd, device=dev) / Kd ** 0.5).bfloat16()
        bias = torch.randn(N, device=dev)
        dys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(nset)]
        outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nset)]
        dxs = [torch.empty(M, Kd, device=dev, dtype=torch.bfloat16) for _ in range(nset)]
        def run(i):
            j = i % nset
            if kind == "fwd":
                K.linear_fwd(xs[j], w, bias, out=outs[j])
            elif kind == "dgrad":
                K.linear_dgrad(dys[j], w, out=dxs[j])
            else:
                K.linear_wgrad(dys[j], xs[j])
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = 2.0 * M * N * Kd / (ms * 1e-3) / 1e12
        res.append({"kind": kind, "M": M, "N": N, "K": Kd, "ms": ms, "tflops": tf})
        print(f"{kind:6s} M={M:6d} N={N:5d} K={Kd:5d}  {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
        del xs, dys, outs, dxs
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_bench.json"), "w"), indent=1)
