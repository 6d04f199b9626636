"""Latency of the [B,*] tail alone (modality-weight head, fusion MLP, class head, cross-entropy; MultimodalModel.py:171-199,
299-313, Trainer.py:68), forward + backward from synthetic pooled features, as a CUDA-graph replay loop -- the chain of
small dependent kernels that sits between the big GEMM phases of the step.  Prints microseconds per replay (forward only,
forward + backward), the library's launch count, and the per-kernel device times of one eager pass (mmsa_prof).

    python scripts/tail_bench.py [--batch 256] [--embed 768] [--contract single|multitask] [--iters 300]
Environment switches (A/B): MMSA_LINEAR_BN=0 (two-launch Linear + BatchNorm)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--embed", type=int, default=768)
    ap.add_argument("--contract", default="single")
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--eager-passes", type=int, default=0, help="only run this many eager passes (for an ncu launch list)")
    args = ap.parse_args()
    import mmsa
    from mmsa import _lib, ops
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, E = args.batch, args.embed
    model = mmsa.MultimodalTransformerModel(embed_dim=E, num_heads=12, wiring="bidirectional", contract=args.contract,
                                            compute_dtype=torch.bfloat16, valence=args.contract == "multitask").to(dev).train()
    feats = [torch.randn(B, E, device=dev, requires_grad=True) for _ in range(4)]          # f0, fv, e1, e2
    lps = [f.detach().to(torch.bfloat16) for f in feats[:2]]
    labels = torch.randint(0, 3, (B,), device=dev)
    one = torch.ones((), device=dev)

    def body(backward: bool):
        model.prepare_step()
        f0, fv, e1, e2 = feats
        arousal, valence = model._tail(f0, fv, (f0, e1, e2), lps[0], lps[1])
        model._drop.commit(dev)
        loss = ops.cross_entropy(arousal, labels)
        if valence is not None:
            loss = loss + ops.cross_entropy(valence, labels)
        if backward:
            for f in feats:
                f.grad = None
            model.zero_grad(set_to_none=True)
            loss.backward(gradient=one)
        return loss

    def timed(backward: bool):
        st = torch.cuda.Stream(device=dev)
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            for _ in range(3):
                body(backward)
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(g):
            body(backward)
        launches = _lib.launch_count() - n0
        for _ in range(20):
            g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / args.iters * 1e3)
        return best, launches

    if args.eager_passes:
        for _ in range(args.eager_passes):
            body(True)
        torch.cuda.synchronize()
        return
    with torch.no_grad():
        fwd_us, fwd_l = timed(False)
    both_us, both_l = timed(True)
    print(f"tail B={B} E={E} contract={args.contract} LINEAR_BN={os.environ.get('MMSA_LINEAR_BN', '1')}: "
          f"forward {fwd_us:.1f} us ({fwd_l} launches), forward+backward {both_us:.1f} us ({both_l} launches)")
    ops.set_overlap(False)
    body(True)
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    for _ in range(5):
        body(True)
    prof = _lib.prof_collect()
    _lib.prof_enable(False)
    tot = 0.0
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        us = v["ms"] / v["count"] * 1e3
        tot += v["ms"] / 5 * 1e3
        print(f"  {k:40s} x{v['count'] / 5:4.1f}  {us:6.1f} us each")
    print(f"  sum of kernel times {tot:.1f} us per pass")


if __name__ == "__main__":
    main()
