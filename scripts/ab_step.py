"""In-process A/B of the configs[1] step (B=256, L=128, bf16, CUDA graph): two captured graphs of the SAME model that differ
in one module-level switch of mmsa.ops, timed in alternating blocks on the same GPU in the same minute -- separate bench.py
runs differ by +-2 % on their own (box, thermal state), which hides changes of a few tens of microseconds.

    python scripts/ab_step.py --switch FUSE_LINEAR_BN [--blocks 8] [--steps 100]
prints the per-block times of A (switch = False) and B (switch = True) and the mean difference."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--switch", default="FUSE_LINEAR_BN")
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--L", type=int, default=128)
    args = ap.parse_args()
    import mmsa
    from mmsa import ops
    from mmsa.step import TrainStep
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = mmsa.MultimodalTransformerModel(num_classes=3, embed_dim=768, num_heads=12, wiring="bidirectional", text_dim=768,
                                            image_dim=2048, contract="single", compute_dtype=torch.bfloat16,
                                            valence=False).to(dev).train()
    g = torch.Generator().manual_seed(1)
    text = torch.randn(args.batch, args.L, 768, generator=g).to(torch.bfloat16)
    image = torch.randn(args.batch, 49, 2048, generator=g).to(torch.bfloat16)
    labels = torch.randint(0, 3, (args.batch,), generator=g)
    steps = {}
    for val in (False, True):
        setattr(ops, args.switch, val)
        st = TrainStep(model, args.batch, args.L, 49, 768, 2048, feature_dtype=torch.bfloat16, n_slots=2, device=dev)
        for sl in st.slots:
            sl.text.copy_(text); sl.image.copy_(image); sl.labels.copy_(labels)
        st.warmup(2)
        st.capture()
        steps[val] = st
    res = {False: [], True: []}
    for blk in range(args.blocks + 1):
        for val in (False, True) if blk % 2 == 0 else (True, False):
            st = steps[val]
            for i in range(10):
                st.run(i & 1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                st.run(i & 1)
            e1.record()
            torch.cuda.synchronize()
            if blk:                                   # block 0 is warm-up
                res[val].append(e0.elapsed_time(e1) / args.steps * 1e3)
    a, b = res[False], res[True]
    ma, mb = sum(a) / len(a), sum(b) / len(b)
    print(f"{args.switch}=False: {' '.join(f'{x:.0f}' for x in a)}  mean {ma:.1f} us  ({steps[False].launches_per_step} launches)")
    print(f"{args.switch}=True : {' '.join(f'{x:.0f}' for x in b)}  mean {mb:.1f} us  ({steps[True].launches_per_step} launches)")
    print(f"True - False = {mb - ma:+.1f} us per step")


if __name__ == "__main__":
    main()
