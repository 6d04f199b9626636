"""Multi-rank parity ON HARDWARE (-m gpu, needs >= 2 GPUs; skipped on a one-GPU box -- run it with `gpurun --gpus 2`):
two ranks, one process per GPU, NCCL.  Each rank runs the sharded step on the CUDA kernels -- batch-sharded InfoNCE with the
global diagonal at rank*B + i (mmsa/dist.py sharded_infonce), reduce-scatter backward of the gathered embeddings, AVG
all-reduce of the parameter gradients (GradAllReducer) -- and compares with the float64 CPU oracle of the same
data-parallel semantics: loss_r = CE over shard r (BatchNorm per shard) + InfoNCE(rows of shard r, ALL columns)
(oracle.infonce(labels2=, row_offset=), MultimodalModel.py:232-260), gradients of mean_r loss_r.
bench.py runs the same check at full size against a single-GPU evaluation (`dp_parity` in its JSON line)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200")


def _worker(rank: int, world: int, port: int, dtype_name: str, reducer_kind: str, q):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import torch.nn.functional as F
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import mmsa
        from mmsa import dist as mdist
        from parity_util import O, build_model, rel_err, zero_grad_bias_keys, grad_err
        cd = torch.float32 if dtype_name == "fp32" else torch.bfloat16
        tol = 1e-5 if dtype_name == "fp32" else 2e-2
        B, L, R = 8, 64, 49
        cfg = O.FusionConfig(embed_dim=768, num_heads=12, wiring="bidirectional", contract="single", valence=False)
        params, buffers = O.init_params(cfg, seed=4)
        params["temperature"] = torch.tensor(0.07)
        inputs, labels = O.synth_inputs(cfg, B * world, L=L, R=R, seed=99)         # every rank knows the global batch
        sl = slice(rank * B, (rank + 1) * B)
        model = build_model(cfg, params, buffers, cd, dev)
        mdist.shard_contrastive(model)
        reducer = (mdist.GradAllReducer if reducer_kind == "flat" else mdist.ArenaGradReducer)(model.parameters())
        text, image = inputs[0][sl].to(dev), inputs[1][sl].to(dev)
        lab = labels[sl].to(dev)
        logits, closs = model(text, image, None, lab)
        loss = mmsa.cross_entropy(logits, lab) + closs.sum()
        loss.backward()
        reducer.step()
        tot = loss.detach().clone()
        dist.all_reduce(tot)
        # float64 oracle of the same data-parallel semantics (and the fp32 / bf16-autocast evaluation as the noise floor)
        def oracle(dt, autocast):
            p = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in params.items()}
            xs = tuple(x.to(dt) for x in inputs)
            ctx = torch.autocast(device_type="cpu", dtype=torch.bfloat16) if autocast else torch.autocast(device_type="cpu", enabled=False)
            with ctx:
                outs = [O.fusion_forward(cfg, p, (xs[0][r * B:(r + 1) * B], xs[1][r * B:(r + 1) * B]), None, training=True)
                        for r in range(world)]
                e2_all = torch.cat([o.feats["slot2"] for o in outs])
                total, per_rank = 0.0, []
                for r, o in enumerate(outs):
                    lr = labels[r * B:(r + 1) * B]
                    c = O.infonce(o.feats["slot1"].float() if autocast else o.feats["slot1"], e2_all.float() if autocast else e2_all,
                                  lr, p["temperature"], labels2=labels, row_offset=r * B)
                    l_r = F.cross_entropy(o.arousal.float(), lr) + (p["contrastive_weight"] * c).sum()
                    per_rank.append((l_r.detach(), o.arousal.detach()))
                    total = total + l_r
                total = total / world
            total.backward()
            return total.detach(), per_rank, {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
        t64, pr64, g64 = oracle(torch.float64, False)
        t32, pr32, g32 = oracle(torch.float32, dtype_name == "bf16")
        zero_keys = zero_grad_bias_keys(g64.keys())
        nm = 3.0 if dtype_name == "fp32" else 1.0
        bad = []

        def chk(name, err, noise):
            if err > max(tol, nm * noise):
                bad.append((name, err, noise))
        chk("mean loss", abs(float(tot) / world - float(t64)) / abs(float(t64)), abs(float(t32) - float(t64)) / abs(float(t64)))
        chk("local loss", rel_err(loss, pr64[rank][0]), rel_err(pr32[rank][0], pr64[rank][0]))
        chk("logits", rel_err(logits, pr64[rank][1]), rel_err(pr32[rank][1], pr64[rank][1]))
        for k, prm in model.named_parameters():
            chk("grad:" + k, grad_err(k, prm.grad, g64[k], g64, zero_keys), grad_err(k, g32[k], g64[k], g64, zero_keys))
        labels_ok = bool(torch.equal(logits.argmax(1).cpu(), pr64[rank][1].argmax(1)))
        q.put((rank, not bad and labels_ok, bad[:5]))
    except Exception as ex:      # surface the failure instead of a queue timeout
        import traceback
        q.put((rank, False, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("reducer_kind", ["arena", "flat"])
@pytest.mark.parametrize("dtype_name", ["fp32", "bf16"])
def test_two_rank_sharded_step_against_oracle(dtype_name, reducer_kind):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, dtype_name, reducer_kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in results), results
