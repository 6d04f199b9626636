"""world_size-2 gloo tests (CPU) of the data-parallel host logic: all-gather with reduce-scatter
backward, the sharded InfoNCE row-block convention, and the flat-bucket gradient all-reduce.
The device kernels are not involved: the arithmetic here is the oracle's, the plumbing is mmsa.dist."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200")


def _worker(rank: int, world: int, port: int, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmsa import dist as mdist
        from oracle import fusion_oracle as O
        torch.manual_seed(0)
        Bg, E = 8, 16
        B = Bg // world
        f1 = torch.randn(Bg, E, dtype=torch.float64)
        f2 = torch.randn(Bg, E, dtype=torch.float64)
        labels = torch.randint(0, 3, (Bg,))
        T = torch.tensor(0.07, dtype=torch.float64)
        # global reference (single process)
        a, b = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
        ref = O.infonce(a, b, labels, T)
        ref.backward()
        # sharded: this rank owns rows [rank*B, (rank+1)*B)
        sl = slice(rank * B, (rank + 1) * B)
        x, y = f1[sl].clone().requires_grad_(True), f2[sl].clone().requires_grad_(True)
        y_all = mdist.all_gather_rows(y)
        lab_all = mdist.gather_labels(labels[sl])
        assert torch.equal(lab_all, labels)
        loss = O.infonce(x, y_all, labels[sl], T, labels2=lab_all, row_offset=rank * B)
        loss.backward()
        # mean over ranks of the local means == global mean; gradients: local rows exact / world,
        # gathered columns reduce-scattered (sum over ranks) / world
        tot = loss.detach().clone()
        dist.all_reduce(tot)
        ok = abs(float(tot) / world - float(ref)) < 1e-12
        ok &= bool(torch.allclose(x.grad / world, a.grad[sl], atol=1e-12))
        ok &= bool(torch.allclose(y.grad / world, b.grad[sl], atol=1e-12))
        # flat-bucket gradient all-reduce (mean)
        lin = torch.nn.Linear(4, 3)
        lin2 = torch.nn.Linear(3, 2)
        with torch.no_grad():
            for p_ in list(lin.parameters()) + list(lin2.parameters()):
                p_.grad = torch.full_like(p_, float(rank + 1))
        lin2.bias.grad = None                                   # a rank-local missing grad is treated as zero
        red = mdist.GradAllReducer(list(lin.parameters()) + list(lin2.parameters()), bucket_mb=1e-5)
        red.step()
        want = sum(range(1, world + 1)) / world
        ok &= all(bool(torch.allclose(p_.grad, torch.full_like(p_, want))) for p_ in lin.parameters())
        ok &= bool(torch.allclose(lin2.bias.grad, torch.zeros_like(lin2.bias)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_contrastive_and_grad_allreduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(ok for _, ok in results), results


def _overlap_worker(rank, world, port, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmsa import dist as mdist
        torch.manual_seed(0)                                   # identical replicas
        net = torch.nn.Sequential(torch.nn.Linear(40, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                  torch.nn.Linear(64, 8), torch.nn.Linear(8, 3, bias=False))
        unused = torch.nn.Parameter(torch.zeros(5))            # never gets a gradient
        import copy
        ref = copy.deepcopy(net)                               # same weights, no hooks: the yardstick
        params = list(net.parameters()) + [unused]
        red = mdist.GradAllReducer(params, overlap=True, late_frac=0.3)
        ok = True
        for it in range(4):
            g = torch.Generator().manual_seed(100 + it)
            xs = torch.randn(world, 6, 40, generator=g)            # every rank knows every shard
            for p_ in params:
                p_.grad = None
            net(xs[rank]).square().sum().backward()
            red.step()
            ref.zero_grad(set_to_none=True)
            for r in range(world):                              # sum over shards of the per-shard gradients
                ref(xs[r]).square().sum().backward()
            ok &= all(bool(torch.allclose(a.grad, b.grad / world, rtol=1e-5, atol=1e-6))
                      for a, b in zip(net.parameters(), ref.parameters()))
            ok &= unused.grad is None
        # first step records the landing order, the other three reduce the early bucket from inside the backward
        ok &= red.overlapped_steps == 3
        # the late bucket is the first layer's weight (its gradient lands last), the early one everything else
        ok &= red._early is not None and 0 not in red._early and len(red._early) == len(params) - 2
        # gradients are views of ONE flat arena in parameter order (FusedClipAdamW's in-place path)
        base = red._flat.data_ptr()
        offs = [p_.grad.data_ptr() - base for p_ in net.parameters()]
        ok &= offs == sorted(offs) and all(0 <= o < red._flat.numel() * 4 for o in offs)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_overlapped_two_bucket_grad_allreduce_gloo():
    """GradAllReducer(overlap=True): landing-order recording, early bucket reduced from the post-accumulate hooks,
    late bucket in step(), results equal to the mean of the per-shard gradients."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(ok for _, ok in results), results


def _arena_worker(rank, world, port, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmsa import dist as mdist
        torch.manual_seed(0)
        shapes = [(6, 5), (6,), (4, 6), (4,), (3, 4), (3,), (2, 3), (1,)]        # "proj", "block1", "block2", "tail" pairs
        params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
        red = mdist.ArenaGradReducer(params, register_sink=False, early_buckets=True)
        single = mdist.ArenaGradReducer([torch.nn.Parameter(torch.zeros(s)) for s in shapes], register_sink=False)   # default: one all-reduce
        ok = True
        for it in range(3):
            g = torch.Generator().manual_seed(50 + it)
            grads = [[torch.randn(s, generator=g) for s in shapes] for _ in range(world)]      # every rank knows every shard
            for p_ in params:
                p_.grad = None
            # tail (last two parameters) arrives through autograd BEFORE the core's backward starts
            params[6].grad = grads[rank][6].clone()
            if it != 1:                                   # one step without a gradient for the last parameter
                params[7].grad = grads[rank][7].clone()
            # the kernels write block1 / block2 / proj straight into the arena views and report them group by group
            for grp in ((2, 3), (4, 5), (0, 1)):
                for i in grp:
                    red.sink(params[i]).copy_(grads[rank][i])
                red.bucket_done([params[i] for i in grp])
            red.step()
            for i, p_ in enumerate(params):
                if i == 7 and it == 1:
                    ok &= p_.grad is None
                    continue
                want = sum(grads[r][i] for r in range(world)) / world
                ok &= bool(torch.allclose(p_.grad, want, rtol=1e-6, atol=1e-7))
                ok &= p_.grad.data_ptr() == red.views[i].data_ptr()        # .grad IS the arena view (optimiser in-place path)
            # launch order: tail rides with the first report, then block1, block2, proj; nothing is left for step()
            first = red.last_launched[0]
            ok &= first[0] == red.offs[6]
            ok &= red.early_bytes == sum((hi - lo) * 4 for lo, hi in red.last_launched)
            ok &= len(red.last_launched) == 4
            # default form: the same gradients, nothing leaves before step(), one all-reduce over the whole arena
            for i, p_ in enumerate(single.params):
                p_.grad = None
                if i >= 6:
                    if not (i == 7 and it == 1):
                        p_.grad = grads[rank][i].clone()
                else:
                    single.sink(p_).copy_(grads[rank][i])
            single.bucket_done(single.params[:6])
            ok &= single.launched == []
            single.step()
            ok &= single.early_bytes == 0 and len(single.last_launched) == (1 if it != 1 else 1)
            for i, p_ in enumerate(single.params):
                if i == 7 and it == 1:
                    continue
                ok &= bool(torch.allclose(p_.grad, params[i].grad, rtol=0, atol=0))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_arena_grad_reducer_buckets_gloo():
    """ArenaGradReducer: gradients written straight into the arena (sink protocol), groups all-reduced as they are reported
    in landing order (tail, block e2p, block p2e, projections), autograd-delivered tail packed, a missing gradient left
    alone, .grad pointed at the arena views; equal to the mean of the per-rank gradients."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_arena_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(ok for _, ok in results), results
