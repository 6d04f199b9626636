#!/usr/bin/env python
"""bench.py -- fusion + ME-MHACL fwd+bwd throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU reference arm

A step = one forward + loss + backward of the hot path over one synthetic batch:
text [B,L,768] and image [B,49,2048] features -> projection GEMMs -> two cross-modal 12-head
attention blocks (gate + LayerNorm) -> token mean-pool -> modality-weight softmax + concat ->
fusion MLP (BatchNorm/GELU/dropout) -> 3-class head -> CE, plus the image-text InfoNCE (learnable
temperature) -- and every gradient back to the parameters.  Workload at N=1: BASELINE.json
configs[1] (L=128, batch 256, bf16).  N>1: the same per-GPU batch on every rank (weak scaling),
InfoNCE columns all-gathered over NCCL, parameter gradients all-reduced.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  value      device-resident throughput (inputs already in HBM, CUDA-graph replay of the step)
  e2e        same metric through mmsa.TrainStep with HOST (pinned) inputs: H2D of the step's
             features + labels and D2H of the loss inside the timed region
  roofline   tensor roofline of the dominant kernel (gemm_tcgen05): algorithmic FLOPs of all its
             launches in a step / their summed device time, measured live with CUDA events
  cpu_baseline  the CPU oracle (port of the reference arithmetic) on this box's host cores
  sustained  the same graph replayed back to back for >= 3 s: ms/step, SM clock and throttle reasons under load
  torch_eager   (N=1) the reference arithmetic run by PyTorch itself on this B200 -- the oracle's functions on cuda, fp32
             and torch.autocast(bf16) (cuBLASLt + ATen), CUDA-event timed -- plus torch.matmul (cuBLAS) on every GEMM shape
             of the step next to this library's kernel on the same shape: SURVEY section 2.1's "kernel to beat"
  dp_parity  (N>1) before timing: one sharded step (dropout off) against the same step evaluated on ONE GPU over the
             gathered global batch (mmsa.dist.emulate_data_parallel_step); the run fails if they disagree
  configs    further BASELINE.json configurations measured in the same launch (configs[2], [3], [4] where they apply)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multimodal-sentiment-aanalysis_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

E, H, R, DT, DI, C = 768, 12, 49, 768, 2048, 3
METRIC = "fusion_fwd_bwd_samples_per_s"
UNIT = "samples/s"


def algorithmic_gflop_per_sample(L: int, Bg: int, Ec: int = E) -> dict:
    """per-sample forward FLOPs (multiply-add = 2) and the canonical fwd+bwd figure."""
    P = 2 * L * DT * E + 2 * R * DI * E
    Xt = 2 * L * E * E + 4 * R * E * E + 4 * L * R * E + 2 * L * E * E + 4 * L * E * E
    Xi = 2 * R * E * E + 4 * L * E * E + 4 * R * L * E + 2 * R * E * E + 4 * R * E * E
    T = 2 * (2 * E * 64 + 64 * 3) + 2 * (3 * E * 256 + 256 * 128) + 2 * (128 * 128 + 128 * C)
    N = 2 * Bg * Ec
    fwd = P + Xt + Xi + T + N
    return {"fwd": fwd / 1e9, "fwd_bwd": (2 * P + 3 * (Xt + Xi + T + N)) / 1e9}


def load_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


# ----------------------------------------------------------------------------- clock sampling
class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
            return self
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def _run(self):
        nv = self._nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    r = int(get_reasons(self._h))
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self) -> dict:
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        med = int(statistics.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_samples_per_s(L: int, sample_batch: int, steps: int, warmup: int, budget_s: float = 25.0) -> dict:
    """fwd+bwd of the oracle (CPU restatement of the reference arithmetic, fp32) on a bounded
    sample of the workload (`sample_batch` samples of the same shapes), all host threads."""
    import torch
    from oracle import fusion_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.FusionConfig(embed_dim=E, num_heads=H, wiring="bidirectional", text_dim=DT, image_dim=DI,
                         contract="single", valence=False)
    params, _ = O.init_params(cfg, seed=0)
    inputs, labels = O.synth_inputs(cfg, sample_batch, L=L, R=R, seed=1234)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    times = []
    t_begin = time.perf_counter()
    for it in range(warmup + steps):
        for v in p.values():
            v.grad = None
        t0 = time.perf_counter()
        loss, _ = O.trainer_loss(cfg, p, inputs, labels, training=True)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 2:
            break
    ms = statistics.median(times) * 1e3
    return {"value": sample_batch / (ms / 1e3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_batch} samples (of the {L}-token workload) x {len(times)} timed steps, fp32, "
                      f"torch {torch.__version__} CPU, median {ms:.1f} ms/step",
            "ms_per_step": ms, "steps": len(times)}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_samples_per_s(args.L, args.cpu_sample_batch, max(args.steps, 2), args.warmup)
    g = algorithmic_gflop_per_sample(args.L, args.batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: ME-MHACL cross-modal 12-head attention fusion fwd+bwd, L={args.L}, "
                               f"E={E}, R={R}; CPU sample of {args.cpu_sample_batch} samples per step",
                   "gflop_per_sample_fwd_bwd": g["fwd_bwd"]},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, BEFORE the pinned host buffers are
    allocated (first touch places them on that NUMA node): at N >= 4 every rank streams ~100 MB of features per step
    over PCIe, and a buffer on the far socket halves that rate.  Returns the CPU list, or None when unavailable."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        handle = None
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------- torch eager on the same GPU (baseline)
def torch_eager_baseline(B: int, L: int, dev, steps: int = 5) -> dict:
    """The reference arithmetic as PyTorch runs it on this GPU (`model.to('cuda')`, main.py:25): the oracle's functional
    restatement of MultimodalModel.py:139-149,232-260 with cuBLASLt/ATen kernels, fwd+bwd, fp32 and bf16 autocast.
    A baseline measurement (like cpu_baseline): nothing of it is on the product path."""
    import torch
    from oracle import fusion_oracle as O
    cfg = O.FusionConfig(embed_dim=E, num_heads=H, wiring="bidirectional", text_dim=DT, image_dim=DI,
                         contract="single", valence=False)
    params, _ = O.init_params(cfg, seed=0)
    inputs, labels = O.synth_inputs(cfg, B, L=L, R=R, seed=1234)
    p = {k: v.to(dev).requires_grad_(True) for k, v in params.items()}
    xs = tuple(x.to(dev) for x in inputs)
    lab = labels.to(dev)
    out = {}
    for mode in ("fp32", "bf16_autocast"):
        xin = xs if mode == "fp32" else tuple(x.bfloat16() for x in xs)

        def one():
            for v in p.values():
                v.grad = None
            if mode == "fp32":
                loss, _ = O.trainer_loss(cfg, p, xin, lab, training=True)
            else:
                with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                    loss, _ = O.trainer_loss(cfg, p, xin, lab, training=True)
            loss.backward()
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3)}
    out["what"] = ("oracle/fusion_oracle.trainer_loss on cuda (torch eager: cuBLASLt + ATen kernels, P materialised), "
                   f"fwd+bwd, B={B}, L={L}, BatchNorm in train mode, the three [B,*] dropout layers as identity, no CUDA graph, "
                   f"{steps} timed steps")
    return out


def cublas_matmul_tflops(prof: dict, dev, prof_steps: int, reps: int = 20) -> dict:
    """torch.matmul (cuBLAS, bf16 in / fp32 accumulate) on every tcgen05 GEMM shape the step launched, same operand majors
    (kk: X W^T, km: dY W, mm: dY^T X), timed back to back with CUDA events over operand sets larger than L2; next to it this
    library's kernel on that shape timed the SAME way through the C ABI (mmsa_b2b_tflops; bf16 output, no bias) and its
    TFLOP/s inside the profiled step (mmsa_in_step_tflops: per-launch events, includes the launch gap)."""
    import torch
    from mmsa import kernels as mk
    out = {}
    for name, v in prof.items():
        if not name.startswith("gemm_tc_"):
            continue
        maj, dims = name[len("gemm_tc_"):].split("_")
        M, N, K = (int(x) for x in dims.split("x"))
        if 2.0 * M * N * K < 1e9:
            continue
        nset = max(2, int(200e6 // ((M * K + N * K) * 2)) + 1)       # rotate operand sets larger than L2, as in a step
        As = [torch.randn((K, M) if maj[0] == "m" else (M, K), device=dev, dtype=torch.bfloat16) for _ in range(nset)]
        Bs = [torch.randn((K, N) if maj[1] == "m" else (N, K), device=dev, dtype=torch.bfloat16) for _ in range(nset)]
        As = [a.t() if maj[0] == "m" else a for a in As]
        Bs = [b if maj[1] == "m" else b.t() for b in Bs]
        for i in range(3):
            torch.matmul(As[i % nset], Bs[i % nset])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            torch.matmul(As[i % nset], Bs[i % nset])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        # the same product through this library (fwd: x W^T; dgrad: dy W; wgrad: dy^T x), same rotation
        raw_a = [a.t() if maj[0] == "m" else a for a in As]          # storage as allocated: [K,M] or [M,K]
        raw_b = [b if maj[1] == "m" else b.t() for b in Bs]          # [K,N] or [N,K]

        def ours(i):
            a, b = raw_a[i % nset], raw_b[i % nset]
            if maj == "kk":
                mk.linear_fwd(a, b, None)
            elif maj == "km":
                mk.linear_dgrad(a, b)
            else:
                mk.linear_wgrad(a, b, want_bias=False)
        ms_o = None
        try:
            for i in range(3):
                ours(i)
            e0.record()
            for i in range(reps):
                ours(i)
            e1.record()
            torch.cuda.synchronize()
            ms_o = e0.elapsed_time(e1) / reps
        except Exception:
            pass
        del As, Bs, raw_a, raw_b
        out[name[len("gemm_tc_"):]] = {"cublas_tflops": 2.0 * M * N * K / (ms * 1e-3) / 1e12,
                                       "mmsa_b2b_tflops": (2.0 * M * N * K / (ms_o * 1e-3) / 1e12) if ms_o else None,
                                       "mmsa_in_step_tflops": v["work"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else None,
                                       "launches_per_step": v["count"] / prof_steps}
    return out


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    import mmsa
    from mmsa import _lib
    from mmsa import dist as mdist
    from mmsa import ops as _ops
    from mmsa.step import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # (NCCL_MAX_CTAS is deliberately left alone: capping NCCL's CTAs to 4 / 8 / 16 made the step 12 / 6 / 1.5 % slower
        #  at N = 2 and 7 % slower at N = 8 with 16 -- profiles/r02_dp_matrix.md)
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    if _lib.load().mmsa_check_device() != 0:
        raise RuntimeError(_lib.load().mmsa_last_error().decode())

    cd = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B, L = args.batch, args.L
    torch.manual_seed(0)
    model = mmsa.MultimodalTransformerModel(num_classes=C, embed_dim=E, num_heads=H, wiring="bidirectional",
                                            text_dim=DT, image_dim=DI, contract="single", compute_dtype=cd,
                                            valence=False).to(dev).train()
    reducer = None
    ablate = set(filter(None, os.environ.get("MMSA_BENCH_ABLATE", "").split(",")))
    if world > 1:
        for p in model.parameters():            # replicas start identical
            dist.broadcast(p.data, 0)
        # MMSA_BENCH_ABLATE (comma list: "shard", "reduce") switches a collective off to attribute the multi-GPU
        # overhead; such a run is a diagnostic, flagged in config.ablate, never a result
        if "shard" not in ablate:
            mdist.shard_contrastive(model)
        if "reduce" not in ablate:
            # default: gradients land in the flat arena (no pack) and leave in ONE all-reduce after the last weight gradient;
            # MMSA_DP_BUCKETS=1 sends them in landing-order buckets under the backward instead (measured slower on this
            # fabric: the NCCL kernels take SMs from the persistent GEMMs), MMSA_DP_REDUCER=flat is round 1's pack-and-reduce
            if os.environ.get("MMSA_DP_REDUCER", "arena") == "flat":
                reducer = mdist.GradAllReducer(model.parameters())
            else:
                reducer = mdist.ArenaGradReducer(model.parameters(),
                                                 early_buckets=os.environ.get("MMSA_DP_BUCKETS", "0") == "1")
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def host_batches(Bc: int, Lc: int, n: int = 2):
        """synthetic features: N(0,1), seeded per rank; host copies pinned, in the feature dtype"""
        g = torch.Generator().manual_seed(1234 + rank)
        return [(torch.randn(Bc, Lc, DT, generator=g).to(cd).pin_memory(),
                 torch.randn(Bc, R, DI, generator=g).to(cd).pin_memory(),
                 torch.randint(0, C, (Bc,), generator=g).pin_memory()) for _ in range(n)]

    # ---- data-parallel parity (N > 1), BEFORE anything is timed ------------------------------------------------------
    dp_parity = None
    if world > 1 and not ablate:
        dp_parity = dp_parity_check(model, reducer, host_batches(B, L, 1)[0], cd, dev, rank, world)
        if not dp_parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "dp_parity": dp_parity, "error": "data-parallel step disagrees with "
                                  "the single-GPU evaluation of the same global batch"}), flush=True)
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(3)

    opt = None
    if args.optimizer:          # fused global-norm clip + AdamW after every step (Trainer.py:80-81)
        # built BEFORE the step graph is captured: the constructor re-homes the parameters into its flat arena
        opt = mmsa.FusedClipAdamW(model.parameters(), lr=1e-4, weight_decay=0.01, max_norm=1.0)

    def build_step(Bc: int, Lc: int, host):
        step = TrainStep(model, Bc, Lc, R, DT, DI, feature_dtype=cd, n_slots=2, use_graph=not args.no_graph,
                         post_backward=(reducer.step if reducer is not None else None), device=dev)
        for k, sl in enumerate(step.slots):
            sl.text.copy_(host[k % len(host)][0]); sl.image.copy_(host[k % len(host)][1]); sl.labels.copy_(host[k % len(host)][2])
        step.warmup(2)
        ok = not args.no_graph
        if ok:
            try:
                step.capture()
            except Exception as ex:                  # e.g. a collective that cannot be captured
                if rank == 0:
                    print(f"bench.py: CUDA-graph capture failed ({type(ex).__name__}: {ex}); timing eager launches",
                          file=sys.stderr)
                ok = False
                for sl in step.slots:
                    sl.graph = None
                torch.cuda.synchronize()
        return step, ok

    def timed_replays(step, n_steps: int, n_warm: int, sample_clocks: bool = False, min_seconds: float = 0.0):
        """device time per step (CUDA events around n_steps back-to-back steps, barrier + synchronize on both sides, max over
        ranks).  min_seconds > 0: keep going in blocks of n_steps until that much wall time has passed (sustained leg)."""
        for _ in range(n_warm):
            for k in (0, 1):
                step.run(k)
                if opt is not None:
                    opt.step()
        # the clock sampler (NVML init: tens of ms) starts BEFORE the barrier: anything rank 0 does between the barrier and
        # its first launch shows up as a stall inside the other ranks' first collective
        sampler = ClockSampler(local_rank).start() if (sample_clocks and rank == 0) else None
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done = 0
        t_wall = time.perf_counter()
        ev0.record()
        while True:
            for i in range(n_steps):
                step.run(i & 1)
                if opt is not None:
                    opt.step()
            done += n_steps
            if min_seconds <= 0:
                break
            torch.cuda.current_stream().synchronize()      # sustained leg only: bound the launch queue, decide on wall time
            flag = torch.tensor([1.0 if time.perf_counter() - t_wall >= min_seconds else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if float(flag.item()) > 0:
                break
        ev1.record()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1) / done)
        clocks = sampler.stop() if sampler is not None else None
        return ms, done, clocks

    # ---- device-resident throughput (`value`) ------------------------------------------------------------------------
    host = host_batches(B, L)
    step, graph_ok = build_step(B, L, host)
    n0 = _lib.launch_count()
    ms_step, _, clocks = timed_replays(step, args.steps, max(args.warmup, 3), sample_clocks=True)
    launches = (step.launches_per_step * args.steps + (_lib.launch_count() - n0)) if graph_ok else (_lib.launch_count() - n0)
    loss_val = float(step.slots[0].loss.item())

    # ---- sustained leg: the same graph back to back for >= 3 s (clocks settle to what the power limit allows) ----------
    sustained = None
    if args.sustained_seconds > 0:
        ms_sus, n_sus, clk_sus = timed_replays(step, max(args.steps, 100), 1, sample_clocks=True,
                                               min_seconds=args.sustained_seconds)
        sustained = {"seconds": ms_sus * n_sus * 1e-3, "steps": n_sus, "ms_per_step": ms_sus,
                     "value": B * world / (ms_sus * 1e-3), "unit": UNIT, "clocks": clk_sus}

    # ---- end-to-end through the public step API with host inputs (`e2e`) ----------------------------------------------
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    in_ready = [torch.cuda.Event() for _ in step.slots]
    slot_free = [torch.cuda.Event() for _ in step.slots]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def e2e_loop(n: int):
        for ev in slot_free:
            ev.record(main)
        for i in range(n):
            k = i & 1
            s = step.slots[k]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(slot_free[k])           # previous step on this slot has finished
                s.text.copy_(host[k][0], non_blocking=True)
                s.image.copy_(host[k][1], non_blocking=True)
                s.labels.copy_(host[k][2], non_blocking=True)
                in_ready[k].record(copy_stream)
            main.wait_event(in_ready[k])
            loss = step.run(k)
            if opt is not None:
                opt.step()
            loss_host.copy_(loss, non_blocking=True)           # D2H read of the step's loss
            slot_free[k].record(main)
        main.synchronize()

    e2e_loop(max(args.warmup, 3))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / args.steps)

    # ---- further BASELINE.json configurations in the same launch ------------------------------------------------------
    extra = {}
    if not args.no_extra_configs and args.batch == 256 and args.L == 128:
        plan = []
        if world > 1 and 4096 % world == 0:
            plan.append(("configs[2]", 4096 // world, 128, "ME-MHACL + image-text InfoNCE, global batch 4096 sharded over "
                         f"{world} GPUs with embedding all-gather (full fwd+bwd step, grads all-reduced)"))
        if world == 8:
            plan.append(("configs[3]", 1024, 128, "full fusion-head train step (attention + CE + contrastive + grad all-reduce) "
                         "at global batch 8192 on 8 GPUs"))
        if world in (1, 8):
            plan.append(("configs[4]", 1024, 512, f"long-text sweep L=512 x 49 regions, per-GPU batch 1024, on {world} GPU(s)"))
        for key, Bc, Lc, what in plan:
            del step
            torch.cuda.empty_cache()
            hostc = host_batches(Bc, Lc)
            step, okc = build_step(Bc, Lc, hostc)
            n_c = max(4, min(args.steps, int(100.0 / (0.0075 * Bc * Lc / 128 + 0.5))))      # ~0.1 s of timed steps
            ms_c, _, clk_c = timed_replays(step, n_c, 2, sample_clocks=True)
            gfc = algorithmic_gflop_per_sample(Lc, Bc * world)
            extra[key] = {"workload": what, "per_gpu_batch": Bc, "global_batch": Bc * world, "L": Lc, "steps": n_c,
                          "ms_per_step": ms_c, "value": Bc * world / (ms_c * 1e-3), "unit": UNIT, "cuda_graph": okc,
                          "step_tflops_per_gpu": Bc / (ms_c * 1e-3) * gfc["fwd_bwd"] / 1e3,
                          "clocks": clk_c, "loss": float(step.slots[0].loss.item())}
            del hostc
        if plan:                                   # back to the headline workload for the profiling pass
            del step
            torch.cuda.empty_cache()
            step, graph_ok2 = build_step(B, L, host)

    # ---- eager pass with per-launch CUDA events: kernel breakdown + roofline of the dominant kernel ----
    for s in step.slots:
        s.graph = None
    _ops.set_overlap(False)          # one kernel at a time: per-launch times must not include a co-running kernel
    step.run(0)
    torch.cuda.synchronize()
    prof_steps = min(args.steps, 5)
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for i in range(prof_steps):
        step.run(i & 1)
    ee1.record()
    torch.cuda.synchronize()
    ms_ungraphed = ee0.elapsed_time(ee1) / prof_steps
    # Per-launch CUDA events only measure kernel time if the GPU never waits for the host: in eager mode the host
    # needs longer to enqueue a step than the GPU to run it, so each profiled step is queued behind a device-side
    # delay (torch.cuda._sleep) long enough for the host to get a whole step ahead.
    delay_cycles = int((ms_ungraphed + 2.0) * 1e-3 * 2.0e9)
    _lib.prof_enable(True)
    for i in range(prof_steps):
        torch.cuda._sleep(delay_cycles)
        step.run(i & 1)
    _lib.prof_enable(False)
    prof = _lib.prof_collect()
    _ops.set_overlap(True)
    kern = {}
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        kern[name] = {"launches_per_step": v["count"] / prof_steps, "ms_per_step": v["ms"] / prof_steps,
                      "share": v["ms"] / tot_ms, "work_per_step": v["work"] / prof_steps}
    gemm = {"count": 0, "ms": 0.0, "work": 0.0}
    for name, v in prof.items():           # tcgen05 GEMM launches are booked per shape: gemm_tc_<majors>_<MxNxK>
        if name.startswith("gemm_tc_"):
            for k in gemm:
                gemm[k] += v[k]
    # Which peak: the profiled launches run one at a time with idle gaps between steps, i.e. at boost clocks (like the
    # 20-step timed region, ~0.04 s) -> the BURST figure of MEASURED_PEAKS.json; the sustained leg above is quoted
    # against the sustained figure.
    burst_peak = peaks.get("bf16_tflops", 1590.0)
    sus_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    roofline = None
    if gemm["ms"] > 0:
        achieved = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12
        traffic, traffic_src = None, None      # DRAM bytes per launch from the committed ncu capture of this workload
        try:
            import glob
            files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic.json")))
            if files and B == 256 and L == 128 and args.dtype == "bf16":
                with open(files[-1]) as f:
                    tj = json.load(f)
                traffic, traffic_src = tj["traffic_bytes_per_launch"], os.path.relpath(files[-1], ROOT)
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved, "peak": burst_peak,
                    "unit": "TFLOP/s", "frac": achieved / burst_peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": f"{peaks['_source']} (MEASURED_PEAKS.json bf16_tflops = BURST: the launches are timed one "
                                   f"at a time at boost clocks; the `sustained` leg is quoted against bf16_tflops_sustained)",
                    "launches_per_step": gemm["count"] / prof_steps,
                    "avg_launch_ms": gemm["ms"] / gemm["count"],
                    "flops_per_launch": gemm["work"] / gemm["count"],
                    "share_of_step": gemm["ms"] / tot_ms}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    membound = {}
    for name in ("gate_ln_pool_fwd", "gate_ln_pool_bwd", "attn_fwd_tc", "attn_bwd_tc", "gate_ln_fwd", "gate_ln_bwd",
                 "attn_fwd_mma", "attn_bwd_dq_mma", "attn_bwd_dkv_mma", "attn_delta", "pool_fwd", "cast_multi"):
        v = prof.get(name)
        if v and v["ms"] > 0:
            gbs = v["work"] / (v["ms"] * 1e-3) / 1e9
            membound[name] = {"achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak}

    # ---- baselines on this box (N = 1 only): torch eager on the same GPU, cuBLAS per GEMM shape, CPU oracle ----------
    eager = None
    cpu = None
    if rank == 0 and world == 1:
        del step
        torch.cuda.empty_cache()
        if not args.no_torch_eager:
            try:
                eager = torch_eager_baseline(B, L, dev)
                eager["gemm_shapes"] = cublas_matmul_tflops(prof, dev, prof_steps)
                # the same launches weighted by their count per step, timed back to back (no launch gaps, cold operands):
                # context for roofline.achieved, which brackets every launch of an eager pass with events
                fl = tm = 0.0
                for nm, v in eager["gemm_shapes"].items():
                    if v.get("mmsa_b2b_tflops"):
                        Mq, Nq, Kq = (int(x) for x in nm.split("_")[1].split("x"))
                        f = 2.0 * Mq * Nq * Kq * v["launches_per_step"]
                        fl += f
                        tm += f / (v["mmsa_b2b_tflops"] * 1e12)
                if roofline is not None and tm > 0:
                    roofline["achieved_back_to_back"] = fl / tm / 1e12
                    roofline["frac_back_to_back"] = fl / tm / 1e12 / burst_peak
                    roofline["back_to_back_note"] = ("the step's big tcgen05 GEMM shapes weighted by launches per step, each timed back "
                                                     "to back over operand sets > L2 (torch_eager.gemm_shapes.mmsa_b2b_tflops)")
                eager["gemm_shapes_note"] = ("cublas_tflops / mmsa_b2b_tflops: torch.matmul bf16 and this library's kernel, each back "
                                             "to back over operand sets > L2, CUDA events; mmsa_in_step_tflops: the same shape "
                                             "inside the profiled step (per-launch events)")
            except Exception as ex:
                eager = {"error": f"{type(ex).__name__}: {ex}"}
        try:      # the CPU baseline is an N = 1 figure (torchrun also pins OMP_NUM_THREADS=1 on the ranks)
            cpu = cpu_reference_samples_per_s(L, args.cpu_sample_batch, 3, 1)
        except Exception as ex:   # the baseline leg must never take the GPU numbers down with it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        gf = algorithmic_gflop_per_sample(L, B * world)
        value = B * world / (ms_step * 1e-3)
        if sustained is not None:
            sustained["step_tflops_per_gpu"] = sustained["value"] * gf["fwd_bwd"] / 1e3 / world
            sustained["step_frac_of_sustained_peak"] = sustained["step_tflops_per_gpu"] / sus_peak
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"configs[1]: ME-MHACL cross-modal 12-head attention fusion fwd+bwd "
                                   f"(projections, 2 cross-attention blocks, pool, fusion MLP, 3-class CE, InfoNCE), "
                                   f"L={L}, R={R}, E={E}, per-GPU batch {B}, global batch {B * world}",
                       "parallelism": f"dp{world}", "cuda_graph": graph_ok,
                       **({"grad_reducer": type(reducer).__name__, "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS")} if reducer is not None else {}),
                       **({"optimizer": "mmsa.FusedClipAdamW after every step (eager: pack + sumsq + clip_adamw), lr 1e-4, wd 0.01, clip 1.0"} if opt is not None else {}),
                       **({"ablate": os.environ["MMSA_BENCH_ABLATE"]} if os.environ.get("MMSA_BENCH_ABLATE") else {}),
                       "dropout": "train mode, in-kernel Philox; stream position in device memory, advanced by a graph node "
                                  "(every replay draws a new mask)",
                       "l2_policy": "per-step working set (> 1 GB of activations, 100 MB of inputs) exceeds the 126 MB L2; "
                                    "two input slots alternate",
                       "gflop_per_sample_fwd_bwd": gf["fwd_bwd"],
                       "step_tflops": value * gf["fwd_bwd"] / 1e3 / world,
                       "step_frac_of_tensor_peak": value * gf["fwd_bwd"] / 1e3 / world / burst_peak,
                       "step_frac_peak_source": "bf16_tflops (burst): the timed region is K steps at boost clocks"},
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e,
                    "host_feature_dtype": args.dtype, "note": "pinned host features; H2D of step i+1 overlaps step i (2 slots)",
                    "numa_bound_cpus": (len(numa_cpus) if numa_cpus else None)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "sustained": sustained,
            "memory_bound_kernels": membound,
            "kernels": kern,
            "ungraphed_ms_per_step": ms_ungraphed,
            "torch_eager": eager,
            "dp_parity": dp_parity,
            "configs": extra or None,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None,
            "loss": loss_val,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dp_parity_check(model, reducer, host_batch, cd, dev, rank: int, world: int) -> dict:
    """One data-parallel step with dropout off -- sharded InfoNCE (global diagonal at rank*B + i), reduce-scatter backward,
    AVG all-reduce of the parameter gradients, all on the CUDA kernels over NCCL -- against the same global batch evaluated
    on rank 0 ALONE (mmsa.dist.emulate_data_parallel_step: gathered inputs, no process group, BatchNorm per shard).
    Compares the mean loss, every rank's InfoNCE value and every parameter gradient; both sides run the same kernels, so
    the only differences are reduction orders (split-K over 256*N rows instead of 256, NCCL's ring order)."""
    import torch
    import torch.distributed as dist
    import mmsa
    from mmsa import dist as mdist
    text, image, labels = (t.to(dev) for t in host_batch)
    p_before = {m: m.p for m in model.modules() if isinstance(m, torch.nn.Dropout)}
    model.set_dropout(0.0)
    model.zero_grad(set_to_none=True)
    model.prepare_step()
    logits, closs = model(text, image, None, labels)
    loss = mmsa.cross_entropy(logits, labels) + closs.sum()
    loss.backward()
    if reducer is not None:
        reducer.step()
    stats = torch.stack([loss.detach().float().reshape(()), closs.detach().float().sum()])
    all_stats = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(all_stats, stats)
    gathered = []
    for t in (text, image, labels):
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), device=dev, dtype=t.dtype)
        dist.all_gather_into_tensor(out, t.contiguous())
        gathered.append(out)
    res = {"ok": True}
    if rank == 0:
        twin = mmsa.MultimodalTransformerModel(num_classes=C, embed_dim=E, num_heads=H, wiring="bidirectional",
                                               text_dim=DT, image_dim=DI, contract="single", compute_dtype=cd,
                                               valence=False).to(dev).train()
        twin.load_state_dict(model.state_dict())
        twin.set_dropout(0.0)
        ref_loss, ref_c, ref_grads = mdist.emulate_data_parallel_step(twin, gathered[0], gathered[1], gathered[2], world)
        w = float(model.contrastive_weight.detach())
        mean_loss = float(sum(float(s[0]) for s in all_stats) / world)
        e_loss = abs(mean_loss - float(ref_loss)) / max(abs(float(ref_loss)), 1e-12)
        e_c = max(abs(float(s[1]) - w * float(c)) / max(abs(w * float(c)), 1e-12) for s, c in zip(all_stats, ref_c))
        worst, worst_name, n = 0.0, None, 0
        for k, prm in model.named_parameters():
            if prm.grad is None or k not in ref_grads:
                continue
            r = ref_grads[k].double()
            e = float((prm.grad.double() - r).abs().max() / max(float(r.abs().max()), 1e-30))
            n += 1
            if e > worst:
                worst, worst_name = e, k
        tol = 2e-3 if cd == torch.bfloat16 else 1e-4
        res = {"ok": bool(e_loss <= tol and e_c <= tol and worst <= tol and n > 0), "tol": tol, "loss_rel": e_loss,
               "infonce_rel_max_over_ranks": e_c, "grad_rel_max": worst, "grad_worst": worst_name, "grads_compared": n,
               "mean_loss": mean_loss, "reference_loss": float(ref_loss), "world": world,
               "what": "sharded step on N ranks (NCCL all-gather / reduce-scatter / AVG all-reduce) vs the same global batch "
                       "on rank 0 alone through mmsa.dist.emulate_data_parallel_step; dropout off; per-shard BatchNorm"}
        del twin, ref_grads
    flag = torch.tensor([1 if res["ok"] else 0], device=dev)
    dist.broadcast(flag, 0)
    res["ok"] = bool(flag.item())
    for m, p in p_before.items():
        m.p = p
    model.zero_grad(set_to_none=True)
    del gathered
    torch.cuda.empty_cache()
    return res


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (configs[1]: 256)")
    ap.add_argument("--L", type=int, default=128, help="text tokens per sample (configs[1]: 128)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--optimizer", action="store_true",
                    help="also run the fused clip+AdamW update after every step (not part of the BASELINE metric)")
    ap.add_argument("--cpu-sample-batch", type=int, default=32, help="samples per CPU reference step")
    ap.add_argument("--sustained-seconds", type=float, default=3.0,
                    help="length of the sustained leg (same graph back to back); 0 switches it off")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the configs[2]/[3]/[4] legs")
    ap.add_argument("--no-torch-eager", action="store_true", help="skip the torch-eager / cuBLAS baseline leg (N=1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
