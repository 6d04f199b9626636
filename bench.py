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


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    import mmsa
    from mmsa import _lib
    from mmsa import dist as mdist
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
    if world > 1:
        for p in model.parameters():            # replicas start identical
            dist.broadcast(p.data, 0)
        # MMSA_BENCH_ABLATE (comma list: "shard", "reduce") switches a collective off to attribute the multi-GPU
        # overhead; such a run is a diagnostic, flagged in config.ablate, never a result
        ablate = set(filter(None, os.environ.get("MMSA_BENCH_ABLATE", "").split(",")))
        if "shard" not in ablate:
            mdist.shard_contrastive(model)
        if "reduce" not in ablate:
            reducer = mdist.GradAllReducer(model.parameters())

    # synthetic features: N(0,1), seeded per rank; host copies pinned, in the feature dtype
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 2
    host = []
    for _ in range(n_host):
        host.append((torch.randn(B, L, DT, generator=g).to(cd).pin_memory(),
                     torch.randn(B, R, DI, generator=g).to(cd).pin_memory(),
                     torch.randint(0, C, (B,), generator=g).pin_memory()))

    opt = None
    if args.optimizer:          # fused global-norm clip + AdamW after every step (Trainer.py:80-81)
        opt = mmsa.FusedClipAdamW(model.parameters(), lr=1e-4, weight_decay=0.01, max_norm=1.0)
        for grp in opt.param_groups:
            opt._arena(grp)     # re-home the parameters into the flat arena BEFORE the step graph captures their addresses
    step = TrainStep(model, B, L, R, DT, DI, feature_dtype=cd, n_slots=2, use_graph=not args.no_graph,
                     post_backward=(reducer.step if reducer is not None else None), device=dev)
    for k, s in enumerate(step.slots):
        s.text.copy_(host[k % n_host][0]); s.image.copy_(host[k % n_host][1]); s.labels.copy_(host[k % n_host][2])
    step.warmup(2)
    graph_ok = not args.no_graph
    if graph_ok:
        try:
            step.capture()
        except Exception as ex:                  # e.g. a collective that cannot be captured
            if rank == 0:
                print(f"bench.py: CUDA-graph capture failed ({type(ex).__name__}: {ex}); timing eager launches",
                      file=sys.stderr)
            graph_ok = False
            for s in step.slots:
                s.graph = None
            torch.cuda.synchronize()

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

    # ---- device-resident throughput (`value`) ----
    for _ in range(max(args.warmup, 3)):
        for k in (0, 1):
            step.run(k)
            if opt is not None:
                opt.step()
    # the clock sampler (NVML init: tens of ms) starts BEFORE the barrier: anything rank 0 does between the barrier and
    # its first launch shows up as a stall inside the other ranks' first collective
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    trace = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)] if os.environ.get("MMSA_BENCH_TRACE") else None
    ev0.record()
    for i in range(args.steps):
        step.run(i & 1)
        if opt is not None:
            opt.step()
        if trace is not None:
            trace[i].record()
    ev1.record()
    barrier()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    if trace is not None:        # diagnostic: per-step device time (where do multi-GPU stalls sit?)
        marks = [ev0] + trace
        print(f"rank {rank} per-step ms: " + " ".join(f"{marks[i].elapsed_time(marks[i + 1]):.2f}" for i in range(args.steps)),
              file=sys.stderr, flush=True)
    clocks = sampler.stop() if sampler is not None else None
    launches = (step.launches_per_step * args.steps + (_lib.launch_count() - n0)) if graph_ok else (_lib.launch_count() - n0)
    loss_val = float(step.slots[0].loss.item())

    # ---- end-to-end through the public step API with host inputs (`e2e`) ----
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

    # ---- eager pass with per-launch CUDA events: kernel breakdown + roofline of the dominant kernel ----
    for s in step.slots:
        s.graph = None
    from mmsa import ops as _ops
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
    ms_eager = ee0.elapsed_time(ee1) / prof_steps
    # Per-launch CUDA events only measure kernel time if the GPU never waits for the host: in eager mode the host
    # needs longer to enqueue a step than the GPU to run it, so each profiled step is queued behind a device-side
    # delay (torch.cuda._sleep) long enough for the host to get a whole step ahead.
    delay_cycles = int((ms_eager + 2.0) * 1e-3 * 2.0e9)
    _lib.prof_enable(True)
    for i in range(prof_steps):
        torch.cuda._sleep(delay_cycles)
        step.run(i & 1)
    _lib.prof_enable(False)
    prof = _lib.prof_collect()
    peaks = load_peaks()
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
    roofline = None
    if gemm["ms"] > 0:
        achieved = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
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
        roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": f"{peaks['_source']} (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)",
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

    if rank == 0:
        cpu = None
        if world == 1:      # the CPU baseline is an N = 1 figure (torchrun also pins OMP_NUM_THREADS=1 on the ranks)
            try:
                cpu = cpu_reference_samples_per_s(L, args.cpu_sample_batch, 3, 1)
            except Exception as ex:   # the baseline leg must never take the GPU numbers down with it
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
        gf = algorithmic_gflop_per_sample(L, B * world)
        value = B * world / (ms_step * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"configs[1]: ME-MHACL cross-modal 12-head attention fusion fwd+bwd "
                                   f"(projections, 2 cross-attention blocks, pool, fusion MLP, 3-class CE, InfoNCE), "
                                   f"L={L}, R={R}, E={E}, per-GPU batch {B}, global batch {B * world}",
                       "parallelism": f"dp{world}", "cuda_graph": graph_ok,
                       **({"optimizer": "mmsa.FusedClipAdamW after every step (eager: pack + sumsq + clip_adamw), lr 1e-4, wd 0.01, clip 1.0"} if opt is not None else {}),
                       **({"ablate": os.environ["MMSA_BENCH_ABLATE"]} if os.environ.get("MMSA_BENCH_ABLATE") else {}), "dropout": "train mode, in-kernel Philox",
                       "l2_policy": "per-step working set (> 1 GB of activations, 100 MB of inputs) exceeds the 126 MB L2; "
                                    "two input slots alternate",
                       "gflop_per_sample_fwd_bwd": gf["fwd_bwd"],
                       "step_tflops": value * gf["fwd_bwd"] / 1e3 / world,
                       "step_frac_of_tensor_peak": value * gf["fwd_bwd"] / 1e3 / world / peaks.get("bf16_tflops_sustained", 1400.0)},
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e,
                    "host_feature_dtype": args.dtype, "note": "pinned host features; H2D of step i+1 overlaps step i (2 slots)",
                    "numa_bound_cpus": (len(numa_cpus) if numa_cpus else None)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "memory_bound_kernels": membound,
            "kernels": kern,
            "eager_ms_per_step": ms_eager,
            "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None,
            "loss": loss_val,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
