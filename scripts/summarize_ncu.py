"""Distil the ncu outputs of scripts/gpu_profile.sh (gpurun_out/) into tracked summaries under profiles/.

    python scripts/summarize_ncu.py r01

writes profiles/<tag>_launches.md   (per-kernel launch list of one bench step: launches, device time, share)
       profiles/<tag>_launches.csv  (the raw `ncu --metrics gpu__time_duration.sum` list)
       profiles/<tag>_ncu_full.md   (`ncu --set full` key counters per captured kernel)
Runs on the CPU box (ncu -i reads the .ncu-rep files; no GPU needed)."""
import collections
import csv
import io
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(PROF, exist_ok=True)


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("mmsa::", "").replace("<unnamed>::", "")


def launches():
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        key = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        d = agg.setdefault(key, [0, 0.0])
        d[0] += 1
        t = float(r["Metric Value"]) / 1e3
        d[1] += t
        tot += t
    ours = sum(t for (n, _, _), (c, t) in agg.items() if not n.startswith("at::") and "nccl" not in n.lower())
    with open(os.path.join(PROF, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: kernel launch list of `python bench.py --steps 2 --warmup 1 --no-graph` (ncu, "
                f"`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        nsteps = sum(1 for r in rows if "ce_fwd" in r["Kernel Name"])
        f.write(f"{len(rows)} consecutive launches captured from the start of the run ({nsteps} cross-entropy launches = steps of "
                f"configs[1]: B=256, L=128, bf16; the two giant bf16 fills are the input slots' one-time torch.zeros). "
                f"Times are cold-cache and serialised: compare SHARES.  Total {tot:.0f} us, of which {ours:.0f} us "
                f"({100 * ours / tot:.1f} %) in this repo's kernels (the rest are torch fill/add plumbing kernels of the autograd tape).\n\n")
        f.write("| kernel | grid | block | launches | avg us | total us | share |\n|---|---|---|---:|---:|---:|---:|\n")
        for (n, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n[:70]}` | {g} | {b} | {c} | {t / c:.1f} | {t:.1f} | {100 * t / tot:.2f} % |\n")
    shutil.copy(path, os.path.join(PROF, f"{tag}_launches.csv"))


KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem thr %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
]


def full(rep: str, f):
    path = os.path.join(OUT, rep)
    if not os.path.exists(path):
        return
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    if r.returncode != 0:
        f.write(f"\n(`{rep}`: ncu -i failed: {r.stderr.strip()[:200]})\n")
        return
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    f.write(f"\n## `{rep}`\n\n| kernel | grid | " + " | ".join(lbl for _, lbl in KEYS) + " |\n|---|---|" + "---:|" * len(KEYS) + "\n")
    for d in data:
        cells = []
        for k, _ in KEYS:
            if k in idx:
                v, u = d[idx[k]], units[idx[k]]
                try:
                    v = f"{float(v):.1f}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("-")
        f.write(f"| `{short(d[idx['Kernel Name']])[:60]}` | {d[idx['Grid Size']]} | " + " | ".join(cells) + " |\n")


def gemm_traffic():
    """per-launch DRAM traffic of the tcgen05 GEMM launches of one step -> profiles/<tag>_gemm_traffic.json
    (bench.py reads it for roofline.traffic)."""
    import json
    path = os.path.join(OUT, "gemm_traffic.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r["Grid Size"]})
        v = float(r["Metric Value"])
        unit = r["Metric Unit"].lower()
        if r["Metric Name"].startswith("dram__bytes"):
            mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            d[r["Metric Name"]] = v * mult
        else:
            mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
            d["us"] = v * mult
    launches = list(per.values())
    n = len(launches)
    if n == 0:
        return
    rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in launches)
    wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in launches)
    out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of the {n} gemm_tcgen05_kernel launches of one "
                     "step of `python bench.py --steps 2 --warmup 1 --no-graph` (configs[1], B=256, L=128, bf16)",
           "launches": n, "dram_read_bytes_per_launch": rd / n, "dram_write_bytes_per_launch": wr / n,
           "traffic_bytes_per_launch": (rd + wr) / n, "total_us": sum(l.get("us", 0.0) for l in launches)}
    with open(os.path.join(PROF, f"{tag}_gemm_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    shutil.copy(path, os.path.join(PROF, f"{tag}_gemm_traffic.csv"))


launches()
gemm_traffic()
with open(os.path.join(PROF, f"{tag}_ncu_full.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` captures (key counters)\n\n"
            "dram rd / dram wr are per launch (`dram__bytes_read.sum`, `dram__bytes_write.sum`); tensor % is "
            "`sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active`.  The .ncu-rep files themselves stay in "
            "gpurun_out/ (scratch, too large to track).\n")
    for rep in ("prof_gemm.ncu-rep", "prof_gemm2.ncu-rep", "prof_mem.ncu-rep", "prof_attn.ncu-rep"):
        full(rep, f)
print("wrote", PROF)
