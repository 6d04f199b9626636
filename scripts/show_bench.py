"""Pretty-print the JSON line of a bench.py log: headline numbers + per-kernel table."""
import json, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.log"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for line in open(path):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    k = d.pop("kernels", None)
    print(f"value={d['value']:.0f} {d['unit']}  ms/step={d['ms_per_step']:.3f}  e2e={d['e2e']['value']:.0f}  "
          f"eager_ms={d.get('eager_ms_per_step', 0):.3f}  gpus={d['n_gpus']}  launches={d.get('gpu_launches')}")
    print("roofline", json.dumps(d.get("roofline")))
    print("membound", json.dumps(d.get("memory_bound_kernels")))
    print("clocks", d.get("clocks"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    if k:
        tot = sum(v["ms_per_step"] for v in k.values())
        print(f"sum of kernel times {tot*1e3:.0f} us")
        for n, v in sorted(k.items(), key=lambda kv: -kv[1]["ms_per_step"])[:top]:
            print(f"{n:36s} x{v['launches_per_step']:4.1f} {v['ms_per_step']*1e3:8.1f}us  {v['share']*100:5.1f}%  "
                  f"{v['work_per_step']/v['ms_per_step']/1e9:8.1f} G/s")
