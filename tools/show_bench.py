"""One-screen digest of a bench.py JSON line.  python tools/show_bench.py profiles/r02_bench_8gpu.json"""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    e = d.get("e2e", {})
    print(path)
    print("  n_gpus", d.get("n_gpus"), " ms/step", round(d["ms_per_step"], 4), " value", round(d["value"]), d["unit"],
          " roofline.frac", round((d.get("roofline") or {}).get("frac", 0), 4), " launches", d.get("gpu_launches"))
    print("  e2e", round(e.get("value", 0)), " ms/step", round(e.get("ms_per_step", 0), 4), " stage threads", e.get("host_stage_threads"))
    for k in ("pinned_fp32_prefetched", "pinned_fp32_blocking_item", "pageable_f64_blocking_item"):
        if isinstance(e.get(k), dict):
            print("   ", k, round(e[k].get("value", 0)), " ms/step", round(e[k].get("ms_per_step", 0), 4))
    if isinstance(d.get("device_feed"), dict):
        print("    device_feed", round(d["device_feed"].get("value", 0)), " ms/step", round(d["device_feed"].get("ms_per_step", 0), 4))
    if d.get("dp_check"):
        print("  dp_check", json.dumps(d["dp_check"])[:400])
    if d.get("score"):
        s = d["score"]
        print("  score", {k: (round(v) if isinstance(v, (int, float)) else v) for k, v in s.items() if "per_s" in k})
    print("  optimizer_step:", d.get("optimizer_step"), "| clocks", d.get("clocks"))
