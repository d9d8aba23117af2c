"""Prints the headline fields of bench.py JSON lines:  python tools/show_line.py FILE [FILE ...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().split("\n")[-1])
    except Exception as exc:
        print(path, "unreadable:", exc)
        continue
    e2e = d.get("e2e") or {}
    print(path, "| n_gpus", d.get("n_gpus"), "| ms/step %.4f" % d.get("ms_per_step", float("nan")), "| value %.3f G" % (d.get("value", 0) / 1e9),
          "| e2e ms %.4f" % e2e.get("ms_per_step", float("nan")), "| roofline", (d.get("roofline") or {}).get("kernel"),
          "%.3f" % (d.get("roofline") or {}).get("frac", float("nan")))
    if d.get("step_timeline_ms"):
        print("   timeline", d["step_timeline_ms"])
    if d.get("parity"):
        print("   parity", {k: d["parity"][k] for k in ("decisions_differ_outside_margin_band", "decisions_differ_total") if k in d["parity"]})
