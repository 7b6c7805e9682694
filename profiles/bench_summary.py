"""One-line summaries of bench JSON files: python profiles/bench_summary.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        txt = [l for l in open(f).read().splitlines() if l.startswith("{")][-1]
        d = json.loads(txt)
        r = d.get("roofline") or {}
        print("%-40s value %.3g ms %.4f e2e %.3g kern %s frac %.3f launches %s" % (
            f.split("/")[-1], d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel_ms_all"), r.get("frac", 0),
            d.get("gpu_launches")))
    except Exception as e:  # noqa
        print(f, "unreadable:", e)
