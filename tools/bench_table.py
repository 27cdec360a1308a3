#!/usr/bin/env python
"""bench.py JSON line(s) -> the markdown tables of DESIGN.md §6 / §7.   python tools/bench_table.py line.json [ref.json]"""
import json
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith("{")]
    return json.loads(lines[-1])


d = load(sys.argv[1])
ref = load(sys.argv[2]) if len(sys.argv) > 2 else None
rows = [(d["config"]["scenario"], "configs[1]", d)] + [(k, v.get("config", ""), v) for k, v in d.get("secondary", {}).items()]
print("| scenario | BASELINE config | batch / GPU | steps/s | ms/step | HBM frac (alg. bytes) | fp64 frac (executed instr.) | binding roof | CPU baseline (cores) | GPU / CPU |")
print("|---|---|---|---|---|---|---|---|---|---|")
for name, cfg, v in rows:
    if "error" in v:
        print("| %s | %s | error: %s |" % (name, cfg, v["error"]))
        continue
    det = v.get("roofline_detail", {})
    hbm = det.get("hbm", {})
    f64 = det.get("fp64", det.get("fp64_algorithmic", {}))
    cb = v.get("cpu_baseline")
    if cb is None and ref is not None:
        cb = (ref if name == ref["config"]["scenario"] else ref.get("secondary", {}).get(name, {})).get("cpu_baseline")
    batch = v.get("batch_per_gpu", v.get("config", {}).get("batch_per_gpu") if isinstance(v.get("config"), dict) else None)
    print("| %s | %s | %s | %.3e | %.4f | %.3f (%d B) | %.3f (%s) | %s %.2f | %s | %s |" % (
        name, cfg, batch, v["value"], v["ms_per_step"], hbm.get("frac", float("nan")),
        hbm.get("algorithmic_bytes_per_step", 0), f64.get("frac", float("nan")),
        ("%.0f" % f64["executed_fp64_inst_per_instance"]) if "executed_fp64_inst_per_instance" in f64 else "alg. flops",
        v["roofline"]["bound"], v["roofline"]["frac"],
        ("%.3e (%d)" % (cb["value"], cb["cores"])) if cb else "-",
        ("%.0fx" % (v["value"] / cb["value"])) if cb else "-"))
if "e2e" in d:
    print("\ne2e %.3e steps/s (h2d %d B, d2h %d B per step); clocks %s" % (d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"],
                                                                       d["e2e"]["d2h_bytes_per_step"], d.get("clocks")))
