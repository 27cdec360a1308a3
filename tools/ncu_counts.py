#!/usr/bin/env python
"""ncu CSV (one `--metrics` pass over `bench.py --secondary-only X` or the headline bench) ->
profiles/r2_ncu_counts.json, which bench.py reads for `roofline.traffic` and the fp64 roofline from the
EXECUTED instruction count.

On the GPU box (one scenario per ncu run; the timed loop of a scenario is only its step kernels):

    M=dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,\
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum
    ncu --metrics $M --clock-control none -k regex:clik_ -s 6 -c 4 --csv --log-file gpurun_out/cnt_X.csv \
        python bench.py --secondary-only X

Here:  python tools/ncu_counts.py scenario=batch:csv [...]
Per scenario the launches of one step are averaged per kernel name and summed over the kernels of a step
(the QP step is a fast + a tail launch).  fp64_inst_per_instance = warp-level fp64-pipe instructions x 32 /
batch: what the fp64 pipe had to issue per instance, idle lanes of divergent warps included."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_ncu_counts.json")


def parse(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    per = {}          # (launch id) -> {kernel, metric: value}
    for r in rows:
        lid = r.get("ID")
        d = per.setdefault(lid, {"kernel": r.get("Kernel Name", "")})
        try:
            d[r["Metric Name"]] = float(str(r["Metric Value"]).replace(",", ""))
        except (KeyError, ValueError):
            pass
        d.setdefault("units", {})[r.get("Metric Name")] = r.get("Metric Unit")
    return list(per.values())


def main():
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    for arg in sys.argv[1:]:
        spec, path = arg.split(":", 1)
        if spec.startswith("range="):
            # range=<scenario>=<launches in the range>:<csv of tools/ncu_range_traffic.py under --replay-mode range>:
            # DRAM bytes of K consecutive launches on K distinct buffer sets, write-back included
            _, name, k = spec.split("=")
            rows = parse(path)
            rd = sum(r.get("dram__bytes_read.sum", 0.0) for r in rows)
            wr = sum(r.get("dram__bytes_write.sum", 0.0) for r in rows)
            data.setdefault(name, {})["dram_bytes_range_per_launch"] = (rd + wr) / int(k)
            data[name]["dram_range"] = {"launches": int(k), "read": rd, "write": wr,
                                        "source": "ncu --replay-mode range, %s" % os.path.basename(path)}
            print(name, "range of %s launches: %.1f MB per launch" % (k, (rd + wr) / int(k) / 1e6))
            continue
        name, batch = spec.split("=")
        batch = int(batch)
        launches = [l for l in parse(path) if l["kernel"].startswith("clik_") and "sizes" not in l["kernel"]
                    and "dfma_peak" not in l["kernel"]]
        by_kernel = {}
        for l in launches:
            by_kernel.setdefault(l["kernel"].split("(")[0], []).append(l)
        tot = {"dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "fp64_warp_inst": 0.0, "warp_inst": 0.0,
               "time_us": 0.0}
        kernels = {}
        for k, ls in by_kernel.items():
            def avg(metric, scale=1.0):
                vals = [l[metric] for l in ls if metric in l]
                return scale * sum(vals) / len(vals) if vals else 0.0
            unit_r = ls[0].get("units", {}).get("dram__bytes_read.sum", "byte")
            unit_w = ls[0].get("units", {}).get("dram__bytes_write.sum", "byte")
            sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            t_unit = ls[0].get("units", {}).get("gpu__time_duration.sum", "ns")
            ts = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(t_unit, 1e-3)
            rec = {"launches_averaged": len(ls),
                   "dram_bytes_read": avg("dram__bytes_read.sum", sc.get(unit_r, 1.0)),
                   "dram_bytes_write": avg("dram__bytes_write.sum", sc.get(unit_w, 1.0)),
                   "fp64_warp_inst": avg("sm__inst_executed_pipe_fp64.sum"),
                   "warp_inst": avg("smsp__inst_executed.sum"),
                   "fp64_pipe_pct_elapsed": avg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                   "time_us": avg("gpu__time_duration.sum", ts)}
            kernels[k] = rec
            for key in tot:
                tot[key] += rec[key]
        keep = {k: v for k, v in data.get(name, {}).items() if k.startswith("dram_range") or k == "dram_bytes_range_per_launch"}
        data[name] = {"batch": batch, "dram_bytes": tot["dram_bytes_read"] + tot["dram_bytes_write"],
                      "dram_bytes_read": tot["dram_bytes_read"], "dram_bytes_write": tot["dram_bytes_write"],
                      "fp64_inst_per_instance": 32.0 * tot["fp64_warp_inst"] / batch,
                      "inst_per_instance": 32.0 * tot["warp_inst"] / batch,
                      "ncu_time_us_per_step": tot["time_us"], "kernels": kernels,
                      "source": "ncu --metrics pass, %s" % os.path.basename(path)}
        data[name].update(keep)
        print(name, json.dumps({k: v for k, v in data[name].items() if k != "kernels"}))
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
