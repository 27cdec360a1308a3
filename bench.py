#!/usr/bin/env python
"""bench.py — controller-steps/sec of the batched CLIK controller step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--scenario ur5_track] [--batch 1048576]

A "step" is one pass of the hot path (PseudoInverseController.solve for every instance of the
batch — reference casclik/controllers/pseudo_inverse.py:512-556) over one batch of synthetic
inputs (BASELINE.json configs[1]: UR5, one EqualityConstraint, 2^20 random joint states/targets
per GPU, fp64).  One JSON line is printed by rank 0; see the keys at the bottom.

Timing: W warm-up steps, then exactly K steps between (barrier + synchronize), timed with CUDA
events on the launch stream, MAX over ranks.  The step rotates over several resident input sets
whose total size exceeds L2 (126 MB), so no step re-reads a cache-resident batch.
`value` times device-resident inputs; `e2e` times the same step through the host-buffer C ABI
(pinned host inputs, H2D + kernel + D2H inside the timed region).
`--impl reference` times the restated reference CPU path (oracle/clik_oracle.c, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "controller-steps/sec (fp64, batch N)"
UNIT = "controller-steps/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def _launch_info(ctrl, is_qp):
    sk = ctrl._skill()
    if is_qp:
        info = sk.launch_info(1)
        if ctrl.kernel_meta.get("qp_split") and os.environ.get("CLIK_QP_SPLIT", "1") != "0":
            info = {"fast": sk.launch_info(3), "tail": sk.launch_info(4), "full (status == NULL)": info,
                    "used": "fast + tail"}
        return info
    info = {"plain": sk.launch_info(0)}
    try:
        info["tma"] = sk.launch_info(2)
        info["used"] = "tma" if os.environ.get("CLIK_TMA", "0") == "1" else "plain"
    except Exception:
        info["used"] = "plain"
    return info


def cpu_reference(scenario, batch, seconds_target=12.0, threads=0):
    """Restated reference CPU path (oracle/clik_oracle.c) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_port
    import clik_oracle as orc
    from casclik_b200 import fk
    if scenario.name != "ur5_track":
        raise SystemExit("CPU baseline is implemented for the ur5_track workload")
    chain = orc.load_chain(fk.UR5_URDF, "base_link", "tool0")
    cores = c_port.max_threads() if threads <= 0 else threads
    n = 4096
    inp = scenario.sample(n, seed=0)
    c_port.pinv_track(chain, inp["q"], inp["y"], threads=cores)          # warm
    t0 = time.perf_counter()
    c_port.pinv_track(chain, inp["q"], inp["y"], threads=cores)
    rate = n / max(time.perf_counter() - t0, 1e-9)
    sample = int(min(batch, max(4096, rate * seconds_target)))
    passes = max(1, int(round(rate * seconds_target / sample)))
    inp = scenario.sample(sample, seed=0)
    t0 = time.perf_counter()
    for _ in range(passes):
        _, used = c_port.pinv_track(chain, inp["q"], inp["y"], threads=cores)
    dt = time.perf_counter() - t0
    return {"value": sample * passes / dt, "unit": UNIT, "cores": int(used), "kind": "port",
            "sample": "%d passes over the first %d instances of the %d-instance batch, %.2f s wall, "
                      "oracle/clik_oracle.c (gcc -O2 -fopenmp)" % (passes, sample, batch, dt)}, sample, dt


def run_reference(args, scenario, rank, world):
    if rank != 0:
        return
    steps, warm = max(args.steps, 1), max(args.warmup, 0)
    # bounded: every step is a sample of the batch sized so that the whole run (warm-up + K
    # steps) costs about two minutes of host time
    budget = 120.0 / (steps + min(warm, 3))
    base, sample, dt = cpu_reference(scenario, args.batch, seconds_target=budget)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_port
    import clik_oracle as orc
    from casclik_b200 import fk
    chain = orc.load_chain(fk.UR5_URDF, "base_link", "tool0")
    inp = scenario.sample(sample, seed=0)
    for _ in range(min(warm, 3)):
        c_port.pinv_track(chain, inp["q"], inp["y"], threads=base["cores"])
    t0 = time.perf_counter()
    for _ in range(steps):
        c_port.pinv_track(chain, inp["q"], inp["y"], threads=base["cores"])
    dt = time.perf_counter() - t0
    value = sample * steps / dt
    base["value"] = value
    base["sample"] = ("%d-instance sample of the %d-instance batch per step, %d steps, %.2f s wall, "
                      "oracle/clik_oracle.c (gcc -O2 -fopenmp)" % (sample, args.batch, steps, dt))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": scenario.description, "scenario": scenario.name,
                   "batch_per_step": sample,
                   "note": "reference CPU path restated in C (CasADi/qpOASES are not installable "
                           "here); host threads only, no GPU"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenario", default="ur5_track")
    ap.add_argument("--batch", type=int, default=1 << 20, help="instances per GPU per step")
    ap.add_argument("--sets", type=int, default=5, help="resident input sets rotated over")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    from casclik_b200 import scenarios
    scenario = scenarios.get(args.scenario)

    if args.impl == "reference":
        run_reference(args, scenario, rank, world)
        return

    import torch
    import torch.distributed as dist
    from casclik_b200 import runtime
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); "
                         "use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        if os.environ.get("CLIK_NUMA_BIND", "1") == "1":
            # one process per GPU: keep each rank (and the pinned host buffers of its e2e leg) on the
            # socket its GPU hangs off.  Not at N = 1, where the cpu_baseline leg wants every host core.
            from casclik_b200 import sharding
            numa = sharding.bind_to_device_numa_node(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctrl = scenario.make_controller()
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    meta = ctrl.kernel_meta
    is_qp = scenario.controller == "qp"
    B = args.batch
    # every rank owns an independent shard (weak scaling: B instances per GPU); no collective
    # on the data path — instances are independent
    sets = []
    for s in range(args.sets):
        inp = scenario.sample(B, seed=1000 * rank + s)
        sets.append(tuple(None if inp[k] is None else torch.from_numpy(np.ascontiguousarray(inp[k])).to(dev)
                          for k in ("t", "q", "x", "y")))
    if is_qp:
        out = (torch.empty((meta["qp_n"], B), dtype=torch.float64, device=dev),
               torch.empty((B,), dtype=torch.int32, device=dev),
               torch.empty((2, B), dtype=torch.int32, device=dev))
    else:
        nq, nx = meta["n_robot"], meta["n_virtual"]
        out = (torch.empty((nq, B), dtype=torch.float64, device=dev),
               torch.empty((nx, B), dtype=torch.float64, device=dev) if nx else None,
               torch.empty((B,), dtype=torch.int32, device=dev))

    def step(i):
        t, q, x, y = sets[i % len(sets)]
        ctrl.solve_batch(t, q, x, y, out=out)

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    # The K timed steps are K launches of the step kernel through the C ABI.  They are captured once
    # into a CUDA graph and replayed, so that host jitter (8 ranks sharing the box's cores with the
    # clock samplers) cannot turn a 29 us kernel into a launch-bound loop; same kernels, same inputs.
    graph = None
    if os.environ.get("CLIK_BENCH_GRAPH", "1") == "1":
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(args.steps):
                    step(i)
            g.replay()                       # untimed: uploads the graph
            graph = g
        except Exception as exc:             # capture not possible: time the plain launch loop
            sys.stderr.write("bench: CUDA graph capture failed (%s); timing direct launches\n" % exc)
            graph = None
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        for i in range(args.steps):
            step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer C ABI (pinned host inputs, copies inside) --------------
    inp = scenario.sample(B, seed=1000 * rank + 77)
    pin = {k: (None if inp[k] is None else torch.from_numpy(np.ascontiguousarray(inp[k])).pin_memory())
           for k in ("t", "q", "x", "y")}
    host = {k: (None if v is None else v.numpy()) for k, v in pin.items()}
    if is_qp:
        hout = (torch.empty((meta["qp_n"], B), dtype=torch.float64).pin_memory().numpy(),
                torch.empty((B,), dtype=torch.int32).pin_memory().numpy(),
                torch.empty((2, B), dtype=torch.int32).pin_memory().numpy())
        d2h = B * (8 * meta["qp_n"] + 4 + 8)
    else:
        hout = (torch.empty((meta["n_robot"], B), dtype=torch.float64).pin_memory().numpy(),
                torch.empty((meta["n_virtual"], B), dtype=torch.float64).pin_memory().numpy()
                if meta["n_virtual"] else None,
                torch.empty((B,), dtype=torch.int32).pin_memory().numpy())
        d2h = B * (8 * (meta["n_robot"] + meta["n_virtual"]) + 4)
    # the host ABI copies only the input rows the compiled kernel reads (read masks in the cubin)
    mk = meta["qp_read_masks" if is_qp else "pinv_read_masks"]
    rows_in = (mk[0] if host["t"].size > 1 else 0)
    for mask, key in ((mk[1], "q"), (mk[2], "x"), (mk[3], "y")):
        if host[key] is not None:
            rows_in += bin(mask & ((1 << host[key].shape[0]) - 1)).count("1")
    h2d = 8 * B * rows_in + (8 if (mk[0] and host["t"].size == 1) else 0)

    def e2e_step():
        ctrl.solve_batch(host["t"], host["q"], host["x"], host["y"], out=hout)

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()          # synchronous: returns when the results are in host memory
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.e2e_steps / float(tdt.item())
    # keep the GPU under load until the sampler has a few readings (the timed region itself can
    # be only milliseconds long); these extra steps are outside every timed region
    t_load = time.perf_counter()
    while len(sampler.rows) < 4 and time.perf_counter() - t_load < 3.0:
        for i in range(200):
            step(i)
        torch.cuda.synchronize()
    clocks = sampler.summary()
    clocks["note"] = ("sampled every 100 ms from warm-up to the end of the e2e loop (plus up to 3 s "
                      "of extra untimed steps when the run is too short for 4 samples)")

    if rank == 0:
        hbm_peak, hbm_src = _peaks()
        fp64_peak = runtime.measure_fp64_peak(local)
        sec = ms * 1e-3 / args.steps
        if is_qp:
            flops, bytes_step = None, meta["qp_bytes_per_step"]
        else:
            flops, bytes_step = meta["pinv_flops_mode0"], meta["pinv_bytes_per_step"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tr = json.load(f)
            if tr.get("scenario") == scenario.name and tr.get("batch") == B:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]   # bytes per launch, from ncu
        ach_gbs = bytes_step * B / sec / 1e9
        roof_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach_gbs / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                    "algorithmic_bytes_per_step": bytes_step}
        roof = roof_hbm
        detail = {"hbm": roof_hbm}
        if flops:
            ach_tf = flops * B / sec / 1e12
            roof_f = {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                      "frac": ach_tf / fp64_peak, "traffic": traffic,
                      "peak_source": "measured in this run (DFMA micro-benchmark, clik_measure_fp64_peak; "
                                     "MEASURED_PEAKS.json has no fp64 entry)",
                      "algorithmic_flops_per_step": flops,
                      "transcendentals_per_step": meta["pinv_eval"]["transcendentals"]}
            detail["fp64"] = roof_f
            if roof_f["frac"] >= roof_hbm["frac"]:
                roof = roof_f
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": scenario.description, "scenario": scenario.name,
                       "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": "independent shards, one per GPU, no collective on the data path",
                       "launch_mode": ("%d step-kernel launches replayed from one CUDA graph" % args.steps
                                       if graph is not None else "direct launches"),
                       "host_affinity": ("rank 0 bound to its GPU's NUMA node: %d of %d CPUs" % (len(numa[1]), len(numa[0]))
                                         if numa else "unbound"),
                       "l2": "rotating %d resident input sets (%d MB total) > 126 MB L2"
                             % (args.sets, args.sets * bytes_step * B // (1 << 20)),
                       "launch": _launch_info(ctrl, is_qp)},
            "roofline": {k: v for k, v in roof.items()},
            "roofline_detail": detail,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
                    "path": "solve_batch(host arrays) -> clik_*_step_host with pinned host buffers: the "
                            "kernel reads inputs from / writes results to mapped host memory over PCIe "
                            "inside the timed region (pageable buffers would take the chunked H2D / "
                            "kernel / D2H pipeline instead)"},
            "gpu_launches": args.steps * (2 if (is_qp and meta.get("qp_split")
                                                and os.environ.get("CLIK_QP_SPLIT", "1") != "0") else 1),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and scenario.name == "ur5_track":
            line["cpu_baseline"], _, _ = cpu_reference(scenario, B)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
