#!/usr/bin/env python
"""Host<->device copy bandwidth of this box (pinned memory): one direction at a time and both at once,
on 1, 2, 4, ... GPUs CONCURRENTLY.  This is the ceiling of bench.py's `e2e` number (64 B in + 52 B out per
UR5 instance) and of solve_batch(devices="all"): if n GPUs together move less than n x the single-GPU
figure, the host side of the box (root complex / memory path of the VM), not the engine, bounds the
multi-GPU end-to-end rate.

    python tools/pcie_probe.py [--mb 256] [--reps 8] > gpurun_out/pcie_probe.txt   (on a multi-GPU box)
"""
import argparse
import json
import threading
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=256)
ap.add_argument("--reps", type=int, default=8)
args = ap.parse_args()
n = args.mb << 20
ndev = torch.cuda.device_count()


class Dev(object):
    def __init__(self, i):
        self.i = i
        self.h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        self.h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        with torch.cuda.device(i):
            self.d_in = torch.empty(n, dtype=torch.uint8, device="cuda:%d" % i)
            self.d_out = torch.empty(n, dtype=torch.uint8, device="cuda:%d" % i)
            self.s1, self.s2 = torch.cuda.Stream(i), torch.cuda.Stream(i)

    def issue(self, h2d, d2h, reps):
        with torch.cuda.device(self.i):
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(self.s1):
                        self.d_in.copy_(self.h_in, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(self.s2):
                        self.h_out.copy_(self.d_out, non_blocking=True)

    def sync(self):
        torch.cuda.synchronize(self.i)


devs = [Dev(i) for i in range(ndev)]


def run(k, h2d, d2h, reps):
    use = devs[:k]
    for d in use:
        d.sync()
    start = threading.Barrier(k + 1)

    def work(d):
        start.wait()
        d.issue(h2d, d2h, reps)
        d.sync()

    th = [threading.Thread(target=work, args=(d,)) for d in use]
    for t in th:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return k * reps * n / dt / 1e9           # GB/s per direction, all devices together


rows = []
k = 1
while k <= ndev:
    run(k, True, True, 2)
    r = {"gpus": k, "h2d_only_GBs": run(k, True, False, args.reps), "d2h_only_GBs": run(k, False, True, args.reps),
         "both_each_GBs": run(k, True, True, args.reps)}
    # steps/s ceiling of the headline skill: 64 B in + 52 B out per instance, both directions at once
    r["ur5_track_e2e_ceiling_steps_per_s"] = min(r["both_each_GBs"] * 1e9 / 64.0, r["both_each_GBs"] * 1e9 / 52.0)
    rows.append(r)
    print("%d GPU(s): H2D only %.1f GB/s, D2H only %.1f GB/s, both %.1f GB/s each  -> ur5_track e2e ceiling %.3e steps/s"
          % (k, r["h2d_only_GBs"], r["d2h_only_GBs"], r["both_each_GBs"], r["ur5_track_e2e_ceiling_steps_per_s"]),
          flush=True)
    k *= 2
print(json.dumps(rows))
