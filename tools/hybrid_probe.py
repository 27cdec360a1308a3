#!/usr/bin/env python
"""End-to-end variants for pinned host batches: (a) the zero-copy kernel (SM loads and stores over PCIe),
(b) hybrid: inputs brought to the device by the copy engine in chunks, the kernel of each chunk reads device
memory and writes its results straight into the mapped host arrays, (c) reverse hybrid: SM loads over PCIe,
results to device, copy engine D2H.   python tools/hybrid_probe.py [scenario] [log2 N]"""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from casclik_b200 import scenarios, runtime

name = sys.argv[1] if len(sys.argv) > 1 else "ur5_track"
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
sc = scenarios.get(name)
ctrl = sc.make_controller(); ctrl.setup_solver()
inp = sc.sample(N, seed=3)
pin = {k: torch.from_numpy(np.ascontiguousarray(inp[k])).pin_memory() for k in ("t", "q", "y")}
hqd = torch.empty((6, N), dtype=torch.float64).pin_memory()
hmd = torch.empty((N,), dtype=torch.int32).pin_memory()
dq = torch.empty((6, N), dtype=torch.float64, device="cuda")
dy = torch.empty((pin["y"].shape[0], N), dtype=torch.float64, device="cuda")
dt_ = pin["t"].cuda()
dqd = torch.empty((6, N), dtype=torch.float64, device="cuda")
dmd = torch.empty((N,), dtype=torch.int32, device="cuda")
lib = runtime.load_library()
sk = ctrl._skill()
qmask = ctrl.kernel_meta["pinv_read_masks"][1]
qrows = [j for j in range(6) if (qmask >> j) & 1]
print("q rows read:", qrows)
P = lambda t_, off=0: ctypes.c_void_p(t_.data_ptr() + off)


def zero_copy():
    ctrl.solve_batch(pin["t"].numpy(), pin["q"].numpy(), None, pin["y"].numpy(), out=(hqd.numpy(), None, hmd.numpy()))


def hybrid(chunk, nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    def run():
        for k, lo in enumerate(range(0, N, chunk)):
            hi = min(N, lo + chunk)
            st = streams[k % nstreams]
            with torch.cuda.stream(st):
                r0, r1 = qrows[0], qrows[-1] + 1
                dq[r0:r1, lo:hi].copy_(pin["q"][r0:r1, lo:hi], non_blocking=True)
                dy[:, lo:hi].copy_(pin["y"][:, lo:hi], non_blocking=True)
                runtime.check(lib.clik_pinv_step_ld(sk.handle, hi - lo, N, P(dt_, 8 * lo), 1, P(dq, 8 * lo), None, P(dy, 8 * lo),
                                                    P(hqd, 8 * lo), None, P(hmd, 4 * lo), ctypes.c_void_p(st.cuda_stream)))
        torch.cuda.synchronize()
    return run


def reverse(chunk, nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    def run():
        for k, lo in enumerate(range(0, N, chunk)):
            hi = min(N, lo + chunk)
            st = streams[k % nstreams]
            with torch.cuda.stream(st):
                runtime.check(lib.clik_pinv_step_ld(sk.handle, hi - lo, N, P(pin["t"], 8 * lo), 1, P(pin["q"], 8 * lo), None, P(pin["y"], 8 * lo),
                                                    P(dqd, 8 * lo), None, P(dmd, 4 * lo), ctypes.c_void_p(st.cuda_stream)))
                hqd[:, lo:hi].copy_(dqd[:, lo:hi], non_blocking=True)
                hmd[lo:hi].copy_(dmd[lo:hi], non_blocking=True)
        torch.cuda.synchronize()
    return run


variants = [("zero-copy kernel (current)", zero_copy)]
for chunk in (1 << 15, 1 << 16, 1 << 17, 1 << 18):
    for ns in (2, 4):
        variants.append(("hybrid H2D copy engine + stores to host, chunk 2^%d, %d streams" % (chunk.bit_length() - 1, ns), hybrid(chunk, ns)))
variants.append(("reverse hybrid (SM loads, D2H copy engine), chunk 2^16, 4 streams", reverse(1 << 16, 4)))
variants.append(("reverse hybrid, chunk 2^17, 4 streams", reverse(1 << 17, 4)))
ref = None
for tag, f in variants:
    for _ in range(3):
        f()
    t0 = time.perf_counter()
    for _ in range(20):
        f()
    dt = (time.perf_counter() - t0) / 20
    same = ""
    if ref is None:
        ref = (hqd.clone(), hmd.clone())
    else:
        same = "  same bits: %s" % bool(torch.equal(ref[0], hqd) and torch.equal(ref[1], hmd))
    print("%-78s %.3f ms  %.3e steps/s%s" % (tag, dt * 1e3, N / dt, same), flush=True)
