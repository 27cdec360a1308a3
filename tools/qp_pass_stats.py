"""Host statistics of the QP working-set prediction (csrc/clik_qp.cuh crash_guess): the kernels' own source compiled
for the host with -DCLIK_QP_STATS (the harness of tests/test_kernel_code_on_host.py), one "thread" per instance, so the
histogram is per instance: how many all-at-once passes an instance needs until its set stops changing, how many enter
the one-row passes and the Goldfarb-Idnani iteration.   python tools/qp_pass_stats.py <scenario> <instances>"""
import os, sys, ctypes, pathlib, numpy as np, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_kernel_code_on_host as H
from casclik_b200 import scenarios, build
name = sys.argv[1]; N = int(sys.argv[2])
sc = scenarios.get(name); ctrl = sc.make_controller()
tmp = pathlib.Path(tempfile.mkdtemp(prefix="qp_stats_"))
saved = (build.compile_cubin, build.kernel_registers)
build.compile_cubin = lambda source, tag="skill", **kw: (b"", str(tmp / "skill.cubin"))
build.kernel_registers = lambda *a, **kw: None
ctrl.setup_problem_functions(load=False)
build.compile_cubin, build.kernel_registers = saved
text = ctrl.kernel_source.replace("__device__ const unsigned short", "static const unsigned short")
extra = '''
extern "C" void get_stats(long long* out) { for (int i = 0; i < 8; ++i) out[i] = clik::qp_stats[i]; for (int k = 0; k < 2; ++k) for (int i = 0; i < 64; ++i) out[8 + k * 64 + i] = clik::qp_pass_hist[k][i]; }
'''
(tmp / "skill.cpp").write_text("#define CLIK_QP_STATS 1\n" + H.SHIM + text + extra)
subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", os.path.join(ROOT, "casclik_b200", "csrc"),
                "-o", str(tmp / "skill.so"), str(tmp / "skill.cpp")], check=True)
lib = ctypes.CDLL(str(tmp / "skill.so"))
inp = {k: v for k, v in sc.sample(N, seed=5).items() if v is not None}
t, q, x, y = H._inputs(inp)
nqp, m = ctrl.kernel_meta["qp_n"], ctrl.kernel_meta["qp_m"]
sol, status = np.full((nqp, N), np.nan), np.full(N, -9, dtype=np.int32)
active = np.zeros((2, N), dtype=np.uint32)
lib.clik_qp_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), H._p(t), ctypes.c_int(1), H._p(q), H._p(x), H._p(y), None, None, H._p(sol), H._p(status), H._p(active), ctypes.c_int(10 * (nqp + m)))
out = (ctypes.c_longlong * 136)(); lib.get_stats(out)
o = np.array(out[:])
print("status", np.unique(status, return_counts=True), "GI entries", o[0], "single phases", o[1])
print("all-at-once passes hist", {i: int(v) for i, v in enumerate(o[8:72]) if v})
print("one-row passes hist", {i: int(v) for i, v in enumerate(o[72:136]) if v})
