"""The kernels' own source — the CUDA translation unit the expression compiler emits for a skill plus
csrc/clik_pinv.cuh / clik_qp.cuh / clik_math.cuh — compiled for the HOST with g++ and a page of shims
(one "thread", grid-stride loop over all instances) and run on the reference-generated fixture
tests/golden/controller_vectors.json.  Like test_qp_core_host.py this is a harness for the device code's
logic in a container without a GPU, not a product path (the product has no CPU fallback); the GPU tests
run the same source as sm_100a code against the same fixture (test_gpu_golden.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import casclik_b200 as cc
from oracle_bridge import close, orc
from test_golden_controllers import PINV, QP, VECTORS, load_case, golden_velocities, golden_qp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r'''
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __constant__ static const
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n)
#define __shared__ static
struct D1 { unsigned x = 1, y = 1, z = 1; };
struct D0 { unsigned x = 0, y = 0, z = 0; };
static D1 gridDim, blockDim;
static D0 blockIdx, threadIdx;
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline int __double2int_rn(double x) { return (int)std::nearbyint(x); }
static inline int __double2hiint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) {
  unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double x; std::memcpy(&x, &b, 8); return x; }
static inline unsigned __activemask() { return 1u; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline int __all_sync(unsigned, int p) { return p; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
#define CLIK_NO_TMA_KERNEL 1
#define CLIK_GROUP 1   // sub-warp mapping with groups of one lane: same code path, one host "thread"
static inline void __syncthreads() {}
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
using std::fma; using std::fmax; using std::fmin; using std::fabs; using std::sqrt;
#define __CUDACC__ 1   // (after the standard headers) enables the shared-memory tail pass of clik_qp.cuh
'''


def _host_library(ctrl, tmp_path):
    """Lower + emit the skill (no nvcc: the host only needs the generated source), build it with g++."""
    from casclik_b200 import build
    saved = (build.compile_cubin, build.kernel_registers)
    build.compile_cubin = lambda source, tag="skill", **kw: (b"", os.path.join(str(tmp_path), "skill.cubin"))
    build.kernel_registers = lambda *a, **kw: None
    try:
        ctrl.setup_problem_functions(load=False)
    finally:
        build.compile_cubin, build.kernel_registers = saved
    text = ctrl.kernel_source.replace("__device__ const unsigned short", "static const unsigned short")
    src = tmp_path / "skill.cpp"
    src.write_text(SHIM + text)
    so = tmp_path / "skill.so"
    subprocess.run(["g++", "-O0", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                    "-I", os.path.join(ROOT, "casclik_b200", "csrc"), "-o", str(so), str(src)], check=True)
    return ctypes.CDLL(str(so))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _inputs(inp):
    c = lambda k: np.ascontiguousarray(inp[k], dtype=np.float64) if k in inp else None  # noqa: E731
    return c("t"), c("q"), c("x"), c("y")


@pytest.mark.parametrize("name", PINV)
def test_pinv_kernel_source_on_host_reproduces_the_reference_outputs(name, tmp_path):
    spec, inp, kwargs, outputs = load_case(name)
    ctrl = cc.PseudoInverseController(spec, **kwargs)
    lib = _host_library(ctrl, tmp_path)
    t, q, x, y = _inputs(inp)
    nq, N = q.shape
    nx = 0 if x is None else x.shape[0]
    qdot, xdot = np.full((nq, N), np.nan), (np.full((nx, N), np.nan) if nx else None)
    mode = np.full(N, -9, dtype=np.int32)
    lib.clik_pinv_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y), _p(qdot), _p(xdot), _p(mode))
    got = qdot if xdot is None else np.vstack([qdot, xdot])
    gv, gmode = golden_velocities(outputs)
    assert np.array_equal(mode, gmode)
    ok = close(got, gv, 1e-9, 1e-12)
    if "kitchen_sink" in name or "conv_last" in name:        # multi-task chains, DESIGN §5
        err = np.linalg.norm(got - gv, axis=0) / np.maximum(np.linalg.norm(gv, axis=0), 1e-300)
        assert err.max() < 1e-9 and ok.mean() > 0.98, (err.max(), ok.mean())
    else:
        assert ok.all(), np.abs(got - gv).max()
    ro = VECTORS[name].get("rollout")
    if ro:                                                   # the closed-loop kernel, same harness
        cfg, n = ro["config"], ro["config"]["n"]
        qs = np.ascontiguousarray(q[:, :n])
        xs = None if x is None else np.ascontiguousarray(x[:, :n])
        ys = None if y is None else np.ascontiguousarray(y[:, :n])
        ts = np.ascontiguousarray(t[:n])
        inf = float("inf")
        lib.clik_pinv_rollout_kernel(
            ctypes.c_longlong(n), ctypes.c_longlong(n), ctypes.c_int(cfg["steps"]), ctypes.c_double(cfg["dt"]), _p(ts), ctypes.c_int(1),
            _p(qs), _p(xs), _p(ys), ctypes.c_double(inf if cfg["max_speed"] is None else cfg["max_speed"]),
            ctypes.c_double(inf if cfg["max_virtual_speed"] is None else cfg["max_virtual_speed"]),
            None, None, None, None)
        want = np.array(ro["q_final"]).T
        assert np.abs(qs - want).max() <= 1e-7 * (1 + np.abs(want).max())


@pytest.mark.parametrize("name", QP)
def test_qp_kernel_source_on_host_reproduces_the_reference_minimisers(name, tmp_path):
    spec, inp, kwargs, outputs = load_case(name)
    ctrl = cc.ReactiveQPController(spec, **kwargs)
    lib = _host_library(ctrl, tmp_path)
    t, q, x, y = _inputs(inp)
    N = q.shape[1]
    gx, gh, gA, glb, gub = golden_qp(outputs)
    nqp = gx.shape[1]
    sol = np.full((nqp, N), np.nan)
    status = np.full(N, -9, dtype=np.int32)
    active = np.zeros((2, N), dtype=np.uint32)
    lib.clik_qp_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y), None, None, _p(sol),
                       _p(status), _p(active), ctypes.c_int(10 * (nqp + gA.shape[1])))
    assert np.all(status == 0)
    for i in range(N):
        assert np.abs(sol[:, i] - gx[i]).max() <= 1e-7 * (1 + np.abs(gx[i]).max())
        kk = orc.kkt_residuals(gh, gA[i], glb[i], gub[i], sol[:, i])
        assert kk["primal"] < 1e-6 and kk["stationarity"] < 1e-6 and kk["sign"] < 1e-6
    ro = VECTORS[name].get("rollout")
    if ro:                                                   # closed loop, warm-started from step to step
        cfg, n = ro["config"], ro["config"]["n"]
        qs = np.ascontiguousarray(q[:, :n])
        xs = None if x is None else np.ascontiguousarray(x[:, :n])
        ys = None if y is None else np.ascontiguousarray(y[:, :n])
        ts = np.ascontiguousarray(t[:n])
        failed = np.full(n, -9, dtype=np.int32)
        inf = float("inf")
        lib.clik_qp_rollout_kernel(
            ctypes.c_longlong(n), ctypes.c_longlong(n), ctypes.c_int(cfg["steps"]), ctypes.c_double(cfg["dt"]), _p(ts), ctypes.c_int(1),
            _p(qs), _p(xs), _p(ys), ctypes.c_double(inf if cfg["max_speed"] is None else cfg["max_speed"]),
            ctypes.c_double(inf if cfg["max_virtual_speed"] is None else cfg["max_virtual_speed"]),
            None, _p(failed), ctypes.c_int(10 * (nqp + gA.shape[1])))
        want = np.array(ro["q_final"]).T
        assert np.all(failed == 0) and np.abs(qs - want).max() <= 1e-7 * (1 + np.abs(want).max())
    if ctrl.kernel_meta.get("qp_split"):                     # prediction pass + tail pass, as the ABI launches them
        sol2, status2, active2 = np.full_like(sol, np.nan), np.full_like(status, -9), np.zeros_like(active)
        args = (ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y), None, None, _p(sol2), _p(status2),
                _p(active2), ctypes.c_int(10 * (nqp + gA.shape[1])))
        lib.clik_qp_fast_kernel(*args)
        assert set(np.unique(status2)) <= {0, 3}
        lib.clik_qp_tail_kernel(*args)
        assert np.array_equal(sol2, sol) and np.array_equal(status2, status) and np.array_equal(active2, active)


@pytest.mark.parametrize("name", ["ur5_track", "ur5_moe2016_pinv", "ur5_moe2016_multidim", "iiwa_multitask"])
def test_benchmark_scenarios_kernel_source_on_host_vs_oracle(name, tmp_path):
    """The BASELINE scenarios at a few thousand instances (all mode-selection paths: static modes, the
    closed-form unit-set modes of the 128-mode iiwa skill, multidim sets) against the oracle."""
    from casclik_b200 import scenarios
    from oracle_bridge import oracle_pinv
    sc = scenarios.get(name)
    ctrl = sc.make_controller()
    lib = _host_library(ctrl, tmp_path)
    N = 3000
    inp = {k: v for k, v in sc.sample(N, seed=17).items() if v is not None}
    t, q, x, y = _inputs(inp)
    nq = q.shape[0]
    qdot, mode = np.full((nq, N), np.nan), np.full(N, -9, dtype=np.int32)
    lib.clik_pinv_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y), _p(qdot), None, _p(mode))
    ref_v, ref_mode = oracle_pinv(sc.spec, inp, dict(sc.options) if sc.options else None)
    assert np.array_equal(mode, ref_mode)
    assert len(np.unique(ref_mode)) >= (1 if name == "ur5_track" else 2)
    assert close(qdot, ref_v, 1e-9, 1e-12).all(), np.abs(qdot - ref_v).max()


@pytest.mark.parametrize("name", ["ur5_qp", "ur5_moe2016_qp"])
def test_benchmark_qp_scenarios_kernel_source_on_host_vs_oracle(name, tmp_path):
    from casclik_b200 import scenarios
    from oracle_bridge import oracle_qp_problem
    sc = scenarios.get(name)
    ctrl = sc.make_controller()
    lib = _host_library(ctrl, tmp_path)
    N = 400
    inp = {k: v for k, v in sc.sample(N, seed=23).items() if v is not None}
    t, q, x, y = _inputs(inp)
    h, A, lb, ub = oracle_qp_problem(sc.spec, inp)
    nqp, m = A.shape[2], A.shape[1]
    sol, status = np.full((nqp, N), np.nan), np.full(N, -9, dtype=np.int32)
    active = np.zeros((2, N), dtype=np.uint32)
    lib.clik_qp_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y), None, None, _p(sol),
                       _p(status), _p(active), ctypes.c_int(10 * (nqp + m)))
    assert np.all(status == 0)
    for i in range(N):
        xo, lamo, sto = orc.solve_qp_single(h, A[i], lb[i], ub[i])
        assert sto == 0 and np.abs(sol[:, i] - xo).max() <= 1e-7 * (1 + np.abs(xo).max())
        up = sum(1 << r for r in range(m) if lamo[r] > 0)
        lo = sum(1 << r for r in range(m) if lamo[r] < 0)
        assert (int(active[0, i]), int(active[1, i])) == (up, lo), i


@pytest.mark.parametrize("env", [{"CLIK_UNIT_SETS": "0"}, {"CLIK_UNIT_SETS": "0", "CLIK_NSTATIC": "1"}],
                         ids=["29_static_then_dynamic", "all_dynamic"])
def test_mode_search_variants_agree_with_the_oracle(env, tmp_path, monkeypatch):
    """The mode-search implementations behind solve_instance — static register modes, the run-time search
    over row lists in local memory (dynamic_mode), and the closed-form unit-set modes (default for this
    skill, covered above) — on the 128-mode iiwa skill with many instances outside their limits.
    (All 128 modes on the static path is not tried: 128 distinct template instantiations take g++ forever.)"""
    from casclik_b200 import scenarios
    from oracle_bridge import oracle_pinv
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sc = scenarios.get("iiwa_multitask")
    ctrl = sc.make_controller()
    lib = _host_library(ctrl, tmp_path)
    assert not ctrl.kernel_meta["pinv_unit_sets"]
    assert ctrl.kernel_meta["pinv_static_modes"] == {"1": 1}.get(env.get("CLIK_NSTATIC"), 29)
    N = 1500
    inp = {k: v for k, v in sc.sample(N, seed=29).items() if v is not None}
    # push more joints over their limits than the benchmark distribution does: deeper modes
    rng = np.random.default_rng(5)
    lower, upper = np.array(sc.spec.constraints[0].set_min), np.array(sc.spec.constraints[0].set_max)
    for i in range(0, N, 3):
        j = rng.choice(7, size=rng.integers(1, 5), replace=False)
        inp["q"][j, i] = np.where(rng.random(len(j)) < 0.5, -3.2, 3.2)
    t, q, x, y = _inputs(inp)
    qdot, mode = np.full((7, N), np.nan), np.full(N, -9, dtype=np.int32)
    lib.clik_pinv_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y), _p(qdot), None, _p(mode))
    ref_v, ref_mode = oracle_pinv(sc.spec, inp)
    assert np.array_equal(mode, ref_mode)
    assert len(np.unique(ref_mode)) > 40 and ref_mode.max() > 60
    assert close(qdot, ref_v, 1e-9, 1e-12).all(), np.abs(qdot - ref_v).max()
    _check_split_and_group_passes(lib, ctrl, (t, q, x, y), qdot, mode, ref_v)


def test_dense_sets_take_the_dynamic_tail(tmp_path):
    """Four SetConstraints on dense expressions (16 modes: 11 static, 5 through dynamic_mode) + two tasks."""
    from casclik_b200 import cs
    from oracle_bridge import oracle_pinv
    t, q = cs.MX.sym("t"), cs.MX.sym("q", 5)
    sets = [cc.SetConstraint("s%d" % k, e, set_min=-0.3, set_max=0.3, priority=k, gain=2.0) for k, e in enumerate(
        [q[0] + 0.5 * q[1], cs.sin(q[1]) - q[2], q[2] * q[3], q[3] + q[4] - 0.2 * cs.cos(t)])]
    tasks = [cc.EqualityConstraint("a", cs.vertcat(q[0] - q[4], q[1] + q[2] - 0.1), priority=8),
             cc.VelocityEqualityConstraint("b", q[3] - q[0], target=0.1, priority=9)]
    spec = cc.SkillSpecification("dense_sets", t, q, constraints=sets + tasks)
    ctrl = cc.PseudoInverseController(spec)
    lib = _host_library(ctrl, tmp_path)
    assert ctrl.n_modes == 16 and ctrl.kernel_meta["pinv_static_modes"] == 11 and not ctrl.kernel_meta["pinv_unit_sets"]
    N = 4000
    rng = np.random.default_rng(8)
    inp = {"t": rng.uniform(0, 5, N), "q": rng.uniform(-0.9, 0.9, (5, N))}
    t_, q_, _, _ = _inputs(inp)
    qdot, mode = np.full((5, N), np.nan), np.full(N, -9, dtype=np.int32)
    lib.clik_pinv_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t_), ctypes.c_int(1), _p(q_), None, None, _p(qdot), None, _p(mode))
    ref_v, ref_mode = oracle_pinv(spec, inp)
    assert np.array_equal(mode, ref_mode)
    assert (ref_mode >= 11).sum() > 20 and (ref_mode == -1).sum() >= 0
    ok = close(qdot, ref_v, 1e-9, 1e-12)
    err = np.linalg.norm(qdot - ref_v, axis=0) / np.maximum(np.linalg.norm(ref_v, axis=0), 1e-300)
    assert ok.mean() > 0.995 and err[np.isfinite(err)].max() < 1e-8, (ok.mean(), err.max())
    _check_split_and_group_passes(lib, ctrl, (t_, q_, None, None), qdot, mode, ref_v)


def _check_split_and_group_passes(lib, ctrl, tqxy, qdot_full, mode_full, ref_v):
    """The two-launch form (clik_pinv_fast_kernel + clik_pinv_group_kernel on what it left pending) and the
    whole step in the sub-warp mapping (group kernel from mode 0; groups of one lane on the host) against
    the single full kernel: same flags, velocities equal to rounding."""
    t, q, x, y = tqxy
    assert ctrl.kernel_meta["pinv_split"] and ctrl.kernel_meta["pinv_group"]
    nq, N = q.shape
    n_static = ctrl.kernel_meta["pinv_static_modes"]
    v2, m2 = np.full((nq, N), np.nan), np.full(N, -9, dtype=np.int32)
    head = (ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y), _p(v2), None, _p(m2))
    lib.clik_pinv_fast_kernel(*head)
    pending = m2 == -2
    assert pending.sum() == (mode_full >= n_static).sum() + (mode_full == -1).sum() and pending.any()
    assert np.array_equal(m2[~pending], mode_full[~pending]) and np.array_equal(v2[:, ~pending], qdot_full[:, ~pending])
    lib.clik_pinv_group_kernel(*head, ctypes.c_int(n_static), ctypes.c_int(1))
    assert np.array_equal(m2, mode_full)
    scale = np.maximum(np.linalg.norm(ref_v, axis=0), 1e-12)
    assert (np.linalg.norm(v2 - qdot_full, axis=0) / scale).max() < 1e-8
    v3, m3 = np.full((nq, N), np.nan), np.full(N, -9, dtype=np.int32)
    lib.clik_pinv_group_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y),
                               _p(v3), None, _p(m3), ctypes.c_int(0), ctypes.c_int(0))
    assert np.array_equal(m3, mode_full)
    assert (np.linalg.norm(v3 - ref_v, axis=0) / scale).max() < 1e-8


@pytest.mark.parametrize("seed", range(12))
def test_fuzzed_skills_kernel_source_on_host_vs_oracle(seed, tmp_path):
    """Random skills (tests/fuzz_skills.py; `tools/fuzz_skills.py A B` runs any seed range — 110 seeds were
    clean when this was written): modes bit-exact, velocities at rounding level unless the random tasks
    are over-determined, where the damped normal equations lose digits for every implementation."""
    from fuzz_skills import make_skill
    from oracle_bridge import oracle_pinv
    spec, opts, inp = make_skill(seed)
    ctrl = cc.PseudoInverseController(spec, options=dict(opts))
    lib = _host_library(ctrl, tmp_path)
    t, q, x, y = _inputs(inp)
    nq, N = q.shape
    nx = ctrl._nx
    if nx and x is None:
        x = np.zeros((nx, N))
        inp = dict(inp, x=x)
    qdot, xdot = np.full((nq, N), np.nan), (np.full((nx, N), np.nan) if nx else None)
    mode = np.full(N, -9, dtype=np.int32)
    lib.clik_pinv_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y if ctrl._ny else None),
                         _p(qdot), _p(xdot), _p(mode))
    got = qdot if xdot is None else np.vstack([qdot, xdot])
    ref_v, ref_mode = oracle_pinv(spec, inp, dict(opts))
    assert np.array_equal(mode, ref_mode)
    err = np.linalg.norm(got - ref_v, axis=0) / np.maximum(np.linalg.norm(ref_v, axis=0), 1e-12)
    assert err.max() < 1e-7, err.max()


@pytest.mark.parametrize("seed", list(range(8)) + ["dense0", "dense1", "dense2"])
def test_fuzzed_qp_skills_kernel_source_on_host_vs_oracle(seed, tmp_path):
    """Random QP skills (hard / soft rows of every constraint class, random weights; some are infeasible for
    part of the batch): status, minimiser and working-set masks against the oracle's solve of the same problem.
    The "dense" seeds have 8-10 dense rows and 12-15 variables: the generic in-kernel solver."""
    from fuzz_skills import make_qp_skill, make_dense_qp_skill
    from oracle_bridge import oracle_qp_problem
    dense = isinstance(seed, str)
    spec, weights, inp = make_dense_qp_skill(int(seed[5:])) if dense else make_qp_skill(seed)
    ctrl = cc.ReactiveQPController(spec, **weights)
    lib = _host_library(ctrl, tmp_path)
    assert bool(ctrl.kernel_meta["qp_structured"]) != dense
    t, q, x, y = _inputs(inp)
    N = q.shape[1]
    if ctrl._nxv and x is None:
        x = np.zeros((ctrl._nxv, N))
        inp = dict(inp, x=x)
    w = {"w_rob": weights["robot_var_weights"]} if "robot_var_weights" in weights else {}
    h, A, lb, ub = oracle_qp_problem(spec, inp, **w)
    m, nqp = A.shape[1], A.shape[2]
    sol, status = np.full((nqp, N), np.nan), np.full(N, -9, dtype=np.int32)
    active = np.zeros((2, N), dtype=np.uint32)
    lib.clik_qp_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), _p(x), _p(y if ctrl._ny else None), None, None,
                       _p(sol), _p(status), _p(active), ctypes.c_int(10 * (nqp + m)))
    for i in range(N):
        xo, lamo, sto = orc.solve_qp_single(h, A[i], lb[i], ub[i])
        assert sto == int(status[i]), i
        if sto == 0:
            assert np.abs(sol[:, i] - xo).max() <= 1e-7 * (1 + np.abs(xo).max()), i
            up = sum(1 << r for r in range(m) if lamo[r] > 0)
            lo = sum(1 << r for r in range(m) if lamo[r] < 0)
            assert (int(active[0, i]), int(active[1, i])) == (up, lo), i


@pytest.mark.parametrize("seed", range(8))
def test_fuzzed_option_skills_kernel_source_on_host_vs_oracle(seed, tmp_path):
    """Random skills for the experimental options: vector-valued sets with multidim_sets (even seeds),
    a final SetConstraint with converge_final_set_to_max (odd seeds).  Modes bit-exact; 1e-6 on the
    velocities because random 3-row boxes on 3 joints are over-determined (24 seeds: worst 2e-7)."""
    from fuzz_skills import make_option_skill
    from oracle_bridge import oracle_pinv
    spec, opts, inp = make_option_skill(seed)
    ctrl = cc.PseudoInverseController(spec, options=dict(opts))
    lib = _host_library(ctrl, tmp_path)
    t, q, x, y = _inputs(inp)
    nq, N = q.shape
    qdot, mode = np.full((nq, N), np.nan), np.full(N, -9, dtype=np.int32)
    lib.clik_pinv_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(q), None, None, _p(qdot), None, _p(mode))
    ref_v, ref_mode = oracle_pinv(spec, inp, dict(opts))
    assert np.array_equal(mode, ref_mode) and len(np.unique(ref_mode)) >= 2
    err = np.linalg.norm(qdot - ref_v, axis=0) / np.maximum(np.linalg.norm(ref_v, axis=0), 1e-12)
    assert err.max() < 1e-6, err.max()


def test_qp_kernel_ignores_input_rows_the_program_does_not_read(tmp_path):
    """The host entry points upload only the rows the compiled program reads (clik_abi.cu rowmask), so
    whatever sits in the other rows of the device copy must influence neither the solution nor the
    NaN / inf screening of the inputs (status 4)."""
    from casclik_b200 import scenarios
    sc = scenarios.get("ur5_qp")
    ctrl = sc.make_controller()
    lib = _host_library(ctrl, tmp_path)
    tmask, qmask, _, _ = ctrl.kernel_meta["qp_read_masks"]
    assert tmask == 0 and qmask == 63, "this tracking QP reads every joint and not the time"
    N = 64
    inp = {k: v for k, v in sc.sample(N, seed=0).items() if v is not None}
    t, q, x, y = _inputs(inp)
    res = []
    for poison in (False, True):
        qq, tt = q.copy(), t.copy()
        if poison:
            tt[:] = np.nan
        sol, status = np.full((9, N), np.nan), np.full(N, -9, dtype=np.int32)
        active = np.zeros((2, N), dtype=np.uint32)
        lib.clik_qp_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(tt), ctypes.c_int(1), _p(qq), _p(x), _p(y), None,
                           None, _p(sol), _p(status), _p(active), ctypes.c_int(240))
        assert np.all(status == 0)
        res.append(sol)
    assert np.array_equal(res[0], res[1])
    qq = q.copy()
    qq[0, 3] = np.inf                                  # a row that IS read: screened
    lib.clik_qp_kernel(ctypes.c_longlong(N), ctypes.c_longlong(N), _p(t), ctypes.c_int(1), _p(qq), _p(x), _p(y), None,
                       None, _p(sol), _p(status), _p(active), ctypes.c_int(240))
    assert status[3] == 4 and np.all(np.delete(status, 3) == 0)
