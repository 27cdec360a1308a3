"""The QP kernel's algorithm (csrc/clik_qp.cuh: qp_dual_active_set) compiled for the HOST with g++
and a few shims, exercised against the oracle on problem classes that are awkward to reach through
a robot skill: +-inf and +-1e10 bounds, equality rows, duplicated and linearly dependent rows, more
rows than variables, infeasible sets.  This is a test harness for the device code's logic, not a
product path (the product has no CPU fallback)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle_bridge import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r'''
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __restrict__
struct D3 { unsigned x = 1; };
static D3 gridDim, blockDim, blockIdx, threadIdx;
static inline double __ldcs(const double* p) { return *p; }
static inline void __stcs(double* p, double v) { *p = v; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline unsigned __activemask() { return 1u; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline int __double2hiint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
using std::fma; using std::fmax; using std::fmin; using std::fabs; using std::sqrt;
#include "%s"
extern "C" int host_qp(int nx, int m, double* A, const double* lb, const double* ub, const double* h,
                       double* x, unsigned* au, unsigned* al, int max_iter) {
  return clik::qp_dual_active_set<16, 32>(nx, m, A, lb, ub, h, nullptr, x, au, al, max_iter);
}
'''


@pytest.fixture(scope="module")
def host_qp(tmp_path_factory):
    d = tmp_path_factory.mktemp("qp_host")
    src = d / "host_qp.cpp"
    src.write_text(SHIM % os.path.join(ROOT, "casclik_b200", "csrc", "clik_qp.cuh"))
    so = d / "host_qp.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))

    def solve(h, A, lb, ub, max_iter=400):
        m, n = A.shape
        Ac = np.ascontiguousarray(A, dtype=np.float64).copy()
        x = np.zeros(n)
        au, al = ctypes.c_uint(), ctypes.c_uint()
        p = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        lbc, ubc, hc = (np.ascontiguousarray(v, dtype=np.float64) for v in (lb, ub, h))
        st = lib.host_qp(n, m, Ac.ctypes.data_as(ctypes.c_void_p), lbc.ctypes.data_as(ctypes.c_void_p),
                         ubc.ctypes.data_as(ctypes.c_void_p), hc.ctypes.data_as(ctypes.c_void_p),
                         x.ctypes.data_as(ctypes.c_void_p), ctypes.byref(au), ctypes.byref(al), max_iter)
        return x, st, au.value, al.value
    return solve


def _masks(lam):
    up = sum(1 << r for r in range(len(lam)) if lam[r] > 0)
    lo = sum(1 << r for r in range(len(lam)) if lam[r] < 0)
    return up, lo


def _check(host_qp, h, A, lb, ub, flags=True):
    x, st, au, al = host_qp(h, A, lb, ub)
    xo, lamo, sto = orc.solve_qp_single(h, A, lb, ub)
    assert st == 0 and sto == 0
    kk = orc.kkt_residuals(h, A, lb, ub, x)
    assert kk["primal"] < 1e-9 and kk["stationarity"] < 1e-9 and kk["sign"] < 1e-9, kk
    assert np.abs(x - xo).max() < 1e-9 * (1 + np.abs(xo).max())
    if flags:
        assert (au, al) == _masks(lamo)


def test_random_problems_with_infinite_huge_and_equality_bounds(host_qp):
    rng = np.random.default_rng(5)
    for trial in range(300):
        n, m = int(rng.integers(1, 10)), int(rng.integers(1, 16))
        h = rng.uniform(0.001, 2.0, n)
        A = rng.normal(size=(m, n))
        r = A @ rng.normal(size=n)
        lb, ub = r - rng.uniform(0, 1, m), r + rng.uniform(0, 1, m)
        for i in range(m):
            c = rng.random()
            if c < 0.1:
                lb[i] = -np.inf
            elif c < 0.2:
                ub[i] = np.inf
            elif c < 0.3:
                lb[i], ub[i] = -1e10, 1e10
            elif c < 0.4:
                lb[i] = ub[i] = r[i]
        _check(host_qp, h, A, lb, ub)


def test_duplicate_and_dependent_rows(host_qp):
    rng = np.random.default_rng(7)
    for trial in range(100):
        n = int(rng.integers(2, 8))
        A0 = rng.normal(size=(4, n))
        A = np.vstack([A0, A0[0], 2.0 * A0[1], A0[2] + A0[3]])      # rows 4..6 depend on rows 0..3
        x_in = rng.normal(size=n)
        r = A @ x_in
        lb, ub = r - rng.uniform(0.01, 1, 7), r + rng.uniform(0.01, 1, 7)
        h = rng.uniform(0.01, 1.0, n)
        x, st, au, al = host_qp(h, A, lb, ub)
        assert st == 0
        kk = orc.kkt_residuals(h, A, lb, ub, x)
        assert kk["primal"] < 1e-9
        xo, _, sto = orc.solve_qp_single(h, A, lb, ub)
        assert sto == 0 and np.abs(x - xo).max() < 1e-8 * (1 + np.abs(xo).max())


def test_more_active_rows_than_variables_and_infeasible(host_qp):
    # box in 2-D cut by many half-planes: feasible, vertex solutions with degenerate ties
    A = np.array([[1., 0.], [0., 1.], [1., 1.], [1., -1.], [2., 1.], [1., 2.]])
    lb = np.array([1., 1., 2., -5., 3., 3.])
    ub = np.full(6, 50.0)
    x, st, _, _ = host_qp(np.array([1.0, 1.0]), A, lb, ub)
    assert st == 0 and np.allclose(x, [1.0, 1.0], atol=1e-12)
    # contradictory rows -> infeasible status, never a wrong "solved"
    x, st, _, _ = host_qp(np.ones(2), np.array([[1., 0.], [1., 0.]]), np.array([1., -3.]), np.array([2., -2.]))
    assert st == 2
    # unconstrained optimum already feasible: zero iterations, nothing active
    x, st, au, al = host_qp(np.ones(3), np.eye(3), -np.ones(3), np.ones(3))
    assert st == 0 and np.all(x == 0.0) and au == 0 and al == 0
    # iteration cap is reported
    rng = np.random.default_rng(0)
    A = rng.normal(size=(12, 6))
    r = A @ rng.normal(size=6)
    x, st, _, _ = host_qp(np.ones(6), A, r - 0.1, r + 0.1)          # feasible (x_in), far from 0
    assert st == 0 and orc.kkt_residuals(np.ones(6), A, r - 0.1, r + 0.1, x)["primal"] < 1e-9
    _, st1, _, _ = host_qp(np.ones(6), A, r - 0.1, r + 0.1, 1)
    assert st1 == 1
    # 12 narrow bands in 6-D that exclude each other: both solvers must say infeasible
    _, st2, _, _ = host_qp(np.ones(6), A, r + 0.5, r + 1.0)
    assert st2 == orc.solve_qp_single(np.ones(6), A, r + 0.5, r + 1.0)[2] == 2


# ---- structured (register-resident) variant ---------------------------------------------------------
SHIM_S = SHIM.split('#include "%s"')[0] + r'''
#include <cstring>
#include "%s"
struct S {
  static constexpr int QN = 6, QMD = 2, QMU = 7, QM = 9;
  static constexpr int QP_FLIP_PASSES = 2;
  static constexpr bool ad_nz(int, int) { return true; }
  static constexpr int dense_row(int a) { return a == 0 ? 1 : 4; }
  static constexpr int unit_row(int i) { constexpr int t[7] = {0, 2, 3, 5, 6, 7, 8}; return t[i]; }
  static constexpr int unit_col(int i) { constexpr int t[7] = {0, 1, 2, 0, 3, 5, 1}; return t[i]; }
  static constexpr double unit_coef(int i) { constexpr double t[7] = {1.0, 1.0, -1.0, 2.0, 1.0, -0.5, 1.0}; return t[i]; }
};
extern "C" int host_crash(const double* Ad, const double* lbd, const double* ubd, const double* lbu,
                          const double* ubu, const double* s, unsigned* wu /*in/out*/, unsigned* wl, double* x) {
  clik::QpSData<S> d;
  std::memcpy(d.Ad, Ad, sizeof(d.Ad)); std::memcpy(d.lbd, lbd, sizeof(d.lbd)); std::memcpy(d.ubd, ubd, sizeof(d.ubd));
  std::memcpy(d.lbu, lbu, sizeof(d.lbu)); std::memcpy(d.ubu, ubu, sizeof(d.ubu)); std::memcpy(d.s, s, sizeof(d.s));
  double xs[S::QN];
  const bool certified = clik::crash_guess<S>(d, wu, wl, xs);
  for (int j = 0; j < S::QN; ++j) x[j] = xs[j];
  return certified ? 1 : 0;
}
extern "C" int host_crash_single(const double* Ad, const double* lbd, const double* ubd, const double* lbu,
                                 const double* ubu, const double* s, unsigned* wu /*in/out*/, unsigned* wl, double* x) {
  // what the non-fast kernels do after the all-at-once budget: one-row-per-pass passes from the last set
  clik::QpSData<S> d;
  std::memcpy(d.Ad, Ad, sizeof(d.Ad)); std::memcpy(d.lbd, lbd, sizeof(d.lbd)); std::memcpy(d.ubd, ubd, sizeof(d.ubd));
  std::memcpy(d.lbu, lbu, sizeof(d.lbu)); std::memcpy(d.ubu, ubu, sizeof(d.ubu)); std::memcpy(d.s, s, sizeof(d.s));
  double xs[S::QN];
  bool certified = clik::crash_guess<S>(d, wu, wl, xs);
  int stage = certified ? 1 : 0;
  if (!certified) {
    certified = clik::crash_guess<S, clik::CRASH_SINGLE_PASSES, true>(d, wu, wl, xs);
    stage = certified ? 2 : 0;
  }
  for (int j = 0; j < S::QN; ++j) x[j] = xs[j];
  return stage;
}
extern "C" int host_qps(const double* Ad, const double* lbd, const double* ubd, const double* lbu,
                        const double* ubu, const double* s, double* x, unsigned* au, unsigned* al, int max_iter,
                        unsigned wu, unsigned wl) {
  clik::QpSData<S> d;
  std::memcpy(d.Ad, Ad, sizeof(d.Ad)); std::memcpy(d.lbd, lbd, sizeof(d.lbd)); std::memcpy(d.ubd, ubd, sizeof(d.ubd));
  std::memcpy(d.lbu, lbu, sizeof(d.lbu)); std::memcpy(d.ubu, ubu, sizeof(d.ubu)); std::memcpy(d.s, s, sizeof(d.s));
  double xs[S::QN];
  int st = clik::qp_structured<S>(d, xs, au, al, max_iter, wu, wl);
  for (int j = 0; j < S::QN; ++j) x[j] = xs[j];
  return st;
}
'''
DENSE_ROWS = [1, 4]
UNIT = [(0, 0, 1.0), (2, 1, 1.0), (3, 2, -1.0), (5, 0, 2.0), (6, 3, 1.0), (7, 5, -0.5), (8, 1, 1.0)]   # (row, col, coef)


@pytest.fixture(scope="module")
def host_qps(tmp_path_factory):
    d = tmp_path_factory.mktemp("qps_host")
    src = d / "host_qps.cpp"
    src.write_text(SHIM_S % os.path.join(ROOT, "casclik_b200", "csrc", "clik_qp.cuh"))
    so = d / "host_qps.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))

    def solve(h, A, lb, ub, max_iter=400, warm=(0, 0)):
        Ad = np.ascontiguousarray(A[DENSE_ROWS])
        ur = [r for r, _, _ in UNIT]
        arrs = [Ad, lb[DENSE_ROWS], ub[DENSE_ROWS], lb[ur], ub[ur], 1.0 / np.sqrt(h)]
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
        x = np.zeros(6)
        au, al = ctypes.c_uint(), ctypes.c_uint()
        st = lib.host_qps(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs],
                          x.ctypes.data_as(ctypes.c_void_p), ctypes.byref(au), ctypes.byref(al), max_iter,
                          ctypes.c_uint(warm[0]), ctypes.c_uint(warm[1]))
        return x, st, au.value, al.value

    def crash(h, A, lb, ub, seed=(0, 0)):
        ur = [r for r, _, _ in UNIT]
        arrs = [A[DENSE_ROWS], lb[DENSE_ROWS], ub[DENSE_ROWS], lb[ur], ub[ur], 1.0 / np.sqrt(h)]
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
        wu, wl = ctypes.c_uint(seed[0]), ctypes.c_uint(seed[1])
        x = np.zeros(6)
        crash.certified = bool(lib.host_crash(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs], ctypes.byref(wu),
                                              ctypes.byref(wl), x.ctypes.data_as(ctypes.c_void_p)))
        crash.x = x
        return wu.value, wl.value
    solve.crash = crash

    def crash_single(h, A, lb, ub):
        ur = [r for r, _, _ in UNIT]
        arrs = [A[DENSE_ROWS], lb[DENSE_ROWS], ub[DENSE_ROWS], lb[ur], ub[ur], 1.0 / np.sqrt(h)]
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
        wu, wl = ctypes.c_uint(0), ctypes.c_uint(0)
        x = np.zeros(6)
        stage = lib.host_crash_single(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs], ctypes.byref(wu),
                                      ctypes.byref(wl), x.ctypes.data_as(ctypes.c_void_p))
        return stage, x, wu.value, wl.value
    solve.crash_single = crash_single
    return solve


def _structured_problem(rng, eq_prob=0.3, tight=False):
    n, m = 6, 9
    A = np.zeros((m, n))
    A[DENSE_ROWS] = rng.normal(size=(2, n))
    for r, c, k in UNIT:
        A[r, c] = k
    h = rng.uniform(0.001, 2.0, n)
    x_in = rng.normal(size=n)
    r = A @ x_in
    w = 0.05 if tight else 1.0
    lb, ub = r - rng.uniform(0, w, m), r + rng.uniform(0, w, m)
    for i in range(m):
        c = rng.random()
        if c < 0.1:
            lb[i] = -np.inf
        elif c < 0.2:
            ub[i] = np.inf
        elif c < 0.3:
            lb[i], ub[i] = -1e10, 1e10
        elif c < 0.3 + 0.1 * eq_prob and i in DENSE_ROWS:
            lb[i] = ub[i] = r[i]
    return h, A, lb, ub


def test_structured_solver_matches_oracle_and_dense_solver(host_qp, host_qps):
    rng = np.random.default_rng(11)
    for trial in range(400):
        h, A, lb, ub = _structured_problem(rng, tight=(trial % 3 == 0))
        x, st, au, al = host_qps(h, A, lb, ub)
        xo, lamo, sto = orc.solve_qp_single(h, A, lb, ub)
        xg, stg, aug, alg = host_qp(h, A, lb, ub)
        assert st == sto == stg == 0, (trial, st, sto, stg)
        kk = orc.kkt_residuals(h, A, lb, ub, x)
        assert kk["primal"] < 1e-9 and kk["stationarity"] < 1e-9 and kk["sign"] < 1e-9, (trial, kk)
        assert np.abs(x - xo).max() < 1e-9 * (1 + np.abs(xo).max()), trial
        assert (au, al) == _masks(lamo) == (aug, alg), trial


def test_structured_solver_infeasible_and_cap(host_qps):
    rng = np.random.default_rng(2)
    h, A, lb, ub = _structured_problem(rng)
    lb2, ub2 = lb.copy(), ub.copy()
    lb2[0], ub2[0] = 1.0, 2.0          # x0 in [1, 2]
    lb2[5], ub2[5] = -8.0, -6.0        # 2*x0 in [-8, -6]  -> contradiction on the same coordinate
    _, st, _, _ = host_qps(h, A, lb2, ub2)
    assert st == 2 and orc.solve_qp_single(h, A, lb2, ub2)[2] == 2
    lb3, ub3 = lb.copy(), ub.copy()
    lb3[:] = np.where(np.isfinite(lb3), lb3, -50.0) + 3.0
    ub3[:] = lb3 + 0.5
    ok = orc.solve_qp_single(h, A, lb3, ub3)[2]
    _, st3, _, _ = host_qps(h, A, lb3, ub3)
    assert st3 == ok
    _, st1, _, _ = host_qps(h, A, lb3, ub3, 1)
    assert st1 in (1, 2)


def test_structured_solver_warm_start_never_changes_the_answer(host_qps):
    """Working-set guesses: the exact final set, the final set of a perturbed problem (what a
    rollout hands over), and random garbage.  Same solution and flags as the cold start; a guess
    that cannot be repaired makes the solver report failure at worst (callers then restart cold),
    never a wrong 'solved'."""
    rng = np.random.default_rng(21)
    n_warm_ok = 0
    for trial in range(300):
        h, A, lb, ub = _structured_problem(rng, tight=(trial % 2 == 0))
        x, st, au, al = host_qps(h, A, lb, ub)
        assert st == 0
        guesses = [(au, al)]
        dl = rng.normal(scale=0.02, size=9)
        xp, stp, aup, alp = host_qps(h, A + 0.0, lb + dl, ub + dl)
        if stp == 0:
            guesses.append((aup, alp))
        guesses.append((int(rng.integers(0, 512)), 0))
        up_bits = int(rng.integers(0, 512))
        guesses.append((up_bits, int(rng.integers(0, 512)) & ~up_bits))
        for g in guesses:
            xw, stw, auw, alw = host_qps(h, A, lb, ub, warm=g)
            if stw != 0:
                continue                      # unrepaired guess: the kernels retry cold
            n_warm_ok += 1
            assert np.abs(xw - x).max() < 1e-9 * (1 + np.abs(x).max()), (trial, g)
            kk = orc.kkt_residuals(h, A, lb, ub, xw)
            assert kk["primal"] < 1e-9 and kk["stationarity"] < 1e-9 and kk["sign"] < 1e-9, (trial, g, kk)
            assert (auw, alw) == (au, al), (trial, g)
    assert n_warm_ok > 1000


def test_single_change_passes_settle_what_the_all_at_once_passes_leave_and_certify_only_minimisers(host_qps):
    """crash_guess<.., SINGLE>: after the all-at-once budget, one row changes per pass.  Whatever it certifies
    is the minimiser with the solver's working set; it leaves (almost) nothing for the iteration."""
    rng = np.random.default_rng(71)
    stages = {0: 0, 1: 0, 2: 0}
    n_feasible = 0
    for trial in range(600):
        h, A, lb, ub = _structured_problem(rng, eq_prob=3.0 if trial % 2 else 0.3, tight=(trial % 3 == 0))
        x, st, au, al = host_qps(h, A, lb, ub)
        stage, xc, wu, wl = host_qps.crash_single(h, A, lb, ub)
        if st == 0:
            n_feasible += 1
            stages[stage] += 1
        if stage:
            assert st == 0 and (wu, wl) == (au, al), trial
            assert np.abs(xc - x).max() < 1e-9 * (1 + np.abs(x).max()), trial
            kk = orc.kkt_residuals(h, A, lb, ub, xc)
            assert kk["primal"] < 1e-9 and kk["stationarity"] < 1e-9 and kk["sign"] < 1e-9, (trial, kk)
    assert stages[2] > 0                                   # the single-change passes did rescue some
    # (what is left on these deliberately nasty problems — dependent rows, two rows of one column active,
    # contradictory bounds — is the Goldfarb-Idnani iteration's job)
    assert stages[0] <= 0.1 * n_feasible, stages


def test_crash_start_guess_is_sound_and_never_changes_the_answer(host_qps):
    """The working-set guess used for cold solves (crash_guess: primal-dual active-set passes over
    equality + single-variable rows): every equality row on one side, at most one single-variable row
    per column; started from it the solver returns the cold-start answer and flags; and for most
    problems it already IS the final working set."""
    rng = np.random.default_rng(33)
    n_guess, n_ok, n_exact, n_cert = 0, 0, 0, 0
    for trial in range(400):
        h, A, lb, ub = _structured_problem(rng, eq_prob=3.0 if trial % 2 else 0.3, tight=(trial % 3 == 0))
        x, st, au, al = host_qps(h, A, lb, ub)
        wu, wl = host_qps.crash(h, A, lb, ub)
        certified, xc = host_qps.crash.certified, host_qps.crash.x.copy()
        assert wu & wl == 0
        if certified:
            # a prediction that certifies itself IS the answer (the kernels return it without iterating)
            assert st == 0 and (wu, wl) == (au, al), trial
            assert np.abs(xc - x).max() < 1e-9 * (1 + np.abs(x).max()), trial
            kk = orc.kkt_residuals(h, A, lb, ub, xc)
            assert kk["primal"] < 1e-9 and kk["stationarity"] < 1e-9 and kk["sign"] < 1e-9, (trial, kk)
            n_cert += 1
        for r in DENSE_ROWS:
            assert (wu | wl) >> r & 1 or lb[r] != ub[r]                   # every equality row is held
        cols = [c for r, c, k in UNIT if (wu | wl) >> r & 1]
        assert len(cols) == len(set(cols))                                # one fixed row per variable
        for r in range(9):                                                # never a row without that bound
            assert not ((wu >> r & 1) and not np.isfinite(ub[r])) and not ((wl >> r & 1) and not np.isfinite(lb[r]))
        if st != 0:
            continue
        n_guess += 1
        xw, stw, auw, alw = host_qps(h, A, lb, ub, warm=(wu, wl))
        if stw != 0:
            continue                                     # the kernels retry cold
        n_ok += 1
        assert np.abs(xw - x).max() < 1e-9 * (1 + np.abs(x).max()), trial
        assert (auw, alw) == (au, al), trial
        n_exact += int((wu | wl) == (au | al))
        # seeded with the solution's set it confirms it; seeded with garbage it is still a sound guess
        assert host_qps.crash(h, A, lb, ub, seed=(au, al)) == (au, al) or (wu | wl) != (au | al)
        gu = int(rng.integers(0, 512))
        su, sl = host_qps.crash(h, A, lb, ub, seed=(gu, int(rng.integers(0, 512)) & ~gu))
        assert su & sl == 0
        xs, sts, aus, als = host_qps(h, A, lb, ub, warm=(su, sl))
        if sts == 0:
            assert np.abs(xs - x).max() < 1e-9 * (1 + np.abs(x).max()) and (aus, als) == (au, al), trial
    assert n_ok > 0.95 * n_guess and n_exact > 0.8 * n_guess and n_cert > 0.8 * n_guess, (n_guess, n_ok, n_exact, n_cert)


# ---- the structured solver + working-set prediction on other row structures -------------------------------
SHIM_G = SHIM.split('#include "%s"')[0] + r'''
#include <cstring>
#include "%(hdr)s"
struct S {
  static constexpr int QN = %(qn)d, QMD = %(md)d, QMU = %(mu)d, QM = %(qm)d;
  static constexpr int QP_FLIP_PASSES = 2;
  static constexpr bool ad_nz(int, int) { return true; }
  static constexpr int dense_row(int a) { constexpr int t[%(md1)d] = {%(dense)s}; return t[a]; }
  static constexpr int unit_row(int i) { constexpr int t[%(mu1)d] = {%(urow)s}; return t[i]; }
  static constexpr int unit_col(int i) { constexpr int t[%(mu1)d] = {%(ucol)s}; return t[i]; }
  static constexpr double unit_coef(int i) { constexpr double t[%(mu1)d] = {%(ucoef)s}; return t[i]; }
};
static void fill(clik::QpSData<S>& d, const double* Ad, const double* lbd, const double* ubd, const double* lbu,
                 const double* ubu, const double* s) {
  std::memcpy(d.Ad, Ad, sizeof(double) * (S::QMD > 0 ? S::QMD : 1) * S::QN);
  std::memcpy(d.lbd, lbd, sizeof(d.lbd)); std::memcpy(d.ubd, ubd, sizeof(d.ubd));
  std::memcpy(d.lbu, lbu, sizeof(d.lbu)); std::memcpy(d.ubu, ubu, sizeof(d.ubu)); std::memcpy(d.s, s, sizeof(d.s));
}
// mode 0: cold Goldfarb-Idnani; 1: prediction, then (if not certified) the iteration from it -- what the kernels do
extern "C" int solve(const double* Ad, const double* lbd, const double* ubd, const double* lbu, const double* ubu,
                     const double* s, double* x, unsigned* au, unsigned* al, int mode, int* certified) {
  clik::QpSData<S> d;
  fill(d, Ad, lbd, ubd, lbu, ubu, s);
  double xs[S::QN];
  unsigned wu = 0, wl = 0;
  int st;
  *certified = 0;
  if (mode == 1 && clik::crash_guess<S>(d, &wu, &wl, xs)) {
    *certified = 1; st = 0; *au = wu; *al = wl;
  } else {
    st = clik::qp_structured<S>(d, xs, au, al, 400, wu, wl);
    if (st != 0 && (wu | wl) != 0) { fill(d, Ad, lbd, ubd, lbu, ubu, s); st = clik::qp_structured<S>(d, xs, au, al, 400, 0u, 0u); }
  }
  for (int j = 0; j < S::QN; ++j) x[j] = xs[j];
  return st;
}
'''

STRUCTURES = {
    # name: (nx, dense rows, [(unit row, column, coefficient)])
    "ur5_like_3_dense_two_bounds_per_joint": (9, [0, 1, 2], [(3 + i, i % 6, 1.0) for i in range(12)]),
    "six_dense_no_unit_rows": (9, [0, 1, 2, 3, 4, 5], []),
    "unit_rows_only_mixed_signs": (5, [], [(0, 0, 1.0), (1, 1, -2.0), (2, 2, 0.5), (3, 0, -1.0), (4, 4, 3.0)]),
    "four_dense_interleaved": (7, [1, 3, 4, 6], [(0, 0, 1.0), (2, 6, -1.0), (5, 3, 2.0), (7, 0, 1.0), (8, 5, 1.0)]),
}


@pytest.mark.parametrize("name", sorted(STRUCTURES))
def test_prediction_plus_iteration_equals_cold_solve_on_other_structures(name, tmp_path):
    nx, dense, unit = STRUCTURES[name]
    m = len(dense) + len(unit)
    ints = lambda v: ", ".join(str(int(i)) for i in (v or [0]))  # noqa: E731
    src = tmp_path / "s.cpp"
    src.write_text(SHIM_G % dict(hdr=os.path.join(ROOT, "casclik_b200", "csrc", "clik_qp.cuh"), qn=nx, md=len(dense),
                                 mu=len(unit), qm=m, md1=max(len(dense), 1), mu1=max(len(unit), 1), dense=ints(dense),
                                 urow=ints([u[0] for u in unit]), ucol=ints([u[1] for u in unit]),
                                 ucoef=", ".join(repr(float(u[2])) for u in unit) or "1.0"))
    so = tmp_path / "s.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))
    urows = [u[0] for u in unit]
    P = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.c_void_p)  # noqa: E731

    def run(h, A, lb, ub, mode):
        pad = lambda v: v if len(v) else np.zeros(1)  # noqa: E731
        arrs = [np.ascontiguousarray(A[dense]) if dense else np.zeros(nx), pad(lb[dense]), pad(ub[dense]),
                pad(lb[urows]), pad(ub[urows]), 1.0 / np.sqrt(h)]
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
        x = np.zeros(nx)
        au, al, cert = ctypes.c_uint(), ctypes.c_uint(), ctypes.c_int()
        st = lib.solve(*[P(a) for a in keep], P(x), ctypes.byref(au), ctypes.byref(al), mode, ctypes.byref(cert))
        return x, st, au.value, al.value, cert.value

    rng = np.random.default_rng(sum(map(ord, name)))
    n_cert = n_ok = 0
    for trial in range(250):
        A = np.zeros((m, nx))
        if dense:
            A[dense] = rng.normal(size=(len(dense), nx))
        for r, c, k in unit:
            A[r, c] = k
        h = rng.uniform(0.001, 2.0, nx)
        r0 = A @ rng.normal(size=nx)
        w = 0.05 if trial % 3 == 0 else 1.0
        lb, ub = r0 - rng.uniform(0, w, m), r0 + rng.uniform(0, w, m)
        for i in range(m):
            u = rng.random()
            if u < 0.08:
                lb[i] = -np.inf
            elif u < 0.16:
                ub[i] = np.inf
            elif u < 0.24:
                lb[i], ub[i] = -1e10, 1e10
            elif u < 0.40 and i in dense:
                lb[i] = ub[i] = r0[i]
        xo, lamo, sto = orc.solve_qp_single(h, A, lb, ub)
        x0, st0, au0, al0, _ = run(h, A, lb, ub, 0)
        x1, st1, au1, al1, cert = run(h, A, lb, ub, 1)
        assert st0 == st1 == sto, (trial, st0, st1, sto)
        if sto != 0:
            continue
        n_ok += 1
        n_cert += cert
        for x in (x0, x1):
            assert np.abs(x - xo).max() < 1e-8 * (1 + np.abs(xo).max()), (trial, cert)
            kk = orc.kkt_residuals(h, A, lb, ub, x)
            assert kk["primal"] < 1e-8 and kk["stationarity"] < 1e-8 and kk["sign"] < 1e-8, (trial, cert, kk)
        assert (au0, al0) == (au1, al1) == _masks(lamo), (trial, cert)
    assert n_ok > 150 and n_cert > 0.6 * n_ok, (n_ok, n_cert)
