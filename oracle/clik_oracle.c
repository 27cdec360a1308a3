/* clik_oracle.c — C restatement of the reference's CPU path for the pseudo-inverse controller
 * step.  TEST / BASELINE INFRASTRUCTURE ONLY: linked by tests/ and by bench.py's cpu_baseline and
 * `--impl reference` legs, never by the product (casclik_b200/).  PARITY UNPINNED in the same sense
 * as clik_oracle.py (no reference golden vectors exist); it is validated against clik_oracle.py
 * in tests/test_oracle_c.py.
 *
 * What it restates (paths under /root/reference/casclik/controllers/):
 *   pseudo_inverse.py:92-105    damped pseudo-inverse, formed EXPLICITLY as a matrix
 *                               (wide: solve(JJ' + lam I, J)' ; tall: solve(J'J + lam I, J'))
 *   pseudo_inverse.py:317-326   first EqualityConstraint: v += P(J) des
 *   pseudo_inverse.py:382-396   ... which then also runs the generic branch:
 *                               N = I - P(J) J ; v += (N P(J)) des        
 *   pseudo_inverse.py:512-556   solve(): one mode, accepted
 * for the skill of BASELINE.json configs[0]/[1]: e = p_fk(q) - y (3 rows), des = -K e.
 * The reference evaluates FK and its Jacobian through CasADi AD + generated C (gcc -O2,
 * pseudo_inverse.py:59-65); here they are hand-written (chain product + geometric Jacobian),
 * which is cheaper than AD code, so as a timing baseline this errs in the reference's favour.
 * Linear systems: Gaussian elimination with partial pivoting (what LAPACK dgesv, used by the
 * NumPy oracle, does).  Build: see oracle/Makefile (gcc -O2 -fopenmp, the reference's JIT flag).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 16 /* max state variables */
#define MAXM 16 /* max rows of one constraint */

typedef struct {
  int type;      /* 0 fixed, 1 revolute */
  double xyz[3]; /* origin translation */
  double R[9];   /* origin rotation, row-major */
  double axis[3];
} clik_joint;

static void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

static void rodrigues(const double* a, double th, double* R) {
  const double c = cos(th), s = sin(th), v = 1.0 - c;
  R[0] = c + a[0] * a[0] * v;        R[1] = a[0] * a[1] * v - a[2] * s; R[2] = a[0] * a[2] * v + a[1] * s;
  R[3] = a[1] * a[0] * v + a[2] * s; R[4] = c + a[1] * a[1] * v;        R[5] = a[1] * a[2] * v - a[0] * s;
  R[6] = a[2] * a[0] * v - a[1] * s; R[7] = a[2] * a[1] * v + a[0] * s; R[8] = c + a[2] * a[2] * v;
}

/* tip position p[3] and geometric position Jacobian J[3*n] (row-major) */
static void fk_position_jacobian(const clik_joint* chain, int njoints, const double* q, double* p,
                                 double* J, int n) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[9], Rq[9];
  double org[MAXN][3], ax[MAXN][3];
  p[0] = p[1] = p[2] = 0.0;
  int k = 0;
  for (int j = 0; j < njoints; ++j) {
    const clik_joint* jt = &chain[j];
    for (int i = 0; i < 3; ++i) p[i] += R[3 * i] * jt->xyz[0] + R[3 * i + 1] * jt->xyz[1] + R[3 * i + 2] * jt->xyz[2];
    mat3_mul(R, jt->R, T);
    memcpy(R, T, sizeof(T));
    if (jt->type == 1) {
      for (int i = 0; i < 3; ++i) {
        ax[k][i] = R[3 * i] * jt->axis[0] + R[3 * i + 1] * jt->axis[1] + R[3 * i + 2] * jt->axis[2];
        org[k][i] = p[i];
      }
      rodrigues(jt->axis, q[k], Rq);
      mat3_mul(R, Rq, T);
      memcpy(R, T, sizeof(T));
      ++k;
    }
  }
  for (int c = 0; c < n; ++c) {
    const double d0 = p[0] - org[c][0], d1 = p[1] - org[c][1], d2 = p[2] - org[c][2];
    J[0 * n + c] = ax[c][1] * d2 - ax[c][2] * d1;
    J[1 * n + c] = ax[c][2] * d0 - ax[c][0] * d2;
    J[2 * n + c] = ax[c][0] * d1 - ax[c][1] * d0;
  }
}

/* solve A X = B in place (A: k x k, B: k x nrhs, row-major), partial pivoting */
static void gauss_solve(double* A, double* B, int k, int nrhs) {
  for (int c = 0; c < k; ++c) {
    int piv = c;
    for (int r = c + 1; r < k; ++r)
      if (fabs(A[r * k + c]) > fabs(A[piv * k + c])) piv = r;
    if (piv != c) {
      for (int j = 0; j < k; ++j) { double t = A[c * k + j]; A[c * k + j] = A[piv * k + j]; A[piv * k + j] = t; }
      for (int j = 0; j < nrhs; ++j) { double t = B[c * nrhs + j]; B[c * nrhs + j] = B[piv * nrhs + j]; B[piv * nrhs + j] = t; }
    }
    for (int r = c + 1; r < k; ++r) {
      const double f = A[r * k + c] / A[c * k + c];
      for (int j = c + 1; j < k; ++j) A[r * k + j] -= f * A[c * k + j];
      for (int j = 0; j < nrhs; ++j) B[r * nrhs + j] -= f * B[c * nrhs + j];
    }
  }
  for (int j = 0; j < nrhs; ++j)
    for (int r = k - 1; r >= 0; --r) {
      double acc = B[r * nrhs + j];
      for (int c = r + 1; c < k; ++c) acc -= A[r * k + c] * B[c * nrhs + j];
      B[r * nrhs + j] = acc / A[r * k + r];
    }
}

/* P (n x m) = damped pseudo-inverse of J (m x n); pseudo_inverse.py:92-105 */
static void damped_pinv(const double* J, int m, int n, double lam, double* P) {
  double A[MAXN * MAXN], B[MAXN * MAXN];
  if (n >= m) {
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) {
        double acc = 0.0;
        for (int c = 0; c < n; ++c) acc += J[i * n + c] * J[j * n + c];
        A[i * m + j] = acc + (i == j ? lam : 0.0);
      }
    memcpy(B, J, sizeof(double) * m * n); /* solve(inner, J): m x n */
    gauss_solve(A, B, m, n);
    for (int i = 0; i < m; ++i)
      for (int c = 0; c < n; ++c) P[c * m + i] = B[i * n + c]; /* .T */
  } else {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int r = 0; r < m; ++r) acc += J[r * n + i] * J[r * n + j];
        A[i * n + j] = acc + (i == j ? lam : 0.0);
      }
    for (int i = 0; i < n; ++i)
      for (int r = 0; r < m; ++r) B[i * m + r] = J[r * n + i]; /* J.T: n x m */
    gauss_solve(A, B, n, m);
    memcpy(P, B, sizeof(double) * n * m);
  }
}

/* single first EqualityConstraint, literal: v = P des + (N P) des with N = I - P J */
static void pinv_first_equality(const double* J, const double* des, int m, int n, double lam, double* v) {
  double P[MAXN * MAXM], Nn[MAXN * MAXN], NJ[MAXN * MAXM];
  damped_pinv(J, m, n, lam, P);
  for (int i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int r = 0; r < m; ++r) acc += P[i * m + r] * des[r];
    v[i] = acc;
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double acc = 0.0;
      for (int r = 0; r < m; ++r) acc += P[i * m + r] * J[r * n + j];
      Nn[i * n + j] = (i == j ? 1.0 : 0.0) - acc;
    }
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < m; ++r) {
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc += Nn[i * n + j] * P[j * m + r];
      NJ[i * m + r] = acc;
    }
  for (int i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int r = 0; r < m; ++r) acc += NJ[i * m + r] * des[r];
    v[i] += acc;
  }
}

/* BASELINE configs[0]/[1]: e = p_fk(q) - y, K = gain, hard, one mode.
 * q: n x N, y: 3 x N, qdot: n x N, all coordinate-major (a[j*N + i]).  Returns threads used. */
int clik_ref_pinv_track(const clik_joint* chain, int njoints, int n, long N, const double* q,
                        const double* y, double gain, double lam, double* qdot, int nthreads) {
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = nthreads > 0 ? nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(static)
#endif
  for (long i = 0; i < N; ++i) {
    double qi[MAXN], p[3], J[3 * MAXN], des[3], v[MAXN];
    for (int j = 0; j < n; ++j) qi[j] = q[(size_t)j * N + i];
    fk_position_jacobian(chain, njoints, qi, p, J, n);
    for (int r = 0; r < 3; ++r) des[r] = -(gain * (p[r] - y[(size_t)r * N + i]));
    pinv_first_equality(J, des, 3, n, lam, v);
    for (int j = 0; j < n; ++j) qdot[(size_t)j * N + i] = v[j];
  }
  return used;
}

int clik_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
