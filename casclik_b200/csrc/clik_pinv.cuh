// clik_pinv.cuh — batched singularity-robust multiple-task-priority (SRMTP) pseudo-inverse step with
// set-based task activation, one controller instance per thread, fp64, sm_100a.
//
// Replaces, for N independent instances at once, what the reference does per instance with a
// Python loop over JIT-compiled CasADi functions:
//   PseudoInverseController.pinv                      reference casclik/controllers/pseudo_inverse.py:92-105
//   PseudoInverseController.get_problem_expressions   :259-451  (per-mode velocity)
//   PseudoInverseController.get_in_tangent_cone_function :132-190
//   PseudoInverseController.solve                     :512-556  (mode search, fallback zeros / -1)
//
// The skill-specific part (constraint values, Jacobians, desired task rates) is straight-line
// code emitted by casclik_b200/codegen as `S::eval`; this header is the hand-written algebra
// around it.  `S` also carries the compile-time skill description (sizes, constraint table,
// Jacobian sparsity, damping), so every loop below has constant bounds and fully unrolls: in the
// common path all matrices live in registers.
//
// Algebra.  The reference forms P(J) = J'(JJ'+lam I)^-1 (wide) or (J'J+lam I)^-1 J' (tall)
// explicitly and multiplies matrices; here P(J) is only ever applied to a vector (one SPD
// factorisation + one right-hand side), which is the same arithmetic up to association.  The
// reference's contract-relevant quirks are kept: the first
// EqualityConstraint is applied twice (A1), "first" means an empty active list (A2), the
// null-space projector is rebuilt from the stacked Jacobians with the damped inverse (A3), the
// in-tangent-cone test uses the 1e-12 thresholds of the code, not of the docstring (A4), active
// sets contribute no velocity (A5), VelocitySetConstraints are ignored (A6).
#pragma once
#include <cstdint>

namespace clik {

enum : int { KIND_EQ = 0, KIND_SET = 1, KIND_VELEQ = 2, KIND_VELSET = 3 };

// Transient value of mode[i] between clik_pinv_fast_kernel (static modes only) and the group pass of
// clik_pinv_group.cuh, which serves the instances every statically compiled mode rejected.
constexpr int PINV_PENDING = -2;

template <int A, int B> struct Max { static constexpr int v = A > B ? A : B; };

// Everything S::eval produces for one instance.  Unused members are never touched, so after
// scalar replacement they cost nothing.
template <class S> struct PinvData {
  double J[Max<S::M * S::NS, 1>::v];   // stacked constraint Jacobians, row-major, M x NS
  double des[Max<S::M, 1>::v];         // Eq: -K e - de/dt ; VelEq: target - de/dt
  double e[Max<S::M, 1>::v];           // Set rows: expression value
  double jt[Max<S::M, 1>::v];          // Set rows: de/dt
  double smin[Max<S::M, 1>::v];        // Set rows: bounds
  double smax[Max<S::M, 1>::v];
  double rmask[Max<S::M, 1>::v];       // multidim_sets only: 1 if the set row is outside its bounds
};

// ---- compile-time row lists ------------------------------------------------------------------
template <int... R> struct Rows {
  static constexpr int size = sizeof...(R);
  __host__ __device__ static constexpr int get(int i) {
    constexpr int a[sizeof...(R) + 1] = {R..., 0};
    return a[i];
  }
};
template <class A, class B> struct Concat;
template <int... X, int... Y> struct Concat<Rows<X...>, Rows<Y...>> { using type = Rows<X..., Y...>; };
template <int R0, int N, int... Acc> struct Range { using type = typename Range<R0, N - 1, R0 + N - 1, Acc...>::type; };
template <int R0, int... Acc> struct Range<R0, 0, Acc...> { using type = Rows<Acc...>; };

// ---- SPD solve in registers (Cholesky, packed lower triangle) -------------------------------------
// G: packed lower triangle, G[i*(i+1)/2 + j], j <= i.  b is overwritten with the solution.
template <int K> __device__ __forceinline__ void spd_solve(double (&G)[K * (K + 1) / 2], double (&b)[K]) {
  double r[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    double s = G[j * (j + 1) / 2 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) s = fma(-G[j * (j + 1) / 2 + k], G[j * (j + 1) / 2 + k], s);
    r[j] = rsqrt(s);
#pragma unroll
    for (int i = j + 1; i < K; ++i) {
      double t = G[i * (i + 1) / 2 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) t = fma(-G[i * (i + 1) / 2 + k], G[j * (j + 1) / 2 + k], t);
      G[i * (i + 1) / 2 + j] = t * r[j];
    }
  }
#pragma unroll
  for (int i = 0; i < K; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-G[i * (i + 1) / 2 + k], b[k], s);
    b[i] = s * r[i];
  }
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    double s = b[i];
#pragma unroll
    for (int k = i + 1; k < K; ++k) s = fma(-G[k * (k + 1) / 2 + i], b[k], s);
    b[i] = s * r[i];
  }
}

// ---- P(J_rows) * b for a compile-time row list -----------------------------------------------------
// wide  (cols >= rows): J' (J J' + lam I)^-1 b          reference pseudo_inverse.py:97-100
// tall  (cols <  rows): (J' J + lam I)^-1 J' b          reference pseudo_inverse.py:101-104
// "standard" (cs.pinv): same two branches, lam = 0, square matrices take the tall branch.
template <class S, class R>
__device__ __forceinline__ void pinv_times(const double (&J)[Max<S::M * S::NS, 1>::v],
                                           const double (&b)[Max<R::size, 1>::v], double (&out)[S::NS]) {
  constexpr int K = R::size;
  constexpr int NS = S::NS;
  constexpr bool wide = S::DAMPED ? (NS >= K) : (K < NS);
  constexpr double lam = S::DAMPED ? S::LAMBDA : 0.0;
  if constexpr (wide) {
    double G[K * (K + 1) / 2];
    double z[K];
#pragma unroll
    for (int a = 0; a < K; ++a) {
#pragma unroll
      for (int c = 0; c <= a; ++c) {
        double acc = 0.0;
        bool first = true;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if (S::jnz(R::get(a), j) && S::jnz(R::get(c), j)) {
            const double p = J[R::get(a) * NS + j], q = J[R::get(c) * NS + j];
            acc = first ? p * q : fma(p, q, acc);
            first = false;
          }
        }
        G[a * (a + 1) / 2 + c] = (a == c) ? acc + lam : acc;
      }
      z[a] = b[a];
    }
    spd_solve<K>(G, z);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int a = 0; a < K; ++a) {
        if (S::jnz(R::get(a), j)) {
          acc = first ? J[R::get(a) * NS + j] * z[a] : fma(J[R::get(a) * NS + j], z[a], acc);
          first = false;
        }
      }
      out[j] = acc;
    }
  } else {
    double G[NS * (NS + 1) / 2];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
#pragma unroll
      for (int c = 0; c <= i; ++c) {
        double acc = 0.0;
        bool first = true;
#pragma unroll
        for (int a = 0; a < K; ++a) {
          if (S::jnz(R::get(a), i) && S::jnz(R::get(a), c)) {
            const double p = J[R::get(a) * NS + i], q = J[R::get(a) * NS + c];
            acc = first ? p * q : fma(p, q, acc);
            first = false;
          }
        }
        G[i * (i + 1) / 2 + c] = (i == c) ? acc + lam : acc;
      }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int a = 0; a < K; ++a) {
        if (S::jnz(R::get(a), j)) {
          acc = first ? J[R::get(a) * NS + j] * b[a] : fma(J[R::get(a) * NS + j], b[a], acc);
          first = false;
        }
      }
      out[j] = acc;
    }
    spd_solve<NS>(G, out);
  }
}

// x <- (I - P(J_rows) J_rows) x        reference pseudo_inverse.py:387-393 (rJ = J without multidim sets)
// With options["multidim_sets"] the reduced stack rJ = S J (S = diag(row is outside its bounds),
// pseudo_inverse.py:289-298, :401-402) replaces J inside the product only: b = rmask .* (J x).
template <class S, class R>
__device__ __forceinline__ void nullspace_apply(const double (&J)[Max<S::M * S::NS, 1>::v],
                                                const double (&rmask)[Max<S::M, 1>::v], double (&x)[S::NS]) {
  constexpr int K = R::size;
  constexpr int NS = S::NS;
  double b[Max<K, 1>::v];
#pragma unroll
  for (int a = 0; a < K; ++a) {
    double acc = 0.0;
    bool first = true;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      if (S::jnz(R::get(a), j)) {
        acc = first ? J[R::get(a) * NS + j] * x[j] : fma(J[R::get(a) * NS + j], x[j], acc);
        first = false;
      }
    }
    if constexpr (S::MULTIDIM) {
      b[a] = (S::row_is_set(R::get(a))) ? acc * rmask[R::get(a)] : acc;
    } else {
      b[a] = acc;
    }
  }
  double corr[NS];
  pinv_times<S, R>(J, b, corr);
#pragma unroll
  for (int j = 0; j < NS; ++j) x[j] -= corr[j];
}

// First EqualityConstraint, both applications at once (reference pseudo_inverse.py:317-326 followed
// by :382-396 with J0toi = rJ0toi = Ji):  v = P des + (I - P J) P des.  With A = J J' + lam I
// (wide) the second term is exactly lam J' A^-2 des, so
//     v = J' A^-1 (des + lam A^-1 des)          wide
//     v = w + lam B^-1 w,  w = B^-1 J' des,  B = J' J + lam I      tall
// i.e. one factorisation, two triangular solves, no J w / J' z round trip and no cancellation
// (the literal form subtracts two nearly equal vectors to obtain a term of relative size lam/sigma^2).
template <class S, class R>
__device__ __forceinline__ void first_equality_twice(const double (&J)[Max<S::M * S::NS, 1>::v],
                                                     const double (&b)[Max<R::size, 1>::v],
                                                     double (&out)[S::NS]) {
  constexpr int K = R::size;
  constexpr int NS = S::NS;
  constexpr bool wide = S::DAMPED ? (NS >= K) : (K < NS);
  constexpr double lam = S::DAMPED ? S::LAMBDA : 0.0;
  if constexpr (wide) {
    double G[K * (K + 1) / 2], G2[K * (K + 1) / 2];
    double z[K], y[K];
#pragma unroll
    for (int a = 0; a < K; ++a) {
#pragma unroll
      for (int c = 0; c <= a; ++c) {
        double acc = 0.0;
        bool first = true;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if (S::jnz(R::get(a), j) && S::jnz(R::get(c), j)) {
            const double p = J[R::get(a) * NS + j], q = J[R::get(c) * NS + j];
            acc = first ? p * q : fma(p, q, acc);
            first = false;
          }
        }
        G[a * (a + 1) / 2 + c] = (a == c) ? acc + lam : acc;
        G2[a * (a + 1) / 2 + c] = G[a * (a + 1) / 2 + c];
      }
      z[a] = b[a];
    }
    spd_solve<K>(G, z);                       // z = A^-1 des
#pragma unroll
    for (int a = 0; a < K; ++a) y[a] = fma(lam, z[a], b[a]);
    spd_solve<K>(G2, y);                      // same factorisation (merged by the compiler)
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int a = 0; a < K; ++a) {
        if (S::jnz(R::get(a), j)) {
          acc = first ? J[R::get(a) * NS + j] * y[a] : fma(J[R::get(a) * NS + j], y[a], acc);
          first = false;
        }
      }
      out[j] = acc;
    }
  } else {
    double w[NS];
    pinv_times<S, R>(J, b, w);
    double G[NS * (NS + 1) / 2];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
#pragma unroll
      for (int c = 0; c <= i; ++c) {
        double acc = 0.0;
        bool first = true;
#pragma unroll
        for (int a = 0; a < K; ++a) {
          if (S::jnz(R::get(a), i) && S::jnz(R::get(a), c)) {
            const double p = J[R::get(a) * NS + i], q = J[R::get(a) * NS + c];
            acc = first ? p * q : fma(p, q, acc);
            first = false;
          }
        }
        G[i * (i + 1) / 2 + c] = (i == c) ? acc + lam : acc;
      }
    }
    double y[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) y[j] = w[j];
    spd_solve<NS>(G, y);
#pragma unroll
    for (int j = 0; j < NS; ++j) out[j] = fma(lam, y[j], w[j]);
  }
}

// P(J_c) des_c of every Eq / VelEq constraint.  It does not depend on the mode (only the null-space
// projector does), so skills with SetConstraints compute it once per instance and every mode that
// is tried reuses it.
template <class S> struct TaskVel {
  double w[Max<S::NEQC, 1>::v][S::NS];
};

template <class S, int C = 0>
__device__ __forceinline__ void compute_task_vel(const PinvData<S>& d, TaskVel<S>& tw) {
  if constexpr (C < S::NC) {
    constexpr int kind = S::kind(C);
    if constexpr (kind == KIND_EQ || kind == KIND_VELEQ ||
                  (kind == KIND_SET && S::CONV_LAST && C == S::NC - 1)) {
      constexpr int r0 = S::row0(C);
      constexpr int m = S::rows(C);
      using RC = typename Range<r0, m>::type;
      double b[Max<m, 1>::v];
#pragma unroll
      for (int a = 0; a < m; ++a) b[a] = d.des[r0 + a];
      pinv_times<S, RC>(d.J, b, tw.w[S::eq_index(C)]);
    }
    compute_task_vel<S, C + 1>(d, tw);
  }
}

// ---- one mode, mode mask known at compile time -----------------------------------------------------
// Walks the priority-sorted constraint table exactly like the reference's loop
// (pseudo_inverse.py:274-443), with the active-row list carried as a type.
template <class S, unsigned MASK, int C, class Stack, bool PRE> struct StaticMode {
  __device__ __forceinline__ static void run(const PinvData<S>& d, const TaskVel<S>& tw, double (&v)[S::NS]) {
    if constexpr (C < S::NC) {
      constexpr int kind = S::kind(C);
      constexpr int r0 = S::row0(C);
      constexpr int m = S::rows(C);
      using RC = typename Range<r0, m>::type;
      if constexpr (kind == KIND_EQ || kind == KIND_VELEQ) {
        double b[Max<m, 1>::v];
#pragma unroll
        for (int a = 0; a < m; ++a) b[a] = d.des[r0 + a];
        double w[S::NS];
        if constexpr (Stack::size == 0) {
          using S1 = typename Concat<Stack, RC>::type;
          if constexpr (kind == KIND_EQ) {
            // :322-326 and, because the reference's next chain starts with a new `if`, :382-396
            // (an empty stack means nothing has contributed yet: v is still zero, so the first
            // contribution is assigned — "0 + w" would cost one fp64 add per coordinate)
            if constexpr (S::FUSE_FIRST_EQ && !PRE) {
              first_equality_twice<S, RC>(d.J, b, w);
#pragma unroll
              for (int j = 0; j < S::NS; ++j) v[j] = w[j];
            } else {
              if constexpr (PRE) {
#pragma unroll
                for (int j = 0; j < S::NS; ++j) w[j] = tw.w[S::eq_index(C)][j];
              } else {
                pinv_times<S, RC>(d.J, b, w);
              }
#pragma unroll
              for (int j = 0; j < S::NS; ++j) v[j] = w[j];
              nullspace_apply<S, S1>(d.J, d.rmask, w);
#pragma unroll
              for (int j = 0; j < S::NS; ++j) v[j] += w[j];
            }
            StaticMode<S, MASK, C + 1, typename Concat<S1, RC>::type, PRE>::run(d, tw, v);
          } else {
            if constexpr (PRE) {
#pragma unroll
              for (int j = 0; j < S::NS; ++j) w[j] = tw.w[S::eq_index(C)][j];
            } else {
              pinv_times<S, RC>(d.J, b, w);                              // :331-335
            }
#pragma unroll
            for (int j = 0; j < S::NS; ++j) v[j] = w[j];
            StaticMode<S, MASK, C + 1, S1, PRE>::run(d, tw, v);
          }
        } else {
          if constexpr (PRE) {
#pragma unroll
            for (int j = 0; j < S::NS; ++j) w[j] = tw.w[S::eq_index(C)][j];
          } else {
            pinv_times<S, RC>(d.J, b, w);
          }
          nullspace_apply<S, Stack>(d.J, d.rmask, w);                     // :387-394 / :434-441
#pragma unroll
          for (int j = 0; j < S::NS; ++j) v[j] += w[j];
          StaticMode<S, MASK, C + 1, typename Concat<Stack, RC>::type, PRE>::run(d, tw, v);
        }
      } else if constexpr (kind == KIND_SET) {
        if constexpr ((MASK >> S::set_index(C)) & 1u) {                   // :399-405
          if constexpr (S::CONV_LAST && C == S::NC - 1 && Stack::size > 0) {
            // converge_final_set_to_max (:337-356): the active final set is also driven to
            // set_max, des = K (max - e) - de/dt, through the null space of everything above it
            double b[Max<m, 1>::v];
#pragma unroll
            for (int a = 0; a < m; ++a) b[a] = d.des[r0 + a];
            double w[S::NS];
            if constexpr (PRE) {
#pragma unroll
              for (int j = 0; j < S::NS; ++j) w[j] = tw.w[S::eq_index(C)][j];
            } else {
              pinv_times<S, RC>(d.J, b, w);
            }
            nullspace_apply<S, Stack>(d.J, d.rmask, w);
#pragma unroll
            for (int j = 0; j < S::NS; ++j) v[j] += w[j];
          }
          StaticMode<S, MASK, C + 1, typename Concat<Stack, RC>::type, PRE>::run(d, tw, v);
        } else {
          StaticMode<S, MASK, C + 1, Stack, PRE>::run(d, tw, v);
        }
      } else {
        StaticMode<S, MASK, C + 1, Stack, PRE>::run(d, tw, v);            // VelocitySet: ignored
      }
    }
  }
};

// in-tangent-cone test of one scalar set (reference pseudo_inverse.py:151-185)
__device__ __forceinline__ bool in_tangent_cone(double e, double de, double smin, double smax) {
  return (smin - e < 1e-12) ? ((e - smax < 1e-12) ? true : (de < 0.0)) : (de > 0.0);
}

// Vector-valued set (options["multidim_sets"], reference pseudo_inverse.py:222-252): inside <=> every
// e - min >= 1e-12 and every e - max <= 1e-12; otherwise the motion must point inwards, with a
// 45-degree rule when every component is outside.  de = de/dt + J v.
template <int MR>
__device__ __forceinline__ bool in_tangent_cone_multidim(const double* e, const double* de, const double* smin,
                                                         const double* smax) {
  bool above = true, below = true, corner = true;
  double proj = 0.0, dd = 0.0, oo = 0.0;
#pragma unroll
  for (int r = 0; r < MR; ++r) {
    const double le = e[r] - smin[r], ue = e[r] - smax[r];
    above = above && (le >= 1e-12);
    below = below && (ue <= 1e-12);
    const double sl = (double)((le > 0.0) - (le < 0.0)), su = (double)((ue > 0.0) - (ue < 0.0));
    corner = corner && (sl == su);
    const double od = (sl + su) / 2.0;
    proj += od * de[r];
    dd += de[r] * de[r];
    oo += od * od;
  }
  if (above && below) return true;
  if (corner) {
    if (!(proj < 0.0)) return false;
    const double dists = (sqrt(dd) + 1e-10) * sqrt(oo);
    return fabs(-proj) / dists < 0.70710678118654757;   // cos(pi/4)
  }
  return proj < 0.0;
}

// in-tangent-cone test of set constraint C for a candidate velocity v (1 row or vector-valued)
template <class S, int C>
__device__ __forceinline__ bool set_admissible(const PinvData<S>& d, const double (&v)[S::NS]) {
  constexpr int r0 = S::row0(C);
  constexpr int m = S::rows(C);
  double de[m];
#pragma unroll
  for (int a = 0; a < m; ++a) {
    double dot = 0.0;
    bool first = true;
#pragma unroll
    for (int j = 0; j < S::NS; ++j) {
      if (S::jnz(r0 + a, j)) {
        dot = first ? d.J[(r0 + a) * S::NS + j] * v[j] : fma(d.J[(r0 + a) * S::NS + j], v[j], dot);
        first = false;
      }
    }
    de[a] = d.jt[r0 + a] + dot;
  }
  if constexpr (m == 1) {
    return in_tangent_cone(d.e[r0], de[0], d.smin[r0], d.smax[r0]);
  } else {
    return in_tangent_cone_multidim<m>(&d.e[r0], de, &d.smin[r0], &d.smax[r0]);
  }
}

template <class S, unsigned MASK, int C = 0>
__device__ __forceinline__ bool inactive_sets_admissible(const PinvData<S>& d, const double (&v)[S::NS]) {
  if constexpr (C < S::NC) {
    bool ok = true;
    if constexpr (S::kind(C) == KIND_SET) {
      if constexpr (!((MASK >> S::set_index(C)) & 1u)) ok = set_admissible<S, C>(d, v);
    }
    return ok && inactive_sets_admissible<S, MASK, C + 1>(d, v);
  } else {
    return true;
  }
}

template <class S, unsigned MASK, bool PRE>
__device__ __forceinline__ bool static_mode(const PinvData<S>& d, const TaskVel<S>& tw, double (&v)[S::NS]) {
#pragma unroll
  for (int j = 0; j < S::NS; ++j) v[j] = 0.0;
  StaticMode<S, MASK, 0, Rows<>, PRE>::run(d, tw, v);
  return inactive_sets_admissible<S, MASK>(d, v);
}

// ---- one mode, mode mask known only at run time (the rare path) -------------------------------------
// Same algebra with run-time loops over a row list held in local memory.  Only instances that
// are rejected by every statically instantiated mode come here.
template <class S>
__device__ void dyn_pinv_times(const double* J, const int* rows, int K, const double* b, double* out) {
  constexpr int NS = S::NS;
  constexpr int D = NS;   // wide: K <= NS ; tall: NS
  double G[D * (D + 1) / 2];
  double z[D];
  double r[D];
  const bool wide = S::DAMPED ? (NS >= K) : (K < NS);
  const double lam = S::DAMPED ? S::LAMBDA : 0.0;
  int n;
  if (wide) {
    n = K;
    for (int a = 0; a < K; ++a) {
      for (int c = 0; c <= a; ++c) {
        double acc = 0.0;
        for (int j = 0; j < NS; ++j) acc = fma(J[rows[a] * NS + j], J[rows[c] * NS + j], acc);
        G[a * (a + 1) / 2 + c] = (a == c) ? acc + lam : acc;
      }
      z[a] = b[a];
    }
  } else {
    n = NS;
    for (int i = 0; i < NS; ++i) {
      for (int c = 0; c <= i; ++c) {
        double acc = 0.0;
        for (int a = 0; a < K; ++a) acc = fma(J[rows[a] * NS + i], J[rows[a] * NS + c], acc);
        G[i * (i + 1) / 2 + c] = (i == c) ? acc + lam : acc;
      }
      double acc = 0.0;
      for (int a = 0; a < K; ++a) acc = fma(J[rows[a] * NS + i], b[a], acc);
      z[i] = acc;
    }
  }
  for (int j = 0; j < n; ++j) {
    double s = G[j * (j + 1) / 2 + j];
    for (int k = 0; k < j; ++k) s = fma(-G[j * (j + 1) / 2 + k], G[j * (j + 1) / 2 + k], s);
    r[j] = rsqrt(s);
    for (int i = j + 1; i < n; ++i) {
      double t = G[i * (i + 1) / 2 + j];
      for (int k = 0; k < j; ++k) t = fma(-G[i * (i + 1) / 2 + k], G[j * (j + 1) / 2 + k], t);
      G[i * (i + 1) / 2 + j] = t * r[j];
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = z[i];
    for (int k = 0; k < i; ++k) s = fma(-G[i * (i + 1) / 2 + k], z[k], s);
    z[i] = s * r[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < n; ++k) s = fma(-G[k * (k + 1) / 2 + i], z[k], s);
    z[i] = s * r[i];
  }
  if (wide) {
    for (int j = 0; j < NS; ++j) {
      double acc = 0.0;
      for (int a = 0; a < K; ++a) acc = fma(J[rows[a] * NS + j], z[a], acc);
      out[j] = acc;
    }
  } else {
    for (int j = 0; j < NS; ++j) out[j] = z[j];
  }
}

template <class S>
__device__ __noinline__ bool dynamic_mode(const PinvData<S>* d, const TaskVel<S>* tw, unsigned mask, double* v) {
  constexpr int NS = S::NS;
  constexpr int MAXK = S::M + S::MAXROWS;   // every row once + the doubled first equality
  int stack[MAXK];
  int k = 0;
  double w[NS], corr[NS], b[Max<MAXK, 1>::v];
  for (int j = 0; j < NS; ++j) v[j] = 0.0;
  for (int c = 0; c < S::NC; ++c) {
    const int kind = S::kind(c), r0 = S::row0(c), m = S::rows(c);
    if (kind == KIND_EQ || kind == KIND_VELEQ) {
      for (int j = 0; j < NS; ++j) w[j] = tw->w[S::eq_index(c)][j];     // mode-independent, precomputed
      const bool first = (k == 0);
      if (first) {
        for (int j = 0; j < NS; ++j) v[j] += w[j];
        for (int a = 0; a < m; ++a) stack[k++] = r0 + a;
      }
      if (!first || kind == KIND_EQ) {
        for (int a = 0; a < k; ++a) {
          double acc = 0.0;
          for (int j = 0; j < NS; ++j) acc = fma(d->J[stack[a] * NS + j], w[j], acc);
          b[a] = (S::MULTIDIM && S::row_is_set(stack[a])) ? acc * d->rmask[stack[a]] : acc;
        }
        dyn_pinv_times<S>(d->J, stack, k, b, corr);
        for (int j = 0; j < NS; ++j) v[j] += w[j] - corr[j];
        for (int a = 0; a < m; ++a) stack[k++] = r0 + a;
      }
    } else if (kind == KIND_SET) {
      if ((mask >> S::set_index(c)) & 1u) {
        if (S::CONV_LAST && c == S::NC - 1 && k > 0) {         // converge_final_set_to_max, :337-356
          for (int j = 0; j < NS; ++j) w[j] = tw->w[S::eq_index(c)][j];
          for (int a = 0; a < k; ++a) {
            double acc = 0.0;
            for (int j = 0; j < NS; ++j) acc = fma(d->J[stack[a] * NS + j], w[j], acc);
            b[a] = (S::MULTIDIM && S::row_is_set(stack[a])) ? acc * d->rmask[stack[a]] : acc;
          }
          dyn_pinv_times<S>(d->J, stack, k, b, corr);
          for (int j = 0; j < NS; ++j) v[j] += w[j] - corr[j];
        }
        for (int a = 0; a < m; ++a) stack[k++] = r0 + a;
      }
    }
  }
  bool ok = true;
  for (int c = 0; c < S::NC; ++c) {
    if (S::kind(c) == KIND_SET && !((mask >> S::set_index(c)) & 1u)) {
      const int r0 = S::row0(c), m = S::rows(c);
      double de[S::MAXROWS];
      for (int a = 0; a < m; ++a) {
        double dot = 0.0;
        for (int j = 0; j < NS; ++j) dot = fma(d->J[(r0 + a) * NS + j], v[j], dot);
        de[a] = d->jt[r0 + a] + dot;
      }
      if (m == 1) {
        ok = ok && in_tangent_cone(d->e[r0], de[0], d->smin[r0], d->smax[r0]);
      } else {
        // run-time row count: same rule as in_tangent_cone_multidim
        bool above = true, below = true, corner = true;
        double proj = 0.0, dd = 0.0, oo = 0.0;
        for (int a = 0; a < m; ++a) {
          const double le = d->e[r0 + a] - d->smin[r0 + a], ue = d->e[r0 + a] - d->smax[r0 + a];
          above = above && (le >= 1e-12);
          below = below && (ue <= 1e-12);
          const double sl = (double)((le > 0.0) - (le < 0.0)), su = (double)((ue > 0.0) - (ue < 0.0));
          corner = corner && (sl == su);
          const double od = (sl + su) / 2.0;
          proj += od * de[a];
          dd += de[a] * de[a];
          oo += od * od;
        }
        bool in_tc;
        if (above && below) in_tc = true;
        else if (corner) in_tc = (proj < 0.0) && (fabs(-proj) / ((sqrt(dd) + 1e-10) * sqrt(oo)) < 0.70710678118654757);
        else in_tc = proj < 0.0;
        ok = ok && in_tc;
      }
    }
  }
  return ok;
}

// ---- the step --------------------------------------------------------------------------------------
// One instance, inputs already in registers: evaluate the skill, try mode 0 on the static path,
// fall back to the run-time mode search, and return the accepted mode index (-1: none).
// Mode with run-time mask for skills whose SetConstraints each bound one state coordinate with a
// constant coefficient (S::UNIT_SETS: joint limits, e = c*q_k) and whose single Eq / VelEq task has
// the lowest priority.  With at least one active set the task is not "first", so
// v = (I - S'(S S' + lam I)^-1 S) w with S S' = diag(c_k^2): the projector only rescales the active
// coordinates.  Same operations, in the same order, as the generic path performs on such rows
// (b = c w_k; Cholesky of the 1x1 block c^2 + lam via rsqrt; two substitutions; w_k - c z).
template <class S>
__device__ __forceinline__ bool unit_set_mode(const PinvData<S>& d, const TaskVel<S>& tw, unsigned mask,
                                              double (&v)[S::NS]) {
  constexpr double lam = S::DAMPED ? S::LAMBDA : 0.0;
#pragma unroll
  for (int j = 0; j < S::NS; ++j) v[j] = tw.w[0][j];
#pragma unroll
  for (int c = 0; c < S::NC; ++c) {
    if (S::kind(c) == KIND_SET) {
      const int k = S::set_index(c);
      const int col = S::set_unit_col(k);
      const double coef = S::set_unit_coef(k);
      if ((mask >> k) & 1u) {
        const double r = rsqrt(coef * coef + lam);
        const double zk = ((coef * v[col]) * r) * r;
        v[col] = v[col] - coef * zk;
      }
    }
  }
  bool ok = true;
#pragma unroll
  for (int c = 0; c < S::NC; ++c) {
    if (S::kind(c) == KIND_SET) {
      const int k = S::set_index(c);
      if (!((mask >> k) & 1u)) {
        const int r = S::row0(c);
        const double de = d.jt[r] + S::set_unit_coef(k) * v[S::set_unit_col(k)];
        ok = ok && in_tangent_cone(d.e[r], de, d.smin[r], d.smax[r]);
      }
    }
  }
  return ok;
}

// Run-time mode index -> statically instantiated mode (the first S::NSTATIC entries of the
// activation map: mode 0 and the single-set modes, or every mode when there are at most 8).
template <class S, int MI> struct StaticDispatch {
  __device__ __forceinline__ static bool run(int mi, const PinvData<S>& d, const TaskVel<S>& tw,
                                             double (&v)[S::NS]) {
    if (mi == MI) return static_mode<S, S::static_mask(MI), (S::NSTATIC < S::NMODES)>(d, tw, v);
    if constexpr (MI + 1 < S::NSTATIC) {
      return StaticDispatch<S, MI + 1>::run(mi, d, tw, v);
    } else {
      return false;
    }
  }
};

// One instance, inputs already in registers: evaluate the skill once, then walk the activation map
// in the reference's order (pseudo_inverse.py:530-550) until a mode passes its in-tangent-cone
// tests.  Mode 0 and the other frequent modes run on the static register path; the rest use the
// run-time path on a local-memory copy of the constraint data.  Returns the accepted mode index
// (-1: none admissible, v = 0, pseudo_inverse.py:551-555).
template <class S, bool SPLIT = false>
__device__ __forceinline__ int solve_instance(const double tv, const double (&qv)[Max<S::NQ, 1>::v],
                                              const double (&xv)[Max<S::NX, 1>::v],
                                              const double (&yv)[Max<S::NY, 1>::v], double (&v)[S::NS]) {
  PinvData<S> d;
  S::eval(tv, qv, xv, yv, d);
  TaskVel<S> tw;
  if constexpr (S::NSETS > 0) {
    // precompute the mode-independent task velocities only when many modes may be tried
    constexpr bool PRE = S::NSTATIC < S::NMODES;
    if constexpr (PRE) compute_task_vel<S>(d, tw);
    if (static_mode<S, 0u, PRE>(d, tw, v)) return 0;
    int accepted = -1;
    if constexpr (S::UNIT_SETS) {
      // Unit sets + task last: for any mode with at least one active set the task is not "first",
      // v = N(S_M) w, and N only rescales the ACTIVE coordinates.  So whether an inactive set k
      // passes its in-tangent-cone test depends on w alone, not on the rest of the mode:
      // mode M is admissible <=> M contains F = {k : set k fails the test under w}.  The
      // activation map is ordered by number of active sets, so the first admissible non-empty
      // mode of the reference's sequential search (pseudo_inverse.py:530-550) is F itself
      // (or the first single-set mode if F is empty: mode 0 failed under its own, doubled,
      // velocity but every set passes under w).  No search loop.
      unsigned F = 0u;
#pragma unroll
      for (int c = 0; c < S::NC; ++c) {
        if (S::kind(c) == KIND_SET) {
          const int k = S::set_index(c);
          const int r = S::row0(c);
          const double de = d.jt[r] + S::set_unit_coef(k) * tw.w[0][S::set_unit_col(k)];
          if (!in_tangent_cone(d.e[r], de, d.smin[r], d.smax[r])) F |= 1u << k;
        }
      }
      accepted = (F == 0u) ? 1 : (int)S::mode_index(F);
      unit_set_mode<S>(d, tw, S::mode_mask(accepted), v);
      return accepted;
    }
    for (int mi = 1; mi < S::NSTATIC; ++mi) {
      if (StaticDispatch<S, 1 < S::NSTATIC ? 1 : 0>::run(mi, d, tw, v)) {
        accepted = mi;
        break;
      }
    }
    if constexpr (S::NSTATIC < S::NMODES && SPLIT) {
      // two-launch form: the run-time tail of the activation map is the group pass's job
      if (accepted < 0) return PINV_PENDING;
    } else if constexpr (S::NSTATIC < S::NMODES) {
      if (accepted < 0) {
        // the run-time path indexes dynamically: work on copies, keep `d` / `tw` in registers
        PinvData<S> copy = d;
        TaskVel<S> twc = tw;
        for (int mi = S::NSTATIC; mi < S::NMODES; ++mi) {
          if (dynamic_mode<S>(&copy, &twc, S::mode_mask(mi), v)) {
            accepted = mi;
            break;
          }
        }
      }
    }
    if (accepted < 0) {
#pragma unroll
      for (int j = 0; j < S::NS; ++j) v[j] = 0.0;
    }
    return accepted;
  } else {
    // no SetConstraints: a single mode with no test to fail
    static_mode<S, 0u, false>(d, tw, v);
    return 0;
  }
}

template <class S>
__device__ __forceinline__ void store_instance(long long ld, long long i, const double (&v)[S::NS],
                                               int accepted, double* __restrict__ qdot,
                                               double* __restrict__ xdot, int* __restrict__ mode) {
#pragma unroll
  for (int j = 0; j < S::NQ; ++j) __stcs(qdot + (long long)j * ld + i, v[j]);
#pragma unroll
  for (int j = 0; j < S::NX; ++j) __stcs(xdot + (long long)j * ld + i, v[S::NQ + j]);
  if (mode != nullptr) __stcs(mode + i, accepted);
}

// Plain driver.  Structure-of-arrays batch: q[j*ld + i] is coordinate j of instance i (same for
// x, y, outputs), so consecutive threads touch consecutive addresses; N instances are processed and
// ld >= N is the row stride (ld = N for a whole batch; ld = size of the whole batch when the call
// covers one shard [lo, lo + N) of it, with every pointer advanced by lo: how one host batch is split
// over several GPUs without repacking).  t has stride t_stride
// (0 = one shared time).  mode[i] = index into the activation map of the accepted mode, -1 (and
// zero velocity) if none.  Used when the TMA driver's alignment conditions do not hold.
template <class S>
__device__ __forceinline__ void load_instance(long long ld, long long i, const double* __restrict__ t,
                                              int t_stride, const double* __restrict__ q,
                                              const double* __restrict__ x, const double* __restrict__ y,
                                              double& tv, double (&qv)[Max<S::NQ, 1>::v],
                                              double (&xv)[Max<S::NX, 1>::v], double (&yv)[Max<S::NY, 1>::v]) {
  tv = __ldcs(t + (long long)t_stride * i);
#pragma unroll
  for (int j = 0; j < S::NQ; ++j) qv[j] = __ldcs(q + (long long)j * ld + i);
#pragma unroll
  for (int j = 0; j < S::NX; ++j) xv[j] = __ldcs(x + (long long)j * ld + i);
#pragma unroll
  for (int j = 0; j < S::NY; ++j) yv[j] = __ldcs(y + (long long)j * ld + i);
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// UNROLL instances per thread: the loads of all of them are issued before the first solve, so
// the memory latency of instance k+1 hides behind the arithmetic of instance k.
// PF > 0: every thread also asks L2 for the inputs of the instance PF CTAs ahead (about one wave
// of resident CTAs), so that the CTA scheduled there later finds its inputs in L2, not in DRAM.
template <class S, int UNROLL, int PF = 0, bool SPLIT = false>
__device__ __forceinline__ void pinv_step(long long N, long long ld, const double* __restrict__ t, int t_stride,
                                          const double* __restrict__ q, const double* __restrict__ x,
                                          const double* __restrict__ y, double* __restrict__ qdot,
                                          double* __restrict__ xdot, int* __restrict__ mode) {
  const long long stride = (long long)gridDim.x * blockDim.x * UNROLL;
  for (long long base = (long long)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; base < N; base += stride) {
    if constexpr (PF > 0) {
      const long long ip = base + (long long)PF * blockDim.x * UNROLL;
      if (ip < N) {
#pragma unroll
        for (int k = 0; k < S::NIN; ++k) prefetch_l2(S::in_row(k, ld, t, q, x, y) + ((k == 0 && S::T_STAGED && t_stride == 0) ? 0 : ip));
      }
    }
    double tv[UNROLL], qv[UNROLL][Max<S::NQ, 1>::v], xv[UNROLL][Max<S::NX, 1>::v], yv[UNROLL][Max<S::NY, 1>::v];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = base + (long long)u * blockDim.x;
      if (i < N) load_instance<S>(ld, i, t, t_stride, q, x, y, tv[u], qv[u], xv[u], yv[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = base + (long long)u * blockDim.x;
      if (i < N) {
        double v[S::NS];
        const int accepted = solve_instance<S, SPLIT>(tv[u], qv[u], xv[u], yv[u], v);
        if (SPLIT && accepted == PINV_PENDING) {
          __stcs(mode + i, accepted);             // (the split form is only launched with a mode array)
        } else {
          store_instance<S>(ld, i, v, accepted, qdot, xdot, mode);
        }
      }
    }
  }
}

// ---- closed-loop rollout -------------------------------------------------------------------------------
// The loop every CASCLIK user runs around solve() (examples/notebooks/ur5_moe2016_example2.ipynb cell 12,
// lines :535-549): res = solve(t_k, q_k); dq = clip(res, +-max_speed); q_{k+1} = q_k + dq*dt, with
// t_k = t0 + k*dt, here for `steps` steps per instance without leaving the GPU.  The state update
// is rounded like the notebook's NumPy code (separate multiply and add, no fma contraction).
template <class S>
__device__ __forceinline__ void pinv_rollout(long long N, long long ld, int steps, double dt, const double* __restrict__ t0,
                                             int t_stride, double* __restrict__ q, double* __restrict__ x,
                                             const double* __restrict__ y, double vmax_q, double vmax_x,
                                             double* __restrict__ qdot_last, double* __restrict__ xdot_last,
                                             int* __restrict__ mode_last, int* __restrict__ n_failed) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    double t0v, qv[Max<S::NQ, 1>::v], xv[Max<S::NX, 1>::v], yv[Max<S::NY, 1>::v];
    load_instance<S>(ld, i, t0, t_stride, q, x, y, t0v, qv, xv, yv);
    double v[S::NS];
#pragma unroll
    for (int j = 0; j < S::NS; ++j) v[j] = 0.0;
    int accepted = 0, failed = 0;
    for (int k = 0; k < steps; ++k) {
      const double tv = __dadd_rn(t0v, __dmul_rn(dt, (double)k));
      accepted = solve_instance<S>(tv, qv, xv, yv, v);
      failed += (accepted < 0) ? 1 : 0;
#pragma unroll
      for (int j = 0; j < S::NQ; ++j) {
        v[j] = fmax(fmin(v[j], vmax_q), -vmax_q);
        qv[j] = __dadd_rn(qv[j], __dmul_rn(v[j], dt));
      }
#pragma unroll
      for (int j = 0; j < S::NX; ++j) {
        v[S::NQ + j] = fmax(fmin(v[S::NQ + j], vmax_x), -vmax_x);
        xv[j] = __dadd_rn(xv[j], __dmul_rn(v[S::NQ + j], dt));
      }
    }
#pragma unroll
    for (int j = 0; j < S::NQ; ++j) {
      q[(long long)j * ld + i] = qv[j];
      if (qdot_last != nullptr) qdot_last[(long long)j * ld + i] = v[j];
    }
#pragma unroll
    for (int j = 0; j < S::NX; ++j) {
      x[(long long)j * ld + i] = xv[j];
      if (xdot_last != nullptr) xdot_last[(long long)j * ld + i] = v[S::NQ + j];
    }
    if (mode_last != nullptr) mode_last[i] = accepted;
    if (n_failed != nullptr) n_failed[i] = failed;
  }
}

// ---- TMA-staged persistent driver -------------------------------------------------------------------
// One CTA per resident slot walks tiles of TILE = blockDim.x instances.  The input rows the skill
// actually reads (S::NIN of them; unused coordinates are never fetched) are brought into shared
// memory with 1-D bulk async copies (cp.async.bulk, the TMA engine) that complete on an mbarrier,
// STAGES tiles ahead of the arithmetic, so HBM reads run continuously underneath the fp64 work
// instead of every warp stalling on its own loads first.  Requirements checked by the host
// (csrc/clik_abi.cu): N even and 16-byte aligned base pointers, so every row segment of a tile is
// a legal bulk copy (16-byte aligned, size a multiple of 16).
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CLIK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra CLIK_DONE;\n"
      "bra CLIK_WAIT;\n"
      "CLIK_DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <class S, int TILE, int STAGES>
__device__ __forceinline__ void pinv_step_tma(long long N, long long ld, const double* __restrict__ t, int t_stride,
                                              const double* __restrict__ q, const double* __restrict__ x,
                                              const double* __restrict__ y, double* __restrict__ qdot,
                                              double* __restrict__ xdot, int* __restrict__ mode) {
  constexpr int NIN = S::NIN;
  __shared__ __align__(128) double buf[STAGES][Max<NIN, 1>::v][TILE];
  __shared__ __align__(8) unsigned long long full[STAGES];
  const int tid = threadIdx.x;
  const long long ntiles = (N + TILE - 1) / TILE;

  auto issue = [&](long long tile, int s) {
    const long long i0 = tile * TILE;
    const unsigned cnt = (unsigned)((N - i0 < TILE) ? (N - i0) : TILE);
    const unsigned bytes = cnt * 8u;
    mbar_expect_tx(&full[s], bytes * (unsigned)NIN);
#pragma unroll
    for (int k = 0; k < NIN; ++k) bulk_load(&buf[s][k][0], S::in_row(k, ld, t, q, x, y) + i0, bytes, &full[s]);
  };

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && NIN > 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const long long tile = (long long)blockIdx.x + (long long)s * gridDim.x;
      if (tile < ntiles) issue(tile, s);
    }
  }
  const double t_shared = (t_stride == 0) ? __ldg(t) : 0.0;
  int s = 0;
  unsigned parity = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i = tile * TILE + tid;
    double qv[Max<S::NQ, 1>::v], xv[Max<S::NX, 1>::v], yv[Max<S::NY, 1>::v];
    double tv = t_shared;
    if constexpr (NIN > 0) {
      mbar_wait(&full[s], parity);
      S::unstage(buf[s], tid, t_stride, tv, qv, xv, yv);
      __syncthreads();                       // every thread has its inputs in registers
      if (tid == 0) {
        const long long nt = tile + (long long)STAGES * gridDim.x;
        if (nt < ntiles) issue(nt, s);       // refill this stage STAGES tiles ahead
      }
    }
    if (i < N) {
      double v[S::NS];
      const int accepted = solve_instance<S>(tv, qv, xv, yv, v);
      store_instance<S>(ld, i, v, accepted, qdot, xdot, mode);
    }
    if (++s == STAGES) {
      s = 0;
      parity ^= 1u;
    }
  }
}

}  // namespace clik
