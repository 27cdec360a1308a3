// clik_pinv_group.cuh — sub-warp ("group per instance") mapping of the SRMTP mode search, fp64, sm_100a.
//
// clik_pinv.cuh maps one controller instance to one thread and keeps every matrix in registers; that
// is the right mapping while the per-instance data fits the register file and all threads of a warp
// walk the same code.  It stops being right in two situations, and this header is the other mapping
// for them (BASELINE.json north_star: "one warp per instance on small dense matrices held in
// registers/shared memory with warp-shuffle reductions"):
//
//   * the run-time mode search (reference pseudo_inverse.py:530-550 past the statically compiled
//     modes): few instances need it, each walks its own sequence of modes over row lists that are
//     only known at run time (dynamic indexing = local memory in the thread mapping, and the other
//     31 lanes of the warp wait);
//   * skills whose stacked Jacobian does not fit a thread's registers (the <= 12 x 30 end of the
//     north star's size range).
//
// Here CLIK_GROUP (8) lanes of a warp cooperate on one instance.  The skill is still evaluated one
// instance per thread (straight-line generated code has no intra-instance parallelism to offer); the
// result is handed to the group through a shared-memory slot, and the algebra is distributed over the
// lanes by matrix ENTRY / ROW: every lane computes whole dot products for the Gram entries, Cholesky
// rows, projected-velocity rows and in-tangent-cone tests it owns, results are exchanged through the
// slot with __syncwarp(group mask), and the group's verdict on a mode is one __all_sync.  A warp
// evaluates 32 instances and then serves them in CLIK_GROUP rounds of 32 / CLIK_GROUP groups.
// The algebra is the literal per-mode walk of dynamic_mode() (pseudo_inverse.py:274-443), so results
// agree with the thread mapping to rounding.
//
// Used as (a) the tail pass after clik_pinv_fast_kernel for skills with a run-time mode tail
// (instances the static modes rejected are handed over through mode[] = PINV_PENDING and gathered
// densely, like the QP tail), and (b) on request (CLIK_PINV_GROUP=1) for whole batches — the A/B
// measurement of the two mappings in DESIGN.md §4.
#pragma once
#include "clik_pinv.cuh"

#ifndef CLIK_GROUP
#define CLIK_GROUP 8
#endif

namespace clik {

constexpr int GRP = CLIK_GROUP;            // lanes per instance (a power of two <= 32)
constexpr int PINV_TAIL_TILE = 1024;

template <class S> struct GroupSlot {
  static constexpr int NS = S::NS, M = S::M;
  static constexpr int MAXK = S::M + S::MAXROWS;   // every row once + the doubled first equality
  double J[Max<M * NS, 1>::v];
  double des[Max<M, 1>::v], e[Max<M, 1>::v], jt[Max<M, 1>::v], smin[Max<M, 1>::v], smax[Max<M, 1>::v];
  double rmask[Max<M, 1>::v];
  double tw[Max<S::NEQC, 1>::v][NS];       // P(J_c) des_c per Eq / VelEq constraint (mode independent)
  double G[NS * (NS + 1) / 2];             // packed lower triangle of the SPD system (order <= NS)
  double rinv[NS], z[NS];
  double b[Max<MAXK, 1>::v];
  double corr[NS], v[NS];
  int accepted;                            // in: SLOT_READY / SLOT_EMPTY, out: accepted mode (-1 none)
};
enum : int { SLOT_READY = -3, SLOT_EMPTY = -4 };

__device__ __forceinline__ void gsync(unsigned gmask) { __syncwarp(gmask); }

// out <- P(J_rows) b for a run-time row list, distributed over the GRP lanes of one group.
// b and out live in the slot (b is read before out is written, so they may not alias).
template <class S>
__device__ void group_pinv_times(GroupSlot<S>& sl, const int* rows, int K, const double* b, double* out,
                                 int gl, unsigned gmask) {
  constexpr int NS = S::NS;
  const bool wide = S::DAMPED ? (NS >= K) : (K < NS);
  const double lam = S::DAMPED ? S::LAMBDA : 0.0;
  const int n = wide ? K : NS;
  const double* J = sl.J;
  // Gram matrix: entry (a, c) is one dot product, entries dealt round-robin to the lanes
  int cnt = 0;
  for (int a = 0; a < n; ++a) {
    for (int c = 0; c <= a; ++c, ++cnt) {
      if (cnt % GRP != gl) continue;
      double acc = 0.0;
      if (wide) {
        for (int j = 0; j < NS; ++j) acc = fma(J[rows[a] * NS + j], J[rows[c] * NS + j], acc);
      } else {
        for (int r = 0; r < K; ++r) acc = fma(J[rows[r] * NS + a], J[rows[r] * NS + c], acc);
      }
      sl.G[a * (a + 1) / 2 + c] = (a == c) ? acc + lam : acc;
    }
  }
  if (wide) {
    for (int a = gl; a < K; a += GRP) sl.z[a] = b[a];
  } else {
    for (int i = gl; i < NS; i += GRP) {
      double acc = 0.0;
      for (int r = 0; r < K; ++r) acc = fma(J[rows[r] * NS + i], b[r], acc);
      sl.z[i] = acc;
    }
  }
  gsync(gmask);
  // Cholesky by columns: the pivot is recomputed by every lane (same value), the rows below it are
  // dealt to the lanes
  for (int j = 0; j < n; ++j) {
    double s = sl.G[j * (j + 1) / 2 + j];
    for (int k = 0; k < j; ++k) s = fma(-sl.G[j * (j + 1) / 2 + k], sl.G[j * (j + 1) / 2 + k], s);
    const double r = rsqrt(s);
    if (gl == 0) sl.rinv[j] = r;
    for (int i = j + 1 + gl; i < n; i += GRP) {
      double t = sl.G[i * (i + 1) / 2 + j];
      for (int k = 0; k < j; ++k) t = fma(-sl.G[i * (i + 1) / 2 + k], sl.G[j * (j + 1) / 2 + k], t);
      sl.G[i * (i + 1) / 2 + j] = t * r;
    }
    gsync(gmask);
  }
  // the two substitutions are a dependent chain of length n: one lane
  if (gl == 0) {
    for (int i = 0; i < n; ++i) {
      double s = sl.z[i];
      for (int k = 0; k < i; ++k) s = fma(-sl.G[i * (i + 1) / 2 + k], sl.z[k], s);
      sl.z[i] = s * sl.rinv[i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = sl.z[i];
      for (int k = i + 1; k < n; ++k) s = fma(-sl.G[k * (k + 1) / 2 + i], sl.z[k], s);
      sl.z[i] = s * sl.rinv[i];
    }
  }
  gsync(gmask);
  for (int j = gl; j < NS; j += GRP) {
    if (wide) {
      double acc = 0.0;
      for (int a = 0; a < K; ++a) acc = fma(J[rows[a] * NS + j], sl.z[a], acc);
      out[j] = acc;
    } else {
      out[j] = sl.z[j];
    }
  }
  gsync(gmask);
}

// v += (I - P(J_stack) rJ_stack) w         pseudo_inverse.py:387-394 / :434-441
template <class S>
__device__ void group_nullspace_add(GroupSlot<S>& sl, const int* stack, int k, const double* w, int gl,
                                    unsigned gmask) {
  constexpr int NS = S::NS;
  for (int a = gl; a < k; a += GRP) {
    double acc = 0.0;
    for (int j = 0; j < NS; ++j) acc = fma(sl.J[stack[a] * NS + j], w[j], acc);
    sl.b[a] = (S::MULTIDIM && S::row_is_set(stack[a])) ? acc * sl.rmask[stack[a]] : acc;
  }
  gsync(gmask);
  group_pinv_times<S>(sl, stack, k, sl.b, sl.corr, gl, gmask);
  for (int j = gl; j < NS; j += GRP) sl.v[j] += w[j] - sl.corr[j];   // lane j % GRP owns v[j] throughout
}

// One mode: the constraint walk of dynamic_mode() with the loops dealt to the lanes.  Uniform result.
template <class S>
__device__ bool group_mode(GroupSlot<S>& sl, unsigned mask, int gl, unsigned gmask) {
  constexpr int NS = S::NS;
  int stack[Max<GroupSlot<S>::MAXK, 1>::v];   // same contents in every lane
  int k = 0;
  for (int j = gl; j < NS; j += GRP) sl.v[j] = 0.0;
  for (int c = 0; c < S::NC; ++c) {
    const int kind = S::kind(c), r0 = S::row0(c), m = S::rows(c);
    if (kind == KIND_EQ || kind == KIND_VELEQ) {
      const double* w = sl.tw[S::eq_index(c)];
      const bool first = (k == 0);
      if (first) {                                                     // :317-326 / :327-335
        for (int j = gl; j < NS; j += GRP) sl.v[j] += w[j];
        for (int a = 0; a < m; ++a) stack[k++] = r0 + a;
      }
      if (!first || kind == KIND_EQ) {                                 // a first Eq also runs :382-396 (A1)
        group_nullspace_add<S>(sl, stack, k, w, gl, gmask);
        for (int a = 0; a < m; ++a) stack[k++] = r0 + a;
      }
    } else if (kind == KIND_SET) {
      if ((mask >> S::set_index(c)) & 1u) {
        if (S::CONV_LAST && c == S::NC - 1 && k > 0)                   // converge_final_set_to_max, :337-356
          group_nullspace_add<S>(sl, stack, k, sl.tw[S::eq_index(c)], gl, gmask);
        for (int a = 0; a < m; ++a) stack[k++] = r0 + a;
      }
    }
  }
  gsync(gmask);                                                        // v is complete
  bool ok = true;
  for (int c = gl; c < S::NC; c += GRP) {                              // in-tangent-cone tests, sets over lanes
    if (S::kind(c) != KIND_SET || ((mask >> S::set_index(c)) & 1u)) continue;
    const int r0 = S::row0(c), m = S::rows(c);
    double de[Max<S::MAXROWS, 1>::v];
    for (int a = 0; a < m; ++a) {
      double dot = 0.0;
      for (int j = 0; j < NS; ++j) dot = fma(sl.J[(r0 + a) * NS + j], sl.v[j], dot);
      de[a] = sl.jt[r0 + a] + dot;
    }
    if (m == 1) {
      ok = ok && in_tangent_cone(sl.e[r0], de[0], sl.smin[r0], sl.smax[r0]);
    } else {                                                           // vector-valued set, :222-252
      bool above = true, below = true, corner = true;
      double proj = 0.0, dd = 0.0, oo = 0.0;
      for (int a = 0; a < m; ++a) {
        const double le = sl.e[r0 + a] - sl.smin[r0 + a], ue = sl.e[r0 + a] - sl.smax[r0 + a];
        above = above && (le >= 1e-12);
        below = below && (ue <= 1e-12);
        const double s1 = (double)((le > 0.0) - (le < 0.0)), s2 = (double)((ue > 0.0) - (ue < 0.0));
        corner = corner && (s1 == s2);
        const double od = (s1 + s2) / 2.0;
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
  return __all_sync(gmask, ok) != 0;
}

// Everything for the instance in the slot: task velocities, then the activation map from `from_mode`
// on, in the reference's order, until a mode is admissible.  Leaves v and the accepted index in the slot.
template <class S>
__device__ void group_solve(GroupSlot<S>& sl, int from_mode, int gl, unsigned gmask) {
  constexpr int NS = S::NS;
  for (int c = 0; c < S::NC; ++c) {
    const int kind = S::kind(c);
    if (kind == KIND_EQ || kind == KIND_VELEQ || (kind == KIND_SET && S::CONV_LAST && c == S::NC - 1)) {
      const int r0 = S::row0(c), m = S::rows(c);
      int rows[Max<S::MAXROWS, 1>::v];
      for (int a = 0; a < m; ++a) rows[a] = r0 + a;
      for (int a = gl; a < m; a += GRP) sl.b[a] = sl.des[r0 + a];
      gsync(gmask);
      group_pinv_times<S>(sl, rows, m, sl.b, sl.tw[S::eq_index(c)], gl, gmask);
    }
  }
  int accepted = -1;
  for (int mi = from_mode; mi < S::NMODES; ++mi) {
    if (group_mode<S>(sl, S::mode_mask(mi), gl, gmask)) {
      accepted = mi;
      break;
    }
  }
  if (accepted < 0) {
    for (int j = gl; j < NS; j += GRP) sl.v[j] = 0.0;                  // pseudo_inverse.py:551-555
  }
  if (gl == 0) sl.accepted = accepted;
}

template <class S>
__device__ __forceinline__ void slot_fill(GroupSlot<S>& sl, const PinvData<S>& d) {
#pragma unroll
  for (int k = 0; k < S::M * S::NS; ++k) sl.J[k] = d.J[k];
#pragma unroll
  for (int r = 0; r < S::M; ++r) {
    if (S::row_is_set(r)) {
      sl.e[r] = d.e[r];
      sl.jt[r] = d.jt[r];
      sl.smin[r] = d.smin[r];
      sl.smax[r] = d.smax[r];
      if constexpr (S::MULTIDIM) sl.rmask[r] = d.rmask[r];
      if constexpr (S::CONV_LAST) {
        if (r >= S::row0(S::NC - 1)) sl.des[r] = d.des[r];
      }
    } else {
      sl.des[r] = d.des[r];
    }
  }
}

// Block size of the group kernel: two warps, or one when the slots of two would not fit the 48 KB of
// static shared memory (large skills).
template <class S> struct GroupGeometry {
  static constexpr int BLOCK = (sizeof(GroupSlot<S>) * (64 / GRP) + 8 * PINV_TAIL_TILE <= 44000) ? 64 : 32;
  static constexpr int NSLOT = BLOCK / GRP;
#ifdef __CUDA_ARCH__   // (the host test harness has no shared-memory limit and runs groups of one lane)
  static_assert(sizeof(GroupSlot<S>) * (32 / GRP) + 4 * PINV_TAIL_TILE <= 48000,
                "skill too large for the static shared-memory slots of the group kernel");
#endif
};

// only_pending != 0: serve the instances clik_pinv_fast_kernel left PINV_PENDING in mode[] (search
// from `from_mode`, the first mode that is not statically compiled); == 0: every instance, from mode
// `from_mode` (0 = the whole controller step in this mapping).
template <class S>
__device__ __forceinline__ void pinv_group_step(long long N, long long ld, const double* __restrict__ t, int t_stride,
                                                const double* __restrict__ q, const double* __restrict__ x,
                                                const double* __restrict__ y, double* __restrict__ qdot,
                                                double* __restrict__ xdot, int* __restrict__ mode, int from_mode,
                                                int only_pending) {
  constexpr int NSLOT = GroupGeometry<S>::NSLOT;
  constexpr int GPW = 32 / GRP;                       // groups per warp
  __shared__ GroupSlot<S> slots[NSLOT];
  __shared__ int list[PINV_TAIL_TILE];
  __shared__ int count;
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int g = lane / GRP, gl = lane % GRP;
  const unsigned gmask = (GRP >= 32) ? 0xffffffffu : (((1u << GRP) - 1u) << (g * GRP));
  for (long long base = (long long)blockIdx.x * PINV_TAIL_TILE; base < N; base += (long long)gridDim.x * PINV_TAIL_TILE) {
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < PINV_TAIL_TILE && base + k < N; k += blockDim.x) {
      if (!only_pending || mode[base + k] == PINV_PENDING) list[atomicAdd(&count, 1)] = k;
    }
    __syncthreads();
    const int n = count;
    for (int b0 = 0; b0 < n; b0 += blockDim.x) {
      const bool have = b0 + (int)threadIdx.x < n;
      const long long i = have ? base + list[b0 + threadIdx.x] : 0;
      PinvData<S> d;
      if (have) {                                     // one instance per thread: evaluate the skill
        double tv, qv[Max<S::NQ, 1>::v], xv[Max<S::NX, 1>::v], yv[Max<S::NY, 1>::v];
        load_instance<S>(ld, i, t, t_stride, q, x, y, tv, qv, xv, yv);
        S::eval(tv, qv, xv, yv, d);
      }
      // GRP rounds: in round r the owners lane / GPW == r hand their instance to group lane % GPW
      for (int r = 0; r < GRP; ++r) {
        const bool my_round = (lane / GPW == r);
        GroupSlot<S>& mine = slots[warp * GPW + lane % GPW];
        if (my_round) {
          if (have) slot_fill<S>(mine, d);
          mine.accepted = have ? SLOT_READY : SLOT_EMPTY;
        }
        __syncwarp();
        GroupSlot<S>& sl = slots[warp * GPW + g];
        if (sl.accepted == SLOT_READY) group_solve<S>(sl, from_mode, gl, gmask);
        __syncwarp();
        if (my_round && have) {
          double v[S::NS];
#pragma unroll
          for (int j = 0; j < S::NS; ++j) v[j] = mine.v[j];
          store_instance<S>(ld, i, v, mine.accepted, qdot, xdot, mode);
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

}  // namespace clik
