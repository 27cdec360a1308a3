// clik_math.cuh — fp64 math helpers for the generated skill code.
//
// sincos: forward kinematics needs sin and cos of every joint angle (5-7 pairs per controller
// step), and in the first profile (profiles/r1_pinv_ncu_summary.md) the CUDA library's inlined
// sincos made up about half of the kernel's instructions, most of them UMOV/IMAD.MOV pairs that
// materialise 64-bit polynomial coefficients as immediates.  This version keeps the same
// structure (Cody-Waite reduction by pi/2, two minimax polynomials on [-pi/4, pi/4], quadrant
// select) but reads every coefficient from the constant bank, where it is a free instruction
// operand, and leaves arguments outside the fast-reduction range to the library routine.
//
// Polynomials: the classic fdlibm __kernel_sin / __kernel_cos coefficient sets (error < 1 ulp on
// the reduced interval).  Reduction: 3-term Cody-Waite, exact for |x| < ~1e5 (same bound the
// CUDA library uses for its fast path).
#pragma once

namespace clik {

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization
// attribute may begin while its predecessor on the stream is still draining.  Every step kernel tells the
// hardware at its start that a successor may be scheduled (free when nothing asks for it) and — as its last
// action, or before its first read when it consumes the predecessor's output — waits for the predecessor to
// have completed and flushed, so kernels still complete in stream order.  Both are no-ops in a launch without
// the attribute; clik_abi.cu sets it only where the caller declared launches independent
// (clik_skill_set_overlap) and between the two launches of one step.
__device__ __forceinline__ void pdl_launch_dependents() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

__constant__ double SC_TAB[20] = {
    6.36619772367581382433e-01,   // 0  2/pi
    1.57079632679489655800e+00,   // 1  pi/2 hi
    6.12323399573676603587e-17,   // 2  pi/2 mid
    -1.49738490485916983294e-33,  // 3  pi/2 lo (residual of hi+mid)
    -1.66666666666666324348e-01,  // 4  S1
    8.33333333332248946124e-03,   // 5  S2
    -1.98412698298579493134e-04,  // 6  S3
    2.75573137070700676789e-06,   // 7  S4
    -2.50507602534068634195e-08,  // 8  S5
    1.58969099521155010221e-10,   // 9  S6
    4.16666666666666019037e-02,   // 10 C1
    -1.38888888888741095749e-03,  // 11 C2
    2.48015872894767294178e-05,   // 12 C3
    -2.75573143513906633035e-07,  // 13 C4
    2.08757232129817482790e-09,   // 14 C5
    -1.13596475577881948265e-11,  // 15 C6
    0.0, 0.0, 0.0, 0.0};

// arguments outside the fast-reduction range (huge, inf, nan): library routine, kept out of line.
// Returned BY VALUE: with pointer outputs the caller's sin / cos variables have their address taken
// and live in local memory for the whole kernel (profiles/r2: 10 STL + 10 LDL per UR5 instance).
struct SinCos { double s, c; };
__device__ __noinline__ SinCos sincos_slow(double x) {
  SinCos r;
  sincos(x, &r.s, &r.c);
  return r;
}

// Fast path only, no range check: the generated code evaluates all joint angles with this
// branch-free routine first (so the independent polynomial chains interleave), then tests all
// arguments at once with sincos_in_range() and re-does the rare out-of-range ones out of line.
// |x| < 1e5 as an INTEGER test on the high word (1e5 = 0x40F86A00'00000000, low word zero, so the
// comparison is exact; NaN / inf have a larger high word and take the slow path): keeps five DSETPs
// per UR5 instance off the fp64 pipe, which is the kernel's busiest unit.
__device__ __forceinline__ bool sincos_in_range(double x) {
  return (unsigned)(__double2hiint(x) & 0x7fffffff) < 0x40F86A00u;
}

__device__ __forceinline__ void sincos_fast(double x, double* sp, double* cp) {
  const int k = __double2int_rn(x * SC_TAB[0]);
  const double kd = (double)k;
  double r = fma(-kd, SC_TAB[1], x);
  r = fma(-kd, SC_TAB[2], r);
  r = fma(-kd, SC_TAB[3], r);
  const double z = r * r;
  // sin(r) = r + r*z*(S1 + z*(S2 + ... ))
  double ps = fma(z, SC_TAB[9], SC_TAB[8]);
  ps = fma(z, ps, SC_TAB[7]);
  ps = fma(z, ps, SC_TAB[6]);
  ps = fma(z, ps, SC_TAB[5]);
  ps = fma(z, ps, SC_TAB[4]);
  const double s = fma(r * z, ps, r);
  // cos(r) = 1 + z*(-1/2 + z*(C1 + z*(C2 + ...)))
  double pc = fma(z, SC_TAB[15], SC_TAB[14]);
  pc = fma(z, pc, SC_TAB[13]);
  pc = fma(z, pc, SC_TAB[12]);
  pc = fma(z, pc, SC_TAB[11]);
  pc = fma(z, pc, SC_TAB[10]);
  pc = fma(z, pc, -0.5);
  const double c = fma(z, pc, 1.0);
  // quadrant: k mod 4 = 0: (s, c), 1: (c, -s), 2: (-s, -c), 3: (-c, s).  The sign flips are XORs on
  // the high word (integer pipe), not fp64 negations + selects.
  const bool swap = (k & 1) != 0;
  const double so = swap ? c : s;
  const double co = swap ? s : c;
  *sp = __hiloint2double(__double2hiint(so) ^ ((k & 2) << 30), __double2loint(so));
  *cp = __hiloint2double(__double2hiint(co) ^ (((k + 1) & 2) << 30), __double2loint(co));
}

}  // namespace clik
