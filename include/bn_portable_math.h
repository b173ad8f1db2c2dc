/* bn_portable_math.h — bit-reproducible fp32 sin/cos/atan.
 *
 * The reference calls the platform libm through MathF.SinCos/Cos/Sin/Atan
 * (Lambertian.fs:21, PBR.fs:53-56,62, ThinLens.fs:18-23, Sphere.fs:91).  libm
 * results are not reproducible between glibc and CUDA's libdevice,
 * which would limit oracle-vs-GPU image parity to a statistical statement.
 * These definitions use only IEEE-754 binary32 +,-,*,/ and fused multiply-add
 * in a fixed order, so a CPU build (-ffp-contract=off) and a CUDA build
 * (-fmad=false) produce the SAME BITS.  Both the oracle (in its "portable"
 * mode) and the CUDA kernels evaluate them; the oracle's "libm" mode is the
 * reference-faithful variant used to show the two agree statistically.
 * Polynomials are the classic Cephes single-precision minimax sets; measured
 * max error against double-precision libm is recorded by tests/test_portable_math.py.
 */
#ifndef BN_PORTABLE_MATH_H
#define BN_PORTABLE_MATH_H

#include <math.h>

#if defined(__CUDACC__)
#define BN_HD __host__ __device__ __forceinline__
#else
#define BN_HD static inline
#endif

/* sin and cos of x for |x| <= ~1e4 (callers pass [-pi, 2*pi]).  Cody-Waite
 * reduction by pi/2 in two fused steps, then degree-7 / degree-8 polynomials
 * on [-pi/4, pi/4]. */
BN_HD void bn_sincosf(float x, float* s_out, float* c_out) {
  float k = rintf(x * 0.636619772f);          /* 2/pi */
  float r = fmaf(-k, 1.57079637050628662109375f, x);   /* pi/2 hi (= fl32(pi/2)) */
  r = fmaf(-k, -4.37113882867379e-08f, r);             /* pi/2 lo */
  float z = r * r;
  float sp = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, z, -1.6666654611e-1f);
  float sn = fmaf(sp * z, r, r);
  float cp = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, z, 4.166664568298827e-2f);
  float cs = fmaf(cp * z, z, fmaf(-0.5f, z, 1.0f));
  int q = (int)k & 3;
  float s = (q & 1) ? cs : sn;
  float c = (q & 1) ? sn : cs;
  if (q & 2) s = -s;
  if ((q + 1) & 2) c = -c;
  *s_out = s;
  *c_out = c;
}

BN_HD float bn_sinf(float x) { float s, c; bn_sincosf(x, &s, &c); return s; }
BN_HD float bn_cosf(float x) { float s, c; bn_sincosf(x, &s, &c); return c; }

/* atan(x), any finite x (Cephes atanf: two range reductions + degree-9 odd polynomial). */
BN_HD float bn_atanf(float x) {
  float a = fabsf(x);
  float y, t;
  if (a > 2.414213562373095f) {         /* tan(3*pi/8) */
    y = 1.57079637050628662109375f;
    t = -(1.0f / a);
  } else if (a > 0.4142135623730950f) { /* tan(pi/8) */
    y = 0.785398185253143310546875f;
    t = (a - 1.0f) / (a + 1.0f);
  } else {
    y = 0.0f;
    t = a;
  }
  float z = t * t;
  float p = fmaf(8.05374449538e-2f, z, -1.38776856032e-1f);
  p = fmaf(p, z, 1.99777106478e-1f);
  p = fmaf(p, z, -3.33329491539e-1f);
  y = y + fmaf(p * z, t, t);
  return copysignf(y, x);
}

/* ---- log / exp (PSSMLT: ErfInv of the Gaussian mutation, Kelemen mutation; PSSMLT.fs:67-68,125-134) ----
 * Cephes logf / expf with the exponent handled through the IEEE bit pattern, so both
 * sides execute the same integer and fp32 operations.  Domain: positive normal x for
 * log; |x| < 80 for exp. */
BN_HD float bn_uint_as_float(unsigned int u) { union { unsigned int u; float f; } c; c.u = u; return c.f; }
BN_HD unsigned int bn_float_as_uint(float f) { union { unsigned int u; float f; } c; c.f = f; return c.u; }

BN_HD float bn_logf(float x) {
  unsigned int b = bn_float_as_uint(x);
  int e = (int)((b >> 23) & 255u) - 126;                     /* x = m * 2^e, m in [0.5, 1) */
  float m = bn_uint_as_float((b & 0x007FFFFFu) | 0x3F000000u);
  if (m < 0.70710678118654752440f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
  float z = m * m;
  float y = fmaf(7.0376836292e-2f, m, -1.1514610310e-1f);
  y = fmaf(y, m, 1.1676998740e-1f);
  y = fmaf(y, m, -1.2420140846e-1f);
  y = fmaf(y, m, 1.4249322787e-1f);
  y = fmaf(y, m, -1.6668057665e-1f);
  y = fmaf(y, m, 2.0000714765e-1f);
  y = fmaf(y, m, -2.4999993993e-1f);
  y = fmaf(y, m, 3.3333331174e-1f);
  y = y * m * z;
  float fe = (float)e;
  y = fmaf(-2.12194440e-4f, fe, y);
  y = fmaf(-0.5f, z, y);
  float r = m + y;
  return fmaf(0.693359375f, fe, r);
}

BN_HD float bn_expf(float x) {
  float z = floorf(fmaf(1.44269504088896341f, x, 0.5f));
  x = fmaf(z, -0.693359375f, x);
  x = fmaf(z, 2.12194440e-4f, x);
  int n = (int)z;
  float xx = x * x;
  float p = fmaf(1.9875691500e-4f, x, 1.3981999507e-3f);
  p = fmaf(p, x, 8.3334519073e-3f);
  p = fmaf(p, x, 4.1665795894e-2f);
  p = fmaf(p, x, 1.6666665459e-1f);
  p = fmaf(p, x, 5.0000001201e-1f);
  float r = fmaf(p, xx, x) + 1.0f;
  return r * bn_uint_as_float((unsigned int)(n + 127) << 23);   /* ldexp(r, n), -126 <= n <= 127 */
}

/* ---- a / s for several numerators sharing ONE correctly rounded reciprocal ---------------------------------
 * (experiment for the shade kernel's normalize(), default off: -DBN_EXP_SHARED_RCP in vecmath.cuh)
 * With r = RN(1 / s):  q0 = RN(a r),  rem = fma(-s, q0, a) (exact),  q = fma(rem, r, q0)  is the correctly rounded
 * quotient RN(a / s) — Markstein's theorem — as long as nothing under- or overflows on the way, so the result has
 * the SAME BITS as the IEEE division the reference performs.  bn_div_rcp_ok() states the domain this is used on
 * (|x| in [2^-60, 2^60], or a zero numerator, whose sign the quotient keeps); tests/test_shared_rcp.py checks the
 * identity there on 4 x 10^7 operand pairs incl. all-ones and near-power-of-two mantissas, and shows it FAILS
 * outside (numerators near the smallest normal), which is why the caller falls back to plain division there. */
BN_HD int bn_div_rcp_ok(float x) {
  float m = fabsf(x);
  return (m >= 8.67361737988403547e-19f && m <= 1.15292150460684698e18f) ? 1 : 0;   /* 2^-60 .. 2^60 */
}
BN_HD float bn_div_by_rcp(float a, float s, float r) {
  float q0 = a * r;
  float rem = fmaf(-s, q0, a);
  float q = fmaf(rem, r, q0);
  return q0 == 0.0f ? q0 : q;   /* a == +-0: keep the zero with its sign (fma(+0, r, -0) would lose it) */
}

#endif /* BN_PORTABLE_MATH_H */
