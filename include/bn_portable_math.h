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

#endif /* BN_PORTABLE_MATH_H */
