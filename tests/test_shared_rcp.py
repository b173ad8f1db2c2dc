"""bn_div_by_rcp (include/bn_portable_math.h): a / s through a shared, correctly rounded reciprocal has the SAME BITS as
the IEEE division on the domain bn_div_rcp_ok() states — and not outside it.  This is the arithmetic behind the
default-off shade experiment -DBN_EXP_SHARED_RCP (vecmath.cuh); the check runs the header's own code on the CPU
(gcc -ffp-contract=off: `1.0f / s` is the correctly rounded reciprocal that __frcp_rn returns on the device)."""
import os
import subprocess

SRC = r"""
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "bn_portable_math.h"
static uint64_t s = 88172645463325252ull;
static uint32_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); }
static float fromb(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static long run(long n, int emin, int emax, long* tested) {
  long bad = 0;
  for (long it = 0; it < n; ++it) {
    float b = fromb((rnd() & 0x007fffffu) | ((uint32_t)(emin + (int)(rnd() % (uint32_t)(emax - emin + 1))) << 23));
    float a = fromb((rnd() & 0x807fffffu) | ((uint32_t)(emin + (int)(rnd() % (uint32_t)(emax - emin + 1))) << 23));
    if (it % 7 == 0) b = fromb((bits(b) & 0xff800000u) | (0x7fffffu - (rnd() % 4)));   /* mantissas of all ones */
    if (it % 11 == 0) a = fromb((bits(a) & 0xff800000u) | (rnd() % 4));               /* just above a power of two */
    if (it % 13 == 0) b = fromb((bits(b) & 0xff800000u) | (rnd() % 4));
    ++*tested;
    if (bits(bn_div_by_rcp(a, b, 1.0f / b)) != bits(a / b)) ++bad;
  }
  return bad;
}
int main(void) {
  long t1 = 0, t2 = 0;
  long inside = run(40000000L, 127 - 60, 127 + 59, &t1);      /* |a|, |b| in [2^-60, 2^60): bn_div_rcp_ok */
  long outside = run(4000000L, 1, 20, &t2);                   /* operands near the smallest normal */
  int zeros = bits(bn_div_by_rcp(0.0f, 3.0f, 1.0f / 3.0f)) == bits(0.0f / 3.0f) && bits(bn_div_by_rcp(-0.0f, 3.0f, 1.0f / 3.0f)) == bits(-0.0f / 3.0f) &&
              bits(bn_div_by_rcp(-0.0f, -3.0f, 1.0f / -3.0f)) == bits(-0.0f / -3.0f);
  int guard = bn_div_rcp_ok(1.0f) && bn_div_rcp_ok(-8.7e-19f) && bn_div_rcp_ok(1.15e18f) && !bn_div_rcp_ok(8.6e-19f) && !bn_div_rcp_ok(1.16e18f) &&
              !bn_div_rcp_ok(0.0f) && !bn_div_rcp_ok(1.0f / 0.0f) && !bn_div_rcp_ok(0.0f / 0.0f);
  printf("%ld %ld %ld %ld %d %d\n", inside, t1, outside, t2, zeros, guard);
  return 0;
}
"""


def test_shared_reciprocal_division_is_bit_identical_on_its_domain(root, tmp_path):
    src, exe = tmp_path / "check.c", tmp_path / "check"
    src.write_text(SRC)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-I", os.path.join(root, "include"), "-o", str(exe), str(src), "-lm"], check=True)
    inside, n_inside, outside, n_outside, zeros, guard = (int(x) for x in subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True, timeout=120).stdout.split())
    assert n_inside == 40_000_000 and inside == 0          # same bits as a / b on the whole guarded domain
    assert outside > 0.001 * n_outside                     # ... and the guard is needed: it does fail near the smallest normal
    assert zeros == 1 and guard == 1
