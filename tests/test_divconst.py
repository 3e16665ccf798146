"""The CUDA warp replaces the two fp64 divisions of Event::apply_project (event.h:164-168), x/127 and
x/10000, by  q0 = x*r;  q = fma(fma(-b, q0, x), r, q0)  with r = rn(1/b).  That is only legitimate if
it returns the correctly rounded quotient for every dividend that can occur -- and the dividends
are always f32 values widened to f64 -- so check ALL 2^32 of them (sampled if the host has no FMA)."""
import os
import subprocess
import tempfile

SRC = r'''
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(int argc, char **argv) {
  const long long stride = argc > 1 ? atoll(argv[1]) : 1;
  const double bs[2] = {127.0, 10000.0};
  long long bad = 0, n = 0;
  for (int k = 0; k < 2; ++k) {
    const double b = bs[k], r = 1.0 / b;
    #pragma omp parallel for reduction(+:bad,n) schedule(static)
    for (long long u = 0; u < (1LL << 32); u += stride) {
      uint32_t bits = (uint32_t)u; float f; memcpy(&f, &bits, 4);
      if (!isfinite(f)) continue;
      double a = f, q = a / b, q0 = a * r;
      double q1 = fma(fma(-b, q0, a), r, q0);
      n++;
      if (q1 != q) bad++;
    }
  }
  printf("%lld %lld\n", n, bad);
  return 0;
}
'''


def test_reciprocal_fma_division_is_exact_for_all_f32_dividends():
    has_fma = " fma " in open("/proc/cpuinfo").read()
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write("#include <stdlib.h>\n" + SRC)
        exe = os.path.join(d, "t")
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp"] + (["-mfma"] if has_fma else []) + [c, "-o", exe, "-lm"])
        out = subprocess.check_output([exe, "1" if has_fma else "4099"], text=True).split()
    n, bad = int(out[0]), int(out[1])
    assert n > (8e9 if has_fma else 1e6)
    assert bad == 0


SRC_F32 = r'''
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
/* unpack_avg_fast (bf_device.cuh): mean = s / f32(c) as  q0 = s*y; q = fmaf(fmaf(-c, q0, s), y, q0),
 * y = RN(1/c) from a table, c = 1..1023 (BF_RCP_TAB). */
static inline float seq(float s, float c, float y) {
  float q0 = s * y;
  return fmaf(fmaf(-c, q0, s), y, q0);
}
int main(int argc, char **argv) {
  const long long stride = atoll(argv[1]);
  const int exhaustive_c = atoi(argv[2]);
  long long bad = 0, n = 0;
  /* (1) every c of the table x a strided sweep over ALL positive finite f32 bit patterns */
  #pragma omp parallel for reduction(+:bad,n) schedule(dynamic, 8)
  for (int c = 1; c < 1024; ++c) {
    const float cf = (float)c, y = 1.0f / cf;
    for (long long u = 0x00800000LL + c; u < 0x7f000000LL; u += stride) {
      uint32_t bits = (uint32_t)u; float s; memcpy(&s, &bits, 4);
      const float want = s / cf;
      if (want < 1.2e-38f) continue;   /* subnormal quotients do not occur (s >= 1e-9, c < 2^17) */
      n++;
      if (seq(s, cf, y) != want) bad++;
    }
    /* (2) quotients engineered to sit next to a rounding boundary: s = RN(c * (Q + 1/2 ulp)) */
    uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)c;
    for (int k = 0; k < 200000; ++k) {
      st = st * 6364136223846793005ull + 1442695040888963407ull;
      const uint32_t Q = 0x800000u | (uint32_t)((st >> 20) & 0x7fffffu);
      const int e = (int)((st >> 50) % 40) - 50;
      const double mid = ldexp((double)Q + 0.5, e);
      float s = (float)(mid * (double)c);
      for (int d = -1; d <= 1; ++d) {
        float sd = d < 0 ? nextafterf(s, 0.0f) : d > 0 ? nextafterf(s, INFINITY) : s;
        const float want = sd / cf;
        n++;
        if (seq(sd, cf, y) != want) bad++;
      }
    }
  }
  /* (3) small counts (the common case) x EVERY f32 in [2^-31, 2): mean times from 0.5 ns to 2 s */
  for (int c = 1; c <= exhaustive_c; ++c) {
    const float cf = (float)c, y = 1.0f / cf;
    #pragma omp parallel for reduction(+:bad,n) schedule(static)
    for (long long u = 0x30000000LL; u < 0x40000000LL; ++u) {
      uint32_t bits = (uint32_t)u; float s; memcpy(&s, &bits, 4);
      n++;
      if (seq(s, cf, y) != s / cf) bad++;
    }
  }
  printf("%lld %lld\n", n, bad);
  return 0;
}
'''


def test_table_reciprocal_f32_division_is_correctly_rounded():
    """The fast unpack's divide (mean time = s / count) against IEEE f32 division."""
    has_fma = " fma " in open("/proc/cpuinfo").read()
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(SRC_F32)
        exe = os.path.join(d, "t")
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp", "-ffp-contract=off"] + (["-mfma"] if has_fma else []) + [c, "-o", exe, "-lm"])
        out = subprocess.check_output([exe, "4099" if has_fma else "400009", "12" if has_fma else "0"], text=True).split()
    n, bad = int(out[0]), int(out[1])
    assert n > (3e9 if has_fma else 1e6)
    assert bad == 0
