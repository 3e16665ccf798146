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
