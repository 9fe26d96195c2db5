/* The division-free quotient of the limited cells' path and of the time-DG solves (csrc: div_rn, div_rn_q):
 *   q = RN(x r),  r = RN(1/h);   result = RN(q + RN(x - h q) r)      (the residual is exact in an fma)
 * equals RN(x / h) -- Markstein's theorem, given the correctly rounded reciprocal.  Checked here bit for bit on
 * the host over random operands of the ranges the kernels see (element widths of uniform and refined meshes,
 * pivots of small dense systems; numerators from rounding level to 1e+6).  usage: div_rn_check [pairs]       */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s = 88172645463325252ULL;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static double ur(void) { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 2000000;
  long bad = 0;
  for (long i = 0; i < n; ++i) {
    double h;
    switch (rnd() % 4) {
      case 0: h = 6.283185307179586 / (double)(1 + rnd() % 4096); break;           /* uniform meshes */
      case 1: h = ldexp(1.0 + ur(), (int)(rnd() % 40) - 30); break;               /* any significand */
      case 2: h = -(0.5 + ur()) * ldexp(1.0, (int)(rnd() % 20) - 10); break;      /* negative pivots */
      default: h = 2.0 / (double)(1 + rnd() % 1024); break;
    }
    const double x = (ur() - 0.5) * ldexp(1.0, (int)(rnd() % 70) - 50);
    const double r = 1.0 / h;
    const double q = x * r;
    const double res = fma(fma(-h, q, x), r, q);
    const double ref = x / h;
    if (memcmp(&res, &ref, sizeof(double)) != 0 && !(res == 0.0 && ref == 0.0)) {   /* (the sign of a zero quotient may differ) */
      if (bad < 5) printf("x=%a h=%a: %a vs %a\n", x, h, res, ref);
      ++bad;
    }
  }
  printf("%ld pairs, %ld differences\n", n, bad);
  /* the stopping rule of the lane-group Newton march (csrc/dgadj_tdg.cu): `sqrt(e2) > tol` decided without the square
   * root outside a band around tol^2 -- the same decision everywhere, densely sampled around the band's edges */
  long badrule = 0;
  const double tols[4] = {1e-7, 1e-10, 3.3e-5, 0.25};
  for (int t = 0; t < 4; ++t) {
    const double tol = tols[t], tol2 = tol * tol, t2hi = tol2 * (1.0 + 1e-12), t2lo = tol2 * (1.0 - 1e-12);
    for (long i = 0; i < n / 4; ++i) {
      double e2;
      switch (rnd() % 3) {
        case 0: e2 = tol2 * (1.0 + (ur() - 0.5) * 4e-12); break;             /* inside and just outside the band */
        case 1: e2 = tol2 * (1.0 + (ur() - 0.5) * 1e-15 * (double)(rnd() % 64)); break;   /* the last ulps around tol^2 */
        default: e2 = tol2 * ldexp(1.0 + ur(), (int)(rnd() % 40) - 20); break;
      }
      int ab = e2 > t2hi;
      if (!ab && !(e2 < t2lo)) ab = sqrt(e2) > tol;
      if (ab != (sqrt(e2) > tol)) ++badrule;
    }
  }
  printf("stopping rule: %ld differences\n", badrule);
  return (bad || badrule) ? 1 : 0;
}
