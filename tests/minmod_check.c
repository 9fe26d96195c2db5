/* minmod without divisions or sign sums (csrc/dgadj_burgers_fused.cu: mm_bc / mm3; dgadj_burgers.cu has the same
 * arithmetic) against the reference's formulation (utils/minmod.m:6-12): s = sum(sign(v))/m; where |s| == 1 the
 * result is s * min(|v|), else 0 -- bit for bit, on random triples rich in zeros, ties and mixed signs.        */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint64_t s = 0x2545F4914F6CDD1DULL;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static double val(void) {
  static const double pool[6] = {0.0, -0.0, 1.0, -1.0, 0.5, -0.5};
  if (rnd() % 4 == 0) return pool[rnd() % 6];                                   /* zeros and ties */
  return ((double)(rnd() >> 11) * (1.0 / 9007199254740992.0) - 0.5) * ldexp(1.0, (int)(rnd() % 12) - 6);
}
typedef struct { int pos, neg; double t; } MmBC;
static MmBC mm_bc(double b, double c) {
  MmBC r;
  r.pos = (b > 0.0) & (c > 0.0);
  r.neg = (b < 0.0) & (c < 0.0);
  r.t = ((b < c) == r.pos) ? b : c;
  return r;
}
static double mm3(double a, MmBC q) {
  const double r = ((a < q.t) == q.pos) ? a : q.t;
  return ((q.pos & (a > 0.0)) | (q.neg & (a < 0.0))) ? r : 0.0;
}
static double sgn(double x) { return (x > 0.0) - (x < 0.0); }
static double minmod_ref(double a, double b, double c) {
  const double sm = (sgn(a) + sgn(b) + sgn(c)) / 3.0;
  if (fabs(sm) != 1.0) return 0.0;
  return sm * fmin(fabs(a), fmin(fabs(b), fabs(c)));
}
int main(void) {
  long bad = 0, nz = 0;
  const long n = 5000000;
  for (long i = 0; i < n; ++i) {
    const double a = val(), b = val(), c = val();
    const double x = mm3(a, mm_bc(b, c)), y = minmod_ref(a, b, c);
    nz += y != 0.0;
    if (memcmp(&x, &y, sizeof(double)) && !(x == 0.0 && y == 0.0)) {
      if (bad < 5) printf("minmod(%a, %a, %a): %a vs %a\n", a, b, c, x, y);
      ++bad;
    }
  }
  printf("%ld triples (%ld with a non-zero result): %ld differences\n", n, nz, bad);
  /* max|u| over a warp as two 32-bit reductions on the bit pattern (warp_max_nn: values >= 0, or exactly -1.0 in
   * idle lanes): the maximum of the high words, then of the low words of the lanes that hold it */
  long badmax = 0;
  for (long i = 0; i < 200000; ++i) {
    double v[32], ref = -1.0;
    const int idle = (int)(rnd() % 33);
    for (int l = 0; l < 32; ++l) {
      v[l] = (l < idle) ? -1.0 : fabs(val());
      if (rnd() % 16 == 0 && l > 0) v[l] = v[l - 1];        /* ties */
      if (v[l] > ref) ref = v[l];
    }
    int32_t mh = INT32_MIN;
    uint32_t ml = 0;
    for (int l = 0; l < 32; ++l) { uint64_t u; memcpy(&u, &v[l], 8); const int32_t hi = (int32_t)(u >> 32); if (hi > mh) mh = hi; }
    for (int l = 0; l < 32; ++l) {
      uint64_t u; memcpy(&u, &v[l], 8);
      const uint32_t lo = ((int32_t)(u >> 32) == mh) ? (uint32_t)u : 0u;
      if (lo > ml) ml = lo;
    }
    const uint64_t ru = ((uint64_t)(uint32_t)mh << 32) | ml;
    double r; memcpy(&r, &ru, 8);
    if (memcmp(&r, &ref, 8)) ++badmax;
  }
  printf("warp maximum on the bit pattern: %ld differences\n", badmax);
  return (bad || badmax || nz < n / 20) ? 1 : 0;
}
