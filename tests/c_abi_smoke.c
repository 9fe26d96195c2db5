/* Plain-C client of libdgadj.so: proves the boundary is a C ABI (no C++/torch types).
 * Without a GPU it must fail with DGADJ_ERR_NO_DEVICE (there is no CPU fallback); with an
 * sm_100 device it marches N=1, K=4 advection forward and checks mass conservation of the
 * periodic upwind scheme.  Built and run by tests/test_host_logic.py / tests/test_gpu_parity.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dgadj.h"

int main(void) {
  dgadj_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.device = 0; cfg.N = 1; cfg.K = 4; cfg.bc = DGADJ_BC_PERIODIC; cfg.inflow = DGADJ_INFLOW_ZERO;
  cfg.functional = DGADJ_FUNC_LINEAR; cfg.scheme = DGADJ_SCHEME_LSERK4; cfg.alpha = 0.0;
  dgadj_handle* h = NULL;
  int rc = dgadj_create(&cfg, &h);
  printf("version %d create rc %d\n", dgadj_version(), rc);
  if (rc == DGADJ_ERR_NO_DEVICE) { printf("no sm_100 device: no CPU fallback (expected without a GPU)\n"); return 3; }
  if (rc != DGADJ_OK) return 1;
  /* N = 1 operators on 4 elements of width 0.25 (StartUp1D by hand) */
  const double Dr[4] = {-0.5, 0.5, -0.5, 0.5}, LIFT[4] = {2.0, -1.0, -1.0, 2.0};
  const double s2 = sqrt(0.5), s32 = sqrt(1.5);
  const double V[4] = {s2, -s32, s2, s32};   /* orthonormal Legendre P~_0, P~_1 at r = -1, +1 */
  double rx[8], Fs[8];
  for (int i = 0; i < 8; ++i) { rx[i] = 8.0; Fs[i] = 8.0; }
  rc = dgadj_set_operators(h, 2, 4, Dr, LIFT, V, rx, Fs);
  if (rc) { printf("set_operators: %s\n", dgadj_last_error(h)); return 1; }
  double u0[8] = {0.1, 0.4, 0.4, 0.9, 0.9, 0.3, 0.3, 0.1}, uT[8];   /* [Np][K] */
  dgadj_march_args a;
  memset(&a, 0, sizeof(a));
  a.B = 1; a.S = 20; a.a = 1.0; a.dt = 1e-2;
  rc = dgadj_forward_host(h, &a, NULL, NULL, u0, uT, NULL);
  if (rc) { printf("forward_host: %s\n", dgadj_last_error(h)); return 1; }
  double m0 = 0, mT = 0;
  for (int i = 0; i < 8; ++i) { m0 += u0[i]; mT += uT[i]; }
  printf("mass %.15f -> %.15f, launches %lld\n", m0, mT, (long long)dgadj_launch_count(h));
  dgadj_destroy(h);
  return fabs(m0 - mT) < 1e-13 ? 0 : 2;
}
