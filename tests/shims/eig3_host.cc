// TEST INFRASTRUCTURE: the host build of noetic-slam_b200/csrc/eig3.cuh (the same source the K3 kernel compiles; on the host
// the approximate reciprocal is a float division and cosf replaces __cosf), exported for tests/test_eig3.py.
#include "eig3.cuh"
extern "C" int plane_fast(const double* a, double* o, long n) {
  int refused = 0;
  for (long i = 0; i < n; i++) {
    ngicp::Sym3 A{a[6 * i], a[6 * i + 1], a[6 * i + 2], a[6 * i + 3], a[6 * i + 4], a[6 * i + 5]}, O;
    if (!ngicp::plane_regularize_fast(A, O)) {
      refused++;
      for (int j = 0; j < 6; j++) o[6 * i + j] = NAN;
      continue;
    }
    o[6 * i] = O.xx; o[6 * i + 1] = O.xy; o[6 * i + 2] = O.xz; o[6 * i + 3] = O.yy; o[6 * i + 4] = O.yz; o[6 * i + 5] = O.zz;
  }
  return refused;
}
