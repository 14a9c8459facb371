// TEST INFRASTRUCTURE ONLY: compiles the host/device header dc_math.cuh with g++ so that the per-point
// fp64 math of the kernels (closed-form 3x3 eigen-solver, SE(3) correction and its reverse mode) can be
// unit-tested without a GPU.  Never loaded by the product package.
#include "../../depth_correction_b200/csrc/dc_math.cuh"

extern "C" void hm_eig(const double* cov9, long n, double* lam3, double* V9) {
  for (long i = 0; i < n; ++i) {
    const double* c = cov9 + 9 * i;
    dc_sym3 C = {c[0], c[3], c[6], c[4], c[7], c[8]};
    double V[9];
    dc_sym3_eig(C, lam3 + 3 * i, V, 3);
    for (int j = 0; j < 3; ++j)
      for (int a = 0; a < 3; ++a) V9[9 * i + 3 * a + j] = V[3 * j + a];
  }
}

extern "C" void hm_pose(const double* P16, const double* d6, long n, double* T12) {
  for (long i = 0; i < n; ++i) dc_pose_compose(P16 + 16 * i, d6 + 6 * i, T12 + 12 * i);
}

extern "C" void hm_pose_bwd(const double* P16, const double* d6, const double* g12, long n, double* gd6) {
  for (long i = 0; i < n; ++i) dc_pose_compose_bwd(P16 + 16 * i, d6 + 6 * i, g12 + 12 * i, gd6 + 6 * i);
}

extern "C" double hm_pow(double g, double e) { return dc_pow_exp(g, e); }
