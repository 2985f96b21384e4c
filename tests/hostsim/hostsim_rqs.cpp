// Host-side harness for the device math headers (TEST ONLY, never shipped and
// never loaded by the cnf_ot_b200 package): runs the exact per-element
// functions the kernels inline, in plain loops, in float or double, so their
// values and hand-derived adjoints can be checked against the autograd oracle
// on a machine without a GPU.
#include <cstdint>
#include "../../cnf_ot_b200/csrc/rqs_math.cuh"

using namespace cnfot;

template <typename T, int K>
static void run(int64_t n, int dir, const T* v, const T* theta, int64_t stride, const T* gout,
                const T* gld, T* out, T* ld, int32_t* idx, T* gin, T* gtheta) {
  SplineConsts<T> c = make_spline_consts<T>(K, -10.0, 10.0, 1e-4, 1e-4);
  constexpr int P = 3 * K + 1;
  for (int64_t i = 0; i < n; ++i) {
    SplineState<T, K> st;
    const T* th = theta + i * stride;
    if (dir == 0) rqs_forward<T, K>(v[i], th, c, st, out[i], ld[i]);
    else rqs_inverse<T, K>(v[i], th, c, st, out[i], ld[i]);
    idx[i] = st.idx;
    if (gin) {
      T g[P];
      gin[i] = dir == 0 ? rqs_forward_bwd<T, K>(v[i], st, c, gout[i], gld[i], g)
                        : rqs_inverse_bwd<T, K>(v[i], st, c, gout[i], gld[i], g);
      for (int j = 0; j < P; ++j) gtheta[i * P + j] = g[j];
    }
  }
}

#define DISPATCH(T)                                                                     \
  switch (K) {                                                                          \
    case 3: run<T, 3>(n, dir, v, theta, stride, gout, gld, out, ld, idx, gin, gtheta); break;   \
    case 5: run<T, 5>(n, dir, v, theta, stride, gout, gld, out, ld, idx, gin, gtheta); break;   \
    case 8: run<T, 8>(n, dir, v, theta, stride, gout, gld, out, ld, idx, gin, gtheta); break;   \
    case 10: run<T, 10>(n, dir, v, theta, stride, gout, gld, out, ld, idx, gin, gtheta); break; \
    default: return 1;                                                                  \
  }                                                                                     \
  return 0;

extern "C" int hs_rqs_f64(int64_t n, int K, int dir, const double* v, const double* theta,
                          int64_t stride, const double* gout, const double* gld, double* out,
                          double* ld, int32_t* idx, double* gin, double* gtheta) {
  DISPATCH(double)
}
extern "C" int hs_rqs_f32(int64_t n, int K, int dir, const float* v, const float* theta,
                          int64_t stride, const float* gout, const float* gld, float* out,
                          float* ld, int32_t* idx, float* gin, float* gtheta) {
  DISPATCH(float)
}

// the compile-time constants of the fused kernels next to the run-time ones (tests pin the literals)
extern "C" void hs_spline_consts(int K, double* fixed6, double* runtime6) {
  SplineConsts<double> c = make_spline_consts<double>(K, -10.0, 10.0, 1e-4, 1e-4);
  double r[6] = {c.lo, c.hi, c.min_bin, c.bin_scale, c.min_slope, c.slope_offset};
  for (int i = 0; i < 6; ++i) runtime6[i] = r[i];
  auto put = [&](auto f) {
    double v[6] = {f.lo, f.hi, f.min_bin, f.bin_scale, f.min_slope, f.slope_offset};
    for (int i = 0; i < 6; ++i) fixed6[i] = v[i];
  };
  if (K == 3) put(FixedSplineConsts<double, 3>());
  else if (K == 5) put(FixedSplineConsts<double, 5>());
  else if (K == 8) put(FixedSplineConsts<double, 8>());
  else for (int i = 0; i < 6; ++i) fixed6[i] = 0.0;
}
