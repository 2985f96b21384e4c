// Host-side harness (TEST ONLY): the draw functions of cnf_ot_b200/csrc/philox.cuh compiled for the host, so the
// generator the kernels inline can be checked against the published known-answer vectors and the numpy
// restatement (oracle/philox.py) without a GPU.  Never loaded by the cnf_ot_b200 package.
#include <cstdint>
#include "../../cnf_ot_b200/csrc/philox.cuh"

using namespace cnfot;

extern "C" {

void hs_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  PhiloxWords w = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = w.w[i];
}

uint64_t hs_philox_salt(int kind, uint64_t n) { return philox_salt(kind, n); }

void hs_philox_rows(uint64_t key, uint32_t step, int source, int64_t global_rows, int64_t row0, int64_t rows, int dim,
                    float* out) {
  const uint64_t kn = key ^ philox_salt(kDrawNormal, (uint64_t)global_rows);
  const uint64_t kc = key ^ philox_salt(kDrawCategorical, (uint64_t)global_rows);
  for (int64_t r = 0; r < rows; ++r) philox_row(kn, kc, step, source, (uint64_t)(row0 + r), dim, out + r * dim);
}

void hs_philox_times(uint64_t key, uint32_t step, int n_t, float horizon, float* out) {
  const uint64_t kt = key ^ philox_salt(kDrawUniform, (uint64_t)n_t);
  for (int i = 0; i < n_t; ++i) out[i] = philox_time(kt, step, i, horizon);
}

}
