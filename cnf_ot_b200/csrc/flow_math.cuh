// Per-row algorithms of the time-conditioned autoregressive RQS flow:
// conditioner MLP forward / backward, one flow pass in either direction with
// its log-det, and the reverse-mode pass that re-computes activations instead
// of storing them.  One "row" is one sample of the batch; on the device one
// thread owns one row and everything below lives in its registers / local
// memory, the conditioner weights are read (broadcast) from shared memory.
//
// Semantics follow the reference exactly:
//   conditioner          /root/reference/cnf_ot/models/flows.py:46-86
//   per-layer algebra    /root/reference/cnf_ot/models/autoregressive.py:76-136
//   stacking, direction  /root/reference/cnf_ot/models/flows.py:138-175,
//                        /root/reference/cnf_ot/models/conditional.py:147-177,217-243
//
// The weight-gradient reduction over rows is delegated to a `Sink` policy:
// the device sink stages (activation, adjoint) pairs in shared memory and
// reduces them CTA-wide (flow_kernels.cuh); the host test harness uses a
// plain accumulating sink.  Control flow is uniform across rows, which the
// device sink relies on (it contains CTA-wide barriers).
#pragma once

#include "rqs_math.cuh"

namespace cnfot {

constexpr int kMaxStateFloats = 64;  // (L+1)*D per pass kept in local memory
constexpr int kMaxDim = 32;          // rows wider than this use the GEMM path

// ---- parameter blob layout ---------------------------------------------------
// One flat fp32 buffer ("weights blob"); the gradient buffer has the same
// layout.  Pp = P rounded up to a multiple of 4 so every matrix row is 16-byte
// aligned (padding entries are zero and stay zero).
//   [ first (Pp) ]
//   for l in 0..L-1, for d in 1..D-1:          (d = position in the permutation)
//     W0 ((d+1) x H) b0 (H)  { Wm (H x H) bm (H) }_{m=1..M-1}  Wout (H x Pp) bout (Pp)
// Matrices are (in, out) row-major like haiku's `w`; each bias directly follows
// its matrix so [W; b] is one (in+1) x out matrix.
struct FlowLayout {
  int D, L, M, H, K, P, Pp;
  int mlp_const;     // floats of one MLP excluding the (d+1) x H input matrix
  int layer_stride;  // floats of one flow layer (all its MLPs)
  int total;         // floats of the whole blob
};

inline FlowLayout make_layout(int D, int L, int M, int H, int K) {
  FlowLayout f;
  f.D = D; f.L = L; f.M = M; f.H = H; f.K = K;
  f.P = 3 * K + 1;
  f.Pp = (f.P + 3) / 4 * 4;
  f.mlp_const = H + (M - 1) * (H * H + H) + H * f.Pp + f.Pp;
  // sum_{d=1}^{D-1} [(d+1) H + mlp_const]
  f.layer_stride = (D - 1) * f.mlp_const + H * ((D - 1) * (D + 2) / 2);
  f.total = f.Pp + L * f.layer_stride;
  return f;
}

template <int H, int K, int M>
struct NetCfg {
  static constexpr int kH = H, kK = K, kM = M;
  static constexpr int kP = 3 * K + 1;
  static constexpr int kPp = (kP + 3) / 4 * 4;
  static constexpr int kMlpConst = H + (M - 1) * (H * H + H) + H * kPp + kPp;
};

// Runtime (DC == 0) or compile-time (DC > 0) flow shape.
template <int DC, int LC>
struct Dims {
  int D_, L_;
  CNFOT_HD int D() const { return DC > 0 ? DC : D_; }
  CNFOT_HD int L() const { return LC > 0 ? LC : L_; }
};

template <class Net>
CNFOT_HD int mlp_offset(int D, int layer, int d) {
  int layer_stride = (D - 1) * Net::kMlpConst + Net::kH * ((D - 1) * (D + 2) / 2);
  return Net::kPp + layer * layer_stride + (d - 1) * Net::kMlpConst +
         Net::kH * ((d - 1) * (d + 2) / 2);
}

// coordinate handled at position d of layer l (alternating identity / reversed
// permutations: flows.py:141-143 with minimum_perm=True)
CNFOT_HD int perm_at(int layer, int d, int D) { return (layer & 1) ? (D - 1 - d) : d; }

// ---- 4-wide weight loads -------------------------------------------------------
template <typename T>
CNFOT_HD void load4(const T* p, T (&v)[4]) {
  v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
}
#if defined(__CUDA_ARCH__)
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 q = *reinterpret_cast<const float4*>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
#endif

// out[0..N) += v * row[0..N)
template <typename T, int N>
CNFOT_HD void axpy_row(T v, const T* row, T* out) {
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    T w[4];
    load4<T>(row + j, w);
#pragma unroll
    for (int q = 0; q < 4; ++q) out[j + q] += v * w[q];
  }
}

// sum_j g[j] * row[j]
template <typename T, int N>
CNFOT_HD T dot_row(const T* g, const T* row) {
  T acc0 = 0, acc1 = 0;
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    T w[4];
    load4<T>(row + j, w);
    acc0 += g[j] * w[0];
    acc1 += g[j + 1] * w[1];
    acc0 += g[j + 2] * w[2];
    acc1 += g[j + 3] * w[3];
  }
  return acc0 + acc1;
}

// Conditioner forward.  `in` holds n_in = d+1 values [t, conditioning coords].
// hid[m*H + j] receives the post-ReLU activations of hidden layer m.
template <typename T, class Net>
CNFOT_HD void mlp_forward(const T* W, int n_in, const T* in, T* hid, T* theta) {
  constexpr int H = Net::kH, M = Net::kM, Pp = Net::kPp;
  const T* b0 = W + n_in * H;
  T acc[H];
#pragma unroll
  for (int j = 0; j < H; ++j) acc[j] = b0[j];
  for (int i = 0; i < n_in; ++i) axpy_row<T, H>(in[i], W + i * H, acc);
#pragma unroll
  for (int j = 0; j < H; ++j) hid[j] = m_max(acc[j], (T)0);
  const T* Wm = b0 + H;
#pragma unroll
  for (int m = 1; m < M; ++m) {
    const T* bm = Wm + H * H;
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = bm[j];
#pragma unroll
    for (int i = 0; i < H; ++i) axpy_row<T, H>(hid[(m - 1) * H + i], Wm + i * H, acc);
#pragma unroll
    for (int j = 0; j < H; ++j) hid[m * H + j] = m_max(acc[j], (T)0);
    Wm = bm + H;
  }
  const T* bo = Wm + H * Pp;
#pragma unroll
  for (int j = 0; j < Pp; ++j) theta[j] = bo[j];
#pragma unroll
  for (int i = 0; i < H; ++i) axpy_row<T, Pp>(hid[(M - 1) * H + i], Wm + i * Pp, theta);
}

// Conditioner backward: pushes (activation, adjoint) pairs of every layer into
// the sink (weight + bias gradients) and returns the adjoint of the inputs
// in gin[1..n_in) (gin[0], the adjoint of t, is not needed by the train step).
template <typename T, class Net, class Sink>
CNFOT_HD void mlp_backward(const T* W, int w_off, int n_in, const T* in, const T* hid,
                           const T* gtheta, T* gin, Sink& sink) {
  constexpr int H = Net::kH, M = Net::kM, Pp = Net::kPp;
  int off_out = w_off + n_in * H + H + (M - 1) * (H * H + H);
  const T* Wout = W + (off_out - w_off);
  sink.template outer<H, Pp>(off_out, H, hid + (M - 1) * H, gtheta);
  T g[H];
#pragma unroll
  for (int i = 0; i < H; ++i) {
    T v = dot_row<T, Pp>(gtheta, Wout + i * Pp);
    g[i] = hid[(M - 1) * H + i] > (T)0 ? v : (T)0;
  }
#pragma unroll
  for (int m = M - 1; m >= 1; --m) {
    int off_m = w_off + n_in * H + H + (m - 1) * (H * H + H);
    const T* Wm = W + (off_m - w_off);
    sink.template outer<H, H>(off_m, H, hid + (m - 1) * H, g);
    T gp[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      T v = dot_row<T, H>(g, Wm + i * H);
      gp[i] = hid[(m - 1) * H + i] > (T)0 ? v : (T)0;
    }
#pragma unroll
    for (int i = 0; i < H; ++i) g[i] = gp[i];
  }
  sink.template outer<kMaxDim, H>(w_off, n_in, in, g);
  for (int i = 1; i < n_in; ++i) gin[i] = dot_row<T, H>(g, W + i * H);
}

// ---- one pass through the flow -----------------------------------------------
// dir 0: "sample direction", latent -> physical: layers 0..L-1, each applying
//        Autoregressive.inverse_and_log_det (conditioners read the layer INPUT,
//        spline inverse formula).
// dir 1: "log-prob direction", physical -> latent: layers L-1..0, each applying
//        Autoregressive.forward_and_log_det (conditioners read the OUTPUT being
//        built, sequential in d, spline forward formula).
// states[0..D) is the input; states[(s+1)*D ..] the result of step s.
// Returns the summed log-det of the pass.
template <typename T, class Net, class DimsT>
CNFOT_CALL T flow_pass(int dir, const DimsT& dm, const T* W, const SplineConsts<T>& sc, T t,
                     T* states) {
  constexpr int H = Net::kH, K = Net::kK, M = Net::kM, Pp = Net::kPp;
  const int D = dm.D(), L = dm.L();
  T ld_total = (T)0;
  for (int s = 0; s < L; ++s) {
    const int layer = dir == 0 ? s : L - 1 - s;
    const T* v = states + s * D;
    T* u = states + (s + 1) * D;
    const T* cvec = dir == 0 ? v : u;
    for (int d = 0; d < D; ++d) {
      const int i = perm_at(layer, d, D);
      T theta[Pp];
      if (d == 0) {
#pragma unroll
        for (int j = 0; j < Pp; ++j) theta[j] = W[j];
      } else {
        T in[kMaxDim + 1];
        in[0] = t;
        for (int j = 0; j < d; ++j) in[1 + j] = cvec[perm_at(layer, j, D)];
        T hid[M * H];
        mlp_forward<T, Net>(W + mlp_offset<Net>(D, layer, d), d + 1, in, hid, theta);
      }
      SplineState<T, K> st;
      T out, ld;
      if (dir == 0) rqs_inverse<T, K>(v[i], theta, sc, st, out, ld);
      else rqs_forward<T, K>(v[i], theta, sc, st, out, ld);
      u[i] = out;
      ld_total += ld;
    }
  }
  return ld_total;
}

// Reverse mode of flow_pass.  On entry g[0..D) is the adjoint of the pass
// output (states[L*D..]), gld the adjoint of the summed log-det; on exit g is
// the adjoint of the pass input.  gfirst[Pp] accumulates the adjoint of the
// shared `first` parameter (flushed to the sink once per kernel).
template <typename T, class Net, class DimsT, class Sink>
CNFOT_CALL void flow_pass_bwd(int dir, const DimsT& dm, const T* W, const SplineConsts<T>& sc,
                            T t, const T* states, T* g, T gld, T* gfirst, Sink& sink) {
  constexpr int H = Net::kH, K = Net::kK, M = Net::kM, Pp = Net::kPp;
  const int D = dm.D(), L = dm.L();
  for (int s = L - 1; s >= 0; --s) {
    const int layer = dir == 0 ? s : L - 1 - s;
    const T* v = states + s * D;
    const T* u = states + (s + 1) * D;
    const T* cvec = dir == 0 ? v : u;
    // dir 0: any order works, ascending keeps the in-place update valid;
    // dir 1: descending, so g[coordinate] is complete before it is consumed.
    for (int dd = 0; dd < D; ++dd) {
      const int d = dir == 0 ? dd : D - 1 - dd;
      const int i = perm_at(layer, d, D);
      T theta[Pp], gtheta[Pp];
      T in[kMaxDim + 1];
      T hid[M * H];
      int w_off = 0;
      if (d == 0) {
#pragma unroll
        for (int j = 0; j < Pp; ++j) theta[j] = W[j];
      } else {
        in[0] = t;
        for (int j = 0; j < d; ++j) in[1 + j] = cvec[perm_at(layer, j, D)];
        w_off = mlp_offset<Net>(D, layer, d);
        mlp_forward<T, Net>(W + w_off, d + 1, in, hid, theta);
      }
      SplineState<T, K> st;
      T out, ld;
#pragma unroll
      for (int j = 0; j < Pp; ++j) gtheta[j] = (T)0;
      if (dir == 0) {
        rqs_inverse<T, K>(v[i], theta, sc, st, out, ld);
        g[i] = rqs_inverse_bwd<T, K>(v[i], st, sc, g[i], gld, gtheta);
      } else {
        rqs_forward<T, K>(v[i], theta, sc, st, out, ld);
        g[i] = rqs_forward_bwd<T, K>(v[i], st, sc, g[i], gld, gtheta);
      }
      if (d == 0) {
#pragma unroll
        for (int j = 0; j < Pp; ++j) gfirst[j] += gtheta[j];
      } else {
        T gin[kMaxDim + 1];
        mlp_backward<T, Net, Sink>(W + w_off, w_off, d + 1, in, hid, gtheta, gin, sink);
        for (int j = 0; j < d; ++j) g[perm_at(layer, j, D)] += gin[1 + j];
      }
    }
  }
}

// log N(x; 0, I)
template <typename T>
CNFOT_HD T base_log_prob(const T* x, int D) {
  T acc = (T)0;
  for (int i = 0; i < D; ++i) acc += x[i] * x[i];
  return (T)-0.5 * acc - (T)0.91893853320467274178 * (T)D;
}

}  // namespace cnfot
