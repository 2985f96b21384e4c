// Host-side harness (TEST ONLY, see hostsim_rqs.cpp): runs the per-row flow and
// loss-term functions the kernels inline, with a plain accumulating gradient
// sink, so values and adjoints can be checked against the autograd oracle
// without a GPU.  Never loaded by the cnf_ot_b200 package.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../cnf_ot_b200/csrc/step_host.h"

using namespace cnfot;

// plain arrays standing in for the CTA's shared-memory row tiles
template <typename T, class Net>
struct HostTiles {
  T in[kMaxDim + 4];
  T hid[Net::kM][Net::kH];
  T gh[Net::kM][Net::kH];
  T gth[Net::kPp];
  RowTiles<T, Net> view() {
    RowTiles<T, Net> tl;
    tl.in = in;
    for (int m = 0; m < Net::kM; ++m) { tl.hid[m] = hid[m]; tl.gh[m] = gh[m]; }
    tl.gth = gth;
    tl.sw_h = 0;
    tl.sw_p = 0;
    return tl;
  }
};

template <typename T, class Net>
struct HostSink {
  static constexpr bool kWarpMlp = false;
  double* G;     // gradient blob (double accumulation)
  const T* Wb;   // weight blob
  FirstKnots<T, Net::kK> fk;   // normalised knots of the shared `first` spline (built once, like the device contexts)
  HostSink(double* G_, const T* Wb_) : G(G_), Wb(Wb_) {
    first_knots_build<T, Net::kK>(Wb, make_spline_consts<T>(Net::kK, -10.0, 10.0, 1e-4, 1e-4), fk);
  }
  const T* first_params() const { return Wb; }
  const FirstKnots<T, Net::kK>& first_knots() const { return fk; }
  bool stash_on() const { return false; }
  const T* weights(int w_off, int) const { return Wb + w_off; }
  void begin() {}
  template <int K, int N>
  void dense_fwd(const T* xt, int sw, const T*, const T* Wm, int, T* y) { dense_fwd_from_tile<T, K, N>(xt, sw, Wm, y); }
  template <int K, int N>
  void dense_bwd(const T*, int, const T* g, const T* Wm, int, T* y) { dense_bwd_from_regs<T, K, N>(g, Wm, y); }
  static void outer(double* dst, int Na, const T* a, int Ng, const T* g) {
    for (int i = 0; i < Na; ++i)
      for (int j = 0; j < Ng; ++j) dst[i * Ng + j] += (double)a[i] * (double)g[j];
    for (int j = 0; j < Ng; ++j) dst[Na * Ng + j] += (double)g[j];
  }
  void commit(int w_off, int n_in, const RowTiles<T, Net>& tl) {
    constexpr int H = Net::kH, M = Net::kM, Pp = Net::kPp;
    outer(G + w_off, n_in, tl.in, H, tl.gh[0]);
    int off = w_off + n_in * H + H;
    for (int m = 1; m < M; ++m) {
      outer(G + off, H, tl.hid[m - 1], H, tl.gh[m]);
      off += H * H + H;
    }
    outer(G + off, H, tl.hid[M - 1], Pp, tl.gth);
  }
};

template <typename T, class Net>
static int flow_eval(int D, int L, int dir, int64_t rows, const T* W, const T* in, const T* cond,
                     int64_t cs, T* out, T* ld, int add_base) {
  if ((L + 1) * D > kMaxStateFloats || D > kMaxDim) return 1;
  Dims<0, 0> dm{D, L};
  SplineConsts<T> sc = make_spline_consts<T>(Net::kK, -10.0, 10.0, 1e-4, 1e-4);
  HostTiles<T, Net> ht;
  RowTiles<T, Net> tl = ht.view();
  HostSink<T, Net> sink(nullptr, W);
  for (int64_t r = 0; r < rows; ++r) {
    T st[kMaxStateFloats];
    for (int i = 0; i < D; ++i) st[i] = in[r * D + i];
    T l = dir == 0 ? flow_pass<0, T, Net, Dims<0, 0>, HostSink<T, Net>>(dm, sc, cond[r * cs], st, tl, sink)
                   : flow_pass<1, T, Net, Dims<0, 0>, HostSink<T, Net>>(dm, sc, cond[r * cs], st, tl, sink);
    for (int i = 0; i < D; ++i) out[r * D + i] = st[L * D + i];
    if (ld) {
      if (add_base) l = dir == 0 ? base_log_prob<T>(st, D) - l : base_log_prob<T>(st + L * D, D) + l;
      ld[r] = l;
    }
  }
  return 0;
}

template <typename T, class Net>
static int flow_vjp(int D, int L, int dir, int64_t rows, const T* W, const T* in, const T* cond,
                    int64_t cs, const T* gout, const T* gld, int add_base, T* gin, double* G) {
  if ((L + 1) * D > kMaxStateFloats || D > kMaxDim) return 1;
  Dims<0, 0> dm{D, L};
  SplineConsts<T> sc = make_spline_consts<T>(Net::kK, -10.0, 10.0, 1e-4, 1e-4);
  HostSink<T, Net> sink(G, W);
  HostTiles<T, Net> ht;
  RowTiles<T, Net> tl = ht.view();
  FirstGrad<T, Net::kK> gfirst = {};
  for (int64_t r = 0; r < rows; ++r) {
    T st[kMaxStateFloats], g[kMaxDim];
    for (int i = 0; i < D; ++i) st[i] = in[r * D + i];
    if (dir == 0) flow_pass<0, T, Net, Dims<0, 0>, HostSink<T, Net>>(dm, sc, cond[r * cs], st, tl, sink);
    else flow_pass<1, T, Net, Dims<0, 0>, HostSink<T, Net>>(dm, sc, cond[r * cs], st, tl, sink);
    T gl = gld ? gld[r] : (T)0;
    for (int i = 0; i < D; ++i) g[i] = gout[r * D + i];
    T gl_pass = gl;
    if (add_base) {
      if (dir == 0) gl_pass = -gl;                                         // lp = logN(in) - fldj
      else for (int i = 0; i < D; ++i) g[i] += gl * (-st[L * D + i]);     // lp = logN(out) + ildj
    }
    if (dir == 0)
      flow_pass_bwd<0, T, Net, Dims<0, 0>, HostSink<T, Net>>(dm, sc, cond[r * cs], st, g, gl_pass,
                                                             gfirst, tl, sink);
    else
      flow_pass_bwd<1, T, Net, Dims<0, 0>, HostSink<T, Net>>(dm, sc, cond[r * cs], st, g, gl_pass,
                                                             gfirst, tl, sink);
    if (add_base && dir == 0) for (int i = 0; i < D; ++i) g[i] += gl * (-st[i]);
    if (gin) for (int i = 0; i < D; ++i) gin[r * D + i] = g[i];
  }
  {
    T graw[Net::kPp] = {};
    first_grad_to_raw<T, Net::kK>(gfirst, sink.first_knots(), sc, graw);
    for (int j = 0; j < Net::kPp; ++j) G[j] += (double)graw[j];
  }
  return 0;
}

template <typename T, class Net>
static int step(int D, int L, const cnfot_problem_desc* pd, const T* W, const T* latent,
                const T* latent_sub, const T* src, const T* tgt, const double* t_batch, int n_t,
                int64_t rows_B, int64_t rows_b, int64_t gB, int64_t gb, double lambda, double* G,
                double* slots) {
  if ((L + 1) * D > kMaxStateFloats || D > kMaxDim) return 1;
  Dims<0, 0> dm{D, L};
  SplineConsts<T> sc = make_spline_consts<T>(Net::kK, -10.0, 10.0, 1e-4, 1e-4);
  StepConsts<T> pc;
  const char* err = nullptr;
  if (make_step_consts<T>(*pd, D, lambda, gB, gb, n_t, &pc, &err)) return 2;
  HostSink<T, Net> sink(G, W);
  HostTiles<T, Net> ht;
  RowTiles<T, Net> tl = ht.view();
  FirstGrad<T, Net::kK> gfirst = {};
  for (int s = 0; s < kNumSlots; ++s) slots[s] = 0.0;
  if (pd->type == CNFOT_OT) {
    for (int64_t r = 0; r < rows_B; ++r) {
      slots[kSlotFit0] += (double)row_nll<T, Net, Dims<0, 0>, HostSink<T, Net>>(
          dm, sc, (T)0, src + r * D, pc.w_fit, gfirst, tl, sink);
      slots[kSlotFitT] += (double)row_nll<T, Net, Dims<0, 0>, HostSink<T, Net>>(
          dm, sc, pc.horizon, tgt + r * D, pc.w_fit, gfirst, tl, sink);
    }
  } else {
    for (int64_t r = 0; r < rows_B; ++r) {
      T lf = 0, lp = 0;
      row_sample_terms<T, Net, Dims<0, 0>, HostSink<T, Net>>(dm, sc, (T)0, latent + r * D, true,
                                                             false, pc, &lf, &lp, gfirst, tl, sink);
      if (pd->type == CNFOT_RWPO)
        row_sample_terms<T, Net, Dims<0, 0>, HostSink<T, Net>>(dm, sc, pc.horizon, latent + r * D,
                                                               false, true, pc, &lf, &lp,
                                                               gfirst, tl, sink);
      slots[kSlotFit0] += (double)lf;
      slots[kSlotPotential] += (double)lp;
    }
  }
  for (int it = 0; it < n_t; ++it)
    for (int64_t r = 0; r < rows_b; ++r) {
      T lk = 0, lp = 0;
      row_kinetic<T, Net, Dims<0, 0>, HostSink<T, Net>>(dm, sc, (T)t_batch[it], latent_sub + r * D,
                                                        pc, &lk, &lp, gfirst, tl, sink);
      slots[kSlotKinetic] += (double)lk;
      slots[kSlotPotential] += (double)lp;
    }
  {
    T graw[Net::kPp] = {};
    first_grad_to_raw<T, Net::kK>(gfirst, sink.first_knots(), sc, graw);
    for (int j = 0; j < Net::kPp; ++j) G[j] += (double)graw[j];
  }
  return 0;
}

#define NET_CASE(H_, K_, M_) \
  if (H == H_ && K == K_ && M == M_) { using Net = NetCfg<H_, K_, M_>; return CALL; }

extern "C" int hs_layout(int D, int L, int M, int H, int K, int64_t* total, int64_t* Pp) {
  FlowLayout f = make_layout(D, L, M, H, K);
  *total = f.total;
  *Pp = f.Pp;
  return 0;
}
extern "C" int64_t hs_mlp_offset(int D, int H, int K, int M, int layer, int d) {
  FlowLayout f = make_layout(D, 1, M, H, K);
  return f.Pp + (int64_t)layer * f.layer_stride + (d - 1) * f.mlp_const + H * ((d - 1) * (d + 2) / 2);
}

extern "C" int hs_flow_eval_f64(int D, int L, int M, int H, int K, int dir, int64_t rows,
                                const double* W, const double* in, const double* cond, int64_t cs,
                                double* out, double* ld, int add_base) {
#define CALL flow_eval<double, Net>(D, L, dir, rows, W, in, cond, cs, out, ld, add_base)
  CNFOT_NET_LIST(NET_CASE)
#undef CALL
  return 3;
}
extern "C" int hs_flow_eval_f32(int D, int L, int M, int H, int K, int dir, int64_t rows,
                                const float* W, const float* in, const float* cond, int64_t cs,
                                float* out, float* ld, int add_base) {
#define CALL flow_eval<float, Net>(D, L, dir, rows, W, in, cond, cs, out, ld, add_base)
  CNFOT_NET_LIST(NET_CASE)
#undef CALL
  return 3;
}
extern "C" int hs_flow_vjp_f64(int D, int L, int M, int H, int K, int dir, int64_t rows,
                               const double* W, const double* in, const double* cond, int64_t cs,
                               const double* gout, const double* gld, int add_base, double* gin,
                               double* G) {
#define CALL flow_vjp<double, Net>(D, L, dir, rows, W, in, cond, cs, gout, gld, add_base, gin, G)
  CNFOT_NET_LIST(NET_CASE)
#undef CALL
  return 3;
}
extern "C" int hs_flow_vjp_f32(int D, int L, int M, int H, int K, int dir, int64_t rows,
                               const float* W, const float* in, const float* cond, int64_t cs,
                               const float* gout, const float* gld, int add_base, float* gin,
                               double* G) {
#define CALL flow_vjp<float, Net>(D, L, dir, rows, W, in, cond, cs, gout, gld, add_base, gin, G)
  CNFOT_NET_LIST(NET_CASE)
#undef CALL
  return 3;
}
extern "C" int hs_step_f64(int D, int L, int M, int H, int K, const cnfot_problem_desc* pd,
                           const double* W, const double* latent, const double* latent_sub,
                           const double* src, const double* tgt, const double* t_batch, int n_t,
                           int64_t rows_B, int64_t rows_b, int64_t gB, int64_t gb, double lambda,
                           double* G, double* slots) {
#define CALL step<double, Net>(D, L, pd, W, latent, latent_sub, src, tgt, t_batch, n_t, rows_B, rows_b, gB, gb, lambda, G, slots)
  CNFOT_NET_LIST(NET_CASE)
#undef CALL
  return 3;
}
extern "C" int hs_step_f32(int D, int L, int M, int H, int K, const cnfot_problem_desc* pd,
                           const float* W, const float* latent, const float* latent_sub,
                           const float* src, const float* tgt, const double* t_batch, int n_t,
                           int64_t rows_B, int64_t rows_b, int64_t gB, int64_t gb, double lambda,
                           double* G, double* slots) {
#define CALL step<float, Net>(D, L, pd, W, latent, latent_sub, src, tgt, t_batch, n_t, rows_B, rows_b, gB, gb, lambda, G, slots)
  CNFOT_NET_LIST(NET_CASE)
#undef CALL
  return 3;
}
