// Fused whole-flow kernels (one thread = one sample row, weights in shared
// memory) behind cnfot_flow_* and cnfot_mfc_step.  Templated on the network
// shape (hidden, bins, mlp layers) and on a runtime / compile-time flow shape.
#pragma once

#include <type_traits>

#include "device_common.cuh"
#include "step_math.cuh"
#include "tc_engine.cuh"
#include "warp_mlp.cuh"

namespace cnfot {

// CTA context by engine: 0 CUDA-core dense layers, 1 the tcgen05 engine (tc_engine.cuh),
// 2 the warp-level tensor-core engine (warp_mlp.cuh); 1 and 2 exist for 16-wide networks
enum Engine { kEngCuda = 0, kEngTc = 1, kEngMma = 2, kEngMmaStream = 3, kEngWide = 4 };
#ifndef CNFOT_MMA_MIN_CTAS
#define CNFOT_MMA_MIN_CTAS 4
#endif
constexpr int kMmaMinCtas = CNFOT_MMA_MIN_CTAS;   // register budget of the warp-MMA kernels: 65536 / (128 * n)
template <class Net, int ENG>
struct CtxSelect { using type = DeviceCtx<Net>; };
template <class Net>
struct CtxSelect<Net, kEngTc> { using type = DeviceCtxTC<Net>; };
template <class Net>
struct CtxSelect<Net, kEngMma> { using type = DeviceCtxMma<Net, true>; };
template <class Net>
struct CtxSelect<Net, kEngMmaStream> { using type = DeviceCtxMma<Net, false>; };

template <class Ctx>
__device__ __forceinline__ void ctx_setup(Ctx& ctx, int D, int L, uint64_t* mbar, uint32_t* slot) {
  ctx.setup(D, L, mbar, slot);
}
template <class Ctx>
__device__ __forceinline__ void ctx_teardown(Ctx& ctx) { ctx.teardown(); }

// ---- forward-only evaluation (model API: sample / forward / inverse / log_prob) ----
struct EvalArgs {
  const float* W;
  const float* frags;   // streamed warp-MMA plan: global fragment buffer (else nullptr)
  const float* in;
  const float* cond;
  int64_t cond_stride;
  int64_t rows;
  float* out;
  float* logdet;
  int dir;
  int add_base;
  int D, L;
  SmemPlan plan;
  SplineConsts<float> sc;
};

template <class Net, class DimsT, int ENG>
__global__ void __launch_bounds__(kTile, ENG >= kEngMma ? kMmaMinCtas : 1) flow_eval_kernel(EvalArgs a) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t tc_mbar;
  __shared__ uint32_t tc_slot;
  using Ctx = typename CtxSelect<Net, ENG>::type;
  Ctx ctx;
  ctx.smem = smem; ctx.gW = a.W; ctx.p = a.plan;
  ctx.bind_partials(nullptr);
  ctx.bind_frags(a.frags);
  ctx_setup(ctx, a.D, a.L, &tc_mbar, &tc_slot);
  const RowTiles<float, Net> tl = make_row_tiles<Net>(smem, a.plan);
  const DimsT dm{a.D, a.L};
  const int D = dm.D(), L = dm.L();
  for (int64_t tile = blockIdx.x; tile * kTile < a.rows; tile += gridDim.x) {
    const int64_t r = tile * kTile + ctx.row_in_tile();
    const bool live = r < a.rows;  // every thread runs the pass: the context has CTA barriers
    float st[kMaxStateFloats];
    for (int i = 0; i < D; ++i) st[i] = live ? a.in[r * D + i] : 0.f;
    const float t = live ? a.cond[r * a.cond_stride] : 0.f;
    float ld = a.dir == 0 ? flow_pass<0, float, Net, DimsT, Ctx>(dm, a.sc, t, st, tl, ctx)
                          : flow_pass<1, float, Net, DimsT, Ctx>(dm, a.sc, t, st, tl, ctx);
    if (!live) continue;
    for (int i = 0; i < D; ++i) a.out[r * D + i] = st[L * D + i];
    if (a.logdet) {
      if (a.add_base)
        ld = a.dir == 0 ? base_log_prob<float>(st, D) - ld
                        : base_log_prob<float>(st + L * D, D) + ld;
      a.logdet[r] = ld;
    }
  }
  ctx_teardown(ctx);
}

// ---- per-CTA partial results -------------------------------------------------------
struct PartialBuf {
  float* grad;     // [n_cta][total]
  double* loss;    // [n_cta][kNumSlots]
};

// sAcc: the CTA's shared-memory gradient accumulator, or nullptr when the context adds straight
// into the CTA's partial row (Ctx::kAccInGlobal)
__device__ inline void flush_partials(const PartialBuf& pb, const float* sAcc, int total,
                                      const double* loss /* kNumSlots, thread-local */,
                                      double* scratch) {
  __syncthreads();
  float* dst = pb.grad + (int64_t)blockIdx.x * total;
  if (sAcc)
    for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = sAcc[i];
  for (int s = 0; s < kNumSlots; ++s) {
    double v = block_sum(loss[s], scratch);
    if (threadIdx.x == 0) pb.loss[(int64_t)blockIdx.x * kNumSlots + s] = v;
  }
}

// ---- generic VJP of one flow pass (what a custom_vjp backward rule calls) ----------
struct VjpArgs {
  const float* W;
  const float* frags;
  const float* in;
  const float* cond;
  int64_t cond_stride;
  int64_t rows;
  const float* g_out;
  const float* g_logdet;
  float* g_in;
  int dir;
  int add_base;
  int D, L;
  SmemPlan plan;
  SplineConsts<float> sc;
  PartialBuf pb;
};

template <class Net, class DimsT, int ENG>
__global__ void __launch_bounds__(kTile, ENG >= kEngMma ? kMmaMinCtas : 1) flow_vjp_kernel(VjpArgs a) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ double scratch[kWarps];
  __shared__ __align__(8) uint64_t tc_mbar;
  __shared__ uint32_t tc_slot;
  using Ctx = typename CtxSelect<Net, ENG>::type;
  float* sAcc = Ctx::kAccInGlobal ? nullptr : smem + a.plan.off_acc;
  Ctx ctx;
  ctx.smem = smem; ctx.gW = a.W; ctx.p = a.plan;
  ctx.bind_partials(a.pb.grad + (int64_t)blockIdx.x * a.plan.total);
  ctx.bind_frags(a.frags);
  ctx_setup(ctx, a.D, a.L, &tc_mbar, &tc_slot);
  const RowTiles<float, Net> tl = make_row_tiles<Net>(smem, a.plan);
  const DimsT dm{a.D, a.L};
  const int D = dm.D(), L = dm.L();
  float gfirst[Net::kPp];
#pragma unroll
  for (int j = 0; j < Net::kPp; ++j) gfirst[j] = 0.f;
  for (int64_t tile = blockIdx.x; tile * kTile < a.rows; tile += gridDim.x) {
    const int64_t r = tile * kTile + ctx.row_in_tile();
    const bool live = r < a.rows;
    float st[kMaxStateFloats], g[kMaxDim];
    for (int i = 0; i < D; ++i) st[i] = live ? a.in[r * D + i] : 0.f;
    const float t = live ? a.cond[r * a.cond_stride] : 0.f;
    if (a.dir == 0) flow_pass<0, float, Net, DimsT, Ctx>(dm, a.sc, t, st, tl, ctx);
    else flow_pass<1, float, Net, DimsT, Ctx>(dm, a.sc, t, st, tl, ctx);
    const float gl = (live && a.g_logdet) ? a.g_logdet[r] : 0.f;
    for (int i = 0; i < D; ++i) g[i] = live ? a.g_out[r * D + i] : 0.f;
    float gl_pass = gl;
    if (a.add_base) {
      if (a.dir == 0) gl_pass = -gl;
      else
        for (int i = 0; i < D; ++i) g[i] -= gl * st[L * D + i];
    }
    if (a.dir == 0)
      flow_pass_bwd<0, float, Net, DimsT, Ctx>(dm, a.sc, t, st, g, gl_pass, gfirst, tl, ctx);
    else
      flow_pass_bwd<1, float, Net, DimsT, Ctx>(dm, a.sc, t, st, g, gl_pass, gfirst, tl, ctx);
    if (a.add_base && a.dir == 0)
      for (int i = 0; i < D; ++i) g[i] -= gl * st[i];
    if (live && a.g_in)
      for (int i = 0; i < D; ++i) a.g_in[r * D + i] = g[i];
  }
  ctx.flush_first(gfirst, tl);
  double zero[kNumSlots];
  for (int s = 0; s < kNumSlots; ++s) zero[s] = 0.0;
  flush_partials(a.pb, sAcc, a.plan.total, zero, scratch);
  ctx_teardown(ctx);
}

// ---- evaluation energies: kinetic energy over a time grid (forward only) --------------------
// tile = (time index, 128-row tile); latent block (ti % latent_blocks) feeds time ti.
struct EnergyArgs {
  const float* W;
  const float* frags;
  int D, L;
  SmemPlan plan;
  SplineConsts<float> sc;
  const float* latent;     // (latent_blocks * batch, D)
  int64_t batch;
  int latent_blocks;
  const float* t_dev;      // (n_t) times, device
  int n_t;
  int with_score;
  float dt, dx, kappa;
  double weight;           // 1 / (2 batch n_t)
  int64_t tiles_per_t, n_tiles;
  unsigned long long* tile_counter;
  PartialBuf pb;
};

template <class Net, class DimsT, int ENG>
__global__ void __launch_bounds__(kTile, ENG >= kEngMma ? kMmaMinCtas : 1) energy_kernel(const __grid_constant__ EnergyArgs a) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ double scratch[kWarps];
  __shared__ long long s_tile;
  __shared__ __align__(8) uint64_t tc_mbar;
  __shared__ uint32_t tc_slot;
  using Ctx = typename CtxSelect<Net, ENG>::type;
  Ctx ctx;
  ctx.smem = smem; ctx.gW = a.W; ctx.p = a.plan;
  ctx.bind_partials(nullptr);
  ctx.bind_frags(a.frags);
  ctx_setup(ctx, a.D, a.L, &tc_mbar, &tc_slot);
  const RowTiles<float, Net> tl = make_row_tiles<Net>(smem, a.plan);
  const DimsT dm{a.D, a.L};
  const int D = dm.D();
  double loss[kNumSlots];
#pragma unroll
  for (int s = 0; s < kNumSlots; ++s) loss[s] = 0.0;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(a.tile_counter, 1ULL);
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= a.n_tiles) break;
    const int ti = (int)(tile / a.tiles_per_t);
    const int64_t r = (tile - (long long)ti * a.tiles_per_t) * kTile + ctx.row_in_tile();
    const bool live = r < a.batch;
    const float* src = a.latent + ((int64_t)(ti % a.latent_blocks) * a.batch + r) * D;
    float row[kMaxDim];
    for (int i = 0; i < D; ++i) row[i] = live ? src[i] : 0.f;
    const float v2 = row_kinetic_value<float, Net, DimsT, Ctx>(dm, a.sc, a.t_dev[ti], row, a.dt, a.with_score != 0,
                                                               a.kappa, a.dx, tl, ctx);
    if (live) loss[kSlotKinetic] += (double)v2 * a.weight;
  }
  flush_partials(a.pb, nullptr, 0, loss, scratch);
  ctx_teardown(ctx);
}

// ---- the fused train step ---------------------------------------------------------------
// One persistent kernel evaluates every term of the configured loss: the work is a list
// of segments (one per loss term and time), cut into 128-row tiles handed out by an
// atomic counter, most expensive segments first.
enum SegmentKind { kSegNll = 0, kSegSample = 1, kSegKinetic = 2 };

struct Segment {
  int kind;
  int slot;          // loss slot of the fit term (kSegNll / kSegSample)
  int do_fit, do_pot;
  float t;
  const float* rows; // (n, D) data or latent rows
  int64_t n;
  int64_t first_tile;
};

constexpr int kMaxSegments = 40;

struct StepArgs {
  const float* W;
  const float* frags;
  int D, L;
  SmemPlan plan;
  SplineConsts<float> sc;
  StepConsts<float> pc;
  int n_seg;
  int64_t n_tiles;
  Segment seg[kMaxSegments];
  unsigned long long* tile_counter;
  PartialBuf pb;
};

template <class Net, class DimsT, int ENG>
__global__ void __launch_bounds__(kTile, ENG >= kEngMma ? kMmaMinCtas : 1) mfc_step_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ double scratch[kWarps];
  __shared__ long long s_tile;
  __shared__ __align__(8) uint64_t tc_mbar;
  __shared__ uint32_t tc_slot;
  using Ctx = typename CtxSelect<Net, ENG>::type;
  float* sAcc = Ctx::kAccInGlobal ? nullptr : smem + a.plan.off_acc;
  Ctx ctx;
  ctx.smem = smem; ctx.gW = a.W; ctx.p = a.plan;
  ctx.bind_partials(a.pb.grad + (int64_t)blockIdx.x * a.plan.total);
  ctx.bind_frags(a.frags);
  ctx_setup(ctx, a.D, a.L, &tc_mbar, &tc_slot);
  const RowTiles<float, Net> tl = make_row_tiles<Net>(smem, a.plan);
  const DimsT dm{a.D, a.L};
  const int D = dm.D();
  float gfirst[Net::kPp];
#pragma unroll
  for (int j = 0; j < Net::kPp; ++j) gfirst[j] = 0.f;
  double loss[kNumSlots];
#pragma unroll
  for (int s = 0; s < kNumSlots; ++s) loss[s] = 0.0;

  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(a.tile_counter, 1ULL);
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= a.n_tiles) break;
    int si = 0;
    while (si + 1 < a.n_seg && tile >= a.seg[si + 1].first_tile) ++si;
    const Segment& sg = a.seg[si];
    const int64_t r = (tile - sg.first_tile) * kTile + ctx.row_in_tile();
    const bool live = r < sg.n;
    float row[kMaxDim];
    for (int i = 0; i < D; ++i) row[i] = live ? sg.rows[r * D + i] : 0.f;
    if (sg.kind == kSegNll) {
      float v = row_nll<float, Net, DimsT, Ctx>(dm, a.sc, sg.t, row,
                                                           live ? a.pc.w_fit : 0.f, gfirst, tl, ctx);
      loss[sg.slot] += (double)v;
    } else if (sg.kind == kSegSample) {
      StepConsts<float> pc = a.pc;
      if (!live) { pc.w_fit = 0.f; pc.w_pot = 0.f; }
      float lf = 0.f, lp = 0.f;
      row_sample_terms<float, Net, DimsT, Ctx>(dm, a.sc, sg.t, row, sg.do_fit != 0,
                                                          sg.do_pot != 0, pc, &lf, &lp, gfirst, tl, ctx);
      loss[sg.slot] += (double)lf;
      loss[kSlotPotential] += (double)lp;
    } else {
      StepConsts<float> pc = a.pc;
      if (!live) { pc.w_kin = 0.f; pc.w_pot = 0.f; }
      float lk = 0.f, lp = 0.f;
      row_kinetic<float, Net, DimsT, Ctx>(dm, a.sc, sg.t, row, pc, &lk, &lp, gfirst,
                                                     tl, ctx);
      loss[kSlotKinetic] += (double)lk;
      loss[kSlotPotential] += (double)lp;
    }
  }
  ctx.flush_first(gfirst, tl);
  flush_partials(a.pb, sAcc, a.plan.total, loss, scratch);
  ctx_teardown(ctx);
}

}  // namespace cnfot
