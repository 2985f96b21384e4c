// XLA FFI shim: every entry point of include/cnfot.h a JAX port of the reference needs, as XLA custom-call handlers.
//
// The reference (`/root/reference`, pure JAX) has no FFI of its own; the seams this plugs into are its ordinary
// Python callables: `bijector_fn` (cnf_ot/models/flows.py:124-132), the `Flow` namedtuple
// (flows.py:213-226) and `jax.value_and_grad(loss_fn)` + `optax.adam` inside `update` (cnf_ot/mfc/solvers.py:90-97).
// Each handler forwards (stream, device buffers, scalar attributes) 1:1 to the C ABI: XLA owns every buffer
// (scratch memory is an extra result buffer sized with the matching *_workspace_bytes()), the library never
// synchronises the stream, errors come back as XLA_FFI_Error.  cnf_ot_b200/jax_ffi.py registers the symbols
// (`jax.ffi.register_ffi_target`) and wraps them in `jax.custom_vjp`.
//
// Build (needs jaxlib's headers, which this image does not have: `python -c "import jax.ffi; print(jax.ffi.include_dir())"`):
//   g++ -O2 -std=c++17 -shared -fPIC -I$JAX_FFI_INCLUDE -I/usr/local/cuda/include -Iinclude \
//       cnf_ot_b200/csrc/xla_ffi_shim.cc -o cnf_ot_b200/libcnfot_xla.so -Lcnf_ot_b200 -lcnfot
// Without the header the file compiles to nothing.  tests/test_ffi_shim.py compiles it against a stand-in of the
// part of the public `xla/ffi/api/ffi.h` API it uses (tests/xla_ffi_standin), which type-checks every handler
// signature against its binding.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define CNFOT_HAVE_XLA_FFI 1
#endif
#endif

#if defined(CNFOT_HAVE_XLA_FFI)

#include <cstdint>

#include "cnfot.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
typedef struct CUstream_st* cudaStream_t;

using F32 = ffi::Buffer<ffi::F32>;
using U8 = ffi::Buffer<ffi::U8>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

static ffi::Error status(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, cnfot_last_error());
}

// RQSFlow(event_shape=(dim,), num_layers, hidden_sizes=[hidden]*mlp_layers, num_bins) with the spline constants the
// reference hard-codes (flows.py:124-132)
static cnfot_flow_desc flow_desc(int64_t dim, int64_t num_layers, int64_t mlp_layers, int64_t hidden, int64_t num_bins) {
  return cnfot_flow_desc{(int32_t)dim, (int32_t)num_layers, (int32_t)mlp_layers, (int32_t)hidden, (int32_t)num_bins,
                         -10.f, 10.f, 1e-4f, 1e-4f};
}
#define CNFOT_FLOW_ATTRS \
  .Attr<int64_t>("dim").Attr<int64_t>("num_layers").Attr<int64_t>("mlp_layers").Attr<int64_t>("hidden").Attr<int64_t>("num_bins")
#define CNFOT_FLOW_PARAMS int64_t dim, int64_t num_layers, int64_t mlp_layers, int64_t hidden, int64_t num_bins
#define CNFOT_FLOW_DESC flow_desc(dim, num_layers, mlp_layers, hidden, num_bins)

// config/mfc.yaml: general.type, the sub-type, rwpo.T / fp.T, rwpo.beta, a, fp.sigma, general.dt, general.dx
static cnfot_problem_desc problem_desc(int64_t type, int64_t subtype, float T, float beta, float a, float sigma, float dt,
                                       float dx) {
  return cnfot_problem_desc{(int32_t)type, (int32_t)subtype, T, beta, a, sigma, dt, dx};
}
#define CNFOT_PROBLEM_ATTRS                                                                                          \
  .Attr<int64_t>("type").Attr<int64_t>("subtype").Attr<float>("T").Attr<float>("beta").Attr<float>("a").Attr<float>( \
      "sigma").Attr<float>("dt").Attr<float>("dx")
#define CNFOT_PROBLEM_PARAMS int64_t type, int64_t subtype, float T, float beta, float a, float sigma, float dt, float dx
#define CNFOT_PROBLEM_DESC problem_desc(type, subtype, T, beta, a, sigma, dt, dx)

// ---- seam 2: distrax.RationalQuadraticSpline(params).forward_and_log_det / inverse_and_log_det -----------------
template <bool INVERSE>
static ffi::Error RqsImpl(cudaStream_t stream, F32 v, F32 params, int64_t num_bins, RF32 out, RF32 logdet, RS32 bin) {
  const int64_t rows = (int64_t)v.element_count();
  auto fn = INVERSE ? cnfot_rqs_inverse : cnfot_rqs_forward;
  return status(fn(stream, v.typed_data(), params.typed_data(), rows, (int32_t)num_bins, -10.f, 10.f, 1e-4f, 1e-4f,
                   out->typed_data(), logdet->typed_data(), bin->typed_data()));
}
template <bool INVERSE>
static ffi::Error RqsVjpImpl(cudaStream_t stream, F32 v, F32 params, F32 g_out, F32 g_logdet, int64_t num_bins, RF32 g_in,
                             RF32 g_params) {
  const int64_t rows = (int64_t)v.element_count();
  auto fn = INVERSE ? cnfot_rqs_inverse_vjp : cnfot_rqs_forward_vjp;
  return status(fn(stream, v.typed_data(), params.typed_data(), g_out.typed_data(), g_logdet.typed_data(), rows,
                   (int32_t)num_bins, -10.f, 10.f, 1e-4f, 1e-4f, g_in->typed_data(), g_params->typed_data()));
}
#define CNFOT_RQS_BINDING                                                                                    \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Attr<int64_t>("num_bins") \
      .Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::S32>>()
#define CNFOT_RQS_VJP_BINDING                                                                               \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()    \
      .Attr<int64_t>("num_bins").Ret<F32>().Ret<F32>()
XLA_FFI_DEFINE_HANDLER_SYMBOL(CnfotRqsForward, RqsImpl<false>, CNFOT_RQS_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(CnfotRqsInverse, RqsImpl<true>, CNFOT_RQS_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(CnfotRqsForwardVjp, RqsVjpImpl<false>, CNFOT_RQS_VJP_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(CnfotRqsInverseVjp, RqsVjpImpl<true>, CNFOT_RQS_VJP_BINDING);

// ---- seam 1: the Flow namedtuple -------------------------------------------------------------------------------
//   forward : flow.forward / sample / sample_and_log_prob (add_base = 1)   inverse : flow.inverse / log_prob (add_base = 1)
// cond holds one time (broadcast, like the (1,) `cond` of log_prob) or one per row.  `scratch` is sized with
// cnfot_flow_workspace_bytes (0 is fine for flows the fused kernels cover: pass a 1-byte buffer).
template <bool INVERSE>
static ffi::Error FlowImpl(cudaStream_t stream, F32 weights, F32 x, F32 cond, CNFOT_FLOW_PARAMS, int64_t add_base, RF32 y,
                           RF32 logdet, RU8 scratch) {
  const cnfot_flow_desc d = CNFOT_FLOW_DESC;
  const int64_t rows = (int64_t)x.element_count() / dim;
  const int64_t cond_stride = cond.element_count() == 1 ? 0 : 1;
  auto fn = INVERSE ? cnfot_flow_inverse_ws : cnfot_flow_forward_ws;
  return status(fn(stream, &d, weights.typed_data(), x.typed_data(), cond.typed_data(), cond_stride, rows, y->typed_data(),
                   logdet->typed_data(), (int32_t)add_base, scratch->typed_data(), (int64_t)scratch->element_count()));
}
template <bool INVERSE>
static ffi::Error FlowVjpImpl(cudaStream_t stream, F32 weights, F32 x, F32 cond, F32 g_out, F32 g_logdet, CNFOT_FLOW_PARAMS,
                              int64_t add_base, RF32 g_x, RF32 g_weights, RU8 scratch) {
  const cnfot_flow_desc d = CNFOT_FLOW_DESC;
  const int64_t rows = (int64_t)x.element_count() / dim;
  const int64_t cond_stride = cond.element_count() == 1 ? 0 : 1;
  auto fn = INVERSE ? cnfot_flow_inverse_vjp : cnfot_flow_forward_vjp;
  return status(fn(stream, &d, weights.typed_data(), x.typed_data(), cond.typed_data(), cond_stride, rows,
                   g_out.typed_data(), g_logdet.typed_data(), (int32_t)add_base, g_x->typed_data(),
                   g_weights->typed_data(), scratch->typed_data(), (int64_t)scratch->element_count()));
}
#define CNFOT_FLOW_BINDING                                                                                 \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>() CNFOT_FLOW_ATTRS \
      .Attr<int64_t>("add_base").Ret<F32>().Ret<F32>().Ret<U8>()
#define CNFOT_FLOW_VJP_BINDING                                                                                        \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()   \
      CNFOT_FLOW_ATTRS.Attr<int64_t>("add_base").Ret<F32>().Ret<F32>().Ret<U8>()
XLA_FFI_DEFINE_HANDLER_SYMBOL(CnfotFlowForward, FlowImpl<false>, CNFOT_FLOW_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(CnfotFlowInverse, FlowImpl<true>, CNFOT_FLOW_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(CnfotFlowForwardVjp, FlowVjpImpl<false>, CNFOT_FLOW_VJP_BINDING);
XLA_FFI_DEFINE_HANDLER_SYMBOL(CnfotFlowInverseVjp, FlowVjpImpl<true>, CNFOT_FLOW_VJP_BINDING);

// ---- seam 3: jax.value_and_grad(loss_fn)(params, rng, _lambda, batch_size), solvers.py:94 ---------------------
// Explicit draws: latent (B, D) [rwpo / fp], latent_sub (b, D), src / tgt (B, D) [ot]; unused ones are passed as
// 0-element buffers.  t_batch: the uniform times, a host-side attribute.  out: [gradient | 8 loss slots].
static const float* opt(const F32& b) { return b.element_count() ? b.typed_data() : nullptr; }
static ffi::Error MfcStepImpl(cudaStream_t stream, F32 weights, F32 latent, F32 latent_sub, F32 src, F32 tgt,
                              CNFOT_FLOW_PARAMS, CNFOT_PROBLEM_PARAMS, ffi::Span<const float> t_batch, int64_t global_B,
                              int64_t global_b, float lambda, RF32 out, RU8 scratch) {
  const cnfot_flow_desc d = CNFOT_FLOW_DESC;
  const cnfot_problem_desc p = CNFOT_PROBLEM_DESC;
  const int64_t rows_B = (int64_t)(type == CNFOT_OT ? src.element_count() : latent.element_count()) / dim;
  const int64_t rows_b = (int64_t)latent_sub.element_count() / dim;
  return status(cnfot_mfc_step(stream, &d, &p, weights.typed_data(), opt(latent), opt(latent_sub), opt(src), opt(tgt),
                               t_batch.begin(), (int32_t)t_batch.size(), rows_B, rows_b, global_B, global_b, lambda,
                               out->typed_data(), scratch->typed_data(), (int64_t)scratch->element_count()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    CnfotMfcStep, MfcStepImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        CNFOT_FLOW_ATTRS CNFOT_PROBLEM_ATTRS.Attr<ffi::Span<const float>>("t_batch").Attr<int64_t>("global_B")
        .Attr<int64_t>("global_b").Attr<float>("lambda").Ret<F32>().Ret<U8>());

// The same with the draws made inside the kernel from (key, step): what the reference's `update` does with its `rng`.
// [row0_B, row0_B + rows_B) / [row0_b, row0_b + rows_b): this device's shard of the global batch.
static ffi::Error MfcStepRngImpl(cudaStream_t stream, F32 weights, CNFOT_FLOW_PARAMS, CNFOT_PROBLEM_PARAMS, int64_t key,
                                 int64_t step, int64_t n_t, int64_t row0_B, int64_t rows_B, int64_t row0_b, int64_t rows_b,
                                 int64_t global_B, int64_t global_b, float lambda, RF32 out, RU8 scratch) {
  const cnfot_flow_desc d = CNFOT_FLOW_DESC;
  const cnfot_problem_desc p = CNFOT_PROBLEM_DESC;
  return status(cnfot_mfc_step_rng(stream, &d, &p, weights.typed_data(), (uint64_t)key, (uint32_t)step, (int32_t)n_t, row0_B,
                                   rows_B, row0_b, rows_b, global_B, global_b, lambda, out->typed_data(),
                                   scratch->typed_data(), (int64_t)scratch->element_count(), nullptr));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    CnfotMfcStepRng, MfcStepRngImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>() CNFOT_FLOW_ATTRS CNFOT_PROBLEM_ATTRS
        .Attr<int64_t>("key").Attr<int64_t>("step").Attr<int64_t>("n_t").Attr<int64_t>("row0_B").Attr<int64_t>("rows_B")
        .Attr<int64_t>("row0_b").Attr<int64_t>("rows_b").Attr<int64_t>("global_B").Attr<int64_t>("global_b")
        .Attr<float>("lambda").Ret<F32>().Ret<U8>());

// ---- `update` itself (solvers.py:90-97) as ONE kernel: draws + value_and_grad + Adam.  state / weights / moments are
// updated in place: bind the call with input_output_aliases {0: 0, 1: 1, 2: 2, 3: 3} (jax_ffi.py).  The train state is
// created once with cnfot_train_state_init (host call, jax_ffi.train_state_init).
static ffi::Error MfcUpdateImpl(cudaStream_t stream, U8 state_in, F32 weights_in, F32 m_in, F32 v_in, CNFOT_FLOW_PARAMS,
                                CNFOT_PROBLEM_PARAMS, int64_t n_t, int64_t row0_B, int64_t rows_B, int64_t row0_b,
                                int64_t rows_b, int64_t global_B, int64_t global_b, float lambda, float lr, float b1, float b2,
                                float eps, RU8 state, RF32 weights, RF32 m, RF32 v, RF32 out) {
  if (state->untyped_data() != state_in.untyped_data() || weights->untyped_data() != weights_in.untyped_data() ||
      m->untyped_data() != m_in.untyped_data() || v->untyped_data() != v_in.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "CnfotMfcUpdate updates in place: alias inputs 0-3 to outputs 0-3");
  const cnfot_flow_desc d = CNFOT_FLOW_DESC;
  const cnfot_problem_desc p = CNFOT_PROBLEM_DESC;
  const cnfot_adam_desc adam{lr, b1, b2, eps};
  return status(cnfot_mfc_update(stream, &d, &p, state->typed_data(), (int64_t)state->element_count(), weights->typed_data(),
                                 m->typed_data(), v->typed_data(), &adam, (int32_t)n_t, row0_B, rows_B, row0_b, rows_b,
                                 global_B, global_b, lambda, out->typed_data(), nullptr, 0, nullptr));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    CnfotMfcUpdate, MfcUpdateImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<U8>().Arg<F32>().Arg<F32>().Arg<F32>() CNFOT_FLOW_ATTRS
        CNFOT_PROBLEM_ATTRS.Attr<int64_t>("n_t").Attr<int64_t>("row0_B").Attr<int64_t>("rows_B").Attr<int64_t>("row0_b")
        .Attr<int64_t>("rows_b").Attr<int64_t>("global_B").Attr<int64_t>("global_b").Attr<float>("lambda").Attr<float>("lr")
        .Attr<float>("b1").Attr<float>("b2").Attr<float>("eps").Ret<U8>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

// optimizer.update + optax.apply_updates (solvers.py:95-96) on a gradient that came from elsewhere; in place (aliases 0, 2, 3)
static ffi::Error AdamImpl(cudaStream_t stream, F32 params_in, F32 grads, F32 m_in, F32 v_in, float lr, float b1, float b2,
                           float eps, int64_t step, RF32 params, RF32 m, RF32 v) {
  if (params->untyped_data() != params_in.untyped_data() || m->untyped_data() != m_in.untyped_data() ||
      v->untyped_data() != v_in.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "CnfotAdam updates in place: alias inputs 0, 2, 3 to outputs 0, 1, 2");
  return status(cnfot_adam_update(stream, params->typed_data(), grads.typed_data(), m->typed_data(), v->typed_data(),
                                  (int64_t)grads.element_count(), lr, b1, b2, eps, step));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CnfotAdam, AdamImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Attr<float>("lr").Attr<float>("b1").Attr<float>("b2").Attr<float>("eps")
                                  .Attr<int64_t>("step").Ret<F32>().Ret<F32>().Ret<F32>());

// ---- evaluation: utils.calc_kinetic_energy / calc_score_kinetic_energy (cnf_ot/utils.py:311-389) ----------------
static ffi::Error KineticEnergyImpl(cudaStream_t stream, F32 weights, F32 latent, CNFOT_FLOW_PARAMS,
                                    ffi::Span<const float> t_values, int64_t latent_blocks, float dt, int64_t with_score,
                                    float kappa, float dx, RF64 out, RU8 scratch) {
  const cnfot_flow_desc d = CNFOT_FLOW_DESC;
  const int64_t batch = (int64_t)latent.element_count() / dim / latent_blocks;
  return status(cnfot_kinetic_energy(stream, &d, weights.typed_data(), latent.typed_data(), batch, (int32_t)latent_blocks,
                                     t_values.begin(), (int32_t)t_values.size(), dt, (int32_t)with_score, kappa, dx,
                                     out->typed_data(), scratch->typed_data(), (int64_t)scratch->element_count()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    CnfotKineticEnergy, KineticEnergyImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>() CNFOT_FLOW_ATTRS
        .Attr<ffi::Span<const float>>("t_values").Attr<int64_t>("latent_blocks").Attr<float>("dt").Attr<int64_t>("with_score")
        .Attr<float>("kappa").Attr<float>("dx").Ret<ffi::Buffer<ffi::F64>>().Ret<U8>());

// ---- evaluation: densities on a grid (utils.py:572-642, solvers.py:184-222,282-301) and at Monte-Carlo samples
// (solvers.py:254-278).  density is (n_t, ny, nx) / (n); sq_err one double (0 when with_ref == 0).
static ffi::Error DensityGridImpl(cudaStream_t stream, F32 weights, CNFOT_FLOW_PARAMS, ffi::Span<const float> t_values,
                                  float x_min, float x_max, float y_min, float y_max, int64_t nx, int64_t ny,
                                  int64_t with_ref, float mix, float var0, float var1, RF32 density, RF64 sq_err,
                                  RU8 scratch) {
  const cnfot_flow_desc d = CNFOT_FLOW_DESC;
  return status(cnfot_density_grid(stream, &d, weights.typed_data(), t_values.begin(), (int32_t)t_values.size(), x_min, x_max,
                                   y_min, y_max, (int32_t)nx, (int32_t)ny, density->typed_data(), (int32_t)with_ref, mix,
                                   var0, var1, sq_err->typed_data(), scratch->typed_data(),
                                   (int64_t)scratch->element_count()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    CnfotDensityGrid, DensityGridImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>() CNFOT_FLOW_ATTRS
        .Attr<ffi::Span<const float>>("t_values").Attr<float>("x_min").Attr<float>("x_max").Attr<float>("y_min")
        .Attr<float>("y_max").Attr<int64_t>("nx").Attr<int64_t>("ny").Attr<int64_t>("with_ref").Attr<float>("mix")
        .Attr<float>("var0").Attr<float>("var1").Ret<F32>().Ret<ffi::Buffer<ffi::F64>>().Ret<U8>());

static ffi::Error DensityMcImpl(cudaStream_t stream, F32 weights, CNFOT_FLOW_PARAMS, float cond, int64_t key, int64_t step,
                                int64_t with_ref, float mix, float var0, float var1, RF32 samples, RF32 density, RF64 sq_err,
                                RU8 scratch) {
  const cnfot_flow_desc d = CNFOT_FLOW_DESC;
  const int64_t n = (int64_t)density->element_count();
  return status(cnfot_density_mc(stream, &d, weights.typed_data(), cond, (uint64_t)key, (uint32_t)step, n,
                                 samples->typed_data(), density->typed_data(), (int32_t)with_ref, mix, var0, var1,
                                 sq_err->typed_data(), scratch->typed_data(), (int64_t)scratch->element_count()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    CnfotDensityMc, DensityMcImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>() CNFOT_FLOW_ATTRS.Attr<float>("cond")
        .Attr<int64_t>("key").Attr<int64_t>("step").Attr<int64_t>("with_ref").Attr<float>("mix").Attr<float>("var0")
        .Attr<float>("var1").Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::F64>>().Ret<U8>());

#endif  // CNFOT_HAVE_XLA_FFI
