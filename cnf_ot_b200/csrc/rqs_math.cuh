// Per-element rational-quadratic-spline math, shared by every kernel.
//
// Mirrors the spline the reference instantiates at
// /root/reference/cnf_ot/models/flows.py:124-132 (distrax
// RationalQuadraticSpline, 'unconstrained' boundary slopes, linear tails) and
// evaluates at /root/reference/cnf_ot/models/autoregressive.py:100,130.
// distrax is not vendored in the reference; the algorithm is the published one
// (SURVEY.md Appendix B).  Everything here is register-resident: knot
// normalisation (two softmaxes, cumulative sums, softplus slopes), the bin
// search, the forward / inverse map, the log-det and their reverse-mode
// adjoints.  Functions are templated on the scalar type so the same source is
// compiled for float on the device and for float/double in the host-side
// test harness (tests/hostsim), which checks the adjoints against autograd.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define CNFOT_HD __host__ __device__ __forceinline__
// real function calls on the device: used for the big per-pass routines that are
// invoked from several places, to bound code size and compile time
#define CNFOT_CALL __host__ __device__ __noinline__
#else
#define CNFOT_HD inline
#define CNFOT_CALL inline
#endif

// Address-space hint: inside device code the pointer is known to address shared memory,
// so loads/stores through it compile to LDS/STS instead of generic LD/ST.
#if defined(__CUDA_ARCH__)
#define CNFOT_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#define CNFOT_ASSUME_LOCAL(p) __builtin_assume(__isLocal(p))
#else
#define CNFOT_ASSUME_SHARED(p) ((void)0)
#define CNFOT_ASSUME_LOCAL(p) ((void)0)
#endif

namespace cnfot {

template <typename T>
struct SplineConsts {
  T lo, hi;        // range_min, range_max
  T min_bin;       // min_bin_size
  T bin_scale;     // (hi - lo) - K * min_bin
  T min_slope;     // min_knot_slope
  T slope_offset;  // log(exp(1 - min_slope) - 1)
};

template <typename T>
inline SplineConsts<T> make_spline_consts(int K, double lo, double hi, double min_bin,
                                          double min_slope) {
  SplineConsts<T> c;
  c.lo = (T)lo;
  c.hi = (T)hi;
  c.min_bin = (T)min_bin;
  c.bin_scale = (T)((hi - lo) - K * min_bin);
  c.min_slope = (T)min_slope;
  c.slope_offset = (T)log(exp(1.0 - min_slope) - 1.0);
  return c;
}

// The constants the reference hard-codes at /root/reference/cnf_ot/models/flows.py:124-132 (range -10 .. 10,
// min_knot_slope 1e-4; min_bin_size 1e-4 is the distrax default) as COMPILE-TIME values: every use becomes an
// immediate operand.  The fused flow kernels use this type (a SplineConsts passed by reference to their
// non-inlined per-pass routines costs a generic load plus descriptor moves per use: 2.7 % of the step kernel's
// instructions, profiles/r02_*); the stand-alone spline kernels and the host harness keep the run-time struct.
template <typename T, int K>
struct FixedSplineConsts {
  static constexpr T lo = (T)-10, hi = (T)10;
  static constexpr T min_bin = (T)1e-4;
  static constexpr T bin_scale = (T)(20.0 - K * 1e-4);
  static constexpr T min_slope = (T)1e-4;
  static constexpr T slope_offset = (T)0.5411666523385311;   // log(exp(1 - 1e-4) - 1)
};
// ---- scalar helpers (overloaded so float uses the f-suffixed device paths) --
// On the device the float versions of exp / log / divide / reciprocal map to the SFU
// approximations (ex2.approx, lg2.approx, rcp.approx: <= 2-3 ulp), which cuts the spline from
// ~250 to ~120 instructions; the parity tests bound the effect (it is below the float32
// rounding already present in the knot positions).  -DCNFOT_PRECISE_MATH restores libm paths.
#if defined(__CUDA_ARCH__) && !defined(CNFOT_PRECISE_MATH)
#define CNFOT_FAST_MATH 1
// One SFU instruction each (the __expf / __logf / __fdividef intrinsics wrap the same instructions in denormal
// scaling and range checks: 5 instructions per call, 25 % of the stand-alone spline kernel in ncu's source view).
// Arguments here are normal numbers: bin sizes >= 1e-4, softmax sums >= 1, slopes >= 1e-4; results below 2^-126
// flush to zero, where the exact value changes nothing at float32 precision.
CNFOT_HD float m_ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
CNFOT_HD float m_lg2(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
CNFOT_HD float m_rcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
CNFOT_HD float m_exp(float v) { return m_ex2(v * 1.4426950408889634f); }
CNFOT_HD float m_log(float v) { return m_lg2(v) * 0.6931471805599453f; }
CNFOT_HD float m_div(float a, float b) { return a * m_rcp(b); }
// exp(u - m) as one FFMA + one SFU instruction: nm = -m log2(e) is computed once per softmax
CNFOT_HD float m_exp_shift_prep(float m) { return -m * 1.4426950408889634f; }
CNFOT_HD float m_exp_shifted(float u, float nm) { return m_ex2(fmaf(u, 1.4426950408889634f, nm)); }
// log(1 + e) for e in [0, 1] (the only use: softplus): e + e^2 Q(e), Q from a degree-8 least-squares fit of log1p(e) / e
// on Chebyshev nodes (fit error 3e-8); the last FFMA rounds the result once: 1.4e-7 maximum, 4e-8 mean relative error
// in float32 (log1pf: ~32 instructions, this: 9)
CNFOT_HD float m_log1p(float e) {
  float q = 0.005253457929939032f;
  q = fmaf(q, e, -0.02958850748836994f);
  q = fmaf(q, e, 0.07836166769266129f);
  q = fmaf(q, e, -0.13674770295619965f);
  q = fmaf(q, e, 0.19111430644989014f);
  q = fmaf(q, e, -0.24844369292259216f);
  q = fmaf(q, e, 0.33319270610809326f);
  q = fmaf(q, e, -0.49999502301216125f);
  return fmaf(q * e, e, e);
}
#else
CNFOT_HD float m_exp(float v) { return expf(v); }
CNFOT_HD float m_log(float v) { return logf(v); }
CNFOT_HD float m_div(float a, float b) { return a / b; }
CNFOT_HD float m_rcp(float a) { return 1.f / a; }
CNFOT_HD float m_exp_shift_prep(float m) { return m; }
CNFOT_HD float m_exp_shifted(float u, float m) { return expf(u - m); }
CNFOT_HD float m_log1p(float v) { return log1pf(v); }
#endif
CNFOT_HD double m_exp(double v) { return exp(v); }
CNFOT_HD double m_log(double v) { return log(v); }
CNFOT_HD double m_div(double a, double b) { return a / b; }
CNFOT_HD double m_rcp(double a) { return 1.0 / a; }
CNFOT_HD double m_exp_shift_prep(double m) { return m; }
CNFOT_HD double m_exp_shifted(double u, double m) { return exp(u - m); }
CNFOT_HD double m_log1p(double v) { return log1p(v); }
CNFOT_HD float m_sqrt(float v) { return sqrtf(v); }
CNFOT_HD double m_sqrt(double v) { return sqrt(v); }
CNFOT_HD float m_abs(float v) { return fabsf(v); }
CNFOT_HD double m_abs(double v) { return fabs(v); }
CNFOT_HD float m_max(float a, float b) { return fmaxf(a, b); }
CNFOT_HD double m_max(double a, double b) { return fmax(a, b); }
CNFOT_HD float m_min(float a, float b) { return fminf(a, b); }
CNFOT_HD double m_min(double a, double b) { return fmin(a, b); }
CNFOT_HD float m_tiny(float) { return 1.17549435e-38f; }
CNFOT_HD double m_tiny(double) { return 2.2250738585072014e-308; }

// softplus(v) = log(1 + e^v), overflow-safe; sigmoid is its derivative.
template <typename T>
CNFOT_HD T softplus(T v) {
  return m_max(v, (T)0) + m_log1p(m_exp(-m_abs(v)));
}
template <typename T>
CNFOT_HD T sigmoid(T v) {
  T e = m_exp(-m_abs(v));
  T r = m_rcp((T)1 + e);
  return v >= (T)0 ? r : e * r;
}

// Everything one spline evaluation needs again in its backward pass.
template <typename T, int K>
struct SplineState {
  T pw[K];    // softmax probabilities of the width logits
  T ph[K];    // softmax probabilities of the height logits
  int idx;    // selected bin (0 in either tail)
  int tail;   // 0 inside, 1 below the range, 2 above it
  T x0, x1, y0, y1, d0, d1;  // gathered knot data of the selected bin
  T u0, u1;   // raw slope logits (+offset) of the two gathered knots
  T s_tail;   // boundary slope used by the active tail (if any)
  T u_tail;   // its logit (+offset)
};

template <typename T, int K, class SC>
CNFOT_HD void softmax_bins(const T* u, const SC& c, T* prob, T* size) {
  T m = u[0];
#pragma unroll
  for (int k = 1; k < K; ++k) m = m_max(m, u[k]);
  const T sh = m_exp_shift_prep(m);
  T sum = (T)0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    prob[k] = m_exp_shifted(u[k], sh);
    sum += prob[k];
  }
  const T inv = m_rcp(sum);
  const T scale = inv * c.bin_scale;   // one FFMA per bin; prob is only read by the backward pass
#pragma unroll
  for (int k = 0; k < K; ++k) {
    size[k] = prob[k] * scale + c.min_bin;
    prob[k] *= inv;
  }
}

// Knot positions from bin sizes: pos[0] = lo, pos[K] = hi exactly (the last
// knot is the constant range end, not a cumulative sum).
template <typename T, int K, class SC>
CNFOT_HD void knot_positions(const T* size, const SC& c, T* pos) {
  pos[0] = c.lo;
  T acc = (T)0;
#pragma unroll
  for (int k = 1; k < K; ++k) {
    acc += size[k - 1];
    pos[k] = c.lo + acc;
  }
  pos[K] = c.hi;
}

// Bin search + gather.  `search` are the knot positions on the axis the input
// lives on, `other` the positions on the opposite axis.  Half-open bins
// [pos[k], pos[k+1]); outside the range the reference falls back to bin 0.
template <typename T, int K, class SC>
CNFOT_HD void locate(T v, const T* search, const T* other, const T* us,
                     const SC& c, SplineState<T, K>& st, T& s0, T& s1,
                     T& o0, T& o1) {
  int tail = 0;
  if (v <= search[0]) tail = 1;
  if (v >= search[K]) tail = 2;
  int idx = 0;
#pragma unroll
  for (int k = 1; k < K; ++k) idx += (v >= search[k]) ? 1 : 0;
  if (v < search[0] || v >= search[K]) idx = 0;
  s0 = search[0]; s1 = search[1];
  o0 = other[0]; o1 = other[1];
  T u0 = us[0], u1 = us[1];
  // the knots ascend, so "v >= search[k]" is monotone in k: overwriting in ascending order leaves the last bin whose
  // left knot is <= v (outside the range the gathered bin is only read by the lower tail, where no test passes)
  const bool inside = v < search[K];
#pragma unroll
  for (int k = 1; k < K; ++k) {
    const bool hit = inside && (v >= search[k]);
    s0 = hit ? search[k] : s0;
    s1 = hit ? search[k + 1] : s1;
    o0 = hit ? other[k] : o0;
    o1 = hit ? other[k + 1] : o1;
    u0 = hit ? us[k] : u0;
    u1 = hit ? us[k + 1] : u1;
  }
  st.idx = idx;
  st.tail = tail;
  st.u0 = u0 + c.slope_offset;
  st.u1 = u1 + c.slope_offset;
  st.d0 = softplus(st.u0) + c.min_slope;
  st.d1 = softplus(st.u1) + c.min_slope;
  st.s_tail = st.d0;
  st.u_tail = st.u0;
  if (tail == 2) {  // rare: needs the last knot's slope
    st.u_tail = us[K] + c.slope_offset;
    st.s_tail = softplus(st.u_tail) + c.min_slope;
  }
}

// Knot normalisation + bin search of one spline evaluation: raw params -> the gathered knot data in `st`.
// INV = false: the input lives on the x axis (forward map); true: on the y axis (inverse map).
template <bool INV, typename T, int K, class SC>
CNFOT_HD void rqs_locate_raw(T v, const T* theta, const SC& c, SplineState<T, K>& st) {
  T w[K], h[K], xp[K + 1], yp[K + 1];
  softmax_bins<T, K>(theta, c, st.pw, w);
  softmax_bins<T, K>(theta + K, c, st.ph, h);
  knot_positions<T, K>(w, c, xp);
  knot_positions<T, K>(h, c, yp);
  if (INV) locate<T, K>(v, yp, xp, theta + 2 * K, c, st, st.y0, st.y1, st.x0, st.x1);
  else locate<T, K>(v, xp, yp, theta + 2 * K, c, st, st.x0, st.x1, st.y0, st.y1);
}

// ---- the shared `first` spline ------------------------------------------------------------------
// Position 0 of every flow layer uses the SAME unconditioned parameters (`~/first`, flows.py:47-55;
// autoregressive.py:88-92): its knots depend on the weights only, not on the row.  They are normalised once per
// kernel (FirstKnots, in shared memory); a row only searches its bin, and the adjoint is accumulated per KNOT
// (FirstGrad) and pulled back through the softmax / cumulative-sum / softplus once per thread at the end.
template <typename T, int K>
struct FirstKnots {
  T xp[K + 1], yp[K + 1];   // knot positions
  T dk[K + 1];              // knot slopes  softplus(u + offset) + min_slope
  T sg[K + 1];              // sigmoid(u + offset): d slope / d raw
  T pw[K], ph[K];           // softmax probabilities of the bin widths / heights
};
template <typename T, int K>
struct FirstGrad {
  T gx[K + 1], gy[K + 1], gd[K + 1];   // adjoints of the knot positions and slopes (entries 0 and K of gx, gy stay unused)
};

template <typename T, int K, class SC>
CNFOT_HD void first_knots_build(const T* theta, const SC& c, FirstKnots<T, K>& fk) {
  T w[K], h[K];
  softmax_bins<T, K>(theta, c, fk.pw, w);
  softmax_bins<T, K>(theta + K, c, fk.ph, h);
  knot_positions<T, K>(w, c, fk.xp);
  knot_positions<T, K>(h, c, fk.yp);
  for (int k = 0; k <= K; ++k) {
    const T u = theta[2 * K + k] + c.slope_offset;
    fk.dk[k] = softplus(u) + c.min_slope;
    fk.sg[k] = sigmoid(u);
  }
}

// Bin search in the precomputed knots (same bin / tail rules as `locate`).
template <bool INV, typename T, int K, class SC>
CNFOT_HD void rqs_locate_first(T v, const FirstKnots<T, K>& fk, const SC& c, SplineState<T, K>& st) {
  CNFOT_ASSUME_SHARED(&fk);   // device: the contexts keep it in shared memory (LDS, not generic loads)
  const T* search = INV ? fk.yp : fk.xp;
  const T* other = INV ? fk.xp : fk.yp;
  int tail = 0;
  if (v <= search[0]) tail = 1;
  if (v >= search[K]) tail = 2;
  int idx = 0;
#pragma unroll
  for (int k = 1; k < K; ++k) idx += (v >= search[k]) ? 1 : 0;
  if (v < search[0] || v >= search[K]) idx = 0;
  T s0 = search[0], s1 = search[1], o0 = other[0], o1 = other[1], d0 = fk.dk[0], d1 = fk.dk[1];
  const bool inside = v < search[K];
#pragma unroll
  for (int k = 1; k < K; ++k) {
    const bool hit = inside && (v >= search[k]);
    s0 = hit ? search[k] : s0;
    s1 = hit ? search[k + 1] : s1;
    o0 = hit ? other[k] : o0;
    o1 = hit ? other[k + 1] : o1;
    d0 = hit ? fk.dk[k] : d0;
    d1 = hit ? fk.dk[k + 1] : d1;
  }
  st.idx = idx;
  st.tail = tail;
  st.d0 = d0;
  st.d1 = d1;
  st.s_tail = tail == 2 ? fk.dk[K] : d0;
  if (INV) { st.y0 = s0; st.y1 = s1; st.x0 = o0; st.x1 = o1; }
  else { st.x0 = s0; st.x1 = s1; st.y0 = o0; st.y1 = o1; }
}

// this row's knot adjoints -> the per-thread accumulators
template <typename T, int K>
CNFOT_HD void first_grad_add(FirstGrad<T, K>& a, const SplineState<T, K>& st, T gx0, T gx1, T gy0, T gy1, T gd0, T gd1,
                             T gs_tail) {
  if (st.tail == 0) {
    const int i = st.idx;
    a.gx[i] += gx0; a.gx[i + 1] += gx1;
    a.gy[i] += gy0; a.gy[i + 1] += gy1;
    a.gd[i] += gd0; a.gd[i + 1] += gd1;
  } else {
    a.gd[st.tail == 1 ? 0 : K] += gs_tail;
  }
}

// accumulated knot adjoints -> adjoint of the raw `first` parameters (the pull-back `scatter_to_raw` does per row)
template <typename T, int K, class SC>
CNFOT_HD void first_grad_to_raw(const FirstGrad<T, K>& a, const FirstKnots<T, K>& fk, const SC& c, T* gtheta) {
  T gsx[K], gsy[K];
  T accx = (T)0, accy = (T)0;
  for (int i = K - 1; i >= 0; --i) {
    if (i + 1 <= K - 1) { accx += a.gx[i + 1]; accy += a.gy[i + 1]; }   // size[i] feeds every interior knot j >= i + 1
    gsx[i] = accx;
    gsy[i] = accy;
  }
  T dotx = (T)0, doty = (T)0;
  for (int k = 0; k < K; ++k) { dotx += fk.pw[k] * gsx[k]; doty += fk.ph[k] * gsy[k]; }
  for (int k = 0; k < K; ++k) {
    gtheta[k] = c.bin_scale * fk.pw[k] * (gsx[k] - dotx);
    gtheta[K + k] = c.bin_scale * fk.ph[k] * (gsy[k] - doty);
  }
  for (int k = 0; k <= K; ++k) gtheta[2 * K + k] = a.gd[k] * fk.sg[k];
}

// The forward map on located knot data: y = S(x), log|S'(x)|.
template <typename T, int K, class SC>
CNFOT_HD void rqs_forward_map(T x, const SplineState<T, K>& st, const SC& c, T& y, T& logdet) {
  if (st.tail == 0) {
    T bw = st.x1 - st.x0, bh = st.y1 - st.y0;
    T ibw = m_rcp(bw);
    T sl = bh * ibw;
    T z = (x - st.x0) * ibw;
    z = m_min(m_max(z, (T)0), (T)1);
    T z2 = z * z, z1 = z - z2, omz = (T)1 - z;
    T stt = st.d1 + st.d0 - (T)2 * sl;
    T den = sl + stt * z1;
    T iden = m_rcp(den);
    y = st.y0 + bh * (sl * z2 + st.d0 * z1) * iden;
    T A = st.d1 * z2 + (T)2 * sl * z1 + st.d0 * omz * omz;
    // 2 log(sl) + log(A) - 2 log(den) as one logarithm
    T r = sl * iden;
    logdet = m_log(r * r * A);
  } else {
    T px = st.tail == 1 ? c.lo : c.hi;
    y = (x - px) * st.s_tail + px;  // range ends coincide on both axes
    logdet = m_log(st.s_tail);
  }
}

// The inverse map on located knot data: x = S^{-1}(y), log|dS^{-1}/dy|.
template <typename T, int K, class SC>
CNFOT_HD void rqs_inverse_map(T y, const SplineState<T, K>& st, const SC& c, T& x, T& logdet) {
  if (st.tail == 0) {
    T bw = st.x1 - st.x0, bh = st.y1 - st.y0;
    T sl = m_div(bh, bw);
    T w_ = m_div(y - st.y0, bh);
    w_ = m_min(m_max(w_, (T)0), (T)1);
    T stt = st.d1 + st.d0 - (T)2 * sl;
    T qc = -sl * w_;
    T qb = st.d0 - stt * w_;
    T qa = sl - qb;
    T disc = qb * qb - (T)4 * qa * qc;
    T root = disc > (T)0 ? m_sqrt(m_max(disc, m_tiny((T)0))) : (T)0;
    T z = qb >= (T)0 ? m_div((T)2 * qc, -qb - root) : m_div(-qb + root, (T)2 * qa);
    z = m_min(m_max(z, (T)0), (T)1);
    x = bw * z + st.x0;
    T z2 = z * z, z1 = z - z2, omz = (T)1 - z;
    T den = sl + stt * z1;
    T A = st.d1 * z2 + (T)2 * sl * z1 + st.d0 * omz * omz;
    // -(2 log(sl) + log(A) - 2 log(den)) as one logarithm
    T r = m_div(den, sl);
    logdet = m_log(m_div(r * r, A));
  } else {
    T px = st.tail == 1 ? c.lo : c.hi;
    x = m_div(y - px, st.s_tail) + px;
    logdet = -m_log(st.s_tail);
  }
}

// y = S(x), log|S'(x)|.  theta: raw params [K widths | K heights | K+1 slopes].
template <typename T, int K, class SC>
CNFOT_HD void rqs_forward(T x, const T* theta, const SC& c,
                          SplineState<T, K>& st, T& y, T& logdet) {
  rqs_locate_raw<false, T, K>(x, theta, c, st);
  rqs_forward_map<T, K>(x, st, c, y, logdet);
}

// x = S^{-1}(y), log|dS^{-1}/dy|.
template <typename T, int K, class SC>
CNFOT_HD void rqs_inverse(T y, const T* theta, const SC& c,
                          SplineState<T, K>& st, T& x, T& logdet) {
  rqs_locate_raw<true, T, K>(y, theta, c, st);
  rqs_inverse_map<T, K>(y, st, c, x, logdet);
}

// Adjoints of the six gathered knot scalars -> adjoints of the raw params.
// Knot positions: pos[j] = lo + sum_{i<j} size[i] for 1 <= j <= K-1; pos[0]
// and pos[K] are constants, so the last bin's size only gets gradient through
// earlier cumulative sums.  Slopes: softplus'(u) = sigmoid(u).
template <typename T, int K, class SC>
CNFOT_HD void scatter_to_raw(const SplineState<T, K>& st, const SC& c,
                             T gx0, T gx1, T gy0, T gy1, T gd0, T gd1, T gs_tail,
                             T* gtheta) {
  T gsx[K], gsy[K];  // adjoints of the bin sizes
  {
    // adjoint of pos[j] for interior knots j = idx (left) and idx+1 (right)
    T accx = (T)0, accy = (T)0;
#pragma unroll
    for (int i = K - 1; i >= 0; --i) {
      // size[i] feeds pos[j] for all interior j >= i+1
      int j = i + 1;
      if (j <= K - 1) {
        accx += (j == st.idx ? gx0 : (T)0) + (j == st.idx + 1 ? gx1 : (T)0);
        accy += (j == st.idx ? gy0 : (T)0) + (j == st.idx + 1 ? gy1 : (T)0);
      }
      gsx[i] = accx;
      gsy[i] = accy;
    }
  }
  T dotx = (T)0, doty = (T)0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    dotx += st.pw[k] * gsx[k];
    doty += st.ph[k] * gsy[k];
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    gtheta[k] = c.bin_scale * st.pw[k] * (gsx[k] - dotx);
    gtheta[K + k] = c.bin_scale * st.ph[k] * (gsy[k] - doty);
  }
  T gu0 = gd0 * sigmoid(st.u0);
  T gu1 = gd1 * sigmoid(st.u1);
#pragma unroll
  for (int k = 0; k <= K; ++k) {
    gtheta[2 * K + k] = (k == st.idx ? gu0 : (T)0) + (k == st.idx + 1 ? gu1 : (T)0);
  }
  if (st.tail == 1) gtheta[2 * K] += gs_tail * sigmoid(st.u_tail);
  if (st.tail == 2) gtheta[3 * K] += gs_tail * sigmoid(st.u_tail);
}

// Adjoints of the gathered knot data of one evaluation.
template <typename T>
struct KnotAdjoints {
  T gx0, gx1, gy0, gy1, gd0, gd1, gst;
};

// Reverse mode of rqs_forward_map: given (gy, gl) = adjoints of (y, logdet), returns gx and the knot adjoints.
template <typename T, int K, class SC>
CNFOT_HD T rqs_forward_map_bwd(T x, const SplineState<T, K>& st, const SC& c, T gy, T gl, KnotAdjoints<T>& ka) {
  T gx;
  T gx0 = 0, gx1 = 0, gy0 = 0, gy1 = 0, gd0 = 0, gd1 = 0, gst = 0;
  if (st.tail == 0) {
    T bw = st.x1 - st.x0, bh = st.y1 - st.y0;
    T ibw = m_rcp(bw);
    T sl = bh * ibw;
    T zr = (x - st.x0) * ibw;
    T z = m_min(m_max(zr, (T)0), (T)1);
    T z2 = z * z, z1 = z - z2, omz = (T)1 - z, o2 = omz * omz;
    T stt = st.d1 + st.d0 - (T)2 * sl;
    T q = sl * z2 + st.d0 * z1;
    T nu = bh * q;
    T den = sl + stt * z1;
    T iden = m_rcp(den);
    T A = st.d1 * z2 + (T)2 * sl * z1 + st.d0 * o2;
    T g_nu = gy * iden;
    T g_den = -gy * nu * iden * iden - (T)2 * gl * iden;
    T g_A = m_div(gl, A);
    T g_sl = m_div((T)2 * gl, sl) + (T)2 * g_A * z1 + g_den;
    gd1 = g_A * z2;
    gd0 = g_A * o2;
    T g_z2 = g_A * st.d1;
    T g_z1 = (T)2 * g_A * sl + g_den * stt;
    T g_o2 = g_A * st.d0;
    T g_stt = g_den * z1;
    T g_bh = g_nu * q;
    T g_q = g_nu * bh;
    g_sl += g_q * z2;
    g_z2 += g_q * sl;
    gd0 += g_q * z1;
    g_z1 += g_q * st.d0;
    gd1 += g_stt;
    gd0 += g_stt;
    g_sl -= (T)2 * g_stt;
    T g_z = g_z1 + (T)2 * z * (g_z2 - g_z1) - (T)2 * omz * g_o2;
    T g_zr = (zr > (T)0 && zr < (T)1) ? g_z : (T)0;
    gx = g_zr * ibw;
    gx0 = -gx;
    T g_bw = -g_zr * zr * ibw - g_sl * sl * ibw;
    g_bh += g_sl * ibw;
    gx1 = g_bw;
    gx0 -= g_bw;
    gy1 = g_bh;
    gy0 = gy - g_bh;
  } else {
    T px = st.tail == 1 ? c.lo : c.hi;
    gx = gy * st.s_tail;
    gst = gy * (x - px) + m_div(gl, st.s_tail);
  }
  ka.gx0 = gx0; ka.gx1 = gx1; ka.gy0 = gy0; ka.gy1 = gy1; ka.gd0 = gd0; ka.gd1 = gd1; ka.gst = gst;
  return gx;
}

// Reverse mode of rqs_forward: given (gy, gl) = adjoints of (y, logdet),
// returns gx and writes gtheta[3K+1] (overwrites).
template <typename T, int K, class SC>
CNFOT_HD T rqs_forward_bwd(T x, const SplineState<T, K>& st, const SC& c,
                           T gy, T gl, T* gtheta) {
  KnotAdjoints<T> ka;
  const T gx = rqs_forward_map_bwd<T, K>(x, st, c, gy, gl, ka);
  scatter_to_raw<T, K>(st, c, ka.gx0, ka.gx1, ka.gy0, ka.gy1, ka.gd0, ka.gd1, ka.gst, gtheta);
  return gx;
}

// Reverse mode of rqs_inverse_map: (gx_out, gl) = adjoints of (x, logdet); returns the adjoint of the input y
// and the knot adjoints.
template <typename T, int K, class SC>
CNFOT_HD T rqs_inverse_map_bwd(T y, const SplineState<T, K>& st, const SC& c, T gxo, T gl, KnotAdjoints<T>& ka) {
  T gyin;
  T gx0 = 0, gx1 = 0, gy0 = 0, gy1 = 0, gd0 = 0, gd1 = 0, gst = 0;
  if (st.tail == 0) {
    T bw = st.x1 - st.x0, bh = st.y1 - st.y0;
    T ibw = m_rcp(bw), ibh = m_rcp(bh);
    T sl = bh * ibw;
    T wr = (y - st.y0) * ibh;
    T w_ = m_min(m_max(wr, (T)0), (T)1);
    T stt = st.d1 + st.d0 - (T)2 * sl;
    T qc = -sl * w_;
    T qb = st.d0 - stt * w_;
    T qa = sl - qb;
    T disc = qb * qb - (T)4 * qa * qc;
    bool pos = disc > (T)0;
    T root = pos ? m_sqrt(m_max(disc, m_tiny((T)0))) : (T)0;
    bool bpos = qb >= (T)0;
    T dn = bpos ? (-qb - root) : ((T)2 * qa);
    T idn = m_rcp(dn);
    T zr = (bpos ? ((T)2 * qc) : (-qb + root)) * idn;
    T z = m_min(m_max(zr, (T)0), (T)1);
    T z2 = z * z, z1 = z - z2, omz = (T)1 - z, o2 = omz * omz;
    T den = sl + stt * z1;
    T A = st.d1 * z2 + (T)2 * sl * z1 + st.d0 * o2;
    // x = bw * z + x0
    T g_bw = gxo * z;
    T g_z = gxo * bw;
    gx0 = gxo;
    // logdet = -(2 log sl + log A - 2 log den)
    T gF = -gl;
    T g_A = m_div(gF, A);
    T g_den = m_div(-(T)2 * gF, den);
    T g_sl = m_div((T)2 * gF, sl) + (T)2 * g_A * z1 + g_den;
    gd1 = g_A * z2;
    gd0 = g_A * o2;
    T g_z2 = g_A * st.d1;
    T g_z1 = (T)2 * g_A * sl + g_den * stt;
    T g_o2 = g_A * st.d0;
    T g_stt = g_den * z1;
    g_z += g_z1 + (T)2 * z * (g_z2 - g_z1) - (T)2 * omz * g_o2;
    T g_zr = (zr > (T)0 && zr < (T)1) ? g_z : (T)0;
    // z = num / dn
    T g_num = g_zr * idn;
    T g_dn = -g_zr * zr * idn;
    T g_qa = 0, g_qb = 0, g_qc = 0, g_root = 0;
    if (bpos) {
      g_qc += (T)2 * g_num;
      g_qb -= g_dn;
      g_root -= g_dn;
    } else {
      g_qb -= g_num;
      g_root += g_num;
      g_qa += (T)2 * g_dn;
    }
    T g_disc = (pos && disc >= m_tiny((T)0)) ? m_div(g_root, (T)2 * root) : (T)0;
    g_qb += (T)2 * qb * g_disc;
    g_qa -= (T)4 * qc * g_disc;
    g_qc -= (T)4 * qa * g_disc;
    // qa = sl - qb
    g_sl += g_qa;
    g_qb -= g_qa;
    // qb = d0 - stt * w ; qc = -sl * w
    gd0 += g_qb;
    g_stt -= g_qb * w_;
    T g_w = -g_qb * stt - g_qc * sl;
    g_sl -= g_qc * w_;
    // stt = d1 + d0 - 2 sl
    gd1 += g_stt;
    gd0 += g_stt;
    g_sl -= (T)2 * g_stt;
    T g_wr = (wr > (T)0 && wr < (T)1) ? g_w : (T)0;
    gyin = g_wr * ibh;
    T g_bh = -g_wr * wr * ibh + g_sl * ibw;
    g_bw -= g_sl * sl * ibw;
    gx1 = g_bw;
    gx0 -= g_bw;
    gy1 = g_bh;
    gy0 = -gyin - g_bh;
  } else {
    T px = st.tail == 1 ? c.lo : c.hi;
    T is = m_rcp(st.s_tail);
    gyin = gxo * is;
    gst = -gxo * (y - px) * is * is - gl * is;
  }
  ka.gx0 = gx0; ka.gx1 = gx1; ka.gy0 = gy0; ka.gy1 = gy1; ka.gd0 = gd0; ka.gd1 = gd1; ka.gst = gst;
  return gyin;
}

// Reverse mode of rqs_inverse: (gx_out, gl) = adjoints of (x, logdet);
// returns the adjoint of the input y and writes gtheta.
template <typename T, int K, class SC>
CNFOT_HD T rqs_inverse_bwd(T y, const SplineState<T, K>& st, const SC& c,
                           T gxo, T gl, T* gtheta) {
  KnotAdjoints<T> ka;
  const T gyin = rqs_inverse_map_bwd<T, K>(y, st, c, gxo, gl, ka);
  scatter_to_raw<T, K>(st, c, ka.gx0, ka.gx1, ka.gy0, ka.gy1, ka.gd0, ka.gd1, ka.gst, gtheta);
  return gyin;
}

}  // namespace cnfot
