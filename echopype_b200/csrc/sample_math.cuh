// Per-sample arithmetic shared by the Sv kernels (sv.cu) and the fused pipeline (pipeline.cu).
// Everything here runs on the FP32 pipes: the row record carries float hi/lo splits of the float64
// range law (include/epb200.h (3)), so  R' = fma(a_h,n,bp_h) + fma(a_l,n,bp_l)  is within ~1 ulp of
// the float64 value even where R - offset cancels, and 20*log10(R') uses MUFU.LG2 (abs error of
// lg2.approx <= ~6e-7 in log2 units => < 4e-6 dB).  Error budget vs the float64 oracle: DESIGN.md.
#pragma once
#include "epb_common.cuh"

namespace epb {

struct RowF {  // 64-byte float block of epb_row, loaded with four 16-byte loads
  float a_h, a_l, r0_h, r0_l;
  float bp_h, bp_l, two_alpha, slog2;
  float fscale, foffK, c0, c1;
  float c2, spow;
  int n_start;
  bool nanrange;
};

__device__ __forceinline__ RowF load_rowf(const epb_row* __restrict__ r) {
  const float4* f = reinterpret_cast<const float4*>(&r->a_h);
  float4 q0 = __ldg(f), q1 = __ldg(f + 1), q2 = __ldg(f + 2), q3 = __ldg(f + 3);
  RowF o;
  o.a_h = q0.x, o.a_l = q0.y, o.r0_h = q0.z, o.r0_l = q0.w;
  o.bp_h = q1.x, o.bp_l = q1.y, o.two_alpha = q1.z, o.slog2 = q1.w;
  o.fscale = q2.x, o.foffK = q2.y, o.c0 = q2.z, o.c1 = q2.w;
  o.c2 = q3.x, o.spow = q3.y;  // q3.z, q3.w: range_last (float64), not used by the sample kernels
  o.n_start = __ldg(&r->n_start);
  o.nanrange = (__ldg(&r->law) & EPB_LAW_NANRANGE) != 0;
  return o;
}

// echo_range R(n) and TVG range R'(n); nf = (float)n exactly (n < 2^24)
__device__ __forceinline__ float range_of(const RowF& r, float nf) {
  return fmaf(r.a_h, nf, r.r0_h) + fmaf(r.a_l, nf, r.r0_l);
}
__device__ __forceinline__ float tvg_range_of(const RowF& r, float nf) {
  return fmaf(r.a_h, nf, r.bp_h) + fmaf(r.a_l, nf, r.bp_l);
}

// dB-domain output for a front-end value fr_minus_K = front(x) - K
__device__ __forceinline__ float sv_db(const RowF& r, int n, float nf, float fr_minus_K) {
  float rp = tvg_range_of(r, nf);
  float v = fr_minus_K + fmaf(r.slog2, fast_log2(rp), r.two_alpha * rp);
  return (n >= r.n_start) ? v : CUDART_NAN_F;
}

}  // namespace epb
