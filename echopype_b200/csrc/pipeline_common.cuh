// Helpers shared by the fused pipeline kernels (pipeline.cu: general kernel, pipeline_fast_impl.cuh: register-resident
// persistent kernel): range-only column terms, exact bin-boundary search, TMA / mbarrier wrappers.
#pragma once
#include "sample_math.cuh"

namespace epb {

struct ColC {  // range-only terms of one column under one row law
  float g;     // 10^((Sv - TL)/10) / e : phase 1 factor (NaN for n < n_start)
  float h;     // 10^(Sv/10) / e        : R'^2 * 10^(2 alpha R'/10)
  float tl;    // 10^(TL/10)            : max(R,1)^2 * 10^(2 alpha R/10)
  float rr;    // echo_range
};

__device__ __forceinline__ ColC col_consts(const RowF& r, int n) {
  const float nf = (float)n;
  const float rp = tvg_range_of(r, nf);
  const float rr = range_of(r, nf);
  const float rm = (rr >= 1.f) ? rr : 1.f;
  ColC c;
  float hh = (rp * rp) * fast_exp2(r.c2 * rp);
  hh = (rp < 0.f) ? CUDART_NAN_F : hh;          // log10 of a negative range is NaN in the reference
  c.h = (n >= r.n_start) ? hh : CUDART_NAN_F;   // R' <= 0 -> NaN (calibrate_ek.py:107)
  c.tl = (rm * rm) * fast_exp2(r.c2 * rr);
  c.g = __fdividef(c.h, c.tl);
  c.rr = rr;
  return c;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// smallest n in [0, R] with law_range(row, n) >= edge (closed left) / > edge (closed right)
__device__ __forceinline__ int first_at_or_above(const epb_row& r, int R, double edge, int closed_right) {
  int lo = 0, hi = R;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const double v = law_range(r, mid);
    if (closed_right ? (v > edge) : (v >= edge))
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ int key_of(const int* __restrict__ bnd, int nR, int n) {
  // number of boundaries <= n, minus 1; valid bins are 0..nR-1
  int lo = 0, hi = nR + 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (bnd[mid] <= n)
      lo = mid + 1;
    else
      hi = mid;
  }
  const int k = lo - 1;
  return (k >= 0 && k < nR) ? k : -1;
}


// ---- mbarrier / bulk-copy (TMA) wrappers: SASS SYNCS.* / UBLKCP ------------------------------------------------
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// global -> shared bulk copy of `bytes` (multiple of 16, both sides 16-byte aligned), completion on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace epb
