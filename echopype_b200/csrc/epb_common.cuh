// Shared device/host helpers for libepb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/epb200.h"

static_assert(sizeof(epb_row) == 192, "epb_row must be 192 bytes");
static_assert(offsetof(epb_row, a_h) == 128, "float block of epb_row must start at byte 128");

void epb_set_error(const char* fmt, ...);
int epb_check_launch(const char* what);
int epb_num_sms();

#define EPB_REQUIRE(cond, msg)                               \
  do {                                                       \
    if (!(cond)) {                                           \
      epb_set_error("%s: %s", __func__, msg);                \
      return EPB_E_BADARG;                                   \
    }                                                        \
  } while (0)

namespace epb {

constexpr float kLog2_10 = 3.321928094887362f;      // log2(10)
constexpr float kDb2Log2 = 0.3321928094887362f;     // log2(10)/10 : 10^(x/10) = 2^(x*kDb2Log2)
// INDEX2POWER = 10 log10(2) / 256 (convert/parse_base.py:24) as a float32 hi/lo pair: fmaf(count, hi, count * lo) is the
// float32 nearest to the reference's float64 product count * INDEX2POWER for every int16 count (tests/test_ingest.py)
constexpr float kIndex2PowerHi = 0.011758984066545963f;
constexpr float kIndex2PowerLo = 1.3907830442860813e-10f;
__device__ __forceinline__ float count_to_db_f(float q) { return fmaf(q, kIndex2PowerHi, __fmul_rn(q, kIndex2PowerLo)); }
constexpr float kLog2ToDb = 3.0102999566398120f;    // 10*log10(2) : 10*log10(x) = log2(x)*kLog2ToDb

__device__ __forceinline__ double cp_at(const epb_cp& a, long long c, long long p) {
  return a.ptr[c * a.sc + p * a.sp];
}

// ---- exact float64 range law (reference operation order, no FMA contraction) ----------------------
__device__ __forceinline__ double law_range(const epb_row& r, int n) {
  if ((r.law & 0xff) == EPB_LAW_EK) {
    // range.py:138  range_sample * sample_interval * sound_speed / 2
    return __dmul_rn(__dmul_rn((double)n, r.p0), r.p1) * 0.5;
  }
  // range.py:81-89  c*L/(2f) + (c/4)*(((2(n+1)-1)*N*1 - 1)/f + tau) - offset
  long long k = (2LL * ((long long)n + 1) - 1) * (long long)r.azfp_N - 1;
  double inner = __dadd_rn(__ddiv_rn((double)k, r.p2), r.p3);
  return __dadd_rn(__dadd_rn(r.p0, __dmul_rn(r.p1, inner)), -r.p4);
}
// R' under the exact law: (R - off1) - off2  (range.py:188-199 subtracts the two offsets in sequence)
__device__ __forceinline__ double law_tvg_range(const epb_row& r, int n) {
  return __dadd_rn(__dadd_rn(law_range(r, n), -r.off1), -r.off2);
}

__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// float atomic min / max that ignore NaN inputs (callers skip NaN)
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  if (v >= 0.f)
    atomicMin((int*)addr, __float_as_int(v));
  else
    atomicMax((unsigned int*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.f)
    atomicMax((int*)addr, __float_as_int(v));
  else
    atomicMin((unsigned int*)addr, __float_as_uint(v));
}

__device__ __forceinline__ void atomic_max_d(double* addr, double v) {  // v >= 0 or any sign, non-NaN
  long long iv = __double_as_longlong(v);
  if (iv >= 0)
    atomicMax((long long*)addr, iv);
  else
    atomicMin((unsigned long long*)addr, (unsigned long long)iv);
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Running min/max over non-NaN values, flushed once per warp.
struct MinMax {
  float lo, hi;
  __device__ __forceinline__ MinMax() {
    lo = __builtin_huge_valf();
    hi = -__builtin_huge_valf();
  }
  __device__ __forceinline__ void add(float v) {
    lo = fminf(lo, v);  // fminf/fmaxf return the non-NaN operand
    hi = fmaxf(hi, v);
  }
  __device__ __forceinline__ void flush(float* dst_min, float* dst_max) {
    float l = warp_min(lo), h = warp_max(hi);
    if ((threadIdx.x & 31) == 0) {
      if (l != CUDART_INF_F) atomic_min_f(dst_min, l);
      if (h != -CUDART_INF_F) atomic_max_f(dst_max, h);
    }
  }
};

// single-MUFU log2 / exp2 (flush-to-zero variants: no denormal pre-scaling code around the MUFU)
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed FP32 pairs (Blackwell FFMA2 / FADD2: one instruction, two lanes of work): a * b + c and a + b elementwise
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// 10^(x/10) and 10*log10(x) in float32 (accuracy budget: DESIGN.md "numerics")
__device__ __forceinline__ float db2lin(float x) { return exp2f(x * kDb2Log2); }
__device__ __forceinline__ float lin2db(float x) { return 10.f * log10f(x); }

}  // namespace epb
