// K1 / K2: fused per-sample Sv / TS.
//
// One pass over HBM: read the raw sample (power dB, AZFP counts, or 4-beam complex), evaluate
//   out = front(x) + slog*log10(R') + 2*alpha*R' - K        
// from the per-row record, write out (+ echo_range).  Replaces the ~10 full-size float64 temporaries of
// calibrate_ek.py:98-171 / calibrate_azfp.py:64-96 / calibrate_ek.py:483-505,564-625 and range.py:138-148.
//
// Layout: rows are contiguous R-float runs; a CTA takes whole rows (grid-stride) so the 128-byte row
// record is read once per row; threads stream float4 (LDG.128 nc / STG.128 cs), 4 independent loads in
// flight per thread.  HBM-bound: 8 B/sample (12 with echo_range); see DESIGN.md for the roofline.
#include "sample_math.cuh"

namespace {
using namespace epb;

// value of one sample given front(x) - K (NaN-propagating); also returns echo_range
__device__ __forceinline__ float sample_out(const RowF& rc, int n, float nf, float frK, float& range_out) {
  range_out = range_of(rc, nf);
  return sv_db(rc, n, nf, frK);
}

// four consecutive input samples as float32 dB: float32 input as stored, or int16 raw power counts (SURVEY.md 8f rank 4:
// K1 reads 2 bytes per sample; count_to_db_f, -32768 = NaN padding, the values epb_ingest_power_i16 would write)
__device__ __forceinline__ float4 load_quad(const float* row, int j) { return ld_stream4(reinterpret_cast<const float4*>(row) + j); }
__device__ __forceinline__ float4 load_quad(const short* row, int j) {
  unsigned lo, hi;
  asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(reinterpret_cast<const uint2*>(row) + j));
  const short q[4] = {(short)(lo & 0xffffu), (short)(lo >> 16), (short)(hi & 0xffffu), (short)(hi >> 16)};
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = (q[k] == (short)-32768) ? CUDART_NAN_F : count_to_db_f((float)q[k]);
  return make_float4(v[0], v[1], v[2], v[3]);
}

template <bool kRange, bool kMinMax, typename XT>
__global__ void __launch_bounds__(256) sv_power_vec4(const XT* __restrict__ x, const epb_row* __restrict__ rows,
                                                     float* __restrict__ out, float* __restrict__ rng,
                                                     float* __restrict__ minmax, long long nrows, int R) {
  const int R4 = R >> 2;
  MinMax mm_v, mm_r;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const RowF rc = load_rowf(rows + row);
    const XT* xin = x + row * (long long)R;
    float4* o4 = reinterpret_cast<float4*>(out + row * (long long)R);
    float4* r4 = kRange ? reinterpret_cast<float4*>(rng + row * (long long)R) : nullptr;
    for (int j0 = threadIdx.x; j0 < R4; j0 += 4 * blockDim.x) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int j = j0 + u * blockDim.x;
        if (j < R4) v[u] = load_quad(xin, j);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int j = j0 + u * blockDim.x;
        if (j >= R4) break;
        float in[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        float o[4], rr[4];
        const float nf0 = (float)(4 * j);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float fr = fmaf(in[k], rc.fscale, rc.foffK);
          o[k] = sample_out(rc, 4 * j + k, nf0 + (float)k, fr, rr[k]);
          if (rc.nanrange && in[k] != in[k]) rr[k] = CUDART_NAN_F;  // range.py:143-148
          if (kMinMax) {
            mm_v.add(o[k]);
            mm_r.add(rr[k]);
          }
        }
        st_stream4(o4 + j, make_float4(o[0], o[1], o[2], o[3]));
        if (kRange) st_stream4(r4 + j, make_float4(rr[0], rr[1], rr[2], rr[3]));
      }
    }
  }
  if (kMinMax) {
    mm_v.flush(minmax + 0, minmax + 1);
    mm_r.flush(minmax + 2, minmax + 3);
  }
}

// K1 on raw counts with rows of R % 8 == 0: one 16-byte load (eight int16 counts) per thread and step, four in flight
template <bool kRange, bool kMinMax>
__global__ void __launch_bounds__(256) sv_power_i16_vec8(const short* __restrict__ x, const epb_row* __restrict__ rows,
                                                         float* __restrict__ out, float* __restrict__ rng,
                                                         float* __restrict__ minmax, long long nrows, int R) {
  const int R8 = R >> 3;
  MinMax mm_v, mm_r;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const RowF rc = load_rowf(rows + row);
    const int4* xin = reinterpret_cast<const int4*>(x + row * (long long)R);
    float4* o4 = reinterpret_cast<float4*>(out + row * (long long)R);
    float4* r4 = kRange ? reinterpret_cast<float4*>(rng + row * (long long)R) : nullptr;
    for (int j0 = threadIdx.x; j0 < R8; j0 += 4 * blockDim.x) {
      int4 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * blockDim.x;
        if (j < R8) asm volatile("ld.global.cs.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(w[u].x), "=r"(w[u].y), "=r"(w[u].z), "=r"(w[u].w) : "l"(xin + j));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * blockDim.x;
        if (j >= R8) break;
        const int ws[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
        float o[8], rr[8];
        const float nf0 = (float)(8 * j);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const short q = (short)((k & 1) ? (ws[k >> 1] >> 16) : (ws[k >> 1] & 0xffff));
          const float in = (q == (short)-32768) ? CUDART_NAN_F : count_to_db_f((float)q);
          const float fr = fmaf(in, rc.fscale, rc.foffK);
          o[k] = sample_out(rc, 8 * j + k, nf0 + (float)k, fr, rr[k]);
          if (rc.nanrange && in != in) rr[k] = CUDART_NAN_F;  // range.py:143-148
          if (kMinMax) {
            mm_v.add(o[k]);
            mm_r.add(rr[k]);
          }
        }
        st_stream4(o4 + 2 * j, make_float4(o[0], o[1], o[2], o[3]));
        st_stream4(o4 + 2 * j + 1, make_float4(o[4], o[5], o[6], o[7]));
        if (kRange) {
          st_stream4(r4 + 2 * j, make_float4(rr[0], rr[1], rr[2], rr[3]));
          st_stream4(r4 + 2 * j + 1, make_float4(rr[4], rr[5], rr[6], rr[7]));
        }
      }
    }
  }
  if (kMinMax) {
    mm_v.flush(minmax + 0, minmax + 1);
    mm_r.flush(minmax + 2, minmax + 3);
  }
}

// scalar fallback for R % 4 != 0 or unaligned bases (ragged last dimension)
template <bool kRange, bool kMinMax, typename XT = float>
__global__ void __launch_bounds__(256) sv_power_scalar(const float* __restrict__ x, const epb_row* __restrict__ rows,
                                                       float* __restrict__ out, float* __restrict__ rng,
                                                       float* __restrict__ minmax, long long nrows, int R) {
  MinMax mm_v, mm_r;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const RowF rc = load_rowf(rows + row);
    const long long base = row * (long long)R;
    for (int n = threadIdx.x; n < R; n += blockDim.x) {
      float in = ld_stream(x + base + n);
      float rr;
      float o = sample_out(rc, n, (float)n, fmaf(in, rc.fscale, rc.foffK), rr);
      if (rc.nanrange && in != in) rr = CUDART_NAN_F;
      out[base + n] = o;
      if (kRange) rng[base + n] = rr;
      if (kMinMax) {
        mm_v.add(o);
        mm_r.add(rr);
      }
    }
  }
  if (kMinMax) {
    mm_v.flush(minmax + 0, minmax + 1);
    mm_r.flush(minmax + 2, minmax + 3);
  }
}

// K2: complex CW samples, beam innermost: one sample = B consecutive floats in each of re / im.
// front = 10*log10(fscale * |nanmean_b x|^2); prx <= 0 -> NaN (calibrate_ek.py:581).
template <int B, bool kRange, bool kMinMax>
__global__ void __launch_bounds__(256) sv_complex_kernel(const float* __restrict__ re, const float* __restrict__ im,
                                                         const epb_row* __restrict__ rows, float* __restrict__ out,
                                                         float* __restrict__ rng, float* __restrict__ minmax,
                                                         long long nrows, int R) {
  MinMax mm_v, mm_r;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const RowF rc = load_rowf(rows + row);
    const long long base = row * (long long)R;
    for (int n = threadIdx.x; n < R; n += blockDim.x) {
      float xr[B], xi[B];
      if (B == 4) {
        float4 a = ld_stream4(reinterpret_cast<const float4*>(re + (base + n) * 4));
        float4 b = ld_stream4(reinterpret_cast<const float4*>(im + (base + n) * 4));
        xr[0] = a.x, xr[1 % B] = a.y, xr[2 % B] = a.z, xr[3 % B] = a.w;
        xi[0] = b.x, xi[1 % B] = b.y, xi[2 % B] = b.z, xi[3 % B] = b.w;
      } else {
#pragma unroll
        for (int b = 0; b < B; ++b) {
          xr[b] = ld_stream(re + (base + n) * B + b);
          xi[b] = ld_stream(im + (base + n) * B + b);
        }
      }
      float sr = 0.f, si = 0.f;
      int cnt = 0;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        bool ok = (xr[b] == xr[b]) && (xi[b] == xi[b]);  // complex NaN if either part is NaN
        sr += ok ? xr[b] : 0.f;
        si += ok ? xi[b] : 0.f;
        cnt += ok;
      }
      float inv = 1.f / (float)cnt;  // cnt == 0 -> inf*0 = NaN below, as nanmean of an all-NaN slice
      float mr = sr * inv, mi = si * inv;
      float prx = rc.fscale * (mr * mr + mi * mi);
      float fr = (prx > 0.f) ? fmaf(kLog2ToDb, fast_log2(prx), rc.foffK) : CUDART_NAN_F;
      float rr;
      float o = sample_out(rc, n, (float)n, fr, rr);
      if (xr[0] != xr[0]) rr = CUDART_NAN_F, o = CUDART_NAN_F;  // range.py:143-145: echo_range (hence R', Sv) is NaN where beam 0 of backscatter_r is
      out[base + n] = o;
      if (kRange) rng[base + n] = rr;
      if (kMinMax) {
        mm_v.add(o);
        mm_r.add(rr);
      }
    }
  }
  if (kMinMax) {
    mm_v.flush(minmax + 0, minmax + 1);
    mm_r.flush(minmax + 2, minmax + 3);
  }
}

__global__ void minmax_init_kernel(float* m) {
  m[0] = CUDART_INF_F;
  m[1] = -CUDART_INF_F;
  m[2] = CUDART_INF_F;
  m[3] = -CUDART_INF_F;
}

}  // namespace

extern "C" int epb_minmax_init(float* minmax, void* stream) {
  EPB_REQUIRE(minmax, "NULL minmax");
  minmax_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(minmax);
  return epb_check_launch("epb_minmax_init");
}

extern "C" int epb_zero(void* ptr, epb_i64 nbytes, void* stream) {
  EPB_REQUIRE(ptr && nbytes >= 0, "bad ptr/size");
  if (cudaMemsetAsync(ptr, 0, (size_t)nbytes, (cudaStream_t)stream) != cudaSuccess)
    return epb_check_launch("epb_zero");
  return EPB_OK;
}

extern "C" int epb_sv_power(const float* x, const epb_row* rows, float* out, float* echo_range, float* minmax,
                            epb_i64 C, epb_i64 P, epb_i64 R, void* stream) {
  EPB_REQUIRE(x && rows && out, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  const long long nrows = C * P;
  const int grid = (int)((nrows < (long long)epb_num_sms() * 8) ? nrows : (long long)epb_num_sms() * 8);
  const bool vec = (R % 4 == 0) && (((uintptr_t)x | (uintptr_t)out | (uintptr_t)echo_range) % 16 == 0);
  using XT = float;
#define EPB_LAUNCH(K)                                                                              \
  do {                                                                                             \
    if (echo_range && minmax)                                                                      \
      K<true, true, XT><<<grid, 256, 0, s>>>(x, rows, out, echo_range, minmax, nrows, (int)R);         \
    else if (echo_range)                                                                           \
      K<true, false, XT><<<grid, 256, 0, s>>>(x, rows, out, echo_range, minmax, nrows, (int)R);        \
    else if (minmax)                                                                               \
      K<false, true, XT><<<grid, 256, 0, s>>>(x, rows, out, echo_range, minmax, nrows, (int)R);        \
    else                                                                                           \
      K<false, false, XT><<<grid, 256, 0, s>>>(x, rows, out, echo_range, minmax, nrows, (int)R);       \
  } while (0)
  if (vec)
    EPB_LAUNCH(sv_power_vec4);
  else
    EPB_LAUNCH(sv_power_scalar);
  return epb_check_launch("epb_sv_power");
}

extern "C" int epb_sv_power_i16(const short* counts, const epb_row* rows, float* out, float* echo_range, float* minmax,
                                epb_i64 C, epb_i64 P, epb_i64 R, void* stream) {
  EPB_REQUIRE(counts && rows && out, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad shape");
  EPB_REQUIRE(R % 4 == 0 && ((uintptr_t)counts % 8) == 0 && (((uintptr_t)out | (uintptr_t)echo_range) % 16) == 0,
              "the int16 kernel needs range_sample % 4 == 0 and aligned arrays (epb_ingest_power_i16 + epb_sv_power otherwise)");
  cudaStream_t s = (cudaStream_t)stream;
  const long long nrows = C * P;
  const int grid = (int)((nrows < (long long)epb_num_sms() * 8) ? nrows : (long long)epb_num_sms() * 8);
  const short* x = counts;
  using XT = short;
  if (R % 8 == 0 && ((uintptr_t)counts % 16) == 0) {
    if (echo_range && minmax)
      sv_power_i16_vec8<true, true><<<grid, 256, 0, s>>>(x, rows, out, echo_range, minmax, nrows, (int)R);
    else if (echo_range)
      sv_power_i16_vec8<true, false><<<grid, 256, 0, s>>>(x, rows, out, echo_range, minmax, nrows, (int)R);
    else if (minmax)
      sv_power_i16_vec8<false, true><<<grid, 256, 0, s>>>(x, rows, out, echo_range, minmax, nrows, (int)R);
    else
      sv_power_i16_vec8<false, false><<<grid, 256, 0, s>>>(x, rows, out, echo_range, minmax, nrows, (int)R);
  } else {
    EPB_LAUNCH(sv_power_vec4);
  }
#undef EPB_LAUNCH
  return epb_check_launch("epb_sv_power_i16");
}

extern "C" int epb_sv_complex(const float* re, const float* im, const epb_row* rows, float* out, float* echo_range,
                              float* minmax, epb_i64 C, epb_i64 P, epb_i64 R, int B, void* stream) {
  EPB_REQUIRE(re && im && rows && out, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad shape");
  EPB_REQUIRE(B >= 1 && B <= 4, "B must be 1..4");
  EPB_REQUIRE(B != 4 || (((uintptr_t)re | (uintptr_t)im) % 16 == 0), "4-beam planes must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const long long nrows = C * P;
  const int grid = (int)((nrows < (long long)epb_num_sms() * 8) ? nrows : (long long)epb_num_sms() * 8);
#define EPB_LAUNCH_B(BB)                                                                                   \
  do {                                                                                                     \
    if (echo_range && minmax)                                                                              \
      sv_complex_kernel<BB, true, true><<<grid, 256, 0, s>>>(re, im, rows, out, echo_range, minmax, nrows, (int)R); \
    else if (echo_range)                                                                                   \
      sv_complex_kernel<BB, true, false><<<grid, 256, 0, s>>>(re, im, rows, out, echo_range, minmax, nrows, (int)R); \
    else if (minmax)                                                                                       \
      sv_complex_kernel<BB, false, true><<<grid, 256, 0, s>>>(re, im, rows, out, echo_range, minmax, nrows, (int)R); \
    else                                                                                                   \
      sv_complex_kernel<BB, false, false><<<grid, 256, 0, s>>>(re, im, rows, out, echo_range, minmax, nrows, (int)R); \
  } while (0)
  switch (B) {
    case 1: EPB_LAUNCH_B(1); break;
    case 2: EPB_LAUNCH_B(2); break;
    case 3: EPB_LAUNCH_B(3); break;
    default: EPB_LAUNCH_B(4); break;
  }
#undef EPB_LAUNCH_B
  return epb_check_launch("epb_sv_complex");
}

// ---- global reductions used for bin-edge construction and actual_range attributes ----------------------------
namespace {
using namespace epb;

// minmax[0..1] = min / max over non-NaN elements, minmax[2] += number of NaN elements (as float, saturating use only)
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ a, long long n, float* __restrict__ minmax) {
  MinMax mm;
  int nan = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = ld_stream(a + i);
    mm.add(v);
    nan |= (v != v);
  }
  mm.flush(minmax + 0, minmax + 1);
  if (__any_sync(0xffffffffu, nan) && (threadIdx.x & 31) == 0) minmax[2] = 1.f;
}

// exact float64 nanmax of the echo_range that compute_Sv would produce: per row, the law value at the last
// sample whose range is not NaN (range.py:143-148: NaN where the input sample is NaN).  Hot case: the last sample
// is valid and the answer is the precomputed rows[row].range_last (one 8-byte and one 4-byte read per row).
__global__ void range_max_kernel(const float* __restrict__ x, const epb_row* __restrict__ rows, long long nrows, int R,
                                 double* __restrict__ out, const int* __restrict__ gate) {
  if (gate && *gate == 0) return;  // the fast fused kernel computed the maximum itself
  const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double best = -CUDART_INF;
  if (row < nrows) {
    const bool nanrule = x && (__ldg(&rows[row].law) & EPB_LAW_NANRANGE);
    const float* xr = x + row * (long long)R;
    if (!nanrule || xr[R - 1] == xr[R - 1]) {
      const double v1 = __ldg(&rows[row].range_last);  // the law is monotone non-decreasing in n
      if (v1 == v1) best = v1;
    } else {
      int n = R - 2;
      while (n >= 0 && xr[n] != xr[n]) --n;
      if (n >= 0) {
        const epb_row r = rows[row];
        const double v1 = law_range(r, n);
        if (v1 == v1) best = v1;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0 && best != -CUDART_INF) atomic_max_d(out, best);
}

__global__ void init_range_max_kernel(double* out) { *out = -CUDART_INF; }

}  // namespace

// consolidate.add_depth: depth = off[c,p] + echo_range * scale[c,p] (float64 parameters, float32 samples)
namespace {
__global__ void __launch_bounds__(256) add_depth_kernel(const float* __restrict__ rng, epb_cp off, epb_cp scale,
                                                        float* __restrict__ out, long long nrows, long long P, int R) {
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const long long c = row / P, p = row % P;
    const float o = (float)cp_at(off, c, p), sc = (float)cp_at(scale, c, p);
    const double od = cp_at(off, c, p), sd = cp_at(scale, c, p);
    const long long base = row * (long long)R;
    if ((R & 3) == 0 && ((((uintptr_t)rng) | ((uintptr_t)out)) & 15) == 0) {
      const float4* r4 = reinterpret_cast<const float4*>(rng + base);
      float4* o4 = reinterpret_cast<float4*>(out + base);
      for (int j = threadIdx.x; j < (R >> 2); j += blockDim.x) {
        const float4 v = ld_stream4(r4 + j);
        // float64 evaluation of two float32-exact inputs, rounded once: within half an ulp of the reference value
        st_stream4(o4 + j, make_float4((float)(od + (double)v.x * sd), (float)(od + (double)v.y * sd),
                                       (float)(od + (double)v.z * sd), (float)(od + (double)v.w * sd)));
      }
    } else {
      for (int n = threadIdx.x; n < R; n += blockDim.x) out[base + n] = (float)(od + (double)rng[base + n] * sd);
    }
    (void)o, (void)sc;
  }
}
}  // namespace

extern "C" int epb_add_depth(const float* echo_range, epb_cp depth_offset, epb_cp scale, float* depth, epb_i64 C, epb_i64 P,
                             epb_i64 R, void* stream) {
  EPB_REQUIRE(echo_range && depth && depth_offset.ptr && scale.ptr, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad shape");
  const long long nrows = C * P;
  const int grid = (int)((nrows < (long long)epb_num_sms() * 8) ? nrows : (long long)epb_num_sms() * 8);
  add_depth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(echo_range, depth_offset, scale, depth, nrows, P, (int)R);
  return epb_check_launch("epb_add_depth");
}

// mask.frequency_differencing / mask.apply_mask (mask/api.py:593-608, :437-438)
namespace {
__device__ __forceinline__ bool cmp_op(float lhs, float diff, int op) {
  switch (op) {
    case 0: return lhs > diff;
    case 1: return lhs < diff;
    case 2: return lhs <= diff;
    case 3: return lhs >= diff;
    default: return lhs == diff;
  }
}
__global__ void __launch_bounds__(256) freq_diff_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        unsigned char* __restrict__ mask, long long n, int op, float diff) {
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n && (((uintptr_t)a | (uintptr_t)b) & 15) == 0 && (((uintptr_t)mask) & 3) == 0) {
      const float4 x = ld_stream4(reinterpret_cast<const float4*>(a + i));
      const float4 y = ld_stream4(reinterpret_cast<const float4*>(b + i));
      const unsigned m = (cmp_op(x.x - y.x, diff, op) ? 1u : 0u) | (cmp_op(x.y - y.y, diff, op) ? 0x100u : 0u) |
                         (cmp_op(x.z - y.z, diff, op) ? 0x10000u : 0u) | (cmp_op(x.w - y.w, diff, op) ? 0x1000000u : 0u);
      *reinterpret_cast<unsigned*>(mask + i) = m;  // NaN differences compare false
    } else {
      for (long long j = i; j < n && j < i + 4; ++j) mask[j] = cmp_op(a[j] - b[j], diff, op) ? 1 : 0;
    }
  }
}
__global__ void __launch_bounds__(256) apply_mask_kernel(const float* __restrict__ src, const unsigned char* __restrict__ mask,
                                                         float* __restrict__ out, long long n, long long plane,
                                                         int mask_has_channel, float fill, const float* __restrict__ fill_plane) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  if ((plane & 3) == 0 && ((((uintptr_t)src) | ((uintptr_t)out) | ((uintptr_t)fill_plane)) & 15) == 0 && (((uintptr_t)mask) & 3) == 0) {
    // four samples per thread and step: 16-byte loads / stores, one 4-byte mask load (the scalar form ran at 0.39 of HBM)
    const long long n4 = n >> 2, plane4 = plane >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
      const long long pi = i % plane4;
      const unsigned m = __ldg(reinterpret_cast<const unsigned*>(mask) + (mask_has_channel ? i : pi));
      const float4 v = ld_stream4(reinterpret_cast<const float4*>(src) + i);
      float4 f = make_float4(fill, fill, fill, fill);
      if (fill_plane) f = __ldg(reinterpret_cast<const float4*>(fill_plane) + pi);
      st_stream4(reinterpret_cast<float4*>(out) + i, make_float4((m & 0xffu) ? v.x : f.x, (m & 0xff00u) ? v.y : f.y,
                                                                  (m & 0xff0000u) ? v.z : f.z, (m & 0xff000000u) ? v.w : f.w));
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long pi = i % plane;
    const long long mi = mask_has_channel ? i : pi;
    out[i] = mask[mi] ? ld_stream(src + i) : (fill_plane ? fill_plane[pi] : fill);
  }
}
}  // namespace

extern "C" int epb_freq_diff_mask(const float* Sv, int chanA, int chanB, int op, float diff, unsigned char* mask, epb_i64 C,
                                  epb_i64 P, epb_i64 R, void* stream) {
  EPB_REQUIRE(Sv && mask, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && chanA >= 0 && chanA < C && chanB >= 0 && chanB < C && op >= 0 && op <= 4, "bad argument");
  const long long n = P * R;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)epb_num_sms() * 16;
  freq_diff_kernel<<<(unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap), 256, 0, (cudaStream_t)stream>>>(
      Sv + chanA * n, Sv + chanB * n, mask, n, op, diff);
  return epb_check_launch("epb_freq_diff_mask");
}

static int apply_mask_impl(const float* src, const unsigned char* mask, int mask_has_channel, float fill_value,
                           const float* fill_plane, float* out, epb_i64 C, epb_i64 P, epb_i64 R, void* stream) {
  EPB_REQUIRE(src && mask && out, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0, "bad shape");
  const long long n = C * P * R;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)epb_num_sms() * 16;
  apply_mask_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(src, mask, out, n, P * R,
                                                                                            mask_has_channel, fill_value, fill_plane);
  return epb_check_launch("epb_apply_mask");
}

extern "C" int epb_apply_mask(const float* src, const unsigned char* mask, int mask_has_channel, float fill_value, float* out,
                              epb_i64 C, epb_i64 P, epb_i64 R, void* stream) {
  return apply_mask_impl(src, mask, mask_has_channel, fill_value, nullptr, out, C, P, R, stream);
}

extern "C" int epb_apply_mask_fill_array(const float* src, const unsigned char* mask, int mask_has_channel,
                                         const float* fill_plane, float* out, epb_i64 C, epb_i64 P, epb_i64 R, void* stream) {
  EPB_REQUIRE(fill_plane, "NULL fill array");
  return apply_mask_impl(src, mask, mask_has_channel, 0.f, fill_plane, out, C, P, R, stream);
}

extern "C" int epb_minmax(const float* a, epb_i64 n, float* minmax, void* stream) {
  EPB_REQUIRE(a && minmax && n > 0, "bad pointer/size");
  const long long blocks = (n + 255) / 256;
  const long long cap = (long long)epb_num_sms() * 16;
  minmax_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(a, n, minmax);
  return epb_check_launch("epb_minmax");
}

extern "C" int epb_range_max(const float* backscatter_r, const epb_row* rows, epb_i64 C, epb_i64 P, epb_i64 R,
                             double* out_max, void* stream) {
  EPB_REQUIRE(rows && out_max && C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad pointer/shape");
  const long long nrows = C * P;
  init_range_max_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out_max);
  range_max_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(backscatter_r, rows, nrows, (int)R,
                                                                                    out_max, nullptr);
  return epb_check_launch("epb_range_max");
}

// used by epb_pipeline_power_mvbs: initialise *out_max, and (gated) compute it when the general kernel does the work
void epb_range_max_init_launch(double* out_max, cudaStream_t s) { init_range_max_kernel<<<1, 1, 0, s>>>(out_max); }
void epb_range_max_gated_launch(const float* x, const epb_row* rows, long long nrows, int R, double* out_max, const int* gate,
                                cudaStream_t s) {
  range_max_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, s>>>(x, rows, nrows, R, out_max, gate);
}
