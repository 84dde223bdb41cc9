// K4 / K5 / K7 on materialised arrays: De Robertis & Higginbottom background-noise estimate and removal
// (clean/api.py:362-433, :436-511) and index-binned MVBS (commongrid/api.py:195-266).
//
// Tile reductions follow xarray's coarsen(ping_time=pn, range_sample=rn, boundary="pad").mean(): the
// volume is cut into (pn x rn) tiles (short edge tiles = NaN padding = fewer members), nanmean per tile.
// Mapping: one CTA per (channel, ping tile); thread-per-column(x4) streams the pn rows with coalesced
// float4 loads and keeps per-column (sum, count) over the rows; the column sums go to shared memory and
// one thread per range tile adds its rn columns (float64) -> mean -> dB; block-wide min for the noise.
// HBM traffic: every input element is read exactly once.
#include "epb_common.cuh"

namespace {
using namespace epb;

// transmission loss of clean/api.py:397-398: 20log10(R if R >= 1 else 1) + 2*alpha*R (NaN R: spreading 0, absorption NaN)
__device__ __forceinline__ float trans_loss(float r, float two_alpha) {
  float rm = (r >= 1.f) ? r : 1.f;
  return fmaf(2.f * kLog2ToDb, fast_log2(rm), two_alpha * r);
}

__device__ __forceinline__ float block_min_skipnan(float v, float* red) {
  // v may be +inf for "no value"; returns the minimum over the block (or +inf)
  v = warp_min(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float x = (lane < (blockDim.x >> 5)) ? red[lane] : CUDART_INF_F;
    x = warp_min(x);
    if (lane == 0) red[0] = x;
  }
  __syncthreads();
  return red[0];
}

// ---- K4: noise[c, tile] ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) noise_estimate_kernel(const float* __restrict__ Sv, const float* __restrict__ rng,
                                                             epb_cp alpha, float* __restrict__ noise, long long P,
                                                             int R, int nPt, int pn, int rn, float noise_max, int vec) {
  extern __shared__ float smem[];
  float* colsum = smem;                                   // [R]
  int* colcnt = reinterpret_cast<int*>(smem + R);         // [R]
  __shared__ float red[32];
  const long long c = blockIdx.x / nPt, i = blockIdx.x % nPt;
  const long long p0 = i * pn, p1 = (p0 + pn < P) ? p0 + pn : P;
  const long long base = c * P * (long long)R;
  if (vec) {
    const int R4 = R >> 2;
    for (int j = threadIdx.x; j < R4; j += blockDim.x) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      int n[4] = {0, 0, 0, 0};
      for (long long p = p0; p < p1; ++p) {
        const float ta = 2.f * (float)cp_at(alpha, c, p);
        const float4 a = ld_stream4(reinterpret_cast<const float4*>(Sv + base + p * R) + j);
        const float4 b = ld_stream4(reinterpret_cast<const float4*>(rng + base + p * R) + j);
        const float sv[4] = {a.x, a.y, a.z, a.w}, rr[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float q = fast_exp2((sv[k] - trans_loss(rr[k], ta)) * kDb2Log2);  // clean/api.py:401
          bool ok = (q == q);
          s[k] += ok ? q : 0.f;
          n[k] += ok;
        }
      }
      *reinterpret_cast<float4*>(colsum + 4 * j) = make_float4(s[0], s[1], s[2], s[3]);
      *reinterpret_cast<int4*>(colcnt + 4 * j) = make_int4(n[0], n[1], n[2], n[3]);
    }
  } else {
    for (int j = threadIdx.x; j < R; j += blockDim.x) {
      float s = 0.f;
      int n = 0;
      for (long long p = p0; p < p1; ++p) {
        const float ta = 2.f * (float)cp_at(alpha, c, p);
        float q = fast_exp2((ld_stream(Sv + base + p * R + j) - trans_loss(ld_stream(rng + base + p * R + j), ta)) * kDb2Log2);
        bool ok = (q == q);
        s += ok ? q : 0.f;
        n += ok;
      }
      colsum[j] = s;
      colcnt[j] = n;
    }
  }
  __syncthreads();
  const int nRt = (R + rn - 1) / rn;
  float best = CUDART_INF_F;
  for (int t = threadIdx.x; t < nRt; t += blockDim.x) {
    const int j0 = t * rn, j1 = (j0 + rn < R) ? j0 + rn : R;
    double s = 0.0;
    int n = 0;
    for (int j = j0; j < j1; ++j) {
      s += (double)colsum[j];
      n += colcnt[j];
    }
    if (n > 0) {
      float m = 10.f * log10f((float)(s / (double)n));  // clean/api.py:402-408
      best = fminf(best, m);                            // min(skipna=True) :411  (fminf drops NaN)
    }
  }
  best = block_min_skipnan(best, red);
  if (threadIdx.x == 0) {
    float v = (best == CUDART_INF_F) ? CUDART_NAN_F : best;
    if (noise_max == noise_max) v = (v < noise_max) ? v : noise_max;  // :418-422 (NaN -> max)
    noise[c * nPt + i] = v;
  }
}

// ---- K5: Sv_noise / Sv_corrected -----------------------------------------------------------------------
// Sv_c = 10log10(10^(Sv/10) - 10^(Sv_noise/10)) evaluated as Sv + 10log10(1 - 10^((Sv_noise-Sv)/10)):
// same value, no cancellation of two large exponentials; "lin > 0" <=> Sv_noise < Sv.
__device__ __forceinline__ void noise_apply_sample(float sv, float r, float two_alpha, float n0, float snr, float& sn,
                                                   float& sc) {
  sn = n0 + trans_loss(r, two_alpha);                       // clean/api.py:425-431
  const float d = sn - sv;
  const float t = fast_exp2(d * kDb2Log2);
  float v = (d < 0.f) ? fmaf(kLog2ToDb, fast_log2(1.f - t), sv) : CUDART_NAN_F;  // :485-486
  sc = (v - sn > snr) ? v : CUDART_NAN_F;                   // :487 (NaN compares false)
}

template <bool kNoise, bool kCorr, bool kMinMax>
__global__ void __launch_bounds__(256) noise_apply_kernel(const float* __restrict__ Sv, const float* __restrict__ rng,
                                                          epb_cp alpha, const float* __restrict__ noise,
                                                          float* __restrict__ out_n, float* __restrict__ out_c,
                                                          float* __restrict__ minmax, long long C, long long P, int R,
                                                          int nPt, int pn, float snr, int vec) {
  MinMax mm_n, mm_c;
  const long long nrows = C * P;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const long long c = row / P, p = row % P;
    const float ta = 2.f * (float)cp_at(alpha, c, p);
    const float n0 = __ldg(noise + c * nPt + p / pn);
    const long long base = row * (long long)R;
    if (vec) {
      const int R4 = R >> 2;
      for (int j = threadIdx.x; j < R4; j += blockDim.x) {
        const float4 a = ld_stream4(reinterpret_cast<const float4*>(Sv + base) + j);
        const float4 b = ld_stream4(reinterpret_cast<const float4*>(rng + base) + j);
        const float sv[4] = {a.x, a.y, a.z, a.w}, rr[4] = {b.x, b.y, b.z, b.w};
        float sn[4], sc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          noise_apply_sample(sv[k], rr[k], ta, n0, snr, sn[k], sc[k]);
          if (kMinMax) {
            mm_n.add(sn[k]);
            mm_c.add(sc[k]);
          }
        }
        if (kNoise) st_stream4(reinterpret_cast<float4*>(out_n + base) + j, make_float4(sn[0], sn[1], sn[2], sn[3]));
        if (kCorr) st_stream4(reinterpret_cast<float4*>(out_c + base) + j, make_float4(sc[0], sc[1], sc[2], sc[3]));
      }
    } else {
      for (int j = threadIdx.x; j < R; j += blockDim.x) {
        float sn, sc;
        noise_apply_sample(ld_stream(Sv + base + j), ld_stream(rng + base + j), ta, n0, snr, sn, sc);
        if (kNoise) out_n[base + j] = sn;
        if (kCorr) out_c[base + j] = sc;
        if (kMinMax) {
          mm_n.add(sn);
          mm_c.add(sc);
        }
      }
    }
  }
  if (kMinMax) {
    mm_n.flush(minmax + 0, minmax + 1);
    mm_c.flush(minmax + 2, minmax + 3);
  }
}

// ---- K7: index binning --------------------------------------------------------------------------------
// out[c,i,j] = 10log10(nanmean over the tile of 10^(Sv/10)); er_out = nanmin of echo_range over the tile.
__global__ void __launch_bounds__(256) coarsen_kernel(const float* __restrict__ Sv, const float* __restrict__ rng,
                                                      float* __restrict__ out, float* __restrict__ er_out, long long P,
                                                      int R, int nPt, int pn, int rn) {
  extern __shared__ float smem[];
  float* colsum = smem;
  int* colcnt = reinterpret_cast<int*>(smem + R);
  float* colmin = smem + 2 * (size_t)R;
  const long long c = blockIdx.x / nPt, i = blockIdx.x % nPt;
  const long long p0 = i * pn, p1 = (p0 + pn < P) ? p0 + pn : P;
  const long long base = c * P * (long long)R;
  if ((R & 3) == 0 && ((((uintptr_t)Sv) | ((uintptr_t)rng)) & 15) == 0) {
    // four columns per thread, two pings in flight: 16-byte loads (the scalar column walk ran at 0.4 of HBM)
    for (int j4 = threadIdx.x; j4 < (R >> 2); j4 += blockDim.x) {
      float s[4] = {0.f, 0.f, 0.f, 0.f}, mn[4] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
      int n[4] = {0, 0, 0, 0};
      auto add = [&](const float4& v, const float4& r) {
        const float vv[4] = {v.x, v.y, v.z, v.w}, rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float q = fast_exp2(vv[k] * kDb2Log2);
          const bool ok = (q == q);
          s[k] += ok ? q : 0.f;
          n[k] += ok;
          mn[k] = fminf(mn[k], rr[k]);
        }
      };
      const float4 inf4 = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
      long long p = p0;
      for (; p + 2 <= p1; p += 2) {
        const float4* a = reinterpret_cast<const float4*>(Sv + base + p * R) + j4;
        const float4 v0 = ld_stream4(a), v1 = ld_stream4(a + (R >> 2));
        float4 r0 = inf4, r1 = inf4;
        if (rng) {
          const float4* b = reinterpret_cast<const float4*>(rng + base + p * R) + j4;
          r0 = ld_stream4(b), r1 = ld_stream4(b + (R >> 2));
        }
        add(v0, r0);
        add(v1, r1);
      }
      if (p < p1) {
        const float4 v0 = ld_stream4(reinterpret_cast<const float4*>(Sv + base + p * R) + j4);
        const float4 r0 = rng ? ld_stream4(reinterpret_cast<const float4*>(rng + base + p * R) + j4) : inf4;
        add(v0, r0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) colsum[4 * j4 + k] = s[k], colcnt[4 * j4 + k] = n[k], colmin[4 * j4 + k] = mn[k];
    }
  } else {
    for (int j = threadIdx.x; j < R; j += blockDim.x) {
      float s = 0.f, mn = CUDART_INF_F;
      int n = 0;
      for (long long p = p0; p < p1; ++p) {
        float q = fast_exp2(ld_stream(Sv + base + p * R + j) * kDb2Log2);
        bool ok = (q == q);
        s += ok ? q : 0.f;
        n += ok;
        if (rng) mn = fminf(mn, ld_stream(rng + base + p * R + j));
      }
      colsum[j] = s;
      colcnt[j] = n;
      colmin[j] = mn;
    }
  }
  __syncthreads();
  const int nRt = (R + rn - 1) / rn;
  for (int t = threadIdx.x; t < nRt; t += blockDim.x) {
    const int j0 = t * rn, j1 = (j0 + rn < R) ? j0 + rn : R;
    double s = 0.0;
    int n = 0;
    float mn = CUDART_INF_F;
    for (int j = j0; j < j1; ++j) {
      s += (double)colsum[j];
      n += colcnt[j];
      mn = fminf(mn, colmin[j]);
    }
    const long long o = (c * nPt + i) * (long long)nRt + t;
    out[o] = (n > 0) ? 10.f * log10f((float)(s / (double)n)) : CUDART_NAN_F;
    if (er_out) er_out[o] = (mn == CUDART_INF_F) ? CUDART_NAN_F : mn;
  }
}

bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) % 16) == 0;
}

}  // namespace

extern "C" int epb_noise_estimate(const float* Sv, const float* echo_range, epb_cp absorption, float* noise, epb_i64 C,
                                  epb_i64 P, epb_i64 R, int ping_num, int range_sample_num, float noise_max,
                                  void* stream) {
  EPB_REQUIRE(Sv && echo_range && absorption.ptr && noise, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad shape");
  EPB_REQUIRE(ping_num > 0 && range_sample_num > 0, "ping_num and range_sample_num must be positive");
  const size_t smem = (size_t)R * 8;
  EPB_REQUIRE(smem <= 200 * 1024, "range_sample dimension too long for the shared-memory column buffer (R <= 25600)");
  const long long nPt = (P + ping_num - 1) / ping_num;
  EPB_REQUIRE(C * nPt < (1LL << 31), "too many ping tiles");
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(noise_estimate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return epb_check_launch("epb_noise_estimate(smem)");
  const int vec = (R % 4 == 0) && aligned16(Sv, echo_range);
  noise_estimate_kernel<<<(unsigned)(C * nPt), 256, smem, (cudaStream_t)stream>>>(
      Sv, echo_range, absorption, noise, P, (int)R, (int)nPt, ping_num, range_sample_num, noise_max, vec);
  return epb_check_launch("epb_noise_estimate");
}

extern "C" int epb_noise_apply(const float* Sv, const float* echo_range, epb_cp absorption, const float* noise,
                               float* Sv_noise, float* Sv_corrected, float* minmax, epb_i64 C, epb_i64 P, epb_i64 R,
                               int ping_num, float snr_threshold, void* stream) {
  EPB_REQUIRE(Sv && echo_range && absorption.ptr && noise, "NULL pointer");
  EPB_REQUIRE(Sv_noise || Sv_corrected || minmax, "nothing to compute");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30) && ping_num > 0, "bad shape");
  const long long nrows = C * P, nPt = (P + ping_num - 1) / ping_num;
  const int grid = (int)((nrows < (long long)epb_num_sms() * 8) ? nrows : (long long)epb_num_sms() * 8);
  const int vec = (R % 4 == 0) && aligned16(Sv, echo_range, Sv_noise, Sv_corrected);
  cudaStream_t s = (cudaStream_t)stream;
#define EPB_NA(N, Cc, M)                                                                                           \
  noise_apply_kernel<N, Cc, M><<<grid, 256, 0, s>>>(Sv, echo_range, absorption, noise, Sv_noise, Sv_corrected,     \
                                                    minmax, C, P, (int)R, (int)nPt, ping_num, snr_threshold, vec)
  const int sel = (Sv_noise ? 4 : 0) | (Sv_corrected ? 2 : 0) | (minmax ? 1 : 0);
  switch (sel) {
    case 1: EPB_NA(false, false, true); break;
    case 2: EPB_NA(false, true, false); break;
    case 3: EPB_NA(false, true, true); break;
    case 4: EPB_NA(true, false, false); break;
    case 5: EPB_NA(true, false, true); break;
    case 6: EPB_NA(true, true, false); break;
    default: EPB_NA(true, true, true); break;
  }
#undef EPB_NA
  return epb_check_launch("epb_noise_apply");
}

extern "C" int epb_coarsen(const float* Sv, const float* echo_range, float* out, float* er_out, epb_i64 C, epb_i64 P,
                           epb_i64 R, int ping_num, int range_sample_num, void* stream) {
  EPB_REQUIRE(Sv && out, "NULL pointer");
  EPB_REQUIRE(!er_out || echo_range, "er_out needs echo_range");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad shape");
  EPB_REQUIRE(ping_num > 0 && range_sample_num > 0, "ping_num and range_sample_num must be positive");
  const size_t smem = (size_t)R * 12;
  EPB_REQUIRE(smem <= 200 * 1024, "range_sample dimension too long for the shared-memory column buffer (R <= 17000)");
  const long long nPt = (P + ping_num - 1) / ping_num;
  EPB_REQUIRE(C * nPt < (1LL << 31), "too many ping tiles");
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(coarsen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return epb_check_launch("epb_coarsen(smem)");
  coarsen_kernel<<<(unsigned)(C * nPt), 256, smem, (cudaStream_t)stream>>>(Sv, er_out ? echo_range : nullptr, out, er_out,
                                                                           P, (int)R, (int)nPt, ping_num,
                                                                           range_sample_num);
  return epb_check_launch("epb_coarsen");
}
