// clean.mask_impulse_noise / clean.mask_transient_noise with use_index_binning=True.
// Reference: echopype/clean/api.py:30-266 and clean/utils.py:109-189 (index_binning_pool_Sv), :263-317
// (index_binning_downsample_upsample_along_depth), :320-337 (echopy_impulse_noise_mask).  Both masks are window means
// of 10^(Sv/10) over range_sample INDEX windows whose length comes from the mean sample spacing of the range variable
// of each channel: the tile-mean machinery of the noise estimate with disjoint blocks (impulse) or sliding windows
// with reflected borders (transient).  HBM-bound passes; window sums are accumulated in float64.
#include <cstdlib>
#include "epb_common.cuh"

namespace {
using namespace epb;

// ---- per-channel sum / count of the forward differences of a range variable (clean/utils.py:131-133, 280-282:
//      np.nanmean(np.diff(range_var, axis=2), axis=(1, 2))).  float32 neighbours subtract exactly (Sterbenz), the sum
//      telescopes in float64.  One CTA per (channel, ping) row, grid-stride. ----------------------------------------
__global__ void __launch_bounds__(256) range_diff_kernel(const float* __restrict__ rng, long long nrows, long long P, int R,
                                                         double* __restrict__ sum, unsigned long long* __restrict__ cnt) {
  __shared__ double s_s[8];
  __shared__ unsigned s_n[8];
  long long acc_c = -1;  // thread 0: running (sum, count) of the channel the CTA is in
  double acc_s = 0.0;
  unsigned long long acc_n = 0;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const float* r = rng + row * (long long)R;
    double s = 0.0;
    unsigned n = 0;
    if ((R & 3) == 0 && (((uintptr_t)rng) & 15) == 0) {  // four differences per 16-byte load plus the next row element
      for (int j4 = threadIdx.x; j4 < (R >> 2); j4 += blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(r) + j4);
        const float nx = (4 * j4 + 4 < R) ? __ldg(r + 4 * j4 + 4) : CUDART_NAN_F;
        const float d[4] = {v.y - v.x, v.z - v.y, v.w - v.z, nx - v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (d[k] == d[k]) s += (double)d[k], ++n;
      }
    } else {
      for (int j = threadIdx.x; j + 1 < R; j += blockDim.x) {
        const float d = r[j + 1] - r[j];
        if (d == d) s += (double)d, ++n;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0) s_s[threadIdx.x >> 5] = s, s_n[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
      double ts = 0.0;
      unsigned long long tn = 0;
      for (int w = 0; w < 8; ++w) ts += s_s[w], tn += s_n[w];
      // the CTA's rows come in increasing order: one atomic pair per (CTA, channel) instead of one per row (400 000 rows
      // adding to four addresses serialised in L2)
      const long long c = row / P;
      if (c != acc_c) {
        if (acc_n) atomicAdd(sum + acc_c, acc_s), atomicAdd(cnt + acc_c, acc_n);
        acc_c = c, acc_s = 0.0, acc_n = 0;
      }
      acc_s += ts, acc_n += tn;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && acc_n) atomicAdd(sum + acc_c, acc_s), atomicAdd(cnt + acc_c, acc_n);
}

// ---- first flat index i with !(a[i] <= thr)  (clean/utils.py:141: np.argmin((range_var <= exclude_above).data)) ----
__global__ void __launch_bounds__(256) first_not_le_kernel(const float* __restrict__ a, long long n, float thr,
                                                           unsigned long long* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    if ((unsigned long long)i >= *reinterpret_cast<volatile unsigned long long*>(out)) return;  // an earlier hit exists
    if (!(a[i] <= thr)) {
      atomicMin(out, (unsigned long long)i);
      return;
    }
  }
}

// ---- impulse noise, step 1: U[c,p,b] = 10 log10(nanmean of 10^(Sv/10) over the block b of nsamp[c] range samples)
//      (coarsen(range_sample=n, boundary="pad").mean(skipna=True), clean/utils.py:292-302).  CTA per row. -------------
__global__ void __launch_bounds__(256) block_mean_kernel(const float* __restrict__ Sv, const int* __restrict__ nsamp,
                                                         float* __restrict__ U, long long nrows, long long P, int R,
                                                         int nbmax) {
  extern __shared__ float s_row[];
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int n = nsamp[row / P];
    const int nb = (R + n - 1) / n;
    const float* sv = Sv + row * (long long)R;
    float chk = 0.f;  // NaN iff the thread staged a NaN sample (the sum of the dB values)
    if ((R & 3) == 0) {  // 16-byte loads, four in flight per thread
      const float4* sv4 = reinterpret_cast<const float4*>(sv);
      const int R4 = R >> 2;
      for (int j = threadIdx.x; j < R4; j += 4 * blockDim.x) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (j + i * (int)blockDim.x < R4) v[i] = ld_stream4(sv4 + j + i * blockDim.x);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (j + i * (int)blockDim.x < R4) {
            chk += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            *reinterpret_cast<float4*>(s_row + 4 * (j + i * blockDim.x)) =
                make_float4(fast_exp2(v[i].x * kDb2Log2), fast_exp2(v[i].y * kDb2Log2), fast_exp2(v[i].z * kDb2Log2),
                            fast_exp2(v[i].w * kDb2Log2));
          }
      }
    } else {
      for (int j = threadIdx.x; j < R; j += blockDim.x) {
        const float v = ld_stream(sv + j);
        chk += v;
        s_row[j] = fast_exp2(v * kDb2Log2);
      }
    }
    const int row_has_nan = __syncthreads_or(chk != chk);
    if (!row_has_nan) {  // the usual row: plain sums, the member count is the block length
      for (int b = threadIdx.x; b < nbmax; b += blockDim.x) {
        float u = CUDART_NAN_F;
        if (b < nb) {
          const int j0 = b * n, j1 = (j0 + n < R) ? j0 + n : R;
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
          int j = j0;
          for (; j + 4 <= j1; j += 4) s4[0] += s_row[j], s4[1] += s_row[j + 1], s4[2] += s_row[j + 2], s4[3] += s_row[j + 3];
          for (; j < j1; ++j) s4[0] += s_row[j];
          u = kLog2ToDb * fast_log2(__fdividef((s4[0] + s4[1]) + (s4[2] + s4[3]), (float)(j1 - j0)));
        }
        U[row * nbmax + b] = u;
      }
      __syncthreads();
      continue;
    }
    for (int b = threadIdx.x; b < nbmax; b += blockDim.x) {
      float u = CUDART_NAN_F;
      if (b < nb) {
        const int j0 = b * n, j1 = (j0 + n < R) ? j0 + n : R;
        // four interleaved float32 partial sums (blocks are ~10^1..10^2 samples: relative error <= n/4 * 2^-24)
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        int m = 0;
        for (int j = j0; j < j1; j += 4) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float q = (j + i < j1) ? s_row[j + i] : CUDART_NAN_F;
            const bool ok = (q == q);
            s4[i] += ok ? q : 0.f;
            m += ok;
          }
        }
        if (m > 0) u = kLog2ToDb * log2f(((s4[0] + s4[1]) + (s4[2] + s4[3])) / (float)m);
      }
      U[row * nbmax + b] = u;
    }
    __syncthreads();
  }
}

// ---- impulse noise, step 2: the two-sided ping comparison on the block means, upsampled by forward fill
//      (clean/utils.py:305-312 reindex ffill, :320-337 echopy_impulse_noise_mask): NaN differences count as +inf. ----
__global__ void __launch_bounds__(256) impulse_mask_kernel(const float* __restrict__ U, const int* __restrict__ nsamp,
                                                           unsigned char* __restrict__ mask, long long nrows, long long P,
                                                           int R, int nbmax, int k, float thr) {
  extern __shared__ unsigned char s_flag[];  // [nbmax] mask value of every block of the row
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const long long c = row / P, p = row - c * P;
    const int n = nsamp[c];
    const int nb = (R + n - 1) / n;
    const float* u0 = U + row * nbmax;
    const float* uf = (p + k < P) ? U + (row + k) * nbmax : nullptr;
    const float* ub = (p - k >= 0) ? U + (row - k) * nbmax : nullptr;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
      const float v = u0[b];
      float f = v - (uf ? uf[b] : CUDART_NAN_F), w = v - (ub ? ub[b] : CUDART_NAN_F);
      f = (f == f) ? f : CUDART_INF_F;
      w = (w == w) ? w : CUDART_INF_F;
      s_flag[b] = (f > thr && w > thr) ? 1 : 0;
    }
    __syncthreads();
    // forward fill: 16 consecutive samples per thread and step (one 16-byte store when the row allows it); the block
    // index advances with a counter instead of a division per sample
    unsigned char* m = mask + row * (long long)R;
    const bool vec = ((R & 15) == 0);
    for (int j0 = threadIdx.x * 16; j0 < R; j0 += blockDim.x * 16) {
      int b = j0 / n, rem = j0 - b * n;
      unsigned fl = s_flag[b];
      unsigned wv[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        wv[i >> 2] |= fl << (8 * (i & 3));
        if (++rem == n) {
          rem = 0;
          ++b;
          fl = (b < nb) ? s_flag[b] : 0u;
        }
      }
      if (vec) {
        *reinterpret_cast<uint4*>(m + j0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      } else {
        for (int i = 0; i < 16 && j0 + i < R; ++i) m[j0 + i] = (unsigned char)((wv[i >> 2] >> (8 * (i & 3))) & 0xffu);
      }
    }
    __syncthreads();
  }
}

// Step 2 for blocks of at least 16 samples (the usual case: a depth bin of metres at centimetre sample spacing): thread
// per 16 consecutive samples, which touch at most two blocks.  The two block flags come straight from U (three rows, L1 /
// L2 resident: nbmax floats per row), the 16 mask bytes are two byte patterns spliced at the block boundary - no shared
// memory, no barrier, one 16-byte store per thread (the shared-memory forward fill above: 2.6 ms on cfg2 at 8 % of DRAM).
__global__ void __launch_bounds__(256) impulse_mask_wide_kernel(const float* __restrict__ U, const int* __restrict__ nsamp,
                                                                unsigned char* __restrict__ mask, long long nrows, long long P,
                                                                int R, int nbmax, int k, float thr) {
  // CTA per row (grid-stride): the row's channel / ping split and block length are CTA-uniform
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const long long c = row / P, p = row - c * P;
    const int n = __ldg(nsamp + c);
    const int nb = (R + n - 1) / n;
    const float* u0 = U + row * nbmax;
    const bool hasf = p + k < P, hasb = p - k >= 0;
    for (int j0 = threadIdx.x << 4; j0 < R; j0 += blockDim.x << 4) {
    const int b = j0 / n, t = n - (j0 - b * n);  // samples j0 .. j0 + t - 1 belong to block b, the rest to b + 1
    auto flag = [&](int bb) -> unsigned {
      if (bb >= nb) return 0u;
      const float v = __ldg(u0 + bb);
      float f = v - (hasf ? __ldg(u0 + (long long)k * nbmax + bb) : CUDART_NAN_F);
      float w = v - (hasb ? __ldg(u0 - (long long)k * nbmax + bb) : CUDART_NAN_F);
      f = (f == f) ? f : CUDART_INF_F;
      w = (w == w) ? w : CUDART_INF_F;
      return (f > thr && w > thr) ? 0x01010101u : 0u;
    };
    unsigned wv[4] = {0u, 0u, 0u, 0u};
    if (n >= 16) {
      const unsigned f0 = flag(b), f1 = (t < 16) ? flag(b + 1) : 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int tt = t - 4 * q;  // bytes of this word that belong to block b
        const unsigned lo = tt >= 4 ? 0xffffffffu : (tt <= 0 ? 0u : (1u << (8 * tt)) - 1u);
        wv[q] = (f0 & lo) | (f1 & ~lo);
      }
    } else {  // blocks shorter than 16 samples: flag per sample
      for (int i = 0; i < 16; ++i) wv[i >> 2] |= (flag((j0 + i) / n) & 1u) << (8 * (i & 3));
    }
    *reinterpret_cast<uint4*>(mask + row * (long long)R + j0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
  }
}

// ---- impulse noise in ONE pass (rows of up to 4096 samples, blocks of at least 16): a CTA walks a (channel, ping-chunk)
// strip ping by ping with the block means of the 2 k + 1 pings around the output ping in a shared-memory ring, so Sv is
// read once and the mask written once (5 bytes per sample; the two kernels above read the block means back from L2 and
// serialise row load -> barrier -> sums -> barrier per row).  The next row is in flight in registers while the current
// one is summed; rows without NaN (detected with one float add per sample) take plain sums.
constexpr int kImpThreads = 256;
__global__ void __launch_bounds__(kImpThreads) impulse_fused_kernel(const float* __restrict__ Sv, const int* __restrict__ nsamp,
                                                                    float* __restrict__ U, unsigned char* __restrict__ mask,
                                                                    long long P, int R, int nbmax, int k, float thr, int chunk,
                                                                    int nchunks, long long nstrips) {
  extern __shared__ __align__(16) float s_imp[];  // [R] linear row, [(2 k + 1)][nbmax] block means (dB) of the window pings
  float* s_row = s_imp;
  float* s_u = s_imp + R;
  const int W = 2 * k + 1, tid = threadIdx.x, R4 = R >> 2, Pi = (int)P;
  for (long long strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
    const long long c = strip / nchunks;
    const int p0 = (int)((strip - c * nchunks) * chunk), p1 = (p0 + chunk < Pi) ? p0 + chunk : Pi;
    const int n = __ldg(nsamp + c), nb = (R + n - 1) / n;
    const float4* base4 = reinterpret_cast<const float4*>(Sv + c * P * (long long)R);
    const int q0 = (p0 - k > 0) ? p0 - k : 0, q1 = (p1 + k < Pi) ? p1 + k : Pi;  // rows whose block means are needed
    float4 v[4];
    auto load_row = [&](int q) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (tid + i * kImpThreads < R4) v[i] = ld_stream4(base4 + (long long)q * R4 + tid + i * kImpThreads);
    };
    load_row(q0);
    const unsigned ninv = (unsigned)((0x100000000ULL + (unsigned)n - 1) / (unsigned)n);  // j / n = umulhi(j, ninv) for j n < 2^32
    int sq = q0 % W;  // ring slot of row q, advanced with q (a modulo per row and warp cost ~25 % of the instructions)
    for (int q = q0; q < q1 + k; ++q, sq = (sq + 1 == W) ? 0 : sq + 1) {  // the last k steps only emit masks
      if (q < q1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (tid + i * kImpThreads < R4) {
            *reinterpret_cast<float4*>(s_row + 4 * (tid + i * kImpThreads)) =
                make_float4(fast_exp2(v[i].x * kDb2Log2), fast_exp2(v[i].y * kDb2Log2), fast_exp2(v[i].z * kDb2Log2),
                            fast_exp2(v[i].w * kDb2Log2));
          }
        if (q + 1 < q1) load_row(q + 1);
        __syncthreads();
        float* su = s_u + sq * nbmax;
        for (int b = tid; b < nb; b += kImpThreads) {
          const int j0 = b * n, j1 = (j0 + n < R) ? j0 + n : R;
          // plain sums (four interleaved float32 partial sums); a NaN member shows in the result, the block is then redone
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          int j = j0;
          for (; j + 4 <= j1; j += 4) s0 += s_row[j], s1 += s_row[j + 1], s2 += s_row[j + 2], s3 += s_row[j + 3];
          for (; j < j1; ++j) s0 += s_row[j];
          float tot = (s0 + s1) + (s2 + s3);
          int m = j1 - j0;
          if (tot != tot) {
            s0 = s1 = s2 = s3 = 0.f, m = 0;
            for (j = j0; j < j1; j += 4) {
              const float x0 = s_row[j], x1 = (j + 1 < j1) ? s_row[j + 1] : CUDART_NAN_F, x2 = (j + 2 < j1) ? s_row[j + 2] : CUDART_NAN_F,
                          x3 = (j + 3 < j1) ? s_row[j + 3] : CUDART_NAN_F;
              s0 += (x0 == x0) ? x0 : 0.f, s1 += (x1 == x1) ? x1 : 0.f, s2 += (x2 == x2) ? x2 : 0.f, s3 += (x3 == x3) ? x3 : 0.f;
              m += (x0 == x0) + (x1 == x1) + (x2 == x2) + (x3 == x3);
            }
            tot = (s0 + s1) + (s2 + s3);
          }
          const float u = (m > 0) ? kLog2ToDb * fast_log2(__fdividef(tot, (float)m)) : CUDART_NAN_F;
          su[b] = u;
          if (q >= p0 && q < p1) U[(c * P + q) * nbmax + b] = u;
        }
        if (q >= p0 && q < p1)
          for (int b = nb + tid; b < nbmax; b += kImpThreads) U[(c * P + q) * nbmax + b] = CUDART_NAN_F;
      }
      __syncthreads();  // block means of row q visible; s_row free for the next row
      const int p = q - k;  // output ping: its window p - k .. p + k is in the ring (rows outside the ping axis: missing)
      if (p >= p0 && p < p1) {
        // rows p + k, p, p - k sit in slots sq, sq - k, sq - 2 k = sq + 1 (mod W = 2 k + 1)
        const float* u0 = s_u + (sq >= k ? sq - k : sq - k + W) * nbmax;
        const float* uf = (p + k < Pi) ? s_u + sq * nbmax : nullptr;
        const float* ub = (p - k >= 0) ? s_u + (sq + 1 == W ? 0 : sq + 1) * nbmax : nullptr;
        auto flag = [&](int bb) -> unsigned {
          if (bb >= nb) return 0u;
          const float x = u0[bb];
          float f = x - (uf ? uf[bb] : CUDART_NAN_F), w = x - (ub ? ub[bb] : CUDART_NAN_F);
          f = (f == f) ? f : CUDART_INF_F;
          w = (w == w) ? w : CUDART_INF_F;
          return (f > thr && w > thr) ? 0x01010101u : 0u;
        };
        unsigned char* mrow = mask + (c * P + p) * (long long)R;
        for (int j0 = tid << 4; j0 < R; j0 += kImpThreads << 4) {
          const int b = (int)__umulhi((unsigned)j0, ninv), t = n - (j0 - b * n);  // samples j0 .. j0 + t - 1: block b, the rest: b + 1
          const unsigned f0 = flag(b), f1 = (t < 16) ? flag(b + 1) : 0u;
          unsigned wv[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int tt = t - 4 * g;
            const unsigned lo = tt >= 4 ? 0xffffffffu : (tt <= 0 ? 0u : (1u << (8 * tt)) - 1u);
            wv[g] = (f0 & lo) | (f1 & ~lo);
          }
          *reinterpret_cast<uint4*>(mrow + j0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
      }
    }
    __syncthreads();  // the next strip rewrites the ring
  }
}

// ---- impulse noise with depth-VALUE binning (use_index_binning=False, clean/utils.py:192-260): per (channel, ping) row
//      the samples are grouped by the interval [e_b, e_b+1) their depth falls into (depth increases along the row, NaN
//      depths of padded pings count as beyond every edge), U[c,p,b] = dB of the nanmean of 10^(Sv/10) over the interval
//      and F[c,p,b] = first sample with depth >= e_b (what np.digitize + np.unique(return_index) give).  Thread per
//      interval: two bisections on the row in shared memory, then a serial sum.  t32: float32 thresholds ceil32(e_b)
//      (x >= e_b <=> x >= ceil32(e_b) for float32 x). ----------------------------------------------------------------
__global__ void __launch_bounds__(256) depth_bin_mean_kernel(const float* __restrict__ Sv, const float* __restrict__ depth,
                                                             const float* __restrict__ t32, int nb, float* __restrict__ U,
                                                             int* __restrict__ F, float* __restrict__ up, long long nrows,
                                                             int R) {
  extern __shared__ __align__(16) float s_buf[];  // [R] depth (NaN -> +inf), [R] linear Sv, [nb + 1] interval starts
  float* s_d = s_buf;
  float* s_l = s_buf + R;
  int* s_j0 = reinterpret_cast<int*>(s_buf + 2 * R);  // [nb + 1] first sample at or beyond every edge
  auto first_ge = [&](float t) {
    int lo = 0, hi = R;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_d[mid] >= t)
        hi = mid;
      else
        lo = mid + 1;
    }
    return lo;
  };
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const float* sv = Sv + row * (long long)R;
    const float* dp = depth + row * (long long)R;
    if ((R & 3) == 0) {  // 16-byte loads, both arrays in flight
      for (int j = threadIdx.x; j < (R >> 2); j += blockDim.x) {
        const float4 d4 = ld_stream4(reinterpret_cast<const float4*>(dp) + j);
        const float4 v4 = ld_stream4(reinterpret_cast<const float4*>(sv) + j);
        *reinterpret_cast<float4*>(s_d + 4 * j) = make_float4((d4.x == d4.x) ? d4.x : CUDART_INF_F, (d4.y == d4.y) ? d4.y : CUDART_INF_F,
                                                              (d4.z == d4.z) ? d4.z : CUDART_INF_F, (d4.w == d4.w) ? d4.w : CUDART_INF_F);
        *reinterpret_cast<float4*>(s_l + 4 * j) = make_float4(fast_exp2(v4.x * kDb2Log2), fast_exp2(v4.y * kDb2Log2),
                                                              fast_exp2(v4.z * kDb2Log2), fast_exp2(v4.w * kDb2Log2));
      }
    } else {
      for (int j = threadIdx.x; j < R; j += blockDim.x) {
        const float d = ld_stream(dp + j);
        s_d[j] = (d == d) ? d : CUDART_INF_F;
        s_l[j] = fast_exp2(ld_stream(sv + j) * kDb2Log2);
      }
    }
    __syncthreads();
    for (int b = threadIdx.x; b <= nb; b += blockDim.x) s_j0[b] = first_ge(t32[b]);  // one bisection per edge
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
      const int j0 = s_j0[b], j1 = s_j0[b + 1];
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      int m = 0;
      for (int j = j0; j < j1; j += 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float q = (j + i < j1) ? s_l[j + i] : CUDART_NAN_F;
          const bool ok = (q == q);
          s4[i] += ok ? q : 0.f;
          m += ok;
        }
      }
      const float u = (m > 0) ? kLog2ToDb * log2f(((s4[0] + s4[1]) + (s4[2] + s4[3])) / (float)m) : CUDART_NAN_F;
      U[row * nb + b] = u;
      F[row * nb + b] = j0;
      s_d[b] = u;  // the depths are no longer needed: interval means for the forward fill below
    }
    __syncthreads();
    // forward fill (np.digitize on the left edges + reindex ffill): sample j takes the mean of the last interval that
    // starts at or before it; four samples per thread and step, one 16-byte store
    float* urow = up + row * (long long)R;
    for (int j0 = 4 * threadIdx.x; j0 < R; j0 += 4 * blockDim.x) {
      int lo = 0, hi = nb;  // number of interval starts <= j0
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_j0[mid] <= j0)
          lo = mid + 1;
        else
          hi = mid;
      }
      int b = lo - 1;
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        while (b + 1 < nb && s_j0[b + 1] <= j0 + i) ++b;
        v[i] = b >= 0 ? s_d[b] : CUDART_NAN_F;
      }
      if ((R & 3) == 0) {
        *reinterpret_cast<float4*>(urow + j0) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        for (int i = 0; i < 4 && j0 + i < R; ++i) urow[j0 + i] = v[i];
      }
    }
    __syncthreads();
  }
}

// ---- impulse noise with depth-VALUE binning in ONE pass (rows of up to 4096 samples in 16-sample units; the upsampled
// array is not materialised): a CTA walks a (channel, ping-chunk) strip ping by ping, stages the row's depth and linear
// Sv in shared memory (the next row is in flight in registers), bisects the interval starts, sums the intervals, and
// keeps (starts, means) of the 2 k + 1 window pings in a shared-memory ring.  The mask of ping p - k is then formed from
// the ring: every thread walks its 16 samples through the interval starts of the three pings involved (each ping has
// its own starts: np.digitize + forward fill per ping).  9 bytes per sample (Sv, depth, mask) instead of the ~25 of
// depth_bin_mean_kernel + impulse_mask_rows_kernel (which write the upsampled array and read it back three times).
__global__ void __launch_bounds__(kImpThreads, 4) impulse_depth_fused_kernel(const float* __restrict__ Sv, const float* __restrict__ depth,
                                                                          const float* __restrict__ t32, int nb, float* __restrict__ U,
                                                                          int* __restrict__ F, unsigned char* __restrict__ mask, long long P,
                                                                          int R, int k, float thr, int chunk, int nchunks, long long nstrips) {
  extern __shared__ __align__(16) float s_idf[];  // [R] depth (NaN -> +inf), [R] linear Sv, [nb + 1] thresholds, ring [W][2 nb + 1]
  float* s_d = s_idf;
  float* s_l = s_idf + R;
  float* s_t = s_idf + 2 * R;
  // per window ping: [nb + 1] interval starts, [nb] means (float bits), [R / 16] interval of every 16th sample
  int* s_ring = reinterpret_cast<int*>(s_t + nb + 1);
  const int W = 2 * k + 1, RS = 2 * nb + 1 + (R >> 4), tid = threadIdx.x, R4 = R >> 2, Pi = (int)P;
  for (int b = tid; b <= nb; b += kImpThreads) s_t[b] = t32[b];
  for (long long strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
    const long long c = strip / nchunks;
    const int p0 = (int)((strip - c * nchunks) * chunk), p1 = (p0 + chunk < Pi) ? p0 + chunk : Pi;
    const float4* sv4 = reinterpret_cast<const float4*>(Sv + c * P * (long long)R);
    const float4* dp4 = reinterpret_cast<const float4*>(depth + c * P * (long long)R);
    const int q0 = (p0 - k > 0) ? p0 - k : 0, q1 = (p1 + k < Pi) ? p1 + k : Pi;  // rows whose intervals are needed
    int sq = q0 % W;  // ring slot of row q
    for (int q = q0; q < q1 + k; ++q, sq = (sq + 1 == W) ? 0 : sq + 1) {  // the last k steps only emit masks
      if (q < q1) {
        // (no register prefetch of the next row: four CTAs per SM at 64 registers overlap the loads of one strip with
        //  the bisections / sums / mask of the others - with 32 registers of prefetch only two CTAs fit: 12.6 ms)
        for (int i4 = tid; i4 < R4; i4 += 2 * kImpThreads) {
          const bool two = i4 + kImpThreads < R4;
          const float4 d0 = ld_stream4(dp4 + (long long)q * R4 + i4), v0 = ld_stream4(sv4 + (long long)q * R4 + i4);
          float4 d1 = d0, v1 = v0;
          if (two) d1 = ld_stream4(dp4 + (long long)q * R4 + i4 + kImpThreads), v1 = ld_stream4(sv4 + (long long)q * R4 + i4 + kImpThreads);
          auto put = [&](int i, const float4& dd, const float4& vv) {
            *reinterpret_cast<float4*>(s_d + 4 * i) = make_float4((dd.x == dd.x) ? dd.x : CUDART_INF_F, (dd.y == dd.y) ? dd.y : CUDART_INF_F,
                                                                  (dd.z == dd.z) ? dd.z : CUDART_INF_F, (dd.w == dd.w) ? dd.w : CUDART_INF_F);
            *reinterpret_cast<float4*>(s_l + 4 * i) = make_float4(fast_exp2(vv.x * kDb2Log2), fast_exp2(vv.y * kDb2Log2),
                                                                  fast_exp2(vv.z * kDb2Log2), fast_exp2(vv.w * kDb2Log2));
          };
          put(i4, d0, v0);
          if (two) put(i4 + kImpThreads, d1, v1);
        }
        __syncthreads();
        int* sj = s_ring + sq * RS;
        const int* sprev = (q > q0) ? s_ring + (sq == 0 ? W - 1 : sq - 1) * RS : nullptr;  // the previous ping's starts
        for (int b = tid; b <= nb; b += kImpThreads) {  // first sample at or beyond every edge
          const float t = s_t[b];
          // neighbouring pings start their intervals within a sample or two of each other: try the previous ping's start
          // (two loads) before the 12-step bisection
          int lo = sprev ? sprev[b] : -1;
          if (!(lo >= 0 && (lo == 0 || !(s_d[lo - 1] >= t)) && (lo == R || s_d[lo] >= t))) {
            lo = 0;
            int hi = R;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (s_d[mid] >= t)
                hi = mid;
              else
                lo = mid + 1;
            }
          }
          sj[b] = lo;
        }
        __syncthreads();
        for (int b = tid; b < nb; b += kImpThreads) {
          const int j0 = sj[b], j1 = sj[b + 1];
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          int j = j0;
          for (; j + 4 <= j1; j += 4) s0 += s_l[j], s1 += s_l[j + 1], s2 += s_l[j + 2], s3 += s_l[j + 3];
          for (; j < j1; ++j) s0 += s_l[j];
          float tot = (s0 + s1) + (s2 + s3);
          int m = j1 - j0;
          if (tot != tot) {  // a NaN member: redo NaN-aware
            s0 = 0.f, m = 0;
            for (j = j0; j < j1; ++j) {
              const float x = s_l[j];
              if (x == x) s0 += x, ++m;
            }
            tot = s0;
          }
          const float u = (m > 0) ? kLog2ToDb * log2f(tot / (float)m) : CUDART_NAN_F;
          sj[nb + 1 + b] = __float_as_int(u);
          if (q >= p0 && q < p1) {
            U[(c * P + q) * nb + b] = u;
            F[(c * P + q) * nb + b] = j0;
          }
        }
        // the interval of every 16th sample (the last one that starts at or before it; -1: none): one bisection per
        // thread here instead of three per thread in the mask phase of each of the three pings that use this row
        for (int g = tid; g < (R >> 4); g += kImpThreads) {
          const int j = g << 4;
          int lo = 0, hi = nb;  // number of interval starts <= j
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (sj[mid] <= j)
              lo = mid + 1;
            else
              hi = mid;
          }
          sj[2 * nb + 1 + g] = lo - 1;
        }
      }
      __syncthreads();  // intervals of row q visible; s_d / s_l free for the next row
      const int p = q - k;
      if (p >= p0 && p < p1) {
        const int* r0 = s_ring + (sq >= k ? sq - k : sq - k + W) * RS;                     // row p
        const int* rf = (p + k < Pi) ? s_ring + sq * RS : nullptr;                         // row p + k
        const int* rb = (p - k >= 0) ? s_ring + (sq + 1 == W ? 0 : sq + 1) * RS : nullptr;  // row p - k
        // per row: the interval a sample lies in (the last one that starts at or before it; -1: none), its mean and the
        // start of the next interval are kept in registers - an interval holds tens of samples, so the ring is touched
        // again only when a sample crosses into the next one (a dependent shared-memory load per sample and row made the
        // first form of this kernel slower than the two-kernel path)
        struct Cur {
          const int* r;
          int b, nx, nx2;  // current interval, start of the next one and of the one after it
          float v, v1;     // mean of the current interval and of the next one
        };
        const int kNone = 0x7fffffff;
        auto open = [&](const int* r, int j) {  // j: a multiple of 16
          Cur cu;
          cu.r = r, cu.b = -1, cu.nx = kNone, cu.nx2 = kNone, cu.v = CUDART_NAN_F, cu.v1 = CUDART_NAN_F;
          if (r) {
            const int lo = r[2 * nb + 1 + (j >> 4)] + 1;  // number of interval starts <= j (table built with the row)
            cu.b = lo - 1;
            cu.nx = (lo < nb) ? r[lo] : kNone;
            cu.nx2 = (lo + 1 < nb) ? r[lo + 1] : kNone;
            if (cu.b >= 0) cu.v = __int_as_float(r[nb + 1 + cu.b]);
            if (lo < nb) cu.v1 = __int_as_float(r[nb + 1 + lo]);
          }
          return cu;
        };
        auto at = [&](Cur& cu, int j) {
          while (j >= cu.nx) {  // rare
            ++cu.b;
            cu.nx = (cu.b + 1 < nb) ? cu.r[cu.b + 1] : kNone;
            cu.v = __int_as_float(cu.r[nb + 1 + cu.b]);
          }
          return cu.v;
        };
        unsigned char* mrow = mask + (c * P + p) * (long long)R;
        for (int j0 = tid << 4; j0 < R; j0 += kImpThreads << 4) {
          Cur c0 = open(r0, j0), cf = open(rf, j0), cb = open(rb, j0);
          unsigned wv[4] = {0u, 0u, 0u, 0u};
          if (c0.nx2 >= j0 + 16 && cf.nx2 >= j0 + 16 && cb.nx2 >= j0 + 16) {
            // the usual case (intervals of at least 16 samples): the 16 samples touch at most two intervals of each ping -
            // branch-free selects (the interval walk below runs its loop body whenever ANY lane crosses a boundary,
            // which with 32 lanes x 16 samples is almost always: 14 000 instructions per row)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int j = j0 + i;
              const float x = (j >= c0.nx) ? c0.v1 : c0.v;
              float f = x - ((j >= cf.nx) ? cf.v1 : cf.v), w = x - ((j >= cb.nx) ? cb.v1 : cb.v);  // missing ping: NaN
              f = (f == f) ? f : CUDART_INF_F;
              w = (w == w) ? w : CUDART_INF_F;
              wv[i >> 2] |= ((f > thr && w > thr) ? 1u : 0u) << (8 * (i & 3));
            }
          } else {
            for (int i = 0; i < 16; ++i) {
              const int j = j0 + i;
              const float x = at(c0, j);
              float f = x - at(cf, j), w = x - at(cb, j);
              f = (f == f) ? f : CUDART_INF_F;
              w = (w == w) ? w : CUDART_INF_F;
              wv[i >> 2] |= ((f > thr && w > thr) ? 1u : 0u) << (8 * (i & 3));
            }
          }
          *reinterpret_cast<uint4*>(mrow + j0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
      }
    }
    __syncthreads();  // the next strip rewrites the ring
  }
}

// echopy_impulse_noise_mask (clean/utils.py:320-337) on a materialised (channel, ping, range_sample) array of
// downsampled-upsampled Sv: elementwise on the rows p, p + k, p - k; 16 samples per thread, one 16-byte store
__global__ void __launch_bounds__(256) impulse_mask_rows_kernel(const float* __restrict__ up, unsigned char* __restrict__ mask,
                                                                long long nrows, long long P, int R, int k, float thr) {
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const long long c = row / P, p = row - c * P;
    const float* u0 = up + row * (long long)R;
    const float* uf = (p + k < P) ? up + (row + k) * (long long)R : nullptr;
    const float* ub = (p - k >= 0) ? up + (row - k) * (long long)R : nullptr;
    unsigned char* m = mask + row * (long long)R;
    const bool vec = (R & 15) == 0;
    for (int j0 = threadIdx.x * 16; j0 < R; j0 += blockDim.x * 16) {
      unsigned wv[4] = {0u, 0u, 0u, 0u};
      if (vec) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = *reinterpret_cast<const float4*>(u0 + j0 + 4 * q);
          const float4 f4 = uf ? *reinterpret_cast<const float4*>(uf + j0 + 4 * q) : make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
          const float4 b4 = ub ? *reinterpret_cast<const float4*>(ub + j0 + 4 * q) : make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
          const float av[4] = {a.x, a.y, a.z, a.w}, fv[4] = {f4.x, f4.y, f4.z, f4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float f = av[i] - fv[i], w = av[i] - bv[i];
            f = (f == f) ? f : CUDART_INF_F;
            w = (w == w) ? w : CUDART_INF_F;
            wv[q] |= ((f > thr && w > thr) ? 1u : 0u) << (8 * i);
          }
        }
        *reinterpret_cast<uint4*>(m + j0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      } else {
        for (int i = 0; i < 16 && j0 + i < R; ++i) {
          const float v = u0[j0 + i];
          float f = v - (uf ? uf[j0 + i] : CUDART_NAN_F), w = v - (ub ? ub[j0 + i] : CUDART_NAN_F);
          f = (f == f) ? f : CUDART_INF_F;
          w = (w == w) ? w : CUDART_INF_F;
          m[j0 + i] = (f > thr && w > thr) ? 1 : 0;
        }
      }
    }
  }
}

// ---- transient noise, step 1: per (channel, ping) row, sum and count of the valid 10^(Sv/10) over the range window
//      [n - w, n + w] of the array sliced at m0, borders reflected (d c b a | a b c d | d c b a: scipy.ndimage
//      mode="reflect", clean/utils.py:158-170).  Float64 prefix sums in shared memory.  CTA per row. ------------------
__global__ void __launch_bounds__(256) pool_rows_kernel(const float* __restrict__ Sv, const int* __restrict__ nsamp,
                                                        float2* __restrict__ S1, long long nrows, long long P, int R, int m0) {
  extern __shared__ double s_pre[];             // [L + 1] prefix sums of the valid linear values
  const int L = R - m0;
  int* s_cpre = reinterpret_cast<int*>(s_pre + (L + 1));  // [L + 1] prefix counts
  __shared__ double s_ps[256];
  __shared__ int s_pc[256];
  const int per = ((L + blockDim.x - 1) / blockDim.x) | 1;  // odd: the chunk walks of a warp hit distinct banks
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int w = nsamp[row / P];
    const float* sv = Sv + row * (long long)R + m0;
    // chunk-local inclusive prefix, then the offsets of the chunks
    const int a = threadIdx.x * per, b = (a + per < L) ? a + per : L;
    double s = 0.0;
    int m = 0;
    for (int j = a; j < b; ++j) {
      const float q = fast_exp2(sv[j] * kDb2Log2);
      if (q == q) s += (double)q, ++m;
      s_pre[j + 1] = s;
      s_cpre[j + 1] = m;
    }
    s_ps[threadIdx.x] = s, s_pc[threadIdx.x] = m;
    __syncthreads();
    // exclusive prefix of the 256 chunk totals: warp 0 scans them (8 per lane) with shuffles
    if (threadIdx.x < 32) {
      double ls[8], run = 0.0;
      int lc[8], crun = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ls[i] = run, lc[i] = crun;
        run += s_ps[threadIdx.x * 8 + i], crun += s_pc[threadIdx.x * 8 + i];
      }
      double inc = run;
      int cinc = crun;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        const int tc = __shfl_up_sync(0xffffffffu, cinc, o);
        if ((int)threadIdx.x >= o) inc += t, cinc += tc;
      }
      const double base = inc - run;
      const int cbase = cinc - crun;
#pragma unroll
      for (int i = 0; i < 8; ++i) s_ps[threadIdx.x * 8 + i] = base + ls[i], s_pc[threadIdx.x * 8 + i] = cbase + lc[i];
    }
    __syncthreads();
    const double off = s_ps[threadIdx.x];
    const int coff = s_pc[threadIdx.x];
    for (int j = a; j < b; ++j) s_pre[j + 1] += off, s_cpre[j + 1] += coff;
    if (threadIdx.x == 0) s_pre[0] = 0.0, s_cpre[0] = 0;
    __syncthreads();
    float2* o = S1 + row * (long long)R + m0;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
      const int lo = j - w, hi = j + w;
      const int ca = lo < 0 ? 0 : lo, cb = hi >= L ? L - 1 : hi;
      double ws = s_pre[cb + 1] - s_pre[ca];
      int wc = s_cpre[cb + 1] - s_cpre[ca];
      if (lo < 0) ws += s_pre[-lo] - s_pre[0], wc += s_cpre[-lo] - s_cpre[0];                           // [0, -lo - 1]
      if (hi >= L) ws += s_pre[L] - s_pre[2 * L - hi - 1], wc += s_cpre[L] - s_cpre[2 * L - hi - 1];    // [2L-hi-1, L-1]
      o[j] = make_float2((float)ws, (float)wc);
    }
    __syncthreads();
  }
}

// ---- transient noise, step 2: sliding sum of the row windows over pings [p - k, p + k] (reflected), pooled Sv =
//      10 log10(sum / count), mask = Sv - pooled > threshold (clean/api.py:163-166); samples above m0 are never
//      masked (pooled NaN, clean/utils.py:174-176).  Thread per column and chunk of pings, float64 running sums. -----
// (a row of the (sum, count) intermediate is read twice, 2 k + 1 pings apart; L2 evict_last / evict_first cache-hint
// policies on the two reads measured slower: 11.3 -> 12.2 ms, as did capping the resident CTAs)
__global__ void __launch_bounds__(128) pool_pings_mask_kernel(const float2* __restrict__ S1, const float* __restrict__ Sv,
                                                              unsigned char* __restrict__ mask, float* __restrict__ pooled,
                                                              long long P, int R, int m0, int k, float thr, int chunk) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= R) return;
  const long long c = blockIdx.z;
  const long long p0 = (long long)blockIdx.y * chunk, p1 = (p0 + chunk < P) ? p0 + chunk : P;
  const long long base = c * P * (long long)R + n;
  if (n < m0) {
    for (long long p = p0; p < p1; ++p) {
      mask[base + p * R] = 0;
      if (pooled) pooled[base + p * R] = CUDART_NAN_F;
    }
    return;
  }
  auto refl = [&](long long q) { return q < 0 ? -q - 1 : (q >= P ? 2 * P - q - 1 : q); };
  double s = 0.0, m = 0.0;
  for (long long q = p0 - k; q <= p0 + k; ++q) {
    const float2 v = S1[base + refl(q) * R];
    s += (double)v.x, m += (double)v.y;
  }
  // U pings per step: their window loads and Sv loads are issued together; the mean is a float32
  // division of the float64 sums and a single-MUFU log2 (abs. error < 4e-6 dB, as in the Sv kernels)
  auto pooled_db = [&]() { return (m > 0.0) ? kLog2ToDb * fast_log2(__fdividef((float)s, (float)m)) : CUDART_NAN_F; };
  long long p = p0;
  constexpr int U = 2;  // 4 measured no faster
  for (; p + U <= p1; p += U) {
    float sv[U];
    float2 in[U], out[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      sv[u] = ld_stream(Sv + base + (p + u) * R);
      in[u] = S1[base + refl(p + u + k + 1) * R];
      out[u] = S1[base + refl(p + u - k) * R];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float pv = pooled_db();
      s += (double)in[u].x - (double)out[u].x;
      m += (double)in[u].y - (double)out[u].y;
      mask[base + (p + u) * R] = (sv[u] - pv > thr) ? 1 : 0;
      if (pooled) pooled[base + (p + u) * R] = pv;
    }
  }
  for (; p < p1; ++p) {
    const float pv = pooled_db();
    mask[base + p * R] = (ld_stream(Sv + base + p * R) - pv > thr) ? 1 : 0;
    if (pooled) pooled[base + p * R] = pv;
    const float2 in = S1[base + refl(p + k + 1) * R], out = S1[base + refl(p - k) * R];
    s += (double)in.x - (double)out.x;
    m += (double)in.y - (double)out.y;
  }
}

// ---- transient noise with depth-VALUE windows (use_index_binning=False, clean/utils.py:28-105 pool_Sv with nanmean).
//      The reference walks every sample in Python; here: (1) per (channel, ping) row inclusive prefix sums of the valid
//      10^(Sv/10) (float64) and of their count, one warp per row; (2) thread per sample: for each of the 2 k + 1 pings
//      of the window two bisections on that ping's depth row (depth increases along range_sample, NaN = beyond every
//      value) give the samples with |depth - d| <= depth_bin, the prefix differences give their sum and count.
//      O((2 k + 1) log R) per sample: seconds on a 1.6e9-sample volume where the reference needs days. -----------------
__global__ void __launch_bounds__(256) row_prefix_kernel(const float* __restrict__ Sv, double* __restrict__ pre,
                                                         int* __restrict__ cnt, long long nrows, int R) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int per = (R + 31) / 32;
  for (long long row = warp0; row < nrows; row += nwarps) {
    const float* sv = Sv + row * (long long)R;
    double* pr = pre + row * (long long)(R + 1);
    int* cr = cnt + row * (long long)(R + 1);
    const int a = lane * per, b = (a + per < R) ? a + per : R;
    double s = 0.0;
    int m = 0;
    for (int j = a; j < b; ++j) {
      const float q = fast_exp2(sv[j] * kDb2Log2);
      if (q == q) s += (double)q, ++m;
    }
    double inc = s;
    int cinc = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, inc, o);
      const int tc = __shfl_up_sync(0xffffffffu, cinc, o);
      if (lane >= o) inc += t, cinc += tc;
    }
    double run = inc - s;
    int crun = cinc - m;
    if (lane == 0) pr[0] = 0.0, cr[0] = 0;
    for (int j = a; j < b; ++j) {
      const float q = fast_exp2(sv[j] * kDb2Log2);
      if (q == q) run += (double)q, ++crun;
      pr[j + 1] = run;
      cr[j + 1] = crun;
    }
  }
}

__global__ void __launch_bounds__(256) pool_depth_mask_kernel(const float* __restrict__ Sv, const float* __restrict__ depth,
                                                              const double* __restrict__ pre, const int* __restrict__ cnt,
                                                              unsigned char* __restrict__ mask, float* __restrict__ pooled,
                                                              long long P, int R, double dmin, double dmax, double bin,
                                                              double exclude_above, int k, float thr, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long row = i / R;
    const long long c = row / P, p = row - c * P;
    const double d = (double)depth[i];
    float pv = CUDART_NAN_F;
    // clean/utils.py:78-84 (NaN depth fails every comparison)
    if ((d - bin >= dmin) && (d + bin <= dmax) && (d - bin >= exclude_above) && (p - k >= 0) && (p + k <= P)) {
      const double lo_v = d - bin, hi_v = d + bin;
      const long long q1 = (p + k < P) ? p + k : P - 1;
      double s = 0.0;
      long long m = 0;
      for (long long q = p - k; q <= q1; ++q) {
        const float* dr = depth + (c * P + q) * (long long)R;
        int lo = 0, hi = R;  // first sample with depth >= lo_v (NaN counts as +inf)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const float v = dr[mid];
          if (!(v == v) || (double)v >= lo_v)
            hi = mid;
          else
            lo = mid + 1;
        }
        const int i0 = lo;
        hi = R;  // first sample with depth > hi_v
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const float v = dr[mid];
          if (!(v == v) || (double)v > hi_v)
            hi = mid;
          else
            lo = mid + 1;
        }
        const long long o = (c * P + q) * (long long)(R + 1);
        s += pre[o + lo] - pre[o + i0];
        m += cnt[o + lo] - cnt[o + i0];
      }
      if (m > 0) pv = kLog2ToDb * log2f((float)(s / (double)m));
    }
    mask[i] = (Sv[i] - pv > thr) ? 1 : 0;
    if (pooled) pooled[i] = pv;
  }
}

// ---- transient noise with func = nanmedian (clean/api.py:132-145: generic_filter(np.nanmedian) / pool_Sv(func)) --------
// The median of the valid 10^(Sv/10) of a window is a selection, not a sum: no prefix trick.  One thread per output
// sample walks its window once per bit of a radix select over the float32 bit patterns (positive floats order like
// their bit patterns): 32 passes for the lower middle element, 32 more for the upper one when the count is even
// (np.nanmedian averages the two).  O(64 x window) per sample - the reference's own cost class (it evaluates
// np.nanmedian per sample in Python / dask) - meant for the volumes this option is used on, not for cfg2.
template <typename ForEach>
__device__ float window_nanmedian(ForEach for_each) {
  int m = 0;
  for_each([&](float v) { m += (v == v) ? 1 : 0; });
  if (m == 0) return CUDART_NAN_F;
  auto select = [&](int kth) {  // kth smallest (0-based) of the valid values
    unsigned prefix = 0u;
    for (int bit = 31; bit >= 0; --bit) {
      const unsigned hi_mask = (bit == 31) ? 0u : (0xffffffffu << (bit + 1));
      int zeros = 0;
      for_each([&](float v) {
        const unsigned u = __float_as_uint(v);
        zeros += ((v == v) && ((u & hi_mask) == prefix) && !((u >> bit) & 1u)) ? 1 : 0;
      });
      if (kth >= zeros) {
        prefix |= 1u << bit;
        kth -= zeros;
      }
    }
    return __uint_as_float(prefix);
  };
  const float a = select((m - 1) / 2);
  if (m & 1) return a;
  const float b = select(m / 2);
  return 0.5f * (a + b);
}

// index windows: (2 k + 1) pings x (2 w + 1) range samples of the array sliced at m0, both reflected (scipy "reflect")
__global__ void __launch_bounds__(128) transient_median_kernel(const float* __restrict__ Sv, const int* __restrict__ nsamp,
                                                               unsigned char* __restrict__ mask, float* __restrict__ pooled,
                                                               long long P, int R, int m0, int k, float thr, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int L = R - m0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long row = i / R;
    const int n = (int)(i - row * R);
    const long long c = row / P, p = row - c * P;
    float pv = CUDART_NAN_F;
    if (n >= m0) {
      const int w = nsamp[c], j = n - m0;
      const float* base = Sv + c * P * (long long)R + m0;
      pv = window_nanmedian([&](auto&& f) {
        for (long long q = p - k; q <= p + k; ++q) {
          const long long qq = q < 0 ? -q - 1 : (q >= P ? 2 * P - q - 1 : q);
          const float* r = base + qq * (long long)R;
          for (int jj = j - w; jj <= j + w; ++jj) {
            const int jr = jj < 0 ? -jj - 1 : (jj >= L ? 2 * L - jj - 1 : jj);
            f(fast_exp2(r[jr] * kDb2Log2));
          }
        }
      });
      pv = kLog2ToDb * log2f(pv);
    }
    mask[i] = (Sv[i] - pv > thr) ? 1 : 0;
    if (pooled) pooled[i] = pv;
  }
}

// depth-value windows (pool_Sv, clean/utils.py:28-105): the samples of pings p - k .. p + k whose depth lies within
// depth_bin of the sample's own depth; valid windows only (same conditions as pool_depth_mask_kernel)
__global__ void __launch_bounds__(128) transient_median_depth_kernel(const float* __restrict__ Sv, const float* __restrict__ depth,
                                                                     unsigned char* __restrict__ mask, float* __restrict__ pooled,
                                                                     long long P, int R, double dmin, double dmax, double bin,
                                                                     double exclude_above, int k, float thr, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long row = i / R;
    const long long c = row / P, p = row - c * P;
    const double d = (double)depth[i];
    float pv = CUDART_NAN_F;
    if ((d - bin >= dmin) && (d + bin <= dmax) && (d - bin >= exclude_above) && (p - k >= 0) && (p + k <= P)) {
      const double lo_v = d - bin, hi_v = d + bin;
      const long long q1 = (p + k < P) ? p + k : P - 1;
      pv = window_nanmedian([&](auto&& f) {
        for (long long q = p - k; q <= q1; ++q) {
          const float* dr = depth + (c * P + q) * (long long)R;
          const float* sr = Sv + (c * P + q) * (long long)R;
          int lo = 0, hi = R;  // first sample with depth >= lo_v (NaN counts as +inf)
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const float v = dr[mid];
            if (!(v == v) || (double)v >= lo_v)
              hi = mid;
            else
              lo = mid + 1;
          }
          for (int n = lo; n < R; ++n) {
            const float v = dr[n];
            if (!(v == v) || (double)v > hi_v) break;
            f(fast_exp2(sr[n] * kDb2Log2));
          }
        }
      });
      pv = kLog2ToDb * log2f(pv);
    }
    mask[i] = (Sv[i] - pv > thr) ? 1 : 0;
    if (pooled) pooled[i] = pv;
  }
}

// ---- transient noise as ONE strip kernel (rows of up to 4096 samples past exclude_above): no (sum, count) intermediate
// A CTA walks a (channel, ping-chunk) strip ping by ping.  Each thread owns 16 adjacent columns (the first thread starts
// at the 16-aligned column at or below m0; columns above exclude_above are never loaded) and keeps, in float64
// registers, the running sum over the 2 k + 1 pings of the window (reflected at the ends of the ping axis) of the valid
// 10^(Sv/10) of its columns: per output ping one row enters and one row leaves.  The range window [n - w, n + w] of the
// array sliced at m0 (reflected at its ends) is a difference of an inclusive prefix sum ACROSS the columns, rebuilt per
// ping in shared memory (thread-local sums, warp scan, warp totals).  The per-ping scan makes a CTA latency / barrier
// bound, so the CTA is kept at <= 256 threads x 128 registers: TWO CTAs per SM work on independent strips.
//  * prefix layout: column x (relative to the first thread's first column, negative / past the row in the mirror zones)
//    at [x & 15][(x >> 4) + TL] with a compile-time row pitch - a warp's accesses to one j are consecutive, so writes and
//    window reads are bank-conflict free, and for a given w & 15 (a 16-way switch per strip) every access of the window
//    phase is ONE base register plus an immediate
//  * reflection at the ends of the sliced axis without a special window path: the owners of the first / last w columns
//    also store the prefix of the MIRRORED sequence (-I(x) before the axis, 2 I(last) - I(x) past it), so every window
//    is one difference of two entries
//  * NaN samples are counted as per-column DEFICITS (in shared memory: touched only when a NaN enters or leaves); while
//    no column of the CTA carries one (the usual case, tracked with the barrier's OR) the window count is the constant
//    (2 k + 1)(2 w + 1) and the count prefix is skipped
//  * the count prefix is rebuilt only on pings where a deficit CHANGED; between those, a thread whose 16 windows hold no
//    deficit (one difference of the count prefix) still takes the constant-count path
//  * the per-ping chain (scan, two barriers) is shorter than a DRAM round trip: the three rows of the ping kPrefetch
//    steps ahead are prefetched into L2 by TMA (cp.async.bulk.prefetch.L2, one instruction each), so that the register
//    loads issued one ping ahead hit L2
//  * (three CTAs per SM at 192 threads x 112 registers, for rows of <= 3072 columns, measured SLOWER: 8.5 vs 6.6 ms - spills)
//  * NaN detection is one float add per sample (the sum of the raw dB values is NaN iff one of them is); NaN -> 0 by
//    fmaxf(ex2(NaN), 0)
// Reading the entering, the leaving and the centre row is all the traffic there is.
constexpr int kStripThreads = 256, kStripCols = 16, kStripPitch = 320;  // pitch: threads + both mirror zones
constexpr int kStripPrefetch = 4;                                         // pings of L2 prefetch distance
__device__ __forceinline__ void strip_prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// one step of an inclusive warp scan: v + (v of lane - o) where that lane exists (the shuffle's own predicate)
__device__ __forceinline__ double scan_up(double v, int o) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 lo, hi, a, b;\n\t.reg .f64 t;\n\t"
      "mov.b64 {lo, hi}, %0;\n\t"
      "shfl.sync.up.b32 a|p, lo, %1, 0, 0xffffffff;\n\t"
      "shfl.sync.up.b32 b, hi, %1, 0, 0xffffffff;\n\t"
      "mov.b64 t, {a, b};\n\t"
      "@p add.f64 %0, %0, t;\n\t}"
      : "+d"(v)
      : "r"(o));
  return v;
}
__device__ __forceinline__ int scan_up(int v, int o) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 a;\n\t"
      "shfl.sync.up.b32 a|p, %0, %1, 0, 0xffffffff;\n\t"
      "@p add.s32 %0, %0, a;\n\t}"
      : "+r"(v)
      : "r"(o));
  return v;
}

// float <-> double without F2F: the conversions share the XU pipe with MUFU, which this kernel already loads with three
// operations per sample (ncu showed the XU pipe as the busiest one).  Integer forms on the ALU pipe: exact for positive
// normal floats (ex2.ftz never returns a denormal; +inf is not expected), truncating double -> float (1 ulp = 2.6e-7 dB)
// with everything below 2^-126 -> 0.
// float -> double for floats known to be positive normal (a zero case costs two more instructions per sample)
__device__ __forceinline__ double f2d_norm(float f) {
  const unsigned b = __float_as_uint(f);
  return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}
__device__ __forceinline__ float d2f_trunc(double d) {
  const int hi = __double2hiint(d);
  const unsigned r = __funnelshift_l((unsigned)__double2loint(d), (unsigned)(hi - 0x38000000), 3);
  return hi >= 0x38100000 ? __uint_as_float(r) : 0.f;
}

template <int W15>
__device__ __forceinline__ void strip_windows(const double* __restrict__ sph, const double* __restrict__ spl, float inv_full,
                                              float (&pv)[kStripCols]) {
  constexpr int L15 = 15 - W15;  // (-w - 1) & 15
#pragma unroll
  for (int j = 0; j < kStripCols; ++j) {
    const double hi = sph[((j + W15) & 15) * kStripPitch + ((j + W15) >> 4)], lo = spl[((j + L15) & 15) * kStripPitch + ((j + L15) >> 4)];
    pv[j] = kLog2ToDb * fast_log2(d2f_trunc(hi - lo) * inv_full);
  }
}

// kDepth: windows of depth VALUES (clean/utils.py:28-105 pool_Sv, use_index_binning=False) on volumes whose depth rows are
// the same for every ping of a channel: the samples with |depth - d| <= depth_bin then form the SAME index interval
// [i0(n), i1(n)) in every ping of the window, so the pooled value is again one difference of the prefix of the running
// column sums - with per-column interval ends (tables from depth_window_table_kernel, staged in shared memory) instead
// of n -+ w, no reflection (pings / depths too close to the ends of their axes give NaN, the window is cut at the last
// ping) and no slicing at exclude_above (it enters through the validity of a column).
template <bool kPooled, bool kDepth = false>
__global__ void __launch_bounds__(kStripThreads, 2)
    transient_strip_kernel(const float* __restrict__ Sv, const int* __restrict__ nsamp, unsigned char* __restrict__ mask,
                           float* __restrict__ pooled, long long P, int R, int m0, int k, float thr, int chunk, int nchunks,
                           long long nstrips, int TL, const unsigned short* __restrict__ wtab = nullptr,
                           const float* __restrict__ depth = nullptr) {
  constexpr int TS = kStripPitch;
  extern __shared__ double s_pre[];                              // [16][TS] inclusive prefix of the running column sums
  int* s_cnt = reinterpret_cast<int*>(s_pre + kStripCols * TS);  // [16][TS] inclusive prefix of the deficits
  int* s_cc = s_cnt + kStripCols * TS;                           // [16][kStripThreads] deficits per column
  unsigned short* s_wt = reinterpret_cast<unsigned short*>(s_cc + kStripCols * kStripThreads);  // kDepth: [3][R] lo slot, hi slot, width
  __shared__ double s_ws[kStripThreads / 32];
  __shared__ int s_wc[kStripThreads / 32];
  const int T = blockDim.x, NW = T >> 5;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int col0 = m0 & ~15, Pi = (int)P;
  const int n0 = col0 + tid * kStripCols;  // first absolute column of this thread
  const bool active = n0 < R;
  // Leading columns of the first thread above the sliced axis: they run through the scan like the others (no patching
  // of loaded values); the prefix then carries the base B = I(m0 - 1), which cancels in every window and enters the
  // mirrored entries before the axis as 2 B - I(x).  Their writers (all in the first warp: w <= 480) read B from the
  // first thread's own entry after a __syncwarp.  Threads past the row load nothing and add nothing.
  const int nkill = m0 - n0 > 0 ? m0 - n0 : 0;
  auto slot = [&](int x) { return (x & 15) * TS + (x >> 4) + TL; };
  if (tid == 0) s_pre[slot(-1)] = 0.0, s_cnt[slot(-1)] = 0;  // "before the first column" when m0 is 16-aligned
  double* const my_pre = s_pre + tid + TL;
  int* const my_cnt = s_cnt + tid + TL;
  int* const my_cc = s_cc + tid;
  for (long long strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
    const long long c = strip / nchunks;
    const int p0 = (int)((strip - c * nchunks) * chunk), p1 = (p0 + chunk < Pi) ? p0 + chunk : Pi;
    const int w = kDepth ? 0 : nsamp[c];
    if (kDepth) {  // the channel's window tables (the previous strip's readers are past its final barrier)
      const uint4* src = reinterpret_cast<const uint4*>(wtab + c * 3 * (long long)R);
      for (int i = tid; i < (3 * R) >> 3; i += T) reinterpret_cast<uint4*>(s_wt)[i] = __ldg(src + i);
    }
    const float* colbase = Sv + c * P * (long long)R + (active ? n0 : 0);
    // index windows: pings reflect at the ends of the axis; depth windows: pings outside the axis do not exist
    auto refl = [&](int q) { return kDepth ? q : (q < 0 ? -q - 1 : (q >= Pi ? 2 * Pi - q - 1 : q)); };
    auto load16 = [&](int q, float (&v)[kStripCols]) {
      if (kDepth && (q < 0 || q >= Pi)) {  // CTA-uniform
#pragma unroll
        for (int j = 0; j < kStripCols; ++j) v[j] = -CUDART_INF_F;  // contributes nothing, never a deficit
        return;
      }
      const float4* r4 = reinterpret_cast<const float4*>(colbase + (long long)q * R);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 a = __ldg(r4 + g);
        v[4 * g] = a.x, v[4 * g + 1] = a.y, v[4 * g + 2] = a.z, v[4 * g + 3] = a.w;
      }
    };
    double cs[kStripCols];
    int tchg = 0;      // a deficit of this thread's columns changed since the count prefix was built
    int has_def = 0;   // CTA-uniform: the count prefix in shared memory is valid and holds deficits
#pragma unroll
    for (int j = 0; j < kStripCols; ++j) cs[j] = 0.0, my_cc[j * kStripThreads] = 0;
    // NaN (and anything below the float range) becomes the smallest normal float, 1.2e-38 (-379 dB): it enters and leaves
    // the running sums as the same value, 28 orders of magnitude below a -100 dB sample, and never reaches an output
    // that has no valid member (the count decides that); in exchange the conversion needs no zero case
    auto lin = [](float v) { return f2d_norm(fmaxf(fast_exp2(v * kDb2Log2), 1.17549435e-38f)); };
    // deficits of this thread's columns: +1 per NaN that enters, -1 per NaN that leaves (rare path)
    auto count_nans = [&](const float (&vi)[kStripCols], const float (&vo)[kStripCols]) {
#pragma unroll
      for (int j = 0; j < kStripCols; ++j) {
        const int d = (vi[j] != vi[j] ? 1 : 0) - (vo[j] != vo[j] ? 1 : 0);
        if (d) my_cc[j * kStripThreads] += d, tchg = 1;
      }
    };
    // window of the first output ping, minus the row that enters in the first step
    if (active) {
      float none[kStripCols];
#pragma unroll
      for (int j = 0; j < kStripCols; ++j) none[j] = 0.f;
      for (int q = p0 - k; q < p0 + k; ++q) {
        float v[kStripCols];
        load16(refl(q), v);
        float chk = 0.f;
#pragma unroll
        for (int j = 0; j < kStripCols; ++j) cs[j] += lin(v[j]), chk += v[j];
        if (chk != chk) count_nans(v, none);
      }
    }
    const int ifull = (2 * k + 1) * (2 * w + 1);
    const float inv_full = 1.f / (float)ifull;
    // owners of the first / last w columns of the sliced axis also write the mirrored prefix
    // columns of this WARP among the first w (sources of the mirrored entries before the axis: [m0, m0 + w - 1]) and among
    // the last w before the final column ([R - 1 - w, R - 2]); empty ranges (lo > hi) for the other warps / depth windows
    const int wx0 = col0 + wid * 32 * kStripCols, wx1 = (wx0 + 32 * kStripCols - 1 < R - 1) ? wx0 + 32 * kStripCols - 1 : R - 1;
    const int mir_lo0 = kDepth ? 1 : (m0 > wx0 ? m0 : wx0), mir_lo1 = kDepth ? 0 : (m0 + w - 1 < wx1 ? m0 + w - 1 : wx1);
    const int mir_hi0 = kDepth ? 1 : (R - 1 - w > wx0 ? R - 1 - w : wx0), mir_hi1 = kDepth ? 0 : (R - 2 < wx1 ? R - 2 : wx1);
    const double* const sph = my_pre + (w >> 4);
    const double* const spl = my_pre + ((-w - 1) >> 4);
    // software pipeline: the entering and leaving rows of the NEXT ping are requested right after this ping's were used
    float nin[kStripCols], nout[kStripCols];
    if (active) load16(refl(p0 + k), nin);
#pragma unroll
    for (int j = 0; j < kStripCols; ++j) nout[j] = -CUDART_INF_F;  // nothing leaves in the first step
    // rows of the sliced axis as TMA prefetches: bytes from the first thread's first column to the end of the row
    const float* rowbase = Sv + c * P * (long long)R + col0;
    const unsigned rowbytes = (unsigned)(R - col0) * 4u;
    auto prefetch_ping = [&](int pp) {  // the three rows ping pp will read
      if (pp < p1) {
        const int qi = pp + k, qo = pp - k - 1;
        if (!kDepth || qi < Pi) strip_prefetch_l2(rowbase + (long long)(qi >= Pi ? 2 * Pi - qi - 1 : qi) * R, rowbytes);
        if (!kDepth || qo >= 0) strip_prefetch_l2(rowbase + (long long)(qo < 0 ? -qo - 1 : qo) * R, rowbytes);
        strip_prefetch_l2(rowbase + (long long)pp * R, rowbytes);
      }
    };
    if (tid == 0)
      for (int d = 1; d < kStripPrefetch; ++d) prefetch_ping(p0 + d);
    for (int p = p0; p < p1; ++p) {
      if (tid == 0) prefetch_ping(p + kStripPrefetch);
      double run = 0.0;
      if (active) {
        float chk = 0.f;
#pragma unroll
        for (int j = 0; j < kStripCols; ++j) {
          chk += nin[j] + nout[j];
          cs[j] += lin(nin[j]);
          cs[j] -= lin(nout[j]);
        }
        if (chk != chk) count_nans(nin, nout);
        if (p + 1 < p1) {
          const int qi = p + 1 + k, qo = p - k;  // reflected at the ends of the ping axis (index windows)
          load16(kDepth ? qi : (qi >= Pi ? 2 * Pi - qi - 1 : qi), nin);
          load16(kDepth ? qo : (qo < 0 ? -qo - 1 : qo), nout);
        }
        run = (((cs[0] + cs[1]) + (cs[2] + cs[3])) + ((cs[4] + cs[5]) + (cs[6] + cs[7]))) +
              (((cs[8] + cs[9]) + (cs[10] + cs[11])) + ((cs[12] + cs[13]) + (cs[14] + cs[15])));
      }
      // inclusive prefix across the columns: thread totals, warp scan, warp totals
      double inc = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) inc = scan_up(inc, o);
      if (lane == 31) s_ws[wid] = inc;
      const int chg = __syncthreads_or(tchg);  // also: the previous ping's window reads are done
      tchg = 0;
      float vc[kStripCols];
      if (active) load16(p, vc);
      {
        // exclusive offset of this warp: every group of 8 lanes adds up the totals of the warps before it
        const double wt = ((lane & 7) < NW) ? s_ws[lane & 7] : 0.0;
        double ex = ((lane & 7) < wid) ? wt : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) ex += __shfl_xor_sync(0xffffffffu, ex, o);
        const double off = (inc - run) + ex;
        if (active) {
          double acc = off;
          if (nkill == 0) {  // four independent chains of four (a 16-long chain of dependent DADDs is ~150 cycles)
            const double b1 = off + ((cs[0] + cs[1]) + (cs[2] + cs[3])), b2 = b1 + ((cs[4] + cs[5]) + (cs[6] + cs[7])),
                         b3 = b2 + ((cs[8] + cs[9]) + (cs[10] + cs[11]));
            double a0 = off, a1 = b1, a2 = b2, a3 = b3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              a0 += cs[j], a1 += cs[4 + j], a2 += cs[8 + j], a3 += cs[12 + j];
              my_pre[j * TS] = a0, my_pre[(4 + j) * TS] = a1, my_pre[(8 + j) * TS] = a2, my_pre[(12 + j) * TS] = a3;
            }
          } else {  // the first thread: nothing above the column before the sliced axis (mirror zone)
#pragma unroll
            for (int j = 0; j < kStripCols; ++j) {
              acc += cs[j];
              if (j >= nkill - 1) my_pre[j * TS] = acc;
            }
          }
        }
        __syncwarp();  // the warp's own entries are visible to its lanes
        // Mirrored entries: the 32 lanes of a warp copy the entries of the warp's columns among the first / last w of the
        // sliced axis (the owners writing their 16 columns each kept the two edge warps ~30 % behind the others)
        if (mir_lo0 <= mir_lo1) {  // warp-uniform
          const double base2 = 2.0 * s_pre[slot(m0 - 1 - col0)];
          for (int x = mir_lo0 + lane; x <= mir_lo1; x += 32) s_pre[slot(2 * m0 - 2 - x - col0)] = base2 - s_pre[slot(x - col0)];
        }
        if (mir_hi0 <= mir_hi1) {
          double tot2 = 0.0;
          for (int i = 0; i < NW; ++i) tot2 += s_ws[i];
          tot2 *= 2.0;
          for (int x = mir_hi0 + lane; x <= mir_hi1; x += 32) s_pre[slot(2 * R - 2 - x - col0)] = tot2 - s_pre[slot(x - col0)];
        }
      }
      if (chg) {  // CTA-uniform and rare: a deficit changed, rebuild the count prefix with the same scan
        int crun = 0;
        if (active) {
#pragma unroll
          for (int j = 0; j < kStripCols; ++j) crun += my_cc[j * kStripThreads];
        }
        int cinc = crun;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) cinc = scan_up(cinc, o);
        if (lane == 31) s_wc[wid] = cinc;
        __syncthreads();
        const int wtc = ((lane & 7) < NW) ? s_wc[lane & 7] : 0;
        int cex = ((lane & 7) < wid) ? wtc : 0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) cex += __shfl_xor_sync(0xffffffffu, cex, o);
        const int coff = (cinc - crun) + cex;
        int ctot = 0;
        for (int i = 0; i < NW; ++i) ctot += s_wc[i];
        has_def = ctot != 0;
        if (active) {
          int acc = coff;
#pragma unroll
          for (int j = 0; j < kStripCols; ++j) {
            acc += my_cc[j * kStripThreads];
            if (j >= nkill - 1) my_cnt[j * TS] = acc;
          }
        }
        __syncwarp();
        if (mir_lo0 <= mir_lo1) {
          const int base2 = 2 * s_cnt[slot(m0 - 1 - col0)];
          for (int x = mir_lo0 + lane; x <= mir_lo1; x += 32) s_cnt[slot(2 * m0 - 2 - x - col0)] = base2 - s_cnt[slot(x - col0)];
        }
        if (mir_hi0 <= mir_hi1)
          for (int x = mir_hi0 + lane; x <= mir_hi1; x += 32) s_cnt[slot(2 * R - 2 - x - col0)] = 2 * ctot - s_cnt[slot(x - col0)];
      }
      __syncthreads();
      // range windows, pooled Sv, mask: 16 outputs per thread
      unsigned char* mrow = mask + (c * P + p) * (long long)R;
      if (active) {
        float pv[kStripCols];
        // no deficit in any of this thread's windows: one difference of the count prefix over their union
        if (kDepth) {
          // clean/utils.py:78-84: the ping keeps p - k >= 0 and p + k <= P; the window holds pings p - k .. min(p + k, P - 1)
          const bool pvalid = (p - k >= 0) && (p + k <= Pi);
          const int nrow = ((p + k < Pi) ? p + k : Pi - 1) - (p - k) + 1;
          unsigned tw[3][8];  // this thread's 16 entries of the three tables, two 16-bit values per word
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint4 q = *reinterpret_cast<const uint4*>(s_wt + a * R + n0 + 8 * g);
              tw[a][4 * g] = q.x, tw[a][4 * g + 1] = q.y, tw[a][4 * g + 2] = q.z, tw[a][4 * g + 3] = q.w;
            }
#pragma unroll
          for (int j = 0; j < kStripCols; ++j) {
            const int sh = 16 * (j & 1);
            const int lo = (tw[0][j >> 1] >> sh) & 0xffff, hi = (tw[1][j >> 1] >> sh) & 0xffff, wd = (tw[2][j >> 1] >> sh) & 0xffff;
            const int wc = nrow * wd - (has_def ? s_cnt[hi] - s_cnt[lo] : 0);
            float rc;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"((float)wc));
            const float m = d2f_trunc(s_pre[hi] - s_pre[lo]) * rc;
            pv[j] = (pvalid && wd != 0 && wc > 0) ? kLog2ToDb * fast_log2(m) : CUDART_NAN_F;
          }
          if (kPooled) {  // a sample without a depth of its own has no window (its Sv is NaN: the mask is False either way)
            const float4* dq = reinterpret_cast<const float4*>(depth + (c * P + p) * (long long)R + n0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 dv = __ldg(dq + g);
              if (!(dv.x == dv.x)) pv[4 * g] = CUDART_NAN_F;
              if (!(dv.y == dv.y)) pv[4 * g + 1] = CUDART_NAN_F;
              if (!(dv.z == dv.z)) pv[4 * g + 2] = CUDART_NAN_F;
              if (!(dv.w == dv.w)) pv[4 * g + 3] = CUDART_NAN_F;
            }
          }
        }
        const bool plain = !kDepth && (!has_def || my_cnt[((15 + w) & 15) * TS + ((15 + w) >> 4)] == my_cnt[((-w - 1) & 15) * TS + ((-w - 1) >> 4)]);
        if (plain) {
          switch (w & 15) {
            case 0: strip_windows<0>(sph, spl, inv_full, pv); break;
            case 1: strip_windows<1>(sph, spl, inv_full, pv); break;
            case 2: strip_windows<2>(sph, spl, inv_full, pv); break;
            case 3: strip_windows<3>(sph, spl, inv_full, pv); break;
            case 4: strip_windows<4>(sph, spl, inv_full, pv); break;
            case 5: strip_windows<5>(sph, spl, inv_full, pv); break;
            case 6: strip_windows<6>(sph, spl, inv_full, pv); break;
            case 7: strip_windows<7>(sph, spl, inv_full, pv); break;
            case 8: strip_windows<8>(sph, spl, inv_full, pv); break;
            case 9: strip_windows<9>(sph, spl, inv_full, pv); break;
            case 10: strip_windows<10>(sph, spl, inv_full, pv); break;
            case 11: strip_windows<11>(sph, spl, inv_full, pv); break;
            case 12: strip_windows<12>(sph, spl, inv_full, pv); break;
            case 13: strip_windows<13>(sph, spl, inv_full, pv); break;
            case 14: strip_windows<14>(sph, spl, inv_full, pv); break;
            default: strip_windows<15>(sph, spl, inv_full, pv); break;
          }
        } else if (!kDepth) {  // deficits in this thread's windows: run-time offsets, counts from the deficit prefix
#pragma unroll
          for (int j = 0; j < kStripCols; ++j) {
            const int hi = ((j + w) & 15) * TS + ((j + w) >> 4), lo = ((j - w - 1) & 15) * TS + ((j - w - 1) >> 4);
            const int wc = ifull - (my_cnt[hi] - my_cnt[lo]);
            float rc;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"((float)wc));
            const float m = d2f_trunc(my_pre[hi] - my_pre[lo]) * rc;
            pv[j] = wc > 0 ? kLog2ToDb * fast_log2(m) : CUDART_NAN_F;
          }
        }
        unsigned mw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < kStripCols; ++j)  // (the first thread's columns above m0 are rewritten below)
          mw[j >> 2] |= ((vc[j] - pv[j] > thr) ? 1u : 0u) << (8 * (j & 3));
        *reinterpret_cast<uint4*>(mrow + n0) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        if (kPooled) {
          float4* o4 = reinterpret_cast<float4*>(pooled + (c * P + p) * (long long)R + n0);
#pragma unroll
          for (int g = 0; g < 4; ++g) o4[g] = make_float4(pv[4 * g], pv[4 * g + 1], pv[4 * g + 2], pv[4 * g + 3]);
        }
        for (int j = 0; j < nkill; ++j) {  // the first thread's columns above the sliced axis (same thread: ordered stores)
          mrow[n0 + j] = 0;
          if (kPooled) pooled[(c * P + p) * (long long)R + n0 + j] = CUDART_NAN_F;
        }
      }
      // columns above the first thread's: outside the sliced axis
      for (int i = tid; i < (col0 >> 4); i += T) {
        *reinterpret_cast<uint4*>(mrow + 16 * i) = make_uint4(0u, 0u, 0u, 0u);
        if (kPooled) {
          float4* o4 = reinterpret_cast<float4*>(pooled + (c * P + p) * (long long)R + 16 * i);
          o4[0] = o4[1] = o4[2] = o4[3] = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        }
      }
    }
    __syncthreads();  // next strip rewrites s_ws / s_pre / the deficits
  }
}

}  // namespace

extern "C" int epb_range_diff_mean(const float* range_var, double* sum, unsigned long long* count, epb_i64 C, epb_i64 P,
                                   epb_i64 R, void* stream) {
  EPB_REQUIRE(range_var && sum && count && C > 0 && P > 0 && R > 1, "bad pointer/shape");
  if (cudaMemsetAsync(sum, 0, C * sizeof(double), (cudaStream_t)stream) != cudaSuccess ||
      cudaMemsetAsync(count, 0, C * sizeof(unsigned long long), (cudaStream_t)stream) != cudaSuccess)
    return epb_check_launch("epb_range_diff_mean(memset)");
  const long long nrows = C * P, cap = (long long)epb_num_sms() * 8;
  range_diff_kernel<<<(unsigned)(nrows < cap ? nrows : cap), 256, 0, (cudaStream_t)stream>>>(range_var, nrows, P, (int)R, sum,
                                                                                             count);
  return epb_check_launch("epb_range_diff_mean");
}

extern "C" int epb_first_not_le(const float* a, epb_i64 n, float threshold, unsigned long long* out, void* stream) {
  EPB_REQUIRE(a && out && n > 0, "bad pointer/size");
  if (cudaMemsetAsync(out, 0xff, sizeof(unsigned long long), (cudaStream_t)stream) != cudaSuccess)
    return epb_check_launch("epb_first_not_le(memset)");
  const long long blocks = (n + 255) / 256, cap = (long long)epb_num_sms() * 8;
  first_not_le_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(a, n, threshold, out);
  return epb_check_launch("epb_first_not_le");
}

extern "C" int epb_impulse_noise_mask(const float* Sv, const int* nsamp, float* block_means, unsigned char* mask, epb_i64 C,
                                      epb_i64 P, epb_i64 R, int nbmax, int num_side_pings, float threshold, void* stream) {
  EPB_REQUIRE(Sv && nsamp && block_means && mask, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R <= 49152 && nbmax > 0 && num_side_pings >= 1, "bad shape / num_side_pings");
  const long long nrows = C * P, cap = (long long)epb_num_sms() * 8;
  const unsigned grid = (unsigned)(nrows < cap ? nrows : cap);
  const size_t smem = (size_t)R * 4;
  // single pass: rows of up to 4096 samples in 16-byte units, blocks of at least 16 samples (nbmax = max ceil(R / nsamp))
  const size_t fsm = (size_t)R * 4 + (size_t)(2 * num_side_pings + 1) * nbmax * 4;
  if ((R & 15) == 0 && R <= 16 * kImpThreads && nbmax <= R / 16 && fsm <= 64 * 1024 && P < (1LL << 30) && ((uintptr_t)Sv % 16) == 0 &&
      ((uintptr_t)mask % 16) == 0 && !getenv("EPB_IMPULSE_TWO_PASS")) {
    const long long want = ((long long)epb_num_sms() * 16 + C - 1) / C;
    long long chunk = (P + want - 1) / want;
    const long long min_chunk = 16LL * (2 * num_side_pings + 1);
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > P) chunk = P;
    const long long nchunks = (P + chunk - 1) / chunk, nstrips = nchunks * C;
    if (cudaFuncSetAttribute(impulse_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm) != cudaSuccess)
      return epb_check_launch("epb_impulse_noise_mask(smem)");
    impulse_fused_kernel<<<(unsigned)(nstrips < cap ? nstrips : cap), kImpThreads, fsm, (cudaStream_t)stream>>>(
        Sv, nsamp, block_means, mask, P, (int)R, nbmax, num_side_pings, threshold, (int)chunk, (int)nchunks, nstrips);
    return epb_check_launch("epb_impulse_noise_mask(fused)");
  }
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(block_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return epb_check_launch("epb_impulse_noise_mask(smem)");
  {  // 128 threads per row: the CTAs an SM holds are bounded by the 2048 threads before the shared-memory rows (13 rows of 16 KB)
    const char* ev = getenv("EPB_BM_THREADS");
    const int bt = ev ? atoi(ev) : 128;
    const long long capb = (long long)epb_num_sms() * (2048 / bt);
    block_mean_kernel<<<(unsigned)(nrows < capb ? nrows : capb), bt, smem, (cudaStream_t)stream>>>(Sv, nsamp, block_means, nrows, P, (int)R, nbmax);
  }
  if ((R & 15) == 0 && ((uintptr_t)mask % 16) == 0) {
    const long long capw = (long long)epb_num_sms() * 8;
    const int wt = R >= 4096 ? 256 : (int)((R / 16 + 31) / 32 * 32);
    impulse_mask_wide_kernel<<<(unsigned)(nrows < capw ? nrows : capw), wt, 0, (cudaStream_t)stream>>>(
        block_means, nsamp, mask, nrows, P, (int)R, nbmax, num_side_pings, threshold);
  } else {
    impulse_mask_kernel<<<grid, 256, (size_t)((nbmax + 15) & ~15), (cudaStream_t)stream>>>(block_means, nsamp, mask, nrows, P, (int)R, nbmax,
                                                                                         num_side_pings, threshold);
  }
  return epb_check_launch("epb_impulse_noise_mask");
}

extern "C" int epb_transient_noise_mask(const float* Sv, const int* nsamp, float* window_sums /* float2 [C,P,R] */,
                                        unsigned char* mask, float* pooled_Sv, epb_i64 C, epb_i64 P, epb_i64 R,
                                        int min_range_sample, int max_nsamp, int num_side_pings, float threshold,
                                        void* stream) {
  EPB_REQUIRE(Sv && nsamp && mask, "NULL pointer");
  EPB_REQUIRE(max_nsamp >= 1 && (max_nsamp <= R - min_range_sample || min_range_sample == R),
              "range window longer than the sliced range axis (single reflection)");
  EPB_REQUIRE(C > 0 && C < 65536 && P > 0 && R > 0 && num_side_pings >= 0, "bad shape");
  EPB_REQUIRE(min_range_sample >= 0 && min_range_sample <= R, "min_range_sample outside the range axis");
  EPB_REQUIRE(num_side_pings <= P, "num_side_pings must not exceed the number of pings (single reflection)");
  EPB_REQUIRE(((uintptr_t)window_sums % 8) == 0, "window_sums must be 8-byte aligned");
  const int L = (int)R - min_range_sample;
  const size_t smem = (size_t)(L + 1) * 12 + 8;
  EPB_REQUIRE(smem <= 200 * 1024, "range_sample dimension too long for the shared-memory prefix sums");
  // up to 4096 samples past exclude_above (16 columns x 256 threads), 16-byte aligned rows: the single-pass strip kernel
  const int col0 = min_range_sample & ~15;
  const int TL = (max_nsamp + 16) / 16 + 1, TH = (max_nsamp + 15) / 16 + 1;  // mirror zones on both sides
  const int threads = (int)(((R - col0) / kStripCols + 31) / 32 * 32);
  if (R % 16 == 0 && R - col0 <= kStripThreads * kStripCols && L > 0 && ((uintptr_t)Sv % 16) == 0 && ((uintptr_t)mask % 16) == 0 &&
      ((uintptr_t)pooled_Sv % 16) == 0 && (long long)(2 * num_side_pings + 1) * (2 * max_nsamp + 1) < (1LL << 24) && max_nsamp < L && max_nsamp <= 480 &&
      P < (1LL << 30) && TL + threads + TH <= kStripPitch) {
    // strips: ~2 per resident CTA (two CTAs per SM), chunks not shorter than 8 windows (the 2 k pings of warm-up are read twice)
    const long long want = ((long long)epb_num_sms() * 4 + C - 1) / C;
    long long chunk = (P + want - 1) / want;
    const long long min_chunk = 8LL * (2 * num_side_pings + 1);
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > P) chunk = P;
    const long long nchunks = (P + chunk - 1) / chunk;
    const size_t sm = (size_t)kStripCols * kStripPitch * 12 + (size_t)kStripCols * kStripThreads * 4;
    auto kern = pooled_Sv ? transient_strip_kernel<true> : transient_strip_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
      return epb_check_launch("epb_transient_noise_mask(smem)");
    const long long nstrips = nchunks * C, capg = (long long)epb_num_sms() * 2;
    kern<<<(unsigned)(nstrips < capg ? nstrips : capg), threads, sm, (cudaStream_t)stream>>>(
        Sv, nsamp, mask, pooled_Sv, P, (int)R, min_range_sample, num_side_pings, threshold, (int)chunk, (int)nchunks, nstrips, TL,
        nullptr, nullptr);
    return epb_check_launch("epb_transient_noise_mask(strip)");
  }
  EPB_REQUIRE(window_sums, "window_sums scratch is needed for rows longer than 4096 samples");
  const long long nrows = C * P, cap = (long long)epb_num_sms() * 8;
  if (L > 0) {
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(pool_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return epb_check_launch("epb_transient_noise_mask(smem)");
    pool_rows_kernel<<<(unsigned)(nrows < cap ? nrows : cap), 256, smem, (cudaStream_t)stream>>>(
        Sv, nsamp, reinterpret_cast<float2*>(window_sums), nrows, P, (int)R, min_range_sample);
  }
  const int chunk = 256;
  dim3 grid((unsigned)((R + 127) / 128), (unsigned)((P + chunk - 1) / chunk), (unsigned)C);
  EPB_REQUIRE(grid.y < 65536, "too many ping chunks");
  // (capping the resident CTAs so that the second read of a row hits L2 measured slower: 12.1 -> 14.7 ms)
  pool_pings_mask_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(window_sums), Sv, mask,
                                                                 pooled_Sv, P, (int)R, min_range_sample, num_side_pings,
                                                                 threshold, chunk);
  return epb_check_launch("epb_transient_noise_mask");
}

namespace {
__global__ void edges_ceil32_kernel(const double* __restrict__ e, float* __restrict__ t, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = __double2float_ru(e[i]);
}
}  // namespace

extern "C" int epb_impulse_noise_mask_depth(const float* Sv, const float* depth, const double* edges, int nbins,
                                            float* bin_means, int* bin_first, float* upsampled, unsigned char* mask, epb_i64 C,
                                            epb_i64 P, epb_i64 R, int num_side_pings, float threshold, float* thresholds_scratch,
                                            void* stream) {
  EPB_REQUIRE(Sv && depth && edges && bin_means && bin_first && mask && thresholds_scratch, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R <= 24576 && nbins > 0 && nbins <= 8192 && num_side_pings >= 1, "bad shape / argument");
  // float32 thresholds of the float64 edges (closed-left intervals): computed on the device by a tiny kernel
  edges_ceil32_kernel<<<(nbins + 1 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(edges, thresholds_scratch, nbins + 1);
  const long long nrows = C * P, cap = (long long)epb_num_sms() * 8;
  EPB_REQUIRE(nbins <= R, "more depth intervals than range samples");
  // the upsampled array is not asked for: single pass (rows of up to 4096 samples in 16-sample units)
  const size_t fsm = (size_t)R * 8 + (size_t)(nbins + 1) * 4 + (size_t)(2 * num_side_pings + 1) * (2 * nbins + 1 + R / 16) * 4;
  if (!upsampled) {
    EPB_REQUIRE((R & 15) == 0 && R <= 16 * kImpThreads && fsm <= 96 * 1024 && P < (1LL << 30) &&
                    (((uintptr_t)Sv | (uintptr_t)depth | (uintptr_t)mask) % 16) == 0,
                "without the upsampled array: range_sample % 16 == 0, range_sample <= 4096, 16-byte aligned arrays, a window ring of <= 96 KB");
    const long long want = ((long long)epb_num_sms() * 16 + C - 1) / C;
    long long chunk = (P + want - 1) / want;
    const long long min_chunk = 16LL * (2 * num_side_pings + 1);
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > P) chunk = P;
    const long long nchunks = (P + chunk - 1) / chunk, nstrips = nchunks * C;
    if (cudaFuncSetAttribute(impulse_depth_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm) != cudaSuccess)
      return epb_check_launch("epb_impulse_noise_mask_depth(smem)");
    impulse_depth_fused_kernel<<<(unsigned)(nstrips < cap ? nstrips : cap), kImpThreads, fsm, (cudaStream_t)stream>>>(
        Sv, depth, thresholds_scratch, nbins, bin_means, bin_first, mask, P, (int)R, num_side_pings, threshold, (int)chunk, (int)nchunks,
        nstrips);
    return epb_check_launch("epb_impulse_noise_mask_depth(fused)");
  }
  const unsigned grid = (unsigned)(nrows < cap ? nrows : cap);
  const size_t smem_a = (size_t)R * 8 + (size_t)(nbins + 1) * 4;
  if (smem_a > 48 * 1024 &&
      cudaFuncSetAttribute(depth_bin_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a) != cudaSuccess)
    return epb_check_launch("epb_impulse_noise_mask_depth(smem)");
  depth_bin_mean_kernel<<<grid, 256, smem_a, (cudaStream_t)stream>>>(Sv, depth, thresholds_scratch, nbins, bin_means, bin_first,
                                                                     upsampled, nrows, (int)R);
  impulse_mask_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(upsampled, mask, nrows, P, (int)R, num_side_pings, threshold);
  return epb_check_launch("epb_impulse_noise_mask_depth");
}

namespace {
// A channel's depth rows are "uniform" when every ping has the same depth at a column wherever its depth is defined, and
// samples without a depth (NaN: the padding of shorter pings, calibrate/range.py:143-148) carry no Sv either - such
// samples are members of no window in the reference (NaN fails both comparisons) and contribute nothing here (NaN Sv).
// One pass over depth and Sv: column-wise maximum AND minimum of the defined depths (ordered-int atomics; uniform <=>
// they agree in every column) and the Sv condition; the decode kernel compares the two and leaves the reference row.
__device__ __forceinline__ int float_order(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__global__ void __launch_bounds__(256) depth_ref_row_kernel(const float* __restrict__ depth, const float* __restrict__ Sv, long long P,
                                                            int R, int pchunk, int* __restrict__ max_ord, int* __restrict__ min_ord,
                                                            int* __restrict__ mismatch) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const long long c = blockIdx.z;
  const long long p0 = (long long)blockIdx.y * pchunk, p1 = (p0 + pchunk < P) ? p0 + pchunk : P;
  if (n >= R) return;
  float mx = -CUDART_INF_F, mn = CUDART_INF_F;
  bool bad = false;
  for (long long p = p0; p < p1; ++p) {
    const float d = __ldg(depth + (c * P + p) * R + n), v = __ldg(Sv + (c * P + p) * R + n);
    mx = fmaxf(mx, d), mn = fminf(mn, d);  // both skip NaN
    bad |= !(d == d) && (v == v);          // an Sv without a depth
  }
  atomicMax(max_ord + c * R + n, float_order(mx));
  atomicMin(min_ord + c * R + n, float_order(mn));
  if (bad) *mismatch = 1;
}
__global__ void __launch_bounds__(256) depth_ref_decode_kernel(int* __restrict__ ref, const int* __restrict__ min_ord, long long total,
                                                               int* __restrict__ mismatch) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int o = ref[i], q = min_ord[i];
  const float f = __int_as_float(o >= 0 ? o : o ^ 0x7fffffff), g = __int_as_float(q >= 0 ? q : q ^ 0x7fffffff);
  if (f != -CUDART_INF_F && f != g) *mismatch = 1;  // two pings disagree on the depth of this column
  ref[i] = __float_as_int(f == -CUDART_INF_F ? CUDART_NAN_F : f);  // a column without any depth
}
// per (channel, column): the index interval of the samples with |depth - d| <= depth_bin in the channel's depth row
// (clean/utils.py:86-93, float64 comparisons of the float32 depths; NaN depths count as beyond every value) as
// prefix slots of the strip kernel (slot(i0 - 1), slot(i1 - 1)) and its width i1 - i0; width 0 where the column's own
// conditions fail (:78-84: d - bin >= min depth, d + bin <= max depth, d - bin >= exclude_above; NaN fails all)
__global__ void __launch_bounds__(256) depth_window_table_kernel(const float* __restrict__ ref, int R, double dmin,
                                                                 double dmax, double bin, double exclude_above, int TL, int col0,
                                                                 unsigned short* __restrict__ wtab, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long c = i / R;
  const int n = (int)(i - c * R);
  const float* dr = ref + c * (long long)R;  // the channel's reference depth row
  const double d = (double)dr[n];
  int i0 = 0, i1 = 0;
  if ((d - bin >= dmin) && (d + bin <= dmax) && (d - bin >= exclude_above)) {
    const double lo_v = d - bin, hi_v = d + bin;
    int lo = 0, hi = R;  // first sample with depth >= lo_v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const float v = dr[mid];
      if (!(v == v) || (double)v >= lo_v)
        hi = mid;
      else
        lo = mid + 1;
    }
    i0 = lo;
    hi = R;  // first sample with depth > hi_v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const float v = dr[mid];
      if (!(v == v) || (double)v > hi_v)
        hi = mid;
      else
        lo = mid + 1;
    }
    i1 = lo;
  }
  // columns before col0 (the caller's bound: shallower than every window) are not part of the strip: no window there
  if (n < col0 || i0 < col0) i0 = i1 = col0;
  auto slot = [&](int x) { return (x & 15) * kStripPitch + (x >> 4) + TL; };  // x relative to col0
  unsigned short* t = wtab + c * 3 * (long long)R;
  t[n] = (unsigned short)slot(i0 - 1 - col0);
  t[R + n] = (unsigned short)slot(i1 - 1 - col0);
  t[2 * R + n] = (unsigned short)(i1 - i0);
}
}  // namespace

extern "C" int epb_depth_rows_uniform(const float* depth, const float* Sv, float* ref_rows, int* mismatch, epb_i64 C, epb_i64 P,
                                      epb_i64 R, void* stream) {
  EPB_REQUIRE(depth && Sv && ref_rows && mismatch && C > 0 && C < 65536 && P > 0 && R > 0, "bad pointer / shape");
  cudaStream_t s = (cudaStream_t)stream;
  int* mx = reinterpret_cast<int*>(ref_rows);
  int* mn = mx + C * R;  // second half of ref_rows: scratch for the column minima
  if (cudaMemsetAsync(mismatch, 0, sizeof(int), s) != cudaSuccess || cudaMemsetAsync(mx, 0x80, (size_t)(C * R) * 4, s) != cudaSuccess ||
      cudaMemsetAsync(mn, 0x7f, (size_t)(C * R) * 4, s) != cudaSuccess)  // 0x80808080 / 0x7f7f7f7f: below / above every ordered float
    return epb_check_launch("epb_depth_rows_uniform(memset)");
  const int pchunk = 256;
  dim3 g1((unsigned)((R + 255) / 256), (unsigned)((P + pchunk - 1) / pchunk), (unsigned)C);
  EPB_REQUIRE(g1.y < 65536, "too many pings");
  depth_ref_row_kernel<<<g1, 256, 0, s>>>(depth, Sv, P, (int)R, pchunk, mx, mn, mismatch);
  depth_ref_decode_kernel<<<(unsigned)((C * R + 255) / 256), 256, 0, s>>>(mx, mn, C * R, mismatch);
  return epb_check_launch("epb_depth_rows_uniform");
}

extern "C" int epb_transient_noise_mask_depth_uniform(const float* Sv, const float* depth, const float* ref_rows, unsigned short* tables, unsigned char* mask,
                                                      float* pooled_Sv, epb_i64 C, epb_i64 P, epb_i64 R, double depth_min,
                                                      double depth_max, double depth_bin, double exclude_above, int num_side_pings,
                                                      float threshold, epb_i64 first_column, void* stream) {
  EPB_REQUIRE(Sv && depth && ref_rows && tables && mask, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && num_side_pings >= 0, "bad shape / argument");
  EPB_REQUIRE(first_column >= 0 && first_column < R && first_column % 16 == 0, "first_column must be a multiple of 16 inside the row");
  const int col0 = (int)first_column;
  const int threads = (int)(((R - col0) / kStripCols + 31) / 32 * 32), TL = 1;
  if (!(R % 16 == 0 && R - col0 <= kStripThreads * kStripCols && P < (1LL << 30) && ((uintptr_t)Sv % 16) == 0 && ((uintptr_t)mask % 16) == 0 &&
        ((uintptr_t)pooled_Sv % 16) == 0 && ((uintptr_t)tables % 16) == 0 && (long long)(2 * num_side_pings + 1) * R < (1LL << 24) &&
        TL + threads + 1 <= kStripPitch)) {
    epb_set_error("%s: %s", __func__, "needs range_sample % 16 == 0, range_sample <= 4096, 16-byte aligned arrays");
    return EPB_E_UNSUPPORTED;
  }
  const long long totc = C * R;
  depth_window_table_kernel<<<(unsigned)((totc + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ref_rows, (int)R, depth_min, depth_max, depth_bin,
                                                                                               exclude_above, TL, col0, tables, totc);
  const long long want = ((long long)epb_num_sms() * 4 + C - 1) / C;
  long long chunk = (P + want - 1) / want;
  const long long min_chunk = 8LL * (2 * num_side_pings + 1);
  if (chunk < min_chunk) chunk = min_chunk;
  if (chunk > P) chunk = P;
  const long long nchunks = (P + chunk - 1) / chunk;
  const size_t sm = (size_t)kStripCols * kStripPitch * 12 + (size_t)kStripCols * kStripThreads * 4 + (size_t)R * 6;
  auto kern = pooled_Sv ? transient_strip_kernel<true, true> : transient_strip_kernel<false, true>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
    return epb_check_launch("epb_transient_noise_mask_depth_uniform(smem)");
  const long long nstrips = nchunks * C, capg = (long long)epb_num_sms() * 2;
  kern<<<(unsigned)(nstrips < capg ? nstrips : capg), threads, sm, (cudaStream_t)stream>>>(
      Sv, nullptr, mask, pooled_Sv, P, (int)R, col0, num_side_pings, threshold, (int)chunk, (int)nchunks, nstrips, TL, tables, depth);
  return epb_check_launch("epb_transient_noise_mask_depth_uniform");
}

extern "C" int epb_transient_noise_mask_depth(const float* Sv, const float* depth, double* prefix_sums, int* prefix_counts,
                                              unsigned char* mask, float* pooled_Sv, epb_i64 C, epb_i64 P, epb_i64 R,
                                              double depth_min, double depth_max, double depth_bin, double exclude_above,
                                              int num_side_pings, float threshold, void* stream) {
  EPB_REQUIRE(Sv && depth && prefix_sums && prefix_counts && mask, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30) && num_side_pings >= 0, "bad shape / argument");
  const long long nrows = C * P, total = nrows * R;
  const long long cap = (long long)epb_num_sms() * 8;
  const long long gw = (nrows + 7) / 8;
  row_prefix_kernel<<<(unsigned)(gw < cap ? gw : cap), 256, 0, (cudaStream_t)stream>>>(Sv, prefix_sums, prefix_counts, nrows, (int)R);
  const long long gb = (total + 255) / 256;
  pool_depth_mask_kernel<<<(unsigned)(gb < cap * 2 ? gb : cap * 2), 256, 0, (cudaStream_t)stream>>>(
      Sv, depth, prefix_sums, prefix_counts, mask, pooled_Sv, P, (int)R, depth_min, depth_max, depth_bin, exclude_above,
      num_side_pings, threshold, total);
  return epb_check_launch("epb_transient_noise_mask_depth");
}

extern "C" int epb_transient_noise_mask_median(const float* Sv, const int* nsamp, unsigned char* mask, float* pooled_Sv, epb_i64 C,
                                               epb_i64 P, epb_i64 R, int min_range_sample, int max_nsamp, int num_side_pings,
                                               float threshold, void* stream) {
  EPB_REQUIRE(Sv && nsamp && mask, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && num_side_pings >= 0, "bad shape");
  EPB_REQUIRE(min_range_sample >= 0 && min_range_sample <= R, "min_range_sample outside the range axis");
  EPB_REQUIRE(max_nsamp >= 1 && (max_nsamp <= R - min_range_sample || min_range_sample == R),
              "range window longer than the sliced range axis (single reflection)");
  EPB_REQUIRE(num_side_pings <= P, "num_side_pings must not exceed the number of pings (single reflection)");
  const long long total = C * P * R, cap = (long long)epb_num_sms() * 16, gb = (total + 127) / 128;
  transient_median_kernel<<<(unsigned)(gb < cap ? gb : cap), 128, 0, (cudaStream_t)stream>>>(
      Sv, nsamp, mask, pooled_Sv, P, (int)R, min_range_sample, num_side_pings, threshold, total);
  return epb_check_launch("epb_transient_noise_mask_median");
}

extern "C" int epb_transient_noise_mask_depth_median(const float* Sv, const float* depth, unsigned char* mask, float* pooled_Sv,
                                                     epb_i64 C, epb_i64 P, epb_i64 R, double depth_min, double depth_max,
                                                     double depth_bin, double exclude_above, int num_side_pings, float threshold,
                                                     void* stream) {
  EPB_REQUIRE(Sv && depth && mask, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && num_side_pings >= 0, "bad shape");
  const long long total = C * P * R, cap = (long long)epb_num_sms() * 16, gb = (total + 127) / 128;
  transient_median_depth_kernel<<<(unsigned)(gb < cap ? gb : cap), 128, 0, (cudaStream_t)stream>>>(
      Sv, depth, mask, pooled_Sv, P, (int)R, depth_min, depth_max, depth_bin, exclude_above, num_side_pings, threshold, total);
  return epb_check_launch("epb_transient_noise_mask_depth_median");
}
