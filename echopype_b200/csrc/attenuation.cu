// clean.mask_attenuated_signal (Ryan et al. 2015): the third noise mask of the reference's clean module.
// Reference: echopype/clean/api.py:269-359 (arguments, the "outside the searching range" early return) and
// clean/utils.py:337-377 (echopy_attenuated_signal_mask): per channel and ping
//   up, lw = argmin |range - upper_limit_sl|, argmin |range - lower_limit_sl|          (np.argmin: first minimum, a NaN wins)
//   ping   = nanmedian of 10^(Sv/10) over Sv[p, up:lw]
//   block  = nanmedian of 10^(Sv/10) over Sv[p - n : p + n, up:lw]                     (2 n pings, p + n excluded)
//   the WHOLE ping is masked when 10 log10(ping) - 10 log10(block) < threshold; pings closer than n to either end of the
//   ping axis and pings whose window is empty or all-NaN are never masked.
// 10^(x/10) is monotone, so both medians are SELECTIONS on the float32 Sv values themselves (order-preserving integer
// keys); only the one or two middle values are taken to the linear domain, in float64 like the reference.
//   attenuated_limits_kernel : one warp per (channel, ping) row, one pass over the range variable (4 B per sample)
//   attenuated_ping_kernel   : one CTA per (channel, ping); the 2 n x (lw - up) window is staged once into shared
//                              memory as keys (it is read from L2: neighbouring pings share all but one row) and the
//                              middle elements are found by a 32-step bitwise radix select with one barrier per step;
//                              windows larger than the staging buffer are selected straight from global memory.
#include "epb_common.cuh"

namespace {
using namespace epb;

constexpr unsigned kNaNKey = 0xffffffffu;  // sorts above every value (incl. +inf = 0xff800000)
constexpr int kThreads = 256;

__device__ __forceinline__ unsigned to_key(float v) {
  if (!(v == v)) return kNaNKey;
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_key(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }

// np.argmin(abs(range_var[p, :] - limit)) for both limits in one pass, float64 like the reference
__global__ void __launch_bounds__(256) attenuated_limits_kernel(const float* __restrict__ rng, long long nrows, int R, double upper,
                                                                double lower, int* __restrict__ limits) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
  for (long long row = warp0; row < nrows; row += nwarps) {
    const float* r = rng + row * (long long)R;
    double bu = CUDART_INF, bl = CUDART_INF;  // best distances; the lane's indices increase, so "<" keeps the first one
    int iu = 0x7fffffff, il = 0x7fffffff, inan = 0x7fffffff;
#pragma unroll 4
    for (int j = lane; j < R; j += 32) {
      const float v = ld_stream(r + j);
      if (!(v == v)) {
        inan = min(inan, j);
      } else {
        const double du = fabs((double)v - upper), dl = fabs((double)v - lower);
        if (du < bu || iu == 0x7fffffff) bu = du, iu = j;
        if (dl < bl || il == 0x7fffffff) bl = dl, il = j;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ou = __shfl_xor_sync(0xffffffffu, bu, o), ol = __shfl_xor_sync(0xffffffffu, bl, o);
      const int oiu = __shfl_xor_sync(0xffffffffu, iu, o), oil = __shfl_xor_sync(0xffffffffu, il, o);
      inan = min(inan, __shfl_xor_sync(0xffffffffu, inan, o));
      // lexicographic (distance, index); a lane without any sample carries index INT_MAX and loses every tie
      if (oiu != 0x7fffffff && (iu == 0x7fffffff || ou < bu || (ou == bu && oiu < iu))) bu = ou, iu = oiu;
      if (oil != 0x7fffffff && (il == 0x7fffffff || ol < bl || (ol == bl && oil < il))) bl = ol, il = oil;
    }
    if (lane == 0) {
      // np.argmin returns the first NaN when there is one (NaN-padded pings: up == lw, an empty window)
      limits[2 * row] = (inan != 0x7fffffff) ? inan : iu;
      limits[2 * row + 1] = (inan != 0x7fffffff) ? inan : il;
    }
  }
}

// sum over the CTA; s_red is [2][8], `parity` alternates between consecutive calls so that one barrier per call suffices
__device__ __forceinline__ int block_sum(int v, int* s_red, int parity) {
  v = __reduce_add_sync(0xffffffffu, v);
  int* s = s_red + 8 * parity;
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += s[i];
  return t;
}
__device__ __forceinline__ unsigned block_min(unsigned v, int* s_red, int parity) {
  v = __reduce_min_sync(0xffffffffu, v);
  unsigned* s = reinterpret_cast<unsigned*>(s_red) + 8 * parity;
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned t = kNaNKey;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t = min(t, s[i]);
  return t;
}

// The middle element(s) of the valid keys `scan` enumerates (every thread of the CTA calls scan(f); f(key) is invoked once
// per element the thread owns).  m = number of valid keys; ka / kb = keys of rank (m-1)/2 and m/2 (np.nanmedian averages
// them).  Returns m; ka, kb undefined when m == 0.  `parity` is the running barrier parity of block_sum.
template <typename Scan>
__device__ int block_median_keys(Scan scan, int* s_red, int& parity, unsigned& ka, unsigned& kb) {
  int cnt = 0;
  scan([&](unsigned key) { cnt += (key != kNaNKey) ? 1 : 0; });
  const int m = block_sum(cnt, s_red, parity);
  parity ^= 1;
  if (m == 0) return 0;
  int k = (m - 1) >> 1;
  unsigned prefix = 0u;
  for (int bit = 31; bit >= 0; --bit) {
    const unsigned hi_mask = (bit == 31) ? 0u : (0xffffffffu << (bit + 1));
    int zeros = 0;
    // NaN keys are all ones: never counted as a zero bit, and k < m keeps the selection among the valid keys
    scan([&](unsigned key) { zeros += (((key & hi_mask) == prefix) && !((key >> bit) & 1u)) ? 1 : 0; });
    const int z = block_sum(zeros, s_red, parity);
    parity ^= 1;
    if (k >= z) {
      prefix |= 1u << bit;
      k -= z;
    }
  }
  ka = kb = prefix;
  if (!(m & 1)) {
    // rank m/2 is the next element: the same value when more copies of it remain (k = rank among the copies),
    // else the smallest key above it
    int eq = 0;
    unsigned above = kNaNKey;
    scan([&](unsigned key) {
      eq += (key == prefix) ? 1 : 0;
      if (key > prefix && key != kNaNKey) above = min(above, key);
    });
    const int neq = block_sum(eq, s_red, parity);
    parity ^= 1;
    const unsigned mn = block_min(above, s_red, parity);
    parity ^= 1;
    if (k + 1 >= neq) kb = mn;
  }
  return m;
}

// 10 log10 of the mean of the two middle values in the linear domain (utils/compute.py:13,29; np.nanmedian), float64
__device__ __forceinline__ double median_db(unsigned ka, unsigned kb) {
  const double a = pow(10.0, (double)from_key(ka) / 10.0), b = pow(10.0, (double)from_key(kb) / 10.0);
  return 10.0 * log10((a + b) * 0.5);
}

__global__ void __launch_bounds__(kThreads) attenuated_ping_kernel(const float* __restrict__ Sv, const int* __restrict__ limits,
                                                                   unsigned char* __restrict__ mask, long long nrows, long long P,
                                                                   int R, int n_side, double thr, int cap_keys) {
  extern __shared__ unsigned s_keys[];
  __shared__ int s_red[16];
  const int tid = threadIdx.x;
  int parity = 0;
  const bool wide = ((R & 15) == 0) && ((reinterpret_cast<uintptr_t>(mask) & 15) == 0);
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const long long p = row % P;
    const int up = limits[2 * row], lw = limits[2 * row + 1];
    const int w = lw - up;  // Sv[p, up:lw] is empty when lw <= up: np.all(np.isnan(empty)) is True, the ping is skipped
    bool flag = false;
    if (w > 0 && p - n_side >= 0 && p + n_side <= P - 1) {
      unsigned pa, pb, ba, bb;
      const float* prow = Sv + row * (long long)R + up;
      const int mp = block_median_keys(
          [&](auto&& f) {
            for (int j = tid; j < w; j += kThreads) f(to_key(__ldg(prow + j)));
          },
          s_red, parity, pa, pb);
      if (mp > 0 && n_side > 0) {  // n_side == 0: the block Sv[p:p] is empty, its median NaN, the comparison False
        const float* brow = Sv + (row - n_side) * (long long)R + up;
        const int rows = 2 * n_side;
        const long long nkeys = (long long)rows * w;
        int mb;
        if (nkeys <= cap_keys) {
          __syncthreads();  // the previous row's selection has finished reading s_keys
          for (int q = 0; q < rows; ++q)
            for (int j = tid; j < w; j += kThreads) s_keys[q * w + j] = to_key(__ldg(brow + q * (long long)R + j));
          __syncthreads();
          const int n = (int)nkeys;
          mb = block_median_keys(
              [&](auto&& f) {
                for (int e = tid; e < n; e += kThreads) f(s_keys[e]);
              },
              s_red, parity, ba, bb);
        } else {
          mb = block_median_keys(
              [&](auto&& f) {
                for (int q = 0; q < rows; ++q)
                  for (int j = tid; j < w; j += kThreads) f(to_key(__ldg(brow + q * (long long)R + j)));
              },
              s_red, parity, ba, bb);
        }
        if (mb > 0) flag = (median_db(pa, pb) - median_db(ba, bb)) < thr;  // NaN (-inf - -inf) compares False
      }
    }
    unsigned char* out = mask + row * (long long)R;
    if (wide) {
      const unsigned v = flag ? 0x01010101u : 0u;
      uint4* o4 = reinterpret_cast<uint4*>(out);
      for (int j = tid; j < (R >> 4); j += kThreads) o4[j] = make_uint4(v, v, v, v);
    } else {
      for (int j = tid; j < R; j += kThreads) out[j] = flag ? 1 : 0;
    }
  }
}

}  // namespace

extern "C" int epb_attenuated_signal_mask(const float* Sv, const float* range_var, int* limits, unsigned char* mask, epb_i64 C,
                                          epb_i64 P, epb_i64 R, double upper_limit_sl, double lower_limit_sl, int num_side_pings,
                                          double threshold, void* stream) {
  EPB_REQUIRE(Sv && range_var && limits && mask, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30), "bad shape");
  EPB_REQUIRE(num_side_pings >= 0 && num_side_pings < (1 << 20), "num_side_pings out of range");
  const long long nrows = C * P;
  const long long sms = epb_num_sms();
  const long long gw = (nrows + 7) / 8, capw = sms * 8;
  attenuated_limits_kernel<<<(unsigned)(gw < capw ? gw : capw), 256, 0, (cudaStream_t)stream>>>(range_var, nrows, (int)R,
                                                                                               upper_limit_sl, lower_limit_sl, limits);
  // staging buffer for the keys of the ping block: 96 KB (24 576 keys; the default 30 pings x 100 m of 0.19 m samples
  // need 15 800), two CTAs per SM
  constexpr int kCapKeys = 24576;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attenuated_ping_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCapKeys * 4) != cudaSuccess) {
      epb_set_error("epb_attenuated_signal_mask: cannot reserve %d bytes of shared memory", kCapKeys * 4);
      return EPB_E_CUDA;
    }
    attr_set = true;
  }
  const long long capb = sms * 2;
  attenuated_ping_kernel<<<(unsigned)(nrows < capb ? nrows : capb), kThreads, kCapKeys * 4, (cudaStream_t)stream>>>(
      Sv, limits, mask, nrows, P, (int)R, num_side_pings, threshold, kCapKeys);
  return epb_check_launch("epb_attenuated_signal_mask");
}
