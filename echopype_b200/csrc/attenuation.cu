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
//                              middle elements are found by a 32-step bitwise radix select - 16-byte shared-memory
//                              reads, one xor-and-compare per key, one barrier per step; the ping's own layer is
//                              selected from a second staged copy.  Windows larger than the staging buffer are
//                              selected straight from global memory.
#include "epb_common.cuh"

namespace {
using namespace epb;

constexpr unsigned kNaNKey = 0xffffffffu;  // sorts above every value (incl. +inf = 0xff800000)
constexpr int kThreads = 512;  // 16 warps: the selection passes are shared-memory latency bound
constexpr int kWarps = kThreads / 32;

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

// CTA-wide reductions.  The scratch lives in shared memory; the slot / parity cursors are per-thread registers that
// advance identically in every thread (all calls sit in CTA-uniform control flow), so ONE barrier per call suffices:
//  * sums: the warp leaders add into cnt[slot]; thread 0 clears cnt[slot + 1] for the next call (its last readers passed
//    the previous barrier, its next writers come after this one); three slots in rotation
//  * min / max: per-warp values in one of two buffers, every thread combines them after the barrier
struct Scratch {
  int cnt[3];
  int cursor;  // compaction write position
  unsigned ext[2][kWarps];
};
struct Cursor {
  int slot = 0, parity = 0;
};
__device__ __forceinline__ int block_sum(int v, Scratch* sc, Cursor& cu) {
  v = __reduce_add_sync(0xffffffffu, v);
  const int nxt = (cu.slot == 2) ? 0 : cu.slot + 1;
  if (threadIdx.x == 0) sc->cnt[nxt] = 0;
  if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd(&sc->cnt[cu.slot], v);
  __syncthreads();
  const int t = sc->cnt[cu.slot];
  cu.slot = nxt;
  return t;
}
template <bool kMax>
__device__ __forceinline__ unsigned block_extreme(unsigned v, Scratch* sc, Cursor& cu) {
  v = kMax ? __reduce_max_sync(0xffffffffu, v) : __reduce_min_sync(0xffffffffu, v);
  unsigned* s = sc->ext[cu.parity];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned t = s[0];
#pragma unroll
  for (int i = 1; i < kWarps; ++i) t = kMax ? max(t, s[i]) : min(t, s[i]);
  cu.parity ^= 1;
  return t;
}

// The middle element(s) of the valid keys `scan` enumerates (every thread of the CTA calls scan(f); f(key) is invoked once
// per element the thread owns).  m = number of valid keys; ka / kb = keys of rank (m-1)/2 and m/2 (np.nanmedian averages
// them).  Returns m; ka, kb undefined when m == 0.  Bitwise radix select: one counting pass per bit.
template <typename Scan>
__device__ int block_median_keys(Scan scan, Scratch* sc, Cursor& cu, unsigned& ka, unsigned& kb) {
  int cnt = 0;
  scan([&](unsigned key) { cnt += (key != kNaNKey) ? 1 : 0; });
  const int m = block_sum(cnt, sc, cu);
  if (m == 0) return 0;
  int k = (m - 1) >> 1;
  unsigned prefix = 0u;
  for (int bit = 31; bit >= 0; --bit) {
    // candidates share the bits above `bit` with the prefix; those with a zero at `bit` differ from the prefix in none of
    // the bits >= bit (the prefix is still zero there).  NaN keys are all ones: never counted, and k < m keeps the
    // selection among the valid keys
    const unsigned test = 0xffffffffu << bit;
    int zeros = 0;
    scan([&](unsigned key) { zeros += (((key ^ prefix) & test) == 0u) ? 1 : 0; });
    const int z = block_sum(zeros, sc, cu);
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
    const int neq = block_sum(eq, sc, cu);
    const unsigned mn = block_extreme<false>(above, sc, cu);
    if (k + 1 >= neq) kb = mn;
  }
  return m;
}

// 10 log10 of the mean of the two middle values in the linear domain (utils/compute.py:13,29; np.nanmedian), float64
__device__ __forceinline__ double median_db(unsigned ka, unsigned kb) {
  const double a = pow(10.0, (double)from_key(ka) / 10.0), b = pow(10.0, (double)from_key(kb) / 10.0);
  return 10.0 * log10((a + b) * 0.5);
}

// keys staged in shared memory, padded with NaN keys to a multiple of 4 * kThreads: every thread reads 16 bytes per step
struct StagedScan {
  const uint4* keys;
  int steps;  // padded length / (4 * kThreads)
  template <typename F>
  __device__ __forceinline__ void operator()(F&& f) const {
#pragma unroll 2
    for (int i = 0; i < steps; ++i) {
      const uint4 v = keys[threadIdx.x + i * kThreads];
      f(v.x), f(v.y), f(v.z), f(v.w);
    }
  }
};
// keys converted on the fly from `rows` rows of `w` samples in global memory (windows beyond the staging buffer)
struct GlobalScan {
  const float* base;
  int rows, w;
  long long pitch;
  template <typename F>
  __device__ __forceinline__ void operator()(F&& f) const {
    for (int q = 0; q < rows; ++q)
      for (int j = threadIdx.x; j < w; j += kThreads) f(to_key(__ldg(base + q * pitch + j)));
  }
};

constexpr int kQuantum = 4 * kThreads;  // staged arrays are NaN-padded to a multiple of this many keys
__device__ __forceinline__ int round_up_quantum(long long n) { return (int)((n + kQuantum - 1) / kQuantum * kQuantum); }

// stage `rows` rows of `w` samples as keys, row after row, NaN keys up to the next multiple of kQuantum; returns the
// number of 16-byte steps per thread.  The caller has made sure that nobody still reads s_keys.
__device__ __forceinline__ int stage_keys(unsigned* s_keys, const float* base, int rows, int w, long long pitch) {
  const int n = rows * w, padded = round_up_quantum(n);
  for (int q = 0; q < rows; ++q)
    for (int j = threadIdx.x; j < w; j += kThreads) s_keys[q * w + j] = to_key(__ldg(base + q * pitch + j));
  for (int e = n + threadIdx.x; e < padded; e += kThreads) s_keys[e] = kNaNKey;
  __syncthreads();
  return padded / kQuantum;
}

// copy the valid keys of `cur` that agree with `prefix` in the bits >= bit to dst (any order), NaN-pad to the quantum.
// `cand` = their number (known from the counting passes).  Warp-aggregated: one shared atomic per warp and group of 32 keys.
__device__ __forceinline__ void compact_keys(const uint4* cur, int steps, unsigned prefix, int bit, unsigned* dst, int cand,
                                             Scratch* sc) {
  if (threadIdx.x == 0) sc->cursor = 0;
  __syncthreads();
  const unsigned test = 0xffffffffu << bit;
  const unsigned lower_lanes = (1u << (threadIdx.x & 31)) - 1u;
  for (int i = 0; i < steps; ++i) {
    const uint4 v = cur[threadIdx.x + i * kThreads];
    const unsigned keys4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned key = keys4[q];
      const bool take = (((key ^ prefix) & test) == 0u) && (key != kNaNKey);
      const unsigned ballot = __ballot_sync(0xffffffffu, take);
      if (ballot != 0u) {  // warp-uniform
        int base = 0;
        if ((threadIdx.x & 31) == 0) base = atomicAdd(&sc->cursor, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (take) dst[base + __popc(ballot & lower_lanes)] = key;
      }
    }
  }
  for (int e = cand + threadIdx.x; e < round_up_quantum(cand); e += kThreads) dst[e] = kNaNKey;  // disjoint from the copies
  __syncthreads();
}

// block_median_keys for keys staged at s_keys[src_off, src_off + steps * kQuantum): the same selection with two short cuts.
//  * the bits that the smallest and the largest valid key share are the median's too: the bit loop starts at the highest
//    bit in which they differ (Sv values of one layer share sign and exponent: nine passes less);
//  * once the candidates (keys agreeing with the prefix so far) are a quarter of the array being scanned they are copied
//    to the scratch region [scr_off, cap) and the remaining passes scan only them; again when they are a quarter of that.
// The staged array itself stays intact (the even-count step needs the smallest key ABOVE the median, which need not be a
// candidate).
__device__ int staged_median_keys(unsigned* s_keys, int src_off, int steps, int scr_off, int cap, Scratch* sc, Cursor& cu,
                                  unsigned& ka, unsigned& kb) {
  const uint4* src = reinterpret_cast<const uint4*>(s_keys + src_off);
  int cnt = 0;
  unsigned lo = kNaNKey, hi = 0u;
  StagedScan{src, steps}([&](unsigned key) {
    const bool ok = key != kNaNKey;
    cnt += ok ? 1 : 0;
    lo = min(lo, key);
    hi = max(hi, ok ? key : 0u);
  });
  const int m = block_sum(cnt, sc, cu);
  if (m == 0) return 0;
  lo = block_extreme<false>(lo, sc, cu);
  hi = block_extreme<true>(hi, sc, cu);
  int k = (m - 1) >> 1;
  unsigned prefix = lo;
  int bit = -1;  // all valid keys equal: nothing to select
  if (lo != hi) {
    bit = 31 - __clz(lo ^ hi);
    prefix = lo & ~((2u << bit) - 1u);  // the shared bits above `bit` (2u << 31 == 0: no shared bit, prefix 0)
  }
  const uint4* cur = src;
  int cur_steps = steps, cand = m, free_off = scr_off;
  for (; bit >= 0; --bit) {
    const unsigned test = 0xffffffffu << bit;
    int zeros = 0;
    StagedScan{cur, cur_steps}([&](unsigned key) { zeros += (((key ^ prefix) & test) == 0u) ? 1 : 0; });
    const int z = block_sum(zeros, sc, cu);
    if (k >= z) {
      prefix |= 1u << bit;
      k -= z;
      cand -= z;
    } else {
      cand = z;
    }
    if (bit > 0 && cur_steps > 1 && cand * 4 <= cur_steps * kQuantum && free_off + round_up_quantum(cand) <= cap) {
      compact_keys(cur, cur_steps, prefix, bit, s_keys + free_off, cand, sc);
      cur = reinterpret_cast<const uint4*>(s_keys + free_off);
      cur_steps = round_up_quantum(cand) / kQuantum;
      free_off += cur_steps * kQuantum;
    }
  }
  ka = kb = prefix;
  if (!(m & 1)) {
    // copies of the median value are all candidates (count them in `cur`); the smallest key above it is looked for in
    // the whole staged array
    int eq = 0;
    unsigned above = kNaNKey;
    StagedScan{cur, cur_steps}([&](unsigned key) { eq += (key == prefix) ? 1 : 0; });
    StagedScan{src, steps}([&](unsigned key) {
      if (key > prefix && key != kNaNKey) above = min(above, key);
    });
    const int neq = block_sum(eq, sc, cu);
    const unsigned mn = block_extreme<false>(above, sc, cu);
    if (k + 1 >= neq) kb = mn;
  }
  return m;
}

__global__ void __launch_bounds__(kThreads) attenuated_ping_kernel(const float* __restrict__ Sv, const int* __restrict__ limits,
                                                                   unsigned char* __restrict__ mask, long long nrows, long long P,
                                                                   int R, int n_side, double thr, int cap_keys) {
  extern __shared__ __align__(16) unsigned s_keys[];
  __shared__ Scratch s_scratch;
  Scratch* sc = &s_scratch;
  Cursor cu;
  const int tid = threadIdx.x;
  if (tid < 3) sc->cnt[tid] = 0;
  __syncthreads();
  const bool wide = ((R & 15) == 0) && ((reinterpret_cast<uintptr_t>(mask) & 15) == 0);
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const long long p = row % P;
    const int up = limits[2 * row], lw = limits[2 * row + 1];
    const int w = lw - up;  // Sv[p, up:lw] is empty when lw <= up: np.all(np.isnan(empty)) is True, the ping is skipped
    bool flag = false;
    // n_side == 0: the block Sv[p:p] is empty, its median NaN, the comparison False
    if (w > 0 && n_side > 0 && p - n_side >= 0 && p + n_side <= P - 1) {
      unsigned pa = 0, pb = 0, ba = 0, bb = 0;
      int mp, mb = 0;
      const float* prow = Sv + row * (long long)R + up;
      const float* brow = prow - n_side * (long long)R;
      const int rows = 2 * n_side;
      const long long nkeys = (long long)rows * w;
      __syncthreads();  // the previous ping's passes have finished reading s_keys
      if (nkeys <= cap_keys && round_up_quantum(nkeys) <= cap_keys) {
        // the whole block in shared memory; the ping's own row is row n_side of it, selected from a second, padded copy
        // behind the block when that fits as well, else from global memory
        const int bsteps = stage_keys(s_keys, brow, rows, w, R);
        const int boff = bsteps * kQuantum;
        if (boff + round_up_quantum(w) <= cap_keys) {
          const int psteps = stage_keys(s_keys + boff, prow, 1, w, R);
          mp = staged_median_keys(s_keys, boff, psteps, boff + psteps * kQuantum, cap_keys, sc, cu, pa, pb);
        } else {
          mp = block_median_keys(GlobalScan{prow, 1, w, R}, sc, cu, pa, pb);
        }
        // the ping's copy is dead now: the block's candidates may be compacted over it
        if (mp > 0) mb = staged_median_keys(s_keys, 0, bsteps, boff, cap_keys, sc, cu, ba, bb);
      } else {
        if (w <= cap_keys && round_up_quantum(w) <= cap_keys) {
          const int psteps = stage_keys(s_keys, prow, 1, w, R);
          mp = staged_median_keys(s_keys, 0, psteps, psteps * kQuantum, cap_keys, sc, cu, pa, pb);
        } else {
          mp = block_median_keys(GlobalScan{prow, 1, w, R}, sc, cu, pa, pb);
        }
        if (mp > 0) mb = block_median_keys(GlobalScan{brow, rows, w, R}, sc, cu, ba, bb);
      }
      if (mp > 0 && mb > 0) flag = (median_db(pa, pb) - median_db(ba, bb)) < thr;  // NaN (-inf - -inf) compares False
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
