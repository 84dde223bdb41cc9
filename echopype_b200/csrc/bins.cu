// K6: linear-domain bin reduction for MVBS / NASC (commongrid/utils.py:504-628 = flox xarray_reduce with
// IntervalIndex groups; compute_raw_NASC :97-207) and its finalisation (mean -> dB).
//
// One warp per (channel, ping) row; lanes stream float4 (coalesced 512 B per warp instruction).  Each sample
// gets a key = range-bin index (or -1: outside every bin, NaN range, ping outside the x bins).  Consecutive
// samples mostly share a key, so the warp performs a SEGMENTED reduction over runs of equal keys with
// shuffles (head flags -> run ids -> 5-step segmented scan) and only the tail lane of each run issues the
// float64 atomics: ~2 runs per 128 samples instead of 128 atomics.
//
// Two ways to obtain the key:
//   generic  - the range variable is an array (float32 / float64): position against the float64 edges
//              with the reference's interval rule (closed left: e[k] <= x < e[k+1]);
//   law      - the range variable was produced by this library (echo_range / depth of compute_Sv): bin
//              boundaries are found in SAMPLE-INDEX space by bisection on the exact float64 range law, so
//              membership is bit-identical with binning the reference's float64 echo_range.
#include "epb_common.cuh"

namespace {
using namespace epb;

constexpr int kWarpsPerCta = 8;
constexpr int kMaxLawBins = 511;  // bounds per warp kept in shared memory

struct Acc3 {
  float sum;   // sum of 10^(Sv/10) over non-NaN members
  int cnt;     // low 16 bits: non-NaN members, high 16 bits: NaN members
  float h;     // NASC: sum of depth differences
};

// position x against edges[0..nR]; returns bin index or -1
__device__ __forceinline__ int bin_of(double x, const double* __restrict__ e, int nR, double inv_w, int closed_right) {
  if (!(x == x)) return -1;
  const double lo = e[0], hi = e[nR];
  if (closed_right ? !(x > lo && x <= hi) : !(x >= lo && x < hi)) return -1;
  int k = (int)((x - lo) * inv_w);  // uniform-edge guess, then exact fix-up against the stored edges
  k = k < 0 ? 0 : (k > nR - 1 ? nR - 1 : k);
  if (closed_right) {
    while (k > 0 && !(x > e[k])) --k;
    while (k < nR - 1 && x > e[k + 1]) ++k;
  } else {
    while (k > 0 && x < e[k]) --k;
    while (k < nR - 1 && x >= e[k + 1]) ++k;
  }
  return k;
}

// float32 range variables: x >= e (float64 edge) <=> x >= ceil32(e) and x > e <=> x > floor32(e) for every float32 x, so
// the interval rule of the reference can be decided with float32 compares against thresholds rounded toward the side
// of the rule (t[k] = ceil32(e[k]) for closed left, floor32(e[k]) for closed right): same keys, no float64 per sample.
__device__ __forceinline__ int bin_of_f32(float x, const float* __restrict__ t, int nR, float inv_w, int closed_right) {
  if (!(x == x)) return -1;
  const float lo = t[0], hi = t[nR];
  if (closed_right ? !(x > lo && x <= hi) : !(x >= lo && x < hi)) return -1;
  int k = (int)((x - lo) * inv_w);
  k = k < 0 ? 0 : (k > nR - 1 ? nR - 1 : k);
  if (closed_right) {
    while (k > 0 && !(x > t[k])) --k;
    while (k < nR - 1 && x > t[k + 1]) ++k;
  } else {
    while (k > 0 && x < t[k]) --k;
    while (k < nR - 1 && x >= t[k + 1]) ++k;
  }
  return k;
}

__device__ __forceinline__ void flush_run(double* __restrict__ cell, const Acc3& a, bool with_h) {
  const int good = a.cnt & 0xffff, bad = a.cnt >> 16;
  if (good) {
    atomicAdd(cell + 0, (double)a.sum);
    atomicAdd(cell + 1, (double)good);
  }
  if (bad) atomicAdd(cell + 2, (double)bad);
  if (with_h && a.h != 0.f) atomicAdd(cell + 3, (double)a.h);
}

// Segmented warp reduction over runs of equal keys; the last lane of each run adds the run total to acc.
template <bool kHeight>
__device__ __forceinline__ void warp_runs_to_acc(int key, Acc3 a, double* __restrict__ acc_row /* [nR][4] */) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync(full, key, 1);
  const bool head = (lane == 0) || (prev != key);
  const unsigned heads = __ballot_sync(full, head);
  if (heads == 1u && key < 0) return;  // whole warp outside every bin
  const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float s = __shfl_up_sync(full, a.sum, d);
    const int c = __shfl_up_sync(full, a.cnt, d);
    const float h = kHeight ? __shfl_up_sync(full, a.h, d) : 0.f;
    const int r = __shfl_up_sync(full, run, d);
    if (lane >= d && r == run) {
      a.sum += s;
      a.cnt += c;
      if (kHeight) a.h += h;
    }
  }
  const int next = __shfl_down_sync(full, run, 1);
  const bool tail = (lane == 31) || (next != run);
  if (tail && key >= 0) flush_run(acc_row + 4 * (long long)key, a, kHeight);
}

__device__ __forceinline__ void add_sample(Acc3& a, float sv) {
  const float v = fast_exp2(sv * kDb2Log2);  // commongrid/utils.py:592  10^(Sv/10)
  const bool ok = (v == v);
  a.sum += ok ? v : 0.f;
  a.cnt += ok ? 1 : (1 << 16);
}

// ---- generic: range variable given as an array ----------------------------------------------------------
template <typename RT, bool kHeight>
__global__ void __launch_bounds__(32 * kWarpsPerCta) bin_reduce_kernel(const float* __restrict__ Sv,
                                                                       const RT* __restrict__ rng,
                                                                       const int* __restrict__ xbin,
                                                                       const double* __restrict__ edges, int nR,
                                                                       int closed_right, double* __restrict__ acc,
                                                                       long long C, long long P, int R, long long nX) {
  extern __shared__ double s_edges[];
  constexpr bool kF32 = sizeof(RT) == 4;
  float* s_t32 = reinterpret_cast<float*>(s_edges + (nR + 1));  // [nR+1] float32 thresholds (kF32)
  for (int k = threadIdx.x; k <= nR; k += blockDim.x) {
    s_edges[k] = edges[k];
    if (kF32) s_t32[k] = closed_right ? __double2float_rd(edges[k]) : __double2float_ru(edges[k]);
  }
  __syncthreads();
  const double inv_w = (double)nR / (s_edges[nR] - s_edges[0]);
  const float inv_w32 = (float)inv_w;
  const int lane = threadIdx.x & 31;
  const long long nrows = C * P;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const bool vec = kF32 && (R & 3) == 0;  // rows start 16-byte aligned: one float4 of Sv and of range per lane
  for (long long row = warp0; row < nrows; row += (long long)gridDim.x * kWarpsPerCta) {
    const long long c = row / P, p = row % P;
    const int xb = __ldg(xbin + p);
    if (xb < 0 || xb >= nX) continue;  // ping outside every x bin: no member
    double* acc_row = acc + ((c * nX + xb) * (long long)nR) * 4;
    const float* sv = Sv + row * (long long)R;
    const RT* rr = rng + row * (long long)R;
    for (int n0 = 0; n0 < R; n0 += 128) {
      // lane owns samples n0 + 4*lane .. +3 ; keys per sample, merged per lane when all four agree
      int keys[4];
      Acc3 a[4];
      float xv[4] = {0.f, 0.f, 0.f, 0.f}, sv4[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec && n0 + 4 * lane < R) {
        const float4 q = ld_stream4(reinterpret_cast<const float4*>(rr + n0 + 4 * lane));
        const float4 w = ld_stream4(reinterpret_cast<const float4*>(sv + n0 + 4 * lane));
        xv[0] = q.x, xv[1] = q.y, xv[2] = q.z, xv[3] = q.w;
        sv4[0] = w.x, sv4[1] = w.y, sv4[2] = w.z, sv4[3] = w.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int n = n0 + 4 * lane + k;
        keys[k] = -1;
        a[k].sum = 0.f, a[k].cnt = 0, a[k].h = 0.f;
        if (n < R) {
          const double x = vec ? (double)xv[k] : (double)rr[n];
          keys[k] = kF32 ? bin_of_f32(vec ? xv[k] : (float)rr[n], s_t32, nR, inv_w32, closed_right)
                         : bin_of(x, s_edges, nR, inv_w, closed_right);
          if (keys[k] >= 0) {
            add_sample(a[k], vec ? sv4[k] : ld_stream(sv + n));
            if (kHeight && n + 1 < R) {
              const double d = (double)rr[n + 1] - x;  // diff(label="lower") commongrid/utils.py:170-172
              if (d == d) a[k].h = (float)d;
            }
          }
        }
      }
      const bool same = (keys[0] == keys[1]) && (keys[1] == keys[2]) && (keys[2] == keys[3]);
      int key = -1;
      Acc3 m;
      m.sum = 0.f, m.cnt = 0, m.h = 0.f;
      if (same) {
        key = keys[0];
        m.sum = (a[0].sum + a[1].sum) + (a[2].sum + a[3].sum);
        m.cnt = a[0].cnt + a[1].cnt + a[2].cnt + a[3].cnt;
        m.h = (a[0].h + a[1].h) + (a[2].h + a[3].h);
      } else {  // a bin boundary inside this lane's four samples (rare): flush them directly
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (keys[k] >= 0) flush_run(acc_row + 4 * (long long)keys[k], a[k], kHeight);
      }
      warp_runs_to_acc<kHeight>(key, m, acc_row);
    }
  }
}

// ---- generic, staged: CTA per run of consecutive rows, rows staged through shared memory ----------------------------------
// The warp-per-row kernel above issues two 16-byte loads per lane and then runs a 5-step segmented scan per 128
// samples: 25 % DRAM utilisation, latency / instruction bound.  This kernel serves the same contract for rows of
// R % 4 == 0 samples and <= kStagedMaxBins range bins:
//   * a persistent CTA owns a contiguous run of (channel, ping) rows; while row i is reduced, row i + 1 streams into the
//     other half of a double buffer with cp.async (16-byte chunks, every byte of the row in flight at once);
//   * a thread reduces a CONTIGUOUS chunk of the row (R / 256 samples, read back with conflict-free swizzled LDS.128):
//     the range variable of neighbouring samples almost always falls into the same bin, so the bin of a sample is found
//     from the bin of its predecessor (two compares) and a run of equal bins is summed in registers;
//   * a thread sees the same chunk of every row, i.e. (for any sensible range variable) the same one or two bins row after
//     row: run totals are merged into a two-entry register cache {bin, sum, counts} that lives ACROSS rows; consecutive
//     pings of a channel share the ping bin, and only when the (channel, ping bin) cell changes the caches go to per-CTA
//     bins in shared memory (shared atomics) and from there to the float64 grid: one global atomic triple per
//     (CTA, cell, range bin) instead of one per (row, run).  (Shared atomics per row measured 6.5 ms: 256 threads on 40 bins.)
// No monotonicity is assumed: an arbitrary range array only makes the runs shorter.
constexpr int kStagedThreads = 256;
constexpr int kStagedMaxBins = 2048;
constexpr int kStages = 3;  // rows in flight per CTA (cp.async ring)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// 16-byte chunk swizzle: a thread's consecutive chunks and its neighbours' land in different bank groups
__device__ __forceinline__ int swz(int chunk) { return chunk ^ ((chunk >> 3) & 7); }

template <typename RT>
__device__ __forceinline__ int locate_bin(RT x, const RT* __restrict__ t, int nR, RT inv_w, int closed_right, int guess) {
  if (!(x == x)) return -1;
  const RT lo = t[0], hi = t[nR];
  if (closed_right ? !(x > lo && x <= hi) : !(x >= lo && x < hi)) return -1;
  int k = guess;
  if (k < 0) {
    k = (int)((x - lo) * inv_w);
    k = k < 0 ? 0 : (k > nR - 1 ? nR - 1 : k);
  }
  if (closed_right) {
    while (k > 0 && !(x > t[k])) --k;
    while (k < nR - 1 && x > t[k + 1]) ++k;
  } else {
    while (k > 0 && x < t[k]) --k;
    while (k < nR - 1 && x >= t[k + 1]) ++k;
  }
  return k;
}

template <typename RT, bool kHeight>
__global__ void __launch_bounds__(kStagedThreads) bin_reduce_staged_kernel(const float* __restrict__ Sv, const RT* __restrict__ rng,
                                                                           const int* __restrict__ xbin,
                                                                           const double* __restrict__ edges, int nR,
                                                                           int closed_right, double* __restrict__ acc, long long C,
                                                                           long long P, int R, long long nX) {
  // shared memory, addressed in 16-byte chunks from one typed base (no integer round trips: the compiler must keep the
  // shared address space, a generic LD here measured 4x slower): [stage buffers kStages x (Sv row | range row)]
  // [thresholds nR + 1][sum nR][good nR][bad nR][h nR]
  extern __shared__ float4 smem4[];
  constexpr int kRC = 16 / (int)sizeof(RT);  // range samples per 16-byte chunk
  const int svc = ((R / 4 + 7) & ~7), rgc = ((R / kRC + 7) & ~7);  // chunks per staged row, padded to the swizzle period
  const int buf_chunks = svc + rgc;
  RT* s_thr = reinterpret_cast<RT*>(smem4 + kStages * buf_chunks);  // [nR + 1] thresholds in the range type
  float* s_sum = reinterpret_cast<float*>(smem4 + kStages * buf_chunks + ((nR + 1) * (int)sizeof(RT) + 15) / 16);  // [nR]
  int* s_good = reinterpret_cast<int*>(s_sum + nR);
  int* s_bad = s_good + nR;
  float* s_h = reinterpret_cast<float*>(s_bad + nR);  // [nR] (kHeight)
  const int tid = threadIdx.x;
  for (int k = tid; k <= nR; k += kStagedThreads) {
    // float32 ranges: x >= e <=> x >= ceil32(e), x > e <=> x > floor32(e) for every float32 x (see bin_of_f32)
    if (sizeof(RT) == 4)
      s_thr[k] = (RT)(closed_right ? __double2float_rd(edges[k]) : __double2float_ru(edges[k]));
    else
      s_thr[k] = (RT)edges[k];
  }
  for (int k = tid; k < nR; k += kStagedThreads) {
    s_sum[k] = 0.f, s_good[k] = 0, s_bad[k] = 0;
    if (kHeight) s_h[k] = 0.f;
  }
  const long long nrows = C * P;
  const long long ra = nrows * (long long)blockIdx.x / gridDim.x, rb = nrows * (long long)(blockIdx.x + 1) / gridDim.x;
  const int nloc = (int)(rb - ra);  // rows of this CTA
  auto stage = [&](int i, int slot) {  // local row i -> buffer slot
    float4* dst = smem4 + slot * buf_chunks;
    const float4* gs = reinterpret_cast<const float4*>(Sv + (ra + i) * (long long)R);
    const float4* gr = reinterpret_cast<const float4*>(rng + (ra + i) * (long long)R);
    for (int c = tid; c < R / 4; c += kStagedThreads) cp_async16(dst + swz(c), gs + c);
    for (int c = tid; c < R / kRC; c += kStagedThreads) cp_async16(dst + svc + swz(c), gr + c);
  };
  // prologue: kStages - 1 rows in flight; one commit group per row slot (empty past the end) keeps the group count uniform
  for (int i = 0; i < kStages - 1; ++i) {
    if (i < nloc) stage(i, i);
    cp_async_commit();
  }
  __syncthreads();
  const RT inv_w = (RT)((double)nR / (edges[nR] - edges[0]));
  // this thread's chunk of every row: samples [n_begin, n_end), a multiple of 4 long
  const int per = (((R + kStagedThreads - 1) / kStagedThreads) + 3) & ~3;
  const int n_begin = tid * per < R ? tid * per : R, n_end = n_begin + per < R ? n_begin + per : R;
  long long cur_cell = -1;
  auto flush_cell = [&]() {  // all threads; shared bins -> float64 grid, then zero them
    double* cell = acc + cur_cell * (long long)nR * 4;
    for (int k = tid; k < nR; k += kStagedThreads) {
      const int g = s_good[k], b = s_bad[k];
      if (g) {
        atomicAdd(cell + 4 * (long long)k + 0, (double)s_sum[k]);
        atomicAdd(cell + 4 * (long long)k + 1, (double)g);
      }
      if (b) atomicAdd(cell + 4 * (long long)k + 2, (double)b);
      if (kHeight && s_h[k] != 0.f) atomicAdd(cell + 4 * (long long)k + 3, (double)s_h[k]);
      s_sum[k] = 0.f, s_good[k] = 0, s_bad[k] = 0;
      if (kHeight) s_h[k] = 0.f;
    }
  };
  int rows_in_cell = 0;
  // Register accumulators of TWO ADJACENT bins (kbase, kbase + 1), kept across the rows of a cell: a thread's chunk of
  // R / 256 samples starts in the same bin row after row and reaches at most the next one unless the bins are narrower
  // than the chunk or the range variable is not monotone.  The hot loop is branch-free: every sample is tested against
  // the three thresholds (a0, a1, a2) of those two bins and against the grid limits; a sample that belongs to a third
  // bin raises `other` and the chunk is redone by the per-sample path (bisection + shared atomics), which is exact for
  // any input.  A warp therefore never diverges on the bin boundaries that some lane meets in almost every row
  // (branching on them measured 9000 warp instructions per row, this form 2000).
  int kbase = -1;
  float sA = 0.f, sB = 0.f, hA = 0.f, hB = 0.f;
  int gA = 0, gB = 0, nA = 0, nB = 0;  // non-NaN members / members
  RT a0 = (RT)0, a1 = (RT)0, a2 = (RT)0;
  auto spill = [&](int k, float sm_, int g, int b_, float h) {  // one accumulator -> shared bins
    if (k >= 0 && k < nR) {
      if (g) atomicAdd(&s_sum[k], sm_), atomicAdd(&s_good[k], g);
      if (b_) atomicAdd(&s_bad[k], b_);
      if (kHeight && h != 0.f) atomicAdd(&s_h[k], h);
    }
  };
  auto spill_both = [&]() {
    if (kbase >= 0) {
      spill(kbase, sA, gA, nA - gA, hA);
      spill(kbase + 1, sB, gB, nB - gB, hB);
    }
    sA = sB = hA = hB = 0.f, gA = gB = nA = nB = 0;
  };
  const RT glo = s_thr[0], ghi = s_thr[nR];
  long long c = ra / P;
  int p = (int)(ra - c * P), slot = 0;
  for (int i = 0; i < nloc; ++i) {
    {  // refill the slot that row i - 1 used (free since the barrier that ended the previous iteration)
      int nslot = slot + kStages - 1;
      nslot = nslot >= kStages ? nslot - kStages : nslot;
      if (i + kStages - 1 < nloc) stage(i + kStages - 1, nslot);
      cp_async_commit();
    }
    cp_async_wait<kStages - 1>();
    __syncthreads();  // row i is in buffer `slot`
    const int xb = __ldg(xbin + p);
    const long long cell = (xb >= 0 && xb < nX) ? c * nX + xb : -1;
    // float32 bins: flush before a bin could collect more than ~2^22 samples (exact integer counts, bounded sums)
    if (cell != cur_cell || rows_in_cell >= 1024) {
      if (cur_cell >= 0) {
        spill_both();
        __syncthreads();
        flush_cell();
        __syncthreads();
      }
      cur_cell = cell, rows_in_cell = 0;
    }
    if (cell >= 0 && n_begin < n_end) {
      ++rows_in_cell;
      const float4* src = smem4 + slot * buf_chunks;
      // bin of the chunk's first member sample; rebase the accumulators when it moved (rare: law change, first row)
      {
        const RT x0 = reinterpret_cast<const RT*>(src + svc + swz(n_begin / kRC))[n_begin % kRC];
        const bool member = closed_right ? (x0 > glo && x0 <= ghi) : (x0 >= glo && x0 < ghi);
        const bool inA0 = kbase >= 0 && (closed_right ? (x0 > a0 && x0 <= a1) : (x0 >= a0 && x0 < a1));
        if (member && !inA0) {
          const int k0 = locate_bin<RT>(x0, s_thr, nR, inv_w, closed_right, kbase);
          spill_both();
          kbase = k0;
          a0 = s_thr[k0], a1 = s_thr[k0 + 1];
          a2 = (k0 + 2 <= nR) ? s_thr[k0 + 2] : a1;  // no bin B past the grid: [a1, a1) is empty
        }
      }
      bool other = false;
      if (kbase >= 0) {
        float rsA = 0.f, rsB = 0.f, rhA = 0.f, rhB = 0.f;
        int rgA = 0, rgB = 0, rnA = 0, rnB = 0;
        for (int n0 = n_begin; n0 < n_end; n0 += 4) {
          const float4 w = src[swz(n0 >> 2)];
          const float sv4[4] = {w.x, w.y, w.z, w.w};
          RT xv[5];
          if (sizeof(RT) == 4) {
            const float4 q = src[svc + swz(n0 >> 2)];
            xv[0] = (RT)q.x, xv[1] = (RT)q.y, xv[2] = (RT)q.z, xv[3] = (RT)q.w;
          } else {
            const double2 q0 = *reinterpret_cast<const double2*>(src + svc + swz(n0 >> 1));
            const double2 q1 = *reinterpret_cast<const double2*>(src + svc + swz((n0 >> 1) + 1));
            xv[0] = (RT)q0.x, xv[1] = (RT)q0.y, xv[2] = (RT)q1.x, xv[3] = (RT)q1.y;
          }
          if (kHeight) {  // the sample after this group (next chunk, possibly another thread's): diff(label="lower")
            const int nn = n0 + 4;
            xv[4] = (RT)0;
            if (nn < R) xv[4] = reinterpret_cast<const RT*>(src + svc + swz(nn / kRC))[nn % kRC];
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const RT x = xv[k];
            const bool inA = closed_right ? (x > a0 && x <= a1) : (x >= a0 && x < a1);
            const bool inB = closed_right ? (x > a1 && x <= a2) : (x >= a1 && x < a2);
            const bool member = closed_right ? (x > glo && x <= ghi) : (x >= glo && x < ghi);
            other = other || (member && !inA && !inB);
            const float v = fast_exp2(sv4[k] * kDb2Log2);  // commongrid/utils.py:592  10^(Sv/10)
            const bool ok = (v == v);
            rsA += (inA && ok) ? v : 0.f;
            rsB += (inB && ok) ? v : 0.f;
            rgA += (inA && ok) ? 1 : 0;
            rgB += (inB && ok) ? 1 : 0;
            rnA += inA ? 1 : 0;
            rnB += inB ? 1 : 0;
            if (kHeight && n0 + k + 1 < R) {
              const double d = (double)xv[k + 1] - (double)x;  // commongrid/utils.py:170-172
              if (d == d) {
                rhA += inA ? (float)d : 0.f;
                rhB += inB ? (float)d : 0.f;
              }
            }
          }
        }
        if (!other) sA += rsA, sB += rsB, hA += rhA, hB += rhB, gA += rgA, gB += rgB, nA += rnA, nB += rnB;
      } else {
        other = true;  // no member seen yet at the chunk start: let the exact path look at every sample
      }
      if (other) {  // exact per-sample path: any bin order, any bin width
        int guess = kbase;
        for (int n = n_begin; n < n_end; ++n) {
          const RT x = reinterpret_cast<const RT*>(src + svc + swz(n / kRC))[n % kRC];
          const int key = locate_bin<RT>(x, s_thr, nR, inv_w, closed_right, guess);
          if (key >= 0) {
            guess = key;
            const float sv = reinterpret_cast<const float*>(src + swz(n >> 2))[n & 3];
            const float v = fast_exp2(sv * kDb2Log2);
            const bool ok = (v == v);
            float h = 0.f;
            if (kHeight && n + 1 < R) {
              const RT xn = reinterpret_cast<const RT*>(src + svc + swz((n + 1) / kRC))[(n + 1) % kRC];
              const double d = (double)xn - (double)x;
              if (d == d) h = (float)d;
            }
            spill(key, ok ? v : 0.f, ok ? 1 : 0, ok ? 0 : 1, h);
          }
        }
      }
    }
    __syncthreads();  // buffer `slot` is free for the stage issued by the next iteration
    if (++slot == kStages) slot = 0;
    if (++p == (int)P) p = 0, ++c;
  }
  if (cur_cell >= 0) {
    spill_both();
    __syncthreads();
    flush_cell();
  }
}

// ---- law: bin boundaries in sample-index space from the exact float64 range law ---------------------------
// value(n) = depth_off[p] + law_range(row, n) * depth_scale[p]   (echo_range: off = 0, scale = 1)
__device__ __forceinline__ double law_value(const epb_row& r, int n, double off, double scale, bool is_depth) {
  const double R = law_range(r, n);
  return is_depth ? __dadd_rn(off, __dmul_rn(R, scale)) : R;
}

// smallest n in [0, R] with value(n) >= edge (closed left) / value(n) > edge (closed right); value monotone in n
__device__ int first_at_or_above(const epb_row& r, int R, double edge, int closed_right, double off, double scale,
                                 bool is_depth) {
  int lo = 0, hi = R;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const double v = law_value(r, mid, off, scale, is_depth);
    const bool pass = closed_right ? (v > edge) : (v >= edge);
    if (pass)
      hi = mid;
    else
      lo = mid + 1;  // also taken for NaN: a NaN law puts every sample outside
  }
  return lo;
}

__global__ void __launch_bounds__(32 * kWarpsPerCta) bin_reduce_law_kernel(
    const float* __restrict__ Sv, const epb_row* __restrict__ rows, const double* __restrict__ depth_off,
    const double* __restrict__ depth_scale, const int* __restrict__ xbin, const double* __restrict__ edges, int nR,
    int closed_right, double* __restrict__ acc, long long C, long long P, int R, long long nX, const int* __restrict__ gate) {
  if (gate && *gate == 0) return;  // the persistent fast kernel did the reduction
  extern __shared__ double s_edges[];                                        // [nR+1]
  int* s_bounds = reinterpret_cast<int*>(s_edges + (nR + 1));                // [kWarpsPerCta][nR+1]
  for (int k = threadIdx.x; k <= nR; k += blockDim.x) s_edges[k] = edges[k];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int* bnd = s_bounds + w * (nR + 1);
  const bool is_depth = depth_off != nullptr;
  const long long nrows = C * P;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerCta + w;
  const bool vec = (R & 3) == 0 && ((uintptr_t)Sv & 15) == 0;
  long long prev_row = -1;  // the row whose boundaries are in bnd
  double prev_off = 0.0, prev_scale = 0.0;
  for (long long row = warp0; row < nrows; row += (long long)gridDim.x * kWarpsPerCta) {
    const long long c = row / P, p = row % P;
    const int xb = __ldg(xbin + p);
    if (xb < 0 || xb >= nX) continue;
    const double off = is_depth ? depth_off[p] : 0.0, scale = is_depth ? depth_scale[p] : 1.0;
    // The boundaries depend on the row's range law and (depth) offset / scale only: a row whose first 128 bytes (the
    // float64 law block of epb_row) and offset / scale equal those of the row the warp did last keeps them - the usual
    // case (one law per channel, a constant transducer depth): no 12-step float64 bisection per edge and row.
    bool same = false;
    if (prev_row >= 0) {
      const unsigned long long* ra = reinterpret_cast<const unsigned long long*>(rows + row);
      const unsigned long long* rb = reinterpret_cast<const unsigned long long*>(rows + prev_row);
      const bool eq = lane < 16 ? (__ldg(ra + lane) == __ldg(rb + lane)) : true;
      same = __all_sync(0xffffffffu, eq) && off == prev_off && scale == prev_scale && row / P == prev_row / P;
    }
    if (!same) {
      const epb_row r = rows[row];
      __syncwarp();
      for (int k = lane; k <= nR; k += 32) bnd[k] = first_at_or_above(r, R, s_edges[k], closed_right, off, scale, is_depth);
      __syncwarp();
      prev_row = row, prev_off = off, prev_scale = scale;
    }
    double* acc_row = acc + ((c * nX + xb) * (long long)nR) * 4;
    const float* sv = Sv + row * (long long)R;
    const int n_lo = bnd[0], n_hi = bnd[nR];  // members are n_lo <= n < n_hi
    int kcur = 0;
    for (int n0 = (n_lo / 128) * 128; n0 < n_hi; n0 += 128) {
      const int nb = n0 + 4 * lane;
      int keys[4];
      Acc3 a[4];
      float sv4[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec && nb < R) {  // one 16-byte load (R % 4 == 0: the four samples are inside the row)
        const float4 q = ld_stream4(reinterpret_cast<const float4*>(sv + nb));
        sv4[0] = q.x, sv4[1] = q.y, sv4[2] = q.z, sv4[3] = q.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int n = nb + k;
        keys[k] = -1;
        a[k].sum = 0.f, a[k].cnt = 0, a[k].h = 0.f;
        if (n >= n_lo && n < n_hi) {
          while (bnd[kcur + 1] <= n) ++kcur;  // bnd is non-decreasing; kcur only moves forward
          keys[k] = kcur;
          add_sample(a[k], vec ? sv4[k] : ld_stream(sv + n));
        }
      }
      const bool same = (keys[0] == keys[1]) && (keys[1] == keys[2]) && (keys[2] == keys[3]);
      int key = -1;
      Acc3 m;
      m.sum = 0.f, m.cnt = 0, m.h = 0.f;
      if (same) {
        key = keys[0];
        m.sum = (a[0].sum + a[1].sum) + (a[2].sum + a[3].sum);
        m.cnt = a[0].cnt + a[1].cnt + a[2].cnt + a[3].cnt;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (keys[k] >= 0) flush_run(acc_row + 4 * (long long)keys[k], a[k], false);
      }
      warp_runs_to_acc<false>(key, m, acc_row);
    }
  }
}

// ---- finalise: (sum, n_good, n_nan, h) -> mean [-> dB] -------------------------------------------------------
__global__ void bin_finalize_kernel(const double* __restrict__ acc, float* __restrict__ out, double* __restrict__ h_out,
                                    long long ncell, int skipna, float fill_value, int to_db) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  const double s = acc[4 * i], good = acc[4 * i + 1], bad = acc[4 * i + 2];
  double m;
  if (good + bad == 0.0)
    m = (double)fill_value;  // no member: flox fill_value
  else if (!skipna && bad > 0.0)
    m = CUDART_NAN;          // func="mean": a NaN member poisons the bin
  else
    m = s / good;            // all-NaN members: 0/0 = NaN, as nanmean
  out[i] = (float)(to_db ? 10.0 * log10(m) : m);  // _lin2log, commongrid/utils.py:92
  if (h_out) h_out[i] = acc[4 * i + 3];
}

}  // namespace

extern "C" int epb_bin_reduce(const float* Sv, const void* range_var, int range_is_f64, const int* xbin,
                              const double* r_edges, int nR, int closed_right, int with_height, double* acc, epb_i64 C,
                              epb_i64 P, epb_i64 R, epb_i64 nX, void* stream) {
  EPB_REQUIRE(Sv && range_var && xbin && r_edges && acc, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30) && nX > 0, "bad shape");
  EPB_REQUIRE(nR > 0 && nR <= 20000, "number of range bins must be in 1..20000");
  const long long nrows = C * P;
  cudaStream_t s = (cudaStream_t)stream;
  // staged kernel: 16-byte aligned rows of R % 4 == 0 samples, bins that fit shared memory
  if (R % 4 == 0 && R >= 64 && nR <= kStagedMaxBins && P < (1LL << 31) && ((uintptr_t)Sv % 16) == 0 && ((uintptr_t)range_var % 16) == 0) {
    const size_t esz = range_is_f64 ? 8 : 4;
    const size_t svc = ((size_t)(R / 4) + 7) & ~(size_t)7, rgc = ((size_t)(R / (16 / esz)) + 7) & ~(size_t)7;
    const size_t sm = (((size_t)(nR + 1) * esz + 15) & ~(size_t)15) + (size_t)nR * (with_height ? 16 : 12) + 16 + kStages * (svc + rgc) * 16;
    if (sm <= 200 * 1024) {
      int per_sm = (int)((227 * 1024) / (sm + 1024));
      per_sm = per_sm > 4 ? 4 : (per_sm < 1 ? 1 : per_sm);
      long long g2 = (long long)epb_num_sms() * per_sm;
      if (g2 > nrows) g2 = nrows;
#define EPB_BRS(T, H)                                                                                                  \
  do {                                                                                                                 \
    cudaFuncSetAttribute(bin_reduce_staged_kernel<T, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);        \
    bin_reduce_staged_kernel<T, H><<<(unsigned)g2, kStagedThreads, sm, s>>>(Sv, (const T*)range_var, xbin, r_edges, nR, \
                                                                            closed_right, acc, C, P, (int)R, nX);      \
  } while (0)
      if (range_is_f64) {
        if (with_height)
          EPB_BRS(double, true);
        else
          EPB_BRS(double, false);
      } else {
        if (with_height)
          EPB_BRS(float, true);
        else
          EPB_BRS(float, false);
      }
#undef EPB_BRS
      return epb_check_launch("epb_bin_reduce(staged)");
    }
  }
  long long grid = (nrows + kWarpsPerCta - 1) / kWarpsPerCta;
  const long long cap = (long long)epb_num_sms() * 8;
  if (grid > cap) grid = cap;
  const size_t smem = (size_t)(nR + 1) * (sizeof(double) + sizeof(float));  // edges + float32 thresholds
#define EPB_BR(T, H)                                                                                        \
  do {                                                                                                      \
    if (smem > 48 * 1024) cudaFuncSetAttribute(bin_reduce_kernel<T, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    bin_reduce_kernel<T, H><<<(unsigned)grid, 32 * kWarpsPerCta, smem, s>>>(Sv, (const T*)range_var, xbin, r_edges, nR, \
                                                                            closed_right, acc, C, P, (int)R, nX); \
  } while (0)
  if (range_is_f64) {
    if (with_height)
      EPB_BR(double, true);
    else
      EPB_BR(double, false);
  } else {
    if (with_height)
      EPB_BR(float, true);
    else
      EPB_BR(float, false);
  }
#undef EPB_BR
  return epb_check_launch("epb_bin_reduce");
}

int epb_pipeline_fast_try(const void* x, int x_i16, const epb_row* rows, const int* xbin, const double* r_edges, int nR,
                          int closed_right, double* acc, float* noise_out, long long C, long long P, long long R,
                          long long nX, int ping_num, int range_sample_num, float noise_max_lin, float snr_lin,
                          double* range_max_out, int sv_input, void* workspace, long long workspace_bytes, cudaStream_t s,
                          float* o_sv = nullptr, float* o_rng = nullptr, float* o_svn = nullptr, float* o_svc = nullptr);

extern "C" int epb_bin_reduce_law(const float* Sv, const epb_row* rows, const double* depth_off,
                                  const double* depth_scale, const int* xbin, const double* r_edges, int nR,
                                  int closed_right, double* acc, epb_i64 C, epb_i64 P, epb_i64 R, epb_i64 nX,
                                  void* workspace, epb_i64 workspace_bytes, void* stream) {
  EPB_REQUIRE(Sv && rows && xbin && r_edges && acc, "NULL pointer");
  EPB_REQUIRE((depth_off == nullptr) == (depth_scale == nullptr), "depth_off and depth_scale go together");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1LL << 30) && nX > 0, "bad shape");
  EPB_REQUIRE(nR > 0 && nR <= kMaxLawBins, "law path supports 1..511 range bins (use epb_bin_reduce)");
  // echo_range binning of a regular volume: the persistent register-accumulating kernel of the fused pipeline, fed
  // with Sv instead of power (pipeline_fast.cu, sv_input); decided on the device like in epb_pipeline_power_mvbs
  const int* gate = nullptr;
  if (workspace && workspace_bytes >= 256 && !depth_off && ((uintptr_t)workspace % 16) == 0 && ((uintptr_t)Sv % 16) == 0 &&
      epb_pipeline_fast_try(Sv, 0, rows, xbin, r_edges, nR, closed_right, acc, nullptr, C, P, R, nX, 0, 0, nanf(""), 0.f, nullptr, 1,
                            workspace, workspace_bytes, (cudaStream_t)stream))
    gate = (const int*)workspace;
  const long long nrows = C * P;
  long long grid = (nrows + kWarpsPerCta - 1) / kWarpsPerCta;
  const long long cap = (long long)epb_num_sms() * 8;
  if (grid > cap) grid = cap;
  const size_t smem = (size_t)(nR + 1) * sizeof(double) + (size_t)kWarpsPerCta * (nR + 1) * sizeof(int);
  bin_reduce_law_kernel<<<(unsigned)grid, 32 * kWarpsPerCta, smem, (cudaStream_t)stream>>>(
      Sv, rows, depth_off, depth_scale, xbin, r_edges, nR, closed_right, acc, C, P, (int)R, nX, gate);
  return epb_check_launch("epb_bin_reduce_law");
}

extern "C" int epb_bin_finalize(const double* acc, float* out, double* h_out, epb_i64 ncell, int skipna,
                                float fill_value, int to_db, void* stream) {
  EPB_REQUIRE(acc && out && ncell > 0, "bad pointer/size");
  bin_finalize_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, (cudaStream_t)stream>>>(acc, out, h_out, ncell, skipna,
                                                                                         fill_value, to_db);
  return epb_check_launch("epb_bin_finalize");
}

// ---- ping-sharded execution: the straddling-bin exchange (pipeline.straddle_reduce) in two launches around ONE
//      all-reduce(sum).  buf: [world][2 S + 1] float64 with S = C * nR * 4: rank r fills its own slot with the first and
//      the last local ping bin of its accumulators (the only bins another rank can also hold) and its exact range
//      maximum, zeros everywhere else; after the sum every rank sees every slot.  unpack adds, for its own first / last
//      bin, the slots of the ranks that share that global bin (src lists from the host plan) and takes the maximum of
//      the range maxima. -------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) straddle_pack_kernel(const double* __restrict__ acc, long long C, long long nXl, long long SC,
                                                            const double* __restrict__ rmax, int nrmax, double* __restrict__ buf,
                                                            int rank, int world, int has_last) {
  const long long S = C * SC, W = 2 * S + 1;
  const long long n = W * world;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long w = i / W, j = i - w * W;
    double v = 0.0;
    if (w == rank) {
      if (j < 2 * S) {
        const long long part = j / S, k = j - part * S, c = k / SC, q = k - c * SC;
        if (part == 0)
          v = acc[(c * nXl) * SC + q];
        else if (has_last)
          v = acc[(c * nXl + (nXl - 1)) * SC + q];
      } else {
        v = -CUDART_INF;
        for (int t = 0; t < nrmax; ++t) v = fmax(v, rmax[t]);
      }
    }
    buf[i] = v;
  }
}

__global__ void __launch_bounds__(256) straddle_unpack_kernel(const double* __restrict__ buf, const int* __restrict__ src,
                                                              int nsrc0, int nsrc1, long long C, long long nXl, long long SC,
                                                              int world, double* __restrict__ acc, double* __restrict__ rmax_out) {
  const long long S = C * SC, W = 2 * S + 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < 2 * S; i += (long long)gridDim.x * blockDim.x) {
    const long long slot = i / S, k = i - slot * S, c = k / SC, q = k - c * SC;
    const int ns = slot == 0 ? nsrc0 : nsrc1;
    if (ns == 0) continue;  // this slot is not shared (or hi == lo and the slot is the same bin as slot 0)
    double v = 0.0;
    for (int t = 0; t < ns; ++t) {
      const int e = src[slot * world + t];  // 2 r: first bin of rank r, 2 r + 1: its last bin
      v += buf[(long long)(e >> 1) * W + (long long)(e & 1) * S + k];
    }
    acc[(c * nXl + (slot == 0 ? 0 : nXl - 1)) * SC + q] = v;
  }
  if (rmax_out && blockIdx.x == 0 && threadIdx.x == 0) {
    double m = -CUDART_INF;
    for (int w = 0; w < world; ++w) m = fmax(m, buf[(long long)w * W + 2 * S]);
    *rmax_out = m;
  }
}
}  // namespace

extern "C" int epb_straddle_pack(const double* acc, epb_i64 C, epb_i64 nXl, epb_i64 nR, const double* rmax, int nrmax,
                                 double* buf, int rank, int world, int has_last, void* stream) {
  EPB_REQUIRE(acc && buf && C > 0 && nXl > 0 && nR > 0 && world > 0 && rank >= 0 && rank < world, "bad pointer/shape");
  EPB_REQUIRE(rmax || nrmax == 0, "rmax pointer missing");
  const long long n = (2 * C * nR * 4 + 1) * (long long)world;
  straddle_pack_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, (cudaStream_t)stream>>>(
      acc, C, nXl, nR * 4, rmax, nrmax, buf, rank, world, has_last);
  return epb_check_launch("epb_straddle_pack");
}

extern "C" int epb_straddle_unpack(const double* buf, const int* src, int nsrc0, int nsrc1, epb_i64 C, epb_i64 nXl, epb_i64 nR,
                                   int world, double* acc, double* rmax_out, void* stream) {
  EPB_REQUIRE(buf && src && acc && C > 0 && nXl > 0 && nR > 0 && world > 0, "bad pointer/shape");
  EPB_REQUIRE(nsrc0 >= 0 && nsrc0 <= world && nsrc1 >= 0 && nsrc1 <= world, "bad source counts");
  const long long n = 2 * C * nR * 4;
  straddle_unpack_kernel<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, (cudaStream_t)stream>>>(
      buf, src, nsrc0, nsrc1, C, nXl, nR * 4, world, acc, rmax_out);
  return epb_check_launch("epb_straddle_unpack");
}
