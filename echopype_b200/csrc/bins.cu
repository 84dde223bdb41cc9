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
  for (long long row = warp0; row < nrows; row += (long long)gridDim.x * kWarpsPerCta) {
    const long long c = row / P, p = row % P;
    const int xb = __ldg(xbin + p);
    if (xb < 0 || xb >= nX) continue;
    const epb_row r = rows[row];
    const double off = is_depth ? depth_off[p] : 0.0, scale = is_depth ? depth_scale[p] : 1.0;
    __syncwarp();
    for (int k = lane; k <= nR; k += 32) bnd[k] = first_at_or_above(r, R, s_edges[k], closed_right, off, scale, is_depth);
    __syncwarp();
    double* acc_row = acc + ((c * nX + xb) * (long long)nR) * 4;
    const float* sv = Sv + row * (long long)R;
    const int n_lo = bnd[0], n_hi = bnd[nR];  // members are n_lo <= n < n_hi
    int kcur = 0;
    for (int n0 = (n_lo / 128) * 128; n0 < n_hi; n0 += 128) {
      const int nb = n0 + 4 * lane;
      int keys[4];
      Acc3 a[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int n = nb + k;
        keys[k] = -1;
        a[k].sum = 0.f, a[k].cnt = 0, a[k].h = 0.f;
        if (n >= n_lo && n < n_hi) {
          while (bnd[kcur + 1] <= n) ++kcur;  // bnd is non-decreasing; kcur only moves forward
          keys[k] = kcur;
          add_sample(a[k], ld_stream(sv + n));
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
  long long grid = (nrows + kWarpsPerCta - 1) / kWarpsPerCta;
  const long long cap = (long long)epb_num_sms() * 8;
  if (grid > cap) grid = cap;
  const size_t smem = (size_t)(nR + 1) * (sizeof(double) + sizeof(float));  // edges + float32 thresholds
  cudaStream_t s = (cudaStream_t)stream;
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
                          double* range_max_out, int sv_input, void* workspace, long long workspace_bytes, cudaStream_t s);

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
