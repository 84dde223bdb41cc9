// Fused pipeline: raw power -> Sv -> background-noise removal -> MVBS accumulators, ONE pass over HBM
// (compute_Sv calibrate/api.py:249 -> remove_background_noise clean/api.py:436 -> compute_MVBS
// commongrid/api.py:31).  4 algorithmic bytes per sample: the power sample is read
// once; Sv, echo_range, Sv_noise, Sv_corrected exist only in registers unless the caller asks for them.
//
// One CTA per (channel, ping tile of `tile` = ping_num rows).  The rows of a tile are contiguous in the
// (channel, ping_time, range_sample) layout, so the whole tile is staged in shared memory with TMA bulk
// copies (cp.async.bulk, one per row, completion on an mbarrier; SASS: UBLKCP) while the CTA prepares the
// per-row constants and the exact bin boundaries.  Then
//   phase 1  e = 10^((front(x)-K)/10) replaces x in shared memory; per-column sums over the tile rows ->
//            (ping_num x range_sample_num) tile means of 10^((Sv-TL)/10) -> min over range tiles = noise
//            (clean/api.py:397-411, entirely in the linear domain: no log per sample);
//   phase 2  sv_lin = e * R'^2 * 10^(2aR'/10);  noise_lin(n) = noise * TL_lin(n);  the sample survives iff
//            sv_lin > noise_lin * (1 + 10^(SNR/10))  (<=> Sv_corrected - Sv_noise > SNR, :485-487) and then
//            contributes sv_lin - noise_lin to its (ping bin, range bin) cell (commongrid/utils.py:592-627).
// Range bins are located in sample-index space by bisection on the exact float64 range law (bit-identical
// membership with the float64 reference); per-thread column sums are merged with a segmented warp-shuffle
// reduction over runs of equal bin index, accumulated per CTA in shared memory and flushed once per touched
// cell with float64 atomics.  When all rows of a tile share one range law (the normal case: constant
// sample interval / sound speed / absorption) all range-only terms are computed once per column.
#include "pipeline_common.cuh"

namespace {
using namespace epb;

constexpr int kThreads = 256;

struct RowS {  // per-row constants in shared memory
  RowF f;
  int slot;    // index of the first tile row with the same ping bin, -1: ping outside every bin
  int xb;
  int bad;     // NaN calibration constants: every output of the row is NaN, but echo_range / bin membership still
               // follow the raw sample (the staged tile keeps x instead of e for such rows)
};

struct Cell {
  float sum;
  int cnt;  // low 16: surviving (non-NaN) members, high 16: NaN members
};

__device__ __forceinline__ void cell_add(Cell* cell, float s, int cnt) {
  if (cnt & 0xffff) atomicAdd(&cell->sum, s);
  if (cnt) atomicAdd(&cell->cnt, cnt);
}

// segmented warp reduction over runs of equal keys (see bins.cu); tail lanes add to the CTA accumulators
__device__ __forceinline__ void warp_runs_to_smem(int key, float s, int cnt, Cell* __restrict__ cells) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync(full, key, 1);
  const bool head = (lane == 0) || (prev != key);
  const unsigned heads = __ballot_sync(full, head);
  if (heads == 1u && key < 0) return;
  const int run = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float s2 = __shfl_up_sync(full, s, d);
    const int c2 = __shfl_up_sync(full, cnt, d);
    const int r2 = __shfl_up_sync(full, run, d);
    if (lane >= d && r2 == run) {
      s += s2;
      cnt += c2;
    }
  }
  const int next = __shfl_down_sync(full, run, 1);
  if (((lane == 31) || (next != run)) && key >= 0) cell_add(cells + key, s, cnt);
}

struct Params {
  const float* x;
  const epb_row* rows;
  const int* xbin;
  const double* edges;
  double* acc;
  float* noise_out;
  float* o_sv;
  float* o_rng;
  float* o_svn;
  float* o_svc;
  long long C, P, nX;
  int R, nR, tile, rs_num, closed_right, do_noise, staged;
  float noise_max_lin;  // NaN: no cap
  float snr_fac;        // 1 + 10^(SNR/10)
  const int* gate;      // NULL, or device flag: run only when *gate != 0 (the fast path declined the launch)
};

__global__ void __launch_bounds__(kThreads, 2) pipeline_kernel(const Params pr) {
  if (pr.gate && *pr.gate == 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = pr.R, nR = pr.nR, T = pr.tile;
  // ---- shared-memory carve-up ------------------------------------------------------------------------------
  size_t off = 0;
  float* s_tile = reinterpret_cast<float*>(smem_raw);
  off += pr.staged ? (size_t)T * R * 4 : 0;
  float* s_colsum = reinterpret_cast<float*>(smem_raw + off);
  off += pr.do_noise ? (size_t)R * 4 : 0;
  unsigned short* s_colcnt = reinterpret_cast<unsigned short*>(smem_raw + off);
  off += pr.do_noise ? (((size_t)R * 2 + 15) & ~(size_t)15) : 0;
  double* s_edges = reinterpret_cast<double*>(smem_raw + off);
  off += (size_t)(nR + 1) * 8;
  RowS* s_rows = reinterpret_cast<RowS*>(smem_raw + off);
  off += (size_t)T * sizeof(RowS);
  off = (off + 15) & ~(size_t)15;
  Cell* s_cells = reinterpret_cast<Cell*>(smem_raw + off);  // [T slots][nR]
  off += (size_t)T * nR * sizeof(Cell);
  int* s_bounds = reinterpret_cast<int*>(smem_raw + off);   // [T][nR+1]
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ float s_red[32];
  __shared__ int s_shared_law;
  __shared__ float s_noise_lin;

  const int nPt = (int)((pr.P + T - 1) / T);
  // grid-stride over the (channel, ping tile) tiles: a launch that loses the device-side dispatch against the fast
  // kernel exits after a few hundred CTAs instead of one per tile
  for (long long tile = blockIdx.x; tile < pr.C * nPt; tile += gridDim.x) {
  const long long c = tile / nPt;
  const int it = (int)(tile % nPt);
  const long long p0 = (long long)it * T;
  const int Ta = (int)((p0 + T <= pr.P) ? T : (pr.P - p0));  // rows actually present in this tile
  const long long row0 = c * pr.P + p0;
  const float* xg = pr.x + row0 * (long long)R;
  const int tid = threadIdx.x;

  // ---- phase 0: start the TMA bulk copies, meanwhile build row constants / bin boundaries -------------------
  if (pr.staged && tid == 0) {
    if (tile != blockIdx.x) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the tile vs the new bulk copies
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t row_bytes = (uint32_t)R * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(row_bytes * (uint32_t)Ta)
                 : "memory");
    for (int t = 0; t < Ta; ++t)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(s_tile + (size_t)t * R)),
                   "l"(xg + (size_t)t * R), "r"(row_bytes), "r"(smem_u32(&s_bar))
                   : "memory");
  }
  for (int k = tid; k <= nR; k += kThreads) s_edges[k] = pr.edges[k];
  for (int i = tid; i < T * nR; i += kThreads) {
    s_cells[i].sum = 0.f;
    s_cells[i].cnt = 0;
  }
  if (tid < Ta) {
    RowS rs;
    rs.f = load_rowf(pr.rows + row0 + tid);
    rs.xb = __ldg(pr.xbin + p0 + tid);
    if (rs.xb < 0 || rs.xb >= pr.nX) rs.xb = -1;
    rs.slot = -1;
    rs.bad = !(rs.f.c0 == rs.f.c0 && rs.f.c1 == rs.f.c1);
    s_rows[tid] = rs;
  }
  __syncthreads();
  if (tid == 0) {
    int shared_law = 1;
    const RowF& a = s_rows[0].f;
    for (int t = 0; t < Ta; ++t) {
      const RowF& b = s_rows[t].f;
      // bitwise-equal range laws (NaN laws never compare equal -> generic path)
      if (!(a.a_h == b.a_h && a.a_l == b.a_l && a.r0_h == b.r0_h && a.r0_l == b.r0_l && a.bp_h == b.bp_h &&
            a.bp_l == b.bp_l && a.c2 == b.c2 && a.n_start == b.n_start && a.nanrange == b.nanrange))
        shared_law = 0;
      int slot = -1;
      if (s_rows[t].xb >= 0) {
        slot = t;
        for (int u = 0; u < t; ++u)
          if (s_rows[u].xb == s_rows[t].xb) {
            slot = u;
            break;
          }
      }
      s_rows[t].slot = slot;
    }
    s_shared_law = shared_law;
  }
  __syncthreads();
  const bool shared_law = s_shared_law != 0;
  {  // exact bin boundaries: one set when the law is shared, else one per row
    const int nsets = shared_law ? 1 : Ta;
    for (int i = tid; i < nsets * (nR + 1); i += kThreads) {
      const int t = i / (nR + 1), k = i % (nR + 1);
      s_bounds[i] = first_at_or_above(pr.rows[row0 + t], R, s_edges[k], pr.closed_right);
    }
  }
  __syncthreads();
  if (pr.staged) {  // wait for the tile
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(smem_u32(&s_bar))
          : "memory");
    }
  }

  const int R4 = R >> 2;  // staged / vector path requires R % 4 == 0 (checked on the host)
  auto load4 = [&](int t, int j) -> float4 {
    if (pr.staged) return *reinterpret_cast<const float4*>(s_tile + (size_t)t * R + 4 * j);
    return ld_stream4(reinterpret_cast<const float4*>(xg + (size_t)t * R) + j);
  };

  // ---- phase 1: noise estimate ----------------------------------------------------------------------------------
  float noise_lin = 0.f;
  if (pr.do_noise) {
    for (int j = tid; j < R4; j += kThreads) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      int cn[4] = {0, 0, 0, 0};
      ColC cc[4];
      if (shared_law) {
#pragma unroll
        for (int k = 0; k < 4; ++k) cc[k] = col_consts(s_rows[0].f, 4 * j + k);
      }
      for (int t = 0; t < Ta; ++t) {
        const RowF& rf = s_rows[t].f;
        const float4 v4 = load4(t, j);
        const float xin[4] = {v4.x, v4.y, v4.z, v4.w};
        float e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          e[k] = fast_exp2(fmaf(xin[k], rf.c1, rf.c0));  // 10^((front(x) - K)/10)
          float q = e[k];
          if (!shared_law) q *= col_consts(rf, 4 * j + k).g;
          const bool ok = (q == q);
          s[k] += ok ? q : 0.f;
          cn[k] += ok;
        }
        if (pr.staged && !s_rows[t].bad)
          *reinterpret_cast<float4*>(s_tile + (size_t)t * R + 4 * j) = make_float4(e[0], e[1], e[2], e[3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (shared_law) {
          s[k] *= cc[k].g;
          if (!(cc[k].g == cc[k].g)) s[k] = 0.f, cn[k] = 0;  // n < n_start: Sv is NaN, skipped by nanmean
        }
        s_colsum[4 * j + k] = s[k];
        s_colcnt[4 * j + k] = (unsigned short)cn[k];
      }
    }
    __syncthreads();
    const int nRt = (R + pr.rs_num - 1) / pr.rs_num;
    float best = CUDART_INF_F;
    for (int t = tid; t < nRt; t += kThreads) {
      const int j0 = t * pr.rs_num, j1 = (j0 + pr.rs_num < R) ? j0 + pr.rs_num : R;
      double s = 0.0;
      int n = 0;
      for (int j = j0; j < j1; ++j) {
        s += (double)s_colsum[j];
        n += s_colcnt[j];
      }
      if (n > 0) best = fminf(best, (float)(s / (double)n));  // min over range tiles of the tile mean (monotone in dB)
    }
    best = warp_min(best);
    if ((tid & 31) == 0) s_red[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
      float b = CUDART_INF_F;
      for (int w = 0; w < kThreads / 32; ++w) b = fminf(b, s_red[w]);
      float v = (b == CUDART_INF_F) ? CUDART_NAN_F : b;
      if (pr.noise_max_lin == pr.noise_max_lin) v = (v < pr.noise_max_lin) ? v : pr.noise_max_lin;  // NaN -> max
      s_noise_lin = v;
      if (pr.noise_out) pr.noise_out[c * nPt + it] = kLog2ToDb * fast_log2(v);
    }
    __syncthreads();
    noise_lin = s_noise_lin;
  }

  // ---- phase 2: noise removal + bin accumulation (+ optional full-size outputs) ----------------------------------
  const bool want_out = pr.o_sv || pr.o_rng || pr.o_svn || pr.o_svc;
  for (int j0 = 0; j0 < R4; j0 += kThreads) {  // every thread runs every iteration: the reduction below is warp-wide
    const int j = j0 + tid;
    const bool live = j < R4;
    ColC cc[4];
    int key[4] = {-1, -1, -1, -1};
    if (live && shared_law) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cc[k] = col_consts(s_rows[0].f, 4 * j + k);
        key[k] = key_of(s_bounds, nR, 4 * j + k);
      }
    }
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    int cn[4] = {0, 0, 0, 0};
    int cur_slot = -2;
    for (int t = 0; t <= Ta; ++t) {
      const int slot = (t < Ta) ? s_rows[t].slot : -3;
      if (t > 0 && (slot != cur_slot || !shared_law)) {  // end of a run of rows sharing (ping bin, bin boundaries): flush
        if (cur_slot >= 0) {
          Cell* cells = s_cells + (size_t)cur_slot * nR;
          const bool same = (key[0] == key[1]) && (key[1] == key[2]) && (key[2] == key[3]);
          int kk = -1;
          float ms = 0.f;
          int mc = 0;
          if (same) {
            kk = key[0];
            ms = (s[0] + s[1]) + (s[2] + s[3]);
            mc = cn[0] + cn[1] + cn[2] + cn[3];
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (key[k] >= 0) cell_add(cells + key[k], s[k], cn[k]);
          }
          warp_runs_to_smem(kk, ms, mc, cells);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = 0.f, cn[k] = 0;
      }
      cur_slot = slot;
      if (t == Ta) break;
      if (!live) continue;
      const RowF& rf = s_rows[t].f;
      const bool rbad = s_rows[t].bad != 0;
      if (!shared_law) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          cc[k] = col_consts(rf, 4 * j + k);
          key[k] = key_of(s_bounds + t * (nR + 1), nR, 4 * j + k);
        }
      }
      const float4 v4 = load4(t, j);
      const float ein[4] = {v4.x, v4.y, v4.z, v4.w};
      float osv[4], orr[4], osn[4], osc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool raw = !(pr.staged && pr.do_noise) || rbad;  // tile holds the raw sample (else e from phase 1)
        const float e = raw ? fast_exp2(fmaf(ein[k], rf.c1, rf.c0)) : ein[k];
        const float v = e * cc[k].h;  // 10^(Sv/10)
        const bool x_nan = rbad ? !(ein[k] == ein[k]) : !(e == e);
        const bool member = !(rf.nanrange && x_nan);  // echo_range is NaN where the sample is NaN (range.py:143-148)
        float contrib;
        bool good;
        if (pr.do_noise) {
          const float nl = noise_lin * cc[k].tl;  // 10^(Sv_noise/10)
          good = v > nl * pr.snr_fac;             // false for NaN
          contrib = v - nl;
          if (want_out) {
            osn[k] = member ? kLog2ToDb * fast_log2(nl) : CUDART_NAN_F;
            osc[k] = good ? kLog2ToDb * fast_log2(contrib) : CUDART_NAN_F;
          }
        } else {
          good = (v == v);
          contrib = v;
        }
        if (want_out) {
          osv[k] = kLog2ToDb * fast_log2(v);
          orr[k] = member ? cc[k].rr : CUDART_NAN_F;
        }
        if (member) {
          s[k] += good ? contrib : 0.f;
          cn[k] += good ? 1 : (1 << 16);
        }
      }
      if (want_out) {
        const size_t o = ((size_t)(row0 + t) * R) + 4 * (size_t)j;
        if (pr.o_sv) st_stream4(reinterpret_cast<float4*>(pr.o_sv + o), make_float4(osv[0], osv[1], osv[2], osv[3]));
        if (pr.o_rng) st_stream4(reinterpret_cast<float4*>(pr.o_rng + o), make_float4(orr[0], orr[1], orr[2], orr[3]));
        if (pr.o_svn) st_stream4(reinterpret_cast<float4*>(pr.o_svn + o), make_float4(osn[0], osn[1], osn[2], osn[3]));
        if (pr.o_svc) st_stream4(reinterpret_cast<float4*>(pr.o_svc + o), make_float4(osc[0], osc[1], osc[2], osc[3]));
      }
    }
  }
  __syncthreads();
  // ---- flush the CTA accumulators: one float64 atomic triple per touched (ping bin, range bin) cell --------------
  for (int i = tid; i < Ta * nR; i += kThreads) {
    const Cell cl = s_cells[i];
    if (cl.cnt == 0) continue;
    const int t = i / nR, k = i % nR;
    double* cell = pr.acc + (((c * pr.nX + s_rows[t].xb) * (long long)nR) + k) * 4;
    const int good = cl.cnt & 0xffff, bad = cl.cnt >> 16;
    if (good) {
      atomicAdd(cell + 0, (double)cl.sum);
      atomicAdd(cell + 1, (double)good);
    }
    if (bad) atomicAdd(cell + 2, (double)bad);
  }
  __syncthreads();  // shared memory (tile, accumulators, mbarrier) is reused by the next tile
  }
}

size_t pipeline_smem(long long R, int nR, int tile, int do_noise, int staged) {
  size_t off = staged ? (size_t)tile * R * 4 : 0;
  if (do_noise) off += (size_t)R * 4 + (((size_t)R * 2 + 15) & ~(size_t)15);
  off += (size_t)(nR + 1) * 8;
  off += (size_t)tile * sizeof(RowS);
  off = (off + 15) & ~(size_t)15;
  off += (size_t)tile * nR * sizeof(Cell);
  off += (size_t)tile * (nR + 1) * 4;
  return off;
}

constexpr size_t kSmemMax = 227 * 1024;

}  // namespace

extern "C" epb_i64 epb_pipeline_smem_bytes(epb_i64 R, int nR, int tile, int do_noise, int staged) {
  return (epb_i64)pipeline_smem(R, nR, tile, do_noise, staged);
}

int epb_pipeline_fast_try(const void* x, int x_i16, const epb_row* rows, const int* xbin, const double* r_edges, int nR,
                          int closed_right, double* acc, float* noise_out, long long C, long long P, long long R,
                          long long nX, int ping_num, int range_sample_num, float noise_max_lin, float snr_lin,
                          double* range_max_out, int sv_input, void* workspace, long long workspace_bytes, cudaStream_t s,
                          float* o_sv = nullptr, float* o_rng = nullptr, float* o_svn = nullptr, float* o_svc = nullptr);
void epb_range_max_init_launch(double* out_max, cudaStream_t s);
void epb_range_max_gated_launch(const float* x, const epb_row* rows, long long nrows, int R, double* out_max, const int* gate,
                                cudaStream_t s);
long long epb_pipeline_fast_workspace(long long C, long long P, long long R, int ping_num);
void epb_ingest_gated_launch(const short* counts, float* out, long long n, const int* gate, cudaStream_t s);

extern "C" epb_i64 epb_pipeline_workspace_bytes(epb_i64 C, epb_i64 P, int ping_num) {
  if (C <= 0 || P <= 0 || ping_num < 0) return 256;
  return epb_pipeline_fast_workspace(C, P, 0, ping_num);
}

extern "C" epb_i64 epb_pipeline_workspace_bytes_r(epb_i64 C, epb_i64 P, epb_i64 R, int ping_num) {
  if (C <= 0 || P <= 0 || ping_num < 0) return 256;
  return epb_pipeline_fast_workspace(C, P, R, ping_num);
}

// counts: NULL, or the int16 raw power counts of which backscatter_r (then a scratch buffer) is the float32 image that
// the ingest kernel writes only when the general kernel has to run.
static int pipeline_power_mvbs_impl(const short* counts, float* backscatter_r, const epb_row* rows, const int* xbin,
                                    const double* r_edges, int nR, int closed_right, double* acc, float* noise_out,
                                    float* Sv, float* echo_range, float* Sv_noise, float* Sv_corrected, epb_i64 C,
                                    epb_i64 P, epb_i64 R, epb_i64 nX, int ping_num, int range_sample_num,
                                    float noise_max, float snr_threshold, double* range_max_out, void* workspace,
                                    epb_i64 workspace_bytes, void* stream) {
  EPB_REQUIRE(backscatter_r && rows && xbin && r_edges && acc, "NULL pointer");
  EPB_REQUIRE(C > 0 && P > 0 && R > 0 && R < (1 << 24) && nX > 0, "bad shape");
  EPB_REQUIRE(R % 4 == 0, "fused pipeline needs range_sample % 4 == 0 (use the separate kernels otherwise)");
  EPB_REQUIRE((((uintptr_t)backscatter_r | (uintptr_t)Sv | (uintptr_t)echo_range | (uintptr_t)Sv_noise |
                (uintptr_t)Sv_corrected) % 16) == 0,
              "arrays must be 16-byte aligned");
  EPB_REQUIRE(nR > 0 && nR <= 4096, "number of range bins must be in 1..4096");
  EPB_REQUIRE(ping_num >= 0 && (ping_num == 0 || range_sample_num > 0), "bad ping_num / range_sample_num");
  EPB_REQUIRE(ping_num <= kThreads, "ping_num must be <= 256");
  EPB_REQUIRE(ping_num > 0 || (!noise_out && !Sv_noise && !Sv_corrected), "noise outputs need ping_num > 0");
  Params pr;
  pr.x = backscatter_r, pr.rows = rows, pr.xbin = xbin, pr.edges = r_edges, pr.acc = acc, pr.noise_out = noise_out;
  pr.o_sv = Sv, pr.o_rng = echo_range, pr.o_svn = Sv_noise, pr.o_svc = Sv_corrected;
  pr.C = C, pr.P = P, pr.nX = nX, pr.R = (int)R, pr.nR = nR;
  pr.do_noise = ping_num > 0;
  pr.tile = pr.do_noise ? ping_num : 4;
  pr.rs_num = range_sample_num, pr.closed_right = closed_right;
  pr.noise_max_lin = (noise_max == noise_max) ? (float)pow(10.0, (double)noise_max / 10.0) : nanf("");
  pr.snr_fac = (float)(1.0 + pow(10.0, (double)snr_threshold / 10.0));
  pr.gate = nullptr;
  if (range_max_out) epb_range_max_init_launch(range_max_out, (cudaStream_t)stream);
  // fast path (pipeline_fast.cu): regular volumes without full-size outputs.  A device-side flag written by its
  // classification kernel decides which of the two kernels does the work; the other returns immediately.
  if (workspace && workspace_bytes >= 256 && ((uintptr_t)workspace % 16) == 0 &&
      epb_pipeline_fast_try(counts ? (const void*)counts : (const void*)backscatter_r, counts != nullptr, rows, xbin, r_edges, nR, closed_right, acc, noise_out, C, P, R, nX, ping_num,
                            range_sample_num, pr.noise_max_lin, (float)pow(10.0, (double)snr_threshold / 10.0),
                            range_max_out, 0, workspace, workspace_bytes, (cudaStream_t)stream, Sv, echo_range, Sv_noise,
                            Sv_corrected))
    pr.gate = (const int*)workspace;
  if (counts) epb_ingest_gated_launch(counts, backscatter_r, C * P * R, pr.gate, (cudaStream_t)stream);
  // stage the tile in shared memory when two CTAs per SM still fit, else when one fits, else stream from global
  pr.staged = 1;
  size_t smem = pipeline_smem(R, nR, pr.tile, pr.do_noise, 1);
  if (smem > kSmemMax) {
    pr.staged = 0;
    smem = pipeline_smem(R, nR, pr.tile, pr.do_noise, 0);
  }
  if (smem > kSmemMax) {
    epb_set_error("epb_pipeline_power_mvbs: tile accumulators need %zu bytes of shared memory (> %zu)", smem, kSmemMax);
    return EPB_E_UNSUPPORTED;
  }
  const long long nPt = (P + pr.tile - 1) / pr.tile;
  EPB_REQUIRE(C * nPt < (1LL << 31), "too many ping tiles");
  if (cudaFuncSetAttribute(pipeline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return epb_check_launch("epb_pipeline_power_mvbs(smem)");
  const long long cap = (long long)epb_num_sms() * 16;
  pipeline_kernel<<<(unsigned)(C * nPt < cap ? C * nPt : cap), kThreads, smem, (cudaStream_t)stream>>>(pr);
  // exact nanmax(echo_range): computed by the fast kernel when it runs, by the (gated) range kernel otherwise
  if (range_max_out) epb_range_max_gated_launch(backscatter_r, rows, C * P, (int)R, range_max_out, pr.gate, (cudaStream_t)stream);
  return epb_check_launch("epb_pipeline_power_mvbs");
}

extern "C" int epb_pipeline_power_mvbs(const float* backscatter_r, const epb_row* rows, const int* xbin,
                                       const double* r_edges, int nR, int closed_right, double* acc, float* noise_out,
                                       float* Sv, float* echo_range, float* Sv_noise, float* Sv_corrected, epb_i64 C,
                                       epb_i64 P, epb_i64 R, epb_i64 nX, int ping_num, int range_sample_num,
                                       float noise_max, float snr_threshold, double* range_max_out, void* workspace,
                                       epb_i64 workspace_bytes, void* stream) {
  return pipeline_power_mvbs_impl(nullptr, const_cast<float*>(backscatter_r), rows, xbin, r_edges, nR, closed_right, acc,
                                  noise_out, Sv, echo_range, Sv_noise, Sv_corrected, C, P, R, nX, ping_num,
                                  range_sample_num, noise_max, snr_threshold, range_max_out, workspace, workspace_bytes, stream);
}

extern "C" int epb_pipeline_power_mvbs_i16(const short* counts, float* scratch, const epb_row* rows, const int* xbin,
                                           const double* r_edges, int nR, int closed_right, double* acc, float* noise_out,
                                           epb_i64 C, epb_i64 P, epb_i64 R, epb_i64 nX, int ping_num, int range_sample_num,
                                           float noise_max, float snr_threshold, double* range_max_out, void* workspace,
                                           epb_i64 workspace_bytes, void* stream) {
  EPB_REQUIRE(counts && scratch, "NULL pointer");
  EPB_REQUIRE(((uintptr_t)counts % 16) == 0, "arrays must be 16-byte aligned");
  return pipeline_power_mvbs_impl(counts, scratch, rows, xbin, r_edges, nR, closed_right, acc, noise_out, nullptr, nullptr,
                                  nullptr, nullptr, C, P, R, nX, ping_num, range_sample_num, noise_max, snr_threshold,
                                  range_max_out, workspace, workspace_bytes, stream);
}
