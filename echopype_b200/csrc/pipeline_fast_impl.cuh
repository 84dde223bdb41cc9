// Fused pipeline, fast path: raw power -> Sv -> background-noise removal -> MVBS accumulators, one pass over HBM
// at 4 algorithmic bytes per sample, 2 on int16 raw counts (same semantics as pipeline.cu).
//
// PERSISTENT kernel (grid = SM count x resident CTAs: one CTA per SM at R = 4096, two at R <= 2048), each CTA owns a
// contiguous run of (channel, ping-tile) tiles.  Design points (DESIGN.md "fused pipeline, fast path"):
//   * a ring of tile slots in shared memory is filled by TMA bulk copies (cp.async.bulk: one copy for the rows of a
//     tile and one for its 144-byte descriptor from prepare_kernel, completing on the tile's mbarrier; SASS UBLKCP /
//     SYNCS), issued as soon as a slot is free, so HBM latency and the per-tile reductions overlap;
//   * a thread owns G x 4 adjacent range samples (one LDS.128 per row and group).  u = 10^((Sv - TL)/10) =
//     2^(x c1 + c0 + lg[column]) of the whole tile (ping_num <= 8 rows) lives in REGISTERS between the noise estimate
//     (phase 1) and the noise removal / binning (phase 2): FFMA2 + one MUFU.EX2 per sample, shared memory is read
//     once per sample;
//   * the range-only terms (lg = log2(h / TL), TL) and the range-bin boundaries are computed once per range law
//     (normally once per channel) and parked in shared memory;
//   * noise estimate: column sums of u -> two partial sums per column group (toward the range tile of its first
//     column and toward the next one) -> L lanes per range tile -> per-warp minima -> warp reduction;
//   * noise removal in the u domain: the sample survives iff u > noise (1 + 10^(SNR/10)) (one tile-wide compare), the
//     surviving u are summed per column and scaled once per tile: sum(Sv_corrected_lin) = TL (sum(u) - n noise);
//   * (sum, count) of a thread's columns accumulate in registers ACROSS tiles while the ping bin does not change; on
//     a bin change the column-group partial sums go through shared memory and eight lanes per range bin issue ONE
//     float64 atomic triple per (CTA, range bin).
// Handles the regular case: every tile's rows share one range law and have finite calibration constants (checked
// on the device by prepare_kernel; otherwise the general kernel of pipeline.cu runs instead), R % 4 == 0 (8 for int16
// input), R <= 4096, range_sample_num >= 4.  Full-size outputs (keep=: Sv, echo_range, Sv_noise, Sv_corrected) come from
// the kKeep instantiations, which stream them out of the register tile between the noise estimate and the binning.
//
// ping_num > 8 (kSweep; the reference's own test setting is remove_background_noise(ping_num=10, range_sample_num=20),
// tests/utils/test_processinglevels_integration.py:111): u of a whole noise tile no longer fits the registers and the tile
// no longer fits the ring (10 x 4096 x 4 B = 160 KB).  The noise tile is cut into S = ceil(ping_num / 8) sub-tiles of T
// rows that stream through the same ring TWICE: sweep 1 adds up the column sums of u (the noise estimate needs nothing
// else), sweep 2 fetches the same rows again - from L2, where sweep 1 left them a few microseconds earlier, so HBM is
// still read once - recomputes u and does the noise removal / binning with the now known noise.
#pragma once
#include "pipeline_common.cuh"

int epb_grid_reserve();  // pipeline_fast.cu

namespace {
using namespace epb;

constexpr int kMaxT = 8;
constexpr int kMaxTilesInFlight = 8;
constexpr int kFlushRows = 248;  // packed 8-bit per-column counters: flush a cell before a counter can overflow
constexpr unsigned kInfBits = 0x7f800000u;

struct TileInfo;

struct FastParams {
  const void* x;  // float32 samples, or int16 raw power counts (kI16; -32768 marks padding)
  const epb_row* rows;
  const int* xbin;
  const double* edges;
  double* acc;
  float* noise_out;
  double* rmax;  // NULL or exact nanmax(echo_range) (atomic max; initialised by the caller)
  float *o_sv, *o_rng, *o_svn, *o_svc;  // kKeep: optional full-size outputs Sv / echo_range / Sv_noise / Sv_corrected [C,P,R]
  const int* irregular;  // workspace flag from prepare_kernel: != 0 -> this kernel does nothing
  const TileInfo* tiles;  // [ntiles] descriptors from prepare_kernel (workspace)
  long long C, P, nX, ntiles;
  int R, nR, rs_num, closed_right, nslots, nPt;
  int PN, S;  // rows of a noise tile (ping_num) and sub-tiles per noise tile (1 unless kSweep); descriptors: [ntiles][S]
  int rt_lanes, rt_tpw, rt_recip;  // noise estimate: lanes per range tile, range tiles per warp, ceil(2^16 / lanes)
  float noise_max_lin;  // NaN: no cap
  float snr1;           // 1 + 10^(SNR/10)
  int sv_input;         // the input is Sv in dB (bin reduction of compute_MVBS): e = 10^(Sv/10), h = 1, NaN = NaN member
};

struct TileInfo {  // 144 bytes; written per tile by prepare_kernel, fetched by TMA together with the tile's rows
  float2 rc[kMaxT];              // per-row (c0, c1): e = 2^(x c1 + c0)
  int run_cell[kMaxT];           // c * nX + ping bin, or -1 (ping outside every bin)
  unsigned char run_end[kMaxT];  // rows [run_end[r-1], run_end[r]) share one accumulator cell
  int nruns;
  int Ta;      // rows present in the tile
  int lawchg;  // the tile's range law differs from the previous tile's (or first tile of a channel)
  int rcsame;  // every row of the tile has the same (c0, c1)
  long long row0;  // first (channel, ping) row of the tile
  long long pad1, pad2;
};
static_assert(sizeof(TileInfo) == 144 && sizeof(TileInfo) % 16 == 0, "TileInfo is moved by cp.async.bulk");

// bitwise comparison of the range law of two rows (exact float64 law + value-form splits)
__device__ __forceinline__ bool same_law(const epb_row& a, const epb_row& b) {
  return a.p0 == b.p0 && a.p1 == b.p1 && a.p2 == b.p2 && a.p3 == b.p3 && a.p4 == b.p4 && a.off1 == b.off1 &&
         a.off2 == b.off2 && a.r0 == b.r0 && a.a == b.a && a.two_alpha == b.two_alpha && a.n_start == b.n_start &&
         a.law == b.law && a.azfp_N == b.azfp_N && a.a_h == b.a_h && a.a_l == b.a_l && a.r0_h == b.r0_h &&
         a.r0_l == b.r0_l && a.bp_h == b.bp_h && a.bp_l == b.bp_l && a.c2 == b.c2;
}

// One thread per tile: build the tile descriptor and flag volumes the fast kernel cannot take (a tile whose rows do
// not share one range law, or rows with NaN calibration constants).  64 us on cfg2 (3.7 % of the step); staging the
// 192-byte row records through shared memory for coalesced loads measured slower (81 us: too few loads in flight).
__global__ void prepare_kernel(const epb_row* __restrict__ rows, const int* __restrict__ xbin, long long P, long long nX,
                               int T, int PN, int S, int nPt, long long ndesc, int sv_input, TileInfo* __restrict__ tiles,
                               int* __restrict__ irregular) {
  const long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (d >= ndesc) return;
  const long long g = d / S;  // noise tile
  const int ksub = (int)(d - g * S);
  const long long c = g / nPt;
  const int itile = (int)(g - c * nPt);
  const long long pt0 = (long long)itile * PN;  // first row of the noise tile
  const long long p0 = pt0 + (long long)ksub * T;
  long long pend = pt0 + ((ksub + 1) * T < PN ? (ksub + 1) * T : PN);
  pend = pend < P ? pend : P;
  const int Ta = (int)(pend > p0 ? pend - p0 : 0);
  const epb_row* r0 = rows + c * P + p0;
  const epb_row* rt0 = rows + c * P + pt0;  // every row of a noise tile must share the law of its first row
  TileInfo ti;
  bool bad = false;
  int nruns = 0, prev_xb = 0, rcsame = 1;
  for (int t = 0; t < kMaxT; ++t) {
    ti.rc[t] = make_float2(0.f, 0.f);
    ti.run_cell[t] = -1;
    ti.run_end[t] = 0;
  }
  for (int t = 0; t < Ta; ++t) {
    const epb_row& r = r0[t];
    if (!sv_input && !(r.c0 == r.c0 && r.c1 == r.c1)) bad = true;
    if (!same_law(rt0[0], r)) bad = true;  // NaN laws never compare equal
    ti.rc[t] = sv_input ? make_float2(0.f, kDb2Log2) : make_float2(r.c0, r.c1);  // Sv input: e = 2^(Sv log2(10)/10)
    if (!(ti.rc[t].x == ti.rc[0].x && ti.rc[t].y == ti.rc[0].y)) rcsame = 0;
    int xb = xbin[p0 + t];
    if (xb < 0 || xb >= nX) xb = -1;
    if (t == 0 || xb != prev_xb) {
      ti.run_cell[nruns] = (xb >= 0) ? (int)(c * nX + xb) : -1;
      ++nruns;
    }
    ti.run_end[nruns - 1] = (unsigned char)(t + 1);
    prev_xb = xb;
  }
  ti.nruns = nruns;
  ti.Ta = Ta;
  ti.lawchg = (ksub == 0) && ((itile == 0) || !same_law(rt0[0], *(rt0 - PN)));
  ti.rcsame = rcsame;
  ti.row0 = c * P + p0;
  ti.pad1 = 0, ti.pad2 = 0;
  tiles[d] = ti;
  if (bad) *irregular = 1;
}

template <int G>
struct Acc {  // per-thread accumulators of the owned columns (G groups of four) for the current (channel, ping bin) cell
  float s[G][4];     // sum of surviving 10^(Sv_corrected/10)
  float good[G][4];  // number of surviving samples (exact in float: < 2^24)
  unsigned nanm[G];  // 4 x 8 bit: rows whose sample is not a member (NaN echo_range)
  int rows;          // rows accumulated into this cell (<= kFlushRows, bounds the 8-bit fields)
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s[g][k] = 0.f, good[g][k] = 0.f;
      nanm[g] = 0u;
    }
    rows = 0;
  }
};

__device__ __forceinline__ void atomic_cell(double* cell, float s, int good, int bad) {
  if (good) {
    atomicAdd(cell + 0, (double)s);
    atomicAdd(cell + 1, (double)good);
  }
  if (bad) atomicAdd(cell + 2, (double)bad);
}

// A + (columns k >= b), B = (columns k < b ? 0 : v): split of a group's four column values at position b (1..4)
__device__ __forceinline__ float2 split_sum(const float (&v)[4], int b) {
  const float A = v[0] + ((b > 1) ? v[1] : 0.f) + (((b > 2) ? v[2] : 0.f) + ((b > 3) ? v[3] : 0.f));
  const float B = ((b > 1) ? 0.f : v[1]) + (((b > 2) ? 0.f : v[2]) + ((b > 3) ? 0.f : v[3]));
  return make_float2(A, B);
}

struct Producer {  // TMA issue cursor, used by one thread only (kept in shared memory, not in registers)
  int tile;        // next local tile to issue
  int ts;          // its tile slot (tile % NT)
  int ds;          // its descriptor slot (tile % (NT + 1))
  int c, it;       // channel / ping tile of `tile`
  int k, ph;       // kSweep: sub-tile and sweep (0 / 1) of the next item; `tile` then counts items
  int li;          // kSweep: local noise tile of the next item
};

// Range-only column terms of the u domain, computed once per range law with the accurate libm variants:
//   lg = log2(h / TL) = 2 log2(R'/Rm) + c2 (R' - R),  TL = Rm^2 2^(c2 R)   (Rm = max(R, 1), clean/api.py:392-431)
// Columns where Sv is undefined (n < n_start, R' < 0) get lg = -inf (u = 0) and TL = 0 (the "undefined" marker).
struct ColT {
  float lg, tl;
};
__device__ __forceinline__ ColT col_tables(const RowF& r, int n) {
  const float nf = (float)n;
  const float rp = tvg_range_of(r, nf);
  const float rr = range_of(r, nf);
  const float rm = (rr >= 1.f) ? rr : 1.f;
  ColT c;
  c.lg = fmaf(2.f, log2f(rp / rm), r.c2 * (rp - rr));
  c.tl = (rm * rm) * exp2f(r.c2 * rr);
  if (!(n >= r.n_start) || !(rp >= 0.f)) c.lg = -CUDART_INF_F, c.tl = 0.f;
  return c;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ bool finite_f(float x) { return x * 0.f == 0.f; }

// T rows per tile (ping_num), G column groups of four per thread (threads = R / (4 G))
#ifndef EPB_G1_THREADS
#define EPB_G1_THREADS 256  // one column group per thread up to this many threads (R <= 1024); wider rows take two groups: R = 2048 as 2 x 256 threads x 128 registers with two resident CTAs per SM measured 1.58 ms vs 1.97 ms for 1 x 512 x 64 registers (4x200000x2048)
#endif
#ifndef EPB_GBIG
#define EPB_GBIG 2  // column groups per thread of the wide variant (R > 4 * EPB_G1_THREADS): 2 -> 512 threads, 4 -> 256
#endif
#ifndef EPB_G1_BLOCKS
#define EPB_G1_BLOCKS 2  // resident CTAs per SM of the one-group variant (R <= 2048): 64 registers per thread
#endif
// raw power counts (int16) -> dB as float32, exactly as the ingest kernel / convert/parse_base.py:24,302 do it:
// the float32 nearest to count * 10 log10(2) / 256 (count_to_db_f).  The 16-bit payloads are spliced into the mantissa of 2^23 (PRMT), so the
// conversion runs on the integer / FMA pipes and leaves the MUFU (where I2F lives) to the ex2 of phase 0.
__device__ __forceinline__ float4 counts_to_db(uint2 w) {
  const unsigned a = w.x ^ 0x80008000u, b = w.y ^ 0x80008000u;  // offset binary: payload = count + 32768
  const float kMagic = 8388608.f + 32768.f;
  float2 f01 = make_float2(__uint_as_float(__byte_perm(a, 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(a, 0x4B000000u, 0x7632)));
  float2 f23 = make_float2(__uint_as_float(__byte_perm(b, 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(b, 0x4B000000u, 0x7632)));
  const float2 mm = make_float2(-kMagic, -kMagic), hi = make_float2(kIndex2PowerHi, kIndex2PowerHi), lo = make_float2(kIndex2PowerLo, kIndex2PowerLo);
  f01 = fadd2(f01, mm), f23 = fadd2(f23, mm);
  f01 = ffma2(f01, hi, fmul2(f01, lo));  // count_to_db_f on pairs
  f23 = ffma2(f23, hi, fmul2(f23, lo));
  return make_float4(f01.x, f01.y, f23.x, f23.y);
}

template <int T, int G, bool kNoise, bool kI16, bool kSweep = false, bool kKeep = false>
__global__ void __launch_bounds__(G == 1 ? EPB_G1_THREADS : 512 / (EPB_GBIG / 2), (G == 1 && EPB_G1_THREADS == 512) ? EPB_G1_BLOCKS : 1)
    pipeline_fast_kernel(const FastParams pr) {
  static_assert(!kSweep || kNoise, "two sweeps only make sense with the noise estimate");
  static_assert(!kKeep || !kI16, "full-size outputs are produced from the float32 image");
  if (*pr.irregular) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_full[kMaxTilesInFlight];  // one mbarrier per tile slot
  __shared__ __align__(16) TileInfo s_tile[kMaxTilesInFlight + 1];  // NT + 1 in rotation: a descriptor outlives its ring slot
  constexpr int kMaxWarps = (EPB_G1_THREADS > 512 ? EPB_G1_THREADS : 512) / 32;
  __shared__ __align__(16) unsigned int s_wmin[kMaxWarps];  // per warp: minimum range-tile mean of the current tile (float bits)
  __shared__ int s_hasnan[2];   // a thread saw a NaN sample in the tile: range-tile counts are corrected by s_def
  __shared__ Producer s_prod;
  __shared__ int s_multi;  // the current law has a column group of four in more than two range bins
  __shared__ int s_last[kMaxT];  // last sample with a defined range of the rows whose final sample is NaN (rare)
  const int R = pr.R, nR = pr.nR, NT = pr.nslots;  // NT tile slots of T rows in the ring
  const int tid = threadIdx.x;
  const int nth = blockDim.x;
  const int lane = tid & 31;
  const bool prod_warp = (tid >> 5) == (nth >> 5) - 1;  // the last warp doubles as descriptor / TMA producer
  const int nRt = kNoise ? (R + pr.rs_num - 1) / pr.rs_num : 0;
  // ---- dynamic shared memory ----------------------------------------------------------------------------------------
  // [lg R][tl R][ga R/4][gb R/4][keys R int16][ctl R/4 uint16][pad to 16][ring NT x T x R][edges nR+1 f64][bounds nR+1]
  // [valid nRt][def 2 nRt]
  float* const s_lg = reinterpret_cast<float*>(smem_raw);  // log2(h / TL): u = 2^(x c1 + c0 + lg) = 10^((Sv - TL)/10);
                                                           // -inf where Sv is undefined (u = 0: never survives)
  float* const s_tl = s_lg + R;                            // TL = 10^(TL_dB/10) >= 1;  0 where Sv is undefined
  float* const s_ga = s_tl + R;                            // per column group: sum of u toward the range tile of its first
  float* const s_gb = s_ga + (R >> 2);                     // column / toward the following range tile
  float2* const s_fs = reinterpret_cast<float2*>(s_ga);    // flush (aliases ga/gb): per column group, cell sums toward the
                                                           // range bin of its first column (.x) / of its last column (.y)
  uint2* const s_fc = reinterpret_cast<uint2*>(s_gb + (R >> 2));  // flush: survivor counts (.x) and non-member counts
                                                                  // (.y), first-bin part | last-bin part << 16
  unsigned char* const s_bsp = reinterpret_cast<unsigned char*>(s_fc + (R >> 2));  // leading columns of a group that
                                                                                   // share the range bin of the first one
  constexpr int kXB = kI16 ? 2 : 4;  // bytes per input sample
  unsigned char* const s_ring = smem_raw + (((size_t)R * 12 + (size_t)R / 4 + 15) & ~(size_t)15);
  double* const s_edges = reinterpret_cast<double*>(s_ring + (((size_t)NT * T * R * kXB + 15) & ~(size_t)15));
  int* const s_bounds = reinterpret_cast<int*>(s_edges + (nR + 1));
  int* const s_valid = s_bounds + (nR + 1);  // columns of each range tile with a defined Sv (n >= n_start, R' >= 0)
  int* const s_def = s_valid + nRt;          // [2][nRt] samples missing (NaN) from each range tile, by tile parity
  // kKeep: log2(TL) (also where Sv is undefined) and echo_range of every column under the current law
  float* const s_ltl = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_def + 2 * nRt) + 15) & ~(uintptr_t)15);
  float* const s_rr = s_ltl + R;

  // ---- tile range of this CTA -----------------------------------------------------------------------------------
  const long long g0 = pr.ntiles * (long long)blockIdx.x / gridDim.x;
  const int ntl = (int)(pr.ntiles * (long long)(blockIdx.x + 1) / gridDim.x - g0);  // local tiles 0..ntl-1
  if (ntl <= 0) return;
  const int nPt = pr.nPt;
  const uint32_t row_bytes = (uint32_t)R * (uint32_t)kXB;
  // one thread: issue every tile whose slot is free (the tile NT before it has been consumed)
  const int S = kSweep ? pr.S : 1;
  const int nitems = kSweep ? ntl * 2 * S : ntl;  // ring items: tiles, or (noise tile, sweep, sub-tile)
  auto issue_tiles = [&](int last_done) {
    Producer p = s_prod;
    while (p.tile < nitems && p.tile - NT <= last_done) {
      long long p0;
      int Ta;
      const TileInfo* desc;
      if (!kSweep) {
        p0 = (long long)p.it * T;
        Ta = (int)((p0 + T <= pr.P) ? T : (pr.P - p0));
        desc = pr.tiles + g0 + p.tile;
      } else {
        const long long pt0 = (long long)p.it * pr.PN;
        p0 = pt0 + (long long)p.k * T;
        long long pend = pt0 + ((p.k + 1) * T < pr.PN ? (p.k + 1) * T : pr.PN);
        pend = pend < pr.P ? pend : pr.P;
        Ta = (int)(pend > p0 ? pend - p0 : 0);
        desc = pr.tiles + (g0 + p.li) * S + p.k;
      }
      const unsigned char* src = reinterpret_cast<const unsigned char*>(pr.x) + ((long long)p.c * pr.P + p0) * (long long)row_bytes;
      unsigned long long* bar = &s_full[p.ts];
      unsigned char* dst = s_ring + (size_t)p.ts * T * row_bytes;
      mbar_expect_tx(bar, row_bytes * (uint32_t)Ta + (uint32_t)sizeof(TileInfo));
      bulk_g2s(&s_tile[p.ds], desc, (uint32_t)sizeof(TileInfo), bar);
      if (Ta > 0) bulk_g2s(dst, src, row_bytes * (uint32_t)Ta, bar);  // the rows of a tile are contiguous on both sides
      ++p.tile;
      if (++p.ts == NT) p.ts = 0;
      if (++p.ds == NT + 1) p.ds = 0;
      if (!kSweep) {
        if (++p.it == nPt) p.it = 0, ++p.c;
      } else if (++p.k == S) {
        p.k = 0;
        if (++p.ph == 2) {
          p.ph = 0, ++p.li;
          if (++p.it == nPt) p.it = 0, ++p.c;
        }
      }
    }
    s_prod = p;
  };

  // ---- prologue -----------------------------------------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kMaxTilesInFlight; ++i) mbar_init(&s_full[i], 1);
    mbar_init_fence();
    for (int i = 0; i < kMaxWarps; ++i) s_wmin[i] = kInfBits;
    s_hasnan[0] = 0, s_hasnan[1] = 0;
    for (int t = 0; t < kMaxT; ++t) s_last[t] = -1;
    const int c0 = (int)(g0 / nPt), it0 = (int)(g0 - (long long)c0 * nPt);
    s_prod.tile = 0, s_prod.ts = 0, s_prod.ds = 0, s_prod.c = c0, s_prod.it = it0;
    s_prod.k = 0, s_prod.ph = 0, s_prod.li = 0;
  }
  for (int k = tid; k <= nR; k += nth) s_edges[k] = pr.edges[k];
  for (int k = tid; k < 2 * nRt; k += nth) s_def[k] = 0;
  __syncthreads();
  if (prod_warp && lane == 0) issue_tiles(-1);

  // group g of this thread owns columns n0 + g * 4 * nth .. +3; threads past the row end work on column 0 (their
  // results are never stored: no partial-sum store, not stored at flush), so the hot loads carry no predicates
  int colg[G];
  bool liveg[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int n = 4 * tid + g * 4 * nth;
    liveg[g] = n < R;
    colg[g] = liveg[g] ? n : 0;
  }
  int bsplit[G];  // columns k < bsplit[g] of group g lie in the range tile of the group's first column
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int rs = kNoise ? pr.rs_num : 4;
    const int b = (colg[g] / rs + 1) * rs - colg[g];
    bsplit[g] = b < 4 ? b : 4;
  }
  bool nanrange = false;
  // exact nanmax(echo_range): the thread that owns the last column watches the final sample of every row
  const int last_group = (R >> 2) - 1;
  const bool is_last = pr.rmax != nullptr && tid == last_group % nth;
  // watcher state lives in shared memory (one thread uses it; as registers it would cost every thread four)
  __shared__ double s_range_last, s_rmax_local;  // range of the final sample under the current law; running maximum
  __shared__ int s_seen_full;                     // some row under the current law had a defined final sample
  if (is_last) s_range_last = -CUDART_INF, s_rmax_local = -CUDART_INF, s_seen_full = 0;
  Acc<G> acc;
  acc.clear();
  int cur_cell = -1;
  int ts = 0;        // tile slot of the current tile and the phase parity of its mbarrier
  unsigned par = 0;
  int tsel = 0;      // li % (NT + 1): descriptor slot

  // Add the register cell (per-column sums / survivor counts of acc.rows rows) to the accumulator grid.  Called by
  // the whole CTA (the conditions are tile properties).  Column-group partial sums go through shared memory, eight
  // lanes per range bin add the groups of the bin and one of them issues the float64 atomics: one triple per
  // (CTA, range bin).  Laws with a group of four columns in more than two bins take per-column atomics instead.
  auto flush = [&]() {
    double* acc_row = pr.acc + (long long)cur_cell * nR * 4;
    if (s_multi) {
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (liveg[g]) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int key = key_of(s_bounds, nR, colg[g] + k);
            if (key >= 0) {
              const int good = (int)acc.good[g][k];
              atomic_cell(acc_row + 4 * (long long)key, acc.s[g][k], good,
                          acc.rows - (int)((acc.nanm[g] >> (8 * k)) & 0xffu) - good);
            }
          }
        }
      acc.clear();
      __syncthreads();  // s_multi is rewritten at the next law change
      return;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int grp = colg[g] >> 2;
      const int b = s_bsp[grp];
      const float2 fs = split_sum(acc.s[g], b), fg = split_sum(acc.good[g], b);
      unsigned nn = 0u;
      if (acc.nanm[g] != 0u) {  // rare: samples that are no members of any bin (NaN echo_range)
        const unsigned m = acc.nanm[g];
        const unsigned lo = (b > 3) ? m : (m & ((1u << (8 * b)) - 1u)), hi = m ^ lo;
        nn = ((lo & 0xffu) + ((lo >> 8) & 0xffu) + ((lo >> 16) & 0xffu) + (lo >> 24)) |
             (((hi & 0xffu) + ((hi >> 8) & 0xffu) + ((hi >> 16) & 0xffu) + (hi >> 24)) << 16);
      }
      if (liveg[g]) {
        s_fs[grp] = fs;
        s_fc[grp] = make_uint2((unsigned)fg.x | ((unsigned)fg.y << 16), nn);
      }
    }
    __syncthreads();
    {
      const int q = tid & 7;
      const int rows = acc.rows;
      for (int k = tid >> 3; k < ((nR + 3) & ~3); k += nth >> 3) {  // warp-uniform trip count
        const bool in = k < nR;
        const int c0 = in ? s_bounds[k] : 0, c1 = in ? s_bounds[k + 1] : 0;
        const int ga = (c0 + 3) >> 2, gb = (c1 + 3) >> 2;
        float sum = 0.f;
        unsigned cg = 0u, cn = 0u;
        if (q == 0 && (c0 & 3) && c1 > c0) {  // the group that straddles into the bin
          sum = s_fs[ga - 1].y;
          const uint2 c = s_fc[ga - 1];
          cg = c.x >> 16, cn = c.y >> 16;
        }
        for (int gq = ga + q; gq < gb; gq += 8) {
          sum += s_fs[gq].x;
          const uint2 c = s_fc[gq];
          cg += c.x & 0xffffu, cn += c.y & 0xffffu;
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          cg += __shfl_xor_sync(0xffffffffu, cg, o);
          cn += __shfl_xor_sync(0xffffffffu, cn, o);
        }
        if (q == 0 && c1 > c0) {
          const int members = rows * (c1 - c0) - (int)cn;
          atomic_cell(acc_row + 4 * (long long)k, sum, (int)cg, members - (int)cg);
        }
      }
    }
    __syncthreads();  // s_fs aliases the column-group sums of the next tile's noise estimate
    acc.clear();
  };

  // rows whose final sample is NaN (need bit t): every thread offers the last column it holds with a valid sample
  // (e >= 0: the NaN sentinel is -2), thread 0 then evaluates the exact range law there
  auto offer_last = [&](unsigned need, const float (&e)[G][T][4]) {
#pragma unroll
    for (int t = 0; t < T; ++t)
      if ((need >> t) & 1u) {
        int best = -1;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (liveg[g] && e[g][t][k] >= 0.f) best = colg[g] + k;
        if (best >= 0) atomicMax(&s_last[t], best);
      }
  };
  auto settle_last = [&](unsigned need, long long row0) {  // thread 0
    for (int t = 0; t < T; ++t)
      if ((need >> t) & 1u) {
        const int n = s_last[t];
        s_last[t] = -1;
        if (n >= 0) {
          const epb_row r = pr.rows[row0 + t];
          const double v = law_range(r, n);
          if (v == v) atomic_max_d(pr.rmax, v);
        }
      }
  };

  float se_acc[kSweep ? G : 1][4];  // kSweep: column sums of u over the sub-tiles of the current noise tile (sweep 1)
#pragma unroll
  for (int g = 0; g < (kSweep ? G : 1); ++g)
#pragma unroll
    for (int k = 0; k < 4; ++k) se_acc[g][k] = 0.f;
  int rows_acc = 0;        // kSweep: rows of the current noise tile seen in sweep 1
  float noise_keep = 0.f;  // kSweep: the noise of the current noise tile (linear), known after sweep 1
  int swli = 0, swk = 0, swph = 0;  // kSweep: noise tile, sub-tile, sweep of the current item
  for (int qi = 0; qi < nitems; ++qi) {
    const int li = kSweep ? swli : qi;
    const bool first_of_tile = !kSweep || (swph == 0 && swk == 0);
    const int it = li & 1;
    // ---- wait for the tile (rows + descriptor) ---------------------------------------------------------------------
    mbar_wait(&s_full[ts], par);
    const unsigned char* tbase = s_ring + (size_t)ts * T * row_bytes;
    unsigned mn16 = 0x7fff7fffu;  // kI16: packed minimum of the thread's counts (finds the padding marker)
    auto ld4 = [&](int t, int g) -> float4 {
      if (!kI16) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(tbase) + t * R + colg[g]);
      const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const short*>(tbase) + t * R + colg[g]);
      mn16 = __vimin3_s16x2(mn16, w.x, w.y);  // VIMNMX3.S16x2
      return counts_to_db(w);
    };
    if (++ts == NT) ts = 0, par ^= 1u;
    const TileInfo* ti = &s_tile[tsel];
    if (++tsel == NT + 1) tsel = 0;
    const int Ta = ti->Ta;

    // ---- new range-law segment: flush, recompute boundaries, column terms, keys ------------------------------------
    if (first_of_tile && (ti->lawchg || qi == 0)) {
      if (cur_cell >= 0) flush();
      cur_cell = -1;
      if (tid == 0) s_multi = 0;
      __syncthreads();  // every warp has used the old tables
      const long long row0 = ti->row0;
      const epb_row& lr = pr.rows[row0];
      for (int k = tid; k <= nR; k += nth) s_bounds[k] = first_at_or_above(lr, R, s_edges[k], pr.closed_right);
      const RowF rf = load_rowf(pr.rows + row0);
      nanrange = rf.nanrange && !pr.sv_input;
      if (is_last) {
        if (s_seen_full && s_range_last == s_range_last) s_rmax_local = fmax(s_rmax_local, s_range_last);
        s_range_last = pr.rows[row0].range_last;
        s_seen_full = 0;
      }
      for (int n = 4 * tid; n < R; n += 4 * nth) {
        float lg[4], tl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const ColT ct = col_tables(rf, n + k);
          lg[k] = pr.sv_input ? 0.f : ct.lg;
          tl[k] = pr.sv_input ? 1.f : ct.tl;
        }
        *reinterpret_cast<float4*>(s_lg + n) = make_float4(lg[0], lg[1], lg[2], lg[3]);
        *reinterpret_cast<float4*>(s_tl + n) = make_float4(tl[0], tl[1], tl[2], tl[3]);
        if (kKeep) {
          float lt[4], rr[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            rr[k] = range_of(rf, (float)(n + k));
            const float rm = (rr[k] >= 1.f) ? rr[k] : 1.f;
            lt[k] = fmaf(2.f, log2f(rm), rf.c2 * rr[k]);
          }
          *reinterpret_cast<float4*>(s_ltl + n) = make_float4(lt[0], lt[1], lt[2], lt[3]);
          *reinterpret_cast<float4*>(s_rr + n) = make_float4(rr[0], rr[1], rr[2], rr[3]);
        }
      }
      __syncthreads();
      for (int n = 4 * tid; n < R; n += 4 * nth) {
        int kk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) kk[k] = key_of(s_bounds, nR, n + k);
        const int b = (kk[1] != kk[0]) ? 1 : (kk[2] != kk[0]) ? 2 : (kk[3] != kk[0]) ? 3 : 4;
        const int changes = (kk[1] != kk[0]) + (kk[2] != kk[1]) + (kk[3] != kk[2]);
        s_bsp[n >> 2] = (unsigned char)b;
        if (changes > 1) s_multi = 1;  // bins narrower than the group: per-column flush for this law
      }
      for (int rt = tid; rt < nRt; rt += nth) {
        const int j0 = rt * pr.rs_num, j1 = (j0 + pr.rs_num < R) ? j0 + pr.rs_num : R;
        int nv = 0;
        for (int j = j0; j < j1; ++j) nv += (s_tl[j] != 0.f);
        s_valid[rt] = nv;
      }
      // no barrier needed here: s_bsp is thread-private, s_valid / s_multi are read after barrier (A)
    }

    // ---- u = 10^((Sv - TL)/10) -> registers ---------------------------------------------------------------------------
    float e[G][T][4];
    if (ti->rcsame) {  // the usual case: one (c0, c1) for the tile, the column term folds into the FFMA2 addend
      const float2 rc = ti->rc[0];
      const float2 c1p = make_float2(rc.y, rc.y);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 l4 = *reinterpret_cast<const float4*>(s_lg + colg[g]);
        const float2 b01 = make_float2(l4.x + rc.x, l4.y + rc.x), b23 = make_float2(l4.z + rc.x, l4.w + rc.x);
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float4 v = ld4(t, g);
          const float2 a01 = ffma2(make_float2(v.x, v.y), c1p, b01), a23 = ffma2(make_float2(v.z, v.w), c1p, b23);
          e[g][t][0] = fast_exp2(a01.x);
          e[g][t][1] = fast_exp2(a01.y);
          e[g][t][2] = fast_exp2(a23.x);
          e[g][t][3] = fast_exp2(a23.y);
        }
      }
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 l4 = *reinterpret_cast<const float4*>(s_lg + colg[g]);
        const float2 l01 = make_float2(l4.x, l4.y), l23 = make_float2(l4.z, l4.w);
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float2 rc = ti->rc[t];  // rows beyond Ta: stale constants and slot data, zeroed below
          const float4 v = ld4(t, g);
          const float2 c0p = make_float2(rc.x, rc.x), c1p = make_float2(rc.y, rc.y);
          const float2 a01 = fadd2(ffma2(make_float2(v.x, v.y), c1p, c0p), l01);
          const float2 a23 = fadd2(ffma2(make_float2(v.z, v.w), c1p, c0p), l23);
          e[g][t][0] = fast_exp2(a01.x);
          e[g][t][1] = fast_exp2(a01.y);
          e[g][t][2] = fast_exp2(a23.x);
          e[g][t][3] = fast_exp2(a23.y);
        }
      }
    }
    if (Ta < T) {  // partial tile (end of a channel): the missing rows contribute nothing
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (t >= Ta) {
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int k = 0; k < 4; ++k) e[g][t][k] = 0.f;
        }
    }
    float se[G][4];
    float chk = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < 4; k += 2) {  // column pairs: FADD2
        float2 acc2 = make_float2(e[g][0][k], e[g][0][k + 1]);
#pragma unroll
        for (int t = 1; t < T; ++t) acc2 = fadd2(acc2, make_float2(e[g][t][k], e[g][t][k + 1]));
        se[g][k] = acc2.x, se[g][k + 1] = acc2.y;
        chk += acc2.x + acc2.y;
      }
    // Rare: some sample of this thread is NaN (padded ping) or e overflowed.  Remember where (bit t*4+k per group),
    // make the column sums NaN-free, tell the range-tile reducer how many samples are missing, and replace the
    // sample by -2 (never survives a threshold >= -1).
    unsigned nanmask[G];
#pragma unroll
    for (int g = 0; g < G; ++g) nanmask[g] = 0u;
    if (kI16 && ((mn16 & 0xffffu) == 0x8000u || (mn16 >> 16) == 0x8000u)) {  // padding marker seen: those samples are NaN
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const short*>(tbase) + t * R + colg[g]);
          if ((w.x & 0xffffu) == 0x8000u) e[g][t][0] = CUDART_NAN_F;
          if ((w.x >> 16) == 0x8000u) e[g][t][1] = CUDART_NAN_F;
          if ((w.y & 0xffffu) == 0x8000u) e[g][t][2] = CUDART_NAN_F;
          if ((w.y >> 16) == 0x8000u) e[g][t][3] = CUDART_NAN_F;
        }
      chk = CUDART_NAN_F;
    }
    if (!finite_f(chk)) {
      if (kNoise) atomicOr(&s_hasnan[it], 1);
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          se[g][k] = 0.f;
          int missing = 0;
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const float v = e[g][t][k];
            const bool ok = finite_f(v);
            if (t < Ta) {
              se[g][k] += ok ? v : 0.f;
              missing += !ok;
              if (!ok) nanmask[g] |= 1u << (4 * t + k);
            }
            e[g][t][k] = ok ? v : -2.f;
          }
          if (kNoise && (!kSweep || swph == 0) && missing && liveg[g] && s_tl[colg[g] + k] != 0.f)
            atomicAdd(&s_def[it * nRt + (colg[g] + k) / pr.rs_num], missing);
        }
    }

    if (is_last && (!kSweep || swph == 1)) {  // rows whose final sample is NaN need a search for the last defined range (flags: bits 8.. of s_hasnan)
      const int gl = last_group / nth;  // the group of this thread that holds the last column
      unsigned nm = 0u;
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (g == gl) nm = nanmask[g];
      if (nm == 0u || !nanrange) {  // the usual tile: every final sample is defined
        s_seen_full = 1;
      } else {
        unsigned need = 0u;
#pragma unroll
        for (int t = 0; t < T; ++t)
          if (t < Ta && ((nm >> (4 * t + 3)) & 1u)) need |= 1u << t;
        if (need != ((1u << Ta) - 1u)) s_seen_full = 1;
        if (need) atomicOr(&s_hasnan[it], (int)(need << 8));
      }
    }

    float noise_lin = 0.f;
    if (kSweep) {
      const bool lastk = swk == S - 1;
      if (swph == 0) {
        // ---- sweep 1: only the column sums of u are kept; the last sub-tile runs the range-tile reduction ----------------
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int k = 0; k < 4; ++k) se_acc[kSweep ? g : 0][k] += se[g][k];
        rows_acc += Ta;
        if (lastk) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int b = bsplit[g];
            const float(&cs)[4] = se_acc[kSweep ? g : 0];
            const float A = cs[0] + ((b > 1) ? cs[1] : 0.f) + (((b > 2) ? cs[2] : 0.f) + ((b > 3) ? cs[3] : 0.f));
            const float B = ((b > 1) ? 0.f : cs[1]) + (((b > 2) ? 0.f : cs[2]) + ((b > 3) ? 0.f : cs[3]));
            if (liveg[g]) s_ga[colg[g] >> 2] = A, s_gb[colg[g] >> 2] = B;
          }
        }
        __syncthreads();  // the sub-tile's slot is free; (last sub-tile) column sums visible
        if (prod_warp && lane == 0) {
          fence_proxy_async();
          issue_tiles(qi);
          if (lastk) s_hasnan[it ^ 1] = 0;
        }
        if (lastk) {
          const int L = pr.rt_lanes, tpw = pr.rt_tpw;
          const int slot = (lane * pr.rt_recip) >> 16;  // lane / L
          const int q = lane - slot * L;
          const bool hasnan = (s_hasnan[it] & 1) != 0;
          unsigned m = kInfBits;
          for (int rb = (tid >> 5) * tpw; rb < nRt; rb += (nth >> 5) * tpw) {  // warp-uniform
            const int rt = rb + slot;
            const bool in = slot < tpw && rt < nRt;
            const int j0 = in ? rt * pr.rs_num : 0;
            const int j1 = in ? ((j0 + pr.rs_num < R) ? j0 + pr.rs_num : R) : 0;
            const int ga = (j0 + 3) >> 2, gb = (j1 + 3) >> 2;
            float sm = (q == 0 && ga > 0 && in) ? s_gb[ga - 1] : 0.f;
            for (int gq = ga + q; gq < gb; gq += L) sm += s_ga[gq];
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
              const float o = __shfl_down_sync(0xffffffffu, sm, d);
              sm += (q + d < L) ? o : 0.f;
            }
            if (in && q == 0) {
              int def = 0;
              if (hasnan) {  // CTA-uniform
                def = s_def[it * nRt + rt];
                s_def[it * nRt + rt] = 0;
              }
              const int n = s_valid[rt] * rows_acc - def;
              if (n > 0) {
                const unsigned u = __float_as_uint(sm * rcp_approx((float)n));  // >= 0: uint order == float order
                m = (u < m) ? u : m;
              }
            }
          }
          m = __reduce_min_sync(0xffffffffu, m);
          if (lane == 0) s_wmin[tid >> 5] = m;
          __syncthreads();
          const unsigned u = __reduce_min_sync(0xffffffffu, lane < kMaxWarps ? s_wmin[lane] : kInfBits);
          float v = (u == kInfBits) ? CUDART_NAN_F : __uint_as_float(u);
          if (pr.noise_max_lin == pr.noise_max_lin) v = (v < pr.noise_max_lin) ? v : pr.noise_max_lin;  // NaN -> max
          noise_keep = v;
          if (tid == 0 && pr.noise_out) pr.noise_out[g0 + li] = kLog2ToDb * fast_log2(v);
          __syncthreads();  // s_wmin / s_ga are rewritten by the next noise tile; s_fs (flush) aliases s_ga
#pragma unroll
          for (int g = 0; g < (kSweep ? G : 1); ++g)
#pragma unroll
            for (int k = 0; k < 4; ++k) se_acc[g][k] = 0.f;
          rows_acc = 0;
        }
        if (++swk == S) swk = 0, swph = 1;
        continue;  // no binning in sweep 1
      }
      // ---- sweep 2: the noise is known; free the slot, then remove the noise and bin --------------------------------------
      __syncthreads();
      if (prod_warp && lane == 0) {
        fence_proxy_async();
        issue_tiles(qi);
      }
      const unsigned need_last = (unsigned)s_hasnan[it] >> 8;  // CTA-uniform, almost always 0
      if (need_last) {
        offer_last(need_last, e);
        __syncthreads();
        if (tid == 0) {
          settle_last(need_last, ti->row0);
          s_hasnan[it] &= 0xff;  // the flags belong to this sub-tile only
        }
        __syncthreads();
      }
      noise_lin = noise_keep;
      if (++swk == S) swk = 0, swph = 0, ++swli;
    } else if (kNoise) {
      // ---- phase 1: column sums of u -> column-group sums -> range-tile means -> min ---------------------------------
      // a group of four columns touches at most two range tiles (range_sample_num >= 4): columns k < bsplit[g] belong
      // to the tile of the first column (sum A), the rest to the next tile (sum B)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int b = bsplit[g];
        const float A = se[g][0] + ((b > 1) ? se[g][1] : 0.f) + (((b > 2) ? se[g][2] : 0.f) + ((b > 3) ? se[g][3] : 0.f));
        const float B = ((b > 1) ? 0.f : se[g][1]) + (((b > 2) ? 0.f : se[g][2]) + ((b > 3) ? 0.f : se[g][3]));
        if (liveg[g]) s_ga[colg[g] >> 2] = A, s_gb[colg[g] >> 2] = B;
      }
      __syncthreads();  // (A) the tile's slot is free; column sums visible
      if (prod_warp && lane == 0) {  // refill at once: a later issue (after (B)) measured slower, the lead time matters
        fence_proxy_async();
        issue_tiles(li);
        s_hasnan[it ^ 1] = 0;
      }
      const unsigned need_last = (unsigned)s_hasnan[it] >> 8;  // CTA-uniform, almost always 0
      if (need_last) offer_last(need_last, e);
      // L lanes per range tile (L = pr.rt_lanes, 32 / L tiles per warp so that one pass over the warps covers the
      // row): the A sums of the groups whose first column lies in the tile, interleaved over the lanes, plus the B sum
      // of the group that straddles into it; per-warp minima are combined by every thread after barrier (B)
      {
        const int L = pr.rt_lanes, tpw = pr.rt_tpw;
        const int slot = (lane * pr.rt_recip) >> 16;  // lane / L
        const int q = lane - slot * L;
        const bool hasnan = (s_hasnan[it] & 1) != 0;
        unsigned m = kInfBits;
        for (int rb = (tid >> 5) * tpw; rb < nRt; rb += (nth >> 5) * tpw) {  // warp-uniform
          const int rt = rb + slot;
          const bool in = slot < tpw && rt < nRt;
          const int j0 = in ? rt * pr.rs_num : 0;
          const int j1 = in ? ((j0 + pr.rs_num < R) ? j0 + pr.rs_num : R) : 0;
          const int ga = (j0 + 3) >> 2, gb = (j1 + 3) >> 2;
          float s = (q == 0 && ga > 0 && in) ? s_gb[ga - 1] : 0.f;
          for (int gq = ga + q; gq < gb; gq += 3 * L) {  // three loads in flight
            const float a0 = s_ga[gq];
            const float a1 = (gq + L < gb) ? s_ga[gq + L] : 0.f;
            const float a2 = (gq + 2 * L < gb) ? s_ga[gq + 2 * L] : 0.f;
            s += (a0 + a1) + a2;
          }
#pragma unroll
          for (int d = 1; d < 8; d <<= 1) {
            const float o = __shfl_down_sync(0xffffffffu, s, d);
            s += (q + d < L) ? o : 0.f;
          }
          if (in && q == 0) {
            int def = 0;
            if (hasnan) {  // CTA-uniform
              def = s_def[it * nRt + rt];
              s_def[it * nRt + rt] = 0;
            }
            const int n = s_valid[rt] * Ta - def;
            if (n > 0) {
              const unsigned u = __float_as_uint(s * rcp_approx((float)n));  // >= 0: uint order == float order
              m = (u < m) ? u : m;
            }
          }
        }
        m = __reduce_min_sync(0xffffffffu, m);
        if (lane == 0) s_wmin[tid >> 5] = m;
      }
      __syncthreads();  // (B)
      {
        // minimum of the per-warp minima: one entry per lane and a warp reduction (entries of absent warps stay +inf)
        static_assert(kMaxWarps <= 32, "one s_wmin entry per lane");
        const unsigned u = __reduce_min_sync(0xffffffffu, lane < kMaxWarps ? s_wmin[lane] : kInfBits);
        float v = (u == kInfBits) ? CUDART_NAN_F : __uint_as_float(u);
        if (pr.noise_max_lin == pr.noise_max_lin) v = (v < pr.noise_max_lin) ? v : pr.noise_max_lin;  // NaN -> max
        noise_lin = v;
        if (tid == 0 && pr.noise_out) pr.noise_out[g0 + li] = kLog2ToDb * fast_log2(v);  // global tile = c * nPt + tile
      }
      if (need_last && tid == 0) settle_last(need_last, ti->row0);
    } else {
      __syncthreads();  // the tile's slot is free
      if (prod_warp && lane == 0) {
        fence_proxy_async();
        issue_tiles(li);
        s_hasnan[it ^ 1] = 0;
      }
      const unsigned need_last = (unsigned)s_hasnan[it] >> 8;  // CTA-uniform, almost always 0
      if (need_last) {
        offer_last(need_last, e);
        __syncthreads();
        if (tid == 0) settle_last(need_last, ti->row0);
      }
    }

    // ---- phase 2: noise removal + accumulation into the register cells ------------------------------------------------
    // survivors: u > thr (one tile-wide threshold);  sum(TL u - TL noise) = TL (sum(u) - n noise)
    const float thr = kNoise ? noise_lin * pr.snr1 : -1.f;    // Sv_c - Sv_noise > SNR  <=>  u > noise (1 + 10^(SNR/10))
    const float nz = (noise_lin == noise_lin) ? noise_lin : 0.f;  // NaN noise: nothing survives, keep the sums clean
    if (kKeep) {
      // Full-size outputs straight from the register tile (u domain): Sv = dB(u TL), Sv_noise = dB(noise TL),
      // Sv_corrected = dB((u - noise) TL) where it survives; a NaN sample (u = -2 sentinel) gives NaN, and with it a
      // NaN echo_range / Sv_noise when the law ties the range to the sample (range.py:143-148).  Streaming stores.
      const float lnz = fast_log2(noise_lin);
      const long long row0 = ti->row0;
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (liveg[g]) {
          const float4 t4 = *reinterpret_cast<const float4*>(s_tl + colg[g]);
          const float4 l4 = *reinterpret_cast<const float4*>(s_ltl + colg[g]);
          const float4 r4 = *reinterpret_cast<const float4*>(s_rr + colg[g]);
          const float tl[4] = {t4.x, t4.y, t4.z, t4.w}, ltl[4] = {l4.x, l4.y, l4.z, l4.w}, rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
          for (int t = 0; t < T; ++t)
            if (t < Ta) {
              const size_t o = (size_t)(row0 + t) * R + colg[g];
              float osv[4], orr[4], osn[4], osc[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float u = e[g][t][k];
                const bool member = !(nanrange && u < 0.f);
                orr[k] = member ? rr[k] : CUDART_NAN_F;
                if (kNoise) osn[k] = member ? kLog2ToDb * (lnz + ltl[k]) : CUDART_NAN_F;
              }
              if (pr.o_sv) {
#pragma unroll
                for (int k = 0; k < 4; ++k) osv[k] = (tl[k] != 0.f) ? kLog2ToDb * (fast_log2(e[g][t][k]) + ltl[k]) : CUDART_NAN_F;
                st_stream4(reinterpret_cast<float4*>(pr.o_sv + o), make_float4(osv[0], osv[1], osv[2], osv[3]));
              }
              if (pr.o_rng) st_stream4(reinterpret_cast<float4*>(pr.o_rng + o), make_float4(orr[0], orr[1], orr[2], orr[3]));
              if (kNoise && pr.o_svn) st_stream4(reinterpret_cast<float4*>(pr.o_svn + o), make_float4(osn[0], osn[1], osn[2], osn[3]));
              if (kNoise && pr.o_svc) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float u = e[g][t][k];
                  osc[k] = (u > thr) ? kLog2ToDb * (fast_log2(u - noise_lin) + ltl[k]) : CUDART_NAN_F;
                }
                st_stream4(reinterpret_cast<float4*>(pr.o_svc + o), make_float4(osc[0], osc[1], osc[2], osc[3]));
              }
            }
        }
    }
    const int nruns = ti->nruns;
    // The usual tile: all T rows fall into ONE ping bin (one run, no row predicates, no run loop).  Straight-line code:
    // per column pair and row FSET x2 + FFMA2 + FADD2.
    if (nruns == 1 && Ta == T && ti->run_cell[0] >= 0) {
      const int cell = ti->run_cell[0];
      if (cell != cur_cell || acc.rows + T > kFlushRows) {
        if (cur_cell >= 0) flush();
        cur_cell = cell;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 t4 = *reinterpret_cast<const float4*>(s_tl + colg[g]);
        const float tl[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
          float2 sg = make_float2(0.f, 0.f), ng = make_float2(0.f, 0.f);
#pragma unroll
          for (int t = 0; t < T; ++t) {
            float2 m;
            m.x = (e[g][t][k] > thr) ? 1.f : 0.f;
            m.y = (e[g][t][k + 1] > thr) ? 1.f : 0.f;
            sg = ffma2(m, make_float2(e[g][t][k], e[g][t][k + 1]), sg);
            ng = fadd2(ng, m);
          }
          if (!kNoise) {  // thr = -1 lets the u = 0 of an undefined column through: not a survivor
            ng.x = (tl[k] != 0.f) ? ng.x : 0.f;
            ng.y = (tl[k + 1] != 0.f) ? ng.y : 0.f;
          }
          acc.s[g][k] = fmaf(tl[k], fmaf(-ng.x, nz, sg.x), acc.s[g][k]);
          acc.s[g][k + 1] = fmaf(tl[k + 1], fmaf(-ng.y, nz, sg.y), acc.s[g][k + 1]);
          acc.good[g][k] += ng.x;
          acc.good[g][k + 1] += ng.y;
        }
        if (nanmask[g] != 0u && nanrange) {  // echo_range is NaN where the sample is NaN (range.py:143-148): not a member
          const unsigned rows_mask = (T >= 8) ? 0xffffffffu : ((1u << (4 * T)) - 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) acc.nanm[g] += (unsigned)__popc(nanmask[g] & rows_mask & (0x11111111u << k)) << (8 * k);
        }
      }
      acc.rows += T;
      continue;
    }
    int ta = 0;
    for (int r = 0; r < nruns; ++r) {
      const int tb = ti->run_end[r];
      const int cell = ti->run_cell[r];
      if (cell != cur_cell || acc.rows + (tb - ta) > kFlushRows) {
        if (cur_cell >= 0) flush();
        cur_cell = cell;
      }
      if (cell >= 0) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float4 t4 = *reinterpret_cast<const float4*>(s_tl + colg[g]);
          const float tl[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int k = 0; k < 4; k += 2) {  // column pairs: mask = (u > thr) as 1.0 / 0.0, FSET x2 + FFMA2 + FADD2 per row
            float2 sg = make_float2(0.f, 0.f), ng = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < T; ++t) {
              float2 m;
              m.x = (t >= ta && t < tb && e[g][t][k] > thr) ? 1.f : 0.f;
              m.y = (t >= ta && t < tb && e[g][t][k + 1] > thr) ? 1.f : 0.f;
              sg = ffma2(m, make_float2(e[g][t][k], e[g][t][k + 1]), sg);
              ng = fadd2(ng, m);
            }
            if (!kNoise) {  // thr = -1 lets the u = 0 of an undefined column through: not a survivor
              ng.x = (tl[k] != 0.f) ? ng.x : 0.f;
              ng.y = (tl[k + 1] != 0.f) ? ng.y : 0.f;
            }
            // sum over the survivors of 10^(Sv_corrected/10) = TL (u - noise)
            acc.s[g][k] = fmaf(tl[k], fmaf(-ng.x, nz, sg.x), acc.s[g][k]);
            acc.s[g][k + 1] = fmaf(tl[k + 1], fmaf(-ng.y, nz, sg.y), acc.s[g][k + 1]);
            acc.good[g][k] += ng.x;
            acc.good[g][k + 1] += ng.y;
          }
          if (nanmask[g] != 0u && nanrange) {  // echo_range is NaN where the sample is NaN (range.py:143-148): not a member
            const unsigned rows_mask = ((tb >= 8) ? 0xffffffffu : ((1u << (4 * tb)) - 1u)) & ~((1u << (4 * ta)) - 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc.nanm[g] += (unsigned)__popc(nanmask[g] & rows_mask & (0x11111111u << k)) << (8 * k);
          }
        }
        acc.rows += tb - ta;
      }
      ta = tb;
    }
  }
  if (cur_cell >= 0) flush();
  if (is_last) {
    if (s_seen_full && s_range_last == s_range_last) s_rmax_local = fmax(s_rmax_local, s_range_last);
    if (s_rmax_local != -CUDART_INF) atomic_max_d(pr.rmax, s_rmax_local);
  }
}

size_t fast_smem(long long R, int T, int nR, int ntiles_ring, int nRt, int xbytes, bool keep = false) {
  return (((size_t)R * 12 + (size_t)R / 4 + 15) & ~(size_t)15) + (((size_t)ntiles_ring * T * R * xbytes + 15) & ~(size_t)15) +
         (size_t)(nR + 1) * 12 + (size_t)nRt * 12 + 16 + (keep ? (size_t)R * 8 + 16 : 0);
}

constexpr size_t kSmemMax = 227 * 1024 - 2048;  // static shared memory of the kernel comes on top

template <int T, int G, bool kNoise, bool kI16, bool kSweep = false, bool kKeep = false>
int launch_fast(const FastParams& pr, int threads, size_t smem, cudaStream_t s) {
  auto kern = pipeline_fast_kernel<T, G, kNoise, kI16, kSweep, kKeep>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) return -1;
  // (epb_set_grid_reserve: SMs left to concurrent kernels of other streams, e.g. the NCCL exchange of the previous step)
  long long grid = (long long)(epb_num_sms() > epb_grid_reserve() ? epb_num_sms() - epb_grid_reserve() : 1) * per_sm;
  if (grid > pr.ntiles) grid = pr.ntiles;
  kern<<<(unsigned)grid, threads, smem, s>>>(pr);
  return 0;
}

}  // namespace

// The instantiations are spread over four translation units (pipeline_fast_{f32a,f32b,i16a,i16b}.cu) so that they
// compile in parallel; each defines one launcher with this macro.  The parameter block crosses the unit boundary as
// const void* (FastParams lives in the unnamed namespace of every unit; the layout is this header's).
// Launchers return 0 when the kernel was launched, -1 when it could not be, -2 when (T, noise) is not theirs.
#define EPB_FAST_LAUNCH_CASE(TT, NZ, I16) \
  ((G != 1) ? launch_fast<TT, EPB_GBIG, NZ, I16>(pr, threads, smem, s) : launch_fast<TT, 1, NZ, I16>(pr, threads, smem, s))
#define EPB_DEFINE_FAST_LAUNCHER(NAME, I16, TA, TB, TC, TD, WITH_NONOISE)                                            \
  int NAME(const void* prv, int T, int G, int noise, int threads, size_t smem, cudaStream_t s) {                      \
    const FastParams& pr = *static_cast<const FastParams*>(prv);                                                       \
    if (!noise) return WITH_NONOISE ? EPB_FAST_LAUNCH_CASE(4, false, I16) : -2;                                        \
    switch (T) {                                                                                                      \
      case TA: return EPB_FAST_LAUNCH_CASE(TA, true, I16);                                                            \
      case TB: return EPB_FAST_LAUNCH_CASE(TB, true, I16);                                                            \
      case TC: return EPB_FAST_LAUNCH_CASE(TC, true, I16);                                                            \
      case TD: return EPB_FAST_LAUNCH_CASE(TD, true, I16);                                                            \
    }                                                                                                                 \
    return -2;                                                                                                        \
  }

// kSweep instantiations (ping_num > 8: sub-tiles of TA..TD rows, two sweeps)
#define EPB_FAST_SWEEP_CASE(TT, I16) \
  ((G != 1) ? launch_fast<TT, EPB_GBIG, true, I16, true>(pr, threads, smem, s) : launch_fast<TT, 1, true, I16, true>(pr, threads, smem, s))
#define EPB_DEFINE_FAST_SWEEP_LAUNCHER(NAME, I16, TA, TB, TC, TD)                                                     \
  int NAME(const void* prv, int T, int G, int noise, int threads, size_t smem, cudaStream_t s) {                      \
    const FastParams& pr = *static_cast<const FastParams*>(prv);                                                       \
    if (!noise) return -2;                                                                                            \
    switch (T) {                                                                                                      \
      case TA: return EPB_FAST_SWEEP_CASE(TA, I16);                                                                   \
      case TB: return EPB_FAST_SWEEP_CASE(TB, I16);                                                                   \
      case TC: return EPB_FAST_SWEEP_CASE(TC, I16);                                                                   \
      case TD: return EPB_FAST_SWEEP_CASE(TD, I16);                                                                   \
    }                                                                                                                 \
    return -2;                                                                                                        \
  }

// kKeep instantiations (float32 input, full-size outputs streamed out of the register tile)
#define EPB_FAST_KEEP_CASE(TT, NZ, SW) \
  ((G != 1) ? launch_fast<TT, EPB_GBIG, NZ, false, SW, true>(pr, threads, smem, s) : launch_fast<TT, 1, NZ, false, SW, true>(pr, threads, smem, s))
#define EPB_DEFINE_FAST_KEEP_LAUNCHER(NAME, SW, TA, TB, TC, TD, WITH_NONOISE)                                         \
  int NAME(const void* prv, int T, int G, int noise, int threads, size_t smem, cudaStream_t s) {                      \
    const FastParams& pr = *static_cast<const FastParams*>(prv);                                                       \
    if (!noise) return WITH_NONOISE ? EPB_FAST_KEEP_CASE(4, false, false) : -2;                                       \
    switch (T) {                                                                                                      \
      case TA: return EPB_FAST_KEEP_CASE(TA, true, SW);                                                               \
      case TB: return EPB_FAST_KEEP_CASE(TB, true, SW);                                                               \
      case TC: return EPB_FAST_KEEP_CASE(TC, true, SW);                                                               \
      case TD: return EPB_FAST_KEEP_CASE(TD, true, SW);                                                               \
    }                                                                                                                 \
    return -2;                                                                                                        \
  }
