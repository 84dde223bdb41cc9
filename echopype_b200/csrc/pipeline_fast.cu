// Fused pipeline, fast path: raw power -> Sv -> background-noise removal -> MVBS accumulators, one pass over HBM
// at 4 algorithmic bytes per sample (same semantics as pipeline.cu).
//
// PERSISTENT kernel, one CTA per SM (grid = SM count x resident CTAs), each CTA owns a contiguous run of
// (channel, ping-tile) tiles.  Design points (DESIGN.md "fused pipeline, fast path"):
//   * a ring of row slots in shared memory is filled by TMA bulk copies (cp.async.bulk, one per ping row, the rows
//     of a tile completing on the tile's mbarrier; SASS UBLKCP / SYNCS) issued tiles ahead of the consumer, so HBM
//     latency and the per-tile reductions of the consumer overlap;
//   * one thread owns four adjacent range samples (one LDS.128 per row).  e = 10^((front(x)-K)/10) of the
//     whole tile (ping_num <= 8 rows) lives in REGISTERS between the noise estimate (phase 1) and the noise
//     removal / binning (phase 2): one FFMA + one MUFU.EX2 per sample, shared memory is read once per sample;
//   * all range-only terms (h = R'^2 10^(2aR'/10) and TL/h) and the range-bin index of every column are computed
//     once per range law (normally once per channel) and parked in shared memory, the exact float64 bin
//     boundaries likewise;
//   * noise removal in the e domain: the sample survives iff e > noise (1 + 10^(SNR/10)) TL / h (one compare), the
//     surviving e are summed per column and scaled once per tile: sum(e h - noise TL) = h sum(e) - n noise TL;
//   * (sum, count) of a thread's four columns accumulate in registers ACROSS tiles while the ping bin does not
//     change; on a bin change a segmented warp-shuffle reduction over runs of equal range-bin index issues one
//     float64 atomic triple per (warp, range bin).
// Handles the regular case: every tile's rows share one range law and have finite calibration constants (checked
// on the device by classify_kernel; otherwise the general kernel of pipeline.cu runs instead), R % 4 == 0,
// R <= 4096, ping_num <= 8, no full-size outputs.
#include "pipeline_common.cuh"

namespace {
using namespace epb;

constexpr int kMaxT = 8;
constexpr int kMaxTilesInFlight = 8;
constexpr int kFlushRows = 248;  // packed 8-bit per-column counters: flush a cell before a counter can overflow
constexpr unsigned kInfBits = 0x7f800000u;

struct FastParams {
  const float* x;
  const epb_row* rows;
  const int* xbin;
  const double* edges;
  double* acc;
  float* noise_out;
  const int* irregular;  // workspace flag from classify_kernel: != 0 -> this kernel does nothing
  long long C, P, nX, ntiles;
  int R, nR, rs_num, closed_right, nslots, nPt;
  float noise_max_lin;  // NaN: no cap
  float snr1;           // 1 + 10^(SNR/10)
};

struct TileInfo {  // written by warp 0 one tile ahead
  float2 rc[kMaxT];     // per-row (c0, c1): e = 2^(x c1 + c0)
  int run_end[kMaxT];   // rows [run_end[r-1], run_end[r]) share one accumulator cell
  int run_cell[kMaxT];  // c * nX + ping bin, or -1 (ping outside every bin)
  int nruns;
  int Ta;      // rows present in the tile
  int lawchg;  // the tile's range law differs from the previous tile's
  long long row0;  // first (channel, ping) row of the tile
};

// bitwise comparison of the range law of two rows (exact float64 law + value-form splits)
__device__ __forceinline__ bool same_law(const epb_row& a, const epb_row& b) {
  return a.p0 == b.p0 && a.p1 == b.p1 && a.p2 == b.p2 && a.p3 == b.p3 && a.p4 == b.p4 && a.off1 == b.off1 &&
         a.off2 == b.off2 && a.r0 == b.r0 && a.a == b.a && a.two_alpha == b.two_alpha && a.n_start == b.n_start &&
         a.law == b.law && a.azfp_N == b.azfp_N && a.a_h == b.a_h && a.a_l == b.a_l && a.r0_h == b.r0_h &&
         a.r0_l == b.r0_l && a.bp_h == b.bp_h && a.bp_l == b.bp_l && a.c2 == b.c2;
}

// one thread per tile: flag tiles whose rows do not share one law or carry NaN constants
__global__ void classify_kernel(const epb_row* __restrict__ rows, long long P, int T, int nPt, long long ntiles,
                                int* __restrict__ irregular) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= ntiles) return;
  const long long c = g / nPt;
  const long long p0 = (g % nPt) * (long long)T;
  const int Ta = (int)((p0 + T <= P) ? T : (P - p0));
  const epb_row* r0 = rows + c * P + p0;
  bool bad = false;
  for (int t = 0; t < Ta; ++t) {
    const epb_row& r = r0[t];
    if (!(r.c0 == r.c0 && r.c1 == r.c1)) bad = true;
    if (!same_law(r0[0], r)) bad = true;  // NaN laws never compare equal
  }
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

// reduce one group of per-thread accumulators of a warp (128 adjacent columns) over runs of equal range-bin keys
// and add each run to the float64 accumulator grid; every lane of the warp must call this.  keys: the thread's
// four range-bin indices (int16, -1 = outside every bin).
__device__ __noinline__ void flush_group(float4 s4, float4 g4, unsigned nanp, int rows, const short* __restrict__ keys,
                                         bool live, double* __restrict__ acc_row) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  short4 k4 = make_short4(-1, -1, -1, -1);
  if (live) k4 = *reinterpret_cast<const short4*>(keys);
  const int key[4] = {k4.x, k4.y, k4.z, k4.w};
  const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
  const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
  int good[4], bad[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    good[k] = (int)gv[k];
    bad[k] = (rows - (int)((nanp >> (8 * k)) & 0xffu)) - good[k];
  }
  const bool same = (key[0] == key[1]) && (key[1] == key[2]) && (key[2] == key[3]);
  int kk = -1;
  float ms = 0.f;
  int mc = 0;  // good | bad << 16 (a warp run holds at most 32 x 4 x 248 < 2^15 of each)
  if (same) {
    kk = key[0];
    ms = (sv[0] + sv[1]) + (sv[2] + sv[3]);
    mc = (good[0] + good[1] + good[2] + good[3]) | ((bad[0] + bad[1] + bad[2] + bad[3]) << 16);
  } else {  // a bin boundary inside the thread's four columns
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (key[k] >= 0) atomic_cell(acc_row + 4 * (long long)key[k], sv[k], good[k], bad[k]);
  }
  const int prev = __shfl_up_sync(full, kk, 1);
  const unsigned heads = __ballot_sync(full, (lane == 0) || (prev != kk));
  if (!(heads == 1u && kk < 0)) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float s2_ = __shfl_up_sync(full, ms, d);
      const int c2 = __shfl_up_sync(full, mc, d);
      // lane-d belongs to the same run iff no run head lies in (lane-d, lane]
      const unsigned between = (lane >= d) ? ((heads >> (lane - d + 1)) & ((1u << d) - 1u)) : 1u;
      if (between == 0u) {
        ms += s2_;
        mc += c2;
      }
    }
    const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
    if (tail && kk >= 0) atomic_cell(acc_row + 4 * (long long)kk, ms, mc & 0xffff, (int)((unsigned)mc >> 16));
  }
}

struct Producer {  // TMA issue cursor, used by one thread only (kept in shared memory, not in registers)
  int tile;        // next local tile to issue
  int slot;        // ring slot of its first row
  int rows;        // rows issued so far
  int bar;         // tile % NB
  int c, it;       // channel / ping tile of `tile`
};

// words of an epb_row (as 48 x 32 bit) that make up the range law: p0..p4, off1, off2, r0, a, two_alpha (0-19),
// n_start, law, azfp_N (28-30), a_h..bp_l, two_alpha_f, slog2 (32-39), c2, spow (44-45)
__device__ __forceinline__ bool law_word(int w) {
  return (w < 20) || (w >= 28 && w <= 30) || (w >= 32 && w <= 39) || w == 44 || w == 45;
}

// T rows per tile (ping_num), G column groups of four per thread (threads = R / (4 G))
template <int T, int G, bool kNoise>
__global__ void __launch_bounds__(512, 1) pipeline_fast_kernel(const FastParams pr) {
  if (*pr.irregular) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_full[kMaxTilesInFlight];  // one mbarrier per tile in flight
  __shared__ TileInfo s_tile[3];  // tile li-1 may still be read while li+2 is written: see describe_store
  __shared__ unsigned int s_min[2];
  __shared__ int s_hasnan[2];   // a thread saw a NaN sample in the tile: range-tile counts come from s_colcnt
  __shared__ unsigned s_lawwords[48];  // first row of the current range-law segment (producer warp only)
  __shared__ Producer s_prod;
  const int R = pr.R, nR = pr.nR, N = pr.nslots;
  const int tid = threadIdx.x;
  const int nth = blockDim.x;
  const int lane = tid & 31;
  const bool prod_warp = (tid >> 5) == (nth >> 5) - 1;  // the last warp doubles as descriptor / TMA producer
  const int nRt = kNoise ? (R + pr.rs_num - 1) / pr.rs_num : 0;
  // ---- dynamic shared memory ----------------------------------------------------------------------------------------
  // [h R][ginv R][colsum R][keys R int16][colcnt R bytes][pad to 16][ring N x R][edges nR+1 f64][bounds nR+1][valid nRt]
  float* const s_h = reinterpret_cast<float*>(smem_raw);  // 10^(Sv/10) / e
  float* const s_ginv = s_h + R;                          // 10^(TL/10) / h
  float* const s_colsum = s_ginv + R;
  short* const s_keys = reinterpret_cast<short*>(s_colsum + R);
  unsigned char* const s_colcnt = reinterpret_cast<unsigned char*>(s_keys + R);
  float* const s_ring = reinterpret_cast<float*>(smem_raw + (((size_t)R * 15 + 15) & ~(size_t)15));
  double* const s_edges = reinterpret_cast<double*>(s_ring + (size_t)N * R);
  int* const s_bounds = reinterpret_cast<int*>(s_edges + (nR + 1));
  int* const s_valid = s_bounds + (nR + 1);  // columns of each range tile with a defined Sv (n >= n_start, R' >= 0)

  // ---- tile range of this CTA -----------------------------------------------------------------------------------
  const long long g0 = pr.ntiles * (long long)blockIdx.x / gridDim.x;
  const int ntl = (int)(pr.ntiles * (long long)(blockIdx.x + 1) / gridDim.x - g0);  // local tiles 0..ntl-1
  if (ntl <= 0) return;
  const int nPt = pr.nPt;
  const int NB = N / T + 1;  // tile barriers in rotation (>= tiles in flight)
  const uint32_t row_bytes = (uint32_t)R * 4u;
  // one thread: issue whole tiles while their rows fit in the ring given `consumed` rows are free again
  // and the tile that last used the tile's mbarrier (tile - NB) has been consumed (last_done = last consumed tile)
  auto issue_tiles = [&](int consumed, int last_done) {
    Producer p = s_prod;
    while (p.tile < ntl && p.tile - NB <= last_done) {
      const long long p0 = (long long)p.it * T;
      const int Ta = (int)((p0 + T <= pr.P) ? T : (pr.P - p0));
      if (p.rows + Ta - consumed > N) break;
      const float* src = pr.x + ((long long)p.c * pr.P + p0) * (long long)R;
      unsigned long long* bar = &s_full[p.bar];
      if (++p.bar == NB) p.bar = 0;
      mbar_expect_tx(bar, row_bytes * (uint32_t)Ta);
      for (int t = 0; t < Ta; ++t) {
        bulk_g2s(s_ring + (size_t)p.slot * R, src + (size_t)t * R, row_bytes, bar);
        if (++p.slot == N) p.slot = 0;
      }
      p.rows += Ta;
      ++p.tile;
      if (++p.it == nPt) p.it = 0, ++p.c;
    }
    s_prod = p;
  };

  // Tile descriptor (producer warp): row constants, accumulator-cell runs, law change.  Split in two so that the
  // global loads are in flight while the warp does its share of the tile: describe_load early, describe_store late.
  struct DescRegs {
    float c0, c1;
    int xb;
    unsigned w0, w1;  // words lane and lane + 32 of the tile's first row record
    int c, itile;
  };
  auto describe_load = [&](int li) {  // local tile li (< ntl)
    DescRegs d;
    const long long g = g0 + li;
    d.c = (int)(g / nPt);
    d.itile = (int)(g - (long long)d.c * nPt);
    const long long p0 = (long long)d.itile * T;
    const int Ta = (int)((p0 + T <= pr.P) ? T : (pr.P - p0));
    const long long row0 = (long long)d.c * pr.P + p0;
    d.c0 = 0.f, d.c1 = 0.f, d.xb = -1;
    if (lane < Ta) {
      const epb_row* r = pr.rows + row0 + lane;
      d.c0 = __ldg(&r->c0);
      d.c1 = __ldg(&r->c1);
      d.xb = __ldg(pr.xbin + p0 + lane);
    }
    const unsigned* w = reinterpret_cast<const unsigned*>(pr.rows + row0);
    d.w0 = __ldg(w + lane);
    d.w1 = (lane < 16) ? __ldg(w + 32 + lane) : 0u;
    return d;
  };
  auto describe_store = [&](const DescRegs& d, TileInfo* ti, bool first) {
    const long long p0 = (long long)d.itile * T;
    const int Ta = (int)((p0 + T <= pr.P) ? T : (pr.P - p0));
    int xb = d.xb;
    if (xb < 0 || xb >= pr.nX) xb = -1;
    if (lane < Ta) ti->rc[lane] = make_float2(d.c0, d.c1);
    const int prev = __shfl_up_sync(0xffffffffu, xb, 1);
    const bool head = (lane < Ta) && (lane == 0 || prev != xb);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    if (head) {
      const int r = __popc(heads & ((1u << lane) - 1u));
      ti->run_cell[r] = (xb >= 0) ? (int)((long long)d.c * pr.nX + xb) : -1;
      const unsigned later = heads & ~((2u << lane) - 1u);
      ti->run_end[r] = later ? (__ffs(later) - 1) : Ta;
    }
    bool diff = law_word(lane) && (d.w0 != s_lawwords[lane]);
    if (lane < 16) diff = diff || (law_word(lane + 32) && d.w1 != s_lawwords[lane + 32]);
    const bool chg = first || __any_sync(0xffffffffu, diff);
    if (chg) {
      s_lawwords[lane] = d.w0;
      if (lane < 16) s_lawwords[lane + 32] = d.w1;
    }
    if (lane == 0) {
      ti->nruns = __popc(heads);
      ti->Ta = Ta;
      ti->row0 = (long long)d.c * pr.P + p0;
      ti->lawchg = chg;
    }
    __syncwarp();
  };

  // ---- prologue -----------------------------------------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kMaxTilesInFlight; ++i) mbar_init(&s_full[i], 1);
    mbar_init_fence();
    s_min[0] = kInfBits, s_min[1] = kInfBits;
    s_hasnan[0] = 0, s_hasnan[1] = 0;
    const int c0 = (int)(g0 / nPt), it0 = (int)(g0 - (long long)c0 * nPt);
    s_prod.tile = 0, s_prod.slot = 0, s_prod.rows = 0, s_prod.bar = 0, s_prod.c = c0, s_prod.it = it0;
  }
  for (int k = tid; k <= nR; k += nth) s_edges[k] = pr.edges[k];
  __syncthreads();
  if (prod_warp) {
    if (lane == 0) issue_tiles(0, -1);
    describe_store(describe_load(0), &s_tile[0], true);
    if (ntl > 1) describe_store(describe_load(1), &s_tile[1], false);
  }
  __syncthreads();

  // group g of this thread owns columns n0 + g * 4 * nth .. +3
  const int n0 = 4 * tid;
  bool nanrange = false;
  Acc<G> acc;
  acc.clear();
  int cur_cell = -1;
  int consumed = 0;  // rows consumed before the current tile
  int slot0 = 0;     // ring slot of the current tile's first row
  int bar_i = 0;     // li % NB and the phase parity of that mbarrier
  unsigned bar_p = 0;
  int tsel = 0;      // li % 3

  auto flush = [&]() {
    double* acc_row = pr.acc + (long long)cur_cell * nR * 4;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int ng = n0 + g * 4 * nth;
      flush_group(make_float4(acc.s[g][0], acc.s[g][1], acc.s[g][2], acc.s[g][3]),
                  make_float4(acc.good[g][0], acc.good[g][1], acc.good[g][2], acc.good[g][3]), acc.nanm[g], acc.rows,
                  s_keys + ng, ng < R, acc_row);
    }
    acc.clear();
  };

  for (int li = 0; li < ntl; ++li) {
    const int it = li & 1;
    const TileInfo* ti = &s_tile[tsel];
    const int tprev = (tsel == 0) ? 2 : tsel - 1;  // slot of tile li+2 (= slot of tile li-1, which is finished)
    tsel = (tsel == 2) ? 0 : tsel + 1;
    const int Ta = ti->Ta;
    // producer warp: start the descriptor loads of tile li+2 now, store them after barrier (A)
    DescRegs dreg;
    const bool have_desc = prod_warp && (li + 2 < ntl);
    if (have_desc) dreg = describe_load(li + 2);

    // ---- new range-law segment: flush, recompute boundaries, column terms, keys ------------------------------------
    if (ti->lawchg) {
      if (cur_cell >= 0) flush();
      cur_cell = -1;
      __syncthreads();  // every warp has used the old keys
      const long long row0 = ti->row0;
      const epb_row& lr = pr.rows[row0];
      for (int k = tid; k <= nR; k += nth) s_bounds[k] = first_at_or_above(lr, R, s_edges[k], pr.closed_right);
      const RowF rf = load_rowf(pr.rows + row0);
      nanrange = rf.nanrange;
      for (int n = n0; n < R; n += 4 * nth) {
        float hh[4], gi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const ColC cc = col_consts(rf, n + k);
          hh[k] = cc.h;
          gi[k] = __fdividef(cc.tl, cc.h);  // NaN where Sv is undefined; inf where R' = 0
          if (!(cc.h == cc.h)) gi[k] = CUDART_NAN_F;
        }
        *reinterpret_cast<float4*>(s_h + n) = make_float4(hh[0], hh[1], hh[2], hh[3]);
        *reinterpret_cast<float4*>(s_ginv + n) = make_float4(gi[0], gi[1], gi[2], gi[3]);
      }
      __syncthreads();
      for (int n = n0; n < R; n += 4 * nth) {
        short kk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) kk[k] = (short)key_of(s_bounds, nR, n + k);
        *reinterpret_cast<short4*>(s_keys + n) = make_short4(kk[0], kk[1], kk[2], kk[3]);
      }
      for (int rt = tid; rt < nRt; rt += nth) {
        const int j0 = rt * pr.rs_num, j1 = (j0 + pr.rs_num < R) ? j0 + pr.rs_num : R;
        int nv = 0;
        for (int j = j0; j < j1; ++j) {
          const float gi = s_ginv[j];
          nv += (gi == gi);
        }
        s_valid[rt] = nv;
      }
      // no barrier needed here: s_keys is thread-private, s_valid is read after barrier (A)
    }

    // ---- wait for the tile, e -> registers ---------------------------------------------------------------------------
    mbar_wait(&s_full[bar_i], bar_p);
    if (++bar_i == NB) bar_i = 0, bar_p ^= 1u;
    float e[G][T][4];
    {
      // rows of the tile sit in consecutive ring slots (wrapping at N)
      const float* src = s_ring + (size_t)slot0 * R + n0;
      int slot = slot0;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float2 rc = ti->rc[t];  // rows beyond Ta: stale constants, stale slot data, zeroed below
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n0 + g * 4 * nth < R) v = *reinterpret_cast<const float4*>(src + g * 4 * nth);
          e[g][t][0] = fast_exp2(fmaf(v.x, rc.y, rc.x));
          e[g][t][1] = fast_exp2(fmaf(v.y, rc.y, rc.x));
          e[g][t][2] = fast_exp2(fmaf(v.z, rc.y, rc.x));
          e[g][t][3] = fast_exp2(fmaf(v.w, rc.y, rc.x));
        }
        src += R;
        if (++slot == N) slot = 0, src = s_ring + n0;
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
    }
    float se[G][4];
    float chk = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        se[g][k] = e[g][0][k];
#pragma unroll
        for (int t = 1; t < T; ++t) se[g][k] += e[g][t][k];
        chk += se[g][k];
      }
    // Rare: some sample of this thread is NaN (padded ping) or e overflowed.  Remember where (bit t*4+k per group),
    // make the column sums NaN-free and replace the sample by -2 (never survives a threshold >= -1).
    unsigned nanmask[G];
    int cn[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      nanmask[g] = 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) cn[g][k] = Ta;
    }
    if (!(chk * 0.f == 0.f)) {
      if (kNoise) s_hasnan[it] = 1;
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          se[g][k] = 0.f;
          cn[g][k] = 0;
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const float v = e[g][t][k];
            const bool ok = (v * 0.f == 0.f);
            if (t < Ta) {
              se[g][k] += ok ? v : 0.f;
              cn[g][k] += ok;
              if (!ok) nanmask[g] |= 1u << (4 * t + k);
            }
            e[g][t][k] = ok ? v : -2.f;
          }
        }
    }

    float noise_lin = 0.f;
    if (kNoise) {
      // ---- phase 1: per-column sums of 10^((Sv-TL)/10) -> range-tile means -> min ----------------------------------------
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int ng = n0 + g * 4 * nth;
        if (ng < R) {
          const float4 g4 = *reinterpret_cast<const float4*>(s_ginv + ng);
          const float gi[4] = {g4.x, g4.y, g4.z, g4.w};
          float cs[4];
          unsigned cc = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool ok = (gi[k] == gi[k]);                   // Sv defined at this column (ginv = inf: Sv = -inf, adds 0)
            cs[k] = ok ? __fdividef(se[g][k], gi[k]) : 0.f;      // sum of 10^((Sv-TL)/10)
            cc |= (ok ? (unsigned)cn[g][k] : 0u) << (8 * k);
          }
          *reinterpret_cast<float4*>(s_colsum + ng) = make_float4(cs[0], cs[1], cs[2], cs[3]);
          *reinterpret_cast<unsigned*>(s_colcnt + ng) = cc;
        }
      }
      __syncthreads();  // (A) the tile's slots are free; column sums visible
      if (prod_warp) {
        if (lane == 0) {
          fence_proxy_async();
          issue_tiles(consumed + Ta, li);
          s_min[it ^ 1] = kInfBits;
          s_hasnan[it ^ 1] = 0;
        }
        if (have_desc) describe_store(dreg, &s_tile[tprev], false);
      }
      // one thread per range tile: float4 loads over the tile's column groups, edge groups masked
      if ((tid & ~31) < nRt) {
        const bool hasnan = s_hasnan[it] != 0;
        unsigned m = kInfBits;
        for (int rt = tid; rt < ((nRt + 31) & ~31); rt += nth) {
          if (rt < nRt) {
            const int j0 = rt * pr.rs_num, j1 = (j0 + pr.rs_num < R) ? j0 + pr.rs_num : R;
            const int ga = j0 >> 2, gb = (j1 - 1) >> 2;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            for (int gq = ga + 1; gq < gb; ++gq) {  // interior groups
              const float4 v = *reinterpret_cast<const float4*>(s_colsum + 4 * gq);
              s0 += v.x, s1 += v.y, s2 += v.z, s3 += v.w;
            }
            {
              const float4 v = *reinterpret_cast<const float4*>(s_colsum + 4 * ga);
              const int c = 4 * ga;
              s0 += (c + 0 >= j0 && c + 0 < j1) ? v.x : 0.f;
              s1 += (c + 1 >= j0 && c + 1 < j1) ? v.y : 0.f;
              s2 += (c + 2 >= j0 && c + 2 < j1) ? v.z : 0.f;
              s3 += (c + 3 < j1) ? v.w : 0.f;
            }
            if (gb > ga) {
              const float4 v = *reinterpret_cast<const float4*>(s_colsum + 4 * gb);
              const int c = 4 * gb;
              s0 += v.x;
              s1 += (c + 1 < j1) ? v.y : 0.f;
              s2 += (c + 2 < j1) ? v.z : 0.f;
              s3 += (c + 3 < j1) ? v.w : 0.f;
            }
            int n = s_valid[rt] * Ta;
            if (hasnan) {
              n = 0;
              for (int j = j0; j < j1; ++j) n += s_colcnt[j];
            }
            if (n > 0) {
              const unsigned u = __float_as_uint(__fdividef((s0 + s1) + (s2 + s3), (float)n));  // >= 0: uint order == float order
              m = (u < m) ? u : m;
            }
          }
        }
        m = __reduce_min_sync(0xffffffffu, m);
        if (lane == 0 && m != kInfBits) atomicMin(&s_min[it], m);
      }
      __syncthreads();  // (B)
      {
        const unsigned u = s_min[it];
        float v = (u == kInfBits) ? CUDART_NAN_F : __uint_as_float(u);
        if (pr.noise_max_lin == pr.noise_max_lin) v = (v < pr.noise_max_lin) ? v : pr.noise_max_lin;  // NaN -> max
        noise_lin = v;
        if (tid == 0 && pr.noise_out) pr.noise_out[g0 + li] = kLog2ToDb * fast_log2(v);  // global tile = c * nPt + tile
      }
    } else {
      __syncthreads();  // the tile's slots are free
      if (prod_warp) {
        if (lane == 0) {
          fence_proxy_async();
          issue_tiles(consumed + Ta, li);
        }
        if (have_desc) describe_store(dreg, &s_tile[tprev], false);
      }
    }

    // ---- phase 2: noise removal + accumulation into the register cells ------------------------------------------------
    // survivors: e > ethr (the SNR test in the e domain);  sum(e h - nl) = h sum(e) - n nl
    const int nruns = ti->nruns;
    const bool simple = (nruns == 1) && (Ta == T);
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
          const int ng = n0 + g * 4 * nth;
          float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = h4;
          if (ng < R) {
            h4 = *reinterpret_cast<const float4*>(s_h + ng);
            g4 = *reinterpret_cast<const float4*>(s_ginv + ng);
          }
          const float h[4] = {h4.x, h4.y, h4.z, h4.w}, gi[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float ethr, nl;
            if (kNoise) {
              const float ne = noise_lin * gi[k];  // noise TL / h: the noise floor in the e domain
              ethr = ne * pr.snr1;                 // Sv_c - Sv_noise > SNR  <=>  e > ne (1 + 10^(SNR/10))
              nl = ne * h[k];                      // 10^(Sv_noise/10)
            } else {
              ethr = (h[k] == h[k]) ? -1.f : CUDART_NAN_F;  // every non-NaN e of a column with defined Sv
              nl = 0.f;
            }
            float sg = 0.f, ng_f = 0.f;
            if (simple) {  // every e finite (or the -2 sentinel): branch-free mask arithmetic, FSET + FFMA + FADD
#pragma unroll
              for (int t = 0; t < T; ++t) {
                const float m = (e[g][t][k] > ethr) ? 1.f : 0.f;
                sg = fmaf(m, e[g][t][k], sg);
                ng_f += m;
              }
            } else {
#pragma unroll
              for (int t = 0; t < T; ++t) {
                const float m = (t >= ta && t < tb && e[g][t][k] > ethr) ? 1.f : 0.f;
                sg = fmaf(m, e[g][t][k], sg);
                ng_f += m;
              }
            }
            const float contrib = fmaf(h[k], sg, -(ng_f * nl));
            acc.s[g][k] += (ng_f > 0.f) ? contrib : 0.f;
            acc.good[g][k] += ng_f;
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
    consumed += Ta;
    slot0 += Ta;
    if (slot0 >= N) slot0 -= N;
  }
  if (cur_cell >= 0) flush();
}

size_t fast_smem(long long R, int nR, int nslots, int nRt) {
  return (((size_t)R * 15 + 15) & ~(size_t)15) + (size_t)nslots * R * 4 + (size_t)(nR + 1) * 12 + (size_t)nRt * 4 + 16;
}

constexpr size_t kSmemMax = 227 * 1024 - 2048;  // static shared memory of the kernel comes on top

template <int T, int G, bool kNoise>
int launch_fast(const FastParams& pr, int threads, size_t smem, cudaStream_t s) {
  auto kern = pipeline_fast_kernel<T, G, kNoise>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) return -1;
  long long grid = (long long)epb_num_sms() * per_sm;
  if (grid > pr.ntiles) grid = pr.ntiles;
  kern<<<(unsigned)grid, threads, smem, s>>>(pr);
  return 0;
}

}  // namespace

// Tries to launch the fast path.  Returns 1 when launched (the general kernel must then be launched with the same
// `irregular` flag so that exactly one of the two does the work), 0 when the static conditions do not hold.
int epb_pipeline_fast_try(const float* x, const epb_row* rows, const int* xbin, const double* r_edges, int nR,
                          int closed_right, double* acc, float* noise_out, long long C, long long P, long long R,
                          long long nX, int ping_num, int range_sample_num, float noise_max_lin, float snr_lin,
                          int* irregular, cudaStream_t s) {
  const bool noise = ping_num > 0;
  const int T = noise ? ping_num : 4;
  if (T > kMaxT || R % 4 != 0 || R > 4096 || R < 128 || C * nX >= (1LL << 31) || nR > 32000) return 0;
  const int G = (R / 4 > 512) ? 2 : 1;  // column groups per thread
  const int threads = (int)(((R / 4 + G - 1) / G + 31) / 32 * 32);
  const int nRt = noise ? (int)((R + range_sample_num - 1) / range_sample_num) : 0;
  // ring: as many row slots as fit (at least one tile, at most kMaxTilesInFlight - 1 tiles)
  int nslots = 0;
  for (int n = (kMaxTilesInFlight - 1) * T; n >= T; --n)
    if (fast_smem(R, nR, n, nRt) <= kSmemMax) {
      nslots = n;
      break;
    }
  if (nslots == 0) return 0;
  const size_t smem = fast_smem(R, nR, nslots, nRt);
  FastParams pr;
  pr.x = x, pr.rows = rows, pr.xbin = xbin, pr.edges = r_edges, pr.acc = acc, pr.noise_out = noise_out;
  pr.irregular = irregular;
  pr.C = C, pr.P = P, pr.nX = nX;
  pr.nPt = (int)((P + T - 1) / T);
  pr.ntiles = C * (long long)pr.nPt;
  pr.R = (int)R, pr.nR = nR, pr.rs_num = range_sample_num, pr.closed_right = closed_right, pr.nslots = nslots;
  pr.noise_max_lin = noise_max_lin;
  pr.snr1 = 1.f + snr_lin;
  if (cudaMemsetAsync(irregular, 0, sizeof(int), s) != cudaSuccess) return 0;
  classify_kernel<<<(unsigned)((pr.ntiles + 255) / 256), 256, 0, s>>>(rows, P, T, pr.nPt, pr.ntiles, irregular);
  int rc = -1;
#define EPB_FAST(TT)                                                                                              \
  case TT:                                                                                                        \
    rc = (G == 2) ? launch_fast<TT, 2, true>(pr, threads, smem, s) : launch_fast<TT, 1, true>(pr, threads, smem, s); \
    break;
  if (!noise)
    rc = (G == 2) ? launch_fast<4, 2, false>(pr, threads, smem, s) : launch_fast<4, 1, false>(pr, threads, smem, s);
  else
    switch (T) {
      EPB_FAST(1)
      EPB_FAST(2)
      EPB_FAST(3)
      EPB_FAST(4)
      EPB_FAST(5)
      EPB_FAST(6)
      EPB_FAST(7)
      EPB_FAST(8)
    }
#undef EPB_FAST
  if (rc != 0) cudaMemsetAsync(irregular, 1, sizeof(int), s);  // could not launch: the general kernel does the work
  return 1;
}
