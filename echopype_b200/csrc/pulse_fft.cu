// K3, FFT form: EK80 broadband pulse compression (matched filter) by overlap-save FFT in shared memory, fused with the
// Sv / TS epilogue.  Same contract as pulse.cu (the direct form, kept for replicas longer than 2049 taps).
//
// Reference: compress_pulse + _convolve_per_channel (calibrate/ek80_complex.py:285-369): per (ping, beam) and channel
//   y[n] = sum_{k<M} x[n+k] conj(tx[k])   (scipy.signal.convolve with the flipped conjugate replica, "full"[M-1:])
// NaN samples zeroed before and restored after, normalised by ||tx||^2 (get_norm_fac :372-391), then
// _get_power_from_complex (calibrate_ek.py:483-490) and the Sv / TS chain of _cal_complex_samples (:564-637).
// scipy is free to evaluate that convolution by FFT; so is this kernel:
//
//   * a ping is cut into segments of L = 4096 - M + 1 outputs; each segment is one 4096-point circular correlation
//     IFFT(FFT(x) * conj(FFT(tx))) whose first L outputs are exact (overlap-save).  8 M flop per sample become
//     ~ 2 * 5 N log2 N / L: 17x less arithmetic at M = 277, which moves K3 from the FP32 pipes to the HBM roofline.
//   * 256 threads, 16 points per thread, radix 16 x 16 x 16.  Forward = decimation in frequency (digit-reversed
//     spectrum), inverse = decimation in time from the digit-reversed spectrum: no reordering pass, every pass is in
//     place on the same 16 positions {i + m j}.  The last forward pass, the spectrum product and the first inverse
//     pass touch the same 16 consecutive points and stay in registers.  The first forward pass reads its points
//     straight from global memory (coalesced: position t + 256 j), the last inverse pass leaves the outputs of the same
//     positions in registers for the epilogue: 4 shared-memory round trips per segment.
//   * shared-memory index p -> p + (p >> 4) (one pad per 16 points): every pass is bank-conflict free.
//   * twiddles come from tables computed in float64 (sincospi) by the setup kernel, laid out k-major so a warp reads
//     consecutive entries; the replica spectrum H = conj(FFT(tx)) / (N ||tx||^2) is computed by the same setup kernel
//     (direct float64 DFT of the M taps) in the register order of the fused pass.
//   * the four beams are summed on the way in (the correlation is linear and NaN padding is normally identical across
//     beams); a segment whose beams carry different NaN masks runs the transform once per beam and forms the exact
//     nanmean (xarray mean(dim="beam") skips NaN, calibrate_ek.py:484).
// Precision: float32 transforms with exact twiddles give ~1.6e-7 of the ping's RMS output per sample (the direct
// float32 sum over 277 taps: 3.7e-7); the reference itself stores the compressed samples as complex64.
#include "sample_math.cuh"

namespace {
using namespace epb;

constexpr int kN = 4096;
constexpr int kT = 256;
constexpr int kPadN = kN + kN / 16;
constexpr int kMaxChan = 32;

struct FftParams {
  const float* re;
  const float* im;
  const float2* H;    // [C][16][256]
  const float2* tw1;  // [16][256]  w_4096^(i k), k major
  const float2* tw2;  // [16][16]   w_256^(i k), k major
  const int* lead0;   // [C] number of leading taps that are exactly zero
  const epb_row* rows;
  float* out;
  float* rng;
  float2* pc_out;
  float* minmax;
  long long nrows, P;
  int R, L, nseg;
};

__device__ __forceinline__ int pad(int p) { return p + (p >> 4); }
// TMA prefetch of a contiguous global range into L2 (one instruction, no registers held): the next segment's samples
// are already on chip when its loads are issued
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
// complex add / subtract as ONE packed FP32 instruction each (Blackwell FADD2 / FFMA2: two lanes per instruction): the
// butterflies are ~60 % of the transform's instructions
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return fadd2(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return ffma2(b, make_float2(-1.f, -1.f), a); }

template <bool kInv>
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  if (!kInv) {  // forward: w4 = -i
    x1 = make_float2(t1.x + t3.y, t1.y - t3.x);
    x3 = make_float2(t1.x - t3.y, t1.y + t3.x);
  } else {
    x1 = make_float2(t1.x - t3.y, t1.y + t3.x);
    x3 = make_float2(t1.x + t3.y, t1.y - t3.x);
  }
}

// 16-point DFT of v[0..15] (natural order in, natural order out); kInv: conjugate kernel, no scaling
template <bool kInv>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
  for (int a = 0; a < 4; ++a) dft4<kInv>(v[a], v[a + 4], v[a + 8], v[a + 12]);  // v[a + 4 d] = S_a[d]
  // twiddle w16^(a d), forward (cos, -sin); inverse conjugate
  const float sg = kInv ? 1.f : -1.f;
  // exponents: (a,d) -> a*d : 1:(1,1) 2:(1,2),(2,1) 3:(1,3),(3,1) 4:(2,2) 6:(2,3),(3,2) 9:(3,3)
  auto tw = [&](float2& z, float cr, float ci) { z = make_float2(z.x * cr - z.y * (sg * ci), z.x * (sg * ci) + z.y * cr); };
  tw(v[1 + 4 * 1], c1, s1);
  tw(v[1 + 4 * 2], h, h);
  tw(v[2 + 4 * 1], h, h);
  tw(v[1 + 4 * 3], s1, c1);
  tw(v[3 + 4 * 1], s1, c1);
  {  // exponent 4: multiply by -i (forward) / +i (inverse)
    float2& z = v[2 + 4 * 2];
    z = kInv ? make_float2(-z.y, z.x) : make_float2(z.y, -z.x);
  }
  tw(v[2 + 4 * 3], -h, h);
  tw(v[3 + 4 * 2], -h, h);
  tw(v[3 + 4 * 3], -c1, -s1);
#pragma unroll
  for (int d = 0; d < 4; ++d) dft4<kInv>(v[4 * d], v[4 * d + 1], v[4 * d + 2], v[4 * d + 3]);  // v[4 d + c] = X[d + 4 c]
  float2 o[16];
#pragma unroll
  for (int d = 0; d < 4; ++d)
#pragma unroll
    for (int c = 0; c < 4; ++c) o[d + 4 * c] = v[4 * d + c];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = o[k];
}

// one 4096-point circular correlation with the replica spectrum: v[j] = x[t + 256 j] in, y[t + 256 j] out
__device__ __forceinline__ void correlate(float2 (&v)[16], float2* __restrict__ s, const float2* __restrict__ s_tw1,
                                          const float2* __restrict__ s_tw2, const float2* __restrict__ Hc, int t) {
  // ---- forward pass 1 (n = 4096, m = 256): registers -> shared ---------------------------------------------------
  dft16<false>(v);
  s[pad(t)] = v[0];
#pragma unroll
  for (int k = 1; k < 16; ++k) s[pad(t + 256 * k)] = cmul(v[k], s_tw1[k * 256 + t]);
  __syncthreads();
  // ---- forward pass 2 (n = 256, m = 16) ------------------------------------------------------------------------------
  const int blk = t >> 4, i2 = t & 15;
  const int b2 = blk * 256 + i2;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = s[pad(b2 + 16 * j)];
  dft16<false>(v);
  s[pad(b2)] = v[0];
#pragma unroll
  for (int k = 1; k < 16; ++k) s[pad(b2 + 16 * k)] = cmul(v[k], s_tw2[k * 16 + i2]);
  __syncthreads();
  // ---- forward pass 3 (n = 16) * H, inverse pass 1 (n = 16): registers only -------------------------------------------
  const int b3 = 17 * t;  // pad(16 t + j) = 17 t + j
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = s[b3 + j];
  dft16<false>(v);
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = cmul(v[k], __ldg(Hc + k * 256 + t));
  dft16<true>(v);
#pragma unroll
  for (int j = 0; j < 16; ++j) s[b3 + j] = v[j];
  __syncthreads();
  // ---- inverse pass 2 (n = 256, m = 16) -------------------------------------------------------------------------------
  v[0] = s[pad(b2)];
#pragma unroll
  for (int k = 1; k < 16; ++k) v[k] = cmulc(s[pad(b2 + 16 * k)], s_tw2[k * 16 + i2]);
  dft16<true>(v);
#pragma unroll
  for (int j = 0; j < 16; ++j) s[pad(b2 + 16 * j)] = v[j];
  __syncthreads();
  // ---- inverse pass 3 (n = 4096, m = 256): shared -> registers ------------------------------------------------------------
  v[0] = s[pad(t)];
#pragma unroll
  for (int k = 1; k < 16; ++k) v[k] = cmulc(s[pad(t + 256 * k)], s_tw1[k * 256 + t]);
  dft16<true>(v);
  __syncthreads();  // the next transform overwrites the work buffer
}

// One 4096-point transform ("item").  Normally it holds one segment of one ping.  The LAST segment of a ping is usually
// short (cfg3: 576 of 3808 outputs), so the tails of up to four consecutive pings of a channel are packed into one
// transform, each in its own sub-block of 1024 (or 2048) positions: a sub-block's correlation window (outputs + M - 1
// samples) fits inside it and everything else is zero, so the circular correlation never mixes pings.
struct Item {
  long long row0;  // sub-block b holds row0 + b (b < nsub)
  int nsub;        // sub-blocks that hold a ping
  int shift;       // point j lies in sub-block j >> shift (4: one block of 4096, 3: two of 2048, 2: four of 1024)
  int n0;          // first sample of every sub-block
  int nout;        // outputs per sub-block
};

struct Plan {  // launch-wide constants of the item enumeration
  int L, nfull, pack, tail_n0, tail_nout, groups_per_chan, nchan;
};

// group g (32-bit: C * ceil(P / pack) groups) -> channel, first row, rows in the group
struct Group {
  int c, rows;
  long long first;
};
__device__ __forceinline__ Group make_group(const Plan& pl, int P, int g) {
  Group gr;
  gr.c = g / pl.groups_per_chan;
  const int gi = g - gr.c * pl.groups_per_chan;
  gr.first = (long long)gr.c * P + (long long)gi * pl.pack;
  gr.rows = min(pl.pack, P - gi * pl.pack);
  return gr;
}
// y -> Sv / TS, echo_range, optional compressed samples and min / max.  SHIFT: sub-block of point j = j >> SHIFT
// (4: the whole transform is one segment).
// cnt4: 4 bits per point = number of beams whose sample is valid (the nanmean divisor); nanre0: bit j = beam-0 real part NaN.
template <int B, int SHIFT, bool kUniform>
__device__ __forceinline__ void epilogue(const FftParams& pr, const float2 (&v)[16], const Item& it, int t, unsigned long long cnt4,
                                         unsigned okmask, unsigned nanre0, const int* __restrict__ s_zf, MinMax& mm_v, MinMax& mm_r) {
  constexpr int kPer = 1 << SHIFT, kSmask = (256 << SHIFT) - 1;
  const int R = pr.R;
#pragma unroll
  for (int sb = 0; sb < 16 / kPer; ++sb) {
    if (SHIFT != 4 && sb >= it.nsub) break;
    const RowF rc = load_rowf(pr.rows + it.row0 + sb);
    const long long base = (it.row0 + sb) * (long long)R + it.n0;
    const int zero_from = s_zf[sb];
    const int nout = min(it.nout, R - it.n0);
#pragma unroll
    for (int jj = 0; jj < kPer; ++jj) {
      const int j = sb * kPer + jj;
      const int q = (t + 256 * j) & kSmask;
      if (q < nout) {
        const int n = it.n0 + q;
        float sc;
        bool any;
        if (kUniform) {  // every beam valid or none
          sc = 1.f / (float)B;
          any = (okmask >> j) & 1u;
        } else {
          const int cnt = (int)((cnt4 >> (4 * j)) & 15ull);
          sc = 1.f;  // 1 / cnt as selects
          sc = (cnt == 2) ? 0.5f : sc;
          sc = (cnt == 3) ? (1.f / 3.f) : sc;
          sc = (cnt == 4) ? 0.25f : sc;
          any = cnt > 0;
        }
        // outputs whose every product is exactly zero (the replica starts with zero taps - the Hann taper does - and only
        // zero samples follow) are exactly zero in the reference's sum, hence NaN after prx.where(prx > 0)
        // (calibrate_ek.py:581); a transform would leave rounding noise there
        const bool exact0 = q >= zero_from;
        const float mr = exact0 ? 0.f : v[j].x * sc, mi = exact0 ? 0.f : v[j].y * sc;
        float prx = rc.fscale * fmaf(mr, mr, mi * mi);
        // deep in a null the transform's rounding noise is quantised and can be exactly (0, 0) where the true sum is tiny
        // but not zero: keep such samples finite (the smallest normal power) so that NaN marks only what is NaN in the
        // reference - structurally zero sums (exact0) and invalid samples
        prx = (!exact0 && prx == 0.f) ? 1.17549435e-38f : prx;
        const float fr = (any && prx > 0.f) ? fmaf(kLog2ToDb, fast_log2(prx), rc.foffK) : CUDART_NAN_F;  // calibrate_ek.py:581
        const float nf = (float)n;
        float rr = range_of(rc, nf);
        float o = sv_db(rc, n, nf, fr);
        if ((nanre0 >> j) & 1u) rr = CUDART_NAN_F, o = CUDART_NAN_F;  // range.py:143-145: NaN where beam 0 of backscatter_r is
        st_stream(pr.out + base + q, o);
        if (pr.rng) st_stream(pr.rng + base + q, rr);
        if (pr.pc_out) pr.pc_out[base + q] = any ? make_float2(mr, mi) : make_float2(CUDART_NAN_F, CUDART_NAN_F);
        if (pr.minmax) {
          mm_v.add(o);
          mm_r.add(rr);
        }
      }
    }
  }
}

// cold path: the beams of this item carry different NaN masks.  One transform per beam; beam b's output counts only
// where beam b's input sample is valid (the exact nanmean over beams of calibrate_ek.py:484).  Everything is passed by
// value: a reference to the kernel parameters or to the caller's running min / max would force them into local memory
// on the hot path as well.
template <int B, int SHIFT>
__device__ __noinline__ void per_beam_item(const FftParams pr, float2* __restrict__ s, const float2* __restrict__ s_tw1,
                                           const float2* __restrict__ s_tw2, const int* __restrict__ s_zf, const float2* __restrict__ Hc,
                                           const Item it, int t, unsigned nanre0) {
  constexpr int kSmask = (256 << SHIFT) - 1;
  MinMax mm_v, mm_r;
  float2 acc[16], v[16];
  unsigned long long valid = 0ull, cnt4 = 0ull;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    acc[j] = make_float2(0.f, 0.f);
    const int sb = j >> SHIFT, n = it.n0 + ((t + 256 * j) & kSmask);
    unsigned fl = 0u;
    if (sb < it.nsub && n < pr.R) {
      const long long e = ((it.row0 + sb) * (long long)pr.R + n) * B;
      for (int b = 0; b < B; ++b) {
        const float xr = pr.re[e + b], xi = pr.im[e + b];
        fl |= (xr == xr && xi == xi) ? (1u << b) : 0u;  // a complex sample is NaN if either part is
      }
    }
    valid |= (unsigned long long)fl << (4 * j);
    cnt4 |= (unsigned long long)__popc(fl) << (4 * j);
  }
  for (int b = 0; b < B; ++b) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int sb = j >> SHIFT, n = it.n0 + ((t + 256 * j) & kSmask);
      const long long e = ((it.row0 + sb) * (long long)pr.R + n) * B + b;
      const bool ok = (valid >> (4 * j + b)) & 1ull;
      v[j] = ok ? make_float2(pr.re[e], pr.im[e]) : make_float2(0.f, 0.f);
    }
    correlate(v, s, s_tw1, s_tw2, Hc, t);
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if ((valid >> (4 * j + b)) & 1ull) acc[j] = cadd(acc[j], v[j]);
  }
  epilogue<B, SHIFT, false>(pr, acc, it, t, cnt4, 0u, nanre0, s_zf, mm_v, mm_r);
  if (pr.minmax) {
    mm_v.flush(pr.minmax + 0, pr.minmax + 1);
    mm_r.flush(pr.minmax + 2, pr.minmax + 3);
  }
}

// load + transform + epilogue of one item
template <int B, int SHIFT>
__device__ __forceinline__ void process_item(const FftParams& pr, const Item& it, float2* __restrict__ s, const float2* __restrict__ s_tw1,
                                             const float2* __restrict__ s_tw2, int* __restrict__ s_wmax, int* __restrict__ s_zf,
                                             const float2* __restrict__ Hc, int lead0, int t, MinMax& mm_v, MinMax& mm_r) {
  constexpr int kPer = 1 << SHIFT, kSmask = (256 << SHIFT) - 1;
  const int R = pr.R;
  // ---- load: beams summed; four points (8 x 16 bytes) in flight.  A NaN anywhere in the point makes the plain sum NaN:
  //      only then the per-beam validity is worked out (NaN -> 0). -------------------------------------------------------------
  float2 v[16];
  unsigned okmask = 0u;   // bit j: every beam of point j is valid
  unsigned nanre0 = 0u;   // bit j: beam-0 real part is NaN (range.py:143-145)
  unsigned nzmask = 0u;   // bit j: the (summed) sample is nonzero
  int nonuniform = 0;     // some point has valid and invalid beams: the per-beam path
#pragma unroll
  for (int gq = 0; gq < 4; ++gq) {
    float xr[4][B], xi[4][B];
    bool inside[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = 4 * gq + u;
      const int sb = j >> SHIFT, n = it.n0 + ((t + 256 * j) & kSmask);
      inside[u] = (SHIFT == 4 || sb < it.nsub) && (n < R);
      const long long e = inside[u] ? ((it.row0 + sb) * (long long)R + n) * B : 0;  // clamped address; value discarded
      if (B == 4) {
        const float4 a4 = ld_stream4(reinterpret_cast<const float4*>(pr.re + e));
        const float4 b4 = ld_stream4(reinterpret_cast<const float4*>(pr.im + e));
        xr[u][0] = a4.x, xr[u][1 % B] = a4.y, xr[u][2 % B] = a4.z, xr[u][3 % B] = a4.w;
        xi[u][0] = b4.x, xi[u][1 % B] = b4.y, xi[u][2 % B] = b4.z, xi[u][3 % B] = b4.w;
      } else {
#pragma unroll
        for (int bb = 0; bb < B; ++bb) {
          xr[u][bb] = ld_stream(pr.re + e + bb);
          xi[u][bb] = ld_stream(pr.im + e + bb);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = 4 * gq + u;
      float sr = xr[u][0], si = xi[u][0];
#pragma unroll
      for (int bb = 1; bb < B; ++bb) sr += xr[u][bb], si += xi[u][bb];
      const bool clean = (sr == sr) && (si == si);  // finite inputs: the sum is NaN iff some part is NaN
      if (inside[u] && !clean) {                     // rare: NaN padding / missing sectors
        unsigned fl = 0u;
        sr = 0.f, si = 0.f;
#pragma unroll
        for (int bb = 0; bb < B; ++bb) {
          const bool ok = (xr[u][bb] == xr[u][bb]) && (xi[u][bb] == xi[u][bb]);
          sr += ok ? xr[u][bb] : 0.f;
          si += ok ? xi[u][bb] : 0.f;
          fl |= ok ? (1u << bb) : 0u;
        }
        nanre0 |= (xr[u][0] != xr[u][0]) ? (1u << j) : 0u;
        nonuniform |= (fl != 0u);  // fl != all here
      }
      okmask |= (inside[u] && clean) ? (1u << j) : 0u;
      v[j] = inside[u] ? make_float2(sr, si) : make_float2(0.f, 0.f);
      nzmask |= (inside[u] && (sr != 0.f || si != 0.f)) ? (1u << j) : 0u;
    }
  }
  // last nonzero input position per sub-block (this thread: highest set bit of its points in the sub-block)
#pragma unroll
  for (int sb = 0; sb < 16 / kPer; ++sb) {
    const unsigned bits = (nzmask >> (sb * kPer)) & ((1u << kPer) - 1u);
    int last = bits ? ((t + 256 * (sb * kPer + 31 - __clz((int)bits))) & kSmask) : -1;
    last = __reduce_max_sync(0xffffffffu, last);
    if ((t & 31) == 0) s_wmax[(t >> 5) * 4 + sb] = last;
  }
  nonuniform = __syncthreads_or(nonuniform);
  if (t < 16 / kPer) {
    int last = -1;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) last = max(last, s_wmax[w * 4 + t]);
    s_zf[t] = last - lead0 + 1;  // outputs q >= this see only zero taps or zero samples (read after correlate's barriers)
  }
  if (!nonuniform) {
    correlate(v, s, s_tw1, s_tw2, Hc, t);
    epilogue<B, SHIFT, true>(pr, v, it, t, 0ull, okmask, nanre0, s_zf, mm_v, mm_r);
  } else {
    per_beam_item<B, SHIFT>(pr, s, s_tw1, s_tw2, s_zf, Hc, it, t, nanre0);
  }
}

// One launch per item kind, so that each kernel carries ONE inlined copy of the transform (with both kinds in one kernel
// 12 % of the issue slots were lost to instruction fetch): SHIFT == 4 walks the full segments (row, seg < nfull) of all
// rows, SHIFT in {2, 3} the packed tails of the row groups.  A CTA owns a contiguous range of items, so the next item
// (prefetched into L2 by TMA while this one is transformed) follows in memory.
template <int B, int SHIFT>
__global__ void __launch_bounds__(kT, 2) pulse_fft_kernel(const FftParams pr, const Plan pl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s = reinterpret_cast<float2*>(smem_raw);  // [kPadN]
  float2* s_tw1 = s + kPadN;                         // [4096]
  float2* s_tw2 = s_tw1 + 4096;                      // [256]
  int* s_wmax = reinterpret_cast<int*>(s_tw2 + 256);  // [8][4] per warp and sub-block: last nonzero input position
  int* s_zf = s_wmax + 32;                            // [4] per sub-block: first structurally-zero output
  const int t = threadIdx.x;
  for (int k = t; k < 4096; k += kT) s_tw1[k] = pr.tw1[k];
  s_tw2[t] = pr.tw2[t];
  __syncthreads();
  const int R = pr.R;
  const int P = (int)pr.P;
  // items of this launch and this CTA's contiguous share of them
  const long long nitems = (SHIFT == 4) ? pr.nrows * (long long)pl.nfull : (long long)pl.nchan * pl.groups_per_chan;
  const long long i0 = nitems * (long long)blockIdx.x / gridDim.x, i1 = nitems * (long long)(blockIdx.x + 1) / gridDim.x;
  auto item_of = [&](long long i, int* chan) {
    Item it;
    if (SHIFT == 4) {
      const long long row = i / pl.nfull;
      it.row0 = row, it.nsub = 1, it.shift = 4, it.n0 = (int)(i - row * pl.nfull) * pl.L, it.nout = pl.L;
      *chan = (int)(row / P);
    } else {
      const Group gr = make_group(pl, P, (int)i);
      it.row0 = gr.first, it.nsub = gr.rows, it.shift = SHIFT, it.n0 = pl.tail_n0, it.nout = pl.tail_nout;
      *chan = gr.c;
    }
    return it;
  };
  MinMax mm_v, mm_r;
  for (long long i = i0; i < i1; ++i) {
    int c;
    const Item it = item_of(i, &c);
    const float2* Hc = pr.H + (size_t)c * kN;
    const int lead0 = __ldg(pr.lead0 + c);
    if (B == 4 && t == 0 && i + 1 < i1) {  // while this item is loaded and transformed, the next one streams from HBM into L2
      int cn;
      const Item nx = item_of(i + 1, &cn);
      const int len = min(256 << nx.shift, R - nx.n0);
      for (int sb = 0; sb < nx.nsub; ++sb) {
        const long long e = ((nx.row0 + sb) * (long long)R + nx.n0) * 4;
        prefetch_l2_bulk(pr.re + e, (unsigned)len * 16u);
        prefetch_l2_bulk(pr.im + e, (unsigned)len * 16u);
      }
    }
    process_item<B, SHIFT>(pr, it, s, s_tw1, s_tw2, s_wmax, s_zf, Hc, lead0, t, mm_v, mm_r);
  }
  if (pr.minmax) {
    mm_v.flush(pr.minmax + 0, pr.minmax + 1);
    mm_r.flush(pr.minmax + 2, pr.minmax + 3);
  }
}

// twiddle tables and the replica spectra, float64 arithmetic (exact to float32 rounding)
__global__ void __launch_bounds__(256) pulse_fft_setup_kernel(const float2* __restrict__ replica, const double* __restrict__ inv_norm,
                                                              int C, const int* __restrict__ off, float2* __restrict__ H,
                                                              float2* __restrict__ tw1, float2* __restrict__ tw2, int* __restrict__ lead0) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.y == 0) {  // tables
    if (g < C) {
      int z = 0;
      const int M = off[g + 1] - off[g];
      while (z < M && replica[off[g] + z].x == 0.f && replica[off[g] + z].y == 0.f) ++z;
      lead0[g] = z;
    }
    if (g < 4096) {
      const int k = g >> 8, i = g & 255;
      double sn, cs;
      sincospi(-2.0 * (double)(i * k) / 4096.0, &sn, &cs);
      tw1[g] = make_float2((float)cs, (float)sn);
    }
    if (g < 256) {
      const int k = g >> 4, i = g & 15;
      double sn, cs;
      sincospi(-2.0 * (double)(i * k) / 256.0, &sn, &cs);
      tw2[g] = make_float2((float)cs, (float)sn);
    }
    return;
  }
  const int c = blockIdx.y - 1;
  if (c >= C || g >= kN) return;
  // entry g = k * 256 + t of channel c: spectrum at the natural frequency f of work position 16 t + k (digit reversed)
  const int k = g >> 8, t = g & 255;
  const int f = (t >> 4) + 16 * (t & 15) + 256 * k;
  const int M = off[c + 1] - off[c];
  const float2* tx = replica + off[c];
  double ar = 0.0, ai = 0.0;
  for (int m = 0; m < M; ++m) {  // T[f] = sum_m tx[m] exp(-2 pi i f m / N)
    double sn, cs;
    sincospi(-2.0 * (double)(((long long)f * m) % kN) / (double)kN, &sn, &cs);
    const double xr = tx[m].x, xi = tx[m].y;
    ar += xr * cs - xi * sn;
    ai += xr * sn + xi * cs;
  }
  const double scale = inv_norm[c] / (double)kN;
  H[(size_t)c * kN + g] = make_float2((float)(ar * scale), (float)(-ai * scale));  // conj(T) / (N ||tx||^2)
}

constexpr size_t kSmemBytes = (size_t)(kPadN + 4096 + 256) * sizeof(float2) + 40 * sizeof(int);

}  // namespace

extern "C" epb_i64 epb_pulse_fft_workspace_bytes(epb_i64 C) {
  if (C <= 0 || C > kMaxChan) return 0;
  return (epb_i64)((size_t)C * kN + 4096 + 256) * (epb_i64)sizeof(float2) + (epb_i64)(2 * kMaxChan + 1) * (epb_i64)sizeof(int) + 64;
}

extern "C" int epb_pulse_fft_max_taps(void) { return kN / 2 + 1; }

extern "C" int epb_pulse_compress_sv_fft(const float* re, const float* im, const float* replica, const int* h_replica_off,
                                         const double* inv_norm, const epb_row* rows, float* out, float* echo_range,
                                         float* pc_out, float* minmax, epb_i64 C, epb_i64 P, epb_i64 R, int B,
                                         void* workspace, epb_i64 workspace_bytes, void* stream) {
  EPB_REQUIRE(re && im && replica && h_replica_off && inv_norm && rows && out && workspace, "NULL pointer");
  EPB_REQUIRE(C > 0 && C <= kMaxChan && P > 0 && R > 0 && R < (1LL << 24), "bad shape (channel <= 32)");
  EPB_REQUIRE(B >= 1 && B <= 4, "B must be 1..4");
  EPB_REQUIRE(B != 4 || (((uintptr_t)re | (uintptr_t)im) % 16 == 0), "4-beam planes must be 16-byte aligned");
  EPB_REQUIRE(((uintptr_t)replica % 8) == 0 && ((uintptr_t)pc_out % 8) == 0, "replica / pc_out must be 8-byte aligned");
  EPB_REQUIRE(workspace_bytes >= epb_pulse_fft_workspace_bytes(C) && ((uintptr_t)workspace % 16) == 0, "workspace too small / misaligned");
  int Mmax = 0;
  int h_off[kMaxChan + 1];
  for (int c = 0; c <= C; ++c) h_off[c] = h_replica_off[c];
  for (int c = 0; c < C; ++c) {
    const int M = h_off[c + 1] - h_off[c];
    EPB_REQUIRE(M > 0, "empty replica");
    Mmax = M > Mmax ? M : Mmax;
  }
  if (Mmax > epb_pulse_fft_max_taps()) {
    epb_set_error("epb_pulse_compress_sv_fft: %d-tap replica exceeds %d (use epb_pulse_compress_sv)", Mmax, epb_pulse_fft_max_taps());
    return EPB_E_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float2* H = reinterpret_cast<float2*>(workspace);
  float2* tw1 = H + (size_t)C * kN;
  float2* tw2 = tw1 + 4096;
  int* d_off = reinterpret_cast<int*>(tw2 + 256);
  int* d_lead0 = d_off + kMaxChan + 1;
  if (cudaMemcpyAsync(d_off, h_off, sizeof(int) * (C + 1), cudaMemcpyHostToDevice, st) != cudaSuccess)
    return epb_check_launch("epb_pulse_compress_sv_fft(offsets)");
  pulse_fft_setup_kernel<<<dim3(kN / 256, (unsigned)C + 1), 256, 0, st>>>(reinterpret_cast<const float2*>(replica), inv_norm, (int)C, d_off, H,
                                                                          tw1, tw2, d_lead0);
  int rc0 = epb_check_launch("epb_pulse_compress_sv_fft(setup)");
  if (rc0 != 0) return rc0;
  FftParams pr;
  pr.re = re, pr.im = im, pr.H = H, pr.tw1 = tw1, pr.tw2 = tw2, pr.lead0 = d_lead0, pr.rows = rows;
  pr.out = out, pr.rng = echo_range, pr.pc_out = reinterpret_cast<float2*>(pc_out), pr.minmax = minmax;
  pr.nrows = C * P, pr.P = P, pr.R = (int)R;
  int L = kN - Mmax + 1;
  if (L > 64) L &= ~31;  // warp-aligned segment starts keep the stores of a segment 128-byte aligned
  pr.L = L;
  pr.nseg = (int)((R + L - 1) / L);
  Plan pl;
  const int nlast = (int)(R - (long long)(pr.nseg - 1) * L);  // outputs of a ping's last segment
  pl.L = L;
  pl.pack = (B != 4) ? 1 : (nlast + Mmax - 1 <= 1024) ? 4 : (nlast + Mmax - 1 <= 2048 ? 2 : 1);
  if (pl.pack > 1) {
    pl.nfull = pr.nseg - 1, pl.tail_n0 = (pr.nseg - 1) * L, pl.tail_nout = nlast;
  } else {
    pl.nfull = pr.nseg, pl.tail_n0 = 0, pl.tail_nout = 0;
  }
  EPB_REQUIRE(P < (1LL << 30), "ping_time too long for one launch");
  pl.groups_per_chan = (int)((P + pl.pack - 1) / pl.pack);
  pl.nchan = (int)C;
  const long long ngroups = C * (long long)pl.groups_per_chan;
  const long long cap = (long long)epb_num_sms() * 2;
#define EPB_FFT(BB, SH, NITEMS)                                                                                              \
  do {                                                                                                                       \
    const long long n_ = (NITEMS);                                                                                           \
    if (n_ > 0) {                                                                                                            \
      auto kern = pulse_fft_kernel<BB, SH>;                                                                                  \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess)           \
        return epb_check_launch("epb_pulse_compress_sv_fft(smem)");                                                          \
      kern<<<(unsigned)(n_ < cap ? n_ : cap), kT, kSmemBytes, st>>>(pr, pl);                                                 \
    }                                                                                                                        \
  } while (0)
  const long long nseg_items = pr.nrows * (long long)pl.nfull;
  switch (B) {
    case 1: EPB_FFT(1, 4, nseg_items); break;
    case 2: EPB_FFT(2, 4, nseg_items); break;
    case 3: EPB_FFT(3, 4, nseg_items); break;
    default:
      EPB_FFT(4, 4, nseg_items);                     // the full segments of every row
      if (pl.pack == 4) EPB_FFT(4, 2, ngroups);      // then the short last segments, four pings per transform
      if (pl.pack == 2) EPB_FFT(4, 3, ngroups);      // ... or two
      break;
  }
#undef EPB_FFT
  return epb_check_launch("epb_pulse_compress_sv_fft");
}
