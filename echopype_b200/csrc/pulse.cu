// K3: EK80 broadband pulse compression (matched filter) fused with the Sv / TS epilogue.
//
// Reference: compress_pulse + _convolve_per_channel (calibrate/ek80_complex.py:285-369): per (ping, beam) and
// channel  y[n] = sum_{k<M} x[n+k] conj(tx[k]),  n in [0, N)  (x = 0 beyond N; NaN samples are zeroed before and
// restored after the convolution), normalised by ||tx||^2 (get_norm_fac :372-391); then
// CalibrateEK80._get_power_from_complex (calibrate_ek.py:483-490): prx = B |nanmean_b y_b|^2 / 8 (|z_er+z_et|/z_er)^2 / z_et
// and the Sv / TS chain of _cal_complex_samples (:564-637), folded into the row record like K2.
//
// One CTA per (channel, ping) row (grid-stride).  The ping is staged ONCE in shared memory: the four beams are
// summed on the way in (the convolution is linear, and NaN padding is normally identical across beams, so
// sum_b conv(x_b) = conv(sum_b x_b): 4x less arithmetic); pings whose beams have different NaN masks take a
// per-beam fallback with the exact nanmean.  Each thread produces J = 8 consecutive outputs with a register
// sliding window: per tap one LDS.64 for the new sample, one broadcast LDS.64 for the tap, 32 FFMA.  The staged
// signal is padded by one element every 8 so that the 64-byte lane stride maps to distinct banks.
// Bound: FP32 pipes (8 M flop per output sample vs 36 bytes): see DESIGN.md for both rooflines.
#include "sample_math.cuh"

namespace {
using namespace epb;

constexpr int kJ = 8;          // outputs per thread and window length
constexpr int kThreads = 512;
constexpr int kMaxChan = 32;

struct PulseParams {
  const float* re;
  const float* im;
  const float2* replica;  // concatenated per-channel transmit replicas tx[c][k]
  const double* inv_norm;  // [C] 1 / ||tx||^2
  const epb_row* rows;
  float* out;
  float* rng;
  float2* pc_out;
  float* minmax;
  long long nrows, P;
  int R, Mmax;
  int off[kMaxChan + 1];  // replica offsets per channel
};

__device__ __forceinline__ int phys(int m) { return m + (m >> 3); }  // one pad element every 8

// y[j] += x[n0 + k + j] * conj(tx[k]) for all taps k in [0, M), j in [0, J)
__device__ __forceinline__ void fir_block(const float2* __restrict__ s_x, const float2* __restrict__ s_tx, int n0, int Mpad,
                                          float (&yr)[kJ], float (&yi)[kJ]) {
  float2 w[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) w[j] = s_x[phys(n0 + j)];
  for (int kb = 0; kb < Mpad; kb += kJ) {
#pragma unroll
    for (int i = 0; i < kJ; ++i) {
      const float2 t = s_tx[kb + i];
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const float2 x = w[(i + j) % kJ];
        yr[j] = fmaf(x.x, t.x, yr[j]);
        yr[j] = fmaf(x.y, t.y, yr[j]);
        yi[j] = fmaf(x.y, t.x, yi[j]);
        yi[j] = fmaf(-x.x, t.y, yi[j]);
      }
      w[i] = s_x[phys(n0 + kb + i + kJ)];  // the window slides by one sample
    }
  }
}

template <int B, bool kRange, bool kMinMax>
__global__ void __launch_bounds__(kThreads, 2) pulse_compress_kernel(const PulseParams pr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = pr.R;
  const int Mcap = (pr.Mmax + kJ - 1) / kJ * kJ;            // taps rounded up to a multiple of J (zero padded)
  const int nx = R + Mcap + kJ;                              // staged samples incl. the zero tail
  float2* s_x = reinterpret_cast<float2*>(smem_raw);         // [phys(nx) + 1]
  float2* s_tx = s_x + (phys(nx) + 2);                       // [Mcap]
  unsigned char* s_flags = reinterpret_cast<unsigned char*>(s_tx + Mcap);  // [R] bits 0-3: beam valid, bit 7: beam-0 re NaN
  const int tid = threadIdx.x;
  int cur_c = -1, Mpad = 0;
  MinMax mm_v, mm_r;
  constexpr unsigned kAll = (1u << B) - 1u;

  for (long long row = blockIdx.x; row < pr.nrows; row += gridDim.x) {
    const int c = (int)(row / pr.P);
    __syncthreads();  // previous row finished with s_x / s_tx / s_flags
    if (c != cur_c) {  // new channel: load its replica
      cur_c = c;
      const int M = pr.off[c + 1] - pr.off[c];
      Mpad = (M + kJ - 1) / kJ * kJ;
      for (int k = tid; k < Mcap; k += kThreads) s_tx[k] = (k < M) ? pr.replica[pr.off[c] + k] : make_float2(0.f, 0.f);
    }
    const RowF rc = load_rowf(pr.rows + row);
    const float inv_norm = (float)pr.inv_norm[c];
    const long long base = row * (long long)R;
    // ---- stage: beam sum (NaN -> 0), per-sample beam validity -----------------------------------------------------
    int nonuniform = 0;
    for (int n = tid; n < nx; n += kThreads) {
      float sr = 0.f, si = 0.f;
      if (n < R) {
        float xr[B], xi[B];
        if (B == 4) {
          const float4 a = ld_stream4(reinterpret_cast<const float4*>(pr.re + (base + n) * 4));
          const float4 b = ld_stream4(reinterpret_cast<const float4*>(pr.im + (base + n) * 4));
          xr[0] = a.x, xr[1 % B] = a.y, xr[2 % B] = a.z, xr[3 % B] = a.w;
          xi[0] = b.x, xi[1 % B] = b.y, xi[2 % B] = b.z, xi[3 % B] = b.w;
        } else {
#pragma unroll
          for (int b = 0; b < B; ++b) {
            xr[b] = ld_stream(pr.re + (base + n) * B + b);
            xi[b] = ld_stream(pr.im + (base + n) * B + b);
          }
        }
        unsigned fl = (xr[0] != xr[0]) ? 0x80u : 0u;
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const bool ok = (xr[b] == xr[b]) && (xi[b] == xi[b]);  // complex NaN if either part is NaN
          sr += ok ? xr[b] : 0.f;
          si += ok ? xi[b] : 0.f;
          fl |= ok ? (1u << b) : 0u;
        }
        s_flags[n] = (unsigned char)fl;
        nonuniform |= ((fl & kAll) != 0u && (fl & kAll) != kAll);
      }
      s_x[phys(n)] = make_float2(sr, si);
    }
    nonuniform = __syncthreads_or(nonuniform);

    // ---- convolve + epilogue: J consecutive outputs per thread ----------------------------------------------------
    for (int n0 = tid * kJ; n0 < R; n0 += kThreads * kJ) {
      float yr[kJ], yi[kJ];
#pragma unroll
      for (int j = 0; j < kJ; ++j) yr[j] = 0.f, yi[j] = 0.f;
      if (!nonuniform) fir_block(s_x, s_tx, n0, Mpad, yr, yi);
      unsigned fl[kJ];
#pragma unroll
      for (int j = 0; j < kJ; ++j) fl[j] = (n0 + j < R) ? s_flags[n0 + j] : 0u;
      if (nonuniform) {
        // per-beam fallback (exact nanmean when the beams' NaN masks differ): x_b is gathered straight from
        // global memory (L2-resident after staging); the outputs of beam b count only where beam b is valid
        for (int b = 0; b < B; ++b) {
          float br_[kJ], bi_[kJ];
#pragma unroll
          for (int j = 0; j < kJ; ++j) br_[j] = 0.f, bi_[j] = 0.f;
          for (int k = 0; k < Mpad; ++k) {
            const float2 t = s_tx[k];
#pragma unroll
            for (int j = 0; j < kJ; ++j) {
              const int m = n0 + j + k;
              float xr = 0.f, xi = 0.f;
              if (m < R && ((s_flags[m] >> b) & 1u)) {
                xr = pr.re[(base + m) * B + b];
                xi = pr.im[(base + m) * B + b];
              }
              br_[j] = fmaf(xr, t.x, br_[j]);
              br_[j] = fmaf(xi, t.y, br_[j]);
              bi_[j] = fmaf(xi, t.x, bi_[j]);
              bi_[j] = fmaf(-xr, t.y, bi_[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < kJ; ++j)
            if ((fl[j] >> b) & 1u) yr[j] += br_[j], yi[j] += bi_[j];
        }
      }
      float o[kJ], rr[kJ];
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int n = n0 + j;
        const int cnt = __popc(fl[j] & kAll);
        const float sc = inv_norm / (float)cnt;  // cnt == 0 -> inf * 0 = NaN: nanmean of an all-NaN slice
        const float mr = yr[j] * sc, mi = yi[j] * sc;
        const float prx = rc.fscale * (mr * mr + mi * mi);
        const float fr = (prx > 0.f) ? fmaf(kLog2ToDb, fast_log2(prx), rc.foffK) : CUDART_NAN_F;  // calibrate_ek.py:581
        const float nf = (float)n;
        rr[j] = range_of(rc, nf);
        o[j] = sv_db(rc, n, nf, fr);
        if (fl[j] & 0x80u) rr[j] = CUDART_NAN_F, o[j] = CUDART_NAN_F;  // range.py:143-145: echo_range (hence R', Sv) is NaN where beam 0 of backscatter_r is
        if (n < R) {
          if (pr.pc_out) pr.pc_out[base + n] = make_float2(mr, mi);
          if (kMinMax) {
            mm_v.add(o[j]);
            mm_r.add(rr[j]);
          }
        }
      }
      if (n0 + kJ <= R && (R % 4) == 0) {
        float4* o4 = reinterpret_cast<float4*>(pr.out + base + n0);
        st_stream4(o4, make_float4(o[0], o[1], o[2], o[3]));
        st_stream4(o4 + 1, make_float4(o[4], o[5], o[6], o[7]));
        if (kRange) {
          float4* r4 = reinterpret_cast<float4*>(pr.rng + base + n0);
          st_stream4(r4, make_float4(rr[0], rr[1], rr[2], rr[3]));
          st_stream4(r4 + 1, make_float4(rr[4], rr[5], rr[6], rr[7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < kJ; ++j)
          if (n0 + j < R) {
            pr.out[base + n0 + j] = o[j];
            if (kRange) pr.rng[base + n0 + j] = rr[j];
          }
      }
    }
  }
  if (kMinMax) {
    mm_v.flush(pr.minmax + 0, pr.minmax + 1);
    mm_r.flush(pr.minmax + 2, pr.minmax + 3);
  }
}

size_t pulse_smem(long long R, int Mmax) {
  const int Mcap = (Mmax + kJ - 1) / kJ * kJ;
  const long long nx = R + Mcap + kJ;
  return (size_t)(nx + (nx >> 3) + 2) * 8 + (size_t)Mcap * 8 + (size_t)R + 16;
}

template <int B>
int launch_pulse(const PulseParams& pr, bool range, bool minmax, int grid, size_t smem, cudaStream_t s) {
#define EPB_PC(RG, MM)                                                                                          \
  do {                                                                                                          \
    auto kern = pulse_compress_kernel<B, RG, MM>;                                                               \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1; \
    kern<<<grid, kThreads, smem, s>>>(pr);                                                                      \
  } while (0)
  if (range && minmax)
    EPB_PC(true, true);
  else if (range)
    EPB_PC(true, false);
  else if (minmax)
    EPB_PC(false, true);
  else
    EPB_PC(false, false);
#undef EPB_PC
  return 0;
}

}  // namespace

extern "C" int epb_pulse_compress_sv(const float* re, const float* im, const float* replica, const int* h_replica_off,
                                     const double* inv_norm, const epb_row* rows, float* out, float* echo_range,
                                     float* pc_out, float* minmax, epb_i64 C, epb_i64 P, epb_i64 R, int B, void* stream) {
  EPB_REQUIRE(re && im && replica && h_replica_off && inv_norm && rows && out, "NULL pointer");
  EPB_REQUIRE(C > 0 && C <= kMaxChan && P > 0 && R > 0 && R < (1LL << 24), "bad shape (channel <= 32)");
  EPB_REQUIRE(B >= 1 && B <= 4, "B must be 1..4");
  EPB_REQUIRE(B != 4 || (((uintptr_t)re | (uintptr_t)im) % 16 == 0), "4-beam planes must be 16-byte aligned");
  EPB_REQUIRE(((uintptr_t)replica % 8) == 0 && ((uintptr_t)pc_out % 8) == 0, "replica / pc_out must be 8-byte aligned");
  EPB_REQUIRE((((uintptr_t)out | (uintptr_t)echo_range) % 16) == 0, "out / echo_range must be 16-byte aligned");
  PulseParams pr;
  pr.re = re, pr.im = im, pr.replica = reinterpret_cast<const float2*>(replica), pr.inv_norm = inv_norm, pr.rows = rows;
  pr.out = out, pr.rng = echo_range, pr.pc_out = reinterpret_cast<float2*>(pc_out), pr.minmax = minmax;
  pr.nrows = C * P, pr.P = P, pr.R = (int)R;
  int Mmax = 0;
  for (int c = 0; c <= C; ++c) pr.off[c] = h_replica_off[c];
  for (int c = 0; c < C; ++c) {
    const int M = pr.off[c + 1] - pr.off[c];
    EPB_REQUIRE(M > 0, "empty replica");
    Mmax = M > Mmax ? M : Mmax;
  }
  pr.Mmax = Mmax;
  const size_t smem = pulse_smem(R, Mmax);
  if (smem > 227 * 1024) {
    epb_set_error("epb_pulse_compress_sv: ping of %lld samples with a %d-tap replica needs %zu bytes of shared memory", (long long)R,
                  Mmax, smem);
    return EPB_E_UNSUPPORTED;
  }
  const long long cap = (long long)epb_num_sms() * 2;
  const int grid = (int)(pr.nrows < cap ? pr.nrows : cap);
  int rc = -1;
  switch (B) {
    case 1: rc = launch_pulse<1>(pr, echo_range != nullptr, minmax != nullptr, grid, smem, (cudaStream_t)stream); break;
    case 2: rc = launch_pulse<2>(pr, echo_range != nullptr, minmax != nullptr, grid, smem, (cudaStream_t)stream); break;
    case 3: rc = launch_pulse<3>(pr, echo_range != nullptr, minmax != nullptr, grid, smem, (cudaStream_t)stream); break;
    default: rc = launch_pulse<4>(pr, echo_range != nullptr, minmax != nullptr, grid, smem, (cudaStream_t)stream); break;
  }
  if (rc != 0) return epb_check_launch("epb_pulse_compress_sv(smem)");
  return epb_check_launch("epb_pulse_compress_sv");
}
