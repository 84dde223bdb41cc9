// Synthetic echosounder volumes generated on the device with Philox4x32-10.
// counter = (float4 index within the row, global ping index, channel, stream id), key = seed, so a
// ping-sharded rank generates exactly the slice of the global volume it owns (ping_offset).
#include "epb_common.cuh"

namespace {

struct U4 {
  unsigned x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(U4 c, unsigned k0, unsigned k1) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    U4 n = {hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

__device__ __forceinline__ float gen_value(unsigned u, unsigned u2, int kind, float scale) {
  if (kind == 0) {  // int16-quantised EK power, convert/parse_base.py:24,302: q * 10log10(2)/256
    int q = -24000 + (int)__umulhi(u, 22001u);
    return epb::count_to_db_f((float)q);
  }
  if (kind == 1) return (float)(u >> 16);  // AZFP counts 0..65535
  // kind 2: N(0,1)*scale via Box-Muller on (u, u2)
  float a = ((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float b = ((float)(u2 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return scale * sqrtf(-2.0f * logf(a)) * cospif(2.0f * b);
}

__global__ void __launch_bounds__(256) synth_kernel(float* __restrict__ out, long long C, long long P, long long row_len,
                                                    int inner, int kind, unsigned k0, unsigned k1, long long ping_offset,
                                                    unsigned nan_tail_q16, float scale) {
  __shared__ long long s_cut;
  const long long nrows = C * P;
  const long long n4 = (row_len + 3) / 4;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    long long c = row / P, pg = row % P + ping_offset;
    if (threadIdx.x == 0) {
      U4 h = philox4x32_10({0u, (unsigned)pg, (unsigned)c, 1u + (unsigned)(pg >> 32)}, k0, k1);
      long long R = row_len / inner;
      long long cut = R;
      if ((h.x & 0xffffu) < nan_tail_q16) cut = R / 4 + (long long)(h.y % (unsigned)(R - R / 4));
      s_cut = cut * inner;
    }
    __syncthreads();
    const long long cut = s_cut;
    float* o = out + row * row_len;
    for (long long j = threadIdx.x; j < n4; j += blockDim.x) {
      U4 r = philox4x32_10({(unsigned)j, (unsigned)pg, (unsigned)c, (unsigned)(pg >> 32) << 8}, k0, k1);
      unsigned u[4] = {r.x, r.y, r.z, r.w};
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        long long e = 4 * j + k;
        v[k] = gen_value(u[k], u[(k + 1) & 3] ^ 0x9E3779B9u, kind, scale);
        if (e >= cut) v[k] = CUDART_NAN_F;
      }
      if (4 * j + 3 < row_len && ((row * row_len) % 4 == 0)) {
        *reinterpret_cast<float4*>(o + 4 * j) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        for (int k = 0; k < 4; ++k)
          if (4 * j + k < row_len) o[4 * j + k] = v[k];
      }
    }
    __syncthreads();
  }
}

// kind 0 as raw counts: the same Philox stream and quantisation as synth_kernel, int16 out, -32768 past the cut
// (the padding marker of the ingest format, ingest.cu), so that epb_ingest_power_i16(counts) == epb_synth_fill(kind 0).
__global__ void __launch_bounds__(256) synth_i16_kernel(short* __restrict__ out, long long C, long long P, long long R,
                                                        unsigned k0, unsigned k1, long long ping_offset,
                                                        unsigned nan_tail_q16) {
  __shared__ long long s_cut;
  const long long nrows = C * P;
  const long long n4 = (R + 3) / 4;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    long long c = row / P, pg = row % P + ping_offset;
    if (threadIdx.x == 0) {
      U4 h = philox4x32_10({0u, (unsigned)pg, (unsigned)c, 1u + (unsigned)(pg >> 32)}, k0, k1);
      long long cut = R;
      if ((h.x & 0xffffu) < nan_tail_q16) cut = R / 4 + (long long)(h.y % (unsigned)(R - R / 4));
      s_cut = cut;
    }
    __syncthreads();
    const long long cut = s_cut;
    short* o = out + row * R;
    for (long long j = threadIdx.x; j < n4; j += blockDim.x) {
      U4 r = philox4x32_10({(unsigned)j, (unsigned)pg, (unsigned)c, (unsigned)(pg >> 32) << 8}, k0, k1);
      unsigned u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        long long e = 4 * j + k;
        if (e < R) o[e] = (e >= cut) ? (short)-32768 : (short)(-24000 + (int)__umulhi(u[k], 22001u));
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int epb_synth_fill(float* out, epb_i64 C, epb_i64 P, epb_i64 R, epb_i64 inner, int kind,
                              unsigned long long seed, epb_i64 ping_offset, unsigned nan_tail_q16, float scale,
                              void* stream) {
  EPB_REQUIRE(out && C > 0 && P > 0 && R > 0 && inner > 0, "bad pointer/shape");
  EPB_REQUIRE(kind >= 0 && kind <= 2, "kind must be 0,1,2");
  EPB_REQUIRE(((uintptr_t)out) % 16 == 0, "out must be 16-byte aligned");
  const long long nrows = C * P;
  const int grid = (int)((nrows < (long long)epb_num_sms() * 8) ? nrows : (long long)epb_num_sms() * 8);
  synth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, C, P, R * inner, (int)inner, kind, (unsigned)seed,
                                                       (unsigned)(seed >> 32), ping_offset, nan_tail_q16, scale);
  return epb_check_launch("epb_synth_fill");
}

extern "C" int epb_synth_fill_i16(short* out, epb_i64 C, epb_i64 P, epb_i64 R, unsigned long long seed, epb_i64 ping_offset,
                                  unsigned nan_tail_q16, void* stream) {
  EPB_REQUIRE(out && C > 0 && P > 0 && R > 0, "bad pointer/shape");
  const long long nrows = C * P;
  const int grid = (int)((nrows < (long long)epb_num_sms() * 8) ? nrows : (long long)epb_num_sms() * 8);
  synth_i16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, C, P, R, (unsigned)seed, (unsigned)(seed >> 32), ping_offset,
                                                           nan_tail_q16);
  return epb_check_launch("epb_synth_fill_i16");
}
