// Raw power ingest: the int16 sample counts of the EK60 / EK80 power datagrams -> the float32
// `backscatter_r` of the Beam group, as convert/parse_base.py does on the host:
//   pad_shorter_ping (parse_base.py:686-730): pings shorter than the longest one are padded with NaN;
//   parse_base.py:24,302: power = counts.astype(float32) * INDEX2POWER, INDEX2POWER = 10 log10(2) / 256.
// The device format keeps the counts as int16 [C,P,R] with -32768 (never produced by the instrument: -385 dB) marking
// the padding, so the volume crosses PCIe / HBM at 2 bytes per sample; the conversion is the float32 nearest to
// count * INDEX2POWER (count_to_db_f, epb_common.cuh) and marker -> NaN.  HBM-bound: 2 B read + 4 B written per sample.
#include "epb_common.cuh"

namespace {
using namespace epb;

__device__ __forceinline__ float count_to_db(short q) {
  return (q == (short)-32768) ? CUDART_NAN_F : count_to_db_f((float)q);
}

// eight samples per thread and step: one 16-byte load, two 16-byte streaming stores
__global__ void __launch_bounds__(256) ingest_kernel(const short* __restrict__ counts, float* __restrict__ out, long long n,
                                                     const int* __restrict__ gate) {
  if (gate && *gate == 0) return;  // the fast fused kernel consumed the counts directly
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int4 w = __ldg(reinterpret_cast<const int4*>(counts) + i);
    const int ws[4] = {w.x, w.y, w.z, w.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = count_to_db((short)(ws[k] & 0xffff));
      v[2 * k + 1] = count_to_db((short)(ws[k] >> 16));
    }
    float4* o = reinterpret_cast<float4*>(out) + 2 * i;
    st_stream4(o, make_float4(v[0], v[1], v[2], v[3]));
    st_stream4(o + 1, make_float4(v[4], v[5], v[6], v[7]));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) out[(n8 << 3) + threadIdx.x] = count_to_db(counts[(n8 << 3) + threadIdx.x]);
}

}  // namespace

void epb_ingest_gated_launch(const short* counts, float* out, long long n, const int* gate, cudaStream_t s) {
  const long long n8 = (n >> 3) > 0 ? (n >> 3) : 1;
  const long long cap = (long long)epb_num_sms() * 16;
  const long long blocks = (n8 + 255) / 256;
  ingest_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(counts, out, n, gate);
}

extern "C" int epb_ingest_power_i16(const short* counts, float* backscatter_r, epb_i64 n, void* stream) {
  EPB_REQUIRE(counts && backscatter_r && n > 0, "bad pointer/size");
  EPB_REQUIRE(((uintptr_t)counts % 16) == 0 && ((uintptr_t)backscatter_r % 16) == 0, "arrays must be 16-byte aligned");
  epb_ingest_gated_launch(counts, backscatter_r, n, nullptr, (cudaStream_t)stream);
  return epb_check_launch("epb_ingest_power_i16");
}
