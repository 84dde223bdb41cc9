// Instantiations of the persistent fused kernel (pipeline_fast_impl.cuh) for rows of 4097 .. 8192 float32 samples: four
// column groups per thread (512 threads), single-row tiles; with the noise estimate the rows of a noise tile stream
// through the ring twice (kSweep).
#include "pipeline_fast_impl.cuh"

int epb_fast_launch_f32w(const void* prv, int T, int G, int noise, int threads, size_t smem, cudaStream_t s) {
  const FastParams& pr = *static_cast<const FastParams*>(prv);
  if (T != 1 || G != 4) return -2;
  return noise ? launch_fast<1, 4, true, false, true>(pr, threads, smem, s) : launch_fast<1, 4, false, false>(pr, threads, smem, s);
}
