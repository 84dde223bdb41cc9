// Instantiations of the persistent fused kernel (pipeline_fast_impl.cuh): int16 raw power counts, ping_num > 8 (two sweeps
// over sub-tiles of 5..8 rows).
#include "pipeline_fast_impl.cuh"

EPB_DEFINE_FAST_SWEEP_LAUNCHER(epb_fast_launch_i16c, true, 5, 6, 7, 8)
