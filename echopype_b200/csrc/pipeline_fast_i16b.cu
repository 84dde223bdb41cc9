// Instantiations of the persistent fused kernel (pipeline_fast_impl.cuh): int16 raw counts, ping_num in {5, 6, 7, 8}.
#include "pipeline_fast_impl.cuh"

EPB_DEFINE_FAST_LAUNCHER(epb_fast_launch_i16b, true, 5, 6, 7, 8, false)
