// Instantiations of the persistent fused kernel (pipeline_fast_impl.cuh) with full-size outputs (keep=): float32 samples,
// ping_num in {1, 2, 3, 4} and the variant without noise removal.
#include "pipeline_fast_impl.cuh"

EPB_DEFINE_FAST_KEEP_LAUNCHER(epb_fast_launch_f32ka, false, 1, 2, 3, 4, true)
