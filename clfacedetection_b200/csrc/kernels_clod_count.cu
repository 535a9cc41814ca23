// kernels_clod_count.cu -- the diagnostic instantiations of k_cascade_tiles (COUNT = true: FP64 fallbacks and
// near-threshold stage sums are counted, clfd_run_stats), compiled from the same source as the production ones
// in a translation unit of their own so that the two build in parallel.  See launch_cascade_tiles.
#define CLFD_TILES_COUNT_TU 1
#include "kernels_clod.cu"
