// marg_stats_kernel for 1 and 2 items (see marg_stats_impl.cuh)
#include "marg_stats_impl.cuh"

namespace gwsem {
#define GW_INST(JJ)                                                                                                   \
  template void launch_marg_stats_j<JJ>(gwsem_engine*, const MargProgram&, const uint8_t*, int64_t, int, int,        \
                                        const uint8_t*, const int32_t*, int, const double*, int, double*, size_t);
GW_INST(1) GW_INST(2)
#undef GW_INST
}  // namespace gwsem
