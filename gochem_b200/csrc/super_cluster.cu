// super_cluster.cu -- placeholder until the shared-memory-resident cluster kernel lands.
#include "common.cuh"
namespace gcb {
bool cluster_kernel_supported(const gcb_traj*) { return false; }
int launch_super_cluster(gcb_traj*, const SuperParams&, bool, bool, int*) {
    set_error("cluster kernel not built");
    return GCB_ERR_ARG;
}
}  // namespace gcb
