// sweeps_fused.cu -- temporally blocked Jacobi sweeps (several sweeps per launch).
#include "pfs_internal.cuh"

namespace pfs {

bool fused_sweeps_supported(int, int) { return false; }

int launch_sweeps_fused(SweepOp, float *, float *, float *, float *, const float *, const SweepParams &, int, int,
                        int *, cudaStream_t)
{
    set_error("fused sweeps not built");
    return PFS_EINVAL;
}

}  // namespace pfs
