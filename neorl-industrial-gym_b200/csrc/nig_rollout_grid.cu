// fused K-step rollout kernels of Grid (see nig_kernels.cuh)
#include "nig_rollout_launch.cuh"
namespace nig {
cudaError_t launch_rollout_grid(const RolloutLaunch& cfg, int64_t pitch, const RolloutArgs& a, const CUtensorMap& map, cudaStream_t st)
{
    return rollout_env<Grid>(cfg, pitch, a, map, st);
}
} // namespace nig
