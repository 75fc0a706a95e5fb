// fused K-step rollout kernels of Robot (see nig_kernels.cuh)
#include "nig_rollout_launch.cuh"
namespace nig {
cudaError_t launch_rollout_robot(const RolloutLaunch& cfg, int64_t pitch, const RolloutArgs& a, const CUtensorMap& map, cudaStream_t st)
{
    return rollout_env<Robot>(cfg, pitch, a, map, st);
}
} // namespace nig
