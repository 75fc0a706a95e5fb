// single-step kernels of the three envs (see nig_kernels.cuh)
#include "nig_launch.h"
namespace nig {
namespace {
template <class Env, int VEC>
cudaError_t go(bool defcons, int64_t pitch, const StepArgs& a, cudaStream_t st)
{
    const unsigned g = grid_for((pitch + VEC - 1) / VEC);
    if (defcons) step_kernel<Env, VEC, true><<<g, kThreads, 0, st>>>(a);
    else step_kernel<Env, VEC, false><<<g, kThreads, 0, st>>>(a);
    return cudaGetLastError();
}
} // namespace
cudaError_t launch_step(int kind, int vec, bool defcons, int64_t pitch, const StepArgs& a, cudaStream_t st)
{
    switch (kind) {
    case NIG_ENV_CHEMICAL_REACTOR:
        return vec == 4 ? go<Reactor, 4>(defcons, pitch, a, st) : vec == 2 ? go<Reactor, 2>(defcons, pitch, a, st) : go<Reactor, 1>(defcons, pitch, a, st);
    case NIG_ENV_POWER_GRID: return vec >= 2 ? go<Grid, 2>(defcons, pitch, a, st) : go<Grid, 1>(defcons, pitch, a, st);
    default: return vec >= 2 ? go<Robot, 2>(defcons, pitch, a, st) : go<Robot, 1>(defcons, pitch, a, st);
    }
}
} // namespace nig
