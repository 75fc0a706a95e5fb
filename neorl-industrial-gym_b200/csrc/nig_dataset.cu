// on-device get_dataset kernels (see nig_kernels.cuh)
#include "nig_launch.h"
namespace nig {
namespace {
template <class Env, bool WRITE>
cudaError_t go(bool defcons, const DatasetArgs& a, cudaStream_t st)
{
    const unsigned g = grid_for(a.n_episodes);
    if (defcons) dataset_kernel<Env, true, WRITE><<<g, kThreads, 0, st>>>(a);
    else dataset_kernel<Env, false, WRITE><<<g, kThreads, 0, st>>>(a);
    return cudaGetLastError();
}
template <bool WRITE>
cudaError_t by_kind(int kind, bool defcons, const DatasetArgs& a, cudaStream_t st)
{
    switch (kind) {
    case NIG_ENV_CHEMICAL_REACTOR: return go<Reactor, WRITE>(defcons, a, st);
    case NIG_ENV_POWER_GRID: return go<Grid, WRITE>(defcons, a, st);
    default: return go<Robot, WRITE>(defcons, a, st);
    }
}
} // namespace
cudaError_t launch_dataset(int kind, bool defcons, bool write, const DatasetArgs& a, cudaStream_t st)
{
    return write ? by_kind<true>(kind, defcons, a, st) : by_kind<false>(kind, defcons, a, st);
}
cudaError_t launch_scan_lengths(const int64_t* len, int64_t* off, int64_t n, int64_t* total, cudaStream_t st)
{
    scan_lengths_kernel<<<1, 1024, 0, st>>>(len, off, n, total);
    return cudaGetLastError();
}
} // namespace nig
