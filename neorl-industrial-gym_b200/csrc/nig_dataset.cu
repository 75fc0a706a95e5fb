// on-device get_dataset kernels (see nig_kernels.cuh)
#include "nig_launch.h"
namespace nig {
namespace {
template <class Env, bool WRITE>
cudaError_t go(int cons, const DatasetArgs& a, cudaStream_t st)
{
    const unsigned g = grid_for(a.n_episodes);
    if (cons == CONS_DEFAULT) dataset_kernel<Env, CONS_DEFAULT, WRITE><<<g, kThreads, 0, st>>>(a);
    else dataset_kernel<Env, CONS_GENERIC, WRITE><<<g, kThreads, 0, st>>>(a);      // prefix sets take the generic path here
    return cudaGetLastError();
}
template <bool WRITE>
cudaError_t by_kind(int kind, int cons, const DatasetArgs& a, cudaStream_t st)
{
    switch (kind) {
    case NIG_ENV_CHEMICAL_REACTOR: return go<Reactor, WRITE>(cons, a, st);
    case NIG_ENV_POWER_GRID: return go<Grid, WRITE>(cons, a, st);
    default: return go<Robot, WRITE>(cons, a, st);
    }
}
} // namespace
cudaError_t launch_dataset(int kind, int cons, bool write, const DatasetArgs& a, cudaStream_t st)
{
    return write ? by_kind<true>(kind, cons, a, st) : by_kind<false>(kind, cons, a, st);
}
cudaError_t launch_scan_lengths(const int64_t* len, int64_t* off, int64_t n, int64_t* total, cudaStream_t st)
{
    scan_lengths_kernel<<<1, 1024, 0, st>>>(len, off, n, total);
    return cudaGetLastError();
}
} // namespace nig
