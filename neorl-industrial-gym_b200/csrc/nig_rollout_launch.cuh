// nig_rollout_launch.cuh -- instantiates the fused rollout kernel for one env kind (included by one .cu per env).
#pragma once
#include "nig_launch.h"

namespace nig {

template <class Env, int CONS, int POLICY, bool TMA, bool TFNOISE, bool EXTREMA = false>
cudaError_t rollout_go(const RolloutLaunch& cfg, int64_t pitch, const RolloutArgs& a, const CUtensorMap& map, cudaStream_t st)
{
    auto kern = rollout_kernel<Env, CONS, POLICY, TMA, TFNOISE, EXTREMA>;
    const int block = TMA ? kThreads : cfg.block;
    const size_t smem = TMA ? (size_t)2 * tma_chunk(Env::A) * Env::A * kThreads * sizeof(float) : 0;
    if (TMA) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return launch_pdl(kern, grid_for(pitch, block), (unsigned)block, smem, st, a, map);
}

template <class Env, int CONS>
cudaError_t rollout_policy(const RolloutLaunch& c, int64_t pitch, const RolloutArgs& a, const CUtensorMap& map, cudaStream_t st)
{
    switch (c.policy) {
    case NIG_POLICY_ACTIONS:
        if (c.tma) {
            if constexpr (Env::NZ > 0) { if (c.tf_noise) return rollout_go<Env, CONS, NIG_POLICY_ACTIONS, true, true>(c, pitch, a, map, st); }
            return rollout_go<Env, CONS, NIG_POLICY_ACTIONS, true, false>(c, pitch, a, map, st);
        }
        if constexpr (Env::NZ > 0) { if (c.tf_noise) return rollout_go<Env, CONS, NIG_POLICY_ACTIONS, false, true>(c, pitch, a, map, st); }
        return rollout_go<Env, CONS, NIG_POLICY_ACTIONS, false, false>(c, pitch, a, map, st);
    case NIG_POLICY_UNIFORM: return rollout_go<Env, CONS, NIG_POLICY_UNIFORM, false, false>(c, pitch, a, map, st);
    case NIG_POLICY_ZERO: return rollout_go<Env, CONS, NIG_POLICY_ZERO, false, false>(c, pitch, a, map, st);
    case NIG_POLICY_PCTRL: return rollout_go<Env, CONS, NIG_POLICY_PCTRL, false, false>(c, pitch, a, map, st);
    case NIG_POLICY_BASELINE: return rollout_go<Env, CONS, NIG_POLICY_BASELINE, false, false>(c, pitch, a, map, st);
    default: return cudaErrorInvalidValue;
    }
}

// NIG_ROLLOUT_EXTREMA: the evaluation flavour (return_min / return_max). Action tensors go through the register-prefetch
// path; teacher-forced noise is a parity-test mode and has no extrema flavour.
template <class Env, int CONS>
cudaError_t rollout_policy_extrema(const RolloutLaunch& c, int64_t pitch, const RolloutArgs& a, const CUtensorMap& map, cudaStream_t st)
{
    switch (c.policy) {
    case NIG_POLICY_ACTIONS: return rollout_go<Env, CONS, NIG_POLICY_ACTIONS, false, false, true>(c, pitch, a, map, st);
    case NIG_POLICY_UNIFORM: return rollout_go<Env, CONS, NIG_POLICY_UNIFORM, false, false, true>(c, pitch, a, map, st);
    case NIG_POLICY_ZERO: return rollout_go<Env, CONS, NIG_POLICY_ZERO, false, false, true>(c, pitch, a, map, st);
    case NIG_POLICY_PCTRL: return rollout_go<Env, CONS, NIG_POLICY_PCTRL, false, false, true>(c, pitch, a, map, st);
    case NIG_POLICY_BASELINE: return rollout_go<Env, CONS, NIG_POLICY_BASELINE, false, false, true>(c, pitch, a, map, st);
    default: return cudaErrorInvalidValue;
    }
}

// CONS_BOUNDS1 / CONS_BOUNDS2 (built-ins + one / two state bounds as straight-line code): the throughput flavours only
// (in-kernel policies and register-prefetched action tensors, with or without return extrema -- the evaluation of a wrapped
// env); TMA staging and teacher-forced noise take the CONS_PREFIX kernels for such a constraint set.
template <class Env, int CONS, bool EXTREMA>
cudaError_t rollout_policy_bounds(const RolloutLaunch& c, int64_t pitch, const RolloutArgs& a, const CUtensorMap& map, cudaStream_t st)
{
    switch (c.policy) {
    case NIG_POLICY_ACTIONS: return rollout_go<Env, CONS, NIG_POLICY_ACTIONS, false, false, EXTREMA>(c, pitch, a, map, st);
    case NIG_POLICY_UNIFORM: return rollout_go<Env, CONS, NIG_POLICY_UNIFORM, false, false, EXTREMA>(c, pitch, a, map, st);
    case NIG_POLICY_ZERO: return rollout_go<Env, CONS, NIG_POLICY_ZERO, false, false, EXTREMA>(c, pitch, a, map, st);
    case NIG_POLICY_PCTRL: return rollout_go<Env, CONS, NIG_POLICY_PCTRL, false, false, EXTREMA>(c, pitch, a, map, st);
    case NIG_POLICY_BASELINE: return rollout_go<Env, CONS, NIG_POLICY_BASELINE, false, false, EXTREMA>(c, pitch, a, map, st);
    default: return cudaErrorInvalidValue;
    }
}

template <class Env>
cudaError_t rollout_env(const RolloutLaunch& c0, int64_t pitch, const RolloutArgs& a, const CUtensorMap& map, cudaStream_t st)
{
    RolloutLaunch c = c0;
    if (c.cons > CONS_PREFIX) {
        if (!c.tma && !c.tf_noise) {
            if (c.extrema)
                return c.cons == CONS_BOUNDS1 ? rollout_policy_bounds<Env, CONS_BOUNDS1, true>(c, pitch, a, map, st)
                                              : rollout_policy_bounds<Env, CONS_BOUNDS2, true>(c, pitch, a, map, st);
            return c.cons == CONS_BOUNDS1 ? rollout_policy_bounds<Env, CONS_BOUNDS1, false>(c, pitch, a, map, st)
                                          : rollout_policy_bounds<Env, CONS_BOUNDS2, false>(c, pitch, a, map, st);
        }
        c.cons = CONS_PREFIX;
    }
    if (c.extrema)
        return c.cons == CONS_DEFAULT ? rollout_policy_extrema<Env, CONS_DEFAULT>(c, pitch, a, map, st)
             : c.cons == CONS_PREFIX  ? rollout_policy_extrema<Env, CONS_PREFIX>(c, pitch, a, map, st)
                                      : rollout_policy_extrema<Env, CONS_GENERIC>(c, pitch, a, map, st);
    return c.cons == CONS_DEFAULT ? rollout_policy<Env, CONS_DEFAULT>(c, pitch, a, map, st)
         : c.cons == CONS_PREFIX  ? rollout_policy<Env, CONS_PREFIX>(c, pitch, a, map, st)
                                  : rollout_policy<Env, CONS_GENERIC>(c, pitch, a, map, st);
}

} // namespace nig
