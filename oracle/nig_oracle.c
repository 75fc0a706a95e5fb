/* nig_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, scalar, one-env-at-a-time restatement of the reference's IndustrialEnv step path
 * (danieleschmidt/neoRL-industrial-gym, src/neorl_industrial/environments/{base,chemical_reactor,
 * power_grid,robot_assembly}.py). Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load this library, and only as the checker / the timed CPU baseline. The product
 * (neorl-industrial-gym_b200/) never links, imports or falls back to it.
 *
 * Pinning: the reference's own tests hold NO numeric vectors for this path, so the
 * oracle is pinned against outputs of the UNMODIFIED reference code executed in the build container
 * (tests/golden/make_golden.py -> tests/golden/ npz fixtures, checked by tests/test_oracle_golden.py).
 *
 * Arithmetic convention (= numpy >= 2 / NEP 50 with float32 actions): every
 * reactor intermediate is IEEE binary32, round-to-nearest, NO fused multiply-add, evaluated
 * left-to-right exactly as the Python source parenthesises; Python-float literals are first
 * rounded to binary32. Compile with -ffp-contract=off (the Makefile does).
 *
 * Two things are NOT reference code but this project's documented spec, restated here
 * independently of the CUDA sources so that free-running runs can be compared bit-for-bit:
 *   - the counter-based RNG (Philox4x32-7 -> inverse-CDF normals by a dyadic-segment table + fmaf cubic), and
 *   - spec_expf(), a <1 ulp fmaf-polynomial exp (exp_mode 1). exp_mode 0 uses libm expf and is
 *     what is compared against the reference goldens (numpy's own float32 exp is a SIMD routine
 *     that is not correctly rounded and is host-dependent, so conc' is a tolerance compare there).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

enum { ORC_REACTOR = 0, ORC_GRID = 1, ORC_ROBOT = 2 };
enum { ORC_CON_BUILTIN = 0, ORC_CON_BOUND = 1, ORC_CON_HOSTMASK = 2 };
enum { ORC_F_TERMINATED = 1, ORC_F_TRUNCATED = 2, ORC_F_CRITICAL = 4, ORC_F_RESET = 8, ORC_F_INACTIVE = 128 };
enum { ORC_MAX_S = 32, ORC_MAX_A = 8, ORC_MAX_NZ = 23, ORC_MAX_CONS = 8 };

typedef struct {
    int32_t kind;     /* ORC_CON_* */
    int32_t id;       /* builtin: 0..2; hostmask: bit index */
    int32_t si;       /* bound: state index */
    int32_t ai;       /* bound: action index or -1 */
    float coef;       /* bound: v = s[si] + coef*a[ai] */
    float lo, hi;     /* bound: lo <= v <= hi */
    float penalty;
    int32_t critical;
} orc_con_t;

typedef struct {
    int32_t kind;
    int32_t max_episode_steps;
    int32_t n_cons;
    int32_t exp_mode;   /* 0 libm expf / sin / cos (what the reference calls), 1 the project's math spec (spec_expf, spec_sincos_f64) */
    int32_t auto_reset;
    int32_t pad;
    uint64_t seed;
    orc_con_t cons[ORC_MAX_CONS];
} orc_cfg_t;

static const int kS[3] = {12, 32, 24};
static const int kA[3] = {3, 8, 7};
static const int kNZ[3] = {2, 23, 0};

ORC_API int orc_state_dim(int kind) { return kS[kind]; }
ORC_API int orc_action_dim(int kind) { return kA[kind]; }
ORC_API int orc_noise_dim(int kind) { return kNZ[kind]; }
ORC_API int orc_cfg_size(void) { return (int)sizeof(orc_cfg_t); }

/* ------------------------------------------------------------------------------------------- */
/* Math spec (DESIGN.md "Math spec"): only exactly-rounded IEEE ops (+,*,fmaf,sqrtf,rintf).     */
/* ------------------------------------------------------------------------------------------- */
static inline float bits_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static float spec_expf(float x)
{
    if (x != x) return x + x;
    float xc = x < -104.0f ? -104.0f : x;
    xc = xc > 89.0f ? 89.0f : xc;
    float n = rintf(xc * 0x1.715476p+0f);
    float r = fmaf(n, -0x1.62e4p-1f, xc);
    r = fmaf(n, -0x1.7f7d1cp-20f, r);
    float p = 0x1.a17e08p-13f;
    p = fmaf(p, r, 0x1.6d7548p-10f);
    p = fmaf(p, r, 0x1.1110a6p-7f);
    p = fmaf(p, r, 0x1.5554acp-5f);
    p = fmaf(p, r, 0x1.555556p-3f);
    p = fmaf(p, r, 0x1.0p-1f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int ni = (int)n;
    int n1 = ni >> 1, n2 = ni - n1;
    float s1 = bits_f((uint32_t)(n1 + 127) << 23);
    float s2 = bits_f((uint32_t)(n2 + 127) << 23);
    return (p * s1) * s2;
}

/* Philox4x32-R (Salmon et al., SC'11). The spec's streams use R = 7, the fewest rounds the paper reports as passing
 * BigCrush ("Crush-resistant"); R = 10 is kept for the published known-answer vectors. */
#define SPEC_PHILOX_ROUNDS 7
static void philox4x32(int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline float u_open(uint32_t x) { return fmaf((float)x, 0x1.0p-32f, 0x1.0p-33f); }  /* (0,1] */
static inline float u_sym(uint32_t x) { return fmaf((float)x, 0x1.0p-31f, -1.0f); }          /* [-1,1] */

/* One standard normal from one 32-bit word (math spec): inverse CDF by a dyadic-segment table.
 * v = 2 (w mod 2^31) + 1 counts the tail (p = v / 2^33); the binary32 exponent of RN(v) and its top four mantissa
 * bits select the segment, a cubic in f (explicit fmaf, Horner; coefficients pre-scaled by powers of two) gives |z|, bit 31 of w the sign.
 * The table is data of the spec (tools/fit_normal_table.py writes the same numbers for the library and for this file). */
#include "nig_normal_table.h"
static const float normal_tab[NIG_NORMAL_TAB_N][4] = { NIG_NORMAL_TAB_VALUES };

static float spec_normal(uint32_t w)
{
    const uint32_t v = (w << 1) | 1u;
    const float f = (float)v;                                  /* round-to-nearest conversion */
    const float* c = normal_tab[(f_bits(f) >> 19) - 127u * 16u];
    const float z = fmaf(fmaf(fmaf(c[3], f, c[2]), f, c[1]), f, c[0]);
    return bits_f(f_bits(z) ^ (w & 0x80000000u));
}

static void normal_pair(uint32_t xa, uint32_t xb, float* z0, float* z1)
{
    *z0 = spec_normal(xa);
    *z1 = spec_normal(xb);
}

enum { STREAM_NOISE = 0, STREAM_RESET = 1, STREAM_POLICY = 2 };

/* 4 standard normals from block j of (env, tick, stream) */
static void normals4(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, uint32_t stream, uint32_t j, float z[4])
{
    uint32_t w[4];
    philox4x32(SPEC_PHILOX_ROUNDS, env, tick, stream, j, (uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32), w);
    normal_pair(w[0], w[1], &z[0], &z[1]);
    normal_pair(w[2], w[3], &z[2], &z[3]);
}
static void words4(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, uint32_t stream, uint32_t j, uint32_t w[4])
{
    philox4x32(SPEC_PHILOX_ROUNDS, env, tick, stream, j, (uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32), w);
}

ORC_API void orc_spec_normals4(uint64_t seed, uint32_t env, uint32_t tick, uint32_t stream, uint32_t j, float* z)
{
    orc_cfg_t c; memset(&c, 0, sizeof c); c.seed = seed; normals4(&c, env, tick, stream, j, z);
}
ORC_API float orc_spec_expf(float x) { return spec_expf(x); }
ORC_API float orc_spec_normal(uint32_t w) { return spec_normal(w); }
/* the two checksums of nig_selftest_normal (include/nig_b200.h), computed on the CPU */
ORC_API void orc_selftest_normal(uint32_t first, uint32_t stride, int64_t count, uint64_t* sums2)
{
    uint64_t s0 = 0, s1 = 0;
#pragma omp parallel for reduction(+ : s0, s1) schedule(static)
    for (int64_t k = 0; k < count; ++k) {
        const uint64_t bits = f_bits(spec_normal(first + (uint32_t)k * stride));
        s0 += bits;
        s1 += bits * ((uint64_t)k + 1u);
    }
    sums2[0] = s0; sums2[1] = s1;
}
ORC_API void orc_philox(int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out)
{ philox4x32(rounds, c0, c1, c2, c3, k0, k1, out); }
ORC_API int orc_spec_philox_rounds(void) { return SPEC_PHILOX_ROUNDS; }
/* the spec's words for a rectangle of (env, tick) counters: out[(e * n_tick + t) * 4 + q] (statistical tests) */
ORC_API void orc_words_batch(uint64_t seed, uint32_t env0, int32_t n_env, uint32_t tick0, int32_t n_tick, uint32_t stream, uint32_t j, uint32_t* out)
{
    for (int32_t e = 0; e < n_env; ++e)
        for (int32_t t = 0; t < n_tick; ++t)
            philox4x32(SPEC_PHILOX_ROUNDS, env0 + (uint32_t)e, tick0 + (uint32_t)t, stream, j, (uint32_t)seed, (uint32_t)(seed >> 32),
                       out + ((size_t)e * (size_t)n_tick + (size_t)t) * 4);
}

/* Python's max(lo, min(hi, v)) on a float: min(hi,v) = v if v < hi else hi; max(lo,m) = m if m > lo else lo */
static inline float py_clamp(float v, float lo, float hi)
{
    float m = (v < hi) ? v : hi;
    return (m > lo) ? m : lo;
}

/* ------------------------------------------------------------------------------------------- */
/* ChemicalReactor-v0 (chemical_reactor.py)                                                     */
/* ------------------------------------------------------------------------------------------- */

/* _dynamics, chemical_reactor.py:109-226. noise[0] = the value np.random.normal(0, 0.1) returned
 * (:149), noise[1] = np.random.normal(0, 500) (:159). */
static void reactor_dynamics(const float* s, const float* a, const float* noise, float* o, int exp_mode)
{
    const float temp = s[0], pressure = s[1], cooling_flow = s[2], feed_flow = s[3];
    const float conc = s[4], cat = s[5], hx = s[6], relief = s[7], estop = s[8], alarm = s[9];
    const float level = s[10], batch_time = s[11];

    float heating_power, cooling_adj, feed_adj;
    if (estop < 0.5f) {                       /* :126-129 */
        heating_power = a[0] * 50000.0f;
        cooling_adj = a[1] * 0.1f;
        feed_adj = a[2] * 0.1f;
    } else {                                  /* :130-134 */
        heating_power = -10000.0f;
        cooling_adj = 0.1f;
        feed_adj = -0.1f;
    }
    const float activity = cat / 100.0f;
    const float kc = 0.1f * conc;             /* reaction_rate_constant * concentration */
    const float reaction_heat = (kc * activity) * 10000.0f;            /* :137-140 */
    const float cooling_heat = ((cooling_flow * 100.0f) * (temp - hx)) * 0.1f;  /* :141 */
    float dT = ((heating_power + reaction_heat) - cooling_heat) / 418000.0f;    /* :143-146 */
    dT = dT + noise[0];                                                         /* :149 */
    const float new_temp = temp + dT * 0.1f;                                    /* :151 */

    const float p_from_temp = pressure * (new_temp / temp);                     /* :155 */
    const float p_from_reaction = (conc * 0.1f) * 1000.0f;                      /* :156 */
    float new_p = p_from_temp + p_from_reaction * 0.1f;                         /* :158 */
    new_p = new_p + noise[1];                                                   /* :159 */

    const float new_relief = py_clamp(relief + (new_p - 506625.0f) * 0.001f, 0.0f, 100.0f); /* :162-163 */
    if (new_relief > 0.0f) {                                                    /* :166-168 */
        const float pr = (new_relief * 0.01f) * 10000.0f;
        const float x = new_p - pr;
        new_p = (x > 101325.0f) ? x : 101325.0f;
    }
    const float cool_v = cooling_flow + cooling_adj;                            /* :171 */
    const float new_cool = py_clamp(cool_v, 10.0f, 100.0f);
    const float feed_v = feed_flow + feed_adj;                                  /* :172 */
    const float new_feed = py_clamp(feed_v, 5.0f, 50.0f);

    const float earg = (-(new_temp - 320.0f)) / 20.0f;                          /* :177 */
    const float ex = exp_mode ? spec_expf(earg) : expf(earg);
    const float reaction_rate = (kc * activity) * ex;                           /* :175-178 */
    /* :180 -- when the clamp returned the Python int 5 the product 5*0.001 is a Python float
     * (0.005) that is then rounded to binary32, which is NOT float32(5)*float32(0.001). */
    float feed_dilution;
    if (!(feed_v < 50.0f)) feed_dilution = 0x1.99999ap-5f;        /* float32(0.05)  */
    else if (!(feed_v > 5.0f)) feed_dilution = 0x1.47ae14p-8f;    /* float32(0.005) */
    else feed_dilution = new_feed * 0.001f;
    const float cv = conc + (reaction_rate - feed_dilution) * 0.1f;             /* :181-182 */
    const float new_conc = (cv > 0.0f) ? cv : 0.0f;

    const float deact = (new_temp > 340.0f) ? 0.001f : 0.0001f;                 /* :185 */
    const float catv = cat - deact;                                             /* :186 */
    const float new_cat = (catv > 50.0f) ? catv : 50.0f;

    const float new_hx = hx + (0.1f * ((290.0f + cooling_flow * 0.1f) - hx)) * 0.1f; /* :189-190 */

    float new_estop = estop, new_alarm = alarm;                                 /* :193-201 */
    if (new_temp > 345.0f || new_p > 480000.0f) new_alarm = 1.0f;
    if (new_temp > 350.0f || new_p > 506625.0f) { new_estop = 1.0f; new_alarm = 1.0f; }

    const float level_change = (new_feed - 20.0f) * 0.1f;                       /* :204 */
    const float new_level = py_clamp(level + level_change * 0.1f, 0.0f, 100.0f); /* :205 */
    const float new_bt = batch_time + 0.1f;                                     /* :208 */

    o[0] = new_temp; o[1] = new_p; o[2] = new_cool; o[3] = new_feed; o[4] = new_conc; o[5] = new_cat;
    o[6] = new_hx; o[7] = new_relief; o[8] = new_estop; o[9] = new_alarm; o[10] = new_level; o[11] = new_bt;
}

/* _compute_reward, chemical_reactor.py:228-270 (fp32 throughout) */
static float reactor_reward(const float* ns, const float* a)
{
    float r = 0.0f;
    r = r + ns[4] * 100.0f;                                   /* :241 */
    r = r - fabsf(ns[0] - 320.0f) * 0.5f;                     /* :244-245 */
    r = r - (fabsf(ns[1] - 253312.5f) / 1000.0f) * 0.1f;      /* :248-249 */
    r = r + (ns[5] / 100.0f) * 10.0f;                         /* :252 */
    if (30.0f <= ns[10] && ns[10] <= 80.0f) r = r + 5.0f;     /* :255-258 */
    else r = r - fabsf(ns[10] - 55.0f) * 0.2f;
    if (ns[9] > 0.5f) r = r - 50.0f;                          /* :261-262 */
    if (ns[8] > 0.5f) r = r - 200.0f;                         /* :263-264 */
    const float asum = (fabsf(a[0]) + fabsf(a[1])) + fabsf(a[2]);  /* np.sum of 3 fp32: sequential */
    r = r - asum * 0.1f;                                      /* :267-268 */
    return r;
}

/* _is_done, chemical_reactor.py:272-290 */
static int reactor_is_done(const float* s)
{
    if (s[8] > 0.5f) return 1;
    if (s[10] < 5.0f || s[10] > 95.0f) return 1;
    if (s[11] > 50.0f) return 1;
    return 0;
}

/* constraint check_fns, chemical_reactor.py:292-305; returns 1 when SATISFIED */
static int reactor_builtin(int id, const float* s, const float* a)
{
    (void)a;
    switch (id) {
    case 0: return s[0] <= 350.0f;
    case 1: return s[1] <= 506625.0f;
    default: return (20.0f <= s[10]) && (s[10] <= 90.0f);
    }
}

/* _get_initial_state distribution (chemical_reactor.py:89-107), drawn from OUR spec'd RNG in fp32 */
static void reactor_reset(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, uint32_t epoch, float* s)
{
    float za[4], zb[4];
    normals4(cfg, env, tick, STREAM_RESET, (epoch << 8) | 0u, za);
    normals4(cfg, env, tick, STREAM_RESET, (epoch << 8) | 1u, zb);
    s[0] = 320.0f + 2.0f * za[0];
    s[1] = 253312.5f + 10000.0f * za[1];
    s[2] = 50.0f + 5.0f * za[2];
    s[3] = 30.0f + 3.0f * za[3];
    s[4] = 0.5f + 0.1f * zb[0];
    s[5] = 95.0f + 2.0f * zb[1];
    s[6] = 295.0f + 1.0f * zb[2];
    s[7] = 0.0f; s[8] = 0.0f; s[9] = 0.0f;
    s[10] = 60.0f + 5.0f * zb[3];
    s[11] = 0.0f;
}

/* ONE Philox block per reactor step, (env, tick, NOISE, 0): words 0, 1 -> the two process-noise normals of the step;
 * words 2, 3 -> the uniform-random policy's action_space.sample() (reactor_uniform_actions) */
static void reactor_noise(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, float* nz)
{
    uint32_t w[4];
    words4(cfg, env, tick, STREAM_NOISE, 0u, w);
    nz[0] = 0.1f * spec_normal(w[0]);
    nz[1] = 500.0f * spec_normal(w[1]);
}
/* a0 / a1 = v * 2^-21 - 1 from bits 21..0 of words 2 / 3, a2 = v * 2^-19 - 1 from their two 10-bit tops: the bits are the
 * mantissa of a float in [1, 2) and one exact fma maps it to [-1, 1) */
static void reactor_uniform_actions(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, float* a)
{
    uint32_t w[4];
    words4(cfg, env, tick, STREAM_NOISE, 0u, w);
    a[0] = fmaf(bits_f((w[2] & 0x3fffffu) | 0x3f800000u), 4.0f, -5.0f);
    a[1] = fmaf(bits_f((w[3] & 0x3fffffu) | 0x3f800000u), 4.0f, -5.0f);
    const uint32_t v2 = ((w[2] >> 22) << 10) | (w[3] >> 22);
    a[2] = fmaf(bits_f(v2 | 0x3f800000u), 16.0f, -17.0f);
}

/* ------------------------------------------------------------------------------------------- */
/* PowerGrid-v0 (power_grid.py)                                                                 */
/* ------------------------------------------------------------------------------------------- */
static inline float pairwise8(const float* x)
{   /* numpy pairwise_sum, n == 8: eight accumulators, combined as a tree */
    return ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
}
static inline double pairwise8d(const double* x)
{
    return ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
}
static const float kBaseLoad[8] = {50, 60, 45, 55, 40, 65, 35, 50};      /* power_grid.py:82 */
static const double kGenCost[8] = {25, 30, 28, 35, 32, 27, 40, 33};      /* power_grid.py:88 (int64 -> fp64 math) */

/* _dynamics, power_grid.py:112-153. noise = V(8), load(8), flow(7) as returned by np.random.normal */
static void grid_dynamics(const float* s, const float* a, const float* nz, float* o)
{
    float gen[8];
    for (int i = 0; i < 8; ++i) {            /* :124 np.clip(generation + action, 0, 100) */
        float g = s[9 + i] + a[i];
        g = g < 0.0f ? 0.0f : g;
        g = g > 100.0f ? 100.0f : g;
        gen[i] = g;
    }
    const float total_gen = pairwise8(gen);               /* :127 */
    const float total_load = pairwise8(s + 17);           /* :128 */
    const float imb = total_gen - total_load;             /* :129 */
    const float fd = ((-1.0f * s[0]) + imb) / 5.0f;       /* :132 */
    o[0] = s[0] + fd * 0.1f;                              /* :133 */
    for (int i = 0; i < 8; ++i) o[1 + i] = s[1 + i] + nz[i];           /* :136-137 */
    for (int i = 0; i < 8; ++i) o[9 + i] = gen[i];
    for (int i = 0; i < 8; ++i) {                                       /* :140-141 */
        float l = s[17 + i] + nz[8 + i];
        o[17 + i] = (l > 0.0f) ? l : ((l != l) ? l : 0.0f);             /* np.maximum propagates NaN */
    }
    for (int i = 0; i < 7; ++i) o[25 + i] = s[25 + i] + nz[16 + i];     /* :144 */
}

/* _compute_reward, power_grid.py:155-177 -- returns the Python float (fp64) the reference returns */
static double grid_reward(const float* ns, const float* a)
{
    const float freq_reward = -100.0f * (ns[0] * ns[0]);          /* :162 */
    float dev2[8], a2[8];
    for (int i = 0; i < 8; ++i) { float d = fabsf(ns[1 + i] - 1.0f); dev2[i] = d * d; }  /* :165-166 */
    const float voltage_reward = -50.0f * pairwise8(dev2);
    double cg[8];
    for (int i = 0; i < 8; ++i) cg[i] = kGenCost[i] * (double)ns[9 + i];                  /* :169 */
    const double economic = -pairwise8d(cg) / 1000.0;                                     /* :170 */
    for (int i = 0; i < 8; ++i) a2[i] = a[i] * a[i];
    const float action_pen = -5.0f * pairwise8(a2);                                       /* :173 */
    const float fv = freq_reward + voltage_reward;                                        /* :175 */
    return ((double)fv + economic) + (double)action_pen;
}

static int grid_is_done(const float* s)
{
    if (fabsf(s[0]) > 1.0f) return 1;                          /* :185 */
    for (int i = 0; i < 8; ++i)
        if (s[1 + i] < 0.9f || s[1 + i] > 1.1f) return 1;      /* :189 */
    return 0;
}

static int grid_builtin(int id, const float* s, const float* a)
{
    switch (id) {
    case 0: return fabsf(s[0]) < 0.5f;                         /* power_grid.py:10-14 */
    case 1:                                                    /* :17-21 */
        for (int i = 0; i < 8; ++i)
            if (!(s[1 + i] >= 0.95f && s[1 + i] <= 1.05f)) return 0;
        return 1;
    default:                                                   /* :24-30 */
        for (int i = 0; i < 8; ++i) {
            const float g = s[9 + i] + a[i];
            if (!(g >= 0.0f && g <= 100.0f)) return 0;
        }
        return 1;
    }
}

/* _get_initial_state distribution (power_grid.py:90-110) from OUR spec'd RNG */
static void grid_reset(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, uint32_t epoch, float* s)
{
    float z[8][4];
    uint32_t w[2][4];
    for (uint32_t j = 0; j < 4; ++j) normals4(cfg, env, tick, STREAM_RESET, (epoch << 8) | j, z[j]);
    for (uint32_t j = 0; j < 2; ++j) words4(cfg, env, tick, STREAM_RESET, (epoch << 8) | (4u + j), w[j]);
    for (uint32_t j = 0; j < 2; ++j) normals4(cfg, env, tick, STREAM_RESET, (epoch << 8) | (6u + j), z[4 + j]);
    s[0] = 0.0f;
    for (int i = 0; i < 8; ++i) s[1 + i] = 1.0f + 0.01f * z[i >> 2][i & 3];
    for (int i = 0; i < 8; ++i) s[9 + i] = kBaseLoad[i] + 2.0f * z[2 + (i >> 2)][i & 3];
    for (int i = 0; i < 8; ++i) s[17 + i] = kBaseLoad[i] * (1.0f + 0.2f * u_sym(w[i >> 2][i & 3]));
    for (int i = 0; i < 7; ++i) s[25 + i] = 10.0f * z[4 + (i >> 2)][i & 3];
}

static void grid_noise(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, float* nz)
{
    float z[6][4];
    for (uint32_t j = 0; j < 6; ++j) normals4(cfg, env, tick, STREAM_NOISE, j, z[j]);
    for (int i = 0; i < 8; ++i) nz[i] = 0.005f * z[i >> 2][i & 3];
    for (int i = 0; i < 8; ++i) nz[8 + i] = 1.0f * z[2 + (i >> 2)][i & 3];
    for (int i = 0; i < 7; ++i) nz[16 + i] = 2.0f * z[4 + (i >> 2)][i & 3];
}

/* ------------------------------------------------------------------------------------------- */
/* RobotAssembly-v0 (robot_assembly.py) -- FK / reward in fp64 like the reference               */
/* ------------------------------------------------------------------------------------------- */
static const double kLink[7] = {0.3, 0.3, 0.25, 0.25, 0.15, 0.1, 0.05};  /* :85 */
static const double kTarget[3] = {0.3, 0.0, 0.4};                        /* :90 */
#define ORC_PI 3.141592653589793

/* sin / cos of a joint angle |x| <= pi in binary64, math spec of this project (exp_mode 1; the CUDA kernels implement the
 * same operation sequence): k = rint(x * 2/pi), r = x - k * pi/2 in two fma steps, the classic degree-13 / degree-14
 * minimax kernels on |r| <= pi/4 evaluated with fma in Horner order, quadrant fix-up. < 1 ulp; libm's sin / cos (what numpy
 * calls upstream, exp_mode 0) are not identical across libm versions, so they cannot be a CPU == GPU contract. */
static void spec_sincos_f64(double x, double* sn, double* cs)
{
    const double kf = rint(x * 0x1.45f306dc9c883p-1);             /* 2/pi */
    double r = fma(-kf, 0x1.921fb54442d18p+0, x);                /* pi/2 hi */
    r = fma(-kf, 0x1.1a62633145c07p-54, r);                       /* pi/2 lo */
    const double z = r * r;
    double ps = 0x1.5d93a5acfd57cp-33;                            /* S6 */
    ps = fma(ps, z, -0x1.ae5e68a2b9cebp-26);                      /* S5 */
    ps = fma(ps, z, 0x1.71de357b1fe7dp-19);                       /* S4 */
    ps = fma(ps, z, -0x1.a01a019c161d5p-13);                      /* S3 */
    ps = fma(ps, z, 0x1.111111110f8a6p-7);                        /* S2 */
    ps = fma(ps, z, -0x1.5555555555549p-3);                       /* S1 */
    const double s0 = fma(r * z, ps, r);
    double pc = -0x1.8fae9be8838d4p-37;                           /* C6 */
    pc = fma(pc, z, 0x1.1ee9ebdb4b1c4p-29);                       /* C5 */
    pc = fma(pc, z, -0x1.27e4f809c52adp-22);                      /* C4 */
    pc = fma(pc, z, 0x1.a01a019cb1590p-16);                       /* C3 */
    pc = fma(pc, z, -0x1.6c16c16c15177p-10);                      /* C2 */
    pc = fma(pc, z, 0x1.555555555554cp-5);                        /* C1 */
    const double c0 = fma(z * z, pc, fma(z, -0.5, 1.0));
    const int k = (int)kf & 3;                                    /* two's complement: -1 -> 3, -2 -> 2 */
    const double ss = (k & 1) ? c0 : s0, cc = (k & 1) ? s0 : c0;
    *sn = (k & 2) ? -ss : ss;
    *cs = ((k + 1) & 2) ? -cc : cc;
}

static void robot_fk(const double* q, double* pos, int spec)
{   /* :94-111 */
    double x = 0.0, y = 0.0, z = 0.0;
    for (int i = 0; i < 7; ++i) {
        double sn, cs;
        if (spec) spec_sincos_f64(q[i], &sn, &cs); else { sn = sin(q[i]); cs = cos(q[i]); }
        if ((i & 1) == 0) { x += kLink[i] * cs; z += kLink[i] * sn; }
        else y += kLink[i] * sn;
    }
    pos[0] = x; pos[1] = y; pos[2] = z;
}

static void robot_dynamics(const float* s, const float* a, float* o, int spec)
{   /* :139-188 */
    double q[7], pos[3];
    for (int i = 0; i < 7; ++i) {
        const float qf = s[7 + i] + a[i] * 0.1f;   /* fp32 array math (:148) ...            */
        double qd = (double)qf;                    /* ... promoted by the fp64 limits (:149) */
        qd = qd < -ORC_PI ? -ORC_PI : qd;
        qd = qd > ORC_PI ? ORC_PI : qd;
        q[i] = qd;
    }
    robot_fk(q, pos, spec);
    double vel[3];
    for (int i = 0; i < 3; ++i) vel[i] = (pos[i] - (double)s[i]) / 0.1;   /* :160 */
    const double dx = pos[0] - kTarget[0], dy = pos[1] - kTarget[1], dz = pos[2] - kTarget[2];
    const double dist = sqrt((dx * dx + dy * dy) + dz * dz);              /* :163 */
    double fz = 0.0;
    if (dist < 0.01) {                                                     /* :164-169 */
        double nf = 0.01 - dist; nf = nf > 0.0 ? nf : 0.0; nf = nf * 1000.0;
        fz = -nf;
    }
    const double aerr = sqrt(dx * dx + dy * dy);                           /* :172 */
    double align = 1.0 - aerr / 0.005; align = align > 0.0 ? align : 0.0;  /* :173 */
    double ins = kTarget[2] - pos[2]; ins = ins > 0.0 ? ins : 0.0;         /* :175 */
    double depth = ins / 0.05; depth = depth < 1.0 ? depth : 1.0;          /* :176 */
    const double comp = align * depth;                                     /* :178 */
    for (int i = 0; i < 24; ++i) o[i] = s[i];
    o[0] = (float)pos[0]; o[1] = (float)pos[1]; o[2] = (float)pos[2];
    o[3] = 0.0f; o[4] = 0.0f; o[5] = 0.0f; o[6] = 1.0f;
    for (int i = 0; i < 7; ++i) o[7 + i] = (float)q[i];
    o[14] = (float)vel[0]; o[15] = (float)vel[1]; o[16] = (float)vel[2]; o[17] = 0.0f;
    o[18] = 0.0f; o[19] = 0.0f; o[20] = (float)fz;
    o[21] = (float)align; o[22] = (float)depth; o[23] = (float)comp;
}

static double robot_reward(const float* ns, const float* a)
{   /* :190-222; slices of the fp32 next_state, scalar math promoted to fp64 by the fp64 target */
    const double completion = (double)(100.0f * ns[23]);        /* int * np.float32 stays fp32 (:197) */
    const double dx = (double)ns[0] - kTarget[0], dy = (double)ns[1] - kTarget[1], dz = (double)ns[2] - kTarget[2];
    const double dist = sqrt((dx * dx + dy * dy) + dz * dz);
    const double distance_reward = -10.0 * dist;
    /* np.linalg.norm of an fp32 slice stays fp32 */
    const float fm = sqrtf((ns[18] * ns[18] + ns[19] * ns[19]) + ns[20] * ns[20]);
    double force_reward = 0.0;
    if (fm > 30.0f) force_reward = (double)(-50.0f * (fm - 30.0f));
    float sa = 0.0f;
    for (int i = 0; i < 7; ++i) sa = sa + a[i] * a[i];                   /* np.sum, n<8: sequential */
    const float action_pen = -0.1f * sa;
    float sv = 0.0f;
    for (int i = 0; i < 4; ++i) sv = sv + ns[14 + i] * ns[14 + i];
    const float vel_pen = -0.5f * sv;
    return (((completion + distance_reward) + force_reward) + (double)action_pen) + (double)vel_pen;
}

static int robot_is_done(const float* s)
{   /* :224-244 */
    if (s[23] > 0.95f) return 1;
    for (int i = 0; i < 3; ++i) if (fabsf(s[18 + i]) > 80.0f) return 1;
    const double lo[3] = {-0.6, -0.6, -0.1}, hi[3] = {0.6, 0.6, 0.9};
    for (int i = 0; i < 3; ++i) if (!((double)s[i] >= lo[i] && (double)s[i] <= hi[i])) return 1;
    return 0;
}

static int robot_builtin(int id, const float* s, const float* a)
{
    (void)a;
    switch (id) {
    case 0: for (int i = 0; i < 3; ++i) if (!(fabsf(s[18 + i]) < 50.0f)) return 0; return 1;   /* :10-15 */
    case 1: {                                                                                  /* :18-25 */
        const double lo[3] = {-0.5, -0.5, 0.0}, hi[3] = {0.5, 0.5, 0.8};
        for (int i = 0; i < 3; ++i) if (!((double)s[i] >= lo[i] && (double)s[i] <= hi[i])) return 0;
        return 1;
    }
    default: for (int i = 0; i < 7; ++i) if (!(fabsf(s[7 + i]) < 2.0f)) return 0; return 1;    /* :28-32 */
    }
}

static void robot_reset(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, uint32_t epoch, float* s)
{   /* :113-137 with q ~ U(-pi/2, pi/2) from OUR spec'd RNG */
    uint32_t w[2][4];
    double q[7], pos[3];
    for (uint32_t j = 0; j < 2; ++j) words4(cfg, env, tick, STREAM_RESET, (epoch << 8) | j, w[j]);
    for (int i = 0; i < 7; ++i) q[i] = (double)(0x1.921fb6p+0f * u_sym(w[i >> 2][i & 3]));
    robot_fk(q, pos, cfg->exp_mode);
    for (int i = 0; i < 24; ++i) s[i] = 0.0f;
    s[0] = (float)pos[0]; s[1] = (float)pos[1]; s[2] = (float)pos[2];
    s[6] = 1.0f;
    for (int i = 0; i < 7; ++i) s[7 + i] = (float)q[i];
}

/* ------------------------------------------------------------------------------------------- */
/* IndustrialEnv.step (environments/base.py:157-213) for ONE env                                */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    float reward;
    uint8_t flags;
    uint8_t viol_mask;
    int32_t n_viol;
    int32_t n_crit;
} orc_out_t;

static int eval_con(const orc_cfg_t* cfg, const orc_con_t* c, const float* s, const float* a, uint8_t hostmask)
{
    switch (c->kind) {
    case ORC_CON_BUILTIN:
        if (cfg->kind == ORC_REACTOR) return reactor_builtin(c->id, s, a);
        if (cfg->kind == ORC_GRID) return grid_builtin(c->id, s, a);
        return robot_builtin(c->id, s, a);
    case ORC_CON_BOUND: {
        float v = s[c->si];
        if (c->ai >= 0) v = v + c->coef * a[c->ai];
        return (c->lo <= v) && (v <= c->hi);
    }
    default: return !((hostmask >> c->id) & 1u);   /* bit set = violated (evaluated by the caller) */
    }
}

static void step_one(const orc_cfg_t* cfg, float* state, int32_t* ep_step, int32_t* ep_viol,
                     const float* action_in, const float* noise, uint8_t hostmask,
                     float* next_state, orc_out_t* out)
{
    const int A = kA[cfg->kind];
    float a[ORC_MAX_A];
    for (int i = 0; i < A; ++i) {            /* base.py:167 np.clip(action, -1, 1) */
        float v = action_in[i];
        v = v < -1.0f ? -1.0f : v;
        v = v > 1.0f ? 1.0f : v;
        a[i] = v;
    }
    /* base.py:170 + :179-183 -- constraints on the PRE-step state (check_fn is evaluated twice in
     * the reference with identical arguments; once here) */
    uint8_t vm = 0; int nv = 0, nc = 0;
    for (int k = 0; k < cfg->n_cons; ++k) {
        if (!eval_con(cfg, &cfg->cons[k], state, a, hostmask)) {
            vm |= (uint8_t)(1u << k); nv++; if (cfg->cons[k].critical) nc++;
        }
    }
    double reward64 = 0.0; float reward32 = 0.0f;
    if (cfg->kind == ORC_REACTOR) {
        reactor_dynamics(state, a, noise, next_state, cfg->exp_mode);      /* base.py:173 */
        reward32 = reactor_reward(next_state, a);                            /* base.py:176 */
        for (int k = 0; k < cfg->n_cons; ++k)                                /* base.py:179-183 */
            if ((vm >> k) & 1u) reward32 = reward32 + cfg->cons[k].penalty;
    } else {
        if (cfg->kind == ORC_GRID) { grid_dynamics(state, a, noise, next_state); reward64 = grid_reward(next_state, a); }
        else { robot_dynamics(state, a, next_state, cfg->exp_mode); reward64 = robot_reward(next_state, a); }
        for (int k = 0; k < cfg->n_cons; ++k)
            if ((vm >> k) & 1u) reward64 = reward64 + (double)cfg->cons[k].penalty;
    }
    *ep_viol += nv;
    *ep_step += 1;                                                           /* base.py:187 */
    int terminated;
    if (cfg->kind == ORC_REACTOR) terminated = reactor_is_done(next_state);  /* base.py:190 */
    else if (cfg->kind == ORC_GRID) terminated = grid_is_done(next_state);
    else terminated = robot_is_done(next_state);
    const int truncated = *ep_step >= cfg->max_episode_steps;                /* base.py:191 */
    uint8_t flags = 0;
    if (nc > 0) {                                                            /* base.py:195-198 */
        terminated = 1; flags |= ORC_F_CRITICAL;
        if (cfg->kind == ORC_REACTOR) reward32 = reward32 - 1000.0f; else reward64 = reward64 - 1000.0;
    }
    if (terminated) flags |= ORC_F_TERMINATED;
    if (truncated) flags |= ORC_F_TRUNCATED;
    out->reward = (cfg->kind == ORC_REACTOR) ? reward32 : (float)reward64;
    out->flags = flags; out->viol_mask = vm; out->n_viol = nv; out->n_crit = nc;
}

static void reset_one(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, uint32_t epoch, float* s)
{
    if (cfg->kind == ORC_REACTOR) reactor_reset(cfg, env, tick, epoch, s);
    else if (cfg->kind == ORC_GRID) grid_reset(cfg, env, tick, epoch, s);
    else robot_reset(cfg, env, tick, epoch, s);
}

static void noise_one(const orc_cfg_t* cfg, uint32_t env, uint32_t tick, float* nz)
{
    if (cfg->kind == ORC_REACTOR) reactor_noise(cfg, env, tick, nz);
    else if (cfg->kind == ORC_GRID) grid_noise(cfg, env, tick, nz);
}

/* stats slots (int64) */
enum { ST_STEPS = 0, ST_EPISODES = 1, ST_TERMINATED = 2, ST_TRUNCATED = 3, ST_CRITICAL = 4, ST_VIOL = 5, ST_CON0 = 8 };

/* Batched step over AoS arrays. state[n][S], actions[n][A], noise[n][NZ] (NULL = spec'd Philox noise),
 * reset_states[n][S] (NULL = spec'd Philox reset draw), hostmask[n] or NULL. Outputs may be NULL. */
ORC_API void orc_step_batch(const orc_cfg_t* cfg, int64_t n, int64_t env_id0, uint32_t tick, uint32_t epoch,
                            float* state, int32_t* ep_step, int32_t* ep_viol, uint8_t* done_latch,
                            const float* actions, const float* noise, const float* reset_states,
                            const uint8_t* hostmask,
                            float* next_obs, float* reward, uint8_t* flags, uint8_t* viol_mask, int64_t* stats,
                            int n_threads)
{
    const int S = kS[cfg->kind], A = kA[cfg->kind], NZ = kNZ[cfg->kind];
    int64_t st[16]; memset(st, 0, sizeof st);
#if defined(_OPENMP)
    if (n_threads <= 0) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
#endif
    {
        int64_t lst[16]; memset(lst, 0, sizeof lst);
#if defined(_OPENMP)
#pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < n; ++i) {
            float* s = state + i * S;
            if (done_latch[i]) {   /* finished env in a no-auto-reset batch: no-op */
                if (reward) reward[i] = 0.0f;
                if (flags) flags[i] = ORC_F_INACTIVE;
                if (viol_mask) viol_mask[i] = 0;
                if (next_obs) memcpy(next_obs + i * S, s, sizeof(float) * S);
                continue;
            }
            float nzbuf[ORC_MAX_NZ];
            const float* nz = NULL;
            if (NZ > 0) {
                if (noise) nz = noise + i * NZ;
                else { noise_one(cfg, (uint32_t)(env_id0 + i), tick, nzbuf); nz = nzbuf; }
            }
            float ns[ORC_MAX_S];
            orc_out_t o;
            step_one(cfg, s, &ep_step[i], &ep_viol[i], actions + i * A, nz, hostmask ? hostmask[i] : 0, ns, &o);
            if (next_obs) memcpy(next_obs + i * S, ns, sizeof(float) * S);
            const int done = (o.flags & (ORC_F_TERMINATED | ORC_F_TRUNCATED)) != 0;
            lst[ST_STEPS]++;
            lst[ST_VIOL] += o.n_viol;
            for (int k = 0; k < cfg->n_cons; ++k) lst[ST_CON0 + k] += (o.viol_mask >> k) & 1;
            if (o.flags & ORC_F_CRITICAL) lst[ST_CRITICAL]++;
            if (done) {
                lst[ST_EPISODES]++;
                if (o.flags & ORC_F_TERMINATED) lst[ST_TERMINATED]++;
                if (o.flags & ORC_F_TRUNCATED) lst[ST_TRUNCATED]++;
                if (cfg->auto_reset) {
                    if (reset_states) memcpy(s, reset_states + i * S, sizeof(float) * S);
                    else reset_one(cfg, (uint32_t)(env_id0 + i), tick + 1u, epoch, s);
                    ep_step[i] = 0; ep_viol[i] = 0;
                    o.flags |= ORC_F_RESET;
                } else {
                    memcpy(s, ns, sizeof(float) * S);
                    done_latch[i] = 1;
                }
            } else {
                memcpy(s, ns, sizeof(float) * S);
            }
            if (reward) reward[i] = o.reward;
            if (flags) flags[i] = o.flags;
            if (viol_mask) viol_mask[i] = o.viol_mask;
        }
#if defined(_OPENMP)
#pragma omp critical
#endif
        for (int k = 0; k < 16; ++k) st[k] += lst[k];
    }
    if (stats) for (int k = 0; k < 16; ++k) stats[k] += st[k];
}

/* reset envs where mask != 0 (mask NULL = all) with the spec'd draw, or copy init_states */
ORC_API void orc_reset_batch(const orc_cfg_t* cfg, int64_t n, int64_t env_id0, uint32_t tick, uint32_t epoch,
                             float* state, int32_t* ep_step, int32_t* ep_viol, uint8_t* done_latch,
                             const uint8_t* mask, const float* init_states)
{
    const int S = kS[cfg->kind];
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        if (init_states) memcpy(state + i * S, init_states + i * S, sizeof(float) * S);
        else reset_one(cfg, (uint32_t)(env_id0 + i), tick, epoch, state + i * S);
        ep_step[i] = 0; ep_viol[i] = 0; done_latch[i] = 0;
    }
}


/* ------------------------------------------------------------------------------------------- */
/* In-kernel policies of the fused rollout / dataset kernels (project spec, DESIGN.md "Policies")  */
/* ------------------------------------------------------------------------------------------- */
enum { ORC_POLICY_ACTIONS = 0, ORC_POLICY_UNIFORM = 1, ORC_POLICY_ZERO = 2, ORC_POLICY_PCTRL = 3 };
typedef struct {
    float p_ctrl, uniform_scale, store_clip;
    int32_t mode;
    float gain[8][2];
    float sigma[8];
} orc_pp_t;

/* the controller branch for given random inputs: z = 8 standard normals, u = 4 uniforms in [-1, 1] */
static void policy_ctrl_from(const orc_cfg_t* cfg, const orc_pp_t* pp, const float* s, const float* z, const float* u, float* a)
{
    if (cfg->kind == ORC_REACTOR) {                 /* chemical_reactor.py:366-385 */
        const float te = (s[0] - 320.0f) / 50.0f, le = (s[10] - 55.0f) / 50.0f;
        for (int k = 0; k < 3; ++k) a[k] = (pp->gain[k][0] * te + pp->gain[k][1] * le) + pp->sigma[k] * z[k];
    } else if (cfg->kind == ORC_GRID) {             /* power_grid.py:216-229 */
        const float imb8 = (pairwise8(s + 17) - pairwise8(s + 9)) / 8.0f;
        for (int k = 0; k < 8; ++k) a[k] = (pp->gain[k][0] * s[0] + pp->gain[k][1] * imb8) + pp->sigma[k] * z[k];
    } else {                                        /* robot_assembly.py:266-287: float64 target_position - float32 obs */
        const double kp = (double)pp->gain[0][0];
        a[0] = (float)(kp * (0.3 - (double)s[0]));
        a[1] = (float)(kp * (0.0 - (double)s[1]));
        a[2] = (float)(kp * (0.4 - (double)s[2]));
        for (int k = 0; k < 4; ++k) a[3 + k] = pp->mode == 0 ? pp->gain[3][0] * s[10 + k] : pp->sigma[3] * u[k];
    }
}

static void policy_one(const orc_cfg_t* cfg, int policy, const orc_pp_t* pp, uint32_t env, uint32_t tick, const float* s, float* a)
{
    const int A = kA[cfg->kind];
    uint32_t w[4];
    if (policy == ORC_POLICY_ZERO) { for (int k = 0; k < A; ++k) a[k] = 0.0f; return; }
    if (policy == ORC_POLICY_UNIFORM) {
        if (cfg->kind == ORC_REACTOR) { reactor_uniform_actions(cfg, env, tick, a); return; }
        for (int j = 0; j < (A + 3) / 4; ++j) {
            words4(cfg, env, tick, STREAM_POLICY, (uint32_t)j, w);
            for (int q = 0; q < 4; ++q) if (4 * j + q < A) a[4 * j + q] = u_sym(w[q]);
        }
        return;
    }
    /* PCTRL: chemical_reactor.py:364-390, power_grid.py:216-232, robot_assembly.py:266-291 */
    uint32_t w0[4];
    words4(cfg, env, tick, STREAM_POLICY, 0u, w0);
    const float coin = u_open(w0[0]);
    if (coin <= pp->p_ctrl) {
        float z[8] = {0}, u[4] = {0};
        if (cfg->kind == ORC_REACTOR) normals4(cfg, env, tick, STREAM_POLICY, 1u, z);
        else if (cfg->kind == ORC_GRID) {
            normals4(cfg, env, tick, STREAM_POLICY, 1u, z);
            normals4(cfg, env, tick, STREAM_POLICY, 2u, z + 4);
        } else {
            words4(cfg, env, tick, STREAM_POLICY, 3u, w);
            for (int k = 0; k < 4; ++k) u[k] = u_sym(w[k]);
        }
        policy_ctrl_from(cfg, pp, s, z, u, a);
    } else {
        for (int k = 0; k < A && k < 3; ++k) a[k] = pp->uniform_scale * u_sym(w0[1 + k]);
        for (int j = 0; 3 + 4 * j < A; ++j) {
            words4(cfg, env, tick, STREAM_POLICY, (uint32_t)(8 + j), w);
            for (int q = 0; q < 4; ++q) if (3 + 4 * j + q < A) a[3 + 4 * j + q] = pp->uniform_scale * u_sym(w[q]);
        }
    }
}

/* get_dataset's policy on caller-supplied random inputs (teacher-forced replay of the reference's own transitions):
 * coin [n] or NULL (= 0.5), z [n][8] standard normals, u [n][8] uniforms in [-1, 1]; actions [n][A] as stored
 * (clipped to +-store_clip when positive) */
ORC_API void orc_policy_forced(const orc_cfg_t* cfg, const orc_pp_t* pp, int64_t n, const float* state, const float* coin,
                               const float* z, const float* u, float* actions)
{
    const int S = kS[cfg->kind], A = kA[cfg->kind];
    static const float zero8[8] = {0};
    for (int64_t i = 0; i < n; ++i) {
        const float* zi = z ? z + i * 8 : zero8;
        const float* ui = u ? u + i * 8 : zero8;
        float* a = actions + i * A;
        if ((coin ? coin[i] : 0.5f) <= pp->p_ctrl) policy_ctrl_from(cfg, pp, state + i * S, zi, ui, a);
        else for (int k = 0; k < A; ++k) a[k] = pp->uniform_scale * ui[k];
        if (pp->store_clip > 0.0f)
            for (int k = 0; k < A; ++k) {
                float v = a[k];
                v = v < -pp->store_clip ? -pp->store_clip : v;
                v = v > pp->store_clip ? pp->store_clip : v;
                a[k] = v;
            }
    }
}

ORC_API void orc_policy_batch(const orc_cfg_t* cfg, int policy, const orc_pp_t* pp, int64_t n, int64_t env_id0,
                              uint32_t tick, const float* state, float* actions, int n_threads)
{
    const int S = kS[cfg->kind], A = kA[cfg->kind];
    (void)n_threads;
#if defined(_OPENMP)
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int64_t i = 0; i < n; ++i) policy_one(cfg, policy, pp, (uint32_t)(env_id0 + i), tick, state + i * S, actions + i * A);
}


/* T free-running steps with an in-kernel policy for every env, each thread owning a contiguous block of envs
 * for the whole horizon (the CPU analogue of the fused rollout kernel; used as the timed CPU baseline and by the
 * full-size parity test). Semantics identical to T x (orc_policy_batch + orc_step_batch). */
ORC_API void orc_rollout_batch(const orc_cfg_t* cfg, int policy, const orc_pp_t* pp, int64_t n, int64_t env_id0,
                               uint32_t tick0, uint32_t epoch, int32_t n_steps,
                               float* state, int32_t* ep_step, int32_t* ep_viol, uint8_t* done_latch,
                               float* reward_sum, int64_t* stats, int n_threads)
{
    const int S = kS[cfg->kind], NZ = kNZ[cfg->kind];
    int64_t st[16]; memset(st, 0, sizeof st);
#if defined(_OPENMP)
    if (n_threads <= 0) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
#endif
    {
        int64_t lst[16]; memset(lst, 0, sizeof lst);
#if defined(_OPENMP)
#pragma omp for schedule(static)
#endif
        for (int64_t i = 0; i < n; ++i) {
            float* s = state + i * S;
            const uint32_t env = (uint32_t)(env_id0 + i);
            float rsum = 0.0f;
            for (int32_t t = 0; t < n_steps && !done_latch[i]; ++t) {
                const uint32_t tick = tick0 + (uint32_t)t;
                float a[ORC_MAX_A], nz[ORC_MAX_NZ], ns[ORC_MAX_S];
                policy_one(cfg, policy, pp, env, tick, s, a);
                if (NZ > 0) noise_one(cfg, env, tick, nz);
                orc_out_t o;
                step_one(cfg, s, &ep_step[i], &ep_viol[i], a, nz, 0, ns, &o);
                rsum = rsum + o.reward;
                lst[ST_STEPS]++; lst[ST_VIOL] += o.n_viol;
                for (int k = 0; k < cfg->n_cons; ++k) lst[ST_CON0 + k] += (o.viol_mask >> k) & 1;
                if (o.flags & ORC_F_CRITICAL) lst[ST_CRITICAL]++;
                if (o.flags & (ORC_F_TERMINATED | ORC_F_TRUNCATED)) {
                    lst[ST_EPISODES]++;
                    if (o.flags & ORC_F_TERMINATED) lst[ST_TERMINATED]++;
                    if (o.flags & ORC_F_TRUNCATED) lst[ST_TRUNCATED]++;
                    if (cfg->auto_reset) { reset_one(cfg, env, tick + 1u, epoch, s); ep_step[i] = 0; ep_viol[i] = 0; }
                    else { memcpy(s, ns, sizeof(float) * S); done_latch[i] = 1; }
                } else memcpy(s, ns, sizeof(float) * S);
            }
            if (reward_sum) reward_sum[i] = rsum;
        }
#if defined(_OPENMP)
#pragma omp critical
#endif
        for (int k = 0; k < 16; ++k) st[k] += lst[k];
    }
    if (stats) for (int k = 0; k < 16; ++k) stats[k] += st[k];
}

/* component functions exposed for direct pinning against the reference's _dynamics/_compute_reward/_is_done */
ORC_API void orc_dynamics(int kind, int exp_mode, int64_t n, const float* s, const float* a, const float* nz, float* o)
{
    const int S = kS[kind], A = kA[kind], NZ = kNZ[kind];
    for (int64_t i = 0; i < n; ++i) {
        if (kind == ORC_REACTOR) reactor_dynamics(s + i * S, a + i * A, nz + i * NZ, o + i * S, exp_mode);
        else if (kind == ORC_GRID) grid_dynamics(s + i * S, a + i * A, nz + i * NZ, o + i * S);
        else robot_dynamics(s + i * S, a + i * A, o + i * S, exp_mode);
    }
}
ORC_API void orc_reward(int kind, int64_t n, const float* ns, const float* a, double* r)
{
    const int S = kS[kind], A = kA[kind];
    for (int64_t i = 0; i < n; ++i) {
        if (kind == ORC_REACTOR) r[i] = (double)reactor_reward(ns + i * S, a + i * A);
        else if (kind == ORC_GRID) r[i] = grid_reward(ns + i * S, a + i * A);
        else r[i] = robot_reward(ns + i * S, a + i * A);
    }
}
ORC_API void orc_is_done(int kind, int64_t n, const float* s, uint8_t* d)
{
    const int S = kS[kind];
    for (int64_t i = 0; i < n; ++i)
        d[i] = (uint8_t)(kind == ORC_REACTOR ? reactor_is_done(s + i * S) : kind == ORC_GRID ? grid_is_done(s + i * S) : robot_is_done(s + i * S));
}
ORC_API int orc_max_threads(void)
{
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* test hook: the binary64 sin / cos of the math spec */
ORC_API void orc_spec_sincos_f64(double x, double* sn, double* cs) { spec_sincos_f64(x, sn, cs); }
