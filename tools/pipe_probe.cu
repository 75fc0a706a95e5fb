// pipe_probe.cu -- issue-rate / latency probe of the sm_100a instructions the fused rollout kernel is made of.
//
// For every op (or mix of ops) one kernel: each thread owns CH independent register chains and executes
// ITERS x UNROLL x CH ops; every CTA = one SM (grid = #SMs, 1 CTA/SM), W warps per SM sub-partition. Each warp
// reads clock64() around its loop; the figure reported is warp-instructions issued per cycle per sub-partition
// (IPC_smsp = W * ops_per_warp / mean cycles) -- 1.0 = the issue limit, 0.5 = a half-rate pipe. With W = 1, CH = 1
// the same loop measures the dependent-issue latency (cycles per op).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/pipe_probe_bin tools/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

enum Op : int {
    FADD_IMM, FADD_RR, FMUL_RR, FFMA_RRR, FFMA_RRI, FADD2, FMUL2, FFMA2,
    FMNMX, FSEL, FSETP, LOP3, IMADW, I2FP, MUFU, IADD3,
    MIX_FMUL_LOP3, MIX_2FMUL_LOP3, MIX_FFMA2_LOP3, MIX_FMUL2_FSETP, MIX_FMUL_FSETP, MIX_FMUL_FMNMX, MIX_IMADW_LOP3,
    MIX_FADD2_FMNMX_LOP3, MIX_FMUL_IMADW,
    DADD_, DMUL_, DFMA_, F2F_64_32, F2F_32_64, FMNMX3_, IMAD_LO, IMAD_HI, SHF_, LDS128_, SHFL_, MIX_DFMA_FMUL, MIX_F2F_FMUL, N_OPS
};
static const char* kNames[N_OPS] = {
    "FADD r,imm", "FADD r,r", "FMUL r,r", "FFMA r,r,r", "FFMA r,r,imm", "FADD2 (f32x2)", "FMUL2 (f32x2)", "FFMA2 (f32x2)",
    "FMNMX", "FSEL", "FSETP (and-chain)", "LOP3 r,r,r", "IMAD.WIDE.U32", "I2FP.U32", "MUFU.RCP", "IADD3",
    "mix 1 FMUL + 1 LOP3", "mix 2 FMUL + 1 LOP3", "mix 1 FFMA2 + 1 LOP3", "mix 1 FMUL2 + 1 FSETP", "mix 1 FMUL + 1 FSETP", "mix 1 FMUL + 1 FMNMX",
    "mix 1 IMAD.WIDE + 1 LOP3", "mix 1 FADD2 + 1 FMNMX + 1 LOP3", "mix 1 FMUL + 1 IMAD.WIDE",
    "DADD", "DMUL", "DFMA", "F2F.F64.F32 (+F2F back)", "F2F.F32.F64 only (see prev)", "FMNMX3", "IMAD lo32", "IMAD.HI", "SHF.R", "LDS.128", "SHFL.IDX", "mix 1 DFMA + 1 FMUL", "mix 1 F2F pair + 2 FMUL"
};
// warp-instructions per chain per unrolled slot
static const int kOpsPerSlot[N_OPS] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 3, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 2, 4};

template <int OP, int CH>
__global__ void __launch_bounds__(1024, 1) probe(float* out, long long* cyc, int iters, float yv, float zv)
{
    __shared__ float smem[2048];
    for (int q = threadIdx.x; q < 2048; q += blockDim.x) smem[q] = (float)q;
    float x[CH], y2[CH];
    unsigned u[CH];
    unsigned long long d[CH];
    float y = yv, z = zv;
    unsigned uy = __float_as_uint(yv) | 1u, uz = __float_as_uint(zv) | 3u;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        x[c] = 1.0f + 1e-3f * (float)(threadIdx.x + c); y2[c] = x[c];
        u[c] = threadIdx.x * 2654435761u + c;
        d[c] = ((unsigned long long)__float_as_uint(x[c]) << 32) | __float_as_uint(x[c] + 0.5f);
    }
    unsigned long long dy = ((unsigned long long)__float_as_uint(yv) << 32) | __float_as_uint(yv);
    unsigned long long dz = ((unsigned long long)__float_as_uint(zv) << 32) | __float_as_uint(zv);
    int pacc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if constexpr (OP == FADD_IMM) asm volatile("add.rn.f32 %0, %0, 0f33D6BF95;" : "+f"(x[c]));
                else if constexpr (OP == FADD_RR) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y));
                else if constexpr (OP == FMUL_RR) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y));
                else if constexpr (OP == FFMA_RRR) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(y), "f"(z));
                else if constexpr (OP == FFMA_RRI) asm volatile("fma.rn.f32 %0, %0, %1, 0f33D6BF95;" : "+f"(x[c]) : "f"(y));
                else if constexpr (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[c]) : "l"(dy));
                else if constexpr (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[c]) : "l"(dy));
                else if constexpr (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[c]) : "l"(dy), "l"(dz));
                else if constexpr (OP == FMNMX) { if (r & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y)); else asm volatile("min.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(z)); }
                else if constexpr (OP == FSEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.f32 %0, %0, %1, p;}" : "+f"(x[c]) : "f"(y), "r"(iters));
                else if constexpr (OP == FSETP) asm volatile("{.reg .pred p; setp.ne.s32 p, %0, 0; setp.gt.and.f32 p, %1, %2, p; selp.s32 %0, 1, 0, p;}" : "+r"(pacc) : "f"(x[c]), "f"(y));
                else if constexpr (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(uy), "r"(uz));
                else if constexpr (OP == IMADW) asm volatile("{.reg .b64 t; .reg .b32 lo; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {lo, %0}, t;}" : "+r"(u[c]));
                else if constexpr (OP == I2FP) asm volatile("{.reg .f32 t; cvt.rn.f32.u32 t, %0; mov.b32 %0, t;}" : "+r"(u[c]));
                else if constexpr (OP == MUFU) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
                else if constexpr (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[c]) : "r"(uy));
                else if constexpr (OP == DADD_) asm volatile("{.reg .f64 t, s; mov.b64 t, %0; mov.b64 s, %1; add.rn.f64 t, t, s; mov.b64 %0, t;}" : "+l"(d[c]) : "l"(dy));
                else if constexpr (OP == DMUL_) asm volatile("{.reg .f64 t, s; mov.b64 t, %0; mov.b64 s, %1; mul.rn.f64 t, t, s; mov.b64 %0, t;}" : "+l"(d[c]) : "l"(dy));
                else if constexpr (OP == DFMA_) asm volatile("{.reg .f64 t, s, q; mov.b64 t, %0; mov.b64 s, %1; mov.b64 q, %2; fma.rn.f64 t, t, s, q; mov.b64 %0, t;}" : "+l"(d[c]) : "l"(dy), "l"(dz));
                else if constexpr (OP == F2F_64_32) asm volatile("{.reg .f64 t; cvt.f64.f32 t, %0; cvt.rn.f32.f64 %0, t;}" : "+f"(x[c]));
                else if constexpr (OP == F2F_32_64) asm volatile("{.reg .f64 t; mov.b64 t, %1; cvt.rn.f32.f64 %0, t;}" : "+f"(x[c]) : "l"(d[c]));
                else if constexpr (OP == FMNMX3_) asm volatile("{.reg .f32 t; max.f32 t, %0, %1; max.f32 %0, t, %2;}" : "+f"(x[c]) : "f"(y), "f"(z));
                else if constexpr (OP == IMAD_LO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(uy), "r"(uz));
                else if constexpr (OP == IMAD_HI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(u[c]) : "r"(uy));
                else if constexpr (OP == SHF_) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(u[c]) : "r"(uy));
                else if constexpr (OP == LDS128_) { float4 v = *reinterpret_cast<const float4*>(&smem[(u[c] & 0x1ffu) * 4]); u[c] = __float_as_uint(v.x) + __float_as_uint(v.w); }
                else if constexpr (OP == SHFL_) u[c] = __shfl_sync(0xffffffffu, u[c], (int)(u[c] & 31u));
                else if constexpr (OP == MIX_DFMA_FMUL) {
                    asm volatile("{.reg .f64 t, s, q; mov.b64 t, %0; mov.b64 s, %1; mov.b64 q, %2; fma.rn.f64 t, t, s, q; mov.b64 %0, t;}" : "+l"(d[c]) : "l"(dy), "l"(dz));
                    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y));
                } else if constexpr (OP == MIX_F2F_FMUL) {
                    asm volatile("{.reg .f64 t; cvt.f64.f32 t, %0; cvt.rn.f32.f64 %0, t;}" : "+f"(x[c]));
                    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(y2[c]) : "f"(y));
                    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(y2[c]) : "f"(z));
                }
                else if constexpr (OP == MIX_FMUL_LOP3) {
                    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(uy), "r"(uz));
                } else if constexpr (OP == MIX_2FMUL_LOP3) {
                    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(uy), "r"(uz));
                    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(z));
                } else if constexpr (OP == MIX_FFMA2_LOP3) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[c]) : "l"(dy), "l"(dz));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(uy), "r"(uz));
                } else if constexpr (OP == MIX_FMUL2_FSETP) {
                    asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[c]) : "l"(dy));
                    asm volatile("{.reg .pred p; setp.ne.s32 p, %0, 0; setp.gt.and.f32 p, %1, %2, p; selp.s32 %0, 1, 0, p;}" : "+r"(pacc) : "f"(x[c]), "f"(y));
                } else if constexpr (OP == MIX_FMUL_FSETP) {
                    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y));
                    asm volatile("{.reg .pred p; setp.ne.s32 p, %0, 0; setp.gt.and.f32 p, %1, %2, p; selp.s32 %0, 1, 0, p;}" : "+r"(pacc) : "f"(x[c]), "f"(y));
                } else if constexpr (OP == MIX_FMUL_FMNMX) {
                    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y));
                    asm volatile("max.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(z));
                } else if constexpr (OP == MIX_IMADW_LOP3) {
                    asm volatile("{.reg .b64 t; .reg .b32 lo; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {lo, %0}, t;}" : "+r"(u[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(uy), "r"(uz));
                } else if constexpr (OP == MIX_FADD2_FMNMX_LOP3) {
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[c]) : "l"(dy));
                    asm volatile("max.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(z));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(uy), "r"(uz));
                } else if constexpr (OP == MIX_FMUL_IMADW) {
                    asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(y));
                    asm volatile("{.reg .b64 t; .reg .b32 lo; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {lo, %0}, t;}" : "+r"(u[c]));
                }
            }
        }
    }
    const long long t1 = clock64();
    float acc = (float)pacc;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc += x[c] + y2[c] + (float)u[c] + (float)(d[c] & 0xffff) + (float)(d[c] >> 48);
    if (acc == 123.456f) out[0] = acc;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

template <int OP, int CH>
static void run(int sms, int warps_per_smsp, int iters, float* out, long long* cyc_d, std::vector<long long>& cyc_h, const char* tag)
{
    const int threads = warps_per_smsp * 4 * 32;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe<OP, CH><<<sms, threads>>>(out, cyc_d, iters / 4, 1.0000001f, 1e-7f);       // warm-up
    CK(cudaEventRecord(e0));
    probe<OP, CH><<<sms, threads>>>(out, cyc_d, iters, 1.0000001f, 1e-7f);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const int nw = sms * threads / 32;
    CK(cudaMemcpy(cyc_h.data(), cyc_d, nw * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0; long long mx = 0;
    for (int i = 0; i < nw; ++i) { mean += (double)cyc_h[i]; mx = cyc_h[i] > mx ? cyc_h[i] : mx; }
    mean /= nw;
    const double ops_per_warp = (double)iters * 8 * CH * kOpsPerSlot[OP];
    const double ipc = warps_per_smsp * ops_per_warp / mean;
    printf("%-34s %s W=%d CH=%d  IPC/smsp %.3f  (cycles/op/warp %.2f)  mean cyc %.0f max %lld  %.3f ms -> %.2f T warp-lane-op/s\n",
           kNames[OP], tag, warps_per_smsp, CH, ipc, mean / ops_per_warp, mean, mx, ms,
           ops_per_warp * nw * 32 / (ms * 1e-3) / 1e12);
}

template <int OP>
static void suite(int sms, int iters, float* out, long long* cyc_d, std::vector<long long>& cyc_h)
{
    run<OP, 8>(sms, 4, iters, out, cyc_d, cyc_h, "tput");
    run<OP, 8>(sms, 1, iters, out, cyc_d, cyc_h, "1warp");
    run<OP, 1>(sms, 1, iters, out, cyc_d, cyc_h, "lat ");
}

int main(int argc, char** argv)
{
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, dev));
    const int sms = pr.multiProcessorCount;
    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    const bool only_new = argc > 2;
    printf("device %s, %d SMs, clock %d kHz; iters %d\n", pr.name, sms, pr.clockRate, iters);
    float* out; long long* cyc_d;
    CK(cudaMalloc(&out, 4));
    CK(cudaMalloc(&cyc_d, sms * 32 * sizeof(long long)));
    std::vector<long long> cyc_h(sms * 32);
    if (!only_new) {
    suite<FADD_IMM>(sms, iters, out, cyc_d, cyc_h);
    suite<FADD_RR>(sms, iters, out, cyc_d, cyc_h);
    suite<FMUL_RR>(sms, iters, out, cyc_d, cyc_h);
    suite<FFMA_RRR>(sms, iters, out, cyc_d, cyc_h);
    suite<FFMA_RRI>(sms, iters, out, cyc_d, cyc_h);
    suite<FADD2>(sms, iters, out, cyc_d, cyc_h);
    suite<FMUL2>(sms, iters, out, cyc_d, cyc_h);
    suite<FFMA2>(sms, iters, out, cyc_d, cyc_h);
    suite<FMNMX>(sms, iters, out, cyc_d, cyc_h);
    suite<FSEL>(sms, iters, out, cyc_d, cyc_h);
    suite<FSETP>(sms, iters, out, cyc_d, cyc_h);
    suite<LOP3>(sms, iters, out, cyc_d, cyc_h);
    suite<IMADW>(sms, iters, out, cyc_d, cyc_h);
    suite<I2FP>(sms, iters, out, cyc_d, cyc_h);
    suite<MUFU>(sms, iters, out, cyc_d, cyc_h);
    suite<IADD3>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_FMUL_LOP3>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_2FMUL_LOP3>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_FFMA2_LOP3>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_FMUL2_FSETP>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_FMUL_FSETP>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_FMUL_FMNMX>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_IMADW_LOP3>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_FADD2_FMNMX_LOP3>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_FMUL_IMADW>(sms, iters, out, cyc_d, cyc_h);
    }
    suite<DADD_>(sms, iters, out, cyc_d, cyc_h);
    suite<DMUL_>(sms, iters, out, cyc_d, cyc_h);
    suite<DFMA_>(sms, iters, out, cyc_d, cyc_h);
    suite<F2F_64_32>(sms, iters, out, cyc_d, cyc_h);
    suite<F2F_32_64>(sms, iters, out, cyc_d, cyc_h);
    suite<FMNMX3_>(sms, iters, out, cyc_d, cyc_h);
    suite<IMAD_LO>(sms, iters, out, cyc_d, cyc_h);
    suite<IMAD_HI>(sms, iters, out, cyc_d, cyc_h);
    suite<SHF_>(sms, iters, out, cyc_d, cyc_h);
    suite<LDS128_>(sms, iters, out, cyc_d, cyc_h);
    suite<SHFL_>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_DFMA_FMUL>(sms, iters, out, cyc_d, cyc_h);
    suite<MIX_F2F_FMUL>(sms, iters, out, cyc_d, cyc_h);
    return 0;
}
