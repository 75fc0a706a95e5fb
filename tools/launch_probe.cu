// Floor of a dependent kernel chain on this GPU: 100 launches captured in a CUDA graph and replayed, each launch reading what the
// previous one wrote (the structure of IndustrialEnv.step called in a loop). Variants: empty kernel / 12-row load-modify-store
// per thread (the reactor step's memory shape, ~no arithmetic); grid = 512 x 128 (65,536 envs) or 128 x 128; with / without
// programmatic dependent launch.    nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/launch_probe_bin tools/launch_probe.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_empty(float* s, int pitch, int pdl)
{
    if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
}
__global__ void k_rows(float* s, int pitch, int pdl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = s[k * pitch + i];
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k * pitch + i] = v[k] * 1.0001f + v[(k + 1) % 12] * 1e-6f;
}
// ~250 dependent FP32 operations per thread between the loads and the stores (the step's arithmetic chain)
__global__ void k_chain(float* s, int pitch, int pdl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = s[k * pitch + i];
    float a = v[0];
#pragma unroll 1
    for (int t = 0; t < 60; ++t) { a = a * 1.0001f + v[1]; a = a * 0.9999f + v[2]; a = a + v[3]; a = a * v[4]; }
    v[0] = a;
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k * pitch + i] = v[k];
}

// the same plus what a step's bookkeeping adds: a CTA-wide reduction of a counter through shared memory and one global atomic
// per CTA (into 64 shards)
__global__ void k_chain_stats(float* s, int pitch, int pdl)
{
    __shared__ unsigned int cnt;
    if (threadIdx.x == 0) cnt = 0u;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = s[k * pitch + i];
    float a = v[0];
#pragma unroll 1
    for (int t = 0; t < 60; ++t) { a = a * 1.0001f + v[1]; a = a * 0.9999f + v[2]; a = a + v[3]; a = a * v[4]; }
    v[0] = a;
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k * pitch + i] = v[k];
    const unsigned int w = __reduce_add_sync(0xffffffffu, a > 0.5f ? 1u : 0u);
    if ((threadIdx.x & 31) == 0) atomicAdd(&cnt, w);
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned int*>(s) + 12 * pitch + (blockIdx.x & 63) * 32, cnt);
}

// the chain kernel + per-WARP bookkeeping into a 256-byte row of a separate allocation (row = warp % g_nrows): MODE 0 = three 64-bit
// atomic adds (RED), 1 = three plain 64-bit stores, 2 = the row loaded right after the wait and written back as old + delta
__device__ unsigned long long* g_rows = nullptr;
__device__ int g_nrows = 1 << 20;
template <int MODE>
__global__ void k_chain_rows(float* s, int pitch, int pdl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
    unsigned long long* row = g_rows + (size_t)((i >> 5) % g_nrows) * 32;
    unsigned long long old0 = 0, old1 = 0, old2 = 0;
    if (MODE == 2 && (threadIdx.x & 31) == 0) { old0 = row[0]; old1 = row[5]; old2 = row[10]; }
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = s[k * pitch + i];
    float a = v[0];
#pragma unroll 1
    for (int t = 0; t < 60; ++t) { a = a * 1.0001f + v[1]; a = a * 0.9999f + v[2]; a = a + v[3]; a = a * v[4]; }
    v[0] = a;
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k * pitch + i] = v[k];
    const unsigned int w = __reduce_add_sync(0xffffffffu, a > 0.5f ? 1u : 0u) + 32u;
    if ((threadIdx.x & 31) == 0) {
        if (MODE == 0) { atomicAdd(&row[0], (unsigned long long)w); atomicAdd(&row[5], (unsigned long long)w); atomicAdd(&row[10], (unsigned long long)w); }
        else { row[0] = old0 + w; row[5] = old1 + w; row[10] = old2 + w; }
    }
}

// the chain kernel with ~1 KB of kernel parameters (what a by-value descriptor block costs per launch)
struct Fat { float* s; int pitch, pdl; unsigned int pad[250]; };
__global__ void k_chain_fat(const __grid_constant__ Fat f)
{
    float* s = f.s; const int pitch = f.pitch, pdl = f.pdl;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;"); }
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = s[k * pitch + i];
    float a = v[0] + __uint_as_float(f.pad[threadIdx.x & 127]);
#pragma unroll 1
    for (int t = 0; t < 60; ++t) { a = a * 1.0001f + v[1]; a = a * 0.9999f + v[2]; a = a + v[3]; a = a * v[4]; }
    v[0] = a;
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k * pitch + i] = v[k];
}

template <class K, class... A>
int run_args(const char* name, K kern, int grid, int pdl, const A&... args)
{
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    for (int l = 0; l < 100; ++l) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = pdl;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, kern, args...));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < 20; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-34s grid %4d x 128  pdl %d: %.3f us per launch\n", name, grid, pdl, ms * 1e3 / 2000);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
    return 0;
}

template <class K>
int run(const char* name, K kern, int grid, int pdl, float* d, int pitch)
{
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    for (int l = 0; l < 100; ++l) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = pdl;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, kern, d, pitch, pdl));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < 20; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-28s grid %4d x 128  pdl %d: %.3f us per launch\n", name, grid, pdl, ms * 1e3 / 2000);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
    return 0;
}

int main()
{
    const int pitch = 65536;
    float* d; CK(cudaMalloc(&d, (size_t)13 * pitch * sizeof(float))); CK(cudaMemset(d, 0, (size_t)13 * pitch * sizeof(float)));
    for (int pdl = 0; pdl < 2; ++pdl)
        for (int grid : {128, 512}) {
            if (run("empty kernel", k_empty, grid, pdl, d, pitch)) return 1;
            if (run("12-row load / store", k_rows, grid, pdl, d, pitch)) return 1;
            if (run("load, 240-op chain, store", k_chain, grid, pdl, d, pitch)) return 1;
            if (run("... + CTA reduce + atomic", k_chain_stats, grid, pdl, d, pitch)) return 1;
            {
                Fat f; memset(&f, 0, sizeof f); f.s = d; f.pitch = pitch; f.pdl = pdl;
                if (run_args("... with 1 KB of parameters", k_chain_fat, grid, pdl, f)) return 1;
                unsigned long long* rows; CK(cudaMalloc(&rows, (size_t)2048 * 32 * 8)); CK(cudaMemset(rows, 0, (size_t)2048 * 32 * 8));
                CK(cudaMemcpyToSymbol(g_rows, &rows, sizeof rows));
                int nr = 2048; CK(cudaMemcpyToSymbol(g_nrows, &nr, sizeof nr));
                if (run_args("... + 3 RED.64 per warp, own row", k_chain_rows<0>, grid, pdl, d, pitch, pdl)) return 1;
                nr = 128; CK(cudaMemcpyToSymbol(g_nrows, &nr, sizeof nr));
                if (run_args("... + 3 RED.64 per warp, 128 rows", k_chain_rows<0>, grid, pdl, d, pitch, pdl)) return 1;
                nr = 2048; CK(cudaMemcpyToSymbol(g_nrows, &nr, sizeof nr));
                if (run_args("... + 3 plain stores per warp", k_chain_rows<1>, grid, pdl, d, pitch, pdl)) return 1;
                if (run_args("... + early row load, 3 stores", k_chain_rows<2>, grid, pdl, d, pitch, pdl)) return 1;
                cudaFree(rows);
            }
        }
    return 0;
}
