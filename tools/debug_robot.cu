#include <cstdio>
#include "../neorl-industrial-gym_b200/csrc/nig_envs.cuh"
using namespace nig;
__global__ void k(const float* st, const float* act) {
    float s[24], a[7], nz[1] = {0}, ns[24];
    for (int i = 0; i < 24; ++i) s[i] = st[i];
    for (int i = 0; i < 7; ++i) a[i] = act[i];
    Robot::dynamics(s, a, nz, ns);
    bool d0 = ns[23] > 0.95f;
    bool f0 = fabsf(ns[18]) > 80.0f, f1 = fabsf(ns[19]) > 80.0f, f2 = fabsf(ns[20]) > 80.0f;
    bool in0 = (double)ns[0] >= -0.6, in1 = (double)ns[0] <= 0.6, in2 = (double)ns[1] >= -0.6, in3 = (double)ns[1] <= 0.6, in4 = (double)ns[2] >= -0.1, in5 = (double)ns[2] <= 0.9;
    printf("d0 %d f %d %d %d in %d %d %d %d %d %d  is_done %d\n", d0, f0, f1, f2, in0, in1, in2, in3, in4, in5, (int)Robot::is_done(ns));
    printf("ns: "); for (int i = 0; i < 24; ++i) printf("%g ", ns[i]); printf("\n");
}
int main() {
    float st[24] = {0.29567f, -0.022489f, 0.419117f, 0, 0, 0, 1, 0.75203f, -0.641178f, 0.700698f, 0.816616f, 3.02913f, -0.255314f, 0.808393f, 0.18368f, -0.235711f, -0.073847f, 0, 0, 0, 0, 0.652022f, 0.80445f, 0.902163f};
    float act[7] = {0.4f, -1.0f, 0.66f, -1.0f, -1.0f, 0.88f, -0.92f};
    float *ds, *da; cudaMalloc(&ds, sizeof st); cudaMalloc(&da, sizeof act);
    cudaMemcpy(ds, st, sizeof st, cudaMemcpyHostToDevice); cudaMemcpy(da, act, sizeof act, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(ds, da); cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
