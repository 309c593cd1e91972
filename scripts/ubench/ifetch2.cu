// Instruction-supply microbenchmark 2: does a taken branch every ~70 instructions (the shape of the generated
// device-evaluation kernels: ~175 taken branches per 15k executed instructions) defeat the sequential
// instruction prefetch?  Variants: straight-line; forward skip over a never-executed block; call to a far
// out-of-line function.  Same DFMA work in all of them.
#include <cstdio>
#include <cuda_runtime.h>
#define F8 a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
#define R5(x) x x x x x
#define R9(x) x x x x x x x x x
#define R10(x) R5(x) R5(x)
#define R250(x) R10(R5(R5(x)))
#define DECL double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
#define FIN out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
__global__ void k_straight(double* out, double b, double c, int flag) { DECL R250(R9(F8)) FIN }
__global__ void k_skip(double* out, double b, double c, int flag) {
    DECL
    R250(R9(F8) if (flag) { R5(F8) b += 1e-9; })
    FIN
}
__device__ __noinline__ double far_fn(double x, double b, double c) { return fma(x, b, c); }
__global__ void k_call(double* out, double b, double c, int flag) {
    DECL
    R250(R9(F8) a0 = far_fn(a0, b, c);)
    FIN
}
// data-dependent two-sided branch: both sides same size, warp-uniform condition alternating per site
__global__ void k_ifelse(double* out, double b, double c, int flag) {
    DECL
    R250(R5(F8) if (flag & 1) { R5(F8) b += 1e-9; } else { R5(F8) c += 1e-9; })
    FIN
}
template <class K>
static void run(const char* name, K k, int threads, int blocks_per_sm, double* d, double instr, int flag) {
    int sms = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int waves = 4, grid = sms * blocks_per_sm * waves;
    const int smem = (200 * 1024 / blocks_per_sm) & ~1023;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<<<grid, threads, smem>>>(d, 0.999999, 1e-9, flag); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double warp_instr = (double)grid * threads / 32 * instr;
    printf("%-10s thr %4d blk/SM %2d warps/SM %2d : %8.3f ms  DFMA IPC/SM %.2f (pipe %.0f%%) %s\n", name, threads, blocks_per_sm,
           threads / 32 * blocks_per_sm, best, warp_instr / (best * 1e-3) / 1.965e9 / sms, 100.0 * warp_instr / (best * 1e-3) / 1.965e9 / sms / 2.0,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}
int main() {
    double* d; cudaMalloc(&d, (size_t)148 * 32 * 4 * 1024 * 8);
    int shapes[][2] = {{128, 5}, {128, 8}, {256, 2}, {256, 4}, {512, 1}};
    for (auto& s : shapes) {
        run("straight", k_straight, s[0], s[1], d, 250 * 72.0, 0);
        run("skip", k_skip, s[0], s[1], d, 250 * 72.0, 0);
        run("call", k_call, s[0], s[1], d, 250 * 73.0, 0);
        run("ifelse", k_ifelse, s[0], s[1], d, 250 * 80.0, 0);
    }
    return 0;
}
