// Instruction-supply microbenchmark: N straight-line DFMAs (no loop, no reuse inside a warp) versus the same
// work in a tight loop, at several CTA shapes.  Answers: how many warp-instructions per clock can one SM issue
// when every warp streams its own copy of a ~320 KB code region (the shape of the generated k_eval kernels)?
#include <cstdio>
#include <cuda_runtime.h>
#define F8 a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
#define R5(x) x x x x x
#define R10(x) R5(x) R5(x)
#define R100(x) R10(R10(x))
#define R2500(x) R100(R5(R5(x)))
template <int SYNC>
__global__ void k_straight(double* out, double b, double c) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    if (SYNC == 0) { R2500(F8) }
    else { R100(R5(R5(F8)) __syncthreads();) }   // barrier every 25*8 = 200 instructions
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_loop(double* out, double b, double c) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
#pragma unroll 1
    for (int i = 0; i < 250; i++) { R10(F8) }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <class K>
static void run(const char* name, K k, int threads, int blocks_per_sm, double* d) {
    int sms = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int waves = 4, grid = sms * blocks_per_sm * waves;
    const int smem = (200 * 1024 / blocks_per_sm) & ~1023;   // dynamic shared memory pins the residency to blocks_per_sm
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<<<grid, threads, smem>>>(d, 0.999999, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double warp_instr = (double)grid * threads / 32 * 20000.0;
    const double clk = 1.965e9;
    printf("%-14s thr %4d blk/SM %2d warps/SM %2d : %8.3f ms  IPC/SM %.2f  (DFMA pipe %.0f%%) %s\n", name, threads, blocks_per_sm,
           threads / 32 * blocks_per_sm, best, warp_instr / (best * 1e-3) / clk / sms, 100.0 * warp_instr / (best * 1e-3) / clk / sms / 2.0,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}
int main() {
    double* d; cudaMalloc(&d, (size_t)148 * 32 * 4 * 1024 * 8);
    int shapes[][2] = {{128, 1}, {128, 2}, {128, 4}, {128, 5}, {128, 8}, {256, 2}, {256, 4}, {512, 1}, {512, 2}, {1024, 1}, {64, 8}, {32, 16}};
    for (auto& s : shapes) {
        run("loop", k_loop, s[0], s[1], d);
        run("straight", k_straight<0>, s[0], s[1], d);
        run("straight+bar", k_straight<1>, s[0], s[1], d);
    }
    return 0;
}
