// Micro-benchmark: issue rate of scalar FFMA vs packed FFMA2 with register / uniform-register multiplicands on sm_100a.
// Every variant runs CH independent accumulator chains per thread, ITER x 9 dependent steps each (a 3x3 stencil's shape),
// 8 warps per CTA, 2 CTAs per SM.  Prints cycles per warp-instruction per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
struct W { float w[9]; };
constexpr int CH = 8, ITER = 256;

template <int V>
__global__ void __launch_bounds__(256, 2) k(const __grid_constant__ W wp, const float *wg, float *out, long long *cyc) {
    float acc[CH];
    u64 acc2[CH];
    float x[CH];
    u64 x2[CH];
    float wr[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wr[q] = (V == 0 || V == 2) ? wg[q + (threadIdx.x & 1)] : wp.w[q];  // per-thread value: vector regs
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        acc[c] = 0.f;
        x[c] = 1.0f + 1e-3f * (threadIdx.x + c);
        acc2[c] = 0;
        x2[c] = pack2(x[c], x[c] + 1.f);
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (V == 0 || V == 1) acc[c] = __fmaf_rn(wr[q], x[c], acc[c]);
                else acc2[c] = fma2(pack2(wr[q], wr[q]), x2[c], acc2[c]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c] + (float)(acc2[c] & 0xffff);
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char *name) {
    W w;
    for (int q = 0; q < 9; ++q) w.w[q] = 0.5f + 0.01f * q;
    float *wg, *out;
    long long *cyc;
    cudaMalloc(&wg, 64);
    cudaMemcpy(wg, w.w, 36, cudaMemcpyHostToDevice);
    const int grid = 148 * 2;
    cudaMalloc(&out, grid * 256 * 4);
    cudaMalloc(&cyc, grid * 8);
    k<V><<<grid, 256>>>(w, wg, out, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    k<V><<<grid, 256>>>(w, wg, out, cyc);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    long long h[296];
    cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += h[i];
    avg /= grid;
    // per SMSP: 16 warps per SM / 4 = 4 warps, each ITER*9*CH instructions
    const double instr_per_smsp = 4.0 * ITER * 9 * CH;
    printf("%-28s %8.1f us  %10.0f cycles  %.2f cycles / warp-instr / SMSP  err=%d\n", name, ms * 1e3, avg, avg / instr_per_smsp,
           (int)cudaGetLastError());
}
int main() {
    run<0>("FFMA  R, R(w), R");
    run<1>("FFMA  R, UR/c(w), R");
    run<2>("FFMA2 R, R(w).F32, R");
    run<3>("FFMA2 R, UR(w).F32, R");
    return 0;
}
