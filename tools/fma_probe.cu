// FP32 pipe microbenchmark on sm_100a: warp-instructions per cycle per SM for (a) FFMA R,R,R  (b) FFMA with a
// constant-bank multiplicand  (c) packed fma.rn.f32x2.  8 independent accumulator chains per thread.
#include <cuda_runtime.h>
#include <cstdio>
struct P { float w[8]; };
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *win, P p, int iters) {
    float a[16];
    float w[8];
    for (int i = 0; i < 8; ++i) w[i] = win[i];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    float x = out[threadIdx.x & 31];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(w[r], a[i], x);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(p.w[r], a[i], x);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long acc, ww, xx;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a[i]), "f"(a[i + 1]));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w[r]));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
                    asm("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(acc) : "l"(ww), "l"(xx));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(acc));
                }
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, float *out, float *w, int flops_per_instr) {
    P p; for (int i = 0; i < 8; ++i) p.w[i] = 0.999f + i * 1e-4f;
    int iters = 2000, blocks = 148 * 4;
    k<MODE><<<blocks, 256>>>(out, w, p, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, w, p, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)blocks * 256 * iters * 8 * 16;  // scalar fma count
    printf("%-22s %.3f ms  %.2f TFLOP/s  (%.2f scalar-FMA lanes / clk / SM @1.965GHz)\n", name, ms, 2 * fmas / ms / 1e9,
           fmas / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float *out, *w; cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&w, 64);
    cudaMemset(out, 0, 148 * 4 * 256 * 4); float hw[8] = {0.9991f, 0.9992f, 0.9993f, 0.9994f, 0.9995f, 0.9996f, 0.9997f, 0.9998f};
    cudaMemcpy(w, hw, 32, cudaMemcpyHostToDevice);
    run<0>("FFMA R,R,R", out, w, 2);
    run<1>("FFMA R,c[],R", out, w, 2);
    run<2>("fma.rn.f32x2", out, w, 4);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
