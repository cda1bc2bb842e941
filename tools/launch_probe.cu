// Per-launch cost of dependent kernels inside a replayed CUDA graph on this GPU (what a V-cycle level costs at minimum):
// empty kernels of several grid shapes, with and without programmatic dependent launch.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/launch_probe tools/launch_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

__global__ void k_empty(int *p) {
    if (p && threadIdx.x == 0 && blockIdx.x == 0xffffff) *p = 1;
}
__global__ void k_pdl(int *p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p && threadIdx.x == 0 && blockIdx.x == 0xffffff) *p = 1;
}
__global__ void k_touch(float *a, int n) {  // one L2 round trip per thread
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = a[i] + 1.0f;
}

template <class K, class... A>
static void launch(K k, int grid, int block, bool pdl, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k, args...);
}

template <class F>
static float time_graph(F body, int nk, cudaStream_t st) {
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < nk; ++i) body();
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    for (int i = 0; i < 50; ++i) cudaGraphLaunch(ge, st);
    cudaEventRecord(b, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
    return ms * 1e3f / (50.0f * nk);
}

int main() {
    cudaStream_t st;
    cudaStreamCreate(&st);
    float *buf;
    cudaMalloc(&buf, 1 << 24);
    cudaMemset(buf, 0, 1 << 24);
    const int nk = 40;
    const int shapes[][2] = {{1, 32}, {16, 352}, {148, 256}, {296, 256}, {1184, 352}, {2368, 256}};
    for (auto &s : shapes) {
        float t0 = time_graph([&] { k_empty<<<s[0], s[1], 0, st>>>(nullptr); }, nk, st);
        float t1 = time_graph([&] { launch(k_pdl, s[0], s[1], false, st, (int *)nullptr); }, nk, st);
        float t2 = time_graph([&] { launch(k_pdl, s[0], s[1], true, st, (int *)nullptr); }, nk, st);
        float t3 = time_graph([&] { launch(k_touch, s[0], s[1], true, st, buf, s[0] * s[1]); }, nk, st);
        float t4 = time_graph([&] { launch(k_touch, s[0], s[1], false, st, buf, s[0] * s[1]); }, nk, st);
        printf("grid %5d x %3d: empty %.2f us | griddep no-attr %.2f | PDL %.2f | touch PDL %.2f | touch serial %.2f\n", s[0],
               s[1], t0, t1, t2, t3, t4);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
