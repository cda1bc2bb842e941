// Standalone probe of TMA tiled-load constraints on sm_100a: which (dtype, box width, start coordinate) combinations
// complete and deliver the right data.  usage: tma_probe <u8|f32> <boxw> <c0> ; prints OK / TIMEOUT / MISMATCH
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../multigrid-feanet_b200/mgfea/csrc/mgfea_ptx.cuh"
using namespace mgfea;

__global__ void k(const __grid_constant__ CUtensorMap map, unsigned char *out, int bytes, int c0, int c1, int *status) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) {
        mbar_init((uint64_t *)&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx((uint64_t *)&bar, bytes);
        tma_load_2d(sm, &map, (uint64_t *)&bar, c0, c1);
    }
    unsigned spins = 0;
    bool ok = false;
    while (spins++ < (1u << 20)) {
        if (mbar_try_wait((uint64_t *)&bar, 0)) { ok = true; break; }
    }
    if (threadIdx.x == 0) *status = ok ? 1 : 0;
    __syncthreads();
    if (ok) for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}

typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    bool u8 = !strcmp(argv[1], "u8");
    int boxw = atoi(argv[2]), c0 = atoi(argv[3]);
    int es = u8 ? 1 : 4, W = 100, H = 40, pitch = 128 * es, boxh = 8, c1 = -1;
    std::vector<unsigned char> h((size_t)H * pitch);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + 3);
    unsigned char *d, *o; int *st;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    int bytes = boxw * es * boxh;
    cudaMalloc(&o, bytes); cudaMalloc(&st, 4);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap map; cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh}, est[2] = {1, 1};
    CUresult r = ((PFN)p)(&map, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides,
                          box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s boxw=%d c0=%d: ENCODE_FAIL %d\n", argv[1], boxw, c0, (int)r); return 0; }
    k<<<1, 128, bytes + 1024>>>(map, o, bytes, c0, c1, st);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s boxw=%d c0=%d: CUDA_ERROR %s\n", argv[1], boxw, c0, cudaGetErrorString(e)); return 0; }
    int hs; cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
    if (!hs) { printf("%s boxw=%d c0=%d: TIMEOUT\n", argv[1], boxw, c0); return 0; }
    std::vector<unsigned char> ho(bytes); cudaMemcpy(ho.data(), o, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < boxh; ++r2) for (int c = 0; c < boxw; ++c) for (int b = 0; b < es; ++b) {
        int gy = c1 + r2, gx = c0 + c; unsigned char want = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) want = h[(size_t)gy * pitch + gx * es + b];
        if (ho[(r2 * boxw + c) * es + b] != want) ++bad;
    }
    printf("%s boxw=%d c0=%d: %s (%d bad)\n", argv[1], boxw, c0, bad ? "MISMATCH" : "OK", bad);
    return 0;
}
