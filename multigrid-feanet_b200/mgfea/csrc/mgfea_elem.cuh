// General per-ELEMENT conductivity operator (data model `material`, one value per element:
// Data/dataset.py:71-104).  The reference ships only the 16-pattern two-phase table (FEANet/mesh.py:103-117); this is that
// operator with the pattern lookup replaced by the element values themselves.  For output node (i,j) with elements
// NW = E(i-1,j-1), NE = E(i-1,j), SW = E(i,j-1), SE = E(i,j) the nine tap weights are `generate_kernel`'s expressions of
// the SOURCE nodes (KNet.forward indexes the table by the source node, FEANet/model.py:22-30), each of which only involves
// elements shared with the output node -- evaluated in the same fp32 order, so on a two-phase map the result is
// bit-identical to the pattern kernels on every interior node (tests/test_gpu_parity.py::test_element_*).
//
// First implementation: one pass per operator (K u / residual / one Jacobi sweep), every thread owns 4 consecutive nodes
// of a row; neighbours come through L1/L2.  HBM bound at ~24 B/node (u, f, a in; u out) per sweep.
#pragma once
#include "mgfea_tile.cuh"

namespace mgfea {

struct ElemParams {
    int N, B, pitch;
    long long plane;
    const float *a;  // [N][pitch] fp32, element (r,c) at a[r*pitch+c] for r,c < N-1, zero elsewhere
    const float *u;
    const float *f;
    float *out;
    float omega;
    float ke[16];  // element matrix, FEANet/mesh.py:28-31 (fp32)
};

__device__ __forceinline__ float elem_ld(const float *a, int n, int pitch, int r, int c) {
    return (r < 0 || c < 0 || r >= n || c >= n) ? 0.0f : __ldg(a + (long long)r * pitch + c);
}

// MODE 0: out = K u (all nodes, zero padding)   MODE 1: out = f - K u   MODE 2: one weighted-Jacobi sweep with the default
// Dirichlet ring (input ring -> 0, u + (omega/d) (f - K u), output ring -> 0; FEANet/jacobi.py:39-47)
template <int MODE>
__global__ void __launch_bounds__(256) mg_elem_kernel(const ElemParams p) {
    const int N = p.N, n = N - 1;
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y, b = blockIdx.z;
    if (y >= N || x0 >= p.pitch) return;
    const float *ub = p.u + (long long)b * p.plane;
    // 3 rows x 6 columns of u (x0-1 .. x0+4), zero outside the domain; MODE 2: the sweep's input reset (ring -> 0)
    float uu[3][6];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int yy = y - 1 + d;
        const bool rowok = (yy >= 0 && yy < N);
        const bool ring_row = (MODE == 2) && (yy == 0 || yy == N - 1);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int xx = x0 - 1 + q;
            float v = (rowok && xx >= 0 && xx < N) ? __ldg(ub + (long long)yy * p.pitch + xx) : 0.0f;
            if (MODE == 2 && (ring_row || xx == 0 || xx == N - 1)) v = 0.0f;
            uu[d][q] = v;
        }
    }
    // elements rows y-1, y; columns x0-1 .. x0+3
    float ea[2][5];
#pragma unroll
    for (int d = 0; d < 2; ++d)
#pragma unroll
        for (int q = 0; q < 5; ++q) ea[d][q] = elem_ld(p.a, n, p.pitch, y - 1 + d, x0 - 1 + q);
    const float *ke = p.ke;
#define KE(r, c) ke[4 * (r) + (c)]
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float nw = ea[0][e], ne = ea[0][e + 1], sw = ea[1][e], se = ea[1][e + 1];
        float w[9];
        w[0] = __fmul_rn(nw, KE(1, 3));
        w[1] = __fadd_rn(__fmul_rn(ne, KE(1, 2)), __fmul_rn(nw, KE(0, 3)));
        w[2] = __fmul_rn(ne, KE(0, 2));
        w[3] = __fadd_rn(__fmul_rn(nw, KE(2, 3)), __fmul_rn(sw, KE(1, 0)));
        w[4] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(sw, KE(0, 0)), __fmul_rn(se, KE(1, 1))), __fmul_rn(ne, KE(2, 2))),
                         __fmul_rn(nw, KE(3, 3)));
        w[5] = __fadd_rn(__fmul_rn(ne, KE(3, 2)), __fmul_rn(se, KE(0, 1)));
        w[6] = __fmul_rn(sw, KE(2, 0));
        w[7] = __fadd_rn(__fmul_rn(se, KE(2, 1)), __fmul_rn(sw, KE(3, 0)));
        w[8] = __fmul_rn(se, KE(3, 1));
        // row-major FMA chain; taps outside the domain are skipped by the reference's zero padding (their u is 0 here and
        // fma(w, 0, acc) == acc exactly, except that a chain of nothing but zeros must stay +0: start from fmul)
        float s = __fmul_rn(w[0], uu[0][e]);
        s = __fmaf_rn(w[1], uu[0][e + 1], s);
        s = __fmaf_rn(w[2], uu[0][e + 2], s);
        s = __fmaf_rn(w[3], uu[1][e], s);
        s = __fmaf_rn(w[4], uu[1][e + 1], s);
        s = __fmaf_rn(w[5], uu[1][e + 2], s);
        s = __fmaf_rn(w[6], uu[2][e], s);
        s = __fmaf_rn(w[7], uu[2][e + 1], s);
        s = __fmaf_rn(w[8], uu[2][e + 2], s);
        const int xx = x0 + e;
        if (MODE == 0) {
            o[e] = s;
        } else {
            const float fv = (xx < N) ? __ldg(p.f + (long long)b * p.plane + (long long)y * p.pitch + xx) : 0.0f;
            const float r = __fsub_rn(fv, s);
            if (MODE == 1) {
                o[e] = r;
            } else {
                const float inv = __fmul_rn(__fdiv_rn(1.0f, w[4]), p.omega);
                const float un = __fadd_rn(__fmul_rn(inv, r), uu[1][e + 1]);
                o[e] = (y == 0 || y == N - 1 || xx == 0 || xx >= N - 1) ? 0.0f : un;
            }
        }
        if (xx >= N) o[e] = 0.0f;  // padding columns stay zero
    }
#undef KE
    *reinterpret_cast<float4 *>(p.out + (long long)b * p.plane + (long long)y * p.pitch + x0) =
        make_float4(o[0], o[1], o[2], o[3]);
}

// JacobiBlockPBC.jacobi_convolution (FEANet/jacobi.py:50-97): periodic 3x3 stencil on the n x n torus evaluated at all
// (n+1)^2 nodes (node n == node 0), single pattern; f_pad is the caller-padded (N+2)^2 load vector the reference asks for
__global__ void __launch_bounds__(256) mg_jacobi_pbc_kernel(const float *u, float *out, const float *fpad, const float *w9,
                                                           const float *invd, int N, int pitch, long long plane, int pitch_f,
                                                           long long plane_f) {
    const int n = N - 1;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, b = blockIdx.z;
    if (x >= pitch || y >= N) return;
    float o = 0.0f;
    if (x < N) {
        const float *ub = u + (long long)b * plane;
        float s = 0.0f;
#pragma unroll
        for (int di = -1; di <= 1; ++di) {
            int yy = y + di;
            yy = yy < 0 ? yy + n : (yy >= n ? yy - n : yy);  // y + di in [-1, n + 1]
            yy = yy >= n ? yy - n : yy;
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                int xx = x + dj;
                xx = xx < 0 ? xx + n : (xx >= n ? xx - n : xx);
                xx = xx >= n ? xx - n : xx;
                const float v = __ldg(ub + (long long)yy * pitch + xx), w = __ldg(w9 + 3 * (di + 1) + (dj + 1));
                s = (di == -1 && dj == -1) ? __fmul_rn(w, v) : __fmaf_rn(w, v, s);
            }
        }
        const float r = __fsub_rn(__ldg(fpad + (long long)b * plane_f + (long long)(y + 1) * pitch_f + x + 1), s);
        const float uc = __ldg(ub + (long long)(y == n ? 0 : y) * pitch + (x == n ? 0 : x));  // reset_boundary(u)
        o = __fadd_rn(__fmul_rn(__ldg(invd), r), uc);
    }
    out[(long long)b * plane + (long long)y * pitch + x] = o;
}

// weight gradient of a zero-padded 3x3 correlation out = w (*) a (the backward pass of an HNet layer):
// acc[t] += sum over batch and nodes of a[i+dy-1][j+dx-1] * g[i][j], t = 3 dy + dx, accumulated in fp64
__global__ void __launch_bounds__(256) corr9_kernel(const float *a, const float *g, double *acc, int N, int pitch, long long plane) {
    __shared__ double red[8][9];
    const int b = blockIdx.y;
    const float *ab = a + (long long)b * plane, *gb = g + (long long)b * plane;
    double s[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) s[t] = 0.0;
    for (int y = blockIdx.x; y < N; y += gridDim.x)
        for (int x = threadIdx.x; x < N; x += 256) {
            const double gv = (double)gb[(long long)y * pitch + x];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int yy = y + dy - 1;
                if (yy < 0 || yy >= N) continue;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int xx = x + dx - 1;
                    if (xx < 0 || xx >= N) continue;
                    s[3 * dy + dx] += (double)ab[(long long)yy * pitch + xx] * gv;
                }
            }
        }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        double v = s[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        atomicAdd(acc + threadIdx.x, v);
    }
}

// coarse element = mean of its four children, summed in fp32 in row-major order, times 0.25
__global__ void __launch_bounds__(256) elem_coarsen_kernel(const float *a, float *ac, int n, int pitch, int pitch_c, int rows_c) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= pitch_c || r >= rows_c) return;
    const int nc = n / 2;
    float v = 0.0f;
    if (r < nc && c < nc) {
        const float *p0 = a + (long long)(2 * r) * pitch + 2 * c, *p1 = p0 + pitch;
        v = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(p0[0], p0[1]), p1[0]), p1[1]), 0.25f);
    }
    ac[(long long)r * pitch_c + c] = v;
}

// sumsq[b] = sum over interior nodes of r^2, fp64, deterministic: one partial per block (rows blockIdx.x, +grid, ...),
// then the last block adds the partials in order
__global__ void __launch_bounds__(256) interior_sumsq_kernel(const float *r, int N, int pitch, long long plane, double *partials,
                                                            unsigned int *counter, double *sumsq) {
    __shared__ double red[8];
    __shared__ int last;
    const int b = blockIdx.y;
    double acc = 0.0;
    for (int y = 1 + blockIdx.x; y < N - 1; y += gridDim.x) {
        const float *row = r + (long long)b * plane + (long long)y * pitch;
        for (int x = 1 + threadIdx.x; x < N - 1; x += 256) {
            const double v = (double)row[x];
            acc += v * v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        partials[(long long)b * gridDim.x + blockIdx.x] = s;
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        for (int bb = 0; bb < (int)gridDim.y; ++bb) {
            double s = 0.0;
            for (int i = 0; i < (int)gridDim.x; ++i) s += __ldcg(partials + (long long)bb * gridDim.x + i);
            sumsq[bb] = s;
        }
        *counter = 0u;
        __threadfence();
    }
}

}  // namespace mgfea
