// fp64 defect correction around the fp32 V-cycle (the reference's own remedy for the fp32 residual
// floor is to run everything in double, MM_poisson.ipynb cell 5 `.double()`).
//
// A V-cycle is the stationary iteration u <- u + B (f - K u); in exact arithmetic it is the same thing as
//     r = f - K u   (fp64)        e = V-cycle(zero guess, rhs r)   (fp32)        u += e   (fp64)
// so keeping u, f and the residual in fp64 and only the correction in fp32 reproduces the fp64 reference's residual
// history (the fp32 rounding of e is relative to e, which shrinks with the error) at fp32 V-cycle cost plus two
// streaming passes.  Weights are the fp32 tables promoted to double, exactly what `.double()` does to the reference's
// nn.Parameters.
//
//   mg_defect_f64_kernel   r32 = (float)(f64 - K u64) on interior nodes (0 on the ring), sum of r64^2 over the interior
//                          -> residual history / convergence control, same epilogue as the other norm kernels
//   mg_correct_f64_kernel  u64 += (double) e32 on interior nodes
#pragma once
#include "mgfea_tile.cuh"

namespace mgfea {

constexpr int F64_TX = 32, F64_TY = 8;  // thread block; every thread owns 2 adjacent columns: tile = 64 x 8 nodes

struct F64Params {
    int N, B, pitch;
    long long plane;
    const double *u;
    const double *f;
    float *r;
    const unsigned char *keys;
    int key_pitch, npat;
    const float *ktab;
    int nbx, nby;  // blocks per sample in x / y
    // row slab (multi-GPU): the arrays hold the global rows [row0, ..); rows [ylo, yhi) are computed (the owned rows
    // plus the 3 ghost rows per side the next down leg reads), the norm covers the owned rows [own0, own1) only.
    // Single GPU: row0 = 0, [ylo, yhi) = [own0, own1) = [0, N)
    int row0, ylo, yhi, own0, own1;
    double *partials;
    unsigned int *counter;
    double *sumsq;
    double *hist;
    void *ctl;
};

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool KEYS>
__global__ void __launch_bounds__(F64_TX *F64_TY) mg_defect_f64_kernel(const F64Params p) {
    __shared__ double tab[MAXPAT * 9];
    __shared__ double red[F64_TY];
    __shared__ int lastflag;
    const int tid = threadIdx.y * F64_TX + threadIdx.x;
    for (int i = tid; i < p.npat * 9; i += F64_TX * F64_TY) tab[i] = (double)p.ktab[i];
    const int solve_done = (p.ctl != nullptr) ? ld_volatile_s32(&reinterpret_cast<const Ctl *>(p.ctl)->done) : 0;
    __syncthreads();
    if (solve_done) return;
    const int N = p.N, b = blockIdx.z;
    const double *ub = p.u + (long long)b * p.plane - (long long)p.row0 * p.pitch;  // indexed by GLOBAL row below
    double part = 0.0;
    // persistent blocks: a block walks over the sample's 64 x 8 tiles with stride gridDim.x and reports ONE partial sum
    // (one fence + one ticket per block; with a block per tile the 33k tickets of a 4097^2 level dominated the kernel)
    const int ntile = p.nbx * p.nby;
    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int by = tile / p.nbx, bx = tile - by * p.nbx;
        const int x = (bx * F64_TX + threadIdx.x) * 2, y = p.ylo + by * F64_TY + threadIdx.y;
        if (y < p.yhi && x < p.pitch) {
            double r0 = 0.0, r1 = 0.0;
            const bool rin = (y >= 1 && y <= N - 2);
            if (rin && x <= N - 2) {  // at least one of the two columns may be interior
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int yy = y - 1 + d;  // 0 .. N-1
                    const double *row = ub + (long long)yy * p.pitch;
                    // columns x-1 .. x+2 (x is even; x-1 >= -1, x+2 <= pitch+1): guard the two outer ones
                    const double2 c = *reinterpret_cast<const double2 *>(row + x);
                    const double l = (x >= 1) ? row[x - 1] : 0.0;
                    const double rr = (x + 2 < p.pitch) ? row[x + 2] : 0.0;
                    const double v[4] = {l, c.x, c.y, rr};
                    int k[4] = {0, 0, 0, 0};
                    if (KEYS) {
                        const unsigned char *kr = p.keys + (long long)yy * p.key_pitch;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int xx = x - 1 + q;
                            k[q] = (xx >= 0 && xx < N) ? kr[xx] : 0;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        acc0 = fma(tab[9 * k[q] + 3 * d + q], v[q], acc0);
                        acc1 = fma(tab[9 * k[q + 1] + 3 * d + q], v[q + 1], acc1);
                    }
                }
                const double2 fv =
                    *reinterpret_cast<const double2 *>(p.f + (long long)b * p.plane + (long long)(y - p.row0) * p.pitch + x);
                if (x >= 1) r0 = fv.x - acc0;
                if (x + 1 <= N - 2) r1 = fv.y - acc1;
            }
            *reinterpret_cast<float2 *>(p.r + (long long)b * p.plane + (long long)(y - p.row0) * p.pitch + x) =
                make_float2((float)r0, (float)r1);
            if (y >= p.own0 && y < p.own1) part += r0 * r0 + r1 * r1;
        }
    }
    // ---- block partial -> per-sample sum by the last block (deterministic order), convergence control
    part = warp_sum_f64(part);
    if (threadIdx.x == 0) red[threadIdx.y] = part;
    __syncthreads();
    const int per = gridDim.x;  // partials per sample
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < F64_TY; ++w) s += red[w];
        p.partials[(long long)b * per + blockIdx.x] = s;
        __threadfence();
        const unsigned int ticket = atomicAdd(p.counter, 1u);
        lastflag = (ticket == (unsigned int)(per * p.B) - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!lastflag) return;
    __threadfence();
    Ctl *ctl = reinterpret_cast<Ctl *>(p.ctl);
    double tot = 0.0, mx = 0.0;
    for (int bb = 0; bb < p.B; ++bb) {
        double v = 0.0;
        for (int i = tid; i < per; i += F64_TX * F64_TY) v += __ldcg(p.partials + (long long)bb * per + i);
        v = warp_sum_f64(v);
        __syncthreads();
        if (threadIdx.x == 0) red[threadIdx.y] = v;
        __syncthreads();
        if (tid == 0) {
            double sum = 0.0;
            for (int w = 0; w < F64_TY; ++w) sum += red[w];
            if (p.sumsq) p.sumsq[bb] = sum;
            if (ctl && p.hist && ctl->cycle < ctl->max_cycles) p.hist[(long long)ctl->cycle * p.B + bb] = sum;
            tot += sum;
            mx = sum > mx ? sum : mx;
        }
    }
    if (tid == 0) {
        if (ctl) {
            const int cyc = ctl->cycle + 1;
            ctl->cycle = cyc;
            const double metric = (ctl->conv_rule == 1) ? mx : tot;
            bool done = false;
            if (ctl->eps2 >= 0.0 && cyc >= ctl->min_cycles && metric <= ctl->eps2) done = true;
            if (cyc >= ctl->max_cycles) done = true;
            if (!(metric == metric) || metric > 1.7e308) done = true;
            if (done) ctl->done = 1;
        }
        *p.counter = 0u;
        __threadfence();
    }
}

// rows [own0, own1) of arrays that hold the global rows [row0, ..) (single GPU: 0, N, 0)
__global__ void __launch_bounds__(256) mg_correct_f64_kernel(double *u, const float *e, int N, int pitch, long long plane,
                                                            int B, const void *ctl, int own0, int own1, int row0) {
    if (ctl != nullptr && ld_volatile_s32(&reinterpret_cast<const Ctl *>(ctl)->done) != 0) return;
    const int half = pitch >> 1;  // column pairs per row
    const int nown = own1 - own0;
    const long long total = (long long)B * nown * half;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xp = (int)(i % half);
        const long long ry = i / half;
        const int y = own0 + (int)(ry % nown);
        const long long b = ry / nown;
        if (y < 1 || y > N - 2) continue;
        const int x = 2 * xp;
        if (x > N - 2) continue;
        const long long o = b * plane + (long long)(y - row0) * pitch + x;
        const float2 ev = *reinterpret_cast<const float2 *>(e + o);
        double2 uv = *reinterpret_cast<double2 *>(u + o);
        if (x >= 1) uv.x += (double)ev.x;
        if (x + 1 <= N - 2) uv.y += (double)ev.y;
        *reinterpret_cast<double2 *>(u + o) = uv;
    }
}

// padded fp32 field -> padded fp64 field (same pitch / plane counts); zero_ring: the Dirichlet ring of the iterate is
// cleared on the way (the reference's first reset_boundary)
__global__ void __launch_bounds__(256) mg_widen_f64_kernel(const float *src, double *dst, int N, int pitch, long long plane,
                                                          int B, int zero_ring) {
    const int half = pitch >> 1;
    const long long total = (long long)B * N * half;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xp = (int)(i % half);
        const long long ry = i / half;
        const int y = (int)(ry % N), x = 2 * xp;
        const long long o = (ry / N) * plane + (long long)y * pitch + x;
        const float2 v = *reinterpret_cast<const float2 *>(src + o);
        double2 w = make_double2((double)v.x, (double)v.y);
        if (x >= N) w.x = 0.0;
        if (x + 1 >= N) w.y = 0.0;
        if (zero_ring) {
            const bool rr = (y == 0 || y == N - 1);
            if (rr || x == 0 || x == N - 1) w.x = 0.0;
            if (rr || x + 1 == N - 1) w.y = 0.0;
        }
        *reinterpret_cast<double2 *>(dst + o) = w;
    }
}

}  // namespace mgfea
