// Backward pass of the table restriction / prolongation: the reference trains the 16-channel R / P
// kernels of FEANet/multigrid.py:50-73 by back-propagating through MultiGrid.iterate (multigrid.py:98-100, 145-185).
// Forward definitions (oracle/mgfea_oracle.c orc_restrict / orc_prolong_table):
//     fc[I][J] = s * sum_{a,c} R[key(y,x)][3a+c] * r[y][x],   y = 2I-1+a, x = 2J-1+c,   1 <= I,J <= Nc-2
//     e[y][x]  = s * sum_{a,c} P[keyc(I,J)][3a+c] * vc[I][J], y = 2I-1+a, x = 2J-1+c,   0 <= I,J < Nc
// Adjoints: the same index sets with the roles of the two grids swapped; weight gradients are per-key reductions.
// Training sizes are small (33^2 .. 129^2, batch <= 32): simple one-thread-per-node kernels, fp64 accumulation.
#pragma once
#include "mgfea_tile.cuh"

namespace mgfea {

struct AdjParams {
    int N, Nc, B, pitch, pitch_c, key_pitch, key_pitch_c, ntab;
    long long plane, plane_c;
    const unsigned char *keys, *keys_c;  // fine / coarse pattern keys (NULL: key 0)
    const float *tab;                    // [ntab][9]
    float scale;
    const float *fine;    // MODE 0: -            MODE 1: g_vf        MODE 2: r       MODE 3: g_vf
    const float *coarse;  // MODE 0: g_fc         MODE 1: -           MODE 2: g_fc    MODE 3: vc
    float *out;           // MODE 0: g_r (fine)   MODE 1: g_vc (coarse)
    double *acc;          // MODE 2 / 3: [16][9] weight gradient (the caller zeroes it)
};

// MODE 0: adjoint of the restriction w.r.t. its input  (out = fine field)
// MODE 1: adjoint of the prolongation w.r.t. its input (out = coarse field)
template <int MODE>
__global__ void __launch_bounds__(256) intergrid_adjoint_kernel(const AdjParams p) {
    const int b = blockIdx.z;
    if (MODE == 0) {
        const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
        if (x >= p.pitch || y >= p.N) return;
        float s = 0.0f;
        if (x < p.N) {
            const int k = (p.keys && p.ntab > 1) ? p.keys[(long long)y * p.key_pitch + x] : 0;
            const float *t = p.tab + 9 * k;
            double acc = 0.0;
            for (int a = 0; a < 3; ++a) {
                if ((y + 1 - a) & 1) continue;
                const int I = (y + 1 - a) / 2;
                if (I < 1 || I > p.Nc - 2) continue;
                for (int c = 0; c < 3; ++c) {
                    if ((x + 1 - c) & 1) continue;
                    const int J = (x + 1 - c) / 2;
                    if (J < 1 || J > p.Nc - 2) continue;
                    acc += (double)t[3 * a + c] * (double)p.coarse[(long long)b * p.plane_c + (long long)I * p.pitch_c + J];
                }
            }
            s = (float)(acc * (double)p.scale);
        }
        p.out[(long long)b * p.plane + (long long)y * p.pitch + x] = s;
    } else {
        const int J = blockIdx.x * 32 + threadIdx.x, I = blockIdx.y * 8 + threadIdx.y;
        if (J >= p.pitch_c || I >= p.Nc) return;
        float s = 0.0f;
        if (J < p.Nc) {
            const int k = (p.keys_c && p.ntab > 1) ? p.keys_c[(long long)I * p.key_pitch_c + J] : 0;
            const float *t = p.tab + 9 * k;
            double acc = 0.0;
            for (int a = 0; a < 3; ++a) {
                const int y = 2 * I - 1 + a;
                if (y < 0 || y >= p.N) continue;
                for (int c = 0; c < 3; ++c) {
                    const int x = 2 * J - 1 + c;
                    if (x < 0 || x >= p.N) continue;
                    acc += (double)t[3 * a + c] * (double)p.fine[(long long)b * p.plane + (long long)y * p.pitch + x];
                }
            }
            s = (float)(acc * (double)p.scale);
        }
        p.out[(long long)b * p.plane_c + (long long)I * p.pitch_c + J] = s;
    }
}

// MODE 2: gradient of the restriction table   acc[key(y,x)][3a+c]  += s * r[y][x] * g_fc[I][J]
// MODE 3: gradient of the prolongation table  acc[keyc(I,J)][3a+c] += s * vc[I][J] * g_vf[y][x]
// one thread per coarse node; per-block shared accumulators, then one atomicAdd per entry and block
template <int MODE>
__global__ void __launch_bounds__(256) intergrid_wgrad_kernel(const AdjParams p) {
    __shared__ double sacc[16 * 9];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < 144; i += 256) sacc[i] = 0.0;
    __syncthreads();
    const int b = blockIdx.z;
    const int J = blockIdx.x * 32 + threadIdx.x, I = blockIdx.y * 8 + threadIdx.y;
    const bool live = (MODE == 2) ? (I >= 1 && J >= 1 && I <= p.Nc - 2 && J <= p.Nc - 2) : (I < p.Nc && J < p.Nc);
    if (live) {
        const double cv = (double)p.coarse[(long long)b * p.plane_c + (long long)I * p.pitch_c + J] * (double)p.scale;
        const int kc = (MODE == 3 && p.keys_c && p.ntab > 1) ? p.keys_c[(long long)I * p.key_pitch_c + J] : 0;
        for (int a = 0; a < 3; ++a) {
            const int y = 2 * I - 1 + a;
            if (y < 0 || y >= p.N) continue;
            for (int c = 0; c < 3; ++c) {
                const int x = 2 * J - 1 + c;
                if (x < 0 || x >= p.N) continue;
                const double fv = (double)p.fine[(long long)b * p.plane + (long long)y * p.pitch + x];
                const int k = (MODE == 2) ? ((p.keys && p.ntab > 1) ? p.keys[(long long)y * p.key_pitch + x] : 0) : kc;
                atomicAdd(&sacc[9 * k + 3 * a + c], cv * fv);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < 9 * p.ntab; i += 256)
        if (sacc[i] != 0.0) atomicAdd(p.acc + i, sacc[i]);
}

}  // namespace mgfea
