// Latency-oriented kernels for the MID levels of the iso V(1,1) cycle (129 <= N <= ~1025, single pattern, default
// Dirichlet ring).  These levels are L2-resident and far too small to be bandwidth bound: in the streaming kernels a
// warp walks its strip row by row, a serial chain of ~8 row steps (~1 us each) per launch, eight launches per cycle.
// Here every thread owns ONE group of 4 columns of ONE row of a 32 x 32 tile, the tile (+ halo) sits in shared memory,
// and a leg is two barrier-separated stages, so the critical path of a launch is two short stencil evaluations.
//
//   MODE 0  down leg from a ZERO guess (every coarse level of a V-cycle): u1 = mask(inv * f) is pointwise, so
//           stage A evaluates r = f - K u1 straight from the f tile (and stores u1), stage B restricts r -> coarse f.
//   MODE 1  up leg: stage A forms uc = mask(u + mask(bilinear P v_c)) in shared memory, stage B is the Jacobi sweep.
//
// Arithmetic: the canonical order shared with the oracle and the other kernels (row-major 9-tap FMA chain, separate
// sub / mul / add in the Jacobi update, ATen's bilinear forms), so results are bit-identical to the streaming / tile path.
#pragma once
#include "mgfea_tile.cuh"

namespace mgfea {

constexpr int MID_T = 32;         // tile edge (fine nodes)
constexpr int MID_G = 10;         // float4 groups per staged row: columns x0-4 .. x0+35
constexpr int MID_S = 48;         // staged row pitch in floats: 4 pad + 40 + 4 pad
constexpr int MID_THREADS = 352;  // >= 34 rows x 10 groups (stage A items)
constexpr int MID_CS = 32;        // coarse tile row pitch: 4 pad + 24 + 4 pad

struct MidParams {
    int N, B, pitch;
    long long plane;
    int nt;  // tiles per edge
    float inv_per, inv_nt;
    const float *u_in;  // up leg: iterate before the correction
    float *u_out;
    const float *f;
    const float *ktab, *invd;
    // down leg
    float *fc;
    int Nc, pitch_c;
    long long plane_c;
    const float *rtab;
    int r_has_scale;
    float r_scale;
    const float *r_scale_dev;
    // up leg
    const float *vc;
    int prolong_seq;
    void *ctl;
    // two-phase meshes / table transfer operators (mg_midk_kernel)
    const unsigned char *keys;    // [N][key_pitch] pattern keys of this level (NULL: single pattern)
    int key_pitch, npat;
    int rtab_n;                   // restriction tables: 1, or one per pattern (by the FINE source node's key)
    int prolong_mode;             // MGFEA_PROLONG_BILINEAR (1) | MGFEA_PROLONG_TABLE (3)
    const float *ptab;            // [ptab_n][9]
    int ptab_n, p_has_scale;
    float p_scale;
    const float *p_scale_dev;
    const unsigned char *keys_c;  // coarse level's key map (table prolongation, ptab_n > 1)
    int key_pitch_c;
};

// exact t / d for 0 <= t < 2^24 (float estimate + one correction step)
__device__ __forceinline__ int mid_div(int t, int d, float inv) {
    int q = __float2int_rz(__int2float_rn(t) * inv);
    const int r = t - q * d;
    if (r < 0) --q;
    else if (r >= d) ++q;
    return q;
}

__device__ __forceinline__ float mid_stencil1(const float (&w)[9], const float *t, const float *m, const float *b) {
    float s = __fmul_rn(w[0], t[0]);
    s = __fmaf_rn(w[1], t[1], s);
    s = __fmaf_rn(w[2], t[2], s);
    s = __fmaf_rn(w[3], m[0], s);
    s = __fmaf_rn(w[4], m[1], s);
    s = __fmaf_rn(w[5], m[2], s);
    s = __fmaf_rn(w[6], b[0], s);
    s = __fmaf_rn(w[7], b[1], s);
    s = __fmaf_rn(w[8], b[2], s);
    return s;
}

// 6 staged values (columns 4g-1 .. 4g+4 of the staged row) around group g
__device__ __forceinline__ void mid_row6(const float *row, int g, float (&a)[6]) {
    const float *p = row + 4 + 4 * g;
    const float4 v = *reinterpret_cast<const float4 *>(p);
    a[0] = p[-1];
    a[1] = v.x;
    a[2] = v.y;
    a[3] = v.z;
    a[4] = v.w;
    a[5] = p[4];
}

template <int MODE>
__global__ void __launch_bounds__(MID_THREADS) mg_mid_kernel(const MidParams p) {
    // MODE 0: A = f rows y0-2..y0+33, Bf = r rows y0-1..y0+32.  MODE 1: A = u rows y0-1..y0+32, Bf = coarse tile.
    __shared__ __align__(16) float A[36 * MID_S];
    __shared__ __align__(16) float Bf[34 * MID_S];
    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int N = p.N;
    float kw[9], rw[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        kw[q] = p.ktab[q];
        rw[q] = (MODE == 0) ? p.rtab[q] : 0.0f;
    }
    const float inv = p.invd[0];
    const float rscale = (MODE == 0 && p.r_has_scale) ? (p.r_scale_dev ? *p.r_scale_dev : p.r_scale) : 1.0f;
    // tile coordinates
    int t = blockIdx.x, b = 0;
    const int per = p.nt * p.nt;
    if (p.B > 1) {
        b = mid_div(t, per, p.inv_per);
        t -= b * per;
    }
    const int ty = mid_div(t, p.nt, p.inv_nt), tx = t - ty * p.nt;
    const int y0 = ty * MID_T, x0 = tx * MID_T;
    // zero the pad columns once (they are read as the out-of-box neighbours of groups 0 and 9; values never matter
    // for a stored result, but they must be finite)
    for (int i = tid; i < 36 * 2; i += MID_THREADS) {
        const int r = i >> 1, side = i & 1;
        *reinterpret_cast<float4 *>(A + r * MID_S + (side ? 44 : 0)) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 0 && r < 34)  // (MODE 1 keeps the coarse tile in Bf, with its own pitch)
            *reinterpret_cast<float4 *>(Bf + r * MID_S + (side ? 44 : 0)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pdl_wait();  // weights above are never written by a kernel; field data is touched only after this point
    // requested now, tested after the tile loads have been issued (one memory round trip instead of two)
    const int solve_done = (p.ctl != nullptr) ? ld_volatile_s32(&reinterpret_cast<const Ctl *>(p.ctl)->done) : 0;

    const float *fb = p.f + (long long)b * p.plane;
    float *uo = p.u_out + (long long)b * p.plane;

    if (MODE == 0) {
        // ---- stage the f tile: rows y0-2 .. y0+33, zero outside the domain
        for (int i = tid; i < 36 * MID_G; i += MID_THREADS) {
            const int r = i / MID_G, g = i - r * MID_G;
            const int y = y0 - 2 + r, x = x0 - 4 + 4 * g;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < N && x >= 0 && x + 3 < p.pitch) v = __ldcg(reinterpret_cast<const float4 *>(fb + (long long)y * p.pitch + x));
            *reinterpret_cast<float4 *>(A + r * MID_S + 4 + 4 * g) = v;
        }
        if (solve_done) return;
        __syncthreads();
        // ---- stage A: r = f - K u1 with u1 = mask(inv * f) on rows y0-1 .. y0+32; store u1 of the tile
        if (tid < 34 * MID_G) {
            const int r = tid / MID_G, g = tid - r * MID_G;  // r: row y0-1+r
            const int y = y0 - 1 + r, x = x0 - 4 + 4 * g;
            float u1[3][6];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float a[6];
                mid_row6(A + (r + d) * MID_S, g, a);
                const int yy = y - 1 + d;
                const bool rin = (yy >= 1 && yy <= N - 2);
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const int xx = x - 1 + q;
                    // Jacobi from a zero guess: K 0 = 0, so u1 = inv * (f - 0) + 0 on interior nodes, 0 elsewhere
                    u1[d][q] = (rin && xx >= 1 && xx <= N - 2) ? __fadd_rn(__fmul_rn(inv, a[q]), 0.0f) : 0.0f;
                }
            }
            const float4 fv = *reinterpret_cast<const float4 *>(A + (r + 1) * MID_S + 4 + 4 * g);
            const float ff[4] = {fv.x, fv.y, fv.z, fv.w};
            float rr[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) rr[e] = __fsub_rn(ff[e], mid_stencil1(kw, &u1[0][e], &u1[1][e], &u1[2][e]));
            *reinterpret_cast<float4 *>(Bf + r * MID_S + 4 + 4 * g) = make_float4(rr[0], rr[1], rr[2], rr[3]);
            if (g >= 1 && g <= 8 && r >= 1 && r <= 32 && y < N && x < N)
                st_global_v4(uo + (long long)y * p.pitch + x, make_float4(u1[1][1], u1[1][2], u1[1][3], u1[1][4]));
        }
        __syncthreads();
        // ---- stage B: coarse rhs of the 16 x 16 coarse nodes under this tile
        if (tid < 256) {
            const int ci = tid >> 4, cj = tid & 15;
            const int I = (y0 >> 1) + ci, J = (x0 >> 1) + cj;
            if (I < p.Nc && J < p.Nc) {
                float out = 0.0f;
                if (I >= 1 && I <= p.Nc - 2 && J >= 1 && J <= p.Nc - 2) {
                    // fine rows 2I-1 .. 2I+1 -> staged r rows (2I-1) - (y0-1) = 2ci .. ; columns 2J-1 .. -> 4 + 4 + 2cj - 1
                    const float *r0 = Bf + (2 * ci) * MID_S + 4 + 4 + 2 * cj - 1;
                    float s = mid_stencil1(rw, r0, r0 + MID_S, r0 + 2 * MID_S);
                    out = p.r_has_scale ? __fmul_rn(rscale, s) : s;
                }
                p.fc[(long long)b * p.plane_c + (long long)I * p.pitch_c + J] = out;
            }
        }
    } else {
        float *VC = Bf;  // coarse rows y0/2-1 .. y0/2+16 (18), coarse columns x0/2-4 .. x0/2+19 (24 = 6 groups)
        const float *ub = p.u_in + (long long)b * p.plane;
        const float *cb = p.vc + (long long)b * p.plane_c;
        const int cy0 = (y0 >> 1) - 1, cx0 = (x0 >> 1) - 4;
        for (int i = tid; i < 34 * MID_G + 18 * 6; i += MID_THREADS) {
            if (i < 34 * MID_G) {
                const int r = i / MID_G, g = i - r * MID_G;
                const int y = y0 - 1 + r, x = x0 - 4 + 4 * g;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (y >= 0 && y < N && x >= 0 && x + 3 < p.pitch) v = __ldcg(reinterpret_cast<const float4 *>(ub + (long long)y * p.pitch + x));
                *reinterpret_cast<float4 *>(A + r * MID_S + 4 + 4 * g) = v;
            } else {
                const int k = i - 34 * MID_G;
                const int r = k / 6, g = k - r * 6;
                const int I = cy0 + r, J = cx0 + 4 * g;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (I >= 0 && I < p.Nc && J >= 0 && J + 3 < p.pitch_c) v = __ldcg(reinterpret_cast<const float4 *>(cb + (long long)I * p.pitch_c + J));
                *reinterpret_cast<float4 *>(VC + r * MID_CS + 4 + 4 * g) = v;
            }
        }
        // this thread's f values for stage B (rows y0 .. y0+31, groups 1..8), issued before the barrier
        const int rB = tid >> 3, gB = 1 + (tid & 7);
        const int yB = y0 + rB, xB = x0 - 4 + 4 * gB;
        const bool actB = (tid < 256) && yB < N && xB < N;
        float4 fB = make_float4(0.f, 0.f, 0.f, 0.f);
        if (actB) fB = __ldcg(reinterpret_cast<const float4 *>(fb + (long long)yB * p.pitch + xB));
        if (solve_done) return;
        __syncthreads();
        // ---- stage A: uc = mask(u + mask(P vc)) in place, rows y0-1 .. y0+32
        if (tid < 34 * MID_G) {
            const int r = tid / MID_G, g = tid - r * MID_G;
            const int y = y0 - 1 + r, x = x0 - 4 + 4 * g;
            float4 *cell = reinterpret_cast<float4 *>(A + r * MID_S + 4 + 4 * g);
            const float4 uv = *cell;
            const float ua[4] = {uv.x, uv.y, uv.z, uv.w};
            float o[4];
            const bool rin = (y >= 1 && y <= N - 2);
            // coarse values: rows I = y>>1 (and I+1 for odd y); columns (x>>1) .. (x>>1)+2
            const int Il = (y >> 1) - cy0;                     // staged coarse row of floor(y/2)
            const float *c0 = VC + Il * MID_CS + 4 + ((x >> 1) - cx0);
            const float *c1 = c0 + MID_CS;
            const float t0 = c0[0], t1 = c0[1], t2 = c0[2];
            float e[4];
            if ((y & 1) == 0) {
                e[0] = t0;
                e[1] = __fadd_rn(__fmul_rn(0.5f, t0), __fmul_rn(0.5f, t1));
                e[2] = t1;
                e[3] = __fadd_rn(__fmul_rn(0.5f, t1), __fmul_rn(0.5f, t2));
            } else {
                const float b0 = c1[0], b1 = c1[1], b2 = c1[2];
                e[0] = __fadd_rn(__fmul_rn(0.5f, t0), __fmul_rn(0.5f, b0));
                e[2] = __fadd_rn(__fmul_rn(0.5f, t1), __fmul_rn(0.5f, b1));
                if (p.prolong_seq) {
                    float v = __fadd_rn(__fmul_rn(0.25f, t0), __fmul_rn(0.25f, t1));
                    v = __fadd_rn(v, __fmul_rn(0.25f, b0));
                    e[1] = __fadd_rn(v, __fmul_rn(0.25f, b1));
                    v = __fadd_rn(__fmul_rn(0.25f, t1), __fmul_rn(0.25f, t2));
                    v = __fadd_rn(v, __fmul_rn(0.25f, b1));
                    e[3] = __fadd_rn(v, __fmul_rn(0.25f, b2));
                } else {
                    const float ta = __fadd_rn(__fmul_rn(0.5f, t0), __fmul_rn(0.5f, t1));
                    const float ba = __fadd_rn(__fmul_rn(0.5f, b0), __fmul_rn(0.5f, b1));
                    e[1] = __fadd_rn(__fmul_rn(0.5f, ta), __fmul_rn(0.5f, ba));
                    const float tb = __fadd_rn(__fmul_rn(0.5f, t1), __fmul_rn(0.5f, t2));
                    const float bb = __fadd_rn(__fmul_rn(0.5f, b1), __fmul_rn(0.5f, b2));
                    e[3] = __fadd_rn(__fmul_rn(0.5f, tb), __fmul_rn(0.5f, bb));
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int xx = x + q;
                o[q] = (rin && xx >= 1 && xx <= N - 2) ? __fadd_rn(ua[q], e[q]) : 0.0f;
            }
            *cell = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        // ---- stage B: Jacobi sweep on the tile
        if (actB) {
            float a[3][6];
#pragma unroll
            for (int d = 0; d < 3; ++d) mid_row6(A + (rB + d) * MID_S, gB, a[d]);  // staged row rB+1 is fine row yB
            const float ff[4] = {fB.x, fB.y, fB.z, fB.w};
            float o[4];
            const bool rin = (yB >= 1 && yB <= N - 2);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float ku = mid_stencil1(kw, &a[0][e], &a[1][e], &a[2][e]);
                const float v = __fadd_rn(__fmul_rn(inv, __fsub_rn(ff[e], ku)), a[1][e + 1]);
                const int xx = xB + e;
                o[e] = (rin && xx >= 1 && xx <= N - 2) ? v : 0.0f;
            }
            st_global_v4(uo + (long long)yB * p.pitch + xB, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same two legs for TWO-PHASE meshes and / or table transfer operators (source-key indexed stiffness stencil,
// restriction table by the fine source node's key, ConvTranspose prolongation table by the coarse node's key): the
// levels 129^2 .. 1025^2 of BASELINE config 3 with the Jacobi smoother.  The generic tile kernel needs ~20 us for each of
// these small launches, nearly all of it instruction fetch (ncu no_instruction 5-6 stalled warps per issue,
// profiles/r02_ncu_tile_cfg3.json); this kernel is a few KB.  Key tiles ride along with the field tiles (bytes, same
// column layout); every weight is looked up in shared-memory tables by the SOURCE node's key, in the canonical
// row-major chain order (oracle: stencil_src, orc_restrict, orc_prolong_table).
__device__ __forceinline__ float midk_stencil1(const float *tab, const int (&kt)[3], const int (&km)[3], const int (&kb)[3],
                                               const float *t, const float *m, const float *b) {
    float s = __fmul_rn(tab[9 * kt[0] + 0], t[0]);
    s = __fmaf_rn(tab[9 * kt[1] + 1], t[1], s);
    s = __fmaf_rn(tab[9 * kt[2] + 2], t[2], s);
    s = __fmaf_rn(tab[9 * km[0] + 3], m[0], s);
    s = __fmaf_rn(tab[9 * km[1] + 4], m[1], s);
    s = __fmaf_rn(tab[9 * km[2] + 5], m[2], s);
    s = __fmaf_rn(tab[9 * kb[0] + 6], b[0], s);
    s = __fmaf_rn(tab[9 * kb[1] + 7], b[1], s);
    s = __fmaf_rn(tab[9 * kb[2] + 8], b[2], s);
    return s;
}
// 6 staged keys (columns 4g-1 .. 4g+4 of the staged key row; bytes, same column layout as the float tiles)
__device__ __forceinline__ void midk_key6(const unsigned char *row, int g, int (&k)[6]) {
    const unsigned char *p = row + 4 + 4 * g;
#pragma unroll
    for (int q = 0; q < 6; ++q) k[q] = p[q - 1];
}

template <int MODE>
__global__ void __launch_bounds__(MID_THREADS) mg_midk_kernel(const MidParams p) {
    __shared__ __align__(16) float A[36 * MID_S];
    __shared__ __align__(16) float Bf[34 * MID_S];
    __shared__ __align__(16) unsigned char KA[36 * MID_S];   // fine keys, rows y0-2 .. y0+33 (MODE 1 uses rows 1 .. 34)
    __shared__ __align__(16) unsigned char KC[18 * MID_CS];  // coarse keys (MODE 1, table prolongation)
    __shared__ float s_k[MAXPAT * 9], s_inv[MAXPAT], s_t[MAXPAT * 9];  // stiffness, omega/d, restriction | prolongation
    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int N = p.N;
    const bool fkeys = (p.keys != nullptr);
    const int ntab = (MODE == 0) ? p.rtab_n : p.ptab_n;
    const float *gtab = (MODE == 0) ? p.rtab : p.ptab;
    for (int i = tid; i < MAXPAT * 9; i += MID_THREADS) {
        s_k[i] = (i < p.npat * 9) ? p.ktab[i] : 0.0f;
        s_t[i] = (gtab != nullptr && i < ntab * 9) ? gtab[i] : 0.0f;
    }
    if (tid < MAXPAT) s_inv[tid] = (tid < p.npat) ? p.invd[tid] : 0.0f;
    const float tscale = (MODE == 0) ? (p.r_has_scale ? (p.r_scale_dev ? *p.r_scale_dev : p.r_scale) : 1.0f)
                                     : (p.p_has_scale ? (p.p_scale_dev ? *p.p_scale_dev : p.p_scale) : 1.0f);
    const bool has_scale = (MODE == 0) ? (p.r_has_scale != 0) : (p.p_has_scale != 0);
    int t = blockIdx.x, b = 0;
    const int per = p.nt * p.nt;
    if (p.B > 1) {
        b = mid_div(t, per, p.inv_per);
        t -= b * per;
    }
    const int ty = mid_div(t, p.nt, p.inv_nt), tx = t - ty * p.nt;
    const int y0 = ty * MID_T, x0 = tx * MID_T;
    for (int i = tid; i < 36 * 2; i += MID_THREADS) {
        const int r = i >> 1, side = i & 1;
        *reinterpret_cast<float4 *>(A + r * MID_S + (side ? 44 : 0)) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 0 && r < 34) *reinterpret_cast<float4 *>(Bf + r * MID_S + (side ? 44 : 0)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // key tiles (never written by a kernel): fine keys of rows y0-2 .. y0+33, columns x0-4 .. x0+43 incl. the pads
    for (int i = tid; i < 36 * (MID_S / 4); i += MID_THREADS) {
        const int r = i / (MID_S / 4), c4 = i - r * (MID_S / 4);
        const int y = y0 - 2 + r, x = x0 - 8 + 4 * c4;  // staged column 4*c4 is fine column x0 - 8 + 4*c4
        unsigned int w = 0u;
        if (fkeys && y >= 0 && y < N && x >= 0 && x < N && x + 3 < p.key_pitch) {
            w = __ldg(reinterpret_cast<const unsigned int *>(p.keys + (long long)y * p.key_pitch + x));
            if (x + 3 >= N) w &= 0xffffffffu >> (8 * (x + 4 - N));  // columns >= N: key 0 (padding bytes are not ours)
        }
        *reinterpret_cast<unsigned int *>(KA + r * MID_S + 4 * c4) = w;
    }
    const bool pkeys = (MODE == 1) && p.prolong_mode == 3 && p.ptab_n > 1 && p.keys_c != nullptr;
    if (MODE == 1) {
        const int cy0k = (y0 >> 1) - 1, cx0k = (x0 >> 1) - 4;
        for (int i = tid; i < 18 * (MID_CS / 4); i += MID_THREADS) {
            const int r = i / (MID_CS / 4), c4 = i - r * (MID_CS / 4);
            const int I = cy0k + r, J = cx0k - 4 + 4 * c4;  // staged column 4 + j is coarse column cx0 + j
            unsigned int w = 0u;
            if (pkeys && I >= 0 && I < p.Nc && J >= 0 && J < p.Nc && J + 3 < p.key_pitch_c) {
                w = __ldg(reinterpret_cast<const unsigned int *>(p.keys_c + (long long)I * p.key_pitch_c + J));
                if (J + 3 >= p.Nc) w &= 0xffffffffu >> (8 * (J + 4 - p.Nc));
            }
            *reinterpret_cast<unsigned int *>(KC + r * MID_CS + 4 * c4) = w;
        }
    }
    pdl_wait();
    const int solve_done = (p.ctl != nullptr) ? ld_volatile_s32(&reinterpret_cast<const Ctl *>(p.ctl)->done) : 0;
    const float *fb = p.f + (long long)b * p.plane;
    float *uo = p.u_out + (long long)b * p.plane;

    if (MODE == 0) {
        for (int i = tid; i < 36 * MID_G; i += MID_THREADS) {
            const int r = i / MID_G, g = i - r * MID_G;
            const int y = y0 - 2 + r, x = x0 - 4 + 4 * g;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < N && x >= 0 && x + 3 < p.pitch) v = __ldcg(reinterpret_cast<const float4 *>(fb + (long long)y * p.pitch + x));
            *reinterpret_cast<float4 *>(A + r * MID_S + 4 + 4 * g) = v;
        }
        if (solve_done) return;
        __syncthreads();
        // ---- stage A: r = f - K u1 with u1 = mask(inv[key] * f) on rows y0-1 .. y0+32; store u1 of the tile
        if (tid < 34 * MID_G) {
            const int r = tid / MID_G, g = tid - r * MID_G;
            const int y = y0 - 1 + r, x = x0 - 4 + 4 * g;
            float u1[3][6];
            int kk[3][6];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float a[6];
                mid_row6(A + (r + d) * MID_S, g, a);
                midk_key6(KA + (r + d) * MID_S, g, kk[d]);
                const int yy = y - 1 + d;
                const bool rin = (yy >= 1 && yy <= N - 2);
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const int xx = x - 1 + q;
                    u1[d][q] = (rin && xx >= 1 && xx <= N - 2) ? __fadd_rn(__fmul_rn(s_inv[kk[d][q]], a[q]), 0.0f) : 0.0f;
                }
            }
            const float4 fv = *reinterpret_cast<const float4 *>(A + (r + 1) * MID_S + 4 + 4 * g);
            const float ff[4] = {fv.x, fv.y, fv.z, fv.w};
            float rr[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kt[3] = {kk[0][e], kk[0][e + 1], kk[0][e + 2]}, km[3] = {kk[1][e], kk[1][e + 1], kk[1][e + 2]},
                          kb[3] = {kk[2][e], kk[2][e + 1], kk[2][e + 2]};
                rr[e] = __fsub_rn(ff[e], midk_stencil1(s_k, kt, km, kb, &u1[0][e], &u1[1][e], &u1[2][e]));
            }
            *reinterpret_cast<float4 *>(Bf + r * MID_S + 4 + 4 * g) = make_float4(rr[0], rr[1], rr[2], rr[3]);
            if (g >= 1 && g <= 8 && r >= 1 && r <= 32 && y < N && x < N)
                st_global_v4(uo + (long long)y * p.pitch + x, make_float4(u1[1][1], u1[1][2], u1[1][3], u1[1][4]));
        }
        __syncthreads();
        // ---- stage B: coarse rhs; the restriction table follows the FINE source node's key
        if (tid < 256) {
            const int ci = tid >> 4, cj = tid & 15;
            const int I = (y0 >> 1) + ci, J = (x0 >> 1) + cj;
            if (I < p.Nc && J < p.Nc) {
                float out = 0.0f;
                if (I >= 1 && I <= p.Nc - 2 && J >= 1 && J <= p.Nc - 2) {
                    const int co = 4 + 4 + 2 * cj - 1;
                    const float *r0 = Bf + (2 * ci) * MID_S + co;
                    const unsigned char *k0 = KA + (2 * ci + 1) * MID_S + co;  // staged key row r+1 is fine row y0-1+r
                    int kt[3], km[3], kb[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        kt[q] = (ntab > 1) ? k0[q] : 0;
                        km[q] = (ntab > 1) ? k0[MID_S + q] : 0;
                        kb[q] = (ntab > 1) ? k0[2 * MID_S + q] : 0;
                    }
                    const float s = midk_stencil1(s_t, kt, km, kb, r0, r0 + MID_S, r0 + 2 * MID_S);
                    out = has_scale ? __fmul_rn(tscale, s) : s;
                }
                p.fc[(long long)b * p.plane_c + (long long)I * p.pitch_c + J] = out;
            }
        }
    } else {
        float *VC = Bf;
        const float *ub = p.u_in + (long long)b * p.plane;
        const float *cb = p.vc + (long long)b * p.plane_c;
        const int cy0 = (y0 >> 1) - 1, cx0 = (x0 >> 1) - 4;
        for (int i = tid; i < 34 * MID_G + 18 * 6; i += MID_THREADS) {
            if (i < 34 * MID_G) {
                const int r = i / MID_G, g = i - r * MID_G;
                const int y = y0 - 1 + r, x = x0 - 4 + 4 * g;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (y >= 0 && y < N && x >= 0 && x + 3 < p.pitch) v = __ldcg(reinterpret_cast<const float4 *>(ub + (long long)y * p.pitch + x));
                *reinterpret_cast<float4 *>(A + r * MID_S + 4 + 4 * g) = v;
            } else {
                const int k = i - 34 * MID_G;
                const int r = k / 6, g = k - r * 6;
                const int I = cy0 + r, J = cx0 + 4 * g;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (I >= 0 && I < p.Nc && J >= 0 && J + 3 < p.pitch_c) v = __ldcg(reinterpret_cast<const float4 *>(cb + (long long)I * p.pitch_c + J));
                *reinterpret_cast<float4 *>(VC + r * MID_CS + 4 + 4 * g) = v;
            }
        }
        const int rB = tid >> 3, gB = 1 + (tid & 7);
        const int yB = y0 + rB, xB = x0 - 4 + 4 * gB;
        const bool actB = (tid < 256) && yB < N && xB < N;
        float4 fB = make_float4(0.f, 0.f, 0.f, 0.f);
        if (actB) fB = __ldcg(reinterpret_cast<const float4 *>(fb + (long long)yB * p.pitch + xB));
        if (solve_done) return;
        __syncthreads();
        // ---- stage A: uc = mask(u + P vc) in place, rows y0-1 .. y0+32 (bilinear: ATen's forms; table: ConvTranspose taps)
        if (tid < 34 * MID_G) {
            const int r = tid / MID_G, g = tid - r * MID_G;
            const int y = y0 - 1 + r, x = x0 - 4 + 4 * g;
            float4 *cell = reinterpret_cast<float4 *>(A + r * MID_S + 4 + 4 * g);
            const float4 uv = *cell;
            const float ua[4] = {uv.x, uv.y, uv.z, uv.w};
            float o[4];
            const bool rin = (y >= 1 && y <= N - 2);
            const int Il = (y >> 1) - cy0;
            const int cofs = Il * MID_CS + 4 + ((x >> 1) - cx0);
            const float *c0 = VC + cofs;
            const float *c1 = c0 + MID_CS;
            const float top[3] = {c0[0], c0[1], c0[2]};
            const float bot[3] = {c1[0], c1[1], c1[2]};
            const bool odd = (y & 1) != 0;
            float e[4];
            if (p.prolong_mode == 3) {
                int kt[3], kb[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    kt[q] = pkeys ? KC[cofs + q] : 0;
                    kb[q] = pkeys ? KC[cofs + MID_CS + q] : 0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int jl = q >> 1;
                    float s;
                    if (!odd) {
                        if (!(q & 1)) {
                            s = __fmul_rn(s_t[9 * kt[jl] + 4], top[jl]);
                        } else {
                            s = __fmul_rn(s_t[9 * kt[jl + 1] + 3], top[jl + 1]);
                            s = __fmaf_rn(s_t[9 * kt[jl] + 5], top[jl], s);
                        }
                    } else {
                        if (!(q & 1)) {
                            s = __fmul_rn(s_t[9 * kb[jl] + 1], bot[jl]);
                            s = __fmaf_rn(s_t[9 * kt[jl] + 7], top[jl], s);
                        } else {
                            s = __fmul_rn(s_t[9 * kb[jl + 1] + 0], bot[jl + 1]);
                            s = __fmaf_rn(s_t[9 * kb[jl] + 2], bot[jl], s);
                            s = __fmaf_rn(s_t[9 * kt[jl + 1] + 6], top[jl + 1], s);
                            s = __fmaf_rn(s_t[9 * kt[jl] + 8], top[jl], s);
                        }
                    }
                    e[q] = has_scale ? __fmul_rn(tscale, s) : s;
                }
            } else {
                const float t0 = top[0], t1 = top[1], t2 = top[2];
                if (!odd) {
                    e[0] = t0;
                    e[1] = __fadd_rn(__fmul_rn(0.5f, t0), __fmul_rn(0.5f, t1));
                    e[2] = t1;
                    e[3] = __fadd_rn(__fmul_rn(0.5f, t1), __fmul_rn(0.5f, t2));
                } else {
                    const float b0 = bot[0], b1 = bot[1], b2 = bot[2];
                    e[0] = __fadd_rn(__fmul_rn(0.5f, t0), __fmul_rn(0.5f, b0));
                    e[2] = __fadd_rn(__fmul_rn(0.5f, t1), __fmul_rn(0.5f, b1));
                    const float ta = __fadd_rn(__fmul_rn(0.5f, t0), __fmul_rn(0.5f, t1));
                    const float ba = __fadd_rn(__fmul_rn(0.5f, b0), __fmul_rn(0.5f, b1));
                    e[1] = __fadd_rn(__fmul_rn(0.5f, ta), __fmul_rn(0.5f, ba));
                    const float tb = __fadd_rn(__fmul_rn(0.5f, t1), __fmul_rn(0.5f, t2));
                    const float bb = __fadd_rn(__fmul_rn(0.5f, b1), __fmul_rn(0.5f, b2));
                    e[3] = __fadd_rn(__fmul_rn(0.5f, tb), __fmul_rn(0.5f, bb));
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int xx = x + q;
                o[q] = (rin && xx >= 1 && xx <= N - 2) ? __fadd_rn(ua[q], e[q]) : 0.0f;
            }
            *cell = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        // ---- stage B: Jacobi sweep on the tile, weights by the source node's key
        if (actB) {
            float a[3][6];
            int kk[3][6];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                mid_row6(A + (rB + d) * MID_S, gB, a[d]);           // staged row rB+1 is fine row yB
                midk_key6(KA + (rB + d + 1) * MID_S, gB, kk[d]);     // key rows start one fine row earlier
            }
            const float ff[4] = {fB.x, fB.y, fB.z, fB.w};
            float o[4];
            const bool rin = (yB >= 1 && yB <= N - 2);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kt[3] = {kk[0][e], kk[0][e + 1], kk[0][e + 2]}, km[3] = {kk[1][e], kk[1][e + 1], kk[1][e + 2]},
                          kb[3] = {kk[2][e], kk[2][e + 1], kk[2][e + 2]};
                const float ku = midk_stencil1(s_k, kt, km, kb, &a[0][e], &a[1][e], &a[2][e]);
                const float v = __fadd_rn(__fmul_rn(s_inv[kk[1][e + 1]], __fsub_rn(ff[e], ku)), a[1][e + 1]);
                const int xx = xB + e;
                o[e] = (rin && xx >= 1 && xx <= N - 2) ? v : 0.0f;
            }
            st_global_v4(uo + (long long)yB * p.pitch + xB, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

}  // namespace mgfea
