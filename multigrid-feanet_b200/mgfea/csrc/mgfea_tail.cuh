// Coarse tail: every multigrid level with N <= 65 processed by ONE CTA per sample in ONE launch, all fields resident
// in shared memory (the levels below ~129^2 are pure launch/latency cost when run as separate grid kernels: 9 levels x
// 2 launches of the 12-level 4097^2 hierarchy).  Arithmetic is the canonical order of oracle/mgfea_oracle.c, so the
// result is bit-identical to running the same levels through the tile kernels.
//
// Shared-memory field layout per level: (N+2) rows x S floats, S = roundup4(N) + 8; node (i,j) lives at
// (i+1)*S + 4 + j.  The ghost row above/below and the pad columns are zero == the reference's zero padding, so the
// stencils need no bounds checks.  Work item = (interior row i, group g of 4 columns 4g..4g+3).
#pragma once
#include "mgfea_tile.cuh"

namespace mgfea {

constexpr int TAIL_MAXLEV = 8;
constexpr int TAIL_THREADS = 1024;
constexpr int TAIL_MAXN = 65;

struct TailLevel {
    int N, S;
    int off_u, off_v, off_f;  // float offsets of the two solution buffers and the rhs
    int off_t0, off_t1, off_t2;  // residual / HNet temporaries in this level's layout (ghost cells stay zero)
    int off_k;                // byte offset of the key array ((N+2) x S bytes) or -1
    int npat, key_pitch;
    const unsigned char *keys;
    const float *ktab, *invd;
};

struct TailParams {
    int nlev, B;
    TailLevel lv[TAIL_MAXLEV];
    int off_tab;        // float offset of per-level tables: nlev x (MAXPAT*9 + MAXPAT)
    int total_floats;   // floats to zero at start
    // global in/out of the top tail level
    const float *f_in;
    float *u_out;
    int pitch;
    long long plane;
    // cycle parameters
    int nu1, nu2, smoother, nlayers, prolong_mode, quirk;
    const float *hw;
    const float *rtab;
    int rtab_n, r_has_scale;
    float r_scale;
    const float *r_scale_dev;
    const float *ptab;
    int ptab_n, p_has_scale;
    float p_scale;
    const float *p_scale_dev;
    void *ctl;
    long long *trace;  // optional: thread 0 appends clock64() after every stage (profiling aid, see mgfea_trace)
};

struct TailTabs {  // shared by all levels
    float rtab[MAXPAT * 9], ptab[MAXPAT * 9], hw[MAXLAYERS * 9];
    float r_scale, p_scale;
};

__device__ __forceinline__ int t_node(const TailLevel &L, int i, int j) { return (i + 1) * L.S + 4 + j; }

// 3 rows x 6 columns around group g of row i
struct Win {
    float a[3][6];
};
__device__ __forceinline__ Win t_window(const float *buf, const TailLevel &L, int i, int g) {
    Win w;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float *row = buf + (i + d) * L.S + 4 + 4 * g;  // row i-1+d
        const float4 v = *reinterpret_cast<const float4 *>(row);
        w.a[d][0] = row[-1];
        w.a[d][1] = v.x;
        w.a[d][2] = v.y;
        w.a[d][3] = v.z;
        w.a[d][4] = v.w;
        w.a[d][5] = row[4];
    }
    return w;
}
struct KWin {
    int k[3][6];
};
__device__ __forceinline__ KWin t_kwindow(const unsigned char *kb, const TailLevel &L, int i, int g) {
    KWin w;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const unsigned char *row = kb + (i + d) * L.S + 4 + 4 * g;
#pragma unroll
        for (int q = 0; q < 6; ++q) w.k[d][q] = row[q - 1];
    }
    return w;
}

// K u at the 4 nodes of the group; tab = this level's [npat][9]
template <bool KEYS>
__device__ __forceinline__ void t_stencil(const Win &w, const KWin &kw, const float *tab, float (&acc)[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float wt = KEYS ? tab[9 * kw.k[d][e + q] + 3 * d + q] : tab[3 * d + q];
                s = (d == 0 && q == 0) ? __fmul_rn(wt, w.a[d][e + q]) : __fmaf_rn(wt, w.a[d][e + q], s);
            }
        acc[e] = s;
    }
}

// one Jacobi sweep: dst = mask(inv*(f - K src) + src) on interior nodes (ring stays 0); optional x = dst - src
template <bool KEYS>
__device__ __forceinline__ void t_jacobi(float *sm, const unsigned char *smb, const TailLevel &L, const float *tab,
                                         const float *invd, const float *src, float *dst, float *xdst) {
    const int G = (L.N - 1 + 3) / 4, ntask = (L.N - 2) * G;
    const float *f = sm + L.off_f;
    const unsigned char *kb = KEYS ? smb + L.off_k : nullptr;
    for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
        const int i = 1 + t / G, g = t - (i - 1) * G;
        const Win w = t_window(src, L, i, g);
        KWin kw;
        if (KEYS) kw = t_kwindow(kb, L, i, g);
        float acc[4];
        t_stencil<KEYS>(w, kw, tab, acc);
        const int o = t_node(L, i, 4 * g);
        const float4 fv = *reinterpret_cast<const float4 *>(f + o);
        const float ff[4] = {fv.x, fv.y, fv.z, fv.w};
        float out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float inv = KEYS ? invd[kw.k[1][e + 1]] : invd[0];
            const float v = __fadd_rn(__fmul_rn(inv, __fsub_rn(ff[e], acc[e])), w.a[1][e + 1]);
            const int j = 4 * g + e;
            out[e] = (j >= 1 && j <= L.N - 2) ? v : 0.0f;
        }
        *reinterpret_cast<float4 *>(dst + o) = make_float4(out[0], out[1], out[2], out[3]);
        if (xdst)
            *reinterpret_cast<float4 *>(xdst + o) =
                make_float4(__fsub_rn(out[0], w.a[1][1]), __fsub_rn(out[1], w.a[1][2]), __fsub_rn(out[2], w.a[1][3]),
                            __fsub_rn(out[3], w.a[1][4]));
    }
}

// one HNet layer: dst = mask(w9 (*) src) [+ base]
__device__ __forceinline__ void t_hlayer(const TailLevel &L, const float *w9, const float *src, float *dst,
                                         const float *base) {
    const int G = (L.N - 1 + 3) / 4, ntask = (L.N - 2) * G;
    float wt[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wt[q] = w9[q];
    for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
        const int i = 1 + t / G, g = t - (i - 1) * G;
        const Win w = t_window(src, L, i, g);
        KWin kw;
        float acc[4];
        t_stencil<false>(w, kw, wt, acc);
        const int o = t_node(L, i, 4 * g);
        float out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * g + e;
            out[e] = (j >= 1 && j <= L.N - 2) ? acc[e] : 0.0f;
        }
        if (base) {
            const float4 jv = *reinterpret_cast<const float4 *>(base + o);
            out[0] = __fadd_rn(jv.x, out[0]);
            out[1] = __fadd_rn(jv.y, out[1]);
            out[2] = __fadd_rn(jv.z, out[2]);
            out[3] = __fadd_rn(jv.w, out[3]);
        }
        *reinterpret_cast<float4 *>(dst + o) = make_float4(out[0], out[1], out[2], out[3]);
    }
}

// r = f - K u on interior nodes -> rdst (ring of r is never used by the restriction's interior coarse nodes... it IS
// used: coarse node 1 reads fine row 1..3 only; fine ring rows 0 / N-1 are not read).  Written for rows 1..N-2, all
// columns 1..N-2; other entries of rdst must be zero-irrelevant (only nodes 1..N-2 are read).
template <bool KEYS>
__device__ __forceinline__ void t_residual(float *sm, const unsigned char *smb, const TailLevel &L, const float *tab,
                                           const float *u, float *rdst) {
    const int G = (L.N - 1 + 3) / 4, ntask = (L.N - 2) * G;
    const float *f = sm + L.off_f;
    const unsigned char *kb = KEYS ? smb + L.off_k : nullptr;
    for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
        const int i = 1 + t / G, g = t - (i - 1) * G;
        const Win w = t_window(u, L, i, g);
        KWin kw;
        if (KEYS) kw = t_kwindow(kb, L, i, g);
        float acc[4];
        t_stencil<KEYS>(w, kw, tab, acc);
        const int o = t_node(L, i, 4 * g);
        const float4 fv = *reinterpret_cast<const float4 *>(f + o);
        *reinterpret_cast<float4 *>(rdst + o) = make_float4(__fsub_rn(fv.x, acc[0]), __fsub_rn(fv.y, acc[1]),
                                                            __fsub_rn(fv.z, acc[2]), __fsub_rn(fv.w, acc[3]));
    }
}

// f_c[I][J] = scale * chain R[key(src)][3a+c] * r[2I-1+a][2J-1+c] on interior coarse nodes
template <bool KEYS>
__device__ __forceinline__ void t_restrict(float *sm, const unsigned char *smb, const TailLevel &L, const TailLevel &C,
                                           const TailTabs &T, const TailParams &p, const float *r) {
    const int n = C.N - 2, ntask = n * n;
    const unsigned char *kb = (KEYS && p.rtab_n > 1) ? smb + L.off_k : nullptr;
    float *fc = sm + C.off_f;
    for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
        const int I = 1 + t / n, J = 1 + (t - (I - 1) * n);
        float s = 0.0f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int o = t_node(L, 2 * I - 1 + a, 2 * J - 1 + c);
                const int k = kb ? kb[o] : 0;
                const float wt = T.rtab[9 * k + 3 * a + c];
                s = (a == 0 && c == 0) ? __fmul_rn(wt, r[o]) : __fmaf_rn(wt, r[o], s);
            }
        fc[t_node(C, I, J)] = p.r_has_scale ? __fmul_rn(T.r_scale, s) : s;
    }
}

// u += scale * P(vc) on all nodes 0..N-1 (bilinear: masked by the default ring; table: unmasked)
template <bool KEYS>
__device__ __forceinline__ void t_prolong(float *sm, const unsigned char *smb, const TailLevel &L, const TailLevel &C,
                                          const TailTabs &T, const TailParams &p, const float *vc, float *u) {
    const int N = L.N, ntask = N * N;
    const bool table = (p.prolong_mode == 3);
    const unsigned char *kc = (KEYS && table && p.ptab_n > 1 && C.off_k >= 0) ? smb + C.off_k : nullptr;
    const bool seq = (N <= 33);
    for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
        const int y = t / N, x = t - y * N;
        const int I = y >> 1, J = x >> 1;
        float e;
        if (!table) {
            const float a = vc[t_node(C, I, J)];
            if ((y & 1) && (x & 1)) {
                const float b = vc[t_node(C, I, J + 1)], c = vc[t_node(C, I + 1, J)], d = vc[t_node(C, I + 1, J + 1)];
                if (seq) {
                    e = __fadd_rn(__fmul_rn(0.25f, a), __fmul_rn(0.25f, b));
                    e = __fadd_rn(e, __fmul_rn(0.25f, c));
                    e = __fadd_rn(e, __fmul_rn(0.25f, d));
                } else {
                    const float tp = __fadd_rn(__fmul_rn(0.5f, a), __fmul_rn(0.5f, b));
                    const float bt = __fadd_rn(__fmul_rn(0.5f, c), __fmul_rn(0.5f, d));
                    e = __fadd_rn(__fmul_rn(0.5f, tp), __fmul_rn(0.5f, bt));
                }
            } else if (x & 1) {
                e = __fadd_rn(__fmul_rn(0.5f, a), __fmul_rn(0.5f, vc[t_node(C, I, J + 1)]));
            } else if (y & 1) {
                e = __fadd_rn(__fmul_rn(0.5f, a), __fmul_rn(0.5f, vc[t_node(C, I + 1, J)]));
            } else {
                e = a;
            }
            if (y == 0 || x == 0 || y == N - 1 || x == N - 1) e = 0.0f;  // fine level's reset_boundary (default ring)
        } else {
            float s = 0.0f;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if ((y + 1 - a) & 1) continue;
                const int II = (y + 1 - a) >> 1;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if ((x + 1 - c) & 1) continue;
                    const int JJ = (x + 1 - c) >> 1;
                    const int o = t_node(C, II, JJ);  // II,JJ in [0, Nc]: ghost cells are zero
                    const int k = kc ? kc[o] : 0;
                    s = __fmaf_rn(T.ptab[9 * k + 3 * a + c], vc[o], s);
                }
            }
            e = p.p_has_scale ? __fmul_rn(T.p_scale, s) : s;
        }
        const int o = t_node(L, y, x);
        u[o] = __fadd_rn(u[o], e);
    }
}

// bilinear variant of t_prolong with one work item per (row, group of 4 columns): three coarse values per coarse row
// serve four fine nodes (same expressions as the streaming / mid kernels)
__device__ __forceinline__ void t_prolong_bilinear4(const TailLevel &L, const TailLevel &C, const float *vc, float *u) {
    const int N = L.N, G = (N + 3) / 4, ntask = N * G;
    const bool seq = (N <= 33);
    for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
        const int y = t / G, g = t - y * G, x = 4 * g;
        if (y == 0 || y == N - 1) continue;  // ring rows: correction masked to zero
        const float *c0 = vc + t_node(C, y >> 1, x >> 1);
        const float t0 = c0[0], t1 = c0[1], t2 = c0[2];  // columns beyond the coarse grid are zero pads
        float e[4];
        if ((y & 1) == 0) {
            e[0] = t0;
            e[1] = __fadd_rn(__fmul_rn(0.5f, t0), __fmul_rn(0.5f, t1));
            e[2] = t1;
            e[3] = __fadd_rn(__fmul_rn(0.5f, t1), __fmul_rn(0.5f, t2));
        } else {
            const float *c1 = c0 + C.S;
            const float b0 = c1[0], b1 = c1[1], b2 = c1[2];
            e[0] = __fadd_rn(__fmul_rn(0.5f, t0), __fmul_rn(0.5f, b0));
            e[2] = __fadd_rn(__fmul_rn(0.5f, t1), __fmul_rn(0.5f, b1));
            if (seq) {
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
        float4 *cell = reinterpret_cast<float4 *>(u + t_node(L, y, x));
        float4 uv = *cell;
        float *up = reinterpret_cast<float *>(&uv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int xx = x + q;
            if (xx >= 1 && xx <= N - 2) up[q] = __fadd_rn(up[q], e[q]);
        }
        *cell = uv;
    }
}

// Down leg of one level from a ZERO guess, single pattern, one Jacobi pre-sweep: u1 = mask(inv * f) is pointwise, so
// u1 (-> dst) and r = f - K u1 (-> rdst) come out of ONE pass over the f window (one barrier less per level).
__device__ __forceinline__ void t_down_fused(float *sm, const TailLevel &L, const float *tab, const float *invd,
                                             float *dst, float *rdst) {
    const int G = (L.N - 1 + 3) / 4, ntask = (L.N - 2) * G;
    const float *f = sm + L.off_f;
    const float inv = invd[0];
    float kw[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) kw[q] = tab[q];
    for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
        const int i = 1 + t / G, g = t - (i - 1) * G;
        const Win w = t_window(f, L, i, g);
        float u1[3][6];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int yy = i - 1 + d;
            const bool rin = (yy >= 1 && yy <= L.N - 2);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int xx = 4 * g - 1 + q;
                u1[d][q] = (rin && xx >= 1 && xx <= L.N - 2) ? __fadd_rn(__fmul_rn(inv, w.a[d][q]), 0.0f) : 0.0f;
            }
        }
        float r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float s = __fmul_rn(kw[0], u1[0][e]);
            s = __fmaf_rn(kw[1], u1[0][e + 1], s);
            s = __fmaf_rn(kw[2], u1[0][e + 2], s);
            s = __fmaf_rn(kw[3], u1[1][e], s);
            s = __fmaf_rn(kw[4], u1[1][e + 1], s);
            s = __fmaf_rn(kw[5], u1[1][e + 2], s);
            s = __fmaf_rn(kw[6], u1[2][e], s);
            s = __fmaf_rn(kw[7], u1[2][e + 1], s);
            s = __fmaf_rn(kw[8], u1[2][e + 2], s);
            r[e] = __fsub_rn(w.a[1][e + 1], s);
        }
        const int o = t_node(L, i, 4 * g);
        *reinterpret_cast<float4 *>(dst + o) = make_float4(u1[1][1], u1[1][2], u1[1][3], u1[1][4]);
        *reinterpret_cast<float4 *>(rdst + o) = make_float4(r[0], r[1], r[2], r[3]);
    }
}

template <bool KEYS>
__global__ void __launch_bounds__(TAIL_THREADS, 1) mg_tail_kernel(const TailParams p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float *sm = reinterpret_cast<float *>(smraw);
    __shared__ TailTabs T;
    pdl_launch_dependents();
    const int tid = threadIdx.x, b = blockIdx.x;
    int ntrace = 0;
    auto stamp = [&]() {
        if (p.trace != nullptr && tid == 0 && b == 0 && ntrace < 64) p.trace[ntrace++] = clock64();
    };
    stamp();
    // ---- zero all fields (ghost cells / pads / zero initial guesses), load tables and keys
    for (int i = tid; i < p.total_floats / 4; i += blockDim.x)
        reinterpret_cast<float4 *>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < MAXPAT * 9; i += blockDim.x) {
        T.rtab[i] = (p.rtab && i < p.rtab_n * 9) ? p.rtab[i] : 0.0f;
        T.ptab[i] = (p.ptab && i < p.ptab_n * 9) ? p.ptab[i] : 0.0f;
    }
    if (tid < MAXLAYERS * 9) T.hw[tid] = (p.hw && tid < p.nlayers * 9) ? p.hw[tid] : 0.0f;
    if (tid == 0) {
        T.r_scale = p.r_scale_dev ? *p.r_scale_dev : p.r_scale;
        T.p_scale = p.p_scale_dev ? *p.p_scale_dev : p.p_scale;
    }
    __syncthreads();
    stamp();
    float *tabs = sm + p.off_tab;
    // all levels' stencil tables in one flat pass (independent loads: one memory round trip, not one per level)
    constexpr int TABW = MAXPAT * 9 + MAXPAT;
    for (int idx = tid; idx < p.nlev * TABW; idx += blockDim.x) {
        const int l = idx / TABW, i = idx - l * TABW;
        const TailLevel &L = p.lv[l];
        float v = 0.0f;
        if (i < L.npat * 9) v = L.ktab[i];
        else if (i >= MAXPAT * 9 && i - MAXPAT * 9 < L.npat) v = L.invd[i - MAXPAT * 9];
        tabs[idx] = v;
    }
    if (KEYS) {
        for (int l = 0; l < p.nlev; ++l) {
            const TailLevel &L = p.lv[l];
            if (L.off_k < 0) continue;
            unsigned char *kb = smraw + L.off_k;  // whole layout: ghost rows / pad columns get key 0
            for (int i = tid; i < (L.N + 2) * L.S; i += blockDim.x) {
                const int y = i / L.S - 1, x = i - (y + 1) * L.S - 4;
                kb[i] = (y >= 0 && y < L.N && x >= 0 && x < L.N) ? L.keys[(long long)y * L.key_pitch + x]
                                                                 : (unsigned char)0;
            }
        }
    }
    stamp();
    pdl_wait();  // tables / keys above are never written by a kernel; the restricted rhs below is
    stamp();
    const int solve_done = (p.ctl != nullptr) ? ld_volatile_s32(&reinterpret_cast<const Ctl *>(p.ctl)->done) : 0;
    {
        const TailLevel &L = p.lv[0];
        const float *fin = p.f_in + (long long)b * p.plane;
        float *f = sm + L.off_f;
        // rows are 16-byte aligned in HBM (pitch % 4 == 0) and in the layout; columns [N, roundup4(N)) are zero in both
        const int G = (L.N + 3) / 4;
        for (int i = tid; i < L.N * G; i += blockDim.x) {
            const int y = i / G, g = i - y * G;
            *reinterpret_cast<float4 *>(f + t_node(L, y, 4 * g)) =
                __ldcg(reinterpret_cast<const float4 *>(fin + (long long)y * p.pitch + 4 * g));
        }
    }
    if (solve_done) return;
    __syncthreads();
    stamp();

    // current solution buffer per level (ping-pong between off_u / off_v)
    int cur[TAIL_MAXLEV];
    for (int l = 0; l < p.nlev; ++l) cur[l] = 0;

    auto relax = [&](int l, int nsweeps) {
        const TailLevel &L = p.lv[l];
        const float *tb = tabs + l * (MAXPAT * 9 + MAXPAT);
        float *tmp0 = sm + L.off_t0, *tmp1 = sm + L.off_t1, *tmp2 = sm + L.off_t2;
        for (int s = 0; s < nsweeps; ++s) {
            float *src = sm + (cur[l] ? L.off_v : L.off_u);
            float *dst = sm + (cur[l] ? L.off_u : L.off_v);
            if (p.smoother == 0) {
                t_jacobi<KEYS>(sm, smraw, L, tb, tb + MAXPAT * 9, src, dst, nullptr);
                __syncthreads();
                stamp();
            } else {
                // J -> tmp0, x -> tmp1, layers ping-pong tmp1/tmp2, last layer adds J into dst
                t_jacobi<KEYS>(sm, smraw, L, tb, tb + MAXPAT * 9, src, tmp0, tmp1);
                __syncthreads();
                stamp();
                float *a = tmp1, *bb = tmp2;
                for (int q = 0; q < p.nlayers; ++q) {
                    const bool last = (q == p.nlayers - 1);
                    t_hlayer(L, T.hw + 9 * q, a, last ? dst : bb, last ? tmp0 : nullptr);
                    __syncthreads();
                    stamp();
                    float *sw = a;
                    a = bb;
                    bb = sw;
                }
            }
            cur[l] ^= 1;
        }
    };

    // ---- down leg
    for (int l = 0; l < p.nlev; ++l) {
        if (!KEYS && !p.quirk && p.nu1 == 1 && p.smoother == 0 && l < p.nlev - 1) {
            // every tail level starts from a zero guess: pre-sweep and residual in one pass
            const TailLevel &L = p.lv[l];
            const float *tb = tabs + l * (MAXPAT * 9 + MAXPAT);
            t_down_fused(sm, L, tb, tb + MAXPAT * 9, sm + (cur[l] ? L.off_u : L.off_v), sm + L.off_t0);
            cur[l] ^= 1;
            __syncthreads();
            stamp();
            t_restrict<KEYS>(sm, smraw, L, p.lv[l + 1], T, p, sm + L.off_t0);
            __syncthreads();
            stamp();
            continue;
        }
        if (!p.quirk) relax(l, p.nu1);
        if (l < p.nlev - 1) {
            const TailLevel &L = p.lv[l];
            const float *tb = tabs + l * (MAXPAT * 9 + MAXPAT);
            const float *u = sm + (cur[l] ? L.off_v : L.off_u);
            t_residual<KEYS>(sm, smraw, L, tb, u, sm + L.off_t0);
            __syncthreads();
            stamp();
            t_restrict<KEYS>(sm, smraw, L, p.lv[l + 1], T, p, sm + L.off_t0);
            __syncthreads();
            stamp();
        }
    }
    // ---- up leg
    for (int l = p.nlev - 1; l >= 0; --l) {
        if (l < p.nlev - 1) {
            const TailLevel &L = p.lv[l], &C = p.lv[l + 1];
            float *u = sm + (cur[l] ? L.off_v : L.off_u);
            const float *vc = sm + (cur[l + 1] ? C.off_v : C.off_u);
            if (p.prolong_mode != 3)
                t_prolong_bilinear4(L, C, vc, u);
            else
                t_prolong<KEYS>(sm, smraw, L, C, T, p, vc, u);
            __syncthreads();
            stamp();
        }
        relax(l, p.nu2);
    }
    // ---- store the top tail level's solution (rows 0..N-1; columns [N, roundup4(N)) are written as zero)
    {
        const TailLevel &L = p.lv[0];
        const float *u = sm + (cur[0] ? L.off_v : L.off_u);
        float *uo = p.u_out + (long long)b * p.plane;
        const int G = (L.N + 3) / 4;
        for (int i = tid; i < L.N * G; i += blockDim.x) {
            const int y = i / G, g = i - y * G;
            const float4 v = *reinterpret_cast<const float4 *>(u + t_node(L, y, 4 * g));
            *reinterpret_cast<float4 *>(uo + (long long)y * p.pitch + 4 * g) = v;
        }
    }
    stamp();
}

}  // namespace mgfea
