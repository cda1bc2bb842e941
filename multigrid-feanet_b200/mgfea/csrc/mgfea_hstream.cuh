// Streaming (register-chained) legs of the V-cycle for the LEARNED smoother (HJacIterator.HRelax + HNet.forward,
// M-FEANet-mg_test.ipynb cells 4, 5) on two-phase or single-pattern meshes with table restriction / prolongation
// (FEANet/multigrid.py:62-73, 124-130, 177-179) -- the config the tile programs run at 0.28 of the HBM roofline because
// every HNet layer is one more barrier-separated pass over a shared-memory box.
//
// Same skeleton as mg_stream2_kernel (mgfea_stream.cuh): one WARP owns a strip of 128 box columns (lane l: columns
// 4l..4l+3), marches down its rows, and every stage of the leg consumes the row the previous stage produced in the same
// step from a rotating 3-row register window.  The chain is longer:
//
//   u row a -> J row a-1 (Jacobi), x = J - u -> L1 row a-3 -> L2 row a-5 -> L3 row a-7, u' = J + L3 (stored)
//           -> r row a-9 (= f - K u') -> down leg: coarse f row (a-10)/2 (restriction) | up leg: sum r^2
//
// Every stage works on the rows its producer finished in EARLIER steps (a software pipeline: two rows of lag per stage), so
// the five stencils of a step are independent instruction streams.  (Measured: not faster than one dependent chain per
// step, 85.7 vs 85.8 us on the 4097^2 down leg -- ptxas interleaves consecutive steps of the unrolled block anyway -- but
// it costs no registers either: profiles/r02_hstream_legs.log.)  The data dependences are those of the operators: column
// halo 8 (112 interior columns per strip), 6 rows above / 4 below a strip; with the pipeline lags a strip of R rows
// streams R + 15.
// u, f (and the key map) are read from HBM once, u' is written once; the three HNet layers never leave the registers.
// Weights of the current material pattern sit in registers; blocks of 6 rows that a material interface crosses take a
// per-node lookup variant (as in mg_stream2_kernel<.., KEYS>).
//
// Arithmetic: operation order of oracle/mgfea_oracle.c (orc_hjacobi, orc_residual, orc_restrict, orc_prolong_table):
// row-major FMA chains, explicit roundings; FFMA2 computes two such chains side by side (half the issue slots: the
// scalar formulation of the same chain measured 118 vs 86 us on the 4097^2 down leg, see DESIGN).
#pragma once
#include "mgfea_stream.cuh"

namespace mgfea {

constexpr int HS_WARPS = 8;
constexpr int HS_TWI = BW - 16;  // interior columns per strip (column halo 8 = chain depth 6, rounded to the lane width)
constexpr int HS_RD = 12;        // u (and coarse) prefetch ring rows: two halves of 6, the unroll factor
constexpr int HS_FD = 18;        // f ring rows (three thirds of 6): the residual stage re-reads f nine rows behind
constexpr int HS_JD = 6;         // J delay line (rows): J row a-1 meets the last layer's row six steps later
constexpr int HS_PD = 6;         // prefetch distance (rows)
constexpr int HS_KD = 32;        // key ring rows (look-back 10 + the block being fetched = 22, rounded to a power of two)
constexpr int HS_NL = 3;         // HNet layers
// float4 units per lane: u ring, f ring, J delay line, coarse float2 ring (up leg)
__host__ __device__ constexpr int hs_ring_f4(int mode) { return HS_RD + HS_FD + HS_JD + (mode == 1 ? HS_RD / 2 : 0); }
__host__ __device__ constexpr size_t hs_smem_bytes(int mode, bool keys) {
    return (size_t)HS_WARPS * ((size_t)hs_ring_f4(mode) * 32 * 16 + (keys ? (size_t)HS_KD * 32 * 4 : 0));
}

// keyed restriction: feed residual row values (columns 4l-1 .. 4l+3) into the lane's two coarse chains; taps 3*row..3*row+2
// with the weight of every SOURCE node's pattern.  first: the chain starts with a product (row 0).
__device__ __noinline__ void hs_restrict_feed_keys(const float *rtab, unsigned int w, int row, float rl, float4 r,
                                                   float &acc0, float &acc1) {
    const unsigned int l = __shfl_up_sync(0xffffffffu, w, 1);
    const int k0 = (int)(l >> 24), k1 = (int)(w & 0xffu), k2 = (int)((w >> 8) & 0xffu), k3 = (int)((w >> 16) & 0xffu),
              k4 = (int)(w >> 24);
    const int t = 3 * row;
    if (row == 0) {
        acc0 = __fmul_rn(rtab[9 * k0 + t], rl);
        acc1 = __fmul_rn(rtab[9 * k2 + t], r.y);
    } else {
        acc0 = __fmaf_rn(rtab[9 * k0 + t], rl, acc0);
        acc1 = __fmaf_rn(rtab[9 * k2 + t], r.y, acc1);
    }
    acc0 = __fmaf_rn(rtab[9 * k1 + t + 1], r.x, acc0);
    acc1 = __fmaf_rn(rtab[9 * k3 + t + 1], r.z, acc1);
    acc0 = __fmaf_rn(rtab[9 * k2 + t + 2], r.y, acc0);
    acc1 = __fmaf_rn(rtab[9 * k4 + t + 2], r.w, acc1);
}

// table prolongation of one fine row (ConvTranspose2d(C->1, 3, stride 2, pad 1), taps in (a asc, c asc) order):
// top = coarse row floor(y/2), bot = coarse row (y+1)/2 (odd rows), each at coarse columns cxl, cxl+1, cxl+2;
// kt / kb = pattern keys of those coarse nodes (all zero for a single table)
__device__ __forceinline__ float4 hs_prolong_table(const float *P, bool odd, const float (&top)[3], const float (&bot)[3],
                                                   const int (&kt)[3], const int (&kb)[3]) {
    float e[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int jl = q >> 1;
        float s;
        if (!odd) {
            if (!(q & 1)) {
                s = __fmul_rn(P[9 * kt[jl] + 4], top[jl]);
            } else {
                s = __fmul_rn(P[9 * kt[jl + 1] + 3], top[jl + 1]);
                s = __fmaf_rn(P[9 * kt[jl] + 5], top[jl], s);
            }
        } else {
            if (!(q & 1)) {
                s = __fmul_rn(P[9 * kb[jl] + 1], bot[jl]);
                s = __fmaf_rn(P[9 * kt[jl] + 7], top[jl], s);
            } else {
                s = __fmul_rn(P[9 * kb[jl + 1] + 0], bot[jl + 1]);
                s = __fmaf_rn(P[9 * kb[jl] + 2], bot[jl], s);
                s = __fmaf_rn(P[9 * kt[jl + 1] + 6], top[jl + 1], s);
                s = __fmaf_rn(P[9 * kt[jl] + 8], top[jl], s);
            }
        }
        e[q] = s;
    }
    return make_float4(e[0], e[1], e[2], e[3]);
}
__device__ __noinline__ float4 hs_prolong_table_keys(const float *P, int odd, float t0, float t1, float t2, float b0,
                                                     float b1, float b2, unsigned int wt, unsigned int wb) {
    // wt / wb: this lane's two coarse keys (bytes 0, 1); the third comes from the next lane
    const unsigned int nt = __shfl_down_sync(0xffffffffu, wt, 1), nb = __shfl_down_sync(0xffffffffu, wb, 1);
    const float top[3] = {t0, t1, t2}, bot[3] = {b0, b1, b2};
    const int kt[3] = {(int)(wt & 0xffu), (int)((wt >> 8) & 0xffu), (int)(nt & 0xffu)};
    const int kb[3] = {(int)(wb & 0xffu), (int)((wb >> 8) & 0xffu), (int)(nb & 0xffu)};
    return hs_prolong_table(P, odd != 0, top, bot, kt, kb);
}

// Rows as STRIDE-2 column pairs: p[j] = (a[j], a[j+2]) for the 6 values a[0..5] = box columns 4l-1 .. 4l+4.  A stencil's
// outputs pair up as A = (e0, e2), B = (e1, e3), which ARE the next stage's p[1] and p[2]: a new row costs two shuffles
// and two register moves (p[0], p[3]).  With adjacent-column pairs (RP, mgfea_stream.cuh) two of the five pairs straddle
// the output pairs and ptxas spent ~17 MOVs per row on them (85 of 233 instructions per step here).
struct RS {
    u64 p[4];
};
__device__ __forceinline__ void stencil_rowsS(const u64 (&w)[9], const RS &t, const RS &m, const RS &b, u64 &A, u64 &B) {
    A = mul2(w[0], t.p[0]);
    B = mul2(w[0], t.p[1]);
    A = fma2(w[1], t.p[1], A);
    B = fma2(w[1], t.p[2], B);
    A = fma2(w[2], t.p[2], A);
    B = fma2(w[2], t.p[3], B);
    A = fma2(w[3], m.p[0], A);
    B = fma2(w[3], m.p[1], B);
    A = fma2(w[4], m.p[1], A);
    B = fma2(w[4], m.p[2], B);
    A = fma2(w[5], m.p[2], A);
    B = fma2(w[5], m.p[3], B);
    A = fma2(w[6], b.p[0], A);
    B = fma2(w[6], b.p[1], B);
    A = fma2(w[7], b.p[1], A);
    B = fma2(w[7], b.p[2], B);
    A = fma2(w[8], b.p[2], A);
    B = fma2(w[8], b.p[3], B);
}
// A = (a1, a3), B = (a2, a4): the lane's four columns
__device__ __forceinline__ RS widenS(u64 A, u64 B) {
    float a1, a2, a3, a4;
    unpack2(A, a1, a3);
    unpack2(B, a2, a4);
    const float a0 = __shfl_up_sync(0xffffffffu, a4, 1);
    const float a5 = __shfl_down_sync(0xffffffffu, a1, 1);
    RS o;
    o.p[0] = pack2(a0, a2);
    o.p[1] = A;
    o.p[2] = B;
    o.p[3] = pack2(a3, a5);
    return o;
}
__device__ __forceinline__ void unpack_rowS(const RS &r, float (&a)[6]) {
    unpack2(r.p[0], a[0], a[2]);
    unpack2(r.p[1], a[1], a[3]);
    float t;
    unpack2(r.p[2], t, a[4]);
    unpack2(r.p[3], t, a[5]);
}
// mask bits of the lane's columns 0..3 applied to A = (e0, e2), B = (e1, e3)
__device__ __forceinline__ void hs_maskS(u64 &A, u64 &B, unsigned int m) {
    float e0, e1, e2, e3;
    unpack2(A, e0, e2);
    unpack2(B, e1, e3);
    e0 = (m & 1u) ? e0 : 0.0f;
    e1 = (m & 2u) ? e1 : 0.0f;
    e2 = (m & 4u) ? e2 : 0.0f;
    e3 = (m & 8u) ? e3 : 0.0f;
    A = pack2(e0, e2);
    B = pack2(e1, e3);
}

// Source-key indexed stencil of one row, inline: the rows a material interface crosses.  tab4 = [pattern][stencil row]
// float4 (taps dj = 0, 1, 2 of that row): ONE 16-byte shared-memory load per source node and stencil row instead of one
// 4-byte load per tap.  Same row-major FMA chain per output as stencil_rowsS.  out = {K A, K B, inv A, inv B}.
__device__ __forceinline__ void hs_stencil_keys(const float *tab4, const float *invt, unsigned int w0, unsigned int w1,
                                                unsigned int w2, const RS &t, const RS &m, const RS &b, u64 *out) {
    const float4 *T4 = reinterpret_cast<const float4 *>(tab4);
    float acc[4];
    float cen[4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const K6 kk = key6(r == 0 ? w0 : (r == 1 ? w1 : w2));
        float a[6];
        unpack_rowS(r == 0 ? t : (r == 1 ? m : b), a);
        float4 w[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) w[j] = T4[3 * kk.k[j] + r];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[e] = (r == 0) ? __fmul_rn(w[e].x, a[e]) : __fmaf_rn(w[e].x, a[e], acc[e]);
            acc[e] = __fmaf_rn(w[e + 1].y, a[e + 1], acc[e]);
            acc[e] = __fmaf_rn(w[e + 2].z, a[e + 2], acc[e]);
        }
        if (r == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) cen[e] = invt[kk.k[e + 1]];
        }
    }
    out[0] = pack2(acc[0], acc[2]);
    out[1] = pack2(acc[1], acc[3]);
    out[2] = pack2(cen[0], cen[2]);
    out[3] = pack2(cen[1], cen[3]);
}

// MODE 0: down leg (HNet sweep, store u, residual, table restriction -> fc); p.u_in == NULL: zero initial guess
// MODE 1: up leg   (bilinear / table prolongation + correction, HNet sweep, store u, optional interior residual norm)
// PTAB (up leg): table prolongation (ConvTranspose taps) instead of bilinear -- compile-time: both variants inline in every
// unrolled block overflowed the instruction caches (ncu: no_instruction was the top stall of the up leg)
// NOL: no HNet layers -- the plain weighted-Jacobi sweep (u' = J) through the same pipeline: the two-phase Jacobi cycle
// with key-indexed table restriction / prolongation, which mg_stream2_kernel<.., KEYS> (single table, bilinear) cannot do
template <int MODE, bool KEYS, bool PTAB = false, bool NOL = false>
__global__ void __launch_bounds__(HS_WARPS * 32, 1) mg_hstream_kernel(const StreamParams p) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    __shared__ double red[HS_WARPS];
    __shared__ int lastflag;
    __shared__ float s_tab[KEYS ? MAXPAT * 9 : 1];                 // stiffness tables of all patterns
    __shared__ __align__(16) float s_tab4[KEYS ? MAXPAT * 12 : 4];  // the same, one float4 per stencil row (per-node path)
    __shared__ float s_inv[KEYS ? MAXPAT : 1];                     // omega / d per pattern
    __shared__ float s_rp[MAXPAT * 9];                             // restriction (MODE 0) / prolongation (MODE 1) tables
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N;
    const int ntab = (MODE == 0) ? p.rtab_n : p.ptab_n;  // 1, or one table per pattern
    const float *gtab = (MODE == 0) ? p.rtab : p.ptab;
    constexpr bool ptable = (MODE == 1) && PTAB;
    const bool pkeys = ptable && ntab > 1 && p.keys_c != nullptr;  // coarse-node keys select the prolongation table
    const bool rkeys = KEYS && (MODE == 0) && ntab > 1;

    // ---- weights in registers for the whole kernel (pattern 0; two-phase strips reload on a pattern change)
    u64 kw2[9], h2[HS_NL][9];
    float tw[9];  // restriction (down leg) / prolongation (up leg, table mode) taps of the current pattern
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const float w = p.ktab[q];
        kw2[q] = pack2(w, w);
        tw[q] = (gtab != nullptr) ? gtab[q] : 0.0f;
#pragma unroll
        for (int l = 0; l < HS_NL; ++l) {
            const float h = NOL ? 0.0f : p.hw[9 * l + q];
            h2[l][q] = pack2(h, h);
        }
    }
    const float inv0 = p.invd[0];
    u64 inv2 = pack2(inv0, inv0);
    const u64 one2 = pack2(p.one, p.one);  // opaque 1.0f: see the Jacobi update
    const float tscale = (MODE == 0) ? (p.r_has_scale ? (p.r_scale_dev ? *p.r_scale_dev : p.r_scale) : 1.0f)
                                     : (p.p_has_scale ? (p.p_scale_dev ? *p.p_scale_dev : p.p_scale) : 1.0f);
    const bool has_scale = (MODE == 0) ? (p.r_has_scale != 0) : (p.p_has_scale != 0);
    if (KEYS) {
        for (int i = threadIdx.x; i < MAXPAT * 9; i += HS_WARPS * 32) s_tab[i] = (i < p.npat * 9) ? p.ktab[i] : 0.0f;
        if (threadIdx.x < MAXPAT) s_inv[threadIdx.x] = (threadIdx.x < p.npat) ? p.invd[threadIdx.x] : 0.0f;
        for (int i = threadIdx.x; i < MAXPAT * 12; i += HS_WARPS * 32) {
            const int pat = i / 12, r = (i % 12) >> 2, c = i & 3;
            s_tab4[i] = (pat < p.npat && c < 3) ? p.ktab[9 * pat + 3 * r + c] : 0.0f;
        }
    }
    for (int i = threadIdx.x; i < MAXPAT * 9; i += HS_WARPS * 32) s_rp[i] = (gtab != nullptr && i < ntab * 9) ? gtab[i] : 0.0f;
    __syncthreads();
    int kcur = 0, ccur = 0;  // fine / coarse pattern whose weights sit in kw, inv (and tw)
    pdl_wait();
    const int solve_done = (p.ctl != nullptr) ? ld_volatile_s32(&reinterpret_cast<const Ctl *>(p.ctl)->done) : 0;
    const bool zero = (p.u_in == nullptr);

    constexpr int PD = HS_PD;
    constexpr int RF4 = hs_ring_f4(MODE);
    float4 *ring_u = reinterpret_cast<float4 *>(st_smem) + warp * (RF4 * 32);
    float4 *ring_f = ring_u + HS_RD * 32;
    float4 *ring_j = ring_f + HS_FD * 32;  // per-lane delay line: every lane reads back what it wrote itself
    float2 *ring_c = reinterpret_cast<float2 *>(ring_j + HS_JD * 32);
    unsigned int *ring_k = reinterpret_cast<unsigned int *>(st_smem + HS_WARPS * RF4 * 32 * 16) + warp * (HS_KD * 32);

    // Strip hand-out.  Static (strip s -> warp s mod resident warps) when the host cuts one strip per resident warp (the
    // default).  Dynamic (one atomic per strip) when it cuts more (hstream_over > 1, two-phase meshes): the strips a
    // material interface runs through cost up to three times a single-pattern strip, and a static assignment left the SMs
    // idle half of the time (ncu: sm__cycles_active 48 % of elapsed on the 4097^2 two-phase down leg).  The queue words
    // live next to the norm ticket in the per-device scratch and are reset by the last CTA to leave -- like the norm
    // ticket they assume that the library's launches on one device are stream-ordered.
    const bool dynq = p.dyn_queue != 0;
    unsigned int *queue = p.counter + 4, *leave = p.counter + 5;
    const int sstep = gridDim.x * HS_WARPS;
    auto next_strip = [&](int prev) {
        if (!dynq) return prev < 0 ? (int)(blockIdx.x * HS_WARPS + warp) : prev + sstep;
        int v = 0;
        if (lane == 0) v = (int)atomicAdd(queue, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    };
    const int total = p.nstrips * p.B;
    for (int s = next_strip(-1); s < total; s = next_strip(s)) {
        int b = 0, rem = s;
        if (p.B > 1) {
            b = __float2int_rz(__int2float_rn(s) * p.inv_nstrips);
            int r0 = s - b * p.nstrips;
            if (r0 < 0) {
                --b;
                r0 += p.nstrips;
            } else if (r0 >= p.nstrips) {
                ++b;
                r0 -= p.nstrips;
            }
            rem = r0;
        }
        int ry = __float2int_rz(__int2float_rn(rem) * p.inv_ntx);
        int tx = rem - ry * p.ntx;
        if (tx < 0) {
            --ry;
            tx += p.ntx;
        } else if (tx >= p.ntx) {
            ++ry;
            tx -= p.ntx;
        }
        const int y0 = ry * p.R;                              // R is even
        const int y1 = (ry == p.nry - 1) ? N : y0 + p.R;      // exclusive
        const int gx = tx * HS_TWI - 8 + 4 * lane;            // first global column of this lane
        const bool lane_int = (lane >= 2 && lane <= 29);
        unsigned int cin = 0, cdom = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (gx + e >= 1 && gx + e <= N - 2) cin |= 1u << e;
            if (gx + e >= 0 && gx + e <= N - 1) cdom |= 1u << e;
        }
        const bool col_ok = (gx >= 0) && (gx + 3 < p.pitch);
        const float *ub = zero ? nullptr : p.u_in + (long long)b * p.plane + gx;
        const float *fb = p.f + (long long)b * p.plane + gx;
        float *uo = p.u_out + (long long)b * p.plane + gx;
        const int cxl = gx >> 1;  // coarse column of this lane's first fine column (gx is a multiple of 4)
        const bool ccol_ok = (MODE == 1) && (cxl >= 0) && (cxl + 1 < p.pitch_c);
        const float *cbp = (MODE == 1) ? p.vc + (long long)b * p.plane_c + cxl : nullptr;
        float *fco = (MODE == 0) ? p.fc + (long long)b * p.plane_c + cxl : nullptr;
        const bool fc_ok = lane_int && cxl >= 0 && cxl <= p.Nc - 1;

        // first streamed row (data dependences only): the down leg's first coarse row y0/2 needs r row y0-1 <- u' y0-2 <-
        // u y0-6 (even start); the up leg's first residual row y0 needs u' y0-1 <- u y0-5 (odd start: the row parity
        // drives the prolongation)
        const int a0 = (MODE == 0) ? y0 - 6 : y0 - 5;
        // last step: the stage that finishes last.  Down leg: the restriction centred on fine row y1-2 closes with r row
        // y1-1 = a-9 (the last strip also writes the coarse ring row (N-1)/2, closed by r row N); up leg: norm row y1-1 =
        // a-9, else the store of u' row y1-1 = a-7
        const int alast = (MODE == 0) ? (y1 == N ? N + 9 : y1 + 8) : (p.want_norm ? y1 + 8 : y1 + 6);
        const int K = alast - a0 + 1;
        const int klo = 0 - a0;               // first k whose row is inside the domain
        const int khi = min(N - a0, K);       // one past the last such k
        const float *pf_u = zero ? nullptr : ub + (long long)a0 * p.pitch;
        const float *pf_f = fb + (long long)a0 * p.pitch;
        int kpf = 0;
        auto fetch_keys = [&](int keyrow0) {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int kr = keyrow0 + j, gy = a0 + kr;
                const bool okk = (gx >= 0 && gx + 3 < p.key_pitch && gy >= 0 && gy < N);
                const void *src = okk ? (const void *)(p.keys + (long long)gy * p.key_pitch + gx) : (const void *)p.f;
                const uint32_t sz = okk ? 4u : 0u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(&ring_k[(kr % HS_KD) * 32 + lane])),
                             "l"(src), "r"(sz)
                             : "memory");
            }
        };
        // prefetch of streamed row kpf into u-ring row `su`, f-ring row `sf`
        auto prefetch = [&](auto check_tag, int su, int sf) {
            constexpr bool CHECK = decltype(check_tag)::value;
            const bool ok = !CHECK || (col_ok && kpf >= klo && kpf < khi);
            if (!zero) st_cp16(&ring_u[su * 32 + lane], ok ? (const void *)pf_u : (const void *)p.f, ok);
            st_cp16(&ring_f[sf * 32 + lane], ok ? (const void *)pf_f : (const void *)p.f, ok);
            if (MODE == 1) {  // coarse row ceil(a/2)
                const int I = (a0 + kpf + 1) >> 1;
                const bool okc = !CHECK || (ccol_ok && (I >= 0) && (I < p.Nc) && kpf < K);
                const void *src = okc ? (const void *)(cbp + (long long)I * p.pitch_c) : (const void *)p.f;
                const uint32_t sz = okc ? 8u : 0u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(&ring_c[su * 32 + lane])), "l"(src),
                             "r"(sz)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (!zero) pf_u += p.pitch;
            pf_f += p.pitch;
            ++kpf;
        };
        // coarse keys of the lane's coarse columns cxl, cxl+1 on coarse row I (two bytes), 0 outside
        auto coarse_keys = [&](int I) -> unsigned int {
            if (!pkeys || I < 0 || I >= p.Nc || cxl < 0 || cxl + 1 >= p.key_pitch_c) return 0u;
            return (unsigned int)__ldg(reinterpret_cast<const unsigned short *>(p.keys_c + (long long)I * p.key_pitch_c + cxl));
        };
        if (KEYS) fetch_keys(0);
#pragma unroll
        for (int k = 0; k < PD; ++k) prefetch(std::true_type{}, k, k);  // at k0 = 0 streamed row k sits in ring row k
        if (solve_done) break;

        // rotating 3-row windows; slot (ph+2)%3 is the one the producer overwrites in this step
        RS A[3], X0[3], X1[3], X2[3], Bw[3];
        u64 rawlo = 0, rawhi = 0;  // un-reset input row a-1 (edge strips)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) A[i].p[j] = X0[i].p[j] = X1[i].p[j] = X2[i].p[j] = Bw[i].p[j] = 0;
        }
        float racc0 = 0.f, racc1 = 0.f;
        float *st_u = uo + (long long)(a0 - 7) * p.pitch;  // u' row a-7 at step 0
        float *st_c = (MODE == 0) ? fco + (long long)((a0 - 10) / 2) * p.pitch_c : nullptr;  // coarse row (a-10)/2
        float vt[3] = {0.f, 0.f, 0.f};  // coarse row floor(a/2) at coarse columns cxl .. cxl+2 (up leg)
        unsigned int kvt = 0;            // its keys (lane's two bytes)
        unsigned int ckw[3] = {0u, 0u, 0u}, ckn[3] = {0u, 0u, 0u};  // coarse keys of this / the next block's 3 coarse rows
        double part = 0.0;
        if (MODE == 1 && a0 >= 1) {  // odd first row: its upper coarse row
            const int I = a0 >> 1;
            const float *row = p.vc + (long long)b * p.plane_c + (long long)I * p.pitch_c;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int c = cxl + q;
                vt[q] = (c >= 0 && c < p.Nc && I >= 0 && I < p.Nc) ? __ldg(row + c) : 0.0f;
            }
            kvt = coarse_keys(I);
        }
        if (MODE == 1 && pkeys) {
            const int I0 = (a0 + 1) >> 1;
#pragma unroll
            for (int j = 0; j < 3; ++j) ckw[j] = coarse_keys(I0 + j);
        }

        // ONE block body for every strip position: pipeline-fill / store-range guards, domain masks and checked prefetches
        // are always on.  Separate steady-state / guarded / edge instantiations (as in mg_stream2_kernel) execute ~15 % fewer
        // instructions per row but were 18-21 % SLOWER here: at 233+ instructions per row step every additional variant in
        // flight evicts the others from the instruction caches (profiles/r02_hstream_legs.log)
        auto block6 = [&](auto keyed_tag, int k0) {
            constexpr bool GUARD = true, EDGE = true;
            constexpr bool KEYED = decltype(keyed_tag)::value;
            const int blk = k0 / 6;
            // u / coarse ring: two halves; streamed row k0 + d, d in [0, 11]
            const int uc = (blk & 1) * 6, un = 6 - uc;
            auto urow = [&](int d) { return d < 6 ? uc + d : un + d - 6; };
            // f ring: three thirds; streamed row k0 + d, d in [-12, 11]
            const int t3 = blk % 3;
            const int fc_ = t3 * 6, fn = ((t3 + 1) % 3) * 6, fp = ((t3 + 2) % 3) * 6;
            auto frow = [&](int d) { return d < -6 ? fn + d + 12 : (d < 0 ? fp + d + 6 : (d < 6 ? fc_ + d : fn + d - 6)); };
            auto krow = [&](int k) { return ((k + HS_KD) % HS_KD) * 32 + lane; };
            if (MODE == 1 && pkeys) {  // the next block's coarse keys: a whole block of latency to hide
                const int I0n = (a0 + k0 + 6 + 1) >> 1;
#pragma unroll
                for (int j = 0; j < 3; ++j) ckn[j] = coarse_keys(I0n + j);
            }
#pragma unroll
            for (int ph = 0; ph < 6; ++ph) {
                const int k = k0 + ph;
                if (GUARD && k >= K) break;
                const int a = a0 + k;
                if (KEYS && ph == 0) fetch_keys(k0 + 6);
                prefetch(std::true_type{}, urow(ph + PD), frow(ph + PD));
                asm volatile("cp.async.wait_group %0;" ::"n"(PD) : "memory");
                // window slots before this step's producers write: oldest (ph+2)%3, middle ph%3, newest (ph+1)%3
                // ================= stage 5: residual row a-9 from the u' rows of earlier steps
                if ((MODE == 0 || p.want_norm) && (!GUARD || k >= 10)) {
                    const int yr = a - 9;
                    const RS &t = Bw[(ph + 2) % 3], &m = Bw[ph % 3], &bq = Bw[(ph + 1) % 3];  // rows a-10, a-9, a-8
                    u64 rlo, rhi;
                    if (KEYED) {
                        u64 out[4];
                        hs_stencil_keys(s_tab4, s_inv, ring_k[krow(k - 10)], ring_k[krow(k - 9)], ring_k[krow(k - 8)], t, m, bq,
                                        out);
                        rlo = out[0];
                        rhi = out[1];
                    } else {
                        stencil_rowsS(kw2, t, m, bq, rlo, rhi);
                    }
                    const float4 f2 = ring_f[frow(ph - 9) * 32 + lane];
                    rlo = sub2(pack2(f2.x, f2.z), rlo);  // (r0, r2)
                    rhi = sub2(pack2(f2.y, f2.w), rhi);  // (r1, r3)
                    float4 r;
                    unpack2(rlo, r.x, r.z);
                    unpack2(rhi, r.y, r.w);
                    if (MODE == 1) {
                        if (lane_int && (!GUARD || (yr >= y0 && yr < y1)) && (!EDGE || (yr >= 1 && yr <= N - 2))) {
                            const float4 q = EDGE ? mask4(r, cin) : r;
                            float s4 = __fmul_rn(q.x, q.x);
                            s4 = __fmaf_rn(q.y, q.y, s4);
                            s4 = __fmaf_rn(q.z, q.z, s4);
                            s4 = __fmaf_rn(q.w, q.w, s4);
                            part += (double)s4;
                        }
                    } else {
                        // restriction fed row by row (row-major taps).  a0 is even: k even <=> residual row a-9 odd = bottom
                        // row of the coarse stencil centred on fine row a-10 and top row of the next one
                        const float rl = __shfl_up_sync(0xffffffffu, r.w, 1);  // column 4l-1
                        const unsigned int kwr = (KEYED && rkeys) ? ring_k[krow(k - 9)] : 0u;
                        if ((ph & 1) == 0) {
                            if (!GUARD || k >= 12) {
                                const int yc = a - 10;
                                float c0, c1;
                                if (KEYED && rkeys) {
                                    c0 = racc0;
                                    c1 = racc1;
                                    hs_restrict_feed_keys(s_rp, kwr, 2, rl, r, c0, c1);
                                } else {
                                    c0 = __fmaf_rn(tw[6], rl, racc0);
                                    c1 = __fmaf_rn(tw[6], r.y, racc1);
                                    c0 = __fmaf_rn(tw[7], r.x, c0);
                                    c1 = __fmaf_rn(tw[7], r.z, c1);
                                    c0 = __fmaf_rn(tw[8], r.y, c0);
                                    c1 = __fmaf_rn(tw[8], r.w, c1);
                                }
                                if (has_scale) {
                                    c0 = __fmul_rn(tscale, c0);
                                    c1 = __fmul_rn(tscale, c1);
                                }
                                if (EDGE) {
                                    const int I = yc >> 1;
                                    const bool Iin = (I >= 1 && I <= p.Nc - 2);
                                    c0 = (Iin && cxl >= 1 && cxl <= p.Nc - 2) ? c0 : 0.0f;
                                    c1 = (Iin && cxl + 1 >= 1 && cxl + 1 <= p.Nc - 2) ? c1 : 0.0f;
                                }
                                st_global_v2_pred(st_c, c0, c1, fc_ok && (!GUARD || (yc >= y0 && yc < y1)));
                            }
                            if (KEYED && rkeys) {
                                hs_restrict_feed_keys(s_rp, kwr, 0, rl, r, racc0, racc1);
                            } else {
                                racc0 = __fmul_rn(tw[0], rl);
                                racc1 = __fmul_rn(tw[0], r.y);
                                racc0 = __fmaf_rn(tw[1], r.x, racc0);
                                racc1 = __fmaf_rn(tw[1], r.z, racc1);
                                racc0 = __fmaf_rn(tw[2], r.y, racc0);
                                racc1 = __fmaf_rn(tw[2], r.w, racc1);
                            }
                        } else {
                            if (KEYED && rkeys) {
                                hs_restrict_feed_keys(s_rp, kwr, 1, rl, r, racc0, racc1);
                            } else {
                                racc0 = __fmaf_rn(tw[3], rl, racc0);
                                racc1 = __fmaf_rn(tw[3], r.y, racc1);
                                racc0 = __fmaf_rn(tw[4], r.x, racc0);
                                racc1 = __fmaf_rn(tw[4], r.z, racc1);
                                racc0 = __fmaf_rn(tw[5], r.y, racc0);
                                racc1 = __fmaf_rn(tw[5], r.w, racc1);
                            }
                        }
                    }
                }
                // ================= stage 4: HNet layer 3 row a-7, u' = J + h, store; new u' row into Bw
                if (!GUARD || k >= 8) {
                    const int y = a - 7;
                    const ulonglong2 jv = *reinterpret_cast<const ulonglong2 *>(&ring_j[ph * 32 + lane]);  // J row a-7: (A, B)
                    u64 ul = jv.x, uh = jv.y;
                    if (!NOL) {
                        u64 lo, hi;
                        stencil_rowsS(h2[2], X2[(ph + 2) % 3], X2[ph % 3], X2[(ph + 1) % 3], lo, hi);  // rows a-8, a-7, a-6
                        if (EDGE) hs_maskS(lo, hi, (y >= 1 && y <= N - 2) ? cin : 0u);
                        ul = add2(jv.x, lo);
                        uh = add2(jv.y, hi);
                    }
                    float o0, o1, o2, o3;
                    unpack2(ul, o0, o2);
                    unpack2(uh, o1, o3);
                    const bool st_ok = lane_int && (!GUARD || (y >= y0 && y < y1)) && (!EDGE || (cdom & 1u));
                    st_global_v4_pred(st_u, o0, o1, o2, o3, st_ok);
                    Bw[(ph + 2) % 3] = widenS(ul, uh);
                }
                // ================= stage 3: layer 2 row a-5
                if (!NOL && (!GUARD || k >= 6)) {
                    u64 lo, hi;
                    stencil_rowsS(h2[1], X1[(ph + 2) % 3], X1[ph % 3], X1[(ph + 1) % 3], lo, hi);  // rows a-6, a-5, a-4
                    if (EDGE) hs_maskS(lo, hi, (a - 5 >= 1 && a - 5 <= N - 2) ? cin : 0u);
                    X2[(ph + 2) % 3] = widenS(lo, hi);
                }
                // ================= stage 2: layer 1 row a-3
                if (!NOL && (!GUARD || k >= 4)) {
                    u64 lo, hi;
                    stencil_rowsS(h2[0], X0[(ph + 2) % 3], X0[ph % 3], X0[(ph + 1) % 3], lo, hi);  // rows a-4, a-3, a-2
                    if (EDGE) hs_maskS(lo, hi, (a - 3 >= 1 && a - 3 <= N - 2) ? cin : 0u);
                    X1[(ph + 2) % 3] = widenS(lo, hi);
                }
                // ================= stage 0: input row a (+ prolongation / correction on the up leg)
                float4 uv = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : ring_u[urow(ph) * 32 + lane];
                const bool arow_in = !EDGE || (a >= 1 && a <= N - 2);
                if (MODE == 1) {
                    const float2 cv = ring_c[urow(ph) * 32 + lane];
                    const float c2 = __shfl_down_sync(0xffffffffu, cv.x, 1);
                    const bool odd = (ph & 1) == 0;  // a0 is odd
                    const unsigned int kb = pkeys ? ckw[ph >> 1] : 0u;  // keys of coarse row ceil(a/2) (a odd at ph = 0)
                    if (!EDGE || (a >= 0 && a <= N - 1)) {
                        float4 e;
                        if (ptable) {
                            float top[3], bot[3];
                            if (!odd) {
                                vt[0] = cv.x;
                                vt[1] = cv.y;
                                vt[2] = c2;
                                kvt = kb;
                            }
                            top[0] = vt[0];
                            top[1] = vt[1];
                            top[2] = vt[2];
                            bot[0] = cv.x;
                            bot[1] = cv.y;
                            bot[2] = c2;
                            if (KEYED) {
                                e = hs_prolong_table_keys(s_rp, odd ? 1 : 0, top[0], top[1], top[2], bot[0], bot[1], bot[2],
                                                          kvt, kb);
                            } else {
                                const int kz[3] = {0, 0, 0};
                                e = hs_prolong_table(tw, odd, top, bot, kz, kz);
                            }
                            if (has_scale) {
                                e.x = __fmul_rn(tscale, e.x);
                                e.y = __fmul_rn(tscale, e.y);
                                e.z = __fmul_rn(tscale, e.z);
                                e.w = __fmul_rn(tscale, e.w);
                            }
                        } else {
                            // bilinear: fl(a/2 + b/2) == fma(0.5, a, 0.5 b) (halving is exact)
                            if (!odd) {
                                vt[0] = cv.x;
                                vt[1] = cv.y;
                                vt[2] = c2;
                                e.x = vt[0];
                                e.y = __fmaf_rn(0.5f, vt[0], __fmul_rn(0.5f, vt[1]));
                                e.z = vt[1];
                                e.w = __fmaf_rn(0.5f, vt[1], __fmul_rn(0.5f, vt[2]));
                            } else {
                                const float vb0 = cv.x, vb1 = cv.y, vb2 = c2;
                                e.x = __fmaf_rn(0.5f, vt[0], __fmul_rn(0.5f, vb0));
                                e.z = __fmaf_rn(0.5f, vt[1], __fmul_rn(0.5f, vb1));
                                const float ta = __fmaf_rn(0.5f, vt[0], __fmul_rn(0.5f, vt[1]));
                                const float ba = __fmaf_rn(0.5f, vb0, __fmul_rn(0.5f, vb1));
                                e.y = __fmaf_rn(0.5f, ta, __fmul_rn(0.5f, ba));
                                const float tb = __fmaf_rn(0.5f, vt[1], __fmul_rn(0.5f, vt[2]));
                                const float bb = __fmaf_rn(0.5f, vb1, __fmul_rn(0.5f, vb2));
                                e.w = __fmaf_rn(0.5f, tb, __fmul_rn(0.5f, bb));
                            }
                            if (EDGE) e = mask4(e, arow_in ? cin : 0u);  // fine level's reset_boundary of the correction
                        }
                        uv.x = __fadd_rn(uv.x, e.x);
                        uv.y = __fadd_rn(uv.y, e.y);
                        uv.z = __fadd_rn(uv.z, e.z);
                        uv.w = __fadd_rn(uv.w, e.w);
                        if (EDGE) uv = mask4(uv, cdom);
                    } else if (!odd) {  // row outside the domain: still advance the carried coarse row
                        vt[0] = cv.x;
                        vt[1] = cv.y;
                        vt[2] = c2;
                        kvt = kb;
                    }
                }
                // x = J - (un-reset u): keep the raw row of edge strips; reset_boundary of the sweep's input
                const u64 rawp_lo = rawlo, rawp_hi = rawhi;  // raw row a-1
                if (EDGE) {
                    rawlo = pack2(uv.x, uv.z);
                    rawhi = pack2(uv.y, uv.w);
                    uv = mask4(uv, arow_in ? cin : 0u);
                }
                A[(ph + 2) % 3] = widenS(pack2(uv.x, uv.z), pack2(uv.y, uv.w));
                // ================= stage 1: Jacobi row a-1, x = J - u; J into the delay line (read above by stage 4)
                if (!GUARD || k >= 2) {
                    const int y = a - 1;
                    const RS &t = A[(ph + 0) % 3], &m = A[(ph + 1) % 3], &bq = A[(ph + 2) % 3];
                    u64 klo_, khi_, ivlo = inv2, ivhi = inv2;
                    if (KEYED) {
                        u64 out[4];
                        hs_stencil_keys(s_tab4, s_inv, ring_k[krow(k - 2)], ring_k[krow(k - 1)], ring_k[krow(k)], t, m, bq, out);
                        klo_ = out[0];
                        khi_ = out[1];
                        ivlo = out[2];
                        ivhi = out[3];
                    } else {
                        stencil_rowsS(kw2, t, m, bq, klo_, khi_);
                    }
                    const float4 fv = ring_f[frow(ph - 1) * 32 + lane];
                    const ulonglong2 ff = make_ulonglong2(pack2(fv.x, fv.z), pack2(fv.y, fv.w));
                    // u + inv * (f - K u) with TWO roundings: fma(p, 1, u) = fl(p + u), 1 is a launch parameter (ptxas would
                    // contract mul + add into one FFMA2 otherwise; see mg_stream2_kernel)
                    const u64 plo = mul2(ivlo, sub2(ff.x, klo_)), phi = mul2(ivhi, sub2(ff.y, khi_));
                    u64 jl = fma2(plo, one2, m.p[1]), jh = fma2(phi, one2, m.p[2]);
                    if (EDGE) hs_maskS(jl, jh, (y >= 1 && y <= N - 2) ? cin : 0u);
                    *reinterpret_cast<ulonglong2 *>(&ring_j[ph * 32 + lane]) = make_ulonglong2(jl, jh);
                    if (!NOL) {
                        const u64 xl = sub2(jl, EDGE ? rawp_lo : m.p[1]), xh = sub2(jh, EDGE ? rawp_hi : m.p[2]);
                        X0[(ph + 2) % 3] = widenS(xl, xh);
                    }
                }
                st_u += p.pitch;
                if (MODE == 0 && (ph & 1) == 0) st_c += p.pitch_c;
            }
            if (MODE == 1 && pkeys) {
#pragma unroll
                for (int j = 0; j < 3; ++j) ckw[j] = ckn[j];
            }
        };
        using T_ = std::true_type;
        using F_ = std::false_type;
        int prev_key = -2, prev2_key = -2, prev_ckey = -2;  // uniform key of the previous blocks (-1 mixed, -2 none)
        for (int k0 = 0; k0 < K; k0 += 6) {
            bool fast = true;
            if (KEYS) {
                // this block's keys travel in the oldest outstanding group (fetched one block ago with row k0's prefetch)
                asm volatile("cp.async.wait_group %0;" ::"n"(PD - 1) : "memory");
                const unsigned int w0 = ring_k[(k0 % HS_KD) * 32 + lane];
                const unsigned int pat = (__shfl_sync(0xffffffffu, w0, 0) & 0xffu) * 0x01010101u;
                unsigned int diff = w0 ^ pat;
#pragma unroll
                for (int j = 1; j < 6; ++j) diff |= ring_k[((k0 + j) % HS_KD) * 32 + lane] ^ pat;
                // the residual stage looks 10 rows back: the two previous blocks must carry the same single pattern
                const int bkey = __all_sync(0xffffffffu, diff == 0u) ? (int)(pat & 0xffu) : -1;
                fast = (bkey >= 0) && (prev_key == -2 || prev_key == bkey) && (prev2_key == -2 || prev2_key == bkey);
                prev2_key = prev_key;
                prev_key = bkey;
                if (fast && bkey != kcur) {
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const float w = s_tab[9 * bkey + q];
                        kw2[q] = pack2(w, w);
                        if (MODE == 0 && ntab > 1) tw[q] = s_rp[9 * bkey + q];
                    }
                    const float iv = s_inv[bkey];
                    inv2 = pack2(iv, iv);
                    kcur = bkey;
                }
            }
            if (MODE == 1 && pkeys) {  // the prolongation table follows the COARSE nodes' keys
                const unsigned int pat = (__shfl_sync(0xffffffffu, ckw[0], 0) & 0xffu) * 0x0101u;
                const unsigned int diff = (ckw[0] ^ pat) | (ckw[1] ^ pat) | (ckw[2] ^ pat) | (k0 == 0 ? (kvt ^ pat) : 0u);
                const int ckey = __all_sync(0xffffffffu, diff == 0u) ? (int)(pat & 0xffu) : -1;
                const bool cfast = (ckey >= 0) && (prev_ckey == -2 || prev_ckey == ckey);
                prev_ckey = ckey;
                if (fast && cfast && ckey != ccur) {
#pragma unroll
                    for (int q = 0; q < 9; ++q) tw[q] = s_rp[9 * ckey + q];
                    ccur = ckey;
                }
                fast = fast && cfast;
            }
            if (fast) block6(F_{}, k0);
            else block6(T_{}, k0);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (MODE == 1 && p.want_norm) {
            part = warp_sum_d(part);
            if (lane == 0) p.partials[s] = part;
        }
    }

    if (dynq) {
        __syncthreads();  // every warp of this CTA has taken its last strip number
        if (threadIdx.x == 0 && atomicAdd(leave, 1u) == gridDim.x - 1) {
            *queue = 0u;
            *leave = 0u;
            __threadfence();
        }
    }
    if (solve_done) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        return;
    }
    // ---- deterministic final reduction of the per-strip partial sums by the last CTA to finish
    if (MODE == 1 && p.want_norm) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int ticket = atomicAdd(p.counter, 1u);
            lastflag = (ticket == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (lastflag) {
            __threadfence();
            Ctl *ctl = reinterpret_cast<Ctl *>(p.ctl);
            double tot = 0.0, mx = 0.0;
            for (int b = 0; b < p.B; ++b) {
                double v = 0.0;
                for (int i = threadIdx.x; i < p.nstrips; i += blockDim.x)
                    v += __ldcg(p.partials + (long long)b * p.nstrips + i);
                v = warp_sum_d(v);
                __syncthreads();
                if (lane == 0) red[warp] = v;
                __syncthreads();
                if (threadIdx.x == 0) {
                    double sum = 0.0;
                    for (int w = 0; w < HS_WARPS; ++w) sum += red[w];
                    if (p.sumsq) p.sumsq[b] = sum;
                    if (ctl && p.hist && ctl->cycle < ctl->max_cycles) p.hist[(long long)ctl->cycle * p.B + b] = sum;
                    tot += sum;
                    mx = sum > mx ? sum : mx;
                }
            }
            if (threadIdx.x == 0) {
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
    }
}

}  // namespace mgfea
