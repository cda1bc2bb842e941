// Streaming (register-chained) V(1,1) kernels for the big levels of single-pattern / default-Dirichlet problems.
//
// One WARP owns a strip of 120 interior columns (128 with a 4-column halo on each side; lane l owns box columns
// 4l..4l+3) and marches down R rows.  All stencil stages of a leg are chained through rotating register windows:
//
//   down leg:  u0 row a  ->  u1 row a-1 (Jacobi)  ->  r row a-2 (= f - K u1)  ->  coarse f row (a-3)/2 (restriction)
//   up leg:    u row a (+ bilinear P v_c)  ->  u2 row a-1 (Jacobi)  ->  r row a-2  ->  sum r^2 (interior residual norm)
//
// so u and f are read from HBM once, u is written once, and nothing is staged in shared memory between stages: no
// __syncthreads, no per-tile prologue.  Global rows are prefetched D rows ahead with cp.async into a PER-LANE ring
// (every lane later reads back exactly the 16 bytes it copied itself, so no cross-lane synchronisation is needed
// either); neighbour columns come from warp shuffles.  cp.async zero-fills rows/columns outside the domain, which is
// the reference's zero padding.  The vertical halo (3+1 rows per strip) is recomputed, the 4-column halos too.
//
// Arithmetic: identical operation order to the tile kernels / oracle (row-major FMA chain, explicit roundings).
#pragma once
#include <type_traits>

#include "mgfea_tile.cuh"

namespace mgfea {

constexpr int ST_WARPS = 8;    // warps per CTA
constexpr int ST_DEPTH = 6;    // prefetch ring depth (rows) == unroll factor of the row loop: ring slots are compile-time
constexpr int ST_TWI = BW - 8; // interior columns per strip (HX = 4)
constexpr int ST_RING_F4 = 2 * ST_DEPTH + ST_DEPTH / 2;  // float4 units per lane: u ring, f ring, coarse float2 ring
constexpr int ST_KDEPTH = 3 * ST_DEPTH;                  // key-word ring depth in rows (KEYS kernels only)

// mg_stream2_kernel: ring depth in rows.  The down leg of single-pattern levels keeps 12 rows (two halves of 6, the
// unroll factor, so slots stay compile-time) and prefetches 9 rows ahead: with 6 rows / 3 ahead only ~49 KB per SM were
// in flight, less than HBM latency x bandwidth asks for (ncu: long_scoreboard was the top stall).  The up leg (coarse
// ring) and the keyed kernels (key ring) have no shared memory to spare at 2 CTAs per SM and stay at 6.
__host__ __device__ constexpr int st2_ring_rows(int mode, bool keys) { return (mode == 0 && !keys) ? 12 : ST_DEPTH; }
__host__ __device__ constexpr int st2_ring_f4(int mode, bool keys) {  // float4 units per lane
    return 2 * st2_ring_rows(mode, keys) + (mode == 1 ? ST_DEPTH / 2 : 0);
}
__device__ __forceinline__ void st_global_v4_pred(float *p, float a, float b, float c, float d, bool ok) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q st.global.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p),
        "f"(a), "f"(b), "f"(c), "f"(d), "r"((int)ok)
        : "memory");
}
__device__ __forceinline__ void st_global_v2_pred(float *p, float a, float b, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.global.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(p), "f"(a),
                 "f"(b), "r"((int)ok)
                 : "memory");
}

struct StreamParams {
    int N, B, pitch;
    long long plane;
    int R, ntx, nry, nstrips;  // rows per strip, strips per row of strips, strip rows, strips per sample
    // row-slab partition (multi-GPU): the arrays hold local rows [row0, row0+nrloc) of the global N x N level (owned rows
    // plus ghost rows filled by the halo exchange); this rank computes / stores the owned global rows [own0, own1).
    // Coarse arrays hold global coarse rows [crow0, crow0+nrc).  Single GPU: row0=0, nrloc=N, own=[0,N), crow0=0, nrc=Nc.
    int row0, nrloc, own0, own1, crow0, nrc;
    float inv_nstrips, inv_ntx;
    float one;  // = 1.0f, deliberately a run-time value
    const float *u_in;   // NULL: zero initial guess (down leg of coarse levels)
    float *u_out;
    const float *f;
    const float *ktab, *invd;  // [npat][9], [npat] (single pattern: [9], [1])
    // material pattern keys (mg_stream2_kernel<.., KEYS = true>): GLOBAL [N][key_pitch] map, also on a slab
    const unsigned char *keys;
    int key_pitch, npat;
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
    int prolong_seq, want_norm;
    double *partials;
    unsigned int *counter;
    double *sumsq;
    double *hist;
    void *ctl;
    // fused halo push (mg_stream2_kernel<.., PUSH>: the finest up leg of the row-slab cycle).  The strips that produce the
    // `push_rows` first / last rows of the owned range [pown0, pown1) store them ALSO straight into the neighbours' ghost
    // rows over NVLink (push_up / push_dn: peer-mapped address of GLOBAL row 0, column 0 of the neighbour's array; NULL:
    // no neighbour), so the transfer overlaps the sweep.  Every pushing strip takes a ticket when its stores are fenced;
    // the last of the npush_* strips raises the flag in the neighbour's mailbox once (release, system scope).
    float *push_up, *push_dn;
    int push_rows, pown0, pown1, npush_up, npush_dn;
    unsigned int *push_ticket;  // 2 local words (up, dn), left at 0
    unsigned int *push_flag_up, *push_flag_dn;
    int ctl_ro;  // row slabs: `ctl` is only read (done flag); the stopping rule belongs to the all-reduce step
    // learned smoother + table transfer operators (mg_hstream_kernel, mgfea_hstream.cuh)
    const float *hw;              // [3][9] HNet layer kernels
    int rtab_n;                   // restriction tables: 1, or one per pattern (indexed by the fine source node's key)
    int prolong_mode;             // MGFEA_PROLONG_BILINEAR | MGFEA_PROLONG_TABLE
    const float *ptab;            // [ptab_n][9]
    int ptab_n, p_has_scale;
    float p_scale;
    const float *p_scale_dev;
    const unsigned char *keys_c;  // coarse level's key map (ptab_n > 1)
    int key_pitch_c;
    int dyn_queue;                // strips handed out by an atomic queue (more than one strip per resident warp)
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void st_cp16(void *dst, const void *src, bool ok) {
    const uint32_t sz = ok ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}

struct R6 {
    float a[6];
};
__device__ __forceinline__ R6 widen(float4 v) {
    R6 o;
    o.a[0] = __shfl_up_sync(0xffffffffu, v.w, 1);
    o.a[1] = v.x;
    o.a[2] = v.y;
    o.a[3] = v.z;
    o.a[4] = v.w;
    o.a[5] = __shfl_down_sync(0xffffffffu, v.x, 1);
    return o;
}
__device__ __forceinline__ float4 stencil_rows(const float (&w)[9], const R6 &t, const R6 &m, const R6 &b) {
    float acc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float s = __fmul_rn(w[0], t.a[e]);
        s = __fmaf_rn(w[1], t.a[e + 1], s);
        s = __fmaf_rn(w[2], t.a[e + 2], s);
        s = __fmaf_rn(w[3], m.a[e], s);
        s = __fmaf_rn(w[4], m.a[e + 1], s);
        s = __fmaf_rn(w[5], m.a[e + 2], s);
        s = __fmaf_rn(w[6], b.a[e], s);
        s = __fmaf_rn(w[7], b.a[e + 1], s);
        s = __fmaf_rn(w[8], b.a[e + 2], s);
        acc[e] = s;
    }
    return make_float4(acc[0], acc[1], acc[2], acc[3]);
}
__device__ __forceinline__ float4 mask4(float4 v, unsigned int m) {
    v.x = (m & 1u) ? v.x : 0.0f;
    v.y = (m & 2u) ? v.y : 0.0f;
    v.z = (m & 4u) ? v.z : 0.0f;
    v.w = (m & 8u) ? v.w : 0.0f;
    return v;
}


// ---- packed fp32x2 helpers (Blackwell FFMA2 / FADD2 / FMUL2): IEEE round-to-nearest per lane, like the scalar ops
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
struct RP {
    u64 q[5];  // q[i] = (a[i], a[i+1]) for the 6 values a[0..5] = box columns 4l-1 .. 4l+4
};
__device__ __forceinline__ RP widen2(float a1, float a2, float a3, float a4) {
    const float a0 = __shfl_up_sync(0xffffffffu, a4, 1);
    const float a5 = __shfl_down_sync(0xffffffffu, a1, 1);
    RP o;
    o.q[0] = pack2(a0, a1);
    o.q[1] = pack2(a1, a2);
    o.q[2] = pack2(a2, a3);
    o.q[3] = pack2(a3, a4);
    o.q[4] = pack2(a4, a5);
    return o;
}
// (e0,e1) -> lo, (e2,e3) -> hi; per lane the same row-major FMA chain as stencil_rows
__device__ __forceinline__ void stencil_rows2(const u64 (&w)[9], const RP &t, const RP &m, const RP &b, u64 &lo,
                                              u64 &hi) {
    lo = mul2(w[0], t.q[0]);
    hi = mul2(w[0], t.q[2]);
    lo = fma2(w[1], t.q[1], lo);
    hi = fma2(w[1], t.q[3], hi);
    lo = fma2(w[2], t.q[2], lo);
    hi = fma2(w[2], t.q[4], hi);
    lo = fma2(w[3], m.q[0], lo);
    hi = fma2(w[3], m.q[2], hi);
    lo = fma2(w[4], m.q[1], lo);
    hi = fma2(w[4], m.q[3], hi);
    lo = fma2(w[5], m.q[2], lo);
    hi = fma2(w[5], m.q[4], hi);
    lo = fma2(w[6], b.q[0], lo);
    hi = fma2(w[6], b.q[2], hi);
    lo = fma2(w[7], b.q[1], lo);
    hi = fma2(w[7], b.q[3], hi);
    lo = fma2(w[8], b.q[2], lo);
    hi = fma2(w[8], b.q[4], hi);
}

// MODE 0: down leg (Jacobi sweep, store u, residual, restriction -> fc)
// MODE 1: up leg   (bilinear prolongation + correction, Jacobi sweep, store u, optional interior residual norm)
template <int MODE, bool ZERO_INIT>
__global__ void __launch_bounds__(ST_WARPS * 32, 2) mg_stream_kernel(const StreamParams p) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    __shared__ double red[ST_WARPS];
    __shared__ int lastflag;
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N;

    // ---- weights in registers for the whole kernel
    float kw[9], rw[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        kw[q] = p.ktab[q];
        rw[q] = (MODE == 0) ? p.rtab[q] : 0.0f;
    }
    const float inv = p.invd[0];
    const float rscale = (MODE == 0 && p.r_has_scale) ? (p.r_scale_dev ? *p.r_scale_dev : p.r_scale) : 1.0f;
    pdl_wait();  // weights above are never written by a kernel; all field data is touched only after this point
    // the solve-control word is requested here but tested only after the first prefetches are in flight, so the two
    // memory round trips overlap (a finished solve returns before anything is stored)
    const int solve_done = (p.ctl != nullptr) ? ld_volatile_s32(&reinterpret_cast<const Ctl *>(p.ctl)->done) : 0;

    // ---- per-lane prefetch ring: [slot][lane] float4 for u and for f, float2 for the coarse row (up leg)
    float4 *ring_u = reinterpret_cast<float4 *>(st_smem) + warp * (ST_RING_F4 * 32);
    float4 *ring_f = ring_u + ST_DEPTH * 32;
    float2 *ring_c = reinterpret_cast<float2 *>(ring_f + ST_DEPTH * 32);

    const int total = p.nstrips * p.B;
    for (int s = blockIdx.x * ST_WARPS + warp; s < total; s += gridDim.x * ST_WARPS) {
        // strip coordinates
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
        const int y0 = p.own0 + ry * p.R;                       // global rows
        const int y1 = (ry == p.nry - 1) ? p.own1 : y0 + p.R;  // exclusive
        const int gx = tx * ST_TWI - 4 + 4 * lane;         // first global column of this lane
        const bool lane_int = (lane >= 1 && lane <= 30);
        // column masks of this lane's 4 columns
        unsigned int cin = 0, cdom = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (gx + e >= 1 && gx + e <= N - 2) cin |= 1u << e;
            if (gx + e >= 0 && gx + e <= N - 1) cdom |= 1u << e;
        }
        const bool col_ok = (gx >= 0) && (gx + 3 < p.pitch);  // the 16-byte chunk exists in memory
        const float *ub = ZERO_INIT ? nullptr : p.u_in + (long long)b * p.plane + gx;
        const float *fb = p.f + (long long)b * p.plane + gx;
        float *uo = p.u_out + (long long)b * p.plane + gx;
        const int cxl = gx >> 1;  // coarse column of this lane's first fine column (gx is even)
        const bool ccol_ok = (MODE == 1) && (cxl >= 0) && (cxl + 1 < p.pitch_c);
        const float *cbp = (MODE == 1) ? p.vc + (long long)b * p.plane_c + cxl : nullptr;
        float *fco = (MODE == 0) ? p.fc + (long long)b * p.plane_c + cxl : nullptr;
        const bool fc_ok = lane_int && cxl >= 0 && cxl <= p.Nc - 1;
        // a strip whose whole streamed box lies strictly inside the domain needs no masks at all
        const bool edge = (y0 - 3 <= 0) || (y1 + 2 >= N - 1) || (tx == 0) || ((tx + 1) * ST_TWI + 4 >= N - 1);

        const int a0 = y0 - 3;  // first streamed row
        // last streamed row: y1+1 (u1 row y1 for the residual row y1-1); the last strip of the down leg goes one
        // further so that the coarse ring row (N-1)/2 is produced (as zeros) too
        const int K = ((MODE == 0 && y1 == N) ? y1 + 2 : y1 + 1) - a0 + 1;
        // prefetch of streamed row k (rows are requested in increasing k): running global pointers, compile-time ring
        // slot, and no validity test in steady-state blocks (CHECK = false)
        const int klo = max(0, p.row0) - a0;                    // first k whose row exists locally / in the domain
        const int khi = min(min(N, p.row0 + p.nrloc) - a0, K);  // one past the last such k
        const float *pf_u = ZERO_INIT ? nullptr : ub + (long long)(a0 - p.row0) * p.pitch;
        const float *pf_f = fb + (long long)(a0 - p.row0) * p.pitch;
        int kpf = 0;
        auto prefetch = [&](auto check_tag, int slot_row) {
            constexpr bool CHECK = decltype(check_tag)::value;
            const bool ok = !CHECK || (col_ok && kpf >= klo && kpf < khi);
            const int slot = slot_row * 32 + lane;
            if (!ZERO_INIT) st_cp16(&ring_u[slot], ok ? (const void *)pf_u : (const void *)p.f, ok);
            st_cp16(&ring_f[slot], ok ? (const void *)pf_f : (const void *)p.f, ok);
            if (MODE == 1) {  // coarse row ceil(a/2): the row an even fine row copies / an odd row's lower partner
                const int I = (a0 + kpf + 1) >> 1;
                const int lI = I - p.crow0;
                const bool okc = !CHECK || (ccol_ok && (I >= 0) && (I < p.Nc) && (lI >= 0) && (lI < p.nrc) && kpf < K);
                const void *src = okc ? (const void *)(cbp + (long long)lI * p.pitch_c) : (const void *)p.f;
                const uint32_t sz = okc ? 8u : 0u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(&ring_c[slot])), "l"(src),
                             "r"(sz)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (!ZERO_INIT) pf_u += p.pitch;
            pf_f += p.pitch;
            ++kpf;
        };
#pragma unroll
        for (int k = 0; k < ST_DEPTH - 1; ++k) prefetch(std::true_type{}, k);
        if (solve_done) break;

        // rotating windows (indices are compile-time after unrolling by 6)
        R6 A[3];      // input rows (u0, or corrected u for the up leg): a-2, a-1, a
        R6 Bw[3];     // smoothed rows: a-3, a-2, a-1
        R6 C[3];      // residual rows (down leg): a-4, a-3, a-2
        float4 F[3];  // f rows a-2, a-1, a
        float vt[3] = {0.f, 0.f, 0.f};  // coarse row floor(a/2) at coarse columns cxl, cxl+1, cxl+2 (up leg)
        double part = 0.0;
        if (MODE == 1 && (a0 & 1) && a0 >= 1) {  // the strip starts on an odd row: fetch its upper coarse row directly
            const int lI = (a0 >> 1) - p.crow0;
            const float *row = p.vc + (long long)b * p.plane_c + (long long)lI * p.pitch_c;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int c = cxl + q;
                vt[q] = (c >= 0 && c < p.Nc && lI >= 0 && lI < p.nrc) ? __ldg(row + c) : 0.0f;
            }
        }

        // one block of 6 row steps.  GUARD: start-up / drain steps (pipeline fill, store ranges); EDGE: masks needed.
        auto block6 = [&](auto guard_tag, auto edge_tag, auto pf_tag, int k0) {
            constexpr bool GUARD = decltype(guard_tag)::value;
            constexpr bool EDGE = decltype(edge_tag)::value;
#pragma unroll
            for (int ph = 0; ph < 6; ++ph) {
                const int k = k0 + ph;
                if (GUARD && k >= K) break;
                const int a = a0 + k;
                prefetch(pf_tag, (ph + ST_DEPTH - 1) % ST_DEPTH);  // k0 is a multiple of ST_DEPTH: slot == phase
                asm volatile("cp.async.wait_group %0;" ::"n"(ST_DEPTH - 1) : "memory");
                const int slot = ph * 32 + lane;
                float4 uv = ZERO_INIT ? make_float4(0.f, 0.f, 0.f, 0.f) : ring_u[slot];
                const float4 fv = ring_f[slot];
                const bool arow_in = !EDGE || (a >= 1 && a <= N - 2);
                if (MODE == 1) {
                    // u += reset(bilinear P v_c) on row a
                    const float2 cv = ring_c[slot];
                    const float c2 = __shfl_down_sync(0xffffffffu, cv.x, 1);
                    if (!EDGE || (a >= 0 && a <= N - 1)) {
                        float4 e;
                        if ((ph & 1) != 0) {  // k odd <=> a even (a0 is odd): copy / horizontal average of coarse row a/2
                            vt[0] = cv.x;
                            vt[1] = cv.y;
                            vt[2] = c2;
                            e.x = vt[0];
                            e.y = __fadd_rn(__fmul_rn(0.5f, vt[0]), __fmul_rn(0.5f, vt[1]));
                            e.z = vt[1];
                            e.w = __fadd_rn(__fmul_rn(0.5f, vt[1]), __fmul_rn(0.5f, vt[2]));
                        } else {  // a odd: between coarse rows (a-1)/2 (vt, kept from the previous step) and (a+1)/2
                            const float vb0 = cv.x, vb1 = cv.y, vb2 = c2;
                            e.x = __fadd_rn(__fmul_rn(0.5f, vt[0]), __fmul_rn(0.5f, vb0));
                            e.z = __fadd_rn(__fmul_rn(0.5f, vt[1]), __fmul_rn(0.5f, vb1));
                            if (p.prolong_seq) {
                                float v = __fadd_rn(__fmul_rn(0.25f, vt[0]), __fmul_rn(0.25f, vt[1]));
                                v = __fadd_rn(v, __fmul_rn(0.25f, vb0));
                                e.y = __fadd_rn(v, __fmul_rn(0.25f, vb1));
                                v = __fadd_rn(__fmul_rn(0.25f, vt[1]), __fmul_rn(0.25f, vt[2]));
                                v = __fadd_rn(v, __fmul_rn(0.25f, vb1));
                                e.w = __fadd_rn(v, __fmul_rn(0.25f, vb2));
                            } else {
                                const float ta = __fadd_rn(__fmul_rn(0.5f, vt[0]), __fmul_rn(0.5f, vt[1]));
                                const float ba = __fadd_rn(__fmul_rn(0.5f, vb0), __fmul_rn(0.5f, vb1));
                                e.y = __fadd_rn(__fmul_rn(0.5f, ta), __fmul_rn(0.5f, ba));
                                const float tb = __fadd_rn(__fmul_rn(0.5f, vt[1]), __fmul_rn(0.5f, vt[2]));
                                const float bb = __fadd_rn(__fmul_rn(0.5f, vb1), __fmul_rn(0.5f, vb2));
                                e.w = __fadd_rn(__fmul_rn(0.5f, tb), __fmul_rn(0.5f, bb));
                            }
                        }
                        if (EDGE) e = mask4(e, arow_in ? cin : 0u);  // fine level's reset_boundary of the correction
                        uv.x = __fadd_rn(uv.x, e.x);
                        uv.y = __fadd_rn(uv.y, e.y);
                        uv.z = __fadd_rn(uv.z, e.z);
                        uv.w = __fadd_rn(uv.w, e.w);
                        if (EDGE) uv = mask4(uv, cdom);
                    }
                }
                // reset_boundary of the sweep's input (ring -> 0, outside stays 0)
                if (EDGE) uv = mask4(uv, arow_in ? cin : 0u);
                A[(ph + 2) % 3] = widen(uv);
                F[(ph + 2) % 3] = fv;
                // ---- Jacobi row a-1
                if (!GUARD || k >= 2) {
                    const int y = a - 1;
                    const R6 &t = A[(ph + 0) % 3], &m = A[(ph + 1) % 3], &bq = A[(ph + 2) % 3];
                    const float4 ku = stencil_rows(kw, t, m, bq);
                    const float4 ff = F[(ph + 1) % 3];
                    float4 o;
                    o.x = __fadd_rn(__fmul_rn(inv, __fsub_rn(ff.x, ku.x)), m.a[1]);
                    o.y = __fadd_rn(__fmul_rn(inv, __fsub_rn(ff.y, ku.y)), m.a[2]);
                    o.z = __fadd_rn(__fmul_rn(inv, __fsub_rn(ff.z, ku.z)), m.a[3]);
                    o.w = __fadd_rn(__fmul_rn(inv, __fsub_rn(ff.w, ku.w)), m.a[4]);
                    if (EDGE) o = mask4(o, (y >= 1 && y <= N - 2) ? cin : 0u);
                    const bool st_ok = lane_int && (!GUARD || (y >= y0 && y < y1)) && (!EDGE || (cdom & 1u));
                    if (st_ok) st_global_v4(uo + (long long)(y - p.row0) * p.pitch, o);
                    Bw[(ph + 2) % 3] = widen(o);
                    // ---- residual row a-2
                    if (!GUARD || k >= 4) {
                        const int yr = a - 2;
                        const float4 ku2 = stencil_rows(kw, Bw[(ph + 0) % 3], Bw[(ph + 1) % 3], Bw[(ph + 2) % 3]);
                        const float4 f2 = F[(ph + 0) % 3];
                        float4 r;
                        r.x = __fsub_rn(f2.x, ku2.x);
                        r.y = __fsub_rn(f2.y, ku2.y);
                        r.z = __fsub_rn(f2.z, ku2.z);
                        r.w = __fsub_rn(f2.w, ku2.w);
                        if (MODE == 1) {
                            if (p.want_norm && lane_int && (!GUARD || (yr >= y0 && yr < y1)) &&
                                (!EDGE || (yr >= 1 && yr <= N - 2))) {
                                // squares of the lane's 4 columns summed in fp32 (fixed order), rows in fp64: one
                                // F2F + DADD per row instead of eight FP64-pipe instructions
                                const float4 q = EDGE ? mask4(r, cin) : r;
                                float s4 = __fmul_rn(q.x, q.x);
                                s4 = __fmaf_rn(q.y, q.y, s4);
                                s4 = __fmaf_rn(q.z, q.z, s4);
                                s4 = __fmaf_rn(q.w, q.w, s4);
                                part += (double)s4;
                            }
                        } else {
                            R6 rr;
                            rr.a[0] = __shfl_up_sync(0xffffffffu, r.w, 1);
                            rr.a[1] = r.x;
                            rr.a[2] = r.y;
                            rr.a[3] = r.z;
                            rr.a[4] = r.w;
                            rr.a[5] = 0.0f;
                            C[(ph + 2) % 3] = rr;
                            // ---- coarse row (a-3)/2 when a-3 is even (k even since y0 is even)
                            if ((ph & 1) == 0 && (!GUARD || k >= 6)) {
                                const int yc = a - 3;
                                if (!GUARD || (yc >= y0 && yc < y1)) {
                                    const R6 &ct = C[(ph + 0) % 3], &cm = C[(ph + 1) % 3], &cb = C[(ph + 2) % 3];
                                    float o2[2];
#pragma unroll
                                    for (int h = 0; h < 2; ++h) {
                                        const int e = 2 * h;
                                        float sacc = __fmul_rn(rw[0], ct.a[e]);
                                        sacc = __fmaf_rn(rw[1], ct.a[e + 1], sacc);
                                        sacc = __fmaf_rn(rw[2], ct.a[e + 2], sacc);
                                        sacc = __fmaf_rn(rw[3], cm.a[e], sacc);
                                        sacc = __fmaf_rn(rw[4], cm.a[e + 1], sacc);
                                        sacc = __fmaf_rn(rw[5], cm.a[e + 2], sacc);
                                        sacc = __fmaf_rn(rw[6], cb.a[e], sacc);
                                        sacc = __fmaf_rn(rw[7], cb.a[e + 1], sacc);
                                        sacc = __fmaf_rn(rw[8], cb.a[e + 2], sacc);
                                        o2[h] = p.r_has_scale ? __fmul_rn(rscale, sacc) : sacc;
                                    }
                                    const int I = yc >> 1;
                                    if (EDGE) {
                                        const bool Iin = (I >= 1 && I <= p.Nc - 2);
                                        o2[0] = (Iin && cxl >= 1 && cxl <= p.Nc - 2) ? o2[0] : 0.0f;
                                        o2[1] = (Iin && cxl + 1 >= 1 && cxl + 1 <= p.Nc - 2) ? o2[1] : 0.0f;
                                    }
                                    if (fc_ok)
                                        *reinterpret_cast<float2 *>(fco + (long long)(I - p.crow0) * p.pitch_c) =
                                            make_float2(o2[0], o2[1]);
                                }
                            }
                        }
                    }
                }
            }
        };
        using T_ = std::true_type;
        using F_ = std::false_type;
        // steady state: k0 >= 6 (pipeline full, all store rows >= y0) and k0 + 5 <= K - 4 (all store rows < y1)
        int k0 = 0;
        if (edge) {
            for (; k0 < K; k0 += 6) block6(T_{}, T_{}, T_{}, k0);
        } else {
            block6(T_{}, F_{}, T_{}, 0);
            // steady state: no pipeline / store-range guards; prefetches unchecked while every prefetched row exists
            for (k0 = 6; k0 + 5 <= K - 4 && k0 + 5 + ST_DEPTH - 1 < khi; k0 += 6) block6(F_{}, F_{}, F_{}, k0);
            for (; k0 + 5 <= K - 4; k0 += 6) block6(F_{}, F_{}, T_{}, k0);
            for (; k0 < K; k0 += 6) block6(T_{}, F_{}, T_{}, k0);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (MODE == 1 && p.want_norm) {
            part = warp_sum_d(part);
            if (lane == 0) p.partials[s] = part;
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
                    for (int w = 0; w < ST_WARPS; ++w) sum += red[w];
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


// ---- source-key indexed stencil (two-phase meshes): slow path for the rows a material interface crosses
struct K6 {
    int k[6];  // pattern keys of box columns 4l-1 .. 4l+4
};
__device__ __forceinline__ K6 key6(unsigned int w) {
    const unsigned int l = __shfl_up_sync(0xffffffffu, w, 1), r = __shfl_down_sync(0xffffffffu, w, 1);
    K6 o;
    o.k[0] = (int)(l >> 24);
    o.k[1] = (int)(w & 0xffu);
    o.k[2] = (int)((w >> 8) & 0xffu);
    o.k[3] = (int)((w >> 16) & 0xffu);
    o.k[4] = (int)(w >> 24);
    o.k[5] = (int)(r & 0xffu);
    return o;
}
__device__ __forceinline__ void unpack_row(const RP &r, float (&a)[6]) {
    unpack2(r.q[0], a[0], a[1]);
    unpack2(r.q[2], a[2], a[3]);
    unpack2(r.q[4], a[4], a[5]);
}
// same row-major FMA chain as stencil_rows2, weights tab[key(source)][tap]; out = {K lo, K hi, inv lo, inv hi}.
// Deliberately NOT inlined: only the few windows a material interface crosses come here, and keeping this register-hungry
// code out of line leaves the single-pattern fast path's register allocation untouched.
__device__ __noinline__ void stencil_rows_keys(const float *tab, const float *invt, unsigned int w0, unsigned int w1,
                                               unsigned int w2, RP t, RP m, RP b, u64 *out) {
    const K6 kt = key6(w0), km = key6(w1), kb = key6(w2);
    float ta[6], ma[6], ba[6], acc[4];
    unpack_row(t, ta);
    unpack_row(m, ma);
    unpack_row(b, ba);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float s = __fmul_rn(tab[9 * kt.k[e] + 0], ta[e]);
        s = __fmaf_rn(tab[9 * kt.k[e + 1] + 1], ta[e + 1], s);
        s = __fmaf_rn(tab[9 * kt.k[e + 2] + 2], ta[e + 2], s);
        s = __fmaf_rn(tab[9 * km.k[e] + 3], ma[e], s);
        s = __fmaf_rn(tab[9 * km.k[e + 1] + 4], ma[e + 1], s);
        s = __fmaf_rn(tab[9 * km.k[e + 2] + 5], ma[e + 2], s);
        s = __fmaf_rn(tab[9 * kb.k[e] + 6], ba[e], s);
        s = __fmaf_rn(tab[9 * kb.k[e + 1] + 7], ba[e + 1], s);
        s = __fmaf_rn(tab[9 * kb.k[e + 2] + 8], ba[e + 2], s);
        acc[e] = s;
    }
    out[0] = pack2(acc[0], acc[1]);
    out[1] = pack2(acc[2], acc[3]);
    out[2] = pack2(invt[km.k[1]], invt[km.k[2]]);
    out[3] = pack2(invt[km.k[3]], invt[km.k[4]]);
}

// Packed variant: the two stencil chains per row run on fma.rn.f32x2 (FFMA2): same IEEE result per lane, half the
// issue slots.  Rows are kept as 5 overlapping column PAIRS q[i] = (a[i], a[i+1]).
// MODE 0: down leg (Jacobi sweep, store u, residual, restriction -> fc)
// MODE 1: up leg   (bilinear prolongation + correction, Jacobi sweep, store u, optional interior residual norm)
// KEYS: two-phase meshes.  The key words (4 keys per lane) of the NEXT block of 6 rows ride the prefetch group of the
// current block's first row into an 18-row ring.  At every block start one vote decides whether the block (and the one
// before it, for the stencils' look-back) carries ONE pattern over the whole strip: then the block runs the very same
// single-pattern code as an iso mesh, with that pattern's weights (reloaded from shared memory when the key changes,
// i.e. twice per strip that crosses the inclusion).  Blocks that touch the interface run a general variant whose
// stencils look every weight up by the source node's key (out of line, so the fast path's code and registers are
// those of the iso kernel).
// ONEV: a strip runs ONE block body from its first row to its last -- the guarded one (pipeline-fill guards, store-range
// predicates, checked prefetches), with the domain masks for the strips on the domain edge and without them for the
// others -- instead of switching between up to four per-phase instantiations (first block, steady state, checked
// prefetch, drain).  More instructions per row, but one loop body per warp in the instruction caches and a kernel image
// of 30-50 KB instead of 107-143 KB: the 4097^2 cycle went 0.1837 -> 0.1759 ms with the masked body everywhere, -> 0.1677
// with the redundant masks removed and the unmasked body for interior strips; most of it on the short coarse launches,
// which start with cold instruction caches (ncu no_instruction 0.5 - 1.05 stalled warps per issue at 2049^2 / 1025^2,
// 0.15 - 0.34 at 4097^2).  Found on mg_hstream_kernel, where three variants cost 18-21 %.
template <int MODE, bool ZERO_INIT, bool KEYS = false, bool PUSH = false, bool ONEV = false>
__global__ void __launch_bounds__(ST_WARPS * 32, 2) mg_stream2_kernel(const StreamParams p) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    __shared__ double red[ST_WARPS];
    __shared__ int lastflag;
    __shared__ float s_tab[KEYS ? MAXPAT * 9 : 1];   // stiffness tables of all patterns
    __shared__ float s_inv[KEYS ? MAXPAT : 1];       // omega / d per pattern
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = p.N;

    // ---- weights in registers for the whole kernel
    float kw[9], rw[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        kw[q] = p.ktab[q];
        rw[q] = (MODE == 0) ? p.rtab[q] : 0.0f;
    }
    const float inv = p.invd[0];
    u64 kw2[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) kw2[q] = pack2(kw[q], kw[q]);
    u64 inv2 = pack2(inv, inv);
    const u64 one2 = pack2(p.one, p.one);  // 1.0f from the launch parameters: opaque to ptxas (see the Jacobi update)
    const float rscale = (MODE == 0 && p.r_has_scale) ? (p.r_scale_dev ? *p.r_scale_dev : p.r_scale) : 1.0f;
    if (KEYS) {
        for (int i = threadIdx.x; i < MAXPAT * 9; i += ST_WARPS * 32) s_tab[i] = (i < p.npat * 9) ? p.ktab[i] : 0.0f;
        if (threadIdx.x < MAXPAT) s_inv[threadIdx.x] = (threadIdx.x < p.npat) ? p.invd[threadIdx.x] : 0.0f;
        __syncthreads();
    }
    int kcur = 0;  // pattern whose weights sit in kw2 / inv2
    constexpr bool zero = ZERO_INIT;  // (a run-time zero-guess case sharing level 0's kernel image measured slower: 0.180 vs 0.176 ms)
    pdl_wait();  // weights above are never written by a kernel; all field data is touched only after this point
    // the solve-control word is requested here but tested only after the first prefetches are in flight, so the two
    // memory round trips overlap (a finished solve returns before anything is stored)
    const int solve_done = (p.ctl != nullptr) ? ld_volatile_s32(&reinterpret_cast<const Ctl *>(p.ctl)->done) : 0;

    // ---- per-lane prefetch ring: [slot][lane] float4 for u and for f, float2 for the coarse row (up leg)
    constexpr int RD = st2_ring_rows(MODE, KEYS);  // ring depth (rows): 6, or two halves of 6
    constexpr int PD = RD - 3;                     // prefetch distance: rows k-2 .. k stay in the ring (f is re-read)
    constexpr int RF4 = st2_ring_f4(MODE, KEYS);
    float4 *ring_u = reinterpret_cast<float4 *>(st_smem) + warp * (RF4 * 32);
    float4 *ring_f = ring_u + RD * 32;
    float2 *ring_c = reinterpret_cast<float2 *>(ring_f + RD * 32);
    // key words: 18 rows deep (the residual stage looks back further than the 6-row data ring keeps its rows)
    unsigned int *ring_k = reinterpret_cast<unsigned int *>(st_smem + ST_WARPS * RF4 * 32 * 16) + warp * (ST_KDEPTH * 32);

    const int total = p.nstrips * p.B;
    for (int s = blockIdx.x * ST_WARPS + warp; s < total; s += gridDim.x * ST_WARPS) {
        // strip coordinates
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
        const int y0 = p.own0 + ry * p.R;                       // global rows
        const int y1 = (ry == p.nry - 1) ? p.own1 : y0 + p.R;  // exclusive
        const int gx = tx * ST_TWI - 4 + 4 * lane;         // first global column of this lane
        const bool lane_int = (lane >= 1 && lane <= 30);
        // column masks of this lane's 4 columns
        unsigned int cin = 0, cdom = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (gx + e >= 1 && gx + e <= N - 2) cin |= 1u << e;
            if (gx + e >= 0 && gx + e <= N - 1) cdom |= 1u << e;
        }
        const bool col_ok = (gx >= 0) && (gx + 3 < p.pitch);  // the 16-byte chunk exists in memory
        const float *ub = zero ? nullptr : p.u_in + (long long)b * p.plane + gx;
        const float *fb = p.f + (long long)b * p.plane + gx;
        float *uo = p.u_out + (long long)b * p.plane + gx;
        const int cxl = gx >> 1;  // coarse column of this lane's first fine column (gx is even)
        const bool ccol_ok = (MODE == 1) && (cxl >= 0) && (cxl + 1 < p.pitch_c);
        const float *cbp = (MODE == 1) ? p.vc + (long long)b * p.plane_c + cxl : nullptr;
        float *fco = (MODE == 0) ? p.fc + (long long)b * p.plane_c + cxl : nullptr;
        const bool fc_ok = lane_int && cxl >= 0 && cxl <= p.Nc - 1;
        // a strip whose whole streamed box lies strictly inside the domain needs no masks at all
        const bool edge = (y0 - 3 <= 0) || (y1 + 2 >= N - 1) || (tx == 0) || ((tx + 1) * ST_TWI + 4 >= N - 1);

        const int a0 = y0 - 3;  // first streamed row
        // last streamed row: y1+1 (u1 row y1 for the residual row y1-1); the last strip of the down leg goes one
        // further so that the coarse ring row (N-1)/2 is produced (as zeros) too
        const int K = ((MODE == 0 && y1 == N) ? y1 + 2 : y1 + 1) - a0 + 1;
        // prefetch of streamed row k (rows are requested in increasing k): running global pointers, compile-time ring
        // slot, and no validity test in steady-state blocks (CHECK = false)
        const int klo = max(0, p.row0) - a0;                    // first k whose row exists locally / in the domain
        const int khi = min(min(N, p.row0 + p.nrloc) - a0, K);  // one past the last such k
        const float *pf_u = zero ? nullptr : ub + (long long)(a0 - p.row0) * p.pitch;
        const float *pf_f = fb + (long long)(a0 - p.row0) * p.pitch;
        int kpf = 0;
        // KEYS: keys of the 6 streamed rows keyrow0 .. keyrow0+5 join the current commit group
        auto fetch_keys = [&](int keyrow0) {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int kr = keyrow0 + j, gy = a0 + kr;
                const bool okk = (gx >= 0 && gx + 3 < p.key_pitch && gy >= 0 && gy < N);
                const void *src = okk ? (const void *)(p.keys + (long long)gy * p.key_pitch + gx) : (const void *)p.f;
                const uint32_t sz = okk ? 4u : 0u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(&ring_k[(kr % ST_KDEPTH) * 32 + lane])),
                             "l"(src), "r"(sz)
                             : "memory");
            }
        };
        auto prefetch = [&](auto check_tag, int slot_row) {
            constexpr bool CHECK = decltype(check_tag)::value;
            const bool ok = !CHECK || (col_ok && kpf >= klo && kpf < khi);
            const int slot = slot_row * 32 + lane;
            // a copy of ZERO source bytes reads nothing (the destination is zero-filled), so the running pointers are
            // passed as they are for rows / columns outside the arrays: no pointer select per row
            if (!zero) st_cp16(&ring_u[slot], pf_u, ok);
            st_cp16(&ring_f[slot], pf_f, ok);
            if (MODE == 1) {  // coarse row ceil(a/2): the row an even fine row copies / an odd row's lower partner
                const int I = (a0 + kpf + 1) >> 1;
                const int lI = I - p.crow0;
                const bool okc = !CHECK || (ccol_ok && (I >= 0) && (I < p.Nc) && (lI >= 0) && (lI < p.nrc) && kpf < K);
                const void *src = (const void *)(cbp + (long long)lI * p.pitch_c);
                const uint32_t sz = okc ? 8u : 0u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(&ring_c[slot])), "l"(src),
                             "r"(sz)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (!zero) pf_u += p.pitch;
            pf_f += p.pitch;
            ++kpf;
        };
#pragma unroll
        if (KEYS) fetch_keys(0);  // block 0's keys travel with the first row
        for (int k = 0; k < PD; ++k) prefetch(std::true_type{}, k);  // at k0 = 0 the ring row of streamed row k is k
        if (solve_done) break;

        // rotating windows (indices are compile-time after unrolling by 6)
        RP A[3];   // input rows (u0, or corrected u for the up leg): a-2, a-1, a
        RP Bw[3];  // smoothed rows: a-3, a-2, a-1
        float racc0 = 0.f, racc1 = 0.f;  // restriction chains of the lane's two coarse columns (down leg), fed row by row
        // running store pointers: row a0 - 1 of u_out / coarse row y0/2 - 3 of fc at step 0, advanced as the rows go by
        float *st_u = uo + (long long)(a0 - 1 - p.row0) * p.pitch;
        // fused halo push: does this strip produce boundary rows a neighbour needs?
        const bool s_up = PUSH && p.push_up != nullptr && y0 < p.pown0 + p.push_rows && y1 > p.pown0;
        const bool s_dn = PUSH && p.push_dn != nullptr && y1 > p.pown1 - p.push_rows && y0 < p.pown1;
        const bool s_push = s_up || s_dn;
        float *st_c = (MODE == 0) ? fco + (long long)((y0 >> 1) - 3 - p.crow0) * p.pitch_c : nullptr;
        float vt[3] = {0.f, 0.f, 0.f};  // coarse row floor(a/2) at coarse columns cxl, cxl+1, cxl+2 (up leg)
        double part = 0.0;
        if (MODE == 1 && (a0 & 1) && a0 >= 1) {  // the strip starts on an odd row: fetch its upper coarse row directly
            const int lI = (a0 >> 1) - p.crow0;
            const float *row = p.vc + (long long)b * p.plane_c + (long long)lI * p.pitch_c;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int c = cxl + q;
                vt[q] = (c >= 0 && c < p.Nc && lI >= 0 && lI < p.nrc) ? __ldg(row + c) : 0.0f;
            }
        }

        // one block of 6 row steps.  GUARD: start-up / drain steps (pipeline fill, store ranges); EDGE: masks needed.
        // KEYED (two-phase meshes only): the block touches a material interface -> per-node weight lookup
        auto block6 = [&](auto guard_tag, auto edge_tag, auto pf_tag, auto keyed_tag, int k0) {
            constexpr bool GUARD = decltype(guard_tag)::value;
            constexpr bool EDGE = decltype(edge_tag)::value;
            constexpr bool KEYED = decltype(keyed_tag)::value;
            // ring row of streamed row k0 + d (d compile-time): a 12-row ring is two halves of 6, `hc` = the half that
            // holds rows k0 .. k0+5 (k0 is a multiple of 6), so every slot is a run-time half base + a constant
            const int hc = (RD == 12) ? ((k0 / 6) & 1) * 6 : 0, ho = (RD == 12) ? 6 - hc : 0;
            auto ring_row = [&](int d) {
                if (RD == 12) return d < 0 ? ho + 6 + d : (d < 6 ? hc + d : (d < 12 ? ho + d - 6 : hc + d - 12));
                return (d + 12) % 6;
            };
#pragma unroll
            for (int ph = 0; ph < 6; ++ph) {
                const int k = k0 + ph;
                if (GUARD && k >= K) break;
                const int a = a0 + k;
                if (KEYS && ph == 0) fetch_keys(k0 + 6);  // next block's keys, same commit group as row k0+3
                prefetch(pf_tag, ring_row(ph + PD));  // rows k-2..k stay in the ring (f is re-read)
                asm volatile("cp.async.wait_group %0;" ::"n"(PD) : "memory");
                const int slot = ring_row(ph) * 32 + lane;
                float4 uv = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : ring_u[slot];
                const bool arow_in = !EDGE || (a >= 1 && a <= N - 2);
                if (MODE == 1) {
                    // u += reset(bilinear P v_c) on row a.  fl(a/2 + b/2) is written fma(0.5, a, 0.5 b): halving is exact, so
                    // this is the same single rounding of the exact sum as the reference's mul, mul, add -- one FMA-pipe
                    // instruction less per value (the up leg is FMA-pipe bound)
                    const float2 cv = ring_c[slot];
                    const float c2 = __shfl_down_sync(0xffffffffu, cv.x, 1);
                    {  // rows outside the domain too: their coarse rows are zero-filled and u + e is reset to zero below
                        float4 e;
                        if ((ph & 1) != 0) {  // k odd <=> a even (a0 is odd): copy / horizontal average of coarse row a/2
                            vt[0] = cv.x;
                            vt[1] = cv.y;
                            vt[2] = c2;
                            e.x = vt[0];
                            e.y = __fmaf_rn(0.5f, vt[0], __fmul_rn(0.5f, vt[1]));
                            e.z = vt[1];
                            e.w = __fmaf_rn(0.5f, vt[1], __fmul_rn(0.5f, vt[2]));
                        } else {  // a odd: between coarse rows (a-1)/2 (vt, kept from the previous step) and (a+1)/2
                            const float vb0 = cv.x, vb1 = cv.y, vb2 = c2;
                            e.x = __fmaf_rn(0.5f, vt[0], __fmul_rn(0.5f, vb0));
                            e.z = __fmaf_rn(0.5f, vt[1], __fmul_rn(0.5f, vb1));
                            if (p.prolong_seq) {
                                float v = __fadd_rn(__fmul_rn(0.25f, vt[0]), __fmul_rn(0.25f, vt[1]));
                                v = __fadd_rn(v, __fmul_rn(0.25f, vb0));
                                e.y = __fadd_rn(v, __fmul_rn(0.25f, vb1));
                                v = __fadd_rn(__fmul_rn(0.25f, vt[1]), __fmul_rn(0.25f, vt[2]));
                                v = __fadd_rn(v, __fmul_rn(0.25f, vb1));
                                e.w = __fadd_rn(v, __fmul_rn(0.25f, vb2));
                            } else {
                                const float ta = __fmaf_rn(0.5f, vt[0], __fmul_rn(0.5f, vt[1]));
                                const float ba = __fmaf_rn(0.5f, vb0, __fmul_rn(0.5f, vb1));
                                e.y = __fmaf_rn(0.5f, ta, __fmul_rn(0.5f, ba));
                                const float tb = __fmaf_rn(0.5f, vt[1], __fmul_rn(0.5f, vt[2]));
                                const float bb = __fmaf_rn(0.5f, vb1, __fmul_rn(0.5f, vb2));
                                e.w = __fmaf_rn(0.5f, tb, __fmul_rn(0.5f, bb));
                            }
                        }
                        // (the fine level's reset_boundary of the correction, and the zeroing of the columns beyond the
                        // domain, are both subsumed by the sweep's own reset_boundary of u + e just below: interior nodes
                        // see the unmasked e, every other node is set to zero there)
                        uv.x = __fadd_rn(uv.x, e.x);
                        uv.y = __fadd_rn(uv.y, e.y);
                        uv.z = __fadd_rn(uv.z, e.z);
                        uv.w = __fadd_rn(uv.w, e.w);
                    }
                }
                // reset_boundary of the sweep's input (ring -> 0, outside stays 0)
                if (EDGE) uv = mask4(uv, arow_in ? cin : 0u);
                A[(ph + 2) % 3] = widen2(uv.x, uv.y, uv.z, uv.w);
                // ---- Jacobi row a-1
                if (!GUARD || k >= 2) {
                    const int y = a - 1;
                    const RP &t = A[(ph + 0) % 3], &m = A[(ph + 1) % 3], &bq = A[(ph + 2) % 3];
                    u64 klo, khi, ivlo = inv2, ivhi = inv2;
                    if (KEYED) {  // rows k-2, k-1, k
                        u64 out[4];
                        stencil_rows_keys(s_tab, s_inv, ring_k[((k - 2 + ST_KDEPTH) % ST_KDEPTH) * 32 + lane],
                                          ring_k[((k - 1 + ST_KDEPTH) % ST_KDEPTH) * 32 + lane],
                                          ring_k[(k % ST_KDEPTH) * 32 + lane], t, m, bq, out);
                        klo = out[0];
                        khi = out[1];
                        ivlo = out[2];
                        ivhi = out[3];
                    } else {
                        stencil_rows2(kw2, t, m, bq, klo, khi);
                    }
                    const ulonglong2 ff = *reinterpret_cast<const ulonglong2 *>(&ring_f[ring_row(ph - 1) * 32 + lane]);
                    // u + inv * (f - K u) with TWO roundings (reference: separate mul and add).  ptxas contracts
                    // mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (measured: the result then carries the fused rounding,
                    // visible as soon as omega/d is not a power of two), so the sum is written as fma(p, 1, u) = fl(p + u):
                    // a product feeding an fma's MULTIPLICAND cannot be contracted (and the 1 is a launch parameter, or
                    // ptxas folds the fma back into an add and contracts again)
                    const u64 plo = mul2(ivlo, sub2(ff.x, klo)), phi = mul2(ivhi, sub2(ff.y, khi));
                    const u64 olo = fma2(plo, one2, m.q[1]), ohi = fma2(phi, one2, m.q[3]);
                    float o0, o1, o2, o3;
                    unpack2(olo, o0, o1);
                    unpack2(ohi, o2, o3);
                    if (EDGE) {
                        const float4 om = mask4(make_float4(o0, o1, o2, o3), (y >= 1 && y <= N - 2) ? cin : 0u);
                        o0 = om.x;
                        o1 = om.y;
                        o2 = om.z;
                        o3 = om.w;
                    }
                    const bool st_ok = lane_int && (!GUARD || (y >= y0 && y < y1)) && (!EDGE || (cdom & 1u));
                    st_global_v4_pred(st_u, o0, o1, o2, o3, st_ok);  // predicated: no branch around the store
                    if (PUSH && s_push) {  // warp-uniform: only the strips that hold boundary rows come here
                        // the same row into the neighbour's ghost rows (peer stores over NVLink)
                        float *pp = p.push_up + (long long)y * p.pitch + gx;
                        st_global_v4_pred(pp, o0, o1, o2, o3, st_ok && s_up && (unsigned)(y - p.pown0) < (unsigned)p.push_rows);
                        pp = p.push_dn + (long long)y * p.pitch + gx;
                        st_global_v4_pred(pp, o0, o1, o2, o3,
                                          st_ok && s_dn && (unsigned)(p.pown1 - 1 - y) < (unsigned)p.push_rows);
                    }
                    Bw[(ph + 2) % 3] = widen2(o0, o1, o2, o3);
                    // ---- residual row a-2
                    if (!GUARD || k >= 4) {
                        const int yr = a - 2;
                        u64 rlo, rhi;
                        if (KEYED) {  // rows k-3, k-2, k-1
                            u64 out[4];
                            stencil_rows_keys(s_tab, s_inv, ring_k[((k - 3 + ST_KDEPTH) % ST_KDEPTH) * 32 + lane],
                                              ring_k[((k - 2 + ST_KDEPTH) % ST_KDEPTH) * 32 + lane],
                                              ring_k[((k - 1 + ST_KDEPTH) % ST_KDEPTH) * 32 + lane], Bw[(ph + 0) % 3],
                                              Bw[(ph + 1) % 3], Bw[(ph + 2) % 3], out);
                            rlo = out[0];
                            rhi = out[1];
                        } else {
                            stencil_rows2(kw2, Bw[(ph + 0) % 3], Bw[(ph + 1) % 3], Bw[(ph + 2) % 3], rlo, rhi);
                        }
                        const ulonglong2 f2 = *reinterpret_cast<const ulonglong2 *>(&ring_f[ring_row(ph - 2) * 32 + lane]);
                        rlo = sub2(f2.x, rlo);
                        rhi = sub2(f2.y, rhi);
                        float4 r;
                        unpack2(rlo, r.x, r.y);
                        unpack2(rhi, r.z, r.w);
                        if (MODE == 1) {
                            if (p.want_norm) {
                                // squares of the lane's 4 columns summed in fp32 (fixed order), rows in fp64: one
                                // F2F + DADD per row instead of eight FP64-pipe instructions; the row / lane conditions
                                // select the addend instead of branching around it
                                const bool nok = lane_int && (!GUARD || (yr >= y0 && yr < y1)) &&
                                                 (!EDGE || (yr >= 1 && yr <= N - 2));
                                const float4 q = EDGE ? mask4(r, cin) : r;
                                float s4 = __fmul_rn(q.x, q.x);
                                s4 = __fmaf_rn(q.y, q.y, s4);
                                s4 = __fmaf_rn(q.z, q.z, s4);
                                s4 = __fmaf_rn(q.w, q.w, s4);
                                part += nok ? (double)s4 : 0.0;
                            }
                        } else {
                            // restriction, fed row by row in the chain's own order (row-major taps): no window of residual
                            // rows.  Residual row a-2 is the bottom row of the coarse stencil centred on fine row a-3 and
                            // the top row of the next one when it is odd (k even, y0 is even), the middle row otherwise.
                            const float rl = __shfl_up_sync(0xffffffffu, r.w, 1);  // column 4l-1
                            if ((ph & 1) == 0) {
                                if (!GUARD || k >= 6) {
                                    const int yc = a - 3;
                                    float c0 = __fmaf_rn(rw[6], rl, racc0), c1 = __fmaf_rn(rw[6], r.y, racc1);
                                    c0 = __fmaf_rn(rw[7], r.x, c0);
                                    c1 = __fmaf_rn(rw[7], r.z, c1);
                                    c0 = __fmaf_rn(rw[8], r.y, c0);
                                    c1 = __fmaf_rn(rw[8], r.w, c1);
                                    if (p.r_has_scale) {
                                        c0 = __fmul_rn(rscale, c0);
                                        c1 = __fmul_rn(rscale, c1);
                                    }
                                    if (EDGE) {
                                        const int I = yc >> 1;
                                        const bool Iin = (I >= 1 && I <= p.Nc - 2);
                                        c0 = (Iin && cxl >= 1 && cxl <= p.Nc - 2) ? c0 : 0.0f;
                                        c1 = (Iin && cxl + 1 >= 1 && cxl + 1 <= p.Nc - 2) ? c1 : 0.0f;
                                    }
                                    st_global_v2_pred(st_c, c0, c1, fc_ok && (!GUARD || (yc >= y0 && yc < y1)));
                                }
                                racc0 = __fmul_rn(rw[0], rl);
                                racc1 = __fmul_rn(rw[0], r.y);
                                racc0 = __fmaf_rn(rw[1], r.x, racc0);
                                racc1 = __fmaf_rn(rw[1], r.z, racc1);
                                racc0 = __fmaf_rn(rw[2], r.y, racc0);
                                racc1 = __fmaf_rn(rw[2], r.w, racc1);
                            } else {
                                racc0 = __fmaf_rn(rw[3], rl, racc0);
                                racc1 = __fmaf_rn(rw[3], r.y, racc1);
                                racc0 = __fmaf_rn(rw[4], r.x, racc0);
                                racc1 = __fmaf_rn(rw[4], r.z, racc1);
                                racc0 = __fmaf_rn(rw[5], r.y, racc0);
                                racc1 = __fmaf_rn(rw[5], r.w, racc1);
                            }
                        }
                    }
                }
                st_u += p.pitch;
                if (MODE == 0 && (ph & 1) == 0) st_c += p.pitch_c;
            }
        };
        using T_ = std::true_type;
        using F_ = std::false_type;
        // steady state: k0 >= 6 (pipeline full, all store rows >= y0) and k0 + 5 <= K - 4 (all store rows < y1)
        int k0 = 0;
        if (!KEYS) {
            if (edge) {
                for (; k0 < K; k0 += 6) block6(T_{}, T_{}, T_{}, F_{}, k0);
            } else if (ONEV) {  // one body for the whole strip: guards and checked prefetches on, no masks
                for (; k0 < K; k0 += 6) block6(T_{}, F_{}, T_{}, F_{}, k0);
            } else {
                block6(T_{}, F_{}, T_{}, F_{}, 0);
                // steady state: no pipeline / store-range guards; prefetches unchecked while every prefetched row exists
                for (k0 = 6; k0 + 5 <= K - 4 && k0 + 5 + PD < khi; k0 += 6) block6(F_{}, F_{}, F_{}, F_{}, k0);
                for (; k0 + 5 <= K - 4; k0 += 6) block6(F_{}, F_{}, T_{}, F_{}, k0);
                for (; k0 < K; k0 += 6) block6(T_{}, F_{}, T_{}, F_{}, k0);
            }
        } else {
            int prev_key = -2;  // uniform key of the previous block (-1: mixed, -2: no previous block)
            asm volatile("cp.async.wait_group %0;" ::"n"(PD - 1) : "memory");  // block 0's keys (first group) are in
            for (; k0 < K; k0 += 6) {
                // one vote per block: do the 6 rows carry ONE pattern over the whole strip?
                const unsigned int w0 = ring_k[(k0 % ST_KDEPTH) * 32 + lane];
                const unsigned int pat = (__shfl_sync(0xffffffffu, w0, 0) & 0xffu) * 0x01010101u;
                unsigned int diff = w0 ^ pat;
#pragma unroll
                for (int j = 1; j < 6; ++j) diff |= ring_k[((k0 + j) % ST_KDEPTH) * 32 + lane] ^ pat;
                const int bkey = __all_sync(0xffffffffu, diff == 0u) ? (int)(pat & 0xffu) : -1;
                const bool fast = (bkey >= 0) && (prev_key == -2 || prev_key == bkey);
                prev_key = bkey;
                if (!fast) {
                    block6(T_{}, T_{}, T_{}, T_{}, k0);
                    continue;
                }
                if (bkey != kcur) {  // entering / leaving the inclusion: this pattern's weights into the registers
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const float w = s_tab[9 * bkey + q];
                        kw2[q] = pack2(w, w);
                    }
                    const float iv = s_inv[bkey];
                    inv2 = pack2(iv, iv);
                    kcur = bkey;
                }
                if (edge) block6(T_{}, T_{}, T_{}, F_{}, k0);
                else if (ONEV || k0 == 0) block6(T_{}, F_{}, T_{}, F_{}, k0);
                else if (k0 + 5 <= K - 4 && k0 + 5 + PD < khi) block6(F_{}, F_{}, F_{}, F_{}, k0);
                else if (k0 + 5 <= K - 4) block6(F_{}, F_{}, T_{}, F_{}, k0);
                else block6(T_{}, F_{}, T_{}, F_{}, k0);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (MODE == 1 && p.want_norm) {
            part = warp_sum_d(part);
            if (lane == 0) p.partials[s] = part;
        }
        if (PUSH && s_push) {
            // all lanes' peer stores precede lane 0's system-scope fence (warp barrier), the fence precedes the ticket;
            // the strip that takes the last ticket publishes the flag: one remote increment per launch and direction
            __syncwarp();
            if (lane == 0) {
                __threadfence_system();
                if (s_up && atomicAdd(p.push_ticket, 1u) == (unsigned)p.npush_up - 1u) {
                    p.push_ticket[0] = 0u;
                    __threadfence_system();
                    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p.push_flag_up) : "memory");
                }
                if (s_dn && atomicAdd(p.push_ticket + 1, 1u) == (unsigned)p.npush_dn - 1u) {
                    p.push_ticket[1] = 0u;
                    __threadfence_system();
                    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p.push_flag_dn) : "memory");
                }
            }
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
            Ctl *ctl = p.ctl_ro ? nullptr : reinterpret_cast<Ctl *>(p.ctl);
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
                    for (int w = 0; w < ST_WARPS; ++w) sum += red[w];
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
