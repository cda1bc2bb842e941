// Tile "programs" for one multigrid level on sm_100a.
//
// One persistent CTA processes a sequence of tiles.  For every tile it stages a BOX of the level's fields in shared
// memory (TMA cp.async.bulk.tensor with hardware zero fill outside the domain == the reference's zero padding;
// cp.async fallback), then runs a short program of stencil STAGES entirely in shared memory/registers:
//
//   [prolong+correct] -> [nsweeps x (Jacobi | HNet-Jacobi)] -> [store u tile] -> [residual -> {store | restrict | norm}]
//
// Each stage consumes one ring of halo ("depth"), so `nsweeps` smoothing sweeps and the residual are temporally
// blocked: the level's u and f are read from HBM once and u is written once per program.
//
// Geometry: the box is BH rows x BW(=128) columns; lane l of every warp owns box columns 4l..4l+3 (one LDS.128 per
// row, neighbours' edge values by warp shuffle), warps split the rows of a stage and slide a 3-row register window
// down their rows.  The tile interior is rows [HT, HT+TH) x columns [HX, BW-HX).
//
// Arithmetic order is the canonical one of oracle/mgfea_oracle.c (row-major FMA chain, explicit roundings).
#pragma once
#include "mgfea_ptx.cuh"

namespace mgfea {

constexpr int BW = 128;        // box width in floats == 32 lanes x 4
constexpr int NTHREADS = 256;  // 8 warps
constexpr int NWARPS = NTHREADS / 32;
constexpr int MAXWARPS = NWARPS;
constexpr int MAXPAT = 16;
constexpr int MAXLAYERS = 8;
// TMA tiled loads (no swizzle) fault with "illegal instruction" unless the innermost start coordinate is a multiple of
// 16 bytes (measured on B200, tools/tma_probe.cu).  The fp32 field boxes start at x0-HX (a multiple of 4 floats); the
// uint8 key box and the coarse boxes are therefore started at the 16-byte boundary below their first needed column
// and carry the offset (kofs / vcofs / kcofs).
constexpr int KBW = 144;  // key box width in bytes: up to 12 bytes of alignment slack + 128
constexpr int CW = 68;    // coarse box width in floats: up to 3 floats of slack + BW/2 + 1
constexpr int KCW = 80;   // coarse key box width in bytes: up to 15 bytes of slack + BW/2 + 1

enum OutMode { OUT_NONE = 0, OUT_RESIDUAL = 1, OUT_KU = 2, OUT_RESTRICT = 3, OUT_NORM = 4 };

struct TileMaps {
    CUtensorMap u, f, k, vc, kc, idx, bval;
};

struct TileParams {
    // level geometry
    int N, B, pitch;
    long long plane;
    int TH, HX, HT, HB, BH, TWI;  // tile interior rows, column halo, top/bottom halo rows, box rows, interior cols
    int ntx, nty, ntiles;         // tiles per sample in x / y, total tiles (all samples)
    float inv_per, inv_ntx;       // reciprocals for the division-free tile index decomposition
    int use_tma, nstages;
    // program
    int zero_init, prolong_mode, prolong_seq, nsweeps, smoother, store_u, out_mode;
    // level data
    const float *u_in;
    float *u_out;
    const float *f;
    const unsigned char *keys;
    int key_pitch, npat;
    const float *ktab, *invd;
    const float *bc_idx, *bc_val;
    long long bc_plane;
    const float *hw;
    int nlayers;
    // coarse partner (prolongation source / restriction target)
    int Nc, pitch_c;
    long long plane_c;
    const float *vc;
    const unsigned char *keys_c;
    int key_pitch_c;
    const float *ptab;
    int ptab_n, p_has_scale;
    float p_scale;
    const float *p_scale_dev;
    float *fc;
    const float *rtab;
    int rtab_n, r_has_scale;
    float r_scale;
    const float *r_scale_dev;
    // residual / Ku output
    float *r_out;
    // norm
    double *tile_partials;
    unsigned int *counter;
    double *sumsq;
    void *ctl;  // mgfea_ctl*
    double *hist;
    // smem carve-up (bytes)
    int off_stage0, stage_bytes, off_w1, off_w2, off_w3;
    int so_u, so_f, so_k, so_vc, so_kc, so_idx, so_bval;  // offsets inside a stage
    int CH;                                                // coarse box rows
    unsigned int tx_bytes;                                 // bytes per stage delivered by TMA
};

struct Ctl {  // mirror of mgfea_ctl
    int cycle, done, min_cycles, max_cycles, conv_rule, pad_;
    double eps2;
};

// ---------------------------------------------------------------------------------------------------------
// shared-memory tables (first 4 KB of dynamic smem)
struct Tables {
    float ktab[MAXPAT * 9];
    float invd[MAXPAT];
    float rtab[MAXPAT * 9];
    float ptab[MAXPAT * 9];
    float hw[MAXLAYERS * 9];
    float r_scale, p_scale;
    unsigned long long mbar[2];
    double red[MAXWARPS];
    int flag;
};
constexpr int TABLES_BYTES = 4096;
static_assert(sizeof(Tables) <= TABLES_BYTES, "tables region too small");

struct Row6 {
    float a[6];  // box columns 4l-1 .. 4l+4
};

// One LDS.128 per lane + the two edge values from the neighbouring lanes.  Lane 0 / 31 get their OWN value for the
// missing neighbour (box column -1 / 128 does not exist); that only perturbs box columns 0 / 127, which are more than
// `depth` columns away from the tile interior, so it never reaches a stored value.
__device__ __forceinline__ Row6 load_row6(const float *rowp) {
    Row6 o;
    const float4 v = *reinterpret_cast<const float4 *>(rowp);
    o.a[0] = __shfl_up_sync(0xffffffffu, v.w, 1);
    o.a[1] = v.x;
    o.a[2] = v.y;
    o.a[3] = v.z;
    o.a[4] = v.w;
    o.a[5] = __shfl_down_sync(0xffffffffu, v.x, 1);
    return o;
}

struct Key6 {
    int k[6];
};
__device__ __forceinline__ Key6 load_key6(const unsigned char *kbuf, int row, int lane) {
    Key6 o;
    const unsigned int w = *reinterpret_cast<const unsigned int *>(kbuf + row * KBW + 4 * lane);
    unsigned int l = __shfl_up_sync(0xffffffffu, w, 1);
    unsigned int r = __shfl_down_sync(0xffffffffu, w, 1);
    o.k[0] = (lane == 0) ? 0 : (int)(l >> 24);
    o.k[1] = (int)(w & 0xffu);
    o.k[2] = (int)((w >> 8) & 0xffu);
    o.k[3] = (int)((w >> 16) & 0xffu);
    o.k[4] = (int)(w >> 24);
    o.k[5] = (lane == 31) ? 0 : (int)(r & 0xffu);
    return o;
}

// 3x3 stencil, one weight set: acc_e = chain_{t row-major} fma(w[t], src_t, acc)
__device__ __forceinline__ void stencil4(const float (&w)[9], const Row6 &t, const Row6 &m, const Row6 &b,
                                         float (&acc)[4]) {
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
}
// source-key indexed weights: w = tab[key(src)][t]
__device__ __forceinline__ void stencil4_keys(const float *tab, const Row6 &t, const Row6 &m, const Row6 &b,
                                              const Key6 &kt, const Key6 &km, const Key6 &kb, float (&acc)[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float s = __fmul_rn(tab[9 * kt.k[e] + 0], t.a[e]);
        s = __fmaf_rn(tab[9 * kt.k[e + 1] + 1], t.a[e + 1], s);
        s = __fmaf_rn(tab[9 * kt.k[e + 2] + 2], t.a[e + 2], s);
        s = __fmaf_rn(tab[9 * km.k[e] + 3], m.a[e], s);
        s = __fmaf_rn(tab[9 * km.k[e + 1] + 4], m.a[e + 1], s);
        s = __fmaf_rn(tab[9 * km.k[e + 2] + 5], m.a[e + 2], s);
        s = __fmaf_rn(tab[9 * kb.k[e] + 6], b.a[e], s);
        s = __fmaf_rn(tab[9 * kb.k[e + 1] + 7], b.a[e + 1], s);
        s = __fmaf_rn(tab[9 * kb.k[e + 2] + 8], b.a[e + 2], s);
        acc[e] = s;
    }
}

// per-tile state shared by the stage functions
struct TileCtx {
    float *U, *F, *W1, *W2, *W3, *VC, *IDX, *BVAL;
    unsigned char *K, *KC;
    int gy0, gx0;  // global row / column of box element (0,0)
    int cy0, cx0;  // global coarse row / column of the first coarse node the box needs
    int vcofs, kcofs;  // column of cx0 inside the (16-byte aligned) coarse value / coarse key boxes
    int b;         // sample
    int BH, N;
    bool keys_uniform;  // every key of the box equals k0
    int k0;
    bool touches_edge;  // box intersects the ring or the outside of the domain
};

// Weights that are uniform for a tile, held in registers for the lifetime of the kernel (single-pattern meshes) or
// refreshed per tile (two-phase meshes, uniform-key tiles).
struct RegW {
    float kw[9];  // stiffness stencil of pattern k0
    float inv;    // omega/d of pattern k0
    float rw[9];  // restriction kernel of pattern k0 (or the single kernel)
};

// split rows [lo, hi) of a stage among the warps: contiguous chunks
__device__ __forceinline__ void warp_rows(int lo, int hi, int warp, int &ra, int &rb) {
    const int n = hi - lo;
    const int per = (n + NWARPS - 1) / NWARPS;
    ra = lo + warp * per;
    rb = min(ra + per, hi);
    if (ra > hi) ra = hi;
}

// default square-ring mask of this lane's 4 columns: bit e set <=> column is interior (1..N-2)
__device__ __forceinline__ unsigned int col_interior_bits(const TileCtx &c, int lane) {
    unsigned int m = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int gx = c.gx0 + 4 * lane + e;
        if (gx >= 1 && gx <= c.N - 2) m |= 1u << e;
    }
    return m;
}
__device__ __forceinline__ unsigned int col_domain_bits(const TileCtx &c, int lane) {
    unsigned int m = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int gx = c.gx0 + 4 * lane + e;
        if (gx >= 0 && gx <= c.N - 1) m |= 1u << e;
    }
    return m;
}

// Row loop over a warp's rows [ra, rb): TWO output rows per iteration from a 4-row register window (rows r-1..r+2).
// The two `body` calls of an iteration are independent (8 FMA chains per lane in flight instead of 4) and both row
// loads are issued together, which hides the LDS.128 -> SHFL -> FFMA latency at the modest occupancy that the staged
// boxes allow.  Unrolled by two iterations so that the window rotation costs no register moves.
// `body(r, t, m, b)` is called once for every row r with the rows r-1, r, r+1 of `src`.
template <class Body>
__device__ __forceinline__ void for_rows3(const float *src, int ra, int rb, int lane, Body body) {
    const float *ps = src + (ra - 1) * BW + 4 * lane;
    Row6 a = load_row6(ps), b = load_row6(ps + BW), c, d;
    int r = ra;
    while (r + 1 < rb) {
        c = load_row6(ps + 2 * BW);
        d = load_row6(ps + 3 * BW);
        body(r, a, b, c);
        body(r + 1, b, c, d);
        r += 2;
        ps += 2 * BW;
        if (r + 1 >= rb) {
            a = c;
            b = d;
            break;
        }
        a = load_row6(ps + 2 * BW);
        b = load_row6(ps + 3 * BW);
        body(r, c, d, a);
        body(r + 1, d, a, b);
        r += 2;
        ps += 2 * BW;
    }
    if (r < rb) {
        c = load_row6(ps + 2 * BW);
        body(r, a, b, c);
    }
}

// ---------------------------------------------------------------------------------------------------------
// reset_boundary on box rows [lo,hi): dst = src*idx + bval   (default BC: ring -> 0, interior unchanged)
template <bool GBC>
__device__ __forceinline__ void stage_reset(const TileCtx &c, const float *src, float *dst, int lo, int hi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int cin = col_interior_bits(c, lane);
    for (int r = lo + warp; r < hi; r += NWARPS) {
        float4 v = *reinterpret_cast<const float4 *>(src + r * BW + 4 * lane);
        if (GBC) {
            const float4 g = *reinterpret_cast<const float4 *>(c.IDX + r * BW + 4 * lane);
            const float4 bv = *reinterpret_cast<const float4 *>(c.BVAL + r * BW + 4 * lane);
            v.x = __fadd_rn(__fmul_rn(v.x, g.x), bv.x);
            v.y = __fadd_rn(__fmul_rn(v.y, g.y), bv.y);
            v.z = __fadd_rn(__fmul_rn(v.z, g.z), bv.z);
            v.w = __fadd_rn(__fmul_rn(v.w, g.w), bv.w);
        } else {
            const int gy = c.gy0 + r;
            const bool rin = (gy >= 1 && gy <= c.N - 2);
            v.x = (rin && (cin & 1u)) ? v.x : 0.0f;
            v.y = (rin && (cin & 2u)) ? v.y : 0.0f;
            v.z = (rin && (cin & 4u)) ? v.z : 0.0f;
            v.w = (rin && (cin & 8u)) ? v.w : 0.0f;
        }
        *reinterpret_cast<float4 *>(dst + r * BW + 4 * lane) = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// One weighted-Jacobi update on box rows [lo,hi) (src must be reset and valid on [lo-1,hi+1)):
//   dst = BC( invd[key] * (f - K src) + src ).   If XOUT: also xdst = dst - raw (HNet input, raw = un-reset u).
// EDGE = the box touches the ring / outside of the domain (default-BC masks needed); interior tiles skip all masking.
template <bool KEYS, bool GBC, bool XOUT, bool EDGE>
__device__ __forceinline__ void stage_jacobi(const TileCtx &c, const Tables &T, const RegW &W, const float *src,
                                             float *dst, const float *raw, float *xdst, int lo, int hi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ra, rb;
    warp_rows(lo, hi, warp, ra, rb);
    if (ra >= rb) return;
    const bool slow = KEYS && !c.keys_uniform;
    if (!slow) {
        unsigned int cin = 0xfu;
        if (EDGE && !GBC) cin = col_interior_bits(c, lane);
        const int lofs = 4 * lane;
        for_rows3(src, ra, rb, lane, [&](int r, const Row6 &t, const Row6 &m, const Row6 &b) {
            float acc[4];
            stencil4(W.kw, t, m, b, acc);
            const int o4 = r * BW + lofs;
            const float4 fv = *reinterpret_cast<const float4 *>(c.F + o4);
            float o[4];
            o[0] = __fadd_rn(__fmul_rn(W.inv, __fsub_rn(fv.x, acc[0])), m.a[1]);
            o[1] = __fadd_rn(__fmul_rn(W.inv, __fsub_rn(fv.y, acc[1])), m.a[2]);
            o[2] = __fadd_rn(__fmul_rn(W.inv, __fsub_rn(fv.z, acc[2])), m.a[3]);
            o[3] = __fadd_rn(__fmul_rn(W.inv, __fsub_rn(fv.w, acc[3])), m.a[4]);
            if (GBC) {
                const float4 g = *reinterpret_cast<const float4 *>(c.IDX + o4);
                const float4 bv = *reinterpret_cast<const float4 *>(c.BVAL + o4);
                o[0] = __fadd_rn(__fmul_rn(o[0], g.x), bv.x);
                o[1] = __fadd_rn(__fmul_rn(o[1], g.y), bv.y);
                o[2] = __fadd_rn(__fmul_rn(o[2], g.z), bv.z);
                o[3] = __fadd_rn(__fmul_rn(o[3], g.w), bv.w);
            } else if (EDGE) {
                const int gy = c.gy0 + r;
                const unsigned int mk = (gy >= 1 && gy <= c.N - 2) ? cin : 0u;
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = ((mk >> e) & 1u) ? o[e] : 0.0f;
            }
            *reinterpret_cast<float4 *>(dst + o4) = make_float4(o[0], o[1], o[2], o[3]);
            if (XOUT) {
                const float4 rv = *reinterpret_cast<const float4 *>(raw + o4);
                *reinterpret_cast<float4 *>(xdst + o4) = make_float4(__fsub_rn(o[0], rv.x), __fsub_rn(o[1], rv.y),
                                                                     __fsub_rn(o[2], rv.z), __fsub_rn(o[3], rv.w));
            }
        });
        return;
    }
    // per-node pattern lookup (tiles crossed by the material interface)
    const unsigned int cin = col_interior_bits(c, lane);
    Row6 t = load_row6(src + (ra - 1) * BW + 4 * lane), m = load_row6(src + ra * BW + 4 * lane);
    Key6 kt = load_key6(c.K, ra - 1, lane), km = load_key6(c.K, ra, lane), kb;
    for (int r = ra; r < rb; ++r) {
        const Row6 b = load_row6(src + (r + 1) * BW + 4 * lane);
        float acc[4], inv[4];
        kb = load_key6(c.K, r + 1, lane);
        stencil4_keys(T.ktab, t, m, b, kt, km, kb, acc);
#pragma unroll
        for (int e = 0; e < 4; ++e) inv[e] = T.invd[km.k[e + 1]];
        kt = km;
        km = kb;
        const float4 fv = *reinterpret_cast<const float4 *>(c.F + r * BW + 4 * lane);
        const float ff[4] = {fv.x, fv.y, fv.z, fv.w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = __fadd_rn(__fmul_rn(inv[e], __fsub_rn(ff[e], acc[e])), m.a[e + 1]);
        if (GBC) {
            const float4 g = *reinterpret_cast<const float4 *>(c.IDX + r * BW + 4 * lane);
            const float4 bv = *reinterpret_cast<const float4 *>(c.BVAL + r * BW + 4 * lane);
            o[0] = __fadd_rn(__fmul_rn(o[0], g.x), bv.x);
            o[1] = __fadd_rn(__fmul_rn(o[1], g.y), bv.y);
            o[2] = __fadd_rn(__fmul_rn(o[2], g.z), bv.z);
            o[3] = __fadd_rn(__fmul_rn(o[3], g.w), bv.w);
        } else {
            const int gy = c.gy0 + r;
            const bool rin = (gy >= 1 && gy <= c.N - 2);
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = (rin && ((cin >> e) & 1u)) ? o[e] : 0.0f;
        }
        *reinterpret_cast<float4 *>(dst + r * BW + 4 * lane) = make_float4(o[0], o[1], o[2], o[3]);
        if (XOUT) {
            const float4 rv = *reinterpret_cast<const float4 *>(raw + r * BW + 4 * lane);
            *reinterpret_cast<float4 *>(xdst + r * BW + 4 * lane) =
                make_float4(__fsub_rn(o[0], rv.x), __fsub_rn(o[1], rv.y), __fsub_rn(o[2], rv.z), __fsub_rn(o[3], rv.w));
        }
        t = m;
        m = b;
    }
}

// ---------------------------------------------------------------------------------------------------------
// One HNet layer on rows [lo,hi): dst = geometry_idx * (w9 (*) src) ; if ADD: dst = base + that  (u = J + H(x))
template <bool GBC, bool ADD, bool EDGE>
__device__ __forceinline__ void stage_hlayer(const TileCtx &c, const float *w9, const float *src, float *dst,
                                             const float *base, int lo, int hi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ra, rb;
    warp_rows(lo, hi, warp, ra, rb);
    if (ra >= rb) return;
    unsigned int cin = 0xfu;
    if (EDGE && !GBC) cin = col_interior_bits(c, lane);
    float w[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) w[t] = w9[t];
    const int lofs = 4 * lane;
    for_rows3(src, ra, rb, lane, [&](int r, const Row6 &t, const Row6 &m, const Row6 &b) {
        float o[4];
        stencil4(w, t, m, b, o);
        const int o4 = r * BW + lofs;
        if (GBC) {
            const float4 g = *reinterpret_cast<const float4 *>(c.IDX + o4);
            o[0] = __fmul_rn(o[0], g.x);
            o[1] = __fmul_rn(o[1], g.y);
            o[2] = __fmul_rn(o[2], g.z);
            o[3] = __fmul_rn(o[3], g.w);
        } else if (EDGE) {
            const int gy = c.gy0 + r;
            const unsigned int mk = (gy >= 1 && gy <= c.N - 2) ? cin : 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = ((mk >> e) & 1u) ? o[e] : 0.0f;
        }
        if (ADD) {
            const float4 jv = *reinterpret_cast<const float4 *>(base + o4);
            o[0] = __fadd_rn(jv.x, o[0]);
            o[1] = __fadd_rn(jv.y, o[1]);
            o[2] = __fadd_rn(jv.z, o[2]);
            o[3] = __fadd_rn(jv.w, o[3]);
        }
        *reinterpret_cast<float4 *>(dst + o4) = make_float4(o[0], o[1], o[2], o[3]);
    });
}

// ---------------------------------------------------------------------------------------------------------
// Output stage on rows [lo,hi): v = K src (MODE OUT_KU) or f - K src (others).
//   OUT_RESIDUAL / OUT_KU : store tile-interior values to global r_out
//   OUT_RESTRICT          : write r into rdst (smem) for stage_restrict
//   OUT_NORM              : accumulate sum of squares over interior nodes of the tile interior
template <bool KEYS, int MODE, bool EDGE>
__device__ __forceinline__ double stage_out(const TileCtx &c, const Tables &T, const RegW &W, const TileParams &p,
                                            const float *src, float *rdst, int lo, int hi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ra, rb;
    warp_rows(lo, hi, warp, ra, rb);
    double part = 0.0;
    if (ra >= rb) return part;
    const bool slow = KEYS && !c.keys_uniform;
    const bool lane_int = (4 * lane >= p.HX) && (4 * lane < BW - p.HX);
    unsigned int cin = 0xfu, cdom = 0xfu;
    if (EDGE) {
        cin = col_interior_bits(c, lane);
        cdom = col_domain_bits(c, lane);
    }
    const int lofs = 4 * lane;
    auto emit = [&](int r, float (&o)[4]) {
        const int o4 = r * BW + lofs;
        if (MODE == OUT_RESTRICT) {
            *reinterpret_cast<float4 *>(rdst + o4) = make_float4(o[0], o[1], o[2], o[3]);
        } else if (MODE == OUT_NORM) {
            bool ok = lane_int && (r >= p.HT) && (r < p.HT + p.TH);
            unsigned int mk = 0xfu;
            if (EDGE) {
                const int gy = c.gy0 + r;
                ok = ok && (gy >= 1 && gy <= c.N - 2);
                mk = cin;
            }
            if (ok) {
                float s0 = (mk & 1u) ? o[0] : 0.0f, s1 = (mk & 2u) ? o[1] : 0.0f;
                float s2 = (mk & 4u) ? o[2] : 0.0f, s3 = (mk & 8u) ? o[3] : 0.0f;
                part += (double)s0 * (double)s0 + (double)s1 * (double)s1;
                part += (double)s2 * (double)s2 + (double)s3 * (double)s3;
            }
        } else {  // global store of the tile interior; columns >= N of a straddling chunk are written as zero
            bool ok = lane_int && (r >= p.HT) && (r < p.HT + p.TH);
            const int gy = c.gy0 + r;
            if (EDGE) {
                ok = ok && (gy >= 0 && gy <= c.N - 1) && (cdom & 1u);
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = ((cdom >> e) & 1u) ? o[e] : 0.0f;
            }
            if (ok) {
                float *gp = p.r_out + (long long)c.b * p.plane + (long long)gy * p.pitch + (c.gx0 + lofs);
                st_global_v4(gp, make_float4(o[0], o[1], o[2], o[3]));
            }
        }
    };
    if (!slow) {
        for_rows3(src, ra, rb, lane, [&](int r, const Row6 &t, const Row6 &m, const Row6 &b) {
            float acc[4], o[4];
            stencil4(W.kw, t, m, b, acc);
            if (MODE == OUT_KU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = acc[e];
            } else {
                const float4 fv = *reinterpret_cast<const float4 *>(c.F + r * BW + lofs);
                o[0] = __fsub_rn(fv.x, acc[0]);
                o[1] = __fsub_rn(fv.y, acc[1]);
                o[2] = __fsub_rn(fv.z, acc[2]);
                o[3] = __fsub_rn(fv.w, acc[3]);
            }
            emit(r, o);
        });
        return part;
    }
    Row6 t = load_row6(src + (ra - 1) * BW + lofs), m = load_row6(src + ra * BW + lofs);
    Key6 kt = load_key6(c.K, ra - 1, lane), km = load_key6(c.K, ra, lane), kb;
    for (int r = ra; r < rb; ++r) {
        const Row6 b = load_row6(src + (r + 1) * BW + lofs);
        float acc[4], o[4];
        kb = load_key6(c.K, r + 1, lane);
        stencil4_keys(T.ktab, t, m, b, kt, km, kb, acc);
        kt = km;
        km = kb;
        if (MODE == OUT_KU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = acc[e];
        } else {
            const float4 fv = *reinterpret_cast<const float4 *>(c.F + r * BW + lofs);
            o[0] = __fsub_rn(fv.x, acc[0]);
            o[1] = __fsub_rn(fv.y, acc[1]);
            o[2] = __fsub_rn(fv.z, acc[2]);
            o[3] = __fsub_rn(fv.w, acc[3]);
        }
        emit(r, o);
        t = m;
        m = b;
    }
    return part;
}

// ---------------------------------------------------------------------------------------------------------
// Full-weighting style restriction of r (smem, valid on rows [HT-1, HT+TH]) to the coarse rows owned by this tile:
//   fc[I][J] = scale * chain_{a,c} R[key(src)][3a+c] * r[2I-1+a][2J-1+c],  1 <= I,J <= Nc-2, ring = 0
template <bool KEYS, bool EDGE>
__device__ __forceinline__ void stage_restrict(const TileCtx &c, const Tables &T, const RegW &W, const TileParams &p,
                                               const float *rbuf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool lane_int = (4 * lane >= p.HX) && (4 * lane < BW - p.HX);
    const bool slow = KEYS && (p.rtab_n > 1) && !c.keys_uniform;
    const int ncr = p.TH / 2;
    const int lofs = 4 * lane;
    const int J0 = (c.gx0 + lofs) >> 1;  // even
    for (int q = warp; q < ncr; q += NWARPS) {
        const int r = p.HT + 2 * q;  // box row of fine row 2I
        const int I = (c.gy0 + r) >> 1;
        if (EDGE && I > p.Nc - 1) break;
        const float *pr = rbuf + (r - 1) * BW + lofs;
        const Row6 t = load_row6(pr), m = load_row6(pr + BW), b = load_row6(pr + 2 * BW);
        float o[2];
        if (slow) {
            const Key6 kt = load_key6(c.K, r - 1, lane), km = load_key6(c.K, r, lane), kb = load_key6(c.K, r + 1, lane);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = 2 * h;  // fine column 4l+2h is the centre -> a[e+1]
                float s = __fmul_rn(T.rtab[9 * kt.k[e] + 0], t.a[e]);
                s = __fmaf_rn(T.rtab[9 * kt.k[e + 1] + 1], t.a[e + 1], s);
                s = __fmaf_rn(T.rtab[9 * kt.k[e + 2] + 2], t.a[e + 2], s);
                s = __fmaf_rn(T.rtab[9 * km.k[e] + 3], m.a[e], s);
                s = __fmaf_rn(T.rtab[9 * km.k[e + 1] + 4], m.a[e + 1], s);
                s = __fmaf_rn(T.rtab[9 * km.k[e + 2] + 5], m.a[e + 2], s);
                s = __fmaf_rn(T.rtab[9 * kb.k[e] + 6], b.a[e], s);
                s = __fmaf_rn(T.rtab[9 * kb.k[e + 1] + 7], b.a[e + 1], s);
                s = __fmaf_rn(T.rtab[9 * kb.k[e + 2] + 8], b.a[e + 2], s);
                o[h] = s;
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = 2 * h;
                float s = __fmul_rn(W.rw[0], t.a[e]);
                s = __fmaf_rn(W.rw[1], t.a[e + 1], s);
                s = __fmaf_rn(W.rw[2], t.a[e + 2], s);
                s = __fmaf_rn(W.rw[3], m.a[e], s);
                s = __fmaf_rn(W.rw[4], m.a[e + 1], s);
                s = __fmaf_rn(W.rw[5], m.a[e + 2], s);
                s = __fmaf_rn(W.rw[6], b.a[e], s);
                s = __fmaf_rn(W.rw[7], b.a[e + 1], s);
                s = __fmaf_rn(W.rw[8], b.a[e + 2], s);
                o[h] = s;
            }
        }
        if (p.r_has_scale) {
            o[0] = __fmul_rn(T.r_scale, o[0]);
            o[1] = __fmul_rn(T.r_scale, o[1]);
        }
        bool ok = lane_int;
        if (EDGE) {
            ok = ok && (J0 <= p.Nc - 1);
            const bool Iin = (I >= 1 && I <= p.Nc - 2);
            o[0] = (Iin && J0 >= 1 && J0 <= p.Nc - 2) ? o[0] : 0.0f;
            o[1] = (Iin && J0 + 1 <= p.Nc - 2) ? o[1] : 0.0f;
        }
        if (ok) {
            float *gp = p.fc + (long long)c.b * p.plane_c + (long long)I * p.pitch_c + J0;
            *reinterpret_cast<float2 *>(gp) = make_float2(o[0], o[1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Prolongation + correction on all box rows: U += scale * P(vc)   (coarse box VC, coarse keys KC)
template <bool GBC, bool EDGE, int PM = -1>
__device__ __forceinline__ void stage_prolong(const TileCtx &c, const Tables &T, const TileParams &p, float *U) {
    const int prolong_mode = (PM >= 0) ? PM : p.prolong_mode;  // PM: the specialised programs of mg_tile_kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!EDGE && !GBC && prolong_mode == 1) {
        // interior tile, bilinear: every node of the box is an interior node of the domain, no masks needed
        const float *vbase = c.VC + c.vcofs + 2 * lane;  // 8-byte aligned (vcofs is even)
        const bool seq = p.prolong_seq != 0;
        for (int r = warp; r < c.BH; r += NWARPS) {
            const int gy = c.gy0 + r;
            const float *vt = vbase + ((gy >> 1) - c.cy0) * CW;
            const float2 t01 = *reinterpret_cast<const float2 *>(vt);
            const float t2 = vt[2];
            float e0, e1, e2, e3;
            if (!(gy & 1)) {
                e0 = t01.x;
                e1 = __fadd_rn(__fmul_rn(0.5f, t01.x), __fmul_rn(0.5f, t01.y));
                e2 = t01.y;
                e3 = __fadd_rn(__fmul_rn(0.5f, t01.y), __fmul_rn(0.5f, t2));
            } else {
                const float2 b01 = *reinterpret_cast<const float2 *>(vt + CW);
                const float b2 = vt[CW + 2];
                e0 = __fadd_rn(__fmul_rn(0.5f, t01.x), __fmul_rn(0.5f, b01.x));
                e2 = __fadd_rn(__fmul_rn(0.5f, t01.y), __fmul_rn(0.5f, b01.y));
                if (seq) {
                    e1 = __fadd_rn(__fmul_rn(0.25f, t01.x), __fmul_rn(0.25f, t01.y));
                    e1 = __fadd_rn(e1, __fmul_rn(0.25f, b01.x));
                    e1 = __fadd_rn(e1, __fmul_rn(0.25f, b01.y));
                    e3 = __fadd_rn(__fmul_rn(0.25f, t01.y), __fmul_rn(0.25f, t2));
                    e3 = __fadd_rn(e3, __fmul_rn(0.25f, b01.y));
                    e3 = __fadd_rn(e3, __fmul_rn(0.25f, b2));
                } else {
                    const float ta = __fadd_rn(__fmul_rn(0.5f, t01.x), __fmul_rn(0.5f, t01.y));
                    const float ba = __fadd_rn(__fmul_rn(0.5f, b01.x), __fmul_rn(0.5f, b01.y));
                    e1 = __fadd_rn(__fmul_rn(0.5f, ta), __fmul_rn(0.5f, ba));
                    const float tb = __fadd_rn(__fmul_rn(0.5f, t01.y), __fmul_rn(0.5f, t2));
                    const float bb = __fadd_rn(__fmul_rn(0.5f, b01.y), __fmul_rn(0.5f, b2));
                    e3 = __fadd_rn(__fmul_rn(0.5f, tb), __fmul_rn(0.5f, bb));
                }
            }
            float4 uv = *reinterpret_cast<float4 *>(U + r * BW + 4 * lane);
            uv.x = __fadd_rn(uv.x, e0);
            uv.y = __fadd_rn(uv.y, e1);
            uv.z = __fadd_rn(uv.z, e2);
            uv.w = __fadd_rn(uv.w, e3);
            *reinterpret_cast<float4 *>(U + r * BW + 4 * lane) = uv;
        }
        return;
    }
    const unsigned int cin = col_interior_bits(c, lane);
    const unsigned int cdom = col_domain_bits(c, lane);
    const int cc = 2 * lane;  // coarse box column of fine box column 4*lane (gx0 even, cx0 = gx0/2)
    const bool table = (prolong_mode == 3);
    const bool pkeys = table && (p.ptab_n > 1) && (c.KC != nullptr);
    for (int r = warp; r < c.BH; r += NWARPS) {
        const int gy = c.gy0 + r;
        if (gy < 0 || gy > c.N - 1) continue;
        const int I0 = (gy >> 1) - c.cy0;  // coarse box row of floor(gy/2)
        const bool odd = gy & 1;
        float top[3], bot[3] = {0.f, 0.f, 0.f};
        int ktop[3] = {0, 0, 0}, kbot[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            top[q] = c.VC[I0 * CW + c.vcofs + cc + q];
            if (odd) bot[q] = c.VC[(I0 + 1) * CW + c.vcofs + cc + q];
            if (pkeys) {
                ktop[q] = c.KC[I0 * KCW + c.kcofs + cc + q];
                if (odd) kbot[q] = c.KC[(I0 + 1) * KCW + c.kcofs + cc + q];
            }
        }
        float e[4];
        if (!table) {
            if (!odd) {
                e[0] = top[0];
                e[1] = __fadd_rn(__fmul_rn(0.5f, top[0]), __fmul_rn(0.5f, top[1]));
                e[2] = top[1];
                e[3] = __fadd_rn(__fmul_rn(0.5f, top[1]), __fmul_rn(0.5f, top[2]));
            } else {
                e[0] = __fadd_rn(__fmul_rn(0.5f, top[0]), __fmul_rn(0.5f, bot[0]));
                e[2] = __fadd_rn(__fmul_rn(0.5f, top[1]), __fmul_rn(0.5f, bot[1]));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float a = top[h], bq = top[h + 1], cq = bot[h], d = bot[h + 1];
                    float v;
                    if (p.prolong_seq) {
                        v = __fadd_rn(__fmul_rn(0.25f, a), __fmul_rn(0.25f, bq));
                        v = __fadd_rn(v, __fmul_rn(0.25f, cq));
                        v = __fadd_rn(v, __fmul_rn(0.25f, d));
                    } else {
                        const float tp = __fadd_rn(__fmul_rn(0.5f, a), __fmul_rn(0.5f, bq));
                        const float bt = __fadd_rn(__fmul_rn(0.5f, cq), __fmul_rn(0.5f, d));
                        v = __fadd_rn(__fmul_rn(0.5f, tp), __fmul_rn(0.5f, bt));
                    }
                    e[2 * h + 1] = v;
                }
            }
            // variant A: the interpolated correction goes through the fine level's reset_boundary
            if (GBC) {
                const float4 g = *reinterpret_cast<const float4 *>(c.IDX + r * BW + 4 * lane);
                const float4 bv = *reinterpret_cast<const float4 *>(c.BVAL + r * BW + 4 * lane);
                e[0] = __fadd_rn(__fmul_rn(e[0], g.x), bv.x);
                e[1] = __fadd_rn(__fmul_rn(e[1], g.y), bv.y);
                e[2] = __fadd_rn(__fmul_rn(e[2], g.z), bv.z);
                e[3] = __fadd_rn(__fmul_rn(e[3], g.w), bv.w);
            } else {
                const bool rin = (gy >= 1 && gy <= c.N - 2);
#pragma unroll
                for (int q = 0; q < 4; ++q) e[q] = (rin && ((cin >> q) & 1u)) ? e[q] : 0.0f;
            }
        } else {
            // transposed conv taps in (a asc, c asc) order; a=y+1-2I: even row -> a=1 (I=I0); odd -> a=0 (I0+1), a=2 (I0)
            const float *P = T.ptab;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int jl = q >> 1;  // left coarse neighbour (box column cc + jl)
                float s = 0.0f;
                if (!odd) {
                    if (!(q & 1)) {
                        s = __fmul_rn(P[9 * ktop[jl] + 4], top[jl]);
                    } else {  // c=0 -> right node, c=2 -> left node
                        s = __fmul_rn(P[9 * ktop[jl + 1] + 3], top[jl + 1]);
                        s = __fmaf_rn(P[9 * ktop[jl] + 5], top[jl], s);
                    }
                } else {
                    if (!(q & 1)) {  // a=0 -> lower coarse row (I0+1), a=2 -> upper (I0)
                        s = __fmul_rn(P[9 * kbot[jl] + 1], bot[jl]);
                        s = __fmaf_rn(P[9 * ktop[jl] + 7], top[jl], s);
                    } else {
                        s = __fmul_rn(P[9 * kbot[jl + 1] + 0], bot[jl + 1]);
                        s = __fmaf_rn(P[9 * kbot[jl] + 2], bot[jl], s);
                        s = __fmaf_rn(P[9 * ktop[jl + 1] + 6], top[jl + 1], s);
                        s = __fmaf_rn(P[9 * ktop[jl] + 8], top[jl], s);
                    }
                }
                e[q] = p.p_has_scale ? __fmul_rn(T.p_scale, s) : s;
            }
        }
        float4 uv = *reinterpret_cast<float4 *>(U + r * BW + 4 * lane);
        uv.x = (cdom & 1u) ? __fadd_rn(uv.x, e[0]) : 0.0f;
        uv.y = (cdom & 2u) ? __fadd_rn(uv.y, e[1]) : 0.0f;
        uv.z = (cdom & 4u) ? __fadd_rn(uv.z, e[2]) : 0.0f;
        uv.w = (cdom & 8u) ? __fadd_rn(uv.w, e[3]) : 0.0f;
        *reinterpret_cast<float4 *>(U + r * BW + 4 * lane) = uv;
    }
}

// store the tile interior of a box buffer to global u_out (columns >= N of a straddling chunk stay zero)
__device__ __forceinline__ void stage_store_u(const TileCtx &c, const TileParams &p, const float *src) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool lane_int = (4 * lane >= p.HX) && (4 * lane < BW - p.HX);
    const unsigned int cdom = col_domain_bits(c, lane);
    if (!lane_int || !(cdom & 1u)) return;
    for (int r = p.HT + warp; r < p.HT + p.TH; r += NWARPS) {
        const int gy = c.gy0 + r;
        if (gy > c.N - 1) break;
        float4 v = *reinterpret_cast<const float4 *>(src + r * BW + 4 * lane);
        v.y = (cdom & 2u) ? v.y : 0.0f;
        v.z = (cdom & 4u) ? v.z : 0.0f;
        v.w = (cdom & 8u) ? v.w : 0.0f;
        float *gp = p.u_out + (long long)c.b * p.plane + (long long)gy * p.pitch + (c.gx0 + 4 * lane);
        st_global_v4(gp, v);
    }
}

}  // namespace mgfea
