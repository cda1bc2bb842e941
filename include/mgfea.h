/*
 * mgfea.h -- C ABI of libmgfea.so: the B200 (sm_100a) multigrid V-cycle kernels behind the FEANet module API.
 *
 * The reference (longfish/Multigrid-FEANet) has no FFI of its own: its hot path is Python calling CPU ATen ops.
 * Each entry point below replaces the ATen call sequence of one reference method (file:line relative to the
 * reference tree); the Python package multigrid-feanet_b200/FEANet binds them through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MGFEA_E* code for argument errors, or a positive
 *     cudaError_t; nothing throws, allocates device memory for the caller, or synchronises the device
 *     (two small internal scratch buffers per device are created lazily at first use);
 *   - all pointers are DEVICE pointers owned by the caller; `stream` is a cudaStream_t passed as void*;
 *   - fields are fp32, row-major, [B][N][pitch]: `pitch` = row pitch in floats, `plane` = sample stride in floats.
 *     Kernels require pitch % 4 == 0 and 16-byte aligned base pointers ("padded-pitch" level buffers; N = 2^k+1 is
 *     odd, so contiguous N*N tensors are packed/unpacked at the API edge with mgfea_pack / mgfea_unpack).
 *     Columns [N, pitch) of every field are kept zero by all kernels;
 *   - material pattern keys are uint8 [N][key_pitch] (key_pitch % 16 == 0), NULL for single-pattern (iso) meshes;
 *   - 3x3 tables are [npat][9] fp32, tap t = 3*(di+1)+(dj+1), read from device memory at launch time so that live
 *     nn.Parameter weights (KNet.net2.weight, FNet.net.weight, RestrictionNet/ProlongationNet.net.weight,
 *     HNet.convLayers[i].weight, MultiGrid.w) behave as in the reference.
 *   - arithmetic: IEEE fp32 in a fixed order (row-major FMA chain per stencil), identical to oracle/mgfea_oracle.c.
 */
#ifndef MGFEA_H
#define MGFEA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGFEA_EINVAL (-1)     /* bad argument (NULL pointer, bad size) */
#define MGFEA_EALIGN (-2)     /* pointer/pitch alignment requirement violated */
#define MGFEA_EUNSUPPORTED (-3)
#define MGFEA_EDRIVER (-4)    /* cuTensorMapEncodeTiled unavailable / failed */

/* One multigrid level ("grid").  Mirrors the per-level state of the reference's SingleGrid
 * (FEANet/multigrid.py:12-47): Knet weights + pattern map, Jacobi diagonal, Dirichlet masks. */
typedef struct mgfea_grid {
    int32_t N;          /* nodes per edge (n+1) */
    int32_t pitch;      /* floats per row of every field of this level */
    int64_t plane;      /* floats per sample */
    int32_t npat;       /* number of material patterns C: 1 (MeshSquare) or 16 (MeshCenterInterface) */
    int32_t key_pitch;  /* bytes per row of `keys` */
    const uint8_t *keys;   /* [N][key_pitch] pattern key per node (FEANet/mesh.py:95-101), or NULL */
    const float *ktab;     /* [npat][9]  KNet.net2.weight[0]            (FEANet/model.py:16,20) */
    const float *invd;     /* [npat]     omega / d per pattern key      (FEANet/jacobi.py:31-37,46) */
    const float *bc_idx;   /* optional general geometry_idx  [B or 1][N][pitch] (FEANet/jacobi.py:19,29); NULL = square ring */
    const float *bc_val;   /* optional general boundary_value, same layout; NULL = 0 */
    int64_t bc_plane;      /* sample stride of bc_idx/bc_val in floats; 0 = shared by all samples */
} mgfea_grid;

/* smoother selection */
#define MGFEA_SMOOTH_JACOBI 0 /* JacobiBlock.jacobi_convolution       FEANet/jacobi.py:39-47 */
#define MGFEA_SMOOTH_HJACOBI 1 /* HJacIterator.HRelax + HNet.forward   M-FEANet-mg_test.ipynb cells 4,5 */

/* prolongation selection */
#define MGFEA_PROLONG_BILINEAR 1 /* F.interpolate(bilinear, align_corners) + fine reset_boundary: MM_Model_convergence.ipynb cell 3 `Interpolate` */
#define MGFEA_PROLONG_TABLE 3    /* ConvTranspose2d(C->1,3,stride 2,pad 1) [* w[1]]: FEANet/multigrid.py:62-73,124-130,177-179 */

/* residual-norm convergence rule evaluated on the device */
#define MGFEA_CONV_SUM 0 /* whole-batch sum of squares  (Multigrid.Solve, MM_Model_convergence.ipynb cell 3) */
#define MGFEA_CONV_MAX 1 /* max over samples            (per-sample torch.norm, M-FEANet-mg_test.ipynb cell 21) */

/* Device-resident solve control block (one per solve; zero it before the first cycle).  Every kernel of a cycle
 * returns immediately when `done` is set, so cycles enqueued past convergence are no-ops and the host may check
 * convergence every few cycles instead of every cycle (the reference syncs with .item() each cycle). */
typedef struct mgfea_ctl {
    int32_t cycle;      /* number of residual norms recorded so far */
    int32_t done;       /* set by the device when converged / max_cycles reached */
    int32_t min_cycles; /* Solve's n_iter */
    int32_t max_cycles; /* capacity of hist (in cycles) */
    int32_t conv_rule;  /* MGFEA_CONV_SUM | MGFEA_CONV_MAX */
    int32_t pad_;
    double eps2;        /* EPS^2 (absolute interior 2-norm threshold, squared); <0 disables */
} mgfea_ctl;

/* ---- library / device -------------------------------------------------------------------------------- */
const char *mgfea_version(void);
const char *mgfea_error_string(int code);
/* 0 = cp.async tile loader, 1 = TMA (cp.async.bulk.tensor) tile loader [default]; returns previous value */
int mgfea_set_loader(int use_tma);
/* kernel-selection thresholds (the MGFEA_* environment knobs of csrc/mgfea.cu at run time; changing one invalidates
 * captured CUDA graphs): "hstream_min_n", "hstream_keys", "hstream_r", "hstream_over", "stream_min_n", "stream_keys",
 * "stream_one_variant", "stream_one_variant_max_n", "stream_one_variant_up_max_n", "tile_prog", "mid_keys",
 * "mid_max_n", "mid_max_n_up".
 * Returns the previous value (>= 0), MGFEA_EINVAL for an unknown name. */
int mgfea_set_option(const char *name, int value);
/* profiling aid: while buf != NULL a one-thread kernel stores %globaltimer (ns) into buf[i++] before and after every
 * fused-leg launch of mgfea_vcycle (i restarts at 0 on every call of mgfea_trace); buf = NULL switches it off */
int mgfea_trace(unsigned long long *buf, int capacity);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
uint64_t mgfea_launch_count(void);

/* ---- setup ------------------------------------------------------------------------------------------- */
/* pattern key map of MeshCenterInterface (FEANet/mesh.py:62-101: the reference loops nodes x elements, O(N^4)) in closed
 * form on the device: keys[N][key_pitch] uint8 (key_pitch % 16 == 0, padding bytes 0); shape 0 = circle r=0.5,
 * 1 = square half-width 0.5, anything else = single phase */
int mgfea_pattern_keys(uint8_t *keys, int N, int key_pitch, int shape, void *stream);

/* ---- layout ------------------------------------------------------------------------------------------ */
/* contiguous [B][N][N] <-> padded [B][N][pitch]; pack zero-fills columns [N,pitch) */
int mgfea_pack(const float *src, float *dst, int N, int pitch, int64_t plane, int B, void *stream);
int mgfea_unpack(const float *src, float *dst, int N, int pitch, int64_t plane, int B, void *stream);

/* ---- operators (one reference method each) ----------------------------------------------------------- */
/* KNet.forward (FEANet/model.py:22-30): out = K u on ALL nodes, zero padding, weights indexed by SOURCE-node key */
int mgfea_stiffness_apply(const mgfea_grid *g, const float *u, float *out, int B, void *stream);
/* FNet.forward (FEANet/model.py:49-61) / any single 3x3 correlation: out = w9 (*) x, zero padding */
int mgfea_load_vector(const float *w9, const float *x, float *out, int N, int pitch, int64_t plane, int B,
                      void *stream);
/* KNet.split_x (FEANet/model.py:37-47): out[b][c] = x[b] where key==c else 0; out contiguous [B][C][N][N] */
int mgfea_split_x(const mgfea_grid *g, const float *x, float *out, int B, void *stream);
/* JacobiBlock.reset_boundary (FEANet/jacobi.py:27-29) */
int mgfea_reset_boundary(const mgfea_grid *g, const float *u, float *out, int B, void *stream);
/* nsweeps of the smoother, temporally blocked inside one tile pass where the halo allows.
 * hw = [nlayers][9] HNet weights (HJACOBI only).  u_in may NOT alias u_out. */
int mgfea_smooth(const mgfea_grid *g, const float *u_in, float *u_out, const float *f, int nsweeps, int smoother,
                 const float *hw, int nlayers, int B, void *stream);
/* r = f - K u on all nodes (the `f - Knet(v)` expression of every driver) */
int mgfea_residual(const mgfea_grid *g, const float *u, const float *f, float *r, int B, void *stream);
/* Restrict (MM_Model_convergence.ipynb cell 3; FEANet/multigrid.py:115-122): fc = scale * R (*) r[1:-1,1:-1] stride 2,
 * zero ring.  rtab = [rtab_n][9], rtab_n = 1 or g->npat (table of the FINE source node's key).
 * scale: *scale_dev if non-NULL, else scale_host; has_scale = 0 skips the multiply. */
int mgfea_restrict(const mgfea_grid *g, const float *r, float *fc, int pitch_c, int64_t plane_c, const float *rtab,
                   int rtab_n, int has_scale, float scale_host, const float *scale_dev, int B, void *stream);
/* API-compat forms taking an already split (B,C,.,.) CONTIGUOUS tensor, as MultiGrid.Restrict / .Interpolate of
 * FEANet/multigrid.py:115-130 receive it: fc (B,1,Nc,Nc) = pad(conv(C->1,3x3,stride 2)(rF[:, :, 1:-1, 1:-1])),
 * out (B,1,2Nc-1,2Nc-1) = convT(C->1,3x3,stride 2,pad 1)(eFC).  No scale is applied. */
int mgfea_restrict_channels(const float *rF, float *fc, const float *rtab, int C, int N, int B, void *stream);
int mgfea_prolong_channels(const float *eFC, float *out, const float *ptab, int C, int Nc, int B, void *stream);
/* fused: nsweeps pre-smoothing (u_in==NULL means u_in = 0), write u_out, then fc = scale*R(f - K u_out) */
int mgfea_smooth_residual_restrict(const mgfea_grid *g, const float *u_in, float *u_out, const float *f, int nsweeps,
                                   int smoother, const float *hw, int nlayers, float *fc, int pitch_c,
                                   int64_t plane_c, const float *rtab, int rtab_n, int has_scale, float scale_host,
                                   const float *scale_dev, int B, void *stream);
/* Interpolate + correct (+ post-smooth): u_out = smooth^nsweeps(u_in + scale * P(vc)).
 * mode BILINEAR: P = bilinear x2 followed by the fine level's reset_boundary (variant A);
 * mode TABLE: transposed conv with ptab[ptab_n][9] indexed by the COARSE node's key (gc->keys), times scale. */
int mgfea_prolong_correct_smooth(const mgfea_grid *g, const mgfea_grid *gc, const float *vc, const float *u_in,
                                 float *u_out, const float *f, int mode, const float *ptab, int ptab_n,
                                 int has_scale, float scale_host, const float *scale_dev, int nsweeps, int smoother,
                                 const float *hw, int nlayers, int B, void *stream);
/* the same with the interior residual sum of squares of the result fused into the last launch (what the finest level
 * of mgfea_vcycle runs when compute_norm is set): sumsq[b] = sum over interior nodes of (f - K u_out)^2 */
int mgfea_prolong_correct_smooth_norm(const mgfea_grid *g, const mgfea_grid *gc, const float *vc, const float *u_in,
                                      float *u_out, const float *f, int mode, const float *ptab, int ptab_n,
                                      int has_scale, float scale_host, const float *scale_dev, int nsweeps,
                                      int smoother, const float *hw, int nlayers, double *sumsq, int B, void *stream);
/* sumsq[b] = sum over interior nodes of (f - K u)^2, accumulated in fp64, deterministic order.
 * If ctl != NULL the value is also appended to hist[ctl->cycle*B + b] and the convergence rule is evaluated. */
int mgfea_residual_norm(const mgfea_grid *g, const float *u, const float *f, double *sumsq, mgfea_ctl *ctl,
                        double *hist, int B, void *stream);

/* ---- row-slab partition (multi-GPU) ------------------------------------------------------------------- */
/* Fine levels are split into contiguous row slabs, one per GPU.  A rank's arrays hold the global
 * rows [row0, row0+nrows) of the N x N level: its owned rows [own0, own1) plus ghost rows that the caller fills by halo
 * exchange (>= 3 above / below the owned range for the fused legs).  own0 must be even, so that fine row 2I and coarse
 * row I live on the same rank.  Single-pattern grids with the default Dirichlet ring only (streaming kernels). */
typedef struct mgfea_slab {
    int32_t row0;   /* global row index of local row 0 */
    int32_t nrows;  /* rows held locally (owned + ghost) */
    int32_t own0;   /* first owned global row */
    int32_t own1;   /* one past the last owned global row */
} mgfea_slab;
/* down leg on a slab: one Jacobi sweep (u_in == NULL: zero guess), u_out (owned rows), fc = scale*R(f - K u_out) on the
 * owned coarse rows [own0/2, own1/2) of the coarse slab `sc` (only sc->row0 / sc->nrows are used) */
int mgfea_slab_smooth_residual_restrict(const mgfea_grid *g, const mgfea_slab *s, const float *u_in, float *u_out,
                                        const float *f, float *fc, const mgfea_slab *sc, int pitch_c, int64_t plane_c,
                                        const float *rtab, int has_scale, float scale_host, const float *scale_dev,
                                        int B, void *stream);
/* up leg on a slab: u_out = smooth(u_in + reset(bilinear P vc)) on the owned rows; vc is the coarse slab `sc` (with
 * ghost rows); if sumsq != NULL the interior residual sum of squares over the OWNED rows is written per sample (the
 * caller all-reduces it) */
int mgfea_slab_prolong_correct_smooth(const mgfea_grid *g, const mgfea_slab *s, const float *vc, const mgfea_slab *sc,
                                      int pitch_c, int64_t plane_c, const float *u_in, float *u_out, const float *f,
                                      double *sumsq, int B, void *stream);

/* The same up leg with the halo push FUSED into the kernel (north_star: exchange overlapped with the sweep): the strips
 * that produce the first / last `rows` owned rows store them also straight into the neighbours' ghost rows (peer-mapped
 * memory, mgfea_peer_open), and the last such strip raises a flag in the neighbour's mailbox (one increment per launch and
 * direction).  A following mgfea_p2p_exchange step with wait2 = this rank's own flags makes the next kernel see complete
 * ghost rows.  up / dn: address of the element (GLOBAL row 0, column 0) of the neighbour's array in this process (may lie
 * before the mapped block; only rows the neighbour holds are touched), NULL = no neighbour on that side.
 * ctl (optional, read only): when ctl->done is set the kernel returns without storing, pushing or signalling -- the
 * solution of a converged slab solve stays what it was (the rule itself is evaluated by the reduce step, mgfea_xchg.ctl). */
typedef struct mgfea_slab_push {
    float *up, *dn;
    uint32_t *flag_up, *flag_dn; /* in the neighbours' mailboxes */
    uint32_t *ticket;            /* 2 zeroed words in local device memory, left at 0 by every launch */
    int32_t rows;                /* boundary rows to push per side (the receiver's ghost depth) */
    int32_t own0, own1;          /* the rank's owned global rows */
} mgfea_slab_push;
int mgfea_slab_prolong_correct_smooth_push(const mgfea_grid *g, const mgfea_slab *s, const float *vc,
                                           const mgfea_slab *sc, int pitch_c, int64_t plane_c, const float *u_in,
                                           float *u_out, const float *f, double *sumsq, const mgfea_slab_push *push,
                                           const mgfea_ctl *ctl, int B, void *stream);

/* ---- peer memory: halo exchange over NVLink without a collective library ----------- */
/* The reference has no distributed code; these entries carry the row-slab exchange of FEANet/distributed.py.  Slab
 * arrays and one mailbox per rank are allocated with mgfea_peer_alloc, exported as 64-byte handles that the host side
 * passes to the other ranks of the node (any byte transport), and mapped there with mgfea_peer_open. */
#define MGFEA_IPC_HANDLE_BYTES 64
int mgfea_peer_alloc(void **ptr, uint64_t bytes);      /* cudaMalloc + zero fill (synchronises once, at setup) */
int mgfea_peer_free(void *ptr);
int mgfea_peer_export(const void *ptr, void *handle);  /* handle[MGFEA_IPC_HANDLE_BYTES] of an mgfea_peer_alloc block */
int mgfea_peer_open(const void *handle, void **ptr);   /* map another rank's block (enables peer access lazily) */
int mgfea_peer_close(void *ptr);

#define MGFEA_XCHG_MAX_JOBS 16
#define MGFEA_XCHG_MAX_PEERS 8
#define MGFEA_XCHG_PUSH 1 /* run the copy jobs; every CTA then increments every `signal` flag once (system scope) */
#define MGFEA_XCHG_WAIT 2 /* wait until every `wait` flag has reached *seq + grid, then advance *seq by grid */
/* One exchange step of one rank (a single kernel, graph-capturable).  All addresses and sizes are multiples of 16 B. */
typedef struct mgfea_xchg {
    int32_t njobs, nsignal, nwait, mode;
    const void *src[MGFEA_XCHG_MAX_JOBS]; /* local rows */
    void *dst[MGFEA_XCHG_MAX_JOBS];       /* the same rows in a peer's array (ghost rows / gathered rows / slots) */
    uint64_t bytes[MGFEA_XCHG_MAX_JOBS];
    uint32_t *signal[MGFEA_XCHG_MAX_PEERS];     /* flags in the TARGET ranks' mailboxes */
    const uint32_t *wait[MGFEA_XCHG_MAX_PEERS]; /* flags in THIS rank's mailbox, incremented by the ranks pushing to it */
    uint32_t *seq;   /* this rank's expected flag value for this flag set (device memory, advanced by the kernel) */
    int32_t *err;    /* device word set to 1 + index of the flag whose wait timed out (dead peer); may be NULL */
    const double *red_src; /* optional: after the wait, *red_dst = sum of nred doubles spaced red_stride bytes apart */
    double *red_dst;
    int32_t nred, red_stride;
    int32_t grid;    /* CTAs of this step, 1..512: MUST be the same on every rank (a flag counts the pushing CTAs) */
    int32_t nwait2;  /* optional second flag set, raised ONCE per launch by the neighbours' fused-push kernels
                        (mgfea_slab_prolong_correct_smooth_push): wait until each has reached *seq2 + 1, advance *seq2 */
    const uint32_t *wait2[2];
    uint32_t *seq2;
    /* optional device-side stopping rule of the row-slab solve (all ranks reduce the same total in the same order, so
     * they all take the same decision): when ctl->done is set the whole step is skipped; otherwise, after the reduction,
     * hist[ctl->cycle] = total (while ctl->cycle < hist_cap), ctl->cycle++, and the mgfea_ctl rule is evaluated */
    mgfea_ctl *ctl;
    double *hist;
    int32_t hist_cap;
    int32_t pad2_;
} mgfea_xchg;
int mgfea_p2p_exchange(const mgfea_xchg *x, void *stream);

/* ---- fp64 defect correction around the fp32 cycle --------------------------------------- */
/* The reference's remedy for the fp32 residual floor is `.double()` on everything (MM_poisson.ipynb cell 5).  Here the
 * iterate, right-hand side and residual are fp64 [B][N][pitch] (same pitch / plane counts as the fp32 fields, 16-byte
 * aligned) and only the correction runs in fp32:  r = f - K u;  e = V-cycle(0, r);  u += e.  Default Dirichlet ring.
 * mgfea_defect_f64: r (fp32, 0 on the ring) and sumsq[b] = sum over interior nodes of (f - K u)^2 in fp64; with ctl
 * it is the residual history / stopping rule of the solve, like mgfea_residual_norm.
 * mgfea_correct_f64: u += (double) e on interior nodes (no-op once ctl->done is set). */
/* padded fp32 field -> padded fp64 field (same pitch / plane counts); zero_ring clears the Dirichlet ring on the way */
int mgfea_widen_f64(const float *src, double *dst, int N, int pitch, int64_t plane, int B, int zero_ring, void *stream);
int mgfea_defect_f64(const mgfea_grid *g, const double *u, const double *f, float *r, double *sumsq, mgfea_ctl *ctl,
                     double *hist, int B, void *stream);
int mgfea_correct_f64(const mgfea_grid *g, double *u, const float *e, const mgfea_ctl *ctl, int B, void *stream);
/* row-slab forms (arrays hold the global rows [s->row0, ..), g->plane = local rows * pitch): the defect is written on
 * the owned rows and on the 3 ghost rows per side the next down leg reads (u needs 4 valid ghost rows), sumsq covers
 * the owned rows only (the caller all-reduces it); the correction touches the owned rows */
int mgfea_slab_defect_f64(const mgfea_grid *g, const mgfea_slab *s, const double *u, const double *f, float *r,
                          double *sumsq, int B, void *stream);
int mgfea_slab_correct_f64(const mgfea_grid *g, const mgfea_slab *s, double *u, const float *e, int B, void *stream);
/* the same defect written on the owned rows and on `ext` rows per side (the deep-halo slab cycle computes its ghost rows
 * redundantly instead of exchanging them: FEANet/distributed.py); u must be valid on ext + 1 rows per side */
int mgfea_slab_defect_f64_ext(const mgfea_grid *g, const mgfea_slab *s, int ext, const double *u, const double *f,
                              float *r, double *sumsq, int B, void *stream);

/* JacobiBlockPBC.jacobi_convolution (FEANet/jacobi.py:50-97): one weighted-Jacobi sweep with PERIODIC boundary conditions,
 * single pattern.  u (N x N nodes, node N-1 == node 0); f_pad = the (N+2) x (N+2) load vector the reference makes its
 * caller pad (its Knet runs on the circularly padded (N+2)^2 array), padded-pitch layout of its own; w9 = the 3x3 kernel,
 * invd = omega/d (both in device memory).  out = invd * (f_pad[1:-1,1:-1] - K_periodic u) + reset_boundary(u). */
int mgfea_smooth_pbc(const float *w9, const float *invd, const float *u_in, float *u_out, const float *f_pad, int N,
                     int pitch, int64_t plane, int pitch_f, int64_t plane_f, int B, void *stream);

/* Backward pass of one HNet layer (the reference back-propagates through HJacIterator.HRelax with autograd,
 * M-FEANet-learn_iterator.ipynb cell 7): weight gradient of the zero-padded correlation out = w (*) a,
 * acc9[3 dy + dx] += sum_{b,i,j} a[b][i+dy-1][j+dx-1] * g[b][i][j] in fp64 (the caller zeroes acc9; atomics across blocks).
 * The input gradient is the same correlation with the flipped kernel (mgfea_load_vector), the adjoint of K is K. */
int mgfea_corr9(const float *a, const float *g, double *acc9, int N, int pitch, int64_t plane, int B, void *stream);

/* Backward of the table restriction / prolongation of FEANet/multigrid.py:50-73,115-130 (the reference trains R / P by
 * back-propagating through MultiGrid.iterate).  Tables [n][9], n = 1 or 16 (by the FINE source node's key for R, by the
 * COARSE node's key for P: g->keys / gc->keys); `scale` = w[0] / w[1].
 *   restrict_adjoint:  g_r  (fine)   = d(fc)/d(r)^T  g_fc      prolong_adjoint:  g_vc (coarse) = d(e)/d(vc)^T g_vf
 *   *_wgrad:           acc[n][9] (fp64, caller-zeroed, atomics) += table gradient */
int mgfea_restrict_adjoint(const mgfea_grid *g, const mgfea_grid *gc, const float *rtab, int rtab_n, float scale,
                           const float *g_fc, float *g_r, int B, void *stream);
int mgfea_prolong_adjoint(const mgfea_grid *g, const mgfea_grid *gc, const float *ptab, int ptab_n, float scale,
                          const float *g_vf, float *g_vc, int B, void *stream);
int mgfea_restrict_wgrad(const mgfea_grid *g, const mgfea_grid *gc, int rtab_n, float scale, const float *r,
                         const float *g_fc, double *acc, int B, void *stream);
int mgfea_prolong_wgrad(const mgfea_grid *g, const mgfea_grid *gc, int ptab_n, float scale, const float *vc,
                        const float *g_vf, double *acc, int B, void *stream);

/* ---- general per-element conductivity ---------------------------------------------------- */
/* The reference's data model carries one conductivity per ELEMENT (`material`, Data/dataset.py:71-104) but its operator
 * only knows the 16 two-phase patterns (FEANet/mesh.py:103-117).  These entries are that operator with the pattern lookup
 * replaced by the element values: a = [N][pitch] fp32 (16-byte aligned), element (r,c) at a[r*pitch + c] for r,c < N-1,
 * zero elsewhere; shared by the batch.  Weights are `generate_kernel`'s fp32 expressions of the source nodes, so on a
 * two-phase map the result equals the pattern operator bit for bit on every interior node.  Default Dirichlet ring. */
int mgfea_elem_stiffness_apply(const float *a, const float *u, float *out, int N, int pitch, int64_t plane, int B,
                               void *stream);                       /* KNet.forward, FEANet/model.py:22-30 */
int mgfea_elem_residual(const float *a, const float *u, const float *f, float *r, int N, int pitch, int64_t plane, int B,
                        void *stream);                              /* f - Knet(u) */
/* one sweep of JacobiBlock.jacobi_convolution (FEANet/jacobi.py:39-47), omega/d = fl(fl(1/d) * omega) per node with d
 * the centre entry of the node's own kernel (jacobi.py:31-37); u_in may not alias u_out */
int mgfea_elem_smooth(const float *a, const float *u_in, float *u_out, const float *f, float omega, int N, int pitch,
                      int64_t plane, int B, void *stream);
/* conductivity of the next coarser level: mean of the four child elements (fp32, row-major order).  Our convention: the
 * reference rediscretises its inclusion on every level instead (FEANet/multigrid.py:19-29) */
int mgfea_elem_coarsen(const float *a, float *ac, int N, int pitch, int pitch_c, void *stream);
/* sumsq[b] = sum over interior nodes of r^2 in fp64, deterministic (Solve's residual norm for the element operator) */
int mgfea_sumsq_interior(const float *r, double *sumsq, int N, int pitch, int64_t plane, int B, void *stream);

/* ---- whole V-cycle ----------------------------------------------------------------------------------- */
typedef struct mgfea_cycle_cfg {
    int32_t nu1, nu2;      /* pre / post sweeps (coarsest level gets nu1 + nu2) */
    int32_t smoother;      /* MGFEA_SMOOTH_* */
    int32_t nlayers;       /* HNet layers */
    const float *hw;       /* [nlayers][9] */
    int32_t prolong_mode;  /* MGFEA_PROLONG_* */
    int32_t rtab_n;        /* 1 or 16 */
    const float *rtab;     /* [rtab_n][9] */
    int32_t ptab_n;
    int32_t r_has_scale;
    const float *ptab;     /* [ptab_n][9] (TABLE mode) */
    float r_scale_host;
    float p_scale_host;
    const float *r_scale_dev; /* MultiGrid.w[0] */
    const float *p_scale_dev; /* MultiGrid.w[1] */
    int32_t p_has_scale;
    int32_t quirk_level0;  /* MM_Interface_error.ipynb cell 2: pre-smooth is always applied to level 0 */
    int32_t tail_max_n;    /* levels with N <= tail_max_n run inside the single coarse-tail kernel (0: default) */
    int32_t compute_norm;  /* 1: fuse the interior residual norm of level 0 into the last kernel */
    int32_t zero_guess;    /* 1: level 0 starts from u = 0 (bufs[0].u is output only): the replicated coarse cycle of the
                              row-slab path, i.e. the `v = zeros` of every coarse level (multigrid.py:171) */
    int32_t pad_;
} mgfea_cycle_cfg;

/* Per-level buffers of one V-cycle: u ping-pong pair and f.  After the call the result is in u[0] again. */
typedef struct mgfea_level_bufs {
    float *u;      /* solution (level 0: in/out; coarser: scratch) */
    float *u_alt;  /* ping-pong partner */
    float *f;      /* right-hand side (level 0: input; coarser: written by the restriction) */
} mgfea_level_bufs;

/* One V(nu1,nu2) cycle over `nlevels` grids (grids[0] finest).  Restates Multigrid.rec_V_cycle
 * (MM_Model_convergence.ipynb cell 3), MultiGrid.Step (M-FEANet-mg_test.ipynb cell 19) and MultiGrid.iterate
 * (FEANet/multigrid.py:159-185).  Graph-capturable: no host synchronisation, no allocation. */
int mgfea_vcycle(const mgfea_grid *grids, const mgfea_level_bufs *bufs, int nlevels, const mgfea_cycle_cfg *cfg,
                 double *sumsq, mgfea_ctl *ctl, double *hist, int B, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MGFEA_H */
