/*
 * mgfea_oracle.c -- CPU restatement of the Multigrid-FEANet V-cycle operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (multigrid-feanet_b200/)
 * may link, import or call this file; it is the checker used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * Parity status: PINNED.  Every function below is checked against outputs of the
 * unmodified reference (imported from /root/reference in the build container by
 * tests/golden/make_golden.py) through the committed fixtures in tests/golden/.
 *
 * All arithmetic is IEEE fp32 with a FIXED operation order (no contraction: build
 * with -ffp-contract=off; the only fused operations are the explicit fmaf() calls),
 * so that the CUDA kernels, which use the same order with __fmaf_rn/__fadd_rn/...,
 * can be compared bit-for-bit.  The reference itself (ATen/oneDNN) uses an
 * unspecified summation order, hence oracle-vs-reference agreement is to fp32
 * rounding (<= a few ulp per operator), oracle-vs-CUDA agreement is exact.
 *
 * Layout: fields are contiguous row-major [B][N][N] float; keys [N][N] uint8
 * (NULL = single pattern 0); tables [C][9] float, tap t = 3*(di+1)+(dj+1).
 *
 * Reference lines restated (relative to /root/reference):
 *   orc_stiffness_apply   FEANet/model.py:22-30   (KNet.forward: split, mask by SOURCE-node pattern, conv)
 *   orc_split_x           FEANet/model.py:37-47
 *   orc_reset_boundary    FEANet/jacobi.py:27-29
 *   orc_jacobi            FEANet/jacobi.py:39-47
 *   orc_hjacobi           M-FEANet-mg_test.ipynb cell 4 (HNet.forward) + cell 5 (HJacIterator.HRelax)
 *   orc_residual          f - Knet(v): MM_Model_convergence.ipynb cell 3 rec_V_cycle; FEANet/multigrid.py:168
 *   orc_restrict          MM_Model_convergence.ipynb cell 3 Restrict (+ "4*"); FEANet/multigrid.py:115-122,50-60,170
 *   orc_prolong_bilinear  MM_Model_convergence.ipynb cell 3 Interpolate (F.interpolate bilinear align_corners + reset_boundary)
 *   orc_prolong_table     FEANet/multigrid.py:62-73,124-130,177-179 (ConvTranspose2d(C->1,3,stride 2,pad 1), *w[1], +)
 *   orc_sumsq_interior    MM_Model_convergence.ipynb cell 3 Solve (sum(residual[:,:,1:-1,1:-1]**2))
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX(i, j) ((size_t)(i) * (size_t)N + (size_t)(j))

static inline int key_at(const uint8_t *keys, int N, int i, int j) { return keys ? keys[IDX(i, j)] : 0; }

/* 3x3 "source-indexed" stencil at (i,j): sum_t W[key(src_t)][t] * u[src_t], zero padding, row-major fma chain */
static inline float stencil_src(const float *u, const uint8_t *keys, const float *tab, int N, int i, int j) {
    float acc = 0.0f;
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
            int ii = i + di, jj = j + dj;
            if (ii < 0 || ii >= N || jj < 0 || jj >= N) continue; /* fmaf(w,0,acc)==acc */
            int t = 3 * (di + 1) + (dj + 1);
            acc = fmaf(tab[9 * key_at(keys, N, ii, jj) + t], u[IDX(ii, jj)], acc);
        }
    return acc;
}

/* plain 3x3 correlation with one table (FNet, HNet layers) */
static inline float stencil_one(const float *u, const float *w9, int N, int i, int j) {
    float acc = 0.0f;
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
            int ii = i + di, jj = j + dj;
            if (ii < 0 || ii >= N || jj < 0 || jj >= N) continue;
            acc = fmaf(w9[3 * (di + 1) + (dj + 1)], u[IDX(ii, jj)], acc);
        }
    return acc;
}

static inline int on_ring(int N, int i, int j) { return i == 0 || j == 0 || i == N - 1 || j == N - 1; }

/* u*idx + bval ; idx==NULL -> default square ring with zero boundary value (FEANet/geo.py:13-30) */
static inline float bc_apply(float v, const float *idx, const float *bval, int N, int i, int j) {
    if (!idx) return on_ring(N, i, j) ? 0.0f : v;
    float t = v * idx[IDX(i, j)];
    return t + (bval ? bval[IDX(i, j)] : 0.0f);
}
static inline float mask_apply(float v, const float *idx, int N, int i, int j) {
    if (!idx) return on_ring(N, i, j) ? 0.0f : v;
    return v * idx[IDX(i, j)];
}

void orc_stiffness_apply(const float *u, float *out, const uint8_t *keys, const float *ktab, int N, int B) {
    for (int b = 0; b < B; ++b) {
        const float *ub = u + (size_t)b * N * N;
        float *ob = out + (size_t)b * N * N;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) ob[IDX(i, j)] = stencil_src(ub, keys, ktab, N, i, j);
    }
}

/* out[b][c][i][j] = (key(i,j)==c) ? x[b][i][j] : 0 */
void orc_split_x(const float *x, float *out, const uint8_t *keys, int C, int N, int B) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j)
                    out[(((size_t)b * C + c) * N + i) * N + j] =
                        (key_at(keys, N, i, j) == c) ? x[(size_t)b * N * N + IDX(i, j)] : 0.0f;
}

/* bc_bstride: 0 = masks shared by all samples, N*N = per-sample masks */
void orc_reset_boundary(const float *u, float *out, const float *idx, const float *bval, size_t bc_bstride, int N,
                        int B) {
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
                out[(size_t)b * N * N + IDX(i, j)] =
                    bc_apply(u[(size_t)b * N * N + IDX(i, j)], idx ? idx + b * bc_bstride : NULL,
                             bval ? bval + b * bc_bstride : NULL, N, i, j);
}

/* one weighted-Jacobi sweep, single sample; tmp = scratch N*N */
static void jacobi_once(const float *u_in, float *u_out, float *tmp, const float *f, const uint8_t *keys,
                        const float *ktab, const float *invd, const float *idx, const float *bval, int N) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) tmp[IDX(i, j)] = bc_apply(u_in[IDX(i, j)], idx, bval, N, i, j);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            float ku = stencil_src(tmp, keys, ktab, N, i, j);
            float res = f[IDX(i, j)] - ku;
            float t = invd[key_at(keys, N, i, j)] * res;
            float un = t + tmp[IDX(i, j)];
            u_out[IDX(i, j)] = bc_apply(un, idx, bval, N, i, j);
        }
}

/* nsweeps of u <- BC(u); u <- BC(u + invd[key]*(f - K u)); invd[c] = fl32(omega)/d_c computed by the caller in fp32.
   u_in may alias u_out. */
void orc_jacobi(const float *u_in, float *u_out, const float *f, const uint8_t *keys, const float *ktab,
                const float *invd, const float *idx, const float *bval, size_t bc_bstride, int N, int B,
                int nsweeps) {
    size_t M = (size_t)N * N;
    float *tmp = (float *)malloc(M * sizeof(float));
    float *cur = (float *)malloc(M * sizeof(float));
    for (int b = 0; b < B; ++b) {
        const float *ib = idx ? idx + b * bc_bstride : NULL;
        const float *vb = bval ? bval + b * bc_bstride : NULL;
        memcpy(cur, u_in + b * M, M * sizeof(float));
        for (int s = 0; s < nsweeps; ++s) jacobi_once(cur, cur, tmp, f + b * M, keys, ktab, invd, ib, vb, N);
        memcpy(u_out + b * M, cur, M * sizeof(float));
    }
    free(tmp);
    free(cur);
}

/* nsweeps of the learned smoother: J = jacobi(u); x = J - u; h = (c3*g)o(c2*g)o(c1*g)(x); u = J + h
   hw = nlayers*9 weights. */
void orc_hjacobi(const float *u_in, float *u_out, const float *f, const uint8_t *keys, const float *ktab,
                 const float *invd, const float *idx, const float *bval, size_t bc_bstride, const float *hw,
                 int nlayers, int N, int B, int nsweeps) {
    size_t M = (size_t)N * N;
    float *tmp = (float *)malloc(M * sizeof(float));
    float *cur = (float *)malloc(M * sizeof(float));
    float *jac = (float *)malloc(M * sizeof(float));
    float *xa = (float *)malloc(M * sizeof(float));
    float *xb = (float *)malloc(M * sizeof(float));
    for (int b = 0; b < B; ++b) {
        const float *ib = idx ? idx + b * bc_bstride : NULL;
        const float *vb = bval ? bval + b * bc_bstride : NULL;
        memcpy(cur, u_in + b * M, M * sizeof(float));
        for (int s = 0; s < nsweeps; ++s) {
            jacobi_once(cur, jac, tmp, f + b * M, keys, ktab, invd, ib, vb, N);
            for (size_t k = 0; k < M; ++k) xa[k] = jac[k] - cur[k];
            for (int l = 0; l < nlayers; ++l) {
#pragma omp parallel for schedule(static)
                for (int i = 0; i < N; ++i)
                    for (int j = 0; j < N; ++j)
                        xb[IDX(i, j)] = mask_apply(stencil_one(xa, hw + 9 * l, N, i, j), ib, N, i, j);
                float *t = xa;
                xa = xb;
                xb = t;
            }
            for (size_t k = 0; k < M; ++k) cur[k] = jac[k] + xa[k];
        }
        memcpy(u_out + b * M, cur, M * sizeof(float));
    }
    free(tmp);
    free(cur);
    free(jac);
    free(xa);
    free(xb);
}

/* r = f - K u on the whole array (ring rows included, as the reference computes them) */
void orc_residual(const float *u, const float *f, float *r, const uint8_t *keys, const float *ktab, int N, int B) {
    for (int b = 0; b < B; ++b) {
        const float *ub = u + (size_t)b * N * N, *fb = f + (size_t)b * N * N;
        float *rb = r + (size_t)b * N * N;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) rb[IDX(i, j)] = fb[IDX(i, j)] - stencil_src(ub, keys, ktab, N, i, j);
    }
}

/* fc[I][J] = scale * sum_{a,b} R[key(src)][3a+b] * r[2I-1+a][2J-1+b], 1<=I,J<=Nc-2 ; ring = 0.
   keys = FINE-level keys (NULL: one table).  has_scale=0 skips the multiply (mg_test variant). */
void orc_restrict(const float *r, float *fc, const uint8_t *keys, const float *rtab, float scale, int has_scale,
                  int N, int B) {
    int Nc = (N - 1) / 2 + 1;
    for (int b = 0; b < B; ++b) {
        const float *rb = r + (size_t)b * N * N;
        float *cb = fc + (size_t)b * Nc * Nc;
#pragma omp parallel for schedule(static)
        for (int I = 0; I < Nc; ++I)
            for (int J = 0; J < Nc; ++J) {
                float v = 0.0f;
                if (I >= 1 && J >= 1 && I <= Nc - 2 && J <= Nc - 2) {
                    float acc = 0.0f;
                    for (int a = 0; a < 3; ++a)
                        for (int c = 0; c < 3; ++c) {
                            int ii = 2 * I - 1 + a, jj = 2 * J - 1 + c;
                            acc = fmaf(rtab[9 * key_at(keys, N, ii, jj) + 3 * a + c], rb[IDX(ii, jj)], acc);
                        }
                    v = has_scale ? scale * acc : acc;
                }
                cb[(size_t)I * Nc + J] = v;
            }
    }
}

/* variant A: u += BC_fine(bilinear_x2(vc)).  Weights are exactly 0, 1/2, 1, so only the ORDER of the additions
   matters, and ATen's CPU upsample_bilinear2d(align_corners=True) has two code paths (observed, torch 2.11):
     fine N <= 33 : scalar loop   e = ((1/4 a + 1/4 b) + 1/4 c) + 1/4 d          ("seq")
     fine N >= 65 : separable     e = 1/2 (1/2 a + 1/2 b) + 1/2 (1/2 c + 1/2 d)  ("hv": horizontal lerp, then vertical)
   They differ only at odd/odd nodes, by at most 1 ulp.  seq_order selects the path (callers pass N <= 33). */
void orc_prolong_bilinear(const float *vc, float *u, const float *idx, const float *bval, size_t bc_bstride,
                          int N, int B, int seq_order) {
    int Nc = (N - 1) / 2 + 1;
    for (int b = 0; b < B; ++b) {
        const float *cb = vc + (size_t)b * Nc * Nc;
        float *ub = u + (size_t)b * N * N;
        const float *ib = idx ? idx + b * bc_bstride : NULL;
        const float *vb = bval ? bval + b * bc_bstride : NULL;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                int I = i >> 1, J = j >> 1;
                float e;
                if ((i & 1) && (j & 1)) {
                    float a = cb[(size_t)I * Nc + J], bq = cb[(size_t)I * Nc + J + 1];
                    float c = cb[(size_t)(I + 1) * Nc + J], d = cb[(size_t)(I + 1) * Nc + J + 1];
                    if (seq_order) {
                        float t = 0.25f * a + 0.25f * bq;
                        t = t + 0.25f * c;
                        e = t + 0.25f * d;
                    } else {
                        float top = 0.5f * a + 0.5f * bq, bot = 0.5f * c + 0.5f * d;
                        e = 0.5f * top + 0.5f * bot;
                    }
                } else if (j & 1) {
                    e = 0.5f * cb[(size_t)I * Nc + J] + 0.5f * cb[(size_t)I * Nc + J + 1];
                } else if (i & 1) {
                    e = 0.5f * cb[(size_t)I * Nc + J] + 0.5f * cb[(size_t)(I + 1) * Nc + J];
                } else
                    e = cb[(size_t)I * Nc + J];
                e = bc_apply(e, ib, vb, N, i, j);
                ub[IDX(i, j)] = ub[IDX(i, j)] + e;
            }
    }
}

/* variant B: u += scale * convT(P[keyc])(vc).  e[y][x] = fma chain over the kernel taps (a asc, c asc) that hit a
   coarse node: I = (y+1-a)/2, J = (x+1-c)/2 (bit-exact with ATen ConvTranspose2d(stride 2, pad 1) as run here). */
void orc_prolong_table(const float *vc, float *u, const uint8_t *keys_c, const float *ptab, float scale,
                       int has_scale, int N, int B) {
    int Nc = (N - 1) / 2 + 1;
    for (int b = 0; b < B; ++b) {
        const float *cb = vc + (size_t)b * Nc * Nc;
        float *ub = u + (size_t)b * N * N;
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) {
                float acc = 0.0f;
                for (int a = 0; a < 3; ++a) {
                    if ((y + 1 - a) & 1) continue;
                    int I = (y + 1 - a) / 2;
                    if (I < 0 || I >= Nc) continue;
                    for (int c = 0; c < 3; ++c) {
                        if ((x + 1 - c) & 1) continue;
                        int J = (x + 1 - c) / 2;
                        if (J < 0 || J >= Nc) continue;
                        int k = keys_c ? keys_c[(size_t)I * Nc + J] : 0;
                        acc = fmaf(ptab[9 * k + 3 * a + c], cb[(size_t)I * Nc + J], acc);
                    }
                }
                float d = has_scale ? scale * acc : acc;
                ub[IDX(y, x)] = ub[IDX(y, x)] + d;
            }
    }
}

/* out[b] = sum over interior of (double)r^2 (sequential, row-major) */
void orc_sumsq_interior(const float *r, double *out, int N, int B) {
    for (int b = 0; b < B; ++b) {
        const float *rb = r + (size_t)b * N * N;
        double s = 0.0;
        for (int i = 1; i < N - 1; ++i)
            for (int j = 1; j < N - 1; ++j) s += (double)rb[IDX(i, j)] * (double)rb[IDX(i, j)];
        out[b] = s;
    }
}

/* FEANet/mesh.py:62-76 in closed form, integer arithmetic: phase of element (r,c), n = N-1 */
static inline int elem_phase(int shape, long n, long r, long c) {
    if (r < 0 || c < 0 || r >= n || c >= n) return 0;
    long a = 2 * c + 1 - n, b = 2 * r + 1 - n;
    if (shape == 0) return 4 * (a * a + b * b) < n * n;
    return (2 * labs(a) < n) && (2 * labs(b) < n);
}

/* pattern key per node (FEANet/mesh.py:23-26,78-101): [e1,e2,e3,e4] = phases of elements (i-1,j),(i-1,j-1),(i,j-1),(i,j) */
void orc_pattern_keys(uint8_t *keys, int N, int shape) {
    static const int ref[16][4] = {{0, 0, 0, 0}, {1, 1, 1, 1}, {0, 0, 0, 1}, {0, 0, 1, 0}, {1, 0, 0, 0}, {0, 1, 0, 0},
                                   {0, 0, 1, 1}, {1, 1, 0, 0}, {0, 1, 1, 0}, {1, 0, 0, 1}, {0, 1, 0, 1}, {1, 0, 1, 0},
                                   {1, 1, 1, 0}, {1, 1, 0, 1}, {0, 1, 1, 1}, {1, 0, 1, 1}};
    long n = N - 1;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            int k = 0;
            if (!on_ring(N, i, j)) {
                int e[4] = {elem_phase(shape, n, i - 1, j), elem_phase(shape, n, i - 1, j - 1),
                            elem_phase(shape, n, i, j - 1), elem_phase(shape, n, i, j)};
                for (int p = 0; p < 16; ++p)
                    if (ref[p][0] == e[0] && ref[p][1] == e[1] && ref[p][2] == e[2] && ref[p][3] == e[3]) k = p;
            }
            keys[IDX(i, j)] = (uint8_t)k;
        }
}

/* ---------------------------------------------------------------------------------------------------------
 * General per-ELEMENT conductivity (data model: `material`, one value per element, Data/dataset.py:71-104).
 * The reference only ships the 16-pattern two-phase operator; this is that operator with the pattern lookup replaced by
 * the element values themselves: the 3x3 kernel of a node is FEANet/mesh.py:103-117 `generate_kernel` evaluated (same
 * fp32 expression order) with a[pattern[i]] replaced by the conductivity of the node's element e_{i+1}, and K u keeps the
 * reference's form (FEANet/model.py:22-30): the weight of tap d is taken from the kernel of the SOURCE node i+d.  Each
 * tap's weight only involves the elements shared by the source node and the output node, i.e. the output node's own four
 * elements NW = E(i-1,j-1), NE = E(i-1,j), SW = E(i,j-1), SE = E(i,j) (E(r,c) = element with nodes (r,c) .. (r+1,c+1)):
 * on a two-phase map this is BIT-IDENTICAL to the pattern operator on every interior node (tests pin that).  Elements
 * outside the plate count as conductivity 0 (only the never-used boundary-ring outputs see them).
 * a: [n][n] fp32, n = N - 1, shared by the batch.  ke: the 4x4 element matrix of FEANet/mesh.py:28-31 in fp32.
 */
static inline float elem_at(const float *a, int n, int r, int c) {
    return (r < 0 || c < 0 || r >= n || c >= n) ? 0.0f : a[(size_t)r * n + c];
}
static inline void elem_weights(const float *a, const float *ke, int N, int i, int j, float *w) {
    const int n = N - 1;
    const float nw = elem_at(a, n, i - 1, j - 1), ne = elem_at(a, n, i - 1, j), sw = elem_at(a, n, i, j - 1),
                se = elem_at(a, n, i, j);
#define KE(r, c) ke[4 * (r) + (c)]
    w[0] = nw * KE(1, 3);
    w[1] = ne * KE(1, 2) + nw * KE(0, 3);
    w[2] = ne * KE(0, 2);
    w[3] = nw * KE(2, 3) + sw * KE(1, 0);
    w[4] = ((sw * KE(0, 0) + se * KE(1, 1)) + ne * KE(2, 2)) + nw * KE(3, 3);
    w[5] = ne * KE(3, 2) + se * KE(0, 1);
    w[6] = sw * KE(2, 0);
    w[7] = se * KE(2, 1) + sw * KE(3, 0);
    w[8] = se * KE(3, 1);
#undef KE
}
static inline float stencil_elem(const float *u, const float *a, const float *ke, int N, int i, int j) {
    float w[9], acc = 0.0f;
    elem_weights(a, ke, N, i, j, w);
    for (int di = -1; di <= 1; ++di)
        for (int dj = -1; dj <= 1; ++dj) {
            int ii = i + di, jj = j + dj;
            if (ii < 0 || ii >= N || jj < 0 || jj >= N) continue;
            acc = fmaf(w[3 * (di + 1) + (dj + 1)], u[IDX(ii, jj)], acc);
        }
    return acc;
}
void orc_elem_stiffness_apply(const float *u, float *out, const float *a, const float *ke, int N, int B) {
    for (int b = 0; b < B; ++b)
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) out[(size_t)b * N * N + IDX(i, j)] = stencil_elem(u + (size_t)b * N * N, a, ke, N, i, j);
}
void orc_elem_residual(const float *u, const float *f, float *r, const float *a, const float *ke, int N, int B) {
    for (int b = 0; b < B; ++b)
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
                r[(size_t)b * N * N + IDX(i, j)] =
                    f[(size_t)b * N * N + IDX(i, j)] - stencil_elem(u + (size_t)b * N * N, a, ke, N, i, j);
}
/* the node's Jacobi diagonal = centre entry of its own kernel (FEANet/jacobi.py:31-37) */
void orc_elem_diag(float *d, const float *a, const float *ke, int N) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            float w[9];
            elem_weights(a, ke, N, i, j, w);
            d[IDX(i, j)] = w[4];
        }
}
/* nsweeps of weighted Jacobi, FEANet/jacobi.py:39-47 with omega/d = fl(fl(1/d) * omega) per node (Tensor.__rtruediv__) */
void orc_elem_jacobi(const float *u_in, float *u_out, const float *f, const float *a, const float *ke, float omega,
                     const float *idx, const float *bval, size_t bc_bstride, int N, int B, int nsweeps) {
    size_t M = (size_t)N * N;
    float *tmp = (float *)malloc(M * sizeof(float));
    float *cur = (float *)malloc(M * sizeof(float));
    for (int b = 0; b < B; ++b) {
        const float *ib = idx ? idx + b * bc_bstride : NULL;
        const float *vb = bval ? bval + b * bc_bstride : NULL;
        const float *fb = f + b * M;
        memcpy(cur, u_in + b * M, M * sizeof(float));
        for (int s = 0; s < nsweeps; ++s) {
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) tmp[IDX(i, j)] = bc_apply(cur[IDX(i, j)], ib, vb, N, i, j);
#pragma omp parallel for schedule(static)
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) {
                    float w[9];
                    elem_weights(a, ke, N, i, j, w);
                    float ku = stencil_elem(tmp, a, ke, N, i, j);
                    float res = fb[IDX(i, j)] - ku;
                    float inv = (1.0f / w[4]) * omega;
                    float t = inv * res;
                    float un = t + tmp[IDX(i, j)];
                    cur[IDX(i, j)] = bc_apply(un, ib, vb, N, i, j);
                }
        }
        memcpy(u_out + b * M, cur, M * sizeof(float));
    }
    free(tmp);
    free(cur);
}

/* ---------------------------------------------------------------------------------------------------------
 * JacobiBlockPBC.jacobi_convolution (FEANet/jacobi.py:50-97), periodic boundary conditions, single pattern:
 *   u_pbc = circular pad of u[:-1,:-1] by (1,2) -> (n+3)^2;  residual = f_pad - Knet(u_pbc) on the padded array (Knet pads
 *   its mask with ones, FEANet/model.py:27-28), cropped [1:-1,1:-1];  u_new = omega/d * residual + reset_boundary(u)
 * i.e. a periodic 3x3 stencil on the n x n torus evaluated at all (n+1)^2 nodes (node n == node 0).
 * f_pad: [B][N+2][N+2] (the caller pads the load vector, as the reference requires); w9: the single 3x3 kernel.
 */
void orc_jacobi_pbc(const float *u_in, float *u_out, const float *f_pad, const float *w9, float invd, int N, int B,
                    int nsweeps) {
    const int n = N - 1, P = N + 2;
    size_t M = (size_t)N * N;
    float *cur = (float *)malloc(M * sizeof(float));
    float *nxt = (float *)malloc(M * sizeof(float));
    for (int b = 0; b < B; ++b) {
        memcpy(cur, u_in + b * M, M * sizeof(float));
        const float *fb = f_pad + (size_t)b * P * P;
        for (int s = 0; s < nsweeps; ++s) {
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) {
                    float acc = 0.0f;
                    for (int di = -1; di <= 1; ++di)
                        for (int dj = -1; dj <= 1; ++dj) {
                            int ii = ((i + di) % n + n) % n, jj = ((j + dj) % n + n) % n;
                            acc = fmaf(w9[3 * (di + 1) + (dj + 1)], cur[IDX(ii, jj)], acc);
                        }
                    float res = fb[(size_t)(i + 1) * P + (j + 1)] - acc;
                    float t = invd * res;
                    nxt[IDX(i, j)] = t + cur[IDX(i % n, j % n)];
                }
            float *tmp = cur;
            cur = nxt;
            nxt = tmp;
        }
        memcpy(u_out + b * M, cur, M * sizeof(float));
    }
    free(cur);
    free(nxt);
}
