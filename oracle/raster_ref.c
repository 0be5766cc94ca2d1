/*
 * oracle/raster_ref.c -- CPU restatement of the reference tile rasterizer.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the *checker* for fluidnexus_b200's CUDA rasterizer.  It is never linked,
 * imported or executed by the product path (fluidnexus_b200/); only tests/, bench.py's
 * cpu_baseline/reference legs and __graft_entry__.smoke() may use it.
 *
 * It restates, in plain scalar C, the algorithm of FluidNexus'
 * FluidDynamics/submodules/gaussian_rasterization_ch{1,3} (R3/ below; ch1 is the same code
 * with NUM_CHANNELS == 1):
 *   preprocess      R3/cuda_rasterizer/forward.cu:148-244  (+ computeCov3D :113-145, computeCov2D :70-108,
 *                   in_frustum/getRect/ndc2Pix R3/cuda_rasterizer/auxiliary.h:39-50,124-147)
 *   tile binning    R3/cuda_rasterizer/rasterizer_impl.cu:67-128 (keys = tile<<32 | float bits of depth,
 *                   stable sort, per-tile [start,end) ranges)
 *   blend forward   R3/cuda_rasterizer/forward.cu:249-373 (front-to-back, median depth, default 15)
 *   blend backward  R3/cuda_rasterizer/backward.cu:384-536
 *   cov2D backward  R3/cuda_rasterizer/backward.cu:137-263
 *   preprocess bwd  R3/cuda_rasterizer/backward.cu:332-381, computeCov3D bwd :267-327
 *
 * Parity pin: the reference ships no golden vectors.  This oracle is pinned
 * against the *compiled reference extension* run on a B200 (oracle/_ref, built by
 * oracle/build_ref.py) through the fixtures in tests/golden/ (tools/make_golden.py), and
 * against closed-form spot checks (tests/test_oracle_raster.py).
 *
 * Build: `make -C oracle` -> oracle/_build/liboracle_f32.so (REAL=float, follows the reference's
 * fp32 arithmetic op by op) and liboracle_f64.so (REAL=double twin used to set gradient
 * tolerances, since the reference's own atomics make its gradients run-to-run noisy).
 *
 * The SH colour path (forward.cu:20-67) is dead on the FluidNexus hot path (all three render
 * pipes pass colors_precomp) and is not restated: colours must be given.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif

#define TILE 16
#define MAXC 4

typedef REAL real;

static inline real r_exp(real x) { return sizeof(real) == 4 ? (real)expf((float)x) : (real)exp((double)x); }
static inline real r_sqrt(real x) { return sizeof(real) == 4 ? (real)sqrtf((float)x) : (real)sqrt((double)x); }
static inline real r_ceil(real x) { return sizeof(real) == 4 ? (real)ceilf((float)x) : (real)ceil((double)x); }
static inline real r_max(real a, real b) { return a > b ? a : b; }
static inline real r_min(real a, real b) { return a < b ? a : b; }
static inline int i_max(int a, int b) { return a > b ? a : b; }
static inline int i_min(int a, int b) { return a < b ? a : b; }

/* math-notation 3x3: m[r][c]; product summed k = 0,1,2 left to right like glm's operator* */
typedef struct { real m[3][3]; } mat3;
static mat3 m3_mul(mat3 A, mat3 B) {
    mat3 R;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            R.m[r][c] = A.m[r][0] * B.m[0][c] + A.m[r][1] * B.m[1][c] + A.m[r][2] * B.m[2][c];
    return R;
}
static mat3 m3_T(mat3 A) {
    mat3 R;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R.m[r][c] = A.m[c][r];
    return R;
}

typedef struct {
    int P, C, W, H, gx, gy;
    /* per Gaussian */
    real *depth, *xy, *cov3D, *conic_o;
    int *radii;
    uint32_t *tiles_touched;
    /* binning */
    int64_t R;
    uint32_t *point_list;
    uint32_t *ranges; /* 2 per tile */
    /* per pixel */
    real *final_T;
    uint32_t *n_contrib;
    /* copies of inputs needed by backward */
    real *means3D, *colors, *scales, *rots, *view, *proj, *bg;
    real scale_modifier, tan_fov_x, tan_fov_y;
    int has_cov_precomp;
} state_t;

/* auxiliary.h:39-41 -- note the reference evaluates this in double (1.0 literals) and returns float */
static real ndc2pix(real v, int S) {
    if (sizeof(real) == 4) return (real)(float)((((double)(float)v + 1.0) * S - 1.0) * 0.5);
    return ((v + 1.0) * S - 1.0) * 0.5;
}

/* auxiliary.h:43-50 */
static void get_rect(real px, real py, int max_radius, int gx, int gy, int *x0, int *y0, int *x1, int *y1) {
    *x0 = i_min(gx, i_max(0, (int)((px - max_radius) / TILE)));
    *y0 = i_min(gy, i_max(0, (int)((py - max_radius) / TILE)));
    *x1 = i_min(gx, i_max(0, (int)((px + max_radius + TILE - 1) / TILE)));
    *y1 = i_min(gy, i_max(0, (int)((py + max_radius + TILE - 1) / TILE)));
}

/* forward.cu:113-145 -- quaternion (r,x,y,z) is used as given, NOT normalised (:121) */
static void cov3d_from_scale_rot(const real *s, real mod, const real *q, real *cov) {
    real r = q[0], x = q[1], y = q[2], z = q[3];
    /* glm::mat3 R(col-major ctor) -> math matrix Rm[row][col] = ctor[col*3+row] */
    mat3 Rg; /* Rg.m[row][col] */
    real c0[3] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)};
    real c1[3] = {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)};
    real c2[3] = {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
    for (int i = 0; i < 3; i++) { Rg.m[i][0] = c0[i]; Rg.m[i][1] = c1[i]; Rg.m[i][2] = c2[i]; }
    mat3 S = {{{mod * s[0], 0, 0}, {0, mod * s[1], 0}, {0, 0, mod * s[2]}}};
    mat3 M = m3_mul(S, Rg);          /* M = S * R */
    mat3 Sig = m3_mul(m3_T(M), M);   /* Sigma = M^T M */
    /* Sigma[c][r] glm -> symmetric anyway */
    cov[0] = Sig.m[0][0]; cov[1] = Sig.m[1][0]; cov[2] = Sig.m[2][0];
    cov[3] = Sig.m[1][1]; cov[4] = Sig.m[2][1]; cov[5] = Sig.m[2][2];
}

/* shared by forward.cu:70-108 and backward.cu:159-193: builds T = W*J and the clamped t */
static void build_T(const real *mean, real fx, real fy, real tfx, real tfy, const real *V,
                    real *t_out, real *txtz_out, real *tytz_out, mat3 *Wm, mat3 *Tm) {
    real t[3] = {V[0] * mean[0] + V[4] * mean[1] + V[8] * mean[2] + V[12],
                 V[1] * mean[0] + V[5] * mean[1] + V[9] * mean[2] + V[13],
                 V[2] * mean[0] + V[6] * mean[1] + V[10] * mean[2] + V[14]};
    real limx = (real)1.3f * tfx, limy = (real)1.3f * tfy;
    real txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = r_min(limx, r_max(-limx, txtz)) * t[2];
    t[1] = r_min(limy, r_max(-limy, tytz)) * t[2];
    /* glm J ctor columns: (fx/tz,0,-(fx tx)/tz^2), (0,fy/tz,-(fy ty)/tz^2), 0 -> math J[row][col] */
    mat3 J = {{{fx / t[2], 0, 0}, {0, fy / t[2], 0}, {-(fx * t[0]) / (t[2] * t[2]), -(fy * t[1]) / (t[2] * t[2]), 0}}};
    /* glm W ctor columns: (v0,v4,v8),(v1,v5,v9),(v2,v6,v10) -> math W[row][col] */
    mat3 Wl = {{{V[0], V[1], V[2]}, {V[4], V[5], V[6]}, {V[8], V[9], V[10]}}};
    *Wm = Wl;
    *Tm = m3_mul(Wl, J);
    t_out[0] = t[0]; t_out[1] = t[1]; t_out[2] = t[2];
    *txtz_out = txtz; *tytz_out = tytz;
}

static mat3 vrk_of(const real *c) {
    mat3 Vrk = {{{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}}};
    return Vrk;
}

typedef struct { uint32_t tile; real depth; uint32_t idx; } inst_t;
static int inst_cmp(const void *a, const void *b) {
    const inst_t *x = a, *y = b;
    if (x->tile != y->tile) return x->tile < y->tile ? -1 : 1;
    if (x->depth != y->depth) return x->depth < y->depth ? -1 : 1; /* float bits of positive floats are monotone */
    return x->idx < y->idx ? -1 : (x->idx > y->idx); /* radix sort is stable; emission order is by idx */
}

static real *dup_f2r(const float *src, size_t n) {
    real *d = malloc(sizeof(real) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) d[i] = (real)src[i];
    return d;
}

void fnx_oracle_free(void *h) {
    state_t *s = h;
    if (!s) return;
    free(s->depth); free(s->xy); free(s->cov3D); free(s->conic_o); free(s->radii); free(s->tiles_touched);
    free(s->point_list); free(s->ranges); free(s->final_T); free(s->n_contrib);
    free(s->means3D); free(s->colors); free(s->scales); free(s->rots); free(s->view); free(s->proj); free(s->bg);
    free(s);
}

/* Rasterizer::forward, rasterizer_impl.cu:184-319.  Returns an opaque state for the backward. */
void *fnx_oracle_forward(int P, int C, int W, int H, const float *bg, const float *means3D, const float *colors,
                         const float *opacities, const float *scales, float scale_modifier, const float *rotations,
                         const float *cov3D_precomp, const float *view, const float *proj, float tan_fov_x,
                         float tan_fov_y, float *out_color, float *out_depth, int *out_radii, int64_t *num_rendered) {
    state_t *s = calloc(1, sizeof(state_t));
    s->P = P; s->C = C; s->W = W; s->H = H;
    s->gx = (W + TILE - 1) / TILE; s->gy = (H + TILE - 1) / TILE;
    s->scale_modifier = scale_modifier; s->tan_fov_x = tan_fov_x; s->tan_fov_y = tan_fov_y;
    s->means3D = dup_f2r(means3D, 3 * (size_t)P);
    s->colors = dup_f2r(colors, (size_t)C * P);
    s->has_cov_precomp = cov3D_precomp != NULL;
    s->scales = dup_f2r(scales, scales ? 3 * (size_t)P : 0);
    s->rots = dup_f2r(rotations, rotations ? 4 * (size_t)P : 0);
    s->view = dup_f2r(view, 16); s->proj = dup_f2r(proj, 16); s->bg = dup_f2r(bg, C);
    s->depth = calloc(P ? P : 1, sizeof(real)); s->xy = calloc(2 * (size_t)(P ? P : 1), sizeof(real));
    s->cov3D = calloc(6 * (size_t)(P ? P : 1), sizeof(real)); s->conic_o = calloc(4 * (size_t)(P ? P : 1), sizeof(real));
    s->radii = calloc(P ? P : 1, sizeof(int)); s->tiles_touched = calloc(P ? P : 1, sizeof(uint32_t));
    const real fy = H / (2.0f * (real)tan_fov_y), fx = W / (2.0f * (real)tan_fov_x); /* rasterizer_impl.cu:207-208 */
    const real *V = s->view, *PM = s->proj;

    /* ---- preprocess, forward.cu:174-243 ---- */
    for (int i = 0; i < P; i++) {
        const real *p = s->means3D + 3 * i;
        s->radii[i] = 0; s->tiles_touched[i] = 0;
        real hom[4] = {PM[0] * p[0] + PM[4] * p[1] + PM[8] * p[2] + PM[12], PM[1] * p[0] + PM[5] * p[1] + PM[9] * p[2] + PM[13],
                       PM[2] * p[0] + PM[6] * p[1] + PM[10] * p[2] + PM[14], PM[3] * p[0] + PM[7] * p[1] + PM[11] * p[2] + PM[15]};
        real pw = 1.0f / (hom[3] + (real)0.0000001f);
        real proj_x = hom[0] * pw, proj_y = hom[1] * pw;
        real pvz = V[2] * p[0] + V[6] * p[1] + V[10] * p[2] + V[14];
        if (pvz <= (real)0.2f) continue; /* auxiliary.h:138 */
        real *cov = s->cov3D + 6 * i;
        if (cov3D_precomp) for (int k = 0; k < 6; k++) cov[k] = (real)cov3D_precomp[6 * i + k];
        else cov3d_from_scale_rot(s->scales + 3 * i, scale_modifier, s->rots + 4 * i, cov);
        real t[3], txtz, tytz; mat3 Wm, Tm;
        build_T(p, fx, fy, tan_fov_x, tan_fov_y, V, t, &txtz, &tytz, &Wm, &Tm);
        /* glm: cov = transpose(T) * transpose(Vrk) * T with T(glm) = W(glm)*J(glm).  In math notation
           glm's product A*B is the ordinary product of the math matrices, and our W,J above are
           already the math matrices of glm's W,J, so Tm is math(T). */
        mat3 Vrk = vrk_of(cov);
        mat3 c2 = m3_mul(m3_mul(m3_T(Tm), m3_T(Vrk)), Tm);
        real ca = c2.m[0][0] + (real)0.3f, cb = c2.m[1][0], cc = c2.m[1][1] + (real)0.3f; /* glm cov[0][1] = col0,row1 */
        real det = ca * cc - cb * cb;
        if (det == 0.0f) continue;
        real det_inv = 1.f / det;
        real con[3] = {cc * det_inv, -cb * det_inv, ca * det_inv};
        real mid = 0.5f * (ca + cc);
        real l1 = mid + r_sqrt(r_max((real)0.1f, mid * mid - det));
        real l2 = mid - r_sqrt(r_max((real)0.1f, mid * mid - det));
        real my_radius = r_ceil(3.f * r_sqrt(r_max(l1, l2)));
        real px = ndc2pix(proj_x, W), py = ndc2pix(proj_y, H);
        int x0, y0, x1, y1;
        get_rect(px, py, (int)my_radius, s->gx, s->gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        s->depth[i] = pvz; s->radii[i] = (int)my_radius;
        s->xy[2 * i] = px; s->xy[2 * i + 1] = py;
        s->conic_o[4 * i] = con[0]; s->conic_o[4 * i + 1] = con[1]; s->conic_o[4 * i + 2] = con[2];
        s->conic_o[4 * i + 3] = (real)opacities[i];
        s->tiles_touched[i] = (uint32_t)((y1 - y0) * (x1 - x0));
    }
    /* ---- binning, rasterizer_impl.cu:259-299 ---- */
    int64_t R = 0;
    for (int i = 0; i < P; i++) R += s->tiles_touched[i];
    s->R = R; *num_rendered = R;
    inst_t *inst = malloc(sizeof(inst_t) * (size_t)(R ? R : 1));
    int64_t off = 0;
    for (int i = 0; i < P; i++) {
        if (s->radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        get_rect(s->xy[2 * i], s->xy[2 * i + 1], s->radii[i], s->gx, s->gy, &x0, &y0, &x1, &y1);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                inst[off].tile = (uint32_t)(y * s->gx + x); inst[off].depth = s->depth[i]; inst[off].idx = (uint32_t)i; off++;
            }
    }
    qsort(inst, (size_t)R, sizeof(inst_t), inst_cmp);
    int ntiles = s->gx * s->gy;
    s->point_list = malloc(sizeof(uint32_t) * (size_t)(R ? R : 1));
    s->ranges = calloc(2 * (size_t)ntiles, sizeof(uint32_t));
    for (int64_t k = 0; k < R; k++) {
        s->point_list[k] = inst[k].idx;
        if (k == 0) s->ranges[2 * inst[k].tile] = 0;
        else if (inst[k].tile != inst[k - 1].tile) { s->ranges[2 * inst[k - 1].tile + 1] = (uint32_t)k; s->ranges[2 * inst[k].tile] = (uint32_t)k; }
        if (k == R - 1) s->ranges[2 * inst[k].tile + 1] = (uint32_t)R;
    }
    free(inst);
    /* ---- blend, forward.cu:249-373 ---- */
    size_t HW = (size_t)H * W;
    s->final_T = malloc(sizeof(real) * HW); s->n_contrib = malloc(sizeof(uint32_t) * HW);
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            int tile = (py / TILE) * s->gx + (px / TILE);
            uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
            real T = 1.0f, Cacc[MAXC] = {0}, D = 15.0f;
            uint32_t contributor = 0, last_contributor = 0;
            for (uint32_t k = r0; k < r1; k++) {
                contributor++;
                uint32_t g = s->point_list[k];
                real dx = s->xy[2 * g] - (real)px, dy = s->xy[2 * g + 1] - (real)py;
                const real *co = s->conic_o + 4 * g;
                real power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                real alpha = r_min((real)0.99f, co[3] * r_exp(power));
                if (alpha < (real)(1.0f / 255.0f)) continue;
                real test_T = T * (1 - alpha);
                if (test_T < (real)0.0001f) break; /* done = true; this Gaussian is NOT blended */
                for (int ch = 0; ch < C; ch++) Cacc[ch] += s->colors[(size_t)g * C + ch] * alpha * T;
                if (T > 0.5f && test_T < 0.5) D = s->depth[g]; /* median depth :351-354 */
                T = test_T;
                last_contributor = contributor;
            }
            size_t pid = (size_t)py * W + px;
            s->final_T[pid] = T; s->n_contrib[pid] = last_contributor;
            for (int ch = 0; ch < C; ch++) out_color[ch * HW + pid] = (float)(Cacc[ch] + T * s->bg[ch]);
            out_depth[pid] = (float)D;
        }
    for (int i = 0; i < P; i++) out_radii[i] = s->radii[i];
    return s;
}

/* Rasterizer::backward, rasterizer_impl.cu:323-414.  All outputs must be zero-initialised by the caller
 * like rasterize_points.cu:150-158 does; dL_dconic is [P,4] (x,y,-,w), dL_dmeans2D is [P,3]. */
void fnx_oracle_backward(void *h, const float *dL_dpix_f, double *dL_dmeans2D, double *dL_dconic, double *dL_dopacity,
                         double *dL_dcolors, double *dL_dmeans3D, double *dL_dcov3D, double *dL_dscales, double *dL_drots) {
    state_t *s = h;
    const int P = s->P, C = s->C, W = s->W, H = s->H;
    size_t HW = (size_t)H * W;
    const real ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    /* accumulate in REAL like the device would (order differs from atomics; see header) */
    real *a_m2 = calloc(3 * (size_t)P + 1, sizeof(real)), *a_con = calloc(4 * (size_t)P + 1, sizeof(real));
    real *a_op = calloc((size_t)P + 1, sizeof(real)), *a_col = calloc((size_t)C * P + 1, sizeof(real));
    /* ---- BACKWARD::render, backward.cu:384-536 ---- */
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            int tile = (py / TILE) * s->gx + (px / TILE);
            uint32_t r0 = s->ranges[2 * tile], r1 = s->ranges[2 * tile + 1];
            size_t pid = (size_t)py * W + px;
            const real T_final = s->final_T[pid];
            real T = T_final;
            uint32_t contributor = r1 - r0;
            const uint32_t last_contributor = s->n_contrib[pid];
            real accum_rec[MAXC] = {0}, dpix[MAXC], last_color[MAXC] = {0}, last_alpha = 0;
            for (int ch = 0; ch < C; ch++) dpix[ch] = (real)dL_dpix_f[ch * HW + pid];
            for (uint32_t kk = 0; kk < r1 - r0; kk++) {
                uint32_t g = s->point_list[r1 - kk - 1];
                contributor--;
                if (contributor >= last_contributor) continue;
                real dx = s->xy[2 * g] - (real)px, dy = s->xy[2 * g + 1] - (real)py;
                const real *co = s->conic_o + 4 * g;
                real power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                real G = r_exp(power);
                real alpha = r_min((real)0.99f, co[3] * G);
                if (alpha < (real)(1.0f / 255.0f)) continue;
                T = T / (1.f - alpha);
                real dchannel_dcolor = alpha * T;
                real dL_dalpha = 0.0f;
                for (int ch = 0; ch < C; ch++) {
                    real c = s->colors[(size_t)g * C + ch];
                    accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                    last_color[ch] = c;
                    dL_dalpha += (c - accum_rec[ch]) * dpix[ch];
                    a_col[(size_t)g * C + ch] += dchannel_dcolor * dpix[ch];
                }
                dL_dalpha *= T;
                last_alpha = alpha;
                real bg_dot = 0;
                for (int ch = 0; ch < C; ch++) bg_dot += s->bg[ch] * dpix[ch];
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                real dL_dG = co[3] * dL_dalpha;
                real gdx = G * dx, gdy = G * dy;
                real dG_ddelx = -gdx * co[0] - gdy * co[1];
                real dG_ddely = -gdy * co[2] - gdx * co[1];
                a_m2[3 * g] += dL_dG * dG_ddelx * ddelx_dx;
                a_m2[3 * g + 1] += dL_dG * dG_ddely * ddely_dy;
                a_con[4 * g] += -0.5f * gdx * dx * dL_dG;
                a_con[4 * g + 1] += -0.5f * gdx * dy * dL_dG;
                a_con[4 * g + 3] += -0.5f * gdy * dy * dL_dG;
                a_op[g] += G * dL_dalpha;
            }
        }
    const real fy = H / (2.0f * s->tan_fov_y), fx = W / (2.0f * s->tan_fov_x);
    const real *V = s->view, *PM = s->proj;
    for (int i = 0; i < P; i++) {
        for (int k = 0; k < 3; k++) dL_dmeans2D[3 * i + k] = a_m2[3 * i + k];
        for (int k = 0; k < 4; k++) dL_dconic[4 * i + k] = a_con[4 * i + k];
        dL_dopacity[i] = a_op[i];
        for (int ch = 0; ch < C; ch++) dL_dcolors[(size_t)i * C + ch] = a_col[(size_t)i * C + ch];
        if (!(s->radii[i] > 0)) continue;
        /* ---- computeCov2DCUDA, backward.cu:137-263 ---- */
        const real *mean = s->means3D + 3 * i;
        const real *cov3D = s->cov3D + 6 * i;
        real dcon[3] = {a_con[4 * i], a_con[4 * i + 1], a_con[4 * i + 3]};
        real t[3], txtz, tytz; mat3 Wm, Tm;
        build_T(mean, fx, fy, s->tan_fov_x, s->tan_fov_y, V, t, &txtz, &tytz, &Wm, &Tm);
        const real limx = (real)1.3f * s->tan_fov_x, limy = (real)1.3f * s->tan_fov_y;
        const real x_grad_mul = (txtz < -limx || txtz > limx) ? 0 : 1;
        const real y_grad_mul = (tytz < -limy || tytz > limy) ? 0 : 1;
        mat3 Vrk = vrk_of(cov3D);
        mat3 c2 = m3_mul(m3_mul(m3_T(Tm), m3_T(Vrk)), Tm);
        real a = c2.m[0][0] + (real)0.3f, b = c2.m[1][0], c = c2.m[1][1] + (real)0.3f;
        real denom = a * c - b * b;
        real dL_da = 0, dL_db = 0, dL_dc = 0;
        real denom2inv = 1.0f / ((denom * denom) + (real)0.0000001f);
        /* glm T[i][j] = column i, row j = Tm.m[j][i] */
#define TG(i, j) Tm.m[j][i]
#define VG(i, j) Vrk.m[j][i]
#define WG(i, j) Wm.m[j][i]
        real dcov[6] = {0, 0, 0, 0, 0, 0};
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcon[0] + 2 * b * c * dcon[1] + (denom - a * c) * dcon[2]);
            dL_dc = denom2inv * (-a * a * dcon[2] + 2 * a * b * dcon[1] + (denom - a * c) * dcon[0]);
            dL_db = denom2inv * 2 * (b * c * dcon[0] - (denom + 2 * b * b) * dcon[1] + a * b * dcon[2]);
            dcov[0] = (TG(0, 0) * TG(0, 0) * dL_da + TG(0, 0) * TG(1, 0) * dL_db + TG(1, 0) * TG(1, 0) * dL_dc);
            dcov[3] = (TG(0, 1) * TG(0, 1) * dL_da + TG(0, 1) * TG(1, 1) * dL_db + TG(1, 1) * TG(1, 1) * dL_dc);
            dcov[5] = (TG(0, 2) * TG(0, 2) * dL_da + TG(0, 2) * TG(1, 2) * dL_db + TG(1, 2) * TG(1, 2) * dL_dc);
            dcov[1] = 2 * TG(0, 0) * TG(0, 1) * dL_da + (TG(0, 0) * TG(1, 1) + TG(0, 1) * TG(1, 0)) * dL_db + 2 * TG(1, 0) * TG(1, 1) * dL_dc;
            dcov[2] = 2 * TG(0, 0) * TG(0, 2) * dL_da + (TG(0, 0) * TG(1, 2) + TG(0, 2) * TG(1, 0)) * dL_db + 2 * TG(1, 0) * TG(1, 2) * dL_dc;
            dcov[4] = 2 * TG(0, 2) * TG(0, 1) * dL_da + (TG(0, 1) * TG(1, 2) + TG(0, 2) * TG(1, 1)) * dL_db + 2 * TG(1, 1) * TG(1, 2) * dL_dc;
        }
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = dcov[k];
        real dT00 = 2 * (TG(0, 0) * VG(0, 0) + TG(0, 1) * VG(0, 1) + TG(0, 2) * VG(0, 2)) * dL_da + (TG(1, 0) * VG(0, 0) + TG(1, 1) * VG(0, 1) + TG(1, 2) * VG(0, 2)) * dL_db;
        real dT01 = 2 * (TG(0, 0) * VG(1, 0) + TG(0, 1) * VG(1, 1) + TG(0, 2) * VG(1, 2)) * dL_da + (TG(1, 0) * VG(1, 0) + TG(1, 1) * VG(1, 1) + TG(1, 2) * VG(1, 2)) * dL_db;
        real dT02 = 2 * (TG(0, 0) * VG(2, 0) + TG(0, 1) * VG(2, 1) + TG(0, 2) * VG(2, 2)) * dL_da + (TG(1, 0) * VG(2, 0) + TG(1, 1) * VG(2, 1) + TG(1, 2) * VG(2, 2)) * dL_db;
        real dT10 = 2 * (TG(1, 0) * VG(0, 0) + TG(1, 1) * VG(0, 1) + TG(1, 2) * VG(0, 2)) * dL_dc + (TG(0, 0) * VG(0, 0) + TG(0, 1) * VG(0, 1) + TG(0, 2) * VG(0, 2)) * dL_db;
        real dT11 = 2 * (TG(1, 0) * VG(1, 0) + TG(1, 1) * VG(1, 1) + TG(1, 2) * VG(1, 2)) * dL_dc + (TG(0, 0) * VG(1, 0) + TG(0, 1) * VG(1, 1) + TG(0, 2) * VG(1, 2)) * dL_db;
        real dT12 = 2 * (TG(1, 0) * VG(2, 0) + TG(1, 1) * VG(2, 1) + TG(1, 2) * VG(2, 2)) * dL_dc + (TG(0, 0) * VG(2, 0) + TG(0, 1) * VG(2, 1) + TG(0, 2) * VG(2, 2)) * dL_db;
        real dJ00 = WG(0, 0) * dT00 + WG(0, 1) * dT01 + WG(0, 2) * dT02;
        real dJ02 = WG(2, 0) * dT00 + WG(2, 1) * dT01 + WG(2, 2) * dT02;
        real dJ11 = WG(1, 0) * dT10 + WG(1, 1) * dT11 + WG(1, 2) * dT12;
        real dJ12 = WG(2, 0) * dT10 + WG(2, 1) * dT11 + WG(2, 2) * dT12;
        real tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        real dtx = x_grad_mul * -fx * tz2 * dJ02;
        real dty = y_grad_mul * -fy * tz2 * dJ12;
        real dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
        /* transformVec4x3Transpose, auxiliary.h:79-86 */
        real dmean[3] = {V[0] * dtx + V[1] * dty + V[2] * dtz, V[4] * dtx + V[5] * dty + V[6] * dtz, V[8] * dtx + V[9] * dty + V[10] * dtz};
        /* ---- preprocessCUDA bwd, backward.cu:332-381 ---- */
        real mw_den = PM[3] * mean[0] + PM[7] * mean[1] + PM[11] * mean[2] + PM[15];
        real m_w = 1.0f / (mw_den + (real)0.0000001f);
        real mul1 = (PM[0] * mean[0] + PM[4] * mean[1] + PM[8] * mean[2] + PM[12]) * m_w * m_w;
        real mul2 = (PM[1] * mean[0] + PM[5] * mean[1] + PM[9] * mean[2] + PM[13]) * m_w * m_w;
        real g2x = a_m2[3 * i], g2y = a_m2[3 * i + 1];
        real dm2[3];
        dm2[0] = (PM[0] * m_w - PM[3] * mul1) * g2x + (PM[1] * m_w - PM[3] * mul2) * g2y;
        dm2[1] = (PM[4] * m_w - PM[7] * mul1) * g2x + (PM[5] * m_w - PM[7] * mul2) * g2y;
        dm2[2] = (PM[8] * m_w - PM[11] * mul1) * g2x + (PM[9] * m_w - PM[11] * mul2) * g2y;
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * i + k] = dmean[k] + dm2[k];
        if (!s->has_cov_precomp) {
            /* computeCov3D bwd, backward.cu:267-327 */
            const real *q = s->rots + 4 * i, *sc = s->scales + 3 * i;
            real r = q[0], x = q[1], y = q[2], z = q[3];
            real c0[3] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)};
            real c1[3] = {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)};
            real c2v[3] = {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
            mat3 Rg;
            for (int k = 0; k < 3; k++) { Rg.m[k][0] = c0[k]; Rg.m[k][1] = c1[k]; Rg.m[k][2] = c2v[k]; }
            real sv[3] = {s->scale_modifier * sc[0], s->scale_modifier * sc[1], s->scale_modifier * sc[2]};
            mat3 S = {{{sv[0], 0, 0}, {0, sv[1], 0}, {0, 0, sv[2]}}};
            mat3 M = m3_mul(S, Rg);
            mat3 dSig = {{{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]}, {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]}, {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}}};
            mat3 dM = m3_mul(M, dSig); /* dL_dM = 2 * M * dL_dSigma */
            for (int rr = 0; rr < 3; rr++) for (int cc2 = 0; cc2 < 3; cc2++) dM.m[rr][cc2] = 2.0f * dM.m[rr][cc2];
            mat3 Rt = m3_T(Rg), dMt = m3_T(dM);
            /* glm Rt[k] = column k of Rt; dot(Rt[k], dMt[k]) */
            for (int k = 0; k < 3; k++)
                dL_dscales[3 * i + k] = Rt.m[0][k] * dMt.m[0][k] + Rt.m[1][k] * dMt.m[1][k] + Rt.m[2][k] * dMt.m[2][k];
            /* dL_dMt[k] *= s[k]  (scales glm column k) */
            for (int k = 0; k < 3; k++) for (int rr = 0; rr < 3; rr++) dMt.m[rr][k] *= sv[k];
#define DG(i, j) dMt.m[j][i] /* glm dL_dMt[i][j] */
            dL_drots[4 * i + 0] = 2 * z * (DG(0, 1) - DG(1, 0)) + 2 * y * (DG(2, 0) - DG(0, 2)) + 2 * x * (DG(1, 2) - DG(2, 1));
            dL_drots[4 * i + 1] = 2 * y * (DG(1, 0) + DG(0, 1)) + 2 * z * (DG(2, 0) + DG(0, 2)) + 2 * r * (DG(1, 2) - DG(2, 1)) - 4 * x * (DG(2, 2) + DG(1, 1));
            dL_drots[4 * i + 2] = 2 * x * (DG(1, 0) + DG(0, 1)) + 2 * r * (DG(2, 0) - DG(0, 2)) + 2 * z * (DG(1, 2) + DG(2, 1)) - 4 * y * (DG(2, 2) + DG(0, 0));
            dL_drots[4 * i + 3] = 2 * r * (DG(0, 1) - DG(1, 0)) + 2 * x * (DG(2, 0) + DG(0, 2)) + 2 * y * (DG(1, 2) + DG(2, 1)) - 4 * z * (DG(1, 1) + DG(0, 0));
        }
    }
    free(a_m2); free(a_con); free(a_op); free(a_col);
}

/* introspection helpers for tests */
void fnx_oracle_get_geom(void *h, double *xy, double *depth, double *conic_o, uint32_t *tiles_touched) {
    state_t *s = h;
    for (int i = 0; i < s->P; i++) {
        xy[2 * i] = s->xy[2 * i]; xy[2 * i + 1] = s->xy[2 * i + 1]; depth[i] = s->depth[i];
        for (int k = 0; k < 4; k++) conic_o[4 * i + k] = s->conic_o[4 * i + k];
        tiles_touched[i] = s->tiles_touched[i];
    }
}
void fnx_oracle_get_image_state(void *h, double *final_T, uint32_t *n_contrib) {
    state_t *s = h;
    size_t HW = (size_t)s->H * s->W;
    for (size_t i = 0; i < HW; i++) { final_T[i] = s->final_T[i]; n_contrib[i] = s->n_contrib[i]; }
}

/* checkFrustum / markVisible, rasterizer_impl.cu:52-63 */
void fnx_oracle_mark_visible(int P, const float *means3D, const float *view, unsigned char *present) {
    for (int i = 0; i < P; i++) {
        const float *p = means3D + 3 * i;
        real z = (real)view[2] * p[0] + (real)view[6] * p[1] + (real)view[10] * p[2] + (real)view[14];
        present[i] = !(z <= (real)0.2f);
    }
}
