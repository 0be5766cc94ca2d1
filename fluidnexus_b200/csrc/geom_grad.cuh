// geom_grad.cuh -- closed-form backward of the per-Gaussian geometry (screen covariance -> 3-D covariance / view-space mean,
// screen mean -> mean, 3-D covariance -> scale / quaternion), derived in matrix form and written over plain float arrays.
//
// What it replaces in the reference: computeCov2DCUDA (R3/cuda_rasterizer/backward.cu:137-263), the projection part of
// preprocessCUDA's backward (:332-381) and computeCov3D's backward (:267-327).  The reference spells those chains out entry by
// entry in glm's column-major indices; here every step is the textbook matrix identity it comes from, so the code can be read
// against the derivation in the comments.  Same function, different (slightly more accurate) roundings: held to the fp64 oracle
// at rel-L2 1e-4 on the GPU (tests/test_raster_gpu.py, tests/test_baseline_sizes_gpu.py) and, because nothing in this header
// needs a GPU, compiled for the HOST by tests/test_geom_grad_host.py (g++) and checked against the oracle and finite differences.
//
// Notation (math indices, row then column):
//   V      view matrix as the reference stores it (16 floats, V[4*j + i] = row i, column j of the 4x4); R(i,j) = V[4*j+i], i,j < 3
//   t      view-space mean R m + V[12..14], with x/z and y/z clamped to +-1.3 tan(fov/2) (forward.cu:78-83)
//   J      2x3 perspective Jacobian at t:  [[fx/z, 0, -fx x/z^2], [0, fy/z, -fy y/z^2]]
//   A      J R (2x3)
//   S      3-D covariance, symmetric, stored as c6 = (xx, xy, xz, yy, yz, zz)
//   S2     A S A^T + 0.3 I = [[a, b], [b, c]]      (screen covariance, forward.cu:99-108)
//   K      S2^-1, the "conic"; the blend evaluates power = -1/2 d^T K d
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FNX_HD __host__ __device__ __forceinline__
#else
#define FNX_HD inline
#endif

namespace fnx {
namespace geomgrad {

struct Sym2 {  // symmetric 2x2
    float xx, xy, yy;
};

// View-space mean, its clamp and A = J R.  x_free / y_free: the coordinate was not clamped (a clamped one carries no gradient,
// backward.cu:171-176).
FNX_HD void view_jacobian(const float *m, const float *V, float fx, float fy, float tan_fov_x, float tan_fov_y, float A[2][3],
                          float t[3], bool &x_free, bool &y_free) {
    for (int i = 0; i < 3; i++) t[i] = V[i] * m[0] + V[4 + i] * m[1] + V[8 + i] * m[2] + V[12 + i];
    const float limx = 1.3f * tan_fov_x, limy = 1.3f * tan_fov_y;
    const float rx = t[0] / t[2], ry = t[1] / t[2];
    x_free = !(rx < -limx || rx > limx);
    y_free = !(ry < -limy || ry > limy);
    t[0] = fminf(limx, fmaxf(-limx, rx)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, ry)) * t[2];
    const float iz = 1.0f / t[2];
    const float j00 = fx * iz, j02 = -fx * t[0] * iz * iz;
    const float j11 = fy * iz, j12 = -fy * t[1] * iz * iz;
    for (int k = 0; k < 3; k++) {  // A(i,k) = sum_c J(i,c) R(c,k),  R(c,k) = V[4k + c]
        A[0][k] = j00 * V[4 * k] + j02 * V[4 * k + 2];
        A[1][k] = j11 * V[4 * k + 1] + j12 * V[4 * k + 2];
    }
}

// dL/dS2 from F = dL/dK, K = S2^-1.
//   g = (F_xx, F_xy, F_yy): the blend backward accumulates F = -1/2 (dL/dpower) d d^T entry by entry, so g_xy is ONE of the two
//   equal off-diagonal entries of F.
//   d(S2^-1) = -K dS2 K  =>  dL/dS2 = -K F K = -(adj F adj) / det^2,   adj = [[c, -b], [-b, a]].
//   The reference regularises 1/det^2 as 1/(det^2 + 1e-7) and drops the whole term when that is 0 (backward.cu:206-210); kept.
// Returned as the symmetric matrix D: dL/da = D.xx, dL/dc = D.yy, and b (which sits in two entries of S2) gets 2 D.xy.
FNX_HD Sym2 screen_cov_grad(float a, float b, float c, float g_xx, float g_xy, float g_yy) {
    const float det = a * c - b * b;
    const float w = 1.0f / (det * det + 0.0000001f);
    Sym2 D = {0.f, 0.f, 0.f};
    if (w != 0.f) {
        // H = adj F
        const float h00 = c * g_xx - b * g_xy, h01 = c * g_xy - b * g_yy;
        const float h10 = a * g_xy - b * g_xx, h11 = a * g_yy - b * g_xy;
        // D = -w H adj
        D.xx = -w * (c * h00 - b * h01);
        D.xy = -w * (a * h01 - b * h00);
        D.yy = -w * (a * h11 - b * h10);
    }
    return D;
}

FNX_HD void sym3_from6(const float *c6, float S[3][3]) {
    S[0][0] = c6[0]; S[0][1] = S[1][0] = c6[1]; S[0][2] = S[2][0] = c6[2];
    S[1][1] = c6[3]; S[1][2] = S[2][1] = c6[4]; S[2][2] = c6[5];
}

// Backward of S2 = A S A^T + 0.3 I given g = dL/dK (see screen_cov_grad).
//   dL/dS = A^T D A   -- ADDED to dcov6, off-diagonal entries doubled (c6 holds each of them once);
//   dL/dA = 2 D A S   -- written to dA.
FNX_HD void screen_cov_backward(const float A[2][3], const float S[3][3], float g_xx, float g_xy, float g_yy, float dcov6[6],
                                float dA[2][3]) {
    float AS[2][3];
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 3; k++) AS[i][k] = A[i][0] * S[0][k] + A[i][1] * S[1][k] + A[i][2] * S[2][k];
    const float a = AS[0][0] * A[0][0] + AS[0][1] * A[0][1] + AS[0][2] * A[0][2] + 0.3f;
    const float b = AS[0][0] * A[1][0] + AS[0][1] * A[1][1] + AS[0][2] * A[1][2];
    const float c = AS[1][0] * A[1][0] + AS[1][1] * A[1][1] + AS[1][2] * A[1][2] + 0.3f;
    const Sym2 D = screen_cov_grad(a, b, c, g_xx, g_xy, g_yy);
    float DA[2][3];
    for (int k = 0; k < 3; k++) {
        DA[0][k] = D.xx * A[0][k] + D.xy * A[1][k];
        DA[1][k] = D.xy * A[0][k] + D.yy * A[1][k];
        dA[0][k] = 2.0f * (D.xx * AS[0][k] + D.xy * AS[1][k]);
        dA[1][k] = 2.0f * (D.xy * AS[0][k] + D.yy * AS[1][k]);
    }
    // E = A^T (D A), symmetric
    float E[3][3];
    for (int k = 0; k < 3; k++)
        for (int l = k; l < 3; l++) E[k][l] = A[0][k] * DA[0][l] + A[1][k] * DA[1][l];
    dcov6[0] += E[0][0];
    dcov6[1] += 2.0f * E[0][1];
    dcov6[2] += 2.0f * E[0][2];
    dcov6[3] += E[1][1];
    dcov6[4] += 2.0f * E[1][2];
    dcov6[5] += E[2][2];
}

// Backward of A = J(t) R and t = R m + const, ADDED to dmean.
//   dL/dJ = dL/dA R^T; J's four non-constant entries: J00 = fx/z, J11 = fy/z, J02 = -fx x/z^2, J12 = -fy y/z^2, so with
//   kx = fx/z^2, ky = fy/z^2:  dJ02/dx = -kx, dJ12/dy = -ky, dJ00/dz = -kx, dJ11/dz = -ky, dJ02/dz = 2 kx x/z, dJ12/dz = 2 ky y/z;
//   dL/dm = R^T dL/dt.
FNX_HD void perspective_backward(const float dA[2][3], const float *V, const float t[3], float fx, float fy, bool x_free, bool y_free,
                                 float dmean[3]) {
    // dJ(i,c) = sum_k dA(i,k) R(c,k)
    const float dJ00 = dA[0][0] * V[0] + dA[0][1] * V[4] + dA[0][2] * V[8];
    const float dJ02 = dA[0][0] * V[2] + dA[0][1] * V[6] + dA[0][2] * V[10];
    const float dJ11 = dA[1][0] * V[1] + dA[1][1] * V[5] + dA[1][2] * V[9];
    const float dJ12 = dA[1][0] * V[2] + dA[1][1] * V[6] + dA[1][2] * V[10];
    const float iz = 1.0f / t[2];
    const float kx = fx * iz * iz, ky = fy * iz * iz;
    float dt[3];
    dt[0] = x_free ? -kx * dJ02 : 0.f;
    dt[1] = y_free ? -ky * dJ12 : 0.f;
    dt[2] = 2.0f * iz * (kx * t[0] * dJ02 + ky * t[1] * dJ12) - (kx * dJ00 + ky * dJ11);
    for (int j = 0; j < 3; j++) dmean[j] += V[4 * j] * dt[0] + V[4 * j + 1] * dt[1] + V[4 * j + 2] * dt[2];
}

// Backward of the screen mean, ADDED to dmean.  ndc = (P m).xy / w', w' = (P m).w + 1e-7 (forward.cu:186-188); (g_x, g_y) is the
// gradient w.r.t. ndc (the blend backward has already applied the ndc -> pixel factors W/2, H/2).
//   d ndc_x / d m_j = (P(0,j) - ndc_x P(3,j)) / w',   P(i,j) = Pm[4j + i].
FNX_HD void ndc_backward(const float *Pm, const float *m, float g_x, float g_y, float dmean[3]) {
    float h[4];
    for (int i = 0; i < 4; i++) h[i] = Pm[i] * m[0] + Pm[4 + i] * m[1] + Pm[8 + i] * m[2] + Pm[12 + i];
    const float iw = 1.0f / (h[3] + 0.0000001f);
    const float nx = h[0] * iw, ny = h[1] * iw;
    for (int j = 0; j < 3; j++)
        dmean[j] += iw * ((Pm[4 * j] - nx * Pm[4 * j + 3]) * g_x + (Pm[4 * j + 1] - ny * Pm[4 * j + 3]) * g_y);
}

// Rotation of the quaternion q = (r, x, y, z), used as given, NOT normalised (forward.cu:121):
//   Q = I + 2 r [v]x + 2 [v]x^2,   v = (x, y, z),  [v]x^2 = v v^T - |v|^2 I.
FNX_HD void quat_rotation(const float *q, float Q[3][3]) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    Q[0][0] = 1.f - 2.f * (y * y + z * z); Q[0][1] = 2.f * (x * y - r * z);       Q[0][2] = 2.f * (x * z + r * y);
    Q[1][0] = 2.f * (x * y + r * z);       Q[1][1] = 1.f - 2.f * (x * x + z * z); Q[1][2] = 2.f * (y * z - r * x);
    Q[2][0] = 2.f * (x * z - r * y);       Q[2][1] = 2.f * (y * z + r * x);       Q[2][2] = 1.f - 2.f * (x * x + y * y);
}

// Backward of S = M^T M, M = diag(s) Q^T (forward.cu:113-145), given dcov6 = dL/dc6 (off-diagonal entries count both copies).
//   E = dL/dS as a full symmetric matrix (off-diagonals halved);  N = dL/dM = 2 M E = 2 diag(s) U,  U = Q^T E;
//   ds_k = sum_a N(k,a) Q(a,k) = 2 s_k sum_a U(k,a) Q(a,k);
//   G = dL/dQ,  G(a,k) = s_k N(k,a) = 2 s_k^2 U(k,a);
//   with ax(G) = (G21 - G12, G02 - G20, G10 - G01):
//       dL/dr = 2 v . ax(G)                              (only the 2 r [v]x term holds r, and sum_ab G_ab [v]x_ab = v . ax(G))
//       dL/dv = 2 r ax(G) + 2 (G + G^T) v - 4 tr(G) v    (the last two from v v^T - |v|^2 I)
// s is the scale the covariance was built from (scale_modifier * scale); like the reference (backward.cu:296-298) the result is
// the gradient w.r.t. THAT, not multiplied by scale_modifier again.
FNX_HD void cov3d_backward(const float s[3], const float *q, const float dcov6[6], float ds[3], float dq[4]) {
    float Q[3][3], E[3][3], G[3][3];
    quat_rotation(q, Q);
    E[0][0] = dcov6[0]; E[1][1] = dcov6[3]; E[2][2] = dcov6[5];
    E[0][1] = E[1][0] = 0.5f * dcov6[1];
    E[0][2] = E[2][0] = 0.5f * dcov6[2];
    E[1][2] = E[2][1] = 0.5f * dcov6[4];
    for (int k = 0; k < 3; k++) {
        float U[3];
        for (int a = 0; a < 3; a++) U[a] = Q[0][k] * E[0][a] + Q[1][k] * E[1][a] + Q[2][k] * E[2][a];
        ds[k] = 2.0f * s[k] * (U[0] * Q[0][k] + U[1] * Q[1][k] + U[2] * Q[2][k]);
        const float w = 2.0f * s[k] * s[k];
        for (int a = 0; a < 3; a++) G[a][k] = w * U[a];
    }
    const float r = q[0], v[3] = {q[1], q[2], q[3]};
    const float ax[3] = {G[2][1] - G[1][2], G[0][2] - G[2][0], G[1][0] - G[0][1]};
    const float tr = G[0][0] + G[1][1] + G[2][2];
    dq[0] = 2.0f * (v[0] * ax[0] + v[1] * ax[1] + v[2] * ax[2]);
    for (int c = 0; c < 3; c++) {
        const float sym = (G[c][0] + G[0][c]) * v[0] + (G[c][1] + G[1][c]) * v[1] + (G[c][2] + G[2][c]) * v[2];
        dq[1 + c] = 2.0f * r * ax[c] + 2.0f * sym - 4.0f * tr * v[c];
    }
}

}  // namespace geomgrad
}  // namespace fnx
