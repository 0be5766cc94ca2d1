// Host build of fluidnexus_b200/csrc/geom_grad.cuh (the chain rule of geom_bwd_kernel) for tests/test_geom_grad_host.py.
// TEST INFRASTRUCTURE: compiled with g++ by the test, never shipped, never loaded by the product (which has no CPU path).
// The loop below is the host twin of geom_bwd_kernel's per-Gaussian body (raster.cu) for one view.
#include <cstddef>

#include "geom_grad.cuh"

using namespace fnx::geomgrad;

extern "C" {

// dL_dmeans2D [P,3] (x, y used), dL_dconic [P,4] (x, y, -, w used) -> dL_dmeans3D [P,3], dL_dcov3D [P,6], dL_dscales [P,3],
// dL_drotations [P,4]; cov3D [P,6] is the covariance the forward used.
void fnx_host_geom_backward(int P, const float *means3D, const float *scales, float scale_modifier, const float *rotations,
                            const float *cov3D, const float *view, const float *proj, int W, int H, float tan_fov_x, float tan_fov_y,
                            const int *radii, const float *dL_dmeans2D, const float *dL_dconic, float *dL_dmeans3D, float *dL_dcov3D,
                            float *dL_dscales, float *dL_drotations) {
    const float fx = W / (2.0f * tan_fov_x), fy = H / (2.0f * tan_fov_y);
    for (int i = 0; i < P; i++) {
        float d_mean[3] = {0, 0, 0}, d_cov[6] = {0, 0, 0, 0, 0, 0}, ds[3] = {0, 0, 0}, dq[4] = {0, 0, 0, 0};
        if (radii[i] > 0) {
            float S[3][3], A[2][3], t[3], dA[2][3];
            bool x_free, y_free;
            sym3_from6(cov3D + 6 * (size_t)i, S);
            view_jacobian(means3D + 3 * (size_t)i, view, fx, fy, tan_fov_x, tan_fov_y, A, t, x_free, y_free);
            screen_cov_backward(A, S, dL_dconic[4 * (size_t)i], dL_dconic[4 * (size_t)i + 1], dL_dconic[4 * (size_t)i + 3], d_cov, dA);
            perspective_backward(dA, view, t, fx, fy, x_free, y_free, d_mean);
            ndc_backward(proj, means3D + 3 * (size_t)i, dL_dmeans2D[3 * (size_t)i], dL_dmeans2D[3 * (size_t)i + 1], d_mean);
            const float s[3] = {scale_modifier * scales[3 * i], scale_modifier * scales[3 * i + 1], scale_modifier * scales[3 * i + 2]};
            cov3d_backward(s, rotations + 4 * (size_t)i, d_cov, ds, dq);
        }
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * (size_t)i + k] = d_mean[k];
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * (size_t)i + k] = d_cov[k];
        for (int k = 0; k < 3; k++) dL_dscales[3 * (size_t)i + k] = ds[k];
        for (int k = 0; k < 4; k++) dL_drotations[4 * (size_t)i + k] = dq[k];
    }
}

// pieces, for the finite-difference tests
void fnx_host_screen_cov_grad(float a, float b, float c, const float *g, float *D) {
    const Sym2 d = screen_cov_grad(a, b, c, g[0], g[1], g[2]);
    D[0] = d.xx; D[1] = d.xy; D[2] = d.yy;
}
void fnx_host_cov3d_backward(const float *s, const float *q, const float *dcov6, float *ds, float *dq) { cov3d_backward(s, q, dcov6, ds, dq); }
void fnx_host_quat_rotation(const float *q, float *Q9) {
    float Q[3][3];
    quat_rotation(q, Q);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) Q9[3 * a + b] = Q[a][b];
}
}
