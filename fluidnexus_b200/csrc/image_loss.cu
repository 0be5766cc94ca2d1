// image_loss.cu -- fused L1 + SSIM image loss (value and gradient) for sm_100a.
//
// Replaces, per view and per step, the reference's torch graph
//   l1_loss                      FD/utils/loss_utils.py:9-10
//   ssim / _ssim                 FD/utils/loss_utils.py:21-64 (window rebuilt on the CPU and uploaded every call :35-39,
//                                five grouped 11x11 conv2d forward + their backward)
//   grey conversion              FD/entries_fluid_nexus/train_physical_particle.py:356-360
//   weighting                    FD/entries_scalar_real/train_physical_particle.py:346-347
// with two kernels: (1) per 16x16 tile with a 5-pixel halo, separable 11-tap Gaussian of {x, y, x^2, y^2, xy}, the SSIM
// map, its three partial-derivative maps and the per-view loss sums; (2) separable Gaussian of the derivative maps ->
// dL/dimage, with the L1 sign term added.  Zero padding like F.conv2d(padding=5).
#include "common.cuh"

namespace fnx {

constexpr int LT = 16;           // tile edge
constexpr int HALO = 5;          // window 11
constexpr int LE = LT + 2 * HALO;  // 26

struct Win {
    float g[11];
};

static Win make_window() {
    // loss_utils.py:21-23: float32 tensor of exp(-(x-5)^2 / (2*1.5^2)), normalised
    Win w;
    float s = 0.f;
    for (int k = 0; k < 11; k++) {
        w.g[k] = (float)exp(-((double)(k - 5) * (k - 5)) / (2.0 * 1.5 * 1.5));
        s += w.g[k];
    }
    for (int k = 0; k < 11; k++) w.g[k] /= s;
    return w;
}

__device__ __forceinline__ float load_px(const float *__restrict__ img, int C, int H, int W, int v, int c, bool grey, int y,
                                         int x) {
    if (x < 0 || y < 0 || x >= W || y >= H) return 0.f;
    const size_t HW = (size_t)H * W;
    if (!grey) return img[((size_t)v * C + c) * HW + (size_t)y * W + x];
    float s = 0.f;
    for (int k = 0; k < C; k++) s += img[((size_t)v * C + k) * HW + (size_t)y * W + x];
    return s / C;  // torch.mean over the channel dim
}

// pass 1: SSIM map + derivative maps + loss sums.  grid (tiles_x, tiles_y, V*Ce), Ce = grey ? 1 : C
__global__ void __launch_bounds__(LT *LT)
ssim_fwd_kernel(int V, int C, int Ce, int H, int W, bool grey, const float *__restrict__ img, const float *__restrict__ gt,
                Win win, float *__restrict__ maps /*[V*Ce][3][H][W]*/, float *__restrict__ l1_sum, float *__restrict__ ssim_sum) {
    __shared__ float sx[LE][LE + 1], sy[LE][LE + 1];
    __shared__ float hq[5][LE][LT + 1];
    __shared__ float red[2][LT * LT / 32];
    const int vc = blockIdx.z, v = vc / Ce, c = vc % Ce;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const int tid = threadIdx.y * LT + threadIdx.x;
    for (int k = tid; k < LE * LE; k += LT * LT) {
        const int ly = k / LE, lx = k % LE;
        sx[ly][lx] = load_px(img, C, H, W, v, c, grey, y0 + ly - HALO, x0 + lx - HALO);
        sy[ly][lx] = load_px(gt, C, H, W, v, c, grey, y0 + ly - HALO, x0 + lx - HALO);
    }
    __syncthreads();
    // horizontal 11-tap of the five products
    for (int k = tid; k < LE * LT; k += LT * LT) {
        const int ly = k / LT, lx = k % LT;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++) {
            const float x = sx[ly][lx + t], y = sy[ly][lx + t], w = win.g[t];
            a0 += w * x; a1 += w * y; a2 += w * x * x; a3 += w * y * y; a4 += w * x * y;
        }
        hq[0][ly][lx] = a0; hq[1][ly][lx] = a1; hq[2][ly][lx] = a2; hq[3][ly][lx] = a3; hq[4][ly][lx] = a4;
    }
    __syncthreads();
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int px = x0 + lx, py = y0 + ly;
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int t = 0; t < 11; t++) {
        const float w = win.g[t];
        mu1 += w * hq[0][ly + t][lx]; mu2 += w * hq[1][ly + t][lx];
        e11 += w * hq[2][ly + t][lx]; e22 += w * hq[3][ly + t][lx]; e12 += w * hq[4][ly + t][lx];
    }
    float l1 = 0.f, ss = 0.f;
    if (px < W && py < H) {
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
        const float sigma1_sq = e11 - mu1_sq, sigma2_sq = e22 - mu2_sq, sigma12 = e12 - mu1_mu2;
        const float A1 = 2.f * mu1_mu2 + C1, A2 = 2.f * sigma12 + C2;
        const float B1 = mu1_sq + mu2_sq + C1, B2 = sigma1_sq + sigma2_sq + C2;
        const float S = (A1 * A2) / (B1 * B2);
        ss = S;
        l1 = fabsf(sx[ly + HALO][lx + HALO] - sy[ly + HALO][lx + HALO]);
        // partial derivatives of S w.r.t. mu1, E[x^2], E[xy] (the window-averaged quantities that depend on x)
        const float dS_dmu1 = (2.f * mu2 * (A2 - A1)) / (B1 * B2) - S * (2.f * mu1 / B1 - 2.f * mu1 / B2);
        const float dS_de11 = -S / B2;
        const float dS_de12 = 2.f * A1 / (B1 * B2);
        const size_t HW = (size_t)H * W, o = (size_t)py * W + px;
        float *m = maps + (size_t)vc * 3 * HW;
        m[o] = dS_dmu1; m[HW + o] = dS_de11; m[2 * HW + o] = dS_de12;
    }
    l1 = warp_sum(l1);
    ss = warp_sum(ss);
    if ((tid & 31) == 0) { red[0][tid >> 5] = l1; red[1][tid >> 5] = ss; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < LT * LT / 32; k++) { a += red[0][k]; b += red[1][k]; }
        atomicAdd(&l1_sum[v], a);
        atomicAdd(&ssim_sum[v], b);
    }
}

// pass 2: dL/dimg = w_l1/N * sign(x-y) - w_ssim/N * (G*M1 + 2x G*M2 + y G*M3)
__global__ void __launch_bounds__(LT *LT)
ssim_bwd_kernel(int V, int C, int Ce, int H, int W, bool grey, const float *__restrict__ img, const float *__restrict__ gt,
                Win win, const float *__restrict__ maps, float w_l1, float w_ssim, float *__restrict__ dL_dimg) {
    __shared__ float sm[3][LE][LE + 1];
    __shared__ float hq[3][LE][LT + 1];
    const int vc = blockIdx.z, v = vc / Ce, c = vc % Ce;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const int tid = threadIdx.y * LT + threadIdx.x;
    const size_t HW = (size_t)H * W;
    const float *m = maps + (size_t)vc * 3 * HW;
    for (int k = tid; k < LE * LE; k += LT * LT) {
        const int ly = k / LE, lx = k % LE;
        const int gx = x0 + lx - HALO, gy = y0 + ly - HALO;
        const bool in = gx >= 0 && gy >= 0 && gx < W && gy < H;
        const size_t o = (size_t)gy * W + gx;
        sm[0][ly][lx] = in ? m[o] : 0.f;
        sm[1][ly][lx] = in ? m[HW + o] : 0.f;
        sm[2][ly][lx] = in ? m[2 * HW + o] : 0.f;
    }
    __syncthreads();
    for (int k = tid; k < LE * LT; k += LT * LT) {
        const int ly = k / LT, lx = k % LT;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++) {
            const float w = win.g[t];
            a0 += w * sm[0][ly][lx + t]; a1 += w * sm[1][ly][lx + t]; a2 += w * sm[2][ly][lx + t];
        }
        hq[0][ly][lx] = a0; hq[1][ly][lx] = a1; hq[2][ly][lx] = a2;
    }
    __syncthreads();
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int px = x0 + lx, py = y0 + ly;
    if (px >= W || py >= H) return;
    float g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
    for (int t = 0; t < 11; t++) {
        const float w = win.g[t];
        g1 += w * hq[0][ly + t][lx]; g2 += w * hq[1][ly + t][lx]; g3 += w * hq[2][ly + t][lx];
    }
    const float x = load_px(img, C, H, W, v, c, grey, py, px), y = load_px(gt, C, H, W, v, c, grey, py, px);
    const float n = (float)Ce * (float)H * (float)W;  // elements the reference's .mean() runs over (per view)
    const float d = x - y;
    const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
    // in grey mode the reference's 3 identical channels each carry 1/3 of the mean, and d grey / d channel = 1/C
    float gout = (w_l1 * sgn - w_ssim * (g1 + 2.f * x * g2 + y * g3)) / n;
    const size_t o = (size_t)py * W + px;
    if (!grey) {
        dL_dimg[((size_t)v * C + c) * HW + o] = gout;
    } else {
        gout /= C;
        for (int k = 0; k < C; k++) dL_dimg[((size_t)v * C + k) * HW + o] = gout;
    }
}

__global__ void scale_means_kernel(int V, float inv_n, float *a, float *b) {
    for (int i = threadIdx.x; i < V; i += blockDim.x) { a[i] *= inv_n; b[i] *= inv_n; }
}

}  // namespace fnx

using namespace fnx;

extern "C" {

size_t fnx_image_loss_bytes(int32_t V, int32_t C, int32_t H, int32_t W) {
    return sizeof(float) * 3 * (size_t)V * C * H * W + 256;
}

int fnx_image_loss(int32_t V, int32_t C, int32_t H, int32_t W, const float *img, const float *gt, int32_t grey, float w_l1,
                   float w_ssim, float *dL_dimg, float *l1_mean, float *ssim_mean, void *scratch, fnx_stream_t stream) {
    FNX_REQUIRE(V >= 1 && C >= 1 && H > 0 && W > 0 && img && gt && l1_mean && ssim_mean && scratch, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    static const Win win = make_window();
    const int Ce = grey ? 1 : C;
    float *maps = (float *)align_up((size_t)scratch, 256);
    FNX_CUDA_TRY(cudaMemsetAsync(l1_mean, 0, sizeof(float) * V, st));
    FNX_CUDA_TRY(cudaMemsetAsync(ssim_mean, 0, sizeof(float) * V, st));
    dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, V * Ce), block(LT, LT);
    prof_begin(SEC_IMAGE_LOSS, st);
    ssim_fwd_kernel<<<grid, block, 0, st>>>(V, C, Ce, H, W, grey != 0, img, gt, win, maps, l1_mean, ssim_mean);
    FNX_LAUNCH_CHECK("ssim_fwd_kernel");
    if (dL_dimg) {
        ssim_bwd_kernel<<<grid, block, 0, st>>>(V, C, Ce, H, W, grey != 0, img, gt, win, maps, w_l1, w_ssim, dL_dimg);
        FNX_LAUNCH_CHECK("ssim_bwd_kernel");
    }
    scale_means_kernel<<<1, 32, 0, st>>>(V, 1.0f / ((float)Ce * (float)H * (float)W), l1_mean, ssim_mean);
    prof_end(SEC_IMAGE_LOSS, st);
    FNX_LAUNCH_CHECK("scale_means_kernel");
    return FNX_OK;
}

}  // extern "C"
