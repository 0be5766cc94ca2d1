// image_loss.cu -- fused L1 + SSIM image loss (value and gradient) for sm_100a.
//
// Replaces, per view and per step, the reference's torch graph
//   l1_loss                      FD/utils/loss_utils.py:9-10
//   ssim / _ssim                 FD/utils/loss_utils.py:21-64 (window rebuilt on the CPU and uploaded every call :35-39,
//                                five grouped 11x11 conv2d forward + their backward)
//   grey conversion              FD/entries_fluid_nexus/train_physical_particle.py:356-360
//   weighting                    FD/entries_scalar_real/train_physical_particle.py:346-347
// with two kernels: (1) per 32x32 tile with a 5-pixel halo, register-tiled separable 11-tap Gaussian of
// {x, y, x^2, y^2, xy}, the SSIM map, its three partial-derivative maps and the per-view loss means; (2) separable
// Gaussian of the derivative maps -> dL/dimage, with the L1 sign term added.  Zero padding like F.conv2d(padding=5).
#include "common.cuh"

namespace fnx {

constexpr int TW = 32, TH = 32;    // output tile of one CTA
constexpr int HALO = 5;            // window 11
constexpr int EW = TW + 2 * HALO;  // 42 input columns
constexpr int EH = TH + 2 * HALO;  // 42 input rows
constexpr int ES = EW + 1;         // padded row stride of the input tile (conflict-free strip reads, see pass A)
constexpr int NT = 256;            // threads per CTA
constexpr int SEG = 8;             // pass A: outputs per thread along x
constexpr int RSEG = 4;            // pass B: outputs per thread along y

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

// Register-tiled separable 11-tap Gaussian.  Pass A (rows): a thread owns SEG consecutive outputs of one input row: it
// pulls SEG+10 inputs into registers once and produces SEG x Q filtered values.  Pass B (columns): a thread owns RSEG
// consecutive outputs of one column and pulls RSEG+10 rows.  Per output that is (SEG+10)/SEG + (RSEG+10)/RSEG shared
// loads per quantity instead of 22.

// pass 1: SSIM map + derivative maps + loss sums.  grid (tiles_x, tiles_y, V*Ce), Ce = grey ? 1 : C
__global__ void __launch_bounds__(NT, 3)
ssim_fwd_kernel(int V, int C, int Ce, int H, int W, bool grey, const float *__restrict__ img, const float *__restrict__ gt,
                Win win, float inv_n, float *__restrict__ maps /*[V*Ce][3][H][W]*/, float *__restrict__ l1_sum,
                float *__restrict__ ssim_sum) {
    __shared__ float sx[EH][ES], sy[EH][ES];
    __shared__ float hq[5][EH][TW];
    __shared__ float red[2][NT / 32];
    const int vc = blockIdx.z, v = vc / Ce, c = vc % Ce;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.x;
    for (int k = tid; k < EH * EW; k += NT) {
        const int ly = k / EW, lx = k % EW;
        sx[ly][lx] = load_px(img, C, H, W, v, c, grey, y0 + ly - HALO, x0 + lx - HALO);
        sy[ly][lx] = load_px(gt, C, H, W, v, c, grey, y0 + ly - HALO, x0 + lx - HALO);
    }
    __syncthreads();
    // pass A: EH rows x (TW / SEG) strips
    for (int k = tid; k < EH * (TW / SEG); k += NT) {
        const int ly = k / (TW / SEG), s0 = (k % (TW / SEG)) * SEG;
        float x[SEG + 10], y[SEG + 10];
#pragma unroll
        for (int i = 0; i < SEG + 10; i++) { x[i] = sx[ly][s0 + i]; y[i] = sy[ly][s0 + i]; }
#pragma unroll
        for (int o = 0; o < SEG; o++) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
            for (int t = 0; t < 11; t++) {
                const float w = win.g[t], xv = x[o + t], yv = y[o + t];
                a0 += w * xv; a1 += w * yv; a2 += w * xv * xv; a3 += w * yv * yv; a4 += w * xv * yv;
            }
            hq[0][ly][s0 + o] = a0; hq[1][ly][s0 + o] = a1; hq[2][ly][s0 + o] = a2; hq[3][ly][s0 + o] = a3; hq[4][ly][s0 + o] = a4;
        }
    }
    __syncthreads();
    // pass B: TW columns x (TH / RSEG) strips; lane = column
    const int lx = tid % TW, r0 = (tid / TW) * RSEG;
    float mu1[RSEG], mu2[RSEG], e11[RSEG], e22[RSEG], e12[RSEG];
#pragma unroll
    for (int o = 0; o < RSEG; o++) mu1[o] = mu2[o] = e11[o] = e22[o] = e12[o] = 0.f;
#pragma unroll
    for (int i = 0; i < RSEG + 10; i++) {
        const float q0 = hq[0][r0 + i][lx], q1 = hq[1][r0 + i][lx], q2 = hq[2][r0 + i][lx], q3 = hq[3][r0 + i][lx], q4 = hq[4][r0 + i][lx];
#pragma unroll
        for (int o = 0; o < RSEG; o++) {
            const int t = i - o;
            if (t >= 0 && t < 11) {
                const float w = win.g[t];
                mu1[o] += w * q0; mu2[o] += w * q1; e11[o] += w * q2; e22[o] += w * q3; e12[o] += w * q4;
            }
        }
    }
    float l1 = 0.f, ss = 0.f;
    const int px = x0 + lx;
    const size_t HW = (size_t)H * W;
    float *m = maps + (size_t)vc * 3 * HW;
#pragma unroll
    for (int o = 0; o < RSEG; o++) {
        const int py = y0 + r0 + o;
        if (px < W && py < H) {
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            const float mu1_sq = mu1[o] * mu1[o], mu2_sq = mu2[o] * mu2[o], mu1_mu2 = mu1[o] * mu2[o];
            const float sigma1_sq = e11[o] - mu1_sq, sigma2_sq = e22[o] - mu2_sq, sigma12 = e12[o] - mu1_mu2;
            const float A1 = 2.f * mu1_mu2 + C1, A2 = 2.f * sigma12 + C2;
            const float B1 = mu1_sq + mu2_sq + C1, B2 = sigma1_sq + sigma2_sq + C2;
            const float invB = 1.f / (B1 * B2);
            const float S = (A1 * A2) * invB;
            ss += S;
            l1 += fabsf(sx[r0 + o + HALO][lx + HALO] - sy[r0 + o + HALO][lx + HALO]);
            // partial derivatives of S w.r.t. mu1, E[x^2], E[xy] (the window-averaged quantities that depend on x)
            const float dS_dmu1 = (2.f * mu2[o] * (A2 - A1)) * invB - S * (2.f * mu1[o] / B1 - 2.f * mu1[o] / B2);
            const float dS_de11 = -S / B2;
            const float dS_de12 = 2.f * A1 * invB;
            const size_t off = (size_t)py * W + px;
            m[off] = dS_dmu1; m[HW + off] = dS_de11; m[2 * HW + off] = dS_de12;
        }
    }
    l1 = warp_sum(l1);
    ss = warp_sum(ss);
    if ((tid & 31) == 0) { red[0][tid >> 5] = l1; red[1][tid >> 5] = ss; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < NT / 32; k++) { a += red[0][k]; b += red[1][k]; }
        atomicAdd(&l1_sum[v], a * inv_n);   // the per-view means of the reference's .mean()
        atomicAdd(&ssim_sum[v], b * inv_n);
    }
}

// pass 2: dL/dimg = w_l1/N * sign(x-y) - w_ssim/N * (G*M1 + 2x G*M2 + y G*M3)
__global__ void __launch_bounds__(NT, 3)
ssim_bwd_kernel(int V, int C, int Ce, int H, int W, bool grey, const float *__restrict__ img, const float *__restrict__ gt,
                int C_src, bool grey_src /* layout of img / gt: the originals, or the precomputed channel means (1, false) */,
                Win win, const float *__restrict__ maps, float w_l1, float w_ssim, float *__restrict__ dL_dimg) {
    __shared__ float sm[3][EH][ES];
    __shared__ float hq[3][EH][TW];
    const int vc = blockIdx.z, v = vc / Ce, c = vc % Ce;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.x;
    const size_t HW = (size_t)H * W;
    const float *m = maps + (size_t)vc * 3 * HW;
    for (int k = tid; k < EH * EW; k += NT) {
        const int ly = k / EW, lx = k % EW;
        const int gx = x0 + lx - HALO, gy = y0 + ly - HALO;
        const bool in = gx >= 0 && gy >= 0 && gx < W && gy < H;
        const size_t o = (size_t)gy * W + gx;
        sm[0][ly][lx] = in ? m[o] : 0.f;
        sm[1][ly][lx] = in ? m[HW + o] : 0.f;
        sm[2][ly][lx] = in ? m[2 * HW + o] : 0.f;
    }
    __syncthreads();
    for (int k = tid; k < EH * (TW / SEG); k += NT) {
        const int ly = k / (TW / SEG), s0 = (k % (TW / SEG)) * SEG;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            float x[SEG + 10];
#pragma unroll
            for (int i = 0; i < SEG + 10; i++) x[i] = sm[q][ly][s0 + i];
#pragma unroll
            for (int o = 0; o < SEG; o++) {
                float a = 0.f;
#pragma unroll
                for (int t = 0; t < 11; t++) a += win.g[t] * x[o + t];
                hq[q][ly][s0 + o] = a;
            }
        }
    }
    __syncthreads();
    const int lx = tid % TW, r0 = (tid / TW) * RSEG;
    float g1[RSEG], g2[RSEG], g3[RSEG];
#pragma unroll
    for (int o = 0; o < RSEG; o++) g1[o] = g2[o] = g3[o] = 0.f;
#pragma unroll
    for (int i = 0; i < RSEG + 10; i++) {
        const float q0 = hq[0][r0 + i][lx], q1 = hq[1][r0 + i][lx], q2 = hq[2][r0 + i][lx];
#pragma unroll
        for (int o = 0; o < RSEG; o++) {
            const int t = i - o;
            if (t >= 0 && t < 11) {
                const float w = win.g[t];
                g1[o] += w * q0; g2[o] += w * q1; g3[o] += w * q2;
            }
        }
    }
    const int px = x0 + lx;
    const float n = (float)Ce * (float)H * (float)W;  // elements the reference's .mean() runs over (per view)
#pragma unroll
    for (int o = 0; o < RSEG; o++) {
        const int py = y0 + r0 + o;
        if (px >= W || py >= H) continue;
        const float x = load_px(img, C_src, H, W, v, c, grey_src, py, px), y = load_px(gt, C_src, H, W, v, c, grey_src, py, px);
        const float d = x - y;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        // in grey mode the reference's 3 identical channels each carry 1/3 of the mean, and d grey / d channel = 1/C
        float gout = (w_l1 * sgn - w_ssim * (g1[o] + 2.f * x * g2[o] + y * g3[o])) / n;
        const size_t off = (size_t)py * W + px;
        if (!grey) {
            dL_dimg[((size_t)v * C + c) * HW + off] = gout;
        } else {
            gout /= C;
            for (int k = 0; k < C; k++) dL_dimg[((size_t)v * C + k) * HW + off] = gout;
        }
    }
}

// channel means of img and gt, once per call in grey mode: the SSIM kernels then read one value per pixel and image
// (with their 1.7x halo re-reads) instead of C
__global__ void grey_kernel(int V, int C, size_t HW, const float *__restrict__ img, const float *__restrict__ gt,
                            float *__restrict__ grey_img, float *__restrict__ grey_gt) {
    // grid (pixels / 256, V): no 64-bit division per pixel (it was 3/4 of this kernel's instructions)
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (o >= HW) return;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < C; k++) { a += img[(v * C + k) * HW + o]; b += gt[(v * C + k) * HW + o]; }
    grey_img[v * HW + o] = a / C;  // torch.mean over the channel dim
    grey_gt[v * HW + o] = b / C;
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
    static_assert(NT == TW * (TH / RSEG), "pass B maps one thread to RSEG rows of one column");
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, V * Ce);
    prof_begin(SEC_IMAGE_LOSS, st);
    const size_t HW = (size_t)H * W;
    const float *src_img = img, *src_gt = gt;
    int C_src = C;
    bool grey_src = grey != 0;
    if (grey && C > 1) {  // scratch holds 3*V*C*HW floats, the maps of grey mode use 3*V*HW of them: room for the two means
        float *grey_img = maps + 3 * (size_t)V * HW, *grey_gt = grey_img + (size_t)V * HW;
        grey_kernel<<<dim3((unsigned)((HW + 255) / 256), V), 256, 0, st>>>(V, C, HW, img, gt, grey_img, grey_gt);
        FNX_LAUNCH_CHECK("grey_kernel");
        src_img = grey_img; src_gt = grey_gt; C_src = 1; grey_src = false;
    }
    ssim_fwd_kernel<<<grid, NT, 0, st>>>(V, C_src, Ce, H, W, grey_src, src_img, src_gt, win, 1.0f / ((float)Ce * (float)H * (float)W), maps,
                                         l1_mean, ssim_mean);
    FNX_LAUNCH_CHECK("ssim_fwd_kernel");
    if (dL_dimg) {
        ssim_bwd_kernel<<<grid, NT, 0, st>>>(V, C, Ce, H, W, grey != 0, src_img, src_gt, C_src, grey_src, win, maps, w_l1, w_ssim, dL_dimg);
        FNX_LAUNCH_CHECK("ssim_bwd_kernel");
    }
    prof_end(SEC_IMAGE_LOSS, st);
    return FNX_OK;
}

}  // extern "C"
