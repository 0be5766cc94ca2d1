// gs_train.cu -- per-Gaussian parameter kernels of the static-background training stage.
//
// Replaces, per iteration of FD/entries_fluid_nexus/train_background.py:160-273, the torch chains around the rasterizer:
//   activations          gm_background.py:29-37,90-107   scaling = exp, opacity = sigmoid, rotation = F.normalize
//   their autograd twins (loss.backward through exp / sigmoid / normalize)
//   scaling regulariser  train_background.py:194-201      mean(max(s_max/s_min - threshold, 0))
//   densification stats  train_background.py:238-243, gm_background.py:472-476
//   torch.optim.Adam over the five groups xyz / color / opacity / scaling / rotation (eps 1e-15), gm_background.py:155-168
// with two launches: fnx_gs_activate (before the rasterizer) and fnx_gs_update (after its backward).
#include "common.cuh"

namespace fnx {

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void gs_activate_kernel(int P, const float *__restrict__ raw_scaling, const float *__restrict__ raw_opacity,
                                   const float *__restrict__ raw_rotation, float *__restrict__ scales, float *__restrict__ opacity,
                                   float *__restrict__ rotation) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
#pragma unroll
    for (int k = 0; k < 3; k++) scales[3 * i + k] = expf(raw_scaling[3 * i + k]);
    opacity[i] = sigmoidf(raw_opacity[i]);
    const float4 q = *reinterpret_cast<const float4 *>(raw_rotation + 4 * (size_t)i);
    // F.normalize: x / max(|x|, 1e-12)
    const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    *reinterpret_cast<float4 *>(rotation + 4 * (size_t)i) = make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
}

struct AdamCoef {
    float beta1, beta2, eps, bc1, bc2_sqrt;
};
__device__ __forceinline__ void adam_update(float &p, float &m, float &v, float g, float lr, const AdamCoef &c) {
    m = m + (g - m) * (1.0f - c.beta1);  // lerp form used by torch
    v = c.beta2 * v + (1.0f - c.beta2) * g * g;
    p = p - (lr / c.bc1) * (m / (sqrtf(v) / c.bc2_sqrt + c.eps));
}

// One thread per Gaussian: chain the rasterizer's gradients (w.r.t. the ACTIVATED attributes) through the activations,
// add the scaling regulariser's gradient, update the densification statistics, Adam-update all five raw tensors.
template <int C>
__global__ void __launch_bounds__(256)
gs_update_kernel(int P, fnx_gs_state s, fnx_gs_grads g, fnx_gs_hparams h, AdamCoef c, const int *__restrict__ radii,
                 float *__restrict__ reg_loss) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float reg = 0.f;
    if (i < P) {
        // ---- scaling: s = exp(raw) ----
        float raw_s[3], sc[3], gs[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            raw_s[k] = s.scaling[3 * i + k];
            sc[k] = expf(raw_s[k]);
            gs[k] = g.dL_dscales ? g.dL_dscales[3 * i + k] : 0.f;
        }
        if (h.lambda_reg_scaling > 0.f) {
            // torch.max / torch.min over dim=1 return the FIRST extremal index; their backward routes to that element
            int imax = 0, imin = 0;
#pragma unroll
            for (int k = 1; k < 3; k++) {
                if (sc[k] > sc[imax]) imax = k;
                if (sc[k] < sc[imin]) imin = k;
            }
            const float ratio = sc[imax] / sc[imin] - h.reg_ratio_threshold;
            if (ratio > 0.f) {
                reg = ratio;
                const float w = h.lambda_reg_scaling / (float)P;
                const float dmax = w / sc[imin], dmin = -w * sc[imax] / (sc[imin] * sc[imin]);
#pragma unroll
                for (int k = 0; k < 3; k++) gs[k] += (k == imax ? dmax : 0.f) + (k == imin ? dmin : 0.f);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            adam_update(s.scaling[3 * i + k], s.m_scaling[3 * i + k], s.v_scaling[3 * i + k], gs[k] * sc[k], h.lr_scaling, c);
        // ---- opacity: o = sigmoid(raw) ----
        {
            const float o = sigmoidf(s.opacity[i]);
            const float go = (g.dL_dopacity ? g.dL_dopacity[i] : 0.f) * o * (1.0f - o);
            adam_update(s.opacity[i], s.m_opacity[i], s.v_opacity[i], go, h.lr_opacity, c);
        }
        // ---- rotation: q_n = q / max(|q|, eps)  =>  dq = (g - q_n (q_n . g)) / |q| ----
        {
            float4 q = *reinterpret_cast<const float4 *>(s.rotation + 4 * (size_t)i);
            const float4 gq = g.dL_drotations ? *reinterpret_cast<const float4 *>(g.dL_drotations + 4 * (size_t)i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
            const float n = fmaxf(nrm, 1e-12f);
            const float4 qn = make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
            // below the clamp the norm is a constant and the map is linear
            const float dot = nrm > 1e-12f ? (qn.x * gq.x + qn.y * gq.y + qn.z * gq.z + qn.w * gq.w) : 0.f;
            float4 m4 = *reinterpret_cast<float4 *>(s.m_rotation + 4 * (size_t)i), v4 = *reinterpret_cast<float4 *>(s.v_rotation + 4 * (size_t)i);
            adam_update(q.x, m4.x, v4.x, (gq.x - qn.x * dot) / n, h.lr_rotation, c);
            adam_update(q.y, m4.y, v4.y, (gq.y - qn.y * dot) / n, h.lr_rotation, c);
            adam_update(q.z, m4.z, v4.z, (gq.z - qn.z * dot) / n, h.lr_rotation, c);
            adam_update(q.w, m4.w, v4.w, (gq.w - qn.w * dot) / n, h.lr_rotation, c);
            *reinterpret_cast<float4 *>(s.rotation + 4 * (size_t)i) = q;
            *reinterpret_cast<float4 *>(s.m_rotation + 4 * (size_t)i) = m4;
            *reinterpret_cast<float4 *>(s.v_rotation + 4 * (size_t)i) = v4;
        }
        // ---- xyz, colour: identity activations ----
#pragma unroll
        for (int k = 0; k < 3; k++)
            adam_update(s.xyz[3 * i + k], s.m_xyz[3 * i + k], s.v_xyz[3 * i + k], g.dL_dmeans3D ? g.dL_dmeans3D[3 * i + k] : 0.f, h.lr_xyz, c);
#pragma unroll
        for (int k = 0; k < C; k++)
            adam_update(s.color[(size_t)C * i + k], s.m_color[(size_t)C * i + k], s.v_color[(size_t)C * i + k],
                        g.dL_dcolors ? g.dL_dcolors[(size_t)C * i + k] : 0.f, h.lr_color, c);
        // ---- densification statistics of the Gaussians visible in this view (radii > 0) ----
        if (h.update_stats && radii != nullptr && radii[i] > 0) {
            s.max_radii2D[i] = fmaxf(s.max_radii2D[i], (float)radii[i]);
            const float gx = g.dL_dmeans2D[3 * i], gy = g.dL_dmeans2D[3 * i + 1];
            s.xyz_gradient_accum[i] += sqrtf(gx * gx + gy * gy);
            s.denom[i] += 1.0f;
        }
    }
    if (reg_loss != nullptr && h.lambda_reg_scaling > 0.f) {
        reg = warp_sum(reg);
        if ((threadIdx.x & 31) == 0 && reg != 0.f) atomicAdd(reg_loss, reg / (float)P);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Level-two ("visual particle") stage: FD/entries_fluid_nexus/train_visual_particle.py:133-222 (ScalarReal twin :129-218).
// Positions are fixed (loaded from the physical stage); colour / opacity / scales / rotation of the V visual particles are
// nn.Parameters (gm_dynamics.py:380-397 training_setup_current_level_two, one Adam group each, eps 1e-15).  Per view the loss
// is the image term + lambda_consistency_X * l2_loss_consistency(X, prev_X) (= mse over the first prev_num rows,
// loss_utils.py:138-146; on the RAW tensors) + lambda_reg_scaling * mean(max(s_max / s_min - threshold, 0)); gradients are
// summed over the views and scaled by 1 / batch (gm_dynamics.py:474-503), so the view-independent terms enter with weight 1.
// One thread per particle: chain the rasterizer's gradients through the activations, add the consistency / regulariser
// gradients, accumulate the loss scalars, Adam-update the tensors that are fitted.
// ---------------------------------------------------------------------------------------------------------------
template <int CP>   // colour channels of the PARAMETER (1: grey particles, repeated to 3 render channels by pipe_dynamics.py:118-120)
__global__ void __launch_bounds__(256)
gs_level_two_kernel(int V, fnx_gs_state s, fnx_gs_grads g, fnx_gs_level_two h, AdamCoef c, int render_channels, float *__restrict__ losses) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float l_col = 0.f, l_op = 0.f, l_sc = 0.f, l_rot = 0.f, reg = 0.f;
    if (i < V) {
        const bool has_prev = i < h.prev_num;
        // ---- colour (identity activation) ----
        if (h.fit_color) {
#pragma unroll
            for (int k = 0; k < CP; k++) {
                float gc = 0.f;
                if (g.dL_dcolors) {
                    if (CP == 1) for (int ch = 0; ch < render_channels; ch++) gc += g.dL_dcolors[(size_t)render_channels * i + ch];
                    else gc = g.dL_dcolors[(size_t)CP * i + k];
                }
                if (has_prev && h.prev_color) {
                    const float d = s.color[(size_t)CP * i + k] - h.prev_color[(size_t)CP * i + k];
                    l_col += d * d;
                    gc += h.lambda_consistency_color * 2.0f * d / ((float)h.prev_num * CP);
                }
                adam_update(s.color[(size_t)CP * i + k], s.m_color[(size_t)CP * i + k], s.v_color[(size_t)CP * i + k], gc, h.lr_color, c);
            }
        }
        // ---- opacity: o = sigmoid(raw) ----
        if (h.fit_opacity) {
            const float raw = s.opacity[i];
            const float o = sigmoidf(raw);
            float go = (g.dL_dopacity ? g.dL_dopacity[i] : 0.f) * o * (1.0f - o);
            if (has_prev && h.prev_opacity) {
                const float d = raw - h.prev_opacity[i];
                l_op += d * d;
                go += h.lambda_consistency_opacity * 2.0f * d / (float)h.prev_num;
            }
            adam_update(s.opacity[i], s.m_opacity[i], s.v_opacity[i], go, h.lr_opacity, c);
        }
        // ---- scales: s = exp(raw), regulariser on the activated values ----
        if (h.fit_scales) {
            float raw_s[3], sc[3], gs[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                raw_s[k] = s.scaling[3 * i + k];
                sc[k] = expf(raw_s[k]);
                gs[k] = (g.dL_dscales ? g.dL_dscales[3 * i + k] : 0.f) * sc[k];
            }
            if (h.lambda_reg_scaling > 0.f) {
                int imax = 0, imin = 0;   // torch.max / torch.min(dim=1): first extremal index, gradient routed there
#pragma unroll
                for (int k = 1; k < 3; k++) {
                    if (sc[k] > sc[imax]) imax = k;
                    if (sc[k] < sc[imin]) imin = k;
                }
                const float ratio = sc[imax] / sc[imin] - h.reg_ratio_threshold;
                if (ratio > 0.f) {
                    reg = ratio;
                    const float w = h.lambda_reg_scaling / (float)V;
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        gs[k] += ((k == imax ? w / sc[imin] : 0.f) + (k == imin ? -w * sc[imax] / (sc[imin] * sc[imin]) : 0.f)) * sc[k];
                }
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (has_prev && h.prev_scales) {
                    const float d = raw_s[k] - h.prev_scales[3 * i + k];
                    l_sc += d * d;
                    gs[k] += h.lambda_consistency_scales * 2.0f * d / ((float)h.prev_num * 3.0f);
                }
                adam_update(s.scaling[3 * i + k], s.m_scaling[3 * i + k], s.v_scaling[3 * i + k], gs[k], h.lr_scaling, c);
            }
        }
        // ---- rotation: q_n = q / max(|q|, 1e-12) ----
        if (h.fit_rotation) {
            float4 q = *reinterpret_cast<const float4 *>(s.rotation + 4 * (size_t)i);
            const float4 gq = g.dL_drotations ? *reinterpret_cast<const float4 *>(g.dL_drotations + 4 * (size_t)i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
            const float n = fmaxf(nrm, 1e-12f);
            const float4 qn = make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
            const float dot = nrm > 1e-12f ? (qn.x * gq.x + qn.y * gq.y + qn.z * gq.z + qn.w * gq.w) : 0.f;
            float gr[4] = {(gq.x - qn.x * dot) / n, (gq.y - qn.y * dot) / n, (gq.z - qn.z * dot) / n, (gq.w - qn.w * dot) / n};
            if (has_prev && h.prev_rotation) {
                const float4 pq = *reinterpret_cast<const float4 *>(h.prev_rotation + 4 * (size_t)i);
                const float d[4] = {q.x - pq.x, q.y - pq.y, q.z - pq.z, q.w - pq.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    l_rot += d[k] * d[k];
                    gr[k] += h.lambda_consistency_rotation * 2.0f * d[k] / ((float)h.prev_num * 4.0f);
                }
            }
            float4 m4 = *reinterpret_cast<float4 *>(s.m_rotation + 4 * (size_t)i), v4 = *reinterpret_cast<float4 *>(s.v_rotation + 4 * (size_t)i);
            adam_update(q.x, m4.x, v4.x, gr[0], h.lr_rotation, c);
            adam_update(q.y, m4.y, v4.y, gr[1], h.lr_rotation, c);
            adam_update(q.z, m4.z, v4.z, gr[2], h.lr_rotation, c);
            adam_update(q.w, m4.w, v4.w, gr[3], h.lr_rotation, c);
            *reinterpret_cast<float4 *>(s.rotation + 4 * (size_t)i) = q;
            *reinterpret_cast<float4 *>(s.m_rotation + 4 * (size_t)i) = m4;
            *reinterpret_cast<float4 *>(s.v_rotation + 4 * (size_t)i) = v4;
        }
    }
    if (losses != nullptr) {   // {color, opacity, scales, rotation consistency (mse), scaling regulariser (mean)}
        const float pn = (float)max(h.prev_num, 1);
        const float vals[5] = {l_col / (pn * CP), l_op / pn, l_sc / (pn * 3.0f), l_rot / (pn * 4.0f), reg / (float)V};
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float r = warp_sum(vals[k]);
            if ((threadIdx.x & 31) == 0 && r != 0.f) atomicAdd(losses + k, r);
        }
    }
}

}  // namespace fnx

using namespace fnx;

extern "C" {

int fnx_gs_activate(int32_t P, const float *raw_scaling, const float *raw_opacity, const float *raw_rotation, float *scales,
                    float *opacity, float *rotation, fnx_stream_t stream) {
    FNX_REQUIRE(P >= 0 && (P == 0 || (raw_scaling && raw_opacity && raw_rotation && scales && opacity && rotation)), "bad arguments");
    if (P == 0) return FNX_OK;
    gs_activate_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, raw_scaling, raw_opacity, raw_rotation, scales, opacity, rotation);
    FNX_LAUNCH_CHECK("gs_activate_kernel");
    return FNX_OK;
}

int fnx_gs_update(int32_t P, int32_t C, const fnx_gs_state *state, const fnx_gs_grads *grads, const fnx_gs_hparams *hp,
                  const int32_t *radii, float *reg_loss, fnx_stream_t stream) {
    FNX_REQUIRE(P >= 0 && (C == 1 || C == 3) && state && grads && hp, "bad arguments");
    FNX_REQUIRE(hp->step >= 1, "step counts from 1 (the count AFTER this update)");
    if (P == 0) return FNX_OK;
    const fnx_gs_state &s = *state;
    FNX_REQUIRE(s.xyz && s.color && s.opacity && s.scaling && s.rotation && s.m_xyz && s.v_xyz && s.m_color && s.v_color && s.m_opacity &&
                    s.v_opacity && s.m_scaling && s.v_scaling && s.m_rotation && s.v_rotation, "state tensors missing");
    FNX_REQUIRE(!hp->update_stats || (radii && grads->dL_dmeans2D && s.max_radii2D && s.xyz_gradient_accum && s.denom),
                "update_stats needs radii, dL_dmeans2D and the three statistics tensors");
    cudaStream_t st = (cudaStream_t)stream;
    AdamCoef c;
    c.beta1 = hp->beta1; c.beta2 = hp->beta2; c.eps = hp->eps;
    c.bc1 = (float)(1.0 - pow((double)hp->beta1, hp->step));
    c.bc2_sqrt = (float)sqrt(1.0 - pow((double)hp->beta2, hp->step));
    if (reg_loss) FNX_CUDA_TRY(cudaMemsetAsync(reg_loss, 0, sizeof(float), st));
    if (C == 3) gs_update_kernel<3><<<(P + 255) / 256, 256, 0, st>>>(P, s, *grads, *hp, c, radii, reg_loss);
    else gs_update_kernel<1><<<(P + 255) / 256, 256, 0, st>>>(P, s, *grads, *hp, c, radii, reg_loss);
    FNX_LAUNCH_CHECK("gs_update_kernel");
    return FNX_OK;
}

int fnx_gs_update_level_two(int32_t V, int32_t render_channels, const fnx_gs_state *state, const fnx_gs_grads *grads,
                            const fnx_gs_level_two *hp, float *losses5, fnx_stream_t stream) {
    FNX_REQUIRE(V >= 0 && (render_channels == 1 || render_channels == 3) && state && grads && hp, "bad arguments");
    FNX_REQUIRE(hp->step >= 1 && (hp->color_channels == 1 || hp->color_channels == 3) && hp->color_channels <= render_channels,
                "step counts from 1; colour parameter channels must be 1 or 3 and not exceed the render channels");
    FNX_REQUIRE(hp->prev_num >= 0 && hp->prev_num <= V, "prev_num must be in [0, V]");
    const fnx_gs_state &s = *state;
    FNX_REQUIRE((!hp->fit_color || (s.color && s.m_color && s.v_color)) && (!hp->fit_opacity || (s.opacity && s.m_opacity && s.v_opacity)) &&
                    (!hp->fit_scales || (s.scaling && s.m_scaling && s.v_scaling)) && (!hp->fit_rotation || (s.rotation && s.m_rotation && s.v_rotation)),
                "state tensors of a fitted attribute missing");
    cudaStream_t st = (cudaStream_t)stream;
    if (losses5) FNX_CUDA_TRY(cudaMemsetAsync(losses5, 0, 5 * sizeof(float), st));
    if (V == 0) return FNX_OK;
    AdamCoef c;
    c.beta1 = hp->beta1; c.beta2 = hp->beta2; c.eps = hp->eps;
    c.bc1 = (float)(1.0 - pow((double)hp->beta1, hp->step));
    c.bc2_sqrt = (float)sqrt(1.0 - pow((double)hp->beta2, hp->step));
    if (hp->color_channels == 1) gs_level_two_kernel<1><<<(V + 255) / 256, 256, 0, st>>>(V, s, *grads, *hp, c, render_channels, losses5);
    else gs_level_two_kernel<3><<<(V + 255) / 256, 256, 0, st>>>(V, s, *grads, *hp, c, render_channels, losses5);
    FNX_LAUNCH_CHECK("gs_level_two_kernel");
    return FNX_OK;
}

}  // extern "C"
