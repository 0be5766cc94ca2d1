// raster.cu -- tile-binned differentiable Gaussian rasterizer for sm_100a (B200).
//
// Replaces the reference extension modules diff_gaussian_rasterization_ch3 / _ch1
// (R3 = FluidDynamics/submodules/gaussian_rasterization_ch3 in the reference tree):
//   forward   R3/cuda_rasterizer/rasterizer_impl.cu:184-319, forward.cu:148-244 (preprocess), :249-373 (blend)
//   backward  R3/cuda_rasterizer/rasterizer_impl.cu:323-414, backward.cu:384-536 (blend), :137-263, :267-381
// It is a new design, not a translation:
//   * V cameras per call (grid.y = view); the reference renders one.
//   * Gaussians are depth-sorted once per view (V*P keys), instances are then emitted in depth order and only
//     *stably partitioned by tile id* (2 radix passes over 8-byte pairs instead of 6 over 12-byte pairs).
//   * opacity-aware exact tile culling: a (tile, Gaussian) instance is emitted only if some pixel of the tile can
//     reach alpha >= 1/255.  Instances the reference would skip at every pixel anyway are never created, so
//     images and gradients are unchanged (`radii` keeps the reference's 3-sigma value).
//   * the sorted instances are re-packed into a contiguous record stream; blend kernels pull each tile's span
//     into shared memory with 1-D bulk async copies (TMA engine, cp.async.bulk + mbarrier), double buffered.
//   * no mid-forward host sync is required (capacity hint + device-side overflow flag, checked after queuing).
//   * backward: a value-splitting warp-shuffle reduction of the per-pixel partials (12 shuffles for 9 values), then one
//     red.global.add.f32 per value into a contiguous 32/48-byte accumulator row per (view, Gaussian); a single fused
//     per-Gaussian kernel turns the rows into dL/d{mean, scale, rotation, opacity, colour} summed over the views.
//   * static + dynamic streams (fnx_raster_static_prepare / _blend_merged / _backward_merged): a frozen set is binned
//     once; per iteration only the moving Gaussians are binned (per-tile buckets sorted in shared memory, no global sort)
//     and merged per tile; tiles without a moving instance keep their pixels; the backward starts at the tile's last
//     moving record from a snapshot the forward took there.
#include <cub/cub.cuh>
#include <cstdlib>
#include <mutex>

#include "geom_grad.cuh"
#include "raster.cuh"

namespace fnx {

// ---------------------------------------------------------------------------------------------------------------
// scratch layout
// ---------------------------------------------------------------------------------------------------------------
struct DepthScanIn {  // reads tiles_touched in depth-sorted order
    const uint32_t *tiles_touched;
    const uint32_t *order;
    __host__ __device__ uint32_t operator()(uint32_t k) const { return tiles_touched[order[k]]; }
};
using DepthScanIter = cub::TransformInputIterator<uint32_t, DepthScanIn, cub::CountingInputIterator<uint32_t>>;

// CUB's size queries walk its dispatch layer (device attribute / occupancy calls): memoise per problem size
static size_t geom_cub_bytes(int n) {
    static thread_local int cached_n = -1;
    static thread_local size_t cached = 0;
    if (n == cached_n) return cached;
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (uint32_t *)nullptr, (uint32_t *)nullptr, n, 0, 64);
    DepthScanIter it(cub::CountingInputIterator<uint32_t>(0), DepthScanIn{nullptr, nullptr});
    cub::DeviceScan::ExclusiveSum(nullptr, b, it, (uint32_t *)nullptr, n);
    cached_n = n;
    cached = a > b ? a : b;
    return cached;
}

GeomView geom_view(void *chunk, int P, int V) {
    GeomView g;
    char *p = (char *)chunk;
    size_t n = (size_t)P * V;
    g.hdr = carve<GeomHeader>(p, 1);
    g.cov3D = carve<float>(p, (size_t)P * 6);
    g.depth = carve<float>(p, n);
    g.xy = carve<float2>(p, n);
    g.conic_o = carve<float4>(p, n);
    g.tiles_touched = carve<uint32_t>(p, n);
    g.dkeys_in = carve<unsigned long long>(p, n);
    g.dkeys_out = carve<unsigned long long>(p, n);
    g.dvals_in = carve<uint32_t>(p, n);
    g.dvals_out = carve<uint32_t>(p, n);
    g.offsets = carve<uint32_t>(p, n + 1);
    g.accum = carve<float>(p, n * 12);
    g.cub_temp_bytes = geom_cub_bytes((int)n);
    g.cub_temp = carve<char>(p, g.cub_temp_bytes);
    return g;
}
size_t geom_bytes(int P, int V) {
    GeomView g = geom_view((void *)0, P, V);
    return (size_t)((char *)g.cub_temp - (char *)0) + g.cub_temp_bytes + 256;
}

ImageView image_view(void *chunk, int W, int H, int V) {
    ImageView im;
    char *p = (char *)chunk;
    size_t hw = (size_t)W * H * V;
    size_t nt = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE) * V;
    im.final_T = carve<float>(p, hw);
    im.n_contrib = carve<uint32_t>(p, hw);
    im.ranges = carve<uint2>(p, nt);
    im.tile_last = carve<uint32_t>(p, nt * TILE_PATCHES);  // per 8x8 patch
    im.mranges = carve<uint2>(p, nt);
    im.tile_src = carve<uint32_t>(p, nt);
    im.tile_dyn_last = carve<uint32_t>(p, nt);
    im.tile_dyn_first = carve<uint32_t>(p, nt);
    im.tile_cached = carve<uint32_t>(p, nt * TILE_PATCHES);  // per 8x8 patch
    im.tile_count = carve<uint32_t>(p, 2 * nt);  // count and cursor are contiguous: one memset clears both
    im.tile_cursor = im.tile_count + nt;
    im.snap = carve<float4>(p, hw);
    return im;
}
size_t image_bytes(int W, int H, int V) {
    ImageView im = image_view((void *)0, W, H, V);
    return (size_t)((char *)im.snap - (char *)0) + (size_t)W * H * V * sizeof(float4) + 256;
}

static size_t bin_cub_bytes(long long cap) {
    static thread_local long long cached_cap[4] = {-1, -1, -1, -1};
    static thread_local size_t cached[4] = {0, 0, 0, 0};
    static thread_local int next = 0;
    for (int k = 0; k < 4; k++)
        if (cached_cap[k] == cap) return cached[k];
    size_t a = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)cap, 0, 32);
    cached_cap[next] = cap;
    cached[next] = a;
    next = (next + 1) & 3;
    return a;
}
BinView bin_view(void *chunk, long long cap, int C) {
    BinView b;
    char *p = (char *)chunk;
    size_t n = (size_t)(cap > 0 ? cap : 1);
    b.tkeys_in = carve<uint32_t>(p, n);
    b.tkeys_out = carve<uint32_t>(p, n);
    b.tvals_in = carve<uint32_t>(p, n);
    b.tvals_out = carve<uint32_t>(p, n);
    b.records = carve<char>(p, n * (size_t)(C == 3 ? 48 : 32));
    b.bkeys = carve<unsigned long long>(p, n);
    b.bkeys2 = carve<unsigned long long>(p, n);
    b.cub_temp_bytes = bin_cub_bytes((long long)n);
    b.cub_temp = carve<char>(p, b.cub_temp_bytes);
    return b;
}
size_t binning_bytes(long long cap, int C) {
    BinView b = bin_view((void *)0, cap, C);
    return (size_t)((char *)b.cub_temp - (char *)0) + b.cub_temp_bytes + 256;
}

// ---------------------------------------------------------------------------------------------------------------
// small column-major 3x3 helper.  m[c][r] = column c, row r; the product sums k = 0,1,2 left to right.
// (The reference does its 3x3 algebra in that storage/ordering; keeping the same expression trees keeps the
// fp32 roundings, hence tile assignment and depth order, identical.)
// ---------------------------------------------------------------------------------------------------------------
struct M3 {
    float m[3][3];
};
__device__ __forceinline__ M3 m3_mul(const M3 &A, const M3 &B) {
    M3 R;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) R.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1] + A.m[2][r] * B.m[c][2];
    return R;
}
__device__ __forceinline__ M3 m3_T(const M3 &A) {
    M3 R;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) R.m[c][r] = A.m[r][c];
    return R;
}
__device__ __forceinline__ M3 m3_cols(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1,
                                      float c2) {
    M3 R;
    R.m[0][0] = a0; R.m[0][1] = a1; R.m[0][2] = a2;
    R.m[1][0] = b0; R.m[1][1] = b1; R.m[1][2] = b2;
    R.m[2][0] = c0; R.m[2][1] = c1; R.m[2][2] = c2;
    return R;
}

__device__ __forceinline__ float3 xform4x3(const float3 &p, const float *m) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float4 xform4x4(const float3 &p, const float *m) {
    return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14], m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}

// rotation matrix of the un-normalised quaternion (r,x,y,z) in the reference's column order (forward.cu:121-132)
__device__ __forceinline__ M3 quat_to_R(const float4 q) {
    float r = q.x, x = q.y, y = q.z, z = q.w;
    return m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y), 2.f * (x * y + r * z),
                   1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x), 2.f * (x * z - r * y), 2.f * (y * z + r * x),
                   1.f - 2.f * (x * x + y * y));
}

// Sigma = (S R)^T (S R), upper triangle (forward.cu:113-145)
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 s, float mod, const float4 q, float *cov) {
    M3 S = m3_cols(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
    S.m[0][0] = mod * s.x;
    S.m[1][1] = mod * s.y;
    S.m[2][2] = mod * s.z;
    M3 R = quat_to_R(q);
    M3 M = m3_mul(S, R);
    M3 Sig = m3_mul(m3_T(M), M);
    cov[0] = Sig.m[0][0]; cov[1] = Sig.m[0][1]; cov[2] = Sig.m[0][2];
    cov[3] = Sig.m[1][1]; cov[4] = Sig.m[1][2]; cov[5] = Sig.m[2][2];
}

// T = W * J with the FoV-clamped view-space mean (forward.cu:70-96 == backward.cu:159-193)
struct ProjJac {
    float3 t;
    float txtz, tytz;
    M3 W, T;
};
__device__ __forceinline__ ProjJac proj_jacobian(const float3 &mean, float fx, float fy, float tfx, float tfy,
                                                 const float *V) {
    ProjJac o;
    float3 t = xform4x3(mean, V);
    const float limx = 1.3f * tfx, limy = 1.3f * tfy;
    o.txtz = t.x / t.z;
    o.tytz = t.y / t.z;
    t.x = min(limx, max(-limx, o.txtz)) * t.z;
    t.y = min(limy, max(-limy, o.tytz)) * t.z;
    M3 J = m3_cols(fx / t.z, 0.0f, -(fx * t.x) / (t.z * t.z), 0.0f, fy / t.z, -(fy * t.y) / (t.z * t.z), 0, 0, 0);
    o.W = m3_cols(V[0], V[4], V[8], V[1], V[5], V[9], V[2], V[6], V[10]);
    o.T = m3_mul(o.W, J);
    o.t = t;
    return o;
}
__device__ __forceinline__ M3 vrk_of(const float *c) { return m3_cols(c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5]); }

// auxiliary.h:39-41 -- evaluated in double by the reference (1.0 literals), result rounded to float
__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// auxiliary.h:43-50
__device__ __forceinline__ void get_rect(const float2 p, int max_radius, uint2 &rmin, uint2 &rmax, int gx, int gy) {
    rmin.x = min(gx, max(0, (int)((p.x - max_radius) / TILE)));
    rmin.y = min(gy, max(0, (int)((p.y - max_radius) / TILE)));
    rmax.x = min(gx, max(0, (int)((p.x + max_radius + TILE - 1) / TILE)));
    rmax.y = min(gy, max(0, (int)((p.y + max_radius + TILE - 1) / TILE)));
}

// Opacity-aware tile test.  A pixel at offset u from the mean gets alpha = min(.99, o*exp(-q(u))) with
// q(u) = .5*(a ux^2 + c uy^2) + b ux uy, and is blended only if alpha >= 1/255, i.e. q(u) <= tau = ln(255 o).
// The tile is needed iff min over its pixel-centre box of q <= tau.  q is convex (checked), so the minimum is 0 if
// the box contains the mean, else it lies on one of the 4 edges (1-D clamped quadratic).  Conservative margins
// cover fp32 differences between this bound and the per-pixel evaluation.
struct TileCull {
    float a, b, c, tau;
    bool active;
};
__device__ __forceinline__ TileCull make_cull(const float4 con_o, bool exact_rect) {
    TileCull t;
    t.a = con_o.x; t.b = con_o.y; t.c = con_o.z;
    float o = con_o.w;
    bool pd = (t.a > 0.f) && (t.c > 0.f) && (t.a * t.c - t.b * t.b > 0.f);
    t.active = !exact_rect && pd && (o == o);
    t.tau = (o > 0.f) ? __logf(255.0f * o) * 1.001f + 0.05f : -1.0f;  // o <= 0 can never reach 1/255
    return t;
}
__device__ __forceinline__ float edge_min(float A, float B, float Cq, float X, float y0, float y1) {
    // min over y in [y0,y1] of .5*A*X^2 + B*X*y + .5*Cq*y^2   (Cq > 0)
    float ys = min(y1, max(y0, -B * X / Cq));
    return 0.5f * A * X * X + B * X * ys + 0.5f * Cq * ys * ys;
}
// does the pixel-centre box [bx, bx+w-1] x [by, by+h-1] contain a pixel that can reach alpha >= 1/255 ?
__device__ __forceinline__ bool box_needed(const TileCull &t, const float2 mean, int bx, int by, int w, int h) {
    if (!t.active) return true;
    if (t.tau < 0.f) return false;
    float x0 = bx - mean.x, x1 = x0 + (w - 1);
    float y0 = by - mean.y, y1 = y0 + (h - 1);
    if (x0 <= 0.f && x1 >= 0.f && y0 <= 0.f && y1 >= 0.f) return true;
    float m = edge_min(t.a, t.b, t.c, x0, y0, y1);
    m = min(m, edge_min(t.a, t.b, t.c, x1, y0, y1));
    m = min(m, edge_min(t.c, t.b, t.a, y0, x0, x1));
    m = min(m, edge_min(t.c, t.b, t.a, y1, x0, x1));
    return m <= t.tau;
}
__device__ __forceinline__ bool tile_needed(const TileCull &t, const float2 mean, int tx, int ty) {
    if (!t.active) return true;
    if (t.tau < 0.f) return false;
    float x0 = tx * TILE - mean.x, x1 = x0 + (TILE - 1);
    float y0 = ty * TILE - mean.y, y1 = y0 + (TILE - 1);
    if (x0 <= 0.f && x1 >= 0.f && y0 <= 0.f && y1 >= 0.f) return true;
    float m = edge_min(t.a, t.b, t.c, x0, y0, y1);
    m = min(m, edge_min(t.a, t.b, t.c, x1, y0, y1));
    m = min(m, edge_min(t.c, t.b, t.a, y0, x0, x1));
    m = min(m, edge_min(t.c, t.b, t.a, y1, x0, x1));
    return m <= t.tau;
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-cooperative tile enumeration.  A Gaussian whose 3-sigma rectangle covers many tiles is handled by all 32
// lanes of its warp (one candidate tile per lane per step) instead of one thread looping over hundreds of tiles.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SERIAL_TILES = 6;  // rectangles up to this many candidate tiles are handled by the owning thread

struct TileJob {  // what a lane needs to enumerate the tiles of one Gaussian
    float2 xy;
    TileCull tc;
    int x0, y0, w, n;  // rect origin, width in tiles, number of candidate tiles
};
__device__ __forceinline__ TileJob bcast_job(const TileJob &j, int src) {
    TileJob o;
    const unsigned m = 0xffffffffu;
    o.xy.x = __shfl_sync(m, j.xy.x, src); o.xy.y = __shfl_sync(m, j.xy.y, src);
    o.tc.a = __shfl_sync(m, j.tc.a, src); o.tc.b = __shfl_sync(m, j.tc.b, src); o.tc.c = __shfl_sync(m, j.tc.c, src);
    o.tc.tau = __shfl_sync(m, j.tc.tau, src);
    o.tc.active = __shfl_sync(m, (int)j.tc.active, src) != 0;
    o.x0 = __shfl_sync(m, j.x0, src); o.y0 = __shfl_sync(m, j.y0, src);
    o.w = __shfl_sync(m, j.w, src); o.n = __shfl_sync(m, j.n, src);
    return o;
}

// ---------------------------------------------------------------------------------------------------------------
// K1: per (view, Gaussian) preprocess.  forward.cu:148-244
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
preprocess_kernel(int P, int V, const float *__restrict__ means3D, const float3 *__restrict__ scales, float scale_modifier,
                  const float4 *__restrict__ rotations, const float *__restrict__ opacities,
                  const float *__restrict__ cov3D_precomp, const float *__restrict__ view_matrix,
                  const float *__restrict__ proj_matrix, int W, int H, float tan_fov_x, float tan_fov_y, float focal_x,
                  float focal_y, int gx, int gy, bool exact_rect, int *__restrict__ radii, GeomView g,
                  uint32_t *__restrict__ tile_count /* bucket binning: per-tile instance histogram, or NULL */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const size_t slot = (size_t)v * P + (i < P ? i : 0);
    const float *Vm = view_matrix + 16 * v;
    const float *Pm = proj_matrix + 16 * v;

    bool visible = false;   // passes every test of the reference's preprocess
    bool big = false;       // tile count is done cooperatively
    uint32_t cnt = 0;
    float depth = 0.f;
    int radius = 0;
    float4 con_o = make_float4(0.f, 0.f, 0.f, 0.f);
    TileJob job;
    job.xy = make_float2(0.f, 0.f);
    job.tc.a = job.tc.b = job.tc.c = job.tc.tau = 0.f;
    job.tc.active = false;
    job.x0 = job.y0 = job.w = job.n = 0;

    if (i < P) do {
        const float3 p_orig = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
        const float4 p_hom = xform4x4(p_orig, Pm);
        const float p_w = 1.0f / (p_hom.w + 0.0000001f);
        const float3 p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);
        const float3 p_view = xform4x3(p_orig, Vm);
        if (p_view.z <= 0.2f) break;  // auxiliary.h:138

        float c6[6];
        if (cov3D_precomp != nullptr) {
#pragma unroll
            for (int k = 0; k < 6; k++) c6[k] = cov3D_precomp[6 * (size_t)i + k];
        } else {
            cov3d_from_scale_rot(scales[i], scale_modifier, rotations[i], c6);
            // every view writes the same six values (benign); a Gaussian culled in view 0 may be visible in view 1
#pragma unroll
            for (int k = 0; k < 6; k++) g.cov3D[6 * (size_t)i + k] = c6[k];
        }
        ProjJac pj = proj_jacobian(p_orig, focal_x, focal_y, tan_fov_x, tan_fov_y, Vm);
        M3 Vrk = vrk_of(c6);
        M3 cov2 = m3_mul(m3_mul(m3_T(pj.T), m3_T(Vrk)), pj.T);
        const float ca = cov2.m[0][0] + 0.3f, cb = cov2.m[0][1], cc = cov2.m[1][1] + 0.3f;

        const float det = (ca * cc - cb * cb);
        if (det == 0.0f) break;
        const float det_inv = 1.f / det;
        const float3 conic = make_float3(cc * det_inv, -cb * det_inv, ca * det_inv);

        const float mid = 0.5f * (ca + cc);
        const float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
        const float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
        const float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
        const float2 point_image = make_float2(ndc2pix(p_proj.x, W), ndc2pix(p_proj.y, H));
        uint2 rmin, rmax;
        get_rect(point_image, (int)my_radius, rmin, rmax, gx, gy);
        if ((rmax.x - rmin.x) * (rmax.y - rmin.y) == 0) break;

        visible = true;
        depth = p_view.z;
        radius = (int)my_radius;
        con_o = make_float4(conic.x, conic.y, conic.z, opacities[i]);
        job.xy = point_image;
        job.tc = make_cull(con_o, exact_rect);
        job.x0 = rmin.x; job.y0 = rmin.y;
        job.w = rmax.x - rmin.x;
        job.n = job.w * (rmax.y - rmin.y);
        // count the tiles that can actually receive a contribution
        if (!job.tc.active && tile_count == nullptr) {
            cnt = job.n;
        } else if (job.n <= SERIAL_TILES) {
            for (int t = 0; t < job.n; t++) {
                const int tx = job.x0 + t % job.w, ty = job.y0 + t / job.w;
                if (!tile_needed(job.tc, job.xy, tx, ty)) continue;
                cnt++;
                if (tile_count != nullptr) atomicAdd(&tile_count[(size_t)v * gx * gy + ty * gx + tx], 1u);
            }
        } else {
            big = true;
        }
    } while (0);

    unsigned pending = __ballot_sync(0xffffffffu, big);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const TileJob j = bcast_job(job, src);
        uint32_t c = 0;
        for (int t = lane; t < j.n; t += 32) {
            const int tx = j.x0 + t % j.w, ty = j.y0 + t / j.w;
            if (!tile_needed(j.tc, j.xy, tx, ty)) continue;
            c++;
            if (tile_count != nullptr) atomicAdd(&tile_count[(size_t)v * gx * gy + ty * gx + tx], 1u);
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == src) cnt = c;
    }

    if (i >= P) return;
    radii[slot] = visible ? radius : 0;
    g.tiles_touched[slot] = cnt;
    if (tile_count == nullptr) {
        g.dvals_in[slot] = (uint32_t)slot;
        // invisible Gaussians sort last in their view
        g.dkeys_in[slot] = ((unsigned long long)v << 32) | (visible ? (unsigned long long)__float_as_uint(depth) : 0xFFFFFFFFull);
    }
    if (visible) {
        g.depth[slot] = depth;
        g.xy[slot] = job.xy;
        g.conic_o[slot] = con_o;
    }
}

// writes total instance count + overflow flag after the scan (one thread)
__global__ void finish_scan_kernel(int n, GeomView g, long long capacity, long long *pinned_out) {
    uint32_t last = g.dvals_out[n - 1];
    long long total = (long long)g.offsets[n - 1] + (long long)g.tiles_touched[last];
    g.hdr->num_rendered = total;
    g.hdr->capacity = capacity;
    g.hdr->overflow = (capacity >= 0 && total > capacity) ? 1 : 0;
    g.hdr->merge_cursor = 0ull;
    g.hdr->static_prepared = 0;
    if (pinned_out) *pinned_out = total;
}

// ---------------------------------------------------------------------------------------------------------------
// K4: emit (tile id, slot) instances in depth order.  Replaces duplicateWithKeys, rasterizer_impl.cu:67-104.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
emit_kernel(int n, int P, int gx, int gy, bool exact_rect, const int *__restrict__ radii, GeomView g, BinView b) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (g.hdr->overflow) return;
    uint32_t slot = 0, cnt = 0, off = 0, tile_base = 0;
    bool big = false;
    TileJob job;
    job.xy = make_float2(0.f, 0.f);
    job.tc.a = job.tc.b = job.tc.c = job.tc.tau = 0.f;
    job.tc.active = false;
    job.x0 = job.y0 = job.w = job.n = 0;
    if (k < n) {
        slot = g.dvals_out[k];
        cnt = g.tiles_touched[slot];
    }
    if (cnt != 0) {
        off = g.offsets[k];
        const int v = slot / P;
        tile_base = (uint32_t)v * gx * gy;
        job.xy = g.xy[slot];
        uint2 rmin, rmax;
        get_rect(job.xy, radii[slot], rmin, rmax, gx, gy);
        job.tc = make_cull(g.conic_o[slot], exact_rect);
        job.x0 = rmin.x; job.y0 = rmin.y;
        job.w = rmax.x - rmin.x;
        job.n = job.w * (rmax.y - rmin.y);
        if (job.n <= SERIAL_TILES) {
            for (int t = 0; t < job.n; t++) {
                const int tx = job.x0 + t % job.w, ty = job.y0 + t / job.w;
                if (!tile_needed(job.tc, job.xy, tx, ty)) continue;
                b.tkeys_in[off] = tile_base + ty * gx + tx;
                b.tvals_in[off] = slot;
                off++;
            }
        } else {
            big = true;
        }
    }
    unsigned pending = __ballot_sync(0xffffffffu, big);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const TileJob j = bcast_job(job, src);
        uint32_t o = __shfl_sync(0xffffffffu, off, src);
        const uint32_t sl = __shfl_sync(0xffffffffu, slot, src);
        const uint32_t tb = __shfl_sync(0xffffffffu, tile_base, src);
        for (int t0 = 0; t0 < j.n; t0 += 32) {
            const int t = t0 + lane;
            const int tx = j.x0 + t % j.w, ty = j.y0 + t / j.w;
            const bool need = t < j.n && tile_needed(j.tc, j.xy, tx, ty);
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (need) {
                const uint32_t pos = o + __popc(m & ((1u << lane) - 1u));
                b.tkeys_in[pos] = tb + ty * gx + tx;
                b.tvals_in[pos] = sl;
            }
            o += __popc(m);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Bucket binning (the dynamic set of merged streams): no global sort at all.  The preprocess histograms the instances
// per tile, one CTA scans the histogram into bucket offsets, the emit below drops each instance's (depth, slot) key
// into its tile's bucket in arbitrary order, and the per-tile merge kernel sorts its bucket in shared memory.  The key
// (depth bits, slot) is a total order equal to the reference's stable (tile, depth) radix sort over index-ordered
// instances (rasterizer_impl.cu:67-104, :285-290), so the result does not depend on the order of the atomics.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int nt, const uint32_t *__restrict__ tile_count, uint2 *__restrict__ ranges, GeomHeader *__restrict__ hdr,
                 long long capacity, long long *pinned_out) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nt; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t c = t < nt ? tile_count[t] : 0u;
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = x + (warp > 0 ? s_warp[warp - 1] : 0u) + s_carry;
        if (t < nt) ranges[t] = make_uint2(incl - c, incl);
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long total = (long long)s_carry;
        hdr->num_rendered = total;
        hdr->capacity = capacity;
        hdr->overflow = (capacity >= 0 && total > capacity) ? 1 : 0;
        hdr->merge_cursor = 0ull;
        hdr->static_prepared = 0;
        if (pinned_out) *pinned_out = total;
    }
}

__global__ void __launch_bounds__(256)
emit_bucket_kernel(int n, int P, int gx, int gy, bool exact_rect, const int *__restrict__ radii, GeomView g,
                   const uint2 *__restrict__ ranges, uint32_t *__restrict__ tile_cursor, unsigned long long *__restrict__ bkeys) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (g.hdr->overflow) return;
    uint32_t slot = 0, cnt = 0, tile_base = 0;
    unsigned long long key = 0ull;
    bool big = false;
    TileJob job;
    job.xy = make_float2(0.f, 0.f);
    job.tc.a = job.tc.b = job.tc.c = job.tc.tau = 0.f;
    job.tc.active = false;
    job.x0 = job.y0 = job.w = job.n = 0;
    if (k < n) {
        slot = (uint32_t)k;
        cnt = g.tiles_touched[slot];
    }
    if (cnt != 0) {
        const int v = slot / P;
        tile_base = (uint32_t)v * gx * gy;
        job.xy = g.xy[slot];
        key = ((unsigned long long)__float_as_uint(g.depth[slot]) << 32) | slot;
        uint2 rmin, rmax;
        get_rect(job.xy, radii[slot], rmin, rmax, gx, gy);
        job.tc = make_cull(g.conic_o[slot], exact_rect);
        job.x0 = rmin.x; job.y0 = rmin.y;
        job.w = rmax.x - rmin.x;
        job.n = job.w * (rmax.y - rmin.y);
        if (job.n <= SERIAL_TILES) {
            for (int t = 0; t < job.n; t++) {
                const int tx = job.x0 + t % job.w, ty = job.y0 + t / job.w;
                if (!tile_needed(job.tc, job.xy, tx, ty)) continue;
                const uint32_t tl = tile_base + ty * gx + tx;
                bkeys[ranges[tl].x + atomicAdd(&tile_cursor[tl], 1u)] = key;
            }
        } else {
            big = true;
        }
    }
    unsigned pending = __ballot_sync(0xffffffffu, big);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const TileJob j = bcast_job(job, src);
        const unsigned long long ky = __shfl_sync(0xffffffffu, key, src);
        const uint32_t tb = __shfl_sync(0xffffffffu, tile_base, src);
        for (int t = lane; t < j.n; t += 32) {
            const int tx = j.x0 + t % j.w, ty = j.y0 + t / j.w;
            if (!tile_needed(j.tc, j.xy, tx, ty)) continue;
            const uint32_t tl = tb + ty * gx + tx;
            bkeys[ranges[tl].x + atomicAdd(&tile_cursor[tl], 1u)] = ky;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K6: pack the sorted instances into the record stream + per-tile ranges (identifyTileRanges, rasterizer_impl.cu:109-128)
// ---------------------------------------------------------------------------------------------------------------
// Sub-tile mask: bit w set <=> the 8x8 pixel patch of warp w can receive a contribution from this instance.  The blend
// kernels iterate only over the set bits of their own patch (a small splat touches 1-2 of a tile's 4 patches).
constexpr uint32_t SLOT_BITS = 24;  // slot id lives in the low 24 bits of the record's slot word when it fits
constexpr uint32_t FROZEN_BIT = 1u << 31;  // record of a Gaussian that needs no gradient (fnx_raster_args.grad_begin/end)
// Blend kernels: a warp owns an 8x8 pixel patch of the tile and every thread blends PPT = 2 pixels of it, (lx, ly) and
// (lx, ly + 4): the per-record costs that do not depend on the pixel (record loads, bit iteration, the warp reduction of
// the backward) are paid once per 64 pixels.  4 warps = 128 threads per tile.
constexpr int PPT = 2;
// Warps per blend CTA.  4: one CTA per tile sharing one staged copy of the span.  1: every 8x8 patch is its own CTA --
// it streams the tile's span by itself, needs no block barrier and stops as soon as ITS 64 pixels are done.  Measured on
// B200 (smoke workload, whole bench): 4 -> 1520 it/s, 2 -> 1414, 1 -> 1400 (at most 32 one-warp CTAs per SM, 4x the
// L2 -> shared traffic and 4x the mask building outweigh the finer scheduling), so 4 is the default.
#ifndef FNX_BLEND_WARPS
#define FNX_BLEND_WARPS 4
#endif
constexpr int BLEND_WARPS = FNX_BLEND_WARPS;
static_assert(BLEND_WARPS == 1 || BLEND_WARPS == 2 || BLEND_WARPS == 4, "1, 2 or 4 warps per blend CTA");
constexpr int BLEND_THREADS = BLEND_WARPS * 32;
constexpr int CTAS_PER_TILE = TILE_PATCHES / BLEND_WARPS;
__device__ __forceinline__ void cta_sync() {
    if (BLEND_WARPS == 1) __syncwarp(); else __syncthreads();
}
__device__ __forceinline__ bool cta_all(bool pred) {  // barrier + "pred holds in every thread of the CTA"
    if (BLEND_WARPS == 1) return __all_sync(0xffffffffu, pred);
    return __syncthreads_count(pred) == BLEND_THREADS;
}
__device__ __forceinline__ uint32_t patch_mask(const float2 xy, const float4 co, int tx, int ty, bool exact_rect) {
    const TileCull tc = make_cull(co, exact_rect);
    if (!tc.active) return (1u << TILE_PATCHES) - 1u;
    if (tc.tau < 0.f) return 0u;
    // axis-aligned extent of the ellipse q <= tau: |ux| <= sqrt(2 tau c / det), |uy| <= sqrt(2 tau a / det)
    const float det = tc.a * tc.c - tc.b * tc.b;
    const float hx = sqrtf(2.f * tc.tau * tc.c / det) * 1.0001f + 0.01f, hy = sqrtf(2.f * tc.tau * tc.a / det) * 1.0001f + 0.01f;
    // splats much larger than a patch touch (almost) every patch their bounding box overlaps: the exact test is
    // only worth its cost for small ones (any superset of the true mask is correct)
    const bool exact = fmaxf(hx, hy) < 10.f;
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < TILE_PATCHES; w++) {
        const int bx = tx * TILE + (w & 1) * PATCH, by = ty * TILE + (w >> 1) * PATCH;
        const bool bbox = (xy.x + hx >= bx) && (xy.x - hx <= bx + PATCH - 1) && (xy.y + hy >= by) && (xy.y - hy <= by + PATCH - 1);
        if (bbox && (!exact || box_needed(tc, xy, bx, by, PATCH, PATCH))) m |= 1u << w;
    }
    return m;
}

template <int C>
__global__ void __launch_bounds__(256)
pack_kernel(long long cap, int P, int gx, int ntiles, bool exact_rect, bool use_mask, int grad_begin, int grad_end,
            const float *__restrict__ colors, bool colors_per_slot, GeomView g, BinView b, uint2 *__restrict__ ranges) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long R = g.hdr->num_rendered;
    if (i >= R || i >= cap || g.hdr->overflow) return;
    uint32_t slot = b.tvals_out[i];
    const uint32_t tile = b.tkeys_out[i];
    const float2 xy = g.xy[slot];
    const float4 co = g.conic_o[slot];
    const uint32_t gi = slot % (uint32_t)P;
    const uint32_t ci = colors_per_slot ? slot : gi;   // SH colours depend on the view: one row per (view, Gaussian)
    const float depth = g.depth[slot];
    if (use_mask) {
        const uint32_t tl = tile % (uint32_t)ntiles;
        slot |= patch_mask(xy, co, tl % gx, tl / gx, exact_rect) << SLOT_BITS;
        if ((int)gi < grad_begin || (int)gi >= grad_end) slot |= FROZEN_BIT;
    }
    float4 *rec = reinterpret_cast<float4 *>(b.records + (size_t)i * RecBytes<C>::value);
    rec[0] = make_float4(xy.x, xy.y, co.x, co.y);
    if (C == 3) {
        const float c0 = colors[3 * (size_t)ci], c1 = colors[3 * (size_t)ci + 1], c2 = colors[3 * (size_t)ci + 2];
        rec[1] = make_float4(co.z, co.w, c0, c1);
        rec[2] = make_float4(c2, __uint_as_float(slot), depth, 0.f);
    } else {
        rec[1] = make_float4(co.z, co.w, colors[ci], __uint_as_float(slot));
    }
    if (i == 0) ranges[tile].x = 0;
    else {
        const uint32_t prev = b.tkeys_out[i - 1];
        if (prev != tile) {
            ranges[prev].y = (uint32_t)i;
            ranges[tile].x = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[tile].y = (uint32_t)R;
}

// ---------------------------------------------------------------------------------------------------------------
// alpha of one (pixel, record) pair -- the ONE place the blend decision is made, shared by forward and backward.
//
// Bit parity with the compiled reference (forward.cu:322-338, backward.cu:470-486) needs three things:
//  (1) `power` rounded like the reference's kernel rounds it.  Its SASS (sm_100, nvcc 12.9) evaluates
//      -0.5f * (a*dx*dx + c*dy*dy) - b*dx*dy  as  s = fma(dx, a*dx, (c*dy)*dy);  power = fma(s, -0.5, -((b*dx)*dy)),
//      i.e. the product c*dy*dy is rounded and a*dx*dx is the fused one.  Left to itself nvcc fuses the OTHER product here
//      (dx is shared by this thread's pixels and gets hoisted), which moves `power` by one ulp for some pairs and with it the
//      1/255 keep/skip decision of pairs that sit on the cut: one pixel in ~10^6 then differs by up to 0.0039 * T * |dc|
//      (measured 1.35e-3 on the 300k-Gaussian scene).  The intrinsics below pin the reference's rounding.
//  (2) alpha = o * expf(power) with the accurate expf of every pair that is KEPT, so that T, the T(1-alpha) < 1e-4
//      termination test, the T < 0.5 median-depth pick and the colour sums see the reference's values.
//  (3) no cost for that on the pairs that are skipped -- most of them: alpha >= 1/255 <=> power >= -ln(255 o) =: the record's
//      cut, computed once per staged record (warp_record_mask) with a margin of 2e-4 (about 1000 ulp of power, far beyond the
//      error of __logf and of the exponentials), so a skipped pair costs one compare and no MUFU at all.
// ---------------------------------------------------------------------------------------------------------------
constexpr float POWER_CUT_MARGIN = 2e-4f;
__device__ __forceinline__ float record_power_cut(float opacity) {
    // power below this can not reach alpha = 1/255 (opacity <= 0 -> +inf: never kept)
    return opacity > 0.f ? -__logf(255.0f * opacity) - POWER_CUT_MARGIN : __int_as_float(0x7f800000);
}
__device__ __forceinline__ bool pair_alpha(float x, float y, float a, float b, float c, float o, float pcut, float px, float py,
                                           float &dx, float &dy, float &G, float &alpha) {
    dx = x - px;
    dy = y - py;
    const float s = __fmaf_rn(dx, __fmul_rn(a, dx), __fmul_rn(__fmul_rn(c, dy), dy));
    const float power = __fmaf_rn(s, -0.5f, -__fmul_rn(__fmul_rn(b, dx), dy));
    if (power > 0.0f || power < pcut) return false;
    G = expf(power);
    alpha = min(ALPHA_MAX, o * G);
    return !(alpha < ALPHA_MIN);
}

// 1/x for x in the normal range as ONE MUFU.RCP (__fdividef(1.f, x) wraps the same instruction in a denormal / overflow guard of
// five more; the blend backward only ever inverts 1 - alpha, which lies in [0.01, 1])
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#ifndef FNX_BLEND_BATCH
#define FNX_BLEND_BATCH 128
#endif
constexpr int BATCH = BLEND_WARPS == 1 ? 64 : FNX_BLEND_BATCH;  // records per smem stage (32 one-warp CTAs per SM need <= 7 KB each)
constexpr int STAGES = 2;

// ---------------------------------------------------------------------------------------------------------------
// K7: blend forward.  One CTA (4 warps) per (tile, view); warp w owns the 8x8 pixel patch (w&1, w>>1), 2 pixels per
// thread.  forward.cu:249-373
// ---------------------------------------------------------------------------------------------------------------
// slot word of record j (slot id in the low SLOT_BITS bits, patch mask above when use_mask)
template <int C>
__device__ __forceinline__ uint32_t rec_slot_word(const float4 *rec, int j) {
    const float *f = reinterpret_cast<const float *>(rec);
    return __float_as_uint(C == 3 ? f[j * 12 + 9] : f[j * 8 + 7]);
}
// this warp's bitmask over the n (<= 128) staged records: bit set <=> the record can touch the warp's 8x8 patch
// ... and the staged records' power cuts (pair_alpha): every warp writes the entries of all n records (the four warps of a
// CTA store identical values), so after the __syncwarp a lane only ever reads what its own warp wrote.
template <int C>
__device__ __forceinline__ void warp_record_mask(const float4 *rec, int n, int warp, int lane, bool use_mask, uint32_t (&words)[BATCH / 32],
                                                 float *pcut) {
#pragma unroll
    for (int k = 0; k < BATCH / 32; k++) {
        const int r = k * 32 + lane;
        bool need = r < n;
        if (need) pcut[r] = record_power_cut(rec[r * (RecBytes<C>::value / 16) + 1].y);
        if (need && use_mask) need = ((rec_slot_word<C>(rec, r) >> (SLOT_BITS + warp)) & 1u) != 0;
        words[k] = __ballot_sync(0xffffffffu, need);
    }
    __syncwarp();
}

// bits [lo, hi) of a 32-bit word, clipped to the word
__device__ __forceinline__ uint32_t bit_range(int lo, int hi) {
    lo = max(lo, 0);
    hi = min(hi, 32);
    if (hi <= lo) return 0u;
    const uint32_t upto_hi = (hi == 32) ? 0xFFFFFFFFu : ((1u << hi) - 1u);
    return upto_hi & ~((1u << lo) - 1u);
}

// Which (view, tile) a CTA works on.  The grid is (tiles * CTAS_PER_TILE, views); CTAs are dispatched in linear block order, so
// with `tile_order` (a permutation of [0, views * tiles): fnx_raster_tile_order lists the units by decreasing work of the last
// forward) the long tiles start first and the short ones fill the tail -- longest-processing-time-first scheduling.  Without it
// the natural order.  An out-of-range entry (a caller's stale buffer) is skipped instead of followed.
__device__ __forceinline__ uint32_t work_unit(const uint32_t *__restrict__ tile_order, int nunits) {
    const uint32_t linear = blockIdx.y * (gridDim.x / CTAS_PER_TILE) + blockIdx.x / CTAS_PER_TILE;
    if (tile_order == nullptr) return linear;
    const uint32_t u = tile_order[linear];
    return u < (uint32_t)nunits ? u : 0xFFFFFFFFu;
}

// Optional per-tile state of the merged (static + dynamic) streams, all NULL for a plain forward:
//   tile_src       1: the tile has no dynamic instance and is blended straight from the static stream
//   tile_cached    (persistent, with the static stream) 1: the caller's out_color / out_depth already hold this
//                  tile's static-only render -- cameras and the frozen set are fixed, so a static-only tile renders the
//                  same pixels every iteration and is blended once, not once per iteration
//   tile_dyn_last  L = 1 + span index of the tile's last dynamic record.  Records at or behind L are all frozen: the
//                  backward needs from them only the transmittance T and the colour behind, per pixel, at L.  The
//                  forward snapshots {T, C} when it passes L and stores {T, (C_final - C) / T} in `snap`, so the
//                  backward starts at L instead of at the last contributor.
// Minimum resident CTAs per SM the forward is compiled for (register cap).  Round 1 (whole bench, it/s): unconstrained
// (77 registers, 6 CTAs/SM) 1454, 10 (48 registers) 1511, 12 (40 registers, 84 B of spills) 1512.  Round 2, with the accurate expf
// on kept pairs (more live registers): 10 (48 registers, 44 B of spills) 2016 it/s, forward 3.60 ms per 16-frame step, one frame
// alone 0.683 ms; 9 (56 registers, 12 B) 2020 / 3.41 / 0.665; 8 (64 registers) 2003 / 3.31 / 0.660.  With the restructured pair
// loop (both pixels' powers ahead of one branch): 9 (56 registers, 28 B of spills) 2215 it/s, 0.200 ms per launch, one frame alone
// 0.636 ms; 8 (64 registers, no spills) 2225 / 0.186 / 0.621.
#ifndef FNX_FWD_MIN_CTAS
#define FNX_FWD_MIN_CTAS 8
#endif
template <int C>
__global__ void __launch_bounds__(BLEND_THREADS, FNX_FWD_MIN_CTAS)
blend_fwd_kernel(int W, int H, int gx, int gy, bool use_mask, const uint32_t *__restrict__ tile_order, const char *__restrict__ records_own,
                 const char *__restrict__ records_static, const uint2 *__restrict__ ranges, const uint32_t *__restrict__ tile_src,
                 uint32_t *__restrict__ tile_cached, const uint32_t *__restrict__ tile_dyn_last, float4 *__restrict__ snap,
                 const float *__restrict__ depth_of_slot, const float *__restrict__ bg, const GeomHeader *__restrict__ hdr,
                 ImageView im, float *__restrict__ out_color, float *__restrict__ out_depth) {
    constexpr int REC = RecBytes<C>::value;
    __shared__ __align__(128) char s_rec[STAGES][BATCH * REC];
    __shared__ float s_pcut[STAGES][BATCH];
    __shared__ __align__(8) uint64_t s_bar[STAGES];

    const int ntiles = gx * gy;
    const uint32_t unit = work_unit(tile_order, ntiles * gridDim.y);
    if (unit == 0xFFFFFFFFu) return;
    const int tile = (int)(unit % (uint32_t)ntiles), v = (int)(unit / (uint32_t)ntiles);
    const size_t tslot = (size_t)v * ntiles + tile;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x % CTAS_PER_TILE) * BLEND_WARPS + (threadIdx.x >> 5);  // which 8x8 patch of the tile
    const size_t pslot = tslot * TILE_PATCHES + warp;
    const bool from_static = tile_src != nullptr && tile_src[tslot] != 0;
    if (tile_cached != nullptr) {
        // the pixels are already in place (the flags of a CTA's patches are always written together: uniform over the CTA)
        if (from_static && tile_cached[pslot - (threadIdx.x >> 5)] != 0) return;
        // (the flag is updated at the end, after the pixels are written)
    }
    const int tx = tile % gx, ty = tile / gx;
    const int px = tx * TILE + (warp & 1) * PATCH + (lane & 7);
    const int py0 = ty * TILE + (warp >> 1) * PATCH + (lane >> 3);  // pixel p of this thread is (px, py0 + 4p)
    const float pxf = (float)px;
    const size_t HW = (size_t)W * H;
    const uint32_t slot_mask = use_mask ? ((1u << SLOT_BITS) - 1u) : 0xFFFFFFFFu;
    bool inside[PPT];
    int done[PPT];  // 0 / 1 (a 32-bit flag: bool arrays compile to byte shuffles in the hot loop)
    float pyf[PPT], T[PPT], D[PPT], Cacc[PPT][C];
    float T_snap[PPT], C_snap[PPT][C];
    uint32_t last_contributor[PPT];
#pragma unroll
    for (int p = 0; p < PPT; p++) {
        inside[p] = px < W && (py0 + 4 * p) < H;
        done[p] = inside[p] ? 0 : 1;
        pyf[p] = (float)(py0 + 4 * p);
        T[p] = 1.0f;
        D[p] = DEPTH_DEFAULT;
        last_contributor[p] = 0;
        T_snap[p] = 1.0f;
#pragma unroll
        for (int ch = 0; ch < C; ch++) Cacc[p][ch] = C_snap[p][ch] = 0.f;
    }

    uint2 range = ranges[tslot];
    const char *records = from_static ? records_static : records_own;
    if (hdr->overflow) range = make_uint2(0, 0);
    const int total = (int)(range.y - range.x);
    const int nbatch = (total + BATCH - 1) / BATCH;
    // snapshot position (span index); -1: no snapshot
    const int L = (snap != nullptr && !from_static && tile_dyn_last != nullptr) ? (int)tile_dyn_last[tslot] : -1;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&s_bar[s], 1);
        mbar_fence_init();
    }
    cta_sync();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++)
            if (s < nbatch) {
                const int n = min(BATCH, total - s * BATCH);
                mbar_arrive_expect_tx(&s_bar[s], n * REC);
                bulk_g2s(s_rec[s], records + ((size_t)range.x + (size_t)s * BATCH) * REC, n * REC, &s_bar[s]);
            }
    }

    int issued = min(STAGES, nbatch);  // batches whose copy has been issued (meaningful in thread 0)
    int bi = 0;
    for (; bi < nbatch; bi++) {
        const int s = bi % STAGES;
        bool all_done = true;
#pragma unroll
        for (int p = 0; p < PPT; p++) all_done = all_done && (done[p] != 0);
        // whole tile finished?  (also orders the previous stage's reads before its buffer is refilled)
        if (cta_all(all_done)) break;
        if (threadIdx.x == 0 && bi >= 1 && bi + STAGES - 1 < nbatch) {
            // refill the stage consumed in the previous iteration
            const int nb = bi + STAGES - 1, ns = nb % STAGES;
            const int n = min(BATCH, total - nb * BATCH);
            mbar_arrive_expect_tx(&s_bar[ns], n * REC);
            bulk_g2s(s_rec[ns], records + ((size_t)range.x + (size_t)nb * BATCH) * REC, n * REC, &s_bar[ns]);
            issued = nb + 1;
        }
        mbar_wait(&s_bar[s], (bi / STAGES) & 1);
        const int n = min(BATCH, total - bi * BATCH);
        const float4 *rec = reinterpret_cast<const float4 *>(s_rec[s]);
        if (__all_sync(0xffffffffu, all_done)) continue;  // this warp's patch is finished
        uint32_t words[BATCH / 32];
        warp_record_mask<C>(rec, n, warp, lane, use_mask, words, s_pcut[s]);
        const float *pcut = s_pcut[s];
        // the batch holding the snapshot position is walked in two segments, [0, Lrel) and [Lrel, BATCH)
        const int Lrel = L - bi * BATCH;
        const int cut = (Lrel >= 0 && Lrel < BATCH) ? Lrel : BATCH;
        for (int seg = 0; seg < 2; seg++) {
            const int lo = seg == 0 ? 0 : cut, hi = seg == 0 ? cut : BATCH;
            if (seg == 1) {
                if (cut == BATCH) break;
#pragma unroll
                for (int p = 0; p < PPT; p++) {
                    T_snap[p] = T[p];
#pragma unroll
                    for (int ch = 0; ch < C; ch++) C_snap[p][ch] = Cacc[p][ch];
                }
            }
#pragma unroll
            for (int k = 0; k < BATCH / 32; k++) {
                uint32_t w = words[k] & bit_range(lo - 32 * k, hi - 32 * k);
                while (w) {
                    const int j = k * 32 + __ffs(w) - 1;
                    w &= w - 1;
                    const float4 r0 = rec[j * (REC / 16)];
                    const float4 r1 = rec[j * (REC / 16) + 1];
                    const float pc = pcut[j];
                    // Both pixels' powers first (independent chains, one divergent branch for the common "neither pixel is touched"
                    // case), then the kept pixels.  Same operations in the same order as pair_alpha.
                    const float dxs = r0.x - pxf;
                    const float adx = __fmul_rn(r0.z, dxs), bdx = __fmul_rn(r0.w, dxs);
                    float power[PPT];
                    bool keep[PPT], any_keep = false;
#pragma unroll
                    for (int p = 0; p < PPT; p++) {
                        const float dy = r0.y - pyf[p];
                        const float s = __fmaf_rn(dxs, adx, __fmul_rn(__fmul_rn(r1.x, dy), dy));
                        power[p] = __fmaf_rn(s, -0.5f, -__fmul_rn(bdx, dy));
                        keep[p] = (T[p] > 0.0f) && !(power[p] > 0.0f) && !(power[p] < pc);
                        any_keep = any_keep || keep[p];
                    }
                    if (!any_keep) continue;
                    float4 r2 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (C == 3) r2 = rec[j * 3 + 2];
                    const uint32_t pos = (uint32_t)(bi * BATCH + j + 1);
#pragma unroll
                    for (int p = 0; p < PPT; p++) {
                        if (!keep[p]) continue;
                        const float alpha = min(ALPHA_MAX, r1.y * expf(power[p]));
                        if (alpha < ALPHA_MIN) continue;
                        const float test_T = T[p] * (1 - alpha);
                        if (test_T < T_EPS) {
                            T[p] = -T[p];   // finished: the sign is the flag (see the declaration of T)
                            continue;
                        }
                        Cacc[p][0] += r1.z * alpha * T[p];
                        if (C == 3) {
                            Cacc[p][1 % C] += r1.w * alpha * T[p];
                            Cacc[p][2 % C] += r2.x * alpha * T[p];
                            if (T[p] > 0.5f && test_T < 0.5) D[p] = r2.z;
                        } else {
                            if (T[p] > 0.5f && test_T < 0.5) D[p] = depth_of_slot[__float_as_uint(r1.w) & slot_mask];
                        }
                        T[p] = test_T;
                        last_contributor[p] = pos;
                    }
                }
            }
        }
    }

    // early exit: bulk copies still in flight must land before this CTA's shared memory is released
    if (threadIdx.x == 0)
        for (int nb = bi; nb < issued; nb++) mbar_wait(&s_bar[nb % STAGES], (nb / STAGES) & 1);

    // tile-wide max of last_contributor: where the backward starts
    uint32_t wl = 0;
#pragma unroll
    for (int p = 0; p < PPT; p++) wl = max(wl, last_contributor[p]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
    if (lane == 0) im.tile_last[pslot] = wl;
#pragma unroll
    for (int p = 0; p < PPT; p++) {
        if (!inside[p]) continue;
        const size_t pix = (size_t)(py0 + 4 * p) * W + px;
        const float Tf = fabsf(T[p]);
        im.final_T[(size_t)v * HW + pix] = Tf;
        im.n_contrib[(size_t)v * HW + pix] = last_contributor[p];
#pragma unroll
        for (int ch = 0; ch < C; ch++) out_color[((size_t)v * C + ch) * HW + pix] = Cacc[p][ch] + Tf * bg[ch];
        out_depth[(size_t)v * HW + pix] = D[p];
        if (L >= 0) {
            // colour accumulated behind the snapshot position, normalised by the transmittance there: what the
            // reference's back-to-front recursion (backward.cu:488-496) holds in accum_rec when it arrives at L
            const float Ts = fabsf(T_snap[p]);
            const float inv = 1.f / Ts;
            snap[(size_t)v * HW + pix] = make_float4(Ts, (Cacc[p][0] - C_snap[p][0]) * inv, (Cacc[p][1 % C] - C_snap[p][1 % C]) * inv,
                                                     (Cacc[p][2 % C] - C_snap[p][2 % C]) * inv);
        }
    }
    if (tile_cached != nullptr) {
        cta_sync();  // every pixel of the CTA's patches is written before their flags say so
        if (lane == 0) tile_cached[pslot] = from_static ? 1u : 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Warp reduction of N per-lane values with a value-splitting butterfly.  A plain butterfly costs 5 shuffles per
// value (45 for N = 9) and shuffles run on the LSU pipe (1 warp-instruction / clk / SM) -- the first profile of the
// backward showed that pipe at 74 % with shuffles as ~85 % of its traffic.  Here, at xor-mask m, the lanes with bit m
// clear keep the first half of their values and the others the second half; each lane sends the half it drops and
// adds the half it receives, so the value count halves every round: 5+3+2+1+1 = 12 shuffles for N = 9 (9 for N = 7).
// Afterwards every lane holds ONE fully reduced value; split_index() tells which (or -1), split_dup_mask() which
// lane bits hold duplicates.
// ---------------------------------------------------------------------------------------------------------------
template <int N, int M>
struct SplitReduce {
    static __device__ __forceinline__ float run(const float *v, int lane) {
        if constexpr (M == 0) {
            return v[0];
        } else if constexpr (N == 1) {
            float w[1] = {v[0] + __shfl_xor_sync(0xffffffffu, v[0], M)};
            return SplitReduce<1, M / 2>::run(w, lane);
        } else {
            constexpr int H0 = (N + 1) / 2;
            const bool bit = (lane & M) != 0;
            float w[H0];
#pragma unroll
            for (int k = 0; k < H0; k++) {
                const float a = v[k];
                const float b = (H0 + k < N) ? v[H0 + k] : 0.f;
                const float send = bit ? a : b, keep = bit ? b : a;
                w[k] = keep + __shfl_xor_sync(0xffffffffu, send, M);
            }
            return SplitReduce<H0, M / 2>::run(w, lane);
        }
    }
    // index (into the original N values) of the value this lane ends up holding, or -1
    static __device__ __forceinline__ int index(int lane) {
        if constexpr (M == 0) {
            return 0;
        } else if constexpr (N == 1) {
            return SplitReduce<1, M / 2>::index(lane);
        } else {
            constexpr int H0 = (N + 1) / 2;
            const int sub = SplitReduce<H0, M / 2>::index(lane);
            if (sub < 0) return -1;
            const int orig = (lane & M) ? H0 + sub : sub;
            return orig < N ? orig : -1;
        }
    }
    static __device__ __forceinline__ constexpr int dup_mask() {
        if constexpr (M == 0) return 0;
        else if constexpr (N == 1) return M | SplitReduce<1, M / 2>::dup_mask();
        else return SplitReduce<(N + 1) / 2, M / 2>::dup_mask();
    }
};

// ---------------------------------------------------------------------------------------------------------------
// K8: blend backward.  backward.cu:384-536.  Walks the tile's records back to front starting at the last record any
// pixel of the tile used; per record the 32 pixels of a warp are reduced with the value-splitting butterfly above,
// then the 6+C lanes that hold a result issue ONE reduction instruction into the (view, Gaussian) accumulator row
// (their 6+C addresses are contiguous: 1-2 L2 sectors).
// ---------------------------------------------------------------------------------------------------------------
// Register cap of the backward (see FNX_FWD_MIN_CTAS).  Measured: 7 CTAs/SM (72 registers) 1989 it/s, 8 (64 registers) 1996,
// 9 (56 registers, 48 B of spills) slower.
#ifndef FNX_BWD_MIN_CTAS
#define FNX_BWD_MIN_CTAS 8
#endif
template <int C>
__global__ void __launch_bounds__(BLEND_THREADS, FNX_BWD_MIN_CTAS)
blend_bwd_kernel(int W, int H, int gx, int gy, bool use_mask, const uint32_t *__restrict__ tile_order, const char *__restrict__ records_own,
                 const char *__restrict__ records_static, const uint2 *__restrict__ ranges, const uint32_t *__restrict__ tile_src,
                 const uint32_t *__restrict__ tile_dyn_last, const uint32_t *__restrict__ tile_dyn_first, const float4 *__restrict__ snap,
                 const float *__restrict__ bg, const GeomHeader *__restrict__ hdr, ImageView im,
                 const float *__restrict__ dL_dpixels, float *__restrict__ accum) {
    constexpr int REC = RecBytes<C>::value;
    constexpr int ACC = AccFloats<C>::value;
    __shared__ __align__(128) char s_rec[STAGES][BATCH * REC];
    __shared__ float s_pcut[STAGES][BATCH];
    __shared__ __align__(8) uint64_t s_bar[STAGES];

    const int ntiles = gx * gy;
    const uint32_t unit = work_unit(tile_order, ntiles * gridDim.y);
    if (unit == 0xFFFFFFFFu) return;
    const int tile = (int)(unit % (uint32_t)ntiles), v = (int)(unit / (uint32_t)ntiles);
    const int tx = tile % gx, ty = tile / gx;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x % CTAS_PER_TILE) * BLEND_WARPS + (threadIdx.x >> 5);  // which 8x8 patch of the tile
    const int px = tx * TILE + (warp & 1) * PATCH + (lane & 7);
    const int py0 = ty * TILE + (warp >> 1) * PATCH + (lane >> 3);
    const float pxf = (float)px;
    const size_t HW = (size_t)W * H;
    const uint32_t slot_mask = use_mask ? ((1u << SLOT_BITS) - 1u) : 0xFFFFFFFFu;

    const size_t tslot = (size_t)v * ntiles + tile;
    // merged streams: a tile blended straight from the static stream holds frozen records only -- nothing to do
    if (tile_src != nullptr && tile_src[tslot] != 0) return;
    const uint2 range = ranges[tslot];
    const char *records = records_own;
    (void)records_static;
    // records [0,total) of the span matter to this CTA's patches
    int total = 0;
#pragma unroll
    for (int w = 0; w < BLEND_WARPS; w++)
        total = max(total, (int)im.tile_last[tslot * TILE_PATCHES + (blockIdx.x % CTAS_PER_TILE) * BLEND_WARPS + w]);
    // merged streams: everything at or behind L is frozen; the forward left {T, colour behind} at L in `snap`
    const int L = (snap != nullptr && tile_dyn_last != nullptr) ? (int)tile_dyn_last[tslot] : 0x7FFFFFFF;
    total = min(total, L);
    // ... and everything in FRONT of the tile's first dynamic record F is frozen too.  The back-to-front walk exists to hand every
    // record with a gradient its transmittance and the colour behind it; once it has passed the nearest one, the records further
    // to the front (about half of a tile's static span in the bench scenes: the near side of the background shell) change
    // nothing that is written.
    const int F = (tile_dyn_first != nullptr && snap != nullptr) ? min((int)tile_dyn_first[tslot], total) : 0;
    if (hdr->overflow) total = 0;
    if (total - F <= 0) return;
    const int nbatch = (total - F + BATCH - 1) / BATCH;
    // batch bi (processing order) covers span indices [lo, hi) with hi = total - bi*BATCH, lo = max(F, hi - BATCH)
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&s_bar[s], 1);
        mbar_fence_init();
    }
    cta_sync();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++)
            if (s < nbatch) {
                const int hi = total - s * BATCH, lo = max(F, hi - BATCH), n = hi - lo;
                mbar_arrive_expect_tx(&s_bar[s], n * REC);
                bulk_g2s(s_rec[s], records + ((size_t)range.x + lo) * REC, n * REC, &s_bar[s]);
            }
    }

    constexpr int NV = 6 + C;  // {dmean2D.xy, dconic.xyw, dopacity, dcolor[C]} == accumulator row order
    using Red = SplitReduce<NV, 16>;
    const int my_val = ((lane & Red::dup_mask()) == 0) ? Red::index(lane) : -1;  // which reduced value this lane publishes

    float pyf[PPT], T_final[PPT], T[PPT], last_alpha[PPT], bg_dot_dpixel[PPT];
    // The reference tracks the colour accumulated behind the current record per channel (accum_rec, last_color:
    // backward.cu:488-496) and dots (c - accum_rec) with dL/dpixel.  dL/dpixel is constant along the walk and the
    // recursion is linear, so the DOT PRODUCTS are tracked instead: accum_dot = accum_rec . dL/dpixel, last_dot = last_color .
    // dL/dpixel -- 2 FMAs instead of 2C per record and 2 registers instead of 2C per pixel.
    float accum_dot[PPT], last_dot[PPT], dL_dpixel[PPT][C];
    int last_contributor[PPT];
    int warp_last = 0;  // records at or beyond this position touch no pixel of this warp
#pragma unroll
    for (int p = 0; p < PPT; p++) {
        const int py = py0 + 4 * p;
        const bool inside = px < W && py < H;
        const size_t pix = (size_t)py * W + px;
        pyf[p] = (float)py;
        T_final[p] = inside ? im.final_T[(size_t)v * HW + pix] : 0.f;
        T[p] = T_final[p];
        last_contributor[p] = inside ? (int)im.n_contrib[(size_t)v * HW + pix] : 0;
        last_alpha[p] = 0.f;
        bg_dot_dpixel[p] = 0.f;
        accum_dot[p] = 0.f;
        last_dot[p] = 0.f;
#pragma unroll
        for (int ch = 0; ch < C; ch++) {
            dL_dpixel[p][ch] = inside ? dL_dpixels[((size_t)v * C + ch) * HW + pix] : 0.f;
            bg_dot_dpixel[p] += bg[ch] * dL_dpixel[p][ch];
        }
        if (last_contributor[p] > L) {  // this pixel blended frozen records behind L: resume from the forward's snapshot
            const float4 sn = snap[(size_t)v * HW + pix];
            T[p] = sn.x;
            accum_dot[p] = sn.y * dL_dpixel[p][0];
            if (C == 3) accum_dot[p] += sn.z * dL_dpixel[p][1 % C] + sn.w * dL_dpixel[p][2 % C];
            last_contributor[p] = L;
        }
        warp_last = max(warp_last, last_contributor[p]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
    // constant factor of the value this lane publishes: dmean2D.x: -(0.5 W), dmean2D.y: -(0.5 H) (backward.cu:444-445,
    // 524-525), dconic.{x,y,w}: -0.5, dopacity and dcolour: 1
    const float post_scale = my_val == 0 ? -0.5f * W : (my_val == 1 ? -0.5f * H : (my_val >= 2 && my_val <= 4 ? -0.5f : 1.0f));

    for (int bi = 0; bi < nbatch; bi++) {
        const int s = bi % STAGES;
        if (bi >= 1) {
            cta_sync();  // everyone finished reading the stage that is about to be refilled
            if (threadIdx.x == 0 && bi + STAGES - 1 < nbatch) {
                const int nb = bi + STAGES - 1, ns = nb % STAGES;
                const int hi = total - nb * BATCH, lo = max(F, hi - BATCH), n = hi - lo;
                mbar_arrive_expect_tx(&s_bar[ns], n * REC);
                bulk_g2s(s_rec[ns], records + ((size_t)range.x + lo) * REC, n * REC, &s_bar[ns]);
            }
        }
        mbar_wait(&s_bar[s], (bi / STAGES) & 1);
        const int hi = total - bi * BATCH, lo = max(F, hi - BATCH), n = hi - lo;
        if (lo >= warp_last) continue;  // nothing in this batch reaches this warp's pixels
        const float4 *rec = reinterpret_cast<const float4 *>(s_rec[s]);
        uint32_t words[BATCH / 32];
        warp_record_mask<C>(rec, min(n, warp_last - lo), warp, lane, use_mask, words, s_pcut[s]);
        const float *pcut = s_pcut[s];
#pragma unroll
        for (int k = BATCH / 32 - 1; k >= 0; k--) {
            uint32_t w = words[k];
            while (w) {
                const int bit = 31 - __clz(w);
                w &= ~(1u << bit);
                const int j = k * 32 + bit;
                const int idx = lo + j;  // 0-based position in the tile's span
                const float4 r0 = rec[j * (REC / 16)];
                const float4 r1 = rec[j * (REC / 16) + 1];
                const float pc = pcut[j];
                // both pixels' powers first (independent chains), ONE warp vote for the common "no pixel of the patch is touched"
                // case, then the exponentials of the pixels that passed -- the operations and their order are pair_alpha's
                const float dx = r0.x - pxf;
                const float adx = __fmul_rn(r0.z, dx), bdx = __fmul_rn(r0.w, dx);
                float dy[PPT], G[PPT], alpha[PPT];
                bool contrib[PPT], any = false;
#pragma unroll
                for (int p = 0; p < PPT; p++) {
                    dy[p] = r0.y - pyf[p];
                    const float s = __fmaf_rn(dx, adx, __fmul_rn(__fmul_rn(r1.x, dy[p]), dy[p]));
                    G[p] = __fmaf_rn(s, -0.5f, -__fmul_rn(bdx, dy[p]));   // the power, for now
                    contrib[p] = (idx < last_contributor[p]) && !(G[p] > 0.0f) && !(G[p] < pc);
                    any = any || contrib[p];
                }
                if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
                for (int p = 0; p < PPT; p++) {
                    alpha[p] = 0.f;
                    if (contrib[p]) {
                        G[p] = expf(G[p]);
                        alpha[p] = min(ALPHA_MAX, r1.y * G[p]);
                        contrib[p] = !(alpha[p] < ALPHA_MIN);
                    }
                }

                float vals[NV];
#pragma unroll
                for (int q = 0; q < NV; q++) vals[q] = 0.f;
                uint32_t slot_word;
                float col[C];
                if (C == 3) {
                    const float4 r2 = rec[j * 3 + 2];
                    col[0] = r1.z; col[1 % C] = r1.w; col[2 % C] = r2.x;
                    slot_word = __float_as_uint(r2.y);
                } else {
                    col[0] = r1.z;
                    slot_word = __float_as_uint(r1.w);
                }
                const uint32_t slot_bits = slot_word & slot_mask;
                if (use_mask && (slot_word & FROZEN_BIT)) {
                    // frozen Gaussian: it occludes (T, accumulated colour behind) but nobody wants its gradient
#pragma unroll
                    for (int p = 0; p < PPT; p++) {
                        if (!contrib[p]) continue;
                        T[p] = T[p] * rcp_approx(1.f - alpha[p]);
                        accum_dot[p] = last_alpha[p] * last_dot[p] + (1.f - last_alpha[p]) * accum_dot[p];
                        float cd = 0.f;
#pragma unroll
                        for (int ch = 0; ch < C; ch++) cd += col[ch] * dL_dpixel[p][ch];
                        last_dot[p] = cd;
                        last_alpha[p] = alpha[p];
                    }
                    continue;
                }
#pragma unroll
                for (int p = 0; p < PPT; p++) {
                    if (!contrib[p]) continue;
                    // one approximate reciprocal (MUFU.RCP, 1-alpha is in [0.01, 1)) replaces the reference's two IEEE
                    // divisions (backward.cu:484,513): ~2 ulp per step, far inside the gradient tolerance
                    const float inv_1ma = rcp_approx(1.f - alpha[p]);
                    T[p] = T[p] * inv_1ma;
                    const float dchannel_dcolor = alpha[p] * T[p];
                    accum_dot[p] = last_alpha[p] * last_dot[p] + (1.f - last_alpha[p]) * accum_dot[p];
                    float cd = 0.f;
#pragma unroll
                    for (int ch = 0; ch < C; ch++) {
                        cd += col[ch] * dL_dpixel[p][ch];
                        vals[6 + ch] += dchannel_dcolor * dL_dpixel[p][ch];
                    }
                    last_dot[p] = cd;
                    float dL_dalpha = (cd - accum_dot[p]) * T[p];
                    last_alpha[p] = alpha[p];
                    dL_dalpha += (-T_final[p] * inv_1ma) * bg_dot_dpixel[p];
                    // geometric terms without their constant factors (-0.5 W, -0.5 H, -0.5: applied once per record after
                    // the warp reduction, see `post_scale`): A = dL/dG * G * dx, B = dL/dG * G * dy
                    const float gG = r1.y * dL_dalpha * G[p];
                    const float A = gG * dx, B = gG * dy[p];
                    vals[0] += A * r0.z + B * r0.w;      // -(dG/ddelx) dL/dG
                    vals[1] += B * r1.x + A * r0.w;      // -(dG/ddely) dL/dG
                    vals[2] += A * dx;
                    vals[3] += A * dy[p];
                    vals[4] += B * dy[p];
                    vals[5] += G[p] * dL_dalpha;
                }
                const float red = Red::run(vals, lane) * post_scale;
                if (my_val >= 0) red_add(accum + (size_t)slot_bits * ACC + my_val, red);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K9: per-Gaussian backward: what computeCov2DCUDA (backward.cu:137-263), preprocessCUDA's backward (:332-381) and
// computeCov3D's backward (:267-327) compute, in one pass over the V views.  The chain rule itself lives in geom_grad.cuh
// (matrix form, host-compilable: tests/test_geom_grad_host.py runs it on the CPU against the oracle).
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256)
geom_bwd_kernel(int P, int V, const float *__restrict__ means3D, const float3 *__restrict__ scales, float scale_modifier,
                const float4 *__restrict__ rotations, const float *__restrict__ cov3D_precomp,
                const float *__restrict__ view_matrix, const float *__restrict__ proj_matrix, int W, int H,
                float tan_fov_x, float tan_fov_y, float focal_x, float focal_y, const int *__restrict__ radii,
                const float *__restrict__ cov3D_geom, const float *__restrict__ accum, int grad_begin, int grad_end,
                fnx_raster_grads out) {
    constexpr int ACC = AccFloats<C>::value;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    if (i < grad_begin || i >= grad_end) {  // frozen: zero rows
        if (out.dL_dmeans3D) { out.dL_dmeans3D[3 * i] = 0.f; out.dL_dmeans3D[3 * i + 1] = 0.f; out.dL_dmeans3D[3 * i + 2] = 0.f; }
        if (out.dL_dmeans2D)
            for (int v = 0; v < V; v++) {
                const size_t sl = (size_t)v * P + i;
                out.dL_dmeans2D[3 * sl] = 0.f; out.dL_dmeans2D[3 * sl + 1] = 0.f; out.dL_dmeans2D[3 * sl + 2] = 0.f;
            }
        if (out.dL_dopacity) out.dL_dopacity[i] = 0.f;
        if (out.dL_dcolors)
            for (int ch = 0; ch < C; ch++) out.dL_dcolors[(size_t)i * C + ch] = 0.f;
        if (out.dL_dcov3D)
            for (int k = 0; k < 6; k++) out.dL_dcov3D[6 * (size_t)i + k] = 0.f;
        if (scales != nullptr) {
            if (out.dL_dscales) { out.dL_dscales[3 * i] = 0.f; out.dL_dscales[3 * i + 1] = 0.f; out.dL_dscales[3 * i + 2] = 0.f; }
            if (out.dL_drotations) *reinterpret_cast<float4 *>(out.dL_drotations + 4 * (size_t)i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    namespace gg = geomgrad;
    const float mean[3] = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
    const float *cov3D = (cov3D_precomp != nullptr ? cov3D_precomp : cov3D_geom) + 6 * (size_t)i;
    float c6[6];
    bool any_visible = false;
    for (int v = 0; v < V; v++) any_visible |= radii[(size_t)v * P + i] > 0;
#pragma unroll
    for (int k = 0; k < 6; k++) c6[k] = any_visible ? cov3D[k] : 0.f;
    float Sig[3][3];
    gg::sym3_from6(c6, Sig);

    float d_mean[3] = {0.f, 0.f, 0.f};
    float d_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float d_op = 0.f, d_col[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) d_col[ch] = 0.f;

    for (int v = 0; v < V; v++) {
        const size_t slot = (size_t)v * P + i;
        const float *row = accum + slot * ACC;
        const float4 a0 = *reinterpret_cast<const float4 *>(row);   // {dmean2D.x, dmean2D.y, dconic.xx, dconic.xy}
        const float4 a1 = *reinterpret_cast<const float4 *>(row + 4);   // {dconic.yy, dopacity, dcol0, dcol1}
        if (out.dL_dmeans2D) {
            out.dL_dmeans2D[3 * slot] = a0.x;
            out.dL_dmeans2D[3 * slot + 1] = a0.y;
            out.dL_dmeans2D[3 * slot + 2] = 0.f;
        }
        d_op += a1.y;
        d_col[0] += a1.z;
        if (C == 3) {
            d_col[1 % C] += a1.w;
            d_col[2 % C] += row[8];
        }
        if (!(radii[slot] > 0)) continue;
        // screen covariance -> 3-D covariance and, through A = J R, the mean; screen mean -> mean (geom_grad.cuh)
        float A[2][3], t[3], dA[2][3];
        bool x_free, y_free;
        gg::view_jacobian(mean, view_matrix + 16 * v, focal_x, focal_y, tan_fov_x, tan_fov_y, A, t, x_free, y_free);
        gg::screen_cov_backward(A, Sig, a0.z, a0.w, a1.x, d_cov, dA);
        gg::perspective_backward(dA, view_matrix + 16 * v, t, focal_x, focal_y, x_free, y_free, d_mean);
        gg::ndc_backward(proj_matrix + 16 * v, mean, a0.x, a0.y, d_mean);
    }

    if (out.dL_dmeans3D) {
        out.dL_dmeans3D[3 * i] = d_mean[0]; out.dL_dmeans3D[3 * i + 1] = d_mean[1]; out.dL_dmeans3D[3 * i + 2] = d_mean[2];
    }
    if (out.dL_dopacity) out.dL_dopacity[i] = d_op;
    if (out.dL_dcolors) {
#pragma unroll
        for (int ch = 0; ch < C; ch++) out.dL_dcolors[(size_t)i * C + ch] = d_col[ch];
    }
    if (out.dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; k++) out.dL_dcov3D[6 * (size_t)i + k] = d_cov[k];
    }
    if (scales != nullptr && (out.dL_dscales || out.dL_drotations)) {
        float ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
        if (any_visible) {
            const float4 q4 = rotations[i];
            const float3 sc = scales[i];
            const float q[4] = {q4.x, q4.y, q4.z, q4.w};
            const float s[3] = {scale_modifier * sc.x, scale_modifier * sc.y, scale_modifier * sc.z};
            gg::cov3d_backward(s, q, d_cov, ds, dq);
        }
        if (out.dL_dscales) {
            out.dL_dscales[3 * i] = ds[0]; out.dL_dscales[3 * i + 1] = ds[1]; out.dL_dscales[3 * i + 2] = ds[2];
        }
        if (out.dL_drotations) *reinterpret_cast<float4 *>(out.dL_drotations + 4 * (size_t)i) = make_float4(dq[0], dq[1], dq[2], dq[3]);
    }
}

__global__ void mark_visible_kernel(int P, const float *__restrict__ means3D, const float *__restrict__ view_matrix,
                                    uint8_t *__restrict__ present) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float3 p = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
    const float3 pv = xform4x3(p, view_matrix);
    present[i] = !(pv.z <= 0.2f);
}

// ---------------------------------------------------------------------------------------------------------------
// Spherical-harmonics colours (R3/cuda_rasterizer/forward.cu:20-67, backward.cu:20-132).  Dead on FluidNexus' own pipes (they
// pass colors_precomp, FD/renderer/pipe_fluid.py:107-118), kept so that the drop-in accepts everything the reference module
// accepts.  colour(view, Gaussian) = max(0, 0.5 + sum_k B_k(dir) sh_k), dir = normalize(mean - campos), real SH basis up to
// degree 3 in the ordering and sign convention of the 3DGS code base.  Implemented as two small kernels next to the existing
// pipeline: the forward fills a per-(view, Gaussian) colour array (+ which channels were clamped) that the pack kernel reads
// instead of colors_precomp; the backward turns the per-(view, Gaussian) colour gradients into dL/dsh and adds the view-direction
// term to dL/dmeans3D.  The basis and its gradient are written out as polynomials in (x, y, z) and their partial derivatives.
// ---------------------------------------------------------------------------------------------------------------
struct ShView {
    float *color;        // [V*P, 3]
    uint32_t *clamped;   // [V*P] bit c: channel c was clamped at 0 (no gradient through it)
};
static size_t sh_extra_bytes(int P, int V) { return align_up((size_t)P * V * 12, 256) + align_up((size_t)P * V * 4, 256) + 512; }
static char *geom_end(const GeomView &g) { return (char *)align_up((size_t)((char *)g.cub_temp + g.cub_temp_bytes), 256); }
static ShView sh_view(char *p, int P, int V) {
    ShView s;
    s.color = carve<float>(p, (size_t)P * V * 3);
    s.clamped = carve<uint32_t>(p, (size_t)P * V);
    return s;
}

constexpr float SH_C0 = 0.28209479177387814f, SH_C1 = 0.4886025119029199f;
__constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f};
__constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f, -0.4570457994644658f,
                               1.445305721320277f, -0.5900435899266435f};

// B_k(x, y, z) for k < (deg+1)^2 and, if wanted, its partial derivatives
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float *b, float *bx, float *by, float *bz) {
    const bool grad = bx != nullptr;
    b[0] = SH_C0;
    if (grad) bx[0] = by[0] = bz[0] = 0.f;
    if (deg < 1) return;
    b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
    if (grad) {
        bx[1] = 0.f; by[1] = -SH_C1; bz[1] = 0.f;
        bx[2] = 0.f; by[2] = 0.f; bz[2] = SH_C1;
        bx[3] = -SH_C1; by[3] = 0.f; bz[3] = 0.f;
    }
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = SH_C2[0] * xy; b[5] = SH_C2[1] * yz; b[6] = SH_C2[2] * (2.f * zz - xx - yy); b[7] = SH_C2[3] * xz; b[8] = SH_C2[4] * (xx - yy);
    if (grad) {
        bx[4] = SH_C2[0] * y; by[4] = SH_C2[0] * x; bz[4] = 0.f;
        bx[5] = 0.f; by[5] = SH_C2[1] * z; bz[5] = SH_C2[1] * y;
        bx[6] = -2.f * SH_C2[2] * x; by[6] = -2.f * SH_C2[2] * y; bz[6] = 4.f * SH_C2[2] * z;
        bx[7] = SH_C2[3] * z; by[7] = 0.f; bz[7] = SH_C2[3] * x;
        bx[8] = 2.f * SH_C2[4] * x; by[8] = -2.f * SH_C2[4] * y; bz[8] = 0.f;
    }
    if (deg < 3) return;
    b[9] = SH_C3[0] * y * (3.f * xx - yy);
    b[10] = SH_C3[1] * xy * z;
    b[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
    b[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
    b[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
    b[14] = SH_C3[5] * z * (xx - yy);
    b[15] = SH_C3[6] * x * (xx - 3.f * yy);
    if (grad) {
        bx[9] = SH_C3[0] * 6.f * xy; by[9] = SH_C3[0] * 3.f * (xx - yy); bz[9] = 0.f;
        bx[10] = SH_C3[1] * yz; by[10] = SH_C3[1] * xz; bz[10] = SH_C3[1] * xy;
        bx[11] = SH_C3[2] * (-2.f * xy); by[11] = SH_C3[2] * (4.f * zz - xx - 3.f * yy); bz[11] = SH_C3[2] * 8.f * yz;
        bx[12] = SH_C3[3] * (-6.f * xz); by[12] = SH_C3[3] * (-6.f * yz); bz[12] = SH_C3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
        bx[13] = SH_C3[4] * (4.f * zz - 3.f * xx - yy); by[13] = SH_C3[4] * (-2.f * xy); bz[13] = SH_C3[4] * 8.f * xz;
        bx[14] = SH_C3[5] * 2.f * xz; by[14] = SH_C3[5] * (-2.f * yz); bz[14] = SH_C3[5] * (xx - yy);
        bx[15] = SH_C3[6] * 3.f * (xx - yy); by[15] = SH_C3[6] * (-6.f * xy); bz[15] = 0.f;
    }
}

__global__ void __launch_bounds__(256)
sh_color_kernel(int P, int V, int deg, int M, const float *__restrict__ means3D, const float *__restrict__ campos,
                const float *__restrict__ sh, const int *__restrict__ radii, ShView out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (i >= P) return;
    const size_t slot = (size_t)v * P + i;
    if (radii[slot] <= 0) { out.clamped[slot] = 0u; return; }
    float dx = means3D[3 * i] - campos[3 * v], dy = means3D[3 * i + 1] - campos[3 * v + 1], dz = means3D[3 * i + 2] - campos[3 * v + 2];
    const float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
    dx *= inv; dy *= inv; dz *= inv;
    float b[16];
    sh_basis(deg, dx, dy, dz, b, nullptr, nullptr, nullptr);
    const int nb = (deg + 1) * (deg + 1);
    const float *c = sh + (size_t)i * M * 3;
    float rgb[3] = {0.5f, 0.5f, 0.5f};
    for (int k = 0; k < nb; k++) {
        rgb[0] += b[k] * c[3 * k]; rgb[1] += b[k] * c[3 * k + 1]; rgb[2] += b[k] * c[3 * k + 2];
    }
    uint32_t cl = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        if (rgb[ch] < 0.f) { cl |= 1u << ch; rgb[ch] = 0.f; }
        out.color[3 * slot + ch] = rgb[ch];
    }
    out.clamped[slot] = cl;
}

// per Gaussian, summed over the views: dL/dsh [P, M, 3] (overwritten) and the SH term ADDED to dL/dmeans3D
__global__ void __launch_bounds__(256)
sh_bwd_kernel(int P, int V, int deg, int M, const float *__restrict__ means3D, const float *__restrict__ campos,
              const float *__restrict__ sh, const int *__restrict__ radii, ShView f, const float *__restrict__ accum,
              float *__restrict__ dL_dsh, float *__restrict__ dL_dmeans3D) {
    constexpr int ACC = AccFloats<3>::value;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int nb = (deg + 1) * (deg + 1);
    const float *c = sh + (size_t)i * M * 3;
    float gsh[16][3];
    for (int k = 0; k < 16; k++) gsh[k][0] = gsh[k][1] = gsh[k][2] = 0.f;
    float gm[3] = {0.f, 0.f, 0.f};
    for (int v = 0; v < V; v++) {
        const size_t slot = (size_t)v * P + i;
        if (radii[slot] <= 0) continue;
        const float *row = accum + slot * ACC;
        const uint32_t cl = f.clamped[slot];
        const float g[3] = {(cl & 1u) ? 0.f : row[6], (cl & 2u) ? 0.f : row[7], (cl & 4u) ? 0.f : row[8]};
        const float vx = means3D[3 * i] - campos[3 * v], vy = means3D[3 * i + 1] - campos[3 * v + 1], vz = means3D[3 * i + 2] - campos[3 * v + 2];
        const float len2 = vx * vx + vy * vy + vz * vz, inv = rsqrtf(len2);
        const float x = vx * inv, y = vy * inv, z = vz * inv;
        float b[16], bx[16], by[16], bz[16];
        sh_basis(deg, x, y, z, b, bx, by, bz);
        float ddx = 0.f, ddy = 0.f, ddz = 0.f;   // dL/ddir
        for (int k = 0; k < nb; k++) {
            const float w = c[3 * k] * g[0] + c[3 * k + 1] * g[1] + c[3 * k + 2] * g[2];
            ddx += bx[k] * w; ddy += by[k] * w; ddz += bz[k] * w;
            gsh[k][0] += b[k] * g[0]; gsh[k][1] += b[k] * g[1]; gsh[k][2] += b[k] * g[2];
        }
        // through dir = v / |v|:  (I - dir dir^T) / |v|
        const float dot = x * ddx + y * ddy + z * ddz;
        gm[0] += (ddx - x * dot) * inv; gm[1] += (ddy - y * dot) * inv; gm[2] += (ddz - z * dot) * inv;
    }
    if (dL_dsh != nullptr)
        for (int k = 0; k < M; k++)
            for (int ch = 0; ch < 3; ch++) dL_dsh[((size_t)i * M + k) * 3 + ch] = k < nb ? gsh[k][ch] : 0.f;
    if (dL_dmeans3D != nullptr) {
        dL_dmeans3D[3 * i] += gm[0]; dL_dmeans3D[3 * i + 1] += gm[1]; dL_dmeans3D[3 * i + 2] += gm[2];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// Ring of pinned int64 slots + events for the instance-count read-back of forwards that are allowed to wait for it.  One ring
// per DEVICE, shared by all host threads (a forward and the fnx_raster_check that goes with it may run on different threads --
// PyTorch's autograd worker -- so nothing here is thread_local); slots are handed out under a mutex.  A slot is reused after N
// later forwards on the same device: fnx_raster_check must be called before that (documented in include/fnx.h).
struct PinnedSlots {
    static constexpr int N = 256;
    long long *host = nullptr;
    cudaEvent_t ev[N];
    int next = 0;
    bool ok = false;
    int init() {
        if (ok) return FNX_OK;
        FNX_CUDA_TRY(cudaHostAlloc((void **)&host, sizeof(long long) * N, cudaHostAllocDefault));
        for (int i = 0; i < N; i++) FNX_CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        ok = true;
        return FNX_OK;
    }
};
constexpr int MAX_DEVICES = 32;
static PinnedSlots g_slots_of_device[MAX_DEVICES];
static std::mutex g_slots_mutex;
static int current_device_index() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
    return d;
}

// The longest-first start order pays in the BACKWARD (-10 % on the bench workload: its CTAs run 8 per SM and the tail of a launch
// is a few long tiles); in the forward it measured 3-4 % SLOWER than raster order (A/B in bench.py, round 2), so the forward keeps
// raster order unless FNX_LPT_FWD=1 is set in the environment (measurement switch).
static const uint32_t *fwd_tile_order(const fnx_raster_args *a) {
    static const bool on = [] { const char *e = getenv("FNX_LPT_FWD"); return e != nullptr && e[0] == '1'; }();
    return on ? a->tile_order : nullptr;
}

static int validate(const fnx_raster_args *a, bool backward = false) {
    FNX_REQUIRE(a != nullptr, "args is NULL");
    FNX_REQUIRE(a->C == 1 || a->C == 3, "C must be 1 or 3 (got %d)", a->C);
    FNX_REQUIRE(a->P >= 0 && a->V >= 1 && a->W > 0 && a->H > 0, "bad sizes P=%d V=%d W=%d H=%d", a->P, a->V, a->W, a->H);
    if (a->sh != nullptr) {
        FNX_REQUIRE(a->C == 3, "For non-RGB, provide precomputed Gaussian colors!");   // rasterizer_impl.cu:226-228
        FNX_REQUIRE(a->colors == nullptr, "provide exactly one of sh / colors");
        FNX_REQUIRE(a->sh_degree >= 0 && a->sh_degree <= 3 && a->sh_coeffs >= (a->sh_degree + 1) * (a->sh_degree + 1) && a->sh_coeffs <= 16,
                    "sh_degree must be 0..3 and sh_coeffs in [(degree+1)^2, 16]");
        FNX_REQUIRE(a->campos != nullptr, "campos [V,3] must be given with sh");
        if (a->flags & (FNX_BUCKET_BINNING | FNX_BIN_ONLY | FNX_NO_HOST_SYNC)) {
            set_error("SH colours are implemented for the plain forward / backward (no workspaces, no merged streams)");
            return FNX_ERR_UNSUPPORTED;
        }
    }
    if (a->P > 0) {
        // the backward reads colours and opacities from the forward's record stream: like the reference's backward
        // (rasterize_points.h:39-59 has no opacity argument) it does not need the pointer
        FNX_REQUIRE(a->means3D && (a->colors || a->sh) && (a->opacities || backward), "means3D / colors (or sh) / opacities must be given");
        FNX_REQUIRE((a->scales && a->rotations) || a->cov3D_precomp, "need scales+rotations or cov3D_precomp");
        FNX_REQUIRE(a->view_matrix && a->proj_matrix && a->bg, "view_matrix / proj_matrix / bg must be given");
    }
    FNX_REQUIRE((long long)a->P * a->V < (1ll << 31), "P*V too large");
    return FNX_OK;
}

constexpr int SORT_CAP = 2048;  // keys sorted in shared memory; larger buckets take the rank-sort path through bkeys2
constexpr int STATIC_DEPTH_SMEM = 2048;  // static depths staged per tile by merge_bucket_kernel

// Bitonic sort of up to 256 * E keys held E per thread by a 256-thread CTA; the key with index e * 256 + tid lives in v[e] of
// thread tid.  Partners at distance j < 32 are exchanged by warp shuffles, at 32 <= j < 256 through shared memory, at j >= 256 inside
// the thread's own registers: of the 36 (E = 1) .. 66 (E = 8) stages of the network only 6 .. 15 need the CTA barrier that EVERY stage
// of a network over a shared array needs.  n2 = padded size (power of two >= the key count; the padding keys are ~0): the stages
// beyond it would only order padding.  The keys end up ascending in s_key[0, 256 * E); ends with a barrier.
template <int E>
__device__ __forceinline__ void sort_bucket_regs(unsigned long long *s_key, const unsigned long long *__restrict__ bkeys, size_t base, int nf,
                                                 int n2) {
    constexpr int N = 256 * E;
    const int tid = threadIdx.x;
    unsigned long long v[E];
#pragma unroll
    for (int e = 0; e < E; e++) v[e] = (e * 256 + tid) < nf ? bkeys[base + e * 256 + tid] : ~0ull;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
        if (k > n2) continue;   // (uniform over the CTA; not a break: the loops must unroll so that v[] stays in registers)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 256) {
                const int je = j / 256;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if (e & je) continue;
                    const bool up = (((e * 256) | tid) & k) == 0;
                    const unsigned long long a = v[e], b = v[e | je];
                    const bool sw = (a > b) == up;
                    v[e] = sw ? b : a;
                    v[e | je] = sw ? a : b;
                }
            } else {
                if (j >= 32) {
#pragma unroll
                    for (int e = 0; e < E; e++) s_key[e * 256 + tid] = v[e];
                    __syncthreads();
                }
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const unsigned long long other = j >= 32 ? s_key[e * 256 + (tid ^ j)] : __shfl_xor_sync(0xffffffffu, v[e], j);
                    const bool up = (((e * 256) | tid) & k) == 0, lower = (tid & j) == 0;
                    v[e] = (lower == up) ? min(v[e], other) : max(v[e], other);
                }
                if (j >= 32) __syncthreads();
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; e++) s_key[e * 256 + tid] = v[e];
    __syncthreads();
}

// Sorts the nf keys bkeys[base, base+nf) of one tile (whole CTA participates; ends with a barrier).  Up to 2048 keys (and a CTA of
// 256 threads): register / shuffle bitonic network (sort_bucket_regs); up to `cap` keys: bitonic network in the shared array s_key
// (cap entries); more: rank sort (keys are unique) into bkeys2.  Returns where the sorted keys are.
__device__ __forceinline__ const unsigned long long *sort_bucket(unsigned long long *s_key, int cap, const unsigned long long *__restrict__ bkeys,
                                                                 unsigned long long *__restrict__ bkeys2, size_t base, int nf) {
    if (nf <= 2048 && cap >= 2048 && blockDim.x == 256) {
        int n2 = 2;
        while (n2 < nf) n2 <<= 1;
        if (nf <= 256) sort_bucket_regs<1>(s_key, bkeys, base, nf, n2);
        else if (nf <= 512) sort_bucket_regs<2>(s_key, bkeys, base, nf, n2);
        else if (nf <= 1024) sort_bucket_regs<4>(s_key, bkeys, base, nf, n2);
        else sort_bucket_regs<8>(s_key, bkeys, base, nf, n2);
        return s_key;
    }
    if (nf <= cap) {
        int n2 = 2;
        while (n2 < nf) n2 <<= 1;
        for (int i = threadIdx.x; i < n2; i += blockDim.x) s_key[i] = i < nf ? bkeys[base + i] : ~0ull;
        __syncthreads();
        for (int k = 2; k <= n2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int idx = threadIdx.x; idx < (n2 >> 1); idx += blockDim.x) {
                    const int i = ((idx & ~(j - 1)) << 1) | (idx & (j - 1));  // bit j clear
                    const int l = i | j;
                    const unsigned long long x = s_key[i], y = s_key[l];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { s_key[i] = y; s_key[l] = x; }
                }
                __syncthreads();
            }
        return s_key;
    }
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        const unsigned long long ki = bkeys[base + i];
        int r = 0;
        for (int q = 0; q < nf; q++) r += bkeys[base + q] < ki ? 1 : 0;
        bkeys2[base + r] = ki;
    }
    __syncthreads();
    return bkeys2 + base;
}

// General (single-stream) bucket binning: sort the tile's bucket and write its records in place -- the bucket offsets ARE
// the tile ranges.  Replaces the depth sort, the scan, emit's ordered output, the tile sort and pack_kernel of the sorted
// pipeline.  Dynamic shared memory: PACK_SORT_CAP keys.
constexpr int PACK_SORT_CAP = 8192;
template <int C>
__global__ void __launch_bounds__(256)
pack_bucket_kernel(int ntiles, int P, int gx, bool exact_rect, bool use_mask, int grad_begin, int grad_end,
                   const uint2 *__restrict__ ranges, const unsigned long long *__restrict__ bkeys, unsigned long long *__restrict__ bkeys2,
                   GeomView g, const float *__restrict__ colors, char *__restrict__ records) {
    extern __shared__ unsigned long long s_dyn_key[];
    const size_t t = (size_t)blockIdx.y * ntiles + blockIdx.x;
    const uint2 f = ranges[t];
    const int nf = (int)(f.y - f.x);
    if (nf == 0 || g.hdr->overflow) return;
    const unsigned long long *sk = sort_bucket(s_dyn_key, PACK_SORT_CAP, bkeys, bkeys2, (size_t)f.x, nf);
    const int tx = (int)blockIdx.x % gx, ty = (int)blockIdx.x / gx;
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        const unsigned long long key = sk[i];
        uint32_t slot = (uint32_t)key;
        const float depth = __uint_as_float((uint32_t)(key >> 32));
        const float2 xy = g.xy[slot];
        const float4 co = g.conic_o[slot];
        const uint32_t gi = slot % (uint32_t)P;
        if (use_mask) {
            slot |= patch_mask(xy, co, tx, ty, exact_rect) << SLOT_BITS;
            if ((int)gi < grad_begin || (int)gi >= grad_end) slot |= FROZEN_BIT;
        }
        float4 *rec = reinterpret_cast<float4 *>(records + ((size_t)f.x + i) * RecBytes<C>::value);
        rec[0] = make_float4(xy.x, xy.y, co.x, co.y);
        if (C == 3) {
            const float c0 = colors[3 * (size_t)gi], c1 = colors[3 * (size_t)gi + 1], c2 = colors[3 * (size_t)gi + 2];
            rec[1] = make_float4(co.z, co.w, c0, c1);
            rec[2] = make_float4(c2, __uint_as_float(slot), depth, 0.f);
        } else {
            rec[1] = make_float4(co.z, co.w, colors[gi], __uint_as_float(slot));
        }
    }
}


template <int C>
static int bin_and_blend(const fnx_raster_args *a, cudaStream_t st, GeomView &g, BinView &b, ImageView &im,
                         long long cap, long long sort_items, const int *radii, float *out_color, float *out_depth) {
    const int P = a->P, V = a->V, n = P * V;
    const int gx = (a->W + TILE - 1) / TILE, gy = (a->H + TILE - 1) / TILE, ntiles = gx * gy;
    const bool exact_rect = (a->flags & FNX_EXACT_RECT) != 0;
    const bool bucket = (a->flags & FNX_BUCKET_BINNING) != 0;
    if (bucket) {  // the tile ranges are the bucket offsets written by tile_scan_kernel; no global sort
        const bool use_mask = (long long)P * V < (1ll << SLOT_BITS);
        const bool all_frozen = (a->flags & FNX_ALL_FROZEN) != 0;
        const bool all_grad = !all_frozen && a->grad_end <= a->grad_begin;
        // (per device / context, so set on every call: it is a host-side table update, not a launch)
        FNX_CUDA_TRY(cudaFuncSetAttribute(pack_bucket_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, PACK_SORT_CAP * 8));
        FNX_CUDA_TRY(cudaMemsetAsync(im.tile_cursor, 0, sizeof(uint32_t) * (size_t)ntiles * V, st));
        prof_begin(SEC_EMIT, st);
        emit_bucket_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, P, gx, gy, exact_rect, radii, g, im.ranges, im.tile_cursor, b.bkeys);
        prof_end(SEC_EMIT, st);
        FNX_LAUNCH_CHECK("emit_bucket_kernel");
        prof_begin(SEC_PACK, st);
        pack_bucket_kernel<C><<<dim3(ntiles, V), 256, PACK_SORT_CAP * 8, st>>>(ntiles, P, gx, exact_rect, use_mask,
                                                                            all_grad ? 0 : (all_frozen ? 0 : a->grad_begin),
                                                                            all_grad ? P : (all_frozen ? 0 : a->grad_end), im.ranges, b.bkeys,
                                                                            b.bkeys2, g, a->colors, b.records);
        prof_end(SEC_PACK, st);
        FNX_LAUNCH_CHECK("pack_bucket_kernel");
    } else {
        FNX_CUDA_TRY(cudaMemsetAsync(im.ranges, 0, sizeof(uint2) * (size_t)ntiles * V, st));
    }
    if (!bucket && sort_items > 0) {
        const bool padded = (a->flags & FNX_NO_HOST_SYNC) != 0 || a->instance_capacity_hint > 0;
        if (padded) FNX_CUDA_TRY(cudaMemsetAsync(b.tkeys_in, 0xFF, sizeof(uint32_t) * (size_t)sort_items, st));
        prof_begin(SEC_EMIT, st);
        emit_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, P, gx, gy, exact_rect, radii, g, b);
        prof_end(SEC_EMIT, st);
        FNX_LAUNCH_CHECK("emit_kernel");
        const int end_bit = ceil_log2_u64((uint64_t)ntiles * V + 1);
        size_t tb = b.cub_temp_bytes;
        prof_begin(SEC_TILE_SORT, st);
        FNX_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b.cub_temp, tb, b.tkeys_in, b.tkeys_out, b.tvals_in, b.tvals_out,
                                                     (int)sort_items, 0, end_bit, st));
        prof_end(SEC_TILE_SORT, st);
        prof_begin(SEC_PACK, st);
        const bool use_mask = (long long)P * V < (1ll << SLOT_BITS);
        const bool all_frozen = (a->flags & FNX_ALL_FROZEN) != 0;
        const bool all_grad = !all_frozen && a->grad_end <= a->grad_begin;
        pack_kernel<C><<<(unsigned)((sort_items + 255) / 256), 256, 0, st>>>(cap, P, gx, ntiles, exact_rect, use_mask,
                                                                            all_grad ? 0 : (all_frozen ? 0 : a->grad_begin),
                                                                            all_grad ? P : (all_frozen ? 0 : a->grad_end), a->sh ? sh_view(geom_end(g), P, V).color : a->colors,
                                                                            a->sh != nullptr, g, b, im.ranges);
        prof_end(SEC_PACK, st);
        FNX_LAUNCH_CHECK("pack_kernel");
    }
    if (a->flags & FNX_BIN_ONLY) return FNX_OK;  // the caller blends a merged stream (fnx_raster_blend_merged)
    dim3 grid(ntiles * CTAS_PER_TILE, V);
    prof_begin(SEC_BLEND_FWD, st);
    blend_fwd_kernel<C><<<grid, BLEND_THREADS, 0, st>>>(a->W, a->H, gx, gy, (long long)P * V < (1ll << SLOT_BITS), fwd_tile_order(a), b.records, nullptr,
                                                        im.ranges, nullptr, nullptr, nullptr, nullptr, g.depth, a->bg, g.hdr, im, out_color,
                                                        out_depth);
    prof_end(SEC_BLEND_FWD, st);
    FNX_LAUNCH_CHECK("blend_fwd_kernel");
    return FNX_OK;
}

template <int C>
static int forward_impl(const fnx_raster_args *a, fnx_alloc_fn ag, void *cg, fnx_alloc_fn ab, void *cb, fnx_alloc_fn ai,
                        void *ci, float *out_color, float *out_depth, int32_t *radii, int64_t *num_rendered_host,
                        fnx_raster_scratch *scratch, cudaStream_t st) {
    FNX_REQUIRE(ag && ab && ai && scratch && num_rendered_host, "allocators / scratch / num_rendered_host must be given");
    FNX_REQUIRE((out_color && out_depth) || (a->flags & FNX_BIN_ONLY), "out_color / out_depth must be given");
    const int P = a->P, V = a->V, W = a->W, H = a->H;
    const size_t HW = (size_t)W * H;
    memset(scratch, 0, sizeof(*scratch));
    *num_rendered_host = 0;
    if (P == 0) {  // rasterize_points.cu:81: zero outputs
        FNX_REQUIRE(!(a->flags & FNX_BIN_ONLY), "FNX_BIN_ONLY needs P > 0");
        FNX_CUDA_TRY(cudaMemsetAsync(out_color, 0, sizeof(float) * HW * C * V, st));
        FNX_CUDA_TRY(cudaMemsetAsync(out_depth, 0, sizeof(float) * HW * V, st));
        return FNX_OK;
    }
    FNX_REQUIRE(radii != nullptr, "radii must be given");
    const int n = P * V;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float focal_y = H / (2.0f * a->tan_fov_y), focal_x = W / (2.0f * a->tan_fov_x);
    const bool exact_rect = (a->flags & FNX_EXACT_RECT) != 0;
    const bool no_sync = (a->flags & FNX_NO_HOST_SYNC) != 0;
    FNX_REQUIRE(!no_sync || a->instance_capacity_hint > 0, "FNX_NO_HOST_SYNC needs instance_capacity_hint > 0");
    int rc = FNX_OK;
    if (!no_sync) {
        std::lock_guard<std::mutex> lock(g_slots_mutex);
        rc = g_slots_of_device[current_device_index()].init();
        if (rc) return rc;
    }

    scratch->geom_bytes = geom_bytes(P, V) + (a->sh ? sh_extra_bytes(P, V) : 0);
    scratch->geom = ag(cg, scratch->geom_bytes);
    scratch->image_bytes = image_bytes(W, H, V);
    scratch->image = ai(ci, scratch->image_bytes);
    if (!scratch->geom || !scratch->image) {
        set_error("allocation callback returned NULL");
        return FNX_ERR_ALLOC;
    }
    GeomView g = geom_view(scratch->geom, P, V);
    ImageView im = image_view(scratch->image, W, H, V);

    dim3 pgrid((P + 255) / 256, V);
    const bool bucket = (a->flags & FNX_BUCKET_BINNING) != 0;
    const bool bucket_dyn = bucket && (a->flags & FNX_BIN_ONLY) != 0;  // dynamic set of merged streams: sorted inside the merge
    if (bucket) {
        FNX_REQUIRE(!bucket_dyn || (C == 3 && no_sync), "FNX_BUCKET_BINNING | FNX_BIN_ONLY needs C == 3 and FNX_NO_HOST_SYNC");
        const int nt = gx * gy * V;
        FNX_CUDA_TRY(cudaMemsetAsync(im.tile_count, 0, sizeof(uint32_t) * 2 * (size_t)nt, st));  // histogram + cursors
    }
    prof_begin(SEC_PREPROCESS, st);
    preprocess_kernel<<<pgrid, 256, 0, st>>>(P, V, a->means3D, (const float3 *)a->scales, a->scale_modifier,
                                             (const float4 *)a->rotations, a->opacities, a->cov3D_precomp, a->view_matrix,
                                             a->proj_matrix, W, H, a->tan_fov_x, a->tan_fov_y, focal_x, focal_y, gx, gy,
                                             exact_rect, radii, g, bucket ? im.tile_count : nullptr);
    if (a->sh != nullptr) {
        sh_color_kernel<<<pgrid, 256, 0, st>>>(P, V, a->sh_degree, a->sh_coeffs, a->means3D, a->campos, a->sh, radii, sh_view(geom_end(g), P, V));
    }
    prof_end(SEC_PREPROCESS, st);
    FNX_LAUNCH_CHECK("preprocess_kernel");
    if (bucket_dyn) {  // histogram -> bucket offsets -> unsorted per-tile buckets; fnx_raster_blend_merged sorts and merges them
        const long long bcap = a->instance_capacity_hint;
        const int nt = gx * gy * V;
        prof_begin(SEC_EMIT, st);
        tile_scan_kernel<<<1, 1024, 0, st>>>(nt, im.tile_count, im.ranges, g.hdr, bcap, (long long *)a->num_rendered_pinned);
        FNX_LAUNCH_CHECK("tile_scan_kernel");
        scratch->binning_bytes = binning_bytes(bcap, C);
        scratch->binning = ab(cb, scratch->binning_bytes);
        if (!scratch->binning) {
            set_error("allocation callback returned NULL");
            return FNX_ERR_ALLOC;
        }
        scratch->binning_capacity = bcap;
        scratch->check_slot = -1;
        BinView b = bin_view(scratch->binning, bcap, C);
        emit_bucket_kernel<<<(P * V + 255) / 256, 256, 0, st>>>(P * V, P, gx, gy, exact_rect, radii, g, im.ranges, im.tile_cursor, b.bkeys);
        prof_end(SEC_EMIT, st);
        FNX_LAUNCH_CHECK("emit_bucket_kernel");
        *num_rendered_host = -1;
        return FNX_OK;
    }
    if (!bucket) {
        prof_begin(SEC_DEPTH_SORT, st);
        size_t tb = g.cub_temp_bytes;
        const int dbits = 32 + ceil_log2_u64((uint64_t)V);
        FNX_CUDA_TRY(cub::DeviceRadixSort::SortPairs(g.cub_temp, tb, g.dkeys_in, g.dkeys_out, g.dvals_in, g.dvals_out, n, 0, dbits, st));
        DepthScanIter it(cub::CountingInputIterator<uint32_t>(0), DepthScanIn{g.tiles_touched, g.dvals_out});
        tb = g.cub_temp_bytes;
        FNX_CUDA_TRY(cub::DeviceScan::ExclusiveSum(g.cub_temp, tb, it, g.offsets, n, st));
        prof_end(SEC_DEPTH_SORT, st);
    }
    // instance count + capacity / overflow header: from the depth-ordered scan (sorted pipeline) or the tile histogram (buckets)
    const int nt_all = gx * gy * V;
    auto write_header = [&](long long capacity, long long *pinned_dst) -> int {
        if (bucket) {
            tile_scan_kernel<<<1, 1024, 0, st>>>(nt_all, im.tile_count, im.ranges, g.hdr, capacity, pinned_dst);
            FNX_LAUNCH_CHECK("tile_scan_kernel");
        } else {
            finish_scan_kernel<<<1, 1, 0, st>>>(n, g, capacity, pinned_dst);
            FNX_LAUNCH_CHECK("finish_scan_kernel");
        }
        return FNX_OK;
    };

    long long cap = a->instance_capacity_hint > 0 ? a->instance_capacity_hint : -1;
    scratch->binning = nullptr;
    scratch->check_slot = -1;

    if (no_sync) {  // no events, no waits: capturable into a CUDA graph
        rc = write_header(cap, (long long *)a->num_rendered_pinned);
        if (rc) return rc;
        scratch->binning_bytes = binning_bytes(cap, C);
        scratch->binning = ab(cb, scratch->binning_bytes);
        if (!scratch->binning) {
            set_error("allocation callback returned NULL");
            return FNX_ERR_ALLOC;
        }
        scratch->binning_capacity = cap;
        BinView b = bin_view(scratch->binning, cap, C);
        rc = bin_and_blend<C>(a, st, g, b, im, cap, cap, radii, out_color, out_depth);
        *num_rendered_host = -1;
        return rc;
    }

    const int dev_id = current_device_index();
    PinnedSlots &g_slots = g_slots_of_device[dev_id];
    int slot_id;
    {
        std::lock_guard<std::mutex> lock(g_slots_mutex);
        slot_id = g_slots.next;
        g_slots.next = (g_slots.next + 1) % PinnedSlots::N;
    }
    long long *pinned = g_slots.host + slot_id;
    *pinned = -1;
    rc = write_header(cap, pinned);
    if (rc) return rc;
    FNX_CUDA_TRY(cudaEventRecord(g_slots.ev[slot_id], st));

    long long R = -1;
    const bool exact = cap < 0;
    if (exact) {  // exact sizing: one host sync, like the reference (rasterizer_impl.cu:263-264)
        FNX_CUDA_TRY(cudaEventSynchronize(g_slots.ev[slot_id]));
        R = *pinned;
        cap = R;
    }
    for (int attempt = 0; attempt < 3; attempt++) {
        scratch->binning_bytes = binning_bytes(cap, C);
        scratch->binning = ab(cb, scratch->binning_bytes);
        if (!scratch->binning) {
            set_error("allocation callback returned NULL");
            return FNX_ERR_ALLOC;
        }
        scratch->binning_capacity = cap;
        BinView b = bin_view(scratch->binning, cap, C);
        if (attempt >= 1) {  // re-arm capacity / overflow flag for the retry
            rc = write_header(cap, nullptr);
            if (rc) return rc;
        }
        rc = bin_and_blend<C>(a, st, g, b, im, cap, exact ? R : cap, radii, out_color, out_depth);
        if (rc) return rc;
        if (exact) break;
        // everything is queued; the count has almost surely landed already -- this wait does not stall the GPU
        FNX_CUDA_TRY(cudaEventSynchronize(g_slots.ev[slot_id]));
        R = *pinned;
        if (R <= cap) break;
        cap = R + R / 8 + 1024;  // overflow: grow and redo binning + blend
    }
    *num_rendered_host = R;
    if (a->num_rendered_pinned) *a->num_rendered_pinned = R;
    scratch->check_slot = slot_id;
    scratch->reserved = dev_id;   // the device whose ring holds the slot
    return FNX_OK;
}

template <int C>
static int backward_impl(const fnx_raster_args *a, const fnx_raster_scratch *scratch, int64_t num_rendered,
                         const int32_t *radii, const float *dL_dout_color, const fnx_raster_grads *gr, cudaStream_t st) {
    FNX_REQUIRE(scratch && gr, "scratch / grads must be given");
    const int P = a->P, V = a->V, W = a->W, H = a->H;
    if (P == 0) return FNX_OK;
    FNX_REQUIRE(scratch->geom && scratch->image && scratch->binning, "scratch buffers missing (forward not run?)");
    FNX_REQUIRE(radii && dL_dout_color, "radii / dL_dout_color must be given");
    (void)num_rendered;
    constexpr int ACC = AccFloats<C>::value;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE, ntiles = gx * gy;
    const float focal_y = H / (2.0f * a->tan_fov_y), focal_x = W / (2.0f * a->tan_fov_x);
    GeomView g = geom_view(scratch->geom, P, V);
    ImageView im = image_view(scratch->image, W, H, V);
    // capacity is only needed to locate the record stream, which sits at a capacity-dependent offset
    long long cap = 0;
    {
        // binning_bytes is monotone in cap; recover cap from the header written by the forward (host copy kept in scratch)
        cap = scratch->binning_capacity;
    }
    BinView b = bin_view(scratch->binning, cap, C);
    FNX_CUDA_TRY(cudaMemsetAsync(g.accum, 0, sizeof(float) * (size_t)P * V * ACC, st));
    dim3 grid(ntiles * CTAS_PER_TILE, V);
    prof_begin(SEC_BLEND_BWD, st);
    blend_bwd_kernel<C><<<grid, BLEND_THREADS, 0, st>>>(W, H, gx, gy, (long long)P * V < (1ll << SLOT_BITS), a->tile_order, b.records, nullptr, im.ranges,
                                                        nullptr, nullptr, nullptr, nullptr, a->bg, g.hdr, im, dL_dout_color, g.accum);
    prof_end(SEC_BLEND_BWD, st);
    FNX_LAUNCH_CHECK("blend_bwd_kernel");
    prof_begin(SEC_GEOM_BWD, st);
    geom_bwd_kernel<C><<<(P + 255) / 256, 256, 0, st>>>(P, V, a->means3D, (const float3 *)a->scales, a->scale_modifier,
                                                        (const float4 *)a->rotations, a->cov3D_precomp, a->view_matrix,
                                                        a->proj_matrix, W, H, a->tan_fov_x, a->tan_fov_y, focal_x, focal_y,
                                                        radii, g.cov3D, g.accum, (a->grad_end <= a->grad_begin) ? 0 : a->grad_begin,
                                                        (a->grad_end <= a->grad_begin) ? P : a->grad_end, *gr);
    if (a->sh != nullptr) {
        if constexpr (C == 3) {
            sh_bwd_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, V, a->sh_degree, a->sh_coeffs, a->means3D, a->campos, a->sh, radii,
                                                           sh_view(geom_end(g), P, V), g.accum, gr->dL_dsh, gr->dL_dmeans3D);
        }
    }
    prof_end(SEC_GEOM_BWD, st);
    FNX_LAUNCH_CHECK("geom_bwd_kernel");
    return FNX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Static + dynamic record streams.  The frozen background set and the cameras do not change within a frame, so
// its depth-sorted, tile-partitioned record stream is built ONCE (FNX_BIN_ONLY | FNX_ALL_FROZEN) and kept in HBM; each
// iteration bins only the moving Gaussians and merges the two streams per tile by depth (dynamic first on ties:
// the dynamic Gaussians precede the static ones in the reference's concatenated array, pipe_dynamics.py:51-57, and
// its sort is stable in the index).  Tiles without dynamic instances are not copied at all: the blend kernels read
// their span straight from the static stream (tile_src = 1).
// ---------------------------------------------------------------------------------------------------------------
constexpr int MERGE_SMEM = 2048;  // dynamic depths staged per tile

__device__ __forceinline__ float rec48_depth(const char *recs, size_t i) { return reinterpret_cast<const float *>(recs + i * 48)[10]; }

__global__ void __launch_bounds__(256)
merge_kernel(int ntiles, const uint2 *__restrict__ ranges_dyn, const uint2 *__restrict__ ranges_stat, const char *__restrict__ rec_dyn,
             const char *__restrict__ rec_stat, char *__restrict__ rec_merged, GeomHeader *__restrict__ hdr_dyn,
             const GeomHeader *__restrict__ hdr_stat, const uint32_t *__restrict__ static_last, const int32_t *__restrict__ view_map,
             uint2 *__restrict__ mranges, uint32_t *__restrict__ tile_src, uint32_t *__restrict__ tile_dyn_last,
             uint32_t *__restrict__ tile_dyn_first) {
    __shared__ float s_depth[MERGE_SMEM];
    __shared__ unsigned long long s_base;
    const size_t t = (size_t)blockIdx.y * ntiles + blockIdx.x;
    // the static stream may hold more cameras than this call renders: view_map[v] = this call's view v in the static stream
    const size_t ts = (size_t)(view_map ? view_map[blockIdx.y] : (int)blockIdx.y) * ntiles + blockIdx.x;
    const uint2 f = ranges_dyn[t], b = ranges_stat[ts];
    int nf = (int)(f.y - f.x);
    int nb = (int)(b.y - b.x);
    // Static records behind the last one that the static-only blend of this tile used can never be reached once more
    // occluders are inserted: a pixel's transmittance at a given static record only shrinks (rounding is monotone),
    // so it terminates no later, and the alpha test does not depend on what lies in front.
    if (hdr_stat->static_prepared) nb = min(nb, (int)max(max(static_last[4 * ts], static_last[4 * ts + 1]), max(static_last[4 * ts + 2], static_last[4 * ts + 3])));
    if (hdr_dyn->overflow) nf = 0;
    if (nf == 0) {
        if (threadIdx.x == 0) { mranges[t] = make_uint2(b.x, b.x + nb); tile_src[t] = 1u; tile_dyn_last[t] = 0u; tile_dyn_first[t] = 0u; }
        return;
    }
    // the merged span of this tile is bump-allocated (tile order inside the merged stream does not matter; the ranges
    // of tiles without instances are (0,0), so they cannot serve as prefix sums)
    if (threadIdx.x == 0) {
        s_base = atomicAdd(&hdr_dyn->merge_cursor, (unsigned long long)(nf + nb));
        mranges[t] = make_uint2((uint32_t)s_base, (uint32_t)(s_base + nf + nb));
        tile_src[t] = 0u;
    }
    const bool staged = nf <= MERGE_SMEM;
    if (staged)
        for (int i = threadIdx.x; i < nf; i += blockDim.x) s_depth[i] = rec48_depth(rec_dyn, (size_t)f.x + i);
    __syncthreads();
    const size_t ms = (size_t)s_base;
    // static records: shifted by the number of dynamic records in front of them (depth <= theirs)
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const float4 *src = reinterpret_cast<const float4 *>(rec_stat + ((size_t)b.x + j) * 48);
        const float4 r0 = src[0], r1 = src[1], r2 = src[2];
        const float d = r2.z;
        int lo = 0, hi = nf;  // upper bound: first dynamic depth > d
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const float dm = staged ? s_depth[mid] : rec48_depth(rec_dyn, (size_t)f.x + mid);
            if (dm <= d) lo = mid + 1; else hi = mid;
        }
        float4 *dst = reinterpret_cast<float4 *>(rec_merged + (ms + j + lo) * 48);
        dst[0] = r0; dst[1] = r1; dst[2] = r2;
    }
    // dynamic records: shifted by the number of static records strictly in front of them
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        const float4 *src = reinterpret_cast<const float4 *>(rec_dyn + ((size_t)f.x + i) * 48);
        const float4 r0 = src[0], r1 = src[1], r2 = src[2];
        const float d = r2.z;
        int lo = 0, hi = nb;  // lower bound: first static depth >= d
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rec48_depth(rec_stat, (size_t)b.x + mid) < d) lo = mid + 1; else hi = mid;
        }
        float4 *dst = reinterpret_cast<float4 *>(rec_merged + (ms + i + lo) * 48);
        dst[0] = r0; dst[1] = r1; dst[2] = r2;
        if (i == nf - 1) tile_dyn_last[t] = (uint32_t)(i + lo + 1);  // dynamic depths ascend: this is the deepest one
        if (i == 0) tile_dyn_first[t] = (uint32_t)lo;                // ... and this the nearest: `lo` static records lie in front of it
    }
}

// Bucket-binned dynamic set (emit_bucket_kernel): sort the tile's (depth, slot) keys in shared memory, build the dynamic
// records from the per-Gaussian state (what pack_kernel does for a sorted stream) and merge them with the static span.
__global__ void __launch_bounds__(256)
merge_bucket_kernel(int ntiles, int P, int gx, bool exact_rect, const uint2 *__restrict__ ranges_dyn, const uint2 *__restrict__ ranges_stat,
                    const unsigned long long *__restrict__ bkeys, unsigned long long *__restrict__ bkeys2, GeomView g,
                    const float *__restrict__ colors, const char *__restrict__ rec_stat, char *__restrict__ rec_merged,
                    const GeomHeader *__restrict__ hdr_stat, const uint32_t *__restrict__ static_last, const int32_t *__restrict__ view_map,
                    uint2 *__restrict__ mranges, uint32_t *__restrict__ tile_src, uint32_t *__restrict__ tile_dyn_last,
                    uint32_t *__restrict__ tile_dyn_first) {
    __shared__ unsigned long long s_key[SORT_CAP];
    __shared__ unsigned long long s_base;
    const size_t t = (size_t)blockIdx.y * ntiles + blockIdx.x;
    const size_t ts = (size_t)(view_map ? view_map[blockIdx.y] : (int)blockIdx.y) * ntiles + blockIdx.x;   // see merge_kernel
    const uint2 f = ranges_dyn[t], b = ranges_stat[ts];
    int nf = (int)(f.y - f.x);
    int nb = (int)(b.y - b.x);
    if (hdr_stat->static_prepared)  // see merge_kernel
        nb = min(nb, (int)max(max(static_last[4 * ts], static_last[4 * ts + 1]), max(static_last[4 * ts + 2], static_last[4 * ts + 3])));
    if (g.hdr->overflow) nf = 0;
    if (nf == 0) {
        if (threadIdx.x == 0) { mranges[t] = make_uint2(b.x, b.x + nb); tile_src[t] = 1u; tile_dyn_last[t] = 0u; tile_dyn_first[t] = 0u; }
        return;
    }
    if (threadIdx.x == 0) {
        s_base = atomicAdd(&g.hdr->merge_cursor, (unsigned long long)(nf + nb));
        mranges[t] = make_uint2((uint32_t)s_base, (uint32_t)(s_base + nf + nb));
        tile_src[t] = 0u;
    }
    // depths of the (truncated) static span, staged once: the dynamic records binary-search them
    __shared__ uint32_t s_sdepth[STATIC_DEPTH_SMEM];
    const bool sdepth_staged = nb <= STATIC_DEPTH_SMEM;
    if (sdepth_staged)
        for (int j = threadIdx.x; j < nb; j += blockDim.x) s_sdepth[j] = __float_as_uint(rec48_depth(rec_stat, (size_t)b.x + j));
    const unsigned long long *sk = sort_bucket(s_key, SORT_CAP, bkeys, bkeys2, (size_t)f.x, nf);   // (ends with a barrier)
    const size_t ms = (size_t)s_base;
    // static records: shifted by the number of dynamic records in front of them (depth <= theirs; depths are positive
    // floats, so their bit patterns order like the values)
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const float4 *src = reinterpret_cast<const float4 *>(rec_stat + ((size_t)b.x + j) * 48);
        const float4 r0 = src[0], r1 = src[1], r2 = src[2];
        const uint32_t d = __float_as_uint(r2.z);
        int lo = 0, hi = nf;  // upper bound: first dynamic depth > d
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((uint32_t)(sk[mid] >> 32) <= d) lo = mid + 1; else hi = mid;
        }
        float4 *dst = reinterpret_cast<float4 *>(rec_merged + (ms + j + lo) * 48);
        dst[0] = r0; dst[1] = r1; dst[2] = r2;
    }
    // dynamic records: built here, shifted by the number of static records strictly in front of them
    const uint32_t tl = (uint32_t)blockIdx.x;
    const int tx = tl % gx, ty = tl / gx;
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        const unsigned long long key = sk[i];
        const uint32_t d = (uint32_t)(key >> 32);
        uint32_t slot = (uint32_t)key;
        int lo = 0, hi = nb;  // lower bound: first static depth >= d
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const uint32_t sd = sdepth_staged ? s_sdepth[mid] : __float_as_uint(rec48_depth(rec_stat, (size_t)b.x + mid));
            if (sd < d) lo = mid + 1; else hi = mid;
        }
        const float2 xy = g.xy[slot];
        const float4 co = g.conic_o[slot];
        const uint32_t gi = slot % (uint32_t)P;
        const float c0 = colors[3 * (size_t)gi], c1 = colors[3 * (size_t)gi + 1], c2 = colors[3 * (size_t)gi + 2];
        slot |= patch_mask(xy, co, tx, ty, exact_rect) << SLOT_BITS;
        float4 *dst = reinterpret_cast<float4 *>(rec_merged + (ms + i + lo) * 48);
        dst[0] = make_float4(xy.x, xy.y, co.x, co.y);
        dst[1] = make_float4(co.z, co.w, c0, c1);
        dst[2] = make_float4(c2, __uint_as_float(slot), __uint_as_float(d), 0.f);
        if (i == nf - 1) tile_dyn_last[t] = (uint32_t)(i + lo + 1);
        if (i == 0) tile_dyn_first[t] = (uint32_t)lo;
    }
}

// marks a static stream as blended (its image scratch now holds tile_last of the static-only blend; tile_cached set by the blend)
__global__ void static_prepared_kernel(GeomHeader *hdr) { hdr->static_prepared = 1; }

static int blend_merged(const fnx_raster_args *a, const fnx_raster_scratch *dyn, const fnx_raster_scratch *stat, int P_static,
                        void *merged_records, float *out_color, float *out_depth, cudaStream_t st) {
    FNX_REQUIRE(a && dyn && stat && merged_records && out_color && out_depth, "bad arguments");
    FNX_REQUIRE(a->C == 3, "merged static+dynamic streams are implemented for 3-channel records (they carry the depth)");
    FNX_REQUIRE(dyn->geom && dyn->binning && dyn->image && stat->geom && stat->binning && stat->image, "scratch missing");
    const int P = a->P, V = a->V, W = a->W, H = a->H;
    FNX_REQUIRE((long long)P * V < (1ll << SLOT_BITS) && (long long)P_static * V < (1ll << SLOT_BITS), "too many Gaussians for masked records");
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE, ntiles = gx * gy;
    const int Vs = a->static_views > 0 ? a->static_views : V;   // cameras the static stream was built for
    FNX_REQUIRE(a->static_view_map != nullptr || Vs == V, "static_views differs from V: static_view_map must be given");
    FNX_REQUIRE((long long)P_static * Vs < (1ll << SLOT_BITS), "too many Gaussians for masked records");
    GeomView g = geom_view(dyn->geom, P, V);
    GeomView gs = geom_view(stat->geom, P_static, Vs);
    ImageView im = image_view(dyn->image, W, H, V);
    ImageView ims = image_view(stat->image, W, H, Vs);
    BinView b = bin_view(dyn->binning, dyn->binning_capacity, 3);
    BinView bs = bin_view(stat->binning, stat->binning_capacity, 3);
    const bool tile_cache = (a->flags & FNX_STATIC_TILE_CACHE) != 0;
    dim3 grid(ntiles, V);
    prof_begin(SEC_PACK, st);
    if (a->flags & FNX_BUCKET_BINNING)
        merge_bucket_kernel<<<grid, 256, 0, st>>>(ntiles, P, gx, (a->flags & FNX_EXACT_RECT) != 0, im.ranges, ims.ranges, b.bkeys, b.bkeys2, g,
                                                  a->colors, bs.records, (char *)merged_records, gs.hdr, ims.tile_last, a->static_view_map,
                                                  im.mranges, im.tile_src, im.tile_dyn_last, im.tile_dyn_first);
    else
        merge_kernel<<<grid, 256, 0, st>>>(ntiles, im.ranges, ims.ranges, b.records, bs.records, (char *)merged_records, g.hdr, gs.hdr,
                                           ims.tile_last, a->static_view_map, im.mranges, im.tile_src, im.tile_dyn_last, im.tile_dyn_first);
    prof_end(SEC_PACK, st);
    FNX_LAUNCH_CHECK("merge_kernel");
    prof_begin(SEC_BLEND_FWD, st);
    blend_fwd_kernel<3><<<dim3(ntiles * CTAS_PER_TILE, V), BLEND_THREADS, 0, st>>>(W, H, gx, gy, true, fwd_tile_order(a), (const char *)merged_records, bs.records, im.mranges, im.tile_src,
                                                        tile_cache ? im.tile_cached : nullptr, im.tile_dyn_last, im.snap, g.depth, a->bg,
                                                        g.hdr, im, out_color, out_depth);
    prof_end(SEC_BLEND_FWD, st);
    FNX_LAUNCH_CHECK("blend_fwd_kernel");
    return FNX_OK;
}

// Blend the static stream alone, once: leaves the static-only image in out_color / out_depth (the buffers every later
// fnx_raster_blend_merged(FNX_STATIC_TILE_CACHE) call must be given), the per-tile depth the static-only blend reaches
// (bounds the static records a merge has to copy) and sets every tile's cached flag.
static int static_prepare(const fnx_raster_args *a, const fnx_raster_scratch *stat, float *out_color, float *out_depth, cudaStream_t st) {
    FNX_REQUIRE(a && stat && out_color && out_depth, "bad arguments");
    FNX_REQUIRE(a->C == 3, "static streams are implemented for 3-channel records");
    FNX_REQUIRE(stat->geom && stat->binning && stat->image, "scratch missing (run fnx_raster_forward(FNX_BIN_ONLY | FNX_ALL_FROZEN) first)");
    const int P = a->P, V = a->V, W = a->W, H = a->H;
    FNX_REQUIRE((long long)P * V < (1ll << SLOT_BITS), "too many Gaussians for masked records");
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE, ntiles = gx * gy;
    GeomView gs = geom_view(stat->geom, P, V);
    ImageView ims = image_view(stat->image, W, H, V);
    BinView bs = bin_view(stat->binning, stat->binning_capacity, 3);
    FNX_CUDA_TRY(cudaMemsetAsync(ims.tile_cached, 0, sizeof(uint32_t) * (size_t)ntiles * V * TILE_PATCHES, st));
    FNX_CUDA_TRY(cudaMemsetAsync(ims.tile_src, 0xFF, sizeof(uint32_t) * (size_t)ntiles * V, st));  // every tile: "from the static stream"
    dim3 grid(ntiles * CTAS_PER_TILE, V);
    blend_fwd_kernel<3><<<grid, BLEND_THREADS, 0, st>>>(W, H, gx, gy, true, nullptr, bs.records, bs.records, ims.ranges, ims.tile_src, ims.tile_cached,
                                                        nullptr, nullptr, gs.depth, a->bg, gs.hdr, ims, out_color, out_depth);
    FNX_LAUNCH_CHECK("blend_fwd_kernel");
    static_prepared_kernel<<<1, 1, 0, st>>>(gs.hdr);
    FNX_LAUNCH_CHECK("static_prepared_kernel");
    return FNX_OK;
}

static int backward_merged(const fnx_raster_args *a, const fnx_raster_scratch *dyn, const fnx_raster_scratch *stat, const void *merged_records,
                           const int32_t *radii, const float *dL_dout_color, const fnx_raster_grads *gr, cudaStream_t st) {
    FNX_REQUIRE(a && dyn && stat && merged_records && radii && dL_dout_color && gr, "bad arguments");
    FNX_REQUIRE(a->C == 3, "merged streams need C == 3");
    const int P = a->P, V = a->V, W = a->W, H = a->H;
    constexpr int ACC = AccFloats<3>::value;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE, ntiles = gx * gy;
    const float focal_y = H / (2.0f * a->tan_fov_y), focal_x = W / (2.0f * a->tan_fov_x);
    GeomView g = geom_view(dyn->geom, P, V);
    ImageView im = image_view(dyn->image, W, H, V);
    BinView bs = bin_view(stat->binning, stat->binning_capacity, 3);
    FNX_CUDA_TRY(cudaMemsetAsync(g.accum, 0, sizeof(float) * (size_t)P * V * ACC, st));
    dim3 grid(ntiles * CTAS_PER_TILE, V);
    prof_begin(SEC_BLEND_BWD, st);
    blend_bwd_kernel<3><<<grid, BLEND_THREADS, 0, st>>>(W, H, gx, gy, true, a->tile_order, (const char *)merged_records, bs.records, im.mranges, im.tile_src,
                                                        im.tile_dyn_last, im.tile_dyn_first, im.snap, a->bg, g.hdr, im, dL_dout_color, g.accum);
    prof_end(SEC_BLEND_BWD, st);
    FNX_LAUNCH_CHECK("blend_bwd_kernel");
    prof_begin(SEC_GEOM_BWD, st);
    geom_bwd_kernel<3><<<(P + 255) / 256, 256, 0, st>>>(P, V, a->means3D, (const float3 *)a->scales, a->scale_modifier,
                                                        (const float4 *)a->rotations, a->cov3D_precomp, a->view_matrix, a->proj_matrix, W, H,
                                                        a->tan_fov_x, a->tan_fov_y, focal_x, focal_y, radii, g.cov3D, g.accum, 0, P, *gr);
    prof_end(SEC_GEOM_BWD, st);
    FNX_LAUNCH_CHECK("geom_bwd_kernel");
    return FNX_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Longest-processing-time-first order of the (view, tile) units for the blend kernels (work_unit): a counting sort, by
// decreasing work of the LAST forward (records the slowest 8x8 patch of the tile walked; 0 for tiles that were served from the
// static tile cache), one CTA.  Cameras and the frozen set are fixed within a frame and the fluid moves little per iteration,
// so an order taken once per frame (or refreshed now and then) stays a good schedule; a stale order costs time, never
// correctness (it is a permutation).
// ---------------------------------------------------------------------------------------------------------------
constexpr int ORDER_BINS = 2048;
__global__ void __launch_bounds__(1024)
tile_order_kernel(int nunits, const uint32_t *__restrict__ tile_last, const uint32_t *__restrict__ tile_src,
                  const uint32_t *__restrict__ tile_cached, uint32_t *__restrict__ order) {
    __shared__ uint32_t hist[ORDER_BINS];
    __shared__ uint32_t part[1024];
    for (int b = threadIdx.x; b < ORDER_BINS; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    auto bin_of = [&](int u) -> int {
        uint32_t w = 0;
#pragma unroll
        for (int q = 0; q < TILE_PATCHES; q++) w = max(w, tile_last[(size_t)u * TILE_PATCHES + q]);
        if (tile_src != nullptr && tile_src[u] != 0 && tile_cached != nullptr && tile_cached[(size_t)u * TILE_PATCHES] != 0) w = 0;
        const uint32_t b = (w + 7u) >> 3;                       // 8 records per bin
        return ORDER_BINS - 1 - (int)min(b, (uint32_t)(ORDER_BINS - 1));   // bin 0 = most work
    };
    for (int u = threadIdx.x; u < nunits; u += blockDim.x) atomicAdd(&hist[bin_of(u)], 1u);
    __syncthreads();
    // exclusive prefix sum over the bins: two bins per thread + a block scan of the per-thread sums
    const uint32_t a0 = hist[2 * threadIdx.x], a1 = hist[2 * threadIdx.x + 1];
    part[threadIdx.x] = a0 + a1;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t add = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += add;
        __syncthreads();
    }
    const uint32_t base = part[threadIdx.x] - (a0 + a1);
    hist[2 * threadIdx.x] = base;
    hist[2 * threadIdx.x + 1] = base + a0;
    __syncthreads();
    for (int u = threadIdx.x; u < nunits; u += blockDim.x) order[atomicAdd(&hist[bin_of(u)], 1u)] = (uint32_t)u;
}

}  // namespace fnx

using namespace fnx;

extern "C" {

size_t fnx_raster_geom_bytes(int32_t P, int32_t V) { return geom_bytes(P, V); }
size_t fnx_raster_image_bytes(int32_t W, int32_t H, int32_t V) { return image_bytes(W, H, V); }
size_t fnx_raster_binning_bytes(int64_t cap, int32_t C) { return binning_bytes(cap, C); }

int fnx_raster_forward(const fnx_raster_args *a, fnx_alloc_fn ag, void *cg, fnx_alloc_fn ab, void *cb, fnx_alloc_fn ai,
                       void *ci, float *out_color, float *out_depth, int32_t *radii, int64_t *num_rendered_host,
                       fnx_raster_scratch *scratch, fnx_stream_t stream) {
    int rc = validate(a);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (a->C == 3) return forward_impl<3>(a, ag, cg, ab, cb, ai, ci, out_color, out_depth, radii, num_rendered_host, scratch, st);
    return forward_impl<1>(a, ag, cg, ab, cb, ai, ci, out_color, out_depth, radii, num_rendered_host, scratch, st);
}
int fnx_raster_forward_ch1(const fnx_raster_args *a, fnx_alloc_fn ag, void *cg, fnx_alloc_fn ab, void *cb, fnx_alloc_fn ai,
                           void *ci, float *out_color, float *out_depth, int32_t *radii, int64_t *num_rendered_host,
                           fnx_raster_scratch *scratch, fnx_stream_t stream) {
    if (a && a->C != 1) { set_error("fnx_raster_forward_ch1 needs C == 1 (got %d)", a->C); return FNX_ERR_INVALID; }
    return fnx_raster_forward(a, ag, cg, ab, cb, ai, ci, out_color, out_depth, radii, num_rendered_host, scratch, stream);
}
int fnx_raster_forward_ch3(const fnx_raster_args *a, fnx_alloc_fn ag, void *cg, fnx_alloc_fn ab, void *cb, fnx_alloc_fn ai,
                           void *ci, float *out_color, float *out_depth, int32_t *radii, int64_t *num_rendered_host,
                           fnx_raster_scratch *scratch, fnx_stream_t stream) {
    if (a && a->C != 3) { set_error("fnx_raster_forward_ch3 needs C == 3 (got %d)", a->C); return FNX_ERR_INVALID; }
    return fnx_raster_forward(a, ag, cg, ab, cb, ai, ci, out_color, out_depth, radii, num_rendered_host, scratch, stream);
}

int fnx_raster_backward(const fnx_raster_args *a, const fnx_raster_scratch *scratch, int64_t num_rendered,
                        const int32_t *radii, const float *dL_dout_color, const fnx_raster_grads *g, fnx_stream_t stream) {
    int rc = validate(a, /*backward=*/true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (a->C == 3) return backward_impl<3>(a, scratch, num_rendered, radii, dL_dout_color, g, st);
    return backward_impl<1>(a, scratch, num_rendered, radii, dL_dout_color, g, st);
}
int fnx_raster_backward_ch1(const fnx_raster_args *a, const fnx_raster_scratch *scratch, int64_t num_rendered,
                            const int32_t *radii, const float *dL_dout_color, const fnx_raster_grads *g, fnx_stream_t stream) {
    if (a && a->C != 1) { set_error("fnx_raster_backward_ch1 needs C == 1 (got %d)", a->C); return FNX_ERR_INVALID; }
    return fnx_raster_backward(a, scratch, num_rendered, radii, dL_dout_color, g, stream);
}
int fnx_raster_backward_ch3(const fnx_raster_args *a, const fnx_raster_scratch *scratch, int64_t num_rendered,
                            const int32_t *radii, const float *dL_dout_color, const fnx_raster_grads *g, fnx_stream_t stream) {
    if (a && a->C != 3) { set_error("fnx_raster_backward_ch3 needs C == 3 (got %d)", a->C); return FNX_ERR_INVALID; }
    return fnx_raster_backward(a, scratch, num_rendered, radii, dL_dout_color, g, stream);
}

int fnx_raster_blend_merged(const fnx_raster_args *dyn_args, const fnx_raster_scratch *dyn, const fnx_raster_scratch *stat,
                            int32_t P_static, void *merged_records, float *out_color, float *out_depth, fnx_stream_t stream) {
    return blend_merged(dyn_args, dyn, stat, P_static, merged_records, out_color, out_depth, (cudaStream_t)stream);
}
int fnx_raster_static_prepare(const fnx_raster_args *static_args, const fnx_raster_scratch *stat, float *out_color, float *out_depth,
                              fnx_stream_t stream) {
    return static_prepare(static_args, stat, out_color, out_depth, (cudaStream_t)stream);
}
int fnx_raster_read_tiles(const fnx_raster_scratch *scratch, int32_t W, int32_t H, int32_t V, int32_t merged, uint32_t *ranges,
                          uint32_t *tile_last, uint32_t *tile_src, uint32_t *tile_dyn_last, fnx_stream_t stream) {
    FNX_REQUIRE(scratch && scratch->image, "no image scratch");
    cudaStream_t st = (cudaStream_t)stream;
    ImageView im = image_view(scratch->image, W, H, V);
    const size_t nt = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE) * V;
    if (ranges) FNX_CUDA_TRY(cudaMemcpyAsync(ranges, merged ? im.mranges : im.ranges, nt * 8, cudaMemcpyDeviceToDevice, st));
    if (tile_last) FNX_CUDA_TRY(cudaMemcpyAsync(tile_last, im.tile_last, nt * 4 * TILE_PATCHES, cudaMemcpyDeviceToDevice, st));
    if (tile_src) FNX_CUDA_TRY(cudaMemcpyAsync(tile_src, im.tile_src, nt * 4, cudaMemcpyDeviceToDevice, st));
    if (tile_dyn_last) FNX_CUDA_TRY(cudaMemcpyAsync(tile_dyn_last, im.tile_dyn_last, nt * 4, cudaMemcpyDeviceToDevice, st));
    return FNX_OK;
}
int fnx_raster_overflow_flag(const fnx_raster_scratch *scratch, const int32_t **flag_dev) {
    FNX_REQUIRE(scratch && scratch->geom && flag_dev, "bad arguments");
    *flag_dev = &reinterpret_cast<const GeomHeader *>(scratch->geom)->overflow;
    return FNX_OK;
}
int fnx_raster_tile_cache_set(const fnx_raster_scratch *dyn, int32_t W, int32_t H, int32_t V, int32_t valid, fnx_stream_t stream) {
    FNX_REQUIRE(dyn && dyn->image && W > 0 && H > 0 && V > 0, "bad arguments");
    ImageView im = image_view(dyn->image, W, H, V);
    const size_t nt = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE) * V;
    // (0x01010101 != 0 reads as "valid" just like 1: the flags are only ever compared with zero)
    FNX_CUDA_TRY(cudaMemsetAsync(im.tile_cached, valid ? 0x01 : 0x00, sizeof(uint32_t) * nt * TILE_PATCHES, (cudaStream_t)stream));
    return FNX_OK;
}
int fnx_raster_tile_order(const fnx_raster_scratch *scratch, const fnx_raster_scratch *stat, int32_t W, int32_t H, int32_t V,
                          uint32_t *order, fnx_stream_t stream) {
    FNX_REQUIRE(scratch && scratch->image && order && W > 0 && H > 0 && V > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    ImageView im = image_view(scratch->image, W, H, V);
    const int nunits = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE) * V;
    const uint32_t *src = nullptr, *cached = nullptr;
    if (stat != nullptr) {   // merged streams: tiles served from the tile cache did no work
        src = im.tile_src;
        cached = im.tile_cached;
    }
    tile_order_kernel<<<1, 1024, 0, st>>>(nunits, im.tile_last, src, cached, order);
    FNX_LAUNCH_CHECK("tile_order_kernel");
    return FNX_OK;
}
int fnx_raster_backward_merged(const fnx_raster_args *dyn_args, const fnx_raster_scratch *dyn, const fnx_raster_scratch *stat,
                               const void *merged_records, const int32_t *radii, const float *dL_dout_color,
                               const fnx_raster_grads *g, fnx_stream_t stream) {
    return backward_merged(dyn_args, dyn, stat, merged_records, radii, dL_dout_color, g, (cudaStream_t)stream);
}

int fnx_raster_check(const fnx_raster_scratch *scratch, int64_t *num_rendered_host, fnx_stream_t stream) {
    (void)stream;
    FNX_REQUIRE(scratch && num_rendered_host, "scratch / num_rendered_host must be given");
    long long R = -1;
    if (scratch->check_slot >= 0) {
        FNX_REQUIRE(scratch->reserved >= 0 && scratch->reserved < MAX_DEVICES, "no forward to check");
        PinnedSlots &g_slots = g_slots_of_device[scratch->reserved];
        FNX_REQUIRE(g_slots.ok && scratch->check_slot < PinnedSlots::N, "no forward to check");
        FNX_CUDA_TRY(cudaEventSynchronize(g_slots.ev[scratch->check_slot]));
        R = g_slots.host[scratch->check_slot];
    } else {  // FNX_NO_HOST_SYNC forward: read the device-side header (blocks on the stream)
        FNX_REQUIRE(scratch->geom, "no forward to check");
        GeomHeader h;
        FNX_CUDA_TRY(cudaMemcpyAsync(&h, scratch->geom, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        FNX_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
        R = h.num_rendered;
    }
    *num_rendered_host = R;
    if (R > scratch->binning_capacity) {
        set_error("instance capacity %lld too small for %lld instances", (long long)scratch->binning_capacity, R);
        return FNX_ERR_CAPACITY;
    }
    return FNX_OK;
}

int fnx_raster_read_geom(const fnx_raster_scratch *scratch, int32_t P, int32_t V, float *xy, float *depth,
                         float *conic_opacity, uint32_t *tiles_touched, fnx_stream_t stream) {
    FNX_REQUIRE(scratch && scratch->geom, "no geom scratch");
    cudaStream_t st = (cudaStream_t)stream;
    GeomView g = geom_view(scratch->geom, P, V);
    const size_t n = (size_t)P * V;
    if (xy) FNX_CUDA_TRY(cudaMemcpyAsync(xy, g.xy, n * 8, cudaMemcpyDeviceToDevice, st));
    if (depth) FNX_CUDA_TRY(cudaMemcpyAsync(depth, g.depth, n * 4, cudaMemcpyDeviceToDevice, st));
    if (conic_opacity) FNX_CUDA_TRY(cudaMemcpyAsync(conic_opacity, g.conic_o, n * 16, cudaMemcpyDeviceToDevice, st));
    if (tiles_touched) FNX_CUDA_TRY(cudaMemcpyAsync(tiles_touched, g.tiles_touched, n * 4, cudaMemcpyDeviceToDevice, st));
    return FNX_OK;
}
int fnx_raster_read_image(const fnx_raster_scratch *scratch, int32_t W, int32_t H, int32_t V, float *final_T,
                          uint32_t *n_contrib, fnx_stream_t stream) {
    FNX_REQUIRE(scratch && scratch->image, "no image scratch");
    cudaStream_t st = (cudaStream_t)stream;
    ImageView im = image_view(scratch->image, W, H, V);
    const size_t n = (size_t)W * H * V;
    if (final_T) FNX_CUDA_TRY(cudaMemcpyAsync(final_T, im.final_T, n * 4, cudaMemcpyDeviceToDevice, st));
    if (n_contrib) FNX_CUDA_TRY(cudaMemcpyAsync(n_contrib, im.n_contrib, n * 4, cudaMemcpyDeviceToDevice, st));
    return FNX_OK;
}

int fnx_mark_visible(int32_t P, const float *means3D, const float *view_matrix, const float *proj_matrix, uint8_t *present,
                     fnx_stream_t stream) {
    (void)proj_matrix;
    if (P == 0) return FNX_OK;
    FNX_REQUIRE(P > 0 && means3D && view_matrix && present, "bad arguments");
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, means3D, view_matrix, present);
    FNX_LAUNCH_CHECK("mark_visible_kernel");
    return FNX_OK;
}

}  // extern "C"
