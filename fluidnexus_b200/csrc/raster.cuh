// raster.cuh -- private layout of the rasterizer scratch + kernel launch prototypes.
#pragma once
#include "common.cuh"

namespace fnx {

constexpr int TILE = 16;                 // 16x16 pixel tiles (R3/cuda_rasterizer/config.h:16-17)
constexpr int TILE_PIX = TILE * TILE;
constexpr int PATCH = 8;                 // a warp of the blend kernels owns an 8x8 pixel patch of a tile
constexpr int TILE_PATCHES = (TILE / PATCH) * (TILE / PATCH);  // 4
constexpr float DEPTH_DEFAULT = 15.0f;   // forward.cu:295
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float T_EPS = 0.0001f;

// Per-instance record in the depth-sorted, tile-partitioned stream the blend kernels consume.
// One contiguous span per tile => a span is a 1-D bulk (TMA) copy into shared memory.
//   C == 3 : 48 B  {x, y, a, b | c, opacity, r, g | b, slot(int bits), depth, pad}
//   C == 1 : 32 B  {x, y, a, b | c, opacity, col, slot(int bits)}   (median depth is fetched from geom.depth[slot])
template <int C>
struct RecBytes {
    static constexpr int value = (C == 3) ? 48 : 32;
};

// accumulator floats per (view, Gaussian) written by the blend backward with vector reductions:
//   {dmean2D.x, dmean2D.y, dconic.x, dconic.y | dconic.w, dopacity, dcol0, dcol1 | dcol2, -, -, -}
template <int C>
struct AccFloats {
    static constexpr int value = (C == 3) ? 12 : 8;
};

struct GeomHeader {
    long long num_rendered;   // instances actually needed (device-written by the scan epilogue)
    long long capacity;       // instances the binning buffers can hold
    int overflow;             // 1 if num_rendered > capacity (nothing rendered)
    int pad;
    unsigned long long merge_cursor;  // bump allocator of fnx_raster_blend_merged's merged stream (reset per forward)
    int static_prepared;      // static stream only: fnx_raster_static_prepare has blended it (tile_last / tile_cached valid)
    int pad2;
};

struct GeomView {  // pointers into the geom scratch
    GeomHeader *hdr;
    float *cov3D;             // [P,6]
    float *depth;             // [V*P]
    float2 *xy;               // [V*P]
    float4 *conic_o;          // [V*P]
    uint32_t *tiles_touched;  // [V*P] (after tile culling)
    unsigned long long *dkeys_in, *dkeys_out;  // [V*P] (view<<32 | depth bits)
    uint32_t *dvals_in, *dvals_out;            // [V*P] slot = v*P + i
    uint32_t *offsets;        // [V*P] exclusive scan of tiles_touched in depth order
    float *accum;             // [V*P*AccFloats] (backward)
    void *cub_temp;
    size_t cub_temp_bytes;
};

struct BinView {
    uint32_t *tkeys_in, *tkeys_out;  // [cap] global tile id = v*ntiles + tile
    uint32_t *tvals_in, *tvals_out;  // [cap] slot
    char *records;                   // [cap * RecBytes]
    unsigned long long *bkeys;       // [cap] bucket binning: (depth bits << 32 | slot), grouped by tile, unsorted inside a tile
    unsigned long long *bkeys2;      // [cap] sorted copy, used only by tiles whose bucket exceeds the shared-memory sort
    void *cub_temp;
    size_t cub_temp_bytes;
};

struct ImageView {
    float *final_T;          // [V*H*W]
    uint32_t *n_contrib;     // [V*H*W]
    uint2 *ranges;           // [V*ntiles]
    uint32_t *tile_last;     // [V*ntiles*4] per 8x8 patch: max n_contrib over its pixels (backward start)
    uint2 *mranges;          // [V*ntiles] merged ranges (static + dynamic streams), see fnx_raster_blend_merged
    uint32_t *tile_src;      // [V*ntiles] 0: the tile's span lives in the call's own record stream, 1: in the static one
    uint32_t *tile_dyn_last; // [V*ntiles] merged streams: 1 + span index of the tile's last dynamic record (0: none)
    uint32_t *tile_dyn_first; // [V*ntiles] merged streams: span index of the tile's FIRST dynamic record (everything in front is frozen)
    uint32_t *tile_cached;   // [V*ntiles*4] static stream, per patch: 1 <=> the caller's out_color/out_depth hold its static-only render
    float4 *snap;            // [V*H*W] merged streams: {T, colour behind} right after the tile's last dynamic record
    uint32_t *tile_count;    // [V*ntiles] bucket binning: instances per tile (histogram written by the preprocess)
    uint32_t *tile_cursor;   // [V*ntiles] bucket binning: fill cursor of the tile's bucket
};

size_t geom_bytes(int P, int V);
size_t image_bytes(int W, int H, int V);
size_t binning_bytes(long long cap, int C);
GeomView geom_view(void *chunk, int P, int V);
ImageView image_view(void *chunk, int W, int H, int V);
BinView bin_view(void *chunk, long long cap, int C);

}  // namespace fnx
