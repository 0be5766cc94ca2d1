// pbf.cu -- fixed-radius neighbour grid + fused particle-physics terms for sm_100a (B200).
//
// Replaces, on the FluidDynamics training hot path (FD = FluidDynamics in the reference tree):
//   torch_cluster.radius / radius_graph (torch-cluster 1.6.3, brute force O(N*M) scan per query) used by
//       FD/gaussian_splatting/gm_fluid.py:1114,1140,1301
//   the ~12-kernel gather / poly6 / index_add_ chains (+ their autograd twins) of
//       P1 get_visual_xyz_from_nn                gm_fluid.py:1291-1336
//       P2 get_gas_constraints_from_exyz_nn      gm_fluid.py:1107-1132
//       P3 get_gas_constraints_from_vel_nn_guess gm_fluid.py:1134-1158, :846-862
//   P5 distance_loss (dense cdist, O(V^2))       FD/utils/loss_utils.py:98-121
//   simple_knn distCUDA2                          FD/submodules/simple-knn/simple_knn.cu:134-202
//   torch_scatter.scatter_min                     gm_fluid.py:1088,1272
// Design: a hashed uniform grid (cell = search radius) built with a counting sort; every term is a *gather* over the 27
// neighbouring cells, walked as 9 x-contiguous bucket rows by 8 lanes per query, with no materialised edge list and no atomics in the gradient passes (deterministic).
// torch_cluster's `max_num_neighbors` rule ("first K hits in index order") is kept exactly through a per-query
// index cut-off kth[c] (= K-th smallest neighbour index, INT_MAX when fewer than K+1 neighbours exist):
//   edge (j -> c) exists  <=>  |x_j - y_c|^2 < r^2  and  j <= kth[c].
#include <cub/cub.cuh>

#include "common.cuh"

namespace fnx {

struct GridHeader {
    int n, M;  // points, hash table size (power of two)
    float cell, inv_cell;
};
struct GridView {
    GridHeader *hdr;
    uint32_t *bucket_start;  // [M+1]
    uint32_t *bucket_fill;   // [M]
    uint32_t *sorted_idx;    // [n]
    uint32_t *tmp_idx;       // [n] bucket contents in atomic arrival order (before the per-bucket index sort)
    float4 *sorted_pos;      // [n] xyz + idx bits
    float4 *aux0, *aux1;     // [n] per-point payloads in the SAME (bucket-sorted) order, packed by a gather kernel's
                             // pre-pass so that the walk reads them next to sorted_pos[a] instead of through the index
    void *cub_temp;
    size_t cub_temp_bytes;
    int M;
};

static int table_size(int n) {
    int M = 1024;
    while (M < 2 * n) M <<= 1;
    return M;
}
static GridView grid_view(void *chunk, int n) {
    GridView g;
    char *p = (char *)chunk;
    g.M = table_size(n);
    g.hdr = carve<GridHeader>(p, 1);
    g.bucket_start = carve<uint32_t>(p, (size_t)g.M + 1);
    g.bucket_fill = carve<uint32_t>(p, (size_t)g.M);
    g.sorted_idx = carve<uint32_t>(p, (size_t)(n > 0 ? n : 1));
    g.tmp_idx = carve<uint32_t>(p, (size_t)(n > 0 ? n : 1));
    g.sorted_pos = carve<float4>(p, (size_t)(n > 0 ? n : 1));
    g.aux0 = carve<float4>(p, (size_t)(n > 0 ? n : 1));
    g.aux1 = carve<float4>(p, (size_t)(n > 0 ? n : 1));
    // memoised CUB size query (its dispatch layer is slow); a handful of distinct table sizes per process
    static thread_local int cached_M[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    static thread_local size_t cached[8];
    size_t tb = 0;
    bool hit = false;
    for (int k = 0; k < 8; k++)
        if (cached_M[k] == g.M) { tb = cached[k]; hit = true; break; }
    if (!hit) {
        cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint32_t *)nullptr, (uint32_t *)nullptr, g.M + 1);
        static thread_local int next = 0;
        cached_M[next] = g.M; cached[next] = tb; next = (next + 1) & 7;
    }
    g.cub_temp_bytes = tb;
    g.cub_temp = carve<char>(p, tb);
    return g;
}
static size_t grid_bytes(int n) {
    GridView g = grid_view((void *)0, n);
    return (size_t)((char *)g.cub_temp - (char *)0) + g.cub_temp_bytes + 256;
}

// The table is built on sub-cells of edge cell / SUB (callers pass the search cell, cell >= r); a query visits (2 SUB + 1)^3
// sub-cells as (2 SUB + 1)^2 x-contiguous rows (walk_neighbors).  SUB = 1: 9 rows of 3 cells, ~33 candidates per row at the
// reference's lattice spacing.  SUB = 2 (125 sub-cells, 15.6 r^3 of candidates instead of 27 r^3 -- the sphere is 4.2 r^3) was
// measured and is SLOWER: 25 rows of ~7 candidates leave the 8 lanes of a query group mostly idle and triple the dependent
// range loads (smoke bench, round 2: SUB = 1 2165 it/s, SUB = 2 1929 it/s; the per-cell walk this replaced: 2072 it/s).
#ifndef FNX_GRID_SUB
#define FNX_GRID_SUB 1
#endif
constexpr int SUB = FNX_GRID_SUB;
constexpr int REACH = SUB;                                   // sub-cells visited on each side of the query's
constexpr int ROW_CELLS = 2 * REACH + 1;                     // x-contiguous buckets of one (dy, dz) row
constexpr int ROWS = ROW_CELLS * ROW_CELLS;                  // rows per query
__device__ __forceinline__ float sub_inv(float inv_cell) { return inv_cell * (float)SUB; }   // exact (power of two)
__device__ __forceinline__ int3 cell_of(float x, float y, float z, float inv_sub) {
    return make_int3(__float2int_rd(x * inv_sub), __float2int_rd(y * inv_sub), __float2int_rd(z * inv_sub));
}
// Sub-cell -> bucket: the table is a periodic A x B x C box of sub-cells (powers of two, each >= 8, A*B*C = M): bucket =
// (x mod A, y mod B, z mod C), x in the low bits.  Unlike a multiplicative hash this makes the (2 REACH + 1)^3 sub-cells around
// any query map to DIFFERENT buckets (2 REACH + 1 <= 5 <= 8), and a point that shares a bucket with a visited sub-cell without
// being in it lies at least A - 2 REACH - 1 sub-cells (>= 5 cells at SUB = 1, 1.5 cells at SUB = 2) away along some axis, i.e.
// farther than the search radius (<= one cell): gathers therefore need no "is this point really in that sub-cell" test and no
// de-duplication, the distance test does it all.
// With x in the low bits the ROW_CELLS sub-cells of one (dy, dz) row are CONTIGUOUS buckets (one [start, end) range of the
// counting sort) unless the row wraps around the period A, when they are two ranges (walk_neighbors).
__device__ __forceinline__ uint32_t hash_cell(int3 c, int M) {
    const int m = 31 - __clz(M);          // M = 2^m, m >= 10
    const int ax = m / 3, az = m / 3, ay = m - ax - az;
    return ((uint32_t)c.x & ((1u << ax) - 1u)) | (((uint32_t)c.y & ((1u << ay) - 1u)) << ax) |
           (((uint32_t)c.z & ((1u << az) - 1u)) << (ax + ay));
}

__global__ void grid_count_kernel(const float *__restrict__ pts, int n, float inv_cell, int M, uint32_t *__restrict__ fill,
                                  GridHeader *__restrict__ h, float cell, uint32_t *__restrict__ bucket_start) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {  // header + the scan's end sentinel (the scan writes entries [0, M))
        h->n = n; h->M = M; h->cell = cell; h->inv_cell = 1.0f / cell;
        bucket_start[M] = (uint32_t)n;
    }
    if (i >= n) return;
    atomicAdd(&fill[hash_cell(cell_of(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], inv_cell), M)], 1u);
}
__global__ void grid_scatter_kernel(const float *__restrict__ pts, int n, float inv_cell, int M,
                                    const uint32_t *__restrict__ start, uint32_t *__restrict__ fill,
                                    uint32_t *__restrict__ sorted_idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = hash_cell(cell_of(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], inv_cell), M);
    sorted_idx[start[b] + atomicAdd(&fill[b], 1u)] = (uint32_t)i;
}
// order every bucket by point index so that results do not depend on atomic arrival order: each entry finds its
// rank inside its bucket (buckets are tiny) and writes itself + the position copy the gathers read
__global__ void grid_finalize_kernel(const float *__restrict__ pts, int n, float inv_cell, int M, const uint32_t *__restrict__ start,
                                     const uint32_t *__restrict__ tmp_idx, uint32_t *__restrict__ sorted_idx,
                                     float4 *__restrict__ sorted_pos) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const uint32_t i = tmp_idx[a];
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const uint32_t b = hash_cell(cell_of(x, y, z, inv_cell), M);
    const uint32_t s = start[b], e = start[b + 1];
    uint32_t rank = 0;
    for (uint32_t k = s; k < e; k++) rank += tmp_idx[k] < i ? 1u : 0u;
    sorted_idx[s + rank] = i;
    sorted_pos[s + rank] = make_float4(x, y, z, __uint_as_float(i));
}
__global__ void grid_header_kernel(GridHeader *h, int n, int M, float cell, uint32_t *bucket_start) {
    h->n = n; h->M = M; h->cell = cell; h->inv_cell = 1.0f / cell;
    bucket_start[M] = (uint32_t)n;  // the scan below writes entries [0, M)
}

static int grid_build(const float *pts, int n, float cell, void *scratch, cudaStream_t st) {
    GridView g = grid_view(scratch, n);
    const float inv_cell = (1.0f / cell) * (float)SUB;   // of the SUB-cell: the same expression the walkers use (sub_inv)
    FNX_CUDA_TRY(cudaMemsetAsync(g.bucket_fill, 0, sizeof(uint32_t) * (size_t)g.M, st));
    if (n > 0) {
        grid_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(pts, n, inv_cell, g.M, g.bucket_fill, g.hdr, cell, g.bucket_start);
        FNX_LAUNCH_CHECK("grid_count_kernel");
    } else {
        grid_header_kernel<<<1, 1, 0, st>>>(g.hdr, n, g.M, cell, g.bucket_start);
        FNX_LAUNCH_CHECK("grid_header_kernel");
    }
    // (a single-CTA scan of the 65 536 counters that also clears them was tried in round 2 to save CUB's two launches: 16.5 us per
    // build against 8 us for the library's decoupled look-back pair -- one CTA cannot hide the memory latency -- and no gain in
    // the step: reverted)
    size_t tb = g.cub_temp_bytes;
    FNX_CUDA_TRY(cub::DeviceScan::ExclusiveSum(g.cub_temp, tb, g.bucket_fill, g.bucket_start, g.M, st));
    FNX_CUDA_TRY(cudaMemsetAsync(g.bucket_fill, 0, sizeof(uint32_t) * (size_t)g.M, st));
    if (n > 0) {
        grid_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(pts, n, inv_cell, g.M, g.bucket_start, g.bucket_fill, g.tmp_idx);
        FNX_LAUNCH_CHECK("grid_scatter_kernel");
        grid_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(pts, n, inv_cell, g.M, g.bucket_start, g.tmp_idx, g.sorted_idx, g.sorted_pos);
        FNX_LAUNCH_CHECK("grid_finalize_kernel");
    }
    return FNX_OK;
}

// Visit every grid point within sqrt(r2) of q: f(j, pj, d2, a) with a = the point's position in the bucket-sorted arrays.
// Requires cell >= r (see hash_cell for why the distance test alone is exact).  LANES consecutive lanes share ONE query and
// stride together over each row's contiguous candidates (coalesced float4 loads, no lane idles while another walks a fuller
// bucket); the next row's ranges are fetched before the current row is walked.  Callers reduce their per-lane partials with
// group_sum().  LANES = 1: one thread per query.
template <int LANES, typename F>
__device__ __forceinline__ void walk_neighbors(const GridView &g, float inv_cell, float3 q, float r2, int lane, F f) {
    const int3 c = cell_of(q.x, q.y, q.z, sub_inv(inv_cell));
    const int m = 31 - __clz(g.M);   // bucket = x | y << ax | z << (ax + ay), see hash_cell
    const int ax = m / 3, ay = m - 2 * ax;
    const uint32_t my = (1u << ay) - 1u, mz = (1u << ax) - 1u;
    const int A = 1 << ax;
    // the row's x-extent: sub-cells x0 .. x0 + len1 - 1 and, when it wraps around the period, 0 .. ROW_CELLS - len1 - 1
    const int x0 = (c.x - REACH) & (A - 1);
    const int len1 = min(ROW_CELLS, A - x0);
    const uint32_t *__restrict__ bs = g.bucket_start;
    // bucket ranges of row (dy, dz): [x, y) and, for a wrapped row, [z, w)   (bs[base + A] is the next row's first entry, bs[M] = n)
    auto ranges = [&](int dy, int dz) {
        const uint32_t base = (((uint32_t)(c.y + dy) & my) << ax) | (((uint32_t)(c.z + dz) & mz) << (ax + ay));
        uint4 r;
        r.x = __ldg(bs + base + x0);
        r.y = __ldg(bs + base + x0 + len1);
        r.z = r.w = 0u;
        if (len1 < ROW_CELLS) {
            r.z = __ldg(bs + base);
            r.w = __ldg(bs + base + (ROW_CELLS - len1));
        }
        return r;
    };
    int dy = -REACH, dz = -REACH;
    uint4 next = ranges(dy, dz);
#pragma unroll 1
    for (int row = 0; row < ROWS; row++) {
        const uint4 cur = next;
        if (++dy > REACH) { dy = -REACH; dz++; }
        if (row + 1 < ROWS) next = ranges(dy, dz);
#pragma unroll 1
        for (int seg = 0; seg < 2; seg++) {
            const uint32_t s = seg == 0 ? cur.x : cur.z, e = seg == 0 ? cur.y : cur.w;
            for (uint32_t a = s + lane; a < e; a += LANES) {
                const float4 p = __ldg(g.sorted_pos + a);
                const float ex = p.x - q.x, ey = p.y - q.y, ez = p.z - q.z;
                const float d2 = ex * ex + ey * ey + ez * ez;
                if (d2 < r2) f((int)__float_as_uint(p.w), p, d2, a);
            }
        }
    }
}
template <typename F>
__device__ __forceinline__ void for_each_neighbor(const GridView &g, float inv_cell, float3 q, float r2, F f) {
    walk_neighbors<1>(g, inv_cell, q, r2, 0, [&](int j, const float4 &p, float d2, uint32_t) { f(j, p, d2); });
}

// Group-cooperative variant: GROUP = 8 consecutive lanes share ONE query.  One thread per query is latency bound (28k queries x
// 9 dependent range walks of ~33 candidates); a whole warp per query leaves half of the lanes idle on a row and gives only 4
// independent queries per 128 threads; 8 lanes walk a row in 4-5 trips and give 16.
#ifndef FNX_GROUP
#define FNX_GROUP 8
#endif
constexpr int GROUP = FNX_GROUP;
constexpr int QPB = 128 / GROUP;  // queries per 128-thread block
// The density kernels (P2 / P3) run on the step's side stream, off the critical path of an iteration: they use groups of 4 lanes --
// fewer idle lanes on a row's last trip (33 candidates: 9 trips of 4 against 5 trips of 8), i.e. fewer instructions for the same
// work, at the price of a longer dependent chain per query.  Measured with ALL gathers at 4 lanes (smoke / scalar): throughput
// +2.2 % / +1.1 %, one frame alone -2 % / -6 % (the advection gathers sit on the critical path): so 8 lanes there, 4 here.
#ifndef FNX_GROUP_SIDE
#define FNX_GROUP_SIDE 4
#endif
constexpr int GROUP_SIDE = FNX_GROUP_SIDE;
constexpr int QPB_SIDE = 128 / GROUP_SIDE;
template <int G = GROUP, typename F>
__device__ __forceinline__ void warp_for_each_neighbor(const GridView &g, float inv_cell, float3 q, float r2, int lane, F f) {
    walk_neighbors<G>(g, inv_cell, q, r2, lane, f);
}
// sums over the G lanes of a query group (xor shuffles stay inside an aligned group; only the group's lanes are named
// in the mask, so groups of one warp may diverge)
template <int G = GROUP>
__device__ __forceinline__ unsigned group_mask() { return ((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)); }
template <int G = GROUP>
__device__ __forceinline__ float group_sum(float v) {
    const unsigned m = group_mask<G>();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}
template <int G = GROUP>
__device__ __forceinline__ int group_sum(int v) {
    const unsigned m = group_mask<G>();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// neighbour counts + the index cut-off that realises torch_cluster's max_num_neighbors rule
// ---------------------------------------------------------------------------------------------------------------
// K-th smallest neighbour index of query q by bisection on the index value (rare path: only when more than K points
// lie within the radius).  Warp-cooperative; every lane returns the same value.
template <int G = GROUP>
__device__ __forceinline__ int kth_by_bisection(const GridView &g, float inv_cell, float3 q, float r2, int lane, int K, int n_x) {
    int lo = 0, hi = n_x - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        int below = 0;
        warp_for_each_neighbor<G>(g, inv_cell, q, r2, lane, [&](int j, const float4 &, float, uint32_t) { below += (j <= mid); });
        below = group_sum<G>(below);
        if (below >= K) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__global__ void __launch_bounds__(128)
radius_count_kernel(GridView g, float inv_cell, const float *__restrict__ y, int ny, float r2, int K, int n_x,
                    int *__restrict__ counts, int *__restrict__ kth) {
    const int c = blockIdx.x * QPB + (threadIdx.x / GROUP);
    const int lane = threadIdx.x % GROUP;
    if (c >= ny) return;
    const float3 q = make_float3(y[3 * c], y[3 * c + 1], y[3 * c + 2]);
    int cnt = 0;
    warp_for_each_neighbor(g, inv_cell, q, r2, lane, [&](int, const float4 &, float, uint32_t) { cnt++; });
    cnt = group_sum(cnt);
    int cut = 0x7fffffff;
    if (cnt > K) {
        cut = kth_by_bisection(g, inv_cell, q, r2, lane, K, n_x);
        cnt = K;
    }
    if (lane == 0) {
        if (counts) counts[c] = cnt;
        if (kth) kth[c] = cut;
    }
}

// fills a torch_cluster-style edge list: for query c the (<= K) neighbours in ascending index order
__global__ void __launch_bounds__(128)
radius_fill_kernel(GridView g, float inv_cell, const float *__restrict__ y, int ny, float r2, const int *__restrict__ kth,
                   const long long *__restrict__ offsets, long long *__restrict__ edge_query,
                   long long *__restrict__ edge_x) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ny) return;
    const float3 q = make_float3(y[3 * c], y[3 * c + 1], y[3 * c + 2]);
    const int cut = kth[c];
    long long o = offsets[c];
    const long long o0 = o;
    for_each_neighbor(g, inv_cell, q, r2, [&](int j, const float4 &, float) {
        if (j <= cut) {
            edge_query[o] = c;
            edge_x[o] = j;
            o++;
        }
    });
    // ascending index order within the query (insertion sort; lists are short)
    for (long long a = o0 + 1; a < o; a++) {
        const long long key = edge_x[a];
        long long k = a;
        while (k > o0 && edge_x[k - 1] > key) {
            edge_x[k] = edge_x[k - 1];
            k--;
        }
        edge_x[k] = key;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// P2 / P3: SPH poly6 density ratio and its gradient
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float poly6(float d2, float H2, float term1) {
    const float t = H2 - d2;
    return d2 < H2 ? term1 * (t * t * t) : 0.f;
}
__device__ __forceinline__ float dpoly6_dd2(float d2, float H2, float term1) {
    const float t = H2 - d2;
    return d2 < H2 ? -3.f * term1 * (t * t) : 0.f;
}

// p_ratio[r] = (sum_{c in N(r), r <= kth[c]} poly6(|X_r - X_c|^2)) / imass[r] / p0      (gm_fluid.py:1117-1130)
__global__ void __launch_bounds__(128)
density_fwd_kernel(GridView g, float inv_cell, const float *__restrict__ X, int N, const float *__restrict__ imass,
                   const int *__restrict__ kth, float H2, float term1, float p0, float *__restrict__ p_ratio) {
    const int r = blockIdx.x * QPB_SIDE + (threadIdx.x / GROUP_SIDE);
    const int lane = threadIdx.x % GROUP_SIDE;
    if (r >= N) return;
    const float3 q = make_float3(X[3 * r], X[3 * r + 1], X[3 * r + 2]);
    float pi = 0.f;
    warp_for_each_neighbor<GROUP_SIDE>(g, inv_cell, q, H2, lane, [&](int c, const float4 &, float d2, uint32_t) {
        if (r <= kth[c]) pi += poly6(d2, H2, term1);
    });
    pi = group_sum<GROUP_SIDE>(pi);
    if (lane == 0) p_ratio[r] = pi / imass[r] / p0;
}

// Neighbour count, cut-off index and density in ONE walk.  The density written here assumes that no neighbour's cap
// binds (r <= kth[c] for every c); *cap_flag is raised when any particle has more than K neighbours, and the exact
// kernel above then recomputes p_ratio with the cut-offs (it returns at once while the flag is down -- the normal case:
// ~46 neighbours at the reference's lattice spacing against K = 100).
__global__ void __launch_bounds__(128)
density_fwd_counted_kernel(GridView g, float inv_cell, const float *__restrict__ X, int N, const float *__restrict__ imass, int K,
                           float H2, float term1, float p0, int *__restrict__ kth, float *__restrict__ p_ratio,
                           int *__restrict__ cap_flag) {
    const int r = blockIdx.x * QPB_SIDE + (threadIdx.x / GROUP_SIDE);
    const int lane = threadIdx.x % GROUP_SIDE;
    if (r >= N) return;
    const float3 q = make_float3(X[3 * r], X[3 * r + 1], X[3 * r + 2]);
    float pi = 0.f;
    int cnt = 0;
    warp_for_each_neighbor<GROUP_SIDE>(g, inv_cell, q, H2, lane, [&](int, const float4 &, float d2, uint32_t) {
        cnt++;
        pi += poly6(d2, H2, term1);
    });
    pi = group_sum<GROUP_SIDE>(pi);
    cnt = group_sum<GROUP_SIDE>(cnt);
    int cut = 0x7fffffff;
    if (cnt > K) {
        cut = kth_by_bisection<GROUP_SIDE>(g, inv_cell, q, H2, lane, K, N);
        if (lane == 0) *cap_flag = 1;
    }
    if (lane == 0) {
        kth[r] = cut;
        p_ratio[r] = pi / imass[r] / p0;
    }
}
__global__ void __launch_bounds__(128)
density_fwd_if_capped_kernel(GridView g, float inv_cell, const float *__restrict__ X, int N, const float *__restrict__ imass,
                             const int *__restrict__ kth, float H2, float term1, float p0, float *__restrict__ p_ratio,
                             const int *__restrict__ cap_flag) {
    if (*cap_flag == 0) return;  // the normal case: a small fixed grid returns at once
    const int lane = threadIdx.x % GROUP_SIDE;
    for (int r = blockIdx.x * QPB_SIDE + (threadIdx.x / GROUP_SIDE); r < N; r += gridDim.x * QPB_SIDE) {  // (r is uniform over a query group)
        const float3 q = make_float3(X[3 * r], X[3 * r + 1], X[3 * r + 2]);
        float pi = 0.f;
        warp_for_each_neighbor<GROUP_SIDE>(g, inv_cell, q, H2, lane, [&](int c, const float4 &, float d2, uint32_t) {
            if (r <= kth[c]) pi += poly6(d2, H2, term1);
        });
        pi = group_sum<GROUP_SIDE>(pi);
        if (lane == 0) p_ratio[r] = pi / imass[r] / p0;
    }
}

// dL/dX_k = sum_{j in N(k)} dpoly6(d2) * 2 (X_k - X_j) * ( gp_k [k <= kth[j]] + gp_j [j <= kth[k]] ),
// gp_i = dL/dp_ratio_i / (imass_i p0)
// pre-pass: aux1[a] = { gp_i = dL/dp_ratio_i / (imass_i p0), kth_i } for the point i stored at sorted position a
__global__ void density_bwd_pack_kernel(GridView g, int N, const float *__restrict__ imass, const int *__restrict__ kth, float p0,
                                        const float *__restrict__ dL_dpratio) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= N) return;
    const uint32_t i = g.sorted_idx[a];
    g.aux1[a] = make_float4(dL_dpratio[i] / imass[i] / p0, __int_as_float(kth[i]), 0.f, 0.f);
}
// the same pre-pass for the training step, where dL/dp_ratio is that of loss = weight * mean((p_ratio - 1)^2)
// (train_physical_particle.py:336-342): computes it on the fly (and stores it: density_bwd_kernel reads it for its own particle),
// and accumulates the un-weighted loss -- the separate ratio_loss launch folded into a pass that exists anyway
__global__ void density_bwd_pack_ratio_kernel(GridView g, int N, const float *__restrict__ imass, const int *__restrict__ kth, float p0,
                                              const float *__restrict__ p_ratio, float weight, float *__restrict__ dL_dpratio,
                                              float *__restrict__ loss) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (a < N) {
        const uint32_t i = g.sorted_idx[a];
        const float d = p_ratio[i] - 1.0f;
        const float gp = weight * 2.0f * d / N;
        l = d * d;
        dL_dpratio[i] = gp;
        g.aux1[a] = make_float4(gp / imass[i] / p0, __int_as_float(kth[i]), 0.f, 0.f);
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(loss, l / N);
}
__global__ void __launch_bounds__(128)
density_bwd_kernel(GridView g, float inv_cell, const float *__restrict__ X, int N, const float *__restrict__ imass,
                   const int *__restrict__ kth, float H2, float term1, float p0, const float *__restrict__ dL_dpratio,
                   float *__restrict__ dL_dX, int accumulate) {
    const int k = blockIdx.x * QPB_SIDE + (threadIdx.x / GROUP_SIDE);
    const int lane = threadIdx.x % GROUP_SIDE;
    if (k >= N) return;
    const float3 q = make_float3(X[3 * k], X[3 * k + 1], X[3 * k + 2]);
    const float gpk = dL_dpratio[k] / imass[k] / p0;
    const int kth_k = kth[k];
    float3 acc = make_float3(0.f, 0.f, 0.f);
    warp_for_each_neighbor<GROUP_SIDE>(g, inv_cell, q, H2, lane, [&](int j, const float4 &pj, float d2, uint32_t a) {
        const float4 pay = g.aux1[a];  // { gp_j, kth_j }
        float w = 0.f;
        if (k <= __float_as_int(pay.y)) w += gpk;
        if (j <= kth_k) w += pay.x;
        const float s = 2.f * dpoly6_dd2(d2, H2, term1) * w;
        acc.x += s * (q.x - pj.x); acc.y += s * (q.y - pj.y); acc.z += s * (q.z - pj.z);
    });
    acc.x = group_sum<GROUP_SIDE>(acc.x); acc.y = group_sum<GROUP_SIDE>(acc.y); acc.z = group_sum<GROUP_SIDE>(acc.z);
    if (lane < 3) {
        const float v = lane == 0 ? acc.x : (lane == 1 ? acc.y : acc.z);
        if (accumulate) dL_dX[3 * k + lane] += v; else dL_dX[3 * k + lane] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// P1: advect visual particles with the poly6-interpolated hidden velocity (gm_fluid.py:1291-1336)
// ---------------------------------------------------------------------------------------------------------------
// pre-pass of P1: aux0[a] = u_j = (X_j - xyz_j) / secs for the hidden particle j stored at sorted position a
__global__ void advect_pack_u_kernel(GridView gh, int N, const float *__restrict__ xyz, float secs) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= N) return;
    const float4 p = gh.sorted_pos[a];
    const uint32_t j = __float_as_uint(p.w);
    gh.aux0[a] = make_float4((p.x - xyz[3 * j]) / secs, (p.y - xyz[3 * j + 1]) / secs, (p.z - xyz[3 * j + 2]) / secs, 0.f);
}
// COUNTED = false: the cut-off kthV[v] is an input (fnx_radius_count ran before).  COUNTED = true: the kernel counts the
// neighbours while it sums; only a query with more than K of them (rare) finds its cut-off by bisection and sums
// again with it, and kthV[v] is an OUTPUT for the backward.
template <bool COUNTED>
__global__ void __launch_bounds__(128)
advect_fwd_kernel(GridView gh, float inv_cell, const float *__restrict__ X /*100*e*/, const float *__restrict__ xyz,
                  const float *__restrict__ vis, int V, int *__restrict__ kthV, int K, int n_x, float H2, float term1, float secs,
                  float eps, float out_div, float *__restrict__ vis_out, float *__restrict__ num_out,
                  float *__restrict__ den_out) {
    const int v = blockIdx.x * QPB + (threadIdx.x / GROUP);
    const int lane = threadIdx.x % GROUP;
    if (v >= V) return;
    const float3 q = make_float3(vis[3 * v], vis[3 * v + 1], vis[3 * v + 2]);
    int cut = COUNTED ? 0x7fffffff : kthV[v];
    float3 num = make_float3(0.f, 0.f, 0.f);
    float den = 0.f;
    int cnt = 0;
    auto gather = [&](int j, const float4 &, float d2, uint32_t a) {
        cnt++;
        if (j > cut) return;
        const float w = poly6(d2, H2, term1);
        const float4 u = gh.aux0[a];  // hidden velocity (X_j - xyz_j) / secs, packed by advect_pack_u_kernel
        num.x += w * u.x;
        num.y += w * u.y;
        num.z += w * u.z;
        den += w;
    };
    warp_for_each_neighbor(gh, inv_cell, q, H2, lane, gather);
    if (COUNTED) {
        cnt = group_sum(cnt);
        if (cnt > K) {
            cut = kth_by_bisection(gh, inv_cell, q, H2, lane, K, n_x);
            num = make_float3(0.f, 0.f, 0.f);
            den = 0.f;
            warp_for_each_neighbor(gh, inv_cell, q, H2, lane, gather);
        }
        if (lane == 0) kthV[v] = cut;
    }
    num.x = group_sum(num.x); num.y = group_sum(num.y); num.z = group_sum(num.z); den = group_sum(den);
    if (lane != 0) return;
    const float dc = fmaxf(den, eps);
    // out_div = scale_factor when the caller wants render units (pipe_fluid.py:45: raw_render_xyz / gm.scale_factor)
    vis_out[3 * v] = (q.x + num.x * secs / dc) / out_div;
    vis_out[3 * v + 1] = (q.y + num.y * secs / dc) / out_div;
    vis_out[3 * v + 2] = (q.z + num.z * secs / dc) / out_div;
    if (num_out) { num_out[3 * v] = num.x; num_out[3 * v + 1] = num.y; num_out[3 * v + 2] = num.z; }
    if (den_out) den_out[v] = den;
}

// dL/dX_j = sum_{v: j in N(v), j <= kthV[v]} [ dw/dX_j * (secs/dc_v) * (G_v.u_j - [den_v > eps] G_v.num_v/den_v)
//                                              + w * (secs/dc_v) * G_v / secs ]
// gathered per hidden particle over the grid of the (un-advected) visual particles.
// pre-pass: per visual particle v (stored at sorted position a of the visual grid)
//   aux0[a] = { g_scale * (G_v [+ G2_v]),  c1 = [den_v > eps] (Gs_v . num_v) / den_v }
//   aux1[a] = { secs / dc_v, 1 / dc_v, kthV_v, - },  dc_v = max(den_v, eps)
__global__ void advect_bwd_pack_kernel(GridView gv, int V, const int *__restrict__ kthV, const float *__restrict__ num,
                                       const float *__restrict__ den, const float *__restrict__ G, const float *__restrict__ G2,
                                       float g_scale, float secs, float eps) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= V) return;
    const uint32_t v = gv.sorted_idx[a];
    float3 Gv = make_float3(G[3 * v], G[3 * v + 1], G[3 * v + 2]);
    if (G2) { Gv.x += G2[3 * v]; Gv.y += G2[3 * v + 1]; Gv.z += G2[3 * v + 2]; }
    Gv.x *= g_scale; Gv.y *= g_scale; Gv.z *= g_scale;
    const float dn = den[v];
    const float dc = fmaxf(dn, eps);
    const float c1 = dn > eps ? (Gv.x * num[3 * v] + Gv.y * num[3 * v + 1] + Gv.z * num[3 * v + 2]) / dn : 0.f;
    gv.aux0[a] = make_float4(Gv.x, Gv.y, Gv.z, c1);
    gv.aux1[a] = make_float4(secs / dc, 1.0f / dc, __int_as_float(kthV[v]), 0.f);
}
__global__ void __launch_bounds__(128)
advect_bwd_kernel(GridView gv, float inv_cell, const float *__restrict__ X, const float *__restrict__ xyz, int N, float H2,
                  float term1, float secs, float *__restrict__ dL_dX, int accumulate) {
    const int j = blockIdx.x * QPB + (threadIdx.x / GROUP);
    const int lane = threadIdx.x % GROUP;
    if (j >= N) return;
    const float3 xj = make_float3(X[3 * j], X[3 * j + 1], X[3 * j + 2]);
    const float3 u = make_float3((xj.x - xyz[3 * j]) / secs, (xj.y - xyz[3 * j + 1]) / secs, (xj.z - xyz[3 * j + 2]) / secs);
    float3 acc = make_float3(0.f, 0.f, 0.f);
    warp_for_each_neighbor(gv, inv_cell, xj, H2, lane, [&](int, const float4 &pv, float d2, uint32_t a) {
        const float4 p1 = gv.aux1[a];
        if (j > __float_as_int(p1.z)) return;
        const float4 p0 = gv.aux0[a];
        const float w = poly6(d2, H2, term1);
        const float dw = dpoly6_dd2(d2, H2, term1);
        const float coef = p0.x * u.x + p0.y * u.y + p0.z * u.z - p0.w;
        // d(d2)/dX_j = 2 (X_j - vis_v)
        const float s = dw * 2.f * coef * p1.x;
        const float t = w * p1.y;  // w * (secs/dc) * (1/secs)
        acc.x += s * (xj.x - pv.x) + t * p0.x;
        acc.y += s * (xj.y - pv.y) + t * p0.y;
        acc.z += s * (xj.z - pv.z) + t * p0.z;
    });
    acc.x = group_sum(acc.x); acc.y = group_sum(acc.y); acc.z = group_sum(acc.z);
    if (lane < 3) {
        const float val = lane == 0 ? acc.x : (lane == 1 ? acc.y : acc.z);
        if (accumulate) dL_dX[3 * j + lane] += val; else dL_dX[3 * j + lane] = val;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// No-grad PBF solver tick: guess_hidden_particles gm_fluid.py:809-844,
// project_gas_constraints :896-1021, confirm_guess_hidden_particles :1160-1175, update_visual_particles :1197-1239.
// The reference builds an edge list with radius_graph(exyz, H, loop=True, K) and runs ~40 gather / index_add_ kernels
// over it per solver iteration; here an iteration is a grid build, a neighbour count and two gather passes.
// Edge convention: row = neighbour j, col = query c, edge exists iff |x_j - x_c|^2 < H^2 and
// j <= kth[c]; every index_add_(0, row, .) therefore is "for node r: sum over c in N(r) with r <= kth[c]".
// ---------------------------------------------------------------------------------------------------------------
constexpr float PBF_EPS = 1e-8f;  // self.EPSILON, gm_fluid.py:98

// spiky_grad(diff, rlen) of gm_fluid.py:171-177 for diff = x_r - x_c, rlen = sqrt(d2 + EPS)
__device__ __forceinline__ float3 spiky_grad(float3 diff, float d2, float H, float sterm) {
    const float rlen = sqrtf(d2 + PBF_EPS);
    if (!(rlen < H && rlen > 0.f)) return make_float3(0.f, 0.f, 0.f);
    const float inv = 1.0f / (rlen + PBF_EPS);
    const float t = (H - rlen) * (H - rlen) * sterm;
    return make_float3(-(diff.x * inv) * t, -(diff.y * inv) * t, -(diff.z * inv) * t);
}

// pass 1: density, neighbour count, gradient sums -> lambda; pressure force correction
__global__ void __launch_bounds__(128)
solver_lambda_kernel(GridView g, float inv_cell, const float *__restrict__ X, int N, const float *__restrict__ imass,
                     const int *__restrict__ kth, float H, float term1, float sterm, float p0, float relaxation, float k_force,
                     const float *__restrict__ velocity, float *__restrict__ force, float *__restrict__ lambda_out,
                     float *__restrict__ nlen_out, float *__restrict__ pratio_out) {
    const int r = blockIdx.x * QPB + (threadIdx.x / GROUP);
    const int lane = threadIdx.x % GROUP;
    if (r >= N) return;
    const float3 q = make_float3(X[3 * r], X[3 * r + 1], X[3 * r + 2]);
    float pi = 0.f, grad_dot = 0.f;
    float3 gr = make_float3(0.f, 0.f, 0.f);
    int cnt = 0;
    warp_for_each_neighbor(g, inv_cell, q, H * H, lane, [&](int c, const float4 &pc, float d2, uint32_t) {
        if (r > kth[c]) return;
        cnt++;
        pi += poly6(d2, H * H, term1);
        if (c == r) return;  // self loop: density and count only
        const float3 sg = spiky_grad(make_float3(q.x - pc.x, q.y - pc.y, q.z - pc.z), d2, H, sterm);
        gr.x += sg.x; gr.y += sg.y; gr.z += sg.z;
        const float ax = sg.x / p0, ay = sg.y / p0, az = sg.z / p0;
        grad_dot += ax * ax + ay * ay + az * az;
    });
    pi = group_sum(pi); grad_dot = group_sum(grad_dot); cnt = group_sum(cnt);
    gr.x = group_sum(gr.x); gr.y = group_sum(gr.y); gr.z = group_sum(gr.z);
    if (lane != 0) return;
    gr.x /= p0; gr.y /= p0; gr.z /= p0;
    const float gr_dot = gr.x * gr.x + gr.y * gr.y + gr.z * gr.z;
    const float p_ratio = pi / imass[r] / p0;
    if (force != nullptr) {
        const float f = (1.0f - p_ratio) * -k_force;
        force[3 * r] += velocity[3 * r] * f; force[3 * r + 1] += velocity[3 * r + 1] * f; force[3 * r + 2] += velocity[3 * r + 2] * f;
    }
    lambda_out[r] = -(p_ratio - 1.0f) / ((grad_dot + gr_dot) + relaxation);
    nlen_out[r] = (float)cnt;
    if (pratio_out) pratio_out[r] = p_ratio;
}

// pass 2: position correction  exyz += sum_c (lambda_r + lambda_c + s_corr) spiky / p0 / (neighbours + counts)
__global__ void __launch_bounds__(128)
solver_delta_kernel(GridView g, float inv_cell, float *__restrict__ X, int N, const int *__restrict__ kth,
                    const float *__restrict__ lambda, const float *__restrict__ nlen, const float *__restrict__ counts, float H,
                    float term1, float sterm, float p0, float K_P, int E_P, float corr_denom) {
    const int r = blockIdx.x * QPB + (threadIdx.x / GROUP);
    const int lane = threadIdx.x % GROUP;
    if (r >= N) return;
    const float3 q = make_float3(X[3 * r], X[3 * r + 1], X[3 * r + 2]);
    const float lam_r = lambda[r];
    float3 d = make_float3(0.f, 0.f, 0.f);
    warp_for_each_neighbor(g, inv_cell, q, H * H, lane, [&](int c, const float4 &pc, float d2, uint32_t) {
        if (c == r || r > kth[c]) return;
        const float3 sg = spiky_grad(make_float3(q.x - pc.x, q.y - pc.y, q.z - pc.z), d2, H, sterm);
        float base = poly6(d2, H * H, term1) / corr_denom, pw = 1.0f;
        for (int e = 0; e < E_P; e++) pw *= base;   // (poly6 / denom) ** E_P, integer exponent (gm_fluid.py:102)
        const float w = (lam_r + lambda[c]) + -K_P * pw;
        d.x += w * sg.x; d.y += w * sg.y; d.z += w * sg.z;
    });
    d.x = group_sum(d.x); d.y = group_sum(d.y); d.z = group_sum(d.z);
    if (lane >= 3) return;
    const float den = nlen[r] + counts[r];
    const float v = (lane == 0 ? d.x : (lane == 1 ? d.y : d.z)) / p0 / den;
    // neighbours read the grid's snapshot of the positions (sorted_pos), so the in-place update is race free
    X[3 * r + lane] = (lane == 0 ? q.x : (lane == 1 ? q.y : q.z)) + v;
}

// bincount(row) of radius_graph(X, r, loop, K): for node r the number of queries c that list r among their neighbours
__global__ void __launch_bounds__(128)
graph_degree_kernel(GridView g, float inv_cell, const float *__restrict__ X, int N, const int *__restrict__ kth, float r2, int loop,
                    int *__restrict__ degree) {
    const int r = blockIdx.x * QPB + (threadIdx.x / GROUP);
    const int lane = threadIdx.x % GROUP;
    if (r >= N) return;
    const float3 q = make_float3(X[3 * r], X[3 * r + 1], X[3 * r + 2]);
    int cnt = 0;
    warp_for_each_neighbor(g, inv_cell, q, r2, lane, [&](int c, const float4 &, float, uint32_t) {
        if (r > kth[c] || (!loop && c == r)) return;
        cnt++;
    });
    cnt = group_sum(cnt);
    if (lane == 0) degree[r] = cnt;
}

// Rigid coupling (gm_fluid.py:1023-1105, 1241-1289): a particle inside the rigid body is moved onto the nearest rigid-body
// sample among its radius() neighbours (x += -(x - nearest)); ties go to the smaller sample index, as scatter_min's argmin
// over index-ordered edges does.  body: 0 cuboid (|x - c| <= half edge per axis), 1 sphere (|x - c| <= radius),
// 2 cylinder (axis z: (x-cx)^2 + (y-cy)^2 <= radius^2, |z - cz| <= half height) -- check_inside_rigid_body, :1024-1056.
struct RigidBody {
    int kind;
    float3 center, prm;
};
__device__ __forceinline__ bool inside_rigid(const RigidBody &b, float3 p) {
    if (b.kind == 0)
        return p.x >= b.center.x - b.prm.x && p.x <= b.center.x + b.prm.x && p.y >= b.center.y - b.prm.y && p.y <= b.center.y + b.prm.y &&
               p.z >= b.center.z - b.prm.z && p.z <= b.center.z + b.prm.z;
    const float dx = p.x - b.center.x, dy = p.y - b.center.y, dz = p.z - b.center.z;
    if (b.kind == 1) return sqrtf(dx * dx + dy * dy + dz * dz) <= b.prm.x;
    return dx * dx + dy * dy <= b.prm.x * b.prm.x && p.z >= b.center.z - b.prm.y && p.z <= b.center.z + b.prm.y;
}
__global__ void __launch_bounds__(128)
rigid_project_kernel(GridView g, float inv_cell, float *__restrict__ xyz, int N, RigidBody body, float r2, int K, int M,
                     int *__restrict__ n_inside) {
    const int i = blockIdx.x * QPB + (threadIdx.x / GROUP);
    const int lane = threadIdx.x % GROUP;
    if (i >= N) return;
    const float3 q = make_float3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (!inside_rigid(body, q)) return;   // uniform over the query group
    int cut = 0x7fffffff;
    if (K > 0) {
        int cnt = 0;
        warp_for_each_neighbor(g, inv_cell, q, r2, lane, [&](int, const float4 &, float, uint32_t) { cnt++; });
        if (group_sum(cnt) > K) cut = kth_by_bisection(g, inv_cell, q, r2, lane, K, M);
    }
    float best = 3.4e38f;
    int best_j = 0x7fffffff;
    float3 bp = q;
    warp_for_each_neighbor(g, inv_cell, q, r2, lane, [&](int j, const float4 &pj, float d2, uint32_t) {
        if (j > cut) return;
        if (d2 < best || (d2 == best && j < best_j)) { best = d2; best_j = j; bp = make_float3(pj.x, pj.y, pj.z); }
    });
    const unsigned m = group_mask();
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(m, best, o);
        const int oj = __shfl_xor_sync(m, best_j, o);
        const float ox = __shfl_xor_sync(m, bp.x, o), oy = __shfl_xor_sync(m, bp.y, o), oz = __shfl_xor_sync(m, bp.z, o);
        if (ob < best || (ob == best && oj < best_j)) { best = ob; best_j = oj; bp = make_float3(ox, oy, oz); }
    }
    if (lane == 0) {
        if (n_inside) atomicAdd(n_inside, 1);
        if (best_j != 0x7fffffff) {
            // x += -(x - nearest), evaluated like the reference (the result is `nearest` up to one rounding)
            xyz[3 * i] = q.x + -(q.x - bp.x); xyz[3 * i + 1] = q.y + -(q.y - bp.y); xyz[3 * i + 2] = q.z + -(q.z - bp.z);
        }
    }
}

// guess_hidden_particles (gm_fluid.py:809-844)
__global__ void solver_guess_kernel(int N, const float *__restrict__ xyz, float *__restrict__ velocity, float *__restrict__ buoyancy,
                                    float *__restrict__ force, float *__restrict__ estimate_xyz, float *__restrict__ counts,
                                    float3 gravity_alpha, float secs, float buoyancy_max_y, float scale_factor, float decay,
                                    int use_wind, float3 wind_force, float wind_power, float wind_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float y = xyz[3 * i + 1];
    const float coeff = buoyancy_max_y > 0.f ? 1.0f - (y / (buoyancy_max_y * scale_factor)) : 1.0f;
    const float b[3] = {gravity_alpha.x, gravity_alpha.y, gravity_alpha.z};   // ones_like(buoyancy) * gravity * alpha
    const float wf[3] = {wind_force.x, wind_force.y, wind_force.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float v = velocity[3 * i + k] + (b[k] * coeff) * secs + secs * force[3 * i + k];
        if (use_wind) {
            const float w = fminf(fmaxf(powf(y / scale_factor, wind_power) * wf[k], 0.0f), wind_max);
            v += w * secs;
        }
        velocity[3 * i + k] = v;
        buoyancy[3 * i + k] = decay > 0.f ? b[k] * decay : b[k];
        force[3 * i + k] = 0.f;
        estimate_xyz[3 * i + k] = xyz[3 * i + k] + secs * v;
    }
    counts[i] = 0.f;
}

// confirm_guess_hidden_particles (gm_fluid.py:1160-1175)
__global__ void solver_confirm_kernel(int N, float *__restrict__ xyz, const float *__restrict__ estimate_xyz,
                                      float *__restrict__ velocity, float secs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float dx = estimate_xyz[3 * i] - xyz[3 * i], dy = estimate_xyz[3 * i + 1] - xyz[3 * i + 1], dz = estimate_xyz[3 * i + 2] - xyz[3 * i + 2];
    const bool still = sqrtf(dx * dx + dy * dy + dz * dz) < PBF_EPS;
    velocity[3 * i] = still ? 0.f : dx / secs; velocity[3 * i + 1] = still ? 0.f : dy / secs; velocity[3 * i + 2] = still ? 0.f : dz / secs;
    if (!still) { xyz[3 * i] = estimate_xyz[3 * i]; xyz[3 * i + 1] = estimate_xyz[3 * i + 1]; xyz[3 * i + 2] = estimate_xyz[3 * i + 2]; }
}

// payload of update_visual_particles: the hidden particles' velocities in grid order
__global__ void advect_pack_vel_kernel(GridView gh, int N, const float *__restrict__ velocity) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= N) return;
    const uint32_t j = gh.sorted_idx[a];
    gh.aux0[a] = make_float4(velocity[3 * j], velocity[3 * j + 1], velocity[3 * j + 2], 0.f);
}

// ---------------------------------------------------------------------------------------------------------------
// P5: pair distance loss  L = sum_{i != j, d_ij < thr} (thr - d_ij)^2  and dL/dp  (loss_utils.py:98-121)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pair_distance_kernel(GridView g, float inv_cell, const float *__restrict__ pts, int n, float thr, float grad_scale,
                     float *__restrict__ loss_out, float *__restrict__ dL_dp) {
    // 8 lanes per point (the grid is sparse at cell = threshold: 25 mostly empty rows per query are pure latency for
    // one thread); out-of-range groups keep running with no work so that the warp-wide loss reduction stays convergent
    const int i = blockIdx.x * QPB + (threadIdx.x / GROUP);
    const int lane = threadIdx.x % GROUP;
    float loss = 0.f;
    float3 acc = make_float3(0.f, 0.f, 0.f);
    if (i < n) {
        const float3 q = make_float3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        warp_for_each_neighbor(g, inv_cell, q, thr * thr, lane, [&](int j, const float4 &pj, float d2, uint32_t) {
            if (j == i) return;
            const float d = sqrtf(d2);
            if (!(d < thr)) return;
            const float m = thr - d;
            loss += m * m;
            if (d > 0.f) {
                // both (i,j) and (j,i) terms of the full-matrix sum depend on p_i
                const float s = -4.f * m / d;
                acc.x += s * (q.x - pj.x); acc.y += s * (q.y - pj.y); acc.z += s * (q.z - pj.z);
            }
        });
    }
    acc.x = group_sum(acc.x); acc.y = group_sum(acc.y); acc.z = group_sum(acc.z);
    if (i < n && dL_dp && lane < 3) dL_dp[3 * i + lane] = grad_scale * (lane == 0 ? acc.x : (lane == 1 ? acc.y : acc.z));
    loss = warp_sum(loss);
    if ((threadIdx.x & 31) == 0 && loss != 0.f) atomicAdd(loss_out, loss);
}

// ---------------------------------------------------------------------------------------------------------------
// distCUDA2: mean squared distance to the 3 nearest OTHER points (simple_knn.cu:134-166).  Expanding ring search on
// the grid: rings of cells are visited until the 3rd best distance is provably final.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void update3(float d, float *best) {
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (best[k] > d) { const float t = best[k]; best[k] = d; d = t; }
}
__global__ void __launch_bounds__(128)
knn3_kernel(GridView g, float cell, const float *__restrict__ pts, int n, int max_ring, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float inv_cell = sub_inv(1.0f / cell);   // the table's sub-cells (grid_build)
    cell /= (float)SUB;
    const float3 q = make_float3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    const int3 c = cell_of(q.x, q.y, q.z, inv_cell);
    float best[3] = {3.4e38f, 3.4e38f, 3.4e38f};
    bool resolved = false;
    for (int ring = 0; ring <= max_ring; ring++) {
        for (int dz = -ring; dz <= ring; dz++)
            for (int dy = -ring; dy <= ring; dy++)
                for (int dx = -ring; dx <= ring; dx++) {
                    if (max(abs(dx), max(abs(dy), abs(dz))) != ring) continue;  // shell only
                    const int3 cc = make_int3(c.x + dx, c.y + dy, c.z + dz);
                    const uint32_t b = hash_cell(cc, g.M);
                    for (uint32_t a = g.bucket_start[b]; a < g.bucket_start[b + 1]; a++) {
                        const float4 p = g.sorted_pos[a];
                        const int3 pc = cell_of(p.x, p.y, p.z, inv_cell);
                        if (pc.x != cc.x || pc.y != cc.y || pc.z != cc.z) continue;
                        if ((int)__float_as_uint(p.w) == i) continue;
                        const float ex = p.x - q.x, ey = p.y - q.y, ez = p.z - q.z;
                        update3(ex * ex + ey * ey + ez * ez, best);
                    }
                }
        // every unvisited point is farther than ring*cell (it lies outside the visited cube of half-width ring*cell
        // around q's cell, hence at least `ring` whole cells away from q along some axis)
        const float safe = ring * cell;
        if (best[2] <= safe * safe) { resolved = true; break; }
    }
    if (!resolved) {  // isolated point (or fewer than 4 points): exhaustive scan
        best[0] = best[1] = best[2] = 3.4028235e38f;  // FLT_MAX like the reference's initial value
        for (int a = 0; a < n; a++) {
            const float4 p = g.sorted_pos[a];
            if ((int)__float_as_uint(p.w) == i) continue;
            const float ex = p.x - q.x, ey = p.y - q.y, ez = p.z - q.z;
            update3(ex * ex + ey * ey + ez * ez, best);
        }
    }
    out[i] = (best[0] + best[1] + best[2]) / 3.0f;
}

// ---------------------------------------------------------------------------------------------------------------
// P3 map: Y(e) = 100 e + secs * ( (100 e - xyz)/secs + b * secs + secs * F ),  b = buoy * (1 - e_y / bmax) or buoy
// (gm_fluid.py:846-862) and its transpose-Jacobian product.
// ---------------------------------------------------------------------------------------------------------------
__global__ void next_tick_fwd_kernel(int N, const float *__restrict__ e, const float *__restrict__ xyz,
                                     const float *__restrict__ buoy, const float *__restrict__ force, float secs,
                                     float bmax, float scale, float *__restrict__ X, float *__restrict__ Y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float coeff = bmax > 0.f ? 1.0f - e[3 * i + 1] / bmax : 1.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float x = e[3 * i + k] * scale;
        const float tmp_v = (x - xyz[3 * i + k]) / secs;
        const float ev = tmp_v + buoy[3 * i + k] * coeff * secs + secs * force[3 * i + k];
        if (X) X[3 * i + k] = x;
        if (Y) Y[3 * i + k] = x + secs * ev;
    }
}
// dL/de = scale*dL/dX + (dY/de)^T dL/dY + 2*lambda_exyz*scale*(scale*e - estimate_xyz)/(3N)
// With adam_p != NULL the Adam update of torch.optim.Adam (see adam_dev_kernel; bc = {1 - beta1^t, sqrt(1 - beta2^t)} from
// adam_prepare_kernel, bc[0] == 0: gated off) is applied to the same element right away: `e` and `adam_p` are the same tensor,
// every thread reads its own three components before it writes them.
struct AdamFused {
    float *p, *m, *v;
    const float *bc;
    float lr, beta1, beta2, eps;
};
__global__ void combine_grad_kernel(int N, const float *__restrict__ e, const float *__restrict__ buoy, float secs, float bmax,
                                    float scale, const float *__restrict__ dL_dX, const float *__restrict__ dL_dY,
                                    const float *__restrict__ estimate_xyz, float w_exyz, float *__restrict__ dL_de,
                                    float *__restrict__ exyz_loss, AdamFused adam) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (i < N) {
        const float e3[3] = {e[3 * i], e[3 * i + 1], e[3 * i + 2]};
        float gy[3] = {0.f, 0.f, 0.f};
        float cross = 0.f;
        if (dL_dY) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                gy[k] = dL_dY[3 * i + k];
                if (bmax > 0.f) cross += gy[k] * (secs * secs * buoy[3 * i + k] * (-1.0f / bmax));
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float gsum = (dL_dX ? scale * dL_dX[3 * i + k] : 0.f) + 2.0f * scale * gy[k];
            if (k == 1) gsum += cross;
            if (estimate_xyz) {
                const float d = e3[k] * scale - estimate_xyz[3 * i + k];
                l += d * d;
                gsum += w_exyz * 2.0f * scale * d / (3.0f * N);
            }
            dL_de[3 * i + k] = gsum;
            if (adam.p != nullptr) {
                const float bc1 = adam.bc[0], bc2_sqrt = adam.bc[1];
                if (bc1 != 0.f) {
                    const int q = 3 * i + k;
                    const float mi = adam.m[q] + (gsum - adam.m[q]) * (1.0f - adam.beta1);
                    const float vi = adam.beta2 * adam.v[q] + (1.0f - adam.beta2) * gsum * gsum;
                    adam.m[q] = mi;
                    adam.v[q] = vi;
                    const float denom = sqrtf(vi) / bc2_sqrt + adam.eps;
                    adam.p[q] = e3[k] - (adam.lr / bc1) * (mi / denom);
                }
            }
        }
    }
    l = warp_sum(l);
    if (exyz_loss && (threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(exyz_loss, l / (3.0f * N));
}

// mean((p_ratio - 1)^2) and its gradient w.r.t. p_ratio scaled by `weight`
__global__ void ratio_loss_kernel(int N, const float *__restrict__ p_ratio, float weight, float *__restrict__ loss,
                                  float *__restrict__ dL_dpratio) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (i < N) {
        const float d = p_ratio[i] - 1.0f;
        l = d * d;
        dL_dpratio[i] = weight * 2.0f * d / N;
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(loss, l / N);
}

// torch.optim.Adam step (no amsgrad / weight decay), grad pre-scaled by grad_scale (= 1/batch, gm_fluid.py:428-430)
__global__ void adam_kernel(long long n, float *__restrict__ p, const float *__restrict__ grad, float *__restrict__ m,
                            float *__restrict__ v, float grad_scale, float lr, float beta1, float beta2, float eps,
                            float bc1, float bc2_sqrt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gI = grad[i] * grad_scale;
    const float mi = m[i] + (gI - m[i]) * (1.0f - beta1);           // lerp form used by torch
    const float vi = beta2 * v[i] + (1.0f - beta2) * gI * gI;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
}

// device-side step counter variant (CUDA-graph replay safe): bumps *step and derives the bias corrections in double
__global__ void adam_prepare_kernel(int *step, float beta1, float beta2, float *bc /*[2]*/, const int *__restrict__ skip_flag) {
    if (skip_flag != nullptr && *skip_flag != 0) {  // the iteration that produced `grad` is void (e.g. the rasterizer's instance
        bc[0] = 0.f;                                // capacity overflowed and it rendered nothing): leave parameters, moments and
        return;                                     // the step count alone.  bc[0] == 0 tells adam_dev_kernel (1 - beta1^t > 0 otherwise)
    }
    const int t = ++(*step);
    bc[0] = (float)(1.0 - pow((double)beta1, (double)t));
    bc[1] = (float)sqrt(1.0 - pow((double)beta2, (double)t));
}
__global__ void adam_dev_kernel(long long n, float *__restrict__ p, const float *__restrict__ grad, float *__restrict__ m,
                                float *__restrict__ v, float grad_scale, float lr, float beta1, float beta2, float eps,
                                const float *__restrict__ bc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float bc1 = bc[0], bc2_sqrt = bc[1];
    if (bc1 == 0.f) return;  // gated off by adam_prepare_kernel
    const float gI = grad[i] * grad_scale;
    const float mi = m[i] + (gI - m[i]) * (1.0f - beta1);
    const float vi = beta2 * v[i] + (1.0f - beta2) * gI * gI;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
}

// scatter_min over int64 index (torch_scatter 2.1.2): out[idx] = min, arg = position of the min (ties: smallest position)
__global__ void scatter_min_init_kernel(int n_out, float *out, long long *arg, long long n_src) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) { out[i] = __int_as_float(0x7f800000); arg[i] = n_src; }
}
__device__ __forceinline__ void atomic_min_float(float *addr, float val) {
    // ordered-int trick, valid for all non-NaN floats
    if (val >= 0.f) atomicMin((int *)addr, __float_as_int(val));
    else atomicMax((unsigned int *)addr, __float_as_uint(val));
}
__global__ void scatter_min_val_kernel(long long n, const float *__restrict__ src, const long long *__restrict__ idx, float *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomic_min_float(&out[idx[i]], src[i]);
}
__global__ void scatter_min_arg_kernel(long long n, const float *__restrict__ src, const long long *__restrict__ idx,
                                       const float *__restrict__ out, long long *arg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && src[i] == out[idx[i]]) atomicMin((unsigned long long *)&arg[idx[i]], (unsigned long long)i);
}
__global__ void scatter_min_fix_kernel(int n_out, float *out, const long long *arg, long long n_src) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out && arg[i] == n_src) out[i] = 0.f;  // empty groups read 0 like torch_scatter
}

}  // namespace fnx

using namespace fnx;

extern "C" {

size_t fnx_grid_bytes(int32_t n) { return grid_bytes(n); }

int fnx_grid_build(const float *pts, int32_t n, float cell, void *grid, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(n >= 0 && cell > 0.f && grid && (pts || n == 0), "bad arguments");
    return grid_build(pts, n, cell, grid, (cudaStream_t)stream);
}

int fnx_radius_count(const void *grid_x, int32_t nx, float cell, const float *y, int32_t ny, float r, int32_t max_num_neighbors,
                     int32_t *counts, int32_t *kth, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid_x && nx >= 0 && ny >= 0 && r > 0.f && r <= cell * 1.000001f && max_num_neighbors > 0, "bad arguments (need r <= cell)");
    if (ny == 0) return FNX_OK;
    GridView g = grid_view((void *)grid_x, nx);
    radius_count_kernel<<<(ny + QPB - 1) / QPB, 128, 0, (cudaStream_t)stream>>>(g, 1.0f / cell, y, ny, r * r, max_num_neighbors, nx, counts, kth);
    FNX_LAUNCH_CHECK("radius_count_kernel");
    return FNX_OK;
}

int fnx_radius_fill(const void *grid_x, int32_t nx, float cell, const float *y, int32_t ny, float r, const int32_t *kth,
                    const int64_t *offsets, int64_t *edge_query, int64_t *edge_x, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid_x && kth && offsets && edge_query && edge_x && r <= cell * 1.000001f, "bad arguments");
    if (ny == 0) return FNX_OK;
    GridView g = grid_view((void *)grid_x, nx);
    radius_fill_kernel<<<(ny + 127) / 128, 128, 0, (cudaStream_t)stream>>>(g, 1.0f / cell, y, ny, r * r, kth, (const long long *)offsets,
                                                                         (long long *)edge_query, (long long *)edge_x);
    FNX_LAUNCH_CHECK("radius_fill_kernel");
    return FNX_OK;
}

int fnx_pbf_density_fwd(const void *grid, const float *X, int32_t N, const float *imass, const int32_t *kth, float H, float p0,
                        float *p_ratio, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && X && imass && kth && p_ratio && N >= 0 && H > 0.f && p0 > 0.f, "bad arguments");
    if (N == 0) return FNX_OK;
    GridView g = grid_view((void *)grid, N);
    const float term1 = (float)(315.0 / (64.0 * 3.14159265358979323846 * pow((double)H, 9)));
    density_fwd_kernel<<<(N + QPB_SIDE - 1) / QPB_SIDE, 128, 0, (cudaStream_t)stream>>>(g, 1.0f / H, X, N, imass, kth, H * H, term1, p0, p_ratio);
    FNX_LAUNCH_CHECK("density_fwd_kernel");
    return FNX_OK;
}

int fnx_pbf_density_fwd_counted(const void *grid, const float *X, int32_t N, const float *imass, int32_t max_num_neighbors, float H,
                                float p0, int32_t *kth_out, float *p_ratio, int32_t *cap_flag, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && X && imass && kth_out && p_ratio && cap_flag && N >= 0 && H > 0.f && p0 > 0.f && max_num_neighbors > 0, "bad arguments");
    if (N == 0) return FNX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    GridView g = grid_view((void *)grid, N);
    const float term1 = (float)(315.0 / (64.0 * 3.14159265358979323846 * pow((double)H, 9)));
    FNX_CUDA_TRY(cudaMemsetAsync(cap_flag, 0, sizeof(int32_t), st));
    density_fwd_counted_kernel<<<(N + QPB_SIDE - 1) / QPB_SIDE, 128, 0, st>>>(g, 1.0f / H, X, N, imass, max_num_neighbors, H * H, term1, p0, kth_out,
                                                                    p_ratio, cap_flag);
    FNX_LAUNCH_CHECK("density_fwd_counted_kernel");
    density_fwd_if_capped_kernel<<<min((N + QPB_SIDE - 1) / QPB_SIDE, 148 * 8), 128, 0, st>>>(g, 1.0f / H, X, N, imass, kth_out, H * H, term1, p0, p_ratio, cap_flag);
    FNX_LAUNCH_CHECK("density_fwd_if_capped_kernel");
    return FNX_OK;
}

int fnx_pbf_density_bwd(const void *grid, const float *X, int32_t N, const float *imass, const int32_t *kth, float H, float p0,
                        const float *dL_dpratio, float *dL_dX, int32_t accumulate, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && X && imass && kth && dL_dpratio && dL_dX && N >= 0, "bad arguments");
    if (N == 0) return FNX_OK;
    GridView g = grid_view((void *)grid, N);
    const float term1 = (float)(315.0 / (64.0 * 3.14159265358979323846 * pow((double)H, 9)));
    density_bwd_pack_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, N, imass, kth, p0, dL_dpratio);
    FNX_LAUNCH_CHECK("density_bwd_pack_kernel");
    density_bwd_kernel<<<(N + QPB_SIDE - 1) / QPB_SIDE, 128, 0, (cudaStream_t)stream>>>(g, 1.0f / H, X, N, imass, kth, H * H, term1, p0, dL_dpratio, dL_dX, accumulate);
    FNX_LAUNCH_CHECK("density_bwd_kernel");
    return FNX_OK;
}

int fnx_pbf_density_bwd_ratio(const void *grid, const float *X, int32_t N, const float *imass, const int32_t *kth, float H, float p0,
                              const float *p_ratio, float weight, float *loss, float *dL_dpratio, float *dL_dX, int32_t accumulate,
                              fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && X && imass && kth && p_ratio && loss && dL_dpratio && dL_dX && N >= 0, "bad arguments");
    FNX_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
    if (N == 0) return FNX_OK;
    GridView g = grid_view((void *)grid, N);
    const float term1 = (float)(315.0 / (64.0 * 3.14159265358979323846 * pow((double)H, 9)));
    density_bwd_pack_ratio_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, N, imass, kth, p0, p_ratio, weight, dL_dpratio, loss);
    FNX_LAUNCH_CHECK("density_bwd_pack_ratio_kernel");
    density_bwd_kernel<<<(N + QPB_SIDE - 1) / QPB_SIDE, 128, 0, (cudaStream_t)stream>>>(g, 1.0f / H, X, N, imass, kth, H * H, term1, p0, dL_dpratio, dL_dX, accumulate);
    FNX_LAUNCH_CHECK("density_bwd_kernel");
    return FNX_OK;
}

int fnx_visual_advect_fwd(const void *grid_hidden, const float *X, const float *xyz, int32_t N, const float *visual, int32_t V,
                          const int32_t *kthV, float H, float secs, float out_div, float *visual_out, float *num_out,
                          float *den_out, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid_hidden && X && xyz && kthV && visual_out && (visual || V == 0) && out_div != 0.f, "bad arguments");
    if (V == 0) return FNX_OK;
    GridView g = grid_view((void *)grid_hidden, N);
    const float term1 = (float)(315.0 / (64.0 * 3.14159265358979323846 * pow((double)H, 9)));
    if (N > 0) {
        advect_pack_u_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, N, xyz, secs);
        FNX_LAUNCH_CHECK("advect_pack_u_kernel");
    }
    advect_fwd_kernel<false><<<(V + QPB - 1) / QPB, 128, 0, (cudaStream_t)stream>>>(g, 1.0f / H, X, xyz, visual, V, (int *)kthV, 0, N, H * H, term1,
                                                                              secs, 1e-8f, out_div, visual_out, num_out, den_out);
    FNX_LAUNCH_CHECK("advect_fwd_kernel");
    return FNX_OK;
}

int fnx_visual_advect_fwd_counted(const void *grid_hidden, const float *X, const float *xyz, int32_t N, const float *visual, int32_t V,
                                  int32_t max_num_neighbors, float H, float secs, float out_div, float *visual_out, float *num_out,
                                  float *den_out, int32_t *kthV_out, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid_hidden && X && xyz && kthV_out && visual_out && (visual || V == 0) && out_div != 0.f && max_num_neighbors > 0,
                "bad arguments");
    if (V == 0) return FNX_OK;
    GridView g = grid_view((void *)grid_hidden, N);
    const float term1 = (float)(315.0 / (64.0 * 3.14159265358979323846 * pow((double)H, 9)));
    if (N > 0) {
        advect_pack_u_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, N, xyz, secs);
        FNX_LAUNCH_CHECK("advect_pack_u_kernel");
    }
    advect_fwd_kernel<true><<<(V + QPB - 1) / QPB, 128, 0, (cudaStream_t)stream>>>(g, 1.0f / H, X, xyz, visual, V, kthV_out, max_num_neighbors, N,
                                                                             H * H, term1, secs, 1e-8f, out_div, visual_out, num_out, den_out);
    FNX_LAUNCH_CHECK("advect_fwd_kernel");
    return FNX_OK;
}

int fnx_visual_advect_bwd(const void *grid_visual, const float *X, const float *xyz, int32_t N, int32_t V, const int32_t *kthV,
                          const float *num, const float *den, const float *dL_dvisual_out, const float *dL_dvisual_out2,
                          float g_scale, float H, float secs, float *dL_dX, int32_t accumulate, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid_visual && X && xyz && kthV && num && den && dL_dvisual_out && dL_dX, "bad arguments");
    if (N == 0) return FNX_OK;
    GridView g = grid_view((void *)grid_visual, V);
    const float term1 = (float)(315.0 / (64.0 * 3.14159265358979323846 * pow((double)H, 9)));
    if (V > 0) {
        advect_bwd_pack_kernel<<<(V + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, V, kthV, num, den, dL_dvisual_out, dL_dvisual_out2, g_scale,
                                                                             secs, 1e-8f);
        FNX_LAUNCH_CHECK("advect_bwd_pack_kernel");
    }
    advect_bwd_kernel<<<(N + QPB - 1) / QPB, 128, 0, (cudaStream_t)stream>>>(g, 1.0f / H, X, xyz, N, H * H, term1, secs, dL_dX, accumulate);
    FNX_LAUNCH_CHECK("advect_bwd_kernel");
    return FNX_OK;
}

int fnx_pbf_guess_hidden(int32_t N, const float *xyz, float *velocity, float *buoyancy, float *force, float *estimate_xyz, float *counts,
                         const float *gravity3_host, float alpha, float secs, float buoyancy_max_y, float scale_factor,
                         float buoyancy_decay_rate, int32_t use_wind, const float *wind_force3_host, float wind_power,
                         float wind_force_max, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(N >= 0 && xyz && velocity && buoyancy && force && estimate_xyz && counts && gravity3_host, "bad arguments");
    FNX_REQUIRE(!use_wind || wind_force3_host, "use_wind needs wind_force");
    if (N == 0) return FNX_OK;
    const float3 ga = make_float3(gravity3_host[0] * alpha, gravity3_host[1] * alpha, gravity3_host[2] * alpha);
    const float3 wf = use_wind ? make_float3(wind_force3_host[0], wind_force3_host[1], wind_force3_host[2]) : make_float3(0.f, 0.f, 0.f);
    solver_guess_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, xyz, velocity, buoyancy, force, estimate_xyz, counts, ga, secs,
                                                                         buoyancy_max_y, scale_factor, buoyancy_decay_rate, use_wind, wf,
                                                                         wind_power, wind_force_max);
    FNX_LAUNCH_CHECK("solver_guess_kernel");
    return FNX_OK;
}

int fnx_pbf_project_gas_constraints(void *grid, float *estimate_xyz, int32_t N, const float *imass, const float *velocity, float *force,
                                    const float *counts, float H, float p0, float k, int32_t max_num_neighbors, float relaxation,
                                    float K_P, int32_t E_P, float DQ_P, int32_t *kth_scratch, float *lambda_scratch,
                                    float *nlen_scratch, float *p_ratio_out, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && estimate_xyz && imass && counts && kth_scratch && lambda_scratch && nlen_scratch && N >= 0 && H > 0.f && p0 > 0.f &&
                    max_num_neighbors > 0 && E_P >= 0 && (force == nullptr || velocity != nullptr), "bad arguments");
    if (N == 0) return FNX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = grid_build(estimate_xyz, N, H, grid, st);
    if (rc) return rc;
    GridView g = grid_view(grid, N);
    radius_count_kernel<<<(N + QPB - 1) / QPB, 128, 0, st>>>(g, 1.0f / H, estimate_xyz, N, H * H, max_num_neighbors, N, nullptr, kth_scratch);
    FNX_LAUNCH_CHECK("radius_count_kernel");
    const double H6 = pow((double)H, 6), H9 = pow((double)H, 9), PI = 3.14159265358979323846;
    const float term1 = (float)(315.0 / (64.0 * PI * H9)), sterm = (float)(45.0 / (PI * H6));
    // lamb_corr_denom = poly6(DQ_P^2 H^2), gm_fluid.py:126
    const double t = (double)H * H - (double)DQ_P * DQ_P * H * H;
    const float corr_denom = (float)((315.0 / (64.0 * PI * H9)) * t * t * t);
    solver_lambda_kernel<<<(N + QPB - 1) / QPB, 128, 0, st>>>(g, 1.0f / H, estimate_xyz, N, imass, kth_scratch, H, term1, sterm, p0, relaxation, k,
                                                              velocity, force, lambda_scratch, nlen_scratch, p_ratio_out);
    FNX_LAUNCH_CHECK("solver_lambda_kernel");
    solver_delta_kernel<<<(N + QPB - 1) / QPB, 128, 0, st>>>(g, 1.0f / H, estimate_xyz, N, kth_scratch, lambda_scratch, nlen_scratch, counts, H,
                                                             term1, sterm, p0, K_P, E_P, corr_denom);
    FNX_LAUNCH_CHECK("solver_delta_kernel");
    return FNX_OK;
}

int fnx_radius_graph_degree(void *grid, const float *X, int32_t N, float r, int32_t loop, int32_t max_num_neighbors, int32_t *kth_scratch,
                            int32_t *degree, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && X && kth_scratch && degree && N >= 0 && r > 0.f && max_num_neighbors > 0, "bad arguments");
    if (N == 0) return FNX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = grid_build(X, N, r, grid, st);
    if (rc) return rc;
    GridView g = grid_view(grid, N);
    // radius_graph(x, r, loop, K) = radius(x, x, r, K if loop else K + 1) minus the self pairs (torch_cluster 1.6.3)
    radius_count_kernel<<<(N + QPB - 1) / QPB, 128, 0, st>>>(g, 1.0f / r, X, N, r * r, loop ? max_num_neighbors : max_num_neighbors + 1, N, nullptr,
                                                             kth_scratch);
    FNX_LAUNCH_CHECK("radius_count_kernel");
    graph_degree_kernel<<<(N + QPB - 1) / QPB, 128, 0, st>>>(g, 1.0f / r, X, N, kth_scratch, r * r, loop, degree);
    FNX_LAUNCH_CHECK("graph_degree_kernel");
    return FNX_OK;
}

int fnx_rigid_project(void *grid, const float *rigid_xyz, int32_t M, float *xyz, int32_t N, int32_t body, const float *center3_host,
                      const float *params3_host, float r, int32_t max_num_neighbors, int32_t *n_inside, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && (rigid_xyz || M == 0) && (xyz || N == 0) && M >= 0 && N >= 0 && body >= 0 && body <= 2 && center3_host && params3_host &&
                    r > 0.f, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_inside) FNX_CUDA_TRY(cudaMemsetAsync(n_inside, 0, sizeof(int32_t), st));
    if (N == 0 || M == 0) return FNX_OK;
    int rc = grid_build(rigid_xyz, M, r, grid, st);
    if (rc) return rc;
    GridView g = grid_view(grid, M);
    RigidBody b;
    b.kind = body;
    b.center = make_float3(center3_host[0], center3_host[1], center3_host[2]);
    b.prm = make_float3(params3_host[0], params3_host[1], params3_host[2]);
    rigid_project_kernel<<<(N + QPB - 1) / QPB, 128, 0, st>>>(g, 1.0f / r, xyz, N, b, r * r, max_num_neighbors, M, n_inside);
    FNX_LAUNCH_CHECK("rigid_project_kernel");
    return FNX_OK;
}

int fnx_pbf_confirm_guess(int32_t N, float *xyz, const float *estimate_xyz, float *velocity, float secs, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(N >= 0 && xyz && estimate_xyz && velocity && secs != 0.f, "bad arguments");
    if (N == 0) return FNX_OK;
    solver_confirm_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, xyz, estimate_xyz, velocity, secs);
    FNX_LAUNCH_CHECK("solver_confirm_kernel");
    return FNX_OK;
}

int fnx_pbf_update_visual(void *grid, const float *estimate_xyz, const float *velocity, int32_t N, float *visual, int32_t V,
                          int32_t max_num_neighbors, float H, float secs, int32_t *kthV_scratch, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && estimate_xyz && velocity && (visual || V == 0) && kthV_scratch && N >= 0 && V >= 0 && max_num_neighbors > 0, "bad arguments");
    if (V == 0 || N == 0) return FNX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = grid_build(estimate_xyz, N, H, grid, st);
    if (rc) return rc;
    GridView g = grid_view(grid, N);
    advect_pack_vel_kernel<<<(N + 255) / 256, 256, 0, st>>>(g, N, velocity);
    FNX_LAUNCH_CHECK("advect_pack_vel_kernel");
    const float term1 = (float)(315.0 / (64.0 * 3.14159265358979323846 * pow((double)H, 9)));
    // visual += secs * sum_j w v_j / max(sum_j w, EPS): the P1 gather with the stored velocities as payload
    advect_fwd_kernel<true><<<(V + QPB - 1) / QPB, 128, 0, st>>>(g, 1.0f / H, estimate_xyz, nullptr, visual, V, kthV_scratch, max_num_neighbors, N,
                                                             H * H, term1, secs, PBF_EPS, 1.0f, visual, nullptr, nullptr);
    FNX_LAUNCH_CHECK("advect_fwd_kernel");
    return FNX_OK;
}

int fnx_pair_distance_loss(const void *grid, const float *pts, int32_t n, float cell, float threshold, float grad_scale, float *loss,
                           float *dL_dpts, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && loss && (pts || n == 0) && threshold > 0.f && threshold <= cell * 1.000001f, "bad arguments (need threshold <= cell)");
    FNX_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
    if (n == 0) return FNX_OK;
    GridView g = grid_view((void *)grid, n);
    pair_distance_kernel<<<(n + QPB - 1) / QPB, 128, 0, (cudaStream_t)stream>>>(g, 1.0f / cell, pts, n, threshold, grad_scale, loss, dL_dpts);
    FNX_LAUNCH_CHECK("pair_distance_kernel");
    return FNX_OK;
}

int fnx_knn3_mean_dist2(const void *grid, const float *pts, int32_t n, float cell, float *mean_dist2, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(grid && mean_dist2 && (pts || n == 0) && cell > 0.f, "bad arguments");
    if (n == 0) return FNX_OK;
    GridView g = grid_view((void *)grid, n);
    knn3_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(g, cell, pts, n, 24 * SUB, mean_dist2);
    FNX_LAUNCH_CHECK("knn3_kernel");
    return FNX_OK;
}

int fnx_pbf_next_tick_fwd(int32_t N, const float *e, const float *xyz, const float *buoyancy, const float *force, float secs,
                          float buoyancy_max_y, float scale_factor, float *X, float *Y, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(N >= 0 && e && xyz && buoyancy && force && (X || Y), "bad arguments");
    if (N == 0) return FNX_OK;
    next_tick_fwd_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, e, xyz, buoyancy, force, secs, buoyancy_max_y, scale_factor, X, Y);
    FNX_LAUNCH_CHECK("next_tick_fwd_kernel");
    return FNX_OK;
}

int fnx_pbf_combine_grad(int32_t N, const float *e, const float *buoyancy, float secs, float buoyancy_max_y, float scale_factor,
                         const float *dL_dX, const float *dL_dY, const float *estimate_xyz, float lambda_exyz, float *dL_de,
                         float *exyz_loss, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(N >= 0 && e && buoyancy && dL_de, "bad arguments");
    if (exyz_loss) FNX_CUDA_TRY(cudaMemsetAsync(exyz_loss, 0, sizeof(float), (cudaStream_t)stream));
    if (N == 0) return FNX_OK;
    combine_grad_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, e, buoyancy, secs, buoyancy_max_y, scale_factor, dL_dX, dL_dY,
                                                                         estimate_xyz, lambda_exyz, dL_de, exyz_loss, AdamFused{});
    FNX_LAUNCH_CHECK("combine_grad_kernel");
    return FNX_OK;
}

int fnx_pbf_combine_grad_adam(int32_t N, float *e, const float *buoyancy, float secs, float buoyancy_max_y, float scale_factor,
                              const float *dL_dX, const float *dL_dY, const float *estimate_xyz, float lambda_exyz, float *dL_de,
                              float *exyz_loss, float *exp_avg, float *exp_avg_sq, float lr, float beta1, float beta2, float eps,
                              int32_t *step_dev, float *bc_dev, const int32_t *skip_flag, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(N >= 0 && e && buoyancy && dL_de && exp_avg && exp_avg_sq && step_dev && bc_dev, "bad arguments");
    if (exyz_loss) FNX_CUDA_TRY(cudaMemsetAsync(exyz_loss, 0, sizeof(float), (cudaStream_t)stream));
    adam_prepare_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, beta1, beta2, bc_dev, skip_flag);
    FNX_LAUNCH_CHECK("adam_prepare_kernel");
    if (N == 0) return FNX_OK;
    AdamFused ad{e, exp_avg, exp_avg_sq, bc_dev, lr, beta1, beta2, eps};
    combine_grad_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, e, buoyancy, secs, buoyancy_max_y, scale_factor, dL_dX, dL_dY,
                                                                         estimate_xyz, lambda_exyz, dL_de, exyz_loss, ad);
    FNX_LAUNCH_CHECK("combine_grad_kernel");
    return FNX_OK;
}

int fnx_pbf_ratio_loss(int32_t N, const float *p_ratio, float weight, float *loss, float *dL_dpratio, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(N >= 0 && p_ratio && loss && dL_dpratio, "bad arguments");
    FNX_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
    if (N == 0) return FNX_OK;
    ratio_loss_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, p_ratio, weight, loss, dL_dpratio);
    FNX_LAUNCH_CHECK("ratio_loss_kernel");
    return FNX_OK;
}

int fnx_adam_step(int64_t n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float grad_scale, float lr,
                  float beta1, float beta2, float eps, int32_t step, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(n >= 0 && param && grad && exp_avg && exp_avg_sq && step >= 1, "bad arguments");
    if (n == 0) return FNX_OK;
    const float bc1 = (float)(1.0 - pow((double)beta1, step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, param, grad, exp_avg, exp_avg_sq, grad_scale, lr, beta1, beta2,
                                                                             eps, bc1, bc2_sqrt);
    FNX_LAUNCH_CHECK("adam_kernel");
    return FNX_OK;
}

int fnx_adam_step_dev(int64_t n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float grad_scale, float lr,
                      float beta1, float beta2, float eps, int32_t *step_dev, float *bc_dev, fnx_stream_t stream) {
    return fnx_adam_step_dev_gated(n, param, grad, exp_avg, exp_avg_sq, grad_scale, lr, beta1, beta2, eps, step_dev, bc_dev, nullptr, stream);
}

int fnx_adam_step_dev_gated(int64_t n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float grad_scale, float lr,
                            float beta1, float beta2, float eps, int32_t *step_dev, float *bc_dev, const int32_t *skip_flag,
                            fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(n >= 0 && param && grad && exp_avg && exp_avg_sq && step_dev && bc_dev, "bad arguments");
    adam_prepare_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, beta1, beta2, bc_dev, skip_flag);
    if (n > 0)
        adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, param, grad, exp_avg, exp_avg_sq, grad_scale, lr,
                                                                                     beta1, beta2, eps, bc_dev);
    FNX_LAUNCH_CHECK("adam_dev_kernel");
    return FNX_OK;
}

int fnx_scatter_min(int64_t n, const float *src, const int64_t *index, int32_t n_out, float *out, int64_t *arg, fnx_stream_t stream) {
    ProfScope _ps(SEC_PHYSICS, (cudaStream_t)stream);
    FNX_REQUIRE(n >= 0 && n_out >= 0 && out && arg && ((src && index) || n == 0), "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_out == 0) return FNX_OK;
    scatter_min_init_kernel<<<(n_out + 255) / 256, 256, 0, st>>>(n_out, out, (long long *)arg, n);
    if (n > 0) {
        scatter_min_val_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, src, (const long long *)index, out);
        scatter_min_arg_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, src, (const long long *)index, out, (long long *)arg);
    }
    scatter_min_fix_kernel<<<(n_out + 255) / 256, 256, 0, st>>>(n_out, out, (const long long *)arg, n);
    FNX_LAUNCH_CHECK("scatter_min kernels");
    return FNX_OK;
}

}  // extern "C"
