// common.cuh -- shared helpers for libfnx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fnx.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfnx is written for sm_100a (B200) only"
#endif

namespace fnx {

// thread-local last error -------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define FNX_CUDA_TRY(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            fnx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FNX_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

extern unsigned long long g_launches;  // hand-written kernels launched by this process (fnx_launch_count)

#define FNX_LAUNCH_CHECK(name)                                                                      \
    do {                                                                                            \
        fnx::g_launches++;                                                                          \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess) {                                                                    \
            fnx::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FNX_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define FNX_REQUIRE(cond, ...)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            fnx::set_error(__VA_ARGS__);                                                            \
            return FNX_ERR_INVALID;                                                                 \
        }                                                                                           \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve a typed array out of a byte chunk, 256-B aligned
template <typename T>
static inline T *carve(char *&p, size_t count) {
    p = (char *)align_up((size_t)p, 256);
    T *r = (T *)p;
    p += count * sizeof(T);
    return r;
}

static inline int ceil_log2_u64(uint64_t n) {  // smallest b with (1<<b) >= n
    int b = 0;
    while (((uint64_t)1 << b) < n) b++;
    return b;
}

// device helpers ------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine; SASS UBLKCP). bytes % 16 == 0, both 16-B aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// vector reductions to global memory (sm_90+): one L2 atomic transaction for 4 / 2 floats
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add(float *addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace fnx

// ---------------------------------------------------------------------------------------------------------------
// optional per-section CUDA-event timing (fnx_profile_*): lets bench.py measure a kernel's launch duration live,
// on the launching stream, inside its timed region
// ---------------------------------------------------------------------------------------------------------------
namespace fnx {
enum Section { SEC_PREPROCESS = 0, SEC_DEPTH_SORT, SEC_EMIT, SEC_TILE_SORT, SEC_PACK, SEC_BLEND_FWD, SEC_BLEND_BWD, SEC_GEOM_BWD,
               SEC_IMAGE_LOSS, SEC_PHYSICS, SEC_COUNT };
void prof_begin(int section, cudaStream_t st);
void prof_end(int section, cudaStream_t st);
struct ProfScope {
    int s; cudaStream_t st;
    ProfScope(int section, cudaStream_t stream) : s(section), st(stream) { prof_begin(s, st); }
    ~ProfScope() { prof_end(s, st); }
};
}  // namespace fnx
