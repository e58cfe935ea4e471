// common.cuh — shared device/host helpers for the sm_100a DMFB / MEDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>

#include "dmfb_b200.h"

namespace dmfb {

// ---------------------------------------------------------------- host side --
extern thread_local char g_last_error[256];
extern std::atomic<uint64_t> g_launches;

inline int cuda_fail(cudaError_t e, const char* what)
{
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return DMFB_ERR_CUDA;
}
#define DMFB_CUDA_TRY(expr)                                      \
    do {                                                         \
        cudaError_t _e = (expr);                                 \
        if (_e != cudaSuccess) return dmfb::cuda_fail(_e, #expr); \
    } while (0)

inline int gcd_int(int a, int b)
{
    while (b) { int t = a % b; a = b; b = t; }
    return a;
}

// Smallest number of envs whose rows (row_bytes each) form a span that is a multiple of 16 bytes,
// scaled up to roughly `target_bytes` per tile.  Tiles of that many envs start 16-byte aligned when
// the tensor base is, which is what the TMA bulk store needs.
// Rows so long (and odd) that even one alignment unit of them would not fit `limit_bytes` of shared memory get a
// smaller tile; its tiles then start unaligned and leave through the ordinary-store path of store_tile().
inline int pick_tile_envs(int row_bytes, int target_bytes, int max_envs, int limit_bytes = 160 * 1024)
{
    int unit = 16 / gcd_int(16, row_bytes);
    int e = unit;
    while (e * 2 <= max_envs && (e * 2) * row_bytes <= target_bytes) e *= 2;
    while (e > 1 && (long long)e * row_bytes > limit_bytes) e >>= 1;
    return e;
}

// -------------------------------------------------------------- device side --
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Shared-memory stores by 32-bit shared address (a generic pointer can make the compiler rebuild the shared window
// base in front of a predicated store).
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}

// Make this thread's generic-proxy shared-memory writes visible to the async (TMA) proxy.
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// TMA 1-D bulk store shared -> global (UBLKCP in SASS).  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes)
{
#ifdef DMFB_OBS_EVICT_FIRST
    // the observation stream is written once and not read again by this library: let it leave L2 first, so that the
    // small state arrays that every step re-reads stay resident
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes), "l"(policy) : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
#endif
}
// 32-bit load that asks L2 to keep the line (small, hot, randomly accessed tables)
__device__ __forceinline__ uint32_t ldg_u32_keep(const uint32_t* p)
{
#ifdef DMFB_BITS_EVICT_LAST
    uint64_t policy;
    uint32_t v;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ void tma_store_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Wait until the bulk stores of all committed groups have finished READING shared memory.
__device__ __forceinline__ void tma_store_wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Philox4x32-10 (Salmon et al. 2011) — counter-based, so a draw is a pure function of
// (seed, env, episode, step, agent) and does not depend on how envs are sharded over GPUs.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
// 53-bit uniform in [0,1), built like CPython's random.random(): (a>>5, b>>6) -> (a*2^26+b)/2^53
__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

enum : uint32_t { kStreamMove = 1, kStreamLayout = 2, kStreamDegrade = 3, kStreamBlocks = 4 };

// The reference's task / obstacle generators redraw until the draw is legal, i.e. for ever when the requested density
// cannot be placed (dmfb.py:212-224,246-250, meda.py:213-233).  A kernel that never ends takes the GPU with it, so the
// device generators give up after this many rounds: they keep the previous layout and raise DMFB_STATUS_SAMPLER_GAVE_UP.
constexpr uint32_t kMaxSamplerRounds = 1u << 22;

// splitmix64 finaliser (Steele/Lea/Flood 2014): cheap counter-based generator for the task sampler
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint4 env_random(uint64_t seed, uint32_t stream, int64_t env, uint32_t episode,
                                            uint32_t a, uint32_t b)
{
    const uint2 key = make_uint2((uint32_t)seed ^ (stream * 0x85EBCA6Bu), (uint32_t)(seed >> 32) ^ (uint32_t)(env >> 32));
    return philox4x32_10(make_uint4((uint32_t)env, episode, a, b), key);
}

__device__ __forceinline__ int load_action(const void* actions, int elem_size, size_t idx)
{
    if (elem_size == 1) return (int)static_cast<const int8_t*>(actions)[idx];
    if (elem_size == 4) return (int)static_cast<const int32_t*>(actions)[idx];
    return (int)static_cast<const long long*>(actions)[idx];
}

// Cooperative store of a finished shared-memory tile to global memory.
//  - fast path: global address 16-byte aligned -> one TMA bulk store for the 16-byte multiple part
//    (issued by thread 0) + at most 15 tail bytes by ordinary stores;
//  - slow path (unaligned tensor base): byte stores.
// Must be called by all threads of the CTA after their last write to the tile; contains the
// proxy fence + barrier, and thread 0 returns only after the TMA has finished reading smem.
// store_tile_issue() returns true to the thread that has a bulk store in flight: it must call
// tma_store_wait_read_all() before the tile is reused or the CTA exits (work that does not touch the tile can go in
// between and overlaps the drain).
__device__ __forceinline__ bool store_tile_issue(int8_t* __restrict__ gdst, const int8_t* tile, uint32_t nbytes)
{
    const bool aligned = ((reinterpret_cast<uintptr_t>(gdst) & 15) == 0);
    if (aligned) {
        fence_proxy_async_smem();
        __syncthreads();
        const uint32_t bulk = nbytes & ~15u;
        if (threadIdx.x == 0 && bulk) {
            tma_store_1d(gdst, tile, bulk);
            tma_store_commit();
        }
        for (uint32_t b = bulk + threadIdx.x; b < nbytes; b += blockDim.x) gdst[b] = tile[b];
        return threadIdx.x == 0 && bulk;
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbytes; b += blockDim.x) gdst[b] = tile[b];
    return false;
}
__device__ __forceinline__ void store_tile(int8_t* __restrict__ gdst, const int8_t* tile, uint32_t nbytes)
{
    if (store_tile_issue(gdst, tile, nbytes)) tma_store_wait_read_all();
}

// Store only the rows (row_bytes each) whose flag is set; used by masked resets.
__device__ __forceinline__ void store_rows_masked(int8_t* __restrict__ gdst, const int8_t* tile, int n_rows,
                                                  int row_bytes, const uint8_t* row_flag)
{
    __syncthreads();
    const bool w4 = ((reinterpret_cast<uintptr_t>(gdst) & 3) == 0) && (row_bytes % 4 == 0);
    for (int r = 0; r < n_rows; ++r) {
        if (!row_flag[r]) continue;
        const int8_t* src = tile + (size_t)r * row_bytes;
        int8_t* dst = gdst + (size_t)r * row_bytes;
        if (w4) {
            for (int k = threadIdx.x; k < row_bytes / 4; k += blockDim.x)
                reinterpret_cast<uint32_t*>(dst)[k] = reinterpret_cast<const uint32_t*>(src)[k];
        } else {
            for (int k = threadIdx.x; k < row_bytes; k += blockDim.x) dst[k] = src[k];
        }
    }
}

#endif  // __CUDACC__
}  // namespace dmfb
