// Shared device/host helpers for libacav_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/acav_b200.h"

#define ACAV_CUDA_TRY(expr)                                  \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return (int)_e;               \
    } while (0)

#define ACAV_LAUNCH_CHECK()                                  \
    do {                                                     \
        cudaError_t _e = cudaPeekAtLastError();              \
        if (_e != cudaSuccess) return (int)_e;               \
    } while (0)

namespace acav {

constexpr int kWarp = 32;
constexpr int kMaxDevices = 64;

// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device, once per device and size (function
// attributes are per device; a process may drive several).  `done` is a zero-initialised static of the call site.
template <typename F>
inline int ensure_dynamic_smem(F *func, size_t bytes, size_t (&done)[kMaxDevices]) {
    int dev = 0;
    ACAV_CUDA_TRY(cudaGetDevice(&dev));
    const int slot = dev >= 0 && dev < kMaxDevices ? dev : 0;
    if (bytes > done[slot] || dev != slot) {                   // devices beyond the table: set every time
        ACAV_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        done[slot] = bytes;
    }
    return 0;
}

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (the k-means step's chain of small kernels) ----
// A kernel launched with launch_pdl() may be scheduled while its predecessor in the stream is still draining; it
// must not touch memory before pdl_begin(), which (a) lets ITS successor be scheduled early in turn and (b) waits until
// the predecessor grid has completed and its writes are visible.  Launched the ordinary way, pdl_begin() is a no-op.
// Works in streams and under stream capture (the graph gets programmatic edges).  ACAV_NO_PDL=1 launches plainly.
__device__ __forceinline__ void pdl_begin() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled();                                    // capi.cu (reads ACAV_NO_PDL once)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// fp32 -> uint32 whose unsigned order equals the float order (-0 canonicalised to +0).
__device__ __forceinline__ uint32_t orderable(float s) {
    if (s == 0.0f) s = 0.0f;
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// (score, position) -> key; larger key = larger score, then smaller position.  0 = "nothing".
__device__ __forceinline__ unsigned long long make_key(float s, uint32_t pos) {
    return ((unsigned long long)orderable(s) << 32) | (unsigned long long)(0xFFFFFFFFu - pos);
}
__device__ __forceinline__ uint32_t key_pos(unsigned long long key) {
    return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
}
__device__ __forceinline__ float key_score(unsigned long long key) {
    return from_orderable((uint32_t)(key >> 32));
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- cross-GPU signalling over NVLink peer memory (greedy-MI mailbox, k-means centroid exchange) ----
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Bounded wait for a peer GPU's flag: the tag of this iteration must appear within `limit_ns` (a rank that
// never launched -- failed setup, host exception, different n_picks -- must not hang the others inside a
// cooperative kernel).  The timer is read once per 1024 polls, so the fast path is the bare acquire load.
__device__ __forceinline__ bool wait_peer_tag(const unsigned int *seq, unsigned int tag, unsigned long long limit_ns) {
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_sys(seq) != tag) {
        if ((++spins & 1023u) == 0u) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > limit_ns) return false;
        }
    }
    return true;
}

}  // namespace acav
