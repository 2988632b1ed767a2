// Pieces shared by the two persistent greedy-MI loops (mi_persistent.cu: candidate stream; mi_cells.cu: cell
// index): the fp32 running-sum update of the reference (mi.py:339-340), the per-CTA publication record, the
// NVLink mailbox record and the grid barrier.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace acav {

constexpr int kSmallCounts = 256;            // per-iteration table of tN(x) for counts below this

__device__ __forceinline__ float xlogx_cnt(uint32_t k, float f0, const float *__restrict__ logs) {
    return k == 0 ? f0 : __fmul_rn((float)k, __ldg(logs + k));
}
__device__ __forceinline__ float bump_sum(float prev, uint32_t k, float f0, const float *__restrict__ logs) {
    return __fadd_rn(__fsub_rn(prev, xlogx_cnt(k, f0, logs)), xlogx_cnt(k + 1, 0.f, logs));
}

struct MiPub {                       // one per CTA and iteration parity: the CTA's best candidate
    unsigned long long key;          // (orderable gain << 32) | (0xFFFFFFFF - global position), 0 = none
    unsigned long long payload;      // (c1 << 48) | (c2 << 32) | table count x of that cell
};

struct MiMail {                      // one per (parity, source rank), written by peers over NVLink
    unsigned long long key;
    unsigned long long payload;
    unsigned int seq;                // iteration tag, stored last with release semantics
    unsigned int pad[3];
};

constexpr int kMaxWorld = 16;

__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int *gen = bar + 1;
        const unsigned int g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace acav
