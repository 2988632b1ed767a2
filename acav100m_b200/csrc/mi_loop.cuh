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

// Returns true when the launch is being aborted (`abort_word` != 0, set by a CTA whose wait for a peer GPU timed
// out): CTAs spinning here must not wait for CTAs that already left the loop.  The word is looked at once per 4096
// polls, and never when it is null (single-GPU runs).
__device__ __forceinline__ bool grid_barrier(unsigned int *bar, unsigned int nblocks, const int *abort_word = nullptr) {
    int aborted = 0;
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
            unsigned int spins = 0;
            while (*gen == g) {
                if (abort_word && (++spins & 4095u) == 0u && *reinterpret_cast<const volatile int *>(abort_word) != 0) {
                    aborted = 1;
                    break;
                }
            }
        }
        __threadfence();
    }
    return __syncthreads_or(aborted) != 0;
}

// Status word of the persistent loops (acav_mi_status): why a launch stopped before n_picks iterations.
constexpr int kMiRunOk = 0;          // all iterations done, or every rank ran out of candidates
constexpr int kMiRunPeerTimeout = 1; // a peer's mailbox entry did not arrive within the spin limit

}  // namespace acav
