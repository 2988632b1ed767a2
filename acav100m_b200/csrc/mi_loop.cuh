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

// One per (parity, source rank), written by peers over NVLink.  The 128 bits of (key, payload) travel in three 64-bit
// words that each carry the low 16 bits of the iteration tag on top: a word is valid when its tag is the expected one
// (8-byte accesses are single-copy atomic), so the sender needs NO release fence -- three plain stores, one NVLink
// one-way latency -- and the receiver no acquire.  (Round 1 stored key, payload and then a flag with st.release.sys: the
// fence waits for the two payload stores to be acknowledged across the link before the flag may leave, ~2 us per
// iteration.)  A slot is rewritten every second iteration, so a stale word is two tags behind, never 65536.
struct MiMail {
    unsigned long long w[3];
    unsigned long long pad;
};

__device__ __forceinline__ void mail_store(MiMail *m, unsigned long long key, unsigned long long pay, unsigned int tag) {
    const unsigned long long t = (unsigned long long)(tag & 0xFFFFu) << 48;
    const unsigned long long w0 = t | (key & 0xFFFFFFFFFFFFull);
    const unsigned long long w1 = t | (key >> 48) | ((pay & 0xFFFFFFFFull) << 16);
    const unsigned long long w2 = t | (pay >> 32);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(&m->w[0]), "l"(w0) : "memory");
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(&m->w[1]), "l"(w1) : "memory");
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(&m->w[2]), "l"(w2) : "memory");
}

// Bounded wait for a peer's entry of this iteration (see wait_peer_tag in common.cuh for why bounded).
__device__ __forceinline__ bool mail_wait(const MiMail *m, unsigned int tag, unsigned long long limit_ns,
                                          unsigned long long &key, unsigned long long &pay) {
    const unsigned long long want = (unsigned long long)(tag & 0xFFFFu);
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
        unsigned long long w0, w1, w2;
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w0) : "l"(&m->w[0]) : "memory");
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w1) : "l"(&m->w[1]) : "memory");
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w2) : "l"(&m->w[2]) : "memory");
        if ((w0 >> 48) == want && (w1 >> 48) == want && (w2 >> 48) == want) {
            key = (w0 & 0xFFFFFFFFFFFFull) | ((w1 & 0xFFFFull) << 48);
            pay = ((w1 >> 16) & 0xFFFFFFFFull) | ((w2 & 0xFFFFFFFFull) << 32);
            return true;
        }
        if ((++spins & 1023u) == 0u) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > limit_ns) return false;
        }
    }
}

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
