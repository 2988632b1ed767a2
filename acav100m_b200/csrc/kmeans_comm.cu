// Multi-GPU k-means step over NVLink peer memory, replacing the two NCCL all-reduces of the reference's
// distributed branch (clustering/code/sgd_clustering.py:114-115 histogram, :125-126 deltas; mps/distributed.py:139-155).
//
// Every rank owns k_own = ceil(K / world) consecutive centroids.  Per step:
//   1. km_hist_exchange_kernel   each rank stores its batch histogram into every rank's arena (4*K bytes per peer),
//                                raises a flag, waits for the others' flags, adds the histograms in rank order
//                                (exact: integers), counts += histogram (:120) and takes the lr decision (:116-119);
//   2. km_update_*<kUpdPush>     (kmeans_update.cu) the row-ordered per-centroid sums of fl32(x*lr) are stored straight
//                                into the OWNER's receive buffer, slot [source rank] -- the compute kernel is the sender;
//   3. km_signal_kernel          system-scope fence + "my deltas are there" flag on every rank;
//   4. km_reduce_broadcast_kernel the owner waits for all flags, adds the world deltas of its centroids IN RANK ORDER
//                                (so the result is the single-process sum order of oracle/kmeans_oracle.py::sgd_step_world,
//                                independent of timing, identical on every rank), applies the decay (:121) and the sum
//                                (:127), and stores the new rows into its own centers and into every peer's inbox;
//   5. km_signal_kernel, km_gather_kernel   flag + copy of the other owners' rows from the inbox into centers.
// Traffic per rank and step: 2 * (world-1)/world * 4*K*D bytes over NVLink (reduce-scatter + all-gather volume), two
// flag round trips, no host involvement, CUDA-graph capturable (the step tag lives in device memory).
// Every wait is bounded (wait_peer_tag): a rank that never arrives sets the status word instead of hanging the others.
#include "common.cuh"
#include "kernels.cuh"

namespace acav {

namespace {

constexpr size_t kAlign = 256;
__host__ __device__ inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

// arena layout (identical on every rank)
struct KmArena {
    size_t flags, hist, red, inbox, total;
};
__host__ __device__ inline KmArena km_arena_layout(int32_t world, int32_t k, int32_t d) {
    const int32_t k_own = (k + world - 1) / world;
    KmArena a;
    a.flags = 0;                                                         // uint32 [3][kKmMaxWorld]: hist, deltas, rows
    a.hist = align_up(3 * kKmMaxWorld * sizeof(unsigned int));           // float [world][k]
    a.red = a.hist + align_up((size_t)world * k * sizeof(float));        // float [world][k_own][d]
    a.inbox = a.red + align_up((size_t)world * k_own * d * sizeof(float));   // float [world * k_own][d]
    a.total = a.inbox + align_up((size_t)world * k_own * d * sizeof(float));
    return a;
}

__device__ __forceinline__ unsigned int *flag_ptr(unsigned char *arena, int which, int src) {
    return reinterpret_cast<unsigned int *>(arena) + which * kKmMaxWorld + src;
}

__global__ void __launch_bounds__(1024)
km_hist_exchange_kernel(KmComm c, const float *__restrict__ counts_b, double lr, float *__restrict__ counts_global,
                        float *__restrict__ lr_eff, int32_t *__restrict__ fallback, float *__restrict__ counts) {
    __shared__ unsigned int sh_seq;
    __shared__ float wmax[32];
    const KmArena L = km_arena_layout(c.world, c.k, c.d);
    if (threadIdx.x == 0) { sh_seq = *c.seq + 1u; *c.seq = sh_seq; }
    __syncthreads();
    const unsigned int seq = sh_seq;
    for (int r = 0; r < c.world; ++r) {
        float *dst = reinterpret_cast<float *>(c.arena[r] + L.hist) + (size_t)c.rank * c.k;
        for (int32_t i = threadIdx.x; i < c.k; i += blockDim.x) dst[i] = counts_b[i];
    }
    __threadfence_system();
    __syncthreads();
    bool timed_out = false;
    if ((int)threadIdx.x < c.world) {
        st_release_sys(flag_ptr(c.arena[threadIdx.x], 0, c.rank), seq);
        timed_out = !wait_peer_tag(flag_ptr(c.arena[c.rank], 0, threadIdx.x), seq, c.spin_limit_ns);
    }
    if (__syncthreads_or(timed_out) && threadIdx.x == 0) *reinterpret_cast<volatile int *>(c.status) = 1;
    const float *hist = reinterpret_cast<const float *>(c.arena[c.rank] + L.hist);
    float m = 0.f;
    for (int32_t i = threadIdx.x; i < c.k; i += blockDim.x) {
        float sum = 0.f;
        for (int r = 0; r < c.world; ++r)
            sum = __fadd_rn(sum, *reinterpret_cast<const volatile float *>(hist + (size_t)r * c.k + i));
        counts_global[i] = sum;
        counts[i] = __fadd_rn(counts[i], sum);                               // :120
        m = fmaxf(m, sum);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x % kWarp == 0) wmax[threadIdx.x / kWarp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {                                                  // :116-119, python-float arithmetic
        float mm = 0.f;
        for (int w = 0; w < (int)(blockDim.x / kWarp); ++w) mm = fmaxf(mm, wmax[w]);
        double eff = lr;
        if ((double)mm * lr >= 1.0) {
            eff = 0.5 / (double)mm;
            if (fallback) *fallback += 1;
        }
        *lr_eff = (float)eff;
    }
}

// which = 1: "my deltas are in your receive buffer", which = 2: "my centroid rows are in your inbox"
__global__ void km_signal_kernel(KmComm c, int which) {
    __threadfence_system();
    if ((int)threadIdx.x < c.world) st_release_sys(flag_ptr(c.arena[threadIdx.x], which, c.rank), *c.seq);
}

// one thread = 4 columns of one owned centroid
__global__ void __launch_bounds__(256)
km_reduce_broadcast_kernel(KmComm c, const float *__restrict__ counts_global, const float *__restrict__ lr_eff,
                           float *__restrict__ centers) {
    const KmArena L = km_arena_layout(c.world, c.k, c.d);
    const unsigned int seq = *c.seq;
    bool timed_out = false;
    if ((int)threadIdx.x < c.world)
        timed_out = !wait_peer_tag(flag_ptr(c.arena[c.rank], 1, threadIdx.x), seq, c.spin_limit_ns);
    if (__syncthreads_or(timed_out) && threadIdx.x == 0) *reinterpret_cast<volatile int *>(c.status) = 1;
    const int32_t d4 = c.d / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t cl = (int32_t)(idx / d4), col = (int32_t)(idx - (int64_t)cl * d4) * 4;
    const int32_t cg = c.rank * c.k_own + cl;
    if (cl >= c.k_own || cg >= c.k) return;
    const float *red = reinterpret_cast<const float *>(c.arena[c.rank] + L.red);
    float4 acc = __ldcv(reinterpret_cast<const float4 *>(red + (size_t)cl * c.d + col));          // rank 0's delta
    for (int r = 1; r < c.world; ++r) {
        const float4 t = __ldcv(reinterpret_cast<const float4 *>(red + ((size_t)r * c.k_own + cl) * c.d + col));
        acc.x = __fadd_rn(acc.x, t.x); acc.y = __fadd_rn(acc.y, t.y);
        acc.z = __fadd_rn(acc.z, t.z); acc.w = __fadd_rn(acc.w, t.w);
    }
    const float decay = __fsub_rn(1.f, __fmul_rn(counts_global[cg], *lr_eff));                   // :121
    float4 *cp = reinterpret_cast<float4 *>(centers + (size_t)cg * c.d + col);
    float4 v = *cp;
    v.x = __fadd_rn(__fmul_rn(v.x, decay), acc.x); v.y = __fadd_rn(__fmul_rn(v.y, decay), acc.y);   // :127
    v.z = __fadd_rn(__fmul_rn(v.z, decay), acc.z); v.w = __fadd_rn(__fmul_rn(v.w, decay), acc.w);
    *cp = v;
    for (int r = 0; r < c.world; ++r) {
        if (r == c.rank) continue;
        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(c.arena[r] + L.inbox) + (size_t)cg * c.d + col) = v;
    }
}

// rows owned by the other ranks: inbox -> centers
__global__ void __launch_bounds__(256)
km_gather_kernel(KmComm c, float *__restrict__ centers) {
    const KmArena L = km_arena_layout(c.world, c.k, c.d);
    const unsigned int seq = *c.seq;
    bool timed_out = false;
    if ((int)threadIdx.x < c.world)
        timed_out = !wait_peer_tag(flag_ptr(c.arena[c.rank], 2, threadIdx.x), seq, c.spin_limit_ns);
    if (__syncthreads_or(timed_out) && threadIdx.x == 0) *reinterpret_cast<volatile int *>(c.status) = 1;
    const int32_t d4 = c.d / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t cg = (int32_t)(idx / d4), col = (int32_t)(idx - (int64_t)cg * d4) * 4;
    if (cg >= c.k || cg / c.k_own == c.rank) return;
    const float *inbox = reinterpret_cast<const float *>(c.arena[c.rank] + L.inbox);
    *reinterpret_cast<float4 *>(centers + (size_t)cg * c.d + col) =
        __ldcv(reinterpret_cast<const float4 *>(inbox + (size_t)cg * c.d + col));
}

}  // namespace

size_t km_comm_arena_bytes(int32_t world, int32_t k, int32_t d) { return km_arena_layout(world, k, d).total; }

KmPush km_comm_push_target(const KmComm &c) {
    KmPush p;
    const KmArena L = km_arena_layout(c.world, c.k, c.d);
    for (int r = 0; r < kKmMaxWorld; ++r)
        p.red[r] = r < c.world ? reinterpret_cast<float *>(c.arena[r] + L.red) : nullptr;
    p.k_own = c.k_own;
    p.rank = c.rank;
    return p;
}

int launch_km_hist_exchange(const KmComm &c, const float *counts_b_local, double lr, float *counts_global,
                            float *lr_eff, int32_t *fallback, float *counts, cudaStream_t st) {
    km_hist_exchange_kernel<<<1, 1024, 0, st>>>(c, counts_b_local, lr, counts_global, lr_eff, fallback, counts);
    ACAV_LAUNCH_CHECK();
    return 0;
}

int launch_km_reduce_broadcast(const KmComm &c, const float *counts_global, const float *lr_eff, float *centers,
                               cudaStream_t st) {
    km_signal_kernel<<<1, 32, 0, st>>>(c, 1);
    ACAV_LAUNCH_CHECK();
    const int64_t own = (int64_t)c.k_own * (c.d / 4);
    km_reduce_broadcast_kernel<<<(unsigned)ceil_div(own, 256), 256, 0, st>>>(c, counts_global, lr_eff, centers);
    ACAV_LAUNCH_CHECK();
    km_signal_kernel<<<1, 32, 0, st>>>(c, 2);
    ACAV_LAUNCH_CHECK();
    const int64_t all = (int64_t)c.k * (c.d / 4);
    km_gather_kernel<<<(unsigned)ceil_div(all, 256), 256, 0, st>>>(c, centers);
    ACAV_LAUNCH_CHECK();
    return 0;
}

}  // namespace acav
