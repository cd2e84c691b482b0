// View-sharded lift, exchange step: reduce + finalise in ONE kernel over NVLink peer memory.
//
// Every rank has run nd_lift_accumulate on its own views into a peer-mapped segment [S1 (C*N) | S2 (C*N) | count (N)].
// Instead of an NCCL all-reduce of the 52.5 MB accumulators followed by a finalise pass on every rank, rank r
//   * owns a slice of the channels,
//   * LOADS the partial sums of its slice from every rank's segment (P2P reads over NVLink / NVSwitch),
//   * adds them in rank order, turns them into mean / exp(-var) with the GLOBAL view count (nerfdet.py:171-181), and
//   * STORES the result rows into every rank's output buffers (P2P writes),
// so the accumulators cross the links once (reduce-scatter traffic), the results once (all-gather traffic), and no
// reduced accumulator is ever written back to memory.  The hand-shake between the ranks is a pair of epoch flags per
// peer inside the segments (st.release.sys / ld.acquire.sys): `ready` = my accumulators of this epoch are complete
// (raised by the first CTA of the finalise kernel, i.e. after the accumulate kernels in stream order), `done` = all my
// reads of your accumulators and all my writes into your outputs have been performed (raised by the last CTA).
// A one-block wait kernel holds the stream until every peer is `done`.  All spins are bounded: a peer that never shows up
// raises the error word of EVERY rank's flag block after a bounded wait (default 4 s) instead of hanging the GPU; ranks
// that see the error skip the reduce / store phase and fill the rows they own with NaN in their own outputs.
//
// The segments are plain cudaMalloc allocations shared with CUDA IPC handles (nd_peer_alloc / nd_peer_open): the only
// place where the library owns device memory, because a handle must cover a whole allocation.
#include <stdlib.h>
#include <string.h>

#include "nd_common.cuh"

namespace nd {

constexpr int kMaxPeers = ND_MAX_PEERS;
constexpr int kPeerThreads = 128;

struct PeerArgs {
    // rank g's accumulators: S1 / S2 rows [c * N + n], counts [n] (S2 null when no variance is wanted)
    const float *s1[kMaxPeers], *s2[kMaxPeers], *cntp[kMaxPeers];
    float *mean[kMaxPeers];
    float *cov[kMaxPeers];           // all null: no covariance wanted
    uint32_t *flags[kMaxPeers];      // [0, P): ready, [P, 2P): done, [2P]: CTA counter, [2P + 1]: error
    int world, rank;
    int owner;                       // >= 0: only this rank receives the finished rows (its scene); -1: every rank
    uint32_t epoch;
    unsigned long long timeout_ns;   // bound of every wait for a peer
    int n_views_total, channels, c_begin, c_end, ch_per_cta, n_tiles, n_items;
    int64_t n_vox;
    const float *alpha;
    int64_t *count;
    // NVLS: multicast addresses of the accumulators / outputs (one address reaches every rank's copy through the switch)
    const float *acc_mc;
    float *mean_mc, *cov_mc;
};

// in-switch reduction: one load returns the sum over all ranks' copies; one store lands in all of them (PTX ISA 8.1, sm_90+)
__device__ __forceinline__ void ld_reduce_mc(const float *p, float (&r)[4]) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void st_mc(float *p, const float (&r)[4]) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]),
                 "f"(r[3])
                 : "memory");
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// waits until *p has reached `epoch` (wrap-safe); false after `timeout_ns`
__device__ __forceinline__ bool spin_until(const uint32_t *p, uint32_t epoch, unsigned long long timeout_ns) {
    const unsigned long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
        __nanosleep(64);
        if (global_ns() - t0 > timeout_ns) return false;
    }
    return true;
}
// A wait that timed out poisons the step on EVERY rank: the error word of every rank's flag block is raised (a rank
// that sees it set skips its reduce / store phase as well), so that no rank hands out rows built from stale or
// partial accumulators without an error it can see.
__device__ __forceinline__ void raise_error_everywhere(uint32_t *const *flags, int world) {
    for (int g = 0; g < world; ++g) atomicExch_system(flags[g] + 2 * kMaxPeers + 1, 1u);
}

template <int V> __device__ __forceinline__ void ld_vec(const float *p, float (&r)[V]);
template <> __device__ __forceinline__ void ld_vec<4>(const float *p, float (&r)[4]) {
    const float4 t = __ldcg(reinterpret_cast<const float4 *>(p));
    r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
}
template <> __device__ __forceinline__ void ld_vec<1>(const float *p, float (&r)[1]) { r[0] = __ldcg(p); }
template <int V> __device__ __forceinline__ void st_vec(float *p, const float (&r)[V]);
template <> __device__ __forceinline__ void st_vec<4>(float *p, const float (&r)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(r[0], r[1], r[2], r[3]);
}
template <> __device__ __forceinline__ void st_vec<1>(float *p, const float (&r)[1]) { *p = r[0]; }

// grid = (voxel tiles of kPeerThreads * V, channel sub-slices of this rank's slice)
// G = compile-time bound of the world size (1, 2, 4, 8), U = channels per iteration: G * U * 2 vector loads are in flight
// per thread before the first is consumed -- remote loads take ~2 us over NVLink, so the link rate is set by the bytes
// in flight (CTAs per SM x loads per thread), not by the instruction count.
// MC: the sums come from multimem.ld_reduce on the multicast address (the switch adds the ranks' copies: (G-1) x fewer
// bytes arrive than with per-peer loads) and the rows leave with one multimem.st instead of G stores.
// T = threads per CTA: 128 (wide grid, one work item per CTA) or 512 (NARROW grid: a few fat CTAs, one per SM, each
// looping over work items -- the exchange is link-bound, so ~20 SMs carry it while the other SMs run the next scene's
// accumulate; fat CTAs because the block scheduler spreads small CTAs over all SMs, where none would leave room for a
// persistent lift CTA).
// Tried in round 2 and removed: the same narrow grid with the SM's copy engine doing the transport (cp.async.bulk pieces of
// 4 KB from every rank into a 192 KB shared-memory ring, bulk stores of the finished rows).  Parity-green on two B200s, but a
// CTA moved only ~6 GB/s over NVLink that way (301 vs 137 us per step at 2 GPUs with 28 CTAs, 489 us with 16): the copy engine
// keeps far fewer remote requests in flight than 512 threads x 16 vector loads do.
template <int V, int G, int U, bool MC, int T, bool COV>
__global__ void __launch_bounds__(T, T > 128 ? 1 : ((G * U <= 4) ? 6 : 4))
k_lift_finalize_peers(const PeerArgs a) {
    static_assert(!MC || (V == 4 && G == 1), "multicast instantiation: 16-byte vectors, one (reduced) load per row");
    const int P = kMaxPeers;
    uint32_t *my_flags = a.flags[a.rank];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    // ---- hand-shake: my accumulators are complete; wait for everybody else's ----
    if (blockIdx.x == 0 && threadIdx.x < a.world)
        st_release_sys(a.flags[threadIdx.x] + a.rank, a.epoch);
    if (threadIdx.x < a.world && threadIdx.x != a.rank) {
        if (!spin_until(my_flags + threadIdx.x, a.epoch, a.timeout_ns)) {
            s_bad = 1;
            raise_error_everywhere(a.flags, a.world);
        }
    }
    if (threadIdx.x == 0 && ld_acquire_sys(my_flags + 2 * P + 1) != 0u) s_bad = 1;     // raised here or by a peer, now or earlier
    __syncthreads();
    const bool bad = s_bad != 0;

    const int64_t cn = (int64_t)a.channels * a.n_vox;
    const int64_t cnt_off = COV ? 2 * cn : cn;                  // accumulators: [S1 | S2 | count], or [S1 | count] without the variance
    // work item = (voxel tile of T * V voxels, channel sub-slice); wide grid: exactly one item per CTA
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int tile = item % a.n_tiles, sub = item / a.n_tiles;
    const int64_t n = ((int64_t)tile * T + threadIdx.x) * V;
    if (bad) {
        // no reduce, no peer traffic: the rows this CTA owns become NaN in the LOCAL outputs (the peers' copies are
        // theirs to poison; their error words are raised), the error word makes the step a hard error on the host
        if (n < a.n_vox) {
            const int c0 = a.c_begin + sub * a.ch_per_cta, c1 = min(a.c_end, c0 + a.ch_per_cta);
            float nanv[V];
#pragma unroll
            for (int j = 0; j < V; ++j) nanv[j] = __int_as_float(0x7fc00000);
            for (int c = c0; c < c1; ++c) {
                st_vec<V>(a.mean[a.rank] + (int64_t)c * a.n_vox + n, nanv);
                if constexpr (COV) st_vec<V>(a.cov[a.rank] + (int64_t)c * a.n_vox + n, nanv);
            }
        }
        continue;
    }
    if (n < a.n_vox) {
        float cnt[V];
#pragma unroll
        for (int j = 0; j < V; ++j) cnt[j] = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (MC || g < a.world) {
                float t[V];
                if constexpr (MC) ld_reduce_mc(a.acc_mc + cnt_off + n, t);
                else ld_vec<V>(a.cntp[g] + n, t);
#pragma unroll
                for (int j = 0; j < V; ++j) cnt[j] += t[j];
            }
        }
        float al[V], inv[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            al[j] = a.alpha != nullptr ? __ldg(a.alpha + n + j) : 1.0f;
            inv[j] = (float)a.n_views_total - cnt[j];
        }
        if (sub == 0 && a.count != nullptr) {
#pragma unroll
            for (int j = 0; j < V; ++j) a.count[n + j] = (int64_t)cnt[j];
        }
        const int c0 = a.c_begin + sub * a.ch_per_cta;
        const int c1 = min(a.c_end, c0 + a.ch_per_cta);
        for (int cb = c0; cb < c1; cb += U) {
            float p1[U][G][V], p2[COV ? U : 1][COV ? G : 1][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {                       // all loads of the U channels in flight together
                const int64_t o = (int64_t)min(cb + u, c1 - 1) * a.n_vox + n;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if constexpr (MC) {
                        ld_reduce_mc(a.acc_mc + o, p1[u][g]);
                        if constexpr (COV) ld_reduce_mc(a.acc_mc + cn + o, p2[u][g]);
                    } else if (g < a.world) {
                        ld_vec<V>(a.s1[g] + o, p1[u][g]);
                        if constexpr (COV) ld_vec<V>(a.s2[g] + o, p2[u][g]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (cb + u >= c1) break;
                const int64_t o = (int64_t)(cb + u) * a.n_vox + n;
                float s1[V], s2[V];
#pragma unroll
                for (int j = 0; j < V; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
#pragma unroll
                for (int g = 0; g < G; ++g) {                   // rank order: every rank would get the same bits
                    if (MC || g < a.world) {
#pragma unroll
                        for (int j = 0; j < V; ++j) {
                            s1[j] += p1[u][g][j];
                            if constexpr (COV) s2[j] += p2[u][g][j];
                        }
                    }
                }
                float m[V], cv[V];
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    m[j] = 0.f;
                    cv[j] = 0.f;
                    if (cnt[j] > 0.f) {                         // same formula as k_lift_finalize (lift.cu)
                        const float mm = s1[j] / cnt[j];
                        if constexpr (COV) {
                            float ssd = fmaxf(fmaf(-mm, s1[j], s2[j]), 0.0f);
                            ssd = fmaf(inv[j] * mm, mm, ssd);
                            cv[j] = expf(-(ssd / cnt[j]));
                        }
                        m[j] = mm * al[j];
                    }
                }
#pragma unroll
                if constexpr (MC) {
                    st_mc(a.mean_mc + o, m);
                    if constexpr (COV) st_mc(a.cov_mc + o, cv);
                } else {
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        if (g < a.world && (a.owner < 0 || g == a.owner)) {
                            st_vec<V>(a.mean[g] + o, m);
                            if constexpr (COV) st_vec<V>(a.cov[g] + o, cv);
                        }
                    }
                }
            }
        }
    }
    }
    // ---- completion: the last CTA tells every peer that this rank is done with their segments ----
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = gridDim.x;
        const uint32_t prev = atomicAdd(my_flags + 2 * P, 1u);
        if (prev == total - 1) {
            my_flags[2 * P] = 0u;                               // ready for the next epoch (next launch is stream-ordered)
            __threadfence_system();
            for (int g = 0; g < a.world; ++g) st_release_sys(a.flags[g] + P + a.rank, a.epoch);
        }
    }
}

__global__ void k_peer_wait_done(const PeerArgs a) {
    uint32_t *my_flags = a.flags[a.rank];
    if (threadIdx.x < a.world) {
        if (!spin_until(my_flags + kMaxPeers + threadIdx.x, a.epoch, a.timeout_ns)) raise_error_everywhere(a.flags, a.world);
    }
}

}  // namespace nd

using namespace nd;

extern "C" {

int nd_peer_alloc(size_t bytes, void **ptr, unsigned char *handle64_host) {
    ND_REQUIRE(ptr && handle64_host && bytes > 0, ND_ERR_BAD_ARG, "nd_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        set_error("nd_peer_alloc: %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    memcpy(handle64_host, &h, 64);
    *ptr = p;
    return ND_OK;
}

int nd_peer_open(const unsigned char *handle64_host, void **ptr) {
    ND_REQUIRE(ptr && handle64_host, ND_ERR_BAD_ARG, "nd_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("nd_peer_open: %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    *ptr = p;
    return ND_OK;
}

int nd_peer_close(void *ptr) {
    if (ptr == nullptr) return ND_OK;
    const cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) {
        set_error("nd_peer_close: %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    return ND_OK;
}

int nd_peer_free(void *ptr) {
    if (ptr == nullptr) return ND_OK;
    const cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) {
        set_error("nd_peer_free: %s", cudaGetErrorString(e));
        return ND_ERR_CUDA;
    }
    return ND_OK;
}

int nd_lift_finalize_peers(const void *const *acc_host, void *const *mean_host, void *const *cov_host,
                           void *const *flags_host, int world, int rank, uint32_t epoch, int n_views_total,
                           int channels, int64_t n_voxels, const float *alpha, int64_t *count, const void *acc_mc,
                           void *mean_mc, void *cov_mc, int max_ctas, int timeout_ms, int owner_rank, void *stream) {
    ND_REQUIRE(acc_host && mean_host && flags_host, ND_ERR_BAD_ARG, "nd_lift_finalize_peers: null pointer table");
    ND_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, ND_ERR_BAD_ARG,
               "nd_lift_finalize_peers: world %d / rank %d outside [1, %d]", world, rank, kMaxPeers);
    ND_REQUIRE(channels > 0 && n_voxels >= 0 && n_views_total > 0, ND_ERR_BAD_SHAPE, "nd_lift_finalize_peers: bad shape");
    ND_REQUIRE(epoch != 0u, ND_ERR_BAD_ARG, "nd_lift_finalize_peers: epoch 0 is the initial state of the flags");
    if (n_voxels == 0) return ND_OK;
    PeerArgs a{};
    bool vec = n_voxels % 4 == 0 && (alpha == nullptr || reinterpret_cast<uintptr_t>(alpha) % 16 == 0);
    for (int g = 0; g < world; ++g) {
        ND_REQUIRE(acc_host[g] && mean_host[g] && flags_host[g], ND_ERR_BAD_ARG, "nd_lift_finalize_peers: null segment of rank %d", g);
        a.s1[g] = static_cast<const float *>(acc_host[g]);
        a.mean[g] = static_cast<float *>(mean_host[g]);
        a.cov[g] = cov_host != nullptr ? static_cast<float *>(cov_host[g]) : nullptr;
        a.flags[g] = static_cast<uint32_t *>(flags_host[g]);
        vec = vec && reinterpret_cast<uintptr_t>(a.s1[g]) % 16 == 0 && reinterpret_cast<uintptr_t>(a.mean[g]) % 16 == 0 &&
              reinterpret_cast<uintptr_t>(a.cov[g]) % 16 == 0 && ((int64_t)channels * n_voxels) % 4 == 0;
    }
    ND_REQUIRE(owner_rank >= -1 && owner_rank < world, ND_ERR_BAD_ARG, "nd_lift_finalize_peers: owner rank %d outside [-1, %d)",
               owner_rank, world);
    ND_REQUIRE(owner_rank < 0 || acc_mc == nullptr, ND_ERR_BAD_ARG, "nd_lift_finalize_peers: the multicast transport stores to every rank");
    a.world = world;
    a.rank = rank;
    a.owner = owner_rank;
    a.epoch = epoch;
    a.timeout_ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 4000) * 1000000ull;
    a.n_views_total = n_views_total;
    a.channels = channels;
    // channel slice of this rank: contiguous, sizes differ by at most one (the split of distributed.view_shard)
    const int base = channels / world, rem = channels % world;
    a.c_begin = rank * base + (rank < rem ? rank : rem);
    a.c_end = a.c_begin + base + (rank < rem ? 1 : 0);
    a.n_vox = n_voxels;
    a.alpha = alpha;
    a.count = count;
    {
        const bool with_s2 = cov_host != nullptr;
        const int64_t cn = (int64_t)channels * n_voxels;
        for (int g = 0; g < world; ++g) {                       // [S1: C x N | S2: C x N | count: N], or [S1 | count]
            const float *blk = a.s1[g];
            a.s2[g] = with_s2 ? blk + cn : nullptr;
            a.cntp[g] = blk + (with_s2 ? 2 : 1) * cn;
        }
    }
    const bool mc = acc_mc != nullptr;
    if (mc) {
        ND_REQUIRE(mean_mc != nullptr && (cov_mc != nullptr) == (cov_host != nullptr), ND_ERR_BAD_ARG,
                   "nd_lift_finalize_peers: multicast needs the multicast address of every buffer in use");
        ND_REQUIRE(vec && (reinterpret_cast<uintptr_t>(acc_mc) | reinterpret_cast<uintptr_t>(mean_mc) |
                           reinterpret_cast<uintptr_t>(cov_mc)) % 16 == 0,
                   ND_ERR_BAD_ALIGNMENT, "nd_lift_finalize_peers: multicast needs 16-byte aligned rows (N %% 4 == 0)");
        a.acc_mc = static_cast<const float *>(acc_mc);
        a.mean_mc = static_cast<float *>(mean_mc);
        a.cov_mc = static_cast<float *>(cov_mc);
    }
    const int v = vec ? 4 : 1;
    const int slice = a.c_end - a.c_begin;
    const int gb = (mc || world <= 1) ? 1 : world <= 2 ? 2 : world <= 4 ? 4 : 8;   // compile-time bound of the instantiation
    const bool narrow = max_ctas > 0 && vec;
    // narrow: 512 threads x 16 vector loads in flight per thread = 128 KB per SM (a remote load takes ~2 us; with 4 loads per
    // thread 20 CTAs of 1024 threads took 164 us for what the wide grid does in 100 us: latency-bound at 32 GB/s per SM)
    const int threads = !narrow ? kPeerThreads : 512;
    const int64_t tiles = ceil_div(n_voxels, (int64_t)threads * v);
    // wide grid: enough CTAs to fill every SM at the instantiation's occupancy; a CTA keeps its tile's counts for its channels
    // measured on B200s (tools/dist_check.py sweep): 2 GPUs 102.9 us at 6 CTAs per SM (104-112 for 2-12); 8 GPUs 176.7 us at 2
    // (191 at 4, 237 at 12) -- the step is bound by the links (~520 GB/s inbound per GPU), more CTAs only add contention
    const int per_sm = gb <= 2 ? 6 : gb <= 4 ? 4 : 2;
    // narrow grid: ~8 work items per CTA so that the CTAs finish together
    int64_t subs = ceil_div(narrow ? (int64_t)max_ctas * 8 : (int64_t)148 * per_sm, tiles);
    if (subs > slice) subs = slice;
    if (subs < 1) subs = 1;
    a.ch_per_cta = slice > 0 ? (int)ceil_div(slice, subs) : 1;
    a.n_tiles = (int)tiles;
    a.n_items = (int)(tiles * (slice > 0 ? ceil_div(slice, a.ch_per_cta) : 1));
    const dim3 grid((unsigned)(narrow && max_ctas < a.n_items ? max_ctas : a.n_items));
    cudaStream_t st = (cudaStream_t)stream;
    const bool with_cov = cov_host != nullptr;
#define ND_PEER_LAUNCH(V_, G_, U_, T_)                                                       \
    do {                                                                                     \
        if (with_cov) k_lift_finalize_peers<V_, G_, U_, false, T_, true><<<grid, T_, 0, st>>>(a);  \
        else k_lift_finalize_peers<V_, G_, U_, false, T_, false><<<grid, T_, 0, st>>>(a);    \
    } while (0)
    if (mc) {
        ND_REQUIRE(with_cov, ND_ERR_BAD_ARG, "nd_lift_finalize_peers: the multicast transport carries mean and cov");
        if (narrow) k_lift_finalize_peers<4, 1, 4, true, 512, true><<<grid, 512, 0, st>>>(a);
        else k_lift_finalize_peers<4, 1, 4, true, kPeerThreads, true><<<grid, kPeerThreads, 0, st>>>(a);
    } else if (narrow) {
        if (gb == 1) ND_PEER_LAUNCH(4, 1, 4, 512);
        else if (gb == 2) ND_PEER_LAUNCH(4, 2, 4, 512);
        else if (gb == 4) ND_PEER_LAUNCH(4, 4, 2, 512);
        else ND_PEER_LAUNCH(4, 8, 1, 512);
    } else if (vec) {
        if (gb == 1) ND_PEER_LAUNCH(4, 1, 4, kPeerThreads);
        else if (gb == 2) ND_PEER_LAUNCH(4, 2, 2, kPeerThreads);
        else if (gb == 4) ND_PEER_LAUNCH(4, 4, 1, kPeerThreads);
        else ND_PEER_LAUNCH(4, 8, 1, kPeerThreads);
    } else {
        if (gb <= 2) ND_PEER_LAUNCH(1, 2, 2, kPeerThreads);
        else ND_PEER_LAUNCH(1, 8, 1, kPeerThreads);
    }
#undef ND_PEER_LAUNCH
    ND_CUDA_LAUNCH_CHECK("k_lift_finalize_peers");
    k_peer_wait_done<<<1, 32, 0, st>>>(a);
    ND_CUDA_LAUNCH_CHECK("k_peer_wait_done");
    return ND_OK;
}

}  // extern "C"
