// sampler.cu -- prioritized sampling, IS weights and priority write-back on
// the dense fp64 heaps.
//
// Replaces (reference, pure Python):
//   pfrl/collections/prioritized.py:56-116   _sample_indices_and_probabilities,
//                                            sample, set_last_priority
//   pfrl/collections/prioritized.py:245-258  _find
//   pfrl/collections/prioritized.py:294-312  SumTreeQueue.prioritized_sample
//   pfrl/replay_buffers/prioritized.py:47-66 priority_from_errors,
//                                            weights_from_probabilities
//
// EXACT mode reproduces the reference's sampled indices bit for bit: the
// draws are a dependent chain (draw k sees the root after the k-1 earlier
// hits were zeroed and their paths re-reduced), so one warp walks them in
// order.  What makes it fast is where the chain's operands live: the top
// TOP_LEVELS levels of the sum tree are staged in shared memory for the
// whole launch, and the remaining levels under the chosen top node are
// fetched by the 32 lanes in ONE round trip (each level's slice of a subtree
// is contiguous in the heap), so a draw costs one global latency, not one per
// level.
#include <math.h>
#include <stdlib.h>

#include "b2rl_internal.cuh"

#define TRY(x)                                                                 \
    do {                                                                       \
        int rc__ = (x);                                                        \
        if (rc__ != B2RL_OK) return rc__;                                      \
    } while (0)

static constexpr int TOP_LEVELS = 14; // heap levels 0..13 -> 16383 nodes = 128 KB

struct SampleArgs {
    double *sum;
    const double *mn;
    B2rlDevState *st;
    const double *u;
    int n;
    int levels;      // leaves are at heap level `levels`
    long long nslots;
    int T;           // levels held in shared memory (nodes [1, 2^T))
    int D;           // levels below the shared part (0: leaves are in shared)
    int32_t *slots_out;
    double *prio_out;
    long long *index_out; // optional
    double *prio_user;    // optional
};

extern __shared__ __align__(128) double smem_d[];

__global__ void __launch_bounds__(256, 1) k_sample_exact(SampleArgs a)
{
    double *top = smem_d;
    double *sub = smem_d + (1 << a.T);
    const int topn = 1 << a.T;
    const int tid = threadIdx.x;
    for (int i = 1 + tid; i < topn; i += blockDim.x) top[i] = a.sum[i];
    __syncthreads();

    if (tid < 32) {
        const int lane = tid;
        const long long mask = a.nslots - 1;
        const long long npop = a.st->npop;
        // The reference's root visits its OLDER half first
        // (collections/prioritized.py:255-258 with the bounds of :229-241).
        const int older = ((npop & mask) >= (a.nslots >> 1)) ? 3 : 2;
        if (lane == 0) {
            a.st->last_total = top[1]; // priority_sums.sum(), :58
            a.st->last_min = a.mn[1];  // priority_mins.min(), :59
            a.st->last_n = a.n;
        }
        double unext = a.n > 0 ? a.u[0] : 0.0;
        for (int k = 0; k < a.n; k++) {
            const double uk = unext;
            if (k + 1 < a.n) unext = a.u[k + 1];
            // np.random.uniform(0.0, root) = 0.0 + (root - 0.0) * u, :302
            double pos = __dmul_rn(top[1], uk);
            // _find, :245-258
            int node = older;
            {
                const double left = top[older];
                if (!(pos < left)) {
                    pos = __dsub_rn(pos, left);
                    node = older ^ 1;
                }
            }
            for (int lv = 1; lv < a.T - 1; lv++) {
                const double left = top[2 * node];
                if (pos < left) {
                    node = 2 * node;
                } else {
                    pos = __dsub_rn(pos, left);
                    node = 2 * node + 1;
                }
            }
            long long leafnode;
            double prio;
            if (a.D > 0) {
                // one round trip: all D levels under `node`
                for (int j = 1; j <= a.D; j++) {
                    const int cnt = 1 << j;
                    const double *src = a.sum + ((long long)node << j);
                    for (int i = lane; i < cnt; i += 32) sub[cnt + i] = src[i];
                }
                __syncwarp();
                int rel = 1;
                for (int j = 0; j < a.D; j++) {
                    const double left = sub[2 * rel];
                    if (pos < left) {
                        rel = 2 * rel;
                    } else {
                        pos = __dsub_rn(pos, left);
                        rel = 2 * rel + 1;
                    }
                }
                leafnode = ((long long)node << a.D) + (rel - (1 << a.D));
                prio = sub[rel];
                // _write(ix, 0.0): zero the leaf, re-reduce the path, :303
                sub[rel] = 0.0;
                if (lane == 0) a.sum[leafnode] = 0.0;
                int dj = a.D - 1;
                for (int p = rel >> 1; p >= 2; p >>= 1, dj--) {
                    const double v = __dadd_rn(sub[2 * p], sub[2 * p + 1]);
                    sub[p] = v;
                    if (lane == 0) a.sum[((long long)node << dj) + (p - (1 << dj))] = v;
                }
                top[node] = __dadd_rn(sub[2], sub[3]);
            } else {
                leafnode = node;
                prio = top[node];
                top[node] = 0.0;
            }
            for (int p = node >> 1; p >= 1; p >>= 1)
                top[p] = __dadd_rn(top[2 * p], top[2 * p + 1]);
            if (lane == 0) {
                const long long slot = leafnode - a.nslots;
                a.slots_out[k] = (int32_t)slot;
                a.prio_out[k] = prio;
                if (a.index_out) a.index_out[k] = (slot - npop) & mask;
                if (a.prio_user) a.prio_user[k] = prio;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // publish the shared-memory levels (zeroed state) back to HBM
    for (int i = 1 + tid; i < topn; i += blockDim.x) a.sum[i] = top[i];
}

// ---------------------------------------------------------------------------
// Latency-engineered EXACT sampler for deep trees (levels >= 13).
//
// The draws are a dependent chain, so throughput = 1 / (latency of one draw).
// Per draw the critical path is: descent (compare / subtract per level) ->
// leaf -> bottom-up re-reduction (one add per level) -> new root.  This
// version shortens it three ways:
//   * speculative descent: the 32 lanes evaluate all 2^R continuations of R
//     levels at once (lane i follows the path whose turn bits are i).  The R
//     `left` operands of every lane depend only on (node, lane), so they are
//     fetched together; the per-lane chain is R predicated subtractions with
//     the comparisons off the critical path; one ballot picks the lane whose
//     comparisons are all consistent.  5 levels cost ~1 shared-memory latency
//     + 5 DADD instead of 5 x (LDS + compare + select).
//   * the D levels below the shared-memory top are fetched as double2 child
//     pairs, all loads of a lane in flight together: one HBM/L2 round trip.
//   * the re-reduction loads all 13 + D siblings first (independent), then
//     runs the 13 + D dependent adds out of registers; stores are fire and
//     forget.
// The arithmetic (order of every compare, subtract and add) is exactly that
// of the reference, so indices stay bit-identical.
// ---------------------------------------------------------------------------
// L2 residency: the sum tree (32 MB at 1M capacity) is re-read by every
// draw while the gather streams ~140 MB per minibatch through the same L2.
// Tree accesses carry an evict_last policy (the gather's stores are .cs /
// evict-first), so the descent's one global round trip per draw tends to be an
// L2 hit instead of an HBM access.
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ double2 ld_tree_pair(const double2 *ptr, uint64_t pol)
{
    double2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(ptr), "l"(pol));
    return v;
}

__device__ __forceinline__ void st_tree(double *ptr, double v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(ptr), "d"(v), "l"(pol)
                 : "memory");
}

// FMA (default since round 2: bit-identity suite green on a B200, 741 vs 802 ns/draw;
// B2RL_SAMPLER_DESCENT=sub selects the subtract + select form): the
// conditional subtraction `if (right) x -= left` becomes ONE fused multiply-add with a
// lane-constant multiplier instead of a subtraction plus a 64-bit select (2 x FSEL) and
// the predicate traffic around it.  fma(-1, left, x) rounds x - left once, exactly like
// __dsub_rn; fma(-0, left, x) = x + (-0) = x because tree values are finite and >= 0.
// The arithmetic of the winning lane is therefore bit for bit the reference's.  Static
// SASS count of k_sample_exact_deep<8>: 1448 -> 1304 instructions (FSEL 114 -> 0,
// ISETP 100 -> 52), i.e. ~14 % of the issue-bound per-draw body.
template <int R, bool FMA>
__device__ __forceinline__ void spec_round(const double *val, int &node, double &pos, int lane)
{
    static_assert(R >= 1 && R <= 5, "one round covers at most 5 levels");
    const int li = lane & ((1 << R) - 1);
    double left[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int nj = (node << j) + (li >> (R - j));
        left[j] = val[2 * nj];
    }
    double x = pos;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < R; j++) {
        const bool right = (li >> (R - 1 - j)) & 1;
        const bool lt = x < left[j];
        ok = ok && (lt != right);
        if constexpr (FMA) {
            x = __fma_rn(right ? -1.0 : -0.0, left[j], x);
        } else {
            if (right) x = __dsub_rn(x, left[j]);
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, ok);
    if constexpr (R < 5) m &= (1u << (1 << R)) - 1u;
    const int win = __ffs(m) - 1;
    pos = __shfl_sync(0xffffffffu, x, win);
    node = (node << R) + win;
}

template <int L, bool FMA = false>
__device__ __forceinline__ void spec_descend(const double *val, int &node, double &pos, int lane)
{
    if constexpr (L > 0) {
        constexpr int R = L >= 5 ? 5 : L;
        spec_round<R, FMA>(val, node, pos, lane);
        spec_descend<L - R, FMA>(val, node, pos, lane);
    }
}

// ---- TMA-style bulk copies (cp.async.bulk, SASS UBLKCP) for the 128 KB top ----
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ int ld_acquire_smem(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_smem(int *p, int v)
{
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// Three warps per launch.
//   warp 0 (main)   walks the draws in order, exactly as the reference does.
//   warps 1, 2 (scouts) run a couple of draws ahead on the not-yet-final
//                   tree (alternating draws), predict which level-13 node a
//                   future draw will pick and stage that node's lower levels
//                   into shared memory, so that the main warp's only global
//                   round trip per draw is usually already done when it gets
//                   there (measured: ~99 % of the draws).
// The prediction is only a prefetch hint: the main warp uses the staged copy
// iff the predicted node equals the node its own exact descent reached AND no
// draw in flight when the copy was taken could have changed that subtree
// (release/acquire on `main_done` + "node != previous node"); otherwise it
// fetches itself.  The arithmetic of the main warp is unchanged, so indices
// stay bit-identical to the reference.
// Shared memory: [top 2^14 f64][sub_own 2^(D+1)][sub_pref 2 x 2^(D+1)][out
// staging][flags][mbarrier].  The top arrives / leaves by bulk async copy.
template <int D, bool FMA>
__global__ void __launch_bounds__(96, 1) k_sample_exact_deep(SampleArgs a)
{
    constexpr int T = TOP_LEVELS; // shared memory holds heap levels 0..T-1
    constexpr int TOPN = 1 << T;
    constexpr int SUBN = 2 << D;
    constexpr int PAIRS = (1 << D) - 1;       // child pairs below the chosen top node
    constexpr int NIT = (PAIRS + 31) / 32;
    constexpr int CHUNK = 32;                 // draws per output / u staging chunk
    constexpr int NSCOUT = 2;                 // scout warps (warp 1..NSCOUT), round-robin over draws
    constexpr int LAG = 2;                    // scout reads the tree as of draw k-1-LAG
    constexpr int NBUF = LAG + 1;             // staged subtrees alive at once
    constexpr int F_READY = 0, F_NODE = 4, F_DONE = 8;
    double *top = smem_d;
    double *sub_own = smem_d + TOPN;
    double *sub_pref = sub_own + SUBN;                            // [NBUF][SUBN]
    double *o_prio = sub_pref + NBUF * SUBN;                      // [CHUNK]
    int *o_slot = reinterpret_cast<int *>(o_prio + CHUNK);        // [CHUNK]
    int *flags = o_slot + CHUNK;  // ready_seq[NBUF] @0, pred_node[NBUF] @4, main_done @8
    uint64_t *bar = reinterpret_cast<uint64_t *>(flags + 16);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; i++) flags[i] = -1;
        flags[F_DONE] = 0;
        mbar_init(bar, 1);
        mbar_expect_tx(bar, TOPN * 8);
#pragma unroll
        for (int c = 0; c < 4; c++)
            bulk_g2s(top + c * (TOPN / 4), a.sum + c * (TOPN / 4), TOPN * 2, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const long long mask = a.nslots - 1;
    const long long npop = a.st->npop;
    const int older = ((npop & mask) >= (a.nslots >> 1)) ? 3 : 2;
    const uint64_t pol = policy_evict_last();
    const double2 *sum2 = reinterpret_cast<const double2 *>(a.sum);

    if (warp >= 1) {
        // ------------------------------ scouts -------------------------------
        // scout s stages draws k = 1 + s, 1 + s + NSCOUT, ...; each runs up to
        // LAG draws ahead of the last COMPLETED draw
        for (int k = warp; k < a.n; k += NSCOUT) {
            const double uk = a.u[k];
            // draws 0..k-1-LAG must be complete (their stores visible) before we read
            while (ld_acquire_smem(&flags[F_DONE]) < k - LAG) __nanosleep(64);
            double pos = uk * top[1]; // approximate: up to LAG draws still in flight
            int node = older;
            {
                const double left = top[older];
                if (!(pos < left)) { pos -= left; node = older ^ 1; }
            }
            spec_descend<T - 2>(top, node, pos, lane);
            double *dst = sub_pref + (k % NBUF) * SUBN;
            const unsigned unode = (unsigned)node;
            double2 tmp[NIT];
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                int q = lane + 32 * it;
                q = q < 1 ? 1 : (q > PAIRS ? PAIRS : q);
                const int dq = 31 - __clz(q);
                tmp[it] = ld_tree_pair(sum2 + ((unode << dq) + (unsigned)(q - (1 << dq))), pol);
            }
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                int q = lane + 32 * it;
                q = q < 1 ? 1 : (q > PAIRS ? PAIRS : q);
                reinterpret_cast<double2 *>(dst)[q] = tmp[it];
            }
            __syncwarp();
            if (lane == 0) {
                flags[F_NODE + (k % NBUF)] = node;
                st_release_smem(&flags[F_READY + (k % NBUF)], k);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------- main --------------------------------
        if (lane == 0) {
            a.st->last_total = top[1];
            a.st->last_min = a.mn[1];
            a.st->last_n = a.n;
        }
        int prev_node[LAG];
#pragma unroll
        for (int i = 0; i < LAG; i++) prev_node[i] = -1;
        int hits = 0, late = 0;
        for (int k0 = 0; k0 < a.n; k0 += CHUNK) {
            const double u_lane = (k0 + lane < a.n) ? a.u[k0 + lane] : 0.0;
            const int kend = (a.n - k0 < CHUNK) ? a.n - k0 : CHUNK;
            for (int kk = 0; kk < kend; kk++) {
                const int k = k0 + kk;
                const double uk = __shfl_sync(0xffffffffu, u_lane, kk);
                double pos = __dmul_rn(top[1], uk);
                int node = older;
                {
                    const double left = top[older];
                    if (!(pos < left)) {
                        pos = __dsub_rn(pos, left);
                        node = older ^ 1;
                    }
                }
                spec_descend<T - 2, FMA>(top, node, pos, lane); // level 1 -> T-1
                const unsigned unode = (unsigned)node;
                // ---- the D levels under `node`: staged by the scout, or fetched here
                const double *sub;
                const int ready_seq = ld_acquire_smem(&flags[F_READY + (k % NBUF)]);
                bool staged = ready_seq == k && flags[F_NODE + (k % NBUF)] == node;
                if (ready_seq != k) late++;
#pragma unroll
                for (int i = 0; i < LAG; i++) staged = staged && node != prev_node[i];
                if (staged) {
                    sub = sub_pref + (k % NBUF) * SUBN;
                    hits++;
                } else {
                    double2 tmp[NIT];
#pragma unroll
                    for (int it = 0; it < NIT; it++) {
                        int q = lane + 32 * it;
                        q = q < 1 ? 1 : (q > PAIRS ? PAIRS : q);
                        const int dq = 31 - __clz(q);
                        tmp[it] =
                            ld_tree_pair(sum2 + ((unode << dq) + (unsigned)(q - (1 << dq))), pol);
                    }
#pragma unroll
                    for (int it = 0; it < NIT; it++) {
                        int q = lane + 32 * it;
                        q = q < 1 ? 1 : (q > PAIRS ? PAIRS : q);
                        reinterpret_cast<double2 *>(sub_own)[q] = tmp[it];
                    }
                    __syncwarp();
                    sub = sub_own;
                }
                int rel = 1;
                spec_descend<D, FMA>(sub, rel, pos, lane);
                const unsigned leafnode = (unode << D) + (unsigned)(rel - (1 << D));
                const double prio = sub[rel];
                // ---- re-reduce the path: siblings first, then the add chain ---
                double sib[D + T - 1];
#pragma unroll
                for (int j = 0; j < D; j++) sib[j] = sub[(rel >> j) ^ 1];
#pragma unroll
                for (int j = 0; j < T - 1; j++) sib[D + j] = top[(node >> j) ^ 1];
                double v = 0.0;
                if (lane == 0) st_tree(a.sum + leafnode, 0.0, pol);
#pragma unroll
                for (int j = 0; j < D; j++) {
                    v = __dadd_rn(v, sib[j]);
                    if (lane == 0) {
                        if (j + 1 < D) {
                            const int dp = D - j - 1; // depth of the relative parent
                            const unsigned p = (unsigned)(rel >> (j + 1));
                            st_tree(a.sum + ((unode << dp) + (p - (1u << dp))), v, pol);
                        } else {
                            top[node] = v;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < T - 1; j++) {
                    v = __dadd_rn(v, sib[D + j]);
                    if (lane == 0) top[node >> (j + 1)] = v;
                }
                if (lane == 0) {
                    o_slot[kk] = (int)(leafnode - (unsigned)a.nslots);
                    o_prio[kk] = prio;
                    // publish "draws 0..k are complete" (global + shared stores above)
                    st_release_smem(&flags[F_DONE], k + 1);
                }
#pragma unroll
                for (int i = LAG - 1; i > 0; i--) prev_node[i] = prev_node[i - 1];
                prev_node[0] = node;
                __syncwarp();
            }
            if (lane < kend) {
                const int k = k0 + lane;
                const long long slot = o_slot[lane];
                const double prio = o_prio[lane];
                a.slots_out[k] = (int32_t)slot;
                a.prio_out[k] = prio;
                if (a.index_out) a.index_out[k] = (slot - npop) & mask;
                if (a.prio_user) a.prio_user[k] = prio;
            }
            __syncwarp();
        }
        if (lane == 0) a.st->pad = hits | (late << 16); // scout diagnostics: hits, not-ready
    }
    __syncthreads();
    // publish the shared-memory levels (zeroed state) back to HBM: bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < 4; c++)
            bulk_s2g(a.sum + c * (TOPN / 4), top + c * (TOPN / 4), TOPN * 2);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <int D, bool FMA>
static cudaError_t launch_deep_as(const SampleArgs &a, cudaStream_t s)
{
    const size_t smem = sizeof(double) * ((size_t(1) << TOP_LEVELS) + 4 * (size_t(2) << D) + 32) +
                        sizeof(int) * (32 + 16) + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_sample_exact_deep<D, FMA>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    k_sample_exact_deep<D, FMA><<<1, 96, smem, s>>>(a);
    return cudaGetLastError();
}

static bool fma_descent_requested()
{
    static int cached = -1;
    if (cached < 0) {
        const char *e = getenv("B2RL_SAMPLER_DESCENT");
        cached = (e && e[0] == 's') ? 0 : 1;
    }
    return cached == 1;
}

template <int D>
static cudaError_t launch_deep(const SampleArgs &a, cudaStream_t s)
{
    return fma_descent_requested() ? launch_deep_as<D, true>(a, s)
                                   : launch_deep_as<D, false>(a, s);
}

static cudaError_t launch_exact_deep(const SampleArgs &a, cudaStream_t s)
{
    switch (a.D) {
    case 1: return launch_deep<1>(a, s);
    case 2: return launch_deep<2>(a, s);
    case 3: return launch_deep<3>(a, s);
    case 4: return launch_deep<4>(a, s);
    case 5: return launch_deep<5>(a, s);
    case 6: return launch_deep<6>(a, s);
    case 7: return launch_deep<7>(a, s);
    case 8: return launch_deep<8>(a, s);
    case 9: return launch_deep<9>(a, s);
    case 10: return launch_deep<10>(a, s);
    default: return cudaErrorInvalidValue;
    }
}

// PARALLEL mode: every draw descends the frozen tree on its own thread.
__global__ void k_sample_parallel(SampleArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const long long mask = a.nslots - 1;
    const long long npop = a.st->npop;
    const double root = a.sum[1];
    if (k == 0) {
        a.st->last_total = root;
        a.st->last_min = a.mn[1];
        a.st->last_n = a.n;
    }
    if (k >= a.n) return;
    const int older = ((npop & mask) >= (a.nslots >> 1)) ? 3 : 2;
    double pos = __dmul_rn(root, a.u[k]);
    long long node = older;
    {
        const double left = a.sum[older];
        if (!(pos < left)) {
            pos = __dsub_rn(pos, left);
            node = older ^ 1;
        }
    }
    for (int lv = 1; lv < a.levels; lv++) {
        const double2 c = *reinterpret_cast<const double2 *>(a.sum + 2 * node);
        if (pos < c.x) {
            node = 2 * node;
        } else {
            pos = __dsub_rn(pos, c.x);
            node = 2 * node + 1;
        }
    }
    const double prio = a.sum[node];
    const long long slot = node - a.nslots;
    a.slots_out[k] = (int32_t)slot;
    a.prio_out[k] = prio;
    if (a.index_out) a.index_out[k] = (slot - npop) & mask;
    if (a.prio_user) a.prio_user[k] = prio;
}

extern "C" int b2rl_per_sample(b2rl_replay *h, const double *u_host, int32_t n, int mode,
                               int64_t *index_dev, double *priority_dev, void *stream)
{
    B2RL_REQUIRE(h && u_host, B2RL_ERR_INVALID, "null argument");
    B2RL_REQUIRE(h->cfg.prioritized, B2RL_ERR_INVALID, "buffer has no priority trees");
    B2RL_REQUIRE(n > 0 && n <= h->cfg.max_batch, B2RL_ERR_RANGE, "sample: n=%d out of 1..max_batch=%d",
                 n, h->cfg.max_batch);
    B2RL_REQUIRE(n <= h->napp - h->npop, B2RL_ERR_RANGE,
                 "sample: n=%d exceeds the %lld stored experiences", n,
                 (long long)(h->napp - h->npop));
    B2RL_REQUIRE(!h->wait_priority, B2RL_ERR_PROTOCOL,
                 "sample() called again before the previous sample's priorities were set "
                 "(collections/prioritized.py:98)");
    B2RL_REQUIRE(mode == B2RL_SAMPLE_EXACT || mode == B2RL_SAMPLE_PARALLEL, B2RL_ERR_INVALID,
                 "unknown sample mode %d", mode);
    cudaStream_t s = (cudaStream_t)stream;
    B2RL_CUDA(cudaSetDevice(h->cfg.device));
    for (int i = 0; i < n; i++)
        B2RL_REQUIRE(u_host[i] >= 0.0 && u_host[i] < 1.0, B2RL_ERR_INVALID,
                     "sample: u[%d]=%g not in [0,1)", i, u_host[i]);
    TRY(b2rl_stage_acquire(h, (size_t)n * 8));
    memcpy(h->pin, u_host, (size_t)n * 8);
    B2RL_CUDA(cudaMemcpyAsync(h->u_dev, h->pin, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    TRY(b2rl_stage_release(h, s));

    SampleArgs a;
    a.sum = h->sum;
    a.mn = h->mn;
    a.st = h->st;
    a.u = h->u_dev;
    a.n = n;
    a.levels = h->levels;
    a.nslots = h->nslots;
    a.T = h->levels + 1 < TOP_LEVELS ? h->levels + 1 : TOP_LEVELS;
    a.D = h->levels - (a.T - 1);
    a.slots_out = h->last_slots;
    a.prio_out = h->last_prio;
    a.index_out = (long long *)index_dev;
    a.prio_user = priority_dev;
    if (mode == B2RL_SAMPLE_EXACT && a.D > 0) {
        B2RL_CUDA(launch_exact_deep(a, s));
    } else if (mode == B2RL_SAMPLE_EXACT) {
        size_t smem = sizeof(double) * ((size_t(1) << a.T) + (size_t(2) << a.D));
        static bool attr_set = false;
        if (!attr_set) {
            B2RL_CUDA(cudaFuncSetAttribute(k_sample_exact,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024));
            attr_set = true;
        }
        k_sample_exact<<<1, 256, smem, s>>>(a);
    } else {
        k_sample_parallel<<<(n + 127) / 128, 128, 0, s>>>(a);
    }
    B2RL_CUDA(cudaGetLastError());
    h->wait_priority = true;
    h->last_n = n;
    h->last_mode = mode;
    return B2RL_OK;
}

// ---------------------------------------------------------------------------
// importance-sampling weights
// ---------------------------------------------------------------------------
struct WeightArgs {
    const B2rlDevState *st;
    const double *prio;
    int n;
    double beta;
    int norm;
    float *weight;
    double *prob;
};

__global__ void __launch_bounds__(1024) k_weights(WeightArgs a)
{
    __shared__ double red[32];
    __shared__ double s_min;
    const double total = a.st->last_total;
    double denom;
    if (a.norm == B2RL_NORM_BATCH) {
        // np.min(probabilities), replay_buffers/prioritized.py:60
        double m = INFINITY;
        for (int k = threadIdx.x; k < a.n; k += blockDim.x) m = fmin(m, a.prio[k] / total);
        for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x < 32) {
            m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : INFINITY;
            for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (threadIdx.x == 0) s_min = m;
        }
        __syncthreads();
        denom = s_min;
    } else if (a.norm == B2RL_NORM_MEMORY) {
        denom = a.st->last_min / total; // collections/prioritized.py:60
    } else {
        denom = 0.0;
    }
    const double len = (double)(a.st->napp - a.st->npop);
    for (int k = threadIdx.x; k < a.n; k += blockDim.x) {
        const double p = a.prio[k] / total; // :79-82 with uniform_ratio == 0
        if (a.prob) a.prob[k] = p;
        if (a.weight) {
            const double base = (a.norm == B2RL_NORM_NONE) ? len * p : p / denom;
            a.weight[k] = (float)pow(base, -a.beta); // :62 / :64
        }
    }
}

extern "C" int b2rl_per_weights(b2rl_replay *h, double beta, int norm, float *weight_dev,
                                double *prob_dev, void *stream)
{
    B2RL_REQUIRE(h, B2RL_ERR_INVALID, "null handle");
    B2RL_REQUIRE(h->wait_priority && h->last_n > 0, B2RL_ERR_PROTOCOL,
                 "weights requested without a pending sample");
    B2RL_REQUIRE(norm >= 0 && norm <= 2, B2RL_ERR_INVALID, "unknown normalisation %d", norm);
    B2RL_CUDA(cudaSetDevice(h->cfg.device));
    WeightArgs a{h->st, h->last_prio, h->last_n, beta, norm, weight_dev, prob_dev};
    k_weights<<<1, 1024, 0, (cudaStream_t)stream>>>(a);
    B2RL_CUDA(cudaGetLastError());
    return B2RL_OK;
}

// ---------------------------------------------------------------------------
// priority write-back
// ---------------------------------------------------------------------------
struct UpdateArgs {
    double *sum, *mn;
    B2rlDevState *st;
    const int32_t *slots;
    double *new_prio;       // [n] priorities (input, or scratch for the error form)
    const void *err;        // optional TD errors
    int err_is_f64;
    double alpha, eps, emin, emax;
    int32_t *winner;
    int n, levels;
    long long nslots;
};

__global__ void __launch_bounds__(1024) k_update(UpdateArgs a)
{
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    if (a.err) {
        // priority_from_errors, replay_buffers/prioritized.py:47-55
        for (int k = tid; k < a.n; k += nt) {
            double d = a.err_is_f64 ? ((const double *)a.err)[k] : (double)((const float *)a.err)[k];
            if (a.emin <= a.emax) d = fmin(fmax(d, a.emin), a.emax);
            d += a.eps;
            a.new_prio[k] = (a.alpha == 0.5) ? sqrt(d) : pow(d, a.alpha);
        }
    }
    // duplicates: the reference writes in order, so the LAST occurrence wins
    for (int k = tid; k < a.n; k += nt) atomicMax(&a.winner[a.slots[k]], k);
    __syncthreads();
    double mx = 0.0;
    for (int k = tid; k < a.n; k += nt) {
        const double p = a.new_prio[k];
        mx = fmax(mx, p); // max_priority sees every value, :114
        const long long leaf = a.nslots + a.slots[k];
        if (a.winner[a.slots[k]] == k) {
            a.sum[leaf] = p;
            a.mn[leaf] = p;
        }
    }
    __syncthreads();
    // lower levels: one node per updated leaf and level, block barrier per level.
    // The leaf index stays in a register (n <= 4 * blockDim), so a level costs
    // one round trip for the four child loads instead of two dependent ones.
    const int topl = a.levels < 10 ? a.levels : 10; // levels < topl are redone in shared memory
    if (a.n <= 4 * nt) {
        long long leaf[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = tid + j * nt;
            leaf[j] = k < a.n ? a.nslots + a.slots[k] : -1;
        }
        for (int lv = a.levels - 1; lv >= topl; lv--) {
            const int sh = a.levels - lv;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (leaf[j] < 0) continue;
                const long long node = leaf[j] >> sh;
                const double2 sc = *reinterpret_cast<const double2 *>(a.sum + 2 * node);
                const double2 mc = *reinterpret_cast<const double2 *>(a.mn + 2 * node);
                a.sum[node] = sc.x + sc.y;
                a.mn[node] = fmin(mc.x, mc.y);
            }
            __syncthreads();
        }
    } else {
        for (int lv = a.levels - 1; lv >= topl; lv--) {
            const int sh = a.levels - lv;
            for (int k = tid; k < a.n; k += nt) {
                const long long node = (a.nslots + a.slots[k]) >> sh;
                a.sum[node] = a.sum[2 * node] + a.sum[2 * node + 1];
                a.mn[node] = fmin(a.mn[2 * node], a.mn[2 * node + 1]);
            }
            __syncthreads();
        }
    }
    // top `topl` levels (<= 1023 nodes per tree): every node is a pure function
    // of its children, so recompute them all from level `topl` in shared memory
    // (cheap barriers) instead of ten more global round trips
    {
        __shared__ double s_sum[2048], s_min[2048];
        const int base = 1 << topl;
        for (int i = tid; i < base; i += nt) {
            s_sum[base + i] = a.sum[base + i];
            s_min[base + i] = a.mn[base + i];
        }
        __syncthreads();
        for (int lv = topl - 1; lv >= 0; lv--) {
            const int w = 1 << lv;
            for (int i = tid; i < w; i += nt) {
                const int node = w + i;
                s_sum[node] = s_sum[2 * node] + s_sum[2 * node + 1];
                s_min[node] = fmin(s_min[2 * node], s_min[2 * node + 1]);
            }
            __syncthreads();
        }
        for (int i = 1 + tid; i < base; i += nt) {
            a.sum[i] = s_sum[i];
            a.mn[i] = s_min[i];
        }
    }
    for (int k = tid; k < a.n; k += nt) a.winner[a.slots[k]] = -1;
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid < 32) {
        mx = tid < (nt >> 5) ? red[tid] : 0.0;
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (tid == 0) {
            if (mx > a.st->max_priority) a.st->max_priority = mx;
            a.st->last_n = 0;
        }
    }
}

static int launch_update(b2rl_replay *h, int32_t n, const void *err, int err_is_f64, double alpha,
                         double eps, double emin, double emax, cudaStream_t s)
{
    UpdateArgs a;
    a.sum = h->sum;
    a.mn = h->mn;
    a.st = h->st;
    a.slots = h->last_slots;
    a.new_prio = h->new_prio;
    a.err = err;
    a.err_is_f64 = err_is_f64;
    a.alpha = alpha;
    a.eps = eps;
    a.emin = emin;
    a.emax = emax;
    a.winner = h->winner;
    a.n = n;
    a.levels = h->levels;
    a.nslots = h->nslots;
    k_update<<<1, 1024, 0, s>>>(a);
    B2RL_CUDA(cudaGetLastError());
    h->wait_priority = false;
    h->last_n = 0;
    return B2RL_OK;
}

static int check_update(b2rl_replay *h, int32_t n)
{
    B2RL_REQUIRE(h, B2RL_ERR_INVALID, "null handle");
    B2RL_REQUIRE(h->wait_priority, B2RL_ERR_PROTOCOL,
                 "priorities set without a pending sample (collections/prioritized.py:108)");
    B2RL_REQUIRE(n == h->last_n, B2RL_ERR_RANGE,
                 "got %d priorities for a sample of %d (collections/prioritized.py:110)", n,
                 h->last_n);
    return B2RL_OK;
}

extern "C" int b2rl_per_update_priorities(b2rl_replay *h, const double *priority, int on_device,
                                          int32_t n, void *stream)
{
    TRY(check_update(h, n));
    B2RL_REQUIRE(priority, B2RL_ERR_INVALID, "null priorities");
    cudaStream_t s = (cudaStream_t)stream;
    B2RL_CUDA(cudaSetDevice(h->cfg.device));
    if (on_device) {
        B2RL_CUDA(cudaMemcpyAsync(h->new_prio, priority, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    } else {
        for (int i = 0; i < n; i++)
            B2RL_REQUIRE(priority[i] > 0.0, B2RL_ERR_INVALID,
                         "priority[%d]=%g must be > 0 (collections/prioritized.py:109)", i,
                         priority[i]);
        TRY(b2rl_stage_acquire(h, (size_t)n * 8));
        memcpy(h->pin, priority, (size_t)n * 8);
        B2RL_CUDA(cudaMemcpyAsync(h->new_prio, h->pin, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        TRY(b2rl_stage_release(h, s));
    }
    return launch_update(h, n, nullptr, 0, 0, 0, 0, 0, s);
}

extern "C" int b2rl_per_update_errors(b2rl_replay *h, const void *err_dev, int err_is_f64,
                                      int32_t n, double alpha, double eps, double error_min,
                                      double error_max, void *stream)
{
    TRY(check_update(h, n));
    B2RL_REQUIRE(err_dev, B2RL_ERR_INVALID, "null errors");
    B2RL_CUDA(cudaSetDevice(h->cfg.device));
    return launch_update(h, n, err_dev, err_is_f64, alpha, eps, error_min, error_max,
                         (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------
extern "C" int b2rl_per_get_info(b2rl_replay *h, b2rl_per_info *out, void *stream)
{
    B2RL_REQUIRE(h && out, B2RL_ERR_INVALID, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    B2RL_CUDA(cudaSetDevice(h->cfg.device));
    B2rlDevState st;
    B2RL_CUDA(cudaMemcpyAsync(&st, h->st, sizeof(st), cudaMemcpyDeviceToHost, s));
    double roots[2] = {0.0, INFINITY};
    if (h->cfg.prioritized) {
        B2RL_CUDA(cudaMemcpyAsync(&roots[0], h->sum + 1, 8, cudaMemcpyDeviceToHost, s));
        B2RL_CUDA(cudaMemcpyAsync(&roots[1], h->mn + 1, 8, cudaMemcpyDeviceToHost, s));
    }
    B2RL_CUDA(cudaStreamSynchronize(s));
    out->total = roots[0];
    out->min = roots[1];
    out->max_priority = st.max_priority;
    out->napp = st.napp;
    out->npop = st.npop;
    out->scout_hits = st.pad;
    out->reserved = 0;
    return B2RL_OK;
}

// Checkpoint restore: PrioritizedBuffer is pickled whole, max_priority included
// (pfrl/replay_buffers/replay_buffer.py:85-94, collections/prioritized.py:32).
extern "C" int b2rl_per_set_max_priority(b2rl_replay *h, double max_priority, void *stream)
{
    B2RL_REQUIRE(h, B2RL_ERR_INVALID, "null argument");
    B2RL_REQUIRE(h->cfg.prioritized, B2RL_ERR_INVALID, "buffer has no priority trees");
    B2RL_REQUIRE(max_priority > 0.0 && max_priority < INFINITY, B2RL_ERR_RANGE,
                 "set_max_priority: value must be positive and finite");
    cudaStream_t s = (cudaStream_t)stream;
    B2RL_CUDA(cudaSetDevice(h->cfg.device));
    B2RL_CUDA(cudaMemcpyAsync(&h->st->max_priority, &max_priority, sizeof(double),
                              cudaMemcpyHostToDevice, s));
    B2RL_CUDA(cudaStreamSynchronize(s)); // the source is a stack variable
    return B2RL_OK;
}

extern "C" int b2rl_per_read_priorities(b2rl_replay *h, int64_t first, int64_t n, double *out,
                                        void *stream)
{
    B2RL_REQUIRE(h && out, B2RL_ERR_INVALID, "null argument");
    B2RL_REQUIRE(h->cfg.prioritized, B2RL_ERR_INVALID, "buffer has no priority trees");
    B2RL_REQUIRE(first >= 0 && n >= 0 && first + n <= h->napp - h->npop, B2RL_ERR_RANGE,
                 "read_priorities: range out of bounds");
    cudaStream_t s = (cudaStream_t)stream;
    B2RL_CUDA(cudaSetDevice(h->cfg.device));
    const int64_t mask = h->nslots - 1;
    int64_t done = 0;
    while (done < n) {
        int64_t slot = (h->npop + first + done) & mask;
        int64_t run = n - done < h->nslots - slot ? n - done : h->nslots - slot;
        B2RL_CUDA(cudaMemcpyAsync(out + done, h->sum + h->nslots + slot, (size_t)run * 8,
                                  cudaMemcpyDeviceToHost, s));
        done += run;
    }
    B2RL_CUDA(cudaStreamSynchronize(s));
    return B2RL_OK;
}
