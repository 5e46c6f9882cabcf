// gemm.cu -- K10: the dense layers of the Q-networks on the 5th-generation tensor cores.
//
//   C[M,N] = A[M,K] . B[N,K]^T (+ bias[N]) (relu)          fp32 in, fp32 out
//
// Replaces the cuBLAS SGEMMs behind (reference, all fp32 on the CUDA cores):
//   pfrl/q_functions/dueling_dqn.py:67-129     main_stream / a_stream / v_stream Linear
//   pfrl/nn/noisy_linear.py:53-70              F.linear with the noisy weight
//   pfrl/nn/atari_cnn.py:17-47                 the Linear(3136, 512) head of the Nature nets
// and their two backward products (dX = dY . W, dW = dY^T . X).
//
// Arithmetic: every fp32 operand x is split into two TF32 numbers, x = hi + lo with
// hi = rn_tf32(x), lo = rn_tf32(x - hi), and the product is accumulated as
// lo.hi + hi.lo + hi.hi in the fp32 accumulator of `tcgen05.mma.kind::tf32` (the lo.lo term,
// <= 2^-22 relative, is dropped).  That keeps the result within a few fp32 ulps of an exact
// fp32 product -- tests/test_gemm_gpu.py measures it against an fp64 product next to
// cuBLAS' own fp32 error -- which is what lets the agents' 1e-5 loss parity stand.
//
// Data movement: no TMA tensor maps.  The split has to touch every element anyway, so sixteen
// loader warps read the fp32 tiles through the read-only path, split them in registers and
// store hi and lo straight into the canonical no-swizzle K-major core-matrix layout the tensor
// core reads (8 rows x 16 B per core matrix, K-neighbours padded to 144 B so that every
// quarter warp stores to 8 different bank groups).  Operands whose contraction index is the
// contiguous one ("K-major": X and W in the forward product) are read with 16-byte loads, four
// whole 128-byte rows per warp instruction;
// operands whose contraction index is the ROW index ("MN-major": dY and X in dW = dY^T . X,
// W in dX = dY . W) are read with four 4-byte loads per thread, each warp instruction one
// coalesced 128-byte run along M/N, and the four k values meet in one 16-byte store -- the
// transposition happens in registers, no transposed copy is ever made and the tensor core only
// ever sees K-major tiles (kind::tf32 accepts MN-major tiles only in a 32-bit swizzled layout).
// The loads of k blocks i+1 and i+2 are in flight while block i is split and stored, so the
// global-load latency hides behind the conversion.
//
// Roles in a CTA of 17 warps, one 128 x 128 output tile (x one K split) per CTA:
//   warp  0     allocates 256 TMEM columns; lane 0 issues 3 MMAs per 8-wide k step (hi.hi into
//               one accumulator, lo.hi + hi.lo into the other)
//   warps 1-16  loaders, 3-stage ring of (A_hi, A_lo, B_hi, B_lo) = 72 KB per stage,
//               full/empty mbarriers; the empty side is armed by tcgen05.commit
//   warps 1-16  then the epilogue: tcgen05.ld the two accumulators (TMEM lane = tile row), add,
//               bias / relu, store (four warps per TMEM lane quarter, interleaved column chunks)
//               (dense outputs are staged through the idle operand ring and leave in whole
//               rows; NCHW scatter outputs are already coalesced across the lanes)
// Small products are split along K so that the grid covers the 148 SMs; the partial tiles go
// to a workspace and `k_gemm_reduce` adds them in a fixed order (deterministic), with the
// bias / relu epilogue.
//
// Bound (measured): every tile streams its own 32 KB of operands per k
// block from L2 -- 4.1-4.6 TB/s over all SMs, the L2 -> SM limit; the tensor pipe is 20-25 %
// active.  Sharing operands across a CTA pair / cluster is the known next step.
//
// Convolutions (pfrl/nn/atari_cnn.py:30-44, pfrl/q_functions/dueling_dqn.py:34-40,91-97: the
// 4x4/2 and 3x3/1 layers of the Nature trunk, forward, input gradient, weight gradient) are the
// same product with one more operand mode, "gather": element (row, k) lives at
//   row_off[row] + k_off[k]
// and, when the operand carries coordinates, exists iff 0 <= y[row] + dy[k] < y_limit and
// 0 <= x[row] + dx[k] < x_limit -- im2col, its transpose (col2im as a gather; a strided layer
// is one product per stride phase) and the pixel-major views of the weight gradient, all read
// in place through small int32 tables per operand (ops/conv.py builds them once per layer
// shape).  The output can be scattered the same way (C[m, n] at row_off[m] + n * col_stride:
// NCHW activations and transposed weight gradients), so no im2col buffer and no layout pass
// ever touches HBM.  uint8 sources
// (float(byte) * scale: the x / 255 of the Atari phi) are read directly.
#include "b2rl_internal.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32; // BK floats = 128 bytes of contraction per stage
constexpr int STAGES = 3;
constexpr int MMA_WARP = 0;
constexpr int LOAD_WARPS = 16;
constexpr int THREADS = (1 + LOAD_WARPS) * 32; // 544
// Shared-memory tile of one operand half (hi or lo): 16 row groups x 8 k chunks of 8 x 16 B core
// matrices.  The cores that are neighbours along K sit 144 bytes apart (128 + 16 of padding),
// so that the 8 lanes which hold one row's eight k chunks store to 8 different bank groups.
constexpr int CORE_K_BYTES = 144;                          // "leading" byte offset
constexpr int CORE_MN_BYTES = 8 * CORE_K_BYTES;            // "stride" byte offset (8-row groups)
constexpr int TILE_BYTES = (BM / 8) * CORE_MN_BYTES;       // 18 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;                // A_hi A_lo B_hi B_lo
constexpr int TMEM_COLS = 256; // two 128-column accumulators: hi.hi and the correction terms
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
constexpr long long WAIT_LIMIT = 4000000000ll; // cycles (~2 s): a lost arrival traps, never hangs

enum { MODE_K_MAJOR = 0, MODE_MN_MAJOR = 1, MODE_GATHER = 2 };

struct Operand {
    const void *p; // float, or uint8_t when u8
    int mode;
    int ld, vec;   // dense modes: leading dimension, 16-byte loads allowed
    // gather mode (see the header comment); k tables are padded to a multiple of 32 entries
    const int *row_off, *row_yx; // element offset; y | x << 16 (NULL: no coordinates)
    const int *k_off, *k_yx;     // element offset; dy | dx << 16, int16 each
    int y_limit, x_limit;
    int along_k;         // consecutive lanes walk k (else rows)
    int u8;
    float scale;
};

struct GemmArgs {
    Operand A, B;
    const float *bias;
    float *C;            // final output (splits == 1) ...
    float *partial;      // ... or [splits][M][N] partial sums
    const int *c_row_tab; // scatter: C[m, n] at c_row_tab[m] + n * c_col_stride
    int c_col_stride;
    int M, N, K;
    int ldc, c_vec;
    int bn;              // N tile: 32, 64 or 128
    int splits, kb_per_split;
    int relu;
    unsigned long long *times; // debug (B2RL_GEMM_TIMES): 8 %globaltimer stamps per CTA, or NULL
};

__device__ __forceinline__ void stamp(const GemmArgs &g, int slot)
{
    if (g.times) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g.times[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{ .reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return done != 0;
}

// try_wait suspends the thread in hardware for a while by itself; the clock is read only every
// 256 polls, to turn a lost arrival into a trap instead of a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    for (uint32_t spins = 1;; spins++) {
        if (mbar_try(bar, parity)) return;
        if ((spins & 255u) == 0 && clock64() - t0 > WAIT_LIMIT) __trap();
    }
}

// 64-bit shared-memory matrix descriptor, no swizzle ("interleave"): start address, the byte
// offset between the core matrices that are neighbours along K (leading) and along M/N
// (stride), both in 16-byte units; bits 46-47 = descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lead_bytes,
                                              uint32_t stride_bytes)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lead_bytes >> 4) << 16) |
           ((uint64_t)(stride_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                         uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :
                 : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                 : "memory");
}

__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                     "r"(bar)
                 : "memory");
}

// 16 consecutive accumulator columns of this thread's TMEM lane (no wait: see tmem_wait)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
                   "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void tmem_wait()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// x = hi + lo (+ <= 2^-22 |x|), both exactly representable in TF32 (10-bit mantissa, low 13
// bits zero); rounding to nearest in integer arithmetic on the sign-magnitude bit pattern
// (plain integer instructions instead of the conversion pipe).  x - hi is exact.  Non-finite
// inputs come out as NaN.
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}

// Four consecutive floats from p (n_valid of them exist), zeros elsewhere.
__device__ __forceinline__ float4 load4(const float *p, bool row_ok, int n_valid, bool vec)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!row_ok || n_valid <= 0) return v;
    if (vec && n_valid >= 4) return __ldg(reinterpret_cast<const float4 *>(p));
    v.x = __ldg(p);
    if (n_valid > 1) v.y = __ldg(p + 1);
    if (n_valid > 2) v.z = __ldg(p + 2);
    if (n_valid > 3) v.w = __ldg(p + 3);
    return v;
}

// One operand tile of one stage: 128 (M or N) x 32 (K) floats, stored K-major: core matrix
// (rg, kc) = rows 8 rg .. 8 rg + 7, k 4 kc .. 4 kc + 3.  Each loader thread owns two 16-byte
// vectors (4 consecutive k of one row) per operand and stage, in one of two lane arrangements:
//   along k   : unit u = rows 4u .. 4u+3 x all 32 k; lanes 8r .. 8r+7 hold row r (a whole
//               128-byte row per 8 lanes: 4 L1 wavefronts per 16-byte load instruction).
//   along rows: unit u = 32 rows x 4 k; lane = row (each warp load instruction is one coalesced
//               run along M/N); the four k values of a thread meet in its 16-byte store -- this
//               is the transposition of the MN-major operands.
struct Frag {
    float4 v[2];
};

template <int MODE> struct Lane { // per thread and operand, fixed for the whole tile
    const float *q[2]; // dense modes: address of this thread's element of k block 0
    int roff[2], ryx[2]; // gather: row table entries
    int tko[3], tkc[3];  // gather: k table entries of three k blocks in flight, lane = k mod 32
    int kofs[2];       // k offset of the thread's vector inside a k block
    int soff[2];       // byte offset of its 16-byte slot inside a tile half, -1: nothing to do
    bool along_k;
};

template <int MODE>
__device__ __forceinline__ void lane_setup(const Operand &o, int mn0, int mn_total, int tile_rows,
                                           int k_first, int lw, int lane, Lane<MODE> &L)
{
    L.along_k = MODE == MODE_K_MAJOR || (MODE == MODE_GATHER && o.along_k);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int u = lw + 16 * i; // 32 units per tile
        int r, kc;
        if (L.along_k) {
            r = 4 * u + (lane >> 3), kc = lane & 7;
        } else {
            r = (u & 3) * 32 + lane, kc = u >> 2;
        }
        L.kofs[i] = kc * 4;
        // rows past the N tile are never read by the tensor core; rows past the matrix are zero
        L.soff[i] = r < tile_rows ? (r >> 3) * CORE_MN_BYTES + kc * CORE_K_BYTES + (r & 7) * 16
                                  : -1;
        const bool ok = r < tile_rows && mn0 + r < mn_total;
        L.q[i] = nullptr;
        L.roff[i] = L.ryx[i] = 0;
        if (!ok) continue;
        if (MODE == MODE_K_MAJOR)
            L.q[i] = (const float *)o.p + (size_t)(mn0 + r) * o.ld + k_first + kc * 4;
        else if (MODE == MODE_MN_MAJOR)
            L.q[i] = (const float *)o.p + (size_t)(k_first + kc * 4) * o.ld + mn0 + r;
        else {
            L.q[i] = (const float *)o.p; // (marks the row as present)
            L.roff[i] = __ldg(o.row_off + mn0 + r);
            if (o.row_yx) L.ryx[i] = __ldg(o.row_yx + mn0 + r);
        }
    }
}

__device__ __forceinline__ float gather_load(const Operand &o, int idx)
{
    if (o.u8) return (float)__ldg((const uint8_t *)o.p + idx) * o.scale;
    return __ldg((const float *)o.p + idx);
}

// k table entries of one k block -> one register per lane (coalesced; the tables are padded to
// a multiple of 32).  Issued one iteration before the data loads that need them, so that the
// gather is not a chain of two dependent global loads.
template <int MODE>
__device__ __forceinline__ void table_fetch(const Operand &o, Lane<MODE> &L, int slot, int k0,
                                            int lane)
{
    if (MODE != MODE_GATHER) return;
    L.tko[slot] = __ldg(o.k_off + k0 + lane);
    L.tkc[slot] = o.k_yx ? __ldg(o.k_yx + k0 + lane) : 0;
}

// Four consecutive k (block-relative kofs .. kofs + 3) of one row; `n` = how many of them lie
// inside K.  All lanes take part in the shuffles, `present` says whether this lane has a row.
__device__ __forceinline__ float4 gather4(const Operand &o, bool present, int roff, int ryx,
                                          int tko, int tkc, int kofs, int n)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int4 ko;
    ko.x = __shfl_sync(0xffffffffu, tko, kofs);
    ko.y = __shfl_sync(0xffffffffu, tko, kofs + 1);
    ko.z = __shfl_sync(0xffffffffu, tko, kofs + 2);
    ko.w = __shfl_sync(0xffffffffu, tko, kofs + 3);
    if (o.row_yx == nullptr) {
        if (!present || n <= 0) return v;
        v.x = gather_load(o, roff + ko.x);
        if (n > 1) v.y = gather_load(o, roff + ko.y);
        if (n > 2) v.z = gather_load(o, roff + ko.z);
        if (n > 3) v.w = gather_load(o, roff + ko.w);
        return v;
    }
    int4 kc;
    kc.x = __shfl_sync(0xffffffffu, tkc, kofs);
    kc.y = __shfl_sync(0xffffffffu, tkc, kofs + 1);
    kc.z = __shfl_sync(0xffffffffu, tkc, kofs + 2);
    kc.w = __shfl_sync(0xffffffffu, tkc, kofs + 3);
    if (!present || n <= 0) return v;
    const int ry = ryx & 0xffff, rx = ryx >> 16;
    const unsigned yl = (unsigned)o.y_limit, xl = (unsigned)o.x_limit;
#define B2RL_OK_AT(c) \
    ((unsigned)(ry + (int)(short)((c) & 0xffff)) < yl && (unsigned)(rx + ((c) >> 16)) < xl)
    if (B2RL_OK_AT(kc.x)) v.x = gather_load(o, roff + ko.x);
    if (n > 1 && B2RL_OK_AT(kc.y)) v.y = gather_load(o, roff + ko.y);
    if (n > 2 && B2RL_OK_AT(kc.z)) v.z = gather_load(o, roff + ko.z);
    if (n > 3 && B2RL_OK_AT(kc.w)) v.w = gather_load(o, roff + ko.w);
#undef B2RL_OK_AT
    return v;
}

// k block `kb` (relative to the first one of this CTA) -> registers.  `full`: the whole block
// lies inside K (warp-uniform), which is the common case and needs no per-element checks.
template <int MODE>
__device__ __forceinline__ void fetch_operand(const Operand &o, const Lane<MODE> &L, int kb,
                                              int slot, int k0, int k_total, Frag &f)
{
    const bool full = k0 + BK <= k_total;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        if (MODE == MODE_GATHER) {
            f.v[i] = gather4(o, L.q[i] != nullptr, L.roff[i], L.ryx[i], L.tko[slot], L.tkc[slot],
                             L.kofs[i], full ? 4 : k_total - (k0 + L.kofs[i]));
            continue;
        }
        f.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (L.q[i] == nullptr) continue;
        if (MODE == MODE_K_MAJOR) {
            const float *q = L.q[i] + (size_t)kb * BK;
            if (full && o.vec)
                f.v[i] = __ldg(reinterpret_cast<const float4 *>(q));
            else
                f.v[i] = load4(q, true, k_total - (k0 + L.kofs[i]), o.vec);
        } else {
            const float *q = L.q[i] + (size_t)kb * BK * o.ld;
            const int left = k_total - (k0 + L.kofs[i]);
            if (full || left > 0) f.v[i].x = __ldg(q);
            if (full || left > 1) f.v[i].y = __ldg(q + o.ld);
            if (full || left > 2) f.v[i].z = __ldg(q + 2 * (size_t)o.ld);
            if (full || left > 3) f.v[i].w = __ldg(q + 3 * (size_t)o.ld);
        }
    }
}

template <int MODE>
__device__ __forceinline__ void store_operand(const Lane<MODE> &L, const Frag &f, uint8_t *s_hi,
                                              uint8_t *s_lo)
{
#pragma unroll
    for (int i = 0; i < 2; i++) {
        if (L.soff[i] < 0) continue;
        uint4 h, l;
        split_tf32(f.v[i].x, h.x, l.x);
        split_tf32(f.v[i].y, h.y, l.y);
        split_tf32(f.v[i].z, h.z, l.z);
        split_tf32(f.v[i].w, h.w, l.w);
        *reinterpret_cast<uint4 *>(s_hi + L.soff[i]) = h;
        *reinterpret_cast<uint4 *>(s_lo + L.soff[i]) = l;
    }
}

template <int AMODE, int BMODE>
__global__ void __launch_bounds__(THREADS, 1) k_gemm_tf32x3(const __grid_constant__ GemmArgs g)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (g.N + g.bn - 1) / g.bn;
    const int tile = blockIdx.x / g.splits, split = blockIdx.x % g.splits;
    const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * g.bn;
    const int kb_total = (g.K + BK - 1) / BK;
    const int kb0 = split * g.kb_per_split;
    const int kb1 = min(kb_total, kb0 + g.kb_per_split);
    const int n_kb = kb1 - kb0; // >= 1 by construction of the launch

    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
    const uint32_t accum_bar = smem_u32(&bars[2 * STAGES]);

    if (threadIdx.x == 0) stamp(g, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full0 + 8 * s, LOAD_WARPS);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (threadIdx.x == 0) stamp(g, 1); // set-up done (barriers, TMEM)

    if (warp > MMA_WARP) {
        // ---------------- loaders ----------------
        const int lw = warp - MMA_WARP - 1;
        Lane<AMODE> la;
        Lane<BMODE> lb;
        lane_setup<AMODE>(g.A, m0, g.M, BM, kb0 * BK, lw, lane, la);
        lane_setup<BMODE>(g.B, n0, g.N, g.bn, kb0 * BK, lw, lane, lb);
        // Two k blocks are in flight in registers while a third is split and stored: the
        // global-load latency (DRAM-cold operands: ~1.5 us) is spread over two iterations.
        // The ring of three register fragments lines up with the three smem stages.
        static_assert(STAGES == 3, "the loader loop is unrolled over the stage ring");
        Frag fa[3], fb[3];
#pragma unroll
        for (int j = 0; j < 3; j++)
            if (j < n_kb) {
                table_fetch<AMODE>(g.A, la, j, (kb0 + j) * BK, lane);
                table_fetch<BMODE>(g.B, lb, j, (kb0 + j) * BK, lane);
            }
#pragma unroll
        for (int j = 0; j < 2; j++)
            if (j < n_kb) {
                fetch_operand<AMODE>(g.A, la, j, j, (kb0 + j) * BK, g.K, fa[j]);
                fetch_operand<BMODE>(g.B, lb, j, j, (kb0 + j) * BK, g.K, fb[j]);
            }
        for (int it0 = 0; it0 < n_kb; it0 += 3) {
#pragma unroll
            for (int s = 0; s < 3; s++) {
                const int it = it0 + s;
                if (it >= n_kb) break;
                const uint32_t ph = (it / STAGES) & 1;
                if (it + 2 < n_kb) {
                    const int k2 = (kb0 + it + 2) * BK;
                    fetch_operand<AMODE>(g.A, la, it + 2, (s + 2) % 3, k2, g.K, fa[(s + 2) % 3]);
                    fetch_operand<BMODE>(g.B, lb, it + 2, (s + 2) % 3, k2, g.K, fb[(s + 2) % 3]);
                }
                if (it + 3 < n_kb) { // table slot s belonged to block `it`, whose loads are out
                    table_fetch<AMODE>(g.A, la, s, (kb0 + it + 3) * BK, lane);
                    table_fetch<BMODE>(g.B, lb, s, (kb0 + it + 3) * BK, lane);
                }
                if (lane == 0) mbar_wait(empty0 + 8 * s, ph ^ 1);
                __syncwarp();
                uint8_t *st = smem + s * STAGE_BYTES;
                store_operand<AMODE>(la, fa[s], st, st + TILE_BYTES);
                store_operand<BMODE>(lb, fb[s], st + 2 * TILE_BYTES, st + 3 * TILE_BYTES);
                // generic-proxy stores -> visible to the tensor core (async proxy)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(full0 + 8 * s);
            }
        }
    } else if (warp == MMA_WARP) {
        // ---------------- MMA issue: one thread ----------------
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = tf32, majors, N >> 3, M >> 4
            // (both operands K-major in shared memory: the loaders transpose)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) |
                                   ((uint32_t)(g.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            // neighbours along K: 144 B; 8-row groups: 1152 B; one k step = two k chunks
            const uint32_t a_lead = CORE_K_BYTES, a_stride = CORE_MN_BYTES;
            const uint32_t b_lead = CORE_K_BYTES, b_stride = CORE_MN_BYTES;
            const uint32_t a_step = 2 * CORE_K_BYTES, b_step = 2 * CORE_K_BYTES;
            for (int it = 0; it < n_kb; it++) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(full0 + 8 * s, ph);
                if (it == 0) stamp(g, 2); // first stage filled
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t sb = sa + 2 * TILE_BYTES;
#pragma unroll
                for (int j = 0; j < BK / 8; j++) {
                    const uint64_t a_hi = make_desc(sa + j * a_step, a_lead, a_stride);
                    const uint64_t a_lo = make_desc(sa + TILE_BYTES + j * a_step, a_lead, a_stride);
                    const uint64_t b_hi = make_desc(sb + j * b_step, b_lead, b_stride);
                    const uint64_t b_lo = make_desc(sb + TILE_BYTES + j * b_step, b_lead, b_stride);
                    // the two correction products have their own accumulator: the tensor core
                    // truncates on every accumulation, so the long hi.hi chain should not carry
                    // the small terms (and vice versa); the epilogue adds the two
                    mma_tf32(tmem_base + BN, a_lo, b_hi, idesc, (it | j) != 0);
                    mma_tf32(tmem_base + BN, a_hi, b_lo, idesc, 1);
                    mma_tf32(tmem_base, a_hi, b_hi, idesc, (it | j) != 0);
                }
                mma_commit(empty0 + 8 * s); // arrives when the MMAs above have read the stage
            }
            mma_commit(accum_bar);
            stamp(g, 3); // last MMA issued
        }
        __syncwarp();
    }
    if (warp >= 1) {
        // ---------------- epilogue: all sixteen loader warps, once their loads are done.  A
        // warp may touch the TMEM lanes 32 (warp % 4) .. 32 (warp % 4) + 31; the four warps
        // that share a quarter take every fourth 16-column chunk each ----------------
        const int quarter = warp & 3, part = (warp - 1) >> 2;
        if (lane == 0) mbar_wait(accum_bar, 0);
        __syncwarp();
        if (warp == 1 && lane == 0) stamp(g, 4); // accumulators complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + quarter * 32 + lane;
        const bool direct = g.splits == 1;
        const bool scatter = direct && g.c_row_tab != nullptr;
        float *out = direct ? g.C : g.partial + (size_t)split * g.M * g.N;
        const int ldo = direct ? g.ldc : g.N;
        const bool vec = direct ? g.c_vec : (g.N % 4 == 0);
        const long long obase = row < g.M ? (scatter ? (long long)__ldg(g.c_row_tab + row)
                                                     : (long long)row * ldo)
                                          : 0;
        const int ostride = scatter ? g.c_col_stride : 1;
        // Dense outputs go through shared memory (the operand ring is idle by now): a thread
        // owns a ROW of the accumulator, so storing it directly would write 16-byte pieces of
        // 32 different rows per instruction (half-empty sectors, 4-byte pieces when C is not
        // 16-byte aligned).  Staged as [128][bn + 4] floats, the tile leaves in whole rows.
        // Scattered outputs (NCHW activations) are already coalesced across the lanes.
        constexpr int SROW = BN + 4;
        float *stage_tile = reinterpret_cast<float *>(smem);
#pragma unroll 1
        for (int c = part; c < g.bn / 16; c += 4) {
            uint32_t r[16], q[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + c * 16;
            tmem_ld16(taddr, r);
            tmem_ld16(taddr + BN, q);
            tmem_wait();
            const int col0 = n0 + c * 16;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]) + __uint_as_float(q[j]);
            if (!scatter) {
                float *o = stage_tile + (quarter * 32 + lane) * SROW + c * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else if (row < g.M && col0 < g.N) {
                float *o = out + obase + (long long)col0 * ostride;
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (col0 + j < g.N) {
                        float x = v[j];
                        if (g.bias) x += __ldg(g.bias + col0 + j);
                        if (g.relu) x = fmaxf(x, 0.f);
                        o[(long long)j * ostride] = x; // lane = row: one run per column
                    }
            }
        }
        if (!scatter) {
            asm volatile("bar.sync 1, %0;" ::"n"(LOAD_WARPS * 32) : "memory");
            const int t = threadIdx.x - 32; // 0 .. 511
            const bool bias_relu = direct;
            if (vec && (n0 & 3) == 0) {
                const int per_row = g.bn / 4;
                for (int idx = t; idx < BM * per_row; idx += LOAD_WARPS * 32) {
                    const int rr = idx / per_row, cc = (idx % per_row) * 4;
                    if (m0 + rr >= g.M || n0 + cc >= g.N) continue;
                    float4 x = *reinterpret_cast<const float4 *>(stage_tile + rr * SROW + cc);
                    float *o = out + (size_t)(m0 + rr) * ldo + n0 + cc;
                    if (n0 + cc + 4 <= g.N) {
                        if (bias_relu && g.bias) {
                            const float4 bv = __ldg(reinterpret_cast<const float4 *>(g.bias + n0 + cc));
                            x.x += bv.x, x.y += bv.y, x.z += bv.z, x.w += bv.w;
                        }
                        if (bias_relu && g.relu)
                            x.x = fmaxf(x.x, 0.f), x.y = fmaxf(x.y, 0.f), x.z = fmaxf(x.z, 0.f),
                            x.w = fmaxf(x.w, 0.f);
                        *reinterpret_cast<float4 *>(o) = x;
                    } else {
                        const float xs[4] = {x.x, x.y, x.z, x.w};
                        for (int j = 0; j < 4 && n0 + cc + j < g.N; j++) {
                            float y = xs[j];
                            if (bias_relu && g.bias) y += __ldg(g.bias + n0 + cc + j);
                            if (bias_relu && g.relu) y = fmaxf(y, 0.f);
                            o[j] = y;
                        }
                    }
                }
            } else {
                for (int idx = t; idx < BM * g.bn; idx += LOAD_WARPS * 32) {
                    const int rr = idx / g.bn, cc = idx % g.bn;
                    if (m0 + rr >= g.M || n0 + cc >= g.N) continue;
                    float y = stage_tile[rr * SROW + cc];
                    if (bias_relu && g.bias) y += __ldg(g.bias + n0 + cc);
                    if (bias_relu && g.relu) y = fmaxf(y, 0.f);
                    out[(size_t)(m0 + rr) * ldo + n0 + cc] = y;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (warp == 1 && lane == 0) stamp(g, 5); // epilogue stores issued
    }
    __syncthreads();
    if (threadIdx.x == 0) stamp(g, 6);
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(TMEM_COLS)
                     : "memory");
    }
}

// C = sum over the K splits + bias, relu; dense or scattered like the epilogue.  The order of
// the additions never depends on timing, so results are run-to-run identical.
// Large outputs: one thread per output, splits added in order (coalesced across the threads).
__global__ void __launch_bounds__(256) k_gemm_reduce(const float *__restrict__ partial, int splits,
                                                     int M, int N, const float *__restrict__ bias,
                                                     int relu, float *__restrict__ C, int ldc,
                                                     const int *__restrict__ row_tab,
                                                     int col_stride)
{
    const size_t total = (size_t)M * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        float acc = partial[i];
        for (int s = 1; s < splits; s++) acc += partial[(size_t)s * total + i];
        const int m = (int)(i / N), n = (int)(i % N);
        if (bias) acc += __ldg(bias + n);
        if (relu) acc = fmaxf(acc, 0.f);
        if (row_tab)
            C[(long long)__ldg(row_tab + m) + (long long)n * col_stride] = acc;
        else
            C[(size_t)m * ldc + n] = acc;
    }
}

// Small outputs with many splits (the weight gradients: 32 K outputs, 36 splits): a block of
// 8 warps works on 32 consecutive outputs, warp w adds the splits w, w + 8, ... (coalesced
// 128-byte reads), the eight partial sums meet in shared memory and are added in warp order.
__global__ void __launch_bounds__(256) k_gemm_reduce_deep(const float *__restrict__ partial,
                                                          int splits, int M, int N,
                                                          const float *__restrict__ bias, int relu,
                                                          float *__restrict__ C, int ldc,
                                                          const int *__restrict__ row_tab,
                                                          int col_stride)
{
    __shared__ float red[8][32];
    const size_t total = (size_t)M * N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (size_t base = (size_t)blockIdx.x * 32; base < total; base += (size_t)gridDim.x * 32) {
        const size_t i = base + lane;
        float acc = 0.f;
        if (i < total)
            for (int s = warp; s < splits; s += 8) acc += partial[(size_t)s * total + i];
        red[warp][lane] = acc;
        __syncthreads();
        if (warp == 0 && i < total) {
            acc = red[0][lane];
#pragma unroll
            for (int w = 1; w < 8; w++) acc += red[w][lane];
            const int m = (int)(i / N), n = (int)(i % N);
            if (bias) acc += __ldg(bias + n);
            if (relu) acc = fmaxf(acc, 0.f);
            if (row_tab)
                C[(long long)__ldg(row_tab + m) + (long long)n * col_stride] = acc;
            else
                C[(size_t)m * ldc + n] = acc;
        }
        __syncthreads();
    }
}

int sm_count_cached()
{
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
    }
    return sm_count;
}

int n_tile(int N) { return N <= 32 ? 32 : N <= 64 ? 64 : 128; }

// How the launch cuts the contraction: enough CTAs to cover the SMs, at least 4 k blocks each.
void plan_splits(int M, int N, int K, int *splits, int *kb_per_split)
{
    const int bn = n_tile(N);
    const int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
    const int kb = (K + BK - 1) / BK;
    int s = sm_count_cached() / tiles;
    if (s > kb / 4) s = kb / 4;
    if (s < 1) s = 1;
    int per = (kb + s - 1) / s;
    s = (kb + per - 1) / per; // no empty split
    *splits = s;
    *kb_per_split = per;
}

int fill_operand(const b2rl_gemm_operand *src, int rows, int K, Operand *dst, const char *name)
{
    B2RL_REQUIRE(src && src->data, B2RL_ERR_INVALID, "gemm: operand %s is null", name);
    dst->p = src->data;
    dst->mode = src->mode;
    dst->ld = src->ld;
    dst->vec = 0;
    dst->row_off = src->row_off, dst->row_yx = src->row_yx;
    dst->k_off = src->k_off, dst->k_yx = src->k_yx;
    dst->y_limit = src->y_limit, dst->x_limit = src->x_limit;
    dst->along_k = src->lanes_along_k ? 1 : 0;
    dst->u8 = src->u8 ? 1 : 0;
    dst->scale = src->scale;
    switch (src->mode) {
    case B2RL_GEMM_K_MAJOR:
    case B2RL_GEMM_MN_MAJOR:
        B2RL_REQUIRE(!src->u8, B2RL_ERR_INVALID, "gemm: dense operand %s must be fp32", name);
        B2RL_REQUIRE(src->ld >= (src->mode == B2RL_GEMM_K_MAJOR ? K : rows), B2RL_ERR_RANGE,
                     "gemm: leading dimension of %s smaller than its rows", name);
        B2RL_REQUIRE(((uintptr_t)src->data & 3) == 0, B2RL_ERR_INVALID,
                     "gemm: operand %s must be 4-byte aligned", name);
        dst->vec = (((uintptr_t)src->data & 15) == 0 && src->ld % 4 == 0);
        break;
    case B2RL_GEMM_GATHER:
        B2RL_REQUIRE(src->row_off && src->k_off, B2RL_ERR_INVALID,
                     "gemm: gather operand %s needs its row and k offset tables", name);
        B2RL_REQUIRE((src->row_yx == nullptr) == (src->k_yx == nullptr), B2RL_ERR_INVALID,
                     "gemm: gather operand %s: coordinates on both tables or on none", name);
        B2RL_REQUIRE(((uintptr_t)src->k_off & 15) == 0 && ((uintptr_t)src->k_yx & 15) == 0,
                     B2RL_ERR_INVALID, "gemm: k tables of %s must be 16-byte aligned", name);
        B2RL_REQUIRE(!src->row_yx || (src->y_limit > 0 && src->x_limit > 0), B2RL_ERR_RANGE,
                     "gemm: gather limits of %s out of range", name);
        break;
    default:
        B2RL_REQUIRE(false, B2RL_ERR_INVALID, "gemm: unknown operand mode %d", src->mode);
    }
    return B2RL_OK;
}

} // namespace

static unsigned long long *g_gemm_times = nullptr; // debug: set by b2rl_gemm_debug_times

// Debug aid (tools/gemm_phases.py): stamps of the phases of every CTA of the next launches go
// to `device_buffer` (8 x uint64 per CTA, large enough for the largest grid); NULL switches off.
extern "C" int b2rl_gemm_debug_times(void *device_buffer)
{
    g_gemm_times = (unsigned long long *)device_buffer;
    return B2RL_OK;
}

extern "C" int64_t b2rl_gemm_workspace_bytes(int32_t M, int32_t N, int32_t K)
{
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    int splits, per;
    plan_splits(M, N, K, &splits, &per);
    return splits > 1 ? (int64_t)splits * M * N * (int64_t)sizeof(float) : 0;
}

extern "C" int b2rl_gemm_tf32x3_ex(const b2rl_gemm_operand *A, const b2rl_gemm_operand *B,
                                   const b2rl_gemm_output *C, int32_t M, int32_t N, int32_t K,
                                   void *workspace, int64_t workspace_bytes, void *stream)
{
    B2RL_REQUIRE(C && C->data, B2RL_ERR_INVALID, "gemm_tf32x3: null output");
    B2RL_REQUIRE(M > 0 && N > 0 && K > 0, B2RL_ERR_RANGE, "gemm_tf32x3: empty product");
    GemmArgs g;
    int rc = fill_operand(A, M, K, &g.A, "A");
    if (rc != B2RL_OK) return rc;
    rc = fill_operand(B, N, K, &g.B, "B");
    if (rc != B2RL_OK) return rc;
    g.bias = C->bias, g.C = C->data;
    g.c_row_tab = C->row_tab, g.c_col_stride = C->col_stride;
    B2RL_REQUIRE(C->row_tab || C->ld >= N, B2RL_ERR_RANGE,
                 "gemm_tf32x3: leading dimension of C smaller than its rows");
    B2RL_REQUIRE(((uintptr_t)C->data & 3) == 0, B2RL_ERR_INVALID,
                 "gemm_tf32x3: C must be 4-byte aligned");
    g.M = M, g.N = N, g.K = K, g.ldc = C->ld;
    g.c_vec = (!C->row_tab && ((uintptr_t)C->data & 15) == 0 && C->ld % 4 == 0);
    g.relu = C->relu ? 1 : 0;
    g.times = g_gemm_times;
    g.bn = n_tile(N);
    plan_splits(M, N, K, &g.splits, &g.kb_per_split);
    g.partial = nullptr;
    if (g.splits > 1) {
        const int64_t need = (int64_t)g.splits * M * N * (int64_t)sizeof(float);
        B2RL_REQUIRE(workspace && workspace_bytes >= need && ((uintptr_t)workspace & 15) == 0,
                     B2RL_ERR_INVALID,
                     "gemm_tf32x3: workspace of %lld bytes (16-byte aligned) required",
                     (long long)need);
        g.partial = (float *)workspace;
    }
    typedef void (*kernel_t)(GemmArgs);
    static const kernel_t kernels[3][3] = {
        {k_gemm_tf32x3<0, 0>, k_gemm_tf32x3<0, 1>, k_gemm_tf32x3<0, 2>},
        {k_gemm_tf32x3<1, 0>, k_gemm_tf32x3<1, 1>, k_gemm_tf32x3<1, 2>},
        {k_gemm_tf32x3<2, 0>, k_gemm_tf32x3<2, 1>, k_gemm_tf32x3<2, 2>}};
    const kernel_t kernel = kernels[g.A.mode][g.B.mode];
    static bool attr_set[64][3][3]; // per device and instantiation
    int cur_dev = 0;
    B2RL_CUDA(cudaGetDevice(&cur_dev));
    if (cur_dev < 0 || cur_dev >= 64 || !attr_set[cur_dev][g.A.mode][g.B.mode]) {
        B2RL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM_BYTES));
        if (cur_dev >= 0 && cur_dev < 64) attr_set[cur_dev][g.A.mode][g.B.mode] = true;
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + g.bn - 1) / g.bn);
    kernel<<<tiles * g.splits, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(g);
    B2RL_CUDA(cudaGetLastError());
    if (g.splits > 1) {
        const long long total = (long long)M * N;
        if (total >= 65536 || g.splits <= 8) {
            long long blocks = (total + 255) / 256;
            if (blocks > 8ll * sm_count_cached()) blocks = 8ll * sm_count_cached();
            k_gemm_reduce<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
                g.partial, g.splits, M, N, C->bias, g.relu, C->data, C->ld, C->row_tab,
                C->col_stride);
        } else {
            long long blocks = (total + 31) / 32;
            if (blocks > 8ll * sm_count_cached()) blocks = 8ll * sm_count_cached();
            k_gemm_reduce_deep<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
                g.partial, g.splits, M, N, C->bias, g.relu, C->data, C->ld, C->row_tab,
                C->col_stride);
        }
        B2RL_CUDA(cudaGetLastError());
    }
    return B2RL_OK;
}

extern "C" int b2rl_gemm_tf32x3(const float *A, int32_t lda, int32_t a_mn_major, const float *B,
                                int32_t ldb, int32_t b_mn_major, const float *bias, int32_t relu,
                                float *C, int32_t ldc, int32_t M, int32_t N, int32_t K,
                                void *workspace, int64_t workspace_bytes, void *stream)
{
    b2rl_gemm_operand a, b;
    b2rl_gemm_output c;
    memset(&a, 0, sizeof a), memset(&b, 0, sizeof b), memset(&c, 0, sizeof c);
    a.data = A, a.ld = lda, a.mode = a_mn_major ? B2RL_GEMM_MN_MAJOR : B2RL_GEMM_K_MAJOR;
    b.data = B, b.ld = ldb, b.mode = b_mn_major ? B2RL_GEMM_MN_MAJOR : B2RL_GEMM_K_MAJOR;
    c.data = C, c.ld = ldc, c.bias = bias, c.relu = relu;
    return b2rl_gemm_tf32x3_ex(&a, &b, &c, M, N, K, workspace, workspace_bytes, stream);
}
