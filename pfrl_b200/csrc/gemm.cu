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
// The loads of k block i+1 are issued before block i is split and stored, so the global-load
// latency hides behind the conversion.
//
// Roles in a CTA of 21 warps, one 128 x 128 output tile (x one K split) per CTA:
//   warps 0-3   epilogue: tcgen05.ld the two accumulators (TMEM lane = tile row), add, bias /
//               relu, store
//   warp  4     allocates 256 TMEM columns; lane 0 issues 3 MMAs per 8-wide k step (hi.hi into
//               one accumulator, lo.hi + hi.lo into the other)
//   warps 5-20  loaders, 3-stage ring of (A_hi, A_lo, B_hi, B_lo) = 64 KB per stage,
//               full/empty mbarriers; the empty side is armed by tcgen05.commit
// Small products are split along K so that the grid covers the 148 SMs; the partial tiles go
// to a workspace and `k_gemm_reduce` adds them in a fixed order (deterministic), with the
// bias / relu epilogue.
//
// Convolutions (pfrl/nn/atari_cnn.py:30-44, pfrl/q_functions/dueling_dqn.py:34-40,91-97: the
// 4x4/2 and 3x3/1 layers of the Nature trunk, forward, input gradient, weight gradient) are the
// same product with one more operand mode, "gather": element (row, k) lives at
//   row_off[row] + k_off[k] + ((y[row] + dy[k]) >> shift) * pitch + ((x[row] + dx[k]) >> shift)
// and exists iff 0 <= y + dy < y_limit, 0 <= x + dx < x_limit and both are multiples of
// 1 << shift -- im2col, its transpose (col2im as a gather, strided layers included) and the
// pixel-major views of the weight gradient, all read in place through two small index tables
// per operand (ops/conv.py builds them once per layer shape).  The output can be scattered the
// same way (C[m, n] at row_off[m] + n * col_stride: NCHW activations and transposed weight
// gradients), so no im2col buffer and no layout pass ever touches HBM.  uint8 sources
// (float(byte) * scale: the x / 255 of the Atari phi) are read directly.
#include "b2rl_internal.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32; // BK floats = 128 bytes of contraction per stage
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 4;
constexpr int MMA_WARP = 4;
constexpr int LOAD_WARPS = 16;
constexpr int THREADS = (EPI_WARPS + 1 + LOAD_WARPS) * 32; // 672
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
    // gather mode (see the header comment)
    const int2 *row_tab; // (element offset, y | x << 16)
    const int2 *k_tab;   // (element offset, dy | dx << 16), int16 each
    int y_limit, x_limit, shift, pitch;
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
};

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

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile("{ .reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (done) return;
        if (clock64() - t0 > WAIT_LIMIT) __trap();
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
    int2 tab[2];       // gather: row table entries
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
        L.tab[i] = make_int2(0, 0);
        if (!ok) continue;
        if (MODE == MODE_K_MAJOR)
            L.q[i] = (const float *)o.p + (size_t)(mn0 + r) * o.ld + k_first + kc * 4;
        else if (MODE == MODE_MN_MAJOR)
            L.q[i] = (const float *)o.p + (size_t)(k_first + kc * 4) * o.ld + mn0 + r;
        else {
            L.q[i] = (const float *)o.p; // (marks the row as present)
            L.tab[i] = __ldg(o.row_tab + mn0 + r);
        }
    }
}

__device__ __forceinline__ float gather1(const Operand &o, int2 rt, int k, int k_total)
{
    if (k >= k_total) return 0.f;
    const int2 kt = __ldg(o.k_tab + k);
    const int yy = (int)(short)(rt.y & 0xffff) + (int)(short)(kt.y & 0xffff);
    const int xx = (rt.y >> 16) + (kt.y >> 16);
    const int mask = (1 << o.shift) - 1;
    if ((unsigned)yy >= (unsigned)o.y_limit || (unsigned)xx >= (unsigned)o.x_limit ||
        ((yy | xx) & mask))
        return 0.f;
    const int idx = rt.x + kt.x + (yy >> o.shift) * o.pitch + (xx >> o.shift);
    if (o.u8) return (float)__ldg((const uint8_t *)o.p + idx) * o.scale;
    return __ldg((const float *)o.p + idx);
}

// k block `kb` (relative to the first one of this CTA) -> registers.  `full`: the whole block
// lies inside K (warp-uniform), which is the common case and needs no per-element checks.
template <int MODE>
__device__ __forceinline__ void fetch_operand(const Operand &o, const Lane<MODE> &L, int kb,
                                              int k0, int k_total, Frag &f)
{
    const bool full = k0 + BK <= k_total;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        f.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (L.q[i] == nullptr) continue;
        if (MODE == MODE_K_MAJOR) {
            const float *q = L.q[i] + (size_t)kb * BK;
            if (full && o.vec)
                f.v[i] = __ldg(reinterpret_cast<const float4 *>(q));
            else
                f.v[i] = load4(q, true, k_total - (k0 + L.kofs[i]), o.vec);
        } else if (MODE == MODE_MN_MAJOR) {
            const float *q = L.q[i] + (size_t)kb * BK * o.ld;
            const int left = k_total - (k0 + L.kofs[i]);
            if (full || left > 0) f.v[i].x = __ldg(q);
            if (full || left > 1) f.v[i].y = __ldg(q + o.ld);
            if (full || left > 2) f.v[i].z = __ldg(q + 2 * (size_t)o.ld);
            if (full || left > 3) f.v[i].w = __ldg(q + 3 * (size_t)o.ld);
        } else {
            const int kk = k0 + L.kofs[i];
            f.v[i].x = gather1(o, L.tab[i], kk, k_total);
            f.v[i].y = gather1(o, L.tab[i], kk + 1, k_total);
            f.v[i].z = gather1(o, L.tab[i], kk + 2, k_total);
            f.v[i].w = gather1(o, L.tab[i], kk + 3, k_total);
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

    if (warp > MMA_WARP) {
        // ---------------- loaders ----------------
        const int lw = warp - MMA_WARP - 1;
        Lane<AMODE> la;
        Lane<BMODE> lb;
        lane_setup<AMODE>(g.A, m0, g.M, BM, kb0 * BK, lw, lane, la);
        lane_setup<BMODE>(g.B, n0, g.N, g.bn, kb0 * BK, lw, lane, lb);
        Frag fa, fb;
        fetch_operand<AMODE>(g.A, la, 0, kb0 * BK, g.K, fa);
        fetch_operand<BMODE>(g.B, lb, 0, kb0 * BK, g.K, fb);
        for (int it = 0; it < n_kb; it++) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            Frag na = fa, nb = fb; // next k block: in flight while this one is split and stored
            if (it + 1 < n_kb) {
                const int k1 = (kb0 + it + 1) * BK;
                fetch_operand<AMODE>(g.A, la, it + 1, k1, g.K, na);
                fetch_operand<BMODE>(g.B, lb, it + 1, k1, g.K, nb);
            }
            if (lane == 0) mbar_wait(empty0 + 8 * s, ph ^ 1);
            __syncwarp();
            uint8_t *st = smem + s * STAGE_BYTES;
            store_operand<AMODE>(la, fa, st, st + TILE_BYTES);
            store_operand<BMODE>(lb, fb, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES);
            // generic-proxy stores -> visible to the tensor core (async proxy)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
            fa = na, fb = nb;
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
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: warp w owns TMEM lanes 32w .. 32w+31 ----------------
        if (lane == 0) mbar_wait(accum_bar, 0);
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + warp * 32 + lane;
        const bool direct = g.splits == 1;
        const bool scatter = direct && g.c_row_tab != nullptr;
        float *out = direct ? g.C : g.partial + (size_t)split * g.M * g.N;
        const int ldo = direct ? g.ldc : g.N;
        const bool vec = direct ? g.c_vec : (g.N % 4 == 0);
        const long long obase = row < g.M ? (scatter ? (long long)__ldg(g.c_row_tab + row)
                                                     : (long long)row * ldo)
                                          : 0;
        const int ostride = scatter ? g.c_col_stride : 1;
#pragma unroll 1
        for (int c = 0; c < g.bn / 16; c++) {
            uint32_t r[16], q[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c * 16;
            tmem_ld16(taddr, r);
            tmem_ld16(taddr + BN, q);
            tmem_wait();
            const int col0 = n0 + c * 16;
            if (row < g.M && col0 < g.N) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    v[j] = __uint_as_float(r[j]) + __uint_as_float(q[j]);
                    if (direct && col0 + j < g.N) {
                        if (g.bias) v[j] += __ldg(g.bias + col0 + j);
                        if (g.relu) v[j] = fmaxf(v[j], 0.f);
                    }
                }
                float *o = out + obase + (long long)col0 * ostride;
                if (!scatter && vec && col0 + 16 <= g.N) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                    // scatter: lane = row, so for each column the warp writes one run
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (col0 + j < g.N) o[(long long)j * ostride] = v[j];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(TMEM_COLS)
                     : "memory");
    }
}

// C = sum over the K splits (fixed order) + bias, relu; dense or scattered like the epilogue.
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
    dst->row_tab = (const int2 *)src->row_tab;
    dst->k_tab = (const int2 *)src->k_tab;
    dst->y_limit = src->y_limit, dst->x_limit = src->x_limit;
    dst->shift = src->shift, dst->pitch = src->pitch;
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
        B2RL_REQUIRE(src->row_tab && src->k_tab, B2RL_ERR_INVALID,
                     "gemm: gather operand %s needs its row and k tables", name);
        B2RL_REQUIRE(((uintptr_t)src->row_tab & 7) == 0 && ((uintptr_t)src->k_tab & 7) == 0,
                     B2RL_ERR_INVALID, "gemm: tables of %s must be 8-byte aligned", name);
        B2RL_REQUIRE(src->shift >= 0 && src->shift < 8 && src->y_limit > 0 && src->x_limit > 0,
                     B2RL_ERR_RANGE, "gemm: gather geometry of %s out of range", name);
        break;
    default:
        B2RL_REQUIRE(false, B2RL_ERR_INVALID, "gemm: unknown operand mode %d", src->mode);
    }
    return B2RL_OK;
}

} // namespace

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
        long long blocks = (total + 255) / 256;
        if (blocks > 4ll * sm_count_cached()) blocks = 4ll * sm_count_cached();
        k_gemm_reduce<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
            g.partial, g.splits, M, N, C->bias, g.relu, C->data, C->ld, C->row_tab,
            C->col_stride);
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
