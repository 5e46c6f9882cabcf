// conv.cu -- fp32 direct convolution for the first Nature-DQN layer.
//
// The Rainbow / DQN update runs the 84x84x4 -> 20x20x32 (8x8, stride 4)
// convolution three times per minibatch (online net on s and s', target net
// on s').  With TF32 disabled (the 1e-5 loss-parity configuration) cuDNN
// serves it with `implicit_convolve_sgemm` at ~8 TFLOP/s (403 us for 512
// images on B200, profiles/README.md); this kernel keeps exact fp32 FFMA
// accumulation and runs the same layer several times faster by keeping one
// whole image (113 KB, brought in by ONE bulk async copy) and the
// re-laid-out filters (32 KB) in shared memory and register-tiling
// 10 output columns x 4 output channels per thread, so that 19 128-bit
// shared loads feed 320 FFMAs.
//
// Replaces nothing in the reference (it calls nn.Conv2d -> cuDNN,
// pfrl/nn/atari_cnn.py:30-36, pfrl/q_functions/dueling_dqn.py:34-40,91-97);
// this is the SURVEY's "K10" dense-contraction kernel for the one layer that
// dominated the measured step.  Backward (weight / bias gradients only: the
// input is data) stays with cuDNN through aten::convolution_backward.
#include "b2rl_internal.cuh"

namespace {

constexpr int C = 4, H = 84, W = 84, KS = 8, ST = 4, O = 32, P = 20; // P = (H - KS) / ST + 1
constexpr int IMG = C * H * W;           // 28224 floats = 112896 B
constexpr int WSZ = O * C * KS * KS;     // 8192 floats
constexpr int THREADS = 320;             // 40 position groups x 8 channel groups
constexpr int COLS = 10;                 // output columns per thread
constexpr int SEG = (COLS - 1) * ST + KS; // 44 input floats per row segment

__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// U8 = true: the images are uint8 (the replay gather's RAW output, 28 KB instead of
// 113 KB per image); they are brought in by bulk async copies, double-buffered so that the
// next image lands while this one is convolved, and expanded to float(x) * scale in
// shared memory -- the very values the ScaleU8 gather would have written to HBM, so the
// result is bit-identical to the f32 path.
template <bool U8>
__global__ void __launch_bounds__(THREADS, 1)
k_conv_nature1(const void *__restrict__ xin, float scale, const float *__restrict__ w,
               const float *__restrict__ bias, float *__restrict__ out, int n_images)
{
    extern __shared__ __align__(128) float smem[];
    float *img = smem;              // [C][H][W]
    float *wsm = smem + IMG;        // [C][KS][KS][O]  (output channel fastest)
    uint64_t *bar = reinterpret_cast<uint64_t *>(wsm + WSZ); // [2]
    uint8_t *raw = reinterpret_cast<uint8_t *>(bar + 2);      // U8: [2][IMG] bytes
    const float *x = static_cast<const float *>(xin);
    const uint8_t *x8 = static_cast<const uint8_t *>(xin);
    const int tid = threadIdx.x;

    // filters: global [O][C][KS][KS] -> shared [C][KS][KS][O]
    for (int i = tid; i < WSZ; i += THREADS) {
        const int o = i / (C * KS * KS), r = i - o * (C * KS * KS);
        wsm[r * O + o] = w[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto fetch_u8 = [&](int n, int buf) { // thread 0: image n -> raw[buf]
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                         smem_addr(bar + buf)),
                     "r"(IMG)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                "r"(smem_addr(raw + (size_t)buf * IMG)),
            "l"(x8 + (size_t)n * IMG), "r"(IMG), "r"(smem_addr(bar + buf))
            : "memory");
    };
    if (U8 && tid == 0 && (int)blockIdx.x < n_images) fetch_u8(blockIdx.x, 0);

    const int cg = tid & 7;        // channels 4*cg .. 4*cg+3
    const int pg = tid >> 3;       // 0..39: output row = pg / 2, column half = pg & 1
    const int oy = pg >> 1;
    const int ox0 = (pg & 1) * COLS;
    float b4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) b4[j] = bias ? bias[4 * cg + j] : 0.f;

    uint32_t phase = 0;       // f32 path: phase of bar[0]
    uint32_t ph8[2] = {0, 0}; // u8 path: phases of bar[0], bar[1]
    int it = 0;
    for (int n = blockIdx.x; n < n_images; n += gridDim.x, it++) {
        if constexpr (U8) {
            const int buf = it & 1;
            asm volatile(
                "{\n.reg .pred p;\nW_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_addr(bar + buf)),
                "r"(ph8[buf])
                : "memory");
            ph8[buf] ^= 1;
            // expand: 4 pixels per step, float(u8) * scale (= the ScaleU8 gather's arithmetic)
            const uint32_t *r4 = reinterpret_cast<const uint32_t *>(raw + (size_t)buf * IMG);
            float4 *i4 = reinterpret_cast<float4 *>(img);
            for (int i = tid; i < IMG / 4; i += THREADS) {
                const uint32_t v = r4[i];
                float4 f;
                f.x = (float)(v & 0xffu) * scale;
                f.y = (float)((v >> 8) & 0xffu) * scale;
                f.z = (float)((v >> 16) & 0xffu) * scale;
                f.w = (float)(v >> 24) * scale;
                i4[i] = f;
            }
            __syncthreads();
            // the other byte buffer was expanded one iteration ago: refill it now
            if (tid == 0 && n + (int)gridDim.x < n_images) fetch_u8(n + gridDim.x, buf ^ 1);
        } else {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                                 smem_addr(bar)),
                             "r"(IMG * 4)
                             : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                        "r"(smem_addr(img)),
                    "l"(x + (size_t)n * IMG), "r"(IMG * 4), "r"(smem_addr(bar))
                    : "memory");
            }
            asm volatile(
                "{\n.reg .pred p;\nW_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_addr(bar)),
                "r"(phase)
                : "memory");
            phase ^= 1;
        }

        float acc[COLS][4];
#pragma unroll
        for (int i = 0; i < COLS; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = b4[j];

#pragma unroll 1
        for (int c = 0; c < C; c++) {
#pragma unroll 2
            for (int ky = 0; ky < KS; ky++) {
                const float4 *row = reinterpret_cast<const float4 *>(
                    img + (c * H + (oy * ST + ky)) * W + ox0 * ST);
                float in[SEG];
#pragma unroll
                for (int q = 0; q < SEG / 4; q++) {
                    const float4 v = row[q];
                    in[4 * q] = v.x;
                    in[4 * q + 1] = v.y;
                    in[4 * q + 2] = v.z;
                    in[4 * q + 3] = v.w;
                }
                const float4 *wr =
                    reinterpret_cast<const float4 *>(wsm + ((c * KS + ky) * KS) * O + 4 * cg);
#pragma unroll
                for (int kx = 0; kx < KS; kx++) {
                    const float4 wv = wr[kx * (O / 4)];
#pragma unroll
                    for (int i = 0; i < COLS; i++) {
                        const float a = in[i * ST + kx];
                        acc[i][0] = fmaf(a, wv.x, acc[i][0]);
                        acc[i][1] = fmaf(a, wv.y, acc[i][1]);
                        acc[i][2] = fmaf(a, wv.z, acc[i][2]);
                        acc[i][3] = fmaf(a, wv.w, acc[i][3]);
                    }
                }
            }
        }
        float *o_img = out + (size_t)n * O * P * P;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float *dst = o_img + ((4 * cg + j) * P + oy) * P + ox0;
#pragma unroll
            for (int i = 0; i < COLS; i += 2)
                *reinterpret_cast<float2 *>(dst + i) = make_float2(acc[i][j], acc[i + 1][j]);
        }
        __syncthreads(); // everyone is done with `img` before the next bulk copy lands
    }
}

} // namespace

template <bool U8>
static int launch_conv1(const void *x, float scale, const float *w, const float *bias,
                        int32_t n_images, float *out, void *stream)
{
    const size_t smem = sizeof(float) * (IMG + WSZ) + 16 + (U8 ? 2 * (size_t)IMG : 0);
    static bool attr_set[64]; // per device
    int cur_dev = 0;
    B2RL_CUDA(cudaGetDevice(&cur_dev));
    if (cur_dev < 0 || cur_dev >= 64 || !attr_set[cur_dev]) {
        B2RL_CUDA(cudaFuncSetAttribute(k_conv_nature1<U8>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (cur_dev >= 0 && cur_dev < 64) attr_set[cur_dev] = true;
    }
    static int sm_count = 0;
    if (!sm_count) {
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, cur_dev);
        if (sm_count <= 0) sm_count = 148;
    }
    const int grid = n_images < sm_count ? n_images : sm_count;
    k_conv_nature1<U8><<<grid, THREADS, smem, (cudaStream_t)stream>>>(x, scale, w, bias, out,
                                                                     n_images);
    B2RL_CUDA(cudaGetLastError());
    return B2RL_OK;
}

extern "C" int b2rl_conv_nature1_fwd(const float *x, const float *w, const float *bias,
                                     int32_t n_images, float *out, void *stream)
{
    B2RL_REQUIRE(x && w && out, B2RL_ERR_INVALID, "conv_nature1_fwd: null argument");
    B2RL_REQUIRE(n_images > 0, B2RL_ERR_RANGE, "conv_nature1_fwd: empty batch");
    B2RL_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 7) == 0, B2RL_ERR_INVALID,
                 "conv_nature1_fwd: x must be 16-byte aligned, out 8-byte aligned");
    return launch_conv1<false>(x, 1.0f, w, bias, n_images, out, stream);
}

extern "C" int b2rl_conv_nature1_fwd_u8(const uint8_t *x, float scale, const float *w,
                                        const float *bias, int32_t n_images, float *out,
                                        void *stream)
{
    B2RL_REQUIRE(x && w && out, B2RL_ERR_INVALID, "conv_nature1_fwd_u8: null argument");
    B2RL_REQUIRE(n_images > 0, B2RL_ERR_RANGE, "conv_nature1_fwd_u8: empty batch");
    B2RL_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 7) == 0, B2RL_ERR_INVALID,
                 "conv_nature1_fwd_u8: x must be 16-byte aligned, out 8-byte aligned");
    return launch_conv1<true>(x, scale, w, bias, n_images, out, stream);
}
