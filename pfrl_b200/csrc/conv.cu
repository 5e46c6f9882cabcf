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

__global__ void __launch_bounds__(THREADS, 1)
k_conv_nature1(const float *__restrict__ x, const float *__restrict__ w,
               const float *__restrict__ bias, float *__restrict__ out, int n_images)
{
    extern __shared__ __align__(128) float smem[];
    float *img = smem;              // [C][H][W]
    float *wsm = smem + IMG;        // [C][KS][KS][O]  (output channel fastest)
    uint64_t *bar = reinterpret_cast<uint64_t *>(wsm + WSZ);
    const int tid = threadIdx.x;

    // filters: global [O][C][KS][KS] -> shared [C][KS][KS][O]
    for (int i = tid; i < WSZ; i += THREADS) {
        const int o = i / (C * KS * KS), r = i - o * (C * KS * KS);
        wsm[r * O + o] = w[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int cg = tid & 7;        // channels 4*cg .. 4*cg+3
    const int pg = tid >> 3;       // 0..39: output row = pg / 2, column half = pg & 1
    const int oy = pg >> 1;
    const int ox0 = (pg & 1) * COLS;
    float b4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) b4[j] = bias ? bias[4 * cg + j] : 0.f;

    uint32_t phase = 0;
    for (int n = blockIdx.x; n < n_images; n += gridDim.x) {
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

extern "C" int b2rl_conv_nature1_fwd(const float *x, const float *w, const float *bias,
                                     int32_t n_images, float *out, void *stream)
{
    B2RL_REQUIRE(x && w && out, B2RL_ERR_INVALID, "conv_nature1_fwd: null argument");
    B2RL_REQUIRE(n_images > 0, B2RL_ERR_RANGE, "conv_nature1_fwd: empty batch");
    B2RL_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 7) == 0, B2RL_ERR_INVALID,
                 "conv_nature1_fwd: x must be 16-byte aligned, out 8-byte aligned");
    const size_t smem = sizeof(float) * (IMG + WSZ) + 16;
    static bool attr_set[64]; // per device
    int cur_dev = 0;
    B2RL_CUDA(cudaGetDevice(&cur_dev));
    if (cur_dev < 0 || cur_dev >= 64 || !attr_set[cur_dev]) {
        B2RL_CUDA(cudaFuncSetAttribute(k_conv_nature1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        if (cur_dev >= 0 && cur_dev < 64) attr_set[cur_dev] = true;
    }
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
    }
    const int grid = n_images < sm_count ? n_images : sm_count;
    k_conv_nature1<<<grid, THREADS, smem, (cudaStream_t)stream>>>(x, w, bias, out, n_images);
    B2RL_CUDA(cudaGetLastError());
    return B2RL_OK;
}
