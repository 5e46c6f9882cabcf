// losses.cu -- fused TD-error / loss kernels of the DQN family.
//
// Replaces chains of small ATen launches + host round trips in the reference:
//   pfrl/agents/dqn.py:44-104,388-470          Huber / MSE value loss, |y-t|
//   pfrl/agents/categorical_dqn.py:7-57        _apply_categorical_projection
//   pfrl/agents/categorical_dqn.py:60-97,178-204  cross entropy, per-sample
//                                              priority error, weighted sum
//   pfrl/agents/iqn.py:176-208                 quantile Huber loss
//
// All arithmetic is fp32 in the reference's operation order (no FMA
// contraction where the reference rounds between ops), so losses agree to
// ~1e-7 relative; the batch reductions are deterministic (fixed-order tree in
// the last CTA to finish).
#include <math.h>

#include "b2rl_internal.cuh"

namespace {

__device__ unsigned int g_done_counter[B2RL_N_TICKETS]; // zero-initialised, self-resetting

__device__ __forceinline__ float warp_sum(float v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic final reduction: the last CTA to arrive sums term[i] for
// i < n in a fixed order and writes *out = sum / divisor.
__device__ void finish_sum(const float *term, int n, float divisor, float *out, float *sh,
                           unsigned ticket)
{
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicInc(&g_done_counter[ticket], gridDim.x - 1);
        last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += __ldcg(term + i);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        s = warp_sum(s);
        if (threadIdx.x == 0) *out = __fdiv_rn(s, divisor);
    }
}

// ---------------------------------------------------------------------------
// C51: one warp per sample.
// ---------------------------------------------------------------------------
constexpr int C51_MAX_ATOMS = 256;
constexpr int C51_WARPS = 8;

struct C51Args {
    const float *y;       // [B, n] predicted probabilities of the taken action
    const float *next_p;  // [B, n] target-net probabilities of the greedy next action
    const float *reward, *discount, *terminal; // [B]
    const float *weights; // [B] or null
    const float *z;       // [n] atom values
    int B, n, mean;
    float *t_out;     // [B, n] projected target (saved for backward)
    float *delta_out; // [B] per-sample loss = priority error
    float *term;      // [B] scratch: w_i * delta_i
    float *loss_out;  // [1]
    unsigned ticket;  // completion counter of this launch
};

__global__ void __launch_bounds__(C51_WARPS * 32) k_c51_fwd(C51Args a)
{
    __shared__ float zt[C51_WARPS][C51_MAX_ATOMS];
    __shared__ float sh[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * C51_WARPS + warp;
    const int n = a.n;
    if (i < a.B) {
        float *row = zt[warp];
        for (int j = lane; j < n; j += 32) row[j] = 0.f;
        __syncwarp();
        const float v_min = a.z[0], v_max = a.z[n - 1];
        const float delta_z = __fsub_rn(a.z[1], a.z[0]);
        const float r = a.reward[i];
        // (1 - terminal) * discount, then * z, then + r: categorical_dqn.py:147-152
        const float scale = __fmul_rn(__fsub_rn(1.0f, a.terminal[i]), a.discount[i]);
        for (int j = lane; j < n; j += 32) {
            float tz = __fadd_rn(r, __fmul_rn(scale, a.z[j]));
            tz = fminf(fmaxf(tz, v_min), v_max);
            float bj = __fdiv_rn(__fsub_rn(tz, v_min), delta_z);
            bj = fminf(fmaxf(bj, 0.f), (float)(n - 1));
            const float lo = floorf(bj), up = ceilf(bj);
            const float frac = __fsub_rn(bj, lo);
            const float p = a.next_p[(size_t)i * n + j];
            atomicAdd(&row[(int)lo], __fmul_rn(p, __fsub_rn(1.0f, frac)));
            atomicAdd(&row[(int)up], __fmul_rn(p, frac));
        }
        __syncwarp();
        float acc = 0.f;
        for (int j = lane; j < n; j += 32) {
            const float t = row[j];
            a.t_out[(size_t)i * n + j] = t;
            const float yc = fminf(fmaxf(a.y[(size_t)i * n + j], 1e-10f), 1.0f);
            acc += __fmul_rn(-t, logf(yc)); // -t * log(clamp(y, 1e-10, 1)), :183
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            a.delta_out[i] = acc;
            a.term[i] = a.weights ? __fmul_rn(acc, a.weights[i]) : acc;
        }
    }
    finish_sum(a.term, a.B, a.mean ? (float)a.B : 1.0f, a.loss_out, sh, a.ticket);
}

struct C51BwdArgs {
    const float *y, *t, *weights, *grad_loss;
    int B, n, mean;
    float *grad_y;
};

__global__ void k_c51_bwd(C51BwdArgs a)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * a.n) return;
    const int i = idx / a.n;
    float g = a.grad_loss[0];
    if (a.mean) g = g / (float)a.B;
    if (a.weights) g *= a.weights[i];
    const float y = a.y[idx];
    // d/dy [-t log(clamp(y))] = -t / y inside the clamp range, 0 outside
    a.grad_y[idx] = (y >= 1e-10f && y <= 1.0f) ? g * (-a.t[idx] / y) : 0.f;
}

// ---------------------------------------------------------------------------
// Scalar TD loss (DQN / DoubleDQN): y = Q(s)[a], t = r + disc (1-term) next_q
// ---------------------------------------------------------------------------
struct TdArgs {
    const float *q;        // [B, nA]
    const long long *action; // [B]
    const float *next_q;   // [B] max / double-Q value of the next state
    const float *reward, *discount, *terminal, *weights;
    int B, nA, clip_delta, mean;
    float *y_out, *t_out, *delta_out, *term, *loss_out;
    unsigned ticket;
};

__global__ void __launch_bounds__(256) k_td_fwd(TdArgs a)
{
    __shared__ float sh[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.B) {
        const float y = a.q[(size_t)i * a.nA + a.action[i]];
        // r + discount * (1 - terminal) * next_q_max, dqn.py:405
        const float t = __fadd_rn(
            a.reward[i],
            __fmul_rn(__fmul_rn(a.discount[i], __fsub_rn(1.0f, a.terminal[i])), a.next_q[i]));
        const float d = __fsub_rn(y, t);
        const float ad = fabsf(d);
        float l;
        if (a.clip_delta)
            l = ad < 1.0f ? __fmul_rn(__fmul_rn(0.5f, d), d) : __fsub_rn(ad, 0.5f);
        else
            l = __fmul_rn(__fmul_rn(d, d), 0.5f);
        a.y_out[i] = y;
        a.t_out[i] = t;
        a.delta_out[i] = ad;
        a.term[i] = a.weights ? __fmul_rn(l, a.weights[i]) : l;
    }
    finish_sum(a.term, a.B, a.mean ? (float)a.B : 1.0f, a.loss_out, sh, a.ticket);
}

struct TdBwdArgs {
    const float *y, *t, *weights, *grad_loss;
    const long long *action;
    int B, nA, clip_delta, mean;
    float *grad_q; // [B, nA]
};

__global__ void k_td_bwd(TdBwdArgs a)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * a.nA) return;
    const int i = idx / a.nA, j = idx - i * a.nA;
    float gq = 0.f;
    if (j == (int)a.action[i]) {
        float g = a.grad_loss[0];
        if (a.mean) g = g / (float)a.B;
        if (a.weights) g *= a.weights[i];
        float d = a.y[i] - a.t[i];
        if (a.clip_delta) d = fminf(fmaxf(d, -1.0f), 1.0f); // Huber'(d), delta = 1
        gq = g * d;
    }
    a.grad_q[idx] = gq;
}

// ---------------------------------------------------------------------------
// Quantile Huber loss (IQN): y [B, N], t [B, N'], taus [B, N]
// ---------------------------------------------------------------------------
struct QhArgs {
    const float *y, *t, *taus, *weights;
    int B, N, Np, mean;
    float *delta_out, *term, *loss_out;
    unsigned ticket;
};

// one CTA per sample, N x N' pairs strided over the threads
__global__ void __launch_bounds__(256) k_qh_fwd(QhArgs a)
{
    __shared__ float sh[32];
    __shared__ float s_tot;
    const int i = blockIdx.x;
    float acc = 0.f;
    const int pairs = a.N * a.Np;
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
        const int n = p / a.Np, m = p - n * a.Np;
        const float y = a.y[(size_t)i * a.N + n];
        const float t = a.t[(size_t)i * a.Np + m];
        const float tau = a.taus[(size_t)i * a.N + n];
        const float d = __fsub_rn(t, y);
        const float ad = fabsf(d);
        // smooth_l1 (beta = 1): iqn.py:198-203
        const float hub = ad < 1.0f ? __fmul_rn(__fmul_rn(0.5f, d), d) : __fsub_rn(ad, 0.5f);
        const float I = d < 0.f ? 1.0f : 0.0f; // bellman error < 0
        acc += __fmul_rn(fabsf(__fsub_rn(tau, I)), hub);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        s = warp_sum(s);
        if (threadIdx.x == 0) s_tot = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // loss per sample: mean over N' then sum over N (iqn.py:210-250)
        const float li = s_tot / (float)a.Np;
        a.delta_out[i] = s_tot / (float)(a.N * a.Np); // eltwise_loss.mean((1, 2)), iqn.py:388
        a.term[i] = a.weights ? __fmul_rn(li, a.weights[i]) : li;
    }
    finish_sum(a.term, a.B, a.mean ? (float)a.B : 1.0f, a.loss_out, sh, a.ticket);
}

struct QhBwdArgs {
    const float *y, *t, *taus, *weights, *grad_loss;
    int B, N, Np, mean;
    float *grad_y; // [B, N]
};

__global__ void k_qh_bwd(QhBwdArgs a)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * a.N) return;
    const int i = idx / a.N;
    float g = a.grad_loss[0] / (float)a.Np;
    if (a.mean) g = g / (float)a.B;
    if (a.weights) g *= a.weights[i];
    const float y = a.y[idx], tau = a.taus[idx];
    float acc = 0.f;
    for (int m = 0; m < a.Np; m++) {
        const float d = a.t[(size_t)i * a.Np + m] - y;
        const float I = d < 0.f ? 1.0f : 0.0f;
        // d hub / dy = -clamp(d, -1, 1)
        acc += fabsf(tau - I) * -fminf(fmaxf(d, -1.0f), 1.0f);
    }
    a.grad_y[idx] = g * acc;
}

} // namespace

extern "C" int b2rl_c51_loss_fwd(const float *y, const float *next_p, const float *reward,
                                 const float *discount, const float *terminal,
                                 const float *weights, const float *z, int32_t B, int32_t n_atoms,
                                 int mean, float *t_out, float *delta_out, float *scratch,
                                 float *loss_out, void *stream)
{
    B2RL_REQUIRE(y && next_p && reward && discount && terminal && z && t_out && delta_out &&
                     scratch && loss_out, B2RL_ERR_INVALID, "c51_loss_fwd: null argument");
    B2RL_REQUIRE(B > 0 && n_atoms >= 2 && n_atoms <= C51_MAX_ATOMS, B2RL_ERR_RANGE,
                 "c51_loss_fwd: need B > 0 and 2 <= n_atoms <= %d", C51_MAX_ATOMS);
    C51Args a{y, next_p, reward, discount, terminal, weights, z, B, n_atoms, mean,
              t_out, delta_out, scratch, loss_out};
    a.ticket = b2rl_next_ticket();
    k_c51_fwd<<<(B + C51_WARPS - 1) / C51_WARPS, C51_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
    B2RL_CUDA(cudaGetLastError());
    return B2RL_OK;
}

extern "C" int b2rl_c51_loss_bwd(const float *y, const float *t, const float *weights,
                                 const float *grad_loss, int32_t B, int32_t n_atoms, int mean,
                                 float *grad_y, void *stream)
{
    B2RL_REQUIRE(y && t && grad_loss && grad_y, B2RL_ERR_INVALID, "c51_loss_bwd: null argument");
    C51BwdArgs a{y, t, weights, grad_loss, B, n_atoms, mean, grad_y};
    const int total = B * n_atoms;
    k_c51_bwd<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    B2RL_CUDA(cudaGetLastError());
    return B2RL_OK;
}

extern "C" int b2rl_td_loss_fwd(const float *q, const int64_t *action, const float *next_q,
                                const float *reward, const float *discount,
                                const float *terminal, const float *weights, int32_t B,
                                int32_t n_actions, int clip_delta, int mean, float *y_out,
                                float *t_out, float *delta_out, float *scratch, float *loss_out,
                                void *stream)
{
    B2RL_REQUIRE(q && action && next_q && reward && discount && terminal && y_out && t_out &&
                     delta_out && scratch && loss_out, B2RL_ERR_INVALID,
                 "td_loss_fwd: null argument");
    B2RL_REQUIRE(B > 0 && n_actions > 0, B2RL_ERR_RANGE, "td_loss_fwd: empty batch");
    TdArgs a{q, (const long long *)action, next_q, reward, discount, terminal, weights, B,
             n_actions, clip_delta, mean, y_out, t_out, delta_out, scratch, loss_out};
    a.ticket = b2rl_next_ticket();
    k_td_fwd<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    B2RL_CUDA(cudaGetLastError());
    return B2RL_OK;
}

extern "C" int b2rl_td_loss_bwd(const float *y, const float *t, const float *weights,
                                const int64_t *action, const float *grad_loss, int32_t B,
                                int32_t n_actions, int clip_delta, int mean, float *grad_q,
                                void *stream)
{
    B2RL_REQUIRE(y && t && action && grad_loss && grad_q, B2RL_ERR_INVALID,
                 "td_loss_bwd: null argument");
    TdBwdArgs a{y, t, weights, grad_loss, (const long long *)action, B, n_actions, clip_delta,
                mean, grad_q};
    const int total = B * n_actions;
    k_td_bwd<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    B2RL_CUDA(cudaGetLastError());
    return B2RL_OK;
}

extern "C" int b2rl_quantile_huber_fwd(const float *y, const float *t, const float *taus,
                                       const float *weights, int32_t B, int32_t N, int32_t Np,
                                       int mean, float *delta_out, float *scratch,
                                       float *loss_out, void *stream)
{
    B2RL_REQUIRE(y && t && taus && delta_out && scratch && loss_out, B2RL_ERR_INVALID,
                 "quantile_huber_fwd: null argument");
    B2RL_REQUIRE(B > 0 && N > 0 && Np > 0, B2RL_ERR_RANGE, "quantile_huber_fwd: empty");
    QhArgs a{y, t, taus, weights, B, N, Np, mean, delta_out, scratch, loss_out};
    a.ticket = b2rl_next_ticket();
    k_qh_fwd<<<B, 256, 0, (cudaStream_t)stream>>>(a);
    B2RL_CUDA(cudaGetLastError());
    return B2RL_OK;
}

extern "C" int b2rl_quantile_huber_bwd(const float *y, const float *t, const float *taus,
                                       const float *weights, const float *grad_loss, int32_t B,
                                       int32_t N, int32_t Np, int mean, float *grad_y,
                                       void *stream)
{
    B2RL_REQUIRE(y && t && taus && grad_loss && grad_y, B2RL_ERR_INVALID,
                 "quantile_huber_bwd: null argument");
    QhBwdArgs a{y, t, taus, weights, grad_loss, B, N, Np, mean, grad_y};
    const int total = B * N;
    k_qh_bwd<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    B2RL_CUDA(cudaGetLastError());
    return B2RL_OK;
}
