// Single-warp latency probes on sm_100a (tools, not product): dependent-issue latency of
// the instructions on the exact sampler's critical path.  nvcc -arch=sm_100a -O3 lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }

__global__ void probe(double *out, long long *cyc, int iters)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    for (int i = lane; i < 4096; i += 32) sm[i] = 1.0 + i * 1e-9;
    __syncwarp();
    // 1. dependent DADD chain
    double x = 1.0 + lane;
    long long t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        x = __dadd_rn(x, 1.0); x = __dadd_rn(x, 1.0); x = __dadd_rn(x, 1.0); x = __dadd_rn(x, 1.0);
        x = __dadd_rn(x, 1.0); x = __dadd_rn(x, 1.0); x = __dadd_rn(x, 1.0); x = __dadd_rn(x, 1.0);
    }
    long long t1 = clk();
    cyc[0] = (t1 - t0);
    // 2. dependent DFMA chain
    t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        x = __fma_rn(x, 1.0000001, 0.5); x = __fma_rn(x, 0.9999999, 0.5); x = __fma_rn(x, 1.0000001, 0.5); x = __fma_rn(x, 0.9999999, 0.5);
        x = __fma_rn(x, 1.0000001, 0.5); x = __fma_rn(x, 0.9999999, 0.5); x = __fma_rn(x, 1.0000001, 0.5); x = __fma_rn(x, 0.9999999, 0.5);
    }
    t1 = clk();
    cyc[1] = (t1 - t0);
    // 3. dependent LDS chain (pointer chasing through shared memory, 64-bit)
    int idx = lane;
    t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) idx = (int)(sm[idx & 4095] * 0.0) + ((idx * 5 + 1) & 4095);
    }
    t1 = clk();
    cyc[2] = (t1 - t0);
    // 4. STS followed by a dependent-address LDS (same warp), 8 pairs
    t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            sm[2048 + ((idx + j) & 1023)] = x;
            idx = (int)(sm[idx & 1023] * 0.0) + ((idx * 5 + 1) & 1023);
        }
    }
    t1 = clk();
    cyc[3] = (t1 - t0);
    // 5. ballot + popc + shfl round
    unsigned acc = lane;
    t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            unsigned b = __ballot_sync(0xffffffffu, (acc + j) & 1);
            acc = __shfl_sync(0xffffffffu, acc + __popc(b), (acc + 1) & 31);
        }
    }
    t1 = clk();
    cyc[4] = (t1 - t0);
    // 6. DSETP -> select -> DADD (compare feeding a predicate feeding an add)
    t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) x = (x < 1e300) ? __dadd_rn(x, 1.0) : x;
    }
    t1 = clk();
    cyc[5] = (t1 - t0);
    // 7. LDS.64 -> DADD -> STS.64 -> LDS.64 same address (read-modify-write chain)
    t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            volatile double *p = sm + 3000 + lane;
            *p = *p - 1e-3;
        }
    }
    t1 = clk();
    cyc[6] = (t1 - t0);
    // 8. ld.acquire.cta.shared + st.release.cta.shared pair
    int *fl = reinterpret_cast<int *>(sm + 4000);
    t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int v;
            unsigned a = (unsigned)__cvta_generic_to_shared(fl);
            asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
            if (lane == 0) asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(a), "r"(v + 1) : "memory");
            acc += v;
        }
    }
    t1 = clk();
    cyc[7] = (t1 - t0);
    // 9. dependent DMUL chain
    t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) x = __dmul_rn(x, 1.0000000001);
    }
    t1 = clk();
    cyc[8] = (t1 - t0);
    // 10. FADD dependent chain (for reference)
    float f = (float)x;
    t0 = clk();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) f = __fadd_rn(f, 1.0f);
    }
    t1 = clk();
    cyc[9] = (t1 - t0);
    out[lane] = x + idx + acc + f;
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
    const int iters = 2000;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; rep++) probe<<<1, 64, 40 * 1024>>>(out, cyc, iters);
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char *names[] = {"DADD chain", "DFMA chain", "LDS chase", "STS+LDS", "ballot+popc+shfl", "DSETP+sel+DADD",
                           "LDS-DADD-STS same addr", "ld.acquire+st.release", "DMUL chain", "FADD chain"};
    for (int i = 0; i < 10; i++) printf("%-26s %7.1f cycles per op\n", names[i], (double)h[i] / (iters * 8.0));
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
