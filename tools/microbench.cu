// tools/microbench.cu -- pipe-rate probes on the GPU box: FFMA, FFMA2 (fma.rn.f32x2), DFMA, mma.sync tf32/bf16.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/microbench tools/microbench.cu && tools/bin/microbench
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048

__global__ void k_ffma(float *out, float s) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    float x = s, y = s * 0.5f;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    }
    float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// rank-1 style: 8x... distinct operands a[r][c] += x[r]*y[c]
__global__ void k_ffma_r1(float *out, float s) {
    float a[4][8], x[4], y[8];
    for (int i = 0; i < 4; ++i) { x[i] = s + i; for (int j = 0; j < 8; ++j) a[i][j] = threadIdx.x + i * j; }
    for (int j = 0; j < 8; ++j) y[j] = s * j;
    for (int it = 0; it < ITERS / 2; ++it) {
        #pragma unroll
        for (int i = 0; i < 4; ++i)
            #pragma unroll
            for (int j = 0; j < 8; ++j) a[i][j] = fmaf(x[i], y[j], a[i][j]);
        x[it & 3] += 1e-9f;
    }
    float r = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) r += a[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_ffma2(float *out, float s) {
    unsigned long long a[16], x, y;
    for (int i = 0; i < 16; ++i) { float lo = threadIdx.x * 0.001f + i, hi = lo + 1; asm("mov.b64 %0, {%1,%2};" : "=l"(a[i]) : "f"(lo), "f"(hi)); }
    asm("mov.b64 %0, {%1,%2};" : "=l"(x) : "f"(s), "f"(s));
    asm("mov.b64 %0, {%1,%2};" : "=l"(y) : "f"(s * 0.5f), "f"(s * 0.25f));
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x), "l"(y));
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); r += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// rank-1 style FFMA2: pairs along rows, y broadcast as a duplicated pair
__global__ void k_ffma2_r1(float *out, float s) {
    unsigned long long a[2][8], x[2], y[8];
    for (int i = 0; i < 2; ++i) {
        float lo = s + i, hi = s - i; asm("mov.b64 %0, {%1,%2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
        for (int j = 0; j < 8; ++j) { float l2 = threadIdx.x + i * j, h2 = l2 + 1; asm("mov.b64 %0, {%1,%2};" : "=l"(a[i][j]) : "f"(l2), "f"(h2)); }
    }
    for (int j = 0; j < 8; ++j) { float v = s * j; asm("mov.b64 %0, {%1,%2};" : "=l"(y[j]) : "f"(v), "f"(v)); }
    for (int it = 0; it < ITERS / 2; ++it) {
        #pragma unroll
        for (int i = 0; i < 2; ++i)
            #pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i][j]) : "l"(x[i]), "l"(y[j]));
    }
    float r = 0;
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 8; ++j) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i][j])); r += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_dfma(float *out, float s) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001 + i;
    double x = s, y = s * 0.5;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double r = 0; for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)r;
}
__global__ void k_mma_tf32(float *out, float s) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = __float_as_uint(s), a1 = a0, a2 = a0, a3 = a0, b0 = a0, b1 = a0;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float r = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) r += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_mma_bf16(float *out, float s) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = __float_as_uint(s), a1 = a0, a2 = a0, a3 = a0, b0 = a0, b1 = a0;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float r = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) r += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_mma_f64(float *out, float s) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 2; ++j) c[i][j] = 0.;
    double a = s, b = s;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double r = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 2; ++j) r += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)r;
}

__global__ void k_shfl(float *out, float s) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    const int src = (threadIdx.x + 5) & 31;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __shfl_sync(0xffffffffu, a[i], src);
    }
    float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// shuffles interleaved with FFMA2 (2 : 1), as in a shuffle-broadcast rank-1 update
__global__ void k_shfl_ffma2(float *out, float s) {
    float a[8];
    unsigned long long acc[16], x;
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 16; ++i) { float lo = threadIdx.x + i, hi = lo + 1; asm("mov.b64 %0, {%1,%2};" : "=l"(acc[i]) : "f"(lo), "f"(hi)); }
    asm("mov.b64 %0, {%1,%2};" : "=l"(x) : "f"(s), "f"(s));
    const int src = (threadIdx.x + 5) & 31;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = __shfl_sync(0xffffffffu, a[i], src);
            unsigned long long y; asm("mov.b64 %0, {%1,%2};" : "=l"(y) : "f"(a[i]), "f"(a[i]));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[2 * i]) : "l"(x), "l"(y));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[2 * i + 1]) : "l"(x), "l"(y));
        }
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); r += lo + hi; }
    for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename K>
static void run(const char *name, K kern, double ops_per_thread_iter, int threads, int blocks_per_sm) {
    int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    float *out; cudaMalloc(&out, sizeof(float) * sms * blocks_per_sm * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<sms * blocks_per_sm, threads>>>(out, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); kern<<<sms * blocks_per_sm, threads>>>(out, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double total = ops_per_thread_iter * (double)ITERS * threads * blocks_per_sm * sms;
    printf("%-12s thr=%4d cta/sm=%d  %8.3f ms  %8.2f Tops/s  %7.1f ops/clk/SM (at %d MHz nominal)  err=%s\n", name, threads, blocks_per_sm, best,
           total / best / 1e9, total / (best * 1e-3) / sms / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int bps = 1; bps <= 2; ++bps) {
        run("ffma", k_ffma, 16, 512, bps);           // ops = FMAs
        run("ffma_r1", k_ffma_r1, 16, 512, bps);
        run("ffma2", k_ffma2, 32, 512, bps);
        run("ffma2_r1", k_ffma2_r1, 16, 512, bps);
        run("dfma", k_dfma, 16, 512, bps);
        run("mma_tf32", k_mma_tf32, 8 * 16 * 8 * 8 / 32.0, 512, bps);   // FMAs per thread
        run("mma_bf16", k_mma_bf16, 8 * 16 * 8 * 16 / 32.0, 512, bps);
        run("mma_f64", k_mma_f64, 8 * 8 * 8 * 4 / 32.0, 512, bps);
        run("shfl (lanes)", k_shfl, 16, 512, bps);          // ops = lane-shuffles: 32 per warp instruction
        run("shfl+2ffma2", k_shfl_ffma2, 8, 512, bps);     // ops = lane-shuffles, two FFMA2 per shuffle alongside
    }
    return 0;
}
