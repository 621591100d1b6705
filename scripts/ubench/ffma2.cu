// Micro-benchmark: issue cost of packed fp32 (fma.rn.f32x2 / add.f32x2 / mul.f32x2, sm_100a) against scalar FFMA,
// alone and mixed with ALU (select / compare) and MUFU work -- does FFMA2 free issue slots in an issue-bound kernel?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned long long p0, p1, p2, p3, pa, pb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(x2), "f"(x3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(x4), "f"(x5));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(x6), "f"(x7));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
    float s0 = 0.f, s1 = 1.f, s2 = 2.f, s3 = 3.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0 || MODE == 2 || MODE == 4) {  // 8 scalar FFMA = 8 fp32 results
                x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
                x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
            } else {  // 4 packed FFMA2 = 8 fp32 results
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
            }
            if (MODE == 2 || MODE == 3) {  // + 4 ALU-pipe ops (FMNMX)
                s0 = fmaxf(s0, s1 + 0.f * 0.f); s1 = fminf(s1, s2); s2 = fmaxf(s2, s3); s3 = fminf(s3, s0);
            }
            if (MODE == 4 || MODE == 5) {  // + 2 MUFU ops
                s0 = __frcp_rn(s0) ; s1 = __expf(s1);
            }
        }
    }
    float lo, hi, acc = s0 + s1 + s2 + s3 + x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p0)); acc += lo + hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p1)); acc += lo + hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p2)); acc += lo + hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p3)); acc += lo + hi;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static void run(const char *name, float *out) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 512>>>(out, 16, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<148, 512>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * 512 * (double)iters * 8 * 8;
    printf("%-34s %.3f ms  %.1f GFMA/s  (%.1f fp32 results / clk / SM at 1.9 GHz)\n", name, ms, fma / ms / 1e6,
           fma / ms / 1e6 / 148 / 1.9);
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 512 * 4);
    run<0>("scalar FFMA", out);
    run<1>("packed FFMA2", out);
    run<2>("scalar FFMA + 4 ALU / 8", out);
    run<3>("packed FFMA2 + 4 ALU / 8", out);
    run<4>("scalar FFMA + 2 MUFU / 8", out);
    run<5>("packed FFMA2 + 2 MUFU / 8", out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
