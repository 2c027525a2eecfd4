// Measures, on the GPU it runs on, the issue rate of scalar FP32 FFMA/FMUL/FADD against the packed
// f32x2 forms (FFMA2 ...) that sm_100 adds -- the rasterisers are FP32-issue bound and process two
// pixels per thread, so whether a packed instruction costs one scalar slot or two decides the design.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_rate fp32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
#define NACC 8

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) {
    u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int MODE>
__global__ void __launch_bounds__(256) rate(float *out, float x, float y) {
    float a[NACC];
    u64 p[NACC];
    const float t = threadIdx.x * 1e-3f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { a[i] = t + i; p[i] = pk(t + i, t - i); }
    const u64 xx = pk(x, x * 1.0001f), yy = pk(y, y * 0.9999f);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], x, y);                       // FFMA, 3 register operands
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(xx), "l"(yy));
            if (MODE == 2) a[i] = __fmul_rn(a[i], x);
            if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(xx));
            if (MODE == 4) a[i] = __fadd_rn(a[i], x);
            if (MODE == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(xx));
            if (MODE == 6) {  // FFMA interleaved with an integer-pipe op (is the slot shared?)
                a[i] = fmaf(a[i], x, y);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(unsigned *)&p[i]) : "r"(it), "r"(i));
            }
            if (MODE == 7) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(xx), "l"(yy));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(unsigned *)&a[i]) : "r"(it), "r"(i));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += a[i] + lo(p[i]);
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
static void run(const char *name, int lanes_per_inst, int sms, float *d) {
    const int blocks = sms * 8;
    rate<MODE><<<blocks, 256>>>(d, 1.0001f, 1e-7f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate<MODE><<<blocks, 256>>>(d, 1.0001f, 1e-7f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst = (double)blocks * 8 * ITER * NACC;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("{\"mode\": \"%s\", \"ms\": %.4f, \"warp_inst_per_clk_per_smsp_at_max_clock\": %.3f, "
           "\"fp32_lane_ops_per_clk_per_sm\": %.1f}\n",
           name, ms, warp_inst / cycles / (sms * 4), warp_inst * 32 * lanes_per_inst / cycles / sms);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    float *d;
    cudaMalloc(&d, 4);
    const int sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    run<0>("FFMA r,r,r", 1, sms, d);
    run<1>("FFMA2 (fma.rn.f32x2)", 2, sms, d);
    run<2>("FMUL", 1, sms, d);
    run<3>("FMUL2 (mul.rn.f32x2)", 2, sms, d);
    run<4>("FADD", 1, sms, d);
    run<5>("FADD2 (add.rn.f32x2)", 2, sms, d);
    run<6>("FFMA + LOP3 interleaved (2 inst per step)", 1, sms, d);
    run<7>("FFMA2 + LOP3 interleaved (2 inst per step)", 2, sms, d);
    return 0;
}
