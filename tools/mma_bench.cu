// Micro-benchmark: issue rate of the warp-level tensor-core instruction mma.sync.m16n8k8 (TF32, SASS HMMA.1688.F32.TF32)
// on one B200, alone and interleaved with packed FP32 (FFMA2) from the same warps - the question behind moving the
// banded mel projection of frames_fast_2048 and the DCT of db_dct to the (otherwise idle) tensor pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int kChains = 4;

// MODE 0: TF32 mma only; 1: bf16 mma only; 2: TF32 mma + 8 FFMA2 per mma; 3: the 8 FFMA2 alone; 4: TF32 + 2 FFMA2 per mma
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed) {
    float d[kChains][4];
    uint32_t a[4], b[2];
    float2 f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(seed * (threadIdx.x + i + 1));
    b[0] = __float_as_uint(seed * 3.0f); b[1] = __float_as_uint(seed * 5.0f);
#pragma unroll
    for (int c = 0; c < kChains; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[c][i] = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = make_float2(seed * i, seed);
    const float2 m = make_float2(1.0001f, 0.9999f), ad = make_float2(seed, -seed);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) {
                if (MODE == 0 || MODE == 2 || MODE == 4) mma_tf32(d[c], a, b);
                if (MODE == 1) mma_bf16(d[c], a, b);
                if (MODE == 2 || MODE == 3) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = __ffma2_rn(f[i], m, ad);
                }
                if (MODE == 4) {
                    f[2 * c] = __ffma2_rn(f[2 * c], m, ad);
                    f[2 * c + 1] = __ffma2_rn(f[2 * c + 1], m, ad);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += d[c][i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i].x + f[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double mma_per_it, double ffma2_per_it, double flop_per_mma) {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int blocks = sms, threads = 512, iters = 4000;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 10, 1e-3f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1e-3f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = double(blocks) * threads / 32.0;
    const double clk = (ms * 1e-3) * (double(khz) * 1e3);
    const double mma = warps * iters * mma_per_it, ff = warps * iters * ffma2_per_it;
    printf("%-34s %8.3f ms  mma %.4f /clk/SMSP (%.2f clk per mma per SMSP)  ffma2 %.3f /clk/SMSP  tensor %.1f TFLOP/s\n", name, ms,
           mma / clk / (sms * 4.0), mma > 0 ? clk * sms * 4.0 / mma : 0.0, ff / clk / (sms * 4.0),
           mma * flop_per_mma / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    run<0>("mma.m16n8k8.tf32 alone", 16, 0, 2.0 * 16 * 8 * 8);
    run<1>("mma.m16n8k16.bf16 alone", 16, 0, 2.0 * 16 * 8 * 16);
    run<3>("8 FFMA2 alone", 0, 128, 0);
    run<2>("tf32 mma + 8 FFMA2 each", 16, 128, 2.0 * 16 * 8 * 8);
    run<4>("tf32 mma + 2 FFMA2 each", 16, 32, 2.0 * 16 * 8 * 8);
    return 0;
}
