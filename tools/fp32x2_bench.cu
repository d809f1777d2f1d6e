// Micro-benchmark: issue rate of Blackwell's packed FP32 instructions (FFMA2 / FADD2 / FMUL2) against the
// scalar FFMA / FADD on one B200.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_bench
// tools/fp32x2_bench.cu.  Prints warp-instructions per clock per SM sub-partition and the implied TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 8;

template <int MODE>
__global__ void __launch_bounds__(512) k(float2* out, int iters, float2 seed) {
    float2 a[kChains], b = seed, c = make_float2(seed.y, -seed.x);
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                if (MODE == 0) {            // scalar FFMA x2
                    a[i].x = fmaf(a[i].x, b.x, c.x);
                    a[i].y = fmaf(a[i].y, b.y, c.y);
                } else if (MODE == 1) {     // FFMA2
                    a[i] = __ffma2_rn(a[i], b, c);
                } else if (MODE == 2) {     // FADD2
                    a[i] = __fadd2_rn(a[i], c);
                } else if (MODE == 3) {     // scalar FADD x2
                    a[i].x += c.x; a[i].y += c.y;
                } else if (MODE == 4) {     // FFMA2 with a swapped / sign-patterned operand (complex rotate-add)
                    a[i] = __ffma2_rn(make_float2(a[i].y, -a[i].x), b, c);
                } else if (MODE == 5) {     // FMUL2
                    a[i] = __fmul2_rn(a[i], b);
                }
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kChains; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, int lanes_ops_per_inst, int inst_per_elem) {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int blocks = sms * 4, threads = 512, iters = 4000;
    float2* out;
    cudaMalloc(&out, sizeof(float2) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 10, make_float2(1.0001f, 0.9999f));
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, make_float2(1.0001f, 0.9999f));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = double(blocks) * threads / 32.0;
    const double inst = warps * iters * 8.0 * kChains * inst_per_elem;           // warp-instructions
    const double per_clk_smsp = inst / (ms * 1e-3) / (double(khz) * 1e3) / (sms * 4.0);
    const double tflops = inst * 32.0 * lanes_ops_per_inst / (ms * 1e-3) / 1e12;
    printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SMSP (at %d MHz nominal)  %7.2f TFLOP/s\n", name, ms, per_clk_smsp,
           khz / 1000, tflops);
    cudaFree(out);
}

int main() {
    run<0>("FFMA  (scalar, 2 per pair)", 2, 2);
    run<1>("FFMA2 (packed)", 4, 1);
    run<3>("FADD  (scalar, 2 per pair)", 1, 2);
    run<2>("FADD2 (packed)", 2, 1);
    run<5>("FMUL2 (packed)", 2, 1);
    run<4>("FFMA2 swapped+negated operand", 4, 1);
    return 0;
}
