// Micro-benchmark: Tensor Memory (TMEM) as a per-lane read-only table store.  Question behind it: can the twiddle and
// mel-weight tables of frames_fast_2048 (166 of its 655 shared-memory wavefronts per frame) be read with tcgen05.ld
// instead of LDS, taking that traffic off the shared-memory data pipe?
//   * correctness: tcgen05.alloc -> tcgen05.st (one warp per 32-lane quarter) -> tcgen05.ld from all 16 warps
//   * throughput of tcgen05.ld.32x32b.x16 with 16 warps per SM, alone and next to a stream of LDS.128
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_lut_bench tools/tmem_lut_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int NCOL = 192;          // columns used (of 256 allocated)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// mode 0: tcgen05.ld only; 1: LDS.128 only (same bytes); 2: both
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, const float* __restrict__ table, int iters) {
    __shared__ uint32_t s_taddr;
    __shared__ __align__(16) float s_tab[32 * NCOL];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * NCOL; i += blockDim.x) s_tab[i] = table[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_taddr)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = s_taddr;
    if (warp < 4) {      // warp q fills lanes 32q .. 32q+31 (every quarter holds the same 32 x NCOL table)
        for (int c = 0; c < NCOL; c += 4) {
            const float4 v = *reinterpret_cast<const float4*>(table + lane * NCOL + c);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tbase + ((uint32_t)(32 * warp) << 16) + c),
                         "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    const uint32_t tq = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    float acc = 0.0f, acc2 = 0.0f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int c = 0; c < NCOL; c += 16) {
            if (MODE == 0 || MODE == 2) {
                float v[16];
                tmem_ld16(tq + c, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += v[i];
            }
            if (MODE == 1 || MODE == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // lane-contiguous float4 rows as the kernel's weight table: [c/4 + j][lane]
                    const float4 w = *reinterpret_cast<const float4*>(s_tab + ((c / 4 + j) * 32 + lane) * 4);
                    acc2 += (w.x + w.y) + (w.z + w.w);
                }
            }
        }
    }
    out[blockIdx.x * blockDim.x + tid] = acc + acc2;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}

template <int MODE>
static void run(const char* name, const float* d_table, const std::vector<float>& h, int sms, int khz) {
    const int blocks = sms, threads = 512, iters = 2000;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    k<MODE><<<blocks, threads>>>(out, d_table, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
    // correctness of one pass
    std::vector<float> got((size_t)blocks * threads);
    cudaMemcpy(got.data(), out, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < blocks * threads; ++t) {
        const int lane = t & 31;
        double want = 0, want2 = 0;
        for (int c = 0; c < NCOL; ++c) want += h[lane * NCOL + c];
        for (int r = 0; r < NCOL / 4; ++r) for (int i = 0; i < 4; ++i) want2 += h[(r * 32 + lane) * 4 + i];
        double w = (MODE == 0 ? want : MODE == 1 ? want2 : want + want2);
        if (fabs(got[t] - w) > 1e-3 * fabs(w) + 1e-3) ++bad;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, d_table, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 * khz * 1e3;
    const double ops = 16.0 * iters * (NCOL / 16);          // per SM: 16 warps x chunks of 16 columns (2 KB per warp each)
    printf("%-28s %8.3f ms  mismatches %d  %.1f clk per 2 KB warp-chunk per SM  = %.1f B/clk/SM per path\n", name, ms, bad,
           clk / ops, 2048.0 * ops / clk);
    cudaFree(out);
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    std::vector<float> h(32 * NCOL);
    for (size_t i = 0; i < h.size(); ++i) h[i] = float((i * 2654435761u) % 1000) * 1e-3f;
    float* d;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    run<0>("tcgen05.ld.32x32b.x16", d, h, sms, khz);
    run<1>("LDS.128 x4 (same bytes)", d, h, sms, khz);
    run<2>("both", d, h, sms, khz);
    return 0;
}
