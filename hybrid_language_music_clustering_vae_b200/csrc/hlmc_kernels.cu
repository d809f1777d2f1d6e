// Hand-written sm_100a kernels for the audio feature hot path.
//
//   frames_fast_2048   persistent, one warp per frame: TMA-bulk-staged waveform
//                      -> pad + window -> 1024-point complex FFT in registers
//                      (two 32-point passes, one shared-memory transpose)
//                      -> real-FFT split -> |X|^2, |X| -> centroid / bandwidth /
//                      rolloff in registers -> banded mel projection -> staged,
//                      coalesced store of the mel-power tile + per-clip max;
//                      ZCR and RMS from the same staged samples.
//   frames_generic     same outputs for any power-of-two n_fft (shared-memory
//                      radix-2 FFT); also the librosa.stft entry point.
//   db_dct             power_to_db (ref=max / top_db per clip) fused with the
//                      orthonormal DCT-II (MFCC).
//   pool / fix_frames  the scripts' time pooling and crop / pad-with-min.
//
// librosa semantics follow SURVEY.md Appendix A; each kernel names the call it
// replaces.  No cuFFT / cuBLAS / Thrust anywhere.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <atomic>

#include "fft_inreg.cuh"
#include "fft_inreg2.cuh"
#include "hlmc_internal.h"

namespace hlmc {

static std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches += n; }
constexpr long long kMaxGridY = 65535;

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
#define FULL 0xffffffffu

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// Sixteen warp sums for the price of 1.6: a reduce-scatter butterfly.  On return lane l holds the warp total of
// v[l >> 1] (both lanes of a pair hold the same value); the pairings and their order are those of warp_sum, so
// each total is bitwise the one warp_sum would give.
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {
        const bool up = (lane & (2 * w)) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const float send = up ? v[i] : v[i + w];
            const float keep = up ? v[i + w] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, 2 * w);
        }
    }
    return v[0] + __shfl_xor_sync(FULL, v[0], 1);
}

// np.pad index maps (librosa.stft / rms pad_mode; zero_crossing_rate is always "edge").
__device__ __forceinline__ int reflect_index(int s, int n) {
    if (n == 1) return 0;
    int period = 2 * (n - 1);
    s %= period;
    if (s < 0) s += period;
    return (s < n) ? s : period - s;
}
__device__ __forceinline__ float sample_padded(const float* __restrict__ clip, int n, int s,
                                               int pad_mode) {
    if (s >= 0 && s < n) return __ldg(clip + s);
    if (pad_mode == 0) return 0.0f;
    if (pad_mode == 1) return __ldg(clip + reflect_index(s, n));
    return __ldg(clip + min(max(s, 0), n - 1));
}
__device__ __forceinline__ float sample_edge(const float* __restrict__ clip, int n, int s) {
    return __ldg(clip + min(max(s, 0), n - 1));
}

// ---------------------------------------------------------------------------
// Fast path: n_fft = 2048.  Persistent CTAs of NW fully independent warps; each
// warp walks a contiguous run of (clip, frame) pairs and owns one shared-memory
// buffer that is first the TMA landing zone for the frame's 2048 samples and
// then the scratch for the FFT transposes and the mel gather.  There is no
// CTA-wide barrier after setup.
// ---------------------------------------------------------------------------
constexpr int kWarpBufFloats = 2 * (1024 + 64 + 1) + 2;
constexpr int kWarpBufFloats4 = (kWarpBufFloats + 3) & ~3;   // padded 1024-bin complex spectrum (>= 32x33 transpose tile)

__host__ __device__ constexpr int pos32(int k) { return fftreg::fft_pos<32>(k); }

// PREF kernels keep the TMA landing zone apart from the scratch of the last phases, so that the next
// frame's samples can be in flight while this frame finishes: [0, kLandOff) power-spectrum scratch,
// [kLandOff, kLandOff + 2048 + 4) landing zone; the transposes run over [0, 2180) once the samples are
// in registers.
constexpr int kLandOff = 1108;
constexpr int kPrefBufFloats = kLandOff + kFastNfft + 4;

struct FastSmemLayout {
    int bars, tables, ntab, wbuf, wb, total;   // float offsets; ntab = table floats staged; wb = floats per warp buffer
};
__host__ __device__ inline FastSmemLayout fast_layout(const FastTables& ft, int nw, bool pref, bool tm = false) {
    FastSmemLayout L;
    L.bars = 0;                                   // nw mbarriers, 8 B each (TM kernels: + the TMEM base address)
    L.tables = (2 * nw + 3 + (tm ? 4 : 0)) & ~3;
    L.ntab = tm ? 0 : pref ? ft.nowin : ft.total; // PREF kernels synthesise the Hann window: no table; TM kernels: all tables in TMEM
    L.wbuf = L.tables + L.ntab;
    const int scr = (ft.scr > kWarpBufFloats) ? ft.scr : kWarpBufFloats;
    L.wb = pref ? kPrefBufFloats : ((scr + 3) & ~3);
    L.total = L.wbuf + nw * L.wb;
    return L;
}
int fast_smem_bytes(const FastTables& ft, int nwarps, int n_fft, int hop, int n_mels) {
    (void)n_fft; (void)hop; (void)n_mels;
    return fast_layout(ft, nwarps, false).total * 4;
}
int pick_fast_warps(int T) { (void)T; return 16; }

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------
// Tensor Memory as a per-lane table store (TM kernels).  The per-lane read-only tables of the frame pipeline -
// window, inter-pass twiddles, split twiddles, banded mel weights, gather start offsets - cost 166 of the 655
// shared-memory wavefronts per frame when they are read with LDS, on the pipe that bounds the kernel.  TMEM is
// a second on-chip memory with its own read path (measured: 256 B/clk/SM, and concurrent with LDS traffic at
// full rate - tools/tmem_lut_bench.cu): thread i of a warp reads N consecutive 32-bit columns of TMEM lane
// 32*(warp%4) + i with ONE tcgen05.ld.32x32b.xN.  Every 32-lane quarter holds the same 32 x kTmCols table.
// Column map (floats, per lane):
// (kTmWin ... kTmAlloc: hlmc_internal.h)

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// The loaded registers are valid only after tcgen05.wait::ld; naming them as in/out operands keeps every use behind it.
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
        : "r"(taddr));
}
__device__ __forceinline__ float2 f2_of(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// CTA-wide: allocate the SM's Tensor Memory (all 512 columns: one CTA per SM), fill the four 32-lane quarters with
// the per-lane tables tmem_tab [32][tmem_cols] (quarter warp%4 by its NW/4 warps, each a contiguous share of
// the columns, 4 loads in flight) and return the base address.  Ends with the fence half of a CTA barrier: the
// caller's next __syncthreads() + tcgen05.fence::after_thread_sync publishes the tables.
template <int NW>
__device__ __forceinline__ uint32_t tmem_tables_setup(const float* __restrict__ tmem_tab, int tmem_cols, uint32_t* s_taddr,
                                                      int warp, int lane) {
    static_assert(NW % 4 == 0, "a whole number of warps per TMEM quarter");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_taddr)), "r"(kTmAlloc));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = *s_taddr;
    const float4* src = reinterpret_cast<const float4*>(tmem_tab + (size_t)lane * tmem_cols);
    const int n4 = tmem_cols / 4, per = (n4 + NW / 4 - 1) / (NW / 4);
    const int c0 = (warp >> 2) * per, c1 = min(n4, c0 + per);
    const uint32_t tqw = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    for (int c = c0; c < c1; c += 4) {
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = (c + u < c1) ? __ldg(src + c + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c + u < c1)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tqw + 4 * (c + u)),
                             "r"(__float_as_uint(q[u].x)), "r"(__float_as_uint(q[u].y)), "r"(__float_as_uint(q[u].z)),
                             "r"(__float_as_uint(q[u].w)));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    return tbase;
}
// CTA-wide, at the end of the kernel (every warp must get here): free the Tensor Memory.
__device__ __forceinline__ void tmem_tables_release(uint32_t tbase, int warp) {
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(kTmAlloc));
}

// Banded mel gather of one (group, lane): n4 float4 steps of 4 taps each.  The step count is
// warp-uniform but only known at run time.  Steps are taken in straight-line chunks of 4, 2 and 1 so
// that all shared-memory loads of a chunk are in flight before its first multiply-add (a step-by-step
// loop pays the full load latency 26 times per frame; ncu showed it as the top short-scoreboard stall).
template <int CH, int WS>
__device__ __forceinline__ void mel_chunk(const float4* __restrict__ wp, const float* __restrict__ pp,
                                          float2& a01, float2& a23) {
    float4 w[CH];
    float2 p01[CH], p23[CH];
#pragma unroll
    for (int s = 0; s < CH; ++s) {
        w[s] = wp[WS * s];
        p01[s] = make_float2(pp[4 * s + 0], pp[4 * s + 1]);
        p23[s] = make_float2(pp[4 * s + 2], pp[4 * s + 3]);
    }
#pragma unroll
    for (int s = 0; s < CH; ++s) {
        a01 = __ffma2_rn(make_float2(w[s].x, w[s].y), p01[s], a01);
        a23 = __ffma2_rn(make_float2(w[s].z, w[s].w), p23[s], a23);
    }
}
// WS = float4 stride between a lane's consecutive steps (lanes per filter round: 32, or L in frames_sub)
template <int WS = 32>
__device__ __forceinline__ void mel_steps(int n4, const float4* __restrict__ wp, const float* __restrict__ pp,
                                          float2& a01, float2& a23) {
    for (; n4 >= 4; n4 -= 4, wp += 4 * WS, pp += 16) mel_chunk<4, WS>(wp, pp, a01, a23);
    if (n4 & 2) { mel_chunk<2, WS>(wp, pp, a01, a23); wp += 2 * WS; pp += 8; }
    if (n4 & 1) mel_chunk<1, WS>(wp, pp, a01, a23);
}

// Mel gather with the weights in Tensor Memory: CH steps = one tcgen05.ld of 4*CH columns; the power-spectrum
// loads of the chunk are in flight while it lands.
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
template <int CH>
__device__ __forceinline__ void mel_chunk_tm(uint32_t taddr, const float* __restrict__ pp, float2& a01, float2& a23) {
    uint32_t wr[16];
    if constexpr (CH == 4) tmem_ld16_issue(taddr, wr);
    else if constexpr (CH == 2) tmem_ld8_issue(taddr, wr);
    else tmem_ld4_issue(taddr, wr);
    float2 p01[CH], p23[CH];
#pragma unroll
    for (int u = 0; u < CH; ++u) {
        p01[u] = make_float2(pp[4 * u + 0], pp[4 * u + 1]);
        p23[u] = make_float2(pp[4 * u + 2], pp[4 * u + 3]);
    }
    if constexpr (CH == 4) tmem_wait16(wr);
    else {
        // (only the registers the load wrote are named: the others hold nothing)
        if constexpr (CH == 2)
            asm volatile("tcgen05.wait::ld.sync.aligned;"
                         : "+r"(wr[0]), "+r"(wr[1]), "+r"(wr[2]), "+r"(wr[3]), "+r"(wr[4]), "+r"(wr[5]), "+r"(wr[6]), "+r"(wr[7]));
        else
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(wr[0]), "+r"(wr[1]), "+r"(wr[2]), "+r"(wr[3]));
    }
#pragma unroll
    for (int u = 0; u < CH; ++u) {
        a01 = __ffma2_rn(f2_of(wr[4 * u], wr[4 * u + 1]), p01[u], a01);
        a23 = __ffma2_rn(f2_of(wr[4 * u + 2], wr[4 * u + 3]), p23[u], a23);
    }
}
// n4 steps of one group; `col` = the group's first TMEM column
__device__ __forceinline__ void mel_steps_tm(int n4, uint32_t taddr, const float* __restrict__ pp, float2& a01, float2& a23) {
    for (; n4 >= 4; n4 -= 4, taddr += 16, pp += 16) mel_chunk_tm<4>(taddr, pp, a01, a23);
    if (n4 & 2) { mel_chunk_tm<2>(taddr, pp, a01, a23); taddr += 8; pp += 8; }
    if (n4 & 1) mel_chunk_tm<1>(taddr, pp, a01, a23);
}

// The default plan (128 Slaney mels at 22.05 kHz: 3, 3, 7, 14 float4 steps for the four filter groups) gets its step
// counts at compile time: every trip count and TMEM column is an immediate (measured: 3.80 against 4.03 ms).
constexpr int kMelUnrDefault = 3 | (3 << 8) | (7 << 16) | (14 << 24);
template <int UNR, int G>
__device__ __forceinline__ float mel_group_tm(uint32_t tq, const float* __restrict__ pp) {
    constexpr int STEPS = (UNR >> (8 * G)) & 255;
    constexpr int COL = kTmMel + 4 * ((G > 0 ? (UNR & 255) : 0) + (G > 1 ? ((UNR >> 8) & 255) : 0) + (G > 2 ? ((UNR >> 16) & 255) : 0));
    float2 a01 = make_float2(0.0f, 0.0f), a23 = a01;
#pragma unroll
    for (int c = 0; c < STEPS / 4; ++c) mel_chunk_tm<4>(tq + COL + 16 * c, pp + 16 * c, a01, a23);
    if constexpr ((STEPS & 3) == 3) {
        mel_chunk_tm<2>(tq + COL + 16 * (STEPS / 4), pp + 16 * (STEPS / 4), a01, a23);
        mel_chunk_tm<1>(tq + COL + 16 * (STEPS / 4) + 8, pp + 16 * (STEPS / 4) + 8, a01, a23);
    }
    if constexpr ((STEPS & 3) == 2) mel_chunk_tm<2>(tq + COL + 16 * (STEPS / 4), pp + 16 * (STEPS / 4), a01, a23);
    if constexpr ((STEPS & 3) == 1) mel_chunk_tm<1>(tq + COL + 16 * (STEPS / 4), pp + 16 * (STEPS / 4), a01, a23);
    a01 = __fadd2_rn(a01, a23);
    return a01.x + a01.y;
}

struct WarpState {
    float* sc;                 // this warp's shared buffer (TMA landing zone, then scratch)
    float2* sc2;
    uint64_t* mbar;
    uint32_t parity;
    const float2 *s_win, *s_tw1, *s_tw2;
    const float4* s_hcs;       // per-lane (cos, cos', sin, sin') of the Hann phase (PREF kernels)
    int lane, zw_base, zlo_base, zhi_base, zhi0;
    uint32_t tq;               // TMEM address of this warp's 32-lane quarter (TM kernels)
    bool pending;              // a TMA copy of the next frame to process is in flight (PREF kernels)
    int pend_off;              // its alignment shift
};

// Start the TMA bulk copy of frame t of `clip` into `land` if the frame holds no padding and is 8-byte
// aligned (16-byte aligned source, 0..3 leading floats).  Returns the shift, or -1 if the frame has to be
// built by hand.
__device__ __forceinline__ int issue_frame_tma(const FrameArgs& a, const WarpState& w, float* land,
                                               const float* clip, int t, int nfft = kFastNfft) {
    const int fs = t * a.hop - a.pad;
    const bool interior = (fs >= 0) && (fs + nfft <= a.n);
    const uintptr_t addr = reinterpret_cast<uintptr_t>(clip + fs);
    const int shift = (int)((addr & 15) >> 2);
    if (!interior || (shift & 1)) return -1;
    if (w.lane == 0) {
        const float* src0 = reinterpret_cast<const float*>(addr & ~uintptr_t(15));
        const int tot = shift + nfft;
        const int bulk = tot & ~3;
        for (int i = bulk; i < tot; ++i) land[i] = __ldg(src0 + i);     // <= 3 tail floats
        fence_proxy_async_smem();       // order earlier generic accesses before the async write
        mbar_arrive_expect_tx(w.mbar, (uint32_t)bulk * 4u);
        tma_bulk_g2s(land, src0, (uint32_t)bulk * 4u, w.mbar);
    }
    return shift;
}

// One frame, from samples to spectrum, shared by the feature kernel and the chroma kernel.
// On return P[i] / S[i] hold |X|^2 / |X| of bin 16*lane + i (.x, the lane's low run) and of bin
// 1024 - 16*lane - i (.y, its mirrored high run); p512 / s512 are bin 512; ss = sum of squares of
// the frame's samples, zc = zero crossings; m0*/m1*/m2* are the lane's magnitude moments (orders
// 0-2) about the centres of its two 16-bin runs.
// All complex arithmetic is packed FP32 (FADD2 / FMUL2 / FFMA2): one float2 = (re, im).
//
// PREF (full-length periodic Hann window only): (1) the window is not read from shared memory but
// synthesised, 0.5*w[2(lane+32j)+c] = 0.25 - 0.25 cos(theta_lane,c + 2 pi j / 32), expanded with the
// angle-addition formula into two FFMA2 on per-lane (cos, sin) pairs and immediates; the 8 KB this
// frees pay for (2) a landing zone of its own, so the TMA copy of the warp's NEXT frame (nclip, nt) is
// started as soon as the transposes are done and flies while the statistics and the mel projection of
// this frame run - the wait at the top of the next frame is then (nearly) free.
template <bool PREF, bool TM = false>
__device__ __forceinline__ void frame_spectrum(const FrameArgs& a, WarpState& w, const float* clip, int t,
                                               const float* nclip, int nt,
                                               float2 (&P)[16], float2 (&S)[16], float& p512, float& s512,
                                               float& ss, int& zc, float& m0l, float& m1l, float& m2l,
                                               float& m0h, float& m1h, float& m2h) {
    float* const sc = w.sc;
    float2* const sc2 = w.sc2;
    uint64_t* const mbar = w.mbar;
    const float2* const s_win = w.s_win;
    const float2* const s_tw1 = w.s_tw1;
    const float2* const s_tw2 = w.s_tw2;
    const int lane = w.lane, zw_base = w.zw_base, zlo_base = w.zlo_base, zhi_base = w.zhi_base, zhi0 = w.zhi0;
    const float zthr = a.zcr_thr;
    uint32_t& parity = w.parity;
    const int fs = t * a.hop - a.pad;              // first sample of the frame (clip coords)
    float* const land = PREF ? sc + kLandOff : sc; // where the frame's samples are staged
    int zc_edge = -1;
    int off;

    // ---- stage the frame's samples in this warp's buffer
    if (!PREF || !w.pending) off = issue_frame_tma(a, w, land, clip, t);
    else off = w.pend_off;
    w.pending = false;
    if (off >= 0) {
        mbar_wait(mbar, parity);
        parity ^= 1u;
    } else {
        // edge frame (or odd alignment): build the padded frame by hand; ZCR pads with "edge"
        int zc = 0;
        unsigned prev_last = 0u;
        for (int c0 = 0; c0 < kFastNfft / 32; c0 += 8) {
            float ve[8], vp[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) ve[u] = sample_edge(clip, a.n, fs + 32 * (c0 + u) + lane);   // 8 loads in flight
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int s = fs + 32 * (c0 + u) + lane;
                const bool inside = (s >= 0) && (s < a.n);
                vp[u] = inside ? ve[u] : ((a.pad_mode == 0) ? 0.0f : (a.pad_mode == 1) ? __ldg(clip + reflect_index(s, a.n)) : ve[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                land[32 * (c0 + u) + lane] = vp[u];
                const unsigned msk = __ballot_sync(FULL, ve[u] < -zthr);
                zc += __popc((msk ^ (msk >> 1)) & 0x7fffffffu);
                if (c0 + u > 0) zc += ((msk & 1u) != prev_last) ? 1 : 0;
                prev_last = msk >> 31;
            }
        }
        zc_edge = zc;
        off = 0;
        __syncwarp();
    }

    unsigned za = 0u, zb = 0u;
    float2 v[32];

    // ---- phase 0: frame -> registers; window; RMS and ZCR partials
    if constexpr (TM) {
        // the window comes from TMEM, 8 sample pairs at a time (their LDS are in flight while the tcgen05.ld lands)
        const float2* xp = reinterpret_cast<const float2*>(land + off);
        const float2 zt = make_float2(zthr, zthr);
        float2 ss2 = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int jc = 0; jc < 4; ++jc) {
            uint32_t wr[16];
            tmem_ld16_issue(w.tq + kTmWin + 16 * jc, wr);
            float2 x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = xp[lane + 32 * (8 * jc + u)];
            tmem_wait16(wr);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                ss2 = __ffma2_rn(x[u], x[u], ss2);
                const float2 xt = __fadd2_rn(x[u], zt);
                za = __funnelshift_l(__float_as_uint(xt.x), za, 1);
                zb = __funnelshift_l(__float_as_uint(xt.y), zb, 1);
                v[8 * jc + u] = __fmul2_rn(x[u], f2_of(wr[2 * u], wr[2 * u + 1]));
            }
        }
        ss = ss2.x + ss2.y;
    } else
    {
        const float2* xp = reinterpret_cast<const float2*>(land + off);
        const float2 zt = make_float2(zthr, zthr);
        float2 ss2 = make_float2(0.0f, 0.0f);
        float2 hc = make_float2(0.f, 0.f), hs = hc;
        if (PREF) { const float4 cs = w.s_hcs[lane]; hc = make_float2(cs.x, cs.y); hs = make_float2(cs.z, cs.w); }
        const float2 quarter = make_float2(0.25f, 0.25f);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float2 x = xp[lane + 32 * j];
            float2 wn;
            if (PREF) {
                // 0.25 - 0.25 cos(theta) cos(phi_j) + 0.25 sin(theta) sin(phi_j), phi_j = 2 pi j / 32
                const float cj = -0.25f * float(fftreg::cos2pi(j, 32)), sj = 0.25f * float(fftreg::sin2pi(j, 32));
                if (j == 0 || j == 16) wn = __ffma2_rn(hc, make_float2(cj, cj), quarter);
                else if (j == 8 || j == 24) wn = __ffma2_rn(hs, make_float2(sj, sj), quarter);
                else wn = __ffma2_rn(hc, make_float2(cj, cj), __ffma2_rn(hs, make_float2(sj, sj), quarter));
            } else {
                wn = s_win[lane + 32 * j];
            }
            ss2 = __ffma2_rn(x, x, ss2);
            // sign bit of (x + thr) <=> x < -thr; sample j ends up at bit 31 - j
            const float2 xt = __fadd2_rn(x, zt);
            za = __funnelshift_l(__float_as_uint(xt.x), za, 1);
            zb = __funnelshift_l(__float_as_uint(xt.y), zb, 1);
            v[j] = __fmul2_rn(x, wn);
        }
        ss = ss2.x + ss2.y;
    }
    {
        // pairs (2m, 2m+1) sit in one lane; pairs (2m+1, 2m+2) straddle to the next lane
        unsigned zn = __shfl_sync(FULL, za, (lane + 1) & 31);
        unsigned msk = FULL;
        if (lane == 31) { zn <<= 1; msk = 0xfffffffeu; }
        zc = __popc(za ^ zb) + __popc((zb ^ zn) & msk);
        zc = warp_sum_i(zc);
        if (zc_edge >= 0) zc = zc_edge;
    }
    ss = warp_sum(ss);
    __syncwarp();                       // every lane has its samples: the buffer becomes scratch

    // ---- phase 1: 32-point FFT over n1 (this lane holds z[lane + 32*n1])
    fftreg2::fft_dif<32>(v);

    // ---- phase 2: inter-pass twiddle W_1024^(lane*k1); the table holds (cos, -sin)
    if constexpr (TM) {
        // two tcgen05.ld in flight per wait (the wait drains every outstanding load)
#pragma unroll
        for (int kc = 0; kc < 4; kc += 2) {
            uint32_t tr[16], tr2[16];
            tmem_ld16_issue(w.tq + kTmTw1 + 16 * kc, tr);
            tmem_ld16_issue(w.tq + kTmTw1 + 16 * kc + 16, tr2);
            tmem_wait16(tr);
            tmem_wait16(tr2);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int k1 = 8 * kc + u + 1;
                if (k1 < 32) {
                    const int p = pos32(k1);
                    const uint32_t c = (u < 8) ? tr[2 * (u & 7)] : tr2[2 * (u & 7)];
                    const uint32_t sn = (u < 8) ? tr[2 * (u & 7) + 1] : tr2[2 * (u & 7) + 1];
                    v[p] = fftreg2::cmul(v[p], __uint_as_float(c), __uint_as_float(sn));
                }
            }
        }
    } else {
#pragma unroll
        for (int k1 = 1; k1 < 32; ++k1) {
            const float2 tw = s_tw1[(k1 - 1) * 32 + lane];
            const int p = pos32(k1);
            v[p] = fftreg2::cmul(v[p], tw.x, tw.y);
        }
    }

    // ---- phase 3: 32x32 complex transpose through shared memory
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) sc2[lane * 33 + k1] = v[pos32(k1)];
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[n2] = sc2[n2 * 33 + lane];
    __syncwarp();

    // ---- phase 4: 32-point FFT over n2; lane = k1, bin k = k1 + 32*k2 at pos32(k2)
    fftreg2::fft_dif<32>(v);

    // ---- phase 5: regroup so each lane owns bins [16*lane, 16*lane+16) (v[0..15]) and their
    //      mirrors 1024-k (v[16..31])
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) sc2[zw_base + 34 * k2] = v[pos32(k2)];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = sc2[zlo_base + i];
    v[16] = sc2[zhi0];
#pragma unroll
    for (int i = 1; i < 16; ++i) v[16 + i] = sc2[zhi_base - i];
    const float2 e512 = sc2[544];
    __syncwarp();

    // ---- the landing zone is free again: start the copy of this warp's next frame
    if (PREF && nclip != nullptr) {
        const int sh = issue_frame_tma(a, w, land, nclip, nt);
        w.pending = sh >= 0;
        w.pend_off = sh;
    }

    // ---- phase 6: real-FFT split, |X|^2 and |X|, local moments of |X|
    float2 M0 = make_float2(0.f, 0.f), M1 = M0, M2 = M0;
    float2 tw_base = make_float2(0.f, 0.f);        // -i * W_2048^(16*lane); bin 16*lane + i needs it times W_2048^i
    uint32_t t2a[16], t2b[16];
    if constexpr (TM) {
        tmem_ld16_issue(w.tq + kTmTw2, t2a);
        tmem_ld16_issue(w.tq + kTmTw2 + 16, t2b);
        tmem_wait16(t2a);
        tmem_wait16(t2b);
    } else {
        tw_base = s_tw2[lane];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float2 tw = TM ? (i < 8 ? f2_of(t2a[2 * i], t2a[2 * i + 1]) : f2_of(t2b[2 * (i - 8)], t2b[2 * (i - 8) + 1]))
                        : (i == 0) ? tw_base
                                   : fftreg2::cmul_conj(tw_base, float(fftreg::cos2pi(i, 2048)), float(fftreg::sin2pi(i, 2048)));
        const float2 za_ = v[i], zb_ = v[16 + i];
        const float2 e = __fadd2_rn(za_, make_float2(zb_.x, -zb_.y));      // Z[k] + conj Z[1024-k]
        const float2 d = __fadd2_rn(za_, make_float2(-zb_.x, zb_.y));      // Z[k] - conj Z[1024-k]
        const float2 tt = fftreg2::cmul(d, tw.x, tw.y);
        // X[k] = e + t, conj X[1024-k] = e - t (both carry the 1/2 folded into the window)
        const float2 xa = __fadd2_rn(e, tt), xb = __fadd2_rn(e, make_float2(-tt.x, -tt.y));
        const float2 pw = make_float2(fmaf(xa.x, xa.x, xa.y * xa.y), fmaf(xb.x, xb.x, xb.y * xb.y));
        const float2 sq = make_float2(fast_sqrt(pw.x), fast_sqrt(pw.y));
        P[i] = pw;
        S[i] = sq;
        const float dd = float(i) - 7.5f;
        M0 = __fadd2_rn(M0, sq);
        M1 = __ffma2_rn(make_float2(sq.x, -sq.y), make_float2(dd, dd), M1);
        M2 = __ffma2_rn(sq, make_float2(dd * dd, dd * dd), M2);
    }
    m0l = M0.x; m0h = M0.y; m1l = M1.x; m1h = M1.y; m2l = M2.x; m2h = M2.y;
    // bin 512 pairs with itself: X[512] = 2*conj(Zhalf[512])
    p512 = 4.0f * fmaf(e512.x, e512.x, e512.y * e512.y);
    s512 = fast_sqrt(p512);
}

template <int NW, bool PIP, bool PREF, bool TM = false, int UNR = 0>
__global__ void __launch_bounds__(NW * 32, 1)
frames_fast_2048(const FrameArgs a, const float* __restrict__ g_tables, const FastTables ft) {
    static_assert(!TM || PREF, "TM kernels use the layout with a landing zone of its own");
    static_assert(UNR == 0 || TM, "compile-time mel step counts: TM kernels only");
    extern __shared__ __align__(16) float smem[];
    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const FastSmemLayout L = fast_layout(ft, NW, PREF, TM);

    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem) + warp;
    float* tab = smem + L.tables;
    const float2* s_win = reinterpret_cast<const float2*>(tab + ft.win);   // (not staged by PREF kernels)
    const float2* s_tw1 = reinterpret_cast<const float2*>(tab + ft.tw1);
    const float2* s_tw2 = reinterpret_cast<const float2*>(tab + ft.tw2);
    const int* s_meta = reinterpret_cast<const int*>(tab + ft.mel_meta);
    const float* s_melw = tab + ft.mel_w;
    float* sc = smem + L.wbuf + warp * L.wb;          // this warp's buffer
    float2* sc2 = reinterpret_cast<float2*>(sc);

    // ---- one-time CTA setup: tables -> smem, zero the warp buffers, one mbarrier per warp
    for (int i = tid; i < L.ntab / 4; i += NT)
        reinterpret_cast<float4*>(tab)[i] = __ldg(reinterpret_cast<const float4*>(g_tables) + i);
    for (int i = L.wbuf + tid; i < L.total; i += NT) smem[i] = 0.0f;
    if (lane == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    // ---- frame assignment: the CTA owns a contiguous run of (clip, frame) pairs and its NW warps
    //      take them round-robin, so at any moment one SM works on NW neighbouring frames of one
    //      clip (their 75 %-overlapping samples are hot in L2) and the whole GPU on ~148 clips
    const long long total = (long long)a.B * a.T;
    const long long per_cta = (total + gridDim.x - 1) / gridDim.x;
    const long long c0 = (long long)blockIdx.x * per_cta;
    const long long g1 = (c0 + per_cta < total) ? c0 + per_cta : total;
    const long long g0 = c0 + warp;
    int b = (int)(g0 / a.T);
    int t = (int)(g0 - (long long)b * a.T);
    WarpState w;
    w.sc = sc; w.sc2 = sc2; w.mbar = mbar; w.parity = 0;
    w.s_win = s_win; w.s_tw1 = s_tw1; w.s_tw2 = s_tw2; w.lane = lane;
    w.s_hcs = reinterpret_cast<const float4*>(tab + ft.hann_cs);
    w.pending = false; w.pend_off = 0;
    uint32_t tbase = 0;
    if constexpr (TM) {
        // the first frame's copy flies while the Tensor Memory tables are filled (what a one-clip call waits for)
        __syncthreads();                         // the landing zones are zeroed, the mbarriers initialised
        if (g0 < g1) {
            const int sh = issue_frame_tma(a, w, sc + kLandOff, a.wave + (long long)b * a.pitch, t);
            w.pending = sh >= 0;
            w.pend_off = sh;
        }
        tbase = tmem_tables_setup<NW>(ft.tmem_tab, ft.tmem_cols, reinterpret_cast<uint32_t*>(smem + 2 * NW), warp, lane);
    }
    __syncthreads();
    if constexpr (TM) asm volatile("tcgen05.fence::after_thread_sync;");
    if (!TM && g0 >= g1) return;                 // (TM kernels: every warp stays for the TMEM deallocation)
    float clip_max = 0.0f;
    // (REDUX leaves its result in a uniform register: the tcgen05.ld addresses derived from it need no R2UR)
    w.tq = __reduce_or_sync(FULL, tbase + ((uint32_t)(32 * (warp & 3)) << 16));
    // per-lane bases of the regroup buffer, layout p(k) = k + k/16 in float2 units: every
    // access is base + immediate and conflict-free (17 is odd)
    w.zw_base = lane + (lane >> 4);              // write: k = lane + 32*k2 -> zw_base + 34*k2
    w.zlo_base = 17 * lane;                      // read:  k = 16*lane + i  -> zlo_base + i
    w.zhi_base = 17 * (63 - lane) + 16;          // read:  k = 1024-16*lane-i (i>=1) -> zhi_base - i
    w.zhi0 = (lane == 0) ? 0 : 17 * (64 - lane); // read:  k = (1024-16*lane) mod 1024

    for (long long g = g0; g < g1; g += NW) {
        const float* clip = a.wave + (long long)b * a.pitch;
        // this warp's next frame (PREF kernels start its copy half-way through this one)
        int nb = b, nt = t + NW;
        while (nt >= a.T) { nt -= a.T; ++nb; }
        const float* nclip = (PREF && g + NW < g1) ? a.wave + (long long)nb * a.pitch : nullptr;
        float2 P[16], S[16];
        float p512, s512, ss, m0l, m1l, m2l, m0h, m1h, m2h;
        int zc;
        frame_spectrum<PREF, TM>(a, w, clip, t, nclip, nt, P, S, p512, s512, ss, zc, m0l, m1l, m2l, m0h, m1h, m2h);

        // ---- chroma_stft: keep the power spectrum for the projection that follows the tuning estimate
        //      (8 coalesced 16-byte stores per lane, in the lane's register order: no second STFT pass)
        if (PIP && a.pstash != nullptr) {
            float* fr = a.pstash + ((size_t)b * a.T + t) * kStashFloats;
            float4* dst = reinterpret_cast<float4*>(fr) + lane * 8;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                dst[q4] = make_float4(P[4 * q4].x, P[4 * q4 + 1].x, P[4 * q4 + 2].x, P[4 * q4 + 3].x);
                dst[4 + q4] = make_float4(P[4 * q4].y, P[4 * q4 + 1].y, P[4 * q4 + 2].y, P[4 * q4 + 3].y);
            }
            if (lane == 31) fr[32 * 32] = p512;
        }

        // ---- centroid / bandwidth (librosa.feature.spectral_centroid / _bandwidth)
        const float kcl = 16.0f * lane + 7.5f;
        const float kch = 1024.0f - 16.0f * lane - 7.5f;
        float s0 = m0l + m0h;
        float s1 = fmaf(kcl, m0l, m1l) + fmaf(kch, m0h, m1h);
        if (lane == 31) { s0 += s512; s1 = fmaf(512.0f, s512, s1); }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        const float denom = (s0 < 1.17549435e-38f) ? 1.0f : s0;   // util.normalize tiny guard
        const float cen = s1 / denom;                            // in bins
        // sum |X| (k - centroid)^2 from the moments about the run centres: with d = k - kc,
        // sum s (d + (kc - c))^2 = m2 + 2 (kc - c) m1 + (kc - c)^2 m0.  |d| <= 7.5, so the cancellation
        // that rules out one-pass moments about the origin (k up to 1024) is bounded by a few tens of
        // ulps of the lane's own energy; librosa's two-pass form differs by < 1e-5 bins^2.
        float q;
        {
            const float dl = kcl - cen, dh = kch - cen;
            q = fmaf(dl, fmaf(dl, m0l, 2.0f * m1l), m2l) + fmaf(dh, fmaf(dh, m0h, 2.0f * m1h), m2h);
            if (lane == 31) { const float dm = 512.0f - cen; q = fmaf(dm * dm, s512, q); }
            q = warp_sum(q);
        }
        const float bw = sqrtf(fmaxf(q, 0.0f) / denom);

        // ---- rolloff (librosa.feature.spectral_rolloff): first bin whose cumulative
        //      magnitude reaches roll_percent * total
        int rbin;
        {
            float pl = m0l;                       // inclusive scan, ascending lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float tv = __shfl_up_sync(FULL, pl, d);
                if (lane >= d) pl += tv;
            }
            float ph = m0h;                       // inclusive scan, descending lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float tv = __shfl_down_sync(FULL, ph, d);
                if (lane + d < 32) ph += tv;
            }
            const float tot_lo = __shfl_sync(FULL, pl, 31);
            const float tot_hi = __shfl_sync(FULL, ph, 0);
            const float mid = tot_lo + s512;      // s512 is warp-uniform (same smem word)
            const float thr = a.roll_percent * (mid + tot_hi);
            const unsigned lo_mask = __ballot_sync(FULL, pl >= thr);
            if (lo_mask) {
                const int tl = __ffs(lo_mask) - 1;
                float cum = pl - m0l;
                int cnt = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) { cum += S[i].x; cnt += (cum < thr) ? 1 : 0; }
                rbin = __shfl_sync(FULL, 16 * lane + min(cnt, 15), tl);
            } else if (mid >= thr) {
                rbin = 512;
            } else {
                const unsigned hi_mask = __ballot_sync(FULL, mid + ph >= thr);
                if (hi_mask) {
                    const int tl = 31 - __clz(hi_mask);
                    float cum = mid + (ph - m0h);
                    int cnt = 0;
#pragma unroll
                    for (int i = 15; i >= 0; --i) { cum += S[i].y; cnt += (cum < thr) ? 1 : 0; }
                    // ascending bins are i = 15..0: the cnt-th of them is i = 15 - cnt
                    rbin = __shfl_sync(FULL, 1024 - 16 * lane - (15 - min(cnt, 15)), tl);
                } else {
                    rbin = 1024;
                }
            }
        }

        // ---- phase 7: power (or magnitude) spectrum -> padded scratch, q(k) = k + k/16
        if (a.use_mag) {
#pragma unroll
            for (int i = 0; i < 16; ++i) sc[17 * lane + i] = S[i].x;
            sc[17 * (64 - lane)] = S[0].y;
#pragma unroll
            for (int i = 1; i < 16; ++i) sc[17 * (63 - lane) + 16 - i] = S[i].y;
            if (lane == 31) sc[544] = s512;
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) sc[17 * lane + i] = P[i].x;
            sc[17 * (64 - lane)] = P[0].y;
#pragma unroll
            for (int i = 1; i < 16; ++i) sc[17 * (63 - lane) + 16 - i] = P[i].y;
            if (lane == 31) sc[544] = p512;
        }
        sc[17 * lane + 16] = 0.0f;
        sc[17 * (63 - lane) + 16] = 0.0f;
        if (lane < 16) sc[1089 + lane] = 0.0f;    // zero-weight taps past bin 1024 must read finite values
        __syncwarp();

        // ---- optional: librosa.piptrack candidates for chroma_stft's tuning estimate
        //      (parabolic interpolation at thresholded local maxima of the power spectrum)
        if (PIP) {
            float pmax = p512;
#pragma unroll
            for (int i = 0; i < 16; ++i) pmax = fmaxf(pmax, fmaxf(P[i].x, P[i].y));
            pmax = warp_max(pmax);
            const float ref = a.pip_threshold * pmax;
            // every frame owns a fixed block of cand_cap slots (at most every other bin of the band
            // can be a peak), so no atomics are needed; the count goes to cand_count[b*T + t]
            float2* slots = a.cand + ((size_t)b * a.T + t) * a.cand_cap;
            int base = 0;
            for (int k0 = a.pip_klo; k0 < a.pip_khi; k0 += 32) {
                const int k = k0 + lane;
                const bool in = k < a.pip_khi;
                const int kk = in ? k : a.pip_klo;
                const float pm1 = sc[(kk - 1) + ((kk - 1) >> 4)];
                const float p0 = sc[kk + (kk >> 4)];
                const float pp1 = sc[(kk + 1) + ((kk + 1) >> 4)];
                const float xm = (pm1 > ref) ? pm1 : 0.0f, x0 = (p0 > ref) ? p0 : 0.0f, xp = (pp1 > ref) ? pp1 : 0.0f;
                const bool peak = in && (x0 > xm) && (x0 >= xp);
                const unsigned bal = __ballot_sync(FULL, peak);
                if (peak) {
                    const float avg = (pp1 - pm1) * 0.5f;
                    const float aa = (pp1 + pm1) - 2.0f * p0;
                    const float shift = (fabsf(avg) >= fabsf(aa)) ? 0.0f : -avg / aa;
                    const float pitch = (float(k) + shift) * a.binhz;
                    const float mag = p0 + (0.5f * avg) * shift;
                    const int slot = base + __popc(bal & ((1u << lane) - 1u));
                    if (slot < a.cand_cap) slots[slot] = make_float2(pitch, mag);
                }
                base += __popc(bal);
            }
            if (lane == 0) a.cand_count[(size_t)b * a.T + t] = min(base, a.cand_cap);
        }

        // ---- phase 8: banded mel projection (librosa.feature.melspectrogram's einsum);
        //      start offsets were shifted on the host so that the 32 lanes hit 32 banks
        if (a.mel_out != nullptr) {
            float wmax = 0.0f;
            const size_t mstride = a.mel_frame_major ? 1 : (size_t)a.T;
            float* outb = a.mel_frame_major ? a.mel_out + ((size_t)b * a.T + t) * a.n_mels
                                            : a.mel_out + ((size_t)b * a.n_mels) * a.T + t;
            if constexpr (TM && UNR != 0) {
                uint32_t st4[4];
                tmem_ld4(w.tq + kTmMeta, st4);       // the lane's first tap in each group
                const float acc0 = mel_group_tm<UNR, 0>(w.tq, sc + st4[0]);
                const float acc1 = mel_group_tm<UNR, 1>(w.tq, sc + st4[1]);
                const float acc2 = mel_group_tm<UNR, 2>(w.tq, sc + st4[2]);
                const float acc3 = mel_group_tm<UNR, 3>(w.tq, sc + st4[3]);
                outb[(size_t)lane * mstride] = acc0;
                outb[(size_t)(32 + lane) * mstride] = acc1;
                outb[(size_t)(64 + lane) * mstride] = acc2;
                outb[(size_t)(96 + lane) * mstride] = acc3;
                wmax = fmaxf(fmaxf(acc0, acc1), fmaxf(acc2, acc3));
            } else if constexpr (TM) {
                uint32_t st4[4], st8[4];
                tmem_ld4(w.tq + kTmMeta, st4);       // the lane's first tap in each group
                if (ft.n_groups > 4) tmem_ld4(w.tq + kTmMeta + 4, st8);
                uint32_t col = w.tq + kTmMel;
#pragma unroll
                for (int g = 0; g < kMaxMelGroups; ++g) {
                    if (g < ft.n_groups) {
                        const int n4 = ft.mel_steps[g];
                        const float* pp = sc + (g < 4 ? st4[g & 3] : st8[g & 3]);
                        float2 a01 = make_float2(0.0f, 0.0f), a23 = a01;
                        mel_steps_tm(n4, col, pp, a01, a23);
                        col += 4 * n4;
                        a01 = __fadd2_rn(a01, a23);
                        const float acc = a01.x + a01.y;
                        const int m = 32 * g + lane;
                        if (m < a.n_mels) outb[(size_t)m * mstride] = acc;
                        wmax = fmaxf(wmax, acc);
                    }
                }
            } else
            for (int g = 0; g < ft.n_groups; ++g) {
                const int n4 = s_meta[g];
                const float4* wp = reinterpret_cast<const float4*>(s_melw + s_meta[kMaxMelGroups + g]) + lane;
                const float* pp = sc + s_meta[2 * kMaxMelGroups + 32 * g + lane];
                float2 a01 = make_float2(0.0f, 0.0f), a23 = a01;
                mel_steps(n4, wp, pp, a01, a23);
                a01 = __fadd2_rn(a01, a23);
                const float acc = a01.x + a01.y;
                const int m = 32 * g + lane;
                if (m < a.n_mels) outb[(size_t)m * mstride] = acc;
                wmax = fmaxf(wmax, acc);
            }
            clip_max = fmaxf(clip_max, wmax);
        }
        if (lane == 0) {
            if (a.stats != nullptr) {
                float* st = a.stats + (size_t)b * 5 * a.T + t;
                st[0] = cen * a.binhz;
                st[(size_t)a.T] = bw * a.binhz;
                st[(size_t)2 * a.T] = float(rbin) * a.binhz;
                st[(size_t)3 * a.T] = float(zc) * (1.0f / float(kFastNfft));
                st[(size_t)4 * a.T] = sqrtf(ss * (1.0f / float(kFastNfft)));
            }
            // librosa.util.valid_audio: a non-finite sample poisons the sum of squares
            if (a.status != nullptr && !(fabsf(ss) <= 3.0e38f)) atomicOr(a.status + b, 1);
        }
        // ---- next frame; flush the running mel maximum when the clip (or this warp's run) ends
        const bool last = (t + NW >= a.T) || (g + NW >= g1);
        if (last && a.clipmax != nullptr && a.mel_out != nullptr) {
            const float m = warp_max(clip_max);
            if (lane == 0) atomicMax(reinterpret_cast<int*>(a.clipmax) + b, __float_as_int(m));
            clip_max = 0.0f;
        }
        t = nt; b = nb;
        __syncwarp();                       // scratch reads are done before the next frame lands
    }
    if constexpr (TM) tmem_tables_release(tbase, warp);
}

template <int NW, bool PIP, bool PREF, bool TM = false, int UNR = 0>
static cudaError_t launch_fast_nw(const FrameArgs& a, const float* d_tables, const FastTables& ft,
                                  int num_sms, cudaStream_t stream) {
    const int smem = fast_layout(ft, NW, PREF, TM).total * 4;
    cudaError_t e = cudaFuncSetAttribute(frames_fast_2048<NW, PIP, PREF, TM, UNR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const long long frames = (long long)a.B * a.T;
    if (frames <= 0) return cudaSuccess;
    long long grid = (frames + NW - 1) / NW;
    if (grid > num_sms) grid = num_sms;
    frames_fast_2048<NW, PIP, PREF, TM, UNR><<<(unsigned)grid, NW * 32, smem, stream>>>(a, d_tables, ft);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_frames_fast(const FrameArgs& a, const float* d_tables, const FastTables& ft,
                               int num_sms, cudaStream_t stream) {
    const bool pip = (a.cand != nullptr);
    constexpr int kMaxSmem = 227 * 1024;
    // default window: 16 warps with the next frame's copy in flight (Hann synthesised in registers)
    static const bool no_pref = [] { const char* e = getenv("HLMC_NO_PREF"); return e && e[0] == '1'; }();
    // per-lane tables in Tensor Memory (any window: it is one of the tables), next frame's copy in flight
    static const bool no_tm = [] { const char* e = getenv("HLMC_NO_TMEM"); return e && e[0] == '1'; }();
    const bool use_tm = !no_pref && !no_tm && !a.no_tmem && ft.tmem_tab != nullptr && fast_layout(ft, 16, true, true).total * 4 <= kMaxSmem;
    if (use_tm) {
        const bool dflt = ft.n_groups == 4 && a.n_mels == 128 &&
                          (ft.mel_steps[0] | (ft.mel_steps[1] << 8) | (ft.mel_steps[2] << 16) | (ft.mel_steps[3] << 24)) == kMelUnrDefault;
        if (dflt)
            return pip ? launch_fast_nw<16, true, true, true, kMelUnrDefault>(a, d_tables, ft, num_sms, stream)
                       : launch_fast_nw<16, false, true, true, kMelUnrDefault>(a, d_tables, ft, num_sms, stream);
    }
    if (use_tm)
        return pip ? launch_fast_nw<16, true, true, true>(a, d_tables, ft, num_sms, stream)
                   : launch_fast_nw<16, false, true, true>(a, d_tables, ft, num_sms, stream);
    if (!no_pref && ft.hann && fast_layout(ft, 16, true).total * 4 <= kMaxSmem)
        return pip ? launch_fast_nw<16, true, true>(a, d_tables, ft, num_sms, stream)
                   : launch_fast_nw<16, false, true>(a, d_tables, ft, num_sms, stream);
    if (fast_layout(ft, 16, false).total * 4 <= kMaxSmem)
        return pip ? launch_fast_nw<16, true, false>(a, d_tables, ft, num_sms, stream)
                   : launch_fast_nw<16, false, false>(a, d_tables, ft, num_sms, stream);
    if (fast_layout(ft, 8, false).total * 4 <= kMaxSmem)
        return pip ? launch_fast_nw<8, true, false>(a, d_tables, ft, num_sms, stream)
                   : launch_fast_nw<8, false, false>(a, d_tables, ft, num_sms, stream);
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------
// Register-FFT path for n_fft = 4096 (full-length periodic Hann window).  One warp per frame as in
// frames_fast_2048, but the 2048-point complex FFT of z[m] = x[2m] + i x[2m+1] does not fit the
// registers of one warp, so it is split by one radix-2 decimation-in-frequency step into two
// 1024-point transforms that run one after the other through the same 32 x 32 machinery:
//     pass 0:  a[m] = z[m] + z[m+1024]                      -> Z[2k']   (even bins of the real FFT)
//     pass 1:  b[m] = (z[m] - z[m+1024]) * W_2048^m         -> Z[2k'+1] (odd bins)
// b is parked in the slot of z[m] (only this lane ever reads it), the Hann window of the second half
// is 0.5 - w of the first (no second synthesis), each pass does its own real-FFT split (the partner of
// bin k is 2048 - k, which has the same parity), its own share of the moments and its own half of the
// mel gather (filterbank columns de-interleaved on the host), so only the 32 magnitudes per lane that
// rolloff needs in bin order are carried from pass 0 to pass 1 (through shared memory).
// Shared memory per warp: 4100 floats landing zone (first half becomes b), 2180 floats transpose /
// power-spectrum scratch (overlapping the second half of the landing zone), 1056 floats magnitudes.
// ---------------------------------------------------------------------------
constexpr int k4Nfft = 4096;
constexpr int k4T = 2052;                        // transposes / power scratch start (past b, 16-byte aligned)
constexpr int k4Sev = k4T + 2180;                // pass-0 magnitudes, [lane][33]
constexpr int k4WarpFloats = k4Sev + 32 * 33;    // 5288
constexpr int k4Warps = 8;

int fast4_smem_bytes(const Fast4Tables& ft, bool tm) {
    return (((2 * k4Warps + 3 + (tm ? 4 : 0)) & ~3) + (tm ? 0 : ft.total) + k4Warps * k4WarpFloats) * 4;
}

// phases 1-5 of one 1024-point pass: v[j] = element lane + 32 j on entry; on exit v[i] / v[16+i] hold the
// lane's low run k' = 16 lane + i and its partners (pass 0: 1024 - k', pass 1: 1023 - k').  `pass` is a
// run-time value on purpose: both passes execute the SAME instructions (the kernel's code would otherwise
// not fit the instruction cache; ncu showed 0.9 no-instruction stall cycles per issue with two copies).
template <bool TM>
__device__ __forceinline__ void fft1024_regroup(float2 (&v)[32], float2* __restrict__ sc2,
                                                const float2* __restrict__ s_tw1, uint32_t tq, int lane, int pass, float2& e512) {
    fftreg2::fft_dif<32>(v);
    if constexpr (TM) {
#pragma unroll
        for (int kc = 0; kc < 4; kc += 2) {
            uint32_t tr[16], tr2[16];
            tmem_ld16_issue(tq + k4TmTw1 + 16 * kc, tr);
            tmem_ld16_issue(tq + k4TmTw1 + 16 * kc + 16, tr2);
            tmem_wait16(tr);
            tmem_wait16(tr2);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int k1 = 8 * kc + u + 1;
                if (k1 < 32) {
                    const int p = pos32(k1);
                    const uint32_t c = (u < 8) ? tr[2 * (u & 7)] : tr2[2 * (u & 7)];
                    const uint32_t sn = (u < 8) ? tr[2 * (u & 7) + 1] : tr2[2 * (u & 7) + 1];
                    v[p] = fftreg2::cmul(v[p], __uint_as_float(c), __uint_as_float(sn));
                }
            }
        }
    } else {
#pragma unroll
        for (int k1 = 1; k1 < 32; ++k1) {
            const float2 tw = s_tw1[(k1 - 1) * 32 + lane];
            const int p = pos32(k1);
            v[p] = fftreg2::cmul(v[p], tw.x, tw.y);
        }
    }
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) sc2[lane * 33 + k1] = v[pos32(k1)];
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[n2] = sc2[n2 * 33 + lane];
    __syncwarp();
    fftreg2::fft_dif<32>(v);
    const int zw_base = lane + (lane >> 4);
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) sc2[zw_base + 34 * k2] = v[pos32(k2)];
    __syncwarp();
    const int zlo_base = 17 * lane;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = sc2[zlo_base + i];
    // partners: pass 0: k' = 1024 - 16 lane - i (i = 0 wraps to k' = 0 for lane 0); pass 1: k' = 1023 - 16 lane - i
    const int zhi_base = 17 * (63 - lane) + 16 - pass;
    const int zhi0 = (pass == 0) ? ((lane == 0) ? 0 : 17 * (64 - lane)) : zhi_base;
    v[16] = sc2[zhi0];
#pragma unroll
    for (int i = 1; i < 16; ++i) v[16 + i] = sc2[zhi_base - i];
    e512 = sc2[544];
    __syncwarp();
}

// real-FFT split of one pass: P[i] = (|X[k]|^2, |X[partner]|^2), S[i] the magnitudes; moments of the
// magnitudes about the run centres in units of k' (the caller rescales to bins).
__device__ __forceinline__ void split_pass(const float2 (&v)[32], float2 tw_base, float2 (&P)[16], float2 (&S)[16],
                                           float2& M0, float2& M1, float2& M2) {
    M0 = make_float2(0.f, 0.f); M1 = M0; M2 = M0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float2 tw = (i == 0) ? tw_base
                                   : fftreg2::cmul_conj(tw_base, float(fftreg::cos2pi(i, 2048)), float(fftreg::sin2pi(i, 2048)));
        const float2 za_ = v[i], zb_ = v[16 + i];
        const float2 e = __fadd2_rn(za_, make_float2(zb_.x, -zb_.y));
        const float2 d = __fadd2_rn(za_, make_float2(-zb_.x, zb_.y));
        const float2 tt = fftreg2::cmul(d, tw.x, tw.y);
        const float2 xa = __fadd2_rn(e, tt), xb = __fadd2_rn(e, make_float2(-tt.x, -tt.y));
        const float2 pw = make_float2(fmaf(xa.x, xa.x, xa.y * xa.y), fmaf(xb.x, xb.x, xb.y * xb.y));
        const float2 sq = make_float2(fast_sqrt(pw.x), fast_sqrt(pw.y));
        P[i] = pw;
        S[i] = sq;
        const float dd = float(i) - 7.5f;
        M0 = __fadd2_rn(M0, sq);
        M1 = __ffma2_rn(make_float2(sq.x, -sq.y), make_float2(dd, dd), M1);
        M2 = __ffma2_rn(sq, make_float2(dd * dd, dd * dd), M2);
    }
}

// TM: every per-lane table (both twiddle sets, split bases, Hann phase, banded mel weights and first taps of both
// passes) is read from Tensor Memory with tcgen05.ld instead of shared memory, as in frames_fast_2048.
template <bool TM>
__global__ void __launch_bounds__(k4Warps * 32, 1)
frames_fast_4096(const FrameArgs a, const float* __restrict__ g_tables, const Fast4Tables ft) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NW = k4Warps, NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem) + warp;
    float* tab = smem + ((2 * NW + 3 + (TM ? 4 : 0)) & ~3);
    const int tab_total = TM ? 0 : ft.total;              // TM: no table is staged in shared memory
    const float2* s_tw1 = reinterpret_cast<const float2*>(tab + ft.tw1);
    const float2* s_tw0 = reinterpret_cast<const float2*>(tab + ft.tw0);
    const float2* s_base = reinterpret_cast<const float2*>(tab + ft.base);
    const float4* s_hcs = reinterpret_cast<const float4*>(tab + ft.hann_cs);
    float* sc = tab + tab_total + warp * k4WarpFloats;       // landing zone; b after pass 0's first phase
    float* scT = sc + k4T;                                  // transposes, then the pass's power spectrum
    float2* scT2 = reinterpret_cast<float2*>(scT);
    float* sev = sc + k4Sev;

    for (int i = tid; i < tab_total / 4; i += NT)
        reinterpret_cast<float4*>(tab)[i] = __ldg(reinterpret_cast<const float4*>(g_tables) + i);
    for (int i = tid; i < NW * k4WarpFloats; i += NT) (tab + tab_total)[i] = 0.0f;
    if (lane == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    uint32_t tbase = 0;
    if constexpr (TM) tbase = tmem_tables_setup<NW>(ft.tmem_tab, ft.tmem_cols, reinterpret_cast<uint32_t*>(smem + 2 * NW), warp, lane);
    __syncthreads();
    if constexpr (TM) asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tq = __reduce_or_sync(FULL, tbase + ((uint32_t)(32 * (warp & 3)) << 16));   // uniform register: no R2UR per tcgen05.ld

    const long long total = (long long)a.B * a.T;
    const long long per_cta = (total + gridDim.x - 1) / gridDim.x;
    const long long c0 = (long long)blockIdx.x * per_cta;
    const long long g1 = (c0 + per_cta < total) ? c0 + per_cta : total;
    const long long g0 = c0 + warp;
    if (!TM && g0 >= g1) return;                      // (TM: every warp stays for the TMEM deallocation)
    int b = (int)(g0 / a.T);
    int t = (int)(g0 - (long long)b * a.T);
    float clip_max = 0.0f;
    WarpState w{};
    w.sc = sc; w.sc2 = reinterpret_cast<float2*>(sc); w.mbar = mbar; w.parity = 0; w.lane = lane;
    const float zthr = a.zcr_thr;
    float4 hcs;
    if constexpr (TM) {
        uint32_t h4[4];
        tmem_ld4(tq + k4TmHcs, h4);
        hcs = make_float4(__uint_as_float(h4[0]), __uint_as_float(h4[1]), __uint_as_float(h4[2]), __uint_as_float(h4[3]));
    } else {
        hcs = s_hcs[lane];
    }
    const float2 hc = make_float2(hcs.x, hcs.y), hs = make_float2(hcs.z, hcs.w);
    // split-twiddle bases of both passes and the first gather taps of the lane's (pass, group) filters
    uint32_t tb4[4] = {0u, 0u, 0u, 0u}, mst[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if constexpr (TM) {
        tmem_ld4(tq + k4TmBase, tb4);
        uint32_t m0[4], m1[4];
        tmem_ld4(tq + k4TmMeta, m0);
        tmem_ld4(tq + k4TmMeta + 4, m1);
#pragma unroll
        for (int i = 0; i < 4; ++i) { mst[i] = m0[i]; mst[4 + i] = m1[i]; }
    }

    for (long long g = g0; g < g1; g += NW) {
        const float* clip = a.wave + (long long)b * a.pitch;
        const int fs = t * a.hop - a.pad;
        // ---- stage the 4096 samples
        int zc_edge = -1;
        int off = issue_frame_tma(a, w, sc, clip, t, k4Nfft);
        if (off >= 0) {
            mbar_wait(mbar, w.parity);
            w.parity ^= 1u;
        } else {
            int zc = 0;
            unsigned prev_last = 0u;
            for (int c0s = 0; c0s < k4Nfft / 32; c0s += 8) {
                float ve[8], vp[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) ve[u] = sample_edge(clip, a.n, fs + 32 * (c0s + u) + lane);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int s = fs + 32 * (c0s + u) + lane;
                    const bool inside = (s >= 0) && (s < a.n);
                    vp[u] = inside ? ve[u] : ((a.pad_mode == 0) ? 0.0f : (a.pad_mode == 1) ? __ldg(clip + reflect_index(s, a.n)) : ve[u]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    sc[32 * (c0s + u) + lane] = vp[u];
                    const unsigned msk = __ballot_sync(FULL, ve[u] < -zthr);
                    zc += __popc((msk ^ (msk >> 1)) & 0x7fffffffu);
                    if (c0s + u > 0) zc += ((msk & 1u) != prev_last) ? 1 : 0;
                    prev_last = msk >> 31;
                }
            }
            zc_edge = zc;
            off = 0;
            __syncwarp();
        }

        float2 v[32];
        float2 P[16], So[16];                   // after the loop So holds the odd-bin magnitudes
        float2 A0, A1, A2, B0, B1, B2;          // moments of the even (A) and odd (B) bins, (low run, high run)
        float ss = 0.0f, p1024 = 0.0f, s1024 = 0.0f;
        int zc = 0;
        // even-bin halves of the lane's (up to) 4 filters.  Indexed by the rolled group loop, i.e. 16 bytes of
        // local memory: unrolling the loop to keep it in registers measured 12 % slower (code size).
        float macc[4];
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) macc[gi] = 0.0f;
        float wmax = 0.0f;

        // The two passes share ONE copy of the transform / split / gather code (rolled loop): with two copies the
        // kernel spent 0.9 cycles per issue waiting for instructions.
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 0) {
                // ---- even bins: a[m] = z[m] + z[m+1024]; b[m] = z[m] - z[m+1024] parked for pass 1
                float2* L2 = reinterpret_cast<float2*>(sc + off);
                const float2 zt = make_float2(zthr, zthr), quarter = make_float2(0.25f, 0.25f), half = make_float2(0.5f, 0.5f);
                float2 ss2 = make_float2(0.f, 0.f);
                unsigned za0 = 0u, zb0 = 0u, za1 = 0u, zb1 = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float2 x0 = L2[lane + 32 * j], x1 = L2[lane + 32 * j + 1024];
                    // 0.5 * hann of samples 2(lane + 32 j) + c:  0.25 - 0.25 cos(theta + 2 pi j / 64); second half = 0.5 - that
                    const float cj = -0.25f * float(fftreg::cos2pi(j, 64)), sj = 0.25f * float(fftreg::sin2pi(j, 64));
                    float2 w0;
                    if (j == 0) w0 = __ffma2_rn(hc, make_float2(cj, cj), quarter);
                    else if (j == 16) w0 = __ffma2_rn(hs, make_float2(sj, sj), quarter);
                    else w0 = __ffma2_rn(hc, make_float2(cj, cj), __ffma2_rn(hs, make_float2(sj, sj), quarter));
                    const float2 w1 = __fadd2_rn(half, make_float2(-w0.x, -w0.y));
                    ss2 = __ffma2_rn(x0, x0, ss2);
                    ss2 = __ffma2_rn(x1, x1, ss2);
                    const float2 t0 = __fadd2_rn(x0, zt), t1 = __fadd2_rn(x1, zt);
                    za0 = __funnelshift_l(__float_as_uint(t0.x), za0, 1);
                    zb0 = __funnelshift_l(__float_as_uint(t0.y), zb0, 1);
                    za1 = __funnelshift_l(__float_as_uint(t1.x), za1, 1);
                    zb1 = __funnelshift_l(__float_as_uint(t1.y), zb1, 1);
                    const float2 z0 = __fmul2_rn(x0, w0), z1 = __fmul2_rn(x1, w1);
                    v[j] = __fadd2_rn(z0, z1);
                    L2[lane + 32 * j] = __fadd2_rn(z0, make_float2(-z1.x, -z1.y));     // b (before its twiddle)
                }
                ss = warp_sum(ss2.x + ss2.y);
                // rows 0..31 (first half) in word 0, rows 32..63 in word 1, row r at bit 31 - (r & 31)
                unsigned zn0 = __shfl_sync(FULL, za0, (lane + 1) & 31);
                unsigned zn1 = __shfl_sync(FULL, za1, (lane + 1) & 31);
                unsigned m1 = FULL;
                if (lane == 31) { zn0 = (zn0 << 1) | (zn1 >> 31); zn1 <<= 1; m1 = 0xfffffffeu; }
                zc = __popc(za0 ^ zb0) + __popc(za1 ^ zb1) + __popc(zb0 ^ zn0) + __popc((zb1 ^ zn1) & m1);
                zc = warp_sum_i(zc);
                if (zc_edge >= 0) zc = zc_edge;
            } else {
                // ---- odd bins: (z[m] - z[m+1024]) * W_2048^m
                const float2* L2 = reinterpret_cast<const float2*>(sc + off);
                if constexpr (TM) {
#pragma unroll
                    for (int jc = 0; jc < 4; ++jc) {
                        uint32_t tr[16];
                        tmem_ld16_issue(tq + k4TmTw0 + 16 * jc, tr);
                        float2 bq[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) bq[u] = L2[lane + 32 * (8 * jc + u)];
                        tmem_wait16(tr);
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            v[8 * jc + u] = fftreg2::cmul(bq[u], __uint_as_float(tr[2 * u]), __uint_as_float(tr[2 * u + 1]));
                    }
                } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float2 bq = L2[lane + 32 * j];
                    const float2 tw = s_tw0[lane + 32 * j];          // W_2048^m = (cos, -sin)
                    v[j] = fftreg2::cmul(bq, tw.x, tw.y);
                }
                }
            }
            __syncwarp();

            float2 e512, S[16], M0, M1, M2;
            fft1024_regroup<TM>(v, scT2, s_tw1, tq, lane, pass, e512);
            split_pass(v, TM ? (pass ? f2_of(tb4[2], tb4[3]) : f2_of(tb4[0], tb4[1])) : s_base[32 * pass + lane], P, S, M0, M1, M2);
            if (pass == 0) {
                A0 = M0; A1 = M1; A2 = M2;
                p1024 = 4.0f * fmaf(e512.x, e512.x, e512.y * e512.y);      // bin 1024 (k' = 512) pairs with itself
                s1024 = fast_sqrt(p1024);
                // magnitudes of the even bins for rolloff's in-order search
#pragma unroll
                for (int i = 0; i < 16; ++i) { sev[lane * 33 + i] = S[i].x; sev[lane * 33 + 16 + i] = S[i].y; }
            } else {
                B0 = M0; B1 = M1; B2 = M2;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) So[i] = S[i];

            // ---- this pass's power (or magnitude) spectrum, index k', layout q(k') = k' + k'/16
            {
                const bool mag = a.use_mag != 0;
                const int hb = 17 * (63 - lane) + 16 - pass;            // partner of run element i >= 1 sits at hb - i
#pragma unroll
                for (int i = 0; i < 16; ++i) scT[17 * lane + i] = mag ? S[i].x : P[i].x;
                scT[hb + 1 - pass] = mag ? S[0].y : P[0].y;             // pass 0: k' = 1024 - 16 lane, pass 1: 1023 - 16 lane
#pragma unroll
                for (int i = 1; i < 16; ++i) scT[hb - i] = mag ? S[i].y : P[i].y;
                if (pass == 0 && lane == 31) scT[544] = mag ? s1024 : p1024;
                scT[17 * lane + 16] = 0.0f;
                scT[17 * (63 - lane) + 16] = 0.0f;
                if (lane < 16) scT[1089 - pass + lane] = 0.0f;          // taps past the last bin read zeros
                __syncwarp();
            }
            // ---- this pass's half of the mel gather
            if (a.mel_out != nullptr) {
                const int* meta = reinterpret_cast<const int*>(tab + ft.mel_meta[pass]);
                const float* melw = tab + ft.mel_w[pass];
                const size_t mstride = a.mel_frame_major ? 1 : (size_t)a.T;
                float* outb = a.mel_frame_major ? a.mel_out + ((size_t)b * a.T + t) * a.n_mels
                                                : a.mel_out + ((size_t)b * a.n_mels) * a.T + t;
                uint32_t col = tq + (uint32_t)(pass ? ft.mel_col[1] : ft.mel_col[0]);
                for (int gi = 0; gi < ft.n_groups; ++gi) {
                    float2 a01 = make_float2(0.f, 0.f), a23 = a01;
                    if constexpr (TM) {
                        // (constant indices only: a run-time index into the kernel parameters costs a local copy of them)
                        const int sel = 4 * pass + gi;
                        uint32_t start = mst[0];
                        int n4 = ft.mel_steps[0];
#pragma unroll
                        for (int i = 1; i < 8; ++i) {
                            start = (sel == i) ? mst[i] : start;
                            n4 = (sel == i) ? ft.mel_steps[i] : n4;
                        }
                        mel_steps_tm(n4, col, scT + start, a01, a23);
                        col += 4 * n4;
                    } else {
                    const int n4 = meta[gi];
                    const float4* wp = reinterpret_cast<const float4*>(melw + meta[kMaxMelGroups + gi]) + lane;
                    const float* pp = scT + meta[2 * kMaxMelGroups + 32 * gi + lane];
                    mel_steps(n4, wp, pp, a01, a23);
                    }
                    a01 = __fadd2_rn(a01, a23);
                    const float part = a01.x + a01.y;
                    if (pass == 0) {
                        macc[gi & 3] = part;
                    } else {
                        const float acc = part + macc[gi & 3];
                        const int m = 32 * gi + lane;
                        if (m < a.n_mels) outb[(size_t)m * mstride] = acc;
                        wmax = fmaxf(wmax, acc);
                    }
                }
            }
            __syncwarp();                       // the power scratch is read: the next transposes may overwrite it
        }
        clip_max = fmaxf(clip_max, wmax);

        // ---- centroid / bandwidth: even bins k = 32 lane + 15 + 2 d (low), 2048 - 32 lane - 15 - 2 d (high);
        //      odd bins k = 32 lane + 16 + 2 d (low), 2047 - 32 lane - 15 - 2 d (high); d = i - 7.5
        const float kel = 32.0f * lane + 15.0f, keh = 2033.0f - 32.0f * lane;
        const float kol = 32.0f * lane + 16.0f, koh = 2032.0f - 32.0f * lane;
        float s0 = (A0.x + A0.y) + (B0.x + B0.y);
        float s1 = fmaf(kel, A0.x, 2.0f * A1.x) + fmaf(keh, A0.y, 2.0f * A1.y) +
                   fmaf(kol, B0.x, 2.0f * B1.x) + fmaf(koh, B0.y, 2.0f * B1.y);
        if (lane == 31) { s0 += s1024; s1 = fmaf(1024.0f, s1024, s1); }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        const float denom = (s0 < 1.17549435e-38f) ? 1.0f : s0;
        const float cen = s1 / denom;
        float q;
        {
            // sum s (2 d + (kc - c))^2 = 4 m2 + 4 (kc - c) m1 + (kc - c)^2 m0
            const float d0 = kel - cen, d1 = keh - cen, d2 = kol - cen, d3 = koh - cen;
            q = fmaf(d0, fmaf(d0, A0.x, 4.0f * A1.x), 4.0f * A2.x) + fmaf(d1, fmaf(d1, A0.y, 4.0f * A1.y), 4.0f * A2.y) +
                fmaf(d2, fmaf(d2, B0.x, 4.0f * B1.x), 4.0f * B2.x) + fmaf(d3, fmaf(d3, B0.y, 4.0f * B1.y), 4.0f * B2.y);
            if (lane == 31) { const float dm = 1024.0f - cen; q = fmaf(dm * dm, s1024, q); }
            q = warp_sum(q);
        }
        const float bw = sqrtf(fmaxf(q, 0.0f) / denom);

        // ---- rolloff: lane totals over [32 lane, 32 lane + 32) and (2016 - 32 lane, 2048 - 32 lane]
        int rbin;
        {
            const float tl = A0.x + B0.x, th = A0.y + B0.y;
            float pl = tl;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float tv = __shfl_up_sync(FULL, pl, d);
                if (lane >= d) pl += tv;
            }
            float ph = th;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float tv = __shfl_down_sync(FULL, ph, d);
                if (lane + d < 32) ph += tv;
            }
            const float tot_lo = __shfl_sync(FULL, pl, 31);
            const float tot_hi = __shfl_sync(FULL, ph, 0);
            const float mid = tot_lo + s1024;
            const float thr = a.roll_percent * (mid + tot_hi);
            const unsigned lo_mask = __ballot_sync(FULL, pl >= thr);
            if (lo_mask) {
                const int tlane = __ffs(lo_mask) - 1;
                float cum = pl - tl;
                int cnt = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {      // ascending: even 32 lane + 2 i, odd 32 lane + 2 i + 1
                    cum += sev[lane * 33 + i]; cnt += (cum < thr) ? 1 : 0;
                    cum += So[i].x; cnt += (cum < thr) ? 1 : 0;
                }
                rbin = __shfl_sync(FULL, 32 * lane + min(cnt, 31), tlane);
            } else if (mid >= thr) {
                rbin = 1024;
            } else {
                const unsigned hi_mask = __ballot_sync(FULL, mid + ph >= thr);
                if (hi_mask) {
                    const int tlane = 31 - __clz(hi_mask);
                    float cum = mid + (ph - th);
                    int cnt = 0;
#pragma unroll
                    for (int i = 15; i >= 0; --i) {  // ascending: odd 2047 - 32 lane - 2 i, even 2048 - 32 lane - 2 i
                        cum += So[i].y; cnt += (cum < thr) ? 1 : 0;
                        cum += sev[lane * 33 + 16 + i]; cnt += (cum < thr) ? 1 : 0;
                    }
                    // the run's ascending bins start at 2017 - 32 lane
                    rbin = __shfl_sync(FULL, 2017 - 32 * lane + min(cnt, 31), tlane);
                } else {
                    rbin = 2048;
                }
            }
        }

        if (lane == 0) {
            if (a.stats != nullptr) {
                float* st = a.stats + (size_t)b * 5 * a.T + t;
                st[0] = cen * a.binhz;
                st[(size_t)a.T] = bw * a.binhz;
                st[(size_t)2 * a.T] = float(rbin) * a.binhz;
                st[(size_t)3 * a.T] = float(zc) * (1.0f / float(k4Nfft));
                st[(size_t)4 * a.T] = sqrtf(ss * (1.0f / float(k4Nfft)));
            }
            if (a.status != nullptr && !(fabsf(ss) <= 3.0e38f)) atomicOr(a.status + b, 1);
        }
        const bool last = (t + NW >= a.T) || (g + NW >= g1);
        if (last && a.clipmax != nullptr && a.mel_out != nullptr) {
            const float m = warp_max(clip_max);
            if (lane == 0) atomicMax(reinterpret_cast<int*>(a.clipmax) + b, __float_as_int(m));
            clip_max = 0.0f;
        }
        t += NW;
        while (t >= a.T) { t -= a.T; ++b; }
        __syncwarp();
    }
    if constexpr (TM) tmem_tables_release(tbase, warp);
}

template <bool TM>
static cudaError_t launch_fast4096_tm(const FrameArgs& a, const float* d_tables, const Fast4Tables& ft, int num_sms,
                                      cudaStream_t stream) {
    const int smem = fast4_smem_bytes(ft, TM);
    cudaError_t e = cudaFuncSetAttribute(frames_fast_4096<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const long long frames = (long long)a.B * a.T;
    if (frames <= 0) return cudaSuccess;
    long long grid = (frames + k4Warps - 1) / k4Warps;
    if (grid > num_sms) grid = num_sms;
    frames_fast_4096<TM><<<(unsigned)grid, k4Warps * 32, smem, stream>>>(a, d_tables, ft);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_frames_fast4096(const FrameArgs& a, const float* d_tables, const Fast4Tables& ft, int num_sms,
                                   cudaStream_t stream) {
    static const bool no_tm = [] { const char* e = getenv("HLMC_NO_TMEM"); return e && e[0] == '1'; }();
    if (!no_tm && !a.no_tmem && ft.tmem_tab != nullptr) return launch_fast4096_tm<true>(a, d_tables, ft, num_sms, stream);
    return launch_fast4096_tm<false>(a, d_tables, ft, num_sms, stream);
}

// ---------------------------------------------------------------------------
// Register-FFT path for n_fft = 1024 and 512.  L = n_fft / 64 lanes (16 or 8) carry one frame,
// so a warp works on FW = 32 / L consecutive frames of one clip at a time, one per lane group.
// Same data flow as frames_fast_2048: TMA-staged samples -> 32-point FFT in registers ->
// transpose -> L-point FFTs -> regroup -> real split -> statistics -> banded mel gather.
// Every cross-lane step is confined to the group (width-L shuffles, group bits of ballots).
// ---------------------------------------------------------------------------
template <int L> struct SubGeom {
    static constexpr int M = 32 * L;                 // complex FFT length
    static constexpr int NFFT = 64 * L;
    static constexpr int FW = 32 / L;                // frames per warp
    static constexpr int C = 32 / L;                 // pass-2 columns per lane
    static constexpr int ROWS = 2 * L;               // 16-bin rows of the spectrum
    static constexpr int GB = (L == 16) ? 1104 : 560;   // floats per group region, == 16 (mod 32)
    static constexpr int WB = FW * GB;               // floats per warp buffer
    static constexpr int PEND = 34 * L + 17;         // the kernel keeps group scratch [0, PEND) finite
};

template <int L> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int d = L / 2; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}
template <int L> __device__ __forceinline__ int group_sum_i(int v) {
#pragma unroll
    for (int d = L / 2; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

int sub_smem_bytes(const FastTables& ft, int nwarps, int L, bool tm) {
    const int wb = (L == 16) ? SubGeom<16>::WB : SubGeom<8>::WB;
    return (((2 * nwarps + 3 + (tm ? 4 : 0)) & ~3) + (tm ? 0 : ft.total) + nwarps * wb) * 4;
}

// TM: the per-lane tables (window, twiddles, mel weights and first taps) are read from Tensor Memory with tcgen05.ld
// instead of shared memory, as in frames_fast_2048 (any window: HANN is then irrelevant and instantiated false).
template <int L, int NW, bool HANN, bool TM = false>
__global__ void __launch_bounds__(NW * 32, 1)
frames_sub(const FrameArgs a, const float* __restrict__ g_tables, const FastTables ft) {
    using G = SubGeom<L>;
    extern __shared__ __align__(16) float smem[];
    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lg = lane & (L - 1), grp = lane / L, gbase = grp * L;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem) + warp;
    float* tab = smem + ((2 * NW + 3 + (TM ? 4 : 0)) & ~3);
    const int tab_total = TM ? 0 : ft.total;          // TM: no table is staged in shared memory
    const float2* s_win = reinterpret_cast<const float2*>(tab + ft.win);
    const float2* s_tw1 = reinterpret_cast<const float2*>(tab + ft.tw1);
    const float2* s_tw2 = reinterpret_cast<const float2*>(tab + ft.tw2);
    const int* s_meta = reinterpret_cast<const int*>(tab + ft.mel_meta);
    const float* s_melw = tab + ft.mel_w;
    float* wbuf = tab + tab_total + warp * G::WB;
    float* sc = wbuf + grp * G::GB;                   // this group's region
    float2* sc2 = reinterpret_cast<float2*>(sc);
    float* scp = sc + ((L == 8) ? 8 * (grp >> 1) : 0);   // power-spectrum scratch, skewed so groups hit different banks

    for (int i = tid; i < tab_total / 4; i += NT)
        reinterpret_cast<float4*>(tab)[i] = __ldg(reinterpret_cast<const float4*>(g_tables) + i);
    for (int i = tid; i < NW * G::WB; i += NT) (tab + tab_total)[i] = 0.0f;
    if (lane == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    uint32_t tbase = 0;
    if constexpr (TM) tbase = tmem_tables_setup<NW>(ft.tmem_tab, ft.tmem_cols, reinterpret_cast<uint32_t*>(smem + 2 * NW), warp, lane);
    __syncthreads();
    if constexpr (TM) asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tq = __reduce_or_sync(FULL, tbase + ((uint32_t)(32 * (warp & 3)) << 16));   // uniform register: no R2UR per tcgen05.ld

    // units of FW consecutive frames; a CTA owns a contiguous run of units, warps take them round-robin
    const int units_per_clip = (a.T + G::FW - 1) / G::FW;
    const long long total = (long long)a.B * units_per_clip;
    const long long per_cta = (total + gridDim.x - 1) / gridDim.x;
    const long long c0 = (long long)blockIdx.x * per_cta;
    const long long u1 = (c0 + per_cta < total) ? c0 + per_cta : total;
    uint32_t parity = 0;
    float clip_max = 0.0f;
    const float zthr = a.zcr_thr;
    const int R = ft.n_groups;                        // mel rounds: ceil(n_mels / L)

    const int zw_base = lg;                           // regroup write: k = lg + L*c + 32*k2
    const int zlo_base = 17 * lg;
    const int zhi_base = 17 * (G::ROWS - 1 - lg) + 16;
    const int zhi0 = (lg == 0) ? 0 : 17 * (G::ROWS - lg);

    for (long long u = c0 + warp; u < u1; u += NW) {
        const int b = (int)(u / units_per_clip);
        const int t0 = (int)(u - (long long)b * units_per_clip) * G::FW;
        const int t = min(t0 + grp, a.T - 1);         // groups past the last frame redo it and are masked
        const bool live = (t0 + grp) < a.T;
        const float* clip = a.wave + (long long)b * a.pitch;
        const int fs = t * a.hop - a.pad;
        const bool interior = (fs >= 0) && (fs + G::NFFT <= a.n);
        const uintptr_t addr = reinterpret_cast<uintptr_t>(clip + fs);
        const int shift = (int)((addr & 15) >> 2);
        const bool fast = __all_sync(FULL, interior && !(shift & 1));
        int zc_edge = -1;
        int off;
        if (fast) {
            // one bulk copy per group, all completing on the warp's mbarrier
            const int tot = shift + G::NFFT;
            const int bulk = tot & ~3;
            const float* src0 = reinterpret_cast<const float*>(addr & ~uintptr_t(15));
            if (lg == 0)
                for (int i = bulk; i < tot; ++i) sc[i] = __ldg(src0 + i);
            unsigned bytes = (lg == 0) ? (unsigned)bulk * 4u : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) bytes += __shfl_xor_sync(FULL, bytes, d);
            __syncwarp();
            if (lg == 0) fence_proxy_async_smem();    // every issuing lane orders its earlier generic accesses
            if (lane == 0) mbar_arrive_expect_tx(mbar, bytes);
            __syncwarp();
            if (lg == 0) tma_bulk_g2s(sc, src0, (unsigned)bulk * 4u, mbar);
            mbar_wait(mbar, parity);
            parity ^= 1u;
            off = shift;
        } else {
            // an edge frame somewhere in the warp: every group builds its padded frame by hand
            constexpr int CH = G::NFFT / L;           // contiguous samples per lane
            int zc = 0;
            float prev = sample_edge(clip, a.n, fs + lg * CH - 1);
            for (int i0 = 0; i0 < CH; i0 += 8) {
                float ve[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) ve[u] = sample_edge(clip, a.n, fs + lg * CH + i0 + u);   // 8 loads in flight
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u, sidx = fs + lg * CH + i;
                    const bool inside = (sidx >= 0) && (sidx < a.n);
                    sc[lg * CH + i] = inside ? ve[u] : ((a.pad_mode == 0) ? 0.0f : (a.pad_mode == 1) ? __ldg(clip + reflect_index(sidx, a.n)) : ve[u]);
                    if ((lg * CH + i) > 0) zc += ((ve[u] < -zthr) != (prev < -zthr)) ? 1 : 0;
                    prev = ve[u];
                }
            }
            zc_edge = group_sum_i<L>(zc);
            off = 0;
            __syncwarp();
        }

        float2 v[32];                                   // packed FP32: one complex value per register pair
        float ss;
        unsigned za = 0u, zb = 0u;
        if constexpr (TM) {
            const float2* xp = reinterpret_cast<const float2*>(sc + off);
            const float2 zt = make_float2(zthr, zthr);
            float2 ss2 = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int jc = 0; jc < 4; ++jc) {
                uint32_t wr[16];
                tmem_ld16_issue(tq + kSubWin + 16 * jc, wr);
                float2 x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = xp[lg + L * (8 * jc + u)];
                tmem_wait16(wr);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    ss2 = __ffma2_rn(x[u], x[u], ss2);
                    const float2 xt = __fadd2_rn(x[u], zt);
                    za = __funnelshift_l(__float_as_uint(xt.x), za, 1);
                    zb = __funnelshift_l(__float_as_uint(xt.y), zb, 1);
                    v[8 * jc + u] = __fmul2_rn(x[u], f2_of(wr[2 * u], wr[2 * u + 1]));
                }
            }
            ss = ss2.x + ss2.y;
        } else {
            const float2* xp = reinterpret_cast<const float2*>(sc + off);
            const float2 zt = make_float2(zthr, zthr);
            float2 ss2 = make_float2(0.0f, 0.0f);
            // HANN: 0.5 * w[2 (lg + L j) + c] = 0.25 - 0.25 cos(theta_lg,c + 2 pi j / 32), synthesised as in
            // frames_fast_2048 instead of read from the window table
            float2 hc = make_float2(0.f, 0.f), hs = hc;
            if (HANN) {
                const float4 cs = reinterpret_cast<const float4*>(tab + ft.hann_cs)[lg];
                hc = make_float2(cs.x, cs.y); hs = make_float2(cs.z, cs.w);
            }
            const float2 quarter = make_float2(0.25f, 0.25f);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float2 x = xp[lg + L * j];
                float2 w;
                if (HANN) {
                    const float cj = -0.25f * float(fftreg::cos2pi(j, 32)), sj = 0.25f * float(fftreg::sin2pi(j, 32));
                    if (j == 0 || j == 16) w = __ffma2_rn(hc, make_float2(cj, cj), quarter);
                    else if (j == 8 || j == 24) w = __ffma2_rn(hs, make_float2(sj, sj), quarter);
                    else w = __ffma2_rn(hc, make_float2(cj, cj), __ffma2_rn(hs, make_float2(sj, sj), quarter));
                } else {
                    w = s_win[lg + L * j];
                }
                ss2 = __ffma2_rn(x, x, ss2);
                const float2 xt = __fadd2_rn(x, zt);
                za = __funnelshift_l(__float_as_uint(xt.x), za, 1);
                zb = __funnelshift_l(__float_as_uint(xt.y), zb, 1);
                v[j] = __fmul2_rn(x, w);
            }
            ss = ss2.x + ss2.y;
        }
        int zc;
        {
            unsigned zn = __shfl_sync(FULL, za, gbase + ((lg + 1) & (L - 1)));
            unsigned msk = FULL;
            if (lg == L - 1) { zn <<= 1; msk = 0xfffffffeu; }
            zc = __popc(za ^ zb) + __popc((zb ^ zn) & msk);
            zc = group_sum_i<L>(zc);
            if (zc_edge >= 0) zc = zc_edge;
        }
        ss = group_sum<L>(ss);
        __syncwarp();

        // pass 1: 32-point FFT over n1 (z[lg + L*n1]); twiddle W_M^(lg*k1)
        fftreg2::fft_dif<32>(v);
        if constexpr (TM) {
#pragma unroll
            for (int kc = 0; kc < 4; kc += 2) {
                uint32_t tr[16], tr2[16];
                tmem_ld16_issue(tq + kSubTw1 + 16 * kc, tr);
                tmem_ld16_issue(tq + kSubTw1 + 16 * kc + 16, tr2);
                tmem_wait16(tr);
                tmem_wait16(tr2);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int k1 = 8 * kc + u + 1;
                    if (k1 < 32) {
                        const int p = pos32(k1);
                        const uint32_t c = (u < 8) ? tr[2 * (u & 7)] : tr2[2 * (u & 7)];
                        const uint32_t sn = (u < 8) ? tr[2 * (u & 7) + 1] : tr2[2 * (u & 7) + 1];
                        v[p] = fftreg2::cmul(v[p], __uint_as_float(c), __uint_as_float(sn));
                    }
                }
            }
        } else {
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) {
                const float2 w = s_tw1[(k1 - 1) * L + lg];
                const int p = pos32(k1);
                v[p] = fftreg2::cmul(v[p], w.x, w.y);
            }
        }
        // transpose inside the group: lane lg then owns columns k1 = lg + L*c
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) sc2[lg * 33 + k1] = v[pos32(k1)];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < G::C; ++c)
#pragma unroll
            for (int n2 = 0; n2 < L; ++n2) v[c * L + n2] = sc2[n2 * 33 + lg + L * c];
        __syncwarp();
        // pass 2: an L-point FFT per column; bin k = (lg + L*c) + 32*k2 sits at c*L + fft_pos<L>(k2)
        if constexpr (L == 16) {
            fftreg2::fft_dif<16, 0>(v);
            fftreg2::fft_dif<16, 16>(v);
        } else {
            fftreg2::fft_dif<8, 0>(v);
            fftreg2::fft_dif<8, 8>(v);
            fftreg2::fft_dif<8, 16>(v);
            fftreg2::fft_dif<8, 24>(v);
        }
        // regroup (layout p(k) = k + k/16): lane lg owns bins [16*lg, 16*lg+16) and mirrors M-k
#pragma unroll
        for (int c = 0; c < G::C; ++c)
#pragma unroll
            for (int k2 = 0; k2 < L; ++k2) {
                const int src = c * L + fftreg::fft_pos<L>(k2);
                sc2[zw_base + (L * c + ((L * c) >> 4)) + 34 * k2] = v[src];
            }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = sc2[zlo_base + i];
        v[16] = sc2[zhi0];
#pragma unroll
        for (int i = 1; i < 16; ++i) v[16 + i] = sc2[zhi_base - i];
        const float2 emid = sc2[17 * L];
        __syncwarp();

        // real-FFT split, |X|^2, |X|, moments of |X| about the centres of the lane's two runs
        float2 P[16], S[16];
        float2 M0 = make_float2(0.f, 0.f), M1 = M0, M2 = M0;
        uint32_t t2a[16], t2b[16];
        if constexpr (TM) {
            tmem_ld16_issue(tq + kSubTw2, t2a);
            tmem_ld16_issue(tq + kSubTw2 + 16, t2b);
            tmem_wait16(t2a);
            tmem_wait16(t2b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 w = TM ? (i < 8 ? f2_of(t2a[2 * i], t2a[2 * i + 1]) : f2_of(t2b[2 * (i - 8)], t2b[2 * (i - 8) + 1]))
                                : s_tw2[i * L + lg];
            const float2 za_ = v[i], zb_ = v[16 + i];
            const float2 e = __fadd2_rn(za_, make_float2(zb_.x, -zb_.y));
            const float2 d = __fadd2_rn(za_, make_float2(-zb_.x, zb_.y));
            const float2 tt = fftreg2::cmul(d, w.x, w.y);
            const float2 xa = __fadd2_rn(e, tt), xb = __fadd2_rn(e, make_float2(-tt.x, -tt.y));
            const float2 pw = make_float2(fmaf(xa.x, xa.x, xa.y * xa.y), fmaf(xb.x, xb.x, xb.y * xb.y));
            const float2 sq = make_float2(fast_sqrt(pw.x), fast_sqrt(pw.y));
            P[i] = pw;
            S[i] = sq;
            const float dd = float(i) - 7.5f;
            M0 = __fadd2_rn(M0, sq);
            M1 = __ffma2_rn(make_float2(sq.x, -sq.y), make_float2(dd, dd), M1);
            M2 = __ffma2_rn(sq, make_float2(dd * dd, dd * dd), M2);
        }
        const float m0l = M0.x, m0h = M0.y, m1l = M1.x, m1h = M1.y, m2l = M2.x, m2h = M2.y;
        const float pmid = 4.0f * fmaf(emid.x, emid.x, emid.y * emid.y);
        const float smid = fast_sqrt(pmid);

        // centroid / bandwidth (moments about the run centres, as in frames_fast_2048)
        const float fM = float(G::M);
        const float kcl = 16.0f * lg + 7.5f;
        const float kch = fM - 16.0f * lg - 7.5f;
        float s0 = m0l + m0h;
        float s1 = fmaf(kcl, m0l, m1l) + fmaf(kch, m0h, m1h);
        if (lg == L - 1) { s0 += smid; s1 = fmaf(0.5f * fM, smid, s1); }
        s0 = group_sum<L>(s0);
        s1 = group_sum<L>(s1);
        const float denom = (s0 < 1.17549435e-38f) ? 1.0f : s0;
        const float cen = s1 / denom;
        float q;
        {
            const float dl = kcl - cen, dh = kch - cen;
            q = fmaf(dl, fmaf(dl, m0l, 2.0f * m1l), m2l) + fmaf(dh, fmaf(dh, m0h, 2.0f * m1h), m2h);
            if (lg == L - 1) { const float dm = 0.5f * fM - cen; q = fmaf(dm * dm, smid, q); }
            q = group_sum<L>(q);
        }
        const float bw = sqrtf(fmaxf(q, 0.0f) / denom);

        // rolloff
        int rbin;
        {
            constexpr unsigned GM = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
            float pl = m0l;
#pragma unroll
            for (int d = 1; d < L; d <<= 1) {
                const float tv = __shfl_up_sync(FULL, pl, d);
                if (lg >= d) pl += tv;
            }
            float ph = m0h;
#pragma unroll
            for (int d = 1; d < L; d <<= 1) {
                const float tv = __shfl_down_sync(FULL, ph, d);
                if (lg + d < L) ph += tv;
            }
            const float tot_lo = __shfl_sync(FULL, pl, gbase + L - 1);
            const float tot_hi = __shfl_sync(FULL, ph, gbase);
            const float mid = tot_lo + smid;
            const float thr = a.roll_percent * (mid + tot_hi);
            const unsigned lo_mask = (__ballot_sync(FULL, pl >= thr) >> gbase) & GM;
            const unsigned hi_mask = (__ballot_sync(FULL, mid + ph >= thr) >> gbase) & GM;
            // both searches run in every lane (groups may take different branches; keep shuffles converged)
            float cum = pl - m0l;
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) { cum += S[i].x; cnt += (cum < thr) ? 1 : 0; }
            const int cand_lo = 16 * lg + min(cnt, 15);
            cum = mid + (ph - m0h);
            cnt = 0;
#pragma unroll
            for (int i = 15; i >= 0; --i) { cum += S[i].y; cnt += (cum < thr) ? 1 : 0; }
            const int cand_hi = G::M - 16 * lg - (15 - min(cnt, 15));
            const int tl_lo = lo_mask ? (__ffs(lo_mask) - 1) : 0;
            const int tl_hi = hi_mask ? (31 - __clz(hi_mask)) : 0;
            const int r_lo = __shfl_sync(FULL, cand_lo, gbase + tl_lo);
            const int r_hi = __shfl_sync(FULL, cand_hi, gbase + tl_hi);
            if (lo_mask) rbin = r_lo;
            else if (mid >= thr) rbin = G::M / 2;
            else if (hi_mask) rbin = r_hi;
            else rbin = G::M;
        }

        // power (or magnitude) spectrum -> padded scratch of the group
        {
            const bool mag = a.use_mag != 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) scp[17 * lg + i] = mag ? S[i].x : P[i].x;
            scp[17 * (G::ROWS - lg)] = mag ? S[0].y : P[0].y;
#pragma unroll
            for (int i = 1; i < 16; ++i) scp[17 * (G::ROWS - 1 - lg) + 16 - i] = mag ? S[i].y : P[i].y;
            if (lg == L - 1) scp[17 * L] = a.use_mag ? smid : pmid;
            scp[17 * lg + 16] = 0.0f;
            scp[17 * (G::ROWS - 1 - lg) + 16] = 0.0f;
            scp[34 * L + 1 + lg] = 0.0f;              // taps past the last bin read zeros
            if (L == 8) scp[34 * L + 9 + lg] = 0.0f;
        }
        __syncwarp();

        // banded mel gather: L filters per round, one per lane of the group
        if (a.mel_out != nullptr) {
            float wmax = 0.0f;
            const size_t mstride = a.mel_frame_major ? 1 : (size_t)a.T;
            float* outb = a.mel_frame_major ? a.mel_out + ((size_t)b * a.T + t) * a.n_mels
                                            : a.mel_out + ((size_t)b * a.n_mels) * a.T + t;
            uint32_t col = tq + kSubMel;
            for (int r = 0; r < R; ++r) {
                float2 a01 = make_float2(0.0f, 0.0f), a23 = a01;
                if constexpr (TM) {
                    const int n4 = ft.mel_steps[r];
                    uint32_t st1[4];
                    tmem_ld4(tq + kSubMeta + (r & ~3), st1);      // first taps of rounds 4*(r/4) .. +3
                    const uint32_t start = (r & 2) ? ((r & 1) ? st1[3] : st1[2]) : ((r & 1) ? st1[1] : st1[0]);
                    mel_steps_tm(n4, col, scp + start, a01, a23);
                    col += 4 * n4;
                } else {
                const int n4 = s_meta[r];
                const float4* wq = reinterpret_cast<const float4*>(s_melw + s_meta[R + r]) + lg;
                const float* pq = scp + s_meta[2 * R + r * L + lg];
                mel_steps<L>(n4, wq, pq, a01, a23);
                }
                a01 = __fadd2_rn(a01, a23);
                const float acc = a01.x + a01.y;
                const int m = L * r + lg;
                if (live && m < a.n_mels) outb[(size_t)m * mstride] = acc;
                if (live) wmax = fmaxf(wmax, acc);
            }
            clip_max = fmaxf(clip_max, wmax);
        }
        if (lg == 0 && live) {
            if (a.stats != nullptr) {
                float* st = a.stats + (size_t)b * 5 * a.T + t;
                st[0] = cen * a.binhz;
                st[(size_t)a.T] = bw * a.binhz;
                st[(size_t)2 * a.T] = float(rbin) * a.binhz;
                st[(size_t)3 * a.T] = float(zc) * (1.0f / float(G::NFFT));
                st[(size_t)4 * a.T] = sqrtf(ss * (1.0f / float(G::NFFT)));
            }
            if (a.status != nullptr && !(fabsf(ss) <= 3.0e38f)) atomicOr(a.status + b, 1);
        }
        // flush the running mel maximum when this warp's next unit belongs to another clip
        const long long un = u + NW;
        const bool last = (un >= u1) || ((int)(un / units_per_clip) != b);
        if (last && a.clipmax != nullptr && a.mel_out != nullptr) {
            const float m = warp_max(clip_max);
            if (lane == 0) atomicMax(reinterpret_cast<int*>(a.clipmax) + b, __float_as_int(m));
            clip_max = 0.0f;
        }
        __syncwarp();
    }
    if constexpr (TM) tmem_tables_release(tbase, warp);
}

template <int L, bool HANN, bool TM = false>
static cudaError_t launch_sub(const FrameArgs& a, const float* d_tables, const FastTables& ft, int num_sms,
                              cudaStream_t stream) {
    constexpr int NW = 16;
    const int smem = sub_smem_bytes(ft, NW, L, TM);
    cudaError_t e = cudaFuncSetAttribute(frames_sub<L, NW, HANN, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const long long units = (long long)a.B * ((a.T + SubGeom<L>::FW - 1) / SubGeom<L>::FW);
    if (units <= 0) return cudaSuccess;
    long long grid = (units + NW - 1) / NW;
    if (grid > num_sms) grid = num_sms;
    frames_sub<L, NW, HANN, TM><<<(unsigned)grid, NW * 32, smem, stream>>>(a, d_tables, ft);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_frames_sub(const FrameArgs& a, const float* d_tables, const FastTables& ft, int num_sms,
                              cudaStream_t stream) {
    static const bool no_tm = [] { const char* e = getenv("HLMC_NO_TMEM"); return e && e[0] == '1'; }();
    if (!no_tm && !a.no_tmem && ft.tmem_tab != nullptr && ft.n_groups <= 32) {
        if (a.n_fft == 1024) return launch_sub<16, false, true>(a, d_tables, ft, num_sms, stream);
        if (a.n_fft == 512) return launch_sub<8, false, true>(a, d_tables, ft, num_sms, stream);
    }
    if (a.n_fft == 1024) return ft.hann ? launch_sub<16, true>(a, d_tables, ft, num_sms, stream)
                                        : launch_sub<16, false>(a, d_tables, ft, num_sms, stream);
    if (a.n_fft == 512) return ft.hann ? launch_sub<8, true>(a, d_tables, ft, num_sms, stream)
                                       : launch_sub<8, false>(a, d_tables, ft, num_sms, stream);
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------
// librosa.estimate_tuning on the piptrack candidates: median of the magnitudes (exact, by a
// three-level radix select on the float bits), then the 100-bin histogram of the pitch
// residuals of the candidates at or above the median; the tuning is the left edge of the
// fullest bin.  One CTA per clip.
// ---------------------------------------------------------------------------
constexpr int kTunThreads = 256;

// Walks every candidate of the clip: one warp per frame, lanes over that frame's list.
// Visits every piptrack candidate of one clip: a warp takes four frames at a time so that four independent
// loads are in flight per lane (the kernel is bound by the latency of these loads: ncu, 12 long-scoreboard
// stall cycles per issue with one frame at a time).
template <class F>
__device__ __forceinline__ void for_each_candidate(const float2* __restrict__ c, const int* __restrict__ cnt,
                                                   int T, int cpf, F f) {
    const int ln = threadIdx.x & 31, wp = threadIdx.x >> 5;
    constexpr int W = kTunThreads / 32, U = 4;
    for (int t0 = wp; t0 < T; t0 += U * W) {
        int m[U], mmax = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 + u * W;
            m[u] = (t < T) ? __ldg(cnt + t) : 0;
            mmax = max(mmax, m[u]);
        }
        for (int j = ln; j < mmax; j += 32) {
            float2 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (j < m[u]) v[u] = __ldg(c + (size_t)(t0 + u * W) * cpf + j);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (j < m[u]) f(v[u]);
        }
    }
}

// Rank search in a digit histogram: every thread owns nb / kTunThreads consecutive bins; block-wide exclusive
// scan of the per-thread sums (warp shuffles + one shared hop), then the owner of rank k walks its bins.
// Writes the digit to out[0] and the rank inside that bin to out[1].  Contains block barriers.
__device__ void digit_search(const unsigned* hist, int nb, unsigned k, unsigned* s_warp, unsigned* out) {
    const int per = nb / kTunThreads;                 // 8 or 4
    unsigned mine = 0u;
    for (int j = 0; j < per; ++j) mine += hist[threadIdx.x * per + j];
    unsigned incl = mine;
    const int ln = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned tv = __shfl_up_sync(FULL, incl, d);
        if (ln >= d) incl += tv;
    }
    __syncthreads();                                  // s_warp may still be read by the previous search
    if (ln == 31) s_warp[wp] = incl;
    __syncthreads();
    unsigned off = 0u;
    for (int q = 0; q < wp; ++q) off += s_warp[q];
    const unsigned excl = off + incl - mine;
    if (k >= excl && k < excl + mine) {
        unsigned cum = excl;
        int d = threadIdx.x * per;
        for (;; ++d) {
            if (cum + hist[d] > k) break;
            cum += hist[d];
        }
        out[0] = (unsigned)d;
        out[1] = k - cum;
    }
}

// The k0-th and k1-th smallest magnitudes (k0 <= k1, usually adjacent: the two middle elements np.median
// averages) by ONE radix descent: keys are the IEEE bits of the (positive) magnitudes, digits of 11 + 11 + 10
// bits; while both ranks share their prefix one histogram serves both, afterwards each candidate is counted
// into the histogram of the prefix it matches - three passes over the candidates in either case.
__device__ void select_two(const float2* __restrict__ c, const int* __restrict__ cnt, int T, int cpf,
                           unsigned k0, unsigned k1, unsigned* hist /* 2 x 2048 */, unsigned* s_misc /* 4 */,
                           unsigned* s_warp, unsigned& bits0, unsigned& bits1) {
    unsigned prefix0 = 0u, prefix1 = 0u, mask = 0u;
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int lvl = 0; lvl < 3; ++lvl) {
        const int sh = shifts[lvl], nb = 1 << widths[lvl];
        const bool split = (prefix0 != prefix1);
        for (int i = threadIdx.x; i < nb; i += kTunThreads) { hist[i] = 0u; hist[2048 + i] = 0u; }
        __syncthreads();
        for_each_candidate(c, cnt, T, cpf, [&](const float2 pm) {
            const unsigned key = __float_as_uint(pm.y), pk = key & mask, d = (key >> sh) & (nb - 1);
            if (pk == prefix0) atomicAdd(&hist[d], 1u);
            else if (split && pk == prefix1) atomicAdd(&hist[2048 + d], 1u);
        });
        __syncthreads();
        digit_search(hist, nb, k0, s_warp, s_misc);
        digit_search(split ? hist + 2048 : hist, nb, k1, s_warp, s_misc + 2);
        __syncthreads();
        prefix0 |= s_misc[0] << sh;
        prefix1 |= s_misc[2] << sh;
        mask |= (unsigned)(nb - 1) << sh;
        k0 = s_misc[1];
        k1 = s_misc[3];
        __syncthreads();
    }
    bits0 = prefix0;
    bits1 = prefix1;
}

__global__ void __launch_bounds__(kTunThreads)
tuning_kernel(const float2* __restrict__ cand, const int* __restrict__ cand_count, int T, int cpf,
              const double* __restrict__ edges, float* __restrict__ tuning, int* __restrict__ tuning_idx) {
    __shared__ unsigned hist[2 * 2048];
    __shared__ unsigned s_misc[4];
    __shared__ unsigned s_warp[kTunThreads / 32];
    __shared__ unsigned counts[kTuningBins];
    __shared__ int s_n;
    const long long b = blockIdx.x;
    const float2* c = cand + (size_t)b * T * cpf;
    const int* cnt = cand_count + (size_t)b * T;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    {
        int part = 0;
        for (int t = threadIdx.x; t < T; t += kTunThreads) part += cnt[t];
        part = warp_sum_i(part);
        if ((threadIdx.x & 31) == 0 && part) atomicAdd(&s_n, part);
    }
    __syncthreads();
    const int n = s_n;
    if (n <= 0) {            // pitch_tuning: no pitches -> 0.0 (bin 50 of linspace(-0.5, 0.5, 101))
        if (threadIdx.x == 0) { if (tuning) tuning[b] = 0.0f; tuning_idx[b] = kTuningBins / 2; }
        return;
    }
    // np.median
    unsigned lo_bits, hi_bits;
    select_two(c, cnt, T, cpf, (unsigned)((n & 1) ? n / 2 : n / 2 - 1), (unsigned)(n / 2), hist, s_misc, s_warp,
               lo_bits, hi_bits);
    float thr = __uint_as_float(hi_bits);
    if ((n & 1) == 0) thr = (__uint_as_float(lo_bits) + thr) * 0.5f;
    for (int i = threadIdx.x; i < kTuningBins; i += kTunThreads) counts[i] = 0u;
    __syncthreads();
    for_each_candidate(c, cnt, T, cpf, [&](const float2 pm) {
        if (!(pm.y >= thr)) return;
        // residual of 12 * log2(f / 27.5) modulo one semitone, folded to [-0.5, 0.5)
        float r = 12.0f * log2f(pm.x / 27.5f);
        r = r - floorf(r);
        if (r >= 0.5f) r -= 1.0f;
        // np.histogram with 100 uniform bins on [-0.5, 0.5]
        const double x = (double)r;
        int idx = (int)((x - edges[0]) / (edges[kTuningBins] - edges[0]) * (double)kTuningBins);
        if (idx >= kTuningBins) idx = kTuningBins - 1;
        if (idx < 0) idx = 0;
        if (x < edges[idx] && idx > 0) --idx;
        else if (idx != kTuningBins - 1 && x >= edges[idx + 1]) ++idx;
        atomicAdd(&counts[idx], 1u);
    });
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 0;
        for (int i = 1; i < kTuningBins; ++i)
            if (counts[i] > counts[best]) best = i;
        if (tuning) tuning[b] = (float)edges[best];
        tuning_idx[b] = best;
    }
}
cudaError_t launch_tuning(const float2* cand, const int* cand_count, int T, int cand_per_frame, long long B,
                          const double* d_edges, float* tuning, int* tuning_idx, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    tuning_kernel<<<(unsigned)B, kTunThreads, 0, stream>>>(cand, cand_count, T, cand_per_frame, d_edges, tuning,
                                                           tuning_idx);
    g_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// librosa.feature.chroma_stft: chroma filterbank (of the clip's estimated tuning) applied to
// the power spectrogram, every frame divided by its largest chroma bin.  A CTA works on one
// clip at a time so that the clip's 12 x 1025 filterbank sits in shared memory; its warps
// take the clip's frames round-robin through the same frame_spectrum() front end.
// ---------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1)
chroma_fast_2048(const FrameArgs a, const ChromaArgs ca, const float* __restrict__ g_tables,
                 const FastTables ft) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntab = ft.total;                         // the whole table blob (twiddles ... window)
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem) + warp;
    float* tab = smem + ((2 * NW + 3) & ~3);
    float* fbs = tab + ntab;                           // this clip's filterbank
    float* sc = fbs + kChromaFbFloats + warp * kWarpBufFloats4;
    for (int i = tid; i < ntab / 4; i += NT)
        reinterpret_cast<float4*>(tab)[i] = __ldg(reinterpret_cast<const float4*>(g_tables) + i);
    for (int i = tid; i < NW * kWarpBufFloats4; i += NT) (fbs + kChromaFbFloats)[i] = 0.0f;
    if (lane == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    __syncthreads();

    WarpState w;
    w.sc = sc; w.sc2 = reinterpret_cast<float2*>(sc); w.mbar = mbar; w.parity = 0;
    w.s_win = reinterpret_cast<const float2*>(tab + ft.win);
    w.s_tw1 = reinterpret_cast<const float2*>(tab + ft.tw1);
    w.s_tw2 = reinterpret_cast<const float2*>(tab + ft.tw2);
    w.s_hcs = nullptr; w.pending = false; w.pend_off = 0;
    w.lane = lane;
    w.zw_base = lane + (lane >> 4);
    w.zlo_base = 17 * lane;
    w.zhi_base = 17 * (63 - lane) + 16;
    w.zhi0 = (lane == 0) ? 0 : 17 * (64 - lane);

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const float* src = ca.fb_all + (size_t)ca.tuning_idx[b] * kChromaFbFloats;
        for (int i = tid; i < kChromaFbFloats / 4; i += NT)
            reinterpret_cast<float4*>(fbs)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        __syncthreads();
        const float* clip = a.wave + (long long)b * a.pitch;
        for (int t = warp; t < a.T; t += NW) {
            float2 P[16], S[16];
            float p512, s512, ss, m0l, m1l, m2l, m0h, m1h, m2h;
            int zc;
            frame_spectrum<false>(a, w, clip, t, nullptr, 0, P, S, p512, s512, ss, zc, m0l, m1l, m2l, m0h, m1h, m2h);
            // filterbank column j of this lane: j < 16 -> bin 16*lane + j, else bin 1024 - 16*lane - (j - 16)
            auto pw = [&](int j) -> float { return j < 16 ? P[j].x : P[j - 16].y; };
            float raw[kChroma];
#pragma unroll
            for (int c = 0; c < kChroma; ++c) {
                const float4* fp = reinterpret_cast<const float4*>(fbs) + (c * 8) * 32 + lane;
                float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 f = fp[j4 * 32];
                    a0 = fmaf(f.x, pw(4 * j4 + 0), a0);
                    a1 = fmaf(f.y, pw(4 * j4 + 1), a1);
                    a0 = fmaf(f.z, pw(4 * j4 + 2), a0);
                    a1 = fmaf(f.w, pw(4 * j4 + 3), a1);
                }
                float r = a0 + a1;
                if (lane == 31) r = fmaf(fbs[kChroma * 32 * 32 + c], p512, r);
                raw[c] = warp_sum(r);
            }
            float mx = 0.0f;
#pragma unroll
            for (int c = 0; c < kChroma; ++c) mx = fmaxf(mx, fabsf(raw[c]));
            const float inv = (mx < 1.17549435e-38f) ? 1.0f : 1.0f / mx;   // util.normalize(norm=inf)
            float mine = raw[0];
#pragma unroll
            for (int c = 1; c < kChroma; ++c) mine = (lane == c) ? raw[c] : mine;
            if (lane < kChroma) ca.chroma[((size_t)b * kChroma + lane) * a.T + t] = mine * inv;
            __syncwarp();
        }
        __syncthreads();                       // the filterbank is replaced for the next clip
    }
}

// chroma projection from the stashed power spectrum (the default chroma path): one CTA per clip keeps the
// clip's filterbank in shared memory; a warp takes two frames at a time so that every 16-byte filterbank
// read feeds four FFMA2 (the filterbank read, 49 KB per frame, is what bounds this kernel).
constexpr int kProjWarps = 8;
__global__ void __launch_bounds__(kProjWarps * 32, 2)
chroma_project(const float* __restrict__ pstash, const ChromaArgs ca, long long B, int T) {
    extern __shared__ __align__(16) float fbs[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        const float* src = ca.fb_all + (size_t)ca.tuning_idx[b] * kChromaFbFloats;
        for (int i = tid; i < kChromaFbFloats / 4; i += kProjWarps * 32)
            reinterpret_cast<float4*>(fbs)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        __syncthreads();
        for (int t0 = 2 * warp; t0 < T; t0 += 2 * kProjWarps) {
            const bool two = (t0 + 1) < T;
            const float* fa = pstash + ((size_t)b * T + t0) * kStashFloats;
            const float* fb = two ? fa + kStashFloats : fa;
            float4 pa[8], pb[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                pa[q] = __ldg(reinterpret_cast<const float4*>(fa) + lane * 8 + q);
                pb[q] = __ldg(reinterpret_cast<const float4*>(fb) + lane * 8 + q);
            }
            const float pa512 = __ldg(fa + 32 * 32), pb512 = __ldg(fb + 32 * 32);
            float rawa[16], rawb[16];
#pragma unroll
            for (int c = kChroma; c < 16; ++c) { rawa[c] = 0.0f; rawb[c] = 0.0f; }
#pragma unroll
            for (int c = 0; c < kChroma; ++c) {
                const float4* fp = reinterpret_cast<const float4*>(fbs) + (c * 8) * 32 + lane;
                float2 aa = make_float2(0.f, 0.f), ab = aa;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 f = fp[j4 * 32];
                    aa = __ffma2_rn(make_float2(f.x, f.y), make_float2(pa[j4].x, pa[j4].y), aa);
                    aa = __ffma2_rn(make_float2(f.z, f.w), make_float2(pa[j4].z, pa[j4].w), aa);
                    ab = __ffma2_rn(make_float2(f.x, f.y), make_float2(pb[j4].x, pb[j4].y), ab);
                    ab = __ffma2_rn(make_float2(f.z, f.w), make_float2(pb[j4].z, pb[j4].w), ab);
                }
                float ra = aa.x + aa.y, rb = ab.x + ab.y;
                if (lane == 31) {
                    const float wm = fbs[kChroma * 32 * 32 + c];
                    ra = fmaf(wm, pa512, ra);
                    rb = fmaf(wm, pb512, rb);
                }
                rawa[c] = ra;
                rawb[c] = rb;
            }
            // lane l ends up with chroma bin l >> 1 of both frames (bins 12..15 are padding)
            const float minea = warp_sum16(rawa, lane), mineb = warp_sum16(rawb, lane);
            const float mxa = warp_max(fabsf(minea)), mxb = warp_max(fabsf(mineb));
            const float inva = (mxa < 1.17549435e-38f) ? 1.0f : 1.0f / mxa;   // util.normalize(norm=inf)
            const float invb = (mxb < 1.17549435e-38f) ? 1.0f : 1.0f / mxb;
            if (!(lane & 1) && (lane >> 1) < kChroma) {
                float* o = ca.chroma + ((size_t)b * kChroma + (lane >> 1)) * T + t0;
                o[0] = minea * inva;
                if (two) o[1] = mineb * invb;
            }
        }
        __syncthreads();                       // the filterbank is replaced for the next clip
    }
}

cudaError_t launch_chroma_project(const float* pstash, const ChromaArgs& c, long long B, int T, int num_sms,
                                  cudaStream_t stream) {
    const int smem = kChromaFbFloats * 4;
    cudaError_t e = cudaFuncSetAttribute(chroma_project, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    if (B <= 0 || T <= 0) return cudaSuccess;
    long long grid = 2LL * num_sms;
    if (grid > B) grid = B;
    chroma_project<<<(unsigned)grid, kProjWarps * 32, smem, stream>>>(pstash, c, B, T);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_chroma_fast(const FrameArgs& a, const ChromaArgs& c, const float* d_tables,
                               const FastTables& ft, int num_sms, cudaStream_t stream) {
    constexpr int NW = 16;
    const int smem = (((2 * NW + 3) & ~3) + ft.total + kChromaFbFloats + NW * kWarpBufFloats4) * 4;
    cudaError_t e = cudaFuncSetAttribute(chroma_fast_2048<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    if (a.B <= 0) return cudaSuccess;
    const int grid = a.B < num_sms ? a.B : num_sms;
    chroma_fast_2048<NW><<<grid, NW * 32, smem, stream>>>(a, c, d_tables, ft);
    g_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Generic path: any power-of-two n_fft, one warp per frame, FFT in shared memory.
// Also serves librosa.stft (writes the complex spectrum when a.spec != NULL).
// ---------------------------------------------------------------------------
constexpr int kGenWarps = 4;

__global__ void __launch_bounds__(kGenWarps * 32)
frames_generic(const FrameArgs a, const GenericTables gt, int logM) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = a.n_fft >> 1, F = M + 1;
    float2* Z = reinterpret_cast<float2*>(smem) + (size_t)warp * M;
    float* Pb = smem + (size_t)kGenWarps * M * 2 + (size_t)warp * (F + 3);

    const long long frame_id = (long long)blockIdx.x * kGenWarps + warp;
    if (frame_id >= (long long)a.B * a.T) return;      // whole warp exits together
    const int b = (int)(frame_id / a.T);
    const int t = (int)(frame_id - (long long)b * a.T);
    const float* clip = a.wave + (long long)b * a.pitch;
    const int fs = t * a.hop - a.pad;

    // ZCR (edge padded, librosa.feature.zero_crossing_rate) and RMS (pad_mode, feature.rms)
    float ss = 0.0f;
    int zc = 0;
    unsigned prev_last = 0u;
    for (int c = 0; c < a.n_fft / 32; ++c) {
        const int s = fs + 32 * c + lane;
        const float x = sample_padded(clip, a.n, s, a.pad_mode);
        const float e = sample_edge(clip, a.n, s);
        ss = fmaf(x, x, ss);
        const unsigned msk = __ballot_sync(FULL, e < -a.zcr_thr);
        zc += __popc((msk ^ (msk >> 1)) & 0x7fffffffu);
        if (c > 0) zc += ((msk & 1u) != prev_last) ? 1 : 0;
        prev_last = msk >> 31;
    }
    ss = warp_sum(ss);

    // windowed frame packed as M complex points, bit-reversed for the DIT passes
    for (int m = lane; m < M; m += 32) {
        const float x0 = sample_padded(clip, a.n, fs + 2 * m, a.pad_mode);
        const float x1 = sample_padded(clip, a.n, fs + 2 * m + 1, a.pad_mode);
        const int r = (int)(__brev((unsigned)m) >> (32 - logM));
        Z[r] = make_float2(x0 * gt.win[2 * m], x1 * gt.win[2 * m + 1]);
    }
    __syncwarp();
    for (int s = 1; s <= logM; ++s) {
        const int half = 1 << (s - 1);
        const int tstride = M >> s;
        for (int idx = lane; idx < M / 2; idx += 32) {
            const int j = idx & (half - 1);
            const int i0 = ((idx - j) << 1) + j, i1 = i0 + half;
            const float2 w = gt.twm[j * tstride];
            const float2 u = Z[i0], v = Z[i1];
            const float tr = fmaf(v.x, w.x, -(v.y * w.y));
            const float ti = fmaf(v.x, w.y, v.y * w.x);
            Z[i0] = make_float2(u.x + tr, u.y + ti);
            Z[i1] = make_float2(u.x - tr, u.y - ti);
        }
        __syncwarp();
    }
    // split into the 1 + n_fft/2 real-FFT bins; power and magnitude
    float s0 = 0.0f, s1 = 0.0f;
    for (int k = lane; k <= M / 2; k += 32) {
        const float2 za = Z[k], zb = Z[(M - k) & (M - 1)];
        const float2 w = gt.tws[k];
        const float ex = za.x + zb.x, ey = za.y - zb.y, dx = za.x - zb.x, dy = za.y + zb.y;
        const float tx = fmaf(w.x, dx, -(w.y * dy)), ty = fmaf(w.x, dy, w.y * dx);
        const float ar = ex + tx, ai = ey + ty, br = ex - tx, bi = -(ey - ty);
        const float pk = fmaf(ar, ar, ai * ai), pm = fmaf(br, br, bi * bi);
        Pb[k] = pk;
        Pb[M - k] = pm;
        if (a.spec != nullptr) {
            float2* sp = reinterpret_cast<float2*>(a.spec) + (size_t)b * F * a.T + t;
            sp[(size_t)k * a.T] = make_float2(ar, ai);
            sp[(size_t)(M - k) * a.T] = make_float2(br, bi);
        }
    }
    __syncwarp();
    // ---- librosa.piptrack candidates for chroma_stft's tuning estimate (any n_fft; same arithmetic as the
    //      epilogue of frames_fast_2048<*, true>)
    if (a.cand != nullptr) {
        float pmax = 0.0f;
        for (int k = lane; k < F; k += 32) pmax = fmaxf(pmax, Pb[k]);
        pmax = warp_max(pmax);
        const float ref = a.pip_threshold * pmax;
        float2* slots = a.cand + ((size_t)b * a.T + t) * a.cand_cap;
        int base = 0;
        for (int k0 = a.pip_klo; k0 < a.pip_khi; k0 += 32) {
            const int k = k0 + lane;
            const bool in = k < a.pip_khi;
            const int kk = in ? k : a.pip_klo;
            const float pm1 = Pb[kk - 1], p0 = Pb[kk], pp1 = Pb[kk + 1];
            const float xm = (pm1 > ref) ? pm1 : 0.0f, x0 = (p0 > ref) ? p0 : 0.0f, xp = (pp1 > ref) ? pp1 : 0.0f;
            const bool peak = in && (x0 > xm) && (x0 >= xp);
            const unsigned bal = __ballot_sync(FULL, peak);
            if (peak) {
                const float avg = (pp1 - pm1) * 0.5f;
                const float aa = (pp1 + pm1) - 2.0f * p0;
                const float shift = (fabsf(avg) >= fabsf(aa)) ? 0.0f : -avg / aa;
                const float pitch = (float(k) + shift) * a.binhz;
                const float mag = p0 + (0.5f * avg) * shift;
                const int slot = base + __popc(bal & ((1u << lane) - 1u));
                if (slot < a.cand_cap) slots[slot] = make_float2(pitch, mag);
            }
            base += __popc(bal);
        }
        if (lane == 0) a.cand_count[(size_t)b * a.T + t] = min(base, a.cand_cap);
    }
    // ---- chroma_stft: the clip's tuned filterbank applied to this frame's power spectrum, util.normalize(norm=inf)
    if (a.chroma_out != nullptr) {
        const float* fb = a.chroma_fb + (size_t)a.chroma_tidx[b] * kChroma * F;
        float mine = 0.0f;
        for (int c = 0; c < kChroma; ++c) {
            float acc = 0.0f;
            for (int k = lane; k < F; k += 32) acc = fmaf(__ldg(fb + (size_t)c * F + k), Pb[k], acc);
            acc = warp_sum(acc);
            if (lane == c) mine = acc;
        }
        const float mx = warp_max(fabsf(mine));        // lanes >= 12 hold 0
        const float inv = (mx < 1.17549435e-38f) ? 1.0f : 1.0f / mx;
        if (lane < kChroma) a.chroma_out[((size_t)b * kChroma + lane) * a.T + t] = mine * inv;
    }
    for (int k = lane; k < F; k += 32) {
        const float s = sqrtf(Pb[k]);
        s0 += s;
        s1 = fmaf(float(k), s, s1);
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    const float denom = (s0 < 1.17549435e-38f) ? 1.0f : s0;
    const float cen = s1 / denom;
    float q = 0.0f;
    for (int k = lane; k < F; k += 32) {
        const float d = float(k) - cen;
        q = fmaf(d * d, sqrtf(Pb[k]), q);
    }
    q = warp_sum(q);
    const float bw = sqrtf(q / denom);
    // rolloff: row-by-row inclusive scan of the magnitudes
    int rbin = F - 1;
    {
        const float thr = a.roll_percent * s0;
        float run = 0.0f;
        bool found = false;
        for (int k0 = 0; k0 < F && !found; k0 += 32) {
            const int k = k0 + lane;
            float v = (k < F) ? sqrtf(Pb[k]) : 0.0f;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float tv = __shfl_up_sync(FULL, v, d);
                if (lane >= d) v += tv;
            }
            const unsigned msk = __ballot_sync(FULL, (k < F) && (run + v >= thr));
            if (msk) { rbin = k0 + __ffs(msk) - 1; found = true; }
            run += __shfl_sync(FULL, v, 31);
        }
    }
    if (lane == 0) {
        if (a.stats != nullptr) {
            float* st = a.stats + (size_t)b * 5 * a.T + t;
            st[0] = cen * a.binhz;
            st[(size_t)a.T] = bw * a.binhz;
            st[(size_t)2 * a.T] = float(rbin) * a.binhz;
            st[(size_t)3 * a.T] = float(zc) / float(a.n_fft);
            st[(size_t)4 * a.T] = sqrtf(ss / float(a.n_fft));
        }
        if (a.status != nullptr && !(fabsf(ss) <= 3.0e38f)) atomicOr(a.status + b, 1);
    }
    if (a.mel_out != nullptr) {
        float wmax = 0.0f;
        for (int m = lane; m < a.n_mels; m += 32) {
            const int lo = gt.mel_lo[m], len = gt.mel_len[m];
            const float* w = gt.mel_w + gt.mel_off[m];
            float acc = 0.0f;
            for (int i = 0; i < len; ++i) {
                const float p = Pb[lo + i];
                acc = fmaf(w[i], a.use_mag ? sqrtf(p) : p, acc);
            }
            if (a.mel_frame_major) a.mel_out[((size_t)b * a.T + t) * a.n_mels + m] = acc;
            else a.mel_out[((size_t)b * a.n_mels + m) * a.T + t] = acc;
            wmax = fmaxf(wmax, acc);
        }
        wmax = warp_max(wmax);
        if (lane == 0 && a.clipmax != nullptr)
            atomicMax(reinterpret_cast<int*>(a.clipmax) + b, __float_as_int(wmax));
    }
}

cudaError_t launch_frames_generic(const FrameArgs& a, const GenericTables& gt, cudaStream_t stream) {
    const int M = a.n_fft / 2;
    int logM = 0;
    while ((1 << logM) < M) ++logM;
    const int smem = kGenWarps * (M * 8 + (M + 4) * 4);
    cudaError_t e = cudaFuncSetAttribute(frames_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const long long frames = (long long)a.B * a.T;
    if (frames <= 0) return cudaSuccess;
    const long long grid = (frames + kGenWarps - 1) / kGenWarps;
    frames_generic<<<(unsigned)grid, kGenWarps * 32, smem, stream>>>(a, gt, logM);
    g_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// power_to_db (per-clip ref=max / top_db) fused with the DCT-II of librosa.feature.mfcc.
// One thread per (clip, frame); DCT matrix broadcast from shared memory.
// ---------------------------------------------------------------------------
// 10*log10(x); __fmul_rn keeps the product from being contracted into the following
// subtraction, so db10(max) - db10(max) is exactly 0 (librosa: ref=np.max peaks at 0.0 dB)
__device__ __forceinline__ float db10(float x) { return __fmul_rn(3.0102999566398120f, __log2f(x)); }

constexpr int kDbThreads = 128;
// frames per thread: every shared-memory broadcast of a DCT row feeds 2x the FMAs (while the accumulators fit)
__host__ __device__ constexpr int db_frames(int nc) { return nc <= 40 ? 2 : 1; }

template <int NC>
__global__ void __launch_bounds__(kDbThreads, (NC <= 32) ? 4 : (NC <= 40) ? 3 : 1) db_dct(const DbArgs a) {
    // The DCT-II basis is symmetric (even k) / antisymmetric (odd k) about the middle of the mel axis:
    // cos(pi (2(N-1-n)+1) k / 2N) = (-1)^k cos(pi (2n+1) k / 2N).  Folding the input into
    // s_n = x_n + x_(N-1-n) and d_n = x_n - x_(N-1-n) halves the multiply-adds: even coefficients
    // come from s, odd ones from d.  sD[n][j]: j < NC/2 -> coefficient 2j, else coefficient 2(j-NC/2)+1.
    // The table is read as warp-wide 16-byte broadcasts, which cost 4 shared-memory cycles each; with
    // one frame per thread that pipe, not the FMA pipe, bounds the kernel - hence kDbFrames frames per thread.
    extern __shared__ __align__(16) float sD[];
    constexpr int H = NC / 2;
    constexpr int F = db_frames(NC);
    const int half = a.n_mels >> 1;                 // mirrored pairs; an odd n_mels has a middle row
    if constexpr (NC > 0) {
        const int rows = (a.n_mels + 1) >> 1;
        for (int i = threadIdx.x; i < rows * NC; i += kDbThreads) {
            const int n = i / NC, j = i - n * NC;
            const int c = (j < H) ? 2 * j : 2 * (j - H) + 1;
            sD[i] = a.dct_t[n * NC + c];
        }
        __syncthreads();
    }
    const long long total = (long long)a.B * a.T;
    // librosa.feature.mfcc always calls power_to_db(S) with ref=1.0, amin=1e-10, top_db=80
    const bool same_amin = (a.amin == 1e-10f);
    bool live[F];
    float ref_db[F], floor_db[F], floor_m[F];
    float* col[F];
    const float* row[F];
    size_t mfcc_off[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        // thread i of the block takes frames i and i + 128 of the block's 256: stores stay coalesced
        long long g = ((long long)blockIdx.x * F + f) * kDbThreads + threadIdx.x;
        live[f] = g < total;
        if (!live[f]) g = total - 1;                // compute on a valid frame, skip the stores
        const int b = (int)(g / a.T), t = (int)(g - (long long)b * a.T);
        const float pmax = __uint_as_float(a.clipmax[b]);
        const float maxa = db10(fmaxf(a.amin, pmax));
        ref_db[f] = (a.ref_mode == 1) ? maxa : db10(fmaxf(a.amin, fabsf(a.ref_value)));
        floor_db[f] = (a.top_db >= 0.0f) ? (maxa - ref_db[f]) - a.top_db : -CUDART_INF_F;
        floor_m[f] = db10(fmaxf(1e-10f, pmax)) - 80.0f;
        col[f] = a.mel + (size_t)b * a.n_mels * a.T + t;
        row[f] = a.mel_in ? a.mel_in + ((size_t)b * a.T + t) * a.n_mels : nullptr;
        mfcc_off[f] = (size_t)b * a.n_mfcc * a.T + t;
    }
    if (!live[0]) return;
    // packed FP32 pairs: acc[f][j] holds sD columns 2j and 2j + 1 (j < H/2: even coefficients, else odd ones)
    float2 acc[F][NC > 0 ? H : 1];
#pragma unroll
    for (int f = 0; f < F; ++f)
#pragma unroll
        for (int c = 0; c < H; ++c) acc[f][c] = make_float2(0.0f, 0.0f);
    const bool vec = (a.mel_in != nullptr) && ((a.n_mels & 7) == 0);
    // one mel row: dB for the log-mel output, the (ref = 1, top_db = 80) dB value for the DCT
    auto to_db = [&](int f, float p, int m) -> float {
        const float adb = db10(fmaxf(a.amin, p));
        if (live[f]) col[f][(size_t)m * a.T] = fmaxf(adb - ref_db[f], floor_db[f]);
        const float x = same_amin ? adb : db10(fmaxf(1e-10f, p));
        return fmaxf(x, floor_m[f]);
    };
    auto fetch = [&](int f, int m) -> float { return row[f] ? row[f][m] : col[f][(size_t)m * a.T]; };
    constexpr int MB = 4;                           // mirrored pairs per batch: 8 values per frame in flight
    for (int n0 = 0; n0 < half; n0 += MB) {
        float lo[F][MB], hi[F][MB];                 // lo[j] = row n0 + j, hi[j] = row N-1-(n0+j)
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (vec) {                              // frame-major scratch: two 16-byte loads
                const float4 u = __ldg(reinterpret_cast<const float4*>(row[f] + n0));
                const float4 v = __ldg(reinterpret_cast<const float4*>(row[f] + a.n_mels - MB - n0));
                lo[f][0] = u.x; lo[f][1] = u.y; lo[f][2] = u.z; lo[f][3] = u.w;
                hi[f][0] = v.w; hi[f][1] = v.z; hi[f][2] = v.y; hi[f][3] = v.x;
            } else {
#pragma unroll
                for (int j = 0; j < MB; ++j) {
                    const bool in = n0 + j < half;
                    lo[f][j] = in ? fetch(f, n0 + j) : 0.0f;
                    hi[f][j] = in ? fetch(f, a.n_mels - 1 - n0 - j) : 0.0f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < MB; ++j) {
            const int n = n0 + j;
            if (n >= half) break;
            float s[F], d[F];
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float xl = to_db(f, lo[f][j], n), xh = to_db(f, hi[f][j], a.n_mels - 1 - n);
                s[f] = xl + xh; d[f] = xl - xh;
            }
            if constexpr (NC > 0) {
                const float4* d4 = reinterpret_cast<const float4*>(sD + n * NC);
#pragma unroll
                for (int c = 0; c < H / 4; ++c) {
                    const float4 e = d4[c], o = d4[H / 4 + c];
#pragma unroll
                    for (int f = 0; f < F; ++f) {       // packed FP32: two coefficients per FFMA2
                        const float2 sf = make_float2(s[f], s[f]), df = make_float2(d[f], d[f]);
                        acc[f][2 * c] = __ffma2_rn(make_float2(e.x, e.y), sf, acc[f][2 * c]);
                        acc[f][2 * c + 1] = __ffma2_rn(make_float2(e.z, e.w), sf, acc[f][2 * c + 1]);
                        acc[f][H / 2 + 2 * c] = __ffma2_rn(make_float2(o.x, o.y), df, acc[f][H / 2 + 2 * c]);
                        acc[f][H / 2 + 2 * c + 1] = __ffma2_rn(make_float2(o.z, o.w), df, acc[f][H / 2 + 2 * c + 1]);
                    }
                }
            }
        }
    }
    if (a.n_mels & 1) {                             // middle row: odd coefficients vanish there
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const float x = to_db(f, fetch(f, half), half);
            if constexpr (NC > 0) {
#pragma unroll
                for (int c = 0; c < H / 2; ++c)
                    acc[f][c] = __ffma2_rn(make_float2(sD[half * NC + 2 * c], sD[half * NC + 2 * c + 1]),
                                           make_float2(x, x), acc[f][c]);
            }
        }
    }
    if constexpr (NC > 0) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (!live[f]) continue;
            float* mo = a.mfcc + mfcc_off[f];
#pragma unroll
            for (int j = 0; j < H; ++j) {       // sD column j = coefficient 2j, column H + j = coefficient 2j + 1
                const float ev = (j & 1) ? acc[f][j >> 1].y : acc[f][j >> 1].x;
                const float od = (j & 1) ? acc[f][H / 2 + (j >> 1)].y : acc[f][H / 2 + (j >> 1)].x;
                if (2 * j < a.n_mfcc) mo[(size_t)(2 * j) * a.T] = ev;
                if (2 * j + 1 < a.n_mfcc) mo[(size_t)(2 * j + 1) * a.T] = od;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// db_dct_small: the same power_to_db + DCT-II for SMALL batches (the reference's per-file loop: one clip = 130 frames).
// db_dct gives a thread two whole frames - 128 mel bands, 16 dependent load -> dB -> 320-FMA rounds - so a single clip
// is one CTA running a 20 us serial chain (longer than the frames kernel).  Here four lanes share a frame: per round
// of 4 mirrored band pairs lane q loads and converts pair q (and writes its two log-mel values), the four (s, d)
// pairs meet in shuffles, and every lane accumulates its own quarter of the DCT coefficients over ALL bands - in
// db_dct's order, from db_dct's folded table, so log-mel AND MFCC are bitwise what db_dct writes whatever the batch
// size.  A warp carries 8 frames, a CTA 32: a clip spreads over 5 CTAs of short chains.  Stores are 32-byte pieces
// (8 frames of one band), which is why large batches stay on db_dct.
// ---------------------------------------------------------------------------
constexpr int kDbsThreads = 128;

template <int NC>
__global__ void __launch_bounds__(kDbsThreads) db_dct_small(const DbArgs a) {
    extern __shared__ __align__(16) float sD[];         // db_dct's folded table: sD[n][j], j < H: coefficient 2j, else 2(j-H)+1
    constexpr int H = NC / 2, HQ = H / 4;                // NC is a multiple of 8: every lane owns HQ even and HQ odd coefficients
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = lane >> 2, q = lane & 3, fbase = lane & ~3;
    const int half = a.n_mels >> 1;
    {
        const int rows = (a.n_mels + 1) >> 1;
        for (int i = tid; i < rows * NC; i += kDbsThreads) {
            const int n = i / NC, j = i - n * NC;
            const int c = (j < H) ? 2 * j : 2 * (j - H) + 1;
            sD[i] = a.dct_t[n * NC + c];
        }
        __syncthreads();
    }
    const long long total = (long long)a.B * a.T;
    long long G = ((long long)blockIdx.x * (kDbsThreads / 32) + warp) * 8 + f;
    const bool live = G < total;
    if (!live) G = total - 1;
    const int b = (int)(G / a.T), t = (int)(G - (long long)b * a.T);
    const float pmax = __uint_as_float(a.clipmax[b]);
    const float maxa = db10(fmaxf(a.amin, pmax));
    const float ref_db = (a.ref_mode == 1) ? maxa : db10(fmaxf(a.amin, fabsf(a.ref_value)));
    const float floor_db = (a.top_db >= 0.0f) ? (maxa - ref_db) - a.top_db : -CUDART_INF_F;
    const float floor_m = db10(fmaxf(1e-10f, pmax)) - 80.0f;   // librosa.feature.mfcc: ref = 1, amin = 1e-10, top_db = 80
    const bool same_amin = (a.amin == 1e-10f);
    float* col = a.mel + (size_t)b * a.n_mels * a.T + t;
    const float* row = a.mel_in ? a.mel_in + ((size_t)b * a.T + t) * a.n_mels : nullptr;
    auto fetch = [&](int m) -> float { return row ? __ldg(row + m) : col[(size_t)m * a.T]; };
    auto to_db = [&](float p, int m) -> float {
        const float adb = db10(fmaxf(a.amin, p));
        if (live) col[(size_t)m * a.T] = fmaxf(adb - ref_db, floor_db);
        const float x = same_amin ? adb : db10(fmaxf(1e-10f, p));
        return fmaxf(x, floor_m);
    };
    float ae[HQ], ao[HQ];                               // coefficients 2 (HQ q + i) and 2 (HQ q + i) + 1
#pragma unroll
    for (int i = 0; i < HQ; ++i) { ae[i] = 0.0f; ao[i] = 0.0f; }
    for (int n0 = 0; n0 < half; n0 += 4) {
        const int n = n0 + q;
        float s = 0.0f, d = 0.0f;
        if (n < half) {                                 // this lane's mirrored pair of the round
            const float xl = to_db(fetch(n), n), xh = to_db(fetch(a.n_mels - 1 - n), a.n_mels - 1 - n);
            s = xl + xh; d = xl - xh;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {                   // all four pairs, in db_dct's order
            const float sj = __shfl_sync(FULL, s, fbase + j), dj = __shfl_sync(FULL, d, fbase + j);
            if (n0 + j < half) {
                const float* dr = sD + (n0 + j) * NC + HQ * q;
#pragma unroll
                for (int i = 0; i < HQ; ++i) {
                    ae[i] = fmaf(dr[i], sj, ae[i]);
                    ao[i] = fmaf(dr[H + i], dj, ao[i]);
                }
            }
        }
    }
    if (a.n_mels & 1) {                                 // middle row: odd coefficients vanish there
        float x = 0.0f;
        if (q == 0) x = to_db(fetch(half), half);
        x = __shfl_sync(FULL, x, fbase);
#pragma unroll
        for (int i = 0; i < HQ; ++i) ae[i] = fmaf(sD[half * NC + HQ * q + i], x, ae[i]);
    }
    if (live) {
        float* mo = a.mfcc + (size_t)b * a.n_mfcc * a.T + t;
#pragma unroll
        for (int i = 0; i < HQ; ++i) {
            const int ce = 2 * (HQ * q + i);
            if (ce < a.n_mfcc) mo[(size_t)ce * a.T] = ae[i];
            if (ce + 1 < a.n_mfcc) mo[(size_t)(ce + 1) * a.T] = ao[i];
        }
    }
}

template <int NC>
static cudaError_t launch_db_small_nc(const DbArgs& a, cudaStream_t stream) {
    const long long frames = (long long)a.B * a.T;
    if (frames <= 0) return cudaSuccess;
    const int smem = ((a.n_mels + 1) / 2) * NC * 4;   // folded DCT table
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(db_dct_small<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    const long long per_block = (kDbsThreads / 32) * 8;
    db_dct_small<NC><<<(unsigned)((frames + per_block - 1) / per_block), kDbsThreads, smem, stream>>>(a);
    g_launches++;
    return cudaGetLastError();
}

template <int NC>
static cudaError_t launch_db_nc(const DbArgs& a, cudaStream_t stream) {
    const long long frames = (long long)a.B * a.T;
    if (frames <= 0) return cudaSuccess;
    const int smem = ((a.n_mels + 1) / 2) * NC * 4;   // folded DCT table
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(db_dct<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    const long long per_block = (long long)kDbThreads * db_frames(NC);
    db_dct<NC><<<(unsigned)((frames + per_block - 1) / per_block), kDbThreads, smem, stream>>>(a);
    g_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// power_to_db + DCT-II + time pooling in one kernel: what 1_preprocessing.py keeps of a clip
// is np.mean / np.std over frames of every log-mel band, MFCC and statistic ([R] src/1_preprocessing.py:
// 115-124), so when the caller does not ask for the (B, n_mels, T) / (B, n_mfcc, T) arrays they are never
// written to HBM.  One CTA of 128 threads per clip, two phases over the clip's frame-major mel power:
//   A. one thread per mel band walks the frames (every load a coalesced n_mels-float row), converts to dB and
//      pools in chunks of 16 frames: chunk mean and squared deviations from registers, chunks merged with
//      Chan's update (a constant row has exactly zero variance, no cancellation far from zero);
//   B. one thread per frame runs db_dct's folded DCT (its second read of the row hits L1 / L2) into a small
//      shared-memory tile of coefficients, which one thread per coefficient pools the same way.
// The 5 statistics rows and the 24 chroma rows are pooled from HBM with pool_kernel's two-pass arithmetic.
// 31 KB of shared memory per CTA: 7 CTAs per SM overlap each other's phases.
// ---------------------------------------------------------------------------
// Threads per CTA = frames per phase-B round, chosen at launch so that the clip's T frames fill the rounds
// evenly (T = 130 -> one round of 160; T = 1292 -> six rounds of 224), between 128 and 256.
constexpr int kPoolMinThreads = 128, kPoolMaxThreads = 256;
constexpr int kPoolRowsPerThread = kMaxMelGroups * 32 / kPoolMinThreads;   // n_mels <= 256
constexpr int kPoolChunk = 16;

static int pool_threads(int T) {
    const int rounds = (T + kPoolMaxThreads - 1) / kPoolMaxThreads;
    int nt = (((T + rounds - 1) / rounds) + 31) & ~31;
    return nt < kPoolMinThreads ? kPoolMinThreads : nt;
}
static int pool_smem_bytes(int n_mels, int nc, int nt) {
    const int rows = (n_mels + 1) / 2;
    return (((rows * nc + 3) & ~3) + nt * (nc + 1)) * 4;
}
bool db_pool_fits(int n_mels, int ncp, int T) { return pool_smem_bytes(n_mels, ncp, pool_threads(T > 0 ? T : 1)) <= 224 * 1024; }

// Chan's update of (mean, M2) over n_a samples with a chunk of n_b samples given by its (mean, M2)
__device__ __forceinline__ void chan_merge(float& mean, float& m2, int na, float cmean, float cm2, int nb) {
    if (na == 0) { mean = cmean; m2 = cm2; return; }
    const float fa = float(na), fb = float(nb), delta = cmean - mean;
    mean = fmaf(delta, fb / (fa + fb), mean);
    m2 += cm2 + delta * delta * (fa * fb / (fa + fb));
}

template <int NC>
__global__ void __launch_bounds__(kPoolMaxThreads) db_pool(const DbArgs a, const PoolArgs pa) {
    extern __shared__ __align__(16) float sP[];
    constexpr int H = NC / 2;
    const int NT = blockDim.x;                      // = frames per phase-B round
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = a.n_mels >> 1;
    const int rows = (a.n_mels + 1) >> 1;
    constexpr int MS = NC + 1;
    float* sD = sP;
    float* tileM = sD + ((rows * NC + 3) & ~3);
    if constexpr (NC > 0) {
        for (int i = tid; i < rows * NC; i += NT) {
            const int n = i / NC, j = i - n * NC;
            const int c = (j < H) ? 2 * j : 2 * (j - H) + 1;
            sD[i] = a.dct_t[n * NC + c];
        }
    }
    __syncthreads();
    const bool same_amin = (a.amin == 1e-10f);
    const bool vec = (a.n_mels & 7) == 0;
    const float invT = 1.0f / float(a.T);
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const float pmax = __uint_as_float(a.clipmax[b]);
        const float maxa = db10(fmaxf(a.amin, pmax));
        const float ref_db = (a.ref_mode == 1) ? maxa : db10(fmaxf(a.amin, fabsf(a.ref_value)));
        const float floor_db = (a.top_db >= 0.0f) ? (maxa - ref_db) - a.top_db : -CUDART_INF_F;
        const float floor_m = db10(fmaxf(1e-10f, pmax)) - 80.0f;
        const float* clip_mel = a.mel_in + (size_t)b * a.T * a.n_mels;
        float* out = pa.pooled + (size_t)b * pa.pooled_w;

        // ---- phase A: log-mel bands.  Thread m reads element m of every frame's row.
#pragma unroll
        for (int r = 0; r < kPoolRowsPerThread; ++r) {
            const int m = tid + r * NT;
            if (m >= a.n_mels) continue;
            float mean = 0.0f, m2 = 0.0f;
            for (int f0 = 0; f0 < a.T; f0 += kPoolChunk) {
                const int nb = (a.T - f0 < kPoolChunk) ? a.T - f0 : kPoolChunk;
                float x[kPoolChunk];
#pragma unroll
                for (int u = 0; u < kPoolChunk; ++u)          // independent loads first
                    x[u] = (u < nb) ? __ldg(clip_mel + (size_t)(f0 + u) * a.n_mels + m) : 0.0f;
                float sum = 0.0f;
#pragma unroll
                for (int u = 0; u < kPoolChunk; ++u) {
                    x[u] = fmaxf(db10(fmaxf(a.amin, x[u])) - ref_db, floor_db);
                    sum += (u < nb) ? x[u] : 0.0f;
                }
                const float cmean = sum / float(nb);
                float q = 0.0f;
#pragma unroll
                for (int u = 0; u < kPoolChunk; ++u) { const float d = (u < nb) ? x[u] - cmean : 0.0f; q = fmaf(d, d, q); }
                chan_merge(mean, m2, f0, cmean, q, nb);
            }
            out[m] = mean;
            out[a.n_mels + m] = sqrtf(m2 * invT);
        }

        // ---- phase B: MFCC.  One thread per frame, tiles of NT frames, then one thread per coefficient.
        if constexpr (NC > 0) {
            if (pa.with_mfcc) {
                float mean_c = 0.0f, m2_c = 0.0f;
                for (int t0 = 0; t0 < a.T; t0 += NT) {
                    const int t = t0 + tid;
                    if (t < a.T) {
                        const float* row = clip_mel + (size_t)t * a.n_mels;
                        float2 acc[H];
#pragma unroll
                        for (int c = 0; c < H; ++c) acc[c] = make_float2(0.f, 0.f);
                        auto mdb = [&](float p) -> float {      // librosa.feature.mfcc: power_to_db(S), ref = 1, top_db = 80
                            const float x = same_amin ? db10(fmaxf(a.amin, p)) : db10(fmaxf(1e-10f, p));
                            return fmaxf(x, floor_m);
                        };
                        constexpr int MB = 4;
                        for (int n0 = 0; n0 < half; n0 += MB) {
                            float lo[MB], hi[MB];
                            if (vec) {
                                const float4 u = __ldg(reinterpret_cast<const float4*>(row + n0));
                                const float4 v = __ldg(reinterpret_cast<const float4*>(row + a.n_mels - MB - n0));
                                lo[0] = u.x; lo[1] = u.y; lo[2] = u.z; lo[3] = u.w;
                                hi[0] = v.w; hi[1] = v.z; hi[2] = v.y; hi[3] = v.x;
                            } else {
#pragma unroll
                                for (int j = 0; j < MB; ++j) {
                                    const bool in = n0 + j < half;
                                    lo[j] = in ? row[n0 + j] : 0.0f;
                                    hi[j] = in ? row[a.n_mels - 1 - n0 - j] : 0.0f;
                                }
                            }
#pragma unroll
                            for (int j = 0; j < MB; ++j) {
                                const int n = n0 + j;
                                if (n >= half) break;
                                const float xl = mdb(lo[j]), xh = mdb(hi[j]);
                                const float2 sf = make_float2(xl + xh, xl + xh), df = make_float2(xl - xh, xl - xh);
                                const float4* d4 = reinterpret_cast<const float4*>(sD + n * NC);
#pragma unroll
                                for (int c = 0; c < H / 4; ++c) {
                                    const float4 e = d4[c], o = d4[H / 4 + c];
                                    acc[2 * c] = __ffma2_rn(make_float2(e.x, e.y), sf, acc[2 * c]);
                                    acc[2 * c + 1] = __ffma2_rn(make_float2(e.z, e.w), sf, acc[2 * c + 1]);
                                    acc[H / 2 + 2 * c] = __ffma2_rn(make_float2(o.x, o.y), df, acc[H / 2 + 2 * c]);
                                    acc[H / 2 + 2 * c + 1] = __ffma2_rn(make_float2(o.z, o.w), df, acc[H / 2 + 2 * c + 1]);
                                }
                            }
                        }
                        if (a.n_mels & 1) {
                            const float x = mdb(row[half]);
#pragma unroll
                            for (int c = 0; c < H / 2; ++c)
                                acc[c] = __ffma2_rn(make_float2(sD[half * NC + 2 * c], sD[half * NC + 2 * c + 1]),
                                                    make_float2(x, x), acc[c]);
                        }
                        float* mrow = tileM + tid * MS;
#pragma unroll
                        for (int j = 0; j < H; ++j) {
                            mrow[2 * j] = (j & 1) ? acc[j >> 1].y : acc[j >> 1].x;
                            mrow[2 * j + 1] = (j & 1) ? acc[H / 2 + (j >> 1)].y : acc[H / 2 + (j >> 1)].x;
                        }
                    }
                    __syncthreads();
                    const int nt = (a.T - t0 < NT) ? a.T - t0 : NT;
                    if (tid < a.n_mfcc) {
                        const float* col = tileM + tid;
                        float sum = 0.0f;
                        for (int f = 0; f < nt; ++f) sum += col[f * MS];
                        const float cmean = sum / float(nt);
                        float q = 0.0f;
                        for (int f = 0; f < nt; ++f) { const float d = col[f * MS] - cmean; q = fmaf(d, d, q); }
                        chan_merge(mean_c, m2_c, t0, cmean, q, nt);
                    }
                    __syncthreads();
                }
                if (tid < a.n_mfcc) {
                    out[2 * a.n_mels + tid] = mean_c;
                    out[2 * a.n_mels + a.n_mfcc + tid] = sqrtf(m2_c * invT);
                }
            }
        }

        // ---- statistics and chroma rows: warp per row, two passes over HBM (pool_kernel's arithmetic)
        const int base = 2 * a.n_mels + 2 * ((NC > 0 && pa.with_mfcc) ? a.n_mfcc : 0);
        for (int r = warp; r < 5 + pa.n_chroma; r += NT / 32) {
            const float* src = (r < 5) ? pa.stats + ((size_t)b * 5 + r) * a.T
                                       : pa.chroma + ((size_t)b * pa.n_chroma + (r - 5)) * a.T;
            float sum = 0.0f;
            for (int i = lane; i < a.T; i += 32) sum += src[i];
            sum = warp_sum(sum);
            const float mean = sum / float(a.T);
            float var = 0.0f;
            for (int i = lane; i < a.T; i += 32) { const float d = src[i] - mean; var = fmaf(d, d, var); }
            var = warp_sum(var);
            if (lane == 0) {
                const int o_mean = (r < 5) ? base + 2 * r : base + 10 + (r - 5);
                const int o_std = (r < 5) ? o_mean + 1 : o_mean + pa.n_chroma;
                out[o_mean] = mean;
                out[o_std] = sqrtf(var / float(a.T));
            }
        }
    }
}

template <int NC>
static cudaError_t launch_db_pool_nc(const DbArgs& a, const PoolArgs& pa, int num_sms, cudaStream_t stream) {
    if (a.B <= 0 || a.T <= 0) return cudaSuccess;
    const int nt = pool_threads(a.T);
    const int smem = pool_smem_bytes(a.n_mels, NC, nt);
    if (smem > 224 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(db_pool<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int per_sm = (224 * 1024) / (smem + 1024);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid = (long long)num_sms * per_sm;
    if (grid > a.B) grid = a.B;
    db_pool<NC><<<(unsigned)grid, nt, smem, stream>>>(a, pa);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_db_pool(const DbArgs& a, const PoolArgs& pa, int num_sms, cudaStream_t stream) {
    if (!pa.with_mfcc || a.n_mfcc <= 0) return launch_db_pool_nc<0>(a, pa, num_sms, stream);
    switch (a.ncp) {
        case 8: return launch_db_pool_nc<8>(a, pa, num_sms, stream);
        case 16: return launch_db_pool_nc<16>(a, pa, num_sms, stream);
        case 24: return launch_db_pool_nc<24>(a, pa, num_sms, stream);
        case 32: return launch_db_pool_nc<32>(a, pa, num_sms, stream);
        case 40: return launch_db_pool_nc<40>(a, pa, num_sms, stream);
        case 64: return launch_db_pool_nc<64>(a, pa, num_sms, stream);
        case 128: return launch_db_pool_nc<128>(a, pa, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

// frames below which the four-lanes-per-frame kernel is faster (measured crossover, tools/gpu_small_batch.py)
constexpr long long kDbSmallFrames = 24000;

cudaError_t launch_db_dct(const DbArgs& a, cudaStream_t stream) {
    static const long long small_frames = [] { const char* e = getenv("HLMC_DB_SMALL"); return e ? atoll(e) : kDbSmallFrames; }();
    if ((long long)a.B * a.T <= small_frames && a.mfcc != nullptr && a.n_mfcc > 0) {
        switch (a.ncp) {
            case 8: return launch_db_small_nc<8>(a, stream);
            case 16: return launch_db_small_nc<16>(a, stream);
            case 24: return launch_db_small_nc<24>(a, stream);
            case 32: return launch_db_small_nc<32>(a, stream);
            case 40: return launch_db_small_nc<40>(a, stream);
            default: break;
        }
    }
    if (a.mfcc == nullptr || a.n_mfcc <= 0) return launch_db_nc<0>(a, stream);
    switch (a.ncp) {
        case 8: return launch_db_nc<8>(a, stream);
        case 16: return launch_db_nc<16>(a, stream);
        case 24: return launch_db_nc<24>(a, stream);
        case 32: return launch_db_nc<32>(a, stream);
        case 40: return launch_db_nc<40>(a, stream);
        case 64: return launch_db_nc<64>(a, stream);
        case 128: return launch_db_nc<128>(a, stream);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------
// Stand-alone librosa.power_to_db on (B, per_clip) arrays.
// ---------------------------------------------------------------------------
__global__ void rowmax_kernel(const float* __restrict__ in, unsigned int* clipmax, long long per_clip) {
    const long long b = blockIdx.y;
    const float* p = in + b * per_clip;
    float m = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_clip;
         i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(p[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(clipmax) + b, __float_as_int(m));
}
cudaError_t launch_rowmax(const float* in, unsigned int* clipmax, long long B, long long per_clip,
                          cudaStream_t stream) {
    if (B <= 0 || per_clip <= 0) return cudaSuccess;
    int gx = (int)((per_clip + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    for (long long b0 = 0; b0 < B; b0 += kMaxGridY) {         // gridDim.y is capped at 65535
        const long long nb = (B - b0 < kMaxGridY) ? B - b0 : kMaxGridY;
        rowmax_kernel<<<dim3(gx, (unsigned)nb), 256, 0, stream>>>(in + b0 * per_clip, clipmax + b0, per_clip);
        g_launches++;
    }
    return cudaGetLastError();
}
__global__ void power_to_db_kernel(const float* __restrict__ in, float* __restrict__ out,
                                   const unsigned int* __restrict__ clipmax, long long per_clip,
                                   int ref_mode, float ref_value, float amin, float top_db) {
    const long long b = blockIdx.y;
    const float pmax = __uint_as_float(clipmax[b]);
    const float maxa = db10(fmaxf(amin, pmax));
    const float ref_db = (ref_mode == 1) ? maxa : db10(fmaxf(amin, fabsf(ref_value)));
    const float floor_db = (top_db >= 0.0f) ? (maxa - ref_db) - top_db : -CUDART_INF_F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_clip;
         i += (long long)gridDim.x * blockDim.x) {
        const float v = db10(fmaxf(amin, fabsf(in[b * per_clip + i]))) - ref_db;
        out[b * per_clip + i] = fmaxf(v, floor_db);
    }
}
cudaError_t launch_power_to_db(const float* in, float* out, const unsigned int* clipmax, long long B,
                               long long per_clip, int ref_mode, float ref_value, float amin,
                               float top_db, cudaStream_t stream) {
    if (B <= 0 || per_clip <= 0) return cudaSuccess;
    int gx = (int)((per_clip + 256 * 4 - 1) / (256 * 4));
    if (gx < 1) gx = 1;
    if (gx > 256) gx = 256;
    for (long long b0 = 0; b0 < B; b0 += kMaxGridY) {
        const long long nb = (B - b0 < kMaxGridY) ? B - b0 : kMaxGridY;
        power_to_db_kernel<<<dim3(gx, (unsigned)nb), 256, 0, stream>>>(in + b0 * per_clip, out + b0 * per_clip,
                                                                       clipmax + b0, per_clip, ref_mode, ref_value,
                                                                       amin, top_db);
        g_launches++;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Time pooling: np.mean / np.std (ddof=0) over frames, one warp per feature row
// ([R] src/1_preprocessing.py:115-124, src/1_preprocessing_advanced.py:144-151).
// ---------------------------------------------------------------------------
__global__ void pool_kernel(const float* __restrict__ logmel, const float* __restrict__ mfcc,
                            const float* __restrict__ stats, const float* __restrict__ chroma, long long B,
                            int n_mels, int n_mfcc, int n_chroma, int T, float* __restrict__ pooled) {
    const int rows = n_mels + n_mfcc + 5 + n_chroma;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= B * rows) return;
    const long long b = wid / rows;
    const int r = (int)(wid - b * rows);
    const float* src;
    int o_mean, o_std;
    if (r < n_mels) {
        src = logmel + ((size_t)b * n_mels + r) * T;
        o_mean = r; o_std = n_mels + r;
    } else if (r < n_mels + n_mfcc) {
        const int c = r - n_mels;
        src = mfcc + ((size_t)b * n_mfcc + c) * T;
        o_mean = 2 * n_mels + c; o_std = 2 * n_mels + n_mfcc + c;
    } else if (r < n_mels + n_mfcc + 5) {
        const int s = r - n_mels - n_mfcc;
        src = stats + ((size_t)b * 5 + s) * T;
        o_mean = 2 * n_mels + 2 * n_mfcc + 2 * s; o_std = o_mean + 1;
    } else {
        const int c = r - n_mels - n_mfcc - 5;
        src = chroma + ((size_t)b * n_chroma + c) * T;
        o_mean = 2 * n_mels + 2 * n_mfcc + 10 + c; o_std = o_mean + n_chroma;
    }
    float sum = 0.0f;
    for (int i = lane; i < T; i += 32) sum += src[i];
    sum = warp_sum(sum);
    const float mean = sum / float(T);
    float var = 0.0f;
    for (int i = lane; i < T; i += 32) { const float d = src[i] - mean; var = fmaf(d, d, var); }
    var = warp_sum(var);
    if (lane == 0) {
        float* out = pooled + (size_t)b * (2 * n_mels + 2 * n_mfcc + 10 + 2 * n_chroma);
        out[o_mean] = mean;
        out[o_std] = sqrtf(var / float(T));
    }
}
cudaError_t launch_pool(const float* logmel, const float* mfcc, const float* stats, const float* chroma,
                        long long B, int n_mels, int n_mfcc, int T, float* pooled, cudaStream_t stream) {
    if (mfcc == nullptr) n_mfcc = 0;
    const int n_chroma = chroma ? kChroma : 0;
    const long long warps = B * (n_mels + n_mfcc + 5 + n_chroma);
    if (warps <= 0) return cudaSuccess;
    pool_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(logmel, mfcc, stats, chroma, B, n_mels,
                                                                 n_mfcc, n_chroma, T, pooled);
    g_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// [R] src/1_preprocessing_advanced.py:108-112: crop to `fixed` frames, or right-pad
// with the clip's minimum.  One CTA per clip.
// ---------------------------------------------------------------------------
__global__ void fix_frames_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int T,
                                  int fixed) {
    __shared__ float s_min[32];
    const long long b = blockIdx.x;
    const float* src = in + (size_t)b * rows * T;
    float* dst = out + (size_t)b * rows * fixed;
    float fill = 0.0f;
    if (T < fixed) {
        float m = CUDART_INF_F;
        for (int i = threadIdx.x; i < rows * T; i += blockDim.x) m = fminf(m, src[i]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, d));
        if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = m;
        __syncthreads();
        m = CUDART_INF_F;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fminf(m, s_min[i]);
        fill = m;
    }
    const int keep = min(T, fixed);
    for (int i = threadIdx.x; i < rows * fixed; i += blockDim.x) {
        const int r = i / fixed, c = i - r * fixed;
        dst[i] = (c < keep) ? src[(size_t)r * T + c] : fill;
    }
}
cudaError_t launch_fix_frames(const float* in, float* out, long long B, int rows, int T, int fixed,
                              cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    fix_frames_kernel<<<(unsigned)B, 256, 0, stream>>>(in, out, rows, T, fixed);
    g_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// sklearn.preprocessing.StandardScaler on device ([R] src/1_preprocessing_advanced.py:376-382:
// the (N, 131072) flattened mel images).  Column statistics in float64, two passes like
// sklearn's _incremental_mean_and_var; one thread per column, rows walked with coalesced reads.
// ---------------------------------------------------------------------------
__global__ void colstats_kernel(const float* __restrict__ x, long long N, long long D,
                                double* __restrict__ mean, double* __restrict__ m2) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    double s = 0.0;
    for (long long r = 0; r < N; ++r) s += (double)x[r * D + c];
    const double mu = (N > 0) ? s / (double)N : 0.0;
    double d1 = 0.0, d2 = 0.0;
    for (long long r = 0; r < N; ++r) {
        const double d = (double)x[r * D + c] - mu;
        d1 += d;
        d2 = fma(d, d, d2);
    }
    mean[c] = mu;
    m2[c] = (N > 0) ? d2 - d1 * d1 / (double)N : 0.0;      // sum of squared deviations (corrected)
}
cudaError_t launch_colstats(const float* x, long long N, long long D, double* mean, double* m2,
                            cudaStream_t stream) {
    if (D <= 0) return cudaSuccess;
    colstats_kernel<<<(unsigned)((D + 127) / 128), 128, 0, stream>>>(x, N, D, mean, m2);
    g_launches++;
    return cudaGetLastError();
}
__global__ void standardize_kernel(const float* __restrict__ x, float* __restrict__ y, long long total,
                                   long long D, const float* __restrict__ mean, const float* __restrict__ scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long c = i % D;
        y[i] = (x[i] - mean[c]) / scale[c];                 // sklearn: X -= mean_; X /= scale_ in X's dtype
    }
}
cudaError_t launch_standardize(const float* x, float* y, long long N, long long D, const float* mean,
                               const float* scale, cudaStream_t stream) {
    const long long total = N * D;
    if (total <= 0) return cudaSuccess;
    long long grid = (total + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    standardize_kernel<<<(unsigned)grid, 256, 0, stream>>>(x, y, total, D, mean, scale);
    g_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Front end of [R] load_audio_file for PCM16 sources: float32 = int16 / 32768 (what
// librosa.load -> soundfile returns) and the scripts' right zero-pad to `n_total` samples,
// done on the device so only the valid int16 samples cross PCIe.
// ---------------------------------------------------------------------------
__global__ void pcm16_to_f32_kernel(const int16_t* __restrict__ raw, long long raw_pitch,
                                    float* __restrict__ out, long long pitch, long long n_valid,
                                    long long n_total) {
    const long long b = blockIdx.y;
    const int16_t* src = raw + b * raw_pitch;
    float* dst = out + b * pitch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_total;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = (i < n_valid) ? float(src[i]) * (1.0f / 32768.0f) : 0.0f;
}
cudaError_t launch_pcm16_to_f32(const int16_t* raw, long long raw_pitch, float* out, long long pitch,
                                long long B, long long n_valid, long long n_total, cudaStream_t stream) {
    if (B <= 0 || n_total <= 0) return cudaSuccess;
    int gx = (int)((n_total + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    if (gx > 128) gx = 128;
    for (long long b0 = 0; b0 < B; b0 += kMaxGridY) {
        const long long nb = (B - b0 < kMaxGridY) ? B - b0 : kMaxGridY;
        pcm16_to_f32_kernel<<<dim3(gx, (unsigned)nb), 256, 0, stream>>>(raw + b0 * raw_pitch, raw_pitch,
                                                                        out + b0 * pitch, pitch, n_valid, n_total);
        g_launches++;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// FP32 roofline denominator: independent FMA chains, every SM sub-partition busy.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float seed) {
    float x0 = seed + threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
    float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    const float a = 0.999f, c = 0.001f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, c); x1 = fmaf(x1, a, c); x2 = fmaf(x2, a, c); x3 = fmaf(x3, a, c);
            x4 = fmaf(x4, a, c); x5 = fmaf(x5, a, c); x6 = fmaf(x6, a, c); x7 = fmaf(x7, a, c);
        }
    }
    const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456f) out[0] = r;
}
cudaError_t measure_fp32_peak(double* tflops) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, grid = sms * 8;
    fma_peak_kernel<<<grid, 256>>>(d, 64, 1.0f);          // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<grid, 256>>>(d, iters, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * 16.0 * double(iters) * 256.0 * double(grid);
        const double tf = flops / (double(ms) * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    g_launches += 6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return cudaGetLastError();
}

}  // namespace hlmc
