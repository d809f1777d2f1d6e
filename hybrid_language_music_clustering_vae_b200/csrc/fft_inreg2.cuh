// In-register complex FFTs on Blackwell's packed FP32 pipe (FADD2 / FMUL2 / FFMA2, sm_100a).
//
// Same decimation-in-frequency radix-4 / radix-2 structure as fft_inreg.cuh, but every complex
// value is ONE float2 = one 64-bit register pair (re, im), so
//   * a complex add / subtract is one FADD2,
//   * a multiplication by -i / +i is free: ptxas folds the (im, -re) swap and the sign into the
//     operand modifiers of the consuming FADD2 (R.F32x2.LO_HI.NP),
//   * a twiddle multiplication is FMUL2 + FFMA2 with the cosine / sine as 32-bit immediates
//     broadcast to both halves,
// i.e. half the issue slots of the scalar form for the same FP32-lane work (measured on B200:
// FFMA2 issues at 0.5 / clk / SM sub-partition, FFMA at 1.0 - tools/fp32x2_bench.cu).
//
// Replaces: scipy.fft.rfft inside librosa.stft; the reference reaches it
// from every feature call of src/1_preprocessing.py:50-83.
#pragma once
#include "fft_inreg.cuh"

namespace fftreg2 {

using fftreg::cos2pi;
using fftreg::fft_pos;
using fftreg::sin2pi;

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a + (-i) b
__device__ __forceinline__ float2 cadd_mi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }
// a + (+i) b
__device__ __forceinline__ float2 cadd_pi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }
// v * (c - i s) = (x c + y s, y c - x s)
__device__ __forceinline__ float2 cmul_conj(float2 v, float c, float s) {
    return __ffma2_rn(v, make_float2(c, c), __fmul2_rn(make_float2(v.y, -v.x), make_float2(s, s)));
}
// v * (c + i s) = (x c - y s, y c + x s)
__device__ __forceinline__ float2 cmul(float2 v, float c, float s) {
    return __ffma2_rn(v, make_float2(c, c), __fmul2_rn(make_float2(-v.y, v.x), make_float2(s, s)));
}

// v *= W_N^P, W_N = exp(-2*pi*i/N); trivial factors are pure operand swizzles.
template <int N, int P> __device__ __forceinline__ void mul_tw(float2& v) {
    constexpr int p = ((P % N) + N) % N;
    if constexpr (p == 0) {
    } else if constexpr (4 * p == N) {          // -i
        v = make_float2(v.y, -v.x);
    } else if constexpr (2 * p == N) {          // -1
        v = make_float2(-v.x, -v.y);
    } else if constexpr (4 * p == 3 * N) {      // +i
        v = make_float2(-v.y, v.x);
    } else {
        constexpr float c = float(cos2pi(p, N));
        constexpr float s = float(sin2pi(p, N));
        v = cmul_conj(v, c, s);
    }
}

template <int N, int OFF, int J> struct R4Stage {
    static __device__ __forceinline__ void run(float2 (&v)[32]) {
        constexpr int Q = N / 4;
        if constexpr (J < Q) {
            constexpr int i0 = OFF + J, i1 = i0 + Q, i2 = i0 + 2 * Q, i3 = i0 + 3 * Q;
            const float2 t0 = cadd(v[i0], v[i2]), t1 = csub(v[i0], v[i2]);
            const float2 t2 = cadd(v[i1], v[i3]), d = csub(v[i1], v[i3]);
            float2 y0 = cadd(t0, t2), y2 = csub(t0, t2);
            float2 y1 = cadd_mi(t1, d), y3 = cadd_pi(t1, d);
            mul_tw<N, J>(y1);
            mul_tw<N, 2 * J>(y2);
            mul_tw<N, 3 * J>(y3);
            v[i0] = y0; v[i1] = y1; v[i2] = y2; v[i3] = y3;
            R4Stage<N, OFF, J + 1>::run(v);
        }
    }
};

template <int N, int OFF> struct FftDif {
    static __device__ __forceinline__ void run(float2 (&v)[32]) {
        if constexpr (N == 2) {
            const float2 a = v[OFF], b = v[OFF + 1];
            v[OFF] = cadd(a, b);
            v[OFF + 1] = csub(a, b);
        } else if constexpr (N >= 4) {
            R4Stage<N, OFF, 0>::run(v);
            FftDif<N / 4, OFF>::run(v);
            FftDif<N / 4, OFF + N / 4>::run(v);
            FftDif<N / 4, OFF + 2 * (N / 4)>::run(v);
            FftDif<N / 4, OFF + 3 * (N / 4)>::run(v);
        }
    }
};

// N-point forward FFT of v[off..off+N), in place; output bin k lands at v[off + fft_pos<N>(k)].
template <int N, int OFF = 0>
__device__ __forceinline__ void fft_dif(float2 (&v)[32]) {
    FftDif<N, OFF>::run(v);
}

}  // namespace fftreg2
