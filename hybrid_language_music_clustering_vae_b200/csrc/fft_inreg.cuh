// In-register complex FFTs (sizes 2..64) for the warp-per-frame STFT kernel.
//
// Decimation-in-frequency, radix-4 with a radix-2 tail (32 = 4*4*2, 16 = 4*4,
// 8 = 4*2).  Every index and every twiddle is a compile-time constant, so after
// inlining the data lives in registers and the twiddles are FFMA immediates.
// Output bin k of an N-point transform lands at register index fft_pos<N>(k)
// (digit reversal), which callers resolve at compile time.
//
// Replaces: scipy.fft.rfft inside librosa.stft; the
// reference reaches it from every feature call of src/1_preprocessing.py:50-83.
#pragma once

namespace fftreg {

constexpr double kPi = 3.14159265358979323846264338327950288;

// Taylor series, |x| <= pi: 30 terms leave < 1e-17 truncation error.
__host__ __device__ constexpr double csin(double x) {
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 30; ++i) { term *= -x2 / double((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
__host__ __device__ constexpr double ccos(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 30; ++i) { term *= -x2 / double((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
// cos / sin of 2*pi*p/n with the argument reduced to [-pi, pi].
__host__ __device__ constexpr double cos2pi(int p, int n) {
    p %= n; if (p < 0) p += n; if (2 * p > n) p -= n;
    return ccos(2.0 * kPi * double(p) / double(n));
}
__host__ __device__ constexpr double sin2pi(int p, int n) {
    p %= n; if (p < 0) p += n; if (2 * p > n) p -= n;
    return csin(2.0 * kPi * double(p) / double(n));
}

// Register index of output bin k.
template <int N> __host__ __device__ constexpr int fft_pos(int k) {
    if constexpr (N == 1) return 0;
    else if constexpr (N == 2) return k;
    else return (k % 4) * (N / 4) + fft_pos<N / 4>(k / 4);
}

// (x + iy) *= W_N^P, W_N = exp(-2*pi*i/N); trivial factors cost no multiply.
template <int N, int P> __device__ __forceinline__ void mul_tw(float& x, float& y) {
    constexpr int p = ((P % N) + N) % N;
    if constexpr (p == 0) {
    } else if constexpr (4 * p == N) {          // -i
        float t = x; x = y; y = -t;
    } else if constexpr (2 * p == N) {          // -1
        x = -x; y = -y;
    } else if constexpr (4 * p == 3 * N) {      // +i
        float t = x; x = -y; y = t;
    } else if constexpr (8 * p == N) {          // (1 - i)/sqrt2
        constexpr float r = 0.70710678118654752440f;
        float a = (x + y) * r, b = (y - x) * r; x = a; y = b;
    } else if constexpr (8 * p == 3 * N) {      // (-1 - i)/sqrt2
        constexpr float r = 0.70710678118654752440f;
        float a = (y - x) * r, b = -(x + y) * r; x = a; y = b;
    } else {
        constexpr float c = float(cos2pi(p, N));
        constexpr float s = float(sin2pi(p, N));
        float a = fmaf(x, c, y * s);            // xc + ys
        float b = fmaf(y, c, -(x * s));         // yc - xs
        x = a; y = b;
    }
}

template <int N, int OFF, int J> struct R4Stage {
    static __device__ __forceinline__ void run(float (&vr)[64], float (&vi)[64]) {
        constexpr int Q = N / 4;
        if constexpr (J < Q) {
            constexpr int i0 = OFF + J, i1 = i0 + Q, i2 = i0 + 2 * Q, i3 = i0 + 3 * Q;
            float t0r = vr[i0] + vr[i2], t0i = vi[i0] + vi[i2];
            float t1r = vr[i0] - vr[i2], t1i = vi[i0] - vi[i2];
            float t2r = vr[i1] + vr[i3], t2i = vi[i1] + vi[i3];
            float dr = vr[i1] - vr[i3], di = vi[i1] - vi[i3];
            // t3 = -i * (a1 - a3) = (di, -dr)
            float y0r = t0r + t2r, y0i = t0i + t2i;
            float y2r = t0r - t2r, y2i = t0i - t2i;
            float y1r = t1r + di, y1i = t1i - dr;
            float y3r = t1r - di, y3i = t1i + dr;
            mul_tw<N, J>(y1r, y1i);
            mul_tw<N, 2 * J>(y2r, y2i);
            mul_tw<N, 3 * J>(y3r, y3i);
            vr[i0] = y0r; vi[i0] = y0i;
            vr[i1] = y1r; vi[i1] = y1i;
            vr[i2] = y2r; vi[i2] = y2i;
            vr[i3] = y3r; vi[i3] = y3i;
            R4Stage<N, OFF, J + 1>::run(vr, vi);
        }
    }
};

template <int N, int OFF> struct FftDif {
    static __device__ __forceinline__ void run(float (&vr)[64], float (&vi)[64]) {
        if constexpr (N == 2) {
            float ar = vr[OFF], ai = vi[OFF], br = vr[OFF + 1], bi = vi[OFF + 1];
            vr[OFF] = ar + br; vi[OFF] = ai + bi;
            vr[OFF + 1] = ar - br; vi[OFF + 1] = ai - bi;
        } else if constexpr (N >= 4) {
            R4Stage<N, OFF, 0>::run(vr, vi);
            FftDif<N / 4, OFF>::run(vr, vi);
            FftDif<N / 4, OFF + N / 4>::run(vr, vi);
            FftDif<N / 4, OFF + 2 * (N / 4)>::run(vr, vi);
            FftDif<N / 4, OFF + 3 * (N / 4)>::run(vr, vi);
        }
    }
};

// N-point forward FFT of (vr[off..off+N), vi[off..off+N)), in place.
// The arrays are declared [64] so one signature serves every size; unused
// slots never materialise (everything is resolved at compile time).
template <int N, int OFF = 0>
__device__ __forceinline__ void fft_dif(float (&vr)[64], float (&vi)[64]) {
    FftDif<N, OFF>::run(vr, vi);
}

}  // namespace fftreg
