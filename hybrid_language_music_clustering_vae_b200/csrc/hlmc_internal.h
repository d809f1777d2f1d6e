// Internal interface between the C ABI (hlmc_capi.cu) and the kernels
// (hlmc_kernels.cu).  Not installed; the public contract is include/hlmc_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hlmc {

constexpr int kMaxMelGroups = 8;     // n_mels <= 256
constexpr int kFastNfft = 2048;      // the register-FFT path is specialised for this size

// Layout (in floats) of the table blob the fast kernel stages into shared memory.  For n_fft = 2048
// the window comes last so that kernels which synthesise it stage only the first `nowin` floats.
struct FastTables {
    int win;        // n_fft floats: 0.5 * window (the real-FFT split's 1/2 is folded in)
    int tw1;        // 31*32 float2: W_1024^(lane*k1), k1 = 1..31
    int tw2;        // n_fft 2048: 32 float2, -i * W_2048^(16*lane) (the kernel rotates it to bins 16*lane + i);
                    // n_fft 1024 / 512: 16*L float2, -i * W_n_fft^(16*lane + i)
    int mel_meta;   // int32: gmax[8], goff[8], qlo[32*n_groups]
    int mel_w;      // floats: per group, [i][lane] weights (zero padded to gmax)
    int total;      // floats, multiple of 4
    int n_groups;
    int scr;        // per-warp scratch floats (>= 1089 + max gmax, multiple of 4)
    int hann;       // 1: the window is the full-length periodic Hann (n_fft 2048: synthesised in registers)
    int hann_cs;    // 32 float4: (cos, cos', sin, sin') of 2*pi*(2*lane + {0,1}) / n_fft
    int nowin;      // floats before the window (multiple of 4)
    int mel_steps[32];   // float4 steps per filter group / round (TM kernels read them from the kernel parameters)
    // per-lane tables for Tensor Memory (TM kernels; global memory, [32 lanes][tmem_cols] floats) or NULL
    const float* tmem_tab;
    int tmem_cols;  // multiple of 4, <= 512
};

// Tables of the n_fft = 4096 register-FFT kernel (frames_fast_4096).
struct Fast4Tables {
    int tw1;          // 31*32 float2: W_1024^(lane*k1)
    int tw0;          // 1024 float2: W_2048^m, the radix-2 twiddle of the odd-bin pass
    int base;         // 2*32 float2: -i W_2048^(16 lane) (even bins), -i W_4096^(32 lane + 1) (odd bins)
    int hann_cs;      // 32 float4: (cos, cos', sin, sin') of 2*pi*(2*lane + {0,1}) / 4096
    int mel_meta[2];  // per parity: int32 gmax[8], goff[8], qlo[32*n_groups]
    int mel_w[2];     // per parity: banded weights of the de-interleaved filterbank columns
    int total;        // floats, multiple of 4
    int n_groups;     // <= 4 (n_mels <= 128)
    // per-lane tables for Tensor Memory (TM variant; global memory, [32 lanes][tmem_cols] floats) or NULL
    const float* tmem_tab;
    int tmem_cols;
    int mel_steps[8];     // float4 steps of (pass, group): [4 * pass + group]
    int mel_col[2];       // first TMEM column of each pass's weights
};

// Column map of the Tensor Memory tables (floats per lane; see hlmc_kernels.cu "Tensor Memory as a per-lane table store")
constexpr int kTmWin = 0;       // 64: 0.5*window[2(lane+32j)], 0.5*window[2(lane+32j)+1], j = 0..31
constexpr int kTmTw1 = 64;      // 64: W_1024^(lane*k1) as (cos, -sin), k1 = 1..31 (last pair unused)
constexpr int kTmTw2 = 128;     // 32: -i*W_2048^(16*lane + i), i = 0..15
constexpr int kTmMeta = 160;    // 8: first tap of the lane's filter in each of the (up to 8) groups (int32 bits)
constexpr int kTmMel = 168;     // banded mel weights, float4 per step, group after group
constexpr int kTmAlloc = 512;   // columns allocated (all of TMEM: one CTA per SM)
// ... and of frames_sub's (n_fft 1024 / 512; lane l reads the entries of lane-in-group l mod L)
constexpr int kSubWin = 0;      // 64: 0.5*window[2(lg+L*j)], [..+1], j = 0..31
constexpr int kSubTw1 = 64;     // 64: W_M^(lg*k1) as (cos, -sin), k1 = 1..31
constexpr int kSubTw2 = 128;    // 32: -i*W_N^(16*lg + i), i = 0..15
constexpr int kSubMeta = 160;   // 32: first tap of the lane's filter in each mel round (int32 bits)
constexpr int kSubMel = 192;    // banded mel weights, float4 per step, round after round
// ... and of frames_fast_4096's
constexpr int k4TmTw1 = 0;      // 64: W_1024^(lane*k1) as (cos, -sin), k1 = 1..31
constexpr int k4TmTw0 = 64;     // 64: W_2048^(lane + 32 j), j = 0..31 (the odd-bin pass's radix-2 twiddle)
constexpr int k4TmBase = 128;   // 4: split-twiddle base of the lane for pass 0, pass 1
constexpr int k4TmHcs = 132;    // 4: (cos, cos', sin, sin') of the lane's Hann phase
constexpr int k4TmMeta = 136;   // 8: first gather tap of the lane's filter, [4 * pass + group]
constexpr int k4TmMel = 144;    // banded mel weights: pass 0's groups, then pass 1's

struct FrameArgs {
    const float* wave;      // (B, pitch)
    long long pitch;
    int B, n, T;
    int n_fft, hop, pad;    // pad = center ? n_fft/2 : 0
    int pad_mode;
    int n_mels;
    int use_mag;            // mel of |X| (power == 1) instead of |X|^2
    float binhz;            // sr / n_fft
    float roll_percent;
    float zcr_thr;
    float* mel_out;         // mel power, (B, n_mels, T) or -- mel_frame_major -- (B, T, n_mels); may be NULL
    int mel_frame_major;    // 1: scratch layout for db_dct (coalesced full-sector stores)
    float* stats;           // (B, 5, T) or NULL
    int* status;            // (B) or NULL
    unsigned int* clipmax;  // (B) float bits, max of mel power; or NULL
    float* spec;            // (B, F, T, 2) complex STFT (generic kernel only) or NULL
    // piptrack candidates for chroma_stft's tuning estimate (register-FFT kernel only) or NULL
    float2* cand;           // (B, T, cand_cap) (pitch Hz, interpolated magnitude)
    int* cand_count;        // (B, T)
    int cand_cap;           // slots per frame
    int pip_klo, pip_khi;   // bins with fmin <= f < fmax
    float pip_threshold;    // 0.1
    // power-spectrum stash for the chroma projection (register-FFT kernel with piptrack only) or NULL:
    // (B, T, kStashFloats), each lane's 32 bins in its register order [lane][32], then bin 512
    float* pstash;
    int no_tmem;            // 1: keep the n_fft 2048 kernel's tables in shared memory (hlmc_plan_set_path, tests)
    // chroma_stft for any n_fft (frames_generic only): projection of the frame's power spectrum onto the clip's
    // tuned filterbank, chroma_out (B, 12, T); chroma_fb is (100, 12, 1 + n_fft/2) dense, chroma_tidx (B)
    float* chroma_out;
    const float* chroma_fb;
    const int* chroma_tidx;
};
constexpr int kStashFloats = 32 * 32 + 4;

struct ChromaArgs {
    const int* tuning_idx;  // (B) index into the 100 pre-built filterbanks
    const float* fb_all;    // (100, kChromaFbFloats) lane-major filterbanks
    float* chroma;          // (B, 12, T)
};
constexpr int kChroma = 12;
constexpr int kChromaFbFloats = kChroma * 32 * 32 + 16;   // [c][j/4][lane][4] + 12 weights of bin 512 (+pad)
constexpr int kTuningBins = 100;

// Generic-kernel tables (global memory).
struct GenericTables {
    const float* win;       // n_fft floats, 0.5 * window
    const float2* twm;      // M/2 entries: W_M^j
    const float2* tws;      // M/2+1 entries: -i * W_{2M}^k
    const int* mel_lo;      // n_mels
    const int* mel_len;     // n_mels
    const int* mel_off;     // n_mels
    const float* mel_w;     // nnz
};

struct DbArgs {
    float* mel;             // out: power_to_db (B, n_mels, T); also the mel-power input when mel_in == NULL
    const float* mel_in;    // mel power in frame-major scratch (B, T, n_mels), or NULL
    float* mfcc;            // (B, n_mfcc, T) or NULL
    const unsigned int* clipmax;
    const float* dct_t;     // (n_mels, ncp) transposed, zero padded DCT matrix
    int B, n_mels, n_mfcc, ncp, T;
    int ref_mode; float ref_value, amin, top_db;
};

// Extra arguments of the fused dB + DCT + time-pooling kernel (db_pool).
struct PoolArgs {
    const float* stats;     // (B, 5, T), written by the frames kernel
    const float* chroma;    // (B, n_chroma, T) or NULL
    int n_chroma;           // 0 or 12
    int with_mfcc;          // pool the MFCC rows too
    float* pooled;          // (B, pooled_w)
    int pooled_w;           // 2*n_mels + 2*n_mfcc*with_mfcc + 10 + 2*n_chroma
};

// launchers (all asynchronous on `stream`; return cudaError_t)
cudaError_t launch_frames_fast(const FrameArgs& a, const float* d_tables, const FastTables& ft,
                               int num_sms, cudaStream_t stream);
cudaError_t launch_frames_fast4096(const FrameArgs& a, const float* d_tables, const Fast4Tables& ft, int num_sms,
                                   cudaStream_t stream);
int fast4_smem_bytes(const Fast4Tables& ft, bool tm = false);
cudaError_t launch_frames_sub(const FrameArgs& a, const float* d_tables, const FastTables& ft, int num_sms,
                              cudaStream_t stream);
int sub_smem_bytes(const FastTables& ft, int nwarps, int L, bool tm = false);
cudaError_t launch_chroma_fast(const FrameArgs& a, const ChromaArgs& c, const float* d_tables,
                               const FastTables& ft, int num_sms, cudaStream_t stream);
cudaError_t launch_chroma_project(const float* pstash, const ChromaArgs& c, long long B, int T, int num_sms,
                                  cudaStream_t stream);
cudaError_t launch_tuning(const float2* cand, const int* cand_count, int T, int cand_per_frame, long long B,
                          const double* d_edges, float* tuning, int* tuning_idx, cudaStream_t stream);
cudaError_t launch_frames_generic(const FrameArgs& a, const GenericTables& gt, cudaStream_t stream);
cudaError_t launch_db_dct(const DbArgs& a, cudaStream_t stream);
cudaError_t launch_db_pool(const DbArgs& a, const PoolArgs& pa, int num_sms, cudaStream_t stream);
bool db_pool_fits(int n_mels, int ncp, int T);   // the fused kernel's tile fits shared memory
cudaError_t launch_rowmax(const float* in, unsigned int* clipmax, long long B, long long per_clip,
                          cudaStream_t stream);
cudaError_t launch_power_to_db(const float* in, float* out, const unsigned int* clipmax, long long B,
                               long long per_clip, int ref_mode, float ref_value, float amin,
                               float top_db, cudaStream_t stream);
cudaError_t launch_pool(const float* logmel, const float* mfcc, const float* stats, const float* chroma,
                        long long B, int n_mels, int n_mfcc, int T, float* pooled, cudaStream_t stream);
cudaError_t launch_fix_frames(const float* in, float* out, long long B, int rows, int T, int fixed,
                              cudaStream_t stream);
cudaError_t launch_pcm16_to_f32(const int16_t* raw, long long raw_pitch, float* out, long long pitch,
                                long long B, long long n_valid, long long n_total, cudaStream_t stream);
cudaError_t launch_colstats(const float* x, long long N, long long D, double* mean, double* m2,
                            cudaStream_t stream);
cudaError_t launch_standardize(const float* x, float* y, long long N, long long D, const float* mean,
                               const float* scale, cudaStream_t stream);
cudaError_t measure_fp32_peak(double* tflops);

// librosa.load's front end (hlmc_frontend.cu)
struct FrontArgs {
    const void* raw;        // (B, raw_pitch, channels) int16 or float32, interleaved channels
    int fmt, channels;      // fmt 0: float32, 1: int16
    long long B, raw_pitch, n_in;
    float* out;             // (B, pitch) float32
    long long pitch, n_total, n_out;   // n_out resampled samples, then zeros up to n_total
    int up, down;           // resample_poly factors after the gcd; 1, 1 = no resampling
    int n_pre_pad;          // zeros scipy prepends to the filter
    long long n_pre_remove; // outputs scipy drops at the front
    int tpp;                // taps per phase
    const float* hpoly;     // (up, tpp): hpoly[p][t] = h[p + t*up] * up, zero past the filter
    const long long* valid; // (B) per-clip input frames (<= n_in) or NULL
};
cudaError_t launch_frontend(const FrontArgs& a, int num_sms, cudaStream_t stream);
// tabular normalisation (hlmc_frontend.cu)
cudaError_t launch_impute_stats(const double* x, long long N, long long D, double* sum, long long* count,
                                cudaStream_t st);
cudaError_t launch_scaler_stats_f64(const double* x, long long N, long long D, const double* fill, double* mean,
                                    double* m2, double* scratch, cudaStream_t st);
cudaError_t launch_impute_scale(const double* x, long long N, long long D, const int* cols, long long Do,
                                const double* fill, const double* mean, const double* scale, double* imputed,
                                double* scaled, cudaStream_t st);
void count_launch(int n);
int fast_smem_bytes(const FastTables& ft, int nwarps, int n_fft, int hop, int n_mels);
int pick_fast_warps(int T);
long long launch_count();

}  // namespace hlmc
