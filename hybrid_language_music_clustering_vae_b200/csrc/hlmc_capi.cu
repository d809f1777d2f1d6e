// C ABI of the B200-native feature extractor: plan construction (host tables in
// float64, rounded once), parameter validation with librosa's error behaviour,
// the device entry points and the chunked host<->device pipeline.
// Public contract: include/hlmc_b200.h.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <string>
#include <vector>

#include "../../include/hlmc_b200.h"
#include "hlmc_internal.h"

using namespace hlmc;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return HLMC_ERR_CUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);      \
    } while (0)

struct HostPipeSlot {
    cudaStream_t stream = nullptr;
    float* d_wave = nullptr; float* d_logmel = nullptr; float* d_mfcc = nullptr;
    float* d_stats = nullptr; float* d_pooled = nullptr; float* d_clipmax = nullptr;
    int32_t* d_status = nullptr;
    void* d_raw = nullptr; size_t raw_bytes = 0;        // staging of int16 / multi-channel / other-rate input
    float* d_fixed = nullptr; size_t fixed_elems = 0;    // (chunk, n_mels, fixed_frames) image of the advanced script
    long long* d_valid = nullptr; size_t valid_elems = 0; // per-clip frame counts
    float* d_melscr = nullptr;                           // frame-major mel-power scratch
    float* d_chroma = nullptr; float* d_tuning = nullptr; void* d_cwork = nullptr; size_t chroma_ws = 0;
};

// scipy.signal.resample_poly's default low-pass for one input rate, regrouped by phase for the device
struct ResampleTaps {
    int sr_in = 0, up = 1, down = 1, half_len = 0, n_pre_pad = 0, tpp = 0;
    long long n_pre_remove = 0;
    std::vector<float> h;            // firwin(...) cast to float32, times up (what scipy convolves with)
    float* d_hpoly = nullptr;        // (up, tpp)
};

struct hlmc_plan {
    hlmc_params p;
    int device = 0, num_sms = 0, F = 0;
    int force_generic = 0, no_tmem = 0;
    bool fast_ok = false;
    std::vector<float> mel_dense, dct;           // host copies
    int ncp = 0;
    // device tables
    float* d_fast = nullptr; FastTables ft{};
    float* d_tmem_tab = nullptr;                                          // per-lane tables for Tensor Memory
    float* d_fast4 = nullptr; Fast4Tables ft4{}; bool fast4_ok = false;   // n_fft = 4096 register-FFT kernel
    float* d_win = nullptr; float2* d_twm = nullptr; float2* d_tws = nullptr;
    int* d_mel_lo = nullptr; int* d_mel_len = nullptr; int* d_mel_off = nullptr; float* d_mel_w = nullptr;
    float* d_dct_t = nullptr;
    // host pipeline
    std::vector<HostPipeSlot> slots;
    int64_t slot_chunk = 0, slot_n = 0; int slot_flags = 0;
    int64_t last_h2d = 0, last_d2h = 0;
    // chroma_stft (built lazily): 100 filterbanks (one per tuning bin) + histogram edges
    float* d_chroma_fb = nullptr; double* d_edges = nullptr;
    int pip_klo = 0, pip_khi = 0, cand_per_frame = 0;
    bool chroma_generic = false;                 // n_fft != 2048: candidates and projection through frames_generic
    // optional per-kernel timing (events recorded on the launching stream)
    int timing = 0;
    std::vector<cudaEvent_t> ev;      // triples: before frames, after frames, after db_dct
    double t_frames_ms = 0.0, t_db_ms = 0.0; int64_t t_calls = 0;   // folded triples (the list is bounded)
    std::vector<ResampleTaps> resamplers;   // built at first use, one per input rate
};

// ---------------------------------------------------------------------------
// librosa.filters.mel restated in float64
// ---------------------------------------------------------------------------
static double hz_to_mel(double f, bool htk) {
    if (htk) return 2595.0 * log10(1.0 + f / 700.0);
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    if (f >= min_log_hz) return min_log_mel + log(f / min_log_hz) / logstep;
    return f / f_sp;
}
static double mel_to_hz(double m, bool htk) {
    if (htk) return 700.0 * (pow(10.0, m / 2595.0) - 1.0);
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    if (m >= min_log_mel) return min_log_hz * exp(logstep * (m - min_log_mel));
    return f_sp * m;
}
static void build_mel(const hlmc_params& p, std::vector<float>& w) {
    const int F = p.n_fft / 2 + 1, nm = p.n_mels;
    const double fmax = (p.fmax > 0.0f) ? double(p.fmax) : double(p.sr) / 2.0;
    const double val = 1.0 / (double(p.n_fft) * (1.0 / double(p.sr)));   // np.fft.rfftfreq
    std::vector<double> mel_f(nm + 2);
    const double m0 = hz_to_mel(p.fmin, p.htk != 0), m1 = hz_to_mel(fmax, p.htk != 0);
    const double step = (m1 - m0) / double(nm + 1);
    for (int i = 0; i < nm + 2; ++i) mel_f[i] = mel_to_hz(i == nm + 1 ? m1 : m0 + i * step, p.htk != 0);
    w.assign((size_t)nm * F, 0.0f);
    for (int i = 0; i < nm; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = (p.mel_norm == HLMC_MELNORM_SLANEY) ? 2.0 / (mel_f[i + 2] - mel_f[i]) : 1.0;
        for (int k = 0; k < F; ++k) {
            const double fk = double(k) * val;
            const double lower = -(mel_f[i] - fk) / fd0, upper = (mel_f[i + 2] - fk) / fd1;
            double v = lower < upper ? lower : upper;
            if (!(v > 0.0)) v = 0.0;
            // librosa stores float32 weights, then scales the float32 array by float64 enorm
            w[(size_t)i * F + k] = float(double(float(v)) * enorm);
        }
    }
}

static int validate(const hlmc_params& p) {
    if (p.sr <= 0) return fail(HLMC_ERR_PARAM, "sr must be positive");
    if (p.n_fft <= 0) return fail(HLMC_ERR_PARAM, "n_fft must be a positive integer");
    if (p.hop_length <= 0) return fail(HLMC_ERR_PARAM, "hop_length must be a positive integer");
    if (p.win_length <= 0 || p.win_length > p.n_fft)
        return fail(HLMC_ERR_PARAM, "win_length must satisfy 0 < win_length <= n_fft (util.pad_center)");
    if (p.n_fft < 64 || p.n_fft > 8192 || (p.n_fft & (p.n_fft - 1)))
        return fail(HLMC_ERR_UNSUPPORTED, "n_fft must be a power of two in [64, 8192]");
    if (p.pad_mode < 0 || p.pad_mode > 2) return fail(HLMC_ERR_PARAM, "unsupported pad_mode");
    if (p.n_mels < 1 || p.n_mels > 32 * kMaxMelGroups) return fail(HLMC_ERR_UNSUPPORTED, "n_mels must be in [1, 256]");
    if (p.n_mfcc < 0) return fail(HLMC_ERR_PARAM, "n_mfcc must be non-negative");
    if (p.n_mfcc > p.n_mels || p.n_mfcc > 128) return fail(HLMC_ERR_UNSUPPORTED, "n_mfcc must be <= min(n_mels, 128)");
    if (!(p.power == 1.0f || p.power == 2.0f)) return fail(HLMC_ERR_UNSUPPORTED, "power must be 1.0 or 2.0");
    if (!(p.amin > 0.0f)) return fail(HLMC_ERR_PARAM, "amin must be strictly positive");
    if (!(p.roll_percent > 0.0f && p.roll_percent < 1.0f)) return fail(HLMC_ERR_PARAM, "roll_percent must lie in the range (0, 1)");
    if (p.lifter < 0.0f) return fail(HLMC_ERR_PARAM, "MFCC lifter must be a non-negative number");
    if (p.fmin < 0.0f) return fail(HLMC_ERR_PARAM, "fmin must be non-negative");
    if (p.zcr_threshold < 0.0f) return fail(HLMC_ERR_PARAM, "zero-crossing threshold must be non-negative");
    if (p.ref_mode != HLMC_REF_VALUE && p.ref_mode != HLMC_REF_MAX) return fail(HLMC_ERR_PARAM, "bad ref_mode");
    return HLMC_OK;
}

// Banded layout of a dense (nm x F) filterbank for the register-FFT kernels' mel gather: filters in groups
// of 32 (one per lane), each group padded to a common number of float4 steps, every lane's first tap
// shifted so that the 32 lanes start on 32 different shared-memory banks.  meta: [0..7] float4 steps per
// group, [8..15] weight offset per group, then the (shifted) first tap of every (group, lane) in the
// padded spectrum layout q(k) = k + k/16.
static void build_banded_groups(const std::vector<float>& dense, int nm, int F, int kReadEnd,
                                std::vector<int>& meta, std::vector<float>& melw) {
    const int ng = (nm + 31) / 32;
    meta.assign(2 * kMaxMelGroups + 32 * ng, 0);
    melw.clear();
    std::vector<int> lo(nm, 0), len(nm, 0);
    for (int m = 0; m < nm; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < F; ++k)
            if (dense[(size_t)m * F + k] != 0.0f) { if (first < 0) first = k; last = k; }
        if (first >= 0) { lo[m] = first; len[m] = last - first + 1; }
    }
    auto qof = [](int k) { return k + (k >> 4); };
    for (int g = 0; g < ng; ++g) {
        int ql[32], qlen[32];
        int gmax = 0;
        for (int l = 0; l < 32; ++l) {
            const int m = 32 * g + l;
            if (m < nm && len[m] > 0) {
                ql[l] = qof(lo[m]);
                qlen[l] = qof(lo[m] + len[m] - 1) - ql[l] + 1;
            } else { ql[l] = 0; qlen[l] = 0; }
            if (qlen[l] > gmax) gmax = qlen[l];
        }
        // Shift each lane's first tap down (zero weights in front) so that the 32 lanes
        // start on 32 different banks: the gather is then conflict-free at every step.
        // Bipartite matching lanes -> banks; grow the common length G until one exists.
        int G = (gmax + 3) & ~3, shiftv[32];
        for (int l = 0; l < 32; ++l) shiftv[l] = 0;
        bool ok = (G == 0);
        for (int tries = 0; !ok && tries < 12; ++tries, G += 4) {
            std::vector<std::vector<int>> opt(32);   // candidate shifts per lane
            bool feasible = true;
            for (int l = 0; l < 32; ++l) {
                if (qlen[l] == 0) { for (int s = 0; s < 32; ++s) opt[l].push_back(-s); continue; }  // start = s
                const int smin = std::max(0, ql[l] + G - kReadEnd);
                const int smax = std::min(G - qlen[l], ql[l]);
                if (smin > smax) { feasible = false; break; }
                for (int s = smin; s <= smax && s < smin + 32; ++s) opt[l].push_back(s);
            }
            if (!feasible) continue;
            int owner[32]; for (int& o : owner) o = -1;
            auto bank_of = [&](int l, int s) { return ((qlen[l] == 0 ? -s : ql[l] - s) % 32 + 32) % 32; };
            std::vector<char> seen(32);
            std::function<bool(int)> aug = [&](int l) -> bool {
                for (int s : opt[l]) {
                    const int bk = bank_of(l, s);
                    if (seen[bk]) continue;
                    seen[bk] = 1;
                    if (owner[bk] < 0 || aug(owner[bk])) { owner[bk] = l; return true; }
                }
                return false;
            };
            int matched = 0;
            for (int l = 0; l < 32; ++l) { std::fill(seen.begin(), seen.end(), 0); if (aug(l)) ++matched; }
            if (matched == 32) {
                for (int bk = 0; bk < 32; ++bk) {
                    const int l = owner[bk];
                    for (int s : opt[l]) if (bank_of(l, s) == bk) { shiftv[l] = s; break; }
                }
                ok = true;
                break;
            }
        }
        if (!ok) {   // no conflict-free placement: keep the natural starts (still correct)
            G = (gmax + 3) & ~3;
            for (int l = 0; l < 32; ++l) shiftv[l] = std::max(0, ql[l] + G - kReadEnd);
        }
        meta[g] = G / 4;
        meta[kMaxMelGroups + g] = (int)melw.size();
        const size_t base = melw.size();
        melw.resize(base + (size_t)32 * G, 0.0f);
        for (int l = 0; l < 32; ++l) {
            const int m = 32 * g + l;
            const int start = (qlen[l] == 0) ? -shiftv[l] : ql[l] - shiftv[l];
            meta[2 * kMaxMelGroups + 32 * g + l] = start;
            if (qlen[l] == 0) continue;
            for (int q = ql[l]; q < ql[l] + qlen[l]; ++q) {
                if (q % 17 == 16) continue;                     // pad slot
                const int k = q - q / 17, i = q - start;
                melw[base + ((size_t)(i / 4) * 32 + l) * 4 + (i % 4)] = dense[(size_t)m * F + k];
            }
        }
    }
}

template <class T>
static cudaError_t upload(T** dptr, const std::vector<T>& h) {
    const size_t bytes = (h.empty() ? 1 : h.size()) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)dptr, bytes);
    if (e != cudaSuccess) return e;
    if (!h.empty()) e = cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

// librosa.filters.chroma(sr, n_fft, tuning, n_chroma=12, ctroct=5, octwidth=2, norm=2, base_c=True)
// restated in float64; returns (12, 1 + n_fft/2) float32, row 0 = C.
static void build_chroma_fb(int sr, int n_fft, double tuning, std::vector<float>& out) {
    const int nc = 12, F = n_fft / 2 + 1;
    std::vector<double> frq(n_fft), bw(n_fft);
    const double a440 = 440.0 * pow(2.0, tuning / nc);
    const double step = double(sr) / double(n_fft);
    for (int i = 1; i < n_fft; ++i) frq[i] = nc * log2((i * step) / (a440 / 16.0));
    frq[0] = frq[1] - 1.5 * nc;
    for (int i = 0; i + 1 < n_fft; ++i) bw[i] = std::max(frq[i + 1] - frq[i], 1.0);
    bw[n_fft - 1] = 1.0;
    std::vector<double> w((size_t)nc * F);
    for (int i = 0; i < F; ++i) {
        double col[12], nrm = 0.0;
        for (int c = 0; c < nc; ++c) {
            double D = fmod(frq[i] - c + 6.0 + 10.0 * nc, double(nc)) - 6.0;
            const double z = 2.0 * D / bw[i];
            col[c] = exp(-0.5 * z * z);
            nrm += col[c] * col[c];
        }
        nrm = sqrt(nrm);
        if (nrm < 2.2250738585072014e-308) nrm = 1.0;
        const double o = (frq[i] / nc - 5.0) / 2.0;
        const double oct = exp(-0.5 * o * o);
        for (int c = 0; c < nc; ++c) w[(size_t)c * F + i] = col[c] / nrm * oct;
    }
    out.resize((size_t)nc * F);
    for (int c = 0; c < nc; ++c)                      // np.roll(wts, -3, axis=0)
        for (int i = 0; i < F; ++i) out[(size_t)c * F + i] = float(w[(size_t)((c + 3) % nc) * F + i]);
}

static int ensure_chroma_tables(hlmc_plan* pl) {
    if (pl->d_chroma_fb) return HLMC_OK;
    // n_fft = 2048 (the scripts' size): piptrack epilogue + power-spectrum stash of the register-FFT kernel.  Any
    // other n_fft: two extra passes of the shared-memory FFT kernel (candidates, then the projection).
    pl->chroma_generic = (pl->p.n_fft != kFastNfft || !pl->fast_ok);
    if (!pl->chroma_generic && pl->p.power != 2.0f)
        return fail(HLMC_ERR_UNSUPPORTED, "chroma_stft on device needs a power=2 plan");
    CK(cudaSetDevice(pl->device));       // callers may be on a fresh host thread whose current device is 0
    const int F = pl->F;
    // bins np.linspace(-0.5, 0.5, 101) of librosa.pitch_tuning (resolution 0.01)
    std::vector<double> edges(kTuningBins + 1);
    const double step = (0.5 - (-0.5)) / double(kTuningBins);
    for (int i = 0; i <= kTuningBins; ++i) edges[i] = (i == kTuningBins) ? 0.5 : i * step + (-0.5);
    const size_t per_fb = pl->chroma_generic ? (size_t)kChroma * F : (size_t)kChromaFbFloats;
    std::vector<float> all((size_t)kTuningBins * per_fb, 0.0f), fb;
    for (int tb = 0; tb < kTuningBins; ++tb) {
        build_chroma_fb(pl->p.sr, pl->p.n_fft, edges[tb], fb);
        float* dst = &all[(size_t)tb * per_fb];
        if (pl->chroma_generic) {                      // dense (12, F) rows, as librosa.filters.chroma returns them
            memcpy(dst, fb.data(), per_fb * 4);
            continue;
        }
        for (int c = 0; c < kChroma; ++c) {
            for (int l = 0; l < 32; ++l)
                for (int j = 0; j < 32; ++j) {
                    const int k = (j < 16) ? 16 * l + j : 1024 - 16 * l - (j - 16);
                    dst[(((size_t)c * 8 + j / 4) * 32 + l) * 4 + (j % 4)] = fb[(size_t)c * F + k];
                }
            dst[kChroma * 32 * 32 + c] = fb[(size_t)c * F + 512];
        }
    }
    // piptrack's frequency mask: fmin=150 <= f < fmax=4000 (librosa.estimate_tuning defaults)
    const double val = 1.0 / (double(pl->p.n_fft) * (1.0 / double(pl->p.sr)));
    const double fmax = std::min(4000.0, double(pl->p.sr) / 2.0);
    int klo = 1, khi = 1;
    while (klo < F - 1 && klo * val < 150.0) ++klo;
    khi = klo;
    while (khi < F - 1 && khi * val < fmax) ++khi;
    pl->pip_klo = klo; pl->pip_khi = khi;
    pl->cand_per_frame = (khi - klo + 1) / 2 + 1;
    cudaError_t e;
    if ((e = upload(&pl->d_chroma_fb, all)) != cudaSuccess) return cuda_fail(e, "upload chroma filterbanks");
    if ((e = upload(&pl->d_edges, edges)) != cudaSuccess) return cuda_fail(e, "upload tuning edges");
    return HLMC_OK;
}

static void free_slots(hlmc_plan* pl) {
    for (auto& s : pl->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.d_wave); cudaFree(s.d_logmel); cudaFree(s.d_mfcc); cudaFree(s.d_stats);
        cudaFree(s.d_pooled); cudaFree(s.d_clipmax); cudaFree(s.d_status); cudaFree(s.d_raw); cudaFree(s.d_melscr);
        cudaFree(s.d_chroma); cudaFree(s.d_tuning); cudaFree(s.d_cwork); cudaFree(s.d_fixed); cudaFree(s.d_valid);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    pl->slots.clear();
    pl->slot_chunk = 0; pl->slot_n = -1; pl->slot_flags = 0;
}

// Fold finished timing triples into the running totals (keeps plan->ev bounded while timing is on).
static int fold_timing(hlmc_plan* plan) {
    for (size_t i = 0; i + 2 < plan->ev.size(); i += 3) {
        float a = 0.f, b = 0.f;
        CK(cudaEventSynchronize(plan->ev[i + 2]));
        CK(cudaEventElapsedTime(&a, plan->ev[i], plan->ev[i + 1]));
        CK(cudaEventElapsedTime(&b, plan->ev[i + 1], plan->ev[i + 2]));
        plan->t_frames_ms += a; plan->t_db_ms += b; ++plan->t_calls;
    }
    for (auto e : plan->ev) cudaEventDestroy(e);
    plan->ev.clear();
    return HLMC_OK;
}

extern "C" {

int hlmc_abi_version(void) { return HLMC_ABI_VERSION; }
const char* hlmc_last_error(void) { return g_err.c_str(); }
int64_t hlmc_launch_count(void) { return launch_count(); }

void hlmc_params_default(hlmc_params* p) {
    if (!p) return;
    p->sr = 22050; p->n_fft = 2048; p->hop_length = 512; p->win_length = 2048;
    p->center = 1; p->pad_mode = HLMC_PAD_CONSTANT;
    p->n_mels = 128; p->fmin = 0.0f; p->fmax = -1.0f; p->htk = 0; p->mel_norm = HLMC_MELNORM_SLANEY;
    p->power = 2.0f; p->n_mfcc = 20; p->lifter = 0.0f;
    p->ref_mode = HLMC_REF_VALUE; p->ref_value = 1.0f; p->amin = 1e-10f; p->top_db = 80.0f;
    p->roll_percent = 0.85f; p->zcr_threshold = 1e-10f;
}

int64_t hlmc_num_frames(const hlmc_params* p, int64_t n) {
    if (!p || p->hop_length <= 0 || p->n_fft <= 0) return fail(HLMC_ERR_PARAM, "bad params");
    if (n < 1) return fail(HLMC_ERR_PARAM, "Input is too short");
    if (n > 0x3fffffff) return fail(HLMC_ERR_UNSUPPORTED, "clips longer than 2^30 samples are not supported");
    if (p->center) return 1 + (n + 2 * (p->n_fft / 2) - p->n_fft) / p->hop_length;
    if (n < p->n_fft) return fail(HLMC_ERR_PARAM, "n_fft is too large for uncentered analysis of this input");
    return 1 + (n - p->n_fft) / p->hop_length;
}

void hlmc_plan_destroy(hlmc_plan* plan) {
    if (!plan) return;
    cudaSetDevice(plan->device);
    free_slots(plan);
    for (auto e : plan->ev) cudaEventDestroy(e);
    for (auto& r : plan->resamplers) cudaFree(r.d_hpoly);
    cudaFree(plan->d_fast); cudaFree(plan->d_tmem_tab); cudaFree(plan->d_fast4); cudaFree(plan->d_win); cudaFree(plan->d_twm); cudaFree(plan->d_tws);
    cudaFree(plan->d_mel_lo); cudaFree(plan->d_mel_len); cudaFree(plan->d_mel_off); cudaFree(plan->d_mel_w);
    cudaFree(plan->d_dct_t); cudaFree(plan->d_chroma_fb); cudaFree(plan->d_edges);
    delete plan;
}

int hlmc_plan_create(const hlmc_params* params, const double* window, const float* mel_basis,
                     int device, hlmc_plan** out) {
    if (!params || !out) return fail(HLMC_ERR_PARAM, "null argument");
    *out = nullptr;
    hlmc_params p = *params;
    if (p.win_length <= 0) p.win_length = p.n_fft;
    int rc = validate(p);
    if (rc != HLMC_OK) return rc;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0)
        return fail(HLMC_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(HLMC_ERR_PARAM, "bad device ordinal");
    CK(cudaSetDevice(device));

    hlmc_plan* pl = new hlmc_plan();
    pl->p = p; pl->device = device; pl->F = p.n_fft / 2 + 1;
    cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, device);
    {   // the per-call mel scratch comes from the stream-ordered pool: do not hand it back at every sync
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    const char* fg = getenv("HLMC_FORCE_GENERIC");
    pl->force_generic = (fg && fg[0] == '1') ? 1 : 0;
    const int N = p.n_fft, M = N / 2, F = pl->F;

    // window: scipy.signal.get_window(...) -> util.pad_center(size=n_fft); 0.5 folded in
    std::vector<float> win(N, 0.0f);
    {
        const int lpad = (N - p.win_length) / 2;
        for (int i = 0; i < p.win_length; ++i) {
            const double w = window ? window[i] : 0.5 - 0.5 * cos(2.0 * M_PI * double(i) / double(p.win_length));
            win[lpad + i] = float(0.5 * w);
        }
    }
    // mel filterbank, dense then banded
    if (mel_basis) pl->mel_dense.assign(mel_basis, mel_basis + (size_t)p.n_mels * F);
    else build_mel(p, pl->mel_dense);
    std::vector<int> lo(p.n_mels, 0), len(p.n_mels, 0), off(p.n_mels, 0);
    std::vector<float> wpack;
    for (int m = 0; m < p.n_mels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < F; ++k)
            if (pl->mel_dense[(size_t)m * F + k] != 0.0f) { if (first < 0) first = k; last = k; }
        off[m] = (int)wpack.size();
        if (first >= 0) {
            lo[m] = first; len[m] = last - first + 1;
            for (int k = first; k <= last; ++k) wpack.push_back(pl->mel_dense[(size_t)m * F + k]);
        }
    }
    // DCT-II (scipy.fftpack.dct type=2 norm="ortho"), first n_mfcc rows, lifter folded in
    const int nm = p.n_mels, nc = p.n_mfcc;
    pl->dct.assign((size_t)nc * nm, 0.0f);
    for (int c = 0; c < nc; ++c) {
        const double s = (c == 0) ? sqrt(1.0 / nm) : sqrt(2.0 / nm);
        // librosa.feature.mfcc lifter: M *= 1 + (lifter/2) * sin(pi * (c+1) / lifter)
        const double lf = (p.lifter > 0.0f) ? 1.0 + (double(p.lifter) / 2.0) * sin(M_PI * double(c + 1) / double(p.lifter)) : 1.0;
        for (int m = 0; m < nm; ++m)
            pl->dct[(size_t)c * nm + m] = float(lf * s * cos(M_PI * double(c) * (2.0 * m + 1.0) / (2.0 * nm)));
    }
    const int ncps[7] = {8, 16, 24, 32, 40, 64, 128};
    pl->ncp = 0;
    for (int v : ncps) if (nc > 0 && v >= nc) { pl->ncp = v; break; }
    std::vector<float> dct_t((size_t)nm * (pl->ncp > 0 ? pl->ncp : 1), 0.0f);
    for (int c = 0; c < nc; ++c)
        for (int m = 0; m < nm; ++m) dct_t[(size_t)m * pl->ncp + c] = pl->dct[(size_t)c * nm + m];

    // generic-kernel tables
    std::vector<float2> twm(M / 2), tws(M / 2 + 1);
    for (int j = 0; j < M / 2; ++j) {
        const double th = 2.0 * M_PI * double(j) / double(M);
        twm[j] = make_float2(float(cos(th)), float(-sin(th)));
    }
    for (int k = 0; k <= M / 2; ++k) {
        const double th = 2.0 * M_PI * double(k) / double(N);
        tws[k] = make_float2(float(-sin(th)), float(-cos(th)));     // -i * W_N^k
    }
    cudaError_t e;
#define UP(dst, vec) if ((e = upload(&pl->dst, vec)) != cudaSuccess) { hlmc_plan_destroy(pl); return cuda_fail(e, "upload " #dst); }
    UP(d_win, win) UP(d_twm, twm) UP(d_tws, tws) UP(d_mel_lo, lo) UP(d_mel_len, len) UP(d_mel_off, off)
    UP(d_mel_w, wpack) UP(d_dct_t, dct_t)

    // fast-kernel table blob (n_fft == 2048 only)
    if (N == kFastNfft) {
        FastTables ft{};
        const int ng = (nm + 31) / 32;
        ft.n_groups = ng;
        // meta: [0..7] float4 steps per group, [8..15] weight offset per group, then the
        // (shifted) first tap of every (group, lane) in the padded layout q(k) = k + k/16
        std::vector<int> meta;
        std::vector<float> melw;
        const int kReadEnd = 1105;           // the kernel keeps scratch [0, 1105) finite
        build_banded_groups(pl->mel_dense, nm, F, kReadEnd, meta, melw);
        auto r4 = [](int x) { return (x + 3) & ~3; };
        for (int g = 0; g < ng; ++g) ft.mel_steps[g] = meta[g];
        ft.tw1 = 0;
        ft.tw2 = ft.tw1 + 31 * 32 * 2;
        ft.hann_cs = ft.tw2 + 32 * 2;
        ft.mel_meta = ft.hann_cs + 32 * 4;
        ft.mel_w = r4(ft.mel_meta + (int)meta.size());
        ft.nowin = r4(ft.mel_w + (int)melw.size());
        ft.win = ft.nowin;
        ft.total = ft.win + N;
        // full-length periodic Hann (the default): the kernel may synthesise the window from 4 values per lane
        ft.hann = (p.win_length == N) ? 1 : 0;
        for (int i = 0; window && ft.hann && i < N; ++i)          // a caller-supplied window counts if it IS that Hann
            if (fabs(window[i] - (0.5 - 0.5 * cos(2.0 * M_PI * double(i) / double(N)))) > 1e-9) ft.hann = 0;
        ft.scr = r4(kReadEnd + 3);
        std::vector<float> blob(ft.total, 0.0f);
        memcpy(&blob[ft.win], win.data(), N * 4);
        for (int k1 = 1; k1 < 32; ++k1)
            for (int l = 0; l < 32; ++l) {
                const double th = 2.0 * M_PI * double((l * k1) % 1024) / 1024.0;
                blob[ft.tw1 + ((k1 - 1) * 32 + l) * 2 + 0] = float(cos(th));
                blob[ft.tw1 + ((k1 - 1) * 32 + l) * 2 + 1] = float(-sin(th));
            }
        for (int l = 0; l < 32; ++l) {
            const double th = 2.0 * M_PI * double(16 * l) / 2048.0;           // -i * W_2048^(16 l)
            blob[ft.tw2 + l * 2 + 0] = float(-sin(th));
            blob[ft.tw2 + l * 2 + 1] = float(-cos(th));
            const double te = 2.0 * M_PI * double(2 * l) / double(N), to = 2.0 * M_PI * double(2 * l + 1) / double(N);
            blob[ft.hann_cs + 4 * l + 0] = float(cos(te));
            blob[ft.hann_cs + 4 * l + 1] = float(cos(to));
            blob[ft.hann_cs + 4 * l + 2] = float(sin(te));
            blob[ft.hann_cs + 4 * l + 3] = float(sin(to));
        }
        memcpy(&blob[ft.mel_meta], meta.data(), meta.size() * 4);
        if (!melw.empty()) memcpy(&blob[ft.mel_w], melw.data(), melw.size() * 4);
        pl->ft = ft;
        UP(d_fast, blob)
        pl->fast_ok = fast_smem_bytes(ft, 8, N, p.hop_length, nm) <= 227 * 1024;
        {
            // the same tables per lane for Tensor Memory (frames_fast_2048<..., TM>): [32][cols], see kTm* in
            // hlmc_internal.h; built whenever the banded weights fit the 512 columns
            int total_steps = 0;
            for (int g = 0; g < ng; ++g) total_steps += meta[g];
            const int cols = kTmMel + 4 * total_steps;
            if (cols <= kTmAlloc) {
                std::vector<float> tm((size_t)32 * cols, 0.0f);
                for (int l = 0; l < 32; ++l) {
                    float* row = &tm[(size_t)l * cols];
                    for (int j = 0; j < 32; ++j) {
                        row[kTmWin + 2 * j] = win[2 * (l + 32 * j)];
                        row[kTmWin + 2 * j + 1] = win[2 * (l + 32 * j) + 1];
                    }
                    for (int k1 = 1; k1 < 32; ++k1) {
                        row[kTmTw1 + 2 * (k1 - 1)] = blob[ft.tw1 + ((k1 - 1) * 32 + l) * 2];
                        row[kTmTw1 + 2 * (k1 - 1) + 1] = blob[ft.tw1 + ((k1 - 1) * 32 + l) * 2 + 1];
                    }
                    for (int i = 0; i < 16; ++i) {
                        const double th = 2.0 * M_PI * double(16 * l + i) / 2048.0;      // -i * W_2048^(16 l + i)
                        row[kTmTw2 + 2 * i] = float(-sin(th));
                        row[kTmTw2 + 2 * i + 1] = float(-cos(th));
                    }
                    int pre = 0;
                    for (int g = 0; g < ng; ++g) {
                        memcpy(&row[kTmMeta + g], &meta[2 * kMaxMelGroups + 32 * g + l], 4);
                        for (int st = 0; st < meta[g]; ++st)
                            for (int c = 0; c < 4; ++c)
                                row[kTmMel + 4 * (pre + st) + c] = melw[(size_t)meta[kMaxMelGroups + g] + ((size_t)st * 32 + l) * 4 + c];
                        pre += meta[g];
                    }
                }
                UP(d_tmem_tab, tm)
                pl->ft.tmem_tab = pl->d_tmem_tab;
                pl->ft.tmem_cols = cols;
            }
        }
    }
    // register-FFT tables for n_fft = 4096 (frames_fast_4096): full-length periodic Hann, n_mels <= 128
    if (N == 4096) {
        bool hann = (p.win_length == N);
        for (int i = 0; window && hann && i < N; ++i)
            if (fabs(window[i] - (0.5 - 0.5 * cos(2.0 * M_PI * double(i) / double(N)))) > 1e-9) hann = false;
        if (hann && nm <= 128) {
            Fast4Tables ft{};
            ft.n_groups = (nm + 31) / 32;
            // filterbank columns de-interleaved: even bins 2k' (k' = 0..1024), odd bins 2k'+1 (k' = 0..1023)
            const int Fh = 1025, kReadEnd = 1105;
            std::vector<float> dense[2] = {std::vector<float>((size_t)nm * Fh, 0.0f), std::vector<float>((size_t)nm * Fh, 0.0f)};
            for (int m = 0; m < nm; ++m)
                for (int k = 0; k < F; ++k) dense[k & 1][(size_t)m * Fh + (k >> 1)] = pl->mel_dense[(size_t)m * F + k];
            std::vector<int> meta[2];
            std::vector<float> melw[2];
            for (int par = 0; par < 2; ++par) build_banded_groups(dense[par], nm, Fh, kReadEnd, meta[par], melw[par]);
            auto r4 = [](int x) { return (x + 3) & ~3; };
            ft.tw1 = 0;
            ft.tw0 = ft.tw1 + 31 * 32 * 2;
            ft.base = ft.tw0 + 1024 * 2;
            ft.hann_cs = ft.base + 2 * 32 * 2;
            int at = ft.hann_cs + 32 * 4;
            for (int par = 0; par < 2; ++par) {
                ft.mel_meta[par] = at;
                ft.mel_w[par] = r4(at + (int)meta[par].size());
                at = r4(ft.mel_w[par] + (int)melw[par].size());
            }
            ft.total = at;
            std::vector<float> blob(ft.total, 0.0f);
            for (int k1 = 1; k1 < 32; ++k1)
                for (int l = 0; l < 32; ++l) {
                    const double th = 2.0 * M_PI * double((l * k1) % 1024) / 1024.0;
                    blob[ft.tw1 + ((k1 - 1) * 32 + l) * 2 + 0] = float(cos(th));
                    blob[ft.tw1 + ((k1 - 1) * 32 + l) * 2 + 1] = float(-sin(th));
                }
            for (int m = 0; m < 1024; ++m) {
                const double th = 2.0 * M_PI * double(m) / 2048.0;               // W_2048^m
                blob[ft.tw0 + 2 * m + 0] = float(cos(th));
                blob[ft.tw0 + 2 * m + 1] = float(-sin(th));
            }
            for (int l = 0; l < 32; ++l) {
                const double t0 = 2.0 * M_PI * double(16 * l) / 2048.0;          // -i * W_2048^(16 l)
                const double t1 = 2.0 * M_PI * double(32 * l + 1) / 4096.0;      // -i * W_4096^(32 l + 1)
                blob[ft.base + 2 * l + 0] = float(-sin(t0));
                blob[ft.base + 2 * l + 1] = float(-cos(t0));
                blob[ft.base + 2 * (32 + l) + 0] = float(-sin(t1));
                blob[ft.base + 2 * (32 + l) + 1] = float(-cos(t1));
                const double te = 2.0 * M_PI * double(2 * l) / double(N), to = 2.0 * M_PI * double(2 * l + 1) / double(N);
                blob[ft.hann_cs + 4 * l + 0] = float(cos(te));
                blob[ft.hann_cs + 4 * l + 1] = float(cos(to));
                blob[ft.hann_cs + 4 * l + 2] = float(sin(te));
                blob[ft.hann_cs + 4 * l + 3] = float(sin(to));
            }
            for (int par = 0; par < 2; ++par) {
                memcpy(&blob[ft.mel_meta[par]], meta[par].data(), meta[par].size() * 4);
                if (!melw[par].empty()) memcpy(&blob[ft.mel_w[par]], melw[par].data(), melw[par].size() * 4);
            }
            pl->ft4 = ft;
            UP(d_fast4, blob)
            pl->fast4_ok = fast4_smem_bytes(ft) <= 227 * 1024;
            pl->fast_ok = pl->fast4_ok;
            {   // the same tables per lane for Tensor Memory (frames_fast_4096<TM>), see k4Tm* in hlmc_internal.h
                int total_steps = 0;
                for (int par = 0; par < 2; ++par)
                    for (int g = 0; g < ft.n_groups; ++g) total_steps += meta[par][g];
                const int cols = k4TmMel + 4 * total_steps;
                if (cols <= kTmAlloc) {
                    std::vector<float> tm((size_t)32 * cols, 0.0f);
                    int mel_col[2] = {0, 0};
                    for (int l = 0; l < 32; ++l) {
                        float* row = &tm[(size_t)l * cols];
                        for (int k1 = 1; k1 < 32; ++k1) {
                            row[k4TmTw1 + 2 * (k1 - 1)] = blob[ft.tw1 + ((k1 - 1) * 32 + l) * 2];
                            row[k4TmTw1 + 2 * (k1 - 1) + 1] = blob[ft.tw1 + ((k1 - 1) * 32 + l) * 2 + 1];
                        }
                        for (int j = 0; j < 32; ++j) {
                            row[k4TmTw0 + 2 * j] = blob[ft.tw0 + (l + 32 * j) * 2];
                            row[k4TmTw0 + 2 * j + 1] = blob[ft.tw0 + (l + 32 * j) * 2 + 1];
                        }
                        for (int par = 0; par < 2; ++par) {
                            row[k4TmBase + 2 * par] = blob[ft.base + (32 * par + l) * 2];
                            row[k4TmBase + 2 * par + 1] = blob[ft.base + (32 * par + l) * 2 + 1];
                        }
                        for (int c = 0; c < 4; ++c) row[k4TmHcs + c] = blob[ft.hann_cs + 4 * l + c];
                        int pre = 0;
                        for (int par = 0; par < 2; ++par) {
                            mel_col[par] = k4TmMel + 4 * pre;
                            for (int g = 0; g < ft.n_groups; ++g) {
                                memcpy(&row[k4TmMeta + 4 * par + g], &meta[par][2 * kMaxMelGroups + 32 * g + l], 4);
                                for (int st = 0; st < meta[par][g]; ++st)
                                    for (int c = 0; c < 4; ++c)
                                        row[k4TmMel + 4 * (pre + st) + c] =
                                            melw[par][(size_t)meta[par][kMaxMelGroups + g] + ((size_t)st * 32 + l) * 4 + c];
                                pre += meta[par][g];
                            }
                        }
                    }
                    UP(d_tmem_tab, tm)
                    pl->ft4.tmem_tab = pl->d_tmem_tab;
                    pl->ft4.tmem_cols = cols;
                    for (int par = 0; par < 2; ++par) {
                        pl->ft4.mel_col[par] = mel_col[par];
                        for (int g = 0; g < ft.n_groups; ++g) pl->ft4.mel_steps[4 * par + g] = meta[par][g];
                    }
                }
            }
        }
    }
    // register-FFT tables for n_fft = 1024 / 512: L = n_fft / 64 lanes per frame (frames_sub kernel)
    if (N == 1024 || N == 512) {
        const int Lg = N / 64, Mh = N / 2, PEND = 34 * Lg + 17;
        FastTables ft{};
        const int R = (nm + Lg - 1) / Lg;                 // mel rounds: one filter per lane of a group
        ft.n_groups = R;
        std::vector<int> meta(2 * R + R * Lg, 0);
        std::vector<float> melw;
        auto qof = [](int k) { return k + (k >> 4); };
        for (int r = 0; r < R; ++r) {
            std::vector<int> ql(Lg, 0), qlen(Lg, 0), shiftv(Lg, 0);
            int gmax = 0;
            for (int l = 0; l < Lg; ++l) {
                const int m = Lg * r + l;
                if (m < nm && len[m] > 0) {
                    ql[l] = qof(lo[m]);
                    qlen[l] = qof(lo[m] + len[m] - 1) - ql[l] + 1;
                }
                gmax = std::max(gmax, qlen[l]);
            }
            // as in the 2048 path: shift every lane's first tap down so that the Lg lanes of a group
            // start on Lg different banks modulo Lg (the groups of a warp are offset by 32 / (32/Lg))
            int G = (gmax + 3) & ~3;
            bool ok = (G == 0);
            for (int tries = 0; !ok && tries < 12; ++tries, G += 4) {
                std::vector<std::vector<int>> opt(Lg);
                bool feasible = true;
                for (int l = 0; l < Lg; ++l) {
                    if (qlen[l] == 0) { for (int s2 = 0; s2 < Lg; ++s2) opt[l].push_back(-s2); continue; }
                    const int smin = std::max(0, ql[l] + G - PEND);
                    const int smax = std::min(G - qlen[l], ql[l]);
                    if (smin > smax) { feasible = false; break; }
                    for (int s2 = smin; s2 <= smax && s2 < smin + Lg; ++s2) opt[l].push_back(s2);
                }
                if (!feasible) continue;
                std::vector<int> owner(Lg, -1);
                auto bank_of = [&](int l, int s2) { return ((qlen[l] == 0 ? -s2 : ql[l] - s2) % Lg + Lg) % Lg; };
                std::vector<char> seen(Lg);
                std::function<bool(int)> aug = [&](int l) -> bool {
                    for (int s2 : opt[l]) {
                        const int bk = bank_of(l, s2);
                        if (seen[bk]) continue;
                        seen[bk] = 1;
                        if (owner[bk] < 0 || aug(owner[bk])) { owner[bk] = l; return true; }
                    }
                    return false;
                };
                int matched = 0;
                for (int l = 0; l < Lg; ++l) { std::fill(seen.begin(), seen.end(), 0); if (aug(l)) ++matched; }
                if (matched == Lg) {
                    for (int bk = 0; bk < Lg; ++bk) {
                        const int l = owner[bk];
                        for (int s2 : opt[l]) if (bank_of(l, s2) == bk) { shiftv[l] = s2; break; }
                    }
                    ok = true;
                    break;
                }
            }
            if (!ok) {
                G = (gmax + 3) & ~3;
                for (int l = 0; l < Lg; ++l) shiftv[l] = std::max(0, ql[l] + G - PEND);
            }
            meta[r] = G / 4;
            meta[R + r] = (int)melw.size();
            const size_t base = melw.size();
            melw.resize(base + (size_t)Lg * G, 0.0f);
            for (int l = 0; l < Lg; ++l) {
                const int m = Lg * r + l;
                const int start = (qlen[l] == 0) ? -shiftv[l] : ql[l] - shiftv[l];
                meta[2 * R + r * Lg + l] = start;
                if (qlen[l] == 0) continue;
                for (int q = ql[l]; q < ql[l] + qlen[l]; ++q) {
                    if (q % 17 == 16) continue;
                    const int k = q - q / 17, i = q - start;
                    melw[base + ((size_t)(i / 4) * Lg + l) * 4 + (i % 4)] = pl->mel_dense[(size_t)m * F + k];
                }
            }
        }
        auto r4 = [](int x) { return (x + 3) & ~3; };
        ft.win = 0;
        ft.tw1 = ft.win + N;
        ft.tw2 = ft.tw1 + 31 * Lg * 2;
        ft.hann_cs = r4(ft.tw2 + 16 * Lg * 2);
        ft.mel_meta = ft.hann_cs + Lg * 4;
        ft.mel_w = r4(ft.mel_meta + (int)meta.size());
        ft.total = r4(ft.mel_w + (int)melw.size());
        ft.nowin = ft.total;
        ft.scr = 0;
        // full-length periodic Hann: frames_sub synthesises the window (its shared-memory pipe is 83 % busy)
        ft.hann = (p.win_length == N) ? 1 : 0;
        for (int i = 0; window && ft.hann && i < N; ++i)
            if (fabs(window[i] - (0.5 - 0.5 * cos(2.0 * M_PI * double(i) / double(N)))) > 1e-9) ft.hann = 0;
        std::vector<float> blob(ft.total, 0.0f);
        memcpy(&blob[ft.win], win.data(), N * 4);
        for (int l = 0; l < Lg; ++l) {
            const double te = 2.0 * M_PI * double(2 * l) / double(N), to = 2.0 * M_PI * double(2 * l + 1) / double(N);
            blob[ft.hann_cs + 4 * l + 0] = float(cos(te));
            blob[ft.hann_cs + 4 * l + 1] = float(cos(to));
            blob[ft.hann_cs + 4 * l + 2] = float(sin(te));
            blob[ft.hann_cs + 4 * l + 3] = float(sin(to));
        }
        for (int k1 = 1; k1 < 32; ++k1)
            for (int l = 0; l < Lg; ++l) {
                const double th = 2.0 * M_PI * double((l * k1) % Mh) / double(Mh);
                blob[ft.tw1 + ((k1 - 1) * Lg + l) * 2 + 0] = float(cos(th));
                blob[ft.tw1 + ((k1 - 1) * Lg + l) * 2 + 1] = float(-sin(th));
            }
        for (int i = 0; i < 16; ++i)
            for (int l = 0; l < Lg; ++l) {
                const double th = 2.0 * M_PI * double(16 * l + i) / double(N);
                blob[ft.tw2 + (i * Lg + l) * 2 + 0] = float(-sin(th));
                blob[ft.tw2 + (i * Lg + l) * 2 + 1] = float(-cos(th));
            }
        memcpy(&blob[ft.mel_meta], meta.data(), meta.size() * 4);
        if (!melw.empty()) memcpy(&blob[ft.mel_w], melw.data(), melw.size() * 4);
        for (int r = 0; r < R && r < 32; ++r) ft.mel_steps[r] = meta[r];
        pl->ft = ft;
        UP(d_fast, blob)
        pl->fast_ok = sub_smem_bytes(ft, 16, Lg) <= 227 * 1024;
        {   // the same tables per lane for Tensor Memory (frames_sub<..., TM>), see kSub* in hlmc_internal.h
            int total_steps = 0;
            for (int r = 0; r < R; ++r) total_steps += meta[r];
            const int cols = kSubMel + 4 * total_steps;
            if (R <= 32 && cols <= kTmAlloc) {
                std::vector<float> tm((size_t)32 * cols, 0.0f);
                for (int l = 0; l < 32; ++l) {
                    const int lg = l & (Lg - 1);
                    float* row = &tm[(size_t)l * cols];
                    for (int j = 0; j < 32; ++j) {
                        row[kSubWin + 2 * j] = win[2 * (lg + Lg * j)];
                        row[kSubWin + 2 * j + 1] = win[2 * (lg + Lg * j) + 1];
                    }
                    for (int k1 = 1; k1 < 32; ++k1) {
                        row[kSubTw1 + 2 * (k1 - 1)] = blob[ft.tw1 + ((k1 - 1) * Lg + lg) * 2];
                        row[kSubTw1 + 2 * (k1 - 1) + 1] = blob[ft.tw1 + ((k1 - 1) * Lg + lg) * 2 + 1];
                    }
                    for (int i = 0; i < 16; ++i) {
                        row[kSubTw2 + 2 * i] = blob[ft.tw2 + (i * Lg + lg) * 2];
                        row[kSubTw2 + 2 * i + 1] = blob[ft.tw2 + (i * Lg + lg) * 2 + 1];
                    }
                    int pre = 0;
                    for (int r = 0; r < R; ++r) {
                        memcpy(&row[kSubMeta + r], &meta[2 * R + r * Lg + lg], 4);
                        for (int st = 0; st < meta[r]; ++st)
                            for (int c = 0; c < 4; ++c)
                                row[kSubMel + 4 * (pre + st) + c] = melw[(size_t)meta[R + r] + ((size_t)st * Lg + lg) * 4 + c];
                        pre += meta[r];
                    }
                }
                UP(d_tmem_tab, tm)
                pl->ft.tmem_tab = pl->d_tmem_tab;
                pl->ft.tmem_cols = cols;
            }
        }
    }
#undef UP
    *out = pl;
    return HLMC_OK;
}

int hlmc_plan_set_path(hlmc_plan* plan, int generic) {
    if (!plan) return fail(HLMC_ERR_PARAM, "null plan");
    plan->force_generic = (generic == HLMC_PATH_GENERIC) ? 1 : 0;
    plan->no_tmem = (generic == HLMC_PATH_FAST_SMEM_TABLES) ? 1 : 0;
    return HLMC_OK;
}
int hlmc_plan_uses_fast_path(const hlmc_plan* plan) {
    return (plan && plan->fast_ok && !plan->force_generic) ? 1 : 0;
}

int hlmc_plan_mel_basis(const hlmc_plan* plan, float* h_out) {
    if (!plan || !h_out) return fail(HLMC_ERR_PARAM, "null argument");
    memcpy(h_out, plan->mel_dense.data(), plan->mel_dense.size() * 4);
    return HLMC_OK;
}
int hlmc_plan_dct_basis(const hlmc_plan* plan, float* h_out) {
    if (!plan || !h_out) return fail(HLMC_ERR_PARAM, "null argument");
    if (!plan->dct.empty()) memcpy(h_out, plan->dct.data(), plan->dct.size() * 4);
    return HLMC_OK;
}

static FrameArgs make_frame_args(const hlmc_plan* pl, const float* d_wave, int64_t B, int64_t n, int64_t pitch, int T) {
    FrameArgs a{};
    a.wave = d_wave; a.pitch = pitch; a.B = (int)B; a.n = (int)n; a.T = T;
    a.n_fft = pl->p.n_fft; a.hop = pl->p.hop_length; a.pad = pl->p.center ? pl->p.n_fft / 2 : 0;
    a.pad_mode = pl->p.pad_mode; a.n_mels = pl->p.n_mels; a.use_mag = (pl->p.power == 1.0f) ? 1 : 0;
    a.binhz = float(double(pl->p.sr) / double(pl->p.n_fft));
    a.roll_percent = pl->p.roll_percent; a.zcr_thr = pl->p.zcr_threshold;
    a.pip_klo = pl->pip_klo; a.pip_khi = pl->pip_khi; a.pip_threshold = 0.1f;
    a.no_tmem = pl->no_tmem;
    return a;
}

static int run_frames(hlmc_plan* pl, const float* d_wave, int64_t B, int64_t n, int64_t pitch, int T,
                      float* d_mel, float* d_stats, int32_t* d_status, float* d_clipmax, float* d_spec,
                      cudaStream_t st, float2* cand = nullptr, int* cand_count = nullptr, int cand_cap = 0,
                      int mel_frame_major = 0, float* pstash = nullptr) {
    FrameArgs a = make_frame_args(pl, d_wave, B, n, pitch, T);
    a.mel_out = d_mel; a.stats = d_stats; a.status = d_status; a.mel_frame_major = mel_frame_major;
    a.clipmax = reinterpret_cast<unsigned int*>(d_clipmax); a.spec = d_spec;
    a.cand = cand; a.cand_count = cand_count; a.cand_cap = cand_cap; a.pstash = pstash;
    if (d_clipmax) CK(cudaMemsetAsync(d_clipmax, 0, (size_t)B * 4, st));
    if (d_status) CK(cudaMemsetAsync(d_status, 0, (size_t)B * 4, st));
    if (pl->fast_ok && !pl->force_generic && d_spec == nullptr) {
        if (pl->p.n_fft == kFastNfft) CK(launch_frames_fast(a, pl->d_fast, pl->ft, pl->num_sms, st));
        else if (pl->p.n_fft == 4096) {
            if (cand) return fail(HLMC_ERR_UNSUPPORTED, "chroma needs n_fft = 2048");
            CK(launch_frames_fast4096(a, pl->d_fast4, pl->ft4, pl->num_sms, st));
        }
        else CK(launch_frames_sub(a, pl->d_fast, pl->ft, pl->num_sms, st));
    } else {
        if (cand) return fail(HLMC_ERR_UNSUPPORTED, "chroma needs the register-FFT kernel");
        GenericTables gt{pl->d_win, pl->d_twm, pl->d_tws, pl->d_mel_lo, pl->d_mel_len, pl->d_mel_off, pl->d_mel_w};
        CK(launch_frames_generic(a, gt, st));
    }
    return HLMC_OK;
}

static int check_batch(const hlmc_plan* pl, const void* wave, int64_t B, int64_t n, int64_t pitch, int64_t* T) {
    if (!pl) return fail(HLMC_ERR_PARAM, "null plan");
    if (B < 0) return fail(HLMC_ERR_PARAM, "negative batch");
    if (B > 0 && !wave) return fail(HLMC_ERR_PARAM, "null waveform pointer");
    if (pitch < n) return fail(HLMC_ERR_PARAM, "pitch < n");
    if (B > 0x7fffffff) return fail(HLMC_ERR_UNSUPPORTED, "batch too large for one call");
    const int64_t t = hlmc_num_frames(&pl->p, n);
    if (t < 0) return (int)t;
    *T = t;
    return HLMC_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// candidates + counts + tuning indices: what the tuning estimate needs
static size_t chroma_ws_core(const hlmc_plan* plan, int64_t B, int64_t T) {
    return align256((size_t)B * T * 4) + align256((size_t)B * 4) +
           align256((size_t)B * T * plan->cand_per_frame * sizeof(float2));
}

// ... plus the power-spectrum stash (B, T, kStashFloats) that lets the projection skip a second STFT pass.  A
// caller that passes only the core size still gets chroma, through the recomputing kernel.
int64_t hlmc_chroma_workspace_bytes(hlmc_plan* plan, int64_t B, int64_t n) {
    if (!plan) return fail(HLMC_ERR_PARAM, "null plan");
    int rc = ensure_chroma_tables(plan);
    if (rc != HLMC_OK) return rc;
    const int64_t T = hlmc_num_frames(&plan->p, n);
    if (T < 0) return T;
    return (int64_t)(chroma_ws_core(plan, B, T) + (plan->chroma_generic ? 0 : (size_t)B * T * kStashFloats * 4));
}

// d_pooled with d_logmel == NULL selects the fused path: dB, DCT and time pooling in one kernel, the
// (B, n_mels, T) / (B, n_mfcc, T) arrays are never written.
struct ImplOwned {                   // what one extract_device call creates before it can fail
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    float* melscr = nullptr;         // stream-ordered scratch, if the caller passed none
    bool handed_over = false;        // events now belong to plan->ev
};

static int extract_device_body(hlmc_plan* plan, const float* d_wave, int64_t B, int64_t n, int64_t pitch,
                               float* d_logmel, float* d_mfcc, float* d_stats, int32_t* d_status,
                               float* d_clipmax, float* d_chroma, float* d_tuning, void* d_work,
                               int64_t work_bytes, void* stream, float* d_melscr,
                               float* d_pooled, int pool_mfcc, int pool_chroma, ImplOwned& own);

static int extract_device_impl(hlmc_plan* plan, const float* d_wave, int64_t B, int64_t n, int64_t pitch,
                               float* d_logmel, float* d_mfcc, float* d_stats, int32_t* d_status,
                               float* d_clipmax, float* d_chroma, float* d_tuning, void* d_work,
                               int64_t work_bytes, void* stream, float* d_melscr,
                               float* d_pooled = nullptr, int pool_mfcc = 0, int pool_chroma = 0) {
    ImplOwned own;
    const int rc = extract_device_body(plan, d_wave, B, n, pitch, d_logmel, d_mfcc, d_stats, d_status, d_clipmax,
                                       d_chroma, d_tuning, d_work, work_bytes, stream, d_melscr, d_pooled, pool_mfcc,
                                       pool_chroma, own);
    if (rc != HLMC_OK) {             // nothing of a failed call outlives it
        const std::string msg = g_err;
        if (own.melscr) cudaFreeAsync(own.melscr, static_cast<cudaStream_t>(stream));
        if (!own.handed_over)
            for (auto e : own.ev) if (e) cudaEventDestroy(e);
        g_err = msg;
    }
    return rc;
}

static int extract_device_body(hlmc_plan* plan, const float* d_wave, int64_t B, int64_t n, int64_t pitch,
                               float* d_logmel, float* d_mfcc, float* d_stats, int32_t* d_status,
                               float* d_clipmax, float* d_chroma, float* d_tuning, void* d_work,
                               int64_t work_bytes, void* stream, float* d_melscr,
                               float* d_pooled, int pool_mfcc, int pool_chroma, ImplOwned& own) {
    int64_t T;
    int rc = check_batch(plan, d_wave, B, n, pitch, &T);
    if (rc != HLMC_OK) return rc;
    if (B == 0) return HLMC_OK;
    const bool fused = (d_pooled != nullptr && d_logmel == nullptr);
    if ((!fused && !d_logmel) || !d_clipmax) return fail(HLMC_ERR_PARAM, "d_logmel and d_clipmax are required");
    if (fused && !d_stats) return fail(HLMC_ERR_PARAM, "the pooled path needs a (B, 5, T) d_stats buffer");
    if (fused && !db_pool_fits(plan->p.n_mels, pool_mfcc ? plan->ncp : 0, (int)T))
        return fail(HLMC_ERR_UNSUPPORTED, "n_mels / n_mfcc too large for the fused pooling kernel: use hlmc_extract_device + hlmc_pool_device");
    if (fused && pool_chroma && !d_chroma) return fail(HLMC_ERR_PARAM, "pooled chroma columns need d_chroma");
    if ((d_mfcc || (fused && pool_mfcc)) && plan->p.n_mfcc <= 0) return fail(HLMC_ERR_PARAM, "plan was created with n_mfcc = 0");
    CK(cudaSetDevice(plan->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // chroma: piptrack candidates come out of the same pass as the other features
    float2* cand = nullptr; int* cand_count = nullptr; int* tuning_idx = nullptr; int cand_cap = 0;
    float* pstash = nullptr;
    if (d_chroma) {
        rc = ensure_chroma_tables(plan);
        if (rc != HLMC_OK) return rc;
        if (plan->force_generic && !plan->chroma_generic) return fail(HLMC_ERR_UNSUPPORTED, "chroma needs the register-FFT kernel");
        const size_t core = chroma_ws_core(plan, B, T);
        if (!d_work || work_bytes < (int64_t)core) return fail(HLMC_ERR_PARAM, "chroma workspace too small");
        char* wsp = static_cast<char*>(d_work);
        cand_count = reinterpret_cast<int*>(wsp);                                   // (B, T), every entry written
        tuning_idx = reinterpret_cast<int*>(wsp + align256((size_t)B * T * 4));
        cand = reinterpret_cast<float2*>(wsp + align256((size_t)B * T * 4) + align256((size_t)B * 4));
        cand_cap = plan->cand_per_frame;
        if (!plan->chroma_generic && work_bytes >= (int64_t)(core + (size_t)B * T * kStashFloats * 4))
            pstash = reinterpret_cast<float*>(wsp + core);
    }
    const bool cgen = d_chroma && plan->chroma_generic;
    cudaEvent_t (&ev3)[3] = own.ev;
    if (plan->timing) {
        if (plan->ev.size() >= 3 * 1024) { rc = fold_timing(plan); if (rc != HLMC_OK) return rc; }
        for (auto& e : ev3) CK(cudaEventCreate(&e));
        CK(cudaMemsetAsync(d_clipmax, 0, (size_t)B * 4, st));   // keep the memsets out of the bracket
    }
    if (plan->timing) CK(cudaEventRecord(ev3[0], st));
    // The frames kernel writes the mel power frame-major (each frame's n_mels values contiguous:
    // full-sector coalesced stores) into a scratch; db_dct reads it back and writes librosa's layout.
    if (!d_melscr) {
        CK(cudaMallocAsync((void**)&own.melscr, (size_t)B * T * plan->p.n_mels * 4, st));
        d_melscr = own.melscr;
    }
    rc = run_frames(plan, d_wave, B, n, pitch, (int)T, d_melscr, d_stats, d_status, d_clipmax, nullptr, st,
                    cgen ? nullptr : cand, cgen ? nullptr : cand_count, cgen ? 0 : cand_cap, 1, pstash);
    if (rc != HLMC_OK) return rc;
    if (plan->timing) CK(cudaEventRecord(ev3[1], st));
    DbArgs d{};
    d.mel = d_logmel; d.mel_in = d_melscr; d.mfcc = d_mfcc; d.clipmax = reinterpret_cast<const unsigned int*>(d_clipmax);
    d.dct_t = plan->d_dct_t; d.B = (int)B; d.n_mels = plan->p.n_mels; d.n_mfcc = plan->p.n_mfcc;
    d.ncp = plan->ncp; d.T = (int)T; d.ref_mode = plan->p.ref_mode; d.ref_value = plan->p.ref_value;
    d.amin = plan->p.amin; d.top_db = plan->p.top_db;
    auto run_chroma = [&]() -> int {
        if (cgen) {
            // any n_fft: candidates from one pass of the shared-memory FFT kernel, the projection from a second
            GenericTables gt{plan->d_win, plan->d_twm, plan->d_tws, plan->d_mel_lo, plan->d_mel_len, plan->d_mel_off, plan->d_mel_w};
            FrameArgs a = make_frame_args(plan, d_wave, B, n, pitch, (int)T);
            a.cand = cand; a.cand_count = cand_count; a.cand_cap = cand_cap;
            CK(launch_frames_generic(a, gt, st));
            CK(launch_tuning(cand, cand_count, (int)T, cand_cap, B, plan->d_edges, d_tuning, tuning_idx, st));
            FrameArgs a2 = make_frame_args(plan, d_wave, B, n, pitch, (int)T);
            a2.chroma_out = d_chroma; a2.chroma_fb = plan->d_chroma_fb; a2.chroma_tidx = tuning_idx;
            CK(launch_frames_generic(a2, gt, st));
            return HLMC_OK;
        }
        CK(launch_tuning(cand, cand_count, (int)T, cand_cap, B, plan->d_edges, d_tuning, tuning_idx, st));
        ChromaArgs ca{tuning_idx, plan->d_chroma_fb, d_chroma};
        if (pstash) {
            CK(launch_chroma_project(pstash, ca, B, (int)T, plan->num_sms, st));
        } else {                               // small workspace: second STFT pass
            FrameArgs a = make_frame_args(plan, d_wave, B, n, pitch, (int)T);
            CK(launch_chroma_fast(a, ca, plan->d_fast, plan->ft, plan->num_sms, st));
        }
        return HLMC_OK;
    };
    if (fused) {
        if (d_chroma) { rc = run_chroma(); if (rc != HLMC_OK) return rc; }     // its pooled columns come from (B, 12, T)
        PoolArgs pa{};
        pa.stats = d_stats; pa.chroma = pool_chroma ? d_chroma : nullptr; pa.n_chroma = pool_chroma ? kChroma : 0;
        pa.with_mfcc = pool_mfcc ? 1 : 0; pa.pooled = d_pooled;
        pa.pooled_w = 2 * d.n_mels + 2 * (pool_mfcc ? d.n_mfcc : 0) + 10 + 2 * pa.n_chroma;
        CK(launch_db_pool(d, pa, plan->num_sms, st));
    } else {
        CK(launch_db_dct(d, st));
    }
    if (plan->timing) {
        CK(cudaEventRecord(ev3[2], st));
        for (auto& e : ev3) plan->ev.push_back(e);
        own.handed_over = true;
    }
    if (own.melscr) { float* m = own.melscr; own.melscr = nullptr; CK(cudaFreeAsync(m, st)); }
    if (d_chroma && !fused) { rc = run_chroma(); if (rc != HLMC_OK) return rc; }
    return HLMC_OK;
}

int hlmc_extract_device_ex(hlmc_plan* plan, const float* d_wave, int64_t B, int64_t n, int64_t pitch,
                           float* d_logmel, float* d_mfcc, float* d_stats, int32_t* d_status,
                           float* d_clipmax, float* d_chroma, float* d_tuning, void* d_work,
                           int64_t work_bytes, void* stream) {
    return extract_device_impl(plan, d_wave, B, n, pitch, d_logmel, d_mfcc, d_stats, d_status, d_clipmax,
                               d_chroma, d_tuning, d_work, work_bytes, stream, nullptr);
}

int hlmc_extract_device(hlmc_plan* plan, const float* d_wave, int64_t B, int64_t n, int64_t pitch,
                        float* d_logmel, float* d_mfcc, float* d_stats, int32_t* d_status,
                        float* d_clipmax, void* stream) {
    return extract_device_impl(plan, d_wave, B, n, pitch, d_logmel, d_mfcc, d_stats, d_status, d_clipmax,
                               nullptr, nullptr, nullptr, 0, stream, nullptr);
}

int hlmc_extract_pooled_device(hlmc_plan* plan, const float* d_wave, int64_t B, int64_t n, int64_t pitch,
                               float* d_pooled, int with_mfcc, int with_chroma, float* d_stats, int32_t* d_status,
                               float* d_clipmax, float* d_chroma, float* d_tuning, void* d_work,
                               int64_t work_bytes, void* stream) {
    if (!d_pooled) return fail(HLMC_ERR_PARAM, "d_pooled is required");
    return extract_device_impl(plan, d_wave, B, n, pitch, nullptr, nullptr, d_stats, d_status, d_clipmax,
                               with_chroma ? d_chroma : nullptr, with_chroma ? d_tuning : nullptr, d_work, work_bytes,
                               stream, nullptr, d_pooled, with_mfcc, with_chroma);
}

struct hlmc_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    float* d_melscr = nullptr;
    int device = 0;
};

void hlmc_graph_destroy(hlmc_graph* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    cudaFree(g->d_melscr);
    delete g;
}

int hlmc_graph_create(hlmc_plan* plan, const float* d_wave, int64_t B, int64_t n, int64_t pitch,
                      float* d_logmel, float* d_mfcc, float* d_stats, int32_t* d_status, float* d_clipmax,
                      hlmc_graph** out) {
    if (!out) return fail(HLMC_ERR_PARAM, "null argument");
    *out = nullptr;
    int64_t T;
    int rc = check_batch(plan, d_wave, B, n, pitch, &T);
    if (rc != HLMC_OK) return rc;
    if (B == 0) return fail(HLMC_ERR_PARAM, "empty batch");
    if (plan->timing) return fail(HLMC_ERR_PARAM, "switch per-kernel timing off before capturing a graph");
    CK(cudaSetDevice(plan->device));
    hlmc_graph* g = new hlmc_graph();
    g->device = plan->device;
    cudaStream_t cs = nullptr;
    cudaError_t e = cudaMalloc((void**)&g->d_melscr, (size_t)B * T * plan->p.n_mels * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    if (e != cudaSuccess) { hlmc_graph_destroy(g); return cuda_fail(e, "graph setup"); }
    // one eager run first: function attributes and lazily built tables must exist before the capture
    rc = extract_device_impl(plan, d_wave, B, n, pitch, d_logmel, d_mfcc, d_stats, d_status, d_clipmax, nullptr,
                             nullptr, nullptr, 0, cs, g->d_melscr);
    if (rc == HLMC_OK && (e = cudaStreamSynchronize(cs)) != cudaSuccess) rc = cuda_fail(e, "graph warm-up");
    if (rc == HLMC_OK) {
        e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamBeginCapture");
        else {
            rc = extract_device_impl(plan, d_wave, B, n, pitch, d_logmel, d_mfcc, d_stats, d_status, d_clipmax,
                                     nullptr, nullptr, nullptr, 0, cs, g->d_melscr);
            e = cudaStreamEndCapture(cs, &g->graph);
            if (rc == HLMC_OK && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamEndCapture");
        }
    }
    if (rc == HLMC_OK && (e = cudaGraphInstantiate(&g->exec, g->graph, 0)) != cudaSuccess)
        rc = cuda_fail(e, "cudaGraphInstantiate");
    cudaStreamDestroy(cs);
    if (rc != HLMC_OK) { hlmc_graph_destroy(g); return rc; }
    *out = g;
    return HLMC_OK;
}

int hlmc_graph_launch(hlmc_graph* g, void* stream) {
    if (!g || !g->exec) return fail(HLMC_ERR_PARAM, "null graph");
    CK(cudaGraphLaunch(g->exec, static_cast<cudaStream_t>(stream)));
    return HLMC_OK;
}

int hlmc_plan_set_timing(hlmc_plan* plan, int enable) {
    if (!plan) return fail(HLMC_ERR_PARAM, "null plan");
    plan->timing = enable ? 1 : 0;
    return HLMC_OK;
}

int hlmc_plan_read_timing(hlmc_plan* plan, double* frames_ms, double* db_ms, int64_t* calls) {
    if (!plan) return fail(HLMC_ERR_PARAM, "null plan");
    int rc = fold_timing(plan);
    if (rc != HLMC_OK) return rc;
    const double f = plan->t_frames_ms, d = plan->t_db_ms;
    const int64_t n = plan->t_calls;
    plan->t_frames_ms = plan->t_db_ms = 0.0; plan->t_calls = 0;
    if (frames_ms) *frames_ms = f;
    if (db_ms) *db_ms = d;
    if (calls) *calls = n;
    return HLMC_OK;
}

int hlmc_melspectrogram_device(hlmc_plan* plan, const float* d_wave, int64_t B, int64_t n, int64_t pitch,
                               float* d_mel, float* d_stats, int32_t* d_status, void* stream) {
    int64_t T;
    int rc = check_batch(plan, d_wave, B, n, pitch, &T);
    if (rc != HLMC_OK) return rc;
    if (B == 0) return HLMC_OK;
    CK(cudaSetDevice(plan->device));
    return run_frames(plan, d_wave, B, n, pitch, (int)T, d_mel, d_stats, d_status, nullptr, nullptr,
                      static_cast<cudaStream_t>(stream));
}

int hlmc_stft_device(hlmc_plan* plan, const float* d_wave, int64_t B, int64_t n, int64_t pitch,
                     float* d_spec, void* stream) {
    int64_t T;
    int rc = check_batch(plan, d_wave, B, n, pitch, &T);
    if (rc != HLMC_OK) return rc;
    if (B == 0) return HLMC_OK;
    if (!d_spec) return fail(HLMC_ERR_PARAM, "null output");
    CK(cudaSetDevice(plan->device));
    return run_frames(plan, d_wave, B, n, pitch, (int)T, nullptr, nullptr, nullptr, nullptr, d_spec,
                      static_cast<cudaStream_t>(stream));
}

int hlmc_power_to_db_device(const float* d_in, float* d_out, int64_t B, int64_t rows, int64_t T,
                            int32_t ref_mode, float ref_value, float amin, float top_db,
                            float* d_clipmax, int device, void* stream) {
    if (!(amin > 0.0f)) return fail(HLMC_ERR_PARAM, "amin must be strictly positive");
    if (B < 0 || rows < 0 || T < 0) return fail(HLMC_ERR_PARAM, "negative shape");
    if (B == 0 || rows * T == 0) return HLMC_OK;
    if (!d_in || !d_out || !d_clipmax) return fail(HLMC_ERR_PARAM, "null argument");
    CK(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaMemsetAsync(d_clipmax, 0, (size_t)B * 4, st));
    CK(launch_rowmax(d_in, reinterpret_cast<unsigned int*>(d_clipmax), B, rows * T, st));
    CK(launch_power_to_db(d_in, d_out, reinterpret_cast<const unsigned int*>(d_clipmax), B, rows * T,
                          ref_mode, ref_value, amin, top_db, st));
    return HLMC_OK;
}

int hlmc_pool_device_ex(hlmc_plan* plan, const float* d_logmel, const float* d_mfcc, const float* d_stats,
                        const float* d_chroma, int64_t B, int64_t T, float* d_pooled, void* stream) {
    if (!plan) return fail(HLMC_ERR_PARAM, "null plan");
    if (B <= 0 || T <= 0) return HLMC_OK;
    if (!d_logmel || !d_stats || !d_pooled) return fail(HLMC_ERR_PARAM, "null argument");
    CK(cudaSetDevice(plan->device));
    CK(launch_pool(d_logmel, d_mfcc, d_stats, d_chroma, B, plan->p.n_mels, plan->p.n_mfcc, (int)T, d_pooled,
                   static_cast<cudaStream_t>(stream)));
    return HLMC_OK;
}

int hlmc_pool_device(hlmc_plan* plan, const float* d_logmel, const float* d_mfcc, const float* d_stats,
                     int64_t B, int64_t T, float* d_pooled, void* stream) {
    return hlmc_pool_device_ex(plan, d_logmel, d_mfcc, d_stats, nullptr, B, T, d_pooled, stream);
}

int hlmc_fix_frames_device(const float* d_in, float* d_out, int64_t B, int64_t rows, int64_t T,
                           int64_t fixed, int device, void* stream) {
    if (!d_in || !d_out) return fail(HLMC_ERR_PARAM, "null argument");
    if (B <= 0 || rows <= 0 || T <= 0 || fixed <= 0) return fail(HLMC_ERR_PARAM, "bad shape");
    CK(cudaSetDevice(device));
    CK(launch_fix_frames(d_in, d_out, B, (int)rows, (int)T, (int)fixed, static_cast<cudaStream_t>(stream)));
    return HLMC_OK;
}

// ---------------------------------------------------------------------------
// Host pipeline: chunk the batch; H2D | kernels | D2H overlap across streams.
// ---------------------------------------------------------------------------
static int alloc_slots(hlmc_plan* pl, int n_streams, int64_t chunk, int64_t n, int64_t T) {
    pl->slots.assign(n_streams, HostPipeSlot());
    const int64_t dp = (n + 3) & ~int64_t(3);
    const int nm = pl->p.n_mels, nc = pl->p.n_mfcc;
    for (auto& s : pl->slots) {
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CK(cudaMalloc((void**)&s.d_wave, (size_t)chunk * dp * 4));
        CK(cudaMalloc((void**)&s.d_logmel, (size_t)chunk * nm * T * 4));
        CK(cudaMalloc((void**)&s.d_melscr, (size_t)chunk * nm * T * 4));
        if (nc > 0) CK(cudaMalloc((void**)&s.d_mfcc, (size_t)chunk * nc * T * 4));
        CK(cudaMalloc((void**)&s.d_stats, (size_t)chunk * 5 * T * 4));
        CK(cudaMalloc((void**)&s.d_pooled, (size_t)chunk * (2 * nm + 2 * nc + 10) * 4));
        CK(cudaMalloc((void**)&s.d_clipmax, (size_t)chunk * 4));
        CK(cudaMalloc((void**)&s.d_status, (size_t)chunk * 4));
    }
    return HLMC_OK;
}

static int ensure_slots(hlmc_plan* pl, int n_streams, int64_t chunk, int64_t n, int64_t T, int flags) {
    if ((int)pl->slots.size() == n_streams && pl->slot_chunk >= chunk && pl->slot_n == n &&
        (pl->slot_flags & flags) == flags)
        return HLMC_OK;
    free_slots(pl);                 // also marks the cache empty: a failed allocation below can never be
                                    // mistaken for a valid configuration by a later call
    const int rc = alloc_slots(pl, n_streams, chunk, n, T);
    if (rc != HLMC_OK) {
        const std::string msg = g_err;
        free_slots(pl);
        cudaGetLastError();
        g_err = msg;
        return rc;
    }
    pl->slot_chunk = chunk; pl->slot_n = n; pl->slot_flags = flags;
    return HLMC_OK;
}

static int host_io_body(hlmc_plan* plan, const hlmc_host_io* io);

// ---------------------------------------------------------------------------
// scipy.signal.resample_poly's filter (librosa.resample(res_type="polyphase")), built in float64:
//   h = firwin(2*half_len + 1, 1/max(up, down), window=("kaiser", 5.0)), half_len = 10*max(up, down),
//   cast to float32, times up.
// ---------------------------------------------------------------------------
static double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = 0.25 * x * x;
    for (int k = 1; k < 200; ++k) {
        term *= q / (double(k) * double(k));
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}
static int64_t gcd64(int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; }

static void build_resample_filter(int sr_in, int sr_out, ResampleTaps& r) {
    const int64_t g = gcd64(sr_in, sr_out);
    r.sr_in = sr_in; r.up = (int)(sr_out / g); r.down = (int)(sr_in / g);
    const int max_rate = std::max(r.up, r.down);
    const double f_c = 1.0 / max_rate;
    r.half_len = 10 * max_rate;
    const int numtaps = 2 * r.half_len + 1;
    std::vector<double> h(numtaps);
    const double alpha = 0.5 * (numtaps - 1), beta = 5.0, i0b = bessel_i0(beta);
    long double sum = 0.0L;
    for (int k = 0; k < numtaps; ++k) {
        const double m = k - alpha;
        const double x = f_c * m;
        const double sinc = (x == 0.0) ? 1.0 : sin(M_PI * x) / (M_PI * x);
        const double rel = (k - alpha) / alpha;
        const double win = bessel_i0(beta * sqrt(std::max(0.0, 1.0 - rel * rel))) / i0b;
        h[k] = f_c * sinc * win;
        sum += h[k];
    }
    r.h.resize(numtaps);
    for (int k = 0; k < numtaps; ++k) {
        const float hf = (float)(h[k] / (double)sum);        // scipy: float64 design cast to x's dtype ...
        r.h[k] = hf * (float)r.up;                           // ... then h *= up in that dtype
    }
    r.n_pre_pad = r.down - r.half_len % r.down;
    r.n_pre_remove = (r.half_len + r.n_pre_pad) / r.down;
    r.tpp = (numtaps + r.up - 1) / r.up;
}

static int get_resampler(hlmc_plan* pl, int sr_in, const ResampleTaps** out) {
    for (auto& r : pl->resamplers) if (r.sr_in == sr_in) { *out = &r; return HLMC_OK; }
    if (sr_in <= 0 || sr_in > 768000) return fail(HLMC_ERR_PARAM, "bad input sample rate");
    ResampleTaps r;
    build_resample_filter(sr_in, pl->p.sr, r);
    if ((int64_t)r.up * r.tpp > (int64_t(1) << 24)) return fail(HLMC_ERR_UNSUPPORTED, "resampling ratio too fine");
    CK(cudaSetDevice(pl->device));
    std::vector<float> poly((size_t)r.up * r.tpp, 0.0f);
    for (int p = 0; p < r.up; ++p)
        for (int t = 0; t < r.tpp; ++t) {
            const int64_t q = p + (int64_t)t * r.up;
            if (q < (int64_t)r.h.size()) poly[(size_t)p * r.tpp + t] = r.h[q];
        }
    CK(cudaMalloc((void**)&r.d_hpoly, poly.size() * 4));
    CK(cudaMemcpy(r.d_hpoly, poly.data(), poly.size() * 4, cudaMemcpyHostToDevice));
    pl->resamplers.push_back(r);
    *out = &pl->resamplers.back();
    return HLMC_OK;
}

static int64_t resampled_length(int64_t n_in, int64_t sr_in, int64_t sr_out) {
    if (sr_in <= 0 || sr_in == sr_out) return n_in;
    const int64_t g = gcd64(sr_in, sr_out);
    const int64_t up = sr_out / g, down = sr_in / g;
    return (n_in * up + down - 1) / down;          // resample_poly's n_out == librosa's ceil(n * ratio)
}

static int run_frontend(hlmc_plan* plan, const void* d_raw, int fmt, int channels, int64_t B, int64_t n_in,
                        int64_t raw_pitch, int sr_in, float* d_wave, int64_t pitch, int64_t n_total,
                        cudaStream_t st, const long long* d_valid = nullptr) {
    FrontArgs a{};
    a.valid = d_valid;
    a.raw = d_raw; a.fmt = fmt; a.channels = channels < 1 ? 1 : channels; a.B = B; a.raw_pitch = raw_pitch;
    a.n_in = n_in; a.out = d_wave; a.pitch = pitch; a.n_total = n_total; a.up = a.down = 1;
    a.n_out = n_in;
    if (sr_in > 0 && sr_in != plan->p.sr) {
        const ResampleTaps* r = nullptr;
        const int rc = get_resampler(plan, sr_in, &r);
        if (rc != HLMC_OK) return rc;
        a.up = r->up; a.down = r->down; a.n_pre_pad = r->n_pre_pad; a.n_pre_remove = r->n_pre_remove;
        a.tpp = r->tpp; a.hpoly = r->d_hpoly;
        a.n_out = resampled_length(n_in, sr_in, plan->p.sr);
    }
    if (a.n_out > n_total) a.n_out = n_total;
    CK(launch_frontend(a, plan->num_sms, st));
    return HLMC_OK;
}

static int host_io_body(hlmc_plan* plan, const hlmc_host_io* io) {
    const int sample_format = io->sample_format;
    const int channels = io->channels > 1 ? io->channels : 1;
    const int sr_in = (io->sr_in > 0 && io->sr_in != plan->p.sr) ? io->sr_in : 0;
    const int64_t B = io->B, n_valid = io->n_valid, h_pitch = io->pitch;
    if (sample_format != HLMC_SAMPLES_F32 && sample_format != HLMC_SAMPLES_PCM16)
        return fail(HLMC_ERR_PARAM, "unknown sample_format");
    if (channels > 64) return fail(HLMC_ERR_PARAM, "more than 64 channels");
    if (n_valid < 1) return fail(HLMC_ERR_PARAM, "Input is too short");
    const int64_t n_res = resampled_length(n_valid, sr_in, plan->p.sr);
    const int64_t n_total = io->n_total > 0 ? io->n_total : n_res;
    int64_t chunk_clips = io->chunk_clips;
    int n_streams = io->n_streams;
    if (n_total < n_res) return fail(HLMC_ERR_PARAM, "n_total is shorter than the (resampled) clips");
    int64_t T;
    const int64_t n = n_total;
    if (h_pitch < n_valid) return fail(HLMC_ERR_PARAM, "pitch < n");
    int rc = check_batch(plan, io->wave, B, n, n, &T);
    if (rc != HLMC_OK) return rc;
    plan->last_h2d = plan->last_d2h = 0;
    if (B == 0) return HLMC_OK;
    if (io->mfcc && plan->p.n_mfcc <= 0) return fail(HLMC_ERR_PARAM, "plan was created with n_mfcc = 0");
    if (io->fixed_logmel && io->fixed_frames <= 0) return fail(HLMC_ERR_PARAM, "fixed_frames must be positive");
    const bool want_chroma = io->chroma || io->tuning || (io->pooled && io->pooled_with_chroma);
    if (want_chroma) {
        rc = ensure_chroma_tables(plan);
        if (rc != HLMC_OK) return rc;
    }
    CK(cudaSetDevice(plan->device));
    if (n_streams <= 0) n_streams = 3;
    if (n_streams > 8) n_streams = 8;
    if (chunk_clips <= 0) {
        chunk_clips = (int64_t(64) << 20) / (n * 4);
        if (chunk_clips < 1) chunk_clips = 1;
    }
    if (chunk_clips > 65535) chunk_clips = 65535;           // gridDim.y of the per-clip helper kernels
    if (chunk_clips > B) chunk_clips = B;
    rc = ensure_slots(plan, n_streams, chunk_clips, n, T, 0);
    if (rc != HLMC_OK) return rc;
    const bool pcm = (sample_format == HLMC_SAMPLES_PCM16);
    const size_t esz = pcm ? 2 : 4;
    const bool front = (channels > 1) || (sr_in != 0) || (io->valid_frames != nullptr);   // front-end kernel needed
    // device row pitch: every row 16-byte aligned - or, for float32 rows of even length that need no pad, the
    // host layout itself (rows 8-byte aligned, which the TMA staging handles), so that the H2D copy is one
    // linear transfer instead of a pitched 2-D one
    const bool linear = !front && !pcm && (n % 2 == 0) && (n_valid == n) && (h_pitch == n);
    const int64_t dp = linear ? n : ((n + 3) & ~int64_t(3));
    const int64_t rp = (n_valid + 7) & ~int64_t(7);           // staging pitch in frames (16-byte rows)
    const size_t frame_bytes = (size_t)channels * esz;
    int64_t chroma_ws = 0;
    for (auto& s : plan->slots) {
        const size_t need_raw = (pcm || front) ? (size_t)plan->slot_chunk * rp * frame_bytes : 0;
        if (s.raw_bytes < need_raw) {
            cudaFree(s.d_raw);
            s.d_raw = nullptr; s.raw_bytes = 0;
            CK(cudaMalloc(&s.d_raw, need_raw));
            s.raw_bytes = need_raw;
        }
        if (io->valid_frames && s.valid_elems < (size_t)plan->slot_chunk) {
            cudaFree(s.d_valid);
            s.d_valid = nullptr; s.valid_elems = 0;
            CK(cudaMalloc((void**)&s.d_valid, (size_t)plan->slot_chunk * 8));
            s.valid_elems = (size_t)plan->slot_chunk;
        }
        const size_t need_fixed = io->fixed_logmel ? (size_t)plan->slot_chunk * plan->p.n_mels * io->fixed_frames : 0;
        if (s.fixed_elems < need_fixed) {
            cudaFree(s.d_fixed);
            s.d_fixed = nullptr; s.fixed_elems = 0;
            CK(cudaMalloc((void**)&s.d_fixed, need_fixed * 4));
            s.fixed_elems = need_fixed;
        }
        if (want_chroma) {
            chroma_ws = hlmc_chroma_workspace_bytes(plan, plan->slot_chunk, n);
            if (chroma_ws < 0) return (int)chroma_ws;
            if (s.chroma_ws < (size_t)chroma_ws) {
                cudaFree(s.d_chroma); cudaFree(s.d_tuning); cudaFree(s.d_cwork); cudaFree(s.d_pooled);
                s.d_chroma = nullptr; s.d_tuning = nullptr; s.d_cwork = nullptr; s.d_pooled = nullptr;
                s.chroma_ws = 0;
                CK(cudaMalloc((void**)&s.d_chroma, (size_t)plan->slot_chunk * kChroma * T * 4));
                CK(cudaMalloc((void**)&s.d_tuning, (size_t)plan->slot_chunk * 4));
                CK(cudaMalloc((void**)&s.d_cwork, (size_t)chroma_ws));
                CK(cudaMalloc((void**)&s.d_pooled, (size_t)plan->slot_chunk *
                                                       (2 * plan->p.n_mels + 2 * plan->p.n_mfcc + 10 + 2 * kChroma) * 4));
                s.chroma_ws = (size_t)chroma_ws;
            }
        }
    }
    const int nm = plan->p.n_mels, nc = plan->p.n_mfcc;
    const bool want_mfcc = (io->mfcc != nullptr) || (io->pooled != nullptr && nc > 0);
    const bool pool_chroma = io->pooled && io->pooled_with_chroma;
    const int pooled_w = 2 * nm + 2 * (want_mfcc ? nc : 0) + 10 + (pool_chroma ? 2 * kChroma : 0);
    const int64_t fixed = io->fixed_frames;
    int64_t done = 0;
    for (int64_t i = 0; done < B; ++i) {
        HostPipeSlot& s = plan->slots[i % n_streams];
        const int64_t c = (B - done < chunk_clips) ? (B - done) : chunk_clips;
        const char* src = static_cast<const char*>(io->wave) + (size_t)done * h_pitch * frame_bytes;
        if (front) {
            // [R] librosa.load: decode -> to_mono -> resample; then the scripts' zero pad, all on the device
            CK(cudaMemcpy2DAsync(s.d_raw, (size_t)rp * frame_bytes, src, (size_t)h_pitch * frame_bytes,
                                 (size_t)n_valid * frame_bytes, (size_t)c, cudaMemcpyHostToDevice, s.stream));
            if (io->valid_frames)
                CK(cudaMemcpyAsync(s.d_valid, io->valid_frames + done, (size_t)c * 8, cudaMemcpyHostToDevice, s.stream));
            rc = run_frontend(plan, s.d_raw, pcm ? 1 : 0, channels, c, n_valid, rp, sr_in, s.d_wave, dp, n, s.stream,
                              io->valid_frames ? s.d_valid : nullptr);
            if (rc != HLMC_OK) return rc;
        } else if (pcm) {
            // [R] librosa.load on a PCM16 file: float32 = int16 / 32768; then the scripts' zero pad
            CK(cudaMemcpy2DAsync(s.d_raw, (size_t)rp * 2, src, (size_t)h_pitch * 2, (size_t)n_valid * 2,
                                 (size_t)c, cudaMemcpyHostToDevice, s.stream));
            CK(launch_pcm16_to_f32(static_cast<const int16_t*>(s.d_raw), rp, s.d_wave, dp, c, n_valid, n, s.stream));
        } else if (linear) {
            CK(cudaMemcpyAsync(s.d_wave, src, (size_t)c * n * 4, cudaMemcpyHostToDevice, s.stream));
        } else {
            CK(cudaMemcpy2DAsync(s.d_wave, (size_t)dp * 4, src, (size_t)h_pitch * 4, (size_t)n_valid * 4,
                                 (size_t)c, cudaMemcpyHostToDevice, s.stream));
            if (n > n_valid)   // [R] np.pad(audio, (0, expected - len(audio))) done on the device
                CK(cudaMemset2DAsync(s.d_wave + n_valid, (size_t)dp * 4, 0, (size_t)(n - n_valid) * 4,
                                     (size_t)c, s.stream));
        }
        plan->last_h2d += c * n_valid * (int64_t)frame_bytes;
        // only pooled columns wanted: dB + DCT + pooling fused, no log-mel / MFCC arrays in HBM
        const bool fused = io->pooled && !io->logmel && !io->mfcc && !io->fixed_logmel &&
                           db_pool_fits(nm, want_mfcc ? plan->ncp : 0, (int)T);
        rc = extract_device_impl(plan, s.d_wave, c, n, dp, fused ? nullptr : s.d_logmel,
                                 (want_mfcc && !fused) ? s.d_mfcc : nullptr,
                                 s.d_stats, s.d_status, s.d_clipmax, want_chroma ? s.d_chroma : nullptr,
                                 want_chroma ? s.d_tuning : nullptr, want_chroma ? s.d_cwork : nullptr,
                                 chroma_ws, s.stream, s.d_melscr, fused ? s.d_pooled : nullptr, want_mfcc ? 1 : 0,
                                 pool_chroma ? 1 : 0);
        if (rc != HLMC_OK) return rc;
        auto d2h = [&](void* dst, const void* srcd, size_t bytes) -> cudaError_t {
            plan->last_d2h += (int64_t)bytes;
            return cudaMemcpyAsync(dst, srcd, bytes, cudaMemcpyDeviceToHost, s.stream);
        };
        if (io->pooled) {
            if (!fused)
                CK(launch_pool(s.d_logmel, want_mfcc ? s.d_mfcc : nullptr, s.d_stats, pool_chroma ? s.d_chroma : nullptr,
                               c, nm, nc, (int)T, s.d_pooled, s.stream));
            CK(d2h(io->pooled + done * pooled_w, s.d_pooled, (size_t)c * pooled_w * 4));
        }
        if (io->fixed_logmel) {
            // [R] src/1_preprocessing_advanced.py:108-112: crop to fixed frames / pad with the clip's minimum
            CK(launch_fix_frames(s.d_logmel, s.d_fixed, c, nm, (int)T, (int)fixed, s.stream));
            CK(d2h(io->fixed_logmel + done * nm * fixed, s.d_fixed, (size_t)c * nm * fixed * 4));
        }
        if (io->logmel) CK(d2h(io->logmel + done * nm * T, s.d_logmel, (size_t)c * nm * T * 4));
        if (io->mfcc) CK(d2h(io->mfcc + done * nc * T, s.d_mfcc, (size_t)c * nc * T * 4));
        if (io->stats) CK(d2h(io->stats + done * 5 * T, s.d_stats, (size_t)c * 5 * T * 4));
        if (io->chroma) CK(d2h(io->chroma + done * kChroma * T, s.d_chroma, (size_t)c * kChroma * T * 4));
        if (io->tuning) CK(d2h(io->tuning + done, s.d_tuning, (size_t)c * 4));
        if (io->status) CK(d2h(io->status + done, s.d_status, (size_t)c * 4));
        if (io->wave_out) {
            plan->last_d2h += c * n * 4;
            CK(cudaMemcpy2DAsync(io->wave_out + done * n, (size_t)n * 4, s.d_wave, (size_t)dp * 4, (size_t)n * 4,
                                 (size_t)c, cudaMemcpyDeviceToHost, s.stream));
        }
        done += c;
    }
    for (auto& s : plan->slots) CK(cudaStreamSynchronize(s.stream));
    return HLMC_OK;
}

int hlmc_extract_host_io(hlmc_plan* plan, const hlmc_host_io* io) {
    if (!plan || !io) return fail(HLMC_ERR_PARAM, "null argument");
    const int rc = host_io_body(plan, io);
    if (rc != HLMC_OK) {
        // work may already be queued (async copies into the caller's buffers included): drain it before
        // the caller sees the error and frees or reuses them
        const std::string msg = g_err;
        for (auto& s : plan->slots) if (s.stream) cudaStreamSynchronize(s.stream);
        cudaGetLastError();
        g_err = msg;
    }
    return rc;
}

int hlmc_extract_host_ex(hlmc_plan* plan, const void* h_wave, int sample_format, int64_t B,
                         int64_t n_valid, int64_t h_pitch, int64_t n_total, float* h_logmel, float* h_mfcc,
                         float* h_stats, int32_t* h_status, float* h_pooled, int64_t chunk_clips,
                         int n_streams) {
    hlmc_host_io io;
    memset(&io, 0, sizeof(io));
    io.wave = h_wave; io.sample_format = sample_format; io.B = B; io.n_valid = n_valid; io.pitch = h_pitch;
    io.n_total = n_total; io.logmel = h_logmel; io.mfcc = h_mfcc; io.stats = h_stats; io.status = h_status;
    io.pooled = h_pooled; io.chunk_clips = chunk_clips; io.n_streams = n_streams;
    return hlmc_extract_host_io(plan, &io);
}

int hlmc_extract_host(hlmc_plan* plan, const float* h_wave, int64_t B, int64_t n, int64_t h_pitch,
                      float* h_logmel, float* h_mfcc, float* h_stats, int32_t* h_status, float* h_pooled,
                      int64_t chunk_clips, int n_streams) {
    return hlmc_extract_host_ex(plan, h_wave, HLMC_SAMPLES_F32, B, n, h_pitch, n, h_logmel, h_mfcc, h_stats,
                                h_status, h_pooled, chunk_clips, n_streams);
}

int hlmc_column_stats_device(const float* d_x, int64_t N, int64_t D, double* d_mean, double* d_m2,
                             int device, void* stream) {
    if (!d_x || !d_mean || !d_m2) return fail(HLMC_ERR_PARAM, "null argument");
    if (N < 0 || D < 0) return fail(HLMC_ERR_PARAM, "negative shape");
    CK(cudaSetDevice(device));
    CK(launch_colstats(d_x, N, D, d_mean, d_m2, static_cast<cudaStream_t>(stream)));
    return HLMC_OK;
}

int hlmc_standardize_device(const float* d_x, float* d_y, int64_t N, int64_t D, const float* d_mean,
                            const float* d_scale, int device, void* stream) {
    if (!d_x || !d_y || !d_mean || !d_scale) return fail(HLMC_ERR_PARAM, "null argument");
    if (N < 0 || D < 0) return fail(HLMC_ERR_PARAM, "negative shape");
    CK(cudaSetDevice(device));
    CK(launch_standardize(d_x, d_y, N, D, d_mean, d_scale, static_cast<cudaStream_t>(stream)));
    return HLMC_OK;
}

int64_t hlmc_resampled_length(int64_t n_in, int32_t sr_in, int32_t sr_out) {
    if (n_in < 0 || sr_in < 0 || sr_out <= 0) return fail(HLMC_ERR_PARAM, "bad argument");
    return resampled_length(n_in, sr_in, sr_out);
}

int hlmc_load_frontend_device(hlmc_plan* plan, const void* d_raw, int sample_format, int channels, int64_t B,
                              int64_t n_in, int64_t raw_pitch, int32_t sr_in, float* d_wave, int64_t pitch,
                              int64_t n_total, const int64_t* d_valid_frames, void* stream) {
    if (!plan || !d_raw || !d_wave) return fail(HLMC_ERR_PARAM, "null argument");
    if (sample_format != HLMC_SAMPLES_F32 && sample_format != HLMC_SAMPLES_PCM16)
        return fail(HLMC_ERR_PARAM, "unknown sample_format");
    if (channels < 1 || channels > 64) return fail(HLMC_ERR_PARAM, "bad channel count");
    if (B < 0 || n_in < 1 || raw_pitch < n_in || pitch < n_total) return fail(HLMC_ERR_PARAM, "bad shape");
    const int sr = (sr_in > 0 && sr_in != plan->p.sr) ? sr_in : 0;
    if (n_total < resampled_length(n_in, sr, plan->p.sr))
        return fail(HLMC_ERR_PARAM, "n_total is shorter than the (resampled) clips");
    if (channels == 2 && ((reinterpret_cast<uintptr_t>(d_raw) | (size_t)raw_pitch * 2 * (sample_format ? 2 : 4)) &
                          (sample_format ? 3 : 7)))
        return fail(HLMC_ERR_PARAM, "stereo rows must be aligned to one frame");
    if (B == 0) return HLMC_OK;
    CK(cudaSetDevice(plan->device));
    return run_frontend(plan, d_raw, sample_format == HLMC_SAMPLES_PCM16 ? 1 : 0, channels, B, n_in, raw_pitch, sr,
                        d_wave, pitch, n_total, static_cast<cudaStream_t>(stream),
                        reinterpret_cast<const long long*>(d_valid_frames));
}

int64_t hlmc_resample_taps(int32_t sr_in, int32_t sr_out, float* h_out, int64_t cap) {
    if (sr_in <= 0 || sr_out <= 0 || sr_in > 768000 || sr_out > 768000) return fail(HLMC_ERR_PARAM, "bad sample rate");
    ResampleTaps r;
    build_resample_filter(sr_in, sr_out, r);
    const int64_t nt = (int64_t)r.h.size();
    if (h_out) memcpy(h_out, r.h.data(), (size_t)std::min(nt, cap < 0 ? int64_t(0) : cap) * 4);
    return nt;
}

int hlmc_impute_stats_device(const double* d_x, int64_t N, int64_t D, double* d_sum, int64_t* d_count,
                             int device, void* stream) {
    if (!d_sum || !d_count || (N > 0 && D > 0 && !d_x)) return fail(HLMC_ERR_PARAM, "null argument");
    if (N < 0 || D < 0) return fail(HLMC_ERR_PARAM, "negative shape");
    CK(cudaSetDevice(device));
    CK(launch_impute_stats(d_x, N, D, d_sum, reinterpret_cast<long long*>(d_count), static_cast<cudaStream_t>(stream)));
    return HLMC_OK;
}

int hlmc_scaler_stats_f64_device(const double* d_x, int64_t N, int64_t D, const double* d_fill, double* d_mean,
                                 double* d_m2, int device, void* stream) {
    if (!d_mean || !d_m2 || (N > 0 && D > 0 && !d_x)) return fail(HLMC_ERR_PARAM, "null argument");
    if (N < 0 || D < 0) return fail(HLMC_ERR_PARAM, "negative shape");
    if (D == 0) return HLMC_OK;
    CK(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* scratch = nullptr;
    CK(cudaMallocAsync((void**)&scratch, (size_t)3 * D * 8, st));
    const cudaError_t e = launch_scaler_stats_f64(d_x, N, D, d_fill, d_mean, d_m2, scratch, st);
    cudaFreeAsync(scratch, st);
    CK(e);
    return HLMC_OK;
}

int hlmc_impute_scale_device(const double* d_x, int64_t N, int64_t D, const int32_t* d_cols, int64_t D_out,
                             const double* d_fill, const double* d_mean, const double* d_scale, double* d_imputed,
                             double* d_scaled, int device, void* stream) {
    if (N < 0 || D < 0 || D_out < 0 || D_out > D) return fail(HLMC_ERR_PARAM, "bad shape");
    if (N * D_out == 0) return HLMC_OK;
    if (!d_x || !d_fill || (d_scaled && (!d_mean || !d_scale))) return fail(HLMC_ERR_PARAM, "null argument");
    CK(cudaSetDevice(device));
    CK(launch_impute_scale(d_x, N, D, d_cols, D_out, d_fill, d_mean, d_scale, d_imputed, d_scaled,
                           static_cast<cudaStream_t>(stream)));
    return HLMC_OK;
}

void hlmc_last_transfer_bytes(const hlmc_plan* plan, int64_t* h2d, int64_t* d2h) {
    if (h2d) *h2d = plan ? plan->last_h2d : 0;
    if (d2h) *d2h = plan ? plan->last_d2h : 0;
}

int hlmc_measure_fp32_peak(int device, double* tflops) {
    if (!tflops) return fail(HLMC_ERR_PARAM, "null argument");
    CK(cudaSetDevice(device));
    CK(measure_fp32_peak(tflops));
    return HLMC_OK;
}

}  // extern "C"
