/*
 * hlmc_b200.h -- C ABI of the B200-native audio feature extractor.
 *
 * Drop-in boundary for the feature-extraction hot path of
 * Shahriar1638/Hybrid-Language-Music-Clustering-VAE.  The reference has no FFI
 * of its own for this path: its boundary is the Python call surface of librosa
 * used in src/1_preprocessing.py and src/1_preprocessing_advanced.py.  Every
 * entry point below names the librosa call (and the reference call site) whose
 * arithmetic it replaces; the ctypes binding a maintainer adds on the
 * reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; no torch / numpy types cross this boundary
 *   - all buffers are caller-owned; the library never returns owning pointers
 *   - "d_" = device pointer on the plan's device, "h_" = host pointer
 *     (pinned for full speed; pageable works, slower)
 *   - every function returns 0 on success or a negative hlmc_status; the text
 *     of the last error on the calling thread is hlmc_last_error()
 *   - device entry points are asynchronous on `stream` (a cudaStream_t passed
 *     as void*; NULL = legacy default stream)
 *   - layouts are librosa's: feature arrays are (clip, feature, frame),
 *     frame fastest (C order)
 */
#ifndef HLMC_B200_H
#define HLMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HLMC_ABI_VERSION 2

typedef enum hlmc_status {
    HLMC_OK = 0,
    HLMC_ERR_PARAM = -1,      /* librosa would raise ParameterError           */
    HLMC_ERR_UNSUPPORTED = -2,/* valid for librosa, not implemented here      */
    HLMC_ERR_CUDA = -3,       /* CUDA runtime error (see hlmc_last_error)     */
    HLMC_ERR_NOMEM = -4
} hlmc_status;

/* pad_mode of librosa.stft / feature.rms (np.pad modes that librosa accepts) */
enum { HLMC_PAD_CONSTANT = 0, HLMC_PAD_REFLECT = 1, HLMC_PAD_EDGE = 2 };
/* `ref` of librosa.power_to_db: a number, or np.max evaluated per clip       */
enum { HLMC_REF_VALUE = 0, HLMC_REF_MAX = 1 };
/* `norm` of librosa.filters.mel */
enum { HLMC_MELNORM_NONE = 0, HLMC_MELNORM_SLANEY = 1 };
/* rows of the `stats` output, in the order the reference stores them
 * ([R] src/1_preprocessing.py:85-91 dict order; _advanced.py:149)            */
enum { HLMC_STAT_CENTROID = 0, HLMC_STAT_BANDWIDTH = 1, HLMC_STAT_ROLLOFF = 2,
       HLMC_STAT_ZCR = 3, HLMC_STAT_RMS = 4, HLMC_NUM_STATS = 5 };
/* per-clip status bits (the batched analogue of the scripts' per-file
 * try/except: [R] src/1_preprocessing.py:238-251, _advanced.py:165-183)      */
enum { HLMC_CLIP_NONFINITE = 1 };
/* Coverage of HLMC_CLIP_NONFINITE: a clip is flagged when a frame's sum of squares is not finite, i.e. when a
 * non-finite sample lies inside some analysis frame.  librosa.util.valid_audio checks the whole buffer; samples
 * that no frame covers (center=False tails, hop_length > n_fft) are therefore not checked here.  The numpy
 * entry points of the Python layer run np.isfinite over the whole clip first and raise ParameterError as
 * librosa does; CUDA-tensor entry points return the status array and leave the decision to the caller.   */

/* One POD block holding every keyword the reference passes (or leaves at its
 * librosa default) at the call sites of SURVEY.md section 8(a).              */
typedef struct hlmc_params {
    int32_t sr;            /* sr=22050                                        */
    int32_t n_fft;         /* n_fft=2048 (power of two, 64..8192)             */
    int32_t hop_length;    /* hop_length=512                                  */
    int32_t win_length;    /* win_length=None -> n_fft                        */
    int32_t center;        /* center=True                                     */
    int32_t pad_mode;      /* HLMC_PAD_*; librosa>=0.10 default "constant"    */
    int32_t n_mels;        /* n_mels=128                                      */
    float   fmin;          /* fmin=0.0                                        */
    float   fmax;          /* fmax=None -> sr/2 (pass <=0 for None)           */
    int32_t htk;           /* htk=False                                       */
    int32_t mel_norm;      /* HLMC_MELNORM_*; default slaney                  */
    float   power;         /* melspectrogram power=2.0 (1.0 or 2.0)           */
    int32_t n_mfcc;        /* n_mfcc=20 (scripts: 40); 0 = no MFCC            */
    float   lifter;        /* lifter=0                                        */
    int32_t ref_mode;      /* HLMC_REF_*; scripts use ref=np.max              */
    float   ref_value;     /* ref when ref_mode == HLMC_REF_VALUE             */
    float   amin;          /* amin=1e-10                                      */
    float   top_db;        /* top_db=80.0; negative = None                    */
    float   roll_percent;  /* spectral_rolloff roll_percent=0.85              */
    float   zcr_threshold; /* zero_crossings threshold=1e-10                  */
} hlmc_params;

typedef struct hlmc_plan hlmc_plan;   /* opaque; one per (params, device)     */

/* ABI / diagnostics ------------------------------------------------------- */
int         hlmc_abi_version(void);
const char *hlmc_last_error(void);
/* Fill *p with the librosa defaults listed above.                            */
void        hlmc_params_default(hlmc_params *p);
/* T of librosa.util.frame after centre padding: 1 + n // hop (center) or
 * 1 + (n - n_fft) // hop.  Returns <0 (HLMC_ERR_PARAM) when librosa raises.  */
int64_t     hlmc_num_frames(const hlmc_params *p, int64_t n);
/* Number of this library's kernels launched by the calling process so far.   */
int64_t     hlmc_launch_count(void);

/* Plan ---------------------------------------------------------------------
 * Builds on the host, in float64 rounded once to float32, and uploads: the
 * analysis window (librosa.filters.get_window + util.pad_center), FFT
 * twiddles, the banded mel filterbank (librosa.filters.mel) and the DCT-II
 * matrix (scipy.fftpack.dct type 2, norm="ortho", first n_mfcc rows).
 *   window    : NULL -> periodic Hann of win_length; else win_length floats
 *               (what scipy.signal.get_window(window, win_length) returned)
 *   mel_basis : NULL -> built from params; else dense (n_mels, 1+n_fft/2)    */
int  hlmc_plan_create(const hlmc_params *params, const double *window,
                      const float *mel_basis, int device, hlmc_plan **out);
void hlmc_plan_destroy(hlmc_plan *plan);
/* Kernel selection.  n_fft == 2048 (the only size the reference uses) runs the
 * register-FFT kernel with its per-lane tables (window, twiddles, mel weights)
 * in Tensor Memory; n_fft 4096 / 1024 / 512 have register-FFT kernels of their
 * own; other power-of-two sizes run the shared-memory FFT kernel.
 * HLMC_PATH_GENERIC forces the latter, HLMC_PATH_FAST_SMEM_TABLES keeps the
 * n_fft 2048 tables in shared memory (tests cross-check the three); the
 * environment variables HLMC_FORCE_GENERIC=1 / HLMC_NO_TMEM=1 do the same.   */
#define HLMC_PATH_AUTO 0
#define HLMC_PATH_GENERIC 1
#define HLMC_PATH_FAST_SMEM_TABLES 2
int  hlmc_plan_set_path(hlmc_plan *plan, int path);
int  hlmc_plan_uses_fast_path(const hlmc_plan *plan);
/* Copy of the filterbank / DCT the plan uses (for inspection and tests).     */
int  hlmc_plan_mel_basis(const hlmc_plan *plan, float *h_out /* n_mels*(1+n_fft/2) */);
int  hlmc_plan_dct_basis(const hlmc_plan *plan, float *h_out /* n_mfcc*n_mels */);

/* Fused extraction, device-resident ---------------------------------------
 * One call replaces, for a batch of B clips, the librosa chain
 *   feature.melspectrogram -> power_to_db          ([R] 1_preprocessing.py:50-57,
 *                                                       _advanced.py:99-106,125-129)
 *   feature.mfcc                                   ([R] 1_preprocessing.py:63-69)
 *   feature.spectral_centroid / spectral_bandwidth / spectral_rolloff /
 *   zero_crossing_rate / rms                       ([R] 1_preprocessing.py:75-83,
 *                                                       _advanced.py:133-137)
 * computed from ONE STFT per frame instead of the reference's five.
 *   d_wave   : (B, n) float32 waveforms, row pitch `pitch` elements (>= n)
 *   d_logmel : (B, n_mels, T) float32   power_to_db(melspectrogram) [required]
 *   d_mfcc   : (B, n_mfcc, T) float32   or NULL
 *   d_stats  : (B, 5, T) float32        rows HLMC_STAT_*, or NULL
 *   d_status : (B) int32                HLMC_CLIP_* bits, or NULL
 *   d_clipmax: (B) float32 workspace    per-clip max of the mel power
 * power_to_db's ref=np.max and top_db are evaluated PER CLIP (the scripts call
 * librosa once per clip), not over the whole batch.  MFCC follows librosa:
 * DCT of power_to_db(mel, ref=1.0, amin=1e-10, top_db=80).                   */
int hlmc_extract_device(hlmc_plan *plan, const float *d_wave, int64_t B, int64_t n,
                        int64_t pitch, float *d_logmel, float *d_mfcc, float *d_stats,
                        int32_t *d_status, float *d_clipmax, void *stream);

/* hlmc_extract_device captured once as a CUDA graph and replayed: for small batches (1-64 clips) the call is
 * launch-bound (two memsets, a scratch allocation, two kernels), and a replay costs one graph launch.  The
 * graph is bound to the buffers, shapes and plan it was captured with; the mel-power scratch belongs to it.
 * The caller's per-file loop ([R] src/1_preprocessing.py:232-251) becomes: copy the next clip(s) into d_wave,
 * hlmc_graph_launch, read the outputs.  */
typedef struct hlmc_graph hlmc_graph;
int hlmc_graph_create(hlmc_plan *plan, const float *d_wave, int64_t B, int64_t n, int64_t pitch,
                      float *d_logmel, float *d_mfcc, float *d_stats, int32_t *d_status,
                      float *d_clipmax, hlmc_graph **out);
int hlmc_graph_launch(hlmc_graph *graph, void *stream);
void hlmc_graph_destroy(hlmc_graph *graph);

/* The same plus librosa.feature.chroma_stft ([R] src/1_preprocessing.py:96-101,
 * src/1_preprocessing_advanced.py:139-141).  n_fft = 2048 (the one configuration both scripts use; needs a
 * power = 2 plan) takes the piptrack epilogue and the power-spectrum stash of the register-FFT kernel; any other
 * power-of-two n_fft gets the same chroma from two extra passes of the shared-memory FFT kernel (candidates,
 * then the projection): correct for every n_fft, fast for the scripts' one:
 *   d_chroma : (B, 12, T) float32, each frame divided by its largest chroma bin, or NULL
 *   d_tuning : (B) float32, the per-clip librosa.estimate_tuning result, or NULL
 *   d_work   : hlmc_chroma_workspace_bytes(plan, B, n) bytes of device scratch
 * The tuning estimate (piptrack candidates -> median magnitude -> 100-bin residual histogram) is
 * per clip; the chroma filterbank of that tuning comes from 100 filterbanks built at first use.  */
int64_t hlmc_chroma_workspace_bytes(hlmc_plan *plan, int64_t B, int64_t n);
int hlmc_extract_device_ex(hlmc_plan *plan, const float *d_wave, int64_t B, int64_t n,
                           int64_t pitch, float *d_logmel, float *d_mfcc, float *d_stats,
                           int32_t *d_status, float *d_clipmax, float *d_chroma, float *d_tuning,
                           void *d_work, int64_t work_bytes, void *stream);

/* Only what 1_preprocessing.py keeps of a clip ([R] src/1_preprocessing.py:115-129, extract_all_features;
 * src/1_preprocessing_advanced.py:144-156 with with_mfcc = 0): np.mean / np.std over frames of every
 * log-mel band, MFCC, statistic and chroma bin, in the scripts' column order
 *   d_pooled : (B, 2*n_mels + 2*n_mfcc*with_mfcc + 10 + 24*with_chroma) float32
 * power_to_db, the DCT and the pooling run in one kernel: the (B, n_mels, T) and (B, n_mfcc, T) arrays are
 * never written.  d_stats (B, 5, T) and d_clipmax (B) are required scratch; d_chroma (B, 12, T), d_tuning (B)
 * and d_work as for hlmc_extract_device_ex when with_chroma != 0.  */
int hlmc_extract_pooled_device(hlmc_plan *plan, const float *d_wave, int64_t B, int64_t n,
                               int64_t pitch, float *d_pooled, int with_mfcc, int with_chroma,
                               float *d_stats, int32_t *d_status, float *d_clipmax, float *d_chroma,
                               float *d_tuning, void *d_work, int64_t work_bytes, void *stream);

/* Per-kernel timing of hlmc_extract_device for the roofline report: when
 * enabled, CUDA events are recorded on the launching stream around the frames
 * kernel and around the dB+DCT kernel.  hlmc_plan_read_timing synchronises on
 * them, returns the summed durations (ms) and the number of calls, and resets. */
int hlmc_plan_set_timing(hlmc_plan *plan, int enable);
int hlmc_plan_read_timing(hlmc_plan *plan, double *frames_ms, double *db_ms, int64_t *calls);

/* Same, but the mel output is |X|^power projected on the filterbank WITHOUT
 * the dB step (librosa.feature.melspectrogram alone).                        */
int hlmc_melspectrogram_device(hlmc_plan *plan, const float *d_wave, int64_t B, int64_t n,
                               int64_t pitch, float *d_mel, float *d_stats,
                               int32_t *d_status, void *stream);

/* librosa.stft: (B, 1+n_fft/2, T) complex64 (interleaved re,im).             */
int hlmc_stft_device(hlmc_plan *plan, const float *d_wave, int64_t B, int64_t n,
                     int64_t pitch, float *d_spec, void *stream);

/* librosa.power_to_db on an arbitrary (B, rows, T) float32 array, ref / top_db
 * per leading index.  In place when d_out == d_in.                           */
int hlmc_power_to_db_device(const float *d_in, float *d_out, int64_t B, int64_t rows,
                            int64_t T, int32_t ref_mode, float ref_value, float amin,
                            float top_db, float *d_clipmax, int device, void *stream);

/* Time pooling of the scripts ([R] 1_preprocessing.py:115-124,
 * _advanced.py:144-151): np.mean / np.std(ddof=0) over frames of every logmel
 * row, every MFCC row and the five statistics.
 *   d_pooled : (B, 2*n_mels + 2*n_mfcc + 10) float32 =
 *              [mel mean | mel std | mfcc mean | mfcc std | (mean,std) x 5]
 *   d_mfcc may be NULL (then the MFCC columns are omitted: 2*n_mels + 10).   */
int hlmc_pool_device(hlmc_plan *plan, const float *d_logmel, const float *d_mfcc,
                     const float *d_stats, int64_t B, int64_t T, float *d_pooled,
                     void *stream);

/* With d_chroma != NULL the pooled row grows by [chroma mean (12) | chroma std (12)]: the full
 * 370 / 290 column layout of the scripts.                                     */
int hlmc_pool_device_ex(hlmc_plan *plan, const float *d_logmel, const float *d_mfcc,
                        const float *d_stats, const float *d_chroma, int64_t B, int64_t T,
                        float *d_pooled, void *stream);

/* [R] _advanced.py:108-112: crop to fixed_time_steps frames, or right-pad
 * with the clip's minimum.  d_out: (B, rows, fixed) float32.                 */
int hlmc_fix_frames_device(const float *d_in, float *d_out, int64_t B, int64_t rows,
                           int64_t T, int64_t fixed, int device, void *stream);

/* Fused extraction, host buffers -------------------------------------------
 * The reference-facing call: waveforms in host memory in, features in host
 * memory out.  Internally chunks the batch, and overlaps H2D copies, kernels
 * and D2H copies on `n_streams` streams.  Blocking.  Any output may be NULL.
 *   chunk_clips <= 0 -> chosen by the library.                               */
int hlmc_extract_host(hlmc_plan *plan, const float *h_wave, int64_t B, int64_t n,
                      int64_t h_pitch, float *h_logmel, float *h_mfcc, float *h_stats,
                      int32_t *h_status, float *h_pooled, int64_t chunk_clips, int n_streams);

/* Same with the front end of the scripts' load_audio_file ([R] src/1_preprocessing.py:137-153,
 * src/1_preprocessing_advanced.py:79-94) folded in:
 *   sample_format HLMC_SAMPLES_PCM16: h_wave holds int16 PCM as stored in the WAV file; the device
 *     converts with float32 = int16 / 32768 (what librosa.load returns for PCM16), so only 2 bytes
 *     per sample cross PCIe;
 *   n_total > n_valid: every clip is right zero-padded on the device to n_total samples (the
 *     scripts' np.pad to sample_rate * duration) instead of shipping the zeros.
 * h_pitch is in elements of the given format.  Frame counts follow n_total.                  */
enum { HLMC_SAMPLES_F32 = 0, HLMC_SAMPLES_PCM16 = 1 };
int hlmc_extract_host_ex(hlmc_plan *plan, const void *h_wave, int sample_format, int64_t B,
                         int64_t n_valid, int64_t h_pitch, int64_t n_total, float *h_logmel,
                         float *h_mfcc, float *h_stats, int32_t *h_status, float *h_pooled,
                         int64_t chunk_clips, int n_streams);

/* The general host entry point: everything above plus chroma_stft, the advanced script's fixed-width
 * image and the rest of librosa.load's front end, as one POD request (zero-initialise it: a zero field
 * means "not wanted" / "default").  Any output pointer may be NULL.  pooled_with_chroma != 0 appends
 * [chroma mean | chroma std] to the pooled rows (the scripts' full 370 / 290 columns).
 *
 * Front end ([R] src/1_preprocessing.py:137-153, src/1_preprocessing_advanced.py:79-94,
 * librosa.load(path, sr=22050, duration=30) then np.pad to sr * duration):
 *   channels > 1 : `wave` holds interleaved frames (as in the WAV data chunk); the device averages the
 *                  channels in float32 in channel order, then divides by the count (librosa.to_mono = np.mean)
 *   sr_in != 0 and != params.sr : the mono signal is resampled sr_in -> params.sr on the device with
 *                  librosa.resample(res_type="polyphase") = scipy.signal.resample_poly(up, down) with its
 *                  default Kaiser(beta=5) low-pass of 20*max(up,down)+1 taps, output length
 *                  ceil(n_valid * sr / sr_in) (librosa's fix_length).  librosa.load's DEFAULT res_type is
 *                  "soxr_hq", a different (also linear-phase, higher-order) low-pass: results differ in the
 *                  transition band above ~0.45 sr; pass res_type="polyphase" on the reference side for parity.
 *   n_valid, pitch : in FRAMES of the input (one frame = `channels` samples); n_total: samples per clip
 *                  AFTER resampling that the features see (right zero pad, the scripts' np.pad).
 * fixed_logmel ([R] src/1_preprocessing_advanced.py:108-112): the log-mel image cropped to
 * fixed_frames frames or right-padded with the clip's minimum, (B, n_mels, fixed_frames).        */
typedef struct hlmc_host_io {
    const void *wave;          /* (B, pitch [, channels]) float32 or int16                */
    int32_t     sample_format; /* HLMC_SAMPLES_*                                          */
    int64_t     B, n_valid, pitch, n_total;   /* n_total <= 0 means the (resampled) clip length */
    float      *logmel;        /* (B, n_mels, T)                                          */
    float      *mfcc;          /* (B, n_mfcc, T)                                          */
    float      *stats;         /* (B, 5, T)                                               */
    float      *chroma;        /* (B, 12, T)                                              */
    float      *tuning;        /* (B)                                                     */
    float      *pooled;        /* (B, 2*n_mels + 2*n_mfcc + 10 [+ 24])                    */
    int32_t    *status;        /* (B)                                                     */
    int32_t     pooled_with_chroma;
    int64_t     chunk_clips;   /* <= 0: chosen by the library                             */
    int32_t     n_streams;     /* <= 0: 3                                                 */
    /* ABI 2 */
    int32_t     channels;      /* <= 1: mono                                              */
    int32_t     sr_in;         /* 0: already at params.sr                                 */
    int32_t     reserved0;
    float      *fixed_logmel;  /* (B, n_mels, fixed_frames) or NULL                       */
    int64_t     fixed_frames;
    float      *wave_out;      /* (B, n_total) float32: the clips as the features saw them (after mono mix,
                                  resampling and padding), or NULL                        */
    const int64_t *valid_frames; /* (B) per-clip frame counts <= n_valid, or NULL (all n_valid): files shorter
                                  than `duration` -- frames past a clip's count are ignored, its resampled
                                  length follows its own count, the rest of n_total is the zero pad  */
} hlmc_host_io;
int hlmc_extract_host_io(hlmc_plan *plan, const hlmc_host_io *io);

/* Bytes moved by the last hlmc_extract_host call on this plan.               */
void hlmc_last_transfer_bytes(const hlmc_plan *plan, int64_t *h2d, int64_t *d2h);

/* sklearn.preprocessing.StandardScaler for the big (N, 131072) mel matrix
 * ([R] src/1_preprocessing_advanced.py:376-382), SURVEY.md 8f-4.
 *   hlmc_column_stats_device: per column, float64 mean and corrected sum of squared deviations
 *     (var = m2 / N), computed like sklearn's _incremental_mean_and_var; per-rank results combine
 *     across GPUs with Chan's formula (the one collective of the whole path, done by the caller);
 *   hlmc_standardize_device: y = (x - mean) / scale in float32, as StandardScaler.transform does
 *     for float32 input.  In place when d_y == d_x.                                           */
int hlmc_column_stats_device(const float *d_x, int64_t N, int64_t D, double *d_mean, double *d_m2,
                             int device, void *stream);
int hlmc_standardize_device(const float *d_x, float *d_y, int64_t N, int64_t D,
                            const float *d_mean, const float *d_scale, int device, void *stream);

/* librosa.load's front end for a device-resident batch (see hlmc_host_io for the semantics):
 * int16 / float32, interleaved channels -> mono float32 -> polyphase resampling -> right zero pad.
 *   d_raw   : (B, raw_pitch, channels) samples, raw_pitch in frames
 *   d_wave  : (B, pitch) float32 out, n_total samples written per clip
 * hlmc_resampled_length = ceil(n_in * sr_out / sr_in), librosa.resample's output length.            */
int64_t hlmc_resampled_length(int64_t n_in, int32_t sr_in, int32_t sr_out);
int hlmc_load_frontend_device(hlmc_plan *plan, const void *d_raw, int sample_format, int channels,
                              int64_t B, int64_t n_in, int64_t raw_pitch, int32_t sr_in,
                              float *d_wave, int64_t pitch, int64_t n_total,
                              const int64_t *d_valid_frames /* (B) or NULL */, void *stream);
/* The taps the resampler uses (scipy.signal.firwin(20*max(up,down)+1, 1/max(up,down), window=("kaiser", 5.0))
 * cast to float32, times up): fills up to `cap` floats, returns the tap count (or < 0).              */
int64_t hlmc_resample_taps(int32_t sr_in, int32_t sr_out, float *h_out, int64_t cap);

/* The scripts' tabular normalisation on device ([R] src/1_preprocessing.py:303-311,
 * src/1_preprocessing_advanced.py:384-391) for the (N, 370) / (N, 290) float64 feature matrix:
 *   np.where(np.isinf(X), np.nan, X) -> SimpleImputer(strategy="mean") -> StandardScaler().
 * hlmc_impute_stats_device : per column, the sum and the count of the finite entries
 *     (SimpleImputer.statistics_ = sum / count; a column with count 0 is dropped by sklearn);
 * hlmc_scaler_stats_f64_device : per column of the imputed matrix (non-finite -> d_fill[c]) the mean and the
 *     corrected sum of squared deviations (var = m2 / N), as sklearn's _incremental_mean_and_var;
 * hlmc_impute_scale_device : d_imputed[r, j] = x[r, cols[j]] or d_fill[cols[j]] where non-finite;
 *     d_scaled[r, j] = (d_imputed[r, j] - mean[j]) / scale[j].  cols (D_out int32) lists the kept
 *     columns; either output may be NULL.  Per-rank sums / counts / (mean, m2) combine across GPUs by
 *     addition / Chan's formula on the caller's side (torch.distributed), the only collectives of the path. */
int hlmc_impute_stats_device(const double *d_x, int64_t N, int64_t D, double *d_sum, int64_t *d_count,
                             int device, void *stream);
int hlmc_scaler_stats_f64_device(const double *d_x, int64_t N, int64_t D, const double *d_fill,
                                 double *d_mean, double *d_m2, int device, void *stream);
int hlmc_impute_scale_device(const double *d_x, int64_t N, int64_t D, const int32_t *d_cols, int64_t D_out,
                             const double *d_fill, const double *d_mean, const double *d_scale,
                             double *d_imputed, double *d_scaled, int device, void *stream);

/* Measurement helper: a dependent-FMA micro-benchmark that returns the
 * achieved FP32 TFLOP/s of the plan's device (the FP32-pipe roofline
 * denominator of SURVEY.md 8(d)); no product path calls it.                  */
int hlmc_measure_fp32_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* HLMC_B200_H */
