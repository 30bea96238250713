/* cpp_audio_b200.h — C ABI of the B200-native rt.resynth hot path (periodic-FFT analysis + sine-bank resynthesis).
 *
 * The reference (OlivierSohn/cpp.audio) has no FFI: its seams are C++ concepts. Each entry point below names the
 * reference interface it stands in for (path:line under the reference tree). C++ adapters that present those
 * concepts on top of this ABI live in cpp.audio_b200/host/.
 *
 * Conventions
 *   - return 0 (CA_OK) or a negative ca_status; no exceptions cross the ABI; ca_last_error(ctx) gives the text.
 *   - pointers named d_* are DEVICE pointers owned by the caller; h_* are HOST pointers (pinned recommended).
 *   - device entry points are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 *   - no allocation happens in the device entry points after the matching ca_*_plan call (mirrors the reference's
 *     "no dynamic allocation after setLambdas" rule, source/rt.resynth.lib.periodicfft.cpp:32).
 *   - one ctx per GPU; a ctx is not thread-safe (the reference drives each operator from a single thread).
 *   - there is NO CPU fallback: every entry point fails with CA_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef CPP_AUDIO_B200_H
#define CPP_AUDIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ca_ctx ca_ctx;
typedef void* ca_stream;

enum ca_status {
  CA_OK = 0,
  CA_ERR_INVALID = -1,     /* bad argument */
  CA_ERR_CUDA = -2,        /* CUDA runtime error (text in ca_last_error) */
  CA_ERR_UNSUPPORTED = -3, /* shape outside what the kernels cover */
  CA_ERR_NOT_PLANNED = -4, /* ca_*_plan was not called or is too small */
  CA_ERR_NOMEM = -5
};

/* ---- context ------------------------------------------------------------------------------------------------ */
const char* ca_version(void);
int ca_init(int device, ca_ctx** ctx);
void ca_destroy(ca_ctx* ctx);
const char* ca_last_error(const ca_ctx* ctx);
int ca_sm_count(const ca_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
int64_t ca_kernel_launches(const ca_ctx* ctx);
/* pinned host memory for the *_host entry points */
void* ca_host_alloc(size_t bytes);
void ca_host_free(void* p);

/* ---- K1: batched forward real FFT ------------------------------------------------------------------------------
 * Stands in for fft::Algo_<Tag,T>::forward as used by FftOperation<T>::operator()
 * (source/rt.resynth.lib.periodicfft.cpp:210-235) and forward_fft (source/main.test_fft.cpp:78).
 * Unscaled: X[k] = sum_n x[n] exp(-2 pi i k n / n_fft). d_in [batch][n_fft] float, d_out [batch][n_fft/2+1] (re,im).
 * n_fft: power of two, 2^9 .. 2^20. */
int ca_rfft_batched(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, ca_stream stream);
int ca_rfft_batched_host(ca_ctx* ctx, const float* h_in, int n_fft, int64_t batch, float* h_out);
/* Inverse: stands in for fft::Algo_<Tag,T>::inverse as the vocoder's carrier path uses it
 * (source/rt.resynth.lib.vocoder.cpp:438-470). d_in [batch][n_fft/2+1] (re,im), Hermitian extension implied (bins 0 and n_fft/2 must
 * be real); d_out [batch][n_fft] = scale * sum_k X[k] exp(+2 pi i k n / n_fft), so
 * irfft(rfft(x)) = n_fft * x at scale 1. n_fft: power of two, 2^9 .. 2^15. */
int ca_irfft_batched(ca_ctx* ctx, const float* d_in, int n_fft, int64_t batch, float* d_out, float scale, ca_stream stream);

/* ---- K2: periodic-FFT analysis ----------------------------------------------------------------------------------
 * Stands in for PeriodicFFT<SqMagFftOperation<Window,T>> (source/rt.resynth.lib.periodicfft.cpp:13-181, :251-325)
 * followed by extractLocalMaxFreqsMags (call site source/rt.resynth.lib.cpp:1592-1595), for whole streams at once:
 * frame f of a stream covers samples [f*hop, f*hop+window); frames = floor((n-window)/hop)+1. */
enum ca_window { CA_WINDOW_RECTANGULAR = 0, CA_WINDOW_GAUSSIAN = 1 };

typedef struct ca_stft_params {
  int32_t window;       /* even (Assert at periodicfft.cpp:274) */
  int32_t hop;          /* window_center_stride in frames, >= 1 */
  int32_t window_type;  /* ca_window */
  float sigma_factor;   /* Gaussian truncation, reference uses 4 (periodicfft.cpp:288) */
  int32_t sample_rate;
  float min_db;         /* peaks below this level are not reported; -INFINITY = none (reference behaviour) */
  int32_t max_peaks;    /* capacity per frame; reference reserves 200 (rt.resynth.lib.cpp:1102) */
} ca_stft_params;

int ca_fft_length_for(int n_samples, int zero_padding_factor);
int64_t ca_stft_num_frames(int64_t n_samples, int window, int hop);
/* builds the window table on the device; the only allocating call of the analysis path. A ctx holds ONE plan for
 * ca_stft_analyze*: a second ca_stft_plan replaces it (the round trip and the vocoder keep plans of their own). */
int ca_stft_plan(ca_ctx* ctx, const ca_stft_params* params);
/* counts the ca_stft_plan calls of the ctx: an operator that shares its ctx with others remembers the value its own plan got
 * and plans again when it has moved on (cpp.audio_b200/host/cuda_periodic_fft.hpp, cpp.audio_b200/analysis.py do) */
uint64_t ca_stft_plan_generation(const ca_ctx* ctx);
/* d_x [n_streams][stream_stride] (n_samples used per stream).
 * d_sqmag (nullable) [n_streams*frames][sqmag_row_stride], row holds n_fft/2+1 values (sqmag_row_stride >= that).
 * peaks (all nullable together): d_peak_bin/freq/db [n_streams*frames][max_peaks] ascending in frequency,
 * d_peak_count [n_streams*frames] = number of peaks found (may exceed max_peaks; only the first max_peaks stored). */
int ca_stft_analyze(ca_ctx* ctx, const float* d_x, int64_t n_samples, int64_t n_streams, int64_t stream_stride,
                    float* d_sqmag, int64_t sqmag_row_stride, int32_t* d_peak_bin, float* d_peak_freq,
                    float* d_peak_db, int32_t* d_peak_count, ca_stream stream);
/* same with host buffers: H2D of x, kernel, D2H of every non-null output, stream sync. Uses ctx scratch (grown on
 * demand, reused across calls). */
int ca_stft_analyze_host(ca_ctx* ctx, const float* h_x, int64_t n_samples, int64_t n_streams, int64_t stream_stride,
                         float* h_sqmag, int64_t sqmag_row_stride, int32_t* h_peak_bin, float* h_peak_freq,
                         float* h_peak_db, int32_t* h_peak_count);

/* ---- K3: sine oscillator bank ----------------------------------------------------------------------------------
 * Stands in for ImplCRTP::compute (include/gen.crtp.h:350-378) over elements
 * VolumeAdjusted<MultiEnveloped<SineOscillatorAlgo,AHDSREnvelope>> (include/audioelement.h:486-657, :1111-1265,
 * :2366-2452), for whole notes at once instead of 16-frame blocks. */
/* itp::interpolation of the envelope phases (include/audioelement.h:314-335, evaluated at :733-737, :994-998). The enum
 * itself lives in cpp.algorithms (absent); the numbering is ours, the equations are Penner's (oracle/SPEC.md S7).
 * rt.resynth uses LINEAR everywhere (rt.resynth.lib.cpp:75-81), tune EASE_OUT_CUBIC / LINEAR (main.tune.cpp:1146-1161).
 * value = CA_ITP_EASE_IN_QUAD + 3*family + mode, family = QUAD, CUBIC, QUART, QUINT, SINE, EXPO, CIRC, mode = IN, OUT, INOUT.
 * The oscillator bank (ca_oscbank_*) evaluates all of them; the round trip and the ramp voices take LINEAR only and
 * fail with CA_ERR_UNSUPPORTED otherwise. */
enum ca_itp {
  CA_ITP_LINEAR = 0,
  CA_ITP_EASE_IN_QUAD = 1, CA_ITP_EASE_OUT_QUAD = 2, CA_ITP_EASE_INOUT_QUAD = 3,
  CA_ITP_EASE_IN_CUBIC = 4, CA_ITP_EASE_OUT_CUBIC = 5, CA_ITP_EASE_INOUT_CUBIC = 6,
  CA_ITP_EASE_IN_QUART = 7, CA_ITP_EASE_OUT_QUART = 8, CA_ITP_EASE_INOUT_QUART = 9,
  CA_ITP_EASE_IN_QUINT = 10, CA_ITP_EASE_OUT_QUINT = 11, CA_ITP_EASE_INOUT_QUINT = 12,
  CA_ITP_EASE_IN_SINE = 13, CA_ITP_EASE_OUT_SINE = 14, CA_ITP_EASE_INOUT_SINE = 15,
  CA_ITP_EASE_IN_EXPO = 16, CA_ITP_EASE_OUT_EXPO = 17, CA_ITP_EASE_INOUT_EXPO = 18,
  CA_ITP_EASE_IN_CIRC = 19, CA_ITP_EASE_OUT_CIRC = 20, CA_ITP_EASE_INOUT_CIRC = 21,
  CA_ITP_COUNT = 22
};

typedef struct ca_ahdsr { /* audioelement::AHDSR, include/audioelement.h:314-335 (durations in samples) */
  int32_t attack;
  int32_t attack_itp;
  int32_t hold;
  int32_t decay;
  int32_t decay_itp;
  int32_t release;
  int32_t release_itp;
  float sustain;
} ca_ahdsr;

enum ca_mix { CA_MIX_NONE = 0, CA_MIX_STEREO = 1, CA_MIX_MONO = 2 };
/* What the caller knows about the *_itp fields of the envelopes of a render (they live in device memory for ca_oscbank_render).
 * LINEAR and eased voices are rendered by two passes; a pass that has nothing to do still costs a launch (90 us on 1024
 * voices x 10 s), so a caller that knows says so. A wrong promise leaves the voices of the other kind silent. */
enum ca_env_kind { CA_ENV_UNKNOWN = 0, CA_ENV_ALL_LINEAR = 1, CA_ENV_ALL_EASED = 2, CA_ENV_MIXED = 3 };

typedef struct ca_oscbank_desc {
  int64_t n_voices;
  int32_t n_partials;
  int64_t n_samples;
  int32_t sample_rate;
  const float* f0_hz;          /* [voices] note frequency (Event::noteOn.frequency) */
  const float* volume;         /* [voices] volume target = baseVolume*velocity (gen.crtp.h:425) */
  const float* start_angle;    /* [voices] nullable; voice start angle in units of pi rad (setStartAngle) */
  const int32_t* key_delay;    /* [voices] nullable; onKeyPressed(delay) */
  const int64_t* release_at;   /* [voices] nullable; onKeyReleased(0) is applied before the step of this sample; <0 never */
  const float* partial_volume; /* [voices][partials] harmonicProperties_t::volume */
  const float* partial_phase;  /* [voices][partials] harmonicProperties_t::phase (units of pi rad) */
  const ca_ahdsr* env;         /* [n_env] */
  int64_t n_env;               /* 1 (shared) or n_voices */
  const float* pan;            /* [voices] nullable, used by CA_MIX_STEREO (include/stereo.h:18-27) */
  int32_t mix;                 /* ca_mix */
  int64_t first_sample;        /* render samples [first_sample, first_sample+n_samples) of the notes' timeline (block-wise or
                                  time-sharded rendering; phases and envelopes are closed-form in the sample index) */
  int32_t env_kind;            /* ca_env_kind; 0 = unknown (always correct). ca_oscbank_render_host works it out itself */
} ca_oscbank_desc;

/* sizes scratch for up to max_voices x max_partials; the only allocating call of the synthesis path */
int ca_oscbank_plan(ca_ctx* ctx, int64_t max_voices, int32_t max_partials);
/* every pointer inside desc is a DEVICE pointer. d_out: CA_MIX_NONE [voices][samples]; CA_MIX_STEREO interleaved
 * [samples][2]; CA_MIX_MONO [samples]. Mixed outputs are overwritten (zeroed first), not accumulated. */
int ca_oscbank_render(ca_ctx* ctx, const ca_oscbank_desc* desc, float* d_out, ca_stream stream);
/* every pointer inside desc is a HOST pointer; H2D of the note parameters, kernels, D2H of h_out, stream sync */
int ca_oscbank_render_host(ca_ctx* ctx, const ca_oscbank_desc* desc, float* h_out);

/* ---- analysis -> pitch pipeline -> resynthesis round trip (BASELINE config 4) -----------------------------------
 * Stands in for the offline RtResynth loop (source/rt.resynth.lib.cpp:1185-1235): per stream, every analysis frame's
 * peak list goes through RtResynth::step (:1670-1759; source/rt.resynth.lib.algo.cpp:10-440: frequencies_to_pitches,
 * aggregate_pitches, reduce_pitches, track_pitches, order by loudness) and synthesize_sounds (:264-382) turns the result
 * into NoteOn / NoteChange / NoteOff on a pool of ResynthElement voices (:5-18: StereoPanned<VolumeAdjusted<Enveloped<
 * FreqCtrl_<SineOscillatorAlgo, InterpolatedFreq>, AHDSREnvelope>>>) whose output is the resynthesised audio.
 * Analysis, the pitch pipeline (frame-sequential per stream: one CTA per stream walks its frames) and voice rendering all run on
 * the GPU; only the audio crosses PCIe. Events of frame k act from sample window-1 + k*hop on. max_channels is limited to 256 on
 * this path, stft.max_peaks to 16384; see ca_resynth_truncated_frames. */
/* Optional stages of RtResynth::step between reduce_pitches and track_pitches (source/rt.resynth.lib.cpp:1687-1707):
 * shift_pitches / harmonize_pitches (source/rt.resynth.lib.algo.cpp:307-371) before and after autotune_pitches (:186-229).
 * All zero = rt.resynth out of the box (pitch_shift_* = pitch_harmonize_* = 0, use_autotune = false). The transform of
 * mkAutotuneFunction (:1762-1869) is passed as data: AutotuneType::Chord and ::FixedSizeIntervals both reduce to "closest
 * entry of the sorted vector allowed_pitches" (CA_AUTOTUNE_TABLE: hand over that vector), ::MusicalScale to
 * MusicalScalePitches::closest_pitch (source/rt.resynth.lib.autotune.cpp:96-146; CA_AUTOTUNE_SCALE: the scale's pitches
 * 0 ... 12 and the root pitch). A frame's list is capped at stft.max_peaks entries at every stage. */
enum ca_autotune { CA_AUTOTUNE_NONE = 0, CA_AUTOTUNE_TABLE = 1, CA_AUTOTUNE_SCALE = 2 };
typedef struct ca_pitch_options {
  float shift_pre_autotune;         /* pitch_shift_pre_autotune, semitones */
  float harmonize_pre_autotune;     /* pitch_harmonize_pre_autotune */
  float shift_post_autotune;
  float harmonize_post_autotune;
  int32_t autotune;                 /* ca_autotune */
  float autotune_max_pitch;         /* pitches above are played but not tuned (autotune_max_pitch) */
  float autotune_tolerance_pitches; /* pitches further than this from their tuned pitch are not played */
  int32_t n_table;                  /* TABLE: <= 4096 entries, may be 0 (nothing is played below max_pitch); SCALE: 2 ... 16 */
  const double* table;              /* HOST pointer. TABLE: allowed pitches, ascending; SCALE: scale pitches, first 0, last 12 */
  double scale_root_pitch;          /* SCALE: A_pitch + transpose + half_tones_distance(La, root note), :1862-1865 */
} ca_pitch_options;

typedef struct ca_roundtrip_params {
  ca_stft_params stft;          /* window / hop / sample rate / min_db; max_peaks = peak capacity per frame */
  float nearby_distance_tones;  /* 0.4   rt.resynth.lib.cpp:948 */
  float min_volume;             /* 1e-4  :947 */
  float max_track_pitches;      /* 1.0   :949 */
  float gain;                   /* analysis_volume, :995 */
  float stereo_spread;          /* :955 */
  ca_ahdsr env;                 /* :957-961 (all zero, sustain 1 by default) */
  int32_t max_channels;         /* 254 = 2*127, :201-210 */
  int32_t host_threads;         /* 0: pitch pipeline on the device (default). n > 0 / -1: its host form on n / all host threads (A/B) */
  ca_pitch_options opts;        /* shift / harmonize / autotune, all zero = off */
} ca_roundtrip_params;
/* h_x [n_streams][n_samples] -> h_out [n_streams][n_samples][2] float32 (stereo). h_stats nullable [n_streams][2]:
 * notes started, notes dropped (pool full). */
int ca_resynth_roundtrip_host(ca_ctx* ctx, const ca_roundtrip_params* params, const float* h_x, int64_t n_samples,
                              int64_t n_streams, float* h_out, int64_t* h_stats);

/* The same with 16-bit PCM input, the usual content of the job's WAV files (RtResynthOfflineJob::read_voice,
 * source/rt.resynth.lib.params.cpp:183-386): h_x [n_streams][n_samples] int16, taken as x / 32768 like a WAV reader does (exact in
 * float32). The input crosses PCIe at 2 bytes per sample instead of 4; the call is copy-bound (multi-GPU). */
int ca_resynth_roundtrip_pcm16_host(ca_ctx* ctx, const ca_roundtrip_params* params, const int16_t* h_x, int64_t n_samples,
                                    int64_t n_streams, float* h_out, int64_t* h_stats);

/* Frames of the last ca_resynth_roundtrip_host call on this ctx that lost something to a capacity: more than stft.max_peaks
 * peaks (only the first max_peaks, in ascending frequency, took part), or more than 256 entries left after aggregation and the
 * min_volume cut (the device pitch pipeline's list; the voice pool has 254 channels). The reference's lists are unbounded
 * (reserve(200) at rt.resynth.lib.cpp:1102 is a reserve, not a cap), so a non-zero value means the result is NOT what the
 * reference computes. stft.max_peaks = n_fft/4 can never overflow (a spectrum has at most that many strict local maxima); it
 * costs 20 bytes of scratch per peak and frame. */
int64_t ca_resynth_truncated_frames(const ca_ctx* ctx);

/* ---- a note schedule on the rt.resynth voice pool -------------------------------------------------------------------
 * Stands in for ImplCRTP::onEvent + ImplCRTP::compute (include/gen.crtp.h:382-622, :350-378) of the synth rt.resynth drives
 * (sine::Synth<2, ResynthElement, ..., SynchronizePhase::Yes, DefaultStartPhase::Random, HandleNoteOff::Yes, 127,
 * ResynthElementInitializer>, source/rt.resynth.lib.cpp:201-210), for a host that owns its note schedule: the synthesis half of
 * the round trip on its own. Events are the reference's (include/events.h:12-107) plus the sample time at which they take
 * effect: the events of sample n are applied, in array order, before the pool steps to produce output sample n
 * (SimpleAudioOutContext::step, include/simple_audio_context.h:69-112).
 *   NoteOn     takes the first free channel (DROPPED_NOTE when none is, counted in h_stats[1]); volume target 0.1 * velocity;
 *              start phase: that of the first other sounding channel with the same angle increment, else random
 *              (setPhase, gen.crtp.h:127-156; random = a hash of the note id, also used for the pan: oracle/SPEC.md S10)
 *   NoteChange new volume target and a frequency glide of `glide_samples` samples on every channel that holds the id
 *   NoteOff    releases the first channel that holds the id
 * Note ids should be unique per note (the reference asks for the same, gen.crtp.h:575-577). */
enum ca_event_type { CA_NOTE_ON = 0, CA_NOTE_OFF = 1, CA_NOTE_CHANGE = 2 }; /* EventType, include/events.h:12-17 */
typedef struct ca_track_event {
  int64_t at_sample;  /* >= 0, non-decreasing along the array */
  int64_t noteid;     /* NoteId */
  int32_t type;       /* ca_event_type */
  float frequency;    /* Hz: noteOn.frequency / noteChange.changed_frequency */
  float velocity;     /* noteOn.velocity / noteChange.changed_velocity */
  int32_t reserved;
} ca_track_event;
typedef struct ca_tracks_desc {
  int64_t n_events;
  const ca_track_event* events; /* HOST pointer in both forms (control-rate data) */
  int64_t n_samples;
  int32_t sample_rate;
  int32_t glide_samples;        /* ResynthElementInitializer's stride: glide duration, volume filter limit 2/stride */
  ca_ahdsr env;                 /* LINEAR interpolations */
  int32_t max_channels;         /* 254 in rt.resynth */
  float stereo_spread;          /* pan = stereo_spread * U(-1,1) per note (CA_MIX_STEREO) */
  int32_t mix;                  /* CA_MIX_STEREO: out [n_samples][2]; CA_MIX_MONO: out [n_samples], no pan */
} ca_tracks_desc;
/* d_out: DEVICE memory, every sample is written; asynchronous on `stream`. h_stats (nullable, host): notes started, dropped. */
int ca_oscbank_render_tracks(ca_ctx* ctx, const ca_tracks_desc* desc, float* d_out, int64_t* h_stats, ca_stream stream);
/* h_out: HOST memory; returns when it is filled */
int ca_oscbank_render_tracks_host(ca_ctx* ctx, const ca_tracks_desc* desc, float* h_out, int64_t* h_stats);

/* The same pool as an object, for a host that posts events as they become known and pulls audio block by block (what
 * ImplCRTP::onEvent + compute(double*, int) do in the reference): ca_tracks_open takes the parameters of `desc` (its events,
 * if any, are posted; n_samples is ignored), ca_tracks_post adds events (ordered; an event earlier than the position acts at
 * the position, as a late MIDI event does, gen.crtp.h:503-510), ca_tracks_pull* renders the next n_samples and advances the
 * position. ca_tracks_state: [0] position, [1] notes started, [2] notes dropped, [3] sounding notes, [4] pending events. */
typedef struct ca_tracks ca_tracks;
int ca_tracks_open(ca_ctx* ctx, const ca_tracks_desc* desc, ca_tracks** tracks);
int ca_tracks_post(ca_tracks* tracks, const ca_track_event* events, int64_t n_events);
int ca_tracks_pull(ca_tracks* tracks, int64_t n_samples, float* d_out, ca_stream stream);
int ca_tracks_pull_host(ca_tracks* tracks, int64_t n_samples, float* h_out);
int ca_tracks_state(const ca_tracks* tracks, int64_t state[5]);
void ca_tracks_close(ca_tracks* tracks);

/* ---- gen.voice ramp voices ---------------------------------------------------------------------------------------
 * Stands in for the inner element of voice::Impl_::doProcessing / SoundEngine::step (include/gen.voice.h:1130-1189;
 * include/soundengine.h:245-257, 412-464): Enveloped<FreqCtrl_<SineOscillatorAlgo, LogRamp>, AHDSREnvelope> with a constant
 * volume. The oscillator's angle increment follows LogRamp::setup(from, to, duration_in_samples, start_sample, LINEAR)
 * (include/audioelement.h:2464-2572): it glides from `from` to `to` in about `duration` samples and bounces back, for ever.
 * A SoundEngine voice (up to three cross-faded ramps) is a sum of such voices with staggered key_delay / release_at.
 * The equal-loudness volume of the full element (Mix<LoudnessAdjusted<Sine>>) is not applied. Host pointers. */
typedef struct ca_ramp_desc {
  int64_t n_voices;
  int64_t n_samples;
  int32_t sample_rate;
  int32_t mix;                /* CA_MIX_NONE: h_out [voices][samples]; CA_MIX_MONO: h_out [samples] = sum of the voices */
  const float* from_hz;       /* [voices] > 0 */
  const float* to_hz;         /* [voices] > 0 */
  const float* duration;      /* [voices] > 0, LogRamp duration_in_samples */
  const float* start_sample;  /* nullable [voices]: initial ramp position, 0 <= start_sample <= duration */
  const float* volume;        /* nullable [voices] (1) */
  const float* start_angle;   /* nullable [voices]: oscillator start angle in units of pi rad */
  const int32_t* key_delay;   /* nullable [voices]: onKeyPressed(delay) */
  const int64_t* release_at;  /* nullable [voices]: onKeyReleased(0) just before this sample; < 0: never */
  ca_ahdsr env;               /* one envelope for all voices; floors of 1 ms and 2.5 periods of `from` apply */
} ca_ramp_desc;
int ca_ramp_render_host(ca_ctx* ctx, const ca_ramp_desc* desc, float* h_out);

/* ---- gen.voice voices: SoundEngine ---------------------------------------------------------------------------------
 * Stands in for voice::Impl_::doProcessing (include/gen.voice.h:1130-1189) over channels of VolumeAdjusted<SoundEngine>
 * (include/soundengine.h:240-468, :738-803): every voice plays its list of ramp specs, each on one of three ramp elements
 * VolumeAdjusted<Enveloped<FreqCtrl_<Mix<LoudnessVolumeAdjusted<Sine>>, LogRamp>, AHDSR>> that fade in and out over xfade_len samples
 * (the orchestration of orchestrate_algos / playNextSpec, :744-803). The specs are what the engine's Markov chain produces at
 * note-on (:496-736): ca_sound_engine_specs below builds them (the chain and the random generator live in cpp.algorithms: its
 * decisions are oracle/SPEC.md S15), or the host supplies its own. Events: NoteOn at key_on (the first spec starts on that sample),
 * NoteOff at key_off (nothing starts any more, the sounding ramp is released). Host pointers. */
typedef struct ca_ramp_spec { /* one entry of SoundEngine::RampSpecs (include/soundengine.h:1030-1122) */
  float from_hz, to_hz;    /* LogRamp::setup(freq_to_angle_increment(from), freq_to_angle_increment(to), duration, start_sample, itp) */
  float duration;          /* samples, > 0 */
  float start_sample;      /* initial ramp position, 0 <= start_sample <= duration */
  int32_t itp;             /* ca_itp: interpolation of the emitted frequency (LINEAR and the 21 easings) */
  float volume;            /* setVolume */
  int32_t silence_follows; /* silenceFollows: the articulative pause applies after this spec */
} ca_ramp_spec;
typedef struct ca_sound_engine_desc {
  int64_t n_voices;
  const int32_t* spec_begin;   /* [voices + 1]: the specs of voice v are specs[spec_begin[v] .. spec_begin[v+1]), at most 30 */
  const ca_ramp_spec* specs;
  const int64_t* key_on;       /* nullable [voices] (0) */
  const int64_t* key_off;      /* nullable [voices]; < 0: never */
  const float* velocity;       /* nullable [voices] (1): the voice's volume (base volume 1, include/audioelement.h:1278) */
  const float* pan;            /* nullable [voices], CA_MIX_STEREO (include/stereo.h:18-27; gen.voice.h:992-1008) */
  int64_t n_samples;
  int32_t sample_rate;
  int32_t mix;                 /* CA_MIX_NONE: h_out [voices][samples]; CA_MIX_MONO: [samples]; CA_MIX_STEREO: [samples][2] */
  int32_t xfade_len;           /* setEnvelopeCharacTime: attack = release of every ramp, samples (at least 1 ms) */
  int32_t articulative_pause;  /* articulative_pause_length, samples */
  int32_t loudness_on;         /* 1: LoudnessVolumeAdjusted<Sine> as in the reference; 0: plain sine */
  int32_t loudness_ref_index;  /* setLoudnessParams (gen.voice.h:934-937): index of the reference frequency in the ISO 226 table */
  float loudness_compensation; /* log_ratio in [0,1] */
  float loudness_level;        /* phons (20 .. 100) */
  float sine_gain;             /* Mix gain of the sine (SetGains, gen.voice.h:951-956) */
} ca_sound_engine_desc;
int ca_sound_engine_render_host(ca_ctx* ctx, const ca_sound_engine_desc* desc, float* h_out);

/* The ramp specs of one note as SoundEngine::initialize_{birds,robot,sweep} + do_initialize produce them (include/soundengine.h
 * :496-736, :883-1003): the Markov chain is run on the host at note-on, min_path_length normalised steps then additional_tries plain
 * steps. mode: 0 BIRDS, 1 ROBOTS, 2 SWEEP (SoundEngineMode). The chain class and mersenne<> are in cpp.algorithms: the walk and the
 * draws follow oracle/SPEC.md S15 with a generator seeded by `seed`. Returns the number of specs written (<= 30), negative on error. */
typedef struct ca_sound_engine_gen {
  int32_t mode;
  int32_t sample_rate;
  uint32_t seed;
  float base_freq;                    /* the note's frequency */
  float length_ms;                    /* set_length */
  float length_exp_min, length_exp_max; /* set_length_exp */
  float freq_scatter;                 /* BIRDS, ROBOTS */
  float phase_ratio1;                 /* ROBOTS, SWEEP: start position = phase_ratio1 * duration */
  int32_t itp;                        /* set_itp */
  int32_t xfade_freq;                 /* BIRDS: FreqXfade 0 No, 1 NonTrivial, 2 All */
  int32_t freq_xfade_samples;         /* BIRDS: set_freq_xfade */
  int32_t freq_itp;                   /* BIRDS: set_freq_interpolation */
  float d1, d2, har_att;              /* ROBOTS */
  float low_freq, high_freq;          /* SWEEP */
  int32_t start_node, pre_tries, min_path_length, additional_tries;
} ca_sound_engine_gen;
int ca_sound_engine_specs(const ca_sound_engine_gen* gen, ca_ramp_spec* specs, int32_t capacity);

/* Harmonics of the reference's non-sine oscillators (FOscillator: saw, square, triangle, include/audioelement.h:1612-1675; wave
 * shapes include/sound.functions.h:86-139) for additive synthesis on the oscillator bank, as the author plans
 * (source/rt.resynth.txt:17-23): volume[h] and phase[h] (units of pi rad) of harmonic h+1 for ca_oscbank_desc.partial_volume /
 * partial_phase, so that sum_h volume[h] sin(pi (h+1) a + pi phase[h]) is the Fourier series of the wave over the angle a in [0,2).
 * The bank's aliasing gain fades harmonics between sr/8 and sr/2 (include/audioelement.h:466-483). waveform: 0 saw, 1 square, 2 triangle. */
enum ca_waveform { CA_WAVE_SAW = 0, CA_WAVE_SQUARE = 1, CA_WAVE_TRIANGLE = 2 };
int ca_waveform_harmonics(int32_t waveform, int32_t n_harmonics, float* volume, float* phase);

/* ---- FFT vocoder --------------------------------------------------------------------------------------------------
 * Stands in for Vocoder / FFTModulator / FFTCarrier (source/rt.resynth.lib.vocoder.cpp:100-257, :396-551, :734-856): the
 * modulator's band amplitudes (PeriodicFFT<SqMagFftOperation<Gaussian>>, window modulator_window_size_seconds, hop = stride)
 * scale the carrier's spectrum band by band (PeriodicFFT<FftOperation>, window 2*stride, hop stride), the frames go back
 * through the inverse real FFT and consecutive frames are cross-faded; output = vol_modulator * modulator + vol_carrier *
 * carrier + vol_vocoder * vocoded, with the vocoded signal 2*stride-1 samples late, as in the reference. Fields follow
 * SetupParams (:5-44; defaults source/rt.resynth.lib.cpp:986-990) and Vocoder::Volumes (:741-745). Host pointers,
 * [n_streams][n_samples] each; stride = good_stride(stride_seconds) (:84-92, odd). */
typedef struct ca_vocoder_params {
  int32_t sample_rate;
  float modulator_window_size_seconds; /* 0.10 */
  float stride_seconds;                /* 0.005 */
  int32_t count_bands;                 /* 5; at most 50 (max_count_bands, :98) */
  float min_freq;                      /* 100 */
  float max_freq;                      /* 20000 */
  float vol_modulator, vol_carrier, vol_vocoder;
} ca_vocoder_params;
int ca_vocoder_host(ca_ctx* ctx, const ca_vocoder_params* params, const float* h_modulator, const float* h_carrier,
                    int64_t n_samples, int64_t n_streams, float* h_out);

/* measured FP32 FMA throughput of this GPU in TFLOP/s (dependent-free FFMA chains, `ms` of work); used as the
 * denominator of the synthesis roofline */
int ca_measure_fp32_peak(ca_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CPP_AUDIO_B200_H */
