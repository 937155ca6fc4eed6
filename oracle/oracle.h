/* oracle/oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the acoustic-scoring hot path of VoiceBridge's
 * Kaldi snapshot (paths below are relative to /root/reference/kaldi-master/src).  It exists so
 * that the CUDA path in voicebridge_b200/csrc can be checked on a box where /root/reference is
 * absent.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here against the
 * reference's own code compiled from /root/reference (oracle/_ref/libvbref.so, oracle/ref_driver.cc)
 * and against tests/golden/ (HTK golden MFCCs of feat/test_data/test.wav, fixtures dumped from the
 * compiled reference by tests/golden/make_golden.py; Kaldi-pitch dumps tests/golden/pitch_golden.npz by
 * tests/golden/make_pitch_golden.py).
 */
#ifndef VB_ORACLE_H_
#define VB_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirror of MfccOptions + FrameExtractionOptions + MelBanksOptions
 * (feat/feature-mfcc.h:38-78, feat/feature-window.h:35-101, feat/mel-computations.h:43-74).
 * Same field order/layout as vbgpu_mfcc_opts in include/vbgpu.h. */
typedef struct orc_mfcc_opts {
  float samp_freq;            /* 16000 */
  float frame_shift_ms;       /* 10 */
  float frame_length_ms;      /* 25 */
  float dither;               /* Kaldi default 1.0; the oracle only supports 0 */
  float preemph_coeff;        /* 0.97 */
  int32_t remove_dc_offset;   /* 1 */
  int32_t window_type;        /* 0 povey, 1 hamming, 2 hanning, 3 rectangular, 4 blackman */
  int32_t round_to_power_of_two; /* 1 (0 unsupported) */
  float blackman_coeff;       /* 0.42 */
  int32_t snip_edges;         /* 1 */
  int32_t num_bins;           /* 23 */
  float low_freq;             /* 20 */
  float high_freq;            /* 0 */
  float vtln_low;             /* 100 */
  float vtln_high;            /* -500 */
  int32_t htk_mode;           /* 0 */
  int32_t num_ceps;           /* 13 */
  int32_t use_energy;         /* Kaldi default 1 */
  float energy_floor;         /* 0 */
  int32_t raw_energy;         /* 1 */
  float cepstral_lifter;      /* 22 */
  int32_t htk_compat;         /* 0 */
} orc_mfcc_opts;

void orc_mfcc_opts_default(orc_mfcc_opts *o);

/* feature-window.cc:28-87 */
int32_t orc_window_shift(const orc_mfcc_opts *o);
int32_t orc_window_size(const orc_mfcc_opts *o);
int32_t orc_padded_window_size(const orc_mfcc_opts *o);
int64_t orc_first_sample_of_frame(int32_t frame, const orc_mfcc_opts *o);
int32_t orc_num_frames(int64_t num_samples, const orc_mfcc_opts *o);

/* Tables (feature-window.cc:109-131, mel-computations.cc:33-144,255-261, matrix-functions.cc:592-608).
 * mel: offsets[num_bins], lens[num_bins], weights[num_bins * (Npad/2)] (row b holds lens[b] weights).
 * Return 0 on success, <0 on bad options. */
int orc_window_table(const orc_mfcc_opts *o, float *window /* WindowSize */);
int orc_mel_banks(const orc_mfcc_opts *o, float vtln_warp, int32_t *offsets, int32_t *lens, float *weights);
void orc_dct_matrix(int32_t num_ceps, int32_t num_bins, float *dct /* num_ceps x num_bins */);
void orc_lifter_coeffs(float Q, int32_t num_ceps, float *coeffs);

/* OfflineFeatureTpl<MfccComputer>::Compute (feature-common-inl.h:61-98) for one utterance.
 * wave: float copies of the int16 samples (wave-reader.cc:302-309).  Returns #frames or <0. */
int orc_mfcc_compute(const orc_mfcc_opts *o, const float *wave, int64_t num_samples, float vtln_warp,
                     float *out, int32_t out_stride);

/* transform/cmvn.cc:30-113.  stats: double[2][D+1] row-major. */
void orc_cmvn_acc(const float *feats, int32_t T, int32_t D, int32_t stride, double *stats);
int orc_cmvn_apply(const double *stats, int32_t D, int32_t norm_vars, float *feats, int32_t T, int32_t stride);

/* feat/feature-functions.cc:54-111,160-171 and :205-226 */
int orc_delta_scales(int32_t order, int32_t window, float *scales /* (order+1) x (2*order*window+1) */, int32_t *lens);
void orc_deltas(int32_t order, int32_t window, const float *in, int32_t T, int32_t D, int32_t in_stride,
                float *out, int32_t out_stride);
void orc_splice(const float *in, int32_t T, int32_t D, int32_t in_stride, int32_t left, int32_t right,
                float *out, int32_t out_stride);
/* transform-feats.cpp:95-107 : cols==D linear, cols==D+1 affine.  Returns 0, or -1 on bad dims. */
int orc_transform(const float *in, int32_t T, int32_t D, int32_t in_stride, const float *mat, int32_t rows,
                  int32_t cols, float *out, int32_t out_stride);

/* gmm/diag-gmm.cc:114-152.  Returns number of +-inf gconsts, or -1 on NaN. */
int orc_gconsts(int32_t M, int32_t D, const float *weights, const float *means_invvars, const float *inv_vars,
                float *gconsts);

/* Dense all-pdf scoring with the decodable's arithmetic (decodable-am-diag-gmm.cc:28-72,
 * kaldi-vector.cc:757-775).  Flattened model: pdf p owns Gaussians [pdf_offsets[p], pdf_offsets[p+1]).
 * out[t*out_stride + p] = LSE_p(x_t).  Returns 0, or -2 if a NaN/Inf was produced. */
int orc_gmm_loglikes(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts,
                     const float *means_invvars, const float *inv_vars, const float *feats, int32_t T,
                     int32_t stride, float prune, float *out, int32_t out_stride);

/* AccumAmDiagGmm::AccumulateForGmm over an alignment (mle-am-diag-gmm.cc:69-79, mle-diag-gmm.cc:171-204,
 * diag-gmm.cc:601-615, kaldi-vector.cc:852-859).  occ[N], mean_acc[N*D], var_acc[N*D] are ADDED to.
 * weights may be NULL (=1.0).  tot_like/tot_frames are added to.  Returns 0 or -2 on NaN/Inf. */
int orc_acc_ali(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts,
                const float *means_invvars, const float *inv_vars, const float *feats, int32_t T,
                int32_t stride, const int32_t *pdf_ids, const float *weights, double *occ, double *mean_acc,
                double *var_acc, double *tot_like, double *tot_frames);

/* AccumulateForGmmTwofeats (mle-am-diag-gmm.cc:81-97): posteriors from feats1, stats from feats2. */
int orc_acc_ali_twofeats(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts,
                         const float *means_invvars, const float *inv_vars, const float *feats1,
                         const float *feats2, int32_t T, int32_t stride, const int32_t *pdf_ids,
                         const float *weights, double *occ, double *mean_acc, double *var_acc,
                         double *tot_like, double *tot_frames);

/* OfflineFeatureTpl<FbankComputer>::ComputeFeatures (feat/feature-fbank.cc:73-123): frame / mel options from o
 * (num_ceps and cepstral_lifter unused); out has num_bins columns, +1 with use_energy (first, or last with htk_compat). */
int orc_fbank_compute(const orc_mfcc_opts *o, int32_t use_log_fbank, int32_t use_power, const float *wave,
                      int64_t n_samp, float vtln_warp, float *out, int32_t out_stride);

/* OfflineFeatureTpl<PlpComputer>::ComputeFeatures (feat/feature-plp.cc:113-188, mel-computations.cc:269-340). */
int orc_plp_compute(const orc_mfcc_opts *o, int32_t lpc_order, float compress_factor, float cepstral_scale,
                    const float *wave, int64_t n_samp, float vtln_warp, float *out, int32_t out_stride);

/* FmllrDiagGmmAccs::AccumulateForGmm over an alignment (transform/fmllr-diag-gmm.cc:30-45,110-121,562-583; driver
 * gmm-est-fmllr.cpp:40-55), update_type "full".  beta, K[D*(D+1)], G[D*(D+1)(D+2)/2] (SpMatrix packing) are ADDED to;
 * tot_like += sum of frame log-likelihoods.  Returns 0, -1 on a bad pdf id, -2 on NaN/Inf. */
int orc_fmllr_acc(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts,
                  const float *means_invvars, const float *inv_vars, const float *feats, int32_t T,
                  int32_t stride, const int32_t *pdf_ids, const float *weights, double *beta, double *K,
                  double *G, double *tot_like);

/* MlltAccs::AccumulateFromGmm over an alignment (transform/mllt.cc:131-170, driver gmm-acc-mllt.cpp:100-112), rand_prune = 0.
 * beta, G[D * D(D+1)/2] (SpMatrix packing) ADDED to; tot_like += sum of loglike * weight. */
int orc_mllt_acc(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts, const float *means_invvars,
                 const float *inv_vars, const float *feats, int32_t T, int32_t stride, const int32_t *pdf_ids,
                 const float *weights, double *beta, double *G, double *tot_like);

/* DiagGmm::ComponentPosteriors of each frame's aligned pdf, scaled by the frame weight (gmm-post-to-gpost). */
int orc_component_posteriors(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *gconsts,
                             const float *means_invvars, const float *inv_vars, const float *feats, int32_t T,
                             int32_t stride, const int32_t *pdf_ids, const float *weights, float *post_out,
                             float *loglikes);

/* Kaldi pitch (feat/pitch-functions.{h,cc}, feat/resample.{h,cc}) — the offline ComputeKaldiPitch path
 * (frames_per_chunk = 0, simulate_first_pass_online = false, nccf_ballast_online = false, max_frames_latency = 0).
 * Same layout as vbgpu_pitch_opts / vbgpu_process_pitch_opts (include/vbgpu.h). */
typedef struct orc_pitch_opts {
  float samp_freq;              /* 16000  PitchExtractionOptions, pitch-functions.h:103-123 */
  float frame_shift_ms;         /* 10 */
  float frame_length_ms;        /* 25 */
  float preemph_coeff;          /* 0 */
  float min_f0;                 /* 50 */
  float max_f0;                 /* 400 */
  float soft_min_f0;            /* 10 */
  float penalty_factor;         /* 0.1 */
  float lowpass_cutoff;         /* 1000 */
  float resample_freq;          /* 4000 */
  float delta_pitch;            /* 0.005 */
  float nccf_ballast;           /* 7000 */
  int32_t lowpass_filter_width; /* 1 */
  int32_t upsample_filter_width;/* 5 */
  int32_t recompute_frame;      /* 500 */
  int32_t snip_edges;           /* 1 */
} orc_pitch_opts;

typedef struct orc_process_pitch_opts {
  float pitch_scale;                  /* 2   ProcessPitchOptions, pitch-functions.h:241-255 */
  float pov_scale;                    /* 2 */
  float pov_offset;                   /* 0 */
  float delta_pitch_scale;            /* 10 */
  float delta_pitch_noise_stddev;     /* 0.005 (Kaldi); the oracle only supports 0 (the reference draws from rand()) */
  int32_t normalization_left_context; /* 75 */
  int32_t normalization_right_context;/* 75 */
  int32_t delta_window;               /* 2 */
  int32_t delay;                      /* 0 */
  int32_t add_pov_feature;            /* 1 */
  int32_t add_normalized_log_pitch;   /* 1 */
  int32_t add_delta_pitch;            /* 1 */
  int32_t add_raw_log_pitch;          /* 0 */
} orc_process_pitch_opts;

void orc_pitch_opts_default(orc_pitch_opts *o);
void orc_process_pitch_opts_default(orc_process_pitch_opts *o);
/* Number of output frames of ComputeKaldiPitch for n_samp input samples (pitch-functions.cc:768-792 after InputFinished). */
/* DownsampleWaveForm (feat/resample.cc:368-376); out nullable; returns the number of output samples, <0 on bad arguments */
int64_t orc_downsample_waveform(float orig_freq, float new_freq, const float *wave, int64_t n, float *out);
int32_t orc_pitch_num_frames(const orc_pitch_opts *o, int64_t n_samp);
/* ComputeKaldiPitch (pitch-functions.cc:1291-1325): out[T][2] = (NCCF at the chosen lag, pitch in Hz).  Returns T or <0.
 * The Viterbi uses the reference's own exhaustive search (pitch_use_naive_search, pitch-functions.cc:334-348), which its
 * interval-tightening search (349-470) reproduces. */
int orc_pitch_compute(const orc_pitch_opts *o, const float *wave, int64_t n_samp, float *out, int32_t out_stride);
/* ProcessPitch (pitch-functions.cc:1581-1595, 1414-1567): in[T][2] -> out[T + delay][dim]; returns rows or <0. */
int orc_process_pitch(const orc_process_pitch_opts *o, const float *in, int32_t T, int32_t in_stride, float *out,
                      int32_t out_stride);

#ifdef __cplusplus
}
#endif
#endif
