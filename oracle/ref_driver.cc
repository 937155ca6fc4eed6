// oracle/ref_driver.cc — TEST INFRASTRUCTURE ONLY.
//
// Thin C-ABI driver (our own code) over the reference's UNMODIFIED Kaldi sources, which are compiled
// where they lie under /root/reference by oracle/Makefile into oracle/_ref/libkaldi_ref.a.  Built as
// oracle/_ref/libvbref.so.  Every ref_* entry point calls the reference's own class/function for that
// step, so tests can (a) pin the plain-C restatement in oracle/oracle.c and (b) compare the CUDA path
// with the real reference.  It is also the "reference" CPU baseline that bench.py times.
//
// Signatures mirror oracle/oracle.h (orc_* -> ref_*); the model is passed flattened exactly as the
// GPU library takes it (include/vbgpu.h).

#include <cstring>
#include <sstream>
#include <thread>
#include <vector>

#include "base/kaldi-common.h"
#include "feat/feature-functions.h"
#include "feat/feature-fbank.h"
#include "feat/feature-mfcc.h"
#include "feat/feature-plp.h"
#include "feat/mel-computations.h"
#include "feat/pitch-functions.h"
#include "feat/resample.h"
#include "feat/wave-reader.h"
#include "gmm/am-diag-gmm.h"
#include "gmm/decodable-am-diag-gmm.h"
#include "gmm/mle-am-diag-gmm.h"
#include "matrix/kaldi-matrix.h"
#include "transform/cmvn.h"
#include "transform/fmllr-diag-gmm.h"
#include "transform/mllt.h"
#include "hmm/transition-model.h"
#include "matrix/compressed-matrix.h"
#include "tree/context-dep.h"
#include "util/kaldi-holder.h"

#include "oracle.h"

using namespace kaldi;

namespace {

MfccOptions ToKaldi(const orc_mfcc_opts *o) {
  MfccOptions m;
  m.frame_opts.samp_freq = o->samp_freq;
  m.frame_opts.frame_shift_ms = o->frame_shift_ms;
  m.frame_opts.frame_length_ms = o->frame_length_ms;
  m.frame_opts.dither = o->dither;
  m.frame_opts.preemph_coeff = o->preemph_coeff;
  m.frame_opts.remove_dc_offset = o->remove_dc_offset != 0;
  static const char *kWin[] = {"povey", "hamming", "hanning", "rectangular", "blackman"};
  m.frame_opts.window_type = kWin[o->window_type];
  m.frame_opts.round_to_power_of_two = o->round_to_power_of_two != 0;
  m.frame_opts.blackman_coeff = o->blackman_coeff;
  m.frame_opts.snip_edges = o->snip_edges != 0;
  m.mel_opts.num_bins = o->num_bins;
  m.mel_opts.low_freq = o->low_freq;
  m.mel_opts.high_freq = o->high_freq;
  m.mel_opts.vtln_low = o->vtln_low;
  m.mel_opts.vtln_high = o->vtln_high;
  m.mel_opts.htk_mode = o->htk_mode != 0;
  m.num_ceps = o->num_ceps;
  m.use_energy = o->use_energy != 0;
  m.energy_floor = o->energy_floor;
  m.raw_energy = o->raw_energy != 0;
  m.cepstral_lifter = o->cepstral_lifter;
  m.htk_compat = o->htk_compat != 0;
  return m;
}

void ToMatrix(const float *data, int32 T, int32 D, int32 stride, Matrix<BaseFloat> *m) {
  m->Resize(T, D, kUndefined);
  for (int32 t = 0; t < T; t++) std::memcpy(m->RowData(t), data + (size_t)t * stride, sizeof(float) * D);
}
void FromMatrix(const Matrix<BaseFloat> &m, float *data, int32 stride) {
  for (int32 t = 0; t < m.NumRows(); t++)
    std::memcpy(data + (size_t)t * stride, m.RowData(t), sizeof(float) * m.NumCols());
}

struct RefModel {
  AmDiagGmm am;
  std::vector<int32> offsets;
};

}  // namespace

extern "C" {

// WaveData::Read on an in-memory image.  Returns the number of samples per channel (<0 on KALDI_ERR); data (nullable)
// receives [channels x samples] floats, cap = its capacity in floats.
int64_t ref_wave_read(const char *bytes, int64_t n, float *samp_freq, int32_t *channels, float *data, int64_t cap) {
  try {
    std::istringstream is(std::string(bytes, bytes + n), std::ios::binary);
    WaveData w;
    w.Read(is);
    *samp_freq = w.SampFreq();
    *channels = w.Data().NumRows();
    const int64_t ns = w.Data().NumCols();
    if (data) {
      if ((int64_t)*channels * ns > cap) return -2;
      for (int32 c = 0; c < *channels; c++)
        for (int64_t i = 0; i < ns; i++) data[c * ns + i] = w.Data()(c, i);
    }
    return ns;
  } catch (const std::exception &) {
    return -1;
  }
}

int32_t ref_num_frames(int64_t n, const orc_mfcc_opts *o) {
  MfccOptions m = ToKaldi(o);
  return NumFrames(n, m.frame_opts);
}

// MelBanks (mel-computations.cc:33-144) -> same table layout as orc_mel_banks.
int ref_mel_banks(const orc_mfcc_opts *o, float vtln_warp, int32_t *offsets, int32_t *lens, float *weights) {
  try {
    MfccOptions m = ToKaldi(o);
    MelBanks mb(m.mel_opts, m.frame_opts, vtln_warp);
    // bins_ is private: recover it by probing MelBanks::Compute with unit power spectra.
    int32 nfft = m.frame_opts.PaddedWindowSize() / 2, B = mb.NumBins();
    std::memset(weights, 0, sizeof(float) * (size_t)B * nfft);
    std::vector<float> dense((size_t)B * nfft, 0.0f);
    Vector<BaseFloat> ps(nfft + 1), me(B);
    for (int32 i = 0; i < nfft; i++) {
      ps.SetZero();
      ps(i) = 1.0;
      mb.Compute(ps, &me);
      for (int32 b = 0; b < B; b++) dense[(size_t)b * nfft + i] = me(b);
    }
    for (int32 b = 0; b < B; b++) {
      int32 first = -1, last = -1;
      for (int32 i = 0; i < nfft; i++)
        if (dense[(size_t)b * nfft + i] != 0.0f) { if (first < 0) first = i; last = i; }
      offsets[b] = first;
      lens[b] = last + 1 - first;
      for (int32 i = 0; i < lens[b]; i++) weights[(size_t)b * nfft + i] = dense[(size_t)b * nfft + first + i];
    }
    return 0;
  } catch (const std::exception &) { return -1; }
}

int ref_window_table(const orc_mfcc_opts *o, float *window) {
  try {
    MfccOptions m = ToKaldi(o);
    FeatureWindowFunction w(m.frame_opts);
    for (int32 i = 0; i < w.window.Dim(); i++) window[i] = w.window(i);
    return 0;
  } catch (const std::exception &) { return -1; }
}

// OfflineFeatureTpl<MfccComputer>::ComputeFeatures (feature-common-inl.h:29-98)
int ref_mfcc_compute(const orc_mfcc_opts *o, const float *wave, int64_t n, float vtln_warp, float *out,
                     int32_t out_stride) {
  try {
    Mfcc mfcc(ToKaldi(o));
    SubVector<BaseFloat> w(const_cast<float *>(wave), (MatrixIndexT)n);
    Matrix<BaseFloat> feats;
    mfcc.ComputeFeatures(w, o->samp_freq, vtln_warp, &feats);
    FromMatrix(feats, out, out_stride);
    return feats.NumRows();
  } catch (const std::exception &) { return -1; }
}

int ref_fbank_compute(const orc_mfcc_opts *o, int32_t use_log_fbank, int32_t use_power, const float *wave, int64_t n,
                      float vtln_warp, float *out, int32_t out_stride) {
  try {
    MfccOptions m = ToKaldi(o);
    FbankOptions f;
    f.frame_opts = m.frame_opts;
    f.mel_opts = m.mel_opts;
    f.use_energy = m.use_energy;
    f.energy_floor = m.energy_floor;
    f.raw_energy = m.raw_energy;
    f.htk_compat = m.htk_compat;
    f.use_log_fbank = use_log_fbank != 0;
    f.use_power = use_power != 0;
    Fbank fbank(f);
    SubVector<BaseFloat> w(const_cast<float *>(wave), (MatrixIndexT)n);
    Matrix<BaseFloat> feats;
    fbank.ComputeFeatures(w, o->samp_freq, vtln_warp, &feats);
    FromMatrix(feats, out, out_stride);
    return feats.NumRows();
  } catch (const std::exception &) { return -1; }
}

int ref_plp_compute(const orc_mfcc_opts *o, int32_t lpc_order, float compress_factor, float cepstral_scale,
                    const float *wave, int64_t n, float vtln_warp, float *out, int32_t out_stride) {
  try {
    MfccOptions m = ToKaldi(o);
    PlpOptions p;
    p.frame_opts = m.frame_opts;
    p.mel_opts = m.mel_opts;
    p.lpc_order = lpc_order;
    p.num_ceps = m.num_ceps;
    p.use_energy = m.use_energy;
    p.energy_floor = m.energy_floor;
    p.raw_energy = m.raw_energy;
    p.compress_factor = compress_factor;
    p.cepstral_lifter = (int32)m.cepstral_lifter;
    p.cepstral_scale = cepstral_scale;
    p.htk_compat = m.htk_compat;
    Plp plp(p);
    SubVector<BaseFloat> w(const_cast<float *>(wave), (MatrixIndexT)n);
    Matrix<BaseFloat> feats;
    plp.ComputeFeatures(w, o->samp_freq, vtln_warp, &feats);
    FromMatrix(feats, out, out_stride);
    return feats.NumRows();
  } catch (const std::exception &) { return -1; }
}

void ref_cmvn_acc(const float *feats, int32_t T, int32_t D, int32_t stride, double *stats) {
  Matrix<BaseFloat> f;
  ToMatrix(feats, T, D, stride, &f);
  Matrix<double> s(2, D + 1);
  AccCmvnStats(f, NULL, &s);
  for (int r = 0; r < 2; r++)
    for (int d = 0; d <= D; d++) stats[r * (D + 1) + d] += s(r, d);
}

int ref_cmvn_apply(const double *stats, int32_t D, int32_t norm_vars, float *feats, int32_t T, int32_t stride) {
  try {
    Matrix<BaseFloat> f;
    ToMatrix(feats, T, D, stride, &f);
    Matrix<double> s(2, D + 1);
    for (int r = 0; r < 2; r++)
      for (int d = 0; d <= D; d++) s(r, d) = stats[r * (D + 1) + d];
    ApplyCmvn(s, norm_vars != 0, &f);
    FromMatrix(f, feats, stride);
    return 0;
  } catch (const std::exception &) { return -1; }
}

void ref_deltas(int32_t order, int32_t window, const float *in, int32_t T, int32_t D, int32_t in_stride, float *out,
                int32_t out_stride) {
  Matrix<BaseFloat> f, g;
  ToMatrix(in, T, D, in_stride, &f);
  DeltaFeaturesOptions opts(order, window);
  ComputeDeltas(opts, f, &g);
  FromMatrix(g, out, out_stride);
}

void ref_splice(const float *in, int32_t T, int32_t D, int32_t in_stride, int32_t left, int32_t right, float *out,
                int32_t out_stride) {
  Matrix<BaseFloat> f, g;
  ToMatrix(in, T, D, in_stride, &f);
  SpliceFrames(f, left, right, &g);
  FromMatrix(g, out, out_stride);
}

// The arithmetic of transform-feats (featbin/transform-feats.cc:95-107, identical in the VoiceBridge copy).
int ref_transform(const float *in, int32_t T, int32_t D, int32_t in_stride, const float *mat, int32_t rows,
                  int32_t cols, float *out, int32_t out_stride) {
  Matrix<BaseFloat> feat, trans(rows, cols);
  ToMatrix(in, T, D, in_stride, &feat);
  for (int r = 0; r < rows; r++) std::memcpy(trans.RowData(r), mat + (size_t)r * cols, sizeof(float) * cols);
  Matrix<BaseFloat> feat_out(T, rows);
  if (cols == D) {
    feat_out.AddMatMat(1.0, feat, kNoTrans, trans, kTrans, 0.0);
  } else if (cols == D + 1) {
    SubMatrix<BaseFloat> linear_part(trans, 0, rows, 0, D);
    feat_out.AddMatMat(1.0, feat, kNoTrans, linear_part, kTrans, 0.0);
    Vector<BaseFloat> offset(rows);
    offset.CopyColFromMat(trans, D);
    feat_out.AddVecToRows(1.0, offset);
  } else {
    return -1;
  }
  FromMatrix(feat_out, out, out_stride);
  return 0;
}

// Build an AmDiagGmm from (weights, means, inverse variances); gconsts via DiagGmm::ComputeGconsts.
void *ref_model_create(int32_t P, int32_t D, const int32_t *pdf_offsets, const float *weights, const float *means,
                       const float *inv_vars) {
  try {
    RefModel *rm = new RefModel;
    rm->offsets.assign(pdf_offsets, pdf_offsets + P + 1);
    for (int32 p = 0; p < P; p++) {
      int32 g0 = pdf_offsets[p], M = pdf_offsets[p + 1] - g0;
      DiagGmm gmm(M, D);
      Vector<BaseFloat> w(M);
      Matrix<BaseFloat> mu(M, D), iv(M, D);
      for (int32 m = 0; m < M; m++) {
        w(m) = weights[g0 + m];
        std::memcpy(mu.RowData(m), means + (size_t)(g0 + m) * D, sizeof(float) * D);
        std::memcpy(iv.RowData(m), inv_vars + (size_t)(g0 + m) * D, sizeof(float) * D);
      }
      gmm.SetWeights(w);
      gmm.SetInvVarsAndMeans(iv, mu);
      gmm.ComputeGconsts();
      rm->am.AddPdf(gmm);
    }
    return rm;
  } catch (const std::exception &) { return NULL; }
}

void ref_model_destroy(void *h) { delete static_cast<RefModel *>(h); }

// Export what DiagGmm holds (diag-gmm.h:174-180): gconsts[N], means_invvars[N*D], inv_vars[N*D].
void ref_model_get(void *h, float *gconsts, float *means_invvars, float *inv_vars) {
  RefModel *rm = static_cast<RefModel *>(h);
  int32 D = rm->am.Dim();
  for (int32 p = 0; p < rm->am.NumPdfs(); p++) {
    const DiagGmm &g = rm->am.GetPdf(p);
    int32 g0 = rm->offsets[p];
    for (int32 m = 0; m < g.NumGauss(); m++) {
      gconsts[g0 + m] = g.gconsts()(m);
      std::memcpy(means_invvars + (size_t)(g0 + m) * D, g.means_invvars().RowData(m), sizeof(float) * D);
      std::memcpy(inv_vars + (size_t)(g0 + m) * D, g.inv_vars().RowData(m), sizeof(float) * D);
    }
  }
}

// Dense T x P scoring through the decodable's own code path
// (DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased, decodable-am-diag-gmm.cc:28-72).
int ref_gmm_loglikes(void *h, const float *feats, int32_t T, int32_t stride, float prune, float *out,
                     int32_t out_stride) {
  try {
    RefModel *rm = static_cast<RefModel *>(h);
    Matrix<BaseFloat> f;
    ToMatrix(feats, T, rm->am.Dim(), stride, &f);
    DecodableAmDiagGmmUnmapped dec(rm->am, f, prune);
    int32 P = rm->am.NumPdfs();
    for (int32 t = 0; t < T; t++)
      for (int32 p = 0; p < P; p++) out[(size_t)t * out_stride + p] = dec.LogLikelihood(t, p + 1);  // one-based index
    return 0;
  } catch (const std::exception &) { return -2; }
}

// Dense scoring in the batched matrix form (DiagGmm::LogLikelihoods(Matrix), diag-gmm.cc:546-562, then
// LogSumExp per row) — the fastest way the reference's own code can produce the T x P matrix on CPU.
int ref_gmm_loglikes_matrix(void *h, const float *feats, int32_t T, int32_t stride, float *out, int32_t out_stride) {
  try {
    RefModel *rm = static_cast<RefModel *>(h);
    Matrix<BaseFloat> f, ll;
    ToMatrix(feats, T, rm->am.Dim(), stride, &f);
    int32 P = rm->am.NumPdfs();
    for (int32 p = 0; p < P; p++) {
      rm->am.GetPdf(p).LogLikelihoods(f, &ll);
      for (int32 t = 0; t < T; t++) out[(size_t)t * out_stride + p] = ll.Row(t).LogSumExp();
    }
    return 0;
  } catch (const std::exception &) { return -2; }
}

static int RefAcc(void *h, const float *feats1, const float *feats2, int32_t T, int32_t stride,
                  const int32_t *pdf_ids, const float *weights, double *occ, double *mean_acc, double *var_acc,
                  double *tot_like, double *tot_frames) {
  try {
    RefModel *rm = static_cast<RefModel *>(h);
    int32 D = rm->am.Dim();
    Matrix<BaseFloat> f1, f2;
    ToMatrix(feats1, T, D, stride, &f1);
    if (feats2) ToMatrix(feats2, T, D, stride, &f2);
    AccumAmDiagGmm acc;
    acc.Init(rm->am, kGmmAll);
    for (int32 t = 0; t < T; t++) {
      BaseFloat w = weights ? weights[t] : 1.0;
      if (feats2) acc.AccumulateForGmmTwofeats(rm->am, f1.Row(t), f2.Row(t), pdf_ids[t], w);
      else acc.AccumulateForGmm(rm->am, f1.Row(t), pdf_ids[t], w);
    }
    for (int32 p = 0; p < rm->am.NumPdfs(); p++) {
      const AccumDiagGmm &a = acc.GetAcc(p);
      int32 g0 = rm->offsets[p];
      for (int32 m = 0; m < a.NumGauss(); m++) {
        occ[g0 + m] += a.occupancy()(m);
        for (int32 d = 0; d < D; d++) {
          mean_acc[(size_t)(g0 + m) * D + d] += a.mean_accumulator()(m, d);
          var_acc[(size_t)(g0 + m) * D + d] += a.variance_accumulator()(m, d);
        }
      }
    }
    *tot_like += acc.TotLogLike();
    *tot_frames += acc.TotCount();
    return 0;
  } catch (const std::exception &) { return -2; }
}

// AccumAmDiagGmm::AccumulateForGmm over an alignment (gmm-acc-stats-ali.cc:89-94).
int ref_acc_ali(void *h, const float *feats, int32_t T, int32_t stride, const int32_t *pdf_ids, const float *weights,
                double *occ, double *mean_acc, double *var_acc, double *tot_like, double *tot_frames) {
  return RefAcc(h, feats, NULL, T, stride, pdf_ids, weights, occ, mean_acc, var_acc, tot_like, tot_frames);
}
int ref_acc_ali_twofeats(void *h, const float *feats1, const float *feats2, int32_t T, int32_t stride,
                         const int32_t *pdf_ids, const float *weights, double *occ, double *mean_acc,
                         double *var_acc, double *tot_like, double *tot_frames) {
  return RefAcc(h, feats1, feats2, T, stride, pdf_ids, weights, occ, mean_acc, var_acc, tot_like, tot_frames);
}

// FmllrDiagGmmAccs::AccumulateForGmm over an alignment (gmm-est-fmllr.cpp:40-55); stats are ADDED to.
int ref_fmllr_acc(void *h, const float *feats, int32_t T, int32_t stride, const int32_t *pdf_ids, const float *weights,
                  double *beta, double *K, double *G, double *tot_like) {
  try {
    RefModel *rm = static_cast<RefModel *>(h);
    int32 D = rm->am.Dim();
    Matrix<BaseFloat> f;
    ToMatrix(feats, T, D, stride, &f);
    FmllrDiagGmmAccs accs(D);
    for (int32 t = 0; t < T; t++)
      *tot_like += accs.AccumulateForGmm(rm->am.GetPdf(pdf_ids[t]), f.Row(t), weights ? weights[t] : 1.0);
    // Flush the frame held back by the single-frame cache the way Update() does (its first statement).
    Matrix<BaseFloat> dummy(D, D + 1);
    dummy.SetUnit();
    FmllrOptions o;
    o.min_count = 1.0e30;  // Update() commits the pending frame, finds too little data and changes nothing
    accs.Update(o, &dummy, NULL, NULL);
    *beta += accs.beta_;
    const int32 np = (D + 1) * (D + 2) / 2;
    for (int32 i = 0; i < D; i++) {
      for (int32 k = 0; k <= D; k++) K[(size_t)i * (D + 1) + k] += accs.K_(i, k);
      for (int32 j = 0; j <= D; j++)
        for (int32 k = 0; k <= j; k++) G[(size_t)i * np + j * (j + 1) / 2 + k] += accs.G_[i](j, k);
    }
    return 0;
  } catch (const std::exception &) { return -2; }
}

// DiagGmm::ComponentPosteriors per frame of an alignment, scaled by the weight (gmm-post-to-gpost).
int ref_component_posteriors(void *h, const float *feats, int32_t T, int32_t stride, const int32_t *pdf_ids,
                             const float *weights, float *post_out, float *loglikes) {
  try {
    RefModel *rm = static_cast<RefModel *>(h);
    int32 D = rm->am.Dim();
    Matrix<BaseFloat> f;
    ToMatrix(feats, T, D, stride, &f);
    size_t o = 0;
    for (int32 t = 0; t < T; t++) {
      const DiagGmm &g = rm->am.GetPdf(pdf_ids[t]);
      Vector<BaseFloat> post(g.NumGauss());
      BaseFloat ll = g.ComponentPosteriors(f.Row(t), &post);
      post.Scale(weights ? weights[t] : 1.0);
      for (int32 m = 0; m < g.NumGauss(); m++) post_out[o + m] = post(m);
      if (loglikes) loglikes[t] = ll;
      o += g.NumGauss();
    }
    return 0;
  } catch (const std::exception &) { return -2; }
}

// MlltAccs::AccumulateFromGmm over an alignment (gmm-acc-mllt.cpp:100-112), rand_prune = 0; stats ADDED to.
int ref_mllt_acc(void *h, const float *feats, int32_t T, int32_t stride, const int32_t *pdf_ids, const float *weights,
                 double *beta, double *G, double *tot_like) {
  try {
    RefModel *rm = static_cast<RefModel *>(h);
    int32 D = rm->am.Dim();
    Matrix<BaseFloat> f;
    ToMatrix(feats, T, D, stride, &f);
    MlltAccs accs(D, 0.0);
    for (int32 t = 0; t < T; t++) {
      BaseFloat w = weights ? weights[t] : 1.0;
      *tot_like += accs.AccumulateFromGmm(rm->am.GetPdf(pdf_ids[t]), f.Row(t), w) * w;
    }
    *beta += accs.beta_;
    const int32 np = D * (D + 1) / 2;
    for (int32 j = 0; j < D; j++)
      for (int32 r = 0; r < D; r++)
        for (int32 c = 0; c <= r; c++) G[(size_t)j * np + r * (r + 1) / 2 + c] += accs.G_[j](r, c);
    return 0;
  } catch (const std::exception &) { return -2; }
}

// MlltAccs::Update (transform/mllt.cc:50-128) on given statistics, starting from the unit matrix.
int ref_mllt_update(int32_t D, double beta, const double *G, float *M, float *objf_impr, float *count) {
  try {
    std::vector<SpMatrix<double> > g(D);
    const int32 np = D * (D + 1) / 2;
    for (int32 j = 0; j < D; j++) {
      g[j].Resize(D);
      for (int32 r = 0; r < D; r++)
        for (int32 c = 0; c <= r; c++) g[j](r, c) = G[(size_t)j * np + r * (r + 1) / 2 + c];
    }
    Matrix<BaseFloat> m(D, D);
    m.SetUnit();
    BaseFloat impr = 0, cnt = 0;
    MlltAccs::Update(beta, g, &m, &impr, &cnt);
    for (int32 i = 0; i < D; i++)
      for (int32 k = 0; k < D; k++) M[(size_t)i * D + k] = m(i, k);
    if (objf_impr) *objf_impr = impr;
    if (count) *count = cnt;
    return 0;
  } catch (const std::exception &) { return -2; }
}

// FmllrDiagGmmAccs::Update (fmllr-diag-gmm.cc:124-170, default options) on given statistics -> D x (D+1) transform.
int ref_fmllr_update(int32_t D, double beta, const double *K, const double *G, float *xform, float *objf_impr,
                     float *count) {
  try {
    FmllrDiagGmmAccs accs(D);
    accs.beta_ = beta;
    const int32 np = (D + 1) * (D + 2) / 2;
    for (int32 i = 0; i < D; i++) {
      for (int32 k = 0; k <= D; k++) accs.K_(i, k) = K[(size_t)i * (D + 1) + k];
      for (int32 j = 0; j <= D; j++)
        for (int32 k = 0; k <= j; k++) accs.G_[i](j, k) = G[(size_t)i * np + j * (j + 1) / 2 + k];
    }
    Matrix<BaseFloat> m(D, D + 1);
    m.SetUnit();
    BaseFloat impr = 0, cnt = 0;
    accs.Update(FmllrOptions(), &m, &impr, &cnt);
    for (int32 i = 0; i < D; i++)
      for (int32 k = 0; k <= D; k++) xform[(size_t)i * (D + 1) + k] = m(i, k);
    if (objf_impr) *objf_impr = impr;
    if (count) *count = cnt;
    return 0;
  } catch (const std::exception &) { return -2; }
}

// ---- wire formats: the reference's own writers / readers on memory buffers (pins vbgpu_io_*) ---------------------------
static int64_t CopyOut(const std::string &s, char *buf, int64_t cap) {
  if ((int64_t)s.size() <= cap && buf) std::memcpy(buf, s.data(), s.size());
  return (int64_t)s.size();
}
// kind: 0 Matrix<float>, 1 Matrix<double>, 2 CompressedMatrix (auto: speech-feature method), 3 kTwoByteAuto, 4 kOneByteAuto;
// written through the table holder, i.e. with the \0B marker (KaldiObjectHolder::Write).
int64_t ref_io_write_matrix(const float *data, int32_t rows, int32_t cols, int32_t stride, int32_t kind, char *buf,
                            int64_t cap) {
  try {
    Matrix<BaseFloat> m;
    ToMatrix(data, rows, cols, stride, &m);
    std::ostringstream os(std::ios::binary);
    if (kind == 0) KaldiObjectHolder<Matrix<BaseFloat> >::Write(os, true, m);
    else if (kind == 1) KaldiObjectHolder<Matrix<double> >::Write(os, true, Matrix<double>(m));
    else {
      CompressionMethod cm = kind == 2 ? kAutomaticMethod : (kind == 3 ? kTwoByteAuto : kOneByteAuto);
      KaldiObjectHolder<CompressedMatrix>::Write(os, true, CompressedMatrix(m, cm));
    }
    return CopyOut(os.str(), buf, cap);
  } catch (const std::exception &) { return -2; }
}
// Reads any matrix object the way feature readers do (Matrix::Read accepts CM too); out needs rows*cols floats.
int ref_io_read_matrix(const char *bytes, int64_t n, int32_t *rows, int32_t *cols, float *out, int64_t cap) {
  try {
    std::istringstream is(std::string(bytes, n), std::ios::binary);
    KaldiObjectHolder<Matrix<BaseFloat> > h;
    if (!h.Read(is)) return -1;
    const Matrix<BaseFloat> &m = h.Value();
    *rows = m.NumRows(), *cols = m.NumCols();
    if ((int64_t)m.NumRows() * m.NumCols() <= cap && out) FromMatrix(m, out, m.NumCols());
    return 0;
  } catch (const std::exception &) { return -2; }
}
int64_t ref_io_write_int32_vector(const int32_t *v, int32_t count, char *buf, int64_t cap) {
  try {
    std::ostringstream os(std::ios::binary);
    BasicVectorHolder<int32>::Write(os, true, std::vector<int32>(v, v + count));
    return CopyOut(os.str(), buf, cap);
  } catch (const std::exception &) { return -2; }
}
int ref_io_read_int32_vector(const char *bytes, int64_t n, int32_t *out, int32_t cap) {
  try {
    std::istringstream is(std::string(bytes, n), std::ios::binary);
    BasicVectorHolder<int32> h;
    if (!h.Read(is)) return -1;
    if ((int32_t)h.Value().size() <= cap) std::copy(h.Value().begin(), h.Value().end(), out);
    return (int)h.Value().size();
  } catch (const std::exception &) { return -2; }
}
// A model file (\0B + TransitionModel + AmDiagGmm, as gmm-init-mono / gmm-est write it) for a monophone system with
// n_phones phones of 3 emitting states, whose 3*n_phones pdfs are the handle's pdfs in order.
int64_t ref_io_write_mdl(void *h, int32_t n_phones, char *buf, int64_t cap) {
  try {
    RefModel *rm = static_cast<RefModel *>(h);
    if (rm->am.NumPdfs() != 3 * n_phones) return -1;
    std::ostringstream topo_txt;
    topo_txt << "<Topology>\n<TopologyEntry>\n<ForPhones>\n";
    for (int32 p = 1; p <= n_phones; p++) topo_txt << p << " ";
    topo_txt << "\n</ForPhones>\n"
             << "<State> 0 <PdfClass> 0 <Transition> 0 0.75 <Transition> 1 0.25 </State>\n"
             << "<State> 1 <PdfClass> 1 <Transition> 1 0.6 <Transition> 2 0.2 <Transition> 3 0.2 </State>\n"
             << "<State> 2 <PdfClass> 2 <Transition> 2 0.75 <Transition> 3 0.25 </State>\n"
             << "<State> 3 </State>\n</TopologyEntry>\n</Topology>\n";
    std::istringstream ti(topo_txt.str());
    HmmTopology topo;
    topo.Read(ti, false);
    std::vector<int32> phones, num_pdf_classes(n_phones + 1, 3);
    for (int32 p = 1; p <= n_phones; p++) phones.push_back(p);
    ContextDependency *ctx = MonophoneContextDependency(phones, num_pdf_classes);
    TransitionModel tm(*ctx, topo);
    delete ctx;
    std::ostringstream os(std::ios::binary);
    InitKaldiOutputStream(os, true);  // the \0B marker Output::Open(binary) puts at the head of the file
    tm.Write(os, true);
    rm->am.Write(os, true);
    return CopyOut(os.str(), buf, cap);
  } catch (const std::exception &) { return -2; }
}
// TransitionModel::Read + AmDiagGmm::Read on the bytes of a model file: what the reference itself sees in it.
int ref_io_read_mdl(const char *bytes, int64_t n, int32_t *num_tids, int32_t *tid2pdf, int32_t tid_cap, float *log_probs,
                    int32_t *P, int32_t *N, float *gconsts, float *miv, float *iv, float *weights, int64_t gauss_cap) {
  try {
    std::istringstream is(std::string(bytes, n), std::ios::binary);
    bool binary;
    if (!InitKaldiInputStream(is, &binary)) return -1;
    TransitionModel tm;
    tm.Read(is, binary);
    AmDiagGmm am;
    am.Read(is, binary);
    *num_tids = tm.NumTransitionIds();
    if (tm.NumTransitionIds() + 1 <= tid_cap)
      for (int32 t = 1; t <= tm.NumTransitionIds(); t++) {
        tid2pdf[t] = tm.TransitionIdToPdf(t);
        log_probs[t] = tm.GetTransitionLogProb(t);
      }
    *P = am.NumPdfs();
    *N = am.NumGauss();
    if (am.NumGauss() <= gauss_cap) {
      int32 g = 0, D = am.Dim();
      for (int32 p = 0; p < am.NumPdfs(); p++) {
        const DiagGmm &d = am.GetPdf(p);
        for (int32 m = 0; m < d.NumGauss(); m++, g++) {
          gconsts[g] = d.gconsts()(m);
          weights[g] = d.weights()(m);
          for (int32 k = 0; k < D; k++) miv[(size_t)g * D + k] = d.means_invvars()(m, k), iv[(size_t)g * D + k] = d.inv_vars()(m, k);
        }
      }
    }
    return 0;
  } catch (const std::exception &) { return -2; }
}
// Reads a statistics file the way gmm-sum-accs / gmm-est do (gmm-sum-accs.cpp:44-50): Vector<double> transition accs
// (if n_trans > 0) then AccumAmDiagGmm::Read.
int ref_io_read_acc(void *h, const char *bytes, int64_t n, int32_t n_trans, double *trans, double *occ, double *mean_acc,
                    double *var_acc, double *tot_like, double *tot_frames) {
  try {
    RefModel *rm = static_cast<RefModel *>(h);
    std::istringstream is(std::string(bytes, n), std::ios::binary);
    bool binary;
    if (!InitKaldiInputStream(is, &binary)) return -1;
    if (n_trans > 0) {
      Vector<double> t;
      t.Read(is, binary);
      if (t.Dim() != n_trans) return -3;
      for (int32 i = 0; i < n_trans; i++) trans[i] = t(i);
    }
    AccumAmDiagGmm acc;
    acc.Read(is, binary, false);
    if (acc.NumAccs() != rm->am.NumPdfs()) return -4;
    int32 D = rm->am.Dim();
    for (int32 p = 0; p < acc.NumAccs(); p++) {
      const AccumDiagGmm &a = acc.GetAcc(p);
      if (a.Flags() != kGmmAll) return -5;
      int32 g0 = rm->offsets[p];
      for (int32 m = 0; m < a.NumGauss(); m++) {
        occ[g0 + m] = a.occupancy()(m);
        for (int32 d = 0; d < D; d++) {
          mean_acc[(size_t)(g0 + m) * D + d] = a.mean_accumulator()(m, d);
          var_acc[(size_t)(g0 + m) * D + d] = a.variance_accumulator()(m, d);
        }
      }
    }
    *tot_like = acc.TotLogLike();
    *tot_frames = acc.TotCount();
    return 0;
  } catch (const std::exception &) { return -2; }
}

// ---------------------------------------------------------------------------------------------------------
// CPU baseline for bench.py: the reference's whole path PCM -> loglikes, in memory, parallelised the
// reference's way (nj host threads over per-speaker splits, BLAS single-threaded inside each worker):
//   Mfcc::ComputeFeatures -> per-speaker AccCmvnStats/ApplyCmvn -> ComputeDeltas | SpliceFrames+LDA ->
//   [per-speaker fMLLR affine] -> dense all-pdf DiagGmm::LogLikelihoods (matrix form) + LogSumExp.
// Utterance u spans pcm[sample_offsets[u], sample_offsets[u+1]); utt2spk maps to [0, n_spk).
// mode 0: deltas(order, window); mode 1: splice(left,right) + lda[lda_rows x lda_cols].
// fmllr: n_spk matrices D x (D+1) or NULL.  loglikes (may be NULL to discard) rows packed by frame_offsets.
// Returns total frames, or <0.
// ---------------------------------------------------------------------------------------------------------
int64_t ref_pcm_to_loglikes(const orc_mfcc_opts *o, void *model_h, const int16_t *pcm, const int64_t *sample_offsets,
                            int32_t n_utts, const int32_t *utt2spk, int32_t n_spk, int32_t norm_vars, int32_t mode,
                            int32_t a, int32_t b, const float *lda, int32_t lda_rows, int32_t lda_cols,
                            const float *fmllr, int32_t nj, float *loglikes, const int64_t *frame_offsets,
                            int32_t ll_stride) {
  RefModel *rm = static_cast<RefModel *>(model_h);
  const int32 D = rm->am.Dim(), P = rm->am.NumPdfs();
  if (nj < 1) nj = 1;
  std::vector<int64_t> frames(nj, 0);
  std::vector<int> status(nj, 0);
  auto worker = [&](int j) {
    try {
      Mfcc mfcc(ToKaldi(o));
      Matrix<BaseFloat> lda_m;
      if (mode == 1) {
        lda_m.Resize(lda_rows, lda_cols);
        for (int r = 0; r < lda_rows; r++)
          std::memcpy(lda_m.RowData(r), lda + (size_t)r * lda_cols, sizeof(float) * lda_cols);
      }
      for (int32 s = j; s < n_spk; s += nj) {  // speaker-level split (utils/split_data.cpp:17-27)
        std::vector<int32> utts;
        for (int32 u = 0; u < n_utts; u++)
          if (utt2spk[u] == s) utts.push_back(u);
        std::vector<Matrix<BaseFloat> > raw(utts.size());
        Matrix<double> stats(2, o->num_ceps + 1);
        for (size_t i = 0; i < utts.size(); i++) {
          int32 u = utts[i];
          int64_t n = sample_offsets[u + 1] - sample_offsets[u];
          Vector<BaseFloat> wave((MatrixIndexT)n, kUndefined);
          for (int64_t k = 0; k < n; k++) wave(k) = pcm[sample_offsets[u] + k];  // wave-reader.cc:302-309
          mfcc.ComputeFeatures(wave, o->samp_freq, 1.0, &raw[i]);
          if (raw[i].NumRows() > 0) AccCmvnStats(raw[i], NULL, &stats);
        }
        for (size_t i = 0; i < utts.size(); i++) {
          int32 u = utts[i];
          if (raw[i].NumRows() == 0) continue;
          ApplyCmvn(stats, norm_vars != 0, &raw[i]);
          Matrix<BaseFloat> feats;
          if (mode == 0) {
            DeltaFeaturesOptions dopts(a, b);
            ComputeDeltas(dopts, raw[i], &feats);
          } else {
            Matrix<BaseFloat> spliced;
            SpliceFrames(raw[i], a, b, &spliced);
            feats.Resize(spliced.NumRows(), lda_rows);
            if (lda_cols == spliced.NumCols()) {
              feats.AddMatMat(1.0, spliced, kNoTrans, lda_m, kTrans, 0.0);
            } else {
              SubMatrix<BaseFloat> lin(lda_m, 0, lda_rows, 0, spliced.NumCols());
              feats.AddMatMat(1.0, spliced, kNoTrans, lin, kTrans, 0.0);
              Vector<BaseFloat> off(lda_rows);
              off.CopyColFromMat(lda_m, spliced.NumCols());
              feats.AddVecToRows(1.0, off);
            }
          }
          if (fmllr) {
            Matrix<BaseFloat> A(D, D + 1), outm(feats.NumRows(), D);
            for (int r = 0; r < D; r++)
              std::memcpy(A.RowData(r), fmllr + ((size_t)s * D + r) * (D + 1), sizeof(float) * (D + 1));
            SubMatrix<BaseFloat> lin(A, 0, D, 0, D);
            outm.AddMatMat(1.0, feats, kNoTrans, lin, kTrans, 0.0);
            Vector<BaseFloat> off(D);
            off.CopyColFromMat(A, D);
            outm.AddVecToRows(1.0, off);
            feats.Swap(&outm);
          }
          Matrix<BaseFloat> ll;
          int32 T = feats.NumRows();
          for (int32 p = 0; p < P; p++) {
            rm->am.GetPdf(p).LogLikelihoods(feats, &ll);
            for (int32 t = 0; t < T; t++) {
              BaseFloat v = ll.Row(t).LogSumExp();
              if (loglikes) loglikes[(size_t)(frame_offsets[u] + t) * ll_stride + p] = v;
            }
          }
          frames[j] += T;
        }
      }
    } catch (const std::exception &) { status[j] = -1; }
  };
  std::vector<std::thread> th;
  for (int j = 0; j < nj; j++) th.emplace_back(worker, j);
  for (auto &t : th) t.join();
  int64_t tot = 0;
  for (int j = 0; j < nj; j++) {
    if (status[j] < 0) return -1;
    tot += frames[j];
  }
  return tot;
}

// ---- Kaldi pitch: the reference's own ComputeKaldiPitch / ProcessPitch (feat/pitch-functions.cc:1291, 1581) ----
static PitchExtractionOptions ToKaldiPitch(const orc_pitch_opts *o) {
  PitchExtractionOptions p;
  p.samp_freq = o->samp_freq; p.frame_shift_ms = o->frame_shift_ms; p.frame_length_ms = o->frame_length_ms;
  p.preemph_coeff = o->preemph_coeff; p.min_f0 = o->min_f0; p.max_f0 = o->max_f0; p.soft_min_f0 = o->soft_min_f0;
  p.penalty_factor = o->penalty_factor; p.lowpass_cutoff = o->lowpass_cutoff; p.resample_freq = o->resample_freq;
  p.delta_pitch = o->delta_pitch; p.nccf_ballast = o->nccf_ballast; p.lowpass_filter_width = o->lowpass_filter_width;
  p.upsample_filter_width = o->upsample_filter_width; p.recompute_frame = o->recompute_frame;
  p.snip_edges = o->snip_edges != 0;
  return p;
}

// DownsampleWaveForm itself (feat/resample.cc:368-376); out nullable, returns the number of output samples
int64_t ref_downsample_waveform(float orig_freq, float new_freq, const float *wave, int64_t n, float *out) {
  try {
    SubVector<BaseFloat> w(const_cast<float *>(wave), (MatrixIndexT)n);
    Vector<BaseFloat> down;
    DownsampleWaveForm(orig_freq, w, new_freq, &down);
    if (out)
      for (MatrixIndexT i = 0; i < down.Dim(); i++) out[i] = down(i);
    return down.Dim();
  } catch (const std::exception &) { return -1; }
}

int ref_pitch_compute(const orc_pitch_opts *o, const float *wave, int64_t n, float *out, int32_t out_stride) {
  try {
    PitchExtractionOptions p = ToKaldiPitch(o);
    SubVector<BaseFloat> w(const_cast<float *>(wave), (MatrixIndexT)n);
    Matrix<BaseFloat> feats;
    ComputeKaldiPitch(p, w, &feats);
    FromMatrix(feats, out, out_stride);
    return feats.NumRows();
  } catch (const std::exception &) { return -1; }
}

int ref_process_pitch(const orc_process_pitch_opts *o, const float *in, int32_t T, int32_t in_stride, float *out,
                      int32_t out_stride) {
  try {
    ProcessPitchOptions p;
    p.pitch_scale = o->pitch_scale; p.pov_scale = o->pov_scale; p.pov_offset = o->pov_offset;
    p.delta_pitch_scale = o->delta_pitch_scale; p.delta_pitch_noise_stddev = o->delta_pitch_noise_stddev;
    p.normalization_left_context = o->normalization_left_context;
    p.normalization_right_context = o->normalization_right_context; p.delta_window = o->delta_window;
    p.delay = o->delay; p.add_pov_feature = o->add_pov_feature != 0;
    p.add_normalized_log_pitch = o->add_normalized_log_pitch != 0; p.add_delta_pitch = o->add_delta_pitch != 0;
    p.add_raw_log_pitch = o->add_raw_log_pitch != 0;
    Matrix<BaseFloat> f, outm;
    ToMatrix(in, T, 2, in_stride, &f);
    ProcessPitch(p, f, &outm);
    FromMatrix(outm, out, out_stride);
    return outm.NumRows();
  } catch (const std::exception &) { return -1; }
}

}  // extern "C"
