// vbgpu_kaldi.h — header-only C++ adaptors: the reference's OWN host types for the acoustic-scoring path, served by
// libvbgpu.so through the C ABI of vbgpu.h.  This is the host side of the drop-in: it compiles against the Kaldi headers
// that VoiceBridge links (paths relative to kaldi-master/src) and is what a VoiceBridge build would include instead of
// the CPU classes (INTEGRATION.md shows the call sites).
//
//   vbgpu::GpuMfcc                 OfflineFeatureTpl<MfccComputer>        feat/feature-common.h:110-178
//   vbgpu::GpuPitch                ComputeKaldiPitch / ProcessPitch       feat/pitch-functions.h:366-395
//   vbgpu::GpuFeaturePipeline      ApplyCmvn + ComputeDeltas | SpliceFrames + transform   transform/cmvn.cc:64-113,
//                                  feat/feature-functions.cc:160-171,205-226, featbin/transform-feats.cc:95-107
//   vbgpu::GpuAmDiagGmm            AmDiagGmm (device-resident copy)       gmm/am-diag-gmm.h:36-105
//   vbgpu::DecodableAmDiagGmmGpu   DecodableAmDiagGmmScaled               gmm/decodable-am-diag-gmm.h:121-160
//   vbgpu::AccumAmDiagGmmGpu       AccumAmDiagGmm                         gmm/mle-am-diag-gmm.h:34-108
//
// Error behaviour follows the reference: a failing call raises KALDI_ERR (std::runtime_error), which the L3 functions of
// VoiceBridge catch and turn into "return -1" (VB/src/featbin/compute-mfcc-feats.cpp:192-197).
// No arithmetic happens in this header: every method is a call into the library.
#ifndef VBGPU_KALDI_H_
#define VBGPU_KALDI_H_

#include <vector>

#include "base/kaldi-common.h"
#include "feat/feature-functions.h"
#include "feat/feature-mfcc.h"
#include "feat/pitch-functions.h"
#include "gmm/am-diag-gmm.h"
#include "gmm/mle-am-diag-gmm.h"
#include "hmm/transition-model.h"
#include "itf/decodable-itf.h"
#include "matrix/kaldi-matrix.h"
#include "transform/fmllr-diag-gmm.h"
#include "transform/mllt.h"

#include "vbgpu.h"

namespace vbgpu {

using kaldi::BaseFloat;
using kaldi::int32;

inline void Check(int64_t rc, const char *what) {
  if (rc < 0) KALDI_ERR << what << ": " << vbgpu_last_error();
}

inline vbgpu_mfcc_opts ToVbgpu(const kaldi::MfccOptions &m) {
  vbgpu_mfcc_opts o;
  vbgpu_mfcc_opts_default(&o);
  const kaldi::FrameExtractionOptions &f = m.frame_opts;
  o.samp_freq = f.samp_freq;
  o.frame_shift_ms = f.frame_shift_ms;
  o.frame_length_ms = f.frame_length_ms;
  o.dither = f.dither;
  o.preemph_coeff = f.preemph_coeff;
  o.remove_dc_offset = f.remove_dc_offset;
  o.window_type = f.window_type == "povey" ? 0 : f.window_type == "hamming" ? 1 : f.window_type == "hanning" ? 2
                : f.window_type == "rectangular" ? 3 : f.window_type == "blackman" ? 4 : -1;
  if (o.window_type < 0) KALDI_ERR << "Invalid window type " << f.window_type;  // feature-window.cc:128
  o.round_to_power_of_two = f.round_to_power_of_two;
  o.blackman_coeff = f.blackman_coeff;
  o.snip_edges = f.snip_edges;
  o.num_bins = m.mel_opts.num_bins;
  o.low_freq = m.mel_opts.low_freq;
  o.high_freq = m.mel_opts.high_freq;
  o.vtln_low = m.mel_opts.vtln_low;
  o.vtln_high = m.mel_opts.vtln_high;
  o.htk_mode = m.mel_opts.htk_mode;
  o.num_ceps = m.num_ceps;
  o.use_energy = m.use_energy;
  o.energy_floor = m.energy_floor;
  o.raw_energy = m.raw_energy;
  o.cepstral_lifter = m.cepstral_lifter;
  o.htk_compat = m.htk_compat;
  return o;
}

// ---- OfflineFeatureTpl<MfccComputer> ------------------------------------------------------------------------------
class GpuMfcc {
 public:
  explicit GpuMfcc(const kaldi::MfccOptions &opts, int device = 0) : opts_(opts), h_(NULL), device_(device) {
    vbgpu_mfcc_opts o = ToVbgpu(opts);
    Check(vbgpu_mfcc_create(&o, device, &h_), "vbgpu_mfcc_create");
  }
  ~GpuMfcc() {
    vbgpu_mfcc_destroy(h_);
    if (down_ != NULL) vbgpu_downsample_destroy(down_);
  }
  int32 Dim() const { return vbgpu_mfcc_dim(h_); }
  // Same contract as OfflineFeatureTpl::ComputeFeatures (feature-common-inl.h:29-59): resizes *output to NumFrames x Dim.
  // A wave sampled faster than the options say is down-sampled on the device (DownsampleWaveForm, resample.cc:368-376) when
  // frame_opts.allow_downsample is set and refused otherwise; a slower one is always refused.
  void ComputeFeatures(const kaldi::VectorBase<BaseFloat> &wave, BaseFloat sample_freq, BaseFloat vtln_warp,
                       kaldi::Matrix<BaseFloat> *output) {
    KALDI_ASSERT(output != NULL);
    const BaseFloat new_sample_freq = opts_.frame_opts.samp_freq;
    if (sample_freq == new_sample_freq) {
      Compute(wave.Data(), wave.Dim(), vtln_warp, output);
    } else if (new_sample_freq < sample_freq) {
      if (!opts_.frame_opts.allow_downsample)
        KALDI_ERR << "Waveform and config sample Frequency mismatch: " << sample_freq << " .vs " << new_sample_freq
                  << " ( use --allow_downsample=true option to allow  downsampling the waveform).";
      if (down_ == NULL || down_freq_ != sample_freq) {
        if (down_ != NULL) vbgpu_downsample_destroy(down_);
        down_ = NULL;
        Check(vbgpu_downsample_create(sample_freq, new_sample_freq, device_, &down_), "vbgpu_downsample_create");
        down_freq_ = sample_freq;
      }
      kaldi::Vector<BaseFloat> down(static_cast<int32>(vbgpu_downsample_num_out(down_, wave.Dim())), kaldi::kUndefined);
      if (down.Dim() > 0) Check(vbgpu_downsample_f32(down_, wave.Data(), wave.Dim(), down.Data()), "vbgpu_downsample_f32");
      Compute(down.Data(), down.Dim(), vtln_warp, output);
    } else {
      KALDI_ERR << "The waveform is allowed to get downsampled. New sample Frequency " << new_sample_freq
                << " is larger than waveform original sampling frequency " << sample_freq;
    }
  }
  vbgpu_mfcc_t handle() const { return h_; }

 private:
  void Compute(const BaseFloat *wave, int64_t n, BaseFloat vtln_warp, kaldi::Matrix<BaseFloat> *output) {
    const int64_t offs[2] = {0, n};
    const int64_t T = vbgpu_mfcc_num_frames(h_, n);
    Check(T, "vbgpu_mfcc_num_frames");
    output->Resize(static_cast<int32>(T), Dim(), kaldi::kUndefined);
    if (T == 0) return;
    Check(vbgpu_mfcc_compute_f32(h_, wave, offs, 1, vtln_warp == 1.0f ? NULL : &vtln_warp, output->Data(), output->Stride()),
          "vbgpu_mfcc_compute_f32");
  }
  kaldi::MfccOptions opts_;
  vbgpu_mfcc_t h_;
  vbgpu_resample_t down_ = NULL;  // tables of the last (wave rate -> samp_freq) pair seen
  BaseFloat down_freq_ = 0;
  int device_ = 0;
  KALDI_DISALLOW_COPY_AND_ASSIGN(GpuMfcc);
};

// ---- compute-kaldi-pitch-feats / process-kaldi-pitch-feats / compute-and-process-kaldi-pitch-feats ----------------------
inline vbgpu_pitch_opts ToVbgpu(const kaldi::PitchExtractionOptions &p) {
  // the offline path only: what compute-kaldi-pitch-feats does with its default flags (pitch-functions.cc:1291-1325)
  if (p.frames_per_chunk != 0 || p.simulate_first_pass_online || p.nccf_ballast_online || p.max_frames_latency != 0)
    KALDI_ERR << "vbgpu pitch: frames-per-chunk, simulate-first-pass-online, nccf-ballast-online and max-frames-latency "
                 "(the online emulation modes) are not served by the GPU path";
  vbgpu_pitch_opts o;
  vbgpu_pitch_opts_default(&o);
  o.samp_freq = p.samp_freq;
  o.frame_shift_ms = p.frame_shift_ms;
  o.frame_length_ms = p.frame_length_ms;
  o.preemph_coeff = p.preemph_coeff;
  o.min_f0 = p.min_f0;
  o.max_f0 = p.max_f0;
  o.soft_min_f0 = p.soft_min_f0;
  o.penalty_factor = p.penalty_factor;
  o.lowpass_cutoff = p.lowpass_cutoff;
  o.resample_freq = p.resample_freq;
  o.delta_pitch = p.delta_pitch;
  o.nccf_ballast = p.nccf_ballast;
  o.lowpass_filter_width = p.lowpass_filter_width;
  o.upsample_filter_width = p.upsample_filter_width;
  o.recompute_frame = p.recompute_frame;
  o.snip_edges = p.snip_edges ? 1 : 0;
  return o;
}

inline vbgpu_process_pitch_opts ToVbgpu(const kaldi::ProcessPitchOptions &p) {
  vbgpu_process_pitch_opts o;
  vbgpu_process_pitch_opts_default(&o);
  o.pitch_scale = p.pitch_scale;
  o.pov_scale = p.pov_scale;
  o.pov_offset = p.pov_offset;
  o.delta_pitch_scale = p.delta_pitch_scale;
  o.delta_pitch_noise_stddev = p.delta_pitch_noise_stddev;
  o.normalization_left_context = p.normalization_left_context;
  o.normalization_right_context = p.normalization_right_context;
  o.delta_window = p.delta_window;
  o.delay = p.delay;
  o.add_pov_feature = p.add_pov_feature ? 1 : 0;
  o.add_normalized_log_pitch = p.add_normalized_log_pitch ? 1 : 0;
  o.add_delta_pitch = p.add_delta_pitch ? 1 : 0;
  o.add_raw_log_pitch = p.add_raw_log_pitch ? 1 : 0;
  return o;
}

class GpuPitch {
 public:
  explicit GpuPitch(const kaldi::PitchExtractionOptions &opts, int device = 0) : h_(NULL) {
    vbgpu_pitch_opts o = ToVbgpu(opts);
    Check(vbgpu_pitch_create(&o, device, &h_), "vbgpu_pitch_create");
  }
  ~GpuPitch() { vbgpu_pitch_destroy(h_); }
  // kaldi::ComputeKaldiPitch(opts, wave, &output): NumFrames x 2 = (NCCF, pitch); 0 x 0 when the wave is too short.
  void ComputeKaldiPitch(const kaldi::VectorBase<BaseFloat> &wave, kaldi::Matrix<BaseFloat> *output) {
    Compute(wave, NULL, 2, 0, output);
  }
  // kaldi::ComputeAndProcessKaldiPitch(pitch_opts, process_opts, wave, &output)
  void ComputeAndProcessKaldiPitch(const kaldi::ProcessPitchOptions &process_opts,
                                   const kaldi::VectorBase<BaseFloat> &wave, kaldi::Matrix<BaseFloat> *output) {
    vbgpu_process_pitch_opts po = ToVbgpu(process_opts);
    Compute(wave, &po, Dim(po), po.delay, output);
  }
  // kaldi::ProcessPitch(opts, input, &output)
  void ProcessPitch(const kaldi::ProcessPitchOptions &process_opts, const kaldi::MatrixBase<BaseFloat> &input,
                    kaldi::Matrix<BaseFloat> *output) {
    KALDI_ASSERT(output != NULL && (input.NumRows() == 0 || input.NumCols() == 2));
    vbgpu_process_pitch_opts po = ToVbgpu(process_opts);
    const int64_t offs[2] = {0, input.NumRows()};
    output->Resize(input.NumRows() > 0 ? input.NumRows() + po.delay : 0, Dim(po), kaldi::kUndefined);
    if (input.NumRows() == 0) return;
    Check(vbgpu_pitch_process(h_, &po, input.Data(), input.Stride(), offs, 1, output->Data(), output->Stride()),
          "vbgpu_pitch_process");
  }

 private:
  static int32 Dim(const vbgpu_process_pitch_opts &po) {
    return (po.add_pov_feature ? 1 : 0) + (po.add_normalized_log_pitch ? 1 : 0) + (po.add_delta_pitch ? 1 : 0) +
           (po.add_raw_log_pitch ? 1 : 0);
  }
  void Compute(const kaldi::VectorBase<BaseFloat> &wave, const vbgpu_process_pitch_opts *po, int32 dim, int32 delay,
               kaldi::Matrix<BaseFloat> *output) {
    KALDI_ASSERT(output != NULL);
    const int64_t offs[2] = {0, wave.Dim()};
    const int64_t T = vbgpu_pitch_num_frames(h_, wave.Dim());
    Check(T, "vbgpu_pitch_num_frames");
    if (T == 0) {
      KALDI_WARN << "No frames output in pitch extraction";
      output->Resize(0, 0);
      return;
    }
    output->Resize(static_cast<int32>(T) + delay, dim, kaldi::kUndefined);
    Check(vbgpu_pitch_compute_f32(h_, wave.Data(), offs, 1, po, output->Data(), output->Stride()),
          "vbgpu_pitch_compute_f32");
  }
  vbgpu_pitch_t h_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(GpuPitch);
};

// ---- apply-cmvn | add-deltas (or splice-feats | transform-feats) [| transform-feats fMLLR] ----------------------------
class GpuFeaturePipeline {
 public:
  // delta mode: ApplyCmvn(norm_vars) then ComputeDeltas(delta_opts).
  GpuFeaturePipeline(int32 in_dim, bool norm_vars, const kaldi::DeltaFeaturesOptions &delta_opts, int device = 0)
      : h_(NULL) {
    vbgpu_feat_opts o;
    vbgpu_feat_opts_default(&o);
    o.norm_vars = norm_vars;
    o.mode = 0;
    o.delta_order = delta_opts.order;
    o.delta_window = delta_opts.window;
    Check(vbgpu_feat_create(&o, in_dim, NULL, 0, 0, device, &h_), "vbgpu_feat_create");
  }
  // lda mode: ApplyCmvn, SpliceFrames(left, right), then the global transform (final.mat).
  GpuFeaturePipeline(int32 in_dim, bool norm_vars, int32 left, int32 right, const kaldi::Matrix<BaseFloat> &transform,
                     int device = 0)
      : h_(NULL) {
    vbgpu_feat_opts o;
    vbgpu_feat_opts_default(&o);
    o.norm_vars = norm_vars;
    o.mode = 1;
    o.splice_left = left;
    o.splice_right = right;
    kaldi::Matrix<BaseFloat> packed(transform.NumRows(), transform.NumCols(), kaldi::kUndefined, kaldi::kStrideEqualNumCols);
    packed.CopyFromMat(transform);
    Check(vbgpu_feat_create(&o, in_dim, packed.Data(), packed.NumRows(), packed.NumCols(), device, &h_), "vbgpu_feat_create");
  }
  ~GpuFeaturePipeline() { vbgpu_feat_destroy(h_); }
  int32 OutDim() const { return vbgpu_feat_out_dim(h_); }
  // One utterance.  cmvn_stats: the 2 x (dim+1) matrix AccCmvnStats produces (per speaker or per utterance);
  // fmllr: NULL, or the speaker's dim x (dim+1) / dim x dim matrix (transform-feats.cpp:95-107).
  void Run(const kaldi::MatrixBase<BaseFloat> &feats, const kaldi::MatrixBase<double> &cmvn_stats,
           const kaldi::MatrixBase<BaseFloat> *fmllr, kaldi::Matrix<BaseFloat> *out) {
    const int64_t fo[2] = {0, feats.NumRows()};
    kaldi::Matrix<double> st(cmvn_stats.NumRows(), cmvn_stats.NumCols(), kaldi::kUndefined, kaldi::kStrideEqualNumCols);
    st.CopyFromMat(cmvn_stats);
    kaldi::Matrix<BaseFloat> fm;
    if (fmllr) {
      fm.Resize(fmllr->NumRows(), fmllr->NumCols(), kaldi::kUndefined, kaldi::kStrideEqualNumCols);
      fm.CopyFromMat(*fmllr);
    }
    out->Resize(feats.NumRows(), OutDim(), kaldi::kUndefined);
    if (feats.NumRows() == 0) return;
    Check(vbgpu_feat_run(h_, feats.Data(), feats.Stride(), fo, 1, NULL, 1, st.Data(), fmllr ? fm.Data() : NULL,
                         fmllr ? fm.NumCols() : 0, out->Data(), out->Stride()),
          "vbgpu_feat_run");
  }
  vbgpu_feat_t handle() const { return h_; }

 private:
  vbgpu_feat_t h_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(GpuFeaturePipeline);
};

// ---- AmDiagGmm on the device -------------------------------------------------------------------------------------------
class GpuAmDiagGmm {
 public:
  explicit GpuAmDiagGmm(const kaldi::AmDiagGmm &am, int device = 0) : h_(NULL) {
    const int32 P = am.NumPdfs(), D = am.Dim();
    offsets_.assign(P + 1, 0);
    for (int32 p = 0; p < P; p++) offsets_[p + 1] = offsets_[p] + am.GetPdf(p).NumGauss();
    const int32 N = offsets_[P];
    std::vector<float> gc(N), miv(static_cast<size_t>(N) * D), iv(static_cast<size_t>(N) * D);
    for (int32 p = 0; p < P; p++) {  // DiagGmm::gconsts()/means_invvars()/inv_vars(), gmm/diag-gmm.h:174-180
      const kaldi::DiagGmm &g = am.GetPdf(p);
      for (int32 m = 0; m < g.NumGauss(); m++) {
        const size_t r = offsets_[p] + m;
        gc[r] = g.gconsts()(m);
        for (int32 d = 0; d < D; d++) {
          miv[r * D + d] = g.means_invvars()(m, d);
          iv[r * D + d] = g.inv_vars()(m, d);
        }
      }
    }
    Check(vbgpu_gmm_create(P, D, offsets_.data(), gc.data(), miv.data(), iv.data(), D, device, &h_), "vbgpu_gmm_create");
  }
  ~GpuAmDiagGmm() { vbgpu_gmm_destroy(h_); }
  int32 NumPdfs() const { return vbgpu_gmm_num_pdfs(h_); }
  int32 NumGauss() const { return vbgpu_gmm_num_gauss(h_); }
  int32 Dim() const { return vbgpu_gmm_dim(h_); }
  const std::vector<int32_t> &PdfOffsets() const { return offsets_; }
  // Dense [T x NumPdfs] log-likelihoods: what DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased fills lazily.
  void LogLikelihoods(const kaldi::MatrixBase<BaseFloat> &feats, kaldi::Matrix<BaseFloat> *loglikes) const {
    loglikes->Resize(feats.NumRows(), NumPdfs(), kaldi::kUndefined);
    if (feats.NumRows() == 0) return;
    Check(vbgpu_gmm_score(h_, feats.Data(), feats.NumRows(), feats.Stride(), loglikes->Data(), loglikes->Stride()),
          "vbgpu_gmm_score");  // NaN/Inf -> KALDI_ERR, as decodable-am-diag-gmm.cc:65-66
  }
  vbgpu_gmm_t handle() const { return h_; }

 private:
  vbgpu_gmm_t h_;
  std::vector<int32_t> offsets_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(GpuAmDiagGmm);
};

// ---- DecodableAmDiagGmmScaled --------------------------------------------------------------------------------------------
class DecodableAmDiagGmmGpu : public kaldi::DecodableInterface {
 public:
  // One launch scores every (frame, pdf); LogLikelihood() is then an array read.
  DecodableAmDiagGmmGpu(const GpuAmDiagGmm &am, const kaldi::TransitionModel &tm,
                        const kaldi::MatrixBase<BaseFloat> &feats, BaseFloat scale)
      : trans_model_(tm), scale_(scale) {
    am.LogLikelihoods(feats, &loglikes_);
  }
  virtual BaseFloat LogLikelihood(int32 frame, int32 tid) {  // tid is 1-based (decodable-itf.h:83-119)
    return scale_ * loglikes_(frame, trans_model_.TransitionIdToPdf(tid));
  }
  virtual int32 NumFramesReady() const { return loglikes_.NumRows(); }
  virtual bool IsLastFrame(int32 frame) const {
    KALDI_ASSERT(frame < NumFramesReady());
    return frame == NumFramesReady() - 1;
  }
  virtual int32 NumIndices() const { return trans_model_.NumTransitionIds(); }
  const kaldi::Matrix<BaseFloat> &loglikes() const { return loglikes_; }

 private:
  const kaldi::TransitionModel &trans_model_;
  BaseFloat scale_;
  kaldi::Matrix<BaseFloat> loglikes_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(DecodableAmDiagGmmGpu);
};

// ---- a whole job's utterances in one scoring launch ---------------------------------------------------
// gmm-align-compiled.cpp:92-130, gmm-rescore-lattice.cpp and gmm-latgen-faster.cpp:103-170 build one decodable per
// utterance.  This scores every utterance of a job (a speaker split) in ONE call on the packed rows and hands out
// per-utterance DecodableInterface views; per-row results do not depend on how rows are batched, so alignments and
// lattices are the ones the per-utterance decodable gives.
class BatchDecodableAmDiagGmmGpu {
 public:
  BatchDecodableAmDiagGmmGpu(const GpuAmDiagGmm &am, const kaldi::TransitionModel &tm,
                             const std::vector<const kaldi::MatrixBase<BaseFloat> *> &feats, BaseFloat scale)
      : trans_model_(tm), scale_(scale), offsets_(1, 0) {
    const int32 D = am.Dim();
    for (size_t u = 0; u < feats.size(); u++) {
      KALDI_ASSERT(feats[u]->NumCols() == D);
      offsets_.push_back(offsets_.back() + feats[u]->NumRows());
    }
    kaldi::Matrix<BaseFloat> packed(offsets_.back(), D, kaldi::kUndefined);
    for (size_t u = 0; u < feats.size(); u++)
      if (feats[u]->NumRows() > 0) packed.RowRange(offsets_[u], feats[u]->NumRows()).CopyFromMat(*feats[u]);
    am.LogLikelihoods(packed, &loglikes_);
    for (size_t u = 0; u < feats.size(); u++) views_.push_back(View(this, offsets_[u], offsets_[u + 1] - offsets_[u]));
  }
  int32 NumUtterances() const { return static_cast<int32>(views_.size()); }
  kaldi::DecodableInterface *Utterance(int32 u) { return &views_[u]; }  // owned by the batch
  const kaldi::Matrix<BaseFloat> &loglikes() const { return loglikes_; }

 private:
  class View : public kaldi::DecodableInterface {
   public:
    View(const BatchDecodableAmDiagGmmGpu *b, int32 first, int32 n) : b_(b), first_(first), n_(n) {}
    virtual BaseFloat LogLikelihood(int32 frame, int32 tid) {
      return b_->scale_ * b_->loglikes_(first_ + frame, b_->trans_model_.TransitionIdToPdf(tid));
    }
    virtual int32 NumFramesReady() const { return n_; }
    virtual bool IsLastFrame(int32 frame) const { return frame == n_ - 1; }
    virtual int32 NumIndices() const { return b_->trans_model_.NumTransitionIds(); }

   private:
    const BatchDecodableAmDiagGmmGpu *b_;
    int32 first_, n_;
  };
  const kaldi::TransitionModel &trans_model_;
  BaseFloat scale_;
  std::vector<int32> offsets_;
  kaldi::Matrix<BaseFloat> loglikes_;
  std::vector<View> views_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(BatchDecodableAmDiagGmmGpu);
};

// ---- sparse consumers: only what the consumer reads crosses PCIe -----------------------------------------
// Forced alignment (gmm-align-compiled.cpp:119-128 -> AlignUtteranceWrapper) only asks its decodable for the pdfs of the
// utterance's own training graph.  PdfsOfGraph() lists them; BatchSubsetDecodableAmDiagGmmGpu scores every utterance of a
// job against ITS list in one call (vbgpu_gmm_score_subset) and hands out per-utterance DecodableInterface views.
inline void PdfsOfGraph(const fst::Fst<fst::StdArc> &graph, const kaldi::TransitionModel &tm, std::vector<int32> *pdfs) {
  std::vector<char> seen(tm.NumPdfs(), 0);
  for (fst::StateIterator<fst::Fst<fst::StdArc> > s(graph); !s.Done(); s.Next())
    for (fst::ArcIterator<fst::Fst<fst::StdArc> > a(graph, s.Value()); !a.Done(); a.Next())
      if (a.Value().ilabel > 0) seen[tm.TransitionIdToPdf(a.Value().ilabel)] = 1;
  pdfs->clear();
  for (int32 p = 0; p < tm.NumPdfs(); p++)
    if (seen[p]) pdfs->push_back(p);
}

class BatchSubsetDecodableAmDiagGmmGpu {
 public:
  BatchSubsetDecodableAmDiagGmmGpu(const GpuAmDiagGmm &am, const kaldi::TransitionModel &tm,
                                   const std::vector<const kaldi::MatrixBase<BaseFloat> *> &feats,
                                   const std::vector<std::vector<int32> > &pdf_subsets, BaseFloat scale)
      : trans_model_(tm), scale_(scale) {
    KALDI_ASSERT(feats.size() == pdf_subsets.size() && !feats.empty());
    const int32 D = am.Dim(), n = static_cast<int32>(feats.size());
    std::vector<int64_t> fo(1, 0), so(1, 0);
    std::vector<int32_t> pdfs;
    for (int32 u = 0; u < n; u++) {
      KALDI_ASSERT(feats[u]->NumCols() == D);
      fo.push_back(fo.back() + feats[u]->NumRows());
      so.push_back(so.back() + static_cast<int64_t>(pdf_subsets[u].size()));
      pdfs.insert(pdfs.end(), pdf_subsets[u].begin(), pdf_subsets[u].end());
    }
    kaldi::Matrix<BaseFloat> packed(fo.back(), D, kaldi::kUndefined);
    for (int32 u = 0; u < n; u++)
      if (feats[u]->NumRows() > 0) packed.RowRange(fo[u], feats[u]->NumRows()).CopyFromMat(*feats[u]);
    out_offsets_.resize(n + 1);
    int64_t total = 0;
    for (int32 u = 0; u < n; u++) total += (fo[u + 1] - fo[u]) * (so[u + 1] - so[u]);
    scores_.resize(std::max<int64_t>(total, 1));
    if (pdfs.empty()) pdfs.push_back(0);
    Check(vbgpu_gmm_score_subset(am.handle(), packed.Data(), packed.NumRows(), packed.Stride(), fo.data(), n, so.data(),
                                 pdfs.data(), scores_.data(), out_offsets_.data()),
          "vbgpu_gmm_score_subset");
    for (int32 u = 0; u < n; u++) {
      std::vector<int32> local(tm.NumPdfs(), -1);
      for (size_t k = 0; k < pdf_subsets[u].size(); k++) local[pdf_subsets[u][k]] = static_cast<int32>(k);
      views_.push_back(View(this, out_offsets_[u], static_cast<int32>(fo[u + 1] - fo[u]),
                            static_cast<int32>(pdf_subsets[u].size()), local));
    }
  }
  int32 NumUtterances() const { return static_cast<int32>(views_.size()); }
  kaldi::DecodableInterface *Utterance(int32 u) { return &views_[u]; }  // owned by the batch
  int64_t NumScores() const { return out_offsets_.back(); }              // floats that crossed PCIe

 private:
  class View : public kaldi::DecodableInterface {
   public:
    View(const BatchSubsetDecodableAmDiagGmmGpu *b, int64_t first, int32 n, int32 cols, const std::vector<int32> &local)
        : b_(b), first_(first), n_(n), cols_(cols), local_(local) {}
    virtual BaseFloat LogLikelihood(int32 frame, int32 tid) {
      const int32 k = local_[b_->trans_model_.TransitionIdToPdf(tid)];
      if (k < 0) KALDI_ERR << "transition-id " << tid << " is not in this utterance's pdf subset";
      return b_->scale_ * b_->scores_[first_ + static_cast<int64_t>(frame) * cols_ + k];
    }
    virtual int32 NumFramesReady() const { return n_; }
    virtual bool IsLastFrame(int32 frame) const { return frame == n_ - 1; }
    virtual int32 NumIndices() const { return b_->trans_model_.NumTransitionIds(); }

   private:
    const BatchSubsetDecodableAmDiagGmmGpu *b_;
    int64_t first_;
    int32 n_, cols_;
    std::vector<int32> local_;  // pdf-id -> column of the utterance's block, -1 = not asked for
  };
  const kaldi::TransitionModel &trans_model_;
  BaseFloat scale_;
  std::vector<int64_t> out_offsets_;
  std::vector<BaseFloat> scores_;
  std::vector<View> views_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(BatchSubsetDecodableAmDiagGmmGpu);
};

// Lattice rescoring (gmm-rescore-lattice.cpp -> RescoreLattice / RescoreCompactLattice, lat/lattice-functions.cc:1214-1401)
// asks for one (frame, transition-id) pair per arc, in an order that depends on the lattice only.  GatherDecodable first
// RECORDS the queries of a dry run of the reference's own RescoreLattice (on a copy of the lattice), GatherScorer scores all
// recorded arcs of all utterances in one call (vbgpu_gmm_score_gather: n floats back instead of frames x pdfs), and the
// same decodable then REPLAYS the answers to the real RescoreLattice run.
class GatherDecodable : public kaldi::DecodableInterface {
 public:
  GatherDecodable(const kaldi::TransitionModel &tm, int32 num_frames, BaseFloat scale)
      : trans_model_(tm), n_(num_frames), scale_(scale), replay_(false), next_(0) {}
  virtual BaseFloat LogLikelihood(int32 frame, int32 tid) {
    if (!replay_) {
      frames_.push_back(frame);
      pdfs_.push_back(trans_model_.TransitionIdToPdf(tid));
      return 0.0;
    }
    KALDI_ASSERT(next_ < frames_.size() && frames_[next_] == frame && pdfs_[next_] == trans_model_.TransitionIdToPdf(tid));
    return scale_ * scores_[next_++];
  }
  virtual int32 NumFramesReady() const { return n_; }
  virtual bool IsLastFrame(int32 frame) const { return frame == n_ - 1; }
  virtual int32 NumIndices() const { return trans_model_.NumTransitionIds(); }
  size_t NumArcs() const { return frames_.size(); }

 private:
  friend class GatherScorer;
  const kaldi::TransitionModel &trans_model_;
  int32 n_;
  BaseFloat scale_;
  bool replay_;
  size_t next_;
  std::vector<int32> frames_, pdfs_;
  std::vector<BaseFloat> scores_;
};

class GatherScorer {
 public:
  // decs[u] has recorded the arcs of utterance u (features feats[u]); afterwards every decs[u] replays.
  static void Score(const GpuAmDiagGmm &am, const std::vector<const kaldi::MatrixBase<BaseFloat> *> &feats,
                    const std::vector<GatherDecodable *> &decs) {
    KALDI_ASSERT(feats.size() == decs.size() && !feats.empty());
    const int32 D = am.Dim();
    std::vector<int64_t> fo(1, 0);
    for (size_t u = 0; u < feats.size(); u++) fo.push_back(fo.back() + feats[u]->NumRows());
    kaldi::Matrix<BaseFloat> packed(fo.back(), D, kaldi::kUndefined);
    std::vector<int32_t> frames, pdfs;
    for (size_t u = 0; u < feats.size(); u++) {
      if (feats[u]->NumRows() > 0) packed.RowRange(fo[u], feats[u]->NumRows()).CopyFromMat(*feats[u]);
      for (size_t i = 0; i < decs[u]->frames_.size(); i++) {
        frames.push_back(static_cast<int32_t>(fo[u]) + decs[u]->frames_[i]);
        pdfs.push_back(decs[u]->pdfs_[i]);
      }
    }
    std::vector<BaseFloat> out(std::max<size_t>(frames.size(), 1));
    Check(vbgpu_gmm_score_gather(am.handle(), packed.Data(), packed.NumRows(), packed.Stride(), frames.data(), pdfs.data(),
                                 static_cast<int64_t>(frames.size()), out.data()),
          "vbgpu_gmm_score_gather");
    size_t k = 0;
    for (size_t u = 0; u < decs.size(); u++) {
      decs[u]->scores_.assign(out.begin() + k, out.begin() + k + decs[u]->frames_.size());
      k += decs[u]->frames_.size();
      decs[u]->replay_ = true;
      decs[u]->next_ = 0;
    }
  }
};

// ---- FmllrDiagGmmAccs for all speakers of a job --------------------------------------------------------
// gmm-est-fmllr.cpp:40-55 calls FmllrDiagGmmAccs::AccumulateForGmm once per (frame, pdf).  Here a packed batch of
// utterances is accumulated per call on the device; CopyTo() fills the reference's own accumulator for one speaker, whose
// Update() (the solver) then runs unchanged.
class FmllrAccsGpu {
 public:
  FmllrAccsGpu(const GpuAmDiagGmm &am, int32 num_spk) : dim_(am.Dim()), h_(NULL) {
    Check(vbgpu_fmllr_create(am.handle(), num_spk, &h_), "vbgpu_fmllr_create");
  }
  ~FmllrAccsGpu() { vbgpu_fmllr_destroy(h_); }
  // feats: packed rows of all utterances; pdf_ids[t] = TransitionIdToPdf(alignment[t]); utterance u owns rows
  // [frame_offsets[u], frame_offsets[u+1]) and belongs to speaker utt2spk[u].  Returns the summed frame log-likelihoods.
  double Accumulate(const kaldi::MatrixBase<BaseFloat> &feats, const std::vector<int32> &pdf_ids,
                    const std::vector<int64_t> &frame_offsets, const std::vector<int32> &utt2spk,
                    const std::vector<BaseFloat> *weights = NULL) {
    KALDI_ASSERT(static_cast<int32>(pdf_ids.size()) == feats.NumRows() && frame_offsets.size() == utt2spk.size() + 1);
    double like = 0.0;
    if (feats.NumRows() == 0) return like;
    Check(vbgpu_fmllr_accumulate(h_, feats.Data(), feats.NumRows(), feats.Stride(), pdf_ids.data(),
                                 weights ? weights->data() : NULL, frame_offsets.data(),
                                 static_cast<int32>(utt2spk.size()), utt2spk.data(), &like),
          "vbgpu_fmllr_accumulate");
    return like;
  }
  // Overwrites beta_, K_ and G_ of `stats` (already Init'ed for this dimension) with speaker spk's device statistics.
  void CopyTo(int32 spk, kaldi::AffineXformStats *stats) const {
    const int32 D = dim_, np = (D + 1) * (D + 2) / 2;
    KALDI_ASSERT(stats->Dim() == D && static_cast<int32>(stats->G_.size()) == D);
    std::vector<double> K(static_cast<size_t>(D) * (D + 1)), G(static_cast<size_t>(D) * np);
    Check(vbgpu_fmllr_download(h_, spk, &stats->beta_, K.data(), G.data()), "vbgpu_fmllr_download");
    for (int32 i = 0; i < D; i++) {
      for (int32 k = 0; k <= D; k++) stats->K_(i, k) = K[static_cast<size_t>(i) * (D + 1) + k];
      kaldi::SubVector<double> packed(G.data() + static_cast<size_t>(i) * np, np);  // SpMatrix packing
      stats->G_[i].CopyFromVec(packed);
    }
  }
  void SetZero() { Check(vbgpu_fmllr_zero(h_), "vbgpu_fmllr_zero"); }

 private:
  int32 dim_;
  vbgpu_fmllr_t h_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(FmllrAccsGpu);
};

// ---- MlltAccs (gmm-acc-mllt.cpp:100-112) -------------------------------------------------------------------------------------
// AccumulateFromGmm for every frame of an utterance on the device, added straight into the reference's accumulator
// (rand_prune = 0); MlltAccs::Update / Write then run unchanged.  Returns the utterance's sum of weight * loglike.
inline double MlltAccumulateForUtterance(const GpuAmDiagGmm &am, const kaldi::MatrixBase<BaseFloat> &feats,
                                         const std::vector<int32> &pdf_ids, kaldi::MlltAccs *accs,
                                         const std::vector<BaseFloat> *weights = NULL) {
  const int32 D = am.Dim(), np = D * (D + 1) / 2;
  KALDI_ASSERT(static_cast<int32>(pdf_ids.size()) == feats.NumRows() && accs->Dim() == D);
  double like = 0.0, beta = 0.0;
  if (feats.NumRows() == 0) return like;
  std::vector<double> G(static_cast<size_t>(D) * np, 0.0);
  Check(vbgpu_mllt_accumulate(am.handle(), feats.Data(), feats.NumRows(), feats.Stride(), pdf_ids.data(),
                              weights ? weights->data() : NULL, &beta, G.data(), &like),
        "vbgpu_mllt_accumulate");
  accs->beta_ += beta;
  for (int32 j = 0; j < D; j++) {
    kaldi::SpMatrix<double> g(D);
    g.CopyFromVec(kaldi::SubVector<double>(G.data() + static_cast<size_t>(j) * np, np));  // SpMatrix packing
    accs->G_[j].AddSp(1.0, g);
  }
  return like;
}

// ---- AccumAmDiagGmm -------------------------------------------------------------------------------------------------------
class AccumAmDiagGmmGpu {
 public:
  explicit AccumAmDiagGmmGpu(const GpuAmDiagGmm &am) : am_(am), h_(NULL) {
    Check(vbgpu_acc_create(am.handle(), &h_), "vbgpu_acc_create");  // Init(model, kGmmAll)
  }
  ~AccumAmDiagGmmGpu() { vbgpu_acc_destroy(h_); }
  void SetZero() { Check(vbgpu_acc_zero(h_), "vbgpu_acc_zero"); }
  // AccumulateForGmm over a whole utterance: pdf_ids[t] = TransitionIdToPdf(alignment[t]) (gmm-acc-stats-ali.cpp:89-94).
  // Returns the utterance's total log-likelihood, as the sum of the reference's per-frame return values.
  double AccumulateForUtterance(const kaldi::MatrixBase<BaseFloat> &feats, const std::vector<int32> &pdf_ids,
                                const std::vector<BaseFloat> *weights = NULL) {
    KALDI_ASSERT(static_cast<int32>(pdf_ids.size()) == feats.NumRows());
    double like = 0.0;
    if (feats.NumRows() == 0) return like;
    Check(vbgpu_acc_accumulate(h_, feats.Data(), NULL, feats.NumRows(), feats.Stride(), pdf_ids.data(),
                               weights ? weights->data() : NULL, &like),
          "vbgpu_acc_accumulate");
    return like;
  }
  // Adds the device statistics into a reference accumulator (already Init'ed on the same model with kGmmAll) through
  // AccumDiagGmm::AddStatsForComponent (mle-diag-gmm.cc:158-168): Write() / gmm-sum-accs / MleAmDiagGmmUpdate then run
  // unchanged.  Returns (total log-likelihood, total frames) for the caller's bookkeeping: AccumAmDiagGmm keeps its
  // total_log_like_ / total_frames_ private, so an .acc file written by acc->Write() after AddTo() carries 0 for both and
  // gmm-sum-accs / gmm-est would log "avg like per frame" as 0/0.  When the file itself is the product, use WriteAccFile().
  std::pair<double, double> AddTo(kaldi::AccumAmDiagGmm *acc) const {
    const int32 N = am_.NumGauss(), D = am_.Dim(), P = am_.NumPdfs();
    std::vector<double> occ(N), mean(static_cast<size_t>(N) * D), var(static_cast<size_t>(N) * D);
    double like = 0.0, frames = 0.0;
    Check(vbgpu_acc_download(h_, occ.data(), mean.data(), var.data(), &like, &frames), "vbgpu_acc_download");
    KALDI_ASSERT(acc->NumAccs() == P);
    const std::vector<int32_t> &po = am_.PdfOffsets();
    for (int32 p = 0; p < P; p++)
      for (int32 g = po[p]; g < po[p + 1]; g++) {
        kaldi::SubVector<double> m(mean.data() + static_cast<size_t>(g) * D, D), v(var.data() + static_cast<size_t>(g) * D, D);
        acc->GetAcc(p).AddStatsForComponent(g - po[p], occ[g], m, v);
      }
    return std::make_pair(like, frames);
  }
  // The statistics file gmm-acc-stats-ali writes (gmm-acc-stats-ali.cpp:124-128: transition_accs.Write(os, binary) then
  // gmm_accs.Write(os, binary)), byte for byte in the reference's format and INCLUDING <total_like> / <total_frames>.
  void WriteAccFile(std::ostream &os, const kaldi::Vector<double> &transition_accs) const {
    const int32 N = am_.NumGauss(), D = am_.Dim(), P = am_.NumPdfs();
    std::vector<double> occ(N), mean(static_cast<size_t>(N) * D), var(static_cast<size_t>(N) * D);
    double like = 0.0, frames = 0.0;
    Check(vbgpu_acc_download(h_, occ.data(), mean.data(), var.data(), &like, &frames), "vbgpu_acc_download");
    const std::vector<int32_t> &po = am_.PdfOffsets();
    const int64_t bytes = vbgpu_io_write_acc(P, D, po.data(), transition_accs.Data(), transition_accs.Dim(), occ.data(),
                                             mean.data(), var.data(), like, frames, NULL, 0);
    if (bytes < 0) KALDI_ERR << "vbgpu_io_write_acc: " << vbgpu_last_error();
    std::vector<char> buf(static_cast<size_t>(bytes));
    vbgpu_io_write_acc(P, D, po.data(), transition_accs.Data(), transition_accs.Dim(), occ.data(), mean.data(), var.data(),
                       like, frames, buf.data(), bytes);
    os.write(buf.data(), bytes);
  }
  vbgpu_acc_t handle() const { return h_; }

 private:
  const GpuAmDiagGmm &am_;
  vbgpu_acc_t h_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(AccumAmDiagGmmGpu);
};

}  // namespace vbgpu
#endif  // VBGPU_KALDI_H_
