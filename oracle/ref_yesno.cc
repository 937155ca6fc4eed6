// oracle/ref_yesno.cc — TEST INFRASTRUCTURE ONLY.
//
// The yes-no scenario of BASELINE.json configs[0] / SURVEY.md §8d cfg 1, end to end inside ONE process, driven through
// the reference's UNMODIFIED Kaldi classes (compiled from /root/reference into oracle/_ref/libkaldi_ref.a):
//   synthetic 8 kHz "yes"/"no" utterances -> 13 MFCC + per-utterance CMVN + delta/delta-delta (39)
//   -> monophone model (SIL, Y, N; 3 states each) trained the way VB/scr/steps/train_mono.cpp chains the tools:
//      gmm-init-mono, compile-train-graphs, align-equal-compiled, gmm-acc-stats-ali, gmm-est (mix-up), gmm-align-compiled
//   -> decoding of held-out utterances with LatticeFasterDecoder on a word-loop HCLG, and forced alignment with FasterDecoder.
// The yes-no data / conf of the real recipe live in an external repository that is not in the reference tree, so the
// corpus is synthesised (tone-burst "words" + noise), as SURVEY.md §8c prescribes.
//
// Modes:
//   ref_yesno cpu            reference only: trains, decodes, prints WER (sanity of the scenario; runs without a GPU)
//   ref_yesno gpu            additionally runs every hot-path stage through libvbgpu.so via include/vbgpu_kaldi.h (the
//                            C++ drop-in adaptors) and compares: features, log-likelihoods, EM statistics of every
//                            training iteration, 1-best word sequences and alignments (these two must be IDENTICAL).
// Output: one JSON line on stdout.  The product never links or loads this file.
#include <cmath>
#include <cstdio>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "base/kaldi-common.h"
#include "decoder/faster-decoder.h"
#include "decoder/lattice-faster-decoder.h"
#include "decoder/training-graph-compiler.h"
#include "feat/feature-functions.h"
#include "feat/feature-mfcc.h"
#include "fstext/fstext-lib.h"
#include "gmm/am-diag-gmm.h"
#include "gmm/decodable-am-diag-gmm.h"
#include "gmm/mle-am-diag-gmm.h"
#include "transform/fmllr-diag-gmm.h"
#include "hmm/hmm-topology.h"
#include "hmm/transition-model.h"
#include "lat/kaldi-lattice.h"
#include "lat/lattice-functions.h"
#include "transform/cmvn.h"
#include "tree/context-dep.h"

#include "feat/pitch-functions.h"
#include "../include/vbgpu_kaldi.h"

using namespace kaldi;
typedef fst::VectorFst<fst::StdArc> StdFst;

namespace {

const int kSil = 1, kY = 2, kN = 3;      // phones
const int kYes = 1, kNo = 2;             // words
const float kFs = 8000.0f;
const BaseFloat kAcwt = 0.083333f;       // decode acoustic scale (decode_gmm.cpp default), 0.1 in training

struct Utt {
  std::vector<int32> words;
  Vector<BaseFloat> wave;
  Matrix<BaseFloat> feats, feats_gpu;
  StdFst graph;
  std::vector<int32> ali;
};

// A "word" is 0.35-0.5 s of three harmonics with a word-specific glide; gaps are 0.1-0.3 s of noise only.
void Synthesise(std::mt19937 *rng, int n_words, Utt *u) {
  std::uniform_real_distribution<float> U(0.0f, 1.0f);
  std::normal_distribution<float> G(0.0f, 1.0f);
  std::vector<float> x;
  auto silence = [&](float secs) {
    for (int i = 0; i < (int)(secs * kFs); i++) x.push_back(0.0f);
  };
  silence(0.15f + 0.2f * U(*rng));
  for (int w = 0; w < n_words; w++) {
    const int word = U(*rng) < 0.5f ? kYes : kNo;
    u->words.push_back(word);
    const float dur = 0.35f + 0.15f * U(*rng), amp = 6000.0f + 4000.0f * U(*rng);
    const int n = (int)(dur * kFs);
    double ph1 = 0, ph2 = 0, ph3 = 0;
    for (int i = 0; i < n; i++) {
      const float r = (float)i / n, env = sinf(3.14159265f * r);
      const float f1 = word == kYes ? 300.0f + 250.0f * r : 650.0f - 300.0f * r;
      const float f2 = word == kYes ? 2300.0f - 300.0f * r : 1000.0f + 100.0f * r;
      const float f3 = word == kYes ? 3000.0f : 2500.0f - 400.0f * r;
      ph1 += 2 * M_PI * f1 / kFs, ph2 += 2 * M_PI * f2 / kFs, ph3 += 2 * M_PI * f3 / kFs;
      x.push_back(amp * env * (0.6f * sinf((float)ph1) + 0.3f * sinf((float)ph2) + 0.1f * sinf((float)ph3)));
    }
    silence(0.1f + 0.2f * U(*rng));
  }
  u->wave.Resize(x.size());
  for (size_t i = 0; i < x.size(); i++) {
    float v = x[i] + 500.0f * G(*rng);
    v = std::max(-32768.0f, std::min(32767.0f, v));
    u->wave(i) = roundf(v);  // what WaveData::Read hands out: integers in int16 range (wave-reader.cc:302-309)
  }
}

MfccOptions MfccOpts() {
  MfccOptions o;
  o.frame_opts.samp_freq = kFs;
  o.frame_opts.dither = 0.0f;
  o.use_energy = false;
  return o;
}

// compute-mfcc-feats | apply-cmvn (per utterance) | add-deltas, with the reference's classes.
void RefFeatures(Mfcc *mfcc, const VectorBase<BaseFloat> &wave, Matrix<BaseFloat> *out) {
  Matrix<BaseFloat> raw;
  mfcc->ComputeFeatures(wave, kFs, 1.0f, &raw);
  Matrix<double> stats;
  InitCmvnStats(raw.NumCols(), &stats);
  AccCmvnStats(raw, NULL, &stats);
  ApplyCmvn(stats, false, &raw);
  ComputeDeltas(DeltaFeaturesOptions(), raw, out);
}

// The same chain through libvbgpu.so (include/vbgpu_kaldi.h).
void GpuFeatures(vbgpu::GpuMfcc *mfcc, vbgpu::GpuFeaturePipeline *fp, const VectorBase<BaseFloat> &wave,
                 Matrix<BaseFloat> *out) {
  Matrix<BaseFloat> raw;
  mfcc->ComputeFeatures(wave, kFs, 1.0f, &raw);
  Matrix<double> stats(2, raw.NumCols() + 1, kSetZero, kStrideEqualNumCols);
  const int64_t fo[2] = {0, raw.NumRows()};
  vbgpu::Check(vbgpu_cmvn_stats(fp->handle(), raw.Data(), raw.Stride(), fo, 1, NULL, 1, stats.Data()), "vbgpu_cmvn_stats");
  fp->Run(raw, stats, NULL, out);
}

HmmTopology MakeTopology() {
  std::ostringstream os;
  os << "<Topology>\n<TopologyEntry>\n<ForPhones> " << kSil << " " << kY << " " << kN << " </ForPhones>\n";
  for (int s = 0; s < 3; s++)
    os << "<State> " << s << " <PdfClass> " << s << " <Transition> " << s << " 0.75 <Transition> " << s + 1 << " 0.25 </State>\n";
  os << "<State> 3 </State>\n</TopologyEntry>\n</Topology>\n";
  std::istringstream is(os.str());
  HmmTopology topo;
  topo.Read(is, false);
  return topo;
}

// Lexicon with optional silence (probability 0.5) at the start and after every word, as utils/make_lexicon_fst does:
// state 0 = start, 1 = loop (final), 2 = "silence next".
StdFst *MakeLexicon() {
  StdFst *L = new StdFst;
  for (int i = 0; i < 3; i++) L->AddState();
  L->SetStart(0);
  const float c = -logf(0.5f);
  L->AddArc(0, fst::StdArc(0, 0, c, 1));
  L->AddArc(0, fst::StdArc(0, 0, c, 2));
  L->AddArc(2, fst::StdArc(kSil, 0, 0.0f, 1));
  const int phone_of[3] = {0, kY, kN};
  for (int w = kYes; w <= kNo; w++) {
    L->AddArc(1, fst::StdArc(phone_of[w], w, c, 1));
    L->AddArc(1, fst::StdArc(phone_of[w], w, c, 2));
  }
  L->SetFinal(1, fst::StdArc::Weight::One());
  return L;
}

StdFst WordLoop() {
  StdFst G;
  G.AddState();
  G.SetStart(0);
  G.AddArc(0, fst::StdArc(kYes, kYes, -logf(0.5f), 0));
  G.AddArc(0, fst::StdArc(kNo, kNo, -logf(0.5f), 0));
  G.SetFinal(0, fst::StdArc::Weight::One());
  return G;
}

// gmm-align-compiled for one utterance: FasterDecoder, beam 10 then 40 (AlignConfig defaults, decoder-wrappers.cc).
bool Align(const StdFst &graph, DecodableInterface *dec, std::vector<int32> *ali) {
  for (BaseFloat beam : {10.0f, 40.0f}) {
    FasterDecoderOptions o;
    o.beam = beam;
    FasterDecoder d(graph, o);
    d.Decode(dec);
    if (!d.ReachedFinal()) continue;
    fst::VectorFst<LatticeArc> best;
    d.GetBestPath(&best);
    std::vector<int32> words;
    LatticeWeight w;
    GetLinearSymbolSequence(best, ali, &words, &w);
    return true;
  }
  return false;
}

bool Decode(const StdFst &hclg, DecodableInterface *dec, std::vector<int32> *ali, std::vector<int32> *words) {
  LatticeFasterDecoderConfig c;  // gmm-latgen-faster defaults of decode_gmm.cpp: beam 13, lattice-beam 6, max-active 7000
  c.beam = 13.0f;
  c.lattice_beam = 6.0f;
  c.max_active = 7000;
  LatticeFasterDecoder d(hclg, c);
  if (!d.Decode(dec)) return false;
  Lattice best;
  d.GetBestPath(&best);
  LatticeWeight w;
  GetLinearSymbolSequence(best, ali, words, &w);
  return true;
}

// gmm-rescore-lattice (VB/src/gmmbin/gmm-rescore-lattice.cpp): the raw lattice of an utterance with its acoustic scores
// removed, then RescoreLattice (lat/lattice-functions.cc:1360-1401) with the given decodable (scale 1.0).
bool RawLatticeWithoutAcoustics(const StdFst &hclg, DecodableInterface *dec, Lattice *lat) {
  LatticeFasterDecoderConfig c;
  c.beam = 13.0f;
  c.lattice_beam = 6.0f;
  c.max_active = 7000;
  LatticeFasterDecoder d(hclg, c);
  if (!d.Decode(dec) || !d.GetRawLattice(lat)) return false;
  fst::ScaleLattice(fst::AcousticLatticeScale(0.0), lat);
  return true;
}

int EditDistance(const std::vector<int32> &a, const std::vector<int32> &b) {
  std::vector<std::vector<int> > d(a.size() + 1, std::vector<int>(b.size() + 1));
  for (size_t i = 0; i <= a.size(); i++) d[i][0] = i;
  for (size_t j = 0; j <= b.size(); j++) d[0][j] = j;
  for (size_t i = 1; i <= a.size(); i++)
    for (size_t j = 1; j <= b.size(); j++)
      d[i][j] = std::min(std::min(d[i - 1][j] + 1, d[i][j - 1] + 1), d[i - 1][j - 1] + (a[i - 1] != b[j - 1]));
  return d[a.size()][b.size()];
}

double RelErr(double got, double want, double floor_) { return std::fabs(got - want) / std::max(std::fabs(want), floor_); }

}  // namespace

int main(int argc, char **argv) {
  try {
    const bool gpu = argc > 1 && std::string(argv[1]) == "gpu";
    const int n_train = 40, n_test = 20, n_iters = 14, max_gauss = 120;
    std::mt19937 rng(1234 + 1);
    std::vector<Utt> train(n_train), test(n_test);
    for (auto &u : train) Synthesise(&rng, 6, &u);
    for (auto &u : test) Synthesise(&rng, 8, &u);

    // ---- features ----
    Mfcc mfcc(MfccOpts());
    vbgpu::GpuMfcc *gmfcc = NULL;
    vbgpu::GpuFeaturePipeline *gfp = NULL;
    if (gpu) {
      gmfcc = new vbgpu::GpuMfcc(MfccOpts());
      gfp = new vbgpu::GpuFeaturePipeline(13, false, DeltaFeaturesOptions());
    }
    double feat_err = 0.0, feat_scale = 0.0;
    int64 n_frames = 0;
    for (std::vector<Utt> *set : {&train, &test})
      for (auto &u : *set) {
        RefFeatures(&mfcc, u.wave, &u.feats);
        n_frames += u.feats.NumRows();
        if (gpu) {
          GpuFeatures(gmfcc, gfp, u.wave, &u.feats_gpu);
          KALDI_ASSERT(u.feats_gpu.NumRows() == u.feats.NumRows() && u.feats_gpu.NumCols() == u.feats.NumCols());
          Matrix<BaseFloat> diff(u.feats_gpu);
          diff.AddMat(-1.0f, u.feats);
          feat_err = std::max(feat_err, (double)std::max(diff.Max(), -diff.Min()));
          feat_scale = std::max(feat_scale, (double)std::max(u.feats.Max(), -u.feats.Min()));
        }
      }
    const int32 dim = train[0].feats.NumCols();

    // ---- Kaldi pitch of the test utterances: the reference's ComputeKaldiPitch / ProcessPitch vs vbgpu::GpuPitch ----
    long long pitch_frames = 0, pitch_same = 0;
    double pitch_nccf_err = 0.0, pitch_rel_err = 0.0, process_err = 0.0;
    if (gpu) {
      PitchExtractionOptions popts;
      ProcessPitchOptions ppopts;
      ppopts.delta_pitch_noise_stddev = 0.0;  // the reference's noise comes from rand()
      vbgpu::GpuPitch gpitch(popts);
      for (auto &u : test) {
        Matrix<BaseFloat> want, got, pw, pg;
        ComputeKaldiPitch(popts, u.wave, &want);
        gpitch.ComputeKaldiPitch(u.wave, &got);
        KALDI_ASSERT(want.NumRows() == got.NumRows() && got.NumCols() == 2);
        for (int32 t = 0; t < want.NumRows(); t++) {
          pitch_frames++;
          const double rel = std::fabs(got(t, 1) - want(t, 1)) / want(t, 1);
          pitch_rel_err = std::max(pitch_rel_err, rel);
          if (rel <= 1e-6) {
            pitch_same++;
            pitch_nccf_err = std::max(pitch_nccf_err, (double)std::fabs(got(t, 0) - want(t, 0)));
          }
        }
        ProcessPitch(ppopts, want, &pw);
        gpitch.ProcessPitch(ppopts, want, &pg);
        KALDI_ASSERT(pw.NumRows() == pg.NumRows() && pw.NumCols() == pg.NumCols());
        pg.AddMat(-1.0f, pw);
        process_err = std::max(process_err, (double)std::max(pg.Max(), -pg.Min()));
      }
    }

    // ---- gmm-init-mono ----
    HmmTopology topo = MakeTopology();
    const std::vector<int32> &phones = topo.GetPhones();
    std::vector<int32> phone2num_pdf_classes(1 + phones.back());
    for (size_t i = 0; i < phones.size(); i++) phone2num_pdf_classes[phones[i]] = topo.NumPdfClasses(phones[i]);
    ContextDependency *ctx_dep = MonophoneContextDependency(phones, phone2num_pdf_classes);
    AmDiagGmm am;
    {
      Vector<double> mean(dim), var(dim);
      double count = 0.0;
      for (int k = 0; k < 10; k++)
        for (int32 t = 0; t < train[k].feats.NumRows(); t++) {
          Vector<double> row(train[k].feats.Row(t));
          mean.AddVec(1.0, row);
          var.AddVec2(1.0, row);
          count += 1.0;
        }
      mean.Scale(1.0 / count);
      var.Scale(1.0 / count);
      var.AddVec2(-1.0, mean);
      var.InvertElements();
      DiagGmm g;
      g.Resize(1, dim);
      Matrix<BaseFloat> inv_var(1, dim), mu(1, dim);
      inv_var.Row(0).CopyFromVec(Vector<BaseFloat>(var));
      mu.Row(0).CopyFromVec(Vector<BaseFloat>(mean));
      Vector<BaseFloat> w(1);
      w.Set(1.0f);
      g.SetInvVarsAndMeans(inv_var, mu);
      g.SetWeights(w);
      g.ComputeGconsts();
      for (int32 i = 0; i < ctx_dep->NumPdfs(); i++) am.AddPdf(g);
    }
    TransitionModel tm(*ctx_dep, topo);

    // ---- compile-train-graphs, align-equal-compiled ----
    TrainingGraphCompilerOptions gopts(1.0f, 0.1f, true);
    TrainingGraphCompiler gc(tm, *ctx_dep, MakeLexicon(), std::vector<int32>(), gopts);
    for (size_t k = 0; k < train.size(); k++) {
      Utt &u = train[k];
      if (!gc.CompileGraphFromText(u.words, &u.graph)) KALDI_ERR << "graph compilation failed";
      StdFst path;
      if (!fst::EqualAlign(u.graph, u.feats.NumRows(), 777 + (int)k, &path)) KALDI_ERR << "EqualAlign failed";
      std::vector<int32> words;
      fst::StdArc::Weight w;
      GetLinearSymbolSequence(path, &u.ali, &words, &w);
    }

    // ---- EM iterations: gmm-acc-stats-ali, gmm-est --mix-up, gmm-align-compiled ----
    double stats_err = 0.0, acc_like_err = 0.0;
    int32 cur_gauss = am.NumGauss();
    const int32 inc = (max_gauss - cur_gauss) / 8;
    for (int it = 0; it < n_iters; it++) {
      if (it > 0 && (it <= 4 || it % 2 == 0))
        for (auto &u : train) {
          DecodableAmDiagGmmScaled dec(am, tm, u.feats, 0.1f);
          if (!Align(u.graph, &dec, &u.ali)) KALDI_ERR << "alignment failed";
        }
      AccumAmDiagGmm acc;
      acc.Init(am, kGmmAll);
      Vector<double> trans_acc;
      tm.InitStats(&trans_acc);
      double tot_like = 0.0;
      for (auto &u : train)
        for (size_t t = 0; t < u.ali.size(); t++) {
          tm.Accumulate(1.0f, u.ali[t], &trans_acc);
          tot_like += acc.AccumulateForGmm(am, u.feats.Row(t), tm.TransitionIdToPdf(u.ali[t]), 1.0f);
        }
      if (gpu) {  // the same E-step on the device, from the device-computed features, through the C++ adaptors
        vbgpu::GpuAmDiagGmm gam(am);
        vbgpu::AccumAmDiagGmmGpu gacc(gam);
        double glike = 0.0;
        for (auto &u : train) {
          std::vector<int32> pdfs(u.ali.size());
          for (size_t t = 0; t < u.ali.size(); t++) pdfs[t] = tm.TransitionIdToPdf(u.ali[t]);
          glike += gacc.AccumulateForUtterance(u.feats_gpu, pdfs);
        }
        AccumAmDiagGmm acc2;
        acc2.Init(am, kGmmAll);
        gacc.AddTo(&acc2);
        acc_like_err = std::max(acc_like_err, RelErr(glike, tot_like, 1.0));
        for (int32 p = 0; p < am.NumPdfs(); p++) {
          const AccumDiagGmm &a = acc.GetAcc(p), &b = acc2.GetAcc(p);
          // scale of a statistic = the pdf's largest entry of that kind (tiny components carry only rounding noise)
          const double so = std::max(a.occupancy().Max(), 1.0), sm = std::max(a.mean_accumulator().LargestAbsElem(), 1.0),
                       sv = std::max(a.variance_accumulator().LargestAbsElem(), 1.0);
          for (int32 g = 0; g < a.NumGauss(); g++) {
            stats_err = std::max(stats_err, std::fabs(a.occupancy()(g) - b.occupancy()(g)) / so);
            for (int32 d = 0; d < dim; d++) {
              stats_err = std::max(stats_err, std::fabs(a.mean_accumulator()(g, d) - b.mean_accumulator()(g, d)) / sm);
              stats_err = std::max(stats_err, std::fabs(a.variance_accumulator()(g, d) - b.variance_accumulator()(g, d)) / sv);
            }
          }
        }
      }
      BaseFloat objf, count;
      tm.MleUpdate(trans_acc, MleTransitionUpdateConfig(), &objf, &count);
      MleAmDiagGmmUpdate(MleDiagGmmOptions(), acc, kGmmAll, &am, &objf, &count);
      if (it < 9) {
        cur_gauss = std::min(max_gauss, cur_gauss + inc);
        Vector<BaseFloat> occs(am.NumPdfs());
        for (int32 p = 0; p < am.NumPdfs(); p++) occs(p) = acc.GetAcc(p).occupancy().Sum();
        am.SplitByCount(occs, cur_gauss, 0.01f, 0.25f, 20.0f);
        am.ComputeGconsts();
      }
    }

    // ---- decode + align the held-out set: reference decodable vs GPU decodable inside the reference's decoders ----
    StdFst hclg;
    {
      StdFst G = WordLoop();
      if (!gc.CompileGraph(G, &hclg)) KALDI_ERR << "HCLG compilation failed";
    }
    vbgpu::GpuAmDiagGmm *gam = gpu ? new vbgpu::GpuAmDiagGmm(am) : NULL;
    int errs = 0, n_ref_words = 0, same_words = 0, same_ali = 0, same_forced = 0, gpu_errs = 0;
    double ll_err = 0.0, ll_mag = 0.0;
    for (auto &u : test) {
      std::vector<int32> ali, words, fali;
      DecodableAmDiagGmmScaled dec(am, tm, u.feats, kAcwt);
      if (!Decode(hclg, &dec, &ali, &words)) KALDI_ERR << "decoding failed";
      errs += EditDistance(words, u.words);
      n_ref_words += u.words.size();
      StdFst fgraph;
      gc.CompileGraphFromText(u.words, &fgraph);
      DecodableAmDiagGmmScaled fdec(am, tm, u.feats, 0.1f);
      if (!Align(fgraph, &fdec, &fali)) KALDI_ERR << "forced alignment failed";
      if (gpu) {
        std::vector<int32> gali, gwords, gfali;
        vbgpu::DecodableAmDiagGmmGpu gdec(*gam, tm, u.feats_gpu, kAcwt);
        if (!Decode(hclg, &gdec, &gali, &gwords)) KALDI_ERR << "decoding (gpu decodable) failed";
        gpu_errs += EditDistance(gwords, u.words);
        same_words += gwords == words;
        same_ali += gali == ali;
        vbgpu::DecodableAmDiagGmmGpu gfdec(*gam, tm, u.feats_gpu, 0.1f);
        if (!Align(fgraph, &gfdec, &gfali)) KALDI_ERR << "forced alignment (gpu decodable) failed";
        same_forced += gfali == fali;
        // dense log-likelihoods against the reference's per-(frame, pdf) arithmetic, on the SAME (reference) features
        Matrix<BaseFloat> gll;
        gam->LogLikelihoods(u.feats, &gll);
        for (int32 t = 0; t < u.feats.NumRows(); t += 3)
          for (int32 p = 0; p < am.NumPdfs(); p++) {
            const BaseFloat want = am.LogLikelihood(p, u.feats.Row(t));
            ll_err = std::max(ll_err, (double)std::fabs(gll(t, p) - want));
            ll_mag = std::max(ll_mag, (double)std::fabs(want));
          }
      }
    }
    // ---- SURVEY 8f n3 / n1 through the C++ adaptors: the whole test set scored in ONE call and aligned through per-utterance
    //      views; fMLLR statistics of two "speakers" against the reference's FmllrDiagGmmAccs and its own solver ----
    int batch_forced_same = 0, rescored_same = 0, subset_forced_same = 0, gather_same = 0;
    long long rescored_arcs = 0, subset_floats = 0, dense_floats = 0, gather_arcs = 0;
    double gather_err = 0.0;
    double fmllr_stats_err = 0.0, fmllr_xform_err = 0.0, rescore_err = 0.0;
    if (gpu) {
      std::vector<const MatrixBase<BaseFloat> *> fl;
      for (auto &u : test) fl.push_back(&u.feats_gpu);
      vbgpu::BatchDecodableAmDiagGmmGpu batch(*gam, tm, fl, 0.1f);
      std::vector<std::vector<int32> > alis(test.size());
      for (size_t i = 0; i < test.size(); i++) {
        StdFst fgraph;
        gc.CompileGraphFromText(test[i].words, &fgraph);
        vbgpu::DecodableAmDiagGmmGpu single(*gam, tm, test[i].feats_gpu, 0.1f);
        std::vector<int32> a1;
        if (!Align(fgraph, batch.Utterance(i), &alis[i]) || !Align(fgraph, &single, &a1)) KALDI_ERR << "batch alignment failed";
        batch_forced_same += alis[i] == a1;
      }
      // lattice rescoring (gmm-rescore-lattice): the reference's RescoreLattice driven by its own decodable and by the
      // batch's views (acoustic scale 1.0); every arc must carry the same acoustic cost within the log-likelihood tolerance
      {
        std::vector<const MatrixBase<BaseFloat> *> fr;
        for (auto &u : test) fr.push_back(&u.feats);
        vbgpu::BatchDecodableAmDiagGmmGpu rbatch(*gam, tm, fr, 1.0f);
        for (size_t i = 0; i < test.size(); i++) {
          DecodableAmDiagGmmScaled dec(am, tm, test[i].feats, kAcwt);
          Lattice lat;
          if (!RawLatticeWithoutAcoustics(hclg, &dec, &lat)) KALDI_ERR << "lattice generation failed";
          Lattice lat_ref(lat), lat_gpu(lat);
          DecodableAmDiagGmmScaled rdec(am, tm, test[i].feats, 1.0f);
          if (!RescoreLattice(&rdec, &lat_ref) || !RescoreLattice(rbatch.Utterance(i), &lat_gpu)) KALDI_ERR << "rescoring failed";
          KALDI_ASSERT(lat_ref.NumStates() == lat_gpu.NumStates());
          for (int32 st = 0; st < lat_ref.NumStates(); st++) {
            fst::ArcIterator<Lattice> a(lat_ref, st), b(lat_gpu, st);
            for (; !a.Done(); a.Next(), b.Next(), rescored_arcs++)
              rescore_err = std::max(rescore_err, (double)std::fabs(a.Value().weight.Value2() - b.Value().weight.Value2()));
          }
          Lattice best_ref, best_gpu;
          fst::ShortestPath(lat_ref, &best_ref);
          fst::ShortestPath(lat_gpu, &best_gpu);
          std::vector<int32> a1, w1, a2, w2;
          LatticeWeight lw;
          GetLinearSymbolSequence(best_ref, &a1, &w1, &lw);
          GetLinearSymbolSequence(best_gpu, &a2, &w2, &lw);
          rescored_same += (a1 == a2 && w1 == w2);
        }
      }
      // the sparse forms (include/vbgpu_kaldi.h): forced alignment on each utterance's own pdf subset, lattice rescoring on
      // the arcs' (frame, pdf) pairs only — what crosses PCIe is counted against the dense matrix
      {
        std::vector<std::vector<int32> > subsets(test.size());
        std::vector<StdFst> graphs(test.size());
        for (size_t i = 0; i < test.size(); i++) {
          gc.CompileGraphFromText(test[i].words, &graphs[i]);
          vbgpu::PdfsOfGraph(graphs[i], tm, &subsets[i]);
        }
        vbgpu::BatchSubsetDecodableAmDiagGmmGpu sbatch(*gam, tm, fl, subsets, 0.1f);
        for (size_t i = 0; i < test.size(); i++) {
          std::vector<int32> a;
          if (!Align(graphs[i], sbatch.Utterance(i), &a)) KALDI_ERR << "subset alignment failed";
          subset_forced_same += a == alis[i];
          dense_floats += (long long)test[i].feats_gpu.NumRows() * am.NumPdfs();
        }
        subset_floats = sbatch.NumScores();
        std::vector<const MatrixBase<BaseFloat> *> fr;
        for (auto &u : test) fr.push_back(&u.feats);
        std::vector<Lattice> lats(test.size());
        std::vector<vbgpu::GatherDecodable *> gdecs;
        for (size_t i = 0; i < test.size(); i++) {
          DecodableAmDiagGmmScaled dec(am, tm, test[i].feats, kAcwt);
          if (!RawLatticeWithoutAcoustics(hclg, &dec, &lats[i])) KALDI_ERR << "lattice generation failed";
          gdecs.push_back(new vbgpu::GatherDecodable(tm, test[i].feats.NumRows(), 1.0f));
          Lattice dry(lats[i]);
          if (!RescoreLattice(gdecs.back(), &dry)) KALDI_ERR << "recording run failed";
          gather_arcs += (long long)gdecs.back()->NumArcs();
        }
        vbgpu::GatherScorer::Score(*gam, fr, gdecs);
        for (size_t i = 0; i < test.size(); i++) {
          Lattice lat_ref(lats[i]), lat_gpu(lats[i]);
          DecodableAmDiagGmmScaled rdec(am, tm, test[i].feats, 1.0f);
          if (!RescoreLattice(&rdec, &lat_ref) || !RescoreLattice(gdecs[i], &lat_gpu)) KALDI_ERR << "gather rescoring failed";
          for (int32 st = 0; st < lat_ref.NumStates(); st++) {
            fst::ArcIterator<Lattice> a(lat_ref, st), b(lat_gpu, st);
            for (; !a.Done(); a.Next(), b.Next())
              gather_err = std::max(gather_err, (double)std::fabs(a.Value().weight.Value2() - b.Value().weight.Value2()));
          }
          Lattice best_ref, best_gpu;
          fst::ShortestPath(lat_ref, &best_ref);
          fst::ShortestPath(lat_gpu, &best_gpu);
          std::vector<int32> a1, w1, a2, w2;
          LatticeWeight lw;
          GetLinearSymbolSequence(best_ref, &a1, &w1, &lw);
          GetLinearSymbolSequence(best_gpu, &a2, &w2, &lw);
          gather_same += (a1 == a2 && w1 == w2);
          delete gdecs[i];
        }
      }
      // fMLLR statistics from those alignments: utterances alternate between two speakers
      const int32 D = am.Dim();
      int64_t T = 0;
      for (auto &u : test) T += u.feats_gpu.NumRows();
      Matrix<BaseFloat> packed(T, D);
      std::vector<int32> pdfs, u2s;
      std::vector<int64_t> fo(1, 0);
      for (size_t i = 0; i < test.size(); i++) {
        packed.RowRange(fo.back(), test[i].feats_gpu.NumRows()).CopyFromMat(test[i].feats_gpu);
        for (size_t t = 0; t < alis[i].size(); t++) pdfs.push_back(tm.TransitionIdToPdf(alis[i][t]));
        fo.push_back(fo.back() + test[i].feats_gpu.NumRows());
        u2s.push_back(i % 2);
      }
      vbgpu::FmllrAccsGpu gf(*gam, 2);
      gf.Accumulate(packed, pdfs, fo, u2s);
      for (int32 spk = 0; spk < 2; spk++) {
        FmllrDiagGmmAccs want(D), got(D);
        for (size_t i = spk; i < test.size(); i += 2)
          for (size_t t = 0; t < alis[i].size(); t++)
            want.AccumulateForGmm(am.GetPdf(tm.TransitionIdToPdf(alis[i][t])), test[i].feats_gpu.Row(t), 1.0);
        Matrix<BaseFloat> xw(D, D + 1), xg(D, D + 1);
        xw.SetUnit();
        xg.SetUnit();
        FmllrOptions fopts;
        fopts.min_count = 100.0;
        want.Update(fopts, &xw, NULL, NULL);  // (also commits the reference's pending frame)
        gf.CopyTo(spk, &got);
        got.Update(fopts, &xg, NULL, NULL);
        double gmax = 0.0;
        for (int32 i = 0; i < D; i++) gmax = std::max(gmax, (double)want.G_[i].Max());
        for (int32 i = 0; i < D; i++)
          for (int32 j = 0; j <= D; j++)
            for (int32 k = 0; k <= j; k++)
              fmllr_stats_err = std::max(fmllr_stats_err, std::fabs(got.G_[i](j, k) - want.G_[i](j, k)) / gmax);
        fmllr_stats_err = std::max(fmllr_stats_err, std::fabs(got.beta_ - want.beta_) / want.beta_);
        xg.AddMat(-1.0, xw);  // relative to the largest entry of the reference's transform (the offsets are O(10) here)
        fmllr_xform_err = std::max(fmllr_xform_err, (double)std::max(xg.Max(), -xg.Min()) / std::max(xw.Max(), -xw.Min()));
      }
    }
    // ---- allow_downsample: a 16 kHz rendering of a test wave through ComputeFeatures of both sides (feature-common-inl.h:29-55) ----
    double down_err = 0.0;
    int down_refused = 0;
    if (gpu) {
      MfccOptions mo = MfccOpts();
      mo.frame_opts.allow_downsample = true;
      Mfcc ref_mfcc(mo);
      vbgpu::GpuMfcc g_mfcc(mo);
      Vector<BaseFloat> w16(24000);
      for (int32 i = 0; i < w16.Dim(); i++)
        w16(i) = 3000.0f * std::sin(0.05f * i) + 1500.0f * std::sin(0.31f * i + 1.0f) + 200.0f * (float)((i * 7919) % 257 - 128) / 128.0f;
      Matrix<BaseFloat> want, got;
      ref_mfcc.ComputeFeatures(w16, 2.0f * kFs, 1.0f, &want);
      g_mfcc.ComputeFeatures(w16, 2.0f * kFs, 1.0f, &got);
      if (want.NumRows() != got.NumRows() || want.NumCols() != got.NumCols()) {
        down_err = 1.0;
      } else {
        for (int32 c = 0; c < want.NumCols(); c++) {
          double scale = 0.0;
          for (int32 t = 0; t < want.NumRows(); t++) scale += (double)want(t, c) * want(t, c);
          scale = std::max(1.0, std::sqrt(scale / std::max(1, want.NumRows())));
          for (int32 t = 0; t < want.NumRows(); t++)
            down_err = std::max(down_err, std::fabs((double)got(t, c) - want(t, c)) / std::max((double)std::fabs(want(t, c)), scale));
        }
      }
      // without the switch both sides refuse the wave
      vbgpu::GpuMfcc strict(MfccOpts());
      try {
        strict.ComputeFeatures(w16, 2.0f * kFs, 1.0f, &got);
      } catch (const std::exception &) { down_refused = 1; }
    }
    printf("{\"mode\": \"%s\", \"train_utts\": %d, \"test_utts\": %d, \"frames\": %lld, \"pdfs\": %d, \"gaussians\": %d, "
           "\"wer_reference\": %.4f",
           gpu ? "gpu" : "cpu", n_train, n_test, (long long)n_frames, am.NumPdfs(), am.NumGauss(),
           (double)errs / n_ref_words);
    if (gpu)
      printf(", \"wer_gpu\": %.4f, \"transcripts_identical\": %d, \"decode_alignments_identical\": %d, "
             "\"forced_alignments_identical\": %d, \"max_feat_rel_err\": %.3e, \"max_loglike_abs_err\": %.3e, "
             "\"max_abs_loglike\": %.1f, \"max_stats_rel_err\": %.3e, \"max_acc_loglike_rel_err\": %.3e",
             (double)gpu_errs / n_ref_words, same_words, same_ali, same_forced, feat_err / std::max(feat_scale, 1e-30),
             ll_err, ll_mag, stats_err, acc_like_err);
    if (gpu)
      printf(", \"batch_forced_alignments_identical\": %d, \"fmllr_stats_rel_err\": %.3e, \"fmllr_xform_rel_err\": %.3e, "
             "\"rescored_lattice_arcs\": %lld, \"rescored_arc_abs_err\": %.3e, \"rescored_best_paths_identical\": %d",
             batch_forced_same, fmllr_stats_err, fmllr_xform_err, rescored_arcs, rescore_err, rescored_same);
    if (gpu)
      printf(", \"subset_forced_alignments_identical\": %d, \"subset_floats\": %lld, \"dense_floats\": %lld, "
             "\"gather_arcs\": %lld, \"gather_arc_abs_err\": %.3e, \"gather_best_paths_identical\": %d",
             subset_forced_same, subset_floats, dense_floats, gather_arcs, gather_err, gather_same);
    if (gpu)
      printf(", \"pitch_frames\": %lld, \"pitch_frames_identical\": %lld, \"pitch_max_rel_err\": %.3e, "
             "\"pitch_nccf_abs_err\": %.3e, \"process_pitch_abs_err\": %.3e",
             pitch_frames, pitch_same, pitch_rel_err, pitch_nccf_err, process_err);
    if (gpu) printf(", \"downsample_mfcc_rel_err\": %.3e, \"downsample_refused_without_switch\": %d", down_err, down_refused);
    printf("}\n");
    delete gam;
    delete gfp;
    delete gmfcc;
    delete ctx_dep;
    return 0;
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_yesno: %s\n", e.what());
    return 1;
  }
}
