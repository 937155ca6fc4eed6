"""ctypes front end for the CPU checkers — TEST INFRASTRUCTURE ONLY.

`load("orc")` -> oracle/liboracle.so  (plain-C restatement, oracle/oracle.c)
`load("ref")` -> oracle/_ref/libvbref.so (the reference's own Kaldi code, oracle/ref_driver.cc)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this
module; nothing under voicebridge_b200/ does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class MfccOpts(C.Structure):
    """Layout shared by orc_mfcc_opts (oracle/oracle.h) and vbgpu_mfcc_opts (include/vbgpu.h)."""
    _fields_ = [
        ("samp_freq", C.c_float), ("frame_shift_ms", C.c_float), ("frame_length_ms", C.c_float),
        ("dither", C.c_float), ("preemph_coeff", C.c_float), ("remove_dc_offset", C.c_int32),
        ("window_type", C.c_int32), ("round_to_power_of_two", C.c_int32), ("blackman_coeff", C.c_float),
        ("snip_edges", C.c_int32), ("num_bins", C.c_int32), ("low_freq", C.c_float), ("high_freq", C.c_float),
        ("vtln_low", C.c_float), ("vtln_high", C.c_float), ("htk_mode", C.c_int32), ("num_ceps", C.c_int32),
        ("use_energy", C.c_int32), ("energy_floor", C.c_float), ("raw_energy", C.c_int32),
        ("cepstral_lifter", C.c_float), ("htk_compat", C.c_int32),
    ]


def default_opts(**kw):
    o = MfccOpts(16000.0, 10.0, 25.0, 1.0, 0.97, 1, 0, 1, 0.42, 1, 23, 20.0, 0.0, 100.0, -500.0, 0, 13, 1, 0.0, 1,
                 22.0, 0)
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


class PitchOpts(C.Structure):
    """Layout shared by orc_pitch_opts (oracle/oracle.h) and vbgpu_pitch_opts (include/vbgpu.h)."""
    _fields_ = [
        ("samp_freq", C.c_float), ("frame_shift_ms", C.c_float), ("frame_length_ms", C.c_float),
        ("preemph_coeff", C.c_float), ("min_f0", C.c_float), ("max_f0", C.c_float), ("soft_min_f0", C.c_float),
        ("penalty_factor", C.c_float), ("lowpass_cutoff", C.c_float), ("resample_freq", C.c_float),
        ("delta_pitch", C.c_float), ("nccf_ballast", C.c_float), ("lowpass_filter_width", C.c_int32),
        ("upsample_filter_width", C.c_int32), ("recompute_frame", C.c_int32), ("snip_edges", C.c_int32),
    ]


class ProcessPitchOpts(C.Structure):
    """Layout shared by orc_process_pitch_opts and vbgpu_process_pitch_opts."""
    _fields_ = [
        ("pitch_scale", C.c_float), ("pov_scale", C.c_float), ("pov_offset", C.c_float),
        ("delta_pitch_scale", C.c_float), ("delta_pitch_noise_stddev", C.c_float),
        ("normalization_left_context", C.c_int32), ("normalization_right_context", C.c_int32),
        ("delta_window", C.c_int32), ("delay", C.c_int32), ("add_pov_feature", C.c_int32),
        ("add_normalized_log_pitch", C.c_int32), ("add_delta_pitch", C.c_int32), ("add_raw_log_pitch", C.c_int32),
    ]


def _set(o, kw):
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


def default_pitch_opts(**kw):
    """PitchExtractionOptions defaults (feat/pitch-functions.h:103-123)."""
    return _set(PitchOpts(16000.0, 10.0, 25.0, 0.0, 50.0, 400.0, 10.0, 0.1, 1000.0, 4000.0, 0.005, 7000.0, 1, 5, 500, 1),
                kw)


def default_process_pitch_opts(**kw):
    """ProcessPitchOptions defaults (feat/pitch-functions.h:241-255), except delta_pitch_noise_stddev = 0: the
    reference draws that noise from rand(), which no other implementation can reproduce."""
    return _set(ProcessPitchOpts(2.0, 2.0, 0.0, 10.0, 0.0, 75, 75, 2, 0, 1, 1, 1, 0), kw)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def stride_of(cols):
    """Kaldi Matrix<float> stride: cols rounded up to 4 floats (kaldi-matrix.cc:797-808)."""
    return (cols + 3) // 4 * 4


class Lib:
    def __init__(self, kind):
        self.kind = kind
        if kind == "orc":
            path = os.path.join(HERE, "liboracle.so")
        elif kind == "ref":
            os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
            path = os.path.join(HERE, "_ref", "libvbref.so")
        else:
            raise ValueError(kind)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        if kind == "ref":  # the wheel's OpenBLAS needs its private libgfortran/libquadmath: preload them
            import glob
            for pat in ("libquadmath-*", "libgfortran-*", "libopenblasp-*"):
                for so in sorted(glob.glob(os.path.join(HERE, "_ref", pat))):
                    C.CDLL(so, mode=C.RTLD_GLOBAL)
        self.lib = C.CDLL(path)
        self.pfx = kind + "_"
        if kind == "ref":
            self.lib.ref_model_create.restype = C.c_void_p
            self.lib.ref_pcm_to_loglikes.restype = C.c_int64

    def fn(self, name):
        return getattr(self.lib, self.pfx + name)

    # ---- front end -------------------------------------------------------------------------------
    def num_frames(self, n, opts):
        if self.kind == "orc":
            return self.lib.orc_num_frames(C.c_int64(n), C.byref(opts))
        return self.lib.ref_num_frames(C.c_int64(n), C.byref(opts))

    def window_table(self, opts):
        L = int(opts.samp_freq * 0.001 * opts.frame_length_ms)
        w = np.zeros(L, np.float32)
        rc = self.fn("window_table")(C.byref(opts), _p(w, C.c_float))
        assert rc == 0
        return w

    def mel_banks(self, opts, vtln_warp=1.0):
        L = int(opts.samp_freq * 0.001 * opts.frame_length_ms)
        npad = 1
        while npad < L:
            npad *= 2
        nfft = npad // 2
        B = opts.num_bins
        offs = np.zeros(B, np.int32)
        lens = np.zeros(B, np.int32)
        w = np.zeros((B, nfft), np.float32)
        rc = self.fn("mel_banks")(C.byref(opts), C.c_float(vtln_warp), _p(offs, C.c_int32), _p(lens, C.c_int32),
                                  _p(w, C.c_float))
        if rc != 0:
            raise RuntimeError("mel_banks rc=%d" % rc)
        return offs, lens, w

    def mfcc(self, opts, wave, vtln_warp=1.0):
        wave = _f32(wave)
        T = self.num_frames(len(wave), opts)
        st = stride_of(opts.num_ceps)
        out = np.zeros((max(T, 1), st), np.float32)
        rc = self.fn("mfcc_compute")(C.byref(opts), _p(wave, C.c_float), C.c_int64(len(wave)), C.c_float(vtln_warp),
                                     _p(out, C.c_float), C.c_int32(st))
        if rc < 0:
            raise RuntimeError("mfcc_compute rc=%d" % rc)
        assert rc == T, (rc, T)
        return out[:T, :opts.num_ceps].copy()

    def fbank(self, opts, wave, vtln_warp=1.0, use_log_fbank=1, use_power=1):
        """OfflineFeatureTpl<FbankComputer>: [T, num_bins (+1 with use_energy)]."""
        wave = _f32(wave)
        T = self.num_frames(len(wave), opts)
        dim = opts.num_bins + (1 if opts.use_energy else 0)
        st = stride_of(dim)
        out = np.zeros((max(T, 1), st), np.float32)
        rc = self.fn("fbank_compute")(C.byref(opts), C.c_int32(use_log_fbank), C.c_int32(use_power), _p(wave, C.c_float),
                                      C.c_int64(len(wave)), C.c_float(vtln_warp), _p(out, C.c_float), C.c_int32(st))
        if rc < 0:
            raise RuntimeError("fbank_compute rc=%d" % rc)
        assert rc == T, (rc, T)
        return out[:T, :dim].copy()

    def plp(self, opts, wave, vtln_warp=1.0, lpc_order=12, compress_factor=0.33333, cepstral_scale=1.0):
        """OfflineFeatureTpl<PlpComputer>: [T, num_ceps]."""
        wave = _f32(wave)
        T = self.num_frames(len(wave), opts)
        st = stride_of(opts.num_ceps)
        out = np.zeros((max(T, 1), st), np.float32)
        rc = self.fn("plp_compute")(C.byref(opts), C.c_int32(lpc_order), C.c_float(compress_factor),
                                    C.c_float(cepstral_scale), _p(wave, C.c_float), C.c_int64(len(wave)),
                                    C.c_float(vtln_warp), _p(out, C.c_float), C.c_int32(st))
        if rc < 0:
            raise RuntimeError("plp_compute rc=%d" % rc)
        assert rc == T, (rc, T)
        return out[:T, :opts.num_ceps].copy()

    def downsample_waveform(self, orig_freq, new_freq, wave):
        """DownsampleWaveForm (feat/resample.cc:368-376)."""
        wave = _f32(wave)
        f = self.fn("downsample_waveform")
        f.restype = C.c_int64
        n = f(C.c_float(orig_freq), C.c_float(new_freq), _p(wave, C.c_float), C.c_int64(len(wave)), None)
        if n < 0:
            raise RuntimeError("downsample_waveform rc=%d" % n)
        out = np.zeros(max(int(n), 1), np.float32)
        f(C.c_float(orig_freq), C.c_float(new_freq), _p(wave, C.c_float), C.c_int64(len(wave)), _p(out, C.c_float))
        return out[:n].copy()

    def pitch(self, opts, wave):
        """ComputeKaldiPitch: [T, 2] = (NCCF, pitch Hz)."""
        wave = _f32(wave)
        cap = len(wave) // max(1, int(opts.samp_freq * opts.frame_shift_ms * 0.001)) + 8
        out = np.zeros((cap, 2), np.float32)
        rc = self.fn("pitch_compute")(C.byref(opts), _p(wave, C.c_float), C.c_int64(len(wave)), _p(out, C.c_float),
                                      C.c_int32(2))
        if rc < 0:
            raise RuntimeError("pitch_compute rc=%d" % rc)
        assert rc <= cap
        return out[:rc].copy()

    def process_pitch(self, opts, raw):
        """ProcessPitch: [T, 2] -> [T + delay, dim]."""
        raw = _f32(raw)
        T = raw.shape[0]
        dim = sum(1 for k in ("add_pov_feature", "add_normalized_log_pitch", "add_delta_pitch", "add_raw_log_pitch")
                  if getattr(opts, k))
        out = np.zeros((T + opts.delay + 1, max(dim, 1)), np.float32)
        rc = self.fn("process_pitch")(C.byref(opts), _p(raw, C.c_float), C.c_int32(T), C.c_int32(2), _p(out, C.c_float),
                                      C.c_int32(max(dim, 1)))
        if rc < 0:
            raise RuntimeError("process_pitch rc=%d" % rc)
        return out[:rc].copy()

    # ---- feature post-processing -----------------------------------------------------------------------
    def cmvn_acc(self, feats, stats=None):
        feats = _f32(feats)
        T, D = feats.shape
        if stats is None:
            stats = np.zeros((2, D + 1), np.float64)
        self.fn("cmvn_acc")(_p(feats, C.c_float), C.c_int32(T), C.c_int32(D), C.c_int32(D), _p(stats, C.c_double))
        return stats

    def cmvn_apply(self, stats, feats, norm_vars=False):
        out = _f32(feats).copy()
        T, D = out.shape
        stats = np.ascontiguousarray(stats, np.float64)
        rc = self.fn("cmvn_apply")(_p(stats, C.c_double), C.c_int32(D), C.c_int32(int(norm_vars)), _p(out, C.c_float),
                                   C.c_int32(T), C.c_int32(D))
        if rc != 0:
            raise RuntimeError("cmvn_apply rc=%d" % rc)
        return out

    def deltas(self, feats, order=2, window=2):
        feats = _f32(feats)
        T, D = feats.shape
        out = np.zeros((T, D * (order + 1)), np.float32)
        self.fn("deltas")(C.c_int32(order), C.c_int32(window), _p(feats, C.c_float), C.c_int32(T), C.c_int32(D),
                          C.c_int32(D), _p(out, C.c_float), C.c_int32(out.shape[1]))
        return out

    def splice(self, feats, left, right):
        feats = _f32(feats)
        T, D = feats.shape
        out = np.zeros((T, D * (left + right + 1)), np.float32)
        self.fn("splice")(_p(feats, C.c_float), C.c_int32(T), C.c_int32(D), C.c_int32(D), C.c_int32(left),
                          C.c_int32(right), _p(out, C.c_float), C.c_int32(out.shape[1]))
        return out

    def transform(self, feats, mat):
        feats = _f32(feats)
        mat = _f32(mat)
        T, D = feats.shape
        out = np.zeros((T, mat.shape[0]), np.float32)
        rc = self.fn("transform")(_p(feats, C.c_float), C.c_int32(T), C.c_int32(D), C.c_int32(D), _p(mat, C.c_float),
                                  C.c_int32(mat.shape[0]), C.c_int32(mat.shape[1]), _p(out, C.c_float),
                                  C.c_int32(out.shape[1]))
        if rc != 0:
            raise RuntimeError("transform rc=%d" % rc)
        return out

    # ---- model -----------------------------------------------------------------------------------------
    def gconsts(self, weights, miv, iv):
        """orc only: DiagGmm::ComputeGconsts for one flat block of Gaussians."""
        assert self.kind == "orc"
        weights, miv, iv = _f32(weights), _f32(miv), _f32(iv)
        M, D = miv.shape
        g = np.zeros(M, np.float32)
        rc = self.lib.orc_gconsts(C.c_int32(M), C.c_int32(D), _p(weights, C.c_float), _p(miv, C.c_float),
                                  _p(iv, C.c_float), _p(g, C.c_float))
        if rc < 0:
            raise RuntimeError("gconsts rc=%d" % rc)
        return g

    def model_params(self, pdf_offsets, weights, means, inv_vars):
        """(gconsts, means_invvars, inv_vars) as the implementation itself holds them."""
        pdf_offsets = np.ascontiguousarray(pdf_offsets, np.int32)
        weights, means, inv_vars = _f32(weights), _f32(means), _f32(inv_vars)
        N, D = means.shape
        if self.kind == "orc":
            miv = (means * inv_vars).astype(np.float32)  # SetInvVarsAndMeans: MulElements in float
            return self.gconsts(weights, miv, inv_vars), miv, inv_vars
        h = self.ref_model(pdf_offsets, weights, means, inv_vars)
        g = np.zeros(N, np.float32)
        miv = np.zeros((N, D), np.float32)
        iv = np.zeros((N, D), np.float32)
        self.lib.ref_model_get(C.c_void_p(h), _p(g, C.c_float), _p(miv, C.c_float), _p(iv, C.c_float))
        self.lib.ref_model_destroy(C.c_void_p(h))
        return g, miv, iv

    def ref_model(self, pdf_offsets, weights, means, inv_vars):
        assert self.kind == "ref"
        pdf_offsets = np.ascontiguousarray(pdf_offsets, np.int32)
        weights, means, inv_vars = _f32(weights), _f32(means), _f32(inv_vars)
        h = self.lib.ref_model_create(C.c_int32(len(pdf_offsets) - 1), C.c_int32(means.shape[1]),
                                      _p(pdf_offsets, C.c_int32), _p(weights, C.c_float), _p(means, C.c_float),
                                      _p(inv_vars, C.c_float))
        if not h:
            raise RuntimeError("ref_model_create failed")
        return h

    # ---- scoring / accumulation (orc: flattened arrays; ref: model handle) ----------------------------
    def gmm_loglikes(self, model, feats, prune=-1.0):
        """model: object with pdf_offsets, gconsts, miv, iv (+ weights, means for ref)."""
        feats = _f32(feats)
        T, D = feats.shape
        P = len(model.pdf_offsets) - 1
        out = np.zeros((T, P), np.float32)
        if self.kind == "orc":
            rc = self.lib.orc_gmm_loglikes(C.c_int32(P), C.c_int32(D), _p(model.pdf_offsets, C.c_int32),
                                           _p(model.gconsts, C.c_float), _p(model.miv, C.c_float),
                                           _p(model.iv, C.c_float), _p(feats, C.c_float), C.c_int32(T), C.c_int32(D),
                                           C.c_float(prune), _p(out, C.c_float), C.c_int32(P))
        else:
            h = self.ref_model(model.pdf_offsets, model.weights, model.means, model.iv)
            rc = self.lib.ref_gmm_loglikes(C.c_void_p(h), _p(feats, C.c_float), C.c_int32(T), C.c_int32(D),
                                           C.c_float(prune), _p(out, C.c_float), C.c_int32(P))
            self.lib.ref_model_destroy(C.c_void_p(h))
        return rc, out

    def gmm_loglikes_matrix(self, model, feats):
        assert self.kind == "ref"
        feats = _f32(feats)
        T, D = feats.shape
        P = len(model.pdf_offsets) - 1
        out = np.zeros((T, P), np.float32)
        h = self.ref_model(model.pdf_offsets, model.weights, model.means, model.iv)
        rc = self.lib.ref_gmm_loglikes_matrix(C.c_void_p(h), _p(feats, C.c_float), C.c_int32(T), C.c_int32(D),
                                              _p(out, C.c_float), C.c_int32(P))
        self.lib.ref_model_destroy(C.c_void_p(h))
        return rc, out

    def acc_ali(self, model, feats, pdf_ids, weights=None, feats2=None):
        feats = _f32(feats)
        T, D = feats.shape
        N = len(model.gconsts)
        pdf_ids = np.ascontiguousarray(pdf_ids, np.int32)
        occ = np.zeros(N, np.float64)
        mean = np.zeros((N, D), np.float64)
        var = np.zeros((N, D), np.float64)
        tl = C.c_double(0.0)
        tf = C.c_double(0.0)
        wp = _p(_f32(weights), C.c_float) if weights is not None else None
        if weights is not None:
            weights = _f32(weights)
            wp = _p(weights, C.c_float)
        tail = (C.c_int32(T), C.c_int32(D), _p(pdf_ids, C.c_int32), wp, _p(occ, C.c_double), _p(mean, C.c_double),
                _p(var, C.c_double), C.byref(tl), C.byref(tf))
        if feats2 is not None:
            feats2 = _f32(feats2)
        if self.kind == "orc":
            head = (C.c_int32(len(model.pdf_offsets) - 1), C.c_int32(D), _p(model.pdf_offsets, C.c_int32),
                    _p(model.gconsts, C.c_float), _p(model.miv, C.c_float), _p(model.iv, C.c_float))
            if feats2 is None:
                rc = self.lib.orc_acc_ali(*head, _p(feats, C.c_float), *tail)
            else:
                rc = self.lib.orc_acc_ali_twofeats(*head, _p(feats, C.c_float), _p(feats2, C.c_float), *tail)
        else:
            h = self.ref_model(model.pdf_offsets, model.weights, model.means, model.iv)
            if feats2 is None:
                rc = self.lib.ref_acc_ali(C.c_void_p(h), _p(feats, C.c_float), *tail)
            else:
                rc = self.lib.ref_acc_ali_twofeats(C.c_void_p(h), _p(feats, C.c_float), _p(feats2, C.c_float), *tail)
            self.lib.ref_model_destroy(C.c_void_p(h))
        return rc, occ, mean, var, tl.value, tf.value


    def fmllr_acc(self, model, feats, pdf_ids, weights=None):
        """FmllrDiagGmmAccs over an alignment: returns rc, beta, K[D, D+1], G[D, (D+1)(D+2)/2] (SpMatrix packing),
        tot_like."""
        feats = _f32(feats)
        T, D = feats.shape
        pdf_ids = np.ascontiguousarray(pdf_ids, np.int32)
        beta, tl = C.c_double(0.0), C.c_double(0.0)
        K = np.zeros((D, D + 1), np.float64)
        G = np.zeros((D, (D + 1) * (D + 2) // 2), np.float64)
        wp = None
        if weights is not None:
            weights = _f32(weights)
            wp = _p(weights, C.c_float)
        tail = (_p(feats, C.c_float), C.c_int32(T), C.c_int32(D), _p(pdf_ids, C.c_int32), wp, C.byref(beta),
                _p(K, C.c_double), _p(G, C.c_double), C.byref(tl))
        if self.kind == "orc":
            rc = self.lib.orc_fmllr_acc(C.c_int32(len(model.pdf_offsets) - 1), C.c_int32(D),
                                        _p(model.pdf_offsets, C.c_int32), _p(model.gconsts, C.c_float),
                                        _p(model.miv, C.c_float), _p(model.iv, C.c_float), *tail)
        else:
            h = self.ref_model(model.pdf_offsets, model.weights, model.means, model.iv)
            rc = self.lib.ref_fmllr_acc(C.c_void_p(h), *tail)
            self.lib.ref_model_destroy(C.c_void_p(h))
        return rc, beta.value, K, G, tl.value

    def component_posteriors(self, model, feats, pdf_ids, weights=None):
        """(rc, post, offsets, loglikes): frame t's Gaussian posteriors at post[offsets[t]:offsets[t+1]]."""
        feats = _f32(feats)
        T, D = feats.shape
        pdf_ids = np.ascontiguousarray(pdf_ids, np.int32)
        offs = np.zeros(T + 1, np.int64)
        offs[1:] = np.cumsum(np.diff(model.pdf_offsets)[pdf_ids])
        post = np.zeros(int(offs[-1]), np.float32)
        ll = np.zeros(T, np.float32)
        wp = None
        if weights is not None:
            weights = _f32(weights)
            wp = _p(weights, C.c_float)
        tail = (_p(feats, C.c_float), C.c_int32(T), C.c_int32(D), _p(pdf_ids, C.c_int32), wp, _p(post, C.c_float),
                _p(ll, C.c_float))
        if self.kind == "orc":
            rc = self.lib.orc_component_posteriors(C.c_int32(len(model.pdf_offsets) - 1), C.c_int32(D),
                                                   _p(model.pdf_offsets, C.c_int32), _p(model.gconsts, C.c_float),
                                                   _p(model.miv, C.c_float), _p(model.iv, C.c_float), *tail)
        else:
            h = self.ref_model(model.pdf_offsets, model.weights, model.means, model.iv)
            rc = self.lib.ref_component_posteriors(C.c_void_p(h), *tail)
            self.lib.ref_model_destroy(C.c_void_p(h))
        return rc, post, offs, ll

    def mllt_acc(self, model, feats, pdf_ids, weights=None):
        """MlltAccs (rand_prune = 0) over an alignment: rc, beta, G[D, D(D+1)/2] (SpMatrix packing), tot_like."""
        feats = _f32(feats)
        T, D = feats.shape
        pdf_ids = np.ascontiguousarray(pdf_ids, np.int32)
        beta, tl = C.c_double(0.0), C.c_double(0.0)
        G = np.zeros((D, D * (D + 1) // 2), np.float64)
        wp = None
        if weights is not None:
            weights = _f32(weights)
            wp = _p(weights, C.c_float)
        tail = (_p(feats, C.c_float), C.c_int32(T), C.c_int32(D), _p(pdf_ids, C.c_int32), wp, C.byref(beta),
                _p(G, C.c_double), C.byref(tl))
        if self.kind == "orc":
            rc = self.lib.orc_mllt_acc(C.c_int32(len(model.pdf_offsets) - 1), C.c_int32(D),
                                       _p(model.pdf_offsets, C.c_int32), _p(model.gconsts, C.c_float),
                                       _p(model.miv, C.c_float), _p(model.iv, C.c_float), *tail)
        else:
            h = self.ref_model(model.pdf_offsets, model.weights, model.means, model.iv)
            rc = self.lib.ref_mllt_acc(C.c_void_p(h), *tail)
            self.lib.ref_model_destroy(C.c_void_p(h))
        return rc, beta.value, G, tl.value

    def mllt_update(self, beta, G):
        """The reference's MlltAccs::Update on given statistics, from the unit matrix: M[D, D], objf_impr, count."""
        assert self.kind == "ref"
        D = G.shape[0]
        G = np.ascontiguousarray(G, np.float64)
        M = np.zeros((D, D), np.float32)
        impr, cnt = C.c_float(0), C.c_float(0)
        rc = self.lib.ref_mllt_update(C.c_int32(D), C.c_double(beta), _p(G, C.c_double), _p(M, C.c_float), C.byref(impr),
                                      C.byref(cnt))
        return rc, M, impr.value, cnt.value

    def fmllr_update(self, beta, K, G):
        """The reference's FmllrDiagGmmAccs::Update (default options) on given statistics: xform, objf_impr, count."""
        assert self.kind == "ref"
        D = K.shape[0]
        K = np.ascontiguousarray(K, np.float64)
        G = np.ascontiguousarray(G, np.float64)
        x = np.zeros((D, D + 1), np.float32)
        impr, cnt = C.c_float(0), C.c_float(0)
        rc = self.lib.ref_fmllr_update(C.c_int32(D), C.c_double(beta), _p(K, C.c_double), _p(G, C.c_double),
                                       _p(x, C.c_float), C.byref(impr), C.byref(cnt))
        return rc, x, impr.value, cnt.value


    # ---- wire formats (reference side only): the reference's own writers / readers on byte strings -----------------
    def io_write_matrix(self, m, kind):
        """kind: 0 FM, 1 DM, 2 CM (automatic method), 3 CM2, 4 CM3; with the table holders' \\0B marker."""
        assert self.kind == "ref"
        m = _f32(m)
        args = (_p(m, C.c_float), C.c_int32(m.shape[0]), C.c_int32(m.shape[1]), C.c_int32(m.shape[1]), C.c_int32(kind))
        self.lib.ref_io_write_matrix.restype = C.c_int64
        n = self.lib.ref_io_write_matrix(*args, None, C.c_int64(0))
        assert n > 0, n
        buf = C.create_string_buffer(n)
        assert self.lib.ref_io_write_matrix(*args, buf, C.c_int64(n)) == n
        return buf.raw

    def io_read_matrix(self, b):
        assert self.kind == "ref"
        r, c = C.c_int32(0), C.c_int32(0)
        rc = self.lib.ref_io_read_matrix(b, C.c_int64(len(b)), C.byref(r), C.byref(c), None, C.c_int64(0))
        assert rc == 0, rc
        out = np.zeros((r.value, c.value), np.float32)
        rc = self.lib.ref_io_read_matrix(b, C.c_int64(len(b)), C.byref(r), C.byref(c), _p(out, C.c_float),
                                         C.c_int64(out.size))
        assert rc == 0, rc
        return out

    def io_write_int32_vector(self, v):
        assert self.kind == "ref"
        v = np.ascontiguousarray(v, np.int32)
        self.lib.ref_io_write_int32_vector.restype = C.c_int64
        n = self.lib.ref_io_write_int32_vector(_p(v, C.c_int32), C.c_int32(len(v)), None, C.c_int64(0))
        buf = C.create_string_buffer(n)
        assert self.lib.ref_io_write_int32_vector(_p(v, C.c_int32), C.c_int32(len(v)), buf, C.c_int64(n)) == n
        return buf.raw

    def io_read_int32_vector(self, b, cap=1 << 20):
        assert self.kind == "ref"
        out = np.zeros(cap, np.int32)
        n = self.lib.ref_io_read_int32_vector(b, C.c_int64(len(b)), _p(out, C.c_int32), C.c_int32(cap))
        assert n >= 0, n
        return out[:n].copy()

    def io_write_mdl(self, model, n_phones):
        """Bytes of a model file (\\0B TransitionModel AmDiagGmm) for a monophone system over `model`'s 3*n_phones pdfs."""
        assert self.kind == "ref"
        h = self.ref_model(model.pdf_offsets, model.weights, model.means, model.iv)
        self.lib.ref_io_write_mdl.restype = C.c_int64
        n = self.lib.ref_io_write_mdl(C.c_void_p(h), C.c_int32(n_phones), None, C.c_int64(0))
        assert n > 0, n
        buf = C.create_string_buffer(n)
        assert self.lib.ref_io_write_mdl(C.c_void_p(h), C.c_int32(n_phones), buf, C.c_int64(n)) == n
        self.lib.ref_model_destroy(C.c_void_p(h))
        return buf.raw

    def io_read_mdl(self, b, dim, tid_cap=1 << 16, gauss_cap=1 << 16):
        assert self.kind == "ref"
        nt, P, N = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        t2p = np.zeros(tid_cap, np.int32)
        lp = np.zeros(tid_cap, np.float32)
        gc, w = np.zeros(gauss_cap, np.float32), np.zeros(gauss_cap, np.float32)
        miv, iv = np.zeros((gauss_cap, dim), np.float32), np.zeros((gauss_cap, dim), np.float32)
        rc = self.lib.ref_io_read_mdl(b, C.c_int64(len(b)), C.byref(nt), _p(t2p, C.c_int32), C.c_int32(tid_cap),
                                      _p(lp, C.c_float), C.byref(P), C.byref(N), _p(gc, C.c_float), _p(miv, C.c_float),
                                      _p(iv, C.c_float), _p(w, C.c_float), C.c_int64(gauss_cap))
        assert rc == 0, rc
        n, t = N.value, nt.value
        return dict(num_pdfs=P.value, tid2pdf=t2p[:t + 1].copy(), trans_log_probs=lp[:t + 1].copy(), gconsts=gc[:n].copy(),
                    weights=w[:n].copy(), means_invvars=miv[:n].copy(), inv_vars=iv[:n].copy())

    def io_read_acc(self, model, b, n_trans):
        assert self.kind == "ref"
        N, D = len(model.gconsts), model.means.shape[1]
        h = self.ref_model(model.pdf_offsets, model.weights, model.means, model.iv)
        tr = np.zeros(max(n_trans, 1), np.float64)
        occ, mean, var = np.zeros(N), np.zeros((N, D)), np.zeros((N, D))
        tl, tf = C.c_double(0.0), C.c_double(0.0)
        rc = self.lib.ref_io_read_acc(C.c_void_p(h), b, C.c_int64(len(b)), C.c_int32(n_trans), _p(tr, C.c_double),
                                      _p(occ, C.c_double), _p(mean, C.c_double), _p(var, C.c_double), C.byref(tl),
                                      C.byref(tf))
        self.lib.ref_model_destroy(C.c_void_p(h))
        assert rc == 0, rc
        return tr[:n_trans], occ, mean, var, tl.value, tf.value


_cache = {}


def load(kind):
    if kind not in _cache:
        _cache[kind] = Lib(kind)
    return _cache[kind]


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libvbref.so"))
