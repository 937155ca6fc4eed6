"""Helpers for the -m gpu parity tests: the oracle chain for a whole batch."""
import numpy as np

from oracle import pyoracle as po
from tests.common import copy_opts


def to_orc_opts(o):
    return copy_opts(o, po.MfccOpts)


def oracle_mfcc_batch(chk, opts, pcm, so, vtln=None):
    oo = to_orc_opts(opts)
    outs = []
    for u in range(len(so) - 1):
        w = np.asarray(pcm[so[u]:so[u + 1]], np.float32)
        outs.append(chk.mfcc(oo, w, 1.0 if vtln is None else float(vtln[u])))
    fo = np.zeros(len(so), np.int64)
    fo[1:] = np.cumsum([len(x) for x in outs])
    return (np.concatenate(outs) if outs else np.zeros((0, opts.num_ceps), np.float32)), fo


def oracle_stats(chk, mfcc, fo, u2s, n_spk):
    D = mfcc.shape[1]
    st = np.zeros((n_spk, 2, D + 1))
    for u in range(len(fo) - 1):
        s = u if u2s is None else int(u2s[u])
        if fo[u + 1] > fo[u]:
            st[s] = chk.cmvn_acc(mfcc[fo[u]:fo[u + 1]], st[s].copy())
    return st


def oracle_feats(chk, mfcc, fo, u2s, stats, fopts, lda=None, fmllr=None):
    outs = []
    for u in range(len(fo) - 1):
        x = mfcc[fo[u]:fo[u + 1]]
        if len(x) == 0:
            continue
        s = u if u2s is None else int(u2s[u])
        if fopts.norm_means or fopts.norm_vars:
            x = chk.cmvn_apply(stats[s], x, bool(fopts.norm_vars))
        if fopts.mode == 0:
            y = chk.deltas(x, fopts.delta_order, fopts.delta_window)
        else:
            y = chk.transform(chk.splice(x, fopts.splice_left, fopts.splice_right), lda)
        if fmllr is not None:
            y = chk.transform(y, fmllr[s])
        outs.append(y)
    return np.concatenate(outs)
