"""Multi-GPU and long-form sharding of the hot path.

Utterances are independent, so a batch is split into contiguous frame-balanced groups, one per
GPU (one process per GPU under torchrun), with NO collective on the data path; a long stream is
split into contiguous time shards whose halos are recomputed locally.  The plans come from the C
library (ce_gpu_partition / ce_gpu_time_shards) so that every rank derives the same one.
"""
import numpy as np

from . import api

FRAME_LEN, FRAME_SHIFT = 400, 160


def my_utterances(sample_offsets, rank, world):
    """(first_utt, end_utt) of this rank's group."""
    fo = api.frame_offsets(sample_offsets)
    pb = api.partition(fo, world)
    return int(pb[rank]), int(pb[rank + 1])


def forward_sharded(model, pcm, sample_offsets, rank, world, **kw):
    """This rank's part of a batch: returns (loglik, argmax, frame_offsets_local, (u0, u1))."""
    off = np.ascontiguousarray(sample_offsets, np.int64)
    u0, u1 = my_utterances(off, rank, world)
    local = off[u0:u1 + 1]
    ll, am, fo = model.forward(pcm, local, **kw)
    return ll, am, fo, (u0, u1)


def frame_to_sample_range(f0, f1):
    """Samples that frames [f0, f1) read: [160 f0, 160 (f1 - 1) + 400)."""
    if f1 <= f0:
        return 0, 0
    return FRAME_SHIFT * f0, FRAME_SHIFT * (f1 - 1) + FRAME_LEN


def forward_longform(models, pcm, want_loglik=True):
    """One long stream over len(models) time shards (one model handle per shard / GPU; the same
    handle may be repeated to run shards back to back on one GPU).  Rows are exact w.r.t. the
    whole-stream evaluation for the AM context; the CMVN running sums restart 600 frames before
    each shard, so they agree to fp32 rounding, not bit for bit."""
    n = len(models)
    m0 = models[0]
    total = int(api.frame_offsets([0, pcm.shape[0]])[-1])
    kb, ke, fb, fe = api.time_shards(total, n, m0.left_context, m0.right_context, 600)
    loglik = None                                  # rows as the handles write them (set_output)
    argmax = np.zeros(total, np.int32)
    for p, m in enumerate(models):
        if ke[p] <= kb[p]:
            continue
        s0, s1 = frame_to_sample_range(int(fb[p]), int(fe[p]))
        ll, am, _ = m.forward(pcm[s0:s1], want_loglik=want_loglik)
        a, b = int(kb[p] - fb[p]), int(ke[p] - fb[p])
        if want_loglik:
            if loglik is None:
                loglik = np.zeros((total,) + ll.shape[1:], ll.dtype)
            loglik[kb[p]:ke[p]] = ll[a:b]
        argmax[kb[p]:ke[p]] = am[a:b]
    return loglik, argmax
