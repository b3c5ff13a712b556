"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d): seeded audio and the
random-init TDNN, written in the reference's own file formats.

Host-side preparation only (numpy); nothing here is on the hot path.
"""
import os

import numpy as np

from . import formats as F

SAMPLE_RATE = 16000
FRAME_LEN = 400
FRAME_SHIFT = 160
UTT_SAMPLES_10S = 160000
AUDIO_SEED = 20261018
MODEL_SEED = 1234


def num_frames(n_samples):
    """Snip-edges frame count, src/fbank.cc:35-42."""
    return 0 if n_samples < FRAME_LEN else 1 + (n_samples - FRAME_LEN) // FRAME_SHIFT


def synth_utterance(u, n_samples=UTT_SAMPLES_10S, seed=AUDIO_SEED):
    """Config-2 audio: round(3000 sin(2 pi f_u i / 16000) + 1000 N(0,1)) clipped to int16,
    f_u ~ U[80, 4000], numpy default_rng(seed + u)."""
    rng = np.random.default_rng(seed + u)
    f = rng.uniform(80.0, 4000.0)
    i = np.arange(n_samples, dtype=np.float64)
    x = 3000.0 * np.sin(2.0 * np.pi * f * i / SAMPLE_RATE) + 1000.0 * rng.standard_normal(n_samples)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def synth_batch(n_utts, n_samples=UTT_SAMPLES_10S, first=0, seed=AUDIO_SEED):
    """Concatenated PCM + sample offsets [n_utts+1] (the layout ce_gpu_forward takes)."""
    pcm = np.empty(n_utts * n_samples, np.int16)
    for u in range(n_utts):
        pcm[u * n_samples:(u + 1) * n_samples] = synth_utterance(first + u, n_samples, seed)
    offsets = np.arange(n_utts + 1, dtype=np.int64) * n_samples
    return pcm, offsets


# -- the TDNN of SURVEY.md section 8(d) -------------------------------------

TDNN_SPLICES = [[-2, -1, 0, 1, 2], [-1, 0, 1], [-1, 0, 1], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3]]


def tdnn_layers(feat_dim=40, hidden=1024, num_pdfs=3072, splices=None, seed=MODEL_SEED):
    """Layer list in convert_am.py order (tool/convert_am.py:272-281): Splice, Narrow,
    Linear, ReLU, BatchNorm per hidden layer; Linear + LogSoftmax at the end.
    W ~ N(0, 1/sqrt(K)), b ~ N(0, 0.1^2), BN scale ~ U[0.5,1.5], offset ~ N(0, 0.1^2)."""
    rng = np.random.default_rng(seed)
    splices = TDNN_SPLICES if splices is None else splices
    layers = []
    dim = feat_dim
    left = right = 0
    for idx in splices:
        k = dim * len(idx)
        nl, nr = -min(min(idx), 0), max(max(idx), 0)
        layers.append({"type": F.SPLICE, "indices": list(idx)})
        layers.append({"type": F.NARROW, "left": nl, "right": nr})
        layers.append({"type": F.LINEAR,
                       "W": (rng.standard_normal((k, hidden)) / np.sqrt(k)).astype(np.float32),
                       "b": (0.1 * rng.standard_normal(hidden)).astype(np.float32)})
        layers.append({"type": F.RELU})
        layers.append({"type": F.BATCHNORM,
                       "scale": rng.uniform(0.5, 1.5, hidden).astype(np.float32),
                       "offset": (0.1 * rng.standard_normal(hidden)).astype(np.float32)})
        dim = hidden
        left += nl
        right += nr
    layers.append({"type": F.LINEAR,
                   "W": (rng.standard_normal((dim, num_pdfs)) / np.sqrt(dim)).astype(np.float32),
                   "b": (0.1 * rng.standard_normal(num_pdfs)).astype(np.float32)})
    layers.append({"type": F.LOGSOFTMAX})
    z = rng.standard_normal(num_pdfs)
    prior = np.exp(z - z.max())
    prior = (prior / prior.sum()).astype(np.float32)
    return layers, left, right, prior


def flops_per_frame(layers):
    """Algorithmic FLOPs per output frame = 2 * sum(in*out) over Linear layers."""
    return 2 * sum(int(l["W"].shape[0]) * int(l["W"].shape[1]) for l in layers if l["type"] == F.LINEAR)


def write_model(dirname, name="tdnn", chunk_size=1 << 20, cmvn_stats=None, **kw):
    """Writes <name>.nnet/.prior/.tid2pdf/.conf (+ .cmvn) under dirname; returns paths + meta."""
    os.makedirs(dirname, exist_ok=True)
    layers, left, right, prior = tdnn_layers(**kw)
    p = {k: os.path.join(dirname, "%s.%s" % (name, k))
         for k in ("nnet", "prior", "tid2pdf", "conf", "cmvn")}
    F.write_nnet(p["nnet"], layers, left, right)
    F.write_vector(p["prior"], prior)
    num_pdfs = int(prior.size)
    F.write_vector(p["tid2pdf"], np.arange(num_pdfs, dtype=np.int32), dtype="<i4")
    extra = {}
    if cmvn_stats is not None:
        F.write_vector(p["cmvn"], cmvn_stats)
        extra["cmvn_stats"] = os.path.basename(p["cmvn"])
    F.write_am_config(p["conf"], p["nnet"], p["prior"], left, right, chunk_size, num_pdfs,
                      p["tid2pdf"], extra)
    p.update(left=left, right=right, num_pdfs=num_pdfs, flops_per_frame=flops_per_frame(layers))
    return p


def default_cmvn_stats(mel=40, count=36162480.0, mean=14.0):
    """Synthetic global CMVN stats (sum per dim, then count) in the layout of
    test/data/cmvn_stats.bin (41 floats for 40 bins)."""
    g = np.empty(mel + 1, np.float32)
    g[:mel] = np.float32(count * mean) * (1.0 + 0.01 * np.arange(mel, dtype=np.float32))
    g[mel] = count
    return g
