"""Decoder feed with fewer bytes per frame: ce_gpu_model_set_output.
The selected rows must be the dense rows' values bit for bit -- Decoder::Process reads
frame_logp(tid2pdf[ilabel]) (src/decoder.cc:97-102), so a gathered row with a remapped tid2pdf is
the same computation -- and the top-k order is fully specified (value descending, pdf ascending)."""
import os

import numpy as np
import pytest

from catears_b200 import api, formats as F, synth

pytestmark = pytest.mark.gpu


def topk_reference(ll, k):
    """(values, pdfs) of the k largest per row, ties by ascending pdf."""
    order = np.lexsort((np.broadcast_to(np.arange(ll.shape[1]), ll.shape), -ll.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(ll, order, axis=1), order.astype(np.int32)


def write_tied_model(dirname, hidden, num_pdfs, tied, seed):
    """A TDNN whose output columns `tied` are copies of column tied[0] (weights, bias and prior),
    so those pdfs have equal log-likelihoods on every frame."""
    layers, left, right, prior = synth.tdnn_layers(hidden=hidden, num_pdfs=num_pdfs, seed=seed)
    out = [l for l in layers if l["type"] == F.LINEAR][-1]
    for j in tied[1:]:
        out["W"][:, j] = out["W"][:, tied[0]]
        out["b"][j] = out["b"][tied[0]]
        prior[j] = prior[tied[0]]
    p = {k: os.path.join(dirname, "tied.%s" % k) for k in ("nnet", "prior", "tid2pdf", "conf")}
    F.write_nnet(p["nnet"], layers, left, right)
    F.write_vector(p["prior"], prior)
    F.write_vector(p["tid2pdf"], np.arange(num_pdfs, dtype=np.int32), dtype="<i4")
    F.write_am_config(p["conf"], p["nnet"], p["prior"], left, right, 1 << 20, num_pdfs, p["tid2pdf"], {})
    return p


@pytest.fixture(scope="module")
def feats():
    rng = np.random.default_rng(77)
    sizes = [1, 40, 3, 211, 90]
    x = (2.0 * rng.standard_normal((sum(sizes), 40))).astype(np.float32)
    return x, np.concatenate([[0], np.cumsum(sizes)])


@pytest.mark.parametrize("prec", ["int8", "fp32"])
def test_subset_rows_equal_dense_columns(small_model, feats, prec):
    x, off = feats
    m = api.AcousticModelGpu(config=small_model["conf"], precision=prec)
    try:
        dense, am = m.nnet(x, off)
        rng = np.random.default_rng(5)
        for ids in (np.array([95]), rng.permutation(96)[:33], np.array([7, 7, 0, 95, 7]), np.arange(96)):
            m.set_output("subset", pdf_ids=ids)
            assert m.output_width() == ids.size
            sub, am2 = m.nnet(x, off)
            assert sub.shape == (x.shape[0], ids.size)
            assert np.array_equal(sub, dense[:, ids])        # bit for bit
            assert np.array_equal(am2, am)                   # argmax stays over all pdfs
        m.set_output("dense")
        again, _ = m.nnet(x, off)
        assert np.array_equal(again, dense)
    finally:
        m.close()


def test_topk_rows_equal_sorted_dense(small_model, feats):
    x, off = feats
    m = api.AcousticModelGpu(config=small_model["conf"], precision="int8")
    try:
        dense, am = m.nnet(x, off)
        for k in (1, 5, 32, 33, 96):
            m.set_output("topk", k=k)
            assert m.output_width() == 2 * k
            best, am2 = m.nnet(x, off)
            assert best.dtype == api.SCORED_PDF and best.shape == (x.shape[0], k)
            want_v, want_i = topk_reference(dense, k)
            assert np.array_equal(best["pdf"], want_i), k
            assert np.array_equal(best["loglik"], want_v), k
            assert np.array_equal(am2, best["pdf"][:, 0])    # entry 0 is the argmax, first maximum wins
            assert np.array_equal(am2, am)
    finally:
        m.close()


@pytest.mark.parametrize("num_pdfs,tied,ks", [
    (96, [3, 17, 18, 40, 77, 95], list(range(1, 97, 5))),
    # 300 equal entries per row: more than the kernel keeps as candidates, so the cut is made by
    # the exact "above the k-th key, then the lowest-numbered ties" path
    (512, list(range(5, 512, 5))[:100] + list(range(301, 501)), [1, 20, 64, 100, 128, 129, 200, 400, 512]),
])
def test_topk_ties_take_lowest_pdfs(tmp_path, feats, num_pdfs, tied, ks):
    """A group of pdfs shares one log-likelihood on every frame; wherever the k-th place falls
    inside that group the lowest-numbered ones are reported, in ascending order."""
    x, off = feats
    p = write_tied_model(str(tmp_path), hidden=64, num_pdfs=num_pdfs, tied=tied, seed=99)
    m = api.AcousticModelGpu(config=p["conf"], precision="fp32")
    try:
        dense, _ = m.nnet(x, off)
        assert np.array_equal(dense[:, tied[0]], dense[:, tied[-1]])
        cut_inside = 0
        for k in ks:
            m.set_output("topk", k=k)
            best, _ = m.nnet(x, off)
            want_v, want_i = topk_reference(dense, k)
            assert np.array_equal(best["pdf"], want_i), k
            assert np.array_equal(best["loglik"], want_v), k
            n_tied = np.isin(best["pdf"], tied).sum(axis=1)
            cut_inside += int(np.sum((n_tied > 0) & (n_tied < len(tied))))
        assert cut_inside > 0                                # the cut did fall inside the tied group
    finally:
        m.close()


def test_selected_output_through_forward_host_and_device(small_model, monkeypatch):
    """The full path (PCM in, host rows out, several chunks) writes the same selected rows."""
    import torch
    monkeypatch.setenv("CE_GPU_CHUNK_ROWS", "256")           # read at load: 6 utterances -> 3+ chunks
    pcm, soff = synth.synth_batch(6, n_samples=16000)
    m = api.AcousticModelGpu(config=small_model["conf"], precision="int8")
    try:
        dense, _, foff = m.forward(pcm, soff)
        ids = np.array([5, 90, 1, 44], np.int32)
        m.set_output("subset", pdf_ids=ids)
        sub, _, _ = m.forward(pcm, soff)
        assert np.array_equal(sub, dense[:, ids])
        m.set_output("topk", k=8)
        best, am, _ = m.forward(pcm, soff)
        want_v, want_i = topk_reference(dense, 8)
        assert np.array_equal(best["pdf"], want_i) and np.array_equal(best["loglik"], want_v)
        # device output buffer: rows of 2 k words
        out = torch.zeros((dense.shape[0], 16), dtype=torch.float32, device="cuda")
        m.forward(torch.from_numpy(pcm).cuda(), soff, loglik=out, want_argmax=False)
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(api.SCORED_PDF)
        assert np.array_equal(got["pdf"], want_i) and np.array_equal(got["loglik"], want_v)
    finally:
        m.close()


def test_set_output_errors(small_model, tmp_path):
    m = api.AcousticModelGpu(config=small_model["conf"], precision="int8")
    try:
        for kw in (dict(mode="topk", k=0), dict(mode="topk", k=97), dict(mode="subset", pdf_ids=[]),
                   dict(mode="subset", pdf_ids=[0, 96]), dict(mode="subset", pdf_ids=[-1])):
            with pytest.raises(api.CeGpuError):
                m.set_output(**kw)
        assert m.output_width() == 96                        # a rejected selection changes nothing
    finally:
        m.close()
    odd = synth.write_model(str(tmp_path), name="odd", hidden=64, num_pdfs=98, seed=3)
    m = api.AcousticModelGpu(config=odd["conf"], precision="int8")
    try:
        with pytest.raises(api.CeGpuError, match="num_pdfs"):
            m.set_output("topk", k=4)
    finally:
        m.close()


def test_rows_callback_per_chunk_rows_complete(small_model, monkeypatch):
    """Row ring: with a callback set, every chunk is announced in order as soon as ITS rows (and
    argmax) are complete in the caller's host buffers -- checked inside the callback, while later
    chunks are still being computed -- and the call returns after the last callback."""
    monkeypatch.setenv("CE_GPU_CHUNK_ROWS", "256")           # 2 utterances of 1 s per chunk
    pcm, soff = synth.synth_batch(9, n_samples=16000)
    soff = np.concatenate([soff[:4], [soff[3] + 100], soff[4:]])   # utterance 3 has no frame, 4 is shorter
    m = api.AcousticModelGpu(config=small_model["conf"], precision="int8")
    try:
        want, want_am, foff = m.forward(pcm, soff)
        ll = np.full_like(want, np.nan)
        am = np.full_like(want_am, -1)
        seen = []

        def ready(first_utt, n_utts, first_frame, n_frames):
            rows = slice(first_frame, first_frame + n_frames)
            seen.append((first_utt, n_utts, first_frame, n_frames,
                         bool(np.array_equal(ll[rows], want[rows])), bool(np.array_equal(am[rows], want_am[rows])),
                         bool(np.isnan(ll[first_frame + n_frames:]).all())))
        m.set_rows_callback(ready)
        m.forward(pcm, soff, loglik=ll, argmax=am)
        assert len(seen) >= 4
        assert all(s[4] and s[5] for s in seen)              # complete at callback time
        assert any(s[6] for s in seen[:-1])                  # ... while later rows had not arrived yet
        utts = [u for s in seen for u in range(s[0], s[0] + s[1])]
        assert utts == list(range(10))                        # every utterance once, in order
        assert sum(s[3] for s in seen) == int(foff[-1])
        for s in seen:
            assert s[2] == foff[s[0]] and s[3] == foff[s[0] + s[1]] - foff[s[0]]
        assert np.array_equal(ll, want) and np.array_equal(am, want_am)
        # top-k rows ride the same ring; a device buffer only gets the notifications
        import torch
        m.set_output("topk", k=4)
        seen.clear()
        best = np.zeros((want.shape[0], 4), api.SCORED_PDF)
        m.set_rows_callback(lambda *a: seen.append(a))
        m.forward(pcm, soff, loglik=best, want_argmax=False)
        assert sum(a[3] for a in seen) == int(foff[-1]) and np.array_equal(best["pdf"][:, 0], want_am)
        n = len(seen)
        d_out = torch.zeros((want.shape[0], 8), dtype=torch.float32, device="cuda")
        m.forward(torch.from_numpy(pcm).cuda(), soff, loglik=d_out, want_argmax=False)
        assert len(seen) >= n + 1
        n = len(seen)
        m.set_rows_callback(None)
        m.forward(pcm, soff, loglik=best, want_argmax=False)
        assert len(seen) == n
    finally:
        m.close()
