"""Full-size runs of the BASELINE configs through size-independent properties (the oracle cannot
finish these sizes in seconds): config 3 (512 ten-second utterances per GPU through the full TDNN,
int8) and config 5 (one hour of audio in 8 time shards)."""
import numpy as np
import pytest

from catears_b200 import api, formats as F, shard, synth

pytestmark = pytest.mark.gpu


def test_config3_full_batch_properties(tmp_path):
    """512 x 10 s utterances (64 distinct, tiled 8x), 6x1024 TDNN + 3072 pdfs, int8, several chunks:
      * copies of an utterance get bit-identical rows wherever they sit in the batch / chunking,
      * a single-utterance call reproduces its rows of the batch bit for bit,
      * every row is a normalised distribution: logsumexp(loglik + log prior) == 0 (1e-3),
      * argmax is the argmax of the returned row."""
    import torch
    stats = synth.default_cmvn_stats()
    m = synth.write_model(str(tmp_path / "tdnn"), name="tdnn", cmvn_stats=stats)
    base, _ = synth.synth_batch(64, 160000)
    n_utts = 512
    pcm = np.tile(base, n_utts // 64)
    off = np.arange(n_utts + 1, dtype=np.int64) * 160000
    am = api.AcousticModelGpu(config=m["conf"], precision="int8")
    d_pcm = torch.from_numpy(pcm).cuda()
    frames = n_utts * 998
    d_ll = torch.empty((frames, am.num_pdfs), dtype=torch.float32, device="cuda")
    d_am = torch.empty(frames, dtype=torch.int32, device="cuda")
    am.forward(d_pcm, off, loglik=d_ll, argmax=d_am, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    ll = d_ll.view(n_utts, 998, am.num_pdfs)
    assert bool(torch.isfinite(d_ll).all())
    for rep in range(1, n_utts // 64):                       # tiled copies: identical rows
        assert bool(torch.equal(ll[:64], ll[rep * 64:(rep + 1) * 64])), rep
    lp = torch.from_numpy(np.log(F.read_vector(m["prior"]))).cuda()
    lse = torch.logsumexp(d_ll[:64 * 998].double() + lp.double(), dim=1)
    assert float(lse.abs().max()) < 1e-3
    assert bool(torch.equal(d_ll.argmax(1).int(), d_am))
    for u in (0, 37, 63):                                    # batch == single, bit for bit
        l1, a1, _ = am.forward(pcm[u * 160000:(u + 1) * 160000])
        assert np.array_equal(l1, ll[u].cpu().numpy())
        assert np.array_equal(a1, d_am[u * 998:(u + 1) * 998].cpu().numpy())
    # selected outputs at full width (3072 pdfs): the 64 best per frame and a
    # 500-pdf subset of the first 64 utterances equal the dense rows' entries bit for bit
    sub_off, n_sub = off[:65], 64 * 998
    want_v, want_i = torch.sort(d_ll[:n_sub], dim=1, descending=True, stable=True)
    for k in (64, 200, 300):                                 # the three code paths of the kernel
        am.set_output("topk", k=k)
        d_best = torch.empty((n_sub, 2 * k), dtype=torch.float32, device="cuda")
        am.forward(d_pcm[:64 * 160000], sub_off, loglik=d_best, want_argmax=False)
        torch.cuda.synchronize()
        got = d_best.view(n_sub, k, 2)
        assert bool(torch.equal(got[:, :, 0], want_v[:, :k])), k
        assert bool(torch.equal(got[:, :, 1].contiguous().view(torch.int32), want_i[:, :k].int())), k
    ids = np.random.default_rng(8).permutation(am.num_pdfs)[:500].astype(np.int32)
    am.set_output("subset", pdf_ids=ids)
    d_sub = torch.empty((n_sub, 500), dtype=torch.float32, device="cuda")
    am.forward(d_pcm[:64 * 160000], sub_off, loglik=d_sub, want_argmax=False)
    torch.cuda.synchronize()
    assert bool(torch.equal(d_sub, d_ll[:n_sub][:, torch.from_numpy(ids).long().cuda()]))
    am.close()


def test_config5_hour_stream_in_8_time_shards(small_model, golden, tmp_path):
    """One hour of 16 kHz audio (359,998 frames): 8 contiguous time shards with recomputed halos ==
    the whole stream (AM context exact; the CMVN sums restart 600 frames before a shard, so values
    agree to fp32 rounding and the per-frame argmax almost everywhere)."""
    stats = tmp_path / "cmvn.bin"
    F.write_vector(str(stats), golden["cmvn_stats"])
    minute = synth.synth_utterance(7, 16000 * 60, seed=7)
    pcm = np.tile(minute, 60)
    m = api.AcousticModelGpu(nnet=small_model["nnet"], prior=small_model["prior"], left_context=13,
                             right_context=13, cmvn_stats=str(stats), precision="fp32")
    whole, am_whole, fo = m.forward(pcm)
    assert whole.shape == (359998, 96) and int(fo[-1]) == 359998
    ll, am = shard.forward_longform([m] * 8, pcm)
    assert ll.shape == whole.shape and np.isfinite(ll).all()
    assert np.abs(ll - whole).max() < 2e-3
    assert np.mean(am == am_whole) > 0.999
    m.close()
