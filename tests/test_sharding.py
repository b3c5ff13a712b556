"""Host-side sharding logic: plans, and a world_size-2 gloo run on the CPU in which
every rank takes its own utterance group and the union equals the single-process result.  The
per-utterance computation in the 2-rank test is the oracle's fbank (tests may use it) -- the
property under test is the partition / reassembly, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest

from catears_b200 import api, shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    for n_utts in (1, 2, 7, 64, 513):
        frames = rng.integers(0, 2000, n_utts)
        fo = np.concatenate([[0], np.cumsum(frames)])
        for parts in (1, 2, 4, 8):
            pb = api.partition(fo, parts)
            assert pb[0] == 0 and pb[-1] == n_utts and (np.diff(pb) >= 0).all()
            load = np.array([fo[pb[p + 1]] - fo[pb[p]] for p in range(parts)])
            assert load.sum() == fo[-1]
            if n_utts >= 8 * parts:
                assert load.max() - fo[-1] / parts <= frames.max()      # within one utterance of ideal
    fo = np.arange(0, 4097) * 998
    pb = api.partition(fo, 8)
    assert list(np.diff(pb)) == [512] * 8                               # config 3: 512 utts per GPU


def test_time_shards_cover_stream_with_halos():
    total = 359998                                                      # 1 h of audio (config 5)
    kb, ke, fb, fe = api.time_shards(total, 8, 13, 13, 600)
    assert kb[0] == 0 and ke[-1] == total and (kb[1:] == ke[:-1]).all()
    assert (fb == np.maximum(0, kb - 613)).all() and (fe == np.minimum(total, ke + 13)).all()
    redundant = ((fe - fb) - (ke - kb)).sum() / total
    assert redundant < 0.015                                            # 1.4 %
    s0, s1 = shard.frame_to_sample_range(int(fb[3]), int(fe[3]))
    assert (s1 - s0 - 400) // 160 + 1 == fe[3] - fb[3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from catears_b200 import synth
    from oracle.port import Port
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sizes = [4000, 16000, 400, 8000, 0, 12000, 5000]
    pcm = np.concatenate([synth.synth_utterance(u, n) for u, n in enumerate(sizes)])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    u0, u1 = shard.my_utterances(off, rank, world)
    port_ = Port()
    mine = [port_.fbank(pcm[off[u]:off[u + 1]]) for u in range(u0, u1) if sizes[u] >= 400]
    gathered = [None] * world
    dist.all_gather_object(gathered, (u0, u1, mine))                    # test-only reassembly
    if rank == 0:
        parts = sorted(gathered)
        assert parts[0][0] == 0 and parts[-1][1] == len(sizes)
        for a, b in zip(parts[:-1], parts[1:]):
            assert a[1] == b[0]
        got = np.concatenate([x for p in parts for x in p[2]])
        want = np.concatenate([port_.fbank(pcm[off[u]:off[u + 1]]) for u in range(len(sizes)) if sizes[u] >= 400])
        assert np.array_equal(got, want)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_partition_roundtrip(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


@pytest.mark.gpu
def test_longform_time_shards_match_whole_stream(small_model, golden, tmp_path):
    """Config 5 in miniature: a 60 s stream in 4 time shards on one GPU == the whole stream
    (AM rows exact in fp32 mode up to the CMVN restart, which is fp32 rounding)."""
    from catears_b200 import formats as F, synth
    stats = tmp_path / "cmvn.bin"
    F.write_vector(str(stats), golden["cmvn_stats"])
    pcm = synth.synth_utterance(7, 16000 * 60, seed=7)
    m = api.AcousticModelGpu(nnet=small_model["nnet"], prior=small_model["prior"], left_context=13,
                             right_context=13, cmvn_stats=str(stats), precision="fp32")
    whole, am_whole, _ = m.forward(pcm)
    ll, am = shard.forward_longform([m, m, m, m], pcm)
    assert ll.shape == whole.shape
    assert np.abs(ll - whole).max() < 1e-3
    assert np.mean(am == am_whole) > 0.999
    m.close()


@pytest.mark.gpu
def test_sharded_batch_equals_whole_batch(small_model):
    """1 GPU == N GPUs: every rank's group evaluated separately and concatenated is bit-identical
    to the whole batch (int8)."""
    from catears_b200 import synth
    pcm, off = synth.synth_batch(6, 16000)
    m = api.AcousticModelGpu(config=small_model["conf"], precision="int8")
    whole, am_whole, fo = m.forward(pcm, off)
    parts = [shard.forward_sharded(m, pcm, off, r, 4) for r in range(4)]
    ll = np.concatenate([p[0] for p in parts])
    assert np.array_equal(ll, whole)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), am_whole)
    m.close()
