#!/usr/bin/env python
"""bench.py -- audio-seconds per second of the fbank + CMVN + acoustic-model path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision int8|bf16|tf32|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's own CPU path, same metric

One step = one pass of the whole hot path (ce_gpu_forward: int16 PCM -> 40-mel fbank -> online
CMVN -> TDNN 6x1024 + 3072 pdfs -> log-likelihoods + argmax) over one batch of synthetic 10 s
utterances (4096 utterances over 8 GPUs = 512 per GPU; weak scaling, no
collective on the data path).  `value` is timed with the PCM already in HBM; `e2e` goes through
the C ABI with pinned HOST buffers (PCM H2D and the per-frame argmax D2H inside the timed region;
the 12 KB/frame log-likelihood matrix stays in HBM, SURVEY H6 -- `e2e_loglik` also copies it out).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-sec/sec (RTFx) fbank+CMVN+AM"
UNIT = "audio-s/s"
UTT_SECONDS = 10.0
FLOPS_PER_FRAME = 38158336          # 2*(200*1024 + 5*3072*1024 + 1024*3072)
FBANK_BYTES_PER_FRAME = 480         # 160 samples * 2 B + 40 mel * 4 B
DTYPES = {"int8": "u8", "bf16": "bf16", "tf32": "tf32", "fp32": "f32", "bf16x3": "bf16x3"}


def workload_config(utts_per_gpu, world, precision):
    frames = world * utts_per_gpu * 998
    return {"workload": "config 3 pipeline: %d synthetic 10 s 16 kHz utterances per GPU (%d in all), "
                        "40-mel fbank + online CMVN + TDNN 6x1024 (+splice) + 3072 pdfs, %s GEMMs"
                        % (utts_per_gpu, world * utts_per_gpu, precision),
            "frames_per_step": frames, "parallelism": "utterances sharded, no collective",
            "l2": "inputs larger than L2 (%.0f MB PCM + %.1f GB activations per GPU per step)"
                  % (utts_per_gpu * 0.32, utts_per_gpu * 998 * 8192 / 1e9)}


def model_dir():
    d = os.path.join(tempfile.gettempdir(), "ce_bench_model_%d" % os.getuid())
    from catears_b200 import synth
    conf = os.path.join(d, "tdnn.conf")
    if not os.path.exists(conf):
        tmp = d + ".tmp%d" % os.getpid()
        synth.write_model(tmp, name="tdnn", cmvn_stats=synth.default_cmvn_stats())
        try:
            os.rename(tmp, d)
        except OSError:
            pass                     # another rank won the race
    return d


# ------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation (oracle/_ref) or the oracle port
# ------------------------------------------------------------------------------------------

_W = {}


def _cpu_worker_init(conf, kind, variant=""):
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from catears_b200 import formats as F, synth
    _W["stats"] = synth.default_cmvn_stats()
    _W["kind"] = kind
    _W["conf"] = conf
    if kind == "reference":
        import ctypes as C
        from oracle import ref as R
        r = R.Ref(variant)
        _W["blas"] = "openblas" if r.set_sgemm("openblas") else "inorder"
        if _W["blas"] == "inorder":
            r.set_sgemm("inorder")
        _W["ref"] = r
        _W["am"] = r.L.ref_am_open(conf.encode())
        if not _W["am"]:
            raise RuntimeError("ref_am_open failed")
        _W["npdf"] = r.L.ref_am_num_pdfs(_W["am"])
    else:
        from oracle.port import Port
        d = os.path.dirname(conf)
        _W["port"] = Port()
        _W["nnet"] = os.path.join(d, "tdnn.nnet")
        _W["prior"] = F.read_vector(os.path.join(d, "tdnn.prior"))
        _W["blas"] = "port-inorder"


def _cpu_one_utt(u):
    """fbank -> CMVN -> AM log-likelihoods of synthetic utterance u on one core."""
    import ctypes as C
    from catears_b200 import synth
    pcm = synth.synth_utterance(u)
    t0 = time.perf_counter()
    if _W["kind"] == "reference":
        r = _W["ref"]
        feats = r.cmvn(_W["stats"], r.fbank(pcm))
        out = np.zeros((feats.shape[0] + 1, _W["npdf"]), np.float32)
        cols = C.c_int()
        n = r.L.ref_am_forward(_W["am"], feats, feats.shape[0], feats.shape[1], out, out.shape[0],
                               C.byref(cols))
        assert n == feats.shape[0], n
    else:
        p = _W["port"]
        feats = p.cmvn(_W["stats"], p.fbank(pcm))
        p.am_forward(_W["nnet"], _W["prior"], 13, 13, feats)
    return time.perf_counter() - t0, _W["blas"]


def _cpu_one_utt_int8(args):
    """The int8 composition (Quantize + MatMat_U8U8F32 + bias per Linear layer, gemmlowp as
    this build of the reference selects its kernel) on `seconds` of synthetic audio, one core."""
    u, seconds = args
    from catears_b200 import synth
    r = _W["ref"]
    d = os.path.dirname(_W["conf"])
    pcm = synth.synth_utterance(u, int(seconds * 16000))
    import ctypes as C
    if "u8" not in _W:                                   # model load + weight quantisation: not timed
        _W["u8"] = r.L.ref_u8_open(os.path.join(d, "tdnn.nnet").encode(), os.path.join(d, "tdnn.prior").encode(),
                                   13, 13)
        if not _W["u8"]:
            raise RuntimeError("ref_u8_open failed")
    t0 = time.perf_counter()
    feats = r.cmvn(_W["stats"], r.fbank(pcm))
    T = feats.shape[0]
    out = np.zeros((T + 26) * 3072, np.float32)
    rr, cc = C.c_int(), C.c_int()
    rc = r.L.ref_u8_forward(_W["u8"], feats, T, feats.shape[1], out, out.size, C.byref(rr), C.byref(cc), -1,
                            None, 0, None, None)
    assert rc == 0 and rr.value == T, (rc, rr.value)
    return time.perf_counter() - t0


def _cpu_worker_loop(args):
    """Runs utterances for at least `budget` seconds; returns (n, elapsed, blas)."""
    first, budget = args
    n, t, blas = 0, 0.0, ""
    while n == 0 or t < budget:
        dt, blas = _cpu_one_utt(first + n)
        t += dt
        n += 1
    return n, t, blas


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_kind():
    from oracle import ref as R
    return "reference" if R.available() else "port"


def make_pool(conf, kind, cores, variant=""):
    import multiprocessing as mp
    return mp.get_context("fork").Pool(cores, initializer=_cpu_worker_init, initargs=(conf, kind, variant))


def load_reference_in_parent():
    """dlopen the reference library in this (parent) process too: the work runs in forked pool workers,
    and the driver's loaded-library record is taken from the process it started."""
    from oracle import ref as R
    if R.available():
        R.Ref()


def cpu_baseline_int8(conf):
    """BASELINE.md section 4 / SURVEY 8d: the reference's int8 path on the host cores, as shipped
    (no -msse4.1: gemmlowp's scalar reference kernel) and built with -msse4.1 (its SSE4
    kernel; integer results identical), one worker per core; plus the float path on ONE core."""
    from oracle import ref as R
    out = {}
    cores = cpu_cores()
    for key, variant, seconds in (("int8_as_shipped", "", 3.0), ("int8_sse4", "_sse4", 10.0)):
        if not R.available(variant):
            continue
        pool = make_pool(conf, "reference", cores, variant)
        try:
            t = pool.map(_cpu_one_utt_int8, [(300000 + i, seconds) for i in range(cores)])
        finally:
            pool.close()
            pool.join()
        out[key] = {"value": round(sum(seconds / x for x in t), 2), "unit": UNIT, "cores": cores,
                    "kind": "reference",
                    "sample": "%d utterances of %.0f s, one per core, Quantize + MatMat_U8U8F32 per Linear layer "
                              "(gemmlowp %s kernel)" % (cores, seconds, "SSE4" if variant else "scalar reference")}
    if R.available():
        pool = make_pool(conf, "reference", 1)
        try:
            res = pool.map(_cpu_worker_loop, [(400000, 4.0)])
        finally:
            pool.close()
            pool.join()
        out["float_single_thread"] = {"value": round(res[0][0] * UTT_SECONDS / res[0][1], 2), "unit": UNIT,
                                      "cores": 1, "kind": "reference",
                                      "sample": "%d synthetic 10 s utterances on one core, float AM (cblas_sgemm = %s)"
                                                % (res[0][0], res[0][2])}
    return out


def cpu_baseline(conf, budget_s=6.0):
    """Bounded sample of the same workload on the host cores (utterance-parallel, one
    single-threaded worker per core)."""
    kind, cores = cpu_kind(), cpu_cores()
    pool = make_pool(conf, kind, cores)
    try:
        res = pool.map(_cpu_worker_loop, [(100000 + 64 * i, budget_s) for i in range(cores)])
    finally:
        pool.close()
        pool.join()
    n = sum(r[0] for r in res)
    value = sum(r[0] * UTT_SECONDS / r[1] for r in res)      # concurrent workers: rates add
    return {"value": round(value, 2), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d synthetic 10 s utterances (>= %.0f s of work per core), float AM "
                      "(cblas_sgemm = %s), fbank+CMVN+AM per utterance on one core, one worker "
                      "per core" % (n, budget_s, res[0][2])}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    conf = os.path.join(model_dir(), "tdnn.conf")
    kind, cores = cpu_kind(), cpu_cores()
    if kind == "reference":
        load_reference_in_parent()
    pool = make_pool(conf, kind, cores)
    try:
        def step(i):
            return pool.map(_cpu_one_utt, [200000 + i * cores + j for j in range(cores)])
        for i in range(args.warmup):
            step(i)
        t0 = time.perf_counter()
        blas = ""
        for i in range(args.steps):
            blas = step(args.warmup + i)[0][1]
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    audio = args.steps * cores * UTT_SECONDS
    value = audio / dt
    sample = ("each step = %d synthetic 10 s utterances (one per host core), float AM "
              "(cblas_sgemm = %s)" % (cores, blas))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.utts_per_gpu, args.gpus, args.precision),
        "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except OSError:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(self.NAMES, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [x for x in vis.split(",") if x.strip()]
        if local_rank < len(ids) and ids[local_rank].strip().isdigit():
            return int(ids[local_rank])
    return local_rank


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    conf = os.path.join(model_dir(), "tdnn.conf")

    # CPU baseline first (fork-based pool must start before CUDA is initialised).
    base = base_more = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        base = cpu_baseline(conf)
        if cpu_kind() == "reference":
            load_reference_in_parent()
            base_more = cpu_baseline_int8(conf)

    import torch
    import torch.distributed as dist
    from catears_b200 import api, synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_utts = args.utts_per_gpu
    n_samples = int(UTT_SECONDS * 16000)
    pcm_np, off = synth.synth_batch(n_utts, n_samples, first=rank * n_utts)
    frames = int(api.frame_offsets(off)[-1])
    model = api.AcousticModelGpu(config=conf, precision=args.precision, device=local_rank)

    h_pcm = torch.from_numpy(pcm_np).pin_memory()
    d_pcm = h_pcm.cuda(non_blocking=True)
    d_ll = torch.empty((frames, model.num_pdfs), dtype=torch.float32, device="cuda")
    d_am = torch.empty(frames, dtype=torch.int32, device="cuda")
    h_am = torch.empty(frames, dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream()

    def step_device():
        model.forward(d_pcm, off, loglik=d_ll, argmax=d_am, stream=stream)

    def step_e2e():
        model.forward(h_pcm.numpy(), off, loglik=d_ll, argmax=h_am.numpy(), stream=stream)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        local = e0.elapsed_time(e1)
        ms = torch.tensor([local], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        timed.local = local
        return float(ms.item())

    audio_per_step = world * n_utts * UTT_SECONDS

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    time.sleep(0.3)
    api.launch_count(reset=True)
    if os.environ.get("BENCH_PROFILE_OVERLAP"):
        api.profile_enable(True)
    ms = timed(step_device, args.steps)
    ms_local = timed.local
    if os.environ.get("BENCH_PROFILE_OVERLAP"):
        sys.stderr.write("overlapped per-kernel ms/step: %s\n" % {k: round(v[0] / args.steps, 3)
                                                                 for k, v in api.profile_read().items()})
        api.profile_enable(False)
    launches = api.launch_count()
    clocks = sampler.stop()
    if world > 1:                                          # every rank watches its own GPU: slowest / fastest median
        mhz = torch.tensor([clocks["sm_mhz"] or 0.0], device="cuda")
        lo, hi = mhz.clone(), mhz.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        clocks["sm_mhz_min_over_ranks"] = float(lo.item())
        clocks["sm_mhz_max_over_ranks"] = float(hi.item())
        mine = torch.tensor([ms_local], device="cuda")
        fastest = mine.clone()
        dist.all_reduce(fastest, op=dist.ReduceOp.MIN)
        clocks["ms_per_step_fastest_rank"] = round(float(fastest.item()) / args.steps, 3)
    value = audio_per_step * args.steps / (ms * 1e-3)

    # Per-kernel durations for the roofline: the same steps once more with CUDA events around
    # every launch (the library's own events, recorded on the stream the kernels run on).
    api.profile_enable(True)
    ms_serial = timed(step_device, args.steps)
    prof = api.profile_read()
    api.profile_enable(False)

    # e2e: host buffers through the C ABI
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e = audio_per_step * args.steps / (ms_e2e * 1e-3)

    e2e_ll = None
    if not args.no_extra_legs and world == 1:
        h_ll = torch.empty((frames, model.num_pdfs), dtype=torch.float32).pin_memory()

        def step_ll():
            model.forward(h_pcm.numpy(), off, loglik=h_ll.numpy(), argmax=h_am.numpy(), stream=stream)
        step_ll()
        n_ll = max(1, min(2, args.steps))
        ms_ll = timed(step_ll, n_ll)
        e2e_ll = audio_per_step * n_ll / (ms_ll * 1e-3)
        del h_ll

    # The same with the k best (loglik, pdf) pairs per frame instead of the dense row (SURVEY 8f
    # rank 4): 8 k bytes a frame over PCIe instead of 4 num_pdfs.
    e2e_topk = None
    if not args.no_extra_legs and args.e2e_topk > 0 and world == 1:
        k = args.e2e_topk
        h_best = torch.empty((frames, 2 * k), dtype=torch.float32).pin_memory()
        model.set_output("topk", k=k)

        def step_topk():
            model.forward(h_pcm.numpy(), off, loglik=h_best.numpy(), argmax=h_am.numpy(), stream=stream)
        step_topk()
        n_tk = max(1, min(3, args.steps))
        ms_topk = timed(step_topk, n_tk)
        e2e_topk = audio_per_step * n_tk / (ms_topk * 1e-3)
        model.set_output("dense")
        del h_best

    # The float paths of config 4 on the same batch (short legs: HBM-resident value only).  bf16x3 and
    # fp32 (3xTF32) meet the 1e-3 log-likelihood bar; bf16 and tf32 single pass are fast modes outside it.
    other = {}
    if not args.no_extra_legs and world == 1 and args.precision == "int8":
        for prec, in_tol in (("bf16x3", True), ("fp32", True), ("bf16", False), ("tf32", False)):
            m2 = api.AcousticModelGpu(config=conf, precision=prec, device=local_rank)

            def step2():
                m2.forward(d_pcm, off, loglik=d_ll, argmax=d_am, stream=stream)
            step2()
            ms2 = timed(step2, 2)
            other[prec] = {"value": round(audio_per_step * 2 / (ms2 * 1e-3), 1), "unit": UNIT,
                           "ms_per_step": round(ms2 / 2, 3), "within_1e-3_of_reference_float": in_tol}
            m2.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    bf16_sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    bf16_burst = peaks.get("bf16_tflops", 1590.0)
    scale = {"int8": 2.0, "bf16": 1.0, "bf16x3": 1.0 / 3.0, "tf32": 0.5, "fp32": 0.5}[args.precision]
    if args.precision == "int8":
        # no int8 figure was measured by the driver: the architectural ratio is 2x the bf16 rate.  The
        # timed region is a fraction of a second at full clocks, so the BURST figure is the denominator.
        tensor_peak = 2.0 * bf16_burst
        peak_note = "2 x bf16_tflops (burst: the timed region runs at full clocks; int8 tcgen05 rate is twice " \
                    "the bf16 rate); %s" % peak_src
    elif args.precision == "bf16":
        tensor_peak = bf16_burst
        peak_note = "bf16_tflops (burst); %s" % peak_src
    elif args.precision == "bf16x3":
        tensor_peak = bf16_burst / 3.0
        peak_note = "bf16_tflops (burst) / 3 (three bf16 products per algorithmic multiply-add); %s" % peak_src
    else:
        tensor_peak = 0.5 * bf16_burst
        peak_note = "0.5 x bf16_tflops (burst; tf32 rate is half the bf16 rate%s); %s" % (
            "; the fp32 path spends 3 tf32 passes per product" if args.precision == "fp32" else "", peak_src)
    gemm_ms, gemm_n = prof["gemm"]
    fb_ms, fb_n = prof["fbank"]
    # int8 models: the output layer runs fused with LogSoftmax + prior + argmax and is timed in the "finalize"
    # category (the library's rule, catears_b200/csrc/nnet.cc); the roofline's dominant kernel is then the
    # plain GEMM of the other six layers, and the fused launch gets its own object below
    fused_out = args.precision == "int8" and os.environ.get("CE_GPU_FUSED_OUTPUT", "1") != "0"
    flops_out = 2.0 * 1024 * model.num_pdfs        # the output layer's share of FLOPS_PER_FRAME
    flops_step = frames * (FLOPS_PER_FRAME - flops_out if fused_out else FLOPS_PER_FRAME)   # this rank's share
    traffic = hbm_step = pipe_ncu = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        tr = json.load(open(tr_path))
        per_row = tr.get("gemm_%s_dram_bytes_per_row" % args.precision)
        if per_row and gemm_n:
            # ncu measured the launches of one 128-utterance chunk; the bytes per activation row (1024
            # rows per 10 s utterance) of the average GEMM launch, times the rows one launch of this
            # run's chunking covers
            launches_per_step = gemm_n / max(1, args.steps)
            traffic = int(per_row * n_utts * 1024 / max(1.0, launches_per_step / (6.0 if fused_out else 7.0)))
        if args.precision == "int8" and tr.get("step_int8_dram_bytes_per_row"):
            hbm_step = int(tr["step_int8_dram_bytes_per_row"] * n_utts * 1024)
        pipe_ncu = tr.get("gemm_%s_tensor_pipe_active_pct" % args.precision)
    achieved = flops_step * args.steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {
        "kernel": "gemm_kernel<%s> (tcgen05 cta_group::2, %d launches/step)" % (args.precision, gemm_n // max(1, args.steps)),
        "bound": "tensor", "achieved": round(achieved, 2), "peak": round(tensor_peak, 1), "unit": "TFLOP/s",
        "frac": round(achieved / tensor_peak, 4), "traffic": traffic,
        "frac_burst": round(achieved / (scale * bf16_burst), 4),
        "frac_sustained": round(achieved / (scale * bf16_sustained), 4),
        "tensor_pipe_active_pct_ncu": pipe_ncu,
        "peak_source": peak_note,
        "algorithmic_flops_per_launch": round(flops_step * args.steps / max(1, gemm_n)),
        "avg_launch_ms": round(gemm_ms / max(1, gemm_n), 4),
        "share_of_step": round(gemm_ms / ms_serial, 4),
        "timing": "CUDA events around every launch in a second pass of the same steps "
                  "(%.3f ms/step with the events in)" % (ms_serial / args.steps),
        "note": ("the six layers in front of the output layer; the output layer runs fused with LogSoftmax + prior + "
                 "argmax and is timed as kernel_ms_per_step.finalize (roofline_output)") if fused_out else
                "all seven layers; log-softmax is a separate launch (kernel_ms_per_step.finalize)",
    }
    roofline_output = None
    fin_ms, fin_n = prof["finalize"]
    if fused_out and fin_ms > 0:
        out_tflops = frames * flops_out * args.steps / (fin_ms * 1e-3) / 1e12
        out_gbs = frames * 4.0 * model.num_pdfs * args.steps / (fin_ms * 1e-3) / 1e9
        roofline_output = {
            "kernel": "gemm_kernel<int8, LSM> (output layer + LogSoftmax + prior + argmax, %d launches/step)"
                      % (fin_n // max(1, args.steps)),
            "bound": "tensor", "achieved": round(out_tflops, 2), "peak": round(tensor_peak, 1), "unit": "TFLOP/s",
            "frac": round(out_tflops / tensor_peak, 4),
            "hbm_write_gbs": round(out_gbs, 1), "hbm_write_frac": round(out_gbs / hbm_peak, 4),
            "avg_launch_ms": round(fin_ms / max(1, fin_n), 4), "share_of_step": round(fin_ms / ms_serial, 4),
            "note": "algorithmic FLOPs counted once: the kernel multiplies every tile twice (row statistics, then the "
                    "finished rows) instead of writing and re-reading 4 x 3072 B of logits per frame; its epilogue "
                    "(2 x 3072 accumulator columns per row) and shared-memory bandwidth bound it, not the tensor "
                    "pipe (DESIGN.md 5.3: MMA floor 507 us, epilogue alone 551 us, together 694 us per 131072 rows)"}
    fb_gbs = frames * FBANK_BYTES_PER_FRAME * args.steps / (fb_ms * 1e-3) / 1e9 if fb_ms > 0 else 0.0
    roofline_fbank = {"kernel": "fbank_kernel", "bound": "hbm", "achieved": round(fb_gbs, 1),
                      "peak": hbm_peak, "unit": "GB/s", "frac": round(fb_gbs / hbm_peak, 4),
                      "avg_launch_ms": round(fb_ms / max(1, fb_n), 4),
                      "share_of_step": round(fb_ms / ms_serial, 4),
                      "frames_per_s": round(frames * args.steps / (fb_ms * 1e-3), 0) if fb_ms > 0 else None,
                      "note": "an fp32 FFT at ~740 warp-instructions a frame is bound by issue slots and "
                              "shared-memory wavefronts, not by HBM (DESIGN.md section 5): the fractions "
                              "below are ncu's, of the same kernel at HEAD (profiles/traffic.json)"}
    if os.path.exists(tr_path):
        tr = json.load(open(tr_path))
        if tr.get("fbank_issue_active_pct"):
            roofline_fbank["issue_frac"] = round(tr["fbank_issue_active_pct"] / 100.0, 4)
            roofline_fbank["smem_frac"] = round(tr["fbank_l1tex_throughput_pct"] / 100.0, 4)
    out = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPES[args.precision],
        "data": "synthetic",
        "config": workload_config(n_utts, world, args.precision),
        "clocks": clocks,
        "e2e": {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": int(pcm_np.nbytes) * world,
                "d2h_bytes_per_step": int(frames * 4) * world,
                "note": "pinned host PCM in, per-frame argmax out; log-likelihoods stay in HBM (H6)"},
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_output": roofline_output, "roofline_fbank": roofline_fbank,
        "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in prof.items()},
    }
    if hbm_step is not None:
        out["hbm_bytes_per_step"] = {"value": hbm_step, "algorithmic": int(frames * (320 + 4 * model.num_pdfs)),
                                     "source": "ncu dram__bytes_read+write of every launch of one 128-utterance "
                                               "chunk (profiles/traffic.json), scaled by rows"}
    if other:
        out["float_paths"] = other
    if base_more:
        out["cpu_baselines_more"] = base_more
    if e2e_ll is not None:
        out["e2e_loglik"] = {"value": round(e2e_ll, 1), "unit": UNIT,
                             "d2h_bytes_per_step": int(frames * (4 + 4 * model.num_pdfs))}
    if e2e_topk is not None:
        out["e2e_topk"] = {"value": round(e2e_topk, 1), "unit": UNIT, "k": args.e2e_topk,
                           "d2h_bytes_per_step": int(frames * (4 + 8 * args.e2e_topk))}
    if base is not None:
        out["cpu_baseline"] = base
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_frontend(args):
    """fbank + CMVN only, 10,000 synthetic 10 s utterances (256 distinct ones
    tiled), 40 or 80 mel bins.  One JSON line with the fbank kernel's HBM roofline."""
    import torch
    from catears_b200 import api, synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    n_utts = args.frontend_utts // world
    mel = args.mel
    base, _ = synth.synth_batch(min(256, n_utts), 160000, first=rank * 256)
    pcm_np = np.tile(base, (n_utts + 255) // 256)[:n_utts * 160000]
    off = np.arange(n_utts + 1, dtype=np.int64) * 160000
    foff = api.frame_offsets(off)
    frames = int(foff[-1])
    stats = synth.default_cmvn_stats(mel=mel)
    h_pcm = torch.from_numpy(pcm_np).pin_memory()
    d_pcm = h_pcm.cuda()
    d_fb = torch.empty((frames, mel), dtype=torch.float32, device="cuda")
    d_out = torch.empty_like(d_fb)
    h_out = torch.empty((frames, mel), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream()

    def step_device():
        api.fbank(d_pcm, off, num_mel=mel, out=d_fb, device=local_rank, stream=stream)
        api.cmvn(stats, d_fb, foff, out=d_out, device=local_rank, stream=stream)

    def step_e2e():
        api.fbank(h_pcm.numpy(), off, num_mel=mel, out=d_fb, device=local_rank, stream=stream)
        api.cmvn(stats, d_fb, foff, out=h_out.numpy(), device=local_rank, stream=stream)

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(args.warmup):
        step_device()
    api.launch_count(reset=True)
    ms = timed(step_device, args.steps)
    launches = api.launch_count()
    api.profile_enable(True)
    timed(step_device, args.steps)
    prof = api.profile_read()
    api.profile_enable(False)
    step_e2e()
    ms_e2e = timed(step_e2e, max(1, args.steps // 2)) / max(1, args.steps // 2) * args.steps
    if rank != 0:
        return
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    audio = world * n_utts * UTT_SECONDS * args.steps
    fb_ms, fb_n = prof["fbank"]
    bytes_per_frame = 320 + 4 * mel
    gbs = frames * bytes_per_frame * args.steps / (fb_ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "audio-sec/sec (RTFx) fbank+CMVN", "value": round(audio / (ms * 1e-3), 1), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config 2: fbank+CMVN only, %d synthetic 10 s 16 kHz utterances per GPU "
                               "(256 distinct, tiled), %d mel bins" % (n_utts, mel),
                   "frames_per_step": frames * world,
                   "l2": "inputs larger than L2 (%.1f GB PCM per GPU per step)" % (pcm_np.nbytes / 1e9)},
        "e2e": {"value": round(audio / (ms_e2e * 1e-3), 1), "unit": UNIT,
                "h2d_bytes_per_step": int(pcm_np.nbytes) * world,
                "d2h_bytes_per_step": int(frames * mel * 4) * world},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "fbank_kernel", "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak,
                     "unit": "GB/s", "frac": round(gbs / hbm_peak, 4), "traffic": None,
                     "algorithmic_bytes_per_launch": frames * bytes_per_frame,
                     "avg_launch_ms": round(fb_ms / max(1, fb_n), 4),
                     "note": "fp32-issue and shared-memory bound, not HBM bound: see DESIGN.md section 5"},
        "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]},
    }))


def run_longform(args):
    """one hour of 16 kHz audio (one minute of synthetic audio tiled), split into
    contiguous time shards with recomputed halos (ce_gpu_time_shards: left context + 600 frames of
    CMVN history before, right context after), one shard per GPU (8 shards back to back at N = 1);
    full fbank + CMVN + TDNN pipeline, log-likelihoods and argmax stay in HBM."""
    import torch
    from catears_b200 import api, shard, synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    conf = os.path.join(model_dir(), "tdnn.conf")
    model = api.AcousticModelGpu(config=conf, precision=args.precision, device=local_rank)
    minute = synth.synth_utterance(7, 16000 * 60, seed=7)
    pcm = np.tile(minute, 60)
    total = int(api.frame_offsets([0, pcm.size])[-1])
    # Shards of --shard-frames kept frames, a contiguous run of them per GPU, evaluated as ONE batch per GPU.
    # (Round 1 needed ~90 shards of 4096 frames -- 15 % halo recomputation -- because the online CMVN chain
    # of a 45 000-frame shard took 5 ms on its single CTA; the chain now costs 16 clocks a frame and a long
    # utterance's bins spread over several CTAs: 16 384-frame shards measure best, 358 k x against 326 k x.)
    n_shards = max(world, (total + args.shard_frames - 1) // args.shard_frames)
    if args.exact:
        # --exact: the hour as ONE utterance -- the online CMVN chain runs from the first frame (bit-identical to
        # the reference's frame-by-frame evaluation of the whole stream: no warm-up halo, no fp32-rounding
        # difference), its bins spread over ten CTAs; single GPU (the chain cannot be cut across GPUs exactly)
        if world > 1:
            raise SystemExit("--workload longform --exact runs on one GPU")
        n_shards = 1
    kb, ke, fb, fe = api.time_shards(total, n_shards, model.left_context, model.right_context, 600)
    mine = list(range(n_shards * rank // world, n_shards * (rank + 1) // world))
    parts, off = [], [0]
    for p in mine:
        s0, s1 = shard.frame_to_sample_range(int(fb[p]), int(fe[p]))
        parts.append(pcm[s0:s1])
        off.append(off[-1] + (s1 - s0))
    off = np.array(off, np.int64)
    d_pcm = torch.from_numpy(np.concatenate(parts)).cuda()
    fed = int(api.frame_offsets(off)[-1])
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    feed = None
    if args.feed == "none":
        d_ll = torch.empty((fed, model.num_pdfs), dtype=torch.float32, device="cuda")
        d_am = torch.empty(fed, dtype=torch.int32, device="cuda")

        def step():
            model.forward(d_pcm, off, loglik=d_ll, argmax=d_am, stream=stream)
    else:
        # config 5 "feeding the CPU decoder": the k best (loglik, pdf) pairs of every frame go to pinned
        # host memory chunk by chunk (row ring, ce_gpu_model_set_rows_callback) and a pool of host
        # threads consumes every chunk as it lands -- a STUB consumer (the reference bundles no HCLG,
        # SURVEY D5): it reads every row once (best score per frame), like a decoder's first touch.
        from concurrent.futures import ThreadPoolExecutor
        k = args.feed_topk
        model.set_output("topk", k=k)
        h_rows = torch.empty((fed, 2 * k), dtype=torch.float32).pin_memory()
        h_am = torch.empty(fed, dtype=torch.int32).pin_memory()
        rows_np = h_rows.numpy()
        pool = ThreadPoolExecutor(max_workers=args.feed_threads)
        pending, seen = [], [0]

        def consume(f0, nf):
            blk = rows_np[f0:f0 + nf]
            return float(blk[:, 0].max()) if nf else 0.0       # touches the chunk's rows

        def on_rows(first_utt, n_utts, first_frame, n_frames):
            seen[0] += n_frames
            pending.append(pool.submit(consume, first_frame, n_frames))
        model.set_rows_callback(on_rows)

        def step():
            model.forward(d_pcm, off, loglik=rows_np, argmax=h_am.numpy(), stream=stream)
            for f in pending:
                f.result()
            del pending[:]
        feed = {"rows": "top-%d (loglik, pdf) pairs per frame, %d B a frame, pinned host" % (k, 8 * k),
                "consumer": "stub: %d host threads, each chunk's rows read once as it lands (no HCLG is bundled, "
                            "SURVEY D5)" % args.feed_threads}

    for _ in range(args.warmup):
        step()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    ms = torch.tensor([wall_ms if feed else e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(ms.item())
        out = {
            "metric": METRIC, "value": round(3600.0 * args.steps / (ms * 1e-3), 1), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": DTYPES[args.precision],
            "data": "synthetic",
            "config": {"workload": "config 5: one hour of 16 kHz audio (%d frames) in %d time shards with "
                                   "recomputed halos (L + 600 CMVN-history frames before, R after), %d shard(s) "
                                   "per GPU as one batch (%d frames fed; 1 shard = the whole stream as one utterance, "
                                   "exact online CMVN), fbank + CMVN + TDNN, %s GEMMs; %s"
                                   % (total, n_shards, len(mine), fed, args.precision,
                                      "rows to a host consumer pool" if feed else "log-likelihoods stay in HBM"),
                       "frames_per_step": total,
                       "timing": "wall clock around the steps incl. the consumers, max over ranks" if feed
                                 else "CUDA events, max over ranks"}}
        if feed:
            feed["frames_per_s"] = round(total * args.steps / (ms * 1e-3), 1)
            out["decoder_feed"] = feed
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_streaming(args):
    """The streaming form: --streams live utterances, each receiving
    --stream-ms milliseconds of new audio per micro-batch call through ce_gpu_streams_* (per-stream
    state on the device), rows = the 64 best (loglik, pdf) pairs per frame to the host.  One step =
    one call; wall time per call (the call returns with the rows on the host) and the kernel time
    inside it.  Single GPU."""
    import time
    from catears_b200 import api, synth
    conf = os.path.join(model_dir(), "tdnn.conf")
    model = api.AcousticModelGpu(config=conf, precision=args.precision)
    model.set_output("topk", k=64)
    n, per_call = args.streams, 16 * args.stream_ms
    streams = api.StreamSet(model, n)
    slots = [streams.open() for _ in range(n)]
    audio = synth.synth_utterance(11, per_call * (2 * args.steps + args.warmup + 1), seed=11)
    eos = [False] * n

    def step(c):
        piece = audio[c * per_call:(c + 1) * per_call]
        return streams.process(slots, [piece] * n, eos)
    for c in range(args.warmup):
        step(c)
    wall = []
    streams.call_stats(reset=True)
    for c in range(args.warmup, args.warmup + args.steps):
        t0 = time.perf_counter()
        rows = step(c)
        wall.append(1e3 * (time.perf_counter() - t0))
    n_calls, enq_us, tot_us = streams.call_stats()
    # kernel time per call: the same calls once more with CUDA events around every launch
    api.profile_enable(True)
    for c in range(args.warmup + args.steps, args.warmup + 2 * args.steps):
        step(c)
    prof = api.profile_read()
    api.profile_enable(False)
    ms = float(np.median(wall))
    print(json.dumps({
        "metric": METRIC, "value": round(n * args.stream_ms * 1e-3 / (ms * 1e-3), 1), "unit": UNIT,
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPES[args.precision],
        "data": "synthetic",
        "kernel_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in prof.items()},
        "abi_ms_per_call": {"total": round(tot_us / max(1, n_calls) * 1e-3, 3),
                            "until_all_work_is_queued": round(enq_us / max(1, n_calls) * 1e-3, 3),
                            "kernels": round(sum(v[0] for v in prof.values()) / args.steps, 3),
                            "note": "inside ce_gpu_streams_process (C ABI); ms_per_step is the Python caller's wall "
                                    "time (ctypes marshalling of the per-slot arrays included)"},
        "config": {"workload": "streaming: %d live streams x %d ms of new audio per call "
                               "(ce_gpu_streams_process, state on the device), fbank + CMVN + TDNN, %s GEMMs, "
                               "rows = 64 best (loglik, pdf) pairs per frame to the host; median wall time per call"
                               % (n, args.stream_ms, args.precision),
                   "rows_per_step": int(sum(r.shape[0] for r in rows))}}))
    streams.close()
    model.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="int8", choices=["int8", "bf16", "tf32", "fp32", "bf16x3"])
    ap.add_argument("--utts-per-gpu", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-loglik", action="store_true", help="(kept for compatibility: now part of the default line)")
    ap.add_argument("--no-extra-legs", action="store_true",
                    help="skip the short extra legs (e2e with dense / top-k rows to the host, float paths)")
    ap.add_argument("--e2e-topk", type=int, default=64,
                    help="also time e2e with the k best (loglik, pdf) pairs per frame copied to the host")
    ap.add_argument("--workload", default="pipeline", choices=["pipeline", "frontend", "longform", "streaming"],
                    help="pipeline = the headline fbank+CMVN+AM step (default); frontend = config 2; "
                         "longform = config 5 (one hour in time shards); streaming = live streams in "
                         "micro-batches (one GPU)")
    ap.add_argument("--feed", default="none", choices=["none", "topk"],
                    help="longform: rows to pinned host memory + a pool of consumer threads (config 5's decoder feed)")
    ap.add_argument("--feed-topk", type=int, default=64)
    ap.add_argument("--shard-frames", type=int, default=16384,
                    help="longform: kept frames per time shard (each shard recomputes L + 600 + R halo frames)")
    ap.add_argument("--exact", action="store_true",
                    help="longform: the whole stream as one utterance (exact online CMVN, one GPU)")
    ap.add_argument("--feed-threads", type=int, default=8)
    ap.add_argument("--streams", type=int, default=512)
    ap.add_argument("--stream-ms", type=int, default=100)
    ap.add_argument("--frontend-utts", type=int, default=10000)
    ap.add_argument("--mel", type=int, default=40)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = max(args.warmup, 1)
    if args.workload == "frontend" and args.impl == "native":
        run_frontend(args)
    elif args.workload == "streaming" and args.impl == "native":
        run_streaming(args)
    elif args.workload == "longform" and args.impl == "native":
        if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
            os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                       "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                                       "--master-port", "29518"] + sys.argv)
        run_longform(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: plain `python bench.py --gpus N` re-launches itself under torchrun
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                                   "--master-port", "29517"] + sys.argv)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
