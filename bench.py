#!/usr/bin/env python
"""bench.py -- DenseCap hot-path throughput on B200 (BASELINE.json metric: images/sec @720x600,
1000 proposals, 15-token greedy decode, random-init VGG-16; synthetic data).

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, C ABI)
  python bench.py --impl reference --gpus N ...            # the reference's CPU path restated (oracle), host cores

A step = one pass of the whole test-time forward (VGG-16 trunk -> LocalizationLayer -> recognition head ->
LSTM decode -> final NMS) over one batch of synthetic images.  `value` times K steps with inputs resident in
HBM (CUDA events, barrier + synchronize on both sides, max over ranks); `e2e` times the same K steps through the
host-buffer C-ABI call (pinned H2D of the image and D2H of the results inside the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_H, IMG_W = 600, 720
NUM_PROPOSALS = 1000
SEQ_LEN = 15
CONV_TRUNK_GFLOP = 264.578          # SURVEY.md 8d (ceil-mode pooling)
METRIC = "images/sec @720x600, 1000 proposals, 15-tok decode"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d.get("bf16_tflops", 1590.0), "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: the sampler runs from before the warm-up,
    every line carries a timestamp, and only lines inside [t0, t1] (host clock around the timed region) are used."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.02 <= ts <= t1 + 0.05):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def _bench_model_inputs():
    from densecap_b200.synth import ModelDims, make_weights
    dims = ModelDims()
    return dims, make_weights(dims, seed=0, box_variant="A")


def cpu_oracle_images_per_sec(threads, n_images=1, num_proposals=NUM_PROPOSALS, verbose=False, weights=None, first_seed=0, keep=None):
    """The reference's CPU path (`th run_model.lua -gpu -1`) cannot run here (no Lua/Torch7); this times its
    line-by-line restatement (oracle/, torch-CPU fp32 + numpy) on the host cores.  keep: list that receives the outputs."""
    import torch
    from oracle import densecap_oracle as O
    from densecap_b200.synth import make_image
    torch.set_num_threads(threads)
    if weights is None:
        _, weights = _bench_model_inputs()
    cfg = {"rpn_nms_thresh": 0.7, "final_nms_thresh": 0.3, "num_proposals": num_proposals, "seq_length": SEQ_LEN}
    times = []
    for i in range(n_images):
        img = make_image(IMG_H, IMG_W, seed=first_seed + i)
        t0 = time.perf_counter()
        out = O.densecap_forward_test(img, weights, cfg, return_intermediates=keep is not None)
        times.append(time.perf_counter() - t0)
        if keep is not None:
            keep.append(out)
        if verbose:
            print("[cpu oracle] image seed %d: %.2f s on %d threads, %d final boxes" % (first_seed + i, times[-1], threads, out["final_boxes"].shape[0]),
                  file=sys.stderr)
    return times


_THREAD_CACHE = "/tmp/densecap_b200_cpu_threads.json"


def pick_cpu_threads(weights=None):
    """Thread count for the CPU baseline.  'All host threads' is not always the fastest setting for oneDNN/MKL on a 128-thread host,
    and a conv micro-benchmark is a noisy proxy, so the REAL workload (one full oracle image) is timed once at three thread counts
    after one warm-up image; the pick is cached for the other bench arm of the same box.  Returns (threads, sweep dict)."""
    n = os.cpu_count() or 1
    try:
        c = json.load(open(_THREAD_CACHE))
        if c.get("cpu_count") == n:
            return int(c["threads"]), c["sweep"]
    except Exception:
        pass
    cands = sorted({max(1, n // 4), max(1, n // 2), n})
    if weights is None:
        _, weights = _bench_model_inputs()
    cpu_oracle_images_per_sec(cands[-1], 1, weights=weights)                         # warm-up (oneDNN primitive caches, page-in)
    sweep = {}
    for c in cands:
        sweep[str(c)] = cpu_oracle_images_per_sec(c, 1, weights=weights)[0]
    best = int(min(sweep, key=lambda k: sweep[k]))
    try:
        json.dump({"cpu_count": n, "threads": best, "sweep": sweep}, open(_THREAD_CACHE, "w"))
    except Exception:
        pass
    return best, sweep


def config0_split(threads, weights):
    """BASELINE configs[0]: single 720x600 image, 300 proposals, CPU path -- LocalizationLayer + LSTM forward timed apart from the
    conv trunk (SURVEY.md 8d last row)."""
    import torch
    from oracle import densecap_oracle as O
    from densecap_b200.synth import make_image
    torch.set_num_threads(threads)
    img = make_image(IMG_H, IMG_W, seed=0)
    t0 = time.perf_counter()
    feats = O.vgg_trunk(img, weights)
    t1 = time.perf_counter()
    ll = O.localization_layer_forward_test(feats, weights, IMG_H, IMG_W, 0.7, 300, True)
    t2 = time.perf_counter()
    codes, obj, trans, fb = O.recog_head(ll["roi_features"], ll["boxes"], weights)
    t3 = time.perf_counter()
    O.lm_sample(codes, weights, SEQ_LEN)
    t4 = time.perf_counter()
    return {"workload": "BASELINE configs[0]: 1 image 720x600, 300 proposals, CPU path", "cores": threads,
            "conv_trunk_s": t1 - t0, "localization_layer_s": t2 - t1, "recog_head_s": t3 - t2, "lstm_decode_s": t4 - t3,
            "localization_layer_plus_lstm_s": (t2 - t1) + (t4 - t3), "proposals": int(ll["boxes"].shape[0])}


def run_reference(args, rank, world):
    if rank != 0:
        return
    _, weights = _bench_model_inputs()
    threads, sweep = pick_cpu_threads(weights)
    per = sweep[str(threads)]
    budget = 120.0
    steps = max(1, min(args.steps, int(budget / max(per, 1e-3))))
    times = cpu_oracle_images_per_sec(threads, steps, verbose=True, weights=weights)
    total = float(np.sum(times))
    value = steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
        "requested_steps": args.steps, "warmup": 1, "ms_per_step": 1000.0 * total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1 image 720x600x3 per step, random-init VGG-16 DenseCap, 1000 proposals, 15-token greedy decode",
                   "global_batch": 1, "note": "reference Lua/Torch7 CPU path is not runnable here (no th/luajit); this is its restatement "
                   "(oracle/densecap_oracle.py, torch-CPU fp32 oneDNN/MKL + numpy) on the host cores; with N > 1 GPUs this arm is still ONE "
                   "CPU process on rank 0 -- compare it with the N = 1 line only"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "host_threads_available": os.cpu_count(), "kind": "port",
                         "sample": "%d full 720x600 image(s), 1000 proposals, 15 tokens" % steps,
                         "thread_sweep_s_per_image": sweep},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "config0": config0_split(threads, weights),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def parity_block(model, dims, weights, threads, dev_image0):
    """Parity of the benched workload, on the bench's own image (seed 0) and weights: the GPU forward against the fp32 oracle and --
    because this workload's saturated objectness makes the pick set ill-conditioned -- the fp32 oracle against its fp64 gold run."""
    import torch
    from oracle import densecap_oracle as O
    from densecap_b200.synth import make_image
    torch.set_num_threads(threads)
    cfg = {"rpn_nms_thresh": 0.7, "final_nms_thresh": 0.3, "num_proposals": NUM_PROPOSALS, "seq_length": SEQ_LEN}
    img = make_image(IMG_H, IMG_W, seed=0)
    t0 = time.perf_counter()
    o32 = O.densecap_forward_test(img, weights, cfg, return_intermediates=True)
    t32 = time.perf_counter() - t0
    with O.precision(np.float64):
        o64 = O.densecap_forward_test(img, weights, cfg, return_intermediates=True)
    model.ctx.set_decode_all(True)
    r = model.forward_device(dev_image0)
    torch.cuda.synchronize()
    K = int(r["num_proposals"][0]); Kf = int(r["num_final"][0])
    g = {"idx": r["proposal_idx"][:K].cpu().numpy(), "boxes": r["all_boxes"][:K].cpu().numpy(), "scores": r["all_scores"][:K].cpu().numpy(),
         "seq": r["all_seq"][:K].cpu().numpy(), "final_idx": r["final_idx"][:Kf].cpu().numpy()}
    model.ctx.set_decode_all(False)

    def pack(o):
        return {"idx": o["ll"]["idx"], "boxes": o["pre_nms"]["final_boxes"], "scores": o["pre_nms"]["obj"][:, 0], "seq": o["pre_nms"]["seq"],
                "final_idx": o["final_idx"]}

    def rel(a, b):
        a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
        return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)) if a.size else 0.0

    def compare(a, b):
        ia = {int(v): i for i, v in enumerate(a["idx"])}
        rows = [(ia[int(v)], j) for j, v in enumerate(b["idx"]) if int(v) in ia]
        ra = np.array([x for x, _ in rows], int); rb = np.array([y for _, y in rows], int)
        fa = a["idx"][a["final_idx"] - 1]; fb = b["idx"][b["final_idx"] - 1]
        return {"rpn_picks_in_common": len(rows), "rpn_pick_order_identical": bool(np.array_equal(a["idx"], b["idx"])),
                "max_rel_box_err_common_rows": rel(a["boxes"][ra], b["boxes"][rb]), "max_rel_score_err_common_rows": rel(a["scores"][ra], b["scores"][rb]),
                "token_rows_differing_common_rows": int((a["seq"][ra] != b["seq"][rb]).any(axis=1).sum()),
                "num_final": [int(len(fa)), int(len(fb))], "final_in_common": int(len(set(fa.tolist()) & set(fb.tolist()))),
                "final_idx_identical": bool(np.array_equal(fa, fb))}

    p32, p64 = pack(o32), pack(o64)
    prob = o32["ll"]["all_prob"]
    return {"image_seed": 0, "oracle_fp32_s": t32,
            "objectness_prob_exactly_1.0": int((prob == np.float32(1.0)).sum()), "rpn_candidates": int(prob.shape[0]),
            "gpu_vs_oracle_fp32": compare(g, p32), "gpu_vs_oracle_fp64": compare(g, p64), "oracle_fp32_vs_oracle_fp64": compare(p32, p64),
            "explanation": "He-init VGG on +-128 input gives conv5 features of O(100); the reference-init RPN scores then differ by std ~20, so "
                           "thousands of the 20,520 objectness probabilities round to exactly 1.0f and the NMS visits them in tie order. Which boxes "
                           "sit in that tie set flips on last-bit noise, so even the fp32 and fp64 runs of the ORACLE pick different proposals "
                           "(oracle_fp32_vs_oracle_fp64); the GPU path agrees with the fp32 oracle at least as well. Integer stages fed the oracle's "
                           "own inputs are bit-exact, and the calibrated full-size model matches end to end (tests/test_gpu_fullsize.py)."}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("DC_BENCH_BATCH", "4")), help="images per step per GPU")
    ap.add_argument("--workload", default="cfg1", choices=["cfg1", "mixed8"],
                    help="cfg1 = BASELINE configs[1] (the metric's configuration); mixed8 = configs[2]: 8 images, long side 600..800, 300 proposals")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle legs (cpu_baseline, parity)")
    ap.add_argument("--no-extras", action="store_true", help="device-resident timing only (for runs under ncu)")
    ap.add_argument("--groups", default="", help="development: dc_set_accumulation_groups a,b,c (k-blocks per TMEM accumulator group)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from densecap_b200 import _lib
    from densecap_b200.model import DenseCapModel, _wrap_result
    from densecap_b200.synth import ModelDims, make_weights, make_image
    from densecap_b200 import parallel as PL

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload == "mixed8":
        run_mixed8(args, rank, world, local_rank, dist)
        return

    B = args.batch
    dims = ModelDims()
    weights = make_weights(dims, seed=0, box_variant="A")      # identical replica on every rank
    model = DenseCapModel({"dims": dims}, weights=weights, device=local_rank, max_batch=B, max_image_hw=(IMG_H, IMG_W))
    model.setTestArgs({"rpn_nms_thresh": 0.7, "final_nms_thresh": 0.3, "num_proposals": NUM_PROPOSALS})
    ctx = model.ctx
    lib = ctx._lib
    if args.groups:
        ctx.check(lib.dc_set_accumulation_groups(ctx.handle, *[int(v) for v in args.groups.split(",")]))
    # a pool of distinct synthetic batches (image i of rank r differs), resident in HBM for `value`,
    # and in pinned host memory for `e2e`
    NPOOL = 3
    host_batches, dev_batches = [], []
    for j in range(NPOOL):
        hb = torch.empty((B, 3, IMG_H, IMG_W), dtype=torch.float32).pin_memory()
        for b in range(B):
            hb[b] = torch.from_numpy(make_image(IMG_H, IMG_W, seed=1000 * rank + 10 * j + b)[0])
        host_batches.append(hb)
        dev_batches.append(hb.cuda())
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    res = _lib.DcResult()
    P = ctx.max_proposals
    T = dims.seq_length
    # fixed-size result record gathered across ranks each step (the path's only exchange): num_final | boxes | scores | seq
    gather_in = torch.zeros((PL.record_numel(B, P, T),), dtype=torch.float32, device="cuda") if world > 1 else None
    gather_out = torch.zeros((world * gather_in.numel(),), dtype=torch.float32, device="cuda") if world > 1 else None

    def make_steps(nb):
        rows = nb * P
        nf = np.zeros((nb,), np.int32); fb = np.zeros((rows, 4), np.float32); fs = np.zeros((rows,), np.float32)
        sq = np.zeros((rows, T), np.int64)
        hp = lambda a: a.ctypes.data_as(C.c_void_p)

        def step_device(j):
            ctx.check(lib.dc_forward_test(ctx.handle, C.c_void_p(dev_batches[j % NPOOL].data_ptr()), nb, IMG_H, IMG_W, C.byref(res), sp))
            if world > 1 and nb == B:
                r = _wrap_result(res, dims, torch.device("cuda", local_rank))
                PL.pack_result_record(r["num_final"], r["final_boxes"], r["final_scores"], r["final_seq"], out=gather_in)
                PL.gather_records(dist, gather_in, out=gather_out)

        def step_host(j):
            ctx.check(lib.dc_forward_test_host(ctx.handle, C.c_void_p(host_batches[j % NPOOL].data_ptr()), nb, IMG_H, IMG_W, hp(nf), hp(fb), hp(fs),
                                               hp(sq)))
        d2h = int(nf.nbytes + fb.nbytes + fs.nbytes + sq.nbytes)
        return step_device, step_host, d2h, (nf, fb, fs, sq)

    step_device, step_host, d2h_bytes, host_out = make_steps(B)
    nf, fb, fs, sq = host_out
    hp = lambda a: a.ctypes.data_as(C.c_void_p)

    # the same call with the decoded uint8 RGB image at the boundary (device-side preprocessing)
    u8_batches = []
    if not args.no_extras:
        for j in range(NPOOL):
            ub = torch.empty((B, IMG_H, IMG_W, 3), dtype=torch.uint8).pin_memory()
            ub.copy_(torch.from_numpy(np.random.RandomState(77 + 10 * j + rank).randint(0, 256, size=(B, IMG_H, IMG_W, 3), dtype=np.uint8)))
            u8_batches.append(ub)

    def step_host_u8(j):
        ctx.check(lib.dc_forward_test_host_u8(ctx.handle, C.c_void_p(u8_batches[j % NPOOL].data_ptr()), B, IMG_H, IMG_W, hp(nf), hp(fb), hp(fs),
                                              hp(sq)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for j in range(warm):
            fn(j)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches()
        wall[0] = time.time()
        e0.record(stream)
        for j in range(steps):
            fn(warm + j)
        e1.record(stream)
        barrier()
        wall[1] = time.time()
        ms = e0.elapsed_time(e1)
        launches = ctx.kernel_launches() - l0
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    wall = [0.0, 0.0]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(1.0)          # let nvidia-smi start streaming before the (short) timed region
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop(wall[0], wall[1]) if sampler else None
    if args.no_extras:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": B * args.steps * world / (ms_dev / 1000.0), "unit": "images/s", "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "gpu_launches": int(launches),
                              "note": "--no-extras run (profiling); not a bench line"}), flush=True)
        model.ctx.close()
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    ms_e2e, _ = timed(step_host, args.steps, 1)
    ms_e2e_u8, _ = timed(step_host_u8, args.steps, 1)

    # the reference API is batch 1 (DenseCapModel.lua:244): the same measurements at one image per call
    batch1 = None
    if B > 1:
        sd1, sh1, d2h1, _ = make_steps(1)
        n1 = max(args.steps, 20)
        ms1, l1 = timed(sd1, n1, 3)
        ms1h, _ = timed(sh1, n1, 1)
        batch1 = {"images_per_step_per_gpu": 1, "value": n1 * world / (ms1 / 1000.0), "unit": "images/s", "ms_per_step": ms1 / n1, "steps": n1,
                  "gpu_launches_per_step": l1 / n1,
                  "e2e": {"value": n1 * world / (ms1h / 1000.0), "unit": "images/s", "ms_per_step": ms1h / n1,
                          "h2d_bytes_per_step": int(3 * IMG_H * IMG_W * 4), "d2h_bytes_per_step": d2h1}}

    # per-phase device times of one step (events only; separate from the timed region)
    ctx.enable_timing(True)
    step_device(0)
    torch.cuda.synchronize()
    phases = ctx.get_times()
    ctx.enable_timing(False)
    final_counts = _wrap_result(res, dims, torch.device("cuda", local_rank))["num_final"].cpu().tolist()
    # the ROI-pooling kernel alone: in a batch > 1 step the per-image side streams overlap it with other images' kernels,
    # so its HBM figure is taken from a batch-1 forward (image 0 of the pool)
    roi_alone_ms = None
    ctx.enable_timing(True)
    ctx.check(lib.dc_forward_test(ctx.handle, C.c_void_p(dev_batches[0].data_ptr()), 1, IMG_H, IMG_W, C.byref(res), sp))
    torch.cuda.synchronize()
    phases_b1 = {}
    for n_, ms_ in ctx.get_times():
        if n_.startswith("kernel:roi"):
            roi_alone_ms = ms_
        else:
            phases_b1[n_] = phases_b1.get(n_, 0.0) + ms_
    ctx.enable_timing(False)

    if rank == 0:
        peaks = measured_peaks()
        imgs = B * args.steps * world
        value = imgs / (ms_dev / 1000.0)
        e2e_value = imgs / (ms_e2e / 1000.0)
        ph = {}
        for n, ms in phases:
            if not n.startswith("kernel:"):
                ph[n] = ph.get(n, 0.0) + ms
        trunk_ms = ph.get("conv_trunk", 0.0)
        fc6_ms = ph.get("recog:fc6", 0.0)
        roi_ms = roi_alone_ms or 0.0
        # Roofline of the dominant kernel (tc_gemm_kernel, the tcgen05 split-fp16 GEMM / implicit-GEMM conv), quoted on its
        # single largest launch, fc6 (2*M*N*K flops), timed live with CUDA events on the launching stream (phase events of
        # dc_get_times).  The conv trunk (12 launches of the same kernel family + conv1_1) and the ROI pooling (HBM) are given beside it.
        fc6_gflop = 2.0 * (B * P) * dims.fc_dim * (dims.feat_dim * 49) / 1e9
        trunk_tflops = (CONV_TRUNK_GFLOP * B / 1e3) / (trunk_ms / 1e3) if trunk_ms > 0 else None
        fc6_tflops = (fc6_gflop / 1e3) / (fc6_ms / 1e3) if fc6_ms > 0 else None
        peak = peaks["bf16_sustained"]
        # DRAM traffic of the fc6 launch from an `ncu --set full` capture AT THE BENCHED BATCH (profiles/roofline_traffic_*.json holds one
        # record per batch size; null when no capture exists for this batch)
        traffic = None
        trunk_ncu = None
        for fn in sorted([f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("roofline_traffic_")], reverse=True):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", fn)))
                rec = tj.get("fc6", {}).get("batch%d" % B)
                if trunk_ncu is None and tj.get("conv_trunk", {}).get("batch%d" % B):
                    trunk_ncu = dict(tj["conv_trunk"]["batch%d" % B], file="profiles/" + fn)
                if rec:
                    traffic = {"bytes_per_launch": int(rec["dram_bytes_read"] + rec["dram_bytes_write"]), "file": "profiles/" + fn, "batch": B,
                               "algorithmic_bytes_per_launch": int(2 * 2 * (dims.fc_dim * dims.feat_dim * 49 + B * P * dims.feat_dim * 49) + 2 * 2 * B * P * dims.fc_dim)}
                    break
            except Exception:
                continue
        roofline = {"bound": "tensor", "kernel": "tc_gemm_kernel<256,gemm,cg2> fc6: M=%d N=%d K=%d" % (B * P, dims.fc_dim, dims.feat_dim * 49),
                    "achieved": fc6_tflops, "peak": peak, "unit": "TFLOP/s", "frac": (fc6_tflops / peak) if fc6_tflops else None,
                    "traffic": (traffic["bytes_per_launch"] if traffic else None), "traffic_detail": traffic,
                    "peak_source": peaks["source"] + ", bf16 sustained (kernel timed inside a long step)",
                    "note": "achieved counts ALGORITHMIC fp32 flops; the kernel issues 3 fp16 MMAs per product (split-fp16, fp32-class accuracy), "
                            "so the tensor pipe does 3x this work (hw_flops_factor); tensor-pipe activity per kernel: profiles/ (ncu)",
                    "hw_flops_factor": 3, "tensor_pipe_frac_3x": (3 * fc6_tflops / peak) if fc6_tflops else None,
                    "conv_trunk": {"ms": trunk_ms, "algorithmic_tflops": trunk_tflops, "frac": (trunk_tflops / peak) if trunk_tflops else None,
                                   "tensor_pipe_frac_3x": (3 * trunk_tflops / peak) if trunk_tflops else None,
                                   # ncu sm__pipe_tensor_cycles_active, time-weighted over every launch of the 13 conv layers (committed capture, not this run)
                                   "tensor_pipe_active_ncu": trunk_ncu},
                    "roi_pool": {"bound": "hbm", "ms": roi_ms, "timed": "ROI-pooling kernel(s) of one image, batch-1 forward, CUDA events",
                                 "algorithmic_bytes": 103870080,
                                 "achieved_GBps": (103870080 / 1e9) / (roi_ms / 1e3) if roi_ms > 0 else None, "peak_GBps": peaks["hbm"],
                                 "frac": ((103870080 / 1e9) / (roi_ms / 1e3) / peaks["hbm"]) if roi_ms > 0 else None}}
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "720x600x3 synthetic images, random-init VGG-16 DenseCap (V=10497), 1000 proposals, 15-token greedy decode, "
                                   "final NMS 0.3 (BASELINE configs[1])",
                       "images_per_step_per_gpu": B, "global_batch": B * world, "parallelism": "image-sharded replicas, dp%d" % world,
                       "arithmetic": "fp32 semantics: split-fp16 (hi+lo) operands, 3 tcgen05 kind::f16 MMAs per product (2 wide ones on 64/128-column tiles), "
                                     "fp32 accumulate in TMEM",
                       "l2": "no explicit flush: each step streams >1 GB (fc6 weights 411 MB + activations), far beyond the 126 MB L2; "
                             "inputs rotate over %d distinct batches" % NPOOL,
                       "batch1": batch1},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(B * 3 * IMG_H * IMG_W * 4),
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps},
            "e2e_u8": {"value": imgs / (ms_e2e_u8 / 1000.0), "unit": "images/s", "h2d_bytes_per_step": int(B * 3 * IMG_H * IMG_W),
                       "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e_u8 / args.steps,
                       "note": "dc_forward_test_host_u8: decoded uint8 RGB image at the boundary, run_model.lua:66-74 preprocessing on the device"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "phases_ms": ph,
            "phases_ms_batch1": phases_b1,
            "conv_trunk_frac_of_flop_roofline": (value * CONV_TRUNK_GFLOP / 1e3) / (peak * world),
            "num_final_last_step": final_counts,
        }
        if not args.no_cpu_baseline and world == 1:
            threads, sweep = pick_cpu_threads(weights)
            t = cpu_oracle_images_per_sec(threads, 1, verbose=True, weights=weights, first_seed=1)
            line["cpu_baseline"] = {"value": 1.0 / t[0], "unit": "images/s", "cores": threads, "host_threads_available": os.cpu_count(),
                                    "kind": "port", "thread_sweep_s_per_image": sweep,
                                    "sample": "1 full image (720x600, 1000 proposals, 15 tokens) after warm-up, at the fastest of 3 thread counts, through "
                                              "oracle/densecap_oracle.py (torch-CPU fp32 + numpy restatement of the reference Lua path; Torch7 itself is "
                                              "not runnable here)"}
            try:
                line["parity"] = parity_block(model, dims, weights, threads, dev_batches[0][:1].contiguous())
            except Exception as e:                                                      # the bench line must survive a checker problem
                line["parity"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    model.ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def mixed8_sizes():
    """8 images, long side drawn from {600, 640, ..., 800}, short side = round(3/4 long), seed = index."""
    rng = np.random.default_rng(np.random.PCG64(3))
    longs = rng.choice(np.arange(600, 801, 40), size=8)
    out = []
    for i, L in enumerate(longs):
        s = int(round(0.75 * int(L)))
        out.append((s, int(L)) if i % 2 == 0 else (int(L), s))          # landscape and portrait
    return out


def run_mixed8(args, rank, world, local_rank, dist):
    """BASELINE configs[2]: batch 8 of mixed-size images, 300 proposals each (BilinearRoiPooling / NMS stress).  The reference API is one
    image per call, and a batch call needs equal sizes, so a step = 8 batch-1 forwards of different sizes on ONE context."""
    import torch
    from densecap_b200 import _lib
    from densecap_b200.model import DenseCapModel
    from densecap_b200.synth import ModelDims, make_weights, make_image
    dims = ModelDims()
    weights = make_weights(dims, seed=0, box_variant="A")
    sizes = mixed8_sizes()
    model = DenseCapModel({"dims": dims, "max_proposals": 300}, weights=weights, device=local_rank, max_batch=1, max_image_hw=(800, 800))
    model.setTestArgs({"rpn_nms_thresh": 0.7, "final_nms_thresh": 0.3, "num_proposals": 300})
    ctx = model.ctx; lib = ctx._lib
    imgs = [torch.from_numpy(make_image(h, w, seed=100 * rank + i)).cuda() for i, (h, w) in enumerate(sizes)]
    stream = torch.cuda.current_stream(); sp = C.c_void_p(stream.cuda_stream)
    res = _lib.DcResult()

    def step():
        for im, (h, w) in zip(imgs, sizes):
            ctx.check(lib.dc_forward_test(ctx.handle, C.c_void_p(im.data_ptr()), 1, h, w, C.byref(res), sp))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    l0 = ctx.kernel_launches()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    # per-image phase times
    ctx.enable_timing(True)
    per = []
    for im, (h, w) in zip(imgs, sizes):
        ctx.check(lib.dc_forward_test(ctx.handle, C.c_void_p(im.data_ptr()), 1, h, w, C.byref(res), sp))
        torch.cuda.synchronize()
        d = {}
        for n_, ms_ in ctx.get_times():
            d[n_] = d.get(n_, 0.0) + ms_
        per.append({"size": [h, w], "phases_ms": d})
    if rank == 0:
        print(json.dumps({"metric": "images/sec, batch 8 mixed 600-800 px long side, 300 proposals (BASELINE configs[2])", "value": 8 * args.steps * world / (ms / 1000.0),
                          "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "mixed8", "sizes_hw": sizes, "proposals": 300}, "gpu_launches": int(ctx.kernel_launches() - l0),
                          "per_image": per}), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
