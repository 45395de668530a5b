#!/usr/bin/env python
"""bench.py — images/sec of the recognition hot path (BASELINE.json metric) on N B200s of one node.

Default workload (`--config 2`, what the driver runs): BASELINE.json configs[1] — single-GPU training step, batch 64,
32x100 synthetic images, target length 20, -input_feed (CNN + BiLSTM encoder + attention decoder + generator + NLL,
forward and backward, per-group clip + SGD).  N>1: the same per-GPU batch on every rank (weak scaling), gradients
summed with NCCL all-reduces of the flat gradient buffer, overlapped with backward, before the identical update.
`--config 3|4|5` select the other BASELINE configs (bucketed decode of batch 256 at widths 100-400; the 256/GPU
data-parallel training step; the 32x800 / target 150 long-sequence stress, global batch 128 split over the ranks).

  value : whole-job images/s with the batch already resident in HBM (CUDA events on the engine's stream)
  e2e   : the same step through the reference-facing call (Model.step) with HOST buffers: H2D of images+targets and
          D2H of the loss (train) or labels/scores (decode) inside the timed region
  decode: the decode half of the metric (greedy pass of max_decoder_l steps + gold pass) with its own roofline block
  roofline / cpu_baseline : see DESIGN.md §8

`--impl reference` times the reference's CPU path instead: the float64 oracle restatement (the Torch7 stack cannot run
here) on the box's host cores, the same config / batch / metric, all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-attention-ocr_b200")]

import numpy as np  # noqa: E402

SEED = 910820
# B = images per GPU (config 5: the GLOBAL batch, split over the ranks); flop = algorithmic FLOP per image
CONFIGS = {
    2: dict(workload="configs[1]: train step, batch 64/GPU, 32x100 gray, target_l 20, -input_feed, max_enc 80, max_dec 50",
            B=64, widths=[100], T=20, max_enc=80, max_dec=50, primary="train", scaling="weak",
            flop_train=6.716e9, flop_decode=5.454e9),
    3: dict(workload="configs[2]: bucketed greedy decode, batch 256/GPU, 32xW gray, W in 100..400 (one bucket per width), "
                     "-input_feed, max_enc 99, max_dec 50",
            B=256, widths=[100, 132, 164, 196, 228, 260, 292, 324, 356, 400], T=20, max_enc=99, max_dec=50,
            primary="decode", scaling="weak", flop_train=None, flop_decode=7.6e9),
    4: dict(workload="configs[3]: data-parallel train step, batch 256/GPU (global 2048 at 8 GPUs), 32x100 gray, target_l 20, "
                     "-input_feed, max_enc 80, max_dec 50",
            B=256, widths=[100], T=20, max_enc=80, max_dec=50, primary="train", scaling="weak",
            flop_train=6.716e9, flop_decode=5.454e9),
    5: dict(workload="configs[4]: long-sequence stress, global batch 128 split over the GPUs, 32x800 gray, target_l 150, "
                     "-input_feed, max_enc 199, max_dec 150",
            B=128, widths=[800], T=150, max_enc=199, max_dec=150, primary="train", scaling="strong",
            flop_train=53.07e9, flop_decode=23.82e9),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


def load_traffic(kind):
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the SHIPPED launch
    mode (profiles/r02_*; the round-1 captures needed AOCR_CLUSTER=1)"""
    names = {"persist": ["r02_executor_ncu_full.json", "r01_executor_ncu_full_final.json"],
             "gemm": ["r02_tc_gemm_ncu_full.json", "r01_tc_gemm_traffic.json"]}[kind]
    for n in names:
        p = os.path.join(ROOT, "profiles", n)
        if os.path.exists(p):
            return json.load(open(p)).get("dram_bytes_per_launch")
    return None


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def per_gpu_batch(c, world):
    return c["B"] // world if c["scaling"] == "strong" else c["B"]


def make_inputs(c, rank, world):
    """one synthetic batch per width bucket of the config (all rows of a batch share W, as in the reference)"""
    # aocr/data.py is plain numpy; it is loaded as a stand-alone module so that the reference arm never imports the
    # package (and with it the ctypes binding of libaocr.so)
    import importlib.util
    spec = importlib.util.spec_from_file_location("aocr_data_standalone", os.path.join(ROOT, "torch-attention-ocr_b200", "aocr", "data.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    synthetic_batch = mod.synthetic_batch
    B = per_gpu_batch(c, world)
    return [synthetic_batch(B, W, c["T"] - 1, seed=SEED + rank + 1000 * i, force_T=c["T"]) for i, W in enumerate(c["widths"])]


def oracle_config(c, B):
    from oracle import Config
    return Config(batch_size=B, max_encoder_l=c["max_enc"], max_decoder_l=c["max_dec"], input_feed=True)


def config_block(c, world, extra=None):
    B = per_gpu_batch(c, world)
    d = {"workload": c["workload"], "global_batch": B * world, "parallelism": f"dp{world}",
         "l2": "flushed (256 MB memset) between timed iterations"}
    if extra:
        d.update(extra)
    return d


def run_reference(args):
    """CPU arm: the oracle restatement of the reference's schedule on the box's host cores, all threads, the SAME batch
    (full per-GPU batch of the config), the same number of warm-up and timed steps as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import Oracle, init_params, init_bn_stats
    c = CONFIGS[args.config]
    world = args.gpus
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = per_gpu_batch(c, world)
    # configs 3-5 at full size take minutes per CPU step: a bounded sample of the batch (flagged), config 2 runs whole
    sample_b = B if args.config == 2 else min(B, 16)
    cfg = oracle_config(c, sample_b)
    batches = make_inputs(c, 0, world)
    sub = [{k: (v[:sample_b] if hasattr(v, "shape") else v) for k, v in b.items()} for b in batches]
    orc = Oracle(cfg, init_params(cfg, SEED), init_bn_stats(cfg))
    train = c["primary"] == "train"

    def step(i):
        b = sub[i % len(sub)]
        if train:
            orc.train_step(b["images"], b["targets"], b["targets_eval"], 0.1)
        else:
            orc.decode_greedy(b["images"], b["targets"], b["targets_eval"])

    warm = args.warmup if args.config == 2 else min(args.warmup, 1)   # exactly the requested warm-up on the full-size config
    for i in range(warm):
        step(i)
    # exactly K steps unless that would take more than ~4 minutes on this host (then as many as fit, reported)
    steps, t0, budget = 0, time.time(), 240.0
    while steps < max(1, args.steps):
        step(steps)
        steps += 1
        if time.time() - t0 > budget:
            break
    dt = (time.time() - t0) / steps
    ips = sample_b / dt
    what = "train steps" if train else "greedy+gold decodes"
    sample = (f"{steps} {what} of the full {B}-image batch" if sample_b == B else
              f"{steps} {what} of {sample_b} of the {B} images per step (rate extrapolated to the full batch)")
    line = {"impl": "reference", "metric": "train_images_per_sec" if train else "greedy_decode_images_per_sec",
            "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(c, world),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": sample + " (float64 oracle, torch CPU primitives, reference schedule)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=int(os.environ.get("AOCR_BENCH_CONFIG", "2")), choices=sorted(CONFIGS))
    ap.add_argument("--gemm-mode", type=int, default=int(os.environ.get("AOCR_GEMM_MODE", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from aocr import Model

    c = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)
    steps = args.steps
    B = per_gpu_batch(c, world)
    train_primary = c["primary"] == "train"

    model = Model(log=lambda m: None, device=local, gemm_mode=args.gemm_mode, dp_rank=rank, dp_world=world,
                  global_batch=B * world if world > 1 else 0)
    # Model.create draws the initial weights (Torch7 reset() distributions) inside the library: same seed on every rank
    model.create(dict(batch_size=B, max_encoder_l=c["max_enc"], max_decoder_l=c["max_dec"], input_feed=True,
                      learning_rate=0.1, seed=SEED))
    h = model.handle
    batches = make_inputs(c, rank, world)
    nb = len(batches)
    lr = 0.1
    stream = torch.cuda.ExternalStream(h.stream(), device=local)
    if world > 1 and train_primary:
        from aocr import dist as aocr_dist
        # exchange = SyncBN statistics + overlapped gradient buckets; native NCCL inside the library (graph-captured),
        # AOCR_DP_HOOK=1 selects the torch.distributed hook flavour instead
        if os.environ.get("AOCR_DP_HOOK"):
            aocr_dist.attach(h, local)
        else:
            aocr_dist.attach_native(h, local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        h.synchronize()

    def stage(i):
        b = batches[i % nb]
        h.stage_batch(b["images"], b["targets"], b["targets_eval"])

    def timed_resident(fn, n):
        """n device-timed runs of fn (enqueue only) with the batch resident; L2 flushed before each; returns avg ms"""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        barrier()
        for i in range(n):
            if nb > 1:
                stage(i)                        # next width bucket (host->device copy outside the timed span)
            with torch.cuda.stream(stream):
                flush.zero_()
                ev[i][0].record(stream)
            fn()
            with torch.cuda.stream(stream):
                ev[i][1].record(stream)
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev) / n

    pins = []
    for b in batches:
        pin = {k: torch.from_numpy(np.ascontiguousarray(b[k])).pin_memory().numpy() for k in ("images", "targets", "targets_eval")}
        pins.append([pin["images"], pin["targets"], pin["targets_eval"], b["num_nonzeros"], None])

    def timed_e2e(forward_only, n):
        for i in range(min(warmup, 3)):
            model.step(pins[i % nb], forward_only)
        barrier()
        t0 = time.perf_counter()
        for i in range(n):
            model.step(pins[i % nb], forward_only)
        barrier()
        return (time.perf_counter() - t0) * 1e3 / n

    sampler = ClockSampler(local)
    ms_train = ms_train_e2e = ms_dec = ms_dec_e2e = float("nan")
    launches = 0
    loss = None
    # ---- training step (resident + e2e)
    if c["flop_train"] is not None:
        stage(0)
        for i in range(warmup):
            if nb > 1:
                stage(i)
            h.train_step_staged(lr, sync=False)
        barrier()
        if train_primary:
            sampler.start()
        l0 = h.launch_count()
        ms_train = timed_resident(lambda: h.train_step_staged(lr, sync=False), steps)
        launches = h.launch_count() - l0
        loss = h.read_loss()
        ms_train_e2e = timed_e2e(False, steps)
        if train_primary:
            sampler.stop_flag = True
            sampler.join(timeout=2)
    # ---- greedy decode (greedy pass of max_decoder_l steps + gold pass over the batch's target length)
    run_decode = c["flop_decode"] is not None and (world == 1 or not train_primary or args.config == 2)
    if run_decode:
        nd = steps if not train_primary else max(3, steps // 2)
        stage(0)
        # every width bucket is visited twice before timing: eager first (builds its programs), captured on the second visit
        for i in range(max(min(warmup, 3), 2 * nb if nb > 1 else 0)):
            if nb > 1:
                stage(i)
            h.decode_greedy_staged(sync=True)
        if not train_primary:
            sampler.start()
        l0 = h.launch_count()
        ms_dec = timed_resident(lambda: h.decode_greedy_staged(sync=False), nd)
        if not train_primary:
            launches = h.launch_count() - l0
        ms_dec_e2e = timed_e2e(True, nd)
        if not train_primary:
            sampler.stop_flag = True
            sampler.join(timeout=2)

    # ---- per-kernel-class timing for the roofline: CUDA events recorded on the engine stream around every call of
    # the class during extra (untimed) steps; no host sync inside the step
    peaks = load_peaks()
    nprof = 3

    def profile(fn):
        stage(0)
        h.prof_enable(True)
        for _ in range(nprof):
            fn()
        h.synchronize()
        prof = [h.prof_read(k) for k in range(4)]
        h.prof_enable(False)
        return prof

    prof_train = profile(lambda: h.train_step_staged(lr, sync=False)) if c["flop_train"] is not None else None
    prof_dec = profile(lambda: h.decode_greedy_staged(sync=True)) if run_decode else None

    t = torch.tensor([ms_train, ms_train_e2e, ms_dec, ms_dec_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_train, ms_train_e2e, ms_dec, ms_dec_e2e = (float(x) for x in t)
    if rank == 0:
        total_imgs = B * world

        def tens(ms, flops, n, name, ms_step):
            ach = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peaks["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                    "peak_source": peaks["src"] + " (sustained bf16)", "launches_per_step": n / nprof,
                    "ms_per_step_in_class": ms / nprof, "share_of_step": (ms / nprof) / ms_step}

        def roofline(prof, ms_step):
            gemm_ms, gemm_n, gemm_flops = prof[0]
            att_ms, att_n, att_bytes = prof[1]
            rec_ms, rec_n, rec_flops = prof[2]
            conv = tens(gemm_ms, gemm_flops, gemm_n,
                        "tc_gemm_kernel (tcgen05 GEMM / implicit-GEMM conv fwd+dgrad+wgrad, incl. operand conversion)", ms_step)
            recur = tens(rec_ms, rec_flops, rec_n,
                         "persist_kernel (persistent recurrence executor: decoder fwd/bwd + encoder directions, tcgen05 GEMM "
                         "tiles + fused cell/attention bodies + grid barriers; lanes overlap, shares can sum > 1)", ms_step)
            if rec_ms >= gemm_ms:
                # The executor is the dominant kernel.  What bounds it is streaming the recurrent weights: every timestep
                # re-reads all weight planes of the step, and its GEMM commands run at the per-SM operand ingest limit
                #  - so the roofline is bytes against HBM.  achieved = operand bytes its GEMM commands read
                # (each weight / activation plane element once per command) / CUDA-event time of the executor launches.
                sb_ms, sb_n, sb_bytes = prof[3]
                ach = sb_bytes / (sb_ms * 1e-3) / 1e9 if sb_ms > 0 else 0.0
                roof = {"bound": "hbm", "kernel": recur["kernel"], "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": ach / peaks["hbm"], "peak_source": peaks["src"] + " (HBM copy bandwidth)",
                        "launches_per_step": sb_n / nprof, "ms_per_step_in_class": sb_ms / nprof,
                        "share_of_step": (sb_ms / nprof) / ms_step,
                        "algorithmic_bytes_per_launch": sb_bytes / max(sb_n, 1),
                        "tensor_view": {k: recur[k] for k in ("achieved", "peak", "unit", "frac")}}
            else:
                roof = dict(conv)
            roof["traffic"] = load_traffic("persist" if rec_ms >= gemm_ms else "gemm")
            roof["other_class"] = conv if rec_ms >= gemm_ms else recur
            roof["note"] = ("tensor figures count every MAC once; the default bf16x3 mode issues 3 MMAs per MAC (DESIGN.md 4), "
                            "and the per-timestep GEMMs have N = batch: operand-streaming / latency-bound, not tensor-bound")
            if att_ms > 0:
                roof["attention_step"] = {"bound": "hbm", "achieved": att_bytes / (att_ms * 1e-3) / 1e9, "peak": peaks["hbm"],
                                          "unit": "GB/s", "frac": att_bytes / (att_ms * 1e-3) / 1e9 / peaks["hbm"],
                                          "launches_per_step": att_n / nprof, "ms_per_step_in_class": att_ms / nprof}
            return roof

        def h2d(pin):
            return int(pin[0].nbytes + pin[1].nbytes + pin[2].nbytes)
        h2d_avg = int(sum(h2d(p) for p in pins) / nb)
        dec_d2h = int(B * c["max_dec"] * 4 + B * 8 * 2 + 8 + 4 + c["max_dec"] * B * 4)   # labels, scores, loss, count, row losses
        train_blk = dec_blk = None
        if c["flop_train"] is not None:
            train_blk = {"metric": "train_images_per_sec", "value": total_imgs / (ms_train / 1e3), "unit": "images/s",
                         "ms_per_step": ms_train,
                         "e2e": {"value": total_imgs / (ms_train_e2e / 1e3), "unit": "images/s", "ms_per_step": ms_train_e2e,
                                 "h2d_bytes_per_step": h2d_avg, "d2h_bytes_per_step": 8},
                         "roofline": roofline(prof_train, ms_train),
                         "step_algorithmic_flops_frac_of_bf16_peak":
                             total_imgs / (ms_train / 1e3) * c["flop_train"] / 1e12 / world / peaks["bf16_sustained"]}
        if run_decode:
            dec_blk = {"metric": "greedy_decode_images_per_sec", "value": total_imgs / (ms_dec / 1e3), "unit": "images/s",
                       "ms_per_batch": ms_dec,
                       "workload": f"greedy decode ({c['max_dec']} steps) + gold pass (the batch's target length, {c['T']} steps: "
                                   f"the padded steps after it carry no loss / score), batch {B}/GPU",
                       "e2e": {"value": total_imgs / (ms_dec_e2e / 1e3), "unit": "images/s", "ms_per_batch": ms_dec_e2e,
                               "h2d_bytes_per_step": h2d_avg, "d2h_bytes_per_step": dec_d2h},
                       "roofline": roofline(prof_dec, ms_dec),
                       "step_algorithmic_flops_frac_of_bf16_peak":
                           total_imgs / (ms_dec / 1e3) * c["flop_decode"] / 1e12 / world / peaks["bf16_sustained"]}
        main_blk = train_blk if train_primary else dec_blk
        line = {"metric": main_blk["metric"], "value": main_blk["value"], "unit": "images/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_train if train_primary else ms_dec, "higher_is_better": True,
                "scaling": c["scaling"], "vs_baseline": None,
                "dtype": {0: "bf16x3 (fp32-grade split operands, fp32 accumulate)", 1: "bf16", 2: "f32"}[args.gemm_mode],
                "data": "synthetic (random-init weights drawn by Model.create)",
                "config": config_block(c, world, {"loss_sum_last_step": loss, "bench_config": args.config,
                                                  "algorithmic_flop_per_image": c["flop_train"] if train_primary else c["flop_decode"]}),
                "clocks": sampler.summary(),
                "e2e": main_blk["e2e"],
                "gpu_launches": int(launches),
                "roofline": main_blk["roofline"],
                "step_algorithmic_flops_frac_of_bf16_peak": main_blk["step_algorithmic_flops_frac_of_bf16_peak"]}
        if train_primary and dec_blk is not None:
            line["decode"] = dec_blk
        if not train_primary and train_blk is not None:
            line["train"] = train_blk
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(c, B, batches, train_primary, run_decode)
        print(json.dumps(line), flush=True)
    model.shutdown()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(c, B, batches, train_primary, with_decode):
    """the float64 oracle on this box's host cores, a bounded sample of the same workload (about 10-30 s)"""
    import torch
    from oracle import Oracle, init_params, init_bn_stats
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full = c["widths"] == [100] and B <= 64
    sb = B if full else min(B, 8)
    cfg = oracle_config(c, sb)
    b = {k: (v[:sb] if hasattr(v, "shape") else v) for k, v in batches[0].items()}
    orc = Oracle(cfg, init_params(cfg, SEED), init_bn_stats(cfg))
    out = {"unit": "images/s", "cores": cores, "kind": "port"}
    t_dec = None
    if with_decode:
        t0 = time.time()
        orc.decode_greedy(b["images"], b["targets"], b["targets_eval"])
        t_dec = time.time() - t0
    t_train, nrep = None, 0
    if c["flop_train"] is not None and (train_primary or full):
        nrep = 3 if full else 1
        t0 = time.time()
        for _ in range(nrep):
            orc.train_step(b["images"], b["targets"], b["targets_eval"], 0.1)
        t_train = (time.time() - t0) / nrep
    frac = f"the full {B}-image batch" if sb == B else f"{sb} of the {B} images (width {c['widths'][0]}; rate extrapolated)"
    if train_primary:
        out["value"] = sb / t_train
        out["sample"] = f"{nrep} train steps of {frac} (float64 oracle, torch CPU primitives, reference schedule)"
        if t_dec is not None:
            out["decode"] = {"value": sb / t_dec, "unit": "images/s",
                             "sample": f"1 greedy+gold decode of {frac}"}
    else:
        out["value"] = sb / t_dec
        out["sample"] = f"1 greedy+gold decode of {frac} (float64 oracle, torch CPU primitives, reference schedule)"
    return out


if __name__ == "__main__":
    main()
