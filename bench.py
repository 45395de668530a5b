#!/usr/bin/env python
"""bench.py — images/sec of the recognition hot path (BASELINE.json metric) on N B200s of one node.

Workload at N=1: BASELINE.json configs[1] — single-GPU training step, batch 64, 32x100 synthetic images,
target length 20, -input_feed (CNN + BiLSTM encoder + attention decoder + generator + NLL, forward and
backward, per-group clip + SGD).  N>1: the same per-GPU batch on every rank (weak scaling), gradients
summed with an NCCL all-reduce of the flat gradient buffer before the identical update on every rank.

  value : whole-job images/s with the batch already resident in HBM (CUDA events on the engine's stream)
  e2e   : the same step through the reference-facing call (Model.step / aocr_train_step) with HOST buffers:
          H2D of images+targets and D2H of the loss inside the timed region
  roofline / cpu_baseline : see DESIGN.md §8

`--impl reference` times the reference's CPU path instead: the float64 oracle restatement (the Torch7
stack cannot run here) on the box's host cores, same config/metric, bounded sample.
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

B_PER_GPU, IMG_W, TGT_T, SEED = 64, 100, 20, 910820
FLOP_PER_IMG_TRAIN = 6.716e9     # 3 x forward (W=100, S=24, T=20)
WORKLOAD = "configs[1]: train step, batch 64/GPU, 32x100 gray, target_l 20, -input_feed, max_enc 80, max_dec 50"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


def load_traffic(kind="gemm"):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)"""
    p = os.path.join(ROOT, "profiles", "r01_executor_ncu_full_final.json" if kind == "persist"
                     else "r01_tc_gemm_traffic.json")
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


def make_inputs(rank):
    from oracle import make_batch
    return make_batch(B_PER_GPU, IMG_W, TGT_T - 1, seed=SEED + rank, force_T=TGT_T, kind="noise")


def oracle_config():
    from oracle import Config
    return Config(batch_size=B_PER_GPU, max_encoder_l=80, max_decoder_l=50, input_feed=True)


def run_reference(args):
    """CPU arm: the oracle restatement of the reference's schedule, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import Oracle, init_params, init_bn_stats
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = oracle_config()
    batch = make_inputs(0)
    sample_b = 16                                   # bounded sample: 16 of the 64 images per step
    sub = {k: (v[:sample_b] if hasattr(v, "shape") else v) for k, v in batch.items()}
    orc = Oracle(cfg, init_params(cfg, SEED), init_bn_stats(cfg))
    warm = min(args.warmup, 3)
    for _ in range(warm):
        orc.train_step(sub["images"], sub["targets"], sub["targets_eval"], 0.1)
    # exactly K steps unless that would take more than ~3 minutes on this host (then as many as fit, reported)
    steps, t0, budget = 0, time.time(), 180.0
    while steps < max(1, args.steps):
        orc.train_step(sub["images"], sub["targets"], sub["targets_eval"], 0.1)
        steps += 1
        if time.time() - t0 > budget:
            break
    dt = (time.time() - t0) / steps
    ips = sample_b / dt
    line = {"impl": "reference", "metric": "train_images_per_sec", "value": ips, "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{sample_b} of {B_PER_GPU} images per step"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} train steps of {sample_b} images (float64 oracle, torch CPU primitives)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cuda_tensor_from_ptr(ptr, n, torch):
    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(w, device="cuda")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--gemm-mode", type=int, default=int(os.environ.get("AOCR_GEMM_MODE", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from aocr import Model
    from oracle import init_params, init_bn_stats   # weights only: the library holds no RNG

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)
    steps = args.steps

    cfg = oracle_config()
    model = Model(log=lambda m: None, device=local, gemm_mode=args.gemm_mode, dp_rank=rank, dp_world=world,
                  global_batch=B_PER_GPU * world if world > 1 else 0)
    model.create(dict(batch_size=B_PER_GPU, max_encoder_l=80, max_decoder_l=50, input_feed=True, learning_rate=0.1))
    model.set_parameters(init_params(cfg, SEED), init_bn_stats(cfg))
    h = model.handle
    batch = make_inputs(rank)
    lr = 0.1
    stream = torch.cuda.ExternalStream(h.stream(), device=local)
    if world > 1:
        from aocr import dist as aocr_dist
        # exchange = SyncBN statistics + 3 overlapped gradient buckets; native NCCL inside the library (graph-captured),
        # AOCR_DP_HOOK=1 selects the torch.distributed hook flavour instead
        if os.environ.get("AOCR_DP_HOOK"):
            aocr_dist.attach(h, local)
        else:
            aocr_dist.attach_native(h, local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def trace(msg):
        if os.environ.get("AOCR_BENCH_TRACE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    def step_resident():
        h.train_step_staged(lr, sync=False)     # dp: the engine calls the exchange hook at its bucket boundaries

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        h.synchronize()

    # ---- device-resident arm (value)
    h.stage_batch(batch["images"], batch["targets"], batch["targets_eval"])
    for _ in range(warmup):
        step_resident()
    trace("warmup enqueued")
    barrier()
    trace("warmup done")
    sampler = ClockSampler(local)
    sampler.start()
    l0 = h.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for i in range(steps):
        with torch.cuda.stream(stream):
            flush.zero_()                       # L2 flush between timed iterations (outside the timed span)
            ev[i][0].record(stream)
        step_resident()
        with torch.cuda.stream(stream):
            ev[i][1].record(stream)
    trace("timed steps enqueued")
    barrier()
    trace("timed steps done")
    launches = h.launch_count() - l0
    ms_resident = sum(a.elapsed_time(b) for a, b in ev) / steps
    loss = h.read_loss()

    # ---- end-to-end arm (e2e): Model.step with host buffers (pinned), H2D + D2H inside the timed region
    pin = {k: torch.from_numpy(np.ascontiguousarray(batch[k])).pin_memory().numpy()
           for k in ("images", "targets", "targets_eval")}
    hb = [pin["images"], pin["targets"], pin["targets_eval"], batch["num_nonzeros"], None]

    def step_e2e():
        return model.step(hb, False)[0]

    for _ in range(warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    trace("e2e done")
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- greedy decode (the metric's second half): greedy pass + gold pass, max_decoder_l = 50 steps each
    nd = max(3, steps // 2)
    for _ in range(2):
        h.decode_greedy_staged(sync=True)
    evd = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nd)]
    barrier()
    for i in range(nd):
        with torch.cuda.stream(stream):
            flush.zero_()
            evd[i][0].record(stream)
        h.decode_greedy_staged(sync=False)
        with torch.cuda.stream(stream):
            evd[i][1].record(stream)
    barrier()
    ms_dec = sum(a.elapsed_time(b) for a, b in evd) / nd
    t0 = time.perf_counter()
    for _ in range(nd):
        model.step(hb, True)
    barrier()
    ms_dec_e2e = (time.perf_counter() - t0) * 1e3 / nd
    trace("decode done")

    # ---- per-kernel-class timing for the roofline: CUDA events recorded on the engine stream around every call of
    # the class during extra (untimed) steps; no host sync inside the step
    peaks = load_peaks()
    h.stage_batch(batch["images"], batch["targets"], batch["targets_eval"])
    h.prof_enable(True)
    nprof = 3
    for _ in range(nprof):
        step_resident()
    h.synchronize()
    prof = [h.prof_read(c) for c in range(4)]
    h.prof_enable(False)

    t = torch.tensor([ms_resident, ms_e2e, ms_dec, ms_dec_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e, ms_dec, ms_dec_e2e = (float(x) for x in t)
    if rank == 0:
        total_imgs = B_PER_GPU * world
        value = total_imgs / (ms_resident / 1e3)
        gemm_ms, gemm_n, gemm_flops = prof[0]
        att_ms, att_n, att_bytes = prof[1]
        rec_ms, rec_n, rec_flops = prof[2]

        def tens(ms, flops, n, name):
            ach = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peaks["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                    "peak_source": peaks["src"] + " (sustained bf16)", "launches_per_step": n / nprof,
                    "ms_per_step_in_class": ms / nprof, "share_of_step": (ms / nprof) / ms_resident}
        conv = tens(gemm_ms, gemm_flops, gemm_n,
                    "tc_gemm_kernel (tcgen05 GEMM / implicit-GEMM conv fwd+dgrad+wgrad, incl. operand conversion)")
        recur = tens(rec_ms, rec_flops, rec_n,
                     "persist_kernel (persistent recurrence executor: decoder fwd/bwd + encoder directions, "
                     "tcgen05 GEMM tiles + fused cell/attention bodies + grid barriers; lanes overlap, shares can sum > 1)")
        if rec_ms >= gemm_ms:
            # The executor is the dominant kernel.  What bounds it is streaming the recurrent weights: every timestep
            # re-reads all weight planes of the step (77 MB for the decoder forward, more than L2 retains), and its GEMM
            # commands run at the per-SM operand ingest limit - so the roofline is bytes against HBM.
            # achieved = operand bytes its GEMM commands read (each weight / activation plane element once per command)
            # / CUDA-event time of the executor launches.  The tensor view (MACs once / bf16 peak) is kept beside it.
            sb_ms, sb_n, sb_bytes = prof[3]
            ach = sb_bytes / (sb_ms * 1e-3) / 1e9 if sb_ms > 0 else 0.0
            roof = {"bound": "hbm", "kernel": recur["kernel"], "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": ach / peaks["hbm"], "peak_source": peaks["src"] + " (HBM copy bandwidth)",
                    "launches_per_step": sb_n / nprof, "ms_per_step_in_class": sb_ms / nprof,
                    "share_of_step": (sb_ms / nprof) / ms_resident,
                    "algorithmic_bytes_per_launch": sb_bytes / max(sb_n, 1),
                    "tensor_view": {k: recur[k] for k in ("achieved", "peak", "unit", "frac")}}
        else:
            roof = dict(conv)
        roof["traffic"] = load_traffic("persist" if rec_ms >= gemm_ms else "gemm")
        roof["other_class"] = conv if rec_ms >= gemm_ms else recur
        roof["note"] = ("tensor figures count every MAC once; the default bf16x3 mode issues 3 MMAs per MAC (DESIGN.md 4), "
                        "and the per-timestep GEMMs have N = batch = 64: operand-streaming / latency-bound, not tensor-bound")
        if att_ms > 0:
            roof["attention_step"] = {"bound": "hbm", "achieved": att_bytes / (att_ms * 1e-3) / 1e9, "peak": peaks["hbm"],
                                      "unit": "GB/s", "frac": att_bytes / (att_ms * 1e-3) / 1e9 / peaks["hbm"],
                                      "launches_per_step": att_n / nprof, "ms_per_step_in_class": att_ms / nprof}
        h2d = int(pin["images"].nbytes + pin["targets"].nbytes + pin["targets_eval"].nbytes)
        line = {"metric": "train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {0: "bf16x3 (fp32-grade split operands, fp32 accumulate)",
                                               1: "bf16", 2: "f32"}[args.gemm_mode],
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": total_imgs, "parallelism": f"dp{world}",
                           "l2": "flushed (256 MB memset) between timed iterations", "loss_sum_last_step": loss,
                           "algorithmic_flop_per_image": FLOP_PER_IMG_TRAIN},
                "clocks": sampler.summary(),
                "e2e": {"value": total_imgs / (ms_e2e / 1e3), "unit": "images/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8},
                "gpu_launches": int(launches),
                "decode": {"metric": "greedy_decode_images_per_sec", "value": total_imgs / (ms_dec / 1e3),
                           "unit": "images/s", "ms_per_batch": ms_dec,
                           "workload": "greedy decode + gold pass, batch 64/GPU, 32x100, 50 decoder steps over 2 x 64 rows (dual pass)",
                           "e2e": {"value": total_imgs / (ms_dec_e2e / 1e3), "unit": "images/s",
                                   "ms_per_batch": ms_dec_e2e}},
                "roofline": roof,
                "step_algorithmic_flops_frac_of_bf16_peak": value * FLOP_PER_IMG_TRAIN / 1e12 / world / peaks["bf16_sustained"]}
        if not args.no_cpu_baseline and world >= 1:
            from oracle import Oracle
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            sb = 16
            orc = Oracle(cfg, init_params(cfg, SEED), init_bn_stats(cfg))
            t0 = time.time()
            nrep = 3
            for _ in range(nrep):
                orc.train_step(batch["images"][:sb], batch["targets"][:sb], batch["targets_eval"][:sb], lr)
            dt = (time.time() - t0) / nrep
            line["cpu_baseline"] = {"value": sb / dt, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"{nrep} train steps of {sb} of the {B_PER_GPU} images "
                                              "(float64 oracle, torch CPU primitives, reference schedule)"}
        print(json.dumps(line), flush=True)
    model.shutdown()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
