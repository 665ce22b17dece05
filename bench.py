#!/usr/bin/env python
"""bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): VeloxSeg AutoPET-II training step, 4 patches (2 x 96^3) per GPU per step
(train_config_bs4.json: batch_size 2 x num_samples 2), full loss (CE + Dice over 4 deep outputs, 0.5 MSE
reconstruction, 2.0 SDKT), backward, AdamW.  Synthetic inputs, seeded He-init weights.  One rank per GPU; weak
scaling (every rank steps its own 4 patches; the flat 9 MB gradient is all-reduced by NCCL inside the captured step graph).

value   patches/s with the batch resident in HBM (CUDA events per step, L2 flushed between steps, max over ranks)
e2e     the same step through veloxseg_b200.train.TrainStep.step with pinned HOST batches: H2D copy of inputs and
        labels and the D2H read of the loss are inside the timed region
roofline  the kernel with the largest share of the step, timed live by the library's event profiler
cpu_baseline / --impl reference   the oracle port of the reference's CPU path (torch fp32 on the host cores)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

CFG_NAME = "autopetii"      # --workload: autopetii = BASELINE configs[1] (default, the headline), brats2021 = configs[2]
PATCHES = int(os.environ.get("VX_PATCHES", "4"))      # 4 = the reference's train_config_bs4 step; other values are scaling probes only
CFG_TITLE = {"autopetii": "AutoPET-II", "brats2021": "BraTS2021", "hecktor2022": "Hecktor2022"}


def workload_name(cfg_name):
    from veloxseg_b200.configs import MODEL_CONFIGS
    cfg = MODEL_CONFIGS[cfg_name]
    return ("VeloxSeg %s train step (models_config_%s + train_config_bs4): %d patches of %dx%s per GPU, CE+Dice x4 deep, "
            "0.5 MSE recon, 2.0 SDKT, AdamW" % (CFG_TITLE[cfg_name], cfg_name, PATCHES, sum(cfg["in_ch"]),
                                                 "96^3" if cfg["input_size"] == [96, 96, 96] else "x".join(map(str, cfg["input_size"]))))
METRIC = "train patches/s"
L2_FLUSH_BYTES = 256 << 20


def synth_batch(cfg, B, seed):
    """randn inputs (speed_test.py style) + labels from seeded ellipsoids (fg a few %)."""
    g = torch.Generator().manual_seed(seed)
    size = cfg["input_size"]
    x = torch.randn(B, sum(cfg["in_ch"]), *size, generator=g)
    zz, yy, xx = torch.meshgrid(*[torch.arange(s, dtype=torch.float32) for s in size], indexing="ij")
    y = torch.zeros(B, 1, *size, dtype=torch.int64)
    for b in range(B):
        for _ in range(3):
            c = torch.rand(3, generator=g) * torch.tensor(size, dtype=torch.float32)
            r = 6 + torch.rand(3, generator=g) * 10
            cls = int(torch.randint(1, cfg["n_classes"], (1,), generator=g))
            y[b, 0][((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2 < 1] = cls
    return x, y


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "measured"
    return 6650.0, 1590.0, "fallback"


def run_ours(args, rank, world, local_rank):
    from veloxseg_b200 import _lib
    from veloxseg_b200.configs import MODEL_CONFIGS, TRAIN
    from veloxseg_b200.nn import VeloxSeg
    from veloxseg_b200.train import TrainStep

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # No library (cuDNN / cuBLAS) kernel is on the model path any more: every convolution is libveloxseg's, fp32-accurate
    # (fp32 SIMT or tcgen05 3xTF32).  TF32 is switched off for whatever torch op remains around the step.
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = os.environ.get("VX_CUDNN_BENCHMARK", "1") == "1"
    cfg = MODEL_CONFIGS[args.workload]
    torch.manual_seed(12345)
    model = VeloxSeg(**cfg)
    ts = TrainStep(model, len(cfg["in_ch"]), dev, lr=TRAIN["lr"], weight_decay=TRAIN["weight_decay"],
                   deep_weights=TRAIN["deep_Loss_weight"], rc_weight=TRAIN["RC_Loss_weight"],
                   feature_weight=TRAIN["Feature_Loss_weight"],
                   use_graph=None if os.environ.get("VX_GRAPH", "1") == "1" else False)
    lib = _lib.get_lib()
    pw_tc = os.environ.get("VX_PW_TC", "1") == "1"
    lib.set_option(1, int(pw_tc))        # VX_OPT_PW_TENSOR_CORES (A/B switch; default on)
    x_h, y_h = synth_batch(cfg, PATCHES, 1000 + rank)
    x_h, y_h = x_h.pin_memory(), y_h.pin_memory()
    x_d, y_d = x_h.to(dev), y_h.to(dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        ts.step(x_d, y_d)
    barrier()
    # ---- device-resident timing
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    n0 = lib.c.vx_launch_count()
    with ClockSampler(local_rank) as clk:
        barrier()
        for a, b in ev:
            flush.zero_()
            a.record()
            ts.step(x_d, y_d)
            b.record()
        barrier()
    launches = int(lib.c.vx_launch_count() - n0) + ts.graph_launches * args.steps * int(ts.use_graph)
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    if os.environ.get("VX_NCU") == "1":        # ncu --profile-from-start off: capture exactly one steady-state step
        torch.cuda.profiler.start()
        ts._step_eager(x_d, y_d)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    # ---- host-side enqueue time of one step (queue empty at the start; the GPU is still busy when step() returns)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts.step(x_d, y_d)
    enqueue_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    # ---- end to end (host batches): every step copies ITS batch from pinned host memory (labels as uint8) and reads its
    # loss back; the copy of the next batch is issued while the current step runs (two distinct host batches alternate)
    xb = [x_h, x_h.clone().pin_memory()]
    yb = [y_h.to(torch.uint8).pin_memory(), y_h.to(torch.uint8).pin_memory()]
    for i in range(2):
        ts.step(xb[i % 2], yb[i % 2], sync=True, prefetch=(xb[(i + 1) % 2], yb[(i + 1) % 2]))
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        last_loss = ts.step(xb[i % 2], yb[i % 2], sync=True, prefetch=(xb[(i + 1) % 2], yb[(i + 1) % 2]))
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    # ---- per-kernel profile of one step (separate pass, not the timed one)
    # every rank runs the same steps (they contain the gradient all-reduces); only rank 0 switches the profiler on
    roof, table, fp32_tflops, hbm, tensor_peak, how = None, [], 0.0, 0.0, 0.0, ""
    nprof = 3
    # The per-kernel pass runs the step SERIALISED on one stream (no forked branches in the model, weight gradients on the main
    # stream): an event pair then brackets exactly one kernel with the device to itself.  With the forked streams of the timed
    # step the intervals also contain the time a kernel waits for SMs held by the other streams' kernels (the pw_tc_kernel
    # average read 25.6 us there against 15.1 us in the ncu launch list, profiles/r3f_*).
    model.parallel_branches = False
    model.encoder.pipeline_branches = False
    lib.set_option(10, 0)               # VX_OPT_SIDE_WGRAD off
    if rank == 0:
        lib.profile(True)
    for _ in range(nprof):
        # Gate the stream with a ~100 ms spin kernel so the host enqueues the whole eager step while the GPU is still busy:
        # the per-launch event pairs then time kernels back to back on the device (no host launch gaps in the intervals).
        torch.cuda._sleep(200_000_000)
        ts._step_eager(x_d, y_d)       # eager: the event profiler brackets individual launches
        torch.cuda.synchronize()
    lib.set_option(10, 1)
    model.parallel_branches = True
    model.encoder.pipeline_branches = True
    if rank == 0:
        rows = lib.profile_report()          # (scope, kernel, launches, total_ms, algorithmic_bytes, algorithmic_flops)
        tl = lib.profile_timeline()
        # ---- calibration of the instrument: an event pair around an EMPTY kernel enqueued behind the same kind of spin gate
        # measures what the profiler adds to every launch (event records + the launch itself); it is subtracted below
        lib.profile(True)
        torch.cuda._sleep(50_000_000)
        st_ = torch.cuda.current_stream().cuda_stream
        for _ in range(200):
            lib.c.vx_microbench(1, 0, None, st_)
        torch.cuda.synchronize()
        null_rows = [r for r in lib.profile_report() if "null_kernel" in r[1]]
        ev_overhead_us = 1e3 * sum(r[3] for r in null_rows) / max(1, sum(r[2] for r in null_rows))
        lib.profile(False)
        # ---- fp32 FMA peak of this device (SURVEY.md section 8d "derive/measure on the box")
        scratch = torch.zeros(4, device=dev)
        fma_flops = 2.0 * 8 * 32 * 2000 * 148 * 8 * 256
        best = 1e9
        for i in range(6):
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            lib.c.vx_microbench(0, 2000, scratch.data_ptr(), st_)
            b_.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, a_.elapsed_time(b_))
        fp32_tflops = fma_flops / (best * 1e-3) / 1e12
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "timeline.json"), "w") as f:      # tools/timeline_digest.py reads it
            json.dump(tl, f)
        rows.sort(key=lambda r: -r[3])
        ours_ms = sum(r[3] for r in rows) / nprof
        n_launch = sum(r[2] for r in rows) / nprof
        with open(os.path.join(ROOT, "gpurun_out", "kernel_table.json"), "w") as f:
            json.dump({"steps": nprof, "own_kernel_ms_per_step": ours_ms, "event_overhead_us_per_launch": ev_overhead_us,
                       "fp32_fma_tflops_measured": fp32_tflops,
                       "rows": [dict(scope=r[0], kernel=r[1], launches=r[2], ms=r[3], alg_bytes=r[4], alg_flops=r[5]) for r in rows]}, f, indent=1)
        hbm, tensor_peak, how = peaks()
        tf32_peak = tensor_peak / 2.0          # dense tf32 = half the measured bf16 rate (nominal 1.1 vs 2.25 PFLOP/s)
        # the dominant kernel = largest summed device time over all its launches in the step (instrument overhead removed)
        by_kernel = {}
        for r in rows:
            # one row per kernel: template instantiations (occupancy / tile variants of one source kernel) are merged
            k = by_kernel.setdefault(r[1].strip("()").split("<")[0], [0, 0.0, 0.0, 0, 0.0])
            k[0] += r[2]; k[1] += max(r[3] - r[2] * ev_overhead_us * 1e-3, 0.05 * r[3]); k[2] += r[4]; k[3] += r[2] if r[4] > 0 else 0; k[4] += r[5]
        ranked = sorted(by_kernel.items(), key=lambda kv: -kv[1][1])
        table = [dict(kernel=k, launches_per_step=v[0] / nprof, ms_per_step=round(v[1] / nprof, 4),
                      alg_gbs=round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 and v[3] == v[0] else None,
                      alg_tflops=round(v[4] / (v[1] * 1e-3) / 1e12, 2) if v[1] > 0 and v[4] > 0 else None) for k, v in ranked[:14]]
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")      # dram bytes per launch from the committed ncu --set full captures
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("kernels", {})
        TENSOR_KERNELS = ("pw_tc_kernel", "pw_wgrad_tc_kernel", "conv3_fwd_tc_kernel", "conv3_wgrad_tc_kernel", "conv3_dgrad_tc_kernel")
        for top_name, top in ranked:
            if not (top[3] == top[0] and top[1] > 0):
                continue                    # kernels without an algorithmic-byte model (attention: flop-bound) are skipped
            sec = top[1] * 1e-3
            ach_b = top[2] / sec / 1e9
            is_tc = top_name.split("<")[0] in TENSOR_KERNELS
            # 3xTF32 issues three tensor-core products per algorithmic one: the pipe peak for ALGORITHMIC flops is tf32 / 3
            pipe_peak = (tf32_peak / 3.0) if is_tc else fp32_tflops
            ach_f = top[4] / sec / 1e12
            t_hbm, t_pipe = top[2] / (hbm * 1e9), top[4] / (pipe_peak * 1e12)
            binds = "hbm" if t_hbm >= t_pipe else ("tensor" if is_tc else "fp32_fma")
            tr = traffic.get(top_name.split("<")[0])
            roof = {"bound": "hbm", "kernel": top_name, "achieved": round(ach_b, 1), "peak": hbm, "unit": "GB/s",
                    "frac": round(ach_b / hbm, 4), "traffic": tr["dram_bytes_per_launch"] if tr else None,
                    "traffic_source": tr["source"] if tr else None, "peak_source": how,
                    "binding_bound": binds, "frac_of_binding_bound": round(max(t_hbm, t_pipe) / sec, 4),
                    "flops": {"achieved_tflops": round(ach_f, 3), "pipe": "tcgen05 tf32 (3 products per algorithmic product)" if is_tc else "fp32 FMA",
                              "pipe_peak_tflops": round(pipe_peak, 1), "frac": round(ach_f / pipe_peak, 4)},
                    "launches_per_step": top[0] / nprof, "alg_bytes_per_launch": round(top[2] / top[0]),
                    "kernel_us_avg": round(1e3 * top[1] / top[0], 2),
                    "event_overhead_us_subtracted": round(ev_overhead_us, 2),
                    "share_of_own_kernel_time": round(top[1] / sum(v[1] for v in by_kernel.values()), 4),
                    "own_kernel_ms_per_step": round(sum(v[1] for v in by_kernel.values()) / nprof, 3),
                    "own_launches_per_step": n_launch,
                    "note": "event-timed launch by launch in an eager step that is serialised on one stream (no forked branches, "
                            "weight gradients on the main stream) and enqueued behind a spin kernel (device back-to-back, no host "
                            "gaps), minus the profiler's own per-launch overhead measured on an empty kernel; in the timed "
                            "graph-replayed step the forked streams overlap these kernels, so the summed kernel time exceeds "
                            "ms_per_step; working sets are L2-resident at 4 patches (DESIGN.md section 3)"}
            break
    # ---- second headline metric: sliding-window inference (all ranks take part)
    used_graph = ts.use_graph
    ts.close()                   # the step graph owns the captured NCCL all-reduce: released before anything else touches the group
    del ts, model
    torch.cuda.empty_cache()
    infer = None if args.no_infer else run_infer(rank, world, dev, volume=tuple(int(v) for v in args.infer_volume.split("x")))
    eager = evalf = None
    if world == 1 and not args.no_eager:
        torch.cuda.empty_cache()
        eager = gpu_eager_port(dev, args.workload, PATCHES)
        evalf = eval_forward_legs(dev, args.workload)
    if world > 1:
        dist.barrier()
    if rank != 0:
        return
    total_patches = PATCHES * world * args.steps
    line = {
        "metric": METRIC, "value": round(total_patches / (ms_total * 1e-3), 3), "unit": "patches/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload),
                   "patches_per_gpu": PATCHES, "global_patches": PATCHES * world, "parallelism": f"dp{world}",
                   "cache": "L2 flushed (256 MiB memset) between timed steps",
                   "launch": ("eager launches" if not used_graph else "whole step replayed as one CUDA graph" if world == 1 else
                              ("one CUDA graph per rank: fwd + bwd + NCCL all-reduce of the flat 9 MB gradient (captured) + AdamW" if os.environ.get("VX_DP_GRAPH", "one") != "split"
                               else "CUDA graph (fwd+bwd) -> eager NCCL all-reduce of the flat 9 MB gradient -> CUDA graph (AdamW)")),
                   "pointwise": ("tcgen05 3xTF32 (fp32-accurate) for S>=1024 (weight gradients S>=512), the level-1/2 FFN pairs as one "
                                 "tcgen05 kernel per direction, fp32 SIMT below (levels 3-4: FFN pairs fused as well)") if pw_tc else "fp32 SIMT",
                   "attention": "forward of windows with L>=128 tokens (level 2) on tcgen05 3xTF32 with P in tensor memory; the other levels and the backward fp32 SIMT",
                   "convolutions": "all libveloxseg: out_conv1 / RC out_conv tcgen05 3xTF32 implicit GEMM with fused bias + PixelShuffle, "
                                   "DownConv / UpConv / heads / stems fp32 SIMT; no cuDNN or cuBLAS kernel in the step",
                   },
        "e2e": {"value": round(total_patches / e2e_s, 3), "unit": "patches/s",
                "h2d_bytes_per_step": int(xb[0].numel() * xb[0].element_size() + yb[0].numel() * yb[0].element_size()),
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "host_enqueue_ms_per_step": round(enqueue_ms, 2), "clocks": clk.summary(), "roofline": roof, "loss": last_loss, "top_kernels": table,
        # whole step against the pipes: algorithmic flops / bytes of every launch that carries a model
        # (contractions, convolutions; the attention and element-wise kernels carry bytes only or nothing) over the timed step
        "step_rates": (lambda fl, by: {"alg_gflop_per_step": round(fl / 1e9, 2), "alg_tflops": round(fl / (ms_total / args.steps * 1e-3) / 1e12, 2),
                                       "frac_of_fp32_fma_peak": round(fl / (ms_total / args.steps * 1e-3) / 1e12 / fp32_tflops, 4) if fp32_tflops else None,
                                       "alg_gbytes_per_step": round(by / 1e9, 3),
                                       "alg_gbs": round(by / (ms_total / args.steps * 1e-3) / 1e9, 1)})(
            sum(r[5] for r in rows) / nprof, sum(r[4] for r in rows) / nprof) if rank == 0 and rows else None,
        "peaks": {"hbm_gbs": hbm, "bf16_tflops": tensor_peak, "source": how, "fp32_fma_tflops_measured": round(fp32_tflops, 1),
                  "fp32_fma_how": "vx_microbench: 8 independent FMA chains per thread, 148 x 8 CTAs of 256 threads, best of 5"},
    }
    if infer is not None:
        line["infer"] = infer
    if eager is not None:
        line["gpu_eager_baseline"] = eager
        line["eval_forward"] = evalf
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_port(steps=3, patches=1, threads=None, cfg_name=args.workload, min_seconds=10.0)   # ~10 s of CPU work
    print(json.dumps(line), flush=True)


def run_infer(rank, world, dev, reps=3, volume=(320, 320, 256)):
    """BASELINE.json configs[3]: Hecktor2022 sliding-window inference on one synthetic PET/CT volume (1,2,320,320,256),
    roi (128,128,64), overlap 0.25 -> 45 windows, sharded round-robin over ranks, partial sums all-reduced."""
    from veloxseg_b200.configs import MODEL_CONFIGS, TRAIN
    from veloxseg_b200.inference import GraphedPredictor, sliding_window_labels, sliding_window_predict, window_starts
    from veloxseg_b200.nn import VeloxSeg
    cfg = MODEL_CONFIGS["hecktor2022"]
    torch.manual_seed(12345)
    model = VeloxSeg(**cfg).to(dev).eval()
    vol_shape = (1, sum(cfg["in_ch"])) + tuple(volume)
    vol_h = torch.randn(vol_shape, generator=torch.Generator().manual_seed(5)).pin_memory()
    roi, sw = cfg["input_size"], 4
    pred = GraphedPredictor(model, sw, vol_shape[1], roi, dev)
    nwin = len(window_starts(vol_shape[2:], roi, TRAIN["sw_overlap"]))
    # default: sharded IO (1/world of the host volume per rank + NVLink all-gather, reduce-scatter of the partial sums, label
    # gather) -- measured 13.5 vs 26.8 ms/volume at 2 GPUs (profiles/r2v_*_2gpu.log); VX_INFER_IO=replicated keeps the
    # round-1 path (every rank copies the whole volume, all-reduce of the logit sums) for A/B runs
    sharded_io = os.environ.get("VX_INFER_IO", "sharded") == "sharded"
    seg_h = torch.zeros(vol_shape[2:], dtype=torch.uint8).pin_memory()
    times = []
    for i in range(reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if sharded_io:                                              # see inference.sliding_window_labels
            seg = sliding_window_labels(vol_h, pred, roi, dev, sw_batch_size=sw, overlap=TRAIN["sw_overlap"], out_host=seg_h)
            seg = seg_h if seg is None else seg                     # ranks other than 0 hold no result
        else:
            vol = vol_h.to(dev, non_blocking=True)                  # host volume in, host label map out
            out = sliding_window_predict(vol, pred, roi, sw_batch_size=sw, overlap=TRAIN["sw_overlap"])
            seg = out.argmax(1).to(torch.uint8).cpu()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i:
            times.append(dt)
    t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"metric": "sliding-window infer ms/volume", "value": round(float(t.item()) * 1e3, 2), "unit": "ms/volume",
            "higher_is_better": False, "scaling": "strong", "n_gpus": world,
            "config": {"workload": "VeloxSeg Hecktor2022 eval, volume 2x%s, roi 128x128x64, overlap 0.25" % "x".join(map(str, volume)),
                       "windows": nwin, "sw_batch": sw,
                       "io": (("pinned host volume uploaded in (plane slab x row slab) blocks on a copy stream in window order, a batch "
                               "of windows waits only for its blocks; finished label planes (count division + arg-max) stream back "
                               "under the remaining windows") if sharded_io and world == 1 else
                              "1/world of the host volume per rank + all-gather; reduce-scatter, arg-max per slab, label gather"
                              if sharded_io else "every rank copies the whole host volume; all-reduce of the logit sums"),
                       "timed": "H2D volume + windows + all-reduce + argmax + D2H labels, best of %d" % reps,
                       "fg_voxels": int(seg.sum())}}


def gpu_eager_port(dev, cfg_name, patches, steps=10, warmup=3):
    """The GPU bar: the same reference math (the oracle's plain-torch
    restatement: F.conv3d / instance_norm / einsum / softmax leaf ops, autograd backward, torch AdamW) under PyTorch eager on
    this B200 -- fp32 (TF32 off) and torch.autocast(bfloat16).  Baseline only; nothing of libveloxseg runs here."""
    from oracle import veloxseg_oracle as O
    from veloxseg_b200.configs import MODEL_CONFIGS, TRAIN
    from veloxseg_b200.nn import VeloxSeg
    cfg = MODEL_CONFIGS[cfg_name]
    spec = O.ModelSpec(cfg)
    x, y = synth_batch(cfg, patches, 1000)
    x, y = x.to(dev), y.to(dev)
    out = {}
    for mode in ("fp32", "autocast_bf16"):
        torch.manual_seed(12345)
        m = VeloxSeg(**cfg)
        p = {k: (v.detach().clone().to(dev).requires_grad_(True) if v.dtype.is_floating_point else v.to(dev)) for k, v in m.state_dict().items()}
        leaves = [v for v in p.values() if v.dtype.is_floating_point]
        opt = torch.optim.AdamW(leaves, lr=TRAIN["lr"], weight_decay=TRAIN["weight_decay"], fused=True)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "autocast_bf16"):
                outs = O.forward(x, p, spec, training=True)
            loss = O.total_loss([o.float() for o in outs], y, x, spec.M, TRAIN["deep_Loss_weight"], TRAIN["RC_Loss_weight"], TRAIN["Feature_Loss_weight"])
            loss.backward()
            opt.step()
            return loss
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            loss = step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        out[mode] = {"ms_per_step": round(ms, 3), "value": round(patches / (ms * 1e-3), 2), "unit": "patches/s", "loss": float(loss)}
        del opt, p, leaves
        torch.cuda.empty_cache()
    out["what"] = ("oracle restatement of the reference (plain torch leaf ops + autograd + fused torch AdamW) under PyTorch eager on this GPU, "
                   "%d patches/step, %d timed steps, cudnn.benchmark on, TF32 off" % (patches, steps))
    return out


def eval_forward_legs(dev, cfg_name, seconds=3.0):
    """BASELINE configs[0] / the reference's only published numbers (README: 599.06 patches/s GPU, 6.67 patches/s CPU):
    speed_test.py's protocol -- eval(), grad off, randn input; GPU: largest power-of-two batch <= 16, synchronize per
    iteration; CPU: batch 1, ONE thread (speed_test.py:27,64-70,102-134).  Time-boxed to `seconds` per leg instead of 60 s."""
    from oracle import veloxseg_oracle as O
    from veloxseg_b200.configs import MODEL_CONFIGS
    from veloxseg_b200.nn import VeloxSeg
    cfg = MODEL_CONFIGS[cfg_name]
    spec = O.ModelSpec(cfg)
    torch.manual_seed(12345)
    model = VeloxSeg(**cfg)
    p_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    p_gpu = {k: v.to(dev) for k, v in p_cpu.items()}
    model = model.to(dev).eval()
    res = {}

    def timed(fn, bs, sync):
        fn()
        fn()
        sync()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds:
            fn()
            sync()
            n += 1
        return round(n * bs / (time.perf_counter() - t0), 2)
    with torch.no_grad():
        for bs in (1, 16):
            xg = torch.randn(bs, sum(cfg["in_ch"]), *cfg["input_size"], device=dev)
            res["ours_fp32_bs%d" % bs] = timed(lambda: model(xg), bs, torch.cuda.synchronize)
            g = torch.cuda.CUDAGraph()
            s_ = torch.cuda.Stream(device=dev)
            s_.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s_):
                model(xg)
            torch.cuda.current_stream(dev).wait_stream(s_)
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                yg = model(xg)
            res["ours_fp32_graph_bs%d" % bs] = timed(g.replay, bs, torch.cuda.synchronize)
            del g, yg
            torch.backends.cudnn.allow_tf32 = False
            res["eager_fp32_bs%d" % bs] = timed(lambda: O.forward(xg, p_gpu, spec, training=False), bs, torch.cuda.synchronize)

            def amp():
                with torch.autocast("cuda", dtype=torch.float16):
                    O.forward(xg, p_gpu, spec, training=False)
            res["eager_autocast_fp16_bs%d" % bs] = timed(amp, bs, torch.cuda.synchronize)
        nthr = torch.get_num_threads()
        torch.set_num_threads(1)
        xc = torch.randn(1, sum(cfg["in_ch"]), *cfg["input_size"])
        res["cpu_1thread_fp32_bs1"] = timed(lambda: O.forward(xc, p_cpu, spec, training=False), 1, lambda: None)
        torch.set_num_threads(nthr)
    res["unit"] = "patches/s"
    res["protocol"] = ("speed_test.py: eval forward, grad off, randn %s input, synchronize per iteration, %.0f s per leg; 'ours' = "
                       "veloxseg_b200.nn.VeloxSeg (eager launches / one CUDA graph), 'eager' = oracle restatement of the reference under "
                       "PyTorch eager on this GPU, cpu = the same on ONE host thread (reference README: 599.06 GPU fp16 on its hardware, "
                       "6.67 CPU)" % (cfg_name, seconds))
    return res


def cpu_port(steps, patches, threads, cfg_name=CFG_NAME, min_seconds=0.0, warmup=1):
    """The oracle's restatement of the reference CPU path: train step (fwd + full loss + bwd + AdamW) on host cores."""
    from oracle import veloxseg_oracle as O
    from veloxseg_b200.configs import MODEL_CONFIGS, TRAIN
    from veloxseg_b200.nn import VeloxSeg
    cores = threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    cfg = MODEL_CONFIGS[cfg_name]
    torch.manual_seed(12345)
    m = VeloxSeg(**cfg)
    p = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in m.state_dict().items()}
    leaves = [v for v in p.values() if v.dtype.is_floating_point]
    opt = torch.optim.AdamW(leaves, lr=TRAIN["lr"], weight_decay=TRAIN["weight_decay"])
    spec = O.ModelSpec(cfg)
    x, y = synth_batch(cfg, patches, 1000)

    def step():
        opt.zero_grad(set_to_none=True)
        outs = O.forward(x, p, spec, training=True)
        loss = O.total_loss(outs, y, x, spec.M, TRAIN["deep_Loss_weight"], TRAIN["RC_Loss_weight"], TRAIN["Feature_Loss_weight"])
        loss.backward()
        opt.step()
        return float(loss.detach())
    for _ in range(max(1, warmup)):      # warm-up (allocator, thread pool)
        step()
    t0 = time.perf_counter()
    done = 0
    while done < steps or (min_seconds and time.perf_counter() - t0 < min_seconds and done < 200):
        step()
        done += 1
    steps = done
    dt = time.perf_counter() - t0
    return {"value": round(steps * patches / dt, 4), "unit": "patches/s", "cores": cores, "kind": "port",
            "sample": f"{steps} train step(s) of {patches} patch(es) ({sum(cfg['in_ch'])}x{'x'.join(map(str, cfg['input_size']))}, fp32, dropout off) after {max(1, warmup)} warm-up step(s), "
                      f"torch CPU with {cores} threads"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    patches = 1
    steps = max(1, min(args.steps, 100))      # one patch per step, ~0.3-0.6 s each on the box's cores: bounded at about a minute
    warm = max(1, min(args.warmup, 5))
    base = cpu_port(steps=steps, patches=patches, threads=None, cfg_name=args.workload, warmup=warm)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "patches/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": round(1e3 * patches / base["value"], 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload),
                       "sample": "reference CPU path (oracle port); each timed step is a bounded sample of 1 of the 4 patches",
                       "parallelism": "host cores"},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "patches/s", "h2d_bytes_per_step": 0,
                                          "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=CFG_NAME, choices=sorted(CFG_TITLE),
                    help="model config of the train step: autopetii = BASELINE configs[1] (the headline metric), brats2021 = configs[2]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="(accepted for old scripts; there is no alternate precision leg any more)")
    ap.add_argument("--no-infer", action="store_true", help="skip the sliding-window inference measurement")
    ap.add_argument("--only-infer", action="store_true", help="run only the sliding-window inference leg and print its object")
    ap.add_argument("--infer-volume", default="320x320x256",
                    help="synthetic Hecktor volume of the sliding-window leg: 320x320x256 (45 windows, default) or 512x512x384 (200 windows)")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager GPU baseline and the eval-forward legs")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.only_infer:
            torch.cuda.set_device(local_rank)
            res = run_infer(rank, world, torch.device("cuda", local_rank), volume=tuple(int(v) for v in args.infer_volume.split("x")))
            if rank == 0:
                print(json.dumps(res), flush=True)
            return
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
