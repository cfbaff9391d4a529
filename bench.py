#!/usr/bin/env python
"""bench.py — PoseTraj denoising hot path on B200 (BASELINE.json metric: UNet+ControlNet denoise steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is ONE denoise step of the reference loop
(/root/reference/pipeline/pipeline_stable_video_diffusion_controlnet.py:530-572): ControlNetSDV forward + UNet forward
on the CFG pair (2 x 14 frames x 8 ch x 40x72 latent, 320x576 px), CFG combine and the Euler-Karras update — the
workload of BASELINE.json configs[1] (25-step sampling, bf16 storage / fp32 accumulate, random-init SVD-shaped
weights, synthetic inputs).

  value      steps/s with the video's inputs resident in HBM (CUDA-graph replay of the per-step kernel list), whole job
             over N GPUs; N > 1 = one independent video per rank (video-batch sharding, weak scaling, no collective
             on the data path: SURVEY.md §8e).
  e2e        the same metric through the public API, `StableVideoDiffusionPipelineControlNet.__call__`, fed HOST
             (pinned) tensors, with the host->device copies of the conditioning and the device->host read of the
             final latents inside the timed region.
  roofline   the dominant kernel (tcgen05 GEMM / implicit-GEMM conv): algorithmic FLOPs / CUDA-event time of its
             launches inside one step, against MEASURED_PEAKS.json.
  cpu_baseline / --impl reference   the oracle (CPU restatement of the reference; the reference itself needs
             diffusers==0.24.0 which does not exist in this image) on all host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "denoise_steps_per_sec"
UNIT = "steps/s"
FRAMES, LAT_H, LAT_W, SAMPLING_STEPS = 14, 40, 72, 25
WORKLOAD = "configs[1]: 25-step Euler-Karras SVD-img2vid, CFG pair x 14 frames x 320x576 (latent 40x72), 1 trajectory"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def _ncu_traffic(kind):
    """Per-launch DRAM traffic of a kernel class from the committed ncu capture (None if the file is absent)."""
    path = os.path.join(ROOT, "profiles", "r3z_dram_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)["classes"][kind]["traffic_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        return None


# ---------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe) during the timed region
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (cpu_baseline of the own arm, and `--impl reference`)
# ---------------------------------------------------------------------------------------------------------------
def _cpu_sample_plan(budget_s_per_step: float):
    """Pick a bounded sample of the workload: (frames, h, w).  ~7 s per frame at 40x72 on 8 cores (measured)."""
    import torch
    cores = torch.get_num_threads()
    per_frame = 7.0 * 8 / max(cores, 1)
    if budget_s_per_step >= 2 * per_frame:
        return 2, LAT_H, LAT_W
    if budget_s_per_step >= per_frame:
        return 1, LAT_H, LAT_W
    return 1, 24, 40  # quarter-ish resolution: still every layer of both networks


def cpu_oracle_steps_per_sec(steps: int, warmup: int, budget_s: float = 150.0, full: bool = False):
    """Times the oracle (fp32, all host threads).  `full`: ONE step of the whole workload (CFG pair x 14 frames x 40x72,
    ~20 s on 16 cores) after a reduced-size warm-up — no extrapolation.  Otherwise a bounded sample (fewer frames /
    lower resolution, every layer of both networks), scaled by the algorithmic FLOP count, sized to `budget_s`."""
    import torch
    from oracle.models import build_models
    from oracle.pipeline import denoise_step, make_inputs
    from oracle.scheduler import EulerKarrasOracle
    from posetraj_b200.config import SVDConfig
    from posetraj_b200.roofline import step_flops
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet, cnet = build_models(seed=0, randomize_zero_convs=True)
    sched = EulerKarrasOracle()
    sched.set_timesteps(SAMPLING_STEPS)

    def one(f, h, w, k):
        inp = make_inputs(num_frames=f, h=h, w=w)
        cond = torch.full((2, f, 3, h * 8, w * 8), -1.0)
        sched._step_index = None
        t0 = time.perf_counter()
        denoise_step(unet, cnet, sched, inp["latents"], k, sched.timesteps[k], inp["image_latents"], inp["image_embeddings"],
                     cond, inp["added_time_ids"], inp["guidance"])
        return time.perf_counter() - t0

    if full:
        one(2, 16, 24, 0)                                # pages the 8.8 GB of fp32 weights in, warms the thread pool
        times = [one(FRAMES, LAT_H, LAT_W, k % SAMPLING_STEPS) for k in range(max(1, steps))]
        t_step = sum(times) / len(times)
        sample = (f"oracle (torch CPU fp32 restatement of the reference), the FULL step: CFG pair x {FRAMES} frames at latent "
                  f"{LAT_H}x{LAT_W}, {len(times)} timed step(s) of {t_step:.2f} s after a reduced-size warm-up; not extrapolated")
        return 1.0 / t_step, t_step, cores, sample
    f, h, w = _cpu_sample_plan(budget_s / max(1, steps + warmup))
    times = []
    for i in range(warmup + steps):
        dt = one(f, h, w, i % SAMPLING_STEPS)
        if i >= warmup:
            times.append(dt)
    t_step = sum(times) / len(times)
    full_fl, _ = step_flops(SVDConfig(), frames=FRAMES, h=LAT_H, w=LAT_W)
    part, _ = step_flops(SVDConfig(), frames=f, h=h, w=w)
    value = (part / full_fl) / t_step
    sample = (f"oracle (torch CPU fp32 restatement of the reference) on {f} of {FRAMES} frames at latent {h}x{w}, "
              f"{steps} timed step(s) of {t_step:.2f} s each, scaled by algorithmic FLOPs {part / 1e12:.2f}/{full_fl / 1e12:.2f} TFLOP")
    return value, t_step, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the whole step costs ~20 s on 16 cores: run it un-sampled when the requested steps fit in a few minutes
    full = args.steps <= 6
    value, t_step, cores, sample = cpu_oracle_steps_per_sec(args.steps, args.warmup, full=full)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "weights": "random-init SVD-shaped UNet 1524.6M + ControlNetSDV 682.0M",
                   "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------------
def time_op_classes(step_ops, stream, reps: int = 2):
    """CUDA-event time of every launch descriptor of one step (eager replay on `stream`), summed per kernel class."""
    import torch
    acc = {}
    for rep in range(reps + 1):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(step_ops) + 1)]
        evs[0].record(stream)
        for i, op in enumerate(step_ops):
            op.launch(stream.cuda_stream)
            evs[i + 1].record(stream)
        stream.synchronize()
        if rep == 0:
            continue  # warm
        for i, op in enumerate(step_ops):
            ms = evs[i].elapsed_time(evs[i + 1])
            d = acc.setdefault(op.kind, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += ms / reps
            d["flops"] += op.alg_flops / reps
            d["bytes"] += op.alg_bytes / reps
            d["launches"] += 1.0 / reps
    return acc


def _synthetic_tracks(frames):
    """One track, `frames` points linearly from (100,80) to (460,240) px."""
    return [[[int(round(100 + (460 - 100) * k / (frames - 1))), int(round(80 + (240 - 80) * k / (frames - 1)))]
             for k in range(frames)]]


def full_pipeline_leg(unet, cnet, tracks, dev):
    """Extra (not the contract's metric): one whole video through the public API on this library's kernels only — host
    uint8 image + host tracks -> R1 rasteriser, anti-aliased resize + CLIP ViT-H/14, VAE encode, 25 denoise steps, VAE
    temporal decode -> 14 x 320 x 576 frames on the host (SURVEY.md §8f rows 2-3 next to the hot path)."""
    import numpy as np
    import torch
    from posetraj_b200.clip import CLIPVisionConfig, CLIPVisionModelWithProjection
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    from posetraj_b200.trajectory import rasterize_tracks
    from posetraj_b200.vae import AutoencoderKLTemporalDecoder, VaeConfig
    try:
        vae = AutoencoderKLTemporalDecoder.from_random(VaeConfig(), dev, seed=0)
        clip = CLIPVisionModelWithProjection.from_random(CLIPVisionConfig(), dev, seed=0)
        pipe = StableVideoDiffusionPipelineControlNet(vae=vae, image_encoder=clip, unet=unet, controlnet=cnet)
        H, W = LAT_H * 8, LAT_W * 8
        image = np.random.default_rng(7).integers(0, 256, size=(H, W, 3), dtype=np.uint8).astype(np.float32) / 255.0

        def call():
            cond = rasterize_tracks(tracks, FRAMES, H, W, dev, output="f32")
            out = pipe(image, cond, height=H, width=W, num_frames=FRAMES, num_inference_steps=SAMPLING_STEPS,
                       generator=torch.Generator().manual_seed(0), output_type="np")
            return out.frames[0]

        frames = call()
        if frames.shape != (FRAMES, H, W, 3) or not np.isfinite(frames).all():
            return {"error": f"bad frames {frames.shape}"}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            call()
        torch.cuda.synchronize()
        s_per_video = (time.perf_counter() - t0) / n
        # stage split with CUDA events (device time of each stage, inputs already on the device)
        def dev_ms(fn):
            fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            return round(e0.elapsed_time(e1), 3)
        img_d = torch.from_numpy(image).permute(2, 0, 1)[None].to(dev)
        lat = torch.randn(FRAMES, 4, LAT_H, LAT_W, device=dev)
        stages = {"resize_clip_ms": dev_ms(lambda: clip.encode_image(img_d)),
                  "vae_encode_ms": dev_ms(lambda: vae.encode(img_d * 2 - 1)),
                  "vae_decode_ms": dev_ms(lambda: vae.decode(lat, num_frames=FRAMES))}
        return {"s_per_video": round(s_per_video, 4), "videos_per_min": round(60.0 / s_per_video, 2), "stages": stages,
                "api": "StableVideoDiffusionPipelineControlNet.__call__(host image, tracks -> rasterize_tracks, output_type='np'): "
                       "CLIP ViT-H/14 632M + VAE 97.7M random-init, 25 steps, frames [14,320,576,3] returned on the host"}
    except Exception as e:  # the extra must never take the contract line down with it
        return {"error": f"{type(e).__name__}: {e}"[:300]}



# ---------------------------------------------------------------------------------------------------------------
# extra legs (BASELINE.json configs[2] / configs[4]); `value` stays the video-batch number
# ---------------------------------------------------------------------------------------------------------------
def _video_inputs(cfg, frames, h, w, seed, pin=True):
    import torch
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(1, frames, 4, h, w, generator=g)
    img = torch.randn(1, 4, h, w, generator=g)
    emb = torch.randn(1, 1, cfg.cross_attention_dim, generator=g)
    out = dict(latents=lat, image_latents=torch.cat([torch.zeros_like(img), img]),
               image_embeddings=torch.cat([torch.zeros_like(emb), emb]))
    return {k: (v.pin_memory() if pin else v) for k, v in out.items()}


def _max_over_ranks(x, dev, world):
    import torch
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _rel_l2(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _time_engine_steps(eng, steps, dev, world, barrier):
    """Device time of `steps` denoise steps of an already captured engine, max over ranks (ms per step)."""
    import torch
    eng.reset()
    eng.step()
    eng.reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        if i and i % SAMPLING_STEPS == 0:
            eng.reset()
        eng.step()
    e1.record()
    barrier()
    return _max_over_ranks(e0.elapsed_time(e1) / steps, dev, world)


def cam_leg(cfg, unet, dev, rank, world, steps, barrier):
    """BASELINE.json configs[2]: controlnet_sdv_cam (camera_control_module branch: cc_projection on [features | camera_RT],
    /root/reference/models/controlnet_sdv_cam_infer.py:96-122), one video (CFG pair) per GPU, N videos on N GPUs."""
    import torch
    from posetraj_b200.models import ControlNetSDVModel
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    from posetraj_b200.trajectory import rasterize_tracks
    cnet = ControlNetSDVModel.from_random(cfg, dev, seed=0, cam=True, faithful_zero_init=False)
    pipe = StableVideoDiffusionPipelineControlNet(unet=unet, controlnet=cnet)
    H, W = LAT_H * 8, LAT_W * 8
    inp = _video_inputs(cfg, FRAMES, LAT_H, LAT_W, 4321 + rank)
    tracks = torch.tensor(_synthetic_tracks(FRAMES), dtype=torch.int32).pin_memory()
    g = torch.Generator().manual_seed(99 + rank)
    cam = (torch.randn(FRAMES, 12, generator=g) * 0.1).pin_memory()
    cam[0] = 0.0   # relative pose of frame 0 is the identity offset (infer/run_inference_vipseg_json_cam_concat_repro.py:485-496)

    def call():
        cond = rasterize_tracks(tracks, FRAMES, H, W, dev, output="f32")
        return pipe(None, cond, camera_cond=cam, height=H, width=W, num_frames=FRAMES, num_inference_steps=SAMPLING_STEPS,
                    latents=inp["latents"], output_type="latent", image_embeddings=inp["image_embeddings"],
                    image_latents=inp["image_latents"]).frames.to("cpu")

    out = call()
    eng = pipe.engine_for(FRAMES, LAT_H, LAT_W, (H, W))
    ms = _time_engine_steps(eng, steps, dev, world, barrier)
    barrier()
    t0 = time.perf_counter()
    call()
    torch.cuda.synchronize()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, dev, world)
    res = {"workload": "configs[2]: controlnet_sdv_cam (camera branch), 14 frames 320x576, one video per GPU",
           "ms_per_step": ms, "value": world * 1000.0 / ms, "unit": UNIT, "videos_per_min": world * 60000.0 / ms / SAMPLING_STEPS,
           "e2e_value": world * SAMPLING_STEPS / e2e_s, "finite": bool(torch.isfinite(out).all())}
    del pipe, cnet, eng
    torch.cuda.empty_cache()
    return res


def cfg_split_leg(cfg, unet, cnet, dev, rank, world, steps, barrier, ms_replica, lat_single_rank0):
    """SURVEY.md 8e "CFG branch": one video on 2 GPUs, rank r of a pair runs row r of the CFG pair through both networks,
    the two 161 KB predictions are all-gathered over NCCL each step, both ranks redo the CFG + Euler update."""
    import torch
    import torch.distributed as dist
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    from posetraj_b200.sharding import cfg_split_ranks
    from posetraj_b200.trajectory import rasterize_tracks
    pairs = cfg_split_ranks(world)
    groups = [dist.new_group(ranks=list(p)) for p in pairs]   # every rank creates every group (collective)
    pair, row = rank // 2, rank % 2
    pipe = StableVideoDiffusionPipelineControlNet(unet=unet, controlnet=cnet)
    pipe.enable_cfg_split(row, group=groups[pair])
    H, W = LAT_H * 8, LAT_W * 8
    inp = _video_inputs(cfg, FRAMES, LAT_H, LAT_W, 1234 + 2 * pair)   # pair 0 = the video rank 0 ran alone
    tracks = torch.tensor(_synthetic_tracks(FRAMES), dtype=torch.int32).pin_memory()

    def call():
        cond = rasterize_tracks(tracks, FRAMES, H, W, dev, output="f32")
        return pipe(None, cond, height=H, width=W, num_frames=FRAMES, num_inference_steps=SAMPLING_STEPS,
                    latents=inp["latents"], output_type="latent", image_embeddings=inp["image_embeddings"],
                    image_latents=inp["image_latents"]).frames.to("cpu")

    out = call()
    eng = pipe.engine_for(FRAMES, LAT_H, LAT_W, (H, W))
    ms = _time_engine_steps(eng, steps, dev, world, barrier)
    barrier()
    t0 = time.perf_counter()
    call()
    torch.cuda.synchronize()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, dev, world)
    res = {"workload": "configs[1], one video per PAIR of GPUs (rank = row of the CFG pair), NCCL all-gather of 2 x 161 KB per step",
           "pairs": len(pairs), "ms_per_step": ms, "value": len(pairs) * 1000.0 / ms, "unit": UNIT,
           "latency_speedup_vs_1gpu": ms_replica / ms, "e2e_value": len(pairs) * SAMPLING_STEPS / e2e_s,
           "backend": dist.get_backend(groups[pair]), "launches_per_step": eng.launches_per_step}
    if rank == 0 and lat_single_rank0 is not None:
        res["rel_l2_latents_25_steps_vs_1gpu"] = _rel_l2(out, lat_single_rank0)
    del pipe, eng
    torch.cuda.empty_cache()
    return res


def frame_sharded_leg(cfg, unet, cnet, dev, rank, world, steps, barrier):
    """BASELINE.json configs[4] / SURVEY.md 8e "Frames": ONE 25-frame 576x1024 video (latent 72x128) over all N GPUs —
    frame-sharded spatial layers, pixel-sharded temporal layers, the exchange fused into the producing GEMM's epilogue
    over NVLink peer memory (PtGemmArgs.scatter_mode), one device barrier per exchange, no NCCL inside the step."""
    import torch
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    from posetraj_b200.roofline import step_flops
    from posetraj_b200.trajectory import rasterize_tracks
    F, h, w = 25, 72, 128
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    inp = _video_inputs(cfg, F, h, w, 777, pin=False)   # the same video on every rank
    tracks = [[[20 + 9 * k, 30 + 5 * k] for k in range(F)]]
    n_par = 3

    def run(pipe, n_steps):
        cond = rasterize_tracks(tracks, F, h * 8, w * 8, dev)
        return pipe(None, cond, height=h * 8, width=w * 8, num_frames=F, num_inference_steps=n_steps, output_type="latent",
                    latents=inp["latents"], image_embeddings=inp["image_embeddings"], image_latents=inp["image_latents"]).frames

    pipe = StableVideoDiffusionPipelineControlNet(unet=unet, controlnet=cnet)
    pipe.enable_frame_sharding(rank, world)
    got = run(pipe, n_par).float().cpu()
    eng = pipe.engine_for(F, h, w, (h * 8, w * 8))
    ms = _time_engine_steps(eng, steps, dev, world, barrier)
    fl, _ = step_flops(cfg, frames=F, h=h, w=w, essential=True)
    res = {"workload": "configs[4]: 25 frames x 576x1024 (latent 72x128), ONE video frame-sharded over all GPUs",
           "ms_per_step": ms, "value": 1000.0 / ms, "unit": UNIT, "aggregate_tflops": fl / ms / 1e9,
           "exchange": "p2p scatter epilogue" if getattr(eng.cplan, "p2p", False) else "nccl all-to-all",
           "graph": eng.graph is not None, "launches_per_step": eng.launches_per_step,
           "barriers_or_collectives_per_step": eng.collectives_per_step,
           "peak_mem_gib_per_rank": torch.cuda.max_memory_allocated() / 2 ** 30}
    eng.graph = None
    del pipe, eng
    torch.cuda.empty_cache()
    barrier()
    # rank 0 alone: the unsharded plan on the same inputs — parity of the sharded latents and the in-run 1-GPU time
    if rank == 0:
        pipe1 = StableVideoDiffusionPipelineControlNet(unet=unet, controlnet=cnet)
        want = run(pipe1, n_par).float().cpu()
        eng1 = pipe1.engine_for(F, h, w, (h * 8, w * 8))
        ms1 = _time_engine_steps(eng1, 3, dev, 1, torch.cuda.synchronize)
        res["rel_l2_latents_3_steps_vs_1gpu"] = _rel_l2(got, want)
        res["ms_per_step_1gpu_same_run"] = ms1
        res["speedup_vs_1gpu"] = ms1 / ms
        res["ideal_ms_per_step"] = ms1 / world
        del pipe1, eng1
        torch.cuda.empty_cache()
    barrier()
    return res



def train_dp_leg(cfg, dev, rank, world, barrier):
    """BASELINE.json configs[3] (controlnet_sdv_bbox pre-training, data-parallel), the parts that exist as kernels
    (posetraj_b200/training.py, DESIGN.md "Training"): forward + backward of one level-0 SpatioTemporalResBlock at the
    per-rank shape (2 videos x 14 frames x 40x72, 320 channels) — 3x3 conv and temporal conv dgrad (pt_gemm) / wgrad
    (pt_wgrad, tcgen05), 4-D and 5-D GroupNorm+SiLU backward — and the data-parallel tail of a step at FULL size: the
    683 M ControlNet gradients in 100 MB fp32 buckets, NCCL all-reduce per bucket, fused AdamW.  (The whole step is the
    next leg, `train_step_leg`.)"""
    import math
    import torch
    from posetraj_b200 import training as T
    from posetraj_b200.config import controlnet_param_shapes
    B, Fr, H, W, Cc, temb = 2, FRAMES, LAT_H, LAT_W, cfg.block_out_channels[0], cfg.block_out_channels[0] * 4
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    rn = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
    params = {}
    for blk, conv in (("spatial_res_block", (Cc, Cc, 3, 3)), ("temporal_res_block", (Cc, Cc, 3, 1, 1))):
        for i in (1, 2):
            params[f"{blk}.norm{i}.weight"], params[f"{blk}.norm{i}.bias"] = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
            params[f"{blk}.conv{i}.weight"] = rn(*conv, sc=1 / math.sqrt(Cc * 9))
            params[f"{blk}.conv{i}.bias"] = rn(Cc, sc=0.02)
    params["time_mixer.mix_factor"] = torch.full((1,), 0.5, device=dev)
    tr = T.ResBlockTrainer(params, B=B, F=Fr, H=H, W=W, eps=1e-6)
    rows = B * Fr * H * W
    x, dout = rn(rows, Cc).to(torch.bfloat16), rn(rows, Cc).to(torch.bfloat16)
    ts, tt = rn(B, Cc, sc=0.1), rn(B, Cc, sc=0.1)

    def fb():
        tr.forward(x, ts, tt)
        return tr.backward(dout)

    fb()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        fb()
    e1.record()
    barrier()
    ms_block = _max_over_ranks(e0.elapsed_time(e1) / n, dev, world)
    conv_fl = 2.0 * rows * Cc * Cc * (9 + 9 + 3 + 3)
    # data-parallel tail at the full ControlNet size
    sizes = [int(math.prod(sh)) for sh in controlnet_param_shapes(cfg, cam=False, bbox=True).values()]
    gb = T.GradientBuckets(sizes, dev, bucket_mb=100.0)
    for f in gb.flat:
        f.normal_(generator=g)
    opt = T.AdamW(gb, [torch.zeros(1, device=dev)] * 0, work=[], lr=1e-5)

    def dp_tail():
        for i in reversed(range(len(sizes))):
            gb.ready(i)
        gb.finish()

    dp_tail()
    barrier()
    e0.record()
    dp_tail()
    e1.record()
    barrier()
    ms_ar = _max_over_ranks(e0.elapsed_time(e1), dev, world)
    opt.step()
    barrier()
    e0.record()
    opt.step()
    e1.record()
    barrier()
    ms_opt = _max_over_ranks(e0.elapsed_time(e1), dev, world)
    nbytes = 4.0 * sum(sizes)
    res = {"workload": "configs[3] building blocks: level-0 SpatioTemporalResBlock fwd+bwd (2 videos x 14 frames x 40x72 per rank); "
                       "all-reduce + AdamW of the full 683 M-parameter ControlNet gradient",
           "resblock_fwd_bwd_ms": ms_block, "resblock_conv_tflops": 3.0 * conv_fl / ms_block / 1e9,
           "grad_bytes": nbytes, "buckets": len(gb.buckets), "allreduce_ms": ms_ar if world > 1 else None,
           "allreduce_busbw_gbs": (2.0 * (world - 1) / world * nbytes / ms_ar / 1e6) if world > 1 else None,
           "adamw_ms": ms_opt, "adamw_gbs": (nbytes * 7.5) / ms_opt / 1e6,
           "see": "configs3_train_step for the whole step (this leg keeps the per-block and the collective-only timings)"}
    del gb, opt, tr
    torch.cuda.empty_cache()
    return res


def train_step_leg(cfg, unet, dev, rank, world, barrier, steps=3):
    """BASELINE.json configs[3]: controlnet_sdv_bbox pre-training step, data-parallel, 2 videos per rank (global batch 16 on
    8 GPUs) x 14 frames x 320x576 — the WHOLE step of scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1404-1475 on the CUDA
    library (posetraj_b200/train_engine.py): bbox ControlNet forward, frozen UNet forward, EDM loss, the one-frame
    "spatial" pass, the reverse pass through both networks (tcgen05 dgrad / wgrad, attention backward ...), per-bucket
    NCCL all-reduce overlapped with the reverse pass, fused AdamW, re-derivation of the bf16 kernel weights."""
    import torch
    from posetraj_b200.models import ControlNetSDVModel
    from posetraj_b200.train_engine import ControlNetTrainer
    B, Fr, H, W = 2, FRAMES, LAT_H, LAT_W
    cnet = ControlNetSDVModel.from_random(cfg, dev, seed=5, bbox=True, faithful_zero_init=False)
    torch.cuda.reset_peak_memory_stats(dev)
    tr = ControlNetTrainer(unet, cnet, batch=B, frames=Fr, height=H, width=W, lr=1e-5)
    tr.use_cuda_graph = os.environ.get("PT_TRAIN_GRAPH", "1") != "0"
    g = torch.Generator(device=dev).manual_seed(11 + rank)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    batch = dict(latents=rn(B, Fr, 4, H, W) * 0.9, noise=rn(B, Fr, 4, H, W), sigmas=torch.tensor([1.3, 0.4], device=dev),
                 image_embeddings=rn(B, 1, cfg.cross_attention_dim),
                 trajectories=(torch.rand(B, Fr, 3, 8 * H, 8 * W, device=dev, generator=g) > 0.97).float() * 2 - 1,
                 motion_values=torch.tensor([127.0, 90.0], device=dev),
                 controlnet_bbox=(torch.rand(B, Fr, 3, 8 * H, 8 * W, device=dev, generator=g) > 0.98).float() * 2 - 1)
    losses = [float(tr.step(ran_idx=3, **batch))]           # warm-up (TMA descriptors, attribute caches, NCCL)
    losses.append(float(tr.step(ran_idx=3, **batch)))       # (graph mode: this call captures and replays once)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * steps + 1)]
    lc0 = int(_lib_launches())
    ev[0].record()
    for i in range(steps):
        if tr.use_cuda_graph:
            for k, v in batch.items():
                tr._static[k].copy_(v, non_blocking=True)
            tr._graphs[3].replay()
            loss = tr.loss
        else:
            loss = tr.forward_backward(ran_idx=3, **batch)
        ev[3 * i + 1].record()
        tr.optimizer_step()
        ev[3 * i + 2].record()
        ev[3 * i + 3].record()
        losses.append(loss.clone())                          # (device copy; read after the timed region)
    barrier()
    launches = tr.graph_launches if tr.use_cuda_graph else (int(_lib_launches()) - lc0) / steps
    ms_step = _max_over_ranks(ev[0].elapsed_time(ev[3 * steps]) / steps, dev, world)
    ms_fb = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(steps)) / steps
    ms_tail = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(steps)) / steps
    losses = [float(x) for x in losses]
    res = {"workload": "configs[3]: controlnet_sdv_bbox training step (ControlNet + frozen UNet forward, EDM loss + one-frame "
                       "spatial pass, reverse pass, all-reduce, AdamW), 2 videos x 14 frames x 320x576 per GPU, data-parallel",
           "ms_per_step": ms_step, "value": 1000.0 * B * world / ms_step, "unit": "videos/s (global batch %d)" % (B * world),
           "forward_backward_ms": ms_fb, "allreduce_wait_adamw_refresh_ms": ms_tail,
           "kernel_launches_per_step": launches, "cuda_graph": bool(tr.use_cuda_graph), "losses": losses[:5],
           "loss_decreases": bool(losses[-1] < losses[0]),
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
           "params_trained": int(sum(tr.buckets.sizes)), "dtype": "bf16 activations / gradients, fp32 master weights + AdamW"}
    del tr, cnet
    torch.cuda.empty_cache()
    return res


def _lib_launches():
    from posetraj_b200 import _lib
    return _lib.lib().pt_launch_count()


def library_baseline_leg(dev, steps=3):
    """The "library bar": the oracle's wiring as plain torch eager bf16 on this GPU (cuDNN / cuBLAS / SDPA, no fusion) —
    what the reference would run here, since it ships no Blackwell kernel.  Comparator only (imports oracle/)."""
    import torch
    from oracle.models import build_models
    from oracle.pipeline import denoise_step, make_inputs
    from oracle.scheduler import EulerKarrasOracle
    torch.manual_seed(0)
    with torch.device(dev):
        unet, cnet = build_models(seed=0, randomize_zero_convs=False)
    unet, cnet = unet.to(torch.bfloat16), cnet.to(torch.bfloat16)
    inp = {k: (v.to(dev, torch.bfloat16) if torch.is_tensor(v) else v) for k, v in make_inputs().items()}
    cond = torch.full((2, FRAMES, 3, LAT_H * 8, LAT_W * 8), -1.0, device=dev, dtype=torch.bfloat16)
    sched = EulerKarrasOracle()
    sched.set_timesteps(SAMPLING_STEPS, device=dev)
    times = []
    for i in range(steps + 2):
        sched._step_index = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        denoise_step(unet, cnet, sched, inp["latents"], 0, sched.timesteps[0], inp["image_latents"], inp["image_embeddings"],
                     cond, inp["added_time_ids"], inp["guidance"])
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    del unet, cnet
    torch.cuda.empty_cache()
    res = {"what": "torch 2.11 eager bf16 of the same wiring (cuDNN / cuBLAS / SDPA), same GPU, same shapes",
           "ms_per_step": ms, "value": 1000.0 / ms, "unit": UNIT}
    # the same bar for the configs[3] training step: torch autograd + torch.optim.AdamW (fused) on the bbox models in bf16
    try:
        from oracle.train import training_step
        torch.manual_seed(0)
        with torch.device(dev):
            unet, cnet = build_models(seed=0, bbox=True, randomize_zero_convs=False)   # (zero-convs at zero: same kernels, same time)
        unet, cnet = unet.to(torch.bfloat16).requires_grad_(False), cnet.to(torch.bfloat16).requires_grad_(True)
        opt = torch.optim.AdamW(cnet.parameters(), lr=1e-5, fused=True)
        g = torch.Generator(device=dev).manual_seed(11)
        B = 2
        rn = lambda *sh: torch.randn(*sh, device=dev, generator=g).to(torch.bfloat16)
        batch = dict(latents=rn(B, FRAMES, 4, LAT_H, LAT_W) * 0.9, noise=rn(B, FRAMES, 4, LAT_H, LAT_W),
                     sigmas=torch.tensor([1.3, 0.4], device=dev), image_embeddings=rn(B, 1, 1024),
                     trajectories=((torch.rand(B, FRAMES, 3, 8 * LAT_H, 8 * LAT_W, device=dev, generator=g) > 0.97).float() * 2 - 1).to(torch.bfloat16),
                     motion_values=torch.tensor([127.0, 90.0], device=dev))
        bbox = ((torch.rand(B, FRAMES, 3, 8 * LAT_H, 8 * LAT_W, device=dev, generator=g) > 0.98).float() * 2 - 1).to(torch.bfloat16)

        def cn(*a, **k):
            return cnet(*a, controlnet_bbox=bbox, **k)

        times = []
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = training_step(unet, cn, ran_idx=3, **batch)
            out["loss"].backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            e1.record()
            torch.cuda.synchronize()
            if i >= 1:
                times.append(e0.elapsed_time(e1))
        res["train_step"] = {"what": "torch eager bf16 autograd + fused AdamW of the configs[3] step (bbox ControlNet, 2 videos x 14 frames x 320x576)",
                             "ms_per_step": sum(times) / len(times), "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        del unet, cnet, opt, out
    except Exception as e:  # noqa: BLE001 - a comparator must not take the line down
        res["train_step"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    return res


def run_own(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from posetraj_b200 import _lib
    from posetraj_b200.config import SVDConfig
    from posetraj_b200.models import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel
    from posetraj_b200.pipeline import StableVideoDiffusionPipelineControlNet
    from posetraj_b200.roofline import step_flops

    cfg = SVDConfig()
    unet = UNetSpatioTemporalConditionControlNetModel.from_random(cfg, dev, seed=0)
    cnet = ControlNetSDVModel.from_random(cfg, dev, seed=0, faithful_zero_init=False)
    pipe = StableVideoDiffusionPipelineControlNet(unet=unet, controlnet=cnet)

    # ---- synthetic inputs of one video, in pinned host memory (each rank its own seed = its own video) ----------
    g = torch.Generator().manual_seed(1234 + rank)
    H, W = LAT_H * 8, LAT_W * 8
    lat_unit = torch.randn(1, FRAMES, 4, LAT_H, LAT_W, generator=g).pin_memory()
    img = torch.randn(1, 4, LAT_H, LAT_W, generator=g)
    image_latents = torch.cat([torch.zeros_like(img), img]).pin_memory()
    emb = torch.randn(1, 1, cfg.cross_attention_dim, generator=g)
    image_embeddings = torch.cat([torch.zeros_like(emb), emb]).pin_memory()
    from posetraj_b200.trajectory import rasterize_tracks
    tracks = torch.tensor(_synthetic_tracks(FRAMES), dtype=torch.int32).pin_memory()   # [K=1, F, 2] host

    def call_pipeline():
        # R1 on the GPU: tracks (host) -> [F,3,H,W] conditioning maps in [-1,1] (cv2-exact), then the sampling loop
        cond = rasterize_tracks(tracks, FRAMES, H, W, dev, output="f32")
        out = pipe(None, cond, height=H, width=W, num_frames=FRAMES, num_inference_steps=SAMPLING_STEPS,
                   latents=lat_unit, output_type="latent", image_embeddings=image_embeddings, image_latents=image_latents)
        return out.frames.to("cpu")

    # first call builds the plans, stages everything, captures the CUDA graph of one step
    lat_final = call_pipeline()
    if not torch.isfinite(lat_final).all():
        raise SystemExit("bench.py: non-finite latents")
    eng = pipe.engine_for(FRAMES, LAT_H, LAT_W, (H, W))
    assert eng.graph is not None
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps/s ----------------------------------------------------------------------------------
    def run_steps(n, start_at=0):
        for i in range(n):
            if (start_at + i) % SAMPLING_STEPS == 0:
                eng.reset()
            eng.graph.replay()

    run_steps(args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run_steps(args.steps, start_at=args.warmup)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total / 1e3)

    # ---- end to end through the public API with host buffers --------------------------------------------------------
    n_calls = max(1, round(args.steps / SAMPLING_STEPS))
    call_pipeline()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(n_calls):
        call_pipeline()
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([max(e0.elapsed_time(e1) / 1e3, wall)], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * n_calls * SAMPLING_STEPS / e2e_s
    h2d = sum(x.numel() * x.element_size() for x in (tracks, lat_unit, image_latents, image_embeddings)) + 26 * 4 + 2 * 3 * 4
    d2h = lat_final.numel() * lat_final.element_size()

    line = None
    if rank == 0:
        # ---- per-kernel-class CUDA-event times of one step (eager replay of the same launch list) -------------------
        eng.reset()
        classes = time_op_classes(eng.step_ops, stream)
        peaks = _peaks()
        gem = classes["gemm"]
        gemm_tf = gem["flops"] / (gem["ms"] * 1e-3) / 1e12
        total_ms = sum(c["ms"] for c in classes.values())
        ess, _ = step_flops(cfg, frames=FRAMES, h=LAT_H, w=LAT_W, essential=True)
        roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (linear / implicit-GEMM conv, all launches of one step)",
                    "achieved": gemm_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": gemm_tf / peaks["tf_sustained"],
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})",
                    "launches_per_step": int(round(gem["launches"])), "avg_launch_ms": gem["ms"] / gem["launches"],
                    "share_of_step": gem["ms"] / total_ms, "traffic": _ncu_traffic("gemm"),
                    "algorithmic_bytes_per_launch": gem["bytes"] / gem["launches"],
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the pt_gemm launches of one step, from the committed ncu capture profiles/r3z_dram_traffic.json (bytes); algorithmic_bytes_per_launch = A rows once + weights + epilogue operands + outputs"}
        per_class = {}
        for k, c in sorted(classes.items(), key=lambda kv: -kv[1]["ms"]):
            e = {"ms": round(c["ms"], 4), "launches": int(round(c["launches"])), "share": round(c["ms"] / total_ms, 4)}
            if c["flops"] > 0:
                e["tflops"] = round(c["flops"] / (c["ms"] * 1e-3) / 1e12, 1)
            if c["bytes"] > 0:
                e["alg_gbs"] = round(c["bytes"] / (c["ms"] * 1e-3) / 1e9, 1)
                e["frac_hbm"] = round(e["alg_gbs"] / peaks["hbm"], 3)
            per_class[k] = e
        step_tf = ess / (ms_per_step * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "weights": "random-init SVD-shaped UNet 1524.6M + ControlNetSDV 682.0M (zero-convs re-randomised)",
                       "sharding": "one video (CFG pair) per GPU, no data-path collective" if world > 1 else "single GPU",
                       "l2": "per-step working set (4.4 GB bf16 weights + activations) >> 126 MB L2; no explicit flush",
                       "videos_per_min": value * 60.0 / SAMPLING_STEPS},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / SAMPLING_STEPS,
                    "d2h_bytes_per_step": d2h / SAMPLING_STEPS, "calls": n_calls,
                    "api": "rasterize_tracks(host tracks) + StableVideoDiffusionPipelineControlNet.__call__(host tensors, output_type='latent') + .cpu()",
                    "videos_per_min": e2e_value * 60.0 / SAMPLING_STEPS},
            "gpu_launches": int(eng.launches_per_step * args.steps + 3 * ((args.steps + SAMPLING_STEPS - 1) // SAMPLING_STEPS)),
            "roofline": roofline,
            "step_roofline": {"essential_tflop_per_step": ess / 1e12, "achieved_tflops": step_tf,
                              "frac_of_sustained_peak": step_tf / peaks["tf_sustained"]},
            "kernel_classes": per_class,
            "lib_launch_count": int(_lib.lib().pt_launch_count()),
        }
        if world == 1:
            line["full_pipeline"] = full_pipeline_leg(unet, cnet, tracks, dev)
    # ---- extra legs: every rank takes part (collectives); rank 0 reports ------------------------------------------
    extras = {}
    if not args.no_legs:
        leg_steps = max(5, min(args.steps, 20))

        def leg(name, fn, *a):
            try:
                extras[name] = fn(*a)
            except Exception as e:  # noqa: BLE001 - an extra must not take the contract line down
                extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}

        leg("configs2_cam", cam_leg, cfg, unet, dev, rank, world, leg_steps, barrier)
        leg("configs3_train_blocks", train_dp_leg, cfg, dev, rank, world, barrier)
        leg("configs3_train_step", train_step_leg, cfg, unet, dev, rank, world, barrier)
        if world >= 2 and world % 2 == 0:
            leg("cfg_split", cfg_split_leg, cfg, unet, cnet, dev, rank, world, leg_steps, barrier, ms_per_step,
                lat_final if rank == 0 else None)
        if world >= 2:
            leg("frame_sharded", frame_sharded_leg, cfg, unet, cnet, dev, rank, world, leg_steps, barrier)
        if world == 1:
            leg("library_baseline", library_baseline_leg, dev)
    if rank == 0:
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            v, t_step, cores, sample = cpu_oracle_steps_per_sec(1, 0, full=True)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the extra legs (cam, CFG split, frame sharding, library baseline)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
