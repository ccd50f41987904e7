#!/usr/bin/env python
"""bench.py — PVR frames/sec embedded (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload uber34x3|conv5]

One step = the embedding hot path over one batch of synthetic observations, `--passes` (default 2) encoder passes of
`obs_per_pass` observations each, so that the K timed steps last >= 2 s (one pass of 768 observations is 52 ms):
uint8 (B, 224, 224, 3n) -> fused preprocessing kernel -> ResNet-50 trunk(s) (tcgen05 implicit GEMM) -> (B, n*O) fp32.
Default workload = BASELINE.json configs[1]: moco_aug_uber_34 (layer3 + layer4 compressed taps, two independent
ResNet-50 trunks as the reference computes them, src/embeddings.py:44-57,225-229), 3-frame observations, bf16.

`value`   frames/s with the batch already resident in HBM (frames = B * n per step, summed over ranks).
`e2e`     the same metric through the public API with HOST buffers: pinned uint8 observations -> H2D ->
          EmbeddingNet.embed -> D2H of the embeddings, every step inside the timed region.
`roofline` dominant kernel = conv_gemm_kernel (all tcgen05 conv launches of a step): algorithmic conv FLOPs of the
          step / summed device time of those launches (CUDA events between ops on the launch stream, measured here).
`bc`, `clip_b16`, `finetune`: the other BASELINE configs measured in the same run, each with its own roofline figure
          and CPU baseline: BC steps/s (configs[4], + a weak-scaling line under torchrun), CLIP ViT-B/16 at 1024
          images per GPU (configs[2]), end-to-end finetuning steps/s (configs[3]).
`cpu_baseline` the oracle's CPU port of the same workload on this host's cores (bounded sample), rank 0, N=1 only.
--impl reference: the reference's own CPU implementation of the path (oracle port: the reference is Python and does
          not travel to the GPU box) on all host threads, same JSON schema.
"""
import argparse
import atexit
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (embedding name, frames per observation, observations per step per GPU)
    # 768 observations = 2304 frames per encoder pass: larger passes amortise the per-launch cost of the 91 conv
    # kernels (profiles/r02_sweep_pass_size_conv5.txt; 192 -> 768: +2 % at the power-capped clock of this pool)
    "uber34x3": ("moco_aug_uber_34", 3, 768),
    "conv5": ("moco_aug", 1, 512),
    "clip_b16": ("clip_vit_b16", 1, 1024),  # BASELINE configs[2]: CLIP-architecture ViT-B/16, batch 1024 per GPU
    "clip_b32": ("clip_vit", 1, 1024),      # the reference's actual `clip_vit` (ViT-B/32)
    "mae_base": ("mae_base", 1, 1024),      # MAE ViT-B/16 (bicubic preprocessing, erf GELU)
    "mae_large": ("mae_large", 1, 512),     # MAE ViT-L/16
    # SURVEY §8(f)-4, measured for DESIGN.md only (no BASELINE config names them)
    "mae_huge": ("mae_huge", 1, 256),       # MAE ViT-H/14: 257 tokens, 16 heads of 80 (attention_mma.cu)
    "clip_rn50": ("clip_rn50", 1, 512),     # CLIP ModifiedResNet + attention pool
    "maskrcnn_l3": ("maskrcnn_l3", 1, 512), # detectron2 R50-C4 through res4 + 1024 -> 11 compression block
}
VARIANTS = {"moco_aug": ["conv5"], "moco_aug_uber_34": ["l3", "l4"]}
CLIP_PATCH = {"clip_vit_b16": 16, "clip_vit": 32}
MAE_NAMES = ("mae_base", "mae_large", "mae_huge")
# encoders driven by a runner object (whole forward timed as one unit): ViTRunner, CLIPRNRunner
VIT_NAMES = tuple(CLIP_PATCH) + MAE_NAMES + ("clip_rn50",)
GFLOP_PER_FRAME = {"moco_aug": 8.174, "moco_aug_uber_34": 14.96,  # SURVEY.md §8(d), convs only
                   "clip_vit_b16": 35.13, "clip_vit": 8.818,
                   "mae_base": 35.13, "mae_large": 123.1,  # 2 x (patch embed + 12W^2 S + 2 S^2 W per layer) MACs
                   "mae_huge": 334.6, "clip_rn50": 12.02,  # runner.flops_per_image
                   "maskrcnn_l3": 6.636}                   # same convolutions as the MoCo layer-3 variant


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms; `mark_start()` / `mark_end()` bracket the timed region
    and only samples read inside it count (the samples around it are idle: with the load gone the power cap lifts and
    the SM clock jumps to its maximum, which is not the clock the measurement ran at)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None   # rows: (host time when read, fields)
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            atexit.register(self._kill)  # never leave the polling nvidia-smi behind, whatever ends the run
        except OSError:
            self.proc = None

    def _kill(self):
        if self.proc is not None and self.proc.poll() is None:
            self.proc.terminate()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    @staticmethod
    def summarise(rows, t0, t1):
        """rows: (time, fields). Samples read inside [t0 + 50 ms, t1 + 50 ms] (a sample describes the ~50 ms before
        it was printed); all samples if the window is empty or was never marked."""
        ok = [(t, r) for t, r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        inside = [(t, r) for t, r in ok if t0 is not None and t1 is not None and t0 + 0.05 <= t <= t1 + 0.05]
        use = inside or ok
        sm = [float(r[0]) for _, r in use]
        mx = [float(r[1]) for _, r in ok if r[1].replace(".", "").isdigit()]
        reasons = set()
        for _, r in use:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(inside)}

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        return self.summarise(list(self.rows), self.t0, self.t1)


def make_observations(n, n_frames, seed):
    """Structured synthetic uint8 frames (smooth gradients + blocks + noise, not iid noise), tiled from a small seeded
    set to bound generation time. Generated here: the GPU arm never imports oracle/."""
    rng = np.random.default_rng(seed)
    c = 3 * n_frames
    yy, xx = np.meshgrid(np.arange(224, dtype=np.float32), np.arange(224, dtype=np.float32), indexing="ij")
    base = np.empty((16, 224, 224, c), dtype=np.uint8)
    for i in range(16):
        for ch in range(c):
            fx, fy, ph = rng.uniform(0.01, 0.08, 3)
            img = 110 + 70 * np.sin(fx * xx + 3 * ph) * np.cos(fy * yy + ph) + rng.normal(0, 12, (224, 224))
            y0, x0 = rng.integers(0, 160, 2)
            img[y0:y0 + 48, x0:x0 + 64] += rng.uniform(-60, 60)
            base[i, :, :, ch] = np.clip(img, 0, 255).astype(np.uint8)
    reps = (n + 15) // 16
    obs = np.concatenate([np.roll(base, shift=7 * r, axis=2) for r in range(reps)])[:n]
    return np.ascontiguousarray(obs)


def oracle_parts(name, seed=1):
    """Synthetic weights for the CPU legs (cpu_baseline / --impl reference only)."""
    from oracle import restate
    if name in CLIP_PATCH:
        from oracle import restate_vit
        return restate_vit.vit_state(CLIP_PATCH[name], seed)
    if name in MAE_NAMES:
        from oracle import restate_mae
        return restate_mae.mae_state(name, seed)
    return [(v, restate.resnet50_state(v, seed + i)) for i, v in enumerate(VARIANTS[name])]


def oracle_embed(name, parts, obs):
    """CPU port of the embedding loop (mini-batches of 64 like main_bc_1.py:130)."""
    from oracle import restate
    if name in CLIP_PATCH:
        from oracle import restate_vit
        return np.concatenate([restate_vit.embedding_forward(parts, obs[i:i + 64]) for i in range(0, len(obs), 64)])
    if name in MAE_NAMES:
        from oracle import restate_mae
        return np.concatenate([restate_mae.embedding_forward(parts, name, obs[i:i + 64]) for i in range(0, len(obs), 64)])
    return restate.embed_observations(parts, obs, batch_size=64)


def build_net(name, device):
    from pvr_habitat_b200.embeddings import EmbeddingNet
    from pvr_habitat_b200.vision_models.moco import allow_random_init
    # random-init weights of the named architecture from the package's own parameter holders (there is no network
    # for checkpoints); BatchNorm statistics are randomised so that the folded scale / bias are not trivial
    torch.manual_seed(1)
    with allow_random_init():
        net = EmbeddingNet(name)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for k, v in net.embedding.state_dict().items():
            if k.endswith("running_var"):
                v.copy_(torch.empty(v.shape).uniform_(0.75, 1.25, generator=g))
            elif k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            elif "bn" in k.split(".")[-2:][0] and k.endswith("weight") and v.dim() == 1:
                v.copy_(torch.empty(v.shape).uniform_(0.4, 0.9, generator=g))
    net.invalidate()
    return net


def cpu_port_frames_per_s(name, n_frames, n_obs, threads):
    """Oracle (CPU port of the reference path) on a bounded sample: returns (frames/s, seconds)."""
    from oracle import restate
    torch.set_num_threads(threads)
    parts = oracle_parts(name)
    obs = make_observations(n_obs, n_frames, 5)
    oracle_embed(name, parts, obs[:2])  # warm-up
    t0 = time.perf_counter()
    oracle_embed(name, parts, obs)
    dt = time.perf_counter() - t0
    return n_obs * n_frames / dt, dt


BC_CFG = dict(n=65536, D=2048, T=64, B=128, A=3)  # BASELINE configs[4]: global batch T*B = 8192, batch_norm=True
BC_GFLOP_PER_STEP = 979.0  # fwd+bwd 119.6 MFLOP/sample x 8192
BC_CPU_STEPS = 8           # CPU baseline sample: about 6-10 s on 16 host cores


def bc_dataset(seed=7):
    """Synthetic pre-embedded BC dataset (obs (n, D) float32, action, done, -): ReLU-like features whose action is a
    noisy linear function of the observation, episodes of random length."""
    rng = np.random.default_rng(seed)
    n, D, A = BC_CFG["n"], BC_CFG["D"], BC_CFG["A"]
    obs = np.maximum(rng.standard_normal((n, D), dtype=np.float32), 0.0)
    w = rng.standard_normal((D, A), dtype=np.float32) / np.sqrt(D)
    action = (obs @ w + 0.3 * rng.standard_normal((n, A), dtype=np.float32)).argmax(1).astype(np.int64)
    done = rng.random(n) < 1.0 / 200.0
    done[0] = True
    return obs, action, done, None


def bench_bc(steps, warmup, world, dist, host_batches, batch_size=None):
    """BC train steps/s on the CUDA policy path. Default: strong scaling, the global batch T*B = 8192 is split over the
    ranks; `batch_size` = 128 * world gives the weak-scaling line (128 sequences per rank)."""
    import random
    from pvr_habitat_b200.bc import BCTrainer
    from pvr_habitat_b200.models import PolicyNet
    obs, action, done, _ = bc_dataset()
    torch.manual_seed(1)
    random.seed(1)
    if host_batches and world > 1:  # the host-side gather of every rank is multi-threaded: share the cores
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    net = PolicyNet((BC_CFG["D"],), BC_CFG["A"], batch_norm=True).cuda().train()
    tr = BCTrainer(net, obs, action, done, batch_size or BC_CFG["B"], BC_CFG["T"], 10 ** 9, host_batches=host_batches)
    for _ in range(warmup):
        tr.step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.step()
        if host_batches:
            loss.item()  # the reference reads the loss on the host (main_bc_2.py:245)
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return steps / (ms / 1e3), ms / steps, float(loss)


def cpu_port_bc_steps_per_s(threads):
    """Oracle CPU port of BC steps at the same shape (bounded sample: BC_CPU_STEPS steps after a small warm-up)."""
    from oracle import restate_policy as rp
    torch.set_num_threads(threads)
    obs, action, done, _ = bc_dataset()
    sd = rp.init_policy_state(BC_CFG["D"], BC_CFG["A"], True, 1)
    rp.bc_train(sd, obs, action, done, 4, 8, 1, 10 ** 9, True)  # warm-up at a tiny shape
    t0 = time.perf_counter()
    rp.bc_train(sd, obs, action, done, BC_CFG["T"], BC_CFG["B"], BC_CPU_STEPS, 10 ** 9, True)
    dt = time.perf_counter() - t0
    return BC_CPU_STEPS / dt, dt


CLIP_B16_IMAGES = 1024     # BASELINE configs[2]: batch 1024 per GPU
FT_CFG = dict(T=100, B=16, hw=64, n_frames=2, n=16384)  # n >= B_global + (B_global - 1)(T - 1) at 8 GPUs  # BASELINE configs[3]: obs (T=100, B=16*G, 64, 64, 6) uint8
# conv trunk 8.04 MFLOP/frame forward (5 x 3x3/s2 convs on 64x64), policy trunk at D = 256: 36.2 MFLOP/sample forward;
# forward + backward = 3 x forward
FT_GFLOP_PER_STEP_PER_SEQ = 3 * (2 * 8.04 + 36.2) * 100 / 1e3


def bench_clip_b16(world, dist, rank, want_cpu):
    """BASELINE configs[2]: CLIP-architecture ViT-B/16 (random init), 1024 images of 224x224 per GPU per pass, frames
    sharded over the ranks (no collective). Whole encoder forward against the sustained bf16 peak."""
    net = build_net("clip_vit_b16", torch.device("cuda", torch.cuda.current_device()))
    n = CLIP_B16_IMAGES
    net.max_images_per_pass = n
    obs = [torch.from_numpy(make_observations(n, 1, 300 + rank * 4 + i)).cuda() for i in range(2)]  # 2 x 154 MB > L2
    out = torch.empty(n, net.out_size, dtype=torch.float32, device="cuda")
    for i in range(3):
        net.embed(obs[i % 2], 1, out=out)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    passes = 48  # ~2 s
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(passes):
        net.embed(obs[i % 2], 1, out=out)
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    v = world * n * passes / (ms / 1e3)
    peaks = load_peaks()
    tf = v / world * GFLOP_PER_FRAME["clip_vit_b16"] / 1e3
    res = {"metric": "pvr_frames_per_sec_embedded", "value": v, "unit": "frames/s", "scaling": "weak",
           "ms_per_pass": ms / passes, "passes": passes,
           "config": {"workload": "clip_vit_b16: CLIP-architecture ViT-B/16, 224x224 uint8 frames, random-init weights "
                                  "(BASELINE configs[2])", "images_per_pass_per_gpu": n,
                      "l2": "2 distinct input batches of 154 MB; activations rewritten every pass"},
           "roofline": {"kernel": "whole ViT forward per GPU: tcgen05 GEMMs (cta_group::2) + vit_attention_kernel + "
                                  "LayerNorm + preprocessing", "bound": "tensor", "achieved": tf,
                        "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_sustained"],
                        "algorithmic_gflop_per_frame": GFLOP_PER_FRAME["clip_vit_b16"], "traffic": None}}
    del net, obs, out
    torch.cuda.empty_cache()
    if want_cpu:
        threads = os.cpu_count() or 1
        vc, dt = cpu_port_frames_per_s("clip_vit_b16", 1, 128, threads)
        res["cpu_baseline"] = {"value": vc, "unit": "frames/s", "cores": threads, "kind": "port",
                               "sample": f"128 images, mini-batch 64, torch CPU fp32 oracle port, {dt:.1f} s"}
    return res


def finetune_dataset(seed=9):
    """Raw 64x64 two-frame uint8 observations (current | goal image) with actions that follow the mean colour of the
    first frame's centre, episodes of random length."""
    rng = np.random.default_rng(seed)
    n, hw, c = FT_CFG["n"], FT_CFG["hw"], 3 * FT_CFG["n_frames"]
    yy, xx = np.meshgrid(np.arange(hw, dtype=np.float32), np.arange(hw, dtype=np.float32), indexing="ij")
    base = np.empty((64, hw, hw, c), dtype=np.uint8)
    for i in range(64):
        for ch in range(c):
            fx, fy, ph = rng.uniform(0.03, 0.25, 3)
            img = 110 + 70 * np.sin(fx * xx + 3 * ph) * np.cos(fy * yy + ph) + rng.normal(0, 12, (hw, hw))
            base[i, :, :, ch] = np.clip(img, 0, 255).astype(np.uint8)
    obs = np.concatenate([np.roll(base, 5 * r, axis=2) for r in range(n // 64)])
    feat = obs[:, 16:48, 16:48, :3].reshape(n, -1, 3).mean(1)
    action = np.argmax(feat + 8 * rng.standard_normal(feat.shape), 1).astype(np.int64)
    done = rng.random(n) < 1.0 / 100.0
    return np.ascontiguousarray(obs), action, done


def cpu_port_finetune_steps_per_s(threads, obs, action, done, n_cpu=2):
    """Oracle CPU port of the finetuning loop (main_bc_finetune.py:167-208) at the bench shape, bounded sample."""
    from oracle import restate_policy as rp
    torch.set_num_threads(threads)
    sd = rp.init_policy_conv_state(FT_CFG["hw"], FT_CFG["n_frames"], 3, True, 1)
    t0 = time.perf_counter()
    rp.bc_train_conv(sd, obs, action, done, FT_CFG["T"], FT_CFG["B"], n_cpu, 10 ** 9)
    dt = time.perf_counter() - t0
    return n_cpu / dt, dt, n_cpu


def bench_finetune(world, dist, want_cpu):
    """BASELINE configs[3]: PolicyNetWithConv((64, 64, 6), 3, batch_norm=True) trained end to end with BC, 16
    sequences of 100 steps per GPU (weak scaling: global batch 16 * G), NCCL gradient all-reduce."""
    import random
    from pvr_habitat_b200.bc import BCTrainer
    from pvr_habitat_b200.models import PolicyNetWithConv
    obs, action, done = finetune_dataset()
    torch.manual_seed(1)
    random.seed(1)
    net = PolicyNetWithConv((FT_CFG["hw"], FT_CFG["hw"], 3 * FT_CFG["n_frames"]), 3, batch_norm=True).cuda().train()
    tr = BCTrainer(net, obs, action, done, FT_CFG["B"] * world, FT_CFG["T"], 10 ** 9)
    for _ in range(4):
        tr.step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    steps = 60
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.step()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    v = steps / (ms / 1e3)
    peaks = load_peaks()
    tf = v * FT_GFLOP_PER_STEP_PER_SEQ * FT_CFG["B"] / 1e3  # per GPU
    res = {"metric": "bc_finetune_steps_per_sec", "value": v, "unit": "steps/s", "ms_per_step": ms / steps,
           "steps": steps, "scaling": "weak", "frames_per_sec": v * FT_CFG["T"] * FT_CFG["B"] * world * FT_CFG["n_frames"],
           "config": {"workload": "PolicyNetWithConv((64, 64, 6), 3, batch_norm=True), main_bc_finetune loop "
                                  "(BASELINE configs[3])", "unroll_length": FT_CFG["T"],
                      "batch_size_per_gpu": FT_CFG["B"], "global_batch_size": FT_CFG["B"] * world,
                      "parallelism": f"dp{world}: sequences split over ranks, BN sums + gradients all-reduced"},
           "roofline": {"kernel": "conv trunk forward / backward GEMMs + policy trunk + persistent LSTM recurrence",
                        "bound": "latency (T = 100 serial LSTM steps, host-side program rebuild); tensor peak reported",
                        "achieved": tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                        "frac": tf / peaks["bf16_sustained"],
                        "algorithmic_gflop_per_step_per_gpu": FT_GFLOP_PER_STEP_PER_SEQ * FT_CFG["B"], "traffic": None},
           "last_loss": float(loss)}
    del tr, net
    torch.cuda.empty_cache()
    if want_cpu:
        threads = os.cpu_count() or 1
        v_cpu, dt, n_cpu = cpu_port_finetune_steps_per_s(threads, obs, action, done)
        res["cpu_baseline"] = {"value": v_cpu, "unit": "steps/s", "cores": threads, "kind": "port",
                               "sample": f"{n_cpu} steps at the same shape (T=100, B=16, 64x64x6), torch CPU fp32 "
                                         f"oracle port, {dt:.1f} s"}
    return res


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads, same schema."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, n_frames, _ = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    from oracle import restate
    torch.set_num_threads(threads)
    parts = oracle_parts(name)
    per_step = 8  # observations per step: bounded so that K steps end within minutes
    obs = make_observations(per_step, n_frames, 5)
    for _ in range(max(1, min(args.warmup, 1))):
        oracle_embed(name, parts, obs[:2])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_embed(name, parts, obs)
    dt = time.perf_counter() - t0
    v = args.steps * per_step * n_frames / dt
    line = {
        "impl": "reference", "metric": "pvr_frames_per_sec_embedded", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name} x {n_frames}-frame 224x224 uint8 observations (BASELINE configs[1])"
                   if args.workload == "uber34x3" else f"{name} 224x224 uint8 frames",
                   "obs_per_step": per_step, "frames_per_step": per_step * n_frames},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x {per_step} observations x {n_frames} frames, torch CPU fp32, "
                                   "oracle/restate.py (the reference is Python and is not present on the GPU box)"},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="uber34x3", choices=list(WORKLOADS))
    ap.add_argument("--obs-per-step", type=int, default=0, help="observations per encoder PASS (per GPU)")
    ap.add_argument("--passes", type=int, default=2, help="encoder passes per step (timed region >= 2 s)")
    ap.add_argument("--no-extra", action="store_true", help="skip the clip_b16 / finetune measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bc", action="store_true", help="skip the BC steps/s measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    name, n_frames, obs_per_pass = WORKLOADS[args.workload]
    if args.obs_per_step:
        obs_per_pass = args.obs_per_step
    passes = max(1, args.passes)
    obs_per_step = obs_per_pass * passes
    frames_per_step = obs_per_step * n_frames
    net = build_net(name, torch.device("cuda", local))
    net.max_images_per_pass = obs_per_pass * n_frames  # `embed` cuts the step's batch into equal passes
    width = n_frames * net.out_size
    # started early (nvidia-smi needs a few hundred ms to print its first sample); only the samples read inside the
    # timed region are used
    sampler = ClockSampler(local) if rank == 0 else None

    # Inputs: a rotation of distinct batches larger than L2 in total (126 MB), so no step re-reads its input from L2.
    bytes_per_batch = obs_per_step * 224 * 224 * 3 * n_frames
    n_rot = max(2, -(-3 * 126 * 2 ** 20 // bytes_per_batch))
    base = make_observations(min(obs_per_step, 256), n_frames, 100 + rank * 16)
    host = []
    for i in range(n_rot):  # distinct batches: rolled copies of a seeded set (generation time is bounded)
        reps = -(-obs_per_step // len(base))
        b = np.concatenate([np.roll(base, 11 * (i * reps + r) + 3, axis=1) for r in range(reps)])[:obs_per_step]
        host.append(torch.from_numpy(np.ascontiguousarray(b)).pin_memory())
    dev = [h.cuda(non_blocking=True) for h in host]
    out = torch.empty(obs_per_step, width, dtype=torch.float32, device="cuda")
    out_host = torch.empty(obs_per_step, width, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for i in range(warmup):
            fn(i)
        barrier()
        if sampler is not None:
            sampler.mark_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        if sampler is not None:
            sampler.mark_end()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput
    def step_resident(i):
        net.embed(dev[i % n_rot], n_frames, out=out)

    ms = timed(step_resident, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if sampler else None
    value = world * frames_per_step * args.steps / (ms / 1e3)

    # ---- end to end: pinned host -> H2D -> embed -> D2H, every step. Two device buffers: the H2D copy of step i+1 is
    # issued on a copy stream while step i computes (what a user streaming frames from host memory does); every step
    # still waits for its own embeddings on the host.
    copy_stream = torch.cuda.Stream()
    dbuf = [torch.empty_like(dev[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    issued = set()

    def prefetch(i):
        if i in issued:
            return
        issued.add(i)
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])  # the embed that read this buffer two steps ago has finished
            dbuf[b].copy_(host[i % n_rot], non_blocking=True)
            ready[b].record(copy_stream)

    def step_e2e(i):
        prefetch(i)
        prefetch(i + 1)
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[i % 2])
        net.embed(dbuf[i % 2], n_frames, out=out)
        consumed[i % 2].record(cur)
        out_host.copy_(out, non_blocking=True)
        cur.synchronize()  # the caller consumes the embeddings (numpy) every step

    ms_e2e = timed(step_e2e, args.steps, 3)
    e2e_value = world * frames_per_step * args.steps / (ms_e2e / 1e3)

    # ---- second headline metric: BC train steps/s (all ranks: data parallel, NCCL gradient all-reduce)
    bc = None
    if not args.no_bc:
        bc_steps = max(600, args.steps)  # >= 2 s of timed steps at ~3 ms per step
        # 6 warm-up steps: 3 eager ones, the CUDA-graph capture of the whole step, 2 replays
        v, ms_bc, last_loss = bench_bc(bc_steps, 6, world, dist, host_batches=False)
        v_e2e, ms_bc_e2e, _ = bench_bc(bc_steps // 2, 6, world, dist, host_batches=True)
        bc = {"metric": "bc_train_steps_per_sec", "value": v, "unit": "steps/s", "ms_per_step": ms_bc,
              "scaling": "strong", "steps": bc_steps,
              "config": {"workload": "PolicyNet((2048,), 3, batch_norm=True) on pre-embedded observations "
                                     "(BASELINE configs[4]), RMSprop + clip 40, LambdaLR",
                         "global_batch_rows": BC_CFG["T"] * BC_CFG["B"], "unroll_length": BC_CFG["T"],
                         "batch_size": BC_CFG["B"], "dataset_rows": BC_CFG["n"],
                         "parallelism": f"dp{world}: sequences split over ranks, BN sums + gradients all-reduced"},
              "tflops_effective": v * BC_GFLOP_PER_STEP / 1e3,
              "frac_of_bf16_sustained": v * BC_GFLOP_PER_STEP / 1e3 / load_peaks()["bf16_sustained"] / world,
              "roofline": {"kernel": "whole BC step per GPU (tcgen05 GEMMs + persistent LSTM recurrence + fused "
                                     "optimizer)", "bound": "latency of 4 x T serial LSTM steps; tensor peak reported",
                           "achieved": v * BC_GFLOP_PER_STEP / 1e3 / world, "peak": load_peaks()["bf16_sustained"],
                           "unit": "TFLOP/s",
                           "frac": v * BC_GFLOP_PER_STEP / 1e3 / world / load_peaks()["bf16_sustained"],
                           "algorithmic_gflop_per_step": BC_GFLOP_PER_STEP, "traffic": None},
              "e2e": {"value": v_e2e, "unit": "steps/s", "ms_per_step": ms_bc_e2e,
                      "h2d_bytes_per_step": BC_CFG["T"] * BC_CFG["B"] // world * (BC_CFG["D"] * 4 + 8 + 1),
                      "d2h_bytes_per_step": 4},
              "last_loss": last_loss}
        if world > 1:  # weak scaling: 128 sequences per rank, so the cost of the collectives is visible on its own
            vw, ms_w, _ = bench_bc(bc_steps // 2, 6, world, dist, host_batches=False, batch_size=BC_CFG["B"] * world)
            bc["weak"] = {"value": vw, "unit": "steps/s", "ms_per_step": ms_w, "scaling": "weak",
                          "global_batch_rows": BC_CFG["T"] * BC_CFG["B"] * world,
                          "rows_per_sec": vw * BC_CFG["T"] * BC_CFG["B"] * world}
    extras = {}
    if not args.no_extra:
        want_cpu = world == 1 and rank == 0 and not args.no_cpu_baseline
        if args.workload != "clip_b16":
            extras["clip_b16"] = bench_clip_b16(world, dist, rank, want_cpu)
        extras["finetune"] = bench_finetune(world, dist, want_cpu)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- per-kernel roofline (rank 0): CUDA events between ops on the launch stream
    enc = net.encoder()
    peaks = load_peaks()
    if name in VIT_NAMES:
        # ViT: GEMM-dominated; the whole encoder forward (GEMMs + LayerNorm + attention) is timed as one unit
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        enc.bind(obs_per_pass * n_frames)
        net.transforms.run(dev[0][:obs_per_pass], n_frames, enc.slot0, enc.input_format, True)
        enc.forward(out[:obs_per_pass], net.out_size)
        e0.record()
        for r in range(3):
            enc.forward(out[:obs_per_pass], net.out_size)
        e1.record()
        torch.cuda.synchronize()
        fwd_ms = e0.elapsed_time(e1) / 3 * passes
        achieved = enc.flops_per_image * frames_per_step / (fwd_ms / 1e3) / 1e12
        roofline = {"kernel": "ViT encoder forward: conv_gemm_kernel (tcgen05 GEMMs) + vit_attention_kernel + LayerNorm",
                    "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_sustained"], "traffic": None,
                    "peak_source": peaks["source"] + ", sustained bf16",
                    "algorithmic_gflop_per_frame": enc.flops_per_image / 1e9, "forward_ms_per_step": fwd_ms}
        n_launch = passes * (1 + enc.launches_per_forward())
        roofline["frac_whole_step"] = enc.flops_per_image * frames_per_step / (ms / args.steps / 1e3) / 1e12 / \
            peaks["bf16_sustained"]
    else:
        roofline, n_launch = resnet_roofline(net, enc, dev, n_rot, n_frames, obs_per_pass, passes, out, peaks, ms, args)
    return finish(args, world, rank, dist, name, n_frames, obs_per_step, frames_per_step, width, n_rot,
                  bytes_per_batch, value, ms, e2e_value, ms_e2e, clocks, roofline, n_launch, bc, extras, passes)


def load_traffic():
    """DRAM bytes per frame of every kernel family, from the ncu capture kept under profiles/ (tools/ncu_traffic.py:
    dram__bytes_read.sum + dram__bytes_write.sum per launch of one default-workload pass); None if absent."""
    p = os.path.join(ROOT, "profiles", "r02_dram_traffic_uber34x3.json")
    return json.load(open(p)) if os.path.exists(p) else None


def resnet_roofline(net, enc, dev, n_rot, n_frames, obs_per_pass, passes, out, peaks, ms, args):
    """Per-kernel figures of ONE pass (obs_per_pass observations), scaled to the step of `passes` passes."""
    conv_ms, conv_flops, other_ms = [], 0.0, []
    reps = 5
    frames_per_pass = obs_per_pass * n_frames
    frames_per_step = frames_per_pass * passes
    enc.bind(frames_per_pass)
    for r in range(reps + 1):
        net.transforms.run(dev[r % n_rot][:obs_per_pass], n_frames, enc.slot0, enc.input_format, True)
        op_ms = enc.forward_timed(out[:obs_per_pass], net.out_size)
        if r == 0:
            continue  # warm
        conv_ms.append(sum(t for t, m in zip(op_ms, enc.op_meta) if m["kind"] == 1))
        other_ms.append(sum(t for t, m in zip(op_ms, enc.op_meta) if m["kind"] != 1))
    conv_flops = sum(m["flops_per_image"] for m in enc.op_meta if m["kind"] == 1) * frames_per_pass
    n_conv = sum(1 for m in enc.op_meta if m["kind"] == 1)
    conv_t = float(np.mean(conv_ms)) / 1e3
    achieved = conv_flops / conv_t / 1e12
    step_s = ms / args.steps / 1e3
    whole = conv_flops * passes / step_s / 1e12  # the same FLOPs over the driver-timed step (everything included)
    # preprocessing kernel timed alone (HBM bound)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(10):
        net.transforms.run(dev[r % n_rot][:obs_per_pass], n_frames, enc.slot0, enc.input_format, True)
    e1.record()
    torch.cuda.synchronize()
    pre_ms = e0.elapsed_time(e1) / 10
    alg_bytes = 224 * 224 * 3 + 224 * 224 * 3 * 2      # uint8 frame in, bf16 (3, 224, 224) out
    fmt_bytes = net.transforms.format_bytes_per_frame(enc.input_format)  # what the chosen stem format moves
    traffic = load_traffic()
    tr_conv = tr_pre = None
    if traffic is not None:
        tr_conv = int(traffic["conv_bytes_per_frame"] * frames_per_pass)
        tr_pre = int(traffic["preprocess_bytes_per_frame"] * frames_per_pass)
    roofline = {
        "kernel": "conv_gemm_kernel + conv3x3_patch_kernel + conv_b2b_kernel (tcgen05 implicit GEMM, %d conv launches "
                  "per pass of %d frames)" % (n_conv, frames_per_pass),
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_sustained"],
        "frac_whole_step": whole / peaks["bf16_sustained"], "achieved_whole_step": whole,
        "frac_of_burst_peak": achieved / peaks["bf16_burst"],
        # per pass like `achieved`: sum over the conv launches of one pass (ncu, profiles/r02_dram_traffic_*.json)
        "traffic": tr_conv, "traffic_source": "profiles/r02_dram_traffic_uber34x3.json" if traffic else None,
        "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
        "algorithmic_gflop_per_frame": conv_flops / frames_per_pass / 1e9,
        "conv_ms_per_pass": conv_t * 1e3, "conv_share_of_step": conv_t * 1e3 * passes / (ms / args.steps),
        "other_ops_ms_per_pass": float(np.mean(other_ms)),
        "preprocess": {"bound": "hbm", "ms_per_pass": pre_ms,
                       "achieved": alg_bytes * frames_per_pass / (pre_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
                       "unit": "GB/s", "frac": alg_bytes * frames_per_pass / (pre_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                       "algorithmic_bytes_per_frame": alg_bytes,
                       "format_bytes_per_frame": fmt_bytes,
                       "frac_of_format_bytes": fmt_bytes * frames_per_pass / (pre_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                       "traffic": tr_pre},
    }
    return roofline, passes * (1 + int(enc.lib.pvr_encoder_launch_count(enc.handle)))


def finish(args, world, rank, dist, name, n_frames, obs_per_step, frames_per_step, width, n_rot, bytes_per_batch,
           value, ms, e2e_value, ms_e2e, clocks, roofline, n_launch, bc, extras, passes):
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        # bounded sample of about 10-15 s of CPU work at the rates these oracle ports reach on 16 host cores
        n_obs = {"moco_aug_uber_34": 80, "moco_aug": 256, "clip_vit_b16": 128, "mae_base": 128, "mae_large": 96,
                 "clip_vit": 384}.get(name, 96)
        v, dt = cpu_port_frames_per_s(name, n_frames, n_obs, threads)
        cpu_baseline = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                        "sample": f"{n_obs} observations x {n_frames} frames of the same workload, mini-batch 64, "
                                  f"torch CPU fp32 oracle port, {dt:.1f} s"}
        if bc is not None:
            v_bc, dt_bc = cpu_port_bc_steps_per_s(threads)
            bc["cpu_baseline"] = {"value": v_bc, "unit": "steps/s", "cores": threads, "kind": "port",
                                  "sample": f"{BC_CPU_STEPS} steps at the same shape (T=64, B=128, D=2048), torch CPU fp32 "
                                            f"oracle port, {dt_bc:.1f} s"}

    line = {
        "metric": "pvr_frames_per_sec_embedded", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": (f"{name}: ResNet-50 MoCo layer3+layer4 compressed taps (2 trunks), {n_frames}-frame "
                         "224x224 uint8 observations, random-init weights (BASELINE configs[1])")
            if args.workload == "uber34x3" else
            (f"{name}: CLIP-architecture ViT-B/{CLIP_PATCH[name]}, 224x224 uint8 frames, random-init weights"
             if name in CLIP_PATCH else
             (f"{name}: MAE ViT-{dict(mae_base='B/16', mae_large='L/16', mae_huge='H/14')[name]} encoder, bicubic "
              "Resize(256)+CenterCrop(224) of 224x224 uint8 frames, random-init weights" if name in MAE_NAMES else
              f"{name}: 224x224 uint8 frames, random-init weights")),
            "obs_per_step_per_gpu": obs_per_step, "frames_per_step_per_gpu": frames_per_step,
            "passes_per_step": passes, "obs_per_pass": obs_per_step // passes,
            "embedding_width": width, "sharding": "observations split over ranks, no data-path collective",
            "l2": f"inputs rotate over {n_rot} distinct batches ({n_rot * bytes_per_batch / 2**20:.0f} MiB > 126 MB L2); "
                  "activations (>1 GB/step) are rewritten every step",
        },
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": bytes_per_batch,
                "d2h_bytes_per_step": obs_per_step * width * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": args.steps * n_launch,
        "tflops_effective": value * GFLOP_PER_FRAME[name] / 1e3,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "bc": bc,
    }
    line.update(extras)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
