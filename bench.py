#!/usr/bin/env python
"""Headline benchmark of the flow_diffuser hot path (BASELINE.json configs[1]):
DDIM-50 flow sampling, Sintel-shaped synthetic 436x1024 frames, batch 8 per GPU.

    python bench.py --gpus N --steps K --warmup W            # this implementation (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU cores

One "step" = one DDIM-50 sampling pass over one batch of 8 frame pairs = 8 flows.
Prints ONE JSON line (rank 0).  Multi-GPU: one process per GPU under torchrun, the batch is sharded
by rank with no data-path collective (sampling has no exchange step) -> weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, BATCH, DDIM_STEPS, TIMESTEPS = 436, 1024, 8, 50, 1000
CONV_GF_PER_SAMPLE = 1590.3      # algorithmic conv GFLOP per UNet forward at 440x1024
FWD_GF_PER_SAMPLE = 1635.28      # whole forward (conv + attention matmuls)
CONV_DRAM_BYTES_PER_FORWARD = 38.411e9   # measured DRAM traffic of the conv launches of one batch-8 forward (profiles/r1_forward_traffic_v5.txt)
CONFIG = {"workload": "flow_diffuser DDIM-50 sampling, Sintel-shaped synthetic 436x1024 (UNet runs on 440x1024), "
                      "target=flow, batch 8 per GPU, random-init weights seed 0",
          "batch_per_gpu": BATCH, "ddim_steps": DDIM_STEPS, "timesteps": TIMESTEPS, "parallelism": "batch-sharded replicas",
          "l2": "working set >> L2: every bf16 activation tensor is 461 MB at full resolution, no reuse between steps"}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "src": "MEASURED_PEAKS.json"}
    except Exception:  # noqa: BLE001
        return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle's fp32 restatement of the reference algorithm
# (the reference tree itself cannot travel to the GPU box) on all host threads.
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_seconds(steps: int, warmup: int):
    """Each step = ONE of the 50 DDIM steps of ONE 436x1024 sample (UNet forward + update); flows/s = 1/(50 t)."""
    from oracle import flowdiff_oracle as O
    from opticalflowdiffusion_b200.unet_params import UnetParams
    # all host cores this process may use (torchrun pins OMP_NUM_THREADS=1 by default)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(0)
    sd = UnetParams(64, channels=5, out_dim=2).state_dict()
    sched = O.make_schedule(TIMESTEPS)
    cond = O.replicate_pad_to_multiple(O.synthetic_frames(1, H, W, seed=0) * 2 - 1)[0]
    x = torch.randn(1, 2, cond.shape[-2], cond.shape[-1], generator=torch.Generator().manual_seed(1234))
    times = O.ddim_times(TIMESTEPS, DDIM_STEPS)
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            tt = torch.full((1,), times[0], dtype=torch.long)
            out = O.unet_forward(sd, x, cond, tt)
            O.ddim_update(sched, x, out, times[0], times[1])
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args, rank: int):
    if rank != 0:
        return
    ts = cpu_reference_step_seconds(args.steps, min(args.warmup, 1))
    threads = torch.get_num_threads()
    t = sum(ts) / len(ts)
    value = 1.0 / (DDIM_STEPS * t)
    sample = "one of the 50 DDIM steps of ONE 436x1024 sample per step (UNet forward + update), x50 extrapolated"
    line = {"impl": "reference", "metric": "flows/sec DDIM-50 @436x1024", "value": value, "unit": "flows/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3 * DDIM_STEPS * BATCH,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": CONFIG,
            "cpu_baseline": {"value": value, "unit": "flows/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "flows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist
    from opticalflowdiffusion_b200 import FlowDiffuser, _lib
    from opticalflowdiffusion_b200.config import compose
    from opticalflowdiffusion_b200.datasets import synthetic_frames
    from opticalflowdiffusion_b200.parallel import max_over_ranks

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load(check_device=True)
    if world > 1:
        # NCCL prints its version banner to stdout when NCCL_DEBUG is set on the box; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    cfg = compose(["algorithm.target=flow", f"algorithm.sampling_timesteps={DDIM_STEPS}",
                   f"algorithm.image_size=[{H},{W}]", "algorithm.return_all_timesteps=true",
                   f"algorithm.use_cuda_graph={'true' if args.graph else 'false'}"])
    torch.manual_seed(0)
    if args.skip_sample:
        tr = run_train_leg(args, cfg.algorithm, rank, world, dev)
        if rank == 0:
            _emit({"train": tr, "n_gpus": world})
        return
    algo = FlowDiffuser(cfg.algorithm).to(dev)
    algo.unet.prepare()
    frames = synthetic_frames(BATCH, H, W, seed=100 + rank)
    cond_host = (2 * frames - 1).pin_memory()
    flow_host = torch.zeros(BATCH, 2, H, W).pin_memory()
    out_host = torch.empty(BATCH, 2, H, W).pin_memory()
    cond_dev = cond_host.to(dev)
    flow_dev = flow_host.to(dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_T = torch.randn(BATCH, 2, H, W, device=dev, generator=gen)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        return algo.sample(cond_dev, flow_dev, x_T=x_T)

    def step_e2e():
        c = cond_host.to(dev, non_blocking=True)
        f = flow_host.to(dev, non_blocking=True)
        _, flows = algo.sample(c, f, x_T=x_T)
        out_host.copy_(flows[:, -1], non_blocking=True)
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        sync()
        return max_over_ranks(s.elapsed_time(e) * 1e-3, device=dev)

    for _ in range(max(args.warmup, 3) if args.warmup >= 0 else 3):
        step_resident()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = lib.fd_launch_count() + algo.model.graph_replayed_launches
    t_res = timed(step_resident, args.steps)
    launches = int(lib.fd_launch_count() + algo.model.graph_replayed_launches - l0)
    t_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop() if rank == 0 else None

    # roofline pass (not timed above): CUDA events around every tensor-core conv launch of one UNet forward
    # (averaged over RF_REPS forwards: a single forward under the power cap varies by +-3 %)
    RF_REPS = 3
    t = torch.full((BATCH,), 999, device=dev, dtype=torch.long)
    torch.set_grad_enabled(False)          # the inference forward (with autograd on, unet(...) is the training forward)
    algo.unet(x_T, cond_dev, t)            # warm: the eager path's allocations are cached before anything is timed
    algo.unet._conv_timing = []
    for _ in range(RF_REPS):
        algo.unet(x_T, cond_dev, t)
    torch.cuda.synchronize()
    conv_s = sum(ev[0].elapsed_time(ev[1]) for _, _, ev in algo.unet._conv_timing) * 1e-3 / RF_REPS
    conv_launches = len(algo.unet._conv_timing) // RF_REPS
    algo.unet._conv_timing = None
    algo.unet(x_T, cond_dev, t)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(RF_REPS):
        algo.unet(x_T, cond_dev, t)
    e.record()
    torch.cuda.synchronize()
    fwd_s = s.elapsed_time(e) * 1e-3 / RF_REPS
    torch.set_grad_enabled(True)

    train = None
    if not args.skip_train:
        del algo
        torch.cuda.empty_cache()
        train = run_train_leg(args, cfg.algorithm, rank, world, dev)
    if rank != 0:
        return
    peaks = load_peaks()
    flows = world * BATCH * args.steps
    value = flows / t_res
    conv_tf = CONV_GF_PER_SAMPLE * BATCH * 1e9 / conv_s / 1e12
    line = {
        "metric": "flows/sec DDIM-50 @436x1024", "value": value, "unit": "flows/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": CONFIG,
        "e2e": {"value": flows / t_e2e, "unit": "flows/s", "h2d_bytes_per_step": cond_host.numel() * 4 + flow_host.numel() * 4,
                "d2h_bytes_per_step": out_host.numel() * 4},
        "gpu_launches": launches, "cuda_graph": bool(args.graph),
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel (tcgen05 implicit GEMM, all %d conv launches of one forward)" % conv_launches,
                     "achieved": conv_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": conv_tf / peaks["tf_sustained"],
                     "traffic": CONV_DRAM_BYTES_PER_FORWARD, "traffic_source": "profiles/r1_forward_traffic_v5.txt: ncu dram__bytes_read.sum + "
                     "dram__bytes_write.sum summed over the conv launches of ONE forward at this batch (same unit as `achieved`: one "
                     "forward's conv launches); algorithmic FLOPs, not bytes, bound these kernels", "peak_source": peaks["src"] + " (sustained: the kernel is timed inside a long step)",
                     "note": "5 of these launches (the full-resolution block2.proj strip convs) also apply the preceding GroupNorm + SiLU "
                             "to their input strips (fd_conv3x3_gnsilu_in) unless FD_FUSE_GN=0; their time is counted, the activation FLOPs are not",
                     "conv_share_of_forward": conv_s / fwd_s, "forward_ms": fwd_s * 1e3,
                     "whole_step_tflops": FWD_GF_PER_SAMPLE * DDIM_STEPS * 1e9 * value / 1e12},
    }
    if train is not None:
        line["train"] = train
    if world == 1 and not args.skip_warp:
        # BASELINE configs[3]: fused backward-warp + photometric / EPE microbench, 8x2x436x1024 flow over 3-channel frames,
        # L2 flushed between iterations; GB/s over the ALGORITHMIC bytes (fwd 40 B/px, bwd 60 B/px)
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_warp
        wm = bench_warp.measure(iters=20, warmup=5)
        line["warp_microbench"] = {k: dict(v, frac_of_hbm_peak=round(v["GBps"] / peaks["hbm"], 3)) for k, v in wm.items()
                                   if k in ("photo_epe_fwd", "photo_epe_bwd", "backwarp_fwd", "backwarp_bwd", "splat_fwd",
                                            "splat_flowgrad")}
        line["warp_microbench"]["hbm_peak_GBps"] = peaks["hbm"]
    if world == 1 and not args.no_cpu_baseline:
        ts = cpu_reference_step_seconds(1, 1)
        tcpu = sum(ts) / len(ts)
        line["cpu_baseline"] = {"value": 1.0 / (DDIM_STEPS * tcpu), "unit": "flows/s", "cores": torch.get_num_threads(),
                                "kind": "port",
                                "sample": "one of the 50 DDIM steps of ONE 436x1024 sample (UNet forward + update) on the host "
                                          "cores after one warm-up, x50 extrapolated"}
    _emit(line)
    if world > 1:
        pass


TRAIN_H, TRAIN_W = 368, 768
TRAIN_GF_PER_SAMPLE = 3.0 * 1019.83      # forward 1019.83 GF at 368x768, fwd + bwd ~ 3x


def cpu_train_baseline():
    """The reference algorithm's training step (p_losses + backward through the oracle's fp32 restatement, no optimiser)
    on the host cores at a 64x64 crop, batch 1, extrapolated to 368x768 by pixel count."""
    from oracle import flowdiff_oracle as O
    from opticalflowdiffusion_b200.unet_params import UnetParams
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        pass
    torch.manual_seed(0)
    sd = {k: v.clone().requires_grad_(True) for k, v in UnetParams(64, channels=5, out_dim=2).state_dict().items()}
    sched = O.make_schedule(TIMESTEPS)
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(1, 2, 64, 64, generator=g) * 2 - 1
    cond = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    noise = torch.randn(1, 2, 64, 64, generator=g)
    t = torch.tensor([500])
    ts = []
    for i in range(3):
        t0 = time.perf_counter()
        loss = O.p_losses_flow(sd, sched, x0, cond, t, noise)
        loss.backward()
        ts.append(time.perf_counter() - t0)
    sec = min(ts[1:])
    scale = (TRAIN_H * TRAIN_W) / (64.0 * 64.0)
    return {"value": 1.0 / (sec * scale), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "forward + backward of ONE 64x64 crop (p_losses, target=flow) after a warm-up, extrapolated x%.0f by "
                      "pixel count to 368x768" % scale}


def run_train_leg(args, algo_cfg, rank: int, world: int, dev):
    """BASELINE configs[2]: one training step (preprocess -> q_sample -> UNet fwd -> loss -> UNet bwd -> gradient
    all-reduce -> fused clip + Adam) at 368x768 crops, batch-sharded data parallel.  Returns the "train" object."""
    import torch.distributed as dist
    from opticalflowdiffusion_b200 import FlowDiffuser, _lib
    from opticalflowdiffusion_b200.datasets import synthetic_frames
    from opticalflowdiffusion_b200.optim import allreduce_gradients
    from opticalflowdiffusion_b200.parallel import max_over_ranks
    lib = _lib.load()
    B = args.train_batch
    torch.manual_seed(0)
    algo = FlowDiffuser(algo_cfg).to(dev)
    algo.train()
    opt = algo.configure_optimizers()
    opt.max_grad_norm = 100.0                                   # experiment=matrix_flow: training.clipping = 100
    g = torch.Generator().manual_seed(7 + rank)
    img_h = synthetic_frames(B, TRAIN_H, TRAIN_W, seed=200 + rank).pin_memory()
    tgt_h = synthetic_frames(B, TRAIN_H, TRAIN_W, seed=300 + rank).pin_memory()
    flow_h = (torch.randn(B, 2, TRAIN_H, TRAIN_W, generator=g) * 5.0).pin_memory()
    batch_dev = tuple(t.to(dev) for t in (img_h, tgt_h, flow_h))
    loss_host = torch.empty((), pin_memory=True)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step(batch):
        first, cond, _ = algo.preprocess(batch, aug=True)          # fused GPU augmentation (augment.py)
        loss = algo.loss(first, cond, None)
        loss.backward()
        allreduce_gradients(opt)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def step_resident():
        return one_step(batch_dev)

    # e2e: every step uploads its own inputs from pinned host memory and reads the loss back; like a DataLoader with
    # pin_memory + prefetch, the upload of step i+1 runs on a copy stream while step i computes (all inside the timed region)
    copy_stream = torch.cuda.Stream(device=dev)
    pending = []

    def upload():
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            b = tuple(t.to(dev, non_blocking=True) for t in (img_h, tgt_h, flow_h))
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return b, ev

    def step_e2e():
        if not pending:
            pending.append(upload())
        batch, ev = pending.pop()
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for t in batch:
            t.record_stream(cur)
        pending.append(upload())
        loss = one_step(batch)
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        sync()
        return max_over_ranks(s.elapsed_time(e) * 1e-3, device=dev)

    for _ in range(3):
        first_loss = step_resident()
    steps = max(args.steps, 3)
    l0 = lib.fd_launch_count()
    t_res = timed(step_resident, steps)
    launches = int(lib.fd_launch_count() - l0)
    t_e2e = timed(step_e2e, steps)
    # phase split of one step (events on the launching stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev_aug = torch.cuda.Event(enable_timing=True)
    ev_aug.record()
    first, cond, _ = algo.preprocess(batch_dev, aug=True)
    ev[0].record()
    loss = algo.loss(first, cond, None)
    ev[1].record()
    loss.backward()
    ev[2].record()
    allreduce_gradients(opt)
    opt.step()
    opt.zero_grad(set_to_none=True)
    ev[3].record()
    torch.cuda.synchronize()
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_train_baseline()
    samples = world * B * steps
    value = samples / t_res
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    return {"metric": "train samples/sec @368x768", "value": value, "unit": "samples/s", "ms_per_step": t_res / steps * 1e3,
            "batch_per_gpu": B, "global_batch": B * world, "steps": steps, "warmup": 3, "dtype": "bf16 activations / fp32 master weights, grads, Adam",
            "e2e": {"value": samples / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": 4 * (img_h.numel() + tgt_h.numel() + flow_h.numel()),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "loss_first": float(first_loss.detach()), "loss_last": float(loss.detach()),
            "phases_ms": {"augment+preprocess": ev_aug.elapsed_time(ev[0]), "forward+loss": ev[0].elapsed_time(ev[1]), "backward": ev[1].elapsed_time(ev[2]),
                          "allreduce+clip+adam": ev[2].elapsed_time(ev[3])},
            "tensor_tflops": TRAIN_GF_PER_SAMPLE * value / 1e3 / world, "peak_mem_gib": peak_mem, "cpu_baseline": cpu,
            "config": "flow_diffuser training step, target=flow, synthetic 368x768 crops, augmentation ON (GpuAugmentor: "
                      "reference Augmentor semantics, decisions on the host, arithmetic in 4 launches), Adam lr 1e-5 wd 1e-6, clip 100, gradient all-reduce over NCCL when n_gpus > 1"}


_REAL_STDOUT = None


def _emit(obj):
    """The one JSON line of this run, on the process's original stdout."""
    text = json.dumps(obj) + "\n"
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="replay the DDIM loop as one CUDA graph (1) or launch eagerly (0)")
    ap.add_argument("--train-batch", type=int, default=8, help="per-GPU batch of the training leg (368x768 crops)")
    ap.add_argument("--skip-warp", action="store_true", help="omit the warp + photometric/EPE microbench (BASELINE configs[3])")
    ap.add_argument("--skip-train", action="store_true", help="omit the training leg (BASELINE configs[2])")
    ap.add_argument("--skip-sample", action="store_true", help="training leg only (debug; prints a reduced line)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that chat on file descriptor 1 (NCCL prints its version banner there)
    # are sent to stderr for the duration of the run; _emit() writes the result to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
