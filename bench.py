#!/usr/bin/env python
"""bench.py - SR frames/s of the RefineNet x4 hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

A step = one RefineNet x4 inference pass (only the consumed output list, i.e. what the reference predictor
reads through `[-1]`) over B synthetic ACDCSR-shaped cine sequences per GPU (LR 54x63, T=30 phases, U=6 warm-up
frames each side -> 42 LR frames in, 30 SR frames 216x252 out, random-init weights, seed 0).
N > 1: one process per GPU (torchrun), sequences sharded by rank, no data-path collective (weak scaling).

Printed JSON (one line, rank 0): value = device-timed frames/s with inputs resident in HBM; e2e = same metric
through the public module call with pinned host inputs (H2D) and the SR frames read back (D2H) every step;
roofline = ConvLSTM-cell tcgen05 kernel (dominant launch class) FLOP/s vs the measured bf16 peak;
cpu_baseline = the pinned CPU oracle (torch fp32 restatement of the reference) on this box's host cores.
`--impl reference` times that CPU implementation alone (the reference itself cannot be installed: see DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

T_FRAMES, U_FRAMES, LR_H, LR_W, SCALE = 30, 6, 54, 63, 4
# BASELINE.json configs: [1] ACDC x4 (the headline workload), [2] the x2 / x3 scale variants, [3] DSB15SR-shaped x4
WORKLOADS = {"acdc_x4": (54, 63, 4, "ACDCSR"), "acdc_x3": (72, 84, 3, "ACDCSR"), "acdc_x2": (108, 126, 2, "ACDCSR"),
             "dsb15_x4": (63, 48, 4, "DSB15SR"),
             # EDSRNet x4 of configs/{train,test}/edsr_net/exp1_x4.yaml on the same conv core
             "edsr_x4": (54, 63, 4, "ACDCSR"),
             # ... and DRFNet x4 (src/model/nets/drf_net.py; 64 features, 6 projection groups)
             "drfnet_x4": (54, 63, 4, "ACDCSR")}
EDSR_FRAMES = 60
DRF_SEQS, DRF_FRAMES = 16, 30
NET_KW = dict(in_channels=1, out_channels=1, num_features=[64, 64, 64], upscale_factor=SCALE, num_stages=3,
              update_memory=True, num_updated_frames=U_FRAMES, refine_window_size=5, positional_encoding=True)
METRIC = "SR frames/s at x4"
UNIT = "frames/s"
WORKLOADS_NAME = "ACDCSR"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback (B200_PROFILING.md)"
    return d


def ncu_traffic(kernel_class):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json), or
    None when the workload is not the captured one."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f).get(kernel_class)
        return d["dram_read_bytes"] + d["dram_write_bytes"], d["source"]
    except Exception:
        return None, None


def synthetic_sequences(batch, seed):
    """ACDCSR-shaped synthetic cine sequences: circular padding and positional code as the reference dataset
    builds them (acdc_vsr_refinenet_dataset.py:74-87, gen_positional_encoding.py:35-38)."""
    from pvsr.synthetic import cine_batch
    return cine_batch(batch, T=T_FRAMES, U=U_FRAMES, h=LR_H, w=LR_W, scale=SCALE, seed=seed)


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.index = None, index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = sorted(sm)[len(sm) // 2:]  # upper half = samples under load
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_model(state_dict=None, scale=None):
    """The CPU implementation of the path: (forward(inputs, pos) -> 3*S lists, kind).  kind = "reference": the
    UNMODIFIED reference RefineNet staged under oracle/_ref by oracle/make_ref.py (all 3*S heads, exactly what the
    reference predictor executes); kind = "port": the pinned restatement oracle/refinenet_oracle.py when the staged
    copy is absent.  `state_dict`: weights to load (default: seed-0 initialisation of the x`SCALE` config)."""
    from oracle import ref_model, refinenet_oracle as O
    scale = scale or SCALE
    if ref_model.available():
        net = ref_model.build_net(state_dict, seed=0, upscale_factor=scale)
        return (lambda inputs, pos: net(inputs, pos)), "reference"
    sd = state_dict if state_dict is not None else O.init_state_dict(upscale_factor=scale, positional_encoding=True, seed=0)
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    kw = dict(num_stages=3, num_updated_frames=U_FRAMES, refine_window_size=5, upscale_factor=scale,
              positional_encoding=True, memory=True, num_layers=3)
    return (lambda inputs, pos: O.refinenet_forward(sd, inputs, pos, **kw)), "port"


def cpu_oracle_time(n_seq_steps, warmup, threads):
    """Times the CPU implementation (all 3*S heads, as the reference predictor executes them) on one sequence per
    step.  Returns (times, kind)."""
    torch.set_num_threads(threads)
    fwd, kind = cpu_model()
    inputs, pos = synthetic_sequences(1, 1234)
    times = []
    with torch.no_grad():
        for i in range(warmup + n_seq_steps):
            t0 = time.perf_counter()
            fwd(inputs, pos)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, kind


# SURVEY.md section 8c gates (bf16 operands, fp32 accumulate / state) - the same numbers tests/test_model_gpu.py uses
GATES = {"max_abs": 2e-2, "rel_l2": 1.5e-2, "psnr_delta_db": 0.01, "ssim_delta": 1e-4}


def gates_for_scale(scale):
    """SSIM gate = 2x the drift of the REFERENCE ITSELF under torch.autocast(bf16) against its own fp32 run on the same
    sequence and synthetic target (tests/test_model_gpu.py: x4 54x63 4.9e-5 -> 1e-4; x2 108x126 1.02e-4 -> 2e-4: noise
    targets 4x the LR grid, where single uint8 flips move the 11x11-window SSIM more)."""
    return dict(GATES, ssim_delta=2e-4 if scale == 2 else GATES["ssim_delta"])


def inference_parity(frames_host, inputs_h, pos_h, hr_h, seqs, state_dict, threads):
    """Checks SR frames of the LAST TIMED e2e step (host copy read back through the HostFrameRing: B sequences x CUDA
    graph replay x rotating output buffers) against the CPU implementation run on the same sequences' own inputs.
    frames_host: [T, B, 1, H, W].  Returns (parity dict, per-run CPU seconds, kind)."""
    from oracle import refinenet_oracle as O
    torch.set_num_threads(threads)
    fwd, kind = cpu_model(state_dict)
    dataset = "dsb15" if WORKLOADS_NAME == "DSB15SR" else "acdc"
    worst = {"max_abs": 0.0, "rel_l2": 0.0, "psnr_delta_db": 0.0, "ssim_delta": 0.0}
    times = []
    with torch.no_grad():
        for i in seqs:
            t0 = time.perf_counter()
            ref = fwd([x[i:i + 1] for x in inputs_h], pos_h[i:i + 1])[-1]
            times.append(time.perf_counter() - t0)
            for t, r in enumerate(ref):
                g = frames_host[t, i:i + 1].float()
                worst["max_abs"] = max(worst["max_abs"], (g - r).abs().max().item())
                worst["rel_l2"] = max(worst["rel_l2"], ((g - r).norm() / r.norm()).item())
                tgt = O.denormalize(hr_h[t][i:i + 1], dataset)
                dg, dr = O.denormalize(g, dataset), O.denormalize(r, dataset)
                worst["psnr_delta_db"] = max(worst["psnr_delta_db"], abs(float(O.psnr(dg, tgt)) - float(O.psnr(dr, tgt))))
                worst["ssim_delta"] = max(worst["ssim_delta"], abs(float(O.ssim(dg, tgt)) - float(O.ssim(dr, tgt))))
    gates = gates_for_scale(SCALE)
    ok = all(worst[k] <= gates[k] for k in gates)
    par = dict(worst, ok=ok, gates=gates, checked_against=kind,
               what=f"SR frames of sequences {list(seqs)} of the last timed e2e step (graph replay, output ring, D2H) "
                    f"vs the CPU {kind} on the same inputs; PSNR / SSIM deltas against a synthetic HR target")
    return par, times, kind


def edsr_cpu_time(frames):
    """CPU oracle of EDSR x4 (32 x 256) on `frames` ACDCSR-shaped LR frames: (seconds, cores)."""
    from oracle import edsr_oracle as O
    from src.model.nets import EDSRNet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = EDSRNet(in_channels=1, out_channels=1, num_resblocks=32, num_features=256, upscale_factor=4,
                 res_scale=0.1).state_dict()
    x = torch.randn(frames, 1, LR_H, LR_W, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        O.edsr_forward(sd, x[:1], 32, 4, 0.1)
        t0 = time.perf_counter()
        O.edsr_forward(sd, x, 32, 4, 0.1)
        return time.perf_counter() - t0, cores


def run_edsr(args, rank, world):
    """`--workload edsr_x4`: single-GPU line for the EDSR widening row (profiles/bench_edsr.py does the device timing)."""
    if rank != 0:
        return
    metric = "SR frames/s at x4 (EDSRNet 32 x 256)"
    cfg = {"workload": f"EDSRNet x4 inference (32 residual blocks x 256 features, 43 M parameters, random init), "
                       f"{EDSR_FRAMES} synthetic ACDCSR-shaped LR frames {LR_H}x{LR_W} per step -> SR frames 216x252",
           "name": "edsr_x4", "frames_per_step": EDSR_FRAMES, "parallelism": "single GPU",
           "l2": "256 MiB flush write between steps"}
    if args.impl == "reference":
        times = []
        for _ in range(max(1, args.steps)):
            dt, cores = edsr_cpu_time(2)
            times.append(dt)
        value = 2 * len(times) / sum(times)
        sample = f"{len(times)} step(s) x 2 LR frames {LR_H}x{LR_W} through the fp32 torch CPU oracle (oracle/edsr_oracle.py)"
        print(json.dumps({"impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": len(times), "warmup": 1, "ms_per_step": 1e3 * sum(times) / len(times),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": dict(cfg, frames_per_step=2),
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import bench_edsr
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    res = bench_edsr.measure(argparse.Namespace(frames=EDSR_FRAMES, steps=args.steps, warmup=args.warmup))
    clocks = sampler.stop()
    inf = res["inference"]
    line = {"metric": metric, "value": inf["frames_per_s"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": inf["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": inf["e2e_frames_per_s"], "unit": UNIT, "h2d_bytes_per_step": inf["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": inf["d2h_bytes_per_step"], "ms_per_step": inf["ms_per_step_e2e"]},
            "gpu_launches": inf["launches_per_step"] * args.steps, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv3x3_halo_kernel<256, EPI_STORE / EPI_PS> (all 71 launches of a step)",
                         "achieved": inf["tflops"], "peak": res["peak_tflops"], "unit": "TFLOP/s",
                         "frac": inf["frac_of_peak"], "traffic": None, "peak_source": res["peak_source"]},
            "train_step": res["train_step"]}
    if not args.no_cpu_baseline:
        dt, cores = edsr_cpu_time(2)
        line["cpu_baseline"] = {"value": 2 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "2 LR frames 54x63 through the fp32 torch CPU oracle (oracle/edsr_oracle.py)"}
    print(json.dumps(line), flush=True)


def drf_cpu_time(frames):
    """CPU oracle of DRFNet x4 (64 features, 6 groups) on one ACDCSR-shaped sequence of `frames` LR frames: (s, cores)."""
    from oracle import drf_oracle as O
    from src.model.nets import DRFNet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = DRFNet(in_channels=1, out_channels=1, num_features=64, num_groups=6, upscale_factor=4).state_dict()
    g = torch.Generator().manual_seed(1234)
    xs = [torch.randn(1, 1, LR_H, LR_W, generator=g) for _ in range(frames)]
    with torch.no_grad():
        O.drf_forward(sd, xs[:1], 6, 4)
        t0 = time.perf_counter()
        O.drf_forward(sd, xs, 6, 4)
        return time.perf_counter() - t0, cores


def run_drf(args, rank, world):
    """`--workload drfnet_x4`: single-GPU line for the DRFNet widening row (profiles/bench_drf.py does the device timing)."""
    if rank != 0:
        return
    metric = "SR frames/s at x4 (DRFNet 64 x 6)"
    cfg = {"workload": f"DRFNet x4 inference (64 features, 6 projection groups, 3.66 M parameters, random init), "
                       f"{DRF_SEQS} synthetic ACDCSR-shaped cine sequences of {DRF_FRAMES} LR frames {LR_H}x{LR_W} per "
                       "step -> SR frames 216x252; the frame recurrence is sequential, the sequences are batched",
           "name": "drfnet_x4", "frames_per_step": DRF_SEQS * DRF_FRAMES, "parallelism": "single GPU",
           "l2": "256 MiB flush write between steps"}
    cpu_frames = 6
    sample = (f"one sequence of {cpu_frames} LR frames {LR_H}x{LR_W} through the fp32 torch CPU oracle "
              "(oracle/drf_oracle.py, pinned to the unmodified reference by tests/golden/drfnet_*.npz)")
    if args.impl == "reference":
        times = []
        for _ in range(max(1, args.steps)):
            dt, cores = drf_cpu_time(cpu_frames)
            times.append(dt)
        value = cpu_frames * len(times) / sum(times)
        print(json.dumps({"impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": len(times), "warmup": 1, "ms_per_step": 1e3 * sum(times) / len(times),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": dict(cfg, frames_per_step=cpu_frames),
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import bench_drf
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    res = bench_drf.measure(argparse.Namespace(seqs=DRF_SEQS, frames=DRF_FRAMES, steps=args.steps, warmup=args.warmup))
    clocks = sampler.stop()
    inf = res["inference"]
    line = {"metric": metric, "value": inf["frames_per_s"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": inf["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": inf["e2e_frames_per_s"], "unit": UNIT, "h2d_bytes_per_step": inf["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": inf["d2h_bytes_per_step"], "ms_per_step": inf["ms_per_step_e2e"]},
            "gpu_launches": inf["launches_per_step"] * args.steps, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv3x3_halo_kernel / conv3x3_kernel <256 | 64, EPI_STORE> (every "
                                                      "launch of a step: 3x3 phase-stacked projection units + 1x1 convs)",
                         "achieved": inf["tflops"], "peak": res["peak_tflops"], "unit": "TFLOP/s",
                         "frac": inf["frac_of_peak"], "traffic": None, "peak_source": res["peak_source"],
                         "executed_tflops": inf["executed_tflops"], "executed_frac": inf["executed_frac_of_peak"],
                         "note": "achieved = ALGORITHMIC FLOPs of the reference's k x k ConvTranspose2d / strided Conv2d; "
                                 "the phase-stacked 3x3 forms execute 9 s^2 / k^2 = 2.25x as many"},
            "train_step": res["train_step"]}
    if not args.no_cpu_baseline:
        dt, cores = drf_cpu_time(cpu_frames)
        line["cpu_baseline"] = {"value": cpu_frames / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    times, kind = cpu_oracle_time(args.steps, args.warmup, cores)
    total = sum(times)
    value = T_FRAMES * len(times) / total
    sample = (f"{len(times)} step(s) x 1 {args.workload} sequence (42 LR frames {LR_H}x{LR_W} -> 30 SR frames), "
              "all 9 heads, fp32, " + ("the unmodified reference RefineNet (oracle/_ref)" if kind == "reference"
                                       else "oracle port (oracle/_ref not staged)"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"RefineNet x{SCALE} inference, synthetic cine sequences (LR {LR_H}x{LR_W}, T=30, U=6), "
                                   "1 sequence per step on the host CPU", "name": args.workload,
                       "sequences_per_step": 1},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


TRAIN_N, TRAIN_T, TRAIN_HW = 16, 7, 32     # configs/train/refine_net/exp1_x4.yaml:20-33: batch 16, 7 frames, 32x32 patches


def nccl_tuning_lines():
    """What this process's NCCL log says about the gradient all-reduce when bench.py itself turned the log on
    (NCCL_DEBUG unset by the caller -> INFO + INIT,TUNING into a private file): the tuner's algorithm / protocol lines
    when NCCL prints them, else the transport summary of the communicator (NVLS / P2P channels, init line); [] otherwise."""
    import re
    path = os.environ.get("PVSR_NCCL_LOG")
    if not path:
        return []
    path = path.replace("%p", str(os.getpid()))
    try:
        with open(path) as f:
            raw = [l.strip() for l in f]
    except OSError:
        return []
    def pick(pattern, limit):
        seen, out = set(), []
        for l in raw:
            if re.search(pattern, l):
                key = re.sub(r"^.*NCCL INFO ", "", l)
                key = re.sub(r"\b\d+ *-> *\d+\b|\[\d+\]", "", key)
                if key not in seen:
                    seen.add(key)
                    out.append(re.sub(r"^.*NCCL INFO ", "", l))
        return out[-limit:]
    tuned = pick(r"(?i)allreduce.*(algo|proto)|algo.*proto", 4)
    return tuned if tuned else pick(r"(?i)nvls|via P2P|Init COMPLETE|nChannels|Connected all", 5)


def tail_flop_saving(n_pixels_r1, backward):
    """FLOPs the rank-1 forms of the head's tail (csrc/tail_rank1.cu) do NOT execute, for `n_pixels_r1` input pixels of
    the last conv + PixelShuffle(2): algorithmic conv-by-conv FLOPs (2*9*64*256 + 4 * 2*9*64 per pixel forward, twice
    that backward) minus the executed ones (forward: 25 taps x 64 ch x 8 mma columns; backward: two 64 x 64 products per
    pixel with a bf16 hi + lo operand)."""
    algo = 2 * 9 * 64 * 256 + 4 * 2 * 9 * 64
    fwd = (algo - 2 * 25 * 64 * 8) * n_pixels_r1
    bwd = (2 * algo - 2 * (2 * 64 * 64 * 2)) * n_pixels_r1 if backward else 0
    return fwd + bwd


GRAD_GATES = {"grad_rel_l2": 4e-2, "grad_cos": 0.999, "loss_rel": 2e-3, "out_rel_l2": 1.5e-2}


def train_parity(net, eng, inputs_d, pos_d, targets_d, inputs_h, pos_h, threads):
    """Parity of the BENCHMARKED training plan (N = 16 per GPU, its own weight-gradient split counts) at the current
    weights, on rank 0, outside the timed region:
      (1) gradients of the fused N=16 step vs the mean of two N=8 steps through the generic autograd path (their own
          plans; the loss is a mean over samples, so the two must agree) - per-tensor rel-L2 and cosine;
      (2) the fused loss vs the trainer's loss formula (acdc_vsr_refinenet_trainer.py:83-93) evaluated on the step's own
          output frames;
      (3) output frames of samples 0 and N-1 (all 9 lists, train mode) vs the CPU implementation's forward."""
    from oracle import refinenet_oracle as O
    N = inputs_d[0].shape[0]
    flat_p, flat_g = eng.flatten_parameters()
    loss16, out16 = eng.loss_and_grads(inputs_d, pos_d, targets_d)
    g16 = flat_g.clone()
    out16 = out16.clone()
    names = [(k, p) for k, p in net.named_parameters()]
    flat_g.zero_()
    h = N // 2
    for lo in (0, h):
        out = net([x[lo:lo + h] for x in inputs_d], pos_d[lo:lo + h])
        O.trainer_loss(out, [t[lo:lo + h] for t in targets_d], training=True).backward()
    torch.cuda.synchronize()
    worst_rel, worst_cos, worst_name = 0.0, 1.0, None
    off = 0
    for k, p_ in names:
        n = p_.numel()
        a = g16[off:off + n].double()
        b = p_.grad.reshape(-1).double() * 0.5 if p_.grad is not None else torch.zeros_like(a)
        off += (n + 3) // 4 * 4
        if float(b.norm()) == 0.0:
            continue
        rel = float((a - b).norm() / b.norm())
        cos = float((a * b).sum() / (a.norm() * b.norm()))
        if rel > worst_rel:
            worst_rel, worst_name = rel, k
        worst_cos = min(worst_cos, cos)
    lists = tuple([out16[l, t].unsqueeze(1) for t in range(out16.shape[1])] for l in range(out16.shape[0]))
    loss_formula = float(O.trainer_loss(lists, targets_d, training=True))
    loss_rel = abs(float(loss16) - loss_formula) / abs(loss_formula)
    # forward of the training plan vs the CPU implementation (train-mode forward == eval forward: no dropout / BN)
    torch.set_num_threads(threads)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    fwd, kind = cpu_model(sd, scale=4)
    out_rel = 0.0
    with torch.no_grad():
        for i in (0, N - 1):
            ref = fwd([x[i:i + 1] for x in inputs_h], pos_h[i:i + 1])
            ref = torch.stack([torch.stack(list(o)) for o in ref])[:, :, 0, 0]
            got = out16[:, :, i].cpu()
            out_rel = max(out_rel, float((got - ref).norm() / ref.norm()))
    flat_g.zero_()
    res = {"grad_rel_l2": worst_rel, "grad_rel_l2_tensor": worst_name, "grad_cos": worst_cos, "loss_rel": loss_rel,
           "out_rel_l2": out_rel, "gates": GRAD_GATES, "checked_against": f"two N={h} autograd steps (gradients), trainer "
           f"loss formula on the step's own frames (loss), CPU {kind} forward of samples 0 and {N - 1} (frames)"}
    res["ok"] = bool(worst_rel <= GRAD_GATES["grad_rel_l2"] and worst_cos >= GRAD_GATES["grad_cos"] and
                     loss_rel <= GRAD_GATES["loss_rel"] and out_rel <= GRAD_GATES["out_rel_l2"])
    return res


def bench_train(args, dev, rank, world, distributed, barrier, parity=True):
    """One RefineNet x4 training step (forward + multi-stage L1 + backward + gradient all-reduce + Adam) per step at
    the reference's training shapes: N=16 per GPU, 7 target frames + 2x6 warm-up frames of 32x32 LR patches.
    Returns a dict for the JSON line (rank 0) - SURVEY.md section 8 config 5."""
    import torch.distributed as dist
    from pvsr.optim import FusedAdam
    from pvsr.parallel import DataParallelStep
    from pvsr.synthetic import cine_batch
    from src.model.nets import RefineNet
    kw = dict(NET_KW, upscale_factor=4)
    torch.manual_seed(0)
    net = RefineNet(**kw).to(dev).train()
    opt = FusedAdam.for_net(net, lr=1e-4)
    dp = DataParallelStep(net, opt)
    eng = net.engine
    eng.use_graph = not args.no_graph
    inputs_h, pos_h, targets_h = cine_batch(TRAIN_N, T=TRAIN_T, U=U_FRAMES, h=TRAIN_HW, w=TRAIN_HW, scale=4,
                                            seed=4321 + rank, end_systole=3, with_targets=True)
    inputs_h = [x.pin_memory() for x in inputs_h]
    targets_h = [x.pin_memory() for x in targets_h]
    pos_h = pos_h.pin_memory()
    inputs_d, pos_d, targets_d = [x.to(dev) for x in inputs_h], pos_h.to(dev), [x.to(dev) for x in targets_h]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss_h = torch.zeros((), dtype=torch.float32).pin_memory()

    def device_step():
        flush.fill_(1)
        loss, _ = eng.loss_and_grads(inputs_d, pos_d, targets_d)
        dp.step()
        return loss

    def e2e_step():
        flush.fill_(1)
        xs = [x.to(dev, non_blocking=True) for x in inputs_h]
        ts = [x.to(dev, non_blocking=True) for x in targets_h]
        ps = pos_h.to(dev, non_blocking=True)
        loss, _ = eng.loss_and_grads(xs, ps, ts)
        dp.step()
        loss_h.copy_(loss, non_blocking=True)
        return loss

    res, ar_ms = {}, None
    for name, fn in (("device", device_step), ("e2e", e2e_step)):
        for _ in range(max(args.warmup, 3)):
            fn()
        barrier()
        dp.time_allreduce = name == "device"
        dp.allreduce_ms()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = fn()
        e1.record()
        barrier()
        res[name] = e0.elapsed_time(e1)
        if name == "device":
            ar_ms = dp.allreduce_ms()
        dp.time_allreduce = False
    t = torch.tensor([res["device"], res["e2e"], ar_ms or 0.0], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ar_ms = t.tolist()
    loss_val = float(loss)
    pl = eng.train_plan(inputs_d)
    prof_f = eng.profile(pl)
    prof_b = eng.profile_backward(pl)
    flops = pl.flops + pl.flops_bwd
    lib = eng.plans[next(iter(eng.plans))].lib
    tail_on = bool(lib.pvsr_get_tail_rank1())
    tail_saved = tail_flop_saving(9 * TRAIN_T * TRAIN_N * (2 * TRAIN_HW) ** 2, True) if tail_on else 0.0
    if tail_on and not lib.pvsr_get_tail_fwd():
        tail_saved -= tail_flop_saving(9 * TRAIN_T * TRAIN_N * (2 * TRAIN_HW) ** 2, False)
    frames = world * TRAIN_N * TRAIN_T
    h2d = sum(x.numel() * 4 for x in inputs_h + targets_h) + pos_h.numel() * 4
    out = {
        "metric": "training target frames/s at x4", "value": frames * args.steps / (ms / 1e3), "unit": "frames/s",
        "ms_per_step": ms / args.steps, "steps_per_s": args.steps / (ms / 1e3), "n_gpus": world,
        "config": {"workload": f"RefineNet x4 training step, N={TRAIN_N} per GPU, {TRAIN_T} target frames + 2x{U_FRAMES} "
                               f"warm-up frames of {TRAIN_HW}x{TRAIN_HW} LR patches (HR 128x128), all 9 heads, L1 multi-stage "
                               "loss, fused Adam", "global_batch": world * TRAIN_N,
                   "parallelism": f"data-parallel x{world}, one NCCL all-reduce of the flat fp32 gradient per step"},
        "e2e": {"value": frames * args.steps / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "allreduce": {"collective": "ncclAllReduce (SUM, fp32) via torch.distributed" if distributed else "none (1 GPU)",
                      "bytes": dp.allreduce_bytes(), "ms_per_step": ar_ms if distributed else 0.0,
                      "timing": "CUDA events on the compute stream around dist.all_reduce, max over ranks, mean over "
                                "the timed steps (includes waiting for the slowest rank's backward)",
                      "nccl_tuning": nccl_tuning_lines() if distributed else []},
        "gpu_launches": int((pl.launches + pl.launches_bwd + 3) * args.steps),
        "algorithmic_tflop_per_step_per_gpu": flops / 1e12,
        "whole_step_tflops_per_gpu": flops / (ms / args.steps / 1e3) / 1e12,
        # x4 heads: the last conv + shuffle + final conv run in their rank-1 forms (tail_rank1.cu), which execute fewer
        # FLOPs than the reference's conv-by-conv evaluation the ALGORITHMIC figure above counts
        "executed_tflop_per_step_per_gpu": (flops - tail_saved) / 1e12,
        "executed_tflops_per_gpu": (flops - tail_saved) / (ms / args.steps / 1e3) / 1e12,
        "flop_accounting": "whole_step_tflops = ALGORITHMIC conv FLOPs of the reference's layer-by-layer evaluation "
                           "(2*9*Cin*Cout per output pixel, SURVEY.md 8d) / time; executed_* subtracts what the rank-1 "
                           "forward / backward of the head's tail does not execute",
        "loss": loss_val,
        "kernel_ms_per_step": {**{k: round(v[0], 3) for k, v in prof_f.items()},
                               **{k: round(v[0], 3) for k, v in prof_b.items()}},
        "kernel_tflops": {k: round(v[2] / (v[0] / 1e3) / 1e12, 1) for k, v in {**prof_f, **prof_b}.items()
                          if v[2] > 0 and v[0] > 0},
    }
    if parity and rank == 0:
        out["parity"] = train_parity(net, eng, inputs_d, pos_d, targets_d, inputs_h, pos_h, os.cpu_count() or 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="cine sequences per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the B = 1 / 8 operating points of the N=1 line")
    ap.add_argument("--no-train", action="store_true", help="skip the train_step object")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-line parity check of the timed frames")
    ap.add_argument("--workload", default="acdc_x4", choices=sorted(WORKLOADS),
                    help="inference workload shape (default: the BASELINE.json headline config)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train", "both"],
                    help="infer: BASELINE.json metric (default; adds a short train_step object at N=1); "
                         "train: the training-step line; both: inference line with the train_step object")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    global LR_H, LR_W, SCALE, METRIC, WORKLOADS_NAME
    LR_H, LR_W, SCALE, shape_name = WORKLOADS[args.workload]
    WORKLOADS_NAME = shape_name
    NET_KW["upscale_factor"] = SCALE
    METRIC = f"SR frames/s at x{SCALE}"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "edsr_x4":
        run_edsr(args, rank, world)
        return
    if args.workload == "drfnet_x4":
        run_drf(args, rank, world)
        return
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    sys.path.insert(0, os.path.join(PKG, "csrc"))
    import build as pvsr_build
    pvsr_build.build()
    from src.model.nets import RefineNet

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (the path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if "NCCL_DEBUG" not in os.environ and "NCCL_DEBUG_FILE" not in os.environ:
            # algorithm / protocol of the gradient all-reduce for the train_step object (private per-process log file;
            # left alone when the caller configured NCCL logging itself)
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ["NCCL_DEBUG_SUBSYS"] = "INIT,TUNING"
            os.environ["PVSR_NCCL_LOG"] = f"/tmp/pvsr_nccl_{os.getppid()}_%p.log"
            os.environ["NCCL_DEBUG_FILE"] = os.environ["PVSR_NCCL_LOG"]
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    if args.mode == "train":
        sampler = ClockSampler(local_rank) if rank == 0 else None
        tr = bench_train(args, dev, rank, world, distributed, barrier)
        if rank == 0:
            peaks = measured_peaks()
            peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
            line = {"metric": tr["metric"], "value": tr["value"], "unit": tr["unit"], "n_gpus": world,
                    "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": tr["ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                    "data": "synthetic", "config": tr["config"], "e2e": tr["e2e"], "gpu_launches": tr["gpu_launches"],
                    "allreduce": tr["allreduce"], "parity": tr.get("parity"),
                    "clocks": sampler.stop(),
                    "roofline": {"bound": "tensor", "kernel": "whole training step (all tcgen05 conv / dgrad / wgrad launches)",
                                 "achieved": tr["whole_step_tflops_per_gpu"], "peak": peak, "unit": "TFLOP/s",
                                 "frac": tr["whole_step_tflops_per_gpu"] / peak, "traffic": None,
                                 "peak_source": peaks["_source"] + ", sustained bf16"},
                    "kernel_ms_per_step": tr["kernel_ms_per_step"], "kernel_tflops": tr["kernel_tflops"],
                    "loss": tr["loss"]}
            print(json.dumps(line), flush=True)
        if distributed:
            dist.destroy_process_group()
        if rank == 0 and tr.get("parity") and not tr["parity"]["ok"]:
            raise SystemExit("bench.py: training parity check FAILED: " + json.dumps(tr["parity"]))
        return

    torch.manual_seed(0)
    net = RefineNet(**NET_KW).to(dev).eval()
    net.only_last_head = True
    net.reuse_output_buffers = True
    net.engine.use_graph = not args.no_graph
    B = args.batch

    # this rank's shard of the synthetic job: B sequences, host-resident (pinned) for the e2e leg
    from pvsr.synthetic import cine_batch
    inputs_h, pos_h, hr_h = cine_batch(B, T=T_FRAMES, U=U_FRAMES, h=LR_H, w=LR_W, scale=SCALE, seed=1234 + rank,
                                       with_targets=True)     # hr_h: synthetic HR targets (PSNR / SSIM deltas only)
    inputs_h = [x.pin_memory() for x in inputs_h]
    stacked_h = torch.stack(inputs_h).pin_memory()      # the same frames as ONE pinned buffer (a collated batch)
    pos_h = pos_h.pin_memory()
    inputs_d = [x.to(dev) for x in inputs_h]
    pos_d = pos_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    eng = net.engine
    plan = eng.plan_for(B, len(inputs_d), LR_H, LR_W, False, dev)

    def device_step():
        flush.fill_(1)                      # L2 flush: 256 MiB write between steps
        eng.run(plan)                       # inputs already staged in HBM

    from pvsr.hostio import HostFrameRing
    ring = HostFrameRing(dev, slots=2)      # pinned host slots + copy stream: the D2H of step i overlaps step i+1
    eng.output_slots = 2                    # ... which needs a second output buffer for step i+1 to write into

    last_slot = [0]

    def e2e_step():
        flush.fill_(1)
        ring.before_launch(eng.next_output_ptr(plan), eng.next_output_bytes(plan))   # buffer reuse vs copies in flight
        xs = list(stacked_h.to(dev, non_blocking=True).unbind(0))        # H2D from pinned host memory (one copy)
        ps = pos_h.to(dev, non_blocking=True)
        with torch.no_grad():
            frames = net(xs, ps)[-1]                                     # the public module call
        last_slot[0] = ring.submit(frames)                               # D2H read of all SR frames of the step
        return frames

    with torch.no_grad():
        eng.stage_inputs(plan, inputs_d, pos_d)
        for _ in range(args.warmup):
            device_step()
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            device_step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None

        # per-class kernel times (eager pass with events around every launch), averaged over 2 passes
        prof = None
        for _ in range(2):
            pr = eng.profile(plan)
            prof = pr if prof is None else {k: (prof[k][0] + v[0], v[1], v[2]) for k, v in pr.items()}
        prof = {k: (v[0] / 2, v[1], v[2]) for k, v in prof.items()}

        # end-to-end leg through the public API with host buffers
        for _ in range(2):
            e2e_step()
        ring.drain()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            e2e_step()
        ring.drain()                        # the timed region ends when the last step's frames are in host memory
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        # what the LAST TIMED e2e step delivered to the host (checked against the CPU implementation below)
        frames_last = ring.result(last_slot[0]).clone() if rank == 0 else None

        # small-batch operating points (SURVEY.md 8d sweep; the reference predictor runs B = 1): device-timed like `value`
        sweep = {}
        if world == 1 and not args.no_sweep:
            for b in (1, 8):
                if b == B:
                    continue
                pb = eng.plan_for(b, len(inputs_d), LR_H, LR_W, False, dev)
                eng.stage_inputs(pb, [x[:b] for x in inputs_d], pos_d[:b])
                for _ in range(args.warmup + 2):            # eager, eager + capture, first replays (graph upload)
                    flush.fill_(1)
                    eng.run(pb)
                torch.cuda.synchronize()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n_rep = max(args.steps, 10) * (4 if b == 1 else 1)      # 4 ms steps: average over enough of them
                g0.record()
                for _ in range(n_rep):
                    flush.fill_(1)
                    eng.run(pb)
                g1.record()
                torch.cuda.synchronize()
                msb = g0.elapsed_time(g1) / n_rep
                sweep[str(b)] = {"sequences_per_step": b, "ms_per_step": msb, "value": b * T_FRAMES / (msb / 1e3),
                                 "unit": UNIT, "tflops": pb.flops / (msb / 1e3) / 1e12, "launches_per_step": pb.launches}
                del pb
                eng.plans.pop((b, len(inputs_d), LR_H, LR_W, False, str(dev), False), None)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # the training step of config 5 (short, same process) rides along on the line at every N: data-parallel with the
    # NCCL gradient all-reduce (acdc_vsr_refinenet_trainer.py:44-47 is the step it replaces)
    train_res = None
    if args.mode == "both" or (args.mode == "infer" and args.workload == "acdc_x4" and not args.no_train):
        targs = argparse.Namespace(**vars(args))
        targs.steps = min(args.steps, 10)
        sd_infer = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()} if rank == 0 else None
        train_res = bench_train(targs, dev, rank, world, distributed, barrier)
    else:
        sd_infer = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()} if rank == 0 else None

    if rank == 0:
        peaks = measured_peaks()
        frames_per_step = world * B * T_FRAMES
        value = frames_per_step * args.steps / (ms / 1e3)
        e2e_value = frames_per_step * args.steps / (ms_e2e / 1e3)
        lstm_ms, lstm_launches, lstm_flops = prof["convlstm_cell"]
        achieved = lstm_flops / (lstm_ms / 1e3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        step_flops = sum(v[2] for v in prof.values())
        tail_saved = (tail_flop_saving(B * T_FRAMES * (2 * LR_H) * (2 * LR_W), False)
                      if (SCALE == 4 and plan.lib.pvsr_get_tail_fwd()) else 0.0)
        traffic, traffic_src = ncu_traffic("convlstm_cell") if (args.workload == "acdc_x4" and B == 32) else (None, None)
        h2d = sum(x.numel() * 4 for x in inputs_h) + pos_h.numel() * 4
        d2h = ring.bytes_per_step
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"RefineNet x{SCALE} inference, {B} synthetic {shape_name}-shaped cine sequences per GPU per "
                                   f"step (LR {LR_H}x{LR_W}, T=30, U=6 -> 42 LR frames in, 30 SR frames "
                                   f"{LR_H * SCALE}x{LR_W * SCALE} out), last output list only",
                       "name": args.workload,
                       "sequences_per_gpu": B, "frames_per_step": frames_per_step, "parallelism": f"sequence-sharded x{world}",
                       "l2": "256 MiB flush write between steps; per-step working set >> 126 MB L2",
                       "cuda_graph": not args.no_graph, "algorithmic_tflop_per_step_per_gpu": step_flops / 1e12,
                       "executed_tflop_per_step_per_gpu": (step_flops - tail_saved) / 1e12},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(plan.launches * args.steps),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv3x3_halo_kernel<256, EPI_LSTM, cta_group::2> (ConvLSTM cell wavefront)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peaks["_source"] + ", sustained bf16",
                         "peak_burst": peaks.get("bf16_tflops"),      # the same source's figure for a kernel timed alone
                         "launches_per_step": lstm_launches, "avg_launch_ms": lstm_ms / max(lstm_launches, 1),
                         "share_of_step": lstm_ms / sum(v[0] for v in prof.values()),
                         "whole_step_tflops": step_flops / (ms / args.steps / 1e3) / 1e12},
            "kernel_ms_per_step": {k: round(v[0], 3) for k, v in prof.items()},
        }
        line["roofline"]["timing"] = ("per-launch CUDA events of an eager pass over the same schedule (the timed region "
                                      "replays these launches as one CUDA graph); share_of_step agrees with the ncu "
                                      "launch list of this command under profiles/")
        if sweep:
            line["batch_sweep"] = sweep
            line["batch_sweep"][str(B)] = {"sequences_per_step": B, "ms_per_step": ms / args.steps, "value": value,
                                           "unit": UNIT, "tflops": step_flops / (ms / args.steps / 1e3) / 1e12,
                                           "launches_per_step": plan.launches}
        if train_res is not None:
            line["train_step"] = {k: train_res[k] for k in ("metric", "value", "unit", "ms_per_step", "steps_per_s",
                                                            "n_gpus", "config", "e2e", "allreduce",
                                                            "algorithmic_tflop_per_step_per_gpu",
                                                            "whole_step_tflops_per_gpu",
                                                            "executed_tflop_per_step_per_gpu",
                                                            "executed_tflops_per_gpu", "flop_accounting",
                                                            "kernel_ms_per_step",
                                                            "kernel_tflops", "loss") if k in train_res}
            if "parity" in train_res:
                line["train_step"]["parity"] = train_res["parity"]
        cores = os.cpu_count() or 1
        failed = []
        if not args.no_parity:
            # sequences 0 and B-1 of the last timed step at N=1; sequence 0 only when other ranks share the host cores
            seqs = [0, B - 1] if (world == 1 and B > 1) else [0]
            par, ptimes, kind = inference_parity(frames_last, inputs_h, pos_h, hr_h, seqs, sd_infer, cores)
            line["parity"] = par
            if not par["ok"]:
                failed.append("inference")
        if train_res is not None and train_res.get("parity") and not train_res["parity"]["ok"]:
            failed.append("training")
        if world == 1 and not args.no_cpu_baseline:
            times, kind = cpu_oracle_time(3, 1, cores)       # 1 warm-up + 3 timed runs, median
            line["cpu_baseline"] = {"value": T_FRAMES / statistics.median(times), "unit": UNIT, "cores": cores,
                                    "kind": kind,
                                    "sample": f"1 {shape_name} x{SCALE} sequence (30 SR frames) per run, all 9 heads as the "
                                              "reference executes them, fp32 on the host CPU ("
                                              + ("the unmodified reference RefineNet staged under oracle/_ref"
                                                 if kind == "reference" else "pinned oracle port")
                                              + "), 1 warm-up + 3 timed runs, median"}
        print(json.dumps(line), flush=True)
        if failed:
            if distributed:
                dist.destroy_process_group()
            raise SystemExit(f"bench.py: parity check FAILED ({', '.join(failed)}): " +
                             json.dumps({"inference": line.get("parity"), "training": (train_res or {}).get("parity")}))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
