#!/usr/bin/env python
"""Vocoder throughput bench: generated audio samples/s of the HiFiGAN V1 generator on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16x3|bf16|fp32]

A step = one mel -> waveform pass over one batch (BASELINE.json configs[1]: V1 random-init, batch 16 x 10 s mels
= [16, 80, 862], fp32-class arithmetic).  N > 1 (torchrun, one rank per GPU): the batch of utterances is sharded,
every rank synthesises its own 16 utterances, no collective on the data path ("weak" scaling).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SAMPLE_RATE = 22050
METRIC = "vocoder_audio_samples_per_sec"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=16, help="utterances per GPU per step")
    ap.add_argument("--frames", type=int, default=862, help="mel frames per utterance (862 = 10 s at hop 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the separately reported bf16 tensor-core mode")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "gbs": float(d["hbm_gbs"]),
                "source": "measured (MEASURED_PEAKS.json; sustained bf16 figure: the kernel is timed inside a long step)"}
    return {"tflops": 1400.0, "gbs": 6650.0, "source": "fallback (B200_PROFILING.md: 6.65 TB/s, ~1.4 PFLOP/s sustained)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
                except ValueError:
                    continue
                for n, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


# ---------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (torch-functional restatement of the reference forward) on the host cores
# ---------------------------------------------------------------------------

def cpu_oracle_throughput(frames: int, steps: int, warmup: int):
    """Bounded sample: B=1 utterance of `frames` mel frames per step, all host threads."""
    import torch

    from oracle import hifigan_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.random_state_dict(O.V1, seed=0)
    mel = torch.from_numpy(O.synthetic_mel(1, frames, seed=1234))
    for _ in range(warmup):
        O.forward(sd, mel, O.V1)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.forward(sd, mel, O.V1)
        ts.append(time.perf_counter() - t0)
    samples = frames * O.V1.hop
    total = sum(ts)
    return {"value": samples * steps / total, "ms_per_step": 1e3 * total / steps, "best_ms": 1e3 * min(ts), "cores": cores,
            "threads": torch.get_num_threads(), "samples_per_step": samples}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    r = cpu_oracle_throughput(args.frames, args.steps, args.warmup)
    sample = (f"1 utterance x {args.frames} frames per step (1/{args.batch} of the step's batch), fp32, oracle/hifigan_oracle.py "
              f"(torch {__import__('torch').__version__} oneDNN conv1d/conv_transpose1d, weight-norm folded once), {r['threads']} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "x_realtime": r["value"] / SAMPLE_RATE,
        "config": {"workload": f"HiFiGAN V1 random-init, {args.batch} x {args.frames}-frame (10 s) mels per GPU, fp32-class",
                   "sampled_as": sample},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# Our arm
# ---------------------------------------------------------------------------

def time_device_steps(eng, stream, mel_dev, out_dev, B, T, precision, steps, warmup, barrier):
    import torch

    for _ in range(warmup):
        eng.forward_ptr(mel_dev.data_ptr(), B, T, out_dev.data_ptr(), precision, mel_on_device=True, wave_on_device=True, sync=False)
    eng.sync()
    eng.profile(True)
    l0 = eng.launch_count
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        eng.forward_ptr(mel_dev.data_ptr(), B, T, out_dev.data_ptr(), precision, mel_on_device=True, wave_on_device=True, sync=False)
    e1.record(stream)
    eng.sync()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    recs = eng.profile_records()
    eng.profile(False)
    return ms, eng.launch_count - l0, recs


def ncu_traffic(precision, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this workload (profiles/traffic_<mode>.json), or None."""
    p = os.path.join(ROOT, "profiles", f"traffic_{precision}.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(d["kernels"][kernel + "_kernel"]["traffic_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        return None


def roofline_from_records(recs, peaks, kernel=None, precision=None):
    """Roofline record of one kernel family from the per-launch event timings (hfg_profile_*).  Default: the dominant kernel,
    i.e. the tensor-core conv family with the largest share of the step (conv_umma2_kernel: the wide ResBlock convs; conv_pair_kernel:
    the fused conv1 -> conv2 ResBlock steps of the C <= 64 stages)."""
    if kernel is None:
        share = {}
        for r in recs:
            if r["kernel"] in ("conv_umma2", "conv_pair", "conv_umma"):
                share[r["kernel"]] = share.get(r["kernel"], 0.0) + r["ms"]
        if not share:
            return None
        kernel = max(share, key=share.get)
    sel = [r for r in recs if r["kernel"] == kernel]
    if not sel:
        return None
    ms = sum(r["ms"] for r in sel)
    flops = sum(r["flops"] for r in sel)
    total_ms = sum(r["ms"] for r in recs)
    achieved = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": kernel + "_kernel", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
            "frac": achieved / peaks["tflops"], "traffic": ncu_traffic(precision, kernel) if precision else None,
            "algorithmic_bytes_per_launch": sum(r["bytes"] for r in sel) / len(sel), "launches": len(sel), "avg_launch_ms": ms / len(sel),
            "share_of_step": ms / total_ms if total_ms else None, "peak_source": peaks["source"],
            # bf16x3 executes three bf16 MMAs per algorithmic one (hi*hi, lo*hi, hi*lo): the tensor pipe's own rate
            "executed_flop_factor": 3 if precision == "bf16x3" else 1,
            "executed_frac": (3 if precision == "bf16x3" else 1) * achieved / peaks["tflops"],
            "algorithmic_flops_per_launch": flops / len(sel),
            "algorithmic_gbs": sum(r["bytes"] for r in sel) / (ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks["gbs"]}


def other_rooflines(recs, peaks, precision, dominant):
    """The same record for the other tensor-core conv families of the step (explains the rest of the time)."""
    out = []
    for k in ("conv_umma2", "conv_pair", "conv_umma"):
        if dominant and dominant["kernel"] == k + "_kernel":
            continue
        r = roofline_from_records(recs, peaks, kernel=k, precision=precision)
        if r:
            out.append(r)
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    from iris_tts_b200 import build as hfg_build

    if rank == 0:
        hfg_build.build()
    barrier()
    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import work

    B, T = args.batch, args.frames
    # BASELINE config 2: the infer_hifigan path on a state dict saved from a seed-0 random-init model
    torch.manual_seed(0)
    ckpt = os.path.join(tempfile.mkdtemp(prefix="hfg_bench_"), "generator.ckpt")
    torch.save(hp.HiFiGANModel().state_dict(), ckpt)
    voc = hp.get_pretrained_hifigan(ckpt, force_reload=True)
    voc.model.precision = args.precision
    eng = voc.model.engine
    hop = eng.hop
    stream = torch.cuda.ExternalStream(eng.stream)

    torch.manual_seed(1234 + rank)
    mel_host = torch.randn(B, 80, T).numpy()
    mel_dev = torch.from_numpy(mel_host).cuda()
    out_dev = torch.empty(B, T * hop, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    peaks = measured_peaks()
    samples_per_step = B * T * hop * world

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, recs = time_device_steps(eng, stream, mel_dev, out_dev, B, T, args.precision, args.steps, args.warmup, barrier)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = samples_per_step * args.steps / (ms * 1e-3)

    # end to end through the drop-in call: numpy in -> numpy out, H2D and D2H inside the timed region
    for _ in range(max(1, min(args.warmup, 3))):
        voc(mel_host)
    barrier()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wav = voc(mel_host)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    s1.record(stream)
    torch.cuda.synchronize()
    barrier()
    e2e_ms = max(s0.elapsed_time(s1), wall_ms)
    assert wav.shape == (B, T * hop) and wav.dtype == np.float32
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = samples_per_step * args.steps / (e2e_ms * 1e-3)

    # the bf16 single-pass tensor-core mode, reported separately (BASELINE config 3)
    secondary = None
    if not args.no_secondary and args.precision != "bf16":
        ms2, _, recs2 = time_device_steps(eng, stream, mel_dev, out_dev, B, T, "bf16", args.steps, args.warmup, barrier)
        if world > 1:
            t = torch.tensor([ms2], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        rl2 = work.layer_roofline_seconds(voc.model.config, B, T, 2, peaks["tflops"] * 1e12, peaks["gbs"] * 1e9)
        secondary = {"dtype": "bf16", "value": samples_per_step * args.steps / (ms2 * 1e-3), "unit": "samples/s",
                     "ms_per_step": ms2 / args.steps, "layer_roofline_ms": rl2 * 1e3, "layer_roofline_frac": rl2 * 1e3 / (ms2 / args.steps),
                     "roofline": roofline_from_records(recs2, peaks, precision="bf16"),
                     "roofline_other_kernels": other_rooflines(recs2, peaks, "bf16", roofline_from_records(recs2, peaks, precision="bf16")),
                     "tolerance": "max-abs 1.5e-1 vs oracle on loud weights (tests/test_gpu_parity.py); 1e-3 at default init"}

    # the north-star's own target case (BASELINE.json: V1 at batch 32 x 10 s, bf16 tensor-core mode, 1 GPU: >= 50 % of the
    # per-layer roofline), reported beside the headline; single GPU only, 5 steps
    north = None
    if world == 1 and not args.no_secondary:
        B32 = 32
        mel32 = torch.randn(B32, 80, T, device="cuda")
        out32 = torch.empty(B32, T * hop, dtype=torch.float32, device="cuda")
        ms32, _, _ = time_device_steps(eng, stream, mel32, out32, B32, T, "bf16", 5, 3, barrier)
        rl32 = work.layer_roofline_seconds(voc.model.config, B32, T, 2, peaks["tflops"] * 1e12, peaks["gbs"] * 1e9)
        north = {"workload": f"HiFiGAN V1, {B32} x {T}-frame (10 s) mels, bf16, 1 GPU", "ms_per_step": ms32 / 5,
                 "value": B32 * T * hop * 5 / (ms32 * 1e-3), "unit": "samples/s", "layer_roofline_ms": rl32 * 1e3,
                 "layer_roofline_frac": rl32 * 1e3 / (ms32 / 5), "target_frac": 0.5}
        del mel32, out32

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # R_layer = sum_l max(F_l/P, Q_l/BW).  fp32-class arithmetic is rated against the TF32-class tensor peak
    # (P/2) with 4-byte activations, the bf16 mode against the bf16 peak with 2-byte activations.
    rl_bf16 = work.layer_roofline_seconds(voc.model.config, B, T, 2, peaks["tflops"] * 1e12, peaks["gbs"] * 1e9)
    rl_fp32 = work.layer_roofline_seconds(voc.model.config, B, T, 4, peaks["tflops"] * 0.5e12, peaks["gbs"] * 1e9)
    rl = rl_bf16 if args.precision == "bf16" else rl_fp32
    ms_step = ms / args.steps
    by_kernel = {}
    for r in recs:
        k = by_kernel.setdefault(r["kernel"], {"ms": 0.0, "launches": 0})
        k["ms"] += r["ms"] / args.steps
        k["launches"] += 1
    for k in by_kernel.values():
        k["launches"] //= args.steps
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3 (split-bf16 operands, 3 tcgen05 MMAs, fp32 accumulate: fp32-class)", "bf16": "bf16",
                  "fp32": "f32"}[args.precision],
        "data": "synthetic", "x_realtime": value / SAMPLE_RATE,
        "config": {"workload": f"HiFiGAN V1 random-init (seed 0) via infer_hifigan path, {B} x {T}-frame (10 s) mels per GPU, "
                               f"{args.precision}", "batch_per_gpu": B, "frames": T, "global_batch": B * world, "sharding": f"batch x{world}, no collective",
                   "l2": "no flush: each step streams ~3 GB of stage activations (452 MB per tensor) >> 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(B * 80 * T * 4),
                "d2h_bytes_per_step": int(B * T * hop * 4), "api": "iris.hifigan_pretrained.HiFiGANGenerator.__call__(np.ndarray)"},
        "gpu_launches": int(launches),
        "roofline": roofline_from_records(recs, peaks, precision=args.precision),
        "roofline_other_kernels": other_rooflines(recs, peaks, args.precision, roofline_from_records(recs, peaks, precision=args.precision)),
        "layer_roofline": {"ms": rl * 1e3, "frac": rl * 1e3 / ms_step,
                           "definition": "sum_l max(F_l/P, Q_l/BW) (SURVEY 8d): " + ("bf16 peak, 2-byte activations" if args.precision == "bf16" else
                                         "fp32-class mode: P = bf16 peak / 2 (TF32-class), 4-byte activations"),
                           "bf16_definition_ms": rl_bf16 * 1e3, "bf16_definition_frac": rl_bf16 * 1e3 / ms_step},
        "kernels_ms_per_step": by_kernel,
    }
    if secondary:
        line["bf16_mode"] = secondary
    if north:
        line["north_star_b32_bf16"] = north
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_oracle_throughput(T, 3, 2)
        line["cpu_baseline"] = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                                "sample": f"1 utterance x {T} frames, 3 timed forwards after 2 warm-ups, oracle/hifigan_oracle.py "
                                          f"(torch oneDNN fp32), {r['threads']} threads", "ms_per_utterance": r["ms_per_step"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    _, _, world = dist_env()
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
