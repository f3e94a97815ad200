#!/usr/bin/env python
"""Vocoder throughput bench: generated audio samples/s of the HiFiGAN V1 generator on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16x3|fp16|bf16|fp32]

A step = one mel -> waveform pass over one batch (BASELINE.json configs[1]: V1 random-init, batch 16 x 10 s mels
= [16, 80, 862], fp32-class arithmetic).  N > 1 (torchrun, one rank per GPU): the batch of utterances is sharded,
every rank synthesises its own 16 utterances, no collective on the data path ("weak" scaling).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for what each key means.

Beside the headline the same line carries (secondary keys, each a separately timed leg):
  bf16_mode / fp16_mode         the single-pass tensor-core modes on the headline workload
  north_star_b32                BASELINE's target case (V1, 32 x 10 s, one GPU), bf16 and fp16
  batch_sweep_bf16, v2_b64, v3_b64   BASELINE configs 3 and 5 (N = 1)
  longform_120s                 BASELINE config 4: one 10,336-frame mel; N > 1: time-chunked + ONE NCCL gather
  strong_scaling_b64            BASELINE config 3 at N > 1: a fixed global batch of 64 sharded over the N GPUs
  per_rank_ms                   min / median / max of the per-rank step times (attributes the weak-scaling loss)
"""
from __future__ import annotations

import argparse
import importlib.util
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
REF_FILE = os.path.join(ROOT, "baseline", "_ref", "iris", "hifigan_pretrained.py")
LONGFORM_FRAMES = 10336   # 120 s at hop 256 / 22050 Hz


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "fp16", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=16, help="utterances per GPU per step")
    ap.add_argument("--frames", type=int, default=862, help="mel frames per utterance (862 = 10 s at hop 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline only: skip every secondary leg")
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
# CPU baseline / reference arm: the reference's own module on the host cores (baseline/_ref, installed by
# baseline/install_ref.py); the oracle port only where that copy did not travel.
# ---------------------------------------------------------------------------

def load_reference_module():
    """The unmodified reference module, loaded by file path under an alias (it cannot shadow this repo's ``iris``)."""
    if not os.path.exists(REF_FILE):
        return None
    import warnings

    warnings.filterwarnings("ignore", category=FutureWarning)   # nn.utils.weight_norm deprecation notice
    spec = importlib.util.spec_from_file_location("_ref_hifigan", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_ref_hifigan"] = mod
    spec.loader.exec_module(mod)
    return mod


class CpuGenerator:
    """``forward(mel tensor [B, 80, T]) -> tensor`` of the seed-0 V1 generator on the host: the reference's stock
    ``HiFiGANModel.forward`` (weight-norm re-folded on every call, as the reference does) or, if baseline/_ref is absent,
    the oracle port (which also folds on every call)."""

    def __init__(self):
        import torch

        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.threads = torch.get_num_threads()
        ref = load_reference_module()
        if ref is not None:
            torch.manual_seed(0)
            self.model = ref.HiFiGANModel().eval()
            self.kind = "reference"
            self.what = (f"baseline/_ref/iris/hifigan_pretrained.py HiFiGANModel.forward (unmodified reference, torch {torch.__version__} "
                         "oneDNN, weight-norm re-folded every forward)")
            self._fwd = lambda m: self.model(m)
        else:
            from oracle import hifigan_oracle as O

            sd = O.random_state_dict(O.V1, seed=0)
            self.kind = "port"
            self.what = f"oracle/hifigan_oracle.py (port; torch {torch.__version__} oneDNN, weight-norm re-folded every forward)"
            self._fwd = lambda m: O.forward(sd, m, O.V1)

    def time(self, batch: int, frames: int, steps: int, warmup: int):
        import torch

        torch.manual_seed(1234)
        mel = torch.randn(batch, 80, frames)
        with torch.no_grad():
            for _ in range(warmup):
                self._fwd(mel)
            ts = []
            for _ in range(steps):
                t0 = time.perf_counter()
                self._fwd(mel)
                ts.append(time.perf_counter() - t0)
        samples = batch * frames * 256
        return {"value": samples * len(ts) / sum(ts), "ms_per_step": 1e3 * sum(ts) / len(ts), "best_ms": 1e3 * min(ts),
                "samples_per_step": samples}


def keras_jax_status():
    try:
        import jax  # noqa: F401
        import keras  # noqa: F401
        return "importable (not timed: the Keras generator is the same graph; see DESIGN.md)"
    except Exception as ex:  # noqa: BLE001
        return f"not runnable: {type(ex).__name__}: {ex} (keras / jax not installed in this image, no network)"


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    gen = CpuGenerator()
    B, T = args.batch, args.frames
    # one untimed full-batch forward decides whether the whole batch fits the time box (a few minutes for K + W steps)
    t0 = time.perf_counter()
    gen.time(B, T, 1, 0)
    t_full = time.perf_counter() - t0
    t0 = time.perf_counter()
    gen.time(B, T, 1, 0)                 # the second forward: oneDNN primitives are cached now
    t_full = time.perf_counter() - t0
    budget_s = 360.0
    n_steps = args.steps + args.warmup
    b_step = B if t_full * n_steps <= budget_s else max(1, int(B * budget_s / (t_full * n_steps)))
    r = gen.time(b_step, T, args.steps, args.warmup)
    same = b_step == B
    sample = (f"{b_step} of the step's {B} utterances x {T} frames per step" if not same else f"the full step: {B} utterances x {T} frames") + \
             f", fp32, {gen.what}, {gen.threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "x_realtime": r["value"] / SAMPLE_RATE,
        "config": {"workload": f"HiFiGAN V1 random-init, {B} x {T}-frame (10 s) mels per GPU, fp32-class",
                   "sampled_as": sample, "same_config": same, "batch_timed": b_step},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": gen.cores, "kind": gen.kind, "sample": sample},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# Our arm
# ---------------------------------------------------------------------------

class Timer:
    """Device timing of engine forwards with inputs resident in HBM (CUDA events on the engine's stream)."""

    def __init__(self, eng, stream, barrier, world):
        self.eng, self.stream, self.barrier, self.world = eng, stream, barrier, world

    def run(self, mel_dev, out_dev, B, T, precision, steps, warmup, profile=False):
        """(ms for `steps` forwards on this rank, launches, per-launch records if profile).  Without `profile` the production path
        runs: one CUDA graph per forward, programmatic dependent launch between the kernels, no events inside."""
        import torch

        eng = self.eng
        fwd = lambda: eng.forward_ptr(mel_dev.data_ptr(), B, T, out_dev.data_ptr(), precision, mel_on_device=True,  # noqa: E731
                                      wave_on_device=True, sync=False)
        for _ in range(max(warmup, 2)):   # >= 2: the second forward of a plan captures its graph
            fwd()
        eng.sync()
        if profile:
            eng.profile(True)
        l0 = eng.launch_count
        self.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fwd()
        e1.record(self.stream)
        eng.sync()
        torch.cuda.synchronize()
        self.barrier()
        ms = e0.elapsed_time(e1)
        recs = None
        if profile:
            recs = eng.profile_records()
            eng.profile(False)
        return ms, eng.launch_count - l0, recs

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms, [ms]
        import torch
        import torch.distributed as dist

        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        allv = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(allv, t)
        vals = [float(v.item()) for v in allv]
        return max(vals), vals


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
            "algorithmic_gbs": sum(r["bytes"] for r in sel) / (ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks["gbs"],
            "timing": "CUDA events around every launch in a SEPARATE profiled pass (the headline pass has no events inside)"}


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


def measure_tf32_peak(seconds=1.5):
    """Sustained dense TF32 throughput the way MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3 back to back): the
    denominator of the fp32-class layer roofline (SURVEY 8(d) asked for a measured TF32 peak instead of bf16 / 2)."""
    import torch

    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device="cuda")
        b = torch.randn(8192, 8192, device="cuda")
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        n, t0 = 0, time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(10):
                a @ b
            n += 10
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
        return 2.0 * 8192 ** 3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e12
    except Exception:  # noqa: BLE001
        return None


def logmel_leg(batch, samples, peaks, reps=20):
    """the log-mel front-end kernel (hfg_logmel_*, csrc/kernels_mel.cu) on `batch` waveforms of `samples` samples,
    device-resident in and out.  Algorithmic bytes = 4 B per audio sample read + the [B, 80, T] mel written (the kernel reads HBM
    exactly that once: profiles/r02d_mel_ncu_summary.txt); the FFT's ~50 flop per byte puts the op on the FP32 issue rate, not on
    HBM, so `frac` (against the HBM peak) is small by nature -- `issue_bound` says so in the line."""
    import torch

    from iris_tts_b200.mel import LogMel

    fe = LogMel()
    audio = torch.randn(batch, samples, device="cuda") * 0.1
    T = fe.frames(samples)
    out = torch.empty(batch, 80, T, device="cuda")
    for _ in range(3):
        fe.forward_ptr(audio.data_ptr(), batch, samples, out.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fe.forward_ptr(audio.data_ptr(), batch, samples, out.data_ptr())     # each call synchronises its stream
    ms = 1e3 * (time.perf_counter() - t0) / reps
    nbytes = batch * samples * 4 + out.numel() * 4
    fe.close()
    return {"workload": f"log-mel of {batch} x {samples} samples (n_fft 1024, hop 256, 80 mels), device-resident, wall clock per synchronous call",
            "ms": ms, "value": batch * samples / (ms * 1e-3), "unit": "audio samples/s", "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9,
            "hbm_peak_gbs": peaks["gbs"], "frac": nbytes / (ms * 1e-3) / 1e9 / peaks["gbs"], "frames": T,
            "issue_bound": "shared-memory FFT on CUDA cores: 70 % issue-active under ncu, 2 % of L2, DRAM read = the audio once"}


def griffin_lim_leg(frames, n_iter=60, reps=3):
    """SURVEY 8(f) f3, the CLI's no-vocoder fallback (reference scripts/synthesize.py:174-194): Griffin-Lim, `n_iter` iterations on
    one utterance of `frames` frames, host magnitudes in / host waveform out through hfg_griffin_lim (inverse STFT, envelope
    normalisation and forward STFT + momentum phase update kernels, csrc/kernels_mel.cu).  Beside it the float64 oracle
    (oracle/griffinlim_oracle.py, numpy FFTs) on 5 iterations, scaled to `n_iter`."""
    import numpy as np

    from iris_tts_b200.griffin_lim import griffin_lim
    from oracle import griffinlim_oracle as G

    rng = np.random.default_rng(3)
    y = rng.standard_normal(256 * (frames - 1)) * 0.1
    S = np.abs(G.stft(y))
    ang = np.exp(2j * np.pi * rng.random(S.shape))
    griffin_lim(S, n_iter=2, angles0=ang)
    t0 = time.perf_counter()
    for _ in range(reps):
        wav = griffin_lim(S, n_iter=n_iter, angles0=ang)
    ms = 1e3 * (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    G.griffinlim(S, ang, n_iter=5)
    cpu_ms = 1e3 * (time.perf_counter() - t0) * (n_iter + 1) / 6.0
    inc = float(np.abs(np.abs(G.stft(wav.astype(np.float64))) - S).mean() / S.mean())
    return {"workload": f"Griffin-Lim, {n_iter} iterations, 1 x {frames} frames (n_fft 1024, hop 256), numpy in / numpy out",
            "ms": ms, "ms_per_iteration": ms / (n_iter + 1), "oracle_cpu_ms_scaled_from_5_iterations": cpu_ms,
            "spectral_inconsistency": inc}


def ragged_leg(voc, max_frames, n_utt=32, reps=3):
    """32 utterances of 32 DISTINCT lengths (300 .. max_frames frames) end to end (numpy in, numpy out) in the
    bf16 mode, three ways that produce the same bits: the engine's native ragged plan (hfg_forward_ragged: one padded call per
    length bucket, every item ended where it ends), the dense-call scheme for plain callables (a padded body pass per bucket + one
    tail pass; HFG_RAGGED=0), and the reference's only option, one batch-1 call per utterance."""
    import numpy as np

    from iris_tts_b200.batching import synthesize_variable

    old = voc.model.precision
    voc.model.precision = "bf16"
    rng = np.random.default_rng(7)
    lengths = sorted(set(int(x) for x in np.linspace(300, max_frames, n_utt)))
    mels = [rng.standard_normal((80, t)).astype(np.float32) for t in lengths]
    b = [voc(m) for m in mels]

    def timed(env, max_pad=0.15):
        prev = os.environ.get("HFG_RAGGED")
        os.environ["HFG_RAGGED"] = env
        try:
            stats = {}
            a = synthesize_variable(voc, mels, stats=stats, length_quantum=64, max_pad=max_pad)      # builds the plans
            synthesize_variable(voc, mels, length_quantum=64, max_pad=max_pad)                         # captures their graphs
            same = all(np.array_equal(x, y) for x, y in zip(a, b))
            t0 = time.perf_counter()
            for _ in range(reps):
                synthesize_variable(voc, mels, length_quantum=64, max_pad=max_pad)
            return 1e3 * (time.perf_counter() - t0) / reps, stats, same
        finally:
            if prev is None:
                os.environ.pop("HFG_RAGGED", None)
            else:
                os.environ["HFG_RAGGED"] = prev

    ms_n, st_n, same_n = timed("1")
    ms_b, st_b, same_b = timed("0")
    # the native plan skips the tiles behind an item's end, so wider buckets (more padding, fewer and larger calls) cost little
    by_pad = {}
    for mp in (0.3, 0.5, 1.0):
        ms_p, st_p, same_p = timed("1", mp)
        by_pad[str(mp)] = {"ms": ms_p, "calls": st_p["calls"], "frames_run": st_p["frames_run"], "bit_identical": bool(same_p)}
    # what the ragged plan's extra launches (one small zero-fill behind every conv) cost: the same padded batch with every length
    # equal to T through both plans, synchronous calls, numpy in / numpy out
    eng = voc.model.engine
    pad = np.zeros((5, 80, 640), dtype=np.float32)
    pad[:] = rng.standard_normal(pad.shape)
    over = {}
    for name, fn in (("dense", lambda: eng.forward(pad, precision="bf16")), ("ragged", lambda: eng.forward_ragged(pad, [640] * 5, precision="bf16"))):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        over[name] = 1e3 * (time.perf_counter() - t0) / 10
    t0 = time.perf_counter()
    for _ in range(reps):
        for m in mels:
            voc(m)
    ms_u = 1e3 * (time.perf_counter() - t0) / reps
    voc.model.precision = old
    samples = sum(lengths) * 256
    return {"workload": f"{len(lengths)} utterances, {len(lengths)} distinct lengths {lengths[0]}..{lengths[-1]} frames, bf16, numpy in -> numpy out",
            "native_ragged_ms": ms_n, "native_ragged_value": samples / (ms_n * 1e-3), "native_calls": st_n["calls"],
            "native_frames_run": st_n["frames_run"], "native_path_taken": bool(st_n.get("native_ragged")),
            "native_by_max_pad": by_pad,
            "bucketed_ms": ms_b, "bucketed_value": samples / (ms_b * 1e-3), "per_utterance_calls_ms": ms_u,
            "per_utterance_value": samples / (ms_u * 1e-3), "unit": "samples/s", "speedup": ms_u / ms_n,
            "speedup_dense_call_scheme": ms_u / ms_b, "dense_calls": st_b["calls"],
            "frames_real": st_b["frames_real"], "frames_run": st_b["frames_run"],
            "bit_identical_to_per_utterance": bool(same_n and same_b),
            "ragged_plan_overhead": {"workload": "5 x 640 frames, all lengths 640, bf16, synchronous call", "dense_ms": over["dense"],
                                     "ragged_ms": over["ragged"]}}


def longform_leg(model, precision, world, rank, reps=5):
    """BASELINE config 4: one 120 s mel (10,336 frames).  N > 1: chunks of T/N frames + halo per rank, ONE NCCL gather to rank 0
    (iris_tts_b200.sharding.synthesize_longform); N = 1: the unchunked forward.  Device-timed, max over ranks."""
    import torch
    import torch.distributed as dist

    from iris_tts_b200 import sharding

    old = model.precision
    model.precision = precision
    torch.manual_seed(4321)
    mel = (torch.randn(1, 80, LONGFORM_FRAMES) * 2.0 - 5.0).cuda()
    halo = sharding.HALO_FRAMES
    assert halo >= sharding.halo_frames(model.config)
    synth = lambda m: model(m)  # noqa: E731   (CUDA tensor in -> CUDA tensor out, no host round trip)
    out = None
    for _ in range(3):
        out = sharding.synthesize_longform(synth, mel, hop=256, halo=halo)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = sharding.synthesize_longform(synth, mel, hop=256, halo=halo)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    rec = None
    if rank == 0:
        full = model(mel).reshape(-1)
        chunks = sharding.time_chunks(LONGFORM_FRAMES, world, halo)
        rec = {"workload": f"one {LONGFORM_FRAMES}-frame (120 s) mel over {world} GPU(s), {precision}" +
                           (f", {halo}-frame halo, one NCCL gather" if world > 1 else ", unchunked"),
               "ms": float(ms.item()), "value": LONGFORM_FRAMES * 256 / (float(ms.item()) * 1e-3), "unit": "samples/s",
               "x_realtime": LONGFORM_FRAMES * 256 / (float(ms.item()) * 1e-3) / SAMPLE_RATE,
               "gather_bytes_per_rank": (max(c.frames for c in chunks) * 256 * 4) if world > 1 else 0,
               "max_abs_vs_unchunked": float((out - full).abs().max()), "samples": int(out.numel())}
        del full
    model.precision = old
    del mel
    return rec


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    affinity = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # one process per GPU: run on (and allocate page-locked staging from) the GPU's own NUMA node
        from iris_tts_b200 import numa

        if os.environ.get("HFG_BIND_NUMA", "1") != "0":
            affinity = numa.bind_process_to_gpu(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()

    from iris_tts_b200 import build as hfg_build

    if rank == 0:
        hfg_build.build()
    barrier()
    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import engine as E
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
    timer = Timer(eng, stream, barrier, world)

    torch.manual_seed(1234 + rank)
    mel_host = torch.randn(B, 80, T).numpy()
    mel_dev = torch.from_numpy(mel_host).cuda()
    out_dev = torch.empty(B, T * hop, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    peaks = measured_peaks()
    samples_per_step = B * T * hop * world
    P, BW = peaks["tflops"] * 1e12, peaks["gbs"] * 1e9

    # ---- headline: device-timed, production path (graph + PDL, no events inside the timed region) ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_rank, launches, _ = timer.run(mel_dev, out_dev, B, T, args.precision, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms, ms_all = timer.max_over_ranks(ms_rank)
    value = samples_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the drop-in call: numpy in -> numpy out, H2D and D2H inside the timed region ----
    # Warm-up in the timed loop's own steady state: the caller keeps the previous waveform while asking for the next one, so the
    # page-locked allocator alternates between TWO 14 MB blocks -- and the first cudaHostAlloc of such a block costs 8-12 ms on this
    # pool's hosts (tools/e2e_loop_probe.py; a warm-up that drops its results creates only one of them).
    wav = None
    for _ in range(max(3, min(args.warmup, 5))):
        wav = voc(mel_host)
    barrier()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler2 = ClockSampler(local_rank) if rank == 0 else None
    if sampler2:
        sampler2.start()
    s0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wav = voc(mel_host)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    s1.record(stream)
    torch.cuda.synchronize()
    e2e_clocks = sampler2.stop() if sampler2 else None
    barrier()
    e2e_ms = max(s0.elapsed_time(s1), wall_ms)
    assert wav.shape == (B, T * hop) and wav.dtype == np.float32
    e2e_ms, e2e_all = timer.max_over_ranks(e2e_ms)
    e2e_value = samples_per_step * args.steps / (e2e_ms * 1e-3)

    # what the host side of that call costs on THIS box (boxes of the pool differ: the same code has shown 1 ms and 7 ms of gap):
    # raw page-locked copies of the step's buffers and the staging copy, each timed alone
    pin_out = torch.empty((B, T * hop), dtype=torch.float32, pin_memory=True)
    pin_in = torch.empty((B, 80, T), dtype=torch.float32, pin_memory=True)
    cs = torch.cuda.Stream()
    host_io = {}
    with torch.cuda.stream(cs):
        for name, fn, nbytes in (("d2h_gbs", lambda: pin_out.copy_(out_dev, non_blocking=True), pin_out.numel() * 4),
                                 ("h2d_gbs", lambda: mel_dev.copy_(pin_in, non_blocking=True), pin_in.numel() * 4)):
            fn()
            cs.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(cs)
            for _ in range(5):
                fn()
            c1.record(cs)
            cs.synchronize()
            host_io[name] = nbytes * 5 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    t0 = time.perf_counter()
    for _ in range(5):
        np.copyto(pin_in.numpy(), mel_host, casting="unsafe")
    host_io["staging_copy_gbs"] = mel_host.nbytes * 5 / (time.perf_counter() - t0) / 1e9
    del pin_out, pin_in

    # ---- per-launch records for the roofline: a separate profiled pass (events around every launch, no graph) ----
    prof_steps = max(1, min(args.steps, 5))
    ms_prof, _, recs = timer.run(mel_dev, out_dev, B, T, args.precision, prof_steps, 1, profile=True)

    def mode_leg(mode, batch, mel_d, out_d, steps, warm, cfg):
        """One single-pass tensor-core mode on (batch, T): production timing + a profiled pass for its kernel records."""
        m, _, _ = timer.run(mel_d, out_d, batch, T, mode, steps, warm)
        m, _ = timer.max_over_ranks(m)
        rl = work.layer_roofline_seconds(cfg, batch, T, 2, P, BW)
        return m / steps, rl

    secondary = {}
    north = None
    sweep = None
    small = {}
    longform = None
    strong = None
    tf32_peak = None
    logmel = None
    glim = None
    ragged = None
    if not args.no_secondary:
        # the single-pass tensor-core modes on the headline workload (BASELINE config 3's mode), reported separately
        for mode in ("bf16", "fp16"):
            if mode == args.precision:
                continue
            ms_m, rl = mode_leg(mode, B, mel_dev, out_dev, args.steps, args.warmup, voc.model.config)
            _, _, recs_m = timer.run(mel_dev, out_dev, B, T, mode, 2, 1, profile=True)
            dom = roofline_from_records(recs_m, peaks, precision=mode)
            secondary[mode] = {
                "dtype": mode, "value": samples_per_step / (ms_m * 1e-3), "unit": "samples/s", "ms_per_step": ms_m,
                "layer_roofline_ms": rl * 1e3, "layer_roofline_frac": rl * 1e3 / ms_m, "roofline": dom,
                "roofline_other_kernels": other_rooflines(recs_m, peaks, mode, dom),
                "tolerance": {"bf16": "max-abs <= 0.15 x output std vs oracle on loud weights (measured 0.064-0.10 std: 1.5e-2 at std 0.23, "
                                      "B=32 x 10 s), <= 1e-3 at default init (tests/test_gpu_parity.py, test_gpu_north_star.py)",
                              "fp16": "max-abs <= 0.025 x output std vs oracle on loud weights (measured 0.006-0.012 std: 2e-3 at std 0.23; "
                                      "TF32-class), <= 1e-3 at default init (tests/test_gpu_parity.py, test_gpu_north_star.py)"}[mode]}
        # the north-star's own target case (BASELINE.json: V1 at batch 32 x 10 s, 16-bit tensor-core mode, 1 GPU: >= 50 % of the
        # per-layer roofline); single GPU only
        if world == 1:
            # BASELINE config 3 at N = 1: batch sweep in the bf16 tensor-core mode.  Ascending and BEFORE the long legs: the
            # latency end of the sweep (1 ms per forward) is timed before the power cap has pulled the clocks down.
            sweep = []
            torch.cuda.synchronize()
            time.sleep(0.5)
            for b in (1, 2, 4, 8, 16, 32, 64):
                mel_b = torch.randn(b, 80, T, device="cuda")
                out_b = torch.empty(b, T * hop, dtype=torch.float32, device="cuda")
                nst = 40 if b <= 2 else (16 if b <= 8 else (8 if b <= 16 else 5))
                ms_m, rl = mode_leg("bf16", b, mel_b, out_b, nst, 3, voc.model.config)
                sweep.append({"batch": b, "ms_per_step": ms_m, "value": b * T * hop / (ms_m * 1e-3), "layer_roofline_frac": rl * 1e3 / ms_m, "steps": nst})
                del mel_b, out_b
            B32 = 32
            mel32 = torch.randn(B32, 80, T, device="cuda")
            out32 = torch.empty(B32, T * hop, dtype=torch.float32, device="cuda")
            north = {"workload": f"HiFiGAN V1, {B32} x {T}-frame (10 s) mels, 1 GPU", "target_frac": 0.5, "steps": 8}
            for mode in ("bf16", "fp16"):
                ms_m, rl = mode_leg(mode, B32, mel32, out32, 8, 3, voc.model.config)
                north[mode] = {"ms_per_step": ms_m, "value": B32 * T * hop / (ms_m * 1e-3), "unit": "samples/s",
                               "layer_roofline_ms": rl * 1e3, "layer_roofline_frac": rl * 1e3 / ms_m}
            del mel32, out32
            # BASELINE config 5: the small generators at batch 64 (memory-bound regime)
            for name, cfg in (("v2_b64", E.V2), ("v3_b64", E.V3)):
                torch.manual_seed(0)
                kw = dict(upsample_rates=list(cfg.upsample_rates), upsample_kernel_sizes=list(cfg.upsample_kernel_sizes),
                          upsample_initial_channel=cfg.upsample_initial_channel, resblock_kernel_sizes=list(cfg.resblock_kernel_sizes),
                          resblock_dilation_sizes=[list(d) for d in cfg.resblock_dilation_sizes])
                m2 = hp.HiFiGANModel(**kw)
                m2.eval().to(f"cuda:{local_rank}")
                e2 = m2.engine
                t2 = Timer(e2, torch.cuda.ExternalStream(e2.stream), barrier, world)
                mel_b = torch.randn(64, 80, T, device="cuda")
                out_b = torch.empty(64, T * e2.hop, dtype=torch.float32, device="cuda")
                rec = {"workload": f"{name.split('_')[0].upper()} ({'C0=128' if name.startswith('v2') else 'C0=256, rates 8/8/4'}), 64 x {T} frames, 1 GPU"}
                for mode in ("bf16", "fp16", "bf16x3"):
                    m, _, _ = t2.run(mel_b, out_b, 64, T, mode, 5, 3)
                    rl = work.layer_roofline_seconds(cfg, 64, T, 2 if mode != "bf16x3" else 4, P if mode != "bf16x3" else P / 2, BW)
                    rec[mode] = {"ms_per_step": m / 5, "value": 64 * T * e2.hop / (m / 5 * 1e-3), "layer_roofline_ms": rl * 1e3,
                                 "layer_roofline_frac": rl * 1e3 / (m / 5)}
                small[name] = rec
                del mel_b, out_b
                e2.close()
            tf32_peak = measure_tf32_peak()
            logmel = logmel_leg(B, T * hop, peaks)
            ragged = ragged_leg(voc, T)
            glim = griffin_lim_leg(T)
            voc.model.precision = args.precision
        else:
            # BASELINE config 3 at N > 1: a FIXED global batch of 64 utterances sharded over the N GPUs (strong scaling)
            from iris_tts_b200 import sharding

            s, e = sharding.batch_shards(64, world)[rank]
            bl = e - s
            mel_b = torch.randn(max(bl, 1), 80, T, device="cuda")
            out_b = torch.empty(max(bl, 1), T * hop, dtype=torch.float32, device="cuda")
            strong = {"workload": f"HiFiGAN V1, global batch 64 x {T} frames sharded over {world} GPUs ({64 // world} per GPU), no collective"}
            for mode in ("bf16", args.precision):
                m, _, _ = timer.run(mel_b, out_b, max(bl, 1), T, mode, 5, 3)
                m, _ = timer.max_over_ranks(m)
                strong[mode] = {"ms_per_step": m / 5, "value": 64 * T * hop / (m / 5 * 1e-3), "unit": "samples/s"}
            del mel_b, out_b
        # BASELINE config 4: the 120 s mel (every rank takes part)
        longform = {}
        for mode in ("bf16x3", "bf16"):
            longform[mode] = longform_leg(voc.model, mode, world, rank)
        voc.model.precision = args.precision

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # R_layer = sum_l max(F_l/P, Q_l/BW).  fp32-class arithmetic is rated against the TF32-class tensor peak
    # with 4-byte activations (measured here when possible, else bf16 peak / 2), the 16-bit modes against the bf16 peak, 2 bytes.
    rl_bf16 = work.layer_roofline_seconds(voc.model.config, B, T, 2, P, BW)
    rl_fp32 = work.layer_roofline_seconds(voc.model.config, B, T, 4, P * 0.5, BW)
    rl_fp32_meas = work.layer_roofline_seconds(voc.model.config, B, T, 4, tf32_peak * 1e12, BW) if tf32_peak else None
    rl = rl_bf16 if args.precision in ("bf16", "fp16") else rl_fp32
    ms_step = ms / args.steps
    by_kernel = {}
    for r in recs:
        k = by_kernel.setdefault(r["kernel"], {"ms": 0.0, "launches": 0})
        k["ms"] += r["ms"] / prof_steps
        k["launches"] += 1
    for k in by_kernel.values():
        k["launches"] //= prof_steps
    dom = roofline_from_records(recs, peaks, precision=args.precision)
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3 (split-bf16 operands, 3 tcgen05 MMAs, fp32 accumulate: fp32-class)", "bf16": "bf16", "fp16": "f16",
                  "fp32": "f32"}[args.precision],
        "data": "synthetic", "x_realtime": value / SAMPLE_RATE,
        "config": {"workload": f"HiFiGAN V1 random-init (seed 0) via infer_hifigan path, {B} x {T}-frame (10 s) mels per GPU, "
                               f"{args.precision}", "batch_per_gpu": B, "frames": T, "global_batch": B * world, "sharding": f"batch x{world}, no collective",
                   "l2": "no flush: each step streams ~3 GB of stage activations (452 MB per tensor) >> 126 MB L2",
                   "timed_path": "production: one CUDA graph per forward, programmatic dependent launch, no events inside the timed region",
                   "rank0_cpu_affinity": (f"{len(affinity)} CPUs local to the GPU (NVML)" if affinity else "unchanged")},
        "clocks": clocks,
        "per_rank_ms": {"min": min(ms_all) / args.steps, "median": statistics.median(ms_all) / args.steps, "max": max(ms_all) / args.steps,
                        "all": [m / args.steps for m in ms_all],
                        "e2e_all": [m / args.steps for m in e2e_all]},
        "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(B * 80 * T * 4),
                "d2h_bytes_per_step": int(B * T * hop * 4), "api": "iris.hifigan_pretrained.HiFiGANGenerator.__call__(np.ndarray)",
                "gap_vs_device": e2e_ms / ms - 1.0, "host_io_rank0": host_io, "clocks": e2e_clocks},
        "gpu_launches": int(launches),
        "roofline": dom,
        "roofline_other_kernels": other_rooflines(recs, peaks, args.precision, dom),
        "layer_roofline": {"ms": rl * 1e3, "frac": rl * 1e3 / ms_step,
                           "definition": "sum_l max(F_l/P, Q_l/BW) (SURVEY 8d): " + ("bf16 peak, 2-byte activations" if args.precision in ("bf16", "fp16") else
                                         "fp32-class mode: P = bf16 peak / 2 (TF32-class, SURVEY's provisional figure), 4-byte activations"),
                           "bf16_definition_ms": rl_bf16 * 1e3, "bf16_definition_frac": rl_bf16 * 1e3 / ms_step,
                           "tf32_peak_measured_tflops": tf32_peak,
                           "measured_tf32_definition_ms": rl_fp32_meas * 1e3 if rl_fp32_meas else None,
                           "measured_tf32_definition_frac": rl_fp32_meas * 1e3 / ms_step if rl_fp32_meas else None},
        "kernels_ms_per_step": by_kernel,
        "profiled_pass_ms_per_step": ms_prof / prof_steps,
    }
    if "bf16" in secondary:
        line["bf16_mode"] = secondary["bf16"]
    if "fp16" in secondary:
        line["fp16_mode"] = secondary["fp16"]
    if north:
        line["north_star_b32"] = north
        line["north_star_b32_bf16"] = dict(north["bf16"], workload=north["workload"] + ", bf16", target_frac=0.5)
    if sweep:
        line["batch_sweep_bf16"] = sweep
    line.update(small)
    if strong:
        line["strong_scaling_b64"] = strong
    if longform:
        line["longform_120s"] = longform
    if logmel:
        line["logmel_frontend"] = logmel
    if glim:
        line["griffin_lim_fallback"] = glim
    if ragged:
        line["ragged_batch"] = ragged
    if world == 1 and not args.no_cpu_baseline:
        gen = CpuGenerator()
        r = gen.time(1, T, 3, 2)
        r1 = gen.time(1, 256, 3, 2)
        line["cpu_baseline"] = {"value": r["value"], "unit": "samples/s", "cores": gen.cores, "kind": gen.kind,
                                "sample": f"1 utterance x {T} frames, 3 timed forwards after 2 warm-ups, {gen.what}, {gen.threads} threads",
                                "ms_per_utterance": r["ms_per_step"],
                                "config1_b1_t256": {"value": r1["value"], "ms": r1["ms_per_step"], "x_realtime": r1["value"] / SAMPLE_RATE,
                                                    "what": "BASELINE config 1 shape (B = 1, 256 frames) on the torch path"},
                                "keras_jax": keras_jax_status()}
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
