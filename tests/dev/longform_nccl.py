"""BASELINE config 4 on real GPUs: one long mel split along time across the ranks (16-frame halo), ONE NCCL gather.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tests/dev/longform_nccl.py [--frames 10336] [--precision bf16x3]

Rank 0 checks the stitched waveform against the unchunked forward on its own GPU (<= 2e-5) and, for the first
4 s, against the CPU oracle (<= 1e-3), and prints one JSON line with device-timed (max over ranks) throughput.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10336)   # 120 s at hop 256 / 22050 Hz
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist

    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import sharding

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    model = hp.HiFiGANModel()
    model.precision = a.precision
    model.eval().to(f"cuda:{local}")
    torch.manual_seed(1234)
    mel = (torch.randn(1, 80, a.frames) * 2.0 - 5.0).cuda()
    synth = lambda m: model(m)  # noqa: E731   (CUDA tensor in -> CUDA tensor out, no host round trip)
    out = sharding.synthesize_longform(synth, mel, hop=256)        # warm-up + result
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        out = sharding.synthesize_longform(synth, mel, hop=256)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.reps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = model(mel).reshape(-1)
        err_full = float((out - full).abs().max())
        from oracle import hifigan_oracle as O
        sd = {k: v.cpu() for k, v in model.state_dict().items()}
        n = min(a.frames, 345)
        ref = O.forward(sd, mel[:, :, :n + 16].cpu(), O.V1).reshape(-1)[: n * 256]
        err_ref = float((out[: n * 256].cpu() - ref).abs().max())
        chunks = sharding.time_chunks(a.frames, world)
        print(json.dumps({"config": f"long-form {a.frames} frames ({a.frames * 256 / 22050:.1f} s) over {world} GPU(s), halo 16, one NCCL gather",
                          "precision": a.precision, "ms": float(ms.item()), "samples_per_s": a.frames * 256 / (float(ms.item()) * 1e-3),
                          "gather_bytes_per_rank": max(c.frames for c in chunks) * 256 * 4,
                          "max_abs_vs_unchunked": err_full, "max_abs_vs_oracle_first_4s": err_ref}), flush=True)
        assert out.numel() == a.frames * 256
        assert err_full <= 2e-5, err_full
        assert err_ref <= 1e-3, err_ref
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
