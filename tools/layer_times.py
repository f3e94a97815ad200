"""Per-layer device times of one forward (CUDA events around every launch, hfg_profile_*), with the
algorithmic TFLOP/s and GB/s of each launch.  Development / profiling tool.

    python tools/layer_times.py --mode bf16 --B 16 --T 862 [--reps 3] [--csv out.csv]
    HFG_NCU_LAYERS=resblocks.2.convs1.2 ncu --profile-from-start off ... python tools/layer_times.py --reps 1 --warm 1
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--cfg", default="v1")
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--T", type=int, default=862)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--csv", default=None)
    a = ap.parse_args()
    import torch

    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import engine as E

    cfg = {"v1": E.V1, "v2": E.V2, "v3": E.V3}[a.cfg]
    torch.manual_seed(0)
    m = hp.HiFiGANModel(cfg.in_channels, list(cfg.upsample_rates), list(cfg.upsample_kernel_sizes), cfg.upsample_initial_channel,
                        list(cfg.resblock_kernel_sizes), [list(d) for d in cfg.resblock_dilation_sizes])
    m.to("cuda:0")
    eng = m.engine
    torch.manual_seed(1234)
    mel = torch.randn(a.B, cfg.in_channels, a.T, device="cuda")
    out = torch.empty(a.B, a.T * eng.hop, device="cuda")
    torch.cuda.synchronize()

    def fwd():
        eng.forward_ptr(mel.data_ptr(), a.B, a.T, out.data_ptr(), a.mode, mel_on_device=True, wave_on_device=True, sync=False)

    for _ in range(a.warm):
        fwd()
    eng.sync()
    eng.profile(True)
    for _ in range(a.reps):
        fwd()
    eng.sync()
    recs = eng.profile_records()
    n = len(recs) // a.reps
    rows = []
    for i in range(n):
        ms = sum(recs[i + r * n]["ms"] for r in range(a.reps)) / a.reps
        r = recs[i]
        rows.append((r["layer"], r["kernel"], ms, r["flops"], r["bytes"]))
    tot = sum(r[2] for r in rows)
    print(f"# {a.cfg} mode {a.mode} B={a.B} T={a.T}: {tot:.3f} ms per forward, {a.B * a.T * eng.hop / tot / 1e3:.1f} M samples/s")
    print(f"{'layer':28s} {'kernel':14s} {'ms':>8s} {'%':>6s} {'TFLOP/s':>9s} {'GB/s(alg)':>10s}")
    for layer, ker, ms, fl, by in rows:
        print(f"{layer:28s} {ker:14s} {ms:8.4f} {100 * ms / tot:6.2f} {fl / ms / 1e9 if ms else 0:9.1f} {by / ms / 1e6 if ms else 0:10.1f}")
    if a.csv:
        with open(a.csv, "w") as f:
            f.write("layer,kernel,ms,flops,bytes\n")
            for row in rows:
                f.write(",".join(str(x) for x in row) + "\n")


if __name__ == "__main__":
    main()
