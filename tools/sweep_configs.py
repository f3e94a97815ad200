"""BASELINE configs 3 and 5 on one GPU: V1 batch sweep 1..64 x 10 s (bf16 and bf16x3), V2 / V3-args at batch 64.
Device-resident mel, CUDA events on the engine's stream, 3 warm-ups + 5 timed forwards.  Prints one JSON line per case.

    python tools/sweep_configs.py [--batches 1,2,4,8,16,32,64] [--modes bf16,bf16x3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,2,4,8,16,32,64")
    ap.add_argument("--modes", default="bf16,bf16x3")
    ap.add_argument("--frames", type=int, default=862)
    ap.add_argument("--v1-only", action="store_true")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch

    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import engine as E
    from iris_tts_b200 import work

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    P = float(peaks.get("bf16_tflops_sustained", 1400.0)) * 1e12
    BW = float(peaks.get("hbm_gbs", 6650.0)) * 1e9

    def model_for(cfg):
        torch.manual_seed(0)
        m = hp.HiFiGANModel(cfg.in_channels, list(cfg.upsample_rates), list(cfg.upsample_kernel_sizes), cfg.upsample_initial_channel,
                            list(cfg.resblock_kernel_sizes), [list(d) for d in cfg.resblock_dilation_sizes])
        m.to("cuda:0")
        return m

    def run(name, cfg, m, B, mode):
        eng = m.engine
        T = a.frames
        torch.manual_seed(1234)
        mel = torch.randn(B, cfg.in_channels, T, device="cuda")
        out = torch.empty(B, T * eng.hop, device="cuda")
        st = torch.cuda.ExternalStream(eng.stream)
        torch.cuda.synchronize()

        def fwd():
            eng.forward_ptr(mel.data_ptr(), B, T, out.data_ptr(), mode, mel_on_device=True, wave_on_device=True, sync=False)

        for _ in range(3):
            fwd()
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            fwd()
        e1.record(st)
        eng.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        rl = work.layer_roofline_seconds(cfg, B, T, 2, P, BW) * 1e3
        print(json.dumps({"tag": a.tag, "model": name, "batch": B, "frames": T, "mode": mode, "ms": round(ms, 4),
                          "samples_per_s": round(B * T * eng.hop / (ms * 1e-3)), "x_realtime": round(B * T * eng.hop / (ms * 1e-3) / 22050, 1),
                          "layer_roofline_bf16_ms": round(rl, 4), "frac_of_bf16_layer_roofline": round(rl / ms, 4),
                          "workspace_gb": round(eng.workspace_bytes(B, T, mode) / 2**30, 2)}), flush=True)

    m1 = model_for(E.V1)
    for mode in a.modes.split(","):
        for B in [int(x) for x in a.batches.split(",")]:
            run("V1", E.V1, m1, B, mode)
    del m1
    if a.v1_only:
        return
    for name, cfg in (("V2", E.V2), ("V3-args", E.V3)):
        m = model_for(cfg)
        for mode in a.modes.split(","):
            run(name, cfg, m, 64, mode)
        del m


if __name__ == "__main__":
    main()
