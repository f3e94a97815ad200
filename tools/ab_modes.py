"""Interleaved A/B of the three tensor-core modes on one box (production path: graph launches), with the SM clock and board power sampled
during each burst -- separates a format's own cost from the power cap's clock response.  python tools/ab_modes.py [B] [T]"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iris.hifigan_pretrained as hp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 862
torch.manual_seed(0)
m = hp.HiFiGANModel(); m.to("cuda:0"); eng = m.engine
mel = torch.randn(B, 80, T, device="cuda"); out = torch.empty(B, T * eng.hop, device="cuda")
stream = torch.cuda.ExternalStream(eng.stream)


def sample(stop, acc):
    while not stop.is_set():
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=5).stdout
            c, p = o.strip().split(",")
            acc.append((float(c), float(p)))
        except Exception:
            pass
        time.sleep(0.05)


def burst(mode, n=60):
    for _ in range(5):
        eng.forward_ptr(mel.data_ptr(), B, T, out.data_ptr(), mode, mel_on_device=True, wave_on_device=True, sync=False)
    eng.sync()
    stop, acc = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, acc)); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(n):
            eng.forward_ptr(mel.data_ptr(), B, T, out.data_ptr(), mode, mel_on_device=True, wave_on_device=True, sync=False)
        e1.record()
    eng.sync(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = e0.elapsed_time(e1) / n
    clk = sorted(a for a, _ in acc)[len(acc) // 2] if acc else 0
    pw = max((b for _, b in acc), default=0)
    return ms, clk, pw


for rnd in range(3):
    for mode in ("bf16", "fp16", "bf16x3", "fp16", "bf16"):
        ms, clk, pw = burst(mode, 60 if mode != "bf16x3" else 25)
        print(f"round {rnd} {mode:7s} {ms:7.3f} ms   sm {clk:6.0f} MHz   power max {pw:5.0f} W   ms*GHz {ms * clk / 1e3:6.2f}")
