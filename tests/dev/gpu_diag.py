"""GPU diagnostic: per-layer and end-to-end error of every precision mode against the CPU oracle.

Development tool (not a test, not the product): prints a table instead of asserting, and runs each
mode in its own subprocess so a trapped kernel in one mode cannot poison the CUDA context of the
next.  Uses oracle/ as the checker, like tests/ do.

    python tests/dev/gpu_diag.py                 # all modes
    python tests/dev/gpu_diag.py --one bf16x3    # one mode, in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def layer_cases(cfg_name):
    if cfg_name == "v1":
        return ["conv_pre", "ups.0", "resblocks.0.convs1.0", "resblocks.1.convs1.1", "resblocks.2.convs1.2",
                "resblocks.2.convs2.2", "ups.1", "resblocks.3.convs1.1", "resblocks.5.convs1.2", "ups.2",
                "resblocks.6.convs1.0", "resblocks.8.convs1.2", "ups.3", "resblocks.9.convs2.0", "resblocks.11.convs1.2",
                "conv_post"]
    return ["conv_pre", "ups.0", "resblocks.0.convs1.0", "ups.1", "resblocks.3.convs1.1", "ups.2", "resblocks.8.convs2.1"]


def run_one(mode: str, cfg_name: str, B: int, L: int) -> int:
    import torch
    import torch.nn.functional as F

    from iris_tts_b200 import Engine
    from iris_tts_b200 import engine as E
    from oracle import hifigan_oracle as O

    ocfg = O.CONFIGS[cfg_name]
    cfg = {"v1": E.V1, "v2": E.V2, "v3": E.V3}[cfg_name]
    sd = O.random_state_dict(ocfg, seed=0, loud=True)
    w = O.folded_weights(sd)
    eng = Engine(cfg, 0)
    eng.load_state_dict(sd, strict=True)
    eng.finalize()
    specs = {n: (tr, d0, d1, k) for n, tr, d0, d1, k in cfg.layer_specs()}
    geo = {n: (kind, cin, cout, k, dil) for n, kind, cin, cout, k, dil, _, _ in O.conv_layers(ocfg)}
    rc = 0
    print(f"== mode {mode} cfg {cfg_name} B={B} L={L}", flush=True)
    for name in layer_cases(cfg_name):
        kind, cin, cout, k, dil = geo[name]
        if mode != "fp32" and name not in ("conv_pre", "conv_post") and (cin % 32 or cout % 32):
            continue   # narrower than a tensor-core tile: these stages run on the fp32 family in every mode
        torch.manual_seed(zlib.crc32(name.encode()) % 1000)
        x = torch.randn(B, cin, L)
        pre = name != "conv_pre"
        xin = F.leaky_relu(x, 0.1) if pre else x
        if kind == "conv":
            ref = F.conv1d(xin.double(), w[name + ".weight"].double(), w[name + ".bias"].double(), dilation=dil,
                           padding=O.get_padding(k, dil))
        else:
            i = int(name.split(".")[1])
            u = cfg.upsample_rates[i]
            ref = F.conv_transpose1d(xin.double(), w[name + ".weight"].double(), w[name + ".bias"].double(), stride=u,
                                     padding=(k - u) // 2)
        t0 = time.time()
        try:
            y = eng.run_layer(name, x.numpy(), pre_lrelu=pre, precision=mode)
        except Exception as ex:  # noqa: BLE001
            print(f"  {name:28s} FAILED: {ex}", flush=True)
            return 2
        err = np.abs(y.astype(np.float64) - ref.numpy())
        scale = float(ref.abs().max())
        bad = int((err > 1e-3 * max(scale, 1.0)).sum())
        print(f"  {name:28s} cin {cin:4d} cout {cout:4d} k {k:2d} d {dil:2d} max|err| {err.max():.3e} ref|max| {scale:.3f} "
              f"rel {err.max() / scale:.2e} bad {bad} ({time.time() - t0:.2f}s)", flush=True)
        if bad and mode != "bf16":
            rc = 1
            idx = np.argwhere(err > 1e-3 * max(scale, 1.0))
            print("     first bad idx (b,c,t):", idx[:6].tolist(), " last:", idx[-3:].tolist(), flush=True)
    # end to end on the golden-style case
    mel = O.synthetic_mel(2, 24, seed=1234)
    taps = {}
    ref = O.forward(sd, torch.from_numpy(mel), ocfg, dtype=torch.float64, taps=taps).numpy()[:, 0]
    out = eng.forward(mel, precision=mode, keep_taps=True)
    print(f"  e2e keep_taps   max|err| {np.abs(out - ref).max():.3e} (out std {ref.std():.3f})", flush=True)
    for tname, t in taps.items():
        if tname == "out":
            continue
        try:
            got = eng.get_tap(tname, t.shape)
        except Exception:  # noqa: BLE001
            continue
        e = np.abs(got - t.numpy()).max()
        print(f"     tap {tname:14s} max|err| {e:.3e}  ref|max| {float(t.abs().max()):.3f}", flush=True)
    out2 = eng.forward(mel, precision=mode)
    e2 = np.abs(out2 - ref).max()
    print(f"  e2e fused       max|err| {e2:.3e}  launches so far {eng.launch_count}", flush=True)
    if mode != "bf16" and e2 > 1e-3:
        rc = 1
    mel = O.synthetic_mel(3, 301, seed=5, realistic=True)
    ref = O.forward(sd, torch.from_numpy(mel), ocfg).numpy()[:, 0]
    out3 = eng.forward(mel, precision=mode)
    e3 = np.abs(out3 - ref).max()
    print(f"  e2e B=3 T=301   max|err| {e3:.3e}", flush=True)
    if mode != "bf16" and e3 > 1e-3:
        rc = 1
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", default=None)
    ap.add_argument("--cfg", default="v1")
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--L", type=int, default=333)
    ap.add_argument("--modes", default="fp32,bf16x3,bf16")
    ap.add_argument("--cfgs", default="v1,v2,v3")
    ap.add_argument("--no-apt", action="store_true", help="skip the HFG_UMMA_A_PER_TAP=1 variants")
    a = ap.parse_args()
    if a.one:
        sys.exit(run_one(a.one, a.cfg, a.B, a.L))
    summary = {}
    for cfg in a.cfgs.split(","):
        for mode in a.modes.split(","):
            for env_extra in ({}, {"HFG_UMMA_A_PER_TAP": "1"}):
                if env_extra and (mode == "fp32" or a.no_apt):
                    continue
                env = dict(os.environ, **env_extra)
                tag = f"{cfg}/{mode}" + ("/a_per_tap" if env_extra else "")
                try:
                    r = subprocess.run([sys.executable, __file__, "--one", mode, "--cfg", cfg, "--B", str(a.B), "--L", str(a.L)],
                                       env=env, timeout=300)
                    summary[tag] = r.returncode
                except subprocess.TimeoutExpired:
                    summary[tag] = "timeout"
                print(f"## {tag}: {summary[tag]}", flush=True)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
