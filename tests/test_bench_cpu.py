"""bench.py contract on the CPU: the reference arm (the oracle timed on host cores) prints ONE JSON line with the keys the
driver reads; under a multi-rank launch only rank 0 prints.  Our own arm needs a GPU and must say so instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "16"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "vocoder_audio_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    # baseline/_ref (the unmodified reference module, baseline/install_ref.py) is timed when present, the oracle port otherwise
    want_kind = "reference" if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "iris", "hifigan_pretrained.py")) else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["same_config"] is True and d["config"]["batch_timed"] == 16
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "16", "--gpus", "2"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
