"""GPU parity at the BASELINE shapes, in every tensor-core mode (north-star: V1 at batch 32 x 10 s; config 5: V2 / V3-args at
batch 64 x 10 s), plus the properties of this round's engine changes: the MRF sum folded into the producers' epilogues, the
CUDA-graph launch path, and a repeated-run determinism stress (the stand-in for a race detector: compute-sanitizer is not
available on the GPU pool).

Each parity test prints max-abs error and error / output std per checked item; the asserted tolerances are those of
tests/test_gpu_parity.py (fp32-class: 1e-3 absolute; fp16: 0.025 std; bf16: 0.15 std).
"""
import os

import numpy as np
import pytest
import torch

from oracle import hifigan_oracle as O
from test_gpu_parity import PAIR_RTOL, TC_MODES, _cfgs, _engine, _pair_ref, e2e_tol

pytestmark = pytest.mark.gpu


def _check_items(eng, sd, ocfg, mel, mode, items, label, abs_tol=None):
    out = eng.forward(mel, precision=mode)
    assert out.shape == (mel.shape[0], mel.shape[2] * 256) and np.isfinite(out).all()
    for b in items:
        ref = O.infer(sd, mel[b:b + 1], ocfg)[0]
        err = float(np.abs(out[b] - ref).max())
        tol = abs_tol if abs_tol is not None else e2e_tol(mode, ref)
        print(f"[parity] {label} {mode} item {b}: max|err| {err:.3e} = {err / ref.std():.3e} std (tol {tol:.3e}, output std {ref.std():.4f})")
        assert err <= tol, f"{label} {mode} item {b}: {err:.3e} > {tol:.3e}"
    return out


@pytest.mark.parametrize("loud", [True, False])
@pytest.mark.parametrize("mode", TC_MODES)
def test_north_star_shape_b32_x_10s(mode, loud):
    """BASELINE north-star case: V1, 32 utterances x 862 frames, every tensor-core mode, loud and default-init weights;
    three items against the oracle (a full-batch oracle pass would take minutes on the host)."""
    eng, sd = _engine("v1", loud=loud)
    mel = O.synthetic_mel(32, 862, seed=1234)
    out = _check_items(eng, sd, O.V1, mel, mode, (0, 13, 31), f"v1 B=32 T=862 {'loud' if loud else 'default'}",
                       abs_tol=None if loud else 1e-3)        # default init: the headline 1e-3, every mode
    if not loud:
        return
    # batch independence at this size: an item run alone reproduces its bits
    np.testing.assert_array_equal(out[13], eng.forward(mel[13:14], precision=mode)[0])


@pytest.mark.parametrize("cfg_name", ["v2", "v3"])
@pytest.mark.parametrize("mode", TC_MODES)
def test_small_generators_b64_x_10s(cfg_name, mode):
    """BASELINE config 5: V2 (C0 = 128) and V3-args at batch 64 x 862 frames, loud weights, three items against the oracle."""
    eng, sd = _engine(cfg_name, loud=True)
    _cfg, ocfg = _cfgs(cfg_name)
    mel = O.synthetic_mel(64, 862, seed=99)
    _check_items(eng, sd, ocfg, mel, mode, (0, 31, 63), f"{cfg_name} B=64 T=862 loud")


# ---------------------------------------------------------------------------
# MRF sum folded into the last convs2 epilogue of a branch (hifigan_pretrained.py:133-137)
# ---------------------------------------------------------------------------

@pytest.mark.parametrize("mode", TC_MODES)
def test_last_resblock_step_with_folded_mrf_sum(mode):
    """y = (x + c2(lrelu(c1(lrelu x))) + s) * scale for the fused pair kernel (C <= 64) and the two-launch plan (C >= 128, and
    HFG_PAIR=0): against float64, at ragged lengths and tile edges; where the pair fuses, its bits equal the two-launch plan's."""
    eng, sd = _engine("v1")
    w = O.folded_weights(sd)
    for (n, m), lengths in (((10, 2), (1, 118, 777)), ((9, 2), (255, 509)), ((7, 2), (123, 700)), ((6, 2), (127, 600)), ((4, 2), (300,)),
                            ((2, 2), (130,))):
        C = 512 >> (n // 3 + 1)
        k, d = (3, 7, 11)[n % 3], (1, 3, 5)[m]
        for L in lengths:
            torch.manual_seed(7 * L + n)
            x = torch.randn(2, C, L)
            s_prev = torch.randn(2, C, L) * 1.5
            for scale in (1.0, 1.0 / 3.0):
                ref = (_pair_ref(w, n, m, k, d, x) + s_prev.double().numpy()) * scale
                y, fused = eng.run_pair(n, m, x.numpy(), precision=mode, mrf_sum=s_prev.numpy(), out_scale=scale)
                tol = PAIR_RTOL[mode] * max(1.0, np.abs(ref).max())
                assert np.abs(y - ref).max() <= tol, (n, m, L, scale, fused, float(np.abs(y - ref).max()))
                if fused:
                    os.environ["HFG_PAIR"] = "0"
                    try:
                        y2, fused2 = eng.run_pair(n, m, x.numpy(), precision=mode, mrf_sum=s_prev.numpy(), out_scale=scale)
                    finally:
                        del os.environ["HFG_PAIR"]
                    assert not fused2
                    np.testing.assert_array_equal(y, y2, err_msg=f"resblocks.{n} pair {m} L={L} scale={scale}")


@pytest.mark.parametrize("mode", TC_MODES)
def test_mrf_fold_equals_separate_combine_pass(mode):
    """HFG_MRF_FOLD=0 restores the separate mrf_combine pass: both plans are the reference's (r0 + r1 + r2) / 3 up to the rounding
    of the partial sums to the operand planes."""
    from iris_tts_b200 import Engine
    from iris_tts_b200.engine import V1
    sd = O.random_state_dict(O.V1, seed=0, loud=True)
    mel = O.synthetic_mel(2, 70, seed=5)
    ref = O.infer(sd, mel)
    outs = []
    for fold in ("1", "0"):
        os.environ["HFG_MRF_FOLD"] = fold
        try:
            eng = Engine(V1, 0)
            eng.load_state_dict(sd, strict=True)
            eng.finalize()
            n0 = eng.launch_count
            outs.append(eng.forward(mel, precision=mode))
            launches = eng.launch_count - n0
            eng.close()
        finally:
            del os.environ["HFG_MRF_FOLD"]
        assert np.abs(outs[-1] - ref).max() <= e2e_tol(mode, ref)
        print(f"[mrf] fold={fold} {mode}: {launches} launches, max|err| {np.abs(outs[-1] - ref).max():.3e}")
    assert np.abs(outs[0] - outs[1]).max() <= {"bf16x3": 5e-5, "fp16": 4e-3, "bf16": 3e-2}[mode]


# ---------------------------------------------------------------------------
# launch paths and determinism
# ---------------------------------------------------------------------------

@pytest.mark.parametrize("mode", ["fp32"] + TC_MODES)
def test_graph_launch_equals_direct_launch(mode):
    """The first forward of a (B, T, precision) plan launches kernel by kernel; from the second on the plan runs as one CUDA
    graph (programmatic-dependent-launch edges preserved).  Same kernels, same arguments: identical bits, repeatedly."""
    from iris_tts_b200 import Engine
    from iris_tts_b200.engine import V1
    sd = O.random_state_dict(O.V1, seed=0, loud=True)
    eng = Engine(V1, 0)
    eng.load_state_dict(sd, strict=True)
    eng.finalize()
    mel = O.synthetic_mel(3, 45, seed=17)
    first = eng.forward(mel, precision=mode)           # direct
    n0 = eng.launch_count
    for _ in range(4):                                 # graph (captured on the second call)
        np.testing.assert_array_equal(eng.forward(mel, precision=mode), first)
    assert n0 > 0 and eng.launch_count == 5 * n0      # the launch counter keeps counting kernels under graph launch
    assert eng.graph_stats == (1, 0)                   # B = 3 x 45 frames: captured WITH the concurrent branch lanes of its stages
    other = O.synthetic_mel(3, 45, seed=18)            # a graph replays the plan, not the data
    ref = O.infer(sd, other)
    assert np.abs(eng.forward(other, precision=mode) - ref).max() <= e2e_tol(mode, ref)
    eng.close()


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_pipelined_host_path_equals_the_synchronous_one(mode):
    """Engine.forward on a large enough batch runs two half-batches back to back without waiting (HFG_NO_SYNC with page-locked
    host pointers): the second half is staged on the host while the first computes, and the first half's waveform is copied out
    on the copy stream while the second computes.  Same bits as the synchronous whole-batch call, call after call."""
    eng, sd = _engine("v1")
    os.environ["HFG_PIPELINE"] = "1"          # off by default (measured: no net gain at the BASELINE shape)
    try:
        for B, T in ((16, 300), (5, 1000), (4, 1001)):
            mel = O.synthetic_mel(B, T, seed=B + T)
            want = eng.forward(mel, precision=mode, pinned=False)          # pageable buffers: one synchronous whole-batch forward
            for _ in range(3):
                got = eng.forward(mel, precision=mode)                       # pipelined halves
                np.testing.assert_array_equal(got, want)
            other = eng.forward(O.synthetic_mel(B, T, seed=1), precision=mode)   # staging buffers are reused across calls, results are not
            assert not np.array_equal(other, want)
            np.testing.assert_array_equal(eng.forward(mel, precision=mode), want)
    finally:
        del os.environ["HFG_PIPELINE"]
    ref = O.infer(sd, mel[1:2])[0]
    assert np.abs(want[1] - ref).max() <= e2e_tol(mode, ref)


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_concurrent_branch_lanes_equal_the_serial_plan(mode):
    """Small inputs run the three ResBlock branches of a stage on three streams (parallel branches of the plan's graph).  Same
    kernels on the same data in per-branch buffers: identical bits to the serial plan (HFG_BRANCH_PAR=0), also when forced on for
    a batch that fills the GPU, also kernel by kernel without a graph."""
    from iris_tts_b200 import Engine
    from iris_tts_b200.engine import V1
    sd = O.random_state_dict(O.V1, seed=0, loud=True)
    outs = {}
    for par in ("0", "1"):
        os.environ["HFG_BRANCH_PAR"] = par
        try:
            eng = Engine(V1, 0)
            eng.load_state_dict(sd, strict=True)
            eng.finalize()
            for B, T in ((1, 862), (2, 97), (8, 300)):
                mel = O.synthetic_mel(B, T, seed=B * 7 + T)
                first = eng.forward(mel, precision=mode)                 # direct launches (lanes as plain streams)
                for _ in range(3):
                    np.testing.assert_array_equal(eng.forward(mel, precision=mode), first)   # graph
                outs.setdefault((B, T), []).append(first)
            assert eng.graph_stats[1] == 0
            eng.close()
        finally:
            del os.environ["HFG_BRANCH_PAR"]
    for key, (serial, parallel) in outs.items():
        np.testing.assert_array_equal(serial, parallel, err_msg=str(key))
    mel = O.synthetic_mel(1, 862, seed=1 * 7 + 862)
    ref = O.infer(sd, mel)
    assert np.abs(outs[(1, 862)][1] - ref).max() <= e2e_tol(mode, ref)


def test_caller_device_is_left_alone():
    """Every ABI call runs on the engine's device and restores the caller's current device."""
    if torch.cuda.device_count() < 2:
        eng, _ = _engine("v2")
        torch.cuda.set_device(0)
        eng.forward(O.synthetic_mel(1, 4, seed=1), precision="bf16")
        assert torch.cuda.current_device() == 0
        return
    from iris_tts_b200 import Engine
    from iris_tts_b200.engine import V2
    sd = O.random_state_dict(O.V2, seed=0, loud=True)
    torch.cuda.set_device(0)
    eng = Engine(V2, 1)
    eng.load_state_dict(sd, strict=True)
    eng.finalize()
    mel = O.synthetic_mel(1, 9, seed=2)
    out = eng.forward(mel, precision="bf16x3")
    assert torch.cuda.current_device() == 0
    assert np.abs(out - O.infer(sd, mel, O.V2)).max() <= 1e-3
    eng.close()
    assert torch.cuda.current_device() == 0


# ring depths that every V1 layer can be planned with; the arithmetic order does not depend on them
_OVERRIDES = [
    {},
    {"HFG_U2_NA": "2", "HFG_U2_NE": "3"},
    {"HFG_U2_NW": "3", "HFG_PAIR_NX": "3"},
    {"HFG_U2_NA": "3", "HFG_U2_NW": "6", "HFG_U2_NE": "4", "HFG_PAIR_NO": "1"},
    {"HFG_SNAKE": "0", "HFG_PDL": "1"},
]


@pytest.mark.parametrize("mode", TC_MODES)
def test_repeated_runs_are_bitwise_identical_under_varied_pipelines(mode):
    """Race evidence without a sanitizer: 200 forwards of BASELINE config 2 (16 x 862 frames) per mode, in five engines whose
    TMA-ring depths / staging slots / tile walk order are forced to different values (HFG_U2_* / HFG_PAIR_* overrides).  A missing
    barrier or a slot reused too early shows up as a run-to-run difference; every output must equal the first one bit for bit."""
    from iris_tts_b200 import Engine
    from iris_tts_b200.engine import V1
    sd = O.random_state_dict(O.V1, seed=0, loud=True)
    mel = torch.from_numpy(O.synthetic_mel(16, 862, seed=1234)).cuda()
    out = torch.empty(16, 862 * 256, dtype=torch.float32, device="cuda")
    first = None
    runs = 0
    for ov in _OVERRIDES:
        os.environ.update(ov)
        try:
            eng = Engine(V1, 0)
            eng.load_state_dict(sd, strict=True)
            eng.finalize()
            mine = None                                    # first output of THIS engine
            for _ in range(40):
                out.zero_()
                torch.cuda.synchronize()
                eng.forward_ptr(mel.data_ptr(), 16, 862, out.data_ptr(), mode, mel_on_device=True, wave_on_device=True)
                if mine is None:
                    mine = out.clone()
                    if first is None:
                        first = mine
                        ref = O.infer(sd, mel[3:4].cpu().numpy())[0]
                        assert np.abs(first[3].cpu().numpy() - ref).max() <= e2e_tol(mode, ref)
                    elif not torch.equal(mine, first):
                        # A forced depth that a layer cannot be planned with sends that layer to the first-generation kernel, whose
                        # K-chunk order differs in bf16x3 (64- vs 32-channel chunks): the same sums in another fp32 order (after ~80
                        # layers the waveforms differ at the mode's own error level, measured 3e-5), not a race.
                        d = float((mine - first).abs().max())
                        print(f"[stress] {mode} overrides {ov}: plan differs from the default one, max|diff| {d:.2e}")
                        assert mode == "bf16x3" and d <= 1e-4, (mode, ov, d)
                else:
                    assert torch.equal(out, mine), f"run {runs} differs from the first run of its engine (overrides {ov})"
                runs += 1
            eng.close()
        finally:
            for k in ov:
                del os.environ[k]
    assert runs == 200


# ---------------------------------------------------------------------------
# generator architectures nobody hand-picked
# ---------------------------------------------------------------------------

def _random_config(rng):
    from iris_tts_b200.engine import GeneratorConfig
    nu = int(rng.integers(2, 5))
    rates = [int(rng.choice([2, 4, 8])) for _ in range(nu)]
    while int(np.prod(rates)) > 512:
        rates[int(np.argmax(rates))] //= 2
    c_last = int(rng.choice([8, 16, 32, 64]))
    c0 = c_last << nu
    if c0 > 512:
        c0, c_last = 512, 512 >> nu
    nk = int(rng.integers(1, 4))
    ks = [int(rng.choice([3, 5, 7, 9, 11])) for _ in range(nk)]
    nd = int(rng.integers(1, 4))
    dils = tuple(tuple(int(rng.integers(1, 7)) for _ in range(nd)) for _ in range(nk))
    return GeneratorConfig(80, tuple(rates), tuple(2 * r for r in rates), c0, tuple(ks), dils)


@pytest.mark.parametrize("seed", list(range(10)))
def test_random_architectures_match_the_oracle(seed):
    """Ten seeded random generator configurations (2-4 upsamplers of rate 2/4/8, 1-3 ResBlock kernels of size 3-11, 1-3 dilations
    of 1-6, 8-64 final channels): every arithmetic mode against the oracle on loud weights, at a ragged length.  Exercises planner
    paths no named configuration reaches (single-branch MRF, one-dilation blocks, time-folded stages behind every rate mix)."""
    from iris_tts_b200 import Engine
    rng = np.random.default_rng(1000 + seed)
    cfg = _random_config(rng)
    ocfg = O.OracleConfig(cfg.in_channels, cfg.upsample_rates, cfg.upsample_kernel_sizes, cfg.upsample_initial_channel,
                          cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)
    sd = O.random_state_dict(ocfg, seed=seed, loud=True)
    eng = Engine(cfg, 0)
    eng.load_state_dict(sd, strict=True)
    eng.finalize()
    for B, T in ((2, 37), (1, 130)):
        mel = O.synthetic_mel(B, T, seed=seed * 10 + T)
        ref = O.infer(sd, mel, ocfg)
        for mode in ("fp32", "bf16x3", "fp16", "bf16"):
            for _ in range(2):                      # direct launches, then the graph
                out = eng.forward(mel, precision=mode)
            assert out.shape == ref.shape
            err = float(np.abs(out - ref).max())
            assert err <= e2e_tol(mode, ref), f"{cfg} B={B} T={T} {mode}: {err:.3e} > {e2e_tol(mode, ref):.3e} (std {ref.std():.3f})"
    eng.close()


ODD_CONFIGS = {
    "128 mel channels": (128, (8, 8, 2, 2), (16, 16, 4, 4), 512, (3, 7, 11), ((1, 3, 5),) * 3),
    "rates 5,4,4,2,2 with kernels 11,8,8,4,4": (80, (5, 4, 4, 2, 2), (11, 8, 8, 4, 4), 512, (3, 7, 11), ((1, 3, 5),) * 3),
    "upsample kernel = rate": (80, (4, 4), (4, 4), 128, (3, 5), ((1, 2), (1, 2))),
    "upsample kernel = 3 x rate": (80, (4, 4), (12, 12), 128, (3, 5), ((1, 2), (1, 2))),
    "rate 3 with kernel 7": (80, (3, 2), (7, 4), 128, (3,), ((1, 3),)),
    "13-tap ResBlock with dilation 7": (80, (8, 4), (16, 8), 128, (13,), ((1, 7),)),
    "one upsampler": (80, (8,), (16,), 64, (3, 7), ((1, 3), (1, 3))),
    "1024 initial channels (C = 512 ResBlocks, 1024 -> 512 upsampler)": (80, (8, 8, 2, 2), (16, 16, 4, 4), 1024, (3, 7, 11), ((1, 3, 5),) * 3),
    "2048 initial channels": (80, (4, 4, 2, 2), (8, 8, 4, 4), 2048, (3,), ((1, 3),)),
}


@pytest.mark.parametrize("name", list(ODD_CONFIGS))
def test_constructor_arguments_outside_the_named_generators(name):
    """The reference builds whatever its constructor is given (hifigan_pretrained.py:77-121).  Configurations no named generator
    uses -- odd upsampling rates, kernels that are not twice the rate, other mel widths, very wide first stages (which leave the
    persistent kernel's limits and run on the first-generation one) -- against the oracle in every mode, loud weights.
    (Found by this test: the wide polyphase upsampler read its bias tile out of bounds for C_out > 256.)"""
    from iris_tts_b200 import Engine
    from iris_tts_b200.engine import GeneratorConfig
    cfg = GeneratorConfig(*ODD_CONFIGS[name])
    ocfg = O.OracleConfig(cfg.in_channels, cfg.upsample_rates, cfg.upsample_kernel_sizes, cfg.upsample_initial_channel,
                          cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)
    sd = O.random_state_dict(ocfg, seed=1, loud=True)
    eng = Engine(cfg, 0)
    eng.load_state_dict(sd, strict=True)
    eng.finalize()
    mel = np.random.default_rng(3).standard_normal((2, cfg.in_channels, 45)).astype(np.float32)
    ref = O.infer(sd, mel, ocfg)
    for mode in ("fp32", "bf16x3", "fp16", "bf16"):
        for _ in range(2):
            out = eng.forward(mel, precision=mode)
        assert out.shape == ref.shape
        err = float(np.abs(out - ref).max())
        assert err <= e2e_tol(mode, ref), f"{name} {mode}: {err:.3e} > {e2e_tol(mode, ref):.3e}"
    eng.close()


def test_unsupported_constructor_arguments_are_refused_with_a_reason():
    """What the engine does not take, it says so at construction (never a wrong waveform): mel widths that are not a multiple of 8
    (TMA rows are 16-byte multiples) final channel counts that are not a power of two in [8, 128], and
    upsamplers whose kernel minus rate is odd (the reference's output would be one sample longer per stage than T * hop)."""
    from iris_tts_b200 import Engine, _abi
    from iris_tts_b200.engine import GeneratorConfig
    for args, word in (((100, (8, 8, 2, 2), (16, 16, 4, 4), 512, (3, 7, 11), ((1, 3, 5),) * 3), "multiple of 8"),
                       ((80, (8, 8, 2, 2), (16, 16, 4, 4), 384, (3, 7, 11), ((1, 3, 5),) * 3), "power of two"),
                       ((80, (3, 2), (6, 4), 128, (3,), ((1,),)), "even difference")):
        with pytest.raises(_abi.HfgError, match=word):
            Engine(GeneratorConfig(*args), 0)
