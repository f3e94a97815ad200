"""The drop-in entry points on a GPU, against what the REFERENCE module returned for the same calls
(tests/golden/api_shapes.npz, made by tests/golden/make_golden.py) and against the oracle."""
import importlib.util
import os
import wave

import numpy as np
import pytest
import torch

from oracle import hifigan_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def loud_ckpt(tmp_path_factory):
    p = tmp_path_factory.mktemp("ckpt") / "generator.ckpt"
    torch.save(O.random_state_dict(O.V1, seed=0, loud=True), p)
    return p


def test_infer_hifigan_shape_rules_and_values_match_the_reference(loud_ckpt):
    import iris.hifigan_pretrained as hp
    z = np.load(os.path.join(GOLD, "api_shapes.npz"))
    mel = z["mel"]                                            # [1, 80, 12]
    a3 = hp.infer_hifigan(mel, checkpoint_path=loud_ckpt)     # [1,80,T] -> [N]   (reference :313-315)
    a2 = hp.infer_hifigan(mel[0], checkpoint_path=loud_ckpt)  # [80,T]   -> [N]
    gen = hp.get_pretrained_hifigan(loud_ckpt)
    g3 = gen(mel)                                             # [1,N]  (HiFiGANGenerator.__call__ keeps the batch dim)
    g64 = gen(mel.astype(np.float64))                         # any float dtype in, float32 out (:228)
    ignored = hp.infer_hifigan(mel, 16000, 123, loud_ckpt)    # sample_rate / hop_length accepted and ignored (:299-300)
    for got, want in ((a3, z["infer3"]), (a2, z["infer2"]), (g3, z["call3"]), (g64, z["call64"]), (ignored, z["infer3"])):
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.abs(got - want).max() <= 1e-3
    np.testing.assert_array_equal(a3, a2)


def test_singleton_semantics(loud_ckpt, tmp_path):
    import iris.hifigan_pretrained as hp
    a = hp.get_pretrained_hifigan(loud_ckpt)
    assert hp.get_pretrained_hifigan(loud_ckpt) is a                       # cached on the resolved path (:267-283)
    assert hp.get_pretrained_hifigan(loud_ckpt, force_reload=True) is not a
    other = tmp_path / "other.ckpt"
    torch.save({"generator": O.random_state_dict(O.V1, seed=5)}, other)    # nested-dict checkpoint format (:172-182)
    b = hp.get_pretrained_hifigan(other)
    assert b.checkpoint_path == other and b.device.type == "cuda"
    mel = O.synthetic_mel(1, 9, seed=2)
    assert np.abs(b(mel) - O.infer(O.random_state_dict(O.V1, seed=5), mel)).max() <= 1e-3


def test_model_object_on_cuda_and_cpu_tensors():
    import iris.hifigan_pretrained as hp
    torch.manual_seed(0)
    m = hp.HiFiGANModel().eval().to("cuda:0")
    sd = m.state_dict()
    mel = torch.from_numpy(O.synthetic_mel(2, 11, seed=4))
    ref = O.forward(sd, mel)
    out_cpu = m(mel)                                         # [B, 1, T*256] like the reference module (:123-143)
    out_gpu = m(mel.cuda())
    assert out_cpu.shape == ref.shape == (2, 1, 11 * 256) and not out_cpu.is_cuda
    assert out_gpu.is_cuda and out_gpu.shape == ref.shape
    assert float((out_cpu - ref).abs().max()) <= 1e-3
    assert torch.equal(out_gpu.cpu(), out_cpu)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 79, 4))


def test_streaming_output_equals_the_full_forward_on_the_engine():
    """sharding.synthesize_streaming on the CUDA engine: 100-frame pieces (+ 16-frame halo) concatenate to the whole waveform."""
    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import sharding
    torch.manual_seed(0)
    m = hp.HiFiGANModel().eval().to("cuda:0")
    mel = torch.from_numpy(O.synthetic_mel(1, 431, seed=9, realistic=True))
    full = m(mel).reshape(-1)
    pieces = list(sharding.synthesize_streaming(m, mel, 100))
    assert [p.numel() for p in pieces] == [25600, 25600, 25600, 25600, 31 * 256]
    assert float((torch.cat(pieces) - full).abs().max()) <= 2e-5


def test_keras_surface_matches_the_oracle_with_permuted_weights(tmp_path):
    """create_vocoder().infer: Keras layouts (Conv1D [k, ci, co], Conv1DTranspose [k, co, ci]); restated mapping."""
    import iris.vocoder as kv
    voc = kv.create_vocoder()
    sd = O.random_state_dict(O.V1, seed=0, loud=True)
    w = O.folded_weights(sd)
    arrays = []
    for key in voc.model.weights:                            # "<layer>/kernel", "<layer>/bias"
        name, kind = key.split("/")
        if kind == "bias":
            arrays.append(w[name + ".bias"].numpy())
        elif name.startswith("ups."):
            arrays.append(O.keras_convT_kernel(w[name + ".weight"].numpy()))
        else:
            arrays.append(O.keras_conv_kernel(w[name + ".weight"].numpy()))
    voc.model.set_weights(arrays)
    mel = O.synthetic_mel(2, 14, seed=6)
    ref = O.infer(sd, mel)
    a = voc.infer(mel)                                       # [B, 80, T] -> [B, N]
    b = voc(mel[0])                                          # [80, T]    -> [N]
    c = voc.model(np.transpose(mel, (0, 2, 1)))              # channels-last model call -> [B, N, 1]
    assert a.shape == ref.shape and b.shape == ref[0].shape and c.shape == ref.shape + (1,)
    assert np.abs(a - ref).max() <= 1e-3 and np.abs(b - ref[0]).max() <= 1e-3
    np.testing.assert_array_equal(c[..., 0], a)
    p = str(tmp_path / "w.npz")
    voc.save_weights(p)
    voc2 = kv.create_vocoder(weights_path=p)
    np.testing.assert_array_equal(voc2.infer(mel), a)
    kv.create_vocoder(weights_path=str(tmp_path / "missing.weights.h5"))   # missing file: logs and stays random (:161-165)


def test_variable_length_batches_are_exact(loud_ckpt):
    import iris.hifigan_pretrained as hp
    from iris_tts_b200.batching import synthesize_variable
    gen = hp.get_pretrained_hifigan(loud_ckpt)
    mels = [O.synthetic_mel(1, t, seed=10 + i)[0] for i, t in enumerate((7, 12, 7, 3, 12, 0))]
    outs = synthesize_variable(gen, mels)
    assert [o.shape[0] for o in outs] == [t * 256 for t in (7, 12, 7, 3, 12, 0)]
    for m, o in zip(mels, outs):
        if m.shape[1]:
            np.testing.assert_array_equal(o, gen(m))        # identical to running the utterance alone


@pytest.mark.parametrize("mode", ["bf16x3", "fp16"])
def test_ragged_batch_by_buckets_and_tails_equals_per_utterance_forwards(loud_ckpt, mode, monkeypatch):
    """f4: 12 utterances of 12 distinct lengths (33 .. 400 frames) in a handful of dense calls -- a zero-padded body pass per
    length bucket plus ONE tail pass over every utterance's last 32 frames -- bit-identical to twelve batch-1 forwards.  This is
    the scheme for plain callables and the fp32 mode (HFG_RAGGED=0 forces it here); the engine's native ragged plan is
    tests/test_gpu_ragged.py."""
    import iris.hifigan_pretrained as hp
    monkeypatch.setenv("HFG_RAGGED", "0")
    from iris_tts_b200 import sharding
    from iris_tts_b200.batching import synthesize_variable
    gen = hp.get_pretrained_hifigan(loud_ckpt)
    old = gen.model.precision
    gen.model.precision = mode
    try:
        assert sharding.halo_frames(gen.model.config) <= sharding.HALO_FRAMES
        lengths = (400, 33, 371, 64, 390, 127, 350, 32, 398, 129, 65, 301)
        mels = [O.synthetic_mel(1, t, seed=50 + i, realistic=(i % 2 == 0))[0] for i, t in enumerate(lengths)]
        stats = {}
        outs = synthesize_variable(gen, mels, stats=stats)
        assert stats["distinct_lengths"] == 12 and stats["calls"] <= 7 and not stats["native_ragged"], stats
        for m, o in zip(mels, outs):
            np.testing.assert_array_equal(o, gen(m))
        ref = O.infer(O.random_state_dict(O.V1, seed=0, loud=True), mels[5])
        assert np.abs(outs[5] - ref).max() <= (1e-3 if mode == "bf16x3" else 0.025 * ref.std())
    finally:
        gen.model.precision = old


def test_cli_end_to_end(loud_ckpt, tmp_path):
    spec = importlib.util.spec_from_file_location("synthesize_cli", os.path.join(ROOT, "scripts", "synthesize.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    mel = O.synthetic_mel(1, 20, seed=1)[0]
    np.save(tmp_path / "mel.npy", mel)
    out = tmp_path / "o.wav"
    assert cli.main(["--mel", str(tmp_path / "mel.npy"), "--output_wav", str(out), "--checkpoint", str(loud_ckpt)]) == 0
    with wave.open(str(out)) as w:
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32767.0
    ref = O.infer(O.random_state_dict(O.V1, seed=0, loud=True), mel)
    assert pcm.shape == ref.shape and np.abs(pcm - ref).max() <= 1e-3 + 1.0 / 32767.0


def test_pickled_reference_model_checkpoint_runs_on_the_engine():
    """A checkpoint that holds a whole model object pickled by the REFERENCE (hifigan_pretrained.py:168-171), non-default
    architecture: loaded through HiFiGANGenerator and run on the GPU, against the output the reference itself produced."""
    import iris.hifigan_pretrained as hp
    gen = hp.HiFiGANGenerator(os.path.join(GOLD, "pickled_model.ckpt"))
    z = np.load(os.path.join(GOLD, "pickled_model_expected.npz"))
    out = gen(z["mel"])
    want = z["out"][:, 0]
    assert out.shape == want.shape and out.dtype == np.float32
    assert np.abs(out - want).max() <= 1e-3


def test_keras_weights_h5_and_keras_archive_through_the_keras_surface(tmp_path):
    """create_vocoder(weights_path='x.weights.h5') (demo_vocoder.py:83; reference vocoder.py:161-170): the HDF5 file is read by
    the pure-Python reader, the kernels are permuted to the torch layout, and the waveform equals the oracle's on those weights."""
    import zipfile

    import _h5write
    import iris.vocoder as kv
    from test_h5lite_cpu import _keras_tree
    src = kv.HiFiGANGenerator(seed=11)
    rng = np.random.default_rng(0)
    for k in src.weights:
        if k.endswith("/bias"):
            src.weights[k] = (rng.standard_normal(src.weights[k].shape) * 0.05).astype(np.float32)
    p = tmp_path / "gen.weights.h5"
    _h5write.write_h5(p, _keras_tree(src, "layers"))
    voc = kv.create_vocoder(weights_path=str(p))
    sd = {}
    for name, *_ in src.config.layer_specs():
        sd[f"{name}.weight"] = torch.from_numpy(np.ascontiguousarray(np.transpose(src.weights[f"{name}/kernel"], (2, 1, 0))))
        sd[f"{name}.bias"] = torch.from_numpy(src.weights[f"{name}/bias"])
    mel = O.synthetic_mel(2, 21, seed=9)
    ref = O.infer(sd, mel)
    out = voc.infer(mel)
    assert out.shape == ref.shape and np.abs(out - ref).max() <= 1e-3
    z = tmp_path / "gen.keras"
    with zipfile.ZipFile(z, "w") as zf:
        zf.write(p, "model.weights.h5")
    np.testing.assert_array_equal(kv.create_vocoder(weights_path=str(z)).infer(mel), out)


def test_fp16_mode_saturates_instead_of_overflowing():
    """HFG_PREC_FP16 stores activations with cvt.rn.satfinite: inputs that drive activations past 65504 give a finite waveform
    (clipped arithmetic), never NaN; the bf16 mode, with fp32's exponent range, handles the same input without clipping."""
    from iris_tts_b200 import Engine
    from iris_tts_b200.engine import V2
    sd = O.random_state_dict(O.V2, seed=0, loud=True)
    eng = Engine(V2, 0)
    eng.load_state_dict(sd, strict=True)
    eng.finalize()
    mel = O.synthetic_mel(1, 12, seed=3) * 3.0e5
    big = eng.forward(mel, precision="fp16")
    assert np.isfinite(big).all() and np.abs(big).max() <= 1.0
    assert np.isfinite(eng.forward(mel, precision="bf16")).all()
    # and at ordinary amplitudes the saturating conversion never engages: the mode stays inside its tolerance
    small = O.synthetic_mel(1, 12, seed=3)
    ref = O.infer(sd, small, O.V2)
    assert np.abs(eng.forward(small, precision="fp16") - ref).max() <= 0.025 * ref.std()
    eng.close()


def test_cli_copy_synthesis_from_a_wav(loud_ckpt, tmp_path):
    """--audio_wav: waveform -> log-mel on the GPU (src/iris/data.py:25-67) -> vocoder -> wav of the frame-aligned length
    (demo_vocoder.py's copy-synthesis)."""
    spec = importlib.util.spec_from_file_location("synthesize_cli", os.path.join(ROOT, "scripts", "synthesize.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    n = 22050
    t = np.arange(n) / 22050.0
    pcm = (0.3 * np.sin(2 * np.pi * 440.0 * t) * 32767).astype("<i2")
    src = tmp_path / "in.wav"
    with wave.open(str(src), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(22050); w.writeframes(pcm.tobytes())
    out = tmp_path / "copy.wav"
    assert cli.main(["--audio_wav", str(src), "--output_wav", str(out), "--checkpoint", str(loud_ckpt)]) == 0
    with wave.open(str(out)) as w:
        assert w.getframerate() == 22050 and w.getnframes() == (1 + n // 256) * 256


def test_oversized_calls_are_refused_not_wrapped():
    """32-bit indexing inside the kernels: a call whose per-item plane would exceed 2^31 elements (V1: T >= 262144 frames), or a
    batch beyond the grid axis, fails with HFG_ERR_UNSUPPORTED before anything is allocated or launched."""
    import ctypes
    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import _abi
    m = hp.HiFiGANModel()
    m.to("cuda:0")
    eng = m.engine
    lib = _abi.load()
    assert lib.hfg_workspace_bytes(eng._h, 1, 262144, _abi.PREC_BF16) == 0
    assert lib.hfg_workspace_bytes(eng._h, 1, 1000, _abi.PREC_BF16) > 0
    dummy = (ctypes.c_float * 4)()
    rc = lib.hfg_forward(eng._h, ctypes.cast(dummy, ctypes.c_void_p), 70000, 4, ctypes.cast(dummy, ctypes.c_void_p), _abi.PREC_BF16, 0)
    assert rc == _abi.ERR_UNSUPPORTED, rc
    rc = lib.hfg_forward(eng._h, ctypes.cast(dummy, ctypes.c_void_p), 1, 300000, ctypes.cast(dummy, ctypes.c_void_p), _abi.PREC_BF16, 0)
    assert rc == _abi.ERR_UNSUPPORTED, rc
    assert b"chunks" in lib.hfg_last_error()


def test_two_engines_driven_from_two_threads_concurrently():
    """A handle is not thread-safe, but two handles are independent: each owns its stream, arena, plans and graphs, and the
    library's only process-wide state is the thread-local error string and idempotent one-time kernel attributes.  Two
    threads (ctypes drops the GIL during the calls) run 30 forwards each, different shapes and modes, while the other is
    running; every result equals the one the same engine gave when it ran alone."""
    import threading
    import iris.hifigan_pretrained as hp
    rng = np.random.default_rng(5)
    jobs = []
    for seed, (B, T, mode) in enumerate([(3, 211, "bf16x3"), (2, 333, "fp16")]):
        torch.manual_seed(seed)
        m = hp.HiFiGANModel()
        m.to("cuda:0")
        mel = rng.standard_normal((B, 80, T)).astype(np.float32)
        alone = m.engine.forward(mel, precision=mode)
        jobs.append((m, mel, mode, alone))
    errors = []
    start = threading.Barrier(2)

    def run(m, mel, mode, alone):
        try:
            start.wait()
            for _ in range(30):
                got = m.engine.forward(mel, precision=mode)
                if not np.array_equal(got, alone):
                    errors.append(f"{mode}: differs from the solo run by {np.abs(got - alone).max():.3e}")
                    return
        except Exception as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=run, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_one_vocoder_called_from_several_threads():
    """The reference's singleton vocoder is a torch module: concurrent callers (a threaded server) are fine there.  Here a handle
    serves one call at a time, so the Python layer queues concurrent callers of ONE engine (engine.py:_locked); every thread gets
    the result it would have got alone, for different shapes at once (plan cache, arena growth and staging buffers are shared)."""
    import threading
    import iris.hifigan_pretrained as hp
    torch.manual_seed(3)
    m = hp.HiFiGANModel()
    m.to("cuda:0")
    m.precision = "bf16x3"
    rng = np.random.default_rng(9)
    mels = [rng.standard_normal((b, 80, t)).astype(np.float32) for b, t in ((1, 97), (3, 160), (2, 233), (1, 311))]
    alone = [m.engine.forward(x, precision="bf16x3") for x in mels]
    from iris_tts_b200.mel import compute_mel_spectrogram
    audio = (rng.standard_normal(22050) * 0.1).astype(np.float32)
    mel_alone = compute_mel_spectrogram(audio)
    errors = []
    start = threading.Barrier(len(mels))

    def run(i):
        try:
            start.wait()
            for _ in range(15):
                if not np.array_equal(m.engine.forward(mels[i], precision="bf16x3"), alone[i]):
                    errors.append(f"thread {i}: vocoder output differs")
                    return
                if not np.array_equal(compute_mel_spectrogram(audio), mel_alone):
                    errors.append(f"thread {i}: log-mel differs")
                    return
        except Exception as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(mels))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_first_cuda_call_of_a_thread_may_be_a_new_shape():
    """A worker thread whose very first CUDA-touching call is a forward of a shape the engine has not planned yet: the ABI entry
    binds the device's context to the thread itself (the planner calls driver entry points, which fail with
    CUDA_ERROR_INVALID_CONTEXT on a thread that has none).  Pageable numpy buffers: nothing before the call touches CUDA."""
    import threading
    eng_sd = O.random_state_dict(O.V1, seed=0, loud=True)
    from iris_tts_b200 import Engine
    from iris_tts_b200.engine import V1
    eng = Engine(V1, 0)
    eng.load_state_dict(eng_sd, strict=True)
    eng.finalize()
    mel = O.synthetic_mel(1, 23, seed=8)
    got = {}

    def run():
        try:
            got["out"] = eng.forward(mel, precision="bf16x3", pinned=False)
        except Exception as exc:  # noqa: BLE001
            got["err"] = repr(exc)

    t = threading.Thread(target=run)
    t.start()
    t.join()
    assert "err" not in got, got
    assert np.abs(got["out"] - O.infer(eng_sd, mel)).max() <= 1e-3
    eng.close()


def test_ragged_batching_reads_the_receptive_field_from_the_generator():
    """synthesize_variable's body / tail split is exact only if its halo covers the generator's receptive field.  For a generator
    that sees much further than V1 (15-tap ResBlocks with dilation 9 in the first stage: 34 frames instead of 15) the halo and the
    hop are taken from the vocoder's own configuration; the result equals the per-utterance forwards bit for bit -- and the V1
    default of 16 frames, forced, does not."""
    from iris_tts_b200 import Engine
    from iris_tts_b200.batching import synthesize_variable
    from iris_tts_b200.engine import GeneratorConfig
    from iris_tts_b200.sharding import halo_frames
    cfg = GeneratorConfig(80, (4, 4), (8, 8), 128, (15, 3), ((1, 9), (1, 3)))
    assert halo_frames(cfg) > 16
    ocfg = O.OracleConfig(80, cfg.upsample_rates, cfg.upsample_kernel_sizes, 128, cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)
    eng = Engine(cfg, 0)
    eng.load_state_dict(O.random_state_dict(ocfg, seed=4, loud=True), strict=True)
    eng.finalize()

    class Voc:
        engine = eng

        def __call__(self, mel):
            return eng.forward(np.asarray(mel) if np.ndim(mel) == 3 else np.asarray(mel)[None], precision="bf16x3")

    voc = Voc()
    rng = np.random.default_rng(2)
    mels = [rng.standard_normal((80, t)).astype(np.float32) for t in (150, 163, 171, 140, 90, 75, 200)]
    want = [voc(m)[0] for m in mels]
    got = synthesize_variable(voc, mels)
    for a, b in zip(got, want):
        assert a.shape == b.shape == (b.size,) and a.size % 16 == 0
        np.testing.assert_array_equal(a, b)
    short = synthesize_variable(voc, mels, halo=16, hop=16)
    assert any(not np.array_equal(a, b) for a, b in zip(short, want))
    eng.close()


def test_a_batch_that_does_not_fit_fails_cleanly_and_the_engine_lives_on():
    """A workspace the device cannot hold (4096 x 862 frames in bf16x3: 0.9 TB) comes back as HFG_ERR_NOMEM with a message -- no
    sticky CUDA error: the next forward on the same engine is correct."""
    import ctypes
    import iris.hifigan_pretrained as hp
    from iris_tts_b200 import _abi
    torch.manual_seed(0)
    m = hp.HiFiGANModel()
    m.to("cuda:0")
    eng = m.engine
    lib = _abi.load()
    assert lib.hfg_workspace_bytes(eng._h, 4096, 862, _abi.PREC_BF16X3) > 500e9
    dummy = (ctypes.c_float * 4)()
    p = ctypes.cast(dummy, ctypes.c_void_p)
    assert lib.hfg_forward(eng._h, p, 4096, 862, p, _abi.PREC_BF16X3, 0) == _abi.ERR_NOMEM
    assert b"out of memory" in lib.hfg_last_error()
    mel = O.synthetic_mel(2, 50, seed=8)
    ref = O.infer({k: v for k, v in m.state_dict().items()}, mel)
    assert np.abs(eng.forward(mel, precision="bf16x3") - ref).max() <= 1e-3
