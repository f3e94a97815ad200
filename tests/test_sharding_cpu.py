"""N > 1 host logic on CPU: world_size-2 gloo processes run the sharding code with the ORACLE standing in for the
CUDA generator (tests may use oracle/ as the checker).  Covers SURVEY.md section 8(e): batch shards need no
collective; a long mel split along time with a 16-frame halo and one gather equals the unchunked forward."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iris_tts_b200 import sharding
from oracle import hifigan_oracle as O


def test_batch_shards_are_contiguous_and_balanced():
    assert sharding.batch_shards(16, 8) == [(2 * i, 2 * i + 2) for i in range(8)]
    assert sharding.batch_shards(5, 4) == [(0, 2), (2, 3), (3, 4), (4, 5)]
    assert sharding.batch_shards(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]      # B < n_gpu: idle ranks
    assert sharding.batch_shards(0, 2) == [(0, 0), (0, 0)]
    with pytest.raises(ValueError):
        sharding.batch_shards(4, 0)


def test_time_chunks_cover_the_mel_with_halos_on_inner_edges_only():
    ch = sharding.time_chunks(10336, 8)            # BASELINE config 4: 120 s over 8 GPUs
    assert [c.frames for c in ch] == [1292] * 8
    assert ch[0].lo == 0 and ch[0].trim_front == 0 and ch[0].trim_back == 16
    assert ch[7].hi == 10336 and ch[7].trim_back == 0 and ch[7].trim_front == 16
    assert all(c.trim_front == 16 and c.trim_back == 16 for c in ch[1:7])
    assert sum(c.frames for c in ch) == 10336
    ragged = sharding.time_chunks(101, 4)
    assert [c.frames for c in ragged] == [26, 25, 25, 25] and ragged[-1].stop == 101
    tiny = sharding.time_chunks(3, 4)
    assert [c.frames for c in tiny] == [1, 1, 1, 0]
    assert tiny[1].lo == 0 and tiny[1].hi == 3     # halo clipped at the true sequence edges


def test_single_process_longform_is_the_plain_forward():
    sd = O.random_state_dict(O.V2, seed=0, loud=True)
    mel = torch.from_numpy(O.synthetic_mel(1, 40, seed=3))
    synth = lambda m: O.forward(sd, m, O.V2)  # noqa: E731
    out = sharding.synthesize_longform(synth, mel, hop=256)
    assert torch.equal(out, synth(mel).reshape(-1))


def test_chunked_equals_unchunked_and_halo_requirement():
    """Receptive field of the generator is +-12.63 frames: halo 16 is exact to fp32 noise, halo 4 is not."""
    sd = O.random_state_dict(O.V2, seed=0, loud=True)
    mel = torch.from_numpy(O.synthetic_mel(1, 120, seed=9, realistic=True))
    synth = lambda m: O.forward(sd, m, O.V2)  # noqa: E731
    full = synth(mel).reshape(-1)

    def stitched(halo):
        return torch.cat([sharding.synthesize_chunk(synth, mel, c, 256) for c in sharding.time_chunks(120, 3, halo)])

    assert stitched(16).shape == full.shape
    # fp32 noise only (oneDNN picks different blockings per length; loud weights put the output at std 0.2)
    assert float((stitched(16) - full).abs().max()) <= 5e-6
    assert float((stitched(4) - full).abs().max()) > 1e-4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd = O.random_state_dict(O.V2, seed=0, loud=True)
        synth = lambda m: O.forward(sd, m, O.V2)  # noqa: E731
        # long-form: odd length so the two chunks differ in size (exercises the padded gather slot)
        mel = torch.from_numpy(O.synthetic_mel(1, 75, seed=11, realistic=True))
        out = sharding.synthesize_longform(synth, mel, hop=256)
        out_all = sharding.synthesize_longform(synth, mel[0], hop=256, all_ranks=True)
        # batch shards: 3 utterances over 2 ranks, no collective
        melb = torch.from_numpy(O.synthetic_mel(3, 20, seed=12))
        part, (s, e) = sharding.synthesize_batch_sharded(synth, melb)
        q.put((rank, None if out is None else out.numpy(), out_all.numpy(), part.numpy(), (s, e)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_streaming_pieces_concatenate_to_the_full_waveform():
    """Streaming chunked output against the oracle forward: pieces of 40 frames + 16-frame halo == unchunked."""
    from oracle import hifigan_oracle as O
    sd = O.random_state_dict(O.V2, seed=0, loud=True)
    mel = torch.from_numpy(O.synthetic_mel(1, 130, seed=5))
    synth = lambda m: O.forward(sd, m, O.V2)                  # noqa: E731
    full = synth(mel).reshape(-1)
    chunks = sharding.stream_chunks(130, 40)
    assert [(c.start, c.stop) for c in chunks] == [(0, 40), (40, 80), (80, 120), (120, 130)]
    assert chunks[0].lo == 0 and chunks[1].lo == 24 and chunks[-1].hi == 130
    pieces = list(sharding.synthesize_streaming(synth, mel, 40))
    assert [p.numel() for p in pieces] == [40 * 256, 40 * 256, 40 * 256, 10 * 256]
    assert float((torch.cat(pieces) - full).abs().max()) <= 2e-6
    with pytest.raises(ValueError):
        sharding.stream_chunks(10, 0)


def test_world_size_2_gloo_longform_gather_and_batch_shards():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        r = q.get(timeout=240)
        results[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sd = O.random_state_dict(O.V2, seed=0, loud=True)
    full = O.forward(sd, torch.from_numpy(O.synthetic_mel(1, 75, seed=11, realistic=True)), O.V2).reshape(-1).numpy()
    out0, all0, part0, r0 = results[0]
    out1, all1, part1, r1 = results[1]
    assert out1 is None and out0.shape == full.shape               # only dst holds the stitched waveform
    assert np.abs(out0 - full).max() <= 5e-6
    np.testing.assert_array_equal(all0, all1)
    np.testing.assert_array_equal(all0, out0)
    fullb = O.forward(sd, torch.from_numpy(O.synthetic_mel(3, 20, seed=12)), O.V2).numpy()
    assert (r0, r1) == ((0, 2), (2, 3))
    assert np.abs(np.concatenate([part0, part1]) - fullb).max() <= 5e-6


# ---------------------------------------------------------------------------
# ragged batches: length buckets + one tail pass (iris_tts_b200/batching.py), with the oracle standing in for the engine
# ---------------------------------------------------------------------------

def test_halo_derived_from_the_constructor_arguments_covers_the_receptive_field():
    from iris_tts_b200.engine import V1, V2, V3
    for cfg in (V1, V2, V3):
        assert 12 <= sharding.halo_frames(cfg) <= sharding.HALO_FRAMES
    # empirically: perturbing the mel `halo` frames away does not change a sample, one frame inside the bound it may
    sd = O.random_state_dict(O.V2, seed=0, loud=True)
    mel = torch.from_numpy(O.synthetic_mel(1, 80, seed=4))
    base = O.forward(sd, mel, O.V2).reshape(-1)
    h = sharding.halo_frames(V2)
    bumped = mel.clone()
    bumped[:, :, 40 + h:] += 1.0
    out = O.forward(sd, bumped, O.V2).reshape(-1)
    assert torch.equal(out[: 40 * 256], base[: 40 * 256])
    assert not torch.equal(out[: (40 + h) * 256], base[: (40 + h) * 256])


def test_length_buckets_bound_the_padding():
    from iris_tts_b200.batching import length_buckets
    lengths = [400, 33, 371, 64, 390, 127, 350, 32, 398, 129, 65, 301]
    b = length_buckets(lengths, max_pad=0.15)
    assert sorted(i for g in b for i in g) == list(range(len(lengths)))
    for g in b:
        top = max(lengths[i] for i in g)
        assert all(lengths[i] >= 0.85 * top for i in g)
    assert len(b) < len(lengths)
    assert all(len(g) <= 2 for g in length_buckets(lengths, max_pad=0.15, max_batch=2))
    assert length_buckets([], 0.1) == []


def test_ragged_batch_scheme_is_exact_against_per_utterance_forwards():
    from iris_tts_b200.batching import synthesize_variable
    sd = O.random_state_dict(O.V2, seed=0, loud=True)
    calls = []

    def vocoder(batch):
        calls.append(batch.shape)
        return O.infer(sd, batch, O.V2)

    lengths = (90, 33, 84, 64, 7, 70, 32, 7, 0, 31)
    mels = [O.synthetic_mel(1, t, seed=20 + i)[0] if t else np.zeros((80, 0), np.float32) for i, t in enumerate(lengths)]
    stats = {}
    outs = synthesize_variable(vocoder, mels, stats=stats)
    n_calls = len(calls)
    assert stats["calls"] == n_calls < len(set(lengths))
    for m, o, t in zip(mels, outs, lengths):
        assert o.shape == (t * 256,) and o.dtype == np.float32
        if t:
            np.testing.assert_allclose(o, O.infer(sd, m[None], O.V2)[0], atol=2e-6)
    # padded lengths quantised to multiples of 64 frames (plan reuse in a serving loop): still exact
    calls.clear()
    outs_q = synthesize_variable(vocoder, mels, length_quantum=64)
    assert all(shape[2] % 64 == 0 for shape in calls if shape[2] > 32 and shape[2] != 2 * 16)
    for a, b_ in zip(outs, outs_q):
        np.testing.assert_allclose(a, b_, atol=2e-6)
    # a halo smaller than the receptive field is NOT exact (the scheme depends on it)
    bad = synthesize_variable(vocoder, mels[:1], halo=4)
    assert np.abs(bad[0] - O.infer(sd, mels[0][None], O.V2)[0]).max() > 1e-4
    with pytest.raises(ValueError):
        synthesize_variable(vocoder, [np.zeros((1, 80, 5), np.float32)])


def test_ragged_shards_balance_frames_not_counts():
    from iris_tts_b200.sharding import ragged_shards
    rng = np.random.default_rng(0)
    lengths = [int(x) for x in rng.integers(50, 900, size=37)]
    for world in (1, 2, 3, 8):
        parts = ragged_shards(lengths, world)
        assert sorted(i for p in parts for i in p) == list(range(37))          # a partition
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)                          # LPT bound
        assert parts == ragged_shards(lengths, world)                           # deterministic: no collective needed
    assert ragged_shards([5, 7], 4) == [[1], [0], [], []]
    assert ragged_shards([], 2) == [[], []]
    with pytest.raises(ValueError):
        ragged_shards([1], 0)


def test_native_ragged_path_host_logic(monkeypatch):
    """A vocoder that offers ``forward_ragged`` in a tensor-core precision gets ONE padded call per length bucket with the items'
    lengths and no tail pass; the fp32 mode, plain callables and HFG_RAGGED=0 keep the dense-call scheme.  (The oracle stands in
    for the engine: the fake forward_ragged runs every item alone, which is what the engine's ragged plan must equal.)"""
    from iris_tts_b200.batching import ragged_forward_of, synthesize_variable
    sd = O.random_state_dict(O.V2, seed=0, loud=True)

    class Model:
        precision = "bf16"
        ragged_calls = []
        dense_calls = 0

        def forward_ragged(self, mel, lengths):
            assert mel.shape[0] == len(lengths) and max(lengths) <= mel.shape[2]
            self.ragged_calls.append((mel.shape, tuple(lengths)))
            out = np.full((mel.shape[0], mel.shape[2] * 256), np.nan, dtype=np.float32)   # behind an item's end: unspecified
            for b, n in enumerate(lengths):
                out[b, : n * 256] = O.infer(sd, np.ascontiguousarray(mel[b:b + 1, :, :n]), O.V2)[0]
            return out

    class Voc:
        model = Model()

        def __call__(self, batch):
            self.model.dense_calls += 1
            return O.infer(sd, batch, O.V2)

    voc = Voc()
    lengths = (90, 33, 84, 64, 7, 70, 0, 31)
    mels = [O.synthetic_mel(1, t, seed=40 + i)[0] if t else np.zeros((80, 0), np.float32) for i, t in enumerate(lengths)]
    stats = {}
    outs = synthesize_variable(voc, mels, stats=stats, hop=256, halo=16)
    assert stats["native_ragged"] and stats["calls"] == len(voc.model.ragged_calls) < len(set(lengths)) and voc.model.dense_calls == 0
    assert sum(len(l) for _s, l in voc.model.ragged_calls) == len([t for t in lengths if t])
    for m, o, t in zip(mels, outs, lengths):
        assert o.shape == (t * 256,) and o.dtype == np.float32 and np.isfinite(o).all()
        if t:
            np.testing.assert_allclose(o, O.infer(sd, m[None], O.V2)[0], atol=2e-6)
    outs_q = synthesize_variable(voc, mels, hop=256, halo=16, length_quantum=64)
    assert all(shape[2] % 64 == 0 for shape, _l in voc.model.ragged_calls[stats["calls"]:])
    for a, b_ in zip(outs, outs_q):
        np.testing.assert_array_equal(a, b_)
    # a vocoder that can take all buckets at once gets them in ONE call
    many_calls = []

    def forward_ragged_batches(self, batches):
        many_calls.append(len(batches))
        return [self.forward_ragged(m, l) for m, l in batches]

    Model.forward_ragged_batches = forward_ragged_batches
    outs_m = synthesize_variable(voc, mels, hop=256, halo=16)
    assert many_calls == [stats["calls"]]
    for a, b_ in zip(outs, outs_m):
        np.testing.assert_array_equal(a, b_)
    # a generator the engine cannot plan ragged (HFG_ERR_UNSUPPORTED) falls back to the dense-call scheme; other errors surface
    class Unsupported(RuntimeError):
        code = -5

    def refuse(self, batches):
        raise Unsupported("no ragged plan")

    Model.forward_ragged_batches = refuse
    before = voc.model.dense_calls
    stats_f = {}
    outs_f = synthesize_variable(voc, mels, stats=stats_f, hop=256, halo=16)
    assert not stats_f["native_ragged"] and voc.model.dense_calls > before
    for a, b_ in zip(outs, outs_f):
        np.testing.assert_allclose(a, b_, atol=2e-6)

    def broken(self, batches):
        raise RuntimeError("something else")

    Model.forward_ragged_batches = broken
    with pytest.raises(RuntimeError, match="something else"):
        synthesize_variable(voc, mels, hop=256, halo=16)
    Model.forward_ragged_batches = forward_ragged_batches
    # no native path: unknown precisions, plain callables, HFG_RAGGED=0
    voc.model.precision = "int8"
    assert ragged_forward_of(voc) is None
    voc.model.precision = "fp16"
    assert ragged_forward_of(voc) is not None and ragged_forward_of(lambda m: m) is None
    monkeypatch.setenv("HFG_RAGGED", "0")
    assert ragged_forward_of(voc) is None
    stats2 = {}
    synthesize_variable(voc, mels[:3], stats=stats2, hop=256, halo=16)
    assert not stats2["native_ragged"] and voc.model.dense_calls > 0


def test_numa_binding_is_a_no_op_without_nvml_or_gpu():
    from iris_tts_b200 import numa
    before = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert numa.gpu_local_cpus(0) is None and numa.bind_process_to_gpu(0) is None
    assert numa.bind_process_to_gpu(10 ** 6) is None          # no such GPU
    assert os.sched_getaffinity(0) == before


def test_chunking_helpers_read_hop_and_halo_from_the_vocoder():
    """resolve_geometry: explicit arguments win; else the generator configuration found on the vocoder object (as .model.engine.config,
    .engine.config, .model.config or .config); else the V1 values for a plain callable."""
    from iris_tts_b200 import sharding
    from iris_tts_b200.engine import V1, V3, GeneratorConfig

    class Eng:
        def __init__(self, cfg):
            self.config = cfg

    class Model:
        def __init__(self, cfg):
            self.engine = Eng(cfg)

    class Gen:
        def __init__(self, cfg):
            self.model = Model(cfg)

    assert sharding.resolve_geometry(lambda m: m, None, None) == (256, 16)
    assert sharding.resolve_geometry(lambda m: m, 64, 20) == (64, 20)
    assert sharding.resolve_geometry(Gen(V1), None, None) == (256, sharding.halo_frames(V1)) == (256, 15)
    assert sharding.resolve_geometry(Model(V3), None, None) == (256, 13)
    wide = GeneratorConfig(80, (4, 4), (8, 8), 128, (15, 3), ((1, 9), (1, 3)))
    hop, halo = sharding.resolve_geometry(Eng(wide), None, None)
    assert hop == 16 and halo == sharding.halo_frames(wide) > 16
    assert sharding.resolve_geometry(Eng(wide), None, 40) == (16, 40)
