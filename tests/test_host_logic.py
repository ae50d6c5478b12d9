"""CPU-only host logic: schedules/tables of the product package vs the oracle, weight-order contract, sharding."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import prodiff_oracle as oracle
from prodiff_b200 import schedule, sharding, synth
from prodiff_b200.engine import state_key_order

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("sched,timesteps,max_beta", [("vpsde", 4, 40), ("vpsde", 100, 0.06), ("linear", 8, 0.06),
                                                      ("cosine", 6, 0.06), ("logsnr", 5, 0.06), ("vpsde", 1000, 0.02)])
def test_module_buffers_equal_oracle_tables(sched, timesteps, max_beta):
    import prodiff_b200 as pb
    gd = pb.GaussianDiffusion(128, pb.WaveNet(128, 256, 1, 256, 1), timesteps=timesteps, schedule_type=sched,
                              max_beta=max_beta, spec_min=[-12.0], spec_max=[0.0])
    tb = oracle.diffusion_tables(timesteps, sched, max_beta)
    for k, v in tb.items():
        got = getattr(gd, k)
        assert got.shape == (timesteps + 1,)
        assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(v)), k
    assert gd.num_timesteps == timesteps and float(gd.timesteps) == timesteps
    assert gd.spec_min.shape == (1, 1, 1) and "spec_min" not in gd.state_dict()


def test_unknown_schedule_raises():
    with pytest.raises(NotImplementedError):
        schedule.get_noise_schedule_list("nope", 4)


def test_explicit_betas_tensor():
    import prodiff_b200 as pb
    betas = torch.linspace(1e-4, 0.05, 11)
    gd = pb.GaussianDiffusion(128, pb.WaveNet(128, 256, 1, 256, 1), timesteps=10, betas=betas, spec_min=[-12.0], spec_max=[0.0])
    assert torch.allclose(gd.betas, betas)


def test_state_key_order_matches_reference_state_dict():
    import prodiff_b200 as pb
    wn = pb.WaveNet(128, 256, 3, 256, 2)
    keys = state_key_order(3)
    assert sorted(keys) == sorted(wn.state_dict().keys()) and len(keys) == 10 + 8 * 3
    st = synth.make_state(synth.DiffNetConfig(128, 256, 3, 256, 2), 0)
    assert sorted(st) == sorted(keys)
    for k in keys:
        assert st[k].shape == wn.state_dict()[k].shape, k
    assert [b.dilation for b in wn.residual_layers] == [1, 2, 1]


def test_wavenet_from_reference_like_module():
    import prodiff_b200 as pb
    src = pb.WaveNet(64, 256, 5, 256, 3)
    src.load_state_dict(synth.make_state(synth.DiffNetConfig(64, 256, 5, 256, 3), 2))
    new = pb.WaveNet.from_reference(src, "fp16")
    assert new.cfg == src.cfg and new.operand_dtype == "fp16"
    for k, v in src.state_dict().items():
        assert torch.equal(v, new.state_dict()[k])


def test_encoder_dropin_state_dict_and_loud_failures():
    """prodiff_b200.FastspeechEncoder: the reference's state_dict keys / shapes / init (tts_modules.py:291-303), the order the
    C ABI expects, and no CPU path."""
    import os
    import sys
    import types
    import prodiff_b200 as pb
    from prodiff_b200.encoder import encoder_state_key_order
    cfg = synth.EncoderConfig(37, 256, 3, 9, 2)
    enc = pb.FastspeechEncoder(cfg.vocab_size, 256, 3, 9, dropout=0.1, num_heads=2)
    sd = enc.state_dict()
    keys = encoder_state_key_order(3)
    assert len(keys) == 3 + 10 * 3 and sorted(keys + ["embed_positions._float_tensor"]) == sorted(sd)
    st = synth.make_encoder_state(cfg, 3)
    assert sorted(st) == sorted(keys) and all(st[k].shape == sd[k].shape for k in keys)
    assert not sd["embed_tokens.weight"][0].any() and abs(float(sd["embed_tokens.weight"][1:].std()) - 256 ** -0.5) < 5e-3
    assert not sd["layers.0.op.ffn.ffn_2.bias"].any() and sd["layers.1.op.self_attn.in_proj_weight"].shape == (768, 256)
    enc.eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.ones(1, 3, dtype=torch.long))
    if os.path.isdir("/root/reference/modules/fastspeech"):
        sys.path.insert(0, "/root/reference")
        sys.modules.setdefault("chardet", types.ModuleType("chardet"))
        try:
            from modules.fastspeech.tts_modules import FastspeechEncoder as Ref
            ref = Ref(cfg.vocab_size, 256, 3, 9, dropout=0.1, num_heads=2)
            assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
            new = pb.FastspeechEncoder.from_reference(ref, "fp16")
            # (embed_positions._float_tensor is torch.FloatTensor(1): uninitialised memory, possibly NaN -- a device marker only)
            assert new.cfg == enc.cfg and all(torch.equal(v, new.state_dict()[k]) for k, v in ref.state_dict().items()
                                              if not k.endswith("_float_tensor"))
        finally:
            sys.path.remove("/root/reference")


def test_shard_bounds():
    assert sharding.shard_bounds(64, 8) == [(8 * i, 8 * i + 8) for i in range(8)]
    assert sharding.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sharding.shard_bounds(0, 3) == [(0, 0)] * 3
    for n, w in [(1, 1), (7, 3), (1024, 8), (5, 8)]:
        b = sharding.shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        sizes = [e - s for s, e in b]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 0)


def test_sample_sharded_without_process_group_is_identity():
    cond = torch.randn(3, 5, 4)
    out = sharding.sample_sharded(lambda c: c[..., :2] * 2, cond, gather=True)
    assert torch.equal(out, cond[..., :2] * 2)


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from prodiff_b200 import sharding
rank, world = int(sys.argv[2]), int(sys.argv[3])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[4], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
cond = torch.randn(5, 7, 4)                       # 5 utterances over 2 ranks: ragged shards (3, 2)
calls = []
def fake_sampler(c):                               # stands in for gd(cond, infer=True): row-independent
    calls.append(c.shape[0])
    return torch.stack([c.sum(-1), c.mean(-1), c.amax(-1)], dim=-1)
full = sharding.sample_sharded(fake_sampler, cond, gather=True)
mine = sharding.sample_sharded(fake_sampler, cond, gather=False)
s, e = sharding.shard_bounds(5, world)[rank]
assert calls == [e - s, e - s], calls
assert torch.equal(full, fake_sampler(cond)), "gathered result differs from the unsharded one"
assert torch.equal(mine, full[s:e])
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
'''


def test_sample_sharded_gloo_world2(tmp_path):
    """N>1 path on CPU: world_size 2, gloo, ragged split, all-gather of the finished shards only."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", port], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in o, o[-2000:]


def test_synth_is_deterministic_and_scaled():
    cfg = synth.DiffNetConfig(128, 256, 2, 256, 1)
    a, b = synth.make_state(cfg, 9), synth.make_state(cfg, 9)
    assert synth.digest(a) == synth.digest(b) != synth.digest(synth.make_state(cfg, 10))
    w = a["residual_layers.0.dilated_conv.weight"]
    assert abs(float(w.std()) - np.sqrt(2 / (256 * 3))) < 2e-3
    assert abs(float(a["output_projection.weight"].std()) - 0.02) < 2e-3
    assert float(synth.make_state(cfg, 9, "zero")["output_projection.weight"].abs().max()) == 0
    cond, xT, noise = synth.make_inputs(cfg, 2, 10, 3)
    assert cond.shape == (2, 10, 256) and xT.shape == (2, 1, 128, 10) and noise.shape == (3, 2, 1, 128, 10)
    assert 0 <= float(xT.min()) and float(xT.max()) < 1


def test_row_chunks_cover_batch_in_balanced_blocks():
    """Engine.sample walks large batches in row blocks (engine.row_chunks): contiguous, covering, balanced, bounded."""
    from prodiff_b200.engine import row_chunks
    assert row_chunks(64, 1000, 65536) == [(0, 64)]
    assert row_chunks(96, 1000, 65536) == [(0, 96)]                # up to 1.5 x the block size stays one call
    for B, T in ((256, 1000), (1024, 1000), (97, 1000), (8, 6000), (30, 6000), (5, 100000), (1, 10)):
        ch = row_chunks(B, T, 65536)
        assert ch[0][0] == 0 and ch[-1][1] == B and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        sizes = [e - s for s, e in ch]
        assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1
        if len(ch) > 1:
            assert max(sizes) <= max(1, 65536 // T)
    assert len(row_chunks(256, 1000, 65536)) == 4 and len(row_chunks(1024, 1000, 65536)) == 16


def test_indexed_dataset_roundtrip_and_reference_bytes(tmp_path):
    """offline.IndexedDatasetBuilder writes the reference's on-disk format (utils/indexed_datasets.py:51-76): round trip
    through our reader, across a segment boundary; when the reference is mounted, the files are byte-identical to the ones
    its own builder writes and its own reader returns the same items."""
    import os
    import sys
    import numpy as np
    from prodiff_b200 import offline
    rng = np.random.default_rng(0)
    items = [{"item_name": f"utt{i}", "x_0": rng.normal(size=(5 + i, 4)).astype(np.float32), "spk_id": i % 3} for i in range(7)]
    a = tmp_path / "ours"
    a.mkdir()
    b = offline.IndexedDatasetBuilder(str(a), "train", segment_size=3)
    for it in items:
        b.add_item(it)
    b.finalize()
    assert sorted(os.listdir(a)) == ["train_0.data", "train_0.idx", "train_1.data", "train_1.idx", "train_2.data", "train_2.idx"]
    ds = offline.IndexedDataset(str(a), "train", segment_size=3)
    assert len(ds) == 7
    for i, it in enumerate(items):
        got = ds[i]
        assert got["item_name"] == it["item_name"] and np.array_equal(got["x_0"], it["x_0"])
    with pytest.raises(IndexError):
        ds[7]
    ds.close()
    ref_root = "/root/reference"
    if os.path.isdir(os.path.join(ref_root, "utils")):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_indexed_datasets", os.path.join(ref_root, "utils", "indexed_datasets.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        r = tmp_path / "ref"
        r.mkdir()
        rb = mod.IndexedDatasetBuilder(str(r), "train", segment_size=3)
        for it in items:
            rb.add_item(it)
        rb.finalize()
        for f in sorted(os.listdir(r)):
            assert (r / f).read_bytes() == (a / f).read_bytes(), f
        rds = mod.IndexedDataset(str(a), "train", segment_size=3)      # the reference reader on OUR files
        assert np.array_equal(rds[4]["x_0"], items[4]["x_0"])


def test_bucket_by_length_properties():
    from prodiff_b200 import offline
    rng = np.random.default_rng(1)
    lengths = [int(v) for v in rng.integers(20, 3000, size=200)]
    batches = offline.bucket_by_length(lengths, frames_per_batch=65536, max_rows=64)
    flat = [i for b in batches for i in b]
    assert sorted(flat) == list(range(200))
    for b in batches:
        ls = [lengths[i] for i in b]
        assert ls == sorted(ls, reverse=True) and len(b) <= 64
        T = -(-ls[0] // 128) * 128
        assert len(b) == 1 or len(b) * T <= 65536
    # padding waste stays small because neighbours in a batch have similar lengths
    waste = sum(len(b) * lengths[b[0]] - sum(lengths[i] for i in b) for b in batches) / sum(lengths)
    assert waste < 0.25
    assert offline.my_batches(10, rank=1, world=4) == [1, 5, 9] and offline.my_batches(3, rank=0, world=1) == [0, 1, 2]
    with pytest.raises(ValueError):
        offline.bucket_by_length([10, 0, 5])


def test_workspace_bucket_bounds_slack_and_collapses_nearby_sizes():
    from prodiff_b200.engine import workspace_bucket
    for n in (1, 1 << 20, (1 << 20) + 1, 123_456_789, 2_200_000_000, 2_205_000_000, 21 << 30):
        b = workspace_bucket(n)
        assert b >= n and b >= 1 << 20 and b <= max(1 << 20, int(n * 1.125) + 1)
    assert workspace_bucket(2_200_000_000) == workspace_bucket(2_205_000_000)   # T = 1000 vs 1002 land in one bucket
    assert len({workspace_bucket(n) for n in range((1 << 30) + 1, (2 << 30) + 1, 1 << 22)}) == 8


def test_synthesizer_routes_arguments_like_the_reference_inferer():
    """pipeline.Synthesizer (handler/infer/handler.py:133-158 for a batch) with stand-ins for the CUDA pieces: conditioning
    keywords reach forward_condition, sampler extras reach diffusion, lengths default to the non-padding frames of mel2ph and
    go to BOTH the sampler and the generator, the generator input is asked for with the 2.30259 scale fused."""
    import types
    import prodiff_b200.pipeline as pl
    seen = {}

    class FakeCond:
        shape = torch.Size((2, 6, 256))

        def lengths(self):
            return [6, 4]

    def fake_forward_condition(teacher, txt, m2p, f0, validate=True, **kw):
        seen["cond_kw"] = kw
        return FakeCond()

    def diffusion(cond, **kw):
        seen["diff_kw"] = kw
        return torch.zeros(2, 128, 6)

    class Gen:
        upp = 4

        def __call__(self, c, f0, **kw):
            seen["gen_kw"] = kw
            return torch.arange(2 * 24, dtype=torch.float32).reshape(2, 1, 24)

    orig = pl.forward_condition
    pl.forward_condition = fake_forward_condition
    try:
        syn = pl.Synthesizer(types.SimpleNamespace(diffusion=diffusion), Gen(), infer_step=4)
        wavs = syn(torch.zeros(2, 3, dtype=torch.long), torch.zeros(2, 6, dtype=torch.long), torch.zeros(2, 6), voicing="v", seed=5,
                   rand_ini="ri", voc_seed=9)
    finally:
        pl.forward_condition = orig
    assert seen["cond_kw"] == {"voicing": "v"}
    assert seen["diff_kw"]["lengths"] == [6, 4] and seen["diff_kw"]["seed"] == 5 and seen["diff_kw"]["infer_step"] == 4
    assert abs(seen["diff_kw"]["vocoder_scale"] - 2.30259) < 1e-9 and seen["diff_kw"]["infer"] is True
    assert seen["gen_kw"] == {"lengths": [6, 4], "rand_ini": "ri", "seed": 9}
    assert [tuple(w.shape) for w in wavs] == [(24,), (16,)] and float(wavs[1][-1]) == 24 + 15
