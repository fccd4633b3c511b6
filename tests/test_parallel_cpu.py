"""CPU, world_size 2 over gloo: the multi-GPU host logic (batch sharding, the single text-embedding broadcast, rank-sliced
noise, gather).  The per-rank sampler is the fp32 oracle loop on a tiny model, so the check is end-to-end: the 2-rank
result must equal the unsharded run (to fp32 round-off on CPU; bitwise on the GPU, see test_parity_gpu.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b200mdm
from b200mdm import parallel
from oracle import mdm_oracle as mo, schedule_oracle as so

L, STEPS, B, T = 1, 3, 5, 12


def _oracle_sampler():
    W = mo.OracleWeights(b200mdm.synthetic_state_dict(num_layers=L, seed=3), L)
    tabs = so.diffusion_tables(so.named_betas("cosine", STEPS))

    def fn(model, shape, noise, model_kwargs, noise_tape, **kw):
        y = model_kwargs["y"]
        tape = [noise] + [noise_tape[k] for k in range(noise_tape.shape[0])]
        return mo.sample_loop(W, tabs, list(range(STEPS)), tape, y["text_embed"], y["scale"], y["lengths"])
    return fn


def _inputs():
    g = torch.Generator().manual_seed(5)
    lengths = torch.tensor([12, 7, 12, 3, 9])
    return dict(y=dict(text_embed=torch.randn(1, B, 512, generator=g), lengths=lengths, scale=torch.tensor([2.5, 1.0, 0.0, 4.0, 2.5]),
                       mask=(torch.arange(T)[None] < lengths[:, None]).view(B, 1, 1, T), text=["p%d" % i for i in range(B)]))


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        kw = _inputs()
        if rank != 0:                       # only rank 0 "encoded the prompts"
            kw["y"]["text_embed"] = torch.zeros_like(kw["y"]["text_embed"])
        out = parallel.sample_sharded(_oracle_sampler(), None, (B, 263, 1, T), kw, n_steps=STEPS, noise_mode="global",
                                      seed=77, device=torch.device("cpu"))
        assert out.shape == (B, 263, 1, T)
        lo, hi = parallel.shard_range(B, rank, world)
        assert (hi - lo) == (3 if rank == 0 else 2)
        if rank == 0:
            torch.save(out, out_path)
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_batch():
    for b in (1, 5, 64, 512, 513):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(b, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_shard_model_kwargs_slices_batch_keys():
    kw = _inputs()
    s = parallel.shard_model_kwargs(kw, 1, 4)["y"]
    assert s["text_embed"].shape == (1, 3, 512) and s["lengths"].tolist() == [7, 12, 3] and s["mask"].shape[0] == 3
    assert s["text"] == ["p1", "p2", "p3"] and kw["y"]["lengths"].shape[0] == B   # original untouched


def test_two_rank_sharded_run_equals_unsharded(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    sharded = torch.load(out_path)
    single = parallel.sample_sharded(_oracle_sampler(), None, (B, 263, 1, T), _inputs(), n_steps=STEPS, noise_mode="global",
                                     seed=77, device=torch.device("cpu"))
    # the CPU oracle's matmuls are not batch-size invariant bit for bit (MKL blocking); the CUDA engine is, and
    # tests/test_parity_gpu.py::test_benchmark_size_properties checks the bitwise version of this on the GPU
    assert torch.allclose(sharded, single, rtol=0, atol=2e-5)
    assert not torch.allclose(sharded[0], sharded[1], atol=1e-2)


# ---------------------------------------------------------------------------------------------------------------------
# DiP: the broadcast object is the (BERT tokens, padding mask) pair; y['prefix'] is sharded with the batch
CTX, PRED, MT = 6, 10, 5


def _dip_sampler():
    W = mo.OracleWeights(b200mdm.synthetic_state_dict(arch="trans_dec", num_layers=L, cond_dim=768, seed=6), L)
    tabs = so.diffusion_tables(so.named_betas("cosine", STEPS))

    def fn(model, shape, noise, model_kwargs, noise_tape, **kw):
        y = model_kwargs["y"]
        enc, tmask = y["text_embed"]
        tape = [noise] + [noise_tape[k] for k in range(noise_tape.shape[0])]
        return mo.sample_loop_dec(W, tabs, list(range(STEPS)), tape, enc, tmask, y["prefix"], y["scale"], y["lengths"])
    return fn


def _dip_inputs():
    enc, tmask, prefix = b200mdm.synthetic_dip_inputs(B, MT, CTX, seed=9)
    lengths = torch.tensor([10, 7, 10, 3, 9])
    return dict(y=dict(text_embed=(enc, tmask), prefix=prefix, lengths=lengths, scale=torch.tensor([7.5, 1.0, 0.0, 4.0, 2.5]),
                       mask=(torch.arange(PRED)[None] < lengths[:, None]).view(B, 1, 1, PRED)))


def _dip_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        kw = _dip_inputs()
        if rank != 0:                       # only rank 0 ran the text tower
            enc, tmask = kw["y"]["text_embed"]
            kw["y"]["text_embed"] = (torch.zeros_like(enc), torch.ones_like(tmask))
        out = parallel.sample_sharded(_dip_sampler(), None, (B, 263, 1, PRED), kw, n_steps=STEPS, noise_mode="global",
                                      seed=78, device=torch.device("cpu"))
        if rank == 0:
            torch.save(out, out_path)
    finally:
        dist.destroy_process_group()


def test_dip_two_rank_sharded_run_equals_unsharded(tmp_path):
    kw = _dip_inputs()
    s = parallel.shard_model_kwargs(kw, 1, 4)["y"]
    assert s["text_embed"][0].shape == (MT, 3, 768) and s["text_embed"][1].shape == (3, MT) and s["prefix"].shape[0] == 3
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out_path = str(tmp_path / "out.pt")
    mp.spawn(_dip_worker, args=(2, port, out_path), nprocs=2, join=True)
    sharded = torch.load(out_path)
    single = parallel.sample_sharded(_dip_sampler(), None, (B, 263, 1, PRED), _dip_inputs(), n_steps=STEPS,
                                     noise_mode="global", seed=78, device=torch.device("cpu"))
    assert sharded.shape == (B, 263, 1, PRED)
    assert torch.allclose(sharded, single, rtol=0, atol=5e-5)


def test_sample_sharded_rejects_batch_smaller_than_world(monkeypatch):
    """a rank with an empty shard used to raise before the all_gather and hang the others."""
    import pytest
    import torch.distributed as dist
    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "get_world_size", lambda group=None: 4)
    monkeypatch.setattr(dist, "get_rank", lambda group=None: 3)
    with pytest.raises(ValueError):
        parallel.sample_sharded(lambda *a, **k: None, None, (2, 263, 1, 8), {"y": {}}, n_steps=3)


def test_sample_sharded_philox_mode_passes_global_sample_index():
    """noise_mode='philox' (default): nothing is drawn on the host; the sampler receives the seed and the global index of
    the shard's first sample, which is what keys the engine's noise stream (G GPUs == 1 GPU, bit for bit)."""
    seen = {}

    def fn(model, shape, model_kwargs, noise_seed, sample_index_base, **kw):
        seen.update(shape=shape, seed=noise_seed, base=sample_index_base, B=model_kwargs["y"]["lengths"].shape[0])
        return torch.zeros(shape)
    y = {"lengths": torch.arange(5), "text_embed": torch.zeros(1, 5, 512)}
    out = parallel.sample_sharded(fn, None, (5, 263, 1, 8), {"y": y}, n_steps=3, seed=42, device=torch.device("cpu"))
    assert seen == dict(shape=(5, 263, 1, 8), seed=42, base=0, B=5) and out.shape == (5, 263, 1, 8)


def test_shard_replications_partition():
    """eval_humanml's replications split over ranks: disjoint, complete, round-robin."""
    for n, world in ((20, 8), (5, 8), (3, 1), (7, 2)):
        parts = [parallel.shard_replications(n, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert all(p == list(range(r, n, world)) for r, p in enumerate(parts))
    assert parallel.gather_objects({"a": 1}) == [{"a": 1}]
