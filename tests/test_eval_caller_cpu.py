"""CPU: the evaluation-time generation caller (reference data_loaders/humanml/motion_loaders/
comp_v6_model_dataset.py:148-283).  The engine is replaced by a stub sampler that is a deterministic function of the
noise it is given / draws, so the test pins the HOST logic: batch stacking of the multimodality repeats, generator
stream order, result bookkeeping -- against the reference's loop structure restated below (sequential calls)."""
from types import SimpleNamespace

import numpy as np
import torch

from b200mdm.data_loaders.humanml.motion_loaders.comp_v6_model_dataset import CompMDMGeneratedDataset

B, D, T, STEPS = 4, 263, 12, 3


class _Stub:
    """p_sample_loop stand-in: x_T + sum_k (k+1) eps_k + lengths (so that y is seen too); draws like the reference when
    no noise is passed (gaussian_diffusion.py:691, :525)."""
    num_timesteps = STEPS

    def __init__(self):
        self.calls = []

    def p_sample_loop(self, model, shape, noise=None, noise_tape=None, model_kwargs=None, **kw):
        self.calls.append(tuple(shape))
        x = noise if noise is not None else torch.randn(*shape)
        tape = noise_tape if noise_tape is not None else torch.stack([torch.randn_like(x) for _ in range(STEPS)])
        out = x.clone()
        for k in range(STEPS):
            out = out + (k + 1) * tape[k]
        y = model_kwargs["y"]
        assert y["lengths"].shape[0] == shape[0] and y["text_embed"].shape[1] == shape[0]
        return out + y["lengths"].view(-1, 1, 1, 1).float() + y["text_embed"][0, :, :1].view(-1, 1, 1, 1)


class _Model(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))


def _loader(n_batches):
    g = torch.Generator().manual_seed(1)
    batches = []
    for i in range(n_batches):
        lengths = torch.randint(4, T + 1, (B,), generator=g)
        y = dict(lengths=lengths, mask=(torch.arange(T)[None] < lengths[:, None]).view(B, 1, 1, T),
                 text=["caption %d %d" % (i, b) for b in range(B)], tokens=["sos/OTHER_walk/VERB_eos/OTHER_unk/OTHER"] * B,
                 text_embed=torch.randn(1, B, 512, generator=g))
        batches.append((torch.zeros(B, D, 1, T), {"y": y}))
    vec = {"sos/OTHER": (np.zeros(3), np.zeros(2)), "walk/VERB": (np.ones(3), np.ones(2)),
           "eos/OTHER": (np.zeros(3), np.zeros(2)), "unk/OTHER": (np.zeros(3), np.zeros(2))}

    class DS(SimpleNamespace):
        def __len__(self):
            return n_batches * B

    class L(list):
        batch_size = B
        dataset = DS(mode="gt", w_vectorizer=vec)
    return L(batches)


def _reference_structure(diffusion, loader, mm_idxs, mm_num_repeats, scale):
    """comp_v6_model_dataset.py:185-252 restated: one sample_fn call per repeat, noise drawn inside the call."""
    gen, mm = [], []
    for i, (motion, kw) in enumerate(loader):
        y = dict(kw["y"])
        y["scale"] = torch.ones(B) * scale
        is_mm = i in mm_idxs
        mm_motions = []
        for t in range(mm_num_repeats if is_mm else 1):
            s = diffusion.p_sample_loop(None, motion.shape, model_kwargs={"y": y})
            if t == 0:
                gen += [s[b].squeeze().permute(1, 0).numpy() for b in range(B)]
            if is_mm:
                mm_motions += [s[b].squeeze().permute(1, 0).numpy() for b in range(B)]
        if is_mm:
            mm += [mm_motions[b::B] for b in range(B)]
    return gen, mm


def test_stacked_repeats_equal_sequential_reference_calls():
    args = SimpleNamespace(autoregressive=False)
    loader = _loader(3)
    np.random.seed(3)
    torch.manual_seed(11)
    stub = _Stub()
    ds = CompMDMGeneratedDataset(args, _Model(), stub, loader, mm_num_samples=B, mm_num_repeats=3, max_motion_length=T,
                                 num_samples_limit=None, scale=2.5)
    np.random.seed(3)
    mm_idxs = np.sort(np.random.choice(3, B // B + 1, replace=False))
    torch.manual_seed(11)
    gen, mm = _reference_structure(_Stub(), _loader(3), mm_idxs, 3, 2.5)
    assert len(ds) == 3 * B and len(ds.mm_generated_motion) == len(mm_idxs) * B
    assert sorted(stub.calls) == sorted([(3 * B if i in mm_idxs else B, D, 1, T) for i in range(3)])   # one call per batch
    for a, b in zip(ds.generated_motion, gen):
        assert a["motion"].shape == (T, D) and np.array_equal(a["motion"], b)
    for a, b in zip(ds.mm_generated_motion, mm):
        assert len(a["mm_motions"]) == 3
        for m, r in zip(a["mm_motions"], b):
            assert np.array_equal(m["motion"], r)
    w, p, cap, sent_len, motion, m_len, tok = ds[5]
    assert cap == "caption 1 1" and sent_len == 3 and w.shape == (4, 3) and p.shape == (4, 2) and tok.count("_") == 3
    assert motion.shape == (T, D) and int(m_len) == int(loader[1][1]["y"]["lengths"][1])


def test_num_samples_limit_and_no_mm():
    args = SimpleNamespace(autoregressive=False)
    stub = _Stub()
    ds = CompMDMGeneratedDataset(args, _Model(), stub, _loader(4), mm_num_samples=0, mm_num_repeats=0, max_motion_length=T,
                                 num_samples_limit=2 * B, scale=1.)
    assert len(ds) == 2 * B and ds.mm_generated_motion == [] and stub.calls == [(B, D, 1, T)] * 2


def test_autoregressive_repeats_keep_the_reference_draw_order():
    """DiP evaluation (args.autoregressive): every repeat is a chain of pred_len chunks, each its own diffusion loop
    (sampler_util.py:41-81).  Stacked repeats + pre-drawn per-chunk noise must equal the sequential reference structure."""
    import b200mdm
    pred, ctx = 5, 2
    args = SimpleNamespace(autoregressive=True, pred_len=pred, context_len=ctx, autoregressive_include_prefix=False)
    g = torch.Generator().manual_seed(2)

    def loader():
        lengths = torch.full((B,), 196)
        y = dict(lengths=lengths, orig_lengths=lengths.clone(), mask=torch.ones(B, 1, 1, 196, dtype=torch.bool),
                 text=["c%d" % b for b in range(B)], tokens=["sos/OTHER_eos/OTHER"] * B,
                 text_embed=torch.randn(1, B, 512, generator=torch.Generator().manual_seed(4)),
                 prefix=torch.randn(B, D, 1, ctx, generator=torch.Generator().manual_seed(5)))

        class DS(SimpleNamespace):
            def __len__(self):
                return 2 * B

        class L(list):
            batch_size = B
            dataset = DS(mode="gt", w_vectorizer={"sos/OTHER": (np.zeros(3), np.zeros(2)), "eos/OTHER": (np.zeros(3), np.zeros(2))})
        return L([(torch.zeros(B, D, 1, 196), {"y": y})])

    class ArStub(_Stub):
        def p_sample_loop(self, model, shape, noise=None, noise_tape=None, model_kwargs=None, **kw):
            y = model_kwargs["y"]
            assert y["prefix"].shape == (shape[0], D, 1, ctx) and shape[-1] == pred
            out = super().p_sample_loop(model, shape, noise=noise, noise_tape=noise_tape, model_kwargs=model_kwargs)
            return out + y["prefix"].mean(dim=-1, keepdim=True)          # the chain depends on the handed-over prefix

    np.random.seed(0)
    torch.manual_seed(21)
    stub = ArStub()
    ds = CompMDMGeneratedDataset(args, _Model(), stub, loader(), mm_num_samples=1, mm_num_repeats=2, max_motion_length=196,
                                 num_samples_limit=None, scale=2.5)
    n_chunks = 196 // pred + 1
    assert stub.calls == [(2 * B, D, 1, pred)] * n_chunks                # one stacked loop per chunk
    # reference structure: for each repeat, AutoRegressiveSampler over sequentially drawn loops
    torch.manual_seed(21)
    ref = ArStub()
    kw = loader()[0][1]
    y = dict(kw["y"])
    y["scale"] = torch.ones(B) * 2.5
    outs = []
    for t in range(2):
        sampler = b200mdm.AutoRegressiveSampler(args, ref.p_sample_loop, required_frames=196)
        outs.append(sampler.sample(None, (B, D, 1, 196), model_kwargs={"y": y}))
    assert len(ds) == B and len(ds.mm_generated_motion) == B
    for b in range(B):
        assert np.array_equal(ds.generated_motion[b]["motion"], outs[0][b].squeeze().permute(1, 0).numpy())
        for t in range(2):
            assert np.array_equal(ds.mm_generated_motion[b]["mm_motions"][t]["motion"], outs[t][b].squeeze().permute(1, 0).numpy())
