"""GPU: parity at the headline sizes (VERDICT r1 'next round' item 1) and the host features added in round 2.

  * BASELINE config 2 exactly (B=64, T=196, 50 steps, CFG 2.5): the GPU runs the batch, the fp32 oracle follows 3 samples;
    Frobenius-relative AND max element-wise errors are reported;
  * BASELINE config 3's per-GPU shard (B=64, 1000 steps, CFG 2.5): the oracle follows 2 samples through all 1000
    recurrent steps, error-vs-step curve printed; the noise is the engine's counter-based stream (no 13 GB tape);
  * a non-identity timestep_map (1000-step model respaced to 50) through the fused loop;
  * MDM.encode_text with stub text towers (clip and bert branches of model/mdm.py:163-187);
  * engine hygiene: workspace pool, schedule change, weight reload, chunked generator draws, Philox stream.
Tolerance: 1e-3 relative (BASELINE.json north_star), Frobenius norm; the element-wise figure is normalised by max|ref|."""
import sys
import types
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import b200mdm
from conftest import default_args, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def _errs(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    fro = float((a - b).norm() / b.norm())
    mx = float((a - b).abs().max() / b.abs().max())
    return fro, mx


def _enc(layers, steps, seed, **over):
    args = default_args(layers=layers, diffusion_steps=steps, **over)
    model, diffusion = b200mdm.create_model_and_diffusion(args, SimpleNamespace(dataset=SimpleNamespace()))
    sd = b200mdm.synthetic_state_dict(num_layers=layers, seed=seed)
    b200mdm.load_model_wo_clip(model, sd)
    model.to("cuda").eval()
    return b200mdm.ClassifierFreeSampleModel(model), model, diffusion, sd


def _y(inp, dev="cuda"):
    return dict(mask=inp["mask"].to(dev), lengths=inp["lengths"].to(dev), text_embed=inp["text_embed"].to(dev),
                scale=inp["scale"].to(dev))


# ---------------------------------------------------------------------------------------------------------------------
def test_c2_exact_config_b64_50_steps():
    """BASELINE config 2 as written: B=64, 196 frames, 263 features, 50 steps, CFG 2.5, L=8 (explicit noise tape)."""
    from oracle import mdm_oracle as mo, schedule_oracle as so
    B, T, steps = 64, 196, 50
    cfg, _, diffusion, sd = _enc(8, steps, 0)
    inp = b200mdm.synthetic_inputs(B, nframes=T, steps=steps, seed=10)
    xT = inp["tape"][0].cuda()
    tape = torch.stack(inp["tape"][1:]).cuda()
    out = diffusion.p_sample_loop(cfg, (B, 263, 1, T), noise=xT, clip_denoised=False, model_kwargs={"y": _y(inp)}, noise_tape=tape)
    assert torch.isfinite(out).all()
    idx = [0, 31, 63]
    W = mo.OracleWeights(sd, 8)
    tabs = so.diffusion_tables(so.named_betas("cosine", steps))
    ref = mo.sample_loop(W, tabs, list(range(steps)), [t[idx] for t in inp["tape"]], inp["text_embed"][:, idx],
                         inp["scale"][idx], inp["lengths"][idx])
    fro, mx = _errs(out[idx], ref)
    print("C2 B=64 x 50 steps x CFG 2.5: Frobenius-relative %.3e, max|diff|/max|ref| %.3e" % (fro, mx))
    assert fro < RTOL and mx < 5 * RTOL


class _LazyTape:
    """tape[0] = x_T, tape[1+k] = eps of the k-th step, fetched from the ENGINE's Philox stream for the followed samples
    (the very values the fused loop consumes; the numpy restatement of the stream is checked separately)."""

    def __init__(self, eng, idx, shape1, seed, n_steps):
        self.eng, self.idx, self.shape1, self.seed, self.n = eng, idx, shape1, seed, n_steps

    def __getitem__(self, k):
        step_id = -1 if k == 0 else self.n - k        # k-th executed step has schedule index n-1-(k-1)
        rows = [self.eng.philox_normal((1,) + self.shape1, self.seed, g, step_id, "cuda") for g in self.idx]
        return torch.cat(rows, 0).cpu()


class _Pick(list):
    def __init__(self, keep):
        super().__init__()
        self.keep, self.count, self.got = set(keep), 0, {}

    def append(self, x):
        self.count += 1
        if self.count in self.keep:
            self.got[self.count] = x


def test_c3_shard_b64_1000_steps_cfg():
    """BASELINE config 3, one GPU's shard: HumanML shapes, CFG 2.5, 1000 recurrent steps of the fused loop."""
    from oracle import mdm_oracle as mo, schedule_oracle as so
    B, T, steps, seed = 64, 196, 1000, 77
    cfg, model, diffusion, sd = _enc(8, steps, 0)
    inp = b200mdm.synthetic_inputs(B, nframes=T, steps=1, seed=12)
    y = _y(inp)
    eng = model.engine()
    torch.cuda.reset_peak_memory_stats()
    base_mem = torch.cuda.memory_allocated()
    marks = [1, 10, 100, 300, 600, 900, 1000]
    # the same loop, in segments, so that intermediate states can be compared (segments do not change the result)
    diffusion._prepare(cfg, (B, 263, 1, T), {"y": y}, torch.device("cuda"), 0.0)
    eng.set_noise_stream(seed, 0)
    x = eng.philox_normal((B, 263, 1, T), seed, 0, -1, "cuda")
    states, done = {}, 0
    for m in marks:
        out = torch.empty_like(x)
        eng.sample_loop_range(b200mdm._lib.MODE_DDPM, steps - 1 - done, m - done, x if done == 0 else None, out, None, 0, True)
        states[m], done = out, m
    whole = diffusion.p_sample_loop(cfg, (B, 263, 1, T), clip_denoised=False, model_kwargs={"y": y}, noise_seed=seed)
    assert torch.equal(whole, states[steps])                       # one call == segmented calls, bit for bit
    peak = torch.cuda.max_memory_allocated() - base_mem
    print("C3 shard: torch peak memory during the 1000-step loops %.1f MB (no noise tape)" % (peak / 2 ** 20))
    assert peak < 2 << 30
    idx = [3, 60]
    W = mo.OracleWeights(sd, 8)
    tabs = so.diffusion_tables(so.named_betas("cosine", steps))
    pick = _Pick(marks)
    mo.sample_loop(W, tabs, list(range(steps)), _LazyTape(eng, idx, (263, 1, T), seed, steps), inp["text_embed"][:, idx],
                   inp["scale"][idx], inp["lengths"][idx], collect=pick)
    worst = 0.0
    for m in marks:
        fro, mx = _errs(states[m][idx], pick.got[m])
        print("C3 shard step %4d: Frobenius-relative %.3e, max|diff|/max|ref| %.3e" % (m, fro, mx))
        worst = max(worst, fro)
    assert worst < RTOL


def test_respaced_timestep_map_through_fused_loop():
    """a5: a 1000-step model sampled with 50 respaced steps (SpacedDiffusion(space_timesteps(1000, '50'))): the device
    gather timestep_map[state.cur] feeds the timestep embedding; schedule = the respaced betas."""
    from oracle import mdm_oracle as mo, schedule_oracle as so
    from b200mdm.diffusion import gaussian_diffusion as gd
    L, B, T = 2, 3, 40
    cfg, model, _, sd = _enc(L, 1000, 6)
    betas = gd.get_named_beta_schedule("cosine", 1000)
    for spec, n in (("50", 50), ("ddim25", 25), ([10, 15, 20], 45)):
        diffusion = b200mdm.SpacedDiffusion(use_timesteps=b200mdm.space_timesteps(1000, spec), betas=betas,
                                            model_mean_type=gd.ModelMeanType.START_X,
                                            model_var_type=gd.ModelVarType.FIXED_SMALL, loss_type=gd.LossType.MSE,
                                            rescale_timesteps=False)
        assert diffusion.num_timesteps == n and diffusion.timestep_map != list(range(n))
        use = so.space_timesteps(1000, spec)
        new_betas, tmap, _ = so.respaced(so.named_betas("cosine", 1000), use)
        assert list(tmap) == list(diffusion.timestep_map)
        inp = b200mdm.synthetic_inputs(B, nframes=T, steps=n, seed=40 + n, lengths=[40, 23, 7], scale=torch.tensor([2.5, 1.0, 4.0]))
        W = mo.OracleWeights(sd, L)
        ref = mo.sample_loop(W, so.diffusion_tables(new_betas), tmap, inp["tape"], inp["text_embed"], inp["scale"], inp["lengths"])
        out = diffusion.p_sample_loop(cfg, (B, 263, 1, T), noise=inp["tape"][0].cuda(), clip_denoised=False,
                                      model_kwargs={"y": _y(inp)}, noise_tape=torch.stack(inp["tape"][1:]).cuda())
        e = rel_err(out, ref)
        print("respacing %r: %d steps, relative error %.3e" % (spec, n, e))
        assert e < RTOL
        if n == 25:                                                 # DDIM on the respaced schedule as well
            ref = mo.sample_loop(W, so.diffusion_tables(new_betas), tmap, inp["tape"], inp["text_embed"], inp["scale"],
                                 inp["lengths"], sampler="ddim", eta=0.0)
            out = diffusion.ddim_sample_loop(cfg, (B, 263, 1, T), noise=inp["tape"][0].cuda(), clip_denoised=False, eta=0.0,
                                             model_kwargs={"y": _y(inp)}, noise_tape=torch.stack(inp["tape"][1:]).cuda())
            assert rel_err(out, ref) < RTOL


def test_encode_text_seam_with_stub_towers(monkeypatch):
    """a23: MDM.encode_text (model/mdm.py:163-187) with stand-in text towers returning known tensors -- the clip branch
    (tokenize to 22 tokens, zero-pad to 77, encode, float, unsqueeze) and the bert branch (permute + mask inversion),
    then y['text'] through p_sample_loop (encoded once, cached into y like gaussian_diffusion.py:633-635)."""
    B, T, steps = 3, 24, 4
    cfg, model, diffusion, sd = _enc(2, steps, 1)
    feats = torch.randn(B, 512, generator=torch.Generator().manual_seed(5))
    seen = {}
    clip_stub = types.ModuleType("clip")

    def tokenize(texts, context_length=77, truncate=False):
        seen["tok"] = (list(texts), context_length, truncate)
        return torch.arange(len(texts) * context_length, dtype=torch.int64).view(len(texts), context_length) % 97 + 1
    clip_stub.tokenize = tokenize
    monkeypatch.setitem(sys.modules, "clip", clip_stub)

    class Tower:
        def encode_text(self, tokens):
            seen["shape"], seen["pad_zero"] = tuple(tokens.shape), bool((tokens[:, 22:] == 0).all())
            seen["device"] = tokens.device.type
            return feats.to(tokens.device).half()                  # CLIP returns fp16 on GPU: .float() in the seam
    model.clip_model = Tower()
    prompts = ["a person walks", "a person jumps", "someone waves"]
    te = model.encode_text(prompts)
    assert seen["tok"] == (prompts, 22, True) and seen["shape"] == (B, 77) and seen["pad_zero"] and seen["device"] == "cuda"
    assert te.shape == (1, B, 512) and te.dtype == torch.float32
    assert torch.equal(te[0].cpu(), feats.half().float())
    inp = b200mdm.synthetic_inputs(B, nframes=T, steps=steps, seed=11, lengths=[24, 17, 5])
    y = _y(inp)
    y.pop("text_embed")
    y["text"] = prompts
    tape, xT = torch.stack(inp["tape"][1:]).cuda(), inp["tape"][0].cuda()
    a = diffusion.p_sample_loop(cfg, (B, 263, 1, T), noise=xT, clip_denoised=False, model_kwargs={"y": y}, noise_tape=tape)
    assert "text_embed" in y and torch.equal(y["text_embed"], te)   # cached into the caller's dict like the reference
    y2 = _y(inp)
    y2["text_embed"] = te
    b = diffusion.p_sample_loop(cfg, (B, 263, 1, T), noise=xT, clip_denoised=False, model_kwargs={"y": y2}, noise_tape=tape)
    assert torch.equal(a, b)
    # bert branch: tower returns (features [B, Mt, 768], mask [B, Mt] True = token present)
    args = default_args(layers=2, diffusion_steps=3, arch="trans_dec", text_encoder_type="bert", context_len=20, pred_len=40)
    dmodel, _ = b200mdm.create_model_and_diffusion(args, SimpleNamespace(dataset=SimpleNamespace()))
    enc = torch.randn(B, 7, 768)
    present = torch.tensor([[1] * 7, [1] * 4 + [0] * 3, [1] + [0] * 6], dtype=torch.bool)
    dmodel.clip_model = lambda texts: (enc, present)
    tok, pad = dmodel.encode_text(prompts)
    assert tok.shape == (7, B, 768) and torch.equal(tok, enc.permute(1, 0, 2)) and torch.equal(pad, ~present)
    bare = b200mdm.create_model_and_diffusion(default_args(layers=1), SimpleNamespace(dataset=SimpleNamespace()))[0]
    with pytest.raises(RuntimeError):
        bare.encode_text(prompts)                                   # no tower attached: explicit error, no fallback


# ---------------------------------------------------------------------------------------------------------------------
def test_philox_stream_vs_numpy_oracle():
    from oracle import philox_oracle as po
    _, model, _, _ = _enc(1, 4, 1)
    eng = model.engine()
    for (B, n, seed, base, step) in [(3, 1001, 1234, 5, 7), (2, 51548, 99, 2 ** 33 + 1, -1), (1, 8, 2 ** 63 + 5, 0, 999)]:
        got = eng.philox_normal((B, n), seed, base, step, "cuda").cpu().numpy()
        want = po.normal(B, n, seed, base, step)
        assert np.isfinite(got).all()
        assert np.abs(got - want).max() < 2e-6, (B, n)               # logf / sincospif differ from numpy in the last ulp


def test_philox_loop_is_split_invariant_and_matches_oracle():
    """noise_seed mode: a sub-batch run with sample_index_base reproduces the rows of the full-batch run bit for bit
    (what makes parallel.sample_sharded's G-GPU result equal the 1-GPU result), and the loop agrees with the oracle."""
    from oracle import mdm_oracle as mo, schedule_oracle as so
    L, steps, B, T, seed = 2, 20, 6, 33, 4242
    cfg, model, diffusion, sd = _enc(L, steps, 3)
    inp = b200mdm.synthetic_inputs(B, nframes=T, steps=1, seed=21, lengths=[33, 30, 20, 9, 2, 1])
    y = _y(inp)
    full = diffusion.p_sample_loop(cfg, (B, 263, 1, T), clip_denoised=False, model_kwargs={"y": y}, noise_seed=seed)
    lo, hi = 2, 5
    ys = dict(mask=y["mask"][lo:hi], lengths=y["lengths"][lo:hi], text_embed=y["text_embed"][:, lo:hi].contiguous(), scale=y["scale"][lo:hi])
    part = diffusion.p_sample_loop(cfg, (hi - lo, 263, 1, T), clip_denoised=False, model_kwargs={"y": ys}, noise_seed=seed,
                                   sample_index_base=lo)
    assert torch.equal(part, full[lo:hi])
    eng = model.engine()
    tape = _LazyTape(eng, list(range(B)), (263, 1, T), seed, steps)
    ref = mo.sample_loop(mo.OracleWeights(sd, L), so.diffusion_tables(so.named_betas("cosine", steps)), list(range(steps)),
                         tape, inp["text_embed"], inp["scale"], inp["lengths"])
    assert rel_err(full, ref) < RTOL


def test_generator_chunked_draws_match_explicit_tape():
    """noise=None over more steps than one NOISE_CHUNK: the side-stream, triple-buffered draws consume torch's generator
    exactly like the reference's per-step th.randn_like (gaussian_diffusion.py:525)."""
    L, steps, B, T = 1, 50, 2, 16
    cfg, _, diffusion, _ = _enc(L, steps, 2)
    inp = b200mdm.synthetic_inputs(B, nframes=T, steps=1, seed=3)
    shape = (B, 263, 1, T)
    assert steps > 2 * diffusion.NOISE_CHUNK
    for skip in (0, 7):
        torch.manual_seed(321)
        a = diffusion.p_sample_loop(cfg, shape, clip_denoised=False, model_kwargs={"y": _y(inp)}, skip_timesteps=skip)
        torch.manual_seed(321)
        xT = torch.randn(*shape, device="cuda")
        tape = torch.stack([torch.randn_like(xT) for _ in range(steps - skip)])
        b = diffusion.p_sample_loop(cfg, shape, noise=xT, clip_denoised=False, model_kwargs={"y": _y(inp)}, noise_tape=tape,
                                    skip_timesteps=skip)
        assert torch.equal(a, b), skip


def test_engine_state_hygiene():
    """(1) two diffusion objects with different step counts on one model and one (B, T) -- the schedule tables
    move / change under a captured graph; (2) a forward before load_state_dict must not leak the old input-projection
    bias (pe_bias) into later forwards; (3) alternating batch shapes reuse pooled workspaces and stay correct."""
    from oracle import mdm_oracle as mo, schedule_oracle as so
    from b200mdm.utils.model_util import create_gaussian_diffusion
    L, B, T = 2, 3, 24
    cfg, model, d1200, sd = _enc(L, 1200, 1)                      # 1200 pre-embedded model timesteps
    d4, d6 = (create_gaussian_diffusion(default_args(layers=L, diffusion_steps=n)) for n in (4, 6))
    W = mo.OracleWeights(sd, L)

    def run(diffusion, steps, batch=B, seed=11):
        inp = b200mdm.synthetic_inputs(batch, nframes=T, steps=steps, seed=seed, lengths=[T] * batch)
        out = diffusion.p_sample_loop(cfg, (batch, 263, 1, T), noise=inp["tape"][0].cuda(), clip_denoised=False,
                                      model_kwargs={"y": _y(inp)}, noise_tape=torch.stack(inp["tape"][1:]).cuda())
        ref = mo.sample_loop(W, so.diffusion_tables(so.named_betas("cosine", steps)), list(range(steps)), inp["tape"],
                             inp["text_embed"], inp["scale"], inp["lengths"])
        return rel_err(out, ref)
    assert run(d4, 4) < RTOL
    assert run(d6, 6) < RTOL                                      # same (B, T), other schedule: graph must not go stale
    assert run(d4, 4) < RTOL
    assert run(d1200, 1200, batch=1, seed=2) < RTOL               # > 1000 steps: the device tables are re-allocated
    assert run(d4, 4) < RTOL
    for batch in (5, B, 7, B, 5, 2, 9, B):                        # pool of workspaces, LRU eviction beyond 4
        assert run(d4, 4, batch=batch, seed=20 + batch) < RTOL
    # weight reload after a forward at the same shape
    sd2 = b200mdm.synthetic_state_dict(num_layers=L, seed=77)
    b200mdm.load_model_wo_clip(model, sd2)
    W = mo.OracleWeights(sd2, L)
    assert run(d4, 4) < RTOL
