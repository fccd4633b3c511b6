"""Host mirror of the reference's evaluation-time generation caller:
`CompMDMGeneratedDataset` of data_loaders/humanml/motion_loaders/comp_v6_model_dataset.py:148-283, the loop behind
`eval_humanml` (12-15 h on the reference: 32-sample loader batches, 10 multimodality repeats on some of them).

Same constructor, same `generated_motion` / `mm_generated_motion` / `__getitem__` contract.  What changes is how the
engine is driven: the `mm_num_repeats` repeats of a multimodality batch are ONE sampling loop over a batch of
repeats x B (the denoiser treats samples independently, so this is the same computation; the engine workspace and its
step graph are sized once per batch shape).  The noise of the stacked loop is drawn repeat by repeat in the reference's
order (x_T, then one eps per step, then the next repeat -- gaussian_diffusion.py:691,525), so a seeded run consumes the
generator exactly like the reference's sequential calls and yields the same motions.
"""
import numpy as np
import torch
from torch.utils.data import Dataset

from ....utils.sampler_util import AutoRegressiveSampler


def _repeat_y(y, times):
    """Stack `times` copies of the reference's y dict along the batch (index = repeat * B + sample)."""
    if times == 1:
        return y
    out = {}
    for k, v in y.items():
        if k == "text_embed" and torch.is_tensor(v):                   # CLIP features [1, B, C]
            out[k] = v.repeat(1, times, 1) if v.shape[1] > 1 else v
        elif k == "text_embed" and isinstance(v, tuple):               # DiP: (tokens [Mt, B, C], mask [B, Mt])
            out[k] = (v[0].repeat(1, times, 1), v[1].repeat(times, 1))
        elif torch.is_tensor(v) and v.dim() > 0:
            out[k] = v.repeat(times, *([1] * (v.dim() - 1)))
        elif isinstance(v, (list, tuple)):
            out[k] = list(v) * times
        else:
            out[k] = v
    return out


def _draw_noise(repeats, chunks, n_steps, shape, device):
    """x_T and eps for `repeats` sequential reference calls, each of `chunks` diffusion loops (1, or the autoregressive
    chunk count), in the order the reference would draw them; stacked along the batch, written straight into ONE
    preallocated pair of tensors (no stack / cat copies)."""
    B = shape[0]
    noise = torch.empty((chunks, repeats * B) + tuple(shape[1:]), device=device)
    tape = torch.empty((chunks, n_steps, repeats * B) + tuple(shape[1:]), device=device)
    for t in range(repeats):
        for c in range(chunks):
            noise[c, t * B:(t + 1) * B] = torch.randn(*shape, device=device)
            for k in range(n_steps):
                tape[c, k, t * B:(t + 1) * B].normal_()
    return noise, tape


# A stacked multimodality batch is only worth it while its noise fits: the tape of ONE repeat of the reference's
# mm_short evaluation (bs 32, 263 x 196, 1000 steps) is 6.6 GB.
TAPE_BUDGET_BYTES = 8 << 30


class _RepeatStreams:
    """The generator streams of `repeats` SEQUENTIAL reference calls, readable in any interleaving.

    The reference runs repeat r's whole loop (x_T, then one randn_like per step) before repeat r+1 starts, all from
    torch's default generator.  Drawing repeat r's k-th eps before repeat r-1 has finished needs a generator positioned
    at `offset_0 + r * (n_steps + 1) * delta`, where delta is what one randn of this shape advances the Philox offset
    by (measured, not assumed).  One cloned generator per repeat provides exactly that; the default generator is left
    where the sequential calls would have left it."""

    def __init__(self, repeats, n_steps, shape, device):
        main = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
        probe = main.clone_state()
        o0 = probe.get_offset()
        torch.randn(*shape, device=device, generator=probe)
        delta = probe.get_offset() - o0
        self.gens = []
        for r in range(repeats):
            g = main.clone_state()
            g.set_offset(main.get_offset() + r * (n_steps + 1) * delta)
            self.gens.append(g)
        main.set_offset(main.get_offset() + repeats * (n_steps + 1) * delta)
        self.B, self.shape, self.device = shape[0], tuple(shape), device

    def x_T(self):
        return torch.cat([torch.randn(*self.shape, device=self.device, generator=g) for g in self.gens], dim=0)

    def fill(self, buf, k0):                                   # buf [n, repeats*B, ...]: eps of steps k0 .. k0+n-1
        for j in range(buf.shape[0]):
            for r, g in enumerate(self.gens):
                buf[j, r * self.B:(r + 1) * self.B].normal_(generator=g)


def _offset_api(device):
    try:
        g = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
        g.clone_state().get_offset()
        return True
    except Exception:
        return False


class CompMDMGeneratedDataset(Dataset):

    def __init__(self, args, model, diffusion, dataloader, mm_num_samples, mm_num_repeats, max_motion_length,
                 num_samples_limit, scale=1.):
        self.args = args
        self.dataloader = dataloader
        self.dataset = dataloader.dataset
        self.model = model
        assert mm_num_samples < len(dataloader.dataset)
        clip_denoised = False                                          # hard-coded in the reference (:157)
        self.max_motion_length = max_motion_length
        sample_fn = diffusion.p_sample_loop                            # use_ddim hard-coded False (:156)
        autoregressive = bool(getattr(args, "autoregressive", False))
        if autoregressive:
            sample_fn = AutoRegressiveSampler(args, sample_fn).sample
        device = next(model.parameters()).device
        n_steps = diffusion.num_timesteps

        real_num_batches = len(dataloader)
        if num_samples_limit is not None:
            real_num_batches = min(num_samples_limit // dataloader.batch_size + 1, real_num_batches)
        generated_motion = []
        mm_generated_motions = []
        if mm_num_samples > 0:
            mm_idxs = np.sort(np.random.choice(real_num_batches, mm_num_samples // dataloader.batch_size + 1, replace=False))
        else:
            mm_idxs = []
        model.eval()
        bs = dataloader.batch_size

        with torch.no_grad():
            for i, (motion, model_kwargs) in enumerate(dataloader):
                if num_samples_limit is not None and len(generated_motion) >= num_samples_limit:
                    break
                y = {k: v.to(device) if torch.is_tensor(v) else v for k, v in model_kwargs["y"].items()}
                tokens = [t.split("_") for t in y["tokens"]]
                if scale != 1.:
                    y["scale"] = torch.ones(motion.shape[0], device=device) * scale
                if "text" in y and "text_embed" not in y:              # encode the prompts once, not once per repeat
                    y["text_embed"] = model.encode_text(y["text"])
                is_mm = i in mm_idxs
                repeat_times = mm_num_repeats if is_mm else 1
                shape = tuple(motion.shape)
                if autoregressive:
                    chunks = 196 // args.pred_len + int(196 % args.pred_len > 0)      # AutoRegressiveSampler default
                    draw_shape = shape[:-1] + (args.pred_len,)
                else:
                    chunks, draw_shape = 1, shape
                ys_full = {k: v for k, v in y.items() if k != "text"}
                kw = dict(clip_denoised=clip_denoised, skip_timesteps=0, init_image=None, progress=False, dump_steps=None,
                          const_noise=False)
                per_repeat = 4 * n_steps * chunks * int(np.prod(draw_shape))
                if repeat_times * per_repeat <= TAPE_BUDGET_BYTES or autoregressive:
                    # small enough: all repeats in one loop, noise drawn up front in the reference's order
                    noise, tape = _draw_noise(repeat_times, chunks, n_steps, draw_shape, device)
                    stacked_shape = (repeat_times * shape[0],) + shape[1:]
                    kw["model_kwargs"] = {"y": _repeat_y(ys_full, repeat_times)}
                    if autoregressive:
                        sample = sample_fn(model, stacked_shape, noise=noise, noise_tape=tape, **kw)
                    else:
                        sample = sample_fn(model, stacked_shape, noise=noise[0], noise_tape=tape[0], **kw)
                elif device.type == "cuda" and _offset_api(device):
                    # long loops (1000 steps): still ONE stacked loop, eps produced chunk by chunk from per-repeat
                    # generator clones positioned where the reference's sequential calls would be
                    streams = _RepeatStreams(repeat_times, n_steps, draw_shape, device)
                    stacked_shape = (repeat_times * shape[0],) + shape[1:]
                    kw["model_kwargs"] = {"y": _repeat_y(ys_full, repeat_times)}
                    sample = sample_fn(model, stacked_shape, noise=streams.x_T(), noise_fn=streams.fill, **kw)
                else:
                    # no offset API: groups of repeats sized by the budget, sequential like the reference
                    group = max(1, int(TAPE_BUDGET_BYTES // per_repeat))
                    parts = []
                    for r0 in range(0, repeat_times, group):
                        g = min(group, repeat_times - r0)
                        noise, tape = _draw_noise(g, 1, n_steps, draw_shape, device)
                        kw["model_kwargs"] = {"y": _repeat_y(ys_full, g)}
                        parts.append(sample_fn(model, (g * shape[0],) + shape[1:], noise=noise[0], noise_tape=tape[0], **kw))
                        del noise, tape
                    sample = torch.cat(parts, dim=0)
                if "prefix" in y:                                      # :216-217
                    y["lengths"] = y["orig_lengths"]
                sample = sample.reshape(repeat_times, shape[0], *sample.shape[1:])
                lengths = y["lengths"].cpu().numpy()
                first = sample[0].squeeze(2).permute(0, 2, 1).cpu().numpy()          # [B, T, D]
                generated_motion += [{
                    "motion": first[b], "length": lengths[b], "caption": y["text"][b], "tokens": tokens[b],
                    "cap_len": tokens[b].index("eos/OTHER") + 1,                     # reference issue #182 fix (:226-229)
                } for b in range(bs)]
                if is_mm:
                    allm = sample.squeeze(3).permute(0, 1, 3, 2).cpu().numpy()      # [repeats, B, T, D]
                    if self.dataset.mode == "eval":                                  # :236-238, T2M normalisation
                        allm = self.dataset.t2m_dataset.inv_transform(allm)
                        allm = (allm - self.dataset.mean_for_eval) / self.dataset.std_for_eval
                    mm_generated_motions += [{
                        "caption": y["text"][b], "tokens": tokens[b], "cap_len": len(tokens[b]),
                        "mm_motions": [{"motion": allm[t, b], "length": lengths[b]} for t in range(repeat_times)],
                    } for b in range(bs)]

        self.generated_motion = generated_motion
        self.mm_generated_motion = mm_generated_motions
        self.w_vectorizer = dataloader.dataset.w_vectorizer

    def __len__(self):
        return len(self.generated_motion)

    def __getitem__(self, item):
        data = self.generated_motion[item]
        motion, m_length, caption, tokens = data["motion"], data["length"], data["caption"], data["tokens"]
        sent_len = data["cap_len"]
        if self.dataset.mode == "eval":                                # T2M evaluators expect their own norms (:269-274)
            motion = (self.dataset.t2m_dataset.inv_transform(motion) - self.dataset.mean_for_eval) / self.dataset.std_for_eval
        pos_one_hots, word_embeddings = [], []
        for token in tokens:
            word_emb, pos_oh = self.w_vectorizer[token]
            pos_one_hots.append(pos_oh[None, :])
            word_embeddings.append(word_emb[None, :])
        return (np.concatenate(word_embeddings, axis=0), np.concatenate(pos_one_hots, axis=0), caption, sent_len, motion,
                m_length, "_".join(tokens))
