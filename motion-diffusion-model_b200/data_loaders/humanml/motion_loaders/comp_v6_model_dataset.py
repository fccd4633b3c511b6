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
    chunk count), in the order the reference would draw them; stacked along the batch."""
    xs = [[None] * repeats for _ in range(chunks)]
    es = [[None] * repeats for _ in range(chunks)]
    for t in range(repeats):
        for c in range(chunks):
            x = torch.randn(*shape, device=device)
            xs[c][t] = x
            es[c][t] = torch.stack([torch.randn_like(x) for _ in range(n_steps)])
    noise = torch.stack([torch.cat(xs[c], dim=0) for c in range(chunks)])            # [chunks, repeats*B, ...]
    tape = torch.stack([torch.cat(es[c], dim=1) for c in range(chunks)])             # [chunks, n_steps, repeats*B, ...]
    return noise, tape


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
                noise, tape = _draw_noise(repeat_times, chunks, n_steps, draw_shape, device)
                ys = _repeat_y({k: v for k, v in y.items() if k != "text"}, repeat_times)
                kw = dict(clip_denoised=clip_denoised, model_kwargs={"y": ys}, skip_timesteps=0, init_image=None,
                          progress=False, dump_steps=None, const_noise=False)
                stacked_shape = (repeat_times * shape[0],) + shape[1:]
                if autoregressive:
                    sample = sample_fn(model, stacked_shape, noise=noise, noise_tape=tape, **kw)
                else:
                    sample = sample_fn(model, stacked_shape, noise=noise[0], noise_tape=tape[0], **kw)
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
