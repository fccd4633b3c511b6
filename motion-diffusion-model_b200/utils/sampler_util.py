"""Sampling-time model wrappers (host mirror of the reference's utils/sampler_util.py)."""
import torch
import torch.nn as nn

from .misc import wrapped_getattr


class ClassifierFreeSampleModel(nn.Module):
    """Classifier-free guidance wrapper (reference utils/sampler_util.py:10-38).

    The reference deep-copies `y`, runs the denoiser twice and blends
    `out_uncond + scale * (out - out_uncond)`.  Here the cond / uncond pair is packed into ONE batch of 2B inside the
    engine and the blend is applied to the hidden rows in front of the (linear) output projection, which is the same
    expression in exact arithmetic; `y` is never copied or mutated.
    """

    def __init__(self, model):
        super().__init__()
        self.model = model
        assert self.model.cond_mask_prob > 0, \
            "Cannot run a guided diffusion on a model that has not been trained with no conditions"
        self.rot2xyz = self.model.rot2xyz
        self.translation = self.model.translation
        self.njoints = self.model.njoints
        self.nfeats = self.model.nfeats
        self.data_rep = self.model.data_rep
        self.cond_mode = self.model.cond_mode
        self.encode_text = self.model.encode_text

    def forward(self, x, timesteps, y=None):
        assert self.model.cond_mode in ["text", "action"]
        from ..model.mdm import _run_model
        return _run_model(self.model, x, timesteps, y, guided=True)

    def __getattr__(self, name, default=None):
        return wrapped_getattr(self, name, default=None)


class AutoRegressiveSampler:
    """DiP's outer loop (reference utils/sampler_util.py:41-81): generate `required_frames` as a chain of `pred_len`
    chunks, each a full diffusion loop of the trans_dec engine conditioned on the last `context_len` frames of the
    previous chunk (y['prefix']).  Host control flow only; every chunk is one `sample_fn` call, i.e. one replay of the
    engine's captured step graph per diffusion step.  The caller's kwargs are never mutated (the reference deep-copies
    them per chunk; here only the dicts that change are rebuilt)."""

    def __init__(self, args, sample_fn, required_frames=196):
        self.sample_fn = sample_fn
        self.args = args
        self.required_frames = required_frames

    def sample(self, model, shape, **kargs):
        pred_len, context_len = self.args.pred_len, self.args.context_len
        n_iterations = self.required_frames // pred_len + int(self.required_frames % pred_len > 0)
        y0 = kargs["model_kwargs"]["y"]
        cur_prefix = y0["prefix"].clone()
        dynamic_text_mode = isinstance(y0["text"][0], list) if "text" in y0 else False   # a prompt per chunk
        samples_buf = [cur_prefix] if getattr(self.args, "autoregressive_include_prefix", False) else []
        ar_shape = list(shape)
        ar_shape[-1] = pred_len
        tape = kargs.get("noise_tape")            # b200mdm extension: one tape per chunk, [n_iterations, n_run+1, ...]
        for i in range(n_iterations):
            y = dict(y0)
            y["prefix"] = cur_prefix
            if dynamic_text_mode:
                y["text"] = [s[i] for s in y0["text"]]
                if getattr(model, "text_encoder_type", "bert") != "bert":
                    raise NotImplementedError("DiP model only supports BERT text encoder at the moment.")
                y["text_embed"] = (y0["text_embed"][0][:, :, i], y0["text_embed"][1][:, i])
            cur = dict(kargs)
            cur["model_kwargs"] = dict(kargs["model_kwargs"], y=y)
            if tape is not None:
                cur["noise_tape"] = tape[i]
            if kargs.get("noise") is not None and kargs["noise"].dim() == len(ar_shape) + 1:   # x_T per chunk
                cur["noise"] = kargs["noise"][i]
            sample = self.sample_fn(model, ar_shape, **cur)
            samples_buf.append(sample[..., -pred_len:].clone())
            cur_prefix = sample[..., -context_len:].clone()
        return torch.cat(samples_buf, dim=-1)[..., :self.required_frames]
