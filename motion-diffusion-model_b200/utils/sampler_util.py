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
