"""Timestep respacing (host mirror of the reference's diffusion/respace.py).

`space_timesteps` is pure integer logic and must be bit-exact, including Python's round-half-to-even at the
fractional stride (reference respace.py:58).  `SpacedDiffusion` re-derives the betas of the kept steps and records
`timestep_map`; the map is handed to the engine (b200mdm_set_schedule) which performs `_WrappedModel`'s gather
`new_ts = map[ts]` (respace.py:125-127) on the device, once per step, from the device-side step counter.
"""
import numpy as np

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Subset of an original process' steps to keep (reference respace.py:9-62)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    per, extra = divmod(num_timesteps, len(section_counts))
    start, picked = 0, []
    for i, count in enumerate(section_counts):
        size = per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            picked.append(start + round(cur))
            cur += frac
        start += size
    return set(picked)


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process that keeps only `use_timesteps` of a base process (reference respace.py:65-115)."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        last = 1.0
        new_betas = []
        for i, acp in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - acp / last)
                last = acp
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def _timestep_map(self):
        return list(self.timestep_map)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t


class _WrappedModel:
    """Callable that remaps spaced indices to model timesteps (reference respace.py:118-134).  Only used when a
    caller drives the model step by step through the reference-style `model(x, ts, **kwargs)` protocol."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps

    def __call__(self, x, ts, **kwargs):
        import torch
        map_tensor = torch.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        new_ts = map_tensor[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)

    def __getattr__(self, name):
        return getattr(self.__dict__["model"], name)
