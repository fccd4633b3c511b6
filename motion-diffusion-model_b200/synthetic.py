"""Deterministic synthetic checkpoints and inputs (no network => no released checkpoints).

`synthetic_state_dict` produces a reference-format ``state_dict`` (same key names and shapes as a
checkpoint written by the reference's train loop) from a numpy
``default_rng`` stream, so the very same weights can be regenerated in the build container, on the
GPU box and inside the reference oracle without shipping 70 MB files.  Scales follow the default
torch initialisers (Linear: U(+-1/sqrt(fan_in)); MHA in_proj: xavier-uniform) but biases and the
LayerNorm affine parameters are made non-trivial so that parity tests exercise them.
"""
import math

import numpy as np
import torch


def _uniform(rng, shape, bound):
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def _normal(rng, shape, std):
    return torch.from_numpy((rng.standard_normal(size=shape) * std).astype(np.float32))


def synthetic_state_dict(arch="trans_enc", latent_dim=512, ff_size=1024, num_layers=8, input_feats=263,
                         cond_dim=512, cond_mode="text", num_actions=1, seed=0):
    rng = np.random.default_rng(seed)
    d = latent_dim
    sd = {}

    def linear(prefix, out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        sd[prefix + ".weight"] = _uniform(rng, (out_f, in_f), b)
        sd[prefix + ".bias"] = _uniform(rng, (out_f,), b)

    def layer_norm(prefix):
        sd[prefix + ".weight"] = 1.0 + _normal(rng, (d,), 0.1)
        sd[prefix + ".bias"] = _normal(rng, (d,), 0.1)

    def mha(prefix):
        xb = math.sqrt(6.0 / (d + 3 * d))
        sd[prefix + ".in_proj_weight"] = _uniform(rng, (3 * d, d), xb)
        sd[prefix + ".in_proj_bias"] = _normal(rng, (3 * d,), 0.02)
        linear(prefix + ".out_proj", d, d)

    linear("input_process.poseEmbedding", d, input_feats)
    linear("embed_timestep.time_embed.0", d, d)
    linear("embed_timestep.time_embed.2", d, d)
    if "text" in cond_mode:
        linear("embed_text", d, cond_dim)
    if "action" in cond_mode:
        sd["embed_action.action_embedding"] = _normal(rng, (num_actions, d), 1.0)
    if arch == "trans_enc":
        for l in range(num_layers):
            p = "seqTransEncoder.layers.%d" % l
            mha(p + ".self_attn")
            linear(p + ".linear1", ff_size, d)
            linear(p + ".linear2", d, ff_size)
            layer_norm(p + ".norm1")
            layer_norm(p + ".norm2")
    elif arch == "trans_dec":
        for l in range(num_layers):
            p = "seqTransDecoder.layers.%d" % l
            mha(p + ".self_attn")
            mha(p + ".multihead_attn")
            linear(p + ".linear1", ff_size, d)
            linear(p + ".linear2", d, ff_size)
            layer_norm(p + ".norm1")
            layer_norm(p + ".norm2")
            layer_norm(p + ".norm3")
    else:
        raise ValueError("unsupported arch %r" % (arch,))
    linear("output_process.poseFinal", input_feats, d)
    return sd


def synthetic_inputs(batch, njoints=263, nfeats=1, nframes=196, steps=50, cond_dim=512, seed=10,
                     lengths=None, scale=2.5, dtype=torch.float32):
    """Noise tape [x_T, eps_{T-1} .. eps_0], text embedding [1,B,cond_dim], lengths, mask, scale."""
    rng = np.random.default_rng(seed)
    shape = (batch, njoints, nfeats, nframes)
    tape = [torch.from_numpy(rng.standard_normal(size=shape).astype(np.float32)) for _ in range(steps + 1)]
    text_embed = torch.from_numpy(rng.standard_normal(size=(1, batch, cond_dim)).astype(np.float32))
    if lengths is None:
        lengths = [nframes] * batch
    lengths = torch.tensor(lengths, dtype=torch.int64)
    mask = (torch.arange(nframes)[None, :] < lengths[:, None]).view(batch, 1, 1, nframes)
    if not torch.is_tensor(scale):
        scale = torch.full((batch,), float(scale), dtype=torch.float32)
    return dict(tape=tape, text_embed=text_embed, lengths=lengths, mask=mask, scale=scale)


def synthetic_dip_inputs(batch, n_tokens, context_len, njoints=263, nfeats=1, cond_dim=768, seed=3):
    """Deterministic DiP conditioning in the reference's layouts (model/mdm.py:180-187,204): BERT token features
    [n_tokens, B, 768], ragged padding mask [B, n_tokens] (True = padding; sample 0 has none), prefix [B, J, F, ctx]."""
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(n_tokens, batch, cond_dim, generator=g)
    tmask = torch.zeros(batch, n_tokens, dtype=torch.bool)
    for b in range(1, batch):
        tmask[b, n_tokens - (b * 2) % n_tokens:] = True
    prefix = torch.randn(batch, njoints, nfeats, context_len, generator=g)
    return enc, tmask, prefix


def synthetic_norm_stats(dim=263, seed=7):
    """Stand-in for the dataset's Mean.npy / Std.npy (data_loaders/humanml/data/dataset.py:248-249): fp32 [dim]."""
    rng = np.random.default_rng(seed)
    mean = rng.standard_normal(dim).astype(np.float32)
    std = rng.uniform(0.2, 2.0, size=dim).astype(np.float32)
    return torch.from_numpy(mean), torch.from_numpy(std)
