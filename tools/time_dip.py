"""Time the DiP configuration (BASELINE config 3): trans_dec, 8 layers, B=128, 40-frame chunks with a 20-frame prefix,
10 diffusion steps per chunk, guidance 7.5 -- ms per chunk (one fused loop) and per 196-frame autoregressive motion batch."""
import os, sys
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200mdm

B, ctx, pred, Mt, steps = 128, 20, 40, 16, 10
args = SimpleNamespace(dataset="humanml", unconstrained=False, latent_dim=512, layers=8, cond_mask_prob=0.1, arch="trans_dec",
                       emb_trans_dec=False, text_encoder_type="bert", pos_embed_max_len=5000, mask_frames=True, pred_len=pred,
                       context_len=ctx, diffusion_steps=steps, noise_schedule="cosine", sigma_small=True, lambda_vel=0.0,
                       lambda_rcxyz=0.0, lambda_fc=0.0, autoregressive_include_prefix=False)
model, diffusion = b200mdm.create_model_and_diffusion(args, SimpleNamespace(dataset=SimpleNamespace()))
b200mdm.load_model_wo_clip(model, b200mdm.synthetic_state_dict(arch="trans_dec", num_layers=8, cond_dim=768, seed=23))
model.to("cuda").eval()
cfg = b200mdm.ClassifierFreeSampleModel(model)
enc, tmask, prefix = b200mdm.synthetic_dip_inputs(B, Mt, ctx, seed=35)
y = dict(mask=torch.ones(B, 1, 1, pred, dtype=torch.bool, device="cuda"), lengths=torch.full((B,), pred, device="cuda"),
         text_embed=(enc.cuda(), tmask.cuda()), prefix=prefix.cuda(), scale=torch.full((B,), 7.5, device="cuda"))
g = torch.Generator(device="cuda").manual_seed(1)
shape = (B, 263, 1, pred)
xT = torch.randn(*shape, device="cuda", generator=g)
tape = torch.randn(steps, *shape, device="cuda", generator=g)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


if "--one-chunk" in sys.argv:                    # for an ncu launch list: two plain chunk loops, nothing else
    for _ in range(2):
        diffusion.p_sample_loop(cfg, shape, noise=xT, clip_denoised=False, model_kwargs={"y": y}, noise_tape=tape)
    torch.cuda.synchronize()
    sys.exit(0)
chunk = timed(lambda: diffusion.p_sample_loop(cfg, shape, noise=xT, clip_denoised=False, model_kwargs={"y": y}, noise_tape=tape))
sampler = b200mdm.AutoRegressiveSampler(args, diffusion.p_sample_loop, required_frames=196)
n5 = torch.stack([xT] * 5)
t5 = torch.stack([tape] * 5)
full = timed(lambda: sampler.sample(cfg, (B, 263, 1, 196), clip_denoised=False, model_kwargs={"y": y}, noise=n5, noise_tape=t5), reps=5)
print("DiP B=%d: %.2f ms per 40-frame chunk (10 steps, CFG 7.5) = %.3f ms per step; 196-frame motions (5 chunks): %.1f ms per batch "
      "-> %.0f motions/s" % (B, chunk, chunk / steps, full, B / full * 1e3))
