"""Time the fused sampling loop at BASELINE config 2 (B=64, T=196, 50 steps, CFG 2.5) with CUDA events: ms per loop.
usage: python tools/time_loop.py [reps] [batch] [steps]"""
import os, sys
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200mdm


def default_args(**over):
    a = dict(dataset="humanml", unconstrained=False, latent_dim=512, layers=8, cond_mask_prob=0.1, arch="trans_enc",
             emb_trans_dec=False, text_encoder_type="clip", pos_embed_max_len=5000, mask_frames=True, pred_len=0,
             context_len=0, diffusion_steps=50, noise_schedule="cosine", sigma_small=True, lambda_vel=0.0,
             lambda_rcxyz=0.0, lambda_fc=0.0)
    a.update(over)
    return SimpleNamespace(**a)


reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
args = default_args(layers=8, diffusion_steps=steps)
model, diffusion = b200mdm.create_model_and_diffusion(args, SimpleNamespace(dataset=SimpleNamespace()))
b200mdm.load_model_wo_clip(model, b200mdm.synthetic_state_dict(num_layers=8, seed=0))
model.to("cuda").eval()
cfg = b200mdm.ClassifierFreeSampleModel(model)
g = torch.Generator(device="cuda").manual_seed(1)
shape = (B, 263, 1, 196)
xT = torch.randn(*shape, device="cuda", generator=g)
tape = torch.randn(steps, *shape, device="cuda", generator=g)
y = dict(mask=torch.ones(B, 1, 1, 196, dtype=torch.bool, device="cuda"), lengths=torch.full((B,), 196, device="cuda"),
         text_embed=torch.randn(1, B, 512, device="cuda", generator=g), scale=torch.full((B,), 2.5, device="cuda"))
run = lambda: diffusion.p_sample_loop(cfg, shape, noise=xT, clip_denoised=False, model_kwargs={"y": y}, noise_tape=tape)
for _ in range(3):
    out = run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print("PDL=%s B=%d steps=%d: loop ms min %.2f median %.2f  -> %.1f motions/s  checksum %.6f" % (
    os.environ.get("B200MDM_PDL", "1"), B, steps, ts[0], ts[len(ts) // 2], B / ts[len(ts) // 2] * 1e3, float(out.double().abs().mean())))
