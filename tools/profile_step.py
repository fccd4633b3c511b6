"""Small driver for ncu: BASELINE config 2 shapes (B=64, T=196, L=8, CFG), a few sampler steps, no CUDA graph so
that every kernel is a plain launch.   ncu ... python tools/profile_step.py [steps]"""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200mdm  # noqa: E402
from bench import make_args, B_PER_GPU, T, J  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
args = make_args()
args.diffusion_steps = steps
model, diffusion = b200mdm.create_model_and_diffusion(args, SimpleNamespace(dataset=SimpleNamespace()))
b200mdm.load_model_wo_clip(model, b200mdm.synthetic_state_dict(num_layers=8, seed=0))
model = b200mdm.ClassifierFreeSampleModel(model.cuda().eval())
inp = b200mdm.synthetic_inputs(B_PER_GPU, nframes=T, steps=steps, seed=10)
y = dict(mask=inp["mask"].cuda(), lengths=inp["lengths"].cuda(), text_embed=inp["text_embed"].cuda(), scale=inp["scale"].cuda())
xT, tape = inp["tape"][0].cuda(), torch.stack(inp["tape"][1:]).cuda()
for _ in range(2):
    out = diffusion.p_sample_loop(model, (B_PER_GPU, J, 1, T), noise=xT, clip_denoised=False, model_kwargs={"y": y},
                                  noise_tape=tape, use_graph=False)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
