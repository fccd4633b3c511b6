"""Host mirror of the reference's data_loaders/humanml/scripts/motion_process.py for the step that follows the sampling
loop: HumanML3D feature vectors -> joint positions, on the GPU.

`recover_from_ric(data, joints_num)` keeps the reference signature and layout (motion_process.py:437-452): data
[..., T, 263|251] de-normalised features -> [..., T, joints_num, 3].  `sample_to_xyz(sample, mean, std)` is the three
lines of sample/generate.py:161-166 in one kernel launch: model output [B, D, 1, T] (normalised, still on the GPU) ->
[B, n_joints, 3, T]; no `sample.cpu()` round trip.  All arithmetic is in libb200mdm.so (csrc/postprocess.cuh); there is
no CPU fallback.
"""
import ctypes

import torch

from .... import _lib
from ...._lib import check


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("b200mdm post-processing runs on the GPU only (tensor is on %s); there is no CPU fallback" % t.device)


def recover_from_ric(data, joints_num):
    """reference motion_process.py:437-452.  data [..., T, D] fp32 on cuda -> [..., T, joints_num, 3]."""
    _need_cuda(data)
    lead = data.shape[:-2]
    T, D = data.shape[-2], data.shape[-1]
    assert D >= 4 + 3 * (joints_num - 1), (D, joints_num)
    x = data.to(torch.float32).reshape(-1, T, D).contiguous()
    out = torch.empty(x.shape[0], T, joints_num, 3, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.load().b200mdm_recover_from_ric(_ptr(x), T * D, 1, D, None, None, _ptr(out), T * joints_num * 3,
                                                   joints_num * 3, 1, x.shape[0], T, joints_num, _stream()))
    return out.reshape(*lead, T, joints_num, 3)


def sample_to_xyz(sample, mean, std):
    """sample/generate.py:161-166: `inv_transform(sample.cpu().permute(0,2,3,1))`, `recover_from_ric`, `permute(0,2,3,1)`.
    sample [B, D, 1, T] on cuda; mean / std: the dataset's Mean.npy / Std.npy ([D], array or tensor).  -> [B, J, 3, T]."""
    _need_cuda(sample)
    B, D, F, T = sample.shape
    assert F == 1 and D in (263, 251), sample.shape
    joints = 22 if D == 263 else 21
    x = sample.to(torch.float32).contiguous()
    m = torch.as_tensor(mean, dtype=torch.float32).to(x.device).contiguous()
    s = torch.as_tensor(std, dtype=torch.float32).to(x.device).contiguous()
    assert m.numel() == D and s.numel() == D
    out = torch.empty(B, joints, 3, T, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.load().b200mdm_recover_from_ric(_ptr(x), D * T, T, 1, _ptr(m), _ptr(s), _ptr(out), joints * 3 * T, 1, T,
                                                   B, T, joints, _stream()))
    return out
