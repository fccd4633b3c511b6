"""Runs the layer GEMMs of BASELINE config 2 (M = 128 seq x 197 tok) through the kernel-level C-ABI hook, for ncu.
usage: python tools/profile_gemm.py [block_n ...]   (512 = CTA-pair kernel, 256 / 128 = single-CTA kernel)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200mdm import _lib  # noqa: E402

lib = _lib.load()
M = 128 * 197
bns = [int(a) for a in sys.argv[1:]] or [512, 256]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (N, K, act) in [(1536, 512, 0), (1024, 512, 1), (512, 1024, 0)]:
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    b = torch.randn(N, device="cuda")
    o = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for bn in bns:
        for _ in range(3):
            _lib.check(lib.b200mdm_test_gemm_f16(a.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), M, N, K, act, bn, st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.check(lib.b200mdm_test_gemm_f16(a.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), M, N, K, act, bn, st))
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 20
        print("N=%d K=%d act=%d block_n=%d: %.1f us  %.0f TFLOP/s" % (N, K, act, bn, us, 2.0 * M * N * K / us / 1e6))
