"""Mainloop efficiency probe: K large so the epilogue is negligible."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200mdm import _lib
lib = _lib.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (M, N, K) in [(148 * 128, 512, 8192), (148 * 128, 256, 8192), (74 * 256, 256, 8192), (25216, 512, 4096)]:
    a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    b = torch.zeros(N, device="cuda"); o = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for bn in (512, 256, 128):
        call = lambda: _lib.check(lib.b200mdm_test_gemm_f16(a.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), M, N, K, 0, bn, st))
        for _ in range(2): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): call()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 200
        print("M=%d N=%d K=%d block_n=%d: %.1f us  %.0f TFLOP/s" % (M, N, K, bn, us, 2.0 * M * N * K / us / 1e6))
