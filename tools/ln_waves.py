"""Fused residual+LayerNorm GEMM: time per launch against the number of 128-row tiles (waves of 74 clusters)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200mdm import _lib
lib = _lib.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for K in (512, 1024):
    for tiles in (37, 74, 111, 148, 197, 222, 296):
        M = 128 * tiles
        a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(512, K, device="cuda") / K ** 0.5).half()
        b = torch.randn(512, device="cuda"); g = torch.ones(512, device="cuda"); be = torch.zeros(512, device="cuda")
        hres = torch.randn(M, 1024, device="cuda").half(); hres[:, 512:] *= 1e-3
        call = lambda: _lib.check(lib.b200mdm_test_gemm_resid_ln(a.data_ptr(), w.data_ptr(), b.data_ptr(), g.data_ptr(), be.data_ptr(), hres.data_ptr(), M, K, st))
        for _ in range(3): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): call()
        e1.record(); torch.cuda.synchronize()
        print("K=%d tiles=%d (%.2f waves): %.1f us" % (K, tiles, tiles / 74, e0.elapsed_time(e1) * 50))
