import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200MDM_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "motion-diffusion-model_b200", "lib", "libb200mdm_trace.so"))   # -DB200_TRACE build: B200MDM_TRACE=1 python -m b200mdm.build
def _ensure_trace_lib():
    import importlib
    if not os.path.exists(os.environ["B200MDM_LIB"]):
        importlib.import_module("motion-diffusion-model_b200.build").build(trace=True)


_ensure_trace_lib()
from b200mdm import _lib
lib = _lib.load()
M = 128 * 197
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
lib.b200mdm_debug_trace.argtypes = [ctypes.c_void_p]
for K in (512, 1024):
    a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(512, K, device="cuda") / K ** 0.5).half()
    b = torch.randn(512, device="cuda"); g = torch.ones(512, device="cuda"); be = torch.zeros(512, device="cuda")
    hres = torch.randn(M, 1024, device="cuda").half(); hres[:, 512:] *= 1e-3
    call = lambda: _lib.check(lib.b200mdm_test_gemm_resid_ln(a.data_ptr(), w.data_ptr(), b.data_ptr(), g.data_ptr(), be.data_ptr(), hres.data_ptr(), M, K, st))
    for _ in range(3): call()
    torch.cuda.synchronize()
    buf.zero_(); lib.b200mdm_debug_trace(buf.data_ptr()); call(); torch.cuda.synchronize(); lib.b200mdm_debug_trace(None)
    t = buf.cpu().tolist()
    for it in range(2):
        r = t[it*16:it*16+11]
        print("K=%d tile %d: acc_full wait %d | pass 1 chunks %s | statistics exchange %d | pass 2 chunks %s | tile total %d"
              % (K, it + 1, r[1]-r[0], [r[i+1]-r[i] for i in range(1, 5)], r[6]-r[5], [r[i+1]-r[i] for i in range(6, 10)], r[10]-r[0]))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("K=%d: %.1f us per launch (L2 flushed), %.0f TFLOP/s" % (K, sum(ts) / len(ts), 2.0 * M * 512 * K / (sum(ts) / len(ts)) / 1e6))
