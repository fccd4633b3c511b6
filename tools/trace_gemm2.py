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
lib.b200mdm_debug_trace.argtypes = [ctypes.c_void_p]
M = 128 * 197
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
for (N, K, act) in [(1536, 512, 0), (512, 1024, 0), (1024, 512, 1)]:
    a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    b = torch.zeros(N, device="cuda"); o = torch.empty(M, N, device="cuda", dtype=torch.float16)
    call = lambda: _lib.check(lib.b200mdm_test_gemm_f16(a.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), M, N, K, act, 512, st))
    for _ in range(3): call()
    torch.cuda.synchronize()
    buf.zero_(); lib.b200mdm_debug_trace(buf.data_ptr()); call(); torch.cuda.synchronize(); lib.b200mdm_debug_trace(None)
    t = buf.cpu().tolist(); t0 = t[0]
    print("N=%d K=%d act=%d   (cycles, relative to the MMA warp's first stamp)" % (N, K, act))
    for it in range(6):
        r = t[it*8:it*8+8]
        print("  tile %d: mma start %6d  acc_empty wait %5d  mma issue span %6d | epi: enter %6d  acc_full wait %6d  epilogue %6d  done at %6d" % (
            it, r[0]-t0, r[1]-r[0], r[2]-r[1], r[4]-t0, r[5]-r[4], r[6]-r[5], r[6]-t0))
    for it in range(2, 6):
        r = t[64 + (it-2)*16: 64 + (it-2)*16 + 8]; base = t[it*8+5]
        print("  tile %d chunks (start,end rel. to acc_full): " % it + "  ".join("[%d..%d]" % (r[2*i]-base, r[2*i+1]-base) for i in range(4)))
