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
n, S, d = 128, 197, 512
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
qkv = torch.randn(n * S, 3 * d, device="cuda").half(); out = torch.empty(n * S, d, device="cuda", dtype=torch.float16)
kv = torch.full((n,), S, device="cuda", dtype=torch.int32)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
call = lambda: _lib.check(lib.b200mdm_test_attention(qkv.data_ptr(), out.data_ptr(), kv.data_ptr(), n, S, d, 0, st))
for _ in range(3): call()
torch.cuda.synchronize()
lib.b200mdm_debug_trace(buf.data_ptr()); call(); torch.cuda.synchronize(); lib.b200mdm_debug_trace(None)
t = buf.cpu().tolist()
names = ["wait Q,K + QK^T (bar_s)", "pass1 max", "pass2 exp + P store", "wait V + PV (bar_o)", "O normalise -> slabs", "TMA store + drain"]
for i, nme in enumerate(names): print("%-28s %6d cycles" % (nme, t[i+1]-t[i]))
print("total", t[6]-t[0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): call()
e1.record(); torch.cuda.synchronize()
print("%.1f us per launch" % (e0.elapsed_time(e1) * 100))
