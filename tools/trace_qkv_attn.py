"""clock64 phase stamps of the fused QKV + attention kernel (cluster 0, items 1 and 2) at the benchmark shape, plus
CUDA-event timing of the kernel alone and of the unfused pair it replaces.   python tools/trace_qkv_attn.py"""
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
h = torch.randn(n * S, 1024, device="cuda").half()
w = (torch.randn(3 * d, d, device="cuda") / d ** 0.5).half()
bias = torch.zeros(3 * d, device="cuda")
out = torch.empty(n * S, d, device="cuda", dtype=torch.float16)
kv = torch.full((n,), S, device="cuda", dtype=torch.int32)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
call = lambda: _lib.check(lib.b200mdm_test_qkv_attention(h.data_ptr(), 1024, w.data_ptr(), bias.data_ptr(), out.data_ptr(), kv.data_ptr(), n, S, st))
for _ in range(3): call()
torch.cuda.synchronize()
lib.b200mdm_debug_trace(buf.data_ptr()); call(); torch.cuda.synchronize(); lib.b200mdm_debug_trace(None)
t = buf.cpu().tolist()
names = ["wait proj_done", "slab drain + bar + Q/K epilogue + arrive", "V epilogue (DSMEM stores)", "wait s_done", "S read + max",
         "exchange + exp + P store + fence + arrive", "wait o_done", "O normalise + slab + store issue"]
for it in range(2):
    b = it * 16
    print("epilogue warp, item %d:" % (it + 1))
    for i, nme in enumerate(names):
        print("   %-44s %6d cycles" % (nme, t[b + i + 1] - t[b + i]))
    print("   item total %d" % (t[b + 8] - t[b]))
mn = ["projection issue (8 k-blocks, waits on TMA)", "wait qkv_ready", "S issue + wait p_ready (+tmem_free)", "PV issue"]
for it in range(2):
    b = 32 + it * 8
    print("MMA thread, item %d:" % (it + 1))
    for i, nme in enumerate(mn):
        print("   %-44s %6d cycles" % (nme, t[b + i + 1] - t[b + i]))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)
us = timeit(call)
fl = n * (2.0 * S * 512 * 1536 + 4.0 * S * S * 512)
print("fused kernel: %.1f us per launch, %.0f TFLOP/s by the SURVEY 8d accounting (49.8 GFLOP)" % (us, fl / us / 1e6))
qkv = torch.empty(n * S, 3 * d, device="cuda", dtype=torch.float16)
a = h[:, :512].contiguous()
g = lambda: _lib.check(lib.b200mdm_test_gemm_f16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), qkv.data_ptr(), n * S, 3 * d, d, 0, 512, st))
at = lambda: _lib.check(lib.b200mdm_test_attention(qkv.data_ptr(), out.data_ptr(), kv.data_ptr(), n, S, d, 0, st))
print("unfused: QKV GEMM %.1f us + attention %.1f us" % (timeit(g), timeit(at)))
