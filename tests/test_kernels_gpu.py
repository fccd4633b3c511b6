"""GPU: each hand-written kernel against a plain torch fp32 restatement of the same op (through the C ABI)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from b200mdm import _lib as L
    return L, L.load()


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("M,N,K,bn,act", [
    (128, 256, 64, 128, 0),          # two N tiles, one k-block
    (128, 256, 512, 128, 0),         # k pipeline wraps the ring
    (300, 512, 512, 128, 0),         # M tail (300 = 2*128 + 44), 4 N tiles
    (25216 // 8, 1536, 512, 128, 0),  # QKV shape (M scaled down), many tiles per CTA -> both accumulator stages
    (1000, 1024, 512, 128, 1),       # FFN up + exact GELU
    (777, 512, 1024, 128, 0),        # FFN down shape
    (394, 512, 792, 128, 0),         # embed GEMM: K = 3*264 (K tail: 792 = 12*64 + 24), BLOCK_N 128
    (394, 288, 1536, 128, 0),        # N tail inside a 128-wide tile, 64-column slab clipped by the TMA store
    (394, 264, 1536, 128, 0),        # N tail that ends inside a 32-column chunk
    (5, 16, 8, 128, 1),              # tiny
    (256, 256, 64, 512, 0),          # CTA-pair kernel: one 256x256 tile, one k-block
    (700, 512, 512, 512, 0),         # pair: M tail inside the second CTA (700 = 2*256 + 188), pipeline wraps
    (25216 // 8, 1536, 512, 512, 0), # pair: QKV shape
    (3000, 1024, 512, 512, 1),       # pair: FFN up + GELU
    (130, 512, 1024, 512, 0),        # pair: second CTA holds 2 live rows only
    (100, 264, 512, 512, 0),         # pair: second CTA entirely out of range, N tail
    (256, 256, 64, 513, 0),          # W-resident pair kernel (gemm2w.cuh): one tile, one resident k-block
    (700, 512, 512, 513, 0),         # W-resident: M tail inside the second CTA, one tile per cluster
    (25216 // 8, 1536, 512, 513, 0), # W-resident: QKV shape, 6 column blocks
    (25216, 1024, 512, 513, 1),      # W-resident: the FFN up-projection at BASELINE config 2 (6 rounds per cluster: A ring
                                     # and both accumulator stages wrap while W stays put) + GELU
    (40000, 512, 320, 513, 0),       # W-resident: 5 k-blocks against a 4-slot A ring, 5 rounds
    (5000, 768, 200, 513, 1),        # W-resident: K tail inside the last k-block (200 = 3*64 + 8)
    (100, 264, 512, 513, 0),         # W-resident: second CTA entirely out of range, N tail
])
def test_gemm_tcgen05(M, N, K, bn, act):
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda", generator=g)).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
    L.check(lib.b200mdm_test_gemm_f16(_p(a), _p(w), _p(bias), _p(out), M, N, K, act, bn, _stream()))
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    if act:
        ref = torch.nn.functional.gelu(ref)
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all()
    assert err < 4e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("n,S,Mt,ld_extra", [(5, 60, 16, 0), (3, 60, 16, 7168), (2, 17, 5, 0), (3, 100, 24, 0), (2, 64, 33, 1024),
                                             (2, 61, 64, 0), (1, 1, 1, 0)])
def test_cross_attention(n, S, Mt, ld_extra):
    """The trans_dec cross-attention core (mma.sync tiles, P = hi + lo) against torch fp32: padding masks, token counts
    that do not fill a key tile, rows that do not fill a 16-row tile, k | v embedded in a wider row (all-layer projection)."""
    L, lib = _lib()
    d, H, dh = 512, 4, 128
    g = torch.Generator(device="cuda").manual_seed(100 * S + Mt)
    q = torch.randn(n * S, d, device="cuda", generator=g).half()
    ld = 2 * d + ld_extra
    kvw = torch.randn(n * Mt, ld, device="cuda", generator=g).half()
    col0 = ld_extra // 2 if ld_extra else 0                          # this "layer"'s k | v columns inside the wide row
    col0 -= col0 % 8
    mask = torch.rand(n, Mt, device="cuda", generator=g) < 0.3
    mask[:, 0] = False                                               # the CLS token is never padding
    out = torch.full((n * S, 2 * d), float("nan"), device="cuda", dtype=torch.float16)
    kv_ptr = kvw.data_ptr() + 2 * col0
    L.check(lib.b200mdm_test_cross_attention(_p(q), ctypes.c_void_p(kv_ptr), _p(mask.to(torch.uint8)), _p(out), n, S, Mt, ld,
                                             _stream()))
    torch.cuda.synchronize()
    k = kvw[:, col0:col0 + d].float().view(n, Mt, H, dh).permute(0, 2, 1, 3)
    v = kvw[:, col0 + d:col0 + 2 * d].float().view(n, Mt, H, dh).permute(0, 2, 1, 3)
    qq = q.float().view(n, S, H, dh).permute(0, 2, 1, 3)
    s = qq @ k.transpose(-1, -2) / dh ** 0.5
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(n * S, d)
    got = out[:, :d].float()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err < 1.5e-3 * max(1.0, ref.abs().max().item()), err      # fp16 output rounding (2^-11 relative) dominates


@pytest.mark.parametrize("impl", [0])
@pytest.mark.parametrize("n,S,kv", [(3, 197, [197, 121, 58]), (2, 41, [41, 1]), (4, 61, [61, 46, 31, 2]), (1, 16, [16]),
                                    (2, 33, [20, 33]), (2, 256, [256, 130]), (2, 129, [129, 128])])
def test_attention(n, S, kv, impl):
    L, lib = _lib()
    d, H, dh = 512, 4, 128
    g = torch.Generator(device="cuda").manual_seed(S)
    qkv = torch.randn(n * S, 3 * d, device="cuda", generator=g).half()
    kvlen = torch.tensor(kv, device="cuda", dtype=torch.int32)
    out = torch.full((n * S, d), float("nan"), device="cuda", dtype=torch.float16)
    L.check(lib.b200mdm_test_attention(_p(qkv), _p(out), _p(kvlen), n, S, d, impl, _stream()))
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(n, S, 3, H, dh).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / dh ** 0.5
    mask = torch.arange(S, device="cuda")[None, :] >= kvlen[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(n * S, d)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("n,S,kv,ld", [(3, 197, [197, 121, 58], 1024), (2, 41, [41, 1], 512), (4, 61, [61, 46, 31, 2], 512),
                                       (1, 16, [16], 512), (2, 256, [256, 130], 1024), (2, 129, [129, 128], 512),
                                       (2, 128, [128, 7], 512), (40, 197, [197] * 39 + [3], 1024)])
def test_qkv_attention_fused(n, S, kv, ld):
    """The fused QKV-projection + attention kernel (one CTA pair per (sample, head)) vs torch fp32 on the same fp16
    operands; `ld` = 1024 feeds it the hi half of an [hi | lo] residual stream like the engine does; 40 samples = 160
    items over 74 clusters exercises the persistent loop (2-3 items per cluster, all barrier phases wrap)."""
    L, lib = _lib()
    d, H, dh = 512, 4, 128
    g = torch.Generator(device="cuda").manual_seed(S * 7 + n)
    h = torch.randn(n * S, ld, device="cuda", generator=g).half()
    w = (torch.randn(3 * d, d, device="cuda", generator=g) / d ** 0.5).half()
    bias = torch.randn(3 * d, device="cuda", generator=g) * 0.3
    kvlen = torch.tensor(kv, device="cuda", dtype=torch.int32)
    out = torch.full((n * S, d), float("nan"), device="cuda", dtype=torch.float16)
    L.check(lib.b200mdm_test_qkv_attention(_p(h), ld, _p(w), _p(bias), _p(out), _p(kvlen), n, S, _stream()))
    torch.cuda.synchronize()
    qkv = (h[:, :d].float() @ w.float().t() + bias).half().float()           # the kernel keeps q, k, v in fp16
    q, k, v = qkv.view(n, S, 3, H, dh).permute(2, 0, 3, 1, 4)
    s_ = q @ k.transpose(-1, -2) / dh ** 0.5
    mask = torch.arange(S, device="cuda")[None, :] >= kvlen[:, None]
    s_ = s_.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s_, -1) @ v).permute(0, 2, 1, 3).reshape(n * S, d)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err < 5e-3, err


def _split_hi_lo(x):
    hi = x.half()
    return torch.cat([hi, (x - hi.float()).half()], dim=1).contiguous()


@pytest.mark.parametrize("M,K", [(256, 512), (25216 // 4, 512), (3000, 1024), (130, 512), (77, 1024), (25216, 512), (25216, 2048)])
def test_gemm_residual_layernorm_fused(M, K):
    """h <- LN(h + A W^T + b): the fused out-projection / FFN-down kernel vs torch fp32.  h travels as the engine's
    residual-stream format, fp16 [hi | lo] (hi + lo ~ 22 bits)."""
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(512, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(512, device="cuda", generator=g) * 0.1
    gamma = 1 + 0.1 * torch.randn(512, device="cuda", generator=g)
    beta = 0.1 * torch.randn(512, device="cuda", generator=g)
    h = torch.randn(M, 512, device="cuda", generator=g) * 1.5 + 0.2
    hres = _split_hi_lo(h)
    h_in = hres[:, :512].float() + hres[:, 512:].float()        # what the kernel reads (|h_in - h| < 1e-6)
    ref = torch.nn.functional.layer_norm(h_in + a.float() @ w.float().t() + bias, (512,), gamma, beta, 1e-5)
    L.check(lib.b200mdm_test_gemm_resid_ln(_p(a), _p(w), _p(bias), _p(gamma), _p(beta), _p(hres), M, K, _stream()))
    torch.cuda.synchronize()
    out = hres[:, :512].float() + hres[:, 512:].float()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-4
    assert (hres[:, :512].float() - ref).abs().max().item() < 5e-3   # the hi half alone is the fp16 GEMM operand
