#!/usr/bin/env python
"""bench.py -- motions/sec of the MDM sampling hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one complete sampling loop over one batch: BASELINE config 2 -- HumanML3D shapes, B=64 motions per
GPU, 196 frames x 263 features, 50 DDPM steps, classifier-free guidance 2.5 (cond/uncond packed to 128 sequences),
trans_enc L=8 d=512 -- synthetic weights / text embeddings / noise (no network for checkpoints).
  value : whole-job motions/s, inputs (x_T, 660 MB noise tape, text embedding) resident in HBM, CUDA events, max over ranks
  e2e   : the same metric through the public API call a user makes (diffusion.p_sample_loop(model, shape, model_kwargs)),
          conditioning copied from pinned host memory and the sample read back to the host inside the timed region;
          noise is drawn on the device by the API exactly as the reference does on a GPU
  roofline     : the dominant kernel (tcgen05 GEMM, FFN-up shape of this workload) timed alone with CUDA events
  cpu_baseline : the CPU restatement of the reference (oracle/, fp32 torch on all host cores) on a bounded sample
Multi-GPU: batch sharded, one NCCL broadcast of the text embedding per loop, nothing inside the loop ("weak" scaling:
64 motions per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, T, J, STEPS, L, D, FF, SCALE = 64, 196, 263, 50, 8, 512, 1024, 2.5


WORKLOAD = "HumanML3D text2motion B=64/GPU T=196 J=263 50 DDPM steps CFG 2.5 trans_enc L8 d512 ff1024 h4"


def flops_per_forward_sample(S=T + 1, d=D, ff=FF, layers=L, jf=J, t=T):
    """F_fwd = L*2S*(3d^2 + d^2 + 2*d*ff + 2*S*d) + 2*(2*T*JF*d)."""
    return layers * 2 * S * (3 * d * d + d * d + 2 * d * ff + 2 * S * d) + 2 * (2 * t * jf * d)


FLOP_PER_MOTION = flops_per_forward_sample() * 2 * STEPS   # two CFG forwards per step


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"],
                    source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, False, None

    # NVML in-process (20 ms period: a 0.4 s timed region still gets ~20 samples); nvidia-smi -lms as the fallback
    def _nvml_open(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nvml, self.handle = pynvml, h
            return True
        except Exception:
            return False

    def _nvml_poll(self):
        n = self.nvml
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                bits = int(get(self.handle))
                row = ["", str(sm), str(mx), "", ""] + ["Active" if bits & b else "Not Active" for b in (0x8, 0x40, 0x20, 0x4)]
                self.rows.append(row)
            except Exception:
                pass
            time.sleep(0.02)
        del names

    def start(self):
        if self._nvml_open():
            self.thread = threading.Thread(target=self._nvml_poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            time.sleep(0.25)
            self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows),
                "source": "nvml 20 ms" if self.thread is not None else "nvidia-smi -lms 200"}


def make_args():
    from types import SimpleNamespace
    return SimpleNamespace(dataset="humanml", unconstrained=False, latent_dim=D, layers=L, cond_mask_prob=0.1,
                           arch="trans_enc", emb_trans_dec=False, text_encoder_type="clip", pos_embed_max_len=5000,
                           mask_frames=True, pred_len=0, context_len=0, diffusion_steps=STEPS, noise_schedule="cosine",
                           sigma_small=True, lambda_vel=0.0, lambda_rcxyz=0.0, lambda_fc=0.0)


# ----------------------------------------------------------------------------------------------------- CPU arms
def host_threads():
    """Threads the CPU arm may use: the cgroup quota if there is one, else the affinity mask (capped at 64)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(per))))
    except Exception:
        pass
    return max(1, min(n, 64))


def cpu_motions_per_sec(batch, threads, budget_s):
    """Oracle port of the reference p_sample_loop (CFG 2.5, L=8, 196 frames) on `batch` motions.  Sampler steps are
    timed one by one (they all cost the same) until `budget_s` seconds are spent; motions/s is extrapolated to the
    full 50-step loop.  Returns (motions_per_sec, steps_timed, seconds)."""
    import torch
    import b200mdm
    from oracle import mdm_oracle as mo, schedule_oracle as so
    torch.set_num_threads(threads)
    W = mo.OracleWeights(b200mdm.synthetic_state_dict(num_layers=L, seed=0), L)
    inp = b200mdm.synthetic_inputs(batch, nframes=T, steps=STEPS, seed=10)
    tabs = so.diffusion_tables(so.named_betas("cosine", STEPS))
    x = inp["tape"][0].clone()
    times = []
    with torch.no_grad():
        for k, i in enumerate(range(STEPS - 1, -1, -1)):
            t0 = time.perf_counter()
            x0 = mo.cfg_denoise_enc(W, x, i, inp["text_embed"], inp["scale"], inp["lengths"])
            x, _ = mo.p_sample_step(tabs, x0, x, i, inp["tape"][1 + k])
            times.append(time.perf_counter() - t0)
            if sum(times) > budget_s and len(times) >= 2:
                break
    steady = times[1:] if len(times) > 1 else times           # first step pays allocator / thread-pool warm-up
    per_step = sum(steady) / len(steady)
    return batch / (per_step * STEPS), len(times), sum(times)


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference_arm(a, rank, world):
    """--impl reference: the reference's CPU implementation of the path (its restatement in oracle/, since
    /root/reference does not exist on the GPU box), all host threads, rank 0 only."""
    if rank != 0:
        return
    threads = host_threads()
    sample_b = 4
    budget = 20.0
    vals, nsteps = [], 0
    for _ in range(max(1, a.steps)):
        v, nsteps, _sec = cpu_motions_per_sec(sample_b, threads, budget / max(1, a.steps))
        vals.append(v)
    val = sum(vals) / len(vals)
    sec = sample_b / val
    line = {"impl": "reference", "metric": "motions/sec", "value": round(val, 4), "unit": "motions/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": 64 * max(1, a.gpus),
                       "sample": "4 of 64 motions per measurement, CPU oracle port of the reference loop"},
            "cpu_baseline": {"value": round(val, 4), "unit": "motions/s", "cores": threads, "kind": "port",
                             "sample": "4 of 64 motions, %d of 50 sampler steps timed per measurement (x2 CFG forwards each), extrapolated to 50; torch fp32; %s" % (nsteps, cpu_model_name())},
            "e2e": {"value": round(val, 4), "unit": "motions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return run_reference_arm(a, rank, world)

    import torch
    import torch.distributed as dist
    import b200mdm
    assert torch.cuda.is_available(), "bench.py needs a B200"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from types import SimpleNamespace
    model, diffusion = b200mdm.create_model_and_diffusion(make_args(), SimpleNamespace(dataset=SimpleNamespace()))
    b200mdm.load_model_wo_clip(model, b200mdm.synthetic_state_dict(num_layers=L, seed=0))
    model = b200mdm.ClassifierFreeSampleModel(model.to(dev).eval())
    eng = model.model.engine()
    B = B_PER_GPU
    shape = (B, J, 1, T)
    inp = b200mdm.synthetic_inputs(B, nframes=T, steps=STEPS, seed=10 + rank)
    # resident inputs for `value`
    xT = inp["tape"][0].to(dev)
    tape = torch.stack(inp["tape"][1:]).to(dev).contiguous()           # 50 x 13.2 MB = 660 MB  (> 126 MB L2)
    text = torch.zeros(1, B * world, 512, device=dev)
    if rank == 0:
        g = torch.Generator().manual_seed(1234)
        text.copy_(torch.randn(1, B * world, 512, generator=g))
    lengths, mask, scale = inp["lengths"].to(dev), inp["mask"].to(dev), inp["scale"].to(dev)

    def one_loop_resident():
        if world > 1:
            dist.broadcast(text, src=0)                                  # the one collective of the path
        y = dict(mask=mask, lengths=lengths, text_embed=text[:, rank * B:(rank + 1) * B], scale=scale)
        return diffusion.p_sample_loop(model, shape, noise=xT, clip_denoised=False, model_kwargs={"y": y}, noise_tape=tape)

    # pinned host buffers for `e2e`
    text_h = torch.randn(1, B, 512).pin_memory()
    scale_h = torch.full((B,), SCALE).pin_memory()
    lengths_h = torch.full((B,), T, dtype=torch.int64).pin_memory()
    out_h = torch.empty(shape, dtype=torch.float32).pin_memory()

    def one_loop_e2e():
        te = text_h.to(dev, non_blocking=True)
        sc = scale_h.to(dev, non_blocking=True)
        ln = lengths_h.to(dev, non_blocking=True)
        if world > 1:
            dist.broadcast(te, src=0)
        y = dict(mask=mask, lengths=ln, text_embed=te, scale=sc)
        s = diffusion.p_sample_loop(model, shape, clip_denoised=False, model_kwargs={"y": y})
        out_h.copy_(s, non_blocking=True)
        return s

    def timed(fn, iters):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def log(msg):
        if rank == 0:
            print("[bench %.1fs] %s" % (time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)

    t_start = time.perf_counter()
    for _ in range(max(a.warmup, 3)):
        one_loop_resident()
    torch.cuda.synchronize()
    log("warm-up done")
    eng.launch_count(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(one_loop_resident, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count(reset=True)
    log("resident loops timed: %.2f ms per loop" % (ms_total / a.steps))
    for _ in range(2):
        one_loop_e2e()
    ms_e2e = timed(one_loop_e2e, a.steps)
    log("e2e loops timed: %.2f ms per loop" % (ms_e2e / a.steps))

    ms_step = ms_total / a.steps
    value = B * world / (ms_step * 1e-3)
    e2e_value = B * world / (ms_e2e / a.steps * 1e-3)
    peaks = measured_peaks()

    line = None
    if rank == 0:
        # ---- the step's kernels timed alone at this workload's shapes (M = 128 sequences x 197 tokens), through the
        # kernel-level C-ABI hooks, CUDA events on the launching stream, L2 flushed between launches.  "roofline" is
        # the kernel with the largest share of the step (profiles/: gemm_resid_ln_cluster, FFN-down shape).
        import ctypes
        from b200mdm import _lib
        lib = _lib.load()
        M = 2 * B * (T + 1)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

        def time_kernel(call):
            for _ in range(3):
                call()
            ts = []
            for _ in range(10):
                flush.zero_()                                          # L2 flush between timed launches
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); call(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sum(ts) / len(ts)

        def gemm_case(N, K, act):
            A = torch.randn(M, K, device=dev).half()
            Wt = (torch.randn(N, K, device=dev) / K ** 0.5).half()
            bias = torch.zeros(N, device=dev)
            O = torch.empty(M, N, device=dev, dtype=torch.float16)
            ms = time_kernel(lambda: _lib.check(lib.b200mdm_test_gemm_f16(A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), O.data_ptr(), M, N, K, act, 512, st)))
            return 2.0 * M * N * K / (ms * 1e-3) / 1e12, ms

        def ln_case(K):
            A = torch.randn(M, K, device=dev).half()
            Wt = (torch.randn(512, K, device=dev) / K ** 0.5).half()
            vec = [torch.zeros(512, device=dev), torch.ones(512, device=dev), torch.zeros(512, device=dev)]
            hres = torch.randn(M, 1024, device=dev).half()      # residual stream, fp16 [hi | lo]
            hres[:, 512:] *= 1e-3
            ms = time_kernel(lambda: _lib.check(lib.b200mdm_test_gemm_resid_ln(A.data_ptr(), Wt.data_ptr(), vec[0].data_ptr(), vec[1].data_ptr(), vec[2].data_ptr(), hres.data_ptr(), M, K, st)))
            return 2.0 * M * 512 * K / (ms * 1e-3) / 1e12, ms

        k_tflops, k_ms = ln_case(FF)
        others = []
        for name, fn in (("gemm_resid_ln_cluster out-proj K=512", lambda: ln_case(D)),
                         ("gemm2_f16_tcgen05<bias> QKV N=1536 K=512", lambda: gemm_case(3 * D, D, 0)),
                         ("gemm2_f16_tcgen05<bias,gelu> FFN-up N=1024 K=512", lambda: gemm_case(FF, D, 1))):
            tf, ms_k = fn()
            others.append({"kernel": name, "achieved": round(tf, 1), "frac": round(tf / peaks["burst"], 4), "us": round(ms_k * 1e3, 1)})
        path_tflops = value / world * FLOP_PER_MOTION / 1e12            # per GPU
        roof = {"bound": "tensor", "kernel": "gemm_resid_ln_cluster (FFN-down + residual + LayerNorm) M=%d N=512 K=%d" % (M, FF),
                "achieved": round(k_tflops, 1), "peak": peaks["burst"], "unit": "TFLOP/s",
                "frac": round(k_tflops / peaks["burst"], 4),
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the ncu --set full
                # capture committed as profiles/r01_c_top_kernel_ncu_metrics.txt (104.4 MB read + 16.6 MB written; the
                # algorithmic bytes are 155.9 MB = A 51.6 + residual in 51.6 + out 51.6 + W 1.0: part of the residual
                # stream stays in the persisting L2 window)
                "traffic": 121.0e6, "us": round(k_ms * 1e3, 1),
                "peak_source": peaks["source"] + " bf16 burst (cuBLAS 8192^3)", "other_kernels": others,
                "path_achieved_tflops_per_gpu": round(path_tflops, 1),
                "path_frac_of_sustained": round(path_tflops / peaks["sustained"], 4)}
        log("kernel roofline timed: %.1f TFLOP/s" % k_tflops)
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            threads = host_threads()
            sb = 4
            v, nst, sec = cpu_motions_per_sec(sb, threads, 15.0)
            cpu = {"value": round(v, 4), "unit": "motions/s", "cores": threads, "kind": "port",
                   "sample": "oracle port of the reference loop: %d of 64 motions, %d of 50 sampler steps timed (%.1f s, x2 CFG "
                             "forwards each), extrapolated to 50 steps; torch fp32; %s" % (sb, nst, sec, cpu_model_name())}
        line = {"metric": "motions/sec", "value": round(value, 2), "unit": "motions/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp16 operands / fp32 accumulate (in/out projections hi-lo split)",
                "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "global_batch": B * world, "parallelism": "batch-sharded x%d, 1 NCCL broadcast of text_embed per loop" % world,
                           "l2": "inputs larger than L2 (660 MB noise tape streamed per loop)", "cuda_graph": True},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": round(e2e_value, 2), "unit": "motions/s",
                        "h2d_bytes_per_step": int(text_h.numel() * 4 + scale_h.numel() * 4 + lengths_h.numel() * 8),
                        "d2h_bytes_per_step": int(out_h.numel() * 4)},
                "roofline": roof, "cpu_baseline": cpu,
                "flop_per_motion": FLOP_PER_MOTION}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
