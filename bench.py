#!/usr/bin/env python
"""bench.py — output tok/s (+ p50 TTFT / TPOT) of the concurrent prefill + decode path.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`; for N > 1 it
is launched under torch.distributed.run, one rank per GPU.  Rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1]): Llama-3-8B, TP = 1, semi-PD with a 50/50 SM split, synthetic
requests of 1024 input / 128 output tokens (token ids as python/sglang/bench_serving.py:772-783,
np.random.seed(1); ignore_eos), random-init weights of that architecture.
A STEP = one serving window: R requests pushed through the engine (prefill worker and decode worker
running concurrently on disjoint SM partitions) until the last token of the last request.
  * `value`  : output tok/s with the prompts already resident in HBM and sampled ids kept on the
               device (no per-step host copies) — the saturated (arrival rate = inf) throughput;
  * `e2e`    : the same windows served from HOST buffers through the public engine API
               (SemiPDEngine.submit / wait): prompt ids H2D from pinned memory every prefill batch,
               sampled ids D2H every decode step;
  * `latency`: a Poisson window at a finite rate (exponential gaps, bench_serving.py:884-899) for
               p50/p99 TTFT and TPOT (TPOT = (t_last - t_first)/(n-1), bench_serving.py:920-921).
N > 1 runs N independent replicas (one engine per GPU, no data-path collective): weak scaling.
`--impl reference` times the CPU restatement of the reference forward (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "output_tok_per_s_llama3_8b_concurrent_prefill_decode"
UNIT = "tok/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="spd", choices=["spd", "reference"])
    ap.add_argument("--model", type=str, default="llama-3-8b")
    ap.add_argument("--dtype", type=str, default="bfloat16")
    ap.add_argument("--requests", type=int, default=192, help="requests per step (serving window)")
    ap.add_argument("--input-len", type=int, default=1024)
    ap.add_argument("--output-len", type=int, default=128)
    ap.add_argument("--prefill-sm-pct", type=int, default=50)
    ap.add_argument("--decode-sm-pct", type=int, default=50)
    ap.add_argument("--partition", type=str, default="green", choices=["green", "none"])
    ap.add_argument("--latency-rate", type=float, default=0.0,
                    help="req/s of the Poisson latency window (0 = 60%% of the measured capacity)")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-layers", type=int, default=2, help="layers in the CPU baseline sample")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def make_prompts(n, input_len, vocab, seed=1):
    rng = np.random.RandomState(seed)
    offsets = rng.randint(0, vocab, size=n)
    return [[int((offsets[i] + i + j) % vocab) for j in range(input_len)] for i in range(n)]


def poisson_arrivals(n, rate, seed=1):
    if not math.isfinite(rate) or rate <= 0:
        return [0.0] * n
    rng = np.random.RandomState(seed)
    return np.cumsum(rng.exponential(1.0 / rate, size=n)).tolist()


def run_window(engine, prompts, dev_prompts, output_len, rate, device_resident):
    """Serve one window; returns (wall seconds, finished Req list)."""
    engine.set_device_resident(device_resident)
    arrivals = poisson_arrivals(len(prompts), rate)
    t0 = time.perf_counter()
    for i, p in enumerate(prompts):
        wait = t0 + arrivals[i] - time.perf_counter()
        if wait > 0:
            time.sleep(wait)
        engine.submit(p, output_len, arrival_time=t0 + arrivals[i],
                      input_ids_device=dev_prompts[i] if device_resident else None)
    reqs = engine.wait(len(prompts))
    torch.cuda.synchronize()
    return time.perf_counter() - t0, reqs


def latency_stats(reqs):
    ttft = np.array([r.first_token_time - r.arrival_time for r in reqs]) * 1e3
    tpot = np.array([(r.token_times[-1] - r.token_times[0]) / max(len(r.token_times) - 1, 1)
                     for r in reqs if len(r.token_times) > 1]) * 1e3
    q = lambda a, p: float(np.percentile(a, p)) if len(a) else None  # noqa: E731
    return {"ttft_ms_p50": q(ttft, 50), "ttft_ms_p99": q(ttft, 99), "tpot_ms_p50": q(tpot, 50),
            "tpot_ms_p99": q(tpot, 99)}


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# roofline of the dominant kernels (measured live, CUDA events on the launching stream)
# ----------------------------------------------------------------------------------------------
def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return pk["hbm_gbs"], pk["bf16_tflops"], pk["bf16_tflops_sustained"], "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, 1590.0, 1400.0, "fallback"


def time_on_stream(fn, stream, iters=10, warm=3, flush=None, burst=1):
    """Average launch duration, CUDA events on the launching stream.  `burst` launches are timed
    between one event pair (for kernels of ~100 us, whose working set exceeds L2 by itself, the
    event / launch gap of a single timed launch is several percent of the kernel)."""
    with torch.cuda.stream(stream):
        for _ in range(warm):
            fn()
        total = 0.0
        for _ in range(iters):
            if flush is not None:
                flush.add_(1)  # > L2-sized write between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(burst):
                fn()
            e1.record(stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
    return total / iters / burst


def roofline_section(engine, cfg, dtype, tokens_per_prefill_batch, decode_bs, ctx):
    from semi_pd_b200 import ops
    hbm, tf_burst, tf_sust, src = load_peaks()
    dev = engine.d_worker.device
    out = {}
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    H, I = cfg.hidden_size, cfg.intermediate_size
    layer = engine.d_worker.model_runner.model.model.layers[0]
    # (1) dominant kernel of the window: the gate_up GEMM (+ fused SiLU*mul) of a prefill batch
    M = tokens_per_prefill_batch
    x = (torch.randn(M, H, device=dev) * 0.5).to(dtype)
    w = layer.mlp.gate_up_proj.weight
    flops = 2.0 * M * (2 * I) * H
    n_sms = ops.device_sm_count(dev.index)
    share = engine.p_sms / float(n_sms)
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tr = json.load(f)
        if tr.get("M") == M:
            traffic = tr.get("gemm_gate_up_silu_dram_bytes")
    except Exception:  # noqa: BLE001
        pass
    # timed ALONE on the whole device -> compared with the measured burst peak of the whole device
    ops.set_sm_budget(0)
    full_stream = torch.cuda.Stream(device=dev)
    ms_full = time_on_stream(lambda: ops.linear(x, w, epilogue=ops.SPD_EPI_SILU_MUL), full_stream, flush=flush)
    out["roofline"] = {"bound": "tensor",
                       "kernel": "gemm_pair_kernel<SILU_MUL> (W_gate_up + fused SiLU*mul, cta_group::2 256x256 tiles)",
                       "achieved": flops / ms_full / 1e9, "peak": tf_burst, "unit": "TFLOP/s",
                       "frac": flops / ms_full / 1e9 / tf_burst, "traffic": traffic,
                       "peak_source": f"{src} burst bf16 {tf_burst} TF/s (kernel timed alone on all {n_sms} SMs)",
                       "launch_ms": ms_full, "M": M, "N": 2 * I, "K": H}
    extra = {}
    # the same kernel where the step runs it: on the prefill partition.  A 72-SM partition is not
    # power-capped the way the whole chip is (SM clock ~1.75-1.95 GHz instead of ~1.2-1.45), so its
    # per-SM rate exceeds the whole-chip measured peak per SM; the ceiling quoted here is therefore
    # the clock-based one: 8192 dense bf16 FLOP/clk/SM x max SM clock x SMs of the partition.
    ops.set_sm_budget(engine.p_sms)
    ms = time_on_stream(lambda: ops.linear(x, w, epilogue=ops.SPD_EPI_SILU_MUL), engine.p_stream, flush=flush)
    try:
        import pynvml
        pynvml.nvmlInit()
        max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(pynvml.nvmlDeviceGetHandleByIndex(dev.index), pynvml.NVML_CLOCK_SM)
    except Exception:  # noqa: BLE001
        max_mhz = 1965
    part_peak = 8192.0 * max_mhz * 1e6 * engine.p_sms / 1e12
    extra["gemm_gate_up_on_prefill_partition"] = {
        "bound": "tensor", "achieved": flops / ms / 1e9, "peak": part_peak, "unit": "TFLOP/s",
        "frac": flops / ms / 1e9 / part_peak, "launch_ms": ms,
        "peak_source": f"8192 FLOP/clk/SM x {max_mhz} MHz x {engine.p_sms} SMs (partition is not power-capped)",
        "frac_of_measured_share": flops / ms / 1e9 / (tf_burst * share)}
    # (2) prefill attention (tensor bound)
    Hq, Hk, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    nreq = max(M // 1024, 1)
    E = M // nreq
    slots = nreq * E + 8
    kp = torch.randn(slots, Hk, D, device=dev).to(dtype)
    vp = torch.randn(slots, Hk, D, device=dev).to(dtype)
    rtt = (torch.arange(nreq * E, device=dev, dtype=torch.int32) + 1).view(nreq, E).contiguous()
    rpi = torch.arange(nreq, device=dev)
    pre = torch.zeros(nreq, dtype=torch.int32, device=dev)
    ext = torch.full((nreq,), E, dtype=torch.int32, device=dev)
    _, st = ops.compute_positions(pre, ext, nreq * E)
    q = torch.randn(nreq * E, Hq * D, device=dev).to(dtype)
    ms = time_on_stream(lambda: ops.attn_prefill(q, kp, vp, rtt, rpi, pre, ext, st, E, Hq, Hk, D, D ** -0.5),
                        engine.p_stream, flush=flush)
    fl = 4.0 * Hq * D * nreq * E * (E + 1) / 2
    extra["attn_prefill"] = {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": tf_burst * share,
                             "unit": "TFLOP/s", "frac": fl / ms / 1e9 / (tf_burst * share), "launch_ms": ms,
                             "shape": f"{nreq} x {E} tokens, prefix 0",
                             "note": f"on the prefill partition ({engine.p_sms} SMs); peak = measured burst x SM share"}
    # the same kernel alone on the whole device, at the bench shape and at a long-context shape
    ops.set_sm_budget(0)
    ms = time_on_stream(lambda: ops.attn_prefill(q, kp, vp, rtt, rpi, pre, ext, st, E, Hq, Hk, D, D ** -0.5),
                        full_stream, flush=flush)
    extra["attn_prefill_full_gpu"] = {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": tf_burst,
                                      "unit": "TFLOP/s", "frac": fl / ms / 1e9 / tf_burst, "launch_ms": ms,
                                      "shape": f"{nreq} x {E} tokens, prefix 0, all {n_sms} SMs"}
    E2 = 4096
    n2 = max(M // E2, 1)
    kp2 = torch.randn(n2 * E2 + 8, Hk, D, device=dev).to(dtype)
    vp2 = torch.randn(n2 * E2 + 8, Hk, D, device=dev).to(dtype)
    rtt2 = (torch.arange(n2 * E2, device=dev, dtype=torch.int32) + 1).view(n2, E2).contiguous()
    rpi2 = torch.arange(n2, device=dev)
    pre2 = torch.zeros(n2, dtype=torch.int32, device=dev)
    ext2 = torch.full((n2,), E2, dtype=torch.int32, device=dev)
    _, st2 = ops.compute_positions(pre2, ext2, n2 * E2)
    q2 = torch.randn(n2 * E2, Hq * D, device=dev).to(dtype)
    ms = time_on_stream(lambda: ops.attn_prefill(q2, kp2, vp2, rtt2, rpi2, pre2, ext2, st2, E2, Hq, Hk, D, D ** -0.5),
                        full_stream, flush=flush)
    fl2 = 4.0 * Hq * D * n2 * E2 * (E2 + 1) / 2
    extra["attn_prefill_full_gpu_4k"] = {"bound": "tensor", "achieved": fl2 / ms / 1e9, "peak": tf_burst,
                                         "unit": "TFLOP/s", "frac": fl2 / ms / 1e9 / tf_burst, "launch_ms": ms,
                                         "shape": f"{n2} x {E2} tokens, prefix 0, all {n_sms} SMs"}
    del kp2, vp2, q2
    ops.set_sm_budget(engine.p_sms)
    # (3) decode attention (HBM bound) on the decode partition
    ops.set_sm_budget(engine.d_sms)
    B, S = decode_bs, ctx
    slots = B * S + 8
    kp = torch.randn(slots, Hk, D, device=dev).to(dtype)
    vp = torch.randn(slots, Hk, D, device=dev).to(dtype)
    rtt = (torch.randperm(B * S, device=dev).int() + 1).view(B, S).contiguous()
    rpi = torch.arange(B, device=dev)
    sl = torch.full((B,), S, dtype=torch.int64, device=dev)
    qd = torch.randn(B, Hq * D, device=dev).to(dtype)
    ms = time_on_stream(lambda: ops.attn_decode(qd, kp, vp, rtt, rpi, sl, Hq, Hk, D, D ** -0.5),
                        engine.d_stream, iters=10, flush=flush, burst=4)
    by = B * S * Hk * D * 2 * 2 + 2 * B * Hq * D * 2 + 4 * B * S
    extra["attn_decode"] = {"bound": "hbm", "achieved": by / ms / 1e6, "peak": hbm, "unit": "GB/s",
                            "frac": by / ms / 1e6 / hbm, "launch_ms": ms, "shape": f"B={B} S={S}",
                            "note": f"decode partition = {engine.d_sms} SMs; peak is the full-device copy bandwidth; "
                                    "4 launches per event pair (each streams > 2 x L2 of K/V, no reuse)"}
    ops.set_sm_budget(0)
    ms = time_on_stream(lambda: ops.attn_decode(qd, kp, vp, rtt, rpi, sl, Hq, Hk, D, D ** -0.5),
                        full_stream, iters=10, flush=flush, burst=4)
    extra["attn_decode_full_gpu"] = {"bound": "hbm", "achieved": by / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                     "frac": by / ms / 1e6 / hbm, "launch_ms": ms,
                                     "shape": f"B={B} S={S}, all {n_sms} SMs",
                                     "note": "4 launches per event pair (each streams > 2 x L2 of K/V, no reuse)"}
    ops.set_sm_budget(engine.d_sms)
    # (4) decode-regime GEMM (weight streaming, HBM bound)
    xd = (torch.randn(B, H, device=dev) * 0.5).to(dtype)
    ms = time_on_stream(lambda: ops.linear(xd, w), engine.d_stream, iters=20, flush=flush)
    by = w.numel() * 2 + B * H * 2 + B * 2 * I * 2
    extra["gemm_decode_gate_up"] = {"bound": "hbm", "achieved": by / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                    "frac": by / ms / 1e6 / hbm, "launch_ms": ms, "shape": f"M={B} N={2*I} K={H}"}
    ops.set_sm_budget(0)
    out["roofline_extra"] = extra
    del flush
    return out


# ----------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference forward; bounded sample)
# ----------------------------------------------------------------------------------------------
def cpu_baseline(model_name, dtype, n_layers, input_len, output_len, weights=None, sample_prefill=256,
                 sample_decode=4):
    """Times the CPU restatement of the reference's forward (oracle/llama_oracle.py, torch eager,
    all host threads) on `n_layers` of the model's layers, a `sample_prefill`-token prefill and
    `sample_decode` batch-1 decode steps, and extrapolates linearly to the full depth and the
    1024-in/128-out request: t_req = t_prefill(1024) + 127 x t_decode_step."""
    from oracle import llama_oracle as O
    from semi_pd_b200.configs import get_model_config
    torch.set_num_threads(os.cpu_count() or 1)
    full = get_model_config(model_name)
    cfg = O.OracleLlamaConfig(hidden_size=full.hidden_size, intermediate_size=full.intermediate_size,
                              num_hidden_layers=n_layers, num_attention_heads=full.num_attention_heads,
                              num_key_value_heads=full.num_key_value_heads, head_dim=full.head_dim,
                              vocab_size=full.vocab_size, rms_norm_eps=full.rms_norm_eps,
                              rope_theta=full.rope_theta, max_position_embeddings=2048)
    if weights is None:
        weights = O.make_weights(cfg, dtype)
    m = O.LlamaOracle(cfg, weights, kv_slots=2048, max_reqs=2, max_ctx=2048, compute="native")
    s = O.OracleSession(m, 2048, 2)
    prompt = make_prompts(1, sample_prefill, cfg.vocab_size)[0]
    t0 = time.perf_counter()
    rids, logits, _ = s.extend([prompt])
    t_pf = time.perf_counter() - t0
    tok = O.LlamaOracle.sample(logits).tolist()
    t0 = time.perf_counter()
    for _ in range(sample_decode):
        logits, _ = s.decode(rids, tok)
        tok = O.LlamaOracle.sample(logits).tolist()
    t_dec = (time.perf_counter() - t0) / sample_decode
    # split head (embedding + lm_head + final norm) from per-layer cost with a 0-layer run
    cfg0 = O.OracleLlamaConfig(**{**cfg.__dict__, "num_hidden_layers": 0})
    w0 = {k: v for k, v in weights.items() if ".layers." not in k}
    m0 = O.LlamaOracle(cfg0, w0, kv_slots=2048, max_reqs=2, max_ctx=2048, compute="native")
    s0 = O.OracleSession(m0, 2048, 2)
    t0 = time.perf_counter()
    r0, lg0, _ = s0.extend([prompt])
    t_pf_head = time.perf_counter() - t0
    t0 = time.perf_counter()
    s0.decode(r0, [1])
    t_dec_head = time.perf_counter() - t0
    L = full.num_hidden_layers
    pf_layer = max(t_pf - t_pf_head, 0.0) / n_layers
    dec_layer = max(t_dec - t_dec_head, 0.0) / n_layers
    t_prefill_full = (pf_layer * L) * (input_len / sample_prefill) + t_pf_head
    t_decode_full = dec_layer * L + t_dec_head
    t_req = t_prefill_full + (output_len - 1) * t_decode_full
    return {"value": output_len / t_req, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": (f"{n_layers} of {L} layers + embedding/lm_head, 1 x {sample_prefill}-token prefill and "
                       f"{sample_decode} batch-1 decode steps, {str(dtype).split('.')[-1]} torch-eager CPU; extrapolated "
                       f"linearly to {L} layers and one {input_len}-in/{output_len}-out request"),
            "prefill_tok_per_s": input_len / t_prefill_full, "decode_tok_per_s": 1.0 / t_decode_full,
            "measured_s": {"prefill_sample": t_pf, "decode_step_sample": t_dec}}


# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    dtype = torch.bfloat16 if args.dtype in ("bfloat16", "bf16") else torch.float16
    config = {"workload": f"{args.model} TP=1 semi-PD {args.prefill_sm_pct}/{args.decode_sm_pct} SM split, "
                          f"synthetic {args.input_len}-in/{args.output_len}-out, {args.requests} requests per step, "
                          f"arrival rate inf (saturated) for value/e2e",
              "requests_per_step": args.requests, "input_len": args.input_len, "output_len": args.output_len,
              "l2": "inputs >> L2: 16 GB of weights + KV stream through the 126 MB L2 every step",
              "parallelism": f"{world} independent replica(s), one engine per GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        vals = []
        for _ in range(max(args.warmup, 0) and 1):
            cpu_baseline(args.model, dtype, 1, args.input_len, args.output_len, sample_prefill=64, sample_decode=1)
        cb = None
        for _ in range(max(args.steps, 1)):
            cb = cpu_baseline(args.model, dtype, args.cpu_layers, args.input_len, args.output_len)
            vals.append(cb["value"])
        v = float(np.mean(vals))
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.output_len / v,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "bf16" if dtype == torch.bfloat16 else "fp16", "data": "synthetic",
                          "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from semi_pd_b200 import _lib
    from semi_pd_b200.configs import ServerArgs, get_model_config
    from semi_pd_b200.semi_pd_scheduler import SemiPDEngine

    cfg = get_model_config(args.model)
    ctx = args.input_len + args.output_len + 8
    kv_tokens = max(4 * args.requests * ctx, 65536)
    graph_bs = [b for b in [1, 2, 4] + [8 * i for i in range(1, 33)] if b <= max(args.requests, 8)]
    if args.requests not in graph_bs:
        graph_bs.append(args.requests)
    sa = ServerArgs(model_path=args.model, dtype="bfloat16" if dtype == torch.bfloat16 else "float16",
                    load_format="dummy", max_total_tokens=kv_tokens, max_running_requests=args.requests,
                    context_length=ctx, enable_semi_pd=True, cuda_graph_bs=sorted(graph_bs),
                    cuda_graph_max_bs=max(graph_bs), chunked_prefill_size=8192, max_prefill_tokens=16384)
    t_setup = time.perf_counter()
    engine = SemiPDEngine(sa, gpu_id=local_rank, model_config=cfg, prefill_sm_percent=args.prefill_sm_pct,
                          decode_sm_percent=args.decode_sm_pct, partition=args.partition)
    engine.warmup(prompt_len=64, n=2, max_new_tokens=4)
    t_setup = time.perf_counter() - t_setup

    prompts = make_prompts(args.requests, args.input_len, cfg.vocab_size)
    dev_prompts = [torch.tensor(p, dtype=torch.int64, device=f"cuda:{local_rank}") for p in prompts]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(device_resident, steps, warmup):
        for _ in range(warmup):
            run_window(engine, prompts, dev_prompts, args.output_len, float("inf"), device_resident)
        barrier()
        n0 = _lib._Counter.launches
        pf0 = engine.prefill.prefill_tokens
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reqs = []
        for _ in range(steps):
            _, r = run_window(engine, prompts, dev_prompts, args.output_len, float("inf"), device_resident)
            reqs += r
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, reqs, _lib._Counter.launches - n0, engine.prefill.prefill_tokens - pf0

    clocks = ClockSampler(local_rank)
    clocks.start()
    ms_dev, _, launches, prefill_tok = timed(True, args.steps, args.warmup)
    # the host-buffer arm repeats the same windows (at most 3 of them, to bound the run time)
    e2e_steps = min(args.steps, 3)
    ms_e2e, reqs_e2e, _, _ = timed(False, e2e_steps, 1)
    clk = clocks.stop()

    out_tokens = args.steps * args.requests * args.output_len * world
    value = out_tokens / (ms_dev / 1e3)
    e2e_value = e2e_steps * args.requests * args.output_len * world / (ms_e2e / 1e3)
    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if dtype == torch.bfloat16 else "fp16",
        "data": "synthetic", "config": config, "clocks": clk, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": args.requests * args.input_len * 8 + args.requests * 64,
                "d2h_bytes_per_step": args.requests * args.output_len * 4,
                "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, **latency_stats(reqs_e2e),
                "note": "saturated window: TTFT includes queueing behind the other requests' prefill"},
        "prefill_tok_per_s_concurrent": prefill_tok * world / (ms_dev / 1e3),
        "sm_partition": {"kind": args.partition, "prefill_sms": engine.p_sms, "decode_sms": engine.d_sms},
        "setup_s": t_setup,
    }
    if rank == 0 and not args.no_latency:
        cap = e2e_value / args.output_len / world  # req/s one GPU sustains
        rate = args.latency_rate if args.latency_rate > 0 else 0.6 * cap
        wall, reqs = run_window(engine, prompts, dev_prompts, args.output_len, rate, False)
        result["latency"] = {"request_rate": rate, "output_tok_per_s": len(reqs) * args.output_len / wall,
                             **latency_stats(reqs)}
    elif not args.no_latency and world > 1:
        pass  # other ranks stay idle during rank 0's latency window
    if rank == 0:
        try:
            result.update(roofline_section(engine, cfg, dtype, min(8192, args.requests * args.input_len),
                                           min(args.requests, 64), args.input_len + args.output_len // 2))
        except Exception as e:  # noqa: BLE001
            result["roofline_error"] = repr(e)
    engine.shutdown()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            result["cpu_baseline"] = cpu_baseline(args.model, dtype, args.cpu_layers, args.input_len, args.output_len)
        except Exception as e:  # noqa: BLE001
            result["cpu_baseline"] = {"error": repr(e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


if __name__ == "__main__":
    main()
