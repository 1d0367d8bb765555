#!/usr/bin/env python
"""bench.py — output tok/s (+ p50 TTFT / TPOT) of the concurrent prefill + decode path.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`; for N > 1 it
is launched under torch.distributed.run, one rank per GPU.  Rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1]): Llama-3-8B, TP = 1, semi-PD with a 50/50 SM split, synthetic
requests of 1024 input / 128 output tokens (token ids as python/sglang/bench_serving.py:772-783,
np.random.seed(1); ignore_eos), random-init weights of that architecture.
A STEP = one serving window: R requests pushed through the engine (prefill worker and decode worker
running concurrently on disjoint SM partitions) until the last token of the last request.
  * `value`   : output tok/s with the prompts already resident in HBM and sampled ids kept on the
                device (no per-step host copies) — the saturated (arrival rate = inf) throughput;
  * `e2e`     : the same K windows served from HOST buffers through the public engine API
                (SemiPDEngine.submit / wait): prompt ids H2D from pinned memory every prefill batch,
                sampled ids D2H every decode step;
  * `parity_check`: greedy ids the e2e arm returned, compared with the CPU oracle's on the same
                weights and prompts (outside the timed region);
  * `latency_sweep`: Poisson windows (exponential gaps, bench_serving.py:884-899) at the rates of
                SURVEY §8d for p50/p99 TTFT and TPOT (TPOT = (t_last - t_first)/(n-1),
                bench_serving.py:920-921); `latency` = the window at 60 % of the measured capacity;
  * `gpu_baseline`: the same request mix through flashinfer's paged wrappers + cuBLAS `F.linear`
                on the same box (scripts/gpu_baseline_flashinfer.py) — the stand-in for the
                reference's own flashinfer build, which cannot be installed here.
N > 1: `value` / `e2e` are N independent replicas (one engine per GPU, no data-path collective,
weak scaling); the `tp` block then serves the same mix with ONE engine sharded TP = N (heads / FFN
columns per GPU, spd_allreduce / NCCL after o_proj and down_proj), `tp_parity` checks a TP = N
model against the unsharded oracle in this very run, and at N = 8 `tp_70b` runs Llama-3-70B TP = 8
(BASELINE.json configs[4]).
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
CPU_SAMPLE_DECODE_STEPS = 8


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
    ap.add_argument("--sweep-rates", type=str, default="4,8,16,32,64", help="req/s per GPU; '' disables")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-tp", action="store_true", help="N > 1: skip the TP = N arm")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--tp-steps", type=int, default=3)
    ap.add_argument("--clocks-csv", type=str, default="", help="write the nvidia-smi samples of the run here")
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
    """Serve one window; returns (wall seconds, finished Req list, rids in submission order)."""
    engine.set_device_resident(device_resident)
    arrivals = poisson_arrivals(len(prompts), rate)
    t0 = time.perf_counter()
    rids = []
    for i, p in enumerate(prompts):
        wait = t0 + arrivals[i] - time.perf_counter()
        if wait > 0:
            time.sleep(wait)
        rids.append(engine.submit(p, output_len, arrival_time=t0 + arrivals[i],
                                  input_ids_device=dev_prompts[i] if device_resident else None))
    reqs = engine.wait(len(prompts))
    torch.cuda.synchronize()
    return time.perf_counter() - t0, reqs, rids


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
            self.lines.append(f"{time.time():.3f}, " + line.strip())

    def stop(self, csv_path=""):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        if csv_path:
            try:
                os.makedirs(os.path.dirname(os.path.abspath(csv_path)), exist_ok=True)
                with open(csv_path, "w") as f:
                    f.write("unix_time, " + self.Q.replace(",", ", ") + "\n" + "\n".join(self.lines) + "\n")
            except OSError:
                pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")][1:]
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


def load_traffic(M):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the newest
    committed `ncu --set full` capture (profiles/traffic.json names the capture)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        if tr.get("M") == M:
            return tr.get("gemm_gate_up_silu_dram_bytes"), tr.get("source")
    except Exception:  # noqa: BLE001
        pass
    return None, None


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
    algo_bytes = (M * H + 2 * I * H + M * I) * 2
    n_sms = ops.device_sm_count(dev.index)
    share = engine.p_sms / float(n_sms)
    traffic, traffic_src = load_traffic(M)
    # timed ALONE on the whole device -> compared with the measured burst peak of the whole device
    ops.set_sm_budget(0)
    full_stream = torch.cuda.Stream(device=dev)
    ms_full = time_on_stream(lambda: ops.linear(x, w, epilogue=ops.SPD_EPI_SILU_MUL), full_stream, flush=flush)
    out["roofline"] = {"bound": "tensor",
                       "kernel": "gemm_pair_kernel<SILU_MUL> (W_gate_up + fused SiLU*mul, cta_group::2 256x256 tiles)",
                       "achieved": flops / ms_full / 1e9, "peak": tf_burst, "unit": "TFLOP/s",
                       "frac": flops / ms_full / 1e9 / tf_burst, "traffic": traffic,
                       "traffic_over_algorithmic": (traffic / algo_bytes) if traffic else None,
                       "algorithmic_bytes": algo_bytes, "traffic_source": traffic_src,
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
    # (2) prefill attention (tensor bound).  Slots are laid out the way the run allocator places
    # them: every request's tokens in consecutive slots (memory_pool.py)
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
    # (3) decode attention (HBM bound) on the decode partition: run-allocated slots (requests'
    # runs interleaved across the pool, each run 128 consecutive slots) and, for comparison, a
    # fully scattered pool (the reference allocator's layout)
    ops.set_sm_budget(engine.d_sms)
    B, S = decode_bs, ctx
    R = 128
    runs_per_req = -(-S // R)
    slots = B * runs_per_req * R + 8
    kp = torch.randn(slots, Hk, D, device=dev).to(dtype)
    vp = torch.randn(slots, Hk, D, device=dev).to(dtype)
    run_ids = torch.randperm(B * runs_per_req, device=dev).view(B, runs_per_req)
    pos = torch.arange(S, device=dev)
    rtt_runs = (run_ids[:, pos // R] * R + 1 + (pos % R)[None, :]).int().contiguous()
    rtt_scat = (torch.randperm(B * S, device=dev).int() + 1).view(B, S).contiguous()
    rpi = torch.arange(B, device=dev)
    sl = torch.full((B,), S, dtype=torch.int64, device=dev)
    qd = torch.randn(B, Hq * D, device=dev).to(dtype)
    by = B * S * Hk * D * 2 * 2 + 2 * B * Hq * D * 2 + 4 * B * S
    for name, rtt_d, stream, budget, note in (
            ("attn_decode", rtt_runs, engine.d_stream, engine.d_sms,
             f"decode partition = {engine.d_sms} SMs, run-allocated pool; peak is the full-device copy bandwidth"),
            ("attn_decode_full_gpu", rtt_runs, full_stream, 0, f"all {n_sms} SMs, run-allocated pool"),
            ("attn_decode_scattered_pool", rtt_scat, engine.d_stream, engine.d_sms,
             f"decode partition = {engine.d_sms} SMs, fully scattered slots (gather fallback)")):
        ops.set_sm_budget(budget)
        ms = time_on_stream(lambda: ops.attn_decode(qd, kp, vp, rtt_d, rpi, sl, Hq, Hk, D, D ** -0.5),
                            stream, iters=10, flush=flush, burst=4)
        extra[name] = {"bound": "hbm", "achieved": by / ms / 1e6, "peak": hbm, "unit": "GB/s",
                       "frac": by / ms / 1e6 / hbm, "launch_ms": ms, "shape": f"B={B} S={S}",
                       "note": note + "; 4 launches per event pair (each streams > 2 x L2 of K/V, no reuse)"}
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
# CPU baseline: the oracle port of the reference forward, FULL model, bounded sample
# ----------------------------------------------------------------------------------------------
def fast_cpu_weights(cfg, dtype):
    """Full-size synthetic weights for the CPU-only `--impl reference` arm, where no device copy
    exists to read them from: one seeded block per parameter, tiled (the values do not change
    what the CPU forward costs; generating 8e9 Gaussians on the host would take longer than the
    measurement)."""
    from oracle import llama_oracle as O
    out = {}
    for idx, (name, shape) in enumerate(O.weight_shapes(cfg).items()):
        n = int(np.prod(shape))
        g = torch.Generator().manual_seed(1234 + idx)
        blk = torch.randn(min(n, 1 << 20), generator=g)
        if len(shape) == 1:
            blk = 1.0 + 0.1 * blk
        elif name == "lm_head.weight":
            blk = blk * (1.5 / math.sqrt(shape[1]))
        elif not name.endswith("embed_tokens.weight"):
            blk = blk / math.sqrt(shape[1])
        blk = blk.to(dtype)
        out[name] = blk.repeat(-(-n // blk.numel()))[:n].view(shape)
    return out


class CpuReference:
    """oracle/llama_oracle.py (torch eager, the model dtype, all host threads) on the FULL model."""

    def __init__(self, model_name, dtype, weights=None, ctx=2048):
        from oracle import llama_oracle as O
        from semi_pd_b200.configs import get_model_config
        self.O = O
        torch.set_num_threads(os.cpu_count() or 1)
        full = get_model_config(model_name)
        self.cfg = O.OracleLlamaConfig(
            hidden_size=full.hidden_size, intermediate_size=full.intermediate_size,
            num_hidden_layers=full.num_hidden_layers, num_attention_heads=full.num_attention_heads,
            num_key_value_heads=full.num_key_value_heads, head_dim=full.head_dim, vocab_size=full.vocab_size,
            rms_norm_eps=full.rms_norm_eps, rope_theta=full.rope_theta, max_position_embeddings=ctx)
        self.weights = weights if weights is not None else fast_cpu_weights(self.cfg, dtype)
        self.dtype = dtype
        self.ctx = ctx

    def sample(self, prompt, decode_steps):
        """One request: prefill + `decode_steps` greedy steps.  → (t_prefill, t_decode_avg, ids,
        margins) — the first id comes from the prefill, as in the reference."""
        O = self.O
        m = O.LlamaOracle(self.cfg, self.weights, kv_slots=self.ctx, max_reqs=2, max_ctx=self.ctx, compute="native")
        s = O.OracleSession(m, self.ctx, 2)
        t0 = time.perf_counter()
        rids, logits, _ = s.extend([prompt])
        t_pf = time.perf_counter() - t0
        ids, margins = [], []
        t_dec = 0.0
        for k in range(decode_steps + 1):
            top2 = logits.topk(2, -1).values[0]
            margins.append(float(top2[0] - top2[1]))
            ids.append(int(logits.argmax(-1)[0]))
            if k == decode_steps:
                break
            t0 = time.perf_counter()
            logits, _ = s.decode(rids, [ids[-1]])
            t_dec += time.perf_counter() - t0
        return t_pf, t_dec / max(decode_steps, 1), ids, margins


def cpu_baseline(ref: CpuReference, prompt, input_len, output_len, decode_steps=CPU_SAMPLE_DECODE_STEPS):
    t_pf, t_dec, ids, margins = ref.sample(prompt, decode_steps)
    t_req = t_pf + (output_len - 1) * t_dec
    L = ref.cfg.num_hidden_layers
    return {"value": output_len / t_req, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": (f"full {L}-layer model: 1 x {input_len}-token prefill + {decode_steps} batch-1 decode steps, "
                       f"{str(ref.dtype).split('.')[-1]} torch-eager CPU (oracle/llama_oracle.py); value = "
                       f"{output_len} / (t_prefill + {output_len - 1} x mean decode step) — no layer extrapolation"),
            "prefill_tok_per_s": input_len / t_pf, "decode_tok_per_s": 1.0 / t_dec,
            "measured_s": {"prefill": t_pf, "decode_step_mean": t_dec, "sample_total": t_pf + decode_steps * t_dec},
            "sample_tok_per_s": (decode_steps + 1) / (t_pf + decode_steps * t_dec)}, ids, margins


def compare_ids(got, want, margins, min_margin):
    """Greedy ids are compared position by position until the oracle's top-1/top-2 margin drops
    below `min_margin` (beyond a near-tie two correct implementations may legitimately diverge)."""
    compared = matched = 0
    for g, w, m in zip(got, want, margins):
        if m <= min_margin:
            break
        compared += 1
        matched += int(g == w)
        if g != w:
            break
    return compared, matched


# ----------------------------------------------------------------------------------------------
# GPU stand-in baseline (flashinfer wrappers + cuBLAS), separate process
# ----------------------------------------------------------------------------------------------
def run_gpu_baseline(args, gpu, rate=0.0):
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "gpu_baseline_flashinfer.py"), "--gpu", str(gpu),
           "--requests", str(args.requests), "--input-len", str(args.input_len),
           "--output-len", str(args.output_len), "--dtype", args.dtype, "--windows", "2", "--rate", str(rate)]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
        for line in reversed(p.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": (p.stderr.strip().splitlines() or ["no output"])[-1][:300]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:300]}


# ----------------------------------------------------------------------------------------------
# TP arm (N > 1)
# ----------------------------------------------------------------------------------------------
TP_PARITY_CFG = dict(hidden_size=1024, intermediate_size=2048, num_hidden_layers=2, num_attention_heads=8,
                     num_key_value_heads=4, head_dim=128, vocab_size=4096, max_position_embeddings=2048)


def tp_parity_check(rank, world, local_rank):
    """TP = N ModelRunner (sharded heads / FFN, spd_allreduce, vocab-sharded lm_head) against the
    UNSHARDED fp32 CPU oracle on identical weights: extend + 3 decode steps (eager and, when the
    runner captured it, through the decode CUDA graph).  → max-abs logit error over all ranks."""
    import torch.distributed as dist
    from oracle import llama_oracle as O
    from semi_pd_b200.configs import ModelConfig, ServerArgs
    from semi_pd_b200.model_runner import ModelRunner
    from tests.helpers import GpuSession
    dtype = torch.float16
    ocfg = O.OracleLlamaConfig(**TP_PARITY_CFG)
    weights = O.make_weights(ocfg, dtype, seed=99)
    out = {}
    for use_graph in (False, True):
        sa = ServerArgs(model_path="tp-parity", dtype="float16", load_format="external", tp_size=world,
                        max_total_tokens=4096, max_running_requests=8, context_length=512,
                        disable_cuda_graph=not use_graph, cuda_graph_bs=[1, 2, 4], enable_semi_pd=True)
        mr = ModelRunner(ModelConfig(name="tp-parity", **TP_PARITY_CFG), 0.5, local_rank, rank, world, 0, sa)
        mr.model.load_weights(weights.items())
        mr.init_attention_backend()
        mr.init_cuda_graphs()
        oracle = O.LlamaOracle(ocfg, weights, kv_slots=2048, max_reqs=9, max_ctx=516, compute="fp32")
        gs, os_ = GpuSession(mr), O.OracleSession(oracle, 2048, 9)
        prompts = [[(5 + 3 * i + 17 * j) % TP_PARITY_CFG["vocab_size"] for i in range(n)]
                   for j, n in enumerate([40, 7, 129])]
        reqs, lg, ids = gs.extend(prompts)
        rids, lc, _ = os_.extend(prompts)
        err = (lg - lc).abs().max().item()
        bad = int((ids.long() != lc.argmax(-1)).sum())
        tok = lc.argmax(-1).tolist()
        for _ in range(3):
            lg, ids = gs.decode(reqs, tok)
            lc, _ = os_.decode(rids, tok)
            if lg.shape == lc.shape:  # a captured step returns only this rank's logits shard
                err = max(err, (lg - lc).abs().max().item())
            bad += int((ids.long() != lc.argmax(-1)).sum())
            tok = lc.argmax(-1).tolist()
        t = torch.tensor([err, float(bad), float(mr.tp_comm.status())], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["graph" if use_graph else "eager"] = {"max_abs": float(t[0]), "wrong_ids": int(t[1]),
                                                  "ar_timeouts": int(t[2]),
                                                  "graph_captured": mr.cuda_graph_runner is not None}
        mr.tp_comm.close()
        del mr, gs
        torch.cuda.empty_cache()
    return {"tp_size": world, "model": "2-layer hidden 1024, 8 q / 4 kv heads, vocab 4096, fp16",
            "tolerance": 1e-2, **out,
            "tp_parity_max_abs": max(out["eager"]["max_abs"], out["graph"]["max_abs"])}


def tp_arm(args, model_name, rank, world, local_rank, dtype, requests, steps, tag):
    """One semi-PD engine sharded TP = world over all GPUs; rank 0 drives, the others serve."""
    import torch.distributed as dist
    from semi_pd_b200.configs import ServerArgs, get_model_config
    from semi_pd_b200.semi_pd_scheduler import SemiPDEngine
    cfg = get_model_config(model_name)
    ctx = args.input_len + args.output_len + 8
    kv_tokens = max(2 * requests * (-(-ctx // 128) * 128), 32768)
    graph_bs = sorted(set([b for b in [1, 2, 4] + [8 * i for i in range(1, 33)] if b <= max(requests, 8)] + [requests]))
    sa = ServerArgs(model_path=model_name, dtype="bfloat16" if dtype == torch.bfloat16 else "float16",
                    load_format="dummy", tp_size=world, max_total_tokens=kv_tokens, max_running_requests=requests,
                    context_length=ctx, enable_semi_pd=True, cuda_graph_bs=graph_bs, cuda_graph_max_bs=max(graph_bs),
                    chunked_prefill_size=8192, max_prefill_tokens=16384)
    t0 = time.perf_counter()
    engine = SemiPDEngine(sa, gpu_id=local_rank, tp_rank=rank, model_config=cfg,
                          prefill_sm_percent=args.prefill_sm_pct, decode_sm_percent=args.decode_sm_pct,
                          partition=args.partition)
    setup = time.perf_counter() - t0
    res = None
    if rank == 0:
        prompts = make_prompts(requests, args.input_len, cfg.vocab_size)
        engine.warmup(prompt_len=64, n=2, max_new_tokens=4)
        run_window(engine, prompts, None, args.output_len, float("inf"), False)  # warm-up window
        walls, reqs_all = [], []
        for _ in range(steps):
            w, reqs, _ = run_window(engine, prompts, None, args.output_len, float("inf"), False)
            walls.append(w)
            reqs_all += reqs
        tok_s = steps * requests * args.output_len / sum(walls)
        cap = tok_s / args.output_len
        w, reqs, _ = run_window(engine, prompts, None, args.output_len, 0.6 * cap, False)
        res = {"model": model_name, "tp_size": world, "requests_per_step": requests, "steps": steps,
               "output_tok_per_s": tok_s, "ms_per_step": 1e3 * sum(walls) / steps,
               "saturated": latency_stats(reqs_all),
               "latency": {"request_rate": 0.6 * cap, "output_tok_per_s": len(reqs) * args.output_len / w,
                           **latency_stats(reqs)},
               "decode_graphs": engine.d_worker.model_runner.cuda_graph_runner is not None,
               "ar_timeouts": engine.d_worker.model_runner.tp_comm.status() + engine.p_worker.model_runner.tp_comm.status(),
               "sm_partition": {"prefill_sms": engine.p_sms, "decode_sms": engine.d_sms}, "setup_s": setup,
               "collectives": "spd_allreduce (peer memory over NVLink, one-/two-shot, graph-captured in decode); "
                              "messages > 32 MiB (prefill batches) go through the same kernel in slot-sized "
                              "pieces; (max, argmax) spd_allgather for the vocab-sharded lm_head; no NCCL "
                              "call on the serving path",
               "timing": "host wall clock on rank 0 around submit .. last token (all ranks run in lockstep)"}
        engine.shutdown()
    else:
        engine.join()
        engine.shutdown()
    dist.barrier()
    torch.cuda.empty_cache()
    return res


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
              "parallelism": f"{world} independent replica(s), one engine per GPU"
                             + (f"; plus a TP={world} engine reported under `tp`" if world > 1 else "")}

    if args.impl == "reference":
        if rank != 0:
            return
        from semi_pd_b200.configs import get_model_config
        vocab = get_model_config(args.model).vocab_size
        prompt = make_prompts(1, args.input_len, vocab)[0]
        ref = CpuReference(args.model, dtype)
        for _ in range(max(args.warmup, 0)):
            cpu_baseline(ref, prompt, args.input_len, args.output_len)
        vals, walls, cb = [], [], None
        for _ in range(max(args.steps, 1)):
            t0 = time.perf_counter()
            cb, _, _ = cpu_baseline(ref, prompt, args.input_len, args.output_len)
            walls.append(time.perf_counter() - t0)
            vals.append(cb["value"])
        v = float(np.mean(vals))
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "bf16" if dtype == torch.bfloat16 else "fp16", "data": "synthetic",
                          "config": config, "cpu_baseline": cb,
                          "note": "a step = the bounded sample named in cpu_baseline.sample (ms_per_step is its "
                                  "measured wall time); value converts the two measured phase times to one "
                                  f"{args.input_len}-in/{args.output_len}-out request",
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
    # every request occupies whole 128-slot runs; room for the running batch + two admitted batches
    kv_tokens = max(2 * args.requests * (-(-ctx // 128) * 128), 65536)
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
        reqs, first_rids = [], None
        for _ in range(steps):
            _, r, rids = run_window(engine, prompts, dev_prompts, args.output_len, float("inf"), device_resident)
            reqs += r
            first_rids = first_rids or rids
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, reqs, _lib._Counter.launches - n0, engine.prefill.prefill_tokens - pf0, first_rids

    clocks = ClockSampler(local_rank)
    clocks.start()
    ms_dev, _, launches, prefill_tok, _ = timed(True, args.steps, args.warmup)
    # the host-buffer arm repeats the same K windows through submit / wait
    ms_e2e, reqs_e2e, _, _, e2e_rids = timed(False, args.steps, 1)
    clk = clocks.stop(args.clocks_csv if rank == 0 else "")
    e2e_by_rid = {r.rid: r for r in reqs_e2e}
    e2e_first_window_ids = [e2e_by_rid[r].output_ids for r in e2e_rids] if e2e_rids else []

    out_tokens = args.steps * args.requests * args.output_len * world
    value = out_tokens / (ms_dev / 1e3)
    e2e_value = out_tokens / (ms_e2e / 1e3)
    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if dtype == torch.bfloat16 else "fp16",
        "data": "synthetic", "config": config, "clocks": clk, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": args.requests * args.input_len * 8 + args.requests * 64,
                "d2h_bytes_per_step": args.requests * args.output_len * 4,
                "ms_per_step": ms_e2e / args.steps, "steps": args.steps, **latency_stats(reqs_e2e),
                "note": "saturated window: TTFT includes queueing behind the other requests' prefill"},
        "prefill_tok_per_s_concurrent": prefill_tok * world / (ms_dev / 1e3),
        "sm_partition": {"kind": args.partition, "prefill_sms": engine.p_sms, "decode_sms": engine.d_sms},
        "kv_pool": {"tokens": engine.max_total_num_tokens, "allocator": "128-slot runs behind page_size = 1",
                    "retractions": engine.decode.num_retracted},
        "setup_s": t_setup,
    }
    if rank == 0 and not args.no_latency:
        cap = e2e_value / args.output_len / world  # req/s one GPU sustains
        rate = args.latency_rate if args.latency_rate > 0 else 0.6 * cap
        wall, reqs, _ = run_window(engine, prompts, dev_prompts, args.output_len, rate, False)
        result["latency"] = {"request_rate": rate, "output_tok_per_s": len(reqs) * args.output_len / wall,
                             **latency_stats(reqs)}
        sweep = []
        for r_s in [float(x) for x in args.sweep_rates.split(",") if x.strip()]:
            # SURVEY §8d rates; each window is bounded to ~6 s of arrivals (at least 32 requests)
            n = int(min(args.requests, max(32, round(r_s * 6))))
            wall, reqs, _ = run_window(engine, prompts[:n], dev_prompts[:n], args.output_len, r_s, False)
            sweep.append({"request_rate": r_s, "requests": n, "saturated": r_s > cap,
                          "output_tok_per_s": len(reqs) * args.output_len / wall, **latency_stats(reqs)})
        result["latency_sweep"] = sweep
    if rank == 0 and not args.no_roofline:
        try:
            result.update(roofline_section(engine, cfg, dtype, min(8192, args.requests * args.input_len),
                                           min(args.requests, 64), args.input_len + args.output_len // 2))
        except Exception as e:  # noqa: BLE001
            result["roofline_error"] = repr(e)
    # weights for the CPU leg / parity check come off the device before the engine goes away
    cpu_weights = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_weights = {k: v.cpu() for k, v in engine.d_worker.model_runner.model.state_dict().items()}
        except Exception as e:  # noqa: BLE001
            result["cpu_baseline"] = {"error": "weights D2H: " + repr(e)}
    engine.shutdown()
    del engine, dev_prompts
    torch.cuda.empty_cache()

    if cpu_weights is not None:
        try:
            ref = CpuReference(args.model, dtype, weights=cpu_weights, ctx=args.input_len + 64)
            cb, ids0, m0 = cpu_baseline(ref, prompts[0], args.input_len, args.output_len)
            result["cpu_baseline"] = cb
            # parity of the e2e arm's returned ids with the oracle (same weights, same prompts);
            # bf16 moves O(1) logits by ~2^-8 per rounding point, hence the wider margin rule
            min_margin = 2e-2 if dtype == torch.float16 else 1e-1
            checks = [(0, ids0, m0)]
            _, _, ids1, m1 = ref.sample(prompts[1], 3)
            checks.append((1, ids1, m1))
            detail, comp, match = [], 0, 0
            for i, ids, margins in checks:
                got = e2e_first_window_ids[i][: len(ids)]
                c, m = compare_ids(got, ids, margins, min_margin)
                comp, match = comp + c, match + m
                detail.append({"request": i, "oracle_ids": ids, "engine_ids": got, "compared": c, "matched": m,
                               "min_margin_seen": min(margins)})
            result["parity_check"] = {"requests": len(checks), "ids_compared": comp, "ids_matched": match,
                                      "ok": comp > 0 and comp == match, "margin_rule": min_margin,
                                      "oracle": "oracle/llama_oracle.py, native " + str(dtype).split(".")[-1]
                                                + ", full 32-layer model, weights copied off the device",
                                      "detail": detail}
        except Exception as e:  # noqa: BLE001
            result.setdefault("cpu_baseline", {"error": repr(e)})
            result["parity_check"] = {"error": repr(e)}
        del cpu_weights
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        # same mix, same box, same run; its Poisson window runs at the rate of our `latency` block
        gb = run_gpu_baseline(args, local_rank, result.get("latency", {}).get("request_rate", 0.0))
        result["gpu_baseline"] = gb
        if "output_tok_per_s" in gb:
            cmp_ = {"output_tok_per_s": result["e2e"]["value"] / gb["output_tok_per_s"],
                    "saturated_tpot_p50": gb["tpot_ms_p50"] / result["e2e"]["tpot_ms_p50"],
                    "saturated_ttft_p50": gb["ttft_ms_p50"] / result["e2e"]["ttft_ms_p50"]}
            if gb.get("latency") and result.get("latency"):
                cmp_["at_rate_ttft_p50"] = gb["latency"]["ttft_ms_p50"] / result["latency"]["ttft_ms_p50"]
                cmp_["at_rate_tpot_p50"] = gb["latency"]["tpot_ms_p50"] / result["latency"]["tpot_ms_p50"]
            result["vs_gpu_baseline"] = {**cmp_, "note": "each > 1 means this build is better (throughput: ours / "
                                                        "baseline; latencies: baseline / ours)"}

    if world > 1 and not args.no_tp:
        # The TP arm must never cost the replica numbers above: if a rank fails or the arm runs
        # away (a hung collective), rank 0 still prints the line, with `tp_error` set.
        def bail(msg):
            if rank == 0:
                result["tp_error"] = msg
                print(json.dumps(result), flush=True)
            os._exit(0)

        tp_limit = 900.0 if world == 8 else 480.0
        watchdog = threading.Timer(tp_limit, bail, args=(f"TP arm exceeded {tp_limit:.0f} s",))
        watchdog.daemon = True
        watchdog.start()
        try:
            par = tp_parity_check(rank, world, local_rank)
            if rank == 0:
                result["tp_parity"] = par
                result["tp_parity_max_abs"] = par["tp_parity_max_abs"]
            tp = tp_arm(args, args.model, rank, world, local_rank, dtype, args.requests, args.tp_steps, "tp")
            if rank == 0:
                result["tp"] = tp
            if world == 8:
                tp70 = tp_arm(args, "llama-3-70b", rank, world, local_rank, dtype, 64, 2, "tp_70b")
                if rank == 0:
                    result["tp_70b"] = tp70
        except BaseException as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            bail(repr(e)[:500])
        watchdog.cancel()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


if __name__ == "__main__":
    main()
