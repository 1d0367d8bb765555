#!/usr/bin/env python
"""GPU stand-in baseline: the reference's forward path rebuilt from the LIBRARY kernels it calls.

The reference's own `--enable-semi-pd` build (torch 2.5.1 / flashinfer 0.2.3 / sgl-kernel 0.0.5 /
vllm <= 0.7.2 pins, MPS) cannot be installed in this image.  What the image does
carry is flashinfer 0.6.11 — the same wrapper API the reference calls — and cuBLAS.  This script
serves the bench's request mix (R x 1024-in / 128-out, saturated) with exactly the kernel sequence
of the reference's Llama layer (python/sglang/srt/models/llama.py:172-183,237-259):

  RMSNorm / fused_add_rmsnorm   flashinfer.norm            (= sgl_kernel.rmsnorm, layernorm.py:47-57)
  qkv / o / gate_up / down      F.linear -> cuBLAS         (linear.py:165-172)
  RoPE                          flashinfer.rope.apply_rope_with_cos_sin_cache_inplace
  KV write                      two index_put_             (memory_pool.py:316-346)
  prefill attention             BatchPrefillWithRaggedKVCacheWrapper, causal
                                (flashinfer_backend.py:425-432; no cached prefix in this mix)
  decode attention              BatchDecodeWithPagedKVCacheWrapper, page_size 1, CUDA graph per
                                batch size, host plan() before every replay
                                (flashinfer_backend.py:479-486, cuda_graph_runner.py:494-504)
  SiLU*mul                      flashinfer.activation.silu_and_mul
  lm_head + greedy              F.linear + argmax          (logits_processor.py:416-440, sampler.py:73-77)

Prefill and decode run CONCURRENTLY from two host threads on two CUDA streams of one process (the
box offers no MPS daemon, so the reference's per-process SM caps become the hardware's own
scheduling of two streams; decode gets the higher stream priority).  Admission mirrors bench.py:
prefill batches of 8 prompts, decode batch grows as prefills finish.  Everything here is library
code by the task's definition — it is the baseline to beat, not part of the product.

Prints ONE JSON line: output tok/s over `--windows` windows, TTFT / TPOT percentiles.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# JIT cache prebuilt in the build container travels with the repo (baseline/_ref is git-ignored,
# not gpurun-ignored).  /root/repo is valid in the build container AND on the GPU box (a symlink
# to the snapshot there), so the absolute paths ninja recorded stay valid and nothing recompiles.
_base = "/root/repo" if os.path.isdir("/root/repo/scripts") else ROOT
os.environ.setdefault("FLASHINFER_WORKSPACE_BASE", os.path.join(_base, "baseline", "_ref", "flashinfer"))
os.environ.setdefault("FLASHINFER_CUDA_ARCH_LIST", "10.0a")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--requests", type=int, default=192)
    ap.add_argument("--input-len", type=int, default=1024)
    ap.add_argument("--output-len", type=int, default=128)
    ap.add_argument("--dtype", type=str, default="bfloat16")
    ap.add_argument("--windows", type=int, default=2)
    ap.add_argument("--prefill-batch", type=int, default=8)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--rate", type=float, default=0.0,
                    help="also serve one Poisson window at this request rate (req/s) for TTFT / TPOT")
    ap.add_argument("--build-only", action="store_true", help="JIT-compile the flashinfer modules and exit (no GPU)")
    return ap.parse_args()


H, I, L, HQ, HK, D, V = 4096, 14336, 32, 32, 8, 128, 128256
EPS, THETA = 1e-5, 500000.0


def build_only(dtype):
    """Compile (nvcc, no GPU needed) the JIT modules the run will request, into the cache dir."""
    import flashinfer  # noqa: F401
    from flashinfer.activation import get_act_and_mul_module
    from flashinfer.decode import get_batch_decode_module
    from flashinfer.norm import get_norm_module
    from flashinfer.prefill import get_batch_prefill_module
    from flashinfer.rope import get_rope_module
    t0 = time.time()
    get_norm_module()
    get_act_and_mul_module("silu")
    get_rope_module()
    get_batch_decode_module(dtype, dtype, dtype, torch.int32, D, D, 0, False, False)
    get_batch_prefill_module("fa2", dtype, dtype, dtype, torch.int32, D, D, 0, False, False, False)
    print(f"flashinfer modules built in {time.time() - t0:.0f} s under {os.environ['FLASHINFER_WORKSPACE_BASE']}")


class Model:
    def __init__(self, dev, dtype, layers):
        g = torch.Generator(device=dev).manual_seed(1234)
        r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).to(dtype)  # noqa: E731
        self.embed = r(V, H)
        self.lm_head = r(V, H, sc=1.5 / H ** 0.5)
        self.layers = []
        for _ in range(layers):
            self.layers.append(dict(
                ln1=torch.ones(H, device=dev, dtype=dtype), ln2=torch.ones(H, device=dev, dtype=dtype),
                qkv=r((HQ + 2 * HK) * D, H, sc=H ** -0.5), o=r(H, HQ * D, sc=(HQ * D) ** -0.5),
                gate_up=r(2 * I, H, sc=H ** -0.5), down=r(H, I, sc=I ** -0.5)))
        self.norm = torch.ones(H, device=dev, dtype=dtype)
        inv = 1.0 / (THETA ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
        fr = torch.outer(torch.arange(4096, dtype=torch.float32), inv)
        self.cos_sin = torch.cat([fr.cos(), fr.sin()], -1).to(dev)
        self.n_layers = layers

    def forward(self, ids, positions, loc, k_pool, v_pool, attn):
        import flashinfer
        h = F.embedding(ids, self.embed)
        res = None
        for li, w in enumerate(self.layers):
            if res is None:
                res = h
                h = flashinfer.norm.rmsnorm(h, w["ln1"], EPS)
            else:
                flashinfer.norm.fused_add_rmsnorm(h, res, w["ln1"], EPS)
            qkv = F.linear(h, w["qkv"])
            q, k, v = qkv.split([HQ * D, HK * D, HK * D], dim=-1)
            flashinfer.rope.apply_rope_with_cos_sin_cache_inplace(positions, q, k, D, self.cos_sin, True)
            k_pool[li][loc] = k.view(-1, HK, D)
            v_pool[li][loc] = v.view(-1, HK, D)
            o = attn(li, q.contiguous().view(-1, HQ, D), k.view(-1, HK, D), v.view(-1, HK, D))
            h = F.linear(o.view(-1, HQ * D), w["o"])
            flashinfer.norm.fused_add_rmsnorm(h, res, w["ln2"], EPS)
            gu = F.linear(h, w["gate_up"])
            h = F.linear(flashinfer.activation.silu_and_mul(gu), w["down"])
        flashinfer.norm.fused_add_rmsnorm(h, res, self.norm, EPS)
        return h


def main():
    a = parse()
    dtype = torch.bfloat16 if a.dtype in ("bfloat16", "bf16") else torch.float16
    if a.build_only:
        build_only(dtype)
        return
    import flashinfer
    dev = torch.device("cuda", a.gpu)
    torch.cuda.set_device(dev)
    R, S_in, S_out, PB = a.requests, a.input_len, a.output_len, a.prefill_batch
    ctx = S_in + S_out
    model = Model(dev, dtype, a.layers)
    slots = R * ctx + 1
    k_pool = [torch.zeros(slots, HK, D, device=dev, dtype=dtype) for _ in range(a.layers)]
    v_pool = [torch.zeros(slots, HK, D, device=dev, dtype=dtype) for _ in range(a.layers)]
    rng = np.random.RandomState(1)
    off = rng.randint(0, V, size=R)
    prompts = torch.tensor([[(off[i] + i + j) % V for j in range(S_in)] for i in range(R)], dtype=torch.int64).pin_memory()
    base = torch.arange(R, dtype=torch.int64) * ctx + 1  # request r owns slots base[r] .. base[r]+ctx-1

    p_stream = torch.cuda.Stream(device=dev, priority=0)
    d_stream = torch.cuda.Stream(device=dev, priority=-1)
    ws_p = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ws_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pre_w = flashinfer.BatchPrefillWithRaggedKVCacheWrapper(ws_p, "NHD", backend="fa2")

    # ---- decode: one CUDA graph + one wrapper per batch size (cuda_graph_runner.py:116-118) ----
    bs_list = sorted(set([1, 2, 4] + [8 * i for i in range(1, R // 8 + 1)] + [R]))
    static = dict(ids=torch.zeros(R, dtype=torch.int64, device=dev), pos=torch.zeros(R, dtype=torch.int64, device=dev),
                  loc=torch.zeros(R, dtype=torch.int64, device=dev))
    graphs, wrappers, outs = {}, {}, {}
    kv_indices_buf = torch.zeros(R * ctx, dtype=torch.int32, device=dev)
    with torch.cuda.stream(d_stream):
        for bs in reversed(bs_list):
            indptr_buf = torch.zeros(bs + 1, dtype=torch.int32, device=dev)
            last_buf = torch.ones(bs, dtype=torch.int32, device=dev)
            w = flashinfer.BatchDecodeWithPagedKVCacheWrapper(
                ws_d, "NHD", use_cuda_graph=True, use_tensor_cores=False,  # GQA group 4 (flashinfer_backend.py:1113-1120)
                paged_kv_indptr_buffer=indptr_buf, paged_kv_indices_buffer=kv_indices_buf,
                paged_kv_last_page_len_buffer=last_buf)
            indptr = torch.arange(bs + 1, dtype=torch.int32) * 8
            w.plan(indptr, torch.arange(bs * 8, dtype=torch.int32), torch.ones(bs, dtype=torch.int32), HQ, HK, D, 1,
                   q_data_type=dtype, kv_data_type=dtype)
            wrappers[bs] = w

            def run(bs=bs, w=w):
                attn = lambda li, q, k, v: w.run(q, (k_pool[li].unsqueeze(1), v_pool[li].unsqueeze(1)))  # noqa: E731
                h = model.forward(static["ids"][:bs], static["pos"][:bs], static["loc"][:bs], k_pool, v_pool, attn)
                return torch.argmax(F.linear(h, model.lm_head).float(), -1)

            run()
            d_stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=d_stream):
                outs[bs] = run()
            graphs[bs] = g
    torch.cuda.synchronize()

    def window(rate=0.0):
        lock = threading.Lock()
        ready = []           # (request index, first token tensor on device, done event)
        t_first, t_last = [0.0] * R, [0.0] * R
        n_tok = [0] * R
        # arrivals: all at once (saturated) or exponential gaps as bench_serving.py:884-899
        arr = [0.0] * R if rate <= 0 else np.cumsum(np.random.RandomState(1).exponential(1.0 / rate, size=R)).tolist()
        t0 = time.perf_counter()

        def prefill_loop():
            with torch.cuda.stream(p_stream):
                nxt = 0
                while nxt < R:
                    now = time.perf_counter() - t0
                    if arr[nxt] > now:
                        time.sleep(min(arr[nxt] - now, 0.002))
                        continue
                    idx = []
                    while nxt < R and len(idx) < PB and arr[nxt] <= now:
                        idx.append(nxt)
                        nxt += 1
                    n = len(idx)
                    ids = prompts[idx].reshape(-1).to(dev, non_blocking=True)
                    pos = torch.arange(S_in, device=dev).repeat(n)
                    loc = (base[idx].to(dev)[:, None] + torch.arange(S_in, device=dev)[None, :]).reshape(-1)
                    qo = torch.arange(n + 1, dtype=torch.int32) * S_in
                    pre_w.plan(qo, qo, HQ, HK, D, causal=True, q_data_type=dtype)
                    attn = lambda li, q, k, v: pre_w.run(q, k, v)  # noqa: E731
                    h = model.forward(ids, pos, loc, k_pool, v_pool, attn)
                    last = h[torch.arange(n, device=dev) * S_in + S_in - 1]
                    tok = torch.argmax(F.linear(last, model.lm_head).float(), -1)
                    ev = torch.cuda.Event()
                    ev.record(p_stream)
                    ev.synchronize()
                    now = time.perf_counter()
                    with lock:
                        for j, r in enumerate(idx):
                            t_first[r] = t_last[r] = now
                            n_tok[r] = 1
                            ready.append((r, tok[j:j + 1]))

        def decode_loop():
            running, cur_tok, lens = [], None, []
            done = 0
            with torch.cuda.stream(d_stream):
                while done < R:
                    with lock:
                        new = list(ready)
                        ready.clear()
                    if new:
                        running += [r for r, _ in new]
                        lens += [S_in] * len(new)
                        add = torch.cat([t for _, t in new])
                        cur_tok = add if cur_tok is None else torch.cat([cur_tok, add])
                    if not running:
                        time.sleep(0.0002)
                        continue
                    n = len(running)
                    bs = next(b for b in bs_list if b >= n)
                    static["ids"][:n].copy_(cur_tok)
                    lens_t = torch.tensor(lens, dtype=torch.int64)
                    static["pos"][:n].copy_(lens_t.to(dev, non_blocking=True))
                    static["loc"][:n].copy_((base[running] + lens_t).to(dev, non_blocking=True))
                    if bs > n:
                        static["loc"][n:bs].zero_()
                        static["pos"][n:bs].zero_()
                    # host-side plan before every replay (the reference's per-step re-plan)
                    kv_lens = torch.tensor([x + 1 for x in lens] + [1] * (bs - n), dtype=torch.int32)
                    indptr = torch.zeros(bs + 1, dtype=torch.int32)
                    indptr[1:] = torch.cumsum(kv_lens, 0)
                    starts = torch.cat([base[running], torch.zeros(bs - n, dtype=torch.int64)])
                    kv_idx = torch.cat([torch.arange(int(s), int(s) + int(k_), dtype=torch.int32)
                                        for s, k_ in zip(starts.tolist(), kv_lens.tolist())])
                    wrappers[bs].plan(indptr, kv_idx, torch.ones(bs, dtype=torch.int32), HQ, HK, D, 1,
                                      q_data_type=dtype, kv_data_type=dtype)
                    graphs[bs].replay()
                    cur_tok = outs[bs][:n].clone()
                    d_stream.synchronize()
                    now = time.perf_counter()
                    keep = []
                    for j, r in enumerate(running):
                        lens[j] += 1
                        n_tok[r] += 1
                        t_last[r] = now
                        if n_tok[r] >= S_out:
                            done += 1
                        else:
                            keep.append(j)
                    if len(keep) != n:
                        running = [running[j] for j in keep]
                        lens = [lens[j] for j in keep]
                        cur_tok = cur_tok[torch.tensor(keep, device=dev, dtype=torch.int64)] if keep else None

        tp = threading.Thread(target=prefill_loop)
        td = threading.Thread(target=decode_loop)
        tp.start()
        td.start()
        tp.join()
        td.join()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ttft = [t_first[r] - t0 - arr[r] for r in range(R)]
        tpot = [(t_last[r] - t_first[r]) / (S_out - 1) for r in range(R)]
        return wall, ttft, tpot

    window()  # warm-up
    walls, ttft, tpot = [], [], []
    for _ in range(a.windows):
        w, t1, t2 = window()
        walls.append(w)
        ttft += t1
        tpot += t2
    q = lambda x, p: float(np.percentile(np.array(x) * 1e3, p))  # noqa: E731
    at_rate = None
    if a.rate > 0:
        w, t1, t2 = window(a.rate)
        at_rate = {"request_rate": a.rate, "output_tok_per_s": R * S_out / w, "ttft_ms_p50": q(t1, 50),
                   "ttft_ms_p99": q(t1, 99), "tpot_ms_p50": q(t2, 50), "tpot_ms_p99": q(t2, 99)}
    print(json.dumps({
        "kind": "flashinfer-" + flashinfer.__version__ + " wrappers + cuBLAS F.linear stand-in (NOT the reference build)",
        "output_tok_per_s": a.windows * R * S_out / sum(walls), "ms_per_window": 1e3 * sum(walls) / a.windows,
        "windows": a.windows, "requests": R, "input_len": S_in, "output_len": S_out, "layers": a.layers,
        "ttft_ms_p50": q(ttft, 50), "ttft_ms_p99": q(ttft, 99), "tpot_ms_p50": q(tpot, 50), "tpot_ms_p99": q(tpot, 99),
        "latency": at_rate,
        "concurrency": "prefill and decode threads on two CUDA streams of one process (no MPS on the box), "
                       "decode stream at higher priority; CUDA-graph decode with a host plan() per step",
        "dtype": a.dtype}))


if __name__ == "__main__":
    main()
