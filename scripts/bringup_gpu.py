"""GPU bring-up checks (scratch tool, not a test): every libspd op against a plain torch
computation on the same device, plus quick timings.  Run under gpurun:
    python scripts/bringup_gpu.py [gemm|elem|attn|all]
"""
import math
import sys
import time

import torch

sys.path.insert(0, ".")
from semi_pd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def check_gemm():
    print("== GEMM ==", flush=True)
    shapes = [
        (1, 256, 128), (5, 256, 256), (16, 384, 512), (64, 4096, 4096), (160, 6144, 4096),
        (200, 512, 1024), (256, 4096, 14336), (300, 1024, 512), (1024, 6144, 4096),
        (1000, 4096, 4096), (2048, 28672, 4096), (1024, 4096, 14336), (8, 128256, 4096),
        (4096, 1280, 4096),
    ]
    for dt in (torch.bfloat16, torch.float16):
        for (M, N, K) in shapes:
            a = (torch.randn(M, K, device=dev) * 0.5).to(dt)
            w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(dt)
            ref = (a.float() @ w.float().t())
            try:
                c = ops.linear(a, w)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f"  {dt} M={M} N={N} K={K}: EXC {e}", flush=True)
                continue
            err = (c.float() - ref).abs().max().item()
            tol = 3e-2 if dt == torch.bfloat16 else 4e-3
            ms = timeit(lambda: ops.linear(a, w), iters=10)
            tf = 2 * M * N * K / ms / 1e9
            gbs = (N * K * 2 + M * K * 2 + M * N * 2) / ms / 1e6
            print(f"  {str(dt):15s} M={M:5d} N={N:6d} K={K:5d} maxerr={err:.4f} {'OK ' if err < tol else 'BAD'}"
                  f" {ms*1e3:8.1f} us {tf:7.1f} TF/s {gbs:7.0f} GB/s", flush=True)
    # fp32-out and fused silu*mul
    M, N, K = 512, 1024, 512
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(2 * N, K, device=dev) / math.sqrt(K)).bfloat16()
    c32 = ops.linear(a, w, out_f32=True)
    ref = (a.float() @ w.float().t()).bfloat16().float()
    print("  out_f32 maxerr", (c32 - ref).abs().max().item(), flush=True)
    fused = ops.linear(a, w, epilogue=ops.SPD_EPI_SILU_MUL)
    gu = (a.float() @ w.float().t()).bfloat16().float()
    ref2 = torch.nn.functional.silu(gu[:, :N]) * gu[:, N:]
    print("  silu_mul epilogue maxerr", (fused.float() - ref2).abs().max().item(), flush=True)
    # cuBLAS comparison point
    for (M, N, K) in [(1024, 6144, 4096), (8192, 4096, 4096), (8192, 28672, 4096), (64, 28672, 4096)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        ms1 = timeit(lambda: ops.linear(a, w), iters=10)
        ms2 = timeit(lambda: torch.nn.functional.linear(a, w), iters=10)
        print(f"  M={M} N={N} K={K}: spd {ms1*1e3:.1f} us ({2*M*N*K/ms1/1e9:.0f} TF/s)  cublas {ms2*1e3:.1f} us"
              f" ({2*M*N*K/ms2/1e9:.0f} TF/s)", flush=True)


def check_elem():
    print("== elementwise ==", flush=True)
    for dt in (torch.bfloat16, torch.float16):
        x = torch.randn(99, 4096, device=dev).to(dt)
        w = torch.randn(4096, device=dev).to(dt)
        r = torch.randn(99, 4096, device=dev).to(dt)
        y = ops.rmsnorm(x, w, 1e-5)
        xf = x.float()
        ref = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()).to(dt)
        print(f"  rmsnorm {dt} maxerr", (y.float() - ref.float()).abs().max().item())
        x2, r2 = x.clone(), r.clone()
        ops.fused_add_rmsnorm(x2, r2, w, 1e-5)
        s = x.float() + r.float()
        ref_r = s.to(dt)
        ref_x = (s * torch.rsqrt(s.pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()).to(dt)
        print(f"  fused_add_rmsnorm {dt} maxerr x", (x2.float() - ref_x.float()).abs().max().item(),
              "res", (r2.float() - ref_r.float()).abs().max().item())
        g = torch.randn(77, 2 * 14336, device=dev).to(dt)
        o = ops.silu_and_mul(g)
        refo = (torch.nn.functional.silu(g[:, :14336].float()) * g[:, 14336:].float()).to(dt)
        print(f"  silu_and_mul {dt} maxerr", (o.float() - refo.float()).abs().max().item())
        # rope
        T, Hq, Hk, D = 50, 32, 8, 128
        q = torch.randn(T, Hq * D, device=dev).to(dt)
        k = torch.randn(T, Hk * D, device=dev).to(dt)
        v = torch.randn(T, Hk * D, device=dev).to(dt)
        pos = torch.randint(0, 4096, (T,), device=dev)
        inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2, device=dev).float() / D))
        fr = torch.arange(8192, device=dev).float()[:, None] * inv[None]
        cache = torch.cat([fr.cos(), fr.sin()], -1)

        def rope_ref(x, H):
            xf = x.float().view(T, H, D)
            c, s = cache[pos][:, None, :D // 2], cache[pos][:, None, D // 2:]
            x1, x2 = xf[..., :D // 2], xf[..., D // 2:]
            return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], -1).view(T, H * D).to(dt)

        qr, kr = rope_ref(q, Hq), rope_ref(k, Hk)
        q1, k1 = q.clone(), k.clone()
        ops.apply_rope_with_cos_sin_cache_inplace(pos, q1, k1, D, cache, True)
        print(f"  rope {dt} maxerr q", (q1.float() - qr.float()).abs().max().item(), "k",
              (k1.float() - kr.float()).abs().max().item())
        kp = torch.zeros(1000, Hk, D, device=dev, dtype=dt)
        vp = torch.zeros(1000, Hk, D, device=dev, dtype=dt)
        loc = torch.randperm(999, device=dev)[:T] + 1
        q2, k2 = q.clone(), k.clone()
        ops.rope_kvwrite(pos, q2, k2, v, kp, vp, loc, D, cache, True)
        print("  rope_kvwrite: q", torch.equal(q2, q1), "k", torch.equal(k2, k1), "kpool",
              torch.equal(kp[loc].view(T, -1), k1), "vpool", torch.equal(vp[loc].view(T, -1), v))
    emb = torch.randn(1000, 4096, device=dev).bfloat16()
    ids = torch.randint(0, 1000, (33,), device=dev)
    print("  gather_rows", torch.equal(ops.gather_rows(emb, ids), emb[ids]))
    lg = torch.randn(7, 128256, device=dev)
    print("  argmax", torch.equal(ops.argmax(lg).long(), lg.argmax(-1)))
    pre = torch.tensor([0, 5, 100, 0], device=dev, dtype=torch.int32)
    ext = torch.tensor([7, 1, 300, 64], device=dev, dtype=torch.int32)
    p, st, tr = ops.compute_positions(pre, ext, int(ext.sum()), want_token_req=True)
    refp = torch.cat([torch.arange(a, a + b) for a, b in zip(pre.tolist(), ext.tolist())]).to(dev)
    print("  compute_positions", torch.equal(p, refp), st.tolist(), tr[[0, 7, 8, 307, 308]].tolist())


def make_paged(B, lens, Hk, D, dt, slots_total):
    rtt = torch.zeros(B + 3, max(lens) + 8, dtype=torch.int32, device=dev)
    perm = torch.randperm(slots_total - 1, device=dev)[: sum(lens)] + 1
    rpi = torch.randperm(B + 3, device=dev)[:B]
    o = 0
    for b, L in enumerate(lens):
        rtt[rpi[b], :L] = perm[o:o + L].int()
        o += L
    kp = torch.randn(slots_total, Hk, D, device=dev).to(dt)
    vp = torch.randn(slots_total, Hk, D, device=dev).to(dt)
    return rtt, rpi, kp, vp


def attn_ref(q, kp, vp, slots, Hq, Hk, D, scale, causal_off=None):
    # q [n, Hq, D] for one request, keys slots [L]; query i sees keys <= causal_off + i (or all)
    k = kp[slots.long()].float().repeat_interleave(Hq // Hk, 1)  # [L, Hq, D]
    v = vp[slots.long()].float().repeat_interleave(Hq // Hk, 1)
    s = torch.einsum("nhd,lhd->hnl", q.float(), k) * scale
    if causal_off is not None:
        n, L = q.shape[0], k.shape[0]
        mask = torch.arange(L, device=dev)[None, :] > (causal_off + torch.arange(n, device=dev))[:, None]
        s = s.masked_fill(mask[None], float("-inf"))
    return torch.einsum("hnl,lhd->nhd", s.softmax(-1), v)


def check_attn(impls=("simt",)):
    print("== attention ==", flush=True)
    Hq, Hk, D = 32, 8, 128
    scale = D ** -0.5
    for dt in (torch.bfloat16, torch.float16):
        for impl in impls:
            for lens in ([1], [17, 130, 1, 64, 700], [1500] * 3, [4000], list(range(1, 70, 3))):
                B = len(lens)
                rtt, rpi, kp, vp = make_paged(B, lens, Hk, D, dt, sum(lens) + 50)
                q = torch.randn(B, Hq * D, device=dev).to(dt)
                sl = torch.tensor(lens, device=dev, dtype=torch.int64)
                try:
                    o = ops.attn_decode(q, kp, vp, rtt, rpi, sl, Hq, Hk, D, scale, impl=impl)
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    print(f"  decode[{impl}] {dt} lens={lens[:4]}: EXC {e}")
                    continue
                err = 0.0
                for b, L in enumerate(lens):
                    r = attn_ref(q[b].view(1, Hq, D), kp, vp, rtt[rpi[b], :L], Hq, Hk, D, scale)
                    err = max(err, (o[b].float().view(Hq, D) - r[0]).abs().max().item())
                print(f"  decode[{impl}] {dt} B={B} maxlen={max(lens)} maxerr={err:.5f}", flush=True)
            for (pres, exts) in (([0], [1]), ([0], [128]), ([0, 0], [100, 257]), ([64, 0, 300], [200, 5, 129]),
                                 ([1000], [24]), ([0], [1024])):
                B = len(pres)
                lens = [a + b for a, b in zip(pres, exts)]
                rtt, rpi, kp, vp = make_paged(B, lens, Hk, D, dt, sum(lens) + 50)
                T = sum(exts)
                q = torch.randn(T, Hq * D, device=dev).to(dt)
                pre_t = torch.tensor(pres, device=dev, dtype=torch.int32)
                ext_t = torch.tensor(exts, device=dev, dtype=torch.int32)
                _, st = ops.compute_positions(pre_t, ext_t, T)
                try:
                    o = ops.attn_prefill(q, kp, vp, rtt, rpi, pre_t, ext_t, st, max(exts), Hq, Hk, D, scale, impl=impl)
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    print(f"  prefill[{impl}] {dt} pres={pres}: EXC {e}")
                    continue
                err, off = 0.0, 0
                for b in range(B):
                    r = attn_ref(q[off:off + exts[b]].view(-1, Hq, D), kp, vp, rtt[rpi[b], :lens[b]], Hq, Hk, D,
                                 scale, causal_off=pres[b])
                    err = max(err, (o[off:off + exts[b]].float().view(-1, Hq, D) - r).abs().max().item())
                    off += exts[b]
                print(f"  prefill[{impl}] {dt} pre={pres} ext={exts} maxerr={err:.5f}", flush=True)
    # timings
    for impl in impls:
        for (B, L) in ((64, 1088), (160, 1088), (8, 4096), (1, 8192)):
            lens = [L] * B
            rtt, rpi, kp, vp = make_paged(B, lens, Hk, D, torch.bfloat16, sum(lens) + 50)
            q = torch.randn(B, Hq * D, device=dev).bfloat16()
            sl = torch.tensor(lens, device=dev, dtype=torch.int64)
            ms = timeit(lambda: ops.attn_decode(q, kp, vp, rtt, rpi, sl, Hq, Hk, D, scale, impl=impl))
            by = B * L * Hk * D * 2 * 2 + 2 * B * Hq * D * 2
            print(f"  decode[{impl}] B={B} L={L}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s", flush=True)
        for (B, E, P) in ((1, 1024, 0), (8, 1024, 0), (2, 2048, 2048)):
            lens = [E + P] * B
            rtt, rpi, kp, vp = make_paged(B, lens, Hk, D, torch.bfloat16, sum(lens) + 50)
            T = B * E
            q = torch.randn(T, Hq * D, device=dev).bfloat16()
            pre_t = torch.full((B,), P, device=dev, dtype=torch.int32)
            ext_t = torch.full((B,), E, device=dev, dtype=torch.int32)
            _, st = ops.compute_positions(pre_t, ext_t, T)
            ms = timeit(lambda: ops.attn_prefill(q, kp, vp, rtt, rpi, pre_t, ext_t, st, E, Hq, Hk, D, scale, impl=impl),
                        iters=5)
            fl = 4 * Hq * D * B * E * (P + (E + 1) / 2)
            print(f"  prefill[{impl}] B={B} E={E} P={P}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TF/s", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    impls = tuple(sys.argv[2].split(",")) if len(sys.argv) > 2 else ("simt",)
    t0 = time.time()
    if what in ("elem", "all"):
        check_elem()
    if what in ("gemm", "all"):
        check_gemm()
    if what in ("attn", "all"):
        check_attn(impls)
    print("done in %.1fs" % (time.time() - t0))
