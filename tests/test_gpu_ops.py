"""Per-op parity of the CUDA kernels (called through the C-ABI via semi_pd_b200.ops) with the CPU
oracle.  Shapes / tolerances mirror the reference's kernel tests: sgl-kernel/tests/test_norm.py
(1e-3), test_activation.py (1e-3), test_rotary_embedding.py (1e-2),
test/srt/test_triton_attention_kernels.py (atol 1e-2), test/srt/test_create_kvindices.py (exact).
Integer / index ops are compared bit-exactly."""
import math

import pytest
import torch

from oracle import llama_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DTYPES = [torch.float16, torch.bfloat16]


@pytest.fixture(scope="module")
def ops():
    from semi_pd_b200 import ops as _ops
    return _ops


def _tol(dt, f16, bf16):
    return f16 if dt == torch.float16 else bf16


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bs,hid", [(1, 128), (19, 4096), (99, 8192), (7, 16384), (989, 1024)])
def test_rmsnorm(ops, dtype, bs, hid):
    torch.manual_seed(0)
    x, w, r = torch.randn(bs, hid).to(dtype), torch.randn(hid).to(dtype), torch.randn(bs, hid).to(dtype)
    y = ops.rmsnorm(x.to(DEV), w.to(DEV), 1e-6).cpu()
    # bf16: one ulp (2^-8 relative) when the fp32 value sits on a rounding boundary
    rt = _tol(dtype, 1e-3, 8e-3)
    torch.testing.assert_close(y, O.rmsnorm(x, w, 1e-6), rtol=rt, atol=_tol(dtype, 1e-3, 8e-3))
    xg, rg = x.to(DEV).clone(), r.to(DEV).clone()
    ops.fused_add_rmsnorm(xg, rg, w.to(DEV), 1e-6)
    ox, orr = O.rmsnorm(x, w, 1e-6, r)
    torch.testing.assert_close(xg.cpu(), ox, rtol=rt, atol=_tol(dtype, 1e-3, 8e-3))
    assert torch.equal(rg.cpu(), orr)  # residual_out = (x + residual).to(dtype): exact


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("d", [128, 1024, 14336])
def test_silu_and_mul(ops, dtype, d):
    torch.manual_seed(1)
    x = torch.randn(33, 2 * d).to(dtype)
    y = ops.silu_and_mul(x.to(DEV)).cpu()
    torch.testing.assert_close(y, O.silu_and_mul(x, fp32_math=True), rtol=1e-3, atol=_tol(dtype, 1e-3, 8e-3))
    with pytest.raises(ValueError):
        ops.silu_and_mul(torch.zeros(2, 6, dtype=dtype, device=DEV))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("D,rot,neox,Hq,Hk", [(128, 128, True, 32, 8), (128, 128, False, 8, 8),
                                             (128, 64, True, 4, 1), (64, 64, True, 8, 2)])
def test_rope_and_kv_write(ops, dtype, D, rot, neox, Hq, Hk):
    torch.manual_seed(2)
    T, maxpos = 77, 4096
    cache = O.cos_sin_cache(rot, maxpos, 500000.0)
    pos = torch.randint(0, maxpos, (T,))
    qkv = torch.randn(T, (Hq + 2 * Hk) * D).to(dtype)
    q, k, v = qkv.split([Hq * D, Hk * D, Hk * D], dim=-1)
    qr, kr = O.apply_rope(pos, q.clone(), k.clone(), D, cache, neox, fp32_math=True)
    g = qkv.to(DEV)
    gq, gk, gv = g.split([Hq * D, Hk * D, Hk * D], dim=-1)  # strided views, as in LlamaAttention
    ops.apply_rope_with_cos_sin_cache_inplace(pos.to(DEV), gq, gk, D, cache.to(DEV), neox)
    tol = _tol(dtype, 2e-3, 1.6e-2)
    torch.testing.assert_close(gq.cpu(), qr, rtol=1e-2, atol=tol)
    torch.testing.assert_close(gk.cpu(), kr, rtol=1e-2, atol=tol)
    # fused RoPE + pool scatter: identical q/k, pool rows == rotated k / raw v, other rows untouched
    g2 = qkv.to(DEV)
    q2, k2, v2 = g2.split([Hq * D, Hk * D, Hk * D], dim=-1)
    kp = torch.full((300, Hk, D), 7.0, dtype=dtype, device=DEV)
    vp = torch.full((300, Hk, D), 7.0, dtype=dtype, device=DEV)
    loc = (torch.randperm(299)[:T] + 1).to(DEV)
    ops.rope_kvwrite(pos.to(DEV), q2, k2, v2, kp, vp, loc, D, cache.to(DEV), neox)
    assert torch.equal(q2, gq) and torch.equal(k2, gk)
    assert torch.equal(kp[loc].view(T, -1), gk) and torch.equal(vp[loc].view(T, -1), gv)
    mask = torch.ones(300, dtype=torch.bool, device=DEV)
    mask[loc] = False
    assert (kp[mask] == 7).all() and (vp[mask] == 7).all()
    # set_kv_buffer alone
    kp2, vp2 = torch.zeros_like(kp), torch.zeros_like(vp)
    ops.kv_write(gk, gv, kp2, vp2, loc)
    assert torch.equal(kp2[loc].view(T, -1), gk) and torch.equal(vp2[loc].view(T, -1), gv)


def test_rope_rejects_non_fp32_cache(ops):
    q = torch.zeros(2, 128, dtype=torch.float16, device=DEV)
    with pytest.raises(ValueError):
        ops.apply_rope_with_cos_sin_cache_inplace(torch.zeros(2, dtype=torch.int64, device=DEV), q, q.clone(),
                                                  128, torch.zeros(8, 128, dtype=torch.float16, device=DEV))


def test_gather_argmax_exact(ops):
    torch.manual_seed(3)
    emb = torch.randn(1000, 4096).bfloat16()
    ids = torch.randint(0, 1000, (257,))
    assert torch.equal(ops.gather_rows(emb.to(DEV), ids.to(DEV)).cpu(), emb[ids])
    for B, V in ((1, 128256), (7, 50272), (160, 1024), (3, 13)):
        lg = torch.randn(B, V)
        assert torch.equal(ops.argmax(lg.to(DEV)).cpu().long(), O.LlamaOracle.sample(lg).long())
    tie = torch.zeros(2, 100)
    tie[0, 40] = tie[0, 70] = 5.0
    assert ops.argmax(tie.to(DEV)).cpu().tolist() == [40, 0]  # first maximal index


@pytest.mark.parametrize("bs", [1, 37, 1786])
def test_metadata_exact(ops, bs):
    """positions / kv_indices / req_to_token writes: bit-exact vs the oracle
    (test/srt/test_create_kvindices.py uses the same batch sizes)."""
    torch.manual_seed(bs)
    pre = torch.randint(0, 50, (bs,), dtype=torch.int32)
    ext = torch.randint(1, 40, (bs,), dtype=torch.int32)
    if bs > 2:
        pre[1], ext[2] = 0, 1
    p_ref, s_ref = O.compute_position(pre.tolist(), ext.tolist())
    p, s, tr = ops.compute_positions(pre.to(DEV), ext.to(DEV), int(ext.sum()), want_token_req=True)
    assert torch.equal(p.cpu(), p_ref) and torch.equal(s.cpu(), s_ref.int())
    assert torch.equal(tr.cpu().long(), torch.repeat_interleave(torch.arange(bs), ext.long()))
    seq = (pre + ext).long()
    assert torch.equal(ops.clamp_position(seq.to(DEV)).cpu(), O.clamp_position(seq))
    assert ops.clamp_position(torch.zeros(3, dtype=torch.int64, device=DEV)).cpu().tolist() == [0, 0, 0]
    # write_req_to_token + kv_indices round trip
    ctx = 96
    rtt = torch.zeros(bs + 5, ctx, dtype=torch.int32)
    reqs = torch.randperm(bs + 5)[:bs]
    loc = torch.randperm(int(ext.sum()) + 10)[: int(ext.sum())] + 1
    O.write_req_to_token(rtt, reqs.tolist(), pre.tolist(), seq.tolist(), loc)
    rtt_g = torch.zeros(bs + 5, ctx, dtype=torch.int32, device=DEV)
    ops.write_req_to_token_pool(rtt_g, reqs.to(DEV), pre.long().to(DEV), seq.to(DEV), ext.long().to(DEV), loc.to(DEV))
    assert torch.equal(rtt_g.cpu(), rtt)
    indptr, idx = O.kv_indices(rtt, reqs.tolist(), ext.tolist(), pre.tolist())
    out = torch.empty(int(ext.sum()), dtype=torch.int32, device=DEV)
    ops.create_kv_indices(rtt_g, reqs.to(DEV), ext.to(DEV), indptr.to(DEV), pre.to(DEV), out)
    assert torch.equal(out.cpu(), idx)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(1, 256, 128), (5, 384, 256), (16, 4096, 4096), (64, 6144, 4096),
                                   (130, 512, 1024), (256, 1024, 14336), (257, 1024, 512),
                                   (1000, 4096, 4096), (1024, 6144, 4096), (384, 28672, 1024),
                                   (8, 128256, 512), (2048, 264, 1032),
                                   # CTA-pair (cta_group::2) kernel: exact tiles, N/K tails, M tails
                                   # incl. a pair tile whose second CTA is entirely out of range
                                   (512, 256, 64), (768, 1000, 4096), (2500, 776, 328), (2100, 264, 1032),
                                   (4096, 2048, 512)])
def test_gemm(ops, dtype, M, N, K):
    """F.linear parity: the oracle is torch CPU matmul of the same 16-bit operands."""
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K) * 0.5).to(dtype)
    w = (torch.randn(N, K) / math.sqrt(K)).to(dtype)
    ref = a.float() @ w.float().t()
    c = ops.linear(a.to(DEV), w.to(DEV)).cpu().float()
    assert (c - ref).abs().max().item() < _tol(dtype, 3e-3, 2.5e-2)
    c32 = ops.linear(a.to(DEV), w.to(DEV), out_f32=True).cpu()
    assert c32.dtype == torch.float32 and torch.equal(c32, c)  # == (a @ w.T).to(dtype).float()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,N,K", [(33, 256, 512), (40, 4096, 4096), (64, 4096, 14336), (100, 6144, 4096),
                                   (128, 4096, 4096), (160, 4096, 14336), (192, 28672, 1024),
                                   (200, 1000, 2048), (256, 6144, 4096), (96, 520, 576), (17, 4096, 4096),
                                   (32, 6144, 4096), (130, 128256, 512)])
def test_gemm_decode_regime(ops, dtype, M, N, K):
    """Swap-AB decode regime: CTA-pair kernel (32 < M <= 256) and single-CTA kernel, whole tiles
    and stream-K with the owner fix-up, on the whole GPU and on small / odd SM budgets (the decode
    partition).  Parity with the fp32 matmul of the same 16-bit operands; every schedule must
    produce bit-identical results run to run."""
    torch.manual_seed(M * 7 + N + K)
    a = (torch.randn(M, K) * 0.5).to(dtype)
    w = (torch.randn(N, K) / math.sqrt(K)).to(dtype)
    ref = a.float() @ w.float().t()
    ad, wd = a.to(DEV), w.to(DEV)
    try:
        for budget in (0, 76, 20, 7, 2):
            ops.set_sm_budget(budget)
            c = ops.linear(ad, wd)
            err = (c.cpu().float() - ref).abs().max().item()
            assert err < _tol(dtype, 3e-3, 2.5e-2), (budget, err)
            for _ in range(3):  # flags / partials are re-armed: replays must match bit for bit
                assert torch.equal(ops.linear(ad, wd), c), budget
    finally:
        ops.set_sm_budget(0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_fused_silu_mul_and_strides(ops, dtype):
    torch.manual_seed(9)
    M, N, K = 700, 1536, 1024
    a = (torch.randn(M, K) * 0.5).to(dtype)
    w = (torch.randn(2 * N, K) / math.sqrt(K)).to(dtype)
    gu = torch.nn.functional.linear(a.float(), w.float()).to(dtype)
    ref = O.silu_and_mul(gu, fp32_math=True).float()
    out = ops.linear(a.to(DEV), w.to(DEV), epilogue=ops.SPD_EPI_SILU_MUL).cpu().float()
    assert (out - ref).abs().max().item() < _tol(dtype, 4e-3, 3e-2)
    # same through the CTA-pair kernel (M a multiple of 256) and through the single-CTA kernel
    for M2 in (1024, 640, 100, 8):  # pair kernel, single-CTA kernel, and the small-M fused form
        a2 = (torch.randn(M2, K) * 0.5).to(dtype)
        gu2 = torch.nn.functional.linear(a2.float(), w.float()).to(dtype)
        ref2 = O.silu_and_mul(gu2, fp32_math=True).float()
        out2 = ops.linear(a2.to(DEV), w.to(DEV), epilogue=ops.SPD_EPI_SILU_MUL).cpu().float()
        assert (out2 - ref2).abs().max().item() < _tol(dtype, 4e-3, 3e-2)
    # strided A (a column slice of a wider activation) and determinism of split-K
    wide = torch.randn(64, 3 * K).to(dtype).to(DEV)
    w2 = (torch.randn(4096, K) / math.sqrt(K)).to(dtype).to(DEV)
    c1 = ops.linear(wide[:, K:2 * K], w2)
    c2 = ops.linear(wide[:, K:2 * K].contiguous(), w2)
    assert torch.equal(c1, c2) and torch.equal(c1, ops.linear(wide[:, K:2 * K], w2))


def test_gemm_argument_errors(ops):
    a = torch.zeros(4, 100, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(8, 100, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        ops.linear(a, w)
    with pytest.raises(RuntimeError, match="dtype mismatch"):
        ops.linear(torch.zeros(4, 128, dtype=torch.float16, device=DEV),
                   torch.zeros(8, 128, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.linear(torch.zeros(4, 128, dtype=torch.bfloat16), torch.zeros(8, 128, dtype=torch.bfloat16))


def _paged(lens, Hk, D, dtype, extra=40, seed=0):
    g = torch.Generator().manual_seed(seed)
    B = len(lens)
    slots = sum(lens) + extra
    rtt = torch.zeros(B + 3, max(lens) + 8, dtype=torch.int32)
    rpi = torch.randperm(B + 3, generator=g)[:B]
    perm = torch.randperm(slots - 1, generator=g)[: sum(lens)] + 1
    o = 0
    for b, L in enumerate(lens):
        rtt[rpi[b], :L] = perm[o:o + L].int()
        o += L
    kp = torch.randn(slots, Hk, D, generator=g).to(dtype)
    vp = torch.randn(slots, Hk, D, generator=g).to(dtype)
    return rtt, rpi, kp, vp


def _paged_runs(lens, Hk, D, dtype, seed=0, mixed=False, run=128):
    """Page tables as the run allocator lays them out (memory_pool.py): positions [k*128, k*128+128)
    of a request are 128 consecutive slots, runs in random order across the pool.  `mixed` also
    produces what a foreign allocator might: runs that start at arbitrary (unaligned) slots, tiles
    with two entries swapped, and tiles that are consecutive except for their last entry — the
    kernels must tell all of these apart from a TMA-able tile on the device."""
    g = torch.Generator().manual_seed(seed)
    B = len(lens)
    n_runs = sum(-(-L // run) for L in lens) + 3
    slots = n_runs * run + 1 + (64 if mixed else 0)
    shift = 37 if mixed else 0
    rtt = torch.zeros(B + 3, max(lens) + 8, dtype=torch.int32)
    rpi = torch.randperm(B + 3, generator=g)[:B]
    order = torch.randperm(n_runs, generator=g).tolist()
    for b, L in enumerate(lens):
        for k in range(-(-L // run)):
            r = order.pop()
            n = min(run, L - k * run)
            ids = torch.arange(n, dtype=torch.int32) + (1 + shift + r * run)
            if mixed and n == run and (b + k) % 3 == 1:
                ids[[5, 77]] = ids[[77, 5]]                      # same slots, two swapped
            elif mixed and n == run and (b + k) % 3 == 2:
                ids[-1], ids[-2] = ids[-2].clone(), ids[-1].clone()  # consecutive but for the tail
            rtt[rpi[b], k * run:k * run + n] = ids
    kp = torch.randn(slots, Hk, D, generator=g).to(dtype)
    vp = torch.randn(slots, Hk, D, generator=g).to(dtype)
    return rtt, rpi, kp, vp


@pytest.mark.parametrize("layout", ["runs", "mixed"])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("Hq,Hk", [(32, 8), (8, 1)])
@pytest.mark.parametrize("lens", [[128], [17, 130, 1, 64, 700], [1500, 128, 129, 256, 384], [4000]],
                         ids=["tile", "ragged", "long", "4k"])
def test_attn_decode_run_layouts(ops, layout, dtype, Hq, Hk, lens):
    """The TMA-staged K/V path (consecutive slots) and its on-device fallback decision."""
    D = 128
    rtt, rpi, kp, vp = _paged_runs(lens, Hk, D, dtype, seed=len(lens), mixed=layout == "mixed")
    torch.manual_seed(5)
    q = torch.randn(len(lens), Hq, D).to(dtype)
    ref = O.attn_decode(q, kp, vp, rtt, rpi, lens, D ** -0.5, fp32=True)
    sl = torch.tensor(lens, dtype=torch.int64, device=DEV)
    args = (q.to(DEV).view(len(lens), -1), kp.to(DEV), vp.to(DEV), rtt.to(DEV), rpi.to(DEV), sl, Hq, Hk, D, D ** -0.5)
    out = ops.attn_decode(*args, impl="tc")
    err = (out.cpu().float().view(-1, Hq, D) - ref.float()).abs().max().item()
    assert err < _tol(dtype, 2e-3, 1.6e-2), err
    try:  # partition-sized grid: every CTA walks several items, TMA and gathered tiles interleaved
        ops.set_sm_budget(8)
        out8 = ops.attn_decode(*args, impl="tc")  # (a different KV split: not bit-comparable)
        err8 = (out8.cpu().float().view(-1, Hq, D) - ref.float()).abs().max().item()
        assert err8 < _tol(dtype, 2e-3, 1.6e-2), err8
    finally:
        ops.set_sm_budget(0)


@pytest.mark.parametrize("layout", ["runs", "mixed"])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("Hq,Hk", [(32, 8), (8, 1)])
@pytest.mark.parametrize("pre,ext", [([0], [128]), ([0, 0], [100, 257]), ([64, 0, 300], [200, 5, 129]),
                                     ([0], [1024]), ([128, 256], [128, 512]), ([0, 40, 0], [700, 1024, 333])],
                         ids=["tile", "two", "prefix", "1k", "aligned-prefix", "multi"])
def test_attn_prefill_run_layouts(ops, layout, dtype, Hq, Hk, pre, ext):
    D = 128
    lens = [a + b for a, b in zip(pre, ext)]
    rtt, rpi, kp, vp = _paged_runs(lens, Hk, D, dtype, seed=sum(lens), mixed=layout == "mixed")
    torch.manual_seed(6)
    T = sum(ext)
    q = torch.randn(T, Hq, D).to(dtype)
    ref = O.attn_extend(q, kp, vp, rtt, rpi, lens, pre, ext, D ** -0.5, fp32=True)
    pre_t = torch.tensor(pre, dtype=torch.int32, device=DEV)
    ext_t = torch.tensor(ext, dtype=torch.int32, device=DEV)
    _, st = ops.compute_positions(pre_t, ext_t, T)
    args = (q.to(DEV).view(T, -1), kp.to(DEV), vp.to(DEV), rtt.to(DEV), rpi.to(DEV), pre_t, ext_t, st, max(ext),
            Hq, Hk, D, D ** -0.5)
    out = ops.attn_prefill(*args, impl="tc")
    err = (out.cpu().float().view(-1, Hq, D) - ref.float()).abs().max().item()
    assert err < _tol(dtype, 2e-3, 1.6e-2), err
    ops.set_sm_budget(8)
    try:
        assert torch.equal(ops.attn_prefill(*args, impl="tc"), out)  # the schedule must not change a bit
    finally:
        ops.set_sm_budget(0)


@pytest.mark.parametrize("impl", ["tc", "simt"])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("Hq,Hk", [(32, 8), (8, 1), (4, 4), (6, 2)])
@pytest.mark.parametrize("lens", [[1], [17, 130, 1, 64, 700], [1500, 128, 129], list(range(1, 70, 3)), [4000]],
                         ids=["one", "ragged", "long", "many", "4k"])
def test_attn_decode(ops, impl, dtype, Hq, Hk, lens):
    D = 128
    rtt, rpi, kp, vp = _paged(lens, Hk, D, dtype, seed=len(lens))
    torch.manual_seed(5)
    q = torch.randn(len(lens), Hq, D).to(dtype)
    ref = O.attn_decode(q, kp, vp, rtt, rpi, lens, D ** -0.5, fp32=True)
    out = ops.attn_decode(q.to(DEV).view(len(lens), -1), kp.to(DEV), vp.to(DEV), rtt.to(DEV), rpi.to(DEV),
                          torch.tensor(lens, dtype=torch.int64, device=DEV), Hq, Hk, D, D ** -0.5, impl=impl)
    err = (out.cpu().float().view(-1, Hq, D) - ref.float()).abs().max().item()
    assert err < _tol(dtype, 2e-3, 1.6e-2), err  # bf16: one output ulp at |o| < 4


@pytest.mark.parametrize("impl", ["tc", "simt"])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("Hq,Hk", [(32, 8), (8, 1), (4, 4), (6, 2)])
@pytest.mark.parametrize("pre,ext", [([0], [1]), ([0], [128]), ([0, 0], [100, 257]),
                                     ([64, 0, 300], [200, 5, 129]), ([1000], [24]), ([0], [1024]),
                                     ([127, 128, 129, 1], [1, 1, 127, 130]), ([0, 40, 0], [700, 1024, 333])],
                         ids=["1tok", "tile", "two", "prefix", "longprefix", "1k", "edges", "multi"])
def test_attn_prefill(ops, impl, dtype, Hq, Hk, pre, ext):
    D = 128
    lens = [a + b for a, b in zip(pre, ext)]
    rtt, rpi, kp, vp = _paged(lens, Hk, D, dtype, seed=sum(lens))
    torch.manual_seed(6)
    T = sum(ext)
    q = torch.randn(T, Hq, D).to(dtype)
    ref = O.attn_extend(q, kp, vp, rtt, rpi, lens, pre, ext, D ** -0.5, fp32=True)
    pre_t = torch.tensor(pre, dtype=torch.int32, device=DEV)
    ext_t = torch.tensor(ext, dtype=torch.int32, device=DEV)
    _, st = ops.compute_positions(pre_t, ext_t, T)
    out = ops.attn_prefill(q.to(DEV).view(T, -1), kp.to(DEV), vp.to(DEV), rtt.to(DEV), rpi.to(DEV), pre_t, ext_t,
                           st, max(ext), Hq, Hk, D, D ** -0.5, impl=impl)
    err = (out.cpu().float().view(-1, Hq, D) - ref.float()).abs().max().item()
    assert err < _tol(dtype, 2e-3, 1.6e-2), err
    if impl == "tc" and len(ext) > 1:
        # the persistent kernel on a small SM budget: every CTA walks many items back to back
        ops.set_sm_budget(8)
        try:
            out8 = ops.attn_prefill(q.to(DEV).view(T, -1), kp.to(DEV), vp.to(DEV), rtt.to(DEV), rpi.to(DEV), pre_t,
                                    ext_t, st, max(ext), Hq, Hk, D, D ** -0.5, impl=impl)
        finally:
            ops.set_sm_budget(0)
        assert torch.equal(out8, out)  # the schedule must not change a single bit


def test_attention_on_reference_golden_vectors(ops, golden):
    """The reference's own TorchNativeAttnBackend outputs (tests/golden) through the CUDA kernels."""
    n = 0
    for c in golden["attention"]:
        D, Hq, Hk = c["D"], c["Hq"], c["Hk"]
        dt = torch.bfloat16 if "bfloat16" in c["dtype"] else torch.float16
        impl = "tc" if D == 128 else "simt"  # the tcgen05 path is specialised for head_dim 128
        pre, ext = c["prefix"], c["extend"]
        seqs = [a + b for a, b in zip(pre, ext)]
        kp, vp = c["k_pool0"].to(dt).to(DEV), c["v_pool0"].to(dt).to(DEV)
        kp[c["out_cache_loc"].to(DEV)] = c["k"].to(dt).to(DEV).view(-1, Hk, D)
        vp[c["out_cache_loc"].to(DEV)] = c["v"].to(dt).to(DEV).view(-1, Hk, D)
        pre_t = torch.tensor(pre, dtype=torch.int32, device=DEV)
        ext_t = torch.tensor(ext, dtype=torch.int32, device=DEV)
        _, st = ops.compute_positions(pre_t, ext_t, sum(ext))
        rtt, rpi = c["req_to_token"].to(DEV), c["req_pool_indices"].to(DEV)
        out = ops.attn_prefill(c["q"].to(dt).to(DEV), kp, vp, rtt, rpi, pre_t, ext_t, st, max(ext), Hq, Hk, D,
                               D ** -0.5, impl=impl)
        tol = 2e-2 if (dt == torch.bfloat16 or "float32" in c["dtype"]) else 4e-3
        assert (out.cpu().float() - c["out_extend"].float()).abs().max().item() < tol
        kp[c["out_cache_loc_d"].to(DEV)] = c["kd"].to(dt).to(DEV).view(-1, Hk, D)
        vp[c["out_cache_loc_d"].to(DEV)] = c["vd"].to(dt).to(DEV).view(-1, Hk, D)
        outd = ops.attn_decode(c["qd"].to(dt).to(DEV), kp, vp, rtt, rpi,
                               torch.tensor([s + 1 for s in seqs], dtype=torch.int64, device=DEV), Hq, Hk, D,
                               D ** -0.5, impl=impl)
        assert (outd.cpu().float() - c["out_decode"].float()).abs().max().item() < tol
        n += 1
    assert n == len(golden["attention"])


def test_attn_properties_full_size(ops):
    """Size-independent properties at BASELINE sizes (Llama-3-8B heads, B = 64, S ≈ 1088):
    (1) tcgen05 == SIMT kernel on the same inputs; (2) softmax convexity: every output lies inside
    [min V, max V] of its kv head; (3) permuting the slot assignment (a different page table over
    the same logical sequence) does not change the result."""
    Hq, Hk, D, B = 32, 8, 128, 64
    lens = [1088 - (i % 7) * 13 for i in range(B)]
    rtt, rpi, kp, vp = _paged(lens, Hk, D, torch.bfloat16, seed=11)
    q = torch.randn(B, Hq * D).bfloat16()
    args = (torch.tensor(lens, dtype=torch.int64, device=DEV), Hq, Hk, D, D ** -0.5)
    a = ops.attn_decode(q.to(DEV), kp.to(DEV), vp.to(DEV), rtt.to(DEV), rpi.to(DEV), *args, impl="tc")
    b = ops.attn_decode(q.to(DEV), kp.to(DEV), vp.to(DEV), rtt.to(DEV), rpi.to(DEV), *args, impl="simt")
    assert (a.float() - b.float()).abs().max().item() < 1e-2
    assert a.float().max() <= vp.float().max() + 1e-2 and a.float().min() >= vp.float().min() - 1e-2
    # relocate every token to a new slot
    slots = kp.shape[0]
    perm = torch.randperm(slots - 1) + 1
    mapping = torch.zeros(slots, dtype=torch.long)
    mapping[1:] = perm
    kp2, vp2 = torch.zeros_like(kp), torch.zeros_like(vp)
    kp2[mapping[1:]] = kp[1:]
    vp2[mapping[1:]] = vp[1:]
    rtt2 = mapping[rtt.long()].int()
    c = ops.attn_decode(q.to(DEV), kp2.to(DEV), vp2.to(DEV), rtt2.to(DEV), rpi.to(DEV), *args, impl="tc")
    assert torch.equal(a, c)
    # (4) the same logical sequences re-housed in 128-slot runs (TMA-staged tiles) give the same
    # bits as the scattered pool (gathered tiles): both paths fill smem identically
    R = 128
    runs_per = [-(-L // R) for L in lens]
    order = torch.randperm(sum(runs_per)).tolist()
    kp3 = torch.zeros(sum(runs_per) * R + 1, Hk, D, dtype=kp.dtype)
    vp3 = torch.zeros_like(kp3)
    rtt3 = torch.zeros_like(rtt)
    for bi, L in enumerate(lens):
        row = int(rpi[bi])
        for k in range(runs_per[bi]):
            n = min(R, L - k * R)
            ids = torch.arange(n) + 1 + order.pop() * R
            old = rtt[row, k * R:k * R + n].long()
            kp3[ids], vp3[ids] = kp[old], vp[old]
            rtt3[row, k * R:k * R + n] = ids.int()
    d = ops.attn_decode(q.to(DEV), kp3.to(DEV), vp3.to(DEV), rtt3.to(DEV), rpi.to(DEV), *args, impl="tc")
    assert torch.equal(a, d)
