"""Pins the CPU oracle (oracle/llama_oracle.py) against (a) outputs of the reference's own
modules recorded by tests/golden/make_golden.py and (b) the closed-form torch references of the
reference's kernel unit tests (sgl-kernel/tests/test_norm.py, test_activation.py,
test_rotary_embedding.py; test/srt/test_create_kvindices.py)."""
import torch

from oracle import llama_oracle as O


def test_positions_match_reference(golden):
    for c in golden["positions"]:
        p, s = O.compute_position(c["prefix"], c["extend"])
        assert torch.equal(p, c["positions"])
        assert torch.equal(s.to(c["start"].dtype), c["start"])


def test_rope_native_matches_reference(golden):
    for c in golden["rope"]:
        cache = O.cos_sin_cache(c["rot"], c["maxpos"], c["base"])
        assert torch.equal(cache[c["pos"]], c["cache_rows"])
        q, k = O.apply_rope(c["pos"], c["q"].clone(), c["k"].clone(), c["D"], cache, c["neox"],
                            fp32_math=False)
        assert torch.equal(q, c["q_out"]) and torch.equal(k, c["k_out"])
        # fp32 inputs: the CUDA-semantics mode is the same arithmetic
        q2, k2 = O.apply_rope(c["pos"], c["q"].clone(), c["k"].clone(), c["D"], cache, c["neox"],
                              fp32_math=True)
        assert torch.equal(q2, c["q_out"]) and torch.equal(k2, c["k_out"])


def test_llama3_rope_scaling_matches_reference(golden):
    c = golden["llama3_cache"]
    cache = O.cos_sin_cache(128, 16384, 500000, c["scaling"])
    assert torch.equal(cache[c["rows"]], c["values"])


def test_attention_matches_reference(golden):
    for c in golden["attention"]:
        Hq, Hk, D = c["Hq"], c["Hk"], c["D"]
        pres, exts = c["prefix"], c["extend"]
        seqs = [a + b for a, b in zip(pres, exts)]
        kp, vp = c["k_pool0"].clone(), c["v_pool0"].clone()
        kp[c["out_cache_loc"]] = c["k"].view(-1, Hk, D)
        vp[c["out_cache_loc"]] = c["v"].view(-1, Hk, D)
        out = O.attn_extend(c["q"].view(-1, Hq, D), kp, vp, c["req_to_token"], c["req_pool_indices"],
                            seqs, pres, exts, D ** -0.5)
        assert torch.equal(out.reshape(-1, Hq * D), c["out_extend"]), c["dtype"]
        kp[c["out_cache_loc_d"]] = c["kd"].view(-1, Hk, D)
        vp[c["out_cache_loc_d"]] = c["vd"].view(-1, Hk, D)
        outd = O.attn_decode(c["qd"].view(-1, Hq, D), kp, vp, c["req_to_token"], c["req_pool_indices"],
                             [s + 1 for s in seqs], D ** -0.5)
        assert torch.equal(outd.reshape(-1, Hq * D), c["out_decode"]), c["dtype"]


def test_rmsnorm_closed_form():
    # sgl-kernel/tests/test_norm.py:8-50 (llama_rms_norm / fused_add_rms_norm), rtol=atol=1e-3
    torch.manual_seed(0)
    for bs, hid in ((1, 111), (19, 4096), (99, 8192)):
        for dt in (torch.float16, torch.bfloat16):
            x, w, r = torch.randn(bs, hid).to(dt), torch.randn(hid).to(dt), torch.randn(bs, hid).to(dt)
            xf = x.float()
            ref = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()).to(dt)
            torch.testing.assert_close(O.rmsnorm(x, w, 1e-6), ref, rtol=1e-3, atol=1e-3)
            s = x.float() + r.float()
            ref_x = (s * torch.rsqrt(s.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()).to(dt)
            ox, orr = O.rmsnorm(x, w, 1e-6, r)
            torch.testing.assert_close(ox, ref_x, rtol=1e-3, atol=1e-3)
            torch.testing.assert_close(orr, s.to(dt), rtol=1e-3, atol=1e-3)


def test_silu_and_mul_closed_form():
    # sgl-kernel/tests/test_activation.py:8-16: silu(x[:d]) * x[d:], 1e-3
    torch.manual_seed(1)
    x = torch.randn(17, 2 * 512).to(torch.float16)
    ref = torch.nn.functional.silu(x[:, :512].float()) * x[:, 512:].float()
    torch.testing.assert_close(O.silu_and_mul(x).float(), ref, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(O.silu_and_mul(x, fp32_math=True).float(), ref, rtol=1e-3, atol=1e-3)


def test_kv_indices_and_write_req_to_token():
    # test/srt/test_create_kvindices.py:41-66 (exact); schedule_batch.py:949-961
    torch.manual_seed(2)
    rtt = torch.zeros(8, 64, dtype=torch.int32)
    reqs, pre, seq = [5, 1, 3], [0, 4, 10], [7, 9, 30]
    loc = torch.randperm(200)[: sum(s - p for s, p in zip(seq, pre))] + 1
    O.write_req_to_token(rtt, reqs, pre, seq, loc)
    assert torch.equal(rtt[5, :7].long(), loc[:7])
    assert torch.equal(rtt[1, 4:9].long(), loc[7:12])
    indptr, idx = O.kv_indices(rtt, reqs, [s - p for s, p in zip(seq, pre)], pre)
    assert torch.equal(idx.long(), loc)
    assert indptr.tolist() == [0, 7, 12, 32]


def test_model_native_close_to_fp32():
    """The two oracle modes agree to bf16/fp16 rounding on a tiny model (sanity of the wiring)."""
    cfg = O.OracleLlamaConfig.tiny()
    for dt, tol in ((torch.float16, 2e-2), (torch.bfloat16, 1.5e-1)):
        w = O.make_weights(cfg, dt)
        outs = []
        for mode in ("native", "fp32"):
            m = O.LlamaOracle(cfg, w, kv_slots=256, max_reqs=4, max_ctx=128, compute=mode)
            s = O.OracleSession(m, 256, 4)
            rids, logits, _ = s.extend([[1, 2, 3, 4, 5, 6, 7], [9, 8, 7]])
            tok = O.LlamaOracle.sample(logits).tolist()
            l2, _ = s.decode(rids, tok)
            outs.append(torch.cat([logits, l2]))
        assert (outs[0] - outs[1]).abs().max().item() < tol


def test_chunked_prefill_equals_one_shot():
    """extend in two chunks (prefix > 0 on the second) == one extend (torch_native_backend semantics)."""
    cfg = O.OracleLlamaConfig.tiny()
    w = O.make_weights(cfg, torch.float32)
    prompt = list(range(3, 40))
    m1 = O.LlamaOracle(cfg, w, 256, 4, 128, "fp32")
    _, l_full, _ = O.OracleSession(m1, 256, 4).extend([prompt])
    m2 = O.LlamaOracle(cfg, w, 256, 4, 128, "fp32")
    s2 = O.OracleSession(m2, 256, 4, seed=3)
    rids, _, _ = s2.extend([prompt[:20]])
    _, l_chunk, _ = s2.extend([prompt[20:]], rids)
    torch.testing.assert_close(l_full, l_chunk, rtol=1e-4, atol=1e-4)
