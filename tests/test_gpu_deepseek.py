"""DeepSeek-V2-Lite path (BASELINE configs[3]) on the GPU: absorbed MLA attention over the latent
pool, block-diagonal / grouped GEMMs, MoE routing + expert dispatch, and the whole model against
oracle/deepseek_oracle.py (the restatement of the reference's models/deepseek_v2.py, pinned to
transformers' DeepseekV2ForCausalLM by tests/test_deepseek_oracle.py — "parity unpinned" against
the reference itself, whose module does not import in this image)."""
import math

import pytest
import torch

from oracle import deepseek_oracle as D

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DTYPES = [torch.float16, torch.bfloat16]


@pytest.fixture(scope="module")
def ops():
    from semi_pd_b200 import ops as _ops
    return _ops


def _tol(dt, f16, bf16):
    return f16 if dt == torch.float16 else bf16


# ------------------------------------------------------------------------------------------------
# ops
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("H", [16, 4, 32])
@pytest.mark.parametrize("lens", [[1], [17, 130, 1, 64, 700], [1500, 33], list(range(1, 40, 3))],
                         ids=["one", "ragged", "long", "many"])
def test_attn_mla_decode(ops, dtype, H, lens):
    R, Dr = 512, 64
    g = torch.Generator().manual_seed(len(lens) + H)
    B = len(lens)
    slots = sum(lens) + 50
    pool = (torch.randn(slots, R + Dr, generator=g) * 0.5).to(dtype)
    rtt = torch.zeros(B + 2, max(lens) + 8, dtype=torch.int32)
    rpi = torch.randperm(B + 2, generator=g)[:B]
    perm = torch.randperm(slots - 1, generator=g)[: sum(lens)] + 1
    o = 0
    for b, L in enumerate(lens):
        rtt[rpi[b], :L] = perm[o:o + L].int()
        o += L
    q = (torch.randn(B, H, R + Dr, generator=g) * 0.5).to(dtype)
    scale = 192 ** -0.5
    ref = torch.empty(B, H, R)
    for b, L in enumerate(lens):
        rows = pool[rtt[rpi[b], :L].long()].float()
        sc = torch.einsum("hl,kl->hk", q[b].float(), rows) * scale
        ref[b] = sc.softmax(-1) @ rows[:, :R]
    ws = torch.zeros(ops.lib.spd_attn_mla_workspace_bytes(H, R), dtype=torch.uint8, device=DEV)
    out = ops.attn_mla(q.to(DEV).view(B, -1), pool.to(DEV), rtt.to(DEV), rpi.to(DEV), H, R, Dr, scale, ws,
                       seq_lens=torch.tensor(lens, dtype=torch.int64, device=DEV))
    err = (out.cpu().float().view(B, H, R) - ref).abs().max().item()
    assert err < _tol(dtype, 2e-3, 1.6e-2), err


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("pre,ext", [([0], [1]), ([0, 0], [100, 257]), ([64, 0, 300], [200, 5, 129]), ([1000], [24])],
                         ids=["1tok", "two", "prefix", "longprefix"])
def test_attn_mla_extend(ops, dtype, pre, ext):
    H, R, Dr = 16, 512, 64
    g = torch.Generator().manual_seed(sum(ext))
    B = len(ext)
    lens = [a + b for a, b in zip(pre, ext)]
    slots = sum(lens) + 50
    pool = (torch.randn(slots, R + Dr, generator=g) * 0.5).to(dtype)
    rtt = torch.zeros(B, max(lens) + 8, dtype=torch.int32)
    perm = torch.randperm(slots - 1, generator=g)[: sum(lens)] + 1
    o = 0
    for b, L in enumerate(lens):
        rtt[b, :L] = perm[o:o + L].int()
        o += L
    T = sum(ext)
    q = (torch.randn(T, H, R + Dr, generator=g) * 0.5).to(dtype)
    scale = 192 ** -0.5
    ref = torch.empty(T, H, R)
    s0 = 0
    for b in range(B):
        rows = pool[rtt[b, :lens[b]].long()].float()
        sc = torch.einsum("qhl,kl->hqk", q[s0:s0 + ext[b]].float(), rows) * scale
        mask = torch.full((ext[b], lens[b]), float("-inf")).triu(pre[b] + 1)
        ref[s0:s0 + ext[b]] = torch.einsum("hqk,kr->qhr", (sc + mask).softmax(-1), rows[:, :R])
        s0 += ext[b]
    pre_t = torch.tensor(pre, dtype=torch.int32, device=DEV)
    ext_t = torch.tensor(ext, dtype=torch.int32, device=DEV)
    _, st = ops.compute_positions(pre_t, ext_t, T)
    ws = torch.zeros(ops.lib.spd_attn_mla_workspace_bytes(H, R), dtype=torch.uint8, device=DEV)
    out = ops.attn_mla(q.to(DEV).view(T, -1), pool.to(DEV), rtt.to(DEV), torch.arange(B, device=DEV), H, R, Dr, scale,
                       ws, prefix_lens=pre_t, extend_start_loc=st)
    err = (out.cpu().float().view(T, H, R) - ref).abs().max().item()
    assert err < _tol(dtype, 2e-3, 1.6e-2), err


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M", [1, 40, 300])
def test_block_diagonal_gemm(ops, dtype, M):
    """The two absorbed-MLA BMMs as block-diagonal GEMMs, incl. strided input and output groups."""
    torch.manual_seed(M)
    H, Dn, Dr, R, Dv = 16, 128, 64, 512, 128
    q = (torch.randn(M, H * (Dn + Dr) + 576) * 0.5).to(dtype)       # the fused q / kv_a GEMM output
    w_kc_t = (torch.randn(H * R, Dn) / math.sqrt(Dn)).to(dtype)
    out = torch.zeros(M, H * (R + Dr), dtype=dtype, device=DEV)
    ops.linear_block_diag(q.to(DEV)[:, : H * (Dn + Dr)], w_kc_t.to(DEV), H, Dn, R, Dn + Dr, out=out,
                          out_group_stride=R + Dr)
    ref = torch.einsum("mhd,hrd->mhr", q[:, : H * (Dn + Dr)].float().view(M, H, Dn + Dr)[:, :, :Dn],
                       w_kc_t.float().view(H, R, Dn))
    got = out.cpu().float().view(M, H, R + Dr)
    assert (got[:, :, :R] - ref).abs().max().item() < _tol(dtype, 4e-3, 3e-2)
    assert got[:, :, R:].abs().max().item() == 0.0               # the rope slots are left alone
    o_lat = (torch.randn(M, H * R) * 0.5).to(dtype)
    w_vc = (torch.randn(H * Dv, R) / math.sqrt(R)).to(dtype)
    got2 = ops.linear_block_diag(o_lat.to(DEV), w_vc.to(DEV), H, R, Dv, R).cpu().float()
    ref2 = torch.einsum("mhr,hdr->mhd", o_lat.float().view(M, H, R), w_vc.float().view(H, Dv, R)).reshape(M, -1)
    assert (got2 - ref2).abs().max().item() < _tol(dtype, 4e-3, 3e-2)


@pytest.mark.parametrize("T,E,k", [(1, 64, 6), (37, 64, 6), (700, 64, 6), (200, 8, 3), (5, 160, 8)])
def test_moe_route_and_align(ops, T, E, k):
    torch.manual_seed(T + E)
    logits = torch.randn(T, E).bfloat16()
    ids, w = ops.moe_route(logits.to(DEV), k, scale=2.5)
    probs = logits.float().softmax(-1)
    # greedy top-k as the reference's topk_softmax kernel: repeated arg-max, ties -> the smaller
    # expert id (16-bit logits DO tie; torch.topk leaves the order of ties unspecified)
    rest, ti, tw = logits.float().clone(), [], []
    for _ in range(k):
        e = rest.argmax(-1)                      # first maximal index
        ti.append(e)
        tw.append(probs.gather(1, e[:, None])[:, 0])
        rest.scatter_(1, e[:, None], -float("inf"))
    ti, tw = torch.stack(ti, 1), torch.stack(tw, 1)
    assert torch.equal(ids.cpu().long(), ti)
    assert torch.allclose(w.cpu(), tw * 2.5, rtol=1e-5, atol=1e-7)
    _, wn = ops.moe_route(logits.to(DEV), k, scale=1.0, renormalize=True)
    assert torch.allclose(wn.cpu().sum(-1), torch.ones(T), atol=1e-5)
    row_token, inv_pos, tile_group = ops.moe_align(ids, E)
    row_token, inv_pos, tile_group, idc = row_token.cpu(), inv_pos.cpu(), tile_group.cpu(), ids.cpu()
    assert sorted(inv_pos.tolist()) == sorted(set(inv_pos.tolist()))          # a permutation into the rows
    for i, pos in enumerate(inv_pos.tolist()):
        assert row_token[pos] == i // k and tile_group[pos // 128] == idc.view(-1)[i]
    assert int((row_token >= 0).sum()) == T * k
    used = tile_group[tile_group >= 0]
    assert torch.equal(used, used.sort().values)                             # experts in ascending tile order


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("T", [3, 64, 500])
def test_moe_grouped_and_dense_paths_agree_with_reference(ops, dtype, T):
    """Routed experts through the grouped-GEMM path and through the dense all-experts path against
    the per-token loop of DeepseekV2MoE (deepseek_v2.py:194-218)."""
    torch.manual_seed(T)
    E, k, H, I = 8, 3, 512, 256
    x = (torch.randn(T, H) * 0.5).to(dtype)
    wg = (torch.randn(E, H) / math.sqrt(H)).to(dtype)
    gu = (torch.randn(E, 2 * I, H) / math.sqrt(H)).to(dtype)
    dn = (torch.randn(E, H, I) / math.sqrt(I)).to(dtype)
    logits = ops.linear(x.to(DEV), wg.to(DEV))
    ids, w = ops.moe_route(logits, k, scale=1.0)
    idc, wc = ids.cpu().long(), w.cpu()
    ref = torch.zeros(T, H)
    for t in range(T):
        for j in range(k):
            e = int(idc[t, j])
            g_u = gu[e].float() @ x[t].float()
            a = (torch.nn.functional.silu(g_u[:I].to(dtype).float()) * g_u[I:].to(dtype).float()).to(dtype).float()
            ref[t] += wc[t, j] * (dn[e].float() @ a)
    w_gate_up = gu.reshape(E * 2 * I, H).to(DEV)
    w_down_cat = dn.permute(1, 0, 2).reshape(H, E * I).contiguous().to(DEV)
    xd = x.to(DEV)
    row_token, inv_pos, tile_group = ops.moe_align(ids, E)
    xs = ops.moe_gather(xd, row_token)
    act = ops.linear_grouped_rows(xs, w_gate_up, tile_group, E, I, epilogue=ops.SPD_EPI_SILU_MUL)
    y = ops.linear_grouped_rows(act, w_down_cat, tile_group, E, H, k_concat=True)
    grouped = ops.moe_combine(y, inv_pos, w).cpu().float()
    gu_all = ops.linear(xd, w_gate_up)
    dense = ops.linear(ops.moe_dense_act(gu_all, ids, w, E, I), w_down_cat).cpu().float()
    tol = _tol(dtype, 6e-3, 4e-2)
    assert (grouped - ref).abs().max().item() < tol
    assert (dense - ref).abs().max().item() < tol


# ------------------------------------------------------------------------------------------------
# whole model
# ------------------------------------------------------------------------------------------------
MINI = dict(vocab_size=2048, hidden_size=512, intermediate_size=1024, moe_intermediate_size=256,
            num_hidden_layers=3, num_attention_heads=16, kv_lora_rank=512, qk_nope_head_dim=128,
            qk_rope_head_dim=64, v_head_dim=128, n_routed_experts=8, num_experts_per_tok=3, n_shared_experts=2,
            first_k_dense_replace=1, routed_scaling_factor=1.0, max_position_embeddings=2048)
YARN = {"type": "yarn", "factor": 8.0, "beta_fast": 32, "beta_slow": 1, "mscale": 0.707, "mscale_all_dim": 0.707,
        "original_max_position_embeddings": 256}


def _weights(cfg: D.DeepseekOracleConfig, dtype, seed=3):
    g = torch.Generator().manual_seed(seed)
    H, qk = cfg.hidden_size, cfg.qk_nope_head_dim + cfg.qk_rope_head_dim
    w = {}

    def mat(*shape, scale=None):
        return (torch.randn(*shape, generator=g) * (scale if scale is not None else shape[-1] ** -0.5)).to(dtype)

    def vec(n):
        return (1.0 + 0.1 * torch.randn(n, generator=g)).to(dtype)

    w["model.embed_tokens.weight"] = mat(cfg.vocab_size, H, scale=1.0)
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        w[p + "input_layernorm.weight"] = vec(H)
        w[p + "post_attention_layernorm.weight"] = vec(H)
        w[p + "self_attn.q_proj.weight"] = mat(cfg.num_attention_heads * qk, H)
        w[p + "self_attn.kv_a_proj_with_mqa.weight"] = mat(cfg.kv_lora_rank + cfg.qk_rope_head_dim, H)
        w[p + "self_attn.kv_a_layernorm.weight"] = vec(cfg.kv_lora_rank)
        w[p + "self_attn.kv_b_proj.weight"] = mat(cfg.num_attention_heads * (cfg.qk_nope_head_dim + cfg.v_head_dim),
                                                  cfg.kv_lora_rank)
        w[p + "self_attn.o_proj.weight"] = mat(H, cfg.num_attention_heads * cfg.v_head_dim)
        if i >= cfg.first_k_dense_replace:
            I = cfg.moe_intermediate_size
            w[p + "mlp.gate.weight"] = mat(cfg.n_routed_experts, H, scale=2.0 * H ** -0.5)
            w[p + "mlp.experts.gate_up_proj"] = mat(cfg.n_routed_experts, 2 * I, H)
            w[p + "mlp.experts.down_proj"] = mat(cfg.n_routed_experts, H, I)
            w[p + "mlp.shared_experts.gate_proj.weight"] = mat(I * cfg.n_shared_experts, H)
            w[p + "mlp.shared_experts.up_proj.weight"] = mat(I * cfg.n_shared_experts, H)
            w[p + "mlp.shared_experts.down_proj.weight"] = mat(H, I * cfg.n_shared_experts)
        else:
            w[p + "mlp.gate_proj.weight"] = mat(cfg.intermediate_size, H)
            w[p + "mlp.up_proj.weight"] = mat(cfg.intermediate_size, H)
            w[p + "mlp.down_proj.weight"] = mat(H, cfg.intermediate_size)
    w["model.norm.weight"] = vec(H)
    w["lm_head.weight"] = mat(cfg.vocab_size, H, scale=1.5 * H ** -0.5)
    return w


def _runner(rope_scaling, dtype, cuda_graph=False):
    from semi_pd_b200.configs import ModelConfig, ServerArgs
    from semi_pd_b200.model_runner import ModelRunner
    ocfg = D.DeepseekOracleConfig(**MINI, rope_scaling=rope_scaling)
    weights = _weights(ocfg, dtype)
    mcfg = ModelConfig(name="mini-deepseek", model_type="deepseek_v2", num_key_value_heads=MINI["num_attention_heads"],
                       head_dim=192, rms_norm_eps=ocfg.rms_norm_eps, rope_theta=ocfg.rope_theta,
                       rope_scaling=rope_scaling, **MINI)
    sa = ServerArgs(model_path="mini-deepseek", dtype="float16" if dtype == torch.float16 else "bfloat16",
                    load_format="external", max_total_tokens=8192, max_running_requests=32, context_length=2000,
                    disable_cuda_graph=not cuda_graph, cuda_graph_bs=[1, 2, 4, 8] if cuda_graph else None,
                    enable_semi_pd=True)
    mr = ModelRunner(mcfg, 0.5, 0, 0, 1, 0, sa)
    mr.model.load_weights(weights.items())
    mr.init_attention_backend()
    mr.init_cuda_graphs()
    oracle = D.DeepseekOracle(ocfg, weights, kv_slots=8192, max_reqs=32, max_ctx=2048)
    return mr, oracle, ocfg


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-2), (torch.bfloat16, 1e-1)], ids=["fp16", "bf16"])
@pytest.mark.parametrize("rope_scaling", [None, YARN], ids=["rope", "yarn"])
def test_deepseek_mini_model_matches_oracle(rope_scaling, dtype, tol):
    from tests.helpers import GpuSession
    mr, oracle, ocfg = _runner(rope_scaling, dtype)
    gs, os_ = GpuSession(mr), D.DeepseekSession(oracle, 8192, seed=1)
    V = ocfg.vocab_size
    # ragged prefill: 5 prompts = 457 tokens (grouped-GEMM MoE path), absorbed MLA without a prefix
    lens = [5, 1, 130, 64, 257]
    prompts = [[(7 + 13 * i + 31 * j) % V for i in range(n)] for j, n in enumerate(lens)]
    reqs, lg, ids = gs.extend(prompts)
    rids, lc = os_.extend(prompts)
    # A request whose tokens include a router near-tie (relative margin between the k-th and the
    # (k+1)-th expert below the rounding noise of the 16-bit router logits: the gate is a model-
    # dtype linear in the reference too, deepseek_v2.py:128-139, so bf16 logits of magnitude ~2
    # resolve 0.016 = 1.6 % of a probability) may legitimately
    # route that token to the other expert: such rows get the loose bound, all others the tight
    # one, and the loose ones must stay a minority.
    noise = 4e-3 if dtype == torch.float16 else 4e-2
    tight = torch.ones(len(lens), dtype=torch.bool)
    if oracle.router_margins:
        m = torch.stack(oracle.router_margins).min(0).values
        ends = torch.tensor(lens).cumsum(0).tolist()
        tight = torch.tensor([bool(m[e - n:e].min() > noise) for e, n in zip(ends, lens)])

    def check(lg, lc):
        err = (lg - lc).abs().max(-1).values
        assert err[tight].max().item() < tol if tight.any() else True, err
        assert err.max().item() < 6 * tol, err

    check(lg, lc)
    assert tight.sum() >= 1, tight
    tok = lc.argmax(-1).tolist()
    for _ in range(3):  # decode, bs = 5 (grouped path with few experts)
        lg, ids = gs.decode(reqs, tok)
        lc = os_.decode(rids, tok)
        if oracle.router_margins:
            tight &= torch.stack(oracle.router_margins).min(0).values > noise
        check(lg, lc)
        tok = lc.argmax(-1).tolist()
    # chunked prefill: the second chunk attends to the cached latent prefix
    long = [(5 + 3 * i) % V for i in range(300)]
    r2, _, _ = gs.extend([long[:170]])
    o2, _ = os_.extend([long[:170]])
    _, lg, _ = gs.extend([long[170:]], r2)
    _, lc = os_.extend([long[170:]], req_ids=o2)
    assert (lg - lc).abs().max().item() < tol


def test_deepseek_decode_dense_moe_path_and_cuda_graph():
    """20 running requests: the decode batch takes the dense all-experts path (16 <= bs <= 256);
    a second runner replays its CUDA graph for bs <= 8 — both against the oracle."""
    from tests.helpers import GpuSession
    dtype, tol = torch.float16, 2e-2
    mr, oracle, ocfg = _runner(None, dtype)
    gs, os_ = GpuSession(mr), D.DeepseekSession(oracle, 8192, seed=2)
    V = ocfg.vocab_size
    prompts = [[(3 + 5 * i + 17 * j) % V for i in range(9 + 3 * j)] for j in range(20)]
    reqs, lg, _ = gs.extend(prompts)
    rids, lc = os_.extend(prompts)
    assert (lg - lc).abs().max().item() < tol
    tok = lc.argmax(-1).tolist()
    for _ in range(2):
        lg, _ = gs.decode(reqs, tok)
        lc = os_.decode(rids, tok)
        assert (lg - lc).abs().max().item() < tol
        tok = lc.argmax(-1).tolist()
    del mr, gs
    mr, oracle, ocfg = _runner(None, dtype, cuda_graph=True)
    assert mr.cuda_graph_runner is not None
    gs, os_ = GpuSession(mr), D.DeepseekSession(oracle, 8192, seed=3)
    prompts = prompts[:3]
    reqs, lg, _ = gs.extend(prompts)
    rids, lc = os_.extend(prompts)
    tok = lc.argmax(-1).tolist()
    for _ in range(3):
        lg, _ = gs.decode(reqs, tok)     # bs 3 -> the bs 4 graph, padded row on the dummy slot
        assert mr.cuda_graph_runner.bs == 4
        lc = os_.decode(rids, tok)
        assert (lg - lc).abs().max().item() < tol
        tok = lc.argmax(-1).tolist()


def test_deepseek_semi_pd_engine_matches_sequential_oracle():
    """BASELINE configs[3] end to end: the semi-PD engine (prefill and decode workers on disjoint SM
    partitions, chunked prefill, decode CUDA graphs over the latent pool) serving the mini
    DeepSeek-V2 model produces the sequential oracle's greedy tokens.  A step is compared only
    where the oracle's top-1 / top-2 margin exceeds 2e-2 and no router decision of that forward
    is a near-tie (see test_deepseek_mini_model_matches_oracle)."""
    import semi_pd_b200.model_runner as MR
    from semi_pd_b200.configs import ModelConfig, ServerArgs
    from semi_pd_b200.semi_pd_scheduler import SemiPDEngine

    dtype = torch.float16
    ocfg = D.DeepseekOracleConfig(**MINI, rope_scaling=None)
    weights = _weights(ocfg, dtype)
    mcfg = ModelConfig(name="mini-deepseek", model_type="deepseek_v2", num_key_value_heads=MINI["num_attention_heads"],
                       head_dim=192, rms_norm_eps=ocfg.rms_norm_eps, rope_theta=ocfg.rope_theta, rope_scaling=None,
                       **MINI)
    sa = ServerArgs(model_path="mini-deepseek", dtype="float16", load_format="external", max_total_tokens=8192,
                    max_running_requests=16, context_length=1024, enable_semi_pd=True, cuda_graph_bs=[1, 2, 4, 8, 16],
                    chunked_prefill_size=96, max_prefill_tokens=4096)
    orig = MR.ModelRunner.load_model

    def load_with_weights(self):
        orig(self)
        self.model.load_weights(weights.items())

    MR.ModelRunner.load_model = load_with_weights
    try:
        eng = SemiPDEngine(sa, gpu_id=0, model_config=mcfg, prefill_sm_percent=50, decode_sm_percent=50)
    finally:
        MR.ModelRunner.load_model = orig
    lens, n_new = [5, 130, 17, 64, 257, 33], 8
    prompts = [[(11 + 7 * i + 29 * j) % MINI["vocab_size"] for i in range(n)] for j, n in enumerate(lens)]
    try:
        assert eng.d_worker.model_runner.cuda_graph_runner is not None
        rids = [eng.submit(p, n_new) for p in prompts]
        done = {r.rid: r for r in eng.wait(len(prompts), timeout=120)}
        got = [done[r].output_ids for r in rids]
    finally:
        eng.shutdown()
    checked = bad = 0
    for p, g_ in zip(prompts, got):
        assert len(g_) == n_new
        oracle = D.DeepseekOracle(ocfg, weights, kv_slots=2048, max_reqs=4, max_ctx=1100)
        s = D.DeepseekSession(oracle, 2048, seed=5)
        r_o, logits = s.extend([p])
        for k in range(n_new):
            top2 = logits.topk(2, -1).values[0]
            near_tie = oracle.router_margins and float(torch.stack(oracle.router_margins).min()) < 4e-3
            if float(top2[0] - top2[1]) <= 2e-2 or near_tie:
                break  # from here on the two sequences may legitimately differ
            t = int(logits.argmax(-1)[0])
            bad += int(g_[k] != t)
            checked += 1
            if k + 1 < n_new:
                logits = s.decode(r_o, [t])
    assert bad == 0 and checked >= 12, (checked, bad)
