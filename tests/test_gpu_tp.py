"""Tensor-parallel forward path on real GPUs (needs >= 2; skipped otherwise): N ranks, sharded
heads / FFN columns, SUM all-reduce after o_proj and down_proj only — libspd's peer-memory
`spd_allreduce` (one-shot / two-shot over IPC-mapped buffers) for decode-sized messages, NCCL above
8 MiB — and a vocab-sharded lm_head whose greedy ids come from an all-gather of (max, argmax) rows
(SURVEY §8e).  Every rank must reproduce the unsharded CPU oracle's logits; decode runs both
eagerly and through the captured CUDA graph (the collective is inside the graph); the semi-PD
engine runs with tp_size = 2 (P and D on separate communicators, rank 0 broadcasting the decode
scheduler's messages) and must produce the sequential oracle's tokens.
The driver's GPUTEST box has one GPU: the pass is recorded by `gpurun --gpus 2` logs under
profiles/ and by `bench.py --gpus N`'s `tp_parity` block in every SCALE run."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CFG = dict(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=8,
           num_key_value_heads=4, head_dim=128, vocab_size=2048, max_position_embeddings=2048)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, use_graph, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import llama_oracle as O
    from semi_pd_b200.configs import ModelConfig, ServerArgs
    from semi_pd_b200.model_runner import ModelRunner
    from tests.helpers import GpuSession

    dtype = torch.float16
    ocfg = O.OracleLlamaConfig(**CFG)
    weights = O.make_weights(ocfg, dtype, seed=99)  # identical full checkpoint on every rank
    sa = ServerArgs(model_path="tp-test", dtype="float16", load_format="external", tp_size=world,
                    max_total_tokens=2048, max_running_requests=8, context_length=512,
                    disable_cuda_graph=not use_graph, cuda_graph_bs=[1, 2, 4], enable_semi_pd=True)
    mr = ModelRunner(ModelConfig(name="tp-test", **CFG), 0.5, rank, rank, world, port, sa)
    mr.model.load_weights(weights.items())  # each rank keeps its shard
    mr.init_attention_backend()
    mr.init_cuda_graphs()
    assert mr.token_to_kv_pool.head_num == CFG["num_key_value_heads"] // world
    oracle = O.LlamaOracle(ocfg, weights, kv_slots=2048, max_reqs=9, max_ctx=516, compute="fp32")
    gs, os_ = GpuSession(mr), O.OracleSession(oracle, 2048, 9)
    prompts = [[(5 + 3 * i + 17 * j) % CFG["vocab_size"] for i in range(n)] for j, n in enumerate([40, 7, 129])]
    reqs, lg, ids = gs.extend(prompts)
    rids, lc, _ = os_.extend(prompts)
    err = (lg - lc).abs().max().item()
    tok = lc.argmax(-1).tolist()
    for _ in range(3):
        lg, ids = gs.decode(reqs, tok)
        lc, _ = os_.decode(rids, tok)
        if lg.shape[1] != lc.shape[1]:
            # a captured step returns this rank's vocab shard only (the ids come from the
            # (max, argmax) all-gather inside the graph)
            n = lg.shape[1]
            assert use_graph and n == CFG["vocab_size"] // world
            top2 = lc.topk(2, -1).values
            clear = (top2[:, 0] - top2[:, 1]) > 2e-2
            assert torch.equal(ids.long()[clear], lc.argmax(-1)[clear])
            lc_cmp = lc[:, rank * n:(rank + 1) * n]
        else:
            lc_cmp = lc
        err = max(err, (lg - lc_cmp).abs().max().item())
        tok = lc.argmax(-1).tolist()
    # the collectives themselves, against NCCL, eager and replayed from a graph
    comm = mr.tp_comm
    assert comm is not None and comm.max_bytes > 0, "libspd communicator missing"
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    for n, algo in [(8, 1), (4096 * 3, 1), (4096 * 64, 1), (4096 * 64, 2), (4096 * 160 + 8, 2), (4096 * 512, 0)]:
        x = torch.randn(n, generator=g, device="cuda", dtype=dtype)
        ref = x.float().clone()
        torch.distributed.all_reduce(ref)
        y = comm.all_reduce(x.clone(), algo=algo)
        assert (y.float() - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1.0), (n, algo)
        # bit-identical on all ranks
        z = y.clone()
        torch.distributed.broadcast(z, src=0)
        assert torch.equal(z, y), (n, algo)
    # the all-reduce fused with the residual-add RMSNorm that consumes it: bit-identical to the two
    # separate kernels, eager and replayed from a graph (odd row counts, 70B-wide rows)
    from semi_pd_b200 import ops
    for rows, dim in [(1, 512), (3, 4096), (64, 4096), (200, 4096), (37, 8192)]:
        xp = torch.randn(rows, dim, generator=g, device="cuda", dtype=dtype)
        res = torch.randn(rows, dim, device="cuda", dtype=dtype)
        torch.distributed.broadcast(res, src=0)      # the residual stream is replicated
        wn = (torch.randn(dim, device="cuda") * 0.1 + 1).to(dtype)
        torch.distributed.broadcast(wn, src=0)
        xa, ra = comm.all_reduce(xp.clone()), res.clone()
        ops.fused_add_rmsnorm(xa, ra, wn, 1e-5)
        xb, rb = xp.clone(), res.clone()
        assert comm.all_reduce_add_rmsnorm(xb, rb, wn, 1e-5, force=True), (rows, dim)
        assert torch.equal(xa, xb) and torch.equal(ra, rb), (rows, dim)
    big = torch.randn(600, 4096, device="cuda", dtype=dtype)
    assert not comm.all_reduce_add_rmsnorm(big, big.clone(), wn[:4096].contiguous(), 1e-5, force=True)  # prefill-sized
    x = torch.randn(4096 * 8, generator=g, device="cuda", dtype=dtype)
    static_in, ref = x.clone(), x.float().clone()
    torch.distributed.all_reduce(ref)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        static_out = comm.all_reduce(static_in * 1, algo=0)
        static_out2 = comm.all_reduce(static_out * 1, algo=1)  # an odd number of calls per replay
        static_out3 = comm.all_reduce(static_in * 1, algo=2)
    for _ in range(5):
        gr.replay()
    torch.cuda.synchronize()
    assert (static_out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert torch.equal(static_out3, static_out)
    assert comm.status() == 0
    ret[rank] = err
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "cuda-graph"])
def test_tp2_matches_unsharded_oracle(use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, use_graph, ret), nprocs=world, join=True)
        errs = dict(ret)
    assert set(errs) == {0, 1}
    assert max(errs.values()) < 1e-2, errs


ECFG = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
            num_key_value_heads=2, head_dim=128, vocab_size=1024, max_position_embeddings=2048)


def _engine_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import llama_oracle as O
    import semi_pd_b200.model_runner as MR
    from semi_pd_b200.configs import ModelConfig, ServerArgs
    from semi_pd_b200.semi_pd_scheduler import SemiPDEngine

    dtype = torch.float16
    ocfg = O.OracleLlamaConfig(**ECFG)
    weights = O.make_weights(ocfg, dtype, seed=77)
    sa = ServerArgs(model_path="tiny", dtype="float16", load_format="external", tp_size=world,
                    max_total_tokens=8192, max_running_requests=16, context_length=1024, enable_semi_pd=True,
                    cuda_graph_bs=[1, 2, 4, 8, 16], chunked_prefill_size=96, max_prefill_tokens=4096)
    orig = MR.ModelRunner.load_model

    def load_with_weights(self):
        orig(self)
        self.model.load_weights(weights.items())

    MR.ModelRunner.load_model = load_with_weights
    eng = SemiPDEngine(sa, gpu_id=rank, tp_rank=rank, model_config=ModelConfig(name="tiny", **ECFG),
                       prefill_sm_percent=50, decode_sm_percent=50)
    MR.ModelRunner.load_model = orig
    lens = [5, 130, 17, 64, 257, 1, 33, 200]
    n_new = 10
    prompts = [[(11 + 7 * i + 29 * j) % ECFG["vocab_size"] for i in range(n)] for j, n in enumerate(lens)]
    if rank == 0:
        try:
            assert eng.d_worker.model_runner.cuda_graph_runner is not None, "decode graphs must be on under TP"
            rids = [eng.submit(p, n_new) for p in prompts]
            done = {r.rid: r for r in eng.wait(len(prompts), timeout=180)}
            got = [done[r].output_ids for r in rids]
        finally:
            eng.shutdown()
        bad = 0
        checked = 0
        for p, g_ in zip(prompts, got):
            m = O.LlamaOracle(ocfg, weights, kv_slots=2048, max_reqs=4, max_ctx=1100, compute="fp32")
            s = O.OracleSession(m, 2048, 4)
            rids_o, logits, _ = s.extend([p])
            for k in range(n_new):
                top2 = logits.topk(2, -1).values[0]
                if float(top2[0] - top2[1]) <= 2e-2:
                    break
                t = int(logits.argmax(-1)[0])
                bad += int(g_[k] != t)
                checked += 1
                if k + 1 < n_new:
                    logits, _ = s.decode(rids_o, [t])
        ret["checked"], ret["bad"] = checked, bad
    else:
        eng.join()
        eng.shutdown()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_tp2_semi_pd_engine_matches_sequential_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_engine_worker, args=(world, port, ret), nprocs=world, join=True)
        out = dict(ret)
    assert out["bad"] == 0 and out["checked"] > 40, out


def _http_tp_main(ret):
    """TP rank 0 of `python -m sglang.launch_server --enable-semi-pd --tp 2` driven in-process: the
    FastAPI app (starlette TestClient) on GPU 0, rank 1 spawned by build_app on GPU 1."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from starlette.testclient import TestClient

    from oracle import llama_oracle as O
    from semi_pd_b200.configs import get_model_config
    from semi_pd_b200.launch_server import build_app, prepare_server_args

    argv = ["--model-path", "tiny-llama-test", "--enable-semi-pd", "--tp", "2", "--dtype", "float16",
            "--max-total-tokens", "8192", "--max-running-requests", "16", "--context-length", "512"]
    sa = prepare_server_args(argv)
    app, engine = build_app(sa, argv=argv)
    try:
        client = TestClient(app)
        assert client.get("/health").json() == {"status": "ok"}
        cfg = get_model_config("tiny-llama-test")
        lens, n_new = [9, 130, 33], 6
        prompts = [[(5 + 3 * i + 11 * j) % cfg.vocab_size for i in range(n)] for j, n in enumerate(lens)]
        outs = [client.post("/generate", json={"input_ids": p, "sampling_params": {"max_new_tokens": n_new}}).json()
                for p in prompts]
        # the dummy model's GLOBAL weights are a function of (shapes, seed): rebuild them for the oracle
        model = engine.d_worker.model_runner.model
        weights = {k: v.cpu() for k, v in model.dummy_full_weights(sa.random_seed)}
        ocfg = O.OracleLlamaConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                                   num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                                   num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim,
                                   vocab_size=cfg.vocab_size, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                                   max_position_embeddings=cfg.max_position_embeddings)
        checked = bad = 0
        for p, body in zip(prompts, outs):
            assert len(body["output_ids"]) == n_new
            m = O.LlamaOracle(ocfg, weights, kv_slots=1024, max_reqs=2, max_ctx=520, compute="fp32")
            s = O.OracleSession(m, 1024, 2)
            rids, logits, _ = s.extend([p])
            for k in range(n_new):
                top2 = logits.topk(2, -1).values[0]
                if float(top2[0] - top2[1]) <= 2e-2:
                    break
                t = int(logits.argmax(-1)[0])
                bad += int(body["output_ids"][k] != t)
                checked += 1
                if k + 1 < n_new:
                    logits, _ = s.decode(rids, [t])
        ret["checked"], ret["bad"] = checked, bad
    finally:
        engine.shutdown()
        for pr in engine.tp_followers:
            pr.join(timeout=60)
        ret["followers_exited"] = all(not pr.is_alive() for pr in engine.tp_followers)


def test_launch_server_tp2_generate_matches_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    with mp.Manager() as m:
        ret = m.dict()
        pr = ctx.Process(target=_http_tp_main, args=(ret,))  # not a daemon: it spawns TP rank 1
        pr.start()
        pr.join(timeout=170)
        alive = pr.is_alive()
        if alive:
            pr.kill()
        out = dict(ret)
    assert not alive and pr.exitcode == 0, (pr.exitcode, out)
    assert out["bad"] == 0 and out["checked"] >= 8 and out["followers_exited"], out
