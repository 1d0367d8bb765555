"""Engine-level parity: requests pushed through the in-process semi-PD engine (prefill worker and
decode worker running CONCURRENTLY on disjoint green-context SM partitions, CUDA-graph decode,
chunked prefill, overlap scheduling) must produce exactly the greedy token ids the CPU oracle
produces one request at a time — wherever the oracle's top-1/top-2 margin exceeds 2e-2.
This is the property the reference never tests ("no Semi-PD tests"): P‖D output ==
sequential output."""
import pytest
import torch

from oracle import llama_oracle as O

pytestmark = pytest.mark.gpu

CFG = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
           num_key_value_heads=2, head_dim=128, vocab_size=1024, max_position_embeddings=2048)


def _oracle_generate(oracle_factory, prompt, n_new):
    """Greedy generation of ONE request; returns (tokens, margins)."""
    m = oracle_factory()
    s = O.OracleSession(m, 2048, 4)
    rids, logits, _ = s.extend([prompt])
    toks, margins = [], []
    for _ in range(n_new):
        top2 = logits.topk(2, -1).values[0]
        margins.append(float(top2[0] - top2[1]))
        t = int(logits.argmax(-1)[0])
        toks.append(t)
        if len(toks) == n_new:
            break
        logits, _ = s.decode(rids, [t])
    return toks, margins


def _build_engine(dtype, weights, chunked_prefill_size, partition="green", max_total_tokens=8192,
                  conservativeness=1.0, cuda_graph_bs=(1, 2, 4, 8, 16), max_running_requests=16):
    from semi_pd_b200.configs import ModelConfig, ServerArgs
    from semi_pd_b200.semi_pd_scheduler import SemiPDEngine
    sa = ServerArgs(model_path="tiny", dtype="float16" if dtype == torch.float16 else "bfloat16",
                    load_format="external", max_total_tokens=max_total_tokens,
                    max_running_requests=max_running_requests,
                    context_length=1024, enable_semi_pd=True, cuda_graph_bs=list(cuda_graph_bs),
                    chunked_prefill_size=chunked_prefill_size, max_prefill_tokens=4096,
                    schedule_conservativeness=conservativeness)
    # external weights must be in place before the CUDA graphs are captured
    import semi_pd_b200.model_runner as MR
    orig = MR.ModelRunner.load_model

    def load_with_weights(self):
        orig(self)
        self.model.load_weights(weights.items())

    MR.ModelRunner.load_model = load_with_weights
    try:
        eng = SemiPDEngine(sa, gpu_id=0, model_config=ModelConfig(name="tiny", **CFG), prefill_sm_percent=50,
                           decode_sm_percent=50, partition=partition)
    finally:
        MR.ModelRunner.load_model = orig
    return eng


@pytest.mark.parametrize("chunk", [4096, 96], ids=["whole-prompts", "chunked-prefill"])
def test_engine_matches_sequential_oracle(chunk):
    dtype = torch.float16
    ocfg = O.OracleLlamaConfig(**CFG)
    weights = O.make_weights(ocfg, dtype, seed=77)
    lens = [5, 130, 17, 64, 257, 1, 33, 200, 9, 128, 75, 301]
    n_new = 12
    prompts = [[(11 + 7 * i + 29 * j) % CFG["vocab_size"] for i in range(n)] for j, n in enumerate(lens)]

    def factory():
        return O.LlamaOracle(ocfg, weights, kv_slots=2048, max_reqs=4, max_ctx=1100, compute="fp32")

    expected = [_oracle_generate(factory, p, n_new) for p in prompts]
    eng = _build_engine(dtype, weights, chunk)
    try:
        assert eng.p_sms + eng.d_sms <= 148 and eng.p_sms >= 8 and eng.d_sms >= 8
        eng.warmup(prompt_len=8, n=2, max_new_tokens=2)
        rids = [eng.submit(p, n_new) for p in prompts]
        done = {r.rid: r for r in eng.wait(len(prompts), timeout=120)}
        pools = (eng.d_worker.model_runner.token_to_kv_pool_allocator, eng.d_worker.model_runner.req_to_token_pool)
    finally:
        eng.shutdown()
    checked = 0
    for rid, (toks, margins) in zip(rids, expected):
        got = done[rid].output_ids
        assert len(got) == n_new
        for k in range(n_new):
            if margins[k] <= 2e-2:
                break  # beyond a near-tie the two greedy paths may legitimately diverge
            assert got[k] == toks[k], (rid, k, got, toks)
            checked += 1
    assert checked > 0.7 * len(prompts) * n_new
    # bookkeeping: every KV run and request slot came back
    assert pools[0].available_size() == pools[0].num_runs * pools[0].run_len
    assert pools[1].available_size() == pools[1].size


def test_engine_timestamps_and_device_resident_arm():
    dtype = torch.float16
    ocfg = O.OracleLlamaConfig(**CFG)
    weights = O.make_weights(ocfg, dtype, seed=5)
    eng = _build_engine(dtype, weights, 4096)
    try:
        prompts = [[(3 + i + j) % 1024 for i in range(40)] for j in range(6)]
        rids = [eng.submit(p, 8) for p in prompts]
        reqs = eng.wait(6, timeout=60)
        for r in reqs:
            assert r.first_token_time >= r.arrival_time and len(r.token_times) == 8
            assert all(b >= a for a, b in zip(r.token_times, r.token_times[1:]))
        host_ids = {r.rid: r.output_ids for r in reqs}
        # device-resident arm: no token ids come back, completion is by launch count
        eng.set_device_resident(True)
        dev = [torch.tensor(p, dtype=torch.int64, device="cuda:0") for p in prompts]
        for p, d in zip(prompts, dev):
            eng.submit(p, 8, input_ids_device=d)
        reqs2 = eng.wait(6, timeout=60)
        assert all(len(r.output_ids) == 8 and r.output_ids[0] == -1 for r in reqs2)
        assert len(host_ids) == 6 and rids
    finally:
        eng.shutdown()


def _check_against_oracle(rids, expected, done, n_new):
    checked = 0
    for rid, (toks, margins) in zip(rids, expected):
        got = done[rid].output_ids
        assert len(got) == n_new, (rid, len(got))
        for k in range(n_new):
            if margins[k] <= 2e-2:
                break
            assert got[k] == toks[k], (rid, k, got, toks)
            checked += 1
    return checked


def test_decode_oom_retraction_reprefill_same_tokens():
    """Forces the path of semi_pd_decode_scheduler.py:102-139: a KV pool of 8 runs and an
    optimistic admission (schedule_conservativeness 0.05) over-admit, the decode batch runs out of
    runs when the requests cross their first 128 tokens, the scheduler retracts requests, frees
    their KV and re-submits prompt + generated tokens to the prefill worker.  Every request must
    still deliver exactly the sequential oracle's tokens, as ONE request with its full history."""
    dtype = torch.float16
    ocfg = O.OracleLlamaConfig(**CFG)
    weights = O.make_weights(ocfg, dtype, seed=31)
    n_new = 60
    lens = [100, 90, 110, 95, 105, 99]
    prompts = [[(17 + 5 * i + 41 * j) % CFG["vocab_size"] for i in range(n)] for j, n in enumerate(lens)]

    def factory():
        return O.LlamaOracle(ocfg, weights, kv_slots=2048, max_reqs=4, max_ctx=1100, compute="fp32")

    expected = [_oracle_generate(factory, p, n_new) for p in prompts]
    eng = _build_engine(dtype, weights, 4096, max_total_tokens=8 * 128, conservativeness=0.05)
    try:
        eng.warmup(prompt_len=8, n=2, max_new_tokens=2)
        t_submit = __import__("time").perf_counter()
        rids = [eng.submit(p, n_new) for p in prompts]
        done = {r.rid: r for r in eng.wait(len(prompts), timeout=180)}
        retracted = eng.decode.num_retracted
        alloc = eng.d_worker.model_runner.token_to_kv_pool_allocator
        free_after = alloc.available_size()
        total = alloc.num_runs * alloc.run_len
    finally:
        eng.shutdown()
    assert retracted > 0, "the test did not exercise retraction"
    assert _check_against_oracle(rids, expected, done, n_new) > 0.3 * len(prompts) * n_new
    for rid, p in zip(rids, prompts):
        r = done[rid]
        assert r.origin_input_ids == p                      # the client sees its own prompt back
        assert len(r.token_times) == n_new and r.first_token_time >= t_submit
        assert r.first_token_time == r.token_times[0]       # TTFT survives the retraction
        assert all(b >= a for a, b in zip(r.token_times, r.token_times[1:]))
    assert free_after == total


def test_eager_steps_above_graph_bs_do_not_corrupt_small_graph_replays():
    """the split-KV workspace captured by the small-bs decode graphs must
    survive eager decode steps with bs > cuda_graph_max_bs.  10 requests run with graphs captured
    for bs <= 4 only: the batch decodes eagerly while it is large, and through the bs 4/2/1 graphs
    (which split the KV across CTAs and use the workspace) once the short requests have finished."""
    dtype = torch.float16
    ocfg = O.OracleLlamaConfig(**CFG)
    weights = O.make_weights(ocfg, dtype, seed=13)
    lens = [40, 33, 57, 21, 64, 9, 130, 77, 5, 48]
    news = [4, 4, 4, 4, 4, 4, 24, 24, 16, 16]
    prompts = [[(23 + 3 * i + 19 * j) % CFG["vocab_size"] for i in range(n)] for j, n in enumerate(lens)]

    def factory():
        return O.LlamaOracle(ocfg, weights, kv_slots=2048, max_reqs=4, max_ctx=1100, compute="fp32")

    expected = [_oracle_generate(factory, p, n) for p, n in zip(prompts, news)]
    eng = _build_engine(dtype, weights, 4096, cuda_graph_bs=(1, 2, 4))
    try:
        ws_ptr = eng.d_worker.model_runner.attn_backend.decode_workspace.data_ptr()
        eng.warmup(prompt_len=8, n=2, max_new_tokens=2)
        rids = [eng.submit(p, n) for p, n in zip(prompts, news)]
        done = {r.rid: r for r in eng.wait(len(prompts), timeout=120)}
        assert eng.d_worker.model_runner.attn_backend.decode_workspace.data_ptr() == ws_ptr
    finally:
        eng.shutdown()
    checked = 0
    for rid, (toks, margins), n in zip(rids, expected, news):
        got = done[rid].output_ids
        assert len(got) == n
        for k in range(n):
            if margins[k] <= 2e-2:
                break
            assert got[k] == toks[k], (rid, k, got, toks)
            checked += 1
    assert checked > 0.3 * sum(news)


def test_request_validation_and_two_engines_in_one_process():
    """over-long inputs are rejected at the engine front, max_new_tokens is
    clipped to the context, and the queue keeps moving.  Also: shutdown() releases the engine's
    device memory and a second engine built afterwards reuses the pooled SM partition."""
    dtype = torch.float16
    ocfg = O.OracleLlamaConfig(**CFG)
    weights = O.make_weights(ocfg, dtype, seed=5)
    for round_ in range(2):
        eng = _build_engine(dtype, weights, 4096)
        try:
            limit = eng.d_worker.max_req_input_len
            with pytest.raises(ValueError):
                eng.submit(list(range(limit + 3)), 4)
            with pytest.raises(ValueError):
                eng.submit([], 4)
            # max_new_tokens far beyond the context is clipped, not an out-of-bounds write
            rid_long = eng.submit([(i * 7) % 1024 for i in range(900)], 100000)
            rid_ok = eng.submit([1, 2, 3, 4], 5)
            done = {r.rid: r for r in eng.wait(2, timeout=120)}
            assert len(done[rid_ok].output_ids) == 5
            assert 900 + len(done[rid_long].output_ids) <= eng.d_worker.max_req_len
            assert len(done[rid_long].output_ids) == eng.d_worker.max_req_len - 900 - 1
            part = eng.partition
        finally:
            eng.shutdown()
        if round_ == 0:
            first_part, mem_after_first = part, torch.cuda.memory_allocated()
        else:
            assert part is first_part                       # green contexts are pooled, not leaked
            assert torch.cuda.memory_allocated() <= mem_after_first + (8 << 20)


def test_http_generate_endpoint_in_process():
    """`python -m sglang.launch_server --enable-semi-pd` serves /generate in token-id space: the
    same FastAPI app driven in-process (starlette TestClient), concurrent requests from a thread
    pool, ids equal to the engine-free sequential oracle; an over-long input is a 400."""
    from concurrent.futures import ThreadPoolExecutor

    from starlette.testclient import TestClient

    import semi_pd_b200.model_runner as MR
    from semi_pd_b200.configs import ModelConfig
    from semi_pd_b200.launch_server import build_app, prepare_server_args

    dtype = torch.float16
    ocfg = O.OracleLlamaConfig(**CFG)
    weights = O.make_weights(ocfg, dtype, seed=21)
    sa = prepare_server_args(["--model-path", "tiny", "--enable-semi-pd", "--dtype", "float16",
                              "--load-format", "external", "--max-total-tokens", "8192",
                              "--max-running-requests", "16", "--context-length", "512"])
    orig = MR.ModelRunner.load_model

    def load_with_weights(self):
        orig(self)
        self.model.load_weights(weights.items())

    MR.ModelRunner.load_model = load_with_weights
    try:
        app, engine = build_app(sa, model_config=ModelConfig(name="tiny", **CFG))
    finally:
        MR.ModelRunner.load_model = orig
    try:
        client = TestClient(app)
        assert client.get("/health").json() == {"status": "ok"}
        info = client.get("/get_server_info").json()
        assert info["prefill_sms"] + info["decode_sms"] <= 148 and info["max_total_num_tokens"] >= 8192
        lens, n_new = [7, 130, 33, 64], 6
        prompts = [[(5 + 3 * i + 11 * j) % CFG["vocab_size"] for i in range(n)] for j, n in enumerate(lens)]

        def post(p):
            return client.post("/generate", json={"input_ids": p, "sampling_params": {"max_new_tokens": n_new}})

        with ThreadPoolExecutor(4) as ex:
            outs = list(ex.map(post, prompts))
        for p, r in zip(prompts, outs):
            assert r.status_code == 200
            body = r.json()
            assert body["meta_info"]["prompt_tokens"] == len(p) and body["meta_info"]["completion_tokens"] == n_new
            m = O.LlamaOracle(ocfg, weights, kv_slots=1024, max_reqs=2, max_ctx=520, compute="fp32")
            s = O.OracleSession(m, 1024, 2)
            rids, logits, _ = s.extend([p])
            for k in range(n_new):
                top2 = logits.topk(2, -1).values[0]
                if float(top2[0] - top2[1]) <= 2e-2:
                    break
                t = int(logits.argmax(-1)[0])
                assert body["output_ids"][k] == t
                if k + 1 < n_new:
                    logits, _ = s.decode(rids, [t])
        bad = client.post("/generate", json={"input_ids": list(range(600)), "sampling_params": {"max_new_tokens": 4}})
        assert bad.status_code == 400 and "exceeds" in bad.json()["detail"]
    finally:
        engine.shutdown()
