"""Host-side logic that needs no GPU: allocators, enums, TP sharding rules, admission policy,
workload generator."""
import math

import numpy as np
import pytest
import torch

from oracle import llama_oracle as O
from semi_pd_b200.configs import ServerArgs, default_cuda_graph_bs, get_model_config
from semi_pd_b200.forward_batch_info import ForwardMode
from semi_pd_b200.layers import MergedColumnParallelLinear, ParallelLMHead, QKVParallelLinear, RowParallelLinear
from semi_pd_b200.memory_pool import ReqToTokenPool, TokenToKVPoolAllocator
from semi_pd_b200.schedule_batch import Req, SamplingParams
from semi_pd_b200.schedule_policy import AddReqResult, PrefillAdder


def test_forward_mode_values_and_predicates():
    # integer values line up with forward_batch_info.py:53-70
    assert [m.value for m in (ForwardMode.EXTEND, ForwardMode.DECODE, ForwardMode.MIXED, ForwardMode.IDLE)] == [1, 2, 3, 4]
    assert ForwardMode.EXTEND.is_extend() and ForwardMode.MIXED.is_extend() and not ForwardMode.DECODE.is_extend()
    assert ForwardMode.DECODE.is_cuda_graph() and ForwardMode.IDLE.is_decode_or_idle()


def test_req_to_token_pool():
    pool = ReqToTokenPool(size=5, max_context_len=16, device="cpu")
    assert pool.req_to_token.dtype == torch.int32 and pool.req_to_token.shape == (5, 16)
    a = pool.alloc(3)
    assert a == [0, 1, 2] and pool.available_size() == 2 and pool.alloc(3) is None
    pool.free(a[1])
    pool.free([a[0]])
    assert pool.available_size() == 4
    with pytest.raises(RuntimeError):
        pool.free(a[1])  # double free is an error, not a silently duplicated slot
    pool.clear()
    assert pool.available_size() == 5


def test_run_allocator_contract_and_runs():
    """External contract of memory_pool.py:124-184 (page_size 1, int64 slot ids, slot 0 reserved,
    alloc / free / available_size) over an allocator that hands out runs of consecutive slots."""
    R = 8
    alloc = TokenToKVPoolAllocator(50, torch.bfloat16, "cpu", kvcache=None, run_len=R)
    assert alloc.page_size == 1 and alloc.num_runs == 6 and alloc.available_size() == 48
    # request-agnostic contract: ids are int64, never 0, consecutive inside a run
    s = alloc.alloc(11)
    assert s.dtype == torch.int64 and s.tolist() == list(range(1, 12)) and alloc.available_size() == 32
    assert alloc.alloc(33) is None and alloc.available_size() == 32
    alloc.free_group_begin()
    alloc.free(s[:3])
    alloc.free(s[3:8])
    assert alloc.available_size() == 32     # nothing returns before the group ends
    alloc.free_group_end()
    assert alloc.available_size() == 40     # run 0 is whole again; run 1 still holds slots 9..11
    alloc.free(s[8:])
    assert alloc.available_size() == 48
    with pytest.raises(RuntimeError):
        alloc.free(s[:1])
    alloc.clear()
    # request-aware entry points: positions [k*R, k*R + R) of a request are R consecutive slots
    loc = alloc.alloc_extend([3, 4], [0, 0], [13, 5])
    a, b = loc[:13].tolist(), loc[13:].tolist()
    assert a[:8] == list(range(a[0], a[0] + 8)) and a[8:] == list(range(a[8], a[8] + 5))
    assert (a[0] - 1) % R == 0 and (a[8] - 1) % R == 0 and (b[0] - 1) % R == 0 and 0 not in a + b
    assert alloc.available_size() == 48 - 3 * R
    # chunked continuation fills the request's last run before taking a new one
    loc2 = alloc.alloc_extend([3], [13], [20]).tolist()
    assert loc2[:3] == [a[12] + 1, a[12] + 2, a[12] + 3] and (loc2[3] - 1) % R == 0
    assert alloc.req_slots(3).tolist() == a + loc2
    # decode: one slot per request, continuing its run; a full run takes a new one
    assert alloc.can_alloc_decode([20, 5])
    d = alloc.alloc_decode([3, 4], [20, 5]).tolist()
    assert d == [loc2[-1] + 1, b[-1] + 1]
    for step in range(21, 24):
        alloc.alloc_decode([3], [step])
    assert alloc.available_size() == 48 - 4 * R  # request 3 holds 24 tokens = 3 runs, request 4 one
    nxt = alloc.alloc_decode([3], [24]).tolist()
    assert (nxt[0] - 1) % R == 0 and alloc.available_size() == 48 - 5 * R
    # exhaustion is reported, not half-applied
    assert alloc.alloc_extend([9], [0], [3 * R]) is None and alloc.available_size() == 8
    with pytest.raises(RuntimeError):
        alloc.alloc_extend([3], [7], [9])    # prefix does not match what the request holds
    assert alloc.free_req(3) == 25 and alloc.free_req(4) == 6 and alloc.available_size() == 48
    assert alloc.free_req(4) == 0


def test_cuda_graph_batch_sizes_and_presets():
    assert default_cuda_graph_bs(160) == [1, 2, 4] + [8 * i for i in range(1, 21)]
    c = get_model_config("meta-llama/Llama-3-8B")
    assert (c.hidden_size, c.num_hidden_layers, c.num_key_value_heads, c.vocab_size) == (4096, 32, 8, 128256)
    assert get_model_config("llama-3-70b").get_num_kv_heads(8) == 1
    assert get_model_config("llama-3-70b").get_num_kv_heads(16) == 1  # replicated when tp > n_kv
    o = get_model_config("facebook/opt-125m")  # BASELINE configs[0]
    assert (o.model_type, o.hidden_size, o.num_hidden_layers, o.head_dim, o.vocab_size) == ("opt", 768, 12, 64, 50272)
    with pytest.raises(ValueError):
        get_model_config("gpt2")
    sa = ServerArgs(enable_semi_pd=True)
    assert sa.disable_radix_cache and sa.chunked_prefill_size == 8192


@pytest.mark.parametrize("tp", [1, 2, 4, 8])
def test_tp_sharding_reconstructs_full_product(tp):
    """Column/row sharding rules (linear.py): concatenating column shards / summing row shards of
    the sharded computation equals the unsharded layer (Megatron identity), incl. the fused
    q‖k‖v and gate‖up row orders."""
    torch.manual_seed(tp)
    H, D, Hq, Hk, I = 64, 8, 8, 4 if tp <= 4 else 8, 48
    x = torch.randn(5, H, dtype=torch.float64)
    wqkv = torch.randn((Hq + 2 * Hk) * D, H, dtype=torch.float64)
    wo = torch.randn(H, Hq * D, dtype=torch.float64)
    wgu = torch.randn(2 * I * tp // tp * 1, H, dtype=torch.float64) if False else torch.randn(2 * 16 * tp, H, dtype=torch.float64)
    wd = torch.randn(H, 16 * tp, dtype=torch.float64)
    full_qkv = x @ wqkv.T
    q_full, k_full, v_full = full_qkv.split([Hq * D, Hk * D, Hk * D], -1)
    o_sum, d_sum = 0, 0
    for r in range(tp):
        sh = QKVParallelLinear.shard(wqkv, D, Hq, Hk, r, tp)
        nq, nk = Hq // tp, max(1, Hk // tp)
        q, k, v = (x @ sh.T).split([nq * D, nk * D, nk * D], -1)
        assert torch.equal(q, q_full[:, r * nq * D:(r + 1) * nq * D])
        kv_rank = r // max(1, tp // Hk)
        assert torch.equal(k, k_full[:, kv_rank * nk * D:(kv_rank + 1) * nk * D])
        assert torch.equal(v, v_full[:, kv_rank * nk * D:(kv_rank + 1) * nk * D])
        o_sum = o_sum + q_full[:, r * nq * D:(r + 1) * nq * D] @ RowParallelLinear.shard(wo, r, tp).T
        gu = x @ MergedColumnParallelLinear.shard(wgu, r, tp).T
        act = O.silu_and_mul(gu)
        d_sum = d_sum + act @ RowParallelLinear.shard(wd, r, tp).T
    torch.testing.assert_close(o_sum, q_full @ wo.T)
    torch.testing.assert_close(d_sum, O.silu_and_mul(x @ wgu.T) @ wd.T)
    lm = torch.randn(32, H, dtype=torch.float64)
    parts = [x @ ParallelLMHead.shard(lm, r, tp).T for r in range(tp)]
    torch.testing.assert_close(torch.cat(parts, -1), x @ lm.T)


def _req(n_in, max_new=128):
    r = Req(f"r{n_in}", list(range(n_in)), SamplingParams(max_new_tokens=max_new))
    r.init_next_round_input(prefix_len=0)
    return r


def test_prefill_adder_budgets_and_chunking():
    free = [100000]
    adder = PrefillAdder(lambda: free[0], [], 0.7, rem_input_tokens=16384, rem_chunk_tokens=8192)
    for _ in range(7):
        assert adder.add_one_req(_req(1024), False) == AddReqResult.CONTINUE
    # the 8th request exactly exhausts the chunk budget
    assert adder.add_one_req(_req(1024), False) == AddReqResult.OTHER
    assert len(adder.can_run_list) == 8 and adder.rem_chunk_tokens == 0 and adder.new_chunked_req is None
    # a request larger than the remaining chunk budget is truncated and becomes THE chunked request
    adder = PrefillAdder(lambda: free[0], [], 0.7, 16384, 8192)
    adder.add_one_req(_req(6000), False)
    big = _req(5000)
    res = adder.add_one_req(big, False)
    assert adder.new_chunked_req is big and big.extend_input_len == 2192 and len(big.fill_ids) == 2192
    assert res == AddReqResult.OTHER
    # continuing it next round: prefix = what was filled, rest fits → no longer chunked
    big.init_next_round_input(prefix_len=len(big.fill_ids))
    adder2 = PrefillAdder(lambda: free[0], [], 0.7, 16384, 8192)
    assert adder2.add_chunked_req(big) is None and big.extend_input_len == 5000 - 2192
    # KV budget: running requests reserve new_token_ratio * remaining tokens
    running = [_req(10) for _ in range(10)]
    small = PrefillAdder(lambda: 2000, running, 1.0, 16384, 8192)
    assert small.rem_total_tokens == 2000 - 10 * 128
    assert small.add_one_req(_req(1024), False) == AddReqResult.NO_TOKEN


def test_req_finish_and_retract_bookkeeping():
    r = Req("a", [1, 2, 3], SamplingParams(max_new_tokens=2))
    r.output_ids.append(5)
    r.check_finished()
    assert not r.finished()
    r.output_ids.append(6)
    r.check_finished()
    assert r.finished() and r.finished_reason == "length"
    e = Req("b", [1], SamplingParams(max_new_tokens=10, ignore_eos=False), eos_token_ids={9})
    e.output_ids.append(9)
    e.check_finished()
    assert e.finished_reason == "stop"
    e.req_pool_idx = 3
    e.reset_for_retract()
    assert e.req_pool_idx is None and e.is_retracted and e.prefix_indices == []


def test_bench_workload_generator_matches_bench_serving_recipe():
    import bench
    p = bench.make_prompts(4, 16, 128256, seed=1)
    rng = np.random.RandomState(1)
    off = rng.randint(0, 128256, size=4)
    assert p[2] == [int((off[2] + 2 + j) % 128256) for j in range(16)]
    arr = bench.poisson_arrivals(1000, 10.0, seed=1)
    gaps = np.diff([0.0] + arr)
    assert abs(gaps.mean() - 0.1) < 0.01 and all(g >= 0 for g in gaps)
    assert bench.poisson_arrivals(3, float("inf")) == [0.0, 0.0, 0.0]


def test_split_sms_normalises_the_reference_percentages():
    from semi_pd_b200.semi_pd_scheduler import split_sms
    assert split_sms(148, 50, 50) == (76, 72)
    assert split_sms(148, 80, 100) == (68, 80)          # the reference's default caps, normalised
    for p in (20, 35, 65, 80):
        ps, ds = split_sms(148, p, 100 - p)
        assert ps + ds == 148 and ds % 8 == 0 and ps >= 8 and ds >= 8
    assert split_sms(148, 1, 99) == (8, 140) and split_sms(148, 99, 1) == (140, 8)
    assert split_sms(132, 50, 50) == (68, 64)
    import pytest as _pt
    with _pt.raises(ValueError):
        split_sms(148, 0, 100)


def test_decode_scheduler_services_its_inbox_while_a_step_is_running():
    """`_drain_results(service_inbox=True)`: while the oldest launched step's event has not fired
    the decode scheduler admits requests / takes prefill reports that arrive, and only then reads
    the step's ids (TTFT does not pay half a decode step twice).  Driven on CPU with fake events."""
    import collections
    import queue
    import threading

    from semi_pd_b200.semi_pd_scheduler import SemiPDDecodeScheduler

    class FakeEvent:
        def __init__(self, polls_until_done):
            self.left = polls_until_done

        def query(self):
            self.left -= 1
            return self.left < 0

    s = SemiPDDecodeScheduler.__new__(SemiPDDecodeScheduler)
    s.tp_size, s.inbox, s.stop_flag = 1, queue.Queue(), threading.Event()
    s.result_queue = collections.deque()
    log = []
    s.recv_requests = lambda: log.append(("recv", s.inbox.get_nowait()))
    s.maybe_schedule_prefill = lambda: log.append(("schedule",))
    s.process_batch_result_decode = lambda reqs, pinned, bs, ev: log.append(("result", bs))
    # two launched steps; keep = 1 -> only the older one is drained
    s.result_queue.append((["r"], None, 3, FakeEvent(polls_until_done=4)))
    s.result_queue.append((["r"], None, 5, FakeEvent(polls_until_done=10 ** 9)))
    s.inbox.put("new request")
    s._drain_results(keep=1, service_inbox=True)
    assert log == [("recv", "new request"), ("schedule",), ("result", 3)]
    assert len(s.result_queue) == 1
    # without servicing (retraction path / TP lockstep) the inbox is left alone
    log.clear()
    s.inbox.put("later")
    s._drain_results(all_of_them=True)
    assert log == [("result", 5)] and s.inbox.qsize() == 1
    # under TP the inbox is only ever processed in lockstep, once per iteration
    s.tp_size = 2
    s.result_queue.append((["r"], None, 7, FakeEvent(polls_until_done=3)))
    log.clear()
    s._drain_results(all_of_them=True, service_inbox=True)
    assert log == [("result", 7)] and s.inbox.qsize() == 1


def test_http_request_model_parses_token_id_requests():
    """POST /generate carries token ids (launch_server.GenerateReq, module-level so that FastAPI
    can resolve it under postponed annotations)."""
    from fastapi import FastAPI
    from starlette.testclient import TestClient

    from semi_pd_b200.launch_server import GenerateReq, prepare_server_args

    app = FastAPI()

    @app.post("/generate")
    def generate(r: GenerateReq):
        return {"n": len(r.input_ids), "max_new_tokens": (r.sampling_params or {}).get("max_new_tokens")}

    c = TestClient(app)
    assert c.post("/generate", json={"input_ids": [1, 2, 3], "sampling_params": {"max_new_tokens": 4}}).json() == \
        {"n": 3, "max_new_tokens": 4}
    assert c.post("/generate", json={"input_ids": "abc"}).status_code == 422
    sa = prepare_server_args(["--model-path", "llama-3-8b", "--enable-semi-pd", "--tp", "4"])
    assert sa.tp_size == 4 and sa.enable_semi_pd and sa.disable_radix_cache and sa.argv[-1] == "4"
