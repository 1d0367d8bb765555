"""world_size-2 `gloo` test (CPU) of the N>1 plumbing: the TP communication ops of
semi_pd_b200.distributed and the sharding rules of the parallel linear layers reproduce the
unsharded oracle layer when each rank computes its shard and the row-parallel partial products are
all-reduced (the only collectives the path issues: after o_proj and down_proj, linear.py:1265-1266;
plus the logits all-gather, logits_processor.py:426-427)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.nn.functional as F

    from oracle import llama_oracle as O
    from semi_pd_b200 import distributed as D
    from semi_pd_b200.layers import (MergedColumnParallelLinear, ParallelLMHead, QKVParallelLinear,
                                     RowParallelLinear)
    D.init_distributed_environment(world, rank, backend="gloo")
    assert D.get_tensor_model_parallel_world_size() == world and D.get_tensor_model_parallel_rank() == rank
    torch.manual_seed(0)  # same "checkpoint" on every rank
    H, Dh, Hq, Hk, I, V, T = 64, 16, 4, 2, 96, 128, 7
    x = torch.randn(T, H)
    wqkv, wo = torch.randn((Hq + 2 * Hk) * Dh, H) / 8, torch.randn(H, Hq * Dh) / 8
    wgu, wd, lm = torch.randn(2 * I, H) / 8, torch.randn(H, I) / 8, torch.randn(V, H) / 8
    # --- attention block with sharded heads (attention itself = per-head SDPA over the T tokens)
    nq, nk = Hq // world, Hk // world
    qkv = F.linear(x, QKVParallelLinear.shard(wqkv, Dh, Hq, Hk, rank, world))
    q, k, v = qkv.split([nq * Dh, nk * Dh, nk * Dh], -1)
    o = F.scaled_dot_product_attention(q.view(T, nq, Dh).transpose(0, 1)[None], k.view(T, nk, Dh).transpose(0, 1)[None],
                                       v.view(T, nk, Dh).transpose(0, 1)[None], is_causal=True, enable_gqa=True)
    o = o[0].transpose(0, 1).reshape(T, nq * Dh)
    attn = D.tensor_model_parallel_all_reduce(F.linear(o, RowParallelLinear.shard(wo, rank, world)))
    # --- MLP block
    act = O.silu_and_mul(F.linear(x, MergedColumnParallelLinear.shard(wgu, rank, world)))
    mlp = D.tensor_model_parallel_all_reduce(F.linear(act, RowParallelLinear.shard(wd, rank, world)))
    # --- vocab-sharded logits, all-gathered
    logits = D.tensor_model_parallel_all_gather(F.linear(x, ParallelLMHead.shard(lm, rank, world)))
    # --- unsharded reference
    qf, kf, vf = F.linear(x, wqkv).split([Hq * Dh, Hk * Dh, Hk * Dh], -1)
    of = F.scaled_dot_product_attention(qf.view(T, Hq, Dh).transpose(0, 1)[None], kf.view(T, Hk, Dh).transpose(0, 1)[None],
                                        vf.view(T, Hk, Dh).transpose(0, 1)[None], is_causal=True, enable_gqa=True)
    attn_ref = F.linear(of[0].transpose(0, 1).reshape(T, Hq * Dh), wo)
    mlp_ref = F.linear(O.silu_and_mul(F.linear(x, wgu)), wd)
    ok = (torch.allclose(attn, attn_ref, atol=1e-4) and torch.allclose(mlp, mlp_ref, atol=1e-4)
          and torch.allclose(logits, F.linear(x, lm), atol=1e-5) and logits.shape == (T, V))
    ret[rank] = bool(ok)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_tp2_gloo_matches_unsharded():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def _lockstep_worker(rank, world, port, ret):
    """Drives SemiPDDecodeScheduler.recv_requests — the per-iteration message broadcast that keeps
    the decode schedulers of all TP ranks in lockstep (scheduler.py:599-659) — on CPU: the class
    is instantiated without its GPU members and the two handlers are recorded instead of run."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import queue
    import threading

    import torch.distributed as dist

    from semi_pd_b200.io_struct import BatchProcessPrefillResultReq, TokenizedGenerateReqInput
    from semi_pd_b200.schedule_batch import SamplingParams
    from semi_pd_b200.semi_pd_scheduler import SemiPDDecodeScheduler

    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = SemiPDDecodeScheduler.__new__(SemiPDDecodeScheduler)
    s.tp_rank, s.tp_size, s.cpu_group = rank, world, dist.group.WORLD
    s.inbox, s.local_prefill_results, s.send_to_p_instance = queue.Queue(), queue.Queue(), queue.Queue()
    s.stop_flag, s._stop_broadcast = threading.Event(), False
    log = []
    s.handle_generate_request = lambda m: log.append(("gen", m.rid, tuple(m.input_ids), m.sampling_params.max_new_tokens))
    s.process_prefill_result = lambda m: log.append(("prefill_result", tuple(m.next_token_ids)))
    if rank == 0:
        # iteration 1: two requests; iteration 2: nothing; iteration 3: a prefill report and a
        # request; iteration 4: stop (set before the iteration starts)
        script = [[TokenizedGenerateReqInput(rid="a", input_text="", input_ids=[1, 2, 3],
                                             sampling_params=SamplingParams(max_new_tokens=5), arrival_time=0.0),
                   TokenizedGenerateReqInput(rid="b", input_text="", input_ids=[4],
                                             sampling_params=SamplingParams(max_new_tokens=2), arrival_time=0.0)],
                  [],
                  [BatchProcessPrefillResultReq(next_token_ids=[7, 8]),
                   TokenizedGenerateReqInput(rid="c", input_text="", input_ids=[9, 9],
                                             sampling_params=SamplingParams(max_new_tokens=1), arrival_time=0.0)]]
    else:
        script = [[], [], []]
        # this rank's own prefill worker reports the same batch (its ids are identical)
        s.local_prefill_results.put(BatchProcessPrefillResultReq(next_token_ids=[7, 8]))
    iters = 0
    for msgs in script:
        for m in msgs:
            s.inbox.put(m)
        s.recv_requests()
        iters += 1
        assert not s._stop_broadcast
    if rank == 0:
        s.stop_flag.set()
    s.recv_requests()
    iters += 1
    ret[rank] = (log, iters, s._stop_broadcast, s.stop_flag.is_set(),
                 [m.rid for m in list(s.send_to_p_instance.queue)])
    dist.barrier()
    dist.destroy_process_group()


def test_tp2_decode_schedulers_process_identical_messages_and_stop_together():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_lockstep_worker, args=(world, port, ret), nprocs=world, join=True)
        out = dict(ret)
    log0, it0, stop0, _, _ = out[0]
    log1, it1, stop1, flag1, to_p1 = out[1]
    assert log0 == log1 == [("gen", "a", (1, 2, 3), 5), ("gen", "b", (4,), 2), ("prefill_result", (7, 8)),
                            ("gen", "c", (9, 9), 1)]
    assert it0 == it1 == 4 and stop0 and stop1 and flag1   # both leave in the SAME iteration
    assert to_p1 == ["a", "b", "c"]                        # rank 1's prefill worker learns every request
