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
