"""Python operator surface over libspd.so.

Function names, argument meaning and error behaviour mirror the reference's `sgl_kernel` Python
API (sgl-kernel/python/sgl_kernel/elementwise.py:9-25,58-76,102-151) and the call sites in
`python/sglang/srt/layers/*`; each wrapper only unwraps `torch.Tensor`s into raw pointers and
launches on torch's current CUDA stream (as sgl-kernel/python/sgl_kernel/utils.py:21-22 does).
torch is plumbing here (device memory + streams); every computation runs in libspd.so.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import SPD_EPI_NONE, SPD_EPI_SILU_MUL, check, lib

_DTYPE = {torch.bfloat16: _lib.SPD_BF16, torch.float16: _lib.SPD_F16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DTYPE[t.dtype]
    except KeyError:
        raise RuntimeError(f"libspd supports bfloat16/float16 tensors, got {t.dtype}") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "libspd ops need CUDA tensors: there is no CPU path (got a tensor on %s)" % t.device
            )


# ------------------------------------------------------------------------------------------------
# SM budget of the calling worker (set by the semi-PD engine per thread: prefill / decode partition)
# ------------------------------------------------------------------------------------------------
import threading

_tls = threading.local()


def set_sm_budget(n: int) -> None:
    """SMs the current thread's kernels may assume (0 = whole device).  Persistent kernels size
    their grids with it; the green-context stream enforces it."""
    _tls.sm_budget = int(n)


def get_sm_budget() -> int:
    return getattr(_tls, "sm_budget", 0)


# ------------------------------------------------------------------------------------------------
# elementwise
# ------------------------------------------------------------------------------------------------
def rmsnorm(input: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(input, weight)
    if out is None:
        out = torch.empty_like(input)
    d = input.shape[-1]
    x2 = input.view(-1, d) if input.is_contiguous() else input
    if x2.dim() != 2 or x2.stride(1) != 1:
        raise RuntimeError("rmsnorm: input must be 2-D with a contiguous last dimension")
    check(lib.spd_rmsnorm(out.data_ptr(), x2.data_ptr(), weight.data_ptr(), x2.shape[0], d,
                          x2.stride(0), out.view(-1, d).stride(0), float(eps), _dt(input), _stream()),
          "spd_rmsnorm")
    return out


def fused_add_rmsnorm(input: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor,
                      eps: float = 1e-6) -> None:
    _cuda(input, residual, weight)
    if not (input.is_contiguous() and residual.is_contiguous()):
        raise RuntimeError("fused_add_rmsnorm: input and residual must be contiguous")
    d = input.shape[-1]
    check(lib.spd_fused_add_rmsnorm(input.data_ptr(), residual.data_ptr(), weight.data_ptr(),
                                    input.numel() // d, d, float(eps), _dt(input), _stream()),
          "spd_fused_add_rmsnorm")


def silu_and_mul(input: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(input)
    if input.shape[-1] * input.dtype.itemsize % 16 != 0:
        raise ValueError("The pointers must be multiple of 16 bytes.")
    d = input.shape[-1] // 2
    if out is None:
        out = torch.empty(input.shape[:-1] + (d,), device=input.device, dtype=input.dtype)
    if not input.is_contiguous():
        raise RuntimeError("silu_and_mul: input must be contiguous")
    check(lib.spd_silu_and_mul(out.data_ptr(), input.data_ptr(), input.numel() // (2 * d), d,
                               _dt(input), _stream()), "spd_silu_and_mul")
    return out


def layernorm(input: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch.nn.LayerNorm over the last dimension (OPT; transformers modeling_opt.py)."""
    _cuda(input, weight, bias)
    d = input.shape[-1]
    x2 = input.view(-1, d)
    if x2.stride(1) != 1:
        raise RuntimeError("layernorm: input must have a contiguous last dimension")
    if out is None:
        out = torch.empty((x2.shape[0], d), device=input.device, dtype=input.dtype)
    check(lib.spd_layernorm(out.data_ptr(), x2.data_ptr(), weight.data_ptr(), bias.data_ptr(), x2.shape[0], d,
                            x2.stride(0), out.stride(0), float(eps), _dt(input), _stream()), "spd_layernorm")
    return out.view(input.shape)


def bias_act_residual(x: torch.Tensor, bias: Optional[torch.Tensor] = None,
                      residual: Optional[torch.Tensor] = None, act: str = "none",
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = act(x + bias) + residual, act in {"none", "relu"}; x / residual contiguous."""
    _cuda(x)
    if not x.is_contiguous() or (residual is not None and not residual.is_contiguous()):
        raise RuntimeError("bias_act_residual: x and residual must be contiguous")
    d = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    check(lib.spd_bias_act_residual(out.data_ptr(), x.data_ptr(), bias.data_ptr() if bias is not None else None,
                                    residual.data_ptr() if residual is not None else None, x.numel() // d, d,
                                    {"none": 0, "relu": 1}[act], _dt(x), _stream()), "spd_bias_act_residual")
    return out


def apply_rope_with_cos_sin_cache_inplace(positions: torch.Tensor, query: torch.Tensor,
                                          key: torch.Tensor, head_size: int,
                                          cos_sin_cache: torch.Tensor, is_neox: bool = True) -> None:
    _cuda(positions, query, key, cos_sin_cache)
    if cos_sin_cache.dtype != torch.float32:
        raise ValueError("cos_sin_cache should be float32")
    if positions.dtype != torch.int64:
        positions = positions.long()
    T = query.shape[0]
    q2, k2 = query.view(T, -1), key.view(T, -1)
    check(lib.spd_rope_inplace(q2.data_ptr(), k2.data_ptr(), cos_sin_cache.data_ptr(),
                               positions.data_ptr(), T, q2.shape[1] // head_size,
                               k2.shape[1] // head_size, head_size, cos_sin_cache.shape[1],
                               q2.stride(0), k2.stride(0), 0 if is_neox else 1, _dt(query), _stream()),
          "spd_rope_inplace")


def rope_kvwrite(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor,
                 value: torch.Tensor, k_pool: torch.Tensor, v_pool: torch.Tensor,
                 out_cache_loc: torch.Tensor, head_size: int, cos_sin_cache: torch.Tensor,
                 is_neox: bool = True) -> None:
    """RoPE(q, k) in place + k_pool[loc] = k, v_pool[loc] = v in one launch."""
    _cuda(positions, query, key, value, k_pool, v_pool, out_cache_loc, cos_sin_cache)
    if cos_sin_cache.dtype != torch.float32:
        raise ValueError("cos_sin_cache should be float32")
    if positions.dtype != torch.int64 or out_cache_loc.dtype != torch.int64:
        raise RuntimeError("rope_kvwrite: positions and out_cache_loc must be int64")
    T = query.shape[0]
    q2, k2, v2 = query.view(T, -1), key.view(T, -1), value.view(T, -1)
    check(lib.spd_rope_kvwrite(q2.data_ptr(), k2.data_ptr(), v2.data_ptr(), k_pool.data_ptr(),
                               v_pool.data_ptr(), cos_sin_cache.data_ptr(), positions.data_ptr(),
                               out_cache_loc.data_ptr(), T, q2.shape[1] // head_size,
                               k2.shape[1] // head_size, head_size, cos_sin_cache.shape[1],
                               q2.stride(0), k2.stride(0), v2.stride(0), 0 if is_neox else 1,
                               _dt(query), _stream()), "spd_rope_kvwrite")


def kv_write(key: torch.Tensor, value: torch.Tensor, k_pool: torch.Tensor, v_pool: torch.Tensor,
             loc: torch.Tensor) -> None:
    _cuda(key, value, k_pool, v_pool, loc)
    T = key.shape[0]
    if T == 0:
        return
    k2, v2 = key.view(T, -1), value.view(T, -1)
    if loc.dtype != torch.int64:
        loc = loc.long()
    check(lib.spd_kv_write(k2.data_ptr(), v2.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(),
                           loc.data_ptr(), T, k2.shape[1], k2.stride(0), v2.stride(0), _dt(key),
                           _stream()), "spd_kv_write")


def gather_rows(table: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(table, idx)
    if idx.dtype != torch.int64:
        idx = idx.long()
    n = idx.numel()
    if out is None:
        out = torch.empty((n, table.shape[1]), device=table.device, dtype=table.dtype)
    check(lib.spd_gather_rows(out.data_ptr(), table.data_ptr(), idx.data_ptr(), n, table.shape[1],
                              table.shape[0], _dt(table), _stream()), "spd_gather_rows")
    return out


def argmax(logits: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(logits)
    if logits.dtype != torch.float32 or logits.dim() != 2 or logits.stride(1) != 1:
        raise RuntimeError("argmax: logits must be fp32 [B, V] with a contiguous last dimension")
    if out is None:
        out = torch.empty((logits.shape[0],), device=logits.device, dtype=torch.int32)
    check(lib.spd_argmax_f32(out.data_ptr(), logits.data_ptr(), logits.shape[0], logits.shape[1],
                             logits.stride(0), _stream()), "spd_argmax_f32")
    return out


def argmax_partial(logits: torch.Tensor, index_offset: int) -> torch.Tensor:
    """[B, 16] uint8 rows {f32 max, i32 index_offset + argmax, 0, 0} of one vocab shard."""
    _cuda(logits)
    if logits.dtype != torch.float32 or logits.dim() != 2 or logits.stride(1) != 1:
        raise RuntimeError("argmax_partial: logits must be fp32 [B, V] with a contiguous last dimension")
    out = torch.empty((logits.shape[0], 16), device=logits.device, dtype=torch.uint8)
    check(lib.spd_argmax_partial_f32(out.data_ptr(), logits.data_ptr(), logits.shape[0], logits.shape[1],
                                     logits.stride(0), int(index_offset), _stream()),
          "spd_argmax_partial_f32")
    return out


def argmax_merge(pairs: torch.Tensor) -> torch.Tensor:
    """pairs [parts, B, 16] uint8 (argmax_partial rows of every shard) → ids [B] int32."""
    _cuda(pairs)
    parts, B = pairs.shape[0], pairs.shape[1]
    out = torch.empty((B,), device=pairs.device, dtype=torch.int32)
    check(lib.spd_argmax_merge(out.data_ptr(), pairs.data_ptr(), parts, B, _stream()), "spd_argmax_merge")
    return out


# ------------------------------------------------------------------------------------------------
# metadata
# ------------------------------------------------------------------------------------------------
def compute_positions(extend_prefix_lens: torch.Tensor, extend_seq_lens: torch.Tensor,
                      extend_num_tokens: int, want_token_req: bool = False):
    """(positions i64 [T], extend_start_loc i32 [bs][, token_req i32 [T]]) —
    compute_position_triton's contract (forward_batch_info.py:393-419)."""
    _cuda(extend_prefix_lens, extend_seq_lens)
    bs = extend_seq_lens.shape[0]
    dev = extend_seq_lens.device
    positions = torch.empty((extend_num_tokens,), dtype=torch.int64, device=dev)
    start = torch.empty((bs,), dtype=torch.int32, device=dev)
    treq = torch.empty((extend_num_tokens,), dtype=torch.int32, device=dev) if want_token_req else None
    has_prefix = extend_prefix_lens is not None and extend_prefix_lens.shape[0] == bs
    check(lib.spd_compute_positions(positions.data_ptr(), start.data_ptr(),
                                    treq.data_ptr() if treq is not None else None,
                                    extend_prefix_lens.data_ptr() if has_prefix else None,
                                    extend_seq_lens.data_ptr(), bs, _stream()),
          "spd_compute_positions")
    return (positions, start, treq) if want_token_req else (positions, start)


def clamp_position(seq_lens: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(seq_lens)
    if seq_lens.dtype != torch.int64:
        seq_lens = seq_lens.long()
    if out is None:
        out = torch.empty_like(seq_lens)
    check(lib.spd_decode_positions(out.data_ptr(), seq_lens.data_ptr(), seq_lens.numel(), _stream()),
          "spd_decode_positions")
    return out


def create_kv_indices(req_to_token: torch.Tensor, req_pool_indices: torch.Tensor,
                      page_kernel_lens: torch.Tensor, kv_indptr: torch.Tensor,
                      kv_start_idx: Optional[torch.Tensor], kv_indices: torch.Tensor) -> None:
    """create_flashinfer_kv_indices_triton's contract (S/layers/attention/utils.py:5-39)."""
    _cuda(req_to_token, req_pool_indices, page_kernel_lens, kv_indptr, kv_indices)
    if req_pool_indices.dtype != torch.int64:
        req_pool_indices = req_pool_indices.long()
    if page_kernel_lens.dtype != torch.int32:
        page_kernel_lens = page_kernel_lens.int()
    check(lib.spd_build_kv_indices(kv_indices.data_ptr(), req_to_token.data_ptr(),
                                   req_pool_indices.data_ptr(), page_kernel_lens.data_ptr(),
                                   kv_indptr.data_ptr(),
                                   kv_start_idx.data_ptr() if kv_start_idx is not None else None,
                                   req_pool_indices.numel(), req_to_token.stride(0), _stream()),
          "spd_build_kv_indices")


def write_req_to_token_pool(req_to_token: torch.Tensor, req_pool_indices: torch.Tensor,
                            pre_lens: torch.Tensor, seq_lens: torch.Tensor,
                            extend_lens: torch.Tensor, out_cache_loc: torch.Tensor) -> None:
    """write_req_to_token_pool_triton's contract (S/managers/schedule_batch.py:1446-1480)."""
    _cuda(req_to_token, req_pool_indices, pre_lens, seq_lens, extend_lens, out_cache_loc)
    args = [t if t.dtype == torch.int64 else t.long()
            for t in (req_pool_indices, pre_lens, seq_lens, extend_lens, out_cache_loc)]
    check(lib.spd_write_req_to_token(req_to_token.data_ptr(), *[a.data_ptr() for a in args],
                                     args[0].numel(), req_to_token.stride(0), _stream()),
          "spd_write_req_to_token")


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
_gemm_ws = {}


def _gemm_workspace(device: torch.device) -> torch.Tensor:
    """Split-K scratch, one per (device, stream) so the prefill and decode workers never share it."""
    key = (device.index, _stream())
    ws = _gemm_ws.get(key)
    if ws is None:
        ws = torch.zeros((lib.spd_gemm_workspace_bytes(),), dtype=torch.uint8, device=device)
        _gemm_ws[key] = ws
    return ws


def linear(x: torch.Tensor, weight: torch.Tensor, out: Optional[torch.Tensor] = None,
           out_f32: bool = False, epilogue: int = SPD_EPI_NONE) -> torch.Tensor:
    """F.linear(x, weight) — x [M, K], weight [N, K] (or [2N, K] with SPD_EPI_SILU_MUL)."""
    _cuda(x, weight)
    if x.dim() != 2 or weight.dim() != 2 or x.stride(1) != 1 or weight.stride(1) != 1:
        raise RuntimeError("linear: x and weight must be 2-D with contiguous last dimension")
    if x.dtype != weight.dtype:
        raise RuntimeError(f"linear: dtype mismatch {x.dtype} vs {weight.dtype}")
    M, K = x.shape
    N = weight.shape[0] // (2 if epilogue == SPD_EPI_SILU_MUL else 1)
    if weight.shape[1] != K:
        raise RuntimeError(f"linear: shape mismatch x{tuple(x.shape)} w{tuple(weight.shape)}")
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    ws = _gemm_workspace(x.device)
    check(lib.spd_gemm(out.data_ptr(), x.data_ptr(), weight.data_ptr(), M, N, K, x.stride(0),
                       weight.stride(0), out.stride(0), _dt(x), 1 if out_f32 else 0, epilogue,
                       ws.data_ptr(), ws.numel(), get_sm_budget(), _stream()), "spd_gemm")
    return out


def linear_block_diag(x: torch.Tensor, weight: torch.Tensor, num_groups: int, k: int, n: int,
                      x_group_stride: int, out: Optional[torch.Tensor] = None,
                      out_group_stride: int = 0) -> torch.Tensor:
    """out[:, g*n:(g+1)*n] = x[:, g*x_group_stride : +k] @ weight[g*n:(g+1)*n, :k].T — the per-head
    products of the absorbed MLA form (torch.bmm(q_nope, w_kc), torch.bmm(attn_out, w_vc)) as one launch."""
    _cuda(x, weight)
    if x.dim() != 2 or weight.dim() != 2 or x.stride(1) != 1 or weight.stride(1) != 1 or x.dtype != weight.dtype:
        raise RuntimeError("linear_block_diag: x and weight must be 2-D, contiguous rows, same dtype")
    M = x.shape[0]
    if out is None:
        out = torch.empty((M, num_groups * n), device=x.device, dtype=x.dtype)
    check(lib.spd_gemm_grouped(out.data_ptr(), x.data_ptr(), weight.data_ptr(), M, n, k, x.stride(0),
                               weight.stride(0), out.stride(0), num_groups, None, x_group_stride, x.shape[1],
                               out_group_stride, 0, _dt(x), SPD_EPI_NONE, get_sm_budget(), _stream()),
          "spd_gemm_grouped")
    return out


def linear_grouped_rows(x_sorted: torch.Tensor, weight: torch.Tensor, tile_group: torch.Tensor, num_groups: int,
                        n: int, epilogue: int = SPD_EPI_NONE, k_concat: bool = False) -> torch.Tensor:
    """Rows of x_sorted are grouped by expert in 128-row tiles; weight = the experts' [n (or 2n), K]
    matrices stacked along the rows — or, k_concat, [n, E * K] concatenated along K.  → [rows, n]."""
    _cuda(x_sorted, weight, tile_group)
    M, K = x_sorted.shape
    out = torch.empty((M, n), device=x_sorted.device, dtype=x_sorted.dtype)
    check(lib.spd_gemm_grouped(out.data_ptr(), x_sorted.data_ptr(), weight.data_ptr(), M, n, K, x_sorted.stride(0),
                               weight.stride(0), out.stride(0), num_groups, tile_group.data_ptr(), 0, K, 0,
                               K if k_concat else 0, _dt(x_sorted), epilogue, get_sm_budget(), _stream()),
          "spd_gemm_grouped")
    return out


def moe_route(logits: torch.Tensor, top_k: int, scale: float = 1.0, renormalize: bool = False):
    """→ (topk_ids int32 [T, k], topk_weights fp32 [T, k])."""
    _cuda(logits)
    T, E = logits.shape
    ids = torch.empty((T, top_k), device=logits.device, dtype=torch.int32)
    w = torch.empty((T, top_k), device=logits.device, dtype=torch.float32)
    check(lib.spd_moe_route(ids.data_ptr(), w.data_ptr(), logits.contiguous().data_ptr(), T, E, top_k, float(scale),
                            1 if renormalize else 0, _dt(logits), _stream()), "spd_moe_route")
    return ids, w


def moe_align(topk_ids: torch.Tensor, num_experts: int):
    """→ (row_token [R], inv_pos [T*k], tile_group [R/128]) with R = spd_moe_align_rows()."""
    _cuda(topk_ids)
    T, k = topk_ids.shape
    R = lib.spd_moe_align_rows(T, k, num_experts)
    dev = topk_ids.device
    row_token = torch.empty((R,), device=dev, dtype=torch.int32)
    inv_pos = torch.empty((T * k,), device=dev, dtype=torch.int32)
    tile_group = torch.empty((R // 128,), device=dev, dtype=torch.int32)
    check(lib.spd_moe_align(row_token.data_ptr(), inv_pos.data_ptr(), tile_group.data_ptr(), topk_ids.data_ptr(),
                            T, k, num_experts, R, _stream()), "spd_moe_align")
    return row_token, inv_pos, tile_group


def moe_gather(x: torch.Tensor, row_token: torch.Tensor) -> torch.Tensor:
    _cuda(x, row_token)
    out = torch.empty((row_token.numel(), x.shape[1]), device=x.device, dtype=x.dtype)
    check(lib.spd_moe_gather(out.data_ptr(), x.data_ptr(), row_token.data_ptr(), row_token.numel(), x.shape[1],
                             x.stride(0), _dt(x), _stream()), "spd_moe_gather")
    return out


def moe_combine(y: torch.Tensor, inv_pos: torch.Tensor, topk_w: torch.Tensor,
                shared: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(y, inv_pos, topk_w)
    T, k = topk_w.shape
    out = torch.empty((T, y.shape[1]), device=y.device, dtype=y.dtype)
    check(lib.spd_moe_combine(out.data_ptr(), y.data_ptr(), shared.data_ptr() if shared is not None else None,
                              inv_pos.data_ptr(), topk_w.data_ptr(), T, k, y.shape[1], _dt(y), _stream()),
          "spd_moe_combine")
    return out


def moe_dense_act(gate_up: torch.Tensor, topk_ids: torch.Tensor, topk_w: torch.Tensor, num_experts: int,
                  inter: int) -> torch.Tensor:
    _cuda(gate_up, topk_ids, topk_w)
    T, k = topk_ids.shape
    out = torch.empty((T, num_experts * inter), device=gate_up.device, dtype=gate_up.dtype)
    check(lib.spd_moe_dense_act(out.data_ptr(), gate_up.data_ptr(), topk_ids.data_ptr(), topk_w.data_ptr(), T, k,
                                num_experts, inter, _dt(gate_up), _stream()), "spd_moe_dense_act")
    return out


def kv_write_single(rows: torch.Tensor, pool: torch.Tensor, loc: torch.Tensor) -> None:
    """pool[loc[t]] = rows[t] — MLATokenToKVPool.set_kv_buffer (one latent buffer per layer)."""
    _cuda(rows, pool, loc)
    T = rows.shape[0]
    if T == 0:
        return
    r2 = rows.view(T, -1)
    if loc.dtype != torch.int64:
        loc = loc.long()
    check(lib.spd_kv_write(r2.data_ptr(), None, pool.data_ptr(), None, loc.data_ptr(), T, r2.shape[1],
                           r2.stride(0), 0, _dt(rows), _stream()), "spd_kv_write")


def attn_mla(q: torch.Tensor, kv_pool: torch.Tensor, req_to_token, req_pool_indices, num_heads: int,
             kv_lora_rank: int, rope_dim: int, sm_scale: float, workspace: torch.Tensor, seq_lens=None,
             prefix_lens=None, extend_start_loc=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Absorbed MLA attention: decode (seq_lens) or ragged extend (prefix_lens + extend_start_loc).
    q [T, H * (kv_lora + rope)] → out [T, H * kv_lora]."""
    _cuda(q, kv_pool, req_to_token, req_pool_indices)
    T = q.shape[0]
    q2 = q.view(T, -1)
    if out is None:
        out = torch.empty((T, num_heads * kv_lora_rank), device=q.device, dtype=q.dtype)
    check(lib.spd_attn_mla(out.data_ptr(), q2.data_ptr(), kv_pool.data_ptr(), req_to_token.data_ptr(),
                           req_pool_indices.data_ptr(), seq_lens.data_ptr() if seq_lens is not None else None,
                           prefix_lens.data_ptr() if prefix_lens is not None else None,
                           extend_start_loc.data_ptr() if extend_start_loc is not None else None,
                           req_pool_indices.numel(), T, num_heads, kv_lora_rank, rope_dim, q2.stride(0),
                           out.stride(0), req_to_token.stride(0), float(sm_scale), _dt(q), workspace.data_ptr(),
                           workspace.numel(), get_sm_budget(), _stream()), "spd_attn_mla")
    return out


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
_attn_ws = {}


def _attn_workspace(device, batch: int, hq: int, d: int) -> torch.Tensor:
    need = lib.spd_attn_decode_workspace_bytes(batch, hq, d)
    key = (device.index, _stream())
    ws = _attn_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros((max(need, 1 << 20),), dtype=torch.uint8, device=device)
        _attn_ws[key] = ws
    return ws


def attn_decode(q, k_pool, v_pool, req_to_token, req_pool_indices, seq_lens, num_q_heads,
                num_kv_heads, head_dim, sm_scale, out=None, impl: str = "tc", workspace=None):
    _cuda(q, k_pool, v_pool, req_to_token, req_pool_indices, seq_lens)
    B = q.shape[0]
    q2 = q.view(B, -1)
    if out is None:
        out = torch.empty_like(q2)
    if req_pool_indices.dtype != torch.int64 or seq_lens.dtype != torch.int64:
        raise RuntimeError("attn_decode: req_pool_indices and seq_lens must be int64")
    ws = workspace if workspace is not None else _attn_workspace(q.device, B, num_q_heads, head_dim)
    if impl == "tc":
        check(lib.spd_attn_decode_paged(out.data_ptr(), q2.data_ptr(), k_pool.data_ptr(),
                                        v_pool.data_ptr(), req_to_token.data_ptr(),
                                        req_pool_indices.data_ptr(), seq_lens.data_ptr(), B,
                                        num_q_heads, num_kv_heads, head_dim, q2.stride(0),
                                        out.stride(0), req_to_token.stride(0), float(sm_scale),
                                        _dt(q), ws.data_ptr(), ws.numel(), get_sm_budget(), _stream()),
              "spd_attn_decode_paged")
    elif impl == "simt":
        check(lib.spd_attn_decode_simt(out.data_ptr(), q2.data_ptr(), k_pool.data_ptr(),
                                       v_pool.data_ptr(), req_to_token.data_ptr(),
                                       req_pool_indices.data_ptr(), seq_lens.data_ptr(), B,
                                       num_q_heads, num_kv_heads, head_dim, q2.stride(0),
                                       out.stride(0), req_to_token.stride(0), float(sm_scale),
                                       _dt(q), ws.data_ptr(), ws.numel(), _stream()),
              "spd_attn_decode_simt")
    else:
        raise ValueError(impl)
    return out


def attn_prefill(q, k_pool, v_pool, req_to_token, req_pool_indices, prefix_lens, extend_lens,
                 extend_start_loc, max_extend_len, num_q_heads, num_kv_heads, head_dim, sm_scale,
                 out=None, impl: str = "tc"):
    _cuda(q, k_pool, v_pool, req_to_token, req_pool_indices, prefix_lens, extend_lens, extend_start_loc)
    T = q.shape[0]
    q2 = q.view(T, -1)
    if out is None:
        out = torch.empty_like(q2)
    if req_pool_indices.dtype != torch.int64:
        raise RuntimeError("attn_prefill: req_pool_indices must be int64")
    for t in (prefix_lens, extend_lens, extend_start_loc):
        if t.dtype != torch.int32:
            raise RuntimeError("attn_prefill: prefix_lens / extend_lens / extend_start_loc must be int32")
    args = (out.data_ptr(), q2.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(),
            req_to_token.data_ptr(), req_pool_indices.data_ptr(), prefix_lens.data_ptr(),
            extend_lens.data_ptr(), extend_start_loc.data_ptr(), req_pool_indices.numel(), T,
            int(max_extend_len), num_q_heads, num_kv_heads, head_dim, q2.stride(0), out.stride(0),
            req_to_token.stride(0), float(sm_scale), _dt(q))
    if impl == "tc":
        check(lib.spd_attn_prefill_ragged(*args, get_sm_budget(), _stream()), "spd_attn_prefill_tc")
    else:
        check(lib.spd_attn_prefill_simt(*args, _stream()), "spd_attn_prefill_simt")
    return out


# ------------------------------------------------------------------------------------------------
# SM partitions (green contexts)
# ------------------------------------------------------------------------------------------------
import ctypes as _C


class SmPartition:
    """Two disjoint SM sets of one GPU with a stream each — the in-process replacement for the
    per-process MPS caps (S/entrypoints/engine.py:591-593,632-634)."""

    def __init__(self, device: int, prefill_sms: int, decode_sms: int):
        h = _C.c_void_p()
        check(lib.spd_partition_create(device, prefill_sms, decode_sms, _C.byref(h)),
              "spd_partition_create")
        self._h = h
        a, b = _C.c_int(), _C.c_int()
        check(lib.spd_partition_sm_counts(h, _C.byref(a), _C.byref(b)), "spd_partition_sm_counts")
        self.prefill_sms, self.decode_sms = a.value, b.value
        s0, s1 = _C.c_void_p(), _C.c_void_p()
        check(lib.spd_partition_stream(h, 0, _C.byref(s0)), "spd_partition_stream")
        check(lib.spd_partition_stream(h, 1, _C.byref(s1)), "spd_partition_stream")
        self.prefill_stream = torch.cuda.ExternalStream(s0.value, device=device)
        self.decode_stream = torch.cuda.ExternalStream(s1.value, device=device)

    def close(self):
        if self._h is not None:
            lib.spd_partition_destroy(self._h)
            self._h = None


_partitions = {}


def get_partition(device: int, prefill_sms: int, decode_sms: int) -> SmPartition:
    """Process-lifetime pool of SM partitions.  Green contexts and their streams are created once
    per (device, split) and REUSED by every engine built afterwards: torch's caching allocator
    and CUDA graphs keep referring to a stream long after the engine that used it is gone, so a
    partition is never destroyed under them (doing so was measured to abort the process);
    `SmPartition.close()` / spd_partition_destroy remain for embedders that own the streams'
    whole lifetime."""
    key = (int(device), int(prefill_sms), int(decode_sms))
    part = _partitions.get(key)
    if part is None:
        part = _partitions[key] = SmPartition(device, prefill_sms, decode_sms)
    return part


def device_sm_count(device: int = 0) -> int:
    n = _C.c_int()
    check(lib.spd_device_sm_count(device, _C.byref(n)), "spd_device_sm_count")
    return n.value
