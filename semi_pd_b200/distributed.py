"""Tensor-parallel plumbing: one process per GPU, `torch.distributed` for rendezvous and the large
collectives (NCCL over NVLink 5 / NVSwitch), libspd's peer-memory all-reduce for everything a
decode step or a small prefill batch sends.  Mirrors the reference's communication_op surface
(python/sglang/srt/distributed/communication_op.py:11-20 → parallel_state.py:376-436): the only
tensor collectives this build issues are the SUM all-reduces after `o_proj` and `down_proj`
(python/sglang/srt/layers/linear.py:1265-1266); the embedding table is replicated and the
vocab-sharded `lm_head` is followed by an all-gather of [bs, 2] (max, argmax) pairs — plus, outside
CUDA-graph capture, of the [bs, vocab/tp] logits themselves as logits_processor.py:426-427 does.

Dispatch (parallel_state.py:376-436 / custom_all_reduce.py:447-480): messages up to
`SPD_AR_MAX_BYTES` (32 MiB) → `spd_allreduce` (one-shot / two-shot over IPC-mapped peer buffers,
graph-capturable); larger ones → the same kernel over slot-sized pieces (or, with SPD_TP_NCCL=1,
ncclAllReduce on the worker's own process group).

Under semi-PD the prefill and the decode worker of a GPU are two threads that share the model
object, so the communicator cannot hang off the layers: each worker owns a `TpComm`
(its own NCCL group + its own libspd communicator, like the reference's separate NCCL worlds for
P and D, managers/scheduler.py:250-259) and installs it as the calling thread's current
communicator before every forward.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import torch
import torch.distributed as dist

from ._lib import check, lib

_TP_RANK = 0
_TP_SIZE = 1
_CPU_GROUP: Optional[dist.ProcessGroup] = None
_DEFAULT_COMM: Optional["TpComm"] = None
_tls = threading.local()

SPD_AR_MAX_BYTES = int(os.environ.get("SPD_AR_MAX_BYTES", 32 << 20))
_DT = {torch.bfloat16: 0, torch.float16: 1}
# SPD_TP_NCCL=1: messages above SPD_AR_MAX_BYTES and the full-logits all-gather go through
# torch.distributed's NCCL group instead of slot-sized pieces of the peer-memory kernels (A/B only)
NCCL_FOR_LARGE = os.environ.get("SPD_TP_NCCL", "0") == "1"
# the all-reduce after o_proj / down_proj fused with the residual-add RMSNorm that consumes it
# (decode-sized messages: every rank reads all W slots, so the one-shot form stops paying at ~2 MiB).
# Bit-identical to the two kernels and tested, but measured NEUTRAL on 2 GPUs (TP = 2 engine, 8B:
# TPOT p50 at 40 req/s 4.31 ms fused vs 4.22 ms unfused, profiles/r02_bench_n2_fused_ar_norm_ab.txt:
# the cross-GPU barrier dominates both, and under PDL the separate RMSNorm already starts inside
# the all-reduce's tail) — so it is opt-in: SPD_FUSE_AR_NORM=1.
FUSE_AR_NORM = os.environ.get("SPD_FUSE_AR_NORM", "0") == "1"
FUSE_AR_NORM_MAX_BYTES = int(os.environ.get("SPD_FUSE_AR_NORM_MAX_BYTES", 2 << 20))


def init_distributed_environment(tp_size: int, tp_rank: int, backend: str = "nccl",
                                 init_method: Optional[str] = None, device_id: Optional[int] = None):
    """model_runner.py:285-344 analogue.  With torchrun the env:// rendezvous is used."""
    global _TP_RANK, _TP_SIZE, _CPU_GROUP
    _TP_RANK, _TP_SIZE = tp_rank, tp_size
    if tp_size == 1:
        return
    if not dist.is_initialized():
        kwargs = {}
        if backend == "nccl" and device_id is not None:
            kwargs["device_id"] = torch.device("cuda", device_id)
        dist.init_process_group(backend=backend, init_method=init_method or "env://",
                                world_size=tp_size, rank=tp_rank, **kwargs)
    if _CPU_GROUP is None:
        # host-side object exchange (IPC handles, the decode scheduler's per-iteration message
        # broadcast): gloo, as the reference's tp_cpu_group (parallel_state.py:213-222)
        _CPU_GROUP = (dist.group.WORLD if dist.get_backend() == "gloo"
                      else dist.new_group(ranks=list(range(tp_size)), backend="gloo"))


def get_tensor_model_parallel_world_size() -> int:
    return _TP_SIZE


def get_tensor_model_parallel_rank() -> int:
    return _TP_RANK


def get_tp_cpu_group():
    return _CPU_GROUP


class TpComm:
    """One worker's tensor-parallel communicator."""

    def __init__(self, device: torch.device, backend: str = "nccl", max_bytes: int = SPD_AR_MAX_BYTES,
                 use_spd: Optional[bool] = None):
        assert _TP_SIZE > 1
        self.rank, self.world, self.device = _TP_RANK, _TP_SIZE, device
        # every TpComm gets its own torch process group (communicator) over the same ranks
        self.group = dist.new_group(ranks=list(range(_TP_SIZE)), backend=backend)
        if backend == "nccl":
            # build the NCCL communicator now, from the constructing thread: its lazy first-use
            # rendezvous must not happen inside a worker loop while the other worker is mid-step
            with torch.cuda.device(device):
                t = torch.zeros(8, device=device)
                dist.all_reduce(t, group=self.group)
                torch.cuda.synchronize()
        self._h = None
        self.max_bytes = 0
        if use_spd is None:
            use_spd = backend == "nccl" and os.environ.get("SPD_DISABLE_CUSTOM_AR", "0") != "1"
        if use_spd:
            h = C.c_void_p()
            handle = C.create_string_buffer(64)
            with torch.cuda.device(device):
                check(lib.spd_comm_create(self.rank, self.world, max_bytes, C.byref(h), handle),
                      "spd_comm_create")
                handles = [None] * self.world
                dist.all_gather_object(handles, handle.raw, group=_CPU_GROUP)
                blob = b"".join(handles)
                check(lib.spd_comm_connect(h, blob), "spd_comm_connect")
            self._h = h
            self.max_bytes = max_bytes
            dist.barrier(group=_CPU_GROUP)  # every rank has mapped every peer before the first call

    # ---- collectives ---------------------------------------------------------------------------
    def all_reduce(self, x: torch.Tensor, algo: int = 0) -> torch.Tensor:
        nbytes = x.numel() * x.element_size()
        spd_ok = (self._h is not None and x.dtype in _DT and x.numel() % 8 == 0 and x.is_contiguous())
        if spd_ok and nbytes <= self.max_bytes:
            check(lib.spd_allreduce(self._h, x.data_ptr(), x.data_ptr(), x.numel(), _DT[x.dtype], algo,
                                    torch.cuda.current_stream().cuda_stream), "spd_allreduce")
            return x
        if spd_ok and not NCCL_FOR_LARGE:
            # a large prefill batch: the same peer-memory kernel over slot-sized pieces (two-shot,
            # stream-ordered) — no second communication library on the worker's stream
            flat, step, st = x.view(-1), self.max_bytes // x.element_size(), torch.cuda.current_stream().cuda_stream
            for off in range(0, flat.numel(), step):
                n = min(step, flat.numel() - off)
                ptr = flat.data_ptr() + off * x.element_size()
                check(lib.spd_allreduce(self._h, ptr, ptr, n, _DT[x.dtype], algo, st), "spd_allreduce")
            return x
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)
        return x

    def all_reduce_add_rmsnorm(self, x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor,
                               eps: float, force: bool = False) -> bool:
        """x ← RMSNorm(residual + Σ_ranks x) · weight, residual ← residual + Σ_ranks x, in ONE
        kernel (`spd_allreduce_add_rmsnorm`); bit-identical to `all_reduce` followed by
        `ops.fused_add_rmsnorm`.  Returns False — and does nothing — when the fused form does not
        apply (no peer-memory communicator, a prefill-sized message, an odd layout): the caller
        then runs the two steps."""
        if self._h is None or not (FUSE_AR_NORM or force) or x.dtype not in _DT or x.dim() != 2:
            return False
        rows, dim = x.shape
        nbytes = rows * dim * x.element_size()
        if (nbytes > min(self.max_bytes, FUSE_AR_NORM_MAX_BYTES) or dim % 8 or dim > 8192
                or not (x.is_contiguous() and residual.is_contiguous() and weight.is_contiguous())
                or residual.shape != x.shape or residual.dtype != x.dtype or weight.dtype != x.dtype):
            return False
        check(lib.spd_allreduce_add_rmsnorm(self._h, x.data_ptr(), x.data_ptr(), residual.data_ptr(),
                                            weight.data_ptr(), rows, dim, float(eps), _DT[x.dtype],
                                            torch.cuda.current_stream().cuda_stream),
              "spd_allreduce_add_rmsnorm")
        return True

    def all_gather(self, x: torch.Tensor) -> torch.Tensor:
        """[rows, n] per rank → [rows, tp * n] (last-dim concat, rank order)."""
        inp = x.contiguous()
        rows = inp.shape[0]
        row_bytes = (inp.numel() // max(rows, 1)) * inp.element_size()
        if (self._h is not None and not NCCL_FOR_LARGE and rows > 0 and row_bytes % 16 == 0
                and row_bytes <= self.max_bytes):
            flat = torch.empty((self.world,) + tuple(inp.shape), dtype=inp.dtype, device=inp.device)
            step, st = max(1, self.max_bytes // row_bytes), torch.cuda.current_stream().cuda_stream
            for r0 in range(0, rows, step):
                piece = inp[r0:r0 + step]
                tmp = flat if step >= rows else torch.empty((self.world,) + tuple(piece.shape),
                                                            dtype=inp.dtype, device=inp.device)
                check(lib.spd_allgather(self._h, tmp.data_ptr(), piece.data_ptr(),
                                        piece.numel() * inp.element_size(), st), "spd_allgather")
                if tmp is not flat:
                    flat[:, r0:r0 + step].copy_(tmp)
        else:
            flat = torch.empty((self.world,) + tuple(inp.shape), dtype=inp.dtype, device=inp.device)
            dist.all_gather_into_tensor(flat.view(self.world * inp.shape[0], *inp.shape[1:]), inp,
                                        group=self.group)
        return flat.movedim(0, -2).reshape(*x.shape[:-1], self.world * x.shape[-1])

    def all_gather_rows16(self, x: torch.Tensor) -> torch.Tensor:
        """x: [rows, 16] uint8 per rank → [tp, rows, 16] (rank order) through libspd's peer-memory
        all-gather: stream-ordered and graph-capturable, unlike the NCCL path."""
        assert x.dtype == torch.uint8 and x.dim() == 2 and x.shape[1] == 16 and x.is_contiguous()
        out = torch.empty((self.world,) + tuple(x.shape), dtype=torch.uint8, device=x.device)
        if self._h is not None and x.numel() <= self.max_bytes:
            check(lib.spd_allgather(self._h, out.data_ptr(), x.data_ptr(), x.numel(),
                                    torch.cuda.current_stream().cuda_stream), "spd_allgather")
        else:
            dist.all_gather_into_tensor(out.view(-1, 16), x, group=self.group)
        return out

    def status(self) -> int:
        if self._h is None:
            return 0
        t = C.c_int()
        check(lib.spd_comm_status(self._h, C.byref(t)), "spd_comm_status")
        return t.value

    def close(self):
        if self._h is not None:
            lib.spd_comm_destroy(self._h)
            self._h = None


def set_current_comm(comm: Optional[TpComm]) -> None:
    """Install the calling thread's communicator (each worker does this before a forward)."""
    _tls.comm = comm


def get_current_comm() -> Optional[TpComm]:
    return getattr(_tls, "comm", None) or _DEFAULT_COMM


def set_default_comm(comm: Optional[TpComm]) -> None:
    global _DEFAULT_COMM
    _DEFAULT_COMM = comm


def tensor_model_parallel_all_reduce(input_: torch.Tensor) -> torch.Tensor:
    if _TP_SIZE == 1:
        return input_
    comm = get_current_comm()
    if comm is None:
        dist.all_reduce(input_, op=dist.ReduceOp.SUM)
        return input_
    return comm.all_reduce(input_)


def tensor_model_parallel_all_gather(input_: torch.Tensor, dim: int = -1) -> torch.Tensor:
    if _TP_SIZE == 1:
        return input_
    assert dim in (-1, input_.dim() - 1)
    comm = get_current_comm()
    if comm is not None:
        return comm.all_gather(input_)
    inp = input_.contiguous()
    flat = torch.empty((_TP_SIZE * inp.shape[0],) + tuple(inp.shape[1:]), dtype=inp.dtype, device=inp.device)
    dist.all_gather_into_tensor(flat, inp)
    parts = flat.view((_TP_SIZE,) + tuple(inp.shape))
    return parts.movedim(0, -2).reshape(*input_.shape[:-1], _TP_SIZE * input_.shape[-1])
