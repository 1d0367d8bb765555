"""KV-cache storage and slot allocators — the classes the reference's
python/sglang/srt/mem_cache/memory_pool.py exposes (`ReqToTokenPool` :46-96,
`TokenToKVPoolAllocator` :124-184, `MHATokenToKVPool` :187-346), with two structural changes
that are the point of this build:

1. ONE pool per GPU that the prefill and the decode worker address directly (they are threads of
   one process on disjoint SM partitions), so `bypass_create_buffers` / the cudaIpcMemHandle
   export-import of every buffer (model_runner.py:346-624, semi-pd-ipc/ipc.cpp) has nothing to do.

2. The allocator hands out RUNS.  The external contract stays the reference's token-granular one
   (page_size = 1, int64 slot ids, slot 0 = dummy slot, `alloc / free / available_size`), but
   internally the pool is carved into runs of `run_len` (128) consecutive slots and a request's
   tokens occupy its runs in position order: positions [k·128, k·128 + 128) of a request are 128
   CONSECUTIVE slots.  A 128-token K/V tile of one kv head is then a strided 2-D box of the pool
   (row stride H_kv·D·2 B) that the attention kernels fetch with one TMA tensor copy instead of
   128 scattered 256-byte gathers; they verify the contiguity per tile on the device and fall back
   to the gather for anything else (a foreign `req_to_token`, the reference's scattered layout).
   The request-aware entry points mirror the reference's paged allocator
   (mem_cache/paged_allocator.py: `alloc_extend(prefix_lens, seq_lens, last_loc, n)`,
   `alloc_decode(seq_lens, last_loc)`), with the bookkeeping on the host so no call synchronises
   with the device.

HBM layout: per layer, `k_buffer[l]` and `v_buffer[l]` are [size + 1, H_kv, D] 16-bit; all layers
live in one allocation [2, L, size + 1, H_kv, D] so a 180 GB part is carved once.
"""
from __future__ import annotations

import abc
import collections
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops

GB = 1024 ** 3
DEFAULT_RUN_LEN = 128


class ReqToTokenPool:
    """Maps a request slot to the KV slots of its tokens: `req_to_token[req, pos]` (int32)."""

    def __init__(self, size: int, max_context_len: int, device: str,
                 enable_memory_saver: bool = False, bypass_create_buffers: bool = False):
        self.size = size
        self.max_context_len = max_context_len
        self.device = device
        # one tensor both workers address; nothing to bypass in a single process
        self.req_to_token = torch.zeros((size, max_context_len), dtype=torch.int32, device=device)
        self.clear()

    def write(self, indices, values):
        self.req_to_token[indices] = values

    def available_size(self) -> int:
        return len(self._free)

    def alloc(self, need_size: int) -> Optional[List[int]]:
        if need_size > len(self._free):
            return None
        out = [self._free.popleft() for _ in range(need_size)]
        self._live.update(out)
        return out

    def free(self, free_index: Union[int, List[int]]):
        for i in ([free_index] if isinstance(free_index, int) else free_index):
            if i not in self._live:
                raise RuntimeError(f"request slot {i} freed twice")
            self._live.discard(i)
            self._free.append(i)

    def clear(self):
        self._free = collections.deque(range(self.size))
        self._live = set()

    @property
    def free_slots(self) -> List[int]:
        return list(self._free)


class KVCache(abc.ABC):
    @abc.abstractmethod
    def get_key_buffer(self, layer_id: int) -> torch.Tensor: ...

    @abc.abstractmethod
    def get_value_buffer(self, layer_id: int) -> torch.Tensor: ...

    @abc.abstractmethod
    def get_kv_buffer(self, layer_id: int) -> Tuple[torch.Tensor, torch.Tensor]: ...

    @abc.abstractmethod
    def set_kv_buffer(self, layer, loc: torch.Tensor, cache_k: torch.Tensor,
                      cache_v: torch.Tensor) -> None: ...


class _PinnedRing:
    """A few reusable pinned staging buffers for the slot-id uploads (a fresh pin_memory()
    allocation costs more than the copy it feeds)."""

    def __init__(self, n: int = 8, elems: int = 16384):
        self.elems = elems
        self.bufs = None
        self.n = n
        self.i = 0

    def stage(self, arr: np.ndarray, device) -> torch.Tensor:
        if not torch.cuda.is_available():
            return torch.from_numpy(arr.copy())
        if arr.size > self.elems:
            return torch.from_numpy(arr).pin_memory().to(device, non_blocking=True)
        if self.bufs is None:
            self.bufs = [torch.empty((self.elems,), dtype=torch.int64).pin_memory() for _ in range(self.n)]
            self.events = [None] * self.n
        k = self.i % self.n
        self.i += 1
        if self.events[k] is not None:
            self.events[k].synchronize()  # the copy that last used this buffer has drained
        buf = self.bufs[k][: arr.size]
        buf.numpy()[:] = arr
        out = buf.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev
        return out


class TokenToKVPoolAllocator:
    """Run allocator behind the token-granular contract (see the module docstring).

    Slots 1 … size are grouped into `size // run_len` runs (run r = slots 1 + r·run_len …); the
    tail that does not fill a run stays unused.  `used[r]` counts the slots of run r that are
    handed out; a run returns to the free list when its count reaches zero.
    """

    def __init__(self, size: int, dtype: torch.dtype, device: str, kvcache: KVCache,
                 run_len: int = DEFAULT_RUN_LEN):
        self.size = size
        self.dtype = dtype
        self.device = device
        self.page_size = 1          # the external contract: every slot id is individually addressable
        self.run_len = int(run_len)
        self.num_runs = size // self.run_len
        if self.num_runs < 1:
            raise ValueError(f"KV pool of {size} slots is smaller than one run of {run_len}")
        self._kvcache = kvcache
        self._ring = _PinnedRing()
        self.is_not_in_free_group = True
        self.free_group: List[torch.Tensor] = []
        self.clear()

    # ---- bookkeeping ---------------------------------------------------------------------------
    def clear(self):
        self._free_runs = collections.deque(range(self.num_runs))
        self._used = np.zeros(self.num_runs, dtype=np.int32)
        self._req_runs: Dict[int, List[int]] = {}   # req_pool_idx -> runs in position order
        self._req_len: Dict[int, int] = {}          # req_pool_idx -> tokens placed
        self.is_not_in_free_group = True
        self.free_group = []

    def _base(self, run: int) -> int:
        return 1 + run * self.run_len  # slot 0 is the dummy slot padded CUDA-graph rows write to

    def available_size(self) -> int:
        """Free capacity in tokens (whole runs only — what a new request can count on)."""
        return len(self._free_runs) * self.run_len

    def get_kvcache(self):
        return self._kvcache

    def runs_needed(self, prefix_len: int, seq_len: int) -> int:
        """New runs a request needs to grow from prefix_len to seq_len tokens."""
        R = self.run_len
        return -(-seq_len // R) - (-(-prefix_len // R))

    def ceil_to_run(self, tokens: int) -> int:
        return -(-tokens // self.run_len) * self.run_len

    # ---- request-aware entry points (paged_allocator.py: alloc_extend / alloc_decode) ----------
    def alloc_extend(self, req_pool_indices: Sequence[int], prefix_lens: Sequence[int],
                     seq_lens: Sequence[int]) -> Optional[torch.Tensor]:
        """Slots for tokens prefix_len … seq_len-1 of every request, concatenated in batch order
        (int64, on the device).  None when the free runs do not suffice — nothing is taken then."""
        need = 0
        for idx, pre, seq in zip(req_pool_indices, prefix_lens, seq_lens):
            have = self._req_len.get(idx, 0) if pre > 0 else 0
            if pre != have:
                raise RuntimeError(f"request slot {idx}: prefix {pre} does not match the {have} tokens placed")
            need += self.runs_needed(pre, seq)
        if need > len(self._free_runs):
            return None
        R = self.run_len
        parts = []
        for idx, pre, seq in zip(req_pool_indices, prefix_lens, seq_lens):
            if pre == 0:
                if self._req_runs.get(idx):
                    raise RuntimeError(f"request slot {idx} still holds runs of a previous request")
                self._req_runs[idx] = []
            runs = self._req_runs[idx]
            for _ in range(self.runs_needed(pre, seq)):
                runs.append(self._free_runs.popleft())
            pos = np.arange(pre, seq, dtype=np.int64)
            bases = np.asarray(runs, dtype=np.int64) * R + 1
            parts.append(bases[pos // R] + pos % R)
            np.add.at(self._used, np.asarray(runs, dtype=np.int64)[pos // R], 1)
            self._req_len[idx] = seq
        flat = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
        return self._ring.stage(flat, self.device)

    def can_alloc_decode(self, seq_lens_before: Sequence[int]) -> bool:
        """One more token for every request of the batch (lens BEFORE the step)."""
        R = self.run_len
        return sum(1 for s in seq_lens_before if s % R == 0) <= len(self._free_runs)

    def alloc_decode(self, req_pool_indices: Sequence[int], seq_lens_before: Sequence[int]
                     ) -> Optional[torch.Tensor]:
        """The slot of position seq_len (the token this decode step writes) of every request."""
        if not self.can_alloc_decode(seq_lens_before):
            return None
        R = self.run_len
        out = np.empty(len(req_pool_indices), dtype=np.int64)
        for i, (idx, s) in enumerate(zip(req_pool_indices, seq_lens_before)):
            runs = self._req_runs[idx]
            if s != self._req_len[idx]:
                raise RuntimeError(f"request slot {idx}: length {s} does not match the {self._req_len[idx]} tokens placed")
            if s % R == 0:
                runs.append(self._free_runs.popleft())
            run = runs[s // R]
            out[i] = run * R + 1 + s % R
            self._used[run] += 1
            self._req_len[idx] = s + 1
        return self._ring.stage(out, self.device)

    def free_req(self, req_pool_idx: int) -> int:
        """Return every run of a request (finished or retracted).  → tokens released."""
        runs = self._req_runs.pop(req_pool_idx, None)
        n = self._req_len.pop(req_pool_idx, 0)
        if runs:
            self._used[runs] = 0
            self._free_runs.extend(runs)
        return n

    def req_slots(self, req_pool_idx: int) -> np.ndarray:
        """Slot ids of a request's tokens in position order (host-side view; tests)."""
        runs = np.asarray(self._req_runs.get(req_pool_idx, []), dtype=np.int64)
        pos = np.arange(self._req_len.get(req_pool_idx, 0), dtype=np.int64)
        return runs[pos // self.run_len] * self.run_len + 1 + pos % self.run_len if len(pos) else pos

    # ---- the reference's request-agnostic contract (memory_pool.py:153-184) --------------------
    def alloc(self, need_size: int) -> Optional[torch.Tensor]:
        """`need_size` slot ids (int64, device) from whole runs; the unused tail of the last run
        stays reserved until the slots next to it are freed."""
        R = self.run_len
        n_runs = -(-need_size // R)
        if n_runs > len(self._free_runs):
            return None
        runs = [self._free_runs.popleft() for _ in range(n_runs)]
        pos = np.arange(need_size, dtype=np.int64)
        flat = (np.asarray(runs, dtype=np.int64)[pos // R] * R + 1 + pos % R) if need_size else pos
        if need_size:
            np.add.at(self._used, np.asarray(runs, dtype=np.int64)[pos // R], 1)
        return self._ring.stage(flat, self.device)

    def free(self, free_index: torch.Tensor):
        """Slot ids back to the pool (synchronises: the ids are read on the host).  The engine
        itself releases through `free_req`."""
        if free_index.numel() == 0:
            return
        if not self.is_not_in_free_group:
            self.free_group.append(free_index)
            return
        ids = free_index.detach().to("cpu", torch.int64).numpy()
        ids = ids[ids > 0]
        runs, counts = np.unique((ids - 1) // self.run_len, return_counts=True)
        self._used[runs] -= counts.astype(np.int32)
        if (self._used[runs] < 0).any():
            raise RuntimeError("KV slots freed twice")
        self._free_runs.extend(int(r) for r in runs[self._used[runs] == 0])

    def free_group_begin(self):
        self.is_not_in_free_group = False
        self.free_group = []

    def free_group_end(self):
        self.is_not_in_free_group = True
        if self.free_group:
            self.free(torch.concat(self.free_group))
            self.free_group = []


class MHATokenToKVPool(KVCache):
    def __init__(self, size: int, page_size: int, dtype: torch.dtype, head_num: int, head_dim: int,
                 layer_num: int, device: str, enable_memory_saver: bool = False,
                 bypass_create_buffers: bool = False):
        if page_size != 1:
            # the reference's semi-PD prefill side raises for page_size > 1 too
            # (managers/schedule_batch.py:936-937)
            raise NotImplementedError("semi-PD path is token-granular: page_size must be 1")
        self.size = size
        self.page_size = page_size
        self.dtype = dtype
        self.store_dtype = dtype
        self.device = device
        self.head_num = head_num
        self.head_dim = head_dim
        self.layer_num = layer_num
        self._create_buffers()

    def _create_buffers(self):
        rows = self.size + self.page_size
        self._slab = torch.zeros((2, self.layer_num, rows, self.head_num, self.head_dim),
                                 dtype=self.store_dtype, device=self.device)
        self.k_buffer = [self._slab[0, i] for i in range(self.layer_num)]
        self.v_buffer = [self._slab[1, i] for i in range(self.layer_num)]

    def get_kv_size_bytes(self):
        per = self._slab[0].numel() * self._slab.element_size()
        return per, per

    def get_key_buffer(self, layer_id: int):
        return self.k_buffer[layer_id]

    def get_value_buffer(self, layer_id: int):
        return self.v_buffer[layer_id]

    def get_kv_buffer(self, layer_id: int):
        return self.k_buffer[layer_id], self.v_buffer[layer_id]

    def set_kv_buffer(self, layer, loc: torch.Tensor, cache_k: torch.Tensor, cache_v: torch.Tensor,
                      k_scale=None, v_scale=None):
        """k_buffer[layer][loc] = cache_k ; v_buffer[layer][loc] = cache_v — one libspd launch
        instead of the reference's two index_put_ kernels (memory_pool.py:316-346).  The Llama
        path normally never calls this: the write is fused into the RoPE kernel."""
        layer_id = layer.layer_id
        if cache_k.dtype != self.dtype:
            raise RuntimeError("KV cache dtype conversion is not on the semi-PD path")
        ops.kv_write(cache_k, cache_v, self.k_buffer[layer_id], self.v_buffer[layer_id], loc)


class MLATokenToKVPool(KVCache):
    """One latent buffer per layer, rows [kv_a_normed(kv_lora_rank) || k_pe_roped(qk_rope_head_dim)]
    (python/sglang/srt/mem_cache/memory_pool.py:379-452): `kv_buffer[l]` is [size + 1, 1, 576] for
    DeepSeek-V2-Lite — 1152 B per token per layer.  Keys are the whole row, values its first
    kv_lora_rank columns (`get_value_buffer` returns that view, :426-433)."""

    def __init__(self, size: int, page_size: int, dtype: torch.dtype, kv_lora_rank: int, qk_rope_head_dim: int,
                 layer_num: int, device: str, enable_memory_saver: bool = False,
                 bypass_create_buffers: bool = False):
        if page_size != 1:
            raise NotImplementedError("semi-PD path is token-granular: page_size must be 1")
        self.size = size
        self.page_size = page_size
        self.dtype = dtype
        self.store_dtype = dtype
        self.device = device
        self.kv_lora_rank = kv_lora_rank
        self.qk_rope_head_dim = qk_rope_head_dim
        self.layer_num = layer_num
        self.head_num = 1
        self.head_dim = kv_lora_rank + qk_rope_head_dim
        self._slab = torch.zeros((layer_num, size + page_size, 1, self.head_dim), dtype=dtype, device=device)
        self.kv_buffer = [self._slab[i] for i in range(layer_num)]

    def get_kv_size_bytes(self):
        return self._slab.numel() * self._slab.element_size()

    def get_key_buffer(self, layer_id: int):
        return self.kv_buffer[layer_id]

    def get_value_buffer(self, layer_id: int):
        return self.kv_buffer[layer_id][..., : self.kv_lora_rank]

    def get_kv_buffer(self, layer_id: int):
        return self.get_key_buffer(layer_id), self.get_value_buffer(layer_id)

    def set_kv_buffer(self, layer, loc: torch.Tensor, cache_k: torch.Tensor, cache_v: torch.Tensor = None):
        if cache_k.dtype != self.dtype:
            raise RuntimeError("KV cache dtype conversion is not on the semi-PD path")
        ops.kv_write_single(cache_k, self.kv_buffer[layer.layer_id], loc)
