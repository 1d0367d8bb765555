"""KV-cache storage and slot allocators — same classes and methods as the reference's
python/sglang/srt/mem_cache/memory_pool.py (`ReqToTokenPool` :46-96, `TokenToKVPoolAllocator`
:124-184, `MHATokenToKVPool` :187-346), with one structural change that is the point of this
build: there is ONE pool per GPU that the prefill and the decode worker address directly (they
are threads of one process on disjoint SM partitions), so `bypass_create_buffers` / the
cudaIpcMemHandle export-import of every buffer (model_runner.py:346-624, semi-pd-ipc/ipc.cpp)
has nothing left to do.

HBM layout: per layer, `k_buffer[l]` and `v_buffer[l]` are [size + page_size, H_kv, D] 16-bit,
token-granular (page_size = 1); all layers live in one contiguous allocation
[2, L, size + 1, H_kv, D] so a 180 GB part is carved once.  Slot 0 is the dummy slot padded
CUDA-graph rows write to (memory_pool.py:179-182,233-250).
"""
from __future__ import annotations

import abc
from typing import List, Tuple, Union

import torch

from . import ops

GB = 1024 ** 3


class ReqToTokenPool:
    """A memory pool that maps a request to its token locations."""

    def __init__(self, size: int, max_context_len: int, device: str,
                 enable_memory_saver: bool = False, bypass_create_buffers: bool = False):
        self.size = size
        self.max_context_len = max_context_len
        self.device = device
        # one shared tensor for both workers; nothing to bypass in a single process
        self.req_to_token = torch.zeros((size, max_context_len), dtype=torch.int32, device=device)
        self.free_slots = list(range(size))

    def write(self, indices, values):
        self.req_to_token[indices] = values

    def available_size(self):
        return len(self.free_slots)

    def alloc(self, need_size: int) -> List[int]:
        if need_size > len(self.free_slots):
            return None
        select_index = self.free_slots[:need_size]
        self.free_slots = self.free_slots[need_size:]
        return select_index

    def free(self, free_index: Union[int, List[int]]):
        if isinstance(free_index, int):
            self.free_slots.append(free_index)
        else:
            self.free_slots.extend(free_index)

    def clear(self):
        self.free_slots = list(range(self.size))


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


class TokenToKVPoolAllocator:
    """An allocator managing the indices to kv cache data (free list = int64 tensor)."""

    def __init__(self, size: int, dtype: torch.dtype, device: str, kvcache: KVCache):
        self.size = size
        self.dtype = dtype
        self.device = device
        self.page_size = 1
        self.free_slots = None
        self.is_not_in_free_group = True
        self.free_group = []
        self.clear()
        self._kvcache = kvcache

    def available_size(self):
        return len(self.free_slots)

    def get_kvcache(self):
        return self._kvcache

    def alloc(self, need_size: int):
        if need_size > len(self.free_slots):
            return None
        select_index = self.free_slots[:need_size]
        self.free_slots = self.free_slots[need_size:]
        return select_index

    def free(self, free_index: torch.Tensor):
        if free_index.numel() == 0:
            return
        if self.is_not_in_free_group:
            self.free_slots = torch.concat((self.free_slots, free_index))
        else:
            self.free_group.append(free_index)

    def free_group_begin(self):
        self.is_not_in_free_group = False
        self.free_group = []

    def free_group_end(self):
        self.is_not_in_free_group = True
        if self.free_group:
            self.free(torch.concat(self.free_group))

    def clear(self):
        # The padded slot 0 is used for writing dummy outputs from padded tokens.
        self.free_slots = torch.arange(1, self.size + 1, dtype=torch.int64, device=self.device)
        self.is_not_in_free_group = True
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
