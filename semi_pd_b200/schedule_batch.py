"""Request / batch bookkeeping the schedulers use to build `ModelWorkerBatch`es — the slice of
python/sglang/srt/managers/schedule_batch.py the semi-PD path exercises:
`Req` (:246-513), `ScheduleBatch.prepare_for_extend` (:790-985, incl. the
`pre_allocated_req_pool_indices` variant the prefill worker uses, :923-935),
`prepare_for_decode` (:1144-1207), `filter_batch` / `merge_batch` (:1209-1306),
`check_decode_mem` / `retract_decode` (:1022-1125) and `get_model_worker_batch` (:1308-1370).
Sampling is greedy-only; logprobs / grammar / multimodal are off this path
(managers/semi_pd_scheduler.py:207-209, semi_pd_decode_scheduler.py:354).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import torch

from . import ops
from .forward_batch_info import ForwardMode, ModelWorkerBatch
from .memory_pool import ReqToTokenPool, TokenToKVPoolAllocator

_bid = 0


@dataclasses.dataclass
class SamplingParams:
    max_new_tokens: int = 128
    ignore_eos: bool = True
    temperature: float = 0.0  # greedy


class Req:
    """The input and output status of a request."""

    def __init__(self, rid: str, origin_input_ids: List[int], sampling_params: SamplingParams,
                 eos_token_ids: Optional[set] = None, arrival_time: float = 0.0):
        self.rid = rid
        self.origin_input_ids = list(origin_input_ids)
        self.output_ids: List[int] = []
        self.fill_ids: List[int] = []
        self.sampling_params = sampling_params
        self.eos_token_ids = eos_token_ids or set()
        self.req_pool_idx: Optional[int] = None
        self.prefix_indices: List[int] = []   # length = tokens already in the KV pool
        self.extend_input_len = 0
        self.is_chunked = 0
        self.is_retracted = False
        self.finished_reason: Optional[str] = None
        # tokens whose computation has been launched (>= len(output_ids) under the overlap loop)
        self.launched_tokens = 0
        # timing (host clock, seconds)
        self.arrival_time = arrival_time
        self.first_token_time: Optional[float] = None
        self.finish_time: Optional[float] = None
        self.token_times: List[float] = []
        # output / timestamps produced before a retraction (this Req then continues the request
        # with prompt + those tokens as its input; stitched back together when it finishes)
        self.prior_output_ids: List[int] = []
        self.prior_token_times: List[float] = []
        self.retracted_from: Optional["Req"] = None

    @property
    def seqlen(self):
        return len(self.origin_input_ids) + len(self.output_ids)

    def finished(self) -> bool:
        return self.finished_reason is not None

    def init_next_round_input(self, prefix_len: Optional[int] = None):
        self.fill_ids = self.origin_input_ids + self.output_ids
        if prefix_len is not None:
            self.prefix_indices = [0] * prefix_len  # only the length matters (no radix cache)
        self.extend_input_len = len(self.fill_ids) - len(self.prefix_indices)

    def check_finished(self):
        if self.finished():
            return
        if len(self.output_ids) >= self.sampling_params.max_new_tokens:
            self.finished_reason = "length"
            return
        if not self.sampling_params.ignore_eos and self.output_ids and \
                self.output_ids[-1] in self.eos_token_ids:
            self.finished_reason = "stop"

    def reset_for_retract(self):
        self.prefix_indices = []
        self.extend_input_len = 0
        self.is_retracted = True
        self.req_pool_idx = None


class ScheduleBatch:
    """Store all information of a batch on the scheduler."""

    def __init__(self, reqs: List[Req], req_to_token_pool: ReqToTokenPool,
                 token_to_kv_pool_allocator: TokenToKVPoolAllocator, device):
        self.reqs = reqs
        self.req_to_token_pool = req_to_token_pool
        self.token_to_kv_pool_allocator = token_to_kv_pool_allocator
        self.device = device
        self.forward_mode: Optional[ForwardMode] = None
        self.input_ids = None
        self.req_pool_indices = None
        self.seq_lens = None
        self.out_cache_loc = None
        self.output_ids = None
        self.seq_lens_sum = 0
        self.seq_lens_host: List[int] = []   # host mirror of `seq_lens` (the run allocator's view)
        self.prefix_lens = None
        self.extend_lens = None
        self.extend_num_tokens = None
        self.decoding_reqs = None

    @classmethod
    def init_new(cls, reqs, req_to_token_pool, token_to_kv_pool_allocator, device=None):
        return cls(reqs, req_to_token_pool, token_to_kv_pool_allocator,
                   device or req_to_token_pool.device)

    def batch_size(self):
        return len(self.reqs)

    def is_empty(self):
        return len(self.reqs) == 0

    # --- allocation -----------------------------------------------------------------------------
    def alloc_req_slots(self, num_reqs: int):
        idx = self.req_to_token_pool.alloc(num_reqs)
        if idx is None:
            raise RuntimeError("alloc_req_slots runs out of memory. Please set a smaller number for "
                               "`--max-running-requests`.")
        return idx

    def _oom(self, phase: str, num_tokens: int):
        return RuntimeError(f"{phase} out of memory. Try to lower your batch size. "
                            f"Try to allocate {num_tokens} tokens; available "
                            f"{self.token_to_kv_pool_allocator.available_size()}.")

    def _h2d(self, values, dtype):
        return torch.tensor(values, dtype=dtype, pin_memory=True).to(self.device, non_blocking=True)

    # --- extend ---------------------------------------------------------------------------------
    def prepare_for_extend(self, pre_allocated_req_pool_indices: Optional[List[int]] = None):
        """Decode-side call (pre_allocated None): allocates request + KV slots and writes
        req_to_token.  Prefill-side call: adopts the request slots the decode side chose and reads
        out_cache_loc back from the shared req_to_token table (schedule_batch.py:923-935)."""
        self.forward_mode = ForwardMode.EXTEND
        bs = len(self.reqs)
        reqs = self.reqs
        input_ids = [r.fill_ids[len(r.prefix_indices):] for r in reqs]
        extend_num_tokens = sum(len(ids) for ids in input_ids)
        seq_lens = [len(r.fill_ids) for r in reqs]
        prefix_lens = [len(r.prefix_indices) for r in reqs]
        extend_lens = [r.extend_input_len for r in reqs]

        if pre_allocated_req_pool_indices is None:
            # a chunked / continued request keeps its slot
            need = [r for r in reqs if r.req_pool_idx is None]
            new_idx = self.alloc_req_slots(len(need))
            for r, i in zip(need, new_idx):
                r.req_pool_idx = i
            req_pool_indices = [r.req_pool_idx for r in reqs]
        else:
            req_pool_indices = list(pre_allocated_req_pool_indices)
            for r, i in zip(reqs, req_pool_indices):
                r.req_pool_idx = i

        if all(getattr(r, "input_ids_device", None) is not None for r in reqs):
            # prompt ids already resident in HBM (bench.py's device-resident arm): slice, no H2D
            parts = [r.input_ids_device[len(r.prefix_indices):len(r.fill_ids)] for r in reqs]
            self.input_ids = parts[0] if len(parts) == 1 else torch.cat(parts)
        else:
            self.input_ids = self._h2d([t for ids in input_ids for t in ids], torch.int64)
        self.req_pool_indices = self._h2d(req_pool_indices, torch.int64)
        self.seq_lens = self._h2d(seq_lens, torch.int64)
        prefix_lens_t = self._h2d(prefix_lens, torch.int64)
        extend_lens_t = self._h2d(extend_lens, torch.int64)

        if pre_allocated_req_pool_indices is None:
            # runs: positions [pre, seq) of every request continue its last run, then take whole
            # new runs (memory_pool.TokenToKVPoolAllocator.alloc_extend)
            self.out_cache_loc = self.token_to_kv_pool_allocator.alloc_extend(
                req_pool_indices, prefix_lens, seq_lens)
            if self.out_cache_loc is None:
                raise self._oom("Prefill", extend_num_tokens)
            ops.write_req_to_token_pool(self.req_to_token_pool.req_to_token, self.req_pool_indices,
                                        prefix_lens_t, self.seq_lens, extend_lens_t,
                                        self.out_cache_loc)
        else:
            # gather out_cache_loc = req_to_token[req, pre:seq] on the device (no host sync)
            kv_indptr = self._h2d([0] + list(_cumsum(extend_lens)), torch.int32)
            loc32 = torch.empty((extend_num_tokens,), dtype=torch.int32, device=self.device)
            ops.create_kv_indices(self.req_to_token_pool.req_to_token, self.req_pool_indices,
                                  extend_lens_t.int(), kv_indptr, prefix_lens_t.int(), loc32)
            self.out_cache_loc = loc32.long()

        self.seq_lens_sum = sum(seq_lens)
        self.seq_lens_host = list(seq_lens)
        self.extend_num_tokens = extend_num_tokens
        self.prefix_lens = prefix_lens
        self.extend_lens = extend_lens

    # --- decode ---------------------------------------------------------------------------------
    def check_decode_mem(self, buf_multiplier: int = 1) -> bool:
        """Can every request of the batch get the slot of its next token?  Exact under the run
        allocator: only requests whose last run is full need a new one."""
        return self.token_to_kv_pool_allocator.can_alloc_decode(self.seq_lens_host)

    def retract_decode(self, retract_decode_steps: int = 20):
        """Retract requests with the fewest generated tokens / longest prompts first until
        `retract_decode_steps` slots per remaining request are free (schedule_batch.py:1034-1111)."""
        order = sorted(range(len(self.reqs)),
                       key=lambda i: (len(self.reqs[i].output_ids), -len(self.reqs[i].origin_input_ids)),
                       reverse=True)
        retracted = []
        alloc = self.token_to_kv_pool_allocator
        while (alloc.available_size() < alloc.ceil_to_run(len(order) * retract_decode_steps)
               and len(order) > 1):
            idx = order.pop()
            req = self.reqs[idx]
            retracted.append(req)
            alloc.free_req(req.req_pool_idx)
            self.req_to_token_pool.free(req.req_pool_idx)
            req.reset_for_retract()
        self.filter_batch(keep_indices=sorted(order))
        return retracted

    def prepare_for_decode(self):
        self.forward_mode = ForwardMode.DECODE
        bs = len(self.reqs)
        # the token sampled by the previous step is this step's input
        self.input_ids = self.output_ids
        self.output_ids = None
        self.out_cache_loc = self.token_to_kv_pool_allocator.alloc_decode(
            [r.req_pool_idx for r in self.reqs], self.seq_lens_host)
        if self.out_cache_loc is None:
            raise self._oom("Decode", bs)
        # req_to_token[req, seq_len] = slot ; seq_lens += 1   (schedule_batch.py:1186-1207)
        ops.write_req_to_token_pool(self.req_to_token_pool.req_to_token, self.req_pool_indices,
                                    self.seq_lens, self.seq_lens + 1,
                                    torch.ones_like(self.seq_lens), self.out_cache_loc)
        self.seq_lens = self.seq_lens + 1
        self.seq_lens_host = [s + 1 for s in self.seq_lens_host]
        self.seq_lens_sum += bs

    def filter_batch(self, keep_indices: Optional[List[int]] = None):
        if keep_indices is None:
            keep_indices = [i for i, r in enumerate(self.reqs) if not r.finished()]
        if len(keep_indices) == len(self.reqs):
            return
        if not keep_indices:
            self.reqs = []
            self.req_pool_indices = self.seq_lens = self.output_ids = None
            self.seq_lens_sum = 0
            self.seq_lens_host = []
            return
        self.reqs = [self.reqs[i] for i in keep_indices]
        self.seq_lens_host = [self.seq_lens_host[i] for i in keep_indices]
        idx = self._h2d(keep_indices, torch.int64)
        self.req_pool_indices = self.req_pool_indices[idx]
        self.seq_lens = self.seq_lens[idx]
        if self.output_ids is not None:
            self.output_ids = self.output_ids[idx]
        self.seq_lens_sum = sum(self.seq_lens_host)

    def merge_batch(self, other: "ScheduleBatch"):
        if self.is_empty():
            self.reqs = list(other.reqs)
            self.req_pool_indices, self.seq_lens = other.req_pool_indices, other.seq_lens
            self.output_ids = other.output_ids
            self.seq_lens_sum = other.seq_lens_sum
            self.seq_lens_host = list(other.seq_lens_host)
            return
        self.req_pool_indices = torch.cat([self.req_pool_indices, other.req_pool_indices])
        self.seq_lens = torch.cat([self.seq_lens, other.seq_lens])
        if self.output_ids is not None and other.output_ids is not None:
            self.output_ids = torch.cat([self.output_ids, other.output_ids])
        self.seq_lens_sum += other.seq_lens_sum
        self.seq_lens_host = self.seq_lens_host + list(other.seq_lens_host)
        self.reqs.extend(other.reqs)

    def get_model_worker_batch(self) -> ModelWorkerBatch:
        global _bid
        _bid += 1
        extend = self.forward_mode.is_extend()
        return ModelWorkerBatch(
            bid=_bid, forward_mode=self.forward_mode, input_ids=self.input_ids,
            req_pool_indices=self.req_pool_indices, seq_lens=self.seq_lens,
            out_cache_loc=self.out_cache_loc, seq_lens_sum=self.seq_lens_sum,
            extend_num_tokens=self.extend_num_tokens if extend else None,
            extend_seq_lens=self.extend_lens if extend else None,
            extend_prefix_lens=self.prefix_lens if extend else None,
            extend_logprob_start_lens=[0] * len(self.reqs) if extend else None,
            gather_logits=getattr(self, "gather_logits", True))


def _cumsum(xs):
    s = 0
    for x in xs:
        s += x
        yield s
