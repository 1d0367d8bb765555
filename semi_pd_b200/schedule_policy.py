"""Prefill admission — the part of python/sglang/srt/managers/schedule_policy.py:272-510
(`PrefillAdder`) the semi-PD decode scheduler runs, for the no-radix-cache case semi-PD forces
(server_args.py:325-331).  Budgets: `rem_total_tokens` (free KV slots minus the expected growth
of the running decode batch, weighted by `new_token_ratio`, minus what this round admitted),
`rem_input_tokens` (= max_prefill_tokens) and `rem_chunk_tokens` (= chunked_prefill_size); a
request that does not fit the chunk budget is truncated and becomes THE chunked request.
"""
from __future__ import annotations

from enum import Enum, auto
from typing import List, Optional

from .schedule_batch import Req

CLIP_MAX_NEW_TOKENS_ESTIMATION = 4096


class AddReqResult(Enum):
    CONTINUE = auto()   # Continue to add requests
    NO_TOKEN = auto()   # No token left
    OTHER = auto()      # Other reasons to stop adding requests


class PrefillAdder:
    def __init__(self, available_size_fn, running_reqs: List[Req], new_token_ratio: float,
                 rem_input_tokens: int, rem_chunk_tokens: Optional[int], run_len: int = 1):
        # KV space is handed out in runs of `run_len` slots per request (memory_pool.py), so every
        # per-request reservation is rounded up to whole runs (the reference's later paged
        # allocator does the same with its page size)
        self.run_len = max(int(run_len), 1)
        self.available_size_fn = available_size_fn
        self.new_token_ratio = new_token_ratio
        self.rem_input_tokens = rem_input_tokens
        self.rem_chunk_tokens = rem_chunk_tokens
        self.rem_total_token_offset = 0
        self.cur_rem_token_offset = 0
        self.can_run_list: List[Req] = []
        self.new_chunked_req: Optional[Req] = None
        self.log_input_tokens = 0
        self.rem_total_token_offset += sum(
            min(r.sampling_params.max_new_tokens - len(r.output_ids), CLIP_MAX_NEW_TOKENS_ESTIMATION)
            * new_token_ratio for r in running_reqs)

    @property
    def rem_total_tokens(self):
        return self.available_size_fn() - self.rem_total_token_offset

    @property
    def cur_rem_tokens(self):
        return self.available_size_fn() - self.cur_rem_token_offset

    def budget_state(self) -> AddReqResult:
        if self.rem_total_tokens <= 0 or self.cur_rem_tokens <= 0:
            return AddReqResult.NO_TOKEN
        if self.rem_input_tokens <= 0 or (self.rem_chunk_tokens is not None and self.rem_chunk_tokens <= 0):
            return AddReqResult.OTHER
        return AddReqResult.CONTINUE

    def _ceil(self, tokens: int) -> int:
        return -(-tokens // self.run_len) * self.run_len

    def _prefill_one_req(self, prefix_len: int, extend_input_len: int, max_new_tokens: int):
        self.rem_total_token_offset += self._ceil(extend_input_len + max_new_tokens)
        self.cur_rem_token_offset += self._ceil(extend_input_len)
        self.rem_input_tokens -= extend_input_len
        if self.rem_chunk_tokens is not None:
            self.rem_chunk_tokens -= extend_input_len
        self.log_input_tokens += extend_input_len

    def add_chunked_req(self, req: Req) -> Optional[Req]:
        """Continue the request that was truncated last round (schedule_policy.py:352-368)."""
        truncated = req.extend_input_len > self.rem_chunk_tokens
        req.extend_input_len = min(req.extend_input_len, self.rem_chunk_tokens)
        req.fill_ids = req.fill_ids[: len(req.prefix_indices) + req.extend_input_len]
        self.can_run_list.append(req)
        self._prefill_one_req(len(req.prefix_indices), req.extend_input_len,
                              0 if truncated else min(req.sampling_params.max_new_tokens,
                                                      CLIP_MAX_NEW_TOKENS_ESTIMATION))
        return req if truncated else None

    def add_one_req(self, req: Req, has_chunked_req: bool) -> AddReqResult:
        total_tokens = self._ceil(req.extend_input_len + min(req.sampling_params.max_new_tokens,
                                                             CLIP_MAX_NEW_TOKENS_ESTIMATION))
        input_tokens = req.extend_input_len
        prefix_len = len(req.prefix_indices)
        if total_tokens >= self.rem_total_tokens:
            return AddReqResult.NO_TOKEN
        if input_tokens > self.rem_input_tokens and len(self.can_run_list) != 0:
            return AddReqResult.OTHER
        if self.rem_chunk_tokens is None or input_tokens <= self.rem_chunk_tokens:
            self.can_run_list.append(req)
            self._prefill_one_req(prefix_len, input_tokens,
                                  min(req.sampling_params.max_new_tokens, CLIP_MAX_NEW_TOKENS_ESTIMATION))
        else:
            if self.rem_chunk_tokens == 0 or has_chunked_req:
                return AddReqResult.OTHER
            trunc_len = self.rem_chunk_tokens
            req.extend_input_len = trunc_len
            req.fill_ids = req.fill_ids[: len(req.prefix_indices) + trunc_len]
            self.can_run_list.append(req)
            self.new_chunked_req = req
            self._prefill_one_req(prefix_len, trunc_len, 0)
        return self.budget_state()
