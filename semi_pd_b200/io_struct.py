"""Messages exchanged between the tokenizer side and the two schedulers — the semi-PD subset of
python/sglang/srt/managers/io_struct.py (`TokenizedGenerateReqInput` :269-318, semi-PD messages
:734-755).  In the reference they are pickled over ZMQ ipc:// sockets between processes; here
they travel through in-process queues, and the two batch messages additionally carry a CUDA
event so the hand-off is ordered ON THE DEVICE (the reference relies on the ZMQ round trip being
slower than the req_to_token kernel).
"""
from __future__ import annotations

import dataclasses
from typing import Any, List, Optional

from .schedule_batch import SamplingParams


@dataclasses.dataclass
class TokenizedGenerateReqInput:
    rid: str
    input_text: str
    input_ids: List[int]
    sampling_params: SamplingParams
    stream: bool = True
    is_retracted: bool = False
    arrival_time: float = 0.0
    input_ids_device: Any = None  # optional: the same ids already resident in HBM (int64 tensor)


@dataclasses.dataclass
class GetNextPrefillBatchInput:
    rids: List[str]


@dataclasses.dataclass
class GetNextPrefillBatchOutput:
    rids: List[str]
    chunked_rid: Optional[str]
    req_pool_indices: List[int]
    prefix_lens: List[int]
    extend_input_lens: List[int]
    # device-side ordering: recorded on the decode stream after req_to_token was written
    ready_event: Any = None


@dataclasses.dataclass
class BatchProcessPrefillResultReq:
    next_token_ids: List[int]
    next_token_logits: Optional[Any] = None
    # recorded on the prefill stream after the forward (KV rows of the batch are in the pool)
    done_event: Any = None
    # host timestamps of the prefill worker (batch fetched, forward launched, ids read back):
    # TTFT breakdown of scripts/debug_ttft_breakdown.py
    timing: Optional[tuple] = None


@dataclasses.dataclass
class BatchTokenIDOut:
    """What the detokenizer would receive (io_struct.py:430-470): one new token per request."""
    rids: List[str]
    output_ids: List[int]
    finished: List[bool]
    timestamp: float
