"""Semi-PD schedulers and launcher — the caller of the hot path, kept protocol-compatible with
the reference (python/sglang/srt/managers/semi_pd_scheduler.py:326-432,
semi_pd_prefill_scheduler.py:74-173, semi_pd_decode_scheduler.py:90-377,
entrypoints/engine.py:540-728) but re-shaped for one process per GPU:

  reference                                   here
  ------------------------------------------  ---------------------------------------------------
  2 OS processes per GPU (P, D)               2 threads, one CUDA context
  MPS caps via CUDA_MPS_ACTIVE_THREAD_PERC.   2 CUDA green contexts with DISJOINT SM sets
  cudaIpcMemHandle export of weights/KV       one weight set + one KV pool, addressed directly
  ZMQ pickles, P blocks on D's loop           in-process queues; D pre-admits prefill batches so P
                                              never waits a decode step (removes the TTFT floor)
  no GPU-side ordering of req_to_token        CUDA events ride on the two batch messages

The decode scheduler remains the single memory authority (admission, request + KV slot
allocation, retraction); the prefill scheduler only runs forwards into slots D chose and returns
the first token ids.
"""
from __future__ import annotations

import collections
import itertools
import logging
import queue
import threading
import time
from typing import Deque, Dict, List, Optional

import torch

from . import distributed, ops
from .configs import (DECODE_ENGINE_SM_PERCENTILE, PREFILL_ENGINE_SM_PERCENTILE, InstanceRole,
                      ModelConfig, ServerArgs, get_model_config)
from .io_struct import (BatchProcessPrefillResultReq, GetNextPrefillBatchInput,
                        GetNextPrefillBatchOutput, TokenizedGenerateReqInput)
from .schedule_batch import Req, SamplingParams, ScheduleBatch
from .schedule_policy import AddReqResult, PrefillAdder
from .tp_worker import TpModelWorker

logger = logging.getLogger(__name__)

# python/sglang/global_config.py:24-32
DEFAULT_INIT_NEW_TOKEN_RATIO = 0.7
DEFAULT_MIN_NEW_TOKEN_RATIO_FACTOR = 0.14
DEFAULT_NEW_TOKEN_RATIO_DECAY_STEPS = 600
RETRACT_DECODE_STEPS = 20


class SemiPDScheduler:
    """State shared by both roles (managers/semi_pd_scheduler.py:40-216)."""

    def __init__(self, server_args: ServerArgs, worker: TpModelWorker, stream: torch.cuda.Stream,
                 sm_budget: int, role: InstanceRole):
        self.server_args = server_args
        self.tp_worker = worker
        self.stream = stream
        self.sm_budget = sm_budget
        self.role = role
        self.device = worker.device
        self.req_to_token_pool, self.token_to_kv_pool_allocator = worker.get_memory_pool()
        self.waiting_queue: List[Req] = []
        self.inbox: "queue.Queue" = queue.Queue()
        self.stop_flag = threading.Event()
        self.error: Optional[BaseException] = None
        self.forward_ct = 0
        self.tp_rank = worker.tp_rank
        self.tp_size = server_args.tp_size
        # bench.py's device-resident arm: prompt ids are sliced from HBM and sampled ids never
        # leave the device (completion is by launch count); off for normal serving
        self.device_resident = False

    def validate_request(self, input_ids: List[int], max_new_tokens: int, strict: bool = True):
        """managers/utils.py:10-44 (`validate_input_length`) + scheduler.py:791-798: an input that
        cannot fit the context / the KV pool is rejected (the reference aborts it with
        BAD_REQUEST unless --allow-auto-truncate), and max_new_tokens is clipped so that
        prompt + output stay inside `max_req_len` — which also bounds the request's
        req_to_token row and RoPE positions.  → (input_ids, max_new_tokens)."""
        w = self.tp_worker
        n = len(input_ids)
        if n == 0:
            if strict:
                raise ValueError("empty input_ids")
            input_ids, n = [0], 1
        if n >= w.max_req_input_len:
            if strict and not getattr(self.server_args, "allow_auto_truncate", False):
                raise ValueError(f"Input length ({n} tokens) exceeds the maximum allowed length "
                                 f"({w.max_req_input_len} tokens). Use a shorter input or enable "
                                 f"--allow-auto-truncate.")
            input_ids = input_ids[: w.max_req_input_len]
            n = len(input_ids)
        max_new_tokens = min(max_new_tokens if max_new_tokens is not None else 1 << 30,
                             w.max_req_len - n - 1)
        return input_ids, max(int(max_new_tokens), 1)

    def handle_generate_request(self, recv_req: TokenizedGenerateReqInput):
        sp = recv_req.sampling_params
        if not recv_req.is_retracted:
            # the engine front (SemiPDEngine.submit) already rejected what must be rejected; a
            # scheduler thread never raises on a request, it clips
            recv_req.input_ids, sp.max_new_tokens = self.validate_request(
                recv_req.input_ids, sp.max_new_tokens, strict=False)
        req = Req(recv_req.rid, recv_req.input_ids, sp, arrival_time=recv_req.arrival_time,
                  eos_token_ids=set(getattr(self.tp_worker.model_config, "eos_token_ids", None) or ()))
        req.input_ids_device = recv_req.input_ids_device
        if recv_req.is_retracted:
            self.waiting_queue.insert(0, req)  # semi_pd_scheduler.py:52-56
        else:
            self.waiting_queue.append(req)
        return req

    def _thread_setup(self):
        torch.cuda.set_device(self.device)
        ops.set_sm_budget(self.sm_budget)
        distributed.set_current_comm(self.tp_worker.model_runner.tp_comm)

    def run(self):
        try:
            self._thread_setup()
            with torch.cuda.stream(self.stream):
                self.event_loop()
        except BaseException as e:  # noqa: BLE001  (reference: SIGQUIT to the parent, engine.py:407-413)
            logger.exception("scheduler %s died", self.role)
            self.error = e
            self.stop_flag.set()


class SemiPDPrefillScheduler(SemiPDScheduler):
    """managers/semi_pd_prefill_scheduler.py:74-173."""

    def __init__(self, server_args, worker, stream, sm_budget, d_inbox: "queue.Queue"):
        super().__init__(server_args, worker, stream, sm_budget, InstanceRole.PREFILL)
        self.send_to_d_instance = d_inbox
        self.bridge_socket: "queue.Queue" = queue.Queue()  # D -> P: GetNextPrefillBatchOutput
        self.chunked_rid: Optional[str] = None
        self.prefill_tokens = 0
        self.prefill_busy_s = 0.0
        # sized by the admission limit the decode side actually enforces (a prefill batch may hold
        # as many short prompts as max_running_requests allows); grown on demand beyond that
        n_ids = max(int(getattr(worker, "max_running_requests", 0) or 0), 256)
        self._pinned_ids = [torch.empty(n_ids, dtype=torch.int32).pin_memory()
                            for _ in range(self.MAX_LAUNCHED_BATCHES + 2)]
        # TP ranks > 0 hand their (identical) result to the local decode scheduler directly; only
        # rank 0's report travels through D's inbox and is broadcast (scheduler.py:599-659)
        self.local_result_queue: Optional["queue.Queue"] = None

    def recv_requests(self):
        while True:
            try:
                msg = self.inbox.get_nowait()
            except queue.Empty:
                return
            if isinstance(msg, TokenizedGenerateReqInput):
                self.handle_generate_request(msg)

    def to_extend_batch(self, resp: GetNextPrefillBatchOutput) -> ScheduleBatch:
        by_rid = {r.rid: r for r in self.waiting_queue}
        can_run_list = [by_rid[rid] for rid in resp.rids]
        # requests of this batch are fully prefilled afterwards — except the one D marked as still
        # chunked, which stays queued for its next chunk (prefill_scheduler.py:79-95)
        done = set(resp.rids) - ({resp.chunked_rid} if resp.chunked_rid else set())
        self.waiting_queue = [r for r in self.waiting_queue if r.rid not in done]
        self.chunked_rid = resp.chunked_rid
        for i, r in enumerate(can_run_list):
            r.extend_input_len = resp.extend_input_lens[i]
            pre_len = resp.prefix_lens[i]
            r.prefix_indices = [0] * pre_len
            r.fill_ids = r.origin_input_ids[: pre_len + r.extend_input_len]
        batch = ScheduleBatch.init_new(can_run_list, self.req_to_token_pool,
                                       self.token_to_kv_pool_allocator, self.device)
        batch.prepare_for_extend(pre_allocated_req_pool_indices=resp.req_pool_indices)
        return batch

    # launched prefill batches whose result has not been handed to D yet (D pre-admits as many)
    MAX_LAUNCHED_BATCHES = 2

    def event_loop(self):
        """The reference's P runs the non-overlapped loop (semi_pd_scheduler.py:422-427): launch a
        batch, wait for its sampled ids, report, fetch the next batch — the partition idles for the
        host work in between (measured here: ~14 % of a saturated window).  Same protocol, but the
        next batch D has already admitted is launched BEFORE the previous one's ids are read back,
        so the prefill stream never runs dry; results are still reported in order, each as soon as
        its event has fired."""
        pending: Deque[tuple] = collections.deque()
        while not self.stop_flag.is_set():
            launched = False
            if len(pending) < self.MAX_LAUNCHED_BATCHES:
                try:
                    resp = self.bridge_socket.get(timeout=0.0005 if pending else 0.002)
                except queue.Empty:
                    resp = None
                if resp is not None:
                    pending.append(self._launch_batch(resp))
                    launched = True
            while pending and pending[0][-1].query():
                self.process_batch_result_prefill(*pending.popleft())
            if pending and not launched and len(pending) >= self.MAX_LAUNCHED_BATCHES:
                # nothing more may be launched until the oldest batch has been reported
                self.process_batch_result_prefill(*pending.popleft())
        while pending:
            self.process_batch_result_prefill(*pending.popleft())

    def _launch_batch(self, resp: GetNextPrefillBatchOutput):
        self.recv_requests()
        t0 = time.perf_counter()
        if resp.ready_event is not None:
            self.stream.wait_event(resp.ready_event)
        batch = self.to_extend_batch(resp)
        batch.gather_logits = self.tp_size == 1  # only the greedy ids are consumed
        mwb = batch.get_model_worker_batch()
        _, next_token_ids = self.tp_worker.forward_batch_generation(mwb)
        self.forward_ct += 1
        self.prefill_tokens += mwb.extend_num_tokens
        n = len(batch.reqs)
        if self.device_resident:
            ids = next_token_ids.to(torch.int64)  # stays in HBM; D consumes it after done_event
        else:
            # async D2H into pinned memory; read after the event (the only host wait of the worker)
            k = self.forward_ct % len(self._pinned_ids)
            if n > self._pinned_ids[k].numel():
                self._pinned_ids[k] = torch.empty(n, dtype=torch.int32).pin_memory()
            ids = self._pinned_ids[k][:n]
            ids.copy_(next_token_ids[:n], non_blocking=True)
        done = torch.cuda.Event()
        done.record(self.stream)
        batch.t_fetched, batch.t_launched = t0, time.perf_counter()
        return batch, ids, t0, done

    def process_batch_result_prefill(self, batch: ScheduleBatch, ids, t0: float, done):
        # In the device-resident arm there is no D2H copy, but D must not merge these requests
        # before their KV exists — otherwise the running decode batch would stall on the device
        # behind this prefill — so the report waits for the event in both arms.
        done.synchronize()
        if not self.device_resident:
            ids = ids.tolist()
        self.prefill_busy_s += time.perf_counter() - t0
        msg = BatchProcessPrefillResultReq(next_token_ids=ids, done_event=done,
                                           timing=(batch.t_fetched, batch.t_launched, time.perf_counter()))
        if self.tp_size > 1 and self.tp_rank != 0:
            self.local_result_queue.put(msg)
        else:
            self.send_to_d_instance.put(msg)


class SemiPDDecodeScheduler(SemiPDScheduler):
    """managers/semi_pd_decode_scheduler.py:90-377 + the overlap loop of scheduler.py:559-597."""

    MAX_INFLIGHT_PREFILL_BATCHES = 2

    def __init__(self, server_args, worker, stream, sm_budget):
        super().__init__(server_args, worker, stream, sm_budget, InstanceRole.DECODE)
        self.send_to_p_instance: Optional["queue.Queue"] = None  # P's inbox (retracted requests)
        self.bridge_socket: Optional["queue.Queue"] = None       # D -> P batch descriptions
        self.running_batch = ScheduleBatch.init_new([], self.req_to_token_pool,
                                                    self.token_to_kv_pool_allocator, self.device)
        self.scheduled_prefill_batches: Deque[ScheduleBatch] = collections.deque()
        self.chunked_req: Optional[Req] = None
        self.max_running_requests = worker.max_running_requests
        self.max_prefill_tokens = server_args.max_prefill_tokens
        self.chunked_prefill_size = server_args.chunked_prefill_size
        self.init_new_token_ratio = min(DEFAULT_INIT_NEW_TOKEN_RATIO * server_args.schedule_conservativeness, 1.0)
        self.min_new_token_ratio = min(self.init_new_token_ratio * DEFAULT_MIN_NEW_TOKEN_RATIO_FACTOR, 1.0)
        self.new_token_ratio_decay = ((self.init_new_token_ratio - self.min_new_token_ratio)
                                      / DEFAULT_NEW_TOKEN_RATIO_DECAY_STEPS)
        self.new_token_ratio = self.init_new_token_ratio
        self.finished: "queue.Queue[Req]" = queue.Queue()
        self.num_finished = 0
        self.num_retracted = 0
        self.decode_tokens = 0
        self.decode_steps = 0
        self.result_queue: Deque = collections.deque()
        self._pinned_ring = [torch.empty((4096,), dtype=torch.int32, pin_memory=True) for _ in range(4)]
        self._ring_i = 0
        self.token_callback = None  # optional: fn(req, token_id, finished) — the detokenizer hook
        # TP > 1: rank 0 broadcasts the messages it drained every iteration so that all ranks'
        # decode schedulers take identical decisions in lockstep (scheduler.py:599-659)
        self.cpu_group = None
        self.local_prefill_results: "queue.Queue" = queue.Queue()
        self._stop_broadcast = False

    # ---- inbound -------------------------------------------------------------------------------
    def _drain_inbox(self) -> list:
        msgs = []
        while True:
            try:
                msgs.append(self.inbox.get_nowait())
            except queue.Empty:
                return msgs

    def _dispatch(self, msg):
        if isinstance(msg, TokenizedGenerateReqInput):
            self.handle_generate_request(msg)
        elif isinstance(msg, BatchProcessPrefillResultReq):
            self.process_prefill_result(msg)
        elif isinstance(msg, GetNextPrefillBatchInput):
            self.get_next_prefill_batch(msg)

    def recv_requests(self):
        if self.tp_size == 1:
            for msg in self._drain_inbox():
                self._dispatch(msg)
            return
        # Lockstep under TP: what rank 0 received this iteration is what every rank processes
        # this iteration (the reference's broadcast_pyobj over the gloo group).  Device handles
        # (CUDA events) cannot travel; every rank substitutes its own prefill worker's report.
        import torch.distributed as dist
        if self.tp_rank == 0:
            msgs = self._drain_inbox()
            wire = []
            for m in msgs:
                if isinstance(m, TokenizedGenerateReqInput):
                    wire.append(("gen", m.rid, m.input_ids, m.sampling_params.max_new_tokens,
                                 m.sampling_params.ignore_eos, m.arrival_time))
                elif isinstance(m, BatchProcessPrefillResultReq):
                    wire.append(("prefill_result",))
            stopping = self.stop_flag.is_set()  # read ONCE: what is broadcast is what is acted on
            if stopping:
                wire.append(("stop",))
            box = [wire]
            dist.broadcast_object_list(box, src=0, group=self.cpu_group)
            if stopping:
                self._stop_broadcast = True
            for m in msgs:
                self._dispatch(m)
            return
        box = [None]
        dist.broadcast_object_list(box, src=0, group=self.cpu_group)
        for w in box[0]:
            if w[0] == "gen":
                msg = TokenizedGenerateReqInput(
                    rid=w[1], input_text="", input_ids=list(w[2]),
                    sampling_params=SamplingParams(max_new_tokens=w[3], ignore_eos=w[4]), arrival_time=w[5])
                self.send_to_p_instance.put(msg)  # P knows the rid before D can schedule it
                self.handle_generate_request(msg)
            elif w[0] == "prefill_result":
                # this rank's prefill worker ran the same batch; wait for ITS report (its ids are
                # identical, its done_event orders this rank's KV writes before the decode steps)
                while True:
                    try:
                        local = self.local_prefill_results.get(timeout=0.5)
                        break
                    except queue.Empty:
                        if self.stop_flag.is_set():
                            return
                self.process_prefill_result(local)
            elif w[0] == "stop":
                self._stop_broadcast = True
                self.stop_flag.set()

    # ---- prefill admission (D decides, P computes) ------------------------------------------------
    def propose_prefill_candidates(self) -> List[str]:
        """The rule the reference's P uses to pick what it asks for (prefill_scheduler.py:123-129)."""
        n, cands = 0, []
        for r in self.waiting_queue:
            if n > self.chunked_prefill_size:
                break
            n += len(r.origin_input_ids)
            cands.append(r.rid)
        return cands

    def maybe_schedule_prefill(self):
        while (len(self.scheduled_prefill_batches) < self.MAX_INFLIGHT_PREFILL_BATCHES
               and (self.waiting_queue or self.chunked_req is not None)):
            before = len(self.scheduled_prefill_batches)
            self.get_next_prefill_batch(GetNextPrefillBatchInput(rids=self.propose_prefill_candidates()))
            if len(self.scheduled_prefill_batches) == before:
                break

    def get_new_batch_prefill(self, rids: List[str]) -> Optional[ScheduleBatch]:
        if (getattr(self.running_batch, "batch_is_full", False) or not self.waiting_queue) \
                and self.chunked_req is None:
            return None
        inflight = sum(len(b.reqs) for b in self.scheduled_prefill_batches)
        running_bs = len(self.running_batch.reqs) + inflight
        if running_bs >= self.max_running_requests and self.chunked_req is None:
            return None
        # slots promised to in-flight prefill batches' future decode growth count as running
        running_reqs = list(self.running_batch.reqs) + [r for b in self.scheduled_prefill_batches for r in b.reqs]
        adder = PrefillAdder(self.token_to_kv_pool_allocator.available_size, running_reqs,
                             self.new_token_ratio, self.max_prefill_tokens, self.chunked_prefill_size,
                             run_len=self.token_to_kv_pool_allocator.run_len)
        if self.chunked_req is not None:
            self.chunked_req.init_next_round_input(prefix_len=len(self.chunked_req.fill_ids))
            self.chunked_req = adder.add_chunked_req(self.chunked_req)
        rid_set = set(rids)
        for req in self.waiting_queue:
            if req.rid not in rid_set:
                continue
            if running_bs + len(adder.can_run_list) >= self.max_running_requests:
                break
            req.init_next_round_input(prefix_len=0)
            res = adder.add_one_req(req, self.chunked_req is not None)
            if res != AddReqResult.CONTINUE:
                break
        can_run_list = adder.can_run_list
        if not can_run_list:
            return None
        run = set(id(r) for r in can_run_list)
        self.waiting_queue = [r for r in self.waiting_queue if id(r) not in run]
        if adder.new_chunked_req is not None:
            assert self.chunked_req is None
            self.chunked_req = adder.new_chunked_req
        if self.chunked_req:
            self.chunked_req.is_chunked += 1
        new_batch = ScheduleBatch.init_new(can_run_list, self.req_to_token_pool,
                                           self.token_to_kv_pool_allocator, self.device)
        new_batch.prepare_for_extend()
        # the request (if any) whose prefill is NOT complete after this batch
        new_batch.chunked_req = self.chunked_req
        new_batch.t_admitted = time.perf_counter()
        self.scheduled_prefill_batches.append(new_batch)
        return new_batch

    def get_next_prefill_batch(self, recv_req: GetNextPrefillBatchInput):
        batch = self.get_new_batch_prefill(recv_req.rids)
        if batch is None:
            return
        ev = torch.cuda.Event()
        ev.record(self.stream)  # req_to_token rows of this batch are written before this point
        self.bridge_socket.put(GetNextPrefillBatchOutput(
            rids=[r.rid for r in batch.reqs],
            chunked_rid=(self.chunked_req.rid if self.chunked_req else None),
            req_pool_indices=[r.req_pool_idx for r in batch.reqs],
            prefix_lens=[len(r.prefix_indices) for r in batch.reqs],
            extend_input_lens=[r.extend_input_len for r in batch.reqs],
            ready_event=ev))

    def process_prefill_result(self, recv_req: BatchProcessPrefillResultReq):
        batch = self.scheduled_prefill_batches.popleft()  # FIFO: P executes in the order D issued
        assert len(batch.reqs) == len(recv_req.next_token_ids)
        if recv_req.done_event is not None:
            self.stream.wait_event(recv_req.done_event)  # KV rows of the batch precede our decode
        now = time.perf_counter()
        keep = []
        on_device = isinstance(recv_req.next_token_ids, torch.Tensor)
        host_ids = [-1] * len(batch.reqs) if on_device else recv_req.next_token_ids
        for i, (req, tok) in enumerate(zip(batch.reqs, host_ids)):
            if req is batch.chunked_req:
                continue  # intermediate chunk: the sampled token is discarded
            req.output_ids.append(int(tok))
            req.launched_tokens = 1
            if req.first_token_time is None:  # a retracted request keeps its original TTFT
                req.first_token_time = now
                req.ttft_stamps = (getattr(batch, "t_admitted", None),) + tuple(recv_req.timing or ())
            req.token_times.append(now)
            req.check_finished()
            if self.token_callback:
                self.token_callback(req, int(tok), req.finished())
            if req.finished():
                self._release(req)
            else:
                keep.append(i)
        if not keep:
            return
        if on_device:
            batch.output_ids = (recv_req.next_token_ids if len(keep) == len(batch.reqs) else
                                recv_req.next_token_ids[torch.tensor(keep, device=self.device)])
        else:
            batch.output_ids = torch.tensor([host_ids[i] for i in keep], dtype=torch.int64,
                                            pin_memory=True).to(self.device, non_blocking=True)
        if len(keep) != len(batch.reqs):
            batch.output_ids, out = None, batch.output_ids
            batch.filter_batch(keep_indices=keep)
            batch.output_ids = out
        self.running_batch.merge_batch(batch)

    # ---- decode ----------------------------------------------------------------------------------
    def _release(self, req: Req):
        """Free the KV runs and the request slot (ChunkCache.cache_finished_req,
        mem_cache/chunk_cache.py:37-42) and hand the request to the output side."""
        self.token_to_kv_pool_allocator.free_req(req.req_pool_idx)
        self.req_to_token_pool.free(req.req_pool_idx)
        req.finish_time = time.perf_counter()
        if req.prior_output_ids:
            # the request was retracted and re-prefilled on the way: the client sees ONE request
            # with its whole output, its original first-token time and every token's timestamp
            req.origin_input_ids = req.origin_input_ids[: len(req.origin_input_ids) - len(req.prior_output_ids)]
            req.output_ids = req.prior_output_ids + req.output_ids
            req.token_times = req.prior_token_times + req.token_times
            req.prior_output_ids, req.prior_token_times = [], []
        self.num_finished += 1
        if self.tp_rank == 0:
            self.finished.put(req)

    def update_running_batch(self, batch: ScheduleBatch) -> ScheduleBatch:
        # requests whose last token has already been LAUNCHED need no further step
        keep = [i for i, r in enumerate(batch.reqs)
                if not r.finished() and r.launched_tokens < r.sampling_params.max_new_tokens]
        if len(keep) != len(batch.reqs):
            batch.filter_batch(keep_indices=keep)
        if batch.is_empty():
            return batch
        if not batch.check_decode_mem():
            torch.cuda.current_stream().synchronize()  # in-flight steps still read the victims' KV
            self._drain_results(all_of_them=True)
            old = self.new_token_ratio
            retracted = batch.retract_decode(RETRACT_DECODE_STEPS)
            self.new_token_ratio = min(old + 0.1, 1.0)
            logger.info("Decode out of memory happened. #retracted_reqs: %d, #new_token_ratio: %.4f -> %.4f",
                        len(retracted), old, self.new_token_ratio)
            for req in retracted:
                # re-submit prompt + generated tokens as a fresh prefill (decode_scheduler.py:117-139)
                msg = TokenizedGenerateReqInput(rid=req.rid, input_text="",
                                                input_ids=req.origin_input_ids + req.output_ids,
                                                sampling_params=SamplingParams(
                                                    max_new_tokens=req.sampling_params.max_new_tokens - len(req.output_ids),
                                                    ignore_eos=req.sampling_params.ignore_eos),
                                                is_retracted=True, arrival_time=req.arrival_time)
                new_req = self.handle_generate_request(msg)
                # history the continuation must report when it finishes (stitched in _release)
                new_req.retracted_from = req
                new_req.prior_output_ids = req.prior_output_ids + req.output_ids
                new_req.prior_token_times = req.prior_token_times + req.token_times
                new_req.first_token_time = req.first_token_time
                self.num_retracted += 1
                self.send_to_p_instance.put(msg)
        else:
            self.new_token_ratio = max(self.new_token_ratio - self.new_token_ratio_decay,
                                       self.min_new_token_ratio)
        batch.prepare_for_decode()
        return batch

    def get_next_batch_to_run(self) -> Optional[ScheduleBatch]:
        if self.running_batch.is_empty():
            return None
        self.running_batch = self.update_running_batch(self.running_batch)
        return None if self.running_batch.is_empty() else self.running_batch

    def run_batch(self, batch: ScheduleBatch):
        batch.gather_logits = self.tp_size == 1
        mwb = batch.get_model_worker_batch()
        logits_output, next_token_ids = self.tp_worker.forward_batch_generation(mwb)
        self.forward_ct += 1
        batch.output_ids = next_token_ids.to(torch.int64)
        bs = len(batch.reqs)
        for r in batch.reqs:
            r.launched_tokens += 1
        pinned = None
        if not self.device_resident:
            pinned = self._pinned_ring[self._ring_i % len(self._pinned_ring)]
            self._ring_i += 1
            if bs > pinned.numel():
                pinned = torch.empty((bs,), dtype=torch.int32, pin_memory=True)
            pinned[:bs].copy_(next_token_ids, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self.result_queue.append((list(batch.reqs), pinned, bs, ev))
        self.decode_steps += 1
        self.decode_tokens += bs

    def process_batch_result_decode(self, reqs: List[Req], pinned: torch.Tensor, bs: int, ev):
        ev.synchronize()
        now = time.perf_counter()
        ids = pinned[:bs].tolist() if pinned is not None else [-1] * bs
        for req, tok in zip(reqs, ids):
            if req.finished():
                continue
            req.output_ids.append(tok)
            req.token_times.append(now)
            req.check_finished()
            if self.token_callback:
                self.token_callback(req, tok, req.finished())
            if req.finished():
                self._release(req)

    def _drain_results(self, all_of_them: bool = False, keep: int = 1, service_inbox: bool = False):
        """Read back the sampled ids of the oldest launched steps.  With `service_inbox` the wait
        for a step that is still running is spent on the inbox: a new request is admitted (so the
        prefill worker starts on it) and a prefill report is taken (so its first token is
        delivered) NOW instead of after the step — at a 4-5 ms decode step each of the two used
        to add half a step to every TTFT.  TP = 1 only: under TP the inbox is processed in
        lockstep, once per iteration."""
        while self.result_queue and (all_of_them or len(self.result_queue) > keep):
            if service_inbox and self.tp_size == 1:
                ev = self.result_queue[0][3]
                while not ev.query() and not self.stop_flag.is_set():
                    if self.inbox.empty():
                        time.sleep(0.00005)
                    else:
                        self.recv_requests()
                        self.maybe_schedule_prefill()
            self.process_batch_result_decode(*self.result_queue.popleft())

    def event_loop(self):
        """Overlap loop: step N is launched before step N-1's tokens are read back."""
        idle_spins = 0
        # TP > 1: every rank leaves the loop in the SAME iteration — when the "stop" rank 0
        # broadcast has been processed — or the others would block in the next broadcast
        while not (self._stop_broadcast if self.tp_size > 1 else self.stop_flag.is_set()):
            self.recv_requests()
            if self._stop_broadcast:
                break
            self.maybe_schedule_prefill()
            batch = self.get_next_batch_to_run()
            if batch is not None:
                self.run_batch(batch)
                self._drain_results(keep=1, service_inbox=True)
                idle_spins = 0
            else:
                self._drain_results(all_of_them=True)
                self.new_token_ratio = self.init_new_token_ratio  # reset when idle (scheduler.py:555)
                idle_spins += 1
                if idle_spins > 50:
                    time.sleep(0.0002)
        self._drain_results(all_of_them=True)


def split_sms(total: int, prefill_pct: int, decode_pct: int, granularity: int = 8):
    """Disjoint (prefill, decode) SM counts for the two green contexts.  The reference's caps are
    percentages that may overlap (80 % + 100 % by default, python/sglang/semi_pd/utils.py:10-11);
    here they are normalised to a split of the device.  Groups come in multiples of `granularity`
    SMs and 148 is not one: the decode group is rounded to the granularity and the prefill group
    gets what remains (50/50 -> 76 + 72; prefill is the throughput-bound side of a long-prompt
    mix).  Both sides always get at least one granule."""
    if prefill_pct <= 0 or decode_pct <= 0:
        raise ValueError("SM percentages must be positive")
    d = int(total * decode_pct / float(prefill_pct + decode_pct) / granularity + 0.5) * granularity
    d = min(max(d, granularity), total - granularity)
    return total - d, d


class SemiPDEngine:
    """`_launch_semi_pd_subprocesses` (entrypoints/engine.py:540-728) for one GPU: builds the D
    worker (weights + pools + CUDA graphs), the P worker (adopts D's parameters through the
    get_ipc_info / share_params_from_ipc call pair), the SM partition, and runs both loops."""

    def __init__(self, server_args: ServerArgs, gpu_id: int = 0, tp_rank: int = 0,
                 model_config: Optional[ModelConfig] = None, prefill_sm_percent: Optional[int] = None,
                 decode_sm_percent: Optional[int] = None, partition: str = "green"):
        assert server_args.enable_semi_pd, "SemiPDEngine needs --enable-semi-pd"
        self.server_args = server_args
        self.gpu_id = gpu_id
        torch.cuda.set_device(gpu_id)
        model_config = model_config or get_model_config(server_args.model_path)
        p_pct = PREFILL_ENGINE_SM_PERCENTILE if prefill_sm_percent is None else prefill_sm_percent
        d_pct = DECODE_ENGINE_SM_PERCENTILE if decode_sm_percent is None else decode_sm_percent
        total = ops.device_sm_count(gpu_id)
        self.partition = None
        if partition == "green":
            # the reference's caps overlap (80 % + 100 %); green contexts are disjoint, so the two
            # percentages are normalised to a split of the device
            p_sms, d_sms = split_sms(total, p_pct, d_pct)
            self.partition = ops.get_partition(gpu_id, p_sms, d_sms)
            self.p_stream, self.d_stream = self.partition.prefill_stream, self.partition.decode_stream
            self.p_sms, self.d_sms = self.partition.prefill_sms, self.partition.decode_sms
        elif partition == "none":
            # two plain streams sharing all SMs (decode gets the higher priority); for A/B runs
            self.p_stream = torch.cuda.Stream(device=gpu_id, priority=0)
            self.d_stream = torch.cuda.Stream(device=gpu_id, priority=-1)
            self.p_sms = self.d_sms = total
        else:
            raise ValueError(partition)

        self.tp_rank, self.tp_size = tp_rank, server_args.tp_size
        if self.tp_size > 1:
            # rendezvous before the workers build their communicators (model_runner.py:285-344)
            distributed.init_distributed_environment(self.tp_size, tp_rank, device_id=gpu_id)

        # D first: owns weights and the KV pool (engine.py:584-627)
        with torch.cuda.stream(self.d_stream):
            ops.set_sm_budget(self.d_sms)
            self.d_worker = TpModelWorker(server_args, gpu_id, tp_rank, instance_role=InstanceRole.DECODE,
                                          model_config=model_config)
            ipc_info = self.d_worker.get_ipc_info()
            self.d_worker.init_attention_backend()
            self.d_worker.init_cuda_graphs()
            self.d_stream.synchronize()
        # then P with bypass_load_weight=True (engine.py:629-661)
        with torch.cuda.stream(self.p_stream):
            ops.set_sm_budget(self.p_sms)
            self.p_worker = TpModelWorker(server_args, gpu_id, tp_rank, bypass_load_weight=True,
                                          instance_role=InstanceRole.PREFILL, model_config=model_config)
            self.p_worker.share_params_from_ipc(ipc_info)
            self.p_worker.init_attention_backend()
        ops.set_sm_budget(0)
        distributed.set_current_comm(None)
        self.max_total_num_tokens = self.d_worker.max_total_num_tokens

        self.decode = SemiPDDecodeScheduler(server_args, self.d_worker, self.d_stream, self.d_sms)
        self.prefill = SemiPDPrefillScheduler(server_args, self.p_worker, self.p_stream, self.p_sms,
                                              self.decode.inbox)
        self.decode.send_to_p_instance = self.prefill.inbox
        self.decode.bridge_socket = self.prefill.bridge_socket
        self.prefill.local_result_queue = self.decode.local_prefill_results
        if self.tp_size > 1:
            import datetime

            import torch.distributed as dist
            self.decode.cpu_group = dist.new_group(ranks=list(range(self.tp_size)), backend="gloo",
                                                   timeout=datetime.timedelta(seconds=300))
        self.threads = [threading.Thread(target=self.decode.run, name="semi-pd-decode", daemon=True),
                        threading.Thread(target=self.prefill.run, name="semi-pd-prefill", daemon=True)]
        for t in self.threads:
            t.start()
        self._rid = itertools.count(1)
        self._closed = False

    def warmup(self, prompt_len: int = 64, n: int = 2, max_new_tokens: int = 4):
        """Run a few requests through both workers (module loads, tensor-map cache, allocator)."""
        vocab = self.d_worker.model_config.vocab_size
        reqs = [self.submit([(3 + i + j) % vocab for j in range(prompt_len)], max_new_tokens) for i in range(n)]
        self.wait(len(reqs))
        return reqs

    def set_device_resident(self, flag: bool):
        """Only while idle: switches both workers between host-buffer serving and the
        device-resident arm."""
        self.decode.device_resident = self.prefill.device_resident = bool(flag)

    def submit(self, input_ids: List[int], max_new_tokens: int, arrival_time: Optional[float] = None,
               ignore_eos: bool = True, input_ids_device: Optional[torch.Tensor] = None) -> str:
        """TokenizerManager.generate_request → push to D and P (tokenizer_manager.py:149-160).
        Raises ValueError for a request the reference would abort with BAD_REQUEST (input longer
        than the context / KV pool allows); max_new_tokens is clipped like scheduler.py:791-798.
        Under TP only rank 0 accepts requests (its decode scheduler broadcasts them)."""
        if self.tp_size > 1 and self.tp_rank != 0:
            raise RuntimeError("TP > 1: requests are submitted on rank 0 only")
        if self.tp_size > 1 and (input_ids_device is not None or self.decode.device_resident):
            raise RuntimeError("the device-resident arm is a TP = 1 measurement mode")
        input_ids, max_new_tokens = self.decode.validate_request(list(input_ids), max_new_tokens)
        rid = f"req-{next(self._rid)}"  # itertools.count: atomic under the GIL (FastAPI threadpool)
        msg = TokenizedGenerateReqInput(rid=rid, input_text="", input_ids=input_ids,
                                        sampling_params=SamplingParams(max_new_tokens=max_new_tokens,
                                                                       ignore_eos=ignore_eos),
                                        arrival_time=time.perf_counter() if arrival_time is None else arrival_time,
                                        input_ids_device=input_ids_device)
        self.prefill.inbox.put(msg)  # P first so that it knows the rid before D can schedule it
        self.decode.inbox.put(msg)
        return rid

    def wait(self, n: int, timeout: float = 600.0) -> List[Req]:
        out, t0 = [], time.perf_counter()
        while len(out) < n:
            self.check_health()
            try:
                out.append(self.decode.finished.get(timeout=0.05))
            except queue.Empty:
                if time.perf_counter() - t0 > timeout:
                    raise TimeoutError(f"only {len(out)}/{n} requests finished in {timeout}s")
        return out

    def check_health(self):
        for s in (self.decode, self.prefill):
            if s.error is not None:
                raise RuntimeError(f"{s.role} scheduler died") from s.error

    def join(self, timeout: Optional[float] = None):
        """TP ranks > 0: serve until rank 0's shutdown has been broadcast.  The broadcast ends the
        decode loop; the prefill loop of this rank has no other way to learn about it."""
        self.threads[0].join(timeout=timeout)
        if not self.threads[0].is_alive():
            self.prefill.stop_flag.set()
        self.threads[1].join(timeout=timeout)

    def shutdown(self):
        """Stop both loops and drop everything the engine owns on the device: CUDA graphs, the
        workers' tensors (weights, KV pool), the TP communicators.  The SM partition itself is a
        process-lifetime object (ops.get_partition): the next engine on this GPU reuses its green
        contexts and streams.  Under TP rank 0's call ends the loops of all ranks; the other
        ranks' calls wait for that."""
        if self._closed:
            return
        self._closed = True
        if self.tp_size == 1 or self.tp_rank == 0:
            self.decode.stop_flag.set()
        self.threads[0].join(timeout=60)
        self.prefill.stop_flag.set()
        self.decode.stop_flag.set()
        for t in self.threads:
            t.join(timeout=10)
        torch.cuda.synchronize()
        for w in (self.d_worker, self.p_worker):
            mr = w.model_runner
            mr.cuda_graph_runner = None
            mr.attn_backend = None
            if mr.tp_comm is not None:
                mr.tp_comm.close()
                mr.tp_comm = None
            mr.model = None
            mr.token_to_kv_pool = mr.token_to_kv_pool_allocator = mr.req_to_token_pool = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
