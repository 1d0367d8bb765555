"""ModelRunner / CudaGraphRunner — the operator surface the scheduler calls
(python/sglang/srt/model_executor/model_runner.py:98-1335, cuda_graph_runner.py:165-558).

Semi-PD specifics (model_runner.py:346-624): the reference's decode process exports every weight,
every per-layer K/V buffer and `req_to_token` as cudaIpcMemHandles (`get_ipc_info`) and the
prefill process maps them (`share_params_from_ipc`).  Here both workers live in one process, so
`get_ipc_info()` returns an `IPCInfo` holding plain object references and
`share_params_from_ipc()` adopts them — same call sequence, zero copies, zero handles, and the
second copy of activations / workspaces that makes the reference scale mem-fraction by 0.9 is the
only thing duplicated.
"""
from __future__ import annotations

import bisect
import dataclasses
import logging
from typing import Optional

import torch

from . import ops
from ._lib import _Counter
from .attention_backend import SpdAttnBackend
from .configs import InstanceRole, ModelConfig, ServerArgs, default_cuda_graph_bs
from .distributed import TpComm, init_distributed_environment, set_current_comm
from .forward_batch_info import ForwardBatch, ForwardMode, LogitsProcessorOutput
from .layers import Sampler
from .llama import LlamaForCausalLM
from .memory_pool import MHATokenToKVPool, MLATokenToKVPool, ReqToTokenPool, TokenToKVPoolAllocator

logger = logging.getLogger(__name__)


@dataclasses.dataclass
class IPCInfo:
    """In-process stand-in for python/sglang/semi_pd/utils.py:14-22 (handles → references)."""
    model: LlamaForCausalLM
    req_to_token_pool: ReqToTokenPool
    token_to_kv_pool: MHATokenToKVPool
    token_to_kv_pool_allocator: TokenToKVPoolAllocator
    max_total_num_tokens: int


class ModelRunner:
    """ModelRunner runs the forward passes of the models."""

    def __init__(self, model_config: ModelConfig, mem_fraction_static: Optional[float], gpu_id: int,
                 tp_rank: int, tp_size: int, nccl_port: int, server_args: ServerArgs,
                 is_draft_worker: bool = False, bypass_load_weight: bool = False,
                 instance_role: InstanceRole = InstanceRole.OTHER):
        self.model_config = model_config
        self.mem_fraction_static = mem_fraction_static
        self.device = torch.device("cuda", gpu_id)
        self.gpu_id = gpu_id
        self.tp_rank = tp_rank
        self.tp_size = tp_size
        self.server_args = server_args
        self.is_draft_worker = is_draft_worker
        self.bypass_load_weight = bypass_load_weight
        self.instance_role = instance_role
        self.dtype = server_args.torch_dtype()
        self.kv_cache_dtype = self.dtype
        self.sliding_window_size = None
        self.page_size = 1
        self.attn_backend = None
        self.cuda_graph_runner = None
        self.model = None
        self.req_to_token_pool = None
        self.token_to_kv_pool = None
        self.token_to_kv_pool_allocator = None
        self.max_total_num_tokens = None
        self.sampler = Sampler()

        torch.cuda.set_device(self.device)
        init_distributed_environment(tp_size, tp_rank, device_id=gpu_id)
        # every worker (the prefill and the decode worker of a GPU alike) owns its communicator —
        # the reference builds separate NCCL worlds for P and D (managers/scheduler.py:250-259);
        # it is installed as the calling thread's current communicator before each forward
        self.tp_comm = TpComm(self.device) if tp_size > 1 else None
        self.tp_group = self.tp_comm.group if self.tp_comm is not None else None

        if not bypass_load_weight:
            self.load_model()
            self.init_memory_pool(server_args.max_total_tokens, server_args.max_running_requests)
        # under semi-PD the attention backend and CUDA graphs are initialised explicitly by the
        # scheduler entry point after the IPC exchange (model_runner.py:222-228)
        if not server_args.enable_semi_pd:
            self.init_attention_backend()
            self.init_cuda_graphs()

    # ------------------------------------------------------------------------------------------
    # model + memory
    # ------------------------------------------------------------------------------------------
    def load_model(self):
        if getattr(self.model_config, "model_type", "llama") == "opt":
            from .opt import OPTForCausalLM
            self.model = OPTForCausalLM(self.model_config, dtype=self.dtype, device=self.device)
        elif self.model_config.model_type == "deepseek_v2":
            from .deepseek_v2 import DeepseekV2ForCausalLM
            self.model = DeepseekV2ForCausalLM(self.model_config, dtype=self.dtype, device=self.device)
        else:
            self.model = LlamaForCausalLM(self.model_config, dtype=self.dtype, device=self.device)
        if self.server_args.load_format == "dummy":
            self.model.init_dummy_weights(self.server_args.random_seed)
        elif self.server_args.load_format != "external":
            raise NotImplementedError("only --load-format dummy (synthetic weights) or weights "
                                      "passed through load_weights() are supported")

    def profile_max_num_token(self, total_gpu_memory_bytes: int) -> int:
        """model_runner.py:924-949: tokens that fit in mem_fraction_static of the device after
        the weights."""
        cfg = self.model_config
        esz = torch.tensor([], dtype=self.kv_cache_dtype).element_size()
        if getattr(cfg, "model_type", "llama") == "deepseek_v2":
            cell = (cfg.kv_lora_rank + cfg.qk_rope_head_dim) * cfg.num_hidden_layers * esz  # model_runner.py:930-936
        else:
            cell = cfg.get_num_kv_heads(self.tp_size) * cfg.head_dim * cfg.num_hidden_layers * 2 * esz
        free, total = torch.cuda.mem_get_info(self.device)
        frac = self.mem_fraction_static if self.mem_fraction_static is not None else 0.85
        rest = free - total * (1 - frac)
        return max(int(rest // cell), 0)

    def init_memory_pool(self, max_total_tokens: Optional[int] = None,
                         max_num_reqs: Optional[int] = None):
        cfg = self.model_config
        profiled = self.profile_max_num_token(0)
        self.max_total_num_tokens = min(profiled, max_total_tokens) if max_total_tokens else profiled
        if self.max_total_num_tokens <= 0:
            raise RuntimeError("Not enough memory. Please try to increase --mem-fraction-static.")
        if max_num_reqs is None:
            max_num_reqs = min(max(int(self.max_total_num_tokens / cfg.context_len * 512), 2048), 4096)
        ctx = self.server_args.context_length or cfg.context_len
        self.req_to_token_pool = ReqToTokenPool(size=max_num_reqs + 1, max_context_len=ctx + 4,
                                                device=self.device)
        if getattr(cfg, "model_type", "llama") == "deepseek_v2":
            self.token_to_kv_pool = MLATokenToKVPool(
                self.max_total_num_tokens, page_size=1, dtype=self.kv_cache_dtype, kv_lora_rank=cfg.kv_lora_rank,
                qk_rope_head_dim=cfg.qk_rope_head_dim, layer_num=cfg.num_hidden_layers, device=self.device)
        else:
            self.token_to_kv_pool = MHATokenToKVPool(
                self.max_total_num_tokens, page_size=1, dtype=self.kv_cache_dtype,
                head_num=cfg.get_num_kv_heads(self.tp_size), head_dim=cfg.head_dim,
                layer_num=cfg.num_hidden_layers, device=self.device)
        self.token_to_kv_pool_allocator = TokenToKVPoolAllocator(
            self.max_total_num_tokens, dtype=self.kv_cache_dtype, device=self.device,
            kvcache=self.token_to_kv_pool)

    # ------------------------------------------------------------------------------------------
    # semi-PD sharing (model_runner.py:346-624)
    # ------------------------------------------------------------------------------------------
    def get_ipc_info(self) -> IPCInfo:
        return IPCInfo(self.model, self.req_to_token_pool, self.token_to_kv_pool,
                       self.token_to_kv_pool_allocator, self.max_total_num_tokens)

    def share_params_from_ipc(self, ipc_info: IPCInfo):
        self.model = ipc_info.model
        self.req_to_token_pool = ipc_info.req_to_token_pool
        self.token_to_kv_pool = ipc_info.token_to_kv_pool
        self.token_to_kv_pool_allocator = ipc_info.token_to_kv_pool_allocator
        self.max_total_num_tokens = ipc_info.max_total_num_tokens

    # ------------------------------------------------------------------------------------------
    def init_attention_backend(self):
        if self.server_args.attention_backend not in ("spd", "flashinfer", None):
            raise ValueError("this build has exactly one attention backend ('spd'); got "
                             f"{self.server_args.attention_backend}")
        self.attn_backend = SpdAttnBackend(self)

    def init_cuda_graphs(self):
        self.cuda_graph_runner = None
        if self.server_args.disable_cuda_graph:
            return
        if self.tp_size > 1 and (self.tp_comm is None or self.tp_comm.max_bytes == 0):
            # only libspd's peer-memory collectives are captured (torch.distributed's NCCL
            # all-reduce inside the decode graph hung on the pool in round 1)
            logger.warning("TP > 1 without the libspd communicator: decode CUDA graphs disabled")
            return
        self.cuda_graph_runner = CudaGraphRunner(self)

    # ------------------------------------------------------------------------------------------
    # forward (model_runner.py:1219-1325)
    # ------------------------------------------------------------------------------------------
    def forward_decode(self, forward_batch: ForwardBatch) -> LogitsProcessorOutput:
        self.attn_backend.init_forward_metadata(forward_batch)
        return self.model.forward(forward_batch.input_ids, forward_batch.positions, forward_batch)

    def forward_extend(self, forward_batch: ForwardBatch, skip_attn_backend_init: bool = False
                       ) -> LogitsProcessorOutput:
        if not skip_attn_backend_init:
            self.attn_backend.init_forward_metadata(forward_batch)
        return self.model.forward(forward_batch.input_ids, forward_batch.positions, forward_batch)

    def forward_idle(self, forward_batch: ForwardBatch):
        return LogitsProcessorOutput(next_token_logits=torch.empty(
            (0, self.model_config.vocab_size), device=self.device, dtype=torch.float32))

    def forward(self, forward_batch: ForwardBatch, skip_attn_backend_init: bool = False
                ) -> LogitsProcessorOutput:
        set_current_comm(self.tp_comm)
        if (forward_batch.forward_mode.is_cuda_graph() and self.cuda_graph_runner
                and self.cuda_graph_runner.can_run(forward_batch)):
            return self.cuda_graph_runner.replay(forward_batch,
                                                 skip_attn_backend_init=skip_attn_backend_init)
        if forward_batch.forward_mode.is_decode():
            return self.forward_decode(forward_batch)
        elif forward_batch.forward_mode.is_extend():
            return self.forward_extend(forward_batch, skip_attn_backend_init=skip_attn_backend_init)
        elif forward_batch.forward_mode.is_idle():
            return self.forward_idle(forward_batch)
        raise ValueError(f"Invalid forward mode: {forward_batch.forward_mode}")

    def sample(self, logits_output: LogitsProcessorOutput, forward_batch: ForwardBatch) -> torch.Tensor:
        if logits_output.next_token_ids is not None:
            return logits_output.next_token_ids
        return self.sampler(logits_output, forward_batch.sampling_info)


class CudaGraphRunner:
    """Captures the DECODE forward for bs ∈ {1,2,4,8,16,…,160}; replay pads the batch up to the
    next captured size (padded rows: seq_len = 1, out_cache_loc = 0 = the dummy slot) and copies
    the five input tensors into static buffers (cuda_graph_runner.py:109-149,455-531).  Unlike the
    flashinfer backend there is no host-side re-plan per replay."""

    def __init__(self, model_runner: ModelRunner):
        self.model_runner = model_runner
        self.graphs = {}
        self.output_buffers = {}
        self.kernels_per_graph = {}
        sa = model_runner.server_args
        bs_list = sa.cuda_graph_bs or default_cuda_graph_bs(sa.cuda_graph_max_bs)
        self.capture_bs = sorted(b for b in bs_list if b <= model_runner.req_to_token_pool.size)
        self.max_bs = max(self.capture_bs)
        model_runner.attn_backend.init_cuda_graph_state(self.max_bs)
        self.seq_len_fill_value = model_runner.attn_backend.get_cuda_graph_seq_len_fill_value()
        dev = model_runner.device
        self.input_ids = torch.zeros((self.max_bs,), dtype=torch.int64, device=dev)
        self.req_pool_indices = torch.zeros((self.max_bs,), dtype=torch.int64, device=dev)
        self.seq_lens = torch.full((self.max_bs,), self.seq_len_fill_value, dtype=torch.int64, device=dev)
        self.out_cache_loc = torch.zeros((self.max_bs,), dtype=torch.int64, device=dev)
        self.positions = torch.zeros((self.max_bs,), dtype=torch.int64, device=dev)
        self.pool = None
        self.capture()

    def can_run(self, forward_batch: ForwardBatch) -> bool:
        return forward_batch.forward_mode.is_decode() and forward_batch.batch_size <= self.max_bs

    def capture(self):
        set_current_comm(self.model_runner.tp_comm)
        stream = torch.cuda.current_stream()
        for bs in reversed(self.capture_bs):
            graph, out = self.capture_one_batch_size(bs, stream)
            self.graphs[bs] = graph
            self.output_buffers[bs] = out

    def capture_one_batch_size(self, bs: int, stream):
        mr = self.model_runner
        fb = ForwardBatch(
            forward_mode=ForwardMode.DECODE, batch_size=bs, input_ids=self.input_ids[:bs],
            req_pool_indices=self.req_pool_indices[:bs], seq_lens=self.seq_lens[:bs],
            out_cache_loc=self.out_cache_loc[:bs], seq_lens_sum=int(bs * self.seq_len_fill_value),
            positions=self.positions[:bs], req_to_token_pool=mr.req_to_token_pool,
            token_to_kv_pool=mr.token_to_kv_pool, attn_backend=mr.attn_backend)
        mr.attn_backend.init_forward_metadata_capture_cuda_graph(
            bs, bs, self.req_pool_indices[:bs], self.seq_lens[:bs], None, ForwardMode.DECODE, None)

        def run_once():
            # positions are derived inside the graph from the static seq_lens buffer
            ops.clamp_position(self.seq_lens[:bs], out=self.positions[:bs])
            out = mr.model.forward(fb.input_ids, fb.positions, fb)
            ids = out.next_token_ids if out.next_token_ids is not None else mr.sampler(out)
            return out.next_token_logits, ids

        for _ in range(2):  # warm up (allocations, tensor-map cache, lazy module loads)
            torch.cuda.synchronize()
            run_once()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _Counter.launches
        with torch.cuda.graph(graph, pool=self.pool, stream=stream if stream != torch.cuda.default_stream() else None):
            out = run_once()
        self.kernels_per_graph[bs] = _Counter.launches - n0  # libspd kernels inside this graph
        self.pool = graph.pool()
        torch.cuda.synchronize()
        return graph, out

    def replay_prepare(self, forward_batch: ForwardBatch):
        raw_bs = forward_batch.batch_size
        index = bisect.bisect_left(self.capture_bs, raw_bs)
        bs = self.capture_bs[index]
        if bs != raw_bs:
            self.seq_lens.fill_(self.seq_len_fill_value)
            self.out_cache_loc.zero_()
        self.input_ids[:raw_bs].copy_(forward_batch.input_ids)
        self.req_pool_indices[:raw_bs].copy_(forward_batch.req_pool_indices)
        self.seq_lens[:raw_bs].copy_(forward_batch.seq_lens)
        self.out_cache_loc[:raw_bs].copy_(forward_batch.out_cache_loc)
        self.model_runner.attn_backend.init_forward_metadata_replay_cuda_graph(
            bs, self.req_pool_indices, self.seq_lens, forward_batch.seq_lens_sum + (bs - raw_bs),
            None, forward_batch.forward_mode, None, seq_lens_cpu=None)
        self.raw_bs = raw_bs
        self.bs = bs

    def replay(self, forward_batch: ForwardBatch, skip_attn_backend_init: bool = False
               ) -> LogitsProcessorOutput:
        if not skip_attn_backend_init:
            self.replay_prepare(forward_batch)
        self.graphs[self.bs].replay()
        _Counter.launches += self.kernels_per_graph[self.bs]
        next_token_logits, next_ids = self.output_buffers[self.bs]
        # greedy ids were computed inside the graph
        return LogitsProcessorOutput(next_token_logits=next_token_logits[: self.raw_bs],
                                     next_token_ids=next_ids[: self.raw_bs])
