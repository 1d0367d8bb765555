"""TpModelWorker — the worker surface `Scheduler.run_batch` calls
(python/sglang/srt/managers/tp_worker.py:43-244): `forward_batch_generation`, `get_worker_info`,
`get_memory_pool`, and the semi-PD extras `get_ipc_info / share_params_from_ipc /
init_attention_backend / init_cuda_graphs` (:140-150).
"""
from __future__ import annotations

from typing import Optional

import torch

from .configs import InstanceRole, ModelConfig, ServerArgs, get_model_config
from .forward_batch_info import ForwardBatch, ModelWorkerBatch
from .model_runner import IPCInfo, ModelRunner


class TpModelWorker:
    """A tensor parallel model worker."""

    def __init__(self, server_args: ServerArgs, gpu_id: int, tp_rank: int, dp_rank: Optional[int] = None,
                 nccl_port: int = 0, is_draft_worker: bool = False, bypass_load_weight: bool = False,
                 instance_role: InstanceRole = InstanceRole.OTHER,
                 model_config: Optional[ModelConfig] = None):
        self.tp_rank = tp_rank
        self.server_args = server_args
        self.model_config = model_config or get_model_config(server_args.model_path)
        if server_args.context_length:
            self.model_config.max_position_embeddings = max(
                self.model_config.max_position_embeddings, server_args.context_length)
        self.model_runner = ModelRunner(
            model_config=self.model_config, mem_fraction_static=server_args.mem_fraction_static,
            gpu_id=gpu_id, tp_rank=tp_rank, tp_size=server_args.tp_size, nccl_port=nccl_port,
            server_args=server_args, is_draft_worker=is_draft_worker,
            bypass_load_weight=bypass_load_weight, instance_role=instance_role)
        self.device = self.model_runner.device
        self.random_seed = server_args.random_seed
        self._finalize_limits()

    def _finalize_limits(self):
        mr = self.model_runner
        if mr.max_total_num_tokens is None:
            return
        self.max_total_num_tokens = mr.max_total_num_tokens
        self.max_prefill_tokens = self.server_args.max_prefill_tokens
        self.max_running_requests = min(
            self.max_total_num_tokens // 2 if self.server_args.max_running_requests is None
            else self.server_args.max_running_requests, mr.req_to_token_pool.size - 1)
        # --context-length overrides the model's (it also sizes the req_to_token rows,
        # model_runner.py:1039-1045), tp_worker.py:100-107
        context_len = self.server_args.context_length or self.model_config.context_len
        self.max_req_len = min(context_len - 1, self.max_total_num_tokens - 1)
        self.max_req_input_len = self.max_req_len - 5

    def get_worker_info(self):
        mr = self.model_runner
        return (self.max_total_num_tokens, self.max_prefill_tokens, self.max_running_requests,
                self.max_req_len, self.max_req_input_len, self.random_seed, self.device,
                None, mr.req_to_token_pool.size, mr.req_to_token_pool.max_context_len,
                mr.token_to_kv_pool.size)

    def get_memory_pool(self):
        return self.model_runner.req_to_token_pool, self.model_runner.token_to_kv_pool_allocator

    def get_tp_cpu_group(self):
        return None

    # semi-PD extras ------------------------------------------------------------------------------
    def get_ipc_info(self) -> IPCInfo:
        return self.model_runner.get_ipc_info()

    def share_params_from_ipc(self, ipc_info: IPCInfo):
        self.model_runner.share_params_from_ipc(ipc_info)
        self._finalize_limits()

    def init_attention_backend(self):
        self.model_runner.init_attention_backend()

    def init_cuda_graphs(self):
        self.model_runner.init_cuda_graphs()

    # ---------------------------------------------------------------------------------------------
    def forward_batch_generation(self, model_worker_batch: ModelWorkerBatch, launch_done=None,
                                 skip_sample: bool = False):
        forward_batch = ForwardBatch.init_new(model_worker_batch, self.model_runner)
        logits_output = self.model_runner.forward(forward_batch)
        if launch_done:
            launch_done.set()
        next_token_ids = None if skip_sample else self.model_runner.sample(logits_output, forward_batch)
        return logits_output, next_token_ids
