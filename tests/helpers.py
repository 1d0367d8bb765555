"""Shared helpers for the GPU parity tests: build the product ModelRunner from the oracle's
weights and drive both with identical batches."""
import torch

from oracle import llama_oracle as O


def make_runner(cfg_kwargs, dtype, kv_slots=4096, max_reqs=16, ctx=2048, cuda_graph=False,
                cuda_graph_bs=None, oracle_mode="fp32", seed=1234, name="tiny-llama-test"):
    from semi_pd_b200.configs import ModelConfig, ServerArgs
    from semi_pd_b200.model_runner import ModelRunner

    ocfg = O.OracleLlamaConfig(**cfg_kwargs)
    weights = O.make_weights(ocfg, dtype, seed=seed)
    mcfg = ModelConfig(name=name, **cfg_kwargs)
    sa = ServerArgs(model_path=name, dtype="bfloat16" if dtype == torch.bfloat16 else "float16",
                    load_format="external", max_total_tokens=kv_slots, max_running_requests=max_reqs,
                    context_length=ctx, disable_cuda_graph=not cuda_graph,
                    cuda_graph_bs=cuda_graph_bs, enable_semi_pd=True)
    mr = ModelRunner(mcfg, 0.5, 0, 0, 1, 0, sa)
    mr.model.load_weights(weights.items())
    mr.init_attention_backend()
    mr.init_cuda_graphs()
    oracle = O.LlamaOracle(ocfg, weights, kv_slots=kv_slots, max_reqs=max_reqs + 1, max_ctx=ctx + 4,
                           compute=oracle_mode)
    return mr, oracle, ocfg


class GpuSession:
    """Drives the product path through ScheduleBatch → ModelWorkerBatch → ForwardBatch →
    ModelRunner.forward, mirroring oracle.OracleSession's extend/decode calls."""

    def __init__(self, mr):
        from semi_pd_b200.schedule_batch import Req, SamplingParams, ScheduleBatch
        self.mr = mr
        self.Req, self.SP, self.SB = Req, SamplingParams, ScheduleBatch
        self.reqs = {}
        self.batch = None

    def _run(self, sb):
        from semi_pd_b200.forward_batch_info import ForwardBatch
        fb = ForwardBatch.init_new(sb.get_model_worker_batch(), self.mr)
        out = self.mr.forward(fb)
        ids = self.mr.sample(out, fb)
        return out.next_token_logits.float().cpu(), ids.cpu()

    def extend(self, prompts, reqs=None):
        if reqs is None:
            reqs = [self.Req(f"r{len(self.reqs) + i}", p, self.SP()) for i, p in enumerate(prompts)]
            for r in reqs:
                self.reqs[r.rid] = r
                r.init_next_round_input()
        else:
            for r, p in zip(reqs, prompts):
                r.prefix_indices = [0] * len(r.fill_ids)
                r.origin_input_ids = r.origin_input_ids + list(p)
                r.init_next_round_input()
        sb = self.SB.init_new(reqs, self.mr.req_to_token_pool, self.mr.token_to_kv_pool_allocator,
                              self.mr.device)
        sb.prepare_for_extend()
        logits, ids = self._run(sb)
        self.batch = sb
        return reqs, logits, ids

    def decode(self, reqs, tokens):
        sb = self.batch
        assert [r.rid for r in sb.reqs] == [r.rid for r in reqs]
        for r, t in zip(reqs, tokens):
            r.output_ids.append(int(t))
        sb.output_ids = torch.tensor(tokens, dtype=torch.int64, device=self.mr.device)
        sb.prepare_for_decode()
        return self._run(sb)
