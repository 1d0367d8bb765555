"""Server entry: `launch_server(server_args)` (python/sglang/srt/entrypoints/http_server.py:603-680).
Builds the semi-PD engine and serves `/generate` in TOKEN-ID space (there is no tokenizer / weights
offline: requests carry `input_ids`, responses carry `output_ids`, as `bench_serving
--dataset-name random` effectively exercises the reference, bench_serving.py:772-783), plus
`/health` and `/get_server_info`.  HTTP, OpenAI protocol and detokenisation are outside the hot
path and intentionally minimal.
"""
from __future__ import annotations

import argparse
import threading
import time
from typing import List, Optional

from pydantic import BaseModel

from .configs import ServerArgs


class GenerateReq(BaseModel):
    """Body of POST /generate (io_struct.GenerateReqInput restricted to token ids)."""
    input_ids: List[int]
    sampling_params: Optional[dict] = None


def prepare_server_args(argv: List[str]) -> ServerArgs:
    parser = argparse.ArgumentParser()
    ServerArgs.add_cli_args(parser)
    parser.add_argument("--sm-partition", type=str, default="green", choices=["green", "none"])
    ns = parser.parse_args(argv)
    sa = ServerArgs.from_cli_args(ns)
    sa.sm_partition = ns.sm_partition
    if sa.enable_semi_pd:
        sa.disable_radix_cache = True  # server_args.py:325-331
    return sa


def build_app(server_args: ServerArgs, model_config=None, gpu_id: int = 0):
    """The FastAPI app + the engine behind it (separate from `launch_server` so that a test can
    drive the endpoints in-process).  A request the engine rejects (input longer than the context
    / KV pool allows) is answered with 400, as the reference aborts it with BAD_REQUEST
    (managers/utils.py:10-44)."""
    if not server_args.enable_semi_pd:
        raise SystemExit("this build implements the semi-PD path: pass --enable-semi-pd")
    if server_args.tp_size != 1:
        raise SystemExit("launch_server runs TP=1 replicas; use bench.py --gpus N for the TP engine")
    from fastapi import FastAPI, HTTPException

    from .semi_pd_scheduler import SemiPDEngine

    engine = SemiPDEngine(server_args, gpu_id=gpu_id, model_config=model_config,
                          partition=getattr(server_args, "sm_partition", "green"))
    engine.warmup()
    results, cond = {}, threading.Condition()

    def collector():
        while not engine._closed:
            try:
                req = engine.decode.finished.get(timeout=0.2)
            except Exception:  # noqa: BLE001  (queue.Empty)
                continue
            with cond:
                results[req.rid] = req
                cond.notify_all()

    threading.Thread(target=collector, daemon=True).start()
    app = FastAPI()

    @app.get("/health")
    def health():
        engine.check_health()
        return {"status": "ok"}

    @app.get("/get_server_info")
    def info():
        return {"max_total_num_tokens": engine.max_total_num_tokens, "prefill_sms": engine.p_sms,
                "decode_sms": engine.d_sms, "model_path": server_args.model_path}

    @app.post("/generate")
    def generate(r: GenerateReq):
        sp = r.sampling_params or {}
        t0 = time.perf_counter()
        try:
            rid = engine.submit(r.input_ids, int(sp.get("max_new_tokens", 128)),
                                ignore_eos=bool(sp.get("ignore_eos", True)))
        except ValueError as e:
            raise HTTPException(status_code=400, detail=str(e))
        with cond:
            if not cond.wait_for(lambda: rid in results, timeout=600):
                raise HTTPException(status_code=504, detail="generation timed out")
            req = results.pop(rid)
        return {"output_ids": req.output_ids,
                "meta_info": {"id": rid, "prompt_tokens": len(r.input_ids),
                              "completion_tokens": len(req.output_ids),
                              "ttft": req.first_token_time - t0, "e2e_latency": time.perf_counter() - t0}}

    return app, engine


def launch_server(server_args: ServerArgs):
    import uvicorn
    app, _ = build_app(server_args)
    uvicorn.run(app, host=server_args.host, port=server_args.port, log_level="warning")


def main(argv: List[str]):
    launch_server(prepare_server_args(argv))
