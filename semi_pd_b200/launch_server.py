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
    sa.argv = list(argv)
    if sa.enable_semi_pd:
        sa.disable_radix_cache = True  # server_args.py:325-331
    return sa


def _tp_follower(rank: int, world: int, port: int, argv: List[str], model_config, partition: str):
    """TP rank > 0 of `--tp N`: one process per GPU (the reference spawns one scheduler process
    pair per TP rank, entrypoints/engine.py:560-700).  It serves no HTTP: rank 0's decode scheduler
    broadcasts every request, and its shutdown."""
    import os

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from .semi_pd_scheduler import SemiPDEngine
    sa = prepare_server_args(argv)
    eng = SemiPDEngine(sa, gpu_id=rank, tp_rank=rank, model_config=model_config, partition=partition)
    eng.join()
    eng.shutdown()


def build_app(server_args: ServerArgs, model_config=None, gpu_id: int = 0, argv: Optional[List[str]] = None):
    """The FastAPI app + the engine behind it (separate from `launch_server` so that a test can
    drive the endpoints in-process).  A request the engine rejects (input longer than the context
    / KV pool allows) is answered with 400, as the reference aborts it with BAD_REQUEST
    (managers/utils.py:10-44)."""
    if not server_args.enable_semi_pd:
        raise SystemExit("this build implements the semi-PD path: pass --enable-semi-pd")
    from fastapi import FastAPI, HTTPException

    from .semi_pd_scheduler import SemiPDEngine

    partition = getattr(server_args, "sm_partition", "green")
    followers = []
    if server_args.tp_size > 1:
        # `--tp N`: this process is TP rank 0 on GPU 0 and serves HTTP; ranks 1..N-1 are spawned,
        # one process per GPU, and rendezvous over 127.0.0.1
        import multiprocessing as mp
        import os
        import socket
        if argv is None:
            raise SystemExit("--tp > 1 needs the command line (the follower ranks re-parse it)")
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        ctx = mp.get_context("spawn")
        for r in range(1, server_args.tp_size):
            pr = ctx.Process(target=_tp_follower, args=(r, server_args.tp_size, port, list(argv), model_config,
                                                        partition), daemon=True)
            pr.start()
            followers.append(pr)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0",
                          WORLD_SIZE=str(server_args.tp_size))
        gpu_id = 0
    engine = SemiPDEngine(server_args, gpu_id=gpu_id, tp_rank=0, model_config=model_config, partition=partition)
    engine.tp_followers = followers
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
    app, engine = build_app(server_args, argv=getattr(server_args, "argv", None))
    try:
        uvicorn.run(app, host=server_args.host, port=server_args.port, log_level="warning")
    finally:
        engine.shutdown()
        for pr in getattr(engine, "tp_followers", []):
            pr.join(timeout=30)


def main(argv: List[str]):
    launch_server(prepare_server_args(argv))
