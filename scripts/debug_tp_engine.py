"""2-GPU debug run of the TP semi-PD engine test with periodic all-thread stack dumps
(gpurun --gpus 2 -- 'timeout 240 python scripts/debug_tp_engine.py').  Writes
gpurun_out/tp_debug_rank{r}.log."""
import faulthandler
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, port, ret):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    f = open(os.path.join(ROOT, "gpurun_out", f"tp_debug_rank{rank}.log"), "w")
    faulthandler.enable(file=f, all_threads=True)
    faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "60")), repeat=True, file=f)
    sys.stderr = f
    from tests.test_gpu_tp import _engine_worker
    _engine_worker(rank, world, port, ret)
    print(f"rank {rank} done", file=f, flush=True)


if __name__ == "__main__":
    from tests.test_gpu_tp import _free_port
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        print(dict(ret))
