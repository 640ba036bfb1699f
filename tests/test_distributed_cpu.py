"""world_size-2 gloo test (CPU) of the multi-GPU host logic: interleaved row sharding and the
all-gather reassembly used by bench.py. The 'render' is faked (each rank writes its global row
index), the collective and the index arithmetic are the real ones."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, height, width, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        mdist = importlib.import_module("monte-carlo-ray-tracer_b200.distributed")
        y_first, y_step, n_rows = mdist.interleaved_rows(rank, world, height)
        local = torch.full((mdist.max_rows(world, height), width, 3), -1.0, dtype=torch.float64)
        for k in range(n_rows):
            y = y_first + k * y_step
            local[k, :, 0] = y
            local[k, :, 1] = torch.arange(width, dtype=torch.float64)
            local[k, :, 2] = rank
        frame = mdist.gather_frame(local, height, world)
        ok = frame.shape == (height, width, 3)
        ok = ok and bool(torch.all(frame[:, 0, 0] == torch.arange(height, dtype=torch.float64)))
        ok = ok and bool(torch.all(frame[:, :, 1] == torch.arange(width, dtype=torch.float64)[None, :]))
        ok = ok and bool(torch.all(frame[:, 0, 2] == torch.arange(height, dtype=torch.float64) % world))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("height", [54, 8])
def test_interleaved_gather_world2(height):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), height, 16, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
