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


def _photon_worker(rank, world, port, total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        mdist = importlib.import_module("monte-carlo-ray-tracer_b200.distributed")
        first, count = mdist.emission_range(total, rank, world)
        # stand-in for the emission kernels: emission w stores (w % 3) photons of 8 floats tagged with w
        mine = torch.tensor([float(w) for w in range(first, first + count) for _ in range((w % 3) * 8)], dtype=torch.float32)
        everything = mdist.all_gather_photons(mine, world)
        expect = torch.tensor([float(w) for w in range(total) for _ in range((w % 3) * 8)], dtype=torch.float32)
        handles = mdist.exchange_handles(bytes([rank] * 64), world)
        ret[rank] = (bool(torch.equal(everything, expect)) and handles == [bytes([r] * 64) for r in range(world)], first, count)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7, 1])
def test_emission_ranges_and_photon_gather_world2(total):
    """Photon pass sharded over ranks (photon-mapper.cpp:61-78 emission index space): the ranges tile [0, total), the
    variable-length photon arrays concatenate in rank order = emission order, and the IPC-handle exchange of the
    peer frames returns every rank's handle in rank order."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_photon_worker, args=(world, _free_port(), total, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world)), dict(ret)
    assert ret[0][1] == 0 and ret[0][1] + ret[0][2] == ret[1][1] and ret[1][1] + ret[1][2] == total


def test_emission_range_partitions():
    mdist = importlib.import_module("monte-carlo-ray-tracer_b200.distributed")
    for total in (0, 1, 5, 1000003):
        for world in (1, 2, 3, 8):
            spans = [mdist.emission_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
