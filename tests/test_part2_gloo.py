"""Sharding logic of the Task-5 part-2 sweep on CPU (world size 2, gloo) with a stand-in for the GPU point
function: every (point, run, estimator) cell is produced by exactly one rank and the result does not depend
on the rank count."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import ofdm_b200  # noqa: F401
from ofdm_b200 import part2

COMBS, RUNS, BLOCK = [4, 9, 100, 256], 10, 4


def fake_point(ctx, pc, dc, bits, n_runs, point_id=0, first_run=0, Ldict=None, **kw):
    r = first_run + np.arange(n_runs)
    nmse = np.stack([(point_id + 1) * 1e-3 * (j + 1) + 1e-6 * r for j in range(4)], axis=1)
    errs = np.array([int(np.sum((r * 7 + point_id + j) % 5)) for j in range(4)], dtype=np.int64)
    return nmse, errs, len(bits), {}


def payload(n):
    return np.zeros(n, dtype=np.uint8)


def test_combs_match_reference_rule():
    combs, amounts = part2.part2_combs()
    assert len(combs) == 57 and combs[0] == 4 and amounts[0] == 256 and amounts[-1] == 4       # 57 distinct pilot counts
    assert len(set(amounts)) == 57 and np.all(np.diff(combs) > 0)
    pc, dc = part2.layout(1024, comb=4)
    assert pc[0] == 1 and pc[-1] == 1021 and len(pc) == 256 and len(dc) == 768


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        NMSE, BER = part2.sweep(None, COMBS, RUNS, payload, block=BLOCK, rank=rank, world=world, run_fn=fake_point)
        np.save(os.path.join(out_dir, f"nmse{rank}.npy"), NMSE)
        np.save(os.path.join(out_dir, f"ber{rank}.npy"), BER)
    finally:
        dist.destroy_process_group()


def test_part2_sweep_independent_of_rank_count(tmp_path):
    N1, B1 = part2.sweep(None, COMBS, RUNS, payload, block=BLOCK, run_fn=fake_point)
    assert N1.shape == (4, len(COMBS))
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"nmse{r}.npy"), N1)
        assert np.array_equal(np.load(tmp_path / f"ber{r}.npy"), B1)
