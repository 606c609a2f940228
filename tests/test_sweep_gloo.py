"""Multi-rank path on CPU: world-size-2 `gloo` run of the sweep's host logic (work partition + the single
all_reduce of int64 counters) with a stand-in compute function -- the product has no CPU compute path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ofdm_b200
from ofdm_b200 import sweep

SNRS = [0.0, 2.5, 5.0, 7.5, 10.0]
SPP, BLOCK = 37, 8        # ragged: 5 blocks per point, the last one of 5 streams


def fake_compute(i, snr_db, s0, n):
    """Deterministic stand-in keyed by global stream ids only (like the Philox-keyed GPU path)."""
    gids = i * SPP + s0 + np.arange(n)
    errors = int(np.sum((gids * 2654435761) % 97 < (20 - snr_db)))
    return errors, n * 43008, int(np.sum(gids % 11 == 0)), int(np.sum(gids % 29 == 0))


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
        res = sweep.run_sweep(SNRS, SPP, BLOCK, fake_compute, rank, world)
        mine = sweep.tiles(len(SNRS), SPP, BLOCK, rank, world)
        np.save(os.path.join(out_dir, f"res{rank}.npy"), res)
        np.save(os.path.join(out_dir, f"n{rank}.npy"), np.array([len(mine)]))
    finally:
        dist.destroy_process_group()


def test_shares_cover_every_stream_once_and_are_balanced():
    total = len(SNRS) * SPP
    for world in (1, 2, 3, 8):
        seen = np.zeros((len(SNRS), SPP), dtype=int)
        sizes = []
        for r in range(world):
            first, count = sweep.share(total, r, world)
            sizes.append(count)
            items = sweep.tiles(len(SNRS), SPP, BLOCK, r, world)
            assert sum(n for _, _, n in items) == count
            for i, s0, n in items:
                assert n <= BLOCK and s0 + n <= SPP            # a tile never straddles an SNR point
                seen[i, s0:s0 + n] += 1
        assert np.all(seen == 1)
        assert max(sizes) - min(sizes) <= 1                    # equal to within one stream for any world size


def test_share_rule_matches_the_library():
    import ctypes as C
    from ofdm_b200 import _cabi
    lib = _cabi.load()
    for total, world in ((499712, 8), (61 * 8192, 3), (17, 5), (0, 2)):
        for r in range(world):
            f, c = C.c_int64(0), C.c_int64(0)
            assert lib.ofdm_sweep_share(total, r, world, C.byref(f), C.byref(c)) == 0
            assert (f.value, c.value) == sweep.share(total, r, world)


@pytest.mark.parametrize("world", [2, 3])
def test_sweep_counts_independent_of_rank_count(tmp_path, world):
    single = sweep.run_sweep(SNRS, SPP, BLOCK, fake_compute, 0, 1)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"res{r}.npy"), single)        # identical on every rank
    assert sum(int(np.load(tmp_path / f"n{r}.npy")[0]) for r in range(world)) >= len(sweep.tiles(len(SNRS), SPP, BLOCK))
    assert single[:, 1].sum() == len(SNRS) * SPP * 43008
