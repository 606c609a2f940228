"""Monte-Carlo BER-vs-SNR sweep sharded over ranks (loop shape of
`Task 3/Main_model_Task_3.m:192-268`, `Task 5/Main_model_Task_5.m:303-346`, and the Task-4 impaired channel
`Task 4/Main_model_Task_4.m:95-110,277-366` swept over SNR).

The compute is ONE C-ABI call per rank, ``ofdm_sweep_ber`` (csrc/sweep.cu): global stream g = snr_index *
streams_per_point + j; rank r of R takes the contiguous share [T r / R, T (r+1) / R) of the T = n_snr * streams_per_point
streams (equal to within one stream for any R) -- **no data-path collective**; the only exchange is one
``all_reduce(SUM)`` of the int64 counters at the end (NCCL on GPUs; the same host logic runs over ``gloo`` in the CPU
tests with a stand-in compute function).  Payload bits, noise, STO and CFO draws are Philox streams keyed by g, so the
counters do not depend on the number of ranks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi

CHAINS = {"task5": 0, "task4": 1}


def share(total_streams, rank, world):
    """(first, count) of the contiguous share of the global stream range -- the rule of ``ofdm_sweep_share``."""
    lo, hi = total_streams * rank // world, total_streams * (rank + 1) // world
    return lo, hi - lo


def tiles(n_snr, streams_per_point, tile, rank=0, world=1):
    """The work items of one rank: (snr_index, first_stream_within_point, n_streams), never straddling an SNR point."""
    first, count = share(n_snr * streams_per_point, rank, world)
    out, g, end = [], first, first + count
    while g < end:
        i = g // streams_per_point
        n = min(tile, end - g, (i + 1) * streams_per_point - g)
        out.append((i, g - i * streams_per_point, n))
        g += n
    return out


def reduce_counts(counts, device=None):
    """Sum an int64 counter tensor over all ranks (no-op without an initialised process group)."""
    t = counts if isinstance(counts, torch.Tensor) else torch.as_tensor(np.asarray(counts), dtype=torch.int64)
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def run_sweep(snrs_db, streams_per_point, tile, compute, rank=0, world=1, device=None, n_counters=4):
    """Generic host driver (used with a stand-in ``compute`` in the gloo tests).
    ``compute(snr_index, snr_db, first_stream, n_streams) -> n_counters ints`` processes one tile on this rank; returns an
    int64 array (n_snr, n_counters) identical on every rank."""
    snrs_db = list(snrs_db)
    local = np.zeros((len(snrs_db), n_counters), dtype=np.int64)
    for (i, s0, n) in tiles(len(snrs_db), streams_per_point, tile, rank, world):
        local[i] += np.asarray(compute(i, snrs_db[i], s0, n), dtype=np.int64)
    return reduce_counts(torch.from_numpy(local), device).cpu().numpy()


def sweep_local(ctx, lp, snrs_db, streams_per_point, taps=None, chain="task5", seed=1, rank=0, world=1, tile=0, near_eps=0.0,
                sto_max=None, cfo_int_max=30, counts=None):
    """This rank's share of the sweep through ``ofdm_sweep_ber``; returns the device counters (n_snr, 4) int64
    {errors, bits, near-boundary symbols, detector failures} WITHOUT synchronising (everything is stream-ordered)."""
    snr = np.ascontiguousarray(np.asarray(snrs_db, dtype=np.float64))
    sp = _cabi.SweepParams()
    sp.chain = CHAINS[chain]
    sp.n_snr = snr.size
    sp.snr_db_host = snr.ctypes.data_as(_cabi.pdbl)
    sp.streams_per_point = int(streams_per_point)
    sp.tile_streams = int(tile)
    sp.rank, sp.world = int(rank), int(world)
    sp.seed = int(seed)
    t = None
    if taps is not None:
        t = np.ascontiguousarray(np.asarray(taps, dtype=np.float64).reshape(-1, 2))
        sp.taps_host = t.ctypes.data_as(_cabi.pdbl)
        sp.n_taps = t.shape[0]
    sp.near_eps = float(near_eps)
    sp.sto_max = int(lp.Nfft + lp.Tg if sto_max is None else sto_max)
    sp.cfo_int_max = int(cfo_int_max)
    if counts is None:
        counts = torch.zeros((snr.size, 4), dtype=torch.int64, device=ctx.device)
    ctx._chk(ctx.lib.ofdm_sweep_ber(ctx.h, C.byref(lp), C.byref(sp), ctx.p(counts)))
    return counts


def ber_sweep(ctx, lp, snrs_db, streams_per_point, taps=None, chain="task5", seed=1, rank=0, world=1, tile=0, near_eps=0.0, **kw):
    """Whole sweep: this rank's share on its GPU, then the closing all-reduce.  Returns (n_snr, 4) int64 on the host,
    identical on every rank and for every world size."""
    acc = sweep_local(ctx, lp, snrs_db, streams_per_point, taps, chain, seed, rank, world, tile, near_eps, **kw)
    if world > 1:
        reduce_counts(acc)                      # NCCL all-reduce on the device tensor, ordered after the sweep's stream work
    ctx.sync()
    return acc.cpu().numpy()
