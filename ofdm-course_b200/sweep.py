"""Monte-Carlo BER-vs-SNR sweep sharded over ranks (loop shape of
`Task 3/Main_model_Task_3.m:192-268` and `Task 5/Main_model_Task_5.m:303-346`).

Streams and SNR points are independent, so the work list -- (SNR point, block of streams) pairs -- is dealt
round-robin to the ranks with **no data-path collective**; the only exchange is one ``all_reduce(SUM)`` of the
int64 counters at the end (NCCL on GPUs; the same code runs over ``gloo`` in the CPU tests with a stand-in
compute function).  Noise is Philox keyed by the *global* stream id, so the counts do not depend on the number
of ranks.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def work_list(n_snr, streams_per_point, block):
    """All (snr_index, first_stream, n_streams) items, in a fixed global order."""
    items = []
    for i in range(n_snr):
        for s0 in range(0, streams_per_point, block):
            items.append((i, s0, min(block, streams_per_point - s0)))
    return items


def my_items(items, rank, world):
    """Round-robin deal: rank r takes items r, r+world, ..."""
    return items[rank::world]


def reduce_counts(counts, device=None):
    """Sum an int64 counter tensor over all ranks (no-op without an initialised process group)."""
    t = counts if isinstance(counts, torch.Tensor) else torch.as_tensor(np.asarray(counts), dtype=torch.int64)
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def run_sweep(snrs_db, streams_per_point, block, compute, rank=0, world=1, device=None):
    """Generic driver.  ``compute(snr_index, snr_db, first_stream, n_streams) -> (errors, bits, near)`` processes one
    work item on this rank; returns an int64 array (n_snr, 3) identical on every rank."""
    snrs_db = list(snrs_db)
    local = np.zeros((len(snrs_db), 3), dtype=np.int64)
    for (i, s0, n) in my_items(work_list(len(snrs_db), streams_per_point, block), rank, world):
        local[i] += np.asarray(compute(i, snrs_db[i], s0, n), dtype=np.int64)
    return reduce_counts(torch.from_numpy(local), device).cpu().numpy()


def ber_sweep_task5(ctx, lp, snrs_db, streams_per_point, block, taps_h, seed=1, rank=0, world=1, near_eps=0.0, sync_every_item=False):
    """Full TX -> AWGN + multipath -> RX sweep on this rank's GPU (`ctx`), one fused kernel per stage and item.
    Payload bits and noise are keyed by the global stream id, so any (rank, world) split gives the same counts."""
    h_dev = ctx.cplx(np.asarray(taps_h)) if taps_h is not None else None
    words = lp.stream_bits // 32
    snrs_db = list(snrs_db)
    n_snr = len(snrs_db)

    # Counters stay on the device (the RX kernel accumulates into the row of its SNR point) and are read once at the end:
    # no host synchronisation between work items, so the launches of item k+1 are queued while item k runs.
    acc = torch.zeros((n_snr, 3), dtype=torch.int64, device=ctx.device)
    for (i, s0, n) in my_items(work_list(n_snr, streams_per_point, block), rank, world):
        gid0 = i * streams_per_point + s0                      # global stream id of the item's first stream
        gen = torch.Generator(device=ctx.device)
        gen.manual_seed(seed * 1_000_003 + gid0)
        bits = torch.randint(-2**31, 2**31 - 1, (n * words,), dtype=torch.int32, device=ctx.device, generator=gen)
        tx, psum = ctx.tx_chain(lp, bits, n, want_power=True)       # stream power measured in the TX kernel's registers
        rx = ctx.channel_t5(tx, snr_db=float(snrs_db[i]), h_dev=h_dev, seed=seed, first_stream_id=gid0, power_sum=psum)
        del tx
        ctx.rx_chain_t5(lp, rx, n, tx_bits_dev=bits, want_bits=False, want_H=False, near_eps=near_eps, counts=acc[i])
        del rx, bits, psum                                     # same stream: the allocator hands these blocks to the next item
        if sync_every_item:
            ctx.sync()
    ctx.sync()
    return reduce_counts(acc.cpu(), ctx.device if world > 1 else None).cpu().numpy()
