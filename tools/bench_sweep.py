#!/usr/bin/env python
"""M5 (BASELINE config 5): full TX -> AWGN + multipath -> RX BER-vs-SNR Monte-Carlo sweep, sharded over the ranks
(one process per GPU under torch.distributed.run, or a single process), one closing all-reduce of the counters.

    python tools/bench_sweep.py [--streams-per-point 8192] [--block 8192] [--snr-step 0.5]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sweep.py

SNR 0:0.5:30 (61 points, `Task 3/Main_model_Task_3.m:192`), 16,384 frames = 8,192 streams x 14 symbols per point,
Task-5 part-2 shape (Nfft 4096, comb 4, 16QAM), channel of `Task 5/Main_model_Task_5.m:112-119`."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ofdm_b200 as G  # noqa: E402
from ofdm_b200 import sweep  # noqa: E402
from ofdm_b200 import layouts  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams-per-point", type=int, default=8192)
    ap.add_argument("--tile", type=int, default=2048)
    ap.add_argument("--chain", default="task5", choices=["task5", "task4"])
    ap.add_argument("--snr-step", type=float, default=0.5)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = G.Context(local, "f32")
    if a.chain == "task5":
        lp, taps = layouts.task5_link(ctx, comb=4), layouts.TAPS_TASK5
    else:
        lp, taps = layouts.task4_link(ctx), layouts.TAPS_TASK4
    snrs = np.arange(0.0, 30.0 + 1e-9, a.snr_step)
    # warm-up at the timed tile size, so that the signal buffers are already in the allocator's pool (a first cudaMalloc of
    # that size costs tens of milliseconds) and every kernel variant has been loaded
    sweep.ber_sweep(ctx, lp, snrs[:2], a.tile * world, taps, a.chain, rank=rank, world=world, tile=a.tile)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc = sweep.sweep_local(ctx, lp, snrs, a.streams_per_point, taps, a.chain, rank=rank, world=world, tile=a.tile)
    if world > 1:
        sweep.reduce_counts(acc)
    e1.record()
    torch.cuda.synchronize()
    res = acc.cpu().numpy()
    dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        syms = len(snrs) * a.streams_per_point * lp.S
        ber = res[:, 0] / np.maximum(res[:, 1], 1)
        print(json.dumps({"workload": "M5 (%s chain): full TX -> channel -> RX BER sweep, SNR 0:%g:30, %d streams x %d symbols per point" % (a.chain, a.snr_step, a.streams_per_point, lp.S),
                          "counters_sha1": __import__("hashlib").sha1(res.tobytes()).hexdigest(), "detector_failures": int(res[:, 3].sum()),
                          "n_gpus": world, "seconds": float(dt.item()), "symbols": syms, "symbols_per_s": syms / float(dt.item()),
                          "kernels_this_rank": int(ctx.launches - l0), "bits_per_point": int(res[0, 1]),
                          "ber_at_snr": {str(float(s)): float(b) for s, b in zip(snrs[::10], ber[::10])},
                          "ber_monotone_non_increasing": bool(np.all(np.diff(ber) <= 1e-4))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
