"""Pilot / data carrier layouts and pilot values of the reference scripts, for building ``ofdm_link_params`` without
touching the test oracle (host-side index arithmetic only; 1-based carrier numbers as in MATLAB)."""
from __future__ import annotations

import numpy as np

from .link import constellation_func

TAPS_TASK5 = [[0, 1], [4, .8], [10, .6], [15, .4], [21, .2], [25, .1]]     # `Task 5/Main_model_Task_5.m:112-119`
TAPS_TASK4 = [[0, 1], [4, .6], [10, .3]]                                   # `Task 4/Main_model_Task_4.m:252-256`


def pilot_layout_comb(N_carrier, comb):
    """`Task 5/Main_model_Task_5.m:18-22,34`: ``pilots = 1:comb:N_carrier``; data = the other carriers."""
    pilots = np.arange(1, N_carrier + 1, int(comb), dtype=np.int64)
    return pilots, np.setdiff1d(np.arange(1, N_carrier + 1, dtype=np.int64), pilots)


def pilot_layout_percent(N_carrier, percent, last_gap=2):
    """`Task 4/Main_model_Task_4.m:14-24`: ``pilots = [1:step:N_carrier-last_gap, N_carrier]`` with
    ``step = floor(N_carrier / round(percent/100*N_carrier))``."""
    amount = int(np.floor(percent / 100 * N_carrier + 0.5))
    step = N_carrier // amount
    pilots = np.unique(np.concatenate([np.arange(1, N_carrier - last_gap + 1, step), [N_carrier]])).astype(np.int64)
    return pilots, np.setdiff1d(np.arange(1, N_carrier + 1, dtype=np.int64), pilots)


def pilot_values(Np, N_symb, Constellation, scale, alternate):
    """`Task 4/Main_model_Task_4.m:31-36`: amplitude ``scale*max|dict|``, phases 0 / pi alternating, the ctranspose of
    ``repmat(pilotValues', 1, N_symb)`` kept (so "-a" carries an imaginary part of about -a*1.2e-16)."""
    d, _ = constellation_func(Constellation)
    amp = scale * np.max(np.abs(d))
    pv = np.full(Np, amp * np.exp(1j * 0), dtype=np.complex128)
    if alternate:
        pv[1::2] = amp * np.exp(1j * np.pi)
    return np.tile(np.conj(pv)[:, None], (1, N_symb))


def task5_link(ctx, comb=4, scale=2.0, alternate=True, Constellation="16QAM", scramble=True):
    """Task-5 part-2 / M1 shape (`Task 5/Task5_part2.m:5-17,46-91`): Nfft 4096, CP 512, 1024 carriers, 2 x 7 symbols."""
    pil, dat = pilot_layout_comb(1024, comb)
    return ctx.link_params(4096, 512, 1024, 14, 7, Constellation, dat, pil, pilot_values(len(pil), 14, Constellation, scale, alternate),
                           scramble=scramble)


def task4_link(ctx, percent=15, scale=4.0 / 3.0, alternate=True, Constellation="16QAM", scramble=True):
    """Task-4 shape (`Task 4/Main_model_Task_4.m:6-36`): Nfft 1024, CP 128, 400 carriers, 10 x 5 symbols."""
    pil, dat = pilot_layout_percent(400, percent)
    return ctx.link_params(1024, 128, 400, 50, 5, Constellation, dat, pil, pilot_values(len(pil), 50, Constellation, scale, alternate),
                           scramble=scramble)
