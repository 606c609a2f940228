"""BASELINE config 1: `Task 1/Main_model.m` through the reference-named drop-ins (examples/main_model_task1.py):
loop-back BER = 0 (SURVEY KAT 3), pilot layout of `Main_model.m:14-24`, PAPR against the oracle, and the AWGN
variant's MER against the power budget (noise spreads over Nfft bins, the boosted pilots take 71 % of the power)."""
import importlib.util
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _example():
    spec = importlib.util.spec_from_file_location("main_model_task1", os.path.join(ROOT, "examples", "main_model_task1.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_task1_main_model_loop_back():
    m = _example()
    r = m.run()
    assert r["ok"] and r["BER"] == 0.0 and r["pilots"] == 101 and r["n_bits"] == 59800        # Np 101, Nd 299, 59,800 payload bits
    assert r["picture"].shape == (360, 360)
    # PAPR of the same stream from the oracle
    bits = (np.random.default_rng(1).random(59800) < 0.337).astype(np.uint8)
    pil, dat = O.pilot_layout_percent(400, 25, 1024, last_gap=2)
    iq, _ = O.mapping(bits, "16QAM")
    amp = 2 * np.max(np.abs(O.constellation_func("16QAM")[0]))
    tx = O.OFDM_modulator(O.OFDM_map_carriers_v1(iq, 50, 1024, dat, pil, amp), 128).ravel(order="F")
    assert abs(r["PAPR_dB"] - O.calculatePAPR(tx)) < 1e-3


def test_task1_awgn_mer():
    m = _example()
    r = m.run(SNR_dB=25.0, seed=3)
    amp2 = (2 * np.max(np.abs(O.constellation_func("16QAM")[0]))) ** 2          # pilot power, data power is 1
    expected = 25.0 + 10 * np.log10(1024 / (299 + 101 * amp2))                    # `Noise.m:3-5` measures the whole stream
    assert r["BER"] < 1e-3 and abs(r["MER_dB"] - expected) < 1.0
