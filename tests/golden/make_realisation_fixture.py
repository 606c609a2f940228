"""Generates tests/golden/realisation_small.mat: one stream's shared realisation (payload bits, AWGN normals,
channel taps, SNR) in the MAT-file layout of ofdm_b200.realisations, together with the oracle's outputs on it
(ref_*).  A MATLAB owner produces the same file from the untouched reference with the recipe in INTEGRATION.md
("Shared realisations"); dropping that file here pins both the oracle and the GPU path to MATLAB's numbers.

    python tests/golden/make_realisation_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ofdm_b200  # noqa: E402,F401
from ofdm_b200 import realisations as R  # noqa: E402
from oracle import chains as OC  # noqa: E402
from oracle import functions as F  # noqa: E402


def small_params():
    p = OC.LinkParams(Nfft=512, N_carrier=128, T_Guard=64, Amount_OFDM_Frames=2, Amount_ODFM_SpF=2)
    p.pilotCarriers, p.dataCarriers = F.pilot_layout_comb(p.N_carrier, 4)
    p.pilotValues, _ = OC.make_pilot_values(len(p.pilotCarriers), p.N_symb, p.Constellation, 2.0, True)
    return p


TAPS = [[0, 1], [4, .8], [10, .6]]
SNR_DB = 9.0

if __name__ == "__main__":
    p = small_params()
    rng = np.random.default_rng(20261018)
    bits = rng.integers(0, 2, p.stream_bits).astype(np.uint8)
    normals = rng.standard_normal((2, p.stream_len))
    tx, _, _ = OC.tx_chain(p, bits)
    rx = OC.channel_task5(p, tx, SNR_DB, TAPS, normals=normals)
    ref = OC.rx_chain_task5(p, rx, bits)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "realisation_small.mat")
    R.export_realisation(out, input_bits=bits, noise_normals=normals.T, channel_taps=np.asarray(TAPS, dtype=np.float64), SNR_dB=SNR_DB,
                         ref_tx=tx, ref_rx=rx, ref_H_LS=ref["H"], ref_bits=ref["bits"].astype(np.float64), ref_errors=float(ref["errors"]))
    print(out, os.path.getsize(out), "bytes; errors", ref["errors"], "of", ref["n_bits"])
