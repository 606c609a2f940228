"""Script-level glue of the reference, restated (test infrastructure; PARITY UNPINNED, see
``oracle/__init__.py``).  Each chain composes the per-function oracle in the order the
reference scripts call it, for ONE serial stream (``N_symb`` OFDM symbols); batches are
Python loops over streams.  Citations are ``path:line`` under ``/root/reference/``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import functions as F


@dataclass
class LinkParams:
    """Literals at the top of the reference scripts (`Task 5/Main_model_Task_5.m:6-46`)."""
    Nfft: int = 4096
    N_carrier: int = 1024
    T_Guard: int = 512
    Amount_OFDM_Frames: int = 2
    Amount_ODFM_SpF: int = 7
    Constellation: str = "16QAM"
    pilotCarriers: np.ndarray = field(default=None)      # 1-based
    dataCarriers: np.ndarray = field(default=None)       # 1-based
    pilotValues: np.ndarray = field(default=None)        # (Np, N_symb)
    Register: np.ndarray = field(default_factory=lambda: F.DEFAULT_REGISTER.copy())

    @property
    def N_symb(self):
        return self.Amount_OFDM_Frames * self.Amount_ODFM_SpF

    @property
    def bps(self):
        return F.constellation_func(self.Constellation)[1]

    @property
    def frame_bits(self):
        return self.Amount_ODFM_SpF * len(self.dataCarriers) * self.bps

    @property
    def stream_bits(self):
        return self.frame_bits * self.Amount_OFDM_Frames

    @property
    def stream_len(self):
        return self.N_symb * (self.Nfft + self.T_Guard)


def make_pilot_values(Np, N_symb, Constellation, scale, alternate):
    """`Task 4/Main_model_Task_4.m:31-36` / `Task 5/Task5_part2.m:85-91`: amplitude
    ``scale*max|dict|``; ``exp(1i*pi)`` and the ctranspose in ``repmat(pilotValues',...)`` are
    kept, so "-a" carries an imaginary part of about -a*1.2e-16."""
    d, _ = F.constellation_func(Constellation)
    amp = scale * np.max(np.abs(d))
    pv = np.zeros(Np, dtype=np.complex128)
    pv[:] = amp * np.exp(1j * 0)
    if alternate:
        pv[1::2] = amp * np.exp(1j * np.pi)
    return np.tile(np.conj(pv)[:, None], (1, N_symb)), amp


def params_task5(comb=4, scale=2.0, alternate=True):
    """M1 / Task-5 part-2 shape (`Task 5/Task5_part2.m:5-17,46-91`)."""
    p = LinkParams()
    p.pilotCarriers, p.dataCarriers = F.pilot_layout_comb(p.N_carrier, comb)
    p.pilotValues, _ = make_pilot_values(len(p.pilotCarriers), p.N_symb, p.Constellation, scale, alternate)
    return p


def params_task4(percent=15, scale=4.0 / 3.0, alternate=True, Constellation="16QAM"):
    """Task-4 shape (`Task 4/Main_model_Task_4.m:6-36`)."""
    p = LinkParams(Nfft=1024, N_carrier=400, T_Guard=128, Amount_OFDM_Frames=10, Amount_ODFM_SpF=5,
                   Constellation=Constellation)
    p.pilotCarriers, p.dataCarriers = F.pilot_layout_percent(p.N_carrier, percent, p.Nfft, last_gap=2)
    p.pilotValues, _ = make_pilot_values(len(p.pilotCarriers), p.N_symb, p.Constellation, scale, alternate)
    return p


def scramble_frames(p: LinkParams, bits, descramble=False, fast=False):
    """Per-frame register reset (`Task 4/Main_model_Task_4.m:43-58`, `:350-364`).  ``fast`` selects the
    vectorised (bit-identical) forms, used only where the oracle is timed as the CPU baseline."""
    if fast:
        fn = F.DeScrambler_fast if descramble else F.Scrambler_fast
    else:
        fn = F.DeScrambler if descramble else F.Scrambler
    L = p.frame_bits
    out = np.empty(p.stream_bits, dtype=np.uint8)
    for i in range(p.Amount_OFDM_Frames):
        out[i * L:(i + 1) * L], _ = fn(p.Register, bits[i * L:(i + 1) * L])
    return out


def tx_chain(p: LinkParams, input_bits, scramble=True, fast=False):
    """bits -> Scrambler -> mapping -> OFDM_map_carriers -> OFDM_modulator -> serial stream
    (`Task 5/Main_model_Task_5.m:53-85`).  Returns (stream, grid, sc_bits)."""
    bits = np.asarray(input_bits).ravel()
    sc = scramble_frames(p, bits, fast=fast) if scramble else bits.astype(np.uint8)
    iq, pad = F.mapping(sc, p.Constellation)
    grid = F.OFDM_map_carriers(iq, p.N_symb, p.Nfft, p.dataCarriers, p.pilotCarriers, p.pilotValues)
    tx = F.OFDM_modulator(grid, p.T_Guard)
    return tx.ravel(order="F"), grid, sc


def channel_task5(p: LinkParams, tx, SNR_dB, taps, normals=None, rng=None):
    """Noise first, then multipath (`Task 5/Main_model_Task_5.m:108,123-127`)."""
    rx = tx
    if SNR_dB is not None:
        rx, _ = F.Noise(SNR_dB, rx, normals=normals, rng=rng)
    if taps is not None:
        h, _ = F.get_MP_channel_resp(taps, p.Nfft)
        rx = F.apply_channel(rx, h)
    return rx


def rx_chain_task5(p: LinkParams, rx_stream, input_bits, descramble=True, method="LS", h_true=None, SNR_dB=None, fast=False):
    """M1 chain: OFDM_demodulator -> LS_CE -> equalize_signal -> get_payload -> demapping ->
    DeScrambler -> BER_func (`Task 5/Task5_part2.m:169-174,269-303`, per-frame descrambler
    reset as `Task 5/Main_model_Task_5.m:262-271`).  Returns dict."""
    X = np.asarray(rx_stream).reshape((p.Nfft + p.T_Guard, p.N_symb), order="F")
    Y = F.OFDM_demodulator(X, p.T_Guard)
    if method == "LS":
        H = F.LS_CE(Y, p.pilotValues, p.pilotCarriers, p.N_carrier)
    elif method == "MMSE":
        h = h_true if h_true is not None else np.fft.ifft(F.LS_CE(Y, p.pilotValues, p.pilotCarriers, p.N_carrier))
        H = F.MMSE_CE(Y, p.pilotValues, p.pilotCarriers, p.Nfft, p.N_carrier, h, SNR_dB)
    else:
        raise ValueError(method)
    eq = F.equalize_signal(Y, H, p.N_carrier)
    rx_iq = F.get_payload(eq, p.dataCarriers).ravel(order="F")
    out_bits = F.demapping(-1, rx_iq, p.Constellation)
    dsc = scramble_frames(p, out_bits, descramble=True, fast=fast) if descramble else out_bits
    tx_bits = np.asarray(input_bits).ravel()
    n_err = int(np.sum(tx_bits != dsc))
    return {"Y": Y, "H": H, "eq": eq, "rx_iq": rx_iq, "bits": dsc, "raw_bits": out_bits,
            "errors": n_err, "n_bits": tx_bits.size}


def impair_task4(p: LinkParams, tx, SNR_dB=None, Time_Delay=None, Freq_Shift=None, taps=None, normals=None, rng=None):
    """Impairments in Task-4 order: Noise -> add_STO -> add_CFO -> multipath
    (`Task 4/Main_model_Task_4.m:95,103,110,263-264`)."""
    rx = np.asarray(tx)
    if SNR_dB is not None:
        rx, _ = F.Noise(SNR_dB, rx, normals=normals, rng=rng)
    if Time_Delay is not None:
        rx = F.add_STO(rx, Time_Delay)
    if Freq_Shift is not None:
        rx = F.add_CFO(rx, Freq_Shift, p.Nfft)
    if taps is not None:
        h, _ = F.get_MP_channel_resp(taps, p.Nfft)
        rx = F.apply_channel(rx, h)
    return rx


def rx_chain_task4(p: LinkParams, rx_stream, input_bits, time_desync=True, freq_desync=True, mp_desync=True):
    """M2 chain (`Task 4/Main_model_Task_4.m:277-366`): AutoCorrFunction -> add_STO x2 -> add_CFO ->
    remove_IFO -> reshape -> OFDM_demodulator -> fine_sync -> estimate_channel -> equalize_signal ->
    get_payload -> demapping -> DeScrambler -> BER."""
    rx = np.asarray(rx_stream).ravel()
    info = {}
    if time_desync or freq_desync:
        ac, tg, fo = F.AutoCorrFunction(rx, p.T_Guard, p.Nfft)
        info.update(TgPosition=tg, FreqOffset=fo)
        if time_desync:
            rx = F.add_STO(rx, tg)
            rx = F.add_STO(rx, -(p.Nfft + p.T_Guard))
    if freq_desync:
        rx = F.add_CFO(rx, -fo, p.Nfft)
        rx, ifo = F.remove_IFO(rx, p.Nfft)
        info.update(IFO=ifo)
    X = rx.reshape((p.Nfft + p.T_Guard, p.N_symb), order="F")
    Y = F.OFDM_demodulator(X, p.T_Guard)
    if time_desync or freq_desync:
        Y, tau, ph = F.fine_sync(Y, p.pilotCarriers, p.pilotValues, time_desync, freq_desync, return_estimates=True)
        info.update(tau=tau, phase_shift=ph)
    if mp_desync:
        H, Hp = F.estimate_channel(Y, np.arange(1, p.Nfft + 1), p.pilotCarriers, p.pilotValues)
        info.update(H=H)
        Y = F.equalize_signal(Y, H, p.N_carrier)
    rx_iq = F.get_payload(Y, p.dataCarriers).ravel(order="F")
    out_bits = F.demapping(-1, rx_iq, p.Constellation)
    dsc = scramble_frames(p, out_bits, descramble=True)
    tx_bits = np.asarray(input_bits).ravel()
    info.update(Y=Y, rx_iq=rx_iq, bits=dsc, errors=int(np.sum(tx_bits != dsc)), n_bits=tx_bits.size)
    return info


def read_payload_bits(tiff_path, size_buffer):
    """`Task 5/file_reader.m:2-12`: ``imbinarize`` = Otsu threshold on the 256-bin histogram,
    column-major flatten, truncate."""
    from PIL import Image
    img = np.asarray(Image.open(tiff_path)).astype(np.float64)
    counts = np.bincount(img.astype(np.int64).ravel(), minlength=256).astype(np.float64)
    p = counts / counts.sum()
    omega = np.cumsum(p)
    mu = np.cumsum(p * np.arange(1, 257))
    mu_t = mu[-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        sigma_b2 = (mu_t * omega - mu) ** 2 / (omega * (1 - omega))
    sigma_b2[~np.isfinite(sigma_b2)] = -np.inf
    mx = sigma_b2.max()
    level = (np.mean(np.nonzero(sigma_b2 == mx)[0]) ) / 255.0     # graythresh: mean of maxima, (idx-1)/(nbins-1)
    bw = (img / 255.0) > level
    return bw.ravel(order="F")[:size_buffer].astype(np.uint8)
