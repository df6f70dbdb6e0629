"""Seeded synthetic inputs shared by tests and bench.py.

Signal model follows the reference's in-tree generator
(src/algorithms/signal_generator/gnuradio_blocks/signal_generator_c.cc:348-383):
sum over SVs of A * code(t - tau) * exp(j 2 pi f_d t), plus complex AWGN with sigma = 1.
"""
from __future__ import annotations

import numpy as np

GPS_L1_FREQ = 1575.42e6
CA_RATE = 1.023e6


def ca_amplitude(cn0_dbhz: float, fs: float) -> float:
    """Amplitude of a unit-modulus code so that C/N0 holds against complex noise of variance 2 (sigma=1 per rail)."""
    return float(np.sqrt(2.0 * 10 ** (cn0_dbhz / 10.0) / fs))


def make_iq(codes: dict, fs: float, n: int, svs: list, seed: int, noise: bool = True, chips_per_table_chip: float = 1.0):
    """codes: prn -> float32 table (one period); svs: list of dict(prn, doppler, code_phase_chips, cn0, phase0).
    Returns complex64[n]."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    x = np.zeros(n, np.complex128)
    for sv in svs:
        tbl = codes[sv["prn"]]
        L = len(tbl)
        rate = CA_RATE * chips_per_table_chip * (1.0 + sv["doppler"] / GPS_L1_FREQ)
        idx = np.floor(sv["code_phase_chips"] + t * (rate / fs)).astype(np.int64) % L
        amp = ca_amplitude(sv.get("cn0", 45.0), fs)
        x += amp * tbl[idx] * np.exp(1j * (sv.get("phase0", 0.0) + 2 * np.pi * sv["doppler"] * t / fs))
    if noise:
        x += rng.standard_normal(n) + 1j * rng.standard_normal(n)
    return x.astype(np.complex64)


def trk_params_for(sv: dict, fs: float, epoch_len: int, n_epochs: int, table_chips_per_chip: float = 1.0, L: int = 1023):
    """Open-loop per-epoch correlator parameters a perfectly locked DLL/PLL would pass to
    do_correlation_step (dll_pll_veml_tracking.cc:1232-1257, update_tracking_vars :1409-1483):
    returns arrays (sample_index, rem_carr, phase_step, rem_code, code_step) of length n_epochs."""
    k = np.arange(n_epochs, dtype=np.float64)
    s = k * epoch_len
    dphi = 2 * np.pi * sv["doppler"] / fs
    rem_carr = np.mod(sv.get("phase0", 0.0) + dphi * s, 2 * np.pi)
    step = CA_RATE * table_chips_per_chip * (1.0 + sv["doppler"] / GPS_L1_FREQ) / fs
    code_phase = np.mod(sv["code_phase_chips"] + step * s, L)
    # the resampler evaluates floor(step*n + shift - rem): rem is minus the code phase at the epoch start
    rem_code = -code_phase
    return (s.astype(np.uint64), rem_carr.astype(np.float32), np.full(n_epochs, dphi, np.float32),
            rem_code.astype(np.float32), np.full(n_epochs, step, np.float32))
