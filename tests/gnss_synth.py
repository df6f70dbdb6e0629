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
    """codes: prn -> float32 table (one period); svs: list of dict(prn, doppler, code_phase_chips, cn0, phase0
    [, symbols, periods_per_symbol]): `symbols` (+-1 array, cycled) modulates the code, one value every
    `periods_per_symbol` code periods (navigation bits, secondary codes).  Returns complex64[n]."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    x = np.zeros(n, np.complex128)
    for sv in svs:
        tbl = codes[sv["prn"]]
        L = len(tbl)
        rate = CA_RATE * chips_per_table_chip * (1.0 + sv["doppler"] / GPS_L1_FREQ)
        chip = np.floor(sv["code_phase_chips"] + t * (rate / fs)).astype(np.int64)
        idx = chip % L
        amp = ca_amplitude(sv.get("cn0", 45.0), fs)
        sig = amp * tbl[idx] * np.exp(1j * (sv.get("phase0", 0.0) + 2 * np.pi * sv["doppler"] * t / fs))
        if sv.get("symbols") is not None:
            sym = np.asarray(sv["symbols"], np.float64)
            period = chip // L  # code period number since t = 0 (may start mid-period: symbol edges follow the code)
            sig = sig * sym[(period // int(sv.get("periods_per_symbol", 1))) % len(sym)]
        x += sig
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


# ---- GPS L1 C/A code tables for synthetic inputs (independent of oracle/: bench.py and tools/ use these) ----------
# IS-GPS-200 Table 3-Ia: G2 output = XOR of two stages (code phase selection) for PRN 1..32.
_G2_TAPS = [(2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10), (2, 3), (3, 4), (5, 6), (6, 7), (7, 8), (8, 9),
            (9, 10), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6), (5, 7), (6, 8), (7, 9), (8, 10), (1, 6), (2, 7),
            (3, 8), (4, 9)]


def gps_ca_code(prn: int) -> np.ndarray:
    """1023 chips of the C/A code of PRN 1..32 as float32 +-1 (chip value 1 -> +1, 0 -> -1), the table the reference's
    gps_l1_ca_code_gen_float produces (checked against the oracle port in tests/test_codes.py)."""
    g1 = np.ones(10, np.uint8)
    g2 = np.ones(10, np.uint8)
    s1, s2 = _G2_TAPS[prn - 1]
    out = np.empty(1023, np.float32)
    for k in range(1023):
        chip = g1[9] ^ g2[s1 - 1] ^ g2[s2 - 1]
        out[k] = 1.0 if chip else -1.0
        f1 = g1[2] ^ g1[9]
        f2 = g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9]
        g1[1:] = g1[:-1]
        g1[0] = f1
        g2[1:] = g2[:-1]
        g2[0] = f2
    return out


def gps_ca_code_complex_sampled(prn: int, fs: int) -> np.ndarray:
    """One code period sampled at fs as (0, +-1) complex64, as gps_l1_ca_code_gen_complex_sampled does
    (float32 index arithmetic, last sample pinned to the last chip)."""
    code = gps_ca_code(prn)
    n = int(float(fs) / (1023000.0 / 1023.0))
    tc = np.float32(1.0) / np.float32(1023000)
    ts = np.float32(1.0) / np.float32(fs)
    i = np.arange(n, dtype=np.float32)
    idx = np.floor((ts * i) / tc).astype(np.int64)
    idx[n - 1] = 1022
    return (1j * code[idx]).astype(np.complex64)
