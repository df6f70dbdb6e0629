"""pcps_acquisition_fine_doppler_cc on the device (coarse grid via b200_acq, fine Doppler via b200_acq_fine) against
the numpy restatement oracle/acq_fine_np.py.  Contract as for pcps_acquisition (FFT boundary unpinned upstream):
indices exact, statistics within 1e-4 relative, fine spectrum within 1e-4 of its peak."""
import numpy as np
import pytest
import scipy.fft as sfft

import gnss_synth as gs
from oracle.acq_fine_np import FineDopplerOracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n1ms", [2000, 4000, 12500, 25000], ids=lambda n: f"N{n}")
def test_zero_padded_transform_matches_fft(n1ms):
    """Eight modulated (10 N)-point transforms == one 80 N-point transform of the zero-padded product: exercises the
    single-level (N=2000: 20000 points, radix-25 + permuted storage) and two-level plans (40000 = 2 x 20000,
    125000 = 5 x 25000, 250000 = 10 x 25000) and the storage -> frequency mapping."""
    from gnss_sdr_b200 import capi
    rng = np.random.default_rng(n1ms)
    m, ext = 10 * n1ms, 80 * n1ms
    k_true = 7 * 8 + 3                      # a tone that falls on sub-bin s = 3
    n = np.arange(m)
    buf = (np.exp(2j * np.pi * k_true * n / ext) + 0.05 * (rng.standard_normal(m) + 1j * rng.standard_normal(m))).astype(np.complex64)
    code = np.where(rng.integers(0, 2, n1ms) > 0, 1.0, -1.0).astype(np.float32) * 1j
    code = code.astype(np.complex64)
    x = np.zeros(ext, np.complex64)
    x[:m] = (buf * np.tile(code, 10) * np.tile(np.conj(code), 10)).astype(np.complex64)   # code wiped twice = tone again
    e = capi.Engine()
    f = capi.AcqFineDoppler(e, n1ms)
    # feed buffer * conj(code) so that the engine's own code multiplication restores the tone
    idx, peak = f.estimate((buf * np.tile(np.conj(code), 10)).astype(np.complex64), code)
    want = sfft.fft(x)
    want_mag = (want.real.astype(np.float64) ** 2 + want.imag.astype(np.float64) ** 2)
    assert idx == k_true == int(np.argmax(want_mag))
    got = f.read_spectrum().astype(np.float64)
    assert np.max(np.abs(got - want_mag)) < 1e-4 * want_mag.max()
    assert abs(peak - want_mag.max()) < 1e-4 * want_mag.max()
    f.close()
    e.close()


CASES = [
    # fs, doppler_max, doppler_step, dwells, threshold, sv doppler, delay, cn0
    (4e6, 5000, 250, 2, 1.5, 1100.0, 1234, 50.0),      # coarse Doppler mis-centred by the upstream grid offset: fine estimate rejected
    (4e6, 1000, 250, 1, 1.5, 600.0, 3001, 50.0),       # |fine - coarse| < 1 kHz: fine estimate accepted
    (4e6, 5000, 500, 3, 1.5, -2300.0, 17, 47.0),       # negative Doppler is outside the (offset) grid: weak peak
    (2e6, 5000, 250, 2, 2.5, 1400.0, 801, 30.0),       # below threshold: negative acquisition, no fine step
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"fs{int(c[0])}_dmax{c[1]}_f{int(c[5])}")
def test_fine_doppler_block_matches_oracle(oracle, case):
    from gnss_sdr_b200 import capi
    from gnss_sdr_b200.fine_doppler import PcpsAcquisitionFineDoppler
    fs, dmax, dstep, dwells, thr, doppler, delay, cn0 = case
    prn = 9
    n = int(fs / 1000)
    code = oracle.port.gps_ca_code(prn)
    codec = oracle.port.gps_ca_code_complex_sampled(prn, int(fs))
    iq = gs.make_iq({prn: code}, fs, n * 14, [dict(prn=prn, doppler=doppler, code_phase_chips=(-delay * 1.023e6 / fs) % 1023, cn0=cn0)], seed=int(doppler) % 97)
    o = FineDopplerOracle(int(fs), float(n), dmax, dstep, dwells, thr)
    o.set_local_code(codec)
    want_pos, want = o.run(iq, codec)
    e = capi.Engine()
    blk = PcpsAcquisitionFineDoppler(e, fs_in=int(fs), samples_per_ms=float(n), doppler_max=dmax, doppler_step=dstep, max_dwells=dwells, threshold=thr)
    blk.set_local_code(codec)
    got_pos, got = blk.run(iq)
    assert got_pos == want_pos
    assert got["index_time"] == want["index_time"] and got["index_doppler"] == want["index_doppler"]
    assert abs(got["test_statistics"] - want["test_statistics"]) < 1e-4 * abs(want["test_statistics"])
    assert abs(got["grid_maximum"] - want["grid_maximum"]) < 1e-4 * want["grid_maximum"]
    assert got["Acq_delay_samples"] == want["Acq_delay_samples"] and got["Acq_samplestamp_samples"] == want["Acq_samplestamp_samples"]
    if want_pos:
        assert got["tmp_index_freq"] == want["tmp_index_freq"]
        spec = blk.fine.read_spectrum()
        assert np.max(np.abs(spec - want["fine_spectrum"])) < 1e-4 * want["fine_spectrum"].max()
    assert got["Acq_doppler_hz"] == want["Acq_doppler_hz"]
    blk.close()
    e.close()


def test_fine_doppler_rejects_unsupported_sizes():
    from gnss_sdr_b200 import capi
    e = capi.Engine()
    for n in (16368, 30000, 1):          # prime factors 11 and 31; 10 N above 10 x 27648; degenerate
        with pytest.raises(capi.B200Error) as ei:
            capi.AcqFineDoppler(e, n)
        assert ei.value.code in (-5, -1), n
    e.close()
