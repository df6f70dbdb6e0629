"""GPU-box probe: acquisition parity against the numpy oracle over a list of FFT sizes (chirp-z and plain)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gnss_sdr_b200.capi as capi  # noqa: E402
import oracle  # noqa: E402
from gnss_synth import make_iq  # noqa: E402
from oracle.acq_np import AcqConf, PcpsAcquisitionOracle  # noqa: E402

eng = capi.Engine(0)
for fs, ms in [(16.368e6, 1), (16.368e6, 2), (20.46e6, 1), (30.69e6, 1), (65.536e6, 1), (32.768e6, 2), (40.92e6, 1), (16.368e6, 3), (16.368e6, 4)]:
    spms = fs / 1000.0
    n = int(spms) * ms
    spchip = int(fs / 1.023e6)
    dmax, dstep, prn = 2000, 250, 7
    codes = {7: oracle.port.gps_ca_code(7)}
    svs = [dict(prn=7, doppler=1260.0, code_phase_chips=321.4, cn0=47.0, phase0=1.0)]
    iq = make_iq(codes, fs, n, svs, seed=3)
    kw = dict(sampled_ms=ms, ms_per_code=ms) if ms > 1 else {}
    conf = AcqConf(fs_in=int(fs), samples_per_ms=spms, samples_per_code=spms, samples_per_chip=spchip, doppler_max=dmax, doppler_step=dstep,
                   pfa=0.001, threshold=0.0, use_CFAR_algorithm_flag=True, max_dwells=1, **kw)
    o = PcpsAcquisitionOracle(conf)
    local = np.tile(oracle.port.gps_ca_code_complex_sampled(prn, int(fs)), ms)
    o.set_local_code(local)
    want = o.acquisition_core(iq)
    try:
        acq = capi.PcpsAcquisition(eng, fs_in=int(fs), samples_per_ms=spms, samples_per_chip=spchip, doppler_max=dmax, doppler_step=dstep,
                                   use_CFAR_algorithm_flag=True, keep_grid=True, **kw)
    except capi.B200Error as e:
        print(n, "create failed", e)
        continue
    acq.set_local_code(0, local)
    got = acq.search(iq, [0])[0]
    g = acq.read_grid(0)
    ref_g = o.magnitude_grid[:, :n]
    err = np.max(np.abs(g - ref_g)) / ref_g.max()
    print(f"N={n} got t={int(got['index_time']) % int(spms)} d={int(got['index_doppler'])} stat={got['test_statistics']:.3f} | "
          f"want t={want['index_time'] % int(spms)} d={want['index_doppler']} stat={want['test_statistics']:.3f} | grid err {err:.2e}", flush=True)
    acq.close()
eng.close()

# the transform alone: chirp-z DFT of random vectors against numpy
eng = capi.Engine(0)
rng = np.random.default_rng(1)
for fs in [5.456e6, 13e6, 16.368e6, 17.391e6, 20.46e6, 30.69e6, 32.736e6, 40.92e6]:
    n = int(fs / 1000)
    try:
        acq = capi.PcpsAcquisition(eng, fs_in=int(fs), samples_per_ms=fs / 1000.0, samples_per_chip=int(fs / 1.023e6), doppler_max=500, doppler_step=250)
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        got = acq.selftest_dft(x)
        ref = np.fft.fft(x.astype(np.complex128))
        e = np.abs(got - ref)
        print(f"DFT N={n}: max err {e.max() / np.abs(ref).max():.2e}; first bad index {np.argmax(e > 1e-3 * np.abs(ref).max()) if (e > 1e-3 * np.abs(ref).max()).any() else -1}, "
              f"bad count {(e > 1e-3 * np.abs(ref).max()).sum()}", flush=True)
        acq.close()
    except capi.B200Error as ex:
        print(n, "failed", ex)
eng.close()
