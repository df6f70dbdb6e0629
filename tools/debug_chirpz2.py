"""GPU-box probe: where does the chirp-z search go wrong for N > 16384?"""
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
for fs, dmax in [(20.46e6, 250), (17.391e6, 250), (32.736e6, 250)]:
    spms = fs / 1000.0
    n = int(spms)
    spchip = int(fs / 1.023e6)
    dstep, prn = 250, 7
    codes = {7: oracle.port.gps_ca_code(7)}
    svs = [dict(prn=7, doppler=160.0, code_phase_chips=321.4, cn0=47.0, phase0=1.0)]
    iq = make_iq(codes, fs, n, svs, seed=3)
    conf = AcqConf(fs_in=int(fs), samples_per_ms=spms, samples_per_code=spms, samples_per_chip=spchip, doppler_max=dmax, doppler_step=dstep,
                   pfa=0.001, threshold=0.0, use_CFAR_algorithm_flag=True, max_dwells=1)
    o = PcpsAcquisitionOracle(conf)
    local = oracle.port.gps_ca_code_complex_sampled(prn, int(fs))
    o.set_local_code(local)
    want = o.acquisition_core(iq)
    acq = capi.PcpsAcquisition(eng, fs_in=int(fs), samples_per_ms=spms, samples_per_chip=spchip, doppler_max=dmax, doppler_step=dstep,
                               use_CFAR_algorithm_flag=True, keep_grid=True)
    acq.set_local_code(0, local)
    cw = acq.selftest_read(1)[0]
    got = acq.search(iq, [0])[0]
    xs = acq.selftest_read(0)
    g = acq.read_grid(0)
    ref_g = o.magnitude_grid[:, :n]
    bins = xs.shape[0]
    ref_xs = np.fft.fft(iq[None, :].astype(np.complex128) * o.grid_doppler_wipeoffs.astype(np.complex128), axis=1)
    e_xs = np.abs(xs - ref_xs).max(axis=1) / np.abs(ref_xs).max()
    k = np.arange(n)
    w = np.exp(1j * np.pi * ((k * k) % (2 * n)) / n)
    ref_cw = np.conj(np.fft.fft(local.astype(np.complex128))) * w
    scale = np.vdot(ref_cw, cw) / np.vdot(ref_cw, ref_cw)
    e_cw = np.abs(cw - scale * ref_cw).max() / np.abs(scale * ref_cw).max()
    e_rows = np.abs(g - ref_g).max(axis=1) / ref_g.max()
    print(f"N={n} bins={bins}: spectra err per bin {np.array2string(e_xs, precision=1)}; CW err {e_cw:.1e} (scale {abs(scale):.3e}, 1/M?);\n"
          f"   grid err per bin {np.array2string(e_rows, precision=1)}; row sums ratio {np.array2string(g.sum(axis=1) / ref_g.sum(axis=1), precision=3)}", flush=True)
    r0 = 0
    e = np.abs(g[r0] - ref_g[r0])
    j = int(np.argmax(e))
    print("   row 0 first 6 gpu", g[r0, :6], "ref", ref_g[r0, :6])
    print("   row 0 worst index", j, "gpu", g[r0, max(0, j - 2):j + 3], "ref", ref_g[r0, max(0, j - 2):j + 3])
    ratio = g[r0] / np.maximum(ref_g[r0], 1e-30)
    print("   row 0 ratio quantiles", np.quantile(ratio, [0.01, 0.25, 0.5, 0.75, 0.99]), "corr of gpu row with ref row", np.corrcoef(g[r0], ref_g[r0])[0, 1])
    for sh in (0, 1, -1, 16384, n - 16384, 20480 - n, 4096):
        print("   corr with ref rolled by", sh, np.corrcoef(g[r0], np.roll(ref_g[r0], sh))[0, 1])
    # energy by segment of 2048 outputs
    seg = 2048
    print("   energy ratio per 2048-segment", np.array2string(np.add.reduceat(g[r0], np.arange(0, n, seg)) / np.add.reduceat(ref_g[r0], np.arange(0, n, seg)), precision=2))
    acq.close()
eng.close()
