"""Committed golden vectors generated from the reference's own code (tests/golden/make_golden.py).
CPU: the C port reproduces them bit for bit.  GPU: the CUDA path matches them within the contract."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES, case_inputs  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "trk_ref_golden.npz"))
GA = np.load(os.path.join(HERE, "golden", "acq_ref_golden.npz"))


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("case", [c for c in CASES if not c[-1]], ids=lambda c: c[0])
def test_port_reproduces_reference_golden_taps(oracle, case):
    name, shifts = case[0], case[6]
    code, iq, (rc, dp, rcode, st) = case_inputs(oracle, case)
    assert np.array_equal(np.array([rc, dp, rcode, st], np.float32), G[f"{name}/params"])   # same seeded inputs
    a = oracle.port.multicorrelator(1, iq, code, shifts, rc, dp, rcode, st)
    g = oracle.port.multicorrelator(0, iq, code, shifts, rc, dp, rcode, st)
    assert np.array_equal(_bits(a), _bits(G[f"{name}/a_avx"]))
    assert np.array_equal(_bits(g), _bits(G[f"{name}/generic"]))


def test_port_reproduces_reference_golden_wipeoff(oracle):
    for key in GA.files:
        if not key.endswith("/head"):
            continue
        _, fs, f, _ = key.split("/")
        fs, f = int(fs), int(f)
        n = int(fs // 1000)
        inc = -np.float32(np.float32(2 * np.pi) * np.float32(f) / np.float32(fs))
        got = np.empty(n, np.complex64)
        ph = C.c_float(0.0)
        oracle.port.lib.port_sincos_avx2(C.c_void_p(got.ctypes.data), C.c_float(float(inc)), C.byref(ph), C.c_uint(n))
        bits = got.view(np.uint32)
        assert np.array_equal(bits[:128], GA[key])
        assert np.array_equal(bits[-128:], GA[key.replace("/head", "/tail")])
        chk = np.array([np.bitwise_xor.reduce(bits), np.sum(bits.astype(np.uint64)) & 0xFFFFFFFFFFFF], np.uint64)
        assert np.array_equal(chk, GA[key.replace("/head", "/xor_sum")])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_gpu_matches_reference_golden_taps(oracle, case):
    import gnss_sdr_b200.capi as capi
    name, shifts, hd = case[0], case[6], case[-1]
    code, iq, (rc, dp, rcode, st) = case_inputs(oracle, case)
    eng = capi.Engine(0)
    mc = capi.Multicorrelator(eng, len(iq), len(shifts))
    mc.set_high_dynamics_resampler(hd)
    mc.set_local_code_and_taps(code, shifts)
    rate = (2e-9, 2e-12) if hd else (0.0, 0.0)
    got = mc.Carrier_wipeoff_multicorrelator_resampler(iq, rc, dp, rate[0], rcode, st, rate[1])
    mc.free()
    eng.close()
    want = G[f"{name}/a_avx"]
    assert np.all(np.abs(got - want) / np.abs(want) < 1e-3)                       # the reference's SIMD-vs-generic bound
    if not hd:
        # and not further from the reference's SIMD result than ~ the reference's own generic kernel is
        # (the high-dynamics rotator has only a generic, cpowf-based implementation whose float phase is
        # itself ~1e-4 off, so that comparison is meaningless there)
        assert np.max(np.abs(got - want)) <= 3 * np.max(np.abs(G[f"{name}/generic"] - want)) + 2e-5 * np.max(np.abs(want))


@pytest.mark.gpu
def test_gpu_wipeoff_matches_reference_golden(oracle):
    import gnss_sdr_b200.capi as capi
    eng = capi.Engine(0)
    for fs, n, dmax, dstep in ((4000000, 4000, 5000, 250), (25000000, 25000, 10125, 250)):
        acq = capi.PcpsAcquisition(eng, fs_in=fs, samples_per_ms=float(n), samples_per_chip=max(1, int(fs / 1.023e6)),
                                   doppler_max=dmax, doppler_step=dstep)
        w = acq.read_wipeoffs()
        for f in (-5000.0, -250.0, 9875.0):
            d = (int(f) + dmax) // dstep
            if not (0 <= d < acq.conf.num_doppler_bins) or -dmax + dstep * d != int(f):
                continue
            bits = w[d].view(np.uint32)
            assert np.array_equal(bits[:128], GA[f"sincos/{fs}/{int(f)}/head"])
            assert np.array_equal(bits[-128:], GA[f"sincos/{fs}/{int(f)}/tail"])
        acq.close()
    eng.close()
