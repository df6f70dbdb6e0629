"""EXTERNAL sanity pin for the restated GPS C/A generator (the reference has no KAT for it,
tests/unit-tests/arithmetic/code_generation_test.cc:28-49): IS-GPS-200 Table 3-Ia lists the first
10 chips of each PRN in octal (PRN 1 = 1440, 2 = 1620, 3 = 1710, 4 = 1744, 5 = 1133, 6 = 1455,
7 = 1131, 8 = 1454, 9 = 1626, 10 = 1504, ..., 32 = 1712)."""
import numpy as np

ICD_FIRST10_OCTAL = {1: 0o1440, 2: 0o1620, 3: 0o1710, 4: 0o1744, 5: 0o1133, 6: 0o1455, 7: 0o1131, 8: 0o1454,
                     9: 0o1626, 10: 0o1504, 11: 0o1642, 12: 0o1750, 13: 0o1764, 14: 0o1772, 15: 0o1775,
                     16: 0o1776, 17: 0o1156, 18: 0o1467, 19: 0o1633, 20: 0o1715, 21: 0o1746, 22: 0o1763,
                     23: 0o1063, 24: 0o1706, 25: 0o1743, 26: 0o1761, 27: 0o1770, 28: 0o1774, 29: 0o1127,
                     30: 0o1453, 31: 0o1625, 32: 0o1712}


def test_ca_first_ten_chips_match_icd(oracle):
    for prn, octal in ICD_FIRST10_OCTAL.items():
        code = oracle.port.gps_ca_code(prn)
        bits = (code[:10] > 0).astype(int)        # reference maps G1^G2 == 1 -> +1
        val = int("".join(str(b) for b in bits), 2)
        assert val == octal, (prn, oct(val), oct(octal))


def test_ca_code_properties(oracle):
    codes = np.stack([oracle.port.gps_ca_code(p) for p in range(1, 33)])
    assert set(np.unique(codes)) == {-1.0, 1.0}
    # balance: 512 ones / 511 zeros -> sum = +-1 ; Gold cross-correlation three-valued {-65, -1, 63}
    assert np.all(np.abs(codes.sum(axis=1)) == 1)
    a, b = codes[0], codes[6]
    xc = np.array([np.dot(a, np.roll(b, k)) for k in range(1023)])
    assert set(np.unique(xc)) <= {-65.0, -1.0, 63.0}
    ac = np.array([np.dot(a, np.roll(a, k)) for k in range(1, 1023)])
    assert set(np.unique(ac)) <= {-65.0, -1.0, 63.0}


def test_sampled_complex_code_is_imaginary_and_last_chip_fixed(oracle):
    c = oracle.port.gps_ca_code_complex_sampled(3, 4000000)
    assert c.size == 4000 and np.all(c.real == 0)
    base = oracle.port.gps_ca_code(3)
    assert c.imag[-1] == base[-1]
    assert np.array_equal(c.imag[:8], np.repeat(base[:3], [4, 4, 4])[:8])


def test_synthetic_input_generators_match_the_oracle(oracle):
    """tests/gnss_synth.py carries its own C/A generator (bench.py and tools/ must not import oracle/ outside the
    CPU-baseline legs); it must produce the very tables the reference's generator does."""
    import gnss_synth as gs
    for prn in range(1, 33):
        assert np.array_equal(gs.gps_ca_code(prn), oracle.port.gps_ca_code(prn)), prn
    for fs in (2000000, 4000000, 25000000):
        for prn in (1, 9, 32):
            assert np.array_equal(gs.gps_ca_code_complex_sampled(prn, fs), oracle.port.gps_ca_code_complex_sampled(prn, fs)), (prn, fs)
