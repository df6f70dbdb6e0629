"""Host-side DLL/PLL loop used by the chain test (BASELINE configs[0]: 1 channel, 4 Msps, acquisition +
dll_pll tracking).  A compact restatement, in the reference's own float/double types and update order, of

  run_dll_pll            src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc:1260-1310
  update_tracking_vars   ...:1409-1483
  pll_cloop_two_quadrant_atan / dll_nc_e_minus_l_normalized   tracking/libs/tracking_discriminators.cc:100-128
  Tracking_FLL_PLL_filter (order 3, PLL only)                 tracking/libs/tracking_FLL_PLL_filter.cc:27-104
  Tracking_loop_filter (order 2, no last integrator)          tracking/libs/tracking_loop_filter.cc:62-165

The loop is scalar host math (row a6, stays on the host); `correlate` is the hot path under test and is
injected, so the same loop runs over the reference's CPU correlator and over the B200 one.
"""
import math

import numpy as np

f32 = np.float32
TWO_PI = 2.0 * math.pi
GPS_L1_FREQ_HZ = 1575.42e6
CODE_RATE = 1.023e6
CODE_LEN = 1023


class FllPllFilter3:
    def __init__(self, pll_bw_hz):
        self.b3, self.a3 = f32(2.4), f32(1.1)
        self.w0p = f32(pll_bw_hz) / f32(0.7845)
        self.w0p2 = self.w0p * self.w0p
        self.w0p3 = self.w0p2 * self.w0p
        self.w = f32(0.0)
        self.x = f32(0.0)

    def initialize(self, doppler_hz):
        self.x = f32(2.0) * f32(doppler_hz)
        self.w = f32(0.0)

    def get_carrier_error(self, pll_disc, T):
        pll_disc, T = f32(pll_disc), f32(T)
        self.w = self.w + T * (self.w0p3 * pll_disc)
        self.x = self.x + T * (f32(0.5) * self.w + self.a3 * self.w0p2 * pll_disc)
        return f32(0.5) * self.x + self.b3 * self.w0p * pll_disc


class LoopFilter2:
    def __init__(self, T, bw):
        T, bw = f32(T), f32(bw)
        zeta = f32(1.0) / np.sqrt(f32(2.0))
        wn = bw * (f32(8.0) * zeta) / (f32(4.0) * zeta * zeta + f32(1.0))
        g1 = wn * wn
        g2 = wn * f32(2.0) * zeta
        self.cin = [f32(g1 * T / 2.0 + g2), f32(g1 * T / 2.0 - g2)]
        self.prev_in = f32(0.0)
        self.prev_out = f32(0.0)

    def apply(self, x):
        x = f32(x)
        result = f32(1.0) * self.prev_out
        result = f32(result + self.cin[0] * x + self.cin[1] * self.prev_in)
        self.prev_in = x
        self.prev_out = result
        return result


def run_tracking(correlate, iq, fs, acq_delay_samples, acq_doppler_hz, n_epochs, pll_bw=35.0, dll_bw=2.0, els=0.5):
    """correlate(in_block, rem_carr, phase_step, rem_code_chips, code_step_chips, n) -> complex64[3] (E,P,L).
    Returns dict of per-epoch arrays."""
    vector_length = int(round(fs / (CODE_RATE / CODE_LEN)))
    carrier_filter = FllPllFilter3(pll_bw)
    code_filter = LoopFilter2(CODE_LEN / CODE_RATE, dll_bw)
    carrier_filter.initialize(acq_doppler_hz)
    # start_tracking (:791-1078): initial NCOs from the acquisition result
    carrier_doppler_hz = float(acq_doppler_hz)
    code_freq_chips = CODE_RATE * (1.0 + carrier_doppler_hz / GPS_L1_FREQ_HZ)   # radial velocity correction
    carrier_phase_step_rad = TWO_PI * carrier_doppler_hz / fs
    code_phase_step_chips = code_freq_chips / fs
    rem_carr_phase_rad = f32(0.0)
    rem_code_phase_samples = 0.0
    rem_code_phase_chips = 0.0
    pos = int(acq_delay_samples)                    # pull-in: align to the next PRN start (:1948-1980)
    out = dict(P=[], doppler=[], code_err=[], carr_err=[], pos=[])
    for _ in range(n_epochs):
        if pos + 2 * vector_length > len(iq):
            break
        taps = correlate(iq[pos:pos + vector_length], float(rem_carr_phase_rad), float(f32(carrier_phase_step_rad)),
                         float(f32(rem_code_phase_chips)), float(f32(code_phase_step_chips)), vector_length)
        E, P, L = taps
        # run_dll_pll
        carr_phase_error_hz = (math.atan(float(P.imag) / float(P.real)) if float(P.real) != 0.0 else 0.0) / TWO_PI
        carr_error_filt_hz = carrier_filter.get_carrier_error(carr_phase_error_hz, CODE_LEN / CODE_RATE)
        carrier_doppler_hz = float(carr_error_filt_hz)
        pe, pl = abs(complex(E)), abs(complex(L))
        code_error_chips = ((1.0 - 1.0 * els) / 1.0) * (pe - pl) / (pe + pl) if (pe + pl) != 0 else 0.0
        code_error_filt_chips = float(code_filter.apply(code_error_chips))
        code_freq_chips = CODE_RATE - code_error_filt_chips
        code_freq_chips += carrier_doppler_hz * CODE_RATE / GPS_L1_FREQ_HZ       # carrier aiding
        # update_tracking_vars
        T_prn_samples = (1.0 / code_freq_chips) * CODE_LEN * fs
        K_blk_samples = T_prn_samples + rem_code_phase_samples
        cur_len = int(math.floor(K_blk_samples))
        carrier_phase_step_rad = TWO_PI * carrier_doppler_hz / fs
        rem_carr_phase_rad = f32(rem_carr_phase_rad + f32(carrier_phase_step_rad * cur_len))
        rem_carr_phase_rad = f32(math.fmod(float(rem_carr_phase_rad), TWO_PI))
        code_phase_step_chips = code_freq_chips / fs
        rem_code_phase_samples = K_blk_samples - cur_len
        rem_code_phase_chips = code_freq_chips * rem_code_phase_samples / fs
        out["P"].append(complex(P))
        out["doppler"].append(carrier_doppler_hz)
        out["code_err"].append(code_error_chips)
        out["carr_err"].append(carr_phase_error_hz)
        out["pos"].append(pos)
        pos += cur_len
    return {k: np.array(v) for k, v in out.items()}
