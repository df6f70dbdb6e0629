"""TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's PCPS acquisition arithmetic, following
src/algorithms/acquisition/gnuradio_blocks/pcps_acquisition.cc line by line:

  ctor sizes          :101-116      set_local_code      :218-251
  wipe-off grid       :275-291      doppler_grid        :522-560
  CFAR statistic      :409-449      first/second peak   :452-519
  compute_threshold   :52-56        update_synchro      :580-602

Third-party arithmetic that is NOT under /root/reference:
  * gr::fft::fft_complex_fwd/rev (GNU Radio gr-fft -> FFTW3f, unnormalised, version unpinned
    ">= 3.7.3", call sites pcps_acquisition.cc:140,144,249,535,541) is restated with
    scipy.fft (pocketfft, scipy 1.1x) in float32 -- same transform definition (unnormalised
    forward exp(-j..), unnormalised backward exp(+j..)), different summation order.
  * upstream VOLK volk_32fc_x2_multiply_32fc / volk_32fc_magnitude_squared_32f /
    volk_32f_x2_add_32f / volk_32fc_conjugate_32fc (call sites :250,531,538,547,551-552) are
    element-wise float32 ops, restated with numpy float32 arithmetic.
PARITY AT THE FFT BOUNDARY IS UNPINNED in the reference (no test holds FFT outputs); the
contract is exact (index_time, index_doppler) and test_statistics within 1e-4 relative.
The wipe-off carrier, arg-max and statistics use the reference's own kernels (oracle.ref) or
their bit-exact C port (oracle.port).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np
import scipy.fft as sfft


@dataclass
class AcqConf:
    """Subset of Acq_Conf (src/algorithms/acquisition/libs/acq_conf.h:33-87) the arithmetic uses."""
    fs_in: int = 4000000
    samples_per_ms: float = 4000.0
    samples_per_code: float = 4000.0
    samples_per_chip: int = 4
    sampled_ms: int = 1
    ms_per_code: int = 1
    doppler_max: int = 5000
    doppler_step: int = 250
    max_dwells: int = 1
    pfa: float = 0.0
    threshold: float = 0.0
    bit_transition_flag: bool = False
    use_CFAR_algorithm_flag: bool = True
    make_2_steps: bool = False
    doppler_step2: float = 125.0
    num_doppler_bins_step2: int = 4
    pfa2: float = 0.0

    def __post_init__(self):
        # pcps_acquisition.cc:107-113
        self.consumed_samples = int(self.sampled_ms * self.samples_per_ms * (2.0 if self.bit_transition_flag else 1.0))
        self.fft_size = self.consumed_samples if self.sampled_ms == self.ms_per_code else self.consumed_samples * 2
        self.effective_fft_size = self.fft_size // 2 if self.bit_transition_flag else self.fft_size
        self.num_doppler_bins = int(math.ceil(float(2 * self.doppler_max) / float(self.doppler_step)))


def gamma_p_inv_int(a: int, p: float) -> float:
    """boost::math::gamma_p_inv(a, p) for integer a (a = 2*max_dwells in compute_threshold):
    solve 1 - exp(-x) * sum_{k<a} x^k/k! = p by Newton iteration in double."""
    from scipy.special import gammaincinv
    return float(gammaincinv(float(a), p))


def compute_threshold(pfa: float, effective_fft_size: int, num_doppler_bins: int, max_dwells: int) -> float:
    """pcps_acquisition.cc:52-56 (note the float cast of num_bins inside the exponent)."""
    num_bins = effective_fft_size * num_doppler_bins
    return float(np.float32(2.0 * gamma_p_inv_int(2 * max_dwells, math.pow(1.0 - pfa, 1.0 / float(np.float32(num_bins))))))


class PcpsAcquisitionOracle:
    """State + arithmetic of one pcps_acquisition block (one channel / one PRN)."""

    def __init__(self, conf: AcqConf, sincos_variant: str = "avx2", workers: int = 1):
        import oracle
        self.c = conf
        self.o = oracle
        self.workers = workers
        self.doppler_center = 0
        self.doppler_bias = 0
        self.sincos_variant = sincos_variant
        self.threshold = (compute_threshold(conf.pfa, conf.effective_fft_size, conf.num_doppler_bins,
                                            1 if conf.bit_transition_flag else conf.max_dwells)
                          if conf.pfa > 0.0 else conf.threshold)
        self.threshold_step_two = (compute_threshold(conf.pfa2, conf.effective_fft_size, conf.num_doppler_bins_step2,
                                                     1 if conf.bit_transition_flag else conf.max_dwells)
                                   if conf.pfa2 > 0.0 else conf.threshold)
        self.step_two = False
        self.doppler_center_step_two = np.float32(0.0)
        self.magnitude_grid = np.zeros((conf.num_doppler_bins, conf.fft_size), np.float32)
        self.num_noncoherent_integrations_counter = 0
        self.input_power = np.float32(0.0)
        self.fft_codes = None
        self.update_grid_doppler_wipeoffs()

    # pcps_acquisition.cc:284-291
    def update_grid_doppler_wipeoffs(self):
        c = self.c
        out = np.empty((c.num_doppler_bins, c.fft_size), np.complex64)
        self.o.port.lib.port_acq_wipeoff_grid(C.c_int(0 if self.sincos_variant == "generic" else 1),
                                               C.c_void_p(out.ctypes.data), C.c_uint(c.fft_size),
                                               C.c_uint(c.num_doppler_bins), C.c_int32(c.doppler_max),
                                               C.c_int32(self.doppler_center), C.c_int32(c.doppler_step),
                                               C.c_int32(self.doppler_bias), C.c_int64(c.fs_in))
        self.grid_doppler_wipeoffs = out

    # pcps_acquisition.cc:294-301
    def update_grid_doppler_wipeoffs_step2(self):
        c = self.c
        out = np.empty((c.num_doppler_bins_step2, c.fft_size), np.complex64)
        self.o.port.lib.port_acq_wipeoff_grid_step2(C.c_int(0 if self.sincos_variant == "generic" else 1),
                                                    C.c_void_p(out.ctypes.data), C.c_uint(c.fft_size),
                                                    C.c_uint(c.num_doppler_bins_step2), C.c_float(float(self.doppler_center_step_two)),
                                                    C.c_float(c.doppler_step2), C.c_int64(c.fs_in))
        self.grid_doppler_wipeoffs_step_two = out

    def set_doppler_center(self, center: int):
        self.doppler_center = int(center)
        self.update_grid_doppler_wipeoffs()

    # pcps_acquisition.cc:218-251
    def set_local_code(self, code: np.ndarray):
        c = self.c
        code = np.asarray(code, np.complex64)
        buf = np.zeros(c.fft_size, np.complex64)
        if c.bit_transition_flag:
            off = c.fft_size // 2
            buf[off:] = code[:off]
        elif c.sampled_ms == c.ms_per_code:
            buf[:c.consumed_samples] = code[:c.consumed_samples]
        else:
            buf[c.consumed_samples:] = code[:c.consumed_samples]
        self.fft_codes = np.conj(sfft.fft(buf).astype(np.complex64))

    # pcps_acquisition.cc:522-560
    def doppler_grid(self, inp: np.ndarray):
        c = self.c
        off = c.effective_fft_size if c.bit_transition_flag else 0
        # volk_32fc_x2_multiply_32fc (float32 complex multiply), all bins at once
        wipe = self.grid_doppler_wipeoffs_step_two if self.step_two else self.grid_doppler_wipeoffs
        nb = wipe.shape[0]
        x = (inp[None, :] * wipe).astype(np.complex64)
        X = sfft.fft(x, axis=1, workers=self.workers)
        Y = (X * self.fft_codes[None, :]).astype(np.complex64)
        y = sfft.ifft(Y, axis=1, norm="forward", workers=self.workers)   # unnormalised backward transform
        y = y[:, off:off + c.effective_fft_size]
        mag = (y.real.astype(np.float32) ** 2 + y.imag.astype(np.float32) ** 2).astype(np.float32)
        if self.num_noncoherent_integrations_counter == 1:
            self.magnitude_grid[:nb, :c.effective_fft_size] = mag
        else:
            self.magnitude_grid[:nb, :c.effective_fft_size] += mag

    def _grid_max(self):
        """the arg-max scan shared by both statistics (:417-426 / :464-473): per-bin first
        maximum (volk_gnsssdr_32f_index_max_32u), strict '>' across ascending bins."""
        c = self.c
        grid_maximum = np.float32(0.0)
        index_doppler = 0
        index_time = 0
        nb = c.num_doppler_bins_step2 if self.step_two else c.num_doppler_bins
        g = self.magnitude_grid[:nb, :c.effective_fft_size]
        idx = np.argmax(g, axis=1)     # numpy argmax returns the FIRST maximum, as the generic kernel
        for i in range(g.shape[0]):
            v = g[i, idx[i]]
            if v > grid_maximum:
                grid_maximum = v
                index_doppler = i
                index_time = int(idx[i])
        return grid_maximum, index_doppler, index_time

    def _doppler_step_two(self, index_doppler):
        # (:436 / :480) static_cast<int32_t>(center2 + (float(idx) - float(floor(bins2/2.0))) * step2)
        c = self.c
        off = np.float32(np.float32(index_doppler) - np.float32(math.floor(c.num_doppler_bins_step2 / 2.0))) * np.float32(c.doppler_step2)
        return int(np.float32(self.doppler_center_step_two) + np.float32(off))

    # pcps_acquisition.cc:409-449
    def max_to_input_power_statistic(self):
        c = self.c
        grid_maximum, index_doppler, index_time = self._grid_max()
        if not self.step_two:
            index_opp = (index_doppler + c.num_doppler_bins // 2) % c.num_doppler_bins
            row = self.magnitude_grid[index_opp, :c.effective_fft_size]
            # std::accumulate with a float init: strictly sequential float32 sum
            acc = np.float32(_seq_sum_f32(row))
            self.input_power = np.float32(float(np.float32(acc / np.float32(c.effective_fft_size))) / 2.0 / self.num_noncoherent_integrations_counter)
            doppler = -int(c.doppler_max) + self.doppler_center + c.doppler_step * int(index_doppler)
        else:
            doppler = self._doppler_step_two(index_doppler)
        if self.input_power < np.finfo(np.float32).eps:
            stat = np.float32(0.0)
        else:
            stat = np.float32(grid_maximum / self.input_power)
        return dict(index_time=index_time, index_doppler=index_doppler, doppler=doppler, test_statistics=float(stat),
                    grid_maximum=float(grid_maximum), input_power=float(self.input_power))

    # pcps_acquisition.cc:452-519
    def first_vs_second_peak_statistic(self):
        c = self.c
        first_peak, index_doppler, index_time = self._grid_max()
        if not self.step_two:
            doppler = -int(c.doppler_max) + self.doppler_center + c.doppler_step * int(index_doppler)
        else:
            doppler = self._doppler_step_two(index_doppler)
        n = c.effective_fft_size
        ex1 = index_time - c.samples_per_chip
        ex2 = index_time + c.samples_per_chip
        if ex1 < 0:
            ex1 = n + ex1
        elif ex2 >= n:
            ex2 = ex2 - n
        tmp = self.magnitude_grid[index_doppler, :n].copy()
        idx = ex1
        while True:
            tmp[idx] = 0.0
            idx += 1
            if idx == n:
                idx = 0
            if idx == ex2:
                break
        second_peak = tmp[int(np.argmax(tmp))]
        stat = np.float32(first_peak) / np.float32(second_peak)
        return dict(index_time=index_time, index_doppler=index_doppler, doppler=doppler, test_statistics=float(stat),
                    grid_maximum=float(first_peak), second_peak=float(second_peak))

    # acquisition_core :648-728, single dwell of `inp` (consumed_samples long)
    def acquisition_core(self, inp: np.ndarray):
        c = self.c
        sig = np.zeros(c.fft_size, np.complex64)
        sig[:c.consumed_samples] = np.asarray(inp, np.complex64)[:c.consumed_samples]
        self.num_noncoherent_integrations_counter += 1
        self.doppler_grid(sig)
        res = self.max_to_input_power_statistic() if c.use_CFAR_algorithm_flag else self.first_vs_second_peak_statistic()
        # update_synchro :580-584
        res["acq_delay_samples"] = float(np.fmod(np.float32(res["index_time"]), np.float32(c.samples_per_code)))
        res["acq_doppler_hz"] = float(res["doppler"])
        th = self.threshold_step_two if self.step_two else self.threshold
        above = bool(res["test_statistics"] > th)
        res["step_two"] = self.step_two
        res["positive"] = above
        if above and c.make_2_steps:
            # handle_threshold_reached (:605-626)
            if self.step_two:
                res["positive"] = True
            else:
                res["positive"] = False
                self.doppler_center_step_two = np.float32(res["doppler"])
                self.update_grid_doppler_wipeoffs_step2()
                self.num_noncoherent_integrations_counter = 0
            self.step_two = not self.step_two
        elif self.num_noncoherent_integrations_counter == c.max_dwells:
            self.step_two = False   # handle_integration_done (:639-645)
        if res["positive"] or self.num_noncoherent_integrations_counter == c.max_dwells or c.bit_transition_flag:
            self.num_noncoherent_integrations_counter = 0
        return res


def _seq_sum_f32(row: np.ndarray) -> np.float32:
    """strictly sequential float32 accumulation (std::accumulate(..., 0.0f))."""
    import oracle
    fn = getattr(oracle.port.lib, "port_seq_sum_f32", None)
    if fn is not None:
        fn.restype = C.c_float
        r = np.ascontiguousarray(row, np.float32)
        return np.float32(fn(C.c_void_p(r.ctypes.data), C.c_uint(r.size)))
    acc = np.float32(0.0)
    for v in row:
        acc = np.float32(acc + v)
    return acc
