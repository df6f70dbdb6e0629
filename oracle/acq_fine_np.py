"""TEST INFRASTRUCTURE ONLY.

numpy restatement of pcps_acquisition_fine_doppler_cc
(src/algorithms/acquisition/gnuradio_blocks/pcps_acquisition_fine_doppler_cc.cc), the acquisition block the
default conf/gnss-sdr.conf selects (GPS_L1_CA_PCPS_Acquisition_Fine_Doppler):

  ctor sizes                    :44-70      (d_num_doppler_points = floor(|2 doppler_max| / doppler_step), d_fft_size = samples_per_ms)
  set_local_code                :130-136    update_carrier_wipeoff   :163-179
  compute_and_accumulate_grid   :266-299    compute_CAF              :182-251
  estimate_Doppler              :316-389    general_work states      :400-557

Restated as found, including what looks like upstream slips: the wipe-off frequencies start at -doppler_step
(:172) while the reported Doppler uses -doppler_max (:246), and the code replica is rotated over fft_size - 1
elements (:338).  FFTs (gr::fft -> FFTW3f, absent) are scipy.fft in float32: parity unpinned at the FFT boundary,
contract = exact indices and 1e-4 relative statistics, as for pcps_acquisition (oracle/acq_np.py).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import scipy.fft as sfft

GPS_L1_CA_CHIP_PERIOD_S = 1.0 / 1.023e6


class FineDopplerOracle:
    def __init__(self, fs_in: int, samples_per_ms: float, doppler_max: int, doppler_step: int, max_dwells: int, threshold: float,
                 sincos_variant: str = "avx2"):
        import oracle
        self.o = oracle
        self.fs_in, self.doppler_max, self.doppler_step = int(fs_in), int(doppler_max), int(doppler_step)
        self.max_dwells, self.threshold = int(max_dwells), float(threshold)
        self.num_doppler_points = int(math.floor(abs(2 * doppler_max) / doppler_step))
        self.fft_size = int(samples_per_ms)
        n, nb = self.fft_size, self.num_doppler_points
        # update_carrier_wipeoff (:163-179): doppler_hz = doppler_step * index - doppler_step
        self.wipe = np.empty((nb, n), np.complex64)
        oracle.port.lib.port_acq_wipeoff_grid(C.c_int(0 if sincos_variant == "generic" else 1), C.c_void_p(self.wipe.ctypes.data),
                                              C.c_uint(n), C.c_uint(nb), C.c_int32(self.doppler_step), C.c_int32(0),
                                              C.c_int32(self.doppler_step), C.c_int32(0), C.c_int64(self.fs_in))
        self.grid = np.zeros((nb, n), np.float32)
        self.well_count = 0
        self.buffer = np.zeros(50 * n, np.complex64)
        self.n_in_buffer = 0
        self.fft_codes = None
        self.result = {}

    def set_local_code(self, code):
        self.fft_codes = np.conj(sfft.fft(np.asarray(code, np.complex64)[:self.fft_size]).astype(np.complex64))

    def reset_grid(self):
        self.well_count = 0
        self.grid[:] = 0.0

    def compute_and_accumulate_grid(self, inp):
        x = (inp[None, :self.fft_size] * self.wipe).astype(np.complex64)
        X = sfft.fft(x, axis=1)
        Y = (X * self.fft_codes[None, :]).astype(np.complex64)
        y = sfft.ifft(Y, axis=1, norm="forward")
        mag = (y.real.astype(np.float32) ** 2 + y.imag.astype(np.float32) ** 2).astype(np.float32)
        self.grid = (self.grid + mag).astype(np.float32)

    def compute_CAF(self, sample_counter=0):
        n = self.fft_size
        first_peak, index_doppler, index_time = np.float32(0.0), 0, 0
        for i in range(self.num_doppler_points):
            t = int(np.argmax(self.grid[i]))          # volk_gnsssdr_32f_index_max_32u: first maximum
            if self.grid[i][t] > first_peak:
                first_peak, index_doppler, index_time = self.grid[i][t], i, t
        spc = int(math.ceil(np.float32(GPS_L1_CA_CHIP_PERIOD_S) * np.float32(self.fs_in)))
        ex1, ex2 = index_time - spc, index_time + spc
        if ex1 < 0:
            ex1 = n + ex1
        elif ex2 >= n:
            ex2 = ex2 - n
        row = self.grid[index_doppler]
        idx = ex1
        while True:
            row[idx] = 0.0
            idx += 1
            if idx == n:
                idx = 0
            if idx == ex2:
                break
        second_peak = row[int(np.argmax(row))]
        stat = np.float32(first_peak) / np.float32(second_peak)
        self.result = dict(index_time=index_time, index_doppler=index_doppler, test_statistics=float(stat),
                           grid_maximum=float(first_peak), second_peak=float(second_peak),
                           Acq_delay_samples=float(index_time),
                           Acq_doppler_hz=float(index_doppler * self.doppler_step - self.doppler_max),
                           Acq_samplestamp_samples=sample_counter, Acq_doppler_step=self.doppler_step, spc=spc)
        return float(stat)

    @staticmethod
    def rotate_code_replica(code_1ms, shift_index):
        """std::rotate(first, first + (N - shift), first + N - 1) of :336-340: the last element stays where it is."""
        c = np.array(code_1ms, np.complex64)
        n = len(c)
        if shift_index != 0:
            mid = n - shift_index
            c[:n - 1] = np.concatenate([c[mid:n - 1], c[:mid]])
        return c

    def fft_freq_bins(self, idx, fft_size_extended):
        fs = np.float32(self.fs_in)
        half = float(np.float32(fft_size_extended)) / 2.0
        if idx < fft_size_extended // 2:
            return np.float32((float(fs) / 2.0) * float(np.float32(idx)) / half)
        k = fft_size_extended - idx
        return np.float32((-float(fs) / 2.0) * float(np.float32(k)) / half)

    def estimate_Doppler(self, code_complex_sampled_1ms):
        n = self.fft_size
        signal_samples = 10 * n
        ext = signal_samples * 8
        rep = self.rotate_code_replica(code_complex_sampled_1ms[:n], int(self.result["Acq_delay_samples"]))
        replica = np.tile(rep, 10)
        x = np.zeros(ext, np.complex64)
        x[:signal_samples] = (self.buffer[:signal_samples] * replica).astype(np.complex64)
        X = sfft.fft(x)
        mag = (X.real.astype(np.float32) ** 2 + X.imag.astype(np.float32) ** 2).astype(np.float32)
        idx = int(np.argmax(mag))
        f = self.fft_freq_bins(idx, ext)
        self.result["tmp_index_freq"] = idx
        self.result["fine_spectrum"] = mag
        if abs(float(f) - self.result["Acq_doppler_hz"]) < 1000:
            self.result["Acq_doppler_hz"] = float(f)
        return rep

    def run(self, samples, code_complex_sampled_1ms):
        """general_work from `set_active` to the positive / negative decision over a contiguous sample vector
        (noutput_items = fft_size per call).  Returns (positive, result dict)."""
        n = self.fft_size
        self.reset_grid()
        self.n_in_buffer = 0
        pos = 0
        sample_counter = 0
        for _ in range(self.max_dwells):                      # state 1
            blk = samples[pos:pos + n]
            self.compute_and_accumulate_grid(blk)
            self.buffer[self.n_in_buffer:self.n_in_buffer + n] = blk
            self.n_in_buffer += n
            self.well_count += 1
            pos += n
            sample_counter += n
        stat = self.compute_CAF(sample_counter)               # state 2
        if not (stat > self.threshold):
            return False, self.result
        remaining = 10 * n - self.n_in_buffer                 # state 3
        if remaining > 0:
            self.buffer[self.n_in_buffer:self.n_in_buffer + remaining] = samples[pos:pos + remaining]
            self.n_in_buffer += remaining
        self.estimate_Doppler(code_complex_sampled_1ms)
        return True, self.result
