"""Host side of pcps_acquisition_fine_doppler_cc over the C ABI (b200_acq + b200_acq_fine).

Mirrors the block's members and state machine
(src/algorithms/acquisition/gnuradio_blocks/pcps_acquisition_fine_doppler_cc.{h,cc}): state 1 accumulates
max_dwells one-millisecond grids (compute_and_accumulate_grid :266-299), state 2 compute_CAF (:182-251) and the
threshold decision (:454-466), state 3 fills the 10 ms buffer and runs estimate_Doppler (:316-389).  The grid
search and the statistics run on the device through b200_acq (first-vs-second-peak statistic, dwell accumulation),
the fine transform through b200_acq_fine; this class keeps what the block keeps on the host: the buffer, the
replica alignment, the frequency mapping and the plausibility check.
"""
from __future__ import annotations

import math

import numpy as np

from . import capi

GPS_L1_CA_CHIP_PERIOD_S = 1.0 / 1.023e6


class PcpsAcquisitionFineDoppler:
    def __init__(self, engine: capi.Engine, *, fs_in: int, samples_per_ms: float, doppler_max: int, doppler_step: int,
                 max_dwells: int, threshold: float):
        self.fs_in, self.doppler_max, self.doppler_step = int(fs_in), int(doppler_max), int(doppler_step)
        self.max_dwells, self.threshold = int(max_dwells), float(threshold)
        self.d_num_doppler_points = int(math.floor(abs(2 * doppler_max) / doppler_step))   # :57
        self.d_fft_size = int(samples_per_ms)                                              # :60
        spc = int(math.ceil(np.float32(GPS_L1_CA_CHIP_PERIOD_S) * np.float32(self.fs_in)))  # :214
        # update_carrier_wipeoff (:163-179) starts the grid at -doppler_step: that is b200_acq's wipe-off formula
        # (-doppler_max + doppler_step * index) with doppler_max := doppler_step
        self.acq = capi.PcpsAcquisition(engine, fs_in=self.fs_in, samples_per_ms=samples_per_ms, samples_per_chip=spc,
                                        doppler_max=self.doppler_step, doppler_step=self.doppler_step,
                                        use_CFAR_algorithm_flag=False, max_dwells=max(2, self.max_dwells), keep_grid=True,
                                        num_doppler_bins=self.d_num_doppler_points)
        self.fine = capi.AcqFineDoppler(engine, self.d_fft_size)
        self.d_10_ms_buffer = np.zeros(50 * self.d_fft_size, np.complex64)
        self.d_n_samples_in_buffer = 0
        self.d_sample_counter = 0
        self.d_test_statistics = 0.0
        self.Acq_delay_samples = 0.0
        self.Acq_doppler_hz = 0.0
        self.Acq_samplestamp_samples = 0
        self.code_complex_sampled = None

    def set_local_code(self, code_complex_sampled):
        self.code_complex_sampled = np.ascontiguousarray(code_complex_sampled, np.complex64)[:self.d_fft_size].copy()
        self.acq.set_local_code(0, self.code_complex_sampled)

    @staticmethod
    def rotate_code_replica(code_1ms, shift_index):
        c = np.array(code_1ms, np.complex64)
        n = len(c)
        if shift_index != 0:                       # std::rotate(first, first + (N - shift), first + N - 1)  (:336-340)
            mid = n - shift_index
            c[:n - 1] = np.concatenate([c[mid:n - 1], c[:mid]])
        return c

    def fft_freq_bins(self, idx: int, fft_size_extended: int) -> np.float32:   # :360-373
        fs = float(np.float32(self.fs_in))
        half = float(np.float32(fft_size_extended)) / 2.0
        if idx < fft_size_extended // 2:
            return np.float32((fs / 2.0) * float(np.float32(idx)) / half)
        return np.float32((-fs / 2.0) * float(np.float32(fft_size_extended - idx)) / half)

    def run(self, samples):
        """general_work from activation to the decision (noutput_items = d_fft_size per call).
        Returns (positive_acquisition, dict of the Gnss_Synchro fields and intermediate indices)."""
        n = self.d_fft_size
        samples = np.ascontiguousarray(samples, np.complex64)
        self.d_n_samples_in_buffer = 0
        pos = 0
        res = None
        for dwell in range(1, self.max_dwells + 1):                                   # state 1
            blk = samples[pos:pos + n]
            res = self.acq.search(blk, [0], dwell_counter=dwell)[0]
            self.d_10_ms_buffer[self.d_n_samples_in_buffer:self.d_n_samples_in_buffer + n] = blk
            self.d_n_samples_in_buffer += n
            self.d_sample_counter += n
            pos += n
        # state 2: compute_CAF - the device already holds first peak / second peak of the accumulated grid
        self.d_test_statistics = float(res["test_statistics"])
        self.Acq_delay_samples = float(res["index_time"])
        self.Acq_doppler_hz = float(int(res["index_doppler"]) * self.doppler_step - self.doppler_max)   # :246
        self.Acq_samplestamp_samples = self.d_sample_counter
        out = dict(index_time=int(res["index_time"]), index_doppler=int(res["index_doppler"]), test_statistics=self.d_test_statistics,
                   grid_maximum=float(res["grid_maximum"]), second_peak=float(res["second_peak"]))
        positive = self.d_test_statistics > self.threshold
        if positive:                                                                  # state 3
            remaining = 10 * n - self.d_n_samples_in_buffer
            if remaining > 0:
                self.d_10_ms_buffer[self.d_n_samples_in_buffer:self.d_n_samples_in_buffer + remaining] = samples[pos:pos + remaining]
                self.d_n_samples_in_buffer += remaining
                self.d_sample_counter += remaining
            rep = self.rotate_code_replica(self.code_complex_sampled, int(self.Acq_delay_samples))
            idx, peak = self.fine.estimate(self.d_10_ms_buffer[:10 * n], rep)
            f = self.fft_freq_bins(idx, 80 * n)
            out["tmp_index_freq"] = idx
            if abs(float(f) - self.Acq_doppler_hz) < 1000:                            # :376-379
                self.Acq_doppler_hz = float(f)
            self.d_n_samples_in_buffer = 0
        out.update(Acq_delay_samples=self.Acq_delay_samples, Acq_doppler_hz=self.Acq_doppler_hz,
                   Acq_samplestamp_samples=self.Acq_samplestamp_samples, Acq_doppler_step=self.doppler_step)
        return positive, out

    def close(self):
        self.fine.close()
        self.acq.close()
