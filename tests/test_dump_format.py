"""On-disk formats, host-only entry points of the product library - no GPU needed.

* acquisition dump: variables of pcps_acquisition::dump_results (pcps_acquisition.cc:354-406) in a Level-5 MAT-file,
  read back with scipy.io.loadmat (an independent implementation of the format) and checked against the names,
  classes and shapes the reference's Acquisition_Dump_Reader expects
  (tests/unit-tests/signal-processing-blocks/libs/acquisition_dump_reader.cc:26-131).
* tracking dump: see tests/test_oracle_loop.py::test_dump_file_is_readable_by_the_reference_reader.
"""
import numpy as np
import scipy.io

from gnss_sdr_b200 import capi


def test_acq_dump_mat_file_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    bins, ne = 41, 4000
    grid = rng.random((bins, ne), dtype=np.float32) * 1e6
    grid[17, 524] = 3.3e9
    fn = capi.acq_dump_filename(str(tmp_path / "acquisition"), "G", "1C", 2, 7, 19)
    assert fn.endswith("acquisition_G_1C_ch_2_7_sat_19.mat")
    capi.acq_dump_write(fn, grid, doppler_max=5000, doppler_step=250, positive_acq=True, acq_doppler_hz=-750.0,
                        acq_delay_samples=524.0, test_statistic=12.5, threshold=2.75, input_power=1.9e6,
                        sample_counter=(1 << 40) + 12345, prn=19, num_dwells=2)
    m = scipy.io.loadmat(fn)
    # d_grid is arma::fmat(effective_fft_size, num_doppler_bins); the reader indexes acq_grid(sample, doppler)
    assert m["acq_grid"].shape == (ne, bins) and m["acq_grid"].dtype == np.float32
    assert np.array_equal(m["acq_grid"], grid.T)
    assert m["acq_grid"][524, 17] == np.float32(3.3e9)
    expect = dict(doppler_max=(np.int32, 5000), doppler_step=(np.int32, 250), positive_acq=(np.int32, 1),
                  acq_doppler_hz=(np.float32, -750.0), acq_delay_samples=(np.float32, 524.0), test_statistic=(np.float32, 12.5),
                  threshold=(np.float32, 2.75), input_power=(np.float32, 1.9e6), sample_counter=(np.uint64, (1 << 40) + 12345),
                  PRN=(np.uint32, 19), num_dwells=(np.int32, 2))
    for name, (dt, val) in expect.items():
        assert m[name].shape == (1, 1) and m[name].dtype == dt, name
        assert m[name][0, 0] == dt(val), name
    assert "acq_grid_narrow" not in m


def test_acq_dump_two_step_variables(tmp_path):
    rng = np.random.default_rng(4)
    grid = rng.random((8, 100), dtype=np.float32)
    narrow = rng.random((5, 100), dtype=np.float32)
    fn = str(tmp_path / "a.mat")
    capi.acq_dump_write(fn, grid, doppler_max=1000, doppler_step=250, positive_acq=False, acq_doppler_hz=0.0, acq_delay_samples=1.0,
                        test_statistic=0.5, threshold=2.0, input_power=1.0, sample_counter=5, prn=1, num_dwells=1,
                        grid_narrow=narrow, doppler_step_narrow=50.0, doppler_grid_narrow_min=-100.0)
    m = scipy.io.loadmat(fn)
    assert np.array_equal(m["acq_grid_narrow"], narrow.T)
    assert m["doppler_step_narrow"][0, 0] == np.float32(50.0) and m["doppler_grid_narrow_min"][0, 0] == np.float32(-100.0)
    assert m["positive_acq"][0, 0] == 0
