#include "b200_dll_pll_veml_loop.h"

#include "b200_multicorrelator_real_codes.h"  // b200::shared_engine

#include <cmath>

namespace b200
{
namespace
{
int g_n_loops = 0;  // loops created on the shared engine (the run call returns every loop's records)
}

bool B200_Dll_Pll_Veml_Loop::init(const Dll_Pll_Conf_Core& conf, const Signal_Core& sig, int band, const float* tracking_code)
{
    b200_engine* eng = shared_engine();
    if (eng == nullptr || tracking_code == nullptr) return false;
    const int taps = sig.veml ? 5 : 3;
    if (b200_trk_channel_create(eng, band, taps, &d_channel) != B200_OK) return false;
    // d_local_code_shift_chips of start_tracking (:1041-1054)
    const float spc = static_cast<float>(sig.code_samples_per_chip);
    float shifts[5];
    if (sig.veml)
        {
            shifts[0] = -conf.very_early_late_space_chips * spc;
            shifts[1] = -conf.early_late_space_chips * spc;
            shifts[2] = 0.0F;
            shifts[3] = conf.early_late_space_chips * spc;
            shifts[4] = conf.very_early_late_space_chips * spc;
        }
    else
        {
            shifts[0] = -conf.early_late_space_chips * spc;
            shifts[1] = 0.0F;
            shifts[2] = conf.early_late_space_chips * spc;
        }
    if (b200_trk_channel_set_code(eng, d_channel, sig.code_samples_per_chip * static_cast<int>(sig.code_length_chips), tracking_code, shifts, 0) != B200_OK)
        return false;
    b200_trk_loop_conf c{};
    c.fs_in = conf.fs_in;
    c.code_chip_rate = sig.code_chip_rate;
    c.signal_carrier_freq = sig.signal_carrier_freq;
    c.code_period = sig.code_period;
    c.carrier_lock_th = conf.carrier_lock_th;
    c.code_length_chips = sig.code_length_chips;
    // the adapters set vector_length = round(fs_in / (code_chip_rate / code_length_chips)) (gps_l1_ca_dll_pll_tracking.cc:59)
    c.vector_length = conf.vector_length != 0U ? conf.vector_length
                                               : static_cast<uint32_t>(std::round(conf.fs_in / (sig.code_chip_rate / static_cast<double>(sig.code_length_chips))));
    c.pull_in_time_s = conf.pull_in_time_s;
    c.bit_synchronization_time_limit_s = conf.bit_synchronization_time_limit_s;
    c.prn = sig.prn;
    c.code_samples_per_chip = sig.code_samples_per_chip;
    c.pll_filter_order = conf.pll_filter_order;
    c.dll_filter_order = conf.dll_filter_order;
    c.cn0_samples = conf.cn0_samples;
    c.cn0_min = conf.cn0_min;
    c.max_code_lock_fail = conf.max_code_lock_fail;
    c.max_carrier_lock_fail = conf.max_carrier_lock_fail;
    c.cn0_smoother_samples = conf.cn0_smoother_samples;
    c.carrier_lock_test_smoother_samples = conf.carrier_lock_test_smoother_samples;
    c.veml = sig.veml ? 1 : 0;
    c.cloop = 1;  // d_cloop = true in start_tracking (:1072)
    c.carrier_aiding = conf.carrier_aiding ? 1 : 0;
    c.enable_fll_pull_in = conf.enable_fll_pull_in ? 1 : 0;
    c.enable_fll_steady_state = conf.enable_fll_steady_state ? 1 : 0;
    c.pll_bw_hz = conf.pll_bw_hz;
    c.dll_bw_hz = conf.dll_bw_hz;
    c.fll_bw_hz = conf.fll_bw_hz;
    c.early_late_space_chips = conf.early_late_space_chips;
    c.slope = conf.slope;
    c.y_intercept = conf.y_intercept;
    c.cn0_smoother_alpha = conf.cn0_smoother_alpha;
    c.carrier_lock_test_smoother_alpha = conf.carrier_lock_test_smoother_alpha;
    if (b200_trk_loop_create(eng, d_channel, &c, &d_loop) != B200_OK) return false;
    if (d_loop + 1 > g_n_loops) g_n_loops = d_loop + 1;
    return true;
}

bool B200_Dll_Pll_Veml_Loop::start_tracking(double acq_delay_samples, double acq_doppler_hz, uint64_t acq_samplestamp_samples, uint64_t nitems_read)
{
    b200_engine* eng = shared_engine();
    if (eng == nullptr || d_loop < 0) return false;
    return b200_trk_loop_start(eng, d_loop, acq_delay_samples, acq_doppler_hz, acq_samplestamp_samples, nitems_read) == B200_OK;
}

bool B200_Dll_Pll_Veml_Loop::status(b200_trk_loop_status* out) const
{
    b200_engine* eng = shared_engine();
    if (eng == nullptr || d_loop < 0 || out == nullptr) return false;
    return b200_trk_loop_status_get(eng, d_loop, out) == B200_OK;
}

bool B200_Dll_Pll_Veml_Loop::run(int max_epochs, std::vector<b200_trk_dump_record>* records)
{
    b200_engine* eng = shared_engine();
    if (eng == nullptr || d_loop < 0 || max_epochs < 1 || records == nullptr) return false;
    std::vector<b200_trk_dump_record> all(static_cast<size_t>(g_n_loops) * max_epochs);
    std::vector<int> counts(g_n_loops, 0);
    if (b200_trk_loop_run(eng, max_epochs, all.data(), counts.data()) != B200_OK) return false;
    records->assign(all.begin() + static_cast<size_t>(d_loop) * max_epochs, all.begin() + static_cast<size_t>(d_loop) * max_epochs + counts[d_loop]);
    return true;
}
}  // namespace b200
