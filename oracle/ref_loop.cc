// TEST INFRASTRUCTURE ONLY - oracle "ref" for the DLL/PLL loop.  Links the reference's OWN library code,
// compiled where it lies under /root/reference (oracle/Makefile, target ref):
//   tracking_discriminators.cc, tracking_FLL_PLL_filter.cc, tracking_loop_filter.cc, lock_detectors.cc,
//   exponential_smoother.cc   (src/algorithms/tracking/libs/)
// The gr::block that calls them (dll_pll_veml_tracking.cc) needs GNU Radio and cannot be built here, so the
// order of calls below restates general_work cases 1-2 (:1898-2015,:2292), cn0_and_tracking_lock_status
// (:1167-1224), run_dll_pll (:1260-1347), update_tracking_vars (:1409-1483) and log_data (:1599-1694) around
// the reference's real Tracking_loop_filter / Tracking_FLL_PLL_filter / Exponential_Smoother objects and
// discriminator / lock-detector functions.  Used to pin oracle/port_loop.c (tests/test_oracle_loop.py) and to
// generate tests/golden/loop_ref_golden.npz.  Nothing in the product path links or loads this file.
#include "b200gnss.h"
#include "exponential_smoother.h"
#include "lock_detectors.h"
#include "tracking_FLL_PLL_filter.h"
#include "tracking_discriminators.h"
#include "tracking_loop_filter.h"

#include <cmath>
#include <cstdint>
#include <vector>

namespace
{
constexpr double TWO_PI_ = 6.283185307179586;

struct RefLoop
{
    b200_trk_loop_conf c{};
    Exponential_Smoother d_cn0_smoother;
    Exponential_Smoother d_carrier_lock_test_smoother;
    Tracking_loop_filter d_code_loop_filter;
    Tracking_FLL_PLL_filter d_carrier_loop_filter;
    std::vector<gr_complex> d_Prompt_buffer;
    double d_acq_code_phase_samples{}, d_acq_carrier_doppler_hz{}, d_current_correlation_time_s{};
    double d_carr_phase_error_hz{}, d_carr_freq_error_hz{}, d_carr_error_filt_hz{}, d_code_error_chips{}, d_code_error_filt_chips{};
    double d_code_freq_chips{}, d_carrier_doppler_hz{}, d_acc_carrier_phase_rad{}, d_rem_code_phase_chips{};
    double d_T_chip_seconds{}, d_T_prn_seconds{}, d_T_prn_samples{}, d_K_blk_samples{};
    double d_carrier_lock_test{}, d_CN0_SNV_dB_Hz{}, d_carrier_lock_threshold{};
    double d_carrier_phase_step_rad{}, d_carrier_phase_rate_step_rad{}, d_code_phase_step_chips{}, d_code_phase_rate_step_chips{};
    double d_rem_code_phase_samples{};
    gr_complex d_VE_accu, d_E_accu, d_P_accu, d_P_accu_old, d_L_accu, d_VL_accu, d_Prompt;
    uint64_t d_acq_sample_stamp{}, nitems_read{};
    float d_rem_carr_phase_rad{}, spc{};
    int32_t d_state{}, d_current_prn_length_samples{}, d_cn0_estimation_counter{}, d_carrier_lock_fail_counter{}, d_code_lock_fail_counter{};
    bool d_pull_in_transitory{true}, d_cloop{true};
    int loss_of_lock{};
    uint64_t epochs{};

    void clear_tracking_vars()
    {
        d_P_accu_old = gr_complex(0.0, 0.0);
        d_carr_phase_error_hz = 0.0;
        d_carr_freq_error_hz = 0.0;
        d_carr_error_filt_hz = 0.0;
        d_code_error_chips = 0.0;
        d_code_error_filt_chips = 0.0;
        d_carrier_phase_rate_step_rad = 0.0;
        d_code_phase_rate_step_chips = 0.0;
    }

    void pull_in_check()
    {
        if (d_pull_in_transitory == true)
            {
                if (c.pull_in_time_s < (nitems_read - d_acq_sample_stamp) / static_cast<int>(c.fs_in))
                    {
                        d_pull_in_transitory = false;
                        d_carrier_lock_fail_counter = 0;
                        d_code_lock_fail_counter = 0;
                    }
            }
    }

    bool lock_status(double coh_integration_time_s)
    {
        if (d_cn0_estimation_counter < c.cn0_samples)
            {
                d_Prompt_buffer[d_cn0_estimation_counter] = d_P_accu;
                d_cn0_estimation_counter++;
                return true;
            }
        d_Prompt_buffer[d_cn0_estimation_counter % c.cn0_samples] = d_P_accu;
        d_cn0_estimation_counter++;
        const float raw = cn0_m2m4_estimator(d_Prompt_buffer.data(), c.cn0_samples, static_cast<float>(coh_integration_time_s));
        d_CN0_SNV_dB_Hz = d_cn0_smoother.smooth(raw);
        d_carrier_lock_test = d_carrier_lock_test_smoother.smooth(carrier_lock_detector(d_Prompt_buffer.data(), 1));
        if (!d_pull_in_transitory)
            {
                if (d_carrier_lock_test < d_carrier_lock_threshold)
                    d_carrier_lock_fail_counter++;
                else if (d_carrier_lock_fail_counter > 0)
                    d_carrier_lock_fail_counter--;
                if (d_CN0_SNV_dB_Hz < c.cn0_min)
                    d_code_lock_fail_counter++;
                else if (d_code_lock_fail_counter > 0)
                    d_code_lock_fail_counter--;
            }
        if (d_carrier_lock_fail_counter > c.max_carrier_lock_fail || d_code_lock_fail_counter > c.max_code_lock_fail)
            {
                d_carrier_lock_fail_counter = 0;
                d_code_lock_fail_counter = 0;
                return false;
            }
        return true;
    }

    void run_dll_pll()
    {
        if (d_cloop)
            d_carr_phase_error_hz = pll_cloop_two_quadrant_atan(d_P_accu) / TWO_PI_;
        else
            d_carr_phase_error_hz = pll_four_quadrant_atan(d_P_accu) / TWO_PI_;
        if ((d_pull_in_transitory == true && c.enable_fll_pull_in) || c.enable_fll_steady_state)
            {
                d_carr_freq_error_hz = fll_diff_atan(d_P_accu_old, d_P_accu, 0, d_current_correlation_time_s) / TWO_PI_;
                d_P_accu_old = d_P_accu;
                if (d_pull_in_transitory == true && c.enable_fll_pull_in)
                    d_carr_error_filt_hz = d_carrier_loop_filter.get_carrier_error(static_cast<float>(d_carr_freq_error_hz), 0.0F, static_cast<float>(d_current_correlation_time_s));
                else
                    d_carr_error_filt_hz = d_carrier_loop_filter.get_carrier_error(static_cast<float>(d_carr_freq_error_hz), static_cast<float>(d_carr_phase_error_hz), static_cast<float>(d_current_correlation_time_s));
            }
        else
            {
                d_carr_error_filt_hz = d_carrier_loop_filter.get_carrier_error(0, static_cast<float>(d_carr_phase_error_hz), static_cast<float>(d_current_correlation_time_s));
            }
        d_carrier_doppler_hz = d_carr_error_filt_hz;
        if (c.veml)
            d_code_error_chips = dll_nc_vemlp_normalized(d_VE_accu, d_E_accu, d_L_accu, d_VL_accu);
        else
            d_code_error_chips = dll_nc_e_minus_l_normalized(d_E_accu, d_L_accu, spc, c.slope, c.y_intercept);
        d_code_error_filt_chips = d_code_loop_filter.apply(static_cast<float>(d_code_error_chips));
        d_code_freq_chips = c.code_chip_rate - d_code_error_filt_chips;
        if (c.carrier_aiding) d_code_freq_chips += d_carrier_doppler_hz * c.code_chip_rate / c.signal_carrier_freq;
    }

    void update_tracking_vars()
    {
        d_T_chip_seconds = 1.0 / d_code_freq_chips;
        d_T_prn_seconds = d_T_chip_seconds * static_cast<double>(static_cast<int32_t>(c.code_length_chips));
        d_T_prn_samples = d_T_prn_seconds * c.fs_in;
        d_K_blk_samples = d_T_prn_samples + d_rem_code_phase_samples;
        d_current_prn_length_samples = static_cast<int32_t>(std::floor(d_K_blk_samples));
        d_carrier_phase_step_rad = TWO_PI_ * (d_carrier_doppler_hz + 0.0) / c.fs_in;
        d_rem_carr_phase_rad += static_cast<float>(d_carrier_phase_step_rad * static_cast<double>(d_current_prn_length_samples) + 0.5 * d_carrier_phase_rate_step_rad * static_cast<double>(d_current_prn_length_samples) * static_cast<double>(d_current_prn_length_samples));
        d_rem_carr_phase_rad = fmod(d_rem_carr_phase_rad, TWO_PI_);
        d_acc_carrier_phase_rad -= (d_carrier_phase_step_rad * static_cast<double>(d_current_prn_length_samples) + 0.5 * d_carrier_phase_rate_step_rad * static_cast<double>(d_current_prn_length_samples) * static_cast<double>(d_current_prn_length_samples));
        d_code_phase_step_chips = d_code_freq_chips / c.fs_in;
        d_rem_code_phase_samples = d_K_blk_samples - static_cast<double>(d_current_prn_length_samples);
        d_rem_code_phase_chips = d_code_freq_chips * d_rem_code_phase_samples / c.fs_in;
    }

    void log_data(b200_trk_dump_record* r) const
    {
        r->abs_VE = c.veml ? std::abs<float>(d_VE_accu) : 0.0F;
        r->abs_E = std::abs<float>(d_E_accu);
        r->abs_P = std::abs<float>(d_P_accu);
        r->abs_L = std::abs<float>(d_L_accu);
        r->abs_VL = c.veml ? std::abs<float>(d_VL_accu) : 0.0F;
        r->prompt_I = d_Prompt.real();
        r->prompt_Q = d_Prompt.imag();
        r->PRN_start_sample_count = nitems_read + static_cast<uint64_t>(d_current_prn_length_samples);
        r->acc_carrier_phase_rad = static_cast<float>(d_acc_carrier_phase_rad);
        r->carrier_doppler_hz = static_cast<float>(d_carrier_doppler_hz);
        r->carrier_doppler_rate_hz_s = static_cast<float>(d_carrier_phase_rate_step_rad * c.fs_in * c.fs_in / TWO_PI_);
        r->code_freq_chips = static_cast<float>(d_code_freq_chips);
        r->code_freq_rate_chips = static_cast<float>(d_code_phase_rate_step_chips * c.fs_in * c.fs_in);
        r->carr_error_hz = static_cast<float>(d_carr_phase_error_hz);
        r->carr_error_filt_hz = static_cast<float>(d_carr_error_filt_hz);
        r->code_error_chips = static_cast<float>(d_code_error_chips);
        r->code_error_filt_chips = static_cast<float>(d_code_error_filt_chips);
        r->CN0_SNV_dB_Hz = static_cast<float>(d_CN0_SNV_dB_Hz);
        r->carrier_lock_test = static_cast<float>(d_carrier_lock_test);
        r->aux1 = static_cast<float>(d_rem_code_phase_samples);
        r->aux2 = static_cast<double>(nitems_read + d_current_prn_length_samples);
        r->PRN = c.prn;
        r->TOW_ms = 0;
        r->WN = 0;
    }
};
}  // namespace

extern "C"
{
    void* ref_loop_create(const b200_trk_loop_conf* conf)
    {
        if (conf->cn0_samples < 1) return nullptr;
        auto* L = new RefLoop();
        L->c = *conf;
        L->d_carrier_lock_threshold = conf->carrier_lock_th;
        L->d_code_freq_chips = conf->code_chip_rate;
        L->d_code_loop_filter = Tracking_loop_filter(static_cast<float>(conf->code_period), conf->dll_bw_hz, conf->dll_filter_order, false);
        L->d_carrier_loop_filter.set_params(conf->fll_bw_hz, conf->pll_bw_hz, conf->pll_filter_order);
        L->d_Prompt_buffer.assign(conf->cn0_samples, gr_complex(0, 0));
        L->d_cn0_smoother = Exponential_Smoother();
        L->d_cn0_smoother.set_alpha(conf->cn0_smoother_alpha);
        if (conf->code_period > 0.0) L->d_cn0_smoother.set_samples_for_initialization(conf->cn0_smoother_samples / static_cast<int>(conf->code_period * 1000.0));
        L->d_carrier_lock_test_smoother = Exponential_Smoother();
        L->d_carrier_lock_test_smoother.set_alpha(conf->carrier_lock_test_smoother_alpha);
        L->d_carrier_lock_test_smoother.set_min_value(-1.0);
        L->d_carrier_lock_test_smoother.set_offset(0.0);
        L->d_carrier_lock_test_smoother.set_samples_for_initialization(conf->carrier_lock_test_smoother_samples);
        L->spc = conf->early_late_space_chips;
        L->clear_tracking_vars();
        return L;
    }

    void ref_loop_destroy(void* h) { delete static_cast<RefLoop*>(h); }

    void ref_loop_start(void* h, double acq_delay_samples, double acq_doppler_hz, uint64_t acq_samplestamp, uint64_t nitems_read)
    {
        auto* L = static_cast<RefLoop*>(h);
        const auto& c = L->c;
        L->d_acq_code_phase_samples = acq_delay_samples;
        L->d_acq_carrier_doppler_hz = acq_doppler_hz;
        L->d_acq_sample_stamp = acq_samplestamp;
        L->nitems_read = nitems_read;
        L->d_carrier_doppler_hz = L->d_acq_carrier_doppler_hz;
        L->d_carrier_phase_step_rad = TWO_PI_ * L->d_carrier_doppler_hz / c.fs_in;
        L->d_carrier_phase_rate_step_rad = 0.0;
        L->d_carrier_lock_fail_counter = 0;
        L->d_code_lock_fail_counter = 0;
        L->d_rem_code_phase_samples = 0.0;
        L->d_rem_carr_phase_rad = 0.0;
        L->d_rem_code_phase_chips = 0.0;
        L->d_acc_carrier_phase_rad = 0.0;
        L->d_cn0_estimation_counter = 0;
        L->d_carrier_lock_test = 1.0;
        L->d_CN0_SNV_dB_Hz = 0.0;
        L->d_current_correlation_time_s = c.code_period;
        L->d_carrier_loop_filter.set_params(c.fll_bw_hz, c.pll_bw_hz, c.pll_filter_order);
        L->d_code_loop_filter.set_noise_bandwidth(c.dll_bw_hz);
        L->d_code_loop_filter.set_update_interval(static_cast<float>(c.code_period));
        L->d_carrier_loop_filter.initialize(static_cast<float>(L->d_acq_carrier_doppler_hz));
        L->d_code_loop_filter.initialize();
        L->d_state = 1;
        L->d_cloop = c.cloop != 0;
        L->d_pull_in_transitory = true;
        L->loss_of_lock = 0;
        L->epochs = 0;
    }

    int ref_loop_prepare(void* h, uint64_t* sample_index, int32_t* n, float* p6)
    {
        auto* L = static_cast<RefLoop*>(h);
        const auto& c = L->c;
        if (L->d_state == 0) return 0;
        L->pull_in_check();
        if (L->d_state == 1)
            {
                const int64_t acq_trk_diff_samples = static_cast<int64_t>(L->nitems_read) - static_cast<int64_t>(L->d_acq_sample_stamp);
                const double delta_trk_to_acq_prn_start_samples = static_cast<double>(acq_trk_diff_samples) - L->d_acq_code_phase_samples;
                L->d_code_freq_chips = c.code_chip_rate;
                L->d_code_phase_step_chips = L->d_code_freq_chips / c.fs_in;
                L->d_code_phase_rate_step_chips = 0.0;
                const double T_chip_mod_seconds = 1.0 / L->d_code_freq_chips;
                const double T_prn_mod_seconds = T_chip_mod_seconds * static_cast<double>(c.code_length_chips);
                const double T_prn_mod_samples = T_prn_mod_seconds * c.fs_in;
                L->d_acq_code_phase_samples = T_prn_mod_samples - std::fmod(delta_trk_to_acq_prn_start_samples, T_prn_mod_samples);
                L->d_current_prn_length_samples = round(T_prn_mod_samples);
                const int32_t samples_offset = round(L->d_acq_code_phase_samples);
                L->d_acc_carrier_phase_rad -= L->d_carrier_phase_step_rad * static_cast<double>(samples_offset);
                L->d_state = 2;
                L->d_cn0_smoother.reset();
                L->d_carrier_lock_test_smoother.reset();
                L->nitems_read += static_cast<uint64_t>(static_cast<int64_t>(samples_offset));
                L->pull_in_check();
            }
        *sample_index = L->nitems_read;
        *n = static_cast<int32_t>(c.vector_length);
        p6[0] = L->d_rem_carr_phase_rad;
        p6[1] = static_cast<float>(L->d_carrier_phase_step_rad);
        p6[2] = static_cast<float>(L->d_carrier_phase_rate_step_rad);
        p6[3] = static_cast<float>(L->d_rem_code_phase_chips) * static_cast<float>(c.code_samples_per_chip);
        p6[4] = static_cast<float>(L->d_code_phase_step_chips) * static_cast<float>(c.code_samples_per_chip);
        p6[5] = static_cast<float>(L->d_code_phase_rate_step_chips) * static_cast<float>(c.code_samples_per_chip);
        return 1;
    }

    int ref_loop_update(void* h, const float* taps, b200_trk_dump_record* rec)
    {
        auto* L = static_cast<RefLoop*>(h);
        const auto& c = L->c;
        const auto* t = reinterpret_cast<const gr_complex*>(taps);
        if (L->d_state != 2) return 0;
        if (c.veml)
            {
                L->d_VE_accu = t[0];
                L->d_E_accu = t[1];
                L->d_P_accu = t[2];
                L->d_L_accu = t[3];
                L->d_VL_accu = t[4];
            }
        else
            {
                L->d_E_accu = t[0];
                L->d_P_accu = t[1];
                L->d_L_accu = t[2];
            }
        L->d_Prompt = L->d_P_accu;
        L->spc = c.early_late_space_chips;
        int logged = 0;
        if (c.bit_synchronization_time_limit_s < (L->nitems_read - L->d_acq_sample_stamp) / static_cast<int>(c.fs_in)) L->d_carrier_lock_fail_counter = 300000;
        if (!L->lock_status(c.code_period))
            {
                L->clear_tracking_vars();
                L->d_state = 0;
                L->loss_of_lock = 1;
            }
        else
            {
                L->run_dll_pll();
                L->update_tracking_vars();
                if (rec) L->log_data(rec);
                logged = 1;
                L->epochs++;
            }
        L->nitems_read += static_cast<uint64_t>(static_cast<int64_t>(L->d_current_prn_length_samples));
        return logged;
    }

    void ref_loop_status(const void* h, b200_trk_loop_status* s)
    {
        const auto* L = static_cast<const RefLoop*>(h);
        s->state = L->d_state;
        s->loss_of_lock = L->loss_of_lock;
        s->sample_counter = L->nitems_read;
        s->epochs = L->epochs;
        s->carrier_doppler_hz = L->d_carrier_doppler_hz;
        s->code_freq_chips = L->d_code_freq_chips;
        s->rem_code_phase_samples = L->d_rem_code_phase_samples;
        s->acc_carrier_phase_rad = L->d_acc_carrier_phase_rad;
        s->CN0_SNV_dB_Hz = L->d_CN0_SNV_dB_Hz;
        s->carrier_lock_test = L->d_carrier_lock_test;
    }

    double ref_disc_pll_cloop(float re, float im) { return pll_cloop_two_quadrant_atan(gr_complex(re, im)); }
    double ref_disc_fll_diff_atan(float re1, float im1, float re2, float im2, double t1, double t2)
    {
        return fll_diff_atan(gr_complex(re1, im1), gr_complex(re2, im2), t1, t2);
    }
    double ref_disc_dll_e_minus_l(float er, float ei, float lr, float li, float spc, float slope, float y_intercept)
    {
        return dll_nc_e_minus_l_normalized(gr_complex(er, ei), gr_complex(lr, li), spc, slope, y_intercept);
    }
    double ref_disc_dll_vemlp(const float* t8)
    {
        const auto* t = reinterpret_cast<const gr_complex*>(t8);
        return dll_nc_vemlp_normalized(t[0], t[1], t[2], t[3]);
    }
    float ref_cn0_m2m4(const float* buf, int length, float T) { return cn0_m2m4_estimator(reinterpret_cast<const gr_complex*>(buf), length, T); }
    float ref_carrier_lock_detector(const float* buf, int length) { return carrier_lock_detector(reinterpret_cast<const gr_complex*>(buf), length); }
}
