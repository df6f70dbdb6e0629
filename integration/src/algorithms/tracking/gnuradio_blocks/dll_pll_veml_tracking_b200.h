/*!
 * \file dll_pll_veml_tracking_b200.h
 * \brief Code DLL + carrier PLL tracking block whose correlations run on a B200 GPU (libb200gnss.so).
 *
 * Drop-in sibling of dll_pll_veml_tracking (same directory in the gnss-sdr tree): same stream signature
 * (in: gr_complex, out: Gnss_Synchro), same message ports ("events" out, "telemetry_to_trk" in), same
 * Dll_Pll_Conf, same public methods, same per-epoch arithmetic through the shared tracking libraries
 * (discriminators, loop filters, lock detectors, smoothers, HistogramBitSynchronizer).  What differs is where
 * the Early/Prompt/Late correlations are computed: every block of the receiver posts its epoch to a per-process
 * coalescer; the samples cross PCIe once per band, the epochs of all channels that are due share one launch.
 *
 * Configuration (same role as the reference block, plus):
 *   Tracking_XX.b200_band=0        index of the sample stream (one per RF band / signal conditioner)
 *   Tracking_XX.b200_coalesce=true false: one synchronous launch per block and epoch (debugging)
 * Build: needs -DB200_GPU_ACCEL (see integration/cmake and the factory patch).
 */
#ifndef GNSS_SDR_DLL_PLL_VEML_TRACKING_B200_H
#define GNSS_SDR_DLL_PLL_VEML_TRACKING_B200_H

#include "b200_multicorrelator_real_codes.h"
#include "bit_synchronizer.h"
#include "dll_pll_conf.h"
#include "exponential_smoother.h"
#include "gnss_block_interface.h"
#include "gnss_time.h"
#include "tow_to_trk.h"
#include "tracking_FLL_PLL_filter.h"
#include "tracking_loop_filter.h"
#include <gnuradio/block.h>
#include <gnuradio/gr_complex.h>
#include <gnuradio/types.h>
#include <pmt/pmt.h>
#include <volk_gnsssdr/volk_gnsssdr_alloc.h>
#include <cstdint>
#include <deque>
#include <fstream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

class Gnss_Synchro;
class dll_pll_veml_tracking_b200;

using dll_pll_veml_tracking_b200_sptr = gnss_shared_ptr<dll_pll_veml_tracking_b200>;

dll_pll_veml_tracking_b200_sptr dll_pll_veml_make_tracking_b200(const Dll_Pll_Conf &conf_, int b200_band = 0, bool b200_coalesce = true);

class dll_pll_veml_tracking_b200 : public gr::block
{
public:
    ~dll_pll_veml_tracking_b200() override;

    void set_channel(uint32_t channel);
    void set_gnss_synchro(Gnss_Synchro *p_gnss_synchro);
    void start_tracking();
    void stop_tracking();

    int general_work(int noutput_items, gr_vector_int &ninput_items,
        gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) override;
    void forecast(int noutput_items, gr_vector_int &ninput_items_required) override;

private:
    friend dll_pll_veml_tracking_b200_sptr dll_pll_veml_make_tracking_b200(const Dll_Pll_Conf &conf_, int b200_band, bool b200_coalesce);
    dll_pll_veml_tracking_b200(const Dll_Pll_Conf &conf_, int b200_band, bool b200_coalesce);

    // what distinguishes one GNSS signal from another for this block
    struct SignalPlan
    {
        bool known{false};
        double carrier_hz{0.0};
        double code_period_s{0.0};
        double chip_rate_cps{0.0};
        int32_t code_length_chips{0};
        int32_t symbols_per_bit{0};
        int32_t correlation_length_ms{1};
        int32_t samples_per_chip{0};
        bool veml{false};
        bool has_secondary{false};
        bool swap_iq{false};
        std::string system_name;
        std::string pretty_name;
        std::string secondary_code;       // symbols removed from the tracked (pilot or data) component
        std::string data_secondary_code;  // symbols removed from the data prompt
    };
    static SignalPlan plan_for(char system, const std::string &signal, bool &track_pilot);
    void load_codes(uint32_t prn);

    enum State : int32_t
    {
        STANDBY = 0,
        PULL_IN = 1,
        WIDE_TRACKING = 2,
        EXTENDED_INTEGRATION = 3,
        NARROW_TRACKING = 4
    };
    struct Taps  // complex correlator sums of one loop update (VE, E, P, L, VL)
    {
        gr_complex ve{0, 0}, e{0, 0}, p{0, 0}, l{0, 0}, vl{0, 0};
        void clear() { ve = e = p = l = vl = gr_complex(0, 0); }
    };

    void on_telemetry_message(const pmt::pmt_t &msg);
    void reset_loop_state();
    void setup_bit_synchronizer();
    void set_tap_spacing(float early_late_chips, float very_early_late_chips);
    bool correlate_epoch(const gr_complex *in, int n_available);
    void fold_epoch_into_symbol();
    bool lock_still_held(double coherent_time_s);
    void close_loops();
    void advance_ncos();
    bool secondary_code_found();
    bool wide_tracking_sync_done();
    void enter_extended_or_narrow();
    void fill_symbol_output(Gnss_Synchro &out);
    void write_dump_record();
    void estimate_tow();

    B200_Multicorrelator_Real_Codes d_correlator;       // tracked component, 3 or 5 taps
    B200_Multicorrelator_Real_Codes d_data_correlator;  // data prompt when a pilot is tracked
    const int d_b200_band;
    const bool d_b200_coalesce;

    Dll_Pll_Conf d_conf;
    SignalPlan d_plan;
    Exponential_Smoother d_cn0_smoother;
    Exponential_Smoother d_lock_test_smoother;
    Tracking_loop_filter d_dll_filter;
    Tracking_FLL_PLL_filter d_pll_filter;
    HistogramBitSynchronizer d_bit_sync;
    Gnss_Synchro *d_synchro{nullptr};  // owned by the Channel

    volk_gnsssdr::vector<float> d_tracking_code;
    volk_gnsssdr::vector<float> d_data_code;
    volk_gnsssdr::vector<float> d_tap_shift_chips;
    volk_gnsssdr::vector<gr_complex> d_tap_out;
    volk_gnsssdr::vector<gr_complex> d_data_prompt;
    std::vector<gr_complex> d_cn0_prompts;
    std::deque<gr_complex> d_sync_prompts;  // last d_sync_length prompts for the secondary-code / preamble search
    std::deque<float> d_dll_history;
    std::deque<std::pair<double, double>> d_code_step_history;
    std::deque<std::pair<double, double>> d_carr_step_history;

    const size_t d_int_type_hash{typeid(int).hash_code()};
    const size_t d_tow_type_hash{typeid(std::shared_ptr<TOW_to_trk>).hash_code()};

    // tap indices inside d_tap_out
    int d_i_ve{-1}, d_i_e{0}, d_i_p{1}, d_i_l{2}, d_i_vl{-1};
    int32_t d_n_taps{3};
    uint32_t d_sync_length{0};  // secondary-code (or preamble) length searched in WIDE_TRACKING

    Taps d_sum;               // accumulators of the running loop update
    gr_complex d_prev_prompt{0, 0};
    gr_complex d_data_sum{0, 0};

    double d_code_freq_chips{0.0};
    double d_acq_code_phase_samples{0.0};
    double d_acq_doppler_hz{0.0};
    double d_loop_time_s{0.0};  // coherent time of one loop update
    double d_carr_phase_err_hz{0.0};
    double d_carr_freq_err_hz{0.0};
    double d_carr_filt_hz{0.0};
    double d_code_err_chips{0.0};
    double d_code_filt_chips{0.0};
    double d_cfo_hz{0.0};
    double d_doppler_hz{0.0};
    double d_acc_phase_rad{0.0};
    double d_rem_code_chips{0.0};
    double d_rem_code_samples{0.0};
    double d_lock_test{1.0};
    double d_cn0_db_hz{0.0};
    double d_carr_step_rad{0.0};
    double d_carr_rate_step_rad{0.0};
    double d_code_step_chips{0.0};
    double d_code_rate_step_chips{0.0};
    float d_rem_carr_rad{0.0F};

    uint64_t d_acq_sample_stamp{0};
    uint64_t d_tow_ms{0};
    int32_t d_week{0};
    int64_t d_bit_edge_epoch{0};
    GnssTime d_last_timetag{};
    std::shared_ptr<TOW_to_trk> d_last_tow;
    uint64_t d_last_timetag_offset{0};
    bool d_timetag_pending{false};

    int32_t d_state{STANDBY};
    int32_t d_epoch_samples{0};  // samples consumed by the current epoch
    int32_t d_ext_count{0};
    int32_t d_ext_symbols{1};
    int32_t d_symbol_idx{0};
    int32_t d_data_symbol_idx{0};
    int32_t d_cn0_count{0};
    int32_t d_carrier_fail{0};
    int32_t d_code_fail{0};
    uint32_t d_channel{0};

    bool d_pull_in{true};
    bool d_doppler_corrected{false};
    bool d_costas{true};
    bool d_acc_phase_started{false};
    bool d_extended{false};
    bool d_phase_180{false};
    bool d_use_hist_sync{false};
    bool d_waiting_bit_edge{false};
    bool d_dump{false};
    bool d_gpu_ok{true};

    std::string d_dump_basename;
    std::ofstream d_dump_file;
};

#endif  // GNSS_SDR_DLL_PLL_VEML_TRACKING_B200_H
