/*!
 * \file dll_pll_veml_tracking_b200.cc
 * \brief see dll_pll_veml_tracking_b200.h.
 *
 * Behavioural contract: for the same Dll_Pll_Conf and the same samples this block emits the Gnss_Synchro stream,
 * the "events" messages and the dump records of dll_pll_veml_tracking
 * (src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc; state machine :1898-2327, loop arithmetic
 * :1167-1224,:1260-1483, symbol accumulation :1486-1596, dump record :1599-1694), within the float tolerance of
 * the correlator outputs.  tests/test_integration_blocks.py runs both blocks side by side (the reference block is
 * compiled where it lies, oracle/Makefile `blocks`).
 */
#include "dll_pll_veml_tracking_b200.h"
#include "GPS_L1_CA.h"
#include "GPS_L2C.h"
#include "GPS_L5.h"
#include "Galileo_E1.h"
#include "Galileo_E5a.h"
#include "MATH_CONSTANTS.h"
#include "galileo_e1_signal_replica.h"
#include "galileo_e5_signal_replica.h"
#include "gnss_satellite.h"
#include "gnss_sdr_create_directory.h"
#include "gnss_sdr_filesystem.h"
#include "gnss_synchro.h"
#include "gps_l2c_signal_replica.h"
#include "gps_l5_signal_replica.h"
#include "gps_sdr_signal_replica.h"
#include "lock_detectors.h"
#include "tracking_discriminators.h"
#include <gnuradio/io_signature.h>
#include <gnuradio/thread/thread.h>
#include <pmt/pmt_sugar.h>
#include <algorithm>
#include <any>
#include <array>
#include <cmath>
#include <exception>
#include <iostream>
#include <numeric>

#if USE_GLOG_AND_GFLAGS
#include <glog/logging.h>
#else
#include <absl/log/log.h>
#endif

namespace
{
template <typename Q, typename V>
void push_bounded(Q &q, size_t capacity, V &&v)
{
    if (capacity == 0) return;
    if (q.size() == capacity) q.pop_front();
    q.push_back(std::forward<V>(v));
}
}  // namespace


dll_pll_veml_tracking_b200_sptr dll_pll_veml_make_tracking_b200(const Dll_Pll_Conf &conf_, int b200_band, bool b200_coalesce)
{
    return dll_pll_veml_tracking_b200_sptr(new dll_pll_veml_tracking_b200(conf_, b200_band, b200_coalesce));
}


// Signal table.  Constants are the reference's own (system_parameters/*.h); the selection mirrors the
// constructor of dll_pll_veml_tracking (:160-597) for the signals the B200 adapters offer.
dll_pll_veml_tracking_b200::SignalPlan dll_pll_veml_tracking_b200::plan_for(char system, const std::string &signal, bool &track_pilot)
{
    SignalPlan p;
    if (system == 'G' && signal == "1C")
        {
            p.known = true;
            p.system_name = "GPS";
            p.pretty_name = "L1 C/A";
            p.carrier_hz = GPS_L1_FREQ_HZ;
            p.code_period_s = GPS_L1_CA_CODE_PERIOD_S;
            p.chip_rate_cps = GPS_L1_CA_CODE_RATE_CPS;
            p.code_length_chips = static_cast<int32_t>(GPS_L1_CA_CODE_LENGTH_CHIPS);
            p.correlation_length_ms = 1;
            p.samples_per_chip = 1;
            track_pilot = false;  // no pilot component
            // bit synchronisation: the telemetry preamble plays the secondary code's part
            p.secondary_code = GPS_CA_PREAMBLE_SYMBOLS_STR;
            p.symbols_per_bit = GPS_CA_TELEMETRY_SYMBOLS_PER_BIT;
        }
    else if (system == 'G' && signal == "2S")
        {
            p.known = true;
            p.system_name = "GPS";
            p.pretty_name = "L2C";
            p.carrier_hz = GPS_L2_FREQ_HZ;
            p.code_period_s = GPS_L2_M_PERIOD_S;
            p.chip_rate_cps = GPS_L2_M_CODE_RATE_CPS;
            p.code_length_chips = static_cast<int32_t>(GPS_L2_M_CODE_LENGTH_CHIPS);
            p.symbols_per_bit = GPS_L2_SAMPLES_PER_SYMBOL;
            p.correlation_length_ms = 20;
            p.samples_per_chip = 2;
            track_pilot = false;
        }
    else if (system == 'G' && signal == "L5")
        {
            p.known = true;
            p.system_name = "GPS";
            p.carrier_hz = GPS_L5_FREQ_HZ;
            p.code_period_s = GPS_L5I_PERIOD_S;
            p.chip_rate_cps = GPS_L5I_CODE_RATE_CPS;
            p.symbols_per_bit = GPS_L5_SAMPLES_PER_SYMBOL;
            p.correlation_length_ms = 1;
            p.samples_per_chip = 1;
            p.code_length_chips = static_cast<int32_t>(GPS_L5I_CODE_LENGTH_CHIPS);
            p.has_secondary = true;
            if (track_pilot)
                {
                    p.secondary_code = GPS_L5Q_NH_CODE_STR;
                    p.data_secondary_code = GPS_L5I_NH_CODE_STR;
                    p.pretty_name = "L5Q";
                }
            else
                {
                    p.secondary_code = GPS_L5I_NH_CODE_STR;
                    p.pretty_name = "L5I";
                    p.swap_iq = true;
                }
        }
    else if (system == 'E' && signal == "1B")
        {
            p.known = true;
            p.system_name = "Galileo";
            p.carrier_hz = GALILEO_E1_FREQ_HZ;
            p.code_period_s = GALILEO_E1_CODE_PERIOD_S;
            p.chip_rate_cps = GALILEO_E1_CODE_CHIP_RATE_CPS;
            p.code_length_chips = static_cast<int32_t>(GALILEO_E1_B_CODE_LENGTH_CHIPS);
            p.symbols_per_bit = 1;
            p.correlation_length_ms = 4;
            p.samples_per_chip = 2;  // sinBOC(1,1) sampled twice per chip
            p.veml = true;
            if (track_pilot)
                {
                    p.has_secondary = true;
                    p.secondary_code = GALILEO_E1_C_SECONDARY_CODE;
                    p.pretty_name = "E1C";
                }
            else
                {
                    p.pretty_name = "E1B";
                }
        }
    else if (system == 'E' && signal == "5X")
        {
            p.known = true;
            p.system_name = "Galileo";
            p.carrier_hz = GALILEO_E5A_FREQ_HZ;
            p.code_period_s = GALILEO_E5A_CODE_PERIOD_S;
            p.chip_rate_cps = GALILEO_E5A_CODE_CHIP_RATE_CPS;
            p.symbols_per_bit = 20;
            p.correlation_length_ms = 1;
            p.samples_per_chip = 1;
            p.code_length_chips = static_cast<int32_t>(GALILEO_E5A_CODE_LENGTH_CHIPS);
            p.has_secondary = true;
            if (track_pilot)
                {
                    // the pilot's secondary code depends on the PRN: filled in load_codes()
                    p.secondary_code = std::string(GALILEO_E5A_Q_SECONDARY_CODE_LENGTH, '0');
                    p.data_secondary_code = GALILEO_E5A_I_SECONDARY_CODE;
                    p.pretty_name = "E5aQ";
                    p.swap_iq = true;
                }
            else
                {
                    p.secondary_code = GALILEO_E5A_I_SECONDARY_CODE;
                    p.pretty_name = "E5aI";
                }
        }
    return p;
}


dll_pll_veml_tracking_b200::dll_pll_veml_tracking_b200(const Dll_Pll_Conf &conf_, int b200_band, bool b200_coalesce)
    : gr::block("dll_pll_veml_tracking_b200", gr::io_signature::make(1, 1, sizeof(gr_complex)),
          gr::io_signature::make(1, 1, sizeof(Gnss_Synchro))),
      d_b200_band(b200_band),
      d_b200_coalesce(b200_coalesce),
      d_conf(conf_),
      d_bit_sync(HistogramBitSynchronizer::Config())
{
#if GNURADIO_GREATER_THAN_38
    this->set_relative_rate(1, static_cast<uint64_t>(d_conf.vector_length));
#else
    this->set_relative_rate(1.0 / static_cast<double>(d_conf.vector_length));
#endif
    this->set_max_noutput_items(1);  // one symbol per call: nothing piles up in the output buffer
    this->message_port_register_out(pmt::mp("events"));
    this->message_port_register_in(pmt::mp("telemetry_to_trk"));
    this->set_msg_handler(pmt::mp("telemetry_to_trk"), [this](const pmt::pmt_t &m) { on_telemetry_message(m); });

    d_plan = plan_for(d_conf.system, std::string(d_conf.signal), d_conf.track_pilot);
    if (!d_plan.known)
        {
            LOG(WARNING) << "dll_pll_veml_tracking_b200: unsupported system/signal " << d_conf.system << "/" << d_conf.signal;
            std::cerr << "Invalid Signal argument when instantiating B200 tracking blocks\n";
        }
    d_conf.spc = d_conf.early_late_space_chips;
    if (d_plan.veml)
        {
            // sinBOC(1,1): discriminator slope and intercept from the autocorrelation model (:365-367)
            d_conf.slope = static_cast<float>(-CalculateSlopeAbs(&SinBocCorrelationFunction<1, 1>, d_conf.spc));
            d_conf.y_intercept = static_cast<float>(GetYInterceptAbs(&SinBocCorrelationFunction<1, 1>, d_conf.spc));
        }
    else
        {
            d_conf.slope = 1.0;
            d_conf.y_intercept = 1.0;
        }
    d_sync_length = static_cast<uint32_t>(d_plan.secondary_code.size());
    d_code_freq_chips = d_plan.chip_rate_cps;

    d_dll_filter = Tracking_loop_filter(static_cast<float>(d_plan.code_period_s), d_conf.dll_bw_hz, d_conf.dll_filter_order, false);
    d_pll_filter.set_params(d_conf.fll_bw_hz, d_conf.pll_bw_hz, d_conf.pll_filter_order);

    const int32_t table_len = 2 * d_plan.code_length_chips;  // room for the 2-samples-per-chip sinBOC replica
    d_tracking_code.resize(table_len, 0.0);
    d_n_taps = d_plan.veml ? 5 : 3;
    d_tap_out = volk_gnsssdr::vector<gr_complex>(d_n_taps);
    d_tap_shift_chips = volk_gnsssdr::vector<float>(d_n_taps);
    if (d_plan.veml)
        {
            d_i_ve = 0;
            d_i_e = 1;
            d_i_p = 2;
            d_i_l = 3;
            d_i_vl = 4;
        }
    std::fill(d_tap_shift_chips.begin(), d_tap_shift_chips.end(), 0.0F);
    set_tap_spacing(d_conf.early_late_space_chips, d_conf.very_early_late_space_chips);

    d_gpu_ok = d_correlator.init(static_cast<int>(2 * d_conf.vector_length), d_n_taps);
    d_correlator.set_high_dynamics_resampler(d_conf.high_dyn);
    d_extended = d_conf.extend_correlation_symbols > 1;
    if (!d_extended) d_conf.extend_correlation_symbols = 1;
    if (d_conf.track_pilot)
        {
            d_gpu_ok = d_data_correlator.init(static_cast<int>(2 * d_conf.vector_length), 1) && d_gpu_ok;
            d_data_correlator.set_high_dynamics_resampler(d_conf.high_dyn);
            d_data_code.resize(table_len, 0.0);
        }
    if (!d_gpu_ok)
        {
            LOG(ERROR) << "dll_pll_veml_tracking_b200: no usable B200 (" << d_correlator.last_error() << "); the channel will report loss of lock";
        }

    d_cn0_prompts.assign(std::max(d_conf.cn0_samples, 1), gr_complex(0, 0));
    d_data_prompt = volk_gnsssdr::vector<gr_complex>(1);
    d_cn0_smoother.set_alpha(d_conf.cn0_smoother_alpha);
    if (d_plan.code_period_s > 0.0)
        {
            d_cn0_smoother.set_samples_for_initialization(d_conf.cn0_smoother_samples / static_cast<int>(d_plan.code_period_s * 1000.0));
        }
    d_lock_test_smoother.set_alpha(d_conf.carrier_lock_test_smoother_alpha);
    d_lock_test_smoother.set_min_value(-1.0);
    d_lock_test_smoother.set_offset(0.0);
    d_lock_test_smoother.set_samples_for_initialization(d_conf.carrier_lock_test_smoother_samples);

    reset_loop_state();

    d_dump = d_conf.dump;
    if (d_dump)
        {
            // <dir>/<basename without extension><channel>.dat, default basename "trk_channel_" (:706-739)
            std::string name = d_conf.dump_filename;
            std::string dir = ".";
            const auto slash = name.find_last_of('/');
            if (slash != std::string::npos)
                {
                    dir = name.substr(0, slash);
                    name = name.substr(slash + 1);
                }
            if (name.empty()) name = "trk_channel_";
            if (name.substr(1).find_last_of('.') != std::string::npos) name = name.substr(0, name.find_last_of('.'));
            d_dump_basename = dir + fs::path::preferred_separator + name;
            if (!gnss_sdr_create_directory(dir))
                {
                    std::cerr << "GNSS-SDR cannot create dump files for the tracking block. Wrong permissions?\n";
                    d_dump = false;
                }
        }
    set_tag_propagation_policy(TPP_DONT);  // time tags are re-issued on the symbol stream in general_work
    d_last_tow = std::make_shared<TOW_to_trk>();
    d_epoch_samples = static_cast<int32_t>(d_conf.vector_length);
}


dll_pll_veml_tracking_b200::~dll_pll_veml_tracking_b200()
{
    if (d_dump_file.is_open())
        {
            try
                {
                    d_dump_file.close();
                }
            catch (const std::exception &ex)
                {
                    LOG(WARNING) << "Exception in Tracking block destructor: " << ex.what();
                }
        }
    d_data_correlator.free();
    d_correlator.free();
}


void dll_pll_veml_tracking_b200::forecast(int noutput_items, gr_vector_int &ninput_items_required)
{
    if (noutput_items != 0)
        {
            ninput_items_required[0] = static_cast<int32_t>(d_conf.vector_length) * 2;
        }
}


void dll_pll_veml_tracking_b200::set_tap_spacing(float early_late_chips, float very_early_late_chips)
{
    const auto spc = static_cast<float>(d_plan.samples_per_chip);
    if (d_plan.veml)
        {
            d_tap_shift_chips[d_i_ve] = -very_early_late_chips * spc;
            d_tap_shift_chips[d_i_vl] = very_early_late_chips * spc;
        }
    d_tap_shift_chips[d_i_e] = -early_late_chips * spc;
    d_tap_shift_chips[d_i_l] = early_late_chips * spc;
    // the correlator reads this array at every epoch (pointer semantics of the CPU class): nothing else to do
}


void dll_pll_veml_tracking_b200::on_telemetry_message(const pmt::pmt_t &msg)
{
    try
        {
            const auto &payload = pmt::any_ref(msg);
            if (payload.type().hash_code() == d_int_type_hash)
                {
                    if (std::any_cast<int>(payload) == 1)
                        {
                            DLOG(INFO) << "Telemetry fault received in ch " << d_channel;
                            gr::thread::scoped_lock lock(d_setlock);
                            d_carrier_fail = 200000;  // the next lock check declares loss of lock
                        }
                }
            if (d_conf.tow_to_trk && payload.type().hash_code() == d_tow_type_hash)
                {
                    const auto tow = std::any_cast<const std::shared_ptr<TOW_to_trk>>(payload);
                    if (tow->signal == std::string(d_conf.signal) && tow->channel == static_cast<int32_t>(d_channel) && d_synchro != nullptr &&
                        tow->prn == d_synchro->PRN)
                        {
                            d_last_tow = tow;
                        }
                }
        }
    catch (const std::exception &ex)
        {
            LOG(WARNING) << "telemetry_to_trk: unexpected message payload: " << ex.what();
        }
}


// Local replicas for the satellite in d_synchro (start_tracking :811-1029, the signals of plan_for()).
void dll_pll_veml_tracking_b200::load_codes(uint32_t prn)
{
    const std::string sig(d_conf.signal);
    std::array<char, 3> sig3{};
    std::copy_n(d_synchro->Signal, 3, sig3.begin());
    if (d_conf.system == 'G' && sig == "1C")
        {
            gps_l1_ca_code_gen_float(d_tracking_code, prn, 0);
        }
    else if (d_conf.system == 'G' && sig == "2S")
        {
            gps_l2c_m_code_gen_float_cl_zeroed(d_tracking_code, prn);
        }
    else if (d_conf.system == 'G' && sig == "L5")
        {
            if (d_conf.track_pilot)
                {
                    gps_l5q_code_gen_float(d_tracking_code, prn);
                    gps_l5i_code_gen_float(d_data_code, prn);
                }
            else
                {
                    gps_l5i_code_gen_float(d_tracking_code, prn);
                }
        }
    else if (d_conf.system == 'E' && sig == "1B")
        {
            if (d_conf.track_pilot)
                {
                    const std::array<char, 3> pilot = {{'1', 'C', '\0'}};
                    galileo_e1_code_gen_sinboc11_float(d_tracking_code, pilot, prn);
                    galileo_e1_code_gen_sinboc11_float(d_data_code, sig3, prn);
                }
            else
                {
                    galileo_e1_code_gen_sinboc11_float(d_tracking_code, sig3, prn);
                }
        }
    else if (d_conf.system == 'E' && sig == "5X")
        {
            volk_gnsssdr::vector<gr_complex> both(d_plan.code_length_chips);
            const std::array<char, 3> e5a = {{'5', 'X', '\0'}};
            galileo_e5_a_code_gen_complex_primary(both, prn, e5a);
            if (d_conf.track_pilot)
                {
                    d_plan.secondary_code = GALILEO_E5A_Q_SECONDARY_CODE[prn - 1];
                    for (int32_t i = 0; i < d_plan.code_length_chips; i++)
                        {
                            d_tracking_code[i] = both[i].imag();
                            d_data_code[i] = both[i].real();
                        }
                }
            else
                {
                    for (int32_t i = 0; i < d_plan.code_length_chips; i++) d_tracking_code[i] = both[i].real();
                }
        }
    const int32_t table_len = d_plan.samples_per_chip * d_plan.code_length_chips;
    if (d_conf.track_pilot)
        {
            d_data_prompt[0] = gr_complex(0.0, 0.0);
            // E5a keeps one value per chip in the data table whatever samples_per_chip says (:873)
            const int32_t data_len = (d_conf.system == 'E' && sig == "1B") ? table_len : d_plan.code_length_chips;
            d_data_correlator.set_local_code_and_taps(data_len, d_data_code.data(), &d_tap_shift_chips[d_i_p]);
        }
    d_correlator.set_local_code_and_taps(table_len, d_tracking_code.data(), d_tap_shift_chips.data());
}


void dll_pll_veml_tracking_b200::start_tracking()
{
    gr::thread::scoped_lock l(d_setlock);
    if (d_synchro == nullptr || !d_plan.known) return;
    d_acq_code_phase_samples = d_synchro->Acq_delay_samples;
    d_acq_doppler_hz = d_synchro->Acq_doppler_hz;
    d_acq_sample_stamp = d_synchro->Acq_samplestamp_samples;

    d_doppler_hz = d_acq_doppler_hz;
    d_carr_step_rad = TWO_PI * d_doppler_hz / d_conf.fs_in;
    d_carr_rate_step_rad = 0.0;
    d_carr_step_history.clear();
    d_code_step_history.clear();
    d_ext_symbols = d_conf.extend_correlation_symbols;

    load_codes(d_synchro->PRN);
    std::fill_n(d_tap_out.begin(), d_n_taps, gr_complex(0.0, 0.0));

    d_carrier_fail = 0;
    d_code_fail = 0;
    d_rem_code_samples = 0.0;
    d_rem_carr_rad = 0.0;
    d_rem_code_chips = 0.0;
    d_acc_phase_rad = 0.0;
    d_cn0_count = 0;
    d_lock_test = 1.0;
    d_cn0_db_hz = 0.0;

    // wide correlator spacing again: a previous run may have left the narrow one in the array the correlator reads
    set_tap_spacing(d_conf.early_late_space_chips, d_conf.very_early_late_space_chips);
    d_loop_time_s = d_plan.code_period_s;

    d_pll_filter.set_params(d_conf.fll_bw_hz, d_conf.pll_bw_hz, d_conf.pll_filter_order);
    d_dll_filter.set_noise_bandwidth(d_conf.dll_bw_hz);
    d_dll_filter.set_update_interval(static_cast<float>(d_plan.code_period_s));
    d_pll_filter.initialize(static_cast<float>(d_acq_doppler_hz));
    d_dll_filter.initialize();

    std::cout << "Tracking of " << d_plan.system_name << " " << d_plan.pretty_name << " signal started on channel " << d_channel
              << " for satellite " << Gnss_Satellite(d_plan.system_name, d_synchro->PRN) << '\n';
    DLOG(INFO) << "Starting B200 tracking of satellite " << Gnss_Satellite(d_plan.system_name, d_synchro->PRN) << " on channel " << d_channel;

    d_state = PULL_IN;
    d_costas = true;
    d_pull_in = true;
    d_sync_prompts.clear();
    d_doppler_corrected = false;
    d_acc_phase_started = false;
    setup_bit_synchronizer();
}


void dll_pll_veml_tracking_b200::stop_tracking()
{
    gr::thread::scoped_lock l(d_setlock);
    d_state = STANDBY;
    d_correlator.idle();
    if (d_conf.track_pilot) d_data_correlator.idle();
}


void dll_pll_veml_tracking_b200::set_channel(uint32_t channel)
{
    gr::thread::scoped_lock l(d_setlock);
    d_channel = channel;
    LOG(INFO) << "Tracking Channel set to " << d_channel;
    if (d_dump && !d_dump_file.is_open())
        {
            const std::string filename = d_dump_basename + std::to_string(d_channel) + ".dat";
            try
                {
                    d_dump_file.exceptions(std::ofstream::failbit | std::ofstream::badbit);
                    d_dump_file.open(filename.c_str(), std::ios::out | std::ios::binary);
                    LOG(INFO) << "Tracking dump enabled on channel " << d_channel << " Log file: " << filename;
                }
            catch (const std::ofstream::failure &e)
                {
                    LOG(WARNING) << "channel " << d_channel << " Exception opening trk dump file " << e.what();
                }
        }
}


void dll_pll_veml_tracking_b200::set_gnss_synchro(Gnss_Synchro *p_gnss_synchro)
{
    gr::thread::scoped_lock l(d_setlock);
    d_synchro = p_gnss_synchro;
}


void dll_pll_veml_tracking_b200::reset_loop_state()
{
    std::fill_n(d_tap_out.begin(), d_n_taps, gr_complex(0.0, 0.0));
    if (d_conf.track_pilot)
        {
            d_data_prompt[0] = gr_complex(0.0, 0.0);
            d_data_sum = gr_complex(0.0, 0.0);
        }
    d_prev_prompt = gr_complex(0.0, 0.0);
    d_carr_phase_err_hz = 0.0;
    d_carr_freq_err_hz = 0.0;
    d_carr_filt_hz = 0.0;
    d_code_err_chips = 0.0;
    d_code_filt_chips = 0.0;
    d_symbol_idx = 0;
    d_data_symbol_idx = 0;
    d_sync_prompts.clear();
    d_carr_rate_step_rad = 0.0;
    d_code_rate_step_chips = 0.0;
    d_tow_ms = 0ULL;
    d_week = 0;
    d_carr_step_history.clear();
    d_code_step_history.clear();
    d_bit_sync.reset();
}


void dll_pll_veml_tracking_b200::setup_bit_synchronizer()
{
    // signals without a secondary code but several code periods per bit look for the bit edge with a histogram (:1383-1406)
    d_use_hist_sync = !d_plan.has_secondary && d_plan.symbols_per_bit > 1 && d_plan.system_name != "Glonass";
    if (d_use_hist_sync)
        {
            HistogramBitSynchronizer::Config cfg;
            cfg.bit_period_ms = d_plan.symbols_per_bit * d_plan.correlation_length_ms;
            cfg.epoch_ms = d_plan.correlation_length_ms;
            cfg.min_events_for_lock = d_conf.bs_min_events_for_lock;
            cfg.stable_best_required = d_conf.bs_stable_best_required;
            cfg.dominance_ratio = d_conf.bs_dominance_ratio;
            cfg.min_prompt_mag = d_conf.bs_min_prompt_mag;
            cfg.use_phase_dot_detector = d_conf.bs_use_phase_dot_detector;
            d_bit_sync = HistogramBitSynchronizer(cfg);
        }
    d_bit_sync.reset();
}


// E/P/L (and the data prompt) of the epoch that starts at in[0] == sample nitems_read(0) of the band.
bool dll_pll_veml_tracking_b200::correlate_epoch(const gr_complex *in, int n_available)
{
    const auto spc = static_cast<float>(d_plan.samples_per_chip);
    const float rem_code = static_cast<float>(d_rem_code_chips) * spc;
    const float code_step = static_cast<float>(d_code_step_chips) * spc;
    const float code_rate = static_cast<float>(d_code_rate_step_chips) * spc;
    const auto carr_step = static_cast<float>(d_carr_step_rad);
    const auto carr_rate = static_cast<float>(d_carr_rate_step_rad);
    const int n = static_cast<int>(d_conf.vector_length);
    d_correlator.set_input_output_vectors(d_tap_out.data(), in);
    if (d_conf.track_pilot) d_data_correlator.set_input_output_vectors(d_data_prompt.data(), in);
    bool ok;
    if (d_b200_coalesce)
        {
            const uint64_t pos = this->nitems_read(0);
            // both correlators go into the same batch
            d_correlator.set_stream_position(d_b200_band, pos, n_available);
            ok = d_correlator.post(d_rem_carr_rad, carr_step, carr_rate, rem_code, code_step, code_rate, n);
            if (ok && d_conf.track_pilot)
                {
                    d_data_correlator.set_stream_position(d_b200_band, pos, n_available);
                    ok = d_data_correlator.post(d_rem_carr_rad, carr_step, carr_rate, rem_code, code_step, code_rate, n);
                    ok = d_data_correlator.wait() && ok;
                }
            ok = d_correlator.wait() && ok;
        }
    else
        {
            ok = d_correlator.Carrier_wipeoff_multicorrelator_resampler(d_rem_carr_rad, carr_step, carr_rate, rem_code, code_step, code_rate, n);
            if (ok && d_conf.track_pilot)
                {
                    ok = d_data_correlator.Carrier_wipeoff_multicorrelator_resampler(d_rem_carr_rad, carr_step, carr_rate, rem_code, code_step, code_rate, n);
                }
        }
    return ok;
}


// Add this epoch's taps to the running loop-update sums with the secondary-code sign removed, and the data prompt to
// the running symbol (:1486-1596).
void dll_pll_veml_tracking_b200::fold_epoch_into_symbol()
{
    float sign = 1.0F;
    if (d_plan.has_secondary)
        {
            if (d_plan.secondary_code[d_symbol_idx] != '0') sign = -1.0F;
            d_symbol_idx = (d_symbol_idx + 1) % static_cast<int32_t>(d_sync_length);
        }
    if (d_plan.veml)
        {
            d_sum.ve += sign * d_tap_out[d_i_ve];
            d_sum.vl += sign * d_tap_out[d_i_vl];
        }
    d_sum.e += sign * d_tap_out[d_i_e];
    d_sum.p += sign * d_tap_out[d_i_p];
    d_sum.l += sign * d_tap_out[d_i_l];

    const gr_complex data = d_conf.track_pilot ? d_data_prompt[0] : d_tap_out[d_i_p];
    if (d_plan.symbols_per_bit > 1)
        {
            if (!d_plan.data_secondary_code.empty())
                {
                    if (d_plan.data_secondary_code[d_data_symbol_idx] == '0')
                        d_data_sum += data;
                    else
                        d_data_sum -= data;
                    d_data_symbol_idx = (d_data_symbol_idx + 1) % static_cast<int32_t>(d_plan.data_secondary_code.size());
                }
            else
                {
                    d_data_sum += data;
                    d_data_symbol_idx = (d_data_symbol_idx + 1) % d_plan.symbols_per_bit;
                }
        }
    else
        {
            d_data_sum = data;
        }
    // a tracked pilot carries no data: four-quadrant discriminator from here on
    d_costas = !d_conf.track_pilot;
}


// C/N0 estimate, carrier lock test and the two fail counters; false = loss of lock (message 3 on "events").
bool dll_pll_veml_tracking_b200::lock_still_held(double coherent_time_s)
{
    const int32_t window = d_conf.cn0_samples;
    if (d_cn0_count < window)
        {
            d_cn0_prompts[d_cn0_count++] = d_sum.p;
            return true;
        }
    d_cn0_prompts[d_cn0_count % window] = d_sum.p;
    d_cn0_count++;
    d_cn0_db_hz = d_cn0_smoother.smooth(cn0_m2m4_estimator(d_cn0_prompts.data(), window, static_cast<float>(coherent_time_s)));
    d_lock_test = d_lock_test_smoother.smooth(carrier_lock_detector(d_cn0_prompts.data(), 1));
    if (!d_pull_in)
        {
            if (d_lock_test < d_conf.carrier_lock_th)
                d_carrier_fail++;
            else if (d_carrier_fail > 0)
                d_carrier_fail--;
            if (d_cn0_db_hz < d_conf.cn0_min)
                d_code_fail++;
            else if (d_code_fail > 0)
                d_code_fail--;
        }
    if (d_carrier_fail > d_conf.max_carrier_lock_fail || d_code_fail > d_conf.max_code_lock_fail)
        {
            std::cout << "Loss of lock in channel " << d_channel << ", satellite " << Gnss_Satellite(d_plan.system_name, d_synchro->PRN) << " !\n";
            LOG(INFO) << "Loss of lock in channel " << d_channel << " (carrier_lock_fail_counter:" << d_carrier_fail
                      << " code_lock_fail_counter : " << d_code_fail << ")";
            this->message_port_pub(pmt::mp("events"), pmt::from_long(3));
            d_carrier_fail = 0;
            d_code_fail = 0;
            return false;
        }
    return true;
}


// One PLL (optionally FLL-assisted) and one DLL update from the accumulated taps (:1260-1347).
void dll_pll_veml_tracking_b200::close_loops()
{
    d_carr_phase_err_hz = (d_costas ? pll_cloop_two_quadrant_atan(d_sum.p) : pll_four_quadrant_atan(d_sum.p)) / TWO_PI;
    const bool fll_pull_in = d_pull_in && d_conf.enable_fll_pull_in;
    const auto T = static_cast<float>(d_loop_time_s);
    if (fll_pull_in || d_conf.enable_fll_steady_state)
        {
            d_carr_freq_err_hz = fll_diff_atan(d_prev_prompt, d_sum.p, 0, d_loop_time_s) / TWO_PI;
            d_prev_prompt = d_sum.p;
            // during pull-in the loop is a pure FLL; afterwards the FLL assists the PLL
            d_carr_filt_hz = d_pll_filter.get_carrier_error(static_cast<float>(d_carr_freq_err_hz), fll_pull_in ? 0.0F : static_cast<float>(d_carr_phase_err_hz), T);
        }
    else
        {
            d_carr_filt_hz = d_pll_filter.get_carrier_error(0, static_cast<float>(d_carr_phase_err_hz), T);
        }
    d_doppler_hz = d_carr_filt_hz;

    d_code_err_chips = d_plan.veml ? dll_nc_vemlp_normalized(d_sum.ve, d_sum.e, d_sum.l, d_sum.vl)
                                   : dll_nc_e_minus_l_normalized(d_sum.e, d_sum.l, d_conf.spc, d_conf.slope, d_conf.y_intercept);
    d_code_filt_chips = d_dll_filter.apply(static_cast<float>(d_code_err_chips));
    d_code_freq_chips = d_plan.chip_rate_cps - d_code_filt_chips;
    if (d_conf.carrier_aiding) d_code_freq_chips += d_doppler_hz * d_plan.chip_rate_cps / d_plan.carrier_hz;

    if (d_conf.enable_doppler_correction && !d_pull_in && !d_doppler_corrected)
        {
            // a persistent DLL command means the carrier Doppler is off: re-seed the carrier filter once (:1320-1345)
            constexpr size_t kHistory = 1000;
            push_bounded(d_dll_history, kHistory, static_cast<float>(d_code_filt_chips));
            if (d_dll_history.size() == kHistory)
                {
                    const float mean = static_cast<float>(std::accumulate(d_dll_history.begin(), d_dll_history.end(), 0.0)) / static_cast<float>(kHistory);
                    if (std::fabs(mean) > 1.0)
                        {
                            const float doppler_err = static_cast<float>(d_plan.carrier_hz) * mean / static_cast<float>(d_plan.chip_rate_cps);
                            LOG(INFO) << "Detected and corrected carrier doppler error: " << doppler_err << " [Hz]";
                            d_pll_filter.initialize(static_cast<float>(d_doppler_hz) - doppler_err);
                            d_doppler_corrected = true;
                        }
                    d_dll_history.clear();
                }
        }
}


// NCO commands for the next epoch: its length in samples, phase steps, remnant phases (:1409-1483).
void dll_pll_veml_tracking_b200::advance_ncos()
{
    const double T_prn_samples = (static_cast<double>(d_plan.code_length_chips) / d_code_freq_chips) * d_conf.fs_in;
    const double K = T_prn_samples + d_rem_code_samples;
    d_epoch_samples = static_cast<int32_t>(std::floor(K));
    const auto len = static_cast<double>(d_epoch_samples);

    auto rate_from_history = [this](std::deque<std::pair<double, double>> &hist, double step, double len_, double &rate, bool need_one_sample) {
        const size_t half = d_conf.smoother_length;
        const size_t cap = half > 0 ? 2 * half : 1;
        push_bounded(hist, cap, std::pair<double, double>(step, len_));
        if (hist.size() != cap) return;
        double older = 0.0, newer = 0.0, samples = 0.0;
        for (size_t k = 0; k < half; k++)
            {
                older += hist[k].first;
                newer += hist[cap - k - 1].first;
                samples += hist[cap - k - 1].second;
            }
        older /= static_cast<double>(half);
        newer /= static_cast<double>(half);
        if (need_one_sample)
            {
                if (samples >= 1.0) rate = (newer - older) / samples;
            }
        else
            {
                rate = (samples != 0) ? (newer - older) / samples : 0.0;
            }
    };

    d_carr_step_rad = TWO_PI * (d_doppler_hz + d_cfo_hz) / d_conf.fs_in;
    if (d_conf.high_dyn) rate_from_history(d_carr_step_history, d_carr_step_rad, len, d_carr_rate_step_rad, false);
    const double dphi = d_carr_step_rad * len + 0.5 * d_carr_rate_step_rad * len * len;
    d_rem_carr_rad += static_cast<float>(dphi);  // float accumulator, as the reference keeps it
    d_rem_carr_rad = fmod(d_rem_carr_rad, TWO_PI);
    d_acc_phase_rad -= dphi;

    d_code_step_chips = d_code_freq_chips / d_conf.fs_in;
    if (d_conf.high_dyn) rate_from_history(d_code_step_history, d_code_step_chips, len, d_code_rate_step_chips, true);
    d_rem_code_samples = K - len;
    d_rem_code_chips = d_code_freq_chips * d_rem_code_samples / d_conf.fs_in;
}


// Hard-decision correlation of the last d_sync_length prompts with the secondary code / preamble (:1093-1136).
bool dll_pll_veml_tracking_b200::secondary_code_found()
{
    int32_t corr = 0;
    for (uint32_t i = 0; i < d_sync_length; i++)
        {
            const bool negative = d_sync_prompts[i].real() < 0.0;
            const bool zero = d_plan.secondary_code[i] == '0';
            corr += (negative == zero) ? 1 : -1;
        }
    if (std::abs(corr) != static_cast<int32_t>(d_sync_length)) return false;
    d_phase_180 = corr < 0;
    return true;
}


// After pull-in: has the symbol (secondary code or data bit) boundary been found?  (:2035-2113)
bool dll_pll_veml_tracking_b200::wide_tracking_sync_done()
{
    if (d_pull_in) return false;
    auto try_code_search = [this](const char *what) {
        push_bounded(d_sync_prompts, d_sync_length, d_tap_out[d_i_p]);
        if (d_sync_prompts.size() != d_sync_length || d_sync_length == 0 || !secondary_code_found()) return false;
        LOG(INFO) << d_plan.system_name << " " << d_plan.pretty_name << " " << what << " locked in channel " << d_channel;
        std::cout << d_plan.system_name << " " << d_plan.pretty_name << " " << what << " locked in channel " << d_channel
                  << " for satellite " << Gnss_Satellite(d_plan.system_name, d_synchro->PRN) << '\n';
        return true;
    };
    if (d_plan.has_secondary) return try_code_search("secondary code");
    if (d_plan.symbols_per_bit <= 1) return true;

    if (d_use_hist_sync)
        {
            if (d_bit_sync.update(d_sum.p, true))
                {
                    d_waiting_bit_edge = true;
                    int wait = d_bit_sync.epochs_until_next_edge() - 1;
                    if (wait < 0) wait += d_bit_sync.bins();
                    d_bit_edge_epoch = d_bit_sync.get_epoch_count() - 1 + wait;
                }
            if (d_waiting_bit_edge && d_bit_sync.get_epoch_count() - 1 == d_bit_edge_epoch)
                {
                    d_waiting_bit_edge = false;
                    d_use_hist_sync = false;  // one lock is enough; later events would be false alarms
                    LOG(INFO) << d_plan.system_name << " " << d_plan.pretty_name << " histogram bit synchronization locked in channel " << d_channel;
                    std::cout << d_plan.system_name << " " << d_plan.pretty_name << " histogram bit synchronization locked in channel " << d_channel
                              << " for satellite " << Gnss_Satellite(d_plan.system_name, d_synchro->PRN) << '\n';
                    return true;
                }
        }
    return try_code_search("tracking bit synchronization");
}


// Leaving WIDE_TRACKING: clear the symbol accumulators; with extended integration switch loops and taps to narrow.
void dll_pll_veml_tracking_b200::enter_extended_or_narrow()
{
    d_sum.clear();
    d_data_sum = gr_complex(0.0, 0.0);
    d_sync_prompts.clear();
    d_symbol_idx = 0;
    d_data_symbol_idx = 0;
    if (!d_extended)
        {
            d_state = NARROW_TRACKING;
            return;
        }
    d_ext_count = 0;
    d_loop_time_s = static_cast<float>(d_ext_symbols) * static_cast<float>(d_plan.code_period_s);
    d_state = EXTENDED_INTEGRATION;
    const int ms = d_ext_symbols * static_cast<int32_t>(d_plan.code_period_s * 1000.0);
    LOG(INFO) << "Enabled " << ms << " ms extended correlator in channel " << d_channel;
    std::cout << "Enabled " << ms << " ms extended correlator in channel " << d_channel << " for satellite "
              << Gnss_Satellite(d_plan.system_name, d_synchro->PRN) << '\n';
    d_dll_filter.set_update_interval(static_cast<float>(d_loop_time_s));
    d_dll_filter.set_noise_bandwidth(d_conf.dll_bw_narrow_hz);
    d_pll_filter.set_params(d_conf.fll_bw_hz, d_conf.pll_bw_narrow_hz, d_conf.pll_filter_order);
    // narrow correlator: rewritten in place, the correlator picks the new values up at the next epoch
    set_tap_spacing(d_conf.early_late_space_narrow_chips, d_conf.very_early_late_space_narrow_chips);
    d_conf.spc = d_conf.early_late_space_narrow_chips;
}


void dll_pll_veml_tracking_b200::fill_symbol_output(Gnss_Synchro &out)
{
    out = *d_synchro;
    const auto re = static_cast<double>(d_data_sum.real());
    const auto im = static_cast<double>(d_data_sum.imag());
    out.Prompt_I = d_plan.swap_iq ? im : re;
    out.Prompt_Q = d_plan.swap_iq ? re : im;
    out.Code_phase_samples = d_rem_code_samples;
    out.Carrier_phase_rads = d_acc_phase_rad;
    out.Carrier_Doppler_hz = d_doppler_hz;
    out.CN0_dB_hz = d_cn0_db_hz;
    out.correlation_length_ms = d_plan.correlation_length_ms;
    out.Flag_valid_symbol_output = true;
    d_data_sum = gr_complex(0.0, 0.0);
}


// Same 108-byte record, field for field, as dll_pll_veml_tracking::log_data (:1599-1694) - read back by
// tests/unit-tests/signal-processing-blocks/libs/tracking_dump_reader.cc and utils/python/lib/dll_pll_veml_read_tracking_dump.py.
void dll_pll_veml_tracking_b200::write_dump_record()
{
    if (!d_dump) return;
    const gr_complex prompt = d_conf.track_pilot ? d_data_prompt[0] : d_tap_out[d_i_p];
    const uint64_t prn_start = this->nitems_read(0) + static_cast<uint64_t>(d_epoch_samples);
    const float f[] = {
        d_plan.veml ? std::abs<float>(d_sum.ve) : 0.0F, std::abs<float>(d_sum.e), std::abs<float>(d_sum.p), std::abs<float>(d_sum.l),
        d_plan.veml ? std::abs<float>(d_sum.vl) : 0.0F, prompt.real(), prompt.imag()};
    const float g[] = {
        static_cast<float>(d_acc_phase_rad), static_cast<float>(d_doppler_hz),
        static_cast<float>(d_carr_rate_step_rad * d_conf.fs_in * d_conf.fs_in / TWO_PI), static_cast<float>(d_code_freq_chips),
        static_cast<float>(d_code_rate_step_chips * d_conf.fs_in * d_conf.fs_in), static_cast<float>(d_carr_phase_err_hz),
        static_cast<float>(d_carr_filt_hz), static_cast<float>(d_code_err_chips), static_cast<float>(d_code_filt_chips),
        static_cast<float>(d_cn0_db_hz), static_cast<float>(d_lock_test), static_cast<float>(d_rem_code_samples)};
    const auto aux = static_cast<double>(prn_start);
    const uint32_t prn = d_synchro->PRN;
    const uint64_t tow = d_tow_ms;
    const auto week = static_cast<uint32_t>(d_week);
    try
        {
            d_dump_file.write(reinterpret_cast<const char *>(f), sizeof(f));
            d_dump_file.write(reinterpret_cast<const char *>(&prn_start), sizeof(prn_start));
            d_dump_file.write(reinterpret_cast<const char *>(g), sizeof(g));
            d_dump_file.write(reinterpret_cast<const char *>(&aux), sizeof(aux));
            d_dump_file.write(reinterpret_cast<const char *>(&prn), sizeof(prn));
            d_dump_file.write(reinterpret_cast<const char *>(&tow), sizeof(tow));
            d_dump_file.write(reinterpret_cast<const char *>(&week), sizeof(week));
        }
    catch (const std::ofstream::failure &e)
        {
            LOG(WARNING) << "Exception writing trk dump file " << e.what();
        }
}


void dll_pll_veml_tracking_b200::estimate_tow()
{
    d_tow_ms = 0ULL;
    d_week = 0;
    if (!d_conf.tow_to_trk || d_synchro == nullptr || d_last_tow->prn != d_synchro->PRN) return;
    const double dt_s = (static_cast<double>(this->nitems_read(0)) + d_epoch_samples - static_cast<double>(d_last_tow->sample_stamp)) / d_conf.fs_in;
    d_tow_ms = (d_last_tow->tow + static_cast<uint64_t>(dt_s * 1000.0)) % static_cast<uint64_t>(604800000);
    d_week = (d_tow_ms < d_last_tow->tow) ? d_last_tow->wn + 1 : d_last_tow->wn;
}


int dll_pll_veml_tracking_b200::general_work(int noutput_items __attribute__((unused)), gr_vector_int &ninput_items,
    gr_vector_const_void_star &input_items, gr_vector_void_star &output_items)
{
    gr::thread::scoped_lock l(d_setlock);
    const auto *in = reinterpret_cast<const gr_complex *>(input_items[0]);
    auto **out = reinterpret_cast<Gnss_Synchro **>(&output_items[0]);
    Gnss_Synchro symbol = Gnss_Synchro();
    symbol.Flag_valid_symbol_output = false;
    bool lost = false;

    const uint64_t since_acq_s = (this->nitems_read(0) - d_acq_sample_stamp) / static_cast<int>(d_conf.fs_in);
    if (d_pull_in && d_conf.pull_in_time_s < since_acq_s)
        {
            d_pull_in = false;
            d_carrier_fail = 0;
            d_code_fail = 0;
        }
    estimate_tow();

    auto give_up = [&]() {
        reset_loop_state();
        d_state = STANDBY;
        lost = true;
        symbol = *d_synchro;
        d_correlator.idle();
        if (d_conf.track_pilot) d_data_correlator.idle();
    };

    switch (d_state)
        {
        case STANDBY:
            consume_each(ninput_items[0]);
            return 0;

        case PULL_IN:
            {
                // skip to the start of the next code period predicted from the acquisition result (:1948-1980)
                const int64_t since_acq = static_cast<int64_t>(this->nitems_read(0)) - static_cast<int64_t>(d_acq_sample_stamp);
                const double to_prn_start = static_cast<double>(since_acq) - d_acq_code_phase_samples;
                d_code_freq_chips = d_plan.chip_rate_cps;
                d_code_step_chips = d_code_freq_chips / d_conf.fs_in;
                d_code_rate_step_chips = 0.0;
                const double T_prn_samples = (1.0 / d_code_freq_chips) * static_cast<double>(d_plan.code_length_chips) * d_conf.fs_in;
                d_acq_code_phase_samples = T_prn_samples - std::fmod(to_prn_start, T_prn_samples);
                d_epoch_samples = round(T_prn_samples);
                const int32_t skip = round(d_acq_code_phase_samples);
                d_acc_phase_rad -= d_carr_step_rad * static_cast<double>(skip);
                d_state = WIDE_TRACKING;
                d_cn0_smoother.reset();
                d_lock_test_smoother.reset();
                LOG(INFO) << "Pull-in: " << since_acq << " samples between acquisition and tracking in channel " << d_channel;
                consume_each(skip);
                return 0;
            }

        case WIDE_TRACKING:
            {
                if (!correlate_epoch(in, ninput_items[0])) d_carrier_fail = 400000;  // GPU failure -> regular loss-of-lock path
                d_sum.clear();
                if (d_plan.veml)
                    {
                        d_sum.ve = d_tap_out[d_i_ve];
                        d_sum.vl = d_tap_out[d_i_vl];
                    }
                d_sum.e = d_tap_out[d_i_e];
                d_sum.p = d_tap_out[d_i_p];
                d_sum.l = d_tap_out[d_i_l];
                d_conf.spc = d_conf.early_late_space_chips;
                if (d_conf.bit_synchronization_time_limit_s < since_acq_s)
                    {
                        d_carrier_fail = 300000;  // symbol synchronisation took too long: give the channel back
                        LOG(INFO) << d_plan.system_name << " " << d_plan.pretty_name << " tracking synchronization time limit reached in channel " << d_channel;
                    }
                if (!lock_still_held(d_plan.code_period_s))
                    {
                        give_up();
                    }
                else
                    {
                        close_loops();
                        advance_ncos();
                        write_dump_record();
                        if (wide_tracking_sync_done()) enter_extended_or_narrow();
                    }
                break;
            }

        case EXTENDED_INTEGRATION:
            {
                if (!correlate_epoch(in, ninput_items[0])) d_carrier_fail = 400000;
                fold_epoch_into_symbol();
                advance_ncos();
                if (d_data_symbol_idx == 0)
                    {
                        write_dump_record();
                        fill_symbol_output(symbol);
                    }
                if (++d_ext_count == d_ext_symbols - 1)
                    {
                        d_ext_count = 0;
                        d_state = NARROW_TRACKING;
                    }
                break;
            }

        case NARROW_TRACKING:
            {
                if (!correlate_epoch(in, ninput_items[0])) d_carrier_fail = 400000;
                fold_epoch_into_symbol();
                if (!lock_still_held(d_plan.code_period_s * static_cast<double>(d_ext_symbols)))
                    {
                        give_up();
                    }
                else
                    {
                        close_loops();
                        advance_ncos();
                        if (!d_acc_phase_started)
                            {
                                d_acc_phase_rad = -d_rem_carr_rad;  // carrier phase observable starts coherent with the NCO
                                d_acc_phase_started = true;
                            }
                        if (d_data_symbol_idx == 0)
                            {
                                write_dump_record();
                                fill_symbol_output(symbol);
                            }
                        d_sum.clear();
                        if (d_extended) d_state = EXTENDED_INTEGRATION;
                    }
                break;
            }
        }

    symbol.TOW_at_current_symbol_ms = d_tow_ms;

    // time tags of File_Timestamp_Signal_Source that fall inside this epoch
    std::vector<gr::tag_t> tags;
    this->get_tags_in_range(tags, 0, this->nitems_read(0), this->nitems_read(0) + d_epoch_samples);
    for (const auto &t : tags)
        {
            try
                {
                    if (pmt::any_ref(t.value).type().hash_code() == typeid(const std::shared_ptr<GnssTime>).hash_code())
                        {
                            d_last_timetag = *std::any_cast<const std::shared_ptr<GnssTime>>(pmt::any_ref(t.value));
                            d_last_timetag_offset = t.offset;
                            d_timetag_pending = true;
                        }
                }
            catch (const std::exception &ex)
                {
                    LOG(WARNING) << "time tag with unexpected payload: " << ex.what();
                }
        }

    consume_each(d_epoch_samples);
    if (!symbol.Flag_valid_symbol_output && !lost) return 0;

    symbol.fs = static_cast<int64_t>(d_conf.fs_in);
    symbol.Tracking_sample_counter = this->nitems_read(0);
    symbol.Flag_valid_symbol_output = !lost;
    symbol.Flag_PLL_180_deg_phase_locked = d_phase_180;
    if (d_timetag_pending)
        {
            // re-issue the time tag on the symbol stream, advanced to this symbol's sample counter (:2281-2300)
            const uint64_t a = symbol.Tracking_sample_counter, b = d_last_timetag_offset;
            const int64_t diff = a > b ? static_cast<int64_t>(a - b) : -static_cast<int64_t>(b - a);
            double whole_ms;
            d_last_timetag.tow_ms_fraction += modf(1000.0 * static_cast<double>(diff) / d_conf.fs_in, &whole_ms);
            auto tag = std::make_shared<GnssTime>(GnssTime());
            tag->week = d_last_timetag.week;
            tag->tow_ms = d_last_timetag.tow_ms + static_cast<int>(whole_ms);
            tag->tow_ms_fraction = d_last_timetag.tow_ms_fraction;
            tag->rx_time = static_cast<double>(symbol.Tracking_sample_counter) / d_conf.fs_in;
            add_item_tag(0, this->nitems_written(0) + 1, pmt::mp("timetag"), pmt::make_any(tag));
            d_timetag_pending = false;
        }
    std::vector<gr::tag_t> sensor_tags;
    get_tags_in_range(sensor_tags, 0, nitems_read(0) - d_epoch_samples, nitems_read(0), pmt::mp("sensor_data"));
    for (const auto &t : sensor_tags) add_item_tag(0, this->nitems_written(0) + 1, t.key, t.value);

    *out[0] = std::move(symbol);
    return 1;
}
