/*!
 * \file b200_dll_pll_veml_loop.h
 * \brief The tracking state of one dll_pll_veml_tracking block kept and advanced on a B200.
 *
 * The per-epoch cycle of src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc (general_work cases 1-2:
 * do_correlation_step :1232-1257, cn0_and_tracking_lock_status :1167-1224, run_dll_pll :1260-1347,
 * update_tracking_vars :1409-1483, log_data :1599-1694) runs on the device (b200_trk_loop_*); this class is the part of
 * the block that configures it from a Dll_Pll_Conf (constructor :97-700), takes the acquisition result
 * (start_tracking :791-1078) and hands back what the block publishes: the per-epoch records and the lock status.
 * The samples reach the device once per band through b200_iq_push, not once per channel.
 */
#ifndef B200_DLL_PLL_VEML_LOOP_H
#define B200_DLL_PLL_VEML_LOOP_H

#include "b200gnss.h"

#include <cstdint>
#include <vector>

namespace b200
{
/*! The Dll_Pll_Conf fields the cycle reads (src/algorithms/tracking/libs/dll_pll_conf.h:32-89), same names and defaults. */
struct Dll_Pll_Conf_Core
{
    double fs_in{2000000.0};
    double carrier_lock_th{0.7};  // FLAGS_carrier_lock_th
    float fll_bw_hz{35.0F};
    float pll_bw_hz{35.0F};
    float dll_bw_hz{2.0F};
    float early_late_space_chips{0.25F};
    float very_early_late_space_chips{0.5F};
    float slope{1.0F};
    float y_intercept{1.0F};
    float cn0_smoother_alpha{0.002F};
    float carrier_lock_test_smoother_alpha{0.002F};
    uint32_t pull_in_time_s{5U};
    uint32_t bit_synchronization_time_limit_s{20U};
    uint32_t vector_length{0U};
    int32_t pll_filter_order{3};
    int32_t dll_filter_order{2};
    int32_t cn0_samples{20};
    int32_t cn0_smoother_samples{200};
    int32_t carrier_lock_test_smoother_samples{25};
    int32_t cn0_min{25};
    int32_t max_code_lock_fail{50};
    int32_t max_carrier_lock_fail{5000};
    bool enable_fll_pull_in{false};
    bool enable_fll_steady_state{false};
    bool carrier_aiding{true};
};

/*! What the constructor of the block derives from the signal type (:159-590), given explicitly. */
struct Signal_Core
{
    double code_chip_rate{1.023e6};      // d_code_chip_rate
    double signal_carrier_freq{1575.42e6};
    double code_period{0.001};
    uint32_t code_length_chips{1023};
    int32_t code_samples_per_chip{1};
    bool veml{false};                    // d_veml: Galileo E1 (five correlators)
    uint32_t prn{1};
};

class B200_Dll_Pll_Veml_Loop
{
public:
    B200_Dll_Pll_Veml_Loop() = default;
    /*! band: the IQ band (b200_iq_create / b200_iq_push on b200::shared_engine()) this channel listens to.
     *  tracking_code: code_samples_per_chip * code_length_chips values (gps_l1_ca_code_gen_float, ...). */
    bool init(const Dll_Pll_Conf_Core& conf, const Signal_Core& sig, int band, const float* tracking_code);
    /*! start_tracking(): Acq_delay_samples, Acq_doppler_hz, Acq_samplestamp_samples of the Gnss_Synchro and the absolute
     *  index of the next sample the block would read (nitems_read). */
    bool start_tracking(double acq_delay_samples, double acq_doppler_hz, uint64_t acq_samplestamp_samples, uint64_t nitems_read);
    bool status(b200_trk_loop_status* out) const;
    int loop_id() const { return d_loop; }

    /*! Advance ALL loops of the shared engine by up to max_epochs cycles (they run side by side in one persistent kernel)
     *  and return this loop's records. */
    bool run(int max_epochs, std::vector<b200_trk_dump_record>* records);

private:
    int d_channel{-1};
    int d_loop{-1};
};
}  // namespace b200
#endif
