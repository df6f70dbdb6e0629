/*!
 * \file b200_multicorrelator_real_codes.h
 * \brief Drop-in for Cpu_Multicorrelator_Real_Codes that runs the carrier wipe-off, code
 * resampling and multi-tap correlation on a B200 through libb200gnss.so.
 *
 * Same public interface, argument meaning and return values as
 * src/algorithms/tracking/libs/cpu_multicorrelator_real_codes.h:37-61 so that
 * dll_pll_veml_tracking (src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc:
 * 663-676 init, :837-866 set_local_code_and_taps, :1232-1257 do_correlation_step) only needs
 * its member type changed.  Differences, all invisible to the caller:
 *   - local code and tap shifts are copied to the device at set_local_code_and_taps()
 *     (the CPU class keeps the caller's pointers) - call it again after changing the shifts,
 *     which is what the tracking block already does (:2135-2143);
 *   - update_local_code() is a no-op: the resampled codes are never materialised;
 *   - failures (CUDA errors, oversize input) return false instead of calling exit().
 */
#ifndef B200_MULTICORRELATOR_REAL_CODES_H
#define B200_MULTICORRELATOR_REAL_CODES_H

#include <complex>

struct b200_engine;
struct b200_trk;

class B200_Multicorrelator_Real_Codes
{
public:
    B200_Multicorrelator_Real_Codes() = default;
    ~B200_Multicorrelator_Real_Codes();
    B200_Multicorrelator_Real_Codes(const B200_Multicorrelator_Real_Codes&) = delete;
    B200_Multicorrelator_Real_Codes& operator=(const B200_Multicorrelator_Real_Codes&) = delete;

    void set_high_dynamics_resampler(bool use_high_dynamics_resampler);
    bool init(int max_signal_length_samples, int n_correlators);
    bool set_local_code_and_taps(int code_length_chips, const float* local_code_in, float* shifts_chips);
    bool set_input_output_vectors(std::complex<float>* corr_out, const std::complex<float>* sig_in);
    void update_local_code(int correlator_length_samples, float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips = 0.0);
    bool Carrier_wipeoff_multicorrelator_resampler(float rem_carrier_phase_in_rad, float phase_step_rad, float phase_rate_step_rad, float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips, int signal_length_samples);
    bool Carrier_wipeoff_multicorrelator_resampler(float rem_carrier_phase_in_rad, float phase_step_rad, float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips, int signal_length_samples);
    bool free();

    //! last error text from the library (empty when the last call succeeded)
    const char* last_error() const;
    //! GPU used by all correlators of this process (default 0, or env B200_DEVICE)
    static int device();

private:
    b200_trk* d_trk{nullptr};
    const std::complex<float>* d_sig_in{nullptr};
    std::complex<float>* d_corr_out{nullptr};
    int d_n_correlators{0};
    bool d_use_high_dynamics_resampler{true};  // same default as the CPU class (.h:60)
};

namespace b200
{
//! process-wide engine (one per process and GPU), created on first use; nullptr on failure
b200_engine* shared_engine();
}  // namespace b200

#endif  // B200_MULTICORRELATOR_REAL_CODES_H
