/*!
 * \file b200_multicorrelator_real_codes.h
 * \brief Drop-in for Cpu_Multicorrelator_Real_Codes that runs the carrier wipe-off, code
 * resampling and multi-tap correlation on a B200 through libb200gnss.so.
 *
 * Same public interface, argument meaning and return values as
 * src/algorithms/tracking/libs/cpu_multicorrelator_real_codes.h:37-61 so that
 * dll_pll_veml_tracking (src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc:
 * 663-676 init, :837-866 set_local_code_and_taps, :1232-1257 do_correlation_step) only needs
 * its member type changed.
 *
 * Semantics kept from the CPU class:
 *   - set_local_code_and_taps() keeps the caller's shifts_chips POINTER (cpu_multicorrelator_real_codes.cc:53-63)
 *     and every correlation reads the array's CURRENT values: the tracking block mutates it in place after the call
 *     (start_tracking :1045-1053, the narrow-correlator switch :2132-2146) and never calls
 *     set_local_code_and_taps again.  The current shifts travel with every correlation (host-mapped control
 *     block, or a descriptor refresh when they changed in coalesced mode).
 *   - the local code table is copied to the device at set_local_code_and_taps() (the CPU class keeps that pointer
 *     too, but the block only rewrites the table right before calling set_local_code_and_taps again, :811-1030).
 * Differences, all invisible to the caller:
 *   - update_local_code() is a no-op: the resampled codes are never materialised;
 *   - failures (CUDA errors, oversize input) return false instead of calling exit();
 *   - the 6-argument overload ignores the high-dynamics flag for BOTH halves (see the .cc).
 *
 * Beyond the CPU class (used by the B200 tracking block, integration/): set_stream_position() tells the correlator
 * where the input pointer sits in the band's sample stream; the correlation then goes through the per-process
 * coalescer (b200_trk_coalescer.h) - samples are copied to the GPU once for all channels and the epochs of all
 * channels that are due share one launch.  post() / wait() split the call so that a block can have its pilot
 * and data correlators in flight together.
 */
#ifndef B200_MULTICORRELATOR_REAL_CODES_H
#define B200_MULTICORRELATOR_REAL_CODES_H

#include <complex>
#include <cstdint>

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

    // ---- extensions -------------------------------------------------------------------------------------------
    /*! Coalesced mode: sig_in[0] (set_input_output_vectors) is sample `abs_index` of band `band`, and
     *  `n_available` samples from there may be offered to the band store.  Call before every correlation
     *  (gr::block::nitems_read(0) and ninput_items[0] in general_work). */
    void set_stream_position(int band, uint64_t abs_index, int n_available);
    //! back to one synchronous launch per call
    void clear_stream_position() { d_coalesced = false; }
    //! coalesced mode only: queue the correlation / collect its taps into corr_out
    bool post(float rem_carrier_phase_in_rad, float phase_step_rad, float phase_rate_step_rad, float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips, int signal_length_samples);
    bool wait();
    //! the channel leaves the batches until it correlates again (stop_tracking, loss of lock)
    void idle();

    //! last error text from the library (empty when the last call succeeded)
    const char* last_error() const;
    //! GPU used by all correlators of this process (default 0, or env B200_DEVICE)
    static int device();

private:
    bool open_coalesced_channel();
    bool refresh_taps();

    b200_trk* d_trk{nullptr};
    const std::complex<float>* d_sig_in{nullptr};
    std::complex<float>* d_corr_out{nullptr};
    float* d_shifts_chips{nullptr};  // the caller's array, read at every correlation (reference semantics)
    int d_n_correlators{0};
    int d_max_len{0};
    bool d_use_high_dynamics_resampler{true};  // same default as the CPU class (.h:60)
    // coalesced mode
    bool d_coalesced{false};
    bool d_done_synchronously{false};  // post() fell back to the synchronous call for this epoch; wait() reports it
    int d_band{0};
    uint64_t d_abs_index{0};
    int d_n_available{0};
    int d_chan{-1};            // coalescer / engine channel id
    int d_chan_band{-1};
    bool d_chan_code_valid{false};
    int d_code_length{0};
    const float* d_code_ptr{nullptr};
    float d_sent_shifts[8] = {0};
};

namespace b200
{
//! process-wide engine (one per process and GPU), created on first use; nullptr on failure
b200_engine* shared_engine();
}  // namespace b200

#endif  // B200_MULTICORRELATOR_REAL_CODES_H
