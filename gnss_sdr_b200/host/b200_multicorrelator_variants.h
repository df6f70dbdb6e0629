/*!
 * \file b200_multicorrelator_variants.h
 * \brief B200 drop-ins for the reference's two other correlator classes: Cpu_Multicorrelator (complex local code,
 * src/algorithms/tracking/libs/cpu_multicorrelator.h:37-58) and Cpu_Multicorrelator_16sc (16-bit samples and code,
 * cpu_multicorrelator_16sc.h:38-59).  Same method names, argument meaning and pointer semantics (the caller's code and
 * shifts pointers are kept; the values current at each correlation are used); failures return false instead of exit().
 * In the reference only the legacy TCP-connector tracking blocks use the first and no block uses the second.
 */
#ifndef B200_MULTICORRELATOR_VARIANTS_H
#define B200_MULTICORRELATOR_VARIANTS_H

#include <complex>
#include <cstdint>

struct b200_trk;

class B200_Multicorrelator
{
public:
    B200_Multicorrelator() = default;
    ~B200_Multicorrelator();
    bool init(int max_signal_length_samples, int n_correlators);
    bool set_local_code_and_taps(int code_length_chips, const std::complex<float>* local_code_in, float* shifts_chips);
    bool set_input_output_vectors(std::complex<float>* corr_out, const std::complex<float>* sig_in);
    void update_local_code(int, float, float) {}
    bool Carrier_wipeoff_multicorrelator_resampler(float rem_carrier_phase_in_rad, float phase_step_rad, float rem_code_phase_chips,
        float code_phase_step_chips, int signal_length_samples);
    bool free();

private:
    b200_trk* d_trk{nullptr};
    const std::complex<float>* d_sig_in{nullptr};
    std::complex<float>* d_corr_out{nullptr};
    const std::complex<float>* d_code{nullptr};
    float* d_shifts_chips{nullptr};
    int d_code_length{0};
};

class B200_Multicorrelator_16sc
{
public:
    typedef std::complex<int16_t> sc16;  // lv_16sc_t
    B200_Multicorrelator_16sc() = default;
    ~B200_Multicorrelator_16sc();
    bool init(int max_signal_length_samples, int n_correlators);
    bool set_local_code_and_taps(int code_length_chips, const sc16* local_code_in, float* shifts_chips);
    bool set_input_output_vectors(sc16* corr_out, const sc16* sig_in);
    void update_local_code(int, float, float) {}
    bool Carrier_wipeoff_multicorrelator_resampler(float rem_carrier_phase_in_rad, float phase_step_rad, float rem_code_phase_chips,
        float code_phase_step_chips, int signal_length_samples);
    bool free();

private:
    b200_trk* d_trk{nullptr};
    const sc16* d_sig_in{nullptr};
    sc16* d_corr_out{nullptr};
    const sc16* d_code{nullptr};
    float* d_shifts_chips{nullptr};
    int d_code_length{0};
};

#endif
