/*!
 * \file b200_multicorrelator_variants.cc
 * \brief see header
 */
#include "b200_multicorrelator_variants.h"
#include "b200_multicorrelator_real_codes.h"  // b200::shared_engine
#include "b200gnss.h"

B200_Multicorrelator::~B200_Multicorrelator() { B200_Multicorrelator::free(); }

bool B200_Multicorrelator::init(int max_signal_length_samples, int n_correlators)
{
    b200_engine* eng = b200::shared_engine();
    if (eng == nullptr) return false;
    B200_Multicorrelator::free();
    return b200_trk_create(eng, &d_trk, max_signal_length_samples, n_correlators) == B200_OK;
}

bool B200_Multicorrelator::set_local_code_and_taps(int code_length_chips, const std::complex<float>* local_code_in, float* shifts_chips)
{
    d_code = local_code_in;
    d_shifts_chips = shifts_chips;
    d_code_length = code_length_chips;
    return d_trk != nullptr;
}

bool B200_Multicorrelator::set_input_output_vectors(std::complex<float>* corr_out, const std::complex<float>* sig_in)
{
    d_sig_in = sig_in;
    d_corr_out = corr_out;
    return true;
}

bool B200_Multicorrelator::Carrier_wipeoff_multicorrelator_resampler(float rem_carrier_phase_in_rad, float phase_step_rad, float rem_code_phase_chips,
    float code_phase_step_chips, int signal_length_samples)
{
    if (d_trk == nullptr || d_code == nullptr || d_sig_in == nullptr || d_corr_out == nullptr) return false;
    // the class keeps pointers (cpu_multicorrelator.cc:53-63): code table and shifts as they are now
    if (b200_trk_set_local_code_and_taps_cplx(d_trk, d_code_length, reinterpret_cast<const b200_cf32*>(d_code), d_shifts_chips) != B200_OK) return false;
    return b200_trk_correlate_cplx(d_trk, reinterpret_cast<const b200_cf32*>(d_sig_in), rem_carrier_phase_in_rad, phase_step_rad, rem_code_phase_chips,
               code_phase_step_chips, signal_length_samples, reinterpret_cast<b200_cf32*>(d_corr_out)) == B200_OK;
}

bool B200_Multicorrelator::free()
{
    if (d_trk != nullptr) b200_trk_destroy(d_trk);
    d_trk = nullptr;
    return true;
}


B200_Multicorrelator_16sc::~B200_Multicorrelator_16sc() { B200_Multicorrelator_16sc::free(); }

bool B200_Multicorrelator_16sc::init(int max_signal_length_samples, int n_correlators)
{
    b200_engine* eng = b200::shared_engine();
    if (eng == nullptr) return false;
    B200_Multicorrelator_16sc::free();
    return b200_trk_create(eng, &d_trk, max_signal_length_samples, n_correlators) == B200_OK;
}

bool B200_Multicorrelator_16sc::set_local_code_and_taps(int code_length_chips, const sc16* local_code_in, float* shifts_chips)
{
    d_code = local_code_in;
    d_shifts_chips = shifts_chips;
    d_code_length = code_length_chips;
    return d_trk != nullptr;
}

bool B200_Multicorrelator_16sc::set_input_output_vectors(sc16* corr_out, const sc16* sig_in)
{
    d_sig_in = sig_in;
    d_corr_out = corr_out;
    return true;
}

bool B200_Multicorrelator_16sc::Carrier_wipeoff_multicorrelator_resampler(float rem_carrier_phase_in_rad, float phase_step_rad, float rem_code_phase_chips,
    float code_phase_step_chips, int signal_length_samples)
{
    if (d_trk == nullptr || d_code == nullptr || d_sig_in == nullptr || d_corr_out == nullptr) return false;
    if (b200_trk_set_local_code_and_taps_16sc(d_trk, d_code_length, reinterpret_cast<const int16_t*>(d_code), d_shifts_chips) != B200_OK) return false;
    return b200_trk_correlate_16sc(d_trk, reinterpret_cast<const int16_t*>(d_sig_in), rem_carrier_phase_in_rad, phase_step_rad, rem_code_phase_chips,
               code_phase_step_chips, signal_length_samples, reinterpret_cast<int16_t*>(d_corr_out)) == B200_OK;
}

bool B200_Multicorrelator_16sc::free()
{
    if (d_trk != nullptr) b200_trk_destroy(d_trk);
    d_trk = nullptr;
    return true;
}
