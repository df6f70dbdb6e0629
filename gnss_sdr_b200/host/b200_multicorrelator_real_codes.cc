/*!
 * \file b200_multicorrelator_real_codes.cc
 * \brief see b200_multicorrelator_real_codes.h.  Mirrors
 * src/algorithms/tracking/libs/cpu_multicorrelator_real_codes.cc:27-160 call for call.
 */
#include "b200_multicorrelator_real_codes.h"
#include "b200gnss.h"
#include <cstdlib>
#include <mutex>

namespace b200
{
b200_engine* shared_engine()
{
    static std::mutex mu;
    static b200_engine* eng = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (eng == nullptr)
        {
            if (b200_engine_create(&eng, B200_Multicorrelator_Real_Codes::device(), nullptr) != B200_OK)
                {
                    eng = nullptr;
                }
        }
    return eng;
}
}  // namespace b200


int B200_Multicorrelator_Real_Codes::device()
{
    const char* env = std::getenv("B200_DEVICE");
    return env != nullptr ? std::atoi(env) : 0;
}


B200_Multicorrelator_Real_Codes::~B200_Multicorrelator_Real_Codes()
{
    if (d_trk != nullptr)
        {
            B200_Multicorrelator_Real_Codes::free();
        }
}


void B200_Multicorrelator_Real_Codes::set_high_dynamics_resampler(bool use_high_dynamics_resampler)
{
    d_use_high_dynamics_resampler = use_high_dynamics_resampler;
    if (d_trk != nullptr)
        {
            b200_trk_set_high_dynamics_resampler(d_trk, use_high_dynamics_resampler ? 1 : 0);
        }
}


bool B200_Multicorrelator_Real_Codes::init(int max_signal_length_samples, int n_correlators)
{
    b200_engine* eng = b200::shared_engine();
    if (eng == nullptr)
        {
            return false;
        }
    if (d_trk != nullptr)
        {
            B200_Multicorrelator_Real_Codes::free();
        }
    if (b200_trk_create(eng, &d_trk, max_signal_length_samples, n_correlators) != B200_OK)
        {
            d_trk = nullptr;
            return false;
        }
    d_n_correlators = n_correlators;
    b200_trk_set_high_dynamics_resampler(d_trk, d_use_high_dynamics_resampler ? 1 : 0);
    return true;
}


bool B200_Multicorrelator_Real_Codes::set_local_code_and_taps(int code_length_chips, const float* local_code_in, float* shifts_chips)
{
    if (d_trk == nullptr)
        {
            return false;
        }
    return b200_trk_set_local_code_and_taps(d_trk, code_length_chips, local_code_in, shifts_chips) == B200_OK;
}


bool B200_Multicorrelator_Real_Codes::set_input_output_vectors(std::complex<float>* corr_out, const std::complex<float>* sig_in)
{
    // Save CPU pointers (cpu_multicorrelator_real_codes.cc:66-72)
    d_sig_in = sig_in;
    d_corr_out = corr_out;
    return true;
}


void B200_Multicorrelator_Real_Codes::update_local_code(int, float, float, float)
{
    // the device kernel resamples the code in registers; nothing to precompute
}


bool B200_Multicorrelator_Real_Codes::Carrier_wipeoff_multicorrelator_resampler(
    float rem_carrier_phase_in_rad,
    float phase_step_rad,
    float phase_rate_step_rad,
    float rem_code_phase_chips,
    float code_phase_step_chips,
    float code_phase_rate_step_chips,
    int signal_length_samples)
{
    if (d_trk == nullptr || d_sig_in == nullptr || d_corr_out == nullptr)
        {
            return false;
        }
    return b200_trk_correlate(d_trk, reinterpret_cast<const b200_cf32*>(d_sig_in),
               rem_carrier_phase_in_rad, phase_step_rad, phase_rate_step_rad,
               rem_code_phase_chips, code_phase_step_chips, code_phase_rate_step_chips,
               signal_length_samples, reinterpret_cast<b200_cf32*>(d_corr_out)) == B200_OK;
}


bool B200_Multicorrelator_Real_Codes::Carrier_wipeoff_multicorrelator_resampler(
    float rem_carrier_phase_in_rad,
    float phase_step_rad,
    float rem_code_phase_chips,
    float code_phase_step_chips,
    float code_phase_rate_step_chips,
    int signal_length_samples)
{
    // The 6-argument CPU overload (cpu_multicorrelator_real_codes.cc:130-144) always takes the
    // plain (non high-dynamics) rotator and ignores phase_rate_step; update_local_code still
    // honours d_use_high_dynamics_resampler.  The device path keeps that asymmetry out: the
    // 6-argument form is the plain path with a zero phase rate.
    return Carrier_wipeoff_multicorrelator_resampler(rem_carrier_phase_in_rad, phase_step_rad, 0.0F,
        rem_code_phase_chips, code_phase_step_chips, code_phase_rate_step_chips, signal_length_samples);
}


bool B200_Multicorrelator_Real_Codes::free()
{
    if (d_trk != nullptr)
        {
            b200_trk_destroy(d_trk);
            d_trk = nullptr;
        }
    return true;
}


const char* B200_Multicorrelator_Real_Codes::last_error() const
{
    return b200_last_error();
}
