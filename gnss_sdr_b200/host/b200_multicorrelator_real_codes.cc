/*!
 * \file b200_multicorrelator_real_codes.cc
 * \brief see b200_multicorrelator_real_codes.h.  Mirrors
 * src/algorithms/tracking/libs/cpu_multicorrelator_real_codes.cc:27-160 call for call.
 */
#include "b200_multicorrelator_real_codes.h"
#include "b200gnss.h"
#include "b200_trk_coalescer.h"
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
    d_max_len = max_signal_length_samples;
    b200_trk_set_high_dynamics_resampler(d_trk, d_use_high_dynamics_resampler ? 1 : 0);
    return true;
}


bool B200_Multicorrelator_Real_Codes::set_local_code_and_taps(int code_length_chips, const float* local_code_in, float* shifts_chips)
{
    if (d_trk == nullptr)
        {
            return false;
        }
    // keep the caller's pointer: every correlation reads the array's current values (cpu_multicorrelator_real_codes.cc:53-63)
    d_shifts_chips = shifts_chips;
    d_code_ptr = local_code_in;
    d_code_length = code_length_chips;
    d_chan_code_valid = false;  // the coalesced channel (if any) takes the new table at its next correlation
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
    if (d_trk == nullptr || d_sig_in == nullptr || d_corr_out == nullptr || d_shifts_chips == nullptr)
        {
            return false;
        }
    if (d_coalesced)
        {
            return post(rem_carrier_phase_in_rad, phase_step_rad, phase_rate_step_rad, rem_code_phase_chips, code_phase_step_chips,
                       code_phase_rate_step_chips, signal_length_samples) &&
                   wait();
        }
    // the shifts as they are NOW (the block may have rewritten the array in place since set_local_code_and_taps)
    b200_trk_set_taps(d_trk, d_shifts_chips);
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
    // DELIBERATE DEVIATION (documented in the header): the 6-argument CPU overload
    // (cpu_multicorrelator_real_codes.cc:130-144) always takes the plain (non high-dynamics) rotator while its
    // update_local_code still honours d_use_high_dynamics_resampler, i.e. HD-resampled codes under a plain
    // rotator.  No block of the reference calls this overload (dll_pll_veml_tracking uses the 7-argument form,
    // :1236-1255).  Here the 6-argument form is the 7-argument one with a zero phase rate, so with the HD flag set
    // BOTH halves are the HD variants.
    return Carrier_wipeoff_multicorrelator_resampler(rem_carrier_phase_in_rad, phase_step_rad, 0.0F,
        rem_code_phase_chips, code_phase_step_chips, code_phase_rate_step_chips, signal_length_samples);
}


bool B200_Multicorrelator_Real_Codes::free()
{
    if (d_chan >= 0)
        {
            b200::Trk_Coalescer* co = b200::Trk_Coalescer::instance();
            if (co != nullptr) co->close_channel(d_chan);
            d_chan = -1;
            d_chan_band = -1;
        }
    if (d_trk != nullptr)
        {
            b200_trk_destroy(d_trk);
            d_trk = nullptr;
        }
    return true;
}


void B200_Multicorrelator_Real_Codes::set_stream_position(int band, uint64_t abs_index, int n_available)
{
    d_coalesced = true;
    d_band = band;
    d_abs_index = abs_index;
    d_n_available = n_available;
}


bool B200_Multicorrelator_Real_Codes::open_coalesced_channel()
{
    b200::Trk_Coalescer* co = b200::Trk_Coalescer::instance();
    if (co == nullptr) return false;
    if (d_chan >= 0 && d_chan_band != d_band)
        {
            co->close_channel(d_chan);
            d_chan = -1;
        }
    if (d_chan < 0)
        {
            d_chan = co->open_channel(d_band, d_n_correlators);
            if (d_chan < 0) return false;
            d_chan_band = d_band;
            d_chan_code_valid = false;
        }
    if (!d_chan_code_valid)
        {
            if (d_code_ptr == nullptr || d_shifts_chips == nullptr) return false;
            if (!co->set_code(d_chan, d_code_length, d_code_ptr, d_shifts_chips, d_use_high_dynamics_resampler)) return false;
            for (int k = 0; k < d_n_correlators; k++) d_sent_shifts[k] = d_shifts_chips[k];
            d_chan_code_valid = true;
        }
    return true;
}


bool B200_Multicorrelator_Real_Codes::refresh_taps()
{
    bool changed = false;
    for (int k = 0; k < d_n_correlators; k++) changed = changed || (d_sent_shifts[k] != d_shifts_chips[k]);
    if (!changed) return true;
    b200::Trk_Coalescer* co = b200::Trk_Coalescer::instance();
    if (co == nullptr || !co->set_taps(d_chan, d_shifts_chips)) return false;
    for (int k = 0; k < d_n_correlators; k++) d_sent_shifts[k] = d_shifts_chips[k];
    return true;
}


bool B200_Multicorrelator_Real_Codes::post(float rem_carrier_phase_in_rad, float phase_step_rad, float phase_rate_step_rad,
    float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips, int signal_length_samples)
{
    if (!d_coalesced || d_sig_in == nullptr || d_corr_out == nullptr || signal_length_samples > d_max_len) return false;
    if (!open_coalesced_channel() || !refresh_taps()) return false;
    b200::Trk_Coalescer* co = b200::Trk_Coalescer::instance();
    int n_offer = d_n_available > signal_length_samples ? d_n_available : signal_length_samples;
    bool behind = false;
    d_done_synchronously = false;
    if (!co->push(d_chan, d_abs_index, d_sig_in, static_cast<uint64_t>(n_offer), &behind))
        {
            if (!behind) return false;
            // this block fell more than a ring behind the others on its band: its samples are still in its own input buffer,
            // so this epoch goes the synchronous way (same kernel, same arithmetic) and wait() hands the result over
            b200_trk_set_taps(d_trk, d_shifts_chips);
            d_done_synchronously = b200_trk_correlate(d_trk, reinterpret_cast<const b200_cf32*>(d_sig_in), rem_carrier_phase_in_rad, phase_step_rad,
                                       phase_rate_step_rad, rem_code_phase_chips, code_phase_step_chips, code_phase_rate_step_chips,
                                       signal_length_samples, reinterpret_cast<b200_cf32*>(d_corr_out)) == B200_OK;
            return d_done_synchronously;
        }
    return co->post(d_chan, d_abs_index, signal_length_samples, rem_carrier_phase_in_rad, phase_step_rad, phase_rate_step_rad,
        rem_code_phase_chips, code_phase_step_chips, code_phase_rate_step_chips);
}


bool B200_Multicorrelator_Real_Codes::wait()
{
    if (d_done_synchronously)
        {
            d_done_synchronously = false;
            return true;
        }
    b200::Trk_Coalescer* co = b200::Trk_Coalescer::instance();
    return co != nullptr && d_chan >= 0 && co->wait(d_chan, d_corr_out);
}


void B200_Multicorrelator_Real_Codes::idle()
{
    b200::Trk_Coalescer* co = b200::Trk_Coalescer::instance();
    if (co != nullptr && d_chan >= 0) co->idle(d_chan);
}


const char* B200_Multicorrelator_Real_Codes::last_error() const
{
    return b200_last_error();
}
