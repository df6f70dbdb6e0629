#include "b200_pcps_acquisition_fine_doppler_core.h"

#include "b200_multicorrelator_real_codes.h"  // b200::shared_engine
#include "b200gnss.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace b200
{
namespace
{
constexpr double GPS_L1_CA_CHIP_PERIOD_S = 1.0 / 1.023e6;  // GPS_L1_CA.h
}

Pcps_Acquisition_Fine_Doppler_Core::Pcps_Acquisition_Fine_Doppler_Core(const Fine_Doppler_Conf& conf)
    : d_fft_size(static_cast<int32_t>(conf.samples_per_ms)),
      d_num_doppler_points(static_cast<int32_t>(std::floor(std::abs(2 * conf.doppler_max) / conf.doppler_step))),
      d_acq_params(conf)
{
    b200_engine* eng = shared_engine();
    if (eng == nullptr || d_fft_size < 2 || d_num_doppler_points < 1) return;
    b200_acq_conf c{};
    c.fft_size = static_cast<uint32_t>(d_fft_size);
    c.effective_fft_size = c.fft_size;
    c.consumed_samples = c.fft_size;
    c.num_doppler_bins = static_cast<uint32_t>(d_num_doppler_points);
    // update_carrier_wipeoff (:163-179): doppler_hz = doppler_step * index - doppler_step
    c.doppler_max = conf.doppler_step;
    c.doppler_step = conf.doppler_step;
    c.fs_in = conf.fs_in;
    c.samples_per_chip = static_cast<uint32_t>(std::ceil(static_cast<float>(GPS_L1_CA_CHIP_PERIOD_S) * static_cast<float>(conf.fs_in)));  // :214
    c.code_layout = 0;
    c.bit_transition_flag = 0;
    c.use_cfar = 0;  // compute_CAF is a first-vs-second-peak statistic
    c.max_dwells = std::max(2U, conf.max_dwells);
    c.n_code_slots = 1;
    c.keep_grid = 1;
    if (b200_acq_create(eng, &c, &d_acq) != B200_OK) d_acq = nullptr;
    if (b200_acq_fine_create(eng, c.fft_size, &d_fine) != B200_OK) d_fine = nullptr;
    d_10_ms_buffer.assign(static_cast<size_t>(50) * d_fft_size, std::complex<float>(0.0F, 0.0F));
}

Pcps_Acquisition_Fine_Doppler_Core::~Pcps_Acquisition_Fine_Doppler_Core()
{
    if (d_fine) b200_acq_fine_destroy(d_fine);
    if (d_acq) b200_acq_destroy(d_acq);
}

void Pcps_Acquisition_Fine_Doppler_Core::set_local_code(std::complex<float>* code)
{
    d_code.assign(code, code + d_fft_size);
    if (d_acq) b200_acq_set_local_code(d_acq, 0, reinterpret_cast<const b200_cf32*>(code));
}

int Pcps_Acquisition_Fine_Doppler_Core::estimate_Doppler()
{
    const int zero_padding_factor = 8;
    const int prn_replicas = 10;
    const int signal_samples = prn_replicas * d_fft_size;
    const int fft_size_extended = signal_samples * zero_padding_factor;
    // 1. local code aligned with the acquisition code phase estimation (:330-340); std::rotate's `last` is
    //    data + d_fft_size - 1 upstream, so the final element keeps its place
    std::vector<std::complex<float>> code_replica(d_code);
    const int shift_index = static_cast<int>(d_gnss_synchro->Acq_delay_samples);
    if (shift_index != 0) std::rotate(code_replica.data(), code_replica.data() + (d_fft_size - shift_index), code_replica.data() + d_fft_size - 1);
    // 2.-4. code wipe-off, zero-padded transform, magnitude, first maximum: on the device
    uint32_t tmp_index_freq = 0;
    if (b200_acq_fine_estimate(d_fine, reinterpret_cast<const b200_cf32*>(d_10_ms_buffer.data()),
            reinterpret_cast<const b200_cf32*>(code_replica.data()), &tmp_index_freq, nullptr) != B200_OK)
        return -1;
    d_tmp_index_freq = tmp_index_freq;
    // fftFreqBins (:360-373), evaluated for the one bin that is read
    float bin_hz;
    const int k = static_cast<int>(tmp_index_freq);
    if (k < fft_size_extended / 2)
        bin_hz = ((static_cast<float>(d_acq_params.fs_in) / 2.0) * static_cast<float>(k)) / (static_cast<float>(fft_size_extended) / 2.0);
    else
        bin_hz = ((-static_cast<float>(d_acq_params.fs_in) / 2.0) * static_cast<float>(fft_size_extended - k)) / (static_cast<float>(fft_size_extended) / 2.0);
    // 5. update the Doppler estimate (:376-379)
    if (std::abs(bin_hz - d_gnss_synchro->Acq_doppler_hz) < 1000) d_gnss_synchro->Acq_doppler_hz = static_cast<double>(bin_hz);
    return d_fft_size;
}

int Pcps_Acquisition_Fine_Doppler_Core::work(const std::complex<float>* in, int noutput_items, int* consumed)
{
    *consumed = 0;
    if (!ok() || d_gnss_synchro == nullptr) return 0;
    if (!d_active)
        {
            d_sample_counter += static_cast<uint64_t>(d_fft_size);
            *consumed = noutput_items;
            return 0;
        }
    switch (d_state)
        {
        case 0:  // S0. StandBy (:433-444)
            d_gnss_synchro->Acq_delay_samples = 0.0;
            d_gnss_synchro->Acq_doppler_hz = 0.0;
            d_gnss_synchro->Acq_samplestamp_samples = 0ULL;
            d_gnss_synchro->Acq_doppler_step = 0U;
            d_well_count = 0;
            d_test_statistics = 0.0;
            d_n_samples_in_buffer = 0;
            d_state = 1;
            break;
        case 1:  // S1. ComputeGrid (:445-456): the grid lives on the device; dwell d_well_count + 1 accumulates into it
            {
                const uint32_t slot = 0;
                b200_acq_result r{};
                if (b200_acq_search(d_acq, reinterpret_cast<const b200_cf32*>(in), &slot, 1, static_cast<uint32_t>(d_well_count + 1), &r) != B200_OK)
                    {
                        // a device failure is a negative acquisition (as in Pcps_Acquisition_Core): consume the block, or a
                        // scheduler would offer the same input for ever
                        d_sample_counter += static_cast<uint64_t>(d_fft_size);
                        *consumed = d_fft_size;
                        d_n_samples_in_buffer = 0;
                        d_state = 5;
                        break;
                    }
                std::copy(in, in + d_fft_size, &d_10_ms_buffer[d_n_samples_in_buffer]);
                d_n_samples_in_buffer += d_fft_size;
                d_well_count++;
                if (d_well_count >= static_cast<int32_t>(d_acq_params.max_dwells))
                    {
                        // compute_CAF (:182-251) of state 2: the last search already reduced the accumulated grid
                        d_test_statistics = r.test_statistics;
                        d_gnss_synchro->Acq_delay_samples = static_cast<double>(r.index_time);
                        d_gnss_synchro->Acq_doppler_hz = static_cast<double>(static_cast<int>(r.index_doppler) * d_acq_params.doppler_step - d_acq_params.doppler_max);
                        d_gnss_synchro->Acq_doppler_step = static_cast<uint32_t>(d_acq_params.doppler_step);
                        d_state = 2;
                    }
                d_sample_counter += static_cast<uint64_t>(d_fft_size);
                *consumed = d_fft_size;
                break;
            }
        case 2:  // decide (:457-468)
            d_gnss_synchro->Acq_samplestamp_samples = d_sample_counter;
            if (d_test_statistics > d_acq_params.threshold)
                {
                    d_state = 3;
                }
            else
                {
                    d_state = 5;
                    d_n_samples_in_buffer = 0;
                }
            break;
        case 3:  // fine Doppler estimation (:469-492)
            {
                const int samples_remaining = 10 * static_cast<int32_t>(d_acq_params.samples_per_ms) - d_n_samples_in_buffer;
                if (samples_remaining > noutput_items)
                    {
                        std::copy(in, in + noutput_items, &d_10_ms_buffer[d_n_samples_in_buffer]);
                        d_n_samples_in_buffer += noutput_items;
                        d_sample_counter += static_cast<uint64_t>(noutput_items);
                        *consumed = noutput_items;
                    }
                else
                    {
                        if (samples_remaining > 0)
                            {
                                std::copy(in, in + samples_remaining, &d_10_ms_buffer[d_n_samples_in_buffer]);
                                d_sample_counter += static_cast<uint64_t>(samples_remaining);
                                *consumed = samples_remaining;
                            }
                        // a failed estimate must not be reported as a positive acquisition
                        d_state = (estimate_Doppler() < 0) ? 5 : 4;
                        d_n_samples_in_buffer = 0;
                    }
                break;
            }
        case 4:  // Positive_Acq (:493-525)
            d_active = false;
            d_state = 0;
            return 1;
        case 5:  // Negative_Acq (:526-546)
            d_active = false;
            d_state = 0;
            return 2;
        default:
            d_state = 0;
            break;
        }
    return 0;
}
}  // namespace b200
