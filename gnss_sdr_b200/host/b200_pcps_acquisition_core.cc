/*!
 * \file b200_pcps_acquisition_core.cc
 * \brief see header.  Line references are to
 * src/algorithms/acquisition/gnuradio_blocks/pcps_acquisition.cc.
 */
#include "b200_pcps_acquisition_core.h"
#include "b200_multicorrelator_real_codes.h"  // b200::shared_engine
#include "b200gnss.h"
#include <cmath>

namespace b200
{
namespace
{
// boost::math::gamma_p_inv(a, p) for INTEGER a (compute_threshold always passes a = 2*max_dwells):
// P(a, x) = 1 - exp(-x) * sum_{k<a} x^k / k!.
// Solve Q(a, x) = q (Q = 1 - P, upper regularised incomplete gamma, integer a) for small q.
// Newton on g(x) = ln Q(a,x) - ln q, which is close to linear (ln Q ~ -x + (a-1) ln x - ln (a-1)!).
double gamma_q_inv_int(uint32_t a, double q)
{
    if (q >= 1.0) return 0.0;
    const double lq = std::log(q);
    double x = -lq + static_cast<double>(a);  // start to the right of the root
    for (int it = 0; it < 100; it++)
        {
            double term = 1.0, sum = 1.0;
            for (uint32_t k = 1; k < a; k++)
                {
                    term *= x / static_cast<double>(k);
                    sum += term;
                }
            // ln Q = -x + ln(sum);  d/dx ln Q = -term / sum   (term == x^(a-1)/(a-1)!)
            const double g = -x + std::log(sum) - lq;
            const double dg = -term / sum;
            const double step = g / dg;
            x -= step;
            if (x <= 0.0) x = 1e-300;
            if (std::fabs(step) < 1e-14 * std::fabs(x)) break;
        }
    return x;
}
}  // namespace

float compute_threshold(float pfa, uint32_t effective_fft_size, uint32_t num_doppler_bins, uint32_t max_dwells)
{
    // :52-56
    const int num_bins = effective_fft_size * num_doppler_bins;
    // gamma_p_inv(a, p) with p = (1-pfa)^(1/num_bins): p is within ~1e-9 of 1, so solve for the
    // complement q = 1-p = -expm1(log1p(-pfa)/num_bins) directly instead of losing digits in 1-p.
    // (Boost evaluates in double on p = std::pow(1.0 - pfa, 1.0 / float(num_bins)); both agree to
    // ~1e-7 relative, far inside the float the result is cast to.)
    const double q = -std::expm1(std::log1p(-static_cast<double>(pfa)) / static_cast<double>(static_cast<float>(num_bins)));
    return static_cast<float>(2.0 * gamma_q_inv_int(2 * max_dwells, q));
}


Pcps_Acquisition_Core::Pcps_Acquisition_Core(const Acq_Conf_Core& conf_)
    : d_consumed_samples(static_cast<uint32_t>(conf_.sampled_ms * conf_.samples_per_ms * (conf_.bit_transition_flag ? 2.0 : 1.0))),
      d_fft_size(conf_.sampled_ms == conf_.ms_per_code ? d_consumed_samples : d_consumed_samples * 2),
      d_effective_fft_size(conf_.bit_transition_flag ? (d_fft_size / 2) : d_fft_size),
      d_num_doppler_bins(static_cast<uint32_t>(std::ceil(static_cast<double>(2 * conf_.doppler_max) / static_cast<double>(conf_.doppler_step)))),
      d_acq_parameters(conf_)
{
    // :113-114
    d_threshold = conf_.pfa > 0.0 ? compute_threshold(conf_.pfa, d_effective_fft_size, d_num_doppler_bins, conf_.bit_transition_flag ? 1 : conf_.max_dwells) : conf_.threshold;
    // :117
    d_threshold_step_two = conf_.pfa2 > 0.0 ? compute_threshold(conf_.pfa2, d_effective_fft_size, conf_.num_doppler_bins_step2, conf_.bit_transition_flag ? 1 : conf_.max_dwells) : conf_.threshold;
    b200_engine* eng = shared_engine();
    if (eng == nullptr) return;
    b200_acq_conf c{};
    c.fft_size = d_fft_size;
    c.effective_fft_size = d_effective_fft_size;
    c.consumed_samples = d_consumed_samples;
    c.num_doppler_bins = d_num_doppler_bins;
    c.doppler_max = static_cast<int32_t>(conf_.doppler_max);
    c.doppler_step = static_cast<int32_t>(conf_.doppler_step);
    c.fs_in = conf_.fs_in;
    c.samples_per_chip = conf_.samples_per_chip;
    c.code_layout = conf_.bit_transition_flag ? 1U : (conf_.sampled_ms == conf_.ms_per_code ? 0U : 2U);
    c.bit_transition_flag = conf_.bit_transition_flag ? 1 : 0;
    c.use_cfar = conf_.use_CFAR_algorithm_flag ? 1 : 0;
    c.max_dwells = conf_.max_dwells;
    c.n_code_slots = 1;
    c.keep_grid = conf_.dump ? 1 : 0;
    if (b200_acq_create(eng, &c, &d_acq) != B200_OK)
        {
            d_acq = nullptr;
        }
}


Pcps_Acquisition_Core::~Pcps_Acquisition_Core()
{
    if (d_acq != nullptr) b200_acq_destroy(d_acq);
}


void Pcps_Acquisition_Core::set_local_code(std::complex<float>* code)
{
    if (d_acq != nullptr) b200_acq_set_local_code(d_acq, 0, reinterpret_cast<const b200_cf32*>(code));
}


void Pcps_Acquisition_Core::set_doppler_center(int32_t doppler_center)
{
    // pcps_acquisition.h:188-200: only regenerate the grid when the centre changes
    if (doppler_center != d_doppler_center)
        {
            d_doppler_center = doppler_center;
            if (d_acq != nullptr) b200_acq_set_doppler_center(d_acq, doppler_center, 0);
        }
}


void Pcps_Acquisition_Core::set_active(bool active) { d_active = active; }


void Pcps_Acquisition_Core::init()
{
    // :196-215
    if (d_gnss_synchro != nullptr)
        {
            d_gnss_synchro->Acq_delay_samples = 0.0;
            d_gnss_synchro->Acq_doppler_hz = 0.0;
            d_gnss_synchro->Acq_samplestamp_samples = 0ULL;
            d_gnss_synchro->Acq_doppler_step = 0U;
        }
    d_input_power = 0.0F;
    d_num_noncoherent_integrations_counter = 0U;
    d_state = 1;
    d_step_two = false;
}


void Pcps_Acquisition_Core::update_synchro(const AcquisitionResult& result)
{
    // :580-596 (automatic resampler branch is upstream of this path and out of scope)
    if (d_gnss_synchro == nullptr) return;
    d_gnss_synchro->Acq_delay_samples = static_cast<double>(std::fmod(static_cast<float>(result.index_time), d_acq_parameters.samples_per_code));
    d_gnss_synchro->Acq_doppler_hz = static_cast<double>(result.doppler);
    d_gnss_synchro->Acq_samplestamp_samples = result.sample_count;
    d_gnss_synchro->fs = d_acq_parameters.fs_in;
}


int Pcps_Acquisition_Core::acquisition_core(const std::complex<float>* in, uint64_t sample_count, AcquisitionResult* out)
{
    return acquisition_core_any(in, false, sample_count, out);
}


int Pcps_Acquisition_Core::acquisition_core_i16(const int16_t* in_iq, uint64_t sample_count, AcquisitionResult* out)
{
    return acquisition_core_any(in_iq, true, sample_count, out);
}


int Pcps_Acquisition_Core::acquisition_core_any(const void* in, bool cshort, uint64_t sample_count, AcquisitionResult* out)
{
    if (d_acq == nullptr) return 2;
    d_num_noncoherent_integrations_counter++;  // :666
    const uint32_t slot = 0;
    b200_acq_result r{};
    // doppler_grid + compute_statistics (:680-682) on the device
    int rc;
    if (cshort)
        {
            const auto* iq = static_cast<const int16_t*>(in);
            rc = d_step_two ? b200_acq_search_step_two_i16(d_acq, iq, slot, d_num_noncoherent_integrations_counter, d_input_power, &r)
                            : b200_acq_search_i16(d_acq, iq, &slot, 1, d_num_noncoherent_integrations_counter, &r);
        }
    else
        {
            const auto* cf = static_cast<const b200_cf32*>(in);
            rc = d_step_two ? b200_acq_search_step_two(d_acq, cf, slot, d_num_noncoherent_integrations_counter, d_input_power, &r)
                            : b200_acq_search(d_acq, cf, &slot, 1, d_num_noncoherent_integrations_counter, &r);
        }
    if (rc != B200_OK)
        {
            // a GPU failure surfaces as a negative acquisition, never as exit()
            d_num_noncoherent_integrations_counter = 0;
            d_active = false;
            d_state = 0;
            d_step_two = false;
            return 2;
        }
    AcquisitionResult result;
    result.index_time = r.index_time;
    result.doppler = r.doppler;
    result.test_statistics = r.test_statistics;
    result.sample_count = sample_count;
    if (!d_step_two) d_input_power = r.input_power;  // the second step keeps the first step's value (:428-438)
    update_synchro(result);  // :686
    if (d_step_two && d_gnss_synchro != nullptr) d_gnss_synchro->Acq_doppler_step = static_cast<uint32_t>(d_acq_parameters.doppler_step2);  // :598-601

    int event = 0;
    // handle_threshold_reached (:605-636)
    auto threshold_reached = [&]() {
        d_state = 0;
        if (d_acq_parameters.make_2_steps)
            {
                if (d_step_two)
                    {
                        result.positive_acq = true;
                        d_active = false;
                        event = 1;
                    }
                else
                    {
                        d_doppler_center_step_two = static_cast<float>(result.doppler);
                        b200_acq_set_step_two(d_acq, d_doppler_center_step_two, d_acq_parameters.doppler_step2, d_acq_parameters.num_doppler_bins_step2);
                        d_num_noncoherent_integrations_counter = 0;
                    }
                d_step_two = !d_step_two;
            }
        else
            {
                result.positive_acq = true;
                d_active = false;
                event = 1;
            }
    };
    // handle_integration_done (:639-645)
    auto integration_done = [&]() {
        if (d_state != 0) event = 2;
        d_active = false;
        d_state = 0;
        d_step_two = false;
    };
    const float th = get_threshold();
    if (!d_acq_parameters.bit_transition_flag)
        {
            if (result.test_statistics > th)
                threshold_reached();
            else
                d_state = 1;
            if (d_num_noncoherent_integrations_counter == d_acq_parameters.max_dwells) integration_done();
        }
    else
        {
            if (result.test_statistics > d_threshold)
                threshold_reached();
            else
                integration_done();
        }
    // :717-725
    if ((d_num_noncoherent_integrations_counter == d_acq_parameters.max_dwells) || result.positive_acq || d_acq_parameters.bit_transition_flag)
        {
            d_num_noncoherent_integrations_counter = 0U;
        }
    if (out != nullptr) *out = result;
    return event;
}


bool Pcps_Acquisition_Core::read_grid(float* grid) const
{
    return d_acq != nullptr && b200_acq_read_grid(d_acq, 0, grid) == B200_OK;
}
}  // namespace b200
