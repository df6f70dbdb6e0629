/*!
 * \file b200_pcps_acquisition.cc
 * \brief see b200_pcps_acquisition.h
 */
#include "b200_pcps_acquisition.h"
#include "GPS_L1_CA.h"
#include "GPS_L5.h"
#include "Galileo_E1.h"
#include "configuration_interface.h"
#include "galileo_e1_signal_replica.h"
#include "gnss_sdr_flags.h"
#include "gps_l5_signal_replica.h"
#include "gps_sdr_signal_replica.h"
#include <algorithm>
#include <array>
#include <cmath>

#if USE_GLOG_AND_GFLAGS
#include <glog/logging.h>
#else
#include <absl/log/log.h>
#endif

#if HAS_STD_SPAN
#include <span>
namespace own = std;
#else
#include <gsl-lite/gsl-lite.hpp>
namespace own = gsl_lite;
#endif

namespace
{
struct Entry
{
    const char* implementation;
    B200PcpsAcquisition::Signal signal;
    double chip_rate;
    double opt_freq;
    double code_length_chips;
    uint32_t ms_per_code;
};
const std::array<Entry, 3> kEntries = {{
    {"GPS_L1_CA_PCPS_Acquisition_B200", B200PcpsAcquisition::Signal::GPS_L1_CA, GPS_L1_CA_CODE_RATE_CPS, GPS_L1_CA_OPT_ACQ_FS_SPS, GPS_L1_CA_CODE_LENGTH_CHIPS, GPS_L1_CA_CODE_PERIOD_MS},
    {"Galileo_E1_PCPS_Ambiguous_Acquisition_B200", B200PcpsAcquisition::Signal::GALILEO_E1, GALILEO_E1_CODE_CHIP_RATE_CPS, GALILEO_E1_OPT_ACQ_FS_SPS, GALILEO_E1_B_CODE_LENGTH_CHIPS, GALILEO_E1_CODE_PERIOD_MS},
    {"GPS_L5i_PCPS_Acquisition_B200", B200PcpsAcquisition::Signal::GPS_L5I, GPS_L5I_CODE_RATE_CPS, GPS_L5_OPT_ACQ_FS_SPS, GPS_L5I_CODE_LENGTH_CHIPS, GPS_L5I_PERIOD_MS},
}};
const Entry& entry_of(B200PcpsAcquisition::Signal s)
{
    return *std::find_if(kEntries.begin(), kEntries.end(), [s](const Entry& e) { return e.signal == s; });
}

Acq_Conf read_conf(const ConfigurationInterface* configuration, const std::string& role, const Entry& e)
{
    Acq_Conf p;
    p.ms_per_code = e.ms_per_code;
    p.sampled_ms = e.ms_per_code;
    p.SetFromConfiguration(configuration, role, e.chip_rate, e.opt_freq);
#if USE_GLOG_AND_GFLAGS
    if (FLAGS_doppler_max != 0) p.doppler_max = FLAGS_doppler_max;
    if (FLAGS_doppler_step != 0) p.doppler_step = FLAGS_doppler_step;
#else
    if (absl::GetFlag(FLAGS_doppler_max) != 0) p.doppler_max = absl::GetFlag(FLAGS_doppler_max);
    if (absl::GetFlag(FLAGS_doppler_step) != 0) p.doppler_step = absl::GetFlag(FLAGS_doppler_step);
#endif
    return p;
}
}  // namespace


const char* const* B200PcpsAcquisition::implementations()
{
    static const char* names[] = {kEntries[0].implementation, kEntries[1].implementation, kEntries[2].implementation, nullptr};
    return names;
}


bool B200PcpsAcquisition::lookup(const std::string& implementation, Signal* signal)
{
    for (const auto& e : kEntries)
        {
            if (implementation == e.implementation)
                {
                    if (signal != nullptr) *signal = e.signal;
                    return true;
                }
        }
    return false;
}


B200PcpsAcquisition::B200PcpsAcquisition(Signal signal, const ConfigurationInterface* configuration, const std::string& role, unsigned int in_streams,
    unsigned int out_streams)
    : signal_(signal),
      role_(role),
      acq_parameters_(read_conf(configuration, role, entry_of(signal))),
      acquire_pilot_(configuration->property(role + ".acquire_pilot", false)),
      cboc_(configuration->property(role + ".cboc", false)),
      supported_item_type_(acq_parameters_.item_type == "gr_complex" || acq_parameters_.item_type == "cshort"),
      vector_length_(std::floor(acq_parameters_.sampled_ms * acq_parameters_.samples_per_ms) * (acq_parameters_.bit_transition_flag ? 2.0 : 1.0)),
      code_length_(static_cast<unsigned int>(std::floor(static_cast<double>(acq_parameters_.resampled_fs) / (entry_of(signal).chip_rate / entry_of(signal).code_length_chips)))),
      code_(vector_length_),
      acquisition_(pcps_make_acquisition_b200(acq_parameters_))
{
    DLOG(INFO) << "role " << role << ", acquisition(" << acquisition_->unique_id() << ")";
    if (!supported_item_type_) LOG(WARNING) << acq_parameters_.item_type << " is not an item type of the B200 acquisition (gr_complex, cshort)";
    if (in_streams > 1) LOG(ERROR) << "This implementation only supports one input stream";
    if (out_streams > 0) LOG(ERROR) << "This implementation does not provide an output stream";
}


std::string B200PcpsAcquisition::implementation()
{
    return entry_of(signal_).implementation;
}


void B200PcpsAcquisition::set_gnss_synchro(Gnss_Synchro* p_gnss_synchro)
{
    gnss_synchro_ = p_gnss_synchro;
    acquisition_->set_gnss_synchro(p_gnss_synchro);
}


void B200PcpsAcquisition::connect(gr::top_block_sptr top_block)
{
    if (top_block)
        { /* a single block: nothing to wire */
        }
}


void B200PcpsAcquisition::disconnect(gr::top_block_sptr top_block)
{
    if (top_block)
        { /* nothing to undo */
        }
}


// One code period sampled at the stream's rate, repeated over the coherent time, handed to the block
// (BasePcpsAcquisition::set_local_code, base_pcps_acquisition.cc:206-222).
void B200PcpsAcquisition::set_local_code()
{
    volk_gnsssdr::vector<std::complex<float>> one_period(code_length_);
    const auto fs = static_cast<int32_t>(acq_parameters_.use_automatic_resampler ? acq_parameters_.resampled_fs : acq_parameters_.fs_in);
    const own::span<std::complex<float>> dest(one_period.data(), one_period.size());
    const uint32_t prn = gnss_synchro_->PRN;
    switch (signal_)
        {
        case Signal::GPS_L1_CA:
            gps_l1_ca_code_gen_complex_sampled(dest, prn, fs, 0);
            break;
        case Signal::GALILEO_E1:
            {
                std::array<char, 3> sig = {{'1', 'C', '\0'}};  // pilot component
                if (!acquire_pilot_)
                    {
                        sig[0] = gnss_synchro_->Signal[0];
                        sig[1] = gnss_synchro_->Signal[1];
                    }
                galileo_e1_code_gen_complex_sampled(dest, sig, cboc_, prn, fs, 0, false);
                break;
            }
        case Signal::GPS_L5I:
            gps_l5i_code_gen_complex_sampled(dest, prn, fs);
            break;
        }
    const auto periods = acq_parameters_.sampled_ms / acq_parameters_.ms_per_code;
    for (unsigned int i = 0; i < periods; i++) std::copy_n(one_period.data(), code_length_, code_.data() + static_cast<size_t>(i) * code_length_);
    acquisition_->set_local_code(code_.data());
}
