/*!
 * \file b200_dll_pll_tracking.cc
 * \brief see b200_dll_pll_tracking.h
 */
#include "b200_dll_pll_tracking.h"
#include "GPS_L1_CA.h"
#include "GPS_L5.h"
#include "Galileo_E1.h"
#include "configuration_interface.h"
#include "display.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <iostream>

#if USE_GLOG_AND_GFLAGS
#include <glog/logging.h>
#else
#include <absl/log/log.h>
#endif

namespace
{
struct Entry
{
    const char* implementation;
    B200DllPllTracking::Signal signal;
    char system;
    std::array<char, 3> code;
    double chips_per_second;
    double chips_per_period;
    const char* label;
};
const std::array<Entry, 3> kEntries = {{
    {"GPS_L1_CA_DLL_PLL_Tracking_B200", B200DllPllTracking::Signal::GPS_L1_CA, 'G', {{'1', 'C', '\0'}}, GPS_L1_CA_CODE_RATE_CPS, GPS_L1_CA_CODE_LENGTH_CHIPS, "GPS L1 C/A"},
    {"Galileo_E1_DLL_PLL_VEML_Tracking_B200", B200DllPllTracking::Signal::GALILEO_E1, 'E', {{'1', 'B', '\0'}}, GALILEO_E1_CODE_CHIP_RATE_CPS, GALILEO_E1_B_CODE_LENGTH_CHIPS, "Galileo E1"},
    {"GPS_L5_DLL_PLL_Tracking_B200", B200DllPllTracking::Signal::GPS_L5, 'G', {{'L', '5', '\0'}}, GPS_L5I_CODE_RATE_CPS, GPS_L5I_CODE_LENGTH_CHIPS, "GPS L5"},
}};
const Entry& entry_of(B200DllPllTracking::Signal s)
{
    return *std::find_if(kEntries.begin(), kEntries.end(), [s](const Entry& e) { return e.signal == s; });
}
}  // namespace


const char* const* B200DllPllTracking::implementations()
{
    static const char* names[] = {kEntries[0].implementation, kEntries[1].implementation, kEntries[2].implementation, nullptr};
    return names;
}


bool B200DllPllTracking::lookup(const std::string& implementation, Signal* signal)
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


B200DllPllTracking::B200DllPllTracking(Signal signal, const ConfigurationInterface* configuration, std::string role, unsigned int in_streams,
    unsigned int out_streams)
    : signal_(signal), role_(std::move(role)), item_size_(sizeof(gr_complex))
{
    trk_params_.SetFromConfiguration(configuration, role_);
    if (in_streams > 1) LOG(ERROR) << "Only one input stream is supported.";
    if (out_streams > 1) LOG(ERROR) << "Only one output stream is supported.";
    police_parameters(configuration);
    if (trk_params_.item_type == "gr_complex")
        {
            const int band = configuration->property(role_ + ".b200_band", 0);
            const bool coalesce = configuration->property(role_ + ".b200_coalesce", true);
            tracking_ = dll_pll_veml_make_tracking_b200(trk_params_, band, coalesce);
            DLOG(INFO) << "tracking(" << tracking_->unique_id() << ")";
        }
    else
        {
            item_size_ = 0;  // the factory drops the channel (gnss_block_factory.cc:1048-1052)
            LOG(WARNING) << trk_params_.item_type << " unknown tracking item type.";
        }
}


std::string B200DllPllTracking::implementation()
{
    return entry_of(signal_).implementation;
}


void B200DllPllTracking::police_parameters(const ConfigurationInterface* configuration)
{
    const Entry& e = entry_of(signal_);
    trk_params_.system = e.system;
    std::copy_n(e.code.data(), 3, trk_params_.signal);
    // one code period of samples per epoch
    trk_params_.vector_length = static_cast<int>(std::round(trk_params_.fs_in / (e.chips_per_second / e.chips_per_period)));

    auto warn = [&e](const std::string& text) { std::cout << TEXT_RED << "WARNING: " << e.label << ": " << text << TEXT_RESET << std::endl; };
    int32_t& ext = trk_params_.extend_correlation_symbols;
    if (ext < 1)
        {
            ext = 1;
            warn("extend_correlation_symbols must be > 0. Coherent integration set to one code period.");
        }
    switch (signal_)
        {
        case Signal::GPS_L1_CA:
            if (ext > 20)
                {
                    ext = 20;
                    warn("extend_correlation_symbols limited to 20 (20 ms).");
                }
            trk_params_.track_pilot = configuration->property(role_ + ".track_pilot", false);
            if (trk_params_.track_pilot)
                {
                    trk_params_.track_pilot = false;
                    warn("no pilot signal. Data tracking enabled instead.");
                }
            break;
        case Signal::GALILEO_E1:
            if (!trk_params_.track_pilot && ext > 1)
                {
                    ext = 1;
                    warn("extended coherent integration is not allowed when tracking the data component. Set to 4 ms.");
                }
            break;
        case Signal::GPS_L5:
            if (!trk_params_.track_pilot && ext > GPS_L5I_NH_CODE_LENGTH)
                {
                    ext = GPS_L5I_NH_CODE_LENGTH;
                    warn("extend_correlation_symbols must be lower than 11 when tracking the data component. Set to 10 ms.");
                }
            break;
        }
    if (ext > 1 && (trk_params_.pll_bw_narrow_hz > trk_params_.pll_bw_hz || trk_params_.dll_bw_narrow_hz > trk_params_.dll_bw_hz))
        {
            warn("narrow tracking bandwidth is higher than the wide one.");
        }
}


void B200DllPllTracking::connect(gr::top_block_sptr top_block)
{
    if (top_block)
        { /* a single block: nothing to wire */
        }
}


void B200DllPllTracking::disconnect(gr::top_block_sptr top_block)
{
    if (top_block)
        { /* nothing to undo */
        }
}
