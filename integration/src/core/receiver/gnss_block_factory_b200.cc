/*!
 * \file gnss_block_factory_b200.cc
 * \brief see gnss_block_factory_b200.h
 */
#include "gnss_block_factory_b200.h"
#include "b200_dll_pll_tracking.h"
#include "b200_pcps_acquisition.h"

std::unique_ptr<AcquisitionInterface> get_b200_acq_block(const std::string& implementation, const ConfigurationInterface* configuration,
    const std::string& role, unsigned int in_streams, unsigned int out_streams)
{
    B200PcpsAcquisition::Signal signal;
    if (!B200PcpsAcquisition::lookup(implementation, &signal)) return nullptr;
    return std::make_unique<B200PcpsAcquisition>(signal, configuration, role, in_streams, out_streams);
}


std::unique_ptr<TrackingInterface> get_b200_trk_block(const std::string& implementation, const ConfigurationInterface* configuration,
    const std::string& role, unsigned int in_streams, unsigned int out_streams)
{
    B200DllPllTracking::Signal signal;
    if (!B200DllPllTracking::lookup(implementation, &signal)) return nullptr;
    return std::make_unique<B200DllPllTracking>(signal, configuration, role, in_streams, out_streams);
}
