/*!
 * \file gnss_block_factory_b200.h
 * \brief The factory arms of the B200 blocks, kept out of gnss_block_factory.cc so that the patch to that file is
 * three lines per function (integration/patches/gnss_block_factory_b200.patch).
 *
 * GNSSBlockFactory maps `Acquisition_XX.implementation` / `Tracking_XX.implementation` strings to adapters in
 * get_acq_block() / get_trk_block() (src/core/receiver/gnss_block_factory.cc:449-579, :582-687); a name it does not
 * know returns nullptr.  These two functions have the same contract for the *_B200 names.
 */
#ifndef GNSS_SDR_GNSS_BLOCK_FACTORY_B200_H
#define GNSS_SDR_GNSS_BLOCK_FACTORY_B200_H

#include "acquisition_interface.h"
#include "tracking_interface.h"
#include <memory>
#include <string>

class ConfigurationInterface;

std::unique_ptr<AcquisitionInterface> get_b200_acq_block(const std::string& implementation, const ConfigurationInterface* configuration,
    const std::string& role, unsigned int in_streams, unsigned int out_streams);

std::unique_ptr<TrackingInterface> get_b200_trk_block(const std::string& implementation, const ConfigurationInterface* configuration,
    const std::string& role, unsigned int in_streams, unsigned int out_streams);

#endif  // GNSS_SDR_GNSS_BLOCK_FACTORY_B200_H
