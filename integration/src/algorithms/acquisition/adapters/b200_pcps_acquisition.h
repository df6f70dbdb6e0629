/*!
 * \file b200_pcps_acquisition.h
 * \brief AcquisitionInterface adapters around pcps_acquisition_b200 (grid search on a B200 GPU).
 *
 * One class, three configuration-file implementations (gnss_block_factory.cc arms in
 * integration/patches/gnss_block_factory_b200.patch):
 *   GPS_L1_CA_PCPS_Acquisition_B200              parameters of GPS_L1_CA_PCPS_Acquisition
 *   Galileo_E1_PCPS_Ambiguous_Acquisition_B200   parameters of Galileo_E1_PCPS_Ambiguous_Acquisition (+ acquire_pilot, cboc)
 *   GPS_L5i_PCPS_Acquisition_B200                parameters of GPS_L5i_PCPS_Acquisition
 * Configuration handling follows BasePcpsAcquisition (src/algorithms/acquisition/adapters/base_pcps_acquisition.cc:38-104,
 * :206-222), which cannot be reused because it creates a pcps_acquisition by type (base_pcps_acquisition.h:167).
 * Item types: gr_complex and cshort go straight into the block (cshort is converted on the device); cbyte is not
 * offered (the reference widens it with two extra flowgraph blocks, base_pcps_acquisition.cc:149-164).
 */
#ifndef GNSS_SDR_B200_PCPS_ACQUISITION_H
#define GNSS_SDR_B200_PCPS_ACQUISITION_H

#include "acq_conf.h"
#include "acquisition_interface.h"
#include "channel_fsm.h"
#include "gnss_synchro.h"
#include "pcps_acquisition_b200.h"
#include <volk_gnsssdr/volk_gnsssdr_alloc.h>
#include <complex>
#include <memory>
#include <string>

class ConfigurationInterface;

class B200PcpsAcquisition : public AcquisitionInterface
{
public:
    enum class Signal
    {
        GPS_L1_CA,
        GALILEO_E1,
        GPS_L5I
    };
    B200PcpsAcquisition(Signal signal, const ConfigurationInterface* configuration, const std::string& role, unsigned int in_streams,
        unsigned int out_streams);
    ~B200PcpsAcquisition() override = default;

    std::string role() override { return role_; }
    std::string implementation() override;
    size_t item_size() override { return supported_item_type_ ? acq_parameters_.it_size : 0; }
    void connect(gr::top_block_sptr top_block) override;
    void disconnect(gr::top_block_sptr top_block) override;
    gr::basic_block_sptr get_left_block() override { return acquisition_; }
    gr::basic_block_sptr get_right_block() override { return acquisition_; }

    void set_gnss_synchro(Gnss_Synchro* p_gnss_synchro) override;
    void set_channel(unsigned int channel) override { acquisition_->set_channel(channel); }
    void set_channel_fsm(std::weak_ptr<ChannelFsm> channel_fsm) override { acquisition_->set_channel_fsm(std::move(channel_fsm)); }
    void set_doppler_center(int doppler_center) override { acquisition_->set_doppler_center(doppler_center); }
    void set_local_code() override;
    signed int mag() override { return static_cast<signed int>(acquisition_->mag()); }
    void reset() override { acquisition_->set_active(true); }
    void stop_acquisition() override { acquisition_->set_active(false); }
    void set_resampler_latency(uint32_t latency_samples) override { acquisition_->set_resampler_latency(latency_samples); }

    static const char* const* implementations();
    static bool lookup(const std::string& implementation, Signal* signal);

private:
    const Signal signal_;
    const std::string role_;
    const Acq_Conf acq_parameters_;
    const bool acquire_pilot_;
    const bool cboc_;
    const bool supported_item_type_;
    const unsigned int vector_length_;
    const unsigned int code_length_;
    Gnss_Synchro* gnss_synchro_{nullptr};
    volk_gnsssdr::vector<std::complex<float>> code_;
    pcps_acquisition_b200_sptr acquisition_;
};

#endif  // GNSS_SDR_B200_PCPS_ACQUISITION_H
