/*!
 * \file b200_dll_pll_tracking.h
 * \brief TrackingInterface adapters around dll_pll_veml_tracking_b200 (B200 GPU correlators).
 *
 * One class, three configuration-file implementations (selected in gnss_block_factory.cc, see
 * integration/patches/gnss_block_factory_b200.patch):
 *   GPS_L1_CA_DLL_PLL_Tracking_B200          same parameters as GPS_L1_CA_DLL_PLL_Tracking
 *   Galileo_E1_DLL_PLL_VEML_Tracking_B200    same parameters as Galileo_E1_DLL_PLL_VEML_Tracking
 *   GPS_L5_DLL_PLL_Tracking_B200             same parameters as GPS_L5_DLL_PLL_Tracking
 * The per-signal parameter policing (vector length, limits on extend_correlation_symbols, pilot handling) is what
 * the reference adapters do in their configure_tracking_parameters()
 * (src/algorithms/tracking/adapters/gps_l1_ca_dll_pll_tracking.cc:50-97,
 *  galileo_e1_dll_pll_veml_tracking.cc:48-71, gps_l5_dll_pll_tracking.cc:48-71).
 * The reference's BaseDllPllTracking cannot be reused: it holds a dll_pll_veml_tracking_sptr by type
 * (base_dll_pll_tracking.h:115).
 */
#ifndef GNSS_SDR_B200_DLL_PLL_TRACKING_H
#define GNSS_SDR_B200_DLL_PLL_TRACKING_H

#include "dll_pll_conf.h"
#include "dll_pll_veml_tracking_b200.h"
#include "tracking_interface.h"
#include <cstddef>
#include <string>

class ConfigurationInterface;

class B200DllPllTracking : public TrackingInterface
{
public:
    enum class Signal
    {
        GPS_L1_CA,
        GALILEO_E1,
        GPS_L5
    };
    B200DllPllTracking(Signal signal, const ConfigurationInterface* configuration, std::string role, unsigned int in_streams, unsigned int out_streams);
    ~B200DllPllTracking() override = default;

    std::string role() override { return role_; }
    std::string implementation() override;
    size_t item_size() override { return item_size_; }
    void connect(gr::top_block_sptr top_block) override;
    void disconnect(gr::top_block_sptr top_block) override;
    gr::basic_block_sptr get_left_block() override { return tracking_; }
    gr::basic_block_sptr get_right_block() override { return tracking_; }

    void set_channel(unsigned int channel) override { tracking_->set_channel(channel); }
    void set_gnss_synchro(Gnss_Synchro* p_gnss_synchro) override { tracking_->set_gnss_synchro(p_gnss_synchro); }
    void start_tracking() override { tracking_->start_tracking(); }
    void stop_tracking() override { tracking_->stop_tracking(); }

    //! the implementation names this adapter answers to; nullptr-terminated
    static const char* const* implementations();
    //! true and *signal set when `implementation` is one of them
    static bool lookup(const std::string& implementation, Signal* signal);

private:
    void police_parameters(const ConfigurationInterface* configuration);

    const Signal signal_;
    const std::string role_;
    size_t item_size_;
    Dll_Pll_Conf trk_params_;
    dll_pll_veml_tracking_b200_sptr tracking_;
};

#endif  // GNSS_SDR_B200_DLL_PLL_TRACKING_H
