/*!
 * \file pcps_acquisition_b200.h
 * \brief Parallel Code Phase Search acquisition block whose Doppler x code-phase grid search runs on a B200 GPU.
 *
 * Drop-in sibling of pcps_acquisition (same directory in the gnss-sdr tree): derives from the same
 * acquisition_impl_interface, same Acq_Conf, same stream signature (in: gr_complex or lv_16sc_t, optional
 * Gnss_Synchro monitor output), same "events" message port (1 = positive, 2 = negative acquisition) and direct
 * ChannelFsm::Event_valid_acquisition() notification.  The sample buffering and the decision state machine of
 * general_work (pcps_acquisition.cc:749-853) live here; the arithmetic and the thresholds live in
 * b200::Pcps_Acquisition_Core over libb200gnss.so (wipe-off x FFT x conj(code) x IFFT x |.|^2 and the peak
 * statistics fused on the device; the magnitude grid never leaves the GPU unless a dump asks for it).
 */
#ifndef GNSS_SDR_PCPS_ACQUISITION_B200_H
#define GNSS_SDR_PCPS_ACQUISITION_B200_H

#include "acq_conf.h"
#include "acquisition_impl_interface.h"
#include "b200_pcps_acquisition_core.h"
#include "channel_fsm.h"
#include <gnuradio/block.h>
#include <gnuradio/gr_complex.h>
#include <gnuradio/thread/thread.h>
#include <gnuradio/types.h>
#include <volk_gnsssdr/volk_gnsssdr_complex.h>  // lv_16sc_t
#include <complex>
#include <cstdint>
#include <memory>
#include <queue>
#include <string>
#include <vector>

class Gnss_Synchro;
class pcps_acquisition_b200;

using pcps_acquisition_b200_sptr = gnss_shared_ptr<pcps_acquisition_b200>;

pcps_acquisition_b200_sptr pcps_make_acquisition_b200(const Acq_Conf& conf_);

class pcps_acquisition_b200 : public acquisition_impl_interface
{
public:
    ~pcps_acquisition_b200() override;

    void set_gnss_synchro(Gnss_Synchro* p_gnss_synchro) override;
    void set_local_code(std::complex<float>* code) override;
    void set_resampler_latency(uint32_t latency_samples);
    uint32_t mag() const override { return 0; }
    void set_active(bool active) override;
    void set_channel(uint32_t channel) override { d_channel = channel; }
    void set_channel_fsm(std::weak_ptr<ChannelFsm> channel_fsm) override { d_channel_fsm = std::move(channel_fsm); }
    void set_doppler_center(int32_t doppler_center);

    int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
        gr_vector_void_star& output_items) override;

private:
    friend pcps_acquisition_b200_sptr pcps_make_acquisition_b200(const Acq_Conf& conf_);
    explicit pcps_acquisition_b200(const Acq_Conf& conf_);

    enum Phase
    {
        RESET_SYNCHRO = 0,  // reference d_state 0
        FILL_BUFFER = 1,    // reference d_state 1
        SEARCH = 2          // reference d_state 2
    };
    void search(uint64_t sample_count);  // acquisition_core(): runs on the scheduler thread (blocking) or on d_worker
    void publish(int event, const b200::AcquisitionResult& result);
    void copy_synchro_out();
    void dump_grid(const b200::AcquisitionResult& result);
    void join_worker();

    const Acq_Conf d_conf;
    const bool d_cshort;
    const uint32_t d_block_samples;  // d_consumed_samples: samples per coherent search
    std::unique_ptr<b200::Pcps_Acquisition_Core> d_core;
    b200::Acq_Synchro d_bridge;  // the Gnss_Synchro fields the core writes

    // guarded by d_setlock
    std::weak_ptr<ChannelFsm> d_channel_fsm;
    std::unique_ptr<gr::thread::thread> d_worker;
    Gnss_Synchro* d_gnss_synchro{nullptr};
    std::queue<Gnss_Synchro> d_monitor_queue;
    std::vector<gr_complex> d_block;      // the samples being collected
    std::vector<lv_16sc_t> d_block_sc;    // same for cshort input
    Phase d_phase{RESET_SYNCHRO};
    uint32_t d_filled{0};
    uint32_t d_channel{0};
    uint32_t d_resampler_latency_samples{0};
    uint64_t d_sample_count{0};
    int64_t d_dump_number{0};
    bool d_active{false};
    bool d_worker_active{false};
    bool d_last_dwell_running{false};
    std::string d_dump_base;
};

#endif  // GNSS_SDR_PCPS_ACQUISITION_B200_H
