/*!
 * \file b200_pcps_acquisition_core.h
 * \brief The arithmetic and decision logic of pcps_acquisition, without GNU Radio, on a B200.
 *
 * pcps_acquisition (src/algorithms/acquisition/gnuradio_blocks/pcps_acquisition.{h,cc}) is a
 * gr::block; GNU Radio is not available in this build environment, so this class holds
 * everything of that block that is NOT GNU Radio plumbing, with the same member names where
 * they exist upstream:
 *   constructor sizes      pcps_acquisition.cc:100-193   set_local_code      :218-251
 *   set_doppler_center     pcps_acquisition.h:188-200    acquisition_core    :648-728
 *   compute_threshold      :52-56                        update_synchro      :580-602
 * The gr::block shell (general_work buffering :749-853, message ports, ChannelFsm notification
 * :318-351) is shown in INTEGRATION.md; it calls acquisition_core() below exactly where the
 * reference calls its own.
 */
#ifndef B200_PCPS_ACQUISITION_CORE_H
#define B200_PCPS_ACQUISITION_CORE_H

#include <complex>
#include <cstdint>
#include <string>

struct b200_acq;

namespace b200
{
/*! The Acq_Conf fields the arithmetic reads (src/algorithms/acquisition/libs/acq_conf.h:33-87). */
struct Acq_Conf_Core
{
    int64_t fs_in{4000000};
    float samples_per_ms{4000.0F};
    float samples_per_code{4000.0F};
    uint32_t samples_per_chip{4};
    uint32_t sampled_ms{1};
    uint32_t ms_per_code{1};
    uint32_t doppler_max{5000};
    uint32_t doppler_step{250};
    uint32_t max_dwells{1};
    float pfa{0.0F};
    float threshold{0.0F};
    bool bit_transition_flag{false};
    bool use_CFAR_algorithm_flag{true};
    bool make_2_steps{false};             // acq_conf.h:74
    float doppler_step2{125.0F};          // :50
    uint32_t num_doppler_bins_step2{4U};  // :62
    float pfa2{0.0F};
    bool dump{false};  // keeps the magnitude grid on the device so read_grid() works
};

/*! The Gnss_Synchro fields acquisition writes (src/core/system_parameters/gnss_synchro.h:50-56). */
struct Acq_Synchro
{
    double Acq_delay_samples{0.0};
    double Acq_doppler_hz{0.0};
    uint64_t Acq_samplestamp_samples{0};
    uint32_t Acq_doppler_step{0};
    int64_t fs{0};
};

struct AcquisitionResult  // pcps_acquisition.h:213-220
{
    uint32_t index_time{0};
    int32_t doppler{0};
    float test_statistics{0.0F};
    uint64_t sample_count{0};
    bool positive_acq{false};
};

float compute_threshold(float pfa, uint32_t effective_fft_size, uint32_t num_doppler_bins, uint32_t max_dwells);

class Pcps_Acquisition_Core
{
public:
    explicit Pcps_Acquisition_Core(const Acq_Conf_Core& conf);
    ~Pcps_Acquisition_Core();
    Pcps_Acquisition_Core(const Pcps_Acquisition_Core&) = delete;
    Pcps_Acquisition_Core& operator=(const Pcps_Acquisition_Core&) = delete;

    bool ok() const { return d_acq != nullptr; }
    void set_gnss_synchro(Acq_Synchro* p_gnss_synchro) { d_gnss_synchro = p_gnss_synchro; }
    void set_local_code(std::complex<float>* code);
    void set_doppler_center(int32_t doppler_center);
    void set_threshold(float threshold) { d_threshold = threshold; }
    float get_threshold() const { return d_step_two ? d_threshold_step_two : d_threshold; }  // :731-734
    void set_active(bool active);
    void init();  // pcps_acquisition::init (:196-215): reset counters and synchro fields
    uint32_t mag() const { return 0; }

    /*! One call of pcps_acquisition::acquisition_core(sample_count) on d_consumed_samples input
     *  samples.  Returns the event the block would emit: 1 positive, 2 negative, 0 none yet
     *  (more dwells needed). */
    int acquisition_core(const std::complex<float>* in, uint64_t sample_count, AcquisitionResult* out);
    /*! the same on cshort input (Acq_Conf::it_size == sizeof(lv_16sc_t), pcps_acquisition.cc:653-656): 2 x d_consumed_samples
     *  int16, converted to float on the device */
    int acquisition_core_i16(const int16_t* in_iq, uint64_t sample_count, AcquisitionResult* out);
    //! state the gr::block shell needs (pcps_acquisition.cc:749-853)
    bool active() const { return d_active; }
    int state() const { return d_state; }
    bool step_two() const { return d_step_two; }
    uint32_t dwell_counter() const { return d_num_noncoherent_integrations_counter; }

    bool read_grid(float* grid) const;  // d_magnitude_grid, bins x effective_fft_size

    uint32_t d_consumed_samples;
    uint32_t d_fft_size;
    uint32_t d_effective_fft_size;
    uint32_t d_num_doppler_bins;
    float d_input_power{0.0F};

private:
    void update_synchro(const AcquisitionResult& result);
    int acquisition_core_any(const void* in, bool cshort, uint64_t sample_count, AcquisitionResult* out);

    Acq_Conf_Core d_acq_parameters;
    b200_acq* d_acq{nullptr};
    Acq_Synchro* d_gnss_synchro{nullptr};
    float d_threshold{0.0F};
    float d_threshold_step_two{0.0F};
    float d_doppler_center_step_two{0.0F};
    bool d_step_two{false};
    int32_t d_doppler_center{0};
    uint32_t d_num_noncoherent_integrations_counter{0};
    int d_state{0};
    bool d_active{false};
};
}  // namespace b200
#endif
