/*!
 * \file b200_pcps_acquisition_fine_doppler_core.h
 * \brief The arithmetic and state machine of pcps_acquisition_fine_doppler_cc, without GNU Radio, on a B200.
 *
 * src/algorithms/acquisition/gnuradio_blocks/pcps_acquisition_fine_doppler_cc.{h,cc} (the acquisition the default
 * conf/gnss-sdr.conf selects).  Member names follow the block:
 *   constructor sizes :44-70          set_local_code :130-136          reset_grid :149-160
 *   compute_and_accumulate_grid :266-299 and compute_CAF :182-251  -> b200_acq (dwell accumulation, first/second peak)
 *   estimate_Doppler :316-389 -> b200_acq_fine (+ replica alignment, frequency mapping and 1 kHz check here)
 *   general_work states 0-5 :400-557  -> work()
 * The gr::block shell (forecast, consume_each, message port "events", monitor output) stays in the gnss-sdr tree.
 */
#ifndef B200_PCPS_ACQUISITION_FINE_DOPPLER_CORE_H
#define B200_PCPS_ACQUISITION_FINE_DOPPLER_CORE_H

#include "b200_pcps_acquisition_core.h"  // Acq_Synchro

#include <complex>
#include <cstdint>
#include <vector>

struct b200_acq;
struct b200_acq_fine;

namespace b200
{
struct Fine_Doppler_Conf  // the Acq_Conf fields this block reads
{
    int64_t fs_in{4000000};
    float samples_per_ms{4000.0F};
    int32_t doppler_max{5000};
    int32_t doppler_step{250};
    uint32_t max_dwells{1};
    float threshold{0.0F};
};

class Pcps_Acquisition_Fine_Doppler_Core
{
public:
    explicit Pcps_Acquisition_Fine_Doppler_Core(const Fine_Doppler_Conf& conf);
    ~Pcps_Acquisition_Fine_Doppler_Core();
    Pcps_Acquisition_Fine_Doppler_Core(const Pcps_Acquisition_Fine_Doppler_Core&) = delete;
    Pcps_Acquisition_Fine_Doppler_Core& operator=(const Pcps_Acquisition_Fine_Doppler_Core&) = delete;

    bool ok() const { return d_acq != nullptr && d_fine != nullptr; }
    void set_gnss_synchro(Acq_Synchro* p) { d_gnss_synchro = p; }
    void set_local_code(std::complex<float>* code);  // gps_l1_ca_code_gen_complex_sampled output, d_fft_size values
    void set_active(bool active) { d_active = active; }
    void reset() { d_state = 0; }
    float test_statistics() const { return d_test_statistics; }
    uint32_t fine_index() const { return d_tmp_index_freq; }

    /*! One general_work call with noutput_items = d_fft_size available input samples.  *consumed = what the block
     *  passes to consume_each.  Returns the event: 0 none, 1 positive acquisition, 2 negative acquisition. */
    int work(const std::complex<float>* in, int noutput_items, int* consumed);

    int32_t d_fft_size;
    int32_t d_num_doppler_points;

private:
    int estimate_Doppler();

    Fine_Doppler_Conf d_acq_params;
    b200_acq* d_acq{nullptr};
    b200_acq_fine* d_fine{nullptr};
    Acq_Synchro* d_gnss_synchro{nullptr};
    std::vector<std::complex<float>> d_10_ms_buffer;
    std::vector<std::complex<float>> d_code;
    uint64_t d_sample_counter{0};
    float d_test_statistics{0.0F};
    uint32_t d_tmp_index_freq{0};
    int32_t d_state{0};
    int32_t d_well_count{0};
    int32_t d_n_samples_in_buffer{0};
    bool d_active{false};
};
}  // namespace b200
#endif
