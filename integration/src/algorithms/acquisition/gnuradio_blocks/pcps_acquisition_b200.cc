/*!
 * \file pcps_acquisition_b200.cc
 * \brief see pcps_acquisition_b200.h.  Line references are to
 * src/algorithms/acquisition/gnuradio_blocks/pcps_acquisition.cc.
 */
#include "pcps_acquisition_b200.h"
#include "b200gnss.h"
#include "gnss_sdr_create_directory.h"
#include "gnss_sdr_filesystem.h"
#include "gnss_synchro.h"
#include <gnuradio/io_signature.h>
#include <pmt/pmt.h>
#include <pmt/pmt_sugar.h>
#include <algorithm>
#include <cmath>
#include <iostream>

#if USE_GLOG_AND_GFLAGS
#include <glog/logging.h>
#else
#include <absl/log/log.h>
#endif

namespace
{
b200::Acq_Conf_Core core_conf(const Acq_Conf& c)
{
    b200::Acq_Conf_Core k;
    k.fs_in = c.use_automatic_resampler ? c.resampled_fs : c.fs_in;  // the wipe-off grid is built for the stream's rate (:277)
    k.samples_per_ms = c.samples_per_ms;
    k.samples_per_code = c.samples_per_code;
    k.samples_per_chip = c.samples_per_chip;
    k.sampled_ms = c.sampled_ms;
    k.ms_per_code = c.ms_per_code;
    k.doppler_max = static_cast<uint32_t>(c.doppler_max);
    k.doppler_step = static_cast<uint32_t>(c.doppler_step);
    k.max_dwells = c.max_dwells;
    k.pfa = c.pfa;
    k.threshold = c.threshold;
    k.bit_transition_flag = c.bit_transition_flag;
    k.use_CFAR_algorithm_flag = c.use_CFAR_algorithm_flag;
    k.make_2_steps = c.make_2_steps;
    k.doppler_step2 = c.doppler_step2;
    k.num_doppler_bins_step2 = c.num_doppler_bins_step2;
    k.pfa2 = c.pfa2;
    k.dump = c.dump;
    return k;
}

// <dir>/<name without extension>, default name "acquisition" (get_dump_filename, :59-97)
std::string dump_base(const std::string& configured)
{
    std::string name = configured.empty() ? std::string("acquisition") : configured;
    std::string dir = ".";
    const auto slash = name.find_last_of('/');
    if (slash != std::string::npos)
        {
            dir = name.substr(0, slash);
            name = name.substr(slash + 1);
        }
    if (name.empty()) name = "acquisition";
    if (name.substr(1).find_last_of('.') != std::string::npos) name = name.substr(0, name.find_last_of('.'));
    if (!gnss_sdr_create_directory(dir))
        {
            std::cerr << "GNSS-SDR cannot create dump file for the Acquisition block. Wrong permissions?\n";
            return {};
        }
    return dir + fs::path::preferred_separator + name;
}
}  // namespace


pcps_acquisition_b200_sptr pcps_make_acquisition_b200(const Acq_Conf& conf_)
{
    return pcps_acquisition_b200_sptr(new pcps_acquisition_b200(conf_));
}


pcps_acquisition_b200::pcps_acquisition_b200(const Acq_Conf& conf_)
    : acquisition_impl_interface("pcps_acquisition_b200", gr::io_signature::make(1, 1, conf_.it_size),
          gr::io_signature::make(0, 1, sizeof(Gnss_Synchro))),
      d_conf(conf_),
      d_cshort(conf_.it_size != sizeof(gr_complex)),
      d_block_samples(static_cast<uint32_t>(conf_.sampled_ms * conf_.samples_per_ms * (conf_.bit_transition_flag ? 2.0 : 1.0))),
      d_core(std::make_unique<b200::Pcps_Acquisition_Core>(core_conf(conf_))),
      d_resampler_latency_samples(conf_.resampler_latency_samples)
{
    this->message_port_register_out(pmt::mp("events"));
    d_core->set_gnss_synchro(&d_bridge);
    if (d_cshort)
        d_block_sc.resize(d_block_samples);
    else
        d_block.resize(d_block_samples);
    if (!d_core->ok())
        {
            LOG(ERROR) << "pcps_acquisition_b200: no usable B200 or unsupported FFT size (" << b200_last_error()
                       << "); every acquisition will be reported negative";
        }
    if (conf_.dump) d_dump_base = dump_base(conf_.dump_filename);
}


pcps_acquisition_b200::~pcps_acquisition_b200()
{
    join_worker();
}


void pcps_acquisition_b200::join_worker()
{
    std::unique_ptr<gr::thread::thread> worker;
    {
        gr::thread::scoped_lock lk(d_setlock);
        worker = std::move(d_worker);
    }
    if (worker && worker->joinable()) worker->join();
}


void pcps_acquisition_b200::set_gnss_synchro(Gnss_Synchro* p_gnss_synchro)
{
    gr::thread::scoped_lock lock(d_setlock);
    d_gnss_synchro = p_gnss_synchro;
}


void pcps_acquisition_b200::set_active(bool active)
{
    {
        gr::thread::scoped_lock lock(d_setlock);
        d_active = active;
        d_core->set_active(active);
    }
    if (!active) join_worker();
}


void pcps_acquisition_b200::set_resampler_latency(uint32_t latency_samples)
{
    gr::thread::scoped_lock lock(d_setlock);
    d_resampler_latency_samples = latency_samples;
}


void pcps_acquisition_b200::set_local_code(std::complex<float>* code)
{
    gr::thread::scoped_lock lock(d_setlock);
    d_core->set_local_code(code);  // placement variants, forward FFT and conjugate (:218-251) on the device
}


void pcps_acquisition_b200::set_doppler_center(int32_t doppler_center)
{
    gr::thread::scoped_lock lock(d_setlock);
    d_core->set_doppler_center(doppler_center);  // rebuilds the wipe-off grid only when the centre changes
}


// Gnss_Synchro <- what the core computed (update_synchro, :580-602, including the resampler compensation)
void pcps_acquisition_b200::copy_synchro_out()
{
    if (d_gnss_synchro == nullptr) return;
    d_gnss_synchro->Acq_delay_samples = d_bridge.Acq_delay_samples;
    d_gnss_synchro->Acq_doppler_hz = d_bridge.Acq_doppler_hz;
    if (d_conf.use_automatic_resampler)
        {
            d_gnss_synchro->Acq_delay_samples = (d_gnss_synchro->Acq_delay_samples * d_conf.resampler_ratio) - static_cast<double>(d_resampler_latency_samples);
            d_gnss_synchro->Acq_samplestamp_samples = rint(static_cast<double>(d_bridge.Acq_samplestamp_samples) * d_conf.resampler_ratio);
            d_gnss_synchro->fs = d_conf.resampled_fs;
        }
    else
        {
            d_gnss_synchro->Acq_samplestamp_samples = d_bridge.Acq_samplestamp_samples;
            d_gnss_synchro->fs = d_conf.fs_in;
        }
    if (d_bridge.Acq_doppler_step != 0U) d_gnss_synchro->Acq_doppler_step = d_bridge.Acq_doppler_step;
}


void pcps_acquisition_b200::publish(int event, const b200::AcquisitionResult& result)
{
    DLOG(INFO) << (event == 1 ? "positive" : "negative") << " acquisition, satellite " << d_gnss_synchro->System << " " << d_gnss_synchro->PRN
               << ", sample_stamp " << result.sample_count << ", test statistics value " << result.test_statistics
               << ", test statistics threshold " << d_core->get_threshold() << ", code phase " << d_gnss_synchro->Acq_delay_samples
               << ", doppler " << static_cast<double>(result.doppler) << ", input signal power " << d_core->d_input_power;
    if (event == 1)
        {
            if (auto fsm = d_channel_fsm.lock())
                {
                    fsm->Event_valid_acquisition();  // straight to the channel state machine: no message latency
                }
            else
                {
                    this->message_port_pub(pmt::mp("events"), pmt::from_long(1));
                }
            LOG(INFO) << "Successful acquisition in channel " << d_channel << " for satellite " << d_gnss_synchro->System << " " << d_gnss_synchro->PRN;
            if (d_conf.enable_monitor_output) d_monitor_queue.push(*d_gnss_synchro);
        }
    else if (event == 2)
        {
            this->message_port_pub(pmt::mp("events"), pmt::from_long(2));
        }
}


// The variables of pcps_acquisition::dump_results (:354-406), written by the library (MAT Level 5).
void pcps_acquisition_b200::dump_grid(const b200::AcquisitionResult& result)
{
    if (d_dump_base.empty() || d_channel != d_conf.dump_channel) return;
    d_dump_number++;
    std::vector<float> grid(static_cast<size_t>(d_core->d_num_doppler_bins) * d_core->d_effective_fft_size);
    if (!d_core->read_grid(grid.data())) return;
    char name[1024];
    const char sig[3] = {d_gnss_synchro->Signal[0], d_gnss_synchro->Signal[1], '\0'};
    if (b200_acq_dump_filename(d_dump_base.c_str(), d_gnss_synchro->System, sig, d_channel, static_cast<uint32_t>(d_dump_number), d_gnss_synchro->PRN,
            name, sizeof(name)) != B200_OK)
        return;
    b200_acq_dump d{};
    d.acq_grid = grid.data();
    d.effective_fft_size = d_core->d_effective_fft_size;
    d.num_doppler_bins = d_core->d_num_doppler_bins;
    d.doppler_max = d_conf.doppler_max;
    d.doppler_step = d_conf.doppler_step;
    d.positive_acq = result.positive_acq ? 1 : 0;
    d.num_dwells = static_cast<int32_t>(d_conf.max_dwells);
    d.prn = d_gnss_synchro->PRN;
    d.acq_doppler_hz = static_cast<float>(d_gnss_synchro->Acq_doppler_hz);
    d.acq_delay_samples = static_cast<float>(d_gnss_synchro->Acq_delay_samples);
    d.test_statistic = result.test_statistics;
    d.threshold = d_core->get_threshold();
    d.input_power = d_core->d_input_power;
    d.sample_counter = result.sample_count;
    if (b200_acq_dump_write(name, &d) != B200_OK) std::cout << "Unable to create or open Acquisition dump file\n";
}


// acquisition_core (:648-728): one coherent search on the collected block, then the decision.
void pcps_acquisition_b200::search(uint64_t sample_count)
{
    gr::thread::scoped_lock lk(d_setlock);
    const uint32_t dwell_before = d_core->dwell_counter();
    d_last_dwell_running = (dwell_before + 1 == d_conf.max_dwells);
    b200::AcquisitionResult result;
    const bool was_step_two = d_core->step_two();
    lk.unlock();
    // the grid search does not touch anything set_* can change: run it without the lock, like the reference (:678-683)
    const int event = d_cshort ? d_core->acquisition_core_i16(reinterpret_cast<const int16_t*>(d_block_sc.data()), sample_count, &result)
                               : d_core->acquisition_core(d_block.data(), sample_count, &result);
    lk.lock();
    copy_synchro_out();
    (void)was_step_two;
    d_active = d_core->active();
    d_phase = d_core->state() == 0 ? RESET_SYNCHRO : FILL_BUFFER;
    d_filled = 0;
    if (event != 0) publish(event, result);
    if (event != 0 || d_conf.bit_transition_flag) dump_grid(result);
    d_worker_active = false;
}


int pcps_acquisition_b200::general_work(int noutput_items __attribute__((unused)), gr_vector_int& ninput_items,
    gr_vector_const_void_star& input_items, gr_vector_void_star& output_items)
{
    gr::thread::scoped_lock lk(d_setlock);
    if (!d_active || d_worker_active)
        {
            // idle, or a search is running: samples flow past (counted) unless this is an intermediate dwell, which
            // must see contiguous blocks (:765-775)
            const bool drain = !d_active || (d_worker_active && d_last_dwell_running);
            if (!d_conf.blocking_on_standby && drain)
                {
                    d_sample_count += static_cast<uint64_t>(ninput_items[0]);
                    consume_each(ninput_items[0]);
                }
            return 0;
        }

    switch (d_phase)
        {
        case RESET_SYNCHRO:
            d_gnss_synchro->Acq_delay_samples = 0.0;
            d_gnss_synchro->Acq_doppler_hz = 0.0;
            d_gnss_synchro->Acq_samplestamp_samples = 0ULL;
            d_gnss_synchro->Acq_doppler_step = 0U;
            d_bridge = b200::Acq_Synchro();
            d_phase = FILL_BUFFER;
            d_filled = 0U;
            break;

        case FILL_BUFFER:
            {
                const uint32_t room = d_block_samples - d_filled;
                const uint32_t take = std::min<uint32_t>(room, static_cast<uint32_t>(ninput_items[0]));
                if (d_cshort)
                    {
                        const auto* in = reinterpret_cast<const lv_16sc_t*>(input_items[0]);
                        std::copy(in, in + take, d_block_sc.begin() + d_filled);
                    }
                else
                    {
                        const auto* in = reinterpret_cast<const gr_complex*>(input_items[0]);
                        std::copy(in, in + take, d_block.begin() + d_filled);
                    }
                d_filled += take;
                d_sample_count += static_cast<uint64_t>(take);
                consume_each(static_cast<int>(take));
                if (d_filled == d_block_samples) d_phase = SEARCH;
                break;
            }

        case SEARCH:
            if (d_conf.blocking)
                {
                    lk.unlock();
                    search(d_sample_count);
                    lk.lock();
                }
            else
                {
                    lk.unlock();
                    join_worker();
                    lk.lock();
                    d_worker_active = true;
                    d_last_dwell_running = (d_core->dwell_counter() + 1 == d_conf.max_dwells);
                    d_worker = std::make_unique<gr::thread::thread>(&pcps_acquisition_b200::search, this, d_sample_count);
                }
            consume_each(0);
            break;
        }

    if (d_conf.enable_monitor_output && !d_monitor_queue.empty())
        {
            auto** out = reinterpret_cast<Gnss_Synchro**>(&output_items[0]);
            const int n = static_cast<int>(d_monitor_queue.size());
            for (int i = 0; i < n; ++i)
                {
                    *out[i] = d_monitor_queue.front();
                    d_monitor_queue.pop();
                }
            return n;
        }
    return 0;
}
