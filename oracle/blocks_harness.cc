/*
 * TEST INFRASTRUCTURE ONLY.  A flat C API (ctypes) that drives ONE receiver channel - an AcquisitionInterface
 * adapter, a TrackingInterface adapter and the reference's ChannelFsm - over an in-memory sample array with
 * gr::shim::Runner (oracle/shim/gnuradio/shim_runner.h) in place of GNU Radio's scheduler.
 *
 * Compiled twice by oracle/Makefile:
 *   oracle/_ref/liboracle_ref_blocks.so     the reference's OWN blocks and adapters, compiled where they lie under
 *                                           /root/reference (the CPU oracle, and bench.py --impl reference for acquisition)
 *   oracle/_ref/libb200_blocks_check.so     -DHARNESS_B200: additionally the B200 blocks of integration/src (build,
 *                                           link and behaviour check of the drop-in sources; needs libb200gnss.so)
 * Blocks are created BY IMPLEMENTATION STRING, as GNSSBlockFactory does (gnss_block_factory.cc:449-687).
 */
#include "acquisition_interface.h"
#include "channel_fsm.h"
#include "concurrent_queue.h"
#include "galileo_e1_dll_pll_veml_tracking.h"
#include "galileo_e1_pcps_ambiguous_acquisition.h"
#include "gnss_synchro.h"
#include "gps_l1_ca_dll_pll_tracking.h"
#include "gps_l1_ca_pcps_acquisition.h"
#include "gps_l5_dll_pll_tracking.h"
#include "gps_l5i_pcps_acquisition.h"
#include "in_memory_configuration.h"
#include "telemetry_decoder_interface.h"
#include "tracking_interface.h"
#include <gnuradio/block.h>
#include <gnuradio/shim_runner.h>
#include <gnuradio/top_block.h>
#include <pmt/pmt.h>
#include "cpu_multicorrelator.h"
#include "cpu_multicorrelator_16sc.h"
#include "galileo_e1_signal_replica.h"
#include "gps_l5_signal_replica.h"
#include "gps_sdr_signal_replica.h"
#include <array>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <pthread.h>
#include <sched.h>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#ifdef HARNESS_B200
#include "b200_trk_coalescer.h"
#include "gnss_block_factory_b200.h"
#endif

extern "C" int ref_blocks_select_arch(const char* arch);
extern "C" int ref_select_arch(const char* arch);

namespace
{
std::unique_ptr<AcquisitionInterface> make_acq(const std::string& impl, const ConfigurationInterface* cfg, const std::string& role)
{
#ifdef HARNESS_B200
    if (auto b = get_b200_acq_block(impl, cfg, role, 1, 0)) return b;
#endif
    if (impl == "GPS_L1_CA_PCPS_Acquisition") return std::make_unique<GpsL1CaPcpsAcquisition>(cfg, role, 1, 0);
    if (impl == "Galileo_E1_PCPS_Ambiguous_Acquisition") return std::make_unique<GalileoE1PcpsAmbiguousAcquisition>(cfg, role, 1, 0);
    if (impl == "GPS_L5i_PCPS_Acquisition") return std::make_unique<GpsL5iPcpsAcquisition>(cfg, role, 1, 0);
    return nullptr;
}
std::unique_ptr<TrackingInterface> make_trk(const std::string& impl, const ConfigurationInterface* cfg, const std::string& role)
{
#ifdef HARNESS_B200
    if (auto b = get_b200_trk_block(impl, cfg, role, 1, 1)) return b;
#endif
    if (impl == "GPS_L1_CA_DLL_PLL_Tracking") return std::make_unique<GpsL1CaDllPllTracking>(cfg, role, 1, 1);
    if (impl == "Galileo_E1_DLL_PLL_VEML_Tracking") return std::make_unique<GalileoE1DllPllVemlTracking>(cfg, role, 1, 1);
    if (impl == "GPS_L5_DLL_PLL_Tracking") return std::make_unique<GpsL5DllPllTracking>(cfg, role, 1, 1);
    return nullptr;
}

// ChannelFsm::start_acquisition() also resets the channel's telemetry decoder (channel_fsm.cc:202-206)
class NullTelemetry : public TelemetryDecoderInterface
{
public:
    std::string role() override { return "TelemetryDecoder"; }
    std::string implementation() override { return "Null"; }
    size_t item_size() override { return sizeof(Gnss_Synchro); }
    void connect(gr::top_block_sptr) override {}
    void disconnect(gr::top_block_sptr) override {}
    gr::basic_block_sptr get_left_block() override { return nullptr; }
    gr::basic_block_sptr get_right_block() override { return nullptr; }
    void reset() override { resets++; }
    void set_satellite(const Gnss_Satellite&) override {}
    void set_channel(int) override {}
    int resets{0};
};

// the reference FSM with its (private) state made observable: 0 idle, 1 acquiring, 2 tracking
class ObservedFsm : public ChannelFsm
{
public:
    bool Event_valid_acquisition() override
    {
        const bool r = ChannelFsm::Event_valid_acquisition();
        if (r) tracking_started++;
        return r;
    }
    int tracking_started{0};
};

struct Channel
{
    Gnss_Synchro synchro{};
    std::shared_ptr<NullTelemetry> nav;
    std::shared_ptr<AcquisitionInterface> acq;
    std::shared_ptr<TrackingInterface> trk;
    std::shared_ptr<ObservedFsm> fsm;
    std::shared_ptr<Concurrent_Queue<pmt::pmt_t>> queue;
    std::unique_ptr<gr::shim::Runner> acq_runner, trk_runner;
    gr::block* acq_block{nullptr};
    gr::block* trk_block{nullptr};
    size_t acq_item_size{0};
};
}  // namespace

extern "C"
{
    // mirrors the Gnss_Synchro fields acquisition and tracking exchange (gnss_synchro.h:38-82)
    struct itf_synchro
    {
        char System;
        char Signal[3];
        uint32_t PRN;
        int32_t Channel_ID;
        double Acq_delay_samples;
        double Acq_doppler_hz;
        uint64_t Acq_samplestamp_samples;
        uint32_t Acq_doppler_step;
        int32_t Flag_valid_acquisition;
        int64_t fs;
        double Prompt_I;
        double Prompt_Q;
        double CN0_dB_hz;
        double Carrier_Doppler_hz;
        double Carrier_phase_rads;
        double Code_phase_samples;
        uint64_t Tracking_sample_counter;
        int32_t Flag_valid_symbol_output;
        int32_t correlation_length_ms;
        int32_t Flag_PLL_180_deg_phase_locked;
        int32_t pad;
    };

    static void to_flat(const Gnss_Synchro& s, itf_synchro* o)
    {
        std::memset(o, 0, sizeof(*o));
        o->System = s.System;
        std::memcpy(o->Signal, s.Signal, 3);
        o->PRN = s.PRN;
        o->Channel_ID = s.Channel_ID;
        o->Acq_delay_samples = s.Acq_delay_samples;
        o->Acq_doppler_hz = s.Acq_doppler_hz;
        o->Acq_samplestamp_samples = s.Acq_samplestamp_samples;
        o->Acq_doppler_step = s.Acq_doppler_step;
        o->Flag_valid_acquisition = s.Flag_valid_acquisition ? 1 : 0;
        o->fs = s.fs;
        o->Prompt_I = s.Prompt_I;
        o->Prompt_Q = s.Prompt_Q;
        o->CN0_dB_hz = s.CN0_dB_hz;
        o->Carrier_Doppler_hz = s.Carrier_Doppler_hz;
        o->Carrier_phase_rads = s.Carrier_phase_rads;
        o->Code_phase_samples = s.Code_phase_samples;
        o->Tracking_sample_counter = s.Tracking_sample_counter;
        o->Flag_valid_symbol_output = s.Flag_valid_symbol_output ? 1 : 0;
        o->correlation_length_ms = s.correlation_length_ms;
        o->Flag_PLL_180_deg_phase_locked = s.Flag_PLL_180_deg_phase_locked ? 1 : 0;
    }

    int itf_has_b200(void)
    {
#ifdef HARNESS_B200
        return 1;
#else
        return 0;
#endif
    }

    // "generic" or "simd": which volk_gnsssdr implementations the reference blocks dispatch to
    int itf_select_arch(const char* arch)
    {
        const std::string a(arch);
        const int r1 = ref_blocks_select_arch(a == "generic" ? "generic" : "simd");
        const int r2 = ref_select_arch(a == "generic" ? "generic" : "u_avx");
        return (r1 == 0 && r2 == 0) ? 0 : -1;
    }

    void* itf_config_create(void) { return new InMemoryConfiguration(); }
    void itf_config_set(void* cfg, const char* key, const char* value) { static_cast<InMemoryConfiguration*>(cfg)->set_property(key, value); }
    void itf_config_destroy(void* cfg) { delete static_cast<InMemoryConfiguration*>(cfg); }

    // acq_impl / trk_impl may be "" (no such block).  Returns NULL when an implementation name is unknown
    // (GNSSBlockFactory returns nullptr, gnss_block_factory.cc:287-292) or an adapter refuses its item type.
    void* itf_channel_create(void* cfg_, const char* acq_impl, const char* acq_role, const char* trk_impl, const char* trk_role, int channel_id)
    {
        auto* cfg = static_cast<InMemoryConfiguration*>(cfg_);
        auto ch = std::make_unique<Channel>();
        ch->synchro = Gnss_Synchro();
        ch->synchro.Channel_ID = channel_id;
        ch->queue = std::make_shared<Concurrent_Queue<pmt::pmt_t>>();
        try
            {
                if (acq_impl && *acq_impl)
                    {
                        ch->acq = make_acq(acq_impl, cfg, acq_role);
                        if (!ch->acq || ch->acq->item_size() == 0) return nullptr;
                        ch->acq_item_size = ch->acq->item_size();
                        ch->acq->set_channel(channel_id);
                        ch->acq->set_gnss_synchro(&ch->synchro);
                        ch->acq_block = dynamic_cast<gr::block*>(ch->acq->get_right_block().get());
                        ch->acq_runner = std::make_unique<gr::shim::Runner>(ch->acq_block, ch->acq_item_size, sizeof(Gnss_Synchro));
                        ch->acq_runner->set_idle_limit(8);
                    }
                if (trk_impl && *trk_impl)
                    {
                        ch->trk = make_trk(trk_impl, cfg, trk_role);
                        if (!ch->trk || ch->trk->item_size() == 0) return nullptr;
                        ch->trk->set_channel(channel_id);
                        ch->trk->set_gnss_synchro(&ch->synchro);
                        ch->trk_block = dynamic_cast<gr::block*>(ch->trk->get_right_block().get());
                        ch->trk_runner = std::make_unique<gr::shim::Runner>(ch->trk_block, sizeof(gr_complex), sizeof(Gnss_Synchro));
                    }
            }
        catch (const std::exception& e)
            {
                std::fprintf(stderr, "itf_channel_create: %s\n", e.what());
                return nullptr;
            }
        if (ch->acq && ch->trk)
            {
                // the reference's own channel state machine (channel.cc:58-75): a positive acquisition starts tracking
                ch->fsm = std::make_shared<ObservedFsm>();
                ch->nav = std::make_shared<NullTelemetry>();
                ch->fsm->set_acquisition(ch->acq);
                ch->fsm->set_tracking(ch->trk);
                ch->fsm->set_telemetry(ch->nav);
                ch->fsm->set_channel(channel_id);
                ch->fsm->set_queue(ch->queue.get());
                ch->acq->set_channel_fsm(ch->fsm);
            }
        return ch.release();
    }

    void itf_channel_destroy(void* h) { delete static_cast<Channel*>(h); }

    const char* itf_implementation(void* h, int which)
    {
        static thread_local std::string s;
        auto* ch = static_cast<Channel*>(h);
        s = which == 0 ? (ch->acq ? ch->acq->implementation() : "") : (ch->trk ? ch->trk->implementation() : "");
        return s.c_str();
    }

    // Channel::set_signal (channel.cc:160-185): satellite assignment, local code, then the FSM starts acquisition
    int itf_set_satellite(void* h, char system, const char* signal, uint32_t prn)
    {
        auto* ch = static_cast<Channel*>(h);
        ch->synchro.System = system;
        std::memset(ch->synchro.Signal, 0, 3);
        std::strncpy(ch->synchro.Signal, signal, 2);
        ch->synchro.PRN = prn;
        if (ch->acq)
            {
                ch->acq->set_local_code();
            }
        return 0;
    }

    int itf_acq_start(void* h)
    {
        auto* ch = static_cast<Channel*>(h);
        if (!ch->acq) return -1;
        if (ch->fsm)
            ch->fsm->Event_start_acquisition();
        else
            ch->acq->reset();
        return 0;
    }

    int itf_acq_set_doppler_center(void* h, int center)
    {
        auto* ch = static_cast<Channel*>(h);
        if (!ch->acq) return -1;
        ch->acq->set_doppler_center(center);
        return 0;
    }

    // feed n_items (item size = the acquisition's) whose first item is absolute item `nitems_read` of the block
    long itf_acq_run(void* h, const void* samples, uint64_t n_items, long max_calls)
    {
        auto* ch = static_cast<Channel*>(h);
        if (!ch->acq_runner) return -1;
        ch->acq_runner->rebase();
        const long calls = ch->acq_runner->run(samples, n_items, max_calls);
        ch->acq->stop_acquisition();  // joins a worker thread that is still searching; no-op otherwise
        return calls;
    }

    // events published on the block's "events" port so far: 1 positive / 2 negative acquisition, 3 loss of lock
    int itf_events(void* h, int which, int* out, int max)
    {
        auto* ch = static_cast<Channel*>(h);
        gr::block* b = which == 0 ? ch->acq_block : ch->trk_block;
        if (!b) return 0;
        int n = 0;
        for (const auto& pm : b->shim_published())
            if (pm.first == "events" && pmt::is_integer(pm.second) && n < max) out[n++] = static_cast<int>(pmt::to_long(pm.second));
        return n;
    }

    // how many times the FSM went acquisition -> tracking (ChannelFsm::Event_valid_acquisition accepted)
    int itf_fsm_tracking_started(void* h)
    {
        auto* ch = static_cast<Channel*>(h);
        return ch->fsm ? ch->fsm->tracking_started : -1;
    }

    void itf_get_synchro(void* h, itf_synchro* out) { to_flat(static_cast<Channel*>(h)->synchro, out); }

    void itf_set_acq_result(void* h, double delay_samples, double doppler_hz, uint64_t samplestamp)
    {
        auto* ch = static_cast<Channel*>(h);
        ch->synchro.Acq_delay_samples = delay_samples;
        ch->synchro.Acq_doppler_hz = doppler_hz;
        ch->synchro.Acq_samplestamp_samples = samplestamp;
    }

    int itf_trk_start(void* h)
    {
        auto* ch = static_cast<Channel*>(h);
        if (!ch->trk) return -1;
        ch->trk->start_tracking();
        return 0;
    }

    int itf_trk_stop(void* h)
    {
        auto* ch = static_cast<Channel*>(h);
        if (!ch->trk) return -1;
        ch->trk->stop_tracking();
        return 0;
    }

    // telemetry -> tracking message (an int in a pmt any, dll_pll_veml_tracking.cc:756-768): 1 = telemetry fault
    int itf_trk_post_telemetry_event(void* h, int event)
    {
        auto* ch = static_cast<Channel*>(h);
        if (!ch->trk_block) return -1;
        ch->trk_block->shim_post("telemetry_to_trk", pmt::make_any(event));
        return 0;
    }

    // run the tracking block over samples[0 .. n) (samples[0] = absolute sample nitems_read of the block); the
    // Gnss_Synchro items it produced are appended to out (at most max_out); returns how many were produced
    long itf_trk_run(void* h, const void* samples, uint64_t n_items, itf_synchro* out, long max_out, long max_calls)
    {
        auto* ch = static_cast<Channel*>(h);
        if (!ch->trk_runner) return -1;
        ch->trk_runner->rebase();
        ch->trk_runner->clear_outputs();
        ch->trk_runner->run(samples, n_items, max_calls);
        const long n = static_cast<long>(ch->trk_runner->outputs());
        const auto* items = reinterpret_cast<const Gnss_Synchro*>(ch->trk_runner->output_bytes().data());
        for (long i = 0; i < n && i < max_out; i++) to_flat(items[i], &out[i]);
        return n;
    }

    uint64_t itf_nitems_read(void* h, int which)
    {
        auto* ch = static_cast<Channel*>(h);
        gr::block* b = which == 0 ? ch->acq_block : ch->trk_block;
        return b ? b->nitems_read(0) : 0;
    }

    // Several channels at once, one thread each over the same sample array - what GNU Radio's thread-per-block
    // scheduler does with the tracking blocks of a receiver.  out: n_channels x max_out_per_channel items.
    void itf_trk_run_parallel(void** hs, int n_channels, const void* samples, uint64_t n_items, itf_synchro* out, long max_out_per_channel,
        long* n_out)
    {
        std::vector<std::thread> th;
        for (int c = 0; c < n_channels; c++)
            th.emplace_back([=] { n_out[c] = itf_trk_run(hs[c], samples, n_items, out + static_cast<size_t>(c) * max_out_per_channel, max_out_per_channel, -1); });
        for (auto& t : th) t.join();
    }

    // the reference's code generators, for synthesising test signals: tracking replica as a float table
    // (1 value per chip, 2 for the Galileo E1 sinBOC(1,1) replica).  Returns the table length or -1.
    int itf_code_float(char system, const char* signal, uint32_t prn, float* out, int max)
    {
        const std::string sig(signal);
        std::vector<float> t;
        if (system == 'G' && sig == "1C")
            {
                t.resize(1023);
                gps_l1_ca_code_gen_float(t, prn, 0);
            }
        else if (system == 'E' && (sig == "1B" || sig == "1C"))
            {
                t.resize(2 * 4092);
                const std::array<char, 3> s3 = {{sig[0], sig[1], '\0'}};
                galileo_e1_code_gen_sinboc11_float(t, s3, prn);
            }
        else if (system == 'G' && sig == "5I")
            {
                t.resize(10230);
                gps_l5i_code_gen_float(t, prn);
            }
        else if (system == 'G' && sig == "5Q")
            {
                t.resize(10230);
                gps_l5q_code_gen_float(t, prn);
            }
        else
            return -1;
        if (static_cast<int>(t.size()) > max) return -1;
        std::memcpy(out, t.data(), sizeof(float) * t.size());
        return static_cast<int>(t.size());
    }

    // coalescer counters of the B200 build: batches, items, window_expired, mean batch round trip [us], mean and max
    // item latency [us], samples copied, samples offered.  Returns 0, or -1 without a usable GPU / in the reference build.
    int itf_coalescer_stats(double* out8, int reset)
    {
#ifdef HARNESS_B200
        b200::Trk_Coalescer* co = b200::Trk_Coalescer::instance();
        if (co == nullptr) return -1;
        const auto st = co->stats();
        out8[0] = static_cast<double>(st.batches);
        out8[1] = static_cast<double>(st.items);
        out8[2] = static_cast<double>(st.window_expired);
        out8[3] = st.batches ? st.sum_batch_us / static_cast<double>(st.batches) : 0.0;
        out8[4] = st.items ? st.sum_latency_us / static_cast<double>(st.items) : 0.0;
        out8[5] = st.max_latency_us;
        out8[6] = static_cast<double>(st.samples_copied);
        out8[7] = static_cast<double>(st.samples_offered);
        if (reset) co->reset_stats();
        return 0;
#else
        (void)out8;
        (void)reset;
        return -1;
#endif
    }

    // CPU timing of the acquisition chain: `threads` channels, each with its own adapter + block created by implementation
    // string, each running `searches_per_thread` complete searches (Channel::set_signal -> set_local_code -> reset ->
    // general_work until the block reports) on the same samples, PRN = 1 + (thread + k * threads) % 32.  Threads are pinned
    // round robin over the process's CPUs and start together.  Returns elapsed seconds (-1 on error); *positives =
    // number of positive acquisitions.  This is the reference arm of bench.py's acquisition figure.
    double itf_acq_bench(void* cfg, const char* impl, const char* role, char system, const char* signal, int threads, int searches_per_thread,
        const void* samples, uint64_t n_items, int pin, int* positives)
    {
        std::vector<void*> chans(threads, nullptr);
        for (int t = 0; t < threads; t++)
            {
                chans[t] = itf_channel_create(cfg, impl, role, "", "", t);
                if (!chans[t]) return -1.0;
            }
        std::vector<int> cpus;
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0)
            for (int c = 0; c < CPU_SETSIZE; c++)
                if (CPU_ISSET(c, &set)) cpus.push_back(c);
        std::atomic<int> ready{0}, pos{0};
        std::atomic<bool> go{false};
        auto worker = [&](int t) {
            if (pin && !cpus.empty())
                {
                    cpu_set_t one;
                    CPU_ZERO(&one);
                    CPU_SET(cpus[t % cpus.size()], &one);
                    pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
                }
            auto* ch = static_cast<Channel*>(chans[t]);
            ready++;
            while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
            for (int k = 0; k < searches_per_thread; k++)
                {
                    itf_set_satellite(chans[t], system, signal, 1U + static_cast<uint32_t>((t + k * threads) % 32));
                    const size_t before = ch->acq_block->shim_published().size();
                    ch->acq->reset();
                    ch->acq_runner->rebase();
                    ch->acq_runner->run(samples, n_items, -1);
                    ch->acq->stop_acquisition();
                    const auto pub = ch->acq_block->shim_published();
                    for (size_t i = before; i < pub.size(); i++)
                        if (pub[i].first == "events" && pmt::is_integer(pub[i].second) && pmt::to_long(pub[i].second) == 1) pos++;
                }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(worker, t);
        while (ready.load() < threads) std::this_thread::yield();
        const auto t0 = std::chrono::steady_clock::now();
        go.store(true, std::memory_order_release);
        for (auto& x : th) x.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (void* c : chans) itf_channel_destroy(c);
        if (positives) *positives = pos.load();
        return dt;
    }

    // The reference's two other correlator classes, compiled where they lie: one call = init, set_local_code_and_taps,
    // set_input_output_vectors, Carrier_wipeoff_multicorrelator_resampler, free.  arch selects the VG dispatch first.
    int ref_mc_cplx_code(const float* sig_iq, int n, const float* code_iq, int code_len, const float* shifts, int taps, float rem_carr, float dphi,
        float rem_code, float step, float* out_iq)
    {
        Cpu_Multicorrelator mc;
        std::vector<float> sh(shifts, shifts + taps);
        volk_gnsssdr::vector<std::complex<float>> in(n), code(code_len), out(taps);
        std::memcpy(static_cast<void*>(in.data()), sig_iq, sizeof(float) * 2 * n);
        std::memcpy(static_cast<void*>(code.data()), code_iq, sizeof(float) * 2 * code_len);
        mc.init(n, taps);
        mc.set_local_code_and_taps(code_len, code.data(), sh.data());
        mc.set_input_output_vectors(out.data(), in.data());
        mc.Carrier_wipeoff_multicorrelator_resampler(rem_carr, dphi, rem_code, step, n);
        std::memcpy(out_iq, out.data(), sizeof(float) * 2 * taps);
        mc.free();
        return 0;
    }

    int ref_mc_16sc(const int16_t* sig_iq, int n, const int16_t* code_iq, int code_len, const float* shifts, int taps, float rem_carr, float dphi,
        float rem_code, float step, int16_t* out_iq)
    {
        Cpu_Multicorrelator_16sc mc;
        std::vector<float> sh(shifts, shifts + taps);
        volk_gnsssdr::vector<lv_16sc_t> in(n), code(code_len), out(taps);
        std::memcpy(in.data(), sig_iq, sizeof(int16_t) * 2 * n);
        std::memcpy(code.data(), code_iq, sizeof(int16_t) * 2 * code_len);
        mc.init(n, taps);
        mc.set_local_code_and_taps(code_len, code.data(), sh.data());
        mc.set_input_output_vectors(out.data(), in.data());
        mc.Carrier_wipeoff_multicorrelator_resampler(rem_carr, dphi, rem_code, step, n);
        std::memcpy(out_iq, out.data(), sizeof(int16_t) * 2 * taps);
        mc.free();
        return 0;
    }
}
