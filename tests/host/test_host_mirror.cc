// Parity test of the C++ host mirror classes against the reference's own class, written the way
// the reference's engine test is (tests/unit-tests/signal-processing-blocks/tracking/
// cpu_multicorrelator_real_codes_test.cc:65-180: same sizes 2048/4096/8192, 3 taps at +-0.5 chip,
// PRN-like code, uniform-random IQ, concurrent correlator objects on std::threads) -- except that
// it CHECKS the E/P/L values instead of only timing them.
//
// Build (tests/test_host_mirror.py does this):
//   g++ -std=c++17 tests/host/test_host_mirror.cc gnss_sdr_b200/host/*.cc -Iinclude -Ignss_sdr_b200/host
//       -Lgnss_sdr_b200 -lb200gnss [-DHAVE_REF oracle/_ref/liboracle_ref.so] -lpthread
#include "b200_multicorrelator_real_codes.h"
#include "b200_pcps_acquisition_core.h"
#include "b200_dll_pll_veml_loop.h"
#include "b200_pcps_acquisition_fine_doppler_core.h"
#include "b200_trk_coalescer.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <string>
#include <thread>
#include <vector>

#ifdef HAVE_REF
extern "C"
{
    void* ref_mc_create(int max_len, int taps, int high_dyn);
    int ref_mc_set_code(void* h, const float* code, int code_len, const float* shifts);
    int ref_mc_correlate(void* h, const float* in_iq, float rem_carr_rad, float phase_step_rad, float phase_rate_step_rad,
        float rem_code_chips, float code_step_chips, float code_rate_step_chips, int n, float* out_taps);
    void ref_mc_destroy(void* h);
    float* ref_mc_shifts(void* h);
    int ref_select_arch(const char* arch);
}
#endif

static int g_fail = 0;
#define CHECK(cond, ...)                      \
    do                                        \
        {                                     \
            if (!(cond))                      \
                {                             \
                    std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
                    std::printf(__VA_ARGS__); \
                    std::printf("\n");        \
                    g_fail++;                 \
                }                             \
        }                                     \
    while (0)

static void correlator_worker(int tid, int n, int iters, const std::vector<float>* code, const std::vector<std::complex<float>>* in,
    std::vector<std::complex<float>>* out)
{
    B200_Multicorrelator_Real_Codes mc;
    float shifts[3] = {-0.5F, 0.0F, 0.5F};
    if (!mc.init(8192, 3))
        {
            std::printf("init failed: %s\n", mc.last_error());
            g_fail++;
            return;
        }
    mc.set_high_dynamics_resampler(false);
    mc.set_local_code_and_taps(static_cast<int>(code->size()), code->data(), shifts);
    mc.set_input_output_vectors(out->data() + 3 * tid, in->data());
    for (int k = 0; k < iters; k++)
        {
            // cpu_multicorrelator_real_codes_test.cc:129-133 parameters
            if (!mc.Carrier_wipeoff_multicorrelator_resampler(0.0F, 0.1F, 0.0F, 0.4F, 0.3F, 0.0F, n)) g_fail++;
        }
    mc.free();
}

// The tracking block rewrites its tap-shift array IN PLACE after set_local_code_and_taps and never calls it again
// (dll_pll_veml_tracking.cc:1030 then :1045-1053 in start_tracking; the narrow-correlator switch :2132-2146).  The
// CPU class keeps the pointer, so it follows; the B200 class must too - on the synchronous and on the coalesced path.
static void shift_pointer_semantics(const std::vector<float>& code, const std::vector<std::complex<float>>& in)
{
    const int n = 4000;
    volatile float wide = 0.5F, narrow = 0.15F;
    for (int coalesced = 0; coalesced < 2; coalesced++)
        {
            float shifts[3] = {-wide, 0.0F, wide};
            B200_Multicorrelator_Real_Codes mc;
            std::complex<float> out_wide[3], out_narrow[3], out_back[3];
            CHECK(mc.init(8192, 3), "init");
            mc.set_high_dynamics_resampler(false);
            mc.set_local_code_and_taps(1023, code.data(), shifts);
            auto run = [&](std::complex<float>* out, uint64_t pos) {
                mc.set_input_output_vectors(out, in.data() + pos);
                if (coalesced) mc.set_stream_position(7, 1000000ULL + pos, n);
                return mc.Carrier_wipeoff_multicorrelator_resampler(0.3F, 0.01F, 0.0F, 0.4F, 0.2557F, 0.0F, n);
            };
            CHECK(run(out_wide, 0), "correlate wide: %s", mc.last_error());
            shifts[0] = -narrow;  // in place, no set_local_code_and_taps
            shifts[2] = narrow;
            CHECK(run(out_narrow, 0), "correlate narrow: %s", mc.last_error());
            shifts[0] = -wide;  // start_tracking restores the wide spacing the same way
            shifts[2] = wide;
            CHECK(run(out_back, 0), "correlate wide again");
            CHECK(out_back[0] == out_wide[0] && out_back[2] == out_wide[2], "wide taps after the round trip differ (coalesced=%d)", coalesced);
            CHECK(std::abs(out_narrow[0] - out_wide[0]) > 1e-3F * std::abs(out_wide[1]), "early tap did not move with the shift array (coalesced=%d)", coalesced);
            CHECK(out_narrow[1] == out_wide[1], "prompt must not change");
#ifdef HAVE_REF
            ref_select_arch("a_avx");
            void* h = ref_mc_create(8192, 3, 0);
            float s0[3] = {-wide, 0.0F, wide};
            ref_mc_set_code(h, code.data(), 1023, s0);
            std::complex<float> want[3];
            float* kept = ref_mc_shifts(h);
            kept[0] = -narrow;
            kept[2] = narrow;
            ref_mc_correlate(h, reinterpret_cast<const float*>(in.data()), 0.3F, 0.01F, 0.0F, 0.4F, 0.2557F, 0.0F, n, reinterpret_cast<float*>(want));
            for (int k = 0; k < 3; k++)
                {
                    const float rel = std::abs(out_narrow[k] - want[k]) / std::abs(want[1]);
                    CHECK(rel < 1e-3F, "narrow taps vs reference class: tap %d rel %g (coalesced=%d)", k, rel, coalesced);
                }
            ref_mc_destroy(h);
#endif
            mc.free();
        }
    std::printf("shift-pointer semantics: ok\n");
}

// N block threads drive the CLASS interface in coalesced mode on one band (what N tracking blocks of a flowgraph do):
// samples cross PCIe once, the epochs of all channels share launches.  Shape of the reference's engine benchmark
// (cpu_multicorrelator_real_codes_test.cc:135-169: N concurrent correlators, each calling in a loop), C2 sizes.
static void coalescer_throughput(int n_threads, int epochs, int window_us)
{
    const int n = 25000;  // 1 ms at 25 Msps
    const uint64_t total = static_cast<uint64_t>(n) * (epochs + 2) + n;
    std::vector<std::complex<float>> iq(total);
    std::mt19937 rng(3);
    std::uniform_real_distribution<float> ud(-1.0F, 1.0F);
    for (auto& v : iq) v = std::complex<float>(ud(rng), ud(rng));
    std::vector<float> code(1023);
    for (auto& c : code) c = (rng() & 1U) ? 1.0F : -1.0F;
    b200::Trk_Coalescer* co = b200::Trk_Coalescer::instance();
    if (co == nullptr)
        {
            std::printf("no GPU: coalescer benchmark skipped\n");
            return;
        }
    co->set_window_us(window_us);
    co->ensure_band(3, 1ULL << 24);
    co->reset_stats();
    std::vector<double> lat_sum(n_threads, 0.0), lat_max(n_threads, 0.0);
    std::atomic<int> failed{0};
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    const uint64_t base = 0;
    auto worker = [&](int t) {
        B200_Multicorrelator_Real_Codes mc;
        float shifts[3] = {-0.5F, 0.0F, 0.5F};
        std::complex<float> out[3];
        if (!mc.init(2 * n, 3))
            {
                failed++;
                return;
            }
        mc.set_high_dynamics_resampler(false);
        mc.set_local_code_and_taps(1023, code.data(), shifts);
        const float step = 1.023e6F / 25.0e6F;
        // one untimed epoch opens the coalescer channel and uploads the code table; the clock starts when all threads are there
        mc.set_input_output_vectors(out, iq.data());
        mc.set_stream_position(3, base, n);
        if (!mc.Carrier_wipeoff_multicorrelator_resampler(0.0F, 0.001F, 0.0F, 0.0F, step, 0.0F, n)) failed++;
        ready++;
        while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
        for (int k = 0; k < epochs; k++)
            {
                const uint64_t pos = static_cast<uint64_t>(k + 1) * n;
                mc.set_input_output_vectors(out, iq.data() + pos);
                mc.set_stream_position(3, base + pos, n);
                const auto t0 = std::chrono::steady_clock::now();
                if (!mc.Carrier_wipeoff_multicorrelator_resampler(0.1F * t, 0.001F + 1e-5F * t, 0.0F, 0.25F * (t % 4), step, 0.0F, n)) failed++;
                const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
                lat_sum[t] += us;
                lat_max[t] = std::max(lat_max[t], us);
            }
        mc.free();
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) pool.emplace_back(worker, t);
    while (ready.load() < n_threads) std::this_thread::yield();
    co->reset_stats();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& t : pool) t.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const auto st = co->stats();
    double mean = 0.0, mx = 0.0;
    for (int t = 0; t < n_threads; t++)
        {
            mean += lat_sum[t] / epochs / n_threads;
            mx = std::max(mx, lat_max[t]);
        }
    CHECK(failed.load() == 0, "%d correlations failed: %s", failed.load(), co->last_error());
    std::printf("COALESCED {\"threads\": %d, \"epochs\": %d, \"window_us\": %d, \"msamples_per_s\": %.1f, \"call_latency_us_mean\": %.1f, "
                "\"call_latency_us_max\": %.1f, \"batches\": %llu, \"items_per_batch\": %.1f, \"batch_round_trip_us\": %.1f, "
                "\"window_expired\": %llu, \"copy_ratio\": %.3f}\n",
        n_threads, epochs, window_us, static_cast<double>(n_threads) * epochs * n / dt / 1e6, mean, mx, static_cast<unsigned long long>(st.batches),
        st.batches ? static_cast<double>(st.items) / st.batches : 0.0, st.batches ? st.sum_batch_us / st.batches : 0.0,
        static_cast<unsigned long long>(st.window_expired), st.samples_offered ? static_cast<double>(st.samples_copied) / st.samples_offered : 0.0);
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::string(argv[1]) == "--thresholds")
        {
            // CPU-only: print compute_threshold for a few configurations (checked against scipy by pytest)
            const float pfas[3] = {0.001F, 0.01F, 1e-6F};
            const uint32_t sizes[3] = {4000, 25000, 8000};
            const uint32_t bins[3] = {40, 81, 80};
            const uint32_t dw[3] = {1, 2, 8};
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++)
                    std::printf("THRESH %g %u %u %u %.9g\n", pfas[i], sizes[j], bins[j], dw[j], b200::compute_threshold(pfas[i], sizes[j], bins[j], dw[j]));
            return 0;
        }
    if (argc > 1 && std::string(argv[1]) == "--coalescer")
        {
            // throughput of the class interface through the coalescer: ./test_host_mirror --coalescer [threads epochs window_us]
            const int threads = argc > 2 ? std::atoi(argv[2]) : 32;
            const int epochs = argc > 3 ? std::atoi(argv[3]) : 200;
            const int window = argc > 4 ? std::atoi(argv[4]) : 200;
            coalescer_throughput(threads, epochs, window);
            return g_fail ? 1 : 0;
        }
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> ud(0.0F, 1.0F);
    std::vector<float> code(1023);
    for (auto& c : code) c = (rng() & 1U) ? 1.0F : -1.0F;
    std::vector<std::complex<float>> in(2 * 8192);
    for (auto& v : in) v = std::complex<float>(ud(rng), ud(rng));

    shift_pointer_semantics(code, in);
    const int sizes[3] = {2048, 4096, 8192};
    for (int n : sizes)
        {
            const int threads = 6;
            std::vector<std::complex<float>> out(3 * threads);
            std::vector<std::thread> pool;
            const auto t0 = std::chrono::steady_clock::now();
            for (int t = 0; t < threads; t++) pool.emplace_back(correlator_worker, t, n, 50, &code, &in, &out);
            for (auto& t : pool) t.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            std::printf("B200 multicorrelator (real codes): %d concurrent correlators, length=%d : %.3e [s] per call\n", threads, n, dt / 50);
            for (int t = 1; t < threads; t++)
                for (int k = 0; k < 3; k++) CHECK(out[3 * t + k] == out[k], "thread %d tap %d differs", t, k);
#ifdef HAVE_REF
            ref_select_arch("a_avx");
            void* h = ref_mc_create(8192, 3, 0);
            float shifts[3] = {-0.5F, 0.0F, 0.5F};
            ref_mc_set_code(h, code.data(), 1023, shifts);
            std::complex<float> want[3];
            ref_mc_correlate(h, reinterpret_cast<const float*>(in.data()), 0.0F, 0.1F, 0.0F, 0.4F, 0.3F, 0.0F, n, reinterpret_cast<float*>(want));
            ref_mc_destroy(h);
            for (int k = 0; k < 3; k++)
                {
                    const float rel = std::abs(out[k] - want[k]) / std::abs(want[k]);
                    CHECK(rel < 1e-3F, "n=%d tap %d: got (%g,%g) want (%g,%g) rel %g", n, k, out[k].real(), out[k].imag(), want[k].real(), want[k].imag(), rel);
                }
#endif
        }

    // ---- acquisition core: PRN-like code delayed by 524 samples and shifted by +1680 Hz -----------
    {
        b200::Acq_Conf_Core conf;
        conf.fs_in = 4000000;
        conf.samples_per_ms = 4000.0F;
        conf.samples_per_code = 4000.0F;
        conf.samples_per_chip = 3;
        conf.doppler_max = 5000;
        conf.doppler_step = 250;
        conf.pfa = 0.001F;
        b200::Pcps_Acquisition_Core acq(conf);
        CHECK(acq.ok(), "acquisition core not created");
        CHECK(acq.d_fft_size == 4000 && acq.d_num_doppler_bins == 40, "sizes");
        // closed form check of the threshold: P(2, th/2) = (1-pfa)^(1/nbins)
        const double x = acq.get_threshold() / 2.0;
        const double q = (1.0 + x) * std::exp(-x);
        CHECK(std::fabs(q / (-std::expm1(std::log1p(-0.001) / 160000.0)) - 1.0) < 1e-5, "threshold %g", acq.get_threshold());
        std::vector<std::complex<float>> sampled(4000);
        for (int i = 0; i < 4000; i++) sampled[i] = std::complex<float>(0.0F, code[static_cast<size_t>(i * 1023.0 / 4000.0)]);
        std::vector<std::complex<float>> sig(4000);
        std::normal_distribution<float> nd(0.0F, 1.0F);
        for (int i = 0; i < 4000; i++)
            {
                const double ph = 2.0 * M_PI * 1680.0 * i / 4e6;
                const std::complex<float> c = sampled[(i + 4000 - 524) % 4000];
                sig[i] = 0.2F * c * std::complex<float>(static_cast<float>(std::cos(ph)), static_cast<float>(std::sin(ph))) + std::complex<float>(nd(rng), nd(rng));
            }
        b200::Acq_Synchro syn;
        acq.set_gnss_synchro(&syn);
        acq.set_local_code(sampled.data());
        acq.init();
        acq.set_active(true);
        b200::AcquisitionResult res;
        const int ev = acq.acquisition_core(sig.data(), 123456, &res);
        std::printf("acquisition: event %d, delay %.0f samples, doppler %.0f Hz, stat %.1f (threshold %.1f)\n", ev, syn.Acq_delay_samples,
            syn.Acq_doppler_hz, res.test_statistics, acq.get_threshold());
        // the reference test's tolerances (gps_l1_ca_pcps_acquisition_test.cc:357-364)
        CHECK(ev == 1, "expected a positive acquisition");
        CHECK(std::fabs(syn.Acq_delay_samples - 524.0) <= 2.0, "delay %g", syn.Acq_delay_samples);
        CHECK(std::fabs(syn.Acq_doppler_hz - 1680.0) <= 666.0, "doppler %g", syn.Acq_doppler_hz);
        CHECK(syn.Acq_samplestamp_samples == 123456, "sample stamp");
        // noise only -> negative acquisition (event 2)
        for (auto& v : sig) v = std::complex<float>(nd(rng), nd(rng));
        acq.init();
        CHECK(acq.acquisition_core(sig.data(), 1, &res) == 2, "noise must not be acquired (stat %g)", res.test_statistics);
    }
    // ---- fine-Doppler acquisition block: same synthetic satellite, 10 ms of signal ----------------------
    {
        b200::Fine_Doppler_Conf conf;
        conf.fs_in = 4000000;
        conf.samples_per_ms = 4000.0F;
        conf.doppler_max = 1000;
        conf.doppler_step = 250;
        conf.max_dwells = 2;
        conf.threshold = 2.0F;
        b200::Pcps_Acquisition_Fine_Doppler_Core acq(conf);
        CHECK(acq.ok(), "fine-doppler core not created");
        CHECK(acq.d_fft_size == 4000 && acq.d_num_doppler_points == 8, "fine-doppler sizes");
        std::vector<std::complex<float>> sampled(4000);
        for (int i = 0; i < 4000; i++) sampled[i] = std::complex<float>(0.0F, code[static_cast<size_t>(i * 1023.0 / 4000.0)]);
        const int total = 4000 * 14;
        std::vector<std::complex<float>> sig(total);
        std::normal_distribution<float> nd(0.0F, 1.0F);
        for (int i = 0; i < total; i++)
            {
                const double ph = 2.0 * M_PI * 640.0 * i / 4e6;
                const std::complex<float> c = sampled[(i + 4000 - 1777) % 4000];
                sig[i] = 0.25F * c * std::complex<float>(static_cast<float>(std::cos(ph)), static_cast<float>(std::sin(ph))) + std::complex<float>(nd(rng), nd(rng));
            }
        b200::Acq_Synchro syn;
        acq.set_gnss_synchro(&syn);
        acq.set_local_code(sampled.data());
        acq.set_active(true);
        int pos = 0, ev = 0, calls = 0;
        while (ev == 0 && pos + 4000 <= total && calls < 64)
            {
                int consumed = 0;
                ev = acq.work(sig.data() + pos, 4000, &consumed);
                pos += consumed;
                calls++;
            }
        std::printf("fine-doppler acquisition: event %d after %d calls, delay %.0f samples, doppler %.2f Hz, stat %.1f, fft bin %u\n", ev, calls,
            syn.Acq_delay_samples, syn.Acq_doppler_hz, acq.test_statistics(), acq.fine_index());
        CHECK(ev == 1, "expected a positive fine-doppler acquisition");
        CHECK(std::fabs(syn.Acq_delay_samples - 1777.0) <= 1.0, "fine-doppler delay %g", syn.Acq_delay_samples);
        // 80 x 4000 bins over 4 MHz: 12.5 Hz per bin
        CHECK(std::fabs(syn.Acq_doppler_hz - 640.0) <= 12.5, "fine doppler %g", syn.Acq_doppler_hz);
        CHECK(syn.Acq_samplestamp_samples == 8000, "fine-doppler sample stamp %llu", static_cast<unsigned long long>(syn.Acq_samplestamp_samples));
    }
    // ---- free-running tracking loop: 0.4 s of a synthetic satellite pushed once, tracked on the device ---------
    {
        b200_engine* eng = b200::shared_engine();
        const int total = 4000 * 400;
        std::vector<std::complex<float>> sig(total);
        std::normal_distribution<float> nd(0.0F, 1.0F);
        const double doppler = -830.0, delay = 2222.0;
        const double rate = 1.023e6 * (1.0 + doppler / 1575.42e6);
        for (int i = 0; i < total; i++)
            {
                const double chip = std::fmod((i - delay) * rate / 4e6 + 1023.0 * 1000.0, 1023.0);
                const double ph = 2.0 * M_PI * doppler * i / 4e6;
                sig[i] = 0.3F * code[static_cast<size_t>(chip)] * std::complex<float>(static_cast<float>(std::cos(ph)), static_cast<float>(std::sin(ph))) +
                         std::complex<float>(nd(rng), nd(rng));
            }
        CHECK(b200_iq_create(eng, 5, total) == B200_OK, "iq_create");
        uint64_t first = 0;
        CHECK(b200_iq_push(eng, 5, reinterpret_cast<const b200_cf32*>(sig.data()), total, &first) == B200_OK && first == 0, "iq_push");
        b200::Dll_Pll_Conf_Core conf;
        conf.fs_in = 4e6;
        conf.early_late_space_chips = 0.5F;
        conf.pull_in_time_s = 1;
        b200::Signal_Core sg;
        sg.prn = 7;
        b200::B200_Dll_Pll_Veml_Loop loop;
        CHECK(loop.init(conf, sg, 5, code.data()), "loop init: %s", b200_last_error());
        CHECK(loop.start_tracking(delay, doppler + 30.0, 0, 0), "start_tracking");
        std::vector<b200_trk_dump_record> recs;
        CHECK(loop.run(390, &recs), "loop run: %s", b200_last_error());
        CHECK(recs.size() == 390, "records %zu", recs.size());
        b200_trk_loop_status st{};
        CHECK(loop.status(&st) && st.state == 2 && st.epochs == 390, "loop status state %d epochs %llu", st.state, static_cast<unsigned long long>(st.epochs));
        double mean_dopp = 0.0;
        for (size_t k = recs.size() - 100; k < recs.size(); k++) mean_dopp += recs[k].carrier_doppler_hz / 100.0;
        std::printf("device tracking loop: %zu epochs, doppler %.1f Hz (true %.1f), CN0 %.1f dB-Hz, next sample %llu\n", recs.size(), mean_dopp, doppler,
            st.CN0_SNV_dB_Hz, static_cast<unsigned long long>(st.sample_counter));
        CHECK(std::fabs(mean_dopp - doppler) < 3.0, "loop doppler %g", mean_dopp);
        CHECK(recs.back().abs_P > 1.5F * recs.back().abs_E, "prompt %g early %g", recs.back().abs_P, recs.back().abs_E);
    }
    std::printf(g_fail ? "HOST_MIRROR_FAILED (%d)\n" : "HOST_MIRROR_OK\n", g_fail);
    return g_fail ? 1 : 0;
}
