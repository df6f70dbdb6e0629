// Host logic of b200::Trk_Coalescer on a CPU-only box: the coalescer is linked against a FAKE engine (the ten C-ABI entry
// points it uses, implemented below on host memory) so that its threading, sample de-duplication, batching, idle handling and
// stream-restart logic run without a GPU.  The fake correlator returns, per tap t, (t + 1) * sum of the band's samples of
// the epoch: a result is right only if the band really holds the samples of THIS stream at the item's absolute indices.
#include "b200_trk_coalescer.h"
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

// ------------------------------------------------------------------------------------------------ fake engine
struct b200_engine
{
    std::mutex mu;
    struct Band
    {
        uint64_t cap{0}, lo{0}, hi{0};
        std::vector<std::complex<float>> ring;
    };
    std::map<int, Band> bands;
    struct Chan
    {
        int band, taps;
    };
    std::vector<Chan> chans;
    std::map<uint64_t, std::vector<b200_cf32>> results;
    std::map<uint64_t, int> strides;
    uint64_t next_ticket{1};
    std::atomic<uint64_t> submits{0}, forgets{0};
    std::atomic<int> range_errors{0};
};

static thread_local char g_err[128] = "";

extern "C"
{
    const char* b200_last_error(void) { return g_err; }

    int b200_iq_create(b200_engine* e, int band, uint64_t capacity)
    {
        std::lock_guard<std::mutex> lk(e->mu);
        auto& b = e->bands[band];
        b.cap = capacity;
        b.ring.assign(capacity, {0.f, 0.f});
        b.lo = b.hi = 0;
        return B200_OK;
    }

    int b200_iq_push_at(b200_engine* e, int band, uint64_t abs_index, const b200_cf32* host, uint64_t n, uint64_t* n_new)
    {
        std::lock_guard<std::mutex> lk(e->mu);
        auto& b = e->bands[band];
        if (n_new) *n_new = 0;
        if (n > b.cap) return B200_ERR_RANGE;
        if (abs_index >= b.lo && abs_index + n <= b.hi) return B200_OK;
        if (abs_index > b.hi || abs_index < b.lo) b.lo = b.hi = abs_index;
        const uint64_t skip = b.hi - abs_index;
        for (uint64_t i = skip; i < n; i++) b.ring[(abs_index + i) & (b.cap - 1)] = {host[i].re, host[i].im};
        b.hi = abs_index + n;
        if (b.hi - b.lo > b.cap) b.lo = b.hi - b.cap;
        if (n_new) *n_new = n - skip;
        return B200_OK;
    }

    int b200_iq_window(b200_engine* e, int band, uint64_t* lo, uint64_t* hi)
    {
        std::lock_guard<std::mutex> lk(e->mu);
        auto& b = e->bands[band];
        *lo = b.lo;
        *hi = b.hi;
        return B200_OK;
    }

    int b200_iq_forget(b200_engine* e, int band)
    {
        std::lock_guard<std::mutex> lk(e->mu);
        auto& b = e->bands[band];
        b.lo = b.hi;
        e->forgets++;
        return B200_OK;
    }

    int b200_trk_channel_create(b200_engine* e, int band, int n_correlators, int* id)
    {
        std::lock_guard<std::mutex> lk(e->mu);
        e->chans.push_back({band, n_correlators});
        *id = static_cast<int>(e->chans.size()) - 1;
        return B200_OK;
    }

    int b200_trk_channel_set_code(b200_engine*, int, int, const float*, const float*, int) { return B200_OK; }
    int b200_trk_channel_set_taps(b200_engine*, int, const float*) { return B200_OK; }

    int b200_trk_submit(b200_engine* e, const b200_trk_item* items, int n_items, int out_stride, uint64_t* ticket)
    {
        std::lock_guard<std::mutex> lk(e->mu);
        std::vector<b200_cf32> out(static_cast<size_t>(n_items) * out_stride, b200_cf32{0.f, 0.f});
        for (int i = 0; i < n_items; i++)
            {
                const auto& ch = e->chans[items[i].channel];
                auto& b = e->bands[ch.band];
                const uint64_t s0 = items[i].sample_index, s1 = s0 + static_cast<uint64_t>(items[i].n);
                if (s0 < b.lo || s1 > b.hi)
                    {
                        e->range_errors++;
                        std::snprintf(g_err, sizeof(g_err), "item %d outside the band window", i);
                        return B200_ERR_RANGE;
                    }
                double re = 0, im = 0;
                for (uint64_t k = s0; k < s1; k++)
                    {
                        re += b.ring[k & (b.cap - 1)].real();
                        im += b.ring[k & (b.cap - 1)].imag();
                    }
                for (int t = 0; t < ch.taps; t++) out[static_cast<size_t>(i) * out_stride + t] = {static_cast<float>((t + 1) * re), static_cast<float>((t + 1) * im)};
            }
        *ticket = e->next_ticket++;
        e->results[*ticket] = std::move(out);
        e->submits++;
        return B200_OK;
    }

    int b200_trk_wait(b200_engine* e, uint64_t ticket, b200_cf32* out)
    {
        std::lock_guard<std::mutex> lk(e->mu);
        auto it = e->results.find(ticket);
        if (it == e->results.end()) return B200_ERR_STATE;
        std::memcpy(out, it->second.data(), it->second.size() * sizeof(b200_cf32));
        e->results.erase(it);
        return B200_OK;
    }
}

namespace b200
{
b200_engine* shared_engine() { return nullptr; }  // instance() is not used here: the test owns its coalescer
}

// ------------------------------------------------------------------------------------------------ the scenarios
static int g_fail = 0;
#define CHECK(cond, ...)                        \
    do                                          \
        {                                       \
            if (!(cond))                        \
                {                               \
                    std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
                    std::printf(__VA_ARGS__);   \
                    std::printf("\n");          \
                    g_fail++;                   \
                }                               \
        }                                       \
    while (0)

// stream `tag`: sample i = (tag * 1e-3 + (i % 97), -(i % 31))
static std::complex<float> sample_of(int tag, uint64_t i) { return {static_cast<float>(tag) * 1e-3f + static_cast<float>(i % 97), -static_cast<float>(i % 31)}; }

static std::complex<double> expected(int tag, uint64_t s0, int n, int tap)
{
    double re = 0, im = 0;
    for (uint64_t k = s0; k < s0 + static_cast<uint64_t>(n); k++)
        {
            const auto v = sample_of(tag, k);
            re += v.real();
            im += v.imag();
        }
    return {(tap + 1) * re, (tap + 1) * im};
}

static std::atomic<int> g_behind{0};

// one block thread: `epochs` epochs of n samples starting at sample `start`, taps correlators
static void channel_thread(b200::Trk_Coalescer* co, int band, int taps, int tag, uint64_t start, int n, int epochs, std::atomic<int>* bad, int sleep_every = 0)
{
    const int id = co->open_channel(band, taps);
    if (id < 0)
        {
            (*bad)++;
            return;
        }
    float code[4] = {1, -1, 1, -1}, shifts[8] = {0};
    co->set_code(id, 4, code, shifts, false);
    std::vector<std::complex<float>> buf(static_cast<size_t>(n) + 64);
    for (int k = 0; k < epochs; k++)
        {
            const uint64_t s0 = start + static_cast<uint64_t>(k) * n;
            // a GNU Radio input buffer: the epoch plus some look-ahead, as ninput_items usually offers
            for (size_t i = 0; i < buf.size(); i++) buf[i] = sample_of(tag, s0 + i);
            bool behind = false;
            if (!co->push(id, s0, buf.data(), buf.size(), &behind))
                {
                    if (behind)
                        {
                            // more than a ring behind the others: the correlator class takes its synchronous path here
                            g_behind++;
                            continue;
                        }
                    (*bad)++;
                    break;
                }
            if (!co->post(id, s0, n, 0.f, 0.f, 0.f, 0.f, 0.25f, 0.f))
                {
                    (*bad)++;
                    break;
                }
            std::complex<float> out[8];
            if (!co->wait(id, out))
                {
                    (*bad)++;
                    break;
                }
            for (int t = 0; t < taps; t++)
                {
                    const auto want = expected(tag, s0, n, t);
                    if (std::abs(static_cast<double>(out[t].real()) - want.real()) > 1e-3 * std::abs(want.real()) + 1e-3 ||
                        std::abs(static_cast<double>(out[t].imag()) - want.imag()) > 1e-3 * std::abs(want.imag()) + 1e-3)
                        (*bad)++;
                }
            if (sleep_every && (k % sleep_every) == sleep_every - 1) std::this_thread::sleep_for(std::chrono::milliseconds(3));
        }
    co->close_channel(id);
}

int main()
{
    b200_engine eng;
    b200::Trk_Coalescer co(&eng);
    co.set_window_us(300);
    co.ensure_band(0, 1 << 16);
    co.ensure_band(1, 1 << 16);

    // 1. 24 block threads on one stream: right answers, one copy of the samples, shared batches
    {
        std::atomic<int> bad{0};
        std::vector<std::thread> th;
        const int threads = 24, n = 2000, epochs = 150;
        for (int c = 0; c < threads; c++) th.emplace_back(channel_thread, &co, 0, 3, 1, 1000 + static_cast<uint64_t>(c % 5), n, epochs, &bad, 0);
        for (auto& t : th) t.join();
        const auto st = co.stats();
        CHECK(bad == 0, "scenario 1: %d wrong or failed correlations", bad.load());
        // (a thread the scheduler kept off the CPU for longer than a ring lasts is refused and correlates on its own)
        CHECK(st.items + static_cast<uint64_t>(g_behind.load()) == static_cast<uint64_t>(threads) * epochs, "items %llu + behind %d",
            static_cast<unsigned long long>(st.items), g_behind.load());
        CHECK(g_behind.load() < threads * epochs / 20, "%d epochs fell out of the band", g_behind.load());
        CHECK(st.batches * 4 < st.items, "epochs do not share launches: %llu batches for %llu items", static_cast<unsigned long long>(st.batches),
            static_cast<unsigned long long>(st.items));
        const double stream = static_cast<double>(n) * epochs;
        CHECK(st.samples_copied < 1.3 * stream, "samples copied %.0f for a stream of %.0f", static_cast<double>(st.samples_copied), stream);
        CHECK(st.samples_offered > 20 * stream, "offered %.0f", static_cast<double>(st.samples_offered));
        CHECK(eng.range_errors == 0, "an item addressed samples outside the band window");
        std::printf("scenario 1: %llu items in %llu batches, copy ratio %.3f\n", static_cast<unsigned long long>(st.items),
            static_cast<unsigned long long>(st.batches), static_cast<double>(st.samples_copied) / static_cast<double>(st.samples_offered));
    }
    // 2. a NEW stream with different content whose indices overlap the old one's (flowgraph restarted in-process): the band must
    //    not answer "already there" with the old samples
    {
        co.reset_stats();
        const uint64_t forgets0 = eng.forgets;
        std::atomic<int> bad{0};
        std::vector<std::thread> th;
        for (int c = 0; c < 6; c++) th.emplace_back(channel_thread, &co, 0, 1, 2, 500, 2000, 40, &bad, 0);
        for (auto& t : th) t.join();
        CHECK(bad == 0, "scenario 2 (restart with overlapping indices): %d wrong or failed correlations", bad.load());
        CHECK(eng.forgets > forgets0, "the band was never declared stale");
    }
    // 3. two bands at once, different streams on them, one straggler thread that pauses every tenth epoch:
    //    everybody still gets the right answer and nobody waits for the straggler for ever
    {
        co.reset_stats();
        std::atomic<int> bad{0};
        std::vector<std::thread> th;
        for (int c = 0; c < 8; c++) th.emplace_back(channel_thread, &co, c % 2, 3, 4 + (c % 2), 0, 1000, 120, &bad, c == 7 ? 10 : 0);
        for (auto& t : th) t.join();
        CHECK(bad == 0, "scenario 3: %d wrong or failed correlations", bad.load());
        CHECK(eng.range_errors == 0, "an item addressed samples outside the band window");
    }
    // 4. ring smaller than the stream (wrap-around) with a fast and a slow channel: back-pressure keeps the slow one's samples alive
    {
        co.reset_stats();
        co.ensure_band(2, 1 << 14);   // 16384 samples: 8 epochs of 2000
        std::atomic<int> bad{0};
        std::thread slow(channel_thread, &co, 2, 3, 6, 0, 2000, 60, &bad, 4);
        std::thread fast(channel_thread, &co, 2, 3, 6, 0, 2000, 60, &bad, 0);
        slow.join();
        fast.join();
        CHECK(bad == 0, "scenario 4 (ring wrap with a slow channel): %d wrong or failed correlations", bad.load());
        CHECK(eng.range_errors == 0, "a channel's samples were overwritten before it used them");
        std::printf("epochs refused because the channel was more than a ring behind (all scenarios): %d\n", g_behind.load());
    }
    std::printf(g_fail ? "COALESCER_CPU FAILED (%d)\n" : "COALESCER_CPU OK\n", g_fail);
    return g_fail ? 1 : 0;
}
