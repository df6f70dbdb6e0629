/*!
 * \file b200_trk_coalescer.cc
 * \brief see header.
 */
#include "b200_trk_coalescer.h"
#include "b200_multicorrelator_real_codes.h"  // b200::shared_engine
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace b200
{
Trk_Coalescer* Trk_Coalescer::instance()
{
    static std::mutex mu;
    static Trk_Coalescer* inst = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (inst == nullptr)
        {
            b200_engine* eng = shared_engine();
            if (eng == nullptr) return nullptr;
            inst = new Trk_Coalescer(eng);  // lives as long as the process (block destructors may still call it at exit)
        }
    return inst;
}


Trk_Coalescer::Trk_Coalescer(b200_engine* engine) : d_engine(engine)
{
    if (const char* env = std::getenv("B200_COALESCE_WINDOW_US")) d_window_us = std::atoi(env);
    d_thread = std::thread([this] { tick_loop(); });
}


Trk_Coalescer::~Trk_Coalescer()
{
    {
        std::lock_guard<std::mutex> lk(d_mu);
        d_stop = true;
    }
    d_cv_tick.notify_all();
    d_cv_done.notify_all();
    d_cv_space.notify_all();
    if (d_thread.joinable()) d_thread.join();
}


bool Trk_Coalescer::ensure_band(int band, uint64_t capacity_samples)
{
    if (band < 0 || band >= 16) return false;
    std::lock_guard<std::mutex> lk(d_mu);
    if (d_band_capacity[band] != 0) return true;
    if (capacity_samples == 0)
        {
            const char* env = std::getenv("B200_BAND_CAPACITY");
            capacity_samples = env ? std::strtoull(env, nullptr, 10) : (1ULL << 23);
        }
    uint64_t cap = 2;
    while (cap < capacity_samples) cap <<= 1;
    if (b200_iq_create(d_engine, band, cap) != B200_OK)
        {
            std::snprintf(d_error, sizeof(d_error), "%s", b200_last_error());
            return false;
        }
    d_band_capacity[band] = cap;
    return true;
}


Trk_Coalescer::Slot* Trk_Coalescer::slot_of(int id)
{
    if (id < 0 || id >= static_cast<int>(d_slots.size()) || d_slots[id].state == FREE) return nullptr;
    return &d_slots[id];
}


int Trk_Coalescer::open_channel(int band, int n_correlators)
{
    if (!ensure_band(band)) return -1;
    int id = -1;
    if (b200_trk_channel_create(d_engine, band, n_correlators, &id) != B200_OK)
        {
            std::snprintf(d_error, sizeof(d_error), "%s", b200_last_error());
            return -1;
        }
    std::lock_guard<std::mutex> lk(d_mu);
    if (id >= static_cast<int>(d_slots.size())) d_slots.resize(id + 1);
    Slot& s = d_slots[id];
    s = Slot();
    s.chan = id;
    s.band = band;
    s.taps = n_correlators;
    s.state = IDLE;
    return id;
}


void Trk_Coalescer::close_channel(int id)
{
    std::unique_lock<std::mutex> lk(d_mu);
    Slot* s = slot_of(id);
    if (s == nullptr) return;
    // an epoch still in flight finishes first (its result is dropped)
    d_cv_done.wait(lk, [&] { return d_stop || (s->state != POSTED && s->state != IN_FLIGHT); });
    s->active = false;
    s->state = IDLE;  // engine channel ids are never reused; the slot just goes quiet
    d_cv_space.notify_all();
}


bool Trk_Coalescer::set_code(int id, int code_length_chips, const float* code, const float* shifts_chips, bool high_dynamics)
{
    {
        std::lock_guard<std::mutex> lk(d_mu);
        if (slot_of(id) == nullptr) return false;
    }
    return b200_trk_channel_set_code(d_engine, id, code_length_chips, code, shifts_chips, high_dynamics ? 1 : 0) == B200_OK;
}


bool Trk_Coalescer::set_taps(int id, const float* shifts_chips)
{
    return b200_trk_channel_set_taps(d_engine, id, shifts_chips) == B200_OK;
}


void Trk_Coalescer::idle(int id)
{
    std::lock_guard<std::mutex> lk(d_mu);
    Slot* s = slot_of(id);
    if (s == nullptr) return;
    s->active = false;
    d_cv_space.notify_all();
    d_cv_tick.notify_all();
}


bool Trk_Coalescer::push(int id, uint64_t abs_index, const std::complex<float>* samples, uint64_t n)
{
    int band;
    {
        std::unique_lock<std::mutex> lk(d_mu);
        Slot* s = slot_of(id);
        if (s == nullptr) return false;
        band = s->band;
        s->cursor = abs_index;
        s->active = true;
        d_cv_space.notify_all();
        // back-pressure: never overwrite what a slower active channel of this band still needs
        const uint64_t cap = d_band_capacity[band];
        auto fits = [&] {
            uint64_t oldest = abs_index;
            for (const Slot& o : d_slots)
                if (o.state != FREE && o.active && o.band == band && o.cursor < oldest) oldest = o.cursor;
            return abs_index + n - oldest <= cap - cap / 8;
        };
        if (!fits())
            {
                // a channel that stopped calling (stalled test thread, block torn down without idle()) must not wedge the
                // rest: after the timeout the laggards are declared idle
                if (!d_cv_space.wait_for(lk, std::chrono::milliseconds(2000), [&] { return d_stop || fits(); }))
                    {
                        for (Slot& o : d_slots)
                            if (o.state != FREE && o.active && o.band == band && abs_index + n - o.cursor > cap - cap / 8) o.active = false;
                    }
            }
    }
    uint64_t n_new = 0;
    const int rc = b200_iq_push_at(d_engine, band, abs_index, reinterpret_cast<const b200_cf32*>(samples), n, &n_new);
    {
        std::lock_guard<std::mutex> lk(d_mu);
        d_stats.samples_offered += n;
        d_stats.samples_copied += n_new;
        if (rc != B200_OK) std::snprintf(d_error, sizeof(d_error), "%s", b200_last_error());
    }
    return rc == B200_OK;
}


bool Trk_Coalescer::post(int id, uint64_t abs_index, int n, float rem_carrier_phase_rad, float phase_step_rad, float phase_rate_step_rad,
    float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips)
{
    std::lock_guard<std::mutex> lk(d_mu);
    Slot* s = slot_of(id);
    if (s == nullptr || s->state == POSTED || s->state == IN_FLIGHT) return false;
    s->item.channel = id;
    s->item.n = n;
    s->item.sample_index = abs_index;
    s->item.rem_carrier_phase_rad = rem_carrier_phase_rad;
    s->item.phase_step_rad = phase_step_rad;
    s->item.phase_rate_step_rad = phase_rate_step_rad;
    s->item.rem_code_phase_chips = rem_code_phase_chips;
    s->item.code_phase_step_chips = code_phase_step_chips;
    s->item.code_phase_rate_step_chips = code_phase_rate_step_chips;
    s->state = POSTED;
    s->active = true;
    s->cursor = abs_index;
    s->t_post = std::chrono::steady_clock::now();
    d_posted++;
    d_cv_tick.notify_one();
    return true;
}


bool Trk_Coalescer::wait(int id, std::complex<float>* out)
{
    std::unique_lock<std::mutex> lk(d_mu);
    Slot* s = slot_of(id);
    if (s == nullptr) return false;
    d_cv_done.wait(lk, [&] { return d_stop || s->state == DONE || s->state == FAILED || s->state == IDLE; });
    const bool ok = s->state == DONE;
    if (ok)
        {
            for (int k = 0; k < s->taps; k++) out[k] = s->out[k];
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - s->t_post).count();
            d_stats.sum_latency_us += us;
            d_stats.max_latency_us = std::max(d_stats.max_latency_us, us);
        }
    if (s->state == DONE || s->state == FAILED) s->state = IDLE;
    return ok;
}


void Trk_Coalescer::tick_loop()
{
    std::vector<b200_trk_item> items;
    std::vector<int> ids;
    std::vector<b200_cf32> taps;
    std::unique_lock<std::mutex> lk(d_mu);
    while (!d_stop)
        {
            d_cv_tick.wait(lk, [&] { return d_stop || d_posted > 0; });
            if (d_stop) break;
            // everyone who took part in the previous batch and is still active is expected again; stragglers get
            // d_window_us from the moment the first item of this batch was seen
            const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(d_window_us);
            bool expired = false;
            for (;;)
                {
                    int expected = 0;
                    for (const Slot& s : d_slots)
                        if (s.state != FREE && s.active) expected++;
                    if (d_posted >= expected) break;
                    if (d_cv_tick.wait_until(lk, deadline) == std::cv_status::timeout)
                        {
                            expired = true;
                            break;
                        }
                    if (d_stop) return;
                }
            items.clear();
            ids.clear();
            for (Slot& s : d_slots)
                if (s.state == POSTED)
                    {
                        items.push_back(s.item);
                        ids.push_back(s.chan);
                        s.state = IN_FLIGHT;
                    }
            d_posted = 0;
            if (items.empty()) continue;
            if (expired)
                {
                    d_stats.window_expired++;
                    // who did not make it is not waited for next time (it posts -> it is active again)
                    for (Slot& s : d_slots)
                        if (s.state != FREE && s.active && s.state != IN_FLIGHT) s.active = false;
                }
            d_last_batch = static_cast<int>(items.size());
            lk.unlock();
            const auto t0 = std::chrono::steady_clock::now();
            taps.resize(items.size() * B200_MAX_TAPS);
            uint64_t ticket = 0;
            int rc = b200_trk_submit(d_engine, items.data(), static_cast<int>(items.size()), B200_MAX_TAPS, &ticket);
            if (rc == B200_OK) rc = b200_trk_wait(d_engine, ticket, taps.data());
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            lk.lock();
            if (rc != B200_OK) std::snprintf(d_error, sizeof(d_error), "%s", b200_last_error());
            for (size_t k = 0; k < ids.size(); k++)
                {
                    Slot& s = d_slots[ids[k]];
                    if (rc == B200_OK)
                        {
                            for (int t = 0; t < s.taps; t++) s.out[t] = std::complex<float>(taps[k * B200_MAX_TAPS + t].re, taps[k * B200_MAX_TAPS + t].im);
                            s.state = DONE;
                        }
                    else
                        s.state = FAILED;
                }
            d_stats.batches++;
            d_stats.items += items.size();
            d_stats.sum_batch_us += us;
            d_cv_done.notify_all();
        }
}


Trk_Coalescer::Stats Trk_Coalescer::stats()
{
    std::lock_guard<std::mutex> lk(d_mu);
    return d_stats;
}


void Trk_Coalescer::reset_stats()
{
    std::lock_guard<std::mutex> lk(d_mu);
    d_stats = Stats();
}
}  // namespace b200
