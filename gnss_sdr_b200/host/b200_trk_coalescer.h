/*!
 * \file b200_trk_coalescer.h
 * \brief Per-process batching of the correlations that N tracking-block threads request concurrently.
 *
 * The reference runs one scheduler thread per tracking block and each calls its own correlator synchronously
 * once per code period (src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc:1232-1257).  On a GPU
 * that is one copy + launch + synchronisation per channel and epoch.  Here the block threads post their work item
 * (channel, absolute sample index, the seven NCO scalars) and sleep; one tick thread turns whatever was posted
 * within a short window into ONE b200_trk_submit / b200_trk_wait pair against the shared IQ band and wakes the
 * posters with their taps.  The samples themselves are offered by every block (each has the same stream in its
 * GNU Radio input buffer) and copied once (b200_iq_push_at).
 *
 * Ring safety: a block that runs ahead may not overwrite samples a slower block still needs; push() blocks
 * while (newest offered index - oldest active cursor) would exceed the band's capacity (a flowgraph's upstream
 * buffer bounds that spread anyway, file-driven tests without back-pressure need it).
 */
#ifndef B200_TRK_COALESCER_H
#define B200_TRK_COALESCER_H

#include "b200gnss.h"
#include <atomic>
#include <chrono>
#include <complex>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace b200
{
class Trk_Coalescer
{
public:
    struct Stats
    {
        uint64_t batches{0};
        uint64_t items{0};
        uint64_t window_expired{0};  // batches launched because the window ran out, not because everyone had posted
        double sum_batch_us{0.0};    // GPU round trip of the batches (submit -> taps on the host)
        double sum_latency_us{0.0};  // post -> result per item
        double max_latency_us{0.0};
        uint64_t samples_copied{0};
        uint64_t samples_offered{0};
    };

    //! process-wide instance over b200::shared_engine(); nullptr when no GPU is usable
    static Trk_Coalescer* instance();
    explicit Trk_Coalescer(b200_engine* engine);
    ~Trk_Coalescer();
    Trk_Coalescer(const Trk_Coalescer&) = delete;
    Trk_Coalescer& operator=(const Trk_Coalescer&) = delete;

    //! window the tick thread waits for stragglers once the first item of a batch is posted [us] (env B200_COALESCE_WINDOW_US, default 200)
    void set_window_us(int us) { d_window_us = us; }
    //! create the band store if needed (capacity in samples, rounded up to a power of two; env B200_BAND_CAPACITY, default 2^23)
    bool ensure_band(int band, uint64_t capacity_samples = 0);

    //! one registration per correlator object; returns the engine channel id or -1
    int open_channel(int band, int n_correlators);
    void close_channel(int id);
    bool set_code(int id, int code_length_chips, const float* code, const float* shifts_chips, bool high_dynamics);
    bool set_taps(int id, const float* shifts_chips);
    //! the channel stops taking part in batches until it posts again (loss of lock, stop_tracking)
    void idle(int id);

    //! offer [abs_index, abs_index + n) of the band's stream; copies what the band does not hold yet
    bool push(int id, uint64_t abs_index, const std::complex<float>* samples, uint64_t n);
    //! post one epoch; returns at once.  wait() blocks until its taps are in `out` (n_correlators values).
    bool post(int id, uint64_t abs_index, int n, float rem_carrier_phase_rad, float phase_step_rad, float phase_rate_step_rad,
        float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips);
    bool wait(int id, std::complex<float>* out);

    Stats stats();
    void reset_stats();
    const char* last_error() const { return d_error; }

private:
    enum State
    {
        FREE = 0,
        IDLE,
        POSTED,
        IN_FLIGHT,
        DONE,
        FAILED
    };
    struct Slot
    {
        int chan{-1};
        int band{0};
        int taps{0};
        State state{FREE};
        bool active{false};
        uint64_t cursor{0};  // first sample the channel still needs
        b200_trk_item item{};
        std::complex<float> out[B200_MAX_TAPS];
        std::chrono::steady_clock::time_point t_post;
    };
    void tick_loop();
    Slot* slot_of(int id);

    b200_engine* d_engine;
    std::mutex d_mu;
    std::condition_variable d_cv_tick, d_cv_done, d_cv_space;
    std::deque<Slot> d_slots;  // indexed by engine channel id; a deque keeps references valid when channels are added
    int d_posted{0};
    int d_last_batch{1};
    int d_window_us{200};
    bool d_stop{false};
    std::thread d_thread;
    Stats d_stats;
    uint64_t d_band_capacity[16] = {0};
    char d_error[256] = "";
};
}  // namespace b200
#endif
