/*!
 * \file b200_trk_coalescer.h
 * \brief Per-process batching of the correlations that N tracking-block threads request concurrently.
 *
 * The reference runs one scheduler thread per tracking block and each calls its own correlator synchronously
 * once per code period (src/algorithms/tracking/gnuradio_blocks/dll_pll_veml_tracking.cc:1232-1257).  On a GPU
 * that is one copy + launch + synchronisation per channel and epoch.  Here the block threads post their work item
 * (channel, absolute sample index, the seven NCO scalars) and wait; one tick thread turns whatever was posted
 * within a short window into ONE b200_trk_submit / b200_trk_wait pair against the shared IQ band and hands the
 * posters their taps.  The samples themselves are offered by every block (each has the same stream in its
 * GNU Radio input buffer) and copied once (b200_iq_push_at).
 *
 * Synchronisation is per slot (one atomic state word per channel: the poster spins briefly, then sleeps on it;
 * the tick thread never takes a lock on the hot path), because with 256 block threads a single mutex +
 * condition variable costs more than the GPU round trip.
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
#include <memory>
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
    static constexpr int kMaxChannels = 4096;
    static constexpr int kMaxBands = 16;

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

    //! offer [abs_index, abs_index + n) of the band's stream; copies what the band does not hold yet.  Returns false with
    //! *behind = true when the epoch starts before the oldest sample the band still holds while other channels are using the
    //! band (this channel fell more than a ring behind them): restarting the band there would pull the samples from under the
    //! others, so nothing is pushed and the caller correlates this epoch on its own (it holds the samples).
    bool push(int id, uint64_t abs_index, const std::complex<float>* samples, uint64_t n, bool* behind = nullptr);
    //! post one epoch; returns at once.  wait() blocks until its taps are in `out` (n_correlators values).
    bool post(int id, uint64_t abs_index, int n, float rem_carrier_phase_rad, float phase_step_rad, float phase_rate_step_rad,
        float rem_code_phase_chips, float code_phase_step_chips, float code_phase_rate_step_chips);
    bool wait(int id, std::complex<float>* out);

    Stats stats();
    void reset_stats();
    const char* last_error() const { return d_error; }

private:
    enum State : int
    {
        FREE = 0,
        IDLE,
        POSTED,
        IN_FLIGHT,
        DONE,
        FAILED
    };
    struct alignas(64) Slot  // one cache line pair per channel: no false sharing between block threads
    {
        std::atomic<int> state{FREE};
        std::atomic<bool> active{false};    // takes part in the batches right now (cleared by idle() and when it misses a window)
        std::atomic<bool> attached{false};  // uses the band's contents (cleared only by idle() / close_channel())
        std::atomic<uint64_t> cursor{0};  // first sample the channel still needs
        int band{0};
        int taps{0};
        b200_trk_item item{};
        std::complex<float> out[B200_MAX_TAPS];
        std::chrono::steady_clock::time_point t_post;
        double lat_sum_us{0.0}, lat_max_us{0.0};  // post -> result, written by the channel's own thread only
    };
    void tick_loop();
    void activate(Slot& s, bool stream_may_be_new);
    void deactivate(Slot& s);
    Slot* slot_of(int id);
    bool fits(int band, uint64_t abs_index, uint64_t n) const;

    b200_engine* d_engine;
    std::unique_ptr<Slot[]> d_slots;        // indexed by engine channel id; fixed size, never reallocated
    std::atomic<int> d_n_slots{0};          // highest id in use + 1
    std::atomic<int> d_posted{0};
    std::atomic<uint32_t> d_done_gen{0};    // batches completed: every waiter sleeps on this one word (one wake call per batch)
    std::atomic<int> d_n_active{0};
    std::atomic<int> d_band_active[kMaxBands];         // active channels per band
    std::atomic<int> d_band_attached[kMaxBands];       // channels per band between their first push and idle() / close
    std::atomic<bool> d_stop{false};
    std::atomic<uint64_t> d_band_written[kMaxBands];   // samples [lo, written) are in the band (fast path of push)
    std::atomic<uint64_t> d_band_lo[kMaxBands];
    uint64_t d_band_capacity[kMaxBands] = {0};
    std::mutex d_band_mu[kMaxBands];        // one producer at a time per band (slow path of push)
    std::condition_variable d_cv_space[kMaxBands];
    std::mutex d_admin_mu;                  // open/close/ensure_band/stats
    int d_window_us{200};
    std::thread d_thread;
    Stats d_stats;                          // tick-thread fields (latencies are kept per slot)
    std::atomic<uint64_t> d_samples_copied{0}, d_samples_offered{0};
    char d_error[256] = "";
};
}  // namespace b200
#endif
