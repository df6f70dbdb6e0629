// Host-side engine state shared by engine.cu (tracking) and acq_engine.cu (acquisition).
#pragma once

#include "common.cuh"
#include "loop.cuh"

#include <mutex>
#include <vector>

namespace b200
{
constexpr int kMaxBands = 16;

struct Band
{
    float2* dev{nullptr};      // owned ring (nullptr when attached)
    const float2* base{nullptr};
    unsigned long long mask{0};
    unsigned long long first_index{0};
    unsigned long long capacity{0};
    unsigned long long write_index{0};  // absolute index of the next pushed sample
    unsigned long long valid_from{0};   // oldest absolute index whose sample is meaningful (after a gap push)
    bool attached{false};
    bool in_use{false};
    // device staging for integer sample pushes: two buffers, so that the next raw copy runs while the previous
    // block is being converted on the compute stream
    void* raw_stage[2]{nullptr, nullptr};
    unsigned long long raw_cap[2]{0, 0};  // bytes
    cudaEvent_t raw_free[2]{nullptr, nullptr};  // conversion that read the buffer has finished
    int raw_next{0};
};

struct Channel
{
    float* code_dev{nullptr};
    int code_cap{0};
    ChanDesc desc{};
};
}  // namespace b200

struct b200_engine
{
    int device{0};
    cudaStream_t stream{nullptr};
    bool own_stream{false};
    cudaStream_t copy_stream{nullptr};
    cudaEvent_t copy_done{nullptr};
    cudaEvent_t t0{nullptr}, t1{nullptr};
    std::mutex mu;
    std::mutex push_mu[b200::kMaxBands];  // serialises competing producers of one band (b200_iq_push_at)
    b200::Band bands[b200::kMaxBands];
    std::vector<b200::Channel> chans;
    // device mirrors
    b200::BandDesc* bands_dev{nullptr};
    b200::ChanDesc* chans_dev{nullptr};
    int chans_dev_cap{0};
    bool tables_dirty{true};
    int max_code_len{0};
    int taps_uniform{-1};
    unsigned int taps_mask{0};   // bit t set: some channel with a code table has t taps (mixed batches run one launch per tap count)
    bool any_high_dyn{false};
    int shared_mode{-1};   // -1 auto, 0 never, 1 always-when-legal (env B200_TRK_SHARED)
    // batch staging
    b200_trk_item* items_dev{nullptr};
    b200_trk_item* items_pin{nullptr};
    float2* out_dev{nullptr};
    float2* out_pin{nullptr};
    int batch_cap{0};
    int out_cap{0};
    float2* partial{nullptr};
    unsigned int* counters{nullptr};
    size_t partial_cap{0};
    int counters_cap{0};
    uint64_t launches{0};
    // asynchronous batches (b200_trk_submit / b200_trk_wait)
    struct Slot
    {
        b200_trk_item* items_pin{nullptr};
        b200_trk_item* items_dev{nullptr};
        float2* out_dev{nullptr};
        float2* out_pin{nullptr};
        int items_cap{0};
        int out_cap{0};
        int n_items{0};
        int out_stride{0};
        cudaEvent_t done{nullptr};
        cudaEvent_t items_ready{nullptr};  // work items have arrived (copy stream)
        bool busy{false};
        bool zero_copy{false};  // items read from / taps written to the pinned host buffers directly (small batches)
        uint64_t ticket{0};
        std::vector<int> perm;  // results k belong to the caller's item perm[k] (empty = identity)
    };
    static constexpr int kSlots = 16;
    Slot slots[kSlots];
    uint64_t next_ticket{1};
    // free-running DLL/PLL loops (loop_engine.cu)
    std::vector<b200::LoopDev> loops;  // host mirror: configuration and start-of-tracking state
    b200::LoopDev* loops_dev{nullptr};
    int loops_dev_cap{0};
    b200_trk_item* loop_items_dev{nullptr};
    float2* loop_taps_dev{nullptr};
    int* loop_nrec_dev{nullptr};
    unsigned int* loop_rec_dev{nullptr};
    size_t loop_rec_cap{0};  // 32-bit words
    int loop_mode{-1};       // 0 persistent kernel (default), 1 one launch pair per epoch with slices = 1, 2 same with automatic slices
};

namespace b200
{
// engine.cu
int upload_tables(b200_engine* e);
int ensure_partials(b200_engine* e, int n_items, int slices);
void loops_free(b200_engine* e);  // loop_engine.cu
}  // namespace b200

