/* TEST INFRASTRUCTURE ONLY (oracle shim; no upstream counterpart): drives ONE gr::block with one input
 * stream over an in-memory array, like GNU Radio's single-threaded scheduler would:
 *   available = items not yet consumed (capped at the buffer size a real flowgraph would offer);
 *   forecast(noutput, required); if available < required[0] -> starved, stop;
 *   general_work(noutput, ninput = {available}, in = {&samples[read]}, out) -> n produced;
 *   advance nitems_read by what the block consumed, nitems_written by n.
 * Output items (if the block has an output stream) are collected as raw bytes. */
#pragma once
#include <gnuradio/block.h>
#include <cstring>
#include <vector>
namespace gr
{
namespace shim
{
class Runner
{
public:
    Runner(block* b, size_t in_item_size, size_t out_item_size, int max_buffer_items = 1 << 20)
        : d_b(b), d_isz(in_item_size), d_osz(out_item_size), d_cap(max_buffer_items) {}

    /* run until the block is starved of input or `max_calls` work calls were made; returns calls made */
    long run(const void* samples, uint64_t n_items, long max_calls = -1)
    {
        const char* base = static_cast<const char*>(samples);
        long calls = 0;
        int idle = 0;
        std::vector<char> outbuf(d_osz ? d_osz * 64 : 1);
        while (max_calls < 0 || calls < max_calls)
            {
                const uint64_t rd = d_b->nitems_read(0) - d_origin;
                if (rd >= n_items) break;
                uint64_t avail = n_items - rd;
                if (avail > static_cast<uint64_t>(d_cap)) avail = d_cap;
                int noutput = d_b->max_noutput_items() > 0 ? d_b->max_noutput_items() : 1;
                gr_vector_int required(1, 0);
                d_b->forecast(noutput, required);
                if (static_cast<uint64_t>(required[0]) > avail) break;  // starved: a real source would deliver more later
                gr_vector_int ninput(1, static_cast<int>(avail));
                gr_vector_const_void_star in(1, base + rd * d_isz);
                gr_vector_void_star out;
                std::vector<void*> outptrs;
                for (int k = 0; k < 64 && d_osz; k++) outptrs.push_back(outbuf.data() + k * d_osz);
                /* the reference writes its single output item through `*out[0]` where out = (T**)&output_items[0] */
                out.assign(outptrs.begin(), outptrs.end());
                if (out.empty()) out.push_back(nullptr);
                const int produced = d_b->general_work(noutput, ninput, in, out);
                calls++;
                const std::vector<int> consumed = d_b->shim_take_consumed();
                const int c0 = consumed.empty() ? 0 : consumed[0];
                d_b->shim_advance(0, static_cast<uint64_t>(c0));
                if (produced > 0 && d_osz)
                    {
                        for (int k = 0; k < produced; k++)
                            d_out.insert(d_out.end(), outbuf.data() + k * d_osz, outbuf.data() + (k + 1) * d_osz);
                        d_b->shim_advance_out(0, static_cast<uint64_t>(produced));
                    }
                if (produced == block::WORK_DONE) break;
                if (c0 == 0 && produced <= 0)
                    {
                        if (++idle > d_idle_limit) break;  // neither consuming nor producing: would spin forever
                    }
                else
                    idle = 0;
            }
        return calls;
    }
    /* the next run() call treats samples[0] as absolute item `nitems_read(0)` at the time of this call */
    void rebase() { d_origin = d_b->nitems_read(0); }
    void set_idle_limit(int n) { d_idle_limit = n; }
    const std::vector<char>& output_bytes() const { return d_out; }
    size_t outputs() const { return d_osz ? d_out.size() / d_osz : 0; }
    void clear_outputs() { d_out.clear(); }

private:
    block* d_b;
    size_t d_isz, d_osz;
    int d_cap;
    uint64_t d_origin{0};
    int d_idle_limit{1000};
    std::vector<char> d_out;
};
}  // namespace shim
}  // namespace gr
