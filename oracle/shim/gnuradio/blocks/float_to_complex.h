/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gnuradio/blocks/float_to_complex.h>; only instantiated by
 * the reference's acquisition adapters for cbyte input (base_pcps_acquisition.cc:94-98), never run here. */
#pragma once
#include <gnuradio/sync_block.h>
#include <memory>
namespace gr
{
namespace blocks
{
class float_to_complex : public sync_block
{
public:
    typedef std::shared_ptr<float_to_complex> sptr;
    static sptr make(size_t vlen = 1) { return sptr(new float_to_complex(vlen)); }
    int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) override
    {
        const auto* re = static_cast<const float*>(input_items[0]);
        const auto* im = input_items.size() > 1 ? static_cast<const float*>(input_items[1]) : nullptr;
        auto* out = static_cast<gr_complex*>(output_items[0]);
        for (int i = 0; i < noutput_items; i++) out[i] = gr_complex(re[i], im ? im[i] : 0.0F);
        return noutput_items;
    }

private:
    explicit float_to_complex(size_t vlen) : sync_block("float_to_complex", io_signature::make(1, 2, sizeof(float) * vlen), io_signature::make(1, 1, sizeof(gr_complex) * vlen)) {}
};
}  // namespace blocks
}  // namespace gr
