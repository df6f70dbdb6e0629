/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gnuradio/sync_block.h>: 1:1 blocks implement work();
 * general_work() consumes what work() produced. */
#pragma once
#include <gnuradio/block.h>
namespace gr
{
class sync_block : public block
{
public:
    virtual int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) = 0;
    int general_work(int noutput_items, gr_vector_int& /*ninput_items*/, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) override
    {
        const int r = work(noutput_items, input_items, output_items);
        if (r > 0) consume_each(r);
        return r;
    }

protected:
    sync_block(const std::string& name, io_signature::sptr in, io_signature::sptr out) : block(name, std::move(in), std::move(out)) {}
};
}  // namespace gr
