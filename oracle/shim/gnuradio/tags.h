/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gnuradio/tags.h>. */
#pragma once
#include <pmt/pmt.h>
#include <cstdint>
namespace gr
{
struct tag_t
{
    uint64_t offset{0};
    pmt::pmt_t key;
    pmt::pmt_t value;
    pmt::pmt_t srcid;
};
}  // namespace gr
