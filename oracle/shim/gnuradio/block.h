/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gnuradio/block.h>: the general_work()
 * contract (forecast, consume, item counters, stream tags) without a scheduler.  gr::shim::Runner
 * (gnuradio/shim_runner.h) drives one block over an in-memory sample array the way the single-threaded
 * scheduler would: ask forecast(), call general_work(), advance by what was consumed / produced. */
#pragma once
#include <gnuradio/basic_block.h>
#include <gnuradio/tags.h>
#include <gnuradio/types.h>
#include <cstdint>
#include <stdexcept>
#include <vector>
namespace gr
{
class block : public basic_block
{
public:
    enum work_return_t
    {
        WORK_CALLED_PRODUCE = -2,
        WORK_DONE = -1
    };
    enum tag_propagation_policy_t
    {
        TPP_DONT = 0,
        TPP_ALL_TO_ALL = 1,
        TPP_ONE_TO_ONE = 2,
        TPP_CUSTOM = 3
    };
    ~block() override = default;

    virtual void forecast(int noutput_items, gr_vector_int& ninput_items_required)
    {
        for (auto& n : ninput_items_required) n = noutput_items + static_cast<int>(d_history) - 1;
    }
    virtual int general_work(int /*noutput_items*/, gr_vector_int& /*ninput_items*/, gr_vector_const_void_star& /*input_items*/,
        gr_vector_void_star& /*output_items*/)
    {
        throw std::runtime_error("gr::block::general_work() not implemented");
    }
    virtual bool start() { return true; }
    virtual bool stop() { return true; }

    void consume(int which_input, int how_many_items)
    {
        if (static_cast<size_t>(which_input) >= d_shim_consumed.size()) d_shim_consumed.resize(which_input + 1, 0);
        d_shim_consumed[which_input] += how_many_items;
    }
    void consume_each(int how_many_items)
    {
        if (d_shim_consumed.empty()) d_shim_consumed.resize(1, 0);
        for (auto& c : d_shim_consumed) c += how_many_items;
    }
    void produce(int which_output, int how_many_items)
    {
        if (static_cast<size_t>(which_output) >= d_shim_produced.size()) d_shim_produced.resize(which_output + 1, 0);
        d_shim_produced[which_output] += how_many_items;
    }
    uint64_t nitems_read(unsigned int which_input) const { return which_input < d_shim_nread.size() ? d_shim_nread[which_input] : 0; }
    uint64_t nitems_written(unsigned int which_output) const { return which_output < d_shim_nwritten.size() ? d_shim_nwritten[which_output] : 0; }

    void set_relative_rate(double r) { d_relative_rate = r; }
    void set_relative_rate(uint64_t interpolation, uint64_t decimation) { d_relative_rate = static_cast<double>(interpolation) / static_cast<double>(decimation); }
    double relative_rate() const { return d_relative_rate; }
    void set_max_noutput_items(int m) { d_max_noutput_items = m; }
    int max_noutput_items() const { return d_max_noutput_items; }
    void set_output_multiple(int m) { d_output_multiple = m; }
    int output_multiple() const { return d_output_multiple; }
    void set_alignment(int) {}
    void set_history(unsigned h) { d_history = h; }
    unsigned history() const { return d_history; }
    void set_tag_propagation_policy(tag_propagation_policy_t p) { d_tpp = p; }
    tag_propagation_policy_t tag_propagation_policy() const { return d_tpp; }

    void add_item_tag(unsigned int which_output, uint64_t abs_offset, const pmt::pmt_t& key, const pmt::pmt_t& value,
        const pmt::pmt_t& srcid = pmt::pmt_t())
    {
        tag_t t;
        t.offset = abs_offset;
        t.key = key;
        t.value = value;
        t.srcid = srcid;
        (void)which_output;
        d_shim_out_tags.push_back(t);
    }
    void get_tags_in_range(std::vector<tag_t>& v, unsigned int /*which_input*/, uint64_t abs_start, uint64_t abs_end)
    {
        v.clear();
        for (const auto& t : d_shim_in_tags)
            if (t.offset >= abs_start && t.offset < abs_end) v.push_back(t);
    }
    void get_tags_in_range(std::vector<tag_t>& v, unsigned int /*which_input*/, uint64_t abs_start, uint64_t abs_end, const pmt::pmt_t& key)
    {
        v.clear();
        for (const auto& t : d_shim_in_tags)
            if (t.offset >= abs_start && t.offset < abs_end && pmt::eqv(t.key, key)) v.push_back(t);
    }

    /* ---- shim-only scheduler hooks (gr::shim::Runner) ---- */
    std::vector<int> shim_take_consumed()
    {
        std::vector<int> c = d_shim_consumed;
        for (auto& x : d_shim_consumed) x = 0;
        return c;
    }
    void shim_advance(unsigned which_input, uint64_t n)
    {
        if (which_input >= d_shim_nread.size()) d_shim_nread.resize(which_input + 1, 0);
        d_shim_nread[which_input] += n;
    }
    void shim_advance_out(unsigned which_output, uint64_t n)
    {
        if (which_output >= d_shim_nwritten.size()) d_shim_nwritten.resize(which_output + 1, 0);
        d_shim_nwritten[which_output] += n;
    }
    void shim_add_input_tag(const tag_t& t) { d_shim_in_tags.push_back(t); }
    std::vector<tag_t>& shim_output_tags() { return d_shim_out_tags; }

protected:
    block(const std::string& name, io_signature::sptr in, io_signature::sptr out) : basic_block(name, std::move(in), std::move(out)) {}
    gr::thread::mutex d_setlock;

private:
    double d_relative_rate{1.0};
    int d_max_noutput_items{0};
    int d_output_multiple{1};
    unsigned d_history{1};
    tag_propagation_policy_t d_tpp{TPP_ALL_TO_ALL};
    std::vector<int> d_shim_consumed, d_shim_produced;
    std::vector<uint64_t> d_shim_nread, d_shim_nwritten;
    std::vector<tag_t> d_shim_in_tags, d_shim_out_tags;
};
typedef std::shared_ptr<block> block_sptr;
}  // namespace gr
