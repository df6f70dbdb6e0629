/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gnuradio/basic_block.h>: names, ids and
 * message ports.  Messages are delivered synchronously to the subscribed handlers (GNU Radio queues them for
 * the subscriber's thread); every publication is also kept in a log the tests can read. */
#pragma once
#include <gnuradio/io_signature.h>
#include <gnuradio/thread/thread.h>
#include <pmt/pmt.h>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>
namespace gr
{
class basic_block;
typedef std::shared_ptr<basic_block> basic_block_sptr;

class basic_block : public std::enable_shared_from_this<basic_block>
{
public:
    typedef std::function<void(pmt::pmt_t)> msg_handler_t;
    virtual ~basic_block() = default;
    long unique_id() const { return d_unique_id; }
    std::string name() const { return d_name; }
    std::string alias() const { return d_name + std::to_string(d_unique_id); }
    io_signature::sptr input_signature() const { return d_input_signature; }
    io_signature::sptr output_signature() const { return d_output_signature; }

    void message_port_register_in(pmt::pmt_t port_id) { d_in_ports[pmt::symbol_to_string(port_id)]; }
    void message_port_register_out(pmt::pmt_t port_id) { d_out_ports[pmt::symbol_to_string(port_id)]; }
    template <typename T>
    void set_msg_handler(pmt::pmt_t which_port, T handler)
    {
        d_in_ports[pmt::symbol_to_string(which_port)] = msg_handler_t(handler);
    }
    void message_port_pub(pmt::pmt_t port_id, pmt::pmt_t msg)
    {
        const std::string port = pmt::symbol_to_string(port_id);
        std::vector<std::pair<basic_block*, std::string>> subs;
        {
            std::lock_guard<std::mutex> lk(d_msg_mu);
            d_published.emplace_back(port, msg);
            subs = d_out_ports[port];
        }
        for (auto& s : subs) s.first->shim_post(s.second, msg);
    }
    void message_port_sub(pmt::pmt_t port_id, basic_block* target, const std::string& target_port)
    {
        std::lock_guard<std::mutex> lk(d_msg_mu);
        d_out_ports[pmt::symbol_to_string(port_id)].emplace_back(target, target_port);
    }
    /* shim-only: deliver a message to an input port (what the scheduler does for a subscribed port) */
    void shim_post(const std::string& port, pmt::pmt_t msg)
    {
        auto it = d_in_ports.find(port);
        if (it != d_in_ports.end() && it->second) it->second(msg);
    }
    /* shim-only: everything this block has published so far, (port, message) in order */
    std::vector<std::pair<std::string, pmt::pmt_t>> shim_published()
    {
        std::lock_guard<std::mutex> lk(d_msg_mu);
        return d_published;
    }

protected:
    basic_block(const std::string& name, io_signature::sptr in, io_signature::sptr out)
        : d_name(name), d_input_signature(std::move(in)), d_output_signature(std::move(out)), d_unique_id(next_id()) {}
    std::string d_name;
    io_signature::sptr d_input_signature, d_output_signature;
    long d_unique_id;

private:
    static long next_id()
    {
        static long id = 0;
        return ++id;
    }
    std::mutex d_msg_mu;
    std::map<std::string, msg_handler_t> d_in_ports;
    std::map<std::string, std::vector<std::pair<basic_block*, std::string>>> d_out_ports;
    std::vector<std::pair<std::string, pmt::pmt_t>> d_published;
};
}  // namespace gr
