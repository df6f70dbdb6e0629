/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gnuradio/top_block.h>: records the edges that
 * adapters connect, nothing runs.  msg_connect wires message ports (synchronous delivery). */
#pragma once
#include <gnuradio/basic_block.h>
#include <string>
#include <tuple>
#include <vector>
namespace gr
{
class top_block
{
public:
    explicit top_block(const std::string& name) : d_name(name) {}
    void connect(basic_block_sptr src, int src_port, basic_block_sptr dst, int dst_port) { d_edges.emplace_back(src, src_port, dst, dst_port); }
    void disconnect(basic_block_sptr src, int src_port, basic_block_sptr dst, int dst_port)
    {
        for (auto it = d_edges.begin(); it != d_edges.end(); ++it)
            if (std::get<0>(*it) == src && std::get<1>(*it) == src_port && std::get<2>(*it) == dst && std::get<3>(*it) == dst_port)
                {
                    d_edges.erase(it);
                    return;
                }
    }
    void msg_connect(basic_block_sptr src, pmt::pmt_t srcport, basic_block_sptr dst, pmt::pmt_t dstport)
    {
        src->message_port_sub(srcport, dst.get(), pmt::symbol_to_string(dstport));
        d_msg_keepalive.push_back(dst);
    }
    void msg_connect(basic_block_sptr src, const std::string& srcport, basic_block_sptr dst, const std::string& dstport)
    {
        msg_connect(src, pmt::intern(srcport), dst, pmt::intern(dstport));
    }
    size_t shim_edge_count() const { return d_edges.size(); }

private:
    std::string d_name;
    std::vector<std::tuple<basic_block_sptr, int, basic_block_sptr, int>> d_edges;
    std::vector<basic_block_sptr> d_msg_keepalive;
};
typedef std::shared_ptr<top_block> top_block_sptr;
inline top_block_sptr make_top_block(const std::string& name) { return std::make_shared<top_block>(name); }
}  // namespace gr
