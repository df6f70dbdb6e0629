/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gnuradio/fxpt_nco.h> (a 32-bit fixed-point phase
 * accumulator with a table-driven sine in GNU Radio).  Only gnss_signal_replica.cc:28-38 (complex_exp_gen,
 * not on the acquisition / tracking paths built here) uses it; libm sincos stands in for the table, so those two
 * helpers are "parity unpinned". */
#pragma once
#include <gnuradio/types.h>
#include <cmath>
#include <cstdint>
namespace gr
{
class fxpt_nco
{
public:
    void set_freq(float angle_rate) { d_inc = static_cast<int32_t>(static_cast<int64_t>(std::llround(angle_rate * (4294967296.0 / (2.0 * M_PI))))); }
    void set_phase(float angle) { d_phase = static_cast<uint32_t>(static_cast<int64_t>(std::llround(angle * (4294967296.0 / (2.0 * M_PI))))); }
    void step() { d_phase += static_cast<uint32_t>(d_inc); }
    void sincos(gr_complex* output, int noutput_items, double ampl = 1.0)
    {
        for (int i = 0; i < noutput_items; i++)
            {
                const double a = static_cast<double>(static_cast<int32_t>(d_phase)) * (2.0 * M_PI / 4294967296.0);
                output[i] = gr_complex(static_cast<float>(std::cos(a) * ampl), static_cast<float>(std::sin(a) * ampl));
                step();
            }
    }

private:
    uint32_t d_phase{0};
    int32_t d_inc{0};
};
}  // namespace gr
