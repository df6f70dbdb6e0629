/* TEST INFRASTRUCTURE ONLY (oracle shim).  gr::fast_atan2f lives in GNU Radio (gnuradio-runtime
 * lib/math/fast_atan2f.cc, a 255-entry table lookup with ~1e-5 rad error), a third-party dependency that is
 * not under /root/reference and not installed here.  This stand-in uses atan2f, so the two functions of
 * tracking_discriminators.cc that call it (fll_four_quadrant_atan, pll_four_quadrant_atan) are
 * "parity unpinned" at the table's error level; everything else in the file is the reference's own code. */
#pragma once
#include <cmath>
namespace gr
{
static inline float fast_atan2f(float y, float x) { return std::atan2(y, x); }
}  // namespace gr
