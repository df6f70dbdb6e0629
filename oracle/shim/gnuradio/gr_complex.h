/* TEST INFRASTRUCTURE ONLY (oracle shim).  GNU Radio is not installed in this image; the reference's
 * tracking libs only need the gr_complex typedef from it (gnuradio/gr_complex.h: std::complex<float>). */
#pragma once
#include <complex>
typedef std::complex<float> gr_complex;
