/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <volk/volk_complex.h> (upstream VOLK is not under
 * /root/reference).  The complex typedefs are the same C++ types volk_gnsssdr_complex.h defines. */
#pragma once
#include <volk_gnsssdr/volk_gnsssdr_complex.h>
