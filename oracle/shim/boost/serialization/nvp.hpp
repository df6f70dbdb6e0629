/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <boost/serialization/nvp.hpp>.  Gnss_Synchro's
 * serialize() template (gnss_synchro.h:200-240) is never instantiated here. */
#pragma once
#define BOOST_SERIALIZATION_NVP(name) name
