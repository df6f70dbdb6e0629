/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <pmt/pmt_sugar.h>. */
#pragma once
#include "pmt.h"
namespace pmt
{
inline pmt_t mp(const std::string& s) { return intern(s); }
inline pmt_t mp(const char* s) { return intern(s); }
inline pmt_t mp(long v) { return from_long(v); }
inline pmt_t mp(int v) { return from_long(v); }
}  // namespace pmt
