/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <glog/logging.h>; neither glog nor abseil is
 * installed.  LOG(x) / DLOG(x) / VLOG(n) swallow the stream. */
#pragma once
#include <iostream>
struct oracle_null_log
{
    template <typename T>
    oracle_null_log& operator<<(const T&) { return *this; }
    oracle_null_log& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
#define LOG(severity) oracle_null_log()
#define DLOG(severity) oracle_null_log()
#define VLOG(level) oracle_null_log()
#define DVLOG(level) oracle_null_log()
#define LOG_IF(severity, cond) oracle_null_log()
#define CHECK(cond) oracle_null_log()
namespace google
{
inline void InitGoogleLogging(const char*) {}
}  // namespace google
