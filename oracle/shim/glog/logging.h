/* TEST INFRASTRUCTURE ONLY (oracle shim).  tracking_loop_filter.cc logs one warning through glog/abseil;
 * neither is installed.  LOG(x) swallows the stream. */
#pragma once
#include <iostream>
struct oracle_null_log
{
    template <typename T>
    oracle_null_log& operator<<(const T&) { return *this; }
};
#define LOG(severity) oracle_null_log()
#define DLOG(severity) oracle_null_log()
