/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gnuradio/thread/thread.h>
 * (GNU Radio: boost::thread / boost::mutex / boost::unique_lock typedefs). */
#pragma once
#include <condition_variable>
#include <mutex>
#include <thread>
namespace gr
{
namespace thread
{
typedef std::thread thread;
typedef std::mutex mutex;
typedef std::unique_lock<std::mutex> scoped_lock;
typedef std::condition_variable condition_variable;
}  // namespace thread
}  // namespace gr
