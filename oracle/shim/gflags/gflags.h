/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gflags/gflags.h>: flags are plain globals that
 * keep their default values (nothing parses a command line here). */
#pragma once
#include <cstdint>
#include <string>
#define DECLARE_string(name) extern std::string FLAGS_##name
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_int32(name) extern int32_t FLAGS_##name
#define DECLARE_uint32(name) extern uint32_t FLAGS_##name
#define DECLARE_int64(name) extern int64_t FLAGS_##name
#define DECLARE_uint64(name) extern uint64_t FLAGS_##name
#define DECLARE_double(name) extern double FLAGS_##name
#define DEFINE_string(name, val, txt) std::string FLAGS_##name = (val)
#define DEFINE_bool(name, val, txt) bool FLAGS_##name = (val)
#define DEFINE_int32(name, val, txt) int32_t FLAGS_##name = (val)
#define DEFINE_uint32(name, val, txt) uint32_t FLAGS_##name = (val)
#define DEFINE_int64(name, val, txt) int64_t FLAGS_##name = (val)
#define DEFINE_uint64(name, val, txt) uint64_t FLAGS_##name = (val)
#define DEFINE_double(name, val, txt) double FLAGS_##name = (val)
#define DEFINE_validator(name, fn) static const bool name##_validator_registered __attribute__((unused)) = ((void)(fn), true)
namespace gflags
{
}
namespace google
{
}
