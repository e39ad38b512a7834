// Minimal stand-in for the cmake-generated gflags header (the reference vendors gflags as a CMake project whose public
// header is generated): just enough for openfst's flags.h and glog's logging.h to parse.  No flag parsing happens.
#pragma once
#include <stdint.h>
#include <string>
namespace google {
typedef int32_t int32; typedef uint32_t uint32; typedef int64_t int64; typedef uint64_t uint64;
inline void SetUsageMessage(const std::string&) {}
inline unsigned ParseCommandLineFlags(int*, char***, bool) { return 0; }
}
namespace gflags = google;
#define DECLARE_VARIABLE(type, shorttype, name, tn) namespace fL##shorttype { extern type FLAGS_##name; } using fL##shorttype::FLAGS_##name
#define DEFINE_VARIABLE_(type, shorttype, name, value) namespace fL##shorttype { type FLAGS_##name = value; } using fL##shorttype::FLAGS_##name
#define DECLARE_bool(name) DECLARE_VARIABLE(bool, B, name, bool)
#define DECLARE_int32(name) DECLARE_VARIABLE(::google::int32, I, name, int32)
#define DECLARE_uint32(name) DECLARE_VARIABLE(::google::uint32, U, name, uint32)
#define DECLARE_int64(name) DECLARE_VARIABLE(::google::int64, I64, name, int64)
#define DECLARE_uint64(name) DECLARE_VARIABLE(::google::uint64, U64, name, uint64)
#define DECLARE_double(name) DECLARE_VARIABLE(double, D, name, double)
#define DECLARE_string(name) namespace fLS { extern std::string& FLAGS_##name; } using fLS::FLAGS_##name
#define DEFINE_bool(name, v, d) DEFINE_VARIABLE_(bool, B, name, v)
#define DEFINE_int32(name, v, d) DEFINE_VARIABLE_(::google::int32, I, name, v)
#define DEFINE_uint32(name, v, d) DEFINE_VARIABLE_(::google::uint32, U, name, v)
#define DEFINE_int64(name, v, d) DEFINE_VARIABLE_(::google::int64, I64, name, v)
#define DEFINE_uint64(name, v, d) DEFINE_VARIABLE_(::google::uint64, U64, name, v)
#define DEFINE_double(name, v, d) DEFINE_VARIABLE_(double, D, name, v)
#define DEFINE_string(name, v, d) namespace fLS { std::string FLAGS_##name##_buf = v; std::string& FLAGS_##name = FLAGS_##name##_buf; } using fLS::FLAGS_##name
