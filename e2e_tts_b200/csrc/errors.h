// Per-thread last-error string behind e2e_last_error_string() (include/e2e_tts_b200.h).
#pragma once
#include <string>

namespace e2e {

inline std::string& last_error() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const std::string& msg) {
  last_error() = msg;
  return code;
}

}  // namespace e2e
