// Test-infrastructure shim: boost::filesystem::path(...).extension() only.
#pragma once
#include <string>
namespace boost { namespace filesystem {
class path {
 public:
  path(const std::string &s) : s_(s) {}
  path(const char *s) : s_(s) {}
  std::string extension() const {
    const auto slash = s_.find_last_of('/');
    const auto dot = s_.find_last_of('.');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
    return s_.substr(dot);
  }
  const std::string &string() const { return s_; }
 private:
  std::string s_;
};
}}  // namespace boost::filesystem
