// Test-infrastructure shim: filtering_istream / filtering_ostream as pass-through wrappers.
// gzip "filters" are no-ops, so files named *.gz written through this shim are plain text.
#pragma once
#include <istream>
#include <ostream>
namespace boost { namespace iostreams {
struct gzip_decompressor {};
struct gzip_compressor {};
class filtering_istream : public std::istream {
 public:
  filtering_istream() : std::istream(nullptr) {}
  void push(const gzip_decompressor &) {}
  void push(std::istream &is) { this->rdbuf(is.rdbuf()); this->clear(); }
};
class filtering_ostream : public std::ostream {
 public:
  filtering_ostream() : std::ostream(nullptr) {}
  ~filtering_ostream() { if (rdbuf()) flush(); }
  void push(const gzip_compressor &) {}
  void push(std::ostream &os) { this->rdbuf(os.rdbuf()); this->clear(); }
};
}}  // namespace boost::iostreams
