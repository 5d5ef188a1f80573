// gz_writer.hpp — gzip output stream of the niqki_b200 host.  The reference writes through
// zstr::ofstream (always gzip; every std::endl closes a gzip member, /root/reference/src/zstr.hpp:
// 337-348); byte parity is defined on the DECOMPRESSED stream, so one member per file
// is written here.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <stdexcept>
#include <string>

namespace nqh {

class GzWriter {
 public:
  GzWriter() = default;
  ~GzWriter() { close(); }
  GzWriter(const GzWriter&) = delete;
  GzWriter& operator=(const GzWriter&) = delete;
  void open(const std::string& path) {
    close();
    gz_ = gzopen(path.c_str(), "wb1");
    if (!gz_) throw std::runtime_error("cannot open output '" + path + "'");
    gzbuffer(gz_, 1 << 20);
  }
  bool is_open() const { return gz_ != nullptr; }
  void write(const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n) {
      const unsigned chunk = n > (1u << 30) ? (1u << 30) : (unsigned)n;
      if (gzwrite(gz_, c, chunk) != (int)chunk) throw std::runtime_error("gzwrite failed");
      c += chunk;
      n -= chunk;
    }
  }
  void write(const std::string& s) { write(s.data(), s.size()); }
  void put_u32(uint32_t v) { write(&v, 4); }
  void close() {
    if (gz_) {
      gzclose(gz_);
      gz_ = nullptr;
    }
  }

 private:
  gzFile gz_ = nullptr;
};

}  // namespace nqh
