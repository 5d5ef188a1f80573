// bio_reader.hpp — sequence-file ingest of the niqki_b200 host: zlib-backed line reader (plain,
// gzip and multi-member gzip alike) and the FASTA/FASTQ record reader.
//
// Behaviour follows the reference's readers (/root/reference/src/niqki_index.cpp:890-952):
//   get_data_type  (:944-952)  a path containing ".fq" or ".fastq" is FASTQ, anything else FASTA;
//   Biogetline 'Q' (:895-900)  four lines per record: header, sequence, two discarded;
//   Biogetline 'A' (:901-909)  one header line, then every line up to the next one starting with
//                              '>' (or EOF) concatenated — '\n' stripped, '\r' kept, as getline does;
//   records shorter than K come back empty (:910-913); callers keep a record only when its
//   length is > K (:395, :423, :450, :512).
// Nothing here touches the GPU.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace nqh {

// Buffered line reader over gzread(): transparent for uncompressed files.
class LineReader {
 public:
  explicit LineReader(const std::string& path) : buf_(1 << 20) {
    gz_ = gzopen(path.c_str(), "rb");
    if (!gz_) throw std::runtime_error("cannot open '" + path + "'");
    gzbuffer(gz_, 1 << 20);
  }
  ~LineReader() {
    if (gz_) gzclose(gz_);
  }
  LineReader(const LineReader&) = delete;
  LineReader& operator=(const LineReader&) = delete;

  // true once a read has hit end of file with nothing buffered (std::istream::eof() analogue for
  // the reference's `while(not in.eof())` loops)
  bool eof() {
    if (pos_ < len_) return false;
    fill();
    return len_ == 0;
  }
  // next byte without consuming it, -1 at EOF (istream::peek)
  int peek() {
    if (pos_ >= len_) fill();
    return pos_ < len_ ? (unsigned char)buf_[pos_] : -1;
  }
  // std::getline: appends the line WITHOUT its '\n' to `out` (cleared first unless append);
  // returns false when nothing at all could be read
  bool getline(std::string& out, bool append = false) {
    if (!append) out.clear();
    bool any = false;
    for (;;) {
      if (pos_ >= len_) {
        fill();
        if (len_ == 0) return any;
      }
      any = true;
      const char* s = buf_.data() + pos_;
      const char* nl = static_cast<const char*>(memchr(s, '\n', len_ - pos_));
      if (nl) {
        out.append(s, nl - s);
        pos_ += (nl - s) + 1;
        return true;
      }
      out.append(s, len_ - pos_);
      pos_ = len_;
    }
  }
  // same, appending straight to a byte vector (record bodies of large genomes)
  bool getline_into(std::vector<char>& out) {
    bool any = false;
    for (;;) {
      if (pos_ >= len_) {
        fill();
        if (len_ == 0) return any;
      }
      any = true;
      const char* s = buf_.data() + pos_;
      const char* nl = static_cast<const char*>(memchr(s, '\n', len_ - pos_));
      if (nl) {
        out.insert(out.end(), s, nl);
        pos_ += (nl - s) + 1;
        return true;
      }
      out.insert(out.end(), s, s + (len_ - pos_));
      pos_ = len_;
    }
  }
  bool skipline() {
    bool any = false;
    for (;;) {
      if (pos_ >= len_) {
        fill();
        if (len_ == 0) return any;
      }
      any = true;
      const char* s = buf_.data() + pos_;
      const char* nl = static_cast<const char*>(memchr(s, '\n', len_ - pos_));
      if (nl) {
        pos_ += (nl - s) + 1;
        return true;
      }
      pos_ = len_;
    }
  }

 private:
  void fill() {
    pos_ = 0;
    const int n = gzread(gz_, buf_.data(), (unsigned)buf_.size());
    if (n < 0) {
      int err = 0;
      const char* msg = gzerror(gz_, &err);
      throw std::runtime_error(std::string("read error: ") + (msg ? msg : "?"));
    }
    len_ = (size_t)n;
  }
  gzFile gz_ = nullptr;
  std::vector<char> buf_;
  size_t pos_ = 0, len_ = 0;
};

inline char data_type_of(const std::string& path) {  // :944-952
  if (path.find(".fq") != std::string::npos) return 'Q';
  if (path.find(".fastq") != std::string::npos) return 'Q';
  return 'A';
}

// One Biogetline call: appends the record's characters to `bases` and returns its length after
// the "< K comes back empty" rule (:910); `header` (nullable) receives the header line.
inline uint64_t read_record(LineReader& in, char type, uint32_t K, std::vector<char>& bases, std::string* header) {
  const size_t start = bases.size();
  std::string scratch;
  if (type == 'Q') {
    if (header) in.getline(*header); else in.skipline();
    in.getline_into(bases);
    in.skipline();
    in.skipline();
  } else {
    if (header) in.getline(*header); else in.skipline();
    for (int c = in.peek(); c != '>' && c != -1; c = in.peek()) in.getline_into(bases);
  }
  uint64_t len = bases.size() - start;
  if (len < K) {
    bases.resize(start);
    if (header) header->clear();
    len = 0;
  }
  return len;
}

}  // namespace nqh
