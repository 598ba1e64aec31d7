// mm_fastx.hpp -- FASTA/FASTQ(.gz) record reader with the record semantics of klib's kseq_read as the
// reference uses it (src/common/kseq.h:171-208; winSketch.hpp:245-252, computeMap.hpp:125-134):
//   * a record starts at the next '>' or '@'; the name is everything up to the first white space;
//   * sequence = all isgraph() characters up to the next '>', '+' or '@' (wherever it occurs);
//   * after '+': skip that line, then read quality characters (33..127) until as many as bases were read;
//   * return value: sequence length, -1 at end of file, -2 for a truncated quality string.
#pragma once
#include <zlib.h>

#include <cctype>
#include <cstdio>
#include <string>
#include <vector>

namespace mmhost {

class FastxReader {
  gzFile fp_ = nullptr;
  std::vector<unsigned char> buf_;
  int begin_ = 0, end_ = 0;
  bool eof_ = false;
  int last_char_ = 0;

  int getc_() {
    if (begin_ >= end_) {
      if (eof_) return -1;
      begin_ = 0;
      end_ = gzread(fp_, buf_.data(), (unsigned)buf_.size());
      if (end_ <= 0) { eof_ = true; end_ = 0; return -1; }
    }
    return (int)buf_[begin_++];
  }

 public:
  std::string name, comment, seq, qual;

  explicit FastxReader(const std::string& path) : buf_(1 << 20) {
    FILE* f = fopen(path.c_str(), "r");
    if (f) fp_ = gzdopen(fileno(f), "r");
  }
  ~FastxReader() { if (fp_) gzclose(fp_); }
  bool ok() const { return fp_ != nullptr; }

  long read() {
    int c;
    if (last_char_ == 0) {
      while ((c = getc_()) != -1 && c != '>' && c != '@') {}
      if (c == -1) return -1;
      last_char_ = c;
    }
    name.clear(); comment.clear(); seq.clear(); qual.clear();
    while ((c = getc_()) != -1 && !isspace(c)) name.push_back((char)c);
    if (c == -1 && name.empty()) return -1;
    if (c != '\n' && c != -1)
      while ((c = getc_()) != -1 && c != '\n') comment.push_back((char)c);
    while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@')
      if (isgraph(c)) seq.push_back((char)c);
    if (c == '>' || c == '@') last_char_ = c;
    if (c != '+') { if (c == -1) last_char_ = 0; return (long)seq.size(); }
    while ((c = getc_()) != -1 && c != '\n') {}
    if (c == -1) return -2;
    while ((c = getc_()) != -1 && qual.size() < seq.size())
      if (c >= 33 && c <= 127) qual.push_back((char)c);
    last_char_ = 0;
    if (qual.size() != seq.size()) return -2;
    return (long)seq.size();
  }
};

}  // namespace mmhost
