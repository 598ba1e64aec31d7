// mm_fastx.hpp -- FASTA/FASTQ(.gz) record reader with the record semantics of klib's kseq_read as the
// reference uses it (src/common/kseq.h:171-208; winSketch.hpp:245-252, computeMap.hpp:125-134):
//   * a record starts at the next '>' or '@'; the name is everything up to the first white space;
//   * sequence = all isgraph() characters up to the next '>', '+' or '@' (wherever it occurs);
//   * after '+': skip that line, then read quality characters (33..127) until as many as bases were read
//     (plus the one character kseq's loop consumes when it notices the count is reached);
//   * return value: sequence length, -1 at end of file, -2 for a truncated quality string.
// Unlike kseq's character-at-a-time getc loop this reader works on whole buffer spans: lines are found with memchr,
// a line made of letters only (the normal case) is appended with one memcpy, quality lines are counted, not stored.
// `read_into` appends the bases straight to the caller's batch buffer.  Plain files bypass zlib.
#pragma once
#include <zlib.h>

#include <cctype>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace mmhost {

class FastxReader {
  gzFile fp_ = nullptr;
  FILE* raw_ = nullptr;              // plain (not gzip) input: read() straight into the buffer, no pass through zlib
  std::vector<unsigned char> buf_;
  size_t begin_ = 0, end_ = 0;
  bool eof_ = false;
  int last_char_ = 0;

  bool fill() {                      // true if at least one byte is available
    if (begin_ < end_) return true;
    if (eof_) return false;
    begin_ = 0;
    int n = raw_ ? (int)fread(buf_.data(), 1, buf_.size(), raw_) : gzread(fp_, buf_.data(), (unsigned)buf_.size());
    if (n <= 0) { eof_ = true; end_ = 0; return false; }
    end_ = (size_t)n;
    return true;
  }
  int getc_() { return fill() ? (int)buf_[begin_++] : -1; }
  // all bytes are letters / other graph characters above '@' (so none of '>' '+' '@', no white space, no control byte)
  static bool clean_span(const unsigned char* p, size_t n) {
    unsigned char ok = 1;
    for (size_t i = 0; i < n; i++) ok &= (unsigned char)((unsigned char)(p[i] - 65) < 62);
    return ok != 0;
  }

 public:
  std::string name, seq;             // seq is filled only by read(); read_into() appends to the caller's buffer instead

  explicit FastxReader(const std::string& path, size_t bufBytes = (size_t)4 << 20) : buf_(bufBytes) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return;
    unsigned char magic[2] = {0, 0};
    const size_t got = fread(magic, 1, 2, f);
    rewind(f);
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) { fp_ = gzdopen(fileno(f), "r"); if (fp_) gzbuffer(fp_, 1 << 20); }
    else { raw_ = f; setvbuf(raw_, nullptr, _IONBF, 0); }
  }
  ~FastxReader() { if (fp_) gzclose(fp_); if (raw_) fclose(raw_); }
  bool ok() const { return fp_ != nullptr || raw_ != nullptr; }

  long read() { seq.clear(); return read_into(seq); }

  // Appends the record's bases to `out` (std::string-like: size(), append(), push_back()).  Returns like kseq_read.
  template <class Buf>
  long read_into(Buf& out) {
    int c;
    if (last_char_ == 0) {           // jump to the next header
      for (;;) {
        if (!fill()) return -1;
        const unsigned char* p = buf_.data() + begin_; const unsigned char* e = buf_.data() + end_;
        while (p < e && *p != '>' && *p != '@') ++p;
        begin_ = (size_t)(p - buf_.data());
        if (p < e) { last_char_ = *p; begin_++; break; }
      }
    }
    name.clear();
    const size_t base = out.size();
    // name: up to the first white space
    c = -1;
    while (fill()) {
      const unsigned char* p = buf_.data() + begin_; const unsigned char* e = buf_.data() + end_; const unsigned char* s = p;
      while (p < e && !isspace(*p)) ++p;
      name.append((const char*)s, (size_t)(p - s));
      begin_ = (size_t)(p - buf_.data());
      if (p < e) { c = *p; begin_++; break; }
    }
    if (c == -1 && name.empty()) return -1;
    if (c != '\n' && c != -1) {      // rest of the header line (comment): skipped
      c = -1;
      while (fill()) {
        const unsigned char* p = buf_.data() + begin_;
        const void* nl = memchr(p, '\n', end_ - begin_);
        if (nl) { begin_ = (size_t)((const unsigned char*)nl - buf_.data()) + 1; c = '\n'; break; }
        begin_ = end_;
      }
    }
    // sequence: graph characters up to the next '>' '+' '@'
    c = -1;
    while (fill()) {
      const unsigned char* p = buf_.data() + begin_; const size_t avail = end_ - begin_;
      const void* nl = memchr(p, '\n', avail);
      const size_t len = nl ? (size_t)((const unsigned char*)nl - p) : avail;
      if (clean_span(p, len)) {
        out.append((const char*)p, len);
        begin_ += len + (nl ? 1 : 0);
        continue;
      }
      // a line with a marker or a non-letter: character by character, like kseq
      size_t i = 0; bool stop = false;
      const size_t lim = len + (nl ? 1 : 0);
      for (; i < lim; i++) {
        const unsigned char ch = p[i];
        if (ch == '>' || ch == '+' || ch == '@') { stop = true; break; }
        if (isgraph(ch)) out.push_back((char)ch);
      }
      if (stop) { c = p[i]; begin_ += i + 1; break; }
      begin_ += lim;
    }
    const size_t n = out.size() - base;
    if (c == '>' || c == '@') last_char_ = c;
    if (c != '+') { if (c == -1) last_char_ = 0; return (long)n; }
    // '+' line
    c = -1;
    while (fill()) {
      const unsigned char* p = buf_.data() + begin_;
      const void* nl = memchr(p, '\n', end_ - begin_);
      if (nl) { begin_ = (size_t)((const unsigned char*)nl - buf_.data()) + 1; c = '\n'; break; }
      begin_ = end_;
    }
    if (c == -1) return -2;
    // quality: count characters in 33..127 until n of them were seen, then kseq consumes one more character
    size_t q = 0;
    while (q < n && fill()) {
      const unsigned char* p = buf_.data() + begin_; const size_t avail = end_ - begin_;
      const void* nl = memchr(p, '\n', avail);
      const size_t len = nl ? (size_t)((const unsigned char*)nl - p) : avail;
      size_t valid = 0;
      for (size_t i = 0; i < len; i++) valid += (size_t)((unsigned char)(p[i] - 33) < 95);
      if (q + valid < n) { q += valid; begin_ += len + (nl ? 1 : 0); continue; }
      size_t i = 0;
      if (valid == len) { i = n - q; q = n; }                 // a line of quality characters only: the count is reached after n - q of them
      else for (; i < len && q < n; i++) q += (size_t)((unsigned char)(p[i] - 33) < 95);
      begin_ += i;
    }
    last_char_ = 0;
    if (q < n) return -2;
    (void)getc_();                   // the read kseq's loop does before it sees qual.l == seq.l
    return (long)n;
  }
};

}  // namespace mmhost
