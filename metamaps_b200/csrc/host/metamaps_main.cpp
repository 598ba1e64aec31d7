// metamaps_main.cpp -- C++ host of the B200 compute core: a `metamaps` command with the reference's
// sub-commands `mapDirectly` and `classify`, the same options and the same output files
// (reference src/map/mash_map.cpp:257-326, parseCmdArgs.hpp:33-117,255-460, mapWrap.h:34-213,407-441,
//  src/meta/fEM.h:466-1133).  All array work goes through the C ABI (include/metamaps_b200.h); this file only
// parses files, prints text with the reference's formatting, and keeps the taxonomy book-keeping.
//
// `index` / `mapAgainstIndex` keep the <prefix>.index manifest contract with a GPU-native index file (mm_index_save) instead
// of Boost archives.  Not supported: `classifyU` (disabled in the reference itself).
#include <sys/stat.h>

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_set>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/metamaps_b200.h"
#include "mm_fastx.hpp"

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// MM_HOST_TIMING=1: wall-clock split of the host stages on stderr (the bench's same-config leg records it)
struct PhaseTimer {
  const char* name; double t0; bool on;
  explicit PhaseTimer(const char* n) : name(n), t0(now_s()), on(getenv("MM_HOST_TIMING") != nullptr) {}
  ~PhaseTimer() { if (on) fprintf(stderr, "[host timing] %s %.3f s\n", name, now_s() - t0); }
};

[[noreturn]] void die(const std::string& m) { std::cerr << m << std::endl; exit(1); }
void ck(int rc, const char* what) { if (rc != 0) die(std::string(what) + ": " + mm_last_error()); }

std::vector<std::string> split(const std::string& s, const std::string& d) {     // util.h:80 semantics
  std::vector<std::string> out; size_t p = 0;
  for (;;) { size_t q = s.find(d, p); if (q == std::string::npos) { out.push_back(s.substr(p)); break; } out.push_back(s.substr(p, q - p)); p = q + d.size(); }
  return out;
}
void erase_nl(std::string& s) { while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); }
template <class T> std::string g6(T v) { std::ostringstream o; o << v; return o.str(); }     // default ostream formatting

// ------------------------------------------------------------------------------------------------ options
struct Options {
  std::map<std::string, std::string> kv; std::set<std::string> flags;
  bool has(const std::string& k) const { return kv.count(k) || flags.count(k); }
  std::string get(const std::string& k) const { return kv.at(k); }
};
// `--opt v`, `--opt=v`, `-o v`; long/short aliases folded to the long name (argvparser.hpp grammar)
Options parse_args(int argc, char** argv, int first, const std::map<std::string, std::string>& alias, const std::set<std::string>& valueless) {
  Options o;
  for (int i = first; i < argc; i++) {
    std::string a = argv[i];
    if (a.size() < 2 || a[0] != '-') die("Unknown argument: " + a);
    a = a.substr(a[1] == '-' ? 2 : 1);
    std::string val; bool hasVal = false;
    size_t eq = a.find('=');
    if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); hasVal = true; }
    auto it = alias.find(a);
    if (it == alias.end()) die("Unknown option: " + a);
    std::string key = it->second;
    if (valueless.count(key)) { o.flags.insert(key); continue; }
    if (!hasVal) { if (i + 1 >= argc) die("Option " + a + " requires a value"); val = argv[++i]; }
    o.kv[key] = val;
  }
  return o;
}

struct Params {          // skch::Parameters (map_parameters.hpp:57-76)
  int kmerSize = 16, windowSize = 0, minReadLength = 1000, alphabetSize = 4, threads = 1;
  uint64_t referenceSize = 0; float percentageIdentity = 80; double p_value = 1e-3;
  std::string ref, query, out, index; bool reportAll = false; uint64_t maximumMemory = 0;
};

uint64_t file_size(const std::string& f) {
  struct stat st; if (stat(f.c_str(), &st) != 0) die("Cannot open " + f + " for size determination.");
  return (uint64_t)st.st_size;
}

// ------------------------------------------------------------------------------------------------ mapDirectly
struct Contig { std::string name; int len; };

// mode: 0 = mapDirectly, 1 = index (no queries), 2 = mapAgainstIndex (no reference; parameters come from the index)
static Params parse_map_options(int argc, char** argv, int mode, int* device) {
  const std::map<std::string, std::string> alias = {
      {"reference", "reference"}, {"r", "reference"}, {"kmer", "kmer"}, {"k", "kmer"}, {"pval", "pval"}, {"p", "pval"},
      {"maxmemory", "maxmemory"}, {"mm", "maxmemory"}, {"window", "window"}, {"w", "window"}, {"minReadLen", "minReadLen"}, {"m", "minReadLen"},
      {"perc_identity", "perc_identity"}, {"pi", "perc_identity"}, {"threads", "threads"}, {"t", "threads"}, {"query", "query"}, {"q", "query"},
      {"all", "all"}, {"output", "output"}, {"o", "output"}, {"device", "device"}, {"index", "index"}, {"i", "index"}};
  Options o = parse_args(argc, argv, 2, alias, {"all"});
  Params P;
  if (mode != 0) {                                                       // parseCmdArgs.hpp:266-281
    if (!o.has("index")) die("Please provide index");
    P.index = o.get("index");
  }
  if (mode != 2) {
    if (!o.has("reference")) die("Provide reference file (s)");
    P.ref = o.get("reference");
    P.referenceSize = file_size(P.ref);                                   // commonFunc.hpp:211
    if (o.has("maxmemory")) P.maximumMemory = (uint64_t)(std::pow(1024, 3) * strtoull(o.get("maxmemory").c_str(), nullptr, 10));
    if (o.has("kmer")) P.kmerSize = atoi(o.get("kmer").c_str());
    if (o.has("pval")) P.p_value = atof(o.get("pval").c_str());
    if (o.has("minReadLen")) P.minReadLength = atoi(o.get("minReadLen").c_str());
    if (o.has("perc_identity")) P.percentageIdentity = (float)atof(o.get("perc_identity").c_str());
    if (o.has("window")) {                                                // parseCmdArgs.hpp:363-374
      P.windowSize = atoi(o.get("window").c_str());
      int s = P.minReadLength * 2 / P.windowSize;
      P.p_value = mm_stat_estimate_pvalue(s, P.kmerSize, P.alphabetSize, P.percentageIdentity, P.minReadLength, P.referenceSize);
    } else {
      P.windowSize = mm_stat_recommended_window(P.p_value, P.kmerSize, P.alphabetSize, P.percentageIdentity, P.minReadLength, P.referenceSize);
    }
  }
  if (mode != 1) {
    if (!o.has("query")) die("Provide query file (s)");
    P.query = o.get("query");
    P.reportAll = o.has("all");
    if (!o.has("output")) die("Provide output file");
    P.out = o.get("output");
  }
  if (o.has("threads")) P.threads = atoi(o.get("threads").c_str());
  *device = o.has("device") ? atoi(o.get("device").c_str()) : 0;
  return P;
}
static void print_params(const Params& P) {
  std::cout << "Parameters used:\n\t- alphabetSize: " << P.alphabetSize << "\n\t- kmerSize: " << P.kmerSize << "\n\t- minReadLength: " << P.minReadLength
            << "\n\t- p_value: " << P.p_value << "\n\t- percentageIdentity: " << P.percentageIdentity << "\n\t- windowSize: " << P.windowSize
            << "\n\t- maximumMemory: ~" << P.maximumMemory / std::pow(1024, 3) << " GB (GPU build: caps the device-memory budget of one index chunk)\n\n" << std::flush;
}
static void print_index_info(mm_index* idx) {
  int64_t nMin = 0, nUniq = 0; int32_t freq = 0, nCont = 0; int64_t bytes = 0;
  mm_index_stats(idx, &nMin, &nUniq, &freq, &nCont, &bytes);
  std::cout << "INFO, skch::Sketch::build, minimizers picked from reference = " << nMin << std::endl;
  if (freq != 0x7fffffff) std::cout << "INFO, skch::Sketch::computeFreqHist, With threshold 0.001%, ignore minimizers occurring >= " << freq << " times during lookup." << std::endl;
  else std::cout << "INFO, skch::Sketch::computeFreqHist, With threshold 0.001%, consider all minimizers during lookup." << std::endl;

}
// skch::Sketch over the reference FASTA (winSketch.hpp:180-365) on the device
// next chunk of the reference (contigs until `budgetBases` bases are in, like winSketch.hpp:284-329 cuts at a memory estimate);
// returns null when the FASTA is exhausted
// occurrence histogram + threshold carried from chunk to chunk, like the reference's one Sketch object does (winSketch.hpp:302-304,452-495)
struct FreqCarry { std::vector<int32_t> v; std::vector<int64_t> c; int32_t thr = 0x7fffffff; };
static mm_index* build_index_chunk(mm_ctx* ctx, const Params& P, mmhost::FastxReader& rd, uint64_t budgetBases, std::vector<Contig>& meta, FreqCarry& carry,
                                   size_t maxContigs = 0) {
  mm_index* idx = nullptr; ck(mm_index_create(ctx, P.kmerSize, P.windowSize, &idx), "mm_index_create");
  meta.clear();
  std::string buf; std::vector<int64_t> off{0}; uint64_t bases = 0;
  auto flush = [&]() {
    if (off.size() > 1) ck(mm_index_add(idx, buf.data(), off.data(), (int32_t)off.size() - 1), "mm_index_add");
    buf.clear(); off.assign(1, 0);
  };
  long len;
  while (bases < budgetBases && (maxContigs == 0 || meta.size() < maxContigs) && (len = rd.read_into(buf)) >= 0) {
    meta.push_back(Contig{rd.name, (int)len});
    off.push_back((int64_t)buf.size()); bases += (uint64_t)len;
    if (buf.size() >= ((size_t)256 << 20)) flush();
  }
  flush();
  if (meta.empty()) { mm_index_destroy(idx); return nullptr; }
  ck(mm_index_set_freq_carry(idx, carry.v.data(), carry.c.data(), (int32_t)carry.v.size(), carry.thr), "mm_index_set_freq_carry");
  ck(mm_index_finalize(idx), "mm_index_finalize");
  int32_t nb = 0; ck(mm_index_get_freq_hist(idx, nullptr, nullptr, 0, &nb, &carry.thr), "mm_index_get_freq_hist");
  carry.v.resize((size_t)nb); carry.c.resize((size_t)nb);
  ck(mm_index_get_freq_hist(idx, carry.v.data(), carry.c.data(), nb, &nb, &carry.thr), "mm_index_get_freq_hist");
  return idx;
}
static mm_index* build_reference_index(mm_ctx* ctx, const Params& P, std::vector<Contig>& meta) {
  PhaseTimer pt("reference: parse FASTA + index build");
  mm_index* idx = nullptr; ck(mm_index_create(ctx, P.kmerSize, P.windowSize, &idx), "mm_index_create");
  {
    mmhost::FastxReader rd(P.ref);
    if (!rd.ok()) die("Cannot open " + P.ref);
    std::string buf; std::vector<int64_t> off{0};
    auto flush = [&]() {
      if (off.size() > 1) ck(mm_index_add(idx, buf.data(), off.data(), (int32_t)off.size() - 1), "mm_index_add");
      buf.clear(); off.assign(1, 0);
    };
    long len;
    while ((len = rd.read_into(buf)) >= 0) {
      meta.push_back(Contig{rd.name, (int)len});
      off.push_back((int64_t)buf.size());
      if (buf.size() >= ((size_t)256 << 20)) flush();
    }
    flush();
    ck(mm_index_finalize(idx), "mm_index_finalize");
  }
  return idx;
}
static void write_meta_and_parameters(const std::string& prefix, const Params& P, const std::string& query, size_t total, size_t tooShort, size_t mapped, size_t notMapped) {
  std::ofstream m(prefix + ".meta");                                  // mapWrap.h:180-183
  m << "TotalReads " << total << "\nReadsTooShort " << tooShort << "\nReadsMapped " << mapped << "\nReadsNotMapped " << notMapped << "\n";
  m.close();
  std::ofstream ps(prefix + ".parameters");                          // mapWrap.h:196-211
  ps << "kmerSize " << P.kmerSize << "\nwindowSize " << P.windowSize << "\nminReadLength " << P.minReadLength << "\nalphabetSize " << P.alphabetSize
     << "\nreferenceSize " << P.referenceSize << "\npercentageIdentity " << P.percentageIdentity << "\np_value " << P.p_value
     << "\nrefSequences [" << P.ref << "]\nquerySequences [" << query << "]\noutFileName " << prefix << "\nreportAll " << (P.reportAll ? 1 : 0)
     << "\nindex \nmaximumMemory " << P.maximumMemory << "\n";
  std::cout << "INFO, skch::Map::mapQuery, [count of mapped reads, reads qualified for mapping, total input reads] = [" << mapped << ", "
            << total - tooShort << ", " << total << "]" << std::endl;
}
// ---- host pipeline: parser thread -> GPU thread -> formatter/writer thread, batches handed over in order ----------
// (the reference: kseq_read on the main thread, a pthread pool with an order-preserving output queue, ThreadPool.hpp:24-33,
//  computeMap.hpp:118-169; here the "pool" is the GPU and the order is the batch order)
template <class T>
class BoundedQueue {
  std::mutex m_; std::condition_variable cv_; std::deque<T> q_; size_t cap_; bool closed_ = false;
 public:
  explicit BoundedQueue(size_t cap) : cap_(cap) {}
  void push(T v) { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return q_.size() < cap_; }); q_.push_back(std::move(v)); cv_.notify_all(); }
  bool pop(T& out) {
    std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) return false;
    out = std::move(q_.front()); q_.pop_front(); cv_.notify_all(); return true;
  }
  void close() { std::lock_guard<std::mutex> l(m_); closed_ = true; cv_.notify_all(); }
};
struct ReadBatch { std::string buf; std::vector<int64_t> off{0}; std::vector<std::string> names; };
struct MappedBatch {            // what the GPU thread hands to the formatter
  std::unique_ptr<ReadBatch> in;
  std::vector<int32_t> read, seq, pos, shared, sketch, strand, mappedRead, status; std::vector<float> identity; std::vector<double> parsed, mapq;
  std::vector<int64_t> readOff;
};
// default-ostream formatting (6 significant digits, %g) without a stream: std::to_chars(general, 6) is specified as printf("%.6g")
inline void put_g6(std::string& o, double v) { char b[40]; auto r = std::to_chars(b, b + sizeof b, v, std::chars_format::general, 6); o.append(b, (size_t)(r.ptr - b)); }
inline void put_int(std::string& o, long long v) { char b[24]; auto r = std::to_chars(b, b + sizeof b, v); o.append(b, (size_t)(r.ptr - b)); }

// parser thread body: FASTA/FASTQ records -> batches below the byte / read caps.  Batch objects come back from the writer through
// `recycled` so that their (hundreds of MB of) text buffers are touched-in once, not mapped and faulted again for every batch.
static void parse_batches(const std::string& path, size_t capBytes, int64_t capReads, BoundedQueue<std::unique_ptr<ReadBatch>>& out,
                          BoundedQueue<std::unique_ptr<ReadBatch>>& recycled, int poolSize) {
  mmhost::FastxReader rd(path);
  if (!rd.ok()) die("Cannot open " + path);
  int made = 0;
  auto fresh = [&]() {
    std::unique_ptr<ReadBatch> b;
    if (made < poolSize) { made++; b = std::make_unique<ReadBatch>(); b->buf.reserve(capBytes + (capBytes >> 3)); }
    else if (!recycled.pop(b)) { b = std::make_unique<ReadBatch>(); }
    b->buf.clear(); b->off.assign(1, 0); b->names.clear();
    return b;
  };
  auto cur = fresh();
  for (;;) {
    const long len = rd.read_into(cur->buf);
    if (len < 0) break;
    cur->names.push_back(rd.name); cur->off.push_back((int64_t)cur->buf.size());
    if (cur->buf.size() >= capBytes || (int64_t)cur->names.size() >= capReads) { out.push(std::move(cur)); cur = fresh(); }
  }
  if (!cur->names.empty()) out.push(std::move(cur));
  out.close();
}

// skch::Map over every query file + unifyFiles + addMappingQualities (computeMap.hpp:104-172, mapWrap.h:34-323)
// chunk < 0: the whole reference is in `idx` -> final files.  chunk >= 0: `idx` is chunk N of the reference -> only the
// 12-column lines of this chunk go to <prefix>.<N> (skch::Map per chunk, mapWrap.h:425-432); unify_files finishes the job.
static void map_queries(mm_ctx* ctx, mm_index* idx, const std::vector<Contig>& meta, const Params& P, int chunk = -1) {
  std::vector<std::string> queries = split(P.query, ","), prefixes = split(P.out, ",");
  if (queries.size() != prefixes.size()) die("Please specify an equal number of input and output files (as comma-separated lists)");
  // batches stay below what the index's 64-bit hit key can number (many / chromosome-scale contigs leave fewer read bits)
  int64_t maxReads = 1 << 20; ck(mm_index_max_batch_reads(idx, &maxReads), "mm_index_max_batch_reads");
  if (maxReads > (1 << 20)) maxReads = 1 << 20;
  size_t capBytes = (size_t)256 << 20;
  if (const char* e = getenv("MM_HOST_BATCH_BYTES")) { long long v = atoll(e); if (v >= 1024) capBytes = (size_t)v; }     // tests: several batches on small inputs
  for (size_t fi = 0; fi < queries.size(); fi++) {
    PhaseTimer pt("queries: parse + map + format + write");
    const std::string prefix = chunk < 0 ? prefixes[fi] : prefixes[fi] + "." + std::to_string(chunk);
    FILE* out = fopen(prefix.c_str(), "wb");
    if (!out) die("Cannot open output file " + prefix);
    FILE* metaLengths = fopen(chunk < 0 ? (prefix + ".meta.unmappedReadsLengths").c_str() : "/dev/null", "wb");
    size_t total = 0, tooShort = 0, mapped = 0, notMapped = 0;
    std::unordered_set<std::string> seenIDs;
    BoundedQueue<std::unique_ptr<ReadBatch>> parsed(2), recycled(8);
    BoundedQueue<std::unique_ptr<MappedBatch>> done(2);
    std::thread parser([&] { parse_batches(queries[fi], capBytes, maxReads, parsed, recycled, /*pool*/6); });
    // formatter / writer: reportReadMappings' 12 columns (computeMap.hpp:546-588) + the two of addMappingQualities (mapWrap.h:313-320)
    std::thread writer([&] {
      std::unique_ptr<MappedBatch> mb;
      while (done.pop(mb)) {
        const ReadBatch& in = *mb->in; const int32_t n = (int32_t)in.names.size();
        const int64_t G = (int64_t)mb->mappedRead.size();
        for (int64_t g = 0; g < G; g++) if (mb->status[(size_t)g]) die("WARNING!\n\tlikelihood_sum: 0\n\treadID: " + in.names[(size_t)mb->mappedRead[(size_t)g]] + "\n========= END ==========");
        // counters, unmapped list, duplicate-id check: in read order
        std::string um; int64_t g = 0;
        for (int32_t r = 0; r < n; r++) {
          const int len = (int)(in.off[(size_t)r + 1] - in.off[(size_t)r]);
          total++;
          if (len < P.windowSize || len < P.kmerSize || len < P.minReadLength) { tooShort++; continue; }
          if (g < G && mb->mappedRead[(size_t)g] == r) {
            mapped++; g++;
            if (!seenIDs.insert(in.names[(size_t)r]).second) die("Seems that read ID " + in.names[(size_t)r] + " has already been processed - this target ID " + in.names[(size_t)r] + "\n");
          } else { notMapped++; put_int(um, len); um += '\t'; um += in.names[(size_t)r]; um += '\n'; }
        }
        if (!um.empty()) fwrite(um.data(), 1, um.size(), metaLengths);
        // the lines: mapped reads cut into one slice per thread, every slice formatted into its own string, written in order
        int nt = P.threads > 1 ? P.threads : 1; if (nt > 64) nt = 64;
        if ((int64_t)nt > G) nt = G > 0 ? (int)G : 1;
        std::vector<std::string> piece((size_t)nt);
#pragma omp parallel for num_threads(nt) schedule(static, 1)
        for (int t = 0; t < nt; t++) {
          std::string& o = piece[(size_t)t];
          const int64_t g0 = G * t / nt, g1 = G * (t + 1) / nt;
          if (g1 > g0) o.reserve((size_t)(mb->readOff[(size_t)g1] - mb->readOff[(size_t)g0]) * 128);
          for (int64_t gg = g0; gg < g1; gg++) {
            const int32_t r = mb->mappedRead[(size_t)gg]; const std::string& nm = in.names[(size_t)r];
            const int len = (int)(in.off[(size_t)r + 1] - in.off[(size_t)r]);
            for (int64_t m = mb->readOff[(size_t)gg]; m < mb->readOff[(size_t)gg + 1]; m++) {
              const Contig& cg = meta[(size_t)mb->seq[(size_t)m]]; const int ps = mb->pos[(size_t)m];
              o += nm; o += ' '; put_int(o, len); o += " 0 "; put_int(o, len - 1); o += (mb->strand[(size_t)m] > 0 ? " + " : " - "); o += cg.name; o += ' ';
              put_int(o, cg.len); o += ' '; put_int(o, ps); o += ' '; put_int(o, ps + len - 1); o += ' '; put_g6(o, (double)mb->identity[(size_t)m]); o += ' ';
              put_int(o, mb->shared[(size_t)m]); o += ' '; put_int(o, mb->sketch[(size_t)m]);
              if (chunk < 0) {                          // mapWrap.h:313-320: identity re-parsed from the printed text, corrected, then the quality
                const float corrected = (float)exp(-(1 - mb->parsed[(size_t)m] / 100.0));
                o += ' '; put_g6(o, (double)(corrected * 100)); o += ' '; put_g6(o, mb->mapq[(size_t)m]);
              }
              o += '\n';
            }
          }
        }
        for (const std::string& o : piece) if (!o.empty()) fwrite(o.data(), 1, o.size(), out);
        recycled.push(std::move(mb->in));                    // the text buffer goes back to the parser
      }
      recycled.close();
    });
    // GPU thread (this one): K0-K5 + identity + mapping quality per batch, results straight from the device-resident table
    std::unique_ptr<ReadBatch> rb;
    while (parsed.pop(rb)) {
      PhaseTimer pb("  batch: GPU calls");
      const int32_t n = (int32_t)rb->names.size();
      mm_map_params mp{P.percentageIdentity, P.minReadLength, P.reportAll ? 1 : 0, 0};
      mm_map_summary sum;
      ck(mm_map_batch(ctx, idx, rb->buf.data(), rb->off.data(), n, &mp, &sum), "mm_map_batch");
      ck(mm_classify_begin(ctx), "mm_classify_begin");
      int64_t M = 0; ck(mm_classify_add_mappings(ctx, 0, &M), "mm_classify_add_mappings");
      mm_classify_summary cs; ck(mm_classify_run(ctx, -1, &cs), "mm_classify_run");      // identity + mapping quality; the EM belongs to `classify`
      auto mb = std::make_unique<MappedBatch>();
      const size_t Mz = (size_t)cs.n_mappings, Gz = (size_t)cs.n_reads_mapped;
      mb->read.resize(Mz); mb->seq.resize(Mz); mb->pos.resize(Mz); mb->shared.resize(Mz); mb->sketch.resize(Mz); mb->strand.resize(Mz);
      mb->identity.resize(Mz); mb->parsed.resize(Mz); mb->mapq.resize(Mz); mb->mappedRead.resize(Gz); mb->status.resize(Gz); mb->readOff.resize(Gz + 1);
      ck(mm_classify_fetch(ctx, mb->read.data(), mb->seq.data(), mb->pos.data(), mb->shared.data(), mb->sketch.data(), mb->strand.data(), mb->identity.data(),
                           mb->parsed.data(), mb->mapq.data(), nullptr, nullptr, nullptr, (int64_t)Mz, mb->mappedRead.data(), mb->readOff.data(), nullptr,
                           mb->status.data(), nullptr, nullptr, 0), "mm_classify_fetch");
      mb->in = std::move(rb);
      done.push(std::move(mb));
    }
    done.close();
    parser.join(); writer.join();
    fclose(out); fclose(metaLengths);
    if (chunk < 0) write_meta_and_parameters(prefix, P, queries[fi], total, tooShort, mapped, notMapped);
  }
}

// mapWrap::unifyFiles (mapWrap.h:34-213): per read, in FASTQ order, the lines of every chunk file in chunk order; mapping
// qualities over the union (addMappingQualities, mapWrap.h:215-323); counters; the chunk files are removed.
static void unify_files(mm_ctx* ctx, const std::string& prefix, const Params& P, const std::vector<std::string>& chunkFiles, const std::string& query) {
  std::ofstream out(prefix);
  if (!out.is_open()) die("Cannot open output file " + prefix);
  std::ofstream metaLengths(prefix + ".meta.unmappedReadsLengths");
  std::vector<std::ifstream> files;
  for (const auto& f : chunkFiles) { files.emplace_back(f); if (!files.back().is_open()) die("Cannot open " + f); }
  std::vector<std::string> pending(files.size()); std::vector<bool> has(files.size(), false);
  auto peek = [&](size_t fI) -> const std::string* {              // next line of chunk file fI, or null at its end
    if (!has[fI]) { if (!std::getline(files[fI], pending[fI])) return nullptr; has[fI] = true; }
    return &pending[fI];
  };
  size_t total = 0, tooShort = 0, mapped = 0, notMapped = 0;
  std::set<std::string> processed;
  struct L { std::string text; double identity; int shared, sketch; };
  std::vector<L> lines; std::vector<int64_t> roff{0}; std::vector<int32_t> rlen; std::vector<std::string> rname;
  auto flush = [&]() {
    if (lines.empty()) { roff.assign(1, 0); rlen.clear(); rname.clear(); return; }
    std::vector<double> id(lines.size()), mq(lines.size()); std::vector<int32_t> sh(lines.size()), ss(lines.size()), st(rlen.size());
    for (size_t i = 0; i < lines.size(); i++) { id[i] = lines[i].identity; sh[i] = lines[i].shared; ss[i] = lines[i].sketch; }
    ck(mm_mapq_batch(ctx, id.data(), sh.data(), ss.data(), rlen.data(), roff.data(), (int64_t)rlen.size(), P.kmerSize, mq.data(), st.data()), "mm_mapq_batch");
    for (size_t r = 0; r < rlen.size(); r++) if (st[r]) die("WARNING!\n\tlikelihood_sum: 0\n\treadID: " + rname[r] + "\n========= END ==========");
    for (size_t i = 0; i < lines.size(); i++) { float corrected = exp(-(1 - id[i])); out << lines[i].text << " " << corrected * 100 << " " << mq[i] << "\n"; }
    lines.clear(); roff.assign(1, 0); rlen.clear(); rname.clear();
  };
  mmhost::FastxReader rd(query);
  if (!rd.ok()) die("Cannot open " + query);
  long len;
  while ((len = rd.read()) >= 0) {
    total++;
    if (len < P.windowSize || len < P.kmerSize || len < P.minReadLength) { tooShort++; continue; }
    const size_t before = lines.size();
    for (size_t fI = 0; fI < files.size(); fI++) {
      const std::string* l;
      while ((l = peek(fI)) != nullptr) {
        size_t sp = l->find(' ');
        if (sp == std::string::npos) { has[fI] = false; continue; }
        const std::string id = l->substr(0, sp);
        if (processed.count(id)) die("Seems that read ID " + id + " has already been processed - this target ID " + rd.name + "\n");
        if (id != rd.name) break;
        std::vector<std::string> f = split(*l, " ");
        if (f.size() != 12) die("Unexpected line in " + chunkFiles[fI] + ": " + *l);
        lines.push_back(L{*l, std::stod(f[9]) / 100.0, atoi(f[10].c_str()), atoi(f[11].c_str())});
        has[fI] = false;
      }
    }
    if (lines.size() == before) { notMapped++; metaLengths << (int)len << "\t" << rd.name << "\n"; }
    else { mapped++; roff.push_back((int64_t)lines.size()); rlen.push_back((int32_t)len); rname.push_back(rd.name); }
    processed.insert(rd.name);
    if (lines.size() >= (1u << 20)) flush();
  }
  flush();
  if (total - tooShort != 0)
    for (size_t fI = 0; fI < files.size(); fI++) if (peek(fI)) die("Error: output file " + std::to_string(fI) + " / " + std::to_string(files.size()) + " not completely processed.\n\tFile: " + chunkFiles[fI]);
  out.close(); metaLengths.close();
  for (auto& f : files) f.close();
  for (const auto& f : chunkFiles) std::remove(f.c_str());
  write_meta_and_parameters(prefix, P, query, total, tooShort, mapped, notMapped);
}

// Does the index fit the device?  ~110 bytes per minimizer while it is being built (8 + 8 + 2 resident, up to 64 of table at load
// 0.25, sort temporaries), 2/(w+1) minimizers per base
static uint64_t chunk_bases_for(double budgetBytes, const Params& P) { return (uint64_t)(budgetBytes / (110.0 * 2.0 / (P.windowSize + 1))); }
// MM_HOST_CHUNK_CONTIGS=n0,n1,...: chunk i takes exactly n_i contigs (tests: the chunk boundaries of a reference --maxmemory run,
// which cuts at its own host-memory estimate, winSketch.hpp:284-329); chunks beyond the list follow the byte budget
static std::vector<size_t> chunk_contig_counts() {
  std::vector<size_t> v;
  if (const char* e = getenv("MM_HOST_CHUNK_CONTIGS")) for (auto& x : split(e, ",")) if (!x.empty()) v.push_back((size_t)strtoull(x.c_str(), nullptr, 10));
  return v;
}
int run_mapDirectly(int argc, char** argv) {
  int device = 0;
  Params P = parse_map_options(argc, argv, 0, &device);
  print_params(P);
  mm_ctx* ctx = nullptr;
  { PhaseTimer pt("mm_ctx_create (CUDA context)"); ck(mm_ctx_create(device, &ctx), "mm_ctx_create"); }
  // Does the index fit the device?  ~75 bytes per minimizer while it is being built (8 + 8 + 2 resident, 32 of table at load
  // 0.5, sort temporaries), 2/(w+1) minimizers per base; --maxmemory (GB) caps the budget like in the reference.
  int64_t freeB = 0, totalB = 0; ck(mm_ctx_mem_info(ctx, &freeB, &totalB), "mm_ctx_mem_info");
  double budget = 0.6 * (double)freeB;
  if (P.maximumMemory > 0 && (double)P.maximumMemory < budget) budget = (double)P.maximumMemory;
  uint64_t chunkBases = chunk_bases_for(budget, P);
  if (const char* e = getenv("MM_HOST_CHUNK_BASES")) chunkBases = strtoull(e, nullptr, 10);       // tests: force chunking
  std::vector<Contig> meta;
  if (P.referenceSize <= chunkBases && !getenv("MM_HOST_CHUNK_CONTIGS")) {      // the usual case: one device-resident index
    mm_index* idx = build_reference_index(ctx, P, meta);
    print_index_info(idx);
    map_queries(ctx, idx, meta, P);
    mm_index_destroy(idx);
  } else {                                                           // the reference's chunk loop (winSketch.hpp:284-329, mapWrap.h:407-441)
    std::vector<std::string> queries = split(P.query, ","), prefixes = split(P.out, ",");
    if (queries.size() != prefixes.size()) die("Please specify an equal number of input and output files (as comma-separated lists)");
    mmhost::FastxReader rd(P.ref);
    if (!rd.ok()) die("Cannot open " + P.ref);
    std::vector<std::vector<std::string>> chunkFiles(prefixes.size());
    int N = 0; FreqCarry carry;
    const std::vector<size_t> cuts = chunk_contig_counts();
    for (mm_index* idx; (idx = build_index_chunk(ctx, P, rd, chunkBases, meta, carry, (size_t)N < cuts.size() ? cuts[(size_t)N] : 0)) != nullptr; N++) {
      std::cout << "Index chunk " << N << ": " << meta.size() << " contigs\n";
      print_index_info(idx);
      map_queries(ctx, idx, meta, P, N);
      for (size_t fi = 0; fi < prefixes.size(); fi++) chunkFiles[fi].push_back(prefixes[fi] + "." + std::to_string(N));
      mm_index_destroy(idx);
    }
    for (size_t fi = 0; fi < prefixes.size(); fi++) unify_files(ctx, prefixes[fi], P, chunkFiles[fi], queries[fi]);
  }
  mm_ctx_destroy(ctx);
  return 0;
}

// ------------------------------------------------------------------------------------------------ index / mapAgainstIndex
// mapWrap::createIndex (mapWrap.h:358-405): <prefix>.index (0 while building, then 1 + the chunk files), <prefix>.arguments
// (the parameters), <prefix>.0 (the index; here the GPU-native dump of mm_index_save plus <prefix>.0.contigs with the
// contig names and lengths).  The index is device-resident and written as ONE chunk.
int run_index(int argc, char** argv) {
  int device = 0;
  Params P = parse_map_options(argc, argv, 1, &device);
  print_params(P);
  { std::ofstream st(P.index + ".index"); if (!st.is_open()) die("Cannot open file " + P.index + ".index"); st << 0 << "\n"; }
  {
    std::ofstream a(P.index + ".arguments");
    if (!a.is_open()) die("Cannot open file " + P.index + ".arguments for serialization.");
    a.precision(17);
    a << "alphabetSize " << P.alphabetSize << "\nkmerSize " << P.kmerSize << "\nminReadLength " << P.minReadLength << "\np_value " << P.p_value
      << "\npercentageIdentity " << P.percentageIdentity << "\nwindowSize " << P.windowSize << "\nreferenceSize " << P.referenceSize
      << "\nrefSequences " << P.ref << "\n";
  }
  mm_ctx* ctx = nullptr; ck(mm_ctx_create(device, &ctx), "mm_ctx_create");
  int64_t freeB = 0, totalB = 0; ck(mm_ctx_mem_info(ctx, &freeB, &totalB), "mm_ctx_mem_info");
  double budget = 0.6 * (double)freeB;
  if (P.maximumMemory > 0 && (double)P.maximumMemory < budget) budget = (double)P.maximumMemory;
  uint64_t chunkBases = chunk_bases_for(budget, P);
  if (const char* e = getenv("MM_HOST_CHUNK_BASES")) chunkBases = strtoull(e, nullptr, 10);
  std::vector<Contig> meta; std::vector<std::string> written;
  auto store = [&](mm_index* idx, int N) {                       // <prefix>.N + its contig list (mapWrap.h:380-393)
    print_index_info(idx);
    const std::string chunk = P.index + "." + std::to_string(N);
    ck(mm_index_save(idx, chunk.c_str()), "mm_index_save");
    std::ofstream c(chunk + ".contigs");
    if (!c.is_open()) die("Cannot open file " + chunk + ".contigs for serialization.");
    for (const Contig& m : meta) c << m.len << "\t" << m.name << "\n";
    std::cout << "Stored state in file " << chunk << "\n" << std::flush;
    written.push_back(chunk);
  };
  if (P.referenceSize <= chunkBases && !getenv("MM_HOST_CHUNK_CONTIGS")) {
    mm_index* idx = build_reference_index(ctx, P, meta);
    store(idx, 0);
    mm_index_destroy(idx);
  } else {                                                       // one file per chunk of the reference, thresholds carried like mapDirectly's loop
    mmhost::FastxReader rd(P.ref);
    if (!rd.ok()) die("Cannot open " + P.ref);
    FreqCarry carry; const std::vector<size_t> cuts = chunk_contig_counts();
    int N = 0;
    for (mm_index* idx; (idx = build_index_chunk(ctx, P, rd, chunkBases, meta, carry, (size_t)N < cuts.size() ? cuts[(size_t)N] : 0)) != nullptr; N++) {
      store(idx, N);
      mm_index_destroy(idx);
    }
  }
  { std::ofstream st(P.index + ".index"); st << 1 << "\n"; for (auto& f : written) st << f << "\n"; }      // mapWrap.h:397-404
  std::cout << "\nIndex construction DONE, wrote " << written.size() << " files.\n\n" << std::flush;
  mm_ctx_destroy(ctx);
  return 0;
}
// mapWrap::mapAgainstIndex (mapWrap.h:443-554)
int run_mapAgainstIndex(int argc, char** argv) {
  int device = 0;
  Params P = parse_map_options(argc, argv, 2, &device);
  std::vector<std::string> lines;
  {
    std::ifstream st(P.index + ".index");
    std::string l;
    while (st.is_open() && std::getline(st, l)) { erase_nl(l); if (!l.empty()) lines.push_back(l); }
  }
  if (lines.empty() || lines[0] != "1") die("The file " + P.index + ".index does not indicate that index " + P.index + " was built successfully, abort.");
  if (lines.size() < 2) die("Index " + P.index + " was built successfully, but no index files present?");
  {
    std::ifstream a(P.index + ".arguments");
    if (!a.is_open()) die("Expected file " + P.index + ".arguments not found - have you supplied a valid index?");
    std::string key, val;
    while (a >> key >> val) {
      if (key == "alphabetSize") P.alphabetSize = atoi(val.c_str());
      else if (key == "kmerSize") P.kmerSize = atoi(val.c_str());
      else if (key == "minReadLength") P.minReadLength = atoi(val.c_str());
      else if (key == "p_value") P.p_value = atof(val.c_str());
      else if (key == "percentageIdentity") P.percentageIdentity = (float)atof(val.c_str());
      else if (key == "windowSize") P.windowSize = atoi(val.c_str());
      else if (key == "referenceSize") P.referenceSize = strtoull(val.c_str(), nullptr, 10);
      else if (key == "refSequences") P.ref = val;
    }
  }
  std::cout << "Parameters restored from index " << P.index << ".arguments\n\t- alphabetSize: " << P.alphabetSize << "\n\t- kmerSize: " << P.kmerSize
            << "\n\t- minReadLength: " << P.minReadLength << "\n\t- p_value: " << P.p_value << "\n\t- percentageIdentity: " << P.percentageIdentity
            << "\n\t- windowSize: " << P.windowSize << "\n\n" << std::flush;
  mm_ctx* ctx = nullptr; ck(mm_ctx_create(device, &ctx), "mm_ctx_create");
  const size_t nChunks = lines.size() - 1;
  std::vector<std::string> queries = split(P.query, ","), prefixes = split(P.out, ",");
  if (queries.size() != prefixes.size()) die("Please specify an equal number of input and output files (as comma-separated lists)");
  std::vector<std::vector<std::string>> chunkFiles(prefixes.size());
  for (size_t ci = 0; ci < nChunks; ci++) {                      // mapWrap.h:500-540: every chunk file in turn, then unifyFiles
    const std::string& file = lines[ci + 1];
    mm_index* idx = nullptr;
    ck(mm_index_load(ctx, file.c_str(), &idx), "mm_index_load");
    int32_t k = 0, w = 0; mm_index_params(idx, &k, &w, nullptr);
    if (k != P.kmerSize || w != P.windowSize) die("Index file " + file + " does not match " + P.index + ".arguments");
    std::vector<Contig> meta;
    {
      std::ifstream c(file + ".contigs");
      if (!c.is_open()) die("Cannot open file " + file + ".contigs for reading -- invalid index " + P.index);
      std::string l;
      while (std::getline(c, l)) { erase_nl(l); size_t t = l.find('\t'); if (t == std::string::npos) continue; meta.push_back(Contig{l.substr(t + 1), atoi(l.substr(0, t).c_str())}); }
    }
    int32_t nCont = 0; mm_index_stats(idx, nullptr, nullptr, nullptr, &nCont, nullptr);
    if ((size_t)nCont != meta.size()) die("Index file " + file + " and its .contigs file disagree");
    print_index_info(idx);
    map_queries(ctx, idx, meta, P, nChunks > 1 ? (int)ci : -1);
    if (nChunks > 1) for (size_t fi = 0; fi < prefixes.size(); fi++) chunkFiles[fi].push_back(prefixes[fi] + "." + std::to_string(ci));
    mm_index_destroy(idx);
  }
  if (nChunks > 1) for (size_t fi = 0; fi < prefixes.size(); fi++) unify_files(ctx, prefixes[fi], P, chunkFiles[fi], queries[fi]);
  mm_ctx_destroy(ctx);
  return 0;
}

// ------------------------------------------------------------------------------------------------ classify
struct TaxNode { std::string parent, rank, name; };
struct Taxonomy {                                                       // taxonomy.h:137-246
  std::map<std::string, TaxNode> T;
  static std::vector<std::string> dmp_fields(std::string line) {
    std::vector<std::string> f = split(line, "|");
    for (auto& x : f) { size_t a = x.find_first_not_of(" \t"); size_t b = x.find_last_not_of(" \t"); x = (a == std::string::npos) ? "" : x.substr(a, b - a + 1); }
    return f;
  }
  explicit Taxonomy(const std::string& dir) {
    std::map<std::string, std::string> names;
    std::ifstream n(dir + "/names.dmp");
    if (!n.is_open()) die("Cannot open file " + dir + "/names.dmp -- is '" + dir + "' a valid NCBI taxonomy?");
    std::string line;
    while (std::getline(n, line)) { erase_nl(line); if (line.empty()) continue; auto f = dmp_fields(line); if (f.size() > 3 && f[3] == "scientific name") names[f[0]] = f[1]; }
    std::ifstream d(dir + "/nodes.dmp");
    if (!d.is_open()) die("Cannot open file " + dir + "/nodes.dmp -- is '" + dir + "' a valid NCBI taxonomy?");
    while (std::getline(d, line)) {
      erase_nl(line); if (line.empty()) continue; auto f = dmp_fields(line);
      if (!names.count(f[0])) die("No name for taxon ID " + f[0] + " in taxonomy directory " + dir);
      T[f[0]] = TaxNode{f[1], f[2], names[f[0]]};
    }
    std::cout << "Read taxonomy from " << dir << " -- have " << T.size() << " nodes." << std::endl;
  }
  std::vector<std::string> upward(std::string id) const { std::vector<std::string> u{id}; while (id != "1") { id = T.at(id).parent; u.push_back(id); } return u; }
  std::map<std::string, std::string> upward_by_ranks(const std::string& id, const std::set<std::string>& ranks) const {      // taxonomy.h:76-110
    std::map<std::string, std::string> r;
    for (auto& n : upward(id)) {
      const std::string& rank = T.at(n).rank;
      if (!ranks.empty() && !ranks.count(rank)) continue;
      if (rank != "no rank") { if (r.count(rank)) die("Node " + id + " has multiple entries for rank " + rank); r[rank] = n; }
    }
    for (auto& t : ranks) if (!r.count(t)) r[t] = "Undefined";
    return r;
  }
  std::string first_non_x(std::string id) const { while (id.find("x") != std::string::npos) id = T.at(id).parent; return id; }
};

std::string extract_taxon(const std::string& contig) {                  // fEM.h:1396-1414, regex kraken:taxid\|(x?\d+)
  size_t p = 0; const std::string tag = "kraken:taxid|";
  while ((p = contig.find(tag, p)) != std::string::npos) {
    size_t q = p + tag.size(), s = q;
    if (q < contig.size() && contig[q] == 'x') q++;
    size_t dstart = q;
    while (q < contig.size() && isdigit((unsigned char)contig[q])) q++;
    if (q > dstart) return contig.substr(s, q - s);
    p += 1;
  }
  die("Could not extract taxon ID from contig identifier '" + contig + "' - did you use the MetMaps build scripts to construct your database?");
}

size_t overlap(size_t a1, size_t a2, size_t b1, size_t b2) {            // util.h:150: closed intervals
  size_t lo = std::max(a1, b1), hi = std::min(a2, b2);
  return hi >= lo ? hi - lo + 1 : 0;
}

// tail sums for the evidence file (fEM.h:1101-1108): Poisson pmf at 0 and binomial cdf
double binom_cdf_host(long k, long n, double p) {
  if (k < 0) return 0; if (k >= n) return 1;
  double s = 0;
  for (long j = 0; j <= k; j++) s += std::exp(std::lgamma(n + 1.0) - std::lgamma(j + 1.0) - std::lgamma(n - j + 1.0) + j * std::log(p) + (n - j) * std::log1p(-p));
  return std::min(1.0, s);
}

void classify_one(const std::string& DB, const std::string& mapped, int device, size_t minReadsPerBest) {
  PhaseTimer pt("classify: whole");
  // ---- read the mappings file, grouped by consecutive read id (fEM.h:1167-1214)
  std::vector<std::vector<std::string>> fields; std::vector<int64_t> readOff{0}; std::vector<std::string> lines;
  {
    std::ifstream in(mapped);
    if (!in.is_open()) die("Cannot open mappings file " + mapped);
    std::string line, running;
    while (std::getline(in, line)) {
      erase_nl(line); if (line.empty()) continue;
      std::vector<std::string> f = split(line, " ");
      if (f.size() < 14) die("File " + mapped + " has weird format - is this a mappings file generated by MetaMap?");
      if (f[0] != running) { if (!fields.empty()) readOff.push_back((int64_t)fields.size()); running = f[0]; }
      fields.push_back(f); lines.push_back(line);
    }
    if (!fields.empty()) readOff.push_back((int64_t)fields.size());
  }
  size_t M = fields.size(); int64_t nReads = (int64_t)readOff.size() - 1;
  std::vector<std::string> mTaxon(M);
  std::set<std::string> relevant;
  for (size_t m = 0; m < M; m++) { mTaxon[m] = extract_taxon(fields[m][5]); relevant.insert(mTaxon[m]); }
  if (relevant.empty()) die("No relevant taxon IDs found in your mappings file - is it possible that none of your reads are mapped?");
  std::map<std::string, size_t> stats;
  {
    std::ifstream s(mapped + ".meta");
    if (!s.is_open()) die("The file " + mapped + ".meta is not present or could not be opened - this file is generated automatically as part of the mapping process, so please check whether the mapping process finished successfully.");
    std::string line; while (std::getline(s, line)) { erase_nl(line); if (line.empty()) continue; auto f = split(line, " "); stats[f[0]] = strtoull(f[1].c_str(), nullptr, 10); }
  }
  size_t nUnmapped = stats.at("ReadsNotMapped"), nTooShort = stats.at("ReadsTooShort"), nTotal = stats.at("TotalReads"), nMapped = stats.at("ReadsMapped");
  // ---- taxonInfo.txt (fEM.h:1320-1364)
  std::map<std::string, std::map<std::string, size_t>> taxonInfo;
  {
    std::ifstream t(DB + "/taxonInfo.txt");
    if (!t.is_open()) die("Could not open file " + DB + "/taxonInfo.txt -- perhaps you have specified an incomplete DB?");
    std::string line;
    while (std::getline(t, line)) {
      erase_nl(line); if (line.empty()) continue;
      auto f = split(line, " ");
      if (f.size() != 2) die("Weird format in " + DB + "/taxonInfo.txt -- wrong number of fields.");
      if (!relevant.count(f[0])) continue;
      for (auto& c : split(f[1], ";")) { auto kv = split(c, "="); taxonInfo[f[0]][kv[0]] = strtoull(kv[1].c_str(), nullptr, 10); }
    }
  }
  Taxonomy T(DB + "/taxonomy");
  std::vector<std::string> taxa(relevant.begin(), relevant.end());       // std::map order of the reference's f
  std::map<std::string, int32_t> tIdx; for (size_t i = 0; i < taxa.size(); i++) tIdx[taxa[i]] = (int32_t)i;
  // ---- per mapping: taxon index, mapq, nloc (fEM.h:246-348)
  std::vector<int32_t> taxon(M); std::vector<double> mapq(M), nloc(M), identity(M); std::vector<long long> rlen((size_t)nReads);
  for (int64_t r = 0; r < nReads; r++) {
    long long L = std::stoi(fields[(size_t)readOff[(size_t)r]][1]); rlen[(size_t)r] = L;
    std::set<std::string> sawContigs; std::map<std::string, double> perTaxon;
    for (int64_t m = readOff[(size_t)r]; m < readOff[(size_t)r + 1]; m++) sawContigs.insert(fields[(size_t)m][5]);
    for (int64_t m = readOff[(size_t)r]; m < readOff[(size_t)r + 1]; m++) {
      const std::string& t = mTaxon[(size_t)m];
      if (!taxonInfo.count(t)) die("Unknown taxonID '" + t + "'; please check that your mappings file was mapped against the database now specified.");
      if (!perTaxon.count(t)) {
        size_t n = 0;
        for (auto& c : taxonInfo.at(t)) { if ((long long)c.second >= L) n += (c.second - L + 1); else if (sawContigs.count(c.first)) n++; }
        perTaxon[t] = (double)n;
      }
      taxon[(size_t)m] = tIdx.at(t); nloc[(size_t)m] = perTaxon[t];
      double q; try { q = std::stod(fields[(size_t)m][13]); } catch (const std::out_of_range&) { q = 0; }
      mapq[(size_t)m] = q; identity[(size_t)m] = std::stod(fields[(size_t)m][9]) / 100.0;
    }
  }
  // ---- EM on the GPU (fEM.h:491-661)
  std::cout << "Starting EM..." << std::endl;
  PhaseTimer pem("classify: CUDA context + EM");
  mm_ctx* ctx = nullptr; ck(mm_ctx_create(device, &ctx), "mm_ctx_create");
  int32_t Tn = (int32_t)taxa.size(), iters = 0;
  std::vector<double> f((size_t)Tn), post(M), ll(4096); std::vector<int64_t> best((size_t)std::max<int64_t>(nReads, 1));
  ck(mm_em_run(ctx, taxon.data(), mapq.data(), nloc.data(), readOff.data(), nReads, Tn, 0, f.data(), post.data(), best.data(), ll.data(), 4096, &iters), "mm_em_run");
  mm_ctx_destroy(ctx);
  for (int i = 0; i < iters && i < 4096; i++) {
    std::cout << "EM round " << i << "\n\n\tLog likelihood: " << ll[(size_t)i] << std::endl;
    if (i > 0) std::cout << "\tImprovement: " << ll[(size_t)i] - ll[(size_t)i - 1] << "\n\tRelative   : " << ll[(size_t)i] / ll[(size_t)i - 1] << std::endl;
  }
  // ---- final pass (fEM.h:663-782)
  std::ofstream oId(mapped + ".EM.lengthAndIdentitiesPerMappingUnit"), oEM(mapped + ".EM"), oR2T(mapped + ".EM.reads2Taxon"), oKrona(mapped + ".EM.reads2Taxon.krona");
  oId << "AnalysisLevel\tID\treadI\tIdentity\tLength\n";
  const size_t W = 1000;
  std::map<std::string, std::map<std::string, std::vector<size_t>>> cov, covReads; std::map<std::string, std::map<std::string, size_t>> lastWin;
  std::map<std::string, size_t> readsPerTaxon; std::map<std::string, std::vector<double>> idPerTaxon; long long maxReadLen = -1;
  std::cout << "Outputting mappings with adjusted alignment qualities." << std::endl;
  for (int64_t r = 0; r < nReads; r++) {
    for (int64_t m = readOff[(size_t)r]; m < readOff[(size_t)r + 1]; m++) {
      std::vector<std::string> f2 = fields[(size_t)m]; f2[13] = std::to_string(post[(size_t)m]);
      for (size_t i = 0; i < f2.size(); i++) oEM << (i ? " " : "") << f2[i];
      oEM << "\n";
    }
    size_t b = (size_t)best[(size_t)r];
    const std::string& bt = mTaxon[b]; const std::string& bc = fields[b][5];
    const std::string& readID = fields[b][0];
    oId << "EqualCoverageUnit\t" << bc << "\t" << r << "\t" << identity[b] << "\t" << rlen[(size_t)r] << "\n";
    oR2T << readID << "\t" << bt << "\n";
    oKrona << readID << "\t" << T.first_non_x(bt) << "\t" << post[b] << "\n";
    idPerTaxon[bt].push_back(identity[b]);
    if (rlen[(size_t)r] > maxReadLen) maxReadLen = rlen[(size_t)r];
    readsPerTaxon[bt]++;
    size_t clen = taxonInfo.at(bt).at(bc);
    if (!cov[bt].count(bc)) {                                            // fEM.h:730-752
      size_t nw = clen / W;
      if (nw == 0) { nw++; lastWin[bt][bc] = clen; }
      else if (nw * W != clen) { nw++; lastWin[bt][bc] = clen - (nw * W); }
      else lastWin[bt][bc] = W;
      cov[bt][bc].assign(nw, 0); covReads[bt][bc].assign(nw, 0);
    }
    size_t start = strtoull(fields[b][7].c_str(), nullptr, 10), stop = strtoull(fields[b][8].c_str(), nullptr, 10);
    size_t stopPos = stop >= clen ? clen - 1 : stop;
    for (size_t p = start; p <= stopPos; p += W) {                        // fEM.h:755-776
      size_t wi = p / W, ws = wi * W, we = (wi + 1) * W - 1;
      if (we > clen) we = clen - 1;
      cov[bt][bc].at(wi) += overlap(ws, we, start, stopPos);
      covReads[bt][bc].at(wi)++;
    }
  }
  {
    std::ifstream u(mapped + ".meta.unmappedReadsLengths"); std::string line;   // fEM.h:785-790
    while (std::getline(u, line)) { erase_nl(line); if (line.empty()) continue; auto f2 = split(line, "\t"); oR2T << f2[1] << "\t0\n"; oKrona << f2[1] << "\t0\t0\n"; }
  }
  oId.close(); oEM.close(); oR2T.close(); oKrona.close();
  // ---- cleanF (fEM.h:1135-1163)
  std::map<std::string, double> fm; for (size_t i = 0; i < taxa.size(); i++) fm[taxa[i]] = f[i];
  {
    double minFreq = 0.9 * (1.0 / (double)nMapped);
    std::set<std::string> del; for (auto& e : fm) if (e.second < minFreq && !readsPerTaxon.count(e.first)) del.insert(e.first);
    for (auto& d : del) fm.erase(d);
    double s = 0; for (auto& e : fm) s += e.second;
    for (auto& e : fm) e.second /= s;
  }
  // ---- producePotFile (fEM.h:52-215)
  {
    const std::set<std::string> levels = {"species", "genus", "family", "order", "phylum", "superkingdom"};
    std::map<std::string, std::set<std::string>> keys; std::map<std::string, std::map<std::string, double>> fl; std::map<std::string, std::map<std::string, size_t>> rc;
    for (auto& e : fm) {
      auto up = T.upward_by_ranks(e.first, levels); up["definedGenomes"] = e.first;
      for (auto& u : up) { fl[u.first][u.second] += e.second; keys[u.first].insert(u.second); if (fl[u.first][u.second] > 1) fl[u.first][u.second] = 1; }
    }
    for (auto& e : readsPerTaxon) {
      auto up = T.upward_by_ranks(e.first, levels); up["definedGenomes"] = e.first;
      for (auto& u : up) {
        if (fl[u.first].count(u.second) == 0) rc[u.first][u.second] = 0;       // (sic) fEM.h:105-106
        rc[u.first][u.second] += e.second; keys[u.first].insert(u.second);
      }
    }
    long long nMappable = (long long)nTotal - (long long)nTooShort, nMap2 = nMappable - (long long)nUnmapped;
    std::ofstream o(mapped + ".EM.WIMP");
    o << "AnalysisLevel\ttaxonID\tName\tAbsolute\tEMFrequency\tPotFrequency\n";
    for (auto& l : keys) {
      const std::string& lv = l.first; std::map<std::string, double> orig; double sumAssigned = 0;
      for (auto& t : l.second) { double v = fl[lv].count(t) ? fl[lv][t] : 0; size_t c = rc[lv].count(t) ? rc[lv][t] : 0; sumAssigned += v; fl[lv][t] = v; rc[lv][t] = c; }
      for (auto& t : l.second) { fl[lv][t] /= sumAssigned; orig[t] = fl[lv][t]; }
      double propMapped = (double)nMap2 / nMappable, propNot = (double)nUnmapped / nMappable;
      for (auto& t : l.second) fl[lv][t] *= propMapped;
      double emUnm = 0; size_t nUnd = nUnmapped;
      for (auto& t : l.second) {
        if (t != "Undefined") o << lv << "\t" << t << "\t" << T.T.at(t).name << "\t" << rc[lv][t] << "\t" << orig[t] << "\t" << fl[lv][t] << "\n";
        else { nUnd += rc[lv][t]; emUnm += orig[t]; propNot += fl[lv][t]; }
      }
      o << lv << "\t0\tUnclassified\t" << nUnd << "\t" << emUnm << "\t" << propNot << "\n";
      o << lv << "\t-3\ttotalReads\t" << nTotal << "\t0\t0\n" << lv << "\t-3\treadsLongEnough\t" << nMappable << "\t0\t0\n" << lv << "\t-3\treadsLongEnough_unmapped\t" << nUnmapped << "\t0\t0\n";
    }
  }
  // ---- contig coverage (fEM.h:805-844)
  std::map<std::string, std::string> contig2taxon;
  {
    std::ofstream o(mapped + ".EM.contigCoverage");
    o << "taxonID\tequalCoverageUnitLabel\tcontigID\tstart\tstop\tnBases\treadCoverage\n";
    for (auto& t : cov) for (auto& c : t.second) {
      for (size_t wi = 0; wi < c.second.size(); wi++) {
        size_t wl = (wi == c.second.size() - 1) ? lastWin.at(t.first).at(c.first) : W;
        size_t nb = c.second[wi];
        o << t.first << "\t" << T.T.at(t.first).name << "\t" << c.first << "\t" << wi * W << "\t" << (wi + 1) * W - 1 << "\t" << nb << "\t" << (double)nb / (double)wl << "\n";
      }
      contig2taxon[c.first] = t.first;
    }
  }
  // ---- evidence for unknown species (fEM.h:846-1132)
  {
    std::string bestTaxon; double bestMedian = 0, oneThird = 0, oneThirdP = 0;
    for (auto e : idPerTaxon) {
      auto ids = e.second;
      if (ids.size() >= 3 && ids.size() >= minReadsPerBest) {
        std::sort(ids.begin(), ids.end());
        double med = ids.at(ids.size() / 2);
        if (bestTaxon.empty() || med > bestMedian) {
          bestMedian = med; bestTaxon = e.first; oneThird = ids.at((size_t)(ids.size() * (1.0 / 3.0)));
          size_t n = 0; for (double v : ids) if (v <= oneThird) n++;
          oneThirdP = (double)n / (double)ids.size();
        }
      }
    }
    std::map<std::string, std::vector<size_t>> Ns;
    {
      std::ifstream w(DB + "/contigNstats_windowSize_1000.txt"); std::string line;
      while (std::getline(w, line)) {
        erase_nl(line); if (line.empty()) continue; auto f2 = split(line, "\t");
        if (f2.size() != 3) continue;
        if (cov.count(f2[0]) && cov.at(f2[0]).count(f2[1])) { std::vector<size_t> v; for (auto& x : split(f2[2], ";")) v.push_back(strtoull(x.c_str(), nullptr, 10)); Ns[f2[1]] = v; }
      }
      for (auto& t : cov) for (auto& c : t.second) if (!Ns.count(c.first)) die("\nMissing entry " + c.first + " in " + DB + "/contigNstats_windowSize_1000.txt\n");
    }
    std::map<std::string, size_t> gw, gwUse, gwReads; std::map<std::string, double> gwZero;
    size_t need = (size_t)maxReadLen;
    for (auto& cd : Ns) {
      const std::string& t = contig2taxon.at(cd.first); size_t n = cd.second.size();
      std::vector<size_t> fw(n, 0), bw(n, 0); size_t run = 0;
      auto wlen = [&](size_t wi) { return wi == n - 1 ? lastWin.at(t).at(cd.first) : W; };
      for (size_t wi = 0; wi < n; wi++) { fw[wi] = run; if ((double)cd.second[wi] / (double)wlen(wi) <= 0.02) run += wlen(wi); else run = 0; }
      run = 0;
      for (long long wi = (long long)n - 1; wi >= 0; wi--) { bw[(size_t)wi] = run; if ((double)cd.second[(size_t)wi] / (double)wlen((size_t)wi) <= 0.02) run += wlen((size_t)wi); else run = 0; }
      size_t use = 0, useReads = 0, useZero = 0;
      for (size_t wi = 0; wi < n; wi++) if (fw[wi] >= need && bw[wi] >= need) { use++; size_t c = covReads.at(t).at(cd.first).at(wi); useReads += c; if (c == 0) useZero++; }
      gw[t] += n; gwUse[t] += use; gwReads[t] += useReads; gwZero[t] += useZero;
    }
    std::ofstream o(mapped + ".EM.evidenceUnknownSpecies");
    o << "taxonID\tspecies\tgenus\tnReads\tpropBottomThirdReadIdentities\texpectedPropBottomThirdReadIdentities\tpValue_BottomThirdReadIdentities\tcoverageWindows_totalGenome"
         "\tcoverageWindows_usable\tcoverageWindows_usable_averageCoverage\tcoverageWindows_usable_coverageIsZero\tcoverageWindows_usable_coverageIsZero_expected"
         "\tcoverageWindows_usable_coverageIsZero_P\n";
    for (auto& e : idPerTaxon) {
      const std::string& t = e.first; const auto& ids = e.second;
      std::string sProp = "NA", sP = "NA", sExp = "NA";
      if (!bestTaxon.empty()) {
        size_t obs = 0; for (double v : ids) if (v <= oneThird) obs++;
        size_t obsNon = ids.size() - obs; double ex = oneThirdP * ids.size(), exNon = ids.size() - ex;
        sExp = std::to_string(oneThirdP);
        double stat = std::pow(obs - ex, 2) / ex + std::pow(obsNon - exNon, 2) / exNon;
        sProp = std::to_string((double)obs / (double)ids.size());
        sP = std::to_string(1 - std::erf(std::sqrt(stat / 2.0)));          // chi-squared, 1 d.f.
      }
      std::string sAvg = "NA", sZeroExp = "NA", sZeroP = "NA";
      if (gwUse.at(t) > 0) {
        double avg = (double)gwReads.at(t) / (double)gwUse.at(t); sAvg = std::to_string(avg);
        if (avg == 0) { sZeroExp = std::to_string(gwUse.at(t)); sZeroP = std::to_string(1); }
        else {
          double p0 = std::exp(-avg); sZeroExp = std::to_string(gwUse.at(t) * p0);
          double pv = 1;
          if (gwZero.at(t) > 0) pv = 1 - binom_cdf_host((long)gwZero.at(t) - 1, (long)gwUse.at(t), p0);
          sZeroP = std::to_string(pv);
        }
      }
      auto up = T.upward_by_ranks(t, {"species", "genus"});
      o << t << "\t" << up.at("species") << "\t" << up.at("genus") << "\t" << ids.size() << "\t" << sProp << "\t" << sExp << "\t" << sP << "\t" << gw.at(t) << "\t"
        << gwUse.at(t) << "\t" << sAvg << "\t" << gwZero.at(t) << "\t" << sZeroExp << "\t" << sZeroP << "\n";
    }
  }
}

int run_classify(int argc, char** argv) {
  const std::map<std::string, std::string> alias = {{"DB", "DB"}, {"mappings", "mappings"}, {"minreads", "minreads"}, {"threads", "threads"}, {"t", "threads"}, {"device", "device"}};
  Options o = parse_args(argc, argv, 2, alias, {});
  if (!o.has("DB") || !o.has("mappings")) die("classify needs --DB and --mappings");
  int device = o.has("device") ? atoi(o.get("device").c_str()) : 0;
  size_t minReads = o.has("minreads") ? strtoull(o.get("minreads").c_str(), nullptr, 10) : 10000;      // parseCmdArgs.hpp:462-471
  for (auto& m : split(o.get("mappings"), ",")) classify_one(o.get("DB"), m, device, minReads);
  return 0;
}

void usage() {
  std::cout << "\nMetaMaps (metamaps_b200: B200 compute core) \n\n  Simultaneous metagenomic classification and mapping.\n\nUsage:\n\n  ./metamaps mapDirectly|classify|index|mapAgainstIndex\n\n"
               "  mapDirectly -r <ref.fa[.gz]> -q <reads.fq[,..]> -o <prefix[,..]> [--all] [-k 16] [-w W | -p 1e-3] [-m 1000] [--pi 80] [-t N] [--maxmemory GB] [--device 0]\n"
               "  classify --DB <dir> --mappings <prefix[,..]> [-t N] [--device 0]\n\n";
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { usage(); return 1; }
  std::string cmd = argv[1];
  if (cmd == "mapDirectly") return run_mapDirectly(argc, argv);
  if (cmd == "classify") return run_classify(argc, argv);
  if (cmd == "index") return run_index(argc, argv);
  if (cmd == "mapAgainstIndex") return run_mapAgainstIndex(argc, argv);
  if (cmd == "classifyU") die("sub-command 'classifyU' is disabled in the reference (mash_map.cpp:323) and not part of the GPU hot path");
  usage();
  return 1;
}
