// mm_lib.cu -- the C ABI declared in include/metamaps_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -> libmetamaps_b200.so (see metamaps_b200/build.py).
// (tests/_emu compiles this same file with g++ -DMM_HOST_EMU to check kernel logic without a GPU;
//  that library is test infrastructure and is never loaded by the product.)
#include "../../include/metamaps_b200.h"

#include "mm_classify.h"
#include "mm_index.h"
#include "mm_map.h"
#include "mm_mapq.h"

#ifndef MM_HOST_EMU
#include <dlfcn.h>
#include <thread>
#endif

using namespace mm;

static thread_local std::string g_err;

struct mm_ctx {
  Runtime rt;
  Prims pr;
  Sketcher sk;
  Mapper mp;
  SeqBatch sketchBatch;
  SketchOut sketchOut;
  double last_ms = 0;
  int64_t last_launches = 0;
  // grow-only scratch of the mapq / EM / fetch entry points (cudaMalloc + cudaFree per call would cost more than the kernels)
  struct {
    DevBuf<double> dId, dMq, dNl;
    DevBuf<uint32_t> gOne, gAll, gSendA, gRecvA, gSendB, gRecvB; DevBuf<int64_t> gRange;      // collectives' slabs (grow-only: no cudaMalloc / cudaFree per batch)
    DevBuf<int32_t> dSh, dSk, dLen, dSt, dTax, tmpA, tmpB;
    DevBuf<int64_t> dOff;
  } scr;
  // double-buffered input staging (mm_stage_reads_async)
  struct Stage { DevBuf<uint8_t> asc; std::vector<int64_t> off; int32_t n = -1;
                 SeqBatch batch; bool packed = false; const uint8_t* hostAsc = nullptr;     // zero-copy mode: packed straight from pinned host memory
#ifndef MM_HOST_EMU
                 cudaEvent_t ready = nullptr; std::thread feeder; int feedErr = 0;
                 void join() { if (feeder.joinable()) feeder.join(); }
#else
                 void join() {}
#endif
  } stage[2];
  // NCCL (multi-GPU EM), bound at run time
  void* comm = nullptr; int nRanks = 1, rank = 0;
  mm_allreduce_fn hostAllreduce = nullptr; void* hostAllreduceUser = nullptr; std::vector<double> hostBuf;
  Classifier cls;
  mm_ctx() : pr(rt), sk(rt, pr), mp(rt, pr, sk), cls(rt, pr) {}
};
struct mm_index {
  mm_ctx* ctx;
  Index ix;
  mm_index(mm_ctx* c, int k, int w) : ctx(c), ix(c->rt, c->pr, c->sk, k, w) {}
};

#define MM_TRY try {
#define MM_CATCH                                                   \
  }                                                                \
  catch (const mm::Error& e) { g_err = e.what(); return e.code; }  \
  catch (const std::bad_alloc&) { g_err = "out of host memory"; return MM_ENOMEM; } \
  catch (const std::exception& e) { g_err = e.what(); return MM_EINVAL; }           \
  return MM_OK;

static void begin_call(mm_ctx* c) {
#ifndef MM_HOST_EMU
  MM_CUDA(cudaSetDevice(c->rt.device));
#endif
  c->rt.launches = 0; c->last_ms = 0;
}
static void end_call(mm_ctx* c) { c->rt.sync(); c->rt.resolve_timers(); c->last_launches = c->rt.launches; }

extern "C" {

const char* mm_last_error(void) { return g_err.c_str(); }
const char* mm_version(void) {
#ifdef MM_HOST_EMU
  return "metamaps_b200 0.1 (HOST EMULATION - test build, not a product path)";
#else
  return "metamaps_b200 0.1 (sm_100a)";
#endif
}

int mm_ctx_create(int device, mm_ctx** out) {
  MM_TRY
  if (!out) throw Error(MM_EINVAL, "mm_ctx_create: out is NULL");
#ifndef MM_HOST_EMU
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); throw Error(MM_ENODEV, "no CUDA device available: this library has no CPU path"); }
  if (device < 0 || device >= n) throw Error(MM_EINVAL, "device ordinal out of range");
  MM_CUDA(cudaSetDevice(device));
#endif
  mm_ctx* c = new mm_ctx();
  c->rt.device = device;
#ifndef MM_HOST_EMU
  cudaDeviceProp prop; MM_CUDA(cudaGetDeviceProperties(&prop, device));
  c->rt.sm_count = prop.multiProcessorCount;
  MM_CUDA(cudaStreamCreateWithFlags(&c->rt.stream, cudaStreamNonBlocking));
  MM_CUDA(cudaStreamCreateWithFlags(&c->rt.side, cudaStreamNonBlocking));
  MM_CUDA(cudaStreamCreateWithFlags(&c->rt.copy, cudaStreamNonBlocking));
#endif
  *out = c;
  MM_CATCH
}
void mm_ctx_destroy(mm_ctx* c) {
  if (!c) return;
#ifndef MM_HOST_EMU
  cudaSetDevice(c->rt.device);
  cudaStreamSynchronize(c->rt.stream);
#endif
  mm_comm_destroy(c);
#ifndef MM_HOST_EMU
  cudaStream_t s = c->rt.stream, s2 = c->rt.side, s3 = c->rt.copy;
  cudaStreamSynchronize(s2); cudaStreamSynchronize(s3);
  for (auto& st : c->stage) { st.join(); if (st.ready) cudaEventDestroy(st.ready); }
  delete c;
  cudaStreamDestroy(s); cudaStreamDestroy(s2); cudaStreamDestroy(s3);
#else
  delete c;
#endif
}
int mm_ctx_last_timing(mm_ctx* c, double* total_ms, int64_t* n_launches) {
  if (!c) return MM_EINVAL;
  if (total_ms) *total_ms = c->last_ms;
  if (n_launches) *n_launches = c->last_launches;
  return MM_OK;
}
int mm_ctx_mem_info(mm_ctx* c, int64_t* free_bytes, int64_t* total_bytes) {
  MM_TRY
  if (!c) throw Error(MM_EINVAL, "null ctx");
#ifndef MM_HOST_EMU
  MM_CUDA(cudaSetDevice(c->rt.device));
  size_t f = 0, t = 0; MM_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
#else
  if (free_bytes) *free_bytes = (int64_t)8 << 30;       // the emulation build pretends to be an 8 GB device
  if (total_bytes) *total_bytes = (int64_t)8 << 30;
#endif
  MM_CATCH
}
int mm_ctx_last_map_stats(mm_ctx* c, double* stage_ms, int64_t* counters) {
  if (!c) return MM_EINVAL;
  for (int i = 0; i < 16; i++) { if (stage_ms) stage_ms[i] = c->mp.st.ms[i]; if (counters) counters[i] = c->mp.st.counters[i]; }
  return MM_OK;
}

// ------------------------------------------------------------------------------------------------ K1
int mm_sketch_batch(mm_ctx* c, const char* seqs, const int64_t* offsets, int32_t n, int k, int w, int64_t* n_total) {
  MM_TRY
  if (!c || !offsets || n < 0 || (!seqs && n > 0)) throw Error(MM_EINVAL, "mm_sketch_batch: bad arguments");
  if (k < 1 || k > 16) throw Error(MM_EINVAL, "k-mer size must be in [1,16] (parseCmdArgs.hpp:62)");
  if (w < 1) throw Error(MM_EINVAL, "window size must be >= 1");
  begin_call(c);
  c->sk.load(c->sketchBatch, seqs, nullptr, offsets, n);
  { StageTimer t(c->rt, &c->last_ms); c->sk.run(c->sketchBatch, k, w, c->sketchOut); }
  end_call(c);
  if (n_total) *n_total = c->sketchOut.n_total;
  MM_CATCH
}
struct UnpackWsFn {
  const uint32_t* ws; int32_t* wpos; int32_t* strand;
  MM_HD void operator()(int64_t i) const { uint32_t v = ldg(ws + i); wpos[i] = (int32_t)(v >> 1); strand[i] = (v & 1u) ? 1 : -1; }
};
int mm_sketch_fetch(mm_ctx* c, int64_t* counts, uint32_t* hash, int32_t* wpos, int32_t* strand) {
  MM_TRY
  if (!c) throw Error(MM_EINVAL, "null ctx");
  begin_call(c);
  SketchOut& o = c->sketchOut;
  if (counts) { d2h(c->rt, counts, o.seqOff.p, sizeof(int64_t) * ((size_t)o.n_seqs + 1)); }
  if (hash) d2h(c->rt, hash, o.hash.p, sizeof(uint32_t) * (size_t)o.n_total);
  if ((wpos || strand) && o.n_total) {
    DevBuf<int32_t> a, b; a.ensure((size_t)o.n_total); b.ensure((size_t)o.n_total);
    foreach(c->rt, o.n_total, UnpackWsFn{o.ws.p, a.p, b.p});
    if (wpos) d2h(c->rt, wpos, a.p, sizeof(int32_t) * (size_t)o.n_total);
    if (strand) d2h(c->rt, strand, b.p, sizeof(int32_t) * (size_t)o.n_total);
    c->rt.sync();
  }
  MM_CATCH
}

// ------------------------------------------------------------------------------------------------ K2
int mm_index_create(mm_ctx* c, int k, int w, mm_index** out) {
  MM_TRY
  if (!c || !out) throw Error(MM_EINVAL, "mm_index_create: bad arguments");
  if (k < 1 || k > 16) throw Error(MM_EINVAL, "k-mer size must be in [1,16] (parseCmdArgs.hpp:62)");
  if (w < 1) throw Error(MM_EINVAL, "window size must be >= 1");
  *out = new mm_index(c, k, w);
  MM_CATCH
}
static int index_add_impl(mm_index* idx, const char* seqs, const void* dev, const int64_t* offsets, int32_t n) {
  MM_TRY
  if (!idx || !offsets || n < 0) throw Error(MM_EINVAL, "mm_index_add: bad arguments");
  mm_ctx* c = idx->ctx;
  begin_call(c);
  c->sk.load(c->sketchBatch, seqs, dev, offsets, n);
  { StageTimer t(c->rt, &c->last_ms); idx->ix.add(c->sketchBatch); }
  end_call(c);
  MM_CATCH
}
int mm_index_add(mm_index* idx, const char* seqs, const int64_t* offsets, int32_t n) { return index_add_impl(idx, seqs, nullptr, offsets, n); }
int mm_index_add_dev(mm_index* idx, const void* dev, const int64_t* offsets, int32_t n) {
  if (!dev) { g_err = "mm_index_add_dev: null device pointer"; return MM_EINVAL; }
  return index_add_impl(idx, nullptr, dev, offsets, n);
}
int mm_index_finalize(mm_index* idx) {
  MM_TRY
  if (!idx) throw Error(MM_EINVAL, "null index");
  begin_call(idx->ctx);
  { StageTimer t(idx->ctx->rt, &idx->ctx->last_ms); idx->ix.finalize(); }
  end_call(idx->ctx);
  MM_CATCH
}
int mm_index_stats(const mm_index* idx, int64_t* n_min, int64_t* n_unique, int32_t* freq, int32_t* n_contigs, int64_t* bytes) {
  if (!idx) { g_err = "null index"; return MM_EINVAL; }
  if (n_min) *n_min = idx->ix.n;
  if (n_unique) *n_unique = idx->ix.n_unique;
  if (freq) *freq = idx->ix.globalSynced ? idx->ix.globalThreshold : idx->ix.freqThreshold;
  if (n_contigs) *n_contigs = idx->ix.n_contigs;
  if (bytes) *bytes = idx->ix.device_bytes();
  return MM_OK;
}
int mm_index_fetch(const mm_index* idx, uint32_t* hash, int32_t* seq_id, int32_t* wpos, int32_t* strand) {
  MM_TRY
  if (!idx) throw Error(MM_EINVAL, "null index");
  mm_ctx* c = idx->ctx; const Index& ix = idx->ix;
  begin_call(c);
  if (hash) d2h(c->rt, hash, ix.miHash.p, sizeof(uint32_t) * (size_t)ix.n);
  if ((wpos || strand) && ix.n) {
    DevBuf<int32_t> a, b; a.ensure((size_t)ix.n); b.ensure((size_t)ix.n);
    foreach(c->rt, ix.n, UnpackWsFn{ix.miWs.p, a.p, b.p});
    if (wpos) d2h(c->rt, wpos, a.p, sizeof(int32_t) * (size_t)ix.n);
    if (strand) d2h(c->rt, strand, b.p, sizeof(int32_t) * (size_t)ix.n);
    c->rt.sync();
  }
  if (seq_id)
    for (int32_t s = 0; s < ix.n_contigs; s++)
      for (int64_t i = ix.h_contigStart[(size_t)s]; i < ix.h_contigStart[(size_t)s + 1]; i++) seq_id[i] = s;
  MM_CATCH
}
// ---- persistent index ------------------------------------------------------------------------------------------
namespace {
struct IndexFileHeader {
  char magic[8]; uint32_t version, k, w, hasSeq16; int64_t n, n_unique, n_dup, total_bases, tableSlots; int32_t n_contigs, freqThreshold, firstContig,
      globalSynced, globalThreshold, reserved[7];
};
const char INDEX_MAGIC[8] = {'M', 'M', 'B', '2', '0', '0', 'I', 'X'};
void put(FILE* f, const void* p, size_t bytes) { if (bytes && fwrite(p, 1, bytes, f) != bytes) throw Error(MM_EINVAL, "mm_index_save: short write"); }
void get(FILE* f, void* p, size_t bytes) { if (bytes && fread(p, 1, bytes, f) != bytes) throw Error(MM_EINVAL, "mm_index_load: truncated index file"); }
// device array <-> file through a host bounce buffer
void put_dev(Runtime& rt, FILE* f, const void* d, size_t bytes, std::vector<char>& bounce) {
  for (size_t o = 0; o < bytes; o += bounce.size()) {
    const size_t nb = std::min(bounce.size(), bytes - o);
    d2h(rt, bounce.data(), (const char*)d + o, nb); put(f, bounce.data(), nb);
  }
}
void get_dev(Runtime& rt, FILE* f, void* d, size_t bytes, std::vector<char>& bounce) {
  for (size_t o = 0; o < bytes; o += bounce.size()) {
    const size_t nb = std::min(bounce.size(), bytes - o);
    get(f, bounce.data(), nb); h2d(rt, (char*)d + o, bounce.data(), nb); rt.sync();
  }
}
}  // namespace
int mm_index_save(const mm_index* idx, const char* path) {
  MM_TRY
  if (!idx || !path || !idx->ix.finalized) throw Error(MM_EINVAL, "mm_index_save: index not finalized");
  const Index& ix = idx->ix; mm_ctx* c = idx->ctx;
  begin_call(c);
  FILE* f = fopen(path, "wb");
  if (!f) throw Error(MM_EINVAL, std::string("mm_index_save: cannot open ") + path);
  try {
    IndexFileHeader h; memset(&h, 0, sizeof h); memcpy(h.magic, INDEX_MAGIC, 8);
    h.version = 1; h.k = (uint32_t)ix.k; h.w = (uint32_t)ix.w; h.hasSeq16 = ix.hasSeq16 ? 1 : 0; h.n = ix.n; h.n_unique = ix.n_unique; h.n_dup = ix.n_dup;
    h.total_bases = ix.total_bases; h.tableSlots = (int64_t)ix.tableMask + 1; h.n_contigs = ix.n_contigs; h.freqThreshold = ix.freqThreshold;
    h.firstContig = ix.firstContig; h.globalSynced = ix.globalSynced ? 1 : 0; h.globalThreshold = ix.globalThreshold;
    put(f, &h, sizeof h);
    put(f, ix.h_contigLen.data(), sizeof(int32_t) * ix.h_contigLen.size());
    put(f, ix.h_contigStart.data(), sizeof(int64_t) * ix.h_contigStart.size());
    std::vector<char> bounce((size_t)64 << 20);
    put_dev(c->rt, f, ix.miHash.p, 4 * (size_t)ix.n, bounce); put_dev(c->rt, f, ix.miWs.p, 4 * (size_t)ix.n, bounce);
    put_dev(c->rt, f, ix.table.p, sizeof(Slot) * (size_t)h.tableSlots, bounce);
    put_dev(c->rt, f, ix.posKey.p, 8 * (size_t)ix.n, bounce);
    if (ix.hasSeq16) put_dev(c->rt, f, ix.posSeq16.p, 2 * (size_t)ix.n, bounce);
    put_dev(c->rt, f, ix.dupBits.p, 4 * (size_t)(ix.n / 32 + 2), bounce);
    put_dev(c->rt, f, ix.dupIdx.p, 4 * (size_t)ix.n_dup, bounce); put_dev(c->rt, f, ix.dupLinks.p, 8 * (size_t)ix.n_dup, bounce);
  } catch (...) { fclose(f); throw; }
  if (fclose(f) != 0) throw Error(MM_EINVAL, "mm_index_save: close failed");
  MM_CATCH
}
int mm_index_load(mm_ctx* c, const char* path, mm_index** out) {
  MM_TRY
  if (!c || !path || !out) throw Error(MM_EINVAL, "mm_index_load: bad arguments");
  begin_call(c);
  FILE* f = fopen(path, "rb");
  if (!f) throw Error(MM_EINVAL, std::string("mm_index_load: cannot open ") + path);
  mm_index* idx = nullptr;
  try {
    IndexFileHeader h; get(f, &h, sizeof h);
    if (memcmp(h.magic, INDEX_MAGIC, 8) != 0 || h.version != 1) throw Error(MM_EINVAL, "mm_index_load: not a metamaps_b200 index file (or another version)");
    if (h.k < 1 || h.k > 16 || h.w < 1 || h.n < 0 || h.n_contigs < 0 || h.n_dup < 0 || h.n_dup > h.n || h.n_unique < 0 || h.n_unique > h.n || h.tableSlots < 1 ||
        h.tableSlots > ((int64_t)1 << 32) || (h.tableSlots & (h.tableSlots - 1)) || h.n >= 0xFFFFFFF0ll)
      throw Error(MM_EINVAL, "mm_index_load: corrupt header");
    {   // the header must agree with the file size before anything is allocated from it
      const int64_t expect = (int64_t)sizeof h + 4 * (int64_t)h.n_contigs + 8 * ((int64_t)h.n_contigs + 1) + 8 * h.n + (int64_t)sizeof(Slot) * h.tableSlots + 8 * h.n +
                             (h.hasSeq16 ? 2 * h.n : 0) + 4 * (h.n / 32 + 2) + 12 * h.n_dup;
      const long at = ftell(f);
      if (fseek(f, 0, SEEK_END) != 0) throw Error(MM_EINVAL, "mm_index_load: cannot seek");
      const int64_t size = (int64_t)ftell(f);
      if (fseek(f, at, SEEK_SET) != 0) throw Error(MM_EINVAL, "mm_index_load: cannot seek");
      if (size != expect) throw Error(MM_EINVAL, "mm_index_load: file size does not match its header (truncated or corrupt index file)");
    }
    idx = new mm_index(c, (int)h.k, (int)h.w);
    Index& ix = idx->ix;
    ix.n = h.n; ix.n_unique = h.n_unique; ix.n_dup = h.n_dup; ix.total_bases = h.total_bases; ix.tableMask = (uint32_t)(h.tableSlots - 1); ix.n_contigs = h.n_contigs;
    ix.freqThreshold = h.freqThreshold; ix.firstContig = h.firstContig; ix.globalSynced = h.globalSynced != 0; ix.globalThreshold = h.globalThreshold; ix.hasSeq16 = h.hasSeq16 != 0;
    ix.h_contigLen.resize((size_t)h.n_contigs); ix.h_contigStart.resize((size_t)h.n_contigs + 1);
    get(f, ix.h_contigLen.data(), sizeof(int32_t) * ix.h_contigLen.size());
    get(f, ix.h_contigStart.data(), sizeof(int64_t) * ix.h_contigStart.size());
    ix.contigStart.ensure(ix.h_contigStart.size()); h2d(c->rt, ix.contigStart.p, ix.h_contigStart.data(), sizeof(int64_t) * ix.h_contigStart.size());
    ix.contigLen.ensure(ix.h_contigLen.size() + 1); h2d(c->rt, ix.contigLen.p, ix.h_contigLen.data(), sizeof(int32_t) * ix.h_contigLen.size());
    c->rt.sync();
    std::vector<char> bounce((size_t)64 << 20);
    ix.miHash.ensure((size_t)ix.n + 1); ix.miWs.ensure((size_t)ix.n + 1);
    get_dev(c->rt, f, ix.miHash.p, 4 * (size_t)ix.n, bounce); get_dev(c->rt, f, ix.miWs.p, 4 * (size_t)ix.n, bounce);
    ix.table.ensure((size_t)h.tableSlots); get_dev(c->rt, f, ix.table.p, sizeof(Slot) * (size_t)h.tableSlots, bounce);
    ix.posKey.ensure((size_t)ix.n + 1); get_dev(c->rt, f, ix.posKey.p, 8 * (size_t)ix.n, bounce);
    if (ix.hasSeq16) { ix.posSeq16.ensure((size_t)ix.n + 8); get_dev(c->rt, f, ix.posSeq16.p, 2 * (size_t)ix.n, bounce); }
    ix.dupBits.ensure((size_t)(ix.n / 32 + 2)); get_dev(c->rt, f, ix.dupBits.p, 4 * (size_t)(ix.n / 32 + 2), bounce);
    ix.dupIdx.ensure((size_t)ix.n_dup + 1); ix.dupLinks.ensure((size_t)ix.n_dup + 1);
    get_dev(c->rt, f, ix.dupIdx.p, 4 * (size_t)ix.n_dup, bounce); get_dev(c->rt, f, ix.dupLinks.p, 8 * (size_t)ix.n_dup, bounce);
    ix.build_dup_rank();
    ix.finalized = true;
  } catch (...) { fclose(f); delete idx; throw; }
  fclose(f);
  *out = idx;
  MM_CATCH
}
int mm_index_params(const mm_index* idx, int32_t* k, int32_t* w, int32_t* contig_len) {
  if (!idx) { g_err = "null index"; return MM_EINVAL; }
  if (k) *k = idx->ix.k;
  if (w) *w = idx->ix.w;
  if (contig_len) for (int32_t i = 0; i < idx->ix.n_contigs; i++) contig_len[i] = idx->ix.h_contigLen[(size_t)i];
  return MM_OK;
}
int mm_index_max_batch_reads(const mm_index* idx, int64_t* max_reads) {
  if (!idx || !max_reads) { g_err = "mm_index_max_batch_reads: bad arguments"; return MM_EINVAL; }
  const HitKeyLayout lay = hit_key_layout(idx->ix);
  const int bits = 64 - lay.seqBits - lay.wsBits;
  *max_reads = bits >= 31 ? (int64_t)0x7fffffff : (bits < 1 ? 1 : ((int64_t)1 << bits));
  return MM_OK;
}
int mm_index_lookup(const mm_index* idx, const uint32_t* hashes, int64_t n, int32_t* counts) {
  MM_TRY
  if (!idx || !idx->ix.finalized) throw Error(MM_EINVAL, "index not finalized");
  mm_ctx* c = idx->ctx;
  begin_call(c);
  DevBuf<uint32_t> h; DevBuf<int32_t> o; h.ensure((size_t)n); o.ensure((size_t)n);
  h2d(c->rt, h.p, hashes, sizeof(uint32_t) * (size_t)n);
  foreach(c->rt, n, LookupFn{idx->ix.table.p, idx->ix.tableMask, h.p, o.p});
  d2h(c->rt, counts, o.p, sizeof(int32_t) * (size_t)n);
  end_call(c);
  MM_CATCH
}
void mm_index_destroy(mm_index* idx) {
  if (!idx) return;
#ifndef MM_HOST_EMU
  cudaSetDevice(idx->ctx->rt.device);
#endif
  delete idx;
}

// ------------------------------------------------------------------------------------------------ K3-K5
static void check_map_args(mm_ctx* c, const mm_index* idx, const mm_map_params* p) {
  if (!c || !idx || !p) throw Error(MM_EINVAL, "mm_map_batch: bad arguments");
  // a finalized index is read-only: any context of the same device may map against it (two host threads with one
  // context each keep two batches in flight and fill each other's synchronisation bubbles)
  if (idx->ctx->rt.device != c->rt.device) throw Error(MM_EINVAL, "index lives on another device");
  if (!idx->ix.finalized) throw Error(MM_EINVAL, "index not finalized");
}
static void fill_summary(mm_map_summary* out, const int64_t* s) {
  if (out) { out->n_reads = s[0]; out->n_too_short = s[1]; out->n_candidates = s[2]; out->n_mappings = s[3]; out->n_reads_mapped = s[4]; out->total_bases_mapped_reads = s[5]; }
}
// --all absent: keep, per read, the mappings with identity >= best - 1.0 (reportReadMappings, computeMap.hpp:551-563); the
// others leave the accepted set, so every consumer (mm_map_fetch_*, the classify stage) sees what the reference would print
struct CandSketchFn { const int32_t* cRead; const int32_t* sOf; int32_t* out; MM_HD void operator()(int64_t c) const { out[c] = ldg(sOf + ldg(cRead + c)); } };
struct BestFilterFn {
  const int64_t* candOff; const float* identity; int32_t* oAccept;
  MM_HD void operator()(int64_t r) const {
    const int64_t b = ldg(candOff + r), e = ldg(candOff + r + 1);
    float best = 0;
    for (int64_t x = b; x < e; x++) if (oAccept[x] && ldg(identity + x) > best) best = ldg(identity + x);
    for (int64_t x = b; x < e; x++) if (oAccept[x] && !(ldg(identity + x) >= best - 1.0)) oAccept[x] = 0;
  }
};
static void apply_best_filter(mm_ctx* c, int64_t* s /*summary[6]*/) {
  Mapper& m = c->mp; Classifier& cl = c->cls; Runtime& rt = c->rt;
  const int64_t nc = m.n_cand;
  if (nc <= 0) return;
  auto& sk = c->scr.tmpA; sk.ensure((size_t)nc + 1);
  foreach(rt, nc, CandSketchFn{m.cRead.p, m.sOf.p, sk.p});
  cl.identity(m.oShared.p, sk.p, nc, m.lastK, m.oAccept.p);
  unsigned long long flagged = 0; d2h(rt, &flagged, cl.cnt.p, sizeof flagged);
  cl.identity_fixups(m.oShared.p, sk.p, nc, m.lastK, flagged, m.oAccept.p);
  foreach(rt, m.n_reads, BestFilterFn{m.candOff.p, cl.id32.p, m.oAccept.p});
  m.red.ensure(2);
  c->pr.reduce_sum<int32_t>(m.oAccept.p, m.red.p, nc);
  int32_t h = 0; d2h(rt, &h, m.red.p, sizeof h);
  s[3] = h; m.st.counters[4] = h;
}
static int map_impl(mm_ctx* c, const mm_index* idx, const char* reads, const void* dev, const int64_t* offsets, int32_t n,
                    const mm_map_params* p, mm_map_summary* out) {
  MM_TRY
  if (!offsets || n < 0) throw Error(MM_EINVAL, "mm_map_batch: bad arguments");
  check_map_args(c, idx, p);
  begin_call(c);
  c->sk.load(c->mp.batch, reads, dev, offsets, n);
  int64_t s[6];
  { StageTimer t(c->rt, &c->last_ms); c->mp.run(idx->ix, c->mp.batch, p->perc_identity, p->min_read_len, s); if (!p->report_all) apply_best_filter(c, s); }
  end_call(c);
  fill_summary(out, s);
  MM_CATCH
}
int mm_map_batch(mm_ctx* c, const mm_index* idx, const char* reads, const int64_t* offsets, int32_t n, const mm_map_params* p, mm_map_summary* out) {
  if (!reads && n > 0) { g_err = "mm_map_batch: null reads"; return MM_EINVAL; }
  return map_impl(c, idx, reads, nullptr, offsets, n, p, out);
}
int mm_map_batch_dev(mm_ctx* c, const mm_index* idx, const void* dev, const int64_t* offsets, int32_t n, const mm_map_params* p, mm_map_summary* out) {
  if (!dev) { g_err = "mm_map_batch_dev: null device pointer"; return MM_EINVAL; }
  return map_impl(c, idx, nullptr, dev, offsets, n, p, out);
}
// ---- contig-sharded ranks: sketch the own block once, all-gather the sketches, map everything against the own shard
struct QReadFillFn {        // qRead[e] = r for e in [qOff[r], qOff[r+1])
  const int64_t* qOff; int32_t* qRead;
  MM_HD void operator()(int64_t r) const { for (int64_t e = ldg(qOff + r), e1 = ldg(qOff + r + 1); e < e1; e++) qRead[e] = (int32_t)r; }
};
struct PackBytesFn { const uint8_t* in; uint32_t* out; int64_t n; MM_HD void operator()(int64_t w) const {
  uint32_t v = 0; for (int j = 0; j < 4; j++) { const int64_t i = 4 * w + j; if (i < n) v |= (uint32_t)ldg(in + i) << (8 * j); } out[w] = v; } };
struct UnpackBytesFn { const uint32_t* in; uint8_t* out; MM_HD void operator()(int64_t i) const { out[i] = (uint8_t)(ldg(in + (i >> 2)) >> (8 * (i & 3))); } };
static void allgather_u32(mm_ctx* c, const uint32_t* send, uint32_t* recv, size_t count);
int mm_map_batch_sharded_dev(mm_ctx* c, const mm_index* idx, const void* dev, const int64_t* offsets, int32_t n, const mm_map_params* p, mm_map_summary* out,
                             int64_t* first_read) {
  MM_TRY
  if (!offsets || n < 0 || (!dev && n > 0)) throw Error(MM_EINVAL, "mm_map_batch_sharded_dev: bad arguments");
  check_map_args(c, idx, p);
  begin_call(c);
  Runtime& rt = c->rt; Mapper& m = c->mp; const int R = c->nRanks;
  c->sk.load(m.batch, nullptr, dev, offsets, n);
  int64_t s[6];
  {
    StageTimer t(rt, &c->last_ms);
    m.sketch_reads(idx->ix.k, idx->ix.w, m.batch, p->min_read_len);
    if (R > 1) {
      m.join_sketch();                                           // the gathered strands must be final
      // 1. per-rank scalars
      auto& one = c->scr.gOne; auto& all = c->scr.gAll; one.ensure(8); all.ensure((size_t)8 * R);
      uint32_t h[8] = {(uint32_t)m.n_reads, (uint32_t)(m.n_q & 0xffffffffu), (uint32_t)((uint64_t)m.n_q >> 32), (uint32_t)m.maxSketch, (uint32_t)m.nShort_,
                       (uint32_t)((uint64_t)m.basesOk_ & 0xffffffffu), (uint32_t)((uint64_t)m.basesOk_ >> 32), (uint32_t)m.n_ambig};
      h2d(rt, one.p, h, sizeof h); allgather_u32(c, one.p, all.p, 8);
      std::vector<uint32_t> ha((size_t)8 * R); d2h(rt, ha.data(), all.p, 4 * ha.size());
      int64_t totReads = 0, totQ = 0, capN = 0, capQ = 0, nShort = 0, bases = 0, nAmb = 0; int32_t maxS = 0; int64_t myFirst = 0;
      std::vector<int64_t> nr((size_t)R), nq((size_t)R);
      for (int r = 0; r < R; r++) {
        nr[(size_t)r] = ha[(size_t)8 * r]; nq[(size_t)r] = (int64_t)(((uint64_t)ha[(size_t)8 * r + 2] << 32) | ha[(size_t)8 * r + 1]);
        if (r < c->rank) myFirst += nr[(size_t)r];
        totReads += nr[(size_t)r]; totQ += nq[(size_t)r]; capN = std::max(capN, nr[(size_t)r]); capQ = std::max(capQ, nq[(size_t)r]);
        maxS = std::max(maxS, (int32_t)ha[(size_t)8 * r + 3]); nShort += ha[(size_t)8 * r + 4];
        bases += (int64_t)(((uint64_t)ha[(size_t)8 * r + 6] << 32) | ha[(size_t)8 * r + 5]); nAmb += ha[(size_t)8 * r + 7];
      }
      if (totReads >= ((int64_t)1 << 31)) throw Error(MM_ERANGE, "more than 2^31 reads in one sharded batch");
      if (first_read) *first_read = myFirst;
      // 2. padded slabs: [readLen | sOf] (2 capN words) and [qHash | qStrand packed 4 per word] (capQ + capQ/4 + 1 words)
      const size_t wq = (size_t)capQ + (size_t)(capQ + 3) / 4;
      auto& sendA = c->scr.gSendA; auto& recvA = c->scr.gRecvA; auto& sendB = c->scr.gSendB; auto& recvB = c->scr.gRecvB;
      sendA.ensure((size_t)2 * capN + 1); recvA.ensure(((size_t)2 * capN + 1) * R); sendB.ensure(wq + 1); recvB.ensure((wq + 1) * R);
      dev_memset(rt, sendA.p, 0, 4 * ((size_t)2 * capN + 1)); dev_memset(rt, sendB.p, 0, 4 * (wq + 1));
      d2d(rt, sendA.p, m.readLen.p, 4 * (size_t)m.n_reads); d2d(rt, sendA.p + capN, m.sOf.p, 4 * (size_t)m.n_reads);
      d2d(rt, sendB.p, m.qHash.p, 4 * (size_t)m.n_q);
      if (m.n_q > 0) foreach(rt, (m.n_q + 3) / 4, PackBytesFn{m.qStrand.p, sendB.p + capQ, m.n_q});
      if (capN > 0) allgather_u32(c, sendA.p, recvA.p, (size_t)2 * capN);
      if (wq > 0) allgather_u32(c, sendB.p, recvB.p, wq);
      // 3. the batch's arrays, blocks in rank order
      m.readLen.ensure((size_t)totReads + 1); m.sOf.ensure((size_t)totReads + 1); m.qOff.ensure((size_t)totReads + 2);
      m.qHash.ensure((size_t)totQ + 1); m.qStrand.ensure((size_t)totQ + 4); m.qRead.ensure((size_t)totQ + 1);
      int64_t pr = 0, pq = 0;
      for (int r = 0; r < R; r++) {
        const uint32_t* a = recvA.p + (size_t)r * 2 * capN; const uint32_t* b = recvB.p + (size_t)r * wq;
        d2d(rt, m.readLen.p + pr, a, 4 * (size_t)nr[(size_t)r]); d2d(rt, m.sOf.p + pr, a + capN, 4 * (size_t)nr[(size_t)r]);
        d2d(rt, m.qHash.p + pq, b, 4 * (size_t)nq[(size_t)r]);
        if (nq[(size_t)r] > 0) foreach(rt, nq[(size_t)r], UnpackBytesFn{b + capQ, m.qStrand.p + pq});
        pr += nr[(size_t)r]; pq += nq[(size_t)r];
      }
      dev_memset(rt, m.sOf.p + totReads, 0, sizeof(int32_t));
      c->pr.exclusive_sum<int32_t, int64_t>(m.sOf.p, m.qOff.p, totReads + 1);
      if (totReads > 0) foreach(rt, totReads, QReadFillFn{m.qOff.p, m.qRead.p});
      m.n_reads = (int32_t)totReads; m.n_q = totQ; m.maxSketch = maxS; m.nShort_ = nShort; m.basesOk_ = bases; m.n_ambig = nAmb;
    } else if (first_read) *first_read = 0;
    m.map_sketched(idx->ix, p->perc_identity, s);
    if (!p->report_all) apply_best_filter(c, s);
  }
  end_call(c);
  fill_summary(out, s);
  MM_CATCH
}
int mm_stage_reads_async(mm_ctx* c, int slot, const char* reads, const int64_t* offsets, int32_t n) {
  MM_TRY
  if (!c || slot < 0 || slot > 1 || !offsets || n < 0 || (!reads && n > 0)) throw Error(MM_EINVAL, "mm_stage_reads_async: bad arguments");
  auto& st = c->stage[slot];
  const int64_t bytes = offsets[n] - offsets[0];
  if (bytes < 0) throw Error(MM_EINVAL, "mm_stage_reads_async: offsets not ascending");
  st.off.assign(offsets, offsets + n + 1);
  for (auto& o : st.off) o -= offsets[0];
  st.n = n;
#ifndef MM_HOST_EMU
  MM_CUDA(cudaSetDevice(c->rt.device));
  if (!st.ready) MM_CUDA(cudaEventCreateWithFlags(&st.ready, cudaEventDisableTiming));
  st.join();
  st.feedErr = 0; st.packed = false;
  // Default: DMA (copy engine) into a device slot, in pieces (below); K0 then packs from device memory at the start of the batch's
  // own step (0.7 ms).  MM_STAGE=zerocopy: for pinned (mapped) host memory K0 itself pulls the bytes over PCIe on the copy stream
  // while the previous batch is mapped -- no K0 on the critical path, but its CTAs sit on the SMs for the whole transfer and
  // slow K1 / K3 of the batch being mapped by 4 ms (config 2: e2e 17 614 Mbp/s against 18 590 with the DMA path, call T).
  cudaPointerAttributes attr; memset(&attr, 0, sizeof attr);
  const bool mapped = bytes > 0 && cudaPointerGetAttributes(&attr, reads + offsets[0]) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer;
  cudaGetLastError();
  const char* mode = getenv("MM_STAGE");
  if (mapped && mode && !strcmp(mode, "zerocopy")) {
    const uint8_t* dp = (const uint8_t*)attr.devicePointer - offsets[0];       // PackFn indexes with the caller's offsets
    std::swap(c->rt.stream, c->rt.copy);                                         // issue on the copy stream
    try {
      c->sk.prepare(st.batch, offsets, n);
      c->sk.pack_async(st.batch, dp, 1);                                         // 148 CTAs keep PCIe busy without crowding the SMs
    } catch (...) { std::swap(c->rt.stream, c->rt.copy); throw; }
    std::swap(c->rt.stream, c->rt.copy);
    st.packed = true; st.hostAsc = dp;
    MM_CUDA(cudaEventRecord(st.ready, c->rt.copy));
    return MM_OK;
  }
  st.asc.ensure((size_t)bytes + 16);
  // The DMA queue is first-in first-out across streams: if the whole batch were queued at once, the small host->device
  // copies of the batch being mapped meanwhile would wait behind 100s of MB.  A feeder thread therefore hands the copy
  // to the engine 8 MB at a time (one piece in flight), and records the slot's event when the last piece is done.
  st.join();
  st.feedErr = 0;
  {
    const int dev = c->rt.device; cudaStream_t cs = c->rt.copy; uint8_t* dst = st.asc.p; const char* src = reads + offsets[0];
    cudaEvent_t ev = st.ready; int* err = &st.feedErr;
    int64_t piece = (int64_t)8 << 20;
    if (const char* e = getenv("MM_STAGE_PIECE_KB")) { long long v = atoll(e); if (v >= 64) piece = v << 10; }
    st.feeder = std::thread([=]() {
      if (cudaSetDevice(dev) != cudaSuccess) { *err = 1; return; }
      for (int64_t o = 0; o < bytes; o += piece) {
        const int64_t nb = std::min<int64_t>(piece, bytes - o);
        if (cudaMemcpyAsync(dst + o, src + o, (size_t)nb, cudaMemcpyHostToDevice, cs) != cudaSuccess || cudaStreamSynchronize(cs) != cudaSuccess) { *err = 1; return; }
      }
      if (cudaEventRecord(ev, cs) != cudaSuccess) *err = 1;
    });
  }
#else
  st.asc.ensure((size_t)bytes + 16);
  if (bytes) memcpy(st.asc.p, reads + offsets[0], (size_t)bytes);
#endif
  MM_CATCH
}
int mm_map_batch_staged(mm_ctx* c, const mm_index* idx, int slot, const mm_map_params* p, mm_map_summary* out) {
  if (!c || slot < 0 || slot > 1 || c->stage[slot].n < 0) { g_err = "mm_map_batch_staged: nothing staged in this slot"; return MM_EINVAL; }
  auto& st = c->stage[slot];
#ifndef MM_HOST_EMU
  st.join();                                    // the last piece has been handed over and the event recorded
  if (st.feedErr) { g_err = "mm_stage_reads_async: the host->device copy failed"; return MM_ECUDA; }
  cudaSetDevice(c->rt.device);
  cudaError_t e = cudaStreamWaitEvent(c->rt.stream, st.ready, 0);
  if (e != cudaSuccess) { g_err = std::string("cudaStreamWaitEvent: ") + cudaGetErrorString(e); return MM_ECUDA; }
#endif
  if (st.packed) {
    MM_TRY
    check_map_args(c, idx, p);
    begin_call(c);
    c->sk.finish_pack(st.batch, st.hostAsc);                 // on the main stream, after the event
    int64_t s[6];
    { StageTimer t(c->rt, &c->last_ms); c->mp.run(idx->ix, st.batch, p->perc_identity, p->min_read_len, s); if (!p->report_all) apply_best_filter(c, s); }
    end_call(c);
    fill_summary(out, s);
    MM_CATCH
  }
  return map_impl(c, idx, nullptr, st.asc.p, st.off.data(), st.n, p, out);
}
struct MinHitsOfFn {
  const int32_t* sOf; const int32_t* tab; int32_t* out;
  MM_HD void operator()(int64_t r) const { int32_t s = ldg(sOf + r); out[r] = s > 0 ? ldg(tab + s) : 0; }
};
int mm_map_fetch_reads(mm_ctx* c, int32_t* sketch_size, int32_t* minimum_hits, int64_t* cand_offsets) {
  MM_TRY
  if (!c) throw Error(MM_EINVAL, "null ctx");
  Mapper& m = c->mp;
  begin_call(c);
  if (sketch_size) d2h(c->rt, sketch_size, m.sOf.p, sizeof(int32_t) * (size_t)m.n_reads);
  if (minimum_hits && m.n_reads) {
    auto& t = c->scr.tmpA; t.ensure((size_t)m.n_reads);
    foreach(c->rt, m.n_reads, MinHitsOfFn{m.sOf.p, m.dMinHits.p, t.p});
    d2h(c->rt, minimum_hits, t.p, sizeof(int32_t) * (size_t)m.n_reads);
  }
  if (cand_offsets) d2h(c->rt, cand_offsets, m.candOff.p, sizeof(int64_t) * ((size_t)m.n_reads + 1));
  c->rt.sync();
  MM_CATCH
}
int mm_map_fetch_candidates(mm_ctx* c, int32_t* seq_id, int32_t* range_start, int32_t* range_end, int32_t* pos, int32_t* shared,
                            int32_t* votes, int32_t* accepted, int32_t* valid, int64_t* opt_start, int64_t* opt_end) {
  MM_TRY
  if (!c) throw Error(MM_EINVAL, "null ctx");
  Mapper& m = c->mp; size_t n = (size_t)m.n_cand;
  begin_call(c);
  if (seq_id) d2h(c->rt, seq_id, m.cSeq.p, 4 * n);
  if (range_start) d2h(c->rt, range_start, m.cStart.p, 4 * n);
  if (range_end) d2h(c->rt, range_end, m.cEnd.p, 4 * n);
  if (pos) d2h(c->rt, pos, m.oPos.p, 4 * n);
  if (shared) d2h(c->rt, shared, m.oShared.p, 4 * n);
  if (votes) d2h(c->rt, votes, m.oVotes.p, 4 * n);
  if (accepted) d2h(c->rt, accepted, m.oAccept.p, 4 * n);
  if (valid) d2h(c->rt, valid, m.oValid.p, 4 * n);
  if (opt_start) d2h(c->rt, opt_start, m.oOptS.p, 8 * n);
  if (opt_end) d2h(c->rt, opt_end, m.oOptE.p, 8 * n);
  c->rt.sync();
  MM_CATCH
}
struct StrandOfFn { const uint8_t* q; int32_t* out; MM_HD void operator()(int64_t i) const { out[i] = ldg(q + i) ? 1 : -1; } };
int mm_map_fetch_sketch(mm_ctx* c, int64_t* offsets, uint32_t* hash, int32_t* strand, int64_t cap) {
  MM_TRY
  if (!c) throw Error(MM_EINVAL, "null ctx");
  Mapper& m = c->mp;
  begin_call(c);
  if (offsets) d2h(c->rt, offsets, m.qOff.p, sizeof(int64_t) * ((size_t)m.n_reads + 1));
  if (cap < m.n_q && (hash || strand)) throw Error(MM_ERANGE, "mm_map_fetch_sketch: capacity too small");
  if (hash) d2h(c->rt, hash, m.qHash.p, sizeof(uint32_t) * (size_t)m.n_q);
  if (strand && m.n_q) {
    DevBuf<int32_t> t; t.ensure((size_t)m.n_q);
    foreach(c->rt, m.n_q, StrandOfFn{m.qStrand.p, t.p});
    d2h(c->rt, strand, t.p, sizeof(int32_t) * (size_t)m.n_q);
  }
  c->rt.sync();
  MM_CATCH
}

struct MapCompactFn {
  const int32_t* oAccept; const int64_t* accIdx; const int32_t* cRead; const int32_t* cSeq; const int32_t* oPos; const int32_t* oShared;
  const int32_t* oVotes; const int32_t* sOf;
  int32_t* mRead; int32_t* mSeq; int32_t* mPos; int32_t* mShared; int32_t* mSketch; int32_t* mStrand;
  MM_HD void operator()(int64_t c) const {
    if (!ldg(oAccept + c)) return;
    const int64_t d = ldg(accIdx + c); const int32_t r = ldg(cRead + c);
    mRead[d] = r; mSeq[d] = ldg(cSeq + c); mPos[d] = ldg(oPos + c); mShared[d] = ldg(oShared + c); mSketch[d] = ldg(sOf + r);
    mStrand[d] = ldg(oVotes + c) > 0 ? 1 : -1;                         // computeMap.hpp:438
  }
};
// the double obtained by printing a float with 6 significant digits and parsing it back
static inline double round6(float x32) {
  const double x = (double)x32;
  if (x >= 10.0 && x < 99.99995) return rint(x * 1e4) / 1e4;       // x*1e4 < 2^24 * 2^14: exact in double
  char buf[64]; snprintf(buf, sizeof buf, "%.6g", x); return strtod(buf, nullptr);
}
int mm_group_sorted(const int32_t* v, int64_t n, int32_t* gv, int64_t* go, int64_t* ng) {
  MM_TRY
  if (n < 0 || (n > 0 && (!v || !gv || !go)) || !ng) throw Error(MM_EINVAL, "mm_group_sorted: bad arguments");
  int64_t g = 0;
  for (int64_t i = 0; i < n; i++) {
    if (i == 0 || v[i] != v[i - 1]) {
      if (i > 0 && v[i] < v[i - 1]) throw Error(MM_EINVAL, "mm_group_sorted: values are not sorted");
      gv[g] = v[i]; go[g] = i; g++;
    }
  }
  if (go) go[g] = n;
  *ng = g;
  MM_CATCH
}
int mm_stat_identity_batch(const int32_t* shared, const int32_t* sketch, int64_t n, int k, float* identity, double* parsed) {
  MM_TRY
  if (n < 0 || (n > 0 && (!shared || !sketch))) throw Error(MM_EINVAL, "mm_stat_identity_batch: bad arguments");
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    float a, b; stats::identity_only(shared[i], sketch[i], k, &a); (void)b;
    if (identity) identity[i] = a;
    if (parsed) parsed[i] = round6(a);
  }
  MM_CATCH
}
int mm_map_fetch_mappings(mm_ctx* c, int32_t* read_idx, int32_t* seq_id, int32_t* ref_start, int32_t* shared, int32_t* sketch, int32_t* strand,
                          float* identity, double* parsed, int64_t cap, int64_t* n_out) {
  MM_TRY
  if (!c) throw Error(MM_EINVAL, "null ctx");
  Mapper& m = c->mp; const int64_t nc = m.n_cand;
  begin_call(c);
  int64_t nm = 0;
  auto& idx = c->scr.dOff; auto& a = c->scr.tmpA; auto& b = c->scr.tmpB; auto& d0 = c->scr.dSh; auto& d1 = c->scr.dSk; auto& d2 = c->scr.dLen; auto& d3 = c->scr.dSt;
  if (nc > 0) {
    idx.ensure((size_t)nc + 2);
    dev_memset(c->rt, m.oAccept.p + nc, 0, sizeof(int32_t));
    c->pr.exclusive_sum<int32_t, int64_t>(m.oAccept.p, idx.p, nc + 1);
    d2h(c->rt, &nm, idx.p + nc, sizeof(int64_t));
  }
  if (n_out) *n_out = nm;
  if (nm > cap && (read_idx || seq_id || ref_start || shared || sketch || strand || identity || parsed)) throw Error(MM_ERANGE, "mm_map_fetch_mappings: capacity too small");
  if (nm > 0) {
    a.ensure((size_t)nm); b.ensure((size_t)nm); d0.ensure((size_t)nm); d1.ensure((size_t)nm); d2.ensure((size_t)nm); d3.ensure((size_t)nm);
    foreach(c->rt, nc, MapCompactFn{m.oAccept.p, idx.p, m.cRead.p, m.cSeq.p, m.oPos.p, m.oShared.p, m.oVotes.p, m.sOf.p, a.p, b.p, d0.p, d1.p, d2.p, d3.p});
    std::vector<int32_t> hs, hk;
    const bool needId = identity || parsed;
    if (needId && !shared) hs.resize((size_t)nm);
    if (needId && !sketch) hk.resize((size_t)nm);
    int32_t* pShared = shared ? shared : (needId ? hs.data() : nullptr);
    int32_t* pSketch = sketch ? sketch : (needId ? hk.data() : nullptr);
    if (read_idx) d2h(c->rt, read_idx, a.p, 4 * (size_t)nm);
    if (seq_id) d2h(c->rt, seq_id, b.p, 4 * (size_t)nm);
    if (ref_start) d2h(c->rt, ref_start, d0.p, 4 * (size_t)nm);
    if (pShared) d2h(c->rt, pShared, d1.p, 4 * (size_t)nm);
    if (pSketch) d2h(c->rt, pSketch, d2.p, 4 * (size_t)nm);
    if (strand) d2h(c->rt, strand, d3.p, 4 * (size_t)nm);
    c->rt.sync();
    if (needId) { int rc = mm_stat_identity_batch(pShared, pSketch, nm, m.lastK, identity, parsed); if (rc != MM_OK) return rc; }
  }
  end_call(c);
  MM_CATCH
}
int mm_nloc_batch(const int32_t* seq_id, const int64_t* read_off, const int32_t* read_len, int64_t n_reads, const int64_t* contig_len,
                  const int32_t* contig_taxon, int32_t n_contigs, int32_t T, int32_t* taxon, double* nloc) {
  MM_TRY
  if (n_reads < 0 || !read_off || !contig_len || !contig_taxon || n_contigs < 0 || T < 1) throw Error(MM_EINVAL, "mm_nloc_batch: bad arguments");
  // contigs of each taxon sorted by length, with prefix sums of the lengths
  std::vector<int64_t> start((size_t)T + 1, 0);
  for (int32_t c_ = 0; c_ < n_contigs; c_++) {
    if (contig_taxon[c_] < 0 || contig_taxon[c_] >= T) throw Error(MM_EINVAL, "mm_nloc_batch: contig taxon out of range");
    start[(size_t)contig_taxon[c_] + 1]++;
  }
  for (int32_t t = 0; t < T; t++) start[(size_t)t + 1] += start[(size_t)t];
  std::vector<int64_t> lens((size_t)n_contigs), fill(start.begin(), start.end() - 1), csum((size_t)n_contigs + 1, 0);
  for (int32_t c_ = 0; c_ < n_contigs; c_++) lens[(size_t)fill[(size_t)contig_taxon[c_]]++] = contig_len[c_];
  for (int32_t t = 0; t < T; t++) std::sort(lens.begin() + start[(size_t)t], lens.begin() + start[(size_t)t + 1]);
  for (int32_t c_ = 0; c_ < n_contigs; c_++) csum[(size_t)c_ + 1] = csum[(size_t)c_] + lens[(size_t)c_];
  const int64_t M = read_off[n_reads];
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < n_reads; r++) {
    const int64_t L = read_len[r];
    const int64_t m0 = read_off[r], m1 = read_off[r + 1];
    for (int64_t m = m0; m < m1; m++) {
      const int32_t sq = seq_id[m];
      if (sq < 0 || sq >= n_contigs) { bad = 1; continue; }
      const int32_t t = contig_taxon[sq];
      if (taxon) taxon[m] = t;
      if (!nloc) continue;
      const int64_t* b = lens.data() + start[(size_t)t]; const int64_t* e = lens.data() + start[(size_t)t + 1];
      const int64_t* p = std::lower_bound(b, e, L);                       // contigs at least as long as the read
      const int64_t nBig = e - p;
      int64_t v = (csum[(size_t)(e - lens.data())] - csum[(size_t)(p - lens.data())]) - nBig * (L - 1);
      if (p != b) {           // shorter contigs of the taxon count once each if this read maps to them (fEM.h:337-345)
        for (int64_t x = m0; x < m1; x++) {
          const int32_t sx = seq_id[x];
          if (sx < 0 || sx >= n_contigs || contig_taxon[sx] != t || contig_len[sx] >= L) continue;
          bool first = true;
          for (int64_t y = m0; y < x; y++) if (seq_id[y] == sx) { first = false; break; }
          if (first) v++;
        }
      }
      nloc[m] = (double)v;
    }
  }
  (void)M;
  if (bad) throw Error(MM_EINVAL, "mm_nloc_batch: contig id out of range");
  MM_CATCH
}

// ------------------------------------------------------------------------------------------------ host statistics
int mm_stat_min_hits_relaxed(int s, int k, float pi) { return stats::estimateMinimumHitsRelaxed(s, k, pi); }
int mm_stat_recommended_window(double p, int k, int alphabet, float pi, int lenQ, uint64_t lenR) { return stats::recommendedWindowSize(p, k, alphabet, pi, lenQ, lenR); }
double mm_stat_estimate_pvalue(int s, int k, int alphabet, float pi, int lenQ, uint64_t lenR) { return stats::estimate_pvalue(s, k, alphabet, pi, lenQ, lenR); }
void mm_stat_identity(int shared, int s, int k, float* a, float* b) { stats::identity(shared, s, k, a, b); }

// ------------------------------------------------------------------------------------------------ K6
int mm_mapq_batch(mm_ctx* c, const double* identity, const int32_t* shared, const int32_t* sketch, const int32_t* read_len,
                  const int64_t* read_off, int64_t n_reads, int k, double* mapq, int32_t* status) {
  MM_TRY
  if (!c || !read_off || n_reads < 0) throw Error(MM_EINVAL, "mm_mapq_batch: bad arguments");
  begin_call(c);
  int64_t M = read_off[n_reads];
  Classifier& cl = c->cls;
  auto& dId = c->scr.dId; auto& dSh = c->scr.dSh; auto& dSk = c->scr.dSk;
  dId.ensure((size_t)M + 1); dSh.ensure((size_t)M + 1); dSk.ensure((size_t)M + 1);
  cl.grpLen.ensure((size_t)n_reads + 1); cl.grpOff.ensure((size_t)n_reads + 2);
  h2d(c->rt, dId.p, identity, 8 * (size_t)M); h2d(c->rt, dSh.p, shared, 4 * (size_t)M); h2d(c->rt, dSk.p, sketch, 4 * (size_t)M);
  h2d(c->rt, cl.grpLen.p, read_len, 4 * (size_t)n_reads); h2d(c->rt, cl.grpOff.p, read_off, 8 * ((size_t)n_reads + 1));
  cl.nGroups = n_reads;
  { StageTimer t(c->rt, &c->last_ms); cl.run_mapq(dId.p, 1.0, dSh.p, dSk.p, M, k); }
  d2h(c->rt, mapq, cl.mapq.p, 8 * (size_t)M);
  if (status) d2h(c->rt, status, cl.status.p, 4 * (size_t)n_reads);
  end_call(c);
  MM_CATCH
}

// ------------------------------------------------------------------------------------------------ NCCL binding
struct NcclUid { char b[128]; };
typedef int (*fn_ncclGetUniqueId)(NcclUid*);
typedef int (*fn_ncclCommInitRank)(void**, int, NcclUid, int);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_ncclCommDestroy)(void*);
typedef int (*fn_ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_ncclSend)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_ncclGroup)(void);
// libnccl is opened ONCE per process and its entry points resolved once (a magic static: thread-safe); the handle is never
// closed -- communicators of any context may still be alive at exit.
struct NcclApi {
  void* lib = nullptr; std::string err;
  fn_ncclGetUniqueId getUniqueId = nullptr; fn_ncclCommInitRank commInitRank = nullptr; fn_ncclAllReduce allReduce = nullptr;
  fn_ncclCommDestroy commDestroy = nullptr; fn_ncclAllGather allGather = nullptr;
  NcclApi() {
#ifdef MM_HOST_EMU
    err = "NCCL is not available in the host-emulation test build";
#else
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    auto sym = [&](const char* n) -> void* { void* p = dlsym(lib, n); if (!p && err.empty()) err = std::string("libnccl is missing ") + n; return p; };
    getUniqueId = (fn_ncclGetUniqueId)sym("ncclGetUniqueId"); commInitRank = (fn_ncclCommInitRank)sym("ncclCommInitRank");
    allReduce = (fn_ncclAllReduce)sym("ncclAllReduce"); commDestroy = (fn_ncclCommDestroy)sym("ncclCommDestroy"); allGather = (fn_ncclAllGather)sym("ncclAllGather");
#endif
  }
};
static const NcclApi& nccl() {
  static NcclApi api;
  if (!api.err.empty()) throw Error(MM_ENODEV, api.err);
  return api;
}
int mm_comm_unique_id(void* id) {
  MM_TRY
  if (!id) throw Error(MM_EINVAL, "null id buffer");
  int rc = nccl().getUniqueId((NcclUid*)id);
  if (rc != 0) throw Error(MM_ECUDA, "ncclGetUniqueId failed: " + std::to_string(rc));
  MM_CATCH
}
int mm_comm_init(mm_ctx* c, int n_ranks, int rank, const void* id) {
  MM_TRY
  if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) throw Error(MM_EINVAL, "mm_comm_init: bad arguments");
  if (c->comm) throw Error(MM_EINVAL, "mm_comm_init: the context already has a communicator (mm_comm_destroy first)");
  begin_call(c);
  NcclUid u; memcpy(&u, id, sizeof(u));
  int rc = nccl().commInitRank(&c->comm, n_ranks, u, rank);
  if (rc != 0) { c->comm = nullptr; throw Error(MM_ECUDA, "ncclCommInitRank failed: " + std::to_string(rc)); }
  c->nRanks = n_ranks; c->rank = rank;
  MM_CATCH
}
int mm_comm_destroy(mm_ctx* c) {
  if (!c) return MM_EINVAL;
  if (c->comm) {
    try { nccl().commDestroy(c->comm); } catch (...) {}
  }
  c->comm = nullptr; c->nRanks = 1; c->rank = 0;
  return MM_OK;
}
int mm_comm_set_allreduce(mm_ctx* c, mm_allreduce_fn fn, void* user) {
  if (!c) { g_err = "null ctx"; return MM_EINVAL; }
  c->hostAllreduce = fn; c->hostAllreduceUser = user;
  return MM_OK;
}
static void allreduce_sum_f64(mm_ctx* c, double* buf, size_t n) {
  if (!c->comm && c->hostAllreduce) {        // host-staged transport
    c->hostBuf.resize(n);
    d2h(c->rt, c->hostBuf.data(), buf, sizeof(double) * n);
    if (c->hostAllreduce(c->hostBuf.data(), (int64_t)n, c->hostAllreduceUser) != 0) throw Error(MM_ECUDA, "host all-reduce callback failed");
    h2d(c->rt, buf, c->hostBuf.data(), sizeof(double) * n);
    return;
  }
  if (!c->comm || c->nRanks == 1) return;
  int rc = nccl().allReduce(buf, buf, n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->comm, c->rt.stream);
  if (rc != 0) throw Error(MM_ECUDA, "ncclAllReduce failed: " + std::to_string(rc));
  c->rt.launches++;
}

int mm_comm_set_rank(mm_ctx* c, int n_ranks, int rank) {
  if (!c || n_ranks < 1 || rank < 0 || rank >= n_ranks) { g_err = "mm_comm_set_rank: bad arguments"; return MM_EINVAL; }
  if (c->comm) { g_err = "mm_comm_set_rank: the context has an NCCL communicator"; return MM_EINVAL; }
  c->nRanks = n_ranks; c->rank = rank;
  return MM_OK;
}
// all-gather of `count` 32-bit words per rank (device buffers); recv holds nRanks * count words in rank order
static void allgather_u32(mm_ctx* c, const uint32_t* send, uint32_t* recv, size_t count) {
  if (count == 0) return;
  if (c->comm && c->nRanks > 1) {
    int rc = nccl().allGather(send, recv, count, /*ncclUint32*/ 3, c->comm, c->rt.stream);
    if (rc != 0) throw Error(MM_ECUDA, "ncclAllGather failed: " + std::to_string(rc));
    c->rt.launches++;
    return;
  }
  if (c->nRanks > 1 && c->hostAllreduce) {      // host transport: every rank fills its own slot of a zero buffer, then sum
    std::vector<uint32_t> mine(count);
    d2h(c->rt, mine.data(), send, 4 * count);
    c->hostBuf.assign(count * (size_t)c->nRanks, 0.0);
    for (size_t i = 0; i < count; i++) c->hostBuf[(size_t)c->rank * count + i] = (double)mine[i];
    if (c->hostAllreduce(c->hostBuf.data(), (int64_t)c->hostBuf.size(), c->hostAllreduceUser) != 0) throw Error(MM_ECUDA, "host all-reduce callback failed");
    std::vector<uint32_t> all(c->hostBuf.size());
    for (size_t i = 0; i < all.size(); i++) all[i] = (uint32_t)c->hostBuf[i];
    h2d(c->rt, recv, all.data(), 4 * all.size());
    c->rt.sync();
    return;
  }
  if (c->nRanks > 1) throw Error(MM_EINVAL, "multi-rank context without a transport (mm_comm_init or mm_comm_set_allreduce)");
  d2d(c->rt, recv, send, 4 * count);
}

int mm_index_set_shard(mm_index* idx, int32_t first_contig_id, int32_t keep_counts) {
  if (!idx || first_contig_id < 0) { g_err = "mm_index_set_shard: bad arguments"; return MM_EINVAL; }
  if (idx->ix.finalized) { g_err = "mm_index_set_shard: call before mm_index_finalize"; return MM_EINVAL; }
  idx->ix.firstContig = first_contig_id; idx->ix.keepUnique = keep_counts != 0;
  return MM_OK;
}
int mm_index_set_freq_carry(mm_index* idx, const int32_t* v, const int64_t* c, int32_t n, int32_t prev_threshold) {
  if (!idx || n < 0 || (n > 0 && (!v || !c))) { g_err = "mm_index_set_freq_carry: bad arguments"; return MM_EINVAL; }
  if (idx->ix.globalSynced) { g_err = "mm_index_set_freq_carry: the index already carries the threshold of the unchunked reference (mm_index_sync_threshold)"; return MM_EINVAL; }
  idx->ix.carryHist.clear();
  for (int32_t i = 0; i < n; i++) idx->ix.carryHist.emplace_back((uint32_t)v[i], c[i]);
  idx->ix.carryThreshold = prev_threshold;
  if (idx->ix.finalized) idx->ix.apply_carry();       // the threshold is applied at probe time: it may be settled after the build
  return MM_OK;
}
int mm_index_get_freq_hist(const mm_index* idx, int32_t* v, int64_t* c, int32_t cap, int32_t* n, int32_t* threshold) {
  if (!idx || !idx->ix.finalized || !n) { g_err = "mm_index_get_freq_hist: index not finalized"; return MM_EINVAL; }
  const auto& h = idx->ix.histCum;
  *n = (int32_t)h.size();
  for (int32_t i = 0; i < (int32_t)h.size() && i < cap; i++) { if (v) v[i] = (int32_t)h[(size_t)i].first; if (c) c[i] = h[(size_t)i].second; }
  if (threshold) *threshold = idx->ix.freqThreshold;
  return MM_OK;
}
int mm_index_sync_threshold(mm_index* idx, int32_t* global_threshold, int64_t* global_unique) {
  MM_TRY
  if (!idx || !idx->ix.finalized) throw Error(MM_EINVAL, "mm_index_sync_threshold: index not finalized");
  Index& ix = idx->ix; mm_ctx* c = idx->ctx; Runtime& rt = c->rt; Prims& pr = c->pr;
  if (!ix.keepUnique) throw Error(MM_EINVAL, "mm_index_sync_threshold: mm_index_set_shard(.., keep_counts = 1) was not called before finalize");
  begin_call(c);
  const int R = c->nRanks; const int64_t nu = ix.n_unique;
  int64_t STEP = (int64_t)1 << 23;
  if (const char* e = getenv("MM_SYNC_STEP")) { long long v = atoll(e); if (v >= 4) STEP = v; }
  // 1. hash ranges: every STEP-th unique hash of rank 0 (the shards are statistically alike)
  DevBuf<uint32_t> one, oneAll; one.ensure(1); oneAll.ensure((size_t)R);
  uint32_t nRmine = (uint32_t)std::max<int64_t>(1, (nu + STEP - 1) / STEP);
  h2d(rt, one.p, &nRmine, 4); allgather_u32(c, one.p, oneAll.p, 1);
  uint32_t nR = 0; d2h(rt, &nR, oneAll.p, 4);                               // rank 0's value
  DevBuf<uint32_t> bnd, bndAll; bnd.ensure(nR + 1); bndAll.ensure((size_t)R * (nR + 1));
  dev_memset(rt, bnd.p, 0, 4 * ((size_t)nR + 1));
  if (c->rank == 0 && nR > 1) foreach(rt, (int64_t)nR - 1, GatherStrideFn{ix.uHash.p + STEP, STEP, bnd.p + 1});
  allgather_u32(c, bnd.p, bndAll.p, nR + 1);                                // slot 0 = rank 0's boundaries; [0] = 0
  // 2. local offsets of the ranges
  DevBuf<int64_t> dLo; dLo.ensure((size_t)nR + 1);
  foreach(rt, (int64_t)nR, LowerBoundFn{ix.uHash.p, nu, bndAll.p, dLo.p});
  std::vector<int64_t> lo((size_t)nR + 1);
  d2h(rt, lo.data(), dLo.p, 8 * (size_t)nR); lo[0] = 0; lo[nR] = nu;
  // 3. every rank's entry count per range
  std::vector<uint32_t> cntMine(nR);
  for (uint32_t j = 0; j < nR; j++) cntMine[j] = (uint32_t)(lo[j + 1] - lo[j]);
  DevBuf<uint32_t> dCnt, dCntAll; dCnt.ensure(nR); dCntAll.ensure((size_t)R * nR);
  h2d(rt, dCnt.p, cntMine.data(), 4 * (size_t)nR); allgather_u32(c, dCnt.p, dCntAll.p, nR);
  std::vector<uint32_t> cntAll((size_t)R * nR); d2h(rt, cntAll.data(), dCntAll.p, 4 * cntAll.size());
  // 4. per range: gather, merge, histogram, look the local hashes up
  DevBuf<uint32_t> gLocal; gLocal.ensure((size_t)nu + 1);
  DevBuf<uint32_t> sendH, sendC, recvH, recvC, catH, catC, srtH, srtC, uh, gc, gcs, hv; DevBuf<int32_t> runLen, hc; DevBuf<int64_t> runStart, nRuns;
  nRuns.ensure(2);
  std::vector<std::pair<uint32_t, int64_t>> hist;          // (count value, number of hashes), unsorted, merged below
  int64_t uniqueGlobal = 0;
  for (uint32_t j = 0; j < nR; j++) {
    uint32_t maxc = 0; int64_t tot = 0;
    for (int r = 0; r < R; r++) { maxc = std::max(maxc, cntAll[(size_t)r * nR + j]); tot += cntAll[(size_t)r * nR + j]; }
    if (tot == 0) continue;
    sendH.ensure(maxc); sendC.ensure(maxc); recvH.ensure((size_t)R * maxc); recvC.ensure((size_t)R * maxc);
    dev_memset(rt, sendH.p, 0, 4 * (size_t)maxc); dev_memset(rt, sendC.p, 0, 4 * (size_t)maxc);
    d2d(rt, sendH.p, ix.uHash.p + lo[j], 4 * (size_t)cntMine[j]); d2d(rt, sendC.p, ix.uCnt.p + lo[j], 4 * (size_t)cntMine[j]);
    allgather_u32(c, sendH.p, recvH.p, maxc); allgather_u32(c, sendC.p, recvC.p, maxc);
    catH.ensure((size_t)tot); catC.ensure((size_t)tot); srtH.ensure((size_t)tot); srtC.ensure((size_t)tot);
    int64_t pos = 0;
    for (int r = 0; r < R; r++) {
      const size_t n_ = cntAll[(size_t)r * nR + j];
      d2d(rt, catH.p + pos, recvH.p + (size_t)r * maxc, 4 * n_); d2d(rt, catC.p + pos, recvC.p + (size_t)r * maxc, 4 * n_);
      pos += (int64_t)n_;
    }
    pr.sort_pairs<uint32_t, uint32_t>(catH.p, srtH.p, catC.p, srtC.p, tot);
    uh.ensure((size_t)tot); runLen.ensure((size_t)tot + 1); runStart.ensure((size_t)tot + 1); gc.ensure((size_t)tot); gcs.ensure((size_t)tot);
    pr.rle<uint32_t>(srtH.p, uh.p, runLen.p, nRuns.p, tot);
    int64_t nuj = 0; d2h(rt, &nuj, nRuns.p, 8);
    pr.exclusive_sum<int32_t, int64_t>(runLen.p, runStart.p, nuj);
    foreach(rt, nuj, RunSumFn{srtC.p, runStart.p, runLen.p, gc.p});
    uniqueGlobal += nuj;
    // histogram of the global counts of this range
    pr.sort_keys<uint32_t>(gc.p, gcs.p, nuj);
    hv.ensure((size_t)nuj); hc.ensure((size_t)nuj + 1);
    pr.rle<uint32_t>(gcs.p, hv.p, hc.p, nRuns.p, nuj);
    int64_t nb = 0; d2h(rt, &nb, nRuns.p, 8);
    std::vector<uint32_t> v((size_t)nb); std::vector<int32_t> k_((size_t)nb);
    d2h(rt, v.data(), hv.p, 4 * (size_t)nb); d2h(rt, k_.data(), hc.p, 4 * (size_t)nb);
    for (int64_t b = 0; b < nb; b++) hist.emplace_back(v[(size_t)b], (int64_t)k_[(size_t)b]);
    if (cntMine[j]) foreach(rt, (int64_t)cntMine[j], GlobalCountFn{ix.uHash.p, lo[j], uh.p, gc.p, nuj, gLocal.p});
    rt.sync();
  }
  // 5. the threshold of the whole reference (computeFreqHist, winSketch.hpp:452-495) and the local flags
  std::sort(hist.begin(), hist.end());
  std::vector<std::pair<uint32_t, int64_t>> hm;
  for (auto& p : hist) { if (!hm.empty() && hm.back().first == p.first) hm.back().second += p.second; else hm.push_back(p); }
  float percentageThreshold = 0.001f;
  int64_t toIgnore = (int64_t)(uniqueGlobal * percentageThreshold / 100);
  int64_t sum = 0; int32_t T = 0x7fffffff;
  for (int64_t b = (int64_t)hm.size() - 1; b >= 0; b--) {
    sum += hm[(size_t)b].second;
    if (sum < toIgnore) T = (int32_t)hm[(size_t)b].first;
    else if (sum == toIgnore) { T = (int32_t)hm[(size_t)b].first; break; }
    else break;
  }
  if (nu > 0 && T != 0x7fffffff) foreach(rt, nu, FlagFreqFn{ix.table.p, ix.tableMask, ix.uHash.p, gLocal.p, (uint32_t)T});
  ix.freqThreshold = 0x7fffffff;            // from now on only the flags decide (flagged counts compare above everything)
  ix.globalSynced = true; ix.globalThreshold = T;
  ix.uHash.release(); ix.uCnt.release();
  end_call(c);
  if (global_threshold) *global_threshold = T;
  if (global_unique) *global_unique = uniqueGlobal;
  MM_CATCH
}

// ------------------------------------------------------------------------------------------------ K7/K8
static void bind_allreduce(mm_ctx* c) {
  const bool multi = (c->comm && c->nRanks > 1) || c->hostAllreduce;
  if (multi) c->cls.allreduce = [c](double* buf, size_t n) { allreduce_sum_f64(c, buf, n); };
  else c->cls.allreduce = nullptr;
  c->cls.hostTransport = !c->comm && c->hostAllreduce;
}
int mm_em_run(mm_ctx* c, const int32_t* taxon, const double* mapq, const double* nloc, const int64_t* read_off, int64_t n_reads,
              int32_t T, int32_t max_iter, double* f_out, double* posterior, int64_t* best, double* ll_hist, int32_t ll_cap, int32_t* n_iter) {
  MM_TRY
  if (!c || !read_off || n_reads < 0 || T < 1) throw Error(MM_EINVAL, "mm_em_run: bad arguments");
  begin_call(c);
  Runtime& rt = c->rt; Classifier& cl = c->cls;
  int64_t M = read_off[n_reads];
  if (M >= ((int64_t)1 << 31)) throw Error(MM_ERANGE, "more than 2^31 mappings on one rank: partition the reads");
  auto& dTax = c->scr.dTax; auto& dMq = c->scr.dMq; auto& dNl = c->scr.dNl; auto& dOff = c->scr.dOff;
  dTax.ensure((size_t)M + 1); dMq.ensure((size_t)M + 1); dNl.ensure((size_t)M + 1); dOff.ensure((size_t)n_reads + 2); cl.w.ensure((size_t)M + 1);
  h2d(rt, dTax.p, taxon, 4 * (size_t)M); h2d(rt, dMq.p, mapq, 8 * (size_t)M); h2d(rt, dNl.p, nloc, 8 * (size_t)M);
  h2d(rt, dOff.p, read_off, 8 * ((size_t)n_reads + 1));
  for (int64_t m = 0; m < M; m++) if (taxon[m] < 0 || taxon[m] >= T) throw Error(MM_EINVAL, "mm_em_run: taxon out of range");
  bind_allreduce(c);
  cl.emMs = 0;
  foreach(rt, M, EmWeightFn{dMq.p, dNl.p, cl.w.p});
  cl.run_em(dTax.p, dOff.p, n_reads, M, T, max_iter, ll_cap > 0 ? ll_cap : 0);
  if (f_out) d2h(rt, f_out, cl.f.p, 8 * (size_t)T);
  if (posterior) d2h(rt, posterior, cl.post.p, 8 * (size_t)M);
  if (best) d2h(rt, best, cl.best.p, 8 * (size_t)n_reads);
  if (ll_hist && ll_cap > 0) d2h(rt, ll_hist, cl.llHist.p, 8 * (size_t)std::min<int32_t>(cl.iters, ll_cap));
  if (n_iter) *n_iter = cl.iters;
  end_call(c);
  c->last_ms += cl.emMs;
  MM_CATCH
}

// ------------------------------------------------------------------------------------------------ classify stage on the device
struct ReadRangeFn {        // first mapping with read >= lo / >= hi in the read-sorted table
  const int32_t* mRead; int64_t n; int32_t lo, hi; int64_t* out;
  MM_HD void operator()(int64_t i) const {
    const int32_t key = i == 0 ? lo : hi;
    int64_t a = 0, b = n;
    while (a < b) { const int64_t m = (a + b) >> 1; if (ldg(mRead + m) < key) a = m + 1; else b = m; }
    out[i] = a;
  }
};
// table -> sorted by read, stable (a read's mappings keep their part order = contig order).  bySeq: the parts did not arrive in
// contig order (ranks that own interleaved chunks of the reference): order by (read, contig) with two stable passes -- mappings on
// the same contig come from one part and keep their (position) order
static void maptable_sort(mm_ctx* c, int32_t n_reads_hint, bool bySeq = false) {
  Classifier& cl = c->cls; MapTable& t = cl.tab; Runtime& rt = c->rt;
  if (t.n <= 1) { t.sorted = true; return; }
  cl.keyA.ensure((size_t)t.n); cl.keyB.ensure((size_t)t.n); cl.permA.ensure((size_t)t.n); cl.permB.ensure((size_t)t.n); cl.tmpI.ensure((size_t)t.n);
  DevBuf<int32_t>* cols[6] = {&t.read, &t.seq, &t.pos, &t.shared, &t.sketch, &t.strand};
  auto pass = [&](const int32_t* key, int bits) {
    foreach(rt, t.n, IotaU32Fn{cl.permA.p});
    c->pr.sort_pairs<uint32_t, uint32_t>((const uint32_t*)key, cl.keyB.p, cl.permA.p, cl.permB.p, t.n, bits);
    for (auto* col : cols) {
      foreach(rt, t.n, GatherI32Fn{cl.permB.p, col->p, cl.tmpI.p});
      d2d(rt, col->p, cl.tmpI.p, 4 * (size_t)t.n);
    }
  };
  if (bySeq) pass(t.seq.p, 31);
  int bits = 1; while (bits < 32 && ((int64_t)1 << bits) <= (int64_t)(n_reads_hint > 1 ? n_reads_hint : 2)) bits++;
  pass(t.read.p, bits);
  t.sorted = true;
}
int mm_classify_setup(mm_ctx* c, const int64_t* contig_len, const int32_t* contig_taxon, int32_t n_contigs, int32_t T) {
  MM_TRY
  if (!c || !contig_len || !contig_taxon || n_contigs < 0 || T < 1) throw Error(MM_EINVAL, "mm_classify_setup: bad arguments");
  begin_call(c);
  c->cls.taxo.upload(c->rt, contig_len, contig_taxon, n_contigs, T);
  MM_CATCH
}
int mm_classify_begin(mm_ctx* c) {
  if (!c) { g_err = "null ctx"; return MM_EINVAL; }
  c->cls.tab.n = 0; c->cls.tab.parts = 0; c->cls.tab.sorted = true; c->cls.readBase = 0; c->cls.readsSeen = 0;
  return MM_OK;
}
int mm_classify_next_batch(mm_ctx* c) {
  if (!c) { g_err = "null ctx"; return MM_EINVAL; }
  c->cls.readBase = c->cls.readsSeen; c->cls.tab.parts = 0;
  return MM_OK;
}
int mm_classify_add_mappings(mm_ctx* c, int32_t first_contig_id, int64_t* n_total) {
  MM_TRY
  if (!c || first_contig_id < 0) throw Error(MM_EINVAL, "mm_classify_add_mappings: bad arguments");
  Mapper& m = c->mp; MapTable& t = c->cls.tab; const int64_t nc = m.n_cand;
  begin_call(c);
  int64_t nm = 0;
  auto& idx = c->scr.dOff;
  if (nc > 0) {
    idx.ensure((size_t)nc + 2);
    dev_memset(c->rt, m.oAccept.p + nc, 0, sizeof(int32_t));
    c->pr.exclusive_sum<int32_t, int64_t>(m.oAccept.p, idx.p, nc + 1);
    d2h(c->rt, &nm, idx.p + nc, sizeof(int64_t));
  }
  if (t.n + nm >= ((int64_t)1 << 31)) throw Error(MM_ERANGE, "more than 2^31 mappings in one classify table: use smaller read batches");
  Classifier& cl = c->cls;
  if (cl.readBase + m.n_reads >= ((int64_t)1 << 31)) throw Error(MM_ERANGE, "more than 2^31 reads in one classify table");
  // the batch's read lengths, at its place in the table's read numbering
  cl.readLenAll.grow(c->rt, (size_t)(cl.readBase + m.n_reads) + 1, (size_t)cl.readsSeen);
  d2d(c->rt, cl.readLenAll.p + cl.readBase, m.readLen.p, 4 * (size_t)m.n_reads);
  cl.readsSeen = std::max<int64_t>(cl.readsSeen, cl.readBase + m.n_reads);
  if (nm > 0) {
    t.reserve(c->rt, t.n + nm);
    StageTimer tm(c->rt, &c->last_ms);
    foreach(c->rt, nc, MapAppendFn{m.oAccept.p, idx.p, m.cRead.p, m.cSeq.p, m.oPos.p, m.oShared.p, m.oVotes.p, m.sOf.p, first_contig_id, (int32_t)cl.readBase, t.n,
                                   t.read.p, t.seq.p, t.pos.p, t.shared.p, t.sketch.p, t.strand.p});
  }
  if (t.parts > 0 && nm > 0 && t.n > 0) t.sorted = false;      // a second part: reads interleave, sort before use
  t.n += nm; t.parts++;
  end_call(c);
  if (n_total) *n_total = t.n;
  MM_CATCH
}
int mm_classify_exchange(mm_ctx* c, int32_t read_lo, int32_t read_hi, int64_t* n_total) {
  MM_TRY
  if (!c || read_lo < 0 || read_hi < read_lo) throw Error(MM_EINVAL, "mm_classify_exchange: bad arguments");
  begin_call(c);
  Runtime& rt = c->rt; Classifier& cl = c->cls; MapTable& t = cl.tab;
  const int R = c->nRanks;
  StageTimer tm(rt, &c->last_ms);
  if (!t.sorted) maptable_sort(c, (int32_t)cl.readsSeen);
  // 1. every rank's count
  auto& one = c->scr.gOne; auto& all = c->scr.gAll; one.ensure(8); all.ensure((size_t)8 * R);
  uint32_t mine = (uint32_t)t.n; h2d(rt, one.p, &mine, 4);
  allgather_u32(c, one.p, all.p, 1);
  std::vector<uint32_t> cnt((size_t)R); d2h(rt, cnt.data(), all.p, 4 * (size_t)R);
  uint32_t cap = 0; int64_t tot = 0; for (int r = 0; r < R; r++) { cap = std::max(cap, cnt[(size_t)r]); tot += cnt[(size_t)r]; }
  if (tot >= ((int64_t)1 << 31)) throw Error(MM_ERANGE, "more than 2^31 mappings in one exchange: use smaller read batches");
  if (cap > 0) {
    // 2. one padded slab of 6 columns per rank, all-gathered
    auto& send = c->scr.gSendA; auto& recv = c->scr.gRecvA; send.ensure((size_t)6 * cap); recv.ensure((size_t)R * 6 * cap);
    DevBuf<int32_t>* cols[6] = {&t.read, &t.seq, &t.pos, &t.shared, &t.sketch, &t.strand};
    dev_memset(rt, send.p, 0, 4 * (size_t)6 * cap);
    for (int a = 0; a < 6; a++) d2d(rt, send.p + (size_t)a * cap, cols[a]->p, 4 * (size_t)t.n);
    allgather_u32(c, send.p, recv.p, (size_t)6 * cap);
    // 3. concatenate in rank (= shard) order, then one stable sort by read
    t.n = 0; t.reserve(rt, tot);
    int64_t pos = 0;
    for (int r = 0; r < R; r++) {
      for (int a = 0; a < 6; a++) d2d(rt, cols[a]->p + pos, recv.p + ((size_t)r * 6 + a) * cap, 4 * (size_t)cnt[(size_t)r]);
      pos += cnt[(size_t)r];
    }
    t.n = tot; t.sorted = false;
    maptable_sort(c, (int32_t)cl.readsSeen, true);      // ranks may own interleaved contig ranges: order by (read, contig)
    // 4. this rank finalises the reads [read_lo, read_hi)
    auto& rg = c->scr.gRange; rg.ensure(2);
    foreach(rt, 2, ReadRangeFn{t.read.p, t.n, read_lo, read_hi, rg.p});
    int64_t h[2]; d2h(rt, h, rg.p, sizeof h);
    const int64_t keep = h[1] - h[0];
    if (h[0] > 0 && keep > 0) {
      cl.tmpI.ensure((size_t)keep);
      for (int a = 0; a < 6; a++) { d2d(rt, cl.tmpI.p, cols[a]->p + h[0], 4 * (size_t)keep); d2d(rt, cols[a]->p, cl.tmpI.p, 4 * (size_t)keep); }
    }
    t.n = keep;
  }
  tm.stop();
  end_call(c);
  if (n_total) *n_total = t.n;
  MM_CATCH
}
int mm_classify_run(mm_ctx* c, int32_t em_max_iter, mm_classify_summary* out) {
  MM_TRY
  if (!c) throw Error(MM_EINVAL, "null ctx");
  Classifier& cl = c->cls; MapTable& t = cl.tab; Runtime& rt = c->rt; Mapper& mp = c->mp;
  if (em_max_iter >= 0 && !cl.taxo.set) throw Error(MM_EINVAL, "mm_classify_run: call mm_classify_setup first (only em_max_iter < 0 -- identity and mapping quality alone -- works without a taxonomy)");
  begin_call(c);
  bind_allreduce(c);
  const int64_t M = t.n; const int k = mp.lastK; const int32_t T = cl.taxo.set ? cl.taxo.T : 1;
  cl.emMs = 0; cl.iters = 0; cl.nGroups = 0; cl.nFix = 0;
  {
    StageTimer tm(rt, &c->last_ms);
    if (!t.sorted) maptable_sort(c, (int32_t)cl.readsSeen);
    if (M > 0) {
      cl.identity(t.shared.p, t.sketch.p, M, k);
      cl.build_groups(t.read.p, M);                                       // host sync: number of mapped reads
      unsigned long long flagged = 0; d2h(rt, &flagged, cl.cnt.p, sizeof flagged);
      cl.identity_fixups(t.shared.p, t.sketch.p, M, k, flagged);
      foreach(rt, cl.nGroups, GroupLenFn{cl.grpRead.p, cl.readLenAll.p, cl.grpLen.p});
      cl.run_mapq(cl.parsed.p, 100.0, t.shared.p, t.sketch.p, M, k);
      cl.tax.ensure((size_t)M + 1); cl.nloc.ensure((size_t)M + 1); cl.w.ensure((size_t)M + 1); cl.bad.ensure(1);
      dev_memset(rt, cl.bad.p, 0, sizeof(int32_t));
      if (cl.taxo.set) foreach(rt, M, NlocFn{t.seq.p, cl.mGrp.p, cl.grpOff.p, cl.grpLen.p, cl.mapq.p, cl.taxo.contigLen.p, cl.taxo.contigTaxon.p, cl.taxo.nContigs,
                            cl.taxo.lens.p, cl.taxo.start.p, cl.taxo.csum.p, cl.tax.p, cl.nloc.p, cl.w.p, cl.bad.p});
    } else { cl.grpOff.ensure(2); dev_memset(rt, cl.grpOff.p, 0, 16); cl.w.ensure(1); cl.tax.ensure(1); }
  }
  if (M > 0) { int32_t bad = 0; d2h(rt, &bad, cl.bad.p, sizeof bad); if (bad) throw Error(MM_EINVAL, "mm_classify_run: contig id outside the taxonomy given to mm_classify_setup"); }
  if (em_max_iter >= 0 && (M > 0 || cl.allreduce)) cl.run_em(cl.tax.p, cl.grpOff.p, cl.nGroups, M, T, em_max_iter, 4096);
  end_call(c);
  c->last_ms += cl.emMs;
  if (out) { out->n_mappings = M; out->n_reads_mapped = cl.nGroups; out->em_iters = cl.iters; out->n_identity_fixups = (int32_t)cl.nFix; out->em_ms = cl.emMs; out->classify_ms = c->last_ms; }
  MM_CATCH
}
int mm_classify_fetch(mm_ctx* c, int32_t* read_idx, int32_t* seq_id, int32_t* ref_start, int32_t* shared, int32_t* sketch, int32_t* strand,
                      float* identity, double* parsed, double* mapq, int32_t* taxon, double* nloc, double* posterior, int64_t cap,
                      int32_t* mapped_read, int64_t* read_off, int64_t* best, int32_t* status, double* f, double* ll_hist, int32_t ll_cap) {
  MM_TRY
  if (!c) throw Error(MM_EINVAL, "null ctx");
  Classifier& cl = c->cls; MapTable& t = cl.tab; Runtime& rt = c->rt;
  const size_t M = (size_t)t.n, G = (size_t)cl.nGroups;
  if ((int64_t)M > cap && (read_idx || seq_id || ref_start || shared || sketch || strand || identity || parsed || mapq || taxon || nloc || posterior))
    throw Error(MM_ERANGE, "mm_classify_fetch: capacity too small");
  begin_call(c);
#ifndef MM_HOST_EMU
  auto get = [&](void* h, const void* d, size_t bytes) { if (h && bytes) MM_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, rt.stream)); };
#else
  auto get = [&](void* h, const void* d, size_t bytes) { if (h && bytes) memcpy(h, d, bytes); };
#endif
  get(read_idx, t.read.p, 4 * M); get(seq_id, t.seq.p, 4 * M); get(ref_start, t.pos.p, 4 * M); get(shared, t.shared.p, 4 * M);
  get(sketch, t.sketch.p, 4 * M); get(strand, t.strand.p, 4 * M);
  if (M) { get(identity, cl.id32.p, 4 * M); get(parsed, cl.parsed.p, 8 * M); get(mapq, cl.mapq.p, 8 * M); if (cl.taxo.set) { get(taxon, cl.tax.p, 4 * M); get(nloc, cl.nloc.p, 8 * M); } }
  if (M && cl.iters > 0) get(posterior, cl.post.p, 8 * M);
  if (G) { get(mapped_read, cl.grpRead.p, 4 * G); get(status, cl.status.p, 4 * G); if (cl.iters > 0) get(best, cl.best.p, 8 * G); }
  if (read_off) { if (G) get(read_off, cl.grpOff.p, 8 * (G + 1)); else read_off[0] = 0; }
  if (cl.iters > 0 && cl.lastT > 0) { get(f, cl.f.p, 8 * (size_t)cl.lastT); if (ll_cap > 0) get(ll_hist, cl.llHist.p, 8 * (size_t)std::min<int32_t>(cl.iters, std::min<int32_t>(ll_cap, 4096))); }
  end_call(c);
  MM_CATCH
}

#ifdef MM_HOST_EMU
// test hook (host-emulation build only): the std::sort replay, checked against the real std::sort by tests
void mm_emu_stdsort(uint64_t* a, int64_t n) { mm::stdsort::sort(a, n); }
void mm_emu_prune_shift(int upper, int lower) { mm::g_emu_prune_shift[0] = upper; mm::g_emu_prune_shift[1] = lower; }
long long mm_emu_rebuild_calls(int reset) { long long v = mm::g_emu_rebuild_calls; if (reset) mm::g_emu_rebuild_calls = 0; return v; }
long long mm_emu_rebuild_elems(int reset) { long long v = mm::g_emu_rebuild_elems; if (reset) mm::g_emu_rebuild_elems = 0; return v; }
long long mm_emu_sweep_iters(int reset) { long long v = mm::g_emu_sweep_iters; if (reset) mm::g_emu_sweep_iters = 0; return v; }
#endif

}  // extern "C"
