// mm_platform.h -- device plumbing: memory, launches, timing, error handling.
//
// Every kernel in this library is written as a functor with `MM_HD void operator()(int64_t item)`
// (or a block-cooperative __global__ kernel in the .cu files) and launched through mm::foreach().
// Defining MM_HOST_EMU compiles the same functors for the host and runs them in a loop; that build
// (tests/_emu) exists ONLY so the kernel logic can be checked against the oracle on a machine without
// a GPU.  It is never shipped, never loaded by the product, and is not a fallback: the product library
// is the nvcc build and mm_ctx_create fails without a CUDA device.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <list>
#include <map>
#include <mutex>
#include <set>

#ifdef MM_HOST_EMU
#define MM_HD inline
#define MM_DEV inline
typedef int cudaStream_t;
struct uint2 { unsigned int x, y; };
struct uint4 { unsigned int x, y, z, w; };
inline uint2 make_uint2(unsigned int x, unsigned int y) { uint2 v; v.x = x; v.y = y; return v; }
#else
#include <cuda_runtime.h>
#define MM_HD __host__ __device__ __forceinline__
#define MM_DEV __device__ __forceinline__
#endif

namespace mm {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#ifndef MM_HOST_EMU
#define MM_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw mm::Error(-5, std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + __FILE__ + \
                              ":" + std::to_string(__LINE__));                                     \
  } while (0)
#endif

// ---- per-context launch bookkeeping -------------------------------------------------------------
struct Runtime {
  int device = 0;
  cudaStream_t stream = 0;
  cudaStream_t side = 0;         // side stream: work that is off the critical path (joined with events)
  cudaStream_t copy = 0;         // copy stream: input staging for the NEXT batch (mm_stage_reads_async)
  int sm_count = 148;
  int64_t launches = 0;          // kernels launched since reset()
  double total_ms = 0;           // filled by StageTimer users
  // cudaFuncSetAttribute is per DEVICE: the opt-ins are remembered per context (one device, one host thread at a time),
  // not in function-local statics -- a process may hold contexts on several GPUs
  std::set<const void*> attrDone;
  bool first(const void* key) { return attrDone.insert(key).second; }
#ifndef MM_HOST_EMU
  struct Pending { cudaEvent_t a = nullptr, b = nullptr; double* acc = nullptr; bool closed = false; };
  std::list<Pending> pending;     // list: StageTimer keeps a pointer to its entry
#endif
  void sync() {
#ifndef MM_HOST_EMU
    MM_CUDA(cudaStreamSynchronize(stream));
#endif
  }
  // call after sync(): folds every closed StageTimer into its accumulator (open ones are left alone)
  void resolve_timers() {
#ifndef MM_HOST_EMU
    for (auto it = pending.begin(); it != pending.end();) {
      if (!it->closed) { ++it; continue; }
      float ms = 0;
      cudaError_t e = cudaEventElapsedTime(&ms, it->a, it->b);
      if (e == cudaSuccess) { if (it->acc) *it->acc += ms; }
      else { fprintf(stderr, "[metamaps_b200] cudaEventElapsedTime failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); }
      cudaEventDestroy(it->a); cudaEventDestroy(it->b);
      it = pending.erase(it);
    }
#endif
  }
};

// ---- memory --------------------------------------------------------------------------------------
#ifndef MM_HOST_EMU
// Freed device blocks of 1 MB and more are kept (per device, up to MM_ALLOC_CACHE_GB, default 16) and handed out again to requests
// of about their size: cudaMalloc / cudaFree of GB-sized blocks cost milliseconds each with 100 GB in use, which is what a
// chunk-streamed reference (one index built and freed per chunk) and growing scratch buffers spend their time on otherwise
// (config-5 slice: 1.3-3.2 s per step from run to run with the same 0.5 s of kernels).  A block is cached only after a device
// synchronisation (what cudaFree does implicitly), so no kernel of any stream can still be using it when it is handed out again.
struct DevCache {
  struct Dev { std::multimap<size_t, void*> free; size_t cached = 0; };
  std::mutex m;
  std::map<int, Dev> dev;
  std::map<void*, std::pair<int, size_t>> live;         // block -> (device, size) of every block this allocator handed out
  size_t cap;
  DevCache() { const char* e = getenv("MM_ALLOC_CACHE_GB"); cap = (size_t)((e ? atof(e) : 16.0) * (double)(1ull << 30)); }
  static DevCache& get() { static DevCache* c = new DevCache(); return *c; }       // never destroyed: the CUDA context may be gone at exit
  static size_t round_up(size_t b) {                    // sixteenths of the power of two below the size (<= 6 % slack), 2 MB at least
    size_t p2 = 1; while ((p2 << 1) <= b) p2 <<= 1;
    size_t q = p2 >> 4; if (q < ((size_t)2 << 20)) q = (size_t)2 << 20;
    return (b + q - 1) / q * q;
  }
  void flush(Dev& d) { for (auto& kv : d.free) cudaFree(kv.second); d.free.clear(); d.cached = 0; cudaGetLastError(); }
};
#endif
inline void* dev_alloc(size_t bytes) {
  if (bytes == 0) bytes = 16;
#ifdef MM_HOST_EMU
  void* p = malloc(bytes);
  if (!p) throw Error(-12, "host-emu alloc failed");
  return p;
#else
  DevCache& C = DevCache::get();
  int dv = 0; cudaGetDevice(&dv);
  const bool big = bytes >= ((size_t)1 << 20) && C.cap > 0;
  const size_t want = big ? DevCache::round_up(bytes) : bytes;
  std::lock_guard<std::mutex> lk(C.m);
  DevCache::Dev& d = C.dev[dv];
  if (big) {
    auto it = d.free.lower_bound(want);
    if (it != d.free.end() && it->first <= want + want / 4) {
      void* p = it->second; const size_t sz = it->first;
      d.free.erase(it); d.cached -= sz;
      C.live[p] = std::make_pair(dv, sz);
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess && d.cached > 0) { cudaGetLastError(); C.flush(d); e = cudaMalloc(&p, want); }
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error(-12, "cudaMalloc(" + std::to_string(want) + " B) failed: " + cudaGetErrorString(e));
  }
  C.live[p] = std::make_pair(dv, want);
  return p;
#endif
}
inline void dev_free(void* p) {
  if (!p) return;
#ifdef MM_HOST_EMU
  free(p);
#else
  DevCache& C = DevCache::get();
  std::unique_lock<std::mutex> lk(C.m);
  auto it = C.live.find(p);
  if (it != C.live.end()) {
    const int dv = it->second.first; const size_t sz = it->second.second;
    C.live.erase(it);
    DevCache::Dev& d = C.dev[dv];
    if (sz >= ((size_t)1 << 20) && d.cached + sz <= C.cap) {
      lk.unlock();
      int cur = 0; cudaGetDevice(&cur);
      if (cur != dv) cudaSetDevice(dv);
      const cudaError_t es = cudaDeviceSynchronize();   // as cudaFree would: nothing in flight may still touch the block
      if (cur != dv) cudaSetDevice(cur);
      if (es == cudaSuccess) {
        lk.lock();
        d.free.emplace(sz, p); d.cached += sz;
        return;
      }
      cudaGetLastError();
      lk.lock();
    }
  }
  lk.unlock();
  cudaError_t e = cudaFree(p);
  if (e != cudaSuccess) { fprintf(stderr, "[metamaps_b200] cudaFree(%p) failed: %s\n", p, cudaGetErrorString(e)); cudaGetLastError(); }
#endif
}
inline void h2d(Runtime& rt, void* d, const void* h, size_t bytes) {
  if (!bytes) return;
#ifdef MM_HOST_EMU
  memcpy(d, h, bytes);
#else
  MM_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, rt.stream));
#endif
}
inline void d2h(Runtime& rt, void* h, const void* d, size_t bytes) {
  if (!bytes) return;
#ifdef MM_HOST_EMU
  memcpy(h, d, bytes);
#else
  MM_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, rt.stream));
  MM_CUDA(cudaStreamSynchronize(rt.stream));
#endif
}
inline void d2d(Runtime& rt, void* dst, const void* src, size_t bytes) {
  if (!bytes) return;
#ifdef MM_HOST_EMU
  memmove(dst, src, bytes);
#else
  MM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, rt.stream));
#endif
}
inline void dev_memset(Runtime& rt, void* d, int v, size_t bytes) {
  if (!bytes) return;
#ifdef MM_HOST_EMU
  memset(d, v, bytes);
#else
  MM_CUDA(cudaMemsetAsync(d, v, bytes, rt.stream));
#endif
}

// Growable device array.  Not copyable; freed on destruction.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;   // elements
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { dev_free(p); }
  void release() { dev_free(p); p = nullptr; cap = 0; }
  // contents are NOT preserved
  T* ensure(size_t n) {
    if (n > cap) {
      dev_free(p); p = nullptr; cap = 0;
      size_t want = n + n / 8 + 64;
      p = (T*)dev_alloc(want * sizeof(T));
      cap = want;
    }
    return p;
  }
  // contents [0,keep) preserved
  T* grow(Runtime& rt, size_t n, size_t keep) {
    if (n > cap) {
      size_t want = n + n / 2 + 64;
      T* q = (T*)dev_alloc(want * sizeof(T));
      if (keep) d2d(rt, q, p, keep * sizeof(T));
#ifndef MM_HOST_EMU
      if (keep) MM_CUDA(cudaStreamSynchronize(rt.stream));
#endif
      dev_free(p);
      p = q; cap = want;
    }
    return p;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

// ---- launches ------------------------------------------------------------------------------------
#ifndef MM_HOST_EMU
template <class F>
__global__ void __launch_bounds__(256) foreach_kernel(int64_t n, F f) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
template <class F>
__global__ void __launch_bounds__(128) foreach_kernel128(int64_t n, F f) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
#endif

// One logical thread per item.  Grid = a multiple of the SM count (grid-stride loop), so the last wave
// is never a sliver (148 SMs, see DESIGN.md "grid sizing").
template <class F>
inline void foreach(Runtime& rt, int64_t n, const F& f, int block = 256, int ctas_per_sm = 8, bool on_side_stream = false) {
  if (n <= 0) return;
#ifdef MM_HOST_EMU
  (void)on_side_stream;
  for (int64_t i = 0; i < n; i++) f(i);
  rt.launches++;
#else
  int64_t need = (n + block - 1) / block;
  int64_t maxg = (int64_t)rt.sm_count * ctas_per_sm;
  int grid = (int)(need < maxg ? need : maxg);
  cudaStream_t st = on_side_stream ? rt.side : rt.stream;
  if (block == 128) foreach_kernel128<F><<<grid, 128, 0, st>>>(n, f);
  else foreach_kernel<F><<<grid, 256, 0, st>>>(n, f);
  MM_CUDA(cudaGetLastError());
  rt.launches++;
#endif
}

// ---- atomics / intrinsics usable from functors -----------------------------------------------------
template <class T>
MM_HD T atomic_add(T* p, T v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  T o = *p; *p = o + v; return o;
#endif
}
MM_HD unsigned long long atomic_add_u64(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  unsigned long long o = *p; *p = o + v; return o;
#endif
}
MM_HD uint32_t atomic_cas_u32(uint32_t* p, uint32_t cmp, uint32_t val) {
#if defined(__CUDA_ARCH__)
  return atomicCAS(p, cmp, val);
#else
  uint32_t o = *p; if (o == cmp) *p = val; return o;
#endif
}
MM_HD uint32_t atomic_or_u32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  return atomicOr(p, v);
#else
  uint32_t o = *p; *p = o | v; return o;
#endif
}
template <class T>
MM_HD T ldg(const T* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// ---- stage timing ----------------------------------------------------------------------------------
// Events are recorded on the context stream without synchronising; Runtime::resolve_timers() (called once
// the entry point has synchronised anyway) turns them into milliseconds.
struct StageTimer {
  Runtime& rt;
#ifndef MM_HOST_EMU
  Runtime::Pending* slot = nullptr;
#endif
  StageTimer(Runtime& r, double* accum) : rt(r) {
#ifndef MM_HOST_EMU
    rt.pending.emplace_back();
    Runtime::Pending& p = rt.pending.back();
    p.acc = accum;
    MM_CUDA(cudaEventCreate(&p.a)); MM_CUDA(cudaEventCreate(&p.b));
    MM_CUDA(cudaEventRecord(p.a, rt.stream));
    slot = &p;
#else
    (void)accum;
#endif
  }
  void stop() {
#ifndef MM_HOST_EMU
    if (!slot) return;
    Runtime::Pending* p = slot; slot = nullptr;
    MM_CUDA(cudaEventRecord(p->b, rt.stream));
    p->closed = true;
#endif
  }
  ~StageTimer() {
#ifndef MM_HOST_EMU
    try { stop(); } catch (...) {}
#endif
  }
};

}  // namespace mm
