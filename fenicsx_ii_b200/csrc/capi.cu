// extern "C" boundary (include/fxii.h).  No exceptions cross it: every entry point catches,
// records a thread-local message and returns a negative status.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace fxii {
void mesh_build(fxii_mesh* m, const double* x, int64_t n_nodes, const int32_t* cell_nodes, int64_t n_cells, int npc,
                double padding, cudaStream_t s);
void device_iota(int32_t* a, int64_t n, cudaStream_t s);
void check_index_range(const int32_t* a, int64_t n, int64_t hi, cudaStream_t s, const char* what);
void pi_assemble(const fxii_space* sp, fxii_rows* rows, cudaStream_t s, fxii_csr* out, fxii_pi_stats* stats,
                 bool keep_points);
void csr_transpose(const fxii_csr* a, int64_t c0, int64_t c1, cudaStream_t s, fxii_csr* out);
void spgemm(const fxii_csr* a, const fxii_csr* b, const double* b_scale, int64_t row_begin, int64_t row_end, cudaStream_t s,
            fxii_csr* out, int64_t* ip_out);
void csr_validate(const fxii_csr* a, cudaStream_t s);
void csr_scale_rows(fxii_csr* a, const double* d_host, cudaStream_t s);
void csr_expand_blocks(const fxii_csr* a, int bs, cudaStream_t s, fxii_csr* out);
void csr_spmv(const fxii_csr* a, const double* u_host, double* y_host, cudaStream_t s);
}  // namespace fxii

static thread_local std::string g_last_error;
#include <chrono>
#include <cstdlib>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>
namespace fxii {
int64_t g_launches = 0;
double Trace::now() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
Trace::Trace(const char* w, cudaStream_t st) : s(st), what(w) {
  static const bool enabled = getenv("FXII_TRACE") != nullptr;
  on = enabled;
  if (on) {
    cudaStreamSynchronize(s);
    t0 = now();
  }
}
void Trace::mark(const char* label) {
  if (!on) return;
  cudaStreamSynchronize(s);
  const double t = now();
  fprintf(stderr, "[fxii %s] %-28s %9.3f ms\n", what, label, t - t0);
  t0 = t;
}

// ---- pools: pinned result slots, timing events ---------------------------------------------------------
namespace {
std::mutex g_pool_mutex;
std::vector<void*> g_free_slots;
std::vector<std::pair<int, cudaEvent_t>> g_free_events;   // (device, event)
std::vector<std::pair<cudaEvent_t, int>> g_event_device;  // every event ever created -> its device
}  // namespace

void* pinned_slot_acquire() {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (g_free_slots.empty()) {
    unsigned char* page = nullptr;
    FXII_CUDA(cudaMallocHost((void**)&page, 4096));  // never freed: slots live as long as the process
    for (int i = 0; i < 64; ++i) g_free_slots.push_back(page + 64 * i);
  }
  void* p = g_free_slots.back();
  g_free_slots.pop_back();
  memset(p, 0, 64);
  return p;
}

void pinned_slot_release(void* p) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  g_free_slots.push_back(p);
}

// events belong to the device that was current when they were created: the pool is keyed by device
cudaEvent_t event_acquire() {
  int dev = 0;
  FXII_CUDA(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (size_t i = g_free_events.size(); i-- > 0;) {
      if (g_free_events[i].first == dev) {
        cudaEvent_t e = g_free_events[i].second;
        g_free_events.erase(g_free_events.begin() + (long)i);
        return e;
      }
    }
  }
  cudaEvent_t e = nullptr;
  FXII_CUDA(cudaEventCreate(&e));
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  g_event_device.push_back({e, dev});
  return e;
}

void event_release(cudaEvent_t e) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  int dev = 0;
  for (const auto& ed : g_event_device)
    if (ed.first == e) dev = ed.second;
  g_free_events.push_back({dev, e});
}

// ---- size-class caching allocator -----------------------------------------------------------------
namespace {
struct Block {
  void* p;
  cudaStream_t last;
};
std::mutex g_mem_mutex;
std::map<std::pair<int, size_t>, std::vector<Block>> g_free;  // (device, size class) -> free blocks
std::unordered_map<void*, std::pair<int, size_t>> g_live;       // live block -> (device, size class)
size_t g_cached_bytes = 0;

size_t size_class(size_t bytes) {
  if (bytes <= 512) return 512;
  // round up to a multiple of 2^(floor(log2)-3): at most 12.5 % slack, 8 classes per octave
  int lg = 63 - __builtin_clzll((unsigned long long)bytes);
  const size_t step = (size_t)1 << (lg - 3);
  return (bytes + step - 1) / step * step;
}
}  // namespace

void* cached_alloc(size_t bytes, cudaStream_t s) {
  const size_t cls = size_class(bytes);
  int dev = 0;
  cudaGetDevice(&dev);
  const std::pair<int, size_t> key{dev, cls};
  {
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    auto it = g_free.find(key);
    if (it != g_free.end() && !it->second.empty()) {
      Block b = it->second.back();
      it->second.pop_back();
      g_cached_bytes -= cls;
      g_live[b.p] = key;
      if (b.last != s) cudaStreamSynchronize(b.last);  // cross-stream reuse: wait for the old stream
      return b.p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, cls);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cache_trim();  // give cached blocks back and retry once
    e = cudaMalloc(&p, cls);
  }
  if (e != cudaSuccess)
    throw Error(FXII_E_CUDA, std::string("cudaMalloc of ") + std::to_string(cls) + " bytes: " + cudaGetErrorString(e));
  std::lock_guard<std::mutex> lock(g_mem_mutex);
  g_live[p] = key;
  return p;
}

void cached_free(void* p, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_mem_mutex);
  auto it = g_live.find(p);
  if (it == g_live.end()) return;
  const std::pair<int, size_t> key = it->second;
  g_live.erase(it);
  g_free[key].push_back(Block{p, s});
  g_cached_bytes += key.second;
}

void cache_trim() {
  std::lock_guard<std::mutex> lock(g_mem_mutex);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : g_free) {
    cudaSetDevice(kv.first.first);
    cudaDeviceSynchronize();
    for (auto& b : kv.second) cudaFree(b.p);
  }
  cudaSetDevice(cur);
  g_free.clear();
  g_cached_bytes = 0;
}
}  // namespace fxii

namespace fxii {
// ---- small device -> host reads through a mapped pinned page (common.cuh: d2h_small / stream_sync) -----------------
namespace {
constexpr size_t kFetchPage = 4096;
struct FetchState {
  unsigned char* page = nullptr;  // pinned; under unified addressing the device stores through the same pointer
  size_t used = 0;
  struct Item { void* host; size_t off, bytes; };
  std::vector<Item> items;
};
thread_local FetchState g_fetch;
__global__ void fetch_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, int bytes) {
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
}  // namespace

void d2h_small(void* host, const void* dev, size_t bytes, cudaStream_t s) {
  FetchState& f = g_fetch;
  if (!f.page) FXII_CUDA(cudaHostAlloc((void**)&f.page, kFetchPage, cudaHostAllocPortable | cudaHostAllocMapped));
  const size_t need = (bytes + 15) / 16 * 16;
  if (bytes > kFetchPage || f.used + need > kFetchPage) {  // does not happen on the path; keep the semantics anyway
    FXII_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s));
    return;
  }
  fetch_kernel<<<1, 64, 0, s>>>(f.page + f.used, static_cast<const unsigned char*>(dev), (int)bytes);
  FXII_LAUNCHED(1);
  FXII_CUDA(cudaGetLastError());
  f.items.push_back({host, f.used, bytes});
  f.used += need;
}

// an error unwound the caller before its stream_sync: forget the pending destinations (they were its locals)
void d2h_small_abandon() {
  g_fetch.items.clear();
  g_fetch.used = 0;
}

void stream_sync(cudaStream_t s) {
  FXII_CUDA(cudaStreamSynchronize(s));
  FetchState& f = g_fetch;
  for (const auto& it : f.items) memcpy(it.host, f.page + it.off, it.bytes);
  f.items.clear();
  f.used = 0;
}
}  // namespace fxii

namespace fxii {
// ---- pipelined upload of large pageable host arrays ----------------------------------------------------------
// A first call on numpy arrays uploads 3.6 GB for config 5's Ω (coordinates, cells, dofmap).  cudaMemcpyAsync from
// pageable memory goes through the driver's single-threaded staging copy: 11 GB/s measured on the B200 boxes, while
// the link moves 55 GB/s from pinned memory and eight host threads copy 69 GB/s into pinned memory
// (profiles/r02_pcie_probe.txt).  Here T worker threads each own two pinned 4 MiB buffers and a stream: a worker
// copies its next chunk into one buffer while the copy engine drains the other.  The buffers live for the process.
namespace {
constexpr size_t kStageChunk = (size_t)4 << 20;
constexpr size_t kStageMin = (size_t)32 << 20;  // below this the plain copy is as fast
constexpr int kStageMaxThreads = 8;
struct StageWorker {
  unsigned char* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
};
std::mutex g_stage_mutex;
std::map<int, std::vector<StageWorker>> g_stage;  // per device
int stage_threads() {
  static int n = [] {
    if (const char* e = getenv("FXII_UPLOAD_THREADS")) return std::max(0, std::min(atoi(e), kStageMaxThreads));
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min<unsigned>(hc / 2, kStageMaxThreads));
  }();
  return n;
}
}  // namespace

void staged_upload(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return;
  const int T = stage_threads();
  bool plain = bytes < kStageMin || T == 0;
  if (!plain) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, src) != cudaSuccess) {
      cudaGetLastError();
      plain = true;
    } else if (attr.type != cudaMemoryTypeUnregistered) {
      plain = true;  // pinned, registered, managed or device memory: the copy engine reads it directly
    }
  }
  if (plain) {
    FXII_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s));
    return;
  }
  int dev = 0;
  FXII_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_stage_mutex);
  std::vector<StageWorker>& ws = g_stage[dev];
  if ((int)ws.size() < T) {
    const size_t have = ws.size();
    ws.resize(T);
    for (size_t t = have; t < ws.size(); ++t) {
      FXII_CUDA(cudaStreamCreateWithFlags(&ws[t].stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        FXII_CUDA(cudaHostAlloc((void**)&ws[t].buf[b], kStageChunk, cudaHostAllocDefault));
        FXII_CUDA(cudaEventCreateWithFlags(&ws[t].ev[b], cudaEventDisableTiming));
      }
    }
  }
  // the destination may still be in use by work queued on s (a recycled block): the worker streams start after it
  cudaEvent_t start = event_acquire();
  FXII_CUDA(cudaEventRecord(start, s));
  const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
  std::vector<cudaError_t> err(T, cudaSuccess);
  std::vector<std::thread> threads;
  threads.reserve(T);
  for (int t = 0; t < T; ++t) {
    threads.emplace_back([&, t] {
      StageWorker& w = ws[t];
      cudaError_t e = cudaSetDevice(dev);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(w.stream, start, 0);
      int b = 0;
      for (size_t c = (size_t)t; c < n_chunks && e == cudaSuccess; c += (size_t)T, b ^= 1) {
        const size_t off = c * kStageChunk;
        const size_t n = std::min(kStageChunk, bytes - off);
        e = cudaEventSynchronize(w.ev[b]);  // the previous copy out of this buffer has finished (no-op the first time)
        if (e != cudaSuccess) break;
        memcpy(w.buf[b], static_cast<const unsigned char*>(src) + off, n);
        e = cudaMemcpyAsync(static_cast<unsigned char*>(dst) + off, w.buf[b], n, cudaMemcpyHostToDevice, w.stream);
        if (e == cudaSuccess) e = cudaEventRecord(w.ev[b], w.stream);
      }
      err[t] = e;
    });
  }
  for (auto& th : threads) th.join();  // the host array has been consumed; the last device copies may be in flight
  event_release(start);
  for (int t = 0; t < T; ++t) {
    // s continues after every worker's last copy (both buffers' events cover them)
    for (int b = 0; b < 2; ++b) FXII_CUDA(cudaStreamWaitEvent(s, ws[t].ev[b], 0));
  }
  for (int t = 0; t < T; ++t)
    if (err[t] != cudaSuccess) throw Error(FXII_E_CUDA, std::string("staged upload: ") + cudaGetErrorString(err[t]));
}

// The mirror image: a synchronous download into a large PAGEABLE host array (what `MatrixCSR.data` / `.indices` hand to
// scipy or PETSc: 1 GB for config 5's Π).  Every worker lets the copy engine fill one of its pinned buffers while it
// copies the other one out to the destination.  Returns after the last byte has reached `dst`.
void staged_download(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return;
  const int T = stage_threads();
  bool plain = bytes < kStageMin || T == 0;
  if (!plain) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, dst) != cudaSuccess) {
      cudaGetLastError();
      plain = true;
    } else if (attr.type != cudaMemoryTypeUnregistered) {
      plain = true;
    }
  }
  if (plain) {
    FXII_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s));
    FXII_CUDA(cudaStreamSynchronize(s));
    return;
  }
  int dev = 0;
  FXII_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_stage_mutex);
  std::vector<StageWorker>& ws = g_stage[dev];
  if ((int)ws.size() < T) {
    const size_t have = ws.size();
    ws.resize(T);
    for (size_t t = have; t < ws.size(); ++t) {
      FXII_CUDA(cudaStreamCreateWithFlags(&ws[t].stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        FXII_CUDA(cudaHostAlloc((void**)&ws[t].buf[b], kStageChunk, cudaHostAllocDefault));
        FXII_CUDA(cudaEventCreateWithFlags(&ws[t].ev[b], cudaEventDisableTiming));
      }
    }
  }
  FXII_CUDA(cudaStreamSynchronize(s));  // the source is complete
  const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
  std::vector<cudaError_t> err(T, cudaSuccess);
  std::vector<std::thread> threads;
  threads.reserve(T);
  for (int t = 0; t < T; ++t) {
    threads.emplace_back([&, t] {
      StageWorker& w = ws[t];
      cudaError_t e = cudaSetDevice(dev);
      for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaEventSynchronize(w.ev[b]);  // buffers free of earlier uploads
      auto issue = [&](size_t c, int b) {
        const size_t off = c * kStageChunk;
        const size_t n = std::min(kStageChunk, bytes - off);
        cudaError_t r = cudaMemcpyAsync(w.buf[b], static_cast<const unsigned char*>(src) + off, n, cudaMemcpyDeviceToHost, w.stream);
        if (r == cudaSuccess) r = cudaEventRecord(w.ev[b], w.stream);
        return r;
      };
      int b = 0;
      size_t c = (size_t)t;
      if (c < n_chunks && e == cudaSuccess) e = issue(c, b);
      for (; c < n_chunks && e == cudaSuccess; c += (size_t)T, b ^= 1) {
        const size_t next = c + (size_t)T;
        if (next < n_chunks) e = issue(next, b ^ 1);  // the copy engine fills the other buffer meanwhile
        if (e != cudaSuccess) break;
        e = cudaEventSynchronize(w.ev[b]);
        if (e != cudaSuccess) break;
        const size_t off = c * kStageChunk;
        memcpy(static_cast<unsigned char*>(dst) + off, w.buf[b], std::min(kStageChunk, bytes - off));
      }
      if (e != cudaSuccess) cudaStreamSynchronize(w.stream);
      err[t] = e;
    });
  }
  for (auto& th : threads) th.join();
  for (int t = 0; t < T; ++t)
    if (err[t] != cudaSuccess) throw Error(FXII_E_CUDA, std::string("staged download: ") + cudaGetErrorString(err[t]));
}
}  // namespace fxii

namespace fxii {
// the live block of the library's allocator that contains p (nullptr: not ours)
void* alloc_base_of(const void* p) {
  std::lock_guard<std::mutex> lock(g_mem_mutex);
  for (const auto& kv : g_live) {
    const char* b = static_cast<const char*>(kv.first);
    if (static_cast<const char*>(p) >= b && static_cast<const char*>(p) < b + kv.second.second) return kv.first;
  }
  return nullptr;
}
}  // namespace fxii

namespace fxii {
// SM copy: peer (NVLink-mapped) or local source -> local destination, four independent vector loads per thread in
// flight.  A copy-engine cudaMemcpyAsync out of an IPC-mapped peer allocation reached ~130 GB/s on this pool (2 GPUs,
// 0.47 GB: 3.5 ms); loads issued from all SMs keep enough bytes in flight to fill the link.
template <typename T>
__global__ void __launch_bounds__(256) peer_copy_kernel(T* __restrict__ dst, const T* __restrict__ src, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const T a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i] = a;
    dst[i + stride] = b;
    dst[i + 2 * stride] = c;
    dst[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) dst[i] = src[i];
}
}  // namespace fxii

namespace fxii {
// All segments of one exchange in ONE launch (8 ranks x 3 arrays would otherwise be 24 copy launches plus 8 pointer
// shifts per step: on a box whose host cores are shared by eight ranks the launch overhead rivals the transfer).
constexpr int kMaxSegments = 32;
struct CopySegments {
  void* dst[kMaxSegments];
  const void* src[kMaxSegments];
  long long bytes[kMaxSegments];
  long long add[kMaxSegments];  // != 0: the segment is int64 data and `add` is added to every element (row-pointer shift)
  int n;
};

template <typename T>
__device__ __forceinline__ void copy_span(T* __restrict__ dst, const T* __restrict__ src, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const T a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i] = a;
    dst[i + stride] = b;
    dst[i + 2 * stride] = c;
    dst[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) copy_segments_kernel(const CopySegments seg) {
  for (int k = 0; k < seg.n; ++k) {
    const long long bytes = seg.bytes[k];
    if (bytes <= 0) continue;
    if (seg.add[k] != 0) {
      long long* d = static_cast<long long*>(seg.dst[k]);
      const long long* s = static_cast<const long long*>(seg.src[k]);
      const long long add = seg.add[k];
      for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < bytes / 8; i += (long long)gridDim.x * blockDim.x)
        d[i] = s[i] + add;
      continue;
    }
    const uintptr_t al = (uintptr_t)seg.dst[k] | (uintptr_t)seg.src[k] | (uintptr_t)bytes;
    if ((al & 15) == 0) copy_span(static_cast<uint4*>(seg.dst[k]), static_cast<const uint4*>(seg.src[k]), bytes / 16);
    else if ((al & 7) == 0) copy_span(static_cast<uint2*>(seg.dst[k]), static_cast<const uint2*>(seg.src[k]), bytes / 8);
    else if ((al & 3) == 0) copy_span(static_cast<uint32_t*>(seg.dst[k]), static_cast<const uint32_t*>(seg.src[k]), bytes / 4);
    else copy_span(static_cast<unsigned char*>(seg.dst[k]), static_cast<const unsigned char*>(seg.src[k]), bytes);
  }
}
}  // namespace fxii

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return FXII_OK;
  } catch (const fxii::Error& e) {
    fxii::d2h_small_abandon();
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    fxii::d2h_small_abandon();
    g_last_error = e.what();
    return FXII_E_INVALID;
  } catch (...) {
    fxii::d2h_small_abandon();
    g_last_error = "unknown error";
    return FXII_E_INVALID;
  }
}

#define S(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int fxii_version(void) { return 100; }

const char* fxii_last_error(void) { return g_last_error.c_str(); }

int fxii_device_count(int* n) {
  return guarded([&] { FXII_CUDA(cudaGetDeviceCount(n)); });
}

int fxii_set_device(int device) {
  return guarded([&] { FXII_CUDA(cudaSetDevice(device)); });
}

int64_t fxii_launch_count(void) { return fxii::g_launches; }

int fxii_trim_memory(void) {
  return guarded([&] { fxii::cache_trim(); });
}

int fxii_sm_count(int* n) {
  return guarded([&] {
    int dev = 0;
    FXII_CUDA(cudaGetDevice(&dev));
    FXII_CUDA(cudaDeviceGetAttribute(n, cudaDevAttrMultiProcessorCount, dev));
  });
}

int fxii_mesh_create_cells(const double* x, int64_t n_nodes, const int32_t* cell_nodes, int64_t n_cells,
                           int32_t nodes_per_cell, double padding, void* stream, fxii_mesh** out) {
  return guarded([&] {
    FXII_CHECK(x && cell_nodes && out, "null argument");
    if (nodes_per_cell != 4 && nodes_per_cell != 8)
      throw fxii::Error(FXII_E_UNSUPPORTED, "cells must be affine tetrahedra (4 nodes) or Q1 hexahedra (8 nodes)");
    auto* m = new fxii_mesh();
    try {
      fxii::mesh_build(m, x, n_nodes, cell_nodes, n_cells, nodes_per_cell, padding, S(stream));
    } catch (...) {
      delete m;
      throw;
    }
    *out = m;
  });
}

int fxii_mesh_create(const double* x, int64_t n_nodes, const int32_t* cell_nodes, int64_t n_cells, double padding,
                     void* stream, fxii_mesh** out) {
  return fxii_mesh_create_cells(x, n_nodes, cell_nodes, n_cells, 4, padding, stream, out);
}

int fxii_mesh_info(const fxii_mesh* m, int64_t* n_cells, int64_t* n_nodes, int64_t* device_bytes, int32_t* max_depth) {
  return guarded([&] {
    FXII_CHECK(m, "null mesh");
    if (n_cells) *n_cells = m->n_cells;
    if (n_nodes) *n_nodes = m->n_nodes;
    if (device_bytes) *device_bytes = m->x.bytes() + m->cell_nodes.bytes() + m->cellgeo.bytes() + m->wnodes.bytes();
    if (max_depth) *max_depth = m->max_depth;
  });
}

int fxii_mesh_destroy(fxii_mesh* m) {
  return guarded([&] { delete m; });
}

int fxii_space_create(fxii_mesh* mesh, const int32_t* dofmap, int32_t nd, int32_t degree, int64_t n_dofs, void* stream,
                      fxii_space** out) {
  return guarded([&] {
    FXII_CHECK(mesh && dofmap && out, "null argument");
    const bool tet_ok = mesh->npc == 4 && ((degree == 1 && nd == 4) || (degree == 2 && nd == 10) || (degree == 3 && nd == 20));
    const bool hex_ok = mesh->npc == 8 && degree == 1 && nd == 8;
    if (!(tet_ok || hex_ok))
      throw fxii::Error(FXII_E_UNSUPPORTED,
                        "supported spaces: Lagrange P1 (nd=4) / P2 (nd=10) / P3 (nd=20) on affine tetrahedra, Q1 (nd=8) on hexahedra");
    auto* s = new fxii_space();
    s->mesh = mesh;
    s->nd = nd;
    s->degree = degree;
    s->n_dofs = n_dofs;
    try {
      s->dofmap.upload(dofmap, (int64_t)nd * mesh->n_cells, S(stream));
      fxii::check_index_range(s->dofmap.p, (int64_t)nd * mesh->n_cells, n_dofs, S(stream), "dof index");  // synchronises
    } catch (...) {
      delete s;
      throw;
    }
    *out = s;
  });
}

int fxii_space_destroy(fxii_space* s) {
  return guarded([&] { delete s; });
}

// Γ rows handle from its description.  `sync == false` leaves the uploads in flight on `s` (the caller
// synchronises before the host arrays may change).
static fxii_rows* make_rows(const fxii_rows_desc& d, cudaStream_t s, bool sync) {
  int32_t nq = d.nq;
  FXII_CHECK(d.nip >= 1 && d.n_cells_k >= 0 && d.n_rows_total >= 0, "bad argument");
  FXII_CHECK(d.kind == FXII_OP_TRACE || d.kind == FXII_OP_CIRCLE || d.kind == FXII_OP_DISK, "bad operator kind");
  FXII_CHECK(d.n_cells_k == 0 || (d.gx && d.gcells && d.xref), "null argument");
  FXII_CHECK(d.cells_k || d.n_cells_k <= d.n_gcells, "cells_k == NULL means cells 0..n_cells_k-1: n_cells_k exceeds the cell count");
  FXII_CHECK(d.row_dofs || d.n_cells_k * d.nip <= d.n_rows_total,
             "row_dofs == NULL means rows 0..n_cells_k*nip-1: more than n_rows_total");
  if (d.kind == FXII_OP_TRACE) nq = 1;
  else FXII_CHECK(nq >= 1 && d.table && d.w && d.radius, "circle/disk operators need table, w and radius");
  {  // range checks as min/max reductions (these loops vectorise; an early-exit loop does not)
    int32_t lo = 0, hi = -1;
    if (d.cells_k) {
      for (int64_t i = 0; i < d.n_cells_k; ++i) {
        lo = d.cells_k[i] < lo ? d.cells_k[i] : lo;
        hi = d.cells_k[i] > hi ? d.cells_k[i] : hi;
      }
      FXII_CHECK(lo >= 0 && hi < d.n_gcells, "cells_k out of range");
    }
    lo = 0, hi = -1;
    for (int64_t i = 0; i < 2 * d.n_gcells; ++i) {
      lo = d.gcells[i] < lo ? d.gcells[i] : lo;
      hi = d.gcells[i] > hi ? d.gcells[i] : hi;
    }
    FXII_CHECK(lo >= 0 && hi < d.n_gnodes, "gamma cell node out of range");
  }
  auto* r = new fxii_rows();
  try {
    r->kind = d.kind;
    r->nq = nq;
    r->nip = d.nip;
    r->n_point_rows = d.n_cells_k * d.nip;
    r->n_rows_total = d.n_rows_total;
    r->radius_per_row = d.radius_per_row;
    r->gx.upload(d.gx, 3 * d.n_gnodes, s);
    r->gcells.upload(d.gcells, 2 * d.n_gcells, s);
    // NULL cells_k / row_dofs: the identity, generated on the device instead of uploaded
    if (d.cells_k) r->cells_k.upload(d.cells_k, d.n_cells_k, s);
    else {
      r->cells_k.alloc(d.n_cells_k, s);
      fxii::device_iota(r->cells_k.p, d.n_cells_k, s);
    }
    r->xref.upload(d.xref, d.nip, s);
    if (d.row_dofs) r->row_dofs.upload(d.row_dofs, d.n_cells_k * d.nip, s);
    else {
      r->row_dofs.alloc(d.n_cells_k * d.nip, s);
      fxii::device_iota(r->row_dofs.p, d.n_cells_k * d.nip, s);
    }
    if (d.kind != FXII_OP_TRACE) {
      r->table.upload(d.table, 3 * (int64_t)nq, s);
      r->w.upload(d.w, nq, s);
      r->radius.upload(d.radius, d.radius_per_row == 1 ? d.n_cells_k * d.nip : (d.radius_per_row == 2 ? d.n_cells_k : 1), s);
    }
    if (sync) FXII_CUDA(cudaStreamSynchronize(s));
  } catch (...) {
    delete r;
    throw;
  }
  return r;
}

int fxii_rows_create(const double* gx, int64_t n_gnodes, const int32_t* gcells, int64_t n_gcells, const int32_t* cells_k,
                     int64_t n_cells_k, const double* xref, int32_t nip, const int32_t* row_dofs, int64_t n_rows_total,
                     int32_t kind, int32_t nq, const double* table, const double* w, const double* radius,
                     int32_t radius_per_row, void* stream, fxii_rows** out) {
  return guarded([&] {
    FXII_CHECK(out, "bad argument");
    const fxii_rows_desc d{gx, n_gnodes, gcells, n_gcells, cells_k, n_cells_k, xref, nip, row_dofs, n_rows_total,
                           kind, nq, table, w, radius, radius_per_row};
    *out = make_rows(d, S(stream), true);
  });
}

int fxii_rows_create_points(const double* points, const double* weights, const double* scales, int64_t n_point_rows,
                            int32_t nq, const int32_t* row_dofs, int64_t n_rows_total, void* stream, fxii_rows** out) {
  return guarded([&] {
    FXII_CHECK(out && nq >= 1 && n_point_rows >= 0, "bad argument");
    FXII_CHECK(n_point_rows == 0 || (points && weights && scales && row_dofs), "null argument");
    auto* r = new fxii_rows();
    cudaStream_t s = S(stream);
    try {
      r->kind = FXII_OP_POINTS;
      r->nq = nq;
      r->nip = 1;
      r->n_point_rows = n_point_rows;
      r->n_rows_total = n_rows_total;
      r->pts_in.upload(points, 3 * n_point_rows * nq, s);
      r->w_in.upload(weights, n_point_rows * nq, s);
      r->s_in.upload(scales, n_point_rows, s);
      r->row_dofs.upload(row_dofs, n_point_rows, s);
      FXII_CUDA(cudaStreamSynchronize(s));
    } catch (...) {
      delete r;
      throw;
    }
    *out = r;
  });
}

int fxii_rows_destroy(fxii_rows* r) {
  return guarded([&] { delete r; });
}

int fxii_rows_keep_points(fxii_rows* r, int32_t flag) {
  return guarded([&] {
    FXII_CHECK(r, "null rows");
    r->keep_points = flag != 0;
  });
}

int fxii_rows_set_mode(fxii_rows* r, int32_t mode) {
  return guarded([&] {
    FXII_CHECK(r && mode >= 0 && mode <= 2, "bad mode");
    r->mode = mode;
  });
}

int fxii_rows_set_nnz_hint(fxii_rows* r, int64_t nnz) {
  return guarded([&] {
    FXII_CHECK(r && nnz >= 0, "bad argument");
    r->nnz_hint = nnz;
  });
}

int64_t fxii_rows_get_nnz_hint(const fxii_rows* r) { return r ? r->nnz_hint : 0; }

int fxii_pi_assemble(const fxii_space* space, fxii_rows* rows, void* stream, fxii_csr** out, fxii_pi_stats* stats) {
  fxii_csr* c = nullptr;
  int rc = guarded([&] {
    FXII_CHECK(space && rows && out, "null argument");
    c = new fxii_csr();
    fxii::pi_assemble(space, rows, S(stream), c, stats, rows->keep_points);
  });
  if (rc == FXII_OK || rc == FXII_E_NOT_FOUND) {
    // on NOT_FOUND the matrix is still returned (entries of unfound points are zero) so that a
    // caller can inspect the diagnostics; the status tells it apart
    *out = c;
  } else {
    delete c;
    if (out) *out = nullptr;
  }
  return rc;
}

int fxii_pi_assemble_host(const fxii_space* space, fxii_rows* rows, void* stream, int64_t* indptr, int32_t* indices,
                          double* values, int64_t capacity, int64_t* nnz_out, fxii_pi_stats* stats) {
  fxii_csr c;
  fxii::HostSink sink{indptr, indices, values, capacity, false};
  int rc = guarded([&] {
    FXII_CHECK(space && rows && indptr && nnz_out && capacity >= 0 && (capacity == 0 || (indices && values)), "null argument");
    rows->sink = &sink;
    try {
      fxii::pi_assemble(space, rows, S(stream), &c, stats, rows->keep_points);
    } catch (...) {
      rows->sink = nullptr;
      throw;
    }
    rows->sink = nullptr;
  });
  if (rows) rows->sink = nullptr;
  if (nnz_out) *nnz_out = c.nnz;
  if (rc != FXII_OK && rc != FXII_E_NOT_FOUND) return rc;
  if (sink.done) return rc;
  const int rc2 = guarded([&] {
    FXII_CHECK(c.nnz <= capacity, "host arrays too small for the matrix (see *nnz_out)");
    cudaStream_t s = S(stream);
    FXII_CUDA(cudaMemcpyAsync(indptr, c.indptr.p, sizeof(int64_t) * (size_t)(c.n_rows + 1), cudaMemcpyDeviceToHost, s));
    if (c.nnz) {
      FXII_CUDA(cudaMemcpyAsync(indices, c.indices.p, sizeof(int32_t) * (size_t)c.nnz, cudaMemcpyDeviceToHost, s));
      FXII_CUDA(cudaMemcpyAsync(values, c.values.p, sizeof(double) * (size_t)c.nnz, cudaMemcpyDeviceToHost, s));
    }
    FXII_CUDA(cudaStreamSynchronize(s));
  });
  return rc2 != FXII_OK ? rc2 : rc;
}

int fxii_interpolation_matrix_host(const fxii_space* space, const fxii_rows_desc* desc, int64_t nnz_hint, void* stream,
                                   int64_t* indptr, int32_t* indices, double* values, int64_t capacity, int64_t* nnz_out,
                                   fxii_pi_stats* stats, fxii_rows** rows_out) {
  fxii_rows* r = nullptr;
  int rc = guarded([&] {
    FXII_CHECK(space && desc, "null argument");
    r = make_rows(*desc, S(stream), false);  // uploads stay in flight: the build synchronises before returning
    if (nnz_hint > 0) r->nnz_hint = nnz_hint;
  });
  if (rc == FXII_OK) rc = fxii_pi_assemble_host(space, r, stream, indptr, indices, values, capacity, nnz_out, stats);
  else cudaStreamSynchronize(S(stream));  // failed half way: nothing may still read the caller's arrays
  if (rows_out && (rc == FXII_OK || rc == FXII_E_NOT_FOUND)) *rows_out = r;
  else {
    if (rows_out) *rows_out = nullptr;
    cudaStreamSynchronize(S(stream));
    delete r;
  }
  return rc;
}

int fxii_rows_download_diag(const fxii_rows* rows, int32_t* cell, uint8_t* n_colliding, double* points, void* stream) {
  return guarded([&] {
    FXII_CHECK(rows, "null rows");
    cudaStream_t s = S(stream);
    const int64_t Np = rows->n_point_rows * rows->nq;
    FXII_CHECK(rows->cell.n == Np, "no assembly has run on this rows handle");
    if (cell && Np) FXII_CUDA(cudaMemcpyAsync(cell, rows->cell.p, sizeof(int32_t) * (size_t)Np, cudaMemcpyDeviceToHost, s));
    if (n_colliding && Np) FXII_CUDA(cudaMemcpyAsync(n_colliding, rows->ncoll.p, (size_t)Np, cudaMemcpyDeviceToHost, s));
    if (points) {
      FXII_CHECK(rows->points.n == 3 * Np, "points were not kept (fxii_rows_keep_points)");
      if (Np) FXII_CUDA(cudaMemcpyAsync(points, rows->points.p, sizeof(double) * 3 * (size_t)Np, cudaMemcpyDeviceToHost, s));
    }
    FXII_CUDA(cudaStreamSynchronize(s));
  });
}

int fxii_rows_download_coo(const fxii_rows* rows, int32_t* cols, double* vals, void* stream) {
  return guarded([&] {
    FXII_CHECK(rows, "null rows");
    cudaStream_t s = S(stream);
    const int64_t E = rows->n_point_rows * rows->nq * rows->nd;
    FXII_CHECK(rows->coo_cols.n == E && rows->nd > 0, "no assembly has run on this rows handle");
    FXII_CHECK(!rows->last_fused, "the last assembly used the fused path: no COO was materialised (fxii_rows_set_mode(rows, 1))");
    if (cols && E) FXII_CUDA(cudaMemcpyAsync(cols, rows->coo_cols.p, sizeof(int32_t) * (size_t)E, cudaMemcpyDeviceToHost, s));
    if (vals && E) FXII_CUDA(cudaMemcpyAsync(vals, rows->coo_vals.p, sizeof(double) * (size_t)E, cudaMemcpyDeviceToHost, s));
    FXII_CUDA(cudaStreamSynchronize(s));
  });
}

static fxii_csr* csr_from(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                          const double* values, cudaMemcpyKind kind, cudaStream_t s) {
  FXII_CHECK(n_rows >= 0 && n_cols >= 0 && nnz >= 0 && indptr, "bad CSR arguments");
  auto* c = new fxii_csr();
  try {
    c->n_rows = n_rows;
    c->n_cols = n_cols;
    c->nnz = nnz;
    c->indptr.alloc(n_rows + 1, s);
    c->indices.alloc(nnz, s);
    c->values.alloc(nnz, s);
    // host sources: large pageable arrays take the pipelined staging copy (staged_upload)
    auto put = [&](void* dst, const void* src, size_t bytes) {
      if (kind == cudaMemcpyHostToDevice) fxii::staged_upload(dst, src, bytes, s);
      else FXII_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, s));
    };
    put(c->indptr.p, indptr, sizeof(int64_t) * (size_t)(n_rows + 1));
    if (nnz) {
      put(c->indices.p, indices, sizeof(int32_t) * (size_t)nnz);
      put(c->values.p, values, sizeof(double) * (size_t)nnz);
    }
    FXII_CUDA(cudaStreamSynchronize(s));
  } catch (...) {
    delete c;
    throw;
  }
  return c;
}

int fxii_csr_upload(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                    const double* values, void* stream, fxii_csr** out) {
  return guarded([&] {
    FXII_CHECK(out, "null out");
    *out = csr_from(n_rows, n_cols, nnz, indptr, indices, values, cudaMemcpyHostToDevice, S(stream));
  });
}

int fxii_csr_from_device(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr_dev,
                         const int32_t* indices_dev, const double* values_dev, void* stream, fxii_csr** out) {
  return guarded([&] {
    FXII_CHECK(out, "null out");
    *out = csr_from(n_rows, n_cols, nnz, indptr_dev, indices_dev, values_dev, cudaMemcpyDeviceToDevice, S(stream));
  });
}

int fxii_csr_wrap_device(int64_t n_rows, int64_t n_cols, int64_t nnz, int64_t* indptr_dev, int32_t* indices_dev,
                         double* values_dev, fxii_csr** out) {
  return guarded([&] {
    FXII_CHECK(out && n_rows >= 0 && n_cols >= 0 && nnz >= 0 && indptr_dev, "bad CSR arguments");
    FXII_CHECK(nnz == 0 || (indices_dev && values_dev), "null array");
    // DevBuf::release only returns blocks the allocator handed out: foreign pointers are left alone
    auto* c = new fxii_csr();
    c->n_rows = n_rows;
    c->n_cols = n_cols;
    c->nnz = nnz;
    c->indptr.p = indptr_dev;
    c->indptr.n = n_rows + 1;
    c->indices.p = indices_dev;
    c->indices.n = nnz;
    c->values.p = values_dev;
    c->values.n = nnz;
    *out = c;
  });
}

int fxii_csr_validate(const fxii_csr* a, void* stream) {
  return guarded([&] {
    FXII_CHECK(a, "null csr");
    fxii::csr_validate(a, S(stream));
  });
}

// ---- peer-to-peer exchange of device arrays between the ranks of one node (CUDA IPC over NVLink) ---------------

int fxii_device_alloc(int64_t bytes, void* stream, void** dev_ptr) {
  return guarded([&] {
    FXII_CHECK(dev_ptr && bytes >= 0, "bad argument");
    *dev_ptr = bytes ? fxii::cached_alloc((size_t)bytes, S(stream)) : nullptr;
  });
}

int fxii_device_free(void* dev_ptr, void* stream) {
  return guarded([&] {
    if (dev_ptr) fxii::cached_free(dev_ptr, S(stream));
  });
}

int fxii_alloc_base(const void* dev_ptr, void** base) {
  return guarded([&] {
    FXII_CHECK(base, "null argument");
    *base = dev_ptr ? fxii::alloc_base_of(dev_ptr) : nullptr;
  });
}

int fxii_ipc_export(const void* base_ptr, unsigned char* handle64) {
  return guarded([&] {
    FXII_CHECK(base_ptr && handle64, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    if (fxii::alloc_base_of(base_ptr) != base_ptr)
      throw fxii::Error(FXII_E_UNSUPPORTED, "fxii_ipc_export: not the base of a block of this library's allocator");
    cudaIpcMemHandle_t h;
    FXII_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(base_ptr)));
    memcpy(handle64, &h, 64);
  });
}

int fxii_ipc_open(const unsigned char* handle64, void** dev_ptr) {
  return guarded([&] {
    FXII_CHECK(handle64 && dev_ptr, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    FXII_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  });
}

int fxii_ipc_close(void* dev_ptr) {
  return guarded([&] {
    if (dev_ptr) FXII_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  });
}


int fxii_copy_dev_async(void* dst_dev, const void* src_dev, int64_t bytes, void* stream) {
  return guarded([&] {
    FXII_CHECK(bytes >= 0 && (bytes == 0 || (dst_dev && src_dev)), "bad argument");
    if (bytes == 0) return;
    const uintptr_t al = (uintptr_t)dst_dev | (uintptr_t)src_dev | (uintptr_t)bytes;
    const int grid = fxii::num_sms() * 8;
    if (bytes < (1 << 20)) {
      FXII_CUDA(cudaMemcpyAsync(dst_dev, src_dev, (size_t)bytes, cudaMemcpyDefault, S(stream)));
    } else if ((al & 15) == 0) {
      fxii::peer_copy_kernel<uint4><<<grid, 256, 0, S(stream)>>>((uint4*)dst_dev, (const uint4*)src_dev, bytes / 16);
      FXII_LAUNCHED(1);
    } else if ((al & 7) == 0) {
      fxii::peer_copy_kernel<uint2><<<grid, 256, 0, S(stream)>>>((uint2*)dst_dev, (const uint2*)src_dev, bytes / 8);
      FXII_LAUNCHED(1);
    } else if ((al & 3) == 0) {
      fxii::peer_copy_kernel<uint32_t><<<grid, 256, 0, S(stream)>>>((uint32_t*)dst_dev, (const uint32_t*)src_dev, bytes / 4);
      FXII_LAUNCHED(1);
    } else {
      FXII_CUDA(cudaMemcpyAsync(dst_dev, src_dev, (size_t)bytes, cudaMemcpyDefault, S(stream)));
    }
    FXII_CUDA(cudaGetLastError());
  });
}

int fxii_copy_segments_async(int32_t n, void* const* dst_dev, const void* const* src_dev, const int64_t* bytes,
                             const int64_t* add_i64, void* stream) {
  return guarded([&] {
    FXII_CHECK(n >= 0 && (n == 0 || (dst_dev && src_dev && bytes)), "bad argument");
    for (int32_t k0 = 0; k0 < n; k0 += fxii::kMaxSegments) {
      fxii::CopySegments seg;
      seg.n = std::min<int32_t>(fxii::kMaxSegments, n - k0);
      for (int k = 0; k < seg.n; ++k) {
        seg.dst[k] = dst_dev[k0 + k];
        seg.src[k] = src_dev[k0 + k];
        seg.bytes[k] = bytes[k0 + k];
        seg.add[k] = add_i64 ? add_i64[k0 + k] : 0;
        FXII_CHECK(seg.bytes[k] >= 0 && (seg.bytes[k] == 0 || (seg.dst[k] && seg.src[k])), "bad segment");
        FXII_CHECK(seg.add[k] == 0 || (seg.bytes[k] % 8 == 0 && ((uintptr_t)seg.dst[k] % 8) == 0 && ((uintptr_t)seg.src[k] % 8) == 0),
                   "shifted segments must be int64 arrays");
      }
      fxii::copy_segments_kernel<<<fxii::num_sms() * 8, 256, 0, S(stream)>>>(seg);
      FXII_LAUNCHED(1);
      FXII_CUDA(cudaGetLastError());
    }
  });
}

int fxii_csr_info(const fxii_csr* a, int64_t* n_rows, int64_t* n_cols, int64_t* nnz) {
  return guarded([&] {
    FXII_CHECK(a, "null csr");
    if (n_rows) *n_rows = a->n_rows;
    if (n_cols) *n_cols = a->n_cols;
    if (nnz) *nnz = a->nnz;
  });
}

int fxii_csr_device_ptrs(const fxii_csr* a, void** indptr_dev, void** indices_dev, void** values_dev) {
  return guarded([&] {
    FXII_CHECK(a, "null csr");
    if (indptr_dev) *indptr_dev = a->indptr.p;
    if (indices_dev) *indices_dev = a->indices.p;
    if (values_dev) *values_dev = a->values.p;
  });
}

int fxii_csr_download(const fxii_csr* a, int64_t* indptr, int32_t* indices, double* values, void* stream) {
  return guarded([&] {
    FXII_CHECK(a, "null csr");
    cudaStream_t s = S(stream);
    // large pageable destinations take the pipelined staging copy (each call returns with its array complete)
    if (indptr) fxii::staged_download(indptr, a->indptr.p, sizeof(int64_t) * (size_t)(a->n_rows + 1), s);
    if (indices && a->nnz) fxii::staged_download(indices, a->indices.p, sizeof(int32_t) * (size_t)a->nnz, s);
    if (values && a->nnz) fxii::staged_download(values, a->values.p, sizeof(double) * (size_t)a->nnz, s);
    FXII_CUDA(cudaStreamSynchronize(s));
  });
}

int fxii_csr_download_async(const fxii_csr* a, int64_t* indptr, int32_t* indices, double* values, void* stream) {
  return guarded([&] {
    FXII_CHECK(a, "null csr");
    cudaStream_t s = S(stream);
    if (indptr) FXII_CUDA(cudaMemcpyAsync(indptr, a->indptr.p, sizeof(int64_t) * (size_t)(a->n_rows + 1), cudaMemcpyDeviceToHost, s));
    if (indices && a->nnz) FXII_CUDA(cudaMemcpyAsync(indices, a->indices.p, sizeof(int32_t) * (size_t)a->nnz, cudaMemcpyDeviceToHost, s));
    if (values && a->nnz) FXII_CUDA(cudaMemcpyAsync(values, a->values.p, sizeof(double) * (size_t)a->nnz, cudaMemcpyDeviceToHost, s));
  });
}

int fxii_csr_destroy(fxii_csr* a) {
  return guarded([&] { delete a; });
}

int fxii_csr_transpose(const fxii_csr* a, void* stream, fxii_csr** out) {
  return guarded([&] {
    FXII_CHECK(a && out, "null argument");
    auto* c = new fxii_csr();
    try {
      fxii::csr_transpose(a, 0, -1, S(stream), c);
      FXII_CUDA(cudaStreamSynchronize(S(stream)));
    } catch (...) {
      delete c;
      throw;
    }
    *out = c;
  });
}

int fxii_csr_transpose_window(const fxii_csr* a, int64_t col_begin, int64_t col_end, void* stream, fxii_csr** out) {
  return guarded([&] {
    FXII_CHECK(a && out, "null argument");
    auto* c = new fxii_csr();
    try {
      fxii::csr_transpose(a, col_begin, col_end, S(stream), c);
      FXII_CUDA(cudaStreamSynchronize(S(stream)));
    } catch (...) {
      delete c;
      throw;
    }
    *out = c;
  });
}

static int spgemm_entry(const fxii_csr* a, const fxii_csr* b, const double* b_scale, int64_t row_begin, int64_t row_end,
                        void* stream, fxii_csr** out, int64_t* ip) {
  return guarded([&] {
    FXII_CHECK(a && b && out, "null argument");
    if (b_scale) {
      cudaPointerAttributes at{};
      const bool on_device = cudaPointerGetAttributes(&at, b_scale) == cudaSuccess && at.type == cudaMemoryTypeDevice;
      if (!on_device) cudaGetLastError();
      FXII_CHECK(on_device, "b_row_scale_dev must be device memory");
    }
    auto* c = new fxii_csr();
    try {
      fxii::spgemm(a, b, b_scale, row_begin, row_end, S(stream), c, ip);
      FXII_CUDA(cudaStreamSynchronize(S(stream)));
    } catch (...) {
      delete c;
      throw;
    }
    *out = c;
  });
}

int fxii_spgemm(const fxii_csr* a, const fxii_csr* b, int64_t row_begin, int64_t row_end, void* stream, fxii_csr** out,
                int64_t* ip) {
  return spgemm_entry(a, b, nullptr, row_begin, row_end, stream, out, ip);
}

int fxii_spgemm_scaled(const fxii_csr* a, const fxii_csr* b, const double* b_row_scale_dev, int64_t row_begin,
                       int64_t row_end, void* stream, fxii_csr** out, int64_t* ip) {
  return spgemm_entry(a, b, b_row_scale_dev, row_begin, row_end, stream, out, ip);
}

int fxii_csr_expand_blocks(const fxii_csr* a, int32_t bs, void* stream, fxii_csr** out) {
  return guarded([&] {
    FXII_CHECK(a && out, "null argument");
    auto* c = new fxii_csr();
    try {
      fxii::csr_expand_blocks(a, bs, S(stream), c);
      FXII_CUDA(cudaStreamSynchronize(S(stream)));
    } catch (...) {
      delete c;
      throw;
    }
    *out = c;
  });
}

int fxii_csr_scale_rows(fxii_csr* a, const double* d, void* stream) {
  return guarded([&] {
    FXII_CHECK(a && d, "null argument");
    fxii::csr_scale_rows(a, d, S(stream));
  });
}

int fxii_csr_spmv(const fxii_csr* a, const double* u, double* y, void* stream) {
  return guarded([&] {
    FXII_CHECK(a && u && y, "null argument");
    fxii::csr_spmv(a, u, y, S(stream));
  });
}

int fxii_csr_spmv_t(const fxii_csr* a, const double* u, double* y, void* stream) {
  return guarded([&] {
    FXII_CHECK(a && u && y, "null argument");
    fxii_csr t;
    fxii::csr_transpose(a, 0, -1, S(stream), &t);
    fxii::csr_spmv(&t, u, y, S(stream));
  });
}

}  // extern "C"
