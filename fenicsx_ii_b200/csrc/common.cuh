// Shared helpers for the fxii CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/fxii.h"

namespace fxii {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define FXII_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      throw ::fxii::Error(FXII_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + \
                                           " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

#define FXII_CHECK(cond, msg)                                       \
  do {                                                              \
    if (!(cond)) throw ::fxii::Error(FXII_E_INVALID, std::string(msg)); \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// grid for a grid-stride kernel: whole waves of the SM count, capped by the work available
inline int grid_for(int64_t work_items, int block, int blocks_per_sm) {
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)num_sms() * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Device memory comes from a size-class cache on top of cudaMalloc (capi.cu): blocks are rounded up to
// 1/8-octave size classes and kept on free lists, so the repeated, differently sized temporaries of a
// build never reach the driver again.  (The stream-ordered cudaMallocAsync pool stalled for up to 230 ms
// when multi-GB temporaries fragmented it.)  A block is reused in stream order; reuse from a different
// stream first synchronises the stream that last used it.
void* cached_alloc(size_t bytes, cudaStream_t s);
void cached_free(void* p, cudaStream_t s);
// pools of small per-handle resources (creating a pinned allocation or an event per handle costs more than
// a launch-bound build): 64-byte pinned host slots and timing events, recycled for the life of the process
void* pinned_slot_acquire();
void pinned_slot_release(void* p);
cudaEvent_t event_acquire();
void event_release(cudaEvent_t e);
void cache_trim();  // return all cached blocks to the driver
// host -> device copy ordered on s; large pageable sources go through a threaded, pipelined pinned staging copy
void staged_upload(void* dst, const void* src, size_t bytes, cudaStream_t s);
// device -> host copy that has COMPLETED on return; large pageable destinations go through the same staging buffers
void staged_download(void* dst, const void* src, size_t bytes, cudaStream_t s);

// Device -> host reads of a few bytes WITHOUT the copy engine.  A cudaMemcpyAsync(DeviceToHost) of 8 bytes queues behind
// every download in flight on any stream (one D2H engine): with a 1 GB CSR download running on a side stream each of
// the size / bin-count reads of a product waited for it to drain (end-to-end: 76 ms of products instead of 24).
// d2h_small enqueues a one-warp kernel that stores the bytes into a pinned, device-mapped page (zero-copy over PCIe);
// stream_sync synchronises the stream and then copies everything fetched by this thread to its destinations.  The
// destinations must stay valid until that stream_sync (they are locals of the calling function).
void d2h_small(void* host, const void* dev, size_t bytes, cudaStream_t s);
void stream_sync(cudaStream_t s);

// stream-ordered device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  int64_t n = 0;
  cudaStream_t st = nullptr;  // stream the buffer was allocated on; frees are ordered on it
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), st(o.st) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; st = o.st; o.p = nullptr; o.n = 0; }
    return *this;
  }
  void alloc(int64_t count, cudaStream_t s) {
    release();
    n = count;
    st = s;
    if (count > 0) p = static_cast<T*>(cached_alloc(sizeof(T) * (size_t)count, s));
  }
  // keep the allocation (and its address) when the size is unchanged; true if the buffer moved
  bool ensure(int64_t count, cudaStream_t s) {
    if (p && n == count) return false;
    alloc(count, s);
    return true;
  }
  void upload(const T* host, int64_t count, cudaStream_t s) {
    alloc(count, s);
    if (count > 0) staged_upload(p, host, sizeof(T) * (size_t)count, s);
  }
  void release() {
    if (p) cached_free(p, st);
    p = nullptr;
    n = 0;
  }
  int64_t bytes() const { return (int64_t)sizeof(T) * n; }
  ~DevBuf() { release(); }
};

// FXII_TRACE=1: print host wall time between marks (each mark synchronises the stream)
struct Trace {
  bool on;
  cudaStream_t s;
  const char* what;
  double t0;
  static double now();
  Trace(const char* w, cudaStream_t st);
  void mark(const char* label);
};
extern int64_t g_launches;  // kernels launched by this library (fxii_launch_count)
#define FXII_LAUNCHED(n) (::fxii::g_launches += (n))

// Host destination of a build (fxii_pi_assemble_host): when the build runs as one cooperative launch with a known
// nnz hint, the CSR download is enqueued behind the kernel BEFORE the single host synchronisation.
struct HostSink {
  int64_t* indptr;
  int32_t* indices;
  double* values;
  int64_t capacity;  // entries indices/values can hold
  bool done;         // the arrays hold the finished matrix
};

// ---- handle layouts -------------------------------------------------------------------------
struct alignas(16) BvhNode {  // 64 B: both child boxes live in the parent
  float lo_l[3], hi_l[3];
  float lo_r[3], hi_r[3];
  int32_t left, right;  // >= 0: internal node index; < 0: leaf, cell = ~child
  int32_t parent, pad;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

// 4-wide node, 64 B: the four grandchildren of a binary LBVH node with boxes quantised to 15 bits on
// the scene grid (rounded outward); hi halfwords carry bit 15 set so that two halfword compares are one
// 32-bit subtraction (SWAR).  lo[a][k] / hi[a][k]: axis a, slots (2k, 2k+1) packed as two
// halfwords.  child >= 0: wide node index; child < 0: leaf, cell = ~child.  Empty slot: lo > hi.
struct alignas(16) WideNode {
  uint32_t lo[3][2];
  uint32_t hi[3][2];
  int32_t child[4];
};
static_assert(sizeof(WideNode) == 64, "WideNode must be 64 bytes");

// scene grid of the quantised boxes: q(x) = floor((x - origin) * scale), clamped to [0, 32767]
struct SceneGrid {
  double origin[3];
  double scale[3];
};

// ---- device helpers shared by the row kernels -----------------------------------------------------
#ifdef __CUDACC__
// bitonic sort of cap (power of two) 64-bit keys in shared memory by `nthreads` cooperating threads
template <bool BLOCK>
__device__ __forceinline__ void group_sync() {
  if (BLOCK) __syncthreads();
  else __syncwarp();
}

template <bool BLOCK>
__device__ void bitonic_sort_u64(uint64_t* keys, uint32_t cap, int tid, int nthreads) {
  for (uint32_t k = 2; k <= cap; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = tid; t < cap / 2; t += nthreads) {
        const uint32_t i = 2 * t - (t & (j - 1));  // index with bit j cleared
        const uint32_t ixj = i | j;
        const bool up = (i & k) == 0;
        const uint64_t a = keys[i], b = keys[ixj];
        if ((a > b) == up) {
          keys[i] = b;
          keys[ixj] = a;
        }
      }
      group_sync<BLOCK>();
    }
  }
}

#endif  // __CUDACC__

}  // namespace fxii

struct fxii_mesh {
  int64_t n_cells = 0, n_nodes = 0;
  int32_t npc = 4;                   // geometry nodes per cell: 4 (affine tetrahedra) or 8 (Q1 hexahedra)
  double padding = 0;
  fxii::DevBuf<double> x;            // [n_nodes,3]
  fxii::DevBuf<int32_t> cell_nodes;  // [n_cells,npc]
  // tetrahedra: [n_cells,12] v0 (3) + K=J^{-1} row-major (9), 96 B per cell (the affine pull-back precomputed);
  // hexahedra:  [n_cells,24] the eight vertices gathered, 192 B per cell (the Newton pull-back and the hull test read them)
  fxii::DevBuf<double> cellgeo;
  fxii::DevBuf<fxii::WideNode> wnodes; // [max(n_cells-1,1)], indexed like the binary LBVH nodes
  fxii::SceneGrid grid;
  int32_t root = 0;
  int32_t max_depth = 0;
};

struct fxii_space {
  const fxii_mesh* mesh = nullptr;
  int32_t nd = 0, degree = 0;
  int64_t n_dofs = 0;
  fxii::DevBuf<int32_t> dofmap;  // [n_cells, nd]
};

struct fxii_rows {
  int32_t kind = 0, nq = 1, nip = 1;
  int64_t n_point_rows = 0;  // Nr
  int64_t n_rows_total = 0;  // rows of Π
  int32_t radius_per_row = 0;
  bool keep_points = false;  // also write the generated 3D points (diagnostics only)
  int mode = 0;              // 0: fused row reduction when possible; 1: always materialise the COO
  bool last_fused = false;
  int64_t nnz_hint = 0;          // expected nnz: the previous fused build on this handle, or set by the caller (0: unknown)
  int n_builds = 0;              // fused builds completed on this handle
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};      // stage timers (fused paths), from the event pool
  cudaEvent_t ev_coo[4] = {nullptr, nullptr, nullptr, nullptr};  // stage timers (COO path)
  // fused path: workspace and, for graph replay, handle-owned CSR buffers + the captured graph
  fxii::DevBuf<uint8_t> ws, scan_tmp;
  fxii::DevBuf<int64_t> g_indptr;
  fxii::DevBuf<int32_t> g_indices;
  fxii::DevBuf<double> g_values;
  void* host_result = nullptr;  // pinned FusedResult
  cudaGraphExec_t graph_exec = nullptr;
  const fxii_space* graph_space = nullptr;
  cudaStream_t graph_stream = nullptr;
  cudaStream_t capture_stream = nullptr;
  int64_t graph_cap = 0;
  int graph_kernels = 0;
  bool buffers_moved = false, last_graph = false, last_coop = false;
  fxii::HostSink* sink = nullptr;  // set for the duration of fxii_pi_assemble_host
  bool not_cell_local = false;     // a fused build found a K dof produced by two point rows
  ~fxii_rows() {
    for (auto& e : ev)
      if (e) fxii::event_release(e);
    for (auto& e : ev_coo)
      if (e) fxii::event_release(e);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (capture_stream) cudaStreamDestroy(capture_stream);
    if (host_result) fxii::pinned_slot_release(host_result);
  }
  // Γ inputs
  fxii::DevBuf<double> gx;
  fxii::DevBuf<int32_t> gcells, cells_k, row_dofs;
  fxii::DevBuf<double> xref, table, w, radius;
  // generic-operator inputs (kind == POINTS)
  fxii::DevBuf<double> pts_in, w_in, s_in;
  // per-build products kept for diagnostics / parity tests
  fxii::DevBuf<double> frames;   // [Nr,16] row frames
  fxii::DevBuf<double> points;   // [Np,3]
  fxii::DevBuf<int32_t> cell;    // [Np]
  fxii::DevBuf<uint8_t> ncoll;   // [Np]
  fxii::DevBuf<int32_t> coo_cols;  // [E]
  fxii::DevBuf<double> coo_vals;   // [E]
  int32_t nd = 0;
};

struct fxii_csr {
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
  fxii::DevBuf<int64_t> indptr;
  fxii::DevBuf<int32_t> indices;
  fxii::DevBuf<double> values;
};
