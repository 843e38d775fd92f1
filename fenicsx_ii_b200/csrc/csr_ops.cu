// CSR products of the block assembly: explicit transpose, row-wise hash SpGEMM, row scaling, SpMV.
//
// Reference calls replaced (file:line under /root/reference/src/fenicsx_ii/):
//   csr_transpose   PETSc Mat.transpose     matrix_assembler.py:184,236; demos/coupled_poisson.py:120-121
//   spgemm          PETSc Mat.matMult       matrix_assembler.py:185,210,237,245; demos/coupled_poisson.py:111,167,177
//   csr_scale_rows  M_Γ·Π for the diagonal quadrature-element mass (demos/coupled_poisson.py:164-167)
//   csr_spmv        K.mult                  assembly.py:67;  csr_spmv_t: vector_assembler.py:177
//
// PETSc semantics kept: the SYMBOLIC product pattern survives (cancellation zeros and Π's explicit
// zeros stay), columns are sorted, and c_ij accumulates a_ik*b_kj over k in A's column order, so
// results are bitwise reproducible run to run.
#include <cub/cub.cuh>
#include <type_traits>

#include "common.cuh"

namespace fxii {

// ---- transpose ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
expand_rows_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, int32_t* __restrict__ row_of) {
  // one warp per row
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const int64_t a = indptr[r], b = indptr[r + 1];
    for (int64_t k = a + lane; k < b; k += 32) row_of[k] = (int32_t)r;
  }
}

// row_of == nullptr: the source row of entry `src` is found by binary search in indptr (window transposes touch
// only a fraction of the entries: no O(nnz) expansion of the row indices)
__global__ void __launch_bounds__(256)
gather_transposed_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ row_of,
                         const int64_t* __restrict__ indptr, int64_t n_rows,
                         const double* __restrict__ values, int64_t nnz, int32_t* __restrict__ t_indices,
                         double* __restrict__ t_values) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t src = perm[k];
    int32_t row;
    if (row_of) {
      row = row_of[src];
    } else {  // last row r with indptr[r] <= src (empty rows share their successor's offset and are skipped)
      int64_t lo = 0, hi = n_rows - 1;
      while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(&indptr[mid]) <= (int64_t)src) lo = mid;
        else hi = mid - 1;
      }
      row = (int32_t)lo;
    }
    t_indices[k] = row;
    t_values[k] = values[src];
  }
}

__global__ void __launch_bounds__(256) iota32_kernel(int32_t* __restrict__ a, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = (int32_t)i;
}

static void exclusive_scan_i64(const long long* in, long long* out, int64_t n, cudaStream_t s) {
  size_t tb = 0;
  FXII_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, s));
  DevBuf<uint8_t> tmp;
  tmp.alloc((int64_t)tb, s);
  FXII_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, n, s));
}

struct InWindow {
  const int32_t* indices;
  int32_t c0, c1;
  __device__ bool operator()(const int32_t& k) const {
    const int32_t c = __ldg(&indices[k]);
    return c >= c0 && c < c1;
  }
};

__global__ void __launch_bounds__(256)
window_keys_kernel(const int32_t* __restrict__ sel, int64_t n_sel, const int32_t* __restrict__ indices, int32_t c0,
                   int32_t* __restrict__ keys, unsigned long long* __restrict__ count) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_sel; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t key = indices[sel[k]] - c0;
    keys[k] = key;
    atomicAdd(&count[key], 1ULL);
  }
}

// (Tried and rejected in round 2: a counting-scatter transpose — column histogram, scan, a warp per row of A scattering
// (row, value) to per-column atomic cursors, then a per-column insertion sort — measured 12.6 ms on config 5 against
// 4.0 ms for this pipeline: 84 M isolated 4- and 8-byte writes cost a sector read-modify-write each, the four radix
// passes stream.)
// Stable sort of the entries by column (radix sort is stable), so every row of Aᵀ lists the
// source rows in ascending order — the order MatTranspose produces.  With a column window
// [c0, c1) only those columns are transposed (rows of the result: c1 - c0): the entries are first
// compacted in order (cub select), so the sort touches only the shard.
void csr_transpose(const fxii_csr* a, int64_t c0, int64_t c1, cudaStream_t s, fxii_csr* out) {
  FXII_CHECK(a->nnz < ((int64_t)1 << 31), "transpose: nnz must fit int32 (shard the matrix)");
  if (c1 < 0) c1 = a->n_cols;
  FXII_CHECK(c0 >= 0 && c0 <= c1 && c1 <= a->n_cols, "transpose: bad column window");
  Trace tr("transpose", s);
  const int64_t n_out_rows = c1 - c0;
  const bool window = !(c0 == 0 && c1 == a->n_cols);
  out->n_rows = n_out_rows;
  out->n_cols = a->n_rows;
  const int64_t nnz = a->nnz;
  DevBuf<unsigned long long> count;
  count.alloc(n_out_rows + 1, s);
  FXII_CUDA(cudaMemsetAsync(count.p, 0, sizeof(unsigned long long) * (size_t)(n_out_rows + 1), s));
  out->indptr.alloc(n_out_rows + 1, s);
  DevBuf<int32_t> row_of, iota, sel, keys, perm, keys_sorted;
  int64_t n_sel = nnz;
  if (nnz > 0) {
    if (window) {
      // ordered selection of the entry positions whose column lies in the window, straight from a counting iterator:
      // the only O(nnz) work of a window transpose is one read of the column indices
      sel.alloc(nnz, s);
      DevBuf<int64_t> n_sel_dev;
      n_sel_dev.alloc(1, s);
      InWindow pred{a->indices.p, (int32_t)c0, (int32_t)c1};
      cub::CountingInputIterator<int32_t> positions(0);
      size_t tb = 0;
      FXII_CUDA(cub::DeviceSelect::If(nullptr, tb, positions, sel.p, n_sel_dev.p, nnz, pred, s));
      DevBuf<uint8_t> tmp;
      tmp.alloc((int64_t)tb, s);
      FXII_CUDA(cub::DeviceSelect::If(tmp.p, tb, positions, sel.p, n_sel_dev.p, nnz, pred, s));
      d2h_small(&n_sel, n_sel_dev.p, sizeof(int64_t), s);
      stream_sync(s);
    } else {
      row_of.alloc(nnz, s);
      iota.alloc(nnz, s);
      { expand_rows_kernel<<<grid_for(a->n_rows * 32, 256, 8), 256, 0, s>>>(a->indptr.p, a->n_rows, row_of.p); FXII_LAUNCHED(1); }
      { iota32_kernel<<<grid_for(nnz, 256, 8), 256, 0, s>>>(iota.p, nnz); FXII_LAUNCHED(1); }
    }
  }
  out->nnz = n_sel;
  out->indices.alloc(n_sel, s);
  out->values.alloc(n_sel, s);
  tr.mark("alloc + expand + select");
  if (n_sel > 0) {
    keys.alloc(n_sel, s);
    { window_keys_kernel<<<grid_for(n_sel, 256, 8), 256, 0, s>>>(window ? sel.p : iota.p, n_sel, a->indices.p, (int32_t)c0, keys.p, count.p); FXII_LAUNCHED(1); }
    FXII_CUDA(cudaGetLastError());
  }
  exclusive_scan_i64((const long long*)count.p, (long long*)out->indptr.p, n_out_rows + 1, s);
  tr.mark("keys + hist + scan");
  if (n_sel == 0) return;
  perm.alloc(n_sel, s);
  keys_sorted.alloc(n_sel, s);
  int end_bit = 1;
  while (end_bit < 31 && ((int64_t)1 << end_bit) < n_out_rows) ++end_bit;
  size_t tb = 0;
  const int32_t* src = window ? sel.p : iota.p;
  FXII_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.p, keys_sorted.p, src, perm.p, n_sel, 0, end_bit, s));
  DevBuf<uint8_t> tmp;
  tmp.alloc((int64_t)tb, s);
  FXII_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys_sorted.p, src, perm.p, n_sel, 0, end_bit, s));
  tr.mark("radix sort");
  { gather_transposed_kernel<<<grid_for(n_sel, 256, 8), 256, 0, s>>>(perm.p, window ? nullptr : row_of.p, a->indptr.p, a->n_rows,
                                                                 a->values.p, n_sel, out->indices.p, out->values.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
}

// ---- SpGEMM -------------------------------------------------------------------------------------
// Row-wise hash SpGEMM, two phases (symbolic count -> scan -> numeric), one GROUP (a warp, or a
// block for long rows) per output row, shared-memory tables sized by the row:
//   symbolic: rows with at most 512 intermediate products get an exact table (>= 2*ub_i slots).
//             Longer rows compress strongly in this problem (C_i is the union of overlapping
//             neighbourhoods), so they first try a 1024-slot warp table; rows that overflow it retry
//             in a 4096-slot warp table, then a 32768-slot block table, and what overflows that is
//             counted in global memory.
//   numeric : bins by the exact nnz_i from the symbolic phase (capacity >= 2*nnz_i).
// (A device-wide cub segmented sort of the finished rows was measured and rejected: 54-230 ms for
// 136 M entries in 17 M mostly empty segments, against ~18 ms for the in-kernel bitonic sort.)

__global__ void __launch_bounds__(256)
spgemm_ub_kernel(const int64_t* __restrict__ a_ptr, const int32_t* __restrict__ a_idx, const int64_t* __restrict__ b_ptr,
                 int64_t row_begin, int64_t n_rows, int64_t* __restrict__ ub) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a0 = a_ptr[row_begin + i], a1 = a_ptr[row_begin + i + 1];
    long long sum = 0;
    for (int64_t k = a0; k < a1; ++k) {
      const int c = a_idx[k];
      sum += b_ptr[c + 1] - b_ptr[c];
    }
    ub[i] = sum;
  }
}

constexpr int32_t kEmpty = -1;
constexpr int kNumBins = 9;  // bins 0..7: shared-memory tables of 64 << b slots; bin 8: long rows
__host__ __device__ constexpr uint32_t bin_cap(int b) { return 64u << b; }  // 64 ... 8192

__device__ __forceinline__ uint32_t hash_col(uint32_t c) { return c * 2654435761u; }

// Insert `key` into an open-addressing table of `cap` (power of two) slots; returns the slot
// (0xffffffff when the table is full).
// A plain load finds the (common) already-present key without an atomic.
__device__ __forceinline__ uint32_t hash_insert(int32_t* keys, uint32_t cap, int32_t key, bool& is_new) {
  uint32_t h = (hash_col((uint32_t)key) >> 7) & (cap - 1);
  is_new = false;
  for (uint32_t probe = 0; probe < cap; ++probe) {
    int32_t cur = *(volatile int32_t*)(keys + h);
    if (cur == key) return h;
    if (cur == kEmpty) {
      cur = atomicCAS(&keys[h], kEmpty, key);
      if (cur == kEmpty) {
        is_new = true;
        return h;
      }
      if (cur == key) return h;
    }
    h = (h + 1) & (cap - 1);
  }
  return 0xffffffffu;  // table full (only possible in an overflow-checked symbolic pass)
}

// NUMERIC=false: row_size[i] = number of distinct columns of row i, or -1 when more than
//   overflow_limit (> 0) distinct columns were seen (the row is retried with a larger table).
// NUMERIC=true : accumulate values — sequentially over k (A's column order), in parallel over the
//   entries of B's row k, whose columns are distinct, so no two lanes touch the same slot between two
//   group barriers and the floating-point sum order is fixed — then compact the occupied slots
//   into (column, slot) keys, bitonic-sort them and write the row in ascending column order.
// Table storage: shared memory (g_keys == nullptr) or a per-row slice of global memory.
// Shared layout per group: vals f64[cap] | sort u64[cap/2] | keys i32[cap] (NUMERIC); keys i32[cap].
template <bool BLOCK, bool NUMERIC>
__global__ void spgemm_rows_kernel(const int64_t* __restrict__ a_ptr, const int32_t* __restrict__ a_idx,
                                   const double* __restrict__ a_val, const int64_t* __restrict__ b_ptr,
                                   const int32_t* __restrict__ b_idx, const double* __restrict__ b_val,
                                   const double* __restrict__ b_scale,
                                   int64_t row_begin, const int32_t* __restrict__ row_list, int64_t n_list, uint32_t cap,
                                   int overflow_limit, int flatten, int* __restrict__ overflow_flag,
                                   int32_t* __restrict__ g_keys,
                                   double* __restrict__ g_vals, uint64_t* __restrict__ g_sort,
                                   const int64_t* __restrict__ g_off,
                                   int64_t* __restrict__ row_size, const int64_t* __restrict__ c_ptr,
                                   int32_t* __restrict__ c_idx, double* __restrict__ c_val) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int nthreads = BLOCK ? blockDim.x : 32;
  const int tid = BLOCK ? threadIdx.x : (threadIdx.x & 31);
  const int groups_per_block = BLOCK ? 1 : blockDim.x / 32;
  const int gid = BLOCK ? 0 : threadIdx.x / 32;
  __shared__ int s_cnt[32];
  __shared__ int s_new[32];  // distinct keys inserted so far, per group (tracked when overflow_limit > 0)
  for (int64_t li = (int64_t)blockIdx.x * groups_per_block + gid; li < n_list; li += (int64_t)gridDim.x * groups_per_block) {
    const int64_t i = row_list[li];  // row index relative to row_begin
    int32_t* keys;
    double* vals = nullptr;
    uint64_t* sortbuf = nullptr;
    uint32_t tcap = cap;
    if (g_keys) {
      const int64_t o0 = g_off[li], o1 = g_off[li + 1];
      tcap = (uint32_t)(o1 - o0);
      keys = g_keys + o0;
      if (NUMERIC) {
        vals = g_vals + o0;
        sortbuf = g_sort + o0 / 2;
      }
    } else {
      unsigned char* base = s_raw + (size_t)gid * cap * (NUMERIC ? 16 : 4);
      if (NUMERIC) {
        vals = reinterpret_cast<double*>(base);
        sortbuf = reinterpret_cast<uint64_t*>(base + (size_t)cap * 8);
        keys = reinterpret_cast<int32_t*>(base + (size_t)cap * 12);
      } else {
        keys = reinterpret_cast<int32_t*>(base);
      }
    }
    for (uint32_t t = tid; t < tcap; t += nthreads) {
      keys[t] = kEmpty;
      if (NUMERIC) vals[t] = 0.0;
    }
    if (tid == 0) {
      s_new[gid] = 0;
      s_cnt[gid] = 0;
    }
    group_sync<BLOCK>();
    const int64_t a0 = a_ptr[row_begin + i], a1 = a_ptr[row_begin + i + 1];
    bool overflow = false;
    if (!NUMERIC && (flatten & 1)) {
      // Symbolic phase: insertion order is irrelevant, so the (k, entry of B_k) pairs of 32 consecutive
      // k are flattened per warp — one lane per intermediate product instead of one warp pass per
      // (short) row of B.  Prefix sums and the k lookup run on shuffles.
      const int lane = threadIdx.x & 31;
      const int wig = BLOCK ? (threadIdx.x >> 5) : 0;        // warp in group
      const int nwarps = BLOCK ? (blockDim.x >> 5) : 1;
      for (int64_t c0 = a0 + 32 * (int64_t)wig; c0 < a1; c0 += 32 * (int64_t)nwarps) {
        if (overflow_limit > 0 && *(volatile int*)&s_new[gid] > overflow_limit) break;  // warp-uniform
        const int64_t ka = c0 + lane;
        int64_t b0 = 0;
        int len = 0;
        if (ka < a1) {
          const int k = __ldg(&a_idx[ka]);
          b0 = __ldg(&b_ptr[k]);
          len = (int)(__ldg(&b_ptr[k + 1]) - b0);
        }
        int incl = len;  // inclusive prefix sum of the row lengths over the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int t0 = 0; t0 < total; t0 += 32) {
          if (overflow_limit > 0 && *(volatile int*)&s_new[gid] > overflow_limit) break;  // warp-uniform
          const int t = t0 + lane;
          // owner lane j: smallest j with incl[j] > t (binary search over the warp's prefix sums)
          int lo = 0, hi = 31;
#pragma unroll
          for (int step = 0; step < 5; ++step) {
            const int mid = (lo + hi) >> 1;
            const int pm = __shfl_sync(0xffffffffu, incl, mid);
            if (pm > t) hi = mid;
            else lo = mid + 1;
          }
          const int j = lo;
          const int incl_j = __shfl_sync(0xffffffffu, incl, j);
          const int len_j = __shfl_sync(0xffffffffu, len, j);
          const int64_t b0_j = __shfl_sync(0xffffffffu, b0, j);
          if (t < total) {
            bool is_new;
            const uint32_t slot = hash_insert(keys, tcap, __ldg(&b_idx[b0_j + (t - (incl_j - len_j))]), is_new);
            if (overflow_limit > 0 && is_new) atomicAdd(&s_new[gid], 1);
            if (slot == 0xffffffffu) atomicAdd(&s_new[gid], (int)tcap);  // full: force the retry
          }
        }
        if (overflow_limit > 0) __syncwarp();
      }
    } else if (NUMERIC && !BLOCK && (flatten & 2)) {
      // Numeric phase, one warp per row: the same flattening, which batches the dependent loads
      // (A column -> B row pointer -> B entries) 32 products at a time.  The sum order stays the
      // sequential one (ascending k for every output column): batches run in order, and inside a
      // batch the lanes that hit the same column (different k, ascending with the lane) add one
      // after the other, ranked by __match_any_sync.
      const int lane = threadIdx.x & 31;
      for (int64_t c0 = a0; c0 < a1; c0 += 32) {
        const int64_t ka = c0 + lane;
        int64_t b0 = 0;
        int len = 0;
        double av = 0.0, dk = 1.0;
        if (ka < a1) {
          const int k = __ldg(&a_idx[ka]);
          av = __ldg(&a_val[ka]);
          b0 = __ldg(&b_ptr[k]);
          len = (int)(__ldg(&b_ptr[k + 1]) - b0);
          if (b_scale) dk = __ldg(&b_scale[k]);
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int t0 = 0; t0 < total; t0 += 32) {
          const int t = t0 + lane;
          int lo = 0, hi = 31;
#pragma unroll
          for (int step = 0; step < 5; ++step) {
            const int mid = (lo + hi) >> 1;
            const int pm = __shfl_sync(0xffffffffu, incl, mid);
            if (pm > t) hi = mid;
            else lo = mid + 1;
          }
          const int j = lo;
          const int incl_j = __shfl_sync(0xffffffffu, incl, j);
          const int len_j = __shfl_sync(0xffffffffu, len, j);
          const int64_t b0_j = __shfl_sync(0xffffffffu, b0, j);
          const double av_j = __shfl_sync(0xffffffffu, av, j);
          const double dk_j = __shfl_sync(0xffffffffu, dk, j);
          const bool active = t < total;
          int32_t col = -2 - lane;  // inactive lanes: distinct keys that match nothing
          double prod = 0.0;
          uint32_t slot = 0;
          if (active) {
            const int64_t kb = b0_j + (t - (incl_j - len_j));
            col = __ldg(&b_idx[kb]);
            prod = __dmul_rn(av_j, __dmul_rn(dk_j, __ldg(&b_val[kb])));  // (M B) first, as the unfused product rounds
            bool is_new;
            slot = hash_insert(keys, tcap, col, is_new);
          }
          const unsigned same = __match_any_sync(0xffffffffu, col);
          const int rank = __popc(same & ((1u << lane) - 1));
          unsigned pending = __ballot_sync(0xffffffffu, active);
          for (int r = 0; pending; ++r) {
            if (active && rank == r) vals[slot] = __dadd_rn(vals[slot], prod);
            __syncwarp();
            pending = __ballot_sync(0xffffffffu, active && rank > r);
          }
        }
      }
    } else
    {
      // k-sequential loop (long rows of B, and every block-per-row bin).  The row pointer chain of the NEXT k
      // (A column -> B row bounds) is requested one iteration ahead, and the entries of B's row are
      // loaded four per lane before the first insert, so that a row costs one memory latency, not
      // one per dependent load.
      int k_n = 0;
      double av_n = 0.0, dk_n = 1.0;
      int64_t b0_n = 0, b1_n = 0;
      if (a0 < a1) {
        k_n = __ldg(&a_idx[a0]);
        av_n = NUMERIC ? __ldg(&a_val[a0]) : 0.0;
        if (NUMERIC && b_scale) dk_n = __ldg(&b_scale[k_n]);
        b0_n = __ldg(&b_ptr[k_n]);
        b1_n = __ldg(&b_ptr[k_n + 1]);
      }
      for (int64_t ka = a0; ka < a1; ++ka) {
        const double av = av_n, dk = dk_n;
        const int64_t b0 = b0_n, b1 = b1_n;
        if (ka + 1 < a1) {
          k_n = __ldg(&a_idx[ka + 1]);
          av_n = NUMERIC ? __ldg(&a_val[ka + 1]) : 0.0;
          if (NUMERIC && b_scale) dk_n = __ldg(&b_scale[k_n]);
          b0_n = __ldg(&b_ptr[k_n]);
          b1_n = __ldg(&b_ptr[k_n + 1]);
        }
        if (overflow_limit > 0) {  // group-uniform: read after a barrier
          // B rows are shorter than the head-room left above the limit, so the table cannot fill up
          // between two checks
          overflow = *(volatile int*)&s_new[gid] > overflow_limit;
          if (overflow) break;
        }
        for (int64_t kb0 = b0 + tid; kb0 < b1; kb0 += 4 * (int64_t)nthreads) {
          int32_t col[4];
          double bv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t kb = kb0 + u * (int64_t)nthreads;
            col[u] = kb < b1 ? __ldg(&b_idx[kb]) : kEmpty;
            bv[u] = (NUMERIC && kb < b1) ? __ldg(&b_val[kb]) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (col[u] == kEmpty) continue;
            bool is_new;
            const uint32_t slot = hash_insert(keys, tcap, col[u], is_new);
            if (NUMERIC) vals[slot] = __dadd_rn(vals[slot], __dmul_rn(av, __dmul_rn(dk, bv[u])));
            if (overflow_limit > 0 && is_new) atomicAdd(&s_new[gid], 1);
            if (!NUMERIC && slot == 0xffffffffu) atomicAdd(&s_new[gid], (int)tcap);  // full: force the retry
          }
        }
        if (NUMERIC || overflow_limit > 0) group_sync<BLOCK>();
      }
    }
    group_sync<BLOCK>();
    if (!NUMERIC) {
      if (overflow_limit > 0) overflow = *(volatile int*)&s_new[gid] > overflow_limit;
      int local = 0;
      for (uint32_t t = tid; t < tcap; t += nthreads) local += keys[t] != kEmpty;
      for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
      if (BLOCK) {
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_cnt[0], local);
        __syncthreads();
        if (threadIdx.x == 0) {
          row_size[i] = overflow ? -1 : s_cnt[0];
          if (overflow) *overflow_flag = 1;
        }
        __syncthreads();
      } else if (tid == 0) {
        row_size[i] = overflow ? -1 : local;
        if (overflow) *overflow_flag = 1;
      }
      group_sync<BLOCK>();
    } else {
      const int64_t out0 = c_ptr[i];
      const uint32_t nnz = (uint32_t)(c_ptr[i + 1] - out0);
      uint32_t scap = 2;
      while (scap < nnz) scap <<= 1;
      for (uint32_t t = nnz + tid; t < scap; t += nthreads) sortbuf[t] = ~0ULL;
      // compact the occupied slots as (column, slot) keys; their order is irrelevant (sorted next)
      for (uint32_t t0 = 0; t0 < tcap; t0 += nthreads) {
        const uint32_t t = t0 + tid;
        const int32_t key = t < tcap ? keys[t] : kEmpty;
        const bool occ = key != kEmpty;
        const unsigned ball = __ballot_sync(0xffffffffu, occ);
        int wbase = 0;
        if ((threadIdx.x & 31) == 0 && ball) wbase = atomicAdd(&s_cnt[gid], __popc(ball));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (occ) sortbuf[wbase + __popc(ball & ((1u << (threadIdx.x & 31)) - 1))] = ((uint64_t)(uint32_t)key << 32) | t;
      }
      group_sync<BLOCK>();
      bitonic_sort_u64<BLOCK>(sortbuf, scap, tid, nthreads);
      for (uint32_t t = tid; t < nnz; t += nthreads) {
        const uint64_t e = sortbuf[t];
        c_idx[out0 + t] = (int32_t)(e >> 32);
        c_val[out0 + t] = vals[(uint32_t)(e & 0xffffffffu)];
      }
      group_sync<BLOCK>();
    }
  }
}

// =================================================================================================
// Lean warp-per-row kernel for SHORT rows of B (mean length <= 32: Π has 4..32 entries per row).
//
// The flattened loops above spend ~4 warp instructions per intermediate product on index arithmetic
// (a 5-step shuffle binary search per product to find its row of B).  Here G lanes (a power of two
// >= the mean row length of B) own one row of B and the warp walks 32/G rows of B per step:
//   stage : every lane loads ONE entry of A's row (k, a_ik), the bounds of B's row k and the optional
//           row scale d_k — 32 independent dependent-load chains in flight;
//   step  : three shuffles hand sub-group g the row j0+g; its lanes load B's entries (the next
//           step's loads are issued before the current step's table work) and insert them.
// Tables: keys i32[cap] + values f64[cap] = 12 B per slot (the round-1 kernel kept a separate 4 B/slot
// sort buffer and sized tables for a load factor of 0.5); rows are binned by nnz <= LF*cap.
// Numeric sum order is the sequential one (ascending k for every output column): entries of ONE row of
// B have distinct columns, so only different sub-groups can meet in a slot, and the sub-groups of a
// step add one after the other (ascending k), separated by __syncwarp().
// After the products the occupied slots are compacted IN PLACE (the write cursor never passes the read
// cursor), then sorted by column: up to 8 (column, position) keys per lane in REGISTERS — lane-crossing
// steps are shuffles — or, for larger rows, a key/value bitonic sort in the table's own storage.
// Symbolic passes detect a table that is too small by a probe limit (no per-insert counter atomics):
// the row is marked -1 and retried in the next larger table.
// =================================================================================================
constexpr uint32_t kNoSlot = 0xffffffffu;
constexpr uint32_t kProbeLimit = 48;  // linear probing exceeds this only beyond ~85 % load

__device__ __forceinline__ uint32_t table_insert(int32_t* keys, uint32_t mask, int shift, int32_t key, uint32_t max_probe) {
  uint32_t h = ((uint32_t)key * 0x9E3779B1u) >> shift;  // Fibonacci hashing: top bits of the product
  for (uint32_t probe = 0; probe <= max_probe; ++probe) {
    int32_t cur = *(volatile int32_t*)(keys + h);
    if (cur == key) return h;  // the common case (rows compress ~10x): no atomic
    if (cur == kEmpty) {
      cur = atomicCAS(&keys[h], kEmpty, key);
      if (cur == kEmpty || cur == key) return h;
    }
    h = (h + 1) & mask;
  }
  return kNoSlot;
}

// bitonic sort of the 32*R keys of a warp held R per lane (element e = R*lane + r)
template <int R>
__device__ __forceinline__ void warp_sort_regs(uint64_t (&key)[R], int lane) {
  constexpr uint32_t N = 32u * R;
#pragma unroll
  for (uint32_t k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      if (j >= (uint32_t)R) {
        // partner element e ^ j lives in lane ^ (j / R), same register; k >= 2R: the direction depends on the lane only
        const int lm = (int)(j / R);
        const bool up = (((uint32_t)R * (uint32_t)lane) & k) == 0;
        const bool take_min = ((lane & lm) == 0) == up;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint64_t other = __shfl_xor_sync(0xffffffffu, key[r], lm);
          key[r] = (take_min == (other < key[r])) ? other : key[r];
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((r & (int)j) == 0) {
            const bool up = ((((uint32_t)R * (uint32_t)lane) | (uint32_t)r) & k) == 0;
            const uint64_t a = key[r], b = key[r | (int)j];
            const bool sw = (a > b) == up;
            key[r] = sw ? b : a;
            key[r | (int)j] = sw ? a : b;
          }
        }
      }
    }
  }
}

// G: lanes per row of B (4, 8, 16, 32).  R: keys per lane of the register sort (numeric; 0: sort in shared memory).
// SCALE: a row scale of B is applied (a_ik * (d_k * b_kj)); compiled out otherwise.
// One warp per output row; W = blockDim.x/32 rows per block; dynamic shared memory W * cap * (NUMERIC ? 12 : 4) bytes.
// Entry offsets of B travel as 32 bits (the host takes this path only while nnz(B) < 2^31).
//
// Instruction budget (ncu, round 2: these kernels are ISSUE bound, 62-87 % issue-active, so instructions per product
// are what counts).  The first version ran ~120 instructions per step of NG rows: a software pipeline that renamed 12
// registers per step, 64-bit address arithmetic, a max-reduction per step for rows of B longer than G.  Now the loads
// of S steps are issued together into registers (no pipeline moves, S*2 loads in flight per lane), the steps are
// unrolled, and only the rows of B that are longer than G take a second pass for their tails.
// Resident blocks per SM the compiler must leave room for (0: no constraint), per register-sort width R.  Round 2 A/B
// on config 5 (profiles/r02_spgemm_occupancy_ab.txt): the two widest numeric bins were latency bound at low occupancy —
// R = 16 (80 registers, 3 blocks/SM, 37 % warps active, 46 % issue active) and R = 8 (64 registers, 4 blocks/SM);
// capping them at 64 and 48 registers (no spills / 8 bytes) gives 4 and 5 blocks/SM: 17.38 -> 16.29 ms for
// C = Pi^T (M_Gamma Pi).  The narrower bins are issue bound: more resident warps changed nothing (R = 4, 2 at 6 blocks).
#ifndef FXII_R16_MINBLOCKS
#define FXII_R16_MINBLOCKS 4
#endif
#ifndef FXII_R8_MINBLOCKS
#define FXII_R8_MINBLOCKS 5
#endif
#ifndef FXII_R4_MINBLOCKS
#define FXII_R4_MINBLOCKS 0
#endif
#ifndef FXII_R2_MINBLOCKS
#define FXII_R2_MINBLOCKS 0
#endif
template <int G, int R, bool NUMERIC, bool SCALE>
__global__ void __launch_bounds__(256, (R == 16) ? FXII_R16_MINBLOCKS : ((R == 8) ? FXII_R8_MINBLOCKS : ((R == 4) ? FXII_R4_MINBLOCKS : ((R == 2) ? FXII_R2_MINBLOCKS : 0))))
spgemm_warp_kernel(const int64_t* __restrict__ a_ptr, const int32_t* __restrict__ a_idx, const double* __restrict__ a_val,
                   const int64_t* __restrict__ b_ptr, const int32_t* __restrict__ b_idx, const double* __restrict__ b_val,
                   const double* __restrict__ b_scale, int64_t row_begin, const int32_t* __restrict__ row_list,
                   int64_t n_list, uint32_t cap, int shift, uint32_t max_probe, int* __restrict__ overflow_flag,
                   int64_t* __restrict__ row_size, const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_idx,
                   double* __restrict__ c_val, unsigned long long* __restrict__ ticket, int chunk) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr int NG = 32 / G;                              // rows of B per step
  constexpr int S = G >= 16 ? 4 : (G == 8 ? 2 : 1);       // steps whose loads are issued together (NG*S divides 32)
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int sub = lane & (G - 1), grp = lane / G;
  unsigned char* base = s_raw + (size_t)wid * cap * (NUMERIC ? 12 : 4);
  double* vals = reinterpret_cast<double*>(base);  // NUMERIC only
  int32_t* keys = reinterpret_cast<int32_t*>(base + (NUMERIC ? (size_t)cap * 8 : 0));
  const uint32_t mask = cap - 1;
  // one table operation of this lane: insert the column, then (numeric) add in sub-group order
  auto process = [&](int32_t col, double bv, double av_j, double dk_j, bool& fail) {
    const bool active = col != kEmpty;
    uint32_t slot = 0;
    if (active) slot = table_insert(keys, mask, shift, col, max_probe);
    if (!NUMERIC) {
      fail = fail || (active && slot == kNoSlot);
    } else {
      // explicit roundings (no FMA contraction, which differs between template instantiations): every term is
      // fl(a_ik * fl(d_k * b_kj)) added with one rounding — the arithmetic of the unfused product and of the oracle
      const double prod = SCALE ? __dmul_rn(av_j, __dmul_rn(dk_j, bv)) : __dmul_rn(av_j, bv);
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (active && grp == g) vals[slot] = __dadd_rn(vals[slot], prod);
        __syncwarp();
      }
    }
  };
  // Rows are handed out dynamically, `chunk` consecutive list entries per atomic (consecutive entries are consecutive
  // rows of A, whose rows of B overlap: a chunk keeps them in one warp's L1).
  for (;;) {
  long long tk = 0;
  if (lane == 0) tk = (long long)atomicAdd(ticket, (unsigned long long)chunk);
  tk = __shfl_sync(0xffffffffu, tk, 0);
  if (tk >= n_list) break;
  const int64_t li_end = min((int64_t)tk + chunk, n_list);
  for (int64_t li = tk; li < li_end; ++li) {
    const int64_t i = row_list[li];
    for (uint32_t t = lane; t < cap; t += 32) {
      keys[t] = kEmpty;
      if (NUMERIC) vals[t] = 0.0;
    }
    __syncwarp();
    const int64_t a0 = a_ptr[row_begin + i], a1 = a_ptr[row_begin + i + 1];
    bool overflow = false;
    for (int64_t c0 = a0; c0 < a1 && !overflow; c0 += 32) {
      // ---- stage: one entry of A's row per lane ----------------------------------------------------
      const int64_t ka = c0 + lane;
      uint32_t b0 = 0;
      int len = 0;
      double av = 0.0, dk = 1.0;
      if (ka < a1) {
        const int k = __ldg(&a_idx[ka]);
        const int64_t p0 = __ldg(&b_ptr[k]);
        b0 = (uint32_t)p0;
        len = (int)(__ldg(&b_ptr[k + 1]) - p0);
        if (NUMERIC) {
          av = __ldg(&a_val[ka]);
          if (SCALE) dk = __ldg(&b_scale[k]);
        }
      }
      const int nk = (int)min((int64_t)32, a1 - c0);
      bool fail = false;
      // ---- pass A: the first G entries of every row of B, NG rows per step, S steps per load group -------
      for (int j0 = 0; j0 < nk; j0 += NG * S) {  // warp-uniform
        int32_t col[S];
        double bv[S];
#pragma unroll
        for (int st = 0; st < S; ++st) {
          const int j = j0 + st * NG + grp;  // <= 31
          const uint32_t b0_j = __shfl_sync(0xffffffffu, b0, j);
          const int len_j = __shfl_sync(0xffffffffu, len, j);  // 0 beyond the row of A
          col[st] = kEmpty;
          bv[st] = 0.0;
          if (sub < len_j) {
            col[st] = __ldg(&b_idx[b0_j + (uint32_t)sub]);
            if (NUMERIC) bv[st] = __ldg(&b_val[b0_j + (uint32_t)sub]);
          }
        }
#pragma unroll
        for (int st = 0; st < S; ++st) {
          if (st > 0 && j0 + st * NG >= nk) break;  // warp-uniform
          const int j = j0 + st * NG + grp;
          double av_j = 0.0, dk_j = 1.0;
          if (NUMERIC) {
            av_j = __shfl_sync(0xffffffffu, av, j);
            if (SCALE) dk_j = __shfl_sync(0xffffffffu, dk, j);
          }
          process(col[st], bv[st], av_j, dk_j, fail);
        }
      }
      // ---- pass B: the tails of the rows of B that are longer than G, NG rows per step --------------------
      // (their contributions follow the heads of the later rows of this batch: the sum order of a column is fixed,
      // ascending k inside each pass)
      unsigned longmask = __ballot_sync(0xffffffffu, len > G);
      while (longmask) {  // warp-uniform
        int j = -1;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int bit = longmask ? __ffs(longmask) - 1 : -1;
          if (grp == g) j = bit;
          longmask &= longmask - 1;  // 0 stays 0
        }
        const int src = j < 0 ? 0 : j;  // every lane takes part in the shuffles; sub-groups without a row get length 0
        const uint32_t b0_j = __shfl_sync(0xffffffffu, b0, src);
        int len_j = __shfl_sync(0xffffffffu, len, src);
        if (j < 0) len_j = 0;
        double av_j = 0.0, dk_j = 1.0;
        if (NUMERIC) {
          av_j = __shfl_sync(0xffffffffu, av, src);
          if (SCALE) dk_j = __shfl_sync(0xffffffffu, dk, src);
        }
        const int maxlen = __reduce_max_sync(0xffffffffu, len_j);
        for (int t0 = G; t0 < maxlen; t0 += G) {  // warp-uniform
          const int t = t0 + sub;
          int32_t col = kEmpty;
          double bv = 0.0;
          if (t < len_j) {
            col = __ldg(&b_idx[b0_j + (uint32_t)t]);
            if (NUMERIC) bv = __ldg(&b_val[b0_j + (uint32_t)t]);
          }
          process(col, bv, av_j, dk_j, fail);
        }
      }
      if (!NUMERIC && __any_sync(0xffffffffu, fail)) overflow = true;  // warp-uniform; checked once per batch of A
    }
    __syncwarp();
    if (!NUMERIC) {
      int local = 0;
      for (uint32_t t = lane; t < cap; t += 32) local += keys[t] != kEmpty;
      for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
      if (lane == 0) {
        row_size[i] = overflow ? -1 : local;
        if (overflow) *overflow_flag = 1;
      }
      __syncwarp();
      continue;
    }
    // ---- in-place compaction of the occupied slots: (keys[e], vals[e]), e < nnz -------------------------
    uint32_t n_occ = 0;
    for (uint32_t t0 = 0; t0 < cap; t0 += 32) {
      const int32_t key = keys[t0 + lane];
      const double v = vals[t0 + lane];
      const bool occ = key != kEmpty;
      const unsigned ball = __ballot_sync(0xffffffffu, occ);
      __syncwarp();  // every lane has read its slot before any lane overwrites one (dst <= t0 + lane)
      if (occ) {
        const uint32_t dst = n_occ + __popc(ball & ((1u << lane) - 1u));
        keys[dst] = key;
        vals[dst] = v;
      }
      n_occ += __popc(ball);
    }
    __syncwarp();
    const int64_t out0 = c_ptr[i];
    const uint32_t nnz = (uint32_t)(c_ptr[i + 1] - out0);  // == n_occ (symbolic count)
    if constexpr (R > 0) {
      // sort (column, position) keys in registers: 32*R keys, or 16*R when the row fits half of that
      auto sort_write = [&](auto rr_tag) {
        constexpr int RR = decltype(rr_tag)::value;
        uint64_t key[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const uint32_t e = (uint32_t)RR * lane + r;
          key[r] = e < nnz ? (((uint64_t)(uint32_t)keys[e] << 32) | e) : ~0ULL;
        }
        warp_sort_regs<RR>(key, lane);
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const uint32_t e = (uint32_t)RR * lane + r;
          if (e < nnz) {
            c_idx[out0 + e] = (int32_t)(key[r] >> 32);
            c_val[out0 + e] = vals[(uint32_t)(key[r] & 0xffffffffu)];
          }
        }
      };
      if (R >= 2 && nnz <= 16u * R) sort_write(std::integral_constant<int, (R >= 2 ? R / 2 : 1)>{});  // warp-uniform
      else sort_write(std::integral_constant<int, R>{});
    } else {
      uint32_t scap = 2;
      while (scap < nnz) scap <<= 1;  // <= cap (nnz <= cap)
      for (uint32_t t = nnz + lane; t < scap; t += 32) keys[t] = 0x7fffffff;
      __syncwarp();
      for (uint32_t k = 2; k <= scap; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t t = lane; t < scap / 2; t += 32) {
            const uint32_t x = 2 * t - (t & (j - 1));
            const uint32_t y = x | j;
            const bool up = (x & k) == 0;
            const int32_t kx = keys[x], ky = keys[y];
            if ((kx > ky) == up) {
              keys[x] = ky;
              keys[y] = kx;
              const double vx = vals[x], vy = vals[y];
              vals[x] = vy;
              vals[y] = vx;
            }
          }
          __syncwarp();
        }
      }
      for (uint32_t t = lane; t < nnz; t += 32) {
        c_idx[out0 + t] = keys[t];
        c_val[out0 + t] = vals[t];
      }
    }
    __syncwarp();  // the table is re-initialised by the next row
  }
  }
}

// bin index by size: first bin b with size <= thr.v[b] (thr.v[b] < 0: bin unused); larger rows go to the last
// bin (long rows); empty rows are skipped
struct BinThresholds {
  int64_t v[kNumBins - 1];
};
__global__ void __launch_bounds__(256)
classify_rows_kernel(const int64_t* __restrict__ size, int64_t n_rows, BinThresholds thr, uint8_t* __restrict__ bin_of,
                     int32_t* __restrict__ rows, unsigned long long* __restrict__ bin_count) {
  __shared__ unsigned int s_count[kNumBins + 1];
  if (threadIdx.x <= kNumBins) s_count[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = size[i];
    int b = kNumBins;  // empty rows sort last and are never launched
    if (u > 0) {
      b = kNumBins - 1;  // long rows
#pragma unroll
      for (int q = kNumBins - 2; q >= 0; --q)
        if (u <= thr.v[q]) b = q;
    }
    bin_of[i] = (uint8_t)b;
    rows[i] = (int32_t)i;
    atomicAdd(&s_count[b], 1u);
  }
  __syncthreads();
  if (threadIdx.x <= kNumBins && s_count[threadIdx.x]) atomicAdd(&bin_count[threadIdx.x], (unsigned long long)s_count[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
global_table_sizes_kernel(const int32_t* __restrict__ row_list, int64_t n_list, const int64_t* __restrict__ size,
                          int64_t size_cap, int64_t* __restrict__ sizes) {
  for (int64_t li = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; li <= n_list; li += (int64_t)gridDim.x * blockDim.x) {
    if (li == n_list) {
      sizes[li] = 0;
      continue;
    }
    uint64_t u = (uint64_t)min(size[row_list[li]], size_cap) * 2;
    uint64_t cap = 1024;
    while (cap < u) cap <<= 1;
    sizes[li] = (int64_t)cap;
  }
}

struct IsMinusOne {
  const int64_t* v;
  __device__ bool operator()(const int32_t& i) const { return v[i] == -1; }
};

// rows grouped by bin: sorted_rows holds the row ids ordered by bin
struct RowBins {
  DevBuf<int32_t> sorted_rows;
  int64_t count[kNumBins + 1] = {0};
  int64_t offset[kNumBins + 1] = {0};
};

static void bin_rows(const int64_t* size, int64_t m, const BinThresholds& thr, cudaStream_t s, RowBins& rb) {
  DevBuf<uint8_t> bin_of, bin_sorted;
  DevBuf<int32_t> rows;
  DevBuf<unsigned long long> counts;
  bin_of.alloc(m, s);
  bin_sorted.alloc(m, s);
  rows.alloc(m, s);
  rb.sorted_rows.alloc(m, s);
  counts.alloc(kNumBins + 1, s);
  FXII_CUDA(cudaMemsetAsync(counts.p, 0, sizeof(unsigned long long) * (kNumBins + 1), s));
  { classify_rows_kernel<<<grid_for(m, 256, 8), 256, 0, s>>>(size, m, thr, bin_of.p, rows.p, counts.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
  size_t tb = 0;
  FXII_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, bin_of.p, bin_sorted.p, rows.p, rb.sorted_rows.p, m, 0, 4, s));
  DevBuf<uint8_t> tmp;
  tmp.alloc((int64_t)tb, s);
  FXII_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, bin_of.p, bin_sorted.p, rows.p, rb.sorted_rows.p, m, 0, 4, s));
  unsigned long long h[kNumBins + 1];
  d2h_small(h, counts.p, sizeof(h), s);
  stream_sync(s);
  int64_t off = 0;
  for (int b = 0; b <= kNumBins; ++b) {
    rb.count[b] = (int64_t)h[b];
    rb.offset[b] = off;
    off += rb.count[b];
  }
}

// bit 0: flattened symbolic loop, bit 1: flattened numeric loop (warp-per-row bins).  The numeric
// flattening pays while B's rows are short (it batches the dependent loads of ~32 rows of B; 36 -> 27 ms
// on the 1M-segment tree with 32-entry rows) and costs a few percent once a single row of B fills
// several warp passes on its own (256-entry rows), so it is chosen from B's mean row length.
// FXII_SPGEMM_FLAT overrides (0 restores the k-sequential loops).
static int flatten_mode(const fxii_csr* b) {
  static const int forced = [] {
    const char* e = getenv("FXII_SPGEMM_FLAT");
    return e ? atoi(e) : -1;
  }();
  if (forced >= 0) return forced;
  return b->nnz <= 64 * std::max<int64_t>(b->n_rows, 1) ? 3 : 1;
}

// FXII_SPGEMM_LEAN=0 restores the round-1 kernels everywhere (A/B measurements); FXII_SPGEMM_LF sets the load factor
// (percent) that bins rows into the numeric tables of the lean kernel.
static bool lean_enabled() {
  static const bool on = [] {
    const char* e = getenv("FXII_SPGEMM_LEAN");
    return !(e && e[0] == '0');
  }();
  return on;
}
// Table load factor (percent) of the lean kernel's numeric bins.  FXII_SPGEMM_LF: all bins (default 50);
// FXII_SPGEMM_LF_HI: the bins of 512 slots and more (default: the same).
static int lean_load_factor_pct(uint32_t cap) {
  static const int lf = [] {
    const char* e = getenv("FXII_SPGEMM_LF");
    const int v = e ? atoi(e) : 50;  // config 5, after the register caps: 30/40/45/50/60/70/80 % -> 18.3/16.1/15.7/15.5/16.0/16.3/16.5 ms
    return v < 30 ? 30 : (v > 85 ? 85 : v);
  }();
  static const int lf_hi = [] {
    const char* e = getenv("FXII_SPGEMM_LF_HI");
    const int v = e ? atoi(e) : lf;
    return v < 30 ? 30 : (v > 85 ? 85 : v);
  }();
  return cap >= 512 ? lf_hi : lf;
}
static int ilog2(uint32_t v) {
  int l = 0;
  while ((1u << l) < v) ++l;
  return l;
}

// C = A[row_begin:row_end, :] · diag(b_scale) · B.  b_scale (device, [b->n_rows]) may be null.
void spgemm(const fxii_csr* a, const fxii_csr* b, const double* b_scale, int64_t row_begin, int64_t row_end, cudaStream_t s,
            fxii_csr* out, int64_t* ip_out) {
  FXII_CHECK(a->n_cols == b->n_rows, "spgemm: inner dimensions differ");
  if (row_end < 0) row_end = a->n_rows;
  FXII_CHECK(row_begin >= 0 && row_begin <= row_end && row_end <= a->n_rows, "spgemm: bad row window");
  const int64_t m = row_end - row_begin;
  FXII_CHECK(m < ((int64_t)1 << 31), "spgemm: row window must fit int32");
  Trace tr("spgemm", s);
  out->n_rows = m;
  out->n_cols = b->n_cols;
  DevBuf<int64_t> ub, row_nnz;
  ub.alloc(m + 1, s);
  row_nnz.alloc(m + 1, s);
  FXII_CUDA(cudaMemsetAsync(ub.p, 0, sizeof(int64_t) * (size_t)(m + 1), s));
  FXII_CUDA(cudaMemsetAsync(row_nnz.p, 0, sizeof(int64_t) * (size_t)(m + 1), s));
  out->indptr.alloc(m + 1, s);
  if (m == 0) {
    FXII_CUDA(cudaMemsetAsync(out->indptr.p, 0, sizeof(int64_t), s));
    out->nnz = 0;
    if (ip_out) *ip_out = 0;
    return;
  }
  { spgemm_ub_kernel<<<grid_for(m, 256, 8), 256, 0, s>>>(a->indptr.p, a->indices.p, b->indptr.p, row_begin, m, ub.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
  DevBuf<int64_t> ip_dev;
  ip_dev.alloc(1, s);
  {
    size_t tb = 0;
    FXII_CUDA(cub::DeviceReduce::Sum(nullptr, tb, ub.p, ip_dev.p, m, s));
    DevBuf<uint8_t> tmp;
    tmp.alloc((int64_t)tb, s);
    FXII_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, ub.p, ip_dev.p, m, s));
  }
  int64_t h_ip = 0;
  d2h_small(&h_ip, ip_dev.p, sizeof(int64_t), s);
  DevBuf<int> overflow;
  overflow.alloc(1, s);

  // launch one round-1 kernel over a row list.  cap == 0: global-memory tables of 2*min(size, n_cols) slots.
  auto launch = [&](bool numeric, const int32_t* list, int64_t nl, uint32_t cap, bool block, int threads, int overflow_limit,
                    const int64_t* size, int32_t* c_idx, double* c_val) {
    if (nl == 0) return;
    const bool global = cap == 0;
    const int groups = block ? 1 : threads / 32;
    const size_t smem = global ? 0 : (size_t)cap * (numeric ? 16 : 4) * groups;
    const int grid = (int)std::min<int64_t>((nl + groups - 1) / groups, (int64_t)num_sms() * 16);
    DevBuf<int64_t> g_sizes, g_off;
    DevBuf<int32_t> g_keys;
    DevBuf<double> g_vals;
    DevBuf<uint64_t> g_sort;
    if (global) {
      g_sizes.alloc(nl + 1, s);
      g_off.alloc(nl + 1, s);
      { global_table_sizes_kernel<<<grid_for(nl + 1, 256, 8), 256, 0, s>>>(list, nl, size, b->n_cols, g_sizes.p); FXII_LAUNCHED(1); }
      exclusive_scan_i64((const long long*)g_sizes.p, (long long*)g_off.p, nl + 1, s);
      int64_t total = 0;
      d2h_small(&total, g_off.p + nl, sizeof(int64_t), s);
      stream_sync(s);
      FXII_CHECK(total < ((int64_t)1 << 33), "spgemm: rows too long for the global-table fallback (shard the product)");
      g_keys.alloc(total, s);
      if (numeric) {
        g_vals.alloc(total, s);
        g_sort.alloc(total / 2 + 1, s);
      }
    }
#define FXII_SPGEMM_LAUNCH(BLK, NUM)                                                                                  \
  do {                                                                                                                \
    if (smem + 1024 > 48 * 1024)                                                                                      \
      FXII_CUDA(cudaFuncSetAttribute(spgemm_rows_kernel<BLK, NUM>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                     (int)smem));                                                                     \
    spgemm_rows_kernel<BLK, NUM><<<grid, threads, smem, s>>>(                                                         \
        a->indptr.p, a->indices.p, a->values.p, b->indptr.p, b->indices.p, b->values.p, b_scale, row_begin, list, nl, \
        cap, overflow_limit, flatten_mode(b), overflow.p, global ? g_keys.p : nullptr, global ? g_vals.p : nullptr,   \
        global ? g_sort.p : nullptr, global ? g_off.p : nullptr, row_nnz.p, out->indptr.p, c_idx, c_val);             \
    FXII_CUDA(cudaGetLastError());                                                                                    \
    FXII_LAUNCHED(1);                                                                                                 \
  } while (0)
    if (block && numeric) FXII_SPGEMM_LAUNCH(true, true);
    else if (block && !numeric) FXII_SPGEMM_LAUNCH(true, false);
    else if (!block && numeric) FXII_SPGEMM_LAUNCH(false, true);
    else FXII_SPGEMM_LAUNCH(false, false);
#undef FXII_SPGEMM_LAUNCH
    if (global) stream_sync(s);  // tables are freed when this scope ends
  };

  // lean warp-per-row kernel (short rows of B): G lanes per row of B from B's mean row length
  const int64_t b_mean = b->nnz / std::max<int64_t>(b->n_rows, 1);
  const bool lean = lean_enabled() && b->nnz <= 32 * std::max<int64_t>(b->n_rows, 1) && b->nnz < ((int64_t)1 << 31);
  const int G = b_mean <= 4 ? 4 : (b_mean <= 8 ? 8 : (b_mean <= 16 ? 16 : 32));
  constexpr int kMaxWarpLaunches = 32;
  DevBuf<unsigned long long> tickets;
  int n_warp_launches = 0;
  if (lean) {
    tickets.alloc(kMaxWarpLaunches, s);
    FXII_CUDA(cudaMemsetAsync(tickets.p, 0, sizeof(unsigned long long) * kMaxWarpLaunches, s));
  }
  auto launch_warp = [&](bool numeric, const int32_t* list, int64_t nl, uint32_t cap, uint32_t max_probe, int32_t* c_idx,
                         double* c_val) {
    if (nl == 0) return;
    FXII_CHECK(n_warp_launches < kMaxWarpLaunches, "spgemm: too many kernel launches in one product");
    unsigned long long* ticket = tickets.p + n_warp_launches++;
    // warps per block: tables of a block within 48 KB, at most 8 warps
    const size_t per_warp = (size_t)cap * (numeric ? 12 : 4);
    int W = (int)std::min<size_t>(8, std::max<size_t>(1, (48 * 1024) / per_warp));
    const size_t smem = per_warp * W;
    const int grid = (int)std::min<int64_t>((nl + W - 1) / W, (int64_t)num_sms() * 32);
    // rows per ticket: about an eighth of a warp's share, 1..16
    const int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(16, nl / ((int64_t)grid * W * 8)));
    const int shift = 32 - ilog2(cap);
    // register sort: R keys per lane cover 32*R >= the largest row of the bin
    const int lf = lean_load_factor_pct(cap);
    const int64_t max_nnz = (int64_t)cap * lf / 100;
    int R = 0;
    if (numeric) R = max_nnz <= 64 ? 2 : (max_nnz <= 128 ? 4 : (max_nnz <= 256 ? 8 : (max_nnz <= 512 ? 16 : 0)));
#define FXII_WARP_LAUNCH(GG, RR, NUM, SC)                                                                              \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      FXII_CUDA(cudaFuncSetAttribute(spgemm_warp_kernel<GG, RR, NUM, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem));                                                                      \
    FXII_CUDA(cudaFuncSetAttribute(spgemm_warp_kernel<GG, RR, NUM, SC>, cudaFuncAttributePreferredSharedMemoryCarveout,\
                                   cudaSharedmemCarveoutMaxShared));                                                   \
    spgemm_warp_kernel<GG, RR, NUM, SC><<<grid, 32 * W, smem, s>>>(                                                    \
        a->indptr.p, a->indices.p, a->values.p, b->indptr.p, b->indices.p, b->values.p, b_scale, row_begin, list, nl,  \
        cap, shift, max_probe, overflow.p, row_nnz.p, out->indptr.p, c_idx, c_val, ticket, chunk);                     \
  } while (0)
#define FXII_WARP_S(GG, RR)                                                                                            \
  do {                                                                                                                 \
    if (b_scale) FXII_WARP_LAUNCH(GG, RR, true, true);                                                                 \
    else FXII_WARP_LAUNCH(GG, RR, true, false);                                                                        \
  } while (0)
#define FXII_WARP_R(GG)                                                                                                \
  do {                                                                                                                 \
    if (!numeric) FXII_WARP_LAUNCH(GG, 0, false, false);                                                               \
    else if (R == 2) FXII_WARP_S(GG, 2);                                                                               \
    else if (R == 4) FXII_WARP_S(GG, 4);                                                                               \
    else if (R == 8) FXII_WARP_S(GG, 8);                                                                               \
    else if (R == 16) FXII_WARP_S(GG, 16);                                                                             \
    else FXII_WARP_S(GG, 0);                                                                                           \
  } while (0)
    if (G == 4) FXII_WARP_R(4);
    else if (G == 8) FXII_WARP_R(8);
    else if (G == 16) FXII_WARP_R(16);
    else FXII_WARP_R(32);
#undef FXII_WARP_R
#undef FXII_WARP_S
#undef FXII_WARP_LAUNCH
    FXII_CUDA(cudaGetLastError());
    FXII_LAUNCHED(1);
  };

  // rows whose count came back as -1 (table overflow) -> list
  DevBuf<int32_t> iota, retry;
  DevBuf<int64_t> n_retry;
  auto select_overflowed = [&]() -> int64_t {
    int h_over = 0;
    d2h_small(&h_over, overflow.p, sizeof(int), s);
    stream_sync(s);
    if (!h_over) return 0;
    if (!iota.p) {
      iota.alloc(m, s);
      retry.alloc(m, s);
      n_retry.alloc(1, s);
      { iota32_kernel<<<grid_for(m, 256, 8), 256, 0, s>>>(iota.p, m); FXII_LAUNCHED(1); }
    }
    IsMinusOne pred{row_nnz.p};
    size_t tb = 0;
    FXII_CUDA(cub::DeviceSelect::If(nullptr, tb, iota.p, retry.p, n_retry.p, m, pred, s));
    DevBuf<uint8_t> tmp;
    tmp.alloc((int64_t)tb, s);
    FXII_CUDA(cub::DeviceSelect::If(tmp.p, tb, iota.p, retry.p, n_retry.p, m, pred, s));
    int64_t h_n = 0;
    d2h_small(&h_n, n_retry.p, sizeof(int64_t), s);
    FXII_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
    stream_sync(s);
    return h_n;
  };

  tr.mark("ub + alloc");
  // ---- symbolic --------------------------------------------------------------------------------
  // exact bins 0..4: tables of 64..1024 slots; round-1 kernels hold ub <= cap/2, the lean kernel ub <= 0.75 cap
  BinThresholds sym_thr;
  for (int q = 0; q < kNumBins - 1; ++q) sym_thr.v[q] = q < 5 ? (lean ? (int64_t)bin_cap(q) * 3 / 4 : (int64_t)bin_cap(q) / 2) : -1;
  RowBins sym;
  bin_rows(ub.p, m, sym_thr, s, sym);  // synchronises, so h_ip is valid
  tr.mark("bin rows (symbolic)");
  if (ip_out) *ip_out = h_ip;
  FXII_CUDA(cudaMemsetAsync(overflow.p, 0, sizeof(int), s));
  for (int bin = 0; bin < 5; ++bin) {
    const int32_t* list = sym.sorted_rows.p + sym.offset[bin];
    if (lean) launch_warp(false, list, sym.count[bin], bin_cap(bin), bin_cap(bin), nullptr, nullptr);
    else launch(false, list, sym.count[bin], bin_cap(bin), false, 256, 0, ub.p, nullptr, nullptr);
  }
  tr.mark("symbolic exact bins");
  const int64_t n_long = sym.count[kNumBins - 1];
  if (n_long > 0) {
    // escalating attempts: rows that overflow a table retry in the next, larger one; the last resort counts in
    // global memory.  Long rows compress strongly here (C_i is a union of overlapping neighbourhoods), so most
    // of them fit the first, 1024-slot table.
    const int32_t* list = sym.sorted_rows.p + sym.offset[kNumBins - 1];
    if (lean) launch_warp(false, list, n_long, 1024, kProbeLimit, nullptr, nullptr);
    else launch(false, list, n_long, 1024, false, 256, 768, ub.p, nullptr, nullptr);
    int64_t n_retry_rows = select_overflowed();
    if (n_retry_rows > 0) {
      if (lean) launch_warp(false, retry.p, n_retry_rows, 8192, kProbeLimit, nullptr, nullptr);
      else launch(false, retry.p, n_retry_rows, 4096, false, 128, 3072, ub.p, nullptr, nullptr);
      n_retry_rows = select_overflowed();
    }
    if (n_retry_rows > 0) {
      launch(false, retry.p, n_retry_rows, 32768, true, 256, 16384, ub.p, nullptr, nullptr);
      n_retry_rows = select_overflowed();
    }
    if (n_retry_rows > 0) launch(false, retry.p, n_retry_rows, 0, true, 256, 0, ub.p, nullptr, nullptr);
  }
  tr.mark("symbolic long rows");
  exclusive_scan_i64((const long long*)row_nnz.p, (long long*)out->indptr.p, m + 1, s);
  int64_t nnz = 0;
  d2h_small(&nnz, out->indptr.p + m, sizeof(int64_t), s);
  // ---- numeric ---------------------------------------------------------------------------------
  // bins 0..7: shared-memory tables of 64..8192 slots; bin 8: exact-size tables in global memory.
  // Lean kernel: bins 0..5 (64..2048 slots, 12 B per slot) hold nnz <= LF*cap; round-1 kernels: nnz <= cap/2.
  BinThresholds num_thr;
  for (int q = 0; q < kNumBins - 1; ++q)
    num_thr.v[q] = (lean && q <= 5) ? (int64_t)bin_cap(q) * lean_load_factor_pct(bin_cap(q)) / 100 : (int64_t)bin_cap(q) / 2;
  RowBins num;
  bin_rows(row_nnz.p, m, num_thr, s, num);  // synchronises
  out->nnz = nnz;
  out->indices.alloc(nnz, s);
  out->values.alloc(nnz, s);
  if (nnz == 0) return;
  tr.mark("bin rows (numeric) + alloc");
  const bool long_b = b->nnz >= 64 * std::max<int64_t>(b->n_rows, 1);
  for (int bin = 0; bin < kNumBins; ++bin) {
    const int32_t* list = num.sorted_rows.p + num.offset[bin];
    const uint32_t cap = bin < kNumBins - 1 ? bin_cap(bin) : 0;  // 0: exact-size tables in global memory
    if (lean && bin <= 5) {
      launch_warp(true, list, num.count[bin], cap, cap, out->indices.p, out->values.p);
    } else if (long_b) {
      // long rows of B (>= 64 entries on average): a whole block works on one output row from the 1024-slot bin
      // on (config 3: 2048-slot bin 85 -> 47 ms; the 512-slot bin measured slower that way) — one pass of the block covers a row of B, and the 16 B/slot table is shared by 4-16 warps instead
      // of pinning 8-32 KB of shared memory per WARP (6 resident warps per SM in the 2048-slot bin)
      if (bin <= 3) launch(true, list, num.count[bin], cap, false, 256, 0, row_nnz.p, out->indices.p, out->values.p);
      else if (bin == 4) launch(true, list, num.count[bin], cap, true, 128, 0, row_nnz.p, out->indices.p, out->values.p);
      else if (bin == 5) launch(true, list, num.count[bin], cap, true, 256, 0, row_nnz.p, out->indices.p, out->values.p);
      else launch(true, list, num.count[bin], cap, true, 512, 0, row_nnz.p, out->indices.p, out->values.p);
    } else
    // warp per row while the tables of a block (16 B per slot and warp) stay within 64 KB; whole blocks above
    if (bin <= 3) launch(true, list, num.count[bin], cap, false, 256, 0, row_nnz.p, out->indices.p, out->values.p);       // <= 512
    else if (bin == 4) launch(true, list, num.count[bin], cap, false, 128, 0, row_nnz.p, out->indices.p, out->values.p);  // 1024
    else if (bin == 5) launch(true, list, num.count[bin], cap, false, 64, 0, row_nnz.p, out->indices.p, out->values.p);   // 2048
    else if (bin == 6) launch(true, list, num.count[bin], cap, true, 128, 0, row_nnz.p, out->indices.p, out->values.p);   // 4096
    else launch(true, list, num.count[bin], cap, true, 256, 0, row_nnz.p, out->indices.p, out->values.p);                 // 8192, global
    if (tr.on && num.count[bin] > 0) {
      char label[64];
      snprintf(label, sizeof label, "numeric bin %d (%lld rows)", bin, (long long)num.count[bin]);
      tr.mark(label);
    }
  }
  tr.mark("numeric");
}

// ---- blocked spaces ---------------------------------------------------------------------------------
// Pi for blocked spaces (K.bs == V.bs == bs): every scalar entry (r, c, v) becomes the bs x bs block with
// v on its diagonal and EXPLICIT zeros elsewhere — the reference inserts all nd*bs unrolled columns for
// every component (interpolation.py:214-248 with utils.unroll_dofmap).
__global__ void __launch_bounds__(256)
expand_indptr_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, int bs, int64_t* __restrict__ out_ptr) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q <= n_rows * bs; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = q / bs, b = q - r * bs;
    // rows r*bs+b hold bs * nnz(r) entries each
    out_ptr[q] = q == n_rows * bs ? indptr[n_rows] * bs * bs : indptr[r] * bs * bs + b * (indptr[r + 1] - indptr[r]) * bs;
  }
}

__global__ void __launch_bounds__(256)
expand_blocks_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const double* __restrict__ values,
                     int64_t n_rows, int bs, const int64_t* __restrict__ out_ptr, int32_t* __restrict__ out_idx,
                     double* __restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const int64_t a = indptr[r], n = indptr[r + 1] - a;
    for (int64_t t = lane; t < n * bs * bs; t += 32) {
      const int64_t b = t / (n * bs);           // component row
      const int64_t rem = t - b * n * bs;
      const int64_t k = rem / bs, bc = rem - k * bs;  // scalar entry, component column
      const int64_t o = out_ptr[r * bs + b] + rem;
      out_idx[o] = indices[a + k] * bs + (int32_t)bc;
      out_val[o] = bc == b ? values[a + k] : 0.0;
    }
  }
}

void csr_expand_blocks(const fxii_csr* a, int bs, cudaStream_t s, fxii_csr* out) {
  FXII_CHECK(bs >= 1 && bs <= 16, "block size must be in 1..16");
  FXII_CHECK(a->n_cols * bs < ((int64_t)1 << 31), "blocked column index must fit int32");
  out->n_rows = a->n_rows * bs;
  out->n_cols = a->n_cols * bs;
  out->nnz = a->nnz * bs * bs;
  out->indptr.alloc(out->n_rows + 1, s);
  out->indices.alloc(out->nnz, s);
  out->values.alloc(out->nnz, s);
  { expand_indptr_kernel<<<grid_for(out->n_rows + 1, 256, 8), 256, 0, s>>>(a->indptr.p, a->n_rows, bs, out->indptr.p); FXII_LAUNCHED(1); }
  if (a->n_rows > 0 && a->nnz > 0) {
    { expand_blocks_kernel<<<grid_for(a->n_rows * 32, 256, 8), 256, 0, s>>>(a->indptr.p, a->indices.p, a->values.p, a->n_rows, bs,
                                                                         out->indptr.p, out->indices.p, out->values.p); FXII_LAUNCHED(1); }
  }
  FXII_CUDA(cudaGetLastError());
}

// ---- row scaling, SpMV --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scale_rows_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, const double* __restrict__ d, double* __restrict__ values) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const double f = d[r];
    for (int64_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) values[k] = f * values[k];
  }
}

// `d_any`: host array (uploaded; the call synchronises) or device array (used in place, stream-ordered)
void csr_scale_rows(fxii_csr* a, const double* d_any, cudaStream_t s) {
  cudaPointerAttributes at{};
  const bool on_device = cudaPointerGetAttributes(&at, d_any) == cudaSuccess && at.type == cudaMemoryTypeDevice;
  if (!on_device) cudaGetLastError();
  DevBuf<double> d;
  const double* dp = d_any;
  if (!on_device) {
    d.upload(d_any, a->n_rows, s);
    dp = d.p;
  }
  if (a->n_rows > 0) { scale_rows_kernel<<<grid_for(a->n_rows * 32, 256, 8), 256, 0, s>>>(a->indptr.p, a->n_rows, dp, a->values.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
  if (!on_device) stream_sync(s);
}

// one warp per row; lanes stride the row, partial sums combined by a fixed shuffle tree
__global__ void __launch_bounds__(256)
spmv_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const double* __restrict__ values,
            int64_t n_rows, const double* __restrict__ u, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    double sum = 0.0;
    for (int64_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) sum += values[k] * __ldg(&u[indices[k]]);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) y[r] = sum;
  }
}

void csr_spmv(const fxii_csr* a, const double* u_host, double* y_host, cudaStream_t s) {
  DevBuf<double> u, y;
  u.upload(u_host, a->n_cols, s);
  y.alloc(a->n_rows, s);
  if (a->n_rows > 0) {
    { spmv_kernel<<<grid_for(a->n_rows * 32, 256, 8), 256, 0, s>>>(a->indptr.p, a->indices.p, a->values.p, a->n_rows, u.p, y.p); FXII_LAUNCHED(1); }
    FXII_CUDA(cudaGetLastError());
    FXII_CUDA(cudaMemcpyAsync(y_host, y.p, sizeof(double) * (size_t)a->n_rows, cudaMemcpyDeviceToHost, s));
  }
  stream_sync(s);
}

// ---- validation of caller-supplied matrices ---------------------------------------------------------
// The products assume what PETSc's assembled AIJ guarantees: indptr monotone from 0 to nnz, columns in range and
// strictly ascending (hence unique) inside every row.  bad[0] = first violation kind (1 indptr, 2 column range,
// 3 order / duplicate).
__global__ void __launch_bounds__(256)
validate_csr_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n_rows, int64_t n_cols,
                    int64_t nnz, int* __restrict__ bad) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = indptr[r], b = indptr[r + 1];
    if (a > b || a < 0 || b > nnz || (r == 0 && a != 0) || (r == n_rows - 1 && b != nnz)) {
      atomicMax(bad, 1);
      continue;
    }
    int32_t prev = -1;
    for (int64_t k = a; k < b; ++k) {
      const int32_t c = indices[k];
      if (c < 0 || c >= n_cols) atomicMax(bad, 2);
      else if (c <= prev) atomicMax(bad, 3);
      prev = c;
    }
  }
}

void csr_validate(const fxii_csr* a, cudaStream_t s) {
  DevBuf<int> bad;
  bad.alloc(1, s);
  FXII_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
  if (a->n_rows > 0) {
    { validate_csr_kernel<<<grid_for(a->n_rows, 256, 8), 256, 0, s>>>(a->indptr.p, a->indices.p, a->n_rows, a->n_cols, a->nnz, bad.p); FXII_LAUNCHED(1); }
    FXII_CUDA(cudaGetLastError());
  }
  int h = 0;
  d2h_small(&h, bad.p, sizeof(int), s);
  stream_sync(s);
  FXII_CHECK(a->n_rows > 0 || a->nnz == 0, "CSR: entries without rows");
  FXII_CHECK(h != 1, "CSR: indptr must rise monotonically from 0 to nnz");
  FXII_CHECK(h != 2, "CSR: column index out of range");
  FXII_CHECK(h != 3, "CSR: columns must be strictly ascending inside every row (sort and sum duplicates first)");
}

}  // namespace fxii
