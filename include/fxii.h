/* fxii.h — C ABI of the B200-native reduction-operator assembly path.
 *
 * Drop-in boundary for the hot path of scientificcomputing/fenicsx_ii
 * (citations are file:line under /root/reference/src/fenicsx_ii/):
 *
 *   create_interpolation_matrix          interpolation.py:15-251
 *   ReductionOperator.compute_quadrature restriction_operators.py:17-47 (PointwiseTrace :60-68,
 *                                        Circle :143-205, Disk :256-315)
 *   Mat.transpose / Mat.matMult products matrix_assembler.py:175-258
 *   K.mult (Π·u)                         assembly.py:61-67
 *
 * The reference has no FFI of its own (pure Python over dolfinx/PETSc); these entry points are
 * what a ctypes / nanobind stub on the reference side binds.
 *
 * Conventions
 *  - every function returns 0 on success, a negative FXII_E_* code on failure;
 *    fxii_last_error() returns a thread-local message.  No exceptions cross the boundary.
 *  - all input arrays are HOST pointers unless the name ends in _dev; inputs are copied, never
 *    retained.  Handles own their device memory until *_destroy.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *    stream-ordered; functions that must return a size to the host synchronise the stream.
 *  - CSR: indptr int64[n_rows+1], indices int32[nnz] ascending and unique per row,
 *    values double[nnz]; explicit zeros are preserved.  Matrices the caller uploads must already have that
 *    form (fxii_csr_validate checks it): the products do not detect duplicate or unsorted columns.
 *  - the library acts on the CURRENT CUDA device of the calling thread.
 */
#ifndef FXII_H
#define FXII_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FXII_OK 0
#define FXII_E_INVALID (-1)   /* bad argument */
#define FXII_E_CUDA (-2)      /* CUDA runtime error */
#define FXII_E_NOT_FOUND (-3) /* some points lie in no cell (mirrors the assert at interpolation.py:78) */
#define FXII_E_UNSUPPORTED (-4)

/* operator kinds (restriction_operators.py) */
#define FXII_OP_TRACE 0  /* PointwiseTrace :50-75 */
#define FXII_OP_CIRCLE 1 /* Circle :78-205 */
#define FXII_OP_DISK 2   /* Disk :208-315 */
#define FXII_OP_POINTS 3 /* generic operator: caller ran compute_quadrature on the host */

typedef struct fxii_mesh fxii_mesh;   /* Ω: packed affine cell geometry + LBVH (replaces the tree
                                         dolfinx builds inside determine_point_ownership,
                                         interpolation.py:66-68) */
typedef struct fxii_space fxii_space; /* V: cell dofmap on a mesh (V.dofmap.list) */
typedef struct fxii_rows fxii_rows;   /* the Γ side of one Π build: cells_K, reference points,
                                         K dof per row, operator tables, radii */
typedef struct fxii_csr fxii_csr;     /* device CSR matrix */

int fxii_version(void);
const char* fxii_last_error(void);
int fxii_device_count(int* n);
int fxii_set_device(int device);
int fxii_sm_count(int* n);
/* number of kernels of THIS library launched so far by the calling process (CUB launches excluded) */
int64_t fxii_launch_count(void);
/* the library caches freed device blocks for reuse; this returns them to the driver */
int fxii_trim_memory(void);

/* ---- Ω mesh (affine tetrahedra) ---------------------------------------------------------- */
/* x: [n_nodes,3] f64, cell_nodes: [n_cells,4] i32 (mesh.geometry.x / mesh.geometry.dofmap).
 * padding: AABB padding of determine_point_ownership (tol of interpolation.py:19). */
int fxii_mesh_create(const double* x, int64_t n_nodes, const int32_t* cell_nodes, int64_t n_cells,
                     double padding, void* stream, fxii_mesh** out);
/* The same for cells with `nodes_per_cell` geometry nodes: 4 = affine tetrahedra, 8 = hexahedra with Q1 geometry,
 * vertices in dolfinx's tensor-product order v = i + 2j + 4k.  Hexahedra are pulled back by Newton's method
 * (interpolation_utils.py:36-38 with a non-affine cmap; tests/test_interpolation_matrix.py:31-34); ownership is the
 * collision with the convex hull of the eight vertices (cells with planar faces). */
int fxii_mesh_create_cells(const double* x, int64_t n_nodes, const int32_t* cell_nodes, int64_t n_cells,
                           int32_t nodes_per_cell, double padding, void* stream, fxii_mesh** out);
int fxii_mesh_info(const fxii_mesh* m, int64_t* n_cells, int64_t* n_nodes, int64_t* device_bytes,
                   int32_t* max_depth);
int fxii_mesh_destroy(fxii_mesh* m);

/* ---- V space ----------------------------------------------------------------------------- */
/* dofmap: [n_cells, nd] i32, basix dof order: tetrahedra nd = 4 (degree 1), 10 (degree 2) or 20 (degree 3, basix's default
 * gll_warped variant; the orientation of shared edges is in the dofmap, as dolfinx builds it); hexahedra nd = 8 (Q1).
 * Indices outside [0, n_dofs) are refused (FXII_E_INVALID, checked on the device). */
int fxii_space_create(fxii_mesh* mesh, const int32_t* dofmap, int32_t nd, int32_t degree,
                      int64_t n_dofs, void* stream, fxii_space** out);
int fxii_space_destroy(fxii_space* s);

/* ---- Γ rows ------------------------------------------------------------------------------ */
/* gx [n_gnodes,3], gcells [n_gcells,2]: Γ geometry.  cells_k [n_cells_k]: Γ cells processed
 * (cells_K of interpolation.py:47-51).  xref [nip]: reference interpolation points of K (:43).
 * row_dofs [n_cells_k*nip]: K.dofmap.list[cells_K] flattened.  cells_k == NULL: cells 0..n_cells_k-1;
 * row_dofs == NULL: rows 0..n_cells_k*nip-1 (both then need no upload).  table [nq,3], w [nq]: operator
 * tables (NULL for TRACE).  radius: [1] when radius_per_row == 0, [n_cells_k*nip] (one per point row) when 1,
 * [n_cells_k] (one per processed cell, expanded on the device) when 2. */
int fxii_rows_create(const double* gx, int64_t n_gnodes, const int32_t* gcells, int64_t n_gcells,
                     const int32_t* cells_k, int64_t n_cells_k, const double* xref, int32_t nip,
                     const int32_t* row_dofs, int64_t n_rows_total, int32_t kind, int32_t nq,
                     const double* table, const double* w, const double* radius,
                     int32_t radius_per_row, void* stream, fxii_rows** out);
/* generic operator: points [n_point_rows,nq,3], weights [n_point_rows,nq], scales [n_point_rows]
 * as returned by ReductionOperator.compute_quadrature (quadrature.py:11-20). */
int fxii_rows_create_points(const double* points, const double* weights, const double* scales,
                            int64_t n_point_rows, int32_t nq, const int32_t* row_dofs,
                            int64_t n_rows_total, void* stream, fxii_rows** out);
int fxii_rows_destroy(fxii_rows* r);
/* diagnostics: also materialise the generated 3D points [Np,3] (off by default: the points are
 * otherwise never written to HBM) */
int fxii_rows_keep_points(fxii_rows* r, int32_t flag);
/* mode 0 (default): reduce rows inside the locate kernel when every matrix row comes from one point row
 * (cell-local K), falling back to the COO path otherwise; mode 1: always materialise the COO
 * (needed by fxii_rows_download_coo); mode 2: fused, but as separate kernels (frames, rows, scan,
 * compaction; CUDA-graph replay on reused handles) instead of the single cooperative launch. */
int fxii_rows_set_mode(fxii_rows* r, int32_t mode);
/* Expected nnz of the matrix this handle will build (e.g. the nnz of the previous build of the same Γ and
 * operator): the CSR arrays are then allocated before the kernels run and the build synchronises with the
 * host once.  A wrong hint only costs a repeated compaction.  get returns the nnz of the last build. */
int fxii_rows_set_nnz_hint(fxii_rows* r, int64_t nnz);
int64_t fxii_rows_get_nnz_hint(const fxii_rows* r);

/* The arguments of fxii_rows_create as one record (same meaning, same NULL conventions). */
typedef struct {
  const double* gx; int64_t n_gnodes;
  const int32_t* gcells; int64_t n_gcells;
  const int32_t* cells_k; int64_t n_cells_k;
  const double* xref; int32_t nip;
  const int32_t* row_dofs; int64_t n_rows_total;
  int32_t kind, nq;
  const double* table; const double* w; const double* radius;
  int32_t radius_per_row;
} fxii_rows_desc;

/* ---- Π assembly -------------------------------------------------------------------------- */
typedef struct {
  int64_t n_points;     /* Np = n_point_rows*nq */
  int64_t n_entries;    /* E = Np*nd COO entries */
  int64_t n_not_found;  /* points owned by no cell */
  int64_t n_ties;       /* points colliding with more than one cell */
  int64_t n_fallback;   /* points assigned by the nearest-cell fallback */
  float ms_frames, ms_locate, ms_csr; /* device time of the three stages (CUDA events) */
  float ms_csr_count, ms_csr_fill;    /* the two row kernels inside ms_csr (COO path) */
  int32_t fused;                      /* 1: rows were reduced inside the locate kernel (no COO in HBM);
                                         2: same, replayed as one CUDA graph launch;
                                         3: same, as ONE cooperative kernel that also computes the row frames,
                                            scans the row lengths and writes the CSR (ms_frames = 0, ms_locate and
                                            ms_csr apportion the launch by its in-kernel timers) */
} fxii_pi_stats;

/* Builds Π (n_rows_total x n_dofs).  Leaves per-point diagnostics on `rows`
 * (fxii_rows_download_diag).  Returns FXII_E_NOT_FOUND when n_not_found > 0. */
int fxii_pi_assemble(const fxii_space* space, fxii_rows* rows, void* stream, fxii_csr** out,
                     fxii_pi_stats* stats);
/* The same build delivered to HOST arrays (the shape of the reference's create_interpolation_matrix,
 * interpolation.py:15-23, which returns a host matrix): indptr i64[n_rows_total+1], indices i32[capacity],
 * values f64[capacity], caller-allocated and ideally pinned.  With an nnz hint on `rows` the download is
 * enqueued behind the kernels before the single host synchronisation.  *nnz_out = nnz of the matrix;
 * FXII_E_INVALID when it exceeds `capacity` (nothing but *nnz_out is then valid). */
int fxii_pi_assemble_host(const fxii_space* space, fxii_rows* rows, void* stream, int64_t* indptr,
                          int32_t* indices, double* values, int64_t capacity, int64_t* nnz_out,
                          fxii_pi_stats* stats);
/* create_interpolation_matrix (interpolation.py:15-23) in ONE call, host arrays in, host CSR out: uploads Γ and the
 * operator (asynchronously when the arrays are pinned), builds Π and downloads it, synchronising `stream` once
 * before returning — the host arrays only have to stay valid during the call.  nnz_hint: nnz of an earlier build
 * of the same problem (0: unknown; the build then takes two synchronisations).  rows_out (may be NULL) receives the
 * rows handle for fxii_rows_download_diag; destroy it with fxii_rows_destroy. */
int fxii_interpolation_matrix_host(const fxii_space* space, const fxii_rows_desc* desc, int64_t nnz_hint,
                                   void* stream, int64_t* indptr, int32_t* indices, double* values,
                                   int64_t capacity, int64_t* nnz_out, fxii_pi_stats* stats,
                                   fxii_rows** rows_out);
/* owning cell i32[Np], colliding-set size u8[Np] (saturating), points f64[Np,3] (any may be NULL) */
int fxii_rows_download_diag(const fxii_rows* rows, int32_t* cell, uint8_t* n_colliding,
                            double* points, void* stream);
/* COO as written by the locate kernel: cols i32[E], vals f64[E]; entry e = p*nd + d belongs to
 * point p = e/nd, point row r = p/nq, matrix row row_dofs[r]. */
int fxii_rows_download_coo(const fxii_rows* rows, int32_t* cols, double* vals, void* stream);

/* ---- CSR objects and products ------------------------------------------------------------ */
int fxii_csr_upload(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr,
                    const int32_t* indices, const double* values, void* stream, fxii_csr** out);
/* adopt copies of DEVICE arrays (e.g. after an NCCL all-gather) */
int fxii_csr_from_device(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr_dev,
                         const int32_t* indices_dev, const double* values_dev, void* stream,
                         fxii_csr** out);
/* a matrix over DEVICE arrays the caller keeps alive and owns (no copy; *_destroy frees only the handle): the
 * gathered row shards of Π after the NCCL exchange */
int fxii_csr_wrap_device(int64_t n_rows, int64_t n_cols, int64_t nnz, int64_t* indptr_dev,
                         int32_t* indices_dev, double* values_dev, fxii_csr** out);
/* Checks what the products assume of a caller-assembled matrix (PETSc's assembled AIJ guarantees it): indptr
 * monotone from 0 to nnz, 0 <= column < n_cols, columns strictly ascending — hence unique — in every row.
 * FXII_E_INVALID otherwise.  Synchronises. */
int fxii_csr_validate(const fxii_csr* a, void* stream);
/* Peer-to-peer exchange between the ranks (processes) of one node: the all-gather-v of Π's row shards as direct
 * NVLink copies out of the peers' own arrays, straight into the gathered CSR (no staging, no padding).
 * fxii_alloc_base: the allocator block of THIS library that contains dev_ptr (NULL when the pointer is not ours,
 * e.g. a torch tensor: such arrays cannot be exported); fxii_ipc_export: 64-byte CUDA IPC handle of such a block;
 * fxii_ipc_open: map a peer's block (enables peer access lazily; keep the mapping and reuse it while the peer's
 * block is unchanged); fxii_copy_dev_async: stream-ordered copy between any two device pointers of the node
 * (unified addressing).  The caller orders the exchange: the source must be complete before the copy is
 * enqueued and must stay alive until it has run. */
/* a block of the library's allocator (exportable with fxii_ipc_export): the persistent send buffers of the exchange */
int fxii_device_alloc(int64_t bytes, void* stream, void** dev_ptr);
int fxii_device_free(void* dev_ptr, void* stream);
int fxii_alloc_base(const void* dev_ptr, void** base);
int fxii_ipc_export(const void* base_ptr, unsigned char* handle64);
int fxii_ipc_open(const unsigned char* handle64, void** dev_ptr);
int fxii_ipc_close(void* dev_ptr);
int fxii_copy_dev_async(void* dst_dev, const void* src_dev, int64_t bytes, void* stream);
/* n copies (HOST arrays of device pointers and sizes) in one kernel launch; add_i64 (may be NULL): a segment with a
 * nonzero entry is int64 data and gets that value added to every element while it is copied (the row pointers of a
 * shard shifted to their place in the gathered matrix). */
int fxii_copy_segments_async(int32_t n, void* const* dst_dev, const void* const* src_dev, const int64_t* bytes,
                             const int64_t* add_i64, void* stream);
int fxii_csr_info(const fxii_csr* a, int64_t* n_rows, int64_t* n_cols, int64_t* nnz);
int fxii_csr_device_ptrs(const fxii_csr* a, void** indptr_dev, void** indices_dev, void** values_dev);
int fxii_csr_download(const fxii_csr* a, int64_t* indptr, int32_t* indices, double* values,
                      void* stream);
/* The same without the final synchronisation: the copies are enqueued on `stream` (pinned destinations overlap with
 * kernels of other streams, e.g. the download of Π with the block products); the caller synchronises the stream and
 * keeps the matrix alive until then. */
int fxii_csr_download_async(const fxii_csr* a, int64_t* indptr, int32_t* indices, double* values,
                            void* stream);
int fxii_csr_destroy(fxii_csr* a);

/* Mat.transpose (matrix_assembler.py:184,236): rows of the result keep the source row order */
int fxii_csr_transpose(const fxii_csr* a, void* stream, fxii_csr** out);
/* rows [col_begin, col_end) of Aᵀ only (col_end < 0: n_cols): the V-dof row shard of one rank */
int fxii_csr_transpose_window(const fxii_csr* a, int64_t col_begin, int64_t col_end, void* stream,
                              fxii_csr** out);
/* Mat.matMult (matrix_assembler.py:185,210,237,245): symbolic pattern kept, columns sorted,
 * c_ij accumulated over k in A's column order.  ip (may be NULL) = intermediate products.
 * Optional row window [row_begin,row_end) of A (the V-dof row shard of SURVEY §8e);
 * pass 0, -1 for all rows. */
int fxii_spgemm(const fxii_csr* a, const fxii_csr* b, int64_t row_begin, int64_t row_end,
                void* stream, fxii_csr** out, int64_t* ip);
/* A · diag(d) · B with d = b_row_scale_dev (DEVICE, double[b.n_rows]): the diagonal quadrature-element mass M_Γ of
 * Πᵀ (M_Γ Π) (demos/coupled_poisson.py:164-167, matrix_assembler.py:225-258) applied while B's rows are read,
 * without materialising M_Γ Π.  Every term is a_ik * (d_k * b_kj): the rounding of the unfused product. */
int fxii_spgemm_scaled(const fxii_csr* a, const fxii_csr* b, const double* b_row_scale_dev,
                       int64_t row_begin, int64_t row_end, void* stream, fxii_csr** out, int64_t* ip);
/* blocked spaces (K.bs == V.bs == bs): every scalar entry becomes a bs x bs block with the value on its
 * diagonal and explicit zeros elsewhere (interpolation.py:214-248 + utils.unroll_dofmap, utils.py:36-47) */
int fxii_csr_expand_blocks(const fxii_csr* a, int32_t bs, void* stream, fxii_csr** out);
/* rows scaled by a diagonal: (diag(d)·A); d [n_rows], host memory (uploaded, synchronises) or device memory (used in
 * place, stream-ordered) — M_Γ·Π for a quadrature-element mass */
int fxii_csr_scale_rows(fxii_csr* a, const double* d, void* stream);
/* y = A·u (assembly.py:67); u host [n_cols], y host [n_rows] */
int fxii_csr_spmv(const fxii_csr* a, const double* u, double* y, void* stream);
/* y = Aᵀ·u (vector_assembler.py:177); u host [n_rows], y host [n_cols] */
int fxii_csr_spmv_t(const fxii_csr* a, const double* u, double* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FXII_H */
