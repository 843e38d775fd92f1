// Ω mesh upload, packed affine cell geometry and LBVH build.
//
// Replaces the BoundingBoxTree dolfinx builds on every determine_point_ownership call
// (reference: interpolation.py:66-68).  Built once per mesh handle:
//   cell_geom_morton  : 1 thread / tet — v0 + J^{-1} (96 B, the affine pull-back of
//                       interpolation_utils.py:36-38 precomputed) and a 63-bit Morton key
//   cub radix sort    : (key, cell)
//   karras_internal   : 1 thread / internal node (Karras 2012), index tie-break on equal keys
//   refit             : 1 thread / leaf climbing with atomic arrival flags; every parent stores
//                       BOTH child boxes (fp32, rounded outward, padded)
//   collapse_wide     : 1 thread / binary node: its four grandchildren with 16-bit boxes on the
//                       scene grid become one 64 B wide node — half the levels and half the bytes
//                       per level of the binary tree (the traversal is bound by the L1 data pipe
//                       and by the length of its dependent-load chain)
#include <cub/cub.cuh>

#include "common.cuh"

namespace fxii {

// ---- ordered encoding of doubles for atomic min/max -------------------------------------------
__device__ __forceinline__ unsigned long long enc_ordered(double d) {
  unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__host__ __device__ __forceinline__ double dec_ordered(unsigned long long e) {
  unsigned long long b = (e & 0x8000000000000000ULL) ? (e & 0x7fffffffffffffffULL) : ~e;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}

// any a[i] outside [0, hi)?  (index arrays are validated on the device: two numpy reductions over config 5's 1.6 GB
// cell array cost 0.25 s on the host, this pass 0.3 ms)
__global__ void __launch_bounds__(256) index_range_kernel(const int32_t* __restrict__ a, int64_t n, int32_t hi,
                                                          int32_t* __restrict__ bad) {
  bool b = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = a[i];
    b = b || v < 0 || v >= hi;
  }
  if (__any_sync(0xffffffffu, b) && (threadIdx.x & 31) == 0) *bad = 1;
}

// bounds[0..2] = encoded min, bounds[3..5] = encoded max
__global__ void __launch_bounds__(256) scene_bounds_kernel(const double* __restrict__ x, int64_t n_nodes,
                                                           unsigned long long* __restrict__ bounds) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double v = x[3 * i + d];
      lo[d] = fmin(lo[d], v);
      hi[d] = fmax(hi[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      atomicMin(&bounds[d], enc_ordered(lo[d]));
      atomicMax(&bounds[3 + d], enc_ordered(hi[d]));
    }
  }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v) {
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffULL;
  x = (x | x << 16) & 0x1f0000ff0000ffULL;
  x = (x | x << 8) & 0x100f00f00f00f00fULL;
  x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}

// One thread per tetrahedron: gather the 4 vertices, write v0 + J^{-1} (cofactor formula, same
// operation order as oracle/pi_oracle.py:cell_jacobian_inverse) and the Morton key of the centroid.
__global__ void __launch_bounds__(256)
cell_geom_morton_kernel(const double* __restrict__ x, const int4* __restrict__ cell_nodes, int64_t n_cells,
                        const unsigned long long* __restrict__ bounds, double* __restrict__ cellgeo,
                        uint64_t* __restrict__ keys, int32_t* __restrict__ ids) {
  const double sx0 = dec_ordered(bounds[0]), sy0 = dec_ordered(bounds[1]), sz0 = dec_ordered(bounds[2]);
  const double ex = fmax(dec_ordered(bounds[3]) - sx0, 1e-300);
  const double ey = fmax(dec_ordered(bounds[4]) - sy0, 1e-300);
  const double ez = fmax(dec_ordered(bounds[5]) - sz0, 1e-300);
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
    int4 nd = __ldg(&cell_nodes[c]);
    double v[4][3];
    const int id[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int d = 0; d < 3; ++d) v[k][d] = __ldg(&x[3 * (int64_t)id[k] + d]);
    }
    const double j00 = v[1][0] - v[0][0], j01 = v[2][0] - v[0][0], j02 = v[3][0] - v[0][0];
    const double j10 = v[1][1] - v[0][1], j11 = v[2][1] - v[0][1], j12 = v[3][1] - v[0][1];
    const double j20 = v[1][2] - v[0][2], j21 = v[2][2] - v[0][2], j22 = v[3][2] - v[0][2];
    const double c00 = j11 * j22 - j12 * j21;
    const double c01 = j12 * j20 - j10 * j22;
    const double c02 = j10 * j21 - j11 * j20;
    const double det = j00 * c00 + j01 * c01 + j02 * c02;
    double g[12];
    g[0] = v[0][0];
    g[1] = v[0][1];
    g[2] = v[0][2];
    g[3] = c00 / det;
    g[4] = (j02 * j21 - j01 * j22) / det;
    g[5] = (j01 * j12 - j02 * j11) / det;
    g[6] = c01 / det;
    g[7] = (j00 * j22 - j02 * j20) / det;
    g[8] = (j02 * j10 - j00 * j12) / det;
    g[9] = c02 / det;
    g[10] = (j01 * j20 - j00 * j21) / det;
    g[11] = (j00 * j11 - j01 * j10) / det;
    double2* out = reinterpret_cast<double2*>(cellgeo + 12 * c);
#pragma unroll
    for (int k = 0; k < 6; ++k) out[k] = make_double2(g[2 * k], g[2 * k + 1]);
    const double cx = 0.25 * (v[0][0] + v[1][0] + v[2][0] + v[3][0]);
    const double cy = 0.25 * (v[0][1] + v[1][1] + v[2][1] + v[3][1]);
    const double cz = 0.25 * (v[0][2] + v[1][2] + v[2][2] + v[3][2]);
    const double s = 2097151.0;  // 2^21 - 1
    uint32_t qx = (uint32_t)fmin(fmax((cx - sx0) / ex * s, 0.0), s);
    uint32_t qy = (uint32_t)fmin(fmax((cy - sy0) / ey * s, 0.0), s);
    uint32_t qz = (uint32_t)fmin(fmax((cz - sz0) / ez * s, 0.0), s);
    keys[c] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    ids[c] = (int32_t)c;
  }
}

// One thread per hexahedron (Q1 geometry, vertices v = i + 2j + 4k): the eight vertices gathered into one 192 B
// record — the Newton pull-back (interpolation_utils.py:36-38 for non-affine cells) and the convex-hull collision
// test read them from there — and the Morton key of the vertex mean.
__global__ void __launch_bounds__(256)
hex_geom_morton_kernel(const double* __restrict__ x, const int32_t* __restrict__ cell_nodes, int64_t n_cells,
                       const unsigned long long* __restrict__ bounds, double* __restrict__ hexgeo,
                       uint64_t* __restrict__ keys, int32_t* __restrict__ ids) {
  const double sx0 = dec_ordered(bounds[0]), sy0 = dec_ordered(bounds[1]), sz0 = dec_ordered(bounds[2]);
  const double ex = fmax(dec_ordered(bounds[3]) - sx0, 1e-300);
  const double ey = fmax(dec_ordered(bounds[4]) - sy0, 1e-300);
  const double ez = fmax(dec_ordered(bounds[5]) - sz0, 1e-300);
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
    double cx = 0.0, cy = 0.0, cz = 0.0;
    double* out = hexgeo + 24 * c;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t id = __ldg(&cell_nodes[8 * c + k]);
      const double vx = __ldg(&x[3 * id]), vy = __ldg(&x[3 * id + 1]), vz = __ldg(&x[3 * id + 2]);
      out[3 * k] = vx;
      out[3 * k + 1] = vy;
      out[3 * k + 2] = vz;
      cx += vx;
      cy += vy;
      cz += vz;
    }
    cx *= 0.125;
    cy *= 0.125;
    cz *= 0.125;
    const double s = 2097151.0;  // 2^21 - 1
    uint32_t qx = (uint32_t)fmin(fmax((cx - sx0) / ex * s, 0.0), s);
    uint32_t qy = (uint32_t)fmin(fmax((cy - sy0) / ey * s, 0.0), s);
    uint32_t qz = (uint32_t)fmin(fmax((cz - sz0) / ez * s, 0.0), s);
    keys[c] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    ids[c] = (int32_t)c;
  }
}

// ---- Karras 2012 ------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256)
karras_internal_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ sorted_cells, int n,
                       BvhNode* __restrict__ nodes, int32_t* __restrict__ leaf_parent) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
      if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
      t = (t + 1) >> 1;
      if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    int left, right;
    if (lo == gamma) {
      left = ~sorted_cells[gamma];
      leaf_parent[gamma] = i;
    } else {
      left = gamma;
      nodes[gamma].parent = i;
    }
    if (hi == gamma + 1) {
      right = ~sorted_cells[gamma + 1];
      leaf_parent[gamma + 1] = i;
    } else {
      right = gamma + 1;
      nodes[gamma + 1].parent = i;
    }
    nodes[i].left = left;
    nodes[i].right = right;
    if (i == 0) nodes[0].parent = -1;
  }
}

struct BoxF {
  float lo[3], hi[3];
};

__device__ __forceinline__ void store_child_box(BvhNode* node, bool is_left, const BoxF& b) {
  float* dst = is_left ? node->lo_l : node->lo_r;  // lo then hi are contiguous (6 floats)
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    __stcg(dst + d, b.lo[d]);
    __stcg(dst + 3 + d, b.hi[d]);
  }
}

// One thread per leaf.  The leaf box is the dolfinx leaf box (min/max of the cell's nodes padded by
// `padding`) widened by the point_in_bbox slack and rounded OUTWARD to fp32, so the fp32 traversal
// never rejects a point the fp64 leaf test would accept.
__global__ void __launch_bounds__(256)
refit_kernel(const double* __restrict__ x, const int32_t* __restrict__ cell_nodes, int npc,
             const int32_t* __restrict__ sorted_cells,
             const int32_t* __restrict__ leaf_parent, int n, double padding, BvhNode* nodes, int32_t* flags) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int cell = sorted_cells[k];
    int id[8];
    if (npc == 4) {
      const int4 nd = __ldg(reinterpret_cast<const int4*>(cell_nodes) + cell);
      id[0] = nd.x, id[1] = nd.y, id[2] = nd.z, id[3] = nd.w;
      id[4] = nd.x, id[5] = nd.x, id[6] = nd.x, id[7] = nd.x;
    } else {
      const int4 n0 = __ldg(reinterpret_cast<const int4*>(cell_nodes) + 2 * (int64_t)cell);
      const int4 n1 = __ldg(reinterpret_cast<const int4*>(cell_nodes) + 2 * (int64_t)cell + 1);
      id[0] = n0.x, id[1] = n0.y, id[2] = n0.z, id[3] = n0.w;
      id[4] = n1.x, id[5] = n1.y, id[6] = n1.z, id[7] = n1.w;
    }
    BoxF box;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double lo = 1e300, hi = -1e300;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q >= npc) break;
        double v = __ldg(&x[3 * (int64_t)id[q] + d]);
        lo = fmin(lo, v);
        hi = fmax(hi, v);
      }
      lo -= padding;
      hi += padding;
      const double slack = 2e-14 * (hi - lo) + 1e-300;
      box.lo[d] = __double2float_rd(lo - slack);
      box.hi[d] = __double2float_ru(hi + slack);
    }
    int child = ~cell;
    int p = leaf_parent[k];
    while (true) {
      BvhNode* np = nodes + p;
      const bool is_left = (np->left == child);
      store_child_box(np, is_left, box);
      __threadfence();
      const int old = atomicAdd(&flags[p], 1);
      if (old == 0) break;  // the sibling subtree finishes this node
      __threadfence();
      // both child boxes are now visible: union them
      BoxF u;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        u.lo[d] = fminf(__ldcg(&np->lo_l[d]), __ldcg(&np->lo_r[d]));
        u.hi[d] = fmaxf(__ldcg(&np->hi_l[d]), __ldcg(&np->hi_r[d]));
      }
      box = u;
      const int parent = np->parent;
      if (parent < 0) break;
      child = p;
      p = parent;
    }
  }
}


// ---- 4-wide collapse --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t quant(double v, double origin, double scale) {
  const double q = floor((v - origin) * scale);  // monotone in v: conservative for lo and hi alike
  return (uint32_t)fmin(fmax(q, 0.0), 32767.0);
}

struct Slot {
  uint32_t lo[3], hi[3];
  int32_t child;
};

__device__ __forceinline__ Slot make_slot(const float* lo, const float* hi, int32_t child, const SceneGrid& g) {
  Slot s;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    s.lo[a] = quant((double)lo[a], g.origin[a], g.scale[a]);
    s.hi[a] = quant((double)hi[a], g.origin[a], g.scale[a]);
  }
  s.child = child;
  return s;
}

// Wide node i = the grandchildren of binary node i (a leaf child fills one slot itself).  Traversal
// starts at node 0 and only ever follows wide-node children, i.e. every second binary level.
__global__ void __launch_bounds__(256)
collapse_wide_kernel(const BvhNode* __restrict__ nodes, int n_internal, SceneGrid grid, WideNode* __restrict__ wide) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_internal; i += gridDim.x * blockDim.x) {
    const BvhNode b = nodes[i];
    Slot slots[4];
    int ns = 0;
    const int kids[2] = {b.left, b.right};
    const float* klo[2] = {b.lo_l, b.lo_r};
    const float* khi[2] = {b.hi_l, b.hi_r};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (klo[k][0] > khi[k][0]) continue;  // empty box (single-cell mesh)
      if (kids[k] < 0) {
        slots[ns++] = make_slot(klo[k], khi[k], kids[k], grid);
      } else {
        const BvhNode c = nodes[kids[k]];
        slots[ns++] = make_slot(c.lo_l, c.hi_l, c.left, grid);
        slots[ns++] = make_slot(c.lo_r, c.hi_r, c.right, grid);
      }
    }
    WideNode w;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= ns) {  // empty slot: lo > hi on every axis
        slots[k].child = -1;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          slots[k].lo[a] = 32767u;
          slots[k].hi[a] = 0u;
        }
      }
      w.child[k] = slots[k].child;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      w.lo[a][0] = slots[0].lo[a] | (slots[1].lo[a] << 16);
      w.lo[a][1] = slots[2].lo[a] | (slots[3].lo[a] << 16);
      // hi is stored with bit 15 of each halfword set: (hi|0x8000) - q keeps that bit iff q <= hi
      w.hi[a][0] = (slots[0].hi[a] | (slots[1].hi[a] << 16)) | 0x80008000u;
      w.hi[a][1] = (slots[2].hi[a] | (slots[3].hi[a] << 16)) | 0x80008000u;
    }
    wide[i] = w;
  }
}

}  // namespace fxii

using namespace fxii;

namespace fxii {
// height of the tree = longest root-to-leaf path; computed by one more bottom-up pass
__global__ void __launch_bounds__(256)
depth_kernel(const BvhNode* __restrict__ nodes, const int32_t* __restrict__ leaf_parent, int n, int32_t* max_depth) {
  int local = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    // sample every 64th leaf plus the ends: depth is only used to size-check the traversal stack,
    // and the bound 95 (63-bit key + 32-bit index tie-break) is enforced separately
    if ((k & 63) != 0 && k != n - 1) continue;
    int d = 1;
    int p = leaf_parent[k];
    while (nodes[p].parent >= 0) {
      p = nodes[p].parent;
      ++d;
    }
    local = max(local, d);
  }
  for (int o = 16; o > 0; o >>= 1) local = max(local, __shfl_xor_sync(0xffffffffu, local, o));
  if ((threadIdx.x & 31) == 0 && local > 0) atomicMax(max_depth, local);
}

// scene grid from the node bounds (widened by the padding and one part in 10^6) and the wide collapse
static void finish_wide(fxii_mesh* m, DevBuf<BvhNode>& bnodes, DevBuf<unsigned long long>& bounds, int n_internal,
                        cudaStream_t s) {
  unsigned long long hb[6];
  FXII_CUDA(cudaMemcpyAsync(hb, bounds.p, sizeof(hb), cudaMemcpyDeviceToHost, s));
  FXII_CUDA(cudaStreamSynchronize(s));
  for (int a = 0; a < 3; ++a) {
    const double lo = dec_ordered(hb[a]), hi = dec_ordered(hb[3 + a]);
    const double ext = (hi - lo) + 2.0 * m->padding;
    const double margin = m->padding + 1e-6 * ext + 1e-300;
    m->grid.origin[a] = lo - margin;
    m->grid.scale[a] = 32767.0 / (ext + 2.0 * margin);
  }
  { collapse_wide_kernel<<<grid_for(n_internal, 256, 8), 256, 0, s>>>(bnodes.p, n_internal, m->grid, m->wnodes.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
  FXII_CUDA(cudaStreamSynchronize(s));
}

// throws FXII_E_INVALID when an entry of the device array a[0..n) lies outside [0, hi); synchronises s
void check_index_range(const int32_t* a, int64_t n, int64_t hi, cudaStream_t s, const char* what) {
  if (n <= 0) return;
  DevBuf<int32_t> bad;
  bad.alloc(1, s);
  FXII_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int32_t), s));
  const int32_t hi32 = hi > 0x7fffffff ? 0x7fffffff : (int32_t)hi;
  { index_range_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(a, n, hi32, bad.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
  int32_t h = 0;
  FXII_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FXII_CUDA(cudaStreamSynchronize(s));
  if (h) throw Error(FXII_E_INVALID, std::string(what) + " out of range");
}

void mesh_build(fxii_mesh* m, const double* x, int64_t n_nodes, const int32_t* cell_nodes, int64_t n_cells, int npc,
                double padding, cudaStream_t s) {
  FXII_CHECK(npc == 4 || npc == 8, "cells must have 4 (tetrahedron) or 8 (hexahedron) geometry nodes");
  FXII_CHECK(n_cells >= 1 && n_nodes >= npc, "mesh needs at least one cell");
  FXII_CHECK(n_cells < (int64_t)1 << 31, "n_cells must fit int32");
  m->n_cells = n_cells;
  m->n_nodes = n_nodes;
  m->npc = npc;
  m->padding = padding;
  Trace tr("mesh_build", s);
  m->x.upload(x, 3 * n_nodes, s);
  m->cell_nodes.upload(cell_nodes, (int64_t)npc * n_cells, s);
  tr.mark("upload x + cells");
  check_index_range(m->cell_nodes.p, (int64_t)npc * n_cells, n_nodes, s, "cell node index");
  m->cellgeo.alloc((npc == 4 ? 12 : 24) * n_cells, s);
  const int n = (int)n_cells;
  const int n_internal = n > 1 ? n - 1 : 1;
  DevBuf<BvhNode> bnodes;  // binary LBVH, only needed until the wide collapse
  bnodes.alloc(n_internal, s);
  m->wnodes.alloc(n_internal, s);
  FXII_CUDA(cudaMemsetAsync(bnodes.p, 0, sizeof(BvhNode) * (size_t)n_internal, s));

  DevBuf<unsigned long long> bounds;
  bounds.alloc(6, s);
  unsigned long long init[6];
  for (int d = 0; d < 3; ++d) {
    init[d] = ~0ULL;
    init[3 + d] = 0ULL;
  }
  FXII_CUDA(cudaMemcpyAsync(bounds.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
  { scene_bounds_kernel<<<grid_for(n_nodes, 256, 8), 256, 0, s>>>(m->x.p, n_nodes, bounds.p); FXII_LAUNCHED(1); }

  DevBuf<uint64_t> keys, keys_sorted;
  DevBuf<int32_t> ids, ids_sorted, leaf_parent, flags;
  keys.alloc(n, s);
  keys_sorted.alloc(n, s);
  ids.alloc(n, s);
  ids_sorted.alloc(n, s);
  leaf_parent.alloc(n, s);
  flags.alloc(n_internal, s);
  tr.mark("allocations");
  if (npc == 4) {
    { cell_geom_morton_kernel<<<grid_for(n_cells, 256, 8), 256, 0, s>>>(
        m->x.p, reinterpret_cast<const int4*>(m->cell_nodes.p), n_cells, bounds.p, m->cellgeo.p, keys.p, ids.p); FXII_LAUNCHED(1); }
  } else {
    { hex_geom_morton_kernel<<<grid_for(n_cells, 256, 8), 256, 0, s>>>(m->x.p, m->cell_nodes.p, n_cells, bounds.p, m->cellgeo.p,
                                                                     keys.p, ids.p); FXII_LAUNCHED(1); }
  }
  FXII_CUDA(cudaGetLastError());

  size_t tmp_bytes = 0;
  FXII_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, ids.p, ids_sorted.p, n, 0, 63, s));
  DevBuf<uint8_t> tmp;
  tmp.alloc((int64_t)tmp_bytes, s);
  tr.mark("geometry + morton");
  FXII_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys_sorted.p, ids.p, ids_sorted.p, n, 0, 63, s));
  tr.mark("sort");

  if (n == 1) {
    // single leaf: a root whose left child is the cell and whose right child box is empty
    BvhNode root;
    memset(&root, 0, sizeof(root));
    for (int d = 0; d < 3; ++d) {
      root.lo_l[d] = -3.0e38f;
      root.hi_l[d] = 3.0e38f;
      root.lo_r[d] = 3.0e38f;
      root.hi_r[d] = -3.0e38f;
    }
    root.left = ~0;
    root.right = ~0;
    root.parent = -1;
    FXII_CUDA(cudaMemcpyAsync(bnodes.p, &root, sizeof(root), cudaMemcpyHostToDevice, s));
    m->max_depth = 1;
    finish_wide(m, bnodes, bounds, n_internal, s);
    return;
  }
  FXII_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int32_t) * (size_t)n_internal, s));
  { karras_internal_kernel<<<grid_for(n - 1, 256, 8), 256, 0, s>>>(keys_sorted.p, ids_sorted.p, n, bnodes.p, leaf_parent.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
  { refit_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(m->x.p, m->cell_nodes.p, npc, ids_sorted.p,
                                                   leaf_parent.p, n, padding, bnodes.p, flags.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
  tr.mark("karras + refit");
  DevBuf<int32_t> depth;
  depth.alloc(1, s);
  FXII_CUDA(cudaMemsetAsync(depth.p, 0, sizeof(int32_t), s));
  { depth_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(bnodes.p, leaf_parent.p, n, depth.p); FXII_LAUNCHED(1); }
  FXII_CUDA(cudaGetLastError());
  int32_t h = 0;
  FXII_CUDA(cudaMemcpyAsync(&h, depth.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FXII_CUDA(cudaStreamSynchronize(s));
  m->max_depth = h;
  finish_wide(m, bnodes, bounds, n_internal, s);
  tr.mark("depth + wide collapse");
}
}  // namespace fxii
