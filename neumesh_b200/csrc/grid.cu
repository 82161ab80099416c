// Spatial index over the mesh vertices and the exact 8-NN + mesh-distance kernels.
//
// Replaces the reference's use of the third-party FRNN package:
//   models/mesh_grid.py:64-74   grid construction (a V x V, K=32 self-query whose only kept result is `grid`)
//   models/mesh_grid.py:109-119 K=8 query  (r = 100 never binds => exact KNN, squared distances ascending)
//   models/mesh_grid.py:121-144 inverse-distance weights + indicator-blended signed distance
//
// B200-first design.  FRNN's uniform grid degenerates to one cell at r=100 (cell = r/2) and brute-forces all V
// vertices per query.  Here the vertices are Morton-sorted once and indexed by a sparse octree whose nodes carry
// TIGHT boxes; a query walks it depth-first, nearest child first, pruning against its current 8th-best distance.
// That is exact for any query position (the renderer probes the whole unit-sphere chord, far from the surface),
// needs no radius, and touches ~100-300 B of L2-resident nodes/points per level instead of 12*V bytes.
// One thread per query; a warp holds 32 neighbouring rays at the same sample index, so the walks are coherent and
// node/point loads are mostly L1 hits.  Distances use un-fused fp32 mul/add so that neighbour selection is
// bit-identical to an IEEE fp32 brute force (the oracle).
#include <cub/cub.cuh>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "grid.cuh"
#include "knn_coop.cuh"
#include "knn_walk.cuh"

namespace nmb {

// The directory start of the thread-per-query walk is compiled in only with -DNMB_KNN_DIRECTORY=1: measured on B200 it
// is bit-identical and not faster (profiles/r2_knn_directory_ab.txt), and merely carrying its code costs the bound scan
// 12 % (66 -> 75 registers per thread: 59.5 -> 67.6 ms per frame), so the shipped build leaves it out.
#ifndef NMB_KNN_DIRECTORY
#define NMB_KNN_DIRECTORY 0
#endif
#define NMB_GV_ARG(gv) (NMB_KNN_DIRECTORY ? &(gv) : nullptr)

// ------------------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

__global__ void bbox_kernel(const float* __restrict__ v, int64_t V, float* __restrict__ out /*6: min xyz, max xyz*/) {
  float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
  float hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < V; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = v[i * 3 + c];
      lo[c] = fminf(lo[c], x);
      hi[c] = fmaxf(hi[c], x);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
    // ordered-int trick: works for any sign
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int a = __float_as_int(lo[c]);
      a = a >= 0 ? a : a ^ 0x7fffffff;
      atomicMin(reinterpret_cast<int*>(out) + c, a);
      int b = __float_as_int(hi[c]);
      b = b >= 0 ? b : b ^ 0x7fffffff;
      atomicMax(reinterpret_cast<int*>(out) + 3 + c, b);
    }
  }
}

__global__ void morton_kernel(const float* __restrict__ v, int64_t V, float3 bmin, float inv_cell, int levels,
                              uint32_t* __restrict__ code, int32_t* __restrict__ iota) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= V) return;
  const uint32_t maxc = (1u << levels) - 1u;
  uint32_t q[3];
  const float b[3] = {bmin.x, bmin.y, bmin.z};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float t = (v[i * 3 + c] - b[c]) * inv_cell;
    int qi = (int)floorf(t);
    q[c] = (uint32_t)min(max(qi, 0), (int)maxc);
  }
  code[i] = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
  iota[i] = (int32_t)i;
}

__global__ void gather_points_kernel(const float* __restrict__ v, const int32_t* __restrict__ order, int64_t V,
                                     float4* __restrict__ pts, int32_t* __restrict__ inv) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= V) return;
  int32_t o = order[i];
  pts[i] = make_float4(v[(int64_t)o * 3 + 0], v[(int64_t)o * 3 + 1], v[(int64_t)o * 3 + 2], __int_as_float(o));
  inv[o] = (int32_t)i;
}

// level build: see DESIGN.md "octree build".  pnode[i] = id of the still-subdividing node holding point i, or -1.
__global__ void lvl_heads_kernel(const uint32_t* __restrict__ code, const int32_t* __restrict__ pnode, int64_t V,
                                 int shift, int32_t* __restrict__ head) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= V) return;
  int32_t p = pnode[i];
  int h = 0;
  if (p >= 0) h = (i == 0) || (pnode[i - 1] != p) || ((code[i] >> shift) != (code[i - 1] >> shift));
  head[i] = h;
}

__global__ void lvl_create_kernel(const uint32_t* __restrict__ code, const int32_t* __restrict__ pnode,
                                  const int32_t* __restrict__ head, const int32_t* __restrict__ cid_excl, int64_t V,
                                  int shift, int32_t lvl_off, int32_t* __restrict__ nbegin, int32_t* __restrict__ nend,
                                  int32_t* __restrict__ nfirst, int32_t* __restrict__ nlast,
                                  int32_t* __restrict__ newnode) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= V) return;
  int32_t p = pnode[i];
  if (p < 0) {
    newnode[i] = -1;
    return;
  }
  int32_t gid = lvl_off + cid_excl[i] + head[i] - 1;
  newnode[i] = gid;
  if (head[i]) {
    nbegin[gid] = (int32_t)i;
    nfirst[gid] = -1;
    nlast[gid] = -1;
  }
  bool last = (i + 1 == V) || (pnode[i + 1] != p) || ((code[i + 1] >> shift) != (code[i] >> shift));
  if (last) nend[gid] = (int32_t)(i + 1);
  if ((int32_t)i == nbegin[p]) nfirst[p] = gid;
  if ((int32_t)(i + 1) == nend[p]) nlast[p] = gid;
}

__global__ void lvl_activate_kernel(const int32_t* __restrict__ newnode, const int32_t* __restrict__ nbegin,
                                    const int32_t* __restrict__ nend, int64_t V, int can_split, int leaf_max,
                                    int32_t* __restrict__ pnode) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= V) return;
  int32_t g = newnode[i];
  int32_t r = -1;
  if (g >= 0 && can_split && (nend[g] - nbegin[g]) > leaf_max) r = g;
  pnode[i] = r;
}

__global__ void lvl_boxes_kernel(int32_t first_node, int32_t n_nodes, const int32_t* __restrict__ nbegin,
                                 const int32_t* __restrict__ nend, const int32_t* __restrict__ nfirst,
                                 const int32_t* __restrict__ nlast, const float4* __restrict__ pts,
                                 float4* __restrict__ nodes) {
  int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes) return;
  int32_t n = first_node + t;
  float3 lo = make_float3(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
  float3 hi = make_float3(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
  int32_t fc = nfirst[n];
  float link, cnt;
  if (fc < 0) {  // leaf: box of its points
    int32_t b = nbegin[n], e = nend[n];
    for (int32_t i = b; i < e; ++i) {
      float4 p = pts[i];
      lo.x = fminf(lo.x, p.x); lo.y = fminf(lo.y, p.y); lo.z = fminf(lo.z, p.z);
      hi.x = fmaxf(hi.x, p.x); hi.y = fmaxf(hi.y, p.y); hi.z = fmaxf(hi.z, p.z);
    }
    link = __int_as_float(b);
    cnt = __int_as_float(-(e - b));
  } else {  // internal: union of the (already final) child boxes
    int32_t lc = nlast[n];
    for (int32_t c = fc; c <= lc; ++c) {
      float4 a = nodes[NODE_F4 * c], b = nodes[NODE_F4 * c + 1];
      lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
      hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
    }
    link = __int_as_float(fc);
    cnt = __int_as_float(lc - fc + 1);
  }
  nodes[NODE_F4 * n] = make_float4(lo.x, lo.y, lo.z, link);
  nodes[NODE_F4 * n + 1] = make_float4(hi.x, hi.y, hi.z, cnt);

  // ---- disc bound: centre c, radius r, axis u, half-thickness t (see knn_walk) ----
  const int32_t b = nbegin[n], e = nend[n];
  const int32_t cntp = e - b;
  float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
  float ux = 0.f, uy = 0.f, uz = 1.f, r, th;
  if (cntp > DISC_MAX_POINTS) {
    const float ex = hi.x - cx, ey = hi.y - cy, ez = hi.z - cz;
    r = sqrtf(ex * ex + ey * ey + ez * ez);
    th = r;
  } else {
    // centroid
    double sx = 0, sy = 0, sz = 0;
    for (int32_t i = b; i < e; ++i) {
      const float4 p = pts[i];
      sx += p.x; sy += p.y; sz += p.z;
    }
    cx = (float)(sx / cntp); cy = (float)(sy / cntp); cz = (float)(sz / cntp);
    // covariance
    float a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    for (int32_t i = b; i < e; ++i) {
      const float4 p = pts[i];
      const float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
      a00 += dx * dx; a01 += dx * dy; a02 += dx * dz; a11 += dy * dy; a12 += dy * dz; a22 += dz * dz;
    }
    // smallest-variance axis = dominant eigenvector of (trace * I - A): power iteration from 3 starts
    const float tr = a00 + a11 + a22;
    if (tr > 0.f) {
      const float m00 = tr - a00, m11 = tr - a11, m22 = tr - a22;
      float best = -1.f;
      for (int s0 = 0; s0 < 3; ++s0) {
        float vx = s0 == 0, vy = s0 == 1, vz = s0 == 2;
        for (int it = 0; it < 24; ++it) {
          const float wx = m00 * vx - a01 * vy - a02 * vz;
          const float wy = -a01 * vx + m11 * vy - a12 * vz;
          const float wz = -a02 * vx - a12 * vy + m22 * vz;
          const float nn = sqrtf(wx * wx + wy * wy + wz * wz);
          if (!(nn > 0.f)) break;
          vx = wx / nn; vy = wy / nn; vz = wz / nn;
        }
        // Rayleigh quotient of (trace I - A): larger is better
        const float q = vx * (m00 * vx - a01 * vy - a02 * vz) + vy * (-a01 * vx + m11 * vy - a12 * vz) +
                        vz * (-a02 * vx - a12 * vy + m22 * vz);
        if (q > best) {
          best = q; ux = vx; uy = vy; uz = vz;
        }
      }
      const float un = sqrtf(ux * ux + uy * uy + uz * uz);
      if (un > 0.5f) { ux /= un; uy /= un; uz /= un; } else { ux = 0.f; uy = 0.f; uz = 1.f; }
    }
    r = 0.f;
    th = 0.f;
    for (int32_t i = b; i < e; ++i) {
      const float4 p = pts[i];
      const float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
      r = fmaxf(r, sqrtf(dx * dx + dy * dy + dz * dz));
      th = fmaxf(th, fabsf(ux * dx + uy * dy + uz * dz));
    }
  }
  // inflate: the bound is evaluated in fp32 and must never exceed the true distance to any point of the node
  r = r * 1.00001f + 1e-7f;
  th = th * 1.00001f + 1e-7f;
  nodes[NODE_F4 * n + 2] = make_float4(cx, cy, cz, r);
  nodes[NODE_F4 * n + 3] = make_float4(ux, uy, uz, th);
}


// Directory tables of the cooperative walk (knn_coop.cuh).  One thread per octree node of depth `depth`: a node AT a
// directory level fills its own cell; a LEAF above a directory level fills every cell it covers at that level.
__global__ void dir_fill_kernel(int32_t first_node, int32_t n_nodes, int depth, int L, const uint32_t* __restrict__ code,
                                const int32_t* __restrict__ nbegin, const int32_t* __restrict__ nfirst, int lmin, int lmax,
                                const int32_t* __restrict__ dir_off /*[lmax - lmin + 1] on device*/,
                                int32_t* __restrict__ dir) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes) return;
  const int32_t n = first_node + t;
  const uint32_t prefix = depth == 0 ? 0u : (code[nbegin[n]] >> (3 * (L - depth)));
  const bool leaf = nfirst[n] < 0;
  for (int l = max(lmin, depth); l <= lmax; ++l) {
    if (l == depth) {
      dir[dir_off[l - lmin] + prefix] = n;
    } else if (leaf) {
      const int sh = 3 * (l - depth);
      const uint32_t b = prefix << sh, e = (prefix + 1u) << sh;
      for (uint32_t c = b; c < e; ++c) dir[dir_off[l - lmin] + c] = n;
    }
  }
}

static int build_grid(const float* vertices, int64_t V, cudaStream_t stream, nmb_grid* g) {
  NMB_CHECK(V >= KNN_K, "mesh needs at least 8 vertices");
  NMB_CHECK(V < (int64_t(1) << 30), "too many vertices");
  g->V = V;
  const int threads = 256;
  const int64_t blocks = ceil_div(V, threads);

  // bounding cube
  StreamBuf bb_buf;
  NMB_CUDA_OK(bb_buf.alloc(6 * sizeof(float), stream));
  struct { float* p; } bb{bb_buf.as<float>()};
  {
    int init[6] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000, (int)0x80800000, (int)0x80800000};
    // ordered-int encodings of +FLT_MAX / -FLT_MAX
    init[3] = init[4] = init[5] = (int)(0xff7fffffu ^ 0x7fffffffu);
    NMB_CUDA_OK(cudaMemcpyAsync(bb.p, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    bbox_kernel<<<(unsigned)(blocks < 1024 ? blocks : 1024), threads, 0, stream>>>(vertices, V, bb.p);
    NMB_LAUNCH_OK();
    int raw[6];
    NMB_CUDA_OK(cudaMemcpyAsync(raw, bb.p, sizeof(raw), cudaMemcpyDeviceToHost, stream));
    NMB_CUDA_OK(cudaStreamSynchronize(stream));
    float lo[3], hi[3];
    for (int c = 0; c < 3; ++c) {
      int a = raw[c];
      a = a >= 0 ? a : a ^ 0x7fffffff;
      int b = raw[3 + c];
      b = b >= 0 ? b : b ^ 0x7fffffff;
      memcpy(&lo[c], &a, 4);
      memcpy(&hi[c], &b, 4);
    }
    float side = 0.f;
    for (int c = 0; c < 3; ++c) side = fmaxf(side, hi[c] - lo[c]);
    NMB_CHECK(side == side && side < 1e30f, "non-finite vertex coordinates");
    side = side * 1.0001f + 1e-6f;
    for (int c = 0; c < 3; ++c) g->bmin[c] = lo[c];
    // depth: aim at ~2-4 points per finest cell for a surface-like point set (#cells ~ 4^L)
    int L = 1;
    while (L < 10 && (double)V / pow(4.0, L) > 2.0) ++L;
    g->levels = L;
    g->inv_cell = (float)(1u << L) / side;
  }
  const int L = g->levels;

  // Morton sort
  // build temporaries are stream-ordered scratch (cached in the device pool: a rebuild allocates nothing new)
  StreamBuf code_in_b, code_b, iota_b;
  NMB_CUDA_OK(code_in_b.alloc(sizeof(uint32_t) * V, stream));
  NMB_CUDA_OK(code_b.alloc(sizeof(uint32_t) * V, stream));
  NMB_CUDA_OK(iota_b.alloc(sizeof(int32_t) * V, stream));
  struct { uint32_t* p; } code_in{code_in_b.as<uint32_t>()}, code{code_b.as<uint32_t>()};
  struct { int32_t* p; } iota{iota_b.as<int32_t>()};
  NMB_CUDA_OK(g->order.alloc(V));
  NMB_CUDA_OK(g->inv.alloc(V));
  NMB_CUDA_OK(g->pts.alloc(V));
  morton_kernel<<<(unsigned)blocks, threads, 0, stream>>>(vertices, V, make_float3(g->bmin[0], g->bmin[1], g->bmin[2]),
                                                          g->inv_cell, L, code_in.p, iota.p);
  NMB_LAUNCH_OK();
  {
    size_t tmp_bytes = 0;
    NMB_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, code_in.p, code.p, iota.p, g->order.p, (int)V, 0,
                                                3 * L, stream));
    StreamBuf tmp;
    NMB_CUDA_OK(tmp.alloc(tmp_bytes, stream));
    NMB_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, code_in.p, code.p, iota.p, g->order.p, (int)V, 0,
                                                3 * L, stream));
    count_launch(4);
  }
  gather_points_kernel<<<(unsigned)blocks, threads, 0, stream>>>(vertices, g->order.p, V, g->pts.p, g->inv.p);
  NMB_LAUNCH_OK();

  // level-by-level subdivision
  // leaf capacity (points per leaf): tunable for experiments, default LEAF_MAX
  const int leaf_max = getenv("NMB_LEAF_MAX") ? std::max(1, atoi(getenv("NMB_LEAF_MAX"))) : LEAF_MAX;
  const int64_t cap = V + (int64_t)(L + 1) * (V / (leaf_max + 1) + 1) + 16;
  StreamBuf node_b, point_b;   // 4 node arrays of `cap` entries; 4 per-point arrays
  NMB_CUDA_OK(node_b.alloc(sizeof(int32_t) * 4 * cap, stream));
  NMB_CUDA_OK(point_b.alloc(sizeof(int32_t) * (4 * V + 4), stream));
  struct I32 { int32_t* p; };
  I32 nbegin{node_b.as<int32_t>()}, nend{nbegin.p + cap}, nfirst{nend.p + cap}, nlast{nfirst.p + cap};
  I32 pnode{point_b.as<int32_t>()}, newnode{pnode.p + V}, head{newnode.p + V}, cid{head.p + V};
  size_t scan_bytes = 0;
  NMB_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, head.p, cid.p, (int)V, stream));
  StreamBuf scan_tmp;
  NMB_CUDA_OK(scan_tmp.alloc(scan_bytes, stream));

  std::vector<int32_t> lvl_off;  // first node id of each level
  lvl_off.push_back(0);
  {
    int32_t root[4] = {0, (int32_t)V, -1, -1};
    NMB_CUDA_OK(cudaMemcpyAsync(nbegin.p, &root[0], 4, cudaMemcpyHostToDevice, stream));
    NMB_CUDA_OK(cudaMemcpyAsync(nend.p, &root[1], 4, cudaMemcpyHostToDevice, stream));
    NMB_CUDA_OK(cudaMemcpyAsync(nfirst.p, &root[2], 4, cudaMemcpyHostToDevice, stream));
    NMB_CUDA_OK(cudaMemcpyAsync(nlast.p, &root[3], 4, cudaMemcpyHostToDevice, stream));
    // all points start in the root (V >= 8 > ... root splits iff V > LEAF_MAX and L > 0)
    int fill = (V > leaf_max && L > 0) ? 0 : -1;
    NMB_CUDA_OK(cudaMemsetAsync(pnode.p, fill == 0 ? 0 : 0xff, sizeof(int32_t) * V, stream));
  }
  int32_t n_nodes = 1;
  lvl_off.push_back(1);
  for (int l = 0; l < L; ++l) {
    const int shift = 3 * (L - (l + 1));
    lvl_heads_kernel<<<(unsigned)blocks, threads, 0, stream>>>(code.p, pnode.p, V, shift, head.p);
    NMB_LAUNCH_OK();
    NMB_CUDA_OK(cub::DeviceScan::ExclusiveSum(scan_tmp.p, scan_bytes, head.p, cid.p, (int)V, stream));
    count_launch(2);
    int32_t last_cid = 0, last_head = 0;
    NMB_CUDA_OK(cudaMemcpyAsync(&last_cid, cid.p + (V - 1), 4, cudaMemcpyDeviceToHost, stream));
    NMB_CUDA_OK(cudaMemcpyAsync(&last_head, head.p + (V - 1), 4, cudaMemcpyDeviceToHost, stream));
    NMB_CUDA_OK(cudaStreamSynchronize(stream));
    const int32_t n_new = last_cid + last_head;
    if (n_new == 0) break;
    NMB_CHECK((int64_t)n_nodes + n_new <= cap, "octree node capacity exceeded");
    lvl_create_kernel<<<(unsigned)blocks, threads, 0, stream>>>(code.p, pnode.p, head.p, cid.p, V, shift, n_nodes,
                                                                 nbegin.p, nend.p, nfirst.p, nlast.p, newnode.p);
    NMB_LAUNCH_OK();
    lvl_activate_kernel<<<(unsigned)blocks, threads, 0, stream>>>(newnode.p, nbegin.p, nend.p, V, (l + 1 < L) ? 1 : 0,
                                                                   leaf_max, pnode.p);
    NMB_LAUNCH_OK();
    n_nodes += n_new;
    lvl_off.push_back(n_nodes);
  }
  g->num_nodes = n_nodes;
  g->lvl_off = lvl_off;
  NMB_CUDA_OK(g->nodes.alloc(NODE_F4 * (int64_t)n_nodes));
  for (int l = (int)lvl_off.size() - 2; l >= 0; --l) {
    const int32_t first = lvl_off[l], cnt = lvl_off[l + 1] - lvl_off[l];
    if (cnt <= 0) continue;
    lvl_boxes_kernel<<<(unsigned)ceil_div(cnt, threads), threads, 0, stream>>>(first, cnt, nbegin.p, nend.p, nfirst.p,
                                                                               nlast.p, g->pts.p, g->nodes.p);
    NMB_LAUNCH_OK();
  }
  // directory tables: levels [3, min(L - 1, 7)] (finer cells than the vertex spacing buy nothing).  OPT-IN
  // (build with -DNMB_KNN_DIRECTORY=1 and set NMB_KNN_DIR=1; the cooperative kernels, NMB_KNN_COOP=1, use it too): measured on B200 the directory start is bit-identical but not faster (knn 130.2 -> 133.7 ms, live
  // lists 43.4 -> 48.3 ms per 800x800 frame, profiles/r2_knn_directory_ab.txt): with a warm bound the ball meets only 1-2
  // children per TOP level, so the levels it skips cost about as much as the 2x2x2-cell seeding does; the expansions that
  // dominate a walk sit at the bottom levels, where cells are as small as the ball.
  g->dir_lmin = 3;
  g->dir_lmax = std::min(L - 1, 7);
  if (getenv("NMB_KNN_DIR_MAX")) g->dir_lmax = std::min(g->dir_lmax, atoi(getenv("NMB_KNN_DIR_MAX")));
  static const bool want_dir = (NMB_KNN_DIRECTORY && getenv("NMB_KNN_DIR") != nullptr) || getenv("NMB_KNN_COOP") != nullptr;
  if (!want_dir) g->dir_lmax = g->dir_lmin - 1;
  if (g->dir_lmax >= g->dir_lmin) {
    int64_t total = 0;
    int32_t offs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int l = g->dir_lmin; l <= g->dir_lmax; ++l) {
      offs[l - g->dir_lmin] = (int32_t)total;
      g->dir_off[l - g->dir_lmin] = (int32_t)total;
      total += int64_t(1) << (3 * l);
    }
    NMB_CUDA_OK(g->dir.alloc(total));
    NMB_CUDA_OK(cudaMemsetAsync(g->dir.p, 0xff, sizeof(int32_t) * total, stream));
    StreamBuf offs_dev;
    NMB_CUDA_OK(offs_dev.alloc(sizeof(offs), stream));
    NMB_CUDA_OK(cudaMemcpyAsync(offs_dev.p, offs, sizeof(offs), cudaMemcpyHostToDevice, stream));
    for (int dpt = 0; dpt + 1 < (int)lvl_off.size() && dpt <= g->dir_lmax; ++dpt) {
      const int32_t first = lvl_off[dpt], cnt = lvl_off[dpt + 1] - lvl_off[dpt];
      if (cnt <= 0) continue;
      dir_fill_kernel<<<(unsigned)ceil_div(cnt, threads), threads, 0, stream>>>(
          first, cnt, dpt, L, code.p, nbegin.p, nfirst.p, g->dir_lmin, g->dir_lmax, offs_dev.as<int32_t>(), g->dir.p);
      NMB_LAUNCH_OK();
    }
    NMB_CUDA_OK(cudaStreamSynchronize(stream));   // `offs` is a stack array
  }
  NMB_CUDA_OK(cudaStreamSynchronize(stream));
  return 0;
}

GridView make_view(const nmb_grid* g) {
  GridView v{};
  v.nodes = g->nodes.p;
  v.pts = g->pts.p;
  static const bool no_dir = getenv("NMB_KNN_NO_DIR") != nullptr;
  v.dir = (g->dir_lmax >= g->dir_lmin && !no_dir) ? g->dir.p : nullptr;
  v.dir_lmin = g->dir_lmin;
  v.dir_lmax = v.dir ? g->dir_lmax : g->dir_lmin - 1;
  v.bmin[0] = g->bmin[0];
  v.bmin[1] = g->bmin[1];
  v.bmin[2] = g->bmin[2];
  v.inv_cell = g->inv_cell;
  v.levels = g->levels;
  return v;
}

__device__ __forceinline__ void load_query(const PointSrc& src, int64_t p, float& qx, float& qy, float& qz) {
  if (src.xyz) {
    qx = src.xyz[p * 3 + 0];
    qy = src.xyz[p * 3 + 1];
    qz = src.xyz[p * 3 + 2];
  } else {
    const int64_t r = p % src.R;
    const float z = src.z[p];
    // pts = rays_o + z * rays_d (renderer.py:85,202,248,264,267): separate mul and add as in torch
    qx = __fadd_rn(src.rays_o[r * 3 + 0], __fmul_rn(z, src.rays_d[r * 3 + 0]));
    qy = __fadd_rn(src.rays_o[r * 3 + 1], __fmul_rn(z, src.rays_d[r * 3 + 1]));
    qz = __fadd_rn(src.rays_o[r * 3 + 2], __fmul_rn(z, src.rays_d[r * 3 + 2]));
  }
}

// mesh_grid.py:121-144 for one query whose neighbours are known.
__device__ __forceinline__ void mesh_distance_point(const float4* __restrict__ pts,
                                                    const float4* __restrict__ indicator, float w1, float qx,
                                                    float qy, float qz, const float (&d2)[KNN_K],
                                                    const int32_t (&ix)[KNN_K], float (&w)[KNN_K], float& ds,
                                                    float (&grad)[3]) {
  float wsum = 0.f;
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) {
    w[k] = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(d2[k]), 1e-7f));  // :123-124
    wsum = __fadd_rn(wsum, w[k]);
  }
  ds = 0.f;
  grad[0] = grad[1] = grad[2] = 0.f;
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) {
    w[k] = __fdiv_rn(w[k], wsum);  // :125
    const float4 p = __ldg(&pts[ix[k]]);
    const float4 nv = __ldg(&indicator[ix[k]]);
    const float vx = __fsub_rn(qx, p.x), vy = __fsub_rn(qy, p.y), vz = __fsub_rn(qz, p.z);  // :134
    const float rho = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
    const float D = __fadd_rn(w1, rho);
    const float mx = __fdiv_rn(__fadd_rn(__fmul_rn(nv.x, w1), __fmul_rn(vx, rho)), D);  // :136
    const float my = __fdiv_rn(__fadd_rn(__fmul_rn(nv.y, w1), __fmul_rn(vy, rho)), D);
    const float mz = __fdiv_rn(__fadd_rn(__fmul_rn(nv.z, w1), __fmul_rn(vz, rho)), D);
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(vx, mx), __fmul_rn(vy, my)), __fmul_rn(vz, mz));
    ds = __fadd_rn(ds, __fmul_rn(w[k], dot));  // :137-142
    // d(dot)/dx = (w1 n + 3 rho v) / D - dot * v / (rho D)   (norm's sub-gradient at rho = 0 is 0)
    const float invD = 1.0f / D;
    const float c2 = rho > 0.f ? dot / (rho * D) : 0.f;
    grad[0] += w[k] * ((w1 * nv.x + 3.f * rho * vx) * invD - c2 * vx);
    grad[1] += w[k] * ((w1 * nv.y + 3.f * rho * vy) * invD - c2 * vy);
    grad[2] += w[k] * ((w1 * nv.z + 3.f * rho * vz) * invD - c2 * vz);
  }
}

__global__ void __launch_bounds__(128)
knn_distance_kernel(const float4* __restrict__ nodes, const float4* __restrict__ pts,
                    const float4* __restrict__ indicator, float w1, PointSrc src, int64_t P, KnnOut out) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= P) return;
  float qx, qy, qz;
  load_query(src, p, qx, qy, qz);
  float d2[KNN_K];
  int32_t ix[KNN_K];
  knn_walk<KNN_K, false>(nodes, pts, qx, qy, qz, d2, ix);
  float w[KNN_K], ds, grad[3];
  mesh_distance_point(pts, indicator, w1, qx, qy, qz, d2, ix, w, ds, grad);
  out.ds[p] = ds;
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) {
    out.slot[k * out.stride + p] = ix[k];
    out.w[k * out.stride + p] = w[k];
  }
  if (out.grad) {
    out.grad[0 * out.stride + p] = grad[0];
    out.grad[1 * out.stride + p] = grad[1];
    out.grad[2 * out.stride + p] = grad[2];
  }
}

// Ray-ordered variant: one thread per RAY walks its S samples in depth order and warm-starts every query with the
// previous sample's neighbours (consecutive samples are <~0.03 apart, so the initial 8th-best bound is already within
// a few percent of the final one and the octree walk prunes almost everything).  A warp = 32 neighbouring rays.
template <int MINB, int ORDER>
__global__ void __launch_bounds__(128, MINB)
knn_rays_kernel(const float4* __restrict__ nodes, const float4* __restrict__ pts, const float4* __restrict__ indicator,
                float w1, PointSrc src, int S, int seg, KnnOut out, const GridView gv) {
  // thread t handles samples [g * seg, (g + 1) * seg) of ray r, t = g * R + r: with few rays (multi-GPU shards) a
  // ray's samples are split over several threads so that the launch still fills the GPU (one cold walk per segment)
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t r = t % src.R;
  const int s_begin = (int)(t / src.R) * seg;
  if (s_begin >= S) return;
  const int s_end = min(s_begin + seg, S);
  const float ox = src.rays_o[r * 3 + 0], oy = src.rays_o[r * 3 + 1], oz = src.rays_o[r * 3 + 2];
  const float dx = src.rays_d[r * 3 + 0], dy = src.rays_d[r * 3 + 1], dz = src.rays_d[r * 3 + 2];
  float d2[KNN_K];
  int32_t ix[KNN_K];
  for (int s = s_begin; s < s_end; ++s) {
    const int64_t p = (int64_t)s * src.R + r;
    const float z = src.z[p];
    const float qx = __fadd_rn(ox, __fmul_rn(z, dx));
    const float qy = __fadd_rn(oy, __fmul_rn(z, dy));
    const float qz = __fadd_rn(oz, __fmul_rn(z, dz));
    if (s == s_begin && s_begin == 0 && src.seed_slot != nullptr) {
      // warm start from an earlier query of this ray (see PointSrc::seed_slot)
      const int64_t sp0 = (int64_t)src.seed_entry[r] * src.R + r;
#pragma unroll
      for (int k = 0; k < KNN_K; ++k) ix[k] = src.seed_slot[k * src.seed_stride + sp0];
      warm_rerank<KNN_K>(pts, qx, qy, qz, d2, ix);
      knn_walk<KNN_K, true, ORDER>(nodes, pts, qx, qy, qz, d2, ix, NMB_GV_ARG(gv));
    } else if (s == s_begin) {
      knn_walk<KNN_K, false, ORDER>(nodes, pts, qx, qy, qz, d2, ix);
    } else {
      warm_rerank<KNN_K>(pts, qx, qy, qz, d2, ix);
      knn_walk<KNN_K, true, ORDER>(nodes, pts, qx, qy, qz, d2, ix, NMB_GV_ARG(gv));
    }
    float w[KNN_K], ds, grad[3];
    mesh_distance_point(pts, indicator, w1, qx, qy, qz, d2, ix, w, ds, grad);
    out.ds[p] = ds;
#pragma unroll
    for (int k = 0; k < KNN_K; ++k) {
      out.slot[k * out.stride + p] = ix[k];
      out.w[k * out.stride + p] = w[k];
    }
    if (out.grad) {
      out.grad[0 * out.stride + p] = grad[0];
      out.grad[1 * out.stride + p] = grad[1];
      out.grad[2 * out.stride + p] = grad[2];
    }
  }
}

// Per-ray lists of explicit points (the compacted live samples of nmb_render): thread r walks its entries
// [off[r], off[r] + cnt[r]) in order - they are consecutive samples of one ray - warm-starting each query with the
// previous one's neighbours.
__global__ void __launch_bounds__(128)
knn_lists_kernel(const float4* __restrict__ nodes, const float4* __restrict__ pts, const float4* __restrict__ indicator,
                 float w1, const float* __restrict__ xyz, const int32_t* __restrict__ off,
                 const int32_t* __restrict__ cnt, int64_t R, int seg, int max_seg, KnnOut out, const GridView gv) {
  // thread t = g * R + r handles entries [g * seg, (g + 1) * seg) of ray r's list (short serial chains)
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t r = t % R;
  const int gseg = (int)(t / R);
  if (gseg >= max_seg) return;
  const int64_t b = off[r];
  const int j_begin = gseg * seg;
  const int n = min(cnt[r], j_begin + seg);
  float d2[KNN_K];
  int32_t ix[KNN_K];
  for (int j = j_begin; j < n; ++j) {
    const int64_t p = b + j;
    const float qx = xyz[p * 3], qy = xyz[p * 3 + 1], qz = xyz[p * 3 + 2];
    if (j == j_begin) {
      knn_walk<KNN_K, false>(nodes, pts, qx, qy, qz, d2, ix);
    } else {
      warm_rerank<KNN_K>(pts, qx, qy, qz, d2, ix);
      knn_walk<KNN_K, true>(nodes, pts, qx, qy, qz, d2, ix, NMB_GV_ARG(gv));
    }
    float w[KNN_K], ds, grad[3];
    mesh_distance_point(pts, indicator, w1, qx, qy, qz, d2, ix, w, ds, grad);
    out.ds[p] = ds;
#pragma unroll
    for (int k = 0; k < KNN_K; ++k) {
      out.slot[k * out.stride + p] = ix[k];
      out.w[k * out.stride + p] = w[k];
    }
    if (out.grad) {
      out.grad[0 * out.stride + p] = grad[0];
      out.grad[1 * out.stride + p] = grad[1];
      out.grad[2 * out.stride + p] = grad[2];
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// group-cooperative kernels (knn_coop.cuh): 8 lanes per query chain, 16 chains per 128-thread block
// ------------------------------------------------------------------------------------------------------------
static bool knn_legacy() {
  // Default: thread-per-query kernels (with the directory start).  NMB_KNN_COOP=1 selects the 8-lanes-per-query
  // cooperative kernels instead: bit-identical results, measured 2.4x SLOWER on B200 (profiles/r2_ncu_knn_coop_summary.csv:
  // 2.15x the warp instructions per query - shuffles, ranking, serial insertions - at 16 of 32 active lanes and 29
  // resident warps), kept as the record of that experiment and as an independent implementation for cross-checks.
  static const bool v = getenv("NMB_KNN_COOP") == nullptr;
  return v;
}

#define NMB_COOP_PROLOGUE()                                                                    \
  __shared__ uint32_t stack_mem[coop::GROUPS_PER_BLOCK * coop::STACK_WORDS];                   \
  const coop::Lane ln = coop::make_lane();                                                     \
  uint32_t* stk = stack_mem + (threadIdx.x / coop::G) * coop::STACK_WORDS;                     \
  const int32_t root_link = __float_as_int(__ldg(&gv.nodes[0]).w);                             \
  const int32_t root_cnt = __float_as_int(__ldg(&gv.nodes[1]).w);                              \
  const int64_t chain = blockIdx.x * (int64_t)coop::GROUPS_PER_BLOCK + threadIdx.x / coop::G;

// explicit points or ray samples in any order: one cold query per group
template <int MINB>
__global__ void __launch_bounds__(coop::BLOCK, MINB)
knn_points_coop_kernel(GridView gv, const float4* __restrict__ indicator, float w1, PointSrc src, int64_t P, KnnOut out) {
  NMB_COOP_PROLOGUE()
  if (chain >= P) return;
  float qx, qy, qz;
  load_query(src, chain, qx, qy, qz);
  float d;
  int32_t ix;
  coop::query(gv, ln, stk, root_link, root_cnt, indicator, w1, qx, qy, qz, false, d, ix, out, chain);
}

// ray-ordered: chain t = g * R + r walks samples [g * seg, (g + 1) * seg) of ray r in depth order, every query
// warm-started with the previous sample's neighbours; the 4 chains of a warp are 4 neighbouring rays
template <int MINB>
__global__ void __launch_bounds__(coop::BLOCK, MINB)
knn_rays_coop_kernel(GridView gv, const float4* __restrict__ indicator, float w1, PointSrc src, int S, int seg, KnnOut out) {
  NMB_COOP_PROLOGUE()
  const int64_t r = chain % src.R;
  const int s_begin = (int)(chain / src.R) * seg;
  if (s_begin >= S) return;
  const int s_end = min(s_begin + seg, S);
  const float ox = src.rays_o[r * 3 + 0], oy = src.rays_o[r * 3 + 1], oz = src.rays_o[r * 3 + 2];
  const float dx = src.rays_d[r * 3 + 0], dy = src.rays_d[r * 3 + 1], dz = src.rays_d[r * 3 + 2];
  float d;
  int32_t ix;
  for (int s = s_begin; s < s_end; ++s) {
    const int64_t p = (int64_t)s * src.R + r;
    const float z = src.z[p];
    const float qx = __fadd_rn(ox, __fmul_rn(z, dx));
    const float qy = __fadd_rn(oy, __fmul_rn(z, dy));
    const float qz = __fadd_rn(oz, __fmul_rn(z, dz));
    coop::query(gv, ln, stk, root_link, root_cnt, indicator, w1, qx, qy, qz, s != s_begin, d, ix, out, p);
  }
}

// per-ray lists of explicit points (the compacted live samples of nmb_render): chain t = g * R + r handles entries
// [g * seg, (g + 1) * seg) of ray r's list
template <int MINB>
__global__ void __launch_bounds__(coop::BLOCK, MINB)
knn_lists_coop_kernel(GridView gv, const float4* __restrict__ indicator, float w1, const float* __restrict__ xyz,
                      const int32_t* __restrict__ off, const int32_t* __restrict__ cnt, int64_t R, int seg, int max_seg,
                      KnnOut out) {
  NMB_COOP_PROLOGUE()
  const int64_t r = chain % R;
  const int gseg = (int)(chain / R);
  if (gseg >= max_seg) return;
  const int64_t b = off[r];
  const int j_begin = gseg * seg;
  const int n = min(cnt[r], j_begin + seg);
  float d;
  int32_t ix;
  for (int j = j_begin; j < n; ++j) {
    const int64_t p = b + j;
    const float qx = xyz[p * 3], qy = xyz[p * 3 + 1], qz = xyz[p * 3 + 2];
    coop::query(gv, ln, stk, root_link, root_cnt, indicator, w1, qx, qy, qz, j != j_begin, d, ix, out, p);
  }
}

// bounded near / far, every sample evaluated (small launches): one cold query per group
template <int MINB>
__global__ void __launch_bounds__(coop::BLOCK, MINB)
bound_scan_coop_kernel(GridView gv, const float4* __restrict__ indicator, float w1, const float* __restrict__ rays_o,
                       const float* __restrict__ dirs, const float* __restrict__ near, const float* __restrict__ far,
                       int64_t R, int n_grid, float thresh, int32_t* __restrict__ bnear, int32_t* __restrict__ bfar) {
  NMB_COOP_PROLOGUE()
  if (chain >= R * n_grid) return;
  const int64_t r = chain % R;
  const int s = (int)(chain / R);
  const float t = linspace01(s, n_grid);
  const float dep = __fadd_rn(__fmul_rn(near[r], __fsub_rn(1.0f, t)), __fmul_rn(far[r], t));  // renderer.py:81
  const float qx = __fadd_rn(rays_o[r * 3 + 0], __fmul_rn(dep, dirs[r * 3 + 0]));
  const float qy = __fadd_rn(rays_o[r * 3 + 1], __fmul_rn(dep, dirs[r * 3 + 1]));
  const float qz = __fadd_rn(rays_o[r * 3 + 2], __fmul_rn(dep, dirs[r * 3 + 2]));
  float d;
  int32_t ix;
  const float ds = coop::query(gv, ln, stk, root_link, root_cnt, indicator, w1, qx, qy, qz, false, d, ix, KnnOut{}, 0);
  if (ds < thresh && ln.gl == 0) {
    atomicMin(&bnear[r], __float_as_int(dep));
    atomicMax(&bfar[r], __float_as_int(dep));
  }
}

// ray-ordered bounded near / far with early exit and the shell certificate (same logic as bound_rays_kernel below):
// chain t = g * R + r scans samples [g * BOUND_SEG, (g + 1) * BOUND_SEG) of ray r from the front to its first hit and
// from the back to its last hit
constexpr int BOUND_SEG_COOP = 32;
template <int MINB>
__global__ void __launch_bounds__(coop::BLOCK, MINB)
bound_rays_coop_kernel(GridView gv, const float4* __restrict__ indicator, float w1, const float* __restrict__ rays_o,
                       const float* __restrict__ dirs, const float* __restrict__ near, const float* __restrict__ far,
                       int64_t R, int n_grid, float thresh, int32_t* __restrict__ bnear, int32_t* __restrict__ bfar,
                       ShellGrid shell) {
  NMB_COOP_PROLOGUE()
  const int64_t r = chain % R;
  const int s_begin = (int)(chain / R) * BOUND_SEG_COOP;
  if (s_begin >= n_grid) return;
  const int s_end = min(s_begin + BOUND_SEG_COOP, n_grid);
  const float ox = rays_o[r * 3 + 0], oy = rays_o[r * 3 + 1], oz = rays_o[r * 3 + 2];
  const float dx = dirs[r * 3 + 0], dy = dirs[r * 3 + 1], dz = dirs[r * 3 + 2];
  const float nr = near[r], fr = far[r];
  float d;
  int32_t ix;
  bool have_prev = false;   // the lanes hold the neighbours of some earlier sample of this ray (valid warm start)
  // mesh distance at sample s, or +inf / -inf when the sample lies in a cell certified outside / inside the shell
  auto ds_at = [&](int s, float& depth) {
    const float tt = linspace01(s, n_grid);
    depth = __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, tt)), __fmul_rn(fr, tt));  // renderer.py:81
    const float qx = __fadd_rn(ox, __fmul_rn(depth, dx));
    const float qy = __fadd_rn(oy, __fmul_rn(depth, dy));
    const float qz = __fadd_rn(oz, __fmul_rn(depth, dz));
    if (shell.cells) {
      const float sc = 0.5f * (float)shell.G / shell.B;
      const float fx = (qx + shell.B) * sc, fy = (qy + shell.B) * sc, fz = (qz + shell.B) * sc;
      if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)shell.G && fy < (float)shell.G && fz < (float)shell.G) {
        const int64_t cell = ((int64_t)(int)fz * shell.G + (int)fy) * shell.G + (int)fx;
        const uint8_t code = __ldg(shell.cells + cell);
        if (code == 1) return CUDART_INF_F;    // proven outside the shell: mask false
        if (code == 2) return -CUDART_INF_F;   // proven inside the shell: mask true (only the depth matters)
      } else {
        const float ex = qx - shell.cx, ey = qy - shell.cy, ez = qz - shell.cz;
        if (ex * ex + ey * ey + ez * ez >= shell.far_r * shell.far_r) return CUDART_INF_F;
      }
    }
    const bool warm = have_prev;
    have_prev = true;
    return coop::query(gv, ln, stk, root_link, root_cnt, indicator, w1, qx, qy, qz, warm, d, ix, KnnOut{}, 0);
  };
  int first = -1;
  float depth = 0.f;
  for (int s = s_begin; s < s_end; ++s) {
    if (ds_at(s, depth) < thresh) {
      first = s;
      if (ln.gl == 0) atomicMin(&bnear[r], __float_as_int(depth));   // depths are >= 0: bit patterns order like values
      break;
    }
  }
  if (first < 0) return;  // no hit in this segment
  for (int s = s_end - 1; s >= first; --s) {
    if (s == first) {   // known to be a hit: the loop always terminates with a far candidate
      if (ln.gl == 0) atomicMax(&bfar[r], __float_as_int(depth));
      break;
    }
    float dd;
    if (ds_at(s, dd) < thresh) {
      if (ln.gl == 0) atomicMax(&bfar[r], __float_as_int(dd));
      break;
    }
  }
}

static int coop_minb() {
  static const int v = getenv("NMB_KNN_MINB") ? atoi(getenv("NMB_KNN_MINB")) : 8;
  return v;
}
// launch helper: picks the instantiation for the tuned number of resident blocks per SM
#define NMB_COOP_LAUNCH(kernel, nblocks, stream, ...)                                                      \
  do {                                                                                                     \
    const int mb_ = coop_minb();                                                                           \
    if (mb_ >= 12) kernel<12><<<(unsigned)(nblocks), coop::BLOCK, 0, stream>>>(__VA_ARGS__);              \
    else if (mb_ >= 10) kernel<10><<<(unsigned)(nblocks), coop::BLOCK, 0, stream>>>(__VA_ARGS__);         \
    else if (mb_ >= 8) kernel<8><<<(unsigned)(nblocks), coop::BLOCK, 0, stream>>>(__VA_ARGS__);           \
    else kernel<6><<<(unsigned)(nblocks), coop::BLOCK, 0, stream>>>(__VA_ARGS__);                         \
  } while (0)

// chains resident on the device at once (for sizing segments)
static int64_t coop_resident_chains() { return (int64_t)sm_count() * coop_minb() * coop::GROUPS_PER_BLOCK; }
constexpr int64_t COOP_RAY_KERNEL_MIN_RAYS = 2048;   // below this the per-point kernels expose more parallelism

int launch_knn_lists(const nmb_grid* g, const float4* indicator_sorted, float w1, const float* xyz, const int32_t* off,
                     const int32_t* cnt, int64_t R, int64_t M, int max_list, KnnOut out, cudaStream_t stream) {
  if (M <= 0 || R <= 0) return 0;
  ProfScope prof(PROF_KNN_LIST, M, stream);
  if (!knn_legacy()) {
    // segments: enough chains for ~4 waves, at least 8 entries each
    int64_t nseg = ceil_div(4 * coop_resident_chains(), R);
    nseg = std::max<int64_t>(1, std::min<int64_t>(nseg, ceil_div(max_list, 8)));
    const int seg = (int)ceil_div(max_list, nseg);
    const int max_seg = (int)ceil_div(max_list, seg);
    NMB_COOP_LAUNCH(knn_lists_coop_kernel, ceil_div(R * max_seg, coop::GROUPS_PER_BLOCK), stream, make_view(g),
                    indicator_sorted, w1, xyz, off, cnt, R, seg, max_seg, out);
    NMB_LAUNCH_OK();
    return 0;
  }
  // entries per thread: 16 when there is plenty of work (one cold walk per 16 queries); shorter segments when the lists
  // of a small shard (multi-GPU single-frame mode) would leave the GPU under-filled - only rays that hit the object have
  // entries at all, so the number of busy threads is ~M / seg, which should cover ~2 waves of the resident threads
  int seg = 16;
  while (seg > 4 && M / seg < (int64_t)sm_count() * 1280 * 2) seg >>= 1;
  const int max_seg = (int)ceil_div(max_list, seg);
  knn_lists_kernel<<<(unsigned)ceil_div(R * max_seg, 128), 128, 0, stream>>>(g->nodes.p, g->pts.p, indicator_sorted, w1,
                                                                            xyz, off, cnt, R, seg, max_seg, out,
                                                                            make_view(g));
  NMB_LAUNCH_OK();
  return 0;
}

constexpr int64_t RAY_KERNEL_MIN_RAYS = 32768;  // below this the per-point kernels expose more parallelism

int launch_knn_distance(const nmb_grid* g, const float4* indicator_sorted, float w1, PointSrc src, int64_t P,
                        KnnOut out, cudaStream_t stream) {
  if (P <= 0) return 0;
  ProfScope prof(PROF_KNN, P, stream);
  if (!knn_legacy()) {
    if (!src.xyz && src.R >= COOP_RAY_KERNEL_MIN_RAYS && P % src.R == 0) {
      const int S = (int)(P / src.R);
      int64_t nseg = ceil_div(4 * coop_resident_chains(), src.R);
      nseg = std::max<int64_t>(1, std::min<int64_t>(nseg, ceil_div(S, 8)));
      const int seg = (int)ceil_div(S, nseg);
      nseg = ceil_div(S, seg);
      NMB_COOP_LAUNCH(knn_rays_coop_kernel, ceil_div(src.R * nseg, coop::GROUPS_PER_BLOCK), stream, make_view(g),
                      indicator_sorted, w1, src, S, seg, out);
    } else {
      NMB_COOP_LAUNCH(knn_points_coop_kernel, ceil_div(P, coop::GROUPS_PER_BLOCK), stream, make_view(g),
                      indicator_sorted, w1, src, P, out);
    }
    NMB_LAUNCH_OK();
    return 0;
  }
  if (!src.xyz && src.R >= RAY_KERNEL_MIN_RAYS && P % src.R == 0) {
    const int S = (int)(P / src.R);
    // segments per ray: enough threads for ~2 waves of the 1280 resident threads per SM (10 blocks of 128) - every segment
    // starts with a cold walk, so no more segments than the occupancy needs
    int64_t nseg = ceil_div((int64_t)sm_count() * 1280 * 2, src.R);
    nseg = std::max<int64_t>(1, std::min<int64_t>(nseg, ceil_div(S, 4)));   // at least 4 samples per segment
    const int seg = (int)ceil_div(S, nseg);
    nseg = ceil_div(S, seg);
    static const int minb = getenv("NMB_KNN_MINB") ? atoi(getenv("NMB_KNN_MINB")) : 10;
    const GridView gv = make_view(g);
    static const int order = getenv("NMB_KNN_ORDER") ? atoi(getenv("NMB_KNN_ORDER")) : 1;   // 0 = full child sort (A/B)
    const unsigned gridn = (unsigned)ceil_div(src.R * nseg, 128);
    if (minb >= 12) { if (order) knn_rays_kernel<12, 1><<<gridn, 128, 0, stream>>>(g->nodes.p, g->pts.p, indicator_sorted, w1, src, S, seg, out, gv); else knn_rays_kernel<12, 0><<<gridn, 128, 0, stream>>>(g->nodes.p, g->pts.p, indicator_sorted, w1, src, S, seg, out, gv); }
    else if (minb >= 10) { if (order) knn_rays_kernel<10, 1><<<gridn, 128, 0, stream>>>(g->nodes.p, g->pts.p, indicator_sorted, w1, src, S, seg, out, gv); else knn_rays_kernel<10, 0><<<gridn, 128, 0, stream>>>(g->nodes.p, g->pts.p, indicator_sorted, w1, src, S, seg, out, gv); }
    else { if (order) knn_rays_kernel<8, 1><<<gridn, 128, 0, stream>>>(g->nodes.p, g->pts.p, indicator_sorted, w1, src, S, seg, out, gv); else knn_rays_kernel<8, 0><<<gridn, 128, 0, stream>>>(g->nodes.p, g->pts.p, indicator_sorted, w1, src, S, seg, out, gv); }
    NMB_LAUNCH_OK();
    return 0;
  }
  knn_distance_kernel<<<(unsigned)ceil_div(P, 128), 128, 0, stream>>>(g->nodes.p, g->pts.p, indicator_sorted, w1, src,
                                                                      P, out);
  NMB_LAUNCH_OK();
  return 0;
}

// renderer.py:79-90,94-95 (compute_bounded_near_far): `n_grid` samples along each ray's sphere chord; keep the min /
// max depth whose mesh distance is below `thresh`.  Nothing per-sample is stored: the two extrema are reduced with
// integer atomics on the (non-negative) depth bit patterns.
__global__ void __launch_bounds__(128)
bound_scan_kernel(const float4* __restrict__ nodes, const float4* __restrict__ pts,
                  const float4* __restrict__ indicator, float w1, const float* __restrict__ rays_o,
                  const float* __restrict__ dirs, const float* __restrict__ near, const float* __restrict__ far,
                  int64_t R, int n_grid, float thresh, int32_t* __restrict__ bnear, int32_t* __restrict__ bfar) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R * n_grid) return;
  const int64_t r = i % R;
  const int s = (int)(i / R);
  const float t = linspace01(s, n_grid);
  const float d = __fadd_rn(__fmul_rn(near[r], __fsub_rn(1.0f, t)), __fmul_rn(far[r], t));  // renderer.py:81
  const float qx = __fadd_rn(rays_o[r * 3 + 0], __fmul_rn(d, dirs[r * 3 + 0]));
  const float qy = __fadd_rn(rays_o[r * 3 + 1], __fmul_rn(d, dirs[r * 3 + 1]));
  const float qz = __fadd_rn(rays_o[r * 3 + 2], __fmul_rn(d, dirs[r * 3 + 2]));
  float d2[KNN_K];
  int32_t ix[KNN_K];
  knn_walk<KNN_K, false>(nodes, pts, qx, qy, qz, d2, ix);
  float w[KNN_K], ds, grad[3];
  mesh_distance_point(pts, indicator, w1, qx, qy, qz, d2, ix, w, ds, grad);
  if (ds < thresh) {
    atomicMin(&bnear[r], __float_as_int(d));
    atomicMax(&bfar[r], __float_as_int(d));
  }
}

// Ray-ordered bounded-near/far scan.  near = min, far = max over the samples with ds < thresh.  The n_grid samples of a
// ray are split into segments of BOUND_SEG consecutive samples, one thread each (t = g * R + r): a thread walks its
// segment from the front to its first hit (candidate for near, atomicMin on the depth bits) and from the back to its
// last hit (candidate for far, atomicMax); samples between the two cannot change either extremum and are not
// evaluated, and samples in cells of the shell certificate grid are decided without evaluation.  Output-identical to
// evaluating all samples; the serial chain per thread is at most BOUND_SEG walks (short tails even with few rays).
constexpr int BOUND_SEG = 32;

__global__ void __launch_bounds__(128)
bound_rays_kernel(const float4* __restrict__ nodes, const float4* __restrict__ pts,
                  const float4* __restrict__ indicator, float w1, const float* __restrict__ rays_o,
                  const float* __restrict__ dirs, const float* __restrict__ near, const float* __restrict__ far,
                  int64_t R, int n_grid, float thresh, int32_t* __restrict__ bnear, int32_t* __restrict__ bfar,
                  ShellGrid shell, const GridView gv) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t r = t % R;
  const int s_begin = (int)(t / R) * BOUND_SEG;
  if (s_begin >= n_grid) return;
  const int s_end = min(s_begin + BOUND_SEG, n_grid);
  const float ox = rays_o[r * 3 + 0], oy = rays_o[r * 3 + 1], oz = rays_o[r * 3 + 2];
  const float dx = dirs[r * 3 + 0], dy = dirs[r * 3 + 1], dz = dirs[r * 3 + 2];
  const float nr = near[r], fr = far[r];
  float d2[KNN_K];
  int32_t ix[KNN_K];
  bool have_prev = false;   // d2 / ix hold the neighbours of some earlier sample of this ray (valid warm start)
  // returns the mesh distance at sample s, or +inf / -inf when the sample lies in a cell certified to be outside /
  // inside the shell
  auto ds_at = [&](int s, float& depth) {
    const float tt = linspace01(s, n_grid);
    depth = __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, tt)), __fmul_rn(fr, tt));  // renderer.py:81
    const float qx = __fadd_rn(ox, __fmul_rn(depth, dx));
    const float qy = __fadd_rn(oy, __fmul_rn(depth, dy));
    const float qz = __fadd_rn(oz, __fmul_rn(depth, dz));
    if (shell.cells) {
      const float sc = 0.5f * (float)shell.G / shell.B;
      const float fx = (qx + shell.B) * sc, fy = (qy + shell.B) * sc, fz = (qz + shell.B) * sc;
      if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)shell.G && fy < (float)shell.G && fz < (float)shell.G) {
        const int64_t cell = ((int64_t)(int)fz * shell.G + (int)fy) * shell.G + (int)fx;
        const uint8_t code = __ldg(shell.cells + cell);
        if (code == 1) return CUDART_INF_F;    // proven outside the shell: mask false
        if (code == 2) return -CUDART_INF_F;   // proven inside the shell: mask true (only the depth matters)
      } else {
        const float ex = qx - shell.cx, ey = qy - shell.cy, ez = qz - shell.cz;
        if (ex * ex + ey * ey + ez * ez >= shell.far_r * shell.far_r) return CUDART_INF_F;
      }
    }
    const bool warm = have_prev;
    have_prev = true;
    if (warm) {
      warm_rerank<KNN_K>(pts, qx, qy, qz, d2, ix);
      knn_walk<KNN_K, true>(nodes, pts, qx, qy, qz, d2, ix, NMB_GV_ARG(gv));
    } else {
      knn_walk<KNN_K, false>(nodes, pts, qx, qy, qz, d2, ix);
    }
    float w[KNN_K], ds, grad[3];
    mesh_distance_point(pts, indicator, w1, qx, qy, qz, d2, ix, w, ds, grad);
    return ds;
  };
  int first = -1;
  float depth = 0.f;
  for (int s = s_begin; s < s_end; ++s) {
    if (ds_at(s, depth) < thresh) {
      first = s;
      atomicMin(&bnear[r], __float_as_int(depth));   // depths are >= 0: their bit patterns order like the values
      break;
    }
  }
  if (first < 0) return;  // no hit in this segment
  for (int s = s_end - 1; s >= first; --s) {
    // s == first is known to be a hit: the loop always terminates with a far candidate
    if (s == first) {
      atomicMax(&bfar[r], __float_as_int(depth));   // `depth` still holds sample `first` unless overwritten below
      break;
    }
    float dd;
    if (ds_at(s, dd) < thresh) {
      atomicMax(&bfar[r], __float_as_int(dd));
      break;
    }
  }
}

// Two-ended variant of the ray-ordered scan (the default for frame-sized batches): only the FIRST and the LAST hit of
// a ray matter, so the front-to-back search and the back-to-front search are separate launches that talk to each other
// through bnear / bfar:
//   launch 1 (BACKWARD = false): thread (segment g ascending, ray r) looks for the first hit of its segment, but gives
//     up as soon as bnear[r] shows a hit in front of the sample it is about to evaluate;
//   launch 2 (BACKWARD = true):  thread (segment g DESCENDING, ray r) looks for the last hit of its segment between the
//     segment's end and the ray's first hit (final after launch 1), and gives up when bfar[r] shows a hit behind it.
// Blocks are scheduled in grid order, so by the time a segment starts the segments that make it redundant have usually
// finished; a stale read only costs work, never correctness (bnear / bfar only move towards their final values, and a
// thread only skips samples that provably cannot change them).  Rays that cross the object evaluate the two OUTER
// crossings of the 0.1 shell only - not the inner boundary of the shell (ds rises above 0.1 again deep inside the
// object) - and rays without any hit are scanned once, not twice.  Depths are taken to be non-decreasing along a ray,
// as in bound_rays_kernel.
template <bool BACKWARD>
__global__ void __launch_bounds__(128)
bound_dir_kernel(const float4* __restrict__ nodes, const float4* __restrict__ pts, const float4* __restrict__ indicator,
                 float w1, const float* __restrict__ rays_o, const float* __restrict__ dirs,
                 const float* __restrict__ near, const float* __restrict__ far, int64_t R, int n_grid, float thresh,
                 int32_t* __restrict__ bnear, int32_t* __restrict__ bfar, ShellGrid shell, const GridView gv) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t r = t % R;
  const int nseg = (n_grid + BOUND_SEG - 1) / BOUND_SEG;
  const int gi = (int)(t / R);
  if (gi >= nseg) return;
  const int s_begin = (BACKWARD ? nseg - 1 - gi : gi) * BOUND_SEG;
  const int s_end = min(s_begin + BOUND_SEG, n_grid);
  const float nr = near[r], fr = far[r];
  auto depth_at = [&](int s) {
    const float tt = linspace01(s, n_grid);
    return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, tt)), __fmul_rn(fr, tt));  // renderer.py:81
  };
  // depths are >= 0, so their bit patterns order like the values (bnear starts at +inf, bfar at -1)
  int32_t first_bits = 0;
  if (BACKWARD) {
    first_bits = bnear[r];
    if (first_bits == 0x7f800000) return;                               // the ray has no hit at all
    if (__float_as_int(depth_at(s_end - 1)) < first_bits) return;      // the whole segment lies in front of the first hit
    if (__ldcg(bfar + r) >= __float_as_int(depth_at(s_end - 1))) return;   // a hit behind this segment is known
  } else {
    if (__ldcg(bnear + r) < __float_as_int(depth_at(s_begin))) return;   // a hit in front of this segment is known
  }
  const float ox = rays_o[r * 3 + 0], oy = rays_o[r * 3 + 1], oz = rays_o[r * 3 + 2];
  const float dx = dirs[r * 3 + 0], dy = dirs[r * 3 + 1], dz = dirs[r * 3 + 2];
  float d2[KNN_K];
  int32_t ix[KNN_K];
  bool have_prev = false;   // d2 / ix hold the neighbours of some earlier sample of this ray (valid warm start)
  for (int i = 0; i < s_end - s_begin; ++i) {
    const int s = BACKWARD ? s_end - 1 - i : s_begin + i;
    const float depth = depth_at(s);
    const int32_t dbits = __float_as_int(depth);
    if (BACKWARD && dbits <= first_bits) {   // reached the first hit: it is the last one as well, as far as this thread knows
      atomicMax(&bfar[r], first_bits);
      return;
    }
    const float qx = __fadd_rn(ox, __fmul_rn(depth, dx));
    const float qy = __fadd_rn(oy, __fmul_rn(depth, dy));
    const float qz = __fadd_rn(oz, __fmul_rn(depth, dz));
    int code = 0;   // 1: proven outside the shell (mask false), 2: proven inside (mask true), 0: evaluate
    if (shell.cells) {
      const float sc = 0.5f * (float)shell.G / shell.B;
      const float fx = (qx + shell.B) * sc, fy = (qy + shell.B) * sc, fz = (qz + shell.B) * sc;
      if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)shell.G && fy < (float)shell.G && fz < (float)shell.G) {
        code = __ldg(shell.cells + ((int64_t)(int)fz * shell.G + (int)fy) * shell.G + (int)fx);
      } else {
        const float ex = qx - shell.cx, ey = qy - shell.cy, ez = qz - shell.cz;
        if (ex * ex + ey * ey + ez * ez >= shell.far_r * shell.far_r) code = 1;
      }
    }
    if (code == 1) continue;
    bool hit = (code == 2);
    if (!hit) {
      // about to walk the octree: worth a look at what the other segments of this ray have found meanwhile
      if (BACKWARD ? (__ldcg(bfar + r) >= dbits) : (__ldcg(bnear + r) < dbits)) return;
      if (have_prev) {
        warm_rerank<KNN_K>(pts, qx, qy, qz, d2, ix);
        knn_walk<KNN_K, true>(nodes, pts, qx, qy, qz, d2, ix, NMB_GV_ARG(gv));
      } else {
        knn_walk<KNN_K, false>(nodes, pts, qx, qy, qz, d2, ix);
        have_prev = true;
      }
      float w[KNN_K], ds, grad[3];
      mesh_distance_point(pts, indicator, w1, qx, qy, qz, d2, ix, w, ds, grad);
      hit = ds < thresh;
    }
    if (hit) {
      if (BACKWARD) atomicMax(&bfar[r], dbits);
      else atomicMin(&bnear[r], dbits);
      return;
    }
  }
}

int launch_bound_scan(const nmb_grid* g, const float4* indicator, float w1, const float* rays_o, const float* dirs,
                      const float* near, const float* far, int64_t R, int n_grid, float thresh, int32_t* bnear,
                      int32_t* bfar, ShellGrid shell, cudaStream_t stream) {
  const int64_t n = R * n_grid;
  if (n <= 0) return 0;
  ProfScope prof(PROF_BOUND, n, stream);
  if (!knn_legacy()) {
    if (R >= COOP_RAY_KERNEL_MIN_RAYS) {
      const int64_t nseg = ceil_div(n_grid, BOUND_SEG_COOP);
      NMB_COOP_LAUNCH(bound_rays_coop_kernel, ceil_div(R * nseg, coop::GROUPS_PER_BLOCK), stream, make_view(g), indicator,
                      w1, rays_o, dirs, near, far, R, n_grid, thresh, bnear, bfar, (thresh == 0.1f) ? shell : ShellGrid{});
    } else {
      NMB_COOP_LAUNCH(bound_scan_coop_kernel, ceil_div(n, coop::GROUPS_PER_BLOCK), stream, make_view(g), indicator, w1,
                      rays_o, dirs, near, far, R, n_grid, thresh, bnear, bfar);
    }
    NMB_LAUNCH_OK();
    return 0;
  }
  static const bool two_ended = getenv("NMB_BOUND_ONE_PASS") == nullptr;
  if (R >= RAY_KERNEL_MIN_RAYS && two_ended) {
    const int64_t nseg = ceil_div(n_grid, BOUND_SEG);
    const ShellGrid sg = (thresh == 0.1f) ? shell : ShellGrid{};
    bound_dir_kernel<false><<<(unsigned)ceil_div(R * nseg, 128), 128, 0, stream>>>(
        g->nodes.p, g->pts.p, indicator, w1, rays_o, dirs, near, far, R, n_grid, thresh, bnear, bfar, sg, make_view(g));
    NMB_LAUNCH_OK();
    bound_dir_kernel<true><<<(unsigned)ceil_div(R * nseg, 128), 128, 0, stream>>>(
        g->nodes.p, g->pts.p, indicator, w1, rays_o, dirs, near, far, R, n_grid, thresh, bnear, bfar, sg, make_view(g));
    NMB_LAUNCH_OK();
    return 0;
  }
  if (R >= RAY_KERNEL_MIN_RAYS) {
    const int64_t nseg = ceil_div(n_grid, BOUND_SEG);
    bound_rays_kernel<<<(unsigned)ceil_div(R * nseg, 128), 128, 0, stream>>>(
        g->nodes.p, g->pts.p, indicator, w1, rays_o, dirs, near, far, R, n_grid, thresh, bnear, bfar,
        (thresh == 0.1f) ? shell : ShellGrid{}, make_view(g));
    NMB_LAUNCH_OK();
    return 0;
  }
  bound_scan_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(g->nodes.p, g->pts.p, indicator, w1, rays_o, dirs,
                                                                    near, far, R, n_grid, thresh, bnear, bfar);
  NMB_LAUNCH_OK();
  return 0;
}

// Generic K (<= 32) for the frnn shim: candidates kept in a local-memory list.
__global__ void __launch_bounds__(128)
knn_generic_kernel(const float4* __restrict__ nodes, const float4* __restrict__ pts, const float* __restrict__ xyz,
                   int64_t M, int K, float r2, float* __restrict__ d2_out, int64_t* __restrict__ idx_out) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float qx = xyz[m * 3], qy = xyz[m * 3 + 1], qz = xyz[m * 3 + 2];
  float d[32];
  int32_t ix[32];
  for (int k = 0; k < K; ++k) {
    d[k] = CUDART_INF_F;
    ix[k] = -1;
  }
  int32_t sn[STACK_MAX];
  float sd[STACK_MAX];
  int sp = 1;
  sn[0] = 0;
  sd[0] = 0.f;
  while (sp > 0) {
    --sp;
    const int32_t n = sn[sp];
    if (sd[sp] >= d[K - 1]) continue;
    const float4 a = __ldg(&nodes[NODE_F4 * n]);
    const float4 b = __ldg(&nodes[NODE_F4 * n + 1]);
    const int32_t link = __float_as_int(a.w);
    const int32_t cnt = __float_as_int(b.w);
    if (cnt < 0) {
      for (int32_t i = link; i < link - cnt; ++i) {
        const float4 p = __ldg(&pts[i]);
        const float dd = sq_dist_rn(qx, qy, qz, p.x, p.y, p.z);
        if (dd < d[K - 1]) {
          int k = K - 1;
          while (k > 0 && dd < d[k - 1]) {
            d[k] = d[k - 1];
            ix[k] = ix[k - 1];
            --k;
          }
          d[k] = dd;
          ix[k] = __float_as_int(p.w);  // original index
        }
      }
    } else {
      // push children farthest-first (selection by repeated max over <= 8 entries)
      float cd[8];
      for (int c = 0; c < 8; ++c) {
        cd[c] = -1.f;
        if (c < cnt) {
          const float4* nc = nodes + NODE_F4 * (link + c);
          const float bd = node_bound(qx, qy, qz, __ldg(nc), __ldg(nc + 1), __ldg(nc + 2), __ldg(nc + 3));
          if (bd < d[K - 1]) cd[c] = bd;
        }
      }
      for (int it = 0; it < cnt; ++it) {
        int best = -1;
        float bv = -1.f;
        for (int c = 0; c < cnt; ++c)
          if (cd[c] > bv) {
            bv = cd[c];
            best = c;
          }
        if (best < 0) break;
        if (sp >= STACK_MAX) __trap();   // cannot happen for depth <= 10; never drop a subtree silently
        sn[sp] = link + best;
        sd[sp] = bv;
        ++sp;
        cd[best] = -1.f;
      }
    }
  }
  for (int k = 0; k < K; ++k) {
    const bool ok = (ix[k] >= 0) && (d[k] <= r2);
    d2_out[m * K + k] = ok ? d[k] : -1.f;
    idx_out[m * K + k] = ok ? (int64_t)ix[k] : (int64_t)-1;
  }
}

// SoA (sorted slots) -> row-major API outputs (original vertex order)
__global__ void export_knn_kernel(const int32_t* __restrict__ order, KnnOut in, int64_t M, float* __restrict__ ds,
                                  int64_t* __restrict__ idx, float* __restrict__ w, float* __restrict__ grad) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= M * KNN_K) return;
  const int64_t m = t / KNN_K;
  const int k = (int)(t % KNN_K);
  if (idx) idx[t] = (int64_t)order[in.slot[k * in.stride + m]];
  if (w) w[t] = in.w[k * in.stride + m];
  if (k == 0 && ds) ds[m] = in.ds[m];
  if (k < 3 && grad) grad[m * 3 + k] = in.grad[k * in.stride + m];
}

int launch_export_knn(const nmb_grid* g, KnnOut in, int64_t M, float* ds, int64_t* idx, float* w, float* grad,
                      cudaStream_t stream) {
  if (M <= 0) return 0;
  export_knn_kernel<<<(unsigned)ceil_div(M * KNN_K, 256), 256, 0, stream>>>(g->order.p, in, M, ds, idx, w, grad);
  NMB_LAUNCH_OK();
  return 0;
}

__global__ void permute_rows4_kernel(const float* __restrict__ src /*[V,3]*/, const int32_t* __restrict__ order,
                                     int64_t V, float4* __restrict__ dst) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= V) return;
  const int64_t o = order[i];
  dst[i] = make_float4(src[o * 3], src[o * 3 + 1], src[o * 3 + 2], 0.f);
}

int permute_indicator(const nmb_grid* g, const float* indicator, float4* dst, cudaStream_t stream) {
  permute_rows4_kernel<<<(unsigned)ceil_div(g->V, 256), 256, 0, stream>>>(indicator, g->order.p, g->V, dst);
  NMB_LAUNCH_OK();
  return 0;
}

}  // namespace nmb

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int nmb_grid_create(const float* vertices, int64_t V, void* stream, nmb_grid** out) {
  if (!out) return 2;
  *out = nullptr;
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
    nmb::set_error("no CUDA device: neumesh_b200 has no CPU path");
    return 3;
  }
  nmb_grid* g = new nmb_grid();
  int rc = nmb::build_grid(vertices, V, static_cast<cudaStream_t>(stream), g);
  if (rc != 0) {
    delete g;
    return rc;
  }
  *out = g;
  return 0;
}

void nmb_grid_destroy(nmb_grid* g) { delete g; }

int64_t nmb_grid_num_vertices(const nmb_grid* g) { return g ? g->V : 0; }

const int32_t* nmb_grid_order(const nmb_grid* g) { return g ? g->order.p : nullptr; }

int nmb_knn(const nmb_grid* g, const float* xyz, int64_t M, int K, float r, float* d2, int64_t* idx, void* stream) {
  NMB_CHECK(g != nullptr, "null grid");
  NMB_CHECK(K >= 1 && K <= 32, "K must be in [1,32]");
  NMB_CHECK(K <= g->V, "K exceeds the number of vertices");
  if (M <= 0) return 0;
  nmb::knn_generic_kernel<<<(unsigned)nmb::ceil_div(M, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      g->nodes.p, g->pts.p, xyz, M, K, r * r, d2, idx);
  NMB_LAUNCH_OK();
  return 0;
}

int nmb_mesh_distance(const nmb_grid* g, const float* indicator, float indicator_weight, const float* xyz, int64_t M,
                      float* ds, int64_t* idx, float* w, float* grad_ds, void* stream_) {
  NMB_CHECK(g != nullptr, "null grid");
  if (M <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // scratch: permuted indicator + SoA outputs (API convenience path; the renderer uses packed fields instead)
  nmb::StreamBuf ind_buf, soa_buf;
  NMB_CUDA_OK(ind_buf.alloc(sizeof(float4) * g->V, stream));
  NMB_CUDA_OK(soa_buf.alloc(sizeof(float) * M * 20, stream));
  float4* ind = ind_buf.as<float4>();
  float* soa = soa_buf.as<float>();
  int rc = nmb::permute_indicator(g, indicator, ind, stream);
  if (rc) return rc;
  nmb::KnnOut out;
  out.ds = soa;
  out.slot = reinterpret_cast<int32_t*>(soa + M);
  out.w = soa + 9 * M;
  out.grad = soa + 17 * M;
  out.stride = M;
  nmb::PointSrc src{xyz, nullptr, nullptr, nullptr, 0};
  rc = nmb::launch_knn_distance(g, ind, indicator_weight, src, M, out, stream);
  if (rc) return rc;
  return nmb::launch_export_knn(g, out, M, ds, idx, w, grad_ds, stream);
}

}  // extern "C"
