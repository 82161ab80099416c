// Group-cooperative exact 8-NN walk: EIGHT LANES PER QUERY (four queries per warp).
//
// The thread-per-query walk of knn_walk.cuh spends its time in the descent: per expanded node up to 8 children x
// (2-4 float4 loads + box/disc bound) and a 19-comparator sorting network, executed by ONE lane while on average 22 of
// the 32 lanes of the warp idle (ncu r1c: 9-10 active lanes per issued instruction).  Here the 8 lanes of a group share
// one query:
//   * expansion : lane c loads child c's 64-byte record (the children of a node are contiguous: one coalesced 512-byte
//                 read per group) and evaluates its box + disc bound; survivors are ranked with 8 shuffles and pushed
//                 farthest-first onto a per-group stack in SHARED memory (one parallel store each);
//   * leaf scan : 8 points per step, one per lane; candidates that beat the current 8th-best are inserted one by one;
//   * result    : the ascending top-8 list is DISTRIBUTED - lane k holds rank k - so an insertion is one shuffle-up and
//                 a select per lane, the warm re-rank is one distance per lane + a 6-stage bitonic network, and the
//                 mesh distance (mesh_grid.py:121-144) evaluates one neighbour per lane followed by shuffle sums.
// Candidates are ranked by the same total order (squared distance, slot) as knn_walk.cuh, so the selected set is unique:
// this walk, the thread-per-query walk and an fp32 brute force agree bit for bit.
//
// Directory start.  A warm-started query already holds 8 real points, i.e. a ball (q, r) that contains every possible
// better neighbour.  The top levels of the octree are replaced by a look-up: per directory level l (dense 8^l table
// indexed by the Morton prefix, built with the octree) the entry of a cell is the id of the octree node with that
// prefix, of the LEAF ancestor if the tree stopped above level l, or -1 if the cell is empty.  The walk starts from the
// finest level whose cells are at least as wide as the ball's extent in x, y and z - the ball then meets at most 2x2x2
// cells, one per lane - instead of from the root.  Exactness: the quantiser cell(x) = clamp(floor((x - bmin) *
// inv_cell)) >> shift is monotone in x (every fp32 operation in it is), every point within r of q has coordinates in
// [q - r', q + r'] (r' = r inflated beyond any rounding), so its cell index lies between those of the two corners.
#pragma once
#include <math_constants.h>

#include "knn_walk.cuh"

namespace nmb {
namespace coop {

constexpr int G = 8;                          // lanes per query
constexpr int STACK = 80;                     // entries per group: <= 7 pushes net per level, depth <= 10, + root
constexpr int STACK_WORDS = STACK * 3 + 1;    // 12-byte entries {bound, link, count}; odd word stride between groups
constexpr int GROUPS_PER_BLOCK = 16;          // 128 threads
constexpr int BLOCK = G * GROUPS_PER_BLOCK;

struct Lane {
  unsigned gmask;   // the 8 lanes of this group inside the warp
  int gl;           // lane within the group: rank held in the distributed list, child / point / cell handled
  int gbase;        // first lane of the group inside the warp
};

__device__ __forceinline__ Lane make_lane() {
  Lane ln;
  const int lane = threadIdx.x & 31;
  ln.gl = lane & (G - 1);
  ln.gbase = lane & ~(G - 1);
  ln.gmask = 0xFFu << ln.gbase;
  return ln;
}

template <typename T>
__device__ __forceinline__ T bcast(const Lane& ln, T v, int src) {
  return __shfl_sync(ln.gmask, v, src, G);
}

__device__ __forceinline__ unsigned group_ballot(const Lane& ln, bool pred) {
  return (__ballot_sync(ln.gmask, pred) >> ln.gbase) & 0xFFu;
}

// Insert the group-uniform candidate (nd, ni) into the distributed ascending list (precondition: it ranks before the
// entry of lane 7).  Lane k keeps its entry if the candidate ranks after it, takes the candidate if it ranks between the
// entries of lanes k-1 and k, and takes lane k-1's entry otherwise.
__device__ __forceinline__ void list_insert(const Lane& ln, float& d, int32_t& ix, float nd, int32_t ni) {
  const float pd = __shfl_up_sync(ln.gmask, d, 1, G);
  const int32_t pi = __shfl_up_sync(ln.gmask, ix, 1, G);
  if (cand_less(nd, ni, d, ix)) {
    const bool from_below = (ln.gl > 0) && cand_less(nd, ni, pd, pi);
    d = from_below ? pd : nd;
    ix = from_below ? pi : ni;
  }
}

// Warm start: recompute the distance of this lane's point to the new query and sort the 8 entries ascending under the
// total order (bitonic network over the group; entries are distinct points, so the order is strict).
__device__ __forceinline__ void list_rerank(const Lane& ln, const float4* __restrict__ pts, float qx, float qy, float qz,
                                            float& d, int32_t& ix) {
  const float4 p = __ldg(&pts[ix]);
  d = sq_dist_rn(qx, qy, qz, p.x, p.y, p.z);
#pragma unroll
  for (int k = 2; k <= G; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float od = __shfl_xor_sync(ln.gmask, d, j, G);
      const int32_t oi = __shfl_xor_sync(ln.gmask, ix, j, G);
      const bool up = (ln.gl & k) == 0;        // ascending block
      const bool lower = (ln.gl & j) == 0;     // lower lane of the pair
      const bool other_less = cand_less(od, oi, d, ix);
      if ((lower == up) ? other_less : !other_less) {
        d = od;
        ix = oi;
      }
    }
  }
}

__device__ __forceinline__ float child_bound(const float4* __restrict__ nc, float qx, float qy, float qz, float worst,
                                             int32_t& link, int32_t& cnt) {
  const float4 a = __ldg(nc);
  const float4 b = __ldg(nc + 1);
  link = __float_as_int(a.w);
  cnt = __float_as_int(b.w);
  float bd = box_dist_rn(qx, qy, qz, a, b);
  if (bd <= worst) {   // the disc bound is only loaded for children the box cannot reject
    bd = fmaxf(bd, disc_bound(qx, qy, qz, __ldg(nc + 2), __ldg(nc + 3)));
  }
  return (bd <= worst) ? bd : CUDART_INF_F;
}

// Rank the surviving candidates of the group (bd < inf) by (bound, lane) and push them farthest-first, so that the
// nearest one is popped first.  Returns the new stack pointer.
__device__ __forceinline__ int push_sorted(const Lane& ln, uint32_t* stk, int sp, float bd, int32_t link, int32_t cnt) {
  const bool alive = bd < CUDART_INF_F;
  const int m = __popc(group_ballot(ln, alive));
  if (m == 0) return sp;
  int rank = 0;
#pragma unroll
  for (int c = 0; c < G; ++c) {
    const float od = bcast(ln, bd, c);
    rank += (od < bd || (od == bd && c < ln.gl)) ? 1 : 0;
  }
  if (sp + m > STACK) {
    // cannot happen for depth <= 10 (7 net pushes per level); never drop a subtree silently
    if (ln.gl == 0) printf("neumesh_b200: KNN traversal stack overflow (sp %d + %d)\n", sp, m);
    __trap();
  }
  if (alive) {
    uint32_t* e = stk + 3 * (sp + m - 1 - rank);
    e[0] = __float_as_uint(bd);
    e[1] = (uint32_t)link;
    e[2] = (uint32_t)cnt;
  }
  __syncwarp(ln.gmask);
  return sp + m;
}

// The walk.  On entry (WARM) lane k holds the k-th of 8 DISTINCT real points and its distance to q, ascending; on exit
// lane k holds the k-th nearest point of the whole set.  WARM = false initialises an empty list.
template <bool WARM>
__device__ __forceinline__ void walk(const GridView& gv, const Lane& ln, uint32_t* stk, int32_t root_link,
                                     int32_t root_cnt, float qx, float qy, float qz, float& d, int32_t& ix) {
  if (!WARM) {
    d = CUDART_INF_F;
    ix = 0x7fffffff;
  }
  float wd = bcast(ln, d, G - 1);
  int32_t wi = bcast(ln, ix, G - 1);
  int sp = 0;
  bool started = false;
  if (WARM && gv.dir != nullptr && gv.dir_lmax >= gv.dir_lmin) {
    // ---- directory start ----
    const float r = sqrtf(wd) * 1.00001f + 1e-6f;
    const int maxc = (1 << gv.levels) - 1;
    const float q[3] = {qx, qy, qz};
    int lo[3], hi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float tl = (__fsub_rn(q[c], r) - gv.bmin[c]) * gv.inv_cell;
      const float th = (__fadd_rn(q[c], r) - gv.bmin[c]) * gv.inv_cell;
      lo[c] = min(max((int)floorf(fmaxf(tl, -1.f)), 0), maxc);
      hi[c] = min(max((int)floorf(fminf(th, 1.0e9f)), 0), maxc);
    }
    int lev = -1;
    for (int l = gv.dir_lmax; l >= gv.dir_lmin; --l) {
      const int sh = gv.levels - l;
      if ((hi[0] >> sh) - (lo[0] >> sh) <= 1 && (hi[1] >> sh) - (lo[1] >> sh) <= 1 && (hi[2] >> sh) - (lo[2] >> sh) <= 1) {
        lev = l;
        break;
      }
    }
    if (lev >= 0) {
      const int sh = gv.levels - lev;
      const int cx = (lo[0] >> sh) + (ln.gl & 1), cy = (lo[1] >> sh) + ((ln.gl >> 1) & 1),
                cz = (lo[2] >> sh) + ((ln.gl >> 2) & 1);
      const bool in = cx <= (hi[0] >> sh) && cy <= (hi[1] >> sh) && cz <= (hi[2] >> sh);
      int32_t nid = -1;
      if (in) {
        const uint32_t m = (spread_bits10((uint32_t)cx) << 2) | (spread_bits10((uint32_t)cy) << 1) | spread_bits10((uint32_t)cz);
        // level l starts at sum_{i = lmin}^{l - 1} 8^i = (8^l - 8^lmin) / 7
        const int32_t off = (int32_t)(((1u << (3 * lev)) - (1u << (3 * gv.dir_lmin))) / 7u);
        nid = __ldg(gv.dir + off + m);
      }
      // several cells may map to one leaf ancestor: keep its first occurrence only
      bool dup = false;
#pragma unroll
      for (int c = 0; c < G; ++c) {
        const int32_t o = bcast(ln, nid, c);
        dup |= (c < ln.gl) && (o == nid);
      }
      float bd = CUDART_INF_F;
      int32_t link = 0, cnt = 0;
      if (nid >= 0 && !dup) bd = child_bound(gv.nodes + NODE_F4 * (int64_t)nid, qx, qy, qz, wd, link, cnt);
      sp = push_sorted(ln, stk, 0, bd, link, cnt);
      started = true;
    }
  }
  if (!started) {
    if (ln.gl == 0) {
      stk[0] = __float_as_uint(0.f);
      stk[1] = (uint32_t)root_link;
      stk[2] = (uint32_t)root_cnt;
    }
    __syncwarp(ln.gmask);
    sp = 1;
  }
  // "while-while": descend through internal nodes until the group holds a leaf, then scan it
  while (true) {
    int32_t leaf_b = 0, leaf_e = 0;
    while (sp > 0) {
      --sp;
      const float sd = __uint_as_float(stk[3 * sp]);
      const int32_t link = (int32_t)stk[3 * sp + 1];
      const int32_t cnt = (int32_t)stk[3 * sp + 2];
      __syncwarp(ln.gmask);   // every lane has read the entry before any lane overwrites the slot
      if (sd > wd) continue;  // '>' (not '>='): an equidistant point with a smaller index may still enter
      if (cnt < 0) {
        leaf_b = link;
        leaf_e = link - cnt;
        break;
      }
      float bd = CUDART_INF_F;
      int32_t clink = 0, ccnt = 0;
      if (ln.gl < cnt) bd = child_bound(gv.nodes + NODE_F4 * (int64_t)(link + ln.gl), qx, qy, qz, wd, clink, ccnt);
      sp = push_sorted(ln, stk, sp, bd, clink, ccnt);
    }
    if (leaf_e == leaf_b) break;   // stack exhausted without another leaf
    for (int32_t i0 = leaf_b; i0 < leaf_e; i0 += G) {
      const int32_t i = i0 + ln.gl;
      float dd = CUDART_INF_F;
      if (i < leaf_e) {
        const float4 p = __ldg(&gv.pts[i]);
        dd = sq_dist_rn(qx, qy, qz, p.x, p.y, p.z);
      }
      unsigned m = group_ballot(ln, (i < leaf_e) && cand_less(dd, i, wd, wi));
      while (m) {
        const int c = __ffs(m) - 1;
        m &= m - 1;
        const float nd = bcast(ln, dd, c);
        const int32_t ni = i0 + c;
        if (cand_less(nd, ni, wd, wi)) {   // the bound may have tightened since the ballot
          const bool present = group_ballot(ln, ix == ni) != 0;   // warm lists / shared leaves: never insert twice
          if (!present) {
            list_insert(ln, d, ix, nd, ni);
            wd = bcast(ln, d, G - 1);
            wi = bcast(ln, ix, G - 1);
          }
        }
      }
    }
  }
}

// mesh_grid.py:121-144 with one neighbour per lane.  w: this lane's normalised weight; ds: the mesh distance (all lanes);
// grad: d ds / d xyz (all lanes).  wsum and ds are accumulated in neighbour order (k = 0..7) exactly as the
// thread-per-query kernels do, so ds / w are bit-identical to them.
__device__ __forceinline__ void mesh_distance(const Lane& ln, const float4* __restrict__ pts,
                                              const float4* __restrict__ indicator, float w1, float qx, float qy,
                                              float qz, float d2, int32_t ix, float& w, float& ds, float (&grad)[3]) {
  float wk = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(d2), 1e-7f));   // :123-124
  float wsum = 0.f;
#pragma unroll
  for (int k = 0; k < G; ++k) wsum = __fadd_rn(wsum, bcast(ln, wk, k));
  wk = __fdiv_rn(wk, wsum);   // :125
  const float4 p = __ldg(&pts[ix]);
  const float4 nv = __ldg(&indicator[ix]);
  const float vx = __fsub_rn(qx, p.x), vy = __fsub_rn(qy, p.y), vz = __fsub_rn(qz, p.z);   // :134
  const float rho = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
  const float D = __fadd_rn(w1, rho);
  const float mx = __fdiv_rn(__fadd_rn(__fmul_rn(nv.x, w1), __fmul_rn(vx, rho)), D);   // :136
  const float my = __fdiv_rn(__fadd_rn(__fmul_rn(nv.y, w1), __fmul_rn(vy, rho)), D);
  const float mz = __fdiv_rn(__fadd_rn(__fmul_rn(nv.z, w1), __fmul_rn(vz, rho)), D);
  const float dot = __fadd_rn(__fadd_rn(__fmul_rn(vx, mx), __fmul_rn(vy, my)), __fmul_rn(vz, mz));
  const float term = __fmul_rn(wk, dot);
  ds = 0.f;
#pragma unroll
  for (int k = 0; k < G; ++k) ds = __fadd_rn(ds, bcast(ln, term, k));   // :137-142
  // d(dot)/dx = (w1 n + 3 rho v) / D - dot * v / (rho D)   (norm's sub-gradient at rho = 0 is 0)
  const float invD = 1.0f / D;
  const float c2 = rho > 0.f ? dot / (rho * D) : 0.f;
  float g0 = wk * ((w1 * nv.x + 3.f * rho * vx) * invD - c2 * vx);
  float g1 = wk * ((w1 * nv.y + 3.f * rho * vy) * invD - c2 * vy);
  float g2 = wk * ((w1 * nv.z + 3.f * rho * vz) * invD - c2 * vz);
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    g0 += __shfl_xor_sync(ln.gmask, g0, o, G);
    g1 += __shfl_xor_sync(ln.gmask, g1, o, G);
    g2 += __shfl_xor_sync(ln.gmask, g2, o, G);
  }
  w = wk;
  grad[0] = g0;
  grad[1] = g1;
  grad[2] = g2;
}

// One query of a chain: cold when `warm` is false, else warm-started from the list the lanes still hold.
// Returns the mesh distance; writes this lane's slice of the outputs when out.ds != nullptr.
__device__ __forceinline__ float query(const GridView& gv, const Lane& ln, uint32_t* stk, int32_t root_link,
                                       int32_t root_cnt, const float4* __restrict__ indicator, float w1, float qx,
                                       float qy, float qz, bool warm, float& d, int32_t& ix, const KnnOut& out,
                                       int64_t p) {
  if (warm) {
    list_rerank(ln, gv.pts, qx, qy, qz, d, ix);
    walk<true>(gv, ln, stk, root_link, root_cnt, qx, qy, qz, d, ix);
  } else {
    walk<false>(gv, ln, stk, root_link, root_cnt, qx, qy, qz, d, ix);
  }
  float w, ds, grad[3];
  mesh_distance(ln, gv.pts, indicator, w1, qx, qy, qz, d, ix, w, ds, grad);
  if (out.ds) {
    out.slot[ln.gl * out.stride + p] = ix;
    out.w[ln.gl * out.stride + p] = w;
    if (ln.gl == 0) out.ds[p] = ds;
    if (out.grad && ln.gl < 3) out.grad[ln.gl * out.stride + p] = ln.gl == 0 ? grad[0] : (ln.gl == 1 ? grad[1] : grad[2]);
  }
  return ds;
}

}  // namespace coop
}  // namespace nmb
