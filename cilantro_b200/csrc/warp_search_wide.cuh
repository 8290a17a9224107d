// Warp-cooperative exact nearest neighbour WITH an exclusion bound (product code, sm_100a).
//
// Same search as warp_grid_nearest() (warp_search.cuh) — same candidates' arithmetic, same conservative
// region bounds, same tie rule, hence the same (index, d2) bit for bit — plus one more output:
//
//   D2 = a lower bound of the squared distance from the query to EVERY reference point other than the
//        returned best ("the second-nearest point is at least sqrt(D2) away").
//
// The device-resident ICP loop (icp_loop.cu) caches (best, sqrt(D2)) per query: in the next iteration the
// query has moved by delta = |T' s - T s|, so every other point is still at least sqrt(D2) - delta away
// (triangle inequality), and if the cached match's new distance is below that, it is provably still the
// exact nearest neighbour — no grid access at all. To make D2 useful the search is WIDENED: after the own
// cell, a region is scanned iff its lower bound is below (sqrt(best) + slack)^2 instead of best, the two
// smallest distances are tracked instead of one, and D2 = min(second smallest scanned, the widened bound,
// the squared distance to the boundary of the scanned 3x3x3 block).
#pragma once
#include "warp_search.cuh"

namespace cb {

struct WideSearchSmem {
  float4 q[32];                // query position
  unsigned long long key[32];  // merged best: d2 bits << 32 | sorted position (0xffffffff = none)
  unsigned int sec[32];        // merged second-smallest d2 (float bits; non-negative floats order like uints)
  uint2 item[kWarpItemsMax];   // .x = first cell index, .y = (#cells << 8) | lane
};

struct WideBest {
  float d2;   // squared distance of the nearest point with d2 < max_d2 (else max_d2)
  int idx;    // its original index, -1 = none
  int pos;    // its position in the cell-sorted array, -1 = none
  float D2;   // every OTHER reference point has true squared distance >= D2 (0 = unknown)
};

// two smallest of the candidates seen so far: (b1, p1) and b2; a candidate bit-equal to b1 lands in b2
__device__ __forceinline__ void two_smallest(float r, int j, float& b1, int& p1, float& b2) {
  if (r < b1) {
    b2 = b1;
    b1 = r;
    p1 = j;
  } else {
    b2 = fminf(b2, r);
  }
}

// scans [b, e) tracking the two smallest distances; position `skip` (already accounted for) is ignored
__device__ __forceinline__ void scan_range_two(const float4* __restrict__ pts, uint32_t b, uint32_t e, float qx, float qy,
                                               float qz, float& b1, int& p1, float& b2, int skip) {
  constexpr int kW = 4;
  auto eval = [&](const float4& p, uint32_t j) {
    const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
    float r = __fmul_rn(dx, dx);
    r = __fadd_rn(r, __fmul_rn(dy, dy));
    r = __fadd_rn(r, __fmul_rn(dz, dz));
    if ((int)j != skip) two_smallest(r, (int)j, b1, p1, b2);
  };
  for (uint32_t j = b; j < e; j += kW) {
    float4 p[kW];
    p[0] = __ldg(pts + j);
#pragma unroll
    for (int u = 1; u < kW; u++)
      if (j + u < e) p[u] = __ldg(pts + j + u);
    eval(p[0], j);
#pragma unroll
    for (int u = 1; u < kW; u++)
      if (j + u < e) eval(p[u], j + u);
  }
}

// All 32 lanes of the warp must call this (inactive lanes pass active = false).
// warm_pos >= 0: sorted position of a point known to be close (the cached match); slack >= 0: widening of the
// search radius beyond the nearest distance, in the units of the coordinates.
__device__ __forceinline__ WideBest warp_grid_nearest_wide(const GridView& g, WideSearchSmem& sm, bool active, float qx,
                                                           float qy, float qz, float max_d2, int warm_pos, float slack) {
  const unsigned int lane = threadIdx.x & 31;
  const unsigned int lt_mask = (1u << lane) - 1u;
  constexpr float kInf = 3.402823466e+38f;
  WideBest out;
  out.d2 = max_d2;
  out.idx = -1;
  out.pos = -1;
  out.D2 = 0.f;

  const float fx = cell_coord(qx, g.ox, g.inv_h), fy = cell_coord(qy, g.oy, g.inv_h), fz = cell_coord(qz, g.oz, g.inv_h);
  const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
  const float hs2 = g.h_safe * g.h_safe;
  const bool inside = active && g.n > 0 && cx >= 0 && cx < g.nx && cy >= 0 && cy < g.ny && cz >= 0 && cz < g.nz;
  bool slow = active && g.n > 0 && !inside;  // outside the grid: per-lane exact search at the end

  float b1 = kInf, b2 = kInf;  // two smallest distances over ALL scanned candidates (regardless of max_d2)
  int p1 = -1;
  if (inside && warm_pos >= 0) {
    const float4 p = __ldg(g.pts + warm_pos);
    const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
    float r = __fmul_rn(dx, dx);
    r = __fadd_rn(r, __fmul_rn(dy, dy));
    r = __fadd_rn(r, __fmul_rn(dz, dz));
    b1 = r;
    p1 = warm_pos;
  }
  sm.q[lane] = make_float4(qx, qy, qz, 0.f);
  __syncwarp();

  // ---- phase A: own cell + work items --------------------------------------------------------------
  unsigned int count = 0;  // warp-uniform number of queued items
  float wide2 = 0.f;       // regions with a lower bound >= wide2 are not scanned
  {
    const uint32_t cbase = inside ? ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx : 0u;
    uint32_t s1 = 0, s2 = 0;
    if (inside) {
      s1 = __ldg(g.cell_start + cbase + cx);
      s2 = __ldg(g.cell_start + cbase + cx + 1);
      scan_range_two(g.pts, s1, s2, qx, qy, qz, b1, p1, b2, warm_pos);
    }
    {
      // widened bound: (sqrt(min(b1, max_d2)) + slack)^2, rounded up
      const float w = __fadd_ru(__fsqrt_ru(fminf(b1, max_d2)), slack);
      wide2 = __fmul_ru(w, w);
    }
    const float gxl = slab_gap(fx, cx, cx - 1), gxr = slab_gap(fx, cx, cx + 1);
    const float gym = slab_gap(fy, cy, cy - 1), gyp = slab_gap(fy, cy, cy + 1);
    const float gzm = slab_gap(fz, cz, cz - 1), gzp = slab_gap(fz, cz, cz + 1);
    const float gy2[3] = {gym * gym, 0.f, gyp * gyp};
    const float gz2[3] = {gzm * gzm, 0.f, gzp * gzp};
    const int xm = max(cx - 1, 0), xp = min(cx + 1, g.nx - 1);
    constexpr int kDy[8] = {-1, 1, 0, 0, -1, 1, -1, 1};
    constexpr int kDz[8] = {0, 0, -1, 1, -1, -1, 1, 1};
    // (Tried: a 10-bit need mask per lane, ONE warp scan and lane-major item order instead of a ballot per region -
    // fewer instructions (this loop is 21 % of the cold kernel's, ncu source page) but 8 % slower: region-major order
    // makes neighbouring lanes of the pooled scan read neighbouring rows of the sorted array.)
#pragma unroll
    for (int t = 0; t < 10; ++t) {
      bool need;
      uint32_t first;
      uint32_t ncells;
      if (t == 0) {  // left x-neighbour
        need = inside && cx > 0 && (gxl * gxl * hs2 < wide2);
        first = cbase + (uint32_t)(cx - 1);
        ncells = 1;
      } else if (t == 1) {  // right x-neighbour
        need = inside && cx < g.nx - 1 && (gxr * gxr * hs2 < wide2);
        first = cbase + (uint32_t)(cx + 1);
        ncells = 1;
      } else {
        const int ry = cy + kDy[t - 2], rz = cz + kDz[t - 2];
        const bool valid = inside && ry >= 0 && ry < g.ny && rz >= 0 && rz < g.nz;
        need = valid && ((gy2[kDy[t - 2] + 1] + gz2[kDz[t - 2] + 1]) * hs2 < wide2);
        first = ((uint32_t)rz * (uint32_t)g.ny + (uint32_t)ry) * (uint32_t)g.nx + (uint32_t)xm;
        ncells = (uint32_t)(xp - xm + 1);
      }
      const unsigned int m = __ballot_sync(0xffffffffu, need);
      if (need) sm.item[count + __popc(m & lt_mask)] = make_uint2(first, (ncells << 8) | lane);
      count += __popc(m);
    }
    // a candidate at or beyond the radius can never be the match: it only bounds the others
    if (b1 < max_d2) {
      sm.key[lane] = pack_key(b1, (unsigned int)p1);
      sm.sec[lane] = __float_as_uint(b2);
    } else {
      sm.key[lane] = pack_key(max_d2, 0xffffffffu);
      sm.sec[lane] = __float_as_uint(b1);  // b1 <= b2
    }
  }
  __syncwarp();

  // ---- phase B: pooled scan of the queued regions ----------------------------------------------------
  for (unsigned int k = lane; k < count; k += 32) {
    const uint2 it = sm.item[k];
    const unsigned int ql = it.y & 31u, nc = it.y >> 8;
    const float4 q = sm.q[ql];
    const uint32_t b = __ldg(g.cell_start + it.x), e = __ldg(g.cell_start + it.x + nc);
    if (b >= e) continue;
    const unsigned long long cur = sm.key[ql];
    float l1 = kInf, l2 = kInf;
    int lp = -1;
    // the only point that can be met twice is the warm seed, and only while it is the running best
    scan_range_two(g.pts, b, e, q.x, q.y, q.z, l1, lp, l2, (int)(unsigned int)(cur & 0xffffffffull));
    if (lp < 0) continue;
    float loser = l1;  // what this region contributes to "second smallest" besides l2
    if (l1 < max_d2) {
      const unsigned long long key = pack_key(l1, (unsigned int)lp);
      const unsigned long long old = atomicMin(&sm.key[ql], key);
      const float od2 = __uint_as_float((unsigned int)(old >> 32));
      // the loser of (previous best, this region's best) is a second-best candidate; the initial "none"
      // sentinel is not a point
      loser = ((unsigned int)(old & 0xffffffffull) == 0xffffffffu) ? kInf : fmaxf(od2, l1);
    }
    atomicMin(&sm.sec[ql], __float_as_uint(fminf(l2, loser)));
  }
  __syncwarp();

  // ---- phase C: merged result, termination, exclusion bound, rare exact fallbacks -----------------------
  if (inside) {
    const unsigned long long key = sm.key[lane];
    const unsigned int pos = (unsigned int)(key & 0xffffffffull);
    out.d2 = __uint_as_float((unsigned int)(key >> 32));
    out.pos = (pos == 0xffffffffu) ? -1 : (int)pos;
    const float sec = __uint_as_float(sm.sec[lane]);
    float cover = kInf;
    bool any = false;
    if (cx - 1 > 0) { cover = fminf(cover, fx - (float)(cx - 1)); any = true; }
    if (cx + 1 < g.nx - 1) { cover = fminf(cover, (float)(cx + 2) - fx); any = true; }
    if (cy - 1 > 0) { cover = fminf(cover, fy - (float)(cy - 1)); any = true; }
    if (cy + 1 < g.ny - 1) { cover = fminf(cover, (float)(cy + 2) - fy); any = true; }
    if (cz - 1 > 0) { cover = fminf(cover, fz - (float)(cz - 1)); any = true; }
    if (cz + 1 < g.nz - 1) { cover = fminf(cover, (float)(cz + 2) - fz); any = true; }
    cover -= kCellMargin;
    const float cover2 = (!any) ? kInf : (cover > 0.f ? cover * cover * hs2 : 0.f);
    const bool done = cover2 >= out.d2;
    const bool tie = out.pos >= 0 && sec == out.d2;  // a second point at a bit-equal distance: index rule
    if (!done || tie) slow = true;
    out.D2 = fminf(fminf(sec, wide2), cover2);
  }
  if (slow) {
    // rare: outside the grid, shell >= 2 needed, or an exact tie to resolve on the original index.
    // No exclusion bound: the next iteration searches this query again.
    const Best bst = grid_nearest(g, qx, qy, qz, max_d2);
    out.d2 = bst.d2;
    out.idx = bst.idx;
    out.pos = bst.pos;
    out.D2 = 0.f;
  } else if (out.pos >= 0) {
    out.idx = __float_as_int(__ldg(&g.pts[out.pos].w));
  }
  __syncwarp();  // the shared arrays are reused by the caller's next query batch
  return out;
}

}  // namespace cb
