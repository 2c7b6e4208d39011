// pt_bvh.cuh -- closest hit for scenes with many geoms: the conservative filter of pt_filter.cuh driven by a
// bounding-volume hierarchy instead of a linear scan (sm_100a).
//
// SURVEY.md 8f rank 1 / BASELINE config "10k spheres/cubes": a linear scan is O(n) per segment.  The hierarchy does
// not change any result: it only skips geoms the filter would have proven to miss.
//   * Every leaf is one geom with the same four-class filter test as the pair scan (scalar here: a ray meets few
//     leaves, and they differ per lane).
//   * Every node stores its two children's world AABBs.  A child's box bounds the INFLATED shapes of all geoms below
//     it at w = 0; the w- and distance-dependent part of the inflation is added per ray as a pad
//     P1*w + P2*D^2 (P1, P2 = maxima over the subtree, D = largest distance from the origin to the box), so "the ray
//     misses the padded box" implies "the filter proves a miss for every geom below".
//   * Pass 1 (bvh_scan) computes the same (lo1, k1, lo2) as the linear scan would over the geoms it visits; a child is
//     skipped if its box's entry distance minus the largest world slack is >= lo2 -- nothing below it can change k1
//     or lower lo2.  The exact test of k1 then decides as usual (resolve in pt_kernels.cuh).
//   * Pass 2 (bvh_exact, the fallback; rare) is the exact scan restricted to the filter's candidates: every candidate
//     leaf whose bound does not exceed the best exact distance so far goes through exact_hit; smaller distance wins,
//     ties go to the lower geom index (the index-order rule of the specification, applied explicitly because the
//     traversal order is not the index order).
// Built on the host (pt_api.cu: build_bvh): median split of the centroids along the widest axis, one geom per leaf.
#pragma once
#include "pt_filter.cuh"

namespace ptd {

// node = 5 float4:  n0 = (min0.xyz, max0.x)  n1 = (max0.yz, min1.xy)  n2 = (min1.z, max1.xyz)
//                   n3 = (P1_0, P2_0, P1_1, P2_1)  n4 = int bits (child0, child1, -, -); child >= 0: node, < 0: leaf ~child
// leaf = 5 float4 (scalar filter record) + int2 (class, geom index):
//   class 0: l0 = (c.xyz, Wc)  l1 = (Ww, Wr, Ew_c, Ew_w)
//   class 2: l0 = (c.xyz, Ew_c)  l1 = (Hc.xyz, Ew_w)  l2 = (Hw.xyz, -)
//   class 1: l0..l2 = rows x,y,z of inverseTransform  l3 = (R2c, R2w, R2r, Ew_c)  l4 = (Ew_w, -, -, -)
//   class 3: l0..l2 = rows x,y,z of inverseTransform  l3 = (hc.xyz, Ew_c)  l4 = (hw.xyz, Ew_w)
constexpr int kBvhNodeRows = 5, kBvhLeafRows = 5, kBvhStack = 48;
struct BvhSoA {
  const float4* nodes;
  const float4* leaves;
  const int2* leaf_meta;
  int n_leaves;              // 0 = no hierarchy (the pair scan is used)
  float ew_c_max, ew_w_max;  // largest world slack E_w = Ew_c + Ew_w * w over all geoms
};

// the filter test of one leaf: false = proven miss, else `lo` = lower bound on the exact world distance
__device__ __forceinline__ bool leaf_filter(int cls, const float4* __restrict__ L, const ScanRay& r, float& lo) {
  if (cls == 0) {
    const float4 C0 = __ldg(L), C1 = __ldg(L + 1);
    const float ocx = r.o.x - C0.x, ocy = r.o.y - C0.y, ocz = r.o.z - C0.z;
    const float b = __fmaf_rn(ocx, r.d.x, __fmaf_rn(ocy, r.d.y, ocz * r.d.z));
    const float oc2 = __fmaf_rn(ocx, ocx, __fmaf_rn(ocy, ocy, ocz * ocz));
    const float R2 = __fmaf_rn(C1.y, oc2, __fmaf_rn(C1.x, r.w, C0.w));
    const float disc = __fmaf_rn(r.a, R2 - oc2, b * b);
    if (disc < 0.0f) return false;
    const float sd = mufu_sqrt(disc), ia = mufu_rcp(r.a);
    if ((sd - b) * ia < 0.0f) return false;
    lo = __fmaf_rn((-b - sd) * ia, r.dl, -__fmaf_rn(C1.w, r.w, C1.z));
    return true;
  }
  if (cls == 2) {
    const float4 C0 = __ldg(L), C1 = __ldg(L + 1), C2 = __ldg(L + 2);
    const float cx = __fmaf_rn(C0.x, r.id.x, -r.od.x), cy = __fmaf_rn(C0.y, r.id.y, -r.od.y), cz = __fmaf_rn(C0.z, r.id.z, -r.od.z);
    const float hx = __fmaf_rn(C2.x, r.w, C1.x), hy = __fmaf_rn(C2.y, r.w, C1.y), hz = __fmaf_rn(C2.z, r.w, C1.z);
    const float ax = fabsf(r.id.x), ay = fabsf(r.id.y), az = fabsf(r.id.z);
    const float tnear = fmaxf(fmaxf(__fmaf_rn(hx, -ax, cx), __fmaf_rn(hy, -ay, cy)), __fmaf_rn(hz, -az, cz));
    const float tfar = fminf(fminf(__fmaf_rn(hx, ax, cx), __fmaf_rn(hy, ay, cy)), __fmaf_rn(hz, az, cz));
    if (tnear > tfar || tfar < 0.0f) return false;
    lo = __fmaf_rn(tnear, r.dl, -__fmaf_rn(C1.w, r.w, C0.w));
    return true;
  }
  // object-space classes: (ro, rw) = inverseTransform * (o, d), fused
  const float4 A0 = __ldg(L), A1 = __ldg(L + 1), A2 = __ldg(L + 2), K0 = __ldg(L + 3), K1 = __ldg(L + 4);
  const float rox = __fmaf_rn(A0.x, r.o.x, __fmaf_rn(A0.y, r.o.y, __fmaf_rn(A0.z, r.o.z, A0.w)));
  const float roy = __fmaf_rn(A1.x, r.o.x, __fmaf_rn(A1.y, r.o.y, __fmaf_rn(A1.z, r.o.z, A1.w)));
  const float roz = __fmaf_rn(A2.x, r.o.x, __fmaf_rn(A2.y, r.o.y, __fmaf_rn(A2.z, r.o.z, A2.w)));
  const float rwx = __fmaf_rn(A0.x, r.d.x, __fmaf_rn(A0.y, r.d.y, A0.z * r.d.z));
  const float rwy = __fmaf_rn(A1.x, r.d.x, __fmaf_rn(A1.y, r.d.y, A1.z * r.d.z));
  const float rwz = __fmaf_rn(A2.x, r.d.x, __fmaf_rn(A2.y, r.d.y, A2.z * r.d.z));
  if (cls == 1) {
    const float a = __fmaf_rn(rwx, rwx, __fmaf_rn(rwy, rwy, rwz * rwz));
    const float b = __fmaf_rn(rox, rwx, __fmaf_rn(roy, rwy, roz * rwz));
    const float ro2 = __fmaf_rn(rox, rox, __fmaf_rn(roy, roy, roz * roz));
    const float R2 = __fmaf_rn(K0.z, ro2, __fmaf_rn(K0.y, r.w, K0.x));
    const float disc = __fmaf_rn(a, R2 - ro2, b * b);
    if (disc < 0.0f) return false;
    const float sd = mufu_sqrt(disc), ia = mufu_rcp(a);
    if ((sd - b) * ia < 0.0f) return false;
    lo = __fmaf_rn((-b - sd) * ia, r.dl, -__fmaf_rn(K1.x, r.w, K0.w));
    return true;
  }
  const float hx = __fmaf_rn(K1.x, r.w, K0.x), hy = __fmaf_rn(K1.y, r.w, K0.y), hz = __fmaf_rn(K1.z, r.w, K0.z);
  const float ix = mufu_rcp(rwx), iy = mufu_rcp(rwy), iz = mufu_rcp(rwz);
  const float cx = -rox * ix, cy = -roy * iy, cz = -roz * iz;
  const float tnear = fmaxf(fmaxf(__fmaf_rn(hx, -fabsf(ix), cx), __fmaf_rn(hy, -fabsf(iy), cy)), __fmaf_rn(hz, -fabsf(iz), cz));
  const float tfar = fminf(fminf(__fmaf_rn(hx, fabsf(ix), cx), __fmaf_rn(hy, fabsf(iy), cy)), __fmaf_rn(hz, fabsf(iz), cz));
  if (tnear > tfar || tfar < 0.0f) return false;
  lo = __fmaf_rn(tnear, r.dl, -__fmaf_rn(K1.w, r.w, K0.w));
  return true;
}

// Entry parameter of the ray into a child's padded box, or +inf if the ray provably misses it.
// A NaN (0 * inf) is ignored by fminf / fmaxf, so the affected slab does not constrain: conservative.
__device__ __forceinline__ float child_entry(float3 bmin, float3 bmax, float p1, float p2, const ScanRay& r) {
  // D = largest distance from the origin to a point of the box (bounds |o - c| of every geom inside)
  const float mx = fmaxf(fabsf(bmin.x - r.o.x), fabsf(bmax.x - r.o.x));
  const float my = fmaxf(fabsf(bmin.y - r.o.y), fabsf(bmax.y - r.o.y));
  const float mz = fmaxf(fabsf(bmin.z - r.o.z), fabsf(bmax.z - r.o.z));
  const float D2 = __fmaf_rn(mx, mx, __fmaf_rn(my, my, mz * mz)) * 1.000001f;
  const float pad = __fmaf_rn(p2, D2, p1 * r.w);
  // (b - o) * (1/d), subtraction first: with d_i = 0 the two planes give -inf / +inf (origin inside the slab: no
  // constraint) or the same infinity twice (outside: miss); b/d - o/d would turn the first case into inf - inf
  const float ax = ((bmin.x - pad) - r.o.x) * r.id.x, bx = ((bmax.x + pad) - r.o.x) * r.id.x;
  const float ay = ((bmin.y - pad) - r.o.y) * r.id.y, by = ((bmax.y + pad) - r.o.y) * r.id.y;
  const float az = ((bmin.z - pad) - r.o.z) * r.id.z, bz = ((bmax.z + pad) - r.o.z) * r.id.z;
  // rounding of the six parameters (a few ulp of |box - o| / |d|) is covered by the 16u*w in P1
  const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
  const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
  if (tn > tf || tf < 0.0f) return INFINITY;
  return fmaxf(tn, 0.0f);  // NaN -> 0: "may be entered at once"
}

// One traversal serves both passes.  EXACT = false: filter scan, result in `best` (k1 = leaf index).
// EXACT = true: exact test of every candidate leaf that can still matter, result in `h`.
template <bool EXACT>
__device__ __forceinline__ void bvh_traverse(const BvhSoA& B, const GeomSoA& g, const ScanRay& r, f3 o, f3 d, ScanBest& best, Hit& h) {
  if (B.n_leaves <= 0) return;
  // largest world slack of any geom, plus room for the rounding differences between a node's slab parameters and a
  // leaf's own entry parameter (a few ulp of w)
  const float ewmax = __fmaf_rn(B.ew_w_max, r.w, B.ew_c_max) + 1e-6f * r.w;
  const float dls = r.dl * 0.999996f;
  int stack[kBvhStack];
  int sp = 0;
  int cur = B.n_leaves == 1 ? ~0 : 0;  // a single geom: the root is leaf 0
  for (;;) {
    if (cur < 0) {
      const int leaf = ~cur;
      const int2 meta = __ldg(B.leaf_meta + leaf);
      float lo;
      if (leaf_filter(meta.x, B.leaves + (size_t)leaf * kBvhLeafRows, r, lo)) {
        lo = fmaxf(lo, 0.0f);
        if (!EXACT) {
          scan_take(best, lo, leaf);
        } else if (!(lo > h.t)) {
          const int gi = meta.y;
          float dist;
          f3 P;
          int ncode;
          if (exact_hit(meta.x < 2 ? 0 : 1, __ldg(g.inv0 + gi), __ldg(g.inv1 + gi), __ldg(g.inv2 + gi), __ldg(g.fwd0 + gi),
                        __ldg(g.fwd1 + gi), __ldg(g.fwd2 + gi), o, d, dist, P, ncode)) {
            // specification: scan in index order, keep the strictly smaller positive distance
            if (dist > 0 && (dist < h.t || (dist == h.t && gi < h.id))) { h.t = dist; h.id = gi; h.p = P; h.ncode = ncode; }
          }
        }
      }
    } else {
      const float4* N = B.nodes + (size_t)cur * kBvhNodeRows;
      const float4 n0 = __ldg(N), n1 = __ldg(N + 1), n2 = __ldg(N + 2), n3 = __ldg(N + 3), n4 = __ldg(N + 4);
      float e0 = child_entry(make_float3(n0.x, n0.y, n0.z), make_float3(n0.w, n1.x, n1.y), n3.x, n3.y, r);
      float e1 = child_entry(make_float3(n1.z, n1.w, n2.x), make_float3(n2.y, n2.z, n2.w), n3.z, n3.w, r);
      // a child whose best possible bound cannot beat the current limit is skipped:
      //   filter pass: bound >= lo2 changes neither k1 nor lo2;  exact pass: bound > best exact distance (ties may still win)
      const float b0 = __fmaf_rn(e0, dls, -ewmax), b1 = __fmaf_rn(e1, dls, -ewmax);
      const bool v0 = e0 < INFINITY && (EXACT ? !(b0 > h.t) : (b0 < best.lo2));
      const bool v1 = e1 < INFINITY && (EXACT ? !(b1 > h.t) : (b1 < best.lo2));
      int c0 = __float_as_int(n4.x), c1 = __float_as_int(n4.y);
      if (v0 && v1) {
        if (e1 < e0) { const int t = c0; c0 = c1; c1 = t; }  // nearer child first
        stack[sp++] = c1;  // depth <= ceil(log2 n) <= 32 < kBvhStack by construction (median split)
        cur = c0;
        continue;
      }
      if (v0 || v1) { cur = v0 ? c0 : c1; continue; }
    }
    if (sp == 0) break;
    cur = stack[--sp];
  }
}

// the exact pass on its own (fallback of resolve_bvh; rare, so not inlined)
__device__ __noinline__ void bvh_exact(const BvhSoA B, const GeomSoA g, f3 o, f3 d, float r_scene, Hit& h) {
  const ScanRay r = make_scan_ray(o, d, r_scene, true);
  ScanBest unused;
  scan_init(unused);
  bvh_traverse<true>(B, g, r, o, d, unused, h);
}

// Resolve a finished BVH filter pass: exact test of the best candidate leaf, accepted if it is a hit closer than
// every other geom's lower bound; otherwise the exact pass.  Returns true if the fallback ran (statistics only).
__device__ __forceinline__ bool resolve_bvh(const ScanBest& best, const BvhSoA& B, const GeomSoA& g, float r_scene, f3 o, f3 d, Hit& h) {
  if (best.k1 < 0) return false;  // every geom is a proven miss
  const int2 meta = __ldg(B.leaf_meta + best.k1);
  const int gi = meta.y;
  float dist;
  f3 P;
  int ncode;
  const bool hit = exact_hit(meta.x < 2 ? 0 : 1, __ldg(g.inv0 + gi), __ldg(g.inv1 + gi), __ldg(g.inv2 + gi), __ldg(g.fwd0 + gi),
                             __ldg(g.fwd1 + gi), __ldg(g.fwd2 + gi), o, d, dist, P, ncode);
  if (hit && dist > 0 && dist < best.lo2) {
    h.t = dist; h.id = gi; h.p = P; h.ncode = ncode;
    return false;
  }
  bvh_exact(B, g, o, d, r_scene, h);
  return true;
}

}  // namespace ptd
