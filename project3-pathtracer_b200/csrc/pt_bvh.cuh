// pt_bvh.cuh -- closest hit for scenes with many geoms: the conservative filter of pt_filter.cuh driven by a
// bounding-volume hierarchy instead of a linear scan (sm_100a).
//
// SURVEY.md 8f rank 1 / BASELINE config "10k spheres/cubes": a linear scan is O(n) per segment.  The hierarchy does
// not change any result: it only skips geoms the filter would have proven to miss.
//   * Every leaf is one geom with the same four-class filter test as the pair scan (scalar here: a ray meets few
//     leaves, and they differ per lane).
//   * Every node stores the world AABBs of up to FOUR children (the binary surface-area tree, collapsed on the host: a
//     traversal is a chain of dependent fetches, and a 4-wide node halves its length).  A child's box bounds the INFLATED shapes of all geoms below
//     it at w = 0; the w- and distance-dependent part of the inflation is added per ray as a pad
//     P1*w + P2*D^2 (P1, P2 = maxima over the subtree, D = largest distance from the origin to the box), so "the ray
//     misses the padded box" implies "the filter proves a miss for every geom below".
//   * Pass 1 (the filter traversal) keeps the THREE smallest lower bounds and the two geoms k1, k2 they belong to; a
//     child is skipped if its box's entry distance minus the largest world slack is >= lo3 (nothing below it can change
//     any of them) or > hi, the smallest upper bound of a geom that is SURELY hit (leaf_filter_rows).  The exact test of
//     k1 decides as usual if it reports a hit closer than lo2; otherwise the exact test of k2 joins in, and the closer
//     of the two exact results is the closest hit if it is closer than lo3 -- every other geom is a proven miss or no
//     closer than its bound >= lo3 (bvh_phase2 in pt_kernels.cuh; the parity entry points below use k1 and lo2 only).
//     Why two candidates: the conservative bound of a small distant sphere is several times its radius (the
//     reference's own cancellation error), so the nearest candidate is often a near miss; settling those with a second
//     traversal that leaves k1 out cost 18 % of the 10 000-geom config's time for 2.5 % of its segments.
//   * Pass 2 (the fallback, ~0.1 % of the rays) is the exact scan restricted to the filter's candidates: every
//     candidate leaf whose bound does not exceed the best exact distance so far goes through exact_hit; smaller
//     distance wins, ties go to the lower geom index (the index-order rule of the specification, applied explicitly
//     because the traversal order is not the index order).  k_bounce_bvh collects the rays that need it and runs them
//     32 at a time (pt_kernels.cuh); the list kernel runs it on the spot (bvh_exact).
// Built on the host (pt_api.cu: build_filter): surface-area heuristic, one geom per leaf.
#pragma once
#include "pt_filter.cuh"

namespace ptd {

#ifdef PT_BVH_STACK_HIST  // measurement build: how deep the stack is when an entry is pushed (pt_debug_sp_hist)
__device__ unsigned long long g_sp_hist[64];  // [0..39] depth at push; 40 node visits, 41 leaf visits, 42 sure hits, 43 entries dropped at pop, 44 pops
#define PT_HIST(i) atomicAdd(&g_sp_hist[i], 1ull)
#else
#define PT_HIST(i)
#endif
#ifndef PT_BVH_SURE
#define PT_BVH_SURE 1
#endif


// node = 9 float4: two blocks of four rows, each holding two children side by side (child a in the low half of each
// float2, child b in the high half), children (0, 1) in rows 0-3 and (2, 3) in rows 4-7:
//   r0 = (min_a.x, min_b.x | min_a.y, min_b.y)  r1 = (min_a.z, min_b.z | max_a.x, max_b.x)  r2 = (max_a.y, max_b.y | max_a.z, max_b.z)
//   r3 = (P1_a, P1_b | P2_a, P2_b)
// row 8 = int bits (child0, child1, child2, child3); child >= 0: node, < 0: leaf ~child, kBvhNoChild: empty slot
// leaf = 5 float4 (scalar filter record) + int2 (class, geom index):
//   class 0: l0 = (c.xyz, Wc)  l1 = (Ww, Wr, Ew_c, Ew_w)
//   class 2: l0 = (c.xyz, Ew_c)  l1 = (Hc.xyz, Ew_w)  l2 = (Hw.xyz, -)
//   class 1: l0..l2 = rows x,y,z of inverseTransform  l3 = (R2c, R2w, R2r, Ew_c)  l4 = (Ew_w, -, -, -)
//   class 3: l0..l2 = rows x,y,z of inverseTransform  l3 = (hc.xyz, Ew_c)  l4 = (hw.xyz, Ew_w)
#ifndef PT_BVH_NODE128
#define PT_BVH_NODE128 0  // 1: 8 rows = one 128-byte line per node, the pads P1, P2 kept per NODE.  Measured: 2.60 instead of 2.98 Gseg/s --
                          // with a 128-byte stride row k of every lane's node falls into the same L1 banks (data pipe 84 % busy instead
                          // of 70 %); the 144-byte stride staggers them
#endif
constexpr int kBvhNodeRows = PT_BVH_NODE128 ? 8 : 9, kBvhLeafRows = 5;
constexpr int kBvhStack = 128;       // entries of a traversal's stack: the builder checks the collapsed tree's worst case
constexpr int kBvhBinaryDepth = 40;  // depth limit of the binary tree before the collapse
// reference of a leaf: ~(leaf index | filter class << kBvhLeafBits)  (< 0; distinct from the two markers below)
constexpr int kBvhLeafBits = 28;
__host__ __device__ inline int bvh_leaf_ref(int leaf, int cls) { return ~(leaf | (cls << kBvhLeafBits)); }
constexpr int kBvhNoChild = (int)0x80000000;
constexpr int kBvhDone = (int)0x80000001;  // `cur` of a finished traversal (k_bounce_bvh)
struct BvhSoA {
  const float4* nodes;
  const float4* leaves;
  const int2* leaf_meta;
  int n_leaves;              // 0 = no hierarchy (the pair scan is used)
  int root;                  // reference of the root: node 0, or the only leaf
  float ew_c_max, ew_w_max;  // largest world slack E_w = Ew_c + Ew_w * w over all geoms
};

// the filter test of one leaf: false = proven miss, else `lo` = lower bound on the exact world distance and `hi` = an
// UPPER bound on it if the exact test is SURE to report a hit, +inf otherwise.
// Sure hit: the error model of pt_filter.cuh is symmetric -- the point the exact path computes for a ray parameter lies
// within the deltas of the point the filter computes -- so a ray that passes through the shape DEFLATED by the same
// deltas (radius^2 1/4 - delta, half extents 1/2 - delta_i) is a ray the exact test reports as a hit, entered no later
// than the deflated shape is; with the origin outside the INFLATED shape (entry parameter > 0) the exact test takes its
// "outside" branch and reports the entry, not the exit.  hi = entry parameter into the deflated shape x |d| (rounded up)
// + the world slack.  Every "sure" needs comparisons to come out TRUE; a direction with a zero component (infinite
// reciprocals, NaN slabs that minimum / maximum would ignore) never gives one (`dlu` = +inf, class 3's own check).
// (the rows L0..L4 of the leaf's record are passed by value: k_bounce_bvh fetches them together with the node rows of the
// lanes that stand at a node, so that a step of a warp waits for memory once -- pt_kernels.cuh)
__device__ __forceinline__ bool leaf_filter_rows(int cls, const float4 L0, const float4 L1, const float4 L2, const float4 L3, const float4 L4,
                                                 const ScanRay& r, float dlu, float& lo, float& hi) {
  hi = INFINITY;
  if (cls == 0) {
    const float4 C0 = L0, C1 = L1;
    const float ocx = r.o.x - C0.x, ocy = r.o.y - C0.y, ocz = r.o.z - C0.z;
    const float b = __fmaf_rn(ocx, r.d.x, __fmaf_rn(ocy, r.d.y, ocz * r.d.z));
    const float oc2 = __fmaf_rn(ocx, ocx, __fmaf_rn(ocy, ocy, ocz * ocz));
    const float R2 = __fmaf_rn(C1.y, oc2, __fmaf_rn(C1.x, r.w, C0.w));
    const float b2 = b * b;
    const float disc = __fmaf_rn(r.a, R2 - oc2, b2);
    if (disc < 0.0f) return false;
    const float sd = mufu_sqrt(disc), ia = mufu_rcp(r.a);
    if ((sd - b) * ia < 0.0f) return false;
    const float tin = (-b - sd) * ia, ew = __fmaf_rn(C1.w, r.w, C1.z);
    lo = __fmaf_rn(tin, r.dl, -ew);
    const float R2d = __fmaf_rn(-C1.y, oc2, __fmaf_rn(-C1.x, r.w, L2.x));
    const float discd = __fmaf_rn(r.a, R2d - oc2, b2);
    if (discd >= 0.0f && tin > 0.0f) hi = __fmaf_rn((-b - mufu_sqrt(discd)) * ia, dlu, ew);
    return true;
  }
  if (cls == 2) {
    const float4 C0 = L0, C1 = L1, C2 = L2;
    const float cx = __fmaf_rn(C0.x, r.id.x, -r.od.x), cy = __fmaf_rn(C0.y, r.id.y, -r.od.y), cz = __fmaf_rn(C0.z, r.id.z, -r.od.z);
    const float hx = __fmaf_rn(C2.x, r.w, C1.x), hy = __fmaf_rn(C2.y, r.w, C1.y), hz = __fmaf_rn(C2.z, r.w, C1.z);
    const float ax = fabsf(r.id.x), ay = fabsf(r.id.y), az = fabsf(r.id.z);
    const float tnear = fmaxf(fmaxf(__fmaf_rn(hx, -ax, cx), __fmaf_rn(hy, -ay, cy)), __fmaf_rn(hz, -az, cz));
    const float tfar = fminf(fminf(__fmaf_rn(hx, ax, cx), __fmaf_rn(hy, ay, cy)), __fmaf_rn(hz, az, cz));
    if (tnear > tfar || tfar < 0.0f) return false;
    const float ew = __fmaf_rn(C1.w, r.w, C0.w);
    lo = __fmaf_rn(tnear, r.dl, -ew);
    const float4 C3 = L3;
    const float hxd = __fmaf_rn(-C2.x, r.w, C3.x), hyd = __fmaf_rn(-C2.y, r.w, C3.y), hzd = __fmaf_rn(-C2.z, r.w, C3.z);
    const float tnd = fmaxf(fmaxf(__fmaf_rn(hxd, -ax, cx), __fmaf_rn(hyd, -ay, cy)), __fmaf_rn(hzd, -az, cz));
    const float tfd = fminf(fminf(__fmaf_rn(hxd, ax, cx), __fmaf_rn(hyd, ay, cy)), __fmaf_rn(hzd, az, cz));
    if (tnd <= tfd && tnear > 0.0f) hi = __fmaf_rn(tnd, dlu, ew);
    return true;
  }
  // object-space classes: (ro, rw) = inverseTransform * (o, d), fused
  const float4 A0 = L0, A1 = L1, A2 = L2, K0 = L3, K1 = L4;
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
    const float b2 = b * b;
    const float disc = __fmaf_rn(a, R2 - ro2, b2);
    if (disc < 0.0f) return false;
    const float sd = mufu_sqrt(disc), ia = mufu_rcp(a);
    if ((sd - b) * ia < 0.0f) return false;
    const float tin = (-b - sd) * ia, ew = __fmaf_rn(K1.x, r.w, K0.w);
    lo = __fmaf_rn(tin, r.dl, -ew);
    const float R2d = __fmaf_rn(-K0.z, ro2, __fmaf_rn(-K0.y, r.w, 0.5f - K0.x));  // (1/4 - delta: 1/2 - K0.x is exact)
    const float discd = __fmaf_rn(a, R2d - ro2, b2);
    if (discd >= 0.0f && tin > 0.0f) hi = __fmaf_rn((-b - mufu_sqrt(discd)) * ia, dlu, ew);
    return true;
  }
  const float hx = __fmaf_rn(K1.x, r.w, K0.x), hy = __fmaf_rn(K1.y, r.w, K0.y), hz = __fmaf_rn(K1.z, r.w, K0.z);
  const float ix = mufu_rcp(rwx), iy = mufu_rcp(rwy), iz = mufu_rcp(rwz);
  const float aix = fabsf(ix), aiy = fabsf(iy), aiz = fabsf(iz);
  const float cx = -rox * ix, cy = -roy * iy, cz = -roz * iz;
  const float tnear = fmaxf(fmaxf(__fmaf_rn(hx, -aix, cx), __fmaf_rn(hy, -aiy, cy)), __fmaf_rn(hz, -aiz, cz));
  const float tfar = fminf(fminf(__fmaf_rn(hx, aix, cx), __fmaf_rn(hy, aiy, cy)), __fmaf_rn(hz, aiz, cz));
  if (tnear > tfar || tfar < 0.0f) return false;
  const float ew = __fmaf_rn(K1.w, r.w, K0.w);
  lo = __fmaf_rn(tnear, r.dl, -ew);
  // (half extents 1/2 - delta_i: 1 - K0.i is exact)
  const float hxd = __fmaf_rn(-K1.x, r.w, 1.0f - K0.x), hyd = __fmaf_rn(-K1.y, r.w, 1.0f - K0.y), hzd = __fmaf_rn(-K1.z, r.w, 1.0f - K0.z);
  const float tnd = fmaxf(fmaxf(__fmaf_rn(hxd, -aix, cx), __fmaf_rn(hyd, -aiy, cy)), __fmaf_rn(hzd, -aiz, cz));
  const float tfd = fminf(fminf(__fmaf_rn(hxd, aix, cx), __fmaf_rn(hyd, aiy, cy)), __fmaf_rn(hzd, aiz, cz));
  if (tnd <= tfd && tnear > 0.0f && (aix + aiy) + aiz < INFINITY) hi = __fmaf_rn(tnd, dlu, ew);
  return true;
}

__device__ __forceinline__ bool leaf_filter(int cls, const float4* __restrict__ L, const ScanRay& r, float dlu, float& lo, float& hi) {
  const float4 z = make_float4(0, 0, 0, 0);
  const float4 L0 = __ldg(L), L1 = __ldg(L + 1), L2 = __ldg(L + 2);  // (class 0 reads three rows, class 2 four, the others five)
  const float4 L3 = cls != 0 ? __ldg(L + 3) : z, L4 = (cls & 1) ? __ldg(L + 4) : z;
  return leaf_filter_rows(cls, L0, L1, L2, L3, L4, r, dlu, lo, hi);
}

// Entry parameters of the ray into the two children's padded boxes (+inf = the ray provably misses the box), both
// children at once: a node stores its children side by side, so every add / multiply is one packed f32x2
// instruction (FADD2 / FMUL2 / FFMA2; min / max have no packed form and stay scalar).
// A NaN (0 * inf) is ignored by fminf / fmaxf, so the affected slab does not constrain: conservative.
__device__ __forceinline__ void child_entries(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const ScanRay& r,
                                              float& e0, float& e1) {
  // box planes relative to the origin, subtraction first: with d_i = 0 the two planes of a slab give -inf / +inf
  // (origin inside the slab: no constraint) or the same infinity twice (outside: miss); b/d - o/d would turn the
  // first case into inf - inf
  const f2 nox = bc2(-r.o.x), noy = bc2(-r.o.y), noz = bc2(-r.o.z);
  const f2 lx = __fadd2_rn(lo2(n0), nox), ly = __fadd2_rn(hi2(n0), noy), lz = __fadd2_rn(lo2(n1), noz);
  const f2 hx = __fadd2_rn(hi2(n1), nox), hy = __fadd2_rn(lo2(n2), noy), hz = __fadd2_rn(hi2(n2), noz);
  // D = largest distance from the origin to a point of the box (bounds |o - c| of every geom inside); the rounding
  // of these few operations is part of P2 (x 1.000002 on the host)
  const f2 mx = make_float2(fmaxf(fabsf(lx.x), fabsf(hx.x)), fmaxf(fabsf(lx.y), fabsf(hx.y)));
  const f2 my = make_float2(fmaxf(fabsf(ly.x), fabsf(hy.x)), fmaxf(fabsf(ly.y), fabsf(hy.y)));
  const f2 mz = make_float2(fmaxf(fabsf(lz.x), fabsf(hz.x)), fmaxf(fabsf(lz.y), fabsf(hz.y)));
  const f2 D2 = fma2(mx, mx, fma2(my, my, mul2(mz, mz)));
  const f2 pad = fma2(hi2(n3), D2, mul2(lo2(n3), bc2(r.w)));
  const f2 npad = neg2(pad), idx = bc2(r.id.x), idy = bc2(r.id.y), idz = bc2(r.id.z);
  const f2 ax = mul2(__fadd2_rn(lx, npad), idx), bx = mul2(__fadd2_rn(hx, pad), idx);
  const f2 ay = mul2(__fadd2_rn(ly, npad), idy), by = mul2(__fadd2_rn(hy, pad), idy);
  const f2 az = mul2(__fadd2_rn(lz, npad), idz), bz = mul2(__fadd2_rn(hz, pad), idz);
  // rounding of the six parameters (a few ulp of |box - o| / |d|) is covered by the 16u*w in P1
  const float tn0 = fmaxf(fmaxf(fminf(ax.x, bx.x), fminf(ay.x, by.x)), fminf(az.x, bz.x));
  const float tf0 = fminf(fminf(fmaxf(ax.x, bx.x), fmaxf(ay.x, by.x)), fmaxf(az.x, bz.x));
  const float tn1 = fmaxf(fmaxf(fminf(ax.y, bx.y), fminf(ay.y, by.y)), fminf(az.y, bz.y));
  const float tf1 = fminf(fminf(fmaxf(ax.y, bx.y), fmaxf(ay.y, by.y)), fmaxf(az.y, bz.y));
  e0 = (tn0 > tf0 || tf0 < 0.0f) ? INFINITY : fmaxf(tn0, 0.0f);  // NaN -> 0: "may be entered at once"
  e1 = (tn1 > tf1 || tf1 < 0.0f) ? INFINITY : fmaxf(tn1, 0.0f);
}

// per-ray constants of a traversal
struct TravRay {
  float ewmax;  // largest world slack of any geom, plus room for the rounding differences between a node's slab
                // parameters and a leaf's own entry parameter (a few ulp of w)
  float dls;    // |d|, rounded down a little further
  float dlu;    // |d|, rounded up with room for the approximations of a sure hit's entry parameter; +inf = no sure hits for
                // this ray (a zero direction component)
};
__device__ __forceinline__ TravRay make_trav_ray(const BvhSoA& B, const ScanRay& r) {
  TravRay t;
  t.ewmax = __fmaf_rn(B.ew_w_max, r.w, B.ew_c_max) + 1e-6f * r.w;
  t.dls = r.dl * 0.999996f;
  t.dlu = PT_BVH_SURE && (fabsf(r.id.x) + fabsf(r.id.y)) + fabsf(r.id.z) < INFINITY ? r.dl * 1.000006f : INFINITY;
  return t;
}
__device__ __forceinline__ int bvh_root(const BvhSoA& B) { return B.root; }

// can anything below a box entered at parameter `e` still matter?
//   filter pass: no if its best possible bound is >= lo3 (it changes none of k1, k2, lo2, lo3) or > hi (a geom that is
//                surely hit lies closer than everything in the box: pt_kernels.cuh resolves among the geoms with bound <= hi);
//   exact pass:  no if the bound exceeds the best exact distance (ties may still win).
template <bool EXACT>
__device__ __forceinline__ bool can_matter(float e, const TravRay& tr, const ScanBest& best, const Hit& h) {
  const float bd = __fmaf_rn(e, tr.dls, -tr.ewmax);
  return EXACT ? !(bd > h.t) : (bd < best.lo3 && !(bd > best.hi));
}

// the leaf `cur` (< 0): filter test (EXACT: exact test of the candidate if it can still matter)
template <bool EXACT>
__device__ __forceinline__ void leaf_visit(const BvhSoA& B, const GeomSoA& g, const ScanRay& r, const TravRay& tr, ScanBest& best, Hit& h,
                                           int cur) {
  const int leaf = ~cur & ((1 << kBvhLeafBits) - 1), cls = ~cur >> kBvhLeafBits;
  float lo, hi;
  PT_HIST(41);
  if (leaf_filter(cls, B.leaves + (size_t)leaf * kBvhLeafRows, r, tr.dlu, lo, hi)) {
    lo = fmaxf(lo, 0.0f);
    if (hi < INFINITY) PT_HIST(42);
    if (!EXACT) {
      scan_take3(best, lo, leaf);
      best.hi = fminf(best.hi, hi);  // (NaN is ignored)
    } else if (!(lo > h.t)) {
      const int gi = __ldg(B.leaf_meta + leaf).y;
      float dist;
      f3 P;
      int ncode;
      if (exact_hit(cls < 2 ? 0 : 1, __ldg(g.inv0 + gi), __ldg(g.inv1 + gi), __ldg(g.inv2 + gi), __ldg(g.fwd0 + gi),
                    __ldg(g.fwd1 + gi), __ldg(g.fwd2 + gi), r.o, r.d, dist, P, ncode)) {
        // specification: scan in index order, keep the strictly smaller positive distance
        if (dist > 0 && (dist < h.t || (dist == h.t && gi < h.id))) { h.t = dist; h.id = gi; h.p = P; h.ncode = ncode; }
      }
    }
  }
}

// a stack entry: (child reference, bits of the entry parameter into its box -- checked again when the entry is popped:
// the limits have usually tightened since it was pushed, and dropping it there saves the fetch of a whole node)
typedef uint2 StackEnt;
// A lane's traversal stack: the lowest S levels in the warp's shared memory (row = level, column = lane: conflict-free,
// and a pop does not wait for L2 -- the local-memory lines of a stack do not survive in an L1 that the nodes stream
// through: 8.5 % hit rate, profiles/r02_bvh_notes.txt), the rest in local memory (10 000 geoms: 99.95 % of the pushes find
// the stack at most 7 deep).  S = 0: local memory only.
template <int S>
struct TravStack {
  StackEnt* sm;  // this lane's column of the warp's S x 32 shared-memory entries (S > 0)
  StackEnt* ov;  // kBvhStack - S entries of local memory (an array of the caller's: a member array would drag `sp` into
                 // local memory with it)
  int sp = 0;
  // (Measured and dropped: the top entry in registers, a pop handing it out at once and requesting the entry below for the
  // NEXT pop -- 2.0 instead of 3.0 Gseg/s: the outstanding load's scoreboard stalls the step it was meant to overlap.)
  __device__ __forceinline__ void push(StackEnt e) {
    PT_CHECK(sp < kBvhStack);
    if (S > 0 && sp < S) sm[sp * 32] = e; else ov[sp - S] = e;
    sp++;
  }
  __device__ __forceinline__ StackEnt pop() {
    --sp;
    if (S > 0 && sp < S) return sm[sp * 32];
    return ov[sp - S];
  }
};


template <bool EXACT, typename Stack>
__device__ __forceinline__ bool node_rows(const float4 a0, const float4 a1, const float4 a2, const float4 a3, const float4 b0, const float4 b1,
                                          const float4 b2, const float4 b3, const int4 ch, const ScanRay& r, const TravRay& tr,
                                          const ScanBest& best, const Hit& h, int& cur, Stack& st);
// the inner node `cur` (>= 0): test its (up to four) children; the nearest one that can matter becomes `cur`, the others
// go on the stack farthest first.  Returns false if no child can matter (the caller pops).
template <bool EXACT, typename Stack>
__device__ __forceinline__ bool node_visit(const BvhSoA& B, const ScanRay& r, const TravRay& tr, const ScanBest& best, const Hit& h,
                                           int& cur, Stack& st) {
  const float4* N = B.nodes + (size_t)cur * kBvhNodeRows;
#if PT_BVH_NODE128
  const float4 a0 = __ldg(N), a1 = __ldg(N + 1), a2 = __ldg(N + 2), b0 = __ldg(N + 3), b1 = __ldg(N + 4), b2 = __ldg(N + 5);
  const int4 ch = __ldg(reinterpret_cast<const int4*>(N + 6));
  const float4 pd = __ldg(N + 7);
  const float4 a3 = make_float4(pd.x, pd.x, pd.y, pd.y), b3 = a3;
#else
  const float4 a0 = __ldg(N), a1 = __ldg(N + 1), a2 = __ldg(N + 2), a3 = __ldg(N + 3);
  const float4 b0 = __ldg(N + 4), b1 = __ldg(N + 5), b2 = __ldg(N + 6), b3 = __ldg(N + 7);
  const int4 ch = __ldg(reinterpret_cast<const int4*>(N + 8));
#endif
  return node_rows<EXACT>(a0, a1, a2, a3, b0, b1, b2, b3, ch, r, tr, best, h, cur, st);
}
// ... the same with the node's nine rows already fetched
template <bool EXACT, typename Stack>
__device__ __forceinline__ bool node_rows(const float4 a0, const float4 a1, const float4 a2, const float4 a3, const float4 b0, const float4 b1,
                                          const float4 b2, const float4 b3, const int4 ch, const ScanRay& r, const TravRay& tr,
                                          const ScanBest& best, const Hit& h, int& cur, Stack& st) {
  PT_HIST(40);
  float e[4];
  child_entries(a0, a1, a2, a3, r, e[0], e[1]);
  child_entries(b0, b1, b2, b3, r, e[2], e[3]);
  // sort keys: the entry parameter's bits (>= 0, so they order like integers) with the slot number in the two lowest
  // bits; a child that cannot matter gets the largest key
  // (an empty slot is excluded by its child word, not by its box: the pad of a box at infinity is 0 * inf = NaN, which
  // child_entries reads as "may be entered")
  const int chs[4] = {ch.x, ch.y, ch.z, ch.w};
  uint32_t k[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const bool v = e[i] < INFINITY && can_matter<EXACT>(e[i], tr, best, h) && chs[i] != kBvhNoChild;
    k[i] = v ? ((__float_as_uint(e[i]) & ~3u) | (uint32_t)i) : 0xffffffffu;
  }
  // sorting network on four keys (ascending)
#define PT_CSWAP(x, y) { const uint32_t lo_ = min(k[x], k[y]), hi_ = max(k[x], k[y]); k[x] = lo_; k[y] = hi_; }
  PT_CSWAP(0, 1) PT_CSWAP(2, 3) PT_CSWAP(0, 2) PT_CSWAP(1, 3) PT_CSWAP(1, 2)
#undef PT_CSWAP
  if (k[0] == 0xffffffffu) return false;
  auto child_of = [&](uint32_t key) { const uint32_t s_ = key & 3u; return s_ == 0 ? ch.x : (s_ == 1 ? ch.y : (s_ == 2 ? ch.z : ch.w)); };
#pragma unroll
  for (int i = 3; i >= 1; i--) {
    // (predicated pushes without the branch were measured: +4 % instructions, -1 % throughput)
    if (k[i] != 0xffffffffu) {
#ifdef PT_BVH_STACK_HIST
      atomicAdd(&g_sp_hist[st.sp < 39 ? st.sp : 39], 1ull);
#endif
      st.push(make_uint2((uint32_t)child_of(k[i]), k[i] & ~3u));  // (the entry parameter, rounded down by the key)
    }
  }
  cur = child_of(k[0]);
  return true;
}

// One step of a traversal: the leaf or the node `cur`, then move on.  Returns false when the traversal is finished.
// EXACT = false: filter scan, result in `best` (k1 = leaf index).  EXACT = true: exact test of every candidate leaf that
// can still matter, result in `h`.  (The builder guarantees that the stack never needs more than kBvhStack entries.)
template <bool EXACT, typename Stack>
__device__ __forceinline__ bool trav_step(const BvhSoA& B, const GeomSoA& g, const ScanRay& r, const TravRay& tr, ScanBest& best, Hit& h,
                                          int& cur, Stack& st) {
  if (cur < 0) leaf_visit<EXACT>(B, g, r, tr, best, h, cur);
  else if (node_visit<EXACT>(B, r, tr, best, h, cur, st)) return true;
  while (st.sp > 0) {
    const StackEnt ent = st.pop();
    PT_HIST(44);
    if (can_matter<EXACT>(__uint_as_float(ent.y), tr, best, h)) { cur = (int)ent.x; return true; }
    PT_HIST(43);
  }
  return false;
}

// One step of the FILTER pass as k_bounce_bvh runs it, lanes of a warp side by side: whatever a lane stands at -- an inner
// node (nine rows) or a leaf (five) -- its rows are requested FIRST, by all lanes together, and only then do the lanes
// part ways into the node code and the leaf classes' code, which are pure arithmetic.  A step of the warp thus waits for
// memory once; with the fetches inside the diverged branches it waited once per branch (node, sphere leaf, cube leaf,
// and the leaf's meta word before that), one L2 round trip after the other: 4 400 cycles per step
// (profiles/r02_bvh_notes.txt).  Returns false when the lane's traversal is finished.
template <typename Stack>
__device__ __forceinline__ bool filter_step(const BvhSoA& B, const ScanRay& r, const TravRay& tr, ScanBest& best, int& cur, Stack& st) {
  const bool at_node = cur >= 0;
  const int leaf = ~cur & ((1 << kBvhLeafBits) - 1), cls = ~cur >> kBvhLeafBits;
  const float4* p = at_node ? B.nodes + (size_t)cur * kBvhNodeRows : B.leaves + (size_t)leaf * kBvhLeafRows;
  const float4 z = make_float4(0, 0, 0, 0);
  const float4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3), q4 = __ldg(p + 4);
  const float4 q5 = at_node ? __ldg(p + 5) : z, q6 = at_node ? __ldg(p + 6) : z, q7 = at_node ? __ldg(p + 7) : z;
#if !PT_BVH_NODE128
  const float4 q8 = at_node ? __ldg(p + 8) : z;
#endif
  const Hit unused{};
  if (at_node) {
#if PT_BVH_NODE128
    const int4 ch = make_int4(__float_as_int(q6.x), __float_as_int(q6.y), __float_as_int(q6.z), __float_as_int(q6.w));
    const float4 pd = make_float4(q7.x, q7.x, q7.y, q7.y);
    if (node_rows<false>(q0, q1, q2, pd, q3, q4, q5, pd, ch, r, tr, best, unused, cur, st)) return true;
#else
    const int4 ch = make_int4(__float_as_int(q8.x), __float_as_int(q8.y), __float_as_int(q8.z), __float_as_int(q8.w));
    if (node_rows<false>(q0, q1, q2, q3, q4, q5, q6, q7, ch, r, tr, best, unused, cur, st)) return true;
#endif
  } else {
    float lo, hi;
    PT_HIST(41);
    if (leaf_filter_rows(cls, q0, q1, q2, q3, q4, r, tr.dlu, lo, hi)) {
      scan_take3(best, lo, leaf);
      best.hi = fminf(best.hi, hi);  // (NaN is ignored)
    }
  }
  while (st.sp > 0) {
    const StackEnt ent = st.pop();
    PT_HIST(44);
    if (can_matter<false>(__uint_as_float(ent.y), tr, best, unused)) { cur = (int)ent.x; return true; }
    PT_HIST(43);
  }
  return false;
}

// a whole traversal by one lane (parity entry points, the exact traversal of deferred paths)
template <bool EXACT>
__device__ __forceinline__ void bvh_traverse(const BvhSoA& B, const GeomSoA& g, const ScanRay& r, ScanBest& best, Hit& h) {
  if (B.n_leaves <= 0) return;
  const TravRay tr = make_trav_ray(B, r);
  StackEnt ov[kBvhStack];
  TravStack<0> st;
  st.sm = nullptr; st.ov = ov;
  int cur = bvh_root(B);
  if (EXACT) {
    while (trav_step<true>(B, g, r, tr, best, h, cur, st)) {}
  } else {
    while (filter_step(B, r, tr, best, cur, st)) {}
  }
}

// the exact test of leaf k on its own: false = the reference's test reports no hit (or a distance <= 0)
__device__ __forceinline__ bool exact_leaf(int k, const BvhSoA& B, const GeomSoA& g, f3 o, f3 d, Hit& e) {
  const int2 meta = __ldg(B.leaf_meta + k);
  const int gi = meta.y;
  e.id = gi;
  const bool hit = exact_hit(meta.x < 2 ? 0 : 1, __ldg(g.inv0 + gi), __ldg(g.inv1 + gi), __ldg(g.inv2 + gi), __ldg(g.fwd0 + gi),
                             __ldg(g.fwd1 + gi), __ldg(g.fwd2 + gi), o, d, e.t, e.p, e.ncode);
  return hit && e.t > 0;
}

// the exact pass on its own (fallback of resolve_bvh; rare, so not inlined)
__device__ __noinline__ void bvh_exact(const BvhSoA B, const GeomSoA g, f3 o, f3 d, float r_scene, Hit& h) {
  const ScanRay r = make_scan_ray(o, d, r_scene, true);
  ScanBest unused;
  scan_init(unused);
  bvh_traverse<true>(B, g, r, unused, h);
}

// the exact test of leaf k on its own, kept out of line (k_bounce_bvh tests two candidates per path: one copy of the code)
__device__ __noinline__ bool exact_leaf_call(int k, const BvhSoA B, const GeomSoA g, f3 o, f3 d, Hit& e) { return exact_leaf(k, B, g, o, d, e); }

// The exact test of the filter pass's best candidate leaf k1: confirmed (h filled in) if it is a hit closer than every
// other geom's lower bound lo2.
__device__ __forceinline__ bool confirm_candidate(int k1, float lo2, const BvhSoA& B, const GeomSoA& g, f3 o, f3 d, Hit& h) {
  const int2 meta = __ldg(B.leaf_meta + k1);
  const int gi = meta.y;
  float dist;
  f3 P;
  int ncode;
  const bool hit = exact_hit(meta.x < 2 ? 0 : 1, __ldg(g.inv0 + gi), __ldg(g.inv1 + gi), __ldg(g.inv2 + gi), __ldg(g.fwd0 + gi),
                             __ldg(g.fwd1 + gi), __ldg(g.fwd2 + gi), o, d, dist, P, ncode);
  if (!(hit && dist > 0 && dist < lo2)) return false;
  h.t = dist; h.id = gi; h.p = P; h.ncode = ncode;
  return true;
}

// Resolve a finished BVH filter pass on the spot (parity entry point; k_bounce_bvh defers the fallback instead):
// the confirmed candidate, otherwise the exact pass.  Returns true if the fallback ran (statistics only).
__device__ __forceinline__ bool resolve_bvh(const ScanBest& best, const BvhSoA& B, const GeomSoA& g, float r_scene, f3 o, f3 d, Hit& h) {
  if (best.k1 < 0) return false;  // every geom is a proven miss
  if (confirm_candidate(best.k1, best.lo2, B, g, o, d, h)) return false;
  bvh_exact(B, g, o, d, r_scene, h);
  return true;
}

}  // namespace ptd
