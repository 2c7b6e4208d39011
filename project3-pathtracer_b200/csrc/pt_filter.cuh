// pt_filter.cuh -- closest hit as "cheap conservative scan + exact evaluation of the winner" (sm_100a).
//
// The arithmetic contract (pt_device.cuh) makes one exact intersection test cost ~200 instructions (unfused
// binary32, IEEE sqrt and division, a binary64 step, two normalisations, a round trip through the forward
// transform).  Running it against every geom made k_bounce issue-bound at a tenth of the HBM roofline
// (profiles/r01_k_bounce_v1_metrics.txt).  Here every geom instead goes through a FILTER: ~45 instructions of fused
// multiply-adds and MUFU approximations that either prove "the exact test cannot report a hit" or give a LOWER
// BOUND on the world distance the exact test would report.  The scan keeps the geom with the smallest lower bound
// (k1) and the second-smallest lower bound (lo2).  The exact test (exact_hit, the reference's arithmetic) then runs
// ONCE, on k1; if it reports a hit closer than lo2, every other geom is provably farther and k1 is the exact
// scan's answer, bit for bit.  Otherwise (two surfaces closer together than the bounds can separate, or k1 was a
// near miss) the ray falls back to the exact scan over all geoms.  The result is therefore always the exact
// scan's result; the filter only decides how much work it takes.
//
// What the filter must guarantee (error model in DESIGN.md "filter"): with E = the exact path's own floating-point
// evaluation of a geom and F = the filter's,
//   (1) if E reports a hit, F does not report "miss";
//   (2) if E reports a hit at world distance dist_E, F's lower bound lo <= dist_E.
// Both come from testing the ray against an INFLATED unit shape: per-axis half extents 0.5 + delta_i for the cube,
// radius^2 0.25 + delta for the sphere, where the deltas bound the distance (in object units) between the point E
// computes and the point F computes for the same ray parameter, plus both paths' rounding in the test itself.
// They are linear in w = 4*max|o_j| + R_scene (and, for the sphere, in |ro|^2, because the reference's radicand
// b^2 - (|ro|^2 - r^2) loses |ro|^2 * 2^-24 to cancellation) with per-geom coefficients computed on the host in
// binary64 (pt_api.cu: build_filter).  A world-space slack E_w (pull-back of 1e-4 object units, rounding of the
// forward transform, residual of transform * inverseTransform - I) is subtracted from the entry distance.
//
// Reference functions behind the exact test: sphereIntersectionTest src/intersections.h:81-117,
// boxIntersectionTest (stub) :74-77, multiplyMV :53-59, getPointOnRay :46-48.
#pragma once
#include "pt_device.cuh"

namespace ptd {

__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- filter geometry: PAIRS of geoms of one class (MESH left out) ----
// Blackwell's packed FFMA2 / FMUL2 / FADD2 take one issue slot for two independent binary32 operations, so the
// filter tests two geoms at once: geom A in the low half, geom B in the high half of every float2.  An odd geom
// out is paired with a copy of itself whose constants are -inf, which the tests below reject (proven miss).
//
// Four classes, in this order; the first and third need no object-space transform at all:
//   class 0  uniformly scaled spheres: the inflated unit sphere is a SPHERE in world space
//            (centre c, radius^2 = Wc + Ww*w + Wr*|o-c|^2); 15 packed instructions per pair
//   class 1  other spheres (ellipsoids): object-space test on (ro, rw) = inverseTransform * (o, d)
//   class 2  cubes whose axes are the world axes (up to a slop far below the other slacks; rotations by multiples of
//            90 degrees): slab test against the world AABB of the inflated cube (centre c, half extents Hc + Hw*w)
//            with the ray's own 1/d, computed once per ray; 13 packed instructions per pair
//   class 3  other cubes: object-space slab test
// Records are 10 consecutive float4 per pair (every lane reads the same pair: shared-memory broadcasts, one pointer
// walks the list).  "|" separates the two float2 halves of a float4, each half = (A, B):
//   class 0: r0 = (c.x | c.y)  r1 = (c.z | Wc)  r2 = (Ww | Wr)  r3 = (Ew_c | Ew_w)
//   class 2: r0 = (c.x | c.y)  r1 = (c.z | Hc.x)  r2 = (Hc.y | Hc.z)  r3 = (Hw.x | Hw.y)  r4 = (Hw.z | Ew_c)  r5 = (Ew_w | -)
//   class 1, 3: r[2k] = (A[k][0], B[k][0], A[k][1], B[k][1])   r[2k+1] = (A[k][2], B[k][2], A[k][3], B[k][3])   k = 0,1,2:
//               rows x,y,z of inverseTransform, interleaved; then
//     class 3: r6 = (hc.x | hc.y)  r7 = (hc.z | hw.x)  r8 = (hw.y | hw.z)  r9 = (Ew_c | Ew_w)     half extent_i = hc_i + hw_i * w
//     class 1: r6 = (R2c | R2w)    r7 = (R2r | Ew_c)   r8 = (Ew_w | -)                            radius^2 = R2c + R2w * w + R2r * |ro|^2
//   world slack E_w = Ew_c + Ew_w * w
constexpr int kFiltRows = 10;
constexpr int kFiltClasses = 4;
struct FiltSoA {  // in HBM
  const float4* rows;       // [n_pairs][kFiltRows]
  const int2* ids;          // pair -> (geom index of A, geom index of B)
  int end[kFiltClasses];    // end[k] = first pair index after class k (end[3] = number of pairs)
  float r_scene;            // bound on |p| over all surface points of the scene
};
__host__ __device__ inline size_t filt_smem_bytes(int cap) { return (size_t)cap * kFiltRows * sizeof(float4); }
// cooperative copy of pairs [first, first+count) into shared memory; caller synchronises
__device__ __forceinline__ void stage_filt(const FiltSoA& g, int first, int count, float4* smem) {
  for (int i = threadIdx.x; i < count * kFiltRows; i += blockDim.x) smem[i] = g.rows[(size_t)first * kFiltRows + i];
}

typedef float2 f2;
__device__ __forceinline__ f2 bc2(float x) { return make_float2(x, x); }
__device__ __forceinline__ f2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ f2 hi2(float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }   // folds into the consumer's operand modifier
__device__ __forceinline__ f2 abs2(f2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }

// per-ray constants of the scan
struct ScanRay {
  f3 o, d;       // the ray (each component is broadcast to both halves by the packed instructions)
  float w;       // 4*max|o_j| + R_scene
  float dl;      // |d| rounded down (parameter -> world distance)
  float a;       // d.d
  f3 id, od;     // 1/d (MUFU) and o/d: the ray's side of the world-space slab test (class 2)
};
__device__ __forceinline__ ScanRay make_scan_ray(f3 o, f3 d, float r_scene, bool need_inverse) {
  ScanRay r;
  r.o = o; r.d = d;
  const float omax = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
  r.w = __fmaf_rn(4.0f, omax, r_scene);
  r.a = __fmaf_rn(d.x, d.x, __fmaf_rn(d.y, d.y, d.z * d.z));
  r.dl = mufu_sqrt(r.a) * 0.99999905f;  // 1 - 2^-20: below |d| whatever the approximation error
  r.id = mk(0, 0, 0); r.od = mk(0, 0, 0);
  if (need_inverse) {
    r.id = mk(mufu_rcp(d.x), mufu_rcp(d.y), mufu_rcp(d.z));
    r.od = mk(o.x * r.id.x, o.y * r.id.y, o.z * r.id.z);
  }
  return r;
}

// running result of the scan
struct ScanBest {
  float lo1, lo2;  // smallest and second-smallest lower bound
  int k1;          // 2*pair + half of the smallest, -1 = every geom so far is a proven miss
  // hierarchy only (pt_bvh.cuh):
  float hi;        // smallest UPPER bound on the exact distance of a geom that is surely hit
  float lo3;       // third-smallest lower bound
  int k2;          // the geom with the second-smallest bound (-1: none)
};
__device__ __forceinline__ void scan_init(ScanBest& b) {
  b.lo1 = INFINITY; b.lo2 = INFINITY; b.k1 = -1; b.hi = INFINITY; b.lo3 = INFINITY; b.k2 = -1;
}
// the best THREE bounds and the two geoms of the smaller ones (hierarchy, pt_bvh.cuh)
__device__ __forceinline__ void scan_take3(ScanBest& b, float lo, int k) {
  lo = fmaxf(lo, 0.0f);  // (NaN -> 0: "no information")
  const bool first = lo < b.lo1, second = !first && lo < b.lo2;
  b.lo3 = first || second ? b.lo2 : fminf(b.lo3, lo);
  b.lo2 = first ? b.lo1 : (second ? lo : b.lo2);
  b.k2 = first ? b.k1 : (second ? k : b.k2);
  b.lo1 = first ? lo : b.lo1;
  b.k1 = first ? k : b.k1;
}
__device__ __forceinline__ void scan_take(ScanBest& b, float lo, int k) {
  lo = fmaxf(lo, 0.0f);  // also turns a NaN bound into 0 (fmaxf ignores NaN): "no information"
  const bool better = lo < b.lo1;
  b.lo2 = fminf(b.lo2, better ? b.lo1 : lo);
  b.k1 = better ? k : b.k1;
  b.lo1 = fminf(b.lo1, lo);
}

// object-space origin and UN-normalised direction of both geoms with fused multiply-adds: the parameter along
// (ro, rw) is the parameter along the world ray
#define PT_FILT_TRANSFORM(V, R)                                                                          \
  const float4 Q0 = (V)[0], Q1 = (V)[1], Q2 = (V)[2], Q3 = (V)[3], Q4 = (V)[4], Q5 = (V)[5];             \
  const f2 rox = fma2(lo2(Q0), bc2((R).o.x), fma2(hi2(Q0), bc2((R).o.y), fma2(lo2(Q1), bc2((R).o.z), hi2(Q1)))); \
  const f2 roy = fma2(lo2(Q2), bc2((R).o.x), fma2(hi2(Q2), bc2((R).o.y), fma2(lo2(Q3), bc2((R).o.z), hi2(Q3)))); \
  const f2 roz = fma2(lo2(Q4), bc2((R).o.x), fma2(hi2(Q4), bc2((R).o.y), fma2(lo2(Q5), bc2((R).o.z), hi2(Q5)))); \
  const f2 rwx = fma2(lo2(Q0), bc2((R).d.x), fma2(hi2(Q0), bc2((R).d.y), mul2(lo2(Q1), bc2((R).d.z))));          \
  const f2 rwy = fma2(lo2(Q2), bc2((R).d.x), fma2(hi2(Q2), bc2((R).d.y), mul2(lo2(Q3), bc2((R).d.z))));          \
  const f2 rwz = fma2(lo2(Q4), bc2((R).d.x), fma2(hi2(Q4), bc2((R).d.y), mul2(lo2(Q5), bc2((R).d.z))));

// Scan pairs [first, last) (global pair indices); `v` points at the record of pair `first`; end[k] = first pair
// index after class k.  Every "miss" needs a comparison to come out TRUE, so a NaN anywhere keeps the geom as a
// candidate.
// Branch-free halves: a proven miss is handed to the sink as such and changes nothing there.  Some lane of a warp nearly
// always needs the body, so branches bought nothing but BSSY / BRA / BSYNC.
// (sqrt of a negative discriminant is NaN, and NaN < 0 is false: the decision is the first comparison's.)
#define PT_SPHERE_HALF(H)                                                                             \
  const float sd_##H = mufu_sqrt(disc.H), ia_##H = mufu_rcp(a.H);                                     \
  const bool miss_##H = (disc.H < 0.0f) || ((sd_##H - b.H) * ia_##H < 0.0f); /* 2nd: the sphere lies behind */ \
  const float lo_##H = __fmaf_rn((-b.H - sd_##H) * ia_##H, r.dl, -ew.H);
#define PT_BOX_HALF(H)                                                                                \
  const float tnear_##H = fmaxf(fmaxf(nx.H, ny.H), nz.H), tfar_##H = fminf(fminf(fx.H, fy.H), fz.H);  \
  const bool miss_##H = (tnear_##H > tfar_##H) || (tfar_##H < 0.0f); /* misses the inflated box, or it lies behind */ \
  const float lo_##H = __fmaf_rn(tnear_##H, r.dl, -ew.H);
// (The four class loops' control is 51 of a unit's 836 instructions in k_bounce_q.  Running a class's pairs as straight-line
// code entered by pair count -- up to four copies of each body behind a switch -- was measured: the same instruction count
// (the compiler's compare chain costs what the loop control did), 40 % more code, issue-slot utilisation 68 -> 60 %:
// 25.4 instead of 29.0 Gseg/s.)
// `sink(lo_A, miss_A, lo_B, miss_B, pair)` receives both halves of every pair: lower bound, and whether the geom is ruled out
template <typename Sink>
__device__ __forceinline__ void filter_scan_to(const float4* v, int first, int last, const int end[kFiltClasses],
                                               const ScanRay& r, Sink& sink) {
  int i = first;
  const f2 w2 = bc2(r.w);
  // ---- class 0: uniformly scaled spheres, world space ----
  for (const int e = min(last, end[0]); i < e; i++, v += kFiltRows) {
    const float4 C0 = v[0], C1 = v[1], C2 = v[2], C3 = v[3];
    const f2 ocx = __fadd2_rn(bc2(r.o.x), neg2(lo2(C0))), ocy = __fadd2_rn(bc2(r.o.y), neg2(hi2(C0))),
             ocz = __fadd2_rn(bc2(r.o.z), neg2(lo2(C1)));
    const f2 b = fma2(ocx, bc2(r.d.x), fma2(ocy, bc2(r.d.y), mul2(ocz, bc2(r.d.z))));
    const f2 oc2 = fma2(ocx, ocx, fma2(ocy, ocy, mul2(ocz, ocz)));
    const f2 R2 = fma2(hi2(C2), oc2, fma2(lo2(C2), w2, hi2(C1)));
    const f2 nc = __fadd2_rn(R2, neg2(oc2));              // -(|o-c|^2 - R^2)
    const f2 disc = fma2(bc2(r.a), nc, mul2(b, b));       // (oc.d)^2 - |d|^2 (|oc|^2 - R^2)
    if (disc.x < 0.0f && disc.y < 0.0f) continue;          // both lines miss their inflated spheres
    const f2 ew = fma2(hi2(C3), w2, lo2(C3));
    const f2 a = bc2(r.a);
    PT_SPHERE_HALF(x)
    PT_SPHERE_HALF(y)
    sink(lo_x, miss_x, lo_y, miss_y, i);
  }
  // ---- class 1: other spheres, object space ----
  for (const int e = min(last, end[1]); i < e; i++, v += kFiltRows) {
    PT_FILT_TRANSFORM(v, r)
    const float4 K0 = v[6], K1 = v[7];
    const f2 a = fma2(rwx, rwx, fma2(rwy, rwy, mul2(rwz, rwz)));
    const f2 b = fma2(rox, rwx, fma2(roy, rwy, mul2(roz, rwz)));
    const f2 ro2 = fma2(rox, rox, fma2(roy, roy, mul2(roz, roz)));
    const f2 R2 = fma2(lo2(K1), ro2, fma2(hi2(K0), w2, lo2(K0)));
    const f2 nc = __fadd2_rn(R2, neg2(ro2));      // -(|ro|^2 - R^2)
    const f2 disc = fma2(a, nc, mul2(b, b));      // (ro.rw)^2 - |rw|^2 (|ro|^2 - R^2)
    if (disc.x < 0.0f && disc.y < 0.0f) continue;  // both lines miss their inflated spheres
    const f2 ew = fma2(lo2(v[8]), w2, hi2(K1));
    PT_SPHERE_HALF(x)
    PT_SPHERE_HALF(y)
    sink(lo_x, miss_x, lo_y, miss_y, i);
  }
  // ---- class 2: world-axis-aligned cubes, world space ----
  for (const int e = min(last, end[2]); i < e; i++, v += kFiltRows) {
    const float4 C0 = v[0], C1 = v[1], C2 = v[2], C3 = v[3], C4 = v[4];
    // parameter of the slab centres, (c_i - o_i) / d_i, and half widths H_i / |d_i|
    const f2 cx = fma2(lo2(C0), bc2(r.id.x), bc2(-r.od.x)), cy = fma2(hi2(C0), bc2(r.id.y), bc2(-r.od.y)),
             cz = fma2(lo2(C1), bc2(r.id.z), bc2(-r.od.z));
    const f2 hx = fma2(lo2(C3), w2, hi2(C1)), hy = fma2(hi2(C3), w2, lo2(C2)), hz = fma2(lo2(C4), w2, hi2(C2));
    const f2 aix = bc2(fabsf(r.id.x)), aiy = bc2(fabsf(r.id.y)), aiz = bc2(fabsf(r.id.z));
    const f2 nx = fma2(hx, neg2(aix), cx), ny = fma2(hy, neg2(aiy), cy), nz = fma2(hz, neg2(aiz), cz);
    const f2 fx = fma2(hx, aix, cx), fy = fma2(hy, aiy, cy), fz = fma2(hz, aiz, cz);
    const f2 ew = fma2(lo2(v[5]), w2, hi2(C4));
    PT_BOX_HALF(x)
    PT_BOX_HALF(y)
    sink(lo_x, miss_x, lo_y, miss_y, i);
  }
  // ---- class 3: other cubes, object space ----
  for (const int e = min(last, end[3]); i < e; i++, v += kFiltRows) {
    PT_FILT_TRANSFORM(v, r)
    const float4 K0 = v[6], K1 = v[7], K2 = v[8], K3 = v[9];
    const f2 hx = fma2(hi2(K1), w2, lo2(K0)), hy = fma2(lo2(K2), w2, hi2(K0)), hz = fma2(hi2(K2), w2, lo2(K1));
    const f2 ix = make_float2(mufu_rcp(rwx.x), mufu_rcp(rwx.y)), iy = make_float2(mufu_rcp(rwy.x), mufu_rcp(rwy.y)),
             iz = make_float2(mufu_rcp(rwz.x), mufu_rcp(rwz.y));
    const f2 cx = mul2(neg2(rox), ix), cy = mul2(neg2(roy), iy), cz = mul2(neg2(roz), iz);  // parameter of the slab centre
    // slab i spans [c_i - h_i|iv_i|, c_i + h_i|iv_i|]; a direction component of 0 gives infinities / NaN, which
    // fmaxf / fminf ignore: that slab then does not constrain (conservative)
    const f2 nx = fma2(hx, neg2(abs2(ix)), cx), ny = fma2(hy, neg2(abs2(iy)), cy), nz = fma2(hz, neg2(abs2(iz)), cz);
    const f2 fx = fma2(hx, abs2(ix), cx), fy = fma2(hy, abs2(iy), cy), fz = fma2(hz, abs2(iz), cz);
    const f2 ew = fma2(hi2(K3), w2, lo2(K3));
    PT_BOX_HALF(x)
    PT_BOX_HALF(y)
    sink(lo_x, miss_x, lo_y, miss_y, i);
  }
}
#undef PT_SPHERE_HALF
#undef PT_BOX_HALF
#undef PT_FILT_TRANSFORM
#ifndef PT_FILT_KEYED
#define PT_FILT_KEYED 1
#endif
struct TakeBest {
  ScanBest& b;
  __device__ __forceinline__ void operator()(float lo_a, bool miss_a, float lo_b, bool miss_b, int pair) {
    scan_take(b, miss_a ? INFINITY : lo_a, 2 * pair);
    scan_take(b, miss_b ? INFINITY : lo_b, 2 * pair + 1);
  }
};
// The same bookkeeping on KEYS: a candidate's index 2*pair + half replaces the low kKeyBits bits of its (non-negative)
// bound, which rounds the bound DOWN by at most 2^-16 of itself -- still a lower bound -- and lets minimum / maximum
// instructions carry the index along: 11 instructions per pair instead of 14.  Keys of non-negative floats order like the
// floats; +inf (no index bits) = nothing.  A bound that is itself +inf or NaN can only lose the candidate or lower lo2:
// +inf cannot be the bound of a hit the exact test accepts (finite distances only), NaN became 0 before.
constexpr int kKeyBits = 7;  // 2 * kMaxSmemPairs = 128 halves
constexpr uint32_t kKeyMask = (1u << kKeyBits) - 1u;
struct TakeBestKeyed {
  float k1 = INFINITY, k2 = INFINITY;
  static __device__ __forceinline__ float key(float lo, bool miss, uint32_t k) {
    const float v = __uint_as_float((__float_as_uint(fmaxf(lo, 0.0f)) & ~kKeyMask) | k);  // (NaN -> 0: "no information")
    return miss ? INFINITY : v;
  }
  __device__ __forceinline__ void operator()(float lo_a, bool miss_a, float lo_b, bool miss_b, int pair) {
    const float ka = key(lo_a, miss_a, 2u * pair), kb = key(lo_b, miss_b, 2u * pair + 1u);
    const float mn = fminf(ka, kb), mx = fmaxf(ka, kb);
    k2 = fminf(fminf(k2, fmaxf(k1, mn)), mx);
    k1 = fminf(k1, mn);
  }
  __device__ __forceinline__ void finish(ScanBest& b) const {
    b.k1 = k1 < INFINITY ? (int)(__float_as_uint(k1) & kKeyMask) : -1;
    b.lo1 = __uint_as_float(__float_as_uint(k1) & ~kKeyMask);
    b.lo2 = __uint_as_float(__float_as_uint(k2) & ~kKeyMask);
  }
};
// (`best` is overwritten: a scan starts from nothing)
__device__ __forceinline__ void filter_scan(const float4* v, int first, int last, const int end[kFiltClasses],
                                            const ScanRay& r, ScanBest& best) {
#if PT_FILT_KEYED
  TakeBestKeyed sink;
  filter_scan_to(v, first, last, end, r, sink);
  sink.finish(best);
#else
  TakeBest sink{best};
  filter_scan_to(v, first, last, end, r, sink);
#endif
}

// ---- the exact test of ONE geom: the reference's arithmetic, unfused, in its order (see pt_device.cuh) ----
// Returns false if the object-space test reports a miss; otherwise the world distance, the world point and the face
// code (cube: axis | (negative ? 4 : 0); sphere: 8).  The caller applies `dist > 0` and the closest-hit rule.
// M: Guarded (range guards inside sqrt / reciprocal, the rule) or DeferredGuard (fast paths unconditionally, the caller checks
// m.bad and does not use the result if it is set)
template <typename M>
__device__ __forceinline__ bool exact_hit(int type, float4 i0, float4 i1, float4 i2, float4 f0, float4 f1, float4 f2,
                                          f3 o, f3 d, float& dist, f3& P, int& ncode, M& m) {
  // intersections.h:85-86: object-space origin and re-normalised direction
  // (Scalar on purpose.  Doing both products side by side with the packed f32x2 instructions would halve these 42
  // operations, but ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad false -- also when
  // they are spelled fma(a, b, -0) and fma(a, 1, b) -- and the fused sums differ in the last bit: parity with the
  // reference's arithmetic was lost (tests/test_gpu_parity.py::test_closest_hit_primary_rays) for no gain.  DESIGN.md 3b.)
  const f3 ro = mulMV(i0, i1, i2, o.x, o.y, o.z, 1.0f);
  const f3 rd = normalize(mulMV(i0, i1, i2, d.x, d.y, d.z, 0.0f), m);
  float t;
  if (type == 0) {
    // sphereIntersectionTest, intersections.h:90-108
    const float vDot = dot(ro, rd);
    // the reference's host build evaluates float*float - (float - pow(.5f,2)) in binary64 (pow -> double)
    const float radicand = (float)((double)(vDot * vDot) - ((double)dot(ro, ro) - 0.25));
    if (radicand < 0) return false;
    const float sq = m.sqrt(radicand);
    const float first_term = -vDot;
    const float t1 = first_term + sq;
    const float t2 = first_term - sq;
    if (t1 < 0 && t2 < 0) return false;
    else if (t1 > 0 && t2 > 0) t = fminf(t1, t2);
    else t = fmaxf(t1, t2);
    ncode = 8;
  } else {
    // boxIntersectionTest (stub in the reference), DESIGN.md "box test": slabs on [-0.5,0.5]^3, IEEE minNum/maxNum
    const float ivx = m.rcp(rd.x), ivy = m.rcp(rd.y), ivz = m.rcp(rd.z);
    const float t1x = (-0.5f - ro.x) * ivx, t2x = (0.5f - ro.x) * ivx;
    const float t1y = (-0.5f - ro.y) * ivy, t2y = (0.5f - ro.y) * ivy;
    const float t1z = (-0.5f - ro.z) * ivz, t2z = (0.5f - ro.z) * ivz;
    const float lx = fminf(t1x, t2x), hx = fmaxf(t1x, t2x);
    const float ly = fminf(t1y, t2y), hy = fmaxf(t1y, t2y);
    const float lz = fminf(t1z, t2z), hz = fmaxf(t1z, t2z);
    const float tnear = fmaxf(fmaxf(lx, ly), lz), tfar = fminf(fminf(hx, hy), hz);
    if (tnear > tfar || tfar < 0) return false;
    const bool outside = tnear > 0;
    int axis;
    if (outside) { t = tnear; axis = lx == tnear ? 0 : (ly == tnear ? 1 : 2); }
    else { t = tfar; axis = hx == tfar ? 0 : (hy == tfar ? 1 : 2); }
    const float rda = axis == 0 ? rd.x : (axis == 1 ? rd.y : rd.z);
    const bool negative = outside ? (rda > 0) : !(rda > 0);
    ncode = axis | (negative ? 4 : 0);
  }
  // intersections.h:110,116: world point of the pulled-back object-space point, world distance
  const f3 po = point_on_ray(ro, rd, t, m);
  P = mulMV(f0, f1, f2, po.x, po.y, po.z, 1.0f);
  dist = length(o - P, m);
  return true;
}

__device__ __forceinline__ bool exact_hit(int type, float4 i0, float4 i1, float4 i2, float4 f0, float4 f1, float4 f2,
                                          f3 o, f3 d, float& dist, f3& P, int& ncode) {
  Guarded m;
  return exact_hit(type, i0, i1, i2, f0, f1, f2, o, d, dist, P, ncode, m);
}

// the exact scan: every geom through exact_hit, index order, strictly smaller positive distance wins
// (the specification of closest hit; also the fallback of closest_hit_filtered)
__device__ __noinline__ void closest_hit_exact(const GeomSoA g, int n_geoms, f3 o, f3 d, Hit& h) {
  for (int i = 0; i < n_geoms; i++) {
    const int type = __ldg(&g.meta[i]).x;
    if (type > 1) continue;  // MESH: no geometry (src/scene.cpp:57-66)
    float dist;
    f3 P;
    int ncode;
    if (!exact_hit(type, __ldg(g.inv0 + i), __ldg(g.inv1 + i), __ldg(g.inv2 + i), __ldg(g.fwd0 + i), __ldg(g.fwd1 + i),
                   __ldg(g.fwd2 + i), o, d, dist, P, ncode))
      continue;
    if (dist > 0 && dist < h.t) { h.t = dist; h.id = i; h.p = P; h.ncode = ncode; }
  }
}

// Resolve a finished scan: exact test of the best candidate, accepted if it is a hit closer than every other
// geom's lower bound; otherwise the exact scan.  Returns true if the fallback ran (statistics only).
// (A fallback that re-runs the filter pass and tests exactly only the geoms whose bound does not exceed the best exact
// distance so far was measured: +0.9 % on the Cornell box, where 0.43 % of the segments fall back, but -2 % on the sample
// scene -- the extra live state around the call costs the common path more than the shorter fallback saves.  DESIGN.md 3b.)
__device__ __forceinline__ bool resolve_scan(const ScanBest& best, const FiltSoA& f, const GeomSoA& g, int n_geoms,
                                             f3 o, f3 d, Hit& h) {
  if (best.k1 < 0) return false;  // every geom is a proven miss
  const int gi = __ldg(reinterpret_cast<const int*>(f.ids) + best.k1);
  const int type = best.k1 < 2 * f.end[1] ? 0 : 1;
  float dist;
  f3 P;
  int ncode;
  DeferredGuard m;  // (an argument outside the fast paths' range sends the ray to the exact scan, which guards every call)
  const bool hit = exact_hit(type, __ldg(g.inv0 + gi), __ldg(g.inv1 + gi), __ldg(g.inv2 + gi), __ldg(g.fwd0 + gi),
                             __ldg(g.fwd1 + gi), __ldg(g.fwd2 + gi), o, d, dist, P, ncode, m);
  if (!m.bad && hit && dist > 0 && dist < best.lo2) {
    h.t = dist; h.id = gi; h.p = P; h.ncode = ncode;
    return false;
  }
  closest_hit_exact(g, n_geoms, o, d, h);
  return true;
}

}  // namespace ptd
