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

// ---- filter geometry: spheres first, then cubes (MESH left out); 5 float4 per geom ----
//   a0,a1,a2 = rows x,y,z of inverseTransform
//   cube:   k0 = (hc.x, hc.y, hc.z, Ew_c)   k1 = (hw.x, hw.y, hw.z, Ew_w)     half extent_i = hc_i + hw_i * w
//   sphere: k0 = (R2c,  R2w,  R2r,  Ew_c)   k1 = (-,    -,    -,    Ew_w)     radius^2 = R2c + R2w * w + R2r * |ro|^2
//   world slack E_w = Ew_c + Ew_w * w
struct FiltSoA {  // in HBM
  const float4 *a0, *a1, *a2, *k0, *k1;
  const int* ids;  // filter index -> geom index
  int n_spheres, n_total;
  float r_scene;   // bound on |p| over all surface points of the scene
};
struct FiltSmem {
  float4 *a0, *a1, *a2, *k0, *k1;
};
__host__ __device__ inline size_t filt_smem_bytes(int cap) { return (size_t)cap * 5 * sizeof(float4); }
__device__ __forceinline__ FiltSmem carve_filt_smem(unsigned char* base, int cap) {
  FiltSmem s;
  float4* f = reinterpret_cast<float4*>(base);
  s.a0 = f; s.a1 = f + cap; s.a2 = f + 2 * cap; s.k0 = f + 3 * cap; s.k1 = f + 4 * cap;
  return s;
}
__device__ __forceinline__ FiltSmem filt_global_view(const FiltSoA& g, int first) {
  FiltSmem s;
  s.a0 = const_cast<float4*>(g.a0) + first; s.a1 = const_cast<float4*>(g.a1) + first;
  s.a2 = const_cast<float4*>(g.a2) + first; s.k0 = const_cast<float4*>(g.k0) + first;
  s.k1 = const_cast<float4*>(g.k1) + first;
  return s;
}
// cooperative copy of filter geoms [first, first+count) into shared memory; caller synchronises
__device__ __forceinline__ void stage_filt(const FiltSoA& g, int first, int count, const FiltSmem& s) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    s.a0[i] = g.a0[first + i]; s.a1[i] = g.a1[first + i]; s.a2[i] = g.a2[first + i];
    s.k0[i] = g.k0[first + i]; s.k1[i] = g.k1[first + i];
  }
}

// per-ray constants of the scan
struct ScanRay {
  f3 o, d;
  float w;   // 4*max|o_j| + R_scene
  float dl;  // |d| rounded down (parameter -> world distance)
};
__device__ __forceinline__ ScanRay make_scan_ray(f3 o, f3 d, float r_scene) {
  ScanRay r;
  r.o = o; r.d = d;
  const float omax = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
  r.w = __fmaf_rn(4.0f, omax, r_scene);
  const float d2 = __fmaf_rn(d.x, d.x, __fmaf_rn(d.y, d.y, d.z * d.z));
  r.dl = mufu_sqrt(d2) * 0.99999905f;  // 1 - 2^-20: below |d| whatever the approximation error
  return r;
}

// running result of the scan
struct ScanBest {
  float lo1, lo2;  // smallest and second-smallest lower bound
  int k1;          // filter index of the smallest, -1 = every geom so far is a proven miss
};
__device__ __forceinline__ void scan_init(ScanBest& b) { b.lo1 = INFINITY; b.lo2 = INFINITY; b.k1 = -1; }
__device__ __forceinline__ void scan_take(ScanBest& b, float lo, int k) {
  lo = fmaxf(lo, 0.0f);  // also turns a NaN bound into 0 (fmaxf ignores NaN): "no information"
  const bool better = lo < b.lo1;
  b.lo2 = fminf(b.lo2, better ? b.lo1 : lo);
  b.k1 = better ? k : b.k1;
  b.lo1 = fminf(b.lo1, lo);
}

// object-space origin and UN-normalised direction with fused multiply-adds: the parameter along (ro, rw) is the
// parameter along the world ray
#define PT_FILT_TRANSFORM(S, I, R)                                                                              \
  const float4 A0 = (S).a0[I], A1 = (S).a1[I], A2 = (S).a2[I];                                                  \
  const float rox = __fmaf_rn(A0.x, (R).o.x, __fmaf_rn(A0.y, (R).o.y, __fmaf_rn(A0.z, (R).o.z, A0.w)));         \
  const float roy = __fmaf_rn(A1.x, (R).o.x, __fmaf_rn(A1.y, (R).o.y, __fmaf_rn(A1.z, (R).o.z, A1.w)));         \
  const float roz = __fmaf_rn(A2.x, (R).o.x, __fmaf_rn(A2.y, (R).o.y, __fmaf_rn(A2.z, (R).o.z, A2.w)));         \
  const float rwx = __fmaf_rn(A0.x, (R).d.x, __fmaf_rn(A0.y, (R).d.y, A0.z * (R).d.z));                         \
  const float rwy = __fmaf_rn(A1.x, (R).d.x, __fmaf_rn(A1.y, (R).d.y, A1.z * (R).d.z));                         \
  const float rwz = __fmaf_rn(A2.x, (R).d.x, __fmaf_rn(A2.y, (R).d.y, A2.z * (R).d.z));

// Scan filter geoms [0, count) of `s` (filter indices base..base+count); the first n_sph of them are spheres.
// Every "miss" needs a comparison to come out TRUE, so a NaN anywhere keeps the geom as a candidate.
__device__ __forceinline__ void filter_scan(const FiltSmem& s, int base, int n_sph, int count, const ScanRay& r,
                                            ScanBest& best) {
  int i = 0;
  for (; i < n_sph; i++) {
    PT_FILT_TRANSFORM(s, i, r)
    const float4 K0 = s.k0[i];
    const float a = __fmaf_rn(rwx, rwx, __fmaf_rn(rwy, rwy, rwz * rwz));
    const float b = __fmaf_rn(rox, rwx, __fmaf_rn(roy, rwy, roz * rwz));
    const float ro2 = __fmaf_rn(rox, rox, __fmaf_rn(roy, roy, roz * roz));
    const float R2 = __fmaf_rn(K0.z, ro2, __fmaf_rn(K0.y, r.w, K0.x));
    const float c = ro2 - R2;
    const float disc = __fmaf_rn(b, b, -(a * c));
    if (disc < 0.0f) continue;  // the line misses the inflated sphere
    const float sd = mufu_sqrt(disc), ia = mufu_rcp(a);
    if ((sd - b) * ia < 0.0f) continue;  // the inflated sphere lies behind the origin
    const float ew = __fmaf_rn(s.k1[i].w, r.w, K0.w);
    scan_take(best, __fmaf_rn((-b - sd) * ia, r.dl, -ew), base + i);
  }
  for (; i < count; i++) {
    PT_FILT_TRANSFORM(s, i, r)
    const float4 K0 = s.k0[i], K1 = s.k1[i];
    const float hx = __fmaf_rn(K1.x, r.w, K0.x), hy = __fmaf_rn(K1.y, r.w, K0.y), hz = __fmaf_rn(K1.z, r.w, K0.z);
    const float ix = mufu_rcp(rwx), iy = mufu_rcp(rwy), iz = mufu_rcp(rwz);
    const float cx = -rox * ix, cy = -roy * iy, cz = -roz * iz;  // parameter of the slab centre
    // slab i spans [c_i - h_i|iv_i|, c_i + h_i|iv_i|]; a direction component of 0 gives infinities / NaN, which
    // fmaxf / fminf ignore: that slab then does not constrain (conservative)
    const float tnear = fmaxf(fmaxf(__fmaf_rn(hx, -fabsf(ix), cx), __fmaf_rn(hy, -fabsf(iy), cy)), __fmaf_rn(hz, -fabsf(iz), cz));
    const float tfar = fminf(fminf(__fmaf_rn(hx, fabsf(ix), cx), __fmaf_rn(hy, fabsf(iy), cy)), __fmaf_rn(hz, fabsf(iz), cz));
    if (tnear > tfar || tfar < 0.0f) continue;  // misses the inflated box, or the box lies behind the origin
    const float ew = __fmaf_rn(K1.w, r.w, K0.w);
    scan_take(best, __fmaf_rn(tnear, r.dl, -ew), base + i);
  }
}
#undef PT_FILT_TRANSFORM

// ---- the exact test of ONE geom: the reference's arithmetic, unfused, in its order (see pt_device.cuh) ----
// Returns false if the object-space test reports a miss; otherwise the world distance, the world point and the face
// code (cube: axis | (negative ? 4 : 0); sphere: 8).  The caller applies `dist > 0` and the closest-hit rule.
__device__ __forceinline__ bool exact_hit(int type, float4 i0, float4 i1, float4 i2, float4 f0, float4 f1, float4 f2,
                                          f3 o, f3 d, float& dist, f3& P, int& ncode) {
  // intersections.h:85-86: object-space origin and re-normalised direction
  const f3 ro = mulMV(i0, i1, i2, o.x, o.y, o.z, 1.0f);
  const f3 rd = normalize(mulMV(i0, i1, i2, d.x, d.y, d.z, 0.0f));
  float t;
  if (type == 0) {
    // sphereIntersectionTest, intersections.h:90-108
    const float vDot = dot(ro, rd);
    // the reference's host build evaluates float*float - (float - pow(.5f,2)) in binary64 (pow -> double)
    const float radicand = (float)((double)(vDot * vDot) - ((double)dot(ro, ro) - 0.25));
    if (radicand < 0) return false;
    const float sq = sqrtf(radicand);
    const float first_term = -vDot;
    const float t1 = first_term + sq;
    const float t2 = first_term - sq;
    if (t1 < 0 && t2 < 0) return false;
    else if (t1 > 0 && t2 > 0) t = fminf(t1, t2);
    else t = fmaxf(t1, t2);
    ncode = 8;
  } else {
    // boxIntersectionTest (stub in the reference), DESIGN.md "box test": slabs on [-0.5,0.5]^3, IEEE minNum/maxNum
    const float ivx = 1.0f / rd.x, ivy = 1.0f / rd.y, ivz = 1.0f / rd.z;
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
  const f3 po = point_on_ray(ro, rd, t);
  P = mulMV(f0, f1, f2, po.x, po.y, po.z, 1.0f);
  dist = length(o - P);
  return true;
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
__device__ __forceinline__ bool resolve_scan(const ScanBest& best, const FiltSoA& f, const GeomSoA& g, int n_geoms,
                                             f3 o, f3 d, Hit& h) {
  if (best.k1 < 0) return false;  // every geom is a proven miss
  const int gi = __ldg(f.ids + best.k1);
  const int type = best.k1 < f.n_spheres ? 0 : 1;
  float dist;
  f3 P;
  int ncode;
  const bool hit = exact_hit(type, __ldg(g.inv0 + gi), __ldg(g.inv1 + gi), __ldg(g.inv2 + gi), __ldg(g.fwd0 + gi),
                             __ldg(g.fwd1 + gi), __ldg(g.fwd2 + gi), o, d, dist, P, ncode);
  if (hit && dist > 0 && dist < best.lo2) {
    h.t = dist; h.id = gi; h.p = P; h.ncode = ncode;
    return false;
  }
  closest_hit_exact(g, n_geoms, o, d, h);
  return true;
}

}  // namespace ptd
