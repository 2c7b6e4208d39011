/* TEST INFRASTRUCTURE ONLY -- see pt_oracle.h for the contract.
 *
 * CPU restatement of the reference's path-tracing hot path.  Each function cites the
 * reference file:line it follows (paths relative to the reference repo root).  Functions
 * marked (R) restate code the reference implements and are pinned bit-for-bit against it;
 * functions marked (S) are the frozen specification of TODO stubs.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fno-fast-math -fopenmp (oracle/Makefile).
 */
#include "pt_oracle.h"

#include <math.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } v3;

/* ---- GLM 0.9.5.4 vector helpers, restated (external/include/glm/detail/func_geometric.inl) ---- */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* compute_dot<tvec3>: tmp = x*y; tmp.x + tmp.y + tmp.z   (func_geometric.inl:66-72) */
static inline float vdot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* cross (func_geometric.inl:216-228) */
static inline v3 vcross(v3 x, v3 y) {
  return V(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
/* length: sqrt(x*x+y*y+z*z) (func_geometric.inl:108-114) */
static inline float vlength(v3 v) { return sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z); }
/* normalize: x * inversesqrt(dot); inversesqrt(float) = 1.0f / sqrt(x)
 * (func_geometric.inl:256-265, func_exponential.inl:226-229) */
static inline v3 vnormalize(v3 v) {
  float sqr = (v.x * v.x + v.y * v.y) + v.z * v.z;
  float inv = 1.0f / sqrtf(sqr);
  return V(v.x * inv, v.y * inv, v.z * inv);
}
static inline v3 ld3(const float* p) { return V(p[0], p[1], p[2]); }
static inline void st3(float* p, v3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

/* (R) src/intersections.h:26-34 */
unsigned int or_hash(unsigned int a) {
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

/* (R) src/intersections.h:37-43: binary64 comparison against EPSILON = 1e-9 (src/utilities.h:24); unused by the
 * reference's implemented code */
int or_epsilonCheck(float a, float b) { return fabs(fabs((double)a) - fabs((double)b)) < 0.000000001 ? 1 : 0; }
/* (R) src/intersections.h:62-70: 1.0 / d in binary64, rounded to binary32 on the way into the vec3; signs as 0 / 1.
 * Slab-test helpers the reference provides and never calls; the box test of this repo uses the binary32 reciprocal. */
void or_ray_helpers(const float d[3], float inv[3], float sign[3]) {
  for (int k = 0; k < 3; k++) {
    inv[k] = (float)(1.0 / (double)d[k]);
    sign[k] = (float)(int)(inv[k] < 0);
  }
}

/* (R) src/intersections.h:53-59: rows x,y,z of the row-stored cudaMat4 times v, left to right. */
static inline v3 mulMV(const float* m, float vx, float vy, float vz, float vw) {
  v3 r;
  r.x = (m[0] * vx) + (m[1] * vy) + (m[2] * vz) + (m[3] * vw);
  r.y = (m[4] * vx) + (m[5] * vy) + (m[6] * vz) + (m[7] * vw);
  r.z = (m[8] * vx) + (m[9] * vy) + (m[10] * vz) + (m[11] * vw);
  return r;
}
void or_multiplyMV(const float m[16], const float v[4], float out[3]) { st3(out, mulMV(m, v[0], v[1], v[2], v[3])); }

/* (R) src/intersections.h:46-48: origin + (t - .0001f) * normalize(direction) */
static inline v3 pointOnRay(v3 o, v3 d, float t) { return vadd(o, vscale(vnormalize(d), (float)(t - .0001f))); }
void or_getPointOnRay(const float o[3], const float d[3], float t, float out[3]) {
  st3(out, pointOnRay(ld3(o), ld3(d), t));
}

/* (R) src/intersections.h:81-117.  Unit sphere radius .5 in object space, object-space solve with a
 * re-normalised direction, 1e-4 pull-back, world-space distance returned, -1 on a miss. */
float or_sphereIntersectionTest(const or_static_geom* g, const float o[3], const float d[3], float p[3], float n[3]) {
  v3 ro = mulMV(g->inverseTransform, o[0], o[1], o[2], 1.0f);
  v3 rd = vnormalize(mulMV(g->inverseTransform, d[0], d[1], d[2], 0.0f));
  float vDotDirection = vdot(ro, rd);
  /* f64 step: pow(float,int) is double on the reference's host build, so the subtraction chain
   * float*float - (float - double) is evaluated in binary64 and rounded once (intersections.h:91). */
  float radicand = (float)((double)(vDotDirection * vDotDirection) - ((double)vdot(ro, ro) - 0.25));
  if (radicand < 0) return -1;
  float squareRoot = sqrtf(radicand);
  float firstTerm = -vDotDirection;
  float t1 = firstTerm + squareRoot;
  float t2 = firstTerm - squareRoot;
  float t;
  if (t1 < 0 && t2 < 0) {
    return -1;
  } else if (t1 > 0 && t2 > 0) {
    t = fminf(t1, t2);
  } else {
    t = fmaxf(t1, t2);
  }
  v3 po = pointOnRay(ro, rd, t);
  v3 realP = mulMV(g->transform, po.x, po.y, po.z, 1.0f);
  v3 realOrigin = mulMV(g->transform, 0.0f, 0.0f, 0.0f, 1.0f);
  st3(p, realP);
  st3(n, vnormalize(vsub(realP, realOrigin)));
  return vlength(vsub(ld3(o), realP));
}

/* (R) src/intersections.h:120-129 */
void or_getRadiuses(const or_static_geom* g, float out[3]) {
  v3 origin = mulMV(g->transform, 0, 0, 0, 1);
  v3 xmax = mulMV(g->transform, .5f, 0, 0, 1);
  v3 ymax = mulMV(g->transform, 0, .5f, 0, 1);
  v3 zmax = mulMV(g->transform, 0, 0, .5f, 1);
  /* glm::distance(p0,p1) = length(p1 - p0) */
  out[0] = vlength(vsub(xmax, origin));
  out[1] = vlength(vsub(ymax, origin));
  out[2] = vlength(vsub(zmax, origin));
}

/* src/utilities.h:20-26 (double literals) */
#define OR_TWO_PI 6.2831853071795864769252867665590057683943
#define OR_SQRT_OF_ONE_THIRD 0.5773502691896257645091487805019574556476
#define OR_RAY_BIAS_AMOUNT 0.0002f

/* shared body of src/interactions.h:62-87; cs/sn are cos/sin of 2*pi*xi2 */
static inline v3 hemisphere_body(v3 normal, float xi1, float cs, float sn) {
  float up = sqrtf(xi1);
  float over = sqrtf(1 - up * up);
  v3 directionNotNormal;
  /* float compared against a double literal (interactions.h:72-78) */
  if ((double)fabsf(normal.x) < OR_SQRT_OF_ONE_THIRD) {
    directionNotNormal = V(1, 0, 0);
  } else if ((double)fabsf(normal.y) < OR_SQRT_OF_ONE_THIRD) {
    directionNotNormal = V(0, 1, 0);
  } else {
    directionNotNormal = V(0, 0, 1);
  }
  v3 p1 = vnormalize(vcross(normal, directionNotNormal));
  v3 p2 = vnormalize(vcross(normal, p1));
  return vadd(vadd(vscale(normal, up), vscale(p1, cs * over)), vscale(p2, sn * over));
}

/* (R) src/interactions.h:62-87 verbatim in behaviour: around = xi2*TWO_PI is an f64 step, cos/sin are libm. */
void or_hemisphere_ref(const float n[3], float xi1, float xi2, float out[3]) {
  float around = (float)((double)xi2 * OR_TWO_PI);
  st3(out, hemisphere_body(ld3(n), xi1, cosf(around), sinf(around)));
}

/* (S) sin/cos of 2*pi*u for u in [0,1): exact quadrant reduction in turns, then the classic single-precision
 * minimax polynomials on [-pi/4, pi/4], Horner, unfused.  Identical op sequence in the CUDA kernels, so the
 * BSDF sampler is bit-reproducible across host and device (libm and CUDA sinf/cosf are not). */
void or_sincos_2pi(float u, float* s, float* c) {
  int q = (int)(u * 4.0f + 0.5f);
  float r = u - 0.25f * (float)q; /* exact */
  float th = r * 6.2831855f;      /* 0x40C90FDB */
  float z = th * th;
  float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * th + th;
  float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
  switch (q & 3) {
    case 0: *s = sp; *c = cp; break;
    case 1: *s = cp; *c = -sp; break;
    case 2: *s = -sp; *c = -cp; break;
    default: *s = -cp; *c = sp; break;
  }
}

/* (S) the sampler the path loop uses: src/interactions.h:62-87 with or_sincos_2pi(xi2). */
void or_hemisphere(const float n[3], float xi1, float xi2, float out[3]) {
  float sn, cs;
  or_sincos_2pi(xi2, &sn, &cs);
  st3(out, hemisphere_body(ld3(n), xi1, cs, sn));
}

/* (S) src/intersections.h:74-77 is a stub returning -1.  Specified to "work in the same way as
 * sphereIntersectionTest" (README.md:121-123): same transform / re-normalise / pull-back / world-distance
 * conventions (intersections.h:85-86,110-116); slab test on [-0.5,0.5]^3; nearest positive root, far root
 * when the origin is inside; outward object-space face normal mapped by the forward transform. */
float or_boxIntersectionTest(const or_static_geom* g, const float o[3], const float d[3], float p[3], float n[3]) {
  v3 ro = mulMV(g->inverseTransform, o[0], o[1], o[2], 1.0f);
  v3 rd = vnormalize(mulMV(g->inverseTransform, d[0], d[1], d[2], 0.0f));
  float roa[3] = {ro.x, ro.y, ro.z};
  float rda[3] = {rd.x, rd.y, rd.z};
  float lo[3], hi[3];
  for (int a = 0; a < 3; a++) {
    float inv = 1.0f / rda[a];
    float t1 = (-0.5f - roa[a]) * inv;
    float t2 = (0.5f - roa[a]) * inv;
    lo[a] = fminf(t1, t2); /* IEEE minNum / maxNum: a NaN operand (0 * inf) is ignored */
    hi[a] = fmaxf(t1, t2);
  }
  float tnear = fmaxf(fmaxf(lo[0], lo[1]), lo[2]);
  float tfar = fminf(fminf(hi[0], hi[1]), hi[2]);
  if (tnear > tfar || tfar < 0) return -1;
  float t;
  int axis;
  float sign;
  if (tnear > 0) { /* entering: the first axis that attains the maximum */
    t = tnear; axis = lo[0] == tnear ? 0 : (lo[1] == tnear ? 1 : 2); sign = rda[axis] > 0 ? -1.0f : 1.0f;
  } else {         /* origin inside: leave through the first axis that attains the minimum */
    t = tfar; axis = hi[0] == tfar ? 0 : (hi[1] == tfar ? 1 : 2); sign = rda[axis] > 0 ? 1.0f : -1.0f;
  }
  v3 no = V(axis == 0 ? sign : 0.0f, axis == 1 ? sign : 0.0f, axis == 2 ? sign : 0.0f);
  v3 po = pointOnRay(ro, rd, t);
  v3 realP = mulMV(g->transform, po.x, po.y, po.z, 1.0f);
  st3(p, realP);
  st3(n, vnormalize(mulMV(g->transform, no.x, no.y, no.z, 0.0f)));
  return vlength(vsub(ld3(o), realP));
}

/* (S) Philox-4x32-10 (Salmon et al., SC'11; Random123).  Counter-based: no state, identical on host and device. */
void or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* (S) uniform in [0,1): top 24 bits */
float or_u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }

/* RNG addressing: counter = (pixel, sample, block, 0), key = seed; block 0 = raygen, 1+depth = bounce. */
static inline void rng4(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t block, float u[4]) {
  uint32_t ctr[4] = {pixel, sample, block, 0u};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  or_philox4x32_10(ctr, key, r);
  for (int i = 0; i < 4; i++) u[i] = or_u01(r[i]);
}

/* (S) src/interactions.h:47-50 stub: mirror of `incident` about `normal` */
static inline v3 reflect3(v3 n, v3 i) { return vsub(i, vscale(n, 2.0f * vdot(i, n))); }
void or_reflect(const float n[3], const float i[3], float out[3]) { st3(out, reflect3(ld3(n), ld3(i))); }

/* (S) src/interactions.h:42-44 stub: Snell refraction; returns 1 on total internal reflection */
static inline int refract3(v3 n, v3 i, float ior_i, float ior_t, v3* out) {
  float eta = ior_i / ior_t;
  float c = -vdot(n, i);
  float k = 1.0f - (eta * eta) * (1.0f - c * c);
  if (k < 0) { *out = V(0, 0, 0); return 1; }
  *out = vadd(vscale(i, eta), vscale(n, eta * c - sqrtf(k)));
  return 0;
}
int or_refract(const float n[3], const float i[3], float ior_i, float ior_t, float out[3]) {
  v3 t; int tir = refract3(ld3(n), ld3(i), ior_i, ior_t, &t); st3(out, t); return tir;
}

/* (S) src/interactions.h:53-59 stub + struct Fresnel (:11-14): unpolarised Fresnel equations */
static inline void fresnel3(v3 n, v3 i, float ior_i, float ior_t, v3 trans, int tir, float* R, float* T) {
  if (tir) { *R = 1.0f; *T = 0.0f; return; }
  float ci = -vdot(n, i);
  float ct = -vdot(n, trans);
  float rpar = (ior_t * ci - ior_i * ct) / (ior_t * ci + ior_i * ct);
  float rperp = (ior_i * ci - ior_t * ct) / (ior_i * ci + ior_t * ct);
  float r = 0.5f * (rpar * rpar + rperp * rperp);
  *R = r; *T = 1.0f - r;
}
void or_fresnel(const float n[3], const float i[3], float ior_i, float ior_t, const float refl[3],
                const float trans[3], int tir, float* R, float* T) {
  (void)refl;
  fresnel3(ld3(n), ld3(i), ior_i, ior_t, ld3(trans), tir, R, T);
}

/* (S) src/raytraceKernel.cu:40-45 stub.  fov are half-angles in degrees (scene.cpp:203-207); image plane at
 * distance 1; buffer x grows toward -right and y toward -up because the save path mirrors x and writes rows
 * top-down (main.cpp:120-125, image.cpp:51-65).  2 RNG dims jitter the pixel, 2 sample a thin lens
 * (aperture 0 => pinhole, bit-identical to having no lens). */
void or_raygen(const or_camera_data* cam, const or_lens* lens, uint64_t seed, uint32_t pixel, uint32_t sample,
               float o[3], float d[3]) {
  int W = (int)cam->resolution[0];
  float fw = cam->resolution[0], fh = cam->resolution[1];
  v3 eye = ld3(cam->position);
  v3 w = vnormalize(ld3(cam->view));
  v3 right = vnormalize(vcross(w, ld3(cam->up)));
  v3 vup = vcross(right, w);
  float tx = tanf(cam->fov[0] * 0.017453292f);
  float ty = tanf(cam->fov[1] * 0.017453292f);
  v3 Hh = vscale(right, tx);
  v3 Vv = vscale(vup, ty);
  float u[4];
  rng4(seed, pixel, sample, 0u, u);
  float x = (float)(pixel % (uint32_t)W), y = (float)(pixel / (uint32_t)W);
  float sx = 1.0f - 2.0f * ((x + u[0]) / fw);
  float sy = 1.0f - 2.0f * ((y + u[1]) / fh);
  v3 dir = vnormalize(vadd(vadd(w, vscale(Hh, sx)), vscale(Vv, sy)));
  v3 org = eye;
  if (lens && lens->aperture > 0.0f) {
    float ft = lens->focal_distance / vdot(dir, w);
    v3 pf = vadd(eye, vscale(dir, ft));
    float r = lens->aperture * sqrtf(u[2]);
    float sn, cs;
    or_sincos_2pi(u[3], &sn, &cs);
    org = vadd(vadd(eye, vscale(right, r * cs)), vscale(vup, r * sn));
    dir = vnormalize(vsub(pf, org));
  }
  st3(o, org);
  st3(d, dir);
}

/* ---- surface-point / direction sampling and absorption (a12, a13, a15, a20) ---- */

/* (R) thrust::default_random_engine = minstd_rand = LCG(48271, 0, 2^31-1) as the reference seeds and draws it
 * (src/intersections.h:135-137, src/raytraceKernel.cu:32-33; thrust/random/detail/linear_congruential_engine.inl:
 * seed s -> s mod m, 0 -> 1) and thrust::uniform_real_distribution<float>(a,b)
 * (detail/uniform_real_distribution.inl): (float)(x - min) / (1.0f + (float)(max - min)) * (b - a) + a. */
typedef struct { uint32_t x; } minstd;
static inline void minstd_seed(minstd* r, uint32_t s) {
  uint32_t v = s % 2147483647u;
  r->x = v == 0 ? 1u : v;
}
static inline float minstd_uniform(minstd* r, float a, float b) {
  r->x = (uint32_t)(((uint64_t)r->x * 48271u) % 2147483647u);
  float res = (float)(r->x - 1u);
  res /= (1.0f + (float)2147483645u);
  return (res * (b - a)) + a;
}

/* (R) generateRandomNumberFromThread, src/raytraceKernel.cu:29-36: what the reference's raytraceRay stub writes into
 * each pixel (:93-104).  index = x + y * resolution.x (binary32), seed = hash((unsigned)(index * time)), three draws.
 * reversed = 1: the draw order of the reference's HOST build (g++ evaluates the constructor's arguments right to
 * left), pinned against it; reversed = 0: left to right (its device build, checked on the GPU box). */
void or_generateRandomNumberFromThread(int W, int H, float time, int x, int y, int reversed, float out[3]) {
  (void)H;
  int index = (int)((float)x + ((float)y * (float)W));
  minstd rng;
  minstd_seed(&rng, or_hash((unsigned int)((float)index * time)));
  float a = minstd_uniform(&rng, 0.0f, 1.0f), b = minstd_uniform(&rng, 0.0f, 1.0f), c = minstd_uniform(&rng, 0.0f, 1.0f);
  if (reversed) { out[0] = c; out[1] = b; out[2] = a; }
  else { out[0] = a; out[1] = b; out[2] = c; }
}
void or_noise_image(int W, int H, float time, int reversed, float* out) {
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) or_generateRandomNumberFromThread(W, H, time, x, y, reversed, out + 3 * ((size_t)y * W + x));
}

/* body of src/intersections.h:140-172 given the three draws: face by area-weighted roulette, then the two
 * in-face coordinates (a = first vec3 argument that is random, b = second) */
/* the five face thresholds of :149-169 and the total area; they depend on the geom only */
static inline float cube_thresholds(const or_static_geom* g, float th[5]) {
  float radii[3];
  or_getRadiuses(g, radii);
  float side1 = radii[0] * radii[1] * 4.0f;
  float side2 = radii[2] * radii[1] * 4.0f;
  float side3 = radii[0] * radii[2] * 4.0f;
  float totalarea = 2.0f * (side1 + side2 + side3);
  th[0] = (side1 / totalarea);
  th[1] = ((side1 * 2) / totalarea);
  th[2] = (((side1 * 2) + (side2)) / totalarea);
  th[3] = (((side1 * 2) + (side2 * 2)) / totalarea);
  th[4] = (((side1 * 2) + (side2 * 2) + (side3)) / totalarea);
  return totalarea;
}
static inline v3 cube_point_th(const or_static_geom* g, const float th[5], float roulette, float a, float b) {
  v3 point;
  if (roulette < th[0]) point = V(a, b, .5f);
  else if (roulette < th[1]) point = V(a, b, -.5f);
  else if (roulette < th[2]) point = V(.5f, a, b);
  else if (roulette < th[3]) point = V(-.5f, a, b);
  else if (roulette < th[4]) point = V(a, .5f, b);
  else point = V(a, -.5f, b);
  return mulMV(g->transform, point.x, point.y, point.z, 1.0f);
}
static inline v3 cube_point_body(const or_static_geom* g, float roulette, float a, float b) {
  float th[5];
  cube_thresholds(g, th);
  return cube_point_th(g, th, roulette, a, b);
}

/* (R) getRandomPointOnCube, src/intersections.h:133-175, as the reference's HOST build behaves: the two u02(rng)
 * arguments of each glm::vec3(...) are evaluated right to left by g++, so the SECOND random
 * coordinate is drawn first.  randomSeed is converted float -> unsigned (defined for 0 <= seed < 2^32). */
void or_getRandomPointOnCube(const or_static_geom* g, float randomSeed, float out[3]) {
  minstd rng;
  minstd_seed(&rng, or_hash((unsigned int)randomSeed));
  float roulette = minstd_uniform(&rng, 0.0f, 1.0f);
  float b = minstd_uniform(&rng, -0.5f, 0.5f);
  float a = minstd_uniform(&rng, -0.5f, 0.5f);
  st3(out, cube_point_body(g, roulette, a, b));
}

/* (S) the same sampler driven by three given uniforms in [0,1) (the path loop feeds it Philox numbers) */
void or_cube_point_u(const or_static_geom* g, float u0, float u1, float u2, float out[3]) {
  st3(out, cube_point_body(g, u0, u1 - 0.5f, u2 - 0.5f));
}

/* (S) uniform direction on the unit sphere from two uniforms: z = 1 - 2*xi1, azimuth 2*pi*xi2
 * (getRandomDirectionInSphere, stub at src/interactions.h:93-95) */
static inline v3 sphere_dir(float xi1, float xi2) {
  float z = 1.0f - 2.0f * xi1;
  float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
  float sn, cs;
  or_sincos_2pi(xi2, &sn, &cs);
  return V(r * cs, r * sn, z);
}
void or_getRandomDirectionInSphere(float xi1, float xi2, float out[3]) { st3(out, sphere_dir(xi1, xi2)); }

/* (S) uniform point on the unit sphere of radius .5 in object space, mapped to world space like the cube sampler */
void or_sphere_point_u(const or_static_geom* g, float u0, float u1, float out[3]) {
  v3 d = sphere_dir(u0, u1);
  st3(out, mulMV(g->transform, 0.5f * d.x, 0.5f * d.y, 0.5f * d.z, 1.0f));
}
/* (S) getRandomPointOnSphere (stub at src/intersections.h:179-182): same generator construction as the cube sampler,
 * two draws in statement order */
void or_getRandomPointOnSphere(const or_static_geom* g, float randomSeed, float out[3]) {
  minstd rng;
  minstd_seed(&rng, or_hash((unsigned int)randomSeed));
  float u0 = minstd_uniform(&rng, 0.0f, 1.0f);
  float u1 = minstd_uniform(&rng, 0.0f, 1.0f);
  or_sphere_point_u(g, u0, u1, out);
}

/* (S) exp(x) in binary32 with +,-,* only (Cody-Waite reduction by ln2 = 0.693359375 - 2.12194440e-4, degree-5
 * polynomial of the classic single-precision expf, scale by 2^k through the exponent field): the same operation
 * sequence in the CUDA kernels, so absorption is bit-reproducible (libm and CUDA expf are not).  x <= -87 (and NaN)
 * gives 0, x is clamped to 88 from above. */
float or_exp(float x) {
  if (!(x > -87.0f)) return 0.0f;
  if (x > 88.0f) x = 88.0f;
  float kf = (float)(int)(x * 1.44269504f + (x < 0 ? -0.5f : 0.5f));
  float r = (x - kf * 0.693359375f) - kf * -2.12194440e-4f;
  float p = 1.9875691500e-4f;
  p = p * r + 1.3981999507e-3f;
  p = p * r + 8.3334519073e-3f;
  p = p * r + 4.1665795894e-2f;
  p = p * r + 1.6666665459e-1f;
  p = p * r + 5.0000001201e-1f;
  float y = (p * (r * r) + r) + 1.0f;
  union { uint32_t u; float f; } sc;
  sc.u = (uint32_t)((int)kf + 127) << 23;
  return y * sc.f;
}

/* (S) calculateTransmission (stub at src/interactions.h:31-33): Beer-Lambert, exp(-sigma_a * distance) per channel */
void or_calculateTransmission(const float absorption[3], float distance, float out[3]) {
  out[0] = or_exp(-(absorption[0] * distance));
  out[1] = or_exp(-(absorption[1] * distance));
  out[2] = or_exp(-(absorption[2] * distance));
}

/* batch forms for the tests: n seeds (or n uniform triples) against ONE geom; type 0 = sphere, else cube */
void or_random_points_batch(const or_static_geom* g, int n, const float* seed, float* out) {
  for (int i = 0; i < n; i++) {
    if (g->type == 0) or_getRandomPointOnSphere(g, seed[i], out + 3 * i);
    else or_getRandomPointOnCube(g, seed[i], out + 3 * i);
  }
}
void or_points_u_batch(const or_static_geom* g, int n, const float* u, float* out) {
  for (int i = 0; i < n; i++) {
    if (g->type == 0) or_sphere_point_u(g, u[3 * i], u[3 * i + 1], out + 3 * i);
    else or_cube_point_u(g, u[3 * i], u[3 * i + 1], u[3 * i + 2], out + 3 * i);
  }
}
void or_sphere_dirs_batch(int n, const float* xi1, const float* xi2, float* out) {
  for (int i = 0; i < n; i++) or_getRandomDirectionInSphere(xi1[i], xi2[i], out + 3 * i);
}
void or_transmission_batch(int n, const float* absorption, const float* distance, float* out) {
  for (int i = 0; i < n; i++) or_calculateTransmission(absorption + 3 * i, distance[i], out + 3 * i);
}

/* (S) closest hit: scan in index order, keep the strictly smaller positive world distance. MESH has no
 * geometry (scene.cpp:57-66) and is never hit. */
int or_closest_hit(const or_static_geom* geoms, int n_geoms, const float o[3], const float d[3], float* t,
                   float p[3], float n[3]) {
  float best = INFINITY;
  int id = -1;
  for (int i = 0; i < n_geoms; i++) {
    float pp[3], nn[3], tt;
    if (geoms[i].type == 0) tt = or_sphereIntersectionTest(&geoms[i], o, d, pp, nn);
    else if (geoms[i].type == 1) tt = or_boxIntersectionTest(&geoms[i], o, d, pp, nn);
    else continue;
    if (tt > 0 && tt < best) {
      best = tt; id = i;
      memcpy(p, pp, sizeof(pp)); memcpy(n, nn, sizeof(nn));
    }
  }
  *t = id >= 0 ? best : -1.0f;
  return id;
}

/* (S) src/interactions.h:99-104 stub (calculateBSDF) + README.md:63-82 feature list.  See SURVEY.md appendix E. */
int or_shade(const or_scene* sc, int geom_id, float t, const float p[3], const float n[3], uint64_t seed, uint32_t pixel,
             uint32_t sample, uint32_t depth, float o[3], float d[3], float thr[3], float L[3]) {
  const or_static_geom* g = &sc->geoms[geom_id];
  const or_material* m = &sc->materials[g->materialid];
  v3 T = ld3(thr), D = ld3(d), P = ld3(p), N = ld3(n);
  if (m->emittance > 0) {
    v3 e = vscale(vmul(T, ld3(m->color)), m->emittance);
    st3(L, e);
    return 3;
  }
  int entering = vdot(D, N) < 0;
  v3 ns = entering ? N : vneg(N); /* the reference's normals always point outward */
  float u[4];
  rng4(seed, pixel, sample, 1u + depth, u);
  int kind;
  v3 nd, no;
  if (m->hasRefractive > 0) {
    float ior = m->indexOfRefraction;
    /* the segment that ends here ran INSIDE the geom if it arrives from within: Beer-Lambert absorption over its
     * world length t with the material's ABSCOEFF (scene.cpp:250-252; calculateTransmission) */
    const float* ab = m->absorptionCoefficient;
    if (!entering && (ab[0] > 0 || ab[1] > 0 || ab[2] > 0)) {
      float tr3[3];
      or_calculateTransmission(ab, t, tr3);
      T = vmul(T, ld3(tr3));
    }
    float ei = entering ? 1.0f : ior, et = entering ? ior : 1.0f;
    v3 refl = reflect3(ns, D);
    v3 tr;
    int tir = refract3(ns, D, ei, et, &tr);
    float R, Tc;
    fresnel3(ns, D, ei, et, tr, tir, &R, &Tc);
    if (tir || u[2] < R) {
      kind = 1; nd = refl; no = vadd(P, vscale(ns, OR_RAY_BIAS_AMOUNT));
      T = vmul(T, ld3(m->specularColor));
    } else {
      /* the hit point sits 1e-4 object-space units before the surface (intersections.h:47,110); step across
       * by that pull-back measured in world space plus RAY_BIAS_AMOUNT (utilities.h:26) */
      v3 rdraw = mulMV(g->inverseTransform, D.x, D.y, D.z, 0.0f);
      float pb = .0001f * (1.0f / sqrtf(vdot(rdraw, rdraw)));
      kind = 2; nd = tr; no = vsub(P, vscale(ns, pb + OR_RAY_BIAS_AMOUNT));
      T = vmul(T, ld3(m->color));
    }
  } else if (m->hasReflective > 0) {
    kind = 1; nd = reflect3(ns, D); no = vadd(P, vscale(ns, OR_RAY_BIAS_AMOUNT));
    T = vmul(T, ld3(m->specularColor));
  } else {
    float nsv[3], out[3];
    st3(nsv, ns);
    or_hemisphere(nsv, u[0], u[1], out);
    kind = 0; nd = ld3(out); no = vadd(P, vscale(ns, OR_RAY_BIAS_AMOUNT));
    T = vmul(T, ld3(m->color));
  }
  st3(o, no); st3(d, nd); st3(thr, T);
  return kind;
}

void or_intersect_rays(const or_static_geom* geoms, int n_geoms, int n_rays, const float* o, const float* d,
                       int* id, float* t, float* p, float* n) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_rays; i++) {
    float pp[3] = {0, 0, 0}, nn[3] = {0, 0, 0}, tt;
    id[i] = or_closest_hit(geoms, n_geoms, o + 3 * i, d + 3 * i, &tt, pp, nn);
    t[i] = tt;
    memcpy(p + 3 * i, pp, sizeof(pp)); memcpy(n + 3 * i, nn, sizeof(nn));
  }
}

void or_raygen_batch(const or_camera_data* cam, const or_lens* lens, uint64_t seed, int n, const uint32_t* pixel,
                     const uint32_t* sample, float* o, float* d) {
  for (int i = 0; i < n; i++) or_raygen(cam, lens, seed, pixel[i], sample[i], o + 3 * i, d + 3 * i);
}

int or_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- (S) direct light sampling (README.md "Sphere surface point sampling") ----
 * Lights = sphere / cube geoms whose material has EMITTANCE > 0, in index order.  At every diffuse shading event that
 * is not the path's last allowed segment, ONE light is picked uniformly and ONE point y on it is drawn with the
 * reference's own area-weighted cube sampler / the sphere sampler above (uniforms: Philox block 65 + depth); the
 * shadow ray from the new path origin towards y is traced with the ordinary closest hit, and if it arrives on that
 * light at y (distance within 1e-3 relative + 1e-3 of |y - o|) the estimate  thr * (1/pi) * Le * cos_s * cos_l / t^2 * area * n_lights  is added,
 * weighted by the BALANCE HEURISTIC against the cosine-weighted direction sampling that continues the path:
 *   G = cos_s * cos_l / t^2,  K = area * n_lights / pi,  p_bsdf / p_light = G * K,  w_light = 1 / (1 + G*K).
 * The continuing path carries cos_b = dot(ns, new direction) (> 0); if it reaches a light by itself at its next hit, at
 * distance t with cosine cos_l' > 0 on the light, its emission is weighted by  w_bsdf = x / (1 + x),  x = (cos_b * cos_l' /
 * t^2) * K  (K of the light that was hit) -- the two weights sum to one for every direction both strategies can produce, so
 * the estimator stays unbiased and neither strategy's bad cases dominate the variance (round 1 suppressed the emission
 * instead: light sampling alone).  Specular and refractive events, and the camera ray, carry 0: their emission counts in
 * full.  area: the cube sampler's totalarea; spheres 4*pi/3 * (rx*ry + ry*rz
 * + rx*rz) (exact for uniform scale). */
typedef struct { int geom; float E[3]; float th[5]; float K; } or_light;  /* K = area * n_lights / pi = 1 / (pi * p_area) */
static int build_lights(const or_scene* sc, or_light* out, int cap) {
  int n = 0;
  float area[1024];
  if (cap > 1024) cap = 1024;
  for (int i = 0; i < sc->n_geoms && n < cap; i++) {
    const or_static_geom* g = &sc->geoms[i];
    if (g->type != 0 && g->type != 1) continue;
    if (!(sc->materials[g->materialid].emittance > 0)) continue;
    out[n].geom = i;
    if (g->type == 1) {
      area[n] = cube_thresholds(g, out[n].th);
    } else {
      float r[3];
      or_getRadiuses(g, r);
      area[n] = 4.1887903f * ((r[0] * r[1] + r[1] * r[2]) + r[0] * r[2]);
      memset(out[n].th, 0, sizeof(out[n].th));
    }
    n++;
  }
  for (int k = 0; k < n; k++) {
    const or_material* m = &sc->materials[sc->geoms[out[k].geom].materialid];
    float kk = (area[k] * (float)n) * 0.31830987f;
    out[k].K = kk;
    st3(out[k].E, vscale(vscale(ld3(m->color), m->emittance), kk));
  }
  return n;
}

/* (S) the per-pixel path loop the reference sketches in raytraceRay (src/raytraceKernel.cu:93-104) and the
 * running accumulation of cudaRaytraceCore (:108-165): one path per (pixel, sample); a path ends on a miss
 * (black background), on an emissive hit (adds throughput*color*emittance) or after max_depth segments. */
double or_render_ex(const or_scene* sc, uint32_t first_sample, uint32_t n_samples, int max_depth, uint64_t seed,
                    uint32_t pix_begin, uint32_t pix_end, float* sum_rgb, uint64_t* live, int threads,
                    uint64_t* shadow_rays) {
  if (max_depth > 64) max_depth = 64;
#ifdef _OPENMP
  int nt = threads > 0 ? threads : omp_get_max_threads();
#else
  int nt = 1;
  (void)threads;
#endif
  or_light lights_buf[1024];  /* 36 KB on the stack; more lights than that are ignored (build_lights caps) */
  or_light* lights = lights_buf;
  int n_lights = 0;
  if (sc->direct_lighting) n_lights = build_lights(sc, lights, 1024);
  double t0 = now_s();
#pragma omp parallel num_threads(nt)
  {
    uint64_t mylive[64], myshadow = 0;
    memset(mylive, 0, sizeof(mylive));
#pragma omp for schedule(dynamic, 256)
    for (int64_t pix = (int64_t)pix_begin; pix < (int64_t)pix_end; pix++) {
      for (uint32_t s = first_sample; s < first_sample + n_samples; s++) {
        float o[3], d[3], thr[3] = {1.0f, 1.0f, 1.0f};
        float cos_b = 0.0f;  /* > 0: the previous event was a diffuse bounce with a light sample (MIS weight at the next light) */
        or_raygen(&sc->cam, &sc->lens, seed, (uint32_t)pix, s, o, d);
        for (int depth = 0; depth < max_depth; depth++) {
          mylive[depth]++;
          float t, p[3], n[3];
          int id = or_closest_hit(sc->geoms, sc->n_geoms, o, d, &t, p, n);
          if (id < 0) break;
          float L[3];
          v3 din = ld3(d);
          int kind = or_shade(sc, id, t, p, n, seed, (uint32_t)pix, s, (uint32_t)depth, o, d, thr, L);
          if (kind == 3) {
            if (cos_b > 0) {
              float clp = -vdot(ld3(n), din);
              if (clp > 0) {
                float Kh = 0.0f;
                for (int k = 0; k < n_lights; k++) if (lights[k].geom == id) Kh = lights[k].K;
                float x = ((cos_b * clp) / (t * t)) * Kh;
                float wb = x / (1.0f + x);
                L[0] = L[0] * wb; L[1] = L[1] * wb; L[2] = L[2] * wb;
              }
            }
            sum_rgb[3 * pix + 0] += L[0];
            sum_rgb[3 * pix + 1] += L[1];
            sum_rgb[3 * pix + 2] += L[2];
            break;
          }
          cos_b = 0.0f;
          if (kind == 0 && n_lights > 0 && depth + 1 < max_depth) {
            v3 N = ld3(n);
            v3 ns = vdot(din, N) < 0 ? N : vneg(N);
            cos_b = vdot(ns, ld3(d));  /* d = the direction the bounce sampled */
            float v[4];
            rng4(seed, (uint32_t)pix, s, 65u + (uint32_t)depth, v);
            int li = (int)(v[0] * (float)n_lights);
            if (li > n_lights - 1) li = n_lights - 1;
            const or_light* Lt = &lights[li];
            const or_static_geom* gl = &sc->geoms[Lt->geom];
            v3 y;
            if (gl->type == 0) {
              float yy[3];
              or_sphere_point_u(gl, v[1], v[2], yy);
              y = ld3(yy);
            } else {
              y = cube_point_th(gl, Lt->th, v[1], v[2] - 0.5f, v[3] - 0.5f);
            }
            v3 wv = vsub(y, ld3(o));
            v3 wd = vnormalize(wv);
            float dy = vlength(wv);
            float cs = vdot(ns, wd);
            if (cs > 0) {
              myshadow++;
              float wdv[3], t2, p2[3], n2[3];
              st3(wdv, wd);
              int id2 = or_closest_hit(sc->geoms, sc->n_geoms, o, wdv, &t2, p2, n2);
              /* y is visible iff the ray arrives ON the light AT y (a point on the light's far side is hidden by the
               * light itself: same geom, shorter distance) */
              if (id2 == Lt->geom && (dy - t2) < 1e-3f * dy + 1e-3f) {
                float cl = -vdot(ld3(n2), wd);
                if (cl > 0) {
                  float G = (cs * cl) / (t2 * t2);
                  float wl = 1.0f / (1.0f + G * Lt->K);
                  v3 Ld = vscale(vscale(vmul(ld3(thr), ld3(Lt->E)), G), wl);
                  sum_rgb[3 * pix + 0] += Ld.x;
                  sum_rgb[3 * pix + 1] += Ld.y;
                  sum_rgb[3 * pix + 2] += Ld.z;
                }
              }
            }
          }
        }
      }
    }
#pragma omp critical
    {
      for (int i = 0; i < max_depth; i++) live[i] += mylive[i];
      if (shadow_rays) *shadow_rays += myshadow;
    }
  }
  return now_s() - t0;
}

double or_render(const or_scene* sc, uint32_t first_sample, uint32_t n_samples, int max_depth, uint64_t seed,
                 uint32_t pix_begin, uint32_t pix_end, float* sum_rgb, uint64_t* live, int threads) {
  return or_render_ex(sc, first_sample, n_samples, max_depth, seed, pix_begin, pix_end, sum_rgb, live, threads, NULL);
}
