/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the path-tracing hot path of CIS565-Fall-2014/Project3-Pathtracer.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or call this.  The product (project3-pathtracer_b200/) never does.
 *
 * Two kinds of functions live here (see pt_oracle.c for the per-function citations):
 *   (R) restatements of functions the reference IMPLEMENTS -- pinned bit-for-bit against
 *       the reference's own code (oracle/_ref, tests/golden/ref_vectors.json);
 *   (S) the SPECIFICATION of what the reference leaves as TODO stubs (box test, raygen,
 *       BSDF, RNG, path loop).  No reference code exists for these:
 *       "parity unpinned" against the reference for (S), pinned only by their own
 *       known-answer vectors (Philox: Random123 published KATs).
 *
 * Arithmetic contract: every operation is IEEE-754 binary32, round-to-nearest-even,
 * UNFUSED, in the order written, except the few binary64 steps that the reference's host
 * build performs (marked "f64 step").  Build with -ffp-contract=off and no fast-math.
 * The CUDA kernels follow the same contract, so results are comparable bit for bit.
 */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-compatible with the reference's types (src/sceneStructs.h:32-48,63-74, src/cudaMat4.h:18-23). */
typedef struct {
  int type;       /* GEOMTYPE: 0 SPHERE, 1 CUBE, 2 MESH (src/sceneStructs.h:14) */
  int materialid;
  float translation[3];
  float rotation[3];
  float scale[3];
  float transform[16];        /* cudaMat4: 4 ROWS x,y,z,w */
  float inverseTransform[16];
} or_static_geom; /* 172 bytes */

typedef struct {
  float color[3];
  float specularExponent;
  float specularColor[3];
  float hasReflective;
  float hasRefractive;
  float indexOfRefraction;
  float hasScatter;
  float absorptionCoefficient[3];
  float reducedScatterCoefficient;
  float emittance;
} or_material; /* 64 bytes */

typedef struct {
  float resolution[2];
  float position[3];
  float view[3];
  float up[3];
  float fov[2]; /* half-angles, degrees (src/scene.cpp:203-207) */
} or_camera_data; /* 52 bytes */

typedef struct {
  float aperture;       /* lens radius; 0 => pinhole */
  float focal_distance; /* distance of the plane of focus along the view axis */
} or_lens;

typedef struct {
  const or_static_geom* geoms;
  int n_geoms;
  const or_material* materials;
  int n_materials;
  or_camera_data cam;
  or_lens lens;
  int direct_lighting; /* 0: emission only where a path arrives (the default); 1: direct light sampling at diffuse hits */
} or_scene;

/* ---- (R) reference-pinned pieces ---- */
unsigned int or_hash(unsigned int a);
int or_epsilonCheck(float a, float b);
void or_ray_helpers(const float d[3], float inv[3], float sign[3]);
void or_multiplyMV(const float m[16], const float v[4], float out[3]);
void or_getPointOnRay(const float o[3], const float d[3], float t, float out[3]);
float or_sphereIntersectionTest(const or_static_geom* g, const float o[3], const float d[3], float p[3], float n[3]);
void or_getRadiuses(const or_static_geom* g, float out[3]);
/* src/intersections.h:133-175 with the host build's argument evaluation order; thrust minstd inside */
void or_getRandomPointOnCube(const or_static_geom* g, float randomSeed, float out[3]);
/* src/raytraceKernel.cu:29-36 (the stub renderer's per-pixel noise); reversed = the host build's draw order */
void or_generateRandomNumberFromThread(int W, int H, float time, int x, int y, int reversed, float out[3]);
void or_noise_image(int W, int H, float time, int reversed, float* out);
/* reference formula with libm sinf/cosf, for pinning only */
void or_hemisphere_ref(const float n[3], float xi1, float xi2, float out[3]);

/* ---- (S) specification of the stubbed pieces ---- */
float or_boxIntersectionTest(const or_static_geom* g, const float o[3], const float d[3], float p[3], float n[3]);
void or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float or_u01(uint32_t x);
void or_sincos_2pi(float u, float* s, float* c);
/* reference formula with or_sincos_2pi in place of libm; the one the path loop uses */
void or_hemisphere(const float n[3], float xi1, float xi2, float out[3]);
void or_reflect(const float n[3], const float i[3], float out[3]);
int or_refract(const float n[3], const float i[3], float ior_i, float ior_t, float out[3]); /* returns 1 on TIR */
void or_fresnel(const float n[3], const float i[3], float ior_i, float ior_t, const float refl[3],
                const float trans[3], int tir, float* R, float* T);
/* a12/a13/a15: surface points and directions; a20: Beer-Lambert transmission with a reproducible exp */
void or_cube_point_u(const or_static_geom* g, float u0, float u1, float u2, float out[3]);
void or_sphere_point_u(const or_static_geom* g, float u0, float u1, float out[3]);
void or_getRandomPointOnSphere(const or_static_geom* g, float randomSeed, float out[3]);
void or_getRandomDirectionInSphere(float xi1, float xi2, float out[3]);
float or_exp(float x);
void or_calculateTransmission(const float absorption[3], float distance, float out[3]);
void or_raygen(const or_camera_data* cam, const or_lens* lens, uint64_t seed, uint32_t pixel, uint32_t sample,
               float o[3], float d[3]);
/* closest hit in index order with strict '<'; returns geom index or -1 */
int or_closest_hit(const or_static_geom* geoms, int n_geoms, const float o[3], const float d[3], float* t,
                   float p[3], float n[3]);
/* One shading event. Returns 0 diffuse, 1 reflected, 2 transmitted (src/interactions.h:97-101),
 * 3 = emissive hit (path ends, radiance written to L). Updates o, d, thr in place. */
int or_shade(const or_scene* sc, int geom_id, float t, const float p[3], const float n[3], uint64_t seed, uint32_t pixel,
             uint32_t sample, uint32_t depth, float o[3], float d[3], float thr[3], float L[3]);

/* ---- batch / whole-frame entry points used by tests and the CPU baseline ---- */
void or_intersect_rays(const or_static_geom* geoms, int n_geoms, int n_rays, const float* o, const float* d,
                       int* id, float* t, float* p, float* n);
void or_random_points_batch(const or_static_geom* g, int n, const float* seed, float* out);
void or_points_u_batch(const or_static_geom* g, int n, const float* u, float* out);
void or_sphere_dirs_batch(int n, const float* xi1, const float* xi2, float* out);
void or_transmission_batch(int n, const float* absorption, const float* distance, float* out);
void or_raygen_batch(const or_camera_data* cam, const or_lens* lens, uint64_t seed, int n, const uint32_t* pixel,
                     const uint32_t* sample, float* o, float* d);
/* Render samples [first_sample, first_sample+n_samples) of pixels [pix_begin, pix_end) and ADD radiance to
 * sum_rgb (W*H*3 floats, indexed by pixel).  live[d] (max_depth entries, uint64) is incremented by the number of
 * paths for which closest-hit was evaluated at depth d.  threads<=0 => all (OpenMP).  Returns seconds of wall
 * time spent in the render loop. */
double or_render(const or_scene* sc, uint32_t first_sample, uint32_t n_samples, int max_depth, uint64_t seed,
                 uint32_t pix_begin, uint32_t pix_end, float* sum_rgb, uint64_t* live, int threads);
/* the same with direct light sampling counted: *shadow_rays (may be NULL) is incremented by the number of shadow
 * rays for which closest-hit was evaluated */
double or_render_ex(const or_scene* sc, uint32_t first_sample, uint32_t n_samples, int max_depth, uint64_t seed,
                    uint32_t pix_begin, uint32_t pix_end, float* sum_rgb, uint64_t* live, int threads,
                    uint64_t* shadow_rays);
int or_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
