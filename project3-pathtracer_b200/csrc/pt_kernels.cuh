// pt_kernels.cuh -- the wavefront kernels (sm_100a).
//
//   k_bounce<FIRST,LAST>  one path segment for every live path of the wavefront, fused:
//                           [FIRST: raygen + depth of field]  (raycastFromCameraKernel, src/raytraceKernel.cu:40-45)
//                           closest hit over SoA geometry staged in shared memory (src/intersections.h:74-117)
//                           BSDF sampling with Philox (calculateBSDF, src/interactions.h:99-104)
//                           radiance accumulation in HBM (the running image of src/raytraceKernel.cu:118-120,154)
//                           stream compaction of the survivors (README.md:63-70): warp ballot -> block scan ->
//                           decoupled look-back across tiles, survivors written once, in order, to the other
//                           ping-pong buffer.
//                         Persistent CTAs take tiles from a ticket counter, so a tile's predecessors are always
//                         resident and the look-back cannot deadlock; the live count never visits the host.
//   k_raygen_list / k_intersect_list   the same device functions on caller-supplied lists (parity entry points)
//   k_compact_u32                      the compaction primitive on its own
//   k_resolve_*                        accumulation buffer -> float RGB / uchar4 (sendImageToPBO, :58-89)
#pragma once
#include "pt_device.cuh"
#include "pt_pairs.cuh"

namespace ptd {

constexpr int kMaxDepth = 64;
constexpr int kTile = 256;  // paths per tile = threads per CTA
#ifndef PT_MIN_BLOCKS
#define PT_MIN_BLOCKS 4  // resident CTAs per SM the register allocation aims for (tuning knob, see profiles/)
#endif

// per-wavefront control block in HBM, zeroed before each wavefront
struct WfCtrl {
  uint32_t count[kMaxDepth + 1];     // count[d] = live paths entering depth d
  uint32_t tile_ctr[kMaxDepth + 1];  // ticket counter of the depth-d launch
};

// ---- decoupled look-back (Merrill & Garland 2016) on 64-bit status words ----
// word = epoch[63:34] | state[33:32] | value[31:0]; a word is valid only if its epoch equals the launch's, so the
// array never needs clearing between launches (or CUDA-graph replays).
constexpr uint32_t kStAggregate = 1u, kStPrefix = 2u;
__device__ __forceinline__ uint64_t st_pack(uint32_t epoch, uint32_t state, uint32_t value) {
  return ((uint64_t)epoch << 34) | ((uint64_t)state << 32) | (uint64_t)value;
}
__device__ __forceinline__ void st_store(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t st_load(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Called by ONE full warp of the CTA that owns `tile`.  aggregate = number of survivors in this tile.
// Returns the number of survivors in all earlier tiles (exclusive prefix) to every lane.
__device__ __forceinline__ uint32_t lookback_exclusive(uint64_t* status, uint32_t tile, uint32_t epoch,
                                                       uint32_t aggregate) {
  const uint32_t lane = threadIdx.x & 31u;
  if (lane == 0) st_store(status + tile, st_pack(epoch, tile == 0 ? kStPrefix : kStAggregate, aggregate));
  if (tile == 0) return 0;
  uint32_t exclusive = 0;
  int look = (int)tile - 1;
  for (;;) {
    const int t = look - (int)lane;
    uint32_t state = kStPrefix, value = 0;  // tiles before the first one: empty prefix
    if (t >= 0) {
      uint64_t w;
      do { w = st_load(status + t); } while ((uint32_t)(w >> 34) != epoch);
      state = (uint32_t)(w >> 32) & 3u;
      value = (uint32_t)w;
    }
    const uint32_t pmask = __ballot_sync(0xffffffffu, state == kStPrefix);
    const int firstp = pmask ? (__ffs(pmask) - 1) : 31;
    uint32_t contrib = ((int)lane <= firstp) ? value : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    exclusive += contrib;
    if (pmask) break;
    look -= 32;
  }
  if (lane == 0) st_store(status + tile, st_pack(epoch, kStPrefix, exclusive + aggregate));
  return exclusive;
}

// Block-level part shared by k_bounce and k_compact_u32: every thread passes its flag, gets its output slot.
// s_warp (>= 8 words) and s_base are shared scratch.  All threads of the CTA must call.
// Returns the global slot of this thread's item (meaningful if keep) and the tile's inclusive total via *incl.
__device__ __forceinline__ uint32_t compact_slot(bool keep, uint64_t* status, uint32_t tile, uint32_t epoch,
                                                 uint32_t* s_warp, uint32_t* s_base, uint32_t* incl) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
  const uint32_t lane_off = __popc(ballot & ((1u << lane) - 1u));
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    const uint32_t nw = blockDim.x >> 5;
    uint32_t c = lane < nw ? s_warp[lane] : 0u;
    uint32_t incl_w = c;  // inclusive scan over the warps' counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(0xffffffffu, incl_w, o);
      if ((int)lane >= o) incl_w += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl_w, 31);
    const uint32_t excl = lookback_exclusive(status, tile, epoch, total);
    if (lane < nw) s_warp[lane] = incl_w - c;  // exclusive offset of each warp within the tile
    if (lane == 0) { s_base[0] = excl; s_base[1] = excl + total; }
  }
  __syncthreads();
  *incl = s_base[1];
  return s_base[0] + s_warp[warp] + lane_off;
}

struct BounceParams {
  const float4 *in_o, *in_d, *in_t;  // path state in:  (origin.xyz, pixel) (direction.xyz, sample) (throughput.xyz, -)
  float4 *out_o, *out_d, *out_t;     // survivors out, compacted
  float4* accum;                     // per-pixel radiance sums
  GeomSoA g;                         // per-geom rows in HBM (winner's normal / material lookup)
  int n_geoms, geom_cap;             // geoms per shared-memory chunk (scalar path)
  PairSoA pairs;                     // type-homogeneous geom pairs, interleaved (packed path)
  int pair_cap;                      // pairs per shared-memory chunk
  PkConsts kc;                       // run-time 1.0 / -0.0 / -1.0 / 0.0 for the packed arithmetic (pt_pairs.cuh)
  const float4* mats;                // 4 float4 per material
  RaygenConsts cam;
  WfCtrl* ctrl;
  uint64_t* status;
  uint32_t epoch, depth;
  uint64_t seed;
  uint32_t first_sample, n_first;    // FIRST only: paths to generate = npix * samples in this wavefront
};

// Work decomposition of k_bounce: a CTA tile is kTileRays = 1024 consecutive paths, cut into 32 sub-tiles of one
// warp each.  The 8 warps of a CTA GRAB sub-tiles from a shared-memory counter, so no warp waits for a slower one
// inside a tile (profiles/r01_k_bounce_v1_*: with one barrier-delimited 256-path tile per step, 57 % of the resident
// warps were parked at __syncthreads()).  Survivors are staged in shared memory in sub-tile-local slots; after
// one barrier a single warp scans the 32 sub-tile counts and runs the decoupled look-back once per 1024 paths, and
// all warps copy the staged survivors out, coalesced and in order (the compaction stays stable).
constexpr int kSubRays = 32;
#ifndef PT_SUB_PER_TILE
#define PT_SUB_PER_TILE 32  // sub-tiles per CTA tile (tuning knob: staging shared memory = 1536 B per sub-tile)
#endif
constexpr int kSubPerTile = PT_SUB_PER_TILE;
constexpr int kTileRays = kSubRays * kSubPerTile;
// shared memory after the geometry: staged survivors (3 float4 arrays of 1024) + per-sub-tile counts and offsets
__host__ __device__ inline size_t stage_smem_bytes() { return (size_t)kTileRays * 3 * sizeof(float4) + 2 * kSubPerTile * sizeof(uint32_t); }

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kTile, PT_MIN_BLOCKS) k_bounce(const BounceParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_tile, s_next;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // ---- geometry: staged once per CTA; scenes too large for shared memory are read through L1/L2 instead ----
#ifdef PT_SCALAR_HIT
  const bool staged = P.n_geoms <= P.geom_cap;
  GeomSmem gs = carve_geom_smem(smem_raw, P.geom_cap);
  unsigned char* after_geom = smem_raw + geom_smem_bytes(P.geom_cap);
  if (staged) stage_geoms(P.g, 0, P.n_geoms, gs);
#else
  const bool staged = P.pairs.n_pairs <= P.pair_cap;
  PairSmem ps = carve_pair_smem(smem_raw, P.pair_cap);
  unsigned char* after_geom = smem_raw + pair_smem_bytes(P.pair_cap);
  const Pk pk = make_pk(P.kc);
  if (staged) stage_pairs(P.pairs, 0, P.pairs.n_pairs, ps);
#endif
  float4* s_o = reinterpret_cast<float4*>(after_geom);
  float4* s_d = s_o + kTileRays;
  float4* s_t = s_d + kTileRays;
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_t + kTileRays);
  uint32_t* s_off = s_cnt + kSubPerTile;

  const uint32_t n_in = FIRST ? P.n_first : P.ctrl->count[P.depth];
  if (FIRST && blockIdx.x == 0 && threadIdx.x == 0) P.ctrl->count[0] = n_in;
  const uint32_t n_tiles = (n_in + kTileRays - 1) / kTileRays;

  for (;;) {
    __syncthreads();  // geometry staged; previous tile fully copied out
    if (threadIdx.x == 0) { s_tile = atomicAdd(&P.ctrl->tile_ctr[P.depth], 1u); s_next = 0u; }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= n_tiles) break;
    const uint32_t tile_base = tile * kTileRays;
    const uint32_t n_sub = min((uint32_t)kSubPerTile, (n_in - tile_base + kSubRays - 1) / kSubRays);

    // ---- phase 1: every warp pulls sub-tiles until the tile is exhausted ----
    for (;;) {
      uint32_t sub = 0;
      if (lane == 0) sub = atomicAdd(&s_next, 1u);
      sub = __shfl_sync(0xffffffffu, sub, 0);
      if (sub >= n_sub) break;
      const uint32_t idx = tile_base + sub * kSubRays + lane;
      const bool valid = idx < n_in;

      f3 o = mk(0, 0, 0), d = mk(0, 0, 1), thr = mk(1, 1, 1);
      uint32_t pixel = 0, sample = 0;
      if (valid) {
        if (FIRST) {
          pixel = idx % P.cam.npix;
          sample = P.first_sample + idx / P.cam.npix;
          raygen(P.cam, P.seed, pixel, sample, o, d);
        } else {
          const float4 a = __ldcs(P.in_o + idx), b = __ldcs(P.in_d + idx), c = __ldcs(P.in_t + idx);
          o = mk(a.x, a.y, a.z); pixel = __float_as_uint(a.w);
          d = mk(b.x, b.y, b.z); sample = __float_as_uint(b.w);
          thr = mk(c.x, c.y, c.z);
        }
      }

      Hit h;
      h.t = INFINITY; h.id = -1; h.p = mk(0, 0, 0); h.ncode = 0;
      if (valid) {
#ifdef PT_SCALAR_HIT
        if (staged) {
          closest_hit_chunk(gs, 0, P.n_geoms, o, d, h);
        } else {
          GeomSmem gg;  // the same rows, straight from HBM (cached)
          gg.inv0 = const_cast<float4*>(P.g.inv0); gg.inv1 = const_cast<float4*>(P.g.inv1); gg.inv2 = const_cast<float4*>(P.g.inv2);
          gg.fwd0 = const_cast<float4*>(P.g.fwd0); gg.fwd1 = const_cast<float4*>(P.g.fwd1); gg.fwd2 = const_cast<float4*>(P.g.fwd2);
          gg.meta = const_cast<int2*>(P.g.meta);
          closest_hit_chunk(gg, 0, P.n_geoms, o, d, h);
        }
#else
        if (staged) {
          closest_hit_pairs(pk, ps, P.pairs.n_pairs, P.g, o, d, h);
        } else {
          PairSmem pg;  // the same pairs, straight from HBM (cached)
          pg.q = const_cast<float4*>(P.pairs.q); pg.meta = const_cast<int4*>(P.pairs.meta); pg.cap = P.pairs.n_pairs;
          closest_hit_pairs(pk, pg, P.pairs.n_pairs, P.g, o, d, h);
        }
#endif
      }

      bool alive = false;
      if (valid && h.id >= 0) {
        const int gi = h.id;
        // the winner's own rows: from HBM through L1 (a handful of distinct addresses per warp)
        const float4 f0 = __ldg(P.g.fwd0 + gi), f1 = __ldg(P.g.fwd1 + gi), f2 = __ldg(P.g.fwd2 + gi);
        const int mat = __ldg(P.g.meta + gi).y;
        const f3 n = hit_normal(f0, f1, f2, h);
        MatRows m;
        m.a = __ldg(P.mats + 4 * mat); m.b = __ldg(P.mats + 4 * mat + 1);
        m.c = __ldg(P.mats + 4 * mat + 2); m.d = __ldg(P.mats + 4 * mat + 3);
        f3 L;
        const int kind = shade(m, P.g, gi, h.p, n, P.seed, pixel, sample, P.depth, o, d, thr, L);
        if (kind == 3) {
          float* px = reinterpret_cast<float*>(P.accum + pixel);
          atomicAdd(px + 0, L.x);
          atomicAdd(px + 1, L.y);
          atomicAdd(px + 2, L.z);
        } else {
          alive = true;
        }
      }

      if (!LAST) {
        // warp ballot -> rank of each survivor inside its sub-tile; survivors staged in sub-tile-local slots
        const uint32_t ballot = __ballot_sync(0xffffffffu, alive);
        if (alive) {
          const uint32_t slot = sub * kSubRays + __popc(ballot & ((1u << lane) - 1u));
          s_o[slot] = make_float4(o.x, o.y, o.z, __uint_as_float(pixel));
          s_d[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(sample));
          s_t[slot] = make_float4(thr.x, thr.y, thr.z, 0.0f);
        }
        if (lane == 0) s_cnt[sub] = __popc(ballot);
      }
    }
    if (LAST) continue;

    __syncthreads();  // all sub-tiles of this tile are staged
    // ---- phase 2: block scan over the 32 sub-tile counts + ONE decoupled look-back for the whole tile ----
    if (warp == 0) {
      static_assert(kSubPerTile <= 32, "one warp scans the sub-tile counts");
      const uint32_t c = lane < n_sub ? s_cnt[lane] : 0u;
      uint32_t incl = c;
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o2);
        if ((int)lane >= o2) incl += v;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
#ifdef PT_ATOMIC_COMPACT
      // experiment: unordered slot reservation (one atomic per tile) instead of the ordered look-back
      uint32_t excl = 0;
      if (lane == 0) excl = atomicAdd(&P.ctrl->count[P.depth + 1], total);
      excl = __shfl_sync(0xffffffffu, excl, 0);
      if (lane < kSubPerTile) s_off[lane] = excl + incl - c;
#else
      const uint32_t excl = lookback_exclusive(P.status, tile, P.epoch, total);
      if (lane < kSubPerTile) s_off[lane] = excl + incl - c;
      if (tile == n_tiles - 1 && lane == 0) P.ctrl->count[P.depth + 1] = excl + total;
#endif
    }
    __syncthreads();
    // ---- phase 3: coalesced, ordered copy-out of the staged survivors ----
    for (uint32_t sub = warp; sub < n_sub; sub += kTile / 32) {
      const uint32_t cnt = s_cnt[sub], off = s_off[sub];
      if (lane < cnt) {
        __stcs(P.out_o + off + lane, s_o[sub * kSubRays + lane]);
        __stcs(P.out_d + off + lane, s_d[sub * kSubRays + lane]);
        __stcs(P.out_t + off + lane, s_t[sub * kSubRays + lane]);
      }
    }
  }
}

// live_total[d] += count[d]; one tiny launch per wavefront
__global__ void k_accum_counts(const WfCtrl* ctrl, unsigned long long* live_total, int max_depth) {
  int d = threadIdx.x;
  if (d < max_depth) live_total[d] += ctrl->count[d];
}

// ---- parity entry points ----
__global__ void k_raygen_list(RaygenConsts C, uint64_t seed, int n, const uint32_t* pixel, const uint32_t* sample,
                              float* o, float* d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  f3 oo, dd;
  raygen(C, seed, pixel[i], sample[i], oo, dd);
  o[3 * i] = oo.x; o[3 * i + 1] = oo.y; o[3 * i + 2] = oo.z;
  d[3 * i] = dd.x; d[3 * i + 1] = dd.y; d[3 * i + 2] = dd.z;
}

__global__ void __launch_bounds__(kTile) k_intersect_list(GeomSoA g, int n_geoms, int geom_cap, PairSoA pairs, int pair_cap,
                                                          PkConsts kc, int n, const float* o, const float* d, int* id, float* t,
                                                          float* p, float* nrm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  f3 oo = mk(0, 0, 0), dd = mk(0, 0, 1);
  if (valid) { oo = mk(o[3 * i], o[3 * i + 1], o[3 * i + 2]); dd = mk(d[3 * i], d[3 * i + 1], d[3 * i + 2]); }
  Hit h;
  h.t = INFINITY; h.id = -1; h.p = mk(0, 0, 0); h.ncode = 0;
#ifdef PT_SCALAR_HIT
  const GeomSmem gs = carve_geom_smem(smem_raw, geom_cap);
  for (int c0 = 0; c0 < n_geoms; c0 += geom_cap) {
    const int cnt = min(geom_cap, n_geoms - c0);
    __syncthreads();
    stage_geoms(g, c0, cnt, gs);
    __syncthreads();
    if (valid) closest_hit_chunk(gs, c0, cnt, oo, dd, h);
  }
#else
  const PairSmem ps = carve_pair_smem(smem_raw, pair_cap);
  const Pk pk = make_pk(kc);
  for (int c0 = 0; c0 < pairs.n_pairs; c0 += pair_cap) {
    const int cnt = min(pair_cap, pairs.n_pairs - c0);
    __syncthreads();
    stage_pairs(pairs, c0, cnt, ps);
    __syncthreads();
    if (valid) closest_hit_pairs(pk, ps, cnt, g, oo, dd, h);
  }
#endif
  if (!valid) return;
  f3 nn = mk(0, 0, 0);
  if (h.id >= 0) nn = hit_normal(__ldg(g.fwd0 + h.id), __ldg(g.fwd1 + h.id), __ldg(g.fwd2 + h.id), h);
  id[i] = h.id;
  t[i] = h.id >= 0 ? h.t : -1.0f;
  p[3 * i] = h.p.x; p[3 * i + 1] = h.p.y; p[3 * i + 2] = h.p.z;
  nrm[3 * i] = nn.x; nrm[3 * i + 1] = nn.y; nrm[3 * i + 2] = nn.z;
}

// stream compaction on its own: same ballot / block scan / look-back code as k_bounce
__global__ void __launch_bounds__(kTile) k_compact_u32(const uint32_t* values, const uint8_t* flags, uint32_t n,
                                                       uint32_t* out, uint32_t* n_out, uint32_t* ticket,
                                                       uint64_t* status, uint32_t epoch) {
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_warp[kTile / 32];
  __shared__ uint32_t s_base[2];
  const uint32_t n_tiles = (n + kTile - 1) / kTile;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= n_tiles) break;
    const uint32_t idx = tile * kTile + threadIdx.x;
    const bool keep = idx < n && flags[idx] != 0;
    uint32_t incl;
    const uint32_t slot = compact_slot(keep, status, tile, epoch, s_warp, s_base, &incl);
    if (keep) out[slot] = values[idx];
    if (tile == n_tiles - 1 && threadIdx.x == 0) *n_out = incl;
  }
}

// ---- exhaustive self-test of the packed IEEE sqrt / reciprocal (pt_pairs.cuh) against the scalar operators ----
// Every one of the 2^32 bit patterns goes through both halves (paired with a different pattern so that mixed
// fast-path / fallback pairs occur).  NaN results compare equal to NaN results.
__global__ void k_selftest_packed(PkConsts kc, unsigned long long* bad_sqrt, unsigned long long* bad_rcp) {
  const Pk pk = make_pk(kc);
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long bs = 0, br = 0;
  for (uint64_t k = i; k < (1ull << 32); k += stride) {
    const uint32_t a = (uint32_t)k, b = a * 2654435761u + 12345u;
    const float x = __uint_as_float(a), y = __uint_as_float(b);
    const f2 s = sqrt2_ieee(pk, make_float2(x, y)), r = rcp2_ieee(pk, make_float2(x, y));
    const float sx = sqrtf(x), sy = sqrtf(y), rx = 1.0f / x, ry = 1.0f / y;
    auto same = [](float u, float v) { return __float_as_uint(u) == __float_as_uint(v) || (u != u && v != v); };
    bs += !same(s.x, sx) + !same(s.y, sy);
    br += !same(r.x, rx) + !same(r.y, ry);
  }
  if (bs) atomicAdd(bad_sqrt, bs);
  if (br) atomicAdd(bad_rcp, br);
}

// ---- image out ----
// packed float RGB (renderCam->image layout) = sum * 1 or sum / spp
__global__ void k_resolve_rgb(const float4* accum, uint32_t npix, float spp, int divide, float* rgb) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  float4 a = accum[i];
  if (divide) { a.x = a.x / spp; a.y = a.y / spp; a.z = a.z / spp; }
  rgb[3 * i] = a.x; rgb[3 * i + 1] = a.y; rgb[3 * i + 2] = a.z;
}
__global__ void k_upload_rgb(const float* rgb, uint32_t npix, float4* accum) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  accum[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 0.0f);
}
// sendImageToPBO, src/raytraceKernel.cu:58-89: c = image*255.0; c > 255 -> 255; stored to uchar (truncation), w = 0.
// image*255.0 is a binary64 product of two binary32-representable values rounded to float, which equals the
// binary32 product (double rounding is innocuous for a single multiply).
__global__ void k_resolve_rgba8(const float4* accum, uint32_t npix, float spp, uchar4* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  float4 a = accum[i];
  float r = (a.x / spp) * 255.0f, g = (a.y / spp) * 255.0f, b = (a.z / spp) * 255.0f;
  if (r > 255) r = 255;
  if (g > 255) g = 255;
  if (b > 255) b = 255;
  uchar4 px;
  px.x = (unsigned char)r; px.y = (unsigned char)g; px.z = (unsigned char)b; px.w = 0;
  out[i] = px;
}

}  // namespace ptd
