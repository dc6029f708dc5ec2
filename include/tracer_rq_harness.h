/* include/tracer_rq_harness.h -- workload generators for tests and bench ("trqh").
 *
 * NOT part of the drop-in boundary (that is tracer_rq.h). These restate, on the host, the few
 * pieces of the reference's CALLERS that are needed to produce the rays the BASELINE configs
 * name: the per-pixel PCG32 (RT_Metal/Metal/Random.metal:3-25, Tracer/pcg_basic.c:60-67), the
 * camera (Tracer.mm:87-125, Camera.hh:59-69, Render.metal:523-527), the diffuse bounce spawn
 * (Render.metal:447-475 with Math.hh:62-74 offset_ray and Sampling.hh:18-34,79-99,125-129), and
 * the NEE shadow-ray spawn (Render.metal:313-337, Square.hh:40-58). Ray / hit structs are those
 * of tracer_rq.h.
 */
#ifndef TRACER_RQ_HARNESS_H
#define TRACER_RQ_HARNESS_H

#include "tracer_rq.h"

#ifdef __cplusplus
extern "C" {
#endif

/* randomF(): ldexp(float(pcg32_random_r()), -32) for one stream (seed, seq). */
void trqh_pcg32_fill_f32(uint64_t seed, uint64_t seq, uint64_t n, float* out);
void trqh_pcg32_fill_u32(uint64_t seed, uint64_t seq, uint64_t n, uint32_t* out);

/* Ray(o, d) normalising ctor (Ray.hh:21-23) applied to n (o,d) pairs given as trq_ray. */
void trqh_normalize_rays(trq_ray* rays, uint64_t n);
/* offset_ray(p, n)  Math.hh:62-74 */
void trqh_offset_ray(const float p[3], const float n[3], float out[3]);

/* C5-style triangle soup: centre uniform in [0,1]^3, each vertex = centre + extent*(xi-0.5)
 * per component, PCG32(seed, seq = triangle index). Writes 3*nTri vertices (32 B each, flat face
 * normal) and 3*nTri indices. */
void trqh_make_soup(uint32_t nTri, uint64_t seed, float extent, void* triList, uint32_t* idxList);

/* Uniform incoherent rays in an axis-aligned box: origin uniform in [lo,hi]^3, direction uniform on
 * the sphere (z = 2xi-1, phi = 2*pi*xi), normalised by the Ray ctor; PCG32(seed, seq = first+i). */
void trqh_gen_random_rays(uint64_t first, uint64_t n, uint64_t seed, const float lo[3], const float hi[3],
                          float tmax, trq_ray* rays);

/* MakeCamera (Tracer.mm:87-125): out = {lookFrom, u, v, vertical, horizontal, cornerLowLeft}, 3 floats each. */
void trqh_make_camera(const float lookFrom[3], const float lookAt[3], const float viewUp[3],
                      float vfov, float aspect, float focus_dist, float out[18]);

/* MakeCamera + castRay with aperture 0: pixel (x,y) -> s = x/W, t = y/H (no jitter), row-major
 * y*W + x. vfov in radians as the reference passes it. */
void trqh_gen_camera_rays(const float lookFrom[3], const float lookAt[3], const float viewUp[3],
                          float vfov, float aspect, float focus_dist, uint32_t W, uint32_t H,
                          trq_ray* rays);

/* Diffuse bounce: for every record with hit != 0, PCG32(seed = seedBase + i, seq = 1), uu =
 * sample2D(); origin = offset_ray(p, sn); dir = stw * CosineSampleHemisphere(uu) with
 * CoordinateSystem(sn); tmax = FLT_MAX. Rays are compacted; srcIndex[k] (optional) receives the
 * index of the record ray k came from. Returns the number of rays written. */
uint64_t trqh_gen_bounce_rays(const trq_hit_record* recs, uint64_t n, uint64_t seedBase,
                              trq_ray* rays, uint32_t* srcIndex);

/* NEE shadow ray toward a point sampled on one of two light squares (272-B Square layout):
 * PCG32(seedBase + i, 1): uu = sample2D(), pick = random() < 0.5 ? lightA : lightB;
 * Square::sample (Square.hh:40-58); ray = Ray(offset_ray(p,sn), normalize(lsr.p - origin)),
 * tmax = length(lsr.p - origin). Compacted like trqh_gen_bounce_rays. */
uint64_t trqh_gen_shadow_rays(const trq_hit_record* recs, uint64_t n, uint64_t seedBase,
                              const void* lightA, const void* lightB, trq_ray* rays, uint32_t* srcIndex);

/* The same two producers drawing from the reference's per-pixel RNG state texture (RGBA32Uint, 4 x uint32 per pixel:
 * state >> 32, state, inc >> 32, inc -- the layout exRNG writes, Render.hh:109-120; see trq_rng_frame_begin in
 * tracer_rq.h for toRNG's entry conversion): record i belongs to pixel pixelOf[i] (NULL: i), its stream is loaded from
 * rngState[4 * pixel], advanced by the draws and stored back; srcIndex[k] receives the PIXEL, so it can be passed as
 * pixelOf of the next wave. rngState == NULL falls back to PCG32(seedBase + i, 1). */
void     trqh_rng_frame_begin(uint32_t* rngState, uint64_t nPixels);   /* host twin of trq_rng_frame_begin */
uint64_t trqh_gen_bounce_rays_rng(const trq_hit_record* recs, uint64_t n, uint64_t seedBase, const uint32_t* pixelOf,
                                  uint32_t* rngState, trq_ray* rays, uint32_t* srcIndex);
uint64_t trqh_gen_shadow_rays_rng(const trq_hit_record* recs, uint64_t n, uint64_t seedBase, const uint32_t* pixelOf,
                                  uint32_t* rngState, const void* lightA, const void* lightB, trq_ray* rays, uint32_t* srcIndex);

#ifdef __cplusplus
}
#endif
#endif
