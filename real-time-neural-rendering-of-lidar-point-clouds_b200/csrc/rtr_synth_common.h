/* Deterministic synthetic LiDAR-hall generator shared by the device generator (rtr_synth.cu),
 * and — through oracle/rtr_oracle.c — by the CPU checker.  Plain C99, integer-only until the last
 * step, so host and device produce bit-identical clouds without storing them.
 *
 * Scene (SURVEY.md §8 d "synthetic inputs"): a box-shaped hall LX x LY x LZ metres (z up) whose
 * six inner faces are sampled uniformly per area, plus NBOX clutter boxes standing on the floor.
 * Every surface is cut into 0.25 m x 0.25 m patches; patch p owns the contiguous index range
 * [p*k, (p+1)*k).  That is the ORDER the reference loader delivers: points grouped by 0.25 m cell,
 * arbitrary cell order, arrival order inside a cell (cloudreader.cpp:47-60, Octreegrid.h:162-170).
 * Range noise: sum of four uniform variates along the face normal, sigma ~ 2.2 mm.
 * Coordinates are fixed point, 2^-16 m, converted with one exact multiply.
 * Colour is a low-frequency procedural pattern of position, stored B,G,R (cloudreader.cpp:168).
 */
#ifndef RTR_SYNTH_COMMON_H
#define RTR_SYNTH_COMMON_H
#include <stdint.h>

#if defined(__CUDACC__)
#define RTR_HD __host__ __device__ __forceinline__
#else
#define RTR_HD static inline
#endif

#define RTR_SYNTH_MAX_FACES 128

typedef struct {
    int32_t o[3];     /* origin, fixed point 2^-16 m */
    int32_t au, av;   /* axis index (0,1,2) of the two in-plane directions */
    int32_t an;       /* axis index of the normal */
    int32_t nu, nv;   /* patches along u and v (0.25 m each) */
    uint32_t first;   /* first global patch index of this face */
} rtr_synth_face;

typedef struct {
    uint64_t seed;
    uint64_t n_points;
    uint64_t per_patch;      /* k = max(1, n_points / n_patches) */
    uint32_t n_patches;
    int32_t n_faces;
    rtr_synth_face faces[RTR_SYNTH_MAX_FACES];
} rtr_synth_scene;

RTR_HD uint64_t rtr_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* Point i of the scene -> fixed-point position (2^-16 m) and packed colour b|g<<8|r<<16|255<<24. */
RTR_HD void rtr_synth_point_fp(const rtr_synth_scene* s, uint64_t i, int32_t p[3], uint32_t* bgra) {
    uint32_t patch = (uint32_t)((i / s->per_patch) % s->n_patches);
    int lo = 0, hi = s->n_faces - 1;
    while (lo < hi) { /* last face with first <= patch */
        int mid = (lo + hi + 1) >> 1;
        if (s->faces[mid].first <= patch) lo = mid; else hi = mid - 1;
    }
    const rtr_synth_face* f = &s->faces[lo];
    uint32_t local = patch - f->first;
    int32_t pu = (int32_t)(local % (uint32_t)f->nu), pv = (int32_t)(local / (uint32_t)f->nu);
    uint64_t r0 = rtr_splitmix64(s->seed ^ (i * 0xD1342543DE82EF95ull));
    uint64_t r1 = rtr_splitmix64(r0);
    int32_t du = (int32_t)(r0 & 0x3FFF), dv = (int32_t)((r0 >> 14) & 0x3FFF); /* 0.25 m = 16384 */
    int32_t noise = (int32_t)((r0 >> 28) & 0xFF) + (int32_t)((r0 >> 36) & 0xFF) +
                    (int32_t)((r0 >> 44) & 0xFF) + (int32_t)((r0 >> 52) & 0xFF) - 510;
    p[0] = f->o[0]; p[1] = f->o[1]; p[2] = f->o[2];
    p[f->au] += pu * 16384 + du;
    p[f->av] += pv * 16384 + dv;
    p[f->an] += noise;
    int32_t b = 64 + ((p[0] >> 11) & 127), g = 64 + ((p[1] >> 11) & 127), r = 64 + ((p[2] >> 10) & 127);
    if (((p[0] >> 15) ^ (p[1] >> 15) ^ (p[2] >> 15)) & 1) { b += 40; g += 40; r += 40; }
    b += (int32_t)(r1 & 15) - 8; g += (int32_t)((r1 >> 4) & 15) - 8; r += (int32_t)((r1 >> 8) & 15) - 8;
    b = b < 0 ? 0 : b > 255 ? 255 : b; g = g < 0 ? 0 : g > 255 ? 255 : g; r = r < 0 ? 0 : r > 255 ? 255 : r;
    *bgra = (uint32_t)b | ((uint32_t)g << 8) | ((uint32_t)r << 16) | 0xFF000000u;
}

/* Exact: |p| < 2^24, scale is a power of two. */
RTR_HD float rtr_synth_fp_to_m(int32_t v) { return (float)v * (1.0f / 65536.0f); }

/* Host-side scene construction (used by both the product's bench support and the oracle).
 * lx, ly, lz in quarter metres; nbox clutter boxes laid out deterministically from the seed. */
static inline void rtr_synth_add_face(rtr_synth_scene* s, int32_t ox, int32_t oy, int32_t oz, int au, int av,
                                      int an, int nu, int nv) {
    if (s->n_faces >= RTR_SYNTH_MAX_FACES || nu <= 0 || nv <= 0) return;
    rtr_synth_face* f = &s->faces[s->n_faces++];
    f->o[0] = ox * 16384; f->o[1] = oy * 16384; f->o[2] = oz * 16384;
    f->au = au; f->av = av; f->an = an; f->nu = nu; f->nv = nv;
    f->first = s->n_patches;
    s->n_patches += (uint32_t)(nu * nv);
}

static inline void rtr_synth_build_scene(rtr_synth_scene* s, uint64_t seed, uint64_t n_points, int lx, int ly,
                                         int lz, int nbox) {
    s->seed = seed; s->n_points = n_points; s->n_patches = 0; s->n_faces = 0;
    rtr_synth_add_face(s, 0, 0, 0, 0, 1, 2, lx, ly);   /* floor   */
    rtr_synth_add_face(s, 0, 0, lz, 0, 1, 2, lx, ly);  /* ceiling */
    rtr_synth_add_face(s, 0, 0, 0, 0, 2, 1, lx, lz);   /* wall y=0  */
    rtr_synth_add_face(s, 0, ly, 0, 0, 2, 1, lx, lz);  /* wall y=ly */
    rtr_synth_add_face(s, 0, 0, 0, 1, 2, 0, ly, lz);   /* wall x=0  */
    rtr_synth_add_face(s, lx, 0, 0, 1, 2, 0, ly, lz);  /* wall x=lx */
    uint64_t r = rtr_splitmix64(seed ^ 0xC1A77E5ull);
    for (int b = 0; b < nbox; ++b) {
        r = rtr_splitmix64(r);
        int sx = 2 + (int)(r & 3), sy = 2 + (int)((r >> 2) & 3), sz = 2 + (int)((r >> 4) & 7); /* 0.5..2.25 m */
        if (sz > lz - 1) sz = lz - 1;
        int x0 = 1 + (int)((r >> 8) % (uint64_t)(lx - sx - 1)), y0 = 1 + (int)((r >> 24) % (uint64_t)(ly - sy - 1));
        rtr_synth_add_face(s, x0, y0, sz, 0, 1, 2, sx, sy);        /* top */
        rtr_synth_add_face(s, x0, y0, 0, 0, 2, 1, sx, sz);         /* y = y0 */
        rtr_synth_add_face(s, x0, y0 + sy, 0, 0, 2, 1, sx, sz);    /* y = y0+sy */
        rtr_synth_add_face(s, x0, y0, 0, 1, 2, 0, sy, sz);         /* x = x0 */
        rtr_synth_add_face(s, x0 + sx, y0, 0, 1, 2, 0, sy, sz);    /* x = x0+sx */
    }
    s->per_patch = n_points / (s->n_patches ? s->n_patches : 1);
    if (s->per_patch == 0) s->per_patch = 1;
}
#endif /* RTR_SYNTH_COMMON_H */
