/*
 * rt_oracle.c -- CPU oracle for the traditional sphere ray-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C, IEEE-double restatement of
 * the reference's algorithm.  It is the *checker* for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  Nothing under ray-tracer-v1_b200/ links or
 * imports it, and the product path has no CPU fallback.
 *
 * Parity pin: every function here is checked against outputs of the UNMODIFIED
 * reference Python modules (imported from /root/reference by
 * oracle/gen_golden.py, vectors committed under tests/golden/) and against the
 * one valid known-answer probe the reference itself stores
 * (RL/Marbles 1.ipynb cell 7).  See tests/test_oracle_golden.py.
 *
 * All arithmetic is double and follows the reference's operation ORDER
 * (Python float == C double; build with -ffp-contract=off so no FMA fusing).
 * Citations are path:line relative to the reference root.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))
#define ORC_NO_ID INT32_MIN

/* ------------------------------------------------------------------ scene */
typedef struct {
    /* spheres: object.py:4-9, material.py:4-8 */
    int32_t n;
    const double *centre;   /* [n,3] */
    const double *radius;   /* [n]   */
    const double *material; /* [n,4] reflective, transparent, emitive, refractive_index */
    const double *colour;   /* [n,3] */
    const int32_t *ids;     /* [n]   */
    /* global lights: light.py:12-21 */
    int32_t nG;
    const double *g_vec;    /* [nG,3] */
    const double *g_col;    /* [nG,3] */
    const double *g_strength;
    const double *g_max_angle;
    const int32_t *g_func;
    /* point lights: light.py:25-37 */
    int32_t nP;
    const int32_t *p_id;
    const double *p_pos;    /* [nP,3] */
    const double *p_col;    /* [nP,3] */
    const double *p_strength;
    const double *p_max_angle;
    const int32_t *p_func;
    double bg[3];
    /* Algorithm-B light list (TraditionalRenderer.light_sources / small_lights,
       fb_vs_traditional_chandelier.py:401-402) */
    int32_t nL;
    const double *l_centre; /* [nL,3] */
    const double *l_colour; /* [nL,3] */
    const int32_t *l_index; /* [nL] scene index of the light sphere, -1 if not in scene */
    const uint8_t *small;   /* [n] sphere is in small_lights */
} orc_scene;

typedef struct { double x, y, z; } v3;

typedef struct {
    int hit;            /* intersects */
    double t;           /* Intersection.distance (signed!) */
    v3 p, n;            /* point, normal */
    int idx;            /* scene index of object */
    int bounces, through;
} isect;

/* ------------------------------------------------------------- vector.py */
static inline v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }   /* vector.py:25-31 */
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }   /* vector.py:33-39 */
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }                        /* vector.py:41-47 */
static inline v3 vscale(v3 a, double l) { return V(a.x * l, a.y * l, a.z * l); }   /* vector.py:49-56 */
static inline double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }/* vector.py:94-95 */
static inline double vmag(v3 a) { return sqrt(vdot(a, a)); }                       /* vector.py:105-108 */
static inline v3 vnorm(v3 a) { double m = vmag(a); return V(a.x / m, a.y / m, a.z / m); } /* vector.py:110-112 */
static inline v3 vcross(v3 a, v3 b) {                                              /* vector.py:97-103 */
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline double vdist(v3 a, v3 b) {                                           /* vector.py:58-59 */
    double dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
    return sqrt(dx * dx + dy * dy + dz * dz);
}
static inline double vangle(v3 a, v3 b) {                                          /* vector.py:61-62 */
    return acos(vdot(a, b) / (vmag(a) * vmag(b)));
}
static inline v3 vreflect(v3 self, v3 B) {                                         /* vector.py:64-67 */
    v3 v = vnorm(self), n = vnorm(B);
    return vnorm(vsub(v, vscale(n, 2 * vdot(v, n))));
}
/* vector.py:69-92; returns 0 on total internal reflection (Python False) */
static inline int vrefract(v3 self, v3 B, double ra, double rb, v3 *out) {
    v3 v = vnorm(self), nrm = vnorm(B);
    double n = ra / rb;
    double cosI = vdot(v, nrm);
    if (cosI < -1) cosI = -1;
    if (cosI > 1) cosI = 1;
    if (cosI < 0) cosI = -cosI;
    double k = 1 - (n * n) * (1 - cosI * cosI);
    if (k < 0) return 0;
    *out = vnorm(vadd(vscale(v, n), vscale(nrm, n * cosI - sqrt(k))));
    return 1;
}
/* vector.py:117-127 (row-vector times R) */
static inline v3 vrotate(v3 s, v3 ang) {
    double a = ang.x, b = ang.y, c = ang.z;
    double R[3][3] = {
        {cos(c) * cos(b) * cos(a) - sin(c) * sin(a), cos(c) * cos(b) * sin(a) + sin(c) * cos(a), -cos(c) * sin(b)},
        {-sin(c) * cos(b) * cos(a) - cos(c) * sin(a), -sin(c) * cos(b) * sin(a) + cos(c) * cos(a), sin(c) * sin(b)},
        {sin(b) * cos(a), sin(b) * sin(a), cos(b)}};
    return V(s.x * R[0][0] + s.y * R[1][0] + s.z * R[2][0],
             s.x * R[0][1] + s.y * R[1][1] + s.z * R[2][1],
             s.x * R[0][2] + s.y * R[1][2] + s.z * R[2][2]);
}

static inline v3 ld3(const double *p, int i) { return V(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

/* ---------------------------------------------------------------- philox */
/* Philox4x32-10 (Salmon et al., SC'11) -- the counter-based RNG the CUDA path
   uses, keyed (pixel, sample, slot>>1); words 2*(slot&1)+{0,1} feed slot.
   slot 0 = camera jitter, slot k+1 = diffuse bounce at depth k.               */
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
ORC_API void orc_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32_10(ctr, key, out); }

/* two uniforms in [0,1) with 24-bit mantissas (exactly representable in f32) */
static inline void rng_pair(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, double *u0, double *u1) {
    uint32_t ctr[4] = {pixel, sample, slot >> 1, 0x52544232u /* "RTB2" */};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    uint32_t a = o[2 * (slot & 1)], b = o[2 * (slot & 1) + 1];
    *u0 = (double)(a >> 8) * (1.0 / 16777216.0);
    *u1 = (double)(b >> 8) * (1.0 / 16777216.0);
}
ORC_API void orc_rng_pair(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, double *out2) {
    rng_pair(seed, pixel, sample, slot, &out2[0], &out2[1]);
}

/* ---------------------------------------------------------------- ray.py */
/* Ray.sphereDiscriminant, ray.py:73-107.  D must already be normalised
   (Ray.__init__, ray.py:69-71).                                               */
static inline isect sphere_discriminant(v3 O, v3 D, v3 C, double r, int point) {
    isect it; memset(&it, 0, sizeof it); it.idx = -1;
    v3 L = vsub(C, O);
    double tca = vdot(L, D);
    if (tca < 0) return it;                           /* ray.py:81-82 */
    double q = vdot(L, L) - tca * tca, d;
    if (q < 0) d = 0; else d = sqrt(q);               /* ray.py:85-88 (math.sqrt raises -> d = 0) */
    if (d > r) return it;                             /* ray.py:89-90 */
    double thc = sqrt(r * r - d * d);
    double t0 = tca - thc, t1 = tca + thc;
    double tmin = point ? t1 : t0;                    /* ray.py:96: may be negative */
    it.hit = 1; it.t = tmin;
    it.p = vadd(O, vscale(D, tmin));
    it.n = vnorm(vsub(it.p, C));
    return it;
}

/* Ray.sphereExitRay, ray.py:109-157.  Returns 0 when the ray is trapped
   (reference prints and returns None) or when the reference would raise
   (entry TIR for ior<1, degenerate chord): callers treat both as None.       */
static int sphere_exit_ray(v3 D, v3 C, double r, double ior, const isect *in, v3 *eo, v3 *ed) {
    v3 refr;
    if (!vrefract(D, in->n, 1, ior, &refr)) return 0;
    isect ex = sphere_discriminant(in->p, vnorm(refr), C, r, 1);
    if (!ex.hit) return 0;
    v3 exit_d; int done = 0, n = 0;
    while (!done && n < 10) {
        n++;
        if (vrefract(refr, vneg(ex.n), ior, 1, &exit_d)) done = 1;
        else {
            refr = vreflect(refr, ex.n);              /* TIR, ray.py:137 */
            ex = sphere_discriminant(ex.p, vnorm(refr), C, r, 1);
            if (!ex.hit) return 0;
        }
    }
    if (!done) return 0;
    *eo = ex.p; *ed = vnorm(exit_d);                  /* Ray(exit point, exit_D) */
    return 1;
}

/* Nearest signed-distance hit over all spheres whose id != suppress
   (ray.py:162-168 + Intersection.nearestIntersection ray.py:10-20).          */
static __thread uint64_t g_queries;   /* nearest-hit / occlusion queries issued by this thread */
static isect nearest_signed(const orc_scene *s, v3 O, v3 D, int32_t suppress) {
    g_queries++;
    isect best; memset(&best, 0, sizeof best); best.idx = -1;
    for (int i = 0; i < s->n; ++i) {
        if (suppress != ORC_NO_ID && s->ids[i] == suppress) continue;
        isect it = sphere_discriminant(O, D, ld3(s->centre, i), s->radius[i], 0);
        if (it.hit && (!best.hit || it.t < best.t)) { best = it; best.idx = i; }
    }
    return best;
}

/* Ray.nearestSphereIntersect, ray.py:160-231, recursion unrolled: a mirror
   that finds nothing returns ITSELF (ray.py:198-201), glass that finds
   nothing returns None (ray.py:226-229), so the result of a dead-ended chain
   is the most recent mirror hit, else None.  D normalised.                   */
static isect trace_terminal(const orc_scene *s, v3 O, v3 D, int32_t suppress, int bounces, int max_bounces, int through) {
    isect fallback; memset(&fallback, 0, sizeof fallback); fallback.idx = -1;
    for (;;) {
        isect h = nearest_signed(s, O, D, suppress);
        if (!h.hit) return fallback;                  /* ray.py:170-171 */
        if (bounces > max_bounces) return fallback;   /* ray.py:173-174 */
        h.bounces = bounces; h.through = through;
        const double *m = s->material + 4 * h.idx;
        if (m[0] == 1.0) {                            /* reflective == True, ray.py:180 */
            fallback = h;
            D = vnorm(vreflect(D, h.n)); O = h.p;     /* Ray() re-normalises */
            bounces += 1; suppress = s->ids[h.idx];
            continue;
        }
        if (m[1] == 1.0) {                            /* transparent == True, ray.py:204 */
            v3 eo, ed;
            if (!sphere_exit_ray(D, ld3(s->centre, h.idx), s->radius[h.idx], m[3], &h, &eo, &ed)) return fallback;
            /* NB a trapped glass ray returns None to ITS caller; an enclosing mirror
               then returns itself -> identical to "return fallback". */
            O = eo; D = ed; bounces += 1; through += 1; suppress = s->ids[h.idx];
            /* glass yields None if nothing found beyond it, but an enclosing mirror still
               returns itself: fallback unchanged. */
            continue;
        }
        return h;
    }
}

static inline double incidence(double angle, double max_angle) {   /* light.py:3-9 */
    if (angle > max_angle) return 0;
    if (angle == 0) return 1;
    return (max_angle - angle) / max_angle;
}

/* Intersection.terminalRGB, ray.py:37-65 (+ colour.py:21-29 illuminate) */
static void terminal_rgb(const orc_scene *s, const isect *h, int shadow_max_bounces, double out[3]) {
    const double *m = s->material + 4 * h->idx;
    const double *col = s->colour + 3 * h->idx;
    double il[3] = {col[0] * m[2], col[1] * m[2], col[2] * m[2]};       /* ray.py:41 */
    for (int g = 0; g < s->nG; ++g) {                                   /* ray.py:43-45 */
        if (s->g_func[g] != 0) continue;   /* reference returns None -> would raise */
        double ang = vangle(h->n, ld3(s->g_vec, g));
        double sc = incidence(ang, s->g_max_angle[g]) * s->g_strength[g];
        for (int c = 0; c < 3; ++c) il[c] = il[c] + s->g_col[3 * g + c] * sc;
    }
    int32_t own = s->ids[h->idx];
    for (int p = 0; p < s->nP; ++p) {                                   /* ray.py:47-62 */
        if (own == s->p_id[p]) continue;
        v3 vl = vsub(ld3(s->p_pos, p), h->p);
        isect t = trace_terminal(s, h->p, vnorm(vl), own, 0, shadow_max_bounces, 0);
        if (!t.hit || s->ids[t.idx] != s->p_id[p]) continue;
        double ang = vangle(h->n, vl), dist = vmag(vl), sc;
        if (s->p_func[p] == -1) sc = incidence(ang, s->p_max_angle[p]) * s->p_strength[p];
        else if (s->p_func[p] == 0) sc = incidence(ang, s->p_max_angle[p]) * s->p_strength[p] / dist;
        else continue;
        for (int c = 0; c < 3; ++c) il[c] = il[c] + s->p_col[3 * p + c] * sc;
    }
    for (int c = 0; c < 3; ++c) out[c] = s->bg[c] + rint(col[c] * (il[c] / 255));  /* half-to-even */
}

/* ------------------------------------------------------ C-ABI: unit KATs */
ORC_API int orc_sphere_discriminant(const double *O, const double *Draw, const double *C, double r, int point, double *out8) {
    isect it = sphere_discriminant(V(O[0], O[1], O[2]), vnorm(V(Draw[0], Draw[1], Draw[2])), V(C[0], C[1], C[2]), r, point);
    out8[0] = it.t; out8[1] = it.p.x; out8[2] = it.p.y; out8[3] = it.p.z; out8[4] = it.n.x; out8[5] = it.n.y; out8[6] = it.n.z;
    return it.hit;
}
ORC_API void orc_reflect(const double *v, const double *n, double *out) {
    v3 r = vreflect(V(v[0], v[1], v[2]), V(n[0], n[1], n[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
ORC_API int orc_refract(const double *v, const double *n, double ra, double rb, double *out) {
    v3 r; if (!vrefract(V(v[0], v[1], v[2]), V(n[0], n[1], n[2]), ra, rb, &r)) return 0;
    out[0] = r.x; out[1] = r.y; out[2] = r.z; return 1;
}
/* batch of rays through nearestSphereIntersect (+ optional terminalRGB).
   rays [m,6] = origin, raw direction.  term [m,11] = hit, idx, bounces, through, p(3), n(3), distance.
   rgb [m,3] (NULL to skip); misses get miss[3].                              */
ORC_API void orc_trace_rays(const orc_scene *s, int m, const double *rays, const int32_t *suppress, const int32_t *bounces0,
                            int max_bounces, int shadow_max_bounces, const double *miss, double *term, double *rgb) {
    for (int i = 0; i < m; ++i) {
        const double *r = rays + 6 * i;
        isect h = trace_terminal(s, V(r[0], r[1], r[2]), vnorm(V(r[3], r[4], r[5])), suppress ? suppress[i] : ORC_NO_ID,
                                 bounces0 ? bounces0[i] : 0, max_bounces, 0);
        double *t = term + 11 * i;
        t[0] = h.hit; t[1] = h.idx; t[2] = h.bounces; t[3] = h.through;
        t[4] = h.p.x; t[5] = h.p.y; t[6] = h.p.z; t[7] = h.n.x; t[8] = h.n.y; t[9] = h.n.z; t[10] = h.t;
        if (rgb) {
            if (h.hit) terminal_rgb(s, &h, shadow_max_bounces, rgb + 3 * i);
            else for (int c = 0; c < 3; ++c) rgb[3 * i + c] = miss[c];
        }
    }
}

/* Intersection.terminalRGB at given hits: hits [m,7] = scene index, point(3), normal(3) -> rgb [m,3] */
ORC_API void orc_shade_hits(const orc_scene *s, int m, const double *hits, int shadow_max_bounces, double *rgb) {
    for (int i = 0; i < m; ++i) {
        const double *q = hits + 7 * i;
        isect h; memset(&h, 0, sizeof h);
        h.hit = 1; h.idx = (int)q[0]; h.p = V(q[1], q[2], q[3]); h.n = V(q[4], q[5], q[6]);
        terminal_rgb(s, &h, shadow_max_bounces, rgb + 3 * i);
    }
}

/* ------------------------------------------- Algorithm A frame (Whitted) */
/* Drivers: RL/output5.py:416-533 render_true_original (spp 1, no int()),
   RL/output5.py:1420-1525 render_custom_scene('traditional') (jitter iff
   spp>1, :1463-1470; int(sum/spp), :1500-1505), notebooks' cell 0 loops.
   sum_out [H,W,3] = sum over samples of terminalRGB (miss -> miss[3]);
   hit_out [H,W] = terminal scene index of the LAST sample (-1 miss).
   prenorm: render_custom_scene normalises the direction before Ray()
   normalises it again (:1476-1483).                                         */
ORC_API void orc_render_whitted(const orc_scene *s, const double *cam, const double *X, const double *Y, int W, int H,
                                int y0, int y1, int spp, int max_bounces, int shadow_max_bounces, const double *miss,
                                uint64_t seed, int prenorm, double *sum_out, int32_t *hit_out, uint64_t *ray_count,
                                int nthreads) {
    v3 O = V(cam[0], cam[1], cam[2]);
    double pitch_x = W > 1 ? X[1] - X[0] : 0, pitch_y = H > 1 ? Y[0] - Y[1] : 0;
    uint64_t rays = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rays)
#endif
    for (int y = y0; y < y1; ++y)
        for (int x = 0; x < W; ++x) {
            double acc[3] = {0, 0, 0}; int last = -1;
            for (int sm = 0; sm < spp; ++sm) {
                double Xj = X[x], Yj = Y[y];
                if (spp > 1) {
                    double u0, u1; rng_pair(seed, (uint32_t)(y * W + x), (uint32_t)sm, 0, &u0, &u1);
                    Xj = X[x] + (u0 - 0.5) * pitch_x; Yj = Y[y] + (u1 - 0.5) * pitch_y;
                }
                v3 d = V(Xj, Yj, -1);
                if (prenorm) d = vnorm(d);
                g_queries = 0;
                isect h = trace_terminal(s, O, vnorm(d), ORC_NO_ID, 0, max_bounces, 0);
                double c[3];
                if (h.hit) { terminal_rgb(s, &h, shadow_max_bounces, c); last = h.idx; }
                else { c[0] = miss[0]; c[1] = miss[1]; c[2] = miss[2]; last = -1; }
                rays += g_queries;       /* primary + mirror/glass continuation + shadow queries */
                acc[0] += c[0]; acc[1] += c[1]; acc[2] += c[2];
            }
            double *o = sum_out + 3 * ((size_t)y * W + x);
            o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
            if (hit_out) hit_out[(size_t)y * W + x] = last;
        }
    if (ray_count) *ray_count = rays;
}

/* --------------------------------- Algorithm B frame (TraditionalRenderer) */
typedef struct { uint64_t rays, inter, light, small, queries; } bstats;
typedef struct {
    const orc_scene *s; int max_bounces; double thr; uint64_t seed; uint32_t pixel, sample; bstats *st;
} bctx;

/* TraditionalRenderer.trace_ray_traditional: FB/fb_vs_traditional_chandelier.py:431-521
   (== FB/fb_vs_traditional_complex.py:299-389 except the mirror threshold,
   complex :349 "> 0.9" vs chandelier :481 "> 0").                            */
static void trace_path(const bctx *c, v3 O, v3 D, int bounce, double out[3]) {
    const orc_scene *s = c->s;
    c->st->rays++;
    if (bounce >= c->max_bounces) { out[0] = 2; out[1] = 2; out[2] = 5; return; }
    c->st->queries++;                                  /* calls that actually loop over the spheres */
    isect best; memset(&best, 0, sizeof best); best.idx = -1; double nd = INFINITY;
    for (int i = 0; i < s->n; ++i) {
        isect it = sphere_discriminant(O, D, ld3(s->centre, i), s->radius[i], 0);
        if (it.hit) { double dist = vdist(it.p, O); if (dist < nd) { nd = dist; best = it; best.idx = i; } }
    }
    if (!best.hit) { out[0] = 2; out[1] = 2; out[2] = 5; return; }
    c->st->inter++;
    const double *m = s->material + 4 * best.idx, *col = s->colour + 3 * best.idx;
    if (m[2] != 0) {                                   /* if material.emitive */
        c->st->light++; if (s->small && s->small[best.idx]) c->st->small++;
        out[0] = col[0]; out[1] = col[1]; out[2] = col[2]; return;
    }
    double direct[3] = {0, 0, 0};
    for (int l = 0; l < s->nL; ++l) {                  /* no occlusion test */
        if (s->l_index[l] == best.idx) continue;
        v3 tl = vsub(ld3(s->l_centre, l), best.p), tln = vnorm(tl);
        double ca = vdot(best.n, tln); if (!(ca > 0)) ca = 0;
        if (ca > 0) {
            double dist = vmag(tl), att = 1.0 / (dist * dist);
            for (int k = 0; k < 3; ++k) direct[k] += (double)(int64_t)(s->l_colour[3 * l + k] * ca * att * 0.3);
        }
    }
    double ind[3];
    v3 o2 = vadd(best.p, vscale(best.n, 0.001));
    if (m[0] > c->thr) {
        v3 rd = vreflect(D, best.n);
        trace_path(c, o2, vnorm(rd), bounce + 1, ind);
    } else {
        double r1, r2; rng_pair(c->seed, c->pixel, c->sample, (uint32_t)bounce + 1, &r1, &r2);
        double theta = acos(sqrt(r1)), phi = 2 * M_PI * r2;
        v3 tg = fabs(best.n.z) > 0.9 ? V(1, 0, 0) : vcross(V(0, 0, 1), best.n);
        tg = vnorm(tg);
        v3 bt = vnorm(vcross(best.n, tg));
        v3 ld = V(sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta));
        v3 bd = vnorm(V(ld.x * tg.x + ld.y * bt.x + ld.z * best.n.x,
                        ld.x * tg.y + ld.y * bt.y + ld.z * best.n.y,
                        ld.x * tg.z + ld.y * bt.z + ld.z * best.n.z));
        trace_path(c, o2, vnorm(bd), bounce + 1, ind);
    }
    for (int k = 0; k < 3; ++k) {
        double tot = direct[k] + ind[k]; if (!(tot < 255)) tot = 255;   /* min(255, .) */
        out[k] = (double)(int64_t)(col[k] * (tot / 255.0));
    }
}

/* WorkingFBRenderer.trace_ray_fb: FB/fb_vs_traditional_complex.py:487-601 -- trace_ray_traditional with the diffuse
   direction taken from a policy (fb_agent.choose_direction(obs22) -> action in [-1,1]^2) with probability
   fb_usage_prob.  Draws: the use-FB decision is word 0 of the Philox block (pixel, sample, bounce, tag "RTFB"); r1, r2
   stay on slot bounce+1 of the path stream. */
typedef void (*orc_policy_fn)(const float *obs22, float *action2, void *user);
typedef struct {
    const orc_scene *s; int max_bounces; double thr, fb_prob; uint64_t seed; uint32_t pixel, sample; bstats *st;
    uint64_t *fb_used; orc_policy_fn policy; void *user;
} fctx;
static void trace_path_fb(const fctx *c, v3 O, v3 D, int bounce, double out[3]) {
    const orc_scene *s = c->s;
    c->st->rays++;
    if (bounce >= c->max_bounces) { out[0] = 2; out[1] = 2; out[2] = 5; return; }
    c->st->queries++;
    isect best; memset(&best, 0, sizeof best); best.idx = -1; double nd = INFINITY;
    for (int i = 0; i < s->n; ++i) {
        isect it = sphere_discriminant(O, D, ld3(s->centre, i), s->radius[i], 0);
        if (it.hit) { double dist = vdist(it.p, O); if (dist < nd) { nd = dist; best = it; best.idx = i; } }
    }
    if (!best.hit) { out[0] = 2; out[1] = 2; out[2] = 5; return; }
    c->st->inter++;
    const double *m = s->material + 4 * best.idx, *col = s->colour + 3 * best.idx;
    if (m[2] != 0) {
        c->st->light++; if (s->small && s->small[best.idx]) c->st->small++;
        out[0] = col[0]; out[1] = col[1]; out[2] = col[2]; return;
    }
    double direct[3] = {0, 0, 0};
    for (int l = 0; l < s->nL; ++l) {
        if (s->l_index[l] == best.idx) continue;
        v3 tl = vsub(ld3(s->l_centre, l), best.p), tln = vnorm(tl);
        double ca = vdot(best.n, tln); if (!(ca > 0)) ca = 0;
        if (ca > 0) {
            double dist = vmag(tl), att = 1.0 / (dist * dist);
            for (int k = 0; k < 3; ++k) direct[k] += (double)(int64_t)(s->l_colour[3 * l + k] * ca * att * 0.3);
        }
    }
    double ind[3];
    v3 o2 = vadd(best.p, vscale(best.n, 0.001));
    if (m[0] > c->thr) {
        trace_path_fb(c, o2, vnorm(vreflect(D, best.n)), bounce + 1, ind);
    } else {
        int use_fb = 0;
        if (c->policy) {                                     /* self.fb_loaded and np.random.random() < fb_usage_prob, :537 */
            uint32_t ctr[4] = {c->pixel, c->sample, (uint32_t)bounce, 0x52544642u /* "RTFB" */};
            uint32_t key[2] = {(uint32_t)c->seed, (uint32_t)(c->seed >> 32)}, o[4];
            philox4x32_10(ctr, key, o);
            use_fb = (double)(o[0] >> 8) * (1.0 / 16777216.0) < c->fb_prob;
        }
        double st_, ct_, sp_, cp_;
        if (use_fb) {
            (*c->fb_used)++;
            float obs[22], act[2];                           /* create_observation, :469-485 (accumulated_color is always 0) */
            obs[0] = (float)best.p.x; obs[1] = (float)best.p.y; obs[2] = (float)best.p.z;
            obs[3] = (float)D.x; obs[4] = (float)D.y; obs[5] = (float)D.z;
            obs[6] = (float)best.n.x; obs[7] = (float)best.n.y; obs[8] = (float)best.n.z;
            obs[9] = (float)m[0]; obs[10] = (float)m[1]; obs[11] = (float)m[2]; obs[12] = (float)m[3];
            obs[13] = obs[14] = obs[15] = 0.0f;
            obs[16] = (float)((double)bounce / c->max_bounces); obs[17] = 0.0f;
            obs[18] = (float)((double)s->ids[best.idx] / 100.0); obs[19] = obs[20] = obs[21] = 0.5f;
            c->policy(obs, act, c->user);
            /* :545-546 in double (the action's float32 values promoted): what the reference computes when the agent
               returns float64, and under NumPy < 2 for float32 too; with NumPy >= 2 a float32 action would drag the
               rest of the branch into float32 */
            double theta = ((double)act[0] + 1) * M_PI / 4, phi = (double)act[1] * M_PI;
            st_ = sin(theta); ct_ = cos(theta); sp_ = sin(phi); cp_ = cos(phi);
        } else {
            double r1, r2; rng_pair(c->seed, c->pixel, c->sample, (uint32_t)bounce + 1, &r1, &r2);
            double theta = acos(sqrt(r1)), phi = 2 * M_PI * r2;
            st_ = sin(theta); ct_ = cos(theta); sp_ = sin(phi); cp_ = cos(phi);
        }
        v3 tg = fabs(best.n.z) > 0.9 ? V(1, 0, 0) : vcross(V(0, 0, 1), best.n);
        tg = vnorm(tg);
        v3 bt = vnorm(vcross(best.n, tg));
        v3 ld = V(st_ * cp_, st_ * sp_, ct_);
        v3 bd = vnorm(V(ld.x * tg.x + ld.y * bt.x + ld.z * best.n.x, ld.x * tg.y + ld.y * bt.y + ld.z * best.n.y,
                        ld.x * tg.z + ld.y * bt.z + ld.z * best.n.z));
        trace_path_fb(c, o2, vnorm(bd), bounce + 1, ind);
    }
    for (int k = 0; k < 3; ++k) {
        double tot = direct[k] + ind[k]; if (!(tot < 255)) tot = 255;
        out[k] = (double)(int64_t)(col[k] * (tot / 255.0));
    }
}

/* WorkingFBRenderer.render: FB/fb_vs_traditional_complex.py:603-640 (same camera and resolve as TraditionalRenderer).
   stats6 = total_rays, total_intersections, light_hits, small_light_hits, queries, fb_used.  Single-threaded: the
   policy is a Python callback. */
ORC_API void orc_render_path_fb(const orc_scene *s, const double *cam, int W, int H, double fov_deg, int s0, int s1,
                                int max_bounces, double mirror_threshold, double fb_prob, uint64_t seed,
                                orc_policy_fn policy, void *user, double *sum_out, uint64_t *stats6) {
    v3 O = V(cam[0], cam[1], cam[2]);
    double aspect = (double)W / (double)H;
    double half_h = tan((fov_deg * (M_PI / 180.0)) / 2), half_w = half_h * aspect;
    bstats st = {0, 0, 0, 0, 0};
    uint64_t fb_used = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double acc[3] = {0, 0, 0};
            for (int sm = s0; sm < s1; ++sm) {
                uint32_t pix = (uint32_t)(y * W + x);
                double u0, u1; rng_pair(seed, pix, (uint32_t)sm, 0, &u0, &u1);
                double sx = 0.5 + (u0 - 0.5), sy = 0.5 + (u1 - 0.5);
                double ndc_x = (x + sx) / W, ndc_y = (y + sy) / H;
                double scx = 2.0 * ndc_x - 1.0, scy = 1.0 - 2.0 * ndc_y;
                scx *= aspect; scx *= half_w; scy *= half_h;
                v3 d = vnorm(V(scx, scy, -1));
                fctx c = {s, max_bounces, mirror_threshold, fb_prob, seed, pix, (uint32_t)sm, &st, &fb_used, policy, user};
                double col[3];
                trace_path_fb(&c, O, vnorm(d), 0, col);
                acc[0] += col[0]; acc[1] += col[1]; acc[2] += col[2];
            }
            double *o = sum_out + 3 * ((size_t)y * W + x);
            o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
        }
    if (stats6) { stats6[0] = st.rays; stats6[1] = st.inter; stats6[2] = st.light; stats6[3] = st.small; stats6[4] = st.queries; stats6[5] = fb_used; }
}

/* TraditionalRenderer.generate_camera_ray + render:
   FB/fb_vs_traditional_chandelier.py:417-429, :523-554.  sum_out [H,W,3] =
   sum over samples [s0,s1) of the per-sample colour for rows [y0,y1).       */
ORC_API void orc_render_path(const orc_scene *s, const double *cam, int W, int H, double fov_deg, int y0, int y1,
                             int s0, int s1, int max_bounces, double mirror_threshold, uint64_t seed,
                             double *sum_out, uint64_t *stats5, int nthreads) {
    v3 O = V(cam[0], cam[1], cam[2]);
    double aspect = (double)W / (double)H;
    double half_h = tan((fov_deg * (M_PI / 180.0)) / 2), half_w = half_h * aspect;
    uint64_t R = 0, I = 0, Lh = 0, Sh = 0, Q = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : R, I, Lh, Sh, Q)
#endif
    for (int y = y0; y < y1; ++y) {
        bstats st = {0, 0, 0, 0, 0};
        for (int x = 0; x < W; ++x) {
            double acc[3] = {0, 0, 0};
            for (int sm = s0; sm < s1; ++sm) {
                uint32_t pix = (uint32_t)(y * W + x);
                double u0, u1; rng_pair(seed, pix, (uint32_t)sm, 0, &u0, &u1);
                double sx = 0.5 + (u0 - 0.5), sy = 0.5 + (u1 - 0.5);
                double ndc_x = (x + sx) / W, ndc_y = (y + sy) / H;
                double scx = 2.0 * ndc_x - 1.0, scy = 1.0 - 2.0 * ndc_y;
                scx *= aspect; scx *= half_w; scy *= half_h;        /* aspect applied twice */
                v3 d = vnorm(V(scx, scy, -1));
                bctx c = {s, max_bounces, mirror_threshold, seed, pix, (uint32_t)sm, &st};
                double col[3];
                trace_path(&c, O, vnorm(d), 0, col);
                acc[0] += col[0]; acc[1] += col[1]; acc[2] += col[2];
            }
            double *o = sum_out + 3 * ((size_t)y * W + x);
            o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
        }
        R += st.rays; I += st.inter; Lh += st.light; Sh += st.small; Q += st.queries;
    }
    if (stats5) { stats5[0] = R; stats5[1] = I; stats5[2] = Lh; stats5[3] = Sh; stats5[4] = Q; }
}

/* ------------------- "Algorithm C": FB/output6.py SimplifiedFBRenderer (fb_usage_prob = 0) */
typedef struct {
    double cam[3];                 /* render_original_style: Vector(0, 0, 1), output6.py:605 */
    double fov;                    /* radians: np.pi / 3, output6.py:622 */
    double sun_pos[3];             /* self.sun_position, output6.py:99 */
    double sun_col[3];             /* self.sun_color, output6.py:101 */
    int32_t sun_id;                /* 7, output6.py:204,478 */
    int32_t max_bounces;           /* 5, output6.py:113 */
} orc_simple_cfg;

/* SimplifiedFBRenderer.calculate_lighting_exact_original: FB/output6.py:197-306 */
static void simple_lighting(const orc_scene *s, const orc_simple_cfg *c, const isect *h, int64_t out[3], uint64_t *sun_hits) {
    if (s->ids[h->idx] == c->sun_id) {                                        /* :204-206 */
        (*sun_hits)++;
        for (int k = 0; k < 3; ++k) out[k] = (int64_t)c->sun_col[k];
        return;
    }
    v3 sun = V(c->sun_pos[0], c->sun_pos[1], c->sun_pos[2]);
    v3 to_sun = vnorm(vsub(sun, h->p));                                       /* :244 */
    v3 gdir = vnorm(V(3, 1, -0.75));                                          /* :247 */
    double gcos = vdot(h->n, gdir); if (!(gcos > 0)) gcos = 0;                /* :248 max(0, .) */
    const double gl[3] = {20, 20, 255};
    int64_t g[3], su[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) g[k] = (int64_t)(gl[k] * gcos * 0.3);         /* :250-254 */
    v3 so = vadd(h->p, vscale(h->n, 0.001)), sd = vnorm(to_sun);              /* :258-261, Ray() normalises */
    double sun_distance = vdist(h->p, sun);                                   /* :264 */
    int visible = 1;
    for (int i = 0; i < s->n; ++i) {                                          /* :266-275 */
        if (i == h->idx || s->ids[i] == c->sun_id) continue;
        isect it = sphere_discriminant(so, sd, ld3(s->centre, i), s->radius[i], 0);
        if (it.hit && vdist(it.p, h->p) < sun_distance) { visible = 0; break; }
    }
    if (visible) {                                                            /* :278-291 */
        double distance = vdist(h->p, sun);
        double att = distance > 0 ? 1.0 / (distance * distance) : 1.0;
        att = att * 100 < 1.0 ? att * 100 : 1.0;
        double ca = vdot(h->n, to_sun); if (!(ca > 0)) ca = 0;
        for (int k = 0; k < 3; ++k) su[k] = (int64_t)(c->sun_col[k] * ca * att * 0.9);
    }
    const double *col = s->colour + 3 * h->idx;
    for (int k = 0; k < 3; ++k) {                                             /* :293-304 */
        int64_t comb = g[k] + su[k]; if (comb > 255) comb = 255;
        out[k] = (int64_t)(col[k] * ((double)comb / 255.0));
    }
}

/* SimplifiedFBRenderer.trace_ray_simple: FB/output6.py:434-577.  Philox slot bounce+1: glass draws word 0,
   diffuse draws words 0,1 (the reference draws from np.random.random in that order). */
static void trace_simple(const orc_scene *s, const orc_simple_cfg *c, v3 O, v3 D, uint64_t seed, uint32_t pixel,
                         int64_t acc[3], uint64_t *rays, uint64_t *sun_hits) {
    acc[0] = acc[1] = acc[2] = 0;
    int bounce = 0;
    while (bounce < c->max_bounces) {
        (*rays)++;
        isect best; memset(&best, 0, sizeof best); best.idx = -1; double nd = INFINITY;
        for (int i = 0; i < s->n; ++i) {                                      /* :451-457 */
            isect it = sphere_discriminant(O, D, ld3(s->centre, i), s->radius[i], 0);
            if (it.hit) { double dist = vdist(it.p, O); if (dist < nd) { nd = dist; best = it; best.idx = i; } }
        }
        if (!best.hit) { if (bounce == 0) { acc[0] = 2; acc[1] = 2; acc[2] = 5; } break; }   /* :459-463 */
        int64_t li[3];
        simple_lighting(s, c, &best, li, sun_hits);
        for (int k = 0; k < 3; ++k) { acc[k] += li[k]; if (acc[k] > 255) acc[k] = 255; }     /* :472-476 */
        if (s->ids[best.idx] == c->sun_id) break;                             /* :479-480 */
        const double *m = s->material + 4 * best.idx;
        v3 nd_;
        if (m[0] != 0) nd_ = vreflect(D, best.n);                             /* truthy reflective, :485-487 */
        else if (m[1] != 0) {                                                 /* truthy transparent, :489-494 */
            double u0, u1; rng_pair(seed, pixel, 0, (uint32_t)bounce + 1, &u0, &u1);
            nd_ = u0 < 0.5 ? vreflect(D, best.n) : D;
        } else {                                                              /* :540-564 */
            double r1, r2; rng_pair(seed, pixel, 0, (uint32_t)bounce + 1, &r1, &r2);
            double theta = acos(sqrt(r1)), phi = 2 * M_PI * r2;
            v3 tg = fabs(best.n.z) > 0.9 ? V(1, 0, 0) : vcross(V(0, 0, 1), best.n);
            tg = vnorm(tg);
            v3 bt = vnorm(vcross(best.n, tg));
            v3 ld = V(sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta));
            nd_ = vnorm(V(ld.x * tg.x + ld.y * bt.x + ld.z * best.n.x, ld.x * tg.y + ld.y * bt.y + ld.z * best.n.y,
                          ld.x * tg.z + ld.y * bt.z + ld.z * best.n.z));
        }
        O = vadd(best.p, vscale(best.n, 0.001));                              /* :567-570 */
        D = vnorm(nd_);
        bounce++;
    }
}

/* render_original_style: FB/output6.py:579-635.  rgb_out [H,W,3] integer-valued colours (image = min(1, c/255));
   rays != NULL: m explicit rays [m,6] (origin + raw direction) instead of the camera grid, W = m, H = 1.
   stats2 = total_rays, sun_hits. */
/* calculate_lighting_exact_original on m given intersections: hits [m,7] = point, normal, scene index -> out [m,3];
   st[1] += sun hits */
ORC_API void orc_simple_lighting(const orc_scene *s, const orc_simple_cfg *c, int m, const double *hits, double *out,
                                 uint64_t *st) {
    uint64_t sh = 0;
    for (int i = 0; i < m; ++i) {
        const double *r = hits + 7 * (size_t)i;
        isect h;
        memset(&h, 0, sizeof h);
        h.hit = 1; h.idx = (int)r[6];
        h.p = V(r[0], r[1], r[2]); h.n = V(r[3], r[4], r[5]);
        int64_t li[3] = {0, 0, 0};
        if (h.idx >= 0 && h.idx < s->n) simple_lighting(s, c, &h, li, &sh);
        for (int k = 0; k < 3; ++k) out[3 * (size_t)i + k] = (double)li[k];
    }
    if (st) st[1] += sh;
}

ORC_API void orc_render_simple(const orc_scene *s, const orc_simple_cfg *c, int W, int H, uint64_t seed,
                               const double *rays, double *rgb_out, uint64_t *stats2, int nthreads) {
    uint64_t R = 0, S = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : R, S)
#endif
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            v3 O, d;
            size_t i = (size_t)y * W + x;
            if (rays) { O = V(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]); d = V(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]); }
            else {
                double u = ((double)x / W - 0.5) * 2.0, v = ((double)y / H - 0.5) * -2.0;   /* :612-613 */
                double aspect = (double)W / (double)H;
                u *= aspect;                                                                 /* :616-617 */
                double t = tan(c->fov / 2);
                d = vnorm(V(u * t, v * t, -1));                                              /* :621 */
                O = V(c->cam[0], c->cam[1], c->cam[2]);
            }
            int64_t acc[3];
            uint64_t r = 0, sh = 0;
            trace_simple(s, c, O, vnorm(d), seed, (uint32_t)i, acc, &r, &sh);
            R += r; S += sh;
            for (int k = 0; k < 3; ++k) rgb_out[3 * i + k] = (double)acc[k];
        }
    if (stats2) { stats2[0] = R; stats2[1] = S; }
}

/* ---------------- FB training trajectories: FB/train_complex_only.py:54-162, :254-334 */
/* sample_cosine_weighted_direction(normal), train_complex_only.py:69-96 (tangent rule |n.z| < 0.999, unlike the
   renderers' |n.z| > 0.9) */
static v3 traj_basis_dir(v3 n, double lx, double ly, double lz, v3 *tg_out, v3 *bt_out) {
    v3 tg = fabs(n.z) < 0.999 ? vnorm(vcross(V(0, 0, 1), n)) : vnorm(vcross(V(1, 0, 0), n));
    v3 bt = vnorm(vcross(n, tg));
    if (tg_out) { *tg_out = tg; *bt_out = bt; }
    return vnorm(V(lx * tg.x + ly * bt.x + lz * n.x, lx * tg.y + ly * bt.y + lz * n.y, lx * tg.z + ly * bt.z + lz * n.z));
}
static v3 traj_cosine_dir(v3 n, double r1, double r2) {
    double theta = acos(sqrt(r1)), phi = 2 * M_PI * r2;
    return traj_basis_dir(n, sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta), NULL, NULL);
}
/* create_observation, train_complex_only.py:130-150 (22 x float32) */
static void traj_obs(const orc_scene *s, v3 p, v3 n, v3 d, int bounce, const double col[3], int idx, int max_bounces, float *o) {
    const double *m = s->material + 4 * idx;
    o[0] = (float)p.x; o[1] = (float)p.y; o[2] = (float)p.z; o[3] = (float)d.x; o[4] = (float)d.y; o[5] = (float)d.z;
    o[6] = (float)n.x; o[7] = (float)n.y; o[8] = (float)n.z;
    o[9] = (float)m[0]; o[10] = (float)m[1]; o[11] = (float)m[2]; o[12] = (float)m[3];
    o[13] = (float)(col[0] / 255.0); o[14] = (float)(col[1] / 255.0); o[15] = (float)(col[2] / 255.0);
    o[16] = (float)((double)bounce / max_bounces); o[17] = 0.0f; o[18] = (float)((double)s->ids[idx] / 100.0);
    o[19] = 0.5f; o[20] = 0.5f; o[21] = 0.5f;
}
/* RayTracedComplexTrainer.generate_trajectory, train_complex_only.py:254-334.  Draws (Philox, trajectory j, sample 0):
   slot 0 = (sphere choice, theta), slot 1 word 0 = phi, slot 2 = incoming direction (r1, r2), slot 3+k = step k.
   obs/next_obs [n,max_steps,22] f32, action [n,max_steps,2] f32, reward [n,max_steps] f32, hit [n,max_steps] u8,
   length [n] (transitions recorded), hit_light [n]. */
ORC_API void orc_generate_trajectories(const orc_scene *s, int n_traj, int max_steps, int max_bounces, uint64_t seed,
                                       float *obs, float *action, float *next_obs, float *reward, uint8_t *hit,
                                       int32_t *length, uint8_t *hit_light) {
    int nl = 0;
    for (int i = 0; i < s->n; ++i) if (s->material[4 * i + 2] == 0) nl++;
    for (int j = 0; j < n_traj; ++j) {
        length[j] = 0; hit_light[j] = 0;
        if (nl == 0) continue;
        double u0, u1, u2, u3;
        rng_pair(seed, (uint32_t)j, 0, 0, &u0, &u1);
        rng_pair(seed, (uint32_t)j, 0, 1, &u2, &u3);
        int pick = (int)(u0 * nl); if (pick > nl - 1) pick = nl - 1;           /* random.choice(non_light) */
        int idx = -1;
        for (int i = 0; i < s->n; ++i) if (s->material[4 * i + 2] == 0 && pick-- == 0) { idx = i; break; }
        double th = 0 + (2 * M_PI - 0) * u1, ph = 0 + (M_PI - 0) * u2;           /* random.uniform */
        v3 unit = V(sin(ph) * cos(th), sin(ph) * sin(th), cos(ph));
        v3 off = vscale(unit, s->radius[idx]);                                   /* scaleByLength */
        v3 p = vadd(ld3(s->centre, idx), off), n = vnorm(off);
        double r1, r2;
        rng_pair(seed, (uint32_t)j, 0, 2, &r1, &r2);
        v3 din = traj_cosine_dir(n, r1, r2);
        const double black[3] = {0, 0, 0};
        float cur[22];
        traj_obs(s, p, n, din, 0, black, idx, max_bounces, cur);
        int bounce = 0;
        while (bounce < max_steps) {
            rng_pair(seed, (uint32_t)j, 0, 3 + (uint32_t)bounce, &r1, &r2);
            v3 nd = traj_cosine_dir(n, r1, r2);
            /* direction_to_action, :99-127 */
            v3 tg, bt;
            traj_basis_dir(n, 0, 0, 1, &tg, &bt);
            double lx = vdot(nd, tg), ly = vdot(nd, bt), lz = vdot(nd, n);
            double cz = lz > 1 ? 1 : lz; if (cz < -1) cz = -1;
            double theta = acos(cz); if (theta > M_PI / 2) theta = M_PI / 2;
            double phi = atan2(ly, lx);
            float a0 = (float)((theta / (M_PI / 2)) * 2 - 1), a1 = (float)(phi / M_PI);
            v3 ro = vadd(p, vscale(n, 0.001)), rd = vnorm(nd);
            isect best; memset(&best, 0, sizeof best); best.idx = -1; double bd = INFINITY;
            for (int i = 0; i < s->n; ++i) {                                      /* nearest_intersection, :153-166 */
                if (s->ids[i] == s->ids[idx]) continue;
                isect it = sphere_discriminant(ro, rd, ld3(s->centre, i), s->radius[i], 0);
                if (it.hit) { double dist = vdist(it.p, ro); if (dist < bd) { bd = dist; best = it; best.idx = i; } }
            }
            if (!best.hit) break;                                                /* ray escaped */
            size_t t = (size_t)j * max_steps + length[j];
            int lit = s->material[4 * best.idx + 2] != 0;
            memcpy(obs + 22 * t, cur, sizeof cur);
            action[2 * t] = a0; action[2 * t + 1] = a1;
            traj_obs(s, best.p, best.n, rd, bounce + 1, lit ? s->colour + 3 * best.idx : black, best.idx, max_bounces, next_obs + 22 * t);
            reward[t] = lit ? 1.0f : 0.0f; hit[t] = (uint8_t)lit;
            length[j]++;
            if (lit) { hit_light[j] = 1; break; }
            p = best.p; n = best.n; idx = best.idx;
            memcpy(cur, next_obs + 22 * t, sizeof cur);
            bounce++;
        }
    }
}

/* -------------------------------------------------- RayTracerEnv (batched) */
/* One record per env; mirrors the attributes of RayTracerEnv
   (RL/ray_tracer_env.py:79-86).                                              */
typedef struct {
    int32_t has_hit, idx, bounce_count, through_count;
    double p[3], n[3], d[3];      /* current_intersection.point/.normal, current_ray.D */
    double acc[3];                /* accumulated_color */
    double total_reward;
    int32_t consec, total_hits;   /* AdaptiveRewardRayTracerEnv.consecutive_light_hits / total_light_hits */
} orc_env;

typedef struct {
    int32_t W, H, max_bounces, flavour;   /* flavour 0 = RL/ray_tracer_env.py, 1 = FB/ray_tracer_env.py */
    double cam[3], cam_angle[3], fov;
    int32_t sun_id;                        /* FB flavour: hard-coded 7 (FB/ray_tracer_env.py:256,419,460) */
    int32_t adaptive;                      /* 1: AdaptiveRewardRayTracerEnv (RL/train_raytracer_optimized.py:16-67) */
    int32_t light_ids[2];                  /* self.light_ids = [99, 100] (:21) */
} orc_env_cfg;

enum { R_NONE = 0, R_RAY_MISSED = 1, R_RAY_ESCAPED = 2, R_MAX_BOUNCES = 3, R_HIT_SUN = 4, R_ALREADY_ON_SUN = 5 };

static void env_obs(const orc_scene *s, const orc_env *e, float *obs) {   /* RL/ray_tracer_env.py:184-222 */
    if (!e->has_hit) { for (int i = 0; i < 18; ++i) obs[i] = 0.f; return; }
    const double *m = s->material + 4 * e->idx;
    obs[0] = (float)e->p[0]; obs[1] = (float)e->p[1]; obs[2] = (float)e->p[2];
    obs[3] = (float)e->d[0]; obs[4] = (float)e->d[1]; obs[5] = (float)e->d[2];
    obs[6] = (float)e->n[0]; obs[7] = (float)e->n[1]; obs[8] = (float)e->n[2];
    obs[9] = (float)m[0]; obs[10] = (float)m[1]; obs[11] = (float)m[2]; obs[12] = (float)m[3];
    for (int c = 0; c < 3; ++c) obs[13 + c] = (float)(e->acc[c] / 255.0);
    obs[16] = (float)e->bounce_count; obs[17] = (float)e->through_count;
}

static isect env_isect(const orc_env *e) {
    isect h; memset(&h, 0, sizeof h);
    h.hit = e->has_hit; h.idx = e->idx; h.p = V(e->p[0], e->p[1], e->p[2]); h.n = V(e->n[0], e->n[1], e->n[2]);
    return h;
}

/* RL: _calculate_reward RL/ray_tracer_env.py:224-252; FB: FB/ray_tracer_env.py:241-278 */
static double env_reward(const orc_scene *s, const orc_env_cfg *cfg, const isect *h, int bounce_count) {
    if (!h->hit) return -0.1;
    if (cfg->flavour == 1 && s->ids[h->idx] == cfg->sun_id) return 10.0;
    double c[3]; terminal_rgb(s, h, 0, c);
    double brightness = (c[0] + c[1] + c[2]) / (3 * 255);
    double pen = -0.01 * bounce_count;
    return brightness + pen;
}

/* AdaptiveRewardRayTracerEnv._calculate_reward: RL/train_raytracer_optimized.py:25-61 */
static double env_reward_adaptive(const orc_scene *s, const orc_env_cfg *cfg, orc_env *e, const isect *h, int bounce_count) {
    if (!h->hit) return -0.5;
    double base = env_reward(s, cfg, h, bounce_count);
    double light_bonus = 0, reflective_bonus = 0, path_length_penalty = 0;
    int id = s->ids[h->idx];
    if (id == cfg->light_ids[0] || id == cfg->light_ids[1]) {
        light_bonus = 2.0; e->consec += 1; e->total_hits += 1;
        if (e->consec > 1) light_bonus += 0.5 * e->consec;
    } else e->consec = 0;
    if (s->material[4 * h->idx] > 0.5) reflective_bonus = 0.3;
    if (bounce_count < 2 && base > 0) path_length_penalty = -0.1;
    return base + light_bonus + reflective_bonus + path_length_penalty;
}
static double env_reward_rl(const orc_scene *s, const orc_env_cfg *cfg, orc_env *e, const isect *h, int bounce_count) {
    return cfg->adaptive ? env_reward_adaptive(s, cfg, e, h, bounce_count) : env_reward(s, cfg, h, bounce_count);
}

/* FB/ray_tracer_env.py:280-336 */
static double env_lighting_reward(const orc_scene *s, const orc_env_cfg *cfg, const isect *h) {
    if (!h->hit) return 0.0;
    if (s->material[4 * h->idx + 2] != 0) return 0.0;
    int sun = -1;
    for (int i = 0; i < s->n; ++i) if (s->ids[i] == cfg->sun_id) { sun = i; break; }
    if (sun < 0) return 0.1;
    v3 sc = ld3(s->centre, sun);
    v3 to_sun = vnorm(vsub(sc, h->p));
    double ca = vdot(h->n, to_sun); if (!(ca > 0)) ca = 0;
    v3 so = vadd(h->p, vscale(h->n, 0.001)), sd = vnorm(to_sun);
    double sun_dist = vmag(vsub(sc, h->p));
    int shadow = 0;
    for (int i = 0; i < s->n; ++i) {
        if (i == h->idx || s->ids[i] == cfg->sun_id) continue;
        isect it = sphere_discriminant(so, sd, ld3(s->centre, i), s->radius[i], 0);
        if (it.hit && vmag(vsub(it.p, h->p)) < sun_dist) { shadow = 1; break; }
    }
    return shadow ? 0.3 : 0.3 + 0.7 * ca;
}

/* reset: RL/ray_tracer_env.py:254-293 (+ _get_initial_ray :121-142) */
ORC_API void orc_env_reset(const orc_scene *s, const orc_env_cfg *cfg, int B, const int32_t *pixels /*[B,2]*/,
                           orc_env *env, float *obs /*[B,18]*/) {
    for (int b = 0; b < B; ++b) {
        orc_env *e = env + b;
        int32_t keep = e->total_hits;                /* total_light_hits survives reset (train_raytracer_optimized.py:63-66) */
        memset(e, 0, sizeof *e);
        e->total_hits = keep;
        double aspect = (double)cfg->W / (double)cfg->H;
        double fr = cfg->fov * M_PI / 180;
        double px = (2 * (pixels[2 * b] + 0.5) / cfg->W - 1) * aspect * tan(fr / 2);
        double py = (1 - 2 * (pixels[2 * b + 1] + 0.5) / cfg->H) * tan(fr / 2);
        v3 d = vnorm(V(px, py, -1));
        if (cfg->cam_angle[0] != 0 || cfg->cam_angle[1] != 0 || cfg->cam_angle[2] != 0)
            d = vrotate(d, V(cfg->cam_angle[0], cfg->cam_angle[1], cfg->cam_angle[2]));
        d = vnorm(d);
        isect h = trace_terminal(s, V(cfg->cam[0], cfg->cam[1], cfg->cam[2]), d, ORC_NO_ID, 0, cfg->max_bounces, 0);
        e->has_hit = h.hit; e->idx = h.idx;
        e->p[0] = h.p.x; e->p[1] = h.p.y; e->p[2] = h.p.z; e->n[0] = h.n.x; e->n[1] = h.n.y; e->n[2] = h.n.z;
        e->d[0] = d.x; e->d[1] = d.y; e->d[2] = d.z;
        env_obs(s, e, obs + 18 * b);
    }
}

/* step: RL/ray_tracer_env.py:295-401, FB/ray_tracer_env.py:378-514.
   actions [B,2] f32 (as the agents hand them over; promoted to double).      */
ORC_API void orc_env_step(const orc_scene *s, const orc_env_cfg *cfg, int B, const float *actions, orc_env *env,
                          float *obs, double *reward, uint8_t *terminated, uint8_t *truncated, int32_t *reason) {
    for (int b = 0; b < B; ++b) {
        orc_env *e = env + b;
        isect cur = env_isect(e);
        reason[b] = R_NONE; terminated[b] = 0; truncated[b] = 0;
        if (!e->has_hit) {                                           /* :313-323 */
            reason[b] = R_RAY_MISSED; reward[b] = -1.0; terminated[b] = 1;
            env_obs(s, e, obs + 18 * b); continue;
        }
        if (e->bounce_count >= cfg->max_bounces) {                   /* :325-337 */
            double fr = cfg->flavour == 1 ? env_lighting_reward(s, cfg, &cur) : env_reward_rl(s, cfg, e, &cur, e->bounce_count);
            e->total_reward += fr; reason[b] = R_MAX_BOUNCES; reward[b] = fr; terminated[b] = 1; truncated[b] = 1;
            env_obs(s, e, obs + 18 * b); continue;
        }
        if (cfg->flavour == 1 && s->ids[e->idx] == cfg->sun_id) {    /* FB :417-431 (total_reward NOT updated) */
            reason[b] = R_ALREADY_ON_SUN; reward[b] = 10.0; terminated[b] = 1;
            env_obs(s, e, obs + 18 * b); continue;
        }
        /* _action_to_direction: RL :144-182, FB :157-198 */
        double theta, phi;
        if (cfg->flavour == 1) { theta = ((double)actions[2 * b] + 1) * M_PI / 4; phi = (double)actions[2 * b + 1] * M_PI; }
        else { theta = (double)actions[2 * b]; phi = (double)actions[2 * b + 1]; }
        v3 ld = V(sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta));
        v3 n = cur.n, tg;
        if (fabs(n.z) < 0.9) tg = vnorm(vcross(V(0, 0, 1), n)); else tg = vnorm(vcross(V(1, 0, 0), n));
        v3 bt = vnorm(vcross(n, tg));
        v3 wd = vnorm(V(ld.x * tg.x + ld.y * bt.x + ld.z * n.x, ld.x * tg.y + ld.y * bt.y + ld.z * n.y,
                        ld.x * tg.z + ld.y * bt.z + ld.z * n.z));
        v3 D = vnorm(wd);                                            /* Ray() normalises again */
        e->bounce_count += 1;
        isect nx = trace_terminal(s, cur.p, D, s->ids[e->idx], e->bounce_count, cfg->max_bounces, e->through_count);
        double rw; int term = 0;
        if (cfg->flavour == 0) rw = env_reward_rl(s, cfg, e, &cur, e->bounce_count);   /* reward at the PRE-update hit, :362 */
        else if (nx.hit) {
            if (s->ids[nx.idx] == cfg->sun_id) { rw = 10.0; reason[b] = R_HIT_SUN; term = 1; }
            else rw = env_lighting_reward(s, cfg, &nx);
        } else { rw = -0.1; reason[b] = R_RAY_MISSED; term = 1; }
        e->total_reward += rw;
        e->d[0] = D.x; e->d[1] = D.y; e->d[2] = D.z;
        e->has_hit = nx.hit; e->idx = nx.idx;
        e->p[0] = nx.p.x; e->p[1] = nx.p.y; e->p[2] = nx.p.z; e->n[0] = nx.n.x; e->n[1] = nx.n.y; e->n[2] = nx.n.z;
        if (nx.hit) {                                                /* :373-381 */
            double c[3]; terminal_rgb(s, &nx, 0, c);
            for (int k = 0; k < 3; ++k) e->acc[k] = e->acc[k] + c[k];
        }
        int trunc = 0;
        if (cfg->flavour == 0) {
            if (!nx.hit) { term = 1; reason[b] = R_RAY_ESCAPED; }
            else if (e->bounce_count >= cfg->max_bounces) { term = 1; trunc = 1; reason[b] = R_MAX_BOUNCES; }
        } else if (!term && e->bounce_count >= cfg->max_bounces) { term = 1; trunc = 1; reason[b] = R_MAX_BOUNCES; }
        reward[b] = rw; terminated[b] = (uint8_t)term; truncated[b] = (uint8_t)trunc;
        env_obs(s, e, obs + 18 * b);
    }
}

ORC_API int orc_sizeof_env(void) { return (int)sizeof(orc_env); }
ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
