/*
 * oracle/orc_pt.inl -- CPU ORACLE (TEST INFRASTRUCTURE) for SURVEY.md section 8 row a14: Mitsuba 3's `path`
 * integrator as /root/reference/scenes/cbox.xml:5-32 configures it.  There is NO reference code for this row:
 * it lives entirely in the un-vendored `mitsuba` wheel (src/integrators/path.cpp, bsdfs/{diffuse,dielectric,
 * conductor}.cpp, emitters/area.cpp, sensors/perspective.cpp, core/{warp,rfilter}), restated here from memory of
 * the Mitsuba 3 sources.  PARITY UNPINNED.  Included by orc_impl.inl (precision-generic).
 */

typedef struct { R x, y, z; } FN(rgb);
#define RGB FN(rgb)

/* C.5 warp::square_to_uniform_disk_concentric / square_to_cosine_hemisphere */
static inline void FN(disk_concentric)(R ux, R uy, R *ox, R *oy) {
    R x = (R) 2 * ux - (R) 1, y = (R) 2 * uy - (R) 1;
    if (x == 0 && y == 0) { *ox = 0; *oy = 0; return; }
    int q = FN(absr)(x) < FN(absr)(y);
    R r = q ? y : x, rp = q ? x : y;
    R phi = (R) 0.25 * (R) M_PI * rp / r;
    if (q) phi = (R) 0.5 * (R) M_PI - phi;
    *ox = r * RCOS_FN(phi); *oy = r * RSIN_FN(phi);
}

/* mitsuba fresnel(cos_theta_i, eta) -> r, cos_theta_t, eta_it, eta_ti */
static inline R FN(fresnel)(R cos_i, R eta, R *cos_t, R *eta_it, R *eta_ti) {
    int outside = cos_i >= 0;
    R rcp_eta = (R) 1 / eta;
    *eta_it = outside ? eta : rcp_eta;
    *eta_ti = outside ? rcp_eta : eta;
    R ct2 = (R) 1 - ((R) 1 - cos_i * cos_i) * (*eta_ti) * (*eta_ti);
    R ci = FN(absr)(cos_i), ct = RSQRT_FN(FN(maxr)(ct2, 0));
    R a_s = (-(*eta_it) * ct + ci) / ((*eta_it) * ct + ci);
    R a_p = (-(*eta_it) * ci + ct) / ((*eta_it) * ci + ct);
    R r = (R) 0.5 * (a_s * a_s + a_p * a_p);
    if (eta == 1) r = 0; else if (ci == 0) r = 1;
    *cos_t = cos_i >= 0 ? -ct : ct;   /* mulsign_neg(ct, cos_i) */
    return r;
}

static inline R FN(mis)(R a, R b) {
    a *= a; b *= b;
    R w = a / (a + b);
    return isfinite((double) w) ? w : 0;
}

typedef struct {
    AFF T;                  /* camera to_world */
    R tan_x, tan_y, near_clip;
    int W, H, max_depth, rr_depth, tent;
} FN(cam);
#define CAM FN(cam)

static void FN(cam_setup)(const orc_render_params *p, CAM *c) {
    double m[12];
    for (int i = 0; i < 12; i++) m[i] = p->to_world[i];
    c->T = FN(aff_from)(m);
    c->W = p->width; c->H = p->height; c->max_depth = p->max_depth; c->rr_depth = p->rr_depth; c->tent = p->rfilter == 1;
    double t = tan(p->fov_deg * M_PI / 360.0);          /* fov along the SMALLER axis (cbox.xml:12) */
    if (p->width <= p->height) { c->tan_x = (R) t; c->tan_y = (R) (t * p->height / p->width); }
    else { c->tan_y = (R) t; c->tan_x = (R) (t * p->width / p->height); }
    c->near_clip = (R) p->near_clip;
}

/* Radiance along one camera path (path.cpp sample()); draws from the path's PCG32 in Mitsuba's order. */
static RGB FN(pt_li)(const orc_scene *sc, const CAM *cam, V3 o, V3 d, uint64_t *state, uint64_t inc, orc_stats *st,
                     uint64_t *shadow_rays) {
    RGB thr = { 1, 1, 1 }, res = { 0, 0, 0 };
    R eta = 1, prev_pdf = 1;
    int prev_delta = 1, depth = 0;
    V3 prev_p = o;
    for (;;) {
        HIT h;
        int valid = FN(closest)(sc, o, d, RINF, &h, st);
        if (valid) { if (st) st->segments++; }
        /* ---- direct emission (with MIS against emitter sampling at the previous vertex) ---- */
        if (valid) {
            const orc_material *m = &sc->materials[h.material];
            if (m->emission[0] > 0 || m->emission[1] > 0 || m->emission[2] > 0) {
                V3 md = FN(neg)(d);
                R cos_e = FN(dot)(md, h.ns);             /* Frame::cos_theta(si.wi) > 0: front side only */
                if (cos_e > 0) {
                    R em_pdf = 0;
                    if (!prev_delta) {
                        V3 dv = FN(sub)(h.p, prev_p);
                        R dist2 = FN(dot)(dv, dv);
                        R dp = FN(absr)(FN(dot)(d, h.ng));
                        int ei = sc->shape_emitter[h.shape];
                        em_pdf = (R) sc->emitter_inv_area[ei] * dist2 / dp / (R) sc->n_emitters;
                        if (!isfinite((double) em_pdf)) em_pdf = 0;
                    }
                    R w = FN(mis)(prev_pdf, em_pdf);
                    res.x += thr.x * (R) m->emission[0] * w;
                    res.y += thr.y * (R) m->emission[1] * w;
                    res.z += thr.z * (R) m->emission[2] * w;
                }
            }
        }
        if (!(depth + 1 < cam->max_depth) || !valid) break;
        const orc_material *m = &sc->materials[h.material];
        V3 md = FN(neg)(d);
        V3 wi = FN(mk)(FN(dot)(md, h.fs), FN(dot)(md, h.ft), FN(dot)(md, h.ns));
        int smooth = m->kind == ORC_MAT_DIFFUSE;
        /* ---- emitter sampling ---- */
        RGB em_w = { 0, 0, 0 };
        V3 wo_em = { 0, 0, 0 };
        R ds_pdf = 0;
        int active_em = 0;
        if (smooth && sc->n_emitters > 0) {
            R u1 = (R) orc_pcg32_next_f32(state, inc), u2 = (R) orc_pcg32_next_f32(state, inc);
            /* Scene::sample_emitter: uniform over emitters, sample.x re-used */
            R fe = u1 * (R) sc->n_emitters;
            int ei = (int) fe;
            if (ei > sc->n_emitters - 1) ei = sc->n_emitters - 1;
            u1 = fe - (R) ei;
            /* Mesh::sample_position: area-weighted face pick with sample.y re-used, then uniform triangle */
            int f0 = sc->emitter_first[ei], f1 = sc->emitter_first[ei + 1];
            R total = (R) sc->em_cdf[f1 - 1];
            R target = u2 * total;
            int f = f0;
            while (f < f1 - 1 && (R) sc->em_cdf[f] < target) f++;
            R lo = f > f0 ? (R) sc->em_cdf[f - 1] : 0;
            u2 = (target - lo) / ((R) sc->em_cdf[f] - lo);
            const double *tv = sc->tri_v + 9 * (size_t) sc->em_tri[f];
            V3 p0 = FN(mk)((R) tv[0], (R) tv[1], (R) tv[2]), p1 = FN(mk)((R) tv[3], (R) tv[4], (R) tv[5]),
               p2 = FN(mk)((R) tv[6], (R) tv[7], (R) tv[8]);
            R tq = RSQRT_FN(FN(maxr)((R) 1 - u1, 0));
            R b1 = (R) 1 - tq, b2 = tq * u2;           /* square_to_uniform_triangle */
            V3 e0 = FN(sub)(p1, p0), e1 = FN(sub)(p2, p0);
            V3 ps = FN(add)(p0, FN(add)(FN(scl)(e0, b1), FN(scl)(e1, b2)));
            V3 pn = FN(normalize)(FN(cross)(e0, e1));
            if (sc->tri_flip[sc->em_tri[f]]) pn = FN(neg)(pn);
            V3 dv = FN(sub)(ps, h.p);
            R dist2 = FN(dot)(dv, dv), dist = RSQRT_FN(dist2);
            V3 dd = FN(scl)(dv, (R) 1 / dist);
            R dp = FN(absr)(FN(dot)(dd, pn));
            R x = dist2 / dp;
            R pdf = (R) sc->emitter_inv_area[ei] * (isfinite((double) x) ? x : 0);
            int ok = FN(dot)(dd, pn) < 0 && pdf != 0;
            if (ok) {
                pdf = pdf / (R) sc->n_emitters;          /* emitter selection pmf */
                /* visibility: spawn_ray_to(ds.p), maxt = dist * (1 - ShadowEpsilon) */
                V3 so = FN(spawn)(h.p, h.ng, dd);
                V3 sv = FN(sub)(ps, so);
                R sd = FN(norm)(sv);
                V3 sdir = FN(scl)(sv, (R) 1 / sd);
                if (shadow_rays) (*shadow_rays)++;
                if (!FN(occluded)(sc, so, sdir, sd * ((R) 1 - (R) (10.0 * ORC_RAY_EPSILON)), st)) {
                    const orc_material *me = &sc->materials[sc->tri_material[sc->em_tri[f]]];
                    em_w.x = (R) me->emission[0] / pdf; em_w.y = (R) me->emission[1] / pdf; em_w.z = (R) me->emission[2] / pdf;
                }
                ds_pdf = pdf;
                active_em = 1;
                wo_em = FN(mk)(FN(dot)(dd, h.fs), FN(dot)(dd, h.ft), FN(dot)(dd, h.ns));
            }
        }
        /* ---- BSDF eval + sample ---- */
        R s1 = (R) orc_pcg32_next_f32(state, inc);
        R s2x = (R) orc_pcg32_next_f32(state, inc), s2y = (R) orc_pcg32_next_f32(state, inc);
        V3 wo = { 0, 0, 0 };
        RGB bw = { 0, 0, 0 };
        R bs_pdf = 0, bs_eta = 1;
        int bs_delta = 0;
        if (m->kind == ORC_MAT_DIFFUSE) {
            R ci = wi.z;
            R dx, dy;
            FN(disk_concentric)(s2x, s2y, &dx, &dy);
            R z = RSQRT_FN(FN(maxr)((R) 1 - dx * dx - dy * dy, 0));   /* square_to_cosine_hemisphere */
            wo = FN(mk)(dx, dy, z);
            bs_pdf = z * (R) (1.0 / M_PI);
            if (ci > 0 && bs_pdf > 0) { bw.x = (R) m->p[0]; bw.y = (R) m->p[1]; bw.z = (R) m->p[2]; }
            if (active_em) {
                R co = wo_em.z;
                if (ci > 0 && co > 0) {
                    R f = co * (R) (1.0 / M_PI);
                    R w = FN(mis)(ds_pdf, f);                 /* bsdf pdf of the emitter direction = cos/pi */
                    res.x += thr.x * (R) m->p[0] * f * em_w.x * w;
                    res.y += thr.y * (R) m->p[1] * f * em_w.y * w;
                    res.z += thr.z * (R) m->p[2] * f * em_w.z * w;
                }
            }
        } else if (m->kind == ORC_MAT_DIELECTRIC) {
            R ct, eit, eti;
            R r = FN(fresnel)(wi.z, (R) (m->p[0] / m->p[1]), &ct, &eit, &eti);
            bs_delta = 1;
            if (s1 <= r) {
                wo = FN(mk)(-wi.x, -wi.y, wi.z); bs_pdf = r; bs_eta = 1; bw.x = bw.y = bw.z = 1;
            } else {
                wo = FN(mk)(-eti * wi.x, -eti * wi.y, ct); bs_pdf = (R) 1 - r; bs_eta = eit;
                bw.x = bw.y = bw.z = eti * eti;               /* radiance mode: * eta_ti^2 */
            }
        } else if (m->kind == ORC_MAT_CONDUCTOR) {
            bs_delta = 1;
            wo = FN(mk)(-wi.x, -wi.y, wi.z); bs_pdf = 1;
            if (wi.z > 0) { bw.x = (R) m->p[0]; bw.y = (R) m->p[1]; bw.z = (R) m->p[2]; }
        } else {
            break;
        }
        /* ---- continue ---- */
        V3 wd = FN(add)(FN(add)(FN(scl)(h.fs, wo.x), FN(scl)(h.ft, wo.y)), FN(scl)(h.ns, wo.z));
        prev_p = h.p;
        o = FN(spawn)(h.p, h.ng, wd);
        d = wd;
        thr.x *= bw.x; thr.y *= bw.y; thr.z *= bw.z;
        eta *= bs_eta;
        prev_pdf = bs_pdf; prev_delta = bs_delta;
        depth++;
        R tmax = FN(maxr)(thr.x, FN(maxr)(thr.y, thr.z));
        R rr_prob = FN(minr)(tmax * eta * eta, (R) 0.95);
        R u_rr = (R) orc_pcg32_next_f32(state, inc);
        int rr_active = depth >= cam->rr_depth;
        if (rr_active) { R inv = (R) 1 / rr_prob; thr.x *= inv; thr.y *= inv; thr.z *= inv; }
        if ((rr_active && !(u_rr < rr_prob)) || tmax == 0) break;
    }
    return res;
}

typedef struct {
    const orc_scene *sc; const CAM *cam; uint64_t seed; uint32_t spp_total, s_offset, s_stride;
    int *next_row; double *film; orc_stats st; uint64_t shadow;
} FN(pt_job);

static void FN(splat)(const CAM *cam, double *film, R px, R py, RGB v) {
    /* ImageBlock::put with a tent filter of radius 1 (pixel centres at i + 0.5); box filter otherwise */
    if (!cam->tent) {
        int x = (int) RFLOOR_FN(px), y = (int) RFLOOR_FN(py);
        if (x < 0 || y < 0 || x >= cam->W || y >= cam->H) return;
        double *q = film + 4 * ((size_t) y * cam->W + x);
        orc_atomic_add(q, v.x); orc_atomic_add(q + 1, v.y); orc_atomic_add(q + 2, v.z); orc_atomic_add(q + 3, 1.0);
        return;
    }
    int x0 = (int) RFLOOR_FN(px - (R) 0.5), y0 = (int) RFLOOR_FN(py - (R) 0.5);
    for (int dy = 0; dy < 2; dy++)
        for (int dx = 0; dx < 2; dx++) {
            int x = x0 + dx, y = y0 + dy;
            if (x < 0 || y < 0 || x >= cam->W || y >= cam->H) continue;
            R wx = (R) 1 - FN(absr)((R) x + (R) 0.5 - px), wy = (R) 1 - FN(absr)((R) y + (R) 0.5 - py);
            R w = FN(maxr)(wx, 0) * FN(maxr)(wy, 0);
            if (w <= 0) continue;
            double *q = film + 4 * ((size_t) y * cam->W + x);
            orc_atomic_add(q, (double) (v.x * w)); orc_atomic_add(q + 1, (double) (v.y * w));
            orc_atomic_add(q + 2, (double) (v.z * w)); orc_atomic_add(q + 3, (double) w);
        }
}

static void *FN(pt_worker)(void *arg) {
    FN(pt_job) *jb = (FN(pt_job) *) arg;
    const CAM *cam = jb->cam;
    for (;;) {
        int y = __atomic_fetch_add(jb->next_row, 1, __ATOMIC_RELAXED);
        if (y >= cam->H) break;
        for (int x = 0; x < cam->W; x++)
            for (uint32_t s = jb->s_offset; s < jb->spp_total; s += jb->s_stride) {
                uint64_t path = ((uint64_t) y * cam->W + x) * (uint64_t) jb->spp_total + s;
                uint64_t state, inc;
                orc_path_rng(jb->seed, path, &state, &inc);
                R jx = (R) orc_pcg32_next_f32(&state, inc), jy = (R) orc_pcg32_next_f32(&state, inc);
                R px = (R) x + jx, py = (R) y + jy;
                /* perspective sensor: x/z = (1 - 2 sx) tan_x, y/z = (1 - 2 sy) tan_y (C.7) */
                R sx = px / (R) cam->W, sy = py / (R) cam->H;
                V3 dl = FN(normalize)(FN(mk)(((R) 1 - (R) 2 * sx) * cam->tan_x, ((R) 1 - (R) 2 * sy) * cam->tan_y, 1));
                V3 o = FN(mk)(cam->T.m[3], cam->T.m[7], cam->T.m[11]);
                V3 d = FN(xvec)(&cam->T, dl);
                o = FN(add)(o, FN(scl)(d, cam->near_clip / dl.z));
                jb->st.paths++;
                RGB L = FN(pt_li)(jb->sc, cam, o, d, &state, inc, &jb->st, &jb->shadow);
                FN(splat)(cam, jb->film, px, py, L);
            }
    }
    return NULL;
}

static int FN(render_path)(const orc_scene *sc, const orc_render_params *p, uint64_t seed, uint32_t spp_total, uint32_t s_offset,
                           uint32_t s_stride, double *film, orc_stats *stats, uint64_t *shadow_rays, int n_threads) {
    CAM cam;
    FN(cam_setup)(p, &cam);
    if (s_stride == 0) s_stride = 1;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    FN(pt_job) jobs[256];
    pthread_t th[256];
    int next_row = 0;
    for (int w = 0; w < n_threads; w++) {
        FN(pt_job) *jb = &jobs[w];
        jb->sc = sc; jb->cam = &cam; jb->seed = seed; jb->spp_total = spp_total; jb->s_offset = s_offset; jb->s_stride = s_stride;
        jb->next_row = &next_row; jb->film = film; jb->shadow = 0;
        memset(&jb->st, 0, sizeof jb->st);
        if (w + 1 < n_threads) pthread_create(&th[w], NULL, FN(pt_worker), jb);
    }
    FN(pt_worker)(&jobs[n_threads - 1]);
    for (int w = 0; w + 1 < n_threads; w++) pthread_join(th[w], NULL);
    orc_stats acc;
    memset(&acc, 0, sizeof acc);
    uint64_t sh = 0;
    for (int w = 0; w < n_threads; w++) {
        acc.paths += jobs[w].st.paths; acc.segments += jobs[w].st.segments; acc.rays += jobs[w].st.rays;
        acc.nodes_visited += jobs[w].st.nodes_visited; acc.tris_tested += jobs[w].st.tris_tested;
        sh += jobs[w].shadow;
    }
    if (stats) *stats = acc;
    if (shadow_rays) *shadow_rays = sh;
    return 0;
}

#undef RGB
#undef CAM
