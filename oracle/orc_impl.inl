/*
 * oracle/orc_impl.inl -- precision-generic body of the CPU oracle (TEST INFRASTRUCTURE).
 * Included twice by orc.c with R = float (suffix _f32) and R = double (suffix _f64).
 * Compiled with -ffp-contract=off so every +,-,*,/,sqrt is a single IEEE operation.
 *
 * Citations "CI:n" = /root/reference/CustomIntegrator.py line n, "CB:n" = /root/reference/CustomBSDF.py
 * line n; "C.x" = SURVEY.md Appendix C (Mitsuba 3 semantics, restated from memory of the
 * un-vendored `mitsuba` wheel); "F" = SURVEY.md Appendix F (canonical single-path specification).
 */

typedef struct { R x, y, z; } FN(v3);
#define V3 FN(v3)

static inline V3 FN(mk)(R x, R y, R z) { V3 r = { x, y, z }; return r; }
static inline V3 FN(add)(V3 a, V3 b) { return FN(mk)(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 FN(sub)(V3 a, V3 b) { return FN(mk)(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 FN(scl)(V3 a, R s) { return FN(mk)(a.x * s, a.y * s, a.z * s); }
static inline V3 FN(neg)(V3 a) { return FN(mk)(-a.x, -a.y, -a.z); }
static inline R  FN(dot)(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 FN(cross)(V3 a, V3 b) {
    return FN(mk)(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline R  FN(norm)(V3 a) { return RSQRT_FN(FN(dot)(a, a)); }
static inline V3 FN(normalize)(V3 a) { R l = FN(norm)(a); return FN(mk)(a.x / l, a.y / l, a.z / l); }
static inline R  FN(maxr)(R a, R b) { return a > b ? a : b; }
static inline R  FN(minr)(R a, R b) { return a < b ? a : b; }
static inline R  FN(absr)(R a) { return a < 0 ? -a : a; }

/* affine 3x4 (row-major) applied to point / vector / normal (C.1) */
typedef struct { R m[12]; } FN(aff);
#define AFF FN(aff)
static inline AFF FN(aff_from)(const double *m) { AFF a; for (int i = 0; i < 12; i++) a.m[i] = (R) m[i]; return a; }
static inline V3 FN(xpoint)(const AFF *a, V3 p) {
    return FN(mk)(a->m[0] * p.x + a->m[1] * p.y + a->m[2] * p.z + a->m[3],
                  a->m[4] * p.x + a->m[5] * p.y + a->m[6] * p.z + a->m[7],
                  a->m[8] * p.x + a->m[9] * p.y + a->m[10] * p.z + a->m[11]);
}
static inline V3 FN(xvec)(const AFF *a, V3 p) {
    return FN(mk)(a->m[0] * p.x + a->m[1] * p.y + a->m[2] * p.z,
                  a->m[4] * p.x + a->m[5] * p.y + a->m[6] * p.z,
                  a->m[8] * p.x + a->m[9] * p.y + a->m[10] * p.z);
}
/* normal: multiply by the transpose of the inverse (pass the INVERSE affine) */
static inline V3 FN(xnormal)(const AFF *inv, V3 n) {
    return FN(mk)(inv->m[0] * n.x + inv->m[4] * n.y + inv->m[8] * n.z,
                  inv->m[1] * n.x + inv->m[5] * n.y + inv->m[9] * n.z,
                  inv->m[2] * n.x + inv->m[6] * n.y + inv->m[10] * n.z);
}

/* C.4 Frame3f(n): Duff et al. 2017 orthonormal basis, as Mitsuba's coordinate_system() */
static inline void FN(coordinate_system)(V3 n, V3 *s, V3 *t) {
    R sign = COPYSIGN_FN((R) 1, n.z);
    R a = -((R) 1 / (sign + n.z));
    R b = n.x * n.y * a;
    /* mulsign(x, n.z) = x * sign(n.z); mulsign_neg = -x * sign(n.z) */
    *s = FN(mk)((n.x * n.x * a) * sign + (R) 1, b * sign, -n.x * sign);
    *t = FN(mk)(b, n.y * (n.y * a) + sign, -n.y);
}

typedef struct {
    R  t;
    V3 p, ng, ns, fs, ft; /* position, geometric normal, shading frame (s, t, n) */
    int prim, shape, material;
} FN(hit);
#define HIT FN(hit)

/* C.2 math::solve_quadratic (numerically stable form) */
static inline int FN(solve_quadratic)(R a, R b, R c, R *x0, R *x1) {
    if (a == 0) {
        if (b == 0) return 0;
        *x0 = *x1 = -c / b;
        return 1;
    }
    R discrim = b * b - (R) 4 * a * c;
    if (!(discrim >= 0)) return 0;
    R sq = RSQRT_FN(discrim);
    R temp = (R) -0.5 * (b + COPYSIGN_FN(sq, b));
    R x0p = temp / a, x1p = c / temp;
    *x0 = FN(minr)(x0p, x1p);
    *x1 = FN(maxr)(x0p, x1p);
    return 1;
}

/* Object-space intersection of one analytic primitive; returns t (world parametrisation is preserved
 * because the object-space direction is NOT re-normalised, C.2) or -1 on miss. */
static R FN(intersect_prim)(const orc_prim *pr, V3 o, V3 d, R tmax) {
    AFF inv = FN(aff_from)(pr->to_object);
    if (pr->kind == ORC_SPHERE) {
        /* C.2 sphere: centre = to_world*(0,0,0), radius = |to_world*(1,0,0)|, quadratic in world space */
        V3 c = FN(mk)((R) pr->to_world[3], (R) pr->to_world[7], (R) pr->to_world[11]);
        R r = (R) pr->radius;
        V3 oc = FN(sub)(o, c);
        R A = FN(dot)(d, d), B = (R) 2 * FN(dot)(oc, d), C = FN(dot)(oc, oc) - r * r;
        R n0, n1;
        if (!FN(solve_quadratic)(A, B, C, &n0, &n1)) return -1;
        if (!(n0 <= tmax && n1 >= 0)) return -1;     /* out of bounds */
        if (n0 < 0 && n1 > tmax) return -1;           /* segment inside the sphere */
        return n0 < 0 ? n1 : n0;
    }
    V3 ol = FN(xpoint)(&inv, o), dl = FN(xvec)(&inv, d);
    if (pr->kind == ORC_RECTANGLE || pr->kind == ORC_DISK) {
        /* C.2 rectangle: t = -o.z/d.z, |x|,|y| <= 1 ; disk: x^2+y^2 <= 1 */
        R t = -ol.z / dl.z;
        if (!(t >= 0 && t <= tmax)) return -1;
        R lx = ol.x + t * dl.x, ly = ol.y + t * dl.y;
        if (pr->kind == ORC_RECTANGLE) {
            if (!(FN(absr)(lx) <= 1 && FN(absr)(ly) <= 1)) return -1;
        } else {
            if (!(lx * lx + ly * ly <= 1)) return -1;
        }
        return t;
    }
    R A, B, C;
    if (pr->kind == ORC_CONE) {
        /* builder-defined: x^2 + y^2 - (1-z)^2 = 0, 0 <= z <= 1 (MitsubaScenes/Cone_*.xml:36) */
        R w = (R) 1 - ol.z;
        A = dl.x * dl.x + dl.y * dl.y - dl.z * dl.z;
        B = (R) 2 * (ol.x * dl.x + ol.y * dl.y + w * dl.z);
        C = ol.x * ol.x + ol.y * ol.y - w * w;
    } else { /* ORC_CYLINDER x^2+y^2 = 1, 0 <= z <= 1 */
        A = dl.x * dl.x + dl.y * dl.y;
        B = (R) 2 * (ol.x * dl.x + ol.y * dl.y);
        C = ol.x * ol.x + ol.y * ol.y - (R) 1;
    }
    R n0, n1;
    if (!FN(solve_quadratic)(A, B, C, &n0, &n1)) return -1;
    R z0 = ol.z + n0 * dl.z, z1 = ol.z + n1 * dl.z;
    if (n0 >= 0 && n0 <= tmax && z0 >= 0 && z0 <= 1) return n0;
    if (n1 >= 0 && n1 <= tmax && z1 >= 0 && z1 <= 1) return n1;
    return -1;
}

/* C.3 initialize_sh_frame: s = normalize(dp_du - n (n.dp_du)), t = n x s ; falls back to Frame3f(n)
 * where dp_du is degenerate (sphere pole / cone apex) */
static inline void FN(finish_frame)(HIT *h, V3 dp_du) {
    V3 s = FN(sub)(dp_du, FN(scl)(h->ns, FN(dot)(h->ns, dp_du)));
    R l2 = FN(dot)(s, s);
    if (l2 > 0) {
        h->fs = FN(normalize)(s);
        h->ft = FN(cross)(h->ns, h->fs);
    } else {
        FN(coordinate_system)(h->ns, &h->fs, &h->ft);
    }
}

static void FN(fill_prim_hit)(const orc_prim *pr, int index, V3 o, V3 d, R t, HIT *h) {
    AFF M = FN(aff_from)(pr->to_world), inv = FN(aff_from)(pr->to_object);
    h->t = t; h->prim = index; h->shape = pr->shape; h->material = pr->material;
    V3 pw = FN(add)(o, FN(scl)(d, t));
    V3 dp_du;
    if (pr->kind == ORC_SPHERE) {
        V3 c = FN(mk)(M.m[3], M.m[7], M.m[11]);
        R r = (R) pr->radius;
        V3 n = FN(normalize)(FN(sub)(pw, c));
        h->p = FN(add)(c, FN(scl)(n, r));           /* re-projected onto the sphere */
        V3 loc = FN(xpoint)(&inv, h->p);
        dp_du = FN(xvec)(&M, FN(mk)(-loc.y, loc.x, 0));
        h->ng = n;
    } else if (pr->kind == ORC_RECTANGLE || pr->kind == ORC_DISK) {
        V3 ol = FN(xpoint)(&inv, o), dl = FN(xvec)(&inv, d);
        V3 loc = FN(mk)(ol.x + t * dl.x, ol.y + t * dl.y, 0);
        h->p = FN(xpoint)(&M, loc);
        h->ng = FN(normalize)(FN(xnormal)(&inv, FN(mk)(0, 0, 1)));
        dp_du = FN(xvec)(&M, FN(mk)(2, 0, 0));
    } else {
        V3 loc = FN(xpoint)(&inv, pw);
        V3 nl = (pr->kind == ORC_CONE) ? FN(mk)(loc.x, loc.y, (R) 1 - loc.z) : FN(mk)(loc.x, loc.y, 0);
        h->p = pw;
        h->ng = FN(normalize)(FN(xnormal)(&inv, nl));
        dp_du = FN(xvec)(&M, FN(mk)(-loc.y, loc.x, 0));
    }
    if (pr->flip) h->ng = FN(neg)(h->ng);
    h->ns = h->ng;
    FN(finish_frame)(h, dp_du);
}

/* C.2 Mesh::ray_intersect_triangle (Moeller-Trumbore) */
static inline int FN(intersect_tri)(const double *tv, V3 o, V3 d, R tmax, R *t_out, R *u_out, R *v_out) {
    V3 p0 = FN(mk)((R) tv[0], (R) tv[1], (R) tv[2]);
    V3 p1 = FN(mk)((R) tv[3], (R) tv[4], (R) tv[5]);
    V3 p2 = FN(mk)((R) tv[6], (R) tv[7], (R) tv[8]);
    V3 e1 = FN(sub)(p1, p0), e2 = FN(sub)(p2, p0);
    V3 pvec = FN(cross)(d, e2);
    R det = FN(dot)(e1, pvec);
    if (det == 0) return 0;
    R inv_det = (R) 1 / det;
    V3 tvec = FN(sub)(o, p0);
    R u = FN(dot)(tvec, pvec) * inv_det;
    if (!(u >= 0 && u <= 1)) return 0;
    V3 qvec = FN(cross)(tvec, e1);
    R v = FN(dot)(d, qvec) * inv_det;
    if (!(v >= 0 && u + v <= 1)) return 0;
    R t = FN(dot)(e2, qvec) * inv_det;
    if (!(t >= 0 && t <= tmax)) return 0;
    *t_out = t; *u_out = u; *v_out = v;
    return 1;
}

static void FN(fill_tri_hit)(const orc_scene *sc, int tri, V3 o, V3 d, R t, R u, R v, HIT *h) {
    const double *tv = sc->tri_v + 9 * (size_t) tri;
    V3 p0 = FN(mk)((R) tv[0], (R) tv[1], (R) tv[2]);
    V3 p1 = FN(mk)((R) tv[3], (R) tv[4], (R) tv[5]);
    V3 p2 = FN(mk)((R) tv[6], (R) tv[7], (R) tv[8]);
    R b0 = (R) 1 - u - v;
    (void) o; (void) d;
    h->t = t; h->prim = sc->n_prims + tri; h->shape = sc->tri_shape[tri]; h->material = sc->tri_material[tri];
    /* C.2: p = b0 p0 + b1 p1 + b2 p2 ; n = normalize(cross(p1-p0, p2-p0)) */
    h->p = FN(add)(FN(add)(FN(scl)(p0, b0), FN(scl)(p1, u)), FN(scl)(p2, v));
    h->ng = FN(normalize)(FN(cross)(FN(sub)(p1, p0), FN(sub)(p2, p0)));
    if (sc->tri_has_n[tri]) {
        const double *tn = sc->tri_n + 9 * (size_t) tri;
        V3 n0 = FN(mk)((R) tn[0], (R) tn[1], (R) tn[2]);
        V3 n1 = FN(mk)((R) tn[3], (R) tn[4], (R) tn[5]);
        V3 n2 = FN(mk)((R) tn[6], (R) tn[7], (R) tn[8]);
        h->ns = FN(normalize)(FN(add)(FN(add)(FN(scl)(n0, b0), FN(scl)(n1, u)), FN(scl)(n2, v)));
    } else {
        h->ns = h->ng;
    }
    if (sc->tri_flip[tri]) { h->ng = FN(neg)(h->ng); h->ns = FN(neg)(h->ns); }
    /* no UVs: (dp_du, dp_dv) = coordinate_system(n) */
    V3 du, dv;
    FN(coordinate_system)(h->ng, &du, &dv);
    FN(finish_frame)(h, du);
}

static inline int FN(box_hit)(const orc_node *nd, V3 o, V3 inv_d, R tmax) {
    R t0 = 0, t1 = tmax;
    const R ox[3] = { o.x, o.y, o.z }, id[3] = { inv_d.x, inv_d.y, inv_d.z };
    for (int a = 0; a < 3; a++) {
        R lo = ((R) nd->lo[a] - ox[a]) * id[a], hi = ((R) nd->hi[a] - ox[a]) * id[a];
        if (lo > hi) { R tmp = lo; lo = hi; hi = tmp; }
        /* NaN (0*inf) slabs are ignored */
        if (lo > t0) t0 = lo;
        if (hi < t1) t1 = hi;
    }
    /* conservative: widen by a few ulps so the BVH never loses a hit that brute force finds */
    return t0 <= t1 * ((R) 1 + (R) 8 * REPS) + (R) 8 * REPS;
}

/* nearest triangle hit; any_hit != 0 returns at the first hit found */
static int FN(closest_tri)(const orc_scene *sc, V3 o, V3 d, R tmax, int any_hit, R *t_best, R *u_best, R *v_best,
                           orc_stats *st) {
    int best = -1;
    R tb = tmax, t, u, v;
    if (!sc->use_bvh || sc->n_nodes == 0) {
        for (int i = 0; i < sc->n_tris; i++) {
            if (st) st->tris_tested++;
            if (FN(intersect_tri)(sc->tri_v + 9 * (size_t) i, o, d, tb, &t, &u, &v) && (best < 0 || t < tb)) {
                best = i; tb = t; *u_best = u; *v_best = v;
                if (any_hit) break;
            }
        }
        *t_best = tb;
        return best;
    }
    V3 inv_d = FN(mk)((R) 1 / d.x, (R) 1 / d.y, (R) 1 / d.z);
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const orc_node *nd = &sc->nodes[stack[--sp]];
        if (st) st->nodes_visited++;
        if (!FN(box_hit)(nd, o, inv_d, tb)) continue;
        if (nd->count > 0) {
            for (int j = 0; j < nd->count; j++) {
                int i = sc->tri_order[nd->first + j];
                if (st) st->tris_tested++;
                if (FN(intersect_tri)(sc->tri_v + 9 * (size_t) i, o, d, tb, &t, &u, &v) &&
                    (best < 0 || t < tb || (t == tb && i < best))) {
                    best = i; tb = t; *u_best = u; *v_best = v;
                    if (any_hit) { *t_best = tb; return best; }
                }
            }
        } else {
            stack[sp++] = nd->right;
            stack[sp++] = nd->left;
        }
    }
    *t_best = tb;
    return best;
}

/* scene.ray_intersect (CI:146,309): nearest hit over analytic primitives and triangles */
static int FN(closest)(const orc_scene *sc, V3 o, V3 d, R tmax, HIT *h, orc_stats *st) {
    int best = -1;
    R tb = tmax;
    for (int i = 0; i < sc->n_prims; i++) {
        R t = FN(intersect_prim)(&sc->prims[i], o, d, tb);
        if (t >= 0 && (best < 0 || t < tb)) { best = i; tb = t; }
    }
    R tt, u = 0, v = 0;
    int tri = sc->n_tris ? FN(closest_tri)(sc, o, d, tb, 0, &tt, &u, &v, st) : -1;
    if (st) st->rays++;
    if (tri >= 0 && (best < 0 || tt < tb)) {
        FN(fill_tri_hit)(sc, tri, o, d, tt, u, v, h);
        return 1;
    }
    if (best < 0) return 0;
    FN(fill_prim_hit)(&sc->prims[best], best, o, d, tb, h);
    return 1;
}

static int FN(occluded)(const orc_scene *sc, V3 o, V3 d, R tmax, orc_stats *st) {
    if (st) st->rays++;
    for (int i = 0; i < sc->n_prims; i++)
        if (FN(intersect_prim)(&sc->prims[i], o, d, tmax) >= 0) return 1;
    R tt, u, v;
    if (sc->n_tris && FN(closest_tri)(sc, o, d, tmax, 1, &tt, &u, &v, st) >= 0) return 1;
    return 0;
}

/* C.3 si.spawn_ray(d): origin = p + n_g * copysign((1 + max|p|) * RayEpsilon, dot(n_g, d)) */
static inline V3 FN(spawn)(V3 p, V3 ng, V3 d) {
    R m = FN(maxr)(FN(absr)(p.x), FN(maxr)(FN(absr)(p.y), FN(absr)(p.z)));
    R mag = ((R) 1 + m) * (R) ORC_RAY_EPSILON; /* 1500 * 2^-24: the fp32 RayEpsilon in both precisions */
    mag = COPYSIGN_FN(mag, FN(dot)(ng, d));
    return FN(add)(p, FN(scl)(ng, mag));
}

/* UltraBSDF.sample, CB:87-175, with _ggx_sample CB:30-61 and ggx_pdf == 1 (CB:81-82); Appendix F. */
static void FN(ultra_bsdf)(V3 wi, V3 ng, V3 ns, R Z, R alpha, R s1, R s2, V3 *dir, R *pdf, R *amp, int *reflect) {
    V3 fs, ft;
    FN(coordinate_system)(ng, &fs, &ft);                               /* CB:32  Frame3f(si.n)      (Q5) */
    V3 w = FN(mk)(FN(dot)(wi, fs), FN(dot)(wi, ft), FN(dot)(wi, ng)); /* CB:33  */
    V3 ws = FN(normalize)(FN(mk)(alpha * w.x, alpha * w.y, w.z));     /* CB:37-38 */
    R inv = (R) 1 / RSQRT_FN(FN(maxr)((R) 1 - ws.z * ws.z, (R) 1e-7)); /* CB:41 */
    V3 T1 = FN(mk)(ws.y * inv, -ws.x * inv, 0);                        /* CB:42-44 */
    V3 T2 = FN(cross)(ws, T1);                                         /* CB:45 */
    /* CB:48 concentric disk of the scalar sample broadcast to (s1,s1): r = 2 s1 - 1, phi = pi/4 (Q4, C.5) */
    R r = (R) 2 * s1 - (R) 1;
    R qx, qy;
    if (r == 0) { qx = 0; qy = 0; }
    else { R phi = (R) 0.25 * (R) M_PI * (r / r); qx = r * RCOS_FN(phi); qy = r * RSIN_FN(phi); }
    R S = (R) 0.5 * ((R) 1 + ws.z);                                    /* CB:51 */
    qy = ((R) 1 - S) * RSQRT_FN(FN(maxr)((R) 1 - qx * qx, 0)) + S * qy; /* CB:52 */
    R zz = RSQRT_FN(FN(maxr)((R) 1 - qx * qx - qy * qy, 0));           /* CB:55 */
    V3 ms = FN(add)(FN(add)(FN(scl)(T1, qx), FN(scl)(T2, qy)), FN(scl)(ws, zz));
    V3 m = FN(normalize)(FN(mk)(alpha * ms.x, alpha * ms.y, ms.z));   /* CB:56-59 */
    if (!(FN(dot)(m, wi) < 0)) m = FN(neg)(m);                         /* CB:100 (Q6) */
    R cwm = FN(dot)(wi, m);                                            /* CB:101 */
    R Z1 = Z, Z2 = (R) 1.2;                                            /* CB:104-107: entering is always False */
    R ratio = Z1 / Z2;                                                 /* CB:111 */
    R cTr = FN(absr)(cwm);                                             /* CB:119 */
    R sq = (R) 1 - (ratio * ratio) * ((R) 1 - cTr * cTr);             /* CB:120 */
    R cTt = RSQRT_FN(FN(maxr)(sq, 0));                                 /* CB:121 */
    R den = Z1 * cTr + Z2 * cTt;                                       /* CB:122 */
    R Ar = (Z1 * cTr - Z2 * cTt) / den;                                /* CB:123 */
    R At = (R) 1 - Ar;                                                 /* CB:124 */
    V3 refl = FN(add)(wi, FN(scl)(m, (R) 2 * cwm));                    /* CB:130 (Q8) */
    V3 trans = FN(add)(FN(scl)(refl, ratio), FN(scl)(m, ratio * cTr - cTt)); /* CB:131 */
    int rf = (sq < 0) || (s2 < Ar * Ar);                               /* CB:137-145 */
    R pdf_r = (R) 1 / ((R) 4 * FN(absr)(cwm));                         /* CB:153-154 (Q7) */
    R cwo = FN(dot)(trans, m);                                         /* CB:155 */
    R anwi = FN(absr)(FN(dot)(ns, wi));                                /* CB:156 */
    R anwo = FN(maxr)(FN(absr)(FN(dot)(ns, trans)), (R) 1e-7);        /* CB:157 */
    R pdf_t = (ratio * ratio) * FN(absr)(cwo) / (anwi * anwo);         /* CB:158 */
    *dir = rf ? refl : trans;                                          /* CB:147 (Q9: local comps used as world) */
    *pdf = rf ? pdf_r : pdf_t;                                         /* CB:166 */
    *amp = rf ? Ar : At;                                               /* CB:170 */
    *reflect = rf;
}

typedef struct {
    AFF T;      /* sensor to_world */
    V3  nT;     /* normalize(T * (0,0,1)) */
    R   c, fs, pitch, two_pi_f, att_k, alpha_m, alpha_c, cos_c, max_len, n_rays;
    int n_a, n_e, Tn, max_depth;
    unsigned qf;
} FN(acq);
#define ACQ FN(acq)

static void FN(acq_setup)(const orc_acq_params *p, ACQ *q) {
    double T12[12];
    for (int i = 0; i < 12; i++) T12[i] = p->sensor_to_world[i];
    q->T = FN(aff_from)(T12);
    q->nT = FN(normalize)(FN(xvec)(&q->T, FN(mk)(0, 0, 1)));      /* CI:123,212 */
    q->c = (R) p->sound_speed; q->fs = (R) p->fs; q->pitch = (R) p->pitch;
    q->two_pi_f = (R) (2.0 * M_PI * p->frequency);                  /* CI:168: python-double product, then Float */
    q->att_k = (R) (-p->attenuation * p->frequency * 1e-6);         /* CI:162 */
    q->alpha_m = (R) (p->main_beam_deg * M_PI / 180.0);             /* dr.deg2rad, CI:184 */
    q->alpha_c = (R) (p->cutoff_deg * M_PI / 180.0);
    q->cos_c = RCOS_FN(q->alpha_c);                                  /* CI:213 */
    q->max_len = (R) p->max_path_len;                                /* 0.2, CI:141 */
    q->n_a = p->n_angles; q->n_e = p->n_elements; q->Tn = p->time_samples; q->max_depth = p->max_depth;
    q->n_rays = (R) (p->n_angles * p->n_elements);                /* num_rays, CI:69 */
    q->qf = p->quirk_flags;
}

static inline R FN(elem_x)(const ACQ *q, int e) {
    /* CI:84  pitch * (e - (n_e - 1) * 0.5) */
    return q->pitch * ((R) e - (R) (q->n_e - 1) * (R) 0.5);
}

/* directivity_weight_i, CI:120-135 / 289-304: alpha = |acos(dot(n_T, -sec_dir))|; 1 up to alpha_m, linear ramp to 0 at alpha_c */
static inline R FN(directivity_wi)(V3 nT, V3 sec, R alpha_m, R alpha_c) {
    R dt = FN(dot)(nT, FN(neg)(sec));                                /* CI:124-125 */
    R al = FN(absr)(RACOS_FN(dt));                                   /* CI:126 */
    return al <= alpha_m ? (R) 1 : (al <= alpha_c ? (alpha_c - al) / (alpha_c - alpha_m) : (R) 0); /* CI:128-133 */
}

/* One path, Appendix F.  buf may be NULL (trace-only); rec may be NULL. */
static void FN(acq_path)(const orc_scene *sc, const ACQ *q, const orc_acq_params *p, uint64_t seed, uint32_t spp_total,
                         int a, int e, uint32_t s, double *buf, orc_seg_record *rec, orc_stats *st) {
    uint64_t path = ((uint64_t) a * (uint64_t) q->n_e + (uint64_t) e) * (uint64_t) spp_total + (uint64_t) s;
    uint64_t state, inc;
    orc_path_rng(seed, path, &state, &inc);
    R theta = (R) (p->angles_deg[a]) * (R) M_PI / (R) 180;         /* CI:78 */
    R xe = FN(elem_x)(q, e);
    R t0 = (xe * RSIN_FN(theta)) / q->c;                            /* CI:87 */
    V3 o = FN(xpoint)(&q->T, FN(mk)(xe, 0, 0));                     /* CI:97,103 */
    V3 d = FN(normalize)(FN(xvec)(&q->T, FN(mk)(RSIN_FN(theta), 0, RCOS_FN(theta)))); /* CI:98,104 */
    R amp = 1, atten = 1, tof = 0, geo = 0;
    int depth = 0;
    R inv_spp = (R) 1 / (R) spp_total;
    if (st) st->paths++;
    while (depth < q->max_depth && geo < q->max_len) {              /* CI:141 / 307 */
        HIT h;
        if (!FN(closest)(sc, o, d, RINF, &h, st)) { if (st) st->misses++; break; } /* CI:146-147 / 309-312 */
        if (st) st->segments++;
        R dist = h.t;
        geo += dist;                                                 /* CI:209 / 315 */
        R tof_here = tof + dist / q->c;                              /* CI:165 / 316 */
        if (!(q->qf & ORC_QF_TOF_LAST_SEGMENT)) tof = tof_here;     /* A.1#2 */
        R u_recv = (R) orc_pcg32_next_f32(&state, inc);              /* CI:153 / 319 */
        R s1 = (R) orc_pcg32_next_f32(&state, inc);                  /* CI:173 / 337 */
        R s2 = (R) orc_pcg32_next_f32(&state, inc);                  /* CI:174 / 337 */
        R u_rr = (R) orc_pcg32_next_f32(&state, inc);                /* CI:219 / 365 */
        int recv = (int) RFLOOR_FN(u_recv * (R) q->n_e);             /* CI:154 */
        if (recv > q->n_e - 1) recv = q->n_e - 1;
        V3 tgt = FN(xpoint)(&q->T, FN(mk)(FN(elem_x)(q, recv), 0, 0)); /* CI:156-157 */
        V3 to_t = FN(sub)(tgt, h.p);
        V3 sec = FN(normalize)(to_t);                                /* CI:158 / 322 */
        R dist_recv = FN(norm)(to_t);                                /* CI:166 / 329 */
        V3 so = FN(spawn)(h.p, h.ng, sec);
        R vis_tmax = (q->qf & ORC_QF_CONNECT_TO_TARGET) ? FN(norm)(FN(sub)(tgt, so)) * ((R) 1 - (R) 1e-4) : RINF;
        int visible = !FN(occluded)(sc, so, sec, vis_tmax, st);      /* CI:159-160 / 324-325 (Q1) */
        atten *= REXP_FN((q->att_k * dist) / (R) 8.686);             /* CI:162-163 / 328 (Q13) */
        R Ttot = (t0 + tof_here) + dist_recv / q->c;                 /* CI:167 / 329 */
        R phase = q->two_pi_f * Ttot;                                /* CI:168 / 330 */
        V3 md = FN(neg)(d);
        V3 wi = FN(mk)(FN(dot)(md, h.fs), FN(dot)(md, h.ft), FN(dot)(md, h.ns)); /* C.3 si.wi */
        const orc_material *mat = &sc->materials[h.material];
        V3 dir; R pdf, a_resp; int reflect;
        FN(ultra_bsdf)(wi, h.ng, h.ns, (R) mat->p[0], (R) mat->p[1], s1, s2, &dir, &pdf, &a_resp, &reflect); /* CI:175 / 338 */
        R cos_theta = FN(dot)(h.ns, md);                             /* CI:176 / 340 */
        amp *= a_resp * cos_theta * FN(maxr)(pdf, (R) 1e-6);         /* CI:177 / 341 (Q2) */
        R w_i = FN(directivity_wi)(q->nT, sec, q->alpha_m, q->alpha_c); /* CI:120-135 */
        R w_o = FN(dot)(d, h.ns) / q->n_rays;                      /* CI:118,184 (Q3) */
        R fd = w_i * w_o;
        R press = atten * amp * fd * RSIN_FN(phase);                 /* CI:187 / 348 */
        R kf = RRINT_FN(Ttot * q->fs);                               /* CI:191 / 351-352 round-half-even */
        long k = (long) kf;
        int in_range = (kf >= 0 && kf < (R) q->Tn);
        if (q->qf & ORC_QF_CLAMP_TIDX) {                             /* CI:192 */
            if (!(kf >= 0)) k = 0;
            if (kf > (R) (q->Tn - 1)) k = q->Tn - 1;
            in_range = 1;
        }
        if (visible && in_range) {
            if (buf) {
                size_t flat = ((size_t) a * q->n_e + recv) * (size_t) q->Tn + (size_t) k; /* CI:197-198 */
                orc_atomic_add(&buf[flat], (double) (press * inv_spp)); /* CI:203 / 354 */
            }
            if (st) st->deposits++;
        }
        d = FN(normalize)(dir);                                      /* CI:205-206 / 358-359 (Q9) */
        o = FN(spawn)(h.p, h.ng, d);
        depth++;                                                     /* CI:210 / 361 */
        R prod = atten * amp;
        R rr = (q->qf & ORC_QF_RR_NO_ABS) ? FN(minr)(prod, 1) : FN(minr)(FN(absr)(prod), 1); /* CI:220 / 364 */
        int survive = u_rr < rr;                                     /* CI:221 */
        atten = survive ? atten / rr : 0;                            /* CI:224 */
        if (rec) {
            orc_seg_record *rc = &rec[depth - 1];
            rc->valid = 1; rc->prim = h.prim; rc->shape = h.shape; rc->recv = recv; rc->visible = visible;
            rc->reflect = reflect; rc->k = (int32_t) k; rc->survive = survive; rc->t = (double) dist;
            rc->total_time = (double) Ttot; rc->press = (double) press; rc->amp = (double) amp;
            rc->atten = (double) atten; rc->dir[0] = d.x; rc->dir[1] = d.y; rc->dir[2] = d.z;
        }
        if (q->qf & ORC_QF_SINGLE_BOUNCE) break;                     /* A.1#8 */
        if (!survive || !(FN(dot)(d, q->nT) >= q->cos_c)) break;    /* CI:212-223 (Q11) */
    }
}

typedef struct {
    const orc_scene *sc; const ACQ *q; const orc_acq_params *p;
    uint64_t seed; uint32_t spp_total, s_offset, s_stride;
    int64_t n_s, total; int64_t *next; double *buf; orc_stats st;
} FN(job);

/* CI:380-399: the reference's thread pool hands out blocks of rays; here blocks of 4096 paths */
static void *FN(worker)(void *arg) {
    FN(job) *jb = (FN(job) *) arg;
    const int64_t CH = 4096;
    for (;;) {
        int64_t j0 = __atomic_fetch_add(jb->next, CH, __ATOMIC_RELAXED);
        if (j0 >= jb->total) break;
        int64_t j1 = j0 + CH < jb->total ? j0 + CH : jb->total;
        for (int64_t j = j0; j < j1; j++) {
            int64_t ae = j / jb->n_s;
            uint32_t s = jb->s_offset + (uint32_t) (j % jb->n_s) * jb->s_stride;
            FN(acq_path)(jb->sc, jb->q, jb->p, jb->seed, jb->spp_total, (int) (ae / jb->q->n_e), (int) (ae % jb->q->n_e), s,
                         jb->buf, NULL, &jb->st);
        }
    }
    return NULL;
}

static int FN(acquire)(const orc_scene *sc, const orc_acq_params *p, uint64_t seed, uint32_t spp_total, uint32_t s_offset,
                       uint32_t s_stride, double *buf, double *tx, orc_stats *stats, int n_threads) {
    ACQ q;
    FN(acq_setup)(p, &q);
    for (int a = 0; a < q.n_a; a++)
        for (int e = 0; e < q.n_e; e++) {
            R theta = (R) (p->angles_deg[a]) * (R) M_PI / (R) 180;
            R xe = FN(elem_x)(&q, e);
            tx[a * q.n_e + e] = (double) ((xe * RSIN_FN(theta)) / q.c); /* CI:87,94 / 254-257 */
        }
    if (s_stride == 0) s_stride = 1;
    uint64_t n_s = s_offset < spp_total ? ((uint64_t) spp_total - s_offset + s_stride - 1) / s_stride : 0;
    int64_t n_ae = (int64_t) q.n_a * q.n_e;
    int64_t total = n_ae * (int64_t) n_s;
    orc_stats acc;
    memset(&acc, 0, sizeof acc);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    FN(job) jobs[256];
    pthread_t th[256];
    int64_t next = 0;
    for (int w = 0; w < n_threads; w++) {
        FN(job) *jb = &jobs[w];
        jb->sc = sc; jb->q = &q; jb->p = p; jb->seed = seed; jb->spp_total = spp_total; jb->s_offset = s_offset;
        jb->s_stride = s_stride; jb->n_s = (int64_t) n_s; jb->total = total; jb->next = &next; jb->buf = buf;
        memset(&jb->st, 0, sizeof jb->st);
        if (w + 1 < n_threads) pthread_create(&th[w], NULL, FN(worker), jb);
    }
    FN(worker)(&jobs[n_threads - 1]);
    for (int w = 0; w + 1 < n_threads; w++) pthread_join(th[w], NULL);
    for (int w = 0; w < n_threads; w++) {
        const orc_stats *loc = &jobs[w].st;
        acc.paths += loc->paths; acc.segments += loc->segments; acc.rays += loc->rays; acc.deposits += loc->deposits;
        acc.misses += loc->misses; acc.nodes_visited += loc->nodes_visited; acc.tris_tested += loc->tris_tested;
    }
    (void) n_ae;
    if (stats) *stats = acc;
    return 0;
}

#include "orc_pt.inl"

#undef V3
#undef AFF
#undef HIT
#undef ACQ
