/* Curve arithmetic "template" for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 * Included twice by cpu_ref.c, once per group:
 *   FE        field element type           (fp_t / fp2_t)
 *   FN(x)     field function name          (fp_##x / fp2_##x)
 *   PN(x)     point function/type name     (g1_##x / g2_##x)
 * Restates ark-ec 0.4 short_weierstrass::Projective (Jacobian; mixed add
 * madd-2007-bl, doubling dbl-2009-l, a = 0) and the algorithms in SURVEY.md
 * Appendix B that the reference reaches through
 *   VariableBaseMSM::msm_bigint       (e.g. legogroth16/src/prover.rs:286, bbs_plus/src/setup.rs:145)
 *   FixedBase::get_window_table / msm (utils/src/msm.rs:18-40)
 *   AffineRepr::mul_bigint            (vb_accumulator/src/witness.rs:190)
 *   CurveGroup::normalize_batch       (vb_accumulator/src/witness.rs:193)
 * ark-ec itself is not vendored under /root/reference (Cargo.toml:32-46). */

typedef struct { FE x, y; int inf; } PN(aff);
typedef struct { FE x, y, z; } PN(jac);

static void PN(jac_zero)(PN(jac) *r) { FN(set_one)(&r->x); FN(set_one)(&r->y); FN(set_zero)(&r->z); }
static int PN(jac_is_zero)(const PN(jac) *p) { return FN(is_zero)(&p->z); }

static void PN(jac_dbl)(PN(jac) *r, const PN(jac) *p) {
    if (PN(jac_is_zero)(p)) { *r = *p; return; }
    FE a, b, c, d, e, f, t;
    FN(sqr)(&a, &p->x);
    FN(sqr)(&b, &p->y);
    FN(sqr)(&c, &b);
    FN(add)(&t, &p->x, &b); FN(sqr)(&t, &t); FN(sub)(&t, &t, &a); FN(sub)(&t, &t, &c);
    FN(add)(&d, &t, &t);
    FN(add)(&e, &a, &a); FN(add)(&e, &e, &a);
    FN(sqr)(&f, &e);
    FE z3; FN(mul)(&z3, &p->y, &p->z); FN(add)(&z3, &z3, &z3);
    FN(sub)(&r->x, &f, &d); FN(sub)(&r->x, &r->x, &d);
    FN(sub)(&t, &d, &r->x); FN(mul)(&t, &e, &t);
    FN(add)(&c, &c, &c); FN(add)(&c, &c, &c); FN(add)(&c, &c, &c);
    FN(sub)(&r->y, &t, &c);
    r->z = z3;
}

/* r = p + q, q affine (mixed add) */
static void PN(jac_add_aff)(PN(jac) *r, const PN(jac) *p, const PN(aff) *q) {
    if (q->inf) { *r = *p; return; }
    if (PN(jac_is_zero)(p)) { r->x = q->x; r->y = q->y; FN(set_one)(&r->z); return; }
    FE z1z1, u2, s2, h, hh, i, j, rr, v, t;
    FN(sqr)(&z1z1, &p->z);
    FN(mul)(&u2, &q->x, &z1z1);
    FN(mul)(&s2, &q->y, &p->z); FN(mul)(&s2, &s2, &z1z1);
    if (FN(eq)(&u2, &p->x)) {
        if (FN(eq)(&s2, &p->y)) { PN(jac_dbl)(r, p); return; }
        PN(jac_zero)(r); return;
    }
    FN(sub)(&h, &u2, &p->x);
    FN(sqr)(&hh, &h);
    FN(add)(&i, &hh, &hh); FN(add)(&i, &i, &i);
    FN(mul)(&j, &h, &i);
    FN(sub)(&rr, &s2, &p->y); FN(add)(&rr, &rr, &rr);
    FN(mul)(&v, &p->x, &i);
    FE x3, y3, z3;
    FN(sqr)(&x3, &rr); FN(sub)(&x3, &x3, &j); FN(sub)(&x3, &x3, &v); FN(sub)(&x3, &x3, &v);
    FN(mul)(&t, &p->y, &j); FN(add)(&t, &t, &t);
    FN(sub)(&y3, &v, &x3); FN(mul)(&y3, &rr, &y3); FN(sub)(&y3, &y3, &t);
    FN(add)(&z3, &p->z, &h); FN(sqr)(&z3, &z3); FN(sub)(&z3, &z3, &z1z1); FN(sub)(&z3, &z3, &hh);
    r->x = x3; r->y = y3; r->z = z3;
}

/* r = p + q, both Jacobian (add-2007-bl) */
static void PN(jac_add)(PN(jac) *r, const PN(jac) *p, const PN(jac) *q) {
    if (PN(jac_is_zero)(p)) { *r = *q; return; }
    if (PN(jac_is_zero)(q)) { *r = *p; return; }
    FE z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
    FN(sqr)(&z1z1, &p->z); FN(sqr)(&z2z2, &q->z);
    FN(mul)(&u1, &p->x, &z2z2); FN(mul)(&u2, &q->x, &z1z1);
    FN(mul)(&s1, &p->y, &q->z); FN(mul)(&s1, &s1, &z2z2);
    FN(mul)(&s2, &q->y, &p->z); FN(mul)(&s2, &s2, &z1z1);
    if (FN(eq)(&u1, &u2)) {
        if (FN(eq)(&s1, &s2)) { PN(jac_dbl)(r, p); return; }
        PN(jac_zero)(r); return;
    }
    FN(sub)(&h, &u2, &u1);
    FN(add)(&i, &h, &h); FN(sqr)(&i, &i);
    FN(mul)(&j, &h, &i);
    FN(sub)(&rr, &s2, &s1); FN(add)(&rr, &rr, &rr);
    FN(mul)(&v, &u1, &i);
    FE x3, y3, z3;
    FN(sqr)(&x3, &rr); FN(sub)(&x3, &x3, &j); FN(sub)(&x3, &x3, &v); FN(sub)(&x3, &x3, &v);
    FN(mul)(&t, &s1, &j); FN(add)(&t, &t, &t);
    FN(sub)(&y3, &v, &x3); FN(mul)(&y3, &rr, &y3); FN(sub)(&y3, &y3, &t);
    FN(add)(&z3, &p->z, &q->z); FN(sqr)(&z3, &z3); FN(sub)(&z3, &z3, &z1z1); FN(sub)(&z3, &z3, &z2z2);
    FN(mul)(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}

static void PN(aff_neg)(PN(aff) *r, const PN(aff) *p) { *r = *p; if (!p->inf) FN(neg)(&r->y, &p->y); }

static void PN(jac_to_aff)(PN(aff) *r, const PN(jac) *p) {
    if (PN(jac_is_zero)(p)) { memset(r, 0, sizeof *r); r->inf = 1; return; }
    FE zi, zi2;
    FN(inv)(&zi, &p->z); FN(sqr)(&zi2, &zi);
    FN(mul)(&r->x, &p->x, &zi2);
    FN(mul)(&zi2, &zi2, &zi); FN(mul)(&r->y, &p->y, &zi2);
    r->inf = 0;
}

/* CurveGroup::normalize_batch: one shared inversion (Montgomery trick), zeros skipped */
static void PN(normalize_batch)(PN(aff) *out, const PN(jac) *in, size_t n) {
    FE *pre = (FE *)malloc((n + 1) * sizeof(FE));
    FE acc; FN(set_one)(&acc);
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (!PN(jac_is_zero)(&in[i])) FN(mul)(&acc, &acc, &in[i].z);
    }
    FE inv; FN(inv)(&inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (PN(jac_is_zero)(&in[i])) { memset(&out[i], 0, sizeof out[i]); out[i].inf = 1; continue; }
        FE zi, zi2;
        FN(mul)(&zi, &inv, &pre[i]);
        FN(mul)(&inv, &inv, &in[i].z);
        FN(sqr)(&zi2, &zi);
        FN(mul)(&out[i].x, &in[i].x, &zi2);
        FN(mul)(&zi2, &zi2, &zi);
        FN(mul)(&out[i].y, &in[i].y, &zi2);
        out[i].inf = 0;
    }
    free(pre);
}

/* AffineRepr::mul_bigint: MSB-first double-and-add over the 256-bit canonical integer */
static void PN(mul_bigint)(PN(jac) *r, const PN(aff) *p, const uint64_t s[4]) {
    PN(jac) acc; PN(jac_zero)(&acc);
    int started = 0;
    for (int i = 255; i >= 0; i--) {
        int bit = (s[i >> 6] >> (i & 63)) & 1;
        if (started) PN(jac_dbl)(&acc, &acc);
        if (bit) { PN(jac_add_aff)(&acc, &acc, p); started = 1; }
    }
    *r = acc;
}

/* VariableBaseMSM::msm_bigint -> msm_bigint_wnaf (ark-ec 0.4).
 * Parallel over windows only, exactly like cfg_into_iter!(0..digits_count). */
static void PN(msm_bigint)(PN(jac) *out, const PN(aff) *bases, const uint64_t *scalars, size_t n) {
    int c = msm_window_size(n);
    int nd = (255 + c - 1) / c;
    int32_t *digits = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1) * nd);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) make_digits(digits + i * nd, scalars + 4 * i, c, nd);
    PN(jac) *wsum = (PN(jac) *)malloc(sizeof(PN(jac)) * nd);
#pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < nd; w++) {
        size_t nb = (size_t)1 << c;   /* ark allocates 1 << c: the top digit is left unsigned */
        PN(jac) *buckets = (PN(jac) *)malloc(sizeof(PN(jac)) * nb);
        for (size_t b = 0; b < nb; b++) PN(jac_zero)(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            int32_t d = digits[i * nd + w];
            if (d > 0) PN(jac_add_aff)(&buckets[d - 1], &buckets[d - 1], &bases[i]);
            else if (d < 0) { PN(aff) nb_; PN(aff_neg)(&nb_, &bases[i]); PN(jac_add_aff)(&buckets[-d - 1], &buckets[-d - 1], &nb_); }
        }
        PN(jac) running, res; PN(jac_zero)(&running); PN(jac_zero)(&res);
        for (size_t b = nb; b-- > 0;) {
            PN(jac_add)(&running, &running, &buckets[b]);
            PN(jac_add)(&res, &res, &running);
        }
        wsum[w] = res;
        free(buckets);
    }
    PN(jac) total; PN(jac_zero)(&total);
    for (int w = nd - 1; w >= 1; w--) {
        PN(jac_add)(&total, &total, &wsum[w]);
        for (int k = 0; k < c; k++) PN(jac_dbl)(&total, &total);
    }
    PN(jac_add)(out, &total, &wsum[0]);
    free(wsum); free(digits);
}

/* FixedBase::get_window_table(255, window, g): table[k][j] = j * 2^(k*window) * g, rows
 * normalised to affine.  Row stride is 2^window (the last row is shorter, as in ark). */
static PN(aff) *PN(fixed_base_table)(const PN(aff) *g, int window, int *outerc_out) {
    int outerc = (255 + window - 1) / window;
    size_t in_window = (size_t)1 << window;
    size_t last = (size_t)1 << (255 - (outerc - 1) * window);
    PN(aff) *table = (PN(aff) *)calloc((size_t)outerc * in_window, sizeof(PN(aff)));
    PN(jac) *gouter = (PN(jac) *)malloc(sizeof(PN(jac)) * outerc);
    PN(jac) cur; cur.x = g->x; cur.y = g->y; if (g->inf) PN(jac_zero)(&cur); else FN(set_one)(&cur.z);
    for (int k = 0; k < outerc; k++) {
        gouter[k] = cur;
        for (int t = 0; t < window; t++) PN(jac_dbl)(&cur, &cur);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < outerc; k++) {
        size_t sz = (k == outerc - 1) ? last : in_window;
        PN(jac) *row = (PN(jac) *)malloc(sizeof(PN(jac)) * sz);
        PN(jac) acc; PN(jac_zero)(&acc);
        for (size_t j = 0; j < sz; j++) { row[j] = acc; PN(jac_add)(&acc, &acc, &gouter[k]); }
        PN(normalize_batch)(table + (size_t)k * in_window, row, sz);
        free(row);
    }
    free(gouter);
    *outerc_out = outerc;
    return table;
}

/* FixedBase::windowed_mul */
static void PN(windowed_mul)(PN(jac) *r, const PN(aff) *table, int window, int outerc, const uint64_t s[4]) {
    PN(jac) acc; PN(jac_zero)(&acc);
    size_t in_window = (size_t)1 << window;
    for (int k = 0; k < outerc; k++) {
        size_t idx = 0;
        for (int b = 0; b < window; b++) {
            int bit = k * window + b;
            if (bit < 256 && ((s[bit >> 6] >> (bit & 63)) & 1)) idx |= (size_t)1 << b;
        }
        PN(jac_add_aff)(&acc, &acc, &table[(size_t)k * in_window + idx]);
    }
    *r = acc;
}
