// GPU test of the C++ host-side mirror (include/dockgpu.hpp), written after the reference's own
// tests: utils/src/msm.rs:116-308 (timing_ark_ops), utils/src/randomized_pairing_check.rs:234-421
// (test_pairing_randomize), utils/src/randomized_mult_checker.rs:136-384.  The CPU oracle
// (oracle/libcpuref.so, test infrastructure) provides inputs and independent expected values.
#include <cstdio>
#include <cstdlib>
#include <thread>
#include "dockgpu.hpp"

extern "C" {
void ref_g1_generator_muls(const uint8_t *scalars, size_t m, uint8_t *out_aff);
void ref_g2_generator_muls(const uint8_t *scalars, size_t m, uint8_t *out_aff);
void ref_batch_mul_g1(const uint8_t *points, const uint8_t *scalars, size_t m, uint8_t *out_jac);
void ref_normalize_batch_g1(const uint8_t *jac, size_t n, uint8_t *out_aff);
void ref_msm_g1(const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t *out_jac);
void ref_multi_miller_loop(const uint8_t *g1s, const uint8_t *g2s, size_t k, uint8_t *out_fp12);
int ref_final_exp(const uint8_t *in_fp12, uint8_t *out_fp12);
}
using namespace dock_gpu;

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
static uint64_t next64() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static Fr rand_fr() {
    for (;;) {
        Fr r;
        for (auto &x : r.l) x = next64();
        r.l[3] &= (1ULL << 63) - 1;
        if (!Fr::geq(r.l, Fr::MODULUS)) return r;
    }
}
static std::vector<G1Affine> rand_g1(size_t n) {
    std::vector<Fr> k(n); for (auto &x : k) x = rand_fr();
    std::vector<G1Affine> out(n);
    ref_g1_generator_muls(k[0].bytes(), n, out[0].b.data());
    return out;
}
static std::vector<G2Affine> rand_g2(size_t n) {
    std::vector<Fr> k(n); for (auto &x : k) x = rand_fr();
    std::vector<G2Affine> out(n);
    ref_g2_generator_muls(k[0].bytes(), n, out[0].b.data());
    return out;
}
static G1Affine oracle_mul(const G1Affine &p, const Fr &s) {
    uint8_t jac[144]; G1Affine a;
    ref_batch_mul_g1(p.b.data(), s.bytes(), 1, jac);
    ref_normalize_batch_g1(jac, 1, a.b.data());
    return a;
}
static G1Projective lift(const G1Affine &a) {           // affine -> projective (x, y, z = R) via [1]P on the oracle
    G1Projective p; Fr one = Fr::one();
    ref_batch_mul_g1(a.b.data(), one.bytes(), 1, p.b.data());
    return p;
}
static PairingOutput oracle_multi_pairing(const std::vector<G1Affine> &a, const std::vector<G2Affine> &b) {
    Fp12 ml, out;
    ref_multi_miller_loop(a[0].b.data(), b[0].b.data(), a.size(), ml.b.data());
    CHECK(ref_final_exp(ml.b.data(), out.b.data()) == 1);
    return out;
}

static void test_fr() {
    Fr a = rand_fr(), b = rand_fr();
    CHECK((a + b) - b == a);
    CHECK(a * Fr::one() == a);
    CHECK((a * b) == (b * a));
    CHECK((a + (-a)).is_zero());
    Fr m1 = -Fr::one();                                   // r - 1
    CHECK(m1 * m1 == Fr::one());
}

static void test_window_table_and_msm() {                 // timing_ark_ops
    auto g = rand_g1(3);
    for (size_t count : {10u, 30u, 300u}) {
        std::vector<Fr> elems(count); for (auto &e : elems) e = rand_fr();
        WindowTable<G1> table(count, lift(g[0]));
        CHECK(table.scalar_size == 255 && table.window_size == WindowTable<G1>::window_size_for(count));
        CHECK(table.num_windows == (255 + table.window_size - 1) / table.window_size);
        auto many = normalize_batch<G1>(table.multiply_many(elems));
        for (size_t i = 0; i < count; i += 7) CHECK(many[i] == oracle_mul(g[0], elems[i]));
        CHECK(into_affine<G1>(table.multiply(elems[1])) == many[1]);
        CHECK(into_affine<G1>(table * elems[2]) == many[2]);
        auto many2 = normalize_batch<G1>(multiply_field_elems_with_same_group_elem<G1>(lift(g[0]), elems));
        CHECK(many2 == many);
    }
    // G::msm([g1, g2], [e1, e2]) == g1*e1 + g2*e2, via the oracle MSM
    std::vector<Fr> e = {rand_fr(), rand_fr(), rand_fr()};
    auto r = VariableBaseMSM<G1>::msm(g, e);
    CHECK(r.ok);
    uint8_t jac[144]; G1Affine exp;
    ref_msm_g1(g[0].b.data(), e[0].bytes(), 3, jac); ref_normalize_batch_g1(jac, 1, exp.b.data());
    CHECK(into_affine<G1>(r.value) == exp);
    e.pop_back();
    auto bad = VariableBaseMSM<G1>::msm(g, e);             // Err(min_len)
    CHECK(!bad.ok && bad.err_min_len == 2);
    CHECK(!VariableBaseMSM<G1>::msm_unchecked(g, e).is_zero());   // truncates instead
    CHECK(VariableBaseMSM<G1>::msm_bigint({}, {}).is_zero());
}

static void test_pairing_randomize() {
    const size_t n = 10;
    auto a1 = rand_g1(n), a2 = rand_g1(n + 5), a3 = rand_g1(n - 2);
    auto b1 = rand_g2(n), b2 = rand_g2(n + 5), b3 = rand_g2(n - 2);
    auto out1 = Bls12_381::multi_pairing(a1, b1), out2 = Bls12_381::multi_pairing(a2, b2), out3 = Bls12_381::multi_pairing(a3, b3);
    CHECK(out1 == oracle_multi_pairing(a1, b1));
    CHECK(out3 == oracle_multi_pairing(a3, b3));
    for (bool lazy : {true, false}) {
        RandomizedPairingChecker checker(rand_fr(), lazy);
        checker.add_multiple_sources_and_target(a1, b1, out1);
        checker.add_multiple_sources_and_target(a2, b2, out2);
        checker.add_multiple_sources_and_target(a3, b3, out3);
        CHECK(checker.verify());
        RandomizedPairingChecker bad(rand_fr(), lazy);       // fail on wrong output
        bad.add_multiple_sources_and_target(a1, b1, out2);
        bad.add_multiple_sources_and_target(a2, b2, out1);
        CHECK(!bad.verify());
        // e(a, b) == e(s*a, s^-1 ... ) style equalities: add_sources with c = a, d = b holds trivially
        RandomizedPairingChecker eq(rand_fr(), lazy);
        eq.add_sources(a1[0], b1[0], a1[0], b1[0]);
        eq.add_multiple_sources({a1[1], a1[2]}, {b1[1], b1[2]}, {a1[2], a1[1]}, {b1[2], b1[1]});
        CHECK(eq.verify());
        RandomizedPairingChecker neq(rand_fr(), lazy);
        neq.add_sources(a1[0], b1[0], a1[1], b1[0]);
        CHECK(!neq.verify());
    }
}

static void test_mult_checker() {
    auto g = rand_g1(3);
    Fr a[6]; for (auto &x : a) x = rand_fr();
    auto add_pts = [&](const G1Affine &p, const G1Affine &q) {   // p + q through a 2-term MSM with scalars 1, 1
        return into_affine<G1>(VariableBaseMSM<G1>::msm_bigint({p, q}, {Fr::one(), Fr::one()}));
    };
    RandomizedMultChecker<G1> ck(rand_fr());
    ck.add_1(g[0], a[0], oracle_mul(g[0], a[0]));
    ck.add_2(g[0], a[1], g[1], a[2], add_pts(oracle_mul(g[0], a[1]), oracle_mul(g[1], a[2])));
    ck.add_3(g[0], a[3], g[1], a[4], g[2], a[5], add_pts(add_pts(oracle_mul(g[0], a[3]), oracle_mul(g[1], a[4])), oracle_mul(g[2], a[5])));
    ck.add_many({g[0], g[1]}, {a[0], a[1]}, add_pts(oracle_mul(g[0], a[0]), oracle_mul(g[1], a[1])));
    CHECK(ck.verify());
    CHECK(ck.len() <= 3 + 4);                                     // g0,g1,g2 deduplicated + 4 targets
    RandomizedMultChecker<G1> bad(rand_fr());
    bad.add_1(g[0], a[0], oracle_mul(g[0], a[1]));
    CHECK(!bad.verify());
}

// The ABI is re-entrant from worker threads (rayon in the reference, SURVEY.md 8b "Threading"):
// every thread gets its own stream and scratch arena.  4 threads issue MSMs of different sizes
// concurrently; every result must equal the oracle's.
static void test_wire_formats() {                         // CanonicalSerialize / CanonicalDeserialize round trips
    std::vector<G1Affine> pts = rand_g1(9);
    pts.push_back(G1Affine{});                                                   // identity
    std::vector<G2Affine> pts2 = rand_g2(5);
    for (auto c : {Compress::Yes, Compress::No}) {
        auto enc = serialize_points<G1>(pts, c);
        CHECK(enc.size() == pts.size() * (c == Compress::Yes ? 48 : 96));
        auto dec = deserialize_points<G1>(enc, c, Validate::Yes);
        CHECK(dec.has_value() && *dec == pts);
        enc[1] ^= 0x55;                                                          // corrupt the first x coordinate
        auto bad = deserialize_points<G1>(enc, c, Validate::Yes);
        CHECK(!bad.has_value() || !((*bad)[0] == pts[0]));
        auto enc2 = serialize_points<G2>(pts2, c);
        auto dec2 = deserialize_points<G2>(enc2, c, Validate::Yes);
        CHECK(dec2.has_value() && *dec2 == pts2);
    }
    // the publicly known compressed generator 97f1d3a7...c6bb
    Fr one = Fr::one();
    G1Affine g;
    ref_g1_generator_muls(one.bytes(), 1, g.b.data());
    auto genc = serialize_points<G1>({g});
    CHECK(genc[0] == 0x97 && genc[1] == 0xf1 && genc[46] == 0xc6 && genc[47] == 0xbb);
}

static void test_concurrent_callers() {
    const size_t sizes[4] = {100, 3000, 257, 20000};
    std::vector<std::vector<G1Affine>> bases(4);
    std::vector<std::vector<Fr>> scalars(4);
    std::vector<G1Affine> expected(4);
    for (int t = 0; t < 4; t++) {
        bases[t] = rand_g1(sizes[t]);
        scalars[t].resize(sizes[t]);
        for (auto &x : scalars[t]) x = rand_fr();
        uint8_t jac[144];
        ref_msm_g1(bases[t][0].b.data(), scalars[t][0].bytes(), sizes[t], jac);
        ref_normalize_batch_g1(jac, 1, expected[t].b.data());
    }
    std::vector<int> okv(4, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < 4; t++)
        th.emplace_back([&, t] {
            int good = 1;
            for (int rep = 0; rep < 5; rep++)
                good &= into_affine<G1>(VariableBaseMSM<G1>::msm_bigint(bases[t], scalars[t])) == expected[t];
            okv[t] = good;
        });
    for (auto &x : th) x.join();
    for (int t = 0; t < 4; t++) CHECK(okv[t]);
}

// one process, several GPUs + the table-free fused update (SURVEY.md 8b "dg_msm_g1_sharded", row a16)
static void test_sharded_and_fused() {
    dg_shutdown();
    int nd = 1;
    try { init_devices({0, 1}); nd = 2; } catch (const Error &) { init_devices({0}); }
    CHECK(device_count() == nd);
    const size_t n = 5000;
    auto g = rand_g1(n);
    std::vector<Fr> e(n); for (auto &x : e) x = rand_fr();
    uint8_t jac[144]; G1Affine exp;
    ref_msm_g1(g[0].b.data(), e[0].bytes(), n, jac); ref_normalize_batch_g1(jac, 1, exp.b.data());
    CHECK(into_affine<G1>(msm_bigint_sharded(g, e)) == exp);
    {
        ShardedBasesG1 hb(g);
        CHECK(into_affine<G1>(hb.msm_bigint(e)) == exp);
        hb.precompute();
        CHECK(into_affine<G1>(hb.msm_bigint(e)) == exp);
        std::vector<Fr> half(e.begin(), e.begin() + n / 2);                 // prefix: the second shard goes empty
        ref_msm_g1(g[0].b.data(), e[0].bytes(), n / 2, jac); G1Affine exp2; ref_normalize_batch_g1(jac, 1, exp2.b.data());
        CHECK(into_affine<G1>(hb.msm_bigint(half)) == exp2);
    }
    CHECK(msm_bigint_sharded({}, {}).is_zero());
    // d_i * C_i + v_i * V == the 2-term MSM of the oracle
    const size_t m = 40;
    auto c = rand_g1(m); auto v = rand_g1(1)[0];
    std::vector<Fr> d(m), vf(m); for (size_t i = 0; i < m; i++) { d[i] = rand_fr(); vf[i] = rand_fr(); }
    auto out = batch_mul_add_same(c, d, v, vf);
    for (size_t i = 0; i < m; i++) {
        G1Affine two[2] = {c[i], v}; Fr sc[2] = {d[i], vf[i]};
        ref_msm_g1(two[0].b.data(), sc[0].bytes(), 2, jac); G1Affine ex; ref_normalize_batch_g1(jac, 1, ex.b.data());
        CHECK(out[i] == ex);
    }
    bool threw = false;
    try { d.pop_back(); batch_mul_add_same(c, d, v, vf); } catch (const Error &) { threw = true; }
    CHECK(threw);
}

// the device-chained prover on a squaring-chain circuit x_(i+1) = x_i^2: every MSM of the call equals
// the same MSM through VariableBaseMSM, and the h MSM equals msm_bigint(h_query, h) with h from dg_qap_h_from_abc on
// A w, B w, C w computed here with the host-side Fr
static std::array<uint8_t, 32> to_mont(const Fr &a, const Fr &r256) { Fr m = a * r256; std::array<uint8_t, 32> o; std::memcpy(o.data(), m.bytes(), 32); return o; }
static void test_chained_prover() {
    const size_t m = 1000, ninputs = 2, nvars = ninputs + m;           // variables: 1, x0 (instance), x1 .. xm (witness)
    Fr r256 = Fr::one();
    for (int i = 0; i < 256; i++) r256 = r256 + r256;                   // 2^256 mod r: Montgomery form = a * 2^256
    std::vector<Fr> w(nvars);
    w[0] = Fr::one(); w[1] = rand_fr();
    for (size_t i = 0; i < m; i++) w[2 + i] = w[1 + i] * w[1 + i];
    CsrMatrix A, B, C;
    for (size_t i = 0; i <= m; i++) { A.row_ptr.push_back((uint32_t)i); B.row_ptr.push_back((uint32_t)i); C.row_ptr.push_back((uint32_t)i); }
    for (size_t i = 0; i < m; i++) {
        A.col.push_back((uint32_t)(1 + i)); B.col.push_back((uint32_t)(1 + i)); C.col.push_back((uint32_t)(2 + i));
        A.coeff_mont.push_back(to_mont(Fr::one(), r256)); B.coeff_mont.push_back(to_mont(Fr::one(), r256)); C.coeff_mont.push_back(to_mont(Fr::one(), r256));
    }
    ResidentR1cs r1cs(A, B, C, m, ninputs, nvars);
    std::vector<std::array<uint8_t, 32>> wm(nvars);
    for (size_t i = 0; i < nvars; i++) wm[i] = to_mont(w[i], r256);
    size_t logd = 0; while ((size_t(1) << logd) < m + ninputs) logd++;
    const size_t D = size_t(1) << logd;
    auto hq = rand_g1(D - 1), aq = rand_g1(nvars), lq = rand_g1(m);
    auto bq = rand_g2(nvars);
    ResidentBases<G1> h_res(hq), a_res(aq), l_res(lq);
    ResidentBases<G2> b_res(bq);
    a_res.precompute();                                                   // one key through a resident table, the others plain
    auto out = groth16_prove_msms(r1cs, wm, h_res.handle, {{a_res.handle, 0, nvars, false}, {b_res.handle, 0, nvars, true}, {l_res.handle, ninputs, m, false}});
    CHECK(out.g1.size() == 2 && out.g2.size() == 1);
    CHECK(into_affine<G1>(out.g1[0]) == into_affine<G1>(VariableBaseMSM<G1>::msm_bigint(aq, w)));
    CHECK(into_affine<G2>(out.g2[0]) == into_affine<G2>(VariableBaseMSM<G2>::msm_bigint(bq, w)));
    std::vector<Fr> wit(w.begin() + ninputs, w.end());
    CHECK(into_affine<G1>(out.g1[1]) == into_affine<G1>(VariableBaseMSM<G1>::msm_bigint(lq, wit)));
    CHECK(into_affine<G1>(out.g1[1]) == into_affine<G1>(l_res.msm_bigint(wit)));
    // h: a_i = b_i = x_i, c_i = x_(i+1) on the constraint rows, then the instance variables in a
    std::vector<std::array<uint8_t, 32>> ea(D), eb(D), ec(D), h(D);
    const auto zero = to_mont(Fr::zero(), r256);
    for (size_t i = 0; i < D; i++) { ea[i] = zero; eb[i] = zero; ec[i] = zero; }
    for (size_t i = 0; i < m; i++) { ea[i] = wm[1 + i]; eb[i] = wm[1 + i]; ec[i] = wm[2 + i]; }
    for (size_t j = 0; j < ninputs; j++) ea[m + j] = wm[j];
    check(dg_qap_h_from_abc(ea[0].data(), eb[0].data(), ec[0].data(), (uint32_t)logd, h[0].data()));
    std::vector<Fr> hb(D);
    check(dg_fr_into_bigint(h[0].data(), D, (uint8_t *)hb[0].l.data()));
    CHECK(into_affine<G1>(out.h_acc) == into_affine<G1>(VariableBaseMSM<G1>::msm_bigint(hq, hb)));
    CHECK(!out.h_acc.is_zero());
}

int main() {
    init(0);
    test_wire_formats();
    test_concurrent_callers();
    test_fr();
    test_window_table_and_msm();
    test_pairing_randomize();
    test_mult_checker();
    test_chained_prover();
    test_sharded_and_fused();
    std::printf("cpp host api ok, launches=%llu\n", (unsigned long long)dg_launch_count());
    return 0;
}
