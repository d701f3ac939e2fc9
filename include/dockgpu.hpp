// dockgpu.hpp -- C++ host-side mirror of the reference's interface for the hot path, over the C ABI
// of dockgpu.h.  The reference is Rust; no Rust toolchain exists in the build image, so the layer
// a Rust maintainer would write as the `dock_gpu` glue crate is expressed here in
// C++ with the SAME names, argument meaning and error behaviour:
//
//   utils::msm::WindowTable<G>, multiply_field_elems_with_same_group_elem   utils/src/msm.rs:8-62
//   ark_ec::VariableBaseMSM::{msm, msm_unchecked, msm_bigint}                SURVEY.md Appendix B
//   CurveGroup::normalize_batch, AffineRepr::mul_bigint (batched)
//   Pairing::{multi_miller_loop, final_exponentiation, multi_pairing, pairing}
//   utils::randomized_pairing_check::RandomizedPairingChecker               utils/src/randomized_pairing_check.rs:24-215
//   utils::randomized_mult_checker::RandomizedMultChecker                   utils/src/randomized_mult_checker.rs:21-126
//
// All curve arithmetic runs on the GPU through libdockgpu.so; the only arithmetic done here is
// scalar-field (mod r) bookkeeping of the checkers (powers of the random challenge), exactly
// what the reference does on the CPU around its arkworks calls.  Header-only; link -ldockgpu.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "dockgpu.h"

namespace dock_gpu {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("libdockgpu error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int32_t rc) {
    if (rc != DG_OK) {
        char buf[512];
        dg_last_error(buf, sizeof buf);
        throw Error(rc, buf);
    }
}
inline void init(int device = -1) { check(dg_init(device)); }

// ---- scalar field Fr (canonical little-endian 4 x u64 = ark BigInt<4>) ---------------------------
struct Fr {
    std::array<uint64_t, 4> l{};
    static constexpr std::array<uint64_t, 4> MODULUS = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                                                        0x73eda753299d7d48ULL};
    static Fr from_u64(uint64_t v) { Fr r; r.l[0] = v; return r; }
    static Fr one() { return from_u64(1); }
    static Fr zero() { return Fr{}; }
    bool is_zero() const { return !(l[0] | l[1] | l[2] | l[3]); }
    bool operator==(const Fr &o) const { return l == o.l; }
    const uint8_t *bytes() const { return reinterpret_cast<const uint8_t *>(l.data()); }     // into_bigint()
    static bool geq(const std::array<uint64_t, 4> &a, const std::array<uint64_t, 4> &b) {
        for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
        return true;
    }
    static void sub_mod(std::array<uint64_t, 4> &a) {
        unsigned __int128 br = 0;
        for (int i = 0; i < 4; i++) { unsigned __int128 t = (unsigned __int128)a[i] - MODULUS[i] - br; a[i] = (uint64_t)t; br = (t >> 64) & 1; }
    }
    Fr operator+(const Fr &o) const {
        Fr r; unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++) { c += (unsigned __int128)l[i] + o.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        if (c || geq(r.l, MODULUS)) sub_mod(r.l);       // r < 2^255 so a + b < 2^256: c is always 0
        return r;
    }
    Fr operator-() const {
        if (is_zero()) return *this;
        Fr r; unsigned __int128 br = 0;
        for (int i = 0; i < 4; i++) { unsigned __int128 t = (unsigned __int128)MODULUS[i] - l[i] - br; r.l[i] = (uint64_t)t; br = (t >> 64) & 1; }
        return r;
    }
    Fr operator-(const Fr &o) const { return *this + (-o); }
    Fr operator*(const Fr &o) const {                    // schoolbook 4x4 then bitwise reduction (bookkeeping only)
        uint64_t t[8] = {0};
        for (int i = 0; i < 4; i++) {
            unsigned __int128 c = 0;
            for (int j = 0; j < 4; j++) { c += (unsigned __int128)l[i] * o.l[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
            t[i + 4] = (uint64_t)c;
        }
        Fr r;                                             // r = t mod MODULUS, MSB-first shift-and-subtract
        for (int bit = 511; bit >= 0; bit--) {
            uint64_t top = r.l[3] >> 63;
            for (int i = 3; i > 0; i--) r.l[i] = (r.l[i] << 1) | (r.l[i - 1] >> 63);
            r.l[0] = (r.l[0] << 1) | ((t[bit >> 6] >> (bit & 63)) & 1);
            if (top || geq(r.l, MODULUS)) sub_mod(r.l);
        }
        return r;
    }
};

// ---- group element records (packed Montgomery limbs, see dockgpu.h) -------------------------------
template <size_t N> struct Rec {
    std::array<uint8_t, N> b{};
    bool all_zero() const { for (auto x : b) if (x) return false; return true; }
    bool operator==(const Rec &o) const { return b == o.b; }
    bool operator<(const Rec &o) const { return b < o.b; }
};
struct G1 {
    static constexpr bool IS_G2 = false;
    struct Affine : Rec<96> { bool is_zero() const { return all_zero(); } };                      // identity = all-zero record
    struct Projective : Rec<144> { bool is_zero() const { for (size_t i = 96; i < 144; i++) if (b[i]) return false; return true; } };
    static int32_t msm(uint64_t h, const uint8_t *p, const uint8_t *s, size_t n, uint8_t *o) { return dg_msm_g1(h, p, s, n, o); }
    static int32_t table(const uint8_t *p, size_t n, uint64_t *h) { return dg_fixed_base_table_g1(p, n, h); }
    static int32_t mul_many(uint64_t h, const uint8_t *s, size_t m, uint8_t *o) { return dg_fixed_base_mul_many_g1(h, s, m, o); }
    static int32_t batch_mul(const uint8_t *p, const uint8_t *s, size_t m, uint8_t *o) { return dg_batch_mul_g1(p, s, m, o); }
    static int32_t normalize(const uint8_t *j, size_t m, uint8_t *o) { return dg_normalize_batch_g1(j, m, o); }
    static int32_t serialize(const uint8_t *a, size_t n, int c, uint8_t *o) { return dg_g1_serialize(a, n, c, o); }
    static int32_t deserialize(const uint8_t *i, size_t n, int c, int v, uint8_t *o, uint8_t *st, size_t *bad) { return dg_g1_deserialize(i, n, c, v, o, st, bad); }
};
struct G2 {
    static constexpr bool IS_G2 = true;
    struct Affine : Rec<192> { bool is_zero() const { return all_zero(); } };
    struct Projective : Rec<288> { bool is_zero() const { for (size_t i = 192; i < 288; i++) if (b[i]) return false; return true; } };
    static int32_t msm(uint64_t h, const uint8_t *p, const uint8_t *s, size_t n, uint8_t *o) { return dg_msm_g2(h, p, s, n, o); }
    static int32_t table(const uint8_t *p, size_t n, uint64_t *h) { return dg_fixed_base_table_g2(p, n, h); }
    static int32_t mul_many(uint64_t h, const uint8_t *s, size_t m, uint8_t *o) { return dg_fixed_base_mul_many_g2(h, s, m, o); }
    static int32_t batch_mul(const uint8_t *p, const uint8_t *s, size_t m, uint8_t *o) { return dg_batch_mul_g2(p, s, m, o); }
    static int32_t normalize(const uint8_t *j, size_t m, uint8_t *o) { return dg_normalize_batch_g2(j, m, o); }
    static int32_t serialize(const uint8_t *a, size_t n, int c, uint8_t *o) { return dg_g2_serialize(a, n, c, o); }
    static int32_t deserialize(const uint8_t *i, size_t n, int c, int v, uint8_t *o, uint8_t *st, size_t *bad) { return dg_g2_deserialize(i, n, c, v, o, st, bad); }
};
using G1Affine = G1::Affine;
using G1Projective = G1::Projective;
using G2Affine = G2::Affine;
using G2Projective = G2::Projective;
struct Fp12 : Rec<576> {};
using MillerLoopOutput = Fp12;      // arkworks newtypes over Fp12
using PairingOutput = Fp12;

static_assert(sizeof(G1Affine) == 96 && sizeof(G1Projective) == 144 && sizeof(G2Affine) == 192 && sizeof(Fr) == 32, "packed records");

// CurveGroup::normalize_batch
template <class G> std::vector<typename G::Affine> normalize_batch(const std::vector<typename G::Projective> &v) {
    std::vector<typename G::Affine> out(v.size());
    if (!v.empty()) check(G::normalize(v[0].b.data(), v.size(), out[0].b.data()));
    return out;
}
template <class G> typename G::Affine into_affine(const typename G::Projective &p) { return normalize_batch<G>({p})[0]; }

// cfg_iter!(points).zip(scalars).map(|(p, s)| p.mul_bigint(s))
template <class G> std::vector<typename G::Projective> mul_bigint_batch(const std::vector<typename G::Affine> &p, const std::vector<Fr> &s) {
    size_t m = p.size() < s.size() ? p.size() : s.size();
    std::vector<typename G::Projective> out(m);
    if (m) check(G::batch_mul(p[0].b.data(), s[0].bytes(), m, out[0].b.data()));
    return out;
}

// ---- ark_serialize::{CanonicalSerialize, CanonicalDeserialize} for vectors of points -------------------------
// (utils/src/serde_utils.rs:13-33; the body after the u64 length prefix of a Vec<Affine>)
enum class Compress { Yes, No };
enum class Validate { Yes, No };
template <class G> std::vector<uint8_t> serialize_points(const std::vector<typename G::Affine> &v, Compress c = Compress::Yes) {
    const size_t rec = sizeof(typename G::Affine) / (c == Compress::Yes ? 2 : 1);
    std::vector<uint8_t> out(rec * v.size());
    if (!v.empty()) check(G::serialize(v[0].b.data(), v.size(), c == Compress::Yes, out.data()));
    return out;
}
// Err(SerializationError::InvalidData) as std::nullopt: any element malformed, off the curve or (Validate::Yes) outside the subgroup
template <class G>
std::optional<std::vector<typename G::Affine>> deserialize_points(const std::vector<uint8_t> &bytes, Compress c = Compress::Yes,
                                                                  Validate val = Validate::Yes) {
    const size_t rec = sizeof(typename G::Affine) / (c == Compress::Yes ? 2 : 1);
    if (bytes.size() % rec) return std::nullopt;
    std::vector<typename G::Affine> out(bytes.size() / rec);
    size_t bad = 0;
    if (!out.empty()) check(G::deserialize(bytes.data(), out.size(), c == Compress::Yes, val == Validate::Yes, out[0].b.data(), nullptr, &bad));
    if (bad) return std::nullopt;
    return out;
}

// ---- ark_ec::VariableBaseMSM ------------------------------------------------------------------------
template <class G> struct MsmResult {            // Result<G, usize>
    bool ok;
    typename G::Projective value;
    size_t err_min_len;
};
template <class G> struct VariableBaseMSM {
    using Affine = typename G::Affine;
    using Projective = typename G::Projective;
    // truncates to the shorter input, never fails
    static Projective msm_bigint(const std::vector<Affine> &bases, const std::vector<Fr> &bigints) {
        size_t n = bases.size() < bigints.size() ? bases.size() : bigints.size();
        Projective out;
        check(G::msm(0, n ? bases[0].b.data() : nullptr, n ? bigints[0].bytes() : nullptr, n, out.b.data()));
        return out;
    }
    // into_bigint() on every scalar (Fr here is already canonical), then msm_bigint
    static Projective msm_unchecked(const std::vector<Affine> &bases, const std::vector<Fr> &scalars) { return msm_bigint(bases, scalars); }
    // Err(min_len) when the lengths differ
    static MsmResult<G> msm(const std::vector<Affine> &bases, const std::vector<Fr> &scalars) {
        if (bases.size() != scalars.size()) return {false, Projective{}, bases.size() < scalars.size() ? bases.size() : scalars.size()};
        return {true, msm_unchecked(bases, scalars), 0};
    }
};

// ---- one process, several GPUs (dg_init_devices / dg_msm_*_sharded) ------------------------------
// The reference is one process with rayon threads; this is the entry point its glue would call for the large MSMs.
inline void init_devices(const std::vector<int> &devices) {
    std::vector<int32_t> d(devices.begin(), devices.end());
    check(dg_init_devices(d.data(), (int32_t)d.size()));
}
inline int device_count() { int32_t n = 0; check(dg_device_count(&n)); return n; }
// Bases resident on every device of init_devices, split by contiguous ranges (RAII over the sharded handle)
class ShardedBasesG1 {
  public:
    explicit ShardedBasesG1(const std::vector<G1::Affine> &bases) : n(bases.size()) {
        check(dg_bases_upload_g1_sharded(bases[0].b.data(), bases.size(), &handle));
    }
    ~ShardedBasesG1() { if (handle) dg_bases_free(handle); }
    ShardedBasesG1(const ShardedBasesG1 &) = delete;
    ShardedBasesG1 &operator=(const ShardedBasesG1 &) = delete;
    void precompute(int window_bits = 0) { check(dg_bases_precompute(handle, window_bits)); }
    // msm_bigint over a prefix of the resident bases: truncates to the shorter side like arkworks
    G1::Projective msm_bigint(const std::vector<Fr> &bigints) const {
        size_t k = bigints.size() < n ? bigints.size() : n;
        G1::Projective out;
        check(dg_msm_g1_sharded(handle, nullptr, k ? bigints[0].bytes() : nullptr, k, out.b.data()));
        return out;
    }
    uint64_t handle = 0;
    size_t n;
};
inline G1::Projective msm_bigint_sharded(const std::vector<G1::Affine> &bases, const std::vector<Fr> &bigints) {
    size_t n = bases.size() < bigints.size() ? bases.size() : bigints.size();
    G1::Projective out;
    check(dg_msm_g1_sharded(0, n ? bases[0].b.data() : nullptr, n ? bigints[0].bytes() : nullptr, n, out.b.data()));
    return out;
}

// ---- resident keys and the device-chained LegoGroth16 prover --------------------------------------
// One query vector of a proving key resident on the device (dg_bases_upload_*; RAII).  precompute() trades memory for the
// 2^(ck)-multiples table of dg_bases_precompute.
template <class G> class ResidentBases {
  public:
    explicit ResidentBases(const std::vector<typename G::Affine> &bases) : n(bases.size()) {
        if (!G::IS_G2) check(dg_bases_upload_g1(bases[0].b.data(), n, &handle));
        else check(dg_bases_upload_g2(bases[0].b.data(), n, &handle));
    }
    ~ResidentBases() { if (handle) dg_bases_free(handle); }
    ResidentBases(const ResidentBases &) = delete;
    ResidentBases &operator=(const ResidentBases &) = delete;
    void precompute(int window_bits = 0) { check(dg_bases_precompute(handle, window_bits)); }
    typename G::Projective msm_bigint(const std::vector<Fr> &bigints) const {
        size_t k = bigints.size() < n ? bigints.size() : n;
        typename G::Projective out;
        check(G::msm(handle, nullptr, k ? bigints[0].bytes() : nullptr, k, out.b.data()));
        return out;
    }
    uint64_t handle = 0;
    size_t n;
};
// ConstraintMatrices (A, B, C in CSR form, coefficients as Montgomery Fr records) resident on the device
struct CsrMatrix {
    std::vector<uint32_t> row_ptr, col;
    std::vector<std::array<uint8_t, 32>> coeff_mont;
};
class ResidentR1cs {
  public:
    ResidentR1cs(const CsrMatrix &a, const CsrMatrix &b, const CsrMatrix &c, size_t num_constraints, size_t num_inputs, size_t num_vars)
        : num_vars(num_vars) {
        const uint32_t *rp[3] = {a.row_ptr.data(), b.row_ptr.data(), c.row_ptr.data()};
        const uint32_t *cl[3] = {a.col.data(), b.col.data(), c.col.data()};
        const uint8_t *co[3] = {a.coeff_mont.empty() ? nullptr : a.coeff_mont[0].data(), b.coeff_mont.empty() ? nullptr : b.coeff_mont[0].data(),
                                c.coeff_mont.empty() ? nullptr : c.coeff_mont[0].data()};
        check(dg_r1cs_upload(rp, cl, co, num_constraints, num_inputs, num_vars, &handle));
    }
    ~ResidentR1cs() { if (handle) dg_r1cs_free(handle); }
    ResidentR1cs(const ResidentR1cs &) = delete;
    ResidentR1cs &operator=(const ResidentR1cs &) = delete;
    uint64_t handle = 0;
    size_t num_vars;
};
struct ProveJob { uint64_t bases; size_t offset, count; bool g2; };
struct ProveMsms {
    G1::Projective h_acc;                              // msm_bigint(h_query, h)
    std::vector<G1::Projective> g1;                    // results of the G1 jobs, in job order
    std::vector<G2::Projective> g2;                    // results of the G2 jobs, in job order
};
// witness_map_from_matrices + into_bigint + every MSM of create_proof_and_committed_witnesses_with_assignment
// (legogroth16/src/prover.rs:267-383) in one call; full_assignment_mont = instance then witness variables, Montgomery form
inline ProveMsms groth16_prove_msms(const ResidentR1cs &r1cs, const std::vector<std::array<uint8_t, 32>> &full_assignment_mont,
                                    uint64_t h_query, const std::vector<ProveJob> &jobs) {
    std::vector<uint64_t> jb, jo, jc;
    for (auto &j : jobs) { jb.push_back(j.bases); jo.push_back(j.offset); jc.push_back(j.count); }
    std::vector<uint8_t> raw(288 * (jobs.size() + 1));
    ProveMsms out;
    check(dg_groth16_prove_msms(r1cs.handle, full_assignment_mont[0].data(), full_assignment_mont.size(), h_query, jb.data(), jo.data(),
                                jc.data(), jobs.size(), out.h_acc.b.data(), raw.data(), nullptr));
    for (size_t j = 0; j < jobs.size(); j++) {
        if (jobs[j].g2) { G2::Projective p; std::copy(raw.begin() + 288 * j, raw.begin() + 288 * j + 288, p.b.begin()); out.g2.push_back(p); }
        else { G1::Projective p; std::copy(raw.begin() + 288 * j, raw.begin() + 288 * j + 144, p.b.begin()); out.g1.push_back(p); }
    }
    return out;
}

// ---- fused accumulator witness update (vb_accumulator/src/witness.rs:269-284): d_i * C_i + v_i * V, normalised ------
inline std::vector<G1::Affine> batch_mul_add_same(const std::vector<G1::Affine> &c, const std::vector<Fr> &d, const G1::Affine &v,
                                                  const std::vector<Fr> &vf) {
    if (c.size() != d.size() || d.size() != vf.size()) throw Error(DG_ERR_BAD_ARG, "NeedSameNoOfElementsAndWitnesses");
    std::vector<G1::Affine> out(c.size());
    if (!c.empty()) check(dg_batch_mul_add_same_g1(c[0].b.data(), d[0].bytes(), v.b.data(), vf[0].bytes(), c.size(), out[0].b.data()));
    return out;
}

// ---- utils::msm::WindowTable -------------------------------------------------------------------------
inline size_t ln_without_floats(size_t a) { size_t l = 0; while ((size_t(1) << l) < a) l++; return l * 69 / 100; }
template <class G> class WindowTable {
  public:
    size_t scalar_size, window_size, num_windows;
    // `num_multiplications` is a performance hint only (utils/src/msm.rs:16-30)
    WindowTable(size_t num_multiplications, const typename G::Projective &group_elem) {
        typename G::Affine a = into_affine<G>(group_elem);
        check(G::table(a.b.data(), num_multiplications, &handle_));
        int32_t w = 0, nw = 0, g2 = 0;
        check(dg_fixed_base_table_info(handle_, &w, &nw, &g2));
        scalar_size = 255;
        window_size = (size_t)w;
        num_windows = (size_t)nw;
    }
    WindowTable(const WindowTable &) = delete;
    WindowTable &operator=(const WindowTable &) = delete;
    ~WindowTable() { if (handle_) dg_fixed_base_table_free(handle_); }
    static size_t window_size_for(size_t num_multiplications) { return num_multiplications < 32 ? 3 : ln_without_floats(num_multiplications); }
    typename G::Projective multiply(const Fr &element) const { return multiply_many({element})[0]; }
    std::vector<typename G::Projective> multiply_many(const std::vector<Fr> &elements) const {
        std::vector<typename G::Projective> out(elements.size());
        if (!elements.empty()) check(G::mul_many(handle_, elements[0].bytes(), elements.size(), out[0].b.data()));
        return out;
    }
    uint64_t handle() const { return handle_; }

  private:
    uint64_t handle_ = 0;
};
template <class G> typename G::Projective operator*(const WindowTable<G> &t, const Fr &s) { return t.multiply(s); }

template <class G>
std::vector<typename G::Projective> multiply_field_elems_with_same_group_elem(const typename G::Projective &group_elem, const std::vector<Fr> &elements) {
    WindowTable<G> table(elements.size(), group_elem);
    return table.multiply_many(elements);
}

// ---- ark_ec::pairing::Pairing for Bls12_381 -------------------------------------------------------------
struct Bls12_381 {
    static MillerLoopOutput multi_miller_loop(const std::vector<G1Affine> &a, const std::vector<G2Affine> &b) {
        size_t k = a.size() < b.size() ? a.size() : b.size();
        MillerLoopOutput out;
        check(dg_multi_miller_loop(k ? a[0].b.data() : nullptr, k ? b[0].b.data() : nullptr, k, out.b.data()));
        return out;
    }
    static std::optional<PairingOutput> final_exponentiation(const MillerLoopOutput &f) {
        PairingOutput out;
        int32_t some = 0;
        check(dg_final_exponentiation(f.b.data(), out.b.data(), &some));
        if (!some) return std::nullopt;
        return out;
    }
    static PairingOutput multi_pairing(const std::vector<G1Affine> &a, const std::vector<G2Affine> &b) {
        size_t k = a.size() < b.size() ? a.size() : b.size();
        PairingOutput out;
        check(dg_multi_pairing(k ? a[0].b.data() : nullptr, k ? b[0].b.data() : nullptr, k, out.b.data()));
        return out;
    }
    static PairingOutput pairing(const G1Affine &p, const G2Affine &q) { return multi_pairing({p}, {q}); }
    static MillerLoopOutput miller_loop(const G1Affine &p, const G2Affine &q) { return multi_miller_loop({p}, {q}); }
};
inline Fp12 fp12_mul(const Fp12 &x, const Fp12 &y) { Fp12 o; check(dg_fp12_mul(x.b.data(), y.b.data(), o.b.data())); return o; }
inline PairingOutput gt_mul_bigint(const PairingOutput &x, const Fr &m) { PairingOutput o; check(dg_gt_pow(x.b.data(), m.bytes(), o.b.data())); return o; }
inline Fp12 fp12_one() { return Bls12_381::multi_pairing({}, {}); }

// ---- utils::randomized_pairing_check::RandomizedPairingChecker --------------------------------------------
class RandomizedPairingChecker {
  public:
    RandomizedPairingChecker(const Fr &random, bool lazy)
        : left(fp12_one()), right(left), lazy(lazy), random(random), current_random(Fr::one()) {}
    void add_sources_and_target(const G1Affine &a, const G2Affine &b, const PairingOutput &out) {
        add_multiple_sources_and_target({a}, {b}, out);
    }
    void add_multiple_sources_and_target(const std::vector<G1Affine> &a, const std::vector<G2Affine> &b, const PairingOutput &out) {
        add_multiple_sources_and_target_with_laziness_choice(a, b, out, lazy);
    }
    void add_multiple_sources(const std::vector<G1Affine> &a, const std::vector<G2Affine> &b, const std::vector<G1Affine> &c,
                              const std::vector<G2Affine> &d) {
        add_multiple_sources_with_laziness_choice(a, b, c, d, lazy);
    }
    void add_sources(const G1Affine &a, const G2Affine &b, const G1Affine &c, const G2Affine &d) {
        add_multiple_sources_with_laziness_choice({a}, {b}, {c}, {d}, lazy);
    }
    void add_multiple_sources_and_target_with_laziness_choice(const std::vector<G1Affine> &a, const std::vector<G2Affine> &b,
                                                              const PairingOutput &out, bool lazy_) {
        const Fr m = current_random;
        std::vector<G1Affine> a_m = scaled(a, m);
        if (lazy_) {
            pending_g1.insert(pending_g1.end(), a_m.begin(), a_m.end());
            pending_g2.insert(pending_g2.end(), b.begin(), b.end());
        } else {
            left = fp12_mul(left, Bls12_381::multi_miller_loop(a_m, b));
        }
        right = fp12_mul(right, gt_mul_bigint(out, m));           // right += out * m
        current_random = current_random * random;
    }
    void add_multiple_sources_with_laziness_choice(const std::vector<G1Affine> &a, const std::vector<G2Affine> &b,
                                                   const std::vector<G1Affine> &c, const std::vector<G2Affine> &d, bool lazy_) {
        const Fr m = current_random;
        std::vector<G1Affine> a_m = scaled(a, m), c_m = scaled(c, -m);     // -(c * m) == c * (r - m)
        if (lazy_) {
            pending_g1.insert(pending_g1.end(), a_m.begin(), a_m.end());
            pending_g2.insert(pending_g2.end(), b.begin(), b.end());
            pending_g1.insert(pending_g1.end(), c_m.begin(), c_m.end());
            pending_g2.insert(pending_g2.end(), d.begin(), d.end());
        } else {
            left = fp12_mul(left, Bls12_381::multi_miller_loop(a_m, b));
            left = fp12_mul(left, Bls12_381::multi_miller_loop(c_m, d));
        }
        current_random = current_random * random;
    }
    bool verify() const {
        MillerLoopOutput l = left;
        if (!pending_g1.empty()) l = fp12_mul(Bls12_381::multi_miller_loop(pending_g1, pending_g2), left);
        auto fe = Bls12_381::final_exponentiation(l);
        if (!fe) throw Error(DG_ERR_BAD_ARG, "final_exponentiation of zero");     // reference unwraps
        return *fe == right;
    }

  private:
    static std::vector<G1Affine> scaled(const std::vector<G1Affine> &pts, const Fr &m) {
        std::vector<Fr> s(pts.size(), m);
        return normalize_batch<G1>(mul_bigint_batch<G1>(pts, s));
    }
    MillerLoopOutput left;
    PairingOutput right;
    bool lazy;
    std::vector<G1Affine> pending_g1;
    std::vector<G2Affine> pending_g2;
    Fr random, current_random;
};

// ---- utils::randomized_mult_checker::RandomizedMultChecker --------------------------------------------------
template <class G> class RandomizedMultChecker {
  public:
    using Affine = typename G::Affine;
    explicit RandomizedMultChecker(const Fr &random) : random(random), current_random(Fr::one()) {}
    void add_1(const Affine &p, const Fr &s, const Affine &t) {
        add(p, current_random * s); add(t, -current_random); current_random = current_random * random;
    }
    void add_2(const Affine &p1, const Fr &s1, const Affine &p2, const Fr &s2, const Affine &t) {
        add(p1, current_random * s1); add(p2, current_random * s2); add(t, -current_random); current_random = current_random * random;
    }
    void add_3(const Affine &p1, const Fr &s1, const Affine &p2, const Fr &s2, const Affine &p3, const Fr &s3, const Affine &t) {
        add(p1, current_random * s1); add(p2, current_random * s2); add(p3, current_random * s3); add(t, -current_random);
        current_random = current_random * random;
    }
    void add_many(const std::vector<Affine> &a, const std::vector<Fr> &b, const Affine &t) {
        size_t n = a.size() < b.size() ? a.size() : b.size();
        for (size_t i = 0; i < n; i++) add(a[i], current_random * b[i]);
        add(t, -current_random);
        current_random = current_random * random;
    }
    size_t len() const { return args.size(); }
    // one MSM over the deduplicated points must be the identity
    bool verify() const {
        std::vector<Affine> points; std::vector<Fr> scalars;
        for (auto &kv : args) { points.push_back(kv.second.second); scalars.push_back(kv.second.first); }
        return VariableBaseMSM<G>::msm_unchecked(points, scalars).is_zero();
    }

  private:
    void add(const Affine &p, const Fr &s) {
        if (p.is_zero()) return;                               // the identity does not affect the result
        std::vector<uint8_t> x(p.b.begin(), p.b.begin() + p.b.size() / 2);
        auto it = args.find(x);
        if (it == args.end()) args.emplace(std::move(x), std::make_pair(s, p));
        else if (it->second.second == p) it->second.first = it->second.first + s;
        else it->second.first = it->second.first - s;          // same x, opposite y: the entry holds -p
    }
    std::map<std::vector<uint8_t>, std::pair<Fr, Affine>> args;    // x-coordinate -> (scalar, point)
    Fr random, current_random;
};

}  // namespace dock_gpu
