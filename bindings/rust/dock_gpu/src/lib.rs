//! Safe entry points over libdockgpu.so for the BLS12-381 hot path of docknetwork/crypto.
//!
//! This is the crate INTEGRATION.md section 1 describes.  It lives outside `legogroth16` / `dock_crypto_utils` because
//! those crates are `#![forbid(unsafe_code)]`.  The `try_*` functions return `None` whenever the GPU does not apply
//! (library not initialised, input below the size threshold, an error code from the library), so the patched arkworks
//! call site falls through to its CPU body: the library itself never falls back.
//!
//! Not compiled in this repository's build image (no Rust toolchain there); `ffi.rs` is generated from the C header.
//!
//! Record layouts (include/dockgpu.h): field elements are their Montgomery limbs, little-endian, exactly ark-ff's
//! `Fp.0 .0`; an affine G1 point is x || y (96 B, the identity is all zeros), a projective one x || y || z (144 B, ark's
//! Jacobian `Projective`); G2 doubles every size with c0 before c1; scalars are the canonical `BigInt<4>` (32 B).
pub mod ffi;

use ark_bls12_381::{Fq, Fq2, Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use ark_ff::{BigInt, PrimeField};
use std::sync::OnceLock;

/// Below this many terms the CPU path of arkworks wins (kernel launches + PCIe latency).
pub const MSM_THRESHOLD: usize = 1 << 10;

static READY: OnceLock<bool> = OnceLock::new();

/// `dg_init_devices` over every visible GPU, once per process; false when there is no usable device.
pub fn init() -> bool {
    *READY.get_or_init(|| unsafe {
        let mut n = 0i32;
        if ffi::dg_device_count(&mut n) != ffi::DG_OK || n <= 0 {
            return false;
        }
        let devs: Vec<i32> = (0..n).collect();
        ffi::dg_init_devices(devs.as_ptr(), n) == ffi::DG_OK
    })
}

pub fn last_error() -> String {
    let mut buf = vec![0u8; 512];
    unsafe { ffi::dg_last_error(buf.as_mut_ptr() as *mut _, buf.len()) };
    let end = buf.iter().position(|&b| b == 0).unwrap_or(buf.len());
    String::from_utf8_lossy(&buf[..end]).into_owned()
}

fn put_fq(f: &Fq, out: &mut [u8]) {
    for (i, l) in f.0 .0.iter().enumerate() {
        out[8 * i..8 * i + 8].copy_from_slice(&l.to_le_bytes());
    }
}
fn get_fq(b: &[u8]) -> Fq {
    Fq::new_unchecked(BigInt::<6>(core::array::from_fn(|i| u64::from_le_bytes(b[8 * i..8 * i + 8].try_into().unwrap()))))
}

pub fn pack_g1(points: &[G1Affine]) -> Vec<u8> {
    let mut out = vec![0u8; 96 * points.len()];
    for (p, rec) in points.iter().zip(out.chunks_exact_mut(96)) {
        if !p.infinity {
            put_fq(&p.x, &mut rec[..48]);
            put_fq(&p.y, &mut rec[48..]);
        }
    }
    out
}
pub fn pack_g2(points: &[G2Affine]) -> Vec<u8> {
    let mut out = vec![0u8; 192 * points.len()];
    for (p, rec) in points.iter().zip(out.chunks_exact_mut(192)) {
        if !p.infinity {
            put_fq(&p.x.c0, &mut rec[..48]);
            put_fq(&p.x.c1, &mut rec[48..96]);
            put_fq(&p.y.c0, &mut rec[96..144]);
            put_fq(&p.y.c1, &mut rec[144..]);
        }
    }
    out
}
pub fn pack_bigints(scalars: &[<Fr as PrimeField>::BigInt]) -> Vec<u8> {
    let mut out = vec![0u8; 32 * scalars.len()];
    for (s, rec) in scalars.iter().zip(out.chunks_exact_mut(32)) {
        for (i, l) in s.0.iter().enumerate() {
            rec[8 * i..8 * i + 8].copy_from_slice(&l.to_le_bytes());
        }
    }
    out
}
pub fn unpack_g1_projective(b: &[u8]) -> G1Projective {
    G1Projective::new_unchecked(get_fq(&b[..48]), get_fq(&b[48..96]), get_fq(&b[96..144]))
}
pub fn unpack_g2_projective(b: &[u8]) -> G2Projective {
    let f2 = |o: usize| Fq2::new(get_fq(&b[o..o + 48]), get_fq(&b[o + 48..o + 96]));
    G2Projective::new_unchecked(f2(0), f2(96), f2(192))
}

/// `VariableBaseMSM::msm_bigint` for G1 (ark-ec scalar_mul/variable_base/mod.rs); truncates to the shorter side like ark.
pub fn try_msm_bigint_g1(bases: &[G1Affine], bigints: &[<Fr as PrimeField>::BigInt]) -> Option<G1Projective> {
    let n = bases.len().min(bigints.len());
    if n < MSM_THRESHOLD || !init() {
        return None;
    }
    let (b, s) = (pack_g1(&bases[..n]), pack_bigints(&bigints[..n]));
    let mut out = [0u8; 144];
    // one call from the calling rayon thread; with several GPUs the library scatters base ranges itself
    let rc = unsafe { ffi::dg_msm_g1_sharded(0, b.as_ptr(), s.as_ptr(), n, out.as_mut_ptr()) };
    (rc == ffi::DG_OK).then(|| unpack_g1_projective(&out))
}
pub fn try_msm_bigint_g2(bases: &[G2Affine], bigints: &[<Fr as PrimeField>::BigInt]) -> Option<G2Projective> {
    let n = bases.len().min(bigints.len());
    if n < MSM_THRESHOLD || !init() {
        return None;
    }
    let (b, s) = (pack_g2(&bases[..n]), pack_bigints(&bigints[..n]));
    let mut out = [0u8; 288];
    let rc = unsafe { ffi::dg_msm_g2_sharded(0, b.as_ptr(), s.as_ptr(), n, out.as_mut_ptr()) };
    (rc == ffi::DG_OK).then(|| unpack_g2_projective(&out))
}

/// Bases that stay on the device across calls (proving keys, signature parameters): uploaded once, freed on drop.
pub struct ResidentG1 {
    handle: u64,
    len: usize,
}
impl ResidentG1 {
    pub fn upload(bases: &[G1Affine]) -> Option<Self> {
        if bases.is_empty() || !init() {
            return None;
        }
        let b = pack_g1(bases);
        let mut handle = 0u64;
        let rc = unsafe { ffi::dg_bases_upload_g1_sharded(b.as_ptr(), bases.len(), &mut handle) };
        (rc == ffi::DG_OK).then_some(Self { handle, len: bases.len() })
    }
    /// Optional table of 2^(c k) multiples for bases reused by many MSMs (`window_bits` = 0: the library's default).
    pub fn precompute(&self, window_bits: i32) -> bool {
        unsafe { ffi::dg_bases_precompute(self.handle, window_bits) == ffi::DG_OK }
    }
    pub fn msm_bigint(&self, bigints: &[<Fr as PrimeField>::BigInt]) -> Option<G1Projective> {
        let n = self.len.min(bigints.len());
        let s = pack_bigints(&bigints[..n]);
        let mut out = [0u8; 144];
        let rc = unsafe { ffi::dg_msm_g1_sharded(self.handle, std::ptr::null(), s.as_ptr(), n, out.as_mut_ptr()) };
        (rc == ffi::DG_OK).then(|| unpack_g1_projective(&out))
    }
}
impl Drop for ResidentG1 {
    fn drop(&mut self) {
        unsafe { ffi::dg_bases_free(self.handle) };
    }
}

/// `utils::msm::WindowTable<G1Projective>` (utils/src/msm.rs:8-45): `new` / `multiply_many`.
pub struct WindowTableG1 {
    handle: u64,
}
impl WindowTableG1 {
    pub fn new(num_multiplications: usize, group_elem: &G1Affine) -> Option<Self> {
        if !init() {
            return None;
        }
        let p = pack_g1(core::slice::from_ref(group_elem));
        let mut handle = 0u64;
        let rc = unsafe { ffi::dg_fixed_base_table_g1(p.as_ptr(), num_multiplications, &mut handle) };
        (rc == ffi::DG_OK).then_some(Self { handle })
    }
    pub fn multiply_many(&self, elements: &[Fr]) -> Option<Vec<G1Projective>> {
        let big: Vec<_> = elements.iter().map(|e| e.into_bigint()).collect();
        let s = pack_bigints(&big);
        let mut out = vec![0u8; 144 * elements.len()];
        let rc = unsafe { ffi::dg_fixed_base_mul_many_g1(self.handle, s.as_ptr(), elements.len(), out.as_mut_ptr()) };
        (rc == ffi::DG_OK).then(|| out.chunks_exact(144).map(unpack_g1_projective).collect())
    }
}
impl Drop for WindowTableG1 {
    fn drop(&mut self) {
        unsafe { ffi::dg_fixed_base_table_free(self.handle) };
    }
}

/// `Bls12_381::multi_pairing(a, b).is_zero()` on the target group, i.e. prod e(a_i, b_i) == 1: the check behind
/// `RandomizedPairingChecker::verify` (utils/src/randomized_pairing_check.rs) and the BBS+ / accumulator verifiers.
pub fn try_multi_pairing_is_one(a: &[G1Affine], b: &[G2Affine]) -> Option<bool> {
    if a.len() != b.len() || a.is_empty() || !init() {
        return None;
    }
    let (g1, g2) = (pack_g1(a), pack_g2(b));
    let mut res = 0i32;
    let rc = unsafe { ffi::dg_multi_pairing_is_one(g1.as_ptr(), g2.as_ptr(), a.len(), &mut res) };
    (rc == ffi::DG_OK).then_some(res != 0)
}
