//! Runs the arkworks calls docknetwork/crypto makes on its hot path over the inputs in
//! ark_inputs.json and writes their outputs, ark-serialize encoded, to ark_vectors.json.
//!
//! Encodings: points = serialize_compressed (48 / 96 bytes, hex); scalars = 32-byte little-endian canonical integers;
//! Fp12 = serialize_compressed (576 bytes: c0.c0.c0, c0.c0.c1, ... each 48-byte little-endian canonical).
use ark_bls12_381::{Bls12_381, Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use ark_ec::{pairing::Pairing, scalar_mul::fixed_base::FixedBase, AffineRepr, CurveGroup, VariableBaseMSM};
use ark_ff::PrimeField;
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize};
use serde_json::{json, Value};
use std::{env, fs};

fn ser<T: CanonicalSerialize>(t: &T) -> String {
    let mut v = Vec::new();
    t.serialize_compressed(&mut v).unwrap();
    hex::encode(v)
}
fn strs(v: &Value) -> Vec<String> {
    v.as_array().unwrap().iter().map(|s| s.as_str().unwrap().to_string()).collect()
}
fn g1s(v: &Value) -> Vec<G1Affine> {
    strs(v).iter().map(|s| G1Affine::deserialize_compressed(&hex::decode(s).unwrap()[..]).unwrap()).collect()
}
fn g2s(v: &Value) -> Vec<G2Affine> {
    strs(v).iter().map(|s| G2Affine::deserialize_compressed(&hex::decode(s).unwrap()[..]).unwrap()).collect()
}
fn frs(v: &Value) -> Vec<Fr> {
    strs(v).iter().map(|s| Fr::from_le_bytes_mod_order(&hex::decode(s).unwrap())).collect()
}

fn main() {
    let args: Vec<String> = env::args().collect();
    let inp: Value = serde_json::from_str(&fs::read_to_string(&args[1]).unwrap()).unwrap();
    let mut out = serde_json::Map::new();

    // VariableBaseMSM::msm_bigint, G1 and G2 (legogroth16/src/prover.rs:286,344)
    let mut msms = Vec::new();
    for case in inp["msm_g1"].as_array().unwrap() {
        let bases = g1s(&case["bases"]);
        let big: Vec<_> = frs(&case["scalars"]).iter().map(|s| s.into_bigint()).collect();
        msms.push(json!({"name": case["name"], "result": ser(&G1Projective::msm_bigint(&bases, &big).into_affine())}));
    }
    out.insert("msm_g1".into(), Value::Array(msms));
    let mut msms = Vec::new();
    for case in inp["msm_g2"].as_array().unwrap() {
        let bases = g2s(&case["bases"]);
        let big: Vec<_> = frs(&case["scalars"]).iter().map(|s| s.into_bigint()).collect();
        msms.push(json!({"name": case["name"], "result": ser(&G2Projective::msm_bigint(&bases, &big).into_affine())}));
    }
    out.insert("msm_g2".into(), Value::Array(msms));

    // FixedBase::get_window_table + msm + normalize_batch, the body of utils::msm::WindowTable (utils/src/msm.rs:18-40)
    let mut fixed = Vec::new();
    for case in inp["fixed_base_g1"].as_array().unwrap() {
        let g = g1s(&case["point"])[0].into_group();
        let hint = case["hint"].as_u64().unwrap() as usize;
        let scalars = frs(&case["scalars"]);
        let scalar_size = Fr::MODULUS_BIT_SIZE as usize;
        let window = FixedBase::get_mul_window_size(hint);
        let table = FixedBase::get_window_table(scalar_size, window, g);
        let res = G1Projective::normalize_batch(&FixedBase::msm::<G1Projective>(scalar_size, window, &table, &scalars));
        let row1: Vec<String> = table[1].iter().take(8).map(|p| ser(p)).collect();
        fixed.push(json!({"name": case["name"], "window": window, "num_windows": table.len(),
                          "results": res.iter().map(|p| ser(p)).collect::<Vec<_>>(), "table_row1_first8": row1}));
    }
    out.insert("fixed_base_g1".into(), Value::Array(fixed));

    // AffineRepr::mul_bigint (vb_accumulator/src/witness.rs:190)
    let mut muls = Vec::new();
    for case in inp["mul_bigint_g1"].as_array().unwrap() {
        let pts = g1s(&case["points"]);
        let sc = frs(&case["scalars"]);
        let res: Vec<G1Projective> = pts.iter().zip(sc.iter()).map(|(p, s)| p.mul_bigint(s.into_bigint())).collect();
        muls.push(json!({"name": case["name"], "results": G1Projective::normalize_batch(&res).iter().map(|p| ser(p)).collect::<Vec<_>>()}));
    }
    out.insert("mul_bigint_g1".into(), Value::Array(muls));

    // Pairing::multi_miller_loop / final_exponentiation / multi_pairing (legogroth16/src/verifier.rs:69-80)
    let mut pairs = Vec::new();
    for case in inp["pairing"].as_array().unwrap() {
        let a = g1s(&case["g1"]);
        let b = g2s(&case["g2"]);
        let ml = Bls12_381::multi_miller_loop(a.clone(), b.clone());
        let miller = ser(&ml.0);
        let fe = Bls12_381::final_exponentiation(ml).unwrap();
        let mp = Bls12_381::multi_pairing(a, b);
        assert_eq!(fe, mp);
        pairs.push(json!({"name": case["name"], "miller_loop": miller, "final_exponentiation": ser(&fe.0)}));
    }
    out.insert("pairing".into(), Value::Array(pairs));

    out.insert("generator".into(), json!({"ark-ec": "0.4", "ark-bls12-381": "0.4", "inputs": args[1]}));
    fs::write(&args[2], serde_json::to_string_pretty(&Value::Object(out)).unwrap()).unwrap();
}
