//! Dumps plonky2 0.2.0's own outputs for the inputs this repository's oracle is tested on.
//! Every quantity below is one the oracle (oracle/oracle.c) and the CUDA path claim to reproduce
//! bit for bit; tests/test_oracle_golden.py::test_plonky2_dump compares them when the dump exists.
//!
//! NOT compiled here (no Rust toolchain).  API names follow plonky2 0.2.0:
//!   PolynomialBatch::from_values / from_coeffs  (fri/oracle.rs)
//!   MerkleTree::new, .cap, .digests, .leaves     (hash/merkle_tree.rs)
//!   PoseidonHash::hash_no_pad / two_to_one       (hash/poseidon.rs, plonk/config.rs)
//!   fri_committed_trees' per-layer steps         (fri/prover.rs)
use plonky2::field::extension::quadratic::QuadraticExtension;
use plonky2::field::extension::FieldExtension;
use plonky2::field::goldilocks_field::GoldilocksField as F;
use plonky2::field::polynomial::{PolynomialCoeffs, PolynomialValues};
use plonky2::field::types::{Field, PrimeField64};
use plonky2::fri::oracle::PolynomialBatch;
use plonky2::hash::hash_types::HashOut;
use plonky2::hash::merkle_tree::MerkleTree;
use plonky2::hash::poseidon::PoseidonHash;
use plonky2::plonk::config::{GenericConfig, Hasher, PoseidonGoldilocksConfig};
use plonky2::plonk::plonk_common::reduce_with_powers;
use plonky2::util::timing::TimingTree;
use plonky2::util::{reverse_index_bits_in_place, transpose};
use sha2::{Digest, Sha256};

type C = PoseidonGoldilocksConfig;
const D: usize = 2;
type FE = QuadraticExtension<F>;
const P: u64 = 0xFFFF_FFFF_0000_0001;

/// SURVEY.md §8(d) workload = verifiable-fhe-paper_b200/plonky2_api.py::synthetic_columns:
/// column c, row i (1-based counter) = splitmix64(seed + c, i) reduced mod p.
fn synthetic_columns(ncols: usize, n: usize, seed: u64) -> Vec<Vec<F>> {
    (0..ncols)
        .map(|c| {
            (1..=n as u64)
                .map(|i| {
                    let mut z = (seed.wrapping_add(c as u64)).wrapping_add(i.wrapping_mul(0x9E3779B97F4A7C15));
                    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
                    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
                    z ^= z >> 31;
                    F::from_canonical_u64(if z >= P { z - P } else { z })
                })
                .collect()
        })
        .collect()
}

fn hex(h: &HashOut<F>) -> Vec<String> {
    h.elements.iter().map(|e| format!("{:016x}", e.to_canonical_u64())).collect()
}
fn sha_of<'a>(it: impl Iterator<Item = &'a F>) -> String {
    // sha256 over the little-endian canonical u64 bytes — numpy's arr.tobytes() of a uint64 array
    let mut s = Sha256::new();
    for e in it {
        s.update(e.to_canonical_u64().to_le_bytes());
    }
    format!("{:x}", s.finalize())
}

fn dump_commit(name: &str, log_n: usize, ncols: usize, rate_bits: usize, cap_height: usize,
               from_coeffs: bool, seed: u64) -> String {
    let cols = synthetic_columns(ncols, 1 << log_n, seed);
    let mut timing = TimingTree::default();
    let batch: PolynomialBatch<F, C, D> = if from_coeffs {
        PolynomialBatch::from_coeffs(cols.into_iter().map(PolynomialCoeffs::new).collect(), rate_bits,
                                     false, cap_height, &mut timing, None)
    } else {
        PolynomialBatch::from_values(cols.into_iter().map(PolynomialValues::new).collect(), rate_bits,
                                     false, cap_height, &mut timing, None)
    };
    let t = &batch.merkle_tree;
    let cap: Vec<Vec<String>> = t.cap.0.iter().map(hex).collect();
    format!(
        "{{\"name\":\"{name}\",\"log_n\":{log_n},\"ncols\":{ncols},\"rate_bits\":{rate_bits},\
         \"cap_height\":{cap_height},\"inputs_are_coeffs\":{from_coeffs},\"seed\":{seed},\
         \"cap\":{cap:?},\"sha256_coeffs\":\"{}\",\"sha256_leaves\":\"{}\",\"sha256_digests\":\"{}\",\
         \"lde_row_5\":{:?}}}",
        sha_of(batch.polynomials.iter().flat_map(|p| p.coeffs.iter())),
        sha_of(t.leaves.iter().flat_map(|l| l.iter())),
        sha_of(t.digests.iter().flat_map(|d| d.elements.iter())),
        batch.get_lde_values(5, 1).iter().map(|e| format!("{:016x}", e.to_canonical_u64())).collect::<Vec<_>>(),
    )
}

fn main() {
    let f = |v: u64| F::from_canonical_u64(v);
    let mut out = vec![];
    // sponge / compression anchors (SURVEY.md §8(c) model anchors)
    let h19 = PoseidonHash::hash_no_pad(&(1..10).map(f).collect::<Vec<_>>());
    let h07 = PoseidonHash::hash_no_pad(&(0..8).map(f).collect::<Vec<_>>());
    let t21 = PoseidonHash::two_to_one(HashOut { elements: [f(1), f(2), f(3), f(4)] },
                                       HashOut { elements: [f(5), f(6), f(7), f(8)] });
    out.push(format!("\"hash_no_pad_1_9\":{:?}", hex(&h19)));
    out.push(format!("\"hash_no_pad_0_7\":{:?}", hex(&h07)));
    out.push(format!("\"two_to_one_1234_5678\":{:?}", hex(&t21)));
    // hash_or_noop threshold: 4 elements are copied, 5 are hashed
    let noop4 = <PoseidonHash as Hasher<F>>::hash_or_noop(&(1..5).map(f).collect::<Vec<_>>());
    let hash5 = <PoseidonHash as Hasher<F>>::hash_or_noop(&(1..6).map(f).collect::<Vec<_>>());
    out.push(format!("\"hash_or_noop_4\":{:?},\"hash_or_noop_5\":{:?}", hex(&noop4), hex(&hash5)));
    // commits: the shapes of tests/golden/oracle_commits.json + BASELINE configs[1] itself
    let commits = vec![
        dump_commit("survey_like_8x9", 3, 9, 1, 1, false, 0x5EED0000),
        dump_commit("t_wires", 13, 135, 3, 4, false, 0x5EED0000),
        dump_commit("quotient_from_coeffs", 10, 16, 3, 4, true, 0x5EED0000),
        dump_commit("all_cap", 2, 7, 1, 3, false, 0x5EED0000),
        dump_commit("microbench_2^16x128", 16, 128, 3, 4, false, 0x5EED0000),
    ];
    out.push(format!("\"commits\":[{}]", commits.join(",")));
    // one FRI commit-phase layer as fri_committed_trees runs it (arity 16, cap height 4)
    let n = 1usize << 12;
    let planes = synthetic_columns(2, n, 0xF1F1); // real parts, imaginary parts
    let coeffs: Vec<FE> = (0..n).map(|i| FE::from_basefield_array([planes[0][i], planes[1][i]])).collect();
    let pc = PolynomialCoeffs::new(coeffs.clone());
    let mut values = pc.coset_fft(F::coset_shift().into()).values;
    reverse_index_bits_in_place(&mut values);
    let leaves: Vec<Vec<F>> = values.chunks(16).map(|c| c.iter().flat_map(|e| e.0.to_vec()).collect()).collect();
    let tree = MerkleTree::<F, <C as GenericConfig<D>>::Hasher>::new(leaves, 4);
    let beta = FE::from_basefield_array([f(0x1234_5678_9ABC_DEF0 % P), f(0x0FED_CBA9_8765_4321)]);
    let folded: Vec<FE> = coeffs.chunks_exact(16).map(|c| reduce_with_powers(c, beta)).collect();
    out.push(format!("\"fri_layer\":{{\"cap\":{:?},\"folded_0\":{:?}}}",
                     tree.cap.0.iter().map(hex).collect::<Vec<_>>(),
                     folded[0].0.iter().map(|e| format!("{:016x}", e.to_canonical_u64())).collect::<Vec<_>>()));
    let _ = transpose::<F>; // (transpose is exercised inside from_values)
    println!("{{{}}}", out.join(","));
}
