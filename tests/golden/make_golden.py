#!/usr/bin/env python3
"""Generates the committed golden fixtures in tests/golden/.  Run in the build container
(it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Fixtures
  ntt_params.npz      the reference's own Goldilocks NTT golden vectors, parsed verbatim from
                      /root/reference/src/ntt/params_{8..2048}.rs (N, NINV, ROOTS, INVROOTS, TESTG,
                      TESTGHAT; exercised upstream by src/vtfhe/crypto/poly.rs:195-208 and
                      src/ntt/mod.rs:82-136).  They pin field mul/add/sub, the 2^k-th roots of
                      unity and a (negacyclic) transform.
  poseidon_kat.json   plonky2 0.2.0's Poseidon test vectors ([P2] hash/poseidon_goldilocks.rs
                      tests::test_vectors: all-zero, 0..11, all -1, one random-looking state).
                      plonky2's source is not on disk; the vectors are as published upstream and
                      each is reproduced by two independent implementations here.
  model_anchors.json  outputs of the independent big-integer model (oracle/model.py) on small
                      seeded inputs, incl. the anchors recorded in SURVEY.md §8(c).  They pin the
                      *reading* of plonky2's conventions (overwrite sponge, hash_or_noop, digests
                      layout, bit-reversed leaves, coset shift 7), not plonky2 itself.
  oracle_commits.json caps + sha256 of coeffs/leaves/digests for mid-size commits computed by the C
                      oracle (after it passed all of the above) — regression pins for sizes the
                      Python model is too slow for.
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = "/root/reference/src/ntt"


def parse_params():
    out = {}
    for n in (8, 16, 32, 64, 128, 256, 512, 1024, 2048):
        src = open(os.path.join(REF, "params_%d.rs" % n)).read()
        assert int(re.search(r"pub const N: usize = (\d+);", src).group(1)) == n
        out["NINV_%d" % n] = np.array([int(re.search(r"pub const NINV: u64 = (\d+);", src).group(1))],
                                      dtype=np.uint64)
        for name in ("ROOTS", "INVROOTS", "TESTG", "TESTGHAT"):
            m = re.search(r"pub const %s: \[u64; \d+\] = \[([^\]]*)\];" % name, src)
            arr = np.array([int(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()],
                           dtype=np.uint64)
            assert arr.size == n, (name, n, arr.size)
            out["%s_%d" % (name, n)] = arr
    np.savez_compressed(os.path.join(HERE, "ntt_params.npz"), **out)


def poseidon_kat():
    p = 2**64 - 2**32 + 1
    kat = [
        ([0] * 12,
         "3c18a9786cb0b359 c4055e3364a246c3 7953db0ab48808f4 c71603f33a1144ca d7709673896996dc "
         "46a84e87642f44ed d032648251ee0b3c 1c687363b207df62 df8565563e8045fe 40f5b37ff4254dae "
         "d070f637b431067c 1792b1c4342109d7"),
        (list(range(12)),
         "d64e1e3efc5b8e9e 53666633020aaa47 d40285597c6a8825 613a4f81e81231d2 414754bfebd051f0 "
         "cb1f8980294a023f 6eb2a9e4d54a9d0f 1902bc3af467e056 f045d5eafdc6021f e4150f77caaa3be5 "
         "c9bfd01d39b50cce 5c0a27fcb0e1459b"),
        ([p - 1] * 12,
         "be0085cfc57a8357 d95af71847d05c09 cf55a13d33c1c953 95803a74f4530e82 fcd99eb30a135df1 "
         "e095905e913a3029 de0392461b42919b 7d3260e24e81d031 10d3d0465d9deaa0 a87571083dfc2a47 "
         "e18263681e9958f8 e28e96f1ae5e60d3"),
        ([0x8ccbbbea4fe5d2b7, 0xc2af59ee9ec49970, 0x90f7e1a9e658446a, 0xdcc0630a3ab8b1b8,
          0x7ff8256bca20588c, 0x5d99a7ca0c44ecfb, 0x48452b17a70fbee3, 0xeb09d654690b6c88,
          0x4a55d3a39c676a88, 0xc0407a38d2285139, 0xa234bac9356386d1, 0xe1633f2bad98a52f],
         "a89280105650c4ec ab542d53860d12ed 5704148e9ccab94f d3a826d4b62da9f5 8a7a6ca87892574f "
         "c7017e1cad1a674e 1f06668922318e34 a3b203bc8102676f fcc781b0ce382bf2 934c69ff3ed14ba5 "
         "504688a5996e8f13 401f3f2ed524a2ba"),
    ]
    json.dump({"source": "[P2] plonky2 0.2.0 src/hash/poseidon_goldilocks.rs tests::test_vectors",
               "vectors": [{"input": ["%016x" % x for x in i], "output": o.split()} for i, o in kat],
               "round_constants_first12": "b585f766f2144405 7746a55f43921ad7 b2fb0d31cee799b4 "
               "0f6760a4803427d7 e10d666650f4e012 8cae14cb07d09bf1 d438539c95f63e9f ef781c7ce35b4c3d "
               "cdc4a239b0c44426 277fa208bf337bff e17653a29da578a1 c54302f225db2c76".split(),
               "round_constants_last4": "4543d9df5476d3cb f172d73e004fc90d dfd1c4febcc81238 "
               "bc8dfb627fe558fc".split()},
              open(os.path.join(HERE, "poseidon_kat.json"), "w"), indent=1)


def hx(v):
    return ["%016x" % int(x) for x in v]


def model_anchors():
    from oracle import model as M
    rng = np.random.default_rng(20240451)
    out = {"source": "oracle/model.py (independent big-integer model); SURVEY.md §8(c) anchors",
           "hash_no_pad_1_9": hx(M.hash_no_pad(list(range(1, 10)))),
           "hash_no_pad_0_7": hx(M.hash_no_pad(list(range(8)))),
           "two_to_one_1234_5678": hx(M.two_to_one([1, 2, 3, 4], [5, 6, 7, 8])),
           "commits": []}
    cases = [dict(name="survey_8x9", log_n=3, ncols=9, rate_bits=1, cap_height=1, coeffs=False,
                  salt=False, cols=[[8 * c + r + 1 for r in range(8)] for c in range(9)])]
    for (log_n, ncols, r, h, co, sa) in [(2, 3, 2, 0, False, False), (4, 5, 3, 4, True, False),
                                         (3, 20, 1, 4, False, True), (2, 4, 1, 3, False, False),
                                         (0, 7, 0, 0, False, False), (1, 2, 3, 2, False, True),
                                         (5, 16, 2, 3, False, False), (4, 135, 1, 2, False, False)]:
        cols = rng.integers(0, 2**64, size=(ncols, 1 << log_n), dtype=np.uint64).tolist()
        cases.append(dict(name="rand_%d_%d_%d_%d_%d_%d" % (log_n, ncols, r, h, co, sa), log_n=log_n,
                          ncols=ncols, rate_bits=r, cap_height=h, coeffs=co, salt=sa, cols=cols))
    for c in cases:
        m = (1 << c["log_n"]) << c["rate_bits"]
        salt = rng.integers(0, 2**64, size=(4, m), dtype=np.uint64).tolist() if c["salt"] else None
        res = M.commit(c["cols"], c["rate_bits"], c["cap_height"], c["coeffs"], salt)
        out["commits"].append(dict(
            name=c["name"], log_n=c["log_n"], ncols=c["ncols"], rate_bits=c["rate_bits"],
            cap_height=c["cap_height"], inputs_are_coeffs=c["coeffs"],
            cols=[hx(col) for col in c["cols"]], salt=[hx(s) for s in salt] if salt else None,
            coeffs=[hx(x) for x in res["coeffs"]], lde=[hx(x) for x in res["lde"]],
            leaves=[hx(x) for x in res["leaves"]], digests=[hx(x) for x in res["digests"]],
            cap=[hx(x) for x in res["cap"]]))
    # [P2] plonk/prover.rs wires_permutation_partial_products_and_zs on a fixed 5-wire x 8-row case
    # (max_degree 2 -> 3 chunks per row: columns Z, pp_0, pp_1)
    P = M.P
    beta, gamma = 0x123456789ABCDEF, 0xFEDCBA987654321
    wa = [[8 * j + i + 1 for i in range(8)] for j in range(5)]
    sa = [[3 * j + 5 * i + 2 for i in range(8)] for j in range(5)]
    ka = [pow(7, j, P) for j in range(5)]
    out["zs_partial_products"] = dict(beta=beta, gamma=gamma, max_degree=2,
                                      columns=[hx(c) for c in M.zs_partial_products(wa, sa, ka, 2, beta, gamma)])
    # [P2] fri/oracle.rs prove_openings up to final_poly on a fixed case: two FRI batches (5 and 2
    # polynomials of 8 coefficients), points and alpha in the quadratic extension
    fb = [[[((11 * b + 5 * j + 3) * (i + 1) ** 2 + 7 * i) % P for i in range(8)] for j in range(nb)]
          for b, nb in enumerate((5, 2))]
    fpts, falpha = [(0x1111111111111111, 0x2222222222222222), (3, 5)], (0x0123456789ABCDEF % P, 0x0FEDCBA987654321 % P)
    out["fri_final_poly"] = dict(batches=[[hx(f) for f in b] for b in fb], points=[hx(z) for z in fpts],
                                 alpha=hx(falpha), final_poly=[hx(c) for c in M.fri_final_poly(fb, fpts, falpha)])
    json.dump(out, open(os.path.join(HERE, "model_anchors.json"), "w"))


def oracle_commits():
    import vfhe_b200 as V
    from oracle import binding as B
    B.build()
    out = {"source": "oracle/liboracle.so on vfhe_b200.synthetic_columns(ncols, n, seed)", "cases": []}
    for (log_n, ncols, r, h, co, seed, canonical) in [
            (13, 135, 3, 4, False, 0x5EED0000, True), (13, 20, 3, 4, False, 0x5EED1000, True),
            (13, 16, 3, 4, True, 0x5EED2000, True), (10, 128, 3, 4, False, 0x5EED3000, False),
            (12, 9, 2, 0, False, 0x5EED4000, True), (9, 3, 1, 5, False, 0x5EED5000, False),
            (16, 8, 3, 4, False, 0x5EED6000, True)]:
        cols = V.synthetic_columns(ncols, 1 << log_n, seed, canonical)
        res = B.commit(cols, r, h, co)
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        out["cases"].append(dict(log_n=log_n, ncols=ncols, rate_bits=r, cap_height=h,
                                 inputs_are_coeffs=co, seed=seed, canonical=canonical,
                                 cap=[hx(x) for x in res["cap"]], sha256_coeffs=sha(res["coeffs"]),
                                 sha256_leaves=sha(res["leaves"]), sha256_digests=sha(res["digests"])))
    json.dump(out, open(os.path.join(HERE, "oracle_commits.json"), "w"), indent=1)


if __name__ == "__main__":
    parse_params()
    poseidon_kat()
    model_anchors()
    oracle_commits()
    print("golden fixtures written to", HERE)
