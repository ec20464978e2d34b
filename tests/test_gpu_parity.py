"""GPU suite: bit-exact parity of the CUDA path (through the C ABI) against the oracle, the
committed golden fixtures and — at BASELINE.json's full sizes — size-independent properties.
All integer work: the bar is exact equality everywhere."""
import hashlib
import random

import numpy as np
import pytest

from conftest import unhex

pytestmark = pytest.mark.gpu
P = 2**64 - 2**32 + 1
EDGE = [0, 1, 2, P - 1, P, P + 1, 2**64 - 1, 2**32 - 1, 2**32, 2**32 + 1, P - 2**32, 2**63]


def rand_u64(rng, shape, edge_frac=0.15):
    a = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    mask = rng.random(shape) < edge_frac
    a[mask] = rng.choice(np.array(EDGE, dtype=np.uint64), size=int(mask.sum()))
    return a


# ------------------------------------------------------------------------------ poseidon / hashes
def test_poseidon_known_answers(V, ctx, poseidon_kat):
    inp = unhex([v["input"] for v in poseidon_kat["vectors"]])
    want = unhex([v["output"] for v in poseidon_kat["vectors"]])
    assert np.array_equal(V.poseidon(inp, ctx), want)


def test_poseidon_random_and_noncanonical(V, ctx, oracle):
    rng = np.random.default_rng(1)
    states = rand_u64(rng, (4096, 12), 0.3)
    got = V.poseidon(states, ctx)
    for k in range(0, 4096, 7):
        assert np.array_equal(got[k], oracle.poseidon(states[k])), k
    assert (got < np.uint64(P)).all()


@pytest.mark.parametrize("width", [0, 1, 3, 4, 5, 7, 8, 9, 15, 16, 17, 20, 128, 135, 139])
def test_hash_or_noop_widths(V, ctx, oracle, width):
    rng = np.random.default_rng(width)
    rows = rand_u64(rng, (257, width), 0.2)
    got = V.hash_or_noop(rows, ctx)
    for k in range(0, 257, 16):
        assert np.array_equal(got[k], oracle.hash_or_noop(rows[k])), (width, k)


def test_two_to_one(V, ctx, oracle, model_anchors):
    assert ["%016x" % int(x) for x in V.two_to_one([1, 2, 3, 4], [5, 6, 7, 8], ctx)[0]] == \
        model_anchors["two_to_one_1234_5678"]
    rng = np.random.default_rng(5)
    l, r = rand_u64(rng, (300, 4)), rand_u64(rng, (300, 4))
    got = V.two_to_one(l, r, ctx)
    for k in range(0, 300, 11):
        assert np.array_equal(got[k], oracle.two_to_one(l[k], r[k]))


# ------------------------------------------------------------------------------ transforms
@pytest.mark.parametrize("n", [8, 16, 32, 64, 128, 256, 512, 1024, 2048])
def test_reference_ntt_vectors(V, ctx, ntt_params, oracle, n):
    """The reference's own golden vectors (src/ntt/params_N.rs) through the CUDA transforms."""
    lg = n.bit_length() - 1
    w = oracle.primitive_root_of_unity(lg + 1)
    rev = lambda i: int(format(i, "0%db" % lg)[::-1], 2)
    g, ghat = ntt_params["TESTG_%d" % n], ntt_params["TESTGHAT_%d" % n]
    ev = V.coset_fft(g, w, ctx)
    assert [int(ev[rev(k)]) for k in range(n)] == [int(x) for x in ghat]
    full = V.fft(np.concatenate([g, np.zeros(n, np.uint64)]), ctx)
    assert [int(full[2 * rev(k) + 1]) for k in range(n)] == [int(x) for x in ghat]
    assert np.array_equal(V.ifft(full, ctx)[:n], g)


@pytest.mark.parametrize("log_n", list(range(0, 21)))
def test_fft_ifft_coset_against_oracle(V, ctx, oracle, log_n):
    rng = np.random.default_rng(100 + log_n)
    v = rand_u64(rng, (1 << log_n,))
    assert np.array_equal(V.fft(v, ctx), oracle.fft(v))
    assert np.array_equal(V.ifft(v, ctx), oracle.ifft(v))
    shift = int(rng.integers(1, 2**64, dtype=np.uint64))
    assert np.array_equal(V.coset_fft(v, shift, ctx), oracle.coset_fft(v, shift))
    assert np.array_equal(V.ifft(V.fft(v, ctx), ctx), v % np.uint64(P))


@pytest.mark.parametrize("log_n,ncols,rate_bits,coeffs", [(0, 3, 3, False), (3, 5, 0, False),
                                                          (5, 9, 3, True), (9, 4, 2, False),
                                                          (12, 3, 3, False)])
def test_lde_values_natural_order(V, ctx, oracle, log_n, ncols, rate_bits, coeffs):
    rng = np.random.default_rng(7 * log_n + ncols)
    cols = rand_u64(rng, (ncols, 1 << log_n))
    co, lde = V.lde_values(cols, rate_bits, coeffs, ctx)
    ref = oracle.commit(cols, rate_bits, 0, coeffs, want_lde=True)
    assert np.array_equal(lde, ref["lde"])
    if not coeffs:
        assert np.array_equal(co, ref["coeffs"])


# ------------------------------------------------------------------------------ Merkle tree
@pytest.mark.parametrize("log_leaves,width,cap_height", [
    (0, 5, 0), (1, 5, 0), (1, 5, 1), (3, 4, 0), (3, 3, 3), (4, 9, 2), (6, 135, 4), (6, 20, 6),
    (10, 16, 4), (10, 128, 0), (12, 8, 5), (13, 33, 4)])
def test_merkle_new_against_oracle(V, ctx, oracle, log_leaves, width, cap_height):
    rng = np.random.default_rng(log_leaves * 100 + width)
    leaves = rand_u64(rng, (1 << log_leaves, width))
    tree = V.MerkleTree.new(leaves, cap_height, ctx)
    digests, cap = oracle.merkle_new(leaves, cap_height)
    assert np.array_equal(tree.cap, cap)
    assert np.array_equal(tree.digests, digests)
    for i in {0, (1 << log_leaves) - 1, (1 << log_leaves) // 3}:
        proof = tree.prove(i)
        assert oracle.merkle_verify(leaves[i], i, proof.siblings, cap)
        assert V.verify_merkle_proof_to_cap(leaves[i], i, tree.cap, proof, ctx)


def test_merkle_new_rejects_bad_arguments(V, ctx):
    with pytest.raises(ValueError):
        V.MerkleTree.new(np.zeros((6, 3), np.uint64), 1, ctx)
    with pytest.raises(ValueError):
        V.MerkleTree.new(np.zeros((8, 3), np.uint64), 4, ctx)
    # and at the C ABI itself
    a = np.zeros((8, 3), np.uint64)
    cap = np.zeros((16, 4), np.uint64)
    rc = ctx.lib.vpbs_merkle_new(ctx.handle, a.ctypes.data_as(V._lib.u64p), 8, 3, 4, None,
                                 cap.ctypes.data_as(V._lib.u64p))
    assert rc == V._lib.VPBS_ERR_ARG and b"cap_height" in ctx.lib.vpbs_last_error(ctx.handle)
    rc = ctx.lib.vpbs_merkle_new(ctx.handle, a.ctypes.data_as(V._lib.u64p), 6, 3, 0, None,
                                 cap.ctypes.data_as(V._lib.u64p))
    assert rc == V._lib.VPBS_ERR_ARG


# ------------------------------------------------------------------------------ PolynomialBatch
def check_batch(V, oracle, batch, cols, rate_bits, cap_height, coeffs, salt=None):
    ref = oracle.commit(cols, rate_bits, cap_height, coeffs, salt)
    assert np.array_equal(batch.merkle_tree.cap, ref["cap"])
    assert np.array_equal(batch.merkle_tree.digests, ref["digests"])
    assert np.array_equal(batch.merkle_tree.leaves, ref["leaves"])
    assert np.array_equal(batch.polynomials, ref["coeffs"])
    return ref


def test_model_anchor_commits(V, ctx, model_anchors):
    """Golden fixtures from the independent model (incl. SURVEY.md §8(c) anchors)."""
    for c in model_anchors["commits"]:
        cols = unhex(c["cols"])
        salt = unhex(c["salt"]) if c["salt"] else None
        f = V.PolynomialBatch.from_coeffs if c["inputs_are_coeffs"] else V.PolynomialBatch.from_values
        b = f(cols, c["rate_bits"], salt is not None, c["cap_height"], ctx=ctx, salt=salt)
        assert np.array_equal(b.polynomials, unhex(c["coeffs"])), c["name"]
        assert np.array_equal(b.merkle_tree.leaves, unhex(c["leaves"])), c["name"]
        want_d = unhex(c["digests"]) if c["digests"] else np.empty((0, 4), np.uint64)
        assert np.array_equal(b.merkle_tree.digests, want_d), c["name"]
        assert np.array_equal(b.merkle_tree.cap, unhex(c["cap"])), c["name"]
        # get_lde_values(i) = natural-order LDE row i without the salt
        lde = unhex(c["lde"])
        for i in {0, min(1, lde.shape[1] - 1), lde.shape[1] - 1}:
            assert np.array_equal(b.get_lde_values(i), lde[:, i]), c["name"]


SHAPES = [(lg, C, r, h) for lg in (0, 1, 2, 3, 5, 8, 9, 11) for C in (1, 3, 4, 5, 8, 9, 16, 20)
          for (r, h) in ((0, 0), (1, 1), (3, 4), (2, 0))]


@pytest.mark.parametrize("log_n,ncols,rate_bits,cap_height",
                         [s for i, s in enumerate(SHAPES) if i % 5 == 0 and s[3] <= s[0] + s[2]])
def test_commit_shapes_from_values(V, ctx, oracle, log_n, ncols, rate_bits, cap_height):
    rng = np.random.default_rng(hash((log_n, ncols, rate_bits, cap_height)) % 2**32)
    cols = rand_u64(rng, (ncols, 1 << log_n))
    b = V.PolynomialBatch.from_values(cols, rate_bits, False, cap_height, ctx=ctx)
    check_batch(V, oracle, b, cols, rate_bits, cap_height, False)


@pytest.mark.parametrize("log_n,ncols,rate_bits,cap_height", [
    (4, 135, 3, 4), (7, 128, 3, 7), (10, 135, 3, 4), (12, 20, 3, 4), (13, 16, 3, 4), (6, 7, 3, 9),
    (14, 2, 1, 15), (16, 3, 3, 4), (17, 2, 2, 0)])
def test_commit_more_shapes(V, ctx, oracle, log_n, ncols, rate_bits, cap_height):
    rng = np.random.default_rng(log_n * 1000 + ncols)
    cols = rand_u64(rng, (ncols, 1 << log_n))
    for coeffs in (False, True):
        f = V.PolynomialBatch.from_coeffs if coeffs else V.PolynomialBatch.from_values
        b = f(cols, rate_bits, False, cap_height, ctx=ctx)
        check_batch(V, oracle, b, cols, rate_bits, cap_height, coeffs)


def test_commit_with_blinding_salt(V, ctx, oracle):
    rng = np.random.default_rng(99)
    cols = rand_u64(rng, (20, 256))
    salt = rand_u64(rng, (4, 256 << 3))
    b = V.PolynomialBatch.from_values(cols, 3, True, 4, ctx=ctx, salt=salt)
    ref = check_batch(V, oracle, b, cols, 3, 4, False, salt)
    assert b.merkle_tree.leaves.shape[1] == 24
    assert np.array_equal(b.get_lde_values(5), ref["leaves"][V.reverse_bits(5, 11)][:20])
    # a fresh random salt is drawn on the host when none is given
    b2 = V.PolynomialBatch.from_values(cols, 3, True, 4, ctx=ctx, rng=np.random.default_rng(1))
    assert not np.array_equal(b2.merkle_tree.cap, b.merkle_tree.cap)
    assert np.array_equal(b2.merkle_tree.leaves[:, :20], b.merkle_tree.leaves[:, :20])


def test_adversarial_columns(V, ctx, oracle):
    """All-zero, all-(p-1), non-canonical and 2^64-1 columns."""
    n = 1 << 10
    rng = np.random.default_rng(4)
    cols = np.stack([np.zeros(n, np.uint64), np.full(n, P - 1, np.uint64),
                     np.full(n, 2**64 - 1, np.uint64), np.full(n, P, np.uint64),
                     rng.integers(P, 2**64, size=n, dtype=np.uint64),
                     np.arange(n, dtype=np.uint64)])
    for coeffs in (False, True):
        f = V.PolynomialBatch.from_coeffs if coeffs else V.PolynomialBatch.from_values
        b = f(cols, 3, False, 4, ctx=ctx)
        check_batch(V, oracle, b, cols, 3, 4, coeffs)
    assert (b.merkle_tree.leaves < np.uint64(P)).all()


def test_commit_rejects_bad_arguments(V, ctx):
    with pytest.raises(ValueError):
        V.PolynomialBatch.from_values(np.zeros((2, 8), np.uint64), 1, False, 5, ctx=ctx)
    cols = np.zeros((1, 8), np.uint64)
    colp = (V._lib.u64p * 1)(cols[0].ctypes.data_as(V._lib.u64p))
    cap = np.zeros((64, 4), np.uint64)
    rc = ctx.lib.vpbs_commit(ctx.handle, colp, 1, 3, 1, 5, 0, None, None, None, None,
                             cap.ctypes.data_as(V._lib.u64p), None)
    assert rc == V._lib.VPBS_ERR_ARG
    rc = ctx.lib.vpbs_commit(ctx.handle, colp, 0, 3, 1, 1, 0, None, None, None, None,
                             cap.ctypes.data_as(V._lib.u64p), None)
    assert rc == V._lib.VPBS_ERR_ARG


def test_oracle_regression_pins_on_gpu(V, ctx, oracle_commits):
    """Committed fixtures (caps + sha256 of every output) incl. the N=8 step shapes and 2^16 rows."""
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for c in oracle_commits["cases"]:
        cols = V.synthetic_columns(c["ncols"], 1 << c["log_n"], c["seed"], c["canonical"])
        f = V.PolynomialBatch.from_coeffs if c["inputs_are_coeffs"] else V.PolynomialBatch.from_values
        b = f(cols, c["rate_bits"], False, c["cap_height"], ctx=ctx)
        assert ["%016x" % int(x) for x in b.merkle_tree.cap.reshape(-1)] == sum(c["cap"], [])
        assert sha(b.polynomials) == c["sha256_coeffs"]
        assert sha(b.merkle_tree.leaves) == c["sha256_leaves"]
        assert sha(b.merkle_tree.digests) == c["sha256_digests"]


# ------------------------------------------------------------------------------ full-size configs
def test_microbench_config_full_size(V, ctx, oracle):
    """BASELINE.json configs[1]: 2^16 rows x 128 columns, rate_bits 3, cap_height 4 — compared
    bit-for-bit with the (multi-threaded) oracle, plus size-independent properties."""
    cols = V.synthetic_columns(128, 1 << 16)
    b = V.PolynomialBatch.from_values(cols, 3, False, 4, ctx=ctx)
    ref = oracle.commit(cols, 3, 4, False)
    assert np.array_equal(b.merkle_tree.cap, ref["cap"])
    assert np.array_equal(b.polynomials, ref["coeffs"])
    assert np.array_equal(b.merkle_tree.digests, ref["digests"])
    assert np.array_equal(b.merkle_tree.leaves, ref["leaves"])
    # properties: round trip, linearity of the LDE, Merkle openings verify
    assert np.array_equal(V.fft(b.polynomials[5], ctx), cols[5])
    rnd = random.Random(3)
    for _ in range(4):
        i = rnd.randrange(1 << 19)
        assert oracle.merkle_verify(b.merkle_tree.get(i), i, b.merkle_tree.prove(i).siblings,
                                    b.merkle_tree.cap)
    # leaf k = natural LDE row bitrev(k): evaluate column 3 directly at 7 * w^bitrev(k)
    w = oracle.primitive_root_of_unity(19)
    for k in (0, 1, 12345, (1 << 19) - 1):
        x = oracle.gl_mul(7, oracle.gl_pow(w, V.reverse_bits(k, 19)))
        acc = 0
        for c in b.polynomials[3][::-1]:
            acc = (acc * x + int(c)) % P
        assert int(b.merkle_tree.leaves[k, 3]) == acc


def test_microbench_config_full_size_adversarial(V, ctx, oracle):
    """the configs[1] shape with raw (non-canonical) u64 inputs and constant columns —
    all zeros, all p - 1, all 2^64 - 1, all p — through the eager and the resident commit."""
    rng = np.random.default_rng(2024)
    cols = rng.integers(0, 2**64, size=(128, 1 << 16), dtype=np.uint64)
    cols[0] = 0
    cols[1] = P - 1
    cols[2] = 2**64 - 1
    cols[3] = P
    cols[127, ::2] = P - 1
    ref = oracle.commit(cols, 3, 4, False)
    b = V.PolynomialBatch.from_values(cols, 3, False, 4, ctx=ctx)
    assert np.array_equal(b.merkle_tree.cap, ref["cap"])
    assert np.array_equal(b.polynomials, ref["coeffs"])
    assert np.array_equal(b.merkle_tree.digests, ref["digests"])
    assert np.array_equal(b.merkle_tree.leaves, ref["leaves"])
    assert (b.merkle_tree.leaves < np.uint64(P)).all() and (b.polynomials < np.uint64(P)).all()
    assert not b.polynomials[0].any() and not b.polynomials[3].any()          # 0 and p are the zero polynomial
    assert int(b.polynomials[1][0]) == P - 1 and not b.polynomials[1][1:].any()  # constants have degree 0
    rb = V.commit_resident(cols, 3, False, 4, ctx=ctx)
    assert np.array_equal(rb.merkle_tree.cap, ref["cap"])
    rb.close()


def test_step_shapes_full_size(V, ctx, oracle):
    """BASELINE.json configs[2] stand-in: the three commits of one N=1024 IVC step, and build()'s
    one-off constants/sigmas commit (~85 columns, ivc_based_vpbs.rs:275)."""
    for (ncols, coeffs, seed) in ((135, False, 0x5EED0000), (20, False, 0x5EED1000), (16, True, 0x5EED2000),
                                  (85, False, 0x5EED3000)):
        cols = V.synthetic_columns(ncols, 1 << 16, seed)
        f = V.PolynomialBatch.from_coeffs if coeffs else V.PolynomialBatch.from_values
        b = f(cols, 3, False, 4, ctx=ctx)
        ref = oracle.commit(cols, 3, 4, coeffs)
        assert np.array_equal(b.merkle_tree.cap, ref["cap"])
        assert np.array_equal(b.merkle_tree.digests, ref["digests"])
        assert np.array_equal(b.merkle_tree.leaves, ref["leaves"])


def test_linearity_of_lde(V, ctx):
    """LDE(a + b) = LDE(a) + LDE(b) on full-size columns (size-independent property)."""
    rng = np.random.default_rng(8)
    a = rng.integers(0, P, size=(2, 1 << 16), dtype=np.uint64)
    s = ((a[0].astype(object) + a[1].astype(object)) % P).astype(np.uint64).reshape(1, -1)
    la = V.PolynomialBatch.from_values(a, 3, False, 4, ctx=ctx).merkle_tree.leaves
    ls = V.PolynomialBatch.from_values(s, 3, False, 4, ctx=ctx).merkle_tree.leaves
    want = (la[:, 0].astype(object) + la[:, 1].astype(object)) % P
    assert np.array_equal(ls[:, 0].astype(object), want)


def test_cpp_host_mirror(V, ctx, oracle, tmp_path):
    """include/vpbs_commit.hpp (C++ mirror of plonky2's API over the C ABI) vs the oracle."""
    import os
    import subprocess
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    exe = str(tmp_path / "test_api")
    libdir = os.path.join(root, "verifiable-fhe-paper_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", os.path.join(root, "tests", "cpp", "test_api.cpp"),
                    "-o", exe, "-L" + libdir, "-lvpbs_commit", "-L" + os.path.join(root, "oracle"),
                    "-loracle", "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(root, "oracle")],
                   check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "cpp host mirror ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("log_n,ncols,rate_bits,cap_height,world", [
    (10, 20, 3, 4, 8), (10, 135, 3, 4, 2), (12, 16, 3, 3, 4), (8, 5, 2, 6, 4), (13, 9, 1, 1, 2)])
def test_row_range_shards_equal_full_commit(V, ctx, oracle, log_n, ncols, rate_bits, cap_height, world):
    """vpbs_commit_shard_dev: every shard (run here one after the other on one GPU) reproduces its
    rows, digests and subtree roots of the full commit bit-for-bit (multi-GPU parity rule)."""
    import torch
    cols = V.synthetic_columns(ncols, 1 << log_n, seed=31337)
    ref = oracle.commit(cols, rate_bits, cap_height)
    d_cols = torch.from_numpy(cols.view(np.int64)).cuda()
    torch.cuda.synchronize()
    caps = []
    for rank in range(world):
        plan, leaves, digests, cap, coeffs, _ = _one_shard(V, ctx, d_cols, ncols, log_n, rate_bits,
                                                            cap_height, rank, world)
        torch.cuda.synchronize()
        assert np.array_equal(leaves.cpu().numpy().view(np.uint64),
                              ref["leaves"][plan.first_leaf: plan.first_leaf + plan.nleaves])
        assert np.array_equal(digests.cpu().numpy().view(np.uint64),
                              ref["digests"][plan.digest_offset: plan.digest_offset + plan.ndigests])
        assert np.array_equal(coeffs.cpu().numpy().view(np.uint64), ref["coeffs"])
        caps.append(cap.cpu().numpy().view(np.uint64))
    assert np.array_equal(np.concatenate(caps), ref["cap"])


def _one_shard(V, ctx, d_cols, ncols, log_n, rate_bits, cap_height, rank, world):
    import torch
    plan = V.shard_plan(log_n, rate_bits, cap_height, rank, world)
    n = 1 << log_n
    dev = d_cols.device
    coeffs = torch.empty((ncols, n), dtype=torch.int64, device=dev)
    leaves = torch.empty((plan.nleaves, ncols), dtype=torch.int64, device=dev)
    digests = torch.empty((max(plan.ndigests, 1), 4), dtype=torch.int64, device=dev)
    roots = torch.empty((plan.ncap, 4), dtype=torch.int64, device=dev)
    V.commit_shard_device(ctx, d_cols.data_ptr(), ncols, log_n, rate_bits, cap_height, False,
                          plan.first_leaf, plan.nleaves, coeffs.data_ptr(), leaves.data_ptr(),
                          digests.data_ptr() if plan.ndigests else 0, roots.data_ptr())
    return plan, leaves, digests[:plan.ndigests], roots, coeffs, None


def test_shard_rejects_bad_ranges(V, ctx):
    import torch
    cols = torch.zeros((2, 16), dtype=torch.int64, device="cuda")
    out = torch.zeros((1024,), dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):   # not a whole LDE block
        V.commit_shard_device(ctx, cols.data_ptr(), 2, 4, 3, 2, False, 8, 8, 0, out.data_ptr(),
                              out.data_ptr(), out.data_ptr())
    with pytest.raises(ValueError):   # smaller than a cap subtree
        V.commit_shard_device(ctx, cols.data_ptr(), 2, 4, 3, 0, False, 0, 16, 0, out.data_ptr(),
                              out.data_ptr(), out.data_ptr())


@pytest.mark.parametrize("log_n,ncols,rate_bits,cap_height,coeffs,salted,nctx", [
    (10, 9, 3, 4, False, False, 2), (10, 9, 3, 4, False, False, 8), (12, 70, 3, 3, False, False, 4),
    (12, 64, 2, 2, True, True, 4), (6, 5, 1, 6, False, False, 2), (13, 135, 3, 4, False, False, 8),
    (4, 3, 3, 7, True, False, 1)])
def test_commit_multi_matches_single_commit(V, oracle, log_n, ncols, rate_bits, cap_height, coeffs,
                                            salted, nctx):
    """vpbs_commit_multi: one commit spread over nctx contexts (row ranges; here the contexts may
    share a device — the data path is the same as on distinct GPUs) equals the oracle bit for bit:
    coefficients, leaves, digests in plonky2 layout, cap."""
    import torch
    ndev = torch.cuda.device_count()
    ctxs = [V.Context(g % ndev) for g in range(nctx)]
    rng = np.random.default_rng(1000 * log_n + ncols + nctx)
    cols = rand_u64(rng, (ncols, 1 << log_n), 0.05)
    m = (1 << log_n) << rate_bits
    salt = rand_u64(rng, (4, m)) if salted else None
    f = V.PolynomialBatch.from_coeffs if coeffs else V.PolynomialBatch.from_values
    b = f(cols, rate_bits, salted, cap_height, ctxs=ctxs, salt=salt)
    ref = oracle.commit(cols, rate_bits, cap_height, coeffs, salt)
    assert np.array_equal(b.merkle_tree.cap, ref["cap"])
    assert np.array_equal(b.merkle_tree.digests, ref["digests"])
    assert np.array_equal(b.merkle_tree.leaves, ref["leaves"])
    assert np.array_equal(b.polynomials, ref["coeffs"])
    assert b.stats["kernel_launches"] > 0
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("log_n,ncols", [(12, 70), (8, 5)])
def test_commit_with_scattered_host_columns(V, ctx, oracle, log_n, ncols):
    """The C ABI takes one pointer per column (a Vec<PolynomialValues<F>> is ncols separate
    allocations).  Adjacent columns are merged into one transfer, so exercise the other cases:
    separately allocated columns, columns in reverse address order, runs of adjacent columns with
    gaps, and NULL entries in coeffs_out (columns the caller does not want back)."""
    import ctypes
    rng = np.random.default_rng(log_n + ncols)
    n = 1 << log_n
    r, h = 2, 3
    m = n << r
    backing = rand_u64(rng, (2 * ncols + 8, n))          # columns live at odd rows, reversed order
    rows = [2 * (ncols - 1 - c) + 1 for c in range(ncols)]
    rows[:4] = [2 * ncols + 1, 2 * ncols + 2, 2 * ncols + 3, 2 * ncols + 5]  # a run of 3, a gap, 1
    cols = np.stack([backing[i] for i in rows])
    u64p = V._lib.u64p
    colp = (u64p * ncols)(*[backing[i].ctypes.data_as(u64p) for i in rows])
    out_back = np.zeros((ncols + 2, n), np.uint64)
    want = [c for c in range(ncols) if c % 3 != 1]
    cop = (u64p * ncols)(*[out_back[c + (c > 2)].ctypes.data_as(u64p) if c in want else None
                           for c in range(ncols)])
    leaves = np.empty((m, ncols), np.uint64)
    digests = np.empty((2 * (m - (1 << h)), 4), np.uint64)
    cap = np.empty((1 << h, 4), np.uint64)
    ctx.check(ctx.lib.vpbs_commit(ctx.handle, colp, ncols, log_n, r, h, 0, None, cop,
                                  leaves.ctypes.data_as(u64p), digests.ctypes.data_as(u64p),
                                  cap.ctypes.data_as(u64p), None))
    ref = oracle.commit(cols, r, h, False)
    assert np.array_equal(cap, ref["cap"])
    assert np.array_equal(digests, ref["digests"])
    assert np.array_equal(leaves, ref["leaves"])
    for c in range(ncols):
        got = out_back[c + (c > 2)]
        if c in want:
            assert np.array_equal(got, ref["coeffs"][c]), c
        else:
            assert not got.any(), c   # skipped outputs stay untouched
    assert not out_back[3].any()      # the gap row between outputs 2 and 3


def test_commit_multi_rejects_bad_context_lists(V, ctx):
    cols = V.synthetic_columns(4, 1 << 6)
    a, b, c = V.Context(0), V.Context(0), V.Context(0)
    with pytest.raises(ValueError):   # not a power of two
        V.PolynomialBatch.from_values(cols, 3, False, 4, ctxs=[a, b, c])
    with pytest.raises(ValueError):   # more GPUs than LDE blocks
        V.PolynomialBatch.from_values(cols, 0, False, 4, ctxs=[a, b])
    with pytest.raises(ValueError):   # more GPUs than cap subtrees
        V.PolynomialBatch.from_values(cols, 3, False, 0, ctxs=[a, b])
    with pytest.raises(ValueError):   # the same context twice
        V.PolynomialBatch.from_values(cols, 3, False, 4, ctxs=[a, a])
    with pytest.raises(ValueError):
        V.PolynomialBatch.from_values(cols, 3, False, 4, ctxs=[])
    for x in (a, b, c):
        x.close()


def _nccl_worker(rank, world, port, q):
    import os
    import sys
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import vfhe_b200 as V
    from oracle import binding as B
    c = V.Context(rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    c.set_stream(stream.cuda_stream)
    log_n, ncols, r, h = 12, 20, 3, 4
    cols = V.synthetic_columns(ncols, 1 << log_n, seed=5)
    d_cols = torch.from_numpy(cols.view(np.int64)).cuda()
    plan, leaves, digests, cap, coeffs, _ = V.commit_sharded(c, d_cols, ncols, log_n, r, h, False, rank, world)
    torch.cuda.synchronize()
    ref = B.commit(cols, r, h)
    ok = (np.array_equal(cap.cpu().numpy().view(np.uint64), ref["cap"])
          and np.array_equal(leaves.cpu().numpy().view(np.uint64),
                             ref["leaves"][plan.first_leaf: plan.first_leaf + plan.nleaves])
          and np.array_equal(digests.cpu().numpy().view(np.uint64),
                             ref["digests"][plan.digest_offset: plan.digest_offset + plan.ndigests]))
    res = [None] * world
    dist.all_gather_object(res, bool(ok))
    if rank == 0:
        q.put(all(res))
    dist.destroy_process_group()


def test_sharded_commit_over_nccl(V):
    """One commit split over all visible GPUs (needs >= 2): only the subtree roots cross NVLink."""
    import socket
    import torch
    import torch.multiprocessing as mp
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 1 << (min(world, 8).bit_length() - 1)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _nccl_proof_worker(rank, world, port, q):
    import os
    import sys
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import vfhe_b200 as V
    from oracle import binding as B
    c = V.Context(rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    c.set_stream(stream.cuda_stream)
    c.set_shard(rank, world)
    sp = V.ShardedProof(rank, world, torch.device("cuda", rank))
    log_n, r, h = 12, 3, 4
    m = (1 << log_n) << r
    ok = True
    for ncols, coeffs in [(21, False), (16, True), (3, False)]:     # 21 and 3: not multiples of the world size
        cols = V.synthetic_columns(ncols, 1 << log_n, seed=50 + ncols)
        handle, cap = sp.commit_from_host(c, cols, r, h, coeffs)
        ref = B.commit(cols, r, h, coeffs)
        ok = ok and np.array_equal(cap, ref["cap"])
        qidx = np.random.default_rng(3).integers(0, m, size=28, dtype=np.uint64)
        qidx[:world] = np.arange(world, dtype=np.uint64) * np.uint64(m // world)   # every rank serves some
        own = sp.owned(qidx, m)
        rows = np.zeros((28, ncols), np.uint64)
        sibs = np.zeros((28, log_n + r - h, 4), np.uint64)
        rb = V.ResidentPolynomialBatch(c, handle, cap, ncols, log_n, r, False, None)
        rows[own] = rb.merkle_tree.get_many(qidx[own])
        sibs[own] = np.stack([p.siblings for p in rb.merkle_tree.prove_many(qidx[own])])
        sp.collect([rows, sibs])
        for k, i in enumerate(qidx):
            ok = ok and np.array_equal(rows[k], ref["leaves"][int(i)])
            ok = ok and B.merkle_verify(rows[k], int(i), sibs[k], ref["cap"])
        rb.close()
    res = [None] * world
    dist.all_gather_object(res, bool(ok))
    if rank == 0:
        q.put(all(res))
    dist.destroy_process_group()


def test_sharded_proof_over_nccl(V):
    """One proof's resident batches sharded over all visible GPUs (needs >= 2): host columns cross
    PCIe once and travel on by NVLink all-gather, every rank commits its row range, caps are completed
    and query openings collected over NCCL; everything equals the oracle's unsharded commit."""
    import socket
    import torch
    import torch.multiprocessing as mp
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 1 << (min(world, 8).bit_length() - 1)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_nccl_proof_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_randomized_shapes_against_oracle(V, ctx, oracle):
    """SURVEY.md §4 test plan (iii): random (log_n, cols, rate_bits, cap_height, from_values/coeffs)
    with seeded inputs incl. non-canonical words, sizes bounded so the oracle stays fast."""
    rnd = random.Random(20240451)
    rng = np.random.default_rng(20240451)
    done = 0
    while done < 48:
        log_n = rnd.randint(0, 17)
        ncols = rnd.choice([1, 3, 4, 5, 8, 9, 16, 20, 33, 128, 135])
        r = rnd.randint(0, 3)
        if (ncols << (log_n + r)) > (1 << 23):
            continue
        h = rnd.randint(0, log_n + r)
        coeffs = rnd.random() < 0.35
        cols = rand_u64(rng, (ncols, 1 << log_n), 0.05)
        f = V.PolynomialBatch.from_coeffs if coeffs else V.PolynomialBatch.from_values
        b = f(cols, r, False, h, ctx=ctx)
        ref = oracle.commit(cols, r, h, coeffs)
        tag = (log_n, ncols, r, h, coeffs)
        assert np.array_equal(b.merkle_tree.cap, ref["cap"]), tag
        assert np.array_equal(b.merkle_tree.digests, ref["digests"]), tag
        assert np.array_equal(b.merkle_tree.leaves, ref["leaves"]), tag
        assert np.array_equal(b.polynomials, ref["coeffs"]), tag
        done += 1


@pytest.mark.parametrize("log_n,ncols,rate_bits,cap_height,coeffs,salted", [
    (10, 135, 3, 4, False, False), (12, 16, 3, 4, True, False), (6, 20, 3, 9, False, False),
    (8, 9, 2, 0, False, True), (0, 5, 3, 1, False, False),
    # wide enough for the column-chunked upload (ragged last chunk; coefficients + salt)
    (12, 70, 2, 3, False, False), (12, 64, 1, 2, True, True)])
def test_resident_batch_lazy_openings(V, ctx, oracle, log_n, ncols, rate_bits, cap_height, coeffs, salted):
    """vpbs_batch_*: commit stays in HBM; rows and Merkle paths fetched on demand match the oracle's
    leaves / MerkleTree::prove, verify against the cap, and download() equals the eager commit."""
    rng = np.random.default_rng(log_n * 31 + ncols)
    cols = rand_u64(rng, (ncols, 1 << log_n))
    m = (1 << log_n) << rate_bits
    salt = rand_u64(rng, (4, m)) if salted else None
    rb = V.commit_resident(cols, rate_bits, salted, cap_height, coeffs, ctx=ctx, salt=salt)
    ref = oracle.commit(cols, rate_bits, cap_height, coeffs, salt)
    assert np.array_equal(rb.merkle_tree.cap, ref["cap"])
    idx = rng.integers(0, m, size=min(28, m), dtype=np.uint64)
    rows = rb.merkle_tree.get_many(idx)
    proofs = rb.merkle_tree.prove_many(idx)
    for k, i in enumerate(idx):
        i = int(i)
        assert np.array_equal(rows[k], ref["leaves"][i])
        assert np.array_equal(proofs[k].siblings, oracle.merkle_prove(ref["digests"], m, cap_height, i))
        assert oracle.merkle_verify(rows[k], i, proofs[k].siblings, ref["cap"])
    j = int(idx[0])
    nat = V.reverse_bits(j, log_n + rate_bits)
    assert np.array_equal(rb.get_lde_values(nat), ref["leaves"][j][:ncols])
    eager = rb.download()
    assert np.array_equal(eager.merkle_tree.leaves, ref["leaves"])
    assert np.array_equal(eager.merkle_tree.digests, ref["digests"])
    if not coeffs:
        assert np.array_equal(eager.polynomials, ref["coeffs"])
    with pytest.raises(ValueError):
        rb.merkle_tree.get(m)
    rb.close()


@pytest.mark.parametrize("log_n,ncols,coeffs", [(13, 128, False), (14, 20, False), (13, 16, True),
                                               (12, 70, False), (20, 3, False)])
@pytest.mark.parametrize("threads", [0, 1, 4])
def test_pageable_host_columns_through_the_pinned_ring(V, oracle, log_n, ncols, coeffs, threads):
    """Host columns in ordinary (pageable) memory — what plonky2's prover passes to from_values —
    travel through the context's pinned staging ring (csrc/host_stage.h): column-chunked wide
    batches, narrow single-chunk batches and columns longer than a ring slot, with 0 (driver
    staging), 1 and 4 copy threads, and columns that are separate allocations.  The resident and
    the eager commit must equal the commit from pinned columns and, at the smaller sizes, the oracle."""
    import ctypes
    rng = np.random.default_rng(1000 * log_n + ncols)
    n = 1 << log_n
    cols = [rand_u64(rng, (n,)) for _ in range(ncols)]          # one allocation per column
    mat = np.stack(cols)
    with V.Context(0) as c:
        c.set_host_threads(threads)
        u64p, lib = V._lib.u64p, c.lib
        pin = lib.vpbs_host_alloc(mat.nbytes)
        pinned = np.ctypeslib.as_array((ctypes.c_uint64 * mat.size).from_address(pin)).reshape(mat.shape)
        pinned[:] = mat
        want = V.commit_resident(pinned, 3, False, 4, coeffs, ctx=c)
        colp = (u64p * ncols)(*[a.ctypes.data_as(u64p) for a in cols])
        cap = np.empty((16, 4), np.uint64)
        for _ in range(3):                                       # ring slots are reused across calls
            h = ctypes.c_void_p()
            c.check(lib.vpbs_batch_commit(c.handle, colp, ncols, log_n, 3, 4, int(coeffs), None,
                                          cap.ctypes.data_as(u64p), ctypes.byref(h), None))
            assert np.array_equal(cap, want.merkle_tree.cap)
            idx = rng.integers(0, n << 3, size=8, dtype=np.uint64)
            rows = np.empty((8, ncols), np.uint64)
            c.check(lib.vpbs_batch_get_leaves(h, idx.ctypes.data_as(u64p), 8, rows.ctypes.data_as(u64p)))
            assert np.array_equal(rows, want.merkle_tree.get_many(idx))
            lib.vpbs_batch_destroy(h)
        if log_n <= 14:
            eager = V.PolynomialBatch._commit(mat, 3, False, 4, coeffs, c, None, None)
            assert np.array_equal(eager.merkle_tree.cap, want.merkle_tree.cap)
            ref = oracle.commit(mat, 3, 4, coeffs, None)
            assert np.array_equal(cap, ref["cap"])
            assert np.array_equal(eager.merkle_tree.leaves, ref["leaves"])
            assert np.array_equal(eager.merkle_tree.digests, ref["digests"])
        want.close()
        lib.vpbs_host_free(pin)


def test_fri_proof_of_work_grind(V, ctx, oracle):
    """vpbs_pow_grind: the smallest witness found on the GPU is exactly the first one the oracle's
    permutation accepts, for the reference's proof_of_work_bits = 16 (plus the 0 extra bits of a
    64-bit field) and for an impossible target."""
    rng = np.random.default_rng(12)
    state = rng.integers(0, P, size=12, dtype=np.uint64)
    pos, bits = 5, 12
    w = V.fri_proof_of_work(state, pos, bits, ctx=ctx)
    assert w is not None
    def response(c):
        s = state.copy(); s[pos] = c
        return int(oracle.poseidon(s)[7])
    assert response(w) >> (64 - bits) == 0
    for c in range(0, w, max(1, w // 300)):     # no smaller qualifying witness (sampled) ...
        assert response(c) >> (64 - bits) != 0 or c == w
    first = next(c for c in range(0, w + 1) if response(c) >> (64 - bits) == 0) if w < 20000 else w
    assert first == w                            # ... and exhaustively when cheap
    assert V.fri_proof_of_work(state, pos, 16, ctx=ctx) is not None
    assert V.fri_proof_of_work(state, pos, 60, count=1 << 16, ctx=ctx) is None
    assert V.fri_proof_of_work(state, pos, bits, first_candidate=w + 1, count=1, ctx=ctx) in (None, w + 1)


@pytest.mark.parametrize("log_n,ncols", [(0, 3), (3, 5), (8, 9), (9, 2), (13, 20), (16, 7)])
def test_eval_ext2_openings(V, ctx, oracle, log_n, ncols):
    """Openings at extension-field points (OpeningSet::new's evaluations) vs the oracle's Horner."""
    rng = np.random.default_rng(log_n + 100 * ncols)
    coeffs = rand_u64(rng, (ncols, 1 << log_n))
    pts = rand_u64(rng, (3, 2))
    pts[2] = (5, 0)   # a base-field point: imaginary parts must come out 0 ... unless coefficients say otherwise
    got = V.eval_ext2(coeffs, pts, ctx)
    for p in range(3):
        assert np.array_equal(got[p], oracle.eval_ext2(coeffs, pts[p])), (log_n, ncols, p)
    assert (got[2][:, 1] == 0).all()


def test_resident_batch_openings(V, ctx, oracle):
    rng = np.random.default_rng(77)
    cols = rand_u64(rng, (20, 1 << 10))
    rb = V.commit_resident(cols, 3, False, 4, ctx=ctx)
    ref = oracle.commit(cols, 3, 4, False)
    zeta = rand_u64(rng, (2, 2))
    got = rb.eval_ext2(zeta)
    for p in range(2):
        assert np.array_equal(got[p], oracle.eval_ext2(ref["coeffs"], zeta[p]))
    rb.close()


@pytest.mark.parametrize("log_len,arity_bits,cap_height", [(4, 2, 1), (8, 4, 4), (12, 4, 4),
                                                            (19, 4, 4), (15, 4, 4), (6, 1, 0), (4, 4, 0)])
def test_fri_commit_phase_layer(V, ctx, oracle, log_len, arity_bits, cap_height):
    """One FRI reduction layer (tree of the bit-reversed, chunked values; fold with beta; next coset
    evaluations) vs the oracle; shapes incl. the N=1024 proof's first layer (2^19 values, arity 16)."""
    rng = np.random.default_rng(log_len * 10 + arity_bits)
    vals = rand_u64(rng, (1 << log_len, 2), 0.02)
    tree = V.fri_layer_commit(vals, arity_bits, cap_height, ctx)
    ref = oracle.fri_layer_commit(vals, arity_bits, cap_height)
    assert np.array_equal(tree.cap, ref["cap"])
    assert np.array_equal(tree.digests, ref["digests"])
    assert np.array_equal(tree.leaves, ref["leaves"])
    if log_len <= 15:
        beta = rng.integers(0, P, size=2, dtype=np.uint64)
        shift = int(oracle.gl_pow(7, 1 << arity_bits))
        co, vo = V.fri_fold(vals, arity_bits, beta, shift, ctx)
        rco, rvo = oracle.fri_fold(vals, arity_bits, beta, shift)
        assert np.array_equal(co, rco)
        assert np.array_equal(vo, rvo)


def test_fri_layers_chain_like_the_prover(V, ctx, oracle):
    """Three chained layers as fri_committed_trees runs them (arity 16, shift <- shift^16):
    values_k = coeffs_k.coset_fft(shift_k) must stay consistent layer after layer."""
    rng = np.random.default_rng(5)
    log_len, a = 14, 4
    coeffs = rng.integers(0, P, size=(1 << log_len, 2), dtype=np.uint64)
    shift = 7
    values = np.stack([oracle.coset_fft(coeffs[:, 0].copy(), shift), oracle.coset_fft(coeffs[:, 1].copy(), shift)], 1)
    for _ in range(3):
        tree = V.fri_layer_commit(values, a, min(4, log_len - a), ctx)
        assert np.array_equal(tree.cap, oracle.fri_layer_commit(values, a, min(4, log_len - a))["cap"])
        beta = rng.integers(0, P, size=2, dtype=np.uint64)
        shift = oracle.gl_pow(shift, 1 << a)
        coeffs, values = V.fri_fold(coeffs, a, beta, shift, ctx)
        log_len -= a
        want = np.stack([oracle.coset_fft(coeffs[:, 0].copy(), shift), oracle.coset_fft(coeffs[:, 1].copy(), shift)], 1)
        assert np.array_equal(values, want)


# ------------------------------------------------------------------------------ permutation argument
def _perm_inputs(rng, num_routed, log_n, noncanonical=True):
    n = 1 << log_n
    wires = rand_u64(rng, (num_routed, n)) if noncanonical else rng.integers(0, P, size=(num_routed, n), dtype=np.uint64)
    sigmas = rand_u64(rng, (num_routed, n))
    return wires, sigmas


@pytest.mark.parametrize("num_routed,log_n,max_degree,nch", [
    (80, 10, 8, 2), (80, 0, 8, 2), (8, 3, 8, 1), (5, 2, 2, 3), (7, 5, 3, 2), (20, 9, 8, 2),
    (33, 13, 8, 2), (80, 16, 8, 2)])
def test_zs_partial_products_against_oracle(V, ctx, oracle, num_routed, log_n, max_degree, nch):
    """[P2] all_wires_permutation_partial_products through the C ABI (host in / host out) against the
    oracle, incl. the N=1024 step's shape (80 routed wires x 2^16 rows, 2 challenges -> 20 columns),
    ragged last chunks, one row, non-canonical inputs."""
    rng = np.random.default_rng(num_routed * 100 + log_n)
    wires, sigmas = _perm_inputs(rng, num_routed, log_n)
    k_is = V.get_unique_coset_shifts(1 << log_n, num_routed)
    # challenges are generic field elements: with an edge value for beta / gamma (say beta = 1,
    # gamma = 0) the edge values among the 2^16 x 80 wires / sigmas do hit wire + beta sigma + gamma
    # == 0, where the library (like upstream) refuses — that case has its own test below
    betas, gammas = rand_u64(rng, nch, edge_frac=0), rand_u64(rng, nch, edge_frac=0)
    sg = V.Sigmas(sigmas, k_is, ctx)
    got = V.all_wires_permutation_partial_products(wires, sg, betas, gammas, max_degree)
    K = -(-num_routed // max_degree)
    assert got.shape == (nch * K, 1 << log_n)
    for c in range(nch):
        ref = oracle.zs_partial_products(wires, sigmas, k_is, max_degree, betas[c], gammas[c])
        assert np.array_equal(got[c], ref[0]), "Z of challenge %d" % c
        assert np.array_equal(got[nch + c * (K - 1): nch + (c + 1) * (K - 1)], ref[1:]), "partial products"
    sg.close()


def test_zs_of_a_real_permutation_closes(V, ctx):
    """Size-independent property at full size: when the wires satisfy a copy-constraint permutation,
    the running product returns to 1 — Z(g^n) = Z(1) — i.e. the last row's full chunk product times
    Z(x_{n-1}) is 1.  Checked through the last partial product and the chunk values on the device."""
    rng = np.random.default_rng(4)
    num_routed, log_n, deg = 16, 12, 8
    n = 1 << log_n
    k_is = V.get_unique_coset_shifts(n, num_routed)
    w = pow(7, (P - 1) >> log_n, P)
    sub = np.array([pow(w, i, P) for i in range(n)], dtype=object)
    ncell = num_routed * n
    perm = rng.permutation(ncell)
    # wire values constant along every cycle of the permutation
    label = np.full(ncell, -1, dtype=np.int64)
    vals = rng.integers(0, P, size=ncell, dtype=np.uint64)
    for c0 in range(ncell):
        c = c0
        while label[c] < 0:
            label[c] = c0
            c = perm[c]
    wires = vals[label].reshape(num_routed, n)
    tj, ti = np.divmod(perm, n)
    sig = np.array([int(k_is[j]) * int(sub[i]) % P for j, i in zip(tj, ti)], dtype=np.uint64).reshape(num_routed, n)
    sg = V.Sigmas(sig, k_is, ctx)
    beta, gamma = rand_u64(rng, 1), rand_u64(rng, 1)
    out = V.all_wires_permutation_partial_products(wires, sg, beta, gamma, deg)
    K = num_routed // deg
    # Z(x_{n-1}) * prod of the last row's quotients == 1: recompute that row's quotient product
    i = n - 1
    num = den = 1
    b, g = int(beta[0]) % P, int(gamma[0]) % P
    for j in range(num_routed):
        num = num * ((int(wires[j, i]) + b * int(k_is[j]) * int(sub[i]) + g) % P) % P
        den = den * ((int(wires[j, i]) + b * int(sig[j, i]) + g) % P) % P
    assert int(out[0, i]) * num % P * pow(den, P - 2, P) % P == 1
    assert int(out[0, 0]) == 1 and K == 2
    sg.close()


def test_zs_zero_denominator_is_an_error(V, ctx):
    """plonky2's batch_multiplicative_inverse panics on a zero denominator; the ABI returns
    VPBS_ERR_ARG (ValueError in the mirror)."""
    rng = np.random.default_rng(8)
    num_routed, log_n = 8, 4
    wires, sigmas = _perm_inputs(rng, num_routed, log_n, noncanonical=False)
    k_is = V.get_unique_coset_shifts(1 << log_n, num_routed)
    beta, gamma = 5, 11
    # force wire + beta * sigma + gamma == 0 at (wire 3, row 7)
    sigmas[3, 7] = (P - (int(wires[3, 7]) + gamma) % P) * pow(beta, P - 2, P) % P
    sg = V.Sigmas(sigmas, k_is, ctx)
    with pytest.raises(ValueError):
        V.all_wires_permutation_partial_products(wires, sg, [beta], [gamma], 8)
    sg.close()


@pytest.mark.parametrize("log_n,ncols,num_routed", [(10, 135, 80), (6, 9, 8), (13, 20, 20)])
def test_resident_zs_commit_matches_host_pipeline(V, ctx, oracle, log_n, ncols, num_routed):
    """prove() steps 2-5 with both batches resident: wires commit -> Z / partial products computed
    from the wires batch's coefficients in HBM -> committed as a new resident batch.  The cap, opened
    rows and coefficients equal the oracle's from_values on the oracle's Z columns."""
    rng = np.random.default_rng(log_n + ncols)
    n = 1 << log_n
    wires = rand_u64(rng, (ncols, n))
    sigmas = rand_u64(rng, (num_routed, n))
    k_is = V.get_unique_coset_shifts(n, num_routed)
    betas, gammas = rand_u64(rng, 2, edge_frac=0), rand_u64(rng, 2, edge_frac=0)
    wb = V.commit_resident(wires, 3, False, 4, ctx=ctx)
    sg = V.Sigmas(sigmas, k_is, ctx)
    zb = V.commit_zs_partial_products(wb, sg, betas, gammas, 8, 3, 4)
    K = -(-num_routed // 8)
    per = [oracle.zs_partial_products(wires[:num_routed], sigmas, k_is, 8, betas[c], gammas[c]) for c in range(2)]
    zcols = np.concatenate([np.stack([per[0][0], per[1][0]]), per[0][1:], per[1][1:]])
    assert zcols.shape[0] == 2 * K == zb.ncols
    ref = oracle.commit(zcols, 3, 4)
    assert np.array_equal(zb.merkle_tree.cap, ref["cap"])
    m = n << 3
    idx = rng.integers(0, m, size=12, dtype=np.uint64)
    rows = zb.merkle_tree.get_many(idx)
    for k, i in enumerate(idx):
        assert np.array_equal(rows[k], ref["leaves"][int(i)])
    eager = zb.download()
    assert np.array_equal(eager.polynomials, ref["coeffs"])
    assert np.array_equal(eager.merkle_tree.digests, ref["digests"])
    zb.close(); wb.close(); sg.close()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("log_n,ncols,num_routed,coeffs", [(10, 135, 80, False), (13, 20, 20, False), (6, 16, 8, True)])
def test_sharded_resident_batches_one_proof_over_several_ranks(V, oracle, world, log_n, ncols, num_routed, coeffs):
    """vpbs_ctx_set_shard: `world` contexts (here all on one GPU) each hold the row range of their
    rank of the SAME commit — wires batch from host columns, Z / partial products computed and
    committed from it — and together reproduce the unsharded commit exactly: the union of the caps,
    opened rows, Merkle paths (verified against the full cap), LDE row blocks, downloaded shard rows
    and digests; rows of another shard are refused; a shard count the commit cannot be split into is
    an argument error."""
    rng = np.random.default_rng(100 * world + log_n)
    n, m = 1 << log_n, (1 << log_n) << 3
    cols = rand_u64(rng, (ncols, n))
    sigmas = rand_u64(rng, (num_routed, n))
    k_is = V.get_unique_coset_shifts(n, num_routed)
    betas, gammas = rand_u64(rng, 2, edge_frac=0), rand_u64(rng, 2, edge_frac=0)
    ref = oracle.commit(cols, 3, 4, coeffs, None)
    zref = None
    if not coeffs:
        per = [oracle.zs_partial_products(cols[:num_routed], sigmas, k_is, 8, betas[c], gammas[c]) for c in range(2)]
        zcols = np.concatenate([np.stack([per[0][0], per[1][0]]), per[0][1:], per[1][1:]])
        zref = oracle.commit(zcols, 3, 4)
    cap_union = np.zeros((16, 4), np.uint64)
    zcap_union = np.zeros((16, 4), np.uint64)
    sub_dig = 2 * (m >> 4) - 2
    for rank in range(world):
        with V.Context(0) as c:
            c.set_shard(rank, world)
            rb = V.commit_resident(cols, 3, False, 4, coeffs, ctx=c)
            first, nl = rb.shard
            assert (first, nl) == (rank * m // world, m // world)
            cap = rb.merkle_tree.cap
            own = slice(rank * 16 // world, (rank + 1) * 16 // world)
            assert np.array_equal(cap[own], ref["cap"][own])
            mask = np.ones(16, bool); mask[own] = False
            assert not cap[mask].any()
            cap_union[own] = cap[own]
            idx = rng.integers(first, first + nl, size=12, dtype=np.uint64)
            idx[0], idx[1] = first, first + nl - 1
            rows = rb.merkle_tree.get_many(idx)
            proofs = rb.merkle_tree.prove_many(idx)
            for k, i in enumerate(idx):
                i = int(i)
                assert np.array_equal(rows[k], ref["leaves"][i])
                assert np.array_equal(proofs[k].siblings, oracle.merkle_prove(ref["digests"], m, 4, i))
                assert oracle.merkle_verify(rows[k], i, proofs[k].siblings, ref["cap"])
            other = (first + nl) % m
            with pytest.raises(ValueError):
                rb.merkle_tree.get_many(np.array([other], np.uint64))
            with pytest.raises(ValueError):
                rb.merkle_tree.prove_many(np.array([other], np.uint64))
            # natural rows i with bitrev(i) inside the shard: i = bitrev(first) + j * world
            i0 = V.reverse_bits(first, log_n + 3)
            got = rb.get_lde_rows(i0, world, 9)
            want = np.array([ref["leaves"][V.reverse_bits(i0 + j * world, log_n + 3)][:ncols] for j in range(9)], np.uint64)
            assert np.array_equal(got, want)
            if world > 1:
                with pytest.raises(ValueError):
                    rb.get_lde_rows(i0, 1, 2)
            eager = rb.download()
            assert np.array_equal(eager.merkle_tree.leaves, ref["leaves"][first:first + nl])
            assert np.array_equal(eager.merkle_tree.digests,
                                  ref["digests"][own.start * sub_dig:own.stop * sub_dig])
            if not coeffs:
                assert np.array_equal(eager.polynomials, ref["coeffs"])
                sg = V.Sigmas(sigmas, k_is, c)
                zb = V.commit_zs_partial_products(rb, sg, betas, gammas, 8, 3, 4)
                assert zb.shard == (first, nl)
                zcap_union[own] = zb.merkle_tree.cap[own]
                zrows = zb.merkle_tree.get_many(idx)
                for k, i in enumerate(idx):
                    assert np.array_equal(zrows[k], zref["leaves"][int(i)])
                opens = rb.eval_ext2(np.array([[3, 5]], np.uint64))   # coefficient consumers are unaffected
                c.set_shard(0, 1)
                full = V.commit_resident(cols, 3, False, 4, coeffs, ctx=c)
                assert np.array_equal(opens, full.eval_ext2(np.array([[3, 5]], np.uint64)))
                full.close(); zb.close(); sg.close()
            rb.close()
    assert np.array_equal(cap_union, ref["cap"])
    if zref is not None:
        assert np.array_equal(zcap_union, zref["cap"])


def test_all_oracles_opened_in_one_round_trip(V, ctx, oracle):
    """vpbs_batches_eval_ext2 / vpbs_batches_open (OpeningSet::new and initial_trees_proof over all
    FRI oracles at once) equal the per-batch calls and the oracle."""
    rng = np.random.default_rng(77)
    log_n, widths = 9, [85, 135, 20, 16]
    m = (1 << log_n) << 3
    mats = [rand_u64(rng, (w, 1 << log_n)) for w in widths]
    batches = [V.commit_resident(a, 3, False, 4, k == 3, ctx=ctx) for k, a in enumerate(mats)]
    pts = rand_u64(rng, (2, 2))
    got = V.open_all_at_points(batches, pts)
    for b, g in zip(batches, got):
        assert np.array_equal(g, b.eval_ext2(pts))
    refs = [oracle.commit(mats[k], 3, 4, k == 3, None) for k in range(4)]
    for count in (28, 13, 1):          # odd counts x odd widths: every region stays 16-byte aligned
        idx = rng.integers(0, m, size=count, dtype=np.uint64)
        opened = V.open_all_at_leaves(batches, idx)
        for k, (b, (rows, sibs)) in enumerate(zip(batches, opened)):
            for q, i in enumerate(idx):
                assert np.array_equal(rows[q], refs[k]["leaves"][int(i)])
                assert np.array_equal(sibs[q], oracle.merkle_prove(refs[k]["digests"], m, 4, int(i)))
    other = V.commit_resident(mats[0], 2, False, 4, ctx=ctx)      # different tree shape
    with pytest.raises(ValueError):
        V.open_all_at_leaves([batches[0], other], idx[:2])
    with pytest.raises(ValueError):
        V.open_all_at_leaves(batches, np.array([m], np.uint64))
    for b in batches + [other]:
        b.close()


def test_shard_count_the_commit_cannot_be_split_into(V):
    with V.Context(0) as c:
        with pytest.raises(ValueError):
            c.set_shard(0, 3)
        with pytest.raises(ValueError):
            c.set_shard(2, 2)
        c.set_shard(1, 16)          # more shards than LDE blocks (2^rate_bits = 8)
        with pytest.raises(ValueError):
            V.commit_resident(np.ones((5, 64), np.uint64), 3, False, 4, ctx=c)
        c.set_shard(1, 4)           # more shards than cap subtrees (cap_height = 1)
        with pytest.raises(ValueError):
            V.commit_resident(np.ones((5, 64), np.uint64), 3, False, 1, ctx=c)


@pytest.mark.parametrize("log_n,ncols,num_routed,deg,qdb,first_sigma", [(6, 12, 8, 4, 2, 3), (10, 135, 80, 8, 3, 5),
                                                                         (5, 9, 7, 2, 3, 0), (13, 20, 16, 8, 1, 2)])
def test_quotient_polys_against_oracle(V, ctx, oracle, log_n, ncols, num_routed, deg, qdb, first_sigma):
    """prove() steps 6-7, gate-independent part (vpbs_batch_quotient_polys): from the resident wires,
    constants/sigmas and Z batches to the committed quotient chunks, with and without alpha-reduced
    gate terms; coefficients, cap and opened rows equal the oracle's compute_quotient_polys + from_coeffs."""
    rng = np.random.default_rng(17 * log_n + ncols)
    n = 1 << log_n
    wires = rand_u64(rng, (ncols, n))
    cs = rand_u64(rng, (first_sigma + num_routed + 1, n))
    sigma_vals = cs[first_sigma:first_sigma + num_routed]
    k_is = V.get_unique_coset_shifts(n, num_routed)
    betas, gammas, alphas = (rand_u64(rng, 2, edge_frac=0) for _ in range(3))
    wb = V.commit_resident(wires, 3, False, 4, ctx=ctx)
    cb = V.commit_resident(cs, 3, False, 4, ctx=ctx)
    sg = V.Sigmas(sigma_vals, k_is, ctx)
    zb = V.commit_zs_partial_products(wb, sg, betas, gammas, deg, 3, 4)
    wc, cc, zc = wb.download().polynomials, cb.download().polynomials, zb.download().polynomials
    gate = rand_u64(rng, (2, n << qdb))
    for gt in (None, gate):
        qb = V.commit_quotient_polys(cb, first_sigma, wb, zb, k_is, deg, qdb, betas, gammas, alphas, 3, 4, gt)
        want = oracle.quotient_polys(wc[:num_routed], cc[first_sigma:first_sigma + num_routed], zc, k_is, deg,
                                     qdb, betas, gammas, alphas, gt)
        assert qb.ncols == 2 << qdb == want.shape[0]
        ref = oracle.commit(want, 3, 4, True)
        assert np.array_equal(qb.merkle_tree.cap, ref["cap"])
        got = qb.download()
        assert np.array_equal(got.polynomials, want)
        assert np.array_equal(got.merkle_tree.leaves, ref["leaves"])
        qb.close()
    with pytest.raises(ValueError):
        V.commit_quotient_polys(cb, first_sigma + 2, wb, zb, k_is, deg, qdb, betas, gammas, alphas, 3, 4)
    with pytest.raises(ValueError):
        V.commit_quotient_polys(cb, first_sigma, wb, zb, k_is, deg, 4, betas, gammas, alphas, 3, 4)
    for b in (wb, cb, zb):
        b.close()
    sg.close()


@pytest.mark.parametrize("log_n,ncols,ncs,qdb", [(6, 12, 12, 2), (10, 40, 30, 3)])
def test_gate_program_against_oracle(V, ctx, oracle, log_n, ncols, ncs, qdb):
    """Gate constraints as a program evaluated on the device (vpbs_gate_program_upload +
    vpbs_batch_quotient_polys): random straight-line programs over wires, constants, immediates and
    public_inputs_hash; the committed quotient chunks equal the oracle's, whose gate terms come from the
    oracle's own interpreter."""
    from test_oracle_golden import _random_gate_program
    rng = np.random.default_rng(5 * log_n + ncols)
    n, num_routed, deg = 1 << log_n, 8, 4
    wires, cs = rand_u64(rng, (ncols, n)), rand_u64(rng, (ncs, n))
    first_sigma = ncs - num_routed
    k_is = V.get_unique_coset_shifts(n, num_routed)
    betas, gammas, alphas = (rand_u64(rng, 2, edge_frac=0) for _ in range(3))
    pih = rand_u64(rng, 4)
    wb = V.commit_resident(wires, 3, False, 4, ctx=ctx)
    cb = V.commit_resident(cs, 3, False, 4, ctx=ctx)
    sg = V.Sigmas(cs[first_sigma:], k_is, ctx)
    zb = V.commit_zs_partial_products(wb, sg, betas, gammas, deg, 3, 4)
    wc, cc, zc = wb.download().polynomials, cb.download().polynomials, zb.download().polynomials
    bld = _random_gate_program(V, rng, ncols, ncs, ngates=4, nops=60, nconstraints=9)
    prog = bld.build(ctx)
    qb = V.commit_quotient_polys(cb, first_sigma, wb, zb, k_is, deg, qdb, betas, gammas, alphas, 3, 4,
                                 program=prog, public_inputs_hash=pih)
    gt = oracle.gate_program_eval(bld.code, bld.imms, bld.nregs, bld.num_constraints, wc, cc, qdb, pih, alphas)
    want = oracle.quotient_polys(wc[:num_routed], cc[first_sigma:], zc, k_is, deg, qdb, betas, gammas, alphas, gt)
    assert np.array_equal(qb.download().polynomials, want)
    assert np.array_equal(qb.merkle_tree.cap, oracle.commit(want, 3, 4, True)["cap"])
    with pytest.raises(ValueError):          # gate_terms and a program at once
        V.commit_quotient_polys(cb, first_sigma, wb, zb, k_is, deg, qdb, betas, gammas, alphas, 3, 4,
                                gate_terms=gt, program=prog)
    bad = V.GateProgramBuilder()
    bad.emit(0, bad.wire(ncols))             # a wire column the batch does not have
    bad.end_gate(bad.imm(1))
    bp = bad.build(ctx)
    with pytest.raises(ValueError):
        V.commit_quotient_polys(cb, first_sigma, wb, zb, k_is, deg, qdb, betas, gammas, alphas, 3, 4, program=bp)
    with pytest.raises(ValueError):          # register index beyond nregs
        V.GateProgram(np.array([0 | 5 << 8 | 3 << 16 | 3 << 20], np.uint64), np.zeros(1, np.uint64), 2, 1, ctx)
    with pytest.raises(ValueError):          # a register read before anything wrote it
        V.GateProgram(np.array([3 | 0 << 16 | 1 << 24], np.uint64), np.zeros(0, np.uint64), 2, 1, ctx)
    for b in (wb, cb, zb, qb):
        b.close()
    sg.close(); prog.close(); bp.close()


def test_quotient_polys_step_shapes_full_size(V, ctx, oracle):
    """BASELINE.json configs[2] stand-in, steps 4-7 at full size: 135 wire columns (80 routed), the 85-column
    constants/sigmas batch, Z / partial products on the device, then the quotient (2^19 points, alpha-reduced
    gate values from the host): the 16 committed chunks and their cap equal the oracle's."""
    n, num_routed = 1 << 16, 80
    wires = V.synthetic_columns(135, n, 0x5EED0000 + 135)
    cs = V.synthetic_columns(85, n, 0x5EED0000 + 85)
    k_is = V.get_unique_coset_shifts(n, num_routed)
    rng = np.random.default_rng(8)
    betas, gammas, alphas = (rng.integers(1, P, size=2, dtype=np.uint64) for _ in range(3))
    gate = rng.integers(0, 2**64, size=(2, n << 3), dtype=np.uint64)
    wb, cb = V.commit_resident(wires, 3, False, 4, ctx=ctx), V.commit_resident(cs, 3, False, 4, ctx=ctx)
    sg = V.Sigmas(cs[5:], k_is, ctx)
    zb = V.commit_zs_partial_products(wb, sg, betas, gammas, 8, 3, 4)
    qb = V.commit_quotient_polys(cb, 5, wb, zb, k_is, 8, 3, betas, gammas, alphas, 3, 4, gate_terms=gate)
    want = oracle.quotient_polys(wb.coefficients()[:num_routed], cb.coefficients()[5:], zb.coefficients(), k_is,
                                 8, 3, betas, gammas, alphas, gate)
    assert np.array_equal(qb.coefficients(), want)
    assert np.array_equal(qb.merkle_tree.cap, oracle.commit(want, 3, 4, True)["cap"])
    for b in (wb, cb, zb, qb):
        b.close()
    sg.close()


def test_gate_program_step_shapes_full_size(V, ctx, oracle):
    """The gate-program interpreter at the N=1024 step's shapes (2^19 points, 135 wire and 85 constant
    columns, a 400-instruction random program with MADs over every operand kind): quotient chunks equal the
    oracle's, whose gate terms come from its own interpreter."""
    from test_oracle_golden import _random_gate_program
    n, num_routed = 1 << 16, 80
    wires = V.synthetic_columns(135, n, 0x5EED0000 + 135)
    cs = V.synthetic_columns(85, n, 0x5EED0000 + 85)
    k_is = V.get_unique_coset_shifts(n, num_routed)
    rng = np.random.default_rng(18)
    betas, gammas, alphas = (rng.integers(1, P, size=2, dtype=np.uint64) for _ in range(3))
    pih = rng.integers(0, P, size=4, dtype=np.uint64)
    bld = _random_gate_program(V, rng, 135, 85, ngates=6, nops=60, nconstraints=20)
    prog = bld.build(ctx)
    wb, cb = V.commit_resident(wires, 3, False, 4, ctx=ctx), V.commit_resident(cs, 3, False, 4, ctx=ctx)
    sg = V.Sigmas(cs[5:], k_is, ctx)
    zb = V.commit_zs_partial_products(wb, sg, betas, gammas, 8, 3, 4)
    qb = V.commit_quotient_polys(cb, 5, wb, zb, k_is, 8, 3, betas, gammas, alphas, 3, 4, program=prog,
                                 public_inputs_hash=pih)
    wc, cc = wb.coefficients(), cb.coefficients()
    gt = oracle.gate_program_eval(bld.code, bld.imms, bld.nregs, bld.num_constraints, wc, cc, 3, pih, alphas)
    want = oracle.quotient_polys(wc[:num_routed], cc[5:], zb.coefficients(), k_is, 8, 3, betas, gammas, alphas, gt)
    assert np.array_equal(qb.coefficients(), want)
    for b in (wb, cb, zb, qb):
        b.close()
    sg.close(); prog.close()


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_quotient_values_and_commit(V, oracle, world):
    """The two halves of the device quotient on sharded batches (vpbs_batch_quotient_values ->
    sum over the ranks -> vpbs_quotient_commit_values), `world` contexts on one GPU standing in for the
    ranks: the summed values, the union of the caps and every rank's rows equal the unsharded
    vpbs_batch_quotient_polys — with a gate program and with host gate terms."""
    import ctypes
    import torch
    from test_oracle_golden import _random_gate_program
    rng = np.random.default_rng(world)
    log_n, ncols, ncs, num_routed, deg, qdb = 9, 20, 14, 8, 4, 3
    n, q = 1 << log_n, (1 << log_n) << qdb
    wires, cs = rand_u64(rng, (ncols, n)), rand_u64(rng, (ncs, n))
    first_sigma = ncs - num_routed
    k_is = V.get_unique_coset_shifts(n, num_routed)
    betas, gammas, alphas = (rand_u64(rng, 2, edge_frac=0) for _ in range(3))
    pih = rand_u64(rng, 4)
    gate_host = rand_u64(rng, (2, q))
    bld = _random_gate_program(V, rng, ncols, ncs, ngates=2, nops=30, nconstraints=6)
    u64p = V._lib.u64p
    ptr = lambda a: a.ctypes.data_as(u64p)
    for use_program in (True, False):
        with V.Context(0) as c0:
            wb, cb = V.commit_resident(wires, 3, False, 4, ctx=c0), V.commit_resident(cs, 3, False, 4, ctx=c0)
            sg = V.Sigmas(cs[first_sigma:], k_is, c0)
            zb = V.commit_zs_partial_products(wb, sg, betas, gammas, deg, 3, 4)
            prog0 = bld.build(c0) if use_program else None
            ref = V.commit_quotient_polys(cb, first_sigma, wb, zb, k_is, deg, qdb, betas, gammas, alphas, 3, 4,
                                          gate_terms=None if use_program else gate_host, program=prog0,
                                          public_inputs_hash=pih)
            ref_cap = ref.merkle_tree.cap.copy()
            ref_rows = ref.download().merkle_tree.leaves
            for b in (wb, cb, zb, ref):
                b.close()
            sg.close()
            if prog0:
                prog0.close()
        ranks, total = [], torch.zeros(2 * q, dtype=torch.int64, device="cuda")
        for rank in range(world):
            c = V.Context(0)
            c.set_shard(rank, world)
            wb, cb = V.commit_resident(wires, 3, False, 4, ctx=c), V.commit_resident(cs, 3, False, 4, ctx=c)
            sg = V.Sigmas(cs[first_sigma:], k_is, c)
            zb = V.commit_zs_partial_products(wb, sg, betas, gammas, deg, 3, 4)
            prog = bld.build(c) if use_program else None
            with pytest.raises(ValueError):      # the one-call form refuses sharded batches
                V.commit_quotient_polys(cb, first_sigma, wb, zb, k_is, deg, qdb, betas, gammas, alphas, 3, 4)
            vals = torch.empty(2 * q, dtype=torch.int64, device="cuda")
            gtp = None if use_program else (u64p * 2)(ptr(gate_host[0]), ptr(gate_host[1]))
            c.check(c.lib.vpbs_batch_quotient_values(cb.handle, first_sigma, wb.handle, zb.handle, ptr(k_is),
                                                     num_routed, deg, qdb, ptr(betas), ptr(gammas), ptr(alphas), 2,
                                                     gtp, prog.handle if prog else None, ptr(pih), vals.data_ptr()))
            torch.cuda.synchronize()
            total += vals
            ranks.append((c, wb, cb, zb, sg, prog))
        torch.cuda.synchronize()
        cap_union = np.zeros((16, 4), np.uint64)
        for rank, (c, wb, cb, zb, sg, prog) in enumerate(ranks):
            cap = np.empty((16, 4), np.uint64)
            h = ctypes.c_void_p()
            c.check(c.lib.vpbs_quotient_commit_values(c.handle, total.data_ptr(), 2, log_n, qdb, 3, 4, ptr(cap),
                                                      ctypes.byref(h), None))
            own = slice(rank * 16 // world, (rank + 1) * 16 // world)
            cap_union[own] = cap[own]
            qb = V.ResidentPolynomialBatch(c, h, cap, 2 << qdb, log_n, 3, False, None)
            first, nl = qb.shard
            idx = rng.integers(first, first + nl, size=6, dtype=np.uint64)
            assert np.array_equal(qb.merkle_tree.get_many(idx), ref_rows[idx])
            for b in (qb, wb, cb, zb):
                b.close()
            sg.close()
            if prog:
                prog.close()
            c.close()
        assert np.array_equal(cap_union, ref_cap)


def test_quotient_polys_plonk_identity_with_gate_program(V, ctx):
    """A valid mini-circuit proved end to end on the device side: three gate types selected per row by a
    selector polynomial (arithmetic: out = c0 m0 m1 + c1 add, twice per row; constant: wire_k = const_k;
    public input: wires 0..3 = public_inputs_hash) with plonky2's filter prod_{i != g}(i - s), copy
    constraints among equal cells, Z / partial products, and the quotient computed from the gate
    program + the permutation terms.  At a random zeta:
        sum_j alpha^j term_j(zeta) = (zeta^n - 1) sum_k zeta^(n k) t_k(zeta)."""
    from oracle import model as M
    rng = np.random.default_rng(11)
    log_n, num_routed, deg, qdb = 7, 8, 4, 3
    n, K = 1 << log_n, 2
    ri = lambda: int(rng.integers(0, P, dtype=np.uint64))
    pih = [ri() for _ in range(4)]
    sel = rng.integers(0, 3, size=n)
    c0, c1 = [ri() for _ in range(n)], [ri() for _ in range(n)]
    wires = [[ri() for _ in range(n)] for _ in range(num_routed)]
    for i in range(n):
        if sel[i] == 0:                                   # ArithmeticGate, 2 operations
            for o in range(2):
                m0, m1, ad = wires[4 * o][i], wires[4 * o + 1][i], wires[4 * o + 2][i]
                wires[4 * o + 3][i] = (c0[i] * m0 * m1 + c1[i] * ad) % P
        elif sel[i] == 1:                                 # ConstantGate, 2 constants
            c0[i] = 5
            wires[0][i], wires[1][i] = c0[i], c1[i]
        else:                                             # PublicInputGate
            for k in range(4):
                wires[k][i] = pih[k]
    # copy constraints: all wire-0 cells of the constant rows hold 5 -> one cycle; the public-input
    # cells of column 2 hold pih[2] -> another; everything else is a fixed point
    w = pow(7, (P - 1) >> log_n, P)
    sub = [pow(w, i, P) for i in range(n)]
    k_is = V.get_unique_coset_shifts(n, num_routed)
    sig = [[int(k_is[j]) * sub[i] % P for i in range(n)] for j in range(num_routed)]
    for col, rows in ((0, [i for i in range(n) if sel[i] == 1]), (2, [i for i in range(n) if sel[i] == 2])):
        for a, b_ in zip(rows, rows[1:] + rows[:1]):
            sig[col][a] = int(k_is[col]) * sub[b_] % P
    wires_np = np.array(wires, dtype=np.uint64)
    cs_np = np.array([list(map(int, sel)), c0, c1] + sig, dtype=np.uint64)
    betas, gammas, alphas = (rng.integers(1, P, size=2, dtype=np.uint64) for _ in range(3))
    wb = V.commit_resident(wires_np, 3, False, 4, ctx=ctx)
    cb = V.commit_resident(cs_np, 3, False, 4, ctx=ctx)
    sg = V.Sigmas(cs_np[3:], k_is, ctx)
    zb = V.commit_zs_partial_products(wb, sg, betas, gammas, deg, 3, 4)
    b = V.GateProgramBuilder()
    for o in range(2):                                    # gate 0
        prod = b.mul(b.mul(b.wire(4 * o), b.wire(4 * o + 1)), b.const(1))
        b.emit(o, b.sub(b.wire(4 * o + 3), b.add(prod, b.mul(b.wire(4 * o + 2), b.const(2)))))
    b.end_gate(b.selector_filter(0, 0, range(3), False))
    for k in range(2):                                    # gate 1
        b.emit(k, b.sub(b.wire(k), b.const(1 + k)))
    b.end_gate(b.selector_filter(0, 1, range(3), False))
    for k in range(4):                                    # gate 2
        b.emit(k, b.sub(b.wire(k), b.pih(k)))
    b.end_gate(b.selector_filter(0, 2, range(3), False))
    prog = b.build(ctx)
    qb = V.commit_quotient_polys(cb, 3, wb, zb, k_is, deg, qdb, betas, gammas, alphas, 3, 4, program=prog,
                                 public_inputs_hash=np.array(pih, dtype=np.uint64))
    t = qb.download().polynomials
    wc, cc, zc = wb.download().polynomials, cb.download().polynomials, zb.download().polynomials
    assert int(zc[0][0]) != 1 or zc[0][1:].any()          # Z is not the constant 1: the cycles are real
    ev = lambda coeffs, x: M.evaluate([int(v) for v in coeffs], x)
    zeta = int(rng.integers(2, P, dtype=np.uint64))
    gz = w * zeta % P
    wz, cz, zz = [ev(c, zeta) for c in wc], [ev(c, zeta) for c in cc], [ev(c, zeta) for c in zc]
    zg = [ev(zc[c], gz) for c in range(2)]
    zh = (pow(zeta, n, P) - 1) % P
    l0 = zh * pow(n * (zeta - 1) % P, P - 2, P) % P
    terms = [l0 * (zz[c] - 1) % P for c in range(2)]
    for c in range(2):
        accs = [zz[c]] + [zz[2 + c * (K - 1) + k] for k in range(K - 1)] + [zg[c]]
        for k in range(K):
            num = den = 1
            for j in range(k * deg, (k + 1) * deg):
                num = num * (wz[j] + int(betas[c]) * int(k_is[j]) * zeta + int(gammas[c])) % P
                den = den * (wz[j] + int(betas[c]) * cz[3 + j] + int(gammas[c])) % P
            terms.append((accs[k] * num - accs[k + 1] * den) % P)
    s = cz[0]
    filt = [(1 - s) * (2 - s) % P, (0 - s) * (2 - s) % P, (0 - s) * (1 - s) % P]
    gate_c = [0] * 4
    for o in range(2):
        gate_c[o] += filt[0] * (wz[4 * o + 3] - (cz[1] * wz[4 * o] * wz[4 * o + 1] + cz[2] * wz[4 * o + 2]))
    for k in range(2):
        gate_c[k] += filt[1] * (wz[k] - cz[1 + k])
    for k in range(4):
        gate_c[k] += filt[2] * (wz[k] - pih[k])
    terms += [g % P for g in gate_c]
    for c in range(2):
        lhs = sum(tm * pow(int(alphas[c]), j, P) for j, tm in enumerate(terms)) % P
        rhs = zh * sum(pow(zeta, n * k, P) * ev(t[(c << qdb) + k], zeta) for k in range(1 << qdb)) % P
        assert lhs == rhs
    for x in (wb, cb, zb, qb):
        x.close()
    sg.close(); prog.close()


def test_quotient_polys_plonk_identity(V, ctx):
    """The property the quotient exists for, on a VALID instance: routed wires that satisfy a
    copy-constraint permutation and an arithmetic gate w3 = c0 w0 w1 + c1 w2 on four further columns.
    Then every vanishing term is zero on the subgroup, the device's quotient chunks t_k are a true
    polynomial quotient, and at a random point zeta outside the domain
        sum_j alpha^j term_j(zeta) = (zeta^n - 1) * sum_k zeta^(n k) t_k(zeta)
    with the terms recomputed from openings in Python integers (no oracle involved)."""
    from oracle import model as M
    rng = np.random.default_rng(9)
    log_n, num_routed, deg, qdb = 8, 16, 8, 3
    n, K = 1 << log_n, 2
    k_is = V.get_unique_coset_shifts(n, num_routed)
    w = pow(7, (P - 1) >> log_n, P)
    sub = [pow(w, i, P) for i in range(n)]
    ncell = num_routed * n
    perm = rng.permutation(ncell)
    label = np.full(ncell, -1, dtype=np.int64)
    vals = rng.integers(0, P, size=ncell, dtype=np.uint64)
    for c0 in range(ncell):
        c = c0
        while label[c] < 0:
            label[c] = c0
            c = perm[c]
    routed = vals[label].reshape(num_routed, n)
    tj, ti = np.divmod(perm, n)
    sig = np.array([int(k_is[j]) * sub[i] % P for j, i in zip(tj, ti)], dtype=np.uint64).reshape(num_routed, n)
    consts = rng.integers(0, P, size=(2, n), dtype=np.uint64)            # c0, c1 per row
    g = rng.integers(0, P, size=(3, n), dtype=np.uint64)                 # w0, w1, w2
    w3 = np.array([(int(consts[0, i]) * int(g[0, i]) * int(g[1, i]) + int(consts[1, i]) * int(g[2, i])) % P
                   for i in range(n)], dtype=np.uint64)
    wires = np.concatenate([routed, g, w3[None]])
    cs = np.concatenate([consts, sig])
    betas, gammas, alphas = (rng.integers(1, P, size=2, dtype=np.uint64) for _ in range(3))
    wb = V.commit_resident(wires, 3, False, 4, ctx=ctx)
    cb = V.commit_resident(cs, 3, False, 4, ctx=ctx)
    sg = V.Sigmas(sig, k_is, ctx)
    zb = V.commit_zs_partial_products(wb, sg, betas, gammas, deg, 3, 4)
    # the gate's constraint on the quotient domain, alpha-reduced (one constraint: alpha^0)
    q = n << qdb
    wl, cl = wb.get_lde_rows(0, 1, q), cb.get_lde_rows(0, 1, q)          # natural order (rate_bits == qdb)
    gate = np.array([(int(r[19]) - (int(c[0]) * int(r[16]) * int(r[17]) + int(c[1]) * int(r[18]))) % P
                     for r, c in zip(wl, cl)], dtype=np.uint64)
    qb = V.commit_quotient_polys(cb, 2, wb, zb, k_is, deg, qdb, betas, gammas, alphas, 3, 4,
                                 np.stack([gate, gate]))
    t = qb.download().polynomials
    wc, cc, zc = wb.download().polynomials, cb.download().polynomials, zb.download().polynomials
    ints = lambda a: [int(x) for x in a]
    zeta = int(rng.integers(2, P, dtype=np.uint64))
    ev = lambda coeffs, x: M.evaluate(ints(coeffs), x)
    gz = w * zeta % P
    wz, cz, zz = [ev(c, zeta) for c in wc], [ev(c, zeta) for c in cc], [ev(c, zeta) for c in zc]
    zg = [ev(zc[c], gz) for c in range(2)]
    zh = (pow(zeta, n, P) - 1) % P
    l0 = zh * pow(n * (zeta - 1) % P, P - 2, P) % P
    terms = [l0 * (zz[c] - 1) % P for c in range(2)]
    for c in range(2):
        accs = [zz[c]] + [zz[2 + c * (K - 1) + k] for k in range(K - 1)] + [zg[c]]
        for k in range(K):
            num = den = 1
            for j in range(k * deg, (k + 1) * deg):
                num = num * (wz[j] + int(betas[c]) * int(k_is[j]) * zeta + int(gammas[c])) % P
                den = den * (wz[j] + int(betas[c]) * cz[2 + j] + int(gammas[c])) % P
            terms.append((accs[k] * num - accs[k + 1] * den) % P)
    terms.append((wz[19] - (cz[0] * wz[16] * wz[17] + cz[1] * wz[18])) % P)   # the gate constraint, last
    for c in range(2):
        lhs = sum(tm * pow(int(alphas[c]), j, P) for j, tm in enumerate(terms)) % P
        rhs = zh * sum(pow(zeta, n * k, P) * ev(t[(c << qdb) + k], zeta) for k in range(1 << qdb)) % P
        assert lhs == rhs
    # deg(vanishing) <= 9 (n - 1), so the true quotient has degree <= 8 n - 9: the top 8 coefficients of
    # the last chunk vanish only because the instance is valid (a random instance fills them)
    for c in range(2):
        assert not t[(c << qdb) + (1 << qdb) - 1][n - 8:].any()
    for b in (wb, cb, zb, qb):
        b.close()
    sg.close()


@pytest.mark.parametrize("log_n,ncols,rate_bits,salted", [(8, 9, 3, False), (10, 135, 3, False), (5, 6, 2, True), (0, 3, 3, False)])
def test_batch_get_lde_rows(V, ctx, oracle, log_n, ncols, rate_bits, salted):
    """vpbs_batch_get_lde_rows == [get_lde_values(first + i * step) for i] (what the quotient reads)."""
    rng = np.random.default_rng(log_n * 7 + ncols)
    cols = rand_u64(rng, (ncols, 1 << log_n))
    m = (1 << log_n) << rate_bits
    salt = rand_u64(rng, (4, m)) if salted else None
    rb = V.commit_resident(cols, rate_bits, salted, min(2, log_n + rate_bits), ctx=ctx, salt=salt)
    ref = oracle.commit(cols, rate_bits, min(2, log_n + rate_bits), False, salt)
    lg = log_n + rate_bits
    for first, step, count in [(0, 1, m), (1, 3, (m - 1 + 2) // 3), (m - 1, 1, 1), (0, 1 << rate_bits, 1 << log_n), (2, 5, 0)]:
        first = min(first, m - 1)
        if count and first + (count - 1) * step >= m:
            count = (m - 1 - first) // step + 1
        got = rb.get_lde_rows(first, step, count)
        want = np.array([ref["leaves"][V.reverse_bits(first + i * step, lg)][:ncols] for i in range(count)],
                        dtype=np.uint64).reshape(count, ncols)
        assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        rb.get_lde_rows(0, 1, m + 1)
    with pytest.raises(ValueError):
        rb.get_lde_rows(m, 1, 1)
    rb.close()


def test_batch_outliving_its_context_is_inert(V):
    """vpbs_ctx_destroy releases the device buffers of live batches and orphans the handles: reads
    fail with VPBS_ERR_STATE, destroy only frees the handle (no use-after-free)."""
    c = V.Context(0)
    rb = V.commit_resident(V.synthetic_columns(5, 64), 3, False, 2, ctx=c)
    sg = V.Sigmas(V.synthetic_columns(4, 64), V.get_unique_coset_shifts(64, 4), c)
    c.close()
    out = np.empty((1, 5), np.uint64)
    idx = np.zeros(1, np.uint64)
    rc = rb.ctx.lib.vpbs_batch_get_leaves(rb.handle, idx.ctypes.data_as(V._lib.u64p), 1,
                                          out.ctypes.data_as(V._lib.u64p))
    assert rc == V._lib.VPBS_ERR_STATE
    rb.close()
    sg.close()


def test_commit_resident_device_columns(V, ctx, oracle):
    """vpbs_batch_commit_dev: resident batch from columns already in HBM."""
    import torch
    cols = V.synthetic_columns(20, 1 << 9, seed=77, canonical=False)
    d = torch.from_numpy(cols.view(np.int64)).cuda()
    torch.cuda.synchronize()
    rb = V.commit_resident_device(ctx, d.data_ptr(), 20, 9, 3, 4)
    ref = oracle.commit(cols, 3, 4)
    assert np.array_equal(rb.merkle_tree.cap, ref["cap"])
    assert np.array_equal(rb.download().merkle_tree.leaves, ref["leaves"])
    rb.close()


@pytest.mark.parametrize("log_n,rate_bits,arities,cap_height", [
    (16, 3, [4, 4, 4], 4),    # the N=1024 step's commit phase: 2^19 -> 2^15 -> 2^11 -> 2^7 values
    (13, 3, [4, 4], 4), (6, 1, [2, 1, 3], 1), (4, 0, [4], 0), (3, 2, [], 2)])
def test_fri_commit_phase_resident_chain(V, ctx, oracle, log_n, rate_bits, arities, cap_height):
    """vpbs_fri_*: the whole commit phase with polynomial and trees resident in HBM equals the oracle's
    layer-by-layer chain: caps, final polynomial, and queried rows + Merkle paths of every layer."""
    rng = np.random.default_rng(log_n * 13 + rate_bits)
    coeffs0 = rand_u64(rng, (1 << log_n, 2))
    fri = V.FriCommitPhase(coeffs0, rate_bits, ctx)
    m = (1 << log_n) << rate_bits
    coeffs = np.zeros((m, 2), np.uint64)
    coeffs[: 1 << log_n] = coeffs0 % np.uint64(P)
    shift = 7
    values = np.stack([oracle.coset_fft(coeffs[:, 0].copy(), shift), oracle.coset_fft(coeffs[:, 1].copy(), shift)], 1)
    lg = log_n + rate_bits
    for k, a in enumerate(arities):
        h = min(cap_height, lg - a)
        ref = oracle.fri_layer_commit(values, a, h)
        cap = fri.commit_layer(a, h)
        assert np.array_equal(cap, ref["cap"]), "cap of layer %d" % k
        nleaves = 1 << (lg - a)
        idx = rng.integers(0, nleaves, size=min(28, nleaves), dtype=np.uint64)
        rows, sib = fri.query(k, idx)
        for q, i in enumerate(idx):
            assert np.array_equal(rows[q], ref["leaves"][int(i)])
            assert np.array_equal(sib[q], oracle.merkle_prove(ref["digests"], nleaves, h, int(i)))
        beta = rand_u64(rng, 2)
        shift = oracle.gl_pow(shift, 1 << a)
        coeffs, values = oracle.fri_fold(coeffs, a, beta, shift)
        fri.fold(beta)
        lg -= a
    assert np.array_equal(fri.final_poly(), coeffs[: coeffs.shape[0] >> rate_bits])
    with pytest.raises(V.VpbsError):
        fri.fold([1, 2])  # nothing committed to fold
    fri.close()


@pytest.mark.parametrize("log_n,widths,sizes", [
    (0, (3,), ((0, 0), (0, 2))), (3, (5, 2), None), (8, (9, 4, 3), None), (12, (20, 16), None),
    (13, (7,), None), (16, (135, 20, 16), None)])
def test_prove_openings_final_poly_from_resident_batches(V, ctx, oracle, log_n, widths, sizes):
    """[P2] fri/oracle.rs prove_openings from its first line to the committed first FRI layer, with the
    batches resident (vpbs_fri_begin_openings): alpha-combination of the coefficient polynomials,
    division by (X - z), shift / accumulate over two FRI batches (all polynomials at zeta; the
    first oracle's leading polynomials at g zeta, as plonky2 opens Z), then LDE + first layer cap.
    The last case has the N=1024 step's shapes (135 + 20 + 16 polynomials of 2^16 coefficients)."""
    rng = np.random.default_rng(log_n * 31 + len(widths))
    n = 1 << log_n
    cols = [rand_u64(rng, (w, n)) for w in widths]
    obs = [V.commit_resident(c, 3, False, min(4, log_n + 3), True, ctx=ctx) for c in cols]
    all_refs = [(o, j) for o, w in enumerate(widths) for j in range(w)]
    next_refs = [(len(widths) - 1, j) for j in range(min(2, widths[-1]))]
    if sizes is not None:
        all_refs, next_refs = [sizes[0]], [sizes[1]]
    batches = [all_refs, next_refs]
    pts = rand_u64(rng, (2, 2), edge_frac=0)
    alpha = rand_u64(rng, 2, edge_frac=0)
    fri = V.FriCommitPhase.from_openings(obs, batches, pts, alpha, 3)
    got = fri.final_poly()  # before any fold: final_poly itself
    want = oracle.fri_final_poly([np.stack([cols[o][j] for (o, j) in b]) for b in batches], pts, alpha)
    assert np.array_equal(got, want)
    if log_n + 3 >= 4:  # the chain goes on exactly as from host coefficients
        ref = V.FriCommitPhase(want, 3, ctx)
        h = min(4, log_n + 3 - 4)
        assert np.array_equal(fri.commit_layer(4, h), ref.commit_layer(4, h))
        ref.close()
    fri.close()
    for o in obs:
        o.close()


def test_prove_openings_rejects_bad_instances(V, ctx):
    rng = np.random.default_rng(3)
    a = V.commit_resident(rand_u64(rng, (3, 16)), 1, False, 0, True, ctx=ctx)
    b = V.commit_resident(rand_u64(rng, (2, 32)), 1, False, 0, True, ctx=ctx)
    pt, al = np.array([[1, 2]], np.uint64), np.array([3, 4], np.uint64)
    with pytest.raises(ValueError):  # FriPolynomialInfo out of range
        V.FriCommitPhase.from_openings([a], [[(0, 3)]], pt, al, 1)
    with pytest.raises(ValueError):  # oracles of different degree
        V.FriCommitPhase.from_openings([a, b], [[(0, 0), (1, 0)]], pt, al, 1)
    with pytest.raises(ValueError):  # empty batch
        V.FriCommitPhase.from_openings([a], [[]], pt, al, 1)
    a.close()
    b.close()


@pytest.mark.parametrize("log_n,ncols,rate_bits,cap_height,coeffs", [
    (12, 64, 3, 4, False), (12, 70, 3, 4, False), (13, 135, 3, 4, False), (12, 100, 2, 0, True),
    (12, 65, 1, 13, False), (14, 128, 3, 4, False)])
def test_resident_commit_of_wide_host_batches(V, ctx, oracle, log_n, ncols, rate_bits, cap_height, coeffs):
    """vpbs_batch_commit on batches wide enough to travel in column chunks (>= 64 columns, >= 2^12
    rows: chunk k's IFFT and its columns of every LDE block run while chunk k+1 is on the bus), incl.
    a ragged last chunk, a last chunk of one column, widths that are not multiples of 8, the all-cap
    tree and from_coeffs.  Everything equals the oracle's commit."""
    rng = np.random.default_rng(log_n * 1000 + ncols)
    cols = rand_u64(rng, (ncols, 1 << log_n))
    rb = V.commit_resident(cols, rate_bits, False, cap_height, coeffs, ctx=ctx)
    ref = oracle.commit(cols, rate_bits, cap_height, coeffs)
    assert np.array_equal(rb.merkle_tree.cap, ref["cap"])
    eager = rb.download()
    assert np.array_equal(eager.merkle_tree.leaves, ref["leaves"])
    assert np.array_equal(eager.merkle_tree.digests, ref["digests"])
    if not coeffs:
        assert np.array_equal(eager.polynomials, ref["coeffs"])
    rb.close()


def test_prove_openings_random_instances(V, ctx, oracle):
    """Random FRI instances (1-3 batches, polynomials drawn with repetition from 1-3 oracles of
    different widths, sizes 2^0..2^11, non-canonical points / alpha) against the oracle."""
    rnd = random.Random(77)
    rng = np.random.default_rng(77)
    for _ in range(24):
        log_n = rnd.randint(0, 11)
        n = 1 << log_n
        widths = [rnd.choice([1, 2, 5, 9, 20]) for _ in range(rnd.randint(1, 3))]
        cols = [rand_u64(rng, (w, n)) for w in widths]
        obs = [V.commit_resident(c, 1, False, 0, True, ctx=ctx) for c in cols]
        batches = []
        for _b in range(rnd.randint(1, 3)):
            k = rnd.randint(1, 12)
            batches.append([(o, rnd.randrange(widths[o])) for o in (rnd.randrange(len(widths)) for _ in range(k))])
        pts = rand_u64(rng, (len(batches), 2))
        alpha = rand_u64(rng, 2)
        fri = V.FriCommitPhase.from_openings(obs, batches, pts, alpha, 1)
        want = oracle.fri_final_poly([np.stack([cols[o][j] for (o, j) in b]) for b in batches], pts, alpha)
        assert np.array_equal(fri.final_poly(), want), (log_n, widths, batches)
        fri.close()
        for o in obs:
            o.close()


def test_fft_of_2p24_points_three_256_point_passes(V, ctx, oracle):
    """log_n = 24 = 8 + 8 + 8: the only size whose strided radix-16 passes work on more than one
    block per column (TMA box coordinate `block` > 0) and apply the four-step twiddle at their
    store; forward, inverse (round trip) and coset transforms against the oracle."""
    rng = np.random.default_rng(24)
    v = rand_u64(rng, 1 << 24, 0.01)
    want = oracle.fft(v.copy())
    got = V.fft(v.copy(), ctx)
    assert np.array_equal(got, want)
    assert np.array_equal(V.ifft(got.copy(), ctx), v % np.uint64(P))
    assert np.array_equal(V.coset_fft(v.copy(), 7, ctx), oracle.coset_fft(v.copy(), 7))
