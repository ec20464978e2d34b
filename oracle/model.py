"""model.py — independent big-integer Python model of plonky2 0.2.0's commitment path.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  It exists as a second opinion on oracle.c: it is
written from the *mathematical definitions* (direct O(n^2) polynomial evaluation instead of an FFT,
Python `%` instead of reduce128, recursion by definition for the Merkle tree) so that a shared bug
between the two is unlikely.  Pure-Python loops: use only for small cases.

Upstream items restated ([P2] = plonky2 0.2.0 family, absent from /root/reference; reached from
/root/reference/src/vtfhe/ivc_based_vpbs.rs:275,302,333,364):
  [P2] plonky2_field/src/goldilocks_field.rs   p, generator 7, POWER_OF_TWO_GENERATOR
  [P2] plonky2_field/src/fft.rs, polynomial/mod.rs   fft / ifft / lde / coset_fft
  [P2] plonky2/src/hash/poseidon{,_goldilocks}.rs    permutation
  [P2] plonky2/src/hash/hashing.rs, plonk/config.rs  hash_n_to_m_no_pad / hash_or_noop / compress
  [P2] plonky2/src/hash/merkle_tree.rs               MerkleTree::new / prove
  [P2] plonky2/src/fri/oracle.rs                     PolynomialBatch::from_values / from_coeffs
"""
from __future__ import annotations

P = 2**64 - 2**32 + 1
GENERATOR = 7
POWER_OF_TWO_GENERATOR = pow(GENERATOR, (P - 1) >> 32, P)  # == 1753635133440165772
COSET_SHIFT = GENERATOR
SALT_SIZE = 4


def primitive_root_of_unity(n_log: int) -> int:
    assert n_log <= 32
    return pow(POWER_OF_TWO_GENERATOR, 1 << (32 - n_log), P)


def bitrev(x: int, bits: int) -> int:
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


# --------------------------------------------------------------------------- polynomials
def evaluate(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def fft(coeffs):
    """[P2] fft.rs fft: out[i] = poly(w_n^i), natural order."""
    n = len(coeffs)
    w = primitive_root_of_unity(n.bit_length() - 1)
    return [evaluate(coeffs, pow(w, i, P)) for i in range(n)]


def ifft(values):
    """[P2] fft.rs ifft: c_j = n^-1 sum_i v_i w^(-ij)."""
    n = len(values)
    w_inv = pow(primitive_root_of_unity(n.bit_length() - 1), P - 2, P)
    n_inv = pow(n, P - 2, P)
    return [evaluate(values, pow(w_inv, j, P)) * n_inv % P for j in range(n)]


def coset_fft(coeffs, shift):
    """[P2] polynomial/mod.rs coset_fft: out[i] = poly(shift * w_n^i)."""
    n = len(coeffs)
    w = primitive_root_of_unity(n.bit_length() - 1)
    return [evaluate(coeffs, shift * pow(w, i, P) % P) for i in range(n)]


def lde(coeffs, rate_bits):
    """[P2] lde(rate_bits) + coset_fft(coset_shift): evaluations on 7*<w_m>, natural order."""
    padded = list(coeffs) + [0] * (len(coeffs) * ((1 << rate_bits) - 1))
    return coset_fft(padded, COSET_SHIFT)


# --------------------------------------------------------------------------- Poseidon
MDS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
MDS_DIAG = [8] + [0] * 11
HALF_N_FULL_ROUNDS = 4
N_PARTIAL_ROUNDS = 22
WIDTH = 12
RATE = 8


def _chacha8_stream(key):
    def rotl(x, r):
        return ((x << r) | (x >> (32 - r))) & 0xFFFFFFFF

    def qr(x, a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 7)

    counter = 0
    while True:
        s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key) + [
            counter & 0xFFFFFFFF, counter >> 32, 0, 0]
        x = list(s)
        for _ in range(4):
            qr(x, 0, 4, 8, 12); qr(x, 1, 5, 9, 13); qr(x, 2, 6, 10, 14); qr(x, 3, 7, 11, 15)
            qr(x, 0, 5, 10, 15); qr(x, 1, 6, 11, 12); qr(x, 2, 7, 8, 13); qr(x, 3, 4, 9, 14)
        for a, b in zip(x, s):
            yield (a + b) & 0xFFFFFFFF
        counter += 1


def round_constants():
    """ALL_ROUND_CONSTANTS regenerated as ChaCha8Rng::seed_from_u64(0) draws."""
    state, key = 0, []
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) % 2**64
        xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        key.append(((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF)
    words = _chacha8_stream(key)
    out = []
    while len(out) < 360:
        v = next(words) | (next(words) << 32)
        wide = v * P
        if wide % 2**64 <= P - 1:
            out.append(wide >> 64)
    return out


RC = round_constants()


def poseidon(state):
    """[P2] Poseidon::poseidon, naive rounds: constants, S-box (x^7; lane 0 only in partial
    rounds), MDS (out[r] = sum_i CIRC[i]*s[(i+r)%12] + DIAG[r]*s[r])."""
    s = [x % P for x in state]
    assert len(s) == WIDTH
    for r in range(2 * HALF_N_FULL_ROUNDS + N_PARTIAL_ROUNDS):
        s = [(x + RC[12 * r + i]) % P for i, x in enumerate(s)]
        if r < HALF_N_FULL_ROUNDS or r >= HALF_N_FULL_ROUNDS + N_PARTIAL_ROUNDS:
            s = [pow(x, 7, P) for x in s]
        else:
            s[0] = pow(s[0], 7, P)
        s = [(sum(MDS_CIRC[i] * s[(i + r2) % 12] for i in range(12)) + MDS_DIAG[r2] * s[r2]) % P
             for r2 in range(12)]
    return s


def hash_no_pad(inputs):
    state = [0] * WIDTH
    for off in range(0, len(inputs), RATE):
        chunk = inputs[off:off + RATE]
        state[:len(chunk)] = [x % P for x in chunk]   # overwrite mode
        state = poseidon(state)
    return state[:4]


def hash_or_noop(inputs):
    if len(inputs) <= 4:
        return [x % P for x in inputs] + [0] * (4 - len(inputs))
    return hash_no_pad(inputs)


def two_to_one(left, right):
    return poseidon(list(left) + list(right) + [0] * 4)[:4]


# --------------------------------------------------------------------------- Merkle tree
def _fill_subtree(leaves):
    """Returns (digests_buf as list of hashes, root) following [P2] fill_subtree's layout:
    left recursive output || left child digest || right child digest || right recursive output."""
    if len(leaves) == 1:
        return [], hash_or_noop(leaves[0])
    half = len(leaves) // 2
    lbuf, ld = _fill_subtree(leaves[:half])
    rbuf, rd = _fill_subtree(leaves[half:])
    return lbuf + [ld, rd] + rbuf, two_to_one(ld, rd)


def merkle_new(leaves, cap_height):
    """[P2] MerkleTree::new -> (digests, cap), each a list of 4-element hashes."""
    n = len(leaves)
    assert n & (n - 1) == 0 and cap_height <= n.bit_length() - 1
    sub = n >> cap_height
    digests, cap = [], []
    for s in range(1 << cap_height):
        buf, root = _fill_subtree(leaves[s * sub:(s + 1) * sub])
        digests += buf
        cap.append(root)
    assert len(digests) == 2 * (n - (1 << cap_height))
    return digests, cap


def merkle_prove(digests, nleaves, cap_height, leaf_index):
    num_layers = nleaves.bit_length() - 1 - cap_height
    tree_len = len(digests) >> cap_height
    tree = digests[tree_len * (leaf_index >> num_layers):][:tree_len]
    pair_index = leaf_index & ((1 << num_layers) - 1)
    sibs = []
    for i in range(num_layers):
        parity = pair_index & 1
        pair_index >>= 1
        sibs.append(tree[2 * ((pair_index << (i + 1)) + (1 << i) - 1) + (1 - parity)])
    return sibs


def merkle_verify(leaf, leaf_index, siblings, cap):
    cur = hash_or_noop(leaf)
    idx = leaf_index
    for sib in siblings:
        cur = two_to_one(sib, cur) if idx & 1 else two_to_one(cur, sib)
        idx >>= 1
    return cur == cap[idx]


# --------------------------------------------------------------------------- PolynomialBatch
def commit(cols, rate_bits, cap_height, inputs_are_coeffs=False, salt_cols=None):
    """[P2] PolynomialBatch::from_values / from_coeffs.
    Returns dict(coeffs, lde (natural order, per column), leaves (bit-reversed rows), digests, cap)."""
    n = len(cols[0])
    log_m = n.bit_length() - 1 + rate_bits
    coeffs = [[x % P for x in c] if inputs_are_coeffs else ifft(c) for c in cols]
    lde_cols = [lde(c, rate_bits) for c in coeffs]
    all_cols = lde_cols + ([[x % P for x in s] for s in salt_cols] if salt_cols else [])
    leaves = [[col[bitrev(k, log_m)] for col in all_cols] for k in range(n << rate_bits)]
    digests, cap = merkle_new(leaves, cap_height)
    return dict(coeffs=coeffs, lde=lde_cols, leaves=leaves, digests=digests, cap=cap)


# --------------------------------------------------------------------------- quadratic extension
EXT_W = 7  # [P2] extension/quadratic.rs: GoldilocksField is Extendable<2> with W = 7


def ext_mul(a, b):
    return ((a[0] * b[0] + EXT_W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def eval_ext2(coeffs, x):
    """[P2] PolynomialCoeffs::to_extension().eval(x): sum_j c_j x^j by explicit powers."""
    acc, pw = (0, 0), (1, 0)
    for c in coeffs:
        acc = ((acc[0] + c % P * pw[0]) % P, (acc[1] + c % P * pw[1]) % P)
        pw = ext_mul(pw, x)
    return acc


# --------------------------------------------------------------------------- FRI commit phase
def fri_layer_commit(values, arity_bits, cap_height):
    """[P2] fri/prover.rs fri_committed_trees, tree of one layer.  values: list of (re, im)."""
    lg = len(values).bit_length() - 1
    rev = [values[bitrev(k, lg)] for k in range(len(values))]
    arity = 1 << arity_bits
    leaves = [[x % P for pair in rev[j * arity:(j + 1) * arity] for x in pair]
              for j in range(len(values) >> arity_bits)]
    digests, cap = merkle_new(leaves, cap_height)
    return leaves, digests, cap


def fri_fold(coeffs, arity_bits, beta, shift_next):
    """coeffs' = sum_i chunk[i] beta^i per chunk of `arity`; values' = coeffs'(shift_next * w^k)."""
    arity = 1 << arity_bits
    out = []
    for j in range(len(coeffs) >> arity_bits):
        acc, pw = (0, 0), (1, 0)
        for i in range(arity):
            t = ext_mul((coeffs[j * arity + i][0] % P, coeffs[j * arity + i][1] % P), pw)
            acc = ((acc[0] + t[0]) % P, (acc[1] + t[1]) % P)
            pw = ext_mul(pw, beta)
        out.append(acc)
    re = coset_fft([c[0] for c in out], shift_next)
    im = coset_fft([c[1] for c in out], shift_next)
    return out, list(zip(re, im))


def fri_final_poly(batches, points, alpha):
    """[P2] fri/oracle.rs prove_openings up to final_poly, from the definitions (no Horner scan):
    batches: per FRI batch a list of coefficient lists; points / alpha: extension elements.
    F_b = sum_j alpha^j f_bj;  Q_b = (F_b(X) - F_b(z)) / (X - z) has coefficients
    q_i = sum_{j > i} F_b[j] z^(j - i - 1)  (padded with q_{n-1} = 0);
    final = final * alpha^len(batch) + Q_b."""
    n = len(batches[0][0])
    final = [(0, 0)] * n
    for polys, z in zip(batches, points):
        comp, pw = [(0, 0)] * n, (1, 0)
        for f in polys:
            comp = [((c[0] + x % P * pw[0]) % P, (c[1] + x % P * pw[1]) % P) for c, x in zip(comp, f)]
            pw = ext_mul(pw, alpha)
        zp = [(1, 0)]
        for _ in range(n):
            zp.append(ext_mul(zp[-1], z))
        quot = []
        for i in range(n):
            acc = (0, 0)
            for j in range(i + 1, n):
                t = ext_mul(comp[j], zp[j - i - 1])
                acc = ((acc[0] + t[0]) % P, (acc[1] + t[1]) % P)
            quot.append(acc)
        shifted = [ext_mul(f, pw) for f in final]  # pw = alpha^len(polys)
        final = [((a[0] + b[0]) % P, (a[1] + b[1]) % P) for a, b in zip(shifted, quot)]
    return final


# --------------------------------------------------------------------------- permutation argument
def zs_partial_products(wires, sigmas, k_is, max_degree, beta, gamma):
    """[P2] plonk/prover.rs wires_permutation_partial_products_and_zs for one (beta, gamma), from the
    definitions: wires[j][i], sigmas[j][i] over the subgroup x_i = w_n^i.  Returns the columns in
    commit order [Z, pp_0, .., pp_{K-2}] (K = ceil(num_routed / max_degree) chunks per row)."""
    num_routed, n = len(wires), len(wires[0])
    w = primitive_root_of_unity(n.bit_length() - 1)
    K = -(-num_routed // max_degree)
    cols = [[0] * n for _ in range(K)]
    z = 1
    for i in range(n):
        x = pow(w, i, P)
        cols[0][i] = z
        acc = z
        for k in range(K):
            for j in range(k * max_degree, min((k + 1) * max_degree, num_routed)):
                num = (wires[j][i] + beta * k_is[j] * x + gamma) % P
                den = (wires[j][i] + beta * sigmas[j][i] + gamma) % P
                acc = acc * num * pow(den, P - 2, P) % P
            if k + 1 < K:
                cols[1 + k][i] = acc
        z = acc
    return cols


def quotient_polys(wire_coeffs, sigma_coeffs, zs_pp_coeffs, k_is, max_degree, qdb, betas, gammas, alphas,
                   gate_terms=None):
    """[P2] plonk/prover.rs compute_quotient_polys restricted to the gate-independent vanishing terms
    (plonk/vanishing_poly.rs eval_vanishing_poly_base_batch: the Z(1) = 1 terms, then the
    partial-product checks of every challenge, reduced with powers of each alpha), straight from the
    definitions: every polynomial is evaluated by Horner at x = 7 w_q^i, the vanishing value divided
    by x^n - 1, the quotient interpolated on the coset by the defining sum and cut into chunks of n.
    gate_terms[c][i] (optional) = alpha-reduced gate constraints at point i; they follow the nc + nc K
    permutation terms in upstream's order, i.e. enter times alpha_c^(nc + nc K).
    Returns nc * 2^qdb coefficient lists of length n."""
    num_routed, n = len(wire_coeffs), len(wire_coeffs[0])
    nc = len(betas)
    K = -(-num_routed // max_degree)
    q = n << qdb
    wq = primitive_root_of_unity(q.bit_length() - 1)
    g = primitive_root_of_unity(n.bit_length() - 1)
    n_inv_mod = lambda v: pow(v, P - 2, P)
    vals = [[0] * q for _ in range(nc)]
    for i in range(q):
        x = 7 * pow(wq, i, P) % P
        gx = g * x % P
        wv = [evaluate(c, x) for c in wire_coeffs]
        sv = [evaluate(c, x) for c in sigma_coeffs]
        zv = [evaluate(c, x) for c in zs_pp_coeffs]
        zg = [evaluate(zs_pp_coeffs[c], gx) for c in range(nc)]
        zh = (pow(x, n, P) - 1) % P
        l0 = zh * n_inv_mod(n * (x - 1) % P) % P
        terms = [l0 * (zv[c] - 1) % P for c in range(nc)]
        for c in range(nc):
            accs = [zv[c]] + [zv[nc + c * (K - 1) + t] for t in range(K - 1)] + [zg[c]]
            for t in range(K):
                num = den = 1
                for j in range(t * max_degree, min((t + 1) * max_degree, num_routed)):
                    num = num * (wv[j] + betas[c] * k_is[j] * x + gammas[c]) % P
                    den = den * (wv[j] + betas[c] * sv[j] + gammas[c]) % P
                terms.append((accs[t] * num - accs[t + 1] * den) % P)
        for c in range(nc):
            res = sum(t * pow(alphas[c], j, P) for j, t in enumerate(terms)) % P
            if gate_terms is not None:
                res = (res + pow(alphas[c], len(terms), P) * gate_terms[c][i]) % P
            vals[c][i] = res * n_inv_mod(zh) % P
    out = []
    w_inv, q_inv, s_inv = n_inv_mod(wq), n_inv_mod(q), n_inv_mod(7)
    for c in range(nc):
        coeffs = [sum(v * pow(w_inv, i * j, P) for i, v in enumerate(vals[c])) % P * q_inv % P * pow(s_inv, j, P) % P
                  for j in range(q)]
        out += [coeffs[t * n:(t + 1) * n] for t in range(1 << qdb)]
    return out


def gate_program_eval(code, imms, num_constraints, wire_coeffs, cs_coeffs, qdb, pih, alphas):
    """The gate-constraint program (format: include/vpbs_commit.h) interpreted from its definition: every
    wire / constant operand is the polynomial evaluated by Horner at x = 7 w_q^i.  Returns
    [[alpha_c-reduced gate constraints at point i] for c]."""
    n = len(wire_coeffs[0])
    q = n << qdb
    wq = primitive_root_of_unity(q.bit_length() - 1)
    out = [[0] * q for _ in alphas]
    for i in range(q):
        x = 7 * pow(wq, i, P) % P
        regs, total, gacc = {}, [0] * len(alphas), [0] * len(alphas)

        def val(kind, idx):
            return (regs[idx] if kind == 0 else evaluate(wire_coeffs[idx], x) if kind == 1 else
                    evaluate(cs_coeffs[idx], x) if kind == 2 else imms[idx] % P if kind == 3 else pih[idx] % P)
        for ins in code:
            op, dst = ins & 0xff, (ins >> 8) & 0xff
            a = val((ins >> 16) & 0xf, (ins >> 24) & 0xffff)
            if op <= 2 or op == 5:
                b = val((ins >> 20) & 0xf, (ins >> 40) & 0xffff)
                regs[dst] = ((a + b) % P if op == 0 else (a - b) % P if op == 1 else a * b % P if op == 2
                             else (regs[dst] + a * b) % P)
            elif op == 3:
                j = (ins >> 40) & 0xffff
                assert j < num_constraints
                gacc = [(g + a * pow(al, j, P)) % P for g, al in zip(gacc, alphas)]
            else:
                total = [(t + g * a) % P for t, g in zip(total, gacc)]
                gacc = [0] * len(alphas)
        for c in range(len(alphas)):
            out[c][i] = total[c]
    return out
