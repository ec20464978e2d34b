"""CPU checks of the generated Poseidon tables the CUDA kernel is driven by (csrc/poseidon_*.inc).

The kernel does not run the textbook round function: the MDS layer is a halved split convolution on
the FP64 pipe whose accumulators start at table values (poseidon_rcs.inc), and the 22 partial rounds
run as 11 pair steps through M^2 with their own constants (poseidon_rcp.inc).  This file restates
that algorithm (csrc/poseidon.cuh: mds_begin / mds_absorb / mds_finish / partial_pair) in exact
integer arithmetic FROM THE TABLES, asserts that every value a double would hold is an integer below
2^53 (so the FP64 arithmetic is exact), and compares the result with the oracle's permutation.  It
needs no GPU; the GPU parity tests cover the real kernel.
"""
import importlib.util
import json
import os
import random
import re

import numpy as np
import pytest

from oracle import binding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "verifiable-fhe-paper_b200", "csrc")
P = 2**64 - 2**32 + 1
M32 = 0xFFFFFFFF
BIAS = 2**52
CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]


def _load_doubles(name):
    txt = open(os.path.join(CSRC, name)).read()
    txt = "\n".join(l for l in txt.splitlines() if not l.strip().startswith("//"))
    vals = [int(v) for v in re.findall(r"(-?\d+)\.0", txt)]
    return vals


def _load_u64(name):
    txt = open(os.path.join(CSRC, name)).read()
    return [int(v, 16) for v in re.findall(r"0x([0-9a-fA-F]{16})ULL", txt)]


RC = _load_u64("poseidon_rc.inc")
RCS = _load_doubles("poseidon_rcs.inc")
RCP = _load_doubles("poseidon_rcp.inc")
RCD = _load_doubles("poseidon_rcd.inc")


def _exact(v):
    """A value held in a double accumulator: must be an integer of magnitude < 2^53."""
    assert isinstance(v, int) and abs(v) < 2**53, v
    return v


def _cplus(j):
    assert (CIRC[j] + CIRC[j + 6]) % 2 == 0
    return (CIRC[j] + CIRC[j + 6]) // 2


def _cminus(j):
    return (CIRC[j] - CIRC[j + 6]) // 2


def _cc(d):
    return sum(CIRC[a] * CIRC[(d - a) % 12] for a in range(12))


def _ccplus(j):
    assert (_cc(j) + _cc(j + 6)) % 2 == 0
    return (_cc(j) + _cc(j + 6)) // 2


def _ccminus(j):
    return (_cc(j) - _cc(j + 6)) // 2


def _combine(dlo, dhi):
    """combine_biased: two biased accumulators -> one state word (any representative < 2^64)."""
    L, H = dlo - BIAS, dhi - BIAS
    assert 0 <= L < 2**52 and 0 <= H < 2**52
    v = (L + (H << 32)) % P
    return v


def _split_conv(s, init, cp, cm):
    """Accumulators after absorbing all lanes: returns zpl, zml, zph, zmh (lists of 6)."""
    zpl = [init[4 * r + 0] for r in range(6)]
    zml = [init[4 * r + 1] for r in range(6)]
    zph = [init[4 * r + 2] for r in range(6)]
    zmh = [init[4 * r + 3] for r in range(6)]
    for t in range(6):
        al, ah = s[t] & M32, s[t] >> 32
        bl, bh = s[t + 6] & M32, s[t + 6] >> 32
        pl, ph, ml, mh = al + bl, ah + bh, al - bl, ah - bh
        for r in range(6):
            j = (t - r) % 6
            c_p = cp(j)
            c_m = cm(j) if j + r < 6 else -cm(j)
            zpl[r] = _exact(zpl[r] + pl * c_p)
            zph[r] = _exact(zph[r] + ph * c_p)
            zml[r] = _exact(zml[r] + ml * c_m)
            zmh[r] = _exact(zmh[r] + mh * c_m)
    return zpl, zml, zph, zmh


def _sbox(x):
    return pow(x, 7, P)


def _sbox_words(x, lift):
    """sbox7_words: x^7 = x^3 * x^4 with the last 128-bit product left unreduced, returned as the
    two "halves" (L, H) = (p0 - u - h1, s1 + u) the kernel's mds_absorb_words forms.  `lift` picks
    non-canonical representatives of x^3 / x^4 where they fit in 64 bits (the kernel's are lazy)."""
    x3, x4 = pow(x, 3, P), pow(x, 4, P)
    if lift and x3 + P < 2**64:
        x3 += P
    if lift and x4 + P < 2**64:
        x4 += P
    prod = x3 * x4
    p0, s1, u, h1 = [(prod >> (32 * k)) & M32 for k in range(4)]
    L, H = p0 - u - h1, s1 + u
    assert (L + (H << 32)) % P == pow(x, 7, P)
    assert -2**33 < L < 2**32 and 0 <= H < 2**33
    return L, H


def _full_round(s, r, lift=False):
    halves = [_sbox_words(x, lift) for x in s]   # (L, H) per lane, as the kernel feeds them
    init = RCS[24 * (r + 1):24 * (r + 2)]
    zpl = [init[4 * q + 0] for q in range(6)]
    zml = [init[4 * q + 1] for q in range(6)]
    zph = [init[4 * q + 2] for q in range(6)]
    zmh = [init[4 * q + 3] for q in range(6)]
    for t in range(6):
        (al, ah), (bl, bh) = halves[t], halves[t + 6]
        pl, ph, ml, mh = al + bl, ah + bh, al - bl, ah - bh
        for q in range(6):
            j = (t - q) % 6
            c_p = _cplus(j)
            c_m = _cminus(j) if j + q < 6 else -_cminus(j)
            zpl[q] = _exact(zpl[q] + pl * c_p)
            zph[q] = _exact(zph[q] + ph * c_p)
            zml[q] = _exact(zml[q] + ml * c_m)
            zmh[q] = _exact(zmh[q] + mh * c_m)
    out = [0] * 12
    for q in range(6):
        s1l, s1h = zpl[q] + zml[q], zph[q] + zmh[q]
        s2l, s2h = zpl[q] - zml[q], zph[q] - zmh[q]
        if q == 0:
            s1l += 8 * halves[0][0]
            s1h += 8 * halves[0][1]
        out[q] = _combine(_exact(s1l), _exact(s1h))
        out[q + 6] = _combine(_exact(s2l), _exact(s2h))
    return out


def _partial_pair(s, pair):
    s = list(s)
    s[0] = _sbox(s[0])
    zpl, zml, zph, zmh = _split_conv(s, RCP[24 * pair:24 * (pair + 1)], _ccplus, _ccminus)
    r1 = 4 + 2 * pair + 1
    xl, xh = RCD[2 * 12 * r1], RCD[2 * 12 * r1 + 1]
    u0l, u0h = s[0] & M32, s[0] >> 32
    for t in range(6):
        al, ah = s[t] & M32, s[t] >> 32
        bl, bh = s[t + 6] & M32, s[t + 6] >> 32
        xl = _exact(xl + (al + bl) * _cplus(t) + (al - bl) * _cminus(t))
        xh = _exact(xh + (ah + bh) * _cplus(t) + (ah - bh) * _cminus(t))
        if t == 0:
            xl = _exact(xl + 8 * u0l)
            xh = _exact(xh + 8 * u0h)
    x1 = _combine(xl, xh)
    u1 = _sbox(x1)
    delta = (u1 - x1) % P
    a_l = 8 * u0l + (delta & M32)
    a_h = 8 * u0h + (delta >> 32)
    G = [17, 20, 34, 18, 39, 13, 13, 28, 2, 16, 41, 15]
    out = [0] * 12
    for q in range(6):
        d1l = zpl[q] + zml[q] + a_l * G[q]
        d1h = zph[q] + zmh[q] + a_h * G[q]
        d2l = zpl[q] - zml[q] + a_l * G[q + 6]
        d2h = zph[q] - zmh[q] + a_h * G[q + 6]
        if q == 0:
            d1l += 8 * (u1 & M32)
            d1h += 8 * (u1 >> 32)
        out[q] = _combine(_exact(d1l), _exact(d1h))
        out[q + 6] = _combine(_exact(d2l), _exact(d2h))
    return out


def permute_from_tables(state, lift=False):
    s = [(x + RC[i]) % P for i, x in enumerate(state)]
    for half in range(2):
        for k in range(4):
            s = _full_round(s, half * 26 + k, lift)
        if half == 0:
            for pair in range(11):
                s = _partial_pair(s, pair)
    return s


def test_tables_are_what_the_generator_writes(tmp_path):
    """csrc/poseidon_*.inc must be the generator's current output (no hand edits, no drift)."""
    spec = importlib.util.spec_from_file_location(
        "gen_consts", os.path.join(ROOT, "verifiable-fhe-paper_b200", "tools", "gen_poseidon_consts.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    rc = gen.all_round_constants()
    assert rc == RC
    assert np.array_equal(orc.round_constants(), np.array(rc, dtype=np.uint64))
    for r in range(31):
        row = rc[12 * r:12 * r + 12] if r < 30 else [0] * 12
        want = []
        for t in range(6):
            want += gen.split_init(row[t], row[t + 6], offset=True)
        assert RCS[24 * r:24 * r + 24] == want
    assert len(RCS) == 31 * 24 and len(RCP) == 11 * 24 and len(RCD) == 31 * 24


def test_split_init_represents_the_constants():
    for r in range(30):
        for t in range(6):
            zpl, zml, zph, zmh = RCS[24 * r + 4 * t:24 * r + 4 * t + 4]
            a = (zpl - BIAS + zml) + ((zph - BIAS + zmh) << 32)
            b = (zpl - BIAS - zml) + ((zph - BIAS - zmh) << 32)
            assert a % P == RC[12 * r + t] and b % P == RC[12 * r + t + 6]
            assert min(zpl - BIAS + zml, zpl - BIAS - zml, zph - BIAS + zmh, zph - BIAS - zmh) >= 0


def test_table_driven_permutation_matches_oracle():
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "poseidon_kat.json")))
    rng = random.Random(1234)
    inputs = [[0] * 12, list(range(12)), [P - 1] * 12]
    inputs += [[rng.randrange(P) for _ in range(12)] for _ in range(12)]
    inputs += [[rng.choice([0, 1, P - 1, P - 2, 2**32 - 1, 2**32, 2**63]) for _ in range(12)] for _ in range(6)]
    for st in inputs:
        want = [int(x) for x in orc.poseidon(np.array(st, dtype=np.uint64))]
        for lift in (False, True):
            got = permute_from_tables(st, lift)
            assert [g % P for g in got] == want
    # and the published vectors themselves
    assert len(kat["vectors"]) >= 3
    for case in kat["vectors"]:
        st = [int(x, 16) for x in case["input"]]
        assert [g % P for g in permute_from_tables(st)] == [int(x, 16) for x in case["output"]]
