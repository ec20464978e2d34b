"""CPU suite, part 2: the C-ABI library loads and exports what include/vpbs_commit.h declares,
fails loudly without a GPU (no CPU fallback), and the host-side logic of the plonky2 mirror."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vpbs_commit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vpbs_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol(V):
    V.build.build()
    lib = V._lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libvpbs_commit.so does not export %s" % s
    assert sorted(V._lib.SIGNATURES) == syms, "binding and header disagree"
    assert lib.vpbs_abi_version() == 3


def test_library_contains_sm100a_code_only(V):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", V._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_host_stage_copy_pool(tmp_path):
    """The copy threads behind the pinned staging ring (csrc/host_stage.h) against plain memcpy."""
    import subprocess
    exe = str(tmp_path / "test_host_stage")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-pthread", "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "cpp", "test_host_stage.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "host stage copy pool ok" in out.stdout, out.stdout + out.stderr


def test_product_does_not_touch_the_oracle():
    """Nothing under the package may import, link or execute oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "verifiable-fhe-paper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".inc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "orc_" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_no_gpu_means_loud_failure(V):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(V.VpbsError):
        V.Context(0)


def test_generated_round_constants_match_oracle(oracle):
    inc = open(os.path.join(ROOT, "verifiable-fhe-paper_b200", "csrc", "poseidon_rc.inc")).read()
    vals = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ULL", inc)]
    assert vals == [int(x) for x in oracle.round_constants()]


def test_merkle_prove_index_arithmetic(V, oracle):
    rng = np.random.default_rng(3)
    for (lg, h, w) in [(5, 0, 7), (5, 2, 3), (4, 4, 9), (3, 1, 135)]:
        leaves = rng.integers(0, 2**64, size=(1 << lg, w), dtype=np.uint64)
        digests, cap = oracle.merkle_new(leaves, h)
        tree = V.MerkleTree(leaves, digests, cap)
        for i in range(1 << lg):
            got = tree.prove(i).siblings
            assert np.array_equal(got, oracle.merkle_prove(digests, 1 << lg, h, i))
            assert oracle.merkle_verify(tree.get(i), i, got, cap)


def test_argument_errors_match_plonky2_asserts(V):
    with pytest.raises(ValueError):
        V.log2_strict(6)
    with pytest.raises(ValueError):
        V.MerkleTree.new(np.zeros((6, 3), np.uint64), 1, ctx=object())
    with pytest.raises(ValueError):
        V.MerkleTree.new(np.zeros((8, 3), np.uint64), 4, ctx=object())
    with pytest.raises(ValueError):
        V.PolynomialBatch.from_values(np.zeros((2, 8), np.uint64), 1, False, 5, ctx=object())
    with pytest.raises(ValueError):
        V.PolynomialBatch.from_values(np.zeros((2, 6), np.uint64), 1, False, 1, ctx=object())


def test_reverse_bits_and_synthetic_inputs(V):
    assert V.reverse_bits(0b0011, 4) == 0b1100
    assert V.reverse_bits(1, 19) == 1 << 18
    a = V.synthetic_columns(3, 16)
    b = V.synthetic_columns(3, 16)
    assert np.array_equal(a, b) and a.shape == (3, 16) and (a < np.uint64(V.P)).all()
    assert not np.array_equal(a[0], a[1])
    raw = V.synthetic_columns(3, 16, canonical=False)
    assert np.array_equal(raw % np.uint64(V.P), a)


def test_bench_and_entry_have_no_undefined_names():
    """bench.py only runs on the GPU box: catch NameErrors (e.g. a dropped assignment) on CPU."""
    import ast
    import builtins
    for fname in ("bench.py", "__graft_entry__.py"):
        tree = ast.parse(open(os.path.join(ROOT, fname)).read())
        module_names = set(dir(builtins))
        for node in tree.body:
            if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
                module_names.add(node.name)
            elif isinstance(node, (ast.Import, ast.ImportFrom)):
                module_names.update((a.asname or a.name).split(".")[0] for a in node.names)
            elif isinstance(node, ast.Assign):
                for t in node.targets:
                    module_names.update(n.id for n in ast.walk(t) if isinstance(n, ast.Name))
        for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:  # closures share scope
            assigned = set(module_names)
            for n in ast.walk(fn):
                if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                    assigned.add(n.id)
                elif isinstance(n, ast.arg):
                    assigned.add(n.arg)
                elif isinstance(n, (ast.Import, ast.ImportFrom)):
                    assigned.update((a.asname or a.name).split(".")[0] for a in n.names)
                elif isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                    assigned.add(n.name)
                elif isinstance(n, ast.ExceptHandler) and n.name:
                    assigned.add(n.name)
            missing = {n.id for n in ast.walk(fn)
                       if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)} - assigned
            assert not missing, "%s: %s uses undefined names %s" % (fname, fn.name, sorted(missing))


def test_product_sources_carry_no_developer_knobs():
    """csrc/ is the product: no getenv() tuning knobs and no alternative kernel forms behind -D
    switches (those are in tools/variants/)."""
    pkg = os.path.join(ROOT, "verifiable-fhe-paper_b200", "csrc")
    for f in os.listdir(pkg):
        text = open(os.path.join(pkg, f)).read()
        assert "getenv" not in text, f
        for flag in ("VPBS_MDS_INT32", "VPBS_MDS_FP64_DENSE", "VPBS_SBOX_REDUCED", "VPBS_SBOX_OUTLINE",
                     "VPBS_NO_PIPE_INTERLEAVE", "VPBS_MUL_C", "VPBS_ADDSUB_C", "VPBS_HALF_I2F",
                     "VPBS_CANON_C", "VPBS_NTT_CANONICAL"):
            assert flag not in text, (f, flag)


def test_bench_help_runs():
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"],
                         capture_output=True, text=True)
    assert out.returncode == 0 and "--impl" in out.stdout


@pytest.mark.parametrize("flag", ["VPBS_MDS_INT32", "VPBS_MDS_FP64_DENSE", "VPBS_SBOX_REDUCED",
                                  "VPBS_SBOX_OUTLINE", "VPBS_NO_PIPE_INTERLEAVE", "VPBS_MUL_C",
                                  "VPBS_ADDSUB_C", "VPBS_HALF_I2F"])
def test_documented_kernel_variants_still_compile(flag, tmp_path):
    """DESIGN.md quotes measurements of alternative kernel forms; they live outside the product, in
    tools/variants/ (round-1 headers with every compile-time switch), and have to keep building for
    sm_100a (their bit-exactness is what tools/selftest.cu checks on a GPU)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "verifiable-fhe-paper_b200")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17",
                        "-I", os.path.join(pkg, "tools", "variants"), "-I", os.path.join(pkg, "csrc"),
                        "-D" + flag, "-c", "-o", str(tmp_path / "v.o"),
                        os.path.join(pkg, "tools", "selftest.cu")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]


def test_gate_program_builder_host_logic(V, oracle):
    """The gate-program assembler (pure host code): instruction packing, immediate dedup, register
    recycling at end_gate, and plonky2's selector filter prod_{i != row}(i - s) [* (UNUSED - s)] — checked
    by running the assembled program through the oracle's interpreter on constant polynomials."""
    P = 2**64 - 2**32 + 1
    b = V.GateProgramBuilder()
    assert b.imm(5) == b.imm(5 + P) and b.imm(7) != b.imm(5) and b.imms == [5, 7]
    r0 = b.add(b.wire(3), b.const(2))
    r1 = b.mad(r0, b.wire(1), b.imm(7))
    assert r0 == r1 == (0, 0) and b.nregs == 1
    op, dst, ka, kb, ia, ib = (b.code[0] & 0xff, (b.code[0] >> 8) & 0xff, (b.code[0] >> 16) & 0xf,
                               (b.code[0] >> 20) & 0xf, (b.code[0] >> 24) & 0xffff, (b.code[0] >> 40) & 0xffff)
    assert (op, dst, ka, kb, ia, ib) == (0, 0, 1, 2, 3, 2)
    assert b.code[1] & 0xff == 5 and (b.code[1] >> 8) & 0xff == 0
    b.emit(4, r1)
    assert b.num_constraints == 5
    f = b.selector_filter(0, 1, range(3), True)
    b.end_gate(f)
    assert b._next == 0                      # registers recycled
    with pytest.raises(ValueError):
        b.mad(b.wire(0), b.wire(1), b.wire(2))   # accumulator must be a register
    with pytest.raises(ValueError):
        b.into(224, b.ADD, b.wire(0), b.wire(1))
    # evaluate on constant polynomials: every wire / constant column is a constant function
    n, qdb = 4, 1
    wires = np.zeros((4, n), np.uint64); wires[:, 0] = [11, 13, 17, 19]      # coefficient form: constants
    cs = np.zeros((3, n), np.uint64); cs[:, 0] = [5, 23, 29]                 # selector s = 5
    alphas = np.array([3], np.uint64)
    got = oracle.gate_program_eval(b.code, b.imms, b.nregs, b.num_constraints, wires, cs, qdb, None, alphas)
    s = 5
    filt = (0 - s) * (2 - s) * (2**32 - 1 - s)
    want = (((19 + 29) + 13 * 7) * pow(3, 4, P) * filt) % P
    assert want != 0 and (got == np.uint64(want)).all()
