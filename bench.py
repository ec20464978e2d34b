#!/usr/bin/env python3
"""bench.py — LDE + Poseidon-Merkle commit throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one PolynomialBatch::from_values commit of BASELINE.json configs[1]
(2^16 rows x 128 Goldilocks columns, rate_bits = 3, Poseidon cap_height = 4) on synthetic values.
N > 1: every rank commits its own independent batch (one proof per GPU, no data-path collective —
north_star "independent PBS proofs shard one per GPU"), so scaling is weak and
value = N * K * 2^16 trace rows / max-over-ranks time.

Prints ONE JSON line (rank 0).  Keys beyond the base contract: roofline (dominant kernel = Poseidon
leaf hashing, integer-issue bound), roofline_hbm (the NTT/LDE kernels against HBM), cpu_baseline
(the C restatement of plonky2's CPU path, oracle/, timed on this box's host cores), e2e (host
buffers through the C ABI with H2D/D2H in the timed region), step_standin (the three commits of
one N=1024 IVC step), clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N, NCOLS, RATE_BITS, CAP_HEIGHT = 16, 128, 3, 4
METRIC = "LDE+Merkle commit rows/s (2^16-row x 128-col Goldilocks batch, rate_bits=3, cap_height=4)"
UNIT = "trace rows/s"
P_GL = 0xFFFFFFFF00000001
# integer work per Poseidon permutation in 32x32->64 multiply-accumulates
# (1,077 full modmuls x 4 + 2,304 + 44 small MACs), and the IMAD.WIDE issue rate it is held against.
IMAD_PER_PERMUTATION = 6700
IMAD_WIDE_LANES_PER_CLK_PER_SM = 64
SM_COUNT = 148


def algorithmic_bytes(ncols, n, r, h, from_values=True):
    m = n << r
    b = 8 * ncols * n + 8 * ncols * m + 64 * (m - (1 << h)) + 32 * (1 << h)
    return b + (8 * ncols * n if from_values else 0)


def permutations(ncols, n, r, h):
    m = n << r
    return (m * ((ncols + 7) // 8) if ncols > 4 else 0) + (m - (1 << h))


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks, reasons, power, mx = [], set(), [], None
        for r in self.rows:
            try:
                clocks.append(float(r[1])); mx = float(r[2]); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load": samples in the upper half of the observed power range
        if power:
            thr = (max(power) + min(power)) / 2
            loaded = [c for c, p in zip(clocks, power) if p >= thr] or clocks
        else:
            loaded = clocks
        return {"sm_mhz": statistics.median(loaded) if loaded else None, "sm_max_mhz": mx,
                "power_w_max": max(power) if power else None, "samples": len(clocks),
                "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process (and so its pinned host buffers, first-touch) to the CPUs closest to its GPU:
    with 8 ranks streaming ~0.7 GB per step over PCIe each, copies that cross the socket
    interconnect are the e2e bottleneck.  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return "cpu affinity = GPU %d's NUMA-local cores (%d cpus)" % (gpu_index, len(os.sched_getaffinity(0)))
    except Exception as e:  # not fatal: affinity is an optimisation
        return "unchanged (%s)" % type(e).__name__


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores.
    plonky2 itself cannot be built here (no Rust toolchain; crates not vendored), so this is the C
    restatement in oracle/ ("port"), with all host threads."""
    if rank != 0:
        return
    import numpy as np
    import vfhe_b200 as V
    from oracle import binding as B
    B.build()
    cores = len(os.sched_getaffinity(0))
    B.set_threads(cores)
    n = 1 << LOG_N
    cols = V.synthetic_columns(NCOLS, n)
    t0 = time.perf_counter()
    B.commit(cols, RATE_BITS, CAP_HEIGHT)          # warm-up, also sizes the sample
    t_est = time.perf_counter() - t0
    budget = 150.0
    timed = max(1, min(args.steps, int(budget / max(t_est, 1e-3))))
    extra_warm = max(0, min(args.warmup - 1, int(30.0 / max(t_est, 1e-3))))
    for _ in range(extra_warm):
        B.commit(cols, RATE_BITS, CAP_HEIGHT)
    t0 = time.perf_counter()
    for _ in range(timed):
        B.commit(cols, RATE_BITS, CAP_HEIGHT)
    dt = (time.perf_counter() - t0) / timed
    value = n / dt
    sample = ("%d of %d steps timed (bounded to ~%ds); each step = one full 2^16x128 commit by the C "
              "restatement of plonky2 0.2.0's CPU path (oracle/liboracle.so: OpenMP, %d threads, %d-lane "
              "SIMD Poseidon and FFT layers)" % (timed, args.steps, int(budget), cores, B.get_simd()))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample, "simd_lanes": B.get_simd(),
                         "us_per_permutation_per_core": cpu_perm_rate(B, np)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_multi_commit(args):
    """Strong scaling of the end-to-end call: one configs[1] commit through vpbs_commit_multi, host
    buffers in, all outputs (coefficients, leaves, digests, cap) back in host memory, G GPUs of
    this process each moving 1/G of the D2H bytes over its own link."""
    import ctypes
    import numpy as np
    import torch
    import vfhe_b200 as V
    G = args.multi_commit
    if not torch.cuda.is_available() or torch.cuda.device_count() < G:
        raise SystemExit("--multi-commit %d needs %d CUDA devices" % (G, G))
    V.build.build()
    ctxs = [V.Context(g) for g in range(G)]
    lib = ctxs[0].lib
    n, m, ncap = 1 << LOG_N, (1 << LOG_N) << RATE_BITS, 1 << CAP_HEIGHT
    host_cols = V.synthetic_columns(NCOLS, n, seed=0x5EED0000)

    def pinned(shape):
        nbytes = int(np.prod(shape)) * 8
        p = lib.vpbs_host_alloc(nbytes)
        if not p:
            raise SystemExit("vpbs_host_alloc failed")
        return np.ctypeslib.as_array((ctypes.c_uint64 * (nbytes // 8)).from_address(p)).reshape(shape)

    h_cols = pinned((NCOLS, n))
    h_cols[:] = host_cols
    h_coeffs, h_leaves = pinned((NCOLS, n)), pinned((m, NCOLS))
    h_digests, h_cap = pinned((2 * (m - ncap), 4)), pinned((ncap, 4))
    u64p = V._lib.u64p
    colp = (u64p * NCOLS)(*[h_cols[c].ctypes.data_as(u64p) for c in range(NCOLS)])
    cop = (u64p * NCOLS)(*[h_coeffs[c].ctypes.data_as(u64p) for c in range(NCOLS)])
    handles = (ctypes.c_void_p * G)(*[c.handle for c in ctxs])
    st = V.VpbsStats()

    def step():
        ctxs[0].check(lib.vpbs_commit_multi(handles, G, colp, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, 0,
                                            None, cop, h_leaves.ctypes.data_as(u64p),
                                            h_digests.ctypes.data_as(u64p),
                                            h_cap.ctypes.data_as(u64p), ctypes.byref(st)))

    sampler = ClockSampler(0)
    for _ in range(max(3, args.warmup)):
        step()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps * 4):
        step()
    dt = (time.perf_counter() - t0) / (args.e2e_steps * 4)
    clocks = sampler.stop()
    # parity with the single-GPU host call on the same inputs (cap, a digest checksum, sampled rows)
    one = V.PolynomialBatch.from_values(host_cols, RATE_BITS, False, CAP_HEIGHT, ctx=ctxs[0])
    same = bool(np.array_equal(one.merkle_tree.cap, h_cap) and
                np.array_equal(one.merkle_tree.digests, h_digests) and
                np.array_equal(one.merkle_tree.leaves[::4099], h_leaves[::4099]) and
                np.array_equal(one.polynomials, h_coeffs))
    h2d = 8 * NCOLS * n  # every column chunk crosses PCIe once (into its owner GPU), then peer copies
    d2h = 8 * NCOLS * n + 8 * m * NCOLS + 32 * 2 * (m - ncap) + 32 * ncap
    print(json.dumps({
        "metric": METRIC, "mode": "multi-commit (strong scaling of one commit, single process)",
        "value": n / dt, "unit": UNIT, "n_gpus": G, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "dtype": "u64", "data": "synthetic",
        "config": dict(workload_config(1), parallelism="one commit over %d GPUs: row ranges "
                       "(whole LDE blocks / cap subtrees) per GPU; input chunks uploaded once, round-robin "
                       "over the GPUs' host links, and passed on by peer copies" % G),
        "e2e": {"value": n / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "vpbs_commit_multi (host C ABI, pinned buffers)",
                "slowest_device_phase_ms": st.as_dict()},
        "matches_single_gpu_commit": same, "gpu_launches": int(st.kernel_launches), "clocks": clocks}),
        flush=True)
    if not same:
        sys.exit(3)


def run_fri_commit_phase(args):
    """[P2] fri/prover.rs fri_committed_trees + fri_proof_of_work at the N=1024 step's size: the
    final polynomial's LDE has 2^19 values in the quadratic extension; ConstantArityBits(4, 5) gives
    arity-16 reduction layers (2^19 -> 2^15 -> 2^11 -> 2^7 values); every layer commits to the
    bit-reversed, chunked values (Merkle tree, cap height 4) and folds with a challenge; then the
    prover grinds a 16-bit proof of work.  The challenger (Fiat-Shamir) stays on the CPU, so every
    layer is one host call with host buffers, exactly how a patched plonky2 would drive it."""
    import numpy as np
    import torch
    import vfhe_b200 as V
    from oracle import binding as orc
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the commit path has no CPU fallback")
    V.build.build()
    orc.build()
    ctx = V.Context(0)
    rng = np.random.default_rng(11)
    log_len, arity_bits, cap_h = LOG_N + RATE_BITS, 4, CAP_HEIGHT
    coeffs0 = rng.integers(0, P_GL, size=(1 << LOG_N, 2), dtype=np.uint64)
    # LDE of the final polynomial: pad to 2^19 coefficients, coset FFT with shift 7
    padded = np.zeros((1 << log_len, 2), np.uint64)
    padded[: 1 << LOG_N] = coeffs0
    values0 = np.stack([V.coset_fft(padded[:, 0].copy(), 7, ctx), V.coset_fft(padded[:, 1].copy(), 7, ctx)], 1)
    betas = rng.integers(0, P_GL, size=(8, 2), dtype=np.uint64)
    pow_state = rng.integers(0, P_GL, size=12, dtype=np.uint64)

    def phase(layer_commit, fold, grind):
        coeffs, values, shift, lg, caps, k = padded, values0, 7, log_len, [], 0
        # [P2] fri/reduction_strategies.rs ConstantArityBits(4, 5): reduce while the degree exceeds
        # 2^5 and the layer still has at least 2^cap_height leaves
        while lg - RATE_BITS > 5 and lg - arity_bits >= cap_h:
            caps.append(layer_commit(values, arity_bits, min(cap_h, lg - arity_bits)))
            shift = pow(shift, 1 << arity_bits, P_GL)
            coeffs, values = fold(coeffs, arity_bits, betas[k], shift)
            lg -= arity_bits
            k += 1
        return caps, coeffs, grind(pow_state)

    gpu = lambda: phase(lambda v, a, h: V.fri_layer_commit(v, a, h, ctx).cap,
                        lambda c, a, b, sh: V.fri_fold(c, a, b, sh, ctx),
                        lambda st: V.fri_proof_of_work(st, 5, 16, ctx=ctx))

    def chain():
        """the same phase as ONE device-resident chain (vpbs_fri_*): coefficients up once, per layer
        only the cap comes back and beta goes down; final polynomial + PoW witness at the end"""
        fri = V.FriCommitPhase(coeffs0, RATE_BITS, ctx)
        lg, caps, k = log_len, [], 0
        while lg - RATE_BITS > 5 and lg - arity_bits >= cap_h:
            caps.append(fri.commit_layer(arity_bits, min(cap_h, lg - arity_bits)))
            fri.fold(betas[k])
            lg -= arity_bits
            k += 1
        final = fri.final_poly()
        w = V.fri_proof_of_work(pow_state, 5, 16, ctx=ctx)
        fri.close()
        return caps, final, w

    sampler = ClockSampler(0)
    for _ in range(3):
        got = gpu()
        got_chain = chain()
    reps = max(5, args.e2e_steps * 2)
    t0 = time.perf_counter()
    for _ in range(reps):
        got = gpu()
    dt = (time.perf_counter() - t0) / reps
    launches0 = ctx.kernel_launches
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(reps):
        got_chain = chain()
    dt_chain = (time.perf_counter() - t0) / reps
    clocks = sampler.stop()
    chain_launches = (ctx.kernel_launches - launches0) // reps

    def cpu_grind(st):
        s = st.copy()
        for c in range(1 << 20):
            s[5] = c
            if int(orc.poseidon(s)[7]) >> 48 == 0:
                return c
        return None

    t0 = time.perf_counter()
    ref = phase(lambda v, a, h: orc.fri_layer_commit(v, a, h)["cap"],
                lambda c, a, b, sh: orc.fri_fold(c, a, b, sh), lambda st: None)
    cpu_dt = time.perf_counter() - t0
    same = all(np.array_equal(a, b) for a, b in zip(got[0], ref[0])) and np.array_equal(got[1], ref[1])
    final_ref = ref[1][: ref[1].shape[0] >> RATE_BITS]
    same_chain = (len(got_chain[0]) == len(ref[0]) and
                  all(np.array_equal(a, b) for a, b in zip(got_chain[0], ref[0])) and
                  np.array_equal(got_chain[1], final_ref))
    w = got_chain[2]
    chk = pow_state.copy()
    chk[5] = w if w is not None else 0
    pow_ok = w is not None and int(orc.poseidon(chk)[7]) >> 48 == 0 and w == got[2]
    nl = len(ref[0])
    pcie_chain = 16 * (1 << LOG_N) + nl * (32 * (1 << cap_h) + 16) + 16 * final_ref.shape[0] + 13 * 8 + 8
    print(json.dumps({
        "metric": "FRI commit phase of one N=1024 step proof (stand-in): final-polynomial LDE (2^16 -> 2^19 "
                  "extension values), 3 arity-16 layers (tree + fold each), final polynomial, 16-bit "
                  "proof-of-work grind; device-resident chain through the host C ABI (vpbs_fri_*)",
        "value": dt_chain * 1e3, "unit": "ms per commit phase", "higher_is_better": False, "n_gpus": 1,
        "steps": reps, "layers": nl, "final_poly_len": int(final_ref.shape[0]),
        "dtype": "u64", "data": "synthetic", "vs_baseline": None,
        "pcie_bytes_per_phase": pcie_chain,
        "per_layer_host_calls_ms": dt * 1e3,
        "per_layer_host_calls_note": "round-1 form (vpbs_fri_layer_commit + vpbs_fri_fold per layer: values "
                                     "up, leaves/digests/cap and folded vectors down), without the initial LDE",
        "cpu_baseline": {"value": cpu_dt * 1e3, "unit": "ms per commit phase (trees + folds, no LDE, no grind)",
                         "cores": len(os.sched_getaffinity(0)), "kind": "port",
                         "sample": "one full commit phase by oracle/liboracle.so (OpenMP)"},
        "matches_oracle": bool(same and same_chain), "pow_witness_valid": bool(pow_ok),
        "gpu_launches": int(chain_launches), "clocks": clocks,
        "note": "the Fiat-Shamir challenger stays on the CPU: per layer the cap comes back and beta goes "
                "down, nothing else crosses PCIe; SURVEY 8(f) row 1"}), flush=True)
    if not (same and same_chain and pow_ok):
        sys.exit(3)


def workload_config(world):
    return {"workload": "configs[1]: standalone commit microbench, 2^16 rows x 128 Goldilocks columns, "
                        "rate_bits=3, Poseidon cap_height=4, from values, blinding off",
            "log_n": LOG_N, "ncols": NCOLS, "rate_bits": RATE_BITS, "cap_height": CAP_HEIGHT,
            "parallelism": "one independent commit per GPU x%d, no collective" % world,
            "l2": "per-step working set ~0.7 GB (64 MiB in, 512 MiB leaves, 34 MB digests) exceeds "
                  "the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chain-steps", type=int, default=0,
                    help="BASELINE configs[3] stand-in: that many sequentially dependent N=1024 IVC "
                         "step stand-ins (3 commits each) through the host C ABI; prints its own line")
    ap.add_argument("--chain-log-n", type=int, default=LOG_N,
                    help="rows of the chained step stand-in (16: N=1024, BASELINE configs[2-4]; 13: the "
                         "N=8 test circuit's size, configs[0])")
    ap.add_argument("--chain-eager", action="store_true",
                    help="with --chain-steps: also time round 1's eager form of the step (all outputs "
                         "of the first two commits downloaded)")
    ap.add_argument("--chain-host-quotient", action="store_true",
                    help="with --chain-steps: the quotient chunks arrive as 16 coefficient columns from the "
                         "host (the form of the first half of round 2) instead of being computed on the "
                         "device from the resident batches (vpbs_batch_quotient_polys)")
    ap.add_argument("--chain-gate-ops", type=int, default=0, metavar="N",
                    help="with --chain-steps: evaluate the gate constraints ON THE DEVICE from a synthetic "
                         "program of about N field operations per point of the quotient domain (Poseidon-gate-"
                         "like rounds: x^7 S-boxes and 12 x 12 linear layers; plonky2's PoseidonGate is a few "
                         "thousand) instead of taking their alpha-reduced values from the host")
    ap.add_argument("--chain-shard", action="store_true",
                    help="with --chain-steps under torchrun: ONE chain for all ranks — every resident "
                         "batch of a step is sharded by row range over the GPUs (vpbs_ctx_set_shard), "
                         "the caps are completed by an NCCL all-gather and the query openings "
                         "collected from the owners; strong scaling of the step latency")
    ap.add_argument("--shard-commit", action="store_true",
                    help="strong scaling of ONE commit: every rank computes its row range "
                         "(vpbs_commit_shard_dev) and the subtree roots are all-gathered over NCCL")
    ap.add_argument("--fri-commit-phase", action="store_true",
                    help="SURVEY 8(f) row 1 stand-in: the FRI commit phase of one N=1024 step proof "
                         "(2^19 extension values, arity-16 layers: tree + fold per layer, then the "
                         "16-bit proof-of-work grind) through the host C ABI, beside the CPU port; "
                         "prints its own line")
    ap.add_argument("--multi-commit", type=int, default=0, metavar="G",
                    help="single process (do not launch under torchrun): ONE commit spread over G "
                         "GPUs through vpbs_commit_multi with host buffers; prints its own line")
    args = ap.parse_args()
    if args.multi_commit:
        return run_multi_commit(args)
    if args.fri_commit_phase:
        return run_fri_commit_phase(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries exactly ONE JSON line: route everything libraries print while we work (e.g.
    # NCCL's version banner) to stderr and restore fd 1 only for the result line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    import vfhe_b200 as V

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the commit path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout (ONE JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    V.build.build()
    ctx = V.Context(local_rank)
    # one explicit (non-default) stream for the library's kernels, the timing events and NCCL:
    # a NULL stream handle means "the context's own stream" to vpbs_ctx_set_stream
    bench_stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(bench_stream)
    ctx.set_stream(bench_stream.cuda_stream)

    n, m = 1 << LOG_N, (1 << LOG_N) << RATE_BITS
    ncap = 1 << CAP_HEIGHT
    host_cols = V.synthetic_columns(NCOLS, n, seed=0x5EED0000 + 1000 * rank)
    dev = torch.device("cuda", local_rank)
    d_cols = torch.from_numpy(host_cols.view(np.int64)).to(dev)
    d_coeffs = torch.empty((NCOLS, n), dtype=torch.int64, device=dev)
    d_leaves = torch.empty((m, NCOLS), dtype=torch.int64, device=dev)
    d_digests = torch.empty((2 * (m - ncap), 4), dtype=torch.int64, device=dev)
    d_cap = torch.empty((ncap, 4), dtype=torch.int64, device=dev)

    def step(stats=True):
        return V.commit_device(ctx, d_cols.data_ptr(), NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, False,
                               d_coeffs.data_ptr(), d_leaves.data_ptr(), d_digests.data_ptr(),
                               d_cap.data_ptr(), want_stats=stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fail_checks = []  # every self-check of this run; any entry makes the process exit non-zero
    checks_passed = [0]

    def check(name, ok):
        if ok:
            checks_passed[0] += 1
        else:
            fail_checks.append(name)
    if args.shard_commit:
        return run_shard_commit(args, V, ctx, rank, world, dev, barrier, max_over_ranks, emit)
    if args.chain_steps:
        return run_chain(args, V, ctx, rank, world, barrier, max_over_ranks, emit)

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step()
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_phase = []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        per_phase.append(step())
    ev1.record()
    barrier()
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.kernel_launches - launches0
    ms_per_step = elapsed_ms / args.steps
    value = world * args.steps * n / (elapsed_ms * 1e-3)

    # ---- end to end through the host C ABI (host buffers in, H2D + D2H inside the timed region)
    lib = ctx.lib
    u64p = V._lib.u64p
    pinned_ptrs = []

    def pinned(shape):
        nbytes = int(np.prod(shape)) * 8
        p = lib.vpbs_host_alloc(nbytes)
        if not p:
            raise SystemExit("vpbs_host_alloc failed")
        pinned_ptrs.append(p)
        buf = (ctypes.c_uint64 * (nbytes // 8)).from_address(p)
        return np.ctypeslib.as_array(buf).reshape(shape)

    def colptrs(a):
        return (u64p * a.shape[0])(*[a[c].ctypes.data_as(u64p) for c in range(a.shape[0])])

    def timed_host_loop(fn, steps):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return dt / steps

    h_cols = pinned((NCOLS, n))
    h_cols[:] = host_cols
    colp = colptrs(h_cols)
    h2d_bytes = 8 * NCOLS * n
    nlayers = LOG_N + RATE_BITS - CAP_HEIGHT

    # (1) e2e — the product path: PolynomialBatch::from_values with the batch left in HBM
    # (vpbs_batch_*).  Per step: H2D of the 128 value columns, the commit, D2H of the cap (the
    # commitment), the openings of all 128 polynomials at two extension points (OpeningSet: zeta,
    # g * zeta) and the 28 FRI-query rows + Merkle paths — everything a proof takes from a batch whose
    # LDE consumers run on the device.
    query_idx = np.random.default_rng(7).integers(0, m, size=28, dtype=np.uint64)
    zeta = np.random.default_rng(8).integers(0, P_GL, size=(2, 2), dtype=np.uint64)
    res_cap = np.empty((ncap, 4), np.uint64)
    res_rows = np.empty((28, NCOLS), np.uint64)
    res_sib = np.empty((28, nlayers, 4), np.uint64)
    res_open = np.empty((2, NCOLS, 2), np.uint64)

    def resident_step():
        h = ctypes.c_void_p()
        ctx.check(lib.vpbs_batch_commit(ctx.handle, colp, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None,
                                        res_cap.ctypes.data_as(u64p), ctypes.byref(h), None))
        ctx.check(lib.vpbs_batch_eval_ext2(h, zeta.ctypes.data_as(u64p), 2, res_open.ctypes.data_as(u64p)))
        ctx.check(lib.vpbs_batch_get_leaves(h, query_idx.ctypes.data_as(u64p), 28,
                                            res_rows.ctypes.data_as(u64p)))
        ctx.check(lib.vpbs_batch_prove(h, query_idx.ctypes.data_as(u64p), 28,
                                       res_sib.ctypes.data_as(u64p)))
        lib.vpbs_batch_destroy(h)

    res_s = timed_host_loop(resident_step, args.e2e_steps)
    res_d2h = 32 * ncap + 2 * NCOLS * 16 + 28 * (8 * NCOLS + 32 * nlayers)

    # (2) e2e_eager — the same commit with EVERY output copied to pinned host memory before the call
    # returns (coefficients, the 512 MiB leaf matrix, digests, cap): what plonky2's stock
    # PolynomialBatch holds, needed only while the LDE's consumers still run on the CPU
    h_coeffs, h_leaves = pinned((NCOLS, n)), pinned((m, NCOLS))
    h_digests, h_cap = pinned((2 * (m - ncap), 4)), pinned((ncap, 4))
    cop = colptrs(h_coeffs)
    e2e_stats = V.VpbsStats()

    def eager_step():
        ctx.check(lib.vpbs_commit(ctx.handle, colp, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None, cop,
                                  h_leaves.ctypes.data_as(u64p), h_digests.ctypes.data_as(u64p),
                                  h_cap.ctypes.data_as(u64p), ctypes.byref(e2e_stats)))

    eager_s = timed_host_loop(eager_step, args.e2e_steps)
    eager_d2h = 8 * NCOLS * n + 8 * m * NCOLS + 32 * 2 * (m - ncap) + 32 * ncap
    check("e2e_eager.cap_matches_device_path", np.array_equal(h_cap.view(np.int64), d_cap.cpu().numpy()))
    check("e2e.cap_matches_eager_path", np.array_equal(res_cap, h_cap))
    check("e2e.opened_rows_match_eager_leaves", np.array_equal(res_rows, h_leaves[query_idx]))
    # Merkle paths of the resident batch against the eager digests (plonky2's prove() index formula)
    eager_tree = V.MerkleTree(h_leaves, h_digests, h_cap)
    check("e2e.merkle_paths_match_eager_digests",
          all(np.array_equal(res_sib[k], eager_tree.prove(int(i)).siblings) for k, i in enumerate(query_idx)))

    # (3) e2e_pageable — the eager call with ordinary (pageable) host buffers, as a caller that
    # passes plain Vec<F> memory sees it (rank 0 only: informational)
    pageable = None
    if rank == 0:
        g_cols = host_cols.copy()
        g_coeffs, g_leaves = np.empty((NCOLS, n), np.uint64), np.empty((m, NCOLS), np.uint64)
        g_digests, g_cap = np.empty((2 * (m - ncap), 4), np.uint64), np.empty((ncap, 4), np.uint64)
        gcolp, gcop = colptrs(g_cols), colptrs(g_coeffs)

        def pageable_step():
            ctx.check(lib.vpbs_commit(ctx.handle, gcolp, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None, gcop,
                                      g_leaves.ctypes.data_as(u64p), g_digests.ctypes.data_as(u64p),
                                      g_cap.ctypes.data_as(u64p), None))

        for _ in range(2):
            pageable_step()
        t0 = time.perf_counter()
        for _ in range(max(2, args.e2e_steps // 2)):
            pageable_step()
        pg_s = (time.perf_counter() - t0) / max(2, args.e2e_steps // 2)
        check("e2e_pageable.outputs_match_pinned_path",
              np.array_equal(g_cap, h_cap) and np.array_equal(g_leaves[::4099], h_leaves[::4099])
              and np.array_equal(g_digests, h_digests))
        pageable = {"value": n / pg_s, "unit": UNIT, "ms_per_step": pg_s * 1e3,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": eager_d2h,
                    "api": "vpbs_commit, pageable host buffers (numpy.empty): the driver stages every "
                           "copy through its own pinned bounce buffers and nothing overlaps"}
        del g_coeffs, g_leaves, g_digests

        # (3b) e2e_pageable_inputs — the product path (resident batch) as a patched plonky2 calls it:
        # the value columns are the prover's own Vec<PolynomialValues<F>> — pageable, one allocation per
        # column — and reach the GPU through the context's pinned staging ring (host_stage.h); the same
        # with the ring switched off (the CUDA driver stages every copy on the calling thread)
        vec_cols = [g_cols[c].copy() for c in range(NCOLS)]
        vcolp = (u64p * NCOLS)(*[a.ctypes.data_as(u64p) for a in vec_cols])
        pg_cap = np.empty((ncap, 4), np.uint64)

        def pageable_resident_step():
            h = ctypes.c_void_p()
            ctx.check(lib.vpbs_batch_commit(ctx.handle, vcolp, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None,
                                            pg_cap.ctypes.data_as(u64p), ctypes.byref(h), None))
            ctx.check(lib.vpbs_batch_eval_ext2(h, zeta.ctypes.data_as(u64p), 2, res_open.ctypes.data_as(u64p)))
            ctx.check(lib.vpbs_batch_get_leaves(h, query_idx.ctypes.data_as(u64p), 28,
                                                res_rows.ctypes.data_as(u64p)))
            ctx.check(lib.vpbs_batch_prove(h, query_idx.ctypes.data_as(u64p), 28,
                                           res_sib.ctypes.data_as(u64p)))
            lib.vpbs_batch_destroy(h)

        pg_in = {}
        for threads in (1, 2, 8, 4, 0):
            ctx.set_host_threads(threads)
            for _ in range(2):
                pageable_resident_step()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                pageable_resident_step()
            pg_in[threads] = (time.perf_counter() - t0) / args.e2e_steps
            check("e2e_pageable_inputs.cap_matches_pinned_path_threads%d" % threads, np.array_equal(pg_cap, h_cap))
        ctx.set_host_threads(4)
        pageable["inputs_only"] = {
            "value": n / pg_in[4], "unit": UNIT, "ms_per_step": pg_in[4] * 1e3,
            "ms_per_step_driver_staging": pg_in[0] * 1e3, "copy_threads": 4,
            "ms_per_step_by_copy_threads": {str(k): round(v * 1e3, 3) for k, v in pg_in.items() if k},
            "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": res_d2h,
            "api": "vpbs_batch_commit + openings + 28 rows/paths, value columns in pageable memory (one "
                   "allocation per column, as plonky2's Vec<PolynomialValues<F>>) through the library's "
                   "pinned staging ring; *_driver_staging: ring off (vpbs_ctx_set_host_threads(0))"}
        del g_cols, vec_cols

    # (4) what the host link of this box can do (explains e2e_eager): H2D alone, D2H alone, both
    pcie = None
    if rank == 0:
        try:
            nel = 32 << 20  # 256 MiB each way, between its own pinned scratch and HBM
            scratch = pinned((2 * nel,))
            flat = torch.from_numpy(scratch.view(np.int64))
            h_a, h_b = flat[:nel], flat[nel:2 * nel]
            d_a = torch.zeros(nel, dtype=torch.int64, device=dev)
            d_b = torch.zeros(nel, dtype=torch.int64, device=dev)
            sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

            def timed(do_h2d, do_d2h):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if do_h2d:
                    with torch.cuda.stream(sa):
                        d_a.copy_(h_a, non_blocking=True)
                if do_d2h:
                    with torch.cuda.stream(sb):
                        h_b.copy_(d_b, non_blocking=True)
                torch.cuda.synchronize()
                return time.perf_counter() - t0

            timed(True, True)
            gb = nel * 8 / 1e9
            t_h = min(timed(True, False) for _ in range(3))
            t_d = min(timed(False, True) for _ in range(3))
            t_b = min(timed(True, True) for _ in range(3))
            pcie = {"h2d_gbs": gb / t_h, "d2h_gbs": gb / t_d, "both_directions_total_gbs": 2 * gb / t_b,
                    "pinned": bool(h_a.is_pinned()),
                    "e2e_eager_floor_ms": max(eager_d2h / (gb / t_d * 1e9),
                                              (h2d_bytes + eager_d2h) / (2 * gb / t_b * 1e9)) * 1e3,
                    "e2e_floor_ms": h2d_bytes / (gb / t_h * 1e9) * 1e3,
                    "note": "256 MiB copies between a pinned scratch buffer and HBM; *_floor_ms = the "
                            "step's PCIe bytes at these rates"}
            del d_a, d_b, flat, h_a, h_b, scratch
        except Exception as ex:  # informational only
            pcie = {"error": repr(ex)}

    # (5) lazy LDE pull: what a CPU quotient (compute_quotient_polys) costs on top of `e2e` while it
    # still runs on the host — every LDE row of a resident batch fetched in 2^16-row blocks
    lde_pull = None
    if rank == 0:
        h = ctypes.c_void_p()
        ctx.check(lib.vpbs_batch_commit(ctx.handle, colp, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None,
                                        res_cap.ctypes.data_as(u64p), ctypes.byref(h), None))
        blk = 1 << 16
        probe = [0, 1, 12345, m - 1]  # natural LDE row i = leaf reverse_bits(i) of the eager matrix
        expect = [h_leaves[V.reverse_bits(i, LOG_N + RATE_BITS)].copy() for i in probe]
        ctx.check(lib.vpbs_batch_get_lde_rows(h, 0, 1, blk, h_leaves.ctypes.data_as(u64p)))
        t0 = time.perf_counter()
        for b0 in range(0, m, blk):
            ctx.check(lib.vpbs_batch_get_lde_rows(h, b0, 1, blk, h_leaves[b0:b0 + blk].ctypes.data_as(u64p)))
        dt = time.perf_counter() - t0
        check("lde_pull.rows_match_eager_leaves_of_reversed_index",
              all(np.array_equal(h_leaves[i], e) for i, e in zip(probe, expect)))
        lib.vpbs_batch_destroy(h)
        lde_pull = {"ms": dt * 1e3, "bytes": 8 * m * NCOLS, "gbs": 8 * m * NCOLS / dt / 1e9,
                    "api": "vpbs_batch_get_lde_rows, 8 blocks of 2^16 natural-order rows into pinned memory"}

    clocks = sampler.stop() if rank == 0 else None

    # ---- the N=1024 step stand-in (BASELINE.json configs[2]): wires / Z / quotient commits
    step_standin = None
    if rank == 0:
        step_standin = run_step_standin(V, ctx, dev, n, m, ncap, d_digests, d_cap, pinned, colptrs,
                                        check, int_peak_gimad(clocks))

    # ---- ONE commit split by row range over all ranks (north_star: "partitioned ... by row range for
    # the Merkle subtrees, with only the subtree roots gathered over NVLink")
    shard = run_shard_commit_record(V, ctx, rank, world, dev, barrier, max_over_ranks, check,
                                    ms_per_step if world == 1 else None)

    if world > 1:  # every rank's self-checks count
        flags = [None] * world
        dist.all_gather_object(flags, list(fail_checks))
        if rank == 0:
            for r, fl in enumerate(flags[1:], 1):
                fail_checks.extend("rank %d: %s" % (r, f) for f in fl)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic = {}
    peaks = measured_peaks()
    hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    leaf_ms = statistics.mean(s["leaf_hash_ms"] for s in per_phase)
    merkle_ms = statistics.mean(s["merkle_ms"] for s in per_phase)
    ifft_ms = statistics.mean(s["ifft_ms"] for s in per_phase)
    fft_ms = statistics.mean(s["fft_ms"] for s in per_phase)
    leaf_perms = m * ((NCOLS + 7) // 8)
    int_peak = int_peak_gimad(clocks)
    int_ach = leaf_perms * IMAD_PER_PERMUTATION / (leaf_ms * 1e-3) / 1e9
    leaf_bytes = 8 * m * NCOLS + 32 * m
    lde_bytes = 8 * NCOLS * n + 8 * NCOLS * m  # coefficients in (once), leaves out
    roofline = {
        "kernel": "merkle::hash_leaves (Poseidon sponge over 2^19 rows x 128, one leaf per thread)",
        "share_of_step": leaf_ms / ms_per_step,
        "bound": "int", "achieved": int_ach, "peak": int_peak, "unit": "G IMAD.WIDE.U32-equivalent/s",
        "frac": int_ach / int_peak,
        "how": "algorithmic 32x32->64 MACs (SURVEY.md §8(d): %d per permutation x %d permutations per "
               "launch) / mean launch duration from CUDA events on the launching stream; peak = 148 SM"
               " x 64 IMAD.WIDE lanes/clk (measured, profiles/microbench_r1.jsonl) x sm_max_mhz"
               % (IMAD_PER_PERMUTATION, leaf_perms),
        "launch_ms": leaf_ms, "permutations_per_launch": leaf_perms,
        "hbm_frac": leaf_bytes / (leaf_ms * 1e-3) / 1e9 / hbm_peak,
        "traffic": traffic.get("hash_leaves_dram_bytes"),
    }
    roofline_hbm = {
        "kernels": "ntt::pass_strided_r16p<fwd> + ntt::pass_final_r16p<fwd, leaf> x 8 LDE blocks (coset LDE "
                   "with fused transpose / bit-reversal; phase \"FFT + blinding\" + \"transpose LDEs\")",
        "bound": "hbm", "achieved": lde_bytes / (fft_ms * 1e-3) / 1e9, "peak": hbm_peak,
        "unit": "GB/s", "frac": lde_bytes / (fft_ms * 1e-3) / 1e9 / hbm_peak,
        "peak_source": hbm_src, "algorithmic_bytes": lde_bytes, "ms": fft_ms,
        "traffic": traffic.get("lde_forward_dram_bytes"),
        "note": "ncu shows these kernels bound by instruction issue (ALU / FMA pipes), not by DRAM: "
                "profiles/r2_ntt_*.txt",
    }
    whole = {"algorithmic_bytes": algorithmic_bytes(NCOLS, n, RATE_BITS, CAP_HEIGHT),
             "permutations": permutations(NCOLS, n, RATE_BITS, CAP_HEIGHT),
             "hbm_frac": algorithmic_bytes(NCOLS, n, RATE_BITS, CAP_HEIGHT) / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
             "int_frac": permutations(NCOLS, n, RATE_BITS, CAP_HEIGHT) * IMAD_PER_PERMUTATION
             / (ms_per_step * 1e-3) / 1e9 / int_peak}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle, all host threads, one full commit
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import binding as B
        B.build()
        try:  # the GPU arm pinned this process to one NUMA node; the CPU arm gets every core
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except OSError:
            pass
        cores = len(os.sched_getaffinity(0))
        B.set_threads(cores)
        t0 = time.perf_counter()
        ref = B.commit(host_cols, RATE_BITS, CAP_HEIGHT)
        dt = time.perf_counter() - t0
        check("cpu_baseline.cap_matches_gpu", np.array_equal(ref["cap"], h_cap))
        check("cpu_baseline.digests_match_gpu", np.array_equal(ref["digests"], h_digests))
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "one full 2^16x128 commit (%.2f s) by the C restatement of plonky2 "
                                  "0.2.0's CPU path (oracle/liboracle.so: OpenMP over %d threads, %d-lane "
                                  "SIMD Poseidon and FFT layers, zero-padding layers skipped); plonky2 "
                                  "itself cannot be built here (no Rust toolchain)" % (dt, cores, B.get_simd()),
                        "simd_lanes": B.get_simd(),
                        "us_per_permutation_per_core": cpu_perm_rate(B, np)}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(world),
        "lde_rows_per_s": value * (1 << RATE_BITS),
        "phase_ms": {"ifft": ifft_ms, "fft_transpose": fft_ms, "merkle": merkle_ms,
                     "leaf_hash": leaf_ms},
        "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_whole_commit": whole,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": world * n / res_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": res_d2h, "ms_per_step": res_s * 1e3, "steps": args.e2e_steps,
                "api": "vpbs_batch_commit (host C ABI, pinned input columns) + vpbs_batch_eval_ext2 at 2 "
                       "points + 28 x (vpbs_batch_get_leaves, vpbs_batch_prove): the batch stays in HBM; "
                       "cap, openings, queried rows and Merkle paths come back",
                "note": "leaves / digests are NOT shipped to the host: e2e_eager is the same commit with "
                        "every output downloaded, lde_pull the cost of fetching the LDE rows on demand "
                        "for a quotient that still runs on the CPU",
                "pcie": pcie, "host_affinity": numa},
        "e2e_eager": {"value": world * n / eager_s, "unit": UNIT, "ms_per_step": eager_s * 1e3,
                      "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": eager_d2h,
                      "api": "vpbs_commit (host C ABI, pinned buffers, all outputs)",
                      "phase_ms_last": e2e_stats.as_dict()},
        "e2e_pageable": pageable,
        "lde_pull": lde_pull,
        "step_standin": step_standin,
        "shard_commit": shard,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "self_checks": {"failed": list(fail_checks), "passed": checks_passed[0]},
    }
    emit(out)
    for p in pinned_ptrs:
        lib.vpbs_host_free(p)
    if world > 1:
        dist.destroy_process_group()
    if fail_checks:
        sys.stderr.write("bench.py: self-check(s) failed: %s\n" % "; ".join(fail_checks))
        sys.exit(3)


def cpu_perm_rate(B, np):
    """Poseidon permutations of the CPU arm on ONE core (microseconds each), for context."""
    st = np.random.default_rng(1).integers(0, P_GL, size=(100000, 12), dtype=np.uint64)
    B.poseidon_batch(st[:800])
    t0 = time.perf_counter()
    B.poseidon_batch(st)
    return (time.perf_counter() - t0) / st.shape[0] * 1e6


def int_peak_gimad(clocks):
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    return SM_COUNT * IMAD_WIDE_LANES_PER_CLK_PER_SM * sm_max * 1e6 / 1e9  # G IMAD.WIDE/s


def run_step_standin(V, ctx, dev, n, m, ncap, d_digests, d_cap, pinned, colptrs, check, int_peak):
    """BASELINE.json configs[2] stand-in: the three commits of one N=1024 IVC step proof
    (135 wire + 20 Z/partial-product value columns, 16 quotient coefficient columns, 2^16 rows),
    (a) kernels only with inputs in HBM, (b) as the device-resident pipeline through the host C ABI:
    wires from host values -> Z / partial products computed on the device from the resident wires
    batch -> quotient chunks from host coefficients; per batch the cap, the openings at two extension
    points and 28 query rows + Merkle paths come back.  The real step proof (witness generation, gate
    constraints of the quotient, the FRI challenger) needs plonky2 and cannot run here."""
    import numpy as np
    import torch
    lib, u64p = ctx.lib, V._lib.u64p
    shapes = ((135, False), (20, False), (16, True))
    bufs = {}
    for (c, coeffs) in shapes:
        bufs[c] = (torch.from_numpy(V.synthetic_columns(c, n, 0x5EED0000 + c).view(np.int64)).to(dev),
                   torch.empty((c, n), dtype=torch.int64, device=dev),
                   torch.empty((m, c), dtype=torch.int64, device=dev))
    tot = []
    for it in range(5):
        t = 0.0
        for (c, coeffs) in shapes:
            a, b, l = bufs[c]
            st = V.commit_device(ctx, a.data_ptr(), c, LOG_N, RATE_BITS, CAP_HEIGHT, coeffs,
                                 b.data_ptr(), l.data_ptr(), d_digests.data_ptr(),
                                 d_cap.data_ptr(), want_stats=True)
            t += st["total_ms"]
        tot.append(t)
    del bufs
    perms = sum(permutations(c, n, RATE_BITS, CAP_HEIGHT) for c, _ in shapes)
    kernels_ms = min(tot[1:])
    # build()'s one-off constants/sigmas commit (~85 columns x 2^16, ivc_based_vpbs.rs:275)
    cs = torch.from_numpy(V.synthetic_columns(85, n, 0x5EED0000 + 85).view(np.int64)).to(dev)
    cs_co = torch.empty((85, n), dtype=torch.int64, device=dev)
    cs_le = torch.empty((m, 85), dtype=torch.int64, device=dev)
    cs_ms = min(V.commit_device(ctx, cs.data_ptr(), 85, LOG_N, RATE_BITS, CAP_HEIGHT, False, cs_co.data_ptr(),
                                cs_le.data_ptr(), d_digests.data_ptr(), d_cap.data_ptr(),
                                want_stats=True)["total_ms"] for _ in range(4))
    del cs, cs_co, cs_le

    # (b) resident pipeline
    num_routed, max_degree = 80, 8
    h_w = pinned((135, n)); h_w[:] = V.synthetic_columns(135, n, 0x5EED0000 + 135)
    h_q = pinned((16, n)); h_q[:] = V.synthetic_columns(16, n, 0x5EED0000 + 16)
    sig = V.Sigmas(V.synthetic_columns(num_routed, n, 0x51630000), V.get_unique_coset_shifts(n, num_routed), ctx)
    rng = np.random.default_rng(3)
    betas = rng.integers(0, P_GL, size=2, dtype=np.uint64)
    gammas = rng.integers(0, P_GL, size=2, dtype=np.uint64)
    zeta = rng.integers(0, P_GL, size=(2, 2), dtype=np.uint64)
    qidx = rng.integers(0, m, size=28, dtype=np.uint64)
    nlayers = LOG_N + RATE_BITS - CAP_HEIGHT
    wp, qp = colptrs(h_w), colptrs(h_q)
    caps = [np.empty((ncap, 4), np.uint64) for _ in range(3)]
    opens = [np.empty((2, c, 2), np.uint64) for c in (135, 20, 16)]
    rows = [np.empty((28, c), np.uint64) for c in (135, 20, 16)]
    sibs = [np.empty((28, nlayers, 4), np.uint64) for _ in range(3)]

    def serve(h, k):
        ctx.check(lib.vpbs_batch_eval_ext2(h, zeta.ctypes.data_as(u64p), 2, opens[k].ctypes.data_as(u64p)))
        ctx.check(lib.vpbs_batch_get_leaves(h, qidx.ctypes.data_as(u64p), 28, rows[k].ctypes.data_as(u64p)))
        ctx.check(lib.vpbs_batch_prove(h, qidx.ctypes.data_as(u64p), 28, sibs[k].ctypes.data_as(u64p)))

    def step():
        hw, hz, hq = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        ctx.check(lib.vpbs_batch_commit(ctx.handle, wp, 135, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None,
                                        caps[0].ctypes.data_as(u64p), ctypes.byref(hw), None))
        ctx.check(lib.vpbs_batch_zs_partial_products(hw, sig.handle, max_degree, betas.ctypes.data_as(u64p),
                                                     gammas.ctypes.data_as(u64p), 2, RATE_BITS, CAP_HEIGHT,
                                                     caps[1].ctypes.data_as(u64p), ctypes.byref(hz), None))
        ctx.check(lib.vpbs_batch_commit(ctx.handle, qp, 16, LOG_N, RATE_BITS, CAP_HEIGHT, 1, None,
                                        caps[2].ctypes.data_as(u64p), ctypes.byref(hq), None))
        for k, h in enumerate((hw, hz, hq)):
            serve(h, k)
        for h in (hq, hz, hw):
            lib.vpbs_batch_destroy(h)

    for _ in range(2):
        step()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        step()
    res_ms = (time.perf_counter() - t0) / reps * 1e3
    # self-check: the device-computed Z batch equals committing the host-computed columns
    zcols = V.all_wires_permutation_partial_products(h_w[:num_routed], sig, betas, gammas, max_degree)
    ref = V.PolynomialBatch.from_values(zcols, RATE_BITS, False, CAP_HEIGHT, ctx=ctx)
    check("step_standin.resident_z_batch_matches_host_pipeline",
          np.array_equal(ref.merkle_tree.cap, caps[1]) and np.array_equal(ref.merkle_tree.leaves[qidx], rows[1]))
    # (c) the quotient of the same step computed on the device from the resident batches
    # (vpbs_batch_quotient_polys: permutation terms, Z_H division, coset IFFT, 16 chunks committed), the
    # gate constraints as alpha-reduced values from pinned host memory (8 MiB, what the 16 columns were)
    h_cs = pinned((85, n)); h_cs[:] = V.synthetic_columns(85, n, 0x5EED0000 + 85)
    h_cs[5:] = V.synthetic_columns(num_routed, n, 0x51630000)
    k_is = V.get_unique_coset_shifts(n, num_routed)
    alphas = rng.integers(0, P_GL, size=2, dtype=np.uint64)
    gate = h_q.reshape(2, 8 * n)
    gp = (u64p * 2)(gate[0].ctypes.data_as(u64p), gate[1].ctypes.data_as(u64p))
    hcs, hw, hz = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    cap_tmp = np.empty((ncap, 4), np.uint64)
    ctx.check(lib.vpbs_batch_commit(ctx.handle, colptrs(h_cs), 85, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None,
                                    cap_tmp.ctypes.data_as(u64p), ctypes.byref(hcs), None))
    ctx.check(lib.vpbs_batch_commit(ctx.handle, wp, 135, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None,
                                    cap_tmp.ctypes.data_as(u64p), ctypes.byref(hw), None))
    ctx.check(lib.vpbs_batch_zs_partial_products(hw, sig.handle, max_degree, betas.ctypes.data_as(u64p),
                                                 gammas.ctypes.data_as(u64p), 2, RATE_BITS, CAP_HEIGHT,
                                                 cap_tmp.ctypes.data_as(u64p), ctypes.byref(hz), None))
    q_ms = []
    for _ in range(5):
        hq = ctypes.c_void_p()
        t0 = time.perf_counter()
        ctx.check(lib.vpbs_batch_quotient_polys(hcs, 5, hw, hz, k_is.ctypes.data_as(u64p), num_routed, max_degree,
                                                RATE_BITS, betas.ctypes.data_as(u64p), gammas.ctypes.data_as(u64p),
                                                alphas.ctypes.data_as(u64p), 2, gp, None, None, RATE_BITS,
                                                CAP_HEIGHT, cap_tmp.ctypes.data_as(u64p), ctypes.byref(hq), None))
        q_ms.append((time.perf_counter() - t0) * 1e3)
        lib.vpbs_batch_destroy(hq)
    for h in (hz, hw, hcs):
        lib.vpbs_batch_destroy(h)
    sig.close()
    h2d = 8 * n * (135 + 16)
    d2h = sum(32 * ncap + 2 * c * 16 + 28 * (8 * c + 32 * nlayers) for c in (135, 20, 16))
    return {"what": "three commits of one N=1024 IVC step (135 wire + 20 Z/partial-product value "
                    "columns, 16 quotient coefficient columns, 2^16 rows)",
            "kernels_ms": kernels_ms, "permutations": perms,
            "constants_sigmas_commit_ms": cs_ms,
            "constants_sigmas_note": "build()'s one-off commit, 85 value columns x 2^16 rows, kernels only",
            "int_frac": perms * IMAD_PER_PERMUTATION / (kernels_ms * 1e-3) / 1e9 / int_peak,
            "resident_pipeline_ms": res_ms, "resident_h2d_bytes": h2d, "resident_d2h_bytes": d2h,
            "resident_api": "vpbs_batch_commit(wires) -> vpbs_batch_zs_partial_products (Z computed and "
                            "committed on the device) -> vpbs_batch_commit(quotient, from coefficients); "
                            "per batch: cap + openings at 2 points + 28 rows and Merkle paths",
            "device_quotient_ms": min(q_ms[1:]),
            "device_quotient_note": "vpbs_batch_quotient_polys on the resident batches of the step: the "
                                    "permutation argument's vanishing terms over 2^19 points, alpha-reduced "
                                    "gate values from the host (8 MiB), Z_H division, coset IFFT, the 16 "
                                    "chunks committed; replaces the from-coefficients commit of the pipeline "
                                    "above (bench.py --chain-steps times the step with it)",
            "full_pbs_730_steps_s_commit_part": 730 * res_ms * 1e-3}


def run_shard_commit_record(V, ctx, rank, world, dev, barrier, max_over_ranks, check, single_ms):
    """ONE configs[1] commit split by row range over the `world` ranks: every rank computes
    m / world leaves + their digests (vpbs_commit_shard_dev) and the subtree roots are all-gathered
    over NCCL.  The gathered cap, the digests and a checksum of the leaves are compared with rank 0's
    single-GPU commit of the same batch."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if world & (world - 1) or world > (1 << RATE_BITS) or world > (1 << CAP_HEIGHT):
        return {"skipped": "world size %d does not divide the commit into whole LDE blocks / cap subtrees" % world}
    n, m, ncap = 1 << LOG_N, (1 << LOG_N) << RATE_BITS, 1 << CAP_HEIGHT
    cols = torch.from_numpy(V.synthetic_columns(NCOLS, n, seed=0x5EED0000).view(np.int64)).to(dev)
    torch.cuda.current_stream(dev).synchronize()
    steps, warm = 10, 3
    for _ in range(warm):
        V.commit_sharded(ctx, cols, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, False, rank, world)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        plan, leaves, digests, cap, coeffs, _ = V.commit_sharded(ctx, cols, NCOLS, LOG_N, RATE_BITS,
                                                                 CAP_HEIGHT, False, rank, world)
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps

    def checksum(t):  # position-dependent, wrapping int64 arithmetic (exact, order-independent)
        f = t.reshape(-1)
        w = torch.arange(1, f.numel() + 1, dtype=torch.int64, device=f.device) * 2 + 1
        return int((f * w).sum().item())

    mine = torch.tensor([checksum(leaves), checksum(digests)], dtype=torch.int64, device=dev)
    if world > 1:
        allsums = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allsums, mine)
    else:
        allsums = [mine]
    rec = None
    if rank == 0:
        full_leaves = torch.empty((m, NCOLS), dtype=torch.int64, device=dev)
        full_dig = torch.empty((2 * (m - ncap), 4), dtype=torch.int64, device=dev)
        full_cap = torch.empty((ncap, 4), dtype=torch.int64, device=dev)
        t_single = []
        for _ in range(4):
            st = V.commit_device(ctx, cols.data_ptr(), NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, False, 0,
                                 full_leaves.data_ptr(), full_dig.data_ptr(), full_cap.data_ptr(),
                                 want_stats=True)
            t_single.append(st["total_ms"])
        torch.cuda.synchronize()
        per = m // world
        dper = 2 * (per - ncap // world)
        ok_cap = bool(torch.equal(cap, full_cap))
        ok_sum = all(int(allsums[r][0].item()) == checksum(full_leaves[r * per:(r + 1) * per]) and
                     int(allsums[r][1].item()) == checksum(full_dig[r * dper:(r + 1) * dper])
                     for r in range(world))
        check("shard_commit.cap_matches_single_gpu", ok_cap)
        check("shard_commit.leaf_and_digest_checksums_match_single_gpu", ok_sum)
        t1 = min(t_single[1:])
        rec = {"what": "ONE 2^16 x 128 commit split by row range over %d rank(s): vpbs_commit_shard_dev per "
                       "rank + all_gather of the subtree roots (NCCL), device-resident" % world,
               "ms_per_step": ms, "single_gpu_ms": t1, "speedup": t1 / ms,
               "strong_scaling_efficiency": t1 / ms / world,
               "matches_single_gpu": bool(ok_cap and ok_sum),
               "collective": "all_gather_into_tensor of %d x 32 B roots per rank" % plan.ncap,
               "cap0": "%016x" % (int(cap[0, 0].item()) & (2**64 - 1))}
        del full_leaves, full_dig
    return rec


def run_chain(args, V, ctx, rank, world, barrier, max_over_ranks, emit):
    """BASELINE.json configs[3]/[4] stand-in (configs[0] with --chain-log-n 13): a chain of
    sequentially dependent IVC-step stand-ins, one chain per GPU.  Step k+1's inputs depend on step
    k's caps (as the real chain's witness contains the previous proof, ivc_based_vpbs.rs:329), so
    nothing can be pipelined across steps.  Two forms of the step are timed:
      resident (the product path): everything of prove() this repository puts on the device —
        wires commit (135 value columns from the host) -> Z / partial products computed and committed
        on the device (20) -> quotient polynomials: permutation-argument terms on the device from the
        resident batches' LDE rows + alpha-reduced gate constraints from the host, Z_H division, coset
        IFFT, 16 chunks committed (--chain-host-quotient: 16 coefficient columns from the host instead;
        sharded chain: values of the own rows per rank, one all-reduce, then every rank's shard) ->
        openings of all 256 polynomials of the four FRI oracles (the circuit's constants/sigmas batch,
        85 columns, is committed once and stays resident) at zeta / g zeta -> prove_openings
        (alpha-combination of the 256 + 2 polynomials, division by X - z, final-polynomial LDE) -> FRI commit phase (3 arity-16 layers: tree, cap to
        the host, beta back, fold) -> final polynomial -> 16-bit proof-of-work grind -> 28 query rounds
        served from the four resident batches and the FRI layers;
      eager (round 1's form): the three commits with every output downloaded.
    The challenges are derived from the caps by plain mixing (a stand-in for the Poseidon
    challenger, which stays on the CPU); witness generation needs plonky2 and is not part of either
    number; the gate constraints of the quotient are either supplied by the host as alpha-reduced values
    (default: their evaluation is then not part of the number) or evaluated on the device from a
    synthetic gate program (--chain-gate-ops N)."""
    import numpy as np
    lib, u64p = ctx.lib, V._lib.u64p
    log_n = args.chain_log_n
    n, m, ncap = 1 << log_n, (1 << log_n) << RATE_BITS, 1 << CAP_HEIGHT
    nlayers = log_n + RATE_BITS - CAP_HEIGHT

    def pinned(shape):
        p = lib.vpbs_host_alloc(int(np.prod(shape)) * 8)
        buf = (ctypes.c_uint64 * int(np.prod(shape))).from_address(p)
        return np.ctypeslib.as_array(buf).reshape(shape)

    def ptrs(a):
        return (u64p * a.shape[0])(*[a[c].ctypes.data_as(u64p) for c in range(a.shape[0])])

    import torch
    shard = bool(args.chain_shard) and world > 1   # ONE chain, every batch sharded over the ranks
    sp = V.ShardedProof(rank, world, torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    sharded_now = [False]

    def set_sharding(on):
        sharded_now[0] = bool(on)
        ctx.set_shard(rank if on else 0, world if on else 1)

    shapes = [(135, False), (20, False), (16, True)]
    ins, caps = [], []
    for i, (c, _) in enumerate(shapes):
        a = pinned((c, n)); a[:] = V.synthetic_columns(c, n, 0x5EED0000 + (0 if shard else 1000 * rank) + c)
        ins.append(a); caps.append(pinned((ncap, 4)))
    pin = [ptrs(a) for a in ins]
    ins0 = [a.copy() for a in ins] if shard else None
    num_routed, max_degree = 80, 8
    sig = V.Sigmas(V.synthetic_columns(num_routed, n, 0x51630000), V.get_unique_coset_shifts(n, num_routed), ctx)
    g_n = pow(7, (P_GL - 1) >> log_n, P_GL)  # generator of the trace subgroup
    k_is = V.get_unique_coset_shifts(n, num_routed)
    gate_view = ins[2].reshape(2, 8 * n)      # alpha-reduced gate constraints on the quotient domain
    gate_ptrs = (u64p * 2)(gate_view[0].ctypes.data_as(u64p), gate_view[1].ctypes.data_as(u64p))

    gate_prog = None
    if args.chain_gate_ops:
        # stand-in for the circuit's compiled gate program (the real one comes from plonky2's gates through
        # a host-side compiler): Poseidon-gate-like rounds on 12 wires — every round
        # constrains its 12 S-box inputs against 12 further wires, raises them to the 7th power and mixes
        # them with a dense 12 x 12 constant layer — until about N operations are reached
        B = V.GateProgramBuilder()
        rng_p = np.random.default_rng(1)
        # registers managed by hand: 0..11 state, 12..23 S-box outputs, 24..25 temporaries, 26..37 mixed
        R = lambda i: (0, i)
        for i in range(12):
            B.into(i, B.ADD, B.wire(i), B.imm(0))
        j = 0
        while len(B.code) < args.chain_gate_ops:
            for i in range(12):
                if j < 123:
                    B.emit(j, B.into(24, B.SUB, R(i), B.wire(12 + j % 120)))
                    j += 1
                B.into(24, B.MUL, R(i), R(i))              # x^2
                B.into(25, B.MUL, R(24), R(24))            # x^4
                B.into(25, B.MUL, R(25), R(24))            # x^6
                B.into(12 + i, B.MUL, R(25), R(i))         # x^7
            for r in range(12):
                B.into(26 + r, B.MUL, R(12), B.imm(int(rng_p.integers(1, 64))))
                for i in range(1, 12):
                    B.mad(R(26 + r), R(12 + i), B.imm(int(rng_p.integers(1, 64))))
            for r in range(12):
                B.into(r, B.ADD, R(26 + r), B.imm(int(rng_p.integers(1, 2**62))))   # + round constant
        B.into(24, B.SUB, B.imm(0), B.const(0))
        B.into(25, B.SUB, B.imm(2), B.const(0))
        B.end_gate(B.into(24, B.MUL, R(24), R(25)))         # filter of gate 1 in a group of three
        gate_prog = B.build(ctx, 123)

    def challenge(cap, k, count):  # stand-in for challenger.get_n_challenges (see docstring)
        x = cap.reshape(-1).astype(np.uint64)
        seed = int(np.bitwise_xor.reduce(x * np.uint64(2 * k + 1))) ^ (k * 0x9E3779B97F4A7C15)
        return np.random.default_rng(seed % (1 << 63)).integers(1, P_GL, size=count, dtype=np.uint64)

    # the circuit's constants/sigmas batch (85 columns, committed once in build(), ivc_based_vpbs.rs:275)
    # is the fourth FRI oracle: resident for the whole chain, opened and queried in every step
    CS = 85
    cs_cols = pinned((CS, n)); cs_cols[:] = V.synthetic_columns(CS, n, 0x5EED0000 + 85)
    cs_cols[CS - num_routed:] = V.synthetic_columns(num_routed, n, 0x51630000)  # the sigma polynomials' values
    cs_cap = np.empty((ncap, 4), np.uint64)
    cs_ptrs = ptrs(cs_cols)
    cs = {"h": None}

    def commit_cs():  # (re)commit the constants/sigmas batch under the current sharding
        if cs["h"] is not None:
            lib.vpbs_batch_destroy(cs["h"])
        if sharded_now[0]:
            cs["h"], full = sp.commit_from_host(ctx, cs_cols, RATE_BITS, CAP_HEIGHT)
            cs_cap[:] = full
            return
        cs["h"] = ctypes.c_void_p()
        ctx.check(lib.vpbs_batch_commit(ctx.handle, cs_ptrs, CS, log_n, RATE_BITS, CAP_HEIGHT, 0, None,
                                        cs_cap.ctypes.data_as(u64p), ctypes.byref(cs["h"]), None))

    set_sharding(shard)
    commit_cs()
    widths = [CS] + [c for c, _ in shapes]  # FRI_ORACLES order: constants_sigmas, wires, zs, quotient
    opens = [np.empty((2, c, 2), np.uint64) for c in widths]
    rows = [np.empty((28, c), np.uint64) for c in widths]
    sibs = [np.empty((28, nlayers, 4), np.uint64) for _ in widths]
    fri_batches = [[(o, j) for o, c in enumerate(widths) for j in range(c)], [(2, 0), (2, 1)]]
    stats = {"fri_layers": 0}
    open_ptrs = (u64p * 4)(*[a.ctypes.data_as(u64p) for a in opens])
    row_ptrs = (u64p * 4)(*[a.ctypes.data_as(u64p) for a in rows])
    sib_ptrs = (u64p * 4)(*[a.ctypes.data_as(u64p) for a in sibs])

    phase = {}

    def lap(name, t0):  # host-side time of one section of the step (every section ends synchronised)
        t1 = time.perf_counter()
        phase[name] = phase.get(name, 0.0) + (t1 - t0)
        return t1

    def resident_step():
        hs = [ctypes.c_void_p() for _ in range(3)]
        t = time.perf_counter()
        if sharded_now[0]:  # every column crosses PCIe once (1 / world per rank), then NVLink all-gather
            hs[0], full = sp.commit_from_host(ctx, ins[0], RATE_BITS, CAP_HEIGHT)
            caps[0][:] = full
        else:
            ctx.check(lib.vpbs_batch_commit(ctx.handle, pin[0], 135, log_n, RATE_BITS, CAP_HEIGHT, 0, None,
                                            caps[0].ctypes.data_as(u64p), ctypes.byref(hs[0]), None))
        t = lap("wires_commit", t)
        bg = challenge(caps[0], 1, 4)
        ctx.check(lib.vpbs_batch_zs_partial_products(hs[0], sig.handle, max_degree, bg[:2].ctypes.data_as(u64p),
                                                     bg[2:].ctypes.data_as(u64p), 2, RATE_BITS, CAP_HEIGHT,
                                                     caps[1].ctypes.data_as(u64p), ctypes.byref(hs[1]), None))
        t = lap("zs_commit", t)
        if sharded_now[0]:
            sp.complete_cap(caps[1])
            t = lap("cap_gather", t)
        ins[2][:, 0] ^= caps[1].reshape(-1)[:16] >> np.uint64(1)  # the quotient depends on alpha <- cap 1
        if sharded_now[0] and args.chain_host_quotient:
            hs[2], full = sp.commit_from_host(ctx, ins[2], RATE_BITS, CAP_HEIGHT, True)
            caps[2][:] = full
        elif sharded_now[0]:
            # the device quotient on sharded batches: every rank computes the values of its own rows, one
            # all-reduce over NVLink completes them (2 x 2^19 x 8 B), every rank commits its shard
            al = challenge(caps[1], 7, 2)
            hs[2], full = sp.quotient_polys(ctx, cs["h"], CS - num_routed, hs[0], hs[1], k_is, max_degree, RATE_BITS,
                                            bg[:2], bg[2:], al, RATE_BITS, CAP_HEIGHT, log_n,
                                            gate_terms=None if gate_prog else gate_view, program=gate_prog)
            caps[2][:] = full
        elif args.chain_host_quotient:
            ctx.check(lib.vpbs_batch_commit(ctx.handle, pin[2], 16, log_n, RATE_BITS, CAP_HEIGHT, 1, None,
                                            caps[2].ctypes.data_as(u64p), ctypes.byref(hs[2]), None))
        else:
            # compute_quotient_polys on the device: the permutation argument's vanishing terms from the
            # resident batches' LDE rows, the gate constraints as alpha-reduced values from the host
            # (the same 8 MiB the 16 coefficient columns were), Z_H division, coset IFFT, chunks, commit
            al = challenge(caps[1], 7, 2)
            ctx.check(lib.vpbs_batch_quotient_polys(
                cs["h"], CS - num_routed, hs[0], hs[1], k_is.ctypes.data_as(u64p), num_routed, max_degree,
                RATE_BITS, bg[:2].ctypes.data_as(u64p), bg[2:].ctypes.data_as(u64p), al.ctypes.data_as(u64p), 2,
                None if gate_prog else gate_ptrs, gate_prog.handle if gate_prog else None, None,
                RATE_BITS, CAP_HEIGHT, caps[2].ctypes.data_as(u64p), ctypes.byref(hs[2]), None))
        t = lap("quotient_commit", t)
        zeta = challenge(caps[2], 2, 2)
        gz = np.array([int(zeta[0]) * g_n % P_GL, int(zeta[1]) * g_n % P_GL], np.uint64)
        pts = np.stack([zeta, gz])
        allh = [cs["h"]] + hs
        # OpeningSet::new: all four oracles at zeta / g zeta in one round trip
        hp = (ctypes.c_void_p * 4)(*allh)
        ctx.check(lib.vpbs_batches_eval_ext2(hp, 4, pts.ctypes.data_as(u64p), 2, open_ptrs))
        t = lap("openings", t)
        alpha = challenge(opens[3][0, :4].copy(), 3, 2)
        obs = [_Resident(ctx, h, log_n) for h in allh]
        fri = V.FriCommitPhase.from_openings(obs, fri_batches, pts, alpha, RATE_BITS)
        t = lap("prove_openings", t)
        lg, k, cap = log_n + RATE_BITS, 0, caps[2]
        while lg - RATE_BITS > 5 and lg - 4 >= CAP_HEIGHT:  # ConstantArityBits(4, 5)
            cap = fri.commit_layer(4, min(CAP_HEIGHT, lg - 4))
            fri.fold(challenge(cap, 4 + k, 2))
            lg -= 4
            k += 1
        stats["fri_layers"] = k
        final = fri.final_poly()
        t = lap("fri_commit_phase", t)
        pow_state = challenge(final[: min(4, final.shape[0])].copy(), 9, 12)
        w = V.fri_proof_of_work(pow_state, 5, 16, ctx=ctx)
        t = lap("pow", t)
        qidx = np.random.default_rng(int(w or 0) + 1).integers(0, m, size=28, dtype=np.uint64)
        if not sharded_now[0]:  # initial_trees_proof of the 28 query rounds: one round trip for all oracles
            ctx.check(lib.vpbs_batches_open(hp, 4, qidx.ctypes.data_as(u64p), 28, row_ptrs, sib_ptrs))
        else:  # every rank opens the rows its shard holds; one all-reduce hands all of them to everyone
            own = sp.owned(qidx, m)
            q_own = np.ascontiguousarray(qidx[own])
            for b in range(4):
                rows[b][:] = 0
                sibs[b][:] = 0
            if len(q_own):
                r = [np.empty((len(q_own), widths[b]), np.uint64) for b in range(4)]
                sb = [np.empty((len(q_own), nlayers, 4), np.uint64) for b in range(4)]
                ctx.check(lib.vpbs_batches_open(hp, 4, q_own.ctypes.data_as(u64p), len(q_own),
                                                (u64p * 4)(*[a.ctypes.data_as(u64p) for a in r]),
                                                (u64p * 4)(*[a.ctypes.data_as(u64p) for a in sb])))
                for b in range(4):
                    rows[b][own] = r[b]
                    sibs[b][own] = sb[b]
            sp.collect(rows + sibs)
        qi = qidx.copy()
        for layer in range(k):
            qi = qi >> np.uint64(4)
            fri.query(layer, qi)
        t = lap("queries", t)
        fri.close()
        for h in reversed(hs):
            lib.vpbs_batch_destroy(h)
        ins[0][:, 0] ^= np.resize(caps[2].reshape(-1) ^ cap.reshape(-1)[:1], 135) >> np.uint64(1)  # next step
        lap("teardown", t)

    class _Resident:  # the two attributes FriCommitPhase.from_openings reads
        def __init__(self, c, h, lg):
            self.ctx, self.handle, self.degree_log = c, h, lg

    out_eager = None
    if args.chain_eager:
        coeffs = [pinned((c, n)) for c, _ in shapes]
        leaves = [pinned((m, c)) for c, _ in shapes[:2]]
        digests = [pinned((2 * (m - ncap), 4)) for _ in shapes[:2]]
        pco = [ptrs(a) for a in coeffs]
        qidx0 = np.random.default_rng(1).integers(0, m, size=28, dtype=np.uint64)

        def eager_step():
            for i in range(2):
                ctx.check(lib.vpbs_commit(ctx.handle, pin[i], shapes[i][0], log_n, RATE_BITS, CAP_HEIGHT, 0,
                                          None, pco[i], leaves[i].ctypes.data_as(u64p),
                                          digests[i].ctypes.data_as(u64p), caps[i].ctypes.data_as(u64p), None))
                ins[i + 1][:, 0] ^= caps[i].reshape(-1)[: shapes[i + 1][0]] >> np.uint64(1)
            h = ctypes.c_void_p()
            ctx.check(lib.vpbs_batch_commit(ctx.handle, pin[2], 16, log_n, RATE_BITS, CAP_HEIGHT, 1, None,
                                            caps[2].ctypes.data_as(u64p), ctypes.byref(h), None))
            ctx.check(lib.vpbs_batch_get_leaves(h, qidx0.ctypes.data_as(u64p), 28, rows[2].ctypes.data_as(u64p)))
            ctx.check(lib.vpbs_batch_prove(h, qidx0.ctypes.data_as(u64p), 28, sibs[2].ctypes.data_as(u64p)))
            lib.vpbs_batch_destroy(h)
            ins[0][:, 0] ^= np.resize(caps[2].reshape(-1), 135) >> np.uint64(1)

        for _ in range(3):
            eager_step()
        barrier()
        k_e = min(args.chain_steps, 64)
        t0 = time.perf_counter()
        for _ in range(k_e):
            eager_step()
        out_eager = max_over_ranks(time.perf_counter() - t0) / k_e * 1e3

    shard_check = None
    if shard:
        # the sharded step must reproduce the unsharded one bit for bit: caps of the three commits, the
        # opened rows and the Merkle paths of all four batches, over two dependent steps from the same start
        def trace(on):
            set_sharding(on)
            commit_cs()
            for a, a0 in zip(ins, ins0):
                a[:] = a0
            out = []
            for _ in range(2):
                resident_step()
                out.append([c.copy() for c in caps] + [cs_cap.copy()] + [r.copy() for r in rows] +
                           [x.copy() for x in sibs])
            return out
        ta, tb = trace(False), trace(True)
        shard_check = all(np.array_equal(x, y) for sa, sb_ in zip(ta, tb) for x, y in zip(sa, sb_))
        for a, a0 in zip(ins, ins0):
            a[:] = a0
    for _ in range(3):
        resident_step()
    barrier()
    phase.clear()
    l0 = ctx.kernel_launches
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.chain_steps):
        resident_step()
    dt = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    launches = (ctx.kernel_launches - l0) // args.chain_steps
    barrier()
    sig.close()
    lib.vpbs_batch_destroy(cs["h"])
    if shard:
        ok = [None] * world
        import torch.distributed as dist
        dist.all_gather_object(ok, bool(shard_check))
        shard_check = all(ok)
    if rank == 0:
        h2d = 8 * n * (135 + 16) + 16 * 8
        d2h = (3 * 32 * ncap + sum(widths) * 2 * 16 + stats["fri_layers"] * 32 * ncap +
               28 * sum(8 * c + 32 * nlayers for c in widths))
        emit(({
            "metric": "N=%d-class vPBS IVC step stand-in (2^%d rows): wires / Z / quotient commits, openings, "
                      "prove_openings and the FRI commit phase device-resident through the host C ABI, "
                      "sequentially dependent chain" % (1024 if log_n == 16 else 8 if log_n == 13 else 0, log_n),
            "value": dt / args.chain_steps * 1e3, "unit": "ms per step (device-side scope of prove())",
            "higher_is_better": False, "n_gpus": world, "steps": args.chain_steps,
            "chains": 1 if shard else world,
            "steps_per_s_all_gpus": (1 if shard else world) * args.chain_steps / dt,
            "full_pbs_730_steps_s": 730 * dt / args.chain_steps, "scaling": "strong" if shard else "weak",
            "sharded_step": None if not shard else {
                "ranks": world, "matches_unsharded_step": shard_check,
                "what": "ONE chain: every resident batch of a step (constants/sigmas, wires, Z, quotient) "
                        "holds the rows of its rank only (vpbs_ctx_set_shard); host columns cross PCIe "
                        "once (1 / ranks of them per rank) and reach the other GPUs by an NCCL "
                        "all-gather over NVLink; per commit the cap is "
                        "completed by an NCCL all-gather of 32 B per entry, per step the 28 x 4 opened "
                        "rows + paths are collected by one all-reduce; the quotient values of a rank's rows "
                        "are completed by one all-reduce (8 MiB); IFFTs, Z, openings and the FRI "
                        "commit phase are computed by every rank (they need all coefficients)"},
            "dtype": "u64",
            "data": "synthetic", "vs_baseline": None, "log_n": log_n,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step_approx": d2h,
            "gpu_launches_per_step": int(launches), "fri_layers": stats["fri_layers"],
            "quotient": ("16 coefficient columns from the host" if args.chain_host_quotient else
                         "device: permutation terms + tail; gate constraints %s" %
                         ("evaluated on the device from a synthetic %d-instruction program (%d registers)"
                          % (len(gate_prog.code), gate_prog.nregs) if gate_prog else
                          "as alpha-reduced values from the host (8 MiB)")),
            "phase_ms_rank0": {k: round(v / args.chain_steps * 1e3, 4) for k, v in phase.items()},
            "eager_commits_ms_per_step": out_eager,
            "eager_note": "round 1's form of the step: the three commits with coefficients, LDE rows and "
                          "digests of the first two downloaded to pinned host memory (--chain-eager)",
            "clocks": clocks,
            "note": "witness generation and the Poseidon challenger run in plonky2 on the CPU and are not part "
                    "of this number; the gate constraints of the quotient: see `quotient` (as values from "
                    "the host their evaluation is not part of it either); challenges are derived from the "
                    "caps by plain mixing"}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if shard and not shard_check:
        sys.exit(3)


def run_shard_commit(args, V, ctx, rank, world, dev, barrier, max_over_ranks, emit):
    """--shard-commit: only the sharded-commit record (see run_shard_commit_record), as its own line."""
    failed = []
    rec = run_shard_commit_record(V, ctx, rank, world, dev, barrier, max_over_ranks,
                                  lambda name, ok: None if ok else failed.append(name), None)
    if rank == 0:
        emit(dict({"metric": METRIC + " — ONE commit sharded by row range", "unit": "ms per commit",
                   "n_gpus": world, "scaling": "strong", "dtype": "u64", "data": "synthetic",
                   "config": workload_config(world), "self_checks_failed": failed}, **(rec or {})))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if failed:
        sys.exit(3)


if __name__ == "__main__":
    main()
