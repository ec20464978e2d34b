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
              "restatement of plonky2 0.2.0's CPU path (oracle/liboracle.so, OpenMP, %d threads)"
              % (timed, args.steps, int(budget), cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
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
    h2d = 8 * NCOLS * n * G
    d2h = 8 * NCOLS * n + 8 * m * NCOLS + 32 * 2 * (m - ncap) + 32 * ncap
    print(json.dumps({
        "metric": METRIC, "mode": "multi-commit (strong scaling of one commit, single process)",
        "value": n / dt, "unit": UNIT, "n_gpus": G, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "dtype": "u64", "data": "synthetic",
        "config": dict(workload_config(1), parallelism="one commit over %d GPUs: row ranges "
                       "(whole LDE blocks / cap subtrees) per GPU, no GPU-to-GPU traffic" % G),
        "e2e": {"value": n / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "vpbs_commit_multi (host C ABI, pinned buffers)",
                "slowest_device_phase_ms": st.as_dict()},
        "matches_single_gpu_commit": same, "gpu_launches": int(st.kernel_launches), "clocks": clocks}),
        flush=True)


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
    sampler = ClockSampler(0)
    for _ in range(3):
        got = gpu()
    sampler.start()
    t0 = time.perf_counter()
    reps = max(5, args.e2e_steps * 2)
    for _ in range(reps):
        got = gpu()
    dt = (time.perf_counter() - t0) / reps
    clocks = sampler.stop()

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
    w = got[2]
    chk = pow_state.copy()
    chk[5] = w if w is not None else 0
    pow_ok = w is not None and int(orc.poseidon(chk)[7]) >> 48 == 0
    print(json.dumps({
        "metric": "FRI commit phase of one N=1024 step proof (stand-in): 3 arity-16 layers from 2^19 "
                  "extension values (tree + fold each) + 16-bit proof-of-work grind, host C ABI",
        "value": dt * 1e3, "unit": "ms per commit phase", "higher_is_better": False, "n_gpus": 1,
        "steps": reps, "layers": len(got[0]), "final_poly_len": int(got[1].shape[0]) >> RATE_BITS,
        "dtype": "u64", "data": "synthetic", "vs_baseline": None,
        "cpu_baseline": {"value": cpu_dt * 1e3, "unit": "ms per commit phase (trees + folds, no grind)",
                         "cores": len(os.sched_getaffinity(0)), "kind": "port",
                         "sample": "one full commit phase by oracle/liboracle.so (OpenMP)"},
        "matches_oracle": bool(same), "pow_witness_valid": bool(pow_ok),
        "gpu_launches": int(ctx.kernel_launches), "clocks": clocks,
        "note": "the Fiat-Shamir challenger stays on the CPU: every layer is one host call (H2D of the "
                "layer's values, D2H of leaves/digests/cap); SURVEY 8(f) row 1"}), flush=True)


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
    if args.shard_commit:
        return run_shard_commit(args, V, ctx, d_cols, rank, world, dev, barrier, max_over_ranks, emit)
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

    # ---- end to end through the host C ABI: pinned host buffers, H2D + D2H inside the timed region
    lib = ctx.lib

    def pinned(shape):
        nbytes = int(np.prod(shape)) * 8
        p = lib.vpbs_host_alloc(nbytes)
        if not p:
            raise SystemExit("vpbs_host_alloc failed")
        buf = (ctypes.c_uint64 * (nbytes // 8)).from_address(p)
        return np.ctypeslib.as_array(buf).reshape(shape), p

    h_cols, p0 = pinned((NCOLS, n))
    h_cols[:] = host_cols
    h_coeffs, p1 = pinned((NCOLS, n))
    h_leaves, p2 = pinned((m, NCOLS))
    h_digests, p3 = pinned((2 * (m - ncap), 4))
    h_cap, p4 = pinned((ncap, 4))
    u64p = V._lib.u64p
    colp = (u64p * NCOLS)(*[h_cols[c].ctypes.data_as(u64p) for c in range(NCOLS)])
    cop = (u64p * NCOLS)(*[h_coeffs[c].ctypes.data_as(u64p) for c in range(NCOLS)])
    e2e_stats = V.VpbsStats()

    def e2e_step():
        ctx.check(lib.vpbs_commit(ctx.handle, colp, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None, cop,
                                  h_leaves.ctypes.data_as(u64p), h_digests.ctypes.data_as(u64p),
                                  h_cap.ctypes.data_as(u64p), ctypes.byref(e2e_stats)))

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * args.e2e_steps * n / e2e_s
    h2d_bytes = 8 * NCOLS * n
    d2h_bytes = 8 * NCOLS * n + 8 * m * NCOLS + 32 * 2 * (m - ncap) + 32 * ncap
    cap_matches = bool(np.array_equal(h_cap.view(np.int64), d_cap.cpu().numpy()))

    # ---- what the host link of this box can do (explains e2e): H2D alone, D2H alone, both at once
    pcie = None
    if rank == 0:
        try:
            nel = 32 << 20  # 256 MiB each way, between its own pinned scratch and HBM
            scratch, p_scratch = pinned((2 * nel,))
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
                    "e2e_floor_ms": max(d2h_bytes / (gb / t_d * 1e9),
                                        (h2d_bytes + d2h_bytes) / (2 * gb / t_b * 1e9)) * 1e3,
                    "note": "256 MiB copies between a pinned scratch buffer and HBM; e2e_floor_ms = the "
                            "step's PCIe bytes at these rates (D2H alone, or all bytes at the "
                            "two-direction total, whichever is larger)"}
            del d_a, d_b, flat, h_a, h_b, scratch
            lib.vpbs_host_free(p_scratch)
        except Exception as ex:  # informational only
            pcie = {"error": repr(ex)}

    # ---- the same call with the batch left in HBM (vpbs_batch_*): only the cap crosses PCIe at
    # commit time; the 28 FRI-query rows + Merkle paths of a proof are fetched on demand
    query_idx = np.random.default_rng(7).integers(0, m, size=28, dtype=np.uint64)
    res_cap = np.empty((ncap, 4), np.uint64)
    res_rows = np.empty((28, NCOLS), np.uint64)
    res_sib = np.empty((28, LOG_N + RATE_BITS - CAP_HEIGHT, 4), np.uint64)

    def resident_step():
        h = ctypes.c_void_p()
        ctx.check(lib.vpbs_batch_commit(ctx.handle, colp, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, 0, None,
                                        res_cap.ctypes.data_as(u64p), ctypes.byref(h), None))
        ctx.check(lib.vpbs_batch_get_leaves(h, query_idx.ctypes.data_as(u64p), 28,
                                            res_rows.ctypes.data_as(u64p)))
        ctx.check(lib.vpbs_batch_prove(h, query_idx.ctypes.data_as(u64p), 28,
                                       res_sib.ctypes.data_as(u64p)))
        lib.vpbs_batch_destroy(h)

    for _ in range(2):
        resident_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        resident_step()
    torch.cuda.synchronize()
    res_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    resident_ok = bool(np.array_equal(res_cap, h_cap) and np.array_equal(res_rows, h_leaves[query_idx]))
    clocks = sampler.stop() if rank == 0 else None

    # ---- the N=1024 step stand-in (BASELINE.json configs[2]): wires / Z / quotient commits
    step_standin = None
    if rank == 0:
        tot = []
        bufs = {}
        for (c, coeffs) in ((135, False), (20, False), (16, True)):
            bufs[c] = (torch.from_numpy(V.synthetic_columns(c, n, 0x5EED0000 + c).view(np.int64)).to(dev),
                       torch.empty((c, n), dtype=torch.int64, device=dev),
                       torch.empty((m, c), dtype=torch.int64, device=dev))
        for it in range(4):
            t = 0.0
            for (c, coeffs) in ((135, False), (20, False), (16, True)):
                a, b, l = bufs[c]
                st = V.commit_device(ctx, a.data_ptr(), c, LOG_N, RATE_BITS, CAP_HEIGHT, coeffs,
                                     b.data_ptr(), l.data_ptr(), d_digests.data_ptr(),
                                     d_cap.data_ptr(), want_stats=True)
                t += st["total_ms"]
            tot.append(t)
        step_standin = {"what": "three commits of one N=1024 IVC step (135 + 20 value columns, 16 "
                                "coefficient columns, 2^16 rows), kernels only, inputs in HBM",
                        "ms": min(tot[1:]), "permutations": sum(permutations(c, n, RATE_BITS, CAP_HEIGHT)
                                                                 for c in (135, 20, 16))}
        del bufs

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
    sm_mhz = (clocks or {}).get("sm_mhz") or (peaks or {}).get("sm_max_mhz", 1965.0)
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    leaf_perms = m * ((NCOLS + 7) // 8)
    int_peak = SM_COUNT * IMAD_WIDE_LANES_PER_CLK_PER_SM * sm_max * 1e6 / 1e9  # G IMAD.WIDE/s
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
        "note": "ncu shows these kernels bound by the ALU pipe and load latency, not by DRAM "
                "(DRAM throughput ~11 % of peak): profiles/r1_ntt_r16p_kernels.txt",
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
        same = bool(np.array_equal(ref["cap"].view(np.int64), d_cap_check(ctx, V, host_cols, np)))
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "one full 2^16x128 commit (%.1f s) by the C restatement of plonky2 "
                                  "0.2.0's CPU path (oracle/liboracle.so, OpenMP); plonky2 itself "
                                  "cannot be built here (no Rust toolchain)" % dt,
                        "cap_matches_gpu": same}

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
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_s / args.e2e_steps * 1e3,
                "steps": args.e2e_steps, "api": "vpbs_commit (host C ABI, pinned buffers)",
                "phase_ms_last": e2e_stats.as_dict(), "cap_matches_device_path": cap_matches,
                "pcie": pcie,
                "host_affinity": numa},
        "e2e_resident": {"value": world * args.e2e_steps * n / res_s, "unit": UNIT,
                         "ms_per_step": res_s / args.e2e_steps * 1e3,
                         "h2d_bytes_per_step": h2d_bytes,
                         "d2h_bytes_per_step": 32 * ncap + 28 * (8 * NCOLS + 32 * (LOG_N + RATE_BITS - CAP_HEIGHT)),
                         "api": "vpbs_batch_commit + 28 x (vpbs_batch_get_leaves, vpbs_batch_prove): batch "
                                "stays in HBM, cap + queried rows/paths only",
                         "matches_eager_path": resident_ok},
        "step_standin": step_standin,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    emit(out)
    for p in (p0, p1, p2, p3, p4):
        lib.vpbs_host_free(p)
    if world > 1:
        dist.destroy_process_group()


def run_chain(args, V, ctx, rank, world, barrier, max_over_ranks, emit):
    """BASELINE.json configs[3]/[4] stand-in: a chain of sequentially dependent IVC-step stand-ins,
    one chain per GPU.  Each step = the three commits of one N=1024 step proof through the HOST C
    ABI: wires (135 value columns) and Z/partial products (20) with full outputs (the CPU prover
    reads their LDE), quotient chunks (16 coefficient columns) as a resident batch opened at 28
    query positions.  Step k+1's inputs depend on step k's caps (as the real chain's witness
    contains the previous proof, ivc_based_vpbs.rs:329), so nothing can be pipelined across steps.
    The real step proof (witness generation, quotient, FRI) needs plonky2 and cannot run here."""
    import numpy as np
    lib, u64p = ctx.lib, V._lib.u64p
    n, m, ncap = 1 << LOG_N, (1 << LOG_N) << RATE_BITS, 1 << CAP_HEIGHT

    def pinned(shape):
        p = lib.vpbs_host_alloc(int(np.prod(shape)) * 8)
        buf = (ctypes.c_uint64 * int(np.prod(shape))).from_address(p)
        return np.ctypeslib.as_array(buf).reshape(shape)

    shapes = [(135, False), (20, False), (16, True)]
    ins, coeffs, leaves, digests, caps = [], [], [], [], []
    for i, (c, _) in enumerate(shapes):
        a = pinned((c, n)); a[:] = V.synthetic_columns(c, n, 0x5EED0000 + 1000 * rank + c)
        ins.append(a); coeffs.append(pinned((c, n))); caps.append(pinned((ncap, 4)))
        if i < 2:
            leaves.append(pinned((m, c))); digests.append(pinned((2 * (m - ncap), 4)))
    qidx = np.random.default_rng(1).integers(0, m, size=28, dtype=np.uint64)
    qrows = np.empty((28, 16), np.uint64)
    qsib = np.empty((28, LOG_N + RATE_BITS - CAP_HEIGHT, 4), np.uint64)

    def ptrs(a):
        return (u64p * a.shape[0])(*[a[c].ctypes.data_as(u64p) for c in range(a.shape[0])])

    pin, pco = [ptrs(a) for a in ins], [ptrs(a) for a in coeffs]

    def step():
        for i in range(2):
            ctx.check(lib.vpbs_commit(ctx.handle, pin[i], shapes[i][0], LOG_N, RATE_BITS, CAP_HEIGHT, 0,
                                      None, pco[i], leaves[i].ctypes.data_as(u64p),
                                      digests[i].ctypes.data_as(u64p), caps[i].ctypes.data_as(u64p), None))
            ins[i + 1][:, 0] ^= caps[i].reshape(-1)[: shapes[i + 1][0]] >> np.uint64(1)  # Fiat-Shamir-like dependency
        h = ctypes.c_void_p()
        ctx.check(lib.vpbs_batch_commit(ctx.handle, pin[2], 16, LOG_N, RATE_BITS, CAP_HEIGHT, 1, None,
                                        caps[2].ctypes.data_as(u64p), ctypes.byref(h), None))
        ctx.check(lib.vpbs_batch_get_leaves(h, qidx.ctypes.data_as(u64p), 28, qrows.ctypes.data_as(u64p)))
        ctx.check(lib.vpbs_batch_prove(h, qidx.ctypes.data_as(u64p), 28, qsib.ctypes.data_as(u64p)))
        lib.vpbs_batch_destroy(h)
        ins[0][:, 0] ^= np.resize(caps[2].reshape(-1), 135) >> np.uint64(1)  # next step depends on this one

    for _ in range(3):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.chain_steps):
        step()
    dt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    if rank == 0:
        emit(({
            "metric": "N=1024 vPBS IVC step stand-in (3 commits: 135 + 20 value columns, 16 coefficient "
                      "columns, 2^16 rows) through the host C ABI, sequentially dependent chain",
            "value": dt / args.chain_steps * 1e3, "unit": "ms per step (commit part only)",
            "higher_is_better": False, "n_gpus": world, "steps": args.chain_steps,
            "chains": world, "steps_per_s_all_gpus": world * args.chain_steps / dt,
            "full_pbs_730_steps_s": 730 * dt / args.chain_steps, "scaling": "weak", "dtype": "u64",
            "data": "synthetic", "vs_baseline": None,
            "note": "commit path only: witness generation, quotient polynomials and FRI of the real "
                    "step proof run in plonky2 on the CPU and are not part of this number"}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_shard_commit(args, V, ctx, d_cols, rank, world, dev, barrier, max_over_ranks, emit):
    """One 2^16 x 128 commit split by row range over all ranks (SURVEY.md §8(e) partitioning B):
    every rank holds all columns, computes m / world leaves + their digests, and only the subtree
    roots (32 B per cap entry) are exchanged.  Strong scaling; prints its own JSON line."""
    import torch
    # every rank must commit the SAME batch
    cols = torch.from_numpy(V.synthetic_columns(NCOLS, 1 << LOG_N, seed=0x5EED0000).view("int64")).to(dev)
    for _ in range(args.warmup):
        V.commit_sharded(ctx, cols, NCOLS, LOG_N, RATE_BITS, CAP_HEIGHT, False, rank, world)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        plan, leaves, digests, cap, coeffs, _ = V.commit_sharded(ctx, cols, NCOLS, LOG_N, RATE_BITS,
                                                                 CAP_HEIGHT, False, rank, world)
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    if rank == 0:
        emit(({
            "metric": METRIC + " — ONE commit sharded by row range", "value": (1 << LOG_N) / (ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": workload_config(world),
            "collective": "all_gather of %d subtree roots per rank (NCCL)" % plan.ncap,
            "cap0": "%016x" % (int(cap[0, 0].item()) & (2**64 - 1))}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def d_cap_check(ctx, V, host_cols, np):
    """cap of the same batch through the host API (for the cpu_baseline cross-check)."""
    b = V.PolynomialBatch.from_values(host_cols, RATE_BITS, False, CAP_HEIGHT, ctx=ctx)
    return b.merkle_tree.cap.view(np.int64)


if __name__ == "__main__":
    main()
