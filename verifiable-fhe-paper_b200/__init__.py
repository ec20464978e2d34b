"""verifiable-fhe-paper_b200 — B200 (sm_100a) polynomial-commitment path of vPBS.

Only what the hot path needs: csrc/ (CUDA kernels + C ABI, built into libvpbs_commit.so),
the ctypes binding and a host-side mirror of plonky2's PolynomialBatch / MerkleTree / fft API.
The directory name is not a Python identifier; import it through the root module `vfhe_b200`.
"""
from . import _lib, build, sharding  # noqa: F401
from ._lib import VpbsError, VpbsStats  # noqa: F401
from .plonky2_api import (  # noqa: F401
    COSET_SHIFT, P, SALT_SIZE, Context, FriCommitPhase, GateProgram, GateProgramBuilder, MerkleProof, MerkleTree, PolynomialBatch,
    ResidentMerkleTree, ResidentPolynomialBatch, Sigmas, all_wires_permutation_partial_products,
    commit_quotient_polys, commit_resident, commit_resident_device, commit_zs_partial_products, coset_fft,
    get_unique_coset_shifts,
    commit_device, commit_shard_device, default_context, eval_ext2, fft, fri_fold, fri_layer_commit,
    fri_proof_of_work, hash_or_noop, ifft, lde_values,
    log2_strict, open_all_at_leaves, open_all_at_points, poseidon, reverse_bits, synthetic_columns, two_to_one,
    verify_merkle_proof_to_cap)
from .sharding import ShardPlan, ShardedProof, commit_sharded, gather_cap, shard_plan  # noqa: F401,E402
