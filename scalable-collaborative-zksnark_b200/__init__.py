"""scz-b200: the dist-primitive hot path of Scalable-Collaborative-zkSNARK on B200.

Everything computes in libscz.so (hand-written sm_100a CUDA behind the C ABI of
include/scz.h).  This package is the Python mirror of the reference's Rust
interface for that path (PackedSharingParams, d_msm, ...); PyTorch is used only
for device memory, streams and torch.distributed.  There is no CPU fallback:
importing works anywhere (so the symbol table can be checked), but creating a
Context without a CUDA device raises.
"""
from .binding import LIB_PATH, SczError, lib  # noqa: F401
from .api import (  # noqa: F401
    Context,
    PackedSharingParams,
    d_msm,
    d_msm_leader,
    msm_batched,
    PolynomialCommitment,
    acc_product_tree,
    c_sumcheck_product,
    d_acc_product,
    d_sumcheck_product,
    degree_reduce,
    fix_variable,
    fr_pointwise,
    pss2ss,
    sumcheck_product,
    sumcheck,
    c_sumcheck,
    d_sumcheck,
    sumcheck_rounds,
    msm,
    PackedProvingParameters,
    HyperPlonkProof,
    ProofReader,
    dhyperplonk,
    dhyperplonk_data_parallel,
    dpermcheck,
    cpermcheck,
    local_hyperplonk,
    c_acc_product_and_share,
    hp_table_sizes,
    msm_g2,
    msm_g2_batched,
    d_msm_g2,
    d_msm_g2_leader,
    g2_op,
    g2_affine_to_jac,
    G2_GENERATOR_AFFINE,
)
from .delegator import Delegator, read_vec_fr  # noqa: F401
from .mpcnet import MPCNetError, MultiplexedStreamID, TorchDistMPCNet  # noqa: F401
