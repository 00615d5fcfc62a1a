"""sorobn_b200: B200-native exact inference for Bayesian networks.

A drop-in for the exact-inference path of MaxHalford/sorobn
(`BayesNet.query(..., algorithm="exact")`, `BayesNet.impute`), with the
factor-product / sum-out loop running as hand-written sm_100a CUDA kernels.
"""
from . import examples, planner, sharding, structure, synthetic, workloads
from .bayes_net import BayesNet

__version__ = "0.1.0"
__all__ = ["BayesNet", "examples", "planner", "sharding", "structure", "synthetic", "workloads"]
