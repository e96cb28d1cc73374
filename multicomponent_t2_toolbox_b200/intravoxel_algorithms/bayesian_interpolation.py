"""Module path of the reference's intravoxel_algorithms/bayesian_interpolation.py — GPU-backed BayesReg."""
from ..reference_api import BayesReg_nnls, nnls  # noqa: F401
