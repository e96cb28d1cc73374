"""Module path of the reference's intravoxel_algorithms/algorithms.py — GPU-backed solvers."""
from ..reference_api import nnls, nnls_gcv, nnls_lcurve_wrapper, nnls_tik, nnls_x2  # noqa: F401
