"""Module path of the reference's flip_angle_algorithms/fa_estimation.py — GPU-backed row workers."""
from ..reference_api import (compute_optimal_FA, fitting_slice_FA_brute_force,  # noqa: F401
                             fitting_slice_FA_spline_method)
