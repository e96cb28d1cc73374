"""Module path of the reference's epg/epg.py — GPU-backed create_Dic_3D / create_met2_design_matrix_epg."""
from ..reference_api import create_Dic_3D, create_met2_design_matrix_epg  # noqa: F401
