"""FV_flux: empty subclass kept for the reference's MRO (src/FVMmodel/FVdiscretization/FVflux.py:21-28)."""
from .FVInterpolation import Interplot


class FV_flux(Interplot):
    def __init__(self):
        super().__init__()
