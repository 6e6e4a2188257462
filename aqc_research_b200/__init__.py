"""
aqc_research_b200 -- B200-native (sm_100a) objective-and-gradient hot path of
qiskit-community/aqc-research behind the reference's own Python API.

Only the hot path lives here: circuit description classes, the
state-vector / matrix / MPS numeric core (CUDA, reached through the C-ABI of
``include/aqc_b200.h``) and the objective classes the SciPy L-BFGS loop calls.
There is no CPU fallback: compute calls raise if the CUDA library or a GPU is missing.
"""

__version__ = "0.1.0"
