"""
Loader for the UNMODIFIED reference (qiskit-community/aqc-research) from
``/root/reference`` -- build-container only.

The reference is pure Python but (a) uses the alias ``numpy.cfloat`` that NumPy 2
removed and (b) imports ``qiskit`` / ``qiskit_aer`` at module level.  We inject the
alias and register stub modules; every pure-NumPy code path
of the reference then runs unmodified.  Nothing of the reference is copied.

This module is used ONLY by ``tests/golden/make_golden.py`` (fixture generation) and
by tests that are skipped when ``/root/reference`` is absent (the GPU box).
"""

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AQC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "aqc_research"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    for key, val in attrs.items():
        setattr(mod, key, val)
    sys.modules[name] = mod
    return mod


class _Missing:  # placeholder for Qiskit classes; instantiation is an error
    def __init__(self, *_, **__):
        raise RuntimeError("qiskit is not installed: stubbed class was instantiated")


def load_reference():
    """Imports the reference package and returns a namespace of its hot-path modules."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    import numpy as np

    if not hasattr(np, "cfloat"):
        np.cfloat = np.complex128  # alias removed in NumPy 2 (reference: checking.py:35)

    if "qiskit" not in sys.modules:
        mk = lambda n: type(n, (_Missing,), {})  # noqa: E731
        _stub("qiskit", QuantumCircuit=mk("QuantumCircuit"))
        _stub("qiskit.quantum_info", Operator=mk("Operator"), Statevector=mk("Statevector"))
        _stub("qiskit.circuit", Parameter=mk("Parameter"))
        _stub("qiskit.circuit.library", QFT=mk("QFT"))
        _stub("qiskit.algorithms")
        _stub(
            "qiskit.algorithms.optimizers",
            L_BFGS_B=mk("L_BFGS_B"),
            ADAM=mk("ADAM"),
            COBYLA=mk("COBYLA"),
            BOBYQA=mk("BOBYQA"),
        )
        _stub("qiskit.algorithms.optimizers.optimizer", OptimizerResult=mk("OptimizerResult"))
        _stub("qiskit_aer", AerSimulator=mk("AerSimulator"))

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    ns = types.SimpleNamespace()
    import aqc_research.parametric_circuit as pc
    import aqc_research.circuit_structures as cs
    import aqc_research.elementary_operations as eo
    import aqc_research.core_operations as cop
    import aqc_research.core_op_matrix as cpm
    import aqc_research.utils as utils

    ns.pc, ns.cs, ns.eo, ns.cop, ns.cpm, ns.utils = pc, cs, eo, cop, cpm, utils
    import aqc_research.mps_operations as mpsop
    import aqc_research.model_sp_lhs.objective_lhs_sur_max as sur_max
    import aqc_research.model_sketching.sk_core as sk_core
    import aqc_research.circuit_transform as ctr
    import aqc_research.model_sp_lhs.trotter.trotter as trotter

    ns.mpsop, ns.sur_max, ns.sk_core, ns.ctr, ns.trotter = mpsop, sur_max, sk_core, ctr, trotter
    return ns
