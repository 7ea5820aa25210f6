"""artensor_b200 -- B200-native (sm_100a) numerical contraction executor for artensor schemes.

Drop-in for the hot path of Fanerst/artensor: `tensor_contraction`, `tensor_contraction_sparse`
(artensor/contraction.py) and the slice loop of `TensorNetworkSimulation.contraction`
(artensor/simulation.py), plus the scheme compilers `contraction_scheme` /
`contraction_scheme_sparse` that turn the reference's contraction tree into executor steps.  The
circuit builder and the order finder stay the reference's own code.  No CPU fallback: the CUDA library must be built (see __graft_entry__.build).
"""
from .plan import SchemeError, SchemeParser
from .backend import ContractionPlan, PlanOptions
from .contraction import tensor_contraction, tensor_contraction_sparse, get_plan
from .simulation import (
    TensorNetworkSimulation,
    tensor_network_contraction,
    quantum_circuit_simulation,
    check_bitstrings,
    partition_slices,
)
from .cases import load_case, save_case, Case
from .scheme import contraction_scheme, contraction_scheme_sparse

__all__ = [
    "tensor_contraction", "tensor_contraction_sparse", "TensorNetworkSimulation",
    "tensor_network_contraction", "quantum_circuit_simulation", "ContractionPlan", "PlanOptions",
    "SchemeError", "load_case", "save_case", "Case", "contraction_scheme", "contraction_scheme_sparse",
]
