"""contractn_b200 -- B200-native (sm_100a) contraction engine for ContracTN.

Drop-in for ONE path of jemisjoky/ContracTN: ``TN.contract()`` -> the stabilised
pairwise contraction executor (reference ``contractn/einsum.py:190-393``).
Same user API (``TN``, ``Node``, ``Edge``, ``contract``), device execution.
"""
from .tn import TN, Node, Edge
from .engine import contract, core_contract, install, Plan, EngineError, device_available

__all__ = ["TN", "Node", "Edge", "contract", "core_contract", "install", "Plan",
           "EngineError", "device_available"]
__version__ = "0.1.0"
