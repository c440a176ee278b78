"""cirkit_b200: B200-native (sm_100a) forward/backward evaluator for cirkit's compiled circuits.

Host side in Python, device side in hand-written CUDA behind the C ABI declared in
`include/cirkit_b200.h` (built into `cirkit_b200/lib/libcirkit_b200.so`, loaded with ctypes).
"""

from .plan import CircuitPlan, LeafSpec, ParamSpec, PlanLayout, StepSpec, build_layout

__all__ = [
    "CircuitPlan",
    "LeafSpec",
    "ParamSpec",
    "PlanLayout",
    "StepSpec",
    "build_layout",
    "B200Circuit",
    "IntegrateQuery",
    "SamplingQuery",
    "PlanRuntime",
    "accelerate",
    "plan_from_torch",
    "register_backend",
]


def __getattr__(name):
    # runtime pieces are imported on first use so that plan tooling works on boxes without
    # the CUDA library; using them without it fails loudly (cirkit_b200._lib).
    if name in ("PlanRuntime",):
        from . import runtime

        return getattr(runtime, name)
    if name in ("B200Circuit",):
        from . import circuit

        return getattr(circuit, name)
    if name in ("IntegrateQuery", "SamplingQuery"):
        from . import queries

        return getattr(queries, name)
    if name in ("accelerate", "plan_from_torch", "register_backend", "UnsupportedCircuitError"):
        from . import adapter

        return getattr(adapter, name)
    raise AttributeError(name)
