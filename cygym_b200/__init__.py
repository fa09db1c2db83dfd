"""cygym_b200: B200-native batched implementation of CyGym's CyberDefenseSimulator step path."""
from .network import Network, synthetic_network, network_from_golden  # noqa: F401


def __getattr__(name):
    # torch / CUDA-dependent members load lazily so that the host-side helpers import anywhere
    if name in ("VectorCyberDefenseEnv", "ActionBatch"):
        from . import vector_env
        return getattr(vector_env, name)
    if name in ("Volt_Typhoon_CyberDefenseEnv", "CyberDefenseEnv"):
        from . import volt_typhoon_env
        return getattr(volt_typhoon_env, name)
    raise AttributeError(name)
