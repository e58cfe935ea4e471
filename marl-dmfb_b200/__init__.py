"""marl-dmfb_b200 — B200-native batched DMFB / MEDA environment step.

Drop-in for the environment step of jesselasse/MARL-DMFB (env/DMFB/dmfb.py, env/MEDA/meda.py):
N independent chips live in HBM as struct-of-arrays and `step` runs as hand-written sm_100a CUDA
kernels behind a C ABI (include/dmfb_b200.h, marl-dmfb_b200/lib/libdmfb_b200.so).

The directory name has a hyphen; import it with
    import importlib; dmfb_b200 = importlib.import_module("marl-dmfb_b200")
or through the alias module at the repo root:  `import marl_dmfb_b200`.
"""
from . import _native  # noqa: F401
from .build import build  # noqa: F401
from .dmfb import BatchedDMFB, DMFBenv, DMFBenv_v0_1  # noqa: F401
from .host import HostDMFB, HostMEDA, gpu_cpu_affinity, pin_to_gpu_numa  # noqa: F401
from .sharding import shard_range  # noqa: F401
from .marl import (CRNN, RNN, BatchedAgents, BatchedRolloutWorker, EpisodeBatch, PhaseTimer, QMixNet,  # noqa: F401
                   QMIXLearner, ReplayBufferGPU, VDNLearner, allreduce_gradients)

try:  # MEDA kernels are part of the same library
    from .meda import BatchedMEDA, MEDAEnv, MEDAEnv_v0_1, MEDAEnv_v0_2  # noqa: F401
except ImportError:  # pragma: no cover - only while the package is being bootstrapped
    pass

__all__ = ["BatchedDMFB", "DMFBenv", "DMFBenv_v0_1", "HostDMFB", "HostMEDA", "BatchedMEDA", "MEDAEnv", "MEDAEnv_v0_1", "MEDAEnv_v0_2", "shard_range", "build"]
