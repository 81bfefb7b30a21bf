"""B200-native batched step engine for DR-MPC's hot path (ORCA human step + path-tracking MPC +
the env step around them). Host side: Python mirror of the reference's interfaces over the C ABI of
libdre.so (include/dre.h). No CPU fallback: importing works anywhere, calling needs the CUDA library."""
from .config import EngineConfig
from .orca import BatchedORCA
from .mpc import BatchedMPC
from . import paths
from . import parallel
from .replay import BatchedReplayBuffer, DeviceReplayBuffer
try:
    from .env import BatchedHAAndPTEnv
except ImportError:   # env module not present in minimal builds
    pass

from . import compat

__all__ = ["compat", "EngineConfig", "BatchedORCA", "BatchedMPC", "BatchedHAAndPTEnv", "BatchedReplayBuffer", "DeviceReplayBuffer", "paths", "parallel"]
