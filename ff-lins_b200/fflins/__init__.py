"""fflins — host-side Python mirror of the FF-LINS LiDAR hot-path interfaces over the B200 C-ABI
(libfflins_b200.so). `api` is the one-to-one ctypes binding, `pipeline` the call sequence of
LINS::runFusion's LiDAR branch (ff_lins/ff_lins/ff_lins.cc:236-262), `synth` the seeded synthetic
sequences. Nothing here computes on the CPU: without the CUDA library the import of a Context
fails."""
from . import api, synth  # noqa: F401
from .api import Context, default_config  # noqa: F401
