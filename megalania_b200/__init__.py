"""B200-native annealing hot path of the Megalania LZMA optimiser.

Only what the path needs lives here:
  csrc/   CUDA kernels (sm_100a) + the C ABI (include/megalania_cuda.h)
  host/   C host: the drop-in `megalania` CLI over the reference's plug-in interfaces
  api.py  ctypes mirror of the host interface for tests and benchmarks
"""
from .api import (Annealer, Context, MegalaniaError, PACKET_DTYPE, literal_slab, load_library,  # noqa: F401
                  anneal_oneshot, LITERAL, MATCH, SHORT_REP, LONG_REP, SCHEDULE_REFERENCE, SCHEDULE_TEMPERATURE, CONTINUE_EVALS, NO_OWNER)
