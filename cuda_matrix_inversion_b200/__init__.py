"""cuda_matrix_inversion_b200 -- B200-native (sm_100a) batched dense inversion and fused GP
mean/variance behind the C API of akuendig/cuda-matrix-inversion.

The product is ``lib/libinvgpu.so`` (hand-written CUDA + a C ABI, see ``include/*.h``); this
package is only its ctypes mirror for tests and measurement.  Importing it without the built
library raises ImportError -- there is no CPU fallback.
"""
from . import api  # noqa: F401
from ._lib import LIB_PATH, lib  # noqa: F401
