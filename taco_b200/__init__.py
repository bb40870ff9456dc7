"""taco_b200 -- B200-native (sm_100a) GPU compute path for tensor-compiler/taco behind the taco_tensor_t C ABI.

The product is libtaco_b200.so (include/taco_b200.h, sources in taco_b200/csrc).  This package is the thin
host-side mirror of the reference's tensor/kernel interface used by tests and bench.py; it contains no compute
fallback -- importing it fails if the library has not been built.
"""
from . import formats  # noqa: F401  (pure-python helpers, no GPU needed)
from ._lib import LIB_PATH, TacoError  # noqa: F401
from .tensor import (BCSR, CSF3, CSR, DCSR, Dense, Format, Kernel, Sparse, Tensor, compile, compressed, dense, launch_count,  # noqa: F401
                     makeBCSR, makeCSF3, makeCSR, makeDCSR, makeDense, pack, partition_pos, read, pinned_empty, pinned_free, set_result_multicast, set_result_peers, set_result_space,
                     synchronize, use_torch_stream)
