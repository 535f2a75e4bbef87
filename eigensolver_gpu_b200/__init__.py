"""eigensolver_gpu_b200 -- B200-native drop-in for the dsygvdx_gpu / zhegvdx_gpu path of NVIDIA/Eigensolver_gpu.

The product is the C-ABI library lib/libeigb200.so (include/eigb200.h), hand-written CUDA for sm_100a.
This Python package is the host-side mirror of the reference's Fortran interface (same names and argument
meaning) used by the tests and the bench; PyTorch only supplies device memory, streams and
torch.distributed.  There is no CPU fallback: every compute call fails loudly without the CUDA library.
"""
from ._lib import Eigb200Error, load  # noqa: F401

__all__ = ["load", "Eigb200Error"]
