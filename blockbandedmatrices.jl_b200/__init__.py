"""bbm_b200 -- B200-native block-banded ``mul!`` / ``ldiv!`` (the hot path of
BlockBandedMatrices.jl) behind the reference's container / lazy-op interface.

The arithmetic lives in ``lib/libbbm_b200.so`` (hand-written CUDA for sm_100a, C ABI declared in
``include/bbm_b200.h``); this package is the host-side mirror of the reference interface used by
the tests and the benchmark (Julia is not available in the build image; INTEGRATION.md shows the
equivalent Julia package extension).
"""
from ._lib import BandError, BBMError, DimensionMismatch, build_library, LIB_PATH, SIGNATURES  # noqa: F401
from .broadcast import (Bidiagonal, Diagonal, SymTridiagonal, Tridiagonal, UniformScaling, accommodating_bandwidths, add, add_diags_,  # noqa: F401
                        axpby_, scaled_add, similar_sum, sub)
from .device import DeviceArray, Handle, PinnedArray  # noqa: F401
from .linalg import PreparedTriangular, add_bandwidths, axpy_, ldiv, ldiv_, prepare_ldiv, lmul_, lmul_adj_, mul, ql_, qr_, mul_, muladd_, sparse  # noqa: F401
from .matrices import (Adjoint, BandedBlockBandedMatrix, BlockBandedMatrix, BlockBandedQL, BlockBandedQR, BlockSkylineMatrix, LowerTriangular,  # noqa: F401
                       UnitLowerTriangular, UnitUpperTriangular, UpperTriangular)

__all__ = ["BandedBlockBandedMatrix", "BlockSkylineMatrix", "BlockBandedMatrix", "UpperTriangular", "LowerTriangular",
           "UnitUpperTriangular", "UnitLowerTriangular", "Adjoint", "BlockBandedQR", "BlockBandedQL", "lmul_adj_", "qr_", "ql_", "sparse", "ldiv", "ldiv_", "add_bandwidths", "axpy_", "lmul_", "Handle", "DeviceArray", "PinnedArray", "mul", "mul_", "muladd_",
           "DimensionMismatch", "BandError", "BBMError", "build_library"]
