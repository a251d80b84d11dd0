"""frad_python_b200 -- B200-native hot path of the FrAD codec (DCT, masking/quantisation, CRC, RS).

The numeric work runs in hand-written sm_100a kernels reached through the C ABI in
``include/fradb200.h`` (``frad_python_b200/lib/libfradb200.so``).  There is no CPU fallback.
"""
from ._abi import Context, FradError, default_context, PREC_REF, PREC_F32, PREC_F64  # noqa: F401
