"""frad_python_b200 -- B200-native hot path of the FrAD codec (DCT, masking/quantisation, CRC, RS).

The numeric work runs in hand-written sm_100a kernels reached through the C ABI in
``include/fradb200.h`` (``frad_python_b200/lib/libfradb200.so``).  There is no CPU fallback.

Public surface mirrors ``libfrad`` (reference src/libfrad/__init__.py:1-9): ``Encoder``, ``Decoder``,
``Repairer``, ``ASFH``, ``fourier.profile{0,1,2}.analogue/digital``, ``tools.ecc.encode/decode`` and the
CRC seams in ``common``.  ``install.patch_reference()`` rebinds the reference's own seams to these.
"""
from ._abi import Context, FradError, default_context, PREC_REF, PREC_F32, PREC_F64  # noqa: F401
from . import common, fourier, tools  # noqa: F401
from .backend import ff_format_to_numpy_type, to_f64, from_f64  # noqa: F401
from .fourier import AVAILABLE, BIT_DEPTHS, SEGMAX, profiles  # noqa: F401
from .tools.asfh import ASFH  # noqa: F401
from .encoder import Encoder, EncodeResult  # noqa: F401
from .decoder import Decoder, DecodeResult  # noqa: F401
from .repairer import Repairer  # noqa: F401
