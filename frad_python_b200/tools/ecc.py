"""Reed-Solomon ECC seam (mirror of libfrad/tools/ecc.py:6-25) running on the GPU."""
from .._abi import default_context


def encode(data: bytes, ecc_dsize: int, ecc_codesize: int, ctx=None) -> bytes:
    """Every ``ecc_dsize``-byte chunk of data becomes chunk || ``ecc_codesize`` parity bytes."""
    if not data:
        return b""
    if ecc_codesize <= 0:
        return bytes(data)
    return (ctx or default_context()).rs_encode_many([bytes(data)], ecc_dsize, ecc_codesize)[0]


def decode(data: bytes, ecc_dsize: int, ecc_codesize: int, repair: bool, ctx=None) -> bytes:
    """Strip the parity; with ``repair`` correct up to codesize/2 byte errors per codeword and
    zero-fill codewords that cannot be corrected (ecc.py:19-23)."""
    if not data:
        return b""
    if ecc_codesize <= 0:
        return bytes(data)
    return (ctx or default_context()).rs_decode_many([bytes(data)], ecc_dsize, ecc_codesize, repair)[0][0]
