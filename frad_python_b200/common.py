"""Stream signatures and CRC seams (mirror of libfrad/common.py:1-10; zlib.crc32 of tools/asfh.py:4).

The per-call functions run on the GPU with batch size 1 (parity surface); the batched codec objects
use ``Context.crc_many`` directly."""
from ._abi import default_context

SIGNATURE = b"fRad"
FRM_SIGN = b"\xff\xd0\xd2\x98"


def crc16_ansi(data: bytes, ctx=None) -> int:
    """CRC-16/ARC of ``data`` (libfrad/common.py:6-10)."""
    return int((ctx or default_context()).crc_many([bytes(data)], False)[0])


def crc32(data: bytes, ctx=None) -> int:
    """CRC-32/ISO-HDLC of ``data`` (what ``zlib.crc32`` returns in tools/asfh.py:69)."""
    return int((ctx or default_context()).crc_many([bytes(data)], True)[0])
