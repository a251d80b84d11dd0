"""Audio Stream Frame Header (mirror of libfrad/tools/asfh.py): same fields, same bytes.

``write`` accepts the payload checksum computed on the GPU by the batched codec objects; when it is
not given the checksum is computed through the GPU CRC seam with batch size 1."""
import struct

from .. import common
from ..fourier.profiles import COMPACT, compact


def encode_pfb(profile: int, isecc: bool, little_endian: bool, bits: int) -> bytes:
    return bytes([(profile << 5) | (int(bool(isecc)) << 4) | (int(bool(little_endian)) << 3) | bits])


def decode_pfb(pfb: bytes):
    v = pfb[0]
    return v >> 5, bool((v >> 4) & 1), bool((v >> 3) & 1), v & 0b111


def encode_css_prf1(channels: int, srate: int, fsize: int, force_flush: bool) -> bytes:
    word = ((channels - 1) << 10) | (compact.get_srate_index(srate) << 6) | (compact.get_samples_index(fsize) << 1)
    return struct.pack(">H", word | int(bool(force_flush)))


def decode_css_prf1(css: bytes):
    word = struct.unpack(">H", css)[0]
    return (word >> 10) + 1, compact.SRATES[(word >> 6) & 0b1111], compact.SAMPLES[(word >> 1) & 0b11111], word & 1


class ASFH:
    def __init__(self):
        self.frmbytes = 0
        self.buffer = b""
        self.all_set = False
        self.header_bytes = 0
        self.endian, self.bit_depth_index = False, 0
        self.channels, self.srate, self.fsize = 0, 0, 0
        self.ecc, self.ecc_dsize, self.ecc_codesize = False, 0, 0
        self.profile, self.overlap_ratio = 0, 0
        self.crc = b""

    def criteq(self, other: "ASFH") -> bool:
        return self.channels == other.channels and self.srate == other.srate

    def write(self, frad: bytes, crc=None) -> bytes:
        head = common.FRM_SIGN + struct.pack(">I", len(frad))
        head += encode_pfb(self.profile, self.ecc, self.endian, self.bit_depth_index)
        if self.profile in COMPACT:
            head += encode_css_prf1(self.channels, self.srate, self.fsize, False)
            head += bytes([max(self.overlap_ratio - 1, 0)])
            if self.ecc:
                if crc is None:
                    crc = common.crc16_ansi(frad)
                head += bytes([self.ecc_dsize, self.ecc_codesize]) + int(crc).to_bytes(2, "big")
        else:
            if crc is None:
                crc = common.crc32(frad)
            head += bytes([self.channels - 1, self.ecc_dsize, self.ecc_codesize])
            head += struct.pack(">I", self.srate) + b"\x00" * 8 + struct.pack(">I", self.fsize)
            head += int(crc).to_bytes(4, "big")
        return head + frad

    def force_flush(self) -> bytes:
        if self.profile not in COMPACT:
            return b""
        head = common.FRM_SIGN + b"\x00" * 4
        head += encode_pfb(self.profile, self.ecc, self.endian, self.bit_depth_index)
        head += encode_css_prf1(max(self.channels, 1), self.srate, self.fsize, True)
        return head + b"\x00"

    # ---- parsing: length first, fields second --------------------------------------------------------------------------
    # The header's length follows from its own prefix: 9 bytes tell the profile and the ECC flag, a compact header's 12 tell a
    # force-flush marker (nothing follows), and an all-ones size field appends a 64-bit size.  `read` tops the buffer up to what
    # the prefix asks for, asks again, and decodes every field in one pass once the length is final.
    _WIDE_SIZE = b"\xff\xff\xff\xff"

    def _span(self):
        """(bytes this header needs as far as the buffered prefix can tell, True when that is the whole header)"""
        have = self.buffer
        if len(have) < 9:
            return 9, False
        if (have[8] >> 5) in COMPACT:
            if len(have) < 12:
                return 12, False
            if have[10] & 1:                                   # force-flush bit of the css word
                return 12, True
            base = 16 if (have[8] >> 4) & 1 else 12
        else:
            base = 32
        return base + (8 if have[4:8] == self._WIDE_SIZE else 0), True

    def fill_buffer(self, buffer: bytes, target_size: int):
        """Tops the header buffer up to target_size bytes from `buffer`; returns (reached, rest of `buffer`)."""
        take = max(target_size - len(self.buffer), 0)
        self.buffer += buffer[:take]
        if len(self.buffer) < target_size:
            return False, buffer[take:]
        self.header_bytes = target_size
        return True, buffer[take:]

    def read(self, buffer: bytes):
        final = False
        while not final:
            want, final = self._span()
            reached, buffer = self.fill_buffer(buffer, want)
            if not reached:
                return "Incomplete", buffer
            final = final and self._span() == (want, True)     # a longer prefix may ask for more
        head = self.buffer
        self.profile, self.ecc, self.endian, self.bit_depth_index = decode_pfb(head[8:9])
        self.frmbytes = struct.unpack_from(">I", head, 4)[0]
        if self.profile in COMPACT:
            self.channels, self.srate, self.fsize, flush = decode_css_prf1(head[9:11])
            if flush:
                return "ForceFlush", buffer
            self.overlap_ratio = head[11] + 1 if head[11] else 0
            if self.ecc:
                self.ecc_dsize, self.ecc_codesize, self.crc = head[12], head[13], head[14:16]
        else:
            chm1, self.ecc_dsize, self.ecc_codesize, self.srate = struct.unpack_from(">BBBI", head, 9)
            self.channels = chm1 + 1
            self.fsize = struct.unpack_from(">I", head, 24)[0]
            self.crc = head[28:32]
        if head[4:8] == self._WIDE_SIZE:
            self.frmbytes = struct.unpack_from(">Q", head, self.header_bytes - 8)[0]
        self.all_set = True
        return "Complete", buffer

    def clear(self):
        self.all_set = False
        self.buffer = b""
