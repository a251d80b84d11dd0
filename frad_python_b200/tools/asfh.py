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

    def fill_buffer(self, buffer: bytes, target_size: int):
        if len(self.buffer) < target_size:
            need = target_size - len(self.buffer)
            self.buffer += buffer[:need]
            buffer = buffer[need:]
            if len(self.buffer) < target_size:
                return False, buffer
        self.header_bytes = target_size
        return True, buffer

    def read(self, buffer: bytes):
        ok, buffer = self.fill_buffer(buffer, 9)
        if not ok:
            return "Incomplete", buffer
        self.frmbytes = struct.unpack(">I", self.buffer[4:8])[0]
        self.profile, self.ecc, self.endian, self.bit_depth_index = decode_pfb(self.buffer[8:9])
        if self.profile in COMPACT:
            ok, buffer = self.fill_buffer(buffer, 12)
            if not ok:
                return "Incomplete", buffer
            self.channels, self.srate, self.fsize, flush = decode_css_prf1(self.buffer[9:11])
            if flush:
                return "ForceFlush", buffer
            self.overlap_ratio = self.buffer[11]
            if self.overlap_ratio != 0:
                self.overlap_ratio += 1
            if self.ecc:
                ok, buffer = self.fill_buffer(buffer, 16)
                if not ok:
                    return "Incomplete", buffer
                self.ecc_dsize, self.ecc_codesize = self.buffer[12], self.buffer[13]
                self.crc = self.buffer[14:16]
        else:
            ok, buffer = self.fill_buffer(buffer, 32)
            if not ok:
                return "Incomplete", buffer
            self.channels = self.buffer[9] + 1
            self.ecc_dsize, self.ecc_codesize = self.buffer[10], self.buffer[11]
            self.srate = struct.unpack(">I", self.buffer[12:16])[0]
            self.fsize = struct.unpack(">I", self.buffer[24:28])[0]
            self.crc = self.buffer[28:32]
        if self.frmbytes == 0xFFFFFFFF:
            ok, buffer = self.fill_buffer(buffer, self.header_bytes + 8)
            if not ok:
                return "Incomplete", buffer
            self.frmbytes = struct.unpack(">Q", self.buffer[-8:])[0]
        self.all_set = True
        return "Complete", buffer

    def clear(self):
        self.all_set = False
        self.buffer = b""
