"""Oracle-backed stand-in for frad_python_b200.Context, for CPU tests of the HOST logic only (header walking, batching,
stream order).  Test infrastructure: nothing in the product imports it."""
import numpy as np

from oracle import frad_oracle as O


class OracleCtx:
    """Implements the byte-side calls the batched Repairer / Decoder make (crc_many, rs_encode_many, rs_decode_many)."""

    def crc_many(self, chunks, bits32: bool):
        f = O.crc32 if bits32 else O.crc16_arc
        return np.array([f(bytes(c)) for c in chunks], np.uint32)

    def rs_encode_many(self, chunks, dsize, csize):
        return [O.ecc_encode(bytes(c), dsize, csize) for c in chunks]

    def rs_decode_many(self, chunks, dsize, csize, repair):
        out = [O.ecc_decode(bytes(c), dsize, csize, bool(repair)) for c in chunks]
        z = np.zeros(len(chunks), np.int32)
        return out, z, z.copy()
