"""Reed-Solomon over GF(2^8) -- CPU oracle (test infrastructure, never shipped).

Restates the code the reference reaches through ``reedsolo.RSCodec`` at
/root/reference/src/libfrad/tools/ecc.py:8,10,16,21 (``reedsolo`` itself is a
third-party, un-vendored, unpinned dependency that is absent from this image --
SURVEY.md F4 -- so this file restates its *published* algorithm):

  * field GF(2^8), primitive polynomial 0x11D, generator alpha = 2
  * first consecutive root fcr = 0:  g(x) = prod_{i=0}^{nsym-1} (x - alpha^i)
  * systematic encoding: parity = msg(x) * x^nsym mod g(x), appended after the data
  * shortened codes: simply fewer data bytes
  * decoding: syndromes S_j = r(alpha^j), Berlekamp-Massey, Chien search limited to
    the positions of the actual (possibly shortened) word, Forney, then a syndrome
    re-check.  Failure (``ReedSolomonError``) when the locator degree exceeds
    nsym/2, when the number of roots found differs from the degree, or when the
    re-check is non-zero.

Any bounded-distance decoder gives identical bytes: success iff a codeword of the
shortened code lies within floor(nsym/2) byte errors of the received word.

PARITY UNPINNED by the reference (it has no RS vectors).  Pinned here by public material
(tests/test_oracle_golden.py): two known-answer vectors -- the QR-code example
``40 d2 75 47 76 17 32 06 27 26 96 c6 c6 96 70 ec`` -> parity
``bc 2a 90 13 6b af ef fd 4b e0`` and ``RSCodec(10).encode(b'hello world')`` --, the
QR version 1-M "HELLO WORLD" codewords, the generator polynomials of the QR specification's
tables (same field, alpha and fcr) against which the encoder is checked by an independent
schoolbook division, and the decoder's bounded-distance limits.  The GPU encoder is checked
by the code's definition as well (every word vanishes at alpha^0 .. alpha^(nsym-1)):
tests/test_gpu_kernels.py::test_rs_encode_is_the_code_by_definition.
"""
from __future__ import annotations

import numpy as np

PRIM = 0x11D


class ReedSolomonError(Exception):
    """Raised when a word cannot be decoded (mirrors reedsolo.ReedSolomonError)."""


def _build_tables():
    exp = np.zeros(512, dtype=np.uint8)
    log = np.zeros(256, dtype=np.int32)
    x = 1
    for i in range(255):
        exp[i] = x
        log[x] = i
        x <<= 1
        if x & 0x100:
            x ^= PRIM
    exp[255:510] = exp[0:255]
    return exp, log


GF_EXP, GF_LOG = _build_tables()


def gf_mul(a: int, b: int) -> int:
    if a == 0 or b == 0:
        return 0
    return int(GF_EXP[int(GF_LOG[a]) + int(GF_LOG[b])])


def gf_div(a: int, b: int) -> int:
    if b == 0:
        raise ZeroDivisionError
    if a == 0:
        return 0
    return int(GF_EXP[(int(GF_LOG[a]) - int(GF_LOG[b])) % 255])


def gf_inv(a: int) -> int:
    return int(GF_EXP[255 - int(GF_LOG[a])])


def gf_pow_alpha(e: int) -> int:
    return int(GF_EXP[e % 255])


def generator_poly(nsym: int) -> list[int]:
    """g(x), highest degree first, g[0] == 1."""
    g = [1]
    for i in range(nsym):
        root = gf_pow_alpha(i)
        nxt = g + [0]
        for j, c in enumerate(g):
            nxt[j + 1] ^= gf_mul(c, root)
        g = nxt
    return g


def _gf_mul_vec(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Element-wise GF product of two uint8 arrays (broadcasting)."""
    a = np.asarray(a, dtype=np.uint8)
    b = np.asarray(b, dtype=np.uint8)
    out = GF_EXP[GF_LOG[a] + GF_LOG[b]]
    return np.where((a == 0) | (b == 0), np.uint8(0), out).astype(np.uint8)


def encode_parity_batch(msgs: np.ndarray, nsym: int) -> np.ndarray:
    """Parity bytes for a batch of equal-length messages.  msgs: uint8 [B, k] -> [B, nsym]."""
    msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
    nb, k = msgs.shape
    par = np.zeros((nb, nsym), dtype=np.uint8)
    if nsym == 0:
        return par
    g = np.array(generator_poly(nsym)[1:], dtype=np.uint8)  # nsym coefficients
    for i in range(k):
        fb = msgs[:, i] ^ par[:, 0]
        par[:, :-1] = par[:, 1:]
        par[:, -1] = 0
        par ^= _gf_mul_vec(fb[:, None], g[None, :])
    return par


def encode_msg(msg: bytes, nsym: int) -> bytes:
    """One systematic codeword: msg || parity."""
    arr = np.frombuffer(bytes(msg), dtype=np.uint8)[None, :]
    return bytes(msg) + encode_parity_batch(arr, nsym)[0].tobytes()


def syndromes_batch(words: np.ndarray, nsym: int) -> np.ndarray:
    """S_j = r(alpha^j), j = 0..nsym-1, Horner over the word.  words: uint8 [B, n]."""
    words = np.ascontiguousarray(words, dtype=np.uint8)
    nb, n = words.shape
    roots = GF_EXP[np.arange(nsym) % 255].astype(np.uint8)
    s = np.zeros((nb, nsym), dtype=np.uint8)
    for i in range(n):
        s = _gf_mul_vec(s, roots[None, :]) ^ words[:, i][:, None]
    return s


def _poly_eval_low_first(p: list[int], x: int) -> int:
    """p[0] + p[1] x + ...  (coefficients lowest degree first)."""
    y = 0
    for c in reversed(p):
        y = gf_mul(y, x) ^ c
    return y


def correct_word(word: bytes, nsym: int, synd=None) -> tuple[bytes, int]:
    """Decode one (possibly shortened) codeword.  Returns (data bytes, n_corrected).

    Raises ReedSolomonError when no codeword lies within nsym//2 errors.
    """
    n = len(word)
    if n > 255:
        raise ValueError("word longer than 255 bytes")
    r = bytearray(word)
    if synd is None:
        synd = syndromes_batch(np.frombuffer(bytes(r), dtype=np.uint8)[None, :], nsym)[0]
    S = [int(v) for v in synd]
    if not any(S):
        return bytes(r[: n - nsym]), 0

    # Berlekamp-Massey; polynomials lowest degree first
    lam = [1]
    prev = [1]
    L, m, b = 0, 1, 1
    for it in range(nsym):
        d = S[it]
        for i in range(1, L + 1):
            if i < len(lam):
                d ^= gf_mul(lam[i], S[it - i])
        if d == 0:
            m += 1
            continue
        coef = gf_div(d, b)
        upd = lam + [0] * max(0, len(prev) + m - len(lam))
        for i, c in enumerate(prev):
            upd[i + m] ^= gf_mul(coef, c)
        if 2 * L <= it:
            prev, lam = lam, upd
            L = it + 1 - L
            b = d
            m = 1
        else:
            lam = upd
            m += 1
    while len(lam) > 1 and lam[-1] == 0:
        lam.pop()
    deg = len(lam) - 1
    if 2 * L > nsym or deg != L:
        raise ReedSolomonError("too many errors")

    # Chien search over the positions that exist in this word
    positions = []
    for i in range(n):
        p = n - 1 - i  # power of x carried by byte i
        xinv = gf_pow_alpha(255 - (p % 255))
        if _poly_eval_low_first(lam, xinv) == 0:
            positions.append(i)
    if len(positions) != deg:
        raise ReedSolomonError("locator roots do not match its degree")

    # Forney: e = X * Omega(X^-1) / Lambda'(X^-1),  Omega = S(x) Lambda(x) mod x^nsym
    omega = [0] * nsym
    for i in range(nsym):
        acc = 0
        for j in range(min(i, deg) + 1):
            acc ^= gf_mul(lam[j], S[i - j])
        omega[i] = acc
    lam_d = [lam[i] if (i & 1) else 0 for i in range(1, deg + 1)]  # derivative coefficients
    for i in positions:
        p = n - 1 - i
        X = gf_pow_alpha(p)
        xinv = gf_pow_alpha(255 - (p % 255))
        den = _poly_eval_low_first(lam_d, xinv)
        if den == 0:
            raise ReedSolomonError("locator derivative vanished")
        num = gf_mul(X, _poly_eval_low_first(omega, xinv))
        r[i] ^= gf_div(num, den)

    chk = syndromes_batch(np.frombuffer(bytes(r), dtype=np.uint8)[None, :], nsym)[0]
    if chk.any():
        raise ReedSolomonError("could not correct")
    return bytes(r[: n - nsym]), deg


class RSCodec:
    """Tiny subset of reedsolo.RSCodec: what tools/ecc.py uses (encode, decode()[0])."""

    def __init__(self, nsym: int = 10, nsize: int = 255):
        if nsize > 255 or nsym > nsize:
            raise ValueError("bad code size")
        self.nsym = nsym
        self.nsize = nsize

    def encode(self, data) -> bytearray:
        data = bytes(data)
        k = self.nsize - self.nsym
        out = bytearray()
        for i in range(0, len(data), k):
            out += encode_msg(data[i:i + k], self.nsym)
        return out

    def decode(self, data):
        data = bytes(data)
        dec = bytearray()
        full = bytearray()
        for i in range(0, len(data), self.nsize):
            chunk = data[i:i + self.nsize]
            msg, _ = correct_word(chunk, self.nsym)
            dec += msg
            full += msg + encode_msg(msg, self.nsym)[len(msg):]
        return dec, full, bytearray()
