"""Drop-in installation into an importable reference ``libfrad`` (level 1).

``patch_reference(libfrad)`` rebinds the module-level seams the reference resolves at call time --
``fourier.profileN.analogue/digital``, ``tools.ecc.encode/decode``, ``asfh.crc32``, ``common.crc16_ansi``
-- to the GPU implementations, so the untouched ``libfrad.Encoder/Decoder/Repairer`` run on the B200.
``unpatch_reference`` restores the originals."""
from . import common, fourier, tools

_saved = {}


def patch_reference(libfrad):
    ref_fourier = libfrad.fourier
    pairs = [
        (ref_fourier.profile0, "analogue", fourier.profile0.analogue),
        (ref_fourier.profile0, "digital", fourier.profile0.digital),
        (ref_fourier.profile1, "analogue", fourier.profile1.analogue),
        (ref_fourier.profile1, "digital", fourier.profile1.digital),
        (ref_fourier.profile2, "analogue", fourier.profile2.analogue),
        (ref_fourier.profile2, "digital", fourier.profile2.digital),
        (ref_fourier.profile4, "analogue", fourier.profile4.analogue),
        (ref_fourier.profile4, "digital", fourier.profile4.digital),
        (libfrad.tools.ecc, "encode", tools.ecc.encode),
        (libfrad.tools.ecc, "decode", tools.ecc.decode),
        (libfrad.tools.asfh, "crc32", common.crc32),
        (libfrad.tools.asfh, "crc16_ansi", common.crc16_ansi),
        (libfrad.common, "crc16_ansi", common.crc16_ansi),
    ]
    for mod, name, fn in pairs:
        _saved.setdefault((mod, name), getattr(mod, name))
        setattr(mod, name, fn)
    if 2 not in ref_fourier.AVAILABLE:
        ref_fourier.AVAILABLE.append(2)


def unpatch_reference():
    for (mod, name), fn in _saved.items():
        setattr(mod, name, fn)
    _saved.clear()
