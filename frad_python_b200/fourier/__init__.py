"""Profile registry (mirror of libfrad/fourier/__init__.py).  Profile 2 is available here because the
north-star path covers it (the reference keeps it disabled in its encoder); profile 4 (PCM archive,
the CLI default) shares profile 0's packing kernels without the transform."""
from . import profiles, profile0, profile1, profile2, profile4  # noqa: F401

AVAILABLE = [0, 1, 2, 4]

SEGMAX = [0xFFFFFFFF, profiles.compact.MAX_SMPL, profiles.compact.MAX_SMPL, 0, 0xFFFFFFFF, 0, 0, 0]

BIT_DEPTHS = [profile0.DEPTHS, profile1.DEPTHS, profile2.DEPTHS, [], profile4.DEPTHS, [], [], []]
