"""Profile constants of the FrAD format (mirror of libfrad/fourier/profiles.py:1-34)."""

LOSSLESS = [0, 4]
COMPACT = [1, 2]


class compact:
    """Legal sample rates / frame sizes of the compact (lossy) profiles and their header indices."""
    SRATES = (96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000)
    SAMPLES = [base << shift for shift in range(8) for base in (128, 160, 192, 224)]
    MAX_SMPL = max(SAMPLES)

    @staticmethod
    def get_valid_srate(srate: int) -> int:
        return min(r for r in compact.SRATES if r >= srate)

    @staticmethod
    def get_srate_index(srate: int) -> int:
        return compact.SRATES.index(compact.get_valid_srate(srate))

    @staticmethod
    def get_samples_min_ge(smpl: int) -> int:
        return min(s for s in compact.SAMPLES if s >= smpl)

    @staticmethod
    def get_samples_index(smpl: int) -> int:
        return compact.SAMPLES.index(compact.get_samples_min_ge(smpl))
