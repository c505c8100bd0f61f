"""Expectations per compression tier: levels 0..9 are byte-identical to the
oracle; levels 10..12 are held to the oracle's SIZE (batch total within 0.5 %) and to validity (the
stream inflates to the input under system zlib) — except streams above 64 KiB, which still run the
step-for-step transliteration (deflate_bt.cuh) and are byte-identical too."""
import zlib

WBITS = {0: -15, 1: 15, 2: 31}


def check_stream(got, src, exp, level, fmt, where=""):
    """One compressed stream against the oracle's (exp = None: the oracle did not fit its bound)."""
    if level < 10 or len(src) > 65536:
        assert got == (exp if exp is not None else b""), (where, level, fmt, len(src))
        return
    if exp is None:
        assert got == b"" or zlib.decompress(got, WBITS[fmt]) == src, (where, level, fmt, len(src))
        return
    assert got != b"", (where, level, fmt, len(src))
    assert zlib.decompress(got, WBITS[fmt]) == src, (where, level, fmt, len(src))


def check_total(pairs, level, tol=1.005):
    """pairs of (got, exp) byte strings: the batch total stays within the tolerance."""
    g = sum(len(a) for a, b in pairs if b is not None)
    e = sum(len(b) for a, b in pairs if b is not None)
    assert g <= e * tol + 64, (level, g, e)
