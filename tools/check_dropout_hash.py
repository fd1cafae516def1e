"""Statistical sanity check (CPU, numpy) of the dropout draw generator in deepsvg_b200/csrc/common.cuh.

Four 16-bit draws per quad index from one shared first mixing stage and two finalisers.  Checks, for several keys:
marginal drop rates of the 4 positions, pairwise joint drop rates inside a quad, joint rates at lags 1..1024 of the
flattened element stream, and chi-square of byte histograms of the draws.  Everything should sit within a few sigma
of the binomial / chi-square expectation.   Usage: python tools/check_dropout_hash.py [log2_quads]
"""
import sys
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def mul(a, c):
    return (a.astype(np.uint64) * np.uint64(c)) & M32


def stage1(q, hikey):
    x = mul(q, 0x9E3779B1) ^ np.uint64(hikey)
    x ^= x >> np.uint64(16)
    x = mul(x, 0x7FEB352D)
    x ^= x >> np.uint64(15)
    return x


def fin(s, c):
    y = mul(s, c)
    return y ^ (y >> np.uint64(16))


def draws(q, key):
    s = stage1(q, key)
    a, b = fin(s, 0x846CA68B), fin(s, 0xC2B2AE35)
    return np.stack([a & np.uint64(0xFFFF), a >> np.uint64(16), b & np.uint64(0xFFFF), b >> np.uint64(16)], 1)


def main():
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
    q = np.arange(n, dtype=np.uint64)
    thr, p = 6554, 6554 / 65536
    worst = 0.0
    for key in (0x1234567, 0xDEADBEEF, 0, 1):
        d = draws(q, key)
        drop = d < thr
        sig = np.sqrt(p * (1 - p) / n)
        z = np.abs(drop.mean(0) - p) / sig
        worst = max(worst, z.max())
        print("key %08x  marginal z-scores %s" % (key, np.round(z, 2)))
        sj = np.sqrt(p * p * (1 - p * p) / n)
        zs = []
        for i in range(4):
            for j in range(i + 1, 4):
                zs.append(abs((drop[:, i] & drop[:, j]).mean() - p * p) / sj)
        flat = drop.reshape(-1)
        for lag in (1, 2, 3, 4, 5, 8, 64, 256, 512, 1024):
            zs.append(abs((flat[:-lag] & flat[lag:]).mean() - p * p) / (sj / 2))
        worst = max(worst, max(zs))
        print("              joint z-scores max %.2f" % max(zs))
        chis = []
        for pos in range(4):
            for v in (d[:, pos] >> np.uint64(8), d[:, pos] & np.uint64(0xFF)):
                h = np.bincount(v.astype(np.int64), minlength=256)
                chis.append(((h - n / 256) ** 2 / (n / 256)).sum())
        print("              byte-histogram chi2 (dof 255): min %.0f max %.0f" % (min(chis), max(chis)))
        assert max(chis) < 255 + 6 * np.sqrt(2 * 255), "byte histogram far from uniform"
    assert worst < 5.0, "drop rates deviate by more than 5 sigma"
    print("OK (worst z-score %.2f)" % worst)


if __name__ == "__main__":
    main()
