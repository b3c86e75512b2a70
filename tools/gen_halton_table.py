"""Generates (and checks) the per-prime reciprocal table c_haltonBase of csrc/shade.cuh:
i / p == umulhi(i, magic) >> shift for every 0 <= i < 2^31 (Granlund & Montgomery round-up reciprocal, N = 31)."""
import numpy as np

PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101, 103, 107,
          109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179, 181, 191, 193, 197, 199, 211, 223, 227, 229,
          233, 239, 241, 251, 257, 263, 269, 271, 277, 281, 283, 293, 307, 311, 313, 317, 331, 337, 347, 349, 353, 359,
          367, 373, 379, 383, 389, 397, 401, 409, 419, 421, 431, 433, 439, 443, 449, 457, 461, 463, 467, 479, 487, 491,
          499, 503, 509, 521, 523, 541]


def entry(p):
    L = (p - 1).bit_length()
    magic = (1 << (31 + L)) // p + 1
    assert magic < (1 << 32) and (1 << (31 + L)) <= magic * p <= (1 << (31 + L)) + (1 << L)  # exactness condition
    inv = (np.float32(1.0) / np.float32(p)).view(np.uint32)
    return p, magic, L - 1, int(inv)


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    rows = []
    for p in PRIMES:
        b, m, sh, inv = entry(p)
        n = np.concatenate([rng.integers(1, 1 << 31, 200000, dtype=np.uint64),
                            np.array([1, 2, p - 1, p, p + 1, (1 << 31) - 1, (1 << 31) - p, 1 << 20], dtype=np.uint64)])
        assert np.array_equal(((n * np.uint64(m)) >> np.uint64(32)) >> np.uint64(sh), n // np.uint64(p)), p
        rows.append("{%du, 0x%08Xu, %du, 0x%08Xu}" % (b, m, sh, inv))
    for k in range(0, 100, 3):
        print("    " + ", ".join(rows[k:k + 3]) + ",")
