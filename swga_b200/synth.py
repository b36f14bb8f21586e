"""Position-indexed synthetic genomes: integer-only, so every shard, the CPU
oracle and the GPU generator kernel (csrc/swga_b200.cu: k_synth) agree bit-for-bit without
communication.

    h(x)  = splitmix64 finaliser
    r(i)  = h((seed << 40) + i) >> 40                      (24 bit)
    base  = A if r < tA, C if r < tC, G if r < tG, else T
    tA = round(2^24 (1-g)/2), tC = tA + round(2^24 g/2), tG = tC + round(2^24 g/2)
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def thresholds(gc):
    ta = int(round((1 << 24) * (1.0 - gc) / 2.0))
    tc = ta + int(round((1 << 24) * gc / 2.0))
    tg = tc + int(round((1 << 24) * gc / 2.0))
    return ta, tc, tg


def bases(seed, start, n, gc):
    """uint8 ASCII bases for positions [start, start+n) of genome `seed`."""
    ta, tc, tg = thresholds(gc)
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 22
    for o in range(0, n, step):
        m = min(step, n - o)
        i = np.arange(start + o, start + o + m, dtype=np.uint64)
        r = (splitmix64((np.uint64(seed) << np.uint64(40)) + i) >> np.uint64(40)).astype(np.uint32)
        b = np.full(m, ord("T"), dtype=np.uint8)
        b[r < tg] = ord("G")
        b[r < tc] = ord("C")
        b[r < ta] = ord("A")
        out[o:o + m] = b
    return out


def equal_records(n, n_rec):
    """rec_starts (uint64[n_rec+1]) splitting n bases into n_rec records of (almost) equal length."""
    return np.array([(n * r) // n_rec for r in range(n_rec + 1)], dtype=np.uint64)


def write_fasta(path, seq, rec_starts, names=None, width=70):
    """Fixed-width FASTA (pyfaidx needs uniform line width)."""
    with open(path, "wb") as f:
        for r in range(len(rec_starts) - 1):
            name = names[r] if names else "rec%d" % (r + 1)
            f.write((">%s\n" % name).encode())
            s = seq[int(rec_starts[r]):int(rec_starts[r + 1])]
            full = (len(s) // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = s[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if len(s) > full:
                f.write(s[full:].tobytes() + b"\n")


# The BASELINE.json configs: name -> (seed, n_bases, gc, n_records)
CONFIGS = {
    "C1_fg": (1, 1_000_000, 0.50, 1),
    "C1_bg": (2, 10_000_000, 0.50, 1),
    "C2_fg": (3, 4_400_000, 0.656, 1),
    "C2_bg": (4, 3_100_000_000, 0.41, 24),
    "C3_fg": (5, 23_000_000, 0.194, 14),
    "C3_bg": (4, 3_100_000_000, 0.41, 24),
    "C5_fg": (6, 1_300_000, 0.352, 1),
    "C5_bg": (7, 140_000_000, 0.42, 6),
}
