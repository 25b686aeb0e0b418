"""Synthetic SingleM-style window sets for tests and bench.py.

Sequences are produced as symbol-index arrays (0=A 1=C 2=G 3=T 4=N/gap) and packed into the
reference's word layout: 12 symbols per u64, 5-bit one-hot code of symbol p at bit 5*(p%12)
of word p//12 (reference src/lib.rs:29-52, codes src/lib.rs:167-184).
"""
import numpy as np

SEED_DB = 0x5AFA0001
SEED_QUERY = 0x5AFA0002
SEED_CLUSTER = 0x5AFA0005

_CODE = np.array([16, 8, 4, 2, 1], dtype=np.uint64)  # A C G T N
_ASCII = np.frombuffer(b"ACGTN", dtype=np.uint8)


def words_for_len(L):
    return (L + 11) // 12


def pack_symbols(sym):
    """uint8 [n, L] symbol indices -> uint64 [n, ceil(L/12)] in the reference bit layout."""
    n, L = sym.shape
    out = np.zeros((n, words_for_len(L)), dtype=np.uint64)
    for lo in range(0, n, 1 << 19):  # in slabs: the u64 code image of 10 M windows would be 4.8 GB at once
        codes = _CODE[sym[lo:lo + (1 << 19)]]
        for p in range(L):
            out[lo:lo + (1 << 19), p // 12] |= codes[:, p] << np.uint64(5 * (p % 12))
    return out


def to_ascii(sym, gap_fraction=0.5, seed=7):
    """Symbol indices -> list[bytes]; symbol 4 is written as 'N' or '-' (both encode to N)."""
    a = _ASCII[sym].copy()
    rng = np.random.default_rng(seed)
    gaps = (sym == 4) & (rng.random(sym.shape) < gap_fraction)
    a[gaps] = ord("-")
    return [row.tobytes() for row in a]


def write_fasta(path, seqs, prefix="seq"):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">" + prefix.encode() + str(i).encode() + b"\n" + s + b"\n")


def _mutate(rng, base, max_subs, noise):
    """Apply s ~ U{0..max_subs} substitutions (positions drawn with replacement, each to a base
    different from the original) and independent N/gap noise with probability `noise`."""
    n, L = base.shape
    out = base.copy()
    s = rng.integers(0, max_subs + 1, size=n)
    for j in range(max_subs):
        rows = np.nonzero(s > j)[0]
        if rows.size == 0:
            break
        pos = rng.integers(0, L, size=rows.size)
        shift = rng.integers(1, 4, size=rows.size).astype(np.uint8)
        out[rows, pos] = (base[rows, pos] + shift) & 3
    if noise > 0:
        # in row slabs: the same draws in the same order as one rng.random(out.shape), without its 8 bytes per symbol at once
        for lo in range(0, n, 1 << 19):
            part = out[lo:lo + (1 << 19)]
            part[rng.random(part.shape) < noise] = 4
    return out


def make_db(D, L=60, seed=SEED_DB, family=16, max_subs=8, noise=0.01):
    """D windows in families of `family` descendants per random root (s=0 gives exact
    duplicates, which exercise ties and --limit-per-sequence)."""
    rng = np.random.default_rng(seed)
    R = max(1, D // family)
    roots = rng.integers(0, 4, size=(R, L), dtype=np.uint8)
    base = roots[np.arange(D) % R]
    return _mutate(rng, base, max_subs, noise)


def make_queries(db_sym, Q, seed=SEED_QUERY, max_subs=10, noise=0.01):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, db_sym.shape[0], size=Q)
    return _mutate(rng, db_sym[src], max_subs, noise)


def make_cluster_input(n, L=60, seed=SEED_CLUSTER, family=20, max_subs=3, dup_fraction=0.02):
    """n sequences: roots x `family` descendants within max_subs of the root, a few exact
    duplicates, shuffled (order-dependent greedy outcomes)."""
    rng = np.random.default_rng(seed)
    R = max(1, n // family)
    roots = rng.integers(0, 4, size=(R, L), dtype=np.uint8)
    sym = _mutate(rng, roots[np.arange(n) % R], max_subs, 0.0)
    ndup = int(n * dup_fraction)
    if ndup:
        dst = rng.integers(0, n, size=ndup)
        sym[dst] = sym[rng.integers(0, n, size=ndup)]
    return sym[rng.permutation(n)]


def random_symbols(n, L, seed, p_n=0.05):
    """Unstructured random windows (uniform ACGT with N at rate p_n)."""
    rng = np.random.default_rng(seed)
    sym = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
    sym[rng.random((n, L)) < p_n] = 4
    return sym


# ---- protein windows (configs[3]; an extension of this build: the reference has no amino-acid mode) ----
SEED_PROTEIN = 0x5AFA0004
_AA_ASCII = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX-*", dtype=np.uint8)  # symbol index s -> byte; number = s + 1


def pack_symbols_aa(sym):
    """uint8 [n, L] protein symbol indices (0..19 amino acids, 20 = X, 21 = '-', 22 = '*') -> uint64 words
    holding the symbol NUMBER (index + 1) in the same 5-bit groups as the nucleotide layout."""
    n, L = sym.shape
    out = np.zeros((n, words_for_len(L)), dtype=np.uint64)
    for lo in range(0, n, 1 << 19):
        codes = sym[lo:lo + (1 << 19)].astype(np.uint64) + np.uint64(1)
        for p in range(L):
            out[lo:lo + (1 << 19), p // 12] |= codes[:, p] << np.uint64(5 * (p % 12))
    return out


def to_ascii_aa(sym):
    return [row.tobytes() for row in _AA_ASCII[sym]]


def _mutate_aa(rng, base, max_subs, noise):
    n, L = base.shape
    out = base.copy()
    s = rng.integers(0, max_subs + 1, size=n)
    for j in range(max_subs):
        rows = np.nonzero(s > j)[0]
        if rows.size == 0:
            break
        pos = rng.integers(0, L, size=rows.size)
        shift = rng.integers(1, 20, size=rows.size).astype(np.uint8)
        out[rows, pos] = (base[rows, pos] + shift) % 20
    if noise > 0:  # X / gap / stop noise, like the N/gap noise of the nucleotide generator
        hit = rng.random(out.shape) < noise
        out[hit] = rng.integers(20, 23, size=int(hit.sum()), dtype=np.uint8)
    return out


def make_db_aa(D, L=20, seed=SEED_PROTEIN, family=16, max_subs=4, noise=0.01):
    """Protein analogue of make_db (same family construction over 20 letters, s ~ U{0..4})."""
    rng = np.random.default_rng(seed)
    R = max(1, D // family)
    roots = rng.integers(0, 20, size=(R, L), dtype=np.uint8)
    return _mutate_aa(rng, roots[np.arange(D) % R], max_subs, noise)


def make_queries_aa(db_sym, Q, seed=SEED_PROTEIN + 1, max_subs=5, noise=0.01):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, db_sym.shape[0], size=Q)
    return _mutate_aa(rng, db_sym[src], max_subs, noise)
