"""numpy restatement of the reference's hot path (TEST INFRASTRUCTURE ONLY).

A second, independent restatement next to the C one (smafa_oracle.c); the two are checked
against each other and against the reference's golden vectors in tests/test_oracle_golden.py.
Citations are relative to the reference checkout (wwood/smafa v0.8.0).
"""
import numpy as np

# src/lib.rs:167-184
_LUT = np.zeros(256, dtype=np.uint8)
for _c in b"Aa":
    _LUT[_c] = 0b10000
for _c in b"Cc":
    _LUT[_c] = 0b01000
for _c in b"Gg":
    _LUT[_c] = 0b00100
for _c in b"TtUu":
    _LUT[_c] = 0b00010
for _c in b"NWSMKRYBDHV-nwsmkrybdhv":
    _LUT[_c] = 0b00001
_DEC = {16: "A", 8: "C", 4: "G", 2: "T", 1: "N"}  # src/lib.rs:120-126


class OraclePanic(Exception):
    """Models a Rust panic! (exit code 101)."""


def words_for_len(L):
    return (L + 11) // 12  # src/lib.rs:32


def encode(seqs):
    """list[bytes] of equal length -> uint64 [n, W] (src/lib.rs:29-52)."""
    n = len(seqs)
    if n == 0:
        return np.zeros((0, 0), dtype=np.uint64)
    L = len(seqs[0])
    W = words_for_len(L)
    out = np.zeros((n, W), dtype=np.uint64)
    for i, s in enumerate(seqs):
        if len(s) != L:
            raise OraclePanic(f"WindowSet seq length is {L}, got a new sequence of length {len(s)}")
        codes = _LUT[np.frombuffer(s, dtype=np.uint8)]
        if (codes == 0).any():
            p = int(np.argmax(codes == 0))
            raise OraclePanic(f"Byte {s[p]} cannot be interpreted as nucleotide, at position {p}")
        for p in range(L):
            out[i, p // 12] |= np.uint64(int(codes[p]) << (5 * (p % 12)))
    return out


def encode_codes(codes):
    """uint8 [n, L] of 5-bit one-hot codes -> uint64 [n, W]; vectorised for big random inputs."""
    n, L = codes.shape
    W = words_for_len(L)
    out = np.zeros((n, W), dtype=np.uint64)
    for p in range(L):
        out[:, p // 12] |= codes[:, p].astype(np.uint64) << np.uint64(5 * (p % 12))
    return out


def decode(words, L):
    """src/lib.rs:113-135"""
    return "".join(_DEC[(int(words[i // 12]) >> (5 * (i % 12))) & 31] for i in range(L))


def distances(db, q, alphabet=0):
    """src/lib.rs:80-88: db [D, W], q [W] -> int64 [D].  alphabet=1: the protein extension (parity
    unpinned -- the reference has no amino-acid mode): positions whose 5-bit symbols differ."""
    if db.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    x = np.bitwise_xor(db, q[None, :])
    if alphabet:
        x = x | (x >> np.uint64(1)) | (x >> np.uint64(2)) | (x >> np.uint64(3)) | (x >> np.uint64(4))
        return np.bitwise_count(x & np.uint64(0x0084210842108421)).sum(axis=1).astype(np.int64)
    return np.bitwise_count(x).sum(axis=1).astype(np.int64) // 2


def query(db, L, queries, q_len, m=None, k=None, r=None):
    """src/lib.rs:224-317 on encoded input -> list of (query, subject, distance) in print order."""
    hits = []
    D = db.shape[0]
    mode_b = k is not None and k != 1  # :224
    for qn in range(queries.shape[0]):
        if D and q_len != L:  # :72-79
            raise OraclePanic(
                f"Cannot compute distances between seq of length {q_len} and windows of lengths {L}")
        d = distances(db, queries[qn])
        if mode_b:
            order = np.lexsort((np.arange(D), d))  # sort by (distance, index), :250
            if k > D:
                if D == 0:
                    raise OraclePanic("called `Option::unwrap()` on a `None` value")
                cutoff = int(d.max())  # :254
            else:
                if k == 0:
                    raise OraclePanic("attempt to subtract with overflow")
                cutoff = int(d[order[k - 1]])  # :255
            last, cnt = None, 0
            for i in order:
                di = int(d[i])
                if di <= cutoff and (m is None or di <= m):  # :262-264
                    if r is not None:  # :269-289
                        key = db[i].tobytes()
                        if last == key:
                            if cnt >= r:
                                continue
                            cnt += 1
                        else:
                            last, cnt = key, 1
                    hits.append((qn, int(i), di))
        else:
            if D == 0:
                raise OraclePanic("called `Option::unwrap()` on a `None` value")
            mn = int(d.min())  # :298
            if r is not None:  # :301-303
                raise OraclePanic("limit_per_sequence is implemented unless max_num_hits > 1.")
            if m is None or mn <= m:  # :306
                for i in np.nonzero(d == mn)[0]:
                    hits.append((qn, int(i), mn))
    return hits


def cluster(enc, t):
    """src/cluster.rs:22-84 on encoded input in file order.

    Returns centroid_of (input index of the assigned centroid, -1 for suppressed duplicates)
    and the number of centroids."""
    n = enc.shape[0]
    seen = set()
    cent = []  # input indices of centroids
    cof = np.full(n, -1, dtype=np.int64)
    for i in range(n):
        key = enc[i].tobytes()
        if key in seen:  # :46-48
            continue
        seen.add(key)
        if cent:
            d = distances(enc[np.array(cent)], enc[i])  # :51
            mn = int(d.min())
        else:
            d, mn = None, 2 * t + 2  # :54-58
        if mn <= t:
            a = int(np.argmax(d == mn))  # first index at the minimum, :62-68
        else:
            a = len(cent)  # :69-74
            cent.append(i)
        cof[i] = cent[a]
    return cof, len(cent)


def finalize_candidates(cands, m=None, k=None):
    """Reference selection applied to a candidate list [(q, subject, d)] that is a superset of
    the true answer for every query (the multi-GPU merge: SURVEY 8e).  Used by the gloo tests
    to check the shard/all-gather plumbing without a GPU."""
    byq = {}
    for q, s, d in cands:
        byq.setdefault(q, []).append((d, s))
    out = []
    mode_b = k is not None and k != 1
    for q in sorted(byq):
        lst = sorted(set(byq[q]))
        if mode_b:
            cutoff = lst[k - 1][0] if k <= len(lst) else lst[-1][0]
            out += [(q, s, d) for d, s in lst if d <= cutoff and (m is None or d <= m)]
        else:
            mn = lst[0][0]
            if m is None or mn <= m:
                out += [(q, s, d) for d, s in lst if d == mn]
    return out
