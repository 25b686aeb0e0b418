"""numpy restatement of the multi-GPU block protocol (smafa_b200/csrc/sharded.cu) and of the sort-free merge
(smafa_b200/csrc/merge.cu), for the CPU tests of the N > 1 path: what a shard sends, how an overflowed or failed block
is seen by every rank, and why summed binary-search ranks give the reference's selection and print order.
Test infrastructure only -- the product path runs the CUDA kernels."""
import numpy as np

KEY_Q_SHIFT, KEY_D_SHIFT = 44, 32


def make_block(local_rows, cap, status=0):
    """local_rows: int [n, 3] (query, GLOBAL subject, distance) in print order -> uint64 [2 + cap] block:
    [0] = n (may exceed cap: overflow), [1] = status, [2:] = the first min(n, cap) keys."""
    r = np.asarray(local_rows, dtype=np.uint64).reshape(-1, 3)
    keys = (r[:, 0] << np.uint64(KEY_Q_SHIFT)) | (r[:, 2] << np.uint64(KEY_D_SHIFT)) | r[:, 1]
    block = np.zeros(2 + cap, dtype=np.uint64)
    block[0], block[1] = len(keys), status
    block[2:2 + min(len(keys), cap)] = keys[:cap]
    return block


def merge_blocks(gathered, cap, k):
    """gathered: uint64 [R, 2 + cap].  k: keep rows <= the k-th smallest distance of their query (1 = Mode A; None =
    keep all).  -> (rows uint32 [n, 3] in print order, need = largest announced count, status, rank of that status)."""
    R = gathered.shape[0]
    need = int(gathered[:, 0].max())
    bad = np.nonzero(gathered[:, 1])[0]
    status, who = (int(gathered[bad[0], 1]), int(bad[0])) if len(bad) else (0, 0)
    keys = [gathered[r, 2:2 + min(int(gathered[r, 0]), cap)] for r in range(R)]
    total = sum(len(x) for x in keys)
    merged = np.zeros(total, dtype=np.uint64)
    keep = np.zeros(total, dtype=bool)
    kk = np.iinfo(np.int64).max if k is None else int(k)
    for r in range(R):
        mine = keys[r]
        q0 = mine & ~np.uint64((1 << KEY_Q_SHIFT) - 1)        # (q, 0, 0): first row of the query
        d0 = mine & ~np.uint64(0xFFFFFFFF)                     # (q, d, 0): first row at this distance
        pos = np.zeros(len(mine), dtype=np.int64)
        less_d = np.zeros(len(mine), dtype=np.int64)
        for o in range(R):
            seg = np.searchsorted(keys[o], q0, side="left")
            at_d = np.searchsorted(keys[o], d0, side="left")
            less_d += at_d - seg
            pos += np.arange(len(mine)) if o == r else np.searchsorted(keys[o], mine, side="left")
        merged[pos] = mine
        keep[pos] = less_d < kk
    sel = merged[keep]
    rows = np.stack([sel >> np.uint64(KEY_Q_SHIFT), sel & np.uint64(0xFFFFFFFF),
                     (sel >> np.uint64(KEY_D_SHIFT)) & np.uint64(0xFFF)], axis=1).astype(np.uint32)
    return rows, need, status, who


def next_cap(need):
    """capacity after an overflow (sharded.cu: pow2_at_least(need + need / 4))"""
    want, p = need + need // 4, 1
    while p < want:
        p <<= 1
    return p
