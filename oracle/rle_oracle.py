"""CPU oracle of the COCO run-length mask encoding — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and bench.py's
cpu legs may import this; the product never does).

The reference turns every predicted mask into this wire format right after the hot path
(projects/IDOL/idol/data/ytvis_eval.py:174-214, projects/SeqFormer/seqformer/data/ytvis_eval.py:171-211:
`mask_util.encode(np.array(mask[:, :, None], order="F", dtype="uint8"))[0]`, then `counts.decode("utf-8")`).  The algorithm
lives in a third-party dependency that is NOT in /root/reference: **pycocotools** (requirements.txt:1 unpinned,
setup.py:165 `pycocotools>=2.0.2`), C file `common/maskApi.c` — `rleEncode` (column-major run lengths, first run counts
zeros), `rleToString` / `rleFrString` (each count minus the count two places before it, for index > 2, written as 5-bit
groups, LSB first, bit 0x20 = "more", + 48 to make it printable) and `rleDecode`.  This file restates that published
algorithm; it is PINNED by the RLE strings the reference's own tests hold (tests/data/test_coco_evaluation.py:24,
tests/test_visualizer.py:54 -> tests/golden/rle_golden.json): every string must decode to exactly size[0]*size[1] pixels and
re-encode to the identical string (tests/test_rle_oracle.py), and `counts_scalar` follows tests/data/test_coco.py:31-42.
"""
import numpy as np


# ------------------------------------------------------------------ scalar restatements (small cases)
def counts_scalar(mask):
    """maskApi.c rleEncode for one [H, W] mask: run lengths of the column-major pixel stream, starting with a zeros run."""
    flat = np.asarray(mask).astype(bool).flatten(order="F").tolist()
    counts, prev, run = [], False, 0
    for v in flat:
        if v != prev:
            counts.append(run)
            run, prev = 0, v
        run += 1
    counts.append(run)
    return counts


def to_string_scalar(counts):
    """maskApi.c rleToString."""
    out = bytearray()
    for i, cnt in enumerate(counts):
        x = int(cnt)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            c = x & 0x1F
            x >>= 5                                   # Python's >> is arithmetic, like C's on a long
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(c + 48)
    return bytes(out)


def from_string(s):
    """maskApi.c rleFrString: bytes / str -> list of counts."""
    if isinstance(s, str):
        s = s.encode("ascii")
    counts, p = [], 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def decode(counts, h, w):
    """maskApi.c rleDecode: counts -> [h, w] uint8 mask (column-major stream)."""
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, v = 0, 0
    for c in counts:
        if v:
            flat[pos:pos + c] = 1
        pos += c
        v ^= 1
    if pos != h * w:
        raise ValueError(f"run lengths sum to {pos}, expected {h * w}")
    return flat.reshape((h, w), order="F")


# ------------------------------------------------------------------ vectorised (full-size masks)
def counts(mask):
    flat = np.asarray(mask).astype(bool).flatten(order="F")
    n = flat.size
    prev = np.concatenate(([False], flat[:-1]))
    starts = np.flatnonzero(flat != prev)             # positions where a new run begins (position 0 only if the mask starts with 1)
    edges = np.concatenate(([0], starts, [n]))
    return np.diff(edges).astype(np.int64)            # first count = zeros before the first 1 (0 if the mask starts with 1)


def to_string(cnts):
    cnts = np.asarray(cnts, dtype=np.int64)
    x = cnts.copy()
    if x.size > 3:
        x[3:] -= cnts[1:-2]
    cols, alive = [], np.ones(x.shape, dtype=bool)
    for _ in range(14):
        c = x & 0x1F
        x = x >> 5
        more = np.where((c & 0x10) != 0, x != -1, x != 0)
        cols.append(np.where(alive, (c | (more.astype(np.int64) << 5)) + 48, 0))
        alive = alive & more
        if not alive.any():
            break
    arr = np.stack(cols, 1)
    return arr[arr != 0].astype(np.uint8).tobytes()


def encode(mask):
    """pycocotools mask.encode of one [H, W] mask -> {"size": [H, W], "counts": bytes}."""
    mask = np.asarray(mask)
    return {"size": [int(mask.shape[0]), int(mask.shape[1])], "counts": to_string(counts(mask))}
