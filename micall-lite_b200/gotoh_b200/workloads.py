"""Synthetic inputs of the shapes BASELINE.json names, seeded and vectorised.

All generators return *packed* arrays (uint8 bytes + int64 offsets) because that is what the
C ABI takes; ``unpacked()`` turns them into Python strings for small cases.  The reference
sequences (HIV-1 HXB2 pol seed, PR/RT/INT, HCV seeds) are data extracted from MiCall-Lite's
projects.json into data/references.json.
"""
import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "references.json")
_refs = None


def references():
    global _refs
    if _refs is None:
        with open(_DATA) as f:
            _refs = json.load(f)
    return _refs


def pol_seed():
    return references()["nucleotide"]["HIV1B-pol-seed"]


def hcv_seeds():
    nt = references()["nucleotide"]
    return [nt[k] for k in sorted(nt) if k.startswith("HCV-")]


def aa_refs():
    aa = references()["amino"]
    return [aa["PR"], aa["RT"], aa["INT"]]


_NT = np.frombuffer(b"ACGT", dtype=np.uint8)
_AA = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)


def _pack_rows(rows, lens):
    """rows: (n, W) uint8 with valid prefix lens[k] -> packed bytes + offsets."""
    n, w = rows.shape
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    mask = np.arange(w)[None, :] < lens[:, None]
    return np.ascontiguousarray(rows[mask]), off


def _windows(ref_arr, n, width, rng):
    starts = rng.integers(0, len(ref_arr) - width + 1, size=n)
    return ref_arr[starts[:, None] + np.arange(width)[None, :]].copy()


def _substitute(rows, alphabet, rate, rng):
    """Replace a fraction `rate` of characters by a different letter of `alphabet` (sparse: only the hit positions are
    touched, so a 1 M x 251 batch costs O(hits), not O(cells))."""
    rows = np.array(rows, dtype=np.uint8, order="C")     # private copy
    flat = rows.reshape(-1)
    k = int(rng.binomial(flat.size, rate))
    pos = rng.integers(0, flat.size, size=k)
    lut = np.zeros(256, dtype=np.int64)                       # letter -> index in `alphabet` (others: 0, they get replaced anyway)
    lut[alphabet] = np.arange(len(alphabet))
    shift = rng.integers(1, len(alphabet), size=k)
    flat[pos] = alphabet[(lut[flat[pos]] + shift) % len(alphabet)]
    return rows


def _indels(rows, lens, alphabet, p_del, p_ins, max_len, rng, extra):
    """At most one deletion and one insertion of 1..max_len characters per row (vectorised over the affected rows)."""
    n, w = rows.shape
    W = w + extra
    out = np.zeros((n, W), dtype=np.uint8)
    out[:, :w] = rows
    lens = lens.copy()
    cols = np.arange(W)[None, :]
    sel = np.nonzero(rng.random(n) < p_del)[0]
    if len(sel):
        d = rng.integers(1, max_len + 1, size=len(sel))
        ok = lens[sel] - d >= 8
        sel, d = sel[ok], d[ok]
        p = (rng.random(len(sel)) * (lens[sel] - d + 1)).astype(np.int64)          # uniform in 0 .. len-d
        src = np.minimum(cols + np.where(cols >= p[:, None], d[:, None], 0), W - 1)
        sub = np.take_along_axis(out[sel], src, axis=1)
        lens[sel] -= d
        sub[cols >= lens[sel][:, None]] = 0
        out[sel] = sub
    sel = np.nonzero(rng.random(n) < p_ins)[0]
    if len(sel):
        d = rng.integers(1, max_len + 1, size=len(sel))
        p = (rng.random(len(sel)) * (lens[sel] + 1)).astype(np.int64)              # uniform in 0 .. len
        src = np.clip(cols - np.where(cols >= (p + d)[:, None], d[:, None], 0), 0, W - 1)
        sub = np.take_along_axis(out[sel], src, axis=1)
        fresh = alphabet[rng.integers(0, len(alphabet), size=sub.shape)]
        new = (cols >= p[:, None]) & (cols < (p + d)[:, None])
        sub[new] = fresh[new]
        lens[sel] += d
        sub[cols >= lens[sel][:, None]] = 0
        out[sel] = sub
    return out, lens


def c2_reads_packed(n, seed=20260101, width=251):
    """C2: n synthetic `width`-nt reads from the HIV-1 HXB2 pol seed (3039 nt): 2 % substitutions,
    p=0.10 one deletion and p=0.10 one insertion of 1-3 nt, p=0.02 one base -> N.
    Returns (ref_str, qry_bytes, qry_off)."""
    rng = np.random.default_rng(seed)
    ref = pol_seed()
    ref_arr = np.frombuffer(ref.encode(), dtype=np.uint8)
    rows = _windows(ref_arr, n, width, rng)
    rows = _substitute(rows, _NT, 0.02, rng)
    lens = np.full(n, width, dtype=np.int64)
    rows, lens = _indels(rows, lens, _NT, 0.10, 0.10, 3, rng, extra=3)
    withn = np.nonzero(rng.random(n) < 0.02)[0]
    if len(withn):
        rows[withn, rng.integers(0, lens[withn])] = ord("N")
    qb, qo = _pack_rows(rows, lens)
    return ref, qb, qo


def c3_queries_packed(n, seed=20260103, width=84):
    """C3: n ~84-aa windows of PR (99) / RT (440) / INT (288), round-robin, 3 % substitutions,
    p=0.05 one 1-aa deletion, p=0.05 one 1-aa insertion.  Returns (refs, ref_idx, qry_bytes, qry_off)."""
    rng = np.random.default_rng(seed)
    refs = aa_refs()
    ref_idx = (np.arange(n) % 3).astype(np.int32)
    rows = np.zeros((n, width), dtype=np.uint8)
    for r in range(3):
        sel = np.nonzero(ref_idx == r)[0]
        arr = np.frombuffer(refs[r].encode(), dtype=np.uint8)
        rows[sel] = _windows(arr, len(sel), width, rng)
    rows = _substitute(rows, _AA, 0.03, rng)
    lens = np.full(n, width, dtype=np.int64)
    rows, lens = _indels(rows, lens, _AA, 0.05, 0.05, 1, rng, extra=1)
    qb, qo = _pack_rows(rows, lens)
    return refs, ref_idx, qb, qo


def c4_pairs_packed(n, seed=20260104):
    """C4: n consensus-vs-genome pairs: ref = one of the 57 HCV seeds (9.1-9.7 kb, contain N/Y/R);
    query = another HCV seed (1 in 4) or a 5 %-mutated copy with up to 10 indels of <= 30 nt.
    Returns (refs, ref_idx, qry_bytes, qry_off)."""
    rng = np.random.default_rng(seed)
    seeds = hcv_seeds()
    ref_idx = rng.integers(0, len(seeds), size=n).astype(np.int32)
    qs = []
    for k in range(n):
        if rng.random() < 0.25:
            q = seeds[int(rng.integers(0, len(seeds)))]
            qs.append(np.frombuffer(q.encode(), dtype=np.uint8))
            continue
        arr = np.frombuffer(seeds[int(ref_idx[k])].encode(), dtype=np.uint8)[None, :]
        arr = _substitute(arr, _NT, 0.05, rng)[0]
        parts = []
        cuts = np.sort(rng.integers(0, len(arr), size=int(rng.integers(0, 11))))
        prev = 0
        for c in cuts:
            parts.append(arr[prev:c])
            d = int(rng.integers(1, 31))
            if rng.random() < 0.5:
                prev = min(len(arr), c + d)           # deletion
            else:
                parts.append(_NT[rng.integers(0, 4, size=d)])  # insertion
                prev = c
        parts.append(arr[prev:])
        qs.append(np.concatenate(parts))
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(q) for q in qs], out=off[1:])
    return seeds, ref_idx, np.ascontiguousarray(np.concatenate(qs)), off


def unpacked(qb, qo):
    buf = qb.tobytes()
    return [buf[int(qo[k]):int(qo[k + 1])].decode("latin-1") for k in range(len(qo) - 1)]


def c2_reads(n, seed=20260101):
    ref, qb, qo = c2_reads_packed(n, seed)
    return ref, unpacked(qb, qo)


def c3_queries(n, seed=20260103):
    refs, _, qb, qo = c3_queries_packed(n, seed)
    return refs, unpacked(qb, qo)


def shard_range(n_pairs, rank, world):
    """Static sharding of independent pairs: contiguous, equal-count ranges."""
    lo = (n_pairs * rank) // world
    hi = (n_pairs * (rank + 1)) // world
    return lo, hi
