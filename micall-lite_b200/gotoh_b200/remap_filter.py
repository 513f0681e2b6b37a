"""Seed-distance filter of ``remap.sam_to_conseqs`` on the GPU.

Reference: micall/core/remap.py:129-138 (``extract_relevant_seed``) and :231-263 (the ``is_filtered`` branch):
every consensus is aligned globally against every seed that produced a consensus (K x K alignments of up to
~9.6 kb x ~9.6 kb, ``Aligner(gop=15, gep=3, is_global=True)``, remap.py:33,248), the seed is cut to the part the
consensus covers and ``Levenshtein.distance`` of the two decides whether the consensus drifted to another seed.

Here the K x K alignments go to the device as ONE batch (gotoh_b200_gotoh2_align_batch) and the K x K edit
distances as one more (gotoh_b200_edit_distance_batch, the score-only forward kernel); the bookkeeping around
them is restated from the reference so that results are identical.  There is no CPU path.
"""
import numpy as np

from . import _ffi, packing
from .gotoh2 import Aligner

# remap.py:33
GAP_OPEN, GAP_EXTEND = 15, 3


def _code_points(x):
    if isinstance(x, str):
        return np.frombuffer(x.encode("utf-32-le"), dtype=np.uint32)
    return np.frombuffer(bytes(x), dtype=np.uint8).astype(np.uint32)


def _dense_pair(a, b):
    """Edit distance only asks whether two symbols are equal: the symbols the two strings share become codes 1..k, every
    other symbol of ``a`` code 29 and of ``b`` code 30 (they match nothing on the other side).  Works on characters, so
    'h\u00e9llo' vs 'hello' is one substitution like in python-Levenshtein, whatever the UTF-8 byte lengths."""
    ua, ub = _code_points(a), _code_points(b)
    common = np.intersect1d(ua, ub)
    if len(common) > 28:
        raise ValueError("distance(): the two strings share %d distinct symbols; this kernel handles up to 28" % len(common))

    def enc(u, other):
        if len(common) == 0 or len(u) == 0:
            return np.full(len(u), other, dtype=np.uint8)
        i = np.minimum(np.searchsorted(common, u), len(common) - 1)
        return np.where(common[i] == u, i + 1, other).astype(np.uint8)
    return enc(ua, 29), enc(ub, 30)


def _device_distances(lib, a_list, b_list, device):
    a, ao = packing.pack(a_list, "a")
    b, bo = packing.pack(b_list, "b")
    out = np.zeros(len(a_list), np.int32)
    if a.size == 0:
        a = np.zeros(1, np.uint8)
    if b.size == 0:
        b = np.zeros(1, np.uint8)
    rc = lib.lib.gotoh_b200_edit_distance_batch(a.ctypes.data, ao.ctypes.data, b.ctypes.data, bo.ctypes.data,
                                                len(a_list), out.ctypes.data, int(device))
    return rc, out


def distance_batch(pairs, library=None, device=0):
    """[(a, b), ...] -> [Levenshtein.distance(a, b), ...] (remap.py:250) in one device call.

    ASCII batches with at most 30 distinct symbols on both sides (every nucleotide / amino-acid batch) go to the device
    as they are.  Anything else - non-ASCII text, or a batch that mixes alphabets - is re-coded pair by pair first
    (``_dense_pair``), which keeps the drop-in contract for arbitrary strings as long as one PAIR shares at most 28
    distinct symbols."""
    pairs = list(pairs)
    if not pairs:
        return []
    lib = library or _ffi.default_library()
    ascii_only = all((isinstance(x, str) and x.isascii()) or isinstance(x, (bytes, bytearray)) for p in pairs for x in p[:2])
    if ascii_only:
        rc, out = _device_distances(lib, [p[0] for p in pairs], [p[1] for p in pairs], device)
        if rc != _ffi.ERANGE:
            lib.check(rc)
            return [int(x) for x in out]
    coded = [_dense_pair(p[0], p[1]) for p in pairs]
    rc, out = _device_distances(lib, [c[0].tobytes() for c in coded], [c[1].tobytes() for c in coded], device)
    lib.check(rc)
    return [int(x) for x in out]


def distance(a, b, library=None, device=0):
    """Drop-in for ``Levenshtein.distance(a, b)``."""
    return distance_batch([(a, b)], library, device)[0]


def extract_relevant_seed(aligned_conseq, aligned_seed):
    """The portion of the seed that mapped to or was surrounded by the consensus (remap.py:129-138): the columns
    from the first to the last non-gap character of ``aligned_conseq``, gaps removed."""
    first = 0
    while first < len(aligned_conseq) and aligned_conseq[first] == "-":
        first += 1
    if first == len(aligned_conseq):
        # the reference's regex needs one non-gap character; an all-gap consensus cannot come out of the aligner
        raise AttributeError("aligned_conseq holds no sequence character")
    last = len(aligned_conseq)
    while aligned_conseq[last - 1] == "-":
        last -= 1
    return aligned_seed[first:last].replace("-", "")


def relevant_conseq(conseq, counts, filter_coverage):
    """Positions of ``conseq`` (1-based keys of ``counts``: {pos: {nuc: count}}) with at least ``filter_coverage``
    reads (remap.py:236-240)."""
    return "".join(c for pos, c in enumerate(conseq, 1) if sum(counts[pos].values()) >= filter_coverage)


def seed_distances(relevant, seeds, names=None, aligner=None, library=None, device=0):
    """{name: relevant_conseq}, {name: seed_ref} -> {name: {seed_name: distance}} for every seed_name in ``names``
    (default: the keys of ``relevant``, like the loops of remap.py:232,246): all alignments in one batch, all edit
    distances in one more."""
    names = sorted(relevant.keys() if names is None else names)
    al = aligner or Aligner(gop=GAP_OPEN, gep=GAP_EXTEND, is_global=True, library=library, device=device)
    todo = [(name, seed_name) for name in sorted(relevant) if relevant[name] for seed_name in names]
    aligned = al.align_batch([(seeds[seed_name], relevant[name]) for name, seed_name in todo])        # remap.py:248
    cut = [(extract_relevant_seed(aconseq, aseed), relevant[name]) for (name, _), (aseed, aconseq, _s) in zip(todo, aligned)]
    dist = distance_batch(cut, library=al._libobj, device=al.device)                                 # remap.py:250
    out = {}
    for (name, seed_name), d in zip(todo, dist):
        out.setdefault(name, {})[seed_name] = d
    return out


def filter_conseqs(new_conseqs, relevant, seeds, read_counts=None, distance_report=None, aligner=None, library=None,
                   device=0):
    """The ``is_filtered`` tail of sam_to_conseqs (remap.py:228-263).

    new_conseqs      {name: consensus}
    relevant         {name: relevant_conseq(...)} for the same names ('' = no acceptable coverage)
    seeds            {name: seed reference}
    read_counts      collections.Counter of reads per reference (only used when nothing passes, remap.py:259-262)
    distance_report  optional dict, filled like remap.py:255-258
    returns          {name: consensus} of the consensuses that are at least as close to their own seed as to any other
    """
    if len(new_conseqs) < 2:                                                                          # remap.py:229
        return dict(new_conseqs)
    names = sorted(new_conseqs.keys())
    dists = seed_distances({n: relevant.get(n, "") for n in names}, seeds, names, aligner, library, device)
    filtered = {}
    for name in names:
        if not relevant.get(name):
            continue                                                                                  # remap.py:241-243
        seed_dist = other_seed = other_dist = None
        for seed_name in names:                                                                       # remap.py:246-254
            d = dists[name][seed_name]
            if seed_name == name:
                seed_dist = d
            elif other_dist is None or d < other_dist:
                other_seed, other_dist = seed_name, d
        if seed_dist <= other_dist:
            filtered[name] = new_conseqs[name]
        if distance_report is not None:
            distance_report[name] = dict(seed_dist=seed_dist, other_dist=other_dist, other_seed=other_seed)
    if not filtered:
        best_ref = read_counts.most_common(1)[0][0]                                                   # remap.py:259-262
        filtered[best_ref] = new_conseqs[best_ref]
    return filtered
