"""Batched form of ``aln2counts.SequenceReport._map_to_coordinate_ref``.

Reference: micall/core/aln2counts.py:174-304.  For one coordinate reference the reference makes seven serial
``aligner.align`` calls (``Aligner(gop=40, gep=10, is_global=False, model='EmpHIV25')``, aln2counts.py:31-37):
the consensus of each of the three reading frames against the coordinate reference (:213-223), the seed
translated in each of the three frames against the coordinate reference (:241-250) and the best seed frame against
the best consensus (:267-268).  The first six are mutually independent, and a seed has up to ten coordinate
references, so ``map_coordinate_refs`` sends 6 x (number of coordinate references) alignments to the device in ONE
batch and the dependent seventh of every coordinate reference in a second one.  The index bookkeeping between the
calls is restated from the reference so that the maps are identical.  There is no CPU path.
"""
from .gotoh2 import Aligner

# aln2counts.py:31-32
GAP_OPEN_COORD, GAP_EXTEND_COORD = 40, 10


def default_aligner(library=None, device=0):
    """The module-level aligner of aln2counts.py:34-37."""
    return Aligner(gop=GAP_OPEN_COORD, gep=GAP_EXTEND_COORD, is_global=False, model="EmpHIV25", library=library, device=device)


def _as_str(s):
    return s.decode("utf-8") if type(s) == bytes else s                                               # aln2counts.py:181-185


def pair_align(reference, query, aligner=None):
    """``SequenceReport._pair_align`` (aln2counts.py:174-189): bytes are decoded, then one alignment."""
    return (aligner or default_aligner()).align(_as_str(reference), _as_str(query))


class CoordinateMap:
    """What _map_to_coordinate_ref derives from its alignments for one coordinate reference.

    reading_frame   best reading frame, or None when no frame scored above min(consensus_length, len(coordinate_ref))
    consensus       consensus of that frame (frame 0's when nothing aligned, aln2counts.py:217-219)
    seed_amino_seq  the seed translation that aligned best to the coordinate reference (None without a reading frame)
    ref2seed        {coordinate index: seed index}      (aln2counts.py:254-264)
    seed2conseq     {seed index: consensus index}       (aln2counts.py:271-281)
    """

    def __init__(self, coordinate_ref, consensus):
        self.coordinate_ref = coordinate_ref
        self.reading_frame = None
        self.consensus = consensus
        self.seed_amino_seq = None
        self.ref2seed = {}
        self.seed2conseq = {}

    def conseq_indexes(self):
        """[(coordinate position (1-based), consensus index or None)] in coordinate order: the ReportAmino list of
        aln2counts.py:288-299 without the count objects."""
        return [(r + 1, self.seed2conseq.get(self.ref2seed[r])) for r in sorted(self.ref2seed)]

    def inserts(self):
        """Consensus indexes that map to no coordinate position (aln2counts.py:283-302)."""
        used = {c for _, c in self.conseq_indexes() if c is not None}
        return set(range(len(self.consensus))) - used


def _walk(aligned_a, aligned_b, seq_a, seq_b):
    """The two-cursor loops of aln2counts.py:254-264 / :271-281: {index in b: index in a}, built while both aligned
    strings are read column by column and a cursor only advances when the column shows that sequence's next
    character."""
    out = {}
    ia = ib = 0
    for ca, cb in zip(aligned_a, aligned_b):
        if ia < len(seq_a) and ca == seq_a[ia]:
            out[ib] = ia
            ia += 1
        if ib < len(seq_b) and cb == seq_b[ib]:
            ib += 1
    return out


def map_coordinate_refs(requests, aligner=None, library=None, device=0):
    """requests: iterable of (coordinate_ref, frame_consensus, consensus_length, seed_amino_seqs)

        coordinate_ref    amino-acid coordinate reference (str or bytes)
        frame_consensus   {reading_frame: consensus} in the iteration order of SequenceReport.seed_aminos
        consensus_length  number of frame-0 positions with counts (aln2counts.py:209)
        seed_amino_seqs   [translate(seed_nuc_seq, offset=f, ambig_char='-') for f in range(3)] (aln2counts.py:243-245)

    returns one CoordinateMap per request; all requests share two device batches."""
    al = aligner or default_aligner(library, device)
    reqs = [(_as_str(cref), dict(fc), int(clen), list(seeds)) for cref, fc, clen, seeds in requests]
    # ---- round 1: consensus frames and seed frames against the coordinate reference ---------------------------
    batch, where = [], []
    for r, (cref, fc, _clen, seeds) in enumerate(reqs):
        for frame, consensus in fc.items():
            where.append((r, "frame", frame)); batch.append((cref, consensus))                        # aln2counts.py:221
        for sf, seed_aa in enumerate(seeds):
            where.append((r, "seed", sf)); batch.append((seed_aa, cref))                              # aln2counts.py:247-248
    scores = al.align_batch(batch)
    maps = []
    best_seed = []
    for r, (cref, fc, clen, seeds) in enumerate(reqs):
        first_frame = next(iter(fc))
        m = CoordinateMap(cref, fc.get(0, fc[first_frame]))                                            # aln2counts.py:217-219
        max_score = min(clen, len(cref))                                                              # aln2counts.py:209-210
        max_seed_score, seed_choice = 0, None
        for (rr, kind, key), (a1, a2, score) in zip(where, scores):
            if rr != r:
                continue
            if kind == "frame":
                if score > max_score:                                                                 # aln2counts.py:222-224
                    max_score = score
                    m.reading_frame, m.consensus = key, fc[key]
            elif score > max_seed_score:                                                              # aln2counts.py:249-251
                max_seed_score = score
                seed_choice = (seeds[key], a1, a2)
        maps.append(m)
        best_seed.append(seed_choice)
    # ---- round 2: the chosen seed translation against the chosen consensus ---------------------------------------
    todo = [r for r, m in enumerate(maps) if m.reading_frame is not None]
    if any(best_seed[r] is None for r in todo):
        # no seed frame scored above 0: the reference unpacks None at aln2counts.py:253 and dies with this TypeError
        raise TypeError("cannot unpack non-iterable NoneType object")
    second = al.align_batch([(best_seed[r][0], maps[r].consensus) for r in todo])                     # aln2counts.py:267-268
    for r, (aseed2, aconseq, _score) in zip(todo, second):
        m = maps[r]
        seed_aa, aseed, aref = best_seed[r]
        m.seed_amino_seq = seed_aa
        m.ref2seed = _walk(aseed, aref, seed_aa, m.coordinate_ref)                                    # aln2counts.py:254-264
        aconseq = aconseq.replace("?", "-")                                                           # aln2counts.py:269
        m.seed2conseq = _walk(aconseq, aseed2, m.consensus, seed_aa)                                  # aln2counts.py:271-281
    return maps
