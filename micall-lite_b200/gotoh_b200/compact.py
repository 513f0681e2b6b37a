"""Compact alignment results: the alignment itself (end cell, op script) instead of its rendering.

``gotoh_b200_align_batch_compact`` (include/gotoh_b200.h) returns per pair one 8-word record and a 2-bit op script -
about 100 bytes for a 251-nt read where the reference's two padded strings (gotoh.cpp:436-513, "ssi" at :650) take
6.6 KB.  ``CompactAlignments`` keeps those arrays and renders the reference's strings on demand: ``strings(k)`` replays
gotoh.cpp's output order - left overhang (:489-496), the traceback ops in reverse (:452-487 then reverse(), :512-513),
right overhang (:434-449) - over the trimmed (AA_RB: degapped) inputs.  It is a format conversion of a finished
alignment (no scoring, no DP); tests compare it byte for byte with the string entry point and the oracle.
"""
import numpy as np

from ._ffi import AA_RB

REC_SCORE, REC_OUT_LEN, REC_I0, REC_J0, REC_END_I, REC_END_J, REC_N_OPS, REC_M_N = range(8)
_WS = b" \t\n\r"
_GAP = ord("-")


class CompactAlignments:
    """Result of ``Aligner.align_packed_compact`` / ``align_batch(..., compact=True)``: sequence of
    (aligned_standard, aligned_seq, score) tuples, rendered lazily."""

    def __init__(self, ref_bytes, ref_off, ref_idx, qry_bytes, qry_off, matrix, rec, ops, ops_off):
        self.ref_bytes, self.ref_off, self.ref_idx = ref_bytes, ref_off, ref_idx
        self.qry_bytes, self.qry_off = qry_bytes, qry_off
        self.matrix = matrix
        self.rec = rec.reshape(-1, 8)
        self.ops = ops
        self.ops_off = ops_off
        self._ref_cache = {}

    def __len__(self):
        return len(self.rec)

    @property
    def scores(self):
        return self.rec[:, REC_SCORE]

    @property
    def out_len(self):
        return self.rec[:, REC_OUT_LEN]

    def nbytes(self):
        """Bytes that crossed PCIe for these results."""
        n_words = int(((self.rec[:, REC_N_OPS].astype(np.int64) + 15) >> 4).sum())
        return self.rec.nbytes + 4 * n_words

    def _clean(self, raw):
        # trim(): gotoh.cpp:545-559; degap(): gotoh.cpp:529-543 (align_it_aa_rb only)
        b = raw.tobytes().strip(_WS)
        if self.matrix == AA_RB:
            b = b.replace(b"-", b"")
        return np.frombuffer(b, dtype=np.uint8)

    def _ref(self, k):
        r = k if self.ref_idx is None else int(self.ref_idx[k])
        a = self._ref_cache.get(r)
        if a is None:
            a = self._clean(self.ref_bytes[int(self.ref_off[r]):int(self.ref_off[r + 1])])
            if len(self._ref_cache) < 4096:
                self._ref_cache[r] = a
        return a

    def op_codes(self, k):
        """The pair's ops in traceback order (from the end cell backwards): 0 diagonal, 1 up, 2 left."""
        n = int(self.rec[k, REC_N_OPS])
        w = self.ops[int(self.ops_off[k]):int(self.ops_off[k]) + ((n + 15) >> 4)]
        return ((w[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).reshape(-1)[:n].astype(np.uint8)

    def arrays(self, k):
        """(aligned_standard, aligned_seq) of pair k as uint8 arrays."""
        a = self._ref(k)
        b = self._clean(self.qry_bytes[int(self.qry_off[k]):int(self.qry_off[k + 1])])
        _, out_len, i0, j0, ei, ej, n, _ = (int(x) for x in self.rec[k])
        M, N = len(a), len(b)
        oa = np.full(out_len, _GAP, dtype=np.uint8)
        ob = np.full(out_len, _GAP, dtype=np.uint8)
        lo = max(i0, j0)                              # left overhang: the leftover prefix of exactly one sequence
        if i0 > j0:
            oa[:lo] = a[:lo]
        else:
            ob[:lo] = b[:lo]
        op = self.op_codes(k)[::-1]                   # alignment order
        ca, cb = op != 2, op != 1                     # consumes a standard / a seq character
        ia = i0 + np.cumsum(ca) - 1
        jb = j0 + np.cumsum(cb) - 1
        mid_a, mid_b = oa[lo:lo + n], ob[lo:lo + n]
        mid_a[ca] = a[ia[ca]]
        mid_b[cb] = b[jb[cb]]
        if ei == M and ej < N:                        # right overhang (gotoh.cpp:429-449)
            ob[lo + n:lo + n + N - ej] = b[ej:]
        else:
            oa[lo + n:lo + n + M - ei] = a[ei:]
        return oa, ob

    def strings(self, k):
        oa, ob = self.arrays(k)
        return oa.tobytes().decode("latin-1"), ob.tobytes().decode("latin-1")

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        oa, ob = self.strings(k)
        return oa, ob, int(self.rec[k, REC_SCORE])

    def __iter__(self):
        return (self[k] for k in range(len(self)))
