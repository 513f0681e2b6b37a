"""GPU-backed mirror of the reference's live aligner class, ``micall.alignment.gotoh2.Aligner``
(gotoh2.py:7-96), the one ``core/remap.py:33,248`` and ``core/aln2counts.py:34-37,187`` call.

Same constructor, attributes and ``align(seq1, seq2) -> (aligned1, aligned2, score)`` contract;
``align_batch`` is the batched form.  Computation happens in libgotoh_b200.so
(gotoh_b200_gotoh2_align_batch); there is no CPU path.
"""
import glob
import os
import re

import numpy as np

from . import _ffi, packing

_MODELS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "models")


def read_matrix_from_csv(handle):
    """Header = alphabet, rows = integer scores (gotoh2.py:47-64)."""
    header = next(handle)
    alphabet = "".join(header.strip("\n").split(","))
    rows = []
    for line in handle:
        if line.strip():
            rows.extend(int(x) for x in line.strip("\n").split(","))
    return rows, alphabet


class Aligner:
    def __init__(self, gop=10, gep=1, is_global=False, model="HYPHY_NUC", library=None, device=0):
        self.gap_open_penalty = gop
        self.gap_extend_penalty = gep
        self.is_global = is_global
        self.device = device
        self._libobj = library or _ffi.default_library()
        self.models = {}
        for path in sorted(glob.glob(os.path.join(_MODELS_DIR, "*.csv"))):     # gotoh2.py:23-33
            with open(path) as handle:
                self.models[os.path.basename(path)[:-4]] = read_matrix_from_csv(handle)
        self.set_model(model)

    def __str__(self):
        return "%s\n%s\nGap open penalty: %s\nGap extend penalty: %s\n" % (
            self.alphabet, self.matrix, self.gap_open_penalty, self.gap_extend_penalty)

    def set_model(self, model):
        if model in self.models:
            self.matrix, self.alphabet = self.models[model]
        else:
            print("ERROR: Unrecognized model name {}".format(model))            # gotoh2.py:66-68

    def clean_sequence(self, seq):
        # replace all non-alphabet characters with ambiguous symbol (gotoh2.py:70-72)
        return re.sub(pattern="[^%s]" % (self.alphabet,), repl="?", string=seq.upper())

    def align(self, seq1, seq2):
        assert type(seq1) is str, "seq1 must be a string"                        # gotoh2.py:82-85
        assert type(seq2) is str, "seq2 must be a string"
        assert len(seq1) > 0, "seq1 cannot be an empty string"
        assert len(seq2) > 0, "seq2 cannot be an empty string"
        return self.align_batch([(seq1, seq2)])[0]

    def align_batch(self, pairs, lazy=False):
        """[(seq1, seq2), ...] -> [(aligned1, aligned2, score), ...]; raises RuntimeError like
        _gotoh2.c:601-603 when a traceback fails.  ``lazy=True`` returns a sequence that builds each tuple when it is
        asked for (hundreds of thousands of result tuples cost more Python time than the GPU spends aligning)."""
        pairs = list(pairs)
        if not pairs:
            return []
        # identical first sequences (one reference against many reads) are packed once and shared through s1_idx
        # (dict.fromkeys / map keep the loops over hundreds of thousands of pairs inside the interpreter's C code)
        firsts = [p[0] for p in pairs]
        s2 = [p[1] for p in pairs]
        uniq = {s: k for k, s in enumerate(dict.fromkeys(firsts))}
        s1_idx = np.fromiter(map(uniq.__getitem__, firsts), np.int32, count=len(pairs))
        # the library cleans the bytes itself (ASCII upper-case, non-alphabet -> '?', gotoh2.py:70-72); only non-ASCII text
        # goes through the reference's regular expression here, because it works on characters, not on UTF-8 bytes
        s1 = [a if a.isascii() else self.clean_sequence(a) for a in uniq]
        try:
            b2, o2 = packing.pack(s2, "seq2", ascii_only=True)
        except UnicodeError:
            b2, o2 = packing.pack([b if b.isascii() else self.clean_sequence(b) for b in s2], "seq2")
        b1, o1 = packing.pack(s1, "seq1")
        n = len(pairs)
        out_off = packing.out_offsets(o1, s1_idx, o2)
        out1 = np.zeros(int(out_off[-1]), np.uint8)
        out2 = np.zeros(int(out_off[-1]), np.uint8)
        out_len = np.zeros(n, np.int32)
        out_score = np.zeros(n, np.int32)
        mat = np.ascontiguousarray(self.matrix, dtype=np.int32)
        rc = self._libobj.lib.gotoh_b200_gotoh2_align_batch(
            b1.ctypes.data, o1.ctypes.data, len(s1), s1_idx.ctypes.data, b2.ctypes.data, o2.ctypes.data, n,
            int(self.gap_open_penalty), int(self.gap_extend_penalty), int(bool(self.is_global)),
            self.alphabet.encode("ascii"), mat.ctypes.data, out1.ctypes.data, out2.ctypes.data,
            out_off.ctypes.data, out_len.ctypes.data, out_score.ctypes.data, int(self.device))
        if rc == _ffi.ETRACEBACK:
            raise RuntimeError("Traceback failed, try local alignment")
        self._libobj.check(rc)
        if lazy:
            return LazyAlignments(packing.PackedStrings(out1, out_off, out_len), packing.PackedStrings(out2, out_off, out_len), out_score)
        a = packing.unpack(out1, out_off, out_len)
        b = packing.unpack(out2, out_off, out_len)
        return list(zip(a, b, out_score.tolist()))


class LazyAlignments:
    """Sequence of (aligned1, aligned2, score) over the packed result arrays."""

    def __init__(self, a, b, score):
        self._a, self._b, self.scores = a, b, score

    def __len__(self):
        return len(self.scores)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        return self._a[k], self._b[k], int(self.scores[k])

    def __iter__(self):
        return (self[k] for k in range(len(self)))
