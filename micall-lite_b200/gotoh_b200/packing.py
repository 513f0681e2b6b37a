"""Pack Python sequences into the contiguous byte + offset arrays the C ABI takes."""
import codecs

import numpy as np


def to_bytes(s, what="sequence"):
    """str/bytes -> bytes with the reference's "s" argument rules (gotoh.cpp:633):
    str is UTF-8 encoded, embedded NUL is rejected, anything else is a TypeError."""
    if isinstance(s, str):
        b = s.encode("utf-8")
    elif isinstance(s, (bytes, bytearray, memoryview)):
        b = bytes(s)
    else:
        raise TypeError("%s must be str, not %s" % (what, type(s).__name__))
    if b"\0" in b:
        raise ValueError("embedded null character")
    return b


def pack(seqs, what="sequence", ascii_only=False):
    """list of str/bytes -> (uint8 array, int64 offsets of len n+1).  ``ascii_only``: raise UnicodeError instead of
    encoding non-ASCII text as UTF-8 (callers that treat such text differently)."""
    seqs = list(seqs)
    n = len(seqs)
    off = np.zeros(n + 1, dtype=np.int64)
    if n == 0:
        return np.zeros(0, dtype=np.uint8), off
    # fast path (hundreds of thousands of short ASCII strings): one join, one encode, lengths from len() - the per-string
    # encode / NUL check of to_bytes() costs ~0.5 us each, more than the GPU spends on an 84-aa window
    if set(map(type, seqs)) == {str}:                   # exactly str, checked without a Python frame per element
        joined = "".join(seqs)
        if joined.isascii():
            if "\0" in joined:
                raise ValueError("embedded null character")
            np.cumsum(np.fromiter(map(len, seqs), dtype=np.int64, count=n), out=off[1:])
            return np.frombuffer(joined.encode("ascii"), dtype=np.uint8), off
        if ascii_only:
            raise UnicodeError("non-ASCII text")
    elif ascii_only and any(isinstance(s, str) and not s.isascii() for s in seqs):
        raise UnicodeError("non-ASCII text")
    bs = [to_bytes(s, what) for s in seqs]
    np.cumsum([len(b) for b in bs], out=off[1:])
    data = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(data), off


def out_offsets(ref_off, ref_idx, qry_off):
    """Output offsets with stride len(ref)+len(query) per pair (>= M+N after trimming)."""
    rlen = np.diff(ref_off)
    qlen = np.diff(qry_off)
    per = (rlen if ref_idx is None else rlen[ref_idx]) + qlen
    off = np.zeros(len(qlen) + 1, dtype=np.int64)
    np.cumsum(per, out=off[1:])
    return off


def unpack(out, off, lens):
    """Packed outputs -> list of str.  One decode of the whole buffer, then str slices over plain-int offsets: numpy
    scalars and a per-element decode cost ~2 us per string, ten times what the slicing does."""
    text = codecs.latin_1_decode(np.ascontiguousarray(out))[0]          # straight from the array's buffer, no bytes copy
    n = len(lens)
    starts = np.asarray(off[:n], dtype=np.int64)
    return [text[a:b] for a, b in zip(starts.tolist(), (starts + np.asarray(lens, dtype=np.int64)).tolist())]


class PackedStrings:
    """Read-only sequence of str over packed bytes: element k is decoded when it is asked for."""

    def __init__(self, out, off, lens):
        self._buf, self._off, self._len = out, off, lens

    def __len__(self):
        return len(self._len)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        o = int(self._off[k])
        return self._buf[o:o + int(self._len[k])].tobytes().decode("latin-1")
