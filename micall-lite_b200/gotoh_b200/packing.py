"""Pack Python sequences into the contiguous byte + offset arrays the C ABI takes."""
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


def pack(seqs, what="sequence"):
    """list of str/bytes -> (uint8 array, int64 offsets of len n+1)."""
    bs = [to_bytes(s, what) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
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
    """Packed outputs -> list of str."""
    buf = out.tobytes()
    return [buf[int(off[k]):int(off[k]) + int(lens[k])].decode("latin-1") for k in range(len(lens))]
