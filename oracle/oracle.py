"""ctypes loaders for the parity oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (micall-lite_b200/gotoh_b200)
never does: it has no CPU path.

Two checkers are exposed with the same Python signature:

* ``Oracle('port')``      - oracle/_build/libgotoh_oracle.so, our C restatement
                            (gotoh_oracle.c) of /root/reference/micall/alignment/gotoh.cpp.
* ``Oracle('reference')`` - oracle/_ref/libgotoh_ref.so, the reference's own gotoh.cpp
                            compiled unmodified (oracle/ref_wrapper.cpp, oracle/Makefile).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_build", "libgotoh_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgotoh_ref.so")

NT, HIV25, AA_RB = 0, 1, 2


def build(quiet=True):
    """Compile the restatement and, when /root/reference is present, the reference."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def have_reference():
    return os.path.exists(REF_SO)


_c_int_p = ctypes.POINTER(ctypes.c_int)
_BATCH_ARGS = [
    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
    ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
    ctypes.c_void_p]


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            if not os.path.exists(PORT_SO):
                build()
            lib = ctypes.CDLL(PORT_SO)
            lib.gotoh_oracle_align.restype = ctypes.c_int
            lib.gotoh_oracle_align.argtypes = [
                ctypes.c_int, ctypes.c_char_p, ctypes.c_long, ctypes.c_char_p, ctypes.c_long,
                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p,
                _c_int_p, _c_int_p]
            lib.gotoh_oracle_align_batch.restype = ctypes.c_int
            lib.gotoh_oracle_align_batch.argtypes = _BATCH_ARGS
            lib.gotoh_oracle_table.argtypes = [ctypes.c_int, ctypes.c_void_p]
        elif kind == "reference":
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " (build it where /root/reference exists: make -C oracle)")
            lib = ctypes.CDLL(REF_SO)
            lib.ref_align.restype = ctypes.c_int
            lib.ref_align.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p]
            lib.ref_align_batch.restype = None
            lib.ref_align_batch.argtypes = _BATCH_ARGS
            lib.ref_pairscore_table.argtypes = [ctypes.c_int, ctypes.c_void_p]
        else:
            raise ValueError(kind)
        self.lib = lib

    # -- single pair, wrapper semantics (gotoh.cpp:624-727) -----------------
    def align(self, matrix_id, standard, seq, gip, gep, term=1):
        a = standard.encode("latin-1") if isinstance(standard, str) else bytes(standard)
        b = seq.encode("latin-1") if isinstance(seq, str) else bytes(seq)
        oa = ctypes.create_string_buffer(len(a) + len(b) + 2)
        ob = ctypes.create_string_buffer(len(a) + len(b) + 2)
        if self.kind == "port":
            ln, sc = ctypes.c_int(0), ctypes.c_int(0)
            rc = self.lib.gotoh_oracle_align(matrix_id, a, len(a), b, len(b), gip, gep, int(term),
                                             oa, ob, ctypes.byref(ln), ctypes.byref(sc))
            if rc:
                raise ValueError("oracle rejected input (code %d)" % rc)
            return (oa.raw[:ln.value].decode("latin-1"), ob.raw[:ln.value].decode("latin-1"), sc.value)
        if b"\0" in a or b"\0" in b:
            raise ValueError("NUL byte")
        sc = self.lib.ref_align(matrix_id, a, b, gip, gep, int(term), oa, ob)
        return (oa.value.decode("latin-1"), ob.value.decode("latin-1"), sc)

    def align_it(self, standard, seq, gip, gep, term):
        return self.align(NT, standard, seq, gip, gep, term)

    def align_it_aa(self, standard, seq, gip, gep, term):
        return self.align(HIV25, standard, seq, gip, gep, term)

    def align_it_aa_rb(self, standard, seq, gip, gep):
        return self.align(AA_RB, standard, seq, gip, gep, 0)[:2]

    # -- packed batch (same layout as the product's C-ABI) -------------------
    def align_batch(self, matrix_id, ref_bytes, ref_off, ref_idx, qry_bytes, qry_off, gip, gep, term,
                    first=0, last=None, out=None):
        """Arrays: uint8 bytes, int64 offsets, int32 ref_idx (or None).  Returns
        (out_ref, out_qry, out_off, out_len, out_score) numpy arrays."""
        n = len(qry_off) - 1
        last = n if last is None else last
        ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
        rlen = np.diff(ref_off)
        qlen = np.diff(qry_off)
        ridx = None if ref_idx is None else np.ascontiguousarray(ref_idx, dtype=np.int32)
        per = (rlen if ridx is None else rlen[ridx]) + qlen
        if out is None:
            out_off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(per, out=out_off[1:])
            out_a = np.zeros(int(out_off[-1]), dtype=np.uint8)
            out_b = np.zeros(int(out_off[-1]), dtype=np.uint8)
            out_len = np.zeros(n, dtype=np.int32)
            out_score = np.zeros(n, dtype=np.int32)
        else:
            out_a, out_b, out_off, out_len, out_score = out
        args = (matrix_id, ref_bytes.ctypes.data, ref_off.ctypes.data,
                None if ridx is None else ridx.ctypes.data, qry_bytes.ctypes.data, qry_off.ctypes.data,
                first, last, gip, gep, int(term), out_a.ctypes.data, out_b.ctypes.data,
                out_off.ctypes.data, out_len.ctypes.data, out_score.ctypes.data)
        if self.kind == "port":
            rc = self.lib.gotoh_oracle_align_batch(*args)
            if rc:
                raise ValueError("oracle rejected input (code %d)" % rc)
        else:
            self.lib.ref_align_batch(*args)
        return out_a, out_b, out_off, out_len, out_score

    def table(self, matrix_id):
        t = np.zeros(127 * 127, dtype=np.int32)
        if self.kind == "port":
            self.lib.gotoh_oracle_table(matrix_id, t.ctypes.data)
        else:
            self.lib.ref_pairscore_table(matrix_id, t.ctypes.data)
        return t.reshape(127, 127)
