"""Host-side mirror of the reference's aligner interface, backed by libgotoh_b200.so.

Drop-in functions (same names, positional-only arguments, argument meaning, return shapes as
the CPython module ``gotoh`` of /root/reference/micall/alignment/gotoh.cpp:624-739):

    align_it(standard, seq, gap_init_penalty, gap_extend_penalty, use_terminal_gap_penalty)
        -> (aligned_standard, aligned_seq, score)                      gotoh.cpp:624-658
    align_it_aa(standard, seq, gip, gep, use_terminal_gap_penalty)
        -> (aligned_standard, aligned_seq, score)                      gotoh.cpp:660-693
    align_it_aa_rb(standard, seq, gip, gep) -> (aligned_standard, aligned_seq)   gotoh.cpp:695-727

plus the batched entry points the GPU needs to be worth using:

    align_batch(refs, queries, gip, gep, term, matrix, ref_idx=None, devices=None)
    Plan(...)  - staged form (create / run / fetch) for resident data and device timing

Everything computes on the GPU; there is no CPU path here.
"""
import ctypes

import numpy as np

from . import _ffi, packing
from ._ffi import AA_RB, HIV25, NT, GotohCapacityError, GotohError, GotohInputError  # noqa: F401
from .compact import CompactAlignments


class Aligner:
    """All entry points bound to one loaded library (the product uses ``default_library()``)."""

    def __init__(self, library=None):
        self._libobj = library or _ffi.default_library()
        self._lib = self._libobj.lib

    # ---- packed, zero-copy form -------------------------------------------------------
    def align_packed(self, ref_bytes, ref_off, ref_idx, qry_bytes, qry_off, gip, gep, term, matrix,
                     out_off=None, out=None, device_mask=1):
        """numpy in / numpy out.  Returns (out_ref, out_qry, out_off, out_len, out_score)."""
        ref_bytes = np.ascontiguousarray(ref_bytes, dtype=np.uint8)
        qry_bytes = np.ascontiguousarray(qry_bytes, dtype=np.uint8)
        ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
        ridx = None if ref_idx is None else np.ascontiguousarray(ref_idx, dtype=np.int32)
        n = len(qry_off) - 1
        if out_off is None:
            out_off = packing.out_offsets(ref_off, ridx, qry_off)
        out_off = np.ascontiguousarray(out_off, dtype=np.int64)
        if out is None:
            out_ref = np.zeros(int(out_off[-1]), dtype=np.uint8)
            out_qry = np.zeros(int(out_off[-1]), dtype=np.uint8)
            out_len = np.zeros(n, dtype=np.int32)
            out_score = np.zeros(n, dtype=np.int32)
        else:
            out_ref, out_qry, out_len, out_score = out
        rc = self._lib.gotoh_b200_align_batch(
            ref_bytes.ctypes.data, ref_off.ctypes.data, len(ref_off) - 1,
            None if ridx is None else ridx.ctypes.data,
            qry_bytes.ctypes.data, qry_off.ctypes.data, n, int(gip), int(gep), int(bool(term)), int(matrix),
            out_ref.ctypes.data, out_qry.ctypes.data, out_off.ctypes.data,
            out_len.ctypes.data, out_score.ctypes.data, int(device_mask))
        self._libobj.check(rc)
        return out_ref, out_qry, out_off, out_len, out_score

    def _packed_args(self, ref_bytes, ref_off, ref_idx, qry_bytes, qry_off):
        ref_bytes = np.ascontiguousarray(ref_bytes, dtype=np.uint8)
        qry_bytes = np.ascontiguousarray(qry_bytes, dtype=np.uint8)
        ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
        ridx = None if ref_idx is None else np.ascontiguousarray(ref_idx, dtype=np.int32)
        return ref_bytes, ref_off, ridx, qry_bytes, qry_off

    def align_packed_tight(self, ref_bytes, ref_off, ref_idx, qry_bytes, qry_off, gip, gep, term, matrix,
                           out=None, device_mask=1):
        """Tight form (gotoh_b200_align_batch_tight): the library lays the aligned strings out back to back.
        Returns (out_ref, out_qry, out_off[n], out_len, out_score); ``out`` = (out_ref, out_qry, out_off, out_len,
        out_score) arrays to fill (e.g. pinned), out_ref/out_qry of equal capacity."""
        ref_bytes, ref_off, ridx, qry_bytes, qry_off = self._packed_args(ref_bytes, ref_off, ref_idx, qry_bytes, qry_off)
        n = len(qry_off) - 1
        if out is None:
            cap = int(packing.out_offsets(ref_off, ridx, qry_off)[-1])
            out = (np.zeros(cap, np.uint8), np.zeros(cap, np.uint8), np.zeros(n, np.int64), np.zeros(n, np.int32),
                   np.zeros(n, np.int32))
        out_ref, out_qry, out_off, out_len, out_score = out
        rc = self._lib.gotoh_b200_align_batch_tight(
            ref_bytes.ctypes.data, ref_off.ctypes.data, len(ref_off) - 1, None if ridx is None else ridx.ctypes.data,
            qry_bytes.ctypes.data, qry_off.ctypes.data, n, int(gip), int(gep), int(bool(term)), int(matrix),
            out_ref.ctypes.data, out_qry.ctypes.data, min(len(out_ref), len(out_qry)), out_off.ctypes.data,
            out_len.ctypes.data, out_score.ctypes.data, int(device_mask))
        self._libobj.check(rc)
        return out_ref, out_qry, out_off, out_len, out_score

    def align_packed_compact(self, ref_bytes, ref_off, ref_idx, qry_bytes, qry_off, gip, gep, term, matrix,
                             out=None, device_mask=1):
        """Compact form (gotoh_b200_align_batch_compact): per pair a record + op script, ~1/60 of the bytes of the
        string forms.  Returns a ``CompactAlignments`` whose ``strings(k)`` / ``[k]`` render the reference's strings on
        demand.  ``out`` = (rec[n*8] int32, ops uint32, ops_off[n] int64) arrays to fill; without it the op buffer is
        sized for typical alignments and the call is repeated with the worst-case bound if it was too small."""
        ref_bytes, ref_off, ridx, qry_bytes, qry_off = self._packed_args(ref_bytes, ref_off, ref_idx, qry_bytes, qry_off)
        n = len(qry_off) - 1
        args = (ref_bytes.ctypes.data, ref_off.ctypes.data, len(ref_off) - 1, None if ridx is None else ridx.ctypes.data,
                qry_bytes.ctypes.data, qry_off.ctypes.data, n, int(gip), int(gep), int(bool(term)), int(matrix))
        if out is not None:
            rec, ops, ops_off = out
            self._libobj.check(self._lib.gotoh_b200_align_batch_compact(
                *args, rec.ctypes.data, ops.ctypes.data, len(ops), ops_off.ctypes.data, int(device_mask)))
        else:
            rlen = np.diff(ref_off)
            rl = rlen if ridx is None else rlen[ridx]
            ql = np.diff(qry_off)
            worst = int(((rl + ql + 15) >> 4).sum())
            typical = int(((np.minimum(rl, ql) * 5 // 4 + 47) >> 4).sum())      # path ~ the shorter sequence + gaps
            rec = np.zeros(n * 8, np.int32)
            ops_off = np.zeros(n, np.int64)
            for cap in sorted({min(typical, worst), worst}):
                ops = np.zeros(max(cap, 1), np.uint32)
                rc = self._lib.gotoh_b200_align_batch_compact(*args, rec.ctypes.data, ops.ctypes.data, cap,
                                                              ops_off.ctypes.data, int(device_mask))
                if rc != _ffi.ECAPACITY or cap == worst:
                    self._libobj.check(rc)
                    break
        return CompactAlignments(ref_bytes, ref_off, ridx, qry_bytes, qry_off, int(matrix), rec, ops, ops_off)

    def d2h_probe(self, host_array, reps=1, device=0):
        """Seconds for `reps` plain device-to-host copies filling host_array (gotoh_b200_d2h_probe)."""
        v = ctypes.c_double(0.0)
        self._libobj.check(self._lib.gotoh_b200_d2h_probe(int(device), host_array.ctypes.data, host_array.nbytes,
                                                          int(reps), ctypes.byref(v)))
        return v.value

    # ---- list-of-strings form ------------------------------------------------------------
    def align_batch(self, refs, queries, gip, gep, term=1, matrix=NT, ref_idx=None, devices=None, compact=False):
        """Align queries[k] against refs[ref_idx[k]] (or refs[k]; a single ref is shared by all).

        Returns a list of (aligned_ref, aligned_query, score) in input order; with ``compact=True`` a
        ``CompactAlignments`` (same items, rendered on demand from the op scripts - only ~100 B per pair leave the GPU)."""
        if isinstance(refs, (str, bytes)):
            refs = [refs]
        queries = list(queries)
        refs = list(refs)
        if ref_idx is None:
            if len(refs) == 1:
                ref_idx = np.zeros(len(queries), dtype=np.int32)
            elif len(refs) != len(queries):
                raise ValueError("need one reference, one per query, or ref_idx")
        if not queries:
            return []
        rb, ro = packing.pack(refs, "standard")
        qb, qo = packing.pack(queries, "seq")
        mask = 1
        if devices is not None:
            mask = 0
            for d in devices:
                if not 0 <= int(d) < 32:
                    raise ValueError("device index %r outside 0..31 (device_mask is 32 bits wide)" % (d,))
                mask |= 1 << int(d)
        if compact:
            return self.align_packed_compact(rb, ro, ref_idx, qb, qo, gip, gep, term, matrix, device_mask=mask)
        o_ref, o_qry, o_off, o_len, o_score = self.align_packed(rb, ro, ref_idx, qb, qo, gip, gep, term, matrix,
                                                                device_mask=mask)
        a = packing.unpack(o_ref, o_off, o_len)
        b = packing.unpack(o_qry, o_off, o_len)
        return list(zip(a, b, o_score.tolist()))

    # ---- the reference's three callables ------------------------------------------------
    def _one(self, standard, seq, gip, gep, term, matrix):
        for name, v in (("gap_init_penalty", gip), ("gap_extend_penalty", gep), ("use_terminal_gap_penalty", term)):
            if isinstance(v, float) or not isinstance(v, (int, bool, np.integer)):
                raise TypeError("%s must be an integer" % name)   # PyArg "i" (gotoh.cpp:633)
        return self.align_batch([standard], [seq], int(gip), int(gep), int(term), matrix)[0]

    def align_it(self, standard, seq, gap_init_penalty, gap_extend_penalty, use_terminal_gap_penalty, /):
        return self._one(standard, seq, gap_init_penalty, gap_extend_penalty, use_terminal_gap_penalty, NT)

    def align_it_aa(self, standard, seq, gap_init_penalty, gap_extend_penalty, use_terminal_gap_penalty, /):
        return self._one(standard, seq, gap_init_penalty, gap_extend_penalty, use_terminal_gap_penalty, HIV25)

    def align_it_aa_rb(self, standard, seq, gap_init_penalty, gap_extend_penalty, /):
        return self._one(standard, seq, gap_init_penalty, gap_extend_penalty, 0, AA_RB)[:2]

    # ---- misc -------------------------------------------------------------------------------
    def pairscore_table(self, matrix):
        t = np.zeros(127 * 127, dtype=np.int32)
        self._libobj.check(self._lib.gotoh_b200_pairscore_table(int(matrix), t.ctypes.data))
        return t.reshape(127, 127)

    def device_count(self):
        return self._libobj.device_count()

    def int_peak(self, which, device=0):
        v = ctypes.c_double(0.0)
        self._libobj.check(self._lib.gotoh_b200_int_peak(int(device), int(which), ctypes.byref(v)))
        return v.value

    def plan(self, *args, **kw):
        return Plan(self, *args, **kw)


class PinnedArray:
    """numpy view over page-locked host memory from gotoh_b200_host_alloc."""

    def __init__(self, aligner, shape, dtype):
        self._lib = aligner._lib
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self._ptr = self._lib.gotoh_b200_host_alloc(max(n, 1))
        if not self._ptr:
            raise MemoryError("gotoh_b200_host_alloc(%d) failed" % n)
        buf = (ctypes.c_uint8 * max(n, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            self._lib.gotoh_b200_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Plan:
    """Staged batch: ``Plan(...)`` validates, packs and uploads; ``run()`` aligns on the device and
    returns (device_ms, forward_ms); ``fetch()`` copies the results back."""

    def __init__(self, aligner, ref_bytes, ref_off, ref_idx, qry_bytes, qry_off, gip, gep, term, matrix,
                 device=0, out_off=None):
        self._a = aligner
        self._lib = aligner._lib
        self.ref_bytes = np.ascontiguousarray(ref_bytes, dtype=np.uint8)
        self.qry_bytes = np.ascontiguousarray(qry_bytes, dtype=np.uint8)
        self.ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
        self.qry_off = np.ascontiguousarray(qry_off, dtype=np.int64)
        self.ref_idx = None if ref_idx is None else np.ascontiguousarray(ref_idx, dtype=np.int32)
        self.n = len(self.qry_off) - 1
        self.out_off = np.ascontiguousarray(
            packing.out_offsets(self.ref_off, self.ref_idx, self.qry_off) if out_off is None else out_off,
            dtype=np.int64)
        h = ctypes.c_void_p(None)
        rc = self._lib.gotoh_b200_plan_create(
            int(device), self.ref_bytes.ctypes.data, self.ref_off.ctypes.data, len(self.ref_off) - 1,
            None if self.ref_idx is None else self.ref_idx.ctypes.data, self.qry_bytes.ctypes.data,
            self.qry_off.ctypes.data, self.n, int(gip), int(gep), int(bool(term)), int(matrix),
            self.out_off.ctypes.data, ctypes.byref(h))
        aligner._libobj.check(rc)
        self._h = h

    def run(self, time_forward=True):
        d, f = ctypes.c_float(0.0), ctypes.c_float(0.0)
        rc = self._lib.gotoh_b200_plan_run(self._h, ctypes.byref(d), ctypes.byref(f) if time_forward else None)
        self._a._libobj.check(rc)
        return d.value, f.value

    def fetch(self, out=None):
        if out is None:
            out = (np.zeros(int(self.out_off[-1]), dtype=np.uint8), np.zeros(int(self.out_off[-1]), dtype=np.uint8),
                   np.zeros(self.n, dtype=np.int32), np.zeros(self.n, dtype=np.int32))
        rc = self._lib.gotoh_b200_plan_fetch(self._h, out[0].ctypes.data, out[1].ctypes.data,
                                             out[2].ctypes.data, out[3].ctypes.data)
        self._a._libobj.check(rc)
        return out

    def stat(self, what):
        return int(self._lib.gotoh_b200_plan_stat(self._h, int(what)))

    @property
    def cells(self):
        return self.stat(0)

    def close(self):
        if self._h:
            self._lib.gotoh_b200_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
