"""Oracle for the live aligner `gotoh2.Aligner.align` / `_gotoh2.align`.
TEST INFRASTRUCTURE ONLY - see oracle.py for who may import this.

``Oracle2('port')``      C restatement oracle/gotoh2_oracle.c
``Oracle2('reference')`` the reference's own _gotoh2.c built as the CPython extension it is
                         (oracle/_ref/_gotoh2*.so), called exactly like gotoh2.py:87-95 does.
"""
import ctypes
import glob
import importlib.util
import os

from . import oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
MODELS_DIR = os.path.join(os.path.dirname(_HERE), "micall-lite_b200", "gotoh_b200", "data", "models")


def read_model(path):
    """header = alphabet, rows = integer scores (gotoh2.py:47-64)."""
    with open(path) as f:
        alphabet = "".join(next(f).strip("\n").split(","))
        rows = []
        for line in f:
            if line.strip():
                rows.extend(int(x) for x in line.strip("\n").split(","))
    return rows, alphabet


def load_models():
    return {os.path.basename(p)[:-4]: read_model(p) for p in sorted(glob.glob(os.path.join(MODELS_DIR, "*.csv")))}


def clean_sequence(seq, alphabet):
    """gotoh2.py:70-72: upper-case, every non-alphabet character becomes '?'."""
    return "".join(c if c in alphabet else "?" for c in seq.upper())


def have_reference():
    return bool(glob.glob(os.path.join(_HERE, "_ref", "_gotoh2*.so")))


class Oracle2:
    def __init__(self, kind="port"):
        self.kind = kind
        self.models = load_models()
        if kind == "port":
            if not os.path.exists(_o.PORT_SO):
                _o.build()
            self.lib = ctypes.CDLL(_o.PORT_SO)
            self.lib.gotoh2_oracle_align.restype = ctypes.c_int
            self.lib.gotoh2_oracle_align.argtypes = [
                ctypes.c_char_p, ctypes.c_long, ctypes.c_char_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                ctypes.c_char_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p,
                ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        elif kind == "reference":
            so = glob.glob(os.path.join(_HERE, "_ref", "_gotoh2*.so"))
            if not so:
                raise FileNotFoundError("oracle/_ref/_gotoh2*.so (build where /root/reference exists: make -C oracle)")
            spec = importlib.util.spec_from_file_location("_gotoh2", so[0])
            self.mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(self.mod)
        else:
            raise ValueError(kind)

    def align(self, seq1, seq2, gop=10, gep=1, is_global=False, model="HYPHY_NUC"):
        """Same contract as gotoh2.Aligner(gop, gep, is_global, model).align(seq1, seq2)."""
        assert type(seq1) is str and type(seq2) is str and len(seq1) > 0 and len(seq2) > 0   # gotoh2.py:82-85
        matrix, alphabet = self.models[model]
        c1, c2 = clean_sequence(seq1, alphabet), clean_sequence(seq2, alphabet)
        if self.kind == "reference":
            return self.mod.align(c1, c2, gop, gep, int(is_global), alphabet, matrix)      # gotoh2.py:87-95
        a, b = c1.encode("ascii"), c2.encode("ascii")
        o1 = ctypes.create_string_buffer(len(a) + len(b) + 1)
        o2 = ctypes.create_string_buffer(len(a) + len(b) + 1)
        ln, sc = ctypes.c_int(0), ctypes.c_int(0)
        d = (ctypes.c_int * len(matrix))(*matrix)
        rc = self.lib.gotoh2_oracle_align(a, len(a), b, len(b), gop, gep, int(is_global), alphabet.encode("ascii"), d,
                                          o1, o2, ctypes.byref(ln), ctypes.byref(sc))
        if rc == -4:
            raise RuntimeError("Traceback failed, try local alignment")                    # _gotoh2.c:601-603
        if rc:
            raise ValueError("gotoh2 oracle rejected input (code %d)" % rc)
        return (o1.raw[:ln.value].decode("ascii"), o2.raw[:ln.value].decode("ascii"), sc.value)


def levenshtein(a, b):
    """Oracle for Levenshtein.distance(a, b) (remap.py:250): oracle/levenshtein_oracle.c."""
    if not os.path.exists(_o.PORT_SO):
        _o.build()
    lib = ctypes.CDLL(_o.PORT_SO)
    if not hasattr(lib, "levenshtein_oracle"):
        _o.build()
        lib = ctypes.CDLL(_o.PORT_SO)
    lib.levenshtein_oracle.restype = ctypes.c_long
    lib.levenshtein_oracle.argtypes = [ctypes.c_char_p, ctypes.c_long, ctypes.c_char_p, ctypes.c_long]
    ab, bb = (a.encode("latin-1") if isinstance(a, str) else bytes(a)), (b.encode("latin-1") if isinstance(b, str) else bytes(b))
    return int(lib.levenshtein_oracle(ab, len(ab), bb, len(bb)))
