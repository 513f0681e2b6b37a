"""gotoh_b200 - B200-native drop-in for MiCall-Lite's Gotoh aligner (align_it / align_it_aa).

``from gotoh_b200 import align_it, align_it_aa, align_it_aa_rb`` mirrors the reference's
``from gotoh import align_it`` (micall/utils/reference_distances.py:6,37); ``align_batch`` is
the batched form.  All of them run on the GPU through libgotoh_b200.so; importing this package
never loads the CPU oracle and there is no CPU fallback.
"""
from . import _ffi
from ._ffi import AA_RB, HIV25, NT, GotohCapacityError, GotohError, GotohInputError  # noqa: F401
from .api import Aligner, PinnedArray, Plan  # noqa: F401

_aligner = None


def _get():
    global _aligner
    if _aligner is None:
        _aligner = Aligner()
    return _aligner


def align_it(standard, seq, gap_init_penalty, gap_extend_penalty, use_terminal_gap_penalty, /):
    """gotoh.cpp:624-658 - nucleotides, init_pairscore(5,4)."""
    return _get().align_it(standard, seq, gap_init_penalty, gap_extend_penalty, use_terminal_gap_penalty)


def align_it_aa(standard, seq, gap_init_penalty, gap_extend_penalty, use_terminal_gap_penalty, /):
    """gotoh.cpp:660-693 - amino acids, empirical HIV 25% matrix."""
    return _get().align_it_aa(standard, seq, gap_init_penalty, gap_extend_penalty, use_terminal_gap_penalty)


def align_it_aa_rb(standard, seq, gap_init_penalty, gap_extend_penalty, /):
    """gotoh.cpp:695-727 - amino acids, ReCall settings (+4/+2, degap, terminal gaps forgiven)."""
    return _get().align_it_aa_rb(standard, seq, gap_init_penalty, gap_extend_penalty)


def align_batch(refs, queries, gip, gep, term=1, matrix=NT, ref_idx=None, devices=None, compact=False):
    return _get().align_batch(refs, queries, gip, gep, term, matrix, ref_idx, devices, compact)


def device_count():
    return _get().device_count()
