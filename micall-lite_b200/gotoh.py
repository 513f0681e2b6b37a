"""Module named like the reference's extension (``from gotoh import align_it``,
micall/utils/reference_distances.py:6; ``#import gotoh`` aln2counts.py:22) so existing call
sites work unchanged once micall-lite_b200/ is on sys.path."""
from gotoh_b200 import align_it, align_it_aa, align_it_aa_rb  # noqa: F401
