"""nwchem_b200 -- B200-native CCSD(T) perturbative triples behind NWChem TCE's sd_t_* call surface.

The product is the C-ABI library nwchem_b200/lib/libnwc_triples.so (include/nwc_triples.h);
this package holds the host-side mirror of the reference driver interface plus the tile/offset
tables and synthetic-input generators that the tests and bench.py need.
"""
from . import tiling, synth  # noqa: F401
