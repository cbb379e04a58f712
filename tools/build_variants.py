"""Builds experiment variants of the library (tools/kbench.py runs them via AEROBULK_GPU_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aerobulk_b200 import build as B
VARIANTS = {
    "smemtab": ["ABM_SMEM_TABLES=1"],
    "b128x6": ["AB_FLUX_BLOCK=128", "AB_MIN_BLOCKS=6"],
    "b128x7": ["AB_FLUX_BLOCK=128", "AB_MIN_BLOCKS=7"],
    "b128x8": ["AB_FLUX_BLOCK=128", "AB_MIN_BLOCKS=8"],
    "b256x4": ["AB_FLUX_BLOCK=256", "AB_MIN_BLOCKS=4"],
    "pass2": ["AB_PASS_UNROLL=2"],
    "sband035": ["AB_SORT_BAND_SKIN=0.35"],
    "ns3": ["AB_MIN_BLOCKS_NOSKIN=3"],
    "ns5": ["AB_MIN_BLOCKS_NOSKIN=5"],
    "b256x2": ["AB_FLUX_BLOCK=256", "AB_MIN_BLOCKS=2"],
    "b128x5": ["AB_FLUX_BLOCK=128", "AB_MIN_BLOCKS=5"],
    "cs2": ["AB_CS_UNROLL=2"],
    "cs5": ["AB_CS_UNROLL=5"],
    "band0": ["AB_SORT_BAND=0.0", "AB_SORT_BAND_SKIN=0.0", "AB_SORT_BAND_NOQ=0.0"],
    "sband075": ["AB_SORT_BAND_SKIN=0.75"],
    "sband1": ["AB_SORT_BAND_SKIN=1.0"],
    "band01": ["AB_SORT_BAND=0.1"],
    "band02": ["AB_SORT_BAND=0.2"],
    "ser64x1": ["AB_SERIES_MIN_BLOCKS=1"],
    "ser64x8": ["AB_SERIES_MIN_BLOCKS=8"],
    "ser128x6": ["AB_SERIES_BLOCK=128", "AB_SERIES_MIN_BLOCKS=6"],
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    for n in names:
        print(n, B.build(defines=VARIANTS[n], tag=n))
