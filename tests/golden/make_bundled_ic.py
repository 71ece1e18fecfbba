"""tests/golden/bundled_ic_bs11.npz: coordinates and types of the reference's four bundled box-11 initial conditions
(VisitorGeneMaps/input_files/IC_ser5P370_RBP550_0RNP_promoter{1,2,3,4}_bs11.data, velocities are zero).  The reference's
launcher starts EVERY repeat of a promoter length from that one file (run_single_GENES_STAGES.py reads
input_files/IC_..._promoter<P>_bs<B>.data), so an ensemble that is to be compared with the reference's tables cell by
cell starts from the same configuration.  Run here (needs /root/reference): python tests/golden/make_bundled_ic.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from enhancershoebox_b200 import datafile

out = {}
for p in (1, 2, 3, 4):
    d = datafile.read_data(f"/root/reference/VisitorGeneMaps/input_files/IC_ser5P370_RBP550_0RNP_promoter{p}_bs11.data")
    g = datafile.make_shoebox(11, p, False, seed=0)
    assert np.array_equal(d.types, g.types) and np.array_equal(d.bonds, g.bonds) and np.array_equal(d.angles, g.angles) and np.all(d.v == 0)
    out[f"x{p}"] = d.x.astype(np.float64)
    out[f"t{p}"] = d.types.astype(np.int8)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bundled_ic_bs11.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
