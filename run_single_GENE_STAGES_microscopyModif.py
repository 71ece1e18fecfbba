#!/usr/bin/env python
"""Drop-in for the reference's run_single_GENE_STAGES_microscopyModif.py: the GENES_STAGES driver with make_microscopy = 1
(same flags, inputs and outputs; microscopy_files/synthetic_microscopy_<time>.npy hold the (C, Z, Y, X) stacks the
reference writes as OME-TIFF).  See enhancershoebox_b200/drivers.py."""
import sys
from enhancershoebox_b200.drivers import main_genes_stages

if __name__ == "__main__":
    sys.exit(main_genes_stages(make_microscopy=True))
