#!/usr/bin/env python
"""Drop-in for the reference's VisitorGeneMaps/run_single_GENES_STAGES.py (same flags, inputs and
outputs), GPU-backed.  See enhancershoebox_b200/drivers.py."""
import sys
from enhancershoebox_b200.drivers import main_genes_stages

if __name__ == "__main__":
    sys.exit(main_genes_stages())
