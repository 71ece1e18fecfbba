#!/usr/bin/env python
"""Drop-in for the reference's run_single_shoebox.py (same flags, inputs and outputs), GPU-backed.
See enhancershoebox_b200/drivers.py."""
import sys
from enhancershoebox_b200.drivers import main_shoebox

if __name__ == "__main__":
    sys.exit(main_shoebox())
