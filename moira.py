#!/usr/bin/env python3
"""`moira.py` on the B200 path.  As a script: the reference's command line (moira/moira.py:1740-1741), same flags, see
moira_b200/cli.py.  As a module (`import moira`): the reference's functions -- calculate_errors_PB, calculate_errors_poisson,
reverse_complement, nw_align, make_contig, process_data, parse_arguments, main, the exception classes -- see
moira_b200/reference_api.py; the reference's own test file reads the same against it (tests/test_moira_module.py)."""
import sys

from moira_b200.cli import run
from moira_b200.reference_api import *  # noqa: F401,F403
from moira_b200.reference_api import __all__  # noqa: F401

if __name__ == "__main__":
    sys.exit(run())
