#!/usr/bin/env python3
"""`moira.py` entry point (reference: moira/moira.py:1740-1741) on the B200 path: same flags, see moira_b200/cli.py."""
import sys

from moira_b200.cli import run

if __name__ == "__main__":
    sys.exit(run())
