#!/usr/bin/env python
"""Launcher for the batch tools under torchrun (the package directory `r-pcc_b200/` is not an importable name, the
alias module at the repo root is):  torchrun ... scripts/datalist_entry.py compress|decompress <tool flags>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpcc_b200.tools import compress_datalist, decompress_datalist  # noqa: E402

if __name__ == "__main__":
    which, argv = sys.argv[1], sys.argv[2:]
    (compress_datalist.main if which == "compress" else decompress_datalist.main)(argv)
