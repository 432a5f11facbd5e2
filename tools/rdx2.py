"""Kept for the golden-vector scripts: the RDX2 reader lives in the package (hibag_b200/rdx2.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hibag_b200.rdx2 import *  # noqa: F401,F403
from hibag_b200.rdx2 import RObj, load, NA_INTEGER  # noqa: F401
