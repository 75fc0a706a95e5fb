"""Shim: the loader of the unmodified reference env modules lives in oracle/ref_loader.py (test infrastructure)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_loader import *  # noqa: E402,F401,F403
from oracle.ref_loader import NoiseQueue, REFERENCE_SRC, load_reference_envs, reference_available  # noqa: E402,F401
