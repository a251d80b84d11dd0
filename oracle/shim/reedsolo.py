"""Stand-in for the absent third-party ``reedsolo`` package so that the
untouched reference (/root/reference/src/libfrad) can be imported in the build container
by oracle/make_golden.py.  Backed by the oracle's RS restatement.  TEST INFRASTRUCTURE."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.rs import RSCodec, ReedSolomonError  # noqa: E402,F401
