#!/usr/bin/env python3
"""Run tools/prof_sqp.py against another build of the library: prof_with_lib.py <lib file name> [prof_sqp args]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pino_locoman_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "pino_locoman_b200", sys.argv[1])
sys.argv = [sys.argv[0]] + sys.argv[2:]
exec(open(os.path.join(ROOT, "tools", "prof_sqp.py")).read())
