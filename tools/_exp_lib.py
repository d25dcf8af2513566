import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from pino_locoman_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "pino_locoman_b200", sys.argv[1])
sys.argv = [sys.argv[0]] + sys.argv[2:]
exec(open(os.path.join(ROOT, "tools", "prof_sqp.py")).read())
