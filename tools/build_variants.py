"""Builds instrumented variants of the library under tools/_trace/ (never the product):
    python tools/build_variants.py NAME DEFINE [DEFINE ...]   ->  tools/_trace/libicebin_b200_NAME.so
Use with ICEBIN_B200_LIB=tools/_trace/libicebin_b200_NAME.so python tools/<probe>.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icebin_b200 import build
name, defs = sys.argv[1], tuple(sys.argv[2:])
d = os.path.join(ROOT, "tools", "_trace")
os.makedirs(d, exist_ok=True)
print(build.build(force=True, defines=defs, out=os.path.join(d, "libicebin_b200_%s.so" % name), objdir=os.path.join(d, "obj_" + name)))
