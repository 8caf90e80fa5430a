"""Builds tools/_trace/libicebin_b200_trace.so: the product sources with -DIB_STREAM_TRACE (per-CTA time stamps in
spmm_stream_kernel).  Use with ICEBIN_B200_LIB=tools/_trace/libicebin_b200_trace.so python tools/stream_trace.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icebin_b200 import build
d = os.path.join(ROOT, "tools", "_trace")
os.makedirs(d, exist_ok=True)
print(build.build(force=True, defines=("IB_STREAM_TRACE",), out=os.path.join(d, "libicebin_b200_trace.so"), objdir=os.path.join(d, "obj")))
