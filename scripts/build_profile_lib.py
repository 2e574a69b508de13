"""Debug build: lib/libnfb200_prof.so = the library with made_chain_bf16.cu compiled -DNF_MC_PROFILE (per-role wait-cycle
counters of CTA 0, read with nf_debug_mc_profile).  Uses the objects of the normal build for everything else.
    python normalizing-flows-study_b200/build.py && python scripts/build_profile_lib.py"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "normalizing-flows-study_b200")
sys.path.insert(0, PKG)
import build as B  # noqa: E402

obj = os.path.join(B.OBJ, "made_chain_bf16_prof.o")
subprocess.check_call([B.NVCC] + B.ARCH + B.FLAGS + ["-DNF_MC_PROFILE", "-c", os.path.join(B.CSRC, "made_chain_bf16.cu"), "-o", obj])
objs = [o for o in glob.glob(os.path.join(B.OBJ, "*.o")) if not o.endswith("made_chain_bf16.o") and not o.endswith("_prof.o")] + [obj]
out = os.path.join(B.LIBDIR, "libnfb200_prof.so")
subprocess.check_call([B.NVCC] + B.ARCH + ["-shared", "-o", out + ".tmp"] + objs + ["-lcudart"])
os.replace(out + ".tmp", out)
print("built", out)
