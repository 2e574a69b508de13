"""A/B build: lib/libnfb200_<tag>.so = the library with the given translation units recompiled with extra nvcc flags
(everything else from the objects of the normal build); select it at run time with NFB200_LIB=<path>.
    python normalizing-flows-study_b200/build.py && python scripts/build_variant_lib.py hint -DNF_MBAR_HINT=1 -- stack_tc gemm_tc2 made_chain_bf16"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "normalizing-flows-study_b200")
sys.path.insert(0, PKG)
import build as B  # noqa: E402

tag = sys.argv[1]
sep = sys.argv.index("--")
flags, units = sys.argv[2:sep], sys.argv[sep + 1:]
objs = [o for o in glob.glob(os.path.join(B.OBJ, "*.o")) if "_var_" not in o and not o.endswith("_prof.o")
        and os.path.basename(o)[:-2] not in units]
for u in units:
    obj = os.path.join(B.OBJ, f"{u}_var_{tag}.o")
    subprocess.check_call([B.NVCC] + B.ARCH + B.FLAGS + flags + ["-c", os.path.join(B.CSRC, u + ".cu"), "-o", obj])
    objs.append(obj)
out = os.path.join(B.LIBDIR, f"libnfb200_{tag}.so")
subprocess.check_call([B.NVCC] + B.ARCH + ["-shared", "-o", out + ".tmp"] + objs + ["-lcudart"])
os.replace(out + ".tmp", out)
print("built", out)
