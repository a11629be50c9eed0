"""Builds an A/B variant of the library: the named sources recompiled with extra flags, every other object reused.

    python tools/build_variant.py NAME file.cu[,file2.cu] -DFLAG [-DFLAG2=3 ...]   ->  ppt_b200/libppt_b200_NAME.so

Select it with PPT_B200_LIB=ppt_b200/libppt_b200_NAME.so (ppt_b200/_lib.py).  Variants are measurement tools: they are
git-ignored like every built file and never loaded by default."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppt_b200 import build as B  # noqa: E402

name, files, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
B.build()
objdir = os.path.join(B.HERE, "_obj_" + name)
os.makedirs(objdir, exist_ok=True)
objs = []
for src in B._sources():
    if src in files:
        obj = os.path.join(objdir, src[:-3] + ".o")
        cmd = [B._nvcc()] + B.ARCH + B.COMMON + B.PER_FILE.get(src, []) + flags + ["-c", os.path.join(B.CSRC, src), "-o", obj]
        subprocess.check_call(cmd)
    else:
        obj = os.path.join(B.OBJ, src[:-3] + ".o")
    objs.append(obj)
lib = os.path.join(B.HERE, "libppt_b200_%s.so" % name)
subprocess.check_call([B._nvcc()] + B.ARCH + ["-shared", "-o", lib] + objs + ["-Xcompiler", "-fvisibility=hidden"])
print(lib)
