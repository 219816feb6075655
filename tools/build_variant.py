"""Build the library from a git ref (or the working tree: ref '.') into tools/bin/libpbe_<name>.so for A/B runs on one box
(PBE_B200_LIB=tools/bin/libpbe_<name>.so python ...).  Usage: python tools/build_variant.py <name> <git-ref|.>"""
import os, shutil, subprocess, sys, tempfile
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbe_b200.build import NVCC_FLAGS, _nvcc

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, ref = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp(prefix="pbe_variant_")
if ref == ".":
    shutil.copytree(os.path.join(root, "pbe_b200", "csrc"), os.path.join(tmp, "pbe_b200", "csrc"))
    shutil.copytree(os.path.join(root, "include"), os.path.join(tmp, "include"))
else:
    tar = subprocess.run(["git", "-C", root, "archive", ref, "pbe_b200/csrc", "include"], capture_output=True, check=True).stdout
    subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
srcs = sorted(f for f in os.listdir(os.path.join(tmp, "pbe_b200", "csrc")) if f.endswith(".cu"))
nvcc = _nvcc()
def cc(f):
    o = os.path.join(tmp, f[:-3] + ".o")
    r = subprocess.run([nvcc, *NVCC_FLAGS, "-c", os.path.join(tmp, "pbe_b200", "csrc", f), "-o", o], capture_output=True, text=True)
    if r.returncode: raise RuntimeError(r.stderr)
    return o
with ThreadPoolExecutor(8) as ex: objs = list(ex.map(cc, srcs))
os.makedirs(os.path.join(root, "tools", "bin"), exist_ok=True)
out = os.path.join(root, "tools", "bin", f"libpbe_{name}.so")
subprocess.run([nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
shutil.rmtree(tmp)
print(out)
