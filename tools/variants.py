"""Development tool: build libsfm_b200 variants with -D overrides (here, no GPU needed) and time
them on the GPU box in one gpurun call.
  python tools/variants.py build name1:-DSFM_QDEPTH=4,-DSFM_STAGES=4 name2:...
  python tools/variants.py run [n_img] [n_desc]     (on the GPU box; prints one JSON line each)
Variant libraries live in build/variants/ (git-ignored, but they travel with gpurun)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "build", "variants")
sys.path.insert(0, ROOT)


def build(specs):
    from sfm_opencv_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for f in os.listdir(VDIR):
        os.remove(os.path.join(VDIR, f))
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        flags = [f for f in flags.split(",") if f]
        objs = []
        for src in b.SOURCES:
            o = os.path.join(VDIR, f"{name}_{src[:-3]}.o")
            if src == "match_knn.cu" or not os.path.exists(os.path.join(b.CSRC, src[:-3] + ".o")):
                subprocess.run([b._nvcc()] + b.NVCC_FLAGS + flags + ["-c", os.path.join(b.CSRC, src), "-o", o],
                               check=True, stderr=subprocess.DEVNULL)
            else:
                o = os.path.join(b.CSRC, src[:-3] + ".o")
            objs.append(o)
        lib = os.path.join(VDIR, f"lib_{name}.so")
        subprocess.run([b._nvcc()] + b.LINK_FLAGS + ["-o", lib] + objs + ["-ldl", "-lpthread", "-lrt"],
                       check=True, stderr=subprocess.DEVNULL)
        print("built", lib)
    for f in os.listdir(VDIR):
        if f.endswith(".o"):
            os.remove(os.path.join(VDIR, f))


def run(n_img, n_desc, modes):
    libs = sorted(f for f in os.listdir(VDIR) if f.endswith(".so"))
    for lib in libs:
        for mode in modes:
            env = dict(os.environ, SFM_B200_LIB=os.path.join(VDIR, lib))
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "exp_one.py"), str(mode), str(n_img),
                                  str(n_desc)], capture_output=True, text=True, env=env, timeout=600)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
            print(lib, line, flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 24
        n_desc = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
        modes = [int(m) for m in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1]
        run(n_img, n_desc, modes)
