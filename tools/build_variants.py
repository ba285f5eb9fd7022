"""Build A/B variants of the C-ABI library (extra -D flags) into echoseal_b200/_variants/lib_<name>.so.
Usage: python tools/build_variants.py name1:-DFOO=1,-DBAR=2 name2:-DFOO=0 ...
Run a variant with ES_B200_LIB=echoseal_b200/_variants/lib_<name>.so (developer tool; the product loads the
in-tree libechoseal_b200.so)."""
import glob, os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G

def build_one(spec):
    name, _, flags = spec.partition(":")
    out = os.path.join(G.PKG, "_variants", f"lib_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(G.PKG, "csrc", "*.cu"))) + sorted(glob.glob(os.path.join(G.PKG, "csrc", "*.cpp")))
    cmd = ["/usr/local/cuda/bin/nvcc"] + G.NVCC_FLAGS + [f for f in flags.split(",") if f] + \
          ["-I", os.path.join(ROOT, "include"), "-o", out] + srcs + G.LIBS
    subprocess.check_call(cmd)
    return out

if __name__ == "__main__":
    with ThreadPoolExecutor(4) as ex:
        for o in ex.map(build_one, sys.argv[1:]):
            print("built", o)
