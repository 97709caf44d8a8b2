"""Builds variant libraries (one per define set) next to the product library: python tools/build_variants.py name=DEF1,DEF2 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from densecap_b200 import build as b
for spec in sys.argv[1:]:
    name, _, defs = spec.partition("=")
    out = os.path.join(b.PKG_DIR, "libdensecap_b200_%s.so" % name)
    print(b.build_library(force=False, defines=tuple(d for d in defs.split(",") if d), out_path=out), flush=True)
