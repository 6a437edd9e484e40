"""Tuning builds of the library beside the product one: python benchmarks/build_variants.py NAME=-DX=1,-DY=2 ...
Each goes to shacira_b200/build_NAME/lib.so (git- and gpurun-ignored dirs are NOT used: the .so must travel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shacira_b200 import build  # noqa: E402

for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    out = os.path.join(ROOT, "shacira_b200", "variant_%s.so" % name)
    build.build(force=True, extra=[f for f in flags.split(",") if f], out=out, tag="_" + name)
    print(out)
