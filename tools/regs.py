"""Print (registers, spill stores, spill loads) of the kernels matching the given substrings."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import test_build_artifacts as t  # noqa: E402

tab = t.ptxas_table()
for k, v in sorted(tab.items()):
    if not sys.argv[1:] or any(a in k for a in sys.argv[1:]):
        print(v, k.replace("ogl::(anonymous namespace)::", "")[:120])
