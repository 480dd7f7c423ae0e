"""Deletes libraries in tla_rust_b200/csrc/native/ that no committed fixture (with the current engine sources) maps to."""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tla_rust_b200 import engine  # noqa: E402
from tla_rust_b200.compiled import load_compiled  # noqa: E402

keep = set()
for f in glob.glob(os.path.join(ROOT, "tests", "golden", "*.tlagz")):
    cm, _, _, _ = load_compiled(f)
    for sc in (False, True):
        keep.add(f"libtlag_{engine._sliced_tag(cm, sc)}.so")
n = 0
for f in os.listdir(engine.NATIVE_DIR):
    pth = os.path.join(engine.NATIVE_DIR, f)
    if f not in keep:
        (shutil.rmtree if os.path.isdir(pth) else os.remove)(pth)
        n += 1
print("removed", n, "kept", len(os.listdir(engine.NATIVE_DIR)))
