"""Test harness ONLY: runs a script written against yxgeee/OpenIBL *unmodified* with this repository's `ibl`
package on the import path.  The one concession is an empty `h5py` module: the reference scripts `import h5py`
at the top (examples/test.py:7) but never use it on this path, and the package is absent from this image.

    python -m torch.distributed.run --nproc-per-node 1 --master-addr 127.0.0.1 tests/fixtures/run_unmodified.py \
        <script.py> [script args...]
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
try:
    import h5py  # noqa: F401
except Exception:
    sys.modules["h5py"] = types.ModuleType("h5py")

if __name__ == "__main__":
    import runpy
    script = sys.argv[1]
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")
