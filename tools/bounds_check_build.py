"""Build the bounds-checked variant of the library (device-side index asserts, WB_DEV_ASSERT):

    python tools/bounds_check_build.py            # -> scratch/variants/libwayne_b200_checked.so
    WAYNE_B200_LIB=$PWD/scratch/variants/libwayne_b200_checked.so python -m pytest tests -m gpu -q

compute-sanitizer is not available on the GPU pool; this is the memory-safety evidence for the
shared-memory tables and tiles of the native kernels: a failed check traps and every later CUDA
call of the test run errors out."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wayne_b200 import build as b  # noqa: E402

out_dir = os.path.join(ROOT, 'scratch', 'variants')
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, 'libwayne_b200_checked.so')
cmd = ['nvcc'] + [f for f in b.NVCC_FLAGS if f not in ('-Xptxas', '-v')] + [
    '-DWB_BOUNDS_CHECK', '-o', out, os.path.join(b.CSRC, 'wayne_b200.cu')]
subprocess.run(cmd, check=True)
print(out)
