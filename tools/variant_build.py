"""Build a variant of the library with extra -D flags (A/B experiments on the GPU box):

    python tools/variant_build.py mb4 -DWB_THROW_MIN_BLOCKS=4     # -> scratch/variants/libwayne_b200_mb4.so
    WAYNE_B200_LIB=$PWD/scratch/variants/libwayne_b200_mb4.so python bench.py ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wayne_b200 import build as b  # noqa: E402

tag, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, 'scratch', 'variants')
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, 'libwayne_b200_%s.so' % tag)
cmd = ['nvcc'] + b.NVCC_FLAGS + flags + ['-o', out, os.path.join(b.CSRC, 'wayne_b200.cu')]
res = subprocess.run(cmd, capture_output=True, text=True)
log = res.stdout + res.stderr
if res.returncode:
    sys.stderr.write(log)
    sys.exit(1)
import re
for m in re.finditer(r"Compiling entry function '(\S*k_throw_philoxILi128ELi64ELb1\S*)'.*?Used (\d+) registers", log, re.S):
    print('k_throw_philox<direct>:', m.group(2), 'registers')
print(out)
