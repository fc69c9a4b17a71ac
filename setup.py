"""Packaging of wayne-b200 (the reference ships setup.py:52-75 with one Cython
extension and the `wayne` console script; here the native artefact is
libwayne_b200.so, compiled in-tree by nvcc for sm_100a).

    python -m pip install --no-build-isolation -e .     # builds the library if nvcc is on PATH
    wayne -p params.yml
"""
import shutil

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py


class BuildWithCuda(build_py):
    def run(self):
        if shutil.which("nvcc") or shutil.which("/usr/local/cuda/bin/nvcc"):
            from wayne_b200.build import build
            build()
        build_py.run(self)


setup(
    name="wayne-b200",
    version="0.1.0",
    description="B200-native per-exposure detector-image synthesis with the API of ucl-exoplanets/wayne",
    packages=find_packages(include=["wayne", "wayne_b200", "wayne_b200.*"]),
    package_data={"wayne_b200": ["libwayne_b200.so", "data/*", "csrc/*"]},
    include_package_data=True,
    python_requires=">=3.9",
    install_requires=["numpy", "pyyaml", "torch"],
    entry_points={"console_scripts": ["wayne = wayne_b200.run_visit:run"]},
    cmdclass={"build_py": BuildWithCuda},
)
