"""Packaging for the drop-in console script.  `pip install -e .` puts a `polycracker` command on PATH whose
hot-path sub-commands run on the GPU; the CUDA library is built in-tree by build.py."""
from setuptools import setup

setup(
    name="polycracker_b200",
    version="0.1.0",
    packages=["polycracker_b200"],
    package_dir={"polycracker_b200": "polycracker-unofficial-mirror_b200"},
    package_data={"polycracker_b200": ["libpolycracker_b200.so", "csrc/*"]},
    install_requires=["numpy", "scipy", "click"],
    entry_points={"console_scripts": ["polycracker=polycracker_b200.cli:main", "polycracker-b200=polycracker_b200.cli:main"]},
)
