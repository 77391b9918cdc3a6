"""Packaging (role of the reference's setup.py).  The native extensions are built in-tree by
`skycomputing_b200._build` (nvcc for sm_100a + g++/pybind11); `pip install -e .` only installs the
Python package - run `python __graft_entry__.py` (or `python -m skycomputing_b200._build`) first."""
from setuptools import find_packages, setup


def _version():
    ns = {}
    exec(open("skycomputing_b200/version.py").read(), ns)
    return ns["__version__"]


setup(
    name="skycomputing_b200",
    version=_version(),
    description="B200-native load-balanced pipeline-model-parallel training (SkyComputing capabilities)",
    packages=find_packages(exclude=("tests", "tools", "baseline", "experiment")),
    package_data={"skycomputing_b200": ["*.so"]},
    python_requires=">=3.10",
    install_requires=["torch>=2.4", "numpy", "psutil", "pybind11"],
)
