"""Build script mirroring the reference's (`python setup.py build_cmake --inplace`,
reference setup.py:43-122): a `build_cmake` command that configures and builds CMakeLists.txt.
`python setup.py install|develop` installs the pure-Python package."""

import os
import subprocess
import sys

from setuptools import Command, find_packages, setup


class BuildCMake(Command):
    description = "configure and build the CUDA library and the pybind11 module with CMake"
    user_options = [("inplace", "i", "kept for compatibility with the reference's flag (always in place)")]

    def initialize_options(self):
        self.inplace = True

    def finalize_options(self):
        pass

    def run(self):
        root = os.path.dirname(os.path.abspath(__file__))
        build = os.path.join(root, "build", "cmake")
        os.makedirs(build, exist_ok=True)
        env = dict(os.environ)
        env.setdefault("CUDACXX", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call(["cmake", root, "-DCMAKE_BUILD_TYPE=Release", "-DPython_EXECUTABLE=" + sys.executable],
                              cwd=build, env=env)
        subprocess.check_call(["cmake", "--build", ".", "--", "-j4"], cwd=build, env=env)


setup(
    name="flamingpy_b200",
    version="0.1.0",
    description="B200-native backend for FlamingPy's Monte-Carlo error-correction trial path",
    packages=find_packages(include=["flamingpy_b200", "flamingpy_b200.*"]),
    package_data={"flamingpy_b200": ["libfpb.so", "cpp/*.so"]},
    python_requires=">=3.9",
    install_requires=["numpy", "networkx"],
    cmdclass={"build_cmake": BuildCMake},
)
