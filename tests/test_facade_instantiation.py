"""Every class template of the header-only C++ facade (include/dealii_cuda_b200/*.h) explicitly instantiated by g++ -fsyntax-only:
an explicit instantiation compiles ALL member functions, also the ones no example calls (that is how LaplaceOperatorGpu::
get_diagonal_inverse went unnoticed with a DiagonalMatrix that could not be instantiated)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")

INSTANCES = {
    "GpuVector": ["double", "float"], "DiagonalMatrix": ["GpuVector<double>", "GpuVector<float>"], "GpuList": ["unsigned int"],
    "HyperCubeMesh": ["2", "3"], "AdaptiveMesh": ["2", "3"], "BallMesh": ["2", "3"], "ConstraintHandlerGpu": ["double", "float"],
    "MatrixFreeGpu": ["2, double", "3, float"], "LaplaceOperatorGpu": ["3, 4, double", "2, 2, float"],
    "MGTransferMatrixFreeGpu": ["3, double", "2, float"], "AdaptiveMultigrid": ["3, double", "2, float"],
    "InterfaceExchange": ["double", "float"], "LocalWorldLevel": ["3, 4, double", "2, 2, float"],
    "PartitionedChebyshev": ["LocalWorldLevel<3, 4, double>, double"], "PartitionedMultigrid": ["3, 4, double", "2, 2, float"],
}


def test_every_facade_template_is_listed():
    found = set()
    for name in os.listdir(os.path.join(INC, "dealii_cuda_b200")):
        if name.endswith(".h"):
            found |= set(re.findall(r"^template <[^>]*> class (\w+)", open(os.path.join(INC, "dealii_cuda_b200", name)).read(), flags=re.M))
    assert found == set(INSTANCES), found ^ set(INSTANCES)


def test_facade_templates_instantiate(tmp_path):
    src = ['#include "dealii_cuda_b200/partitioned_mg.h"', "using namespace dealii_cuda_b200;"]
    for cls, argss in INSTANCES.items():
        src += ["template class dealii_cuda_b200::%s<%s>;" % (cls, a) for a in argss]
    f = tmp_path / "inst.cc"
    f.write_text("\n".join(src) + "\nint main() { return 0; }\n")
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I", INC, str(f)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
