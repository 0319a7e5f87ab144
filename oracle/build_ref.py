"""Compile the UNMODIFIED reference querier extension into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.

Sources are read where they lie under /root/reference (never copied into the repo):
  pointnerf/models/neural_points/cuda/query_worldcoords.cpp, query_worldcoords.cu
Output: oracle/_ref/query_worldcoords_cuda.so (git-ignored; travels to the GPU box with the
snapshot).  It can only EXECUTE on a GPU; tests/test_gpu_reference_querier.py uses it there as
the cross-check for rows G1/G2/Q, and bench.py never touches it.

    python oracle/build_ref.py            # ~2-3 min (nvcc sm_100a + torch headers)
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/pointnerf/models/neural_points/cuda"
NAME = "query_worldcoords_cuda"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=[os.path.join(SRC, "query_worldcoords.cpp"), os.path.join(SRC, "query_worldcoords.cu")],
         build_directory=OUT, verbose=verbose, is_python_module=False)
    return so_path() if os.path.exists(so_path()) else None


def load_module():
    """Import the built extension (needs `import torch` first for libtorch symbols)."""
    import torch  # noqa: F401
    p = so_path()
    if not os.path.exists(p):
        return None
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
