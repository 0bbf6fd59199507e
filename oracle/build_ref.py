"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY.

Builds the UNMODIFIED reference `hv_cuda` extension (houghvoting/src/hv_cuda.cpp,
hv_cuda_kernel.cu; build recipe houghvoting/setup.py:7-10) from the sources where
they lie under /root/reference into oracle/_ref/hv_cuda_ref.so, for sm_100a.

Nothing is copied into the repository: the only adaptation is the force-included
header oracle/ref_compat_shim.h (an API-compat macro re-definition; arithmetic
untouched).  The reference has NO CPU path (hv_cuda.cpp:26-28 rejects CPU
tensors), so this module can only run on the GPU box, where it is used
 * by tests/ (-m gpu) as the real-reference checker for our kernels, and
 * by tools/make_golden.py to generate the golden fixtures under tests/golden/.

/root/reference does not exist on the GPU box: there only `load_ref()` is used,
which imports the prebuilt .so (oracle/_ref/ is git-ignored but travels with the
gpurun snapshot).
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/houghvoting/src"
NAME = "hv_cuda_ref"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    """Compile the reference extension. Returns the .so path, or None when the
    reference tree is not present (GPU box)."""
    if not os.path.isdir(REF_SRC):
        return so_path() if os.path.exists(so_path()) else None
    srcs = [os.path.join(REF_SRC, "hv_cuda.cpp"), os.path.join(REF_SRC, "hv_cuda_kernel.cu")]
    if os.path.exists(so_path()) and all(
            os.path.getmtime(so_path()) >= os.path.getmtime(s) for s in srcs + [os.path.join(HERE, "ref_compat_shim.h")]):
        return so_path()
    os.makedirs(OUT, exist_ok=True)
    from torch.utils.cpp_extension import load
    shim = os.path.join(HERE, "ref_compat_shim.h")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    load(name=NAME, sources=srcs,
         extra_cflags=["-include", shim, "-O2"],
         # same flags the reference's setup.py would give (no fast-math), sm_100a target
         extra_cuda_cflags=["-include", shim, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"],
         build_directory=OUT, is_python_module=False, verbose=verbose)
    return so_path()


def load_ref():
    """Import the prebuilt reference extension (needs a GPU to run anything)."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[NAME] = mod
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
