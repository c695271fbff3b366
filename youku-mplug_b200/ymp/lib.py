"""ctypes binding of libymp_b200.so (the C ABI declared in include/ymp.h).

There is deliberately NO fallback: if the shared library is missing the import raises, and every
op raises ``YmpError`` when a kernel call fails.  PyTorch is used only as the owner of device
memory and streams; tensors cross this boundary as raw pointers.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libymp_b200.so")


class YmpError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C youku-mplug_b200/csrc` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

ACT_NONE, ACT_GELU_ERF, ACT_GELU_TANH = 0, 1, 2
DT_BF16, DT_F32 = 0, 1

c_i32 = C.c_int32
c_vp = C.c_void_p


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", c_vp), ("B", c_vp), ("D", c_vp),
        ("M", c_i32), ("N", c_i32), ("K", c_i32),
        ("lda", c_i32), ("ldb", c_i32), ("ldd", c_i32),
        ("a_mn_major", c_i32), ("b_mn_major", c_i32),
        ("bias", c_vp), ("residual", c_vp), ("ldr", c_i32),
        ("act", c_i32), ("aux_out", c_vp), ("aux_in", c_vp),
        ("out_dtype", c_i32), ("accumulate", c_i32), ("split_k", c_i32),
        ("alpha", C.c_float), ("tile_n", c_i32),
    ]


lib.ymp_last_error.restype = C.c_char_p
lib.ymp_abi_version.restype = C.c_int
lib.ymp_launch_count.restype = C.c_uint64


def _declare(name, argstruct):
    fn = getattr(lib, name)
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(argstruct), c_vp]
    return fn


_gemm = _declare("ymp_gemm", GemmArgs)


def check(rc, what):
    if rc != 0:
        raise YmpError(f"{what} failed ({rc}): {lib.ymp_last_error().decode()}")


def launch_count():
    return int(lib.ymp_launch_count())


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def cur_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def call(fn, args, what):
    check(fn(C.byref(args), cur_stream()), what)
