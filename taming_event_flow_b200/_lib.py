"""ctypes binding of ``libtef_b200.so`` (C ABI declared in ``include/tef_b200.h``).

There is no CPU fallback: every public function of this package ends in a call
into the CUDA library, and loading fails loudly when the library is missing.
"""
import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("TEF_B200_LIB", os.path.join(_HERE, "libtef_b200.so"))   # override: kernel-variant experiments
CSRC = os.path.join(_HERE, "csrc")

MAX_PASSES = 31
MAX_SCALES = 6
MAX_FLOWS = 8

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # bit-faithful per-event arithmetic, see csrc/tef_device.cuh
    "-Xcompiler", "-fPIC", "-shared",
]


class TefError(RuntimeError):
    pass


class TefShapeError(ValueError, RuntimeError):
    """A tensor handed to the host mirror has the wrong shape.  The reference fails with a torch RuntimeError (a broadcasting
    or indexing error somewhere inside the loss); this one is caught as either."""


class EmptyWindowError(RuntimeError, ValueError):
    """A temporal scale has no window to concatenate.  The reference fails in ``torch.cat([])`` (loss/flow.py:689), which
    raises RuntimeError up to torch 2.x and ValueError in recent releases (2.11 here): this one is caught by either."""


class CmDesc(ctypes.Structure):
    """Mirror of ``tef_cm_desc``."""
    _fields_ = (
        [(k, ctypes.c_int) for k in ("B", "H", "W", "P", "F", "S", "mode", "border_comp", "loss_scaling", "deterministic")]
        + [
            ("ev", (ctypes.c_void_p * MAX_PASSES) * 2),
            ("mk", (ctypes.c_void_p * MAX_PASSES) * 2),
            ("n", (ctypes.c_int * MAX_PASSES) * 2),
            ("flow", ctypes.c_void_p),
            ("gflow", ctypes.c_void_p),
            ("img", ctypes.c_void_p),
            ("acc_sum", ctypes.c_void_p),
            ("acc_nnz", ctypes.c_void_p),
            ("den", ctypes.c_void_p),
            ("loss", ctypes.c_void_p),
            ("grad_out", ctypes.c_void_p),
            ("sort_bins", ctypes.c_void_p),
            ("sort_sums", ctypes.c_void_p),
            ("sorted_ev", ctypes.c_void_p),
            ("posbuf", ctypes.c_void_p),
            ("alivebuf", ctypes.c_void_p),
            ("gimg", ctypes.c_void_p),
            ("hist_done", ctypes.c_int),
            ("reserved_", ctypes.c_int),
            ("flowq", ctypes.c_void_p),
            ("gimgq", ctypes.c_void_p),
        ]
    )


class UpdateDesc(ctypes.Structure):
    """Mirror of ``tef_update_desc``."""
    _fields_ = (
        [(k, ctypes.c_int) for k in ("F", "t", "P", "B", "H", "W")]
        + [
            ("flow_maps", ctypes.c_void_p * MAX_FLOWS),
            ("packed", ctypes.c_void_p),
            ("packedq", ctypes.c_void_p),
            ("events", ctypes.c_void_p * 2),
            ("masks", ctypes.c_void_p * 2),
            ("ev_out", ctypes.c_void_p * 2),
            ("mk_out", ctypes.c_void_p * 2),
            ("rows", ctypes.c_long * 2),
            ("pass_index", ctypes.c_float * 2),
            ("ts_override", ctypes.c_void_p * 2),
            ("sort_bins", ctypes.c_void_p),
            ("hist", ctypes.c_int),
            ("zero_bins", ctypes.c_int),
            ("strided", ctypes.c_int * 2),
            ("ev_strides", (ctypes.c_long * 3) * 2),
            ("mk_strides", (ctypes.c_long * 3) * 2),
        ]
    )


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(_ROOT, "include", "tef_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a into ``libtef_b200.so`` (in-tree, so it ships to the GPU box)."""
    if not (force or needs_build()):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("TEF_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(_ROOT, "include"), "-o", LIB_PATH] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise TefError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def lib():
    """The loaded CUDA library.  Raises if it has not been built -- there is no other code path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TefError(
                "libtef_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "this package has no CPU or PyTorch fallback." % LIB_PATH
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.tef_strerror.restype = ctypes.c_char_p
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().tef_strerror(int(rc)).decode()
        if rc == -3:
            raise EmptyWindowError("%s: %s" % (what, msg))  # reference: torch.cat([]) fails
        if rc == -4:
            raise TypeError("%s: %s" % (what, msg))        # reference: TypeError (None in torch.cat)
        raise TefError("%s failed: %s (code %d)" % (what, msg, rc))


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    import torch

    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def require_cuda(*tensors):
    import torch

    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise TefError("taming_event_flow_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU path" % t.device)
        if t.device.index != torch.cuda.current_device():
            # the C ABI launches on the current device's stream: a tensor of another GPU would be an illegal address
            raise TefError("tensor lives on %s but the current CUDA device is cuda:%d; use torch.cuda.set_device / torch.cuda.device"
                           % (t.device, torch.cuda.current_device()))
