"""ctypes binding of ``libmflow_b200.so`` (C ABI declared in ``include/mflow_b200.h``).

The library is loaded lazily: building a model (``nn.Module`` construction, ``state_dict``
handling) works on a CPU-only box, but the first kernel call without the library or without a
CUDA tensor raises - there is no CPU fallback and no eager-PyTorch fallback on this path.
"""
import contextlib
import ctypes
import os
import threading

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "lib", "libmflow_b200.so")
_lib = None
_lock = threading.Lock()
launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)

c_f32p = ctypes.c_void_p
i64 = ctypes.c_int64
i32 = ctypes.c_int32
f32 = ctypes.c_float


class RqsDesc(ctypes.Structure):
    """mflow_rqs_desc"""

    _fields_ = [
        ("n_outer", i64), ("n_chan", i64), ("n_inner", i64),
        ("x_so", i64), ("x_sc", i64), ("y_so", i64), ("y_sc", i64),
        ("p_so", i64), ("p_sc", i64), ("p_sp", i64), ("p_si", i64),
        ("num_bins", i32), ("inverse", i32),
        ("tail_bound", f32), ("min_bin_width", f32), ("min_bin_height", f32), ("min_derivative", f32),
        ("wh_divisor", f32),
        ("n_id_chan", i64), ("id_off", i64), ("id_sc", i64),
        ("p_hw", i64), ("t_off", i64),
    ]


class RqsBoxDesc(ctypes.Structure):
    """mflow_rqs_box_desc"""

    _fields_ = [
        ("n", i64), ("num_bins", i32), ("inverse", i32),
        ("left", f32), ("right", f32), ("bottom", f32), ("top", f32),
        ("min_bin_width", f32), ("min_bin_height", f32), ("min_derivative", f32),
    ]


class MlpDesc(ctypes.Structure):
    """mflow_mlp_desc"""

    _fields_ = [("rows", i64), ("n_in", i32), ("n_ctx", i32), ("hidden", i32), ("n_out", i32), ("n_blocks", i32)]


class ConvDesc(ctypes.Structure):
    """mflow_conv_desc"""

    _fields_ = [
        ("batch", i64), ("c_in", i64), ("c_out", i64), ("height", i64), ("width", i64),
        ("kernel", i32), ("relu_in", i32), ("relu_out", i32),
        ("n_pad", i64), ("k_pad", i64),
    ]


class MixDesc(ctypes.Structure):
    """mflow_mix_desc"""

    _fields_ = [
        ("n_outer", i64), ("n_chan", i64), ("n_inner", i64),
        ("x_so", i64), ("x_sc", i64), ("x_si", i64),
        ("y_so", i64), ("y_sc", i64), ("y_si", i64),
    ]


_SIGNATURES = {
    "mflow_version": (ctypes.c_int, []),
    "mflow_last_error_string": (ctypes.c_char_p, []),
    "mflow_check_device": (ctypes.c_int, []),
    "mflow_rqs_workspace_bytes": (i64, [ctypes.POINTER(RqsDesc)]),
    "mflow_rqs_apply": (ctypes.c_int, [ctypes.POINTER(RqsDesc)] + [c_f32p] * 9),
    "mflow_rqs_backward": (ctypes.c_int, [ctypes.POINTER(RqsDesc)] + [c_f32p] * 10),
    "mflow_rqs_box_apply": (ctypes.c_int, [ctypes.POINTER(RqsBoxDesc)] + [c_f32p] * 6),
    "mflow_rqs_box_backward": (ctypes.c_int, [ctypes.POINTER(RqsBoxDesc)] + [c_f32p] * 8),
    "mflow_rqs_search": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, i64, i32, c_f32p]),
    "mflow_mlp_workspace_floats": (i64, [ctypes.POINTER(MlpDesc)]),
    "mflow_mlp_forward": (ctypes.c_int, [ctypes.POINTER(MlpDesc), c_f32p, i64, i64, c_f32p, i64, c_f32p, c_f32p, c_f32p, c_f32p]),
    "mflow_mlp_backward": (ctypes.c_int, [ctypes.POINTER(MlpDesc), c_f32p, i64, i64, c_f32p, i64, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
                                          c_f32p, c_f32p, c_f32p]),
    "mflow_conv2d_tf32x3": (ctypes.c_int, [ctypes.POINTER(ConvDesc)] + [c_f32p] * 8),
    "mflow_conv2d_f16x3": (ctypes.c_int, [ctypes.POINTER(ConvDesc)] + [c_f32p] * 11),
    "mflow_conv2d_wgrad_f16x3": (ctypes.c_int, [ctypes.POINTER(ConvDesc)] + [c_f32p] * 8),
    "mflow_absmax": (ctypes.c_int, [c_f32p, i64, c_f32p, c_f32p]),
    "mflow_conv2d_wgrad_workspace_bytes": (i64, [ctypes.POINTER(ConvDesc)]),
    "mflow_conv2d_rqs_workspace_bytes": (i64, [ctypes.POINTER(ConvDesc)]),
    "mflow_conv2d_rqs_f16x3": (ctypes.c_int, [ctypes.POINTER(ConvDesc)] + [c_f32p] * 6 + [ctypes.POINTER(RqsDesc)] + [c_f32p] * 5),
    "mflow_conv2d_wgrad_tf32x3": (ctypes.c_int, [ctypes.POINTER(ConvDesc)] + [c_f32p] * 6),
    "mflow_chanmix_apply": (ctypes.c_int, [ctypes.POINTER(MixDesc)] + [c_f32p] * 5),
    "mflow_chanmix_wgrad_workspace_bytes": (i64, [ctypes.POINTER(MixDesc)]),
    "mflow_chanmix_wgrad": (ctypes.c_int, [ctypes.POINTER(MixDesc)] + [c_f32p] * 6),
    "mflow_channel_sum_workspace_bytes": (i64, [i64, i64, i64]),
    "mflow_channel_sum": (ctypes.c_int, [c_f32p, c_f32p, i64, i64, i64, c_f32p, c_f32p]),
    "mflow_squeeze2x2": (ctypes.c_int, [c_f32p, c_f32p, i64, i64, i64, i64, i32, f32, f32, c_f32p]),
    "mflow_affine_scalar": (ctypes.c_int, [c_f32p, c_f32p, i64, f32, f32, i32, c_f32p]),
    "mflow_permute_rows": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, i64, i64, c_f32p]),
    "mflow_logtanh_apply": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, i64, i64, i32, f32, c_f32p]),
    "mflow_logtanh_backward": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, i64, i64, i32, f32, c_f32p]),
    "mflow_add_amax": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, i64, c_f32p]),
    "mflow_glu_residual": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, i64, c_f32p]),
    "mflow_glu_residual_backward": (ctypes.c_int, [c_f32p] * 7 + [i64, c_f32p]),
    "mflow_gauss_logprob": (ctypes.c_int, [c_f32p, i64, i64, c_f32p, i64, i64, f32, f32, c_f32p, c_f32p, c_f32p, i64, c_f32p]),
    "mflow_gauss_logprob_backward": (ctypes.c_int, [c_f32p, i64, i64, c_f32p, i64, i64, f32, f32, c_f32p, c_f32p, c_f32p, i64, c_f32p]),
    "mflow_jtj_logdet": (ctypes.c_int, [c_f32p, c_f32p, i64, i32, i32, c_f32p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class MflowError(RuntimeError):
    pass


def library_path():
    return _LIB_PATH


def _load(path):
    if not os.path.exists(path):
        raise MflowError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C manifold-flow_b200/csrc). The B200 path has no CPU or eager fallback."
        )
    handle = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, args
    return handle


def lib():
    """The loaded library; raises if it was never built (no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                _lib = _load(_LIB_PATH)
    return _lib


_variants = {}


@contextlib.contextmanager
def library_variant(name):
    """Run the enclosed calls through ``libmflow_b200_<name>.so`` (same ABI).  ``exact``: the spline kernels compiled with
    -DMFLOW_RQS_EXACT_ORDER, i.e. every fp32 operation of the reference separately rounded in the reference's order
    (csrc/rqs_math.cuh) - the parity tests use it to separate algorithmic agreement from the default build's
    FMA contraction and ex2.approx."""
    global _lib
    if name not in _variants:
        _variants[name] = _load(_LIB_PATH.replace(".so", f"_{name}.so"))
    prev = lib()
    _lib = _variants[name]
    try:
        yield _lib
    finally:
        _lib = prev


def check(rc, what=""):
    if rc != 0:
        msg = lib().mflow_last_error_string().decode()
        if rc == 1:
            raise ValueError(f"{what}: {msg}" if what else msg)
        raise MflowError(f"{what}: {msg} (code {rc})")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MflowError(
                "manifold_flow (B200 build) runs on CUDA tensors only: got a tensor on "
                f"{t.device}. Move the model and its inputs to the GPU (there is no CPU fallback)."
            )
        if t is not None and t.dtype != torch.float32:
            raise MflowError(f"the B200 kernels compute in float32; got {t.dtype}")
        if t is not None and t.device.index != torch.cuda.current_device():
            # kernels launch on the current device's current stream: a tensor of another device would be an illegal
            # address (or a silent peer access)
            raise MflowError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: wrap the call "
                             "in `with torch.cuda.device(tensor.device):` (the flows' public methods do this for host inputs)")


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_workspaces = {}


def workspace(tag, nbytes, device):
    """Zero-initialised scratch, cached per (purpose, device, stream).  The kernels keep their
    arrival counters in a fixed region at the start and leave them zeroed."""
    key = (tag, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def count_launch(n=1):
    global launch_count
    launch_count += n


# ---- optional per-kernel device timing (bench.py: roofline.achieved) -----------------------------
class KernelTimer:
    """Brackets selected launches with CUDA events on the launching stream and accumulates, per tag,
    the device time and the algorithmic bytes.  Off unless bench.py installs one."""

    def __init__(self):
        self.records = []  # (tag, bytes, start_event, end_event)

    def begin(self, tag, nbytes):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return (tag, nbytes, ev)

    def end(self, token):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.records.append((token[0], token[1], token[2], ev))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for tag, nbytes, e0, e1 in self.records:
            d = out.setdefault(tag, dict(launches=0, ms=0.0, bytes=0))
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["bytes"] += nbytes
        return out


kernel_timer = None
