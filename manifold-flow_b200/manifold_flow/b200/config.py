"""Run-time switches of the B200 back end."""
import os


class _Config:
    # "fp32": conditioner GEMMs/convolutions reproduce the reference's fp32 results (parity mode)
    # "tf32": single-pass TF32 tensor-core math (throughput sweeps; not parity-clean)
    conditioner_math = os.environ.get("MFLOW_CONDITIONER_MATH", "fp32")
    # image-conditioner convolutions on the tcgen05 tensor cores (3xTF32); False = cuDNN fp32 library calls
    conv_tensor_cores = os.environ.get("MFLOW_CONV_TC", "1") != "0"
    # operand format of those kernels' error-compensated split: "fp16" (kind::f16, power-of-two operand scaling from the
    # tensors' absolute maxima; twice the tensor-core rate) or "tf32" (kind::tf32, no scaling)
    conv_operand_format = os.environ.get("MFLOW_CONV_FORMAT", "fp16")
    # weight gradients of those convolutions on the tensor cores too; False = cuDNN / cuBLAS fp32 library calls
    conv_wgrad_tensor_cores = os.environ.get("MFLOW_CONV_WGRAD_TC", "1") != "0"
    # smallest feature-map width that goes to the tensor-core weight-gradient kernel (4, 8, 16 or 32); below it the
    # reduction has too few pixel blocks to fill the machine and the library call is as fast
    conv_wgrad_min_width = int(os.environ.get("MFLOW_CONV_WGRAD_MIN_WIDTH", "4"))
    # vector conditioners (ResidualNet): from this many rows on, the whole net runs on the tensor-core convolution
    # kernels (a Linear layer is a 1x1 convolution once the batch rows are laid out as pixels); below it the
    # 2-CTA launches are latency bound and the cuBLAS sgemm calls are faster
    linear_tc_min_rows = int(os.environ.get("MFLOW_LINEAR_TC_MIN_ROWS", "32768"))
    # below that row count: the whole ResidualNet as one fused launch (mflow_mlp_forward / _backward); False = one cuBLAS
    # GEMM and a few ATen launches per layer
    fuse_vector_conditioner = os.environ.get("MFLOW_FUSE_MLP", "1") != "0"
    # evaluation (no-grad) passes of the image couplings: the spline runs in the epilogue of the conditioner's final 1x1
    # convolution and the parameter tensor is never written
    fuse_spline_epilogue = os.environ.get("MFLOW_FUSE_SPLINE", "1") != "0"
    # build the dense L / U / W / W^-1 matrices of all same-width LU layers of a model in one batched pass
    batch_lu_builds = os.environ.get("MFLOW_BATCH_LU", "1") != "0"
    # fuse ActNorm + 1x1 convolution (+ permutation + LU) into one channel-mix launch
    fuse_linear_glue = os.environ.get("MFLOW_FUSE_GLUE", "1") != "0"
    # full_jacobian=True: the reference's Permutation Jacobian is the transpose of the exact one
    # (permutations.py:49); False reproduces the reference (parity), True is mathematically exact
    exact_permutation_jacobian = os.environ.get("MFLOW_EXACT_PERM_JACOBIAN", "0") == "1"


    # Bumped by ``invalidate_caches()``: part of the key of every parameter-derived cache (dense L / U, composed channel
    # mixes, split convolution weights).  Writes through ``.data`` or by collectives do not bump ``Tensor._version``;
    # whoever writes parameters that way calls ``invalidate_caches()``.
    cache_epoch = 0


config = _Config()


def invalidate_caches():
    """Drop every parameter-derived cache (next use rebuilds).  Called after writes that bypass the version counter:
    ActNorm's data-dependent initialisation, ``ddp.broadcast_parameters``, and by users who assign ``param.data``."""
    config.cache_epoch += 1
