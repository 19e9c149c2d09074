"""Serving-framework glue for the Petit FP4 linear layers.

The reference ships only the ops; the helpers that call them live in its users (vLLM's and
SGLang's ``petit_utils``: ``verify_petit_nvfp4_supported``, ``prepare_nvfp4_layer_for_petit``,
``apply_petit_nvfp4_linear``).  They are restated here with the same names, arguments and
behaviour so that a framework can import them from the kernel package instead, and so that the
whole weight-load -> forward path is covered by this repo's parity tests.  What they do is
fixed by the op contracts of ``petit_kernel/__init__.py:17-60``: repack the *local TP shard*
once after loading, keep the packed tensors as frozen ``nn.Parameter``s, call ``mul_*_a16``
with the activations flattened to 2-D, add the bias in place, restore the leading shape.

Added on top (no counterpart in the frameworks):

* a **layout version tag**: packed tensors are opaque and differ between builds (this build:
  ``ops.packed_layout_version()``); ``export_packed_state`` / ``load_packed_state`` carry the
  version, the format and the shapes, and refuse a mismatch instead of multiplying garbage;
* MXFP4 twins of the NVFP4 helpers.

Everything here runs the CUDA ops; there is no CPU or eager fallback.
"""
from __future__ import annotations

import torch

from . import ops

# vLLM / SGLang: the quantisation methods and the group size the Petit path accepts
_SUPPORTED = {"NVFP4": 16, "MXFP4": 32}


def _check_petit_supported(fmt: str, quant_method: str, group_size):
    if quant_method != fmt:
        return False, (f"Petit currently only supports: {fmt} quantizations in this path. "
                       f"Got quant_method={quant_method!r}.")
    if group_size is not None and group_size != _SUPPORTED[fmt]:
        return False, (f"Petit currently only supports: group_size={_SUPPORTED[fmt]} for {fmt}. "
                       f"Got group_size={group_size}.")
    return True, None


def _check_petit_nvfp4_supported(quant_method: str, group_size):
    return _check_petit_supported("NVFP4", quant_method, group_size)


def verify_petit_nvfp4_supported(quant_method: str, group_size) -> None:
    supported, error_msg = _check_petit_nvfp4_supported(quant_method, group_size)
    if not supported:
        raise ValueError(error_msg)


def verify_petit_mxfp4_supported(quant_method: str, group_size) -> None:
    supported, error_msg = _check_petit_supported("MXFP4", quant_method, group_size)
    if not supported:
        raise ValueError(error_msg)


def _tag(layer: torch.nn.Module, fmt: str, size_n: int, size_k: int) -> None:
    layer.petit_format = fmt
    layer.petit_layout_version = ops.packed_layout_version()
    layer.petit_size_n, layer.petit_size_k = int(size_n), int(size_k)


def repack_nvfp4_for(qw: torch.Tensor, size_n: int, size_k: int, a_dtype=None) -> torch.Tensor:
    """``petit_kernel.repack_nvfp4`` with the activation type the weights will meet.
    ``a_dtype=torch.float16``: the fp16-native packed layout, whose decode kernel converts a
    pair of weights with one ``cvt`` + one multiply (the packed tensor stays opaque and keeps
    its dtype / byte count; ``mul_nvfp4_a16`` recognises the layout from the tensor's shape and
    rejects bfloat16 activations for it).  ``None`` / bfloat16: the default layout, which every
    activation type can use."""
    return ops.repack_nvfp4(qw, size_n, size_k, a_dtype)


def interleave_gate_up(t: torch.Tensor) -> torch.Tensor:
    """Row order the fused SiLU * mul epilogue needs (petit.h, PETIT_ACT_SILU_MUL): ``t`` is a
    merged gate_up tensor ``[2 I, ...]`` = gate rows then up rows (weights, block scales or a
    bias); the result holds, per 128 rows, 64 gate rows followed by the 64 matching up rows."""
    two_i = t.shape[0]
    assert two_i % 128 == 0, "gate_up rows must be a multiple of 128"
    i = two_i // 2
    gate = t[:i].reshape(i // 64, 64, *t.shape[1:])
    up = t[i:].reshape(i // 64, 64, *t.shape[1:])
    return torch.cat((gate, up), dim=1).reshape(t.shape).contiguous()


def prepare_nvfp4_layer_for_petit(layer: torch.nn.Module, fuse_silu_mul: bool = False) -> None:
    """``process_weights_after_loading`` of an NVFP4 linear layer: ``layer.weight`` is the
    checkpoint's packed e2m1 bytes ``[N, K/2]`` (any 1-byte dtype) of the local shard,
    ``layer.weight_scale`` its ``float8_e4m3fn [N, K/16]`` block scales.
    ``fuse_silu_mul=True`` (a merged gate_up projection): interleaves gate and up rows so that
    ``apply_petit_nvfp4_linear(..., silu_mul=True)`` returns ``silu(gate) * up`` directly."""
    part_size_n = layer.output_size_per_partition
    part_size_k = layer.input_size_per_partition
    if fuse_silu_mul:
        layer.weight = torch.nn.Parameter(interleave_gate_up(layer.weight.data), requires_grad=False)
        layer.weight_scale = torch.nn.Parameter(
            interleave_gate_up(layer.weight_scale.data.view(torch.uint8)).view(layer.weight_scale.dtype),
            requires_grad=False)
        if getattr(layer, "bias", None) is not None:
            layer.bias = torch.nn.Parameter(interleave_gate_up(layer.bias.data), requires_grad=False)
        layer.petit_silu_mul = True
    qweight = layer.weight.view(torch.int32).contiguous()
    # a float16 model gets the fp16-native layout (the layer's activations are params_dtype)
    a_dtype = torch.float16 if getattr(layer, "params_dtype", None) == torch.float16 else None
    petit_qweight = ops.repack_nvfp4(qweight, part_size_n, part_size_k, a_dtype)
    layer.weight = torch.nn.Parameter(petit_qweight, requires_grad=False)
    weight_scale = ops.process_nvfp4_scales(layer.weight_scale.data.contiguous(), part_size_n,
                                            part_size_k)
    layer.weight_scale = torch.nn.Parameter(weight_scale, requires_grad=False)
    _tag(layer, "NVFP4", part_size_n, part_size_k)


def prepare_mxfp4_layer_for_petit(layer: torch.nn.Module) -> None:
    """Same for MXFP4: ``layer.weight_scale`` is ``uint8`` e8m0 ``[N, K/32]``."""
    part_size_n = layer.output_size_per_partition
    part_size_k = layer.input_size_per_partition
    qweight = layer.weight.view(torch.int32).contiguous()
    layer.weight = torch.nn.Parameter(ops.repack_nvfp4(qweight, part_size_n, part_size_k),
                                      requires_grad=False)
    layer.weight_scale = torch.nn.Parameter(
        ops.process_mxfp4_scales(layer.weight_scale.data.contiguous(), part_size_n, part_size_k),
        requires_grad=False)
    _tag(layer, "MXFP4", part_size_n, part_size_k)


def _apply(mul, input, weight, weight_scale, weight_scale_2, size_n, size_k, bias,
           residual=None, silu_mul=False):
    reshaped_x = input.reshape(-1, input.shape[-1])
    out_shape = input.shape[:-1] + (size_n // 2 if silu_mul else size_n,)
    # solution_id=-1: the library's chooser, which honours the tuned-solution table
    # (petit_kernel.tuning) -- the frameworks' "TODO: use auto-tuning" lives there
    if silu_mul:
        if bias is not None:
            bias = bias.to(reshaped_x.dtype).contiguous()
        output = ops.mul_fp4_a16_ex_out(None, reshaped_x, weight, weight_scale, weight_scale_2,
                                        reshaped_x.size(0), size_n, size_k, -1,
                                        mul is ops.mul_mxfp4_a16, bias, None, True)
    elif bias is None and residual is None:
        output = mul(reshaped_x, weight, weight_scale, weight_scale_2, reshaped_x.size(0), size_n,
                     size_k, -1)
    else:
        # fused epilogue: what the frameworks do as `output.add_(bias)` (and the residual add)
        # right after the op happens on the fp32 accumulator, before the one rounding
        if bias is not None:
            bias = bias.to(reshaped_x.dtype).contiguous()
        if residual is not None:
            residual = residual.reshape(-1, size_n).contiguous()
        output = ops.mul_fp4_a16_ex_out(None, reshaped_x, weight, weight_scale, weight_scale_2,
                                        reshaped_x.size(0), size_n, size_k, -1,
                                        mul is ops.mul_mxfp4_a16, bias, residual)
    return output.reshape(out_shape)


def apply_petit_nvfp4_linear(input: torch.Tensor, weight: torch.Tensor, weight_scale: torch.Tensor,
                             weight_scale_2: torch.Tensor, size_n: int, size_k: int,
                             bias: torch.Tensor | None = None,
                             residual: torch.Tensor | None = None,
                             silu_mul: bool = False) -> torch.Tensor:
    """Forward of an NVFP4 linear layer; ``weight_scale_2`` is the float32 device tensor with
    the global scale (read inside the kernel: no host sync, CUDA-graph safe).  ``bias`` and the
    optional ``residual`` (same shape as the output) are added inside the GEMM epilogue;
    ``silu_mul=True`` (layer prepared with ``fuse_silu_mul=True``) returns
    ``silu(gate) * up`` of shape ``[..., size_n / 2]``."""
    return _apply(ops.mul_nvfp4_a16, input, weight, weight_scale, weight_scale_2, size_n, size_k,
                  bias, residual, silu_mul)


def apply_petit_mxfp4_linear(input: torch.Tensor, weight: torch.Tensor, weight_scale: torch.Tensor,
                             weight_scale_2: torch.Tensor, size_n: int, size_k: int,
                             bias: torch.Tensor | None = None,
                             residual: torch.Tensor | None = None,
                             silu_mul: bool = False) -> torch.Tensor:
    return _apply(ops.mul_mxfp4_a16, input, weight, weight_scale, weight_scale_2, size_n, size_k,
                  bias, residual, silu_mul)


# ---- versioned packed state -----------------------------------------------------
def export_packed_state(layer: torch.nn.Module) -> dict:
    """The repacked tensors of a prepared layer plus what is needed to trust them later."""
    if not hasattr(layer, "petit_layout_version"):
        raise ValueError("layer has not been prepared with prepare_*_layer_for_petit")
    return {
        "petit_format": layer.petit_format,
        "petit_layout_version": layer.petit_layout_version,
        "size_n": layer.petit_size_n,
        "size_k": layer.petit_size_k,
        "weight": layer.weight.data,
        "weight_scale": layer.weight_scale.data,
    }


def check_packed_state(state: dict, fmt: str | None = None) -> None:
    """Raise if `state` was not produced by THIS build's layout for `fmt`."""
    have = ops.packed_layout_version()
    if state.get("petit_layout_version") != have:
        raise ValueError(
            f"packed tensors have layout version {state.get('petit_layout_version')!r}, this "
            f"build reads version {have}: re-run repack_*/process_* on the checkpoint tensors")
    if fmt is not None and state.get("petit_format") != fmt:
        raise ValueError(f"packed tensors are {state.get('petit_format')!r}, expected {fmt!r}")
    n, k = state["size_n"], state["size_k"]
    group = _SUPPORTED[state["petit_format"]]
    if state["weight"].numel() * state["weight"].element_size() != n * k // 2:
        raise ValueError("packed weight size does not match size_n x size_k")
    if state["weight_scale"].numel() * state["weight_scale"].element_size() != n * k // group:
        raise ValueError("packed scale size does not match size_n x size_k")


def load_packed_state(layer: torch.nn.Module, state: dict) -> None:
    """Install previously exported packed tensors into `layer` (skips the repack)."""
    check_packed_state(state)
    layer.weight = torch.nn.Parameter(state["weight"], requires_grad=False)
    layer.weight_scale = torch.nn.Parameter(state["weight_scale"], requires_grad=False)
    _tag(layer, state["petit_format"], state["size_n"], state["size_k"])
