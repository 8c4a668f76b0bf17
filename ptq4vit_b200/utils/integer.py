"""Integer export of a calibrated model with the reference's surface (utils/integer.py:8-129): int8 weights, int8 /
uint8 (twin-uniform) activations.  The quantise-and-pack pass runs in the CUDA library (csrc/export.cu); results are
device tensors (the reference moves them to the host, a `.cpu()` away).

Differences from the reference, both documented there as limits of its own code: weights use the module's real block
structure (the reference's `weight / w_interval` only broadcasts for n_V = n_H = 1, integer.py:15), activations honour
`n_a` chunks (the reference broadcasts `[n_a, 1]` against the feature axis, valid for n_a = 1, integer.py:55, :66).
"""
import ctypes

import torch

from .. import _lib
from ..quant_layers.linear import GELU_MIN_NEG, MinMaxQuantLinear
from ..quant_layers.matmul import MinMaxQuantMatMul

MODE_INT8, MODE_GELU_TWIN, MODE_SOS_TWIN = 0, 1, 2


def _export(src, delta, rows_per_block, n_row_blocks, cols_per_block, n_col_blocks, mode, bit, d_neg=0.0, split=None):
    if src.device.type != "cuda":
        raise RuntimeError("ptq4vit_b200.utils.integer needs CUDA tensors (no CPU path)")
    x = src.detach().contiguous().float()
    cols = x.shape[-1]
    rows = x.numel() // cols
    d = torch.as_tensor(delta, dtype=torch.float32, device=x.device).reshape(-1).contiguous()
    assert d.numel() == n_row_blocks * n_col_blocks, f"step-size table has {d.numel()} entries, expected {n_row_blocks}x{n_col_blocks}"
    sp = None if split is None else torch.as_tensor(split, dtype=torch.float32, device=x.device).reshape(1).contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8 if mode != MODE_INT8 else torch.int8, device=x.device)
    _lib.check(_lib.lib().p4v_export_quantized(_lib.ptr(x), rows, cols, _lib.ptr(d), int(rows_per_block), int(n_row_blocks),
                                                int(cols_per_block), int(n_col_blocks), int(mode), int(bit), float(d_neg),
                                                _lib.ptr(sp), _lib.ptr(out),
                                                ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
               "p4v_export_quantized")
    return out


def quantize_int_weight(module):
    """reference: integer.py:8-18 -- int8 weight of a calibrated Linear (bias stays FP32)."""
    assert hasattr(module, "weight"), f"module {module} does not have weight"
    assert module.w_bit == 8, f"module {module}'s weight is quantized with {module.w_bit} bits"
    n_V, n_H = getattr(module, "n_V", 1), getattr(module, "n_H", 1)
    O, K = module.weight.shape[0], module.weight[0].numel()
    w2 = module.weight.detach().reshape(O, K)
    return _export(w2, module.w_interval, O // n_V, n_V, K // n_H, n_H, MODE_INT8, module.w_bit).view(module.weight.shape)


def dequantize_int_weight(module, w_int):
    """reference: integer.py:20-25"""
    n_V, n_H = getattr(module, "n_V", 1), getattr(module, "n_H", 1)
    wi = torch.as_tensor(module.w_interval, dtype=torch.float32, device=w_int.device).reshape(n_V, 1, n_H, 1)
    O, K = w_int.shape[0], w_int[0].numel()
    return (w_int.float().reshape(n_V, O // n_V, n_H, K // n_H) * wi).reshape(w_int.shape)


def quantize_matmul_input(input, interval, qmax, n_G, n_V, n_H, crb_groups, crb_rows, crb_cols):
    """reference: integer.py:27-42 -- int8 operand of a MatMul; head-wise layout (crb_groups = n_V = n_H = 1), the one
    the Batching classes force (matmul.py:411-417)."""
    if (crb_groups, n_V, n_H) != (1, 1, 1):
        raise NotImplementedError("quantize_matmul_input: head-wise layout only (crb_groups = n_V = n_H = 1)")
    bit = int(qmax).bit_length()
    return _export(input, interval, input.shape[2], n_G, input.shape[3], 1, MODE_INT8, bit).float()


def quantize_int_activation(module, input):
    """reference: integer.py:44-109 -- forward pre-hook that stores the integer inputs of a calibrated module in
    `module.int_input`.  Twin-uniform layouts: post-GELU and post-softmax activations are uint8 with the reference's
    `+128` region bit."""
    from ..quant_layers.linear import PostGeluPTQSLQuantLinear, PostGeluPTQSLBatchingQuantLinear
    if isinstance(module, MinMaxQuantLinear):
        assert module.a_bit == 8, f"module {module}'s activation is quantized with {module.a_bit} bits"
        x = input[0]
        n_a = getattr(module, "n_a", 1)
        K = x.shape[-1]
        if isinstance(module, (PostGeluPTQSLQuantLinear, PostGeluPTQSLBatchingQuantLinear)):
            a_int = module.a_interval[0] if isinstance(module.a_interval, (list, tuple)) else module.a_interval
            module.int_input = [_export(x, a_int, 0, 1, K // n_a, n_a, MODE_GELU_TWIN, module.a_bit, d_neg=GELU_MIN_NEG / module.a_qmax)]
        else:
            module.int_input = [_export(x, module.a_interval, 0, 1, K // n_a, n_a, MODE_INT8, module.a_bit)]
    elif isinstance(module, MinMaxQuantMatMul):
        assert module.A_bit == 8, f"module {module}'s matrix A is quantized with {module.A_bit} bits"
        assert module.B_bit == 8, f"module {module}'s matrix B is quantized with {module.B_bit} bits"
        A, B = input[0], input[1]
        H = A.shape[1]
        if getattr(module, "sos", False):
            A_int = _export(A, module.A_interval, 0, 1, A.shape[-1], 1, MODE_SOS_TWIN, module.A_bit, split=module.split)
        else:
            A_int = _export(A, module.A_interval, A.shape[2], H, A.shape[3], 1, MODE_INT8, module.A_bit)
        B_int = _export(B, module.B_interval, B.shape[2], H, B.shape[3], 1, MODE_INT8, module.B_bit)
        module.int_input = [A_int, B_int]


def get_model_int_weight(wrapped_modules):
    """reference: integer.py:113-129 -- {name: int8 weight} of every wrapped module that has one."""
    int_weights = {}
    for name, m in wrapped_modules.items():
        try:
            int_weights[name] = quantize_int_weight(m)
        except (AssertionError, AttributeError):
            pass
    return int_weights
