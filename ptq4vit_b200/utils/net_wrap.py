"""reference: utils/net_wrap.py:39-81 -- replace Linear / MatMul submodules by quant operators
built through cfg.get_module(module_type, *ctor_args)."""
import torch.nn as nn

from .models import MatMul

MODULE_TYPES = {"qkv": "qlinear_qkv", "proj": "qlinear_proj", "fc1": "qlinear_MLP_1", "fc2": "qlinear_MLP_2",
                "head": "qlinear_classifier", "matmul1": "qmatmul_qk", "matmul2": "qmatmul_scorev",
                "reduction": "qlinear_reduction"}


def wrap_modules_in_net(net, cfg, wrap_conv=True):
    wrapped_modules = {}
    module_dict = {}
    for name, m in list(net.named_modules()):
        module_dict[name] = m
        idx = name.rfind(".")
        father = module_dict[name[:idx] if idx != -1 else ""] if (name[:idx] if idx != -1 else "") in module_dict else None
        leaf = name[idx + 1:] if idx != -1 else name
        if father is None:
            continue
        if isinstance(m, nn.Conv2d):
            if not wrap_conv:
                continue      # the embedding conv stays FP32
            new_m = cfg.get_module("qconv", m.in_channels, m.out_channels, m.kernel_size, m.stride, m.padding, m.dilation,
                                   m.groups, m.bias is not None, m.padding_mode)
            new_m.weight.data = m.weight.data
            new_m.bias = m.bias
            new_m.to(m.weight.device)
        elif isinstance(m, nn.Linear):
            new_m = cfg.get_module(MODULE_TYPES[leaf], m.in_features, m.out_features)
            new_m.weight.data = m.weight.data
            new_m.bias = m.bias
            new_m.to(m.weight.device)
        elif isinstance(m, MatMul):
            new_m = cfg.get_module(MODULE_TYPES[leaf])
        else:
            continue
        wrapped_modules[name] = new_m
        setattr(father, leaf, new_m)
    return wrapped_modules
