"""Operator factory with the reference's contract (configs/PTQ4ViT.py:1-80): module-level
kwargs dicts that experiment code mutates in place, and get_module(module_type, *args)."""
from ..quant_layers.conv import ChannelwiseBatchingQuantConv2d
from ..quant_layers.linear import PTQSLBatchingQuantLinear, PostGeluPTQSLBatchingQuantLinear
from ..quant_layers.matmul import PTQSLBatchingQuantMatMul, SoSPTQSLBatchingQuantMatMul

no_softmax = False
no_postgelu = False

bit = 8
conv_fc_name_list = ["qconv", "qlinear_qkv", "qlinear_proj", "qlinear_MLP_1", "qlinear_MLP_2", "qlinear_classifier", "qlinear_reduction"]
matmul_name_list = ["qmatmul_qk", "qmatmul_scorev"]
w_bit = {name: bit for name in conv_fc_name_list}
a_bit = {name: bit for name in conv_fc_name_list}
A_bit = {name: bit for name in matmul_name_list}
B_bit = {name: bit for name in matmul_name_list}

ptqsl_conv2d_kwargs = {"metric": "hessian", "eq_alpha": 0.01, "eq_beta": 1.2, "eq_n": 100, "search_round": 3, "n_V": 1, "n_H": 1}
ptqsl_linear_kwargs = {"metric": "hessian", "eq_alpha": 0.01, "eq_beta": 1.2, "eq_n": 100, "search_round": 3,
                       "n_V": 1, "n_H": 1, "n_a": 1, "bias_correction": True}
ptqsl_matmul_kwargs = {"metric": "hessian", "eq_alpha": 0.01, "eq_beta": 1.2, "eq_n": 100, "search_round": 3,
                       "n_G_A": 1, "n_V_A": 1, "n_H_A": 1, "n_G_B": 1, "n_V_B": 1, "n_H_B": 1}


def _linear_class(module_type):
    if module_type == "qlinear_MLP_2" and not no_postgelu:
        return PostGeluPTQSLBatchingQuantLinear          # twin-uniform post-GELU input (configs/PTQ4ViT.py:61-65)
    return PTQSLBatchingQuantLinear


def _matmul_class(module_type):
    if module_type == "qmatmul_scorev" and not no_softmax:
        return SoSPTQSLBatchingQuantMatMul               # split-of-softmax A operand (configs/PTQ4ViT.py:75-79)
    return PTQSLBatchingQuantMatMul


# per-type edits of the shared Linear kwargs: q, k and v each get their own row blocks; the classifier head keeps one
_N_V_RULE = {"qlinear_qkv": lambda n_V: 3 * n_V, "qlinear_classifier": lambda n_V: 1}


def get_module(module_type, *args, **kwargs):
    """reference: configs/PTQ4ViT.py:51-80.  The kwargs dicts and the two switches are read at call time, so an
    experiment's cfg_modifier can edit them between calls (example/test_all.py:59-75)."""
    if module_type == "qconv":
        opts = {**kwargs, **ptqsl_conv2d_kwargs}
        return ChannelwiseBatchingQuantConv2d(*args, **opts, w_bit=w_bit["qconv"], a_bit=32)   # activation quantizer off
    if "qlinear" in module_type:
        opts = {**kwargs, **ptqsl_linear_kwargs}
        if module_type in _N_V_RULE:
            opts["n_V"] = _N_V_RULE[module_type](opts["n_V"])
        return _linear_class(module_type)(*args, **opts, w_bit=w_bit[module_type], a_bit=a_bit[module_type])
    if "qmatmul" in module_type:
        opts = {**kwargs, **ptqsl_matmul_kwargs}
        return _matmul_class(module_type)(*args, **opts, A_bit=A_bit[module_type], B_bit=B_bit[module_type])
    raise NotImplementedError(f"unknown module type {module_type}")
