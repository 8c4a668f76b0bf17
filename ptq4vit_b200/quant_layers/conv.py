"""Conv2d quant operators with the reference's class surface (quant_layers/conv.py).

PTQ4ViT wraps exactly one convolution per network, the patch embedding, with
`ChannelwiseBatchingQuantConv2d(..., a_bit=32)` (configs/PTQ4ViT.py:52-54): per-output-channel weight
step sizes, activations left in FP32.  The search runs in the CUDA library as a batched product over
the images: rows = output channels (candidate planes of the quantised kernel), columns = output
positions (im2col of the FP32 input, split exactly into three bf16 terms), one score per channel.
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from ._metric import metric_weight


class MinMaxQuantConv2d(nn.Conv2d):
    """reference: quant_layers/conv.py:9-89"""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, padding=0, dilation=1, groups: int = 1,
                 bias: bool = True, padding_mode: str = "zeros", mode="raw", w_bit=8, a_bit=8, bias_bit=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
        self.n_calibration_steps = 2
        self.mode = mode
        self.w_bit = w_bit
        self.a_bit = a_bit
        self.bias_bit = bias_bit
        assert bias_bit is None, "No support bias bit now"
        self.w_interval = None
        self.a_interval = None
        self.bias_interval = None
        self.raw_input = None
        self.raw_out = None
        self.metric = None
        self.next_nodes = []
        self.w_qmax = 2 ** (self.w_bit - 1)
        self.a_qmax = 2 ** (self.a_bit - 1)

    def forward(self, x):
        if self.mode == "raw":
            out = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        elif self.mode == "quant_forward":
            out = self.quant_forward(x)
        elif self.mode == "calibration_step1":
            out = self.calibration_step1(x)
        elif self.mode == "calibration_step2":
            out = self.calibration_step2(x)
        else:
            raise NotImplementedError
        return out

    def quant_weight_bias(self):
        """reference: conv.py:53-62"""
        wi = torch.as_tensor(self.w_interval, dtype=torch.float32, device=self.weight.device)
        w_sim = (self.weight / wi).round_().clamp_(-self.w_qmax, self.w_qmax - 1).mul_(wi)
        return w_sim, self.bias

    def quant_input(self, x):
        """reference: conv.py:64-67"""
        ai = torch.as_tensor(self.a_interval, dtype=torch.float32, device=x.device)
        return (x / ai).round_().clamp_(-self.a_qmax, self.a_qmax - 1).mul_(ai)

    def quant_forward(self, x):
        """reference: conv.py:69-74"""
        assert self.calibrated is not None, f"You should run calibrate_forward before run quant_forward for {self}"
        w_sim, bias_sim = self.quant_weight_bias()
        x_sim = self.quant_input(x) if self.a_bit < 32 else x
        return F.conv2d(x_sim, w_sim, bias_sim, self.stride, self.padding, self.dilation, self.groups)

    def calibration_step1(self, x):
        """reference: conv.py:76-81"""
        out = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        self.raw_input = x.cpu().detach()
        self.raw_out = out.cpu().detach()
        return out

    def calibration_step2(self, x):
        """reference: conv.py:83-89 (layer-wise min-max)"""
        self.w_interval = (self.weight.data.abs().max() / (self.w_qmax - 0.5)).detach()
        self.a_interval = (x.abs().max() / (self.a_qmax - 0.5)).detach()
        self.calibrated = True
        return self.quant_forward(x)


class PTQSLQuantConv2d(MinMaxQuantConv2d):
    """reference: quant_layers/conv.py:126-277 -- constructor surface (the sub-layerwise search of the non-batching
    class is not part of PTQ4ViT's configuration; only the channel-wise batching class below searches natively)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, padding=0, dilation=1, groups: int = 1,
                 bias: bool = True, padding_mode: str = "zeros", mode="raw", w_bit=8, a_bit=8, bias_bit=None,
                 metric="L2_norm", search_round=1, eq_alpha=0.1, eq_beta=2, eq_n=100, parallel_eq_n=10, n_V=1, n_H=1,
                 init_layerwise=False):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                         groups=groups, bias=bias, padding_mode=padding_mode, mode=mode, w_bit=w_bit, a_bit=a_bit,
                         bias_bit=bias_bit)
        self.metric = metric
        self.search_round = search_round
        self.eq_alpha = eq_alpha
        self.eq_beta = eq_beta
        self.eq_n = int(eq_n)
        self.n_H = n_H
        self.n_V = n_V
        self.parallel_eq_n = parallel_eq_n
        self.crb_rows = out_channels // n_V
        self.crb_cols = in_channels // n_H
        self.init_layerwise = init_layerwise
        self.raw_grad = None
        self.keep_scores = False
        self.last_scores = None


class ChannelwiseBatchingQuantConv2d(PTQSLQuantConv2d):
    """reference: quant_layers/conv.py:444-613.  `a_bit >= 32` turns the activation quantizer off (the only way
    PTQ4ViT uses this class); the weight step size is searched per output channel."""
    batching = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.n_V = self.out_channels
        self.n_H = 1
        self.calib_size = None
        self.calib_batch_size = None
        self.calib_need_batching = False

    def _initialize_calib_parameters(self):
        """reference: conv.py:467-480; a whole layer fits in HBM, no batching."""
        self.calib_size = int(self.raw_input.shape[0])
        self.calib_batch_size = int(self.raw_input.shape[0])

    def _grad_for_metric(self, y):
        """Per-element weight of the metric (conv.py:509-522); see _metric.py."""
        return metric_weight(self.metric, y, self.raw_grad, "_get_similarity")

    def calibration_step2(self):
        """reference: conv.py:591-603.  The weight search does not depend on anything the rounds change when the
        activations are not quantized, so its result is the same in every round: it is run once."""
        if self.a_bit < 32:
            raise NotImplementedError("ChannelwiseBatchingQuantConv2d: the B200 path implements a_bit >= 32 "
                                      "(activation quantizer off), as configs/PTQ4ViT.py:54 uses it")
        if self.groups != 1 or self.init_layerwise:
            raise NotImplementedError("ChannelwiseBatchingQuantConv2d: groups == 1 and init_layerwise=False only")
        self._initialize_calib_parameters()
        dev = self.weight.device
        if dev.type != "cuda":
            raise RuntimeError("ptq4vit_b200 quant layers need their parameters on a CUDA device (no CPU path)")
        x = self.raw_input.to(dev).float()
        y = self.raw_out.to(dev).float().contiguous()
        g = self._grad_for_metric(y).to(dev).float().contiguous()
        n, oc = y.shape[0], y.shape[1]
        L = y.shape[2] * y.shape[3]
        # im2col of the FP32 input: [n, L, K] (row l = one output position), K = ic*kh*kw in the kernel's own order
        cols = F.unfold(x, self.kernel_size, self.dilation, self.padding, self.stride).transpose(1, 2).contiguous()
        K = cols.shape[2]
        w2 = self.weight.detach().reshape(oc, K).contiguous().float()
        b = None if self.bias is None else self.bias.detach().contiguous().float()
        d = _lib.ConvDesc()
        d.images, d.out_channels, d.K, d.positions = n, oc, K, L
        d.w_bit, d.eq_n = int(self.w_bit), int(self.eq_n)
        d.eq_alpha, d.eq_beta = float(self.eq_alpha), float(self.eq_beta)
        d.has_bias = 0 if b is None else 1
        d.kernel = _lib.default_kernel()
        lib = _lib.lib()
        nbytes = ctypes.c_size_t()
        _lib.check(lib.p4v_conv_workspace_bytes(ctypes.byref(d), ctypes.byref(nbytes)), "p4v_conv_workspace_bytes")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        w_int = torch.empty(oc, dtype=torch.float32, device=dev)
        log = torch.empty(self.eq_n * oc, dtype=torch.float32, device=dev) if self.keep_scores else None
        _lib.check(lib.p4v_conv_calibrate(ctypes.byref(d), _lib.ptr(cols), _lib.ptr(w2), _lib.ptr(b), _lib.ptr(y.view(n, oc, L)),
                                          _lib.ptr(g.view(n, oc, L)), _lib.ptr(ws), nbytes.value, _lib.ptr(w_int), _lib.ptr(log),
                                          ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "p4v_conv_calibrate")
        self.w_interval = w_int.view(oc, 1, 1, 1)
        self.a_interval = (x.abs().max() / (self.a_qmax - 0.5)).detach().view(1)   # conv.py:490-496 (unused when a_bit >= 32)
        self.last_scores = [log.view(self.eq_n, oc)] * int(self.search_round) if log is not None else None
        self.calibrated = True
        del self.raw_input, self.raw_out, self.raw_grad
        return None
