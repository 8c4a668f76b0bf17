"""Linear quant operators with the reference's class surface (quant_layers/linear.py).

Same constructors, attributes (`w_interval [n_V,1,n_H,1]`, `a_interval [n_a,1]`,
`calibrated`, `mode`, `raw_input/raw_out/raw_grad`) and methods; the interval search
and the quantized forward run in the CUDA library (ptq4vit_b200._lib).  There is no
PyTorch fallback for the search: a missing library or a CPU tensor raises.
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from ._metric import metric_weight

GELU_MIN_NEG = 0.16997124254703522  # reference: quant_layers/linear.py:574


def _flat2d(t):
    return t.reshape(-1, t.shape[-1]).contiguous().float()


class _QuantLinearFn(torch.autograd.Function):
    """Keeps the native quantized forward inside autograd.  In the reference `quant_forward` is
    F.linear(x_sim, w_sim, bias) with x_sim / w_sim built by `.round_()` (linear.py:46-67), whose derivative is zero:
    no gradient reaches x or the weight, but the output carries a grad_fn (through the bias / weight Parameters), so
    that with sequential=True the gradient hooks of the modules BEHIND an already-quantized layer still fire
    (utils/quant_calib.py:330-341).  Same here: the output requires grad, every input gradient is zero."""

    @staticmethod
    def forward(ctx, module, x, weight, bias):
        ctx.x_meta = (x.shape, x.dtype, x.device) if x.requires_grad else None
        return module._quant_forward_native(x)

    @staticmethod
    def backward(ctx, grad_out):
        gx = None if ctx.x_meta is None else torch.zeros(ctx.x_meta[0], dtype=ctx.x_meta[1], device=ctx.x_meta[2])
        return None, gx, None, None


class MinMaxQuantLinear(nn.Linear):
    """reference: quant_layers/linear.py:6-92"""

    post_gelu = False

    def __init__(self, in_features: int, out_features: int, bias: bool = True, mode="raw", w_bit=8, a_bit=8,
                 bias_bit=None, bias_correction=False):
        super().__init__(in_features, out_features, bias)
        self.n_calibration_step = 2
        self.mode = mode
        self.w_bit = w_bit
        self.a_bit = a_bit
        self.bias_bit = bias_bit
        assert bias_bit is None, "No support bias bit now"
        self.w_interval = None
        self.a_interval = None
        self.raw_input = None
        self.raw_out = None
        self.metric = None
        self.next_nodes = []
        self.w_qmax = 2 ** (self.w_bit - 1)
        self.a_qmax = 2 ** (self.a_bit - 1)
        self.bias_correction = bias_correction
        # block structure of the base class: one block
        self.n_V = self.n_H = self.n_a = 1

    def forward(self, x):
        if self.mode == "raw":
            out = F.linear(x, self.weight, self.bias)
        elif self.mode == "quant_forward":
            out = self.quant_forward(x)
        elif self.mode == "calibration_step1":
            out = self.calibration_step1(x)
        elif self.mode == "calibration_step2":
            out = self.calibration_step2(x)
        else:
            raise NotImplementedError
        return out

    # ---- native plumbing -------------------------------------------------
    def _desc(self, rows, tokens, search_round=1, eq=(0.0, 1.0, 1)):
        d = _lib.LinearDesc()
        d.rows, d.tokens = int(rows), int(tokens)
        d.in_features, d.out_features = self.in_features, self.out_features
        d.n_V, d.n_H, d.n_a = int(self.n_V), int(self.n_H), int(self.n_a)
        d.w_bit, d.a_bit = int(self.w_bit), int(self.a_bit)
        d.eq_n, d.search_round = int(eq[2]), int(search_round)
        d.eq_alpha, d.eq_beta = float(eq[0]), float(eq[1])
        d.post_gelu = 1 if self.post_gelu else 0
        d.has_bias = 0 if self.bias is None else 1
        d.operand = _lib.default_operand()
        d.kernel = _lib.default_kernel()
        d.init_layerwise = 1 if getattr(self, "init_layerwise", False) else 0
        return d

    def _device(self):
        dev = self.weight.device
        if dev.type != "cuda":
            raise RuntimeError("ptq4vit_b200 quant layers need their parameters on a CUDA device "
                               "(no CPU path; the reference semantics live in oracle/ for tests only)")
        return dev

    def _w_flat(self):
        return torch.as_tensor(self.w_interval, dtype=torch.float32, device=self._device()).reshape(-1).contiguous()

    def _a_flat(self):
        a = self.a_interval
        if isinstance(a, (list, tuple)):        # non-batching PostGelu keeps [pos, neg]
            a = a[0]
        return torch.as_tensor(a, dtype=torch.float32, device=self._device()).reshape(-1).contiguous()

    def quant_forward(self, x):
        """reference: linear.py:62-67 -- fq(x) @ fq(W)^T + b on the tensor cores."""
        assert self.calibrated is not None, f"You should run calibrate_forward before run quant_forward for {self}"
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            return _QuantLinearFn.apply(self, x, self.weight, self.bias)
        return self._quant_forward_native(x)

    def _quant_forward_native(self, x):
        dev = self._device()
        x2 = _flat2d(x.to(dev))
        d = self._desc(x2.shape[0], 1)
        lib = _lib.lib()
        nbytes = ctypes.c_size_t()
        _lib.check(lib.p4v_linear_quant_forward_workspace_bytes(ctypes.byref(d), ctypes.byref(nbytes)), "quant_forward_workspace")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        out = torch.empty(x2.shape[0], self.out_features, dtype=torch.float32, device=dev)
        w = self.weight.detach().contiguous().float()
        b = None if self.bias is None else self.bias.detach().contiguous().float()
        wi, ai = self._w_flat(), self._a_flat()
        _lib.check(lib.p4v_linear_quant_forward(ctypes.byref(d), _lib.ptr(x2), _lib.ptr(w), _lib.ptr(b), _lib.ptr(wi),
                                                _lib.ptr(ai), _lib.ptr(ws), nbytes.value, _lib.ptr(out),
                                                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "p4v_linear_quant_forward")
        return out.reshape(*x.shape[:-1], self.out_features)

    def quant_weight_bias(self):
        """reference: linear.py:46-55 / :152-162 (fake-quantized weight as a float tensor)."""
        wi = torch.as_tensor(self.w_interval, dtype=torch.float32, device=self.weight.device).reshape(self.n_V, 1, self.n_H, 1)
        w = self.weight.view(self.n_V, self.out_features // self.n_V, self.n_H, self.in_features // self.n_H)
        w_sim = (w / wi).round_().clamp_(-self.w_qmax, self.w_qmax - 1).mul_(wi).view(self.out_features, self.in_features)
        return w_sim, self.bias

    def quant_input(self, x):
        """reference: linear.py:57-60 / :164-169 / :601-607."""
        ai = self._a_flat().to(x.device).reshape(self.n_a, 1)
        xv = x.reshape(*x.shape[:-1], self.n_a, self.in_features // self.n_a)
        if self.post_gelu:
            neg = GELU_MIN_NEG / self.a_qmax
            x_pos = (xv / ai).round_().clamp_(0, self.a_qmax - 1).mul_(ai)
            x_neg = (xv / neg).round_().clamp_(-self.a_qmax, 0).mul_(neg)
            return (x_pos + x_neg).reshape_as(x)
        return (xv / ai).round_().clamp_(-self.a_qmax, self.a_qmax - 1).mul_(ai).reshape_as(x)

    def _bias_correction_quant_forward(self, x):
        """reference: linear.py:69-77"""
        if self.bias_correction and self.bias is not None:
            w_sim = self.quant_weight_bias()[0]
            x_sim = self.quant_input(x)
            eps = F.linear(x_sim, w_sim - self.weight.data, None)
            eps = torch.mean(eps, dim=(list(range(len(eps.shape) - 1))), keepdim=False)
            self.bias -= eps
            self.bias_correction = False
        return self.quant_forward(x)

    def calibration_step1(self, x):
        """reference: linear.py:79-84"""
        out = F.linear(x, self.weight, self.bias)
        self.raw_input = x.cpu().detach()
        self.raw_out = out.cpu().detach()
        return out

    def calibration_step2(self, x):
        """reference: linear.py:86-92 (layer-wise min-max)"""
        self.w_interval = (self.weight.data.abs().max() / (self.w_qmax - 0.5)).detach()
        self.a_interval = (x.abs().max() / (self.a_qmax - 0.5)).detach()
        self.calibrated = True
        out = self._bias_correction_quant_forward(x)
        return out


class PTQSLQuantLinear(MinMaxQuantLinear):
    """reference: quant_layers/linear.py:94-260"""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, mode="raw", w_bit=8, a_bit=8,
                 bias_bit=None, bias_correction=False, metric="L2_norm", search_round=1, eq_alpha=0, eq_beta=1,
                 eq_n=100, parallel_eq_n=10, n_H=1, n_V=1, n_a=1, init_layerwise=False):
        super().__init__(in_features, out_features, bias=bias, mode=mode, w_bit=w_bit, a_bit=a_bit, bias_bit=bias_bit,
                         bias_correction=bias_correction)
        self.metric = metric
        self.search_round = search_round
        self.eq_alpha = eq_alpha
        self.eq_beta = eq_beta
        self.eq_n = eq_n
        self.n_H = n_H
        self.n_V = n_V
        self.n_a = n_a
        self.crb_rows = out_features // n_V
        self.crb_cols = in_features // n_H  # ignore remnent != 0 situations
        self.crb_acts = in_features // n_a
        self.parallel_eq_n = parallel_eq_n   # kept for signature parity; the B200 path holds a whole layer in HBM
        self.init_layerwise = init_layerwise
        self.raw_grad = None
        self.last_scores = None              # optional per-step score tables (set P4V_SCORE_LOG=1 or keep_scores=True)
        self.keep_scores = False

    # ---- native search ---------------------------------------------------
    def _grad_for_metric(self, y):
        """Per-element weight of the metric (linear.py:406-422); see _metric.py."""
        return metric_weight(self.metric, y, self.raw_grad, "_get_similarity")

    def _native_calibrate(self, x, y, g):
        dev = self._device()
        tokens = 1
        if x.dim() > 2:
            for s in x.shape[1:-1]:
                tokens *= int(s)
        x2, y2, g2 = _flat2d(x.to(dev)), _flat2d(y.to(dev)), _flat2d(g.to(dev))
        d = self._desc(x2.shape[0], tokens, self.search_round, (self.eq_alpha, self.eq_beta, self.eq_n))
        lib = _lib.lib()
        nbytes, nlog = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(lib.p4v_linear_workspace_bytes(ctypes.byref(d), ctypes.byref(nbytes)), "p4v_linear_workspace_bytes")
        _lib.check(lib.p4v_linear_score_log_floats(ctypes.byref(d), ctypes.byref(nlog)), "p4v_linear_score_log_floats")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        w = self.weight.detach().contiguous().float()
        b = None if self.bias is None else self.bias.detach().contiguous().float()
        w_int = torch.empty(self.n_V * self.n_H, dtype=torch.float32, device=dev)
        a_int = torch.empty(self.n_a, dtype=torch.float32, device=dev)
        log = torch.empty(nlog.value, dtype=torch.float32, device=dev) if self.keep_scores else None
        _lib.check(lib.p4v_linear_calibrate(ctypes.byref(d), _lib.ptr(x2), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y2),
                                            _lib.ptr(g2), _lib.ptr(ws), nbytes.value, _lib.ptr(w_int), _lib.ptr(a_int),
                                            _lib.ptr(log), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "p4v_linear_calibrate")
        self.w_interval = w_int.view(self.n_V, 1, self.n_H, 1)
        self.a_interval = a_int.view(self.n_a, 1)
        self.last_scores = self._split_log(log) if log is not None else None

    def _split_log(self, log):
        out, o = [], 0
        for _ in range(self.search_round):
            for _h in range(self.n_H):
                out.append(log[o:o + self.eq_n * self.n_V].view(self.eq_n, self.n_V)); o += self.eq_n * self.n_V
            for _a in range(self.n_a):
                out.append(log[o:o + self.eq_n]); o += self.eq_n
        return out

    def calibration_step2(self, x):
        """reference: linear.py:235-260 (x already on the device; raw_out/raw_grad attributes)"""
        y = self.raw_out
        self._native_calibrate(x, y, self._grad_for_metric(y))
        if self.post_gelu:   # the non-batching PostGelu class stores [pos, neg] (linear.py:316-320)
            self.a_interval = [self.a_interval, GELU_MIN_NEG / self.a_qmax]
        self.calibrated = True
        with torch.no_grad():
            out = self._bias_correction_quant_forward(x)
        del self.raw_input, self.raw_out, self.raw_grad
        return out


class PostGeluPTQSLQuantLinear(PTQSLQuantLinear):
    """reference: quant_layers/linear.py:262-347 (twin-uniform post-GELU activations)"""
    post_gelu = True


class PTQSLBatchingQuantLinear(PTQSLQuantLinear):
    """reference: quant_layers/linear.py:349-555"""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.calib_size = None
        self.calib_batch_size = None
        self.calib_need_batching = False

    def _initialize_calib_parameters(self):
        """reference: linear.py:365-378.  180 GB of HBM hold a whole layer: no batching."""
        self.calib_size = int(self.raw_input.shape[0])
        self.calib_batch_size = int(self.raw_input.shape[0])
        self.calib_need_batching = False

    def calibration_step2(self):
        """reference: linear.py:536-555 -- only uses the cached raw inputs / outs / grads."""
        self._initialize_calib_parameters()
        y = self.raw_out
        self._native_calibrate(self.raw_input, y, self._grad_for_metric(y))
        self.calibrated = True
        del self.raw_input, self.raw_out, self.raw_grad
        return None


class PostGeluPTQSLBatchingQuantLinear(PTQSLBatchingQuantLinear):
    """reference: quant_layers/linear.py:557-642"""
    post_gelu = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.a_neg_interval = GELU_MIN_NEG / self.a_qmax
