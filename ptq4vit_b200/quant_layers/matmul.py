"""MatMul quant operators with the reference's class surface (quant_layers/matmul.py).

Result attributes as in the reference: `A_interval`, `B_interval` of shape
[1, n_G, 1, n_V, 1, n_H, 1] (head-wise: n_G = heads, n_V = n_H = 1); the split-of-softmax
variant keeps a 0-d `split` and a 0-d `A_interval = split / (A_qmax - 1)`.
"""
import ctypes

import torch
from torch import nn

from .. import _lib
from ._metric import metric_weight


class _QuantMatMulFn(torch.autograd.Function):
    """The native quantized product inside autograd: the reference's quant_forward (matmul.py:40-45) rounds both
    operands in place, so no gradient flows to A or B, but the output still has a grad_fn whenever an input requires
    grad -- which the gradient hooks of later modules rely on with sequential=True (utils/quant_calib.py:330-341)."""

    @staticmethod
    def forward(ctx, module, A, B):
        ctx.meta = [(t.shape, t.dtype, t.device) if t.requires_grad else None for t in (A, B)]
        return module._quant_forward_native(A, B)

    @staticmethod
    def backward(ctx, grad_out):
        g = [None if m is None else torch.zeros(m[0], dtype=m[1], device=m[2]) for m in ctx.meta]
        return None, g[0], g[1]


class MinMaxQuantMatMul(nn.Module):
    """reference: quant_layers/matmul.py:8-60"""
    sos = False

    def __init__(self, A_bit=8, B_bit=8, mode="raw"):
        super().__init__()
        self.A_bit = A_bit
        self.B_bit = B_bit
        self.A_interval = None
        self.B_interval = None
        self.A_qmax = 2 ** (self.A_bit - 1)
        self.B_qmax = 2 ** (self.B_bit - 1)
        self.mode = mode
        self.raw_input = None
        self.raw_out = None

    def forward(self, A, B):
        if self.mode == "raw":
            out = A @ B
        elif self.mode == "quant_forward":
            out = self.quant_forward(A, B)
        elif self.mode == "calibration_step1":
            out = self.calibration_step1(A, B)
        elif self.mode == "calibration_step2":
            out = self.calibration_step2(A, B)
        else:
            raise NotImplementedError
        return out

    def quant_input(self, x, interval, qmax):
        x_sim = (x / interval).round_().clamp_(-qmax, qmax - 1)
        x_sim.mul_(interval)
        return x_sim

    def calibration_step1(self, A, B):
        self.raw_input = A.cpu().detach(), B.cpu().detach()
        out = A @ B
        self.raw_out = out.cpu().detach()
        return out

    def calibration_step2(self, A, B):
        """reference: matmul.py:54-60 (layer-wise min-max)"""
        H = A.shape[1]
        self.A_interval = (A.data.abs().max() / (self.A_qmax - 0.5)).detach().view(1, 1, 1, 1, 1, 1, 1).repeat(1, H, 1, 1, 1, 1, 1)
        self.B_interval = (B.data.abs().max() / (self.B_qmax - 0.5)).detach().view(1, 1, 1, 1, 1, 1, 1).repeat(1, H, 1, 1, 1, 1, 1)
        self.calibrated = True
        return self.quant_forward(A, B)

    # ---- native plumbing -------------------------------------------------
    def _desc(self, A, B, search_round=1, eq=(0.0, 1.0, 1)):
        assert A.dim() == 4 and B.dim() == 4 and A.shape[:2] == B.shape[:2] and A.shape[3] == B.shape[2], \
            f"expected A [b,H,S1,S2] and B [b,H,S2,S3], got {tuple(A.shape)} and {tuple(B.shape)}"
        d = _lib.MatMulDesc()
        d.batch, d.heads, d.S1, d.S2, d.S3 = int(A.shape[0]), int(A.shape[1]), int(A.shape[2]), int(A.shape[3]), int(B.shape[3])
        d.A_bit, d.B_bit = int(self.A_bit), int(self.B_bit)
        d.eq_n, d.search_round = int(eq[2]), int(search_round)
        d.eq_alpha, d.eq_beta = float(eq[0]), float(eq[1])
        d.sos = 1 if self.sos else 0
        d.operand = _lib.default_operand()
        d.kernel = _lib.default_kernel()
        d.init_layerwise = 1 if getattr(self, "init_layerwise", False) else 0
        return d

    @staticmethod
    def _cuda(t):
        if t.device.type != "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError("ptq4vit_b200 MatMul quant layers need a CUDA device (no CPU path)")
            t = t.cuda()
        return t.contiguous().float()

    def quant_forward(self, A, B):
        """reference: matmul.py:40-45 / :140-145 -- fq(A) @ fq(B) on the tensor cores."""
        assert self.calibrated is not None, f"You should run calibrate_forward before run quant_forward for {self}"
        if torch.is_grad_enabled() and (A.requires_grad or B.requires_grad):
            return _QuantMatMulFn.apply(self, A, B)
        return self._quant_forward_native(A, B)

    def _quant_forward_native(self, A, B):
        A_, B_ = self._cuda(A), self._cuda(B)
        dev = A_.device
        d = self._desc(A_, B_)
        lib = _lib.lib()
        nbytes = ctypes.c_size_t()
        _lib.check(lib.p4v_matmul_quant_forward_workspace_bytes(ctypes.byref(d), ctypes.byref(nbytes)), "matmul_quant_forward_workspace")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        out = torch.empty(d.batch, d.heads, d.S1, d.S3, dtype=torch.float32, device=dev)
        H = d.heads
        a_int = torch.as_tensor(self.A_interval, dtype=torch.float32, device=dev).reshape(-1)
        a_int = (a_int if a_int.numel() == H or self.sos else a_int.expand(H)).contiguous()
        b_int = torch.as_tensor(self.B_interval, dtype=torch.float32, device=dev).reshape(-1)
        b_int = (b_int if b_int.numel() == H else b_int.expand(H)).contiguous()
        split = torch.as_tensor(self.split, dtype=torch.float32, device=dev).reshape(1) if self.sos else None
        _lib.check(lib.p4v_matmul_quant_forward(ctypes.byref(d), _lib.ptr(A_), _lib.ptr(B_), _lib.ptr(a_int), _lib.ptr(b_int),
                                                _lib.ptr(split), _lib.ptr(ws), nbytes.value, _lib.ptr(out),
                                                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "p4v_matmul_quant_forward")
        return out


class PTQSLQuantMatMul(MinMaxQuantMatMul):
    """reference: quant_layers/matmul.py:62-282.  Block structure: only the head-wise layout that
    the Batching classes force (n_G = heads, n_V = n_H = 1) is implemented by the B200 path."""

    def __init__(self, A_bit=8, B_bit=8, mode="raw", metric="L2_norm", search_round=1, eq_alpha=0.1, eq_beta=2,
                 eq_n=100, parallel_eq_n=10, n_G_A=1, n_V_A=1, n_H_A=1, n_G_B=1, n_V_B=1, n_H_B=1, init_layerwise=False):
        super().__init__(A_bit=A_bit, B_bit=B_bit, mode=mode)
        self.metric = metric
        self.search_round = search_round
        self.eq_alpha = eq_alpha
        self.eq_beta = eq_beta
        self.eq_n = eq_n
        self.parallel_eq_n = parallel_eq_n
        self.n_G_A, self.n_V_A, self.n_H_A = n_G_A, n_V_A, n_H_A
        self.n_G_B, self.n_V_B, self.n_H_B = n_G_B, n_V_B, n_H_B
        self.crb_groups_A = self.crb_groups_B = None
        self.crb_rows_A = self.crb_cols_A = self.crb_rows_B = self.crb_cols_B = None
        self.pad_groups_A = self.pad_groups_B = None
        self.pad_rows_A = self.pad_rows_B = self.pad_cols_A = self.pad_cols_B = None
        self.raw_grad = None
        self.init_layerwise = init_layerwise
        self.split = None
        self.keep_scores = False
        self.last_scores = None

    _force_headwise = False       # the Batching classes set n_G = heads (matmul.py:411-417)

    def _get_padding_parameters(self, A, B):
        """reference: matmul.py:109-122 (groups of consecutive heads, zero padding).  The B200 path implements the two
        layouts PTQ4ViT meets: one group per head (what the Batching classes force, :411-417) and one group for all
        heads (the constructor default n_G = 1 of the non-batching classes)."""
        H = A.shape[1]
        if self._force_headwise:
            self.n_G_A = self.n_G_B = H
        for n_G in (self.n_G_A, self.n_G_B):
            if n_G not in (1, H):
                raise NotImplementedError(f"ptq4vit_b200 MatMul search: n_G must be 1 or the number of heads ({H}), got {n_G}")
        self.crb_groups_A, self.crb_groups_B = H // self.n_G_A, H // self.n_G_B
        self.crb_rows_A, self.crb_cols_A = A.shape[2], A.shape[3]
        self.crb_rows_B, self.crb_cols_B = B.shape[2], B.shape[3]
        self.pad_groups_A = self.pad_groups_B = 0
        self.pad_rows_A = self.pad_rows_B = self.pad_cols_A = self.pad_cols_B = 0

    def _grad_for_metric(self, y):
        """Per-element weight of the metric (matmul.py:465-479); see _metric.py."""
        return metric_weight(self.metric, y, self.raw_grad, "PTQSLBatchingQuantMatMul")

    def _native_calibrate(self, A, B, Y, G):
        if (self.n_V_A, self.n_H_A, self.n_V_B, self.n_H_B) != (1, 1, 1, 1):
            raise NotImplementedError("ptq4vit_b200 MatMul search implements the head-wise layout only (n_V = n_H = 1)")
        A_, B_, Y_, G_ = self._cuda(A), self._cuda(B), self._cuda(Y), self._cuda(G)
        dev = A_.device
        self._get_padding_parameters(A_, B_)
        n_G = self.n_G_B if self.sos else self.n_G_A
        if not self.sos and self.n_G_A != self.n_G_B:
            raise NotImplementedError("ptq4vit_b200 MatMul search: A and B must use the same group layout")
        if n_G == 1 and A_.shape[1] > 1:      # one group for all heads: the heads become part of the batch
            A_, B_, Y_, G_ = [t.reshape(-1, 1, t.shape[2], t.shape[3]) for t in (A_, B_, Y_, G_)]
        d = self._desc(A_, B_, self.search_round, (self.eq_alpha, self.eq_beta, self.eq_n))
        H = d.heads
        lib = _lib.lib()
        nbytes, nlog = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(lib.p4v_matmul_workspace_bytes(ctypes.byref(d), ctypes.byref(nbytes)), "p4v_matmul_workspace_bytes")
        _lib.check(lib.p4v_matmul_score_log_floats(ctypes.byref(d), ctypes.byref(nlog)), "p4v_matmul_score_log_floats")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        a_int = torch.empty(H, dtype=torch.float32, device=dev)
        b_int = torch.empty(H, dtype=torch.float32, device=dev)
        split = torch.empty(1, dtype=torch.float32, device=dev) if self.sos else None
        log = torch.empty(nlog.value, dtype=torch.float32, device=dev) if self.keep_scores else None
        _lib.check(lib.p4v_matmul_calibrate(ctypes.byref(d), _lib.ptr(A_), _lib.ptr(B_), _lib.ptr(Y_), _lib.ptr(G_), _lib.ptr(ws),
                                            nbytes.value, _lib.ptr(a_int), _lib.ptr(b_int), _lib.ptr(split), _lib.ptr(log),
                                            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "p4v_matmul_calibrate")
        if self.sos:
            self.split = split[0]
            self.A_interval = a_int[0]
        else:
            self.A_interval = a_int.view(1, H, 1, 1, 1, 1, 1)
        self.B_interval = b_int.view(1, H, 1, 1, 1, 1, 1)
        if log is not None:
            out, o = [], 0
            for _ in range(self.search_round):
                n1 = 20 if self.sos else self.eq_n * H
                out.append(log[o:o + n1] if self.sos else log[o:o + n1].view(self.eq_n, H)); o += n1
                out.append(log[o:o + self.eq_n * H].view(self.eq_n, H)); o += self.eq_n * H
            self.last_scores = out

    def calibration_step2(self, A, B):
        """reference: matmul.py:257-282"""
        Y = self.raw_out
        self._native_calibrate(A, B, Y, self._grad_for_metric(Y))
        self.calibrated = True
        del self.raw_input, self.raw_out, self.raw_grad
        return self.quant_forward(A, B)


class SoSPTQSLQuantMatMul(PTQSLQuantMatMul):
    """reference: quant_layers/matmul.py:284-388 (split-of-softmax, twin-uniform A)"""
    sos = True

    def __init__(self, *args, split=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.n_G_A = self.n_V_A = self.n_H_A = 1
        self.A_qmax = 2 ** (self.A_bit - 1)
        self.split = split
        if split is not None:
            self.A_interval = self.split / (self.A_qmax - 1)


class PTQSLBatchingQuantMatMul(PTQSLQuantMatMul):
    """reference: quant_layers/matmul.py:390-576"""
    _force_headwise = True

    def _initialize_calib_parameters(self):
        """reference: matmul.py:396-409; a whole layer fits in HBM, no batching."""
        self.calib_size = int(self.raw_input[0].shape[0])
        self.calib_batch_size = int(self.raw_input[0].shape[0])
        self.calib_need_batching = False

    def calibration_step2(self):
        """reference: matmul.py:565-576"""
        self._initialize_calib_parameters()
        Y = self.raw_out
        self._native_calibrate(self.raw_input[0], self.raw_input[1], Y, self._grad_for_metric(Y))
        self.calibrated = True
        del self.raw_input, self.raw_out, self.raw_grad


class SoSPTQSLBatchingQuantMatMul(PTQSLBatchingQuantMatMul):
    """reference: quant_layers/matmul.py:578-644"""
    sos = True

    def __init__(self, *args, split=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.n_G_A = self.n_V_A = self.n_H_A = 1
        self.A_qmax = 2 ** (self.A_bit - 1)
        self.split = split
        if split is not None:
            self.A_interval = self.split / (self.A_qmax - 1)
