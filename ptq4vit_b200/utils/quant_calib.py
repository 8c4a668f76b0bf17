"""Calibration drivers with the reference's surface (utils/quant_calib.py:9-378).

`HessianQuantCalibrator(net, wrapped_modules, calib_loader, sequential=False, batch_size=1)
.batching_quant_calib()` is the entry point the reference's experiments time
(example/test_all.py:31-34).  B200-first changes, results unchanged:

* capture: with sequential=False every module stays in "raw" mode while the others calibrate
  (quant_calib.py:369-372), so the captured (input, output, grad) tensors do not depend on the
  order -- ONE forward+backward sweep over the calibration images with hooks on all modules
  replaces the reference's one-sweep-per-module loop (quant_calib.py:317-356), and the tensors
  stay in HBM instead of bouncing through host memory (quant_calib.py:173-201).  When the
  tensors do not fit (`capture="per_module"`) or sequential=True the reference's loop is used.
* search: `module.calibration_step2()` runs the CUDA search.
* multi-GPU: modules are independent => static LPT sharding over ranks, every rank captures,
  searches its share, and one all_gather of the chosen step sizes ends the job.
"""
import math

import torch
import torch.nn.functional as F

from ..quant_layers.linear import MinMaxQuantLinear
from ..quant_layers.matmul import MinMaxQuantMatMul


# ---------------------------------------------------------------- hooks (device resident)
def grad_hook(module, grad_input, grad_output):
    if module.raw_grad is None:
        module.raw_grad = []
    module.raw_grad.append(grad_output[0].detach())


def linear_forward_hook(module, input, output):
    if module.raw_input is None:
        module.raw_input = []
    if module.raw_out is None:
        module.raw_out = []
    module.raw_input.append(input[0].detach())
    module.raw_out.append(output.detach())


def matmul_forward_hook(module, input, output):
    if module.raw_input is None:
        module.raw_input = [[], []]
    if module.raw_out is None:
        module.raw_out = []
    module.raw_input[0].append(input[0].detach())
    module.raw_input[1].append(input[1].detach())
    module.raw_out.append(output.detach())


def _cat_captured(module):
    if isinstance(module, MinMaxQuantLinear):
        module.raw_input = torch.cat(module.raw_input, dim=0)
        module.raw_out = torch.cat(module.raw_out, dim=0)
    if isinstance(module, MinMaxQuantMatMul):
        module.raw_input = [torch.cat(_, dim=0) for _ in module.raw_input]
        module.raw_out = torch.cat(module.raw_out, dim=0)
    if getattr(module, "raw_grad", None) is not None and isinstance(module.raw_grad, list):
        module.raw_grad = torch.cat(module.raw_grad, dim=0)


# ---------------------------------------------------------------- work model + sharding
def module_cost(module, n_img, tokens_hint=197):
    """Relative cost of one module's search in candidate-GEMM operations."""
    if isinstance(module, MinMaxQuantLinear):
        rounds = getattr(module, "search_round", 1)
        units = rounds * (getattr(module, "n_H", 1) + getattr(module, "n_a", 1)) * getattr(module, "eq_n", 1)
        return float(units) * 2.0 * n_img * tokens_hint * module.in_features * module.out_features / max(1, getattr(module, "n_H", 1))
    if isinstance(module, MinMaxQuantMatMul):
        rounds = getattr(module, "search_round", 1)
        return float(rounds * 2 * getattr(module, "eq_n", 1)) * 2.0 * n_img * 12 * tokens_hint * tokens_hint * 64
    return 0.0


def shard_modules(names, costs, world_size):
    """Static longest-processing-time assignment: returns owner[name] = rank.  Deterministic."""
    order = sorted(range(len(names)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    owner = {}
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[names[i]] = r
        load[r] += costs[i]
    return owner


def pack_result(module, width):
    """Flatten a calibrated module's step sizes into a fixed-width fp32 row for the all_gather."""
    vals = []
    if isinstance(module, MinMaxQuantLinear):
        vals += [torch.as_tensor(module.w_interval, dtype=torch.float32).reshape(-1),
                 torch.as_tensor(module.a_interval, dtype=torch.float32).reshape(-1)]
    else:
        vals += [torch.as_tensor(module.A_interval, dtype=torch.float32).reshape(-1),
                 torch.as_tensor(module.B_interval, dtype=torch.float32).reshape(-1)]
        if getattr(module, "sos", False):
            vals.append(torch.as_tensor(module.split, dtype=torch.float32).reshape(-1))
    flat = torch.cat([v.to(vals[0].device) for v in vals])
    assert flat.numel() <= width, f"result of {type(module).__name__} does not fit the gather row ({flat.numel()} > {width})"
    row = torch.zeros(width, dtype=torch.float32, device=flat.device)
    row[:flat.numel()] = flat
    return row


def unpack_result(module, row, heads=None):
    """Inverse of pack_result (the module's static block structure gives the split points)."""
    if isinstance(module, MinMaxQuantLinear):
        nw = module.n_V * module.n_H
        module.w_interval = row[:nw].clone().view(module.n_V, 1, module.n_H, 1)
        module.a_interval = row[nw:nw + module.n_a].clone().view(module.n_a, 1)
    else:
        H = heads if heads is not None else module.n_G_B
        if getattr(module, "sos", False):
            module.A_interval = row[0].clone()
            module.B_interval = row[1:1 + H].clone().view(1, H, 1, 1, 1, 1, 1)
            module.split = row[1 + H].clone()
        else:
            module.A_interval = row[:H].clone().view(1, H, 1, 1, 1, 1, 1)
            module.B_interval = row[H:2 * H].clone().view(1, H, 1, 1, 1, 1, 1)
    module.calibrated = True


def result_width(modules):
    w = 1
    for m in modules:
        if isinstance(m, MinMaxQuantLinear):
            w = max(w, m.n_V * m.n_H + m.n_a)
        else:
            w = max(w, 2 * 64 + 1)        # up to 64 heads + split
    return w


# ---------------------------------------------------------------- calibrators
# ---------------------------------------------------------------- host-resident captures
def search_from_host(items, device, out_host=None):
    """Search a list of modules whose captured tensors live in (pinned) HOST memory, as the reference's calibrators
    keep them (`.cpu()` in its hooks, `.cuda()` per module before the search; quant_calib.py:173-201, :317-356).

    items: [(module, {"x"|"A","B", "y", "g": pinned cpu tensors})].  The host->device copies of module i+1 run on a
    side stream while module i searches; the chosen step sizes are copied back asynchronously into pinned buffers and
    one synchronisation ends the call.  Returns (h2d_bytes, d2h_bytes)."""
    import os, time
    dbg = os.environ.get("P4V_E2E_DEBUG")
    t_stage = t_cal = t_out = 0.0
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    h2d = d2h = 0

    def stage(i):
        with torch.cuda.stream(copy_stream):
            dev = {k: v.to(device, non_blocking=True) for k, v in items[i][1].items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dev, ev

    results = []
    nxt = stage(0) if items else None
    for i, (m, hb) in enumerate(items):
        dev, ev = nxt
        t0 = time.perf_counter()
        nxt = stage(i + 1) if i + 1 < len(items) else None
        t_stage += time.perf_counter() - t0
        main.wait_event(ev)
        for v in dev.values():
            v.record_stream(main)
        h2d += sum(v.numel() * v.element_size() for v in dev.values())
        if "x" in dev:
            m.raw_input, m.raw_out, m.raw_grad = dev["x"], dev["y"], dev["g"]
        else:
            m.raw_input, m.raw_out, m.raw_grad = [dev["A"], dev["B"]], dev["y"], dev["g"]
        t0 = time.perf_counter()
        with torch.no_grad():
            m.calibration_step2()
        t_cal += time.perf_counter() - t0
        t0 = time.perf_counter()
        m.raw_input = m.raw_out = m.raw_grad = None
        outs = [m.w_interval, m.a_interval] if "x" in dev else [torch.as_tensor(m.A_interval, device=device), m.B_interval]
        for j, o in enumerate(outs):
            o = torch.as_tensor(o, device=device).detach().reshape(-1).float()
            buf = torch.empty(o.numel(), dtype=torch.float32, pin_memory=True) if out_host is None else out_host[i][j]
            buf.copy_(o, non_blocking=True)
            results.append(buf)
            d2h += o.numel() * 4
        t_out += time.perf_counter() - t0
    t0 = time.perf_counter()
    main.synchronize()
    if dbg:
        print(f"[search_from_host] host time: stage {t_stage:.3f}s calibrate {t_cal:.3f}s outputs {t_out:.3f}s final sync {time.perf_counter() - t0:.3f}s", flush=True)
    return h2d, d2h


class QuantCalibrator():
    """reference: utils/quant_calib.py:9-171"""

    def __init__(self, net, wrapped_modules, calib_loader, sequential=True):
        self.net = net
        self.wrapped_modules = wrapped_modules
        self.calib_loader = calib_loader
        self.sequential = sequential
        self.calibrated = False
        self.batch_size = getattr(calib_loader, "batch_size", 1)   # the reference forgets this attribute (quant_calib.py:131)

    def _device(self):
        return next(self.net.parameters()).device

    def quant_calib(self):
        """reference: quant_calib.py:95-104 (step1: collect, step2: per-module search on the cached tensors)"""
        for name, module in self.wrapped_modules.items():
            module.mode = "raw" if hasattr(module, "calibrated") else "calibration_step1"
        with torch.no_grad():
            for inp, target in self.calib_loader:
                self.net(inp.to(self._device()))
        for name, module in self.wrapped_modules.items():
            if hasattr(module, "calibrated"):
                continue
            module.mode = "calibration_step2"
            with torch.no_grad():
                if isinstance(module, MinMaxQuantLinear):
                    module.forward(module.raw_input.to(self._device()))
                elif isinstance(module, MinMaxQuantMatMul):
                    module.forward(module.raw_input[0].to(self._device()), module.raw_input[1].to(self._device()))
        for name, module in self.wrapped_modules.items():
            module.mode = "quant_forward"
        self.calibrated = True


class HessianQuantCalibrator(QuantCalibrator):
    """reference: utils/quant_calib.py:203-378"""

    def __init__(self, net, wrapped_modules, calib_loader, sequential=False, batch_size=1, capture="auto",
                 distributed=None, target_noise=0.0):
        super().__init__(net, wrapped_modules, calib_loader, sequential=sequential)
        self.batch_size = batch_size
        self.capture = capture
        self.distributed = distributed
        self.target_noise = target_noise     # synthetic benches: perturb the KL target so that gradients are not ~0
        self.timings = {}

    # -- target distribution (quant_calib.py:308-313)
    def _raw_pred_softmax(self):
        dev = self._device()
        preds = []
        with torch.no_grad():
            for inp, _ in self.calib_loader:
                preds.append(F.softmax(self.net(inp.to(dev)), dim=-1).detach())
        return torch.cat(preds, dim=0)

    def _fwd_bwd(self, raw_pred_softmax):
        """quant_calib.py:333-341: KL(self) backward in mini-batches of self.batch_size."""
        dev = self._device()
        off = 0
        for inp, target in self.calib_loader:
            n = inp.shape[0]
            for batch_st in range(0, n, self.batch_size):
                self.net.zero_grad()
                inp_ = inp[batch_st:batch_st + self.batch_size].to(dev)
                pred = self.net(inp_)
                tgt = raw_pred_softmax[off + batch_st:off + batch_st + self.batch_size]
                loss = F.kl_div(F.log_softmax(pred, dim=-1), tgt, reduction="batchmean")
                loss.backward()
            off += n

    def _hooks_for(self, module):
        hooks = []
        if isinstance(module, MinMaxQuantLinear):
            hooks.append(module.register_forward_hook(linear_forward_hook))
        if isinstance(module, MinMaxQuantMatMul):
            hooks.append(module.register_forward_hook(matmul_forward_hook))
        if hasattr(module, "metric"):
            hooks.append(module.register_full_backward_hook(grad_hook))
        return hooks

    def _my_modules(self):
        """Layer-wise sharding (only meaningful when sequential=False)."""
        names = list(self.wrapped_modules.keys())
        dist = self.distributed
        if dist is None or not dist.is_initialized() or dist.get_world_size() == 1 or self.sequential:
            return names, None
        n_img = sum(inp.shape[0] for inp, _ in self.calib_loader)
        costs = [module_cost(self.wrapped_modules[n], n_img) for n in names]
        owner = shard_modules(names, costs, dist.get_world_size())
        return [n for n in names if owner[n] == dist.get_rank()], owner

    def _gather(self, owner):
        dist = self.distributed
        names = list(self.wrapped_modules.keys())
        mods = [self.wrapped_modules[n] for n in names]
        width = result_width(mods)
        dev = self._device()
        mine = torch.zeros(len(names), width, dtype=torch.float32, device=dev)
        heads = {}
        for i, n in enumerate(names):
            if owner[n] == dist.get_rank():
                mine[i] = pack_result(mods[i], width).to(dev)
        # head counts are only known to the owner of a MatMul module: ship them in the last column
        meta = torch.zeros(len(names), dtype=torch.float32, device=dev)
        for i, n in enumerate(names):
            if owner[n] == dist.get_rank() and isinstance(mods[i], MinMaxQuantMatMul):
                meta[i] = float(mods[i].n_G_B)
        gathered = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        gmeta = [torch.zeros_like(meta) for _ in range(dist.get_world_size())]
        dist.all_gather(gathered, mine)           # the one collective of the job
        dist.all_gather(gmeta, meta)
        for i, n in enumerate(names):
            r = owner[n]
            if r != dist.get_rank():
                h = int(gmeta[r][i].item()) if isinstance(mods[i], MinMaxQuantMatMul) else None
                unpack_result(mods[i], gathered[r][i], heads=h)

    def batching_quant_calib(self):
        """reference: quant_calib.py:300-378"""
        raw_pred_softmax = self._raw_pred_softmax()
        if self.target_noise > 0:
            gen = torch.Generator(device=raw_pred_softmax.device).manual_seed(1234)
            logits = raw_pred_softmax.clamp_min(1e-30).log()
            logits = logits + self.target_noise * torch.randn(logits.shape, generator=gen, device=logits.device)
            raw_pred_softmax = F.softmax(logits, dim=-1)
        my_names, owner = self._my_modules()
        single_pass = (not self.sequential) and self.capture in ("auto", "single_pass")
        if single_pass:
            hooks = []
            for name in my_names:
                hooks += self._hooks_for(self.wrapped_modules[name])
            self._fwd_bwd(raw_pred_softmax)
            for h in hooks:
                h.remove()
            for name in my_names:
                module = self.wrapped_modules[name]
                _cat_captured(module)
                with torch.no_grad():
                    module.calibration_step2()
                module.mode = "raw"
        else:
            for name in my_names:
                module = self.wrapped_modules[name]
                hooks = self._hooks_for(module)
                self._fwd_bwd(raw_pred_softmax)
                _cat_captured(module)
                for h in hooks:
                    h.remove()
                with torch.no_grad():
                    module.calibration_step2()
                module.mode = "quant_forward" if self.sequential else "raw"
        if owner is not None:
            self._gather(owner)
        for name, module in self.wrapped_modules.items():
            module.mode = "quant_forward"
        self.calibrated = True

    # the reference's non-batching entry point maps onto the same machinery (quant_calib.py:216-298)
    def quant_calib(self):
        return self.batching_quant_calib()
