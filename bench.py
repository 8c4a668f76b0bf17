#!/usr/bin/env python
"""bench.py -- candidate-GEMMs/s of the PTQ4ViT scale-factor search on B200.

A "step" = the full `calibration_step2` search of every wrapped Linear / MatMul of the workload
(ViT-B/224, 32 synthetic images, W8A8, n_V=n_H=24 (qkv 72, head 1), n_a=1, eq_n=100, 3 rounds, hessian
metric = BASELINE.json configs[2] at one bit width) over tensors already resident in HBM.
`value` = candidate-GEMM units of ALL ranks / max-over-ranks device time.  `e2e` = the same through the
reference-facing call (`module.calibration_step2()`) with pinned HOST tensors, copies inside the timing.

  python bench.py --gpus 1 --steps 3 --warmup 3
  torchrun ... bench.py --gpus N ...          (layer-sharded, one all_gather of the step sizes per step)
  python bench.py --impl reference            (reference algorithm on the host cores, bounded sample)
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "candidate-GEMMs/s (scale-factor search, ViT-B/224 32-img W8A8)"
UNIT = "cand-GEMM/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="vit_base_patch16_224")
    ap.add_argument("--images", type=int, default=32)
    ap.add_argument("--blocks", type=int, default=24, help="n_V = n_H of the Linear layers (BASELINE: 24)")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--bit", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    return ap.parse_args()


def workload_name(a):
    return (f"{a.model} {a.images} synthetic imgs W{a.bit}A{a.bit} n_V=n_H={a.blocks} (qkv x3, head 1) n_a=1 "
            f"eq_n=100 rounds={a.rounds} hessian")


# ------------------------------------------------------------------ unit accounting
def units_of(module):
    from ptq4vit_b200.quant_layers.linear import MinMaxQuantLinear
    if isinstance(module, MinMaxQuantLinear):
        return module.search_round * (module.n_H + module.n_a) * module.eq_n
    return module.search_round * ((20 if module.sos else module.eq_n) + module.eq_n)


def full_gemm_ops(module, shapes):
    """2*M*K*O per Linear unit / 2*B*H*S1*S2*S3 per MatMul unit."""
    from ptq4vit_b200.quant_layers.linear import MinMaxQuantLinear
    if isinstance(module, MinMaxQuantLinear):
        return 2.0 * shapes["rows"] * module.in_features * module.out_features
    b, h, s1, s2, s3 = shapes["bmm"]
    return 2.0 * b * h * s1 * s2 * s3


# ------------------------------------------------------------------ workload construction
def build_workload(a, device, rank, world):
    """Synthetic ViT + one fwd/bwd capture sweep; returns {name: (module, tensors)} for THIS rank's shard."""
    import importlib
    from ptq4vit_b200.configs import PTQ4ViT as cfg
    from ptq4vit_b200.utils import quant_calib as Q
    from ptq4vit_b200.utils.models import get_net
    from ptq4vit_b200.utils.net_wrap import wrap_modules_in_net
    importlib.reload(cfg)
    for d in (cfg.w_bit, cfg.a_bit, cfg.A_bit, cfg.B_bit):
        for k in d:
            d[k] = a.bit
    cfg.ptqsl_linear_kwargs.update(n_V=a.blocks, n_H=a.blocks, n_a=1, search_round=a.rounds)
    cfg.ptqsl_matmul_kwargs.update(search_round=a.rounds)
    net = get_net(a.model, device=device, seed=0)
    wrapped = wrap_modules_in_net(net, cfg)
    gen = torch.Generator().manual_seed(3)               # mirrors calib_loader(seed=3), utils/datasets.py:88
    size = 384 if "384" in a.model else 224
    images = torch.randn(a.images, 3, size, size, generator=gen)
    loader = [(images, None)]
    names = list(wrapped.keys())
    costs = [Q.module_cost(wrapped[n], a.images) for n in names]
    owner = Q.shard_modules(names, costs, world)
    mine = [n for n in names if owner[n] == rank]
    cal = Q.HessianQuantCalibrator(net, wrapped, loader, sequential=False, batch_size=4, target_noise=1.0)
    raw = cal._raw_pred_softmax()
    g = torch.Generator(device=raw.device).manual_seed(1234)
    logits = raw.clamp_min(1e-30).log() + torch.randn(raw.shape, generator=g, device=raw.device)
    raw = torch.softmax(logits, dim=-1)
    hooks = []
    for n in mine:
        hooks += cal._hooks_for(wrapped[n])
    cal._fwd_bwd(raw)
    for h in hooks:
        h.remove()
    work = {}
    for n in mine:
        m = wrapped[n]
        Q._cat_captured(m)
        if isinstance(m.raw_input, list):
            t = dict(A=m.raw_input[0].contiguous(), B=m.raw_input[1].contiguous(), y=m.raw_out.contiguous(), g=m.raw_grad.contiguous())
        else:
            t = dict(x=m.raw_input.contiguous(), y=m.raw_out.contiguous(), g=m.raw_grad.contiguous())
        m.raw_input = m.raw_out = m.raw_grad = None
        work[n] = (m, t)
    del net, cal
    torch.cuda.empty_cache()
    return wrapped, work, owner, names


def run_module(m, t):
    """One module's search through the reference-facing call, tensors already on the device."""
    if "x" in t:
        m.raw_input, m.raw_out, m.raw_grad = t["x"], t["y"], t["g"]
    else:
        m.raw_input, m.raw_out, m.raw_grad = [t["A"], t["B"]], t["y"], t["g"]
    with torch.no_grad():
        m.calibration_step2()


def gather_results(wrapped, owner, names, dist, device):
    from ptq4vit_b200.utils import quant_calib as Q
    if dist is None:
        return
    cal = Q.HessianQuantCalibrator(torch.nn.Linear(1, 1).to(device), wrapped, [], distributed=dist)
    cal._gather(owner)


class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU reference arm (oracle port)
def cpu_reference_rate(a, seconds_budget):
    """Time the reference algorithm (oracle port: fp32 fake-quant + full GEMM + Hessian-weighted error per
    candidate) on the host cores, one layer of each type at the workload's sizes, a few candidates each,
    and extrapolate: total_time = sum(units_type / rate_type)."""
    from oracle import ptq_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    dims = {"vit_base": (768, 12), "deit_base": (768, 12), "vit_small": (384, 6), "deit_small": (384, 6), "vit_tiny": (192, 3)}
    D, H = next(v for k, v in dims.items() if a.model.startswith(k))
    tok = (384 // 16) ** 2 + 1 if "384" in a.model else 197
    depth = 12
    nb = a.blocks
    lin_types = {"qkv": (D, 3 * D, 3 * nb, False), "proj": (D, D, nb, False), "fc1": (D, 4 * D, nb, False), "fc2": (4 * D, D, nb, True)}
    counts, rates, detail = {}, {}, {}
    per_type_budget = seconds_budget / 6.0
    for name, (K, Oo, nV, gelu) in lin_types.items():
        x, W, b, y, g = O.make_linear_fixture(1, a.images, tok, K, Oo, post_gelu=gelu)
        sp = O.LinearSpec(K, Oo, n_V=nV, n_H=nb, n_a=1, w_bit=a.bit, a_bit=a.bit, eq_n=2, search_round=1, post_gelu=gelu)
        w_int, a_int = O.linear_initial_intervals(sp, W, x)
        f = O.candidate_factors(0.01, 1.2, 100)[:3]
        wc = f.view(-1, 1, 1, 1, 1) * w_int.unsqueeze(0); ac = f.view(1, 1, -1) * a_int.unsqueeze(-1)
        t0 = time.time(); n = 0
        while time.time() - t0 < per_type_budget or n == 0:
            O.linear_search_w(sp, W, b, x, y, g, w_int, a_int, wc, chunk=2, h_list=[0]); n += 2
            O.linear_search_a(sp, W, b, x, y, g, w_int, a_int, ac); n += 2
        dt = time.time() - t0
        rates[name] = n / dt
        counts[name] = depth * a.rounds * (nb + 1) * 100
        detail[name] = {"cands": n, "s": round(dt, 2)}
    for name, (sos, S2, S3) in {"matmul1": (False, D // H, tok), "matmul2": (True, tok, D // H)}.items():
        A, B, Y, G = O.make_matmul_fixture(2, a.images, H, tok, S2, S3, softmax_A=sos)
        sp = O.MatMulSpec(A_bit=a.bit, B_bit=a.bit, eq_n=2, search_round=1, sos=sos)
        A_int, B_int = O.matmul_initial_intervals(sp, A, B)
        fB = O.candidate_factors(0.01, 1.2, 100)[:3].view(-1, 1, 1, 1, 1, 1, 1, 1) * B_int.unsqueeze(0)
        split = torch.tensor(0.5); Ai = split / (sp.A_qmax - 1) if sos else A_int
        t0 = time.time(); n = 0
        while time.time() - t0 < per_type_budget or n == 0:
            O.matmul_search_B(sp, A, B, Y, G, Ai, B_int, fB, split if sos else None); n += 2
        dt = time.time() - t0
        rates[name] = n / dt
        counts[name] = depth * a.rounds * ((20 if sos else 100) + 100)
        detail[name] = {"cands": n, "s": round(dt, 2)}
    total_units = sum(counts.values())
    total_time = sum(counts[k] / rates[k] for k in counts)
    return total_units / total_time, {"per_type_rate": {k: round(v, 3) for k, v in rates.items()}, "timed": detail,
                                      "extrapolated_full_job_s": round(total_time, 1)}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return
        t0 = time.time()
        vals = []
        per = max(5.0, min(a.cpu_seconds, 60.0))
        for i in range(a.warmup + a.steps):
            v, info = cpu_reference_rate(a, per / max(1, a.warmup + a.steps) * 3)
            if i >= a.warmup:
                vals.append(v)
        value = sum(vals) / len(vals)
        out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a)},
               "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "oracle port of the reference search, one layer per type (qkv, proj, fc1, fc2, matmul1, matmul2) at full "
                                          "size, a few candidates each, extrapolated by unit counts; " + json.dumps(info)},
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "wall_s": round(time.time() - t0, 1)}
        print(json.dumps(out))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=device)
        dist = dist_
    from ptq4vit_b200 import _lib
    lib = _lib.lib()

    wrapped, work, owner, names = build_workload(a, device, rank, world)
    my_units = sum(units_of(m) for m, _ in work.values())
    units_t = torch.tensor([float(my_units)], device=device)
    if dist:
        dist.all_reduce(units_t)
    total_units = float(units_t.item())

    def step():
        for m, t in work.values():
            run_module(m, t)
        gather_results(wrapped, owner, names, dist, device)

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    sync_all()
    sampler = ClockSampler(local);
    if rank == 0:
        sampler.start()
    lib.p4v_profile_enable(1)
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    lib.p4v_profile_enable(0)
    sweep_ms, sweep_n, sweep_ops = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
    lib.p4v_profile_collect(ctypes.byref(sweep_ms), ctypes.byref(sweep_n), ctypes.byref(sweep_ops))
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([ms], device=device)
    if dist:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = total_units * a.steps / (ms / 1e3)

    # ---- end to end through the public call with HOST (pinned) tensors
    e2e = None
    if not a.no_e2e:
        host = {}
        h2d = 0
        for n, (m, t) in work.items():
            host[n] = {k: v.cpu().pin_memory() for k, v in t.items()}
            h2d += sum(v.numel() * 4 for v in t.values())
        saved = {n: t for n, (m, t) in work.items()}
        for n in work:
            work[n] = (work[n][0], None)
        del saved
        torch.cuda.empty_cache()
        d2h = 0

        from ptq4vit_b200.utils.quant_calib import search_from_host
        items = [(m, host[n]) for n, (m, _) in work.items()]

        def e2e_step():
            # pinned host tensors -> (copy stream, one module ahead) -> search -> step sizes back to pinned host memory
            nonlocal d2h
            _, d2h = search_from_host(items, device)
            gather_results(wrapped, owner, names, dist, device)

        e2e_step()
        sync_all()
        k_e2e = max(1, min(a.steps, 2))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
        f0.record()
        for _ in range(k_e2e):
            e2e_step()
        f1.record()
        sync_all()
        e2e_clocks = sampler2.stop() if rank == 0 else None
        ems = f0.elapsed_time(f1)
        et = torch.tensor([ems], device=device)
        hb = torch.tensor([float(h2d), float(d2h)], device=device)
        if dist:
            dist.all_reduce(et, op=dist.ReduceOp.MAX); dist.all_reduce(hb)
        e2e = {"value": total_units * k_e2e / (float(et.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(hb[0].item()), "d2h_bytes_per_step": int(hb[1].item()), "steps": k_e2e,
               "ms_per_step": float(et.item()) / k_e2e, "clocks": e2e_clocks}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16, measured)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    achieved = sweep_ops.value / (sweep_ms.value / 1e3) / 1e12 if sweep_ms.value > 0 else 0.0
    roofline = {"bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s", "frac": achieved / bf16_peak,
                # DRAM bytes (read+write) of the dominant sweep launch from profiles/r01_sweep_xstep_qkv.ncu-rep (ncu --set full):
                # the ViT-B qkv activation step, 4.90 ms, 2.26 TFLOP executed; algorithmic bytes of that launch =
                # candidate planes of X 968 MB + y,g once 116 MB + weight image 3.5 MB + partial scores 11.5 MB = 1.10 GB
                "traffic": 1230711296, "traffic_of": "qkv activation-step launch (algorithmic 1.10e9 B)", "kernel": "sweep_tc_kernel", "launches": int(sweep_n.value),
                "avg_launch_ms": sweep_ms.value / max(1, sweep_n.value), "share_of_step": sweep_ms.value / ms,
                "note": "achieved = EXECUTED tensor-core ops (slab-incremental: only the K segment a candidate changes is multiplied) / "
                        "summed CUDA-event time of the sweep launches of this rank; peak = " + peak_src +
                        "; slab sweeps only (the Gram GEMM of the weight steps is a separate kernel: 1.53 PFLOP/s executed, "
                        "profiles/r01_final_ncu_summary.csv); the sweep is bound by the per-accumulator hand-over (fp32 epilogue, "
                        "TMEM load latency, single-warp issue loop) at 32-wide slabs, see DESIGN.md 4.1"}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "int8/bf16-int operands, s32/f32 accumulate, f32 error", "data": "synthetic",
           "config": {"workload": workload_name(a), "parallelism": f"layer-sharded x{world}", "units_per_step": total_units,
                      "l2": "inputs (9.5 GB of staged tensors per step) are larger than L2; no explicit flush"},
           "calib_search_wallclock_s": ms / a.steps / 1e3,
           "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline}
    if e2e:
        out["e2e"] = e2e
    if not a.no_cpu:
        v, info = cpu_reference_rate(a, a.cpu_seconds)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "oracle port, one layer per type at full size, a few candidates each, extrapolated by unit counts; " + json.dumps(info)}
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
