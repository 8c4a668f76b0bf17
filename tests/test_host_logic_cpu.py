"""Host logic that needs no GPU: class surface, factory contract, LPT sharding, result packing and the
world_size-2 gloo all_gather of the chosen step sizes."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ptq4vit_b200.configs import PTQ4ViT as cfg
from ptq4vit_b200.quant_layers import linear as L, matmul as M
from ptq4vit_b200.utils import quant_calib as Q
from ptq4vit_b200.utils.models import VisionTransformer
from ptq4vit_b200.utils.net_wrap import wrap_modules_in_net


def test_constructor_surface_matches_reference():
    m = L.PTQSLBatchingQuantLinear(768, 2304, True, "raw", 8, 8, None, False, "hessian", 3, 0.01, 1.2, 100, 10, 24, 72, 1, False)
    assert (m.n_H, m.n_V, m.n_a, m.crb_rows, m.crb_cols, m.crb_acts) == (24, 72, 1, 32, 32, 768)
    assert m.w_qmax == 128 and m.a_qmax == 128 and m.mode == "raw" and m.raw_grad is None
    assert not hasattr(m, "calibrated")                       # exists only after calibration (quant_calib.py:39)
    p = L.PostGeluPTQSLBatchingQuantLinear(3072, 768, a_bit=6)
    assert abs(p.a_neg_interval - 0.16997124254703522 / 32) < 1e-12
    with pytest.raises(AssertionError):
        L.MinMaxQuantLinear(4, 4, bias_bit=8)
    mm = M.SoSPTQSLBatchingQuantMatMul(A_bit=6, B_bit=6, metric="hessian", split=0.01)
    assert mm.A_qmax == 32 and abs(mm.A_interval - 0.01 / 31) < 1e-12
    x = torch.randn(2, 5, 768)
    assert torch.equal(m(x), torch.nn.functional.linear(x, m.weight, m.bias))    # raw mode
    m.mode = "bogus"
    with pytest.raises(NotImplementedError):
        m(x)


def test_factory_contract():
    import importlib
    importlib.reload(cfg)
    cfg.ptqsl_linear_kwargs["n_V"] = 24; cfg.ptqsl_linear_kwargs["n_H"] = 24
    qkv = cfg.get_module("qlinear_qkv", 768, 2304)
    assert type(qkv) is L.PTQSLBatchingQuantLinear and qkv.n_V == 72 and qkv.n_H == 24 and qkv.metric == "hessian"
    assert type(cfg.get_module("qlinear_MLP_2", 3072, 768)) is L.PostGeluPTQSLBatchingQuantLinear
    assert cfg.get_module("qlinear_classifier", 768, 1000).n_V == 1
    assert type(cfg.get_module("qmatmul_qk")) is M.PTQSLBatchingQuantMatMul
    assert type(cfg.get_module("qmatmul_scorev")) is M.SoSPTQSLBatchingQuantMatMul
    cfg.no_softmax = True
    assert type(cfg.get_module("qmatmul_scorev")) is M.PTQSLBatchingQuantMatMul
    importlib.reload(cfg)


def test_wrap_counts_match_reference_module_inventory():
    import importlib
    importlib.reload(cfg)
    net = VisionTransformer(img_size=32, patch=16, dim=64, depth=2, num_heads=2, num_classes=10)
    wrapped = wrap_modules_in_net(net, cfg)
    assert len(wrapped) == 1 + 2 * 6 + 1                      # patch-embedding conv + per block qkv, proj, fc1, fc2, matmul1, matmul2 + head
    from ptq4vit_b200.quant_layers.conv import ChannelwiseBatchingQuantConv2d
    assert isinstance(net.patch_embed.proj, ChannelwiseBatchingQuantConv2d) and net.patch_embed.proj.n_V == 64 and net.patch_embed.proj.a_bit == 32
    assert isinstance(net.blocks[0].attn.matmul2, M.SoSPTQSLBatchingQuantMatMul)
    out = net(torch.randn(2, 3, 32, 32))
    assert out.shape == (2, 10)


def test_lpt_sharding_is_balanced_and_deterministic():
    names = [f"m{i}" for i in range(74)]
    costs = [float((i * 37) % 11 + 1) for i in range(74)]
    o1 = Q.shard_modules(names, costs, 8)
    assert o1 == Q.shard_modules(names, costs, 8)
    load = [0.0] * 8
    for n, c in zip(names, costs):
        load[o1[n]] += c
    assert max(load) - min(load) <= max(costs)
    assert set(Q.shard_modules(names, costs, 1).values()) == {0}


def test_pack_unpack_roundtrip():
    lin = L.PTQSLBatchingQuantLinear(64, 96, n_V=3, n_H=2, n_a=2)
    lin.w_interval = torch.rand(3, 1, 2, 1); lin.a_interval = torch.rand(2, 1)
    row = Q.pack_result(lin, 200)
    lin2 = L.PTQSLBatchingQuantLinear(64, 96, n_V=3, n_H=2, n_a=2)
    Q.unpack_result(lin2, row)
    assert torch.equal(lin2.w_interval, lin.w_interval) and torch.equal(lin2.a_interval, lin.a_interval) and lin2.calibrated
    sos = M.SoSPTQSLBatchingQuantMatMul()
    sos.A_interval = torch.tensor(0.5 / 127); sos.B_interval = torch.rand(1, 6, 1, 1, 1, 1, 1); sos.split = torch.tensor(0.5)
    sos2 = M.SoSPTQSLBatchingQuantMatMul()
    Q.unpack_result(sos2, Q.pack_result(sos, 200), heads=6)
    assert torch.equal(sos2.B_interval, sos.B_interval) and float(sos2.split) == 0.5


def _gather_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Linear(2, 2)
    mods = {}
    for i in range(5):
        mods[f"lin{i}"] = L.PTQSLBatchingQuantLinear(32, 32, n_V=2, n_H=2, n_a=1)
    mods["mm"] = M.PTQSLBatchingQuantMatMul()
    from ptq4vit_b200.quant_layers.conv import ChannelwiseBatchingQuantConv2d
    mods["conv"] = ChannelwiseBatchingQuantConv2d(3, 40, 4, stride=4, a_bit=32)    # widest row of this set: 40 + 1
    names = list(mods)
    owner = Q.shard_modules(names, [1.0 + i for i in range(len(names))], world)
    for i, n in enumerate(names):
        if owner[n] == rank:
            if n == "conv":
                mods[n].w_interval = torch.arange(40, dtype=torch.float32).view(40, 1, 1, 1) + 0.5
                mods[n].a_interval = torch.tensor([7.0])
            elif n == "mm":
                mods[n].n_G_B = 4
                mods[n].A_interval = torch.full((1, 4, 1, 1, 1, 1, 1), 10.0 + i)
                mods[n].B_interval = torch.full((1, 4, 1, 1, 1, 1, 1), 20.0 + i)
            else:
                mods[n].w_interval = torch.full((2, 1, 2, 1), float(i)); mods[n].a_interval = torch.full((1, 1), 100.0 + i)
    cal = Q.HessianQuantCalibrator(net, mods, [], distributed=dist)
    cal._gather(owner)
    ok = all(float(mods[f"lin{i}"].w_interval.mean()) == float(i) and float(mods[f"lin{i}"].a_interval) == 100.0 + i for i in range(5))
    ok = ok and float(mods["mm"].B_interval.mean()) == 25.0 and mods["mm"].A_interval.shape == (1, 4, 1, 1, 1, 1, 1)
    ok = ok and mods["conv"].w_interval.shape == (40, 1, 1, 1) and float(mods["conv"].w_interval[39]) == 39.5 and float(mods["conv"].a_interval) == 7.0
    ret[rank] = ok
    dist.destroy_process_group()


def test_two_rank_gloo_gather_of_step_sizes():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(120)
        assert p_.exitcode == 0
    assert ret[0] and ret[1]
