"""CUDA-graph replay of the whole-model forwards (face_mask_inpaint_b200/graphs.py) equals the eager forward: same kernels,
same order, bit-identical where no atomics are involved. SpectralNorm's u/v advance by one power iteration per call in both
(external_function.py:44-57), so call k of the eager model is compared with call k of the captured one."""
import copy
import types

import pytest
import torch

from conftest import rel_err
from golden_util import fill_by_name, mean_z, picnet_inputs, refpsp_inputs

pytestmark = pytest.mark.gpu


def _picnet():
    from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
    m = fill_by_name(build_picnet_ref()).eval()
    m.decoder.get_z = types.MethodType(mean_z, m.decoder)
    return m


def test_picnet_graph_replay_matches_eager():
    from face_mask_inpaint_b200.graphs import CapturedForward
    base = _picnet()
    eager, graphed = copy.deepcopy(base).cuda(), copy.deepcopy(base).cuda()
    eager.decoder.get_z = types.MethodType(mean_z, eager.decoder)
    graphed.decoder.get_z = types.MethodType(mean_z, graphed.decoder)
    src, ref, mask = (t.cuda() for t in picnet_inputs(2))
    src2, ref2, mask2 = (t.cuda() for t in picnet_inputs(2, seed=8))
    fwd = CapturedForward(graphed, src, ref, mask, warmup=3)     # 3 eager calls inside; capture itself runs nothing
    with torch.no_grad():
        for _ in range(3):
            eager(src, ref, mask)
        want1 = eager(src, ref, mask)
        want2 = eager(src2, ref2, mask2)
    got1 = fwd(src, ref, mask).clone()
    got2 = fwd(src2, ref2, mask2).clone()
    assert got1.shape == want1.shape == (2, 3, 256, 256)
    assert rel_err(got1, want1) <= 1e-5, rel_err(got1, want1)
    assert rel_err(got2, want2) <= 1e-5, rel_err(got2, want2)
    assert rel_err(got1, got2) > 1e-3          # the replay really consumed the new inputs
    u_e = eager.decoder.decoder0.conv1.module.weight_u
    u_g = graphed.decoder.decoder0.conv1.module.weight_u
    assert rel_err(u_g, u_e) <= 1e-5           # power-iteration state advanced under replay as in eager mode


def test_picnet_graph_rejects_other_shapes():
    from face_mask_inpaint_b200.graphs import CapturedForward
    m = _picnet().cuda()
    src, ref, mask = (t.cuda() for t in picnet_inputs(1))
    fwd = CapturedForward(m, src, ref, mask, warmup=1)
    with pytest.raises(RuntimeError):
        fwd(*(t.cuda() for t in picnet_inputs(2)))
    with pytest.raises(RuntimeError):
        CapturedForward(m, src.cpu(), ref.cpu(), mask.cpu())


def test_refpsp_graph_replay_matches_eager():
    from face_mask_inpaint_b200.graphs import CapturedForward
    from face_mask_inpaint_b200.modules.psp import pSp, refpsp_opts
    net = fill_by_name(pSp(refpsp_opts(output_size=256))).eval().cuda()
    x, ref, mask = (t.cuda() for t in refpsp_inputs(2))
    x2, ref2, mask2 = (t.cuda() for t in refpsp_inputs(2, seed=12))
    fwd = CapturedForward(net, x, ref=ref, src_mask=mask, resize=True, randomize_noise=False)
    with torch.no_grad():
        want1 = net(x, ref=ref, src_mask=mask, resize=True, randomize_noise=False)
        want2 = net(x2, ref=ref2, src_mask=mask2, resize=True, randomize_noise=False)
    got1 = fwd(x, ref=ref, src_mask=mask).clone()
    got2 = fwd(x2, ref=ref2, src_mask=mask2).clone()
    assert rel_err(got1, want1) <= 1e-5, rel_err(got1, want1)
    assert rel_err(got2, want2) <= 1e-5, rel_err(got2, want2)


def test_auto_graph_wrapper_keys_graphs_by_shape_and_falls_back_under_autograd():
    """graphs.auto_graph (what the installer puts around ReferenceFill.forward with FMI_CUDA_GRAPH=1): one captured graph per
    input shape, clones returned, eager forward whenever autograd is on or the module is in training mode."""
    from face_mask_inpaint_b200.graphs import auto_graph, module_cache
    from face_mask_inpaint_b200.modules.picnet import ReferenceFill
    base = _picnet()
    eager, wrapped = copy.deepcopy(base).cuda(), copy.deepcopy(base).cuda()
    for m in (eager, wrapped):
        m.decoder.get_z = types.MethodType(mean_z, m.decoder)

    class Graphed(ReferenceFill):
        forward = auto_graph(ReferenceFill.forward)

    wrapped.__class__ = Graphed
    a1 = tuple(t.cuda() for t in picnet_inputs(1))
    a2 = tuple(t.cuda() for t in picnet_inputs(2, seed=9))
    with torch.no_grad():
        got = [wrapped(*a1), wrapped(*a2), wrapped(*a1)]
        assert len(module_cache(wrapped)["graphs"]) == 2           # one graph per shape, the third call replays the first
    assert got[0].shape == (1, 3, 256, 256) and got[1].shape == (2, 3, 256, 256)
    assert got[0].data_ptr() != got[2].data_ptr()                  # clones, not the graph's static output
    # the eager model with the same number of forwards (= SpectralNorm power iterations): a capture is 3 warm-up calls, then
    # every wrapped call is one replay
    with torch.no_grad():
        want = []
        for args, calls in ((a1, 4), (a2, 4), (a1, 1)):
            for _ in range(calls):
                o = eager(*args)
            want.append(o)
    for g, w in zip(got, want):
        assert rel_err(g, w) <= 1e-5, rel_err(g, w)
    # autograd on: no graph, gradients flow
    wrapped.train()
    out = wrapped(*a1)
    assert out.requires_grad and len(module_cache(wrapped)["graphs"]) == 2


def test_captured_train_step_matches_eager():
    """graphs.CapturedStep: forward + loss + backward + Adam(capturable) of a SpectralNorm conv block as ONE CUDA graph; after k
    replays parameters, Adam state and SpectralNorm's u / v equal k eager steps (same kernels, same order)."""
    import copy
    from torch import nn
    from face_mask_inpaint_b200.graphs import CapturedStep
    from face_mask_inpaint_b200.modules.picnet import ResBlock
    torch.manual_seed(1)
    net = ResBlock(32, 32, 32, norm_layer=None, nonlinearity=nn.LeakyReLU(0.1), sample_type='down', use_spect=True).cuda()
    ref = copy.deepcopy(net)
    x = torch.randn(2, 32, 16, 16, device="cuda")
    xs = [torch.randn(2, 32, 16, 16, device="cuda") for _ in range(3)]

    def make_step(m, opt):
        def step(inp):
            loss = (m(inp) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss
        return step

    opt_a = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)
    opt_b = torch.optim.Adam(ref.parameters(), lr=1e-3, capturable=True)
    cap = CapturedStep(make_step(net, opt_a), x, warmup=3, modules=(net,))     # 3 warm-up steps on x (the capture itself runs nothing)
    eager = make_step(ref, opt_b)
    for _ in range(3):
        eager(x)
    for inp in xs:
        lg = cap(inp).clone()
        le = eager(inp)
        assert abs(float(lg) - float(le)) <= 1e-5 * max(1.0, abs(float(le)))
    for (n1, p1), (_, p2) in zip(net.named_parameters(), ref.named_parameters()):
        assert (p1 - p2).abs().max().item() <= 1e-5 * max(1.0, p2.abs().max().item()), n1
