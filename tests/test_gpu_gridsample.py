"""GPU parity of the generic grid_sample plugin (C-ABI section 1) against the CPU oracle, reproducing
the reference's own tests third_party/cuda_gridsample_grad2/test3d.py on the new Function:
known-answer inputs (:17-25), gradcheck/gradgradcheck in fp64 incl. OOB + zeros + align_corners=False
(:32-35, :69-72), strided inputs (:80-118), the conv use-case with nondet_tol (:121-151).
Tolerances: forward 1e-5 relative, first/second-order gradients 1e-4 relative (BASELINE.json)."""
import numpy as np
import pytest
import torch
from torch.autograd import grad

from helpers import rel_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def cu():
    from miso_b200 import cuda_gridsample
    return cuda_gridsample


def _cmp_all_orders(image, optical, padding_mode, align_corners, dtype, channels_last=False, tol_f=1e-5, tol_g=1e-4):
    """Forward, first- and second-order gradients vs the gather oracle (the shape of test3d.cmp_with_naive)."""
    img_c = torch.tensor(image, dtype=dtype, requires_grad=True)
    opt_c = torch.tensor(optical, dtype=dtype, requires_grad=True)
    o = O.trilinear_sample(img_c, opt_c, padding_mode, align_corners)
    ol = torch.sum(o ** 2)
    og_i, og_o = grad(ol, [img_c, opt_c], create_graph=True)
    og2_i, og2_o = grad(torch.sum(og_i) + torch.sum(og_o * og_o), [img_c, opt_c])

    img = torch.tensor(image, dtype=dtype, device="cuda")
    if channels_last:
        img = img.contiguous(memory_format=torch.channels_last_3d)
    img.requires_grad_(True)
    opt = torch.tensor(optical, dtype=dtype, device="cuda", requires_grad=True)
    out = cu().grid_sample_3d(img, opt, padding_mode=padding_mode, align_corners=align_corners)
    assert out.shape == o.shape
    l = torch.sum(out ** 2)
    g_i, g_o = grad(l, [img, opt], create_graph=True)
    g2_i, g2_o = grad(torch.sum(g_i) + torch.sum(g_o * g_o), [img, opt])
    assert rel_err(out, o) < tol_f
    assert rel_err(g_i, og_i) < tol_g
    assert rel_err(g_o, og_o) < tol_g
    assert rel_err(g2_i, og2_i) < tol_g
    assert rel_err(g2_o, og2_o) < tol_g


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_known_answer_arange27(dtype):
    image = np.arange(27).reshape(1, 1, 3, 3, 3)
    optical = np.array([0.1, 0.1, 0.1]).reshape(1, 1, 1, 1, 3)
    for pad in ("zeros", "border"):
        for ac in (True, False):
            _cmp_all_orders(image, optical, pad, ac, dtype)


def test_known_answer_oob():
    image = np.array([[1, 2], [3, 4], [5, 6], [7, 8]]).reshape(1, 1, 2, 2, 2)
    optical = np.array([0.1, 1.1, 0.1]).reshape(1, 1, 1, 1, 3)
    for pad in ("zeros", "border"):
        _cmp_all_orders(image, optical, pad, True, torch.float64)
        _cmp_all_orders(image, optical, pad, False, torch.float64)


def _random_input(rng, oob=False, max_dim=20):
    bs, c, d, h, w = [rng.randint(1, max_dim) for _ in range(5)]
    dg, hg, wg = [rng.randint(1, max_dim) for _ in range(3)]
    lo, hi = (-2, 2) if oob else (-1, 1)
    grid = rng.uniform(lo, hi, size=(bs, dg, hg, wg, 3))
    inp = rng.normal(size=(bs, c, d, h, w))
    return inp, grid


@pytest.mark.parametrize("pad,ac", [("zeros", False), ("zeros", True), ("border", True), ("border", False)])
def test_random_vs_oracle_fp64(pad, ac):
    rng = np.random.RandomState(0)
    for i in range(6):
        inp, grid = _random_input(rng, oob=True, max_dim=8)
        _cmp_all_orders(inp, grid, pad, ac, torch.float64, tol_f=1e-12, tol_g=1e-10)


def test_random_vs_oracle_fp32_channels_last_vec4():
    """The 128-bit path: channels_last_3d grid with C % 4 == 0 (MISO's C=4), zeros / align_corners=False."""
    rng = np.random.RandomState(1)
    for c in (4, 8, 16):
        inp = rng.normal(size=(1, c, 7, 5, 9))
        grid = rng.uniform(-1.2, 1.2, size=(1, 300, 1, 1, 3))
        _cmp_all_orders(inp, grid, "zeros", False, torch.float32, channels_last=True)
        _cmp_all_orders(inp, grid, "zeros", False, torch.float32, channels_last=False)


def _gradchecks(inp, grid, pad, ac):
    image = torch.tensor(inp, dtype=torch.float64, device="cuda", requires_grad=True)
    optical = torch.tensor(grid, dtype=torch.float64, device="cuda", requires_grad=True)
    fn = lambda a, b: cu().grid_sample_3d(a, b, padding_mode=pad, align_corners=ac)
    assert torch.autograd.gradcheck(fn, (image, optical), nondet_tol=1e-9)
    assert torch.autograd.gradgradcheck(fn, (image, optical), nondet_tol=1e-9)


def test_gradcheck_constant_oob_zeros():
    image = np.arange(27).reshape(1, 1, 3, 3, 3).astype(np.float64)
    optical = np.array([-2.1, 0.1, 0.1]).reshape(1, 1, 1, 1, 3)
    _gradchecks(image, optical, "zeros", True)


@pytest.mark.parametrize("pad,ac,oob", [("border", True, False), ("zeros", True, True), ("border", True, True),
                                        ("zeros", False, True)])
def test_gradcheck_random(pad, ac, oob):
    rng = np.random.RandomState(2)
    for i in range(4):
        inp, grid = _random_input(rng, oob=oob, max_dim=5)
        _gradchecks(inp, grid, pad, ac)


def test_grad_output_gradcheck():
    """test3d.py:52-67: gradcheck of the backward Function w.r.t. grad_output, input and grid."""
    from miso_b200.cuda_gridsample import _GridSample3dBackward
    rng = np.random.RandomState(3)
    for i in range(4):
        inp, grid = _random_input(rng, oob=True, max_dim=5)
        go = rng.normal(size=(inp.shape[0], inp.shape[1]) + grid.shape[1:4])
        a = torch.tensor(go, dtype=torch.float64, device="cuda", requires_grad=True)
        b = torch.tensor(inp, dtype=torch.float64, device="cuda", requires_grad=True)
        c = torch.tensor(grid, dtype=torch.float64, device="cuda", requires_grad=True)
        assert torch.autograd.gradcheck(lambda x, y, z: _GridSample3dBackward.apply(x, y, z, 0, True), (a, b, c),
                                        nondet_tol=1e-9)


def test_strided_inputs():
    rng = np.random.RandomState(4)
    for i in range(8):
        inp, grid = _random_input(rng, max_dim=5)
        image = torch.tensor(inp, dtype=torch.float64, device="cuda", requires_grad=True)
        optical = torch.tensor(grid, dtype=torch.float64, device="cuda", requires_grad=True)
        s = [rng.randint(1, 3) for _ in range(5)]
        g = [rng.randint(1, 3) for _ in range(3)]
        im = image[::s[0], ::s[1], ::s[2], ::s[3], ::s[4]]
        op = optical[::s[0], ::g[0], ::g[1], ::g[2]]
        fn = lambda a, b: cu().grid_sample_3d(a, b)
        assert torch.autograd.gradcheck(fn, (im, op), nondet_tol=1e-9)
        assert torch.autograd.gradgradcheck(fn, (im, op), nondet_tol=1e-9)


def test_use_case_conv_relu_sample():
    """test3d.py:121-151."""
    rng = np.random.RandomState(5)
    torch.manual_seed(0)
    inp, grid = _random_input(rng, max_dim=6)
    c = inp.shape[1]
    l1 = torch.nn.Conv3d(c, c, 1).double().cuda()
    l2 = torch.nn.Conv3d(c, 1, 1).double().cuda()
    image = torch.tensor(inp, dtype=torch.float64, device="cuda", requires_grad=True)
    optical = torch.tensor(grid, dtype=torch.float64, device="cuda", requires_grad=True)

    def fn(image, optical):
        out = torch.nn.functional.relu(l1(image))
        out = cu().grid_sample_3d(out, optical)
        out = l2(out)
        return (out * out).sum()

    assert torch.autograd.gradcheck(fn, (image, optical), nondet_tol=1e-5)
    assert torch.autograd.gradgradcheck(fn, (image, optical), nondet_tol=1e-5)


def test_empty_and_errors():
    image = torch.randn(1, 4, 3, 3, 3, device="cuda")
    empty = torch.zeros(1, 0, 1, 1, 3, device="cuda")
    out = cu().grid_sample_3d(image, empty, padding_mode="zeros", align_corners=False)
    assert out.shape == (1, 4, 0, 1, 1)
    with pytest.raises(AssertionError):
        cu().grid_sample_3d(image, empty, padding_mode="reflection")
    with pytest.raises(RuntimeError):
        cu().grid_sample_3d(image.cpu(), empty.cpu())


def test_far_out_of_bounds_is_zero():
    """Points beyond the bound by more than half a voxel give exactly 0 features and 0 gradients
    (SURVEY.md section 9)."""
    image = torch.randn(1, 4, 5, 6, 7, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    image.requires_grad_(True)
    g = torch.tensor([[3.0, 0.0, 0.0], [0.0, -1e9, 0.0], [0.0, 0.0, 1.5]], device="cuda").reshape(1, 3, 1, 1, 3)
    g.requires_grad_(True)
    out = cu().grid_sample_3d(image, g, padding_mode="zeros", align_corners=False)
    assert torch.count_nonzero(out) == 0
    gi, gg = grad(out.sum(), [image, g])
    assert torch.count_nonzero(gi) == 0 and torch.count_nonzero(gg) == 0


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
@pytest.mark.parametrize("padding_mode,align_corners", [("zeros", False), ("zeros", True), ("border", False)])
def test_double_backward_matches_reference_extension(dtype, tol, padding_mode, align_corners):
    """Value parity against the reference's OWN double-backward kernel (grid_sampler_3d_grad2_kernel,
    gridsample_cuda.cu:212-533), compiled unmodified into oracle/_ref/gridsample_grad2.so by oracle/build_ref.py and
    driven through the reference's plugin structure (oracle/ref_gpu.py): first-order grads, the double-backward's three
    outputs (gg_output via a second cotangent, g_input, g_grid) on the same inputs."""
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref/gridsample_grad2.so not built (python oracle/build_ref.py in the build container)")
    from miso_b200 import cuda_gridsample as cu
    g = torch.Generator().manual_seed(3)
    inp0 = torch.randn(2, 4, 5, 6, 7, generator=g, dtype=dtype)
    grid0 = (torch.rand(2, 300, 1, 1, 3, generator=g, dtype=dtype) * 2.4 - 1.2)
    go = torch.randn(2, 4, 300, 1, 1, generator=g, dtype=dtype).cuda()
    g2 = torch.randn(2, 300, 1, 1, 3, generator=g, dtype=dtype).cuda()
    g2i = torch.randn(2, 4, 5, 6, 7, generator=g, dtype=dtype).cuda()
    res = []
    for fn in (cu.grid_sample_3d, ref_gpu.grid_sample_3d):
        inp = inp0.clone().cuda().requires_grad_(True)
        grid = grid0.clone().cuda().requires_grad_(True)
        gor = go.clone().requires_grad_(True)
        out = fn(inp, grid, padding_mode=padding_mode, align_corners=align_corners)
        gi, gg = torch.autograd.grad(out, (inp, grid), gor, create_graph=True)
        ((gg * g2).sum() + (gi * g2i).sum()).backward()
        res.append((out.detach(), gi.detach(), gg.detach(), inp.grad, grid.grad, gor.grad))
    names = ["out", "grad_input", "grad_grid", "dbl.g_input", "dbl.g_grid", "dbl.gg_output"]
    for n, a, b in zip(names, res[0], res[1]):
        assert rel_err(a, b) < tol, (n, rel_err(a, b))
