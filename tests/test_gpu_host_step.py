"""-m gpu: the end-to-end entry points dgpmp2_gn_step_host_f32 / _f64 (HOST buffers in, HOST buffers out; what
bench.py's `e2e` number and the planner's CPU-tensor path run through) against the live reference's goldens, with the
tolerances of tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden, rel_err, step_cases

pytestmark = pytest.mark.gpu

TOL = {torch.float64: dict(dth=1e-9, err=1e-11), torch.float32: dict(dth=1e-5, err=1e-6)}
STATIC = [n for n in step_cases() if bool(load_golden(n)['static'])]


def _host(a, dtype, pinned):
    t = torch.as_tensor(np.ascontiguousarray(a)).to(dtype).contiguous()
    return t.pin_memory() if pinned else t


def _stepper(g, dtype, shared_sdf=False):
    from dgpmp2_b200 import _lib, ops
    from tests.gpu_helpers import cparams
    B, T, d = g['th'].shape
    H, W = g['sdf'].shape[-2:]
    cp = cparams(T, B=B, H=H, W=W, x_lims=g['x_lims'], y_lims=g['y_lims'])
    _lib.set_sdf_shape(cp, H, W, 0 if shared_sdf else H * W)
    return ops.HostStepper(cp, dtype)


def _check(out, g, dtype, rows=slice(None)):
    dth, err, err_ext, status = out
    tol = TOL[dtype]
    assert int(status.abs().max()) == 0
    assert rel_err(dth[rows], g['dth'][rows]) < tol['dth']
    np.testing.assert_allclose(err.double().numpy()[rows], g['err'].reshape(-1)[rows], rtol=tol['err'])
    np.testing.assert_allclose(err_ext.double().numpy()[rows], g['err_ext'].reshape(-1)[rows], rtol=tol['err'])


@pytest.mark.parametrize('pinned', [True, False], ids=['pinned', 'pageable'])
@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('name', STATIC)
def test_host_step_vs_reference_golden(name, dtype, pinned):
    g = load_golden(name)
    B, T, d = g['th'].shape
    hs = _stepper(g, dtype)
    th, start, goal, sdf = (_host(g[k], dtype, pinned) for k in ('th', 'start', 'goal', 'sdf'))
    out = hs.step(th, start.reshape(B, d), goal.reshape(B, d), sdf.reshape(B, *g['sdf'].shape[-2:]))
    assert all(not t.is_cuda for t in out)
    _check(out, g, dtype)
    first = [t.clone() for t in out]
    # the SDF stays in the device workspace: a second call may skip the 4*B*H*W-byte copy and must give the same bits
    out2 = hs.step(th, start.reshape(B, d), goal.reshape(B, d), None, sdf_resident=True)
    for a, b in zip(first, out2):
        assert torch.equal(a, b)
    # ... also for a different trajectory against the same resident SDF: equal to the device-resident entry point
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams
    th2 = (th + 0.01).contiguous()
    out3 = [t.clone() for t in hs.step(th2, start.reshape(B, d), goal.reshape(B, d), None, sdf_resident=True)]
    cp = cparams(T, x_lims=g['x_lims'], y_lims=g['y_lims'])
    ref = ops.gn_step(cp, th2.cuda(), start.cuda(), goal.cuda(), sdf.cuda())
    for a, b in zip(out3, ref):
        assert torch.equal(a, b.cpu())


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_host_step_shared_sdf_stride_zero(dtype):
    """sdf_stride_b = 0: ONE SDF for the whole batch (H*W elements cross the bus instead of B*H*W)."""
    g = load_golden('step_static_B4_T64_k0')
    B, T, d = g['th'].shape
    hs = _stepper(g, dtype, shared_sdf=True)
    assert hs.sdf_bytes == g['sdf'].shape[-1] * g['sdf'].shape[-2] * (4 if dtype == torch.float32 else 8)
    th, start, goal = (_host(g[k], dtype, True) for k in ('th', 'start', 'goal'))
    sdf0 = _host(g['sdf'][0, 0], dtype, True)
    out = hs.step(th, start.reshape(B, d), goal.reshape(B, d), sdf0)
    _check(out, g, dtype, rows=slice(0, 1))          # problem 0 sees its own SDF -> the reference's numbers
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams
    cp = cparams(T, x_lims=g['x_lims'], y_lims=g['y_lims'])
    ref = ops.gn_step(cp, th.cuda(), start.cuda(), goal.cuda(), sdf0.cuda().reshape(1, 1, *sdf0.shape))
    for a, b in zip(out, ref):
        assert torch.equal(a, b.cpu())
    # ... and read in place from the pinned buffer (one shared field, every problem's taps over PCIe)
    out_z = hs.step(th, start.reshape(B, d), goal.reshape(B, d), sdf0, in_place=True)
    assert hs.last_sdf_read_in_place
    for a, b in zip(out_z, ref):
        assert torch.equal(a, b.cpu())


def test_host_step_validates_its_arguments():
    from dgpmp2_b200 import _lib
    g = load_golden('step_static_B2_T3')
    B, T, d = g['th'].shape
    hs = _stepper(g, torch.float32)
    th, start, goal, sdf = (_host(g[k], torch.float32, False) for k in ('th', 'start', 'goal', 'sdf'))
    start, goal = start.reshape(B, d), goal.reshape(B, d)
    with pytest.raises(_lib.Dgpmp2Error):
        hs.step(th, start, goal, None, sdf_resident=True)          # nothing staged yet
    with pytest.raises(TypeError):
        hs.step(th.double(), start, goal, sdf)
    with pytest.raises(ValueError):
        hs.step(th[:, :, :2], start, goal, sdf)                    # not contiguous / wrong size
    with pytest.raises(ValueError):
        hs.step(th, start, goal, sdf[:1])
    with pytest.raises(_lib.Dgpmp2Error):
        hs.step(th.cuda(), start, goal, sdf)
    hs.step(th, start, goal, sdf)


def test_bit_packed_maps_give_the_same_field_and_the_same_step():
    """dgpmp2_sdf_from_occupancy_bits_f32 == the float-image EDT (itself bit-identical to scipy's sdf_2d, test_gpu_api),
    and dgpmp2_gn_step_host_occ_f32 (maps cross the bus as bits, SDF built on the device) == the host-SDF entry point."""
    from dgpmp2_b200 import _lib, ops
    from dgpmp2_b200.datasets.synthetic import make_problems
    from tests.gpu_helpers import cparams
    B, T = 37, 64
    for size in (128, 100, 33):
        pr = make_problems(B, T, im_size=size, seed=size, unique_envs=B)
        im = pr['im'][:, 0].contiguous()
        res = 10.0 / size
        ref = ops.sdf_from_occupancy(im.cuda(), padlen=0, res=res)
        bits = ops.pack_occupancy_bits(im)
        got = ops.sdf_from_occupancy_bits(bits.cuda(), size, res=res)
        assert torch.equal(got, ref)
        cp = cparams(T, B=B, H=size, W=size)
        _lib.set_sdf_shape(cp, size, size, size * size)
        hs, ho = ops.HostStepper(cp, torch.float32), ops.HostOccStepper(cp)
        th, start, goal = pr['th_init'].contiguous(), pr['start'].reshape(B, 4).contiguous(), pr['goal'].reshape(B, 4).contiguous()
        a = [t.clone() for t in hs.step(th, start, goal, ref.cpu().contiguous())]
        b = ho.step(th, start, goal, bits)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        assert ho.occ_bytes == 4 * B * size * ((size + 31) // 32)
    np_sdf = pr['sdf'][:, 0].numpy()
    np.testing.assert_allclose(got.cpu().numpy(), np_sdf, rtol=1e-6, atol=1e-6)       # the dataset's scipy field


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_chunked_host_step_is_bit_identical(dtype):
    """B >= 256 with per-problem SDFs: the host entry point pipelines 4 chunks over two helper streams; same bits as the
    single-launch path (DGPMP2_HOST_CHUNKS=1) and as the device-resident entry point, pinned or pageable buffers."""
    import os
    from dgpmp2_b200 import _lib, ops
    from dgpmp2_b200.datasets.synthetic import make_problems
    from tests.gpu_helpers import cparams
    B, T = 301, 64
    pr = make_problems(B, T, im_size=64, seed=9, unique_envs=16, dtype=dtype)
    cp = cparams(T, B=B, H=64, W=64)
    _lib.set_sdf_shape(cp, 64, 64, 64 * 64)
    hs = ops.HostStepper(cp, dtype)
    th, start, goal, sdf = pr['th_init'].contiguous(), pr['start'].reshape(B, 4).contiguous(), pr['goal'].reshape(B, 4).contiguous(), pr['sdf'][:, 0].contiguous()
    ref = [t.cpu() for t in ops.gn_step(cparams(T), th.cuda(), start.cuda(), goal.cuda(), sdf.cuda())]
    saved = os.environ.pop('DGPMP2_HOST_CHUNKS', None)
    try:
        for pinned in (False, True):
            args = [t.pin_memory() if pinned else t for t in (th, start, goal, sdf)]
            for chunks in (None, '1', '7'):
                os.environ.pop('DGPMP2_HOST_CHUNKS', None)
                if chunks:
                    os.environ['DGPMP2_HOST_CHUNKS'] = chunks
                out = hs.step(*args)
                for a, b in zip(out, ref):
                    assert torch.equal(a, b), (pinned, chunks)
    finally:
        os.environ.pop('DGPMP2_HOST_CHUNKS', None)
        if saved is not None:
            os.environ['DGPMP2_HOST_CHUNKS'] = saved


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_sdf_read_in_place_from_pinned_host_memory(dtype):
    """DGPMP2_SDF_IN_PLACE: a pinned SDF is not copied -- the kernel gathers its taps from the host buffer over PCIe.
    Same bits as the copying call and as the reference golden's tolerance; nothing is staged (a following
    sdf_resident call must be refused); a pageable SDF silently takes the copying path and is staged."""
    from dgpmp2_b200 import _lib
    name = STATIC[0]
    g = load_golden(name)
    B, T, d = g['th'].shape
    hw = g['sdf'].shape[-2:]
    th, start, goal, sdf = (_host(g[k], dtype, True) for k in ('th', 'start', 'goal', 'sdf'))
    args = (th, start.reshape(B, d), goal.reshape(B, d), sdf.reshape(B, *hw))
    lib = _lib.load()
    assert lib.dgpmp2_host_pointer_is_mapped(args[3].data_ptr()) == 1
    assert lib.dgpmp2_host_pointer_is_mapped(args[3].clone().data_ptr()) == 0          # pageable copy
    hs = _stepper(g, dtype)
    out = [t.clone() for t in hs.step(*args, in_place=True)]
    assert hs.last_sdf_read_in_place
    _check(out, g, dtype)
    with pytest.raises(_lib.Dgpmp2Error):
        hs.step(args[0], args[1], args[2], None, sdf_resident=True)                  # nothing was staged
    ref = [t.clone() for t in hs.step(*args)]                                         # copying call
    for a, b in zip(out, ref):
        assert torch.equal(a, b)
    # the host buffer is read at call time: change it in place, the next in-place call sees the new field
    sdf2 = (args[3] * 0.5).contiguous().pin_memory()
    want = [t.clone() for t in hs.step(args[0], args[1], args[2], sdf2)]
    args[3].copy_(sdf2)
    got = hs.step(*args, in_place=True)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    # pageable: falls back to the copy and stages it
    hs2 = _stepper(g, dtype)
    pageable = args[3].clone()
    out_p = [t.clone() for t in hs2.step(args[0], args[1], args[2], pageable, in_place=True)]
    assert not hs2.last_sdf_read_in_place
    for a, b in zip(out_p, want):
        assert torch.equal(a, b)
    out_r = hs2.step(args[0], args[1], args[2], None, sdf_resident=True)
    for a, b in zip(out_r, want):
        assert torch.equal(a, b)
