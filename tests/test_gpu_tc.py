"""Tensor-core (tcgen05) path on the GPU: operand-layout self test, then parity vs reference golden."""
import ctypes as C

import numpy as np
import pytest
import torch

from alignsdf_b200 import _lib, engine, mesh as amesh, packer, tc_pack
from oracle import alignsdf_oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_umma_layout_selftest():
    """D = A.B^T (M=128 over a CTA pair, N=256, K=64) through the production operand layouts."""
    rng = np.random.default_rng(0)
    a = rng.standard_normal((128, 64)).astype(np.float16)
    b = rng.standard_normal((256, 64)).astype(np.float16)
    tiles = np.stack([tc_pack.swizzle_tile(b[:128]), tc_pack.swizzle_tile(b[128:])])
    dev = torch.device("cuda")
    a_d = torch.from_numpy(a).to(dev)
    b_d = torch.from_numpy(tiles).to(dev)
    d_d = torch.full((128, 256), float("nan"), dtype=torch.float32, device=dev)
    rc = _lib.lib().asdf_tc_selftest(_lib.ptr(a_d), _lib.ptr(b_d), _lib.ptr(d_d), _lib.stream_ptr(dev))
    _lib.check(rc, "asdf_tc_selftest")
    torch.cuda.synchronize()
    want = a.astype(np.float64) @ b.astype(np.float64).T
    got = d_d.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() < 1e-3, np.abs(got - want).max()


@pytest.mark.parametrize("name", ["sep_both9_n24", "sep_nerf3_n16", "sep_hand51_n16", "sep_hand6_n12",
                                  "sep_obj6_n12", "sep_both54_n12", "sep_both9_n32_handonly",
                                  "sep_both9_n20_objonly"])
def test_tc_two_pass_fields_match_reference_golden(name):
    meta, g, dec, sample = helpers.load_case(name)
    s = helpers.to_cuda(sample)
    hb, ob = meta.get("hand_branch", True), meta.get("obj_branch", True)
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, meta["N"], hb, ob, path="tc")
    assert np.float32(float(vols["voxel"])) == g["new_voxel"]
    assert np.array_equal(vols["origin"].numpy(), g["new_origin"])
    for key, vol in (("pass1_hand", vols["pass1_hand"]), ("pass1_obj", vols["pass1_obj"]),
                     ("pass2_hand", vols["hand"]), ("pass2_obj", vols["obj"])):
        if key in g:
            err = np.abs(vol.cpu().numpy() - g[key]).max()
            assert err <= TOL, (key, err)


def test_tc_matches_generic_fp32_kernel_and_is_deterministic():
    meta, g, dec, sample = helpers.load_case("sep_both9_n24")
    s = helpers.to_cuda(sample)
    bound = engine.get_engine(dec, torch.device("cuda")).bind(s.latent, s.specs, s.mano_results, s.obj_results)
    N = 40                                    # 64000 points: 500 tiles, > 3 waves of 74 CTA pairs
    ht, ot, _, bt = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="tc")
    hs, os_, _, bs = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="simt")
    assert (ht - hs).abs().max() <= 2e-6 and (ot - os_).abs().max() <= 2e-6
    ht2, ot2, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="tc")
    assert torch.equal(ht, ht2) and torch.equal(ot, ot2)
    # ragged ranges: not a multiple of the 128-point tile, odd begin
    h3, o3, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], begin=77, end=77 + 1001, path="tc")
    assert torch.equal(h3, ht[77:77 + 1001]) and torch.equal(o3, ot[77:77 + 1001])
    # explicit points
    xyz = (torch.rand(777, 3, generator=torch.Generator().manual_seed(2)) * 2 - 1).cuda()
    hp, op, _ = bound.eval_points(xyz, path="tc")
    hq, oq, _ = bound.eval_points(xyz, path="simt")
    assert (hp - hq).abs().max() <= 2e-6 and (op - oq).abs().max() <= 2e-6


V2_CASES = ["sep_both9_n24", "sep_nerf3_n16", "sep_hand51_n16", "sep_hand6_n12", "sep_obj6_n12",
            "sep_both54_n12", "sep_both9_n32_handonly", "sep_both9_n20_objonly"]


@pytest.mark.parametrize("name", V2_CASES)
def test_tc2_two_pass_fields_match_reference_golden(name):
    meta, g, dec, sample = helpers.load_case(name)
    s = helpers.to_cuda(sample)
    hb, ob = meta.get("hand_branch", True), meta.get("obj_branch", True)
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, meta["N"], hb, ob, path="tc2")
    assert np.float32(float(vols["voxel"])) == g["new_voxel"]
    assert np.array_equal(vols["origin"].numpy(), g["new_origin"])
    for key, vol in (("pass1_hand", vols["pass1_hand"]), ("pass1_obj", vols["pass1_obj"]),
                     ("pass2_hand", vols["hand"]), ("pass2_obj", vols["obj"])):
        if key in g:
            err = np.abs(vol.cpu().numpy() - g[key]).max()
            assert err <= TOL, (key, err)


def test_tc2_matches_generic_fp32_kernel_and_is_deterministic():
    meta, g, dec, sample = helpers.load_case("sep_both9_n24")
    s = helpers.to_cuda(sample)
    bound = engine.get_engine(dec, torch.device("cuda")).bind(s.latent, s.specs, s.mano_results, s.obj_results)
    N = 48                                    # 110592 points: 432 tiles of 256, > 5 waves of 74 CTA pairs
    ht, ot, _, bt = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="tc2")
    hs, os_, _, bs = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="simt")
    assert (ht - hs).abs().max() <= 3e-6 and (ot - os_).abs().max() <= 3e-6
    ht2, ot2, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="tc2")
    assert torch.equal(ht, ht2) and torch.equal(ot, ot2)
    h3, o3, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], begin=77, end=77 + 1001, path="tc2")
    assert torch.equal(h3, ht[77:77 + 1001]) and torch.equal(o3, ot[77:77 + 1001])
    xyz = (torch.rand(777, 3, generator=torch.Generator().manual_seed(2)) * 2 - 1).cuda()
    hp, op, _ = bound.eval_points(xyz, path="tc2")
    hq, oq, _ = bound.eval_points(xyz, path="simt")
    assert (hp - hq).abs().max() <= 3e-6 and (op - oq).abs().max() <= 3e-6
    far = xyz * 7.0                            # outside the default point-operand range: re-bound automatically
    hp, op, _ = bound.eval_points(far, path="tc2")
    hq, oq, _ = bound.eval_points(far, path="simt")
    assert (hp - hq).abs().max() <= 1e-5 and (op - oq).abs().max() <= 1e-5


@pytest.mark.parametrize("name", V2_CASES)
def test_tc3_two_pass_fields_match_reference_golden(name):
    """k1_tc3.cu (fp16 main product + e4m3 correction products) vs the real reference: <= 1e-5, and no
    silent fallback to the all-fp16 kernel."""
    meta, g, dec, sample = helpers.load_case(name)
    s = helpers.to_cuda(sample)
    hb, ob = meta.get("hand_branch", True), meta.get("obj_branch", True)
    before = engine.FALLBACKS["tc3_to_tc2"]
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, meta["N"], hb, ob, path="tc3")
    assert engine.FALLBACKS["tc3_to_tc2"] == before
    assert np.float32(float(vols["voxel"])) == g["new_voxel"]
    assert np.array_equal(vols["origin"].numpy(), g["new_origin"])
    for key, vol in (("pass1_hand", vols["pass1_hand"]), ("pass1_obj", vols["pass1_obj"]),
                     ("pass2_hand", vols["hand"]), ("pass2_obj", vols["obj"])):
        if key in g:
            err = np.abs(vol.cpu().numpy() - g[key]).max()
            assert err <= TOL, (key, err)
            assert err <= 6e-6, (key, err)          # the margin the design relies on (emulated: 2.3e-6)


def test_tc3_matches_generic_fp32_kernel_emulator_and_is_deterministic():
    from alignsdf_b200 import tc3_pack
    from tests.tc3_emulate import emulate
    meta, g, dec, sample = helpers.load_case("sep_both9_n24")
    s = helpers.to_cuda(sample)
    eng = engine.get_engine(dec, torch.device("cuda"))
    bound = eng.bind(s.latent, s.specs, s.mano_results, s.obj_results)
    N = 48
    before = engine.FALLBACKS["tc3_to_tc2"]
    ht, ot, _, bt = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="tc3")
    hs, os_, _, bs = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="simt")
    assert (ht - hs).abs().max() <= 6e-6 and (ot - os_).abs().max() <= 6e-6
    ht2, ot2, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="tc3")
    assert torch.equal(ht, ht2) and torch.equal(ot, ot2)
    h3, o3, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], begin=77, end=77 + 1001, path="tc3")
    assert torch.equal(h3, ht[77:77 + 1001]) and torch.equal(o3, ot[77:77 + 1001])
    xyz = (torch.rand(777, 3, generator=torch.Generator().manual_seed(2)) * 2 - 1).cuda()
    hp, op, _ = bound.eval_points(xyz, path="tc3")
    hq, oq, _ = bound.eval_points(xyz, path="simt")
    assert (hp - hq).abs().max() <= 6e-6 and (op - oq).abs().max() <= 6e-6
    assert engine.FALLBACKS["tc3_to_tc2"] == before
    # the CPU emulation of the packed bytes predicts the kernel to fp32-accumulation-order noise
    tc3 = bound._tc3_for(2.0)
    he, oe = emulate(eng.tc3_static.cpu().numpy(), tc3.sample.cpu().numpy(), xyz.cpu().numpy())
    assert np.abs(he - hp.cpu().numpy()).max() <= 1e-6 and np.abs(oe - op.cpu().numpy()).max() <= 1e-6


def test_tc3_falls_back_to_fp16_corrections_when_activations_leave_the_fp8_range():
    """Scaling layer 0 up / layer 1 down by 2^9 keeps the function but pushes x1 beyond 448: the
    kernel must raise its status flag and the engine must re-run the query through k1_tc2.cu."""
    from alignsdf_b200 import synthetic
    dec = synthetic.make_decoder(3)
    with torch.no_grad():
        for p in ("linh", "lino"):
            l0, l1 = getattr(dec, p + "0"), getattr(dec, p + "1")
            l0.weight_g.mul_(512.0); l0.bias.mul_(512.0)
            l1.weight_g.mul_(1.0 / 512.0)
    s = synthetic.make_sample(3).to(torch.device("cuda"))
    bound = engine.get_engine(dec, torch.device("cuda")).bind(s.latent, s.specs, s.mano_results, s.obj_results)
    N = 24
    before = engine.FALLBACKS["tc3_to_tc2"]
    ht, ot, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="tc3")
    assert engine.FALLBACKS["tc3_to_tc2"] == before + 1
    hs, os_, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="simt")
    assert (ht - hs).abs().max() <= 1e-5 and (ot - os_).abs().max() <= 1e-5


def test_tc3_full_256_grid_within_contract_of_the_fp32_kernel():
    """BASELINE's full size: all 16.7 M points of the 256^3 pass-1 grid, product kernel (fp16 + fp8 corrections)
    vs the exact-fp32 generic kernel (itself pinned to the reference's golden fields): <= 1e-5 everywhere, same
    bounding box, no range fallback."""
    from alignsdf_b200 import synthetic
    dec = synthetic.make_decoder(0)
    s = synthetic.make_sample(0).to(torch.device("cuda"))
    bound = engine.get_engine(dec, torch.device("cuda")).bind(s.latent, s.specs, s.mano_results, s.obj_results)
    N = 256
    before = engine.FALLBACKS["tc3_to_tc2"]
    ht, ot, _, bt = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="tc3")
    assert engine.FALLBACKS["tc3_to_tc2"] == before
    hs, os_, _, bs = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="simt")
    eh, eo = float((ht - hs).abs().max()), float((ot - os_).abs().max())
    assert eh <= TOL and eo <= TOL, (eh, eo)
    assert eh <= 7e-6 and eo <= 7e-6, (eh, eo)        # observed ~3e-6 at this size
    assert torch.equal(bt, bs)
