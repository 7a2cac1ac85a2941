"""PixelAlign (specs['PixelAlign'], utils/utils.py:536-566) on the GPU: per-point latents sampled in the kernel from
projected feature maps.  The grid passes are covered by test_gpu_parity.py::test_two_pass_fields_match_reference_golden
(fixtures sep_pa_both9_n12 / comb_pa_xyz3_n10 captured from the real reference); here: the arbitrary-point API, the
out-of-image rule and the drop-in call."""
import os

import numpy as np
import pytest
import torch

from alignsdf_b200 import engine, mesh as amesh, utils as autils
from oracle import alignsdf_oracle as orc
from tests import helpers

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = torch.device("cuda")


@pytest.mark.parametrize("name", ["sep_pa_both9_n12", "comb_pa_xyz3_n10"])
def test_decode_sdf_multi_output_with_pixel_aligned_latents(name):
    """The reference's call (embedded queries in, utils/utils.py:561-572) incl. ragged sizes and points far outside
    the image (mean feature)."""
    meta, g, dec, sample = helpers.load_case(name)
    s = helpers.to_cuda(sample)
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    cfg = orc.decoder_cfg(dec)
    gen = torch.Generator().manual_seed(5)
    for P in (1, 33, 1000):
        xyz = torch.rand(P, 3, generator=gen) * 2.4 - 1.2
        xyz[::7] *= 4.0                                              # well outside the image
        feats = orc.embed(xyz, sample.specs, sample.mano_results, sample.obj_results)
        with torch.no_grad():
            rh, ro, _ = orc.decode_points(sd, cfg, sample.latent, xyz, sample.specs, sample.mano_results,
                                          sample.obj_results, sample.cam_intr)
        h, o, _ = autils.decode_sdf_multi_output(dec, s.latent, feats.to(DEV), s.mano_results, s.cam_intr, s.specs)
        assert h.shape == (P, 1) and o.shape == (P, 1)
        assert (h.cpu() - rh).abs().max() <= TOL and (o.cpu() - ro).abs().max() <= TOL
        # raw-xyz fast path (pose-align folded, projection folded into the point affine)
        bound = engine.get_engine(dec, DEV).bind(s.latent, s.specs, s.mano_results, s.obj_results, cam_intr=s.cam_intr)
        h2, o2, _ = bound.eval_points(xyz.to(DEV))
        assert (h2.cpu() - rh[:, 0]).abs().max() <= TOL and (o2.cpu() - ro[:, 0]).abs().max() <= TOL
        assert bound.kinds_used == {"simt"}                         # per-point latents never take the folded tensor-core path


def test_pixel_align_needs_camera_and_joints():
    meta, g, dec, sample = helpers.load_case("sep_pa_both9_n12")
    s = helpers.to_cuda(sample)
    bound = engine.get_engine(dec, DEV).bind(s.latent, s.specs, s.mano_results, s.obj_results)      # no cam_intr
    with pytest.raises(ValueError):
        bound.eval_points(torch.zeros(4, 3, device=DEV))
    with pytest.raises(ValueError):                                  # a plain latent vector is not a feature map
        engine.get_engine(dec, DEV).bind(s.latent[:, :, 0, 0], s.specs, s.mano_results, s.obj_results,
                                         cam_intr=s.cam_intr).eval_points(torch.zeros(4, 3, device=DEV))


def test_create_mesh_with_pixel_align(tmp_path):
    meta, g, dec, sample = helpers.load_case("sep_pa_both9_n12")
    s = helpers.to_cuda(sample)
    res = amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results,
                                             s.cam_intr, s.specs, str(tmp_path / "pa"), N=meta["N"])
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, meta["N"], cam_intr=s.cam_intr)
    assert np.abs(vols["hand"].cpu().numpy() - g["pass2_hand"]).max() <= TOL
    for tag in ("hand", "obj"):
        has_surface = bool((g["pass2_" + tag] < 0).any() and (g["pass2_" + tag] >= 0).any())
        assert (res[tag] is not None) == has_surface
        assert os.path.exists(tmp_path / f"pa_{tag}.ply") == has_surface


def test_pipelined_batch_api_with_pixel_align(tmp_path):
    meta, g, dec, sample = helpers.load_case("sep_pa_both9_n12")
    res = amesh.create_meshes_pipelined(dec, [sample, sample], [str(tmp_path / "a"), str(tmp_path / "b")], N=meta["N"],
                                        device=DEV)
    one = amesh.create_mesh_combined_decoder(True, True, False, dec, *[getattr(helpers.to_cuda(sample), k) for k in
                                             ("latent", "mano_results", "obj_results", "cam_intr", "specs")],
                                             str(tmp_path / "c"), N=meta["N"])
    for r in res:
        for tag in ("hand", "obj"):
            assert (r[tag] is None) == (one[tag] is None)
            if r[tag] is not None:
                assert np.array_equal(r[tag].vertices, one[tag].vertices) and np.array_equal(r[tag].faces, one[tag].faces)
