"""GPU parity tests: the CUDA path (through the C ABI) vs the oracle and vs the reference's golden
vectors.  Tolerance for SDF values: 1e-5 absolute (BASELINE.json north_star); grids, re-grid
parameters, class labels, bounding boxes, marching-cubes output: bit-exact."""
import os

import numpy as np
import pytest
import torch

from alignsdf_b200 import engine, mesh as amesh, synthetic, utils as autils
from alignsdf_b200.deep_sdf import mesh as legacy_mesh, utils as legacy_utils
from oracle import alignsdf_oracle as orc
from oracle import mc_oracle as mo
from tests import helpers

pytestmark = pytest.mark.gpu
TOL = 1e-5
PATHS = ["simt", "auto"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda")


def test_library_sees_a_blackwell_device(dev):
    from alignsdf_b200 import _lib
    assert _lib.lib().asdf_device_ok() == 1, _lib.lib().asdf_last_error()


def test_grid_coordinates_bit_exact(dev):
    meta, g, _, _ = helpers.load_case("sep_both9_n24")
    got = engine.grid_points(24, float(g["new_voxel"]), g["new_origin"].tolist()).cpu().numpy()
    assert np.array_equal(got, g["xyz2"])                       # captured from the real reference
    w = np.load(os.path.join(helpers.GOLD, "grid512_windows.npz"))
    for key in w.files:                                           # float(i) rounding above 2**24
        _, a, b = key.split("_")
        got = engine.grid_points(512, 2.0 / 511, [-1, -1, -1], "reference", int(a), int(b)).cpu().numpy()
        assert np.array_equal(got, w[key]), key
    reg = engine.grid_points(16, 2.0 / 15, [-1, -1, -1], "regular").cpu().numpy()
    assert np.array_equal(reg, orc.grid_points(16, 2.0 / 15, [-1, -1, -1], "regular").numpy())


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", helpers.field_cases())
def test_two_pass_fields_match_reference_golden(dev, name, path):
    meta, g, dec, sample = helpers.load_case(name)
    s = helpers.to_cuda(sample)
    hb, ob = meta.get("hand_branch", True), meta.get("obj_branch", True)
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, meta["N"], hb, ob,
                             cls_branch="pass2_cls" in g, path=path, cam_intr=s.cam_intr)
    assert np.float32(float(vols["voxel"])) == g["new_voxel"]
    assert np.array_equal(vols["origin"].numpy(), g["new_origin"])
    for key, vol in (("pass1_hand", vols["pass1_hand"]), ("pass1_obj", vols["pass1_obj"]),
                     ("pass2_hand", vols["hand"]), ("pass2_obj", vols["obj"])):
        if key in g:
            err = np.abs(vol.cpu().numpy() - g[key]).max()
            assert err <= TOL, (key, err)
    if "pass2_cls" in g:
        assert np.array_equal(vols["cls"].cpu().numpy().reshape(-1), g["pass2_cls"])


def test_slab_ranges_are_bit_identical_to_full_grid(dev):
    """z-slab sharding property: evaluating [begin,end) pieces == evaluating the whole grid."""
    meta, g, dec, sample = helpers.load_case("sep_both9_n24")
    s = helpers.to_cuda(sample)
    bound = engine.get_engine(dec, dev).bind(s.latent, s.specs, s.mano_results, s.obj_results)
    N = 24
    full_h, full_o, _, box = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3)
    cuts = [0, 5 * N * N, 5 * N * N + 77, 17 * N * N, N ** 3]
    hs, os_, boxes = [], [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        h, o, _, bx = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], begin=a, end=b, bbox_mask=3)
        hs.append(h); os_.append(o); boxes.append(bx)
    assert torch.equal(torch.cat(hs), full_h) and torch.equal(torch.cat(os_), full_o)
    bs = torch.stack(boxes).cpu()
    merged = torch.cat([bs[:, 0:3].min(0).values, bs[:, 3:6].max(0).values,
                        bs[:, 6:9].min(0).values, bs[:, 9:12].max(0).values])
    assert torch.equal(merged, box.cpu())
    # and the fused bbox equals nonzero(sdf<0) of the oracle-style reduction
    idx = torch.nonzero(full_h.view(N, N, N) < 0)
    assert box[0:3].tolist() == idx.min(0).values.tolist() and box[3:6].tolist() == idx.max(0).values.tolist()


def test_empty_negative_set_follows_reference_convention(dev):
    """No sdf<0 anywhere -> min=max=(0,0,0) (utils/mesh.py:209-211) -> cube of 4 voxels at -1-2vs."""
    dec = synthetic.make_decoder(0)
    with torch.no_grad():
        dec.linh4.bias += 5.0
        dec.lino4.bias += 5.0
    s = helpers.to_cuda(synthetic.make_sample(0))
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, 12)
    sd = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
    res = orc.two_pass_field(sd, orc.decoder_cfg(dec), sample_cpu(s).latent, s.specs,
                             sample_cpu(s).mano_results, sample_cpu(s).obj_results, 12)
    assert float(vols["voxel"]) == float(res["voxel"])
    assert np.array_equal(vols["origin"].numpy(), res["origin"].numpy())


def sample_cpu(s):
    return s.to(torch.device("cpu"))


@pytest.mark.parametrize("path", PATHS)
def test_points_and_feature_queries_match_oracle(dev, path):
    meta, g, dec, sample = helpers.load_case("sep_both9_n24")
    s = helpers.to_cuda(sample)
    gen = torch.Generator().manual_seed(3)
    xyz = torch.rand(1000, 3, generator=gen) * 2 - 1           # ragged size (not a tile multiple)
    sd = dec.state_dict()
    with torch.no_grad():
        h, o, _ = orc.decode_points(sd, orc.decoder_cfg(dec), sample.latent, xyz, sample.specs,
                                    sample.mano_results, sample.obj_results)
    gh, go, _ = autils.decode_sdf_points(dec, s.latent, xyz.to(dev), s.mano_results, s.obj_results, s.specs)
    assert (gh.cpu() - h[:, 0]).abs().max() <= TOL and (go.cpu() - o[:, 0]).abs().max() <= TOL
    if path == "simt":
        feats = autils.kinematic_embedding(xyz.to(dev), s.mano_results, xyz.shape[0], 9,
                                           s.specs["SdfScaleFactor"], s.obj_results, "both")
        ref_feats = orc.embed(xyz, sample.specs, sample.mano_results, sample.obj_results)
        assert (feats.cpu() - ref_feats).abs().max() <= 5e-6
        fh, fo, third = autils.decode_sdf_multi_output(dec, s.latent, feats, s.mano_results, None, s.specs)
        assert fh.shape == (1000, 1)
        assert third.shape == (1,) and float(third[0]) == 0.0                # networks/model.py:350 Tensor([0])
        assert (fh[:, 0].cpu() - h[:, 0]).abs().max() <= TOL and (fo[:, 0].cpu() - o[:, 0]).abs().max() <= TOL
    # empty query
    eh, eo, _ = autils.decode_sdf_points(dec, s.latent, xyz[:0].to(dev), s.mano_results, s.obj_results, s.specs)
    assert eh.numel() == 0 and eo.numel() == 0


def test_decode_sdf_multi_output_returns_the_class_logits(dev):
    """the third output of a decoder with a classifier head is the raw logits [P, num_class]
    (networks/model.py:161-162,188) -- reference-style callers take .argmax(dim=1) of it (utils/mesh.py:60,112,157)."""
    meta, g, dec, sample = helpers.load_case("comb_cls_n12")
    s = helpers.to_cuda(sample)
    xyz = torch.rand(700, 3, generator=torch.Generator().manual_seed(4)) * 2 - 1
    feats = autils.kinematic_embedding(xyz.to(dev), s.mano_results, 700, 9, s.specs["SdfScaleFactor"], s.obj_results, "both")
    h, o, logits = autils.decode_sdf_multi_output(dec, s.latent, feats, s.mano_results, None, s.specs)
    assert h.shape == (700, 1) and o.shape == (700, 1) and logits.shape == (700, 6) and logits.dtype == torch.float32
    with torch.no_grad():
        rh, ro, rl = orc.decode_points(dec.state_dict(), orc.decoder_cfg(dec), sample.latent, xyz, sample.specs,
                                       sample.mano_results, sample.obj_results)
    assert (h.cpu() - rh).abs().max() <= TOL and (o.cpu() - ro).abs().max() <= TOL
    assert (logits.cpu() - rl).abs().max() <= 1e-4 * max(1.0, float(rl.abs().max()))
    top2 = rl.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(logits.argmax(dim=1).cpu()[clear], rl.argmax(dim=1)[clear])


def test_torch_fp32_reference_forward_on_gpu(dev):
    """Plain eager-PyTorch fp32 forward of the same module on the GPU vs the CUDA path."""
    meta, g, dec, sample = helpers.load_case("sep_both9_n24")
    s = helpers.to_cuda(sample)
    xyz = (torch.rand(4096, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(dev)
    feats = autils.kinematic_embedding(xyz, s.mano_results, 4096, 9, s.specs["SdfScaleFactor"], s.obj_results, "both")
    dgpu = synthetic.make_decoder(meta["seed"]).to(dev)
    with torch.no_grad():
        rh, ro, _ = dgpu(torch.cat([s.latent.expand(4096, -1), feats], 1))
    gh, go, _ = autils.decode_sdf_points(dec, s.latent, xyz, s.mano_results, s.obj_results, s.specs)
    assert (gh - rh[:, 0]).abs().max() <= TOL and (go - ro[:, 0]).abs().max() <= TOL


def _mc_equal(vol, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)):
    v, f, k = mo.marching_cubes(vol, 0.0, spacing)
    out = engine.marching_cubes(torch.from_numpy(np.ascontiguousarray(vol)).cuda(), 0.0, spacing, origin,
                                want_keys=True)
    assert np.array_equal(out["keys"].cpu().numpy().astype(np.uint64), k)
    assert np.array_equal(out["faces"].cpu().numpy(), f)
    gv = out["verts"].cpu().numpy()
    assert np.array_equal(gv.view(np.uint32), v.view(np.uint32)), np.abs(gv - v).max()
    pts = out["points"].cpu().numpy()
    assert np.array_equal(pts, (np.asarray(origin, np.float32)[None] + v).astype(np.float32))
    return v, f


def test_marching_cubes_matches_oracle_bit_exact(dev):
    rng = np.random.default_rng(0)
    _mc_equal(rng.standard_normal((9, 10, 11)).astype(np.float32))            # every case + variant
    _mc_equal(rng.standard_normal((33, 17, 40)).astype(np.float32), (0.5, 0.25, 2.0), (1.0, -2.0, 0.5))
    ax = np.linspace(-1, 1, 48, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    v, f = _mc_equal(np.sqrt(x * x + y * y + z * z) - np.float32(0.6), [2 / 47] * 3, [-1, -1, -1])
    inv = mo.mesh_invariants(v, f)
    assert inv["closed"] and inv["oriented"] and inv["euler"] == 2
    tor = np.sqrt((np.sqrt(x * x + y * y) - 0.55) ** 2 + z * z) - np.float32(0.2)
    v, f = _mc_equal(tor)
    assert mo.mesh_invariants(v, f)["euler"] == 0
    vol = np.ones((5, 5, 5), np.float32); vol[2, 2, 2] = 0.0                   # value == iso exactly
    _mc_equal(vol - 0.0)


def test_marching_cubes_level_outside_range(dev):
    with pytest.raises(ValueError, match="Surface level must be within volume data range"):
        engine.marching_cubes(torch.ones(6, 6, 6, device=dev), 0.0)


def test_marching_cubes_slabs_stitch_to_the_single_gpu_mesh(dev):
    meta, g, _, _ = helpers.load_case("sep_both9_n24")
    vol = g["pass2_hand"]
    full = engine.marching_cubes(torch.from_numpy(vol).cuda(), 0.0, [0.1] * 3, want_keys=True)
    parts = []
    for a, b in ((0, 9), (8, 17), (16, 24)):        # slabs with one halo plane each
        parts.append(engine.marching_cubes(torch.from_numpy(np.ascontiguousarray(vol[a:b])).cuda(), 0.0,
                                           [0.1] * 3, index0_offset=a, want_keys=True, check_range=False))
    keys = torch.cat([p["keys"] for p in parts]).cpu().numpy()
    verts = torch.cat([p["verts"] for p in parts]).cpu().numpy()
    uk, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    off, faces = 0, []
    for p in parts:
        faces.append(inv[p["faces"].cpu().numpy().astype(np.int64) + off]); off += p["keys"].numel()
    assert np.array_equal(uk, full["keys"].cpu().numpy())
    assert np.array_equal(verts[first], full["verts"].cpu().numpy())
    assert np.array_equal(np.concatenate(faces), full["faces"].cpu().numpy())


def test_marching_cubes_512_windows_match_the_oracle(dev):
    """BASELINE config #5 (512^3, the marching-cubes stress): the mesh of the FULL volume -- interior-tile path of
    mc_classify, 32-bit offsets in the millions -- restricted to 40^3 windows equals the oracle's mesh of each window:
    same vertices (keys, bit-equal positions), same triangles in the same order."""
    N, W = 512, 40
    dec = synthetic.make_decoder(0, init="default")
    s = synthetic.make_sample(0).to(dev)
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, N, keep_pass1=False)
    vs = float(vols["voxel"])
    vol = vols["hand"]
    full = engine.marching_cubes(vol, 0.0, [vs] * 3, want_keys=True)
    keys = full["keys"].cpu().numpy().astype(np.uint64)
    verts = full["verts"].cpu().numpy()
    tri_keys = keys[full["faces"].cpu().numpy()]
    assert np.all(np.diff(keys.astype(np.int64)) > 0)                      # vertex order = key order
    host = vol.cpu().numpy()
    rng = np.random.default_rng(3)
    for pick in rng.choice(len(keys), 3, replace=False):
        p = int(keys[pick] // 4)
        c = np.array([p // (N * N), (p // N) % N, p % N])
        lo = np.clip(c - W // 2, 0, N - W - 1)
        sub = host[lo[0]:lo[0] + W + 1, lo[1]:lo[1] + W + 1, lo[2]:lo[2] + W + 1]
        ov, of, ok = mo.marching_cubes(sub, 0.0, [vs] * 3, tuple(int(x) for x in lo), (N, N, N))
        assert len(ok) > 100
        pos = np.searchsorted(keys, ok)
        assert np.array_equal(keys[pos], ok)
        assert np.array_equal(verts[pos], ov)
        inside = np.isin(tri_keys, ok).all(1)
        assert np.array_equal(tri_keys[inside], ok[of])


@pytest.mark.parametrize("N,cuts", [(128, (0, 61, 128)), (256, (0, 30, 94, 158, 222, 256))])
def test_slab_stitch_by_key_lookup_equals_the_full_volume_mesh(dev, N, cuts):
    """slab.stitch (drop each slab's copies of the next slab's first-plane vertices, re-index by key look-up) on an
    OPEN surface that leaves the volume through every face -- at sizes that run the interior-tile path of
    mc_classify -- equals marching cubes of the whole volume."""
    from alignsdf_b200 import slab
    dec = synthetic.make_decoder(0, init="default")
    s = synthetic.make_sample(0).to(dev)
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, N, keep_pass1=False)
    vs, org = float(vols["voxel"]), vols["origin"].tolist()
    for tag in ("hand", "obj"):
        vol = vols[tag]
        full = engine.marching_cubes(vol, 0.0, [vs] * 3, org, want_keys=True)
        parts, bounds = [], []
        for a, b in zip(cuts[:-1], cuts[1:]):
            hi = min(b + 1, N)                                    # one halo plane, except on the last slab
            p = engine.marching_cubes(vol[a:hi].contiguous(), 0.0, [vs] * 3, org, index0_offset=a, want_keys=True,
                                      check_range=False)
            parts.append((p["verts"], p["faces"], p["keys"]))
            bounds.append(b * N * N * 4 if b < N else slab.INT64_MAX)
        verts, faces = slab.stitch(parts, bounds)
        assert torch.equal(verts, full["verts"])
        assert torch.equal(faces, full["faces"])


@pytest.mark.parametrize("name", ["sep_both9_n24", "comb_cls_n12", "sep_both9_n32_handonly"])
def test_create_mesh_combined_decoder_end_to_end(dev, tmp_path, name):
    """Public API: files written, meshes == oracle marching cubes of the GPU volumes."""
    meta, g, dec, sample = helpers.load_case(name)
    s = helpers.to_cuda(sample)
    hb, ob = meta.get("hand_branch", True), meta.get("obj_branch", True)
    label = "pass2_cls" in g
    prefix = str(tmp_path / "img0")
    res = amesh.create_mesh_combined_decoder(hb, ob, label, dec, s.latent, s.mano_results, s.obj_results,
                                             None, s.specs, prefix, N=meta["N"], max_batch=2 ** 18,
                                             label_out=label, viz=label)
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, meta["N"], hb, ob, cls_branch=label)
    vs = float(vols["voxel"]); org = vols["origin"].tolist()
    for tag, use in (("hand", hb), ("obj", ob)):
        path = f"{prefix}_{tag}.ply"
        if not use:
            assert not os.path.exists(path)
            continue
        vol = vols[tag].cpu().numpy()
        if not (vol.min() <= 0 <= vol.max()):
            assert res[tag] is None and not os.path.exists(path)
            continue
        v, f, _ = mo.marching_cubes(vol, 0.0, [vs] * 3)
        pts = (np.asarray(org, np.float32)[None] + v).astype(np.float32)
        if tag == "obj" and hb:
            pts = pts * np.array([1]) + np.array([0, 0, 0])        # utils/mesh.py:366-369 with the hand's values
        ev, ef = mo.largest_component_if_split(pts, f)
        rv, rf = mo.read_ply(path)
        assert np.array_equal(rf, ef) and np.array_equal(rv, ev.astype(np.float32))
        assert np.array_equal(np.asarray(res[tag].faces), ef)
    if label:
        # utils/mesh.py:137-184: the decoder re-queried at the marching-cubes vertices, argmax of the class logits
        lab = np.load(prefix + "_hand_label.npz")
        assert lab["points"].shape[0] == lab["labels"].shape[0] > 0
        v, _, _ = mo.marching_cubes(vols["hand"].cpu().numpy(), 0.0, [vs] * 3)
        want_pts = v.copy()
        for k in range(3):
            want_pts[:, k] = np.float32(org[k]) + want_pts[:, k]
        assert np.array_equal(lab["points"], want_pts)                       # every vertex, origin-shifted
        with torch.no_grad():
            _, _, logits = orc.decode_points(dec.state_dict(), orc.decoder_cfg(dec), sample.latent,
                                             torch.from_numpy(want_pts), sample.specs, sample.mano_results,
                                             sample.obj_results)
        top2 = logits.topk(2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-4                             # ties within fp32 noise may go either way
        assert clear.float().mean() > 0.99
        assert np.array_equal(lab["labels"][clear.numpy()], logits.argmax(1).float().numpy()[clear.numpy()])
        # viz files (utils/mesh.py:258-278,300-329): one line per vertex; raw marching-cubes faces, coloured by label
        obj_lines = open(prefix + "_hand_label.obj").read().splitlines()
        assert len(obj_lines) == len(want_pts) and obj_lines[0].startswith("v ")
        ply_lines = open(prefix + "_hand_color.ply").read().splitlines()
        _, f_raw, _ = mo.marching_cubes(vols["hand"].cpu().numpy(), 0.0, [vs] * 3)
        assert f"element vertex {len(want_pts)}" in ply_lines and f"element face {len(f_raw)}" in ply_lines
        assert ply_lines[-1] == "3 %d %d %d" % tuple(f_raw[-1])


def test_label_out_without_classifier_raises_like_reference(dev, tmp_path):
    meta, g, dec, sample = helpers.load_case("sep_both9_n24")
    s = helpers.to_cuda(sample)
    with pytest.raises(IndexError):
        amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results,
                                           None, s.specs, str(tmp_path / "x"), N=16, label_out=True)


def test_no_surface_is_not_an_error(dev, tmp_path, caplog):
    dec = synthetic.make_decoder(0)
    with torch.no_grad():
        dec.linh4.bias += 5.0
        dec.lino4.bias += 5.0
    s = helpers.to_cuda(synthetic.make_sample(0))
    res = amesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results, s.obj_results,
                                             None, s.specs, str(tmp_path / "none"), N=10)
    assert res == {"hand": None, "obj": None}
    assert "Cannot reconstruct mesh" in caplog.text
    v = amesh.convert_sdf_samples_to_ply(torch.ones(8, 8, 8), [-1, -1, -1], 0.1, str(tmp_path / "n.ply"))
    assert v[0] is None and v[1] is None and v[2].tolist() == [0, 0, 0] and v[3].tolist() == [1]


def test_legacy_deep_sdf_api(dev, tmp_path):
    meta = helpers.golden_index()["legacy_n16"]
    g = np.load(os.path.join(helpers.GOLD, "legacy_n16.npz"))
    dec = synthetic.make_decoder(meta["seed"], "combined", 256, 3, "nerf")
    sample = synthetic.make_sample(meta["seed"], 256, 3, "nerf")

    class FirstOutput(torch.nn.Module):                    # DeepSDF-style single-output adapter
        def __init__(self, comb):
            super().__init__(); self.comb = comb

        def forward(self, x):
            return self.comb(x)[0]
    wrapped = FirstOutput(dec)
    N = meta["N"]
    xyz = orc.grid_points(N, 2.0 / (N - 1), [-1, -1, -1]).to(dev)
    sdf = legacy_utils.decode_sdf(wrapped, sample.latent.to(dev), xyz)
    assert sdf.shape == (N ** 3, 1)
    assert np.abs(sdf[:, 0].cpu().numpy().reshape(N, N, N) - g["volume"]).max() <= TOL
    m = legacy_mesh.create_mesh(wrapped, sample.latent.to(dev), str(tmp_path / "legacy"), N=N)
    rv, rf = mo.read_ply(str(tmp_path / "legacy.ply"))
    assert rf.shape[0] == m.faces.shape[0] > 0
    v, f, _ = mo.marching_cubes(sdf[:, 0].cpu().numpy().reshape(N, N, N), 0.0, [2.0 / (N - 1)] * 3)
    assert np.array_equal(rf, f)
    assert np.array_equal(rv, (np.float32(-1) + v).astype(np.float32))


def test_full_size_properties_128(dev):
    """BASELINE config #2 size (128^3 hand+obj): properties that need no CPU oracle run."""
    dec = synthetic.make_decoder(0)
    s = helpers.to_cuda(synthetic.make_sample(0))
    N = 128
    v1 = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, N)
    v2 = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, N)
    assert torch.equal(v1["hand"], v2["hand"]) and torch.equal(v1["obj"], v2["obj"])      # deterministic
    # sampled oracle check on 4096 random grid indices of pass 2
    idx = torch.randint(0, N ** 3, (4096,), generator=torch.Generator().manual_seed(0))
    xyz = orc.grid_points(N, v1["voxel"], v1["origin"])[idx]
    sc = s.to(torch.device("cpu"))
    with torch.no_grad():
        h, o, _ = orc.decode_points(dec.state_dict(), orc.decoder_cfg(dec), sc.latent, xyz, sc.specs,
                                    sc.mano_results, sc.obj_results)
    assert (v1["hand"].view(-1)[idx.to(dev)].cpu() - h[:, 0]).abs().max() <= TOL
    assert (v1["obj"].view(-1)[idx.to(dev)].cpu() - o[:, 0]).abs().max() <= TOL
    for tag in ("hand", "obj"):
        out = engine.marching_cubes(v1[tag], 0.0, [float(v1["voxel"])] * 3)
        inv = mo.mesh_invariants(out["verts"].cpu().numpy(), out["faces"].cpu().numpy())
        assert inv["nonmanifold_edges"] == 0 and inv["oriented"] and inv["closed"], (tag, inv)
        # every vertex lies within half a voxel of the iso-surface of the trilinear field
        assert out["faces"].shape[0] > 1000


def test_gpu_component_selection_matches_host_split(dev):
    """csrc/cc.cu (union-find labels, per-component area / open flag, compaction) == the generic
    edge-adjacency split + largest-area selection of the oracle, on meshes with several closed
    pieces, pieces cut by the volume boundary, a single piece and an empty mesh."""
    from alignsdf_b200 import trimesh_lite
    ax = np.linspace(-1, 1, 48, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    sph = lambda c, r: np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - np.float32(r)
    sp = 2 / 47
    cases = [np.minimum.reduce([sph((-0.4, 0, 0), 0.33), sph((0.5, 0.1, 0), 0.25), sph((0, 0.95, 0), 0.3),
                                sph((0, -0.6, 0.6), 0.12), sph((0.6, 0.6, 0.6), 0.2)]),
             sph((0, 0, 0), 0.5),
             sph((0, 0.95, 0), 0.3),
             np.minimum(sph((0.97, 0, 0), 0.3), sph((-0.3, 0, 0), 0.2)),
             np.minimum(sph((-0.4, 0, 0), 0.3), sph((0.4, 0, 0), 0.3))]      # equal areas: the first piece wins
    for vol in cases:
        out = engine.marching_cubes(torch.from_numpy(vol).to(dev), 0.0, [sp] * 3, [-1.0, -1.0, -1.0])
        p, f, info = engine.select_component(out["points"], out["faces"], out["verts"], vol.shape, [sp] * 3)
        pts, faces, verts = out["points"].cpu().numpy(), out["faces"].cpu().numpy(), out["verts"].cpu().numpy()
        ev, ef = mo.largest_component_if_split(pts, faces)
        assert np.array_equal(p.cpu().numpy(), ev) and np.array_equal(f.cpu().numpy(), ef), info
        m = trimesh_lite.largest_watertight_component_mc(pts, faces, verts, vol.shape, [sp] * 3)
        assert np.array_equal(m.vertices, ev) and np.array_equal(m.faces, ef)
        rec = engine.ply_face_records(f).cpu().numpy()
        assert rec.shape == (len(ef), 13) and np.all(rec[:, 0] == 3)
        assert np.array_equal(np.ascontiguousarray(rec[:, 1:]).view("<i4").reshape(-1, 3), ef)
    empty = torch.empty((0, 3), device=dev)
    p, f, info = engine.select_component(empty, torch.empty((0, 3), dtype=torch.int32, device=dev), empty, (4, 4, 4), [1.0] * 3)
    assert p.shape[0] == 0 and f.shape[0] == 0 and info["kept"] == "whole"


def test_nerf_embedder_api_matches_oracle(dev):
    """utils.get_nerf_embedder (utils/utils.py:521-533) on the GPU vs the oracle's torch-CPU restatement."""
    xyz = torch.rand(4, 333, 3, generator=torch.Generator().manual_seed(5)) * 2 - 1
    for multires in (1, 4, 10):
        embed, dim = autils.get_nerf_embedder(multires)
        got = embed(xyz.to(dev))
        want = orc.nerf_embedding(xyz, multires)
        assert dim == 3 + 6 * multires and tuple(got.shape) == (4, 333, dim)
        assert torch.equal(got[..., :3].cpu(), xyz)
        assert (got.cpu() - want).abs().max() <= 2e-6      # sinf/cosf of arguments up to 2^9: a few ulp


def test_pipelined_batch_equals_one_call_per_sample(dev, tmp_path):
    """mesh.create_meshes_pipelined (bind / grid passes / mesh extraction of consecutive samples overlapped
    on three threads and streams) writes byte-identical files to the one-call-per-sample loop."""
    dec = synthetic.make_decoder(0)
    samples = [synthetic.make_sample(i) for i in range(5)]
    N = 40
    seq = []
    for i, s in enumerate(samples):
        sc = s.to(dev)
        seq.append(amesh.create_mesh_combined_decoder(True, True, False, dec, sc.latent, sc.mano_results, sc.obj_results,
                                                      None, sc.specs, str(tmp_path / f"seq{i}"), N=N))
    got = amesh.create_meshes_pipelined(dec, samples, [str(tmp_path / f"pipe{i}") for i in range(5)], N=N)
    assert len(got) == 5
    for i in range(5):
        for tag in ("hand", "obj"):
            a, b = tmp_path / f"seq{i}_{tag}.ply", tmp_path / f"pipe{i}_{tag}.ply"
            assert os.path.exists(a) == os.path.exists(b)
            if os.path.exists(a):
                assert open(a, "rb").read() == open(b, "rb").read()
                assert np.array_equal(seq[i][tag].faces, got[i][tag].faces)
                assert np.array_equal(seq[i][tag].vertices, got[i][tag].vertices)
