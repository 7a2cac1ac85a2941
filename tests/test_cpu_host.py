"""CPU-side tests: oracle vs the reference's golden vectors, host logic, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from alignsdf_b200 import packer, synthetic, trimesh_lite
from oracle import alignsdf_oracle as orc
from oracle import mc_oracle as mo
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", helpers.field_cases())
def test_synthetic_generator_is_bit_reproducible(name):
    meta, g, dec, sample = helpers.load_case(name)
    assert synthetic.state_digest(dec) == meta["digest"]


@pytest.mark.parametrize("name", helpers.field_cases())
def test_oracle_matches_reference_golden(name):
    """oracle/alignsdf_oracle.py vs outputs captured from the real reference (tol 1e-6; 2e-6 for the
    LayerNorm decoders, whose outputs span the whole tanh range and amplify fp32 summation-order noise)."""
    meta, g, dec, sample = helpers.load_case(name)
    # non-engineered decoders with a large last-layer gain: two faithful fp32 evaluations differ by summation-order
    # noise that grows with the output range (oracle/make_golden.py)
    tol = 2e-6 if meta.get("weight_norm") is False else 1e-6 * max(1.0, meta.get("out_gain", 1.0) / 2.0)
    sd = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    res = orc.two_pass_field(sd, orc.decoder_cfg(dec), sample.latent, sample.specs,
                             sample.mano_results, sample.obj_results, meta["N"],
                             meta.get("hand_branch", True), meta.get("obj_branch", True), cam_intr=sample.cam_intr)
    assert np.float32(float(res["voxel"])) == g["new_voxel"]
    assert np.array_equal(res["origin"].numpy(), g["new_origin"])
    for key, arr in (("pass1_hand", res["pass1_hand"]), ("pass1_obj", res["pass1_obj"]),
                     ("pass2_hand", res["hand"]), ("pass2_obj", res["obj"])):
        if key in g:
            assert np.abs(arr.numpy() - g[key]).max() <= tol, key
    if "pass2_cls" in g:
        assert np.array_equal(res["cls"].numpy().reshape(-1).astype(np.int32), g["pass2_cls"])
    if "xyz2" in g:
        xyz = orc.grid_points(meta["N"], torch.tensor(g["new_voxel"]), torch.tensor(g["new_origin"]))
        assert np.array_equal(xyz.numpy(), g["xyz2"])


def test_oracle_grid_above_2_pow_24():
    g = np.load(os.path.join(helpers.GOLD, "grid512_windows.npz"))
    for key in g.files:
        _, a, b = key.split("_")
        got = orc.grid_points(512, 2.0 / 511, [-1, -1, -1], "reference", int(a), int(b)).numpy()
        assert np.array_equal(got, g[key])
    # the shear really is there: axis 1 picks up z/N
    p = orc.grid_points(512, 1.0, [0, 0, 0], "reference", 5 * 512 + 7, 5 * 512 + 8)[0]
    assert p[2] == 7 and abs(float(p[1]) - (5 + 7 / 512)) < 1e-5
    r = orc.grid_points(512, 1.0, [0, 0, 0], "regular", 5 * 512 + 7, 5 * 512 + 8)[0]
    assert r.tolist() == [0.0, 5.0, 7.0]


@pytest.mark.parametrize("name", helpers.field_cases())
def test_folded_network_matches_reference_golden(name):
    """packer.fold_decoder (latent -> bias, pose-align -> [out,3]) vs the reference (tol 1e-6)."""
    meta, g, dec, sample = helpers.load_case(name)
    if meta.get("pixel_align"):
        pytest.skip("per-point latents are not folded: test_pixel_align_projected_maps_match_reference_golden below")
    topo = packer.decoder_topology(dec)
    N = meta["N"]
    xyz = orc.grid_points(N, 2.0 / (N - 1), [-1, -1, -1])
    nerf = packer.nerf_freqs(sample.specs, sample.mano_results)
    br = packer.fold_decoder(topo, sample.latent, sample.specs, sample.mano_results, sample.obj_results,
                             feature_mode=nerf > 0)
    if nerf:          # positional encoding is not affine in xyz: the kernel encodes, the fold sees features
        assert nerf == (meta["pf"] - 3) // 6
        u = orc.nerf_embedding(xyz, nerf).numpy()
    else:
        u = xyz.numpy()
    out = packer.folded_forward_numpy(br, u, topo.pre_tanh)
    hand, obj = (out[0][:, 0], out[1][:, 0]) if topo.kind == "separate" else (out[0][:, 0], out[0][:, 1])
    tol = 3e-6 if meta.get("weight_norm") is False else 1e-6 * max(1.0, meta.get("out_gain", 1.0) / 2.0)
    if "pass1_hand" in g:
        assert np.abs(hand - g["pass1_hand"].reshape(-1)).max() <= tol
    if "pass1_obj" in g:
        assert np.abs(obj - g["pass1_obj"].reshape(-1)).max() <= tol


def test_feature_mode_fold_matches_oracle():
    meta, g, dec, sample = helpers.load_case("sep_both9_n24")
    topo = packer.decoder_topology(dec)
    xyz = orc.grid_points(8, 2.0 / 7, [-1, -1, -1])
    feats = orc.embed(xyz, sample.specs, sample.mano_results, sample.obj_results)
    br = packer.fold_decoder(topo, sample.latent, sample.specs, None, None, feature_mode=True)
    outs = []
    for b, (tag, _) in zip(br, topo.branches):
        idx = packer.branch_feature_index(topo, tag)
        outs.append(packer.folded_forward_numpy([b], feats.numpy()[:, idx])[0][:, 0])
    sd = dec.state_dict()
    h, o, _ = orc.decode_points(sd, orc.decoder_cfg(dec), sample.latent, xyz, sample.specs,
                                sample.mano_results, sample.obj_results)
    assert np.abs(outs[0] - h[:, 0].detach().numpy()).max() <= 1e-6
    assert np.abs(outs[1] - o[:, 0].detach().numpy()).max() <= 1e-6


def test_embedding_affine_matches_explicit_embedding():
    for pf, style in ((9, "both"), (51, "hand"), (6, "hand"), (6, "obj"), (54, "both")):
        s = synthetic.make_sample(3, 256, pf, style)
        xyz = torch.rand(64, 3) * 2 - 1
        ref = orc.kinematic_embedding(xyz, s.mano_results, pf, s.specs["SdfScaleFactor"], s.obj_results, style)
        got = packer.embed_points_torch(xyz, s)
        assert torch.abs(ref - got).max() < 5e-6


def test_non_rigid_pose_is_rejected():
    s = synthetic.make_sample(0)
    s.obj_results["obj_trans"][0, 3, 0] = 0.1            # projective row -> w != 1
    with pytest.raises(ValueError, match="not affine"):
        packer.embedding_affine(s.specs, s.mano_results, s.obj_results)


def test_unsupported_variants_raise():
    s = synthetic.make_sample(0, 256, 9, "both")
    dec = synthetic.make_decoder(0)
    topo = packer.decoder_topology(dec)
    # PixelAlign: the fold leaves the latent out (it is applied per point from projected feature maps) ...
    br = packer.fold_decoder(topo, None, dict(s.specs, PixelAlign=True), s.mano_results, s.obj_results)
    zero = packer.fold_decoder(topo, torch.zeros(1, 256), s.specs, s.mano_results, s.obj_results)
    assert all(np.array_equal(a.B, b.B) for x, y in zip(br, zero) for a, b in zip(x.layers, y.layers))
    # ... and the set-up refuses what it cannot project
    from alignsdf_b200 import pixel_align
    with pytest.raises(ValueError, match="feature map"):
        pixel_align.setup(topo, s.latent, dict(s.specs, PixelAlign=True), s.mano_results, None, None, True, "cpu")
    with pytest.raises(ValueError, match="cam_intr"):
        pixel_align.setup(topo, torch.zeros(1, 256, 4, 4), dict(s.specs, PixelAlign=True), s.mano_results, None, None,
                          True, "cpu")
    with pytest.raises(ValueError, match="not an affine map"):
        packer.embedding_affine(dict(s.specs, PointFeatSize=9, EncodeStyle="nerf"), None, None)
    with pytest.raises(ValueError, match="3 \\+ 6"):
        packer.nerf_freqs(dict(s.specs, PointFeatSize=10, EncodeStyle="nerf"), None)
    with pytest.raises(AttributeError):
        from alignsdf_b200.decoders import SeparateDecoder
        SeparateDecoder(256, 9, "both", [512] * 4, use_classifier=True)


def test_reference_style_state_dict_prefixes_are_accepted():
    dec = synthetic.make_decoder(1)
    sd = {"module.decoder." + k: v for k, v in dec.state_dict().items()}

    class Holder:
        point_feat_size, encode_style, latent_in = 9, "both", (2,)

        def state_dict(self):
            return sd
    t = packer.decoder_topology(Holder())
    assert t.kind == "separate" and t.latent_size == 256 and t.n_layers == 5


def test_shared_library_exports_every_declared_symbol():
    """The C-ABI library loads on a GPU-less host and exports what include/*.h declares."""
    from alignsdf_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "alignsdf_b200.h")).read()
    declared = set(re.findall(r"\b(asdf_[a-z0-9_]+)\s*\(", header))
    assert {"asdf_simt_eval", "asdf_tc_eval", "asdf_mc_count", "asdf_mc_emit", "asdf_grid_points"} <= declared
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.asdf_abi_version.restype = ctypes.c_int
    assert lib.asdf_abi_version() == 4


def test_ctypes_structs_match_header_sizes(tmp_path):
    """sizeof / offsetof as gcc sees include/alignsdf_b200.h == the ctypes mirrors in _lib.py."""
    import subprocess
    from alignsdf_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "alignsdf_b200.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(asdf_query), sizeof(asdf_simt_desc),'
        ' sizeof(asdf_tc_launch), sizeof(asdf_mc_params), offsetof(asdf_query, points_dev),'
        ' offsetof(asdf_simt_desc, table), offsetof(asdf_mc_params, spacing), offsetof(asdf_tc_launch, status_dev),'
        ' sizeof(asdf_tc_bind_desc), offsetof(asdf_tc_bind_desc, decoder_stride), offsetof(asdf_tc_bind_desc, w_scale),'
        ' offsetof(asdf_tc_launch, grid_dev));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.Query), ctypes.sizeof(_lib.SimtDesc), ctypes.sizeof(_lib.TcLaunch),
            ctypes.sizeof(_lib.McParams), _lib.Query.points_dev.offset, _lib.SimtDesc.table.offset,
            _lib.McParams.spacing.offset, _lib.TcLaunch.status_dev.offset,
            ctypes.sizeof(_lib.TcBindDesc), _lib.TcBindDesc.decoder_stride.offset, _lib.TcBindDesc.w_scale.offset,
            _lib.TcLaunch.grid_dev.offset]
    assert got == want


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the GPU-less failure mode")
def test_product_path_fails_loudly_without_gpu(tmp_path):
    from alignsdf_b200 import mesh
    from alignsdf_b200._lib import AsdfError
    dec, s = synthetic.make_decoder(0), synthetic.make_sample(0)
    with pytest.raises(AsdfError, match="no CPU path"):
        mesh.create_mesh_combined_decoder(True, True, False, dec, s.latent, s.mano_results,
                                          s.obj_results, None, s.specs, str(tmp_path / "x"), N=8)


def test_trimesh_lite_split_and_ply_roundtrip(tmp_path):
    ax = np.linspace(-1, 1, 28, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    two = np.minimum(np.sqrt((x + 0.45) ** 2 + y ** 2 + z ** 2) - 0.35,
                     np.sqrt((x - 0.5) ** 2 + y ** 2 + z ** 2) - 0.22)
    v, f, _ = mo.marching_cubes(two, 0.0)
    m = trimesh_lite.Mesh(v, f)
    parts = trimesh_lite.split(m)
    assert len(parts) == 2 and all(p.is_watertight for p in parts)
    big = max(parts, key=lambda p: p.area)
    v2, f2 = mo.largest_component_if_split(v, f)
    assert np.array_equal(big.vertices, v2) and np.array_equal(big.faces, f2)
    path = str(tmp_path / "m.ply")
    big.export(path)
    rv, rf = mo.read_ply(path)
    assert np.array_equal(rv, v2.astype(np.float32)) and np.array_equal(rf, f2)
    open_mesh = trimesh_lite.Mesh(v, f[:-3])             # punch a hole: no longer watertight
    assert len(trimesh_lite.split(open_mesh)) == 1


def test_fast_component_filter_matches_generic_split():
    """trimesh_lite.largest_watertight_component_mc (O(V+F), MC meshes only) == generic split + max area."""
    ax = np.linspace(-1, 1, 40, dtype=np.float32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    sph = lambda c, r: np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - np.float32(r)
    sp = 2 / 39
    cases = [np.minimum.reduce([sph((-0.4, 0, 0), 0.33), sph((0.5, 0.1, 0), 0.25), sph((0, 0.95, 0), 0.3),
                                sph((0, -0.6, 0.6), 0.12)]),        # one piece is cut by the volume boundary
             sph((0, 0, 0), 0.5),                                   # single closed piece
             sph((0, 0.95, 0), 0.3),                                # single open piece
             np.minimum(sph((0.97, 0, 0), 0.3), sph((-0.3, 0, 0), 0.2))]   # open + closed -> one candidate -> whole
    for vol in cases:
        v, f, _ = mo.marching_cubes(vol, 0.0, [sp] * 3)
        pts = (np.float32(-1) + v).astype(np.float32)
        m = trimesh_lite.largest_watertight_component_mc(pts, f, v, vol.shape, [sp] * 3)
        ev, ef = mo.largest_component_if_split(pts, f)
        assert np.array_equal(m.vertices, ev) and np.array_equal(m.faces, ef)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's torch-CPU path through the oracle port) runs without a GPU
    and prints one JSON line carrying the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--N", "64"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mq/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == dict(value=line["value"], unit="Mq/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    # both arms print the SAME config object (the driver compares them): one helper, no run-dependent figures in it
    sys.path.insert(0, root)
    import bench
    assert line["config"] == bench.workload_config(64, 16, 1)
    assert set(line["config"]) == {"workload", "decoder", "parallelism", "l2"}
    src = open(os.path.join(root, "bench.py")).read()
    assert src.count("config=workload_config(") == 2


def test_ply_writer_from_device_style_face_records(tmp_path):
    """export_ply_records (face records serialised as uint8 [F,13], what engine.ply_face_records builds on the
    GPU) writes the same bytes as export_ply."""
    rng = np.random.default_rng(3)
    v = rng.standard_normal((37, 3)).astype(np.float32)
    f = rng.integers(0, 37, (55, 3)).astype(np.int32)
    rec = np.empty((55, 13), np.uint8)
    rec[:, 0] = 3
    rec[:, 1:] = f.view(np.uint8).reshape(55, 12)
    a, b = str(tmp_path / "a.ply"), str(tmp_path / "b.ply")
    trimesh_lite.export_ply(a, v, f)
    trimesh_lite.export_ply_records(b, v, rec)
    assert open(a, "rb").read() == open(b, "rb").read()
    rv, rf = mo.read_ply(b)
    assert np.array_equal(rv, v) and np.array_equal(rf, f)
    trimesh_lite.export_ply_records(b, v[:0], rec[:0])          # empty mesh: header only
    rv, rf = mo.read_ply(b)
    assert rv.shape == (0, 3) and rf.shape[0] == 0


def _bits(x):
    return int(np.array([x], np.float32).view(np.int32)[0])


def test_kernel_selection_logic_from_flag_words():
    """BoundSample.decide: calibration error bounds and operand-range words -> kernel level (pure host logic; the
    words come from the device, or MAX-reduced from all ranks of a slab group)."""
    from alignsdf_b200 import engine
    dec = synthetic.make_decoder(5)
    s = synthetic.make_sample(5)
    eng = engine.DecoderEngine(dec, "cuda")                      # no device work happens at construction
    assert eng.level == engine.LEVEL_F8 and eng.fast_tau() is None

    def fresh():
        b = eng.bind(s.latent, s.specs, s.mano_results, s.obj_results)
        b._calib = object()                                      # "a calibration result is pending"
        return b
    # benign sample: fp16 + e4m3 accepted; its single-product error becomes the decoder's bound
    b = fresh()
    assert b.decide([0, 0, _bits(2e-6), _bits(2e-7), _bits(4e-5)]) == engine.LEVEL_F8
    assert eng.level == engine.LEVEL_F8 and abs(eng.calib["f1"] - 4e-5) < 1e-9
    assert abs(eng.fast_tau() - engine.FAST_TAU_FACTOR * 4e-5) < 1e-9 and eng.fast_tau(2048) is None
    # a fast bounding-box pass whose threshold this sample's own error does not respect is void
    b = fresh()
    b._fast_tau = 1e-4
    assert b.decide([0, 0, _bits(2e-6), _bits(2e-7), _bits(9e-5)]) == engine.LEVEL_F8 and b.redo_fast
    b = fresh()
    b._fast_tau = 4e-4
    b.decide([0, 0, _bits(2e-6), _bits(2e-7), _bits(9e-5)])
    assert not b.redo_fast
    # e4m3 operand range exceeded on some rank -> all-fp16 kind, for good
    b = fresh()
    assert b.decide([1, 0, 0, 0, 0]) == engine.LEVEL_F16 and eng.level == engine.LEVEL_F16
    # calibration rejects both tensor-core kinds (also NaN) -> fp32 kernel
    eng2 = engine.DecoderEngine(dec, "cuda")
    b = eng2.bind(s.latent, s.specs, s.mano_results, s.obj_results)
    b._calib = object()
    assert b.decide([0, 0, _bits(3e-5), _bits(float("nan")), _bits(1e-3)]) == engine.LEVEL_SIMT
    assert eng2.level == engine.LEVEL_SIMT and engine.STATS["tc_to_simt"] >= 1
    # a bind failure (operands do not fit fp16) leaves only the fp32 kernel
    eng3 = engine.DecoderEngine(dec, "cuda")
    assert eng3.bind(s.latent, s.specs, s.mano_results, s.obj_results).decide([2, 0, 0, 0, 0]) == engine.LEVEL_SIMT
    # a forced path is not overridden (sticky level untouched)
    eng4 = engine.DecoderEngine(dec, "cuda")
    eng4.path = "f8"
    b = eng4.bind(s.latent, s.specs, s.mano_results, s.obj_results)
    b._calib = object()
    assert b.decide([0, 0, _bits(3e-5), _bits(1e-7), 0]) == engine.LEVEL_F16 and eng4.level == engine.LEVEL_F8


def test_label_visualisation_files_equal_the_reference(tmp_path):
    """viz outputs of the label pass (utils/mesh.py:258-278,300-329) byte for byte against files written by the
    reference's own functions (oracle/make_golden_viz.py)."""
    from alignsdf_b200 import mesh as amesh
    g = np.load(os.path.join(helpers.GOLD, "viz_label.npz"))
    pts, labels = torch.from_numpy(g["points"]), torch.from_numpy(g["labels"])
    for tag, off, sc in (("plain", None, None), ("moved", g["offset"], g["scale"])):
        amesh.write_verts_label_to_obj(pts, labels, str(tmp_path / "a.obj"), off, sc)
        amesh.write_color_labeled_ply(pts, g["faces"], labels, str(tmp_path / "a.ply"), off, sc)
        assert open(tmp_path / "a.obj", "rb").read() == g[f"obj_{tag}"].tobytes()
        assert open(tmp_path / "a.ply", "rb").read() == g[f"ply_{tag}"].tobytes()


def _pa_taps_numpy(pa, xyz):
    """Numpy statement of csrc/k1_simt.cu::pixel_align_taps in float32: per point 16 (map row, weight) taps."""
    f = np.float32
    P = pa["point_affine"].reshape(3, 4).astype(f)
    C = pa["cam"].reshape(3, 4).astype(f)
    x = xyz.astype(f)
    c = (x @ P[:, :3].T + P[:, 3]).astype(f)
    h = (c @ C[:, :3].T + C[:, 3]).astype(f)
    size = f(pa["image_size"])
    with np.errstate(all="ignore"):
        u = (h[:, 0] / h[:, 2] / size * f(2) - f(1)).astype(f)
        v = (h[:, 1] / h[:, 2] / size * f(2) - f(1)).astype(f)
    fh, fw = pa["fh"], pa["fw"]
    n = len(x)
    ti = np.zeros((n, 16), np.int64)
    tw = np.zeros((n, 16), f)
    inside = (u >= -1) & (u <= 1) & (v >= -1) & (v <= 1)
    ti[~inside, 0], tw[~inside, 0] = fh * fw, 1.0
    ix, iy = (u + f(1)) * f(0.5) * f(fw - 1), (v + f(1)) * f(0.5) * f(fh - 1)
    fx, fy = np.floor(ix), np.floor(iy)

    def coeffs(t):
        A = f(-0.75)
        x0, x1, x2, x3 = t + f(1), t, f(1) - t, f(2) - t
        return np.stack([((A * x0 - f(5) * A) * x0 + f(8) * A) * x0 - f(4) * A,
                         ((A + f(2)) * x1 - (A + f(3))) * x1 * x1 + f(1),
                         ((A + f(2)) * x2 - (A + f(3))) * x2 * x2 + f(1),
                         ((A * x3 - f(5) * A) * x3 + f(8) * A) * x3 - f(4) * A], 1).astype(f)
    cx, cy = coeffs((ix - fx).astype(f)), coeffs((iy - fy).astype(f))
    for j in range(4):
        for k in range(4):
            xx, yy = fx.astype(np.int64) - 1 + k, fy.astype(np.int64) - 1 + j
            ok = inside & (xx >= 0) & (xx < fw) & (yy >= 0) & (yy < fh)
            ti[ok, 4 * j + k] = (yy * fw + xx)[ok]
            tw[ok, 4 * j + k] = (cy[:, j] * cx[:, k])[ok]
    return ti, tw, inside


@pytest.mark.parametrize("name", ["sep_pa_both9_n12", "comb_pa_xyz3_n10"])
def test_pixel_align_projected_maps_match_reference_golden(name):
    """PixelAlign without a GPU: the host side of the path (pixel_align.setup: latent columns applied to the feature
    map once per sample, projection folded into one affine map) + a numpy statement of the kernel's 16-tap bicubic
    gather reproduce the REAL reference's pass-1 fields (utils/utils.py:536-566 through F.grid_sample)."""
    from alignsdf_b200 import pixel_align
    meta, g, dec, sample = helpers.load_case(name)
    topo = packer.decoder_topology(dec)
    N = meta["N"]
    xyz = orc.grid_points(N, 2.0 / (N - 1), [-1, -1, -1]).numpy()
    affine = packer.embedding_affine(sample.specs, sample.mano_results, sample.obj_results)
    br = packer.fold_decoder(topo, None, sample.specs, sample.mano_results, sample.obj_results, affine=affine)
    pa = pixel_align.setup(topo, sample.latent, sample.specs, sample.mano_results, sample.cam_intr, affine, False, "cpu")
    ti, tw, inside = _pa_taps_numpy(pa, xyz)
    assert inside.any() and (~inside).any()                      # both rules are exercised: bicubic taps and the mean feature
    maps = pa["maps"].numpy()
    outs = []
    for b, branch in enumerate(br):
        x = None
        for l, fl in enumerate(branch.layers):
            y = np.broadcast_to(fl.B, (len(xyz), fl.B.shape[0])).astype(np.float32).copy()
            if fl.Wx is not None:
                y += x @ fl.Wx.T
            if fl.M is not None:
                y += xyz.astype(np.float32) @ fl.M.T
            if l in pa["layers"]:
                G = maps[b, pa["layers"].index(l)]                # [fh fw + 1, npad]
                y += np.einsum("pt,ptn->pn", tw, G[ti][:, :, :y.shape[1]])
            if l == len(branch.layers) - 1:
                y = np.tanh(np.tanh(y)) if topo.pre_tanh else np.tanh(y)
            else:
                assert fl.ln is None
                y = np.maximum(y, 0)
            x = y
        outs.append(x)
    hand, obj = (outs[0][:, 0], outs[1][:, 0]) if topo.kind == "separate" else (outs[0][:, 0], outs[0][:, 1])
    assert np.abs(hand - g["pass1_hand"].reshape(-1)).max() <= 2e-6
    assert np.abs(obj - g["pass1_obj"].reshape(-1)).max() <= 2e-6


def test_deep_sdf_package_surface_mirrors_the_reference():
    """deep_sdf/__init__.py:4-9 star-imports its sub-modules; the names of the path and of its evaluation resolve the
    same way here (``deep_sdf.create_mesh``, ``deep_sdf.metrics.chamfer.compute_trimesh_chamfer`` of evaluate.py:64 ...)."""
    import inspect
    import alignsdf_b200.deep_sdf as deep_sdf
    for name in ("create_mesh", "convert_sdf_samples_to_ply", "decode_sdf", "ICP_T_S", "compute_trimesh_chamfer",
                 "procrustes", "procrustes_without_rot", "icp", "transform_points"):
        assert callable(getattr(deep_sdf, name)), name
    assert deep_sdf.metrics.chamfer.compute_trimesh_chamfer is deep_sdf.compute_trimesh_chamfer
    assert deep_sdf.metrics.icp_trans_scale.ICP_T_S is deep_sdf.ICP_T_S
    # same positional signatures as the reference's functions
    assert list(inspect.signature(deep_sdf.compute_trimesh_chamfer).parameters)[:4] == \
        ["gt_mesh_filename", "pred_mesh_filename", "optim", "rot"]
    assert list(inspect.signature(deep_sdf.create_mesh).parameters)[:5] == ["decoder", "latent_vec", "filename", "N", "max_batch"]
    assert list(inspect.signature(deep_sdf.decode_sdf).parameters) == ["decoder", "latent_vector", "queries"]
    assert list(inspect.signature(deep_sdf.icp).parameters) == ["a", "b", "initial", "threshold", "max_iterations", "rot"]
    assert list(inspect.signature(deep_sdf.procrustes).parameters) == ["a", "b", "reflection", "translation", "scale", "return_cost"]
    icp = deep_sdf.ICP_T_S
    for m in ("sample_mesh", "run_icp_f", "run_icp", "get_trans_scale", "export_source_mesh"):
        assert callable(getattr(icp, m)), m
    assert list(inspect.signature(icp.run_icp_f).parameters)[1:] == ["max_iter", "stop_error", "stop_improvement", "verbose"]


def test_top_level_package_resolves_the_path_functions_like_the_reference_utils_package():
    """utils/__init__.py:3-4 star-imports utils.mesh and utils.utils: ``utils.<name>`` and ``utils.mesh.<name>`` are the
    same objects there; same here, with the argument order of the reference's functions."""
    import inspect
    import alignsdf_b200 as pkg
    from alignsdf_b200 import mesh as amesh, utils as autils
    assert pkg.create_mesh_combined_decoder is amesh.create_mesh_combined_decoder
    assert pkg.convert_sdf_samples_to_ply is amesh.convert_sdf_samples_to_ply
    assert pkg.decode_sdf_multi_output is autils.decode_sdf_multi_output
    assert pkg.kinematic_embedding is autils.kinematic_embedding
    want = ["hand_branch", "obj_branch", "cls_branch", "decoder", "latent_vec", "mano_results", "obj_results", "cam_intr",
            "specs", "filename", "N", "max_batch", "offset", "scale", "device", "label_out", "viz", "eval_mode", "task"]
    assert list(inspect.signature(pkg.create_mesh_combined_decoder).parameters)[:len(want)] == want     # utils/mesh.py:17
    assert list(inspect.signature(pkg.decode_sdf_multi_output).parameters) == \
        ["decoder", "latent_vector", "queries", "mano_results", "cam_intr", "specs"]                    # utils/utils.py:561
    assert list(inspect.signature(pkg.convert_sdf_samples_to_ply).parameters)[:8] == \
        ["pytorch_3d_sdf_tensor", "voxel_grid_origin", "voxel_size", "ply_filename_out", "offset", "scale", "eval_mode", "task"]
    assert "create_mesh_combined_decoder" in dir(pkg)
    with pytest.raises(AttributeError):
        pkg.no_such_function
