"""Generate tests/golden/*.npz by running the REAL reference.  TEST INFRASTRUCTURE.

Runs only in the authoring container (needs /root/reference, which does not
exist on the GPU box).  It imports the reference's own, unmodified
``utils.mesh.create_mesh_combined_decoder`` / ``deep_sdf.mesh.create_mesh`` /
``networks.model.{SeparateDecoder,CombinedDecoder}`` with the three shims of
SURVEY.md Appendix C:

1. empty stub modules for packages that are not installed (trimesh, lmdb,
   skimage, plyfile, chumpy, ...);
2. ``torch.Tensor.cuda = identity`` (GPU-less host; the reference hard-codes
   ``.cuda()`` at networks/model.py:188,350 and utils/mesh.py:48);
3. ``skimage.measure.marching_cubes_lewiner`` := a capture hook that records
   the volume + spacing and raises (the reference swallows the exception,
   utils/mesh.py:353-358).

For every case it (a) captures the reference's grid coordinates, pass-1 and
pass-2 volumes and re-grid parameters, (b) asserts oracle/alignsdf_oracle.py
reproduces them (grid bit-exact, fields <= 1e-6) -- this is what PINS the
oracle -- and (c) writes the captured reference outputs as a fixture.

    python oracle/make_golden.py            # regenerate everything
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from alignsdf_b200 import synthetic  # noqa: E402
from oracle import alignsdf_oracle as orc  # noqa: E402

CASES = [
    # name, kind, pf, style, N, seed, hand_branch, obj_branch, extras
    dict(name="sep_both9_n24", kind="separate", pf=9, style="both", N=24, seed=0),
    dict(name="sep_both9_n32_handonly", kind="separate", pf=9, style="both", N=32, seed=1,
         obj_branch=False),
    dict(name="sep_nerf3_n16", kind="separate", pf=3, style="nerf", N=16, seed=2),
    dict(name="sep_hand51_n16", kind="separate", pf=51, style="hand", N=16, seed=3),
    dict(name="sep_hand6_n12", kind="separate", pf=6, style="hand", N=12, seed=4),
    dict(name="sep_obj6_n12", kind="separate", pf=6, style="obj", N=12, seed=5),
    dict(name="sep_both54_n12", kind="separate", pf=54, style="both", N=12, seed=6),
    dict(name="comb_both9_n16", kind="combined", pf=9, style="both", N=16, seed=7),
    # classifier + xyz_in_all crashes inside the reference itself (classifier_head is
    # Linear(512, 6) but the penultimate activations are 512-pf wide), so they are separate cases
    dict(name="comb_cls_n12", kind="combined", pf=9, style="both", N=12, seed=8,
         use_classifier=True, cls_branch=True),
    dict(name="comb_xyzall_n12", kind="combined", pf=9, style="both", N=12, seed=10,
         xyz_in_all=True),
    dict(name="sep_both9_n20_objonly", kind="separate", pf=9, style="both", N=20, seed=9,
         hand_branch=False),
    # NeRF positional encoding of xyz (utils/mesh.py:54-55): pf = 3 + 6 * multires
    dict(name="sep_nerf27_n12", kind="separate", pf=27, style="nerf", N=12, seed=11),
    dict(name="comb_nerf15_n12", kind="combined", pf=15, style="nerf", N=12, seed=12),
    # LayerNorm instead of weight norm (NetworkSpecs.weight_norm = false, networks/model.py:254-255,317-319)
    dict(name="sep_ln_both9_n12", kind="separate", pf=9, style="both", N=12, seed=13, weight_norm=False),
    dict(name="comb_ln_both9_n12", kind="combined", pf=9, style="both", N=12, seed=14, weight_norm=False),
    # NON-engineered decoders: the plain random generator without the ellipsoid wiring, last layer
    # x1 / x4 / x16 (|sdf| up to ~0.1 / 0.45 / 0.97), and torch's default initialisation; last-layer
    # biases shifted so ~30 % of the cube is inside (the shifts are recorded in the fixture's meta)
    dict(name="sep_plain_g1_n32", kind="separate", pf=9, style="both", N=32, seed=20, init="plain", out_gain=1.0),
    dict(name="sep_plain_g4_n32", kind="separate", pf=9, style="both", N=32, seed=21, init="plain", out_gain=4.0),
    dict(name="sep_plain_g16_n32", kind="separate", pf=9, style="both", N=32, seed=22, init="plain", out_gain=16.0),
    dict(name="sep_default_n32", kind="separate", pf=9, style="both", N=32, seed=23, init="default"),
    dict(name="comb_plain_g4_n24", kind="combined", pf=9, style="both", N=24, seed=24, init="plain", out_gain=4.0),
    dict(name="comb_default_n16", kind="combined", pf=9, style="both", N=16, seed=25, init="default"),
    # PixelAlign (utils/utils.py:536-566): the latent is an image feature map sampled per point (bicubic); ~40 % of
    # the cube projects outside the image and takes the mean feature
    dict(name="sep_pa_both9_n12", kind="separate", pf=9, style="both", N=12, seed=30, pixel_align=(10, 12)),
    dict(name="comb_pa_xyz3_n10", kind="combined", pf=3, style="nerf", N=10, seed=31, pixel_align=(7, 5)),
]


def install_shims():
    for name in ["trimesh", "lmdb", "skimage", "skimage.measure", "plyfile", "chumpy",
                 "soft_renderer", "soft_renderer.functional"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REF not in sys.path:
        sys.path.insert(0, REF)


class Capture:
    def __init__(self):
        self.mc_calls = []       # (volume copy, spacing) in call order
        self.cube = None         # (hand_in, obj_in, new_voxel, new_origin)
        self.xyz = []            # per chunk embedded-or-raw xyz before embedding
        self.logits = []

    def mc_hook(self, vol, level=0.0, spacing=(1., 1., 1.)):
        self.mc_calls.append((np.array(vol, copy=True), [float(s) for s in spacing]))
        raise ValueError("captured by oracle/make_golden.py")


def run_reference_case(case):
    import utils.mesh as ref_mesh            # the reference's module
    import networks.model as ref_model
    import skimage.measure

    kind, pf, style, N, seed = case["kind"], case["pf"], case["style"], case["N"], case["seed"]
    ns = dict(synthetic.NETWORK_SPECS)
    if case.get("xyz_in_all"):
        ns["xyz_in_all"] = True
    if case.get("weight_norm") is False:
        ns["weight_norm"] = False
    use_cls = bool(case.get("use_classifier", False))
    mine = synthetic.make_decoder(seed, kind, 256, pf, style, ns, use_classifier=use_cls,
                                  init=case.get("init", "engineered"), out_gain=case.get("out_gain", 1.0))
    sample = synthetic.make_sample(seed, 256, pf, style, pixel_align=case.get("pixel_align"))

    ref_cls = ref_model.SeparateDecoder if kind == "separate" else ref_model.CombinedDecoder
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_dec = ref_cls(256, pf, style, use_classifier=use_cls, **ns)
    missing = ref_dec.load_state_dict(mine.state_dict(), strict=True)   # key/shape compatibility
    assert not missing.missing_keys and not missing.unexpected_keys
    ref_dec.eval()

    cap = Capture()
    skimage.measure.marching_cubes_lewiner = cap.mc_hook
    orig_cube = ref_mesh.get_higher_res_cube
    orig_decode = ref_mesh.decode_sdf_multi_output
    orig_embed = ref_mesh.kinematic_embedding
    orig_nerf = ref_mesh.get_nerf_embedder

    def cube_wrap(hb, ob, vh, vo, n, origin, vs):
        nv, no = orig_cube(hb, ob, vh, vo, n, origin, vs)
        cap.cube = (None if vh is None else vh.clone().numpy(),
                    None if vo is None else vo.clone().numpy(), nv.clone(), no.clone())
        return nv, no

    def embed_wrap(xyz, *a, **k):
        cap.xyz.append(xyz.clone().numpy())
        return orig_embed(xyz, *a, **k)

    def decode_wrap(decoder, latent, queries, *a, **k):
        if pf == 3:
            cap.xyz.append(queries.clone().numpy())
        out = orig_decode(decoder, latent, queries, *a, **k)
        if out[2].dim() == 2:
            cap.logits.append(out[2].clone().numpy())
        return out

    ref_mesh.get_higher_res_cube = cube_wrap
    ref_mesh.decode_sdf_multi_output = decode_wrap
    ref_mesh.kinematic_embedding = embed_wrap

    def nerf_wrap(multires):
        fn, dim = orig_nerf(multires)

        def fn2(x):
            cap.xyz.append(x.clone().numpy())
            return fn(x)
        return fn2, dim
    ref_mesh.get_nerf_embedder = nerf_wrap
    hb, ob = case.get("hand_branch", True), case.get("obj_branch", True)
    try:
        with tempfile.TemporaryDirectory() as td, torch.no_grad():
            ref_mesh.create_mesh_combined_decoder(
                hb, ob, bool(case.get("cls_branch", False)), ref_dec, sample.latent,
                sample.mano_results, sample.obj_results, sample.cam_intr, sample.specs,
                os.path.join(td, "x"), N=N, max_batch=2 ** 18)
    finally:
        ref_mesh.get_higher_res_cube = orig_cube
        ref_mesh.decode_sdf_multi_output = orig_decode
        ref_mesh.kinematic_embedding = orig_embed
        ref_mesh.get_nerf_embedder = orig_nerf

    # pass-2 volumes arrive via the MC hook (hand first if requested, then obj)
    vols = {}
    calls = list(cap.mc_calls)
    if hb:
        vols["hand"], sp = calls.pop(0)
    if ob:
        vols["obj"], sp = calls.pop(0)
    p1h, p1o, nv, no = cap.cube
    n3 = N ** 3
    xyz_all = np.concatenate(cap.xyz, 0)
    assert xyz_all.shape[0] == 2 * n3, xyz_all.shape
    xyz1, xyz2 = xyz_all[:n3], xyz_all[n3:]

    # ---------------- pin the oracle against what the reference just did ----
    sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
    cfg = orc.decoder_cfg(mine)
    g1 = orc.grid_points(N, 2.0 / (N - 1), [-1, -1, -1]).numpy()
    assert np.array_equal(g1, xyz1), "oracle pass-1 grid is not bit-exact vs reference"
    res = orc.two_pass_field(sd, cfg, sample.latent, sample.specs, sample.mano_results,
                             sample.obj_results, N, hb, ob, cam_intr=sample.cam_intr)
    assert float(res["voxel"]) == float(nv) and np.array_equal(res["origin"].numpy(), no.numpy()), \
        "oracle re-grid parameters differ from the reference"
    g2 = orc.grid_points(N, nv, no).numpy()
    assert np.array_equal(g2, xyz2), "oracle pass-2 grid is not bit-exact vs reference"
    errs = {}
    if hb:
        errs["p1h"] = float(np.abs(res["pass1_hand"].numpy() - p1h).max())
        errs["p2h"] = float(np.abs(res["hand"].numpy() - vols["hand"]).max())
    if ob:
        errs["p1o"] = float(np.abs(res["pass1_obj"].numpy() - p1o).max())
        errs["p2o"] = float(np.abs(res["obj"].numpy() - vols["obj"]).max())
    # LayerNorm decoders amplify fp32 summation-order noise (outputs span the whole tanh range): 2e-6 there
    # non-engineered decoders with a large last-layer gain: the fp32 summation-order noise of two faithful fp32
    # evaluations grows with the output range (measured 1.0e-6 .. 1.9e-6 at gain 4, 4.5e-6 at gain 16)
    tol = 2e-6 if case.get("weight_norm") is False else 1e-6 * max(1.0, case.get("out_gain", 1.0) / 2.0)
    assert max(errs.values()) <= tol, errs
    cls = None
    if cap.logits:
        lg = np.concatenate(cap.logits, 0)
        cls = lg[n3:].argmax(1).astype(np.int32)          # pass-2 classes
        assert np.array_equal(res["cls"].numpy().reshape(-1).astype(np.int32), cls)

    meta = dict(case)
    if getattr(mine, "bias_shift", None) is not None:
        meta["bias_shift"] = [float(x) for x in mine.bias_shift]
    meta.update(latent_size=256, network_specs=ns, digest=synthetic.state_digest(mine),
                oracle_vs_reference_maxerr=errs, torch=torch.__version__,
                frac_neg_hand=float((p1h < 0).mean()) if p1h is not None else None,
                frac_neg_obj=float((p1o < 0).mean()) if p1o is not None else None,
                xyz1_sha=hashlib.sha256(xyz1.tobytes()).hexdigest(),
                xyz2_sha=hashlib.sha256(xyz2.tobytes()).hexdigest())
    out = dict(meta=np.array(json.dumps(meta)), new_voxel=np.float32(float(nv)),
               new_origin=no.numpy().astype(np.float32), spacing=np.array(sp, np.float64))
    if p1h is not None:
        out["pass1_hand"] = p1h.astype(np.float32)
    if p1o is not None:
        out["pass1_obj"] = p1o.astype(np.float32)
    for k, v in vols.items():
        out["pass2_" + k] = v.astype(np.float32)
    if cls is not None:
        out["pass2_cls"] = cls
    if case["name"] == "sep_both9_n24":
        out["xyz2"] = xyz2.astype(np.float32)               # full sheared pass-2 grid
    return out, meta


class _LegacyDecoder(torch.nn.Module):
    """DeepSDF-style single-output decoder for deep_sdf.mesh.create_mesh."""

    def __init__(self, comb):
        super().__init__()
        self.comb = comb

    def forward(self, x):
        return self.comb(x)[0]


def run_legacy_case():
    """deep_sdf/mesh.py::create_mesh (dead code in the reference, named by north_star)."""
    import deep_sdf.mesh as legacy
    import skimage.measure
    N, seed = 16, 11
    mine = synthetic.make_decoder(seed, "combined", 256, 3, "nerf")
    sample = synthetic.make_sample(seed, 256, 3, "nerf")
    cap = Capture()
    skimage.measure.marching_cubes_lewiner = cap.mc_hook
    try:
        with tempfile.TemporaryDirectory() as td, torch.no_grad():
            legacy.create_mesh(_LegacyDecoder(mine), sample.latent, os.path.join(td, "x"), N=N,
                               max_batch=32 ** 3)
    except ValueError as e:            # the legacy path does not swallow the MC exception
        assert "captured" in str(e)
    vol, sp = cap.mc_calls[0]
    sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
    cfg = orc.decoder_cfg(mine)
    with torch.no_grad():
        h, _, _ = orc.eval_volume(sd, cfg, sample.latent, sample.specs, None, None, N,
                                  2.0 / (N - 1), [-1, -1, -1])
    err = float(np.abs(h.numpy().reshape(N, N, N) - vol).max())
    assert err <= 1e-6, err
    meta = dict(name="legacy_n16", kind="combined", pf=3, style="nerf", N=N, seed=seed,
                latent_size=256, network_specs=dict(synthetic.NETWORK_SPECS),
                digest=synthetic.state_digest(mine), oracle_vs_reference_maxerr=dict(vol=err),
                torch=torch.__version__)
    out = dict(meta=np.array(json.dumps(meta)), volume=vol.astype(np.float32),
               spacing=np.array(sp, np.float64))
    return out, meta


def grid_512_fixture():
    """float(i) is inexact above 2**24 (N=512): pin the oracle's coordinates there with the
    reference's own expressions (utils/mesh.py:32-40) evaluated on index windows."""
    N = 512
    vs = 2.0 / (N - 1)
    wins = [(0, 4096), (2 ** 24 - 2048, 2 ** 24 + 2048), (N ** 3 - 4096, N ** 3),
            (100_000_007, 100_000_007 + 4096)]
    out = {}
    for a, b in wins:
        idx = torch.arange(a, b, 1, out=torch.LongTensor())
        s = torch.zeros(b - a, 3)
        s[:, 2] = idx % N
        s[:, 1] = (idx.long() / N) % N
        s[:, 0] = ((idx.long() / N) / N) % N
        s[:, 0] = (s[:, 0] * vs) + -1
        s[:, 1] = (s[:, 1] * vs) + -1
        s[:, 2] = (s[:, 2] * vs) + -1
        assert np.array_equal(s.numpy(), orc.grid_points(N, vs, [-1, -1, -1], "reference", a, b).numpy())
        out[f"win_{a}_{b}"] = s.numpy()
    return out


def main():
    install_shims()
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    index = {}
    only = set(sys.argv[1:])
    if only:                                   # regenerate just the named cases, keep the rest of the index
        with open(os.path.join(gold, "index.json")) as f:
            index = json.load(f)
        for case in CASES:
            if case["name"] in only:
                out, meta = run_reference_case(case)
                np.savez_compressed(os.path.join(gold, case["name"] + ".npz"), **out)
                index[case["name"]] = meta
                print(case["name"], meta["oracle_vs_reference_maxerr"], "neg frac", meta["frac_neg_hand"], meta["frac_neg_obj"])
        with open(os.path.join(gold, "index.json"), "w") as f:
            json.dump(index, f, indent=1, sort_keys=True)
        return
    for case in CASES:
        out, meta = run_reference_case(case)
        np.savez_compressed(os.path.join(gold, case["name"] + ".npz"), **out)
        index[case["name"]] = meta
        print(case["name"], meta["oracle_vs_reference_maxerr"],
              "neg frac", meta["frac_neg_hand"], meta["frac_neg_obj"])
    out, meta = run_legacy_case()
    np.savez_compressed(os.path.join(gold, "legacy_n16.npz"), **out)
    index["legacy_n16"] = meta
    print("legacy_n16", meta["oracle_vs_reference_maxerr"])
    np.savez_compressed(os.path.join(gold, "grid512_windows.npz"), **grid_512_fixture())
    with open(os.path.join(gold, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
