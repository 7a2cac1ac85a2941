"""Tensor-core (tcgen05) path on the GPU: device-side bind and re-grid, both precision kinds against the
reference's golden fields (engineered AND non-engineered decoders), kernel selection by calibration, batched
launches, CombinedDecoder, full-size grids against the exact-fp32 kernel."""
import numpy as np
import pytest
import torch

from alignsdf_b200 import _lib, engine, mesh as amesh, packer, synthetic, tc_pack
from tests import helpers
from tests.tc_emulate import emulate

pytestmark = pytest.mark.gpu
TOL = 1e-5
ENGINEERED = ["sep_both9_n24", "sep_nerf3_n16", "sep_hand51_n16", "sep_hand6_n12", "sep_obj6_n12",
              "sep_both54_n12", "sep_both9_n32_handonly", "sep_both9_n20_objonly"]
PLAIN = ["sep_plain_g1_n32", "sep_plain_g4_n32", "sep_plain_g16_n32", "sep_default_n32"]
COMBINED = ["comb_both9_n16", "comb_plain_g4_n24", "comb_default_n16"]
DEV = torch.device("cuda")


def _check_fields(vols, g, tol=TOL):
    assert np.float32(float(vols["voxel"])) == g["new_voxel"]
    assert np.array_equal(vols["origin"].numpy(), g["new_origin"])
    worst = 0.0
    for key, vol in (("pass1_hand", vols["pass1_hand"]), ("pass1_obj", vols["pass1_obj"]),
                     ("pass2_hand", vols["hand"]), ("pass2_obj", vols["obj"])):
        if key in g:
            err = float(np.abs(vol.cpu().numpy() - g[key]).max())
            assert err <= tol, (key, err)
            worst = max(worst, err)
    return worst


def _volumes(name, path):
    meta, g, dec, sample = helpers.load_case(name)
    s = helpers.to_cuda(sample)
    hb, ob = meta.get("hand_branch", True), meta.get("obj_branch", True)
    vols = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, meta["N"], hb, ob, path=path)
    return vols, g, dec


@pytest.mark.parametrize("name", ["sep_both9_n24", "sep_hand51_n16", "sep_plain_g16_n32", "comb_both9_n16"])
@pytest.mark.parametrize("kind", [tc_pack.F16X3, tc_pack.F16_F8])
def test_device_bind_matches_the_numpy_statement(name, kind):
    """asdf_tc_bind (fold in float64, scale choice, fp16 split, swizzled tiles -- all on the device) against
    tc_pack.pack_sample_numpy on the host fold: same scales, same hi halves, hi + lo equal to double-rounding noise."""
    meta, g, dec, sample = helpers.load_case(name)
    s = helpers.to_cuda(sample)
    eng = engine.get_engine(dec, DEV)
    bound = eng.bind(s.latent, s.specs, s.mano_results, s.obj_results)
    blocks, status = bound.tc_blocks(kind, 2.0)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    got = blocks[0].cpu().numpy()
    br = packer.fold_decoder(eng.topo, sample.latent, sample.specs, sample.mano_results, sample.obj_results)
    want, info = tc_pack.pack_sample_numpy(br, eng.w_scale, 2.0 * 1.01, kind)
    nd = eng.n_dec
    assert np.array_equal(got[tc_pack.SAMPLE_TILE_BYTES:].view(np.float32)[:4][[0, 2, 3] if nd == 1 else [0, 1, 2, 3]],
                          want[tc_pack.SAMPLE_TILE_BYTES:].view(np.float32)[:4][[0, 2, 3] if nd == 1 else [0, 1, 2, 3]]), info
    gt = got[:tc_pack.SAMPLE_TILE_BYTES].view(np.float16).reshape(2, 2, tc_pack.P_TILES, tc_pack.TILE_ELEMS)
    wt = want[:tc_pack.SAMPLE_TILE_BYTES].view(np.float16).reshape(2, 2, tc_pack.P_TILES, tc_pack.TILE_ELEMS)
    for d in range(nd):
        for c in range(2):
            for t in range(tc_pack.P_TILES):
                a = tc_pack.unswizzle_tile(gt[d, c, t]).astype(np.float64)[:, :16]
                b = tc_pack.unswizzle_tile(wt[d, c, t]).astype(np.float64)[:, :16]
                va, vb = a[:, 0:4] + a[:, 8:12], b[:, 0:4] + b[:, 8:12]          # hi + lo of (M, B)
                assert np.array_equal(a[:, 4:7], a[:, 0:3]) and not a[:, 7].any() and not a[:, 12:].any()
                # one fp16 ulp of the lo half (the two float64 folds differ in summation order); subnormal lo: 2^-24
                assert (np.abs(va - vb) <= 5e-7 * np.abs(vb) + 1.3e-7).all(), (d, c, t)


def test_regrid_kernel_is_bit_equal_to_the_host_arithmetic():
    """asdf_regrid vs mesh._regrid (torch f32 ops, pinned to the reference by the golden fixtures): random boxes,
    empty branches, single branches."""
    rng = np.random.default_rng(0)
    for N in (24, 128, 256, 512):
        S = 64
        lo = rng.integers(0, N // 2, (S, 2, 3))
        hi = lo + rng.integers(0, N // 2, (S, 2, 3))
        box = np.concatenate([lo, hi], 2).reshape(S, 12).astype(np.int32)
        box[::7, 0:3], box[::7, 3:6] = engine.INT_MAX, -1           # empty hand branch
        box[::5, 6:9], box[::5, 9:12] = engine.INT_MAX, -1          # empty object branch
        box_d = torch.from_numpy(box).to(DEV)
        for mask in (1, 2, 3):
            grid = torch.empty((S, 4), dtype=torch.float32, device=DEV)
            mm = torch.empty((S, 6), dtype=torch.float32, device=DEV)
            _lib.check(_lib.lib().asdf_regrid(_lib.ptr(box_d), S, mask, N, float(np.float32(2.0 / (N - 1))),
                                              _lib.ptr(grid), _lib.ptr(mm), _lib.stream_ptr(DEV)), "asdf_regrid")
            got = grid.cpu()
            for i in range(S):
                mn, mx = amesh._bbox_to_minmax(torch.from_numpy(box[i]), bool(mask & 1), bool(mask & 2))
                v, o = amesh._regrid(mn, mx, N, 2.0 / (N - 1))
                assert float(v) == float(got[i, 0]) and torch.equal(o, got[i, 1:4]), (N, mask, i)


@pytest.mark.parametrize("name", ENGINEERED + ["sep_plain_g1_n32", "sep_plain_g4_n32", "sep_default_n32"])
def test_f16x3_two_pass_fields_match_reference_golden(name):
    """All-fp16 split precision: inside the contract up to |sdf| ~ 0.5 with margin, engineered decoder or not.
    (At last-layer gain 16 it reaches 1.0e-5 .. 1.6e-5 -- two faithful fp32 evaluations differ by 4.5e-6 there --
    which is why "auto" checks it against the fp32 kernel per sample: test_auto_path_*.)"""
    vols, g, _ = _volumes(name, "f16")
    assert vols["bound"].kinds_used == {"f16x3"}
    worst = _check_fields(vols, g)
    print(f"F16X3 {name}: max |sdf - reference| = {worst:.2e}")


@pytest.mark.parametrize("name", ENGINEERED + ["sep_default_n32"])
def test_f16_f8_two_pass_fields_match_reference_golden(name):
    """fp16 main product + e4m3 corrections on the decoders it is valid for: <= 1e-5 with margin, no range flag."""
    before = dict(engine.STATS)
    vols, g, _ = _volumes(name, "f8")
    assert engine.STATS["f8_rejected"] == before["f8_rejected"] and vols["bound"].kinds_used == {"f16+2xe4m3"}
    worst = _check_fields(vols, g)
    assert worst <= 6e-6, worst                     # the margin the design relies on (emulated: 2.3e-6)


@pytest.mark.parametrize("name", ENGINEERED + PLAIN)
def test_auto_path_picks_a_kernel_inside_the_contract(name):
    """"auto" calibrates EVERY sample -- exact-fp32 kernel vs the tensor-core kinds on 4 k random
    points, compared on the device -- and uses the fastest kind that agrees to CALIB_TOL (else the fp32 kernel
    itself); whatever it picks is <= 1e-5 from the real reference, engineered decoder or not."""
    vols, g, dec = _volumes(name, "auto")
    bound = vols["bound"]
    eng = bound.engine
    worst = _check_fields(vols, g)
    e8, e16 = bound.calib_err
    print(f"auto {name}: {bound.last_kind}, calibration f8 {e8:.2e} f16 {e16:.2e}, max |sdf - reference| = {worst:.2e}")
    assert eng.calib["samples"] >= 1
    want = "f16+2xe4m3" if e8 <= engine.CALIB_TOL else ("f16x3" if e16 <= engine.CALIB_TOL else "simt")
    assert bound.last_kind == want and eng.level == {"f16+2xe4m3": 0, "f16x3": 1, "simt": 2}[want]
    if name in ENGINEERED + ["sep_default_n32"]:
        assert want == "f16+2xe4m3"                 # benign / default-initialised decoders keep the fast kind
    if name.startswith("sep_plain"):                # ~1e-4 x range with e4m3 corrections: rejected, sticky
        assert want != "f16+2xe4m3" and e8 > engine.CALIB_TOL
        meta, _, _, sample = helpers.load_case(name)
        s = helpers.to_cuda(sample)
        again = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, meta["N"])
        assert again["bound"].last_kind == want     # no second attempt at the rejected kind
        assert torch.equal(again["hand"], vols["hand"]) and torch.equal(again["obj"], vols["obj"])


@pytest.mark.parametrize("kind,path,tol", [(tc_pack.F16X3, "f16", 3e-6), (tc_pack.F16_F8, "f8", 6e-6)])
def test_kernel_matches_fp32_kernel_emulator_and_is_deterministic(kind, path, tol):
    meta, g, dec, sample = helpers.load_case("sep_both9_n24")
    s = helpers.to_cuda(sample)
    eng = engine.get_engine(dec, DEV)
    bound = eng.bind(s.latent, s.specs, s.mano_results, s.obj_results)
    N = 48                                    # 110592 points: 432 tiles of 256, > 5 waves of 74 CTA pairs
    ht, ot, _, bt = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path=path)
    hs, os_, _, bs = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="simt")
    assert (ht - hs).abs().max() <= tol and (ot - os_).abs().max() <= tol
    assert torch.equal(bt, bs)
    ht2, ot2, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path=path)
    assert torch.equal(ht, ht2) and torch.equal(ot, ot2)
    # ragged ranges: not a multiple of the 256-point tile, odd begin
    h3, o3, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], begin=77, end=77 + 1001, path=path)
    assert torch.equal(h3, ht[77:77 + 1001]) and torch.equal(o3, ot[77:77 + 1001])
    # explicit points
    xyz = (torch.rand(777, 3, generator=torch.Generator().manual_seed(2)) * 2 - 1).cuda()
    hp, op, _ = bound.eval_points(xyz, path=path)
    hq, oq, _ = bound.eval_points(xyz, path="simt")
    assert (hp - hq).abs().max() <= tol and (op - oq).abs().max() <= tol
    # the CPU emulation of the packed bytes predicts the kernel to fp32-accumulation-order noise
    blocks, _ = bound.tc_blocks(kind, 2.0)
    he, oe = emulate(eng.tc_static(kind).cpu().numpy(), blocks[0].cpu().numpy(), xyz.cpu().numpy(), kind)
    assert np.abs(he - hp.cpu().numpy()).max() <= 1e-6 and np.abs(oe - op.cpu().numpy()).max() <= 1e-6
    far = xyz * 7.0                            # outside the default point-operand range: re-bound automatically
    hp, op, _ = bound.eval_points(far, path=path)
    hq, oq, _ = bound.eval_points(far, path="simt")
    assert (hp - hq).abs().max() <= 1e-5 and (op - oq).abs().max() <= 1e-5


@pytest.mark.parametrize("path", ["f16", "f8"])
def test_batched_launch_is_bit_identical_to_single_sample_launches(path):
    """One launch over S samples (per-sample P tiles, per-sample re-grid read from device memory, per-sample
    boxes) == S single-sample launches, bit for bit; no host round trip between the passes."""
    dec = synthetic.make_decoder(0)
    eng = engine.get_engine(dec, DEV)
    samples = [synthetic.make_sample(i).to(DEV) for i in range(5)]
    N = 20                                    # 8000 points: 32 tiles per sample, 160 items over 74 pairs (mixed samples per pair)
    kind = tc_pack.F16X3 if path == "f16" else tc_pack.F16_F8
    batch = eng.bind_batch([(s.latent, s.specs, s.mano_results, s.obj_results) for s in samples])
    lvl = engine.LEVEL_F16 if path == "f16" else engine.LEVEL_F8
    r = batch.two_pass(N, 3, "reference", lvl, keep_pass1=True)
    assert batch.verify() <= lvl
    for i, s in enumerate(samples):
        one = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, N, path=path)
        assert torch.equal(r["grid"][i].cpu(), torch.cat([one["voxel"].reshape(1), one["origin"]])), i
        assert torch.equal(r["pass1_hand"][i], one["pass1_hand"].reshape(-1)) and torch.equal(r["pass1_obj"][i], one["pass1_obj"].reshape(-1))
        assert torch.equal(r["hand"][i], one["hand"].reshape(-1)) and torch.equal(r["obj"][i], one["obj"].reshape(-1)), i
    # bounding-box-only pass 1 (what create_mesh_combined_decoder runs) gives the same lattice
    r2 = batch.two_pass(N, 3, "reference", lvl, keep_pass1=False)
    assert r2["pass1_hand"] is None and torch.equal(r2["grid"], r["grid"]) and torch.equal(r2["hand"], r["hand"])


@pytest.mark.parametrize("name", COMBINED)
@pytest.mark.parametrize("path", ["f16", "auto"])
def test_combined_decoder_on_the_tensor_core_kernel(name, path):
    """CombinedDecoder (one MLP, two outputs) through k1_tc: golden fields of the real reference."""
    vols, g, dec = _volumes(name, path)
    worst = _check_fields(vols, g)
    print(f"combined {name} [{path}]: {vols['bound'].last_kind}, max |sdf - reference| = {worst:.2e}")
    assert vols["bound"].last_kind in (("f16x3",) if path == "f16" else ("f16x3", "f16+2xe4m3"))


def test_f16_f8_falls_back_when_activations_leave_the_fp8_range():
    """Scaling layer 0 up / layer 1 down by 2^9 keeps the function but pushes x1 beyond 448: the
    kernel must raise its status flag and the engine must re-run the query through the all-fp16 kind."""
    dec = synthetic.make_decoder(3)
    with torch.no_grad():
        for p in ("linh", "lino"):
            l0, l1 = getattr(dec, p + "0"), getattr(dec, p + "1")
            l0.weight_g.mul_(512.0); l0.bias.mul_(512.0)
            l1.weight_g.mul_(1.0 / 512.0)
    s = synthetic.make_sample(3).to(DEV)
    bound = engine.get_engine(dec, DEV).bind(s.latent, s.specs, s.mano_results, s.obj_results)
    N = 24
    before = engine.STATS["f8_rejected"]
    ht, ot, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="f8")
    assert engine.STATS["f8_rejected"] == before + 1 and bound.last_kind == "f16x3"
    hs, os_, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="simt")
    assert (ht - hs).abs().max() <= 1e-5 and (ot - os_).abs().max() <= 1e-5


@pytest.mark.parametrize("init,gain,N", [("default", 1.0, 256), ("plain", 4.0, 256), ("engineered", 1.0, 128),
                                         ("plain", 16.0, 128)])
def test_full_grid_auto_path_within_contract_of_the_fp32_kernel(init, gain, N):
    """BASELINE's sizes: every point of the N^3 pass-1 grid, the kernel "auto" picks vs the exact-fp32 generic
    kernel (itself pinned to the reference's golden fields): <= 1e-5 everywhere, same bounding box."""
    dec = synthetic.make_decoder(31, init=init, out_gain=gain)
    s = synthetic.make_sample(31).to(DEV)
    bound = engine.get_engine(dec, DEV).bind(s.latent, s.specs, s.mano_results, s.obj_results)
    ht, ot, _, bt = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="auto")
    kind = bound.last_kind
    hs, os_, _, bs = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], bbox_mask=3, path="simt")
    eh, eo = float((ht - hs).abs().max()), float((ot - os_).abs().max())
    print(f"{init} gain {gain} N={N}: {kind}, max |sdf - fp32 kernel| = {max(eh, eo):.2e}")
    assert eh <= TOL and eo <= TOL, (eh, eo, kind)
    assert (kind == "f16+2xe4m3") == (init != "plain"), kind
    assert max(eh, eo) <= 2.5 * engine.CALIB_TOL        # 16.7 M points vs the 4 k calibration points: the tail stays close
    # the box may differ only where a value sits within the kernels' rounding of zero
    near0 = ((hs.abs() < 2e-5).sum() + (os_.abs() < 2e-5).sum()).item()
    assert torch.equal(bt, bs) or near0 > 0


def test_full_512_grid_against_the_fp32_kernel():
    """Config #5's size (float(i) inexact above 2^24): 134 M points, auto path vs exact-fp32 kernel."""
    dec = synthetic.make_decoder(0, init="default")
    s = synthetic.make_sample(0).to(DEV)
    bound = engine.get_engine(dec, DEV).bind(s.latent, s.specs, s.mano_results, s.obj_results)
    N = 512
    ht, ot, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], path="auto")
    worst = 0.0
    step = 2 ** 24                            # the fp32 kernel in 8 slices (keeps memory at two volumes)
    for a in range(0, N ** 3, step):
        hs, os_, _, _ = bound.eval_grid(N, 2.0 / (N - 1), [-1, -1, -1], begin=a, end=min(a + step, N ** 3), path="simt")
        worst = max(worst, float((ht[a:a + step] - hs).abs().max()), float((ot[a:a + step] - os_).abs().max()))
    assert worst <= TOL, worst


# ----------------------------------------------------------------------------------------------------------------
# pass 1 on the single-product kind (bounding boxes only) + exact re-evaluation of the shell around the surface
# ----------------------------------------------------------------------------------------------------------------
def _exact_box(bound, N, kind, mask=3, begin=0, end=None):
    end = N ** 3 if end is None else end
    q = engine.make_query(_lib.QUERY_GRID_REFERENCE, N, begin, end, 2.0 / (N - 1), (-1.0, -1.0, -1.0), bbox_mask=mask)
    box = engine.new_bbox(DEV, bound.S)
    bound.launch_tc(kind, q, end - begin, False, box)
    return q, box


@pytest.mark.parametrize("kind_name,init,gain,N", [("separate", "default", 1.0, 96), ("separate", "plain", 4.0, 64),
                                                   ("separate", "engineered", 1.0, 80), ("combined", "plain", 4.0, 64)])
def test_fast_bounding_box_pass_equals_the_exact_pass(kind_name, init, gain, N):
    dec = synthetic.make_decoder(41, kind_name, init=init, out_gain=gain)
    smp = [s.to(DEV) for s in synthetic.make_batch(2, base_seed=41)]
    eng = engine.get_engine(dec, DEV)
    bound = eng.bind_batch([(s.latent, s.specs, s.mano_results, s.obj_results) for s in smp])
    bound._calibrate()
    assert bound.verify() < engine.LEVEL_SIMT
    e1 = eng.calib["f1"]
    assert e1 is not None and 0 < e1 < 5e-3
    kind = engine.LEVEL_KIND[eng.level]
    q, want = _exact_box(bound, N, kind)
    before = dict(engine.STATS)
    for tau in (eng.fast_tau(), 2 * eng.fast_tau()):                      # a wider shell must give the same boxes
        box = engine.new_bbox(DEV, bound.S)
        bound.fast_bbox_pass(kind, q, N ** 3, box, tau)
        assert torch.equal(box, want), (tau, box.tolist(), want.tolist())
    assert engine.STATS["fast_bbox_passes"] == before["fast_bbox_passes"] + 2
    assert engine.STATS["fast_bbox_ambiguous"] > before["fast_bbox_ambiguous"]
    assert engine.STATS["fast_bbox_redone"] == before["fast_bbox_redone"]
    # a sub-range of the grid (z-slab) and a single branch
    q2, want2 = _exact_box(bound, N, kind, mask=2, begin=5 * N * N + 7, end=31 * N * N)
    box = engine.new_bbox(DEV, bound.S)
    bound.fast_bbox_pass(kind, q2, 31 * N * N - (5 * N * N + 7), box, eng.fast_tau())
    assert torch.equal(box, want2)
    # a threshold so large that the list overflows: the pass falls back to the exact kind
    box = engine.new_bbox(DEV, bound.S)
    bound.fast_bbox_pass(kind, q, N ** 3, box, 2.0)
    assert torch.equal(box, want) and engine.STATS["fast_bbox_redone"] == before["fast_bbox_redone"] + 1


def test_fast_bounding_box_threshold_is_checked_per_sample(tmp_path):
    """The drop-in call uses the fast pass from a decoder's second sample on; a threshold that this sample's own
    calibration does not support voids the pass (it is repeated exactly) -- the result never depends on it."""
    N = 64
    dec = synthetic.make_decoder(43, init="default")
    smp = [s.to(DEV) for s in synthetic.make_batch(3, base_seed=43)]
    eng = engine.get_engine(dec, DEV)
    ref = []
    try:
        engine.FAST_BBOX = False
        dec2 = synthetic.make_decoder(43, init="default")
        for s in smp:
            v = amesh.sdf_volumes(dec2, s.latent, s.mano_results, s.obj_results, s.specs, N, keep_pass1=False)
            ref.append((float(v["voxel"]), v["origin"].clone(), v["hand"].clone()))
    finally:
        engine.FAST_BBOX = True
    before = dict(engine.STATS)
    for k, s in enumerate(smp):
        if k == 2:
            eng.calib["f1"] = eng.calib["f1"] * 1e-3                      # pretend earlier samples were far more benign
        v = amesh.sdf_volumes(dec, s.latent, s.mano_results, s.obj_results, s.specs, N, keep_pass1=False)
        assert float(v["voxel"]) == ref[k][0] and torch.equal(v["origin"], ref[k][1]) and torch.equal(v["hand"], ref[k][2])
        assert v["bound"].redo_fast == (k == 2)
    assert engine.STATS["fast_bbox_passes"] == before["fast_bbox_passes"] + 2          # samples 1 and 2 (0 has no bound yet)
