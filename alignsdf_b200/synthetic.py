"""Seeded synthetic decoders, latents and poses (no checkpoints / datasets ship
with the reference).

Everything here is produced from ``torch.rand`` (mt19937 -> uniform) followed
only by IEEE-exact elementwise arithmetic (+, -, *, /, sqrt), so a seed yields
bit-identical tensors on any host with the same torch build.  That is what lets
golden vectors made from the real reference in the authoring container be
replayed on the GPU box from just (seed, config, last-layer weights).

Random-init decoders have no zero crossing, so a block of
hidden units is wired to produce an ellipsoid-like level set for the hand and
for the object (``_engineer``), perturbed by the remaining random units.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from .decoders import CombinedDecoder, SeparateDecoder

SDF_SCALE_OBMAN = 7.018621123357809  # experiments/obman/*.json "SdfScaleFactor"

NETWORK_SPECS = dict(  # experiments/obman/30k_1e2d_mlp5.json:62-89
    dims=[512, 512, 512, 512], dropout=[0, 1, 2, 3], dropout_prob=0.2,
    norm_layers=[0, 1, 2, 3], latent_in=[2], num_class=6, xyz_in_all=False,
    use_tanh=False, latent_dropout=False, weight_norm=True)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def _uniform(g, *shape):
    return torch.rand(*shape, generator=g, dtype=torch.float64)


def _gauss(g, *shape):
    """Approximately N(0,1): Irwin-Hall sum of 4 uniforms, exact IEEE ops only."""
    u = torch.rand(4, *shape, generator=g, dtype=torch.float64)
    return (u[0] + u[1] + u[2] + u[3] - 2.0) * math.sqrt(3.0)


def _sumsq(t):
    """Row sum of squares with a FIXED association order (bit-reproducible)."""
    acc = t[:, 0:1] * t[:, 0:1]
    for k in range(1, t.shape[1]):
        acc = acc + t[:, k:k + 1] * t[:, k:k + 1]
    return acc


def _rigid(g, n, rot_sigma=0.35, trans_sigma=0.05):
    """n random rigid 4x4 transforms from normalised quaternions (exact ops)."""
    q = torch.cat([torch.ones(n, 1, dtype=torch.float64), rot_sigma * _gauss(g, n, 3)], 1)
    q = q / torch.sqrt(_sumsq(q))
    w, x, y, z = q.unbind(1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)
    T = torch.zeros(n, 4, 4, dtype=torch.float64)
    T[:, :3, :3] = R
    T[:, :3, 3] = trans_sigma * _gauss(g, n, 3)
    T[:, 3, 3] = 1.0
    return T.float()


@dataclass
class Sample:
    """One synthetic test image's worth of decoder conditioning."""
    latent: torch.Tensor                      # [1, L] f32
    mano_results: dict | None                 # global_trans [1,16,4,4], rot_center [1,1,3]
    obj_results: dict | None                  # obj_trans [1,4,4]
    specs: dict = field(default_factory=dict)
    cam_intr: torch.Tensor | None = None      # [1,3,4] camera projection (PixelAlign samples only)

    def to(self, device):
        mv = lambda d: None if d is None else {k: v.to(device) for k, v in d.items()}
        return Sample(self.latent.to(device), mv(self.mano_results), mv(self.obj_results), self.specs,
                      None if self.cam_intr is None else self.cam_intr.to(device))


def make_sample(seed: int, latent_size=256, point_feat_size=9, encode_style="both",
                scale_factor=SDF_SCALE_OBMAN, pixel_align=None) -> Sample:
    """``pixel_align=(fh, fw)``: a PixelAlign sample (specs['PixelAlign'], utils/utils.py:536-566) -- the latent is
    an image feature map [1, latent_size, fh, fw], with the camera intrinsics and the predicted root joint that
    project query points into it (about a quarter of the cube projects outside the image)."""
    g = _gen(1_000_003 * (seed + 1))
    latent = (0.5 * _gauss(g, 1, latent_size)).float()
    mano = dict(global_trans=_rigid(g, 16).unsqueeze(0),
                rot_center=(0.05 * _gauss(g, 1, 1, 3)).float())
    obj = dict(obj_trans=_rigid(g, 1))
    specs = dict(PointFeatSize=point_feat_size, EncodeStyle=encode_style,
                 SdfScaleFactor=scale_factor, PixelAlign=False, ImageSize=[256, 256],
                 LatentSize=latent_size)
    if encode_style == "nerf":
        mano_out, obj_out = None, None
    else:
        mano_out, obj_out = mano, obj
    if pixel_align is not None:
        fh, fw = pixel_align
        gp = _gen(3_000_017 * (seed + 1))
        # smooth-ish feature map around the per-sample latent: what the head of the image encoder would produce
        latent = (latent.double()[:, :, None, None] + 0.35 * _gauss(gp, 1, latent_size, fh, fw)).float()
        joints = (0.05 * _gauss(gp, 1, 21, 3)).double()
        joints[:, :, 2] += 0.55                                  # ~55 cm in front of the camera
        mano_out = dict(mano_out or {}, joints=joints.float())
        cam = torch.tensor([[[210.0, 0.0, 128.0, 0.0], [0.0, 210.0, 128.0, 0.0], [0.0, 0.0, 1.0, 0.0]]])
        specs = dict(specs, PixelAlign=True)
        return Sample(latent, mano_out, obj_out, specs, cam)
    return Sample(latent, mano_out, obj_out, specs)


def make_batch(n: int, base_seed: int = 0, latent_jitter: float = 0.02, **kw):
    """n samples for ONE decoder whose level sets all exist: independent poses, latents within
    ``latent_jitter`` (std per entry) of the base sample's.  Independent latents would not do for the
    non-engineered decoders: their output offset depends far more on the latent than on the query point, so
    the last-layer bias shift fitted for the base sample leaves most other samples without a zero crossing."""
    base = make_sample(base_seed, **kw)
    out = [base]
    for i in range(1, n):
        s = make_sample(base_seed + i, **kw)
        g = _gen(5_000_011 * (base_seed + i + 1))
        s.latent = (base.latent.double() + latent_jitter * _gauss(g, *base.latent.shape)).float()
        out.append(s)
    return out


def _fill_linear(g, mod, hidden_gain):
    """Overwrite one (WN)Linear with reproducible values (variance preserving)."""
    out_f, in_f = (mod.weight_v.shape if hasattr(mod, "weight_v") else mod.weight.shape)
    v = (_gauss(g, out_f, in_f) / math.sqrt(in_f)).float()
    b = (0.1 * _gauss(g, out_f)).float()
    with torch.no_grad():
        if hasattr(mod, "weight_v"):
            mod.weight_v.copy_(v)
            mod.weight_g.copy_((hidden_gain * (0.8 + 0.4 * _uniform(g, out_f, 1))).float())
        else:
            mod.weight.copy_(v * hidden_gain)
        mod.bias.copy_(b)


def _fill_default(g, mod):
    """torch's default nn.Linear initialisation as a distribution (kaiming_uniform(a=sqrt 5) ->
    U(-1/sqrt(in), 1/sqrt(in)) for weight and bias; weight_norm starts at g = ||v||, i.e. W = v), drawn
    from our own generator with exact IEEE ops only so that the decoder is bit-reproducible from the seed."""
    out_f, in_f = (mod.weight_v.shape if hasattr(mod, "weight_v") else mod.weight.shape)
    bound = 1.0 / math.sqrt(in_f)
    v = ((2.0 * _uniform(g, out_f, in_f) - 1.0) * bound).float()
    b = ((2.0 * _uniform(g, out_f) - 1.0) * bound).float()
    with torch.no_grad():
        if hasattr(mod, "weight_v"):
            mod.weight_v.copy_(v)
            mod.weight_g.copy_(torch.sqrt(_sumsq(v.double())).float())
        else:
            mod.weight.copy_(v)
        mod.bias.copy_(b)


def make_decoder(seed: int, kind="separate", latent_size=256, point_feat_size=9,
                 encode_style="both", network_specs=None, use_classifier=False,
                 init="engineered", out_gain=1.0, bias_shift=None):
    """Reproducible decoder; hidden layers roughly variance preserving.

    ``init``: "engineered" -- random hidden units + a block wired to an ellipsoid level set (_engineer:
                              numerically benign, a closed surface inside the cube);
              "plain"      -- the same random generator WITHOUT the wiring, last layer scaled by
                              ``out_gain`` (|sdf| up to ~0.1 / 0.4 / 0.95 for gain 1 / 4 / 16);
              "default"    -- torch's default initialisation, |sdf| ~ 0.02.
    The non-engineered variants get their last-layer biases shifted so that ~30 % of a coarse grid is
    negative for the sample of the same seed (random nets have no zero crossing);
    ``bias_shift`` replays recorded shifts (fixtures) instead of recomputing them (quantile + tanh are
    not bit-reproducible across hosts); the shifts used are left in ``dec.bias_shift``."""
    ns = dict(NETWORK_SPECS if network_specs is None else network_specs)
    cls = SeparateDecoder if kind == "separate" else CombinedDecoder
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(0)  # nn.Linear default init is overwritten below anyway
        dec = cls(latent_size, point_feat_size, encode_style, use_classifier=use_classifier, **ns)
    g = _gen(7_000_001 * (seed + 1))
    has_ln = False
    if init not in ("engineered", "plain", "default"):
        raise ValueError(f"unknown init {init!r}")
    for name, mod in dec.named_children():  # registration order == deterministic
        if hasattr(mod, "weight_v") or isinstance(mod, torch.nn.Linear):
            if init == "default":
                _fill_default(g, mod)
            else:
                _fill_linear(g, mod, 1.0)
        elif isinstance(mod, torch.nn.LayerNorm):
            has_ln = True
            with torch.no_grad():
                mod.weight.copy_((1.0 + 0.2 * _gauss(g, *mod.weight.shape)).float())
                mod.bias.copy_((0.1 * _gauss(g, *mod.bias.shape)).float())
    if init != "engineered":
        sep = isinstance(dec, SeparateDecoder)
        n_lin = (dec.num_hand_layers if sep else dec.num_layers) - 1
        with torch.no_grad():
            for prefix in (("linh", "lino") if sep else ("lin",)):
                getattr(dec, f"{prefix}{n_lin - 1}").weight.mul_(float(out_gain))
    if has_ln or init != "engineered":
        return _center_outputs(dec, seed, latent_size, point_feat_size, encode_style,
                               shifts=bias_shift).eval()
    return _engineer(dec, seed).eval()


def _center_outputs(dec, seed, latent_size, point_feat_size, encode_style, frac_negative=0.3, shifts=None):
    """LayerNorm decoders: the ellipsoid wiring of _engineer does not survive the normalisation, so the
    last-layer biases are shifted until ``frac_negative`` of a coarse grid (for the sample of the same
    seed) is inside -- a zero level set exists and the bbox re-grid is exercised."""
    from . import packer
    sep = isinstance(dec, SeparateDecoder)
    n_lin = (dec.num_hand_layers if sep else dec.num_layers) - 1
    if shifts is not None:                 # replay recorded shifts (fixtures)
        with torch.no_grad():
            for o, prefix in enumerate(("linh", "lino") if sep else ("lin", "lin")):
                getattr(dec, f"{prefix}{n_lin - 1}").bias[o if not sep else 0] -= float(shifts[o])
        dec.bias_shift = [float(x) for x in shifts]
        return dec
    sample = make_sample(seed, latent_size, point_feat_size, encode_style)
    ax = torch.linspace(-1, 1, 9)
    xyz = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    feats = packer.embed_points_torch(xyz, sample) if packer.uses_kinematic_embedding(sample.specs, sample.mano_results) else xyz
    x = torch.cat([sample.latent.expand(xyz.shape[0], -1), feats], 1)
    dec.eval()
    with torch.no_grad():
        out = dec(x)
        used = []
        for o, prefix in enumerate(("linh", "lino") if sep else ("lin", "lin")):
            pre = torch.atanh(out[o][:, 0].double().clamp(-0.999999, 0.999999))
            shift = torch.quantile(pre, frac_negative).float()
            last = getattr(dec, f"{prefix}{n_lin - 1}")
            last.bias[o if not sep else 0] -= shift
            used.append(float(shift))
    dec.bias_shift = used
    return dec


# ----------------------------------------------------------------------------
# engineered level sets
# ----------------------------------------------------------------------------
# A random-init decoder has no zero crossing and a
# least-squares fit of its last layer is so ill-conditioned that the reference's
# own fp32 forward is only reproducible to ~2e-5.  Instead a block of J hidden
# units is wired, through the layer-2 skip connection, to compute
#   u_j = relu(d_j . S (p + k - c))         (d_j random unit directions)
# which is passed through layer 3 unchanged and summed by the last layer with
# equal positive weights: (4/J) sum_j u_j ~ |S(p-c)|, i.e. the output is
# ~ amp * (|S(p-c)| - 1), the SDF of an ellipsoid with semi-axes 1/S, perturbed
# by the remaining random units.  No cancellation -> benign fp32 numerics, a
# closed surface well inside the cube, and everything stays bit-reproducible.
_BLOBS = dict(hand=dict(c=(-0.05, 0.05, 0.0), axes=(0.34, 0.50, 0.22)),
              obj=dict(c=(0.30, -0.16, 0.26), axes=(0.24, 0.18, 0.15)))
_AMP = 0.15
_NOISE = 0.004


def _unit_dirs(g, n):
    d = _gauss(g, n, 3)
    return d / torch.sqrt(_sumsq(d))


def _set_row(mod, rows, W, b, nrm):
    """Write effective weight rows (W [r,in], row norms nrm [r,1]) + bias into a (WN)Linear."""
    with torch.no_grad():
        if hasattr(mod, "weight_v"):
            mod.weight_v[rows] = (W / nrm).float()
            mod.weight_g[rows] = nrm.float()
        else:
            mod.weight[rows] = W.float()
        mod.bias[rows] = b.float()


def _engineer(dec, seed):
    g = _gen(9_000_011 * (seed + 1))
    sep = isinstance(dec, SeparateDecoder)
    n_lin = (dec.num_hand_layers if sep else dec.num_layers) - 1
    skips = [l for l in dec.latent_in if 1 <= l <= n_lin - 2]
    if not skips:
        return dec          # no skip connection to wire through: plain random net
    ls = skips[0]
    L = dec.latent_size
    plan = ([("linh", ["hand"]), ("lino", ["obj"])] if sep else [("lin", ["hand", "obj"])])
    for prefix, outs in plan:
        J = 128 // len(outs)
        skip = getattr(dec, f"{prefix}{ls}")
        in_skip = (skip.weight_v if hasattr(skip, "weight_v") else skip.weight).shape[1]
        d0 = (getattr(dec, f"{prefix}0").weight_v if hasattr(getattr(dec, f"{prefix}0"), "weight_v")
              else getattr(dec, f"{prefix}0").weight).shape[1]
        h = in_skip - d0
        last = getattr(dec, f"{prefix}{n_lin - 1}")
        with torch.no_grad():
            last.weight.copy_((_NOISE * _gauss(g, *last.weight.shape)).float())
            last.bias.zero_()
        for o, tag in enumerate(outs):
            rows = torch.arange(o * J, (o + 1) * J)
            S = torch.tensor([1.0 / a for a in _BLOBS[tag]["axes"]], dtype=torch.float64)
            c = torch.tensor(_BLOBS[tag]["c"], dtype=torch.float64)
            dS = _unit_dirs(g, J) * S
            W = torch.zeros(J, in_skip, dtype=torch.float64)
            W[:, h + L:h + L + 3] = dS                 # first three point features ~ p + const
            _set_row(skip, rows, W, -(dS[:, 0] * c[0] + dS[:, 1] * c[1] + dS[:, 2] * c[2]),
                     torch.sqrt(_sumsq(dS)))
            for l in range(ls + 1, n_lin - 1):          # identity pass-through
                mod = getattr(dec, f"{prefix}{l}")
                in_l = (mod.weight_v if hasattr(mod, "weight_v") else mod.weight).shape[1]
                I = torch.zeros(J, in_l, dtype=torch.float64)
                I[torch.arange(J), rows] = 1.0
                _set_row(mod, rows, I, torch.zeros(J, dtype=torch.float64),
                         torch.ones(J, 1, dtype=torch.float64))
            with torch.no_grad():
                last.weight[o if len(outs) > 1 else 0, rows] = float(_AMP * 4.0 / J)
                last.bias[o if len(outs) > 1 else 0] = -_AMP
    return dec


def state_digest(dec) -> str:
    """sha256 over the state dict: fixtures use it to prove bit-reproducibility."""
    import hashlib
    h = hashlib.sha256()
    for k, v in sorted(dec.state_dict().items()):
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
