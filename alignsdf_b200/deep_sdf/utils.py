"""Drop-in for ``deep_sdf/utils.py::decode_sdf`` (deep_sdf/utils.py:64-75): single-output decoder."""
from __future__ import annotations

from .. import engine as _engine


def _legacy_specs(decoder, latent_vector, queries):
    pf = int(queries.shape[1])
    return dict(PointFeatSize=pf, EncodeStyle="nerf", SdfScaleFactor=1.0, PixelAlign=False)


def decode_sdf(decoder, latent_vector, queries):
    """sdf [P,1] = decoder(cat([latent.expand, queries])) with queries = xyz [P,3]."""
    if latent_vector is None:
        raise NotImplementedError("decode_sdf without a latent vector is not supported")
    inner = _engine.unwrap_decoder(decoder)
    eng = _engine.get_engine(inner, queries.device)
    bound = eng.bind(latent_vector, _legacy_specs(inner, latent_vector, queries), None, None)
    hand, _, _ = bound.eval_points(queries)
    return hand.unsqueeze(1)
