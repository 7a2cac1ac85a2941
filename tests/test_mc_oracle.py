"""Marching cubes: generated tables vs independent oracle tracing, and mesh invariants."""
import numpy as np
import pytest

from alignsdf_b200 import mc_tables
from oracle import mc_oracle as mo
from tests.mc_table_sim import table_mc


def _grid(n):
    ax = np.linspace(-1, 1, n, dtype=np.float32)
    return np.meshgrid(ax, ax, ax, indexing="ij")


def sphere(n, r=0.6, c=(0, 0, 0)):
    x, y, z = _grid(n)
    return np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - np.float32(r)


def torus(n, R=0.55, r=0.22):
    x, y, z = _grid(n)
    return np.sqrt((np.sqrt(x ** 2 + y ** 2) - R) ** 2 + z ** 2) - np.float32(r)


def test_tables_cover_all_cases():
    t = mc_tables.build_tables()
    assert t["var_offset"].shape == (256,)
    assert t["n_tris"][t["var_offset"][0]] == 0 and t["n_tris"][t["var_offset"][255]] == 0
    assert int(t["n_tris"].max()) <= mc_tables.MAX_TRIS
    for c in range(256):
        n_amb = bin(int(t["amb_mask"][c])).count("1")
        nxt = t["var_offset"][c + 1] if c < 255 else len(t["n_tris"])
        assert nxt - t["var_offset"][c] == 1 << n_amb


def test_committed_header_matches_generator():
    with open(mc_tables.header_path()) as f:
        assert f.read() == mc_tables.header_text()


@pytest.mark.parametrize("seed", [0, 1])
def test_tables_equal_oracle_tracing_on_noise(seed):
    """White noise hits every sign configuration and decider variant."""
    rng = np.random.default_rng(seed)
    vol = rng.standard_normal((9, 10, 11)).astype(np.float32)
    verts, faces, keys = mo.marching_cubes(vol, 0.0)
    k2, f2 = table_mc(vol)
    assert np.array_equal(keys, k2)
    assert np.array_equal(faces, f2)
    inv = mo.mesh_invariants(verts, faces)
    assert inv["nonmanifold_edges"] == 0
    assert inv["oriented"] or inv["boundary_edges"] > 0   # open only at the volume boundary


def test_noise_interior_is_watertight():
    """Pad noise with a positive shell: the surface cannot reach the boundary -> closed."""
    rng = np.random.default_rng(5)
    vol = np.full((14, 14, 14), 1.0, np.float32)
    vol[1:-1, 1:-1, 1:-1] = rng.standard_normal((12, 12, 12)).astype(np.float32)
    verts, faces, _ = mo.marching_cubes(vol, 0.0)
    inv = mo.mesh_invariants(verts, faces)
    assert inv["closed"] and inv["oriented"] and inv["nonmanifold_edges"] == 0


def test_sphere_invariants_and_orientation():
    n = 24
    vol = sphere(n)
    sp = 2.0 / (n - 1)
    verts, faces, keys = mo.marching_cubes(vol, 0.0, spacing=[sp] * 3)
    inv = mo.mesh_invariants(verts, faces)
    assert inv["closed"] and inv["oriented"] and inv["euler"] == 2 and inv["n_components"] == 1
    p = verts.astype(np.float64) - 1.0            # back to [-1,1] coordinates
    r = np.linalg.norm(p, axis=1)
    assert np.abs(r - 0.6).max() < 0.01           # vertices sit on the level set
    a, b, c = p[faces[:, 0]], p[faces[:, 1]], p[faces[:, 2]]
    nrm = np.cross(b - a, c - a)
    assert np.all((nrm * (a + b + c)).sum(1) > 0)  # outward (towards increasing value)
    assert abs(mo.face_areas(p, faces).sum() - 4 * np.pi * 0.36) < 0.05
    assert np.all(np.diff(keys.astype(np.int64)) > 0)


def test_torus_and_two_spheres_topology():
    v, f, _ = mo.marching_cubes(torus(32), 0.0)
    inv = mo.mesh_invariants(v, f)
    assert inv["closed"] and inv["euler"] == 0 and inv["n_components"] == 1
    two = np.minimum(sphere(32, 0.35, (-0.45, 0, 0)), sphere(32, 0.25, (0.5, 0.1, 0)))
    v, f, _ = mo.marching_cubes(two, 0.0)
    inv = mo.mesh_invariants(v, f)
    assert inv["closed"] and inv["euler"] == 4 and inv["n_components"] == 2
    v2, f2 = mo.largest_component_if_split(v, f)
    inv2 = mo.mesh_invariants(v2, f2)
    assert inv2["n_components"] == 1 and inv2["euler"] == 2
    assert np.abs(np.linalg.norm(v2.astype(np.float64) * (2 / 31) - 1.0 - np.array([-0.45, 0, 0]), axis=1)
                  - 0.35).max() < 0.02


def test_level_out_of_range_raises():
    with pytest.raises(ValueError, match="Surface level must be within volume data range"):
        mo.marching_cubes(np.ones((4, 4, 4), np.float32), 0.0)


def test_slab_keys_and_coordinates_are_global():
    vol = sphere(20)
    v, f, k = mo.marching_cubes(vol, 0.0, spacing=[0.1] * 3)
    # mesh the two z-slabs [0,10] and [10,19] separately (shared plane 10) and merge by key
    va, fa, ka = mo.marching_cubes(vol[:11], 0.0, [0.1] * 3, (0, 0, 0), vol.shape)
    vb, fb, kb = mo.marching_cubes(vol[10:], 0.0, [0.1] * 3, (10, 0, 0), vol.shape)
    keys = np.concatenate([ka, kb]); verts = np.concatenate([va, vb])
    uk, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    faces = np.concatenate([inv[fa], inv[fb + len(ka)]])
    assert np.array_equal(uk, k)
    assert np.array_equal(verts[first], v)
    assert np.array_equal(faces.astype(np.int32), f)
