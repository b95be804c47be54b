"""Known-answer tests of the oracle for the source types beyond points and spheres
(src/sources/source_type.f90: emit_from_extern_sph :748, emit_from_extern_box :822,
emit_from_plane_parallel :935, emit_from_point_collection :570).

The reference holds no golden file for these emitters, so the oracle is pinned by physics: in an
optically thin grid the track length per unit volume is known in closed form for each of them, and
the deposit grid is  sum(path * kappa * E)  (grid_propagate_3d.f90:148-160).
"""
import numpy as np
import pytest

from helpers import pc, lsun
from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatSource


def _thin_model(sources, n=8, evenly=False):
    dust = syn.grey_dust(n_temp=10)
    w = np.linspace(-pc, pc, n + 1)
    rho = np.full((1, n, n, n), 1.e-30)      # optically thin: no attenuation, no interactions
    return FlatModel(w, w, w, rho, [dust], sources, FlatConf(sample_sources_evenly=evenly)), dust


def _track_density(model, dust, n_photons, rank=0):
    """Track length per unit volume and per emitted packet, from the deposit grid."""
    from oracle import oracle
    o = oracle.Oracle(model, rank=rank)
    o.lucy_begin()
    o.lucy_photons(n_photons)
    sums = o.get_energy_sum()[0]
    st = o.lucy_finish().as_dict()
    assert st["n_absorptions"] == 0 and st["n_scatterings"] == 0
    kappa = float(dust.chi[0] * (1.0 - dust.albedo[0]))
    assert np.allclose(dust.chi, dust.chi[0]) and np.allclose(dust.albedo, dust.albedo[0])   # grey
    vol = np.diff(model.w3)[:, None, None] * np.diff(model.w2)[None, :, None] * np.diff(model.w1)[None, None, :]
    return sums / kappa / vol / o.energy_current


def test_extern_sphere_fills_its_interior_uniformly():
    """Cosine-law emission from a sphere inwards gives a uniform isotropic field inside it: track
    length per volume and packet = mean chord / volume = 1 / (pi R^2)."""
    R = 0.99 * pc                    # the reference needs the emission points inside the grid
    src = FlatSource(type=5, luminosity=lsun, temperature=5000., position=(0., 0., 0.), radius=R)
    model, dust = _thin_model([src])
    t = _track_density(model, dust, 400000)
    x = 0.5 * (model.w1[1:] + model.w1[:-1])
    zz, yy, xx = np.meshgrid(x, x, x, indexing="ij")
    half = 0.5 * (model.w1[1] - model.w1[0])
    inside = np.sqrt(xx ** 2 + yy ** 2 + zz ** 2) + half * np.sqrt(3.0) < R
    assert inside.sum() > 100
    expect = 1.0 / (np.pi * R * R)
    rel = t[inside] / expect - 1.0
    assert abs(rel.mean()) < 0.005, rel.mean()
    assert rel.std() < 0.05, rel.std()


def test_extern_box_fills_its_interior_uniformly():
    """Same for a box emitting inwards from its six faces, chosen by area: 4 / S per volume and packet."""
    b = (-0.9 * pc, 0.7 * pc, -0.8 * pc, 0.95 * pc, -0.6 * pc, 0.9 * pc)
    src = FlatSource(type=6, luminosity=lsun, temperature=5000., bounds=b)
    model, dust = _thin_model([src])
    t = _track_density(model, dust, 400000)
    dx, dy, dz = b[1] - b[0], b[3] - b[2], b[5] - b[4]
    S = 2.0 * (dx * dy + dy * dz + dz * dx)
    lo, hi = model.w1[:-1], model.w1[1:]
    ins = lambda a0, a1: (lo >= a0) & (hi <= a1)
    inside = ins(b[4], b[5])[:, None, None] & ins(b[2], b[3])[None, :, None] & ins(b[0], b[1])[None, None, :]
    assert inside.sum() > 50
    rel = t[inside] / (4.0 / S) - 1.0
    assert abs(rel.mean()) < 0.005, rel.mean()
    assert rel.std() < 0.05, rel.std()


def test_plane_parallel_beam():
    """A beam of radius r along (theta, phi): 1 / (pi r^2) per volume and packet inside the beam,
    nothing outside, and every packet leaves through the far side."""
    r = 0.4 * pc
    src = FlatSource(type=7, luminosity=lsun, temperature=5000., position=(0.05 * pc, -0.03 * pc, -0.99 * pc),
                     radius=r, direction=(0.0, 0.0), peeloff=False)      # travelling along +z
    model, dust = _thin_model([src], n=10)
    t = _track_density(model, dust, 300000)
    x = 0.5 * (model.w1[1:] + model.w1[:-1])
    half = 0.5 * (model.w1[1] - model.w1[0])
    yy, xx = np.meshgrid(x, x, indexing="ij")
    d = np.sqrt((xx - 0.05 * pc) ** 2 + (yy + 0.03 * pc) ** 2)
    inside = d + half * np.sqrt(2.0) < r
    outside = d - half * np.sqrt(2.0) > r
    assert inside.sum() >= 4
    col = t[1:, :, :]                     # the first layer holds the starting disk
    rel = col[:, inside] / (1.0 / (np.pi * r * r)) - 1.0
    assert abs(rel.mean()) < 0.01, rel.mean()
    assert np.all(col[:, outside] == 0.0)
    # oblique beam: total track length = N x the chord of each ray, energy conserved
    src2 = FlatSource(type=7, luminosity=lsun, temperature=5000., position=(-0.5 * pc, 0., 0.), radius=0.1 * pc,
                      direction=(60.0, 30.0), peeloff=False)
    model2, dust2 = _thin_model([src2], n=10)
    t2 = _track_density(model2, dust2, 50000)
    v = np.array([np.sin(np.radians(60.)) * np.cos(np.radians(30.)), np.sin(np.radians(60.)) * np.sin(np.radians(30.)),
                  np.cos(np.radians(60.))])
    p0 = np.array([-0.5 * pc, 0., 0.])
    chord = min(((np.sign(v[k]) * pc) - p0[k]) / v[k] for k in range(3))
    vol = (model2.w1[1] - model2.w1[0]) ** 3
    assert abs(t2.sum() * vol / chord - 1.0) < 0.03


def test_point_collection_is_a_sum_of_point_sources():
    """Track density of a collection = sum_i p_i / (4 pi d_i^2), p_i = L_i / sum L."""
    pts = np.array([[-0.5 * pc, -0.5 * pc, -0.5 * pc], [0.5 * pc, 0.25 * pc, 0.5 * pc], [0.0, 0.5 * pc, -0.25 * pc]])
    lum = np.array([1.0, 3.0, 2.0]) * lsun
    src = FlatSource(type=8, temperature=4000., points=pts, points_luminosity=lum)
    model, dust = _thin_model([src], n=16)
    t = _track_density(model, dust, 600000)
    x = 0.5 * (model.w1[1:] + model.w1[:-1])
    zz, yy, xx = np.meshgrid(x, x, x, indexing="ij")
    expect = np.zeros_like(t)
    dmin = np.full(t.shape, np.inf)
    for p, l in zip(pts, lum / lum.sum()):
        d2 = (xx - p[0]) ** 2 + (yy - p[1]) ** 2 + (zz - p[2]) ** 2
        expect += l / (4.0 * np.pi * d2)
        dmin = np.minimum(dmin, np.sqrt(d2))
    far = dmin > 5.0 * (model.w1[1] - model.w1[0])      # cell size small against the distance
    assert far.sum() > 500
    rel = t[far] / expect[far] - 1.0
    assert abs(rel.mean()) < 0.01, rel.mean()
    assert rel.std() < 0.12, rel.std()


def test_mixed_sources_share_the_luminosity_pdf():
    """emit (source.f90:100-179) picks the source by luminosity (or evenly with weights): the emitted
    energy and the per-source packet shares follow sum L."""
    from oracle import oracle
    srcs = [FlatSource(type=1, luminosity=1.0 * lsun, temperature=6000., position=(0., 0., 0.)),
            FlatSource(type=5, luminosity=2.0 * lsun, temperature=5000., position=(0., 0., 0.), radius=0.9 * pc),
            FlatSource(type=6, luminosity=3.0 * lsun, temperature=4000., bounds=(-0.9 * pc, 0.9 * pc) * 3),
            FlatSource(type=8, temperature=3000., points=np.array([[0.1 * pc, 0., 0.], [0., 0.2 * pc, 0.]]),
                       points_luminosity=np.array([1.5, 2.5]) * lsun)]
    for evenly in (False, True):
        model, dust = _thin_model(srcs, evenly=evenly)
        o = oracle.Oracle(model)
        o.lucy_begin()
        o.lucy_photons(40000)
        st = o.lucy_finish().as_dict()
        assert st["n_photons"] == 40000
        # energy_current sums the weights: N exactly when sampling by luminosity, N on average when even
        assert abs(o.energy_current / 40000 - 1.0) < (1e-12 if not evenly else 0.02)


def test_map_source_far_field():
    """emit_from_map: cells drawn by luminosity, uniform position inside the cell.  Two lit cells
    with weights 1 : 3 look like two point sources from afar: sum_i p_i / (4 pi d_i^2)."""
    n = 16
    lum = np.zeros((n, n, n))
    lum[3, 4, 5] = 1.0        # [z, y, x]
    lum[12, 10, 9] = 3.0
    src = FlatSource(type=4, luminosity=lsun, temperature=4000., map=lum)
    model, dust = _thin_model([src], n=n)
    t = _track_density(model, dust, 600000)
    x = 0.5 * (model.w1[1:] + model.w1[:-1])
    zz, yy, xx = np.meshgrid(x, x, x, indexing="ij")
    expect = np.zeros_like(t)
    dmin = np.full(t.shape, np.inf)
    for (iz, iy, ix), w in (((3, 4, 5), 0.25), ((12, 10, 9), 0.75)):
        d2 = (xx - x[ix]) ** 2 + (yy - x[iy]) ** 2 + (zz - x[iz]) ** 2
        expect += w / (4.0 * np.pi * d2)
        dmin = np.minimum(dmin, np.sqrt(d2))
    far = dmin > 5.0 * (model.w1[1] - model.w1[0])
    assert far.sum() > 500
    rel = t[far] / expect[far] - 1.0
    assert abs(rel.mean()) < 0.01, rel.mean()
    assert rel.std() < 0.12, rel.std()


def test_map_source_uniform_cube_conserves_track_length():
    """A uniform map over a cube: the mean track length of a packet is the mean chord seen from a
    uniform interior point, i.e. total track length / N = (1 / V) int int ds dV -- checked against a
    direct numerical average of the distance to the boundary over positions and directions."""
    n = 8
    src = FlatSource(type=4, luminosity=lsun, temperature=4000., map=np.ones((n, n, n)))
    model, dust = _thin_model([src], n=n)
    t = _track_density(model, dust, 200000)
    vol = (model.w1[1] - model.w1[0]) ** 3
    mean_track = t.sum() * vol
    rng = np.random.default_rng(3)
    m = 400000
    p = rng.uniform(-pc, pc, (m, 3))
    mu = rng.uniform(-1, 1, m)
    ph = rng.uniform(0, 2 * np.pi, m)
    v = np.stack([np.sqrt(1 - mu ** 2) * np.cos(ph), np.sqrt(1 - mu ** 2) * np.sin(ph), mu], axis=1)
    with np.errstate(divide="ignore"):
        d = np.min(np.where(v > 0, (pc - p) / v, (-pc - p) / v), axis=1)
    assert abs(mean_track / d.mean() - 1.0) < 0.01


def test_spot_lights_only_its_cap():
    """A spherical source whose luminosity sits in one spot (source type 3, source_type.f90:150-188,
    421-427, 632-637): seen from above the spot centre the light comes from a disk of radius
    R sin(spot radius); from the opposite side nothing is seen (the 4 mu peel-off weight vanishes)."""
    from oracle import oracle
    from hyperion_b200.flatmodel import FlatPeeledGroup
    R, size = 0.4 * pc, 30.0
    # angle3d_deg(longitude, latitude) makes the reference read the pair as (theta, phi): the cap sits
    # around the direction (theta = longitude, phi = latitude)
    lon, lat = 60.0, 20.0
    star = FlatSource(type=2, luminosity=1e-12 * lsun, temperature=5000., position=(0., 0., 0.), radius=R,
                      spots=[dict(luminosity=lsun, longitude=lon, latitude=lat, radius=size, temperature=8000.)])
    model, _ = _thin_model([star], n=4)
    common = dict(wavelengths=(1, 0.01, 1000.), stokes=False, image=(40, 40, -R, R, -R, R))
    model.peeled = [FlatPeeledGroup(theta=[lon], phi=[lat], **common),
                    FlatPeeledGroup(theta=[180. - lon], phi=[lat + 180.], **common)]
    o = oracle.Oracle(model)
    o.final_begin()
    o.final_photons(40000, False)
    o.final_finish()
    front = o.image(0)[0, 0, 0, :, :, 0]
    back = o.image(1)[0, 0, 0, :, :, 0]
    x = (np.arange(40) + 0.5) / 40 * 2 * R - R
    yy, xx = np.meshgrid(x, x, indexing="ij")
    rr = np.sqrt(xx ** 2 + yy ** 2)
    pix = 2 * R / 40
    assert front.sum() > 0
    assert front[rr > R * np.sin(np.radians(size)) + pix].sum() == 0.0
    assert front[rr < R * np.sin(np.radians(size)) - pix].min() > 0.0
    assert back.sum() < 1e-9 * front.sum()


def test_additional_specific_energy_is_added_after_each_iteration():
    """specific_energy_type = 'additional' (grid_physics_3d.f90:213-235, 537-545): the iterations start from
    the minimum specific energy and the given array is added to what the packets deposit."""
    from oracle import oracle
    from helpers import bitlevel_model
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "bitlevel_car.npz"))
    base = bitlevel_model(z, False, True)
    extra = np.random.default_rng(2).uniform(0.5, 2.0, base.density.shape) * 1e-3
    plain = oracle.Oracle(base)
    plain.run_lucy_iteration(20000)
    e0 = plain.get_specific_energy()
    add = bitlevel_model(z, False, True)
    add.specific_energy = extra
    add.conf.specific_energy_additional = True
    o = oracle.Oracle(add)
    o.run_lucy_iteration(20000)
    e1 = o.get_specific_energy()
    # same random stream, same starting state (the minimum): the deposits are identical
    assert np.allclose(e1, e0 + extra, rtol=1e-13, atol=0)
