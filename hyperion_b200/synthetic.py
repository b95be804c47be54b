"""Synthetic inputs for tests and benchmarks (no HDF5, no astropy needed).

Produces the same kind of tables the reference's Python front end writes into a
dust file, following the published recipe:

* isotropic / Henyey-Greenstein scattering matrices
  (``hyperion/dust/dust_type.py:35-40,525-585``),
* Planck / reciprocal-Planck / Rosseland mean opacities and
  ``specific_energy = 4 sigma T^4 kappa_P`` (``hyperion/dust/mean_opacities.py:30-110``),
* LTE emissivities ``j_nu = kappa_nu B_nu(T)`` on the merged frequency grid
  (``hyperion/dust/emissivities.py:33-65``, ``hyperion/util/functions.py:111-152``).

and the synthetic grids BASELINE.json names.
"""
from __future__ import annotations

import numpy as np

from .flatmodel import FlatConf, FlatDust, FlatModel, FlatSource

# cgs constants, hyperion/util/constants.py
h = 6.626068e-27
k = 1.3806503e-16
c = 2.99792458e10
sigma = 5.67051e-5
pc = 3.08568025e18
lsun = 3.846e33
rsun = 6.95508e10
au = 1.49598e13


def planck_nu_range(tmin, tmax):
    alpha = 2.821439
    nu_min = np.log10(alpha / h * k * tmin / 100.)
    nu_max = np.log10(alpha / h * k * tmax * 10.)
    n_nu = int((nu_max - nu_min) * 100.)
    return np.logspace(nu_min, nu_max, n_nu)


def nu_common(nu1, nu2):
    nu = np.sort(np.hstack([nu1, nu2]))
    keep = np.hstack([(nu[1:] - nu[:-1]) / nu[:-1] > 1.e-10, True])
    return nu[keep]


def B_nu(nu, T):
    x = h * nu / k / T
    f = np.zeros(nu.shape)
    main = (1.e-8 <= x) & (x < 700.)
    f[main] = 2. * h * nu[main] ** 3. / c ** 2. / np.expm1(x[main])
    small = x < 1.e-8
    f[small] = 2. * h * nu[small] ** 3. / c ** 2. / x[small]
    return f


def dB_nu_dT(nu, T):
    b = B_nu(nu, T)
    x = h * nu / k / T
    f = np.zeros(nu.shape)
    main = x >= 1.e-14
    f[main] = x[main] / T / (-np.expm1(-x[main])) * b[main]
    f[~main] = b[~main] / T
    return f


def interp_loglog(x, y, xv):
    with np.errstate(divide="ignore"):
        ly = np.log10(y)
    out = 10. ** np.interp(np.log10(xv), np.log10(x), ly)
    return out


def integrate_loglog(x, y):
    """Integral of a piecewise power law (zero segments where y is zero)."""
    x1, x2, y1, y2 = x[:-1], x[1:], y[:-1], y[1:]
    ok = (y1 > 0) & (y2 > 0) & (x2 > x1)
    out = np.zeros(len(x) - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        b = np.log10(y1[ok] / y2[ok]) / np.log10(x1[ok] / x2[ok])
        near = np.abs(b + 1.) < 1.e-10
        seg = np.where(near, x1[ok] * y1[ok] * np.log(x2[ok] / x1[ok]),
                       y1[ok] * (x2[ok] * (x2[ok] / x1[ok]) ** b - x1[ok]) / (b + 1.))
    out[ok] = seg
    return out.sum()


def make_dust(nu, albedo, chi, g=None, p_lin_max=None, n_temp=1200, temp_min=0.1, temp_max=100000.,
              sublimation_mode=0, sublimation_specific_energy=0.0):
    """IsotropicDust (g is None) or HenyeyGreensteinDust with LTE emissivities."""
    nu = np.asarray(nu, dtype=float)
    albedo = np.asarray(albedo, dtype=float)
    chi = np.asarray(chi, dtype=float)
    order = np.argsort(nu)
    nu, albedo, chi = nu[order], albedo[order], chi[order]
    if g is None:
        mu = np.linspace(-1., 1., 2)
        P1 = np.ones((len(nu), 2))
        P2 = np.zeros((len(nu), 2))
        P3 = np.ones((len(nu), 2))
        P4 = np.zeros((len(nu), 2))
    else:
        gg = np.broadcast_to(np.asarray(g, dtype=float), nu.shape)[order][:, None]
        pl = np.broadcast_to(np.asarray(p_lin_max, dtype=float), nu.shape)[order][:, None]
        mu = np.linspace(-1., 1., 100)
        m = mu[None, :]
        P1 = (1. - gg * gg) / (1. + gg * gg - 2. * gg * m) ** 1.5
        P2 = -pl * P1 * (1. - m * m) / (1. + m * m)
        P3 = P1 * 2. * m / (1. + m * m)
        P4 = np.zeros_like(P1)
    kappa = chi * (1. - albedo)

    temperatures = np.logspace(np.log10(temp_min), np.log10(temp_max), n_temp)
    temperatures[0], temperatures[-1] = temp_min, temp_max
    pn = planck_nu_range(temp_min, temp_max)
    nuc = nu_common(pn, nu)
    nuc = nuc[(nuc >= nu.min()) & (nuc <= nu.max())]
    chi_c = interp_loglog(nu, chi, nuc)
    kap_c = interp_loglog(nu, kappa, nuc)
    n = len(temperatures)
    chi_p, kap_p, chi_ip, kap_ip, chi_r, kap_r = (np.zeros(n) for _ in range(6))
    jnu = np.zeros((len(nuc), n))
    with np.errstate(divide="ignore", invalid="ignore"):
        for it, T in enumerate(temperatures):
            b = B_nu(nuc, T)
            db = dB_nu_dT(nuc, T)
            ib = integrate_loglog(nuc, b)
            chi_p[it] = integrate_loglog(nuc, b * chi_c) / ib
            kap_p[it] = integrate_loglog(nuc, b * kap_c) / ib
            chi_ip[it] = ib / integrate_loglog(nuc, np.where(chi_c > 0, b / chi_c, 0.))
            kap_ip[it] = ib / integrate_loglog(nuc, np.where(kap_c > 0, b / kap_c, 0.))
            idb = integrate_loglog(nuc, db)
            chi_r[it] = idb / integrate_loglog(nuc, np.where(chi_c > 0, db / chi_c, 0.))
            kap_r[it] = idb / integrate_loglog(nuc, np.where(kap_c > 0, db / kap_c, 0.))
            jnu[:, it] = kap_c * b
    specific_energy = 4. * sigma * temperatures ** 4. * kap_p
    d = FlatDust(nu=nu, albedo=albedo, chi=chi, mu=mu, P1=P1, P2=P2, P3=P3, P4=P4,
                 specific_energy=specific_energy, chi_planck=chi_p, kappa_planck=kap_p,
                 chi_inv_planck=chi_ip, kappa_inv_planck=kap_ip, chi_rosseland=chi_r, kappa_rosseland=kap_r,
                 emiss_nu=nuc, emiss_jnu=jnu, jnu_var=specific_energy, version=2, is_lte=True,
                 sublimation_mode=sublimation_mode, sublimation_specific_energy=sublimation_specific_energy)
    d.temperature = temperatures
    return d


def grey_dust(n_temp=10, **kw):
    """hyperion/model/tests/test_helpers.py:14-18: albedo 0.5, chi = 1 cm^2/g."""
    return make_dust([3.e9, 3.e16], [0.5, 0.5], [1., 1.], n_temp=n_temp, temp_min=0.1, temp_max=1600., **kw)


REALISTIC_NU = [3.e7, 1.e10, 2.e11, 2.e12, 2.e13, 2.e14, 2.e15, 2.e16, 2.e17]
REALISTIC_CHI = [1.e-11, 2.e-6, 2.e-3, 0.2, 13., 90., 1000., 700., 700.]
REALISTIC_ALBEDO = [0., 0., 0., 0., 0.1, 0.5, 0.4, 0.4, 0.4]


def realistic_dust(n_temp=40, **kw):
    """hyperion/model/tests/test_helpers.py:21-30 (isotropic scattering)."""
    return make_dust(REALISTIC_NU, REALISTIC_ALBEDO, REALISTIC_CHI, n_temp=n_temp, **kw)


def hg_dust(g=0.6, p_lin_max=0.5, n_temp=40, **kw):
    """The realistic table with Henyey-Greenstein scattering tabulated at 100 mu points."""
    return make_dust(REALISTIC_NU, REALISTIC_ALBEDO, REALISTIC_CHI, g=g, p_lin_max=p_lin_max, n_temp=n_temp, **kw)


def chi_at(dust: FlatDust, nu0):
    return float(interp_loglog(dust.nu, dust.chi, np.array([nu0]))[0])


def cartesian_point_source_model(n=256, tau_edge=1.0, dust=None, temperature=6000., seed=1,
                                 uniform=False, n_photons=0, n_iter=1, lam_ref_um=0.5):
    """SURVEY.md section 8d 'C1' / 'C-headline': n^3 Cartesian grid over [-pc, pc]^3, density
    1+U(0,1) times a base value chosen so that the optical depth from the centre to the
    face centre at ``lam_ref_um`` microns is ``tau_edge``; one 6000 K point source at the origin."""
    if dust is None:
        dust = realistic_dust(n_temp=1200)
    w = np.linspace(-pc, pc, n + 1)
    chi0 = chi_at(dust, c / (lam_ref_um * 1.e-4))
    mean_factor = 1.0 if uniform else 1.5
    rho0 = tau_edge / (chi0 * pc * mean_factor)
    if uniform:
        rho = np.full((1, n, n, n), rho0)
    else:
        rng = np.random.default_rng(seed)
        rho = rng.random((1, n, n, n), dtype=np.float64)
        rho += 1.0
        rho *= rho0
    src = FlatSource(type=1, luminosity=lsun, temperature=temperature, position=(0., 0., 0.))
    conf = FlatConf(n_initial_iter=n_iter, n_initial_photons=n_photons)
    return FlatModel(w, w, w, rho, [dust], [src], conf)


def spherical_disk_model(n_r=399, n_theta=199, n_phi=1, tau_edge=10.0, dust=None, temperature=4000.,
                         lam_ref_um=0.5, n_photons=0, n_iter=1, stellar_sphere=False):
    """SURVEY.md section 8d 'C3': spherical polar (r, theta[, phi]) grid of a flared disk as the
    AnalyticalYSOModel front end lays it out (hyperion/model/analytical_yso_model.py:490-626: r walls
    [0, rmin, rmin (1 + logspace)], theta walls linspace(0, pi) + sin(2 theta)/6, which concentrates
    cells towards the midplane; hyperion/densities/flared_disk.py:286-351: rho ~ (r0/w)^(beta-p)
    exp(-(z/h)^2/2), h = h0 (w/r0)^beta).  The density is scaled so that the midplane optical depth
    from rmin to rmax at ``lam_ref_um`` is ``tau_edge``.  The star is a point source at the origin, or with
    ``stellar_sphere`` the SphericalSource of the tutorial model (R = 2 R_sun: packets that come back to the
    star are re-absorbed and re-emitted from its surface, src/main/iter_lucy.f90:158-185)."""
    if dust is None:
        dust = realistic_dust(n_temp=200)
    rstar = 2. * rsun
    rmin, rmax = 10. * rstar, 200. * au
    beta, p, r0, h0 = 1.25, -1.0, 100. * au, 10. * au
    rnext = rmin * 1.e-3
    w1 = np.hstack([0., rmin, rmin * (1. + np.logspace(np.log10(rnext / rmin), np.log10((rmax - rmin) / rmin), n_r - 1))])
    t = np.linspace(0., np.pi, n_theta + 1)
    w2 = t + np.sin(2. * t) / 6.
    w2[0], w2[-1] = 0., np.pi
    if n_theta % 2 == 0:
        w2[n_theta // 2] = np.pi / 2.
    w3 = np.linspace(0., 2. * np.pi, n_phi + 1)
    rc = np.sqrt(np.maximum(w1[:-1], 1e-30) * w1[1:])
    rc[0] = 0.5 * w1[1]
    tc = 0.5 * (w2[:-1] + w2[1:])
    R, T = np.meshgrid(rc, tc)                       # [n_theta, n_r]
    w = R * np.sin(T)
    z = R * np.cos(T)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        h = h0 * (w / r0) ** beta
        rho = (r0 / w) ** (beta - p) * np.exp(-0.5 * (z / h) ** 2)
    rho[~np.isfinite(rho)] = 0.
    rho[(w < rmin) | (w > rmax)] = 0.
    # midplane optical depth: integrate along the row closest to theta = pi/2
    chi0 = chi_at(dust, c / (lam_ref_um * 1.e-4))
    imid = int(np.argmin(np.abs(tc - np.pi / 2.)))
    tau_mid = float((rho[imid] * np.diff(w1)).sum() * chi0)
    rho *= tau_edge / tau_mid
    rho = np.broadcast_to(rho[None, None], (1, n_phi, n_theta, n_r)).copy()
    if stellar_sphere:
        src = FlatSource(type=2, luminosity=lsun, temperature=temperature, position=(0., 0., 0.), radius=rstar)
    else:
        src = FlatSource(type=1, luminosity=lsun, temperature=temperature, position=(0., 0., 0.))
    conf = FlatConf(n_initial_iter=n_iter, n_initial_photons=n_photons)
    return FlatModel(w1, w2, w3, rho, [dust], [src], conf, grid_type="sph")


def random_octree(max_depth=5, p_refine=0.6, seed=4, min_depth=1):
    """Depth-first refinement flags of a seeded random octree (children in x-fastest order,
    docs/advanced/indepth_oct.rst): a node at depth < min_depth is always refined, deeper nodes with
    probability p_refine until max_depth."""
    rng = np.random.default_rng(seed)
    refined = []

    def build(depth):
        r = depth < max_depth and (depth < min_depth or rng.random() < p_refine)
        refined.append(1 if r else 0)
        if r:
            for _ in range(8):
                build(depth + 1)

    build(0)
    return np.array(refined, dtype=np.int32)


def octree_point_sources_model(refined=None, tau_edge=2.0, dust=None, n_sources=4, seed=4, lam_ref_um=0.5,
                               n_photons=0, n_iter=1, **kw):
    """SURVEY.md section 8d 'C4': a seeded octree over [-pc, pc]^3 with a few point sources at random
    positions and density 1+U(0,1) per leaf (scaled so that the optical depth from the centre to a face
    at ``lam_ref_um`` is ``tau_edge``)."""
    if refined is None:
        refined = random_octree(seed=seed, **kw)
    if dust is None:
        dust = hg_dust(n_temp=40)
    rng = np.random.default_rng(seed + 1)
    chi0 = chi_at(dust, c / (lam_ref_um * 1.e-4))
    rho0 = tau_edge / (chi0 * pc * 1.5)
    rho = (1. + rng.random((1, len(refined)))) * rho0
    src = [FlatSource(type=1, luminosity=lsun * (0.5 + rng.random()), temperature=float(rng.uniform(3000., 9000.)),
                      position=tuple(rng.uniform(-0.8 * pc, 0.8 * pc, 3))) for _ in range(n_sources)]
    conf = FlatConf(n_initial_iter=n_iter, n_initial_photons=n_photons)
    return FlatModel(None, None, None, rho, [dust], src, conf, grid_type="oct", refined=refined,
                     oct_center=(0., 0., 0.), oct_half=(pc, pc, pc))


def amr_point_sources_model(n_root=64, n_levels=3, n_patches=3, seed=5, tau_edge=2.0, dust=None, n_sources=2,
                            lam_ref_um=0.5, n_photons=0, n_iter=1):
    """SURVEY.md section 8d 'C5': a block-structured AMR hierarchy standing in for the yt sample:
    one root grid of n_root^3 cells over [-pc, pc]^3; every finer level has ``n_patches`` seeded
    non-overlapping boxes per parent patch region (refinement factor 2, aligned with parent cells and
    nested inside a patch of the previous level).  Density 1+U(0,1) per cell, scaled so that the
    optical depth from the centre to a face is ``tau_edge``."""
    if dust is None:
        dust = hg_dust(n_temp=40)
    rng = np.random.default_rng(seed)
    levels = [[(n_root, n_root, n_root, -pc, pc, -pc, pc, -pc, pc)]]
    # patches as integer boxes in the index space of their level: (lo, hi) per axis
    boxes = [[((0, 0, 0), (n_root, n_root, n_root))]]
    for lev in range(1, n_levels):
        new_boxes, new_grids = [], []
        nlev = n_root * 2 ** lev
        width = 2. * pc / nlev
        for lo, hi in boxes[-1]:
            # candidate sub-boxes in parent index space, refined by 2; keep those that do not touch each other
            placed = []
            for _ in range(8 * n_patches):
                if len(placed) == n_patches:
                    break
                size = [max(2, int((hi[a] - lo[a]) * rng.uniform(0.2, 0.4))) for a in range(3)]
                start = [int(rng.integers(lo[a] + 1, max(lo[a] + 2, hi[a] - size[a]))) for a in range(3)]
                end = [min(start[a] + size[a], hi[a] - 1) for a in range(3)]
                if any(end[a] - start[a] < 1 for a in range(3)):
                    continue
                if any(all(start[a] <= e2[a] and s2[a] <= end[a] for a in range(3)) for s2, e2 in placed):
                    continue
                placed.append((start, end))
            for start, end in placed:
                s2 = tuple(2 * v for v in start)
                e2 = tuple(2 * v for v in end)
                new_boxes.append((s2, e2))
                new_grids.append((e2[0] - s2[0], e2[1] - s2[1], e2[2] - s2[2],
                                  -pc + s2[0] * width, -pc + e2[0] * width, -pc + s2[1] * width, -pc + e2[1] * width,
                                  -pc + s2[2] * width, -pc + e2[2] * width))
        if not new_grids:
            break
        boxes.append(new_boxes)
        levels.append(new_grids)
    n_cells = sum(g[0] * g[1] * g[2] for lev in levels for g in lev)
    chi0 = chi_at(dust, c / (lam_ref_um * 1.e-4))
    rho0 = tau_edge / (chi0 * pc * 1.5)
    rho = (1. + rng.random((1, n_cells))) * rho0
    src = [FlatSource(type=1, luminosity=lsun * (0.5 + rng.random()), temperature=float(rng.uniform(3000., 9000.)),
                      position=tuple(rng.uniform(-0.7 * pc, 0.7 * pc, 3))) for _ in range(n_sources)]
    conf = FlatConf(n_initial_iter=n_iter, n_initial_photons=n_photons)
    return FlatModel(None, None, None, rho, [dust], src, conf, grid_type="amr", amr_levels=levels)


def voronoi_mesh(sites, box):
    """The Voronoi tessellation of ``sites`` [n, 3] clipped to ``box`` = (xmin, xmax, ymin, ymax, zmin, zmax) in the
    layout ``VoronoiGrid.write`` gives the Fortran code (hyperion/grid/voronoi_grid.py:417-478; the front end computes
    it with voro++): cell volumes, bounding boxes and neighbour lists, walls of the box numbered -1 .. -6.
    Made with scipy: the sites are mirrored in the six walls, so that every cell of the unbounded tessellation of
    sites + mirror images is exactly the clipped cell, and a neighbour that is a mirror image in wall w means the
    cell touches that wall."""
    from scipy.spatial import ConvexHull, Voronoi
    sites = np.asarray(sites, dtype=np.float64)
    n = len(sites)
    box = np.asarray(box, dtype=np.float64)
    pts = [sites]
    for w in range(6):
        m = sites.copy()
        m[:, w // 2] = 2.0 * box[w] - m[:, w // 2]
        pts.append(m)
    vor = Voronoi(np.concatenate(pts))
    neigh = [[] for _ in range(n)]
    for a, b in vor.ridge_points:
        for i, j in ((a, b), (b, a)):
            if i < n:
                neigh[i].append(j if j < n else -(j // n))        # mirror block k (1..6) -> wall -k
    volume, bb_min, bb_max = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3))
    for i in range(n):
        reg = vor.regions[vor.point_region[i]]
        if -1 in reg or len(reg) < 4:
            raise ValueError("open Voronoi cell: is a site outside the box?")
        v = vor.vertices[reg]
        volume[i] = ConvexHull(v).volume
        bb_min[i], bb_max[i] = v.min(0), v.max(0)
    idx = np.concatenate([[0], np.cumsum([len(set(x)) for x in neigh])]).astype(np.int32)
    flat = np.concatenate([sorted(set(x), reverse=True) for x in neigh]).astype(np.int32)
    return dict(coordinates=sites, bb_min=bb_min, bb_max=bb_max, volume=volume, sparse_neighs=flat, sparse_idx=idx, box=box)


def lattice_voronoi(w1, w2, w3):
    """The Voronoi mesh whose cells are the cells of a Cartesian grid: sites at the cell centres, the six face
    neighbours (or walls of the box) of every cell, in x-fastest order."""
    w1, w2, w3 = [np.asarray(w, dtype=np.float64) for w in (w1, w2, w3)]
    n1, n2, n3 = len(w1) - 1, len(w2) - 1, len(w3) - 1
    c1, c2, c3 = [0.5 * (w[1:] + w[:-1]) for w in (w1, w2, w3)]
    Z, Y, X = np.meshgrid(c3, c2, c1, indexing="ij")
    sites = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    lo = np.stack(np.meshgrid(w3[:-1], w2[:-1], w1[:-1], indexing="ij")[::-1], axis=-1).reshape(-1, 3)
    hi = np.stack(np.meshgrid(w3[1:], w2[1:], w1[1:], indexing="ij")[::-1], axis=-1).reshape(-1, 3)
    neigh = []
    for i3 in range(n3):
        for i2 in range(n2):
            for i1 in range(n1):
                ic = (i3 * n2 + i2) * n1 + i1
                neigh.append([ic - 1 if i1 > 0 else -1, ic + 1 if i1 < n1 - 1 else -2,
                              ic - n1 if i2 > 0 else -3, ic + n1 if i2 < n2 - 1 else -4,
                              ic - n1 * n2 if i3 > 0 else -5, ic + n1 * n2 if i3 < n3 - 1 else -6])
    idx = (6 * np.arange(len(neigh) + 1)).astype(np.int32)
    return dict(coordinates=sites, bb_min=lo, bb_max=hi, volume=np.prod(hi - lo, axis=1),
                sparse_neighs=np.array(neigh, dtype=np.int32).ravel(), sparse_idx=idx,
                box=np.array([w1[0], w1[-1], w2[0], w2[-1], w3[0], w3[-1]]))
