"""Synthetic stand-ins for the configurations BASELINE.json names: the model each one
runs and what it does with it.  Used by ``bench.py --workload`` / its ``other_workloads`` rows and by the
parity tests at configuration scale (``tests/test_gpu_configs.py``).

=====  ====================================================================================================
c1     3-D Cartesian 128^3, one point source, isotropic dust, 5 Lucy iterations
c3     spherical polar (399, 199, 1) flared disk as AnalyticalYSOModel lays it out
       (hyperion/model/analytical_yso_model.py:490-626, hyperion/densities/flared_disk.py:286-351) around a
       stellar SPHERE of 2 R_sun (re-absorption on the star, docs/tutorials/scripts/class2_sed_setup.py),
       Lucy iteration + final iteration peeled into a 150-wavelength SED seen from 10 inclinations
c4     octree of about a million leaves, 4 point sources, Henyey-Greenstein dust (g = 0.6, p_lin = 0.5 tabulated
       at 100 angles, hyperion/dust/dust_type.py:565-585), Lucy iteration
c5     block-structured AMR (3 levels, root 64^3) with the modified random walk (gamma = 2), final imaging
       iteration peeled into an image and an SED
tau5   the 256^3 headline at a centre-to-face optical depth of 5
=====  ====================================================================================================

Photon counts are per GPU and sized so that a step takes about a second; the configuration strings' totals
(1e8, 1e9) only set how long the reference's run would last, the rate is what is measured.
"""
from __future__ import annotations

import numpy as np

from . import synthetic as syn
from .flatmodel import FlatPeeledGroup, FlatSource

NAMES = ("c1", "c3", "c4", "c5", "tau5")


def octree_refined(n_leaves_target=1_000_000, seed=4):
    """Refinement flags (depth-first, docs/advanced/indepth_oct.rst) of an octree with about
    ``n_leaves_target`` leaves: fully refined down to a base depth, then every node of that depth is refined
    once more with the probability that gives the target on average (seeded)."""
    base = 1
    while 8 ** (base + 1) <= n_leaves_target:
        base += 1
    n_base = 8 ** base
    p = min(1.0, max(0.0, (n_leaves_target / n_base - 1.0) / 7.0))
    rng = np.random.default_rng(seed)
    extra = rng.random(n_base) < p
    # nodes of the base depth in depth-first order: [0] or [1, 0 x 8]; the levels above are fully refined
    parts = []
    it = iter(extra)

    def walk(depth):
        if depth == base:
            parts.append(_REFINED_ONCE if next(it) else _LEAF)
            return
        parts.append(_NODE)
        for _ in range(8):
            walk(depth + 1)

    walk(0)
    return np.concatenate(parts)


_LEAF = np.zeros(1, dtype=np.int32)
_NODE = np.ones(1, dtype=np.int32)
_REFINED_ONCE = np.array([1] + [0] * 8, dtype=np.int32)


def build(name, scale=1.0):
    """(model, plan) of workload ``name``.  plan: dict(kind='lucy'|'final', photons=..., iterations=...,
    description=...).  ``scale`` < 1 shrinks grids and photon counts for the tests."""
    if name == "c1":
        n = max(16, int(round(128 * scale ** (1. / 3.))))
        m = syn.cartesian_point_source_model(n=n, tau_edge=1.0, dust=syn.realistic_dust(n_temp=1200), seed=1)
        return m, dict(kind="lucy", photons=int(2e7 * scale), iterations=5,
                       description="cartesian_%d^3_point_source_isotropic_dust_5_lucy_iterations" % n)
    if name == "tau5":
        n = max(16, int(round(256 * scale ** (1. / 3.))))
        m = syn.cartesian_point_source_model(n=n, tau_edge=5.0, dust=syn.realistic_dust(n_temp=1200), seed=1)
        return m, dict(kind="lucy", photons=int(2e7 * scale), iterations=1,
                       description="cartesian_%d^3_point_source_6000K_isotropic_dust_tau5" % n)
    if name == "c3":
        n_r = max(20, int(round(399 * scale ** 0.5)))
        n_t = max(11, int(round(199 * scale ** 0.5)) | 1)
        m = syn.spherical_disk_model(n_r=n_r, n_theta=n_t, n_phi=1, tau_edge=10.0, dust=syn.realistic_dust(n_temp=200),
                                     stellar_sphere=True)
        rmax = float(m.w1[-1])
        m.peeled = [FlatPeeledGroup(theta=np.linspace(0., 90., 10), phi=np.zeros(10), wavelengths=(150, 0.02, 2000.),
                                    sed=(1, 0., rmax), stokes=True)]
        return m, dict(kind="lucy+final", photons=int(4e6 * scale), final_photons=int(1e6 * scale), iterations=1,
                       description="spherical_%dx%d_flared_disk_stellar_sphere_lucy_plus_sed_10_views" % (n_r, n_t))
    if name == "c4":
        refined = octree_refined(int(1_000_000 * scale), seed=4)
        m = syn.octree_point_sources_model(refined=refined, tau_edge=2.0, dust=syn.hg_dust(g=0.6, p_lin_max=0.5, n_temp=200),
                                           n_sources=4, seed=4)
        return m, dict(kind="lucy", photons=int(4e6 * scale), iterations=1,
                       description="octree_%d_leaves_4_point_sources_hg_dust" % int((refined == 0).sum()))
    if name == "c5":
        n_root = max(8, int(round(64 * scale ** (1. / 3.))))
        m = syn.amr_point_sources_model(n_root=n_root, n_levels=3, n_patches=3, seed=5, tau_edge=2.0,
                                        dust=syn.hg_dust(g=0.6, p_lin_max=0.5, n_temp=200), n_sources=2)
        m.conf.use_mrw = True
        m.conf.mrw_gamma = 2.0
        m.conf.n_mrw_max = 1000
        half = syn.pc
        m.peeled = [FlatPeeledGroup(theta=[45., 90.], phi=[30., 120.], wavelengths=(20, 0.1, 1000.),
                                    image=(128, 128, -half, half, -half, half), sed=(1, 0., 2. * half), stokes=True)]
        return m, dict(kind="final", photons=int(2e6 * scale), iterations=1,
                       description="amr_3_levels_root_%d^3_mrw_gamma2_final_imaging_2_views_128x128x20" % n_root)
    raise ValueError("unknown workload %r (one of %s)" % (name, ", ".join(NAMES)))
