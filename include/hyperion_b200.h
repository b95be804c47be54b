/*
 * hyperion_b200.h -- C ABI of the B200-native photon-packet propagation engine.
 *
 * The reference (hyperion-rt/hyperion) has no FFI on this path: the Python
 * front end writes an .rtin file and shells out to a Fortran binary
 * (hyperion/model/model.py:1025-1080, scripts/hyperion:39-92, program main
 * src/main/main.f90:1).  This header is the seam a maintainer would bind
 * instead of that process boundary.  Each entry point
 * cites the reference routine whose job it takes over.
 *
 * Conventions
 *   - plain C, no C++ / torch types; all arrays are caller-owned, contiguous,
 *     C-ordered HOST buffers unless a name says "device"
 *   - every function returns 0 on success, <0 on error; the message is
 *     available from hyp_last_error() (thread-local); nothing calls exit()
 *   - one hyp_ctx drives ONE GPU (one process per GPU); multi-GPU runs shard
 *     photon packets across processes and all-reduce the deposit grid
 *     (hyp_lucy_device_buffers + NCCL, replacing src/mpi/mpi_routines.f90)
 *   - grids use the .rtin layout: density[n_dust][n3][n2][n1] (x fastest),
 *     identical in bytes to the Fortran (n_cells, n_dust) column-major arrays
 *     (src/core/type_cell_id_3d.f90:97-102)
 */
#ifndef HYPERION_B200_H
#define HYPERION_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hyp_ctx hyp_ctx;

/* error codes */
#define HYP_OK 0
#define HYP_ERR_INVALID -1      /* bad argument / inconsistent model */
#define HYP_ERR_CUDA -2         /* CUDA runtime failure */
#define HYP_ERR_STATE -3        /* call out of order */
#define HYP_ERR_PHYSICS -4      /* reference error() condition hit (message matches the reference text) */

/* Dust tables for one dust type, exactly the columns of a Hyperion dust file
 * (read by dust_setup, src/dust/dust_type_4elem.f90:78-293).  2-D tables are
 * in file (C) order. */
typedef struct {
  int32_t version;                 /* root attr 'version' (1 or 2) */
  int32_t is_lte;                  /* root attr 'lte' */
  int32_t sublimation_mode;        /* 0 no, 1 fast, 2 slow, 3 cap */
  double sublimation_specific_energy;
  int32_t n_nu;                    /* optical_properties rows */
  const double *nu, *albedo, *chi; /* [n_nu] */
  int32_t n_mu;                    /* scattering_angles rows */
  const double *mu;                /* [n_mu] */
  const double *P1, *P2, *P3, *P4; /* [n_nu][n_mu] */
  int32_t n_e;                     /* mean_opacities rows */
  const double *specific_energy;   /* [n_e] */
  const double *chi_planck, *kappa_planck;
  const double *chi_inv_planck, *kappa_inv_planck; /* version 1 files: pass the rosseland columns (dust_type_4elem.f90:232-238) */
  const double *chi_rosseland, *kappa_rosseland;
  int32_t n_emiss_nu;              /* emissivities rows */
  const double *emiss_nu;          /* [n_emiss_nu] */
  int32_t n_jnu;                   /* emissivity_variable rows */
  const double *emiss_jnu;         /* [n_emiss_nu][n_jnu] */
  const double *jnu_var;           /* [n_jnu] specific energies */
} hyp_dust_tables;

/* One source (source_read, src/sources/source_type.f90:102-282). */
#define HYP_SOURCE_POINT 1
#define HYP_SOURCE_SPHERE 2
/* the reference's type 3 (sphere with spots) is HYP_SOURCE_SPHERE with n_spots > 0 */
#define HYP_SOURCE_MAP 4               /* emit_from_map, source_type.f90:713-746 */
#define HYP_SOURCE_EXTERN_SPH 5        /* emit_from_extern_sph, source_type.f90:748-820 */
#define HYP_SOURCE_EXTERN_BOX 6        /* emit_from_extern_box, source_type.f90:822-933 */
#define HYP_SOURCE_PLANE_PARALLEL 7    /* emit_from_plane_parallel, source_type.f90:935-980 */
#define HYP_SOURCE_POINT_COLLECTION 8  /* emit_from_point_collection, source_type.f90:570-598 */
#define HYP_SPECTRUM_TABLE 1
#define HYP_SPECTRUM_BLACKBODY 2
#define HYP_SPECTRUM_LTE 3             /* map sources only: emissivity of the dust in the emitting cell */
/* A spot on a spherical source (source_read, src/sources/source_type.f90:150-188): a cap of angular
 * radius `radius` around (longitude, latitude) with its own luminosity and spectrum. */
typedef struct {
  double luminosity;
  double longitude, latitude, radius;   /* degrees */
  int32_t spectrum_type;                /* HYP_SPECTRUM_TABLE / _BLACKBODY */
  double temperature;
  int32_t n_spec;
  const double *spec_nu, *spec_fnu;
} hyp_spot;

typedef struct {
  int32_t type;            /* HYP_SOURCE_* (the reference's numbering, source_read) */
  int32_t peeloff;
  double luminosity;       /* point collection: ignored, the sum of points_lum is used (source_type.f90:268) */
  double x, y, z;          /* point, sphere, extern_sph, plane_parallel */
  double radius;           /* sphere, extern_sph, plane_parallel (radius of the disk the beam starts from) */
  int32_t limb_darkening;  /* sphere only */
  int32_t spectrum_type;   /* HYP_SPECTRUM_* */
  double temperature;      /* blackbody */
  int32_t n_spec;          /* tabulated spectrum */
  const double *spec_nu, *spec_fnu;
  double box[6];           /* extern_box: xmin, xmax, ymin, ymax, zmin, zmax */
  double theta, phi;       /* plane_parallel: direction of travel in degrees ('theta', 'phi') */
  int64_t n_points;        /* point collection */
  const double *points_xyz;  /* [n_points][3] */
  const double *points_lum;  /* [n_points] */
  int64_t n_map;           /* map: number of cells */
  const double *map;       /* map: luminosity per cell, in cell-id order (the 'Luminosity map' dataset) */
  int32_t n_spots;         /* sphere: spots (the reference's source type 3); luminosity is the star's own */
  const hyp_spot *spots;
} hyp_source;

/* Run configuration: the root attributes of the .rtin file
 * (setup_initial, src/main/setup_rt.f90:38-157; defaults
 * hyperion/conf/conf_files.py:48-73). */
typedef struct {
  int64_t seed;                    /* 'seed' (default -124902) */
  int64_t n_inter_max;             /* 'n_inter_max' */
  int64_t n_reabs_max;             /* 'n_reabs_max' */
  int32_t kill_on_absorb;
  int32_t kill_on_scatter;
  int32_t sample_sources_evenly;
  int32_t enforce_energy_range;
  int32_t use_mrw;                 /* modified random walk (grid_mrw_3d.f90), on the device */
  double mrw_gamma;
  int64_t n_mrw_max;
  double propagation_check_frequency; /* reference self-check rate; see DESIGN.md */
  /* final iteration (src/main/iter_final.f90:191-209, src/main/forced_interaction.f90) */
  int32_t forced_first_interaction;            /* default on (hyperion/conf/conf_files.py:66) */
  int32_t forced_first_interaction_algorithm;  /* HYP_FFI_WR99 / HYP_FFI_BAES16 */
  double baes16_xi;
  /* specific_energy_type = 'additional' (setup_grid_physics, src/grid/grid_physics_3d.f90:213-235;
   * update_energy_abs :537-545): the specific energy passed in is an extra heating term that is added
   * after every Lucy iteration, and the iterations start from the minimum specific energy. */
  int32_t specific_energy_additional;
  /* 'pda' (setup_rt.f90:75): after every Lucy iteration the specific energy of the cells fewer than
   * max(30, 0.005 x mean) packets visited is replaced by the solution of the diffusion equation between their
   * better-sampled neighbours (solve_pda, src/grid/grid_pda_3d.f90:105-169; Cartesian and polar grids). */
  int32_t use_pda;
  /* keep the per-cell packet counter n_photons without the PDA (output_n_photons /= 'none',
   * src/grid/grid_physics_3d.f90:308-317) */
  int32_t count_photons;
} hyp_run_conf;

#define HYP_FFI_WR99 1
#define HYP_FFI_BAES16 2

/* One peeled image / SED group: the attributes of Output/Peeled/group_%05i in the .rtin file
 * (peeled_images_setup, src/images/images_peeled.f90:272-382; image_setup,
 * src/images/image_type.f90:153-335; written by hyperion/conf/conf_files.py PeeledImageConf). */
#define HYP_TRACK_NO 0
#define HYP_TRACK_BASIC 1
#define HYP_TRACK_DETAILED 2
#define HYP_TRACK_SCATTERINGS 3
typedef struct {
  int32_t n_view;
  const double *theta, *phi;                 /* [n_view] viewing angles, degrees (table 'angles') */
  int32_t inside_observer;                   /* the observer sits at the peeloff origin; theta/phi = where it looks */
  int32_t ignore_optical_depth;
  double peeloff_x, peeloff_y, peeloff_z;    /* peeloff origin */
  double d_min, d_max;                       /* depth cut along the line of sight (+-inf: none) */
  int32_t compute_image, n_x, n_y;
  double x_min, x_max, y_min, y_max;
  int32_t compute_sed, n_ap;
  double ap_min, ap_max;
  int32_t n_wav;
  double wav_min, wav_max;                   /* microns */
  int32_t track_origin;                      /* HYP_TRACK_* */
  int32_t track_n_scat;
  int32_t uncertainties;
  int32_t compute_stokes;
  int32_t io_bytes;                          /* 4 or 8: precision the caller will store */
  /* Binned images (Output/Binned/group_00001, src/images/images_binned.f90): instead of peel-offs, every
   * packet that escapes in the final iteration is binned by its own direction into n_theta x n_phi
   * views (cos theta in [-1, 1], phi in [0, 2 pi]).  Then n_view must be n_theta * n_phi, theta / phi
   * are ignored, and at most one such group may exist; it needs forced_first_interaction = 0
   * (setup_rt.f90:327-329). */
  int32_t binned, n_theta, n_phi;
  /* Filter convolution (image_setup, src/images/image_type.f90:174-183,274-284; image_bin :467-476): the
   * n_wav channels are n_filt = n_wav filters; a packet of frequency nu adds energy x transmission(nu)
   * to every filter whose (linearly interpolated) transmission is > 0.  wav_min / wav_max are ignored.
   * Filters cannot be combined with raytracing (image_type.f90:541). */
  int32_t use_filters;
  const int32_t *filt_n;     /* [n_wav] points of each filter curve */
  const double *filt_nu;     /* concatenated 'nu' columns of Output/Peeled/group/filter_%05i */
  const double *filt_tr;     /* concatenated 'tn' columns (normalised transmission) */
  const double *filt_nu0;    /* [n_wav] central frequencies (attribute nu0), written back as filt_nu0 */
  /* Monochromatic mode (image_setup, src/images/image_type.f90:243-258): the n_wav channels are the
   * frequencies inu_min .. inu_max (1-based) of hyp_set_monochromatic; wav_min / wav_max are ignored.
   * 0 = not monochromatic. */
  int32_t inu_min, inu_max;
} hyp_image_conf;

/* Per-iteration counters (killed_photons_* attrs of main.f90:225-230 plus the
 * work counters the roofline needs). */
typedef struct {
  double energy_emitted;      /* sum of emitted packet weights (energy_current, source.f90:163) */
  int64_t n_photons;          /* packets run by this ctx */
  int64_t killed_geo;
  int64_t killed_int;
  int64_t n_crossings;        /* cell crossings in grid_integrate */
  int64_t n_absorptions;      /* absorb + re-emit events */
  int64_t n_scatterings;
  int64_t n_escaped;
  double kernel_ms;           /* device time of the photon loop: all rounds of emit/flight/interact (CUDA events) */
  double epilogue_ms;         /* device time of scale/clamp/jnu_var kernels */
  double flight_ms;           /* device time of the flight kernel alone (the HBM-bound part) */
  int64_t n_rounds;           /* rounds of the packet pool */
  int64_t n_launches;         /* kernels of this library launched for the iteration */
  int64_t n_peel_crossings;   /* cell crossings of peel-off marches (grid_escape_tau / _column_density) */
  int64_t n_peeloffs;         /* peel-off contributions binned */
  int64_t n_peel_cached;      /* part of n_peel_crossings served by the point-source column cache (not marched) */
  int64_t n_wave_rounds;      /* rounds of n_rounds run by the wave engine (tile visits in shared memory); 0: direct kernels only */
} hyp_iter_stats;

const char *hyp_last_error(void);
int hyp_version(void);
/* sizeof() of the ABI structs as compiled: 0 hyp_dust_tables, 1 hyp_source, 2 hyp_run_conf, 3 hyp_iter_stats
 * (lets a binding verify its mirror of this header) */
int hyp_sizeof(int which);

/* replaces: program start-up, mp_initialize (src/mpi/mpi_core.f90:35) */
int hyp_ctx_create(int device_id, hyp_ctx **out);
void hyp_ctx_destroy(hyp_ctx *ctx);

/* replaces: setup_grid_geometry (src/grid/grid_geometry_cartesian_3d.f90:77-135).
 * w1/w2/w3 are the n+1 wall positions (Grid/Geometry walls_1..3). */
int hyp_set_grid_cartesian(hyp_ctx *ctx, int32_t n1, int32_t n2, int32_t n3,
                           const double *w1, const double *w2, const double *w3);

/* replaces: setup_grid_geometry (src/grid/grid_geometry_spherical_3d.f90:92-203).
 * w1 = r walls (n1+1), w2 = theta walls in [0, pi] (n2+1), w3 = phi walls in [0, 2 pi] (n3+1)
 * (Grid/Geometry walls_1 'r', walls_2 't', walls_3 'p' of a 'sph_pol' grid).  Cell ids and the
 * density layout are as for Cartesian grids: density[n_dust][n3][n2][n1], r fastest. */
int hyp_set_grid_spherical(hyp_ctx *ctx, int32_t n1, int32_t n2, int32_t n3,
                           const double *w1, const double *w2, const double *w3);

/* replaces: setup_grid_geometry (src/grid/grid_geometry_cylindrical_3d.f90:90-177).
 * w1 = cylindrical radius walls (n1+1), w2 = z walls (n2+1), w3 = phi walls in [0, 2 pi] (n3+1)
 * (walls_1 'w', walls_2 'z', walls_3 'p' of a 'cyl_pol' grid); density[n_dust][n3][n2][n1]. */
int hyp_set_grid_cylindrical(hyp_ctx *ctx, int32_t n1, int32_t n2, int32_t n3,
                             const double *w1, const double *w2, const double *w3);

/* replaces: setup_grid_geometry + octree_setup_indiv (src/grid/grid_geometry_octree.f90:148-262).
 * refined[n_cells]: depth-first (pre-order) refinement flags, children in x-fastest order
 * (Grid/Geometry table 'cells', column 'refined'); (x, y, z) centre and (dx, dy, dz) HALF-widths of the root
 * cell (group attributes).  Quantities then have one entry per NODE: density[n_dust][n_cells]; refined nodes
 * are masked out (density forced to zero), as setup_grid_physics does. */
int hyp_set_grid_octree(hyp_ctx *ctx, int32_t n_cells, const int32_t *refined,
                        double x, double y, double z, double dx, double dy, double dz);

/* replaces: read_grid / read_level / setup_grid_geometry (src/grid/grid_geometry_amr.f90:111-507).
 * n_grids[n_levels] grids per level (level 1 = coarsest); then per grid, level-major, in file order
 * (Grid/Geometry/level_%05d/grid_%05d): dims[3] = n1, n2, n3 and bounds[6] = xmin, xmax, ymin, ymax, zmin, zmax.
 * Quantities have one entry per cell of every grid, concatenated in the same order, x fastest inside a
 * grid (src/core/type_cell_id_amr.f90:115-133): density[n_dust][n_cells]; cells covered by a finer grid
 * are masked out. */
int hyp_set_grid_amr(hyp_ctx *ctx, int32_t n_levels, const int32_t *n_grids, const int32_t *dims, const double *bounds);

/* replaces: setup_grid_geometry (src/grid/grid_geometry_voronoi.f90:92-187).
 * The 'cells' table and neighbour lists VoronoiGrid.write stores (hyperion/grid/voronoi_grid.py:417-478):
 * coords / bb_min / bb_max [n_cells][3] (sites and bounding boxes), volume[n_cells] (cells with volume <= 0 are
 * masked out), sparse_idx[n_cells + 1] / sparse_neighs[sparse_idx[n_cells]] (CSR neighbour lists in the file's
 * numbering: >= 0 a cell, -1 .. -6 the walls xmin, xmax, ymin, ymax, zmin, zmax), box[6] = xmin, xmax, ymin, ymax,
 * zmin, zmax.  Quantities have one entry per cell: density[n_dust][n_cells].  No PDA and no modified random walk
 * on this grid (grid_pda_disabled.f90; distance_to_closest_wall, grid_geometry_voronoi.f90:314-320). */
int hyp_set_grid_voronoi(hyp_ctx *ctx, int32_t n_cells, const double *coords, const double *bb_min, const double *bb_max,
                         const double *volume, const int32_t *sparse_idx, const int32_t *sparse_neighs, const double *box);

/* replaces: the specific_energy_spectrum_bin_edges input (src/main/setup_rt.f90:98-104) and the spectrum part of
 * setup_grid_physics (src/grid/grid_physics_3d.f90:124-143,269-284).  n_edges strictly increasing frequencies (Hz):
 * the Lucy iterations then also keep the deposits per frequency bin of the absorbed packets
 * (grid_propagate_3d.f90:155-158,217-225; MRW deposits by the local emissivity, grid_physics_3d.f90:367-395).
 * Call before hyp_finalize_setup.  With several processes the per-bin sums travel behind the scalars and the packet
 * counts in hyp_lucy_device_buffers (mp_collect_physical_arrays, src/mpi/mpi_routines.f90:292-301). */
int hyp_set_specific_energy_spectrum_bins(hyp_ctx *ctx, int32_t n_edges, const double *edges);

/* replaces: output_grid 'specific_energy_spectrum' (src/grid/grid_generic.f90:68-84): out[n_bins][n_dust][n_cells] */
int hyp_get_specific_energy_spectrum(hyp_ctx *ctx, double *out);

/* replaces: dust_setup (src/dust/dust_type_4elem.f90:78-293); call once per dust type, in order */
int hyp_add_dust(hyp_ctx *ctx, const hyp_dust_tables *dust);

/* replaces: source_read / setup_sources (src/sources/source.f90:48-80) */
int hyp_add_source(hyp_ctx *ctx, const hyp_source *src);

/* replaces: the root-attribute block of setup_initial (src/main/setup_rt.f90:38-157) */
int hyp_set_run_conf(hyp_ctx *ctx, const hyp_run_conf *conf);

/* replaces: setup_grid_physics (src/grid/grid_physics_3d.f90:111-322).
 * density: [n_dust][n_cells]; specific_energy may be NULL (then the minimum is
 * used); minimum_specific_energy: [n_dust] or NULL (zeros).  After hyp_finalize_setup
 * hyp_set_density replaces the densities in place and `density` may also be a device
 * pointer on the context's GPU (e.g. a buffer filled by an NCCL broadcast from the
 * rank that read the file, src/mpi/mpi_io.f90:213-242). */
int hyp_set_density(hyp_ctx *ctx, int32_t n_dust, const double *density);
int hyp_set_specific_energy(hyp_ctx *ctx, const double *specific_energy,
                            const double *minimum_specific_energy);

/* Builds the sampling tables and uploads everything to the device; after this
 * the model is frozen except for density / specific_energy. */
int hyp_finalize_setup(hyp_ctx *ctx);

/* replaces: do_lucy (src/main/iter_lucy.f90:66-237) in three steps so that a
 * multi-process host can all-reduce between the photon loop and the scaling:
 *   begin   = grid_reset_energy + precompute_jnu_var        (iter_lucy.f90:101-107)
 *   photons = the photon loop for packets [first_id, first_id+n)   (iter_lucy.f90:119-209)
 *   finish  = update_energy_abs(energy_total/energy_current) + sublimate_dust (iter_lucy.f90:224-235)
 * hyp_run_lucy_iteration does all three for a single process. */
int hyp_lucy_begin(hyp_ctx *ctx);
int hyp_lucy_photons(hyp_ctx *ctx, int64_t first_id, int64_t n_photons, int64_t iteration);
/* Device pointers for the host's collective: sum grid [n_dust*n_cells] fp64 followed
 * directly by 11 fp64 scalars (energy_emitted, killed_geo, killed_int, crossings,
 * absorptions, scatterings, escaped, photons, peel crossings, peel-offs, cached peel crossings);
 * n_values = n_dust*n_cells + 11.
 * replaces: mp_collect_physical_arrays + mp_sync (src/mpi/mpi_routines.f90:272-361) */
int hyp_lucy_device_buffers(hyp_ctx *ctx, void **sum_and_scalars, int64_t *n_values);
int hyp_lucy_finish(hyp_ctx *ctx, hyp_iter_stats *stats);
int hyp_run_lucy_iteration(hyp_ctx *ctx, int64_t n_photons, int64_t iteration, hyp_iter_stats *stats);

/* replaces: output_grid 'specific_energy' (src/grid/grid_generic.f90:50-63): [n_dust][n_cells] */
int hyp_get_specific_energy(hyp_ctx *ctx, double *out);
int hyp_get_density(hyp_ctx *ctx, double *out);
/* replaces: the n_photons dataset of output_grid (src/grid/grid_generic.f90:40-46): number of packets that visited
 * each cell in the last Lucy iteration, [n3][n2][n1] (grid_propagate_3d.f90:90-95,175-180), summed over the
 * processes once the host has reduced hyp_lucy_device_buffers.  Needs use_pda or count_photons. */
int hyp_get_n_photons(hyp_ctx *ctx, int64_t *out);
/* replaces: solve_pda (src/grid/grid_pda_3d.f90:105-169) on the current specific energy: the cells that fewer than
 * max(30, 0.005 x mean) packets visited (and that do not lie on the edge of the grid) get the solution of the
 * diffusion equation between their neighbours.  n_photons: [n3][n2][n1] packet counts, or NULL for the counts of the
 * last Lucy iteration.  hyp_lucy_finish calls this itself when use_pda is set.  n_pda_cells may be NULL. */
int hyp_solve_pda(hyp_ctx *ctx, const int64_t *n_photons, int64_t *n_pda_cells);
/* raw deposit sums of the last iteration (specific_energy_sum, grid_physics_3d.f90:40) */
int hyp_get_energy_sum(hyp_ctx *ctx, double *out);

/* ---- final (imaging) and raytracing iterations ------------------------------------------------
 * replaces: setup_final_iteration (src/main/setup_rt.f90:306-347) -- call once per peeled group,
 * in file order, any time before hyp_run_final. */
int hyp_add_peeled_group(hyp_ctx *ctx, const hyp_image_conf *conf);
/* replaces: do_final (src/main/iter_final.f90:60-145) for packets [first_id, first_id+n).
 * peeloff_scattering_only is main.f90:274's use_raytracing.  Images accumulate in device memory;
 * hyp_final_finish applies peeled_images_adjust_scale(energy_total / energy_current)
 * (iter_final.f90:140-143) after the host has all-reduced hyp_image_device_buffers. */
int hyp_final_begin(hyp_ctx *ctx);
int hyp_final_photons(hyp_ctx *ctx, int64_t first_id, int64_t n_photons, int32_t peeloff_scattering_only);
/* Monochromatic mode ('monochromatic' = yes, the table /frequencies and 'monochromatic_energy_threshold' of
 * the .rtin file; src/main/setup_rt.f90:49-56,220-222).  Call before hyp_add_peeled_group / hyp_finalize_setup. */
int hyp_set_monochromatic(hyp_ctx *ctx, int32_t n_nu, const double *frequencies, double energy_threshold);
/* replaces: do_final_mono (src/main/iter_final_mono.f90:58-229) for ONE frequency inu (1-based), between
 * hyp_final_begin and hyp_final_finish: source packets [first_source_id, +n_sources) of the job's n_total_sources
 * emitted AT that frequency with the spectrum's probability as weight, then thermal packets [first_dust_id,
 * +n_dust) of n_total_dust drawn from the per-cell emission probability at that frequency
 * (src/grid/grid_monochromatic.f90:50-176); every interaction is a scattering that multiplies the energy by the
 * albedo, packets die below energy_threshold x their initial energy.  The cubes need no scaling afterwards. */
int hyp_final_mono_photons(hyp_ctx *ctx, int32_t inu, int64_t first_source_id, int64_t n_sources, int64_t n_total_sources,
                           int64_t first_dust_id, int64_t n_dust, int64_t n_total_dust, int32_t peeloff_scattering_only);
int hyp_final_finish(hyp_ctx *ctx, hyp_iter_stats *stats);
/* replaces: do_raytracing (src/main/iter_raytracing.f90:31-141): n_sources packets from the
 * sources [first_source_id ...) and n_dust packets from random cells [first_dust_id ...), each
 * peeled off polychromatically.  n_total_* are the whole job's counts (the weights divide by them). */
int hyp_raytracing_photons(hyp_ctx *ctx, int64_t first_source_id, int64_t n_sources, int64_t n_total_sources,
                           int64_t first_dust_id, int64_t n_dust, int64_t n_total_dust, hyp_iter_stats *stats);
/* All image / SED accumulators of all groups as one contiguous device buffer (fp64) followed by
 * the same 11 scalars, for the host's collective.  replaces: mp_collect_images (src/mpi/mpi_routines.f90:363-471) */
int hyp_image_device_buffers(hyp_ctx *ctx, void **buffer, int64_t *n_values);
/* Shapes in file order: seds (n_stokes, n_orig, n_view, n_ap, n_wav), images (n_stokes, n_orig,
 * n_view, n_y, n_x, n_wav) (src/images/image_type.f90:291,299 reversed, as HDF5 stores them). */
int hyp_image_shape(hyp_ctx *ctx, int32_t group, int32_t which /* 0 sed, 1 image */, int64_t dims[6], int32_t *ndim);
/* replaces: image_write (src/images/image_type.f90:608-788): values as the reference writes them
 * (divided by the relative bin width, apertures cumulative); unc may be NULL. */
int hyp_get_sed(hyp_ctx *ctx, int32_t group, double *sed, double *unc);
int hyp_get_image(hyp_ctx *ctx, int32_t group, double *image, double *unc);

/* stream handle (cudaStream_t) the ctx launches on, for callers that time with events */
void *hyp_stream(hyp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* HYPERION_B200_H */
