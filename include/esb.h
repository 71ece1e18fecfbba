/*
 * esb.h -- C ABI of libesb.so, the B200 (sm_100a) engine for the EnhancerShoeBox hot path.
 *
 * The reference has no plugin API: its seam is "Python driver <-> the `lammps` module"
 * (ctypes -> liblammps).  Each entry point below names the reference call it stands in for
 * (paths relative to the reference checkout; VGM = VisitorGeneMaps).  One handle owns a BATCH of
 * independent replicas of one system layout (same box / topology / bead count; positions, types,
 * seeds and activation parameters per replica) on one GPU.  All calls on one handle must come
 * from one host thread at a time.  Every function returns 0 on success or a negative ESB_E*
 * code; esb_last_error() describes the last failure of the calling thread.
 *
 * Pointers are plain host pointers unless the parameter is documented as a device pointer.
 * No torch types cross this boundary; a torch user passes tensor.data_ptr() and a raw
 * cudaStream_t (torch.cuda.current_stream().cuda_stream) as void*.
 */
#ifndef ESB_H
#define ESB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESB_MAXT 12          /* atom types are 1..11 (run_single_shoebox.py:23-33); index 0 unused */
#define ESB_MAX_TABLES 48
#define ESB_N_SHELL 50       /* cluster_r = linspace(0.65, 2.5, 50) (run_single_shoebox.py:425)    */
#define ESB_N_COUNTERS 12
#define ESB_FRAME_DOUBLES 9  /* geneTrack.txt columns (run_single_shoebox.py:1496)                   */

enum {
    ESB_OK = 0,
    ESB_EINVAL = -1,   /* bad argument */
    ESB_ECUDA = -2,    /* CUDA runtime error (message in esb_last_error) */
    ESB_ENODEV = -3,   /* no CUDA device / wrong architecture: there is NO CPU fallback */
    ESB_ESTATE = -4,   /* call order violated (e.g. run before tables were set) */
    ESB_ENOMEM = -5
};

/* per-replica status bits (esb_get_status).  LAMMPS aborts the run in each of these cases
 * (README.md:203-205; run_parallel_shoebox.sh:1-14 then deletes the run folder and retries);
 * here the replica stops advancing and the host decides. */
enum {
    ESB_ST_WALL = 1,       /* "Particle on or inside fix wall surface" / lost atom (boundary f f f) */
    ESB_ST_INNER = 2,      /* "Pair distance < table inner cutoff" */
    ESB_ST_NONFINITE = 4,
    ESB_ST_NEIGH = 8       /* neighbour list overflow (LAMMPS: neigh_modify one/page exceeded) */
};

typedef struct esb_handle esb_handle;

/* Closed-form description of one LJsoft table section (ljsoft_potential.py:10-27, 81-86):
 * F(r) = 24 eps_eff r^5/s^6 (2/x^3 - 1/x^2), x = c + (r/s)^6 for r < rc, else 0.
 * valid = 0: the table is only known as arrays (any user-supplied table file). */
typedef struct {
    int valid;
    double eps_eff, s, rc, c;
} esb_closed_form;

/* One Python timestep ("chunk") of the drivers' main loop (run_single_shoebox.py:759-899). */
typedef struct {
    int chunk_index;        /* i                                                             */
    int n_steps;            /* tRun = 200                                                    */
    int pair_mode;          /* 1 = pair_style soft, 2 = pair_style table                     */
    int map_id;             /* which esb_set_table_map() map applies (pair_mode 2)           */
    int ramp_on;            /* `run N start S stop E` given (soft phase, :895-896)           */
    int64_t ramp_start, ramp_stop;
    int adapt_soft;         /* fix s1: pair soft a ramp(0,60), every step (:627-628)         */
    int adapt_bond1;        /* fix s2: bond r0 type 1 ramp(0.033,1.0), every 10 (:635-636)   */
    int adapt_bond2;        /* fix s3: bond r0 type 2 ramp(0.033,0.3), every 10 (:638-639)   */
    int observe;            /* 1: measurements + gene state machine + frame record          */
    int expected_rnp;       /* int(f_rnp*len(rbprnp_atoms)) for this chunk, -1: no ramp (:1457-1460) */
    int induce_on;          /* i >= t_induction_on and i < t_induction_off                    (:1304)      */
    int activate_on;        /* i >= t_activation_on and i < t_activation_off (Flavopiridol: off from NRuns, :885-886, 1322) */
    int inactivate_on;      /* i >= t_activation_on                                           (:1380)      */
    int microscopy;         /* k > 0: record the synthetic-microscopy voxel counts of this chunk as frame k-1
                             * ((i+1) % 10 == 0 and make_microscopy, run_single_shoebox.py:1498-1502) */
    int actin_flags;        /* actin events of this chunk (run_single_shoebox.py:969-1298): 1 = nucleation + plus-end
                             * polymerisation ((i+1) >= t_polymerization_on), 2 = minus-end polymerisation
                             * (i % (t_on_plus*t_on_minus) == 0), 4 = minus-end depolymerisation (i % t_off_minus == 0),
                             * 8 = plus-end depolymerisation (i % t_off_plus == 0, t_off_plus = 2 NRuns: :496, 1220-1250) */
} esb_chunk_desc;

/* Actin polymerisation dynamics (run_single_shoebox.py:398-423, 489-496, 969-1298).  Actin beads are the atoms
 * [first, first+count) (0-based, all of type 3); bonds/angles of type 2 among them given to esb_set_topology()
 * are the initial filaments (make_input_file.py:83-87; the lower atom id is the plus end, :540-553).  With this
 * call bonds, angles and the 1-2/1-3/1-4 exclusions of the actin beads become per-replica state that the
 * per-chunk events rewrite (create_bonds single/bond 2, single/angle 2, delete_bonds ... remove special). */
typedef struct {
    int first, count;
    double p_nucleation;        /* -n                                                   (:300, 981)  */
    double p_on;                /* pon = 0.3                                            (:496)       */
    double monomer_radius;      /* monomerEffectRadius = 2                              (:418)       */
    double end_radius;          /* filEndEffectRadius = 3                               (:417)       */
    double shell_lo, shell_hi;  /* sig_actin - w_sa, sig_actin + w_sa                   (:428-429, 1001-1003) */
    double adj_min;             /* 2**(1/2) * sig_actin: distance to the adjacent bead  (:1074)      */
    double free_box;            /* 1: half edge of the cube a released monomer is placed in (:1186)  */
    double free_min;            /* sig_actin: its minimal distance to the old neighbour (:1188)      */
    int max_tries;              /* 2000                                                 (:1004)      */
} esb_actin_params;
int esb_set_actin(esb_handle *h, const esb_actin_params *p);
/* makeFakeMicroscopyImages (run_single_shoebox.py:152-213; run_single_GENE_STAGES_microscopyModif.py:301-375): the five
 * np.histogramdd(positions, bins=(x_edges, y_edges, z_edges)) voxel counts -- Ser5P {8}, Ser2P {7,11}, chromatin
 * {1,2,6,7,9,10,11}, gene {2,6,7,10,11}, actin {3} -- of every chunk whose descriptor asks for it, fused on the device.
 * Edges as numpy passes them (np.linspace(-11, 11, 22) ...); n_frames = frames to keep per replica.  The Gaussian blur
 * and the image files stay on the host.  esb_get_microscopy: out [R][n_frames][5][nx][ny][nz] counts. */
int esb_set_microscopy(esb_handle *h, int n_x_edges, const double *x_edges, int n_y_edges, const double *y_edges,
                       int n_z_edges, const double *z_edges, int n_frames);
int esb_get_microscopy(esb_handle *h, int frame_begin, int n_frames, unsigned short *out);
/* Filament bookkeeping of every replica after the last run: plus[R][count], minus[R][count] = neighbour of each
 * actin bead towards the plus / minus end (atom id, -1 none); n_free[R], free_list[R][count] = free_actin in list
 * order; n_fil[R], fil_plus[R][count] = plus-end atom of every filament in filament_details order. */
int esb_get_actin(esb_handle *h, int *plus, int *minus, int *n_free, int *free_list, int *n_fil, int *fil_plus);
/* actinTrack.txt / actinAroundCluster.txt rows (:1523-1544): stats [R][n_chunks][6] = n_filaments, sum of lengths,
 * sum of squared lengths, n_free_actin, min and max length; hist [R][n_chunks][49] or NULL. */
int esb_get_actin_frames(esb_handle *h, int chunk_begin, int n_chunks, int *stats, int *hist);

const char *esb_last_error(void);
int esb_version(void);

/* lammps() + read_data (run_single_shoebox.py:323, 529-531).  n_atoms per replica. */
int esb_create(int device, int n_replicas, int n_atoms, int n_types, esb_handle **out);
int esb_destroy(esb_handle *h);                                   /* lammps.close()                  */

/* box bounds from the data file header (make_input_file.py:197-199) */
int esb_set_box(esb_handle *h, const double lo[3], const double hi[3]);
/* Bonds / Angles sections (make_input_file.py:283-328) as (type,i,j) / (type,i,j,k), 0-based atoms;
 * frozen = the Anchors group + fix setforce 0 0 0 (run_single_shoebox.py:651-671).
 * 1-2/1-3/1-4 pair exclusions are derived from the bonds (LAMMPS default special_bonds). */
int esb_set_topology(esb_handle *h, int n_bonds, const int *bonds, int n_angles, const int *angles,
                     int n_frozen, const int *frozen);
/* bond_coeff / angle_coeff (run_single_shoebox.py:583-584, 632-633); arrays indexed by type 1..2 */
int esb_set_coeffs(esb_handle *h, const double bond_k[3], const double bond_r0[3], const double angle_k[3]);
/* timestep, fix langevin T T damp, fix wall/harmonic eps cutoff (run_single_shoebox.py:643-647, 680).
 * langevin_mode: 1 drag+noise (reference), 2 drag only (zero-noise parity runs), 0 off. */
int esb_set_fixes(esb_handle *h, double dt, double t_target, double t_damp, int langevin_mode,
                  double wall_eps, double wall_cut);
/* neighbor SKIN multi; neigh_modify every 1 delay DELAY check yes (run_single_shoebox.py:532-533: 0.3, 10).
 * A list build happens when DELAY steps have passed since the last one and a bead has moved > SKIN/2;
 * every run command builds at set-up.  delay 0 = build as soon as a bead has moved > SKIN/2. */
int esb_set_neighbor(esb_handle *h, double skin, int delay);
/* pair_style soft + pair_coeff block (run_single_shoebox.py:589-625): symmetric MAXT x MAXT cutoffs */
int esb_set_soft(esb_handle *h, const double *cut);
/* pair_style table lookup N + the distinct `pair_coeff I J file KEY cut` tables
 * (run_single_shoebox.py:800-874).  f,e: [n_tables][tlm1] LOOKUP arrays (bin-midpoint F/r and E),
 * innersq/delta/cut per table; cf: optional closed forms (NULL or valid=0 -> array lookups). */
int esb_set_tables(esb_handle *h, int n_tables, int tlm1, const double *f, const double *e,
                   const double *innersq, const double *delta, const double *cut,
                   const esb_closed_form *cf);
/* What LAMMPS does inside `pair_coeff I J file KEY cut` for pair_style table lookup N
 * ([LAMMPS-ext] PairTable::read_table RSQ re-grid, spline_table, compute_table): file columns
 * E,F of an `N n RSQ rlo rhi` section -> tablength-1 bin-midpoint values e_out, f_out (F/r);
 * params_out = {innersq, delta, invdelta}.  Pure host code (set-up, not the hot path). */
int esb_table_from_section(int n_input, double rlo, double rhi, const double *e_file, const double *f_file,
                           double cut, int tablength, double *e_out, double *f_out, double *params_out);
/* effective (type_i,type_j) -> table index after replaying a pair_coeff stream; map_id in [0,4) */
int esb_set_table_map(esb_handle *h, int map_id, const int *map);

/* index sets the drivers derive after read_data (run_single_shoebox.py:683-704, 941-946) and
 * cluster_r (:425): gene start bead, promoter length, enhancer beads, Ser5P beads (0-based). */
int esb_set_layout(esb_handle *h, int gene_start, int promoter_length, int n_enhancer,
                   const int *enhancer, int n_ser5p, const int *ser5p, const double *cluster_r);
/* per replica: Langevin/decision seed, -x ser5p_to_activate, -a/1000 p_gene_activation
 * (run_single_shoebox.py:282-284, 644) */
int esb_set_replica_params(esb_handle *h, const uint64_t *seed, const int *threshold, const double *p_activation);

/* atom coordinates / velocities / types per replica, ID order: x,v [R][N][3] doubles, types [R][N].
 * replicate != 0: the arrays hold ONE replica that is copied to all.  (read_data Atoms/Velocities;
 * `set atom ID type T`, run_single_shoebox.py:1308.) */
int esb_upload_state(esb_handle *h, const double *x, const double *v, const int *types, int replicate);
/* lmp.gather_atoms('x',1,3) / ('type',0,1) (run_single_shoebox.py:913-914), for all replicas */
int esb_download_state(esb_handle *h, double *x, double *v, int *types);
/* fp32 fast path of the two calls above with caller-pinned buffers on `stream`:
 * pos4/vel4 = [R][n_pad][4] floats (x,y,z,type-as-int-bits / vx,vy,vz,0). */
int esb_upload_state_f32(esb_handle *h, const float *pos4, const float *vel4, void *stream);
int esb_download_state_f32(esb_handle *h, float *pos4, float *vel4, void *stream);
int esb_n_pad(const esb_handle *h);

/* the per-chunk programme of the drivers' main loop; copied to the device */
int esb_set_schedule(esb_handle *h, int n_chunks, const esb_chunk_desc *descs);
/* L.run(tRun) x n_chunks with everything the drivers do in between fused on the device
 * (run_single_shoebox.py:895-899, 913-961, 1304-1400, 1457-1519).  Asynchronous on `stream`
 * (a cudaStream_t, may be NULL). */
int esb_run(esb_handle *h, int chunk_begin, int n_chunks, void *stream);
int esb_sync(esb_handle *h);
/* Scheduling hint for launches that SHARE the device with the launch of another handle (two half batches on two
 * streams, so that one half's host copies overlap the other half's steps): every chunk is cut into n_seg consecutive
 * work items (same arithmetic, same bits) so that the two persistent kernels interleave at that grain instead of one
 * waiting for whole chunks of the other; co_scheduled != 0 keeps the two-CTAs-per-SM kernel even for a batch that would
 * fit one CTA per SM.  n_seg = 0, co_scheduled = 0: the defaults (the launcher decides). */
int esb_set_segments(esb_handle *h, int n_seg, int co_scheduled);

/* Parity hook: forces [R][N][3] and energies [R][4] = (pair, bond, angle, wall) at the current
 * positions.  pair_mode/map_id as in esb_chunk_desc; soft_a: prefactor A for pair_mode 1.
 * with_post != 0 adds langevin (noise evaluation index `eval`), walls and setforce.
 * use_arrays != 0 forces LOOKUP-array evaluation even when closed forms are present. */
int esb_forces(esb_handle *h, int pair_mode, int map_id, double soft_a, int with_post, uint64_t eval,
               int use_arrays, double *f_out, double *e_out);

/* geneTrack records: out [R][n_chunks][9] doubles (t, probe xyz*60, d_rp*60, d_rg*60, n_s5p_prom,
 * n_s5p_gene, state) (run_single_shoebox.py:1496); hist [R][n_chunks][49] ints or NULL (:1504-1519) */
int esb_get_frames(esb_handle *h, int chunk_begin, int n_chunks, double *out, int *hist);
/* d_rp, d_rg of the same frames in bead units, out [R][n_chunks][2]: gene_stats.txt tests d_rg < 3
 * before scaling by 60 nm (run_single_shoebox.py:1629-1642) */
int esb_get_raw_distances(esb_handle *h, int chunk_begin, int n_chunks, double *out);
/* same records, asynchronously on `stream` into a caller-pinned buffer (no histograms) */
int esb_get_frames_async(esb_handle *h, int chunk_begin, int n_chunks, double *out, void *stream);
int esb_get_status(esb_handle *h, int *status);                   /* [R] ESB_ST_* bits               */
/* [R][ESB_N_COUNTERS]: force evaluations, neighbour builds, in-cutoff (ordered) pairs, listed (ordered)
 * pairs summed over evaluations, steps, total_active_steps, then SM cycles spent in the per-atom
 * phase, list builds, force phase, set-up/observation, and the number of "dangerous builds" (a list
 * build that was already due at the first step `neigh_modify delay` permits: what LAMMPS prints at the end
 * of a run), and the largest pair-force component any bead has seen, in units of 2^-8 (the pair forces are
 * accumulated in fixed point: this is the headroom diagnostic) */
int esb_get_counters(esb_handle *h, int64_t *out);
int esb_reset_counters(esb_handle *h);
/* current bond r0 (after fix adapt) and MD step per handle */
int esb_get_bond_r0(esb_handle *h, double r0[3]);
int64_t esb_get_step(const esb_handle *h);

/* raw device pointers for tensor hand-off: which = 0 pos4 [R][n_pad] float4, 1 vel4, 2 frames,
 * 3 status, 4 counters */
void *esb_device_ptr(esb_handle *h, int which);

/* dist_genestages_grouped.py (VGM/dist_genestages_grouped.py:70-106) on the device.
 * frames: DEVICE or host pointer [n_runs][n_frames][9] doubles (geneTrack columns), group = 5.
 * out: host [n_runs/group][4] = S5PInt, S2PInt, Contact, DistActivation (NaN where the reference
 * takes the mean of an empty list).  frames_on_device selects the pointer kind. */
int esb_aggregate_grouped(int device, const double *frames, int frames_on_device, int n_runs, int n_frames,
                          int group, double contact_dist, int t_skip, double *out, void *stream);

/* Per-run rows of VGM/dist_genestages_grouped_5percentile_10xaveraging.py:78-89 (and of
 * dist_genestages_all_5percentile.py): out [n_runs][6] = mean s5p_promoter, S2PInt, Contact, mean
 * activation distance (NaN if none), np.percentile(d_rg[t_skip:], q), np.percentile(d_rp[t_skip:], q).
 * prev_index/gamma: lower order statistic and weight of numpy's 'linear' percentile for that q and
 * length (computed by the caller with numpy's own expression). */
int esb_aggregate_runs(int device, const double *frames, int frames_on_device, int n_runs, int n_frames,
                       double contact_dist, int t_skip, int prev_index, double gamma, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
