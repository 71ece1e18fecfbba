/*
 * esb_oracle.c -- CPU (fp64) restatement of the EnhancerShoeBox simulation hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * path (enhancershoebox_b200/) never does and fails loudly without its CUDA extension.
 *
 * PARITY UNPINNED for the MD part: the arithmetic of the reference hot path lives in
 * LAMMPS (pinned by prose only: "stable_23Jun2022_update2", /root/reference/README.md:59),
 * which is not vendored in /root/reference, is not installed here and cannot be fetched.
 * The reference holds no test, golden trajectory, force dump or thermo log for this path.
 * What follows restates the published LAMMPS algorithms for exactly the commands the
 * reference issues (call sites cited per function).  The pieces that CAN be pinned are
 * pinned in tests/: the pair-table file contents (against the reference's own generator,
 * ljsoft_potential.py, run in this container -> tests/golden/), and the aggregation tables
 * (against dist_genestages_grouped.py run in this container).
 *
 * Command stream restated (VGM = /root/reference/VisitorGeneMaps):
 *   atom_style angle / boundary f f f / read_data        run_single_shoebox.py:529-531
 *   pair_style soft + fix adapt (A ramp)                  run_single_shoebox.py:589-628
 *   pair_style table lookup 30000 + pair_coeff            run_single_shoebox.py:800-874,
 *                                                         VGM/run_single_GENES_STAGES.py:743-791
 *   bond_style harmonic (+ fix adapt r0 ramps)            run_single_shoebox.py:631-639
 *   angle_style cosine                                    run_single_shoebox.py:582-584
 *   fix nve, fix langevin 1 1 0.5 seed                    run_single_shoebox.py:643-644
 *   fix wall/harmonic x6                                  run_single_shoebox.py:647
 *   fix setforce 0 0 0 on the anchors                     run_single_shoebox.py:651-671
 *   run 200 [start 0 stop 2000]                           run_single_shoebox.py:895-899
 *
 * One deliberate, documented departure: LAMMPS draws Langevin noise from a per-run
 * Marsaglia stream in local-atom order, which no independent implementation can reproduce
 * (SURVEY.md section 8 a10).  The noise here is uniform(-0.5,0.5) exactly as LAMMPS's default
 * (variance-matched uniform, not Gaussian) but drawn from Philox4x32-10 keyed by
 * (replica seed; atom, force-evaluation index) so that the CUDA path and this oracle see
 * identical noise and short noisy trajectories can be compared step by step.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ESBO_MAXT 12 /* atom types are 1..11 (run_single_shoebox.py:23-33) */

/* ------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11).  Same function is implemented on the device.    */
/* ------------------------------------------------------------------------------------ */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0];
    const uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

void esbo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; r++) {
        if (r) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
        philox_round(c, k);
    }
    memcpy(out, c, sizeof(c));
}

/* u32 -> uniform in [-0.5, 0.5]: the bits are read as a signed int, rounded to fp32 and
 * scaled by 2^-32.  Rounding through fp32 is part of the definition (device does the same). */
static inline double u32_to_centered(uint32_t u)
{
    return (double)((float)(int32_t)u * 0x1p-32f);
}
/* u32 -> uniform in [0,1) with 24 bits, used for the chunk-level (gene / RNP) draws. */
double esbo_u32_to_unit(uint32_t u) { return (double)(u >> 8) * 0x1p-24; }

/* ------------------------------------------------------------------------------------ */
/* pair_style table: file section -> LOOKUP arrays                                        */
/* [LAMMPS-ext] PairTable::read_table (RSQ re-grid), spline_table, compute_table, spline, */
/* splint.  Reference call sites: run_single_shoebox.py:800-874; file format written by   */
/* ljsoft_potential.py:95-108 ("N 30000 RSQ rlo rhi").                                    */
/* ------------------------------------------------------------------------------------ */
static void nr_spline(const double *x, const double *y, int n, double yp1, double ypn, double *y2)
{
    double *u = (double *)malloc(sizeof(double) * (size_t)n);
    if (yp1 > 0.99e30) { y2[0] = 0.0; u[0] = 0.0; }
    else {
        y2[0] = -0.5;
        u[0] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1);
    }
    for (int i = 1; i < n - 1; i++) {
        const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
        const double p = sig * y2[i - 1] + 2.0;
        y2[i] = (sig - 1.0) / p;
        double t = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
        u[i] = (6.0 * t / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
    }
    double qn, un;
    if (ypn > 0.99e30) { qn = 0.0; un = 0.0; }
    else {
        qn = 0.5;
        un = (3.0 / (x[n - 1] - x[n - 2])) * (ypn - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
    }
    y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.0);
    for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
    free(u);
}

static double nr_splint(const double *xa, const double *ya, const double *y2a, int n, double x)
{
    int klo = 0, khi = n - 1;
    while (khi - klo > 1) {
        const int k = (khi + klo) >> 1;
        if (xa[k] > x) khi = k; else klo = k;
    }
    const double h = xa[khi] - xa[klo];
    const double a = (xa[khi] - x) / h;
    const double b = (x - xa[klo]) / h;
    return a * ya[klo] + b * ya[khi] +
           ((a * a * a - a) * y2a[klo] + (b * b * b - b) * y2a[khi]) * (h * h) / 6.0;
}

/* Build the LOOKUP arrays for one `pair_coeff I J file KEY cut`.
 * efile/ffile: the ninput E,F columns of section KEY; rlo,rhi: the RSQ header bounds.
 * tablength: N of `pair_style table lookup N`.  Outputs e_out/f_out hold tablength-1 bin
 * midpoint values (f as F/r); params_out = {innersq, delta, invdelta}.
 * Returns 0, or -1 if cut is outside (rlo, rhi] ("Pair table cutoff outside of table"). */
int esbo_table_build(int ninput, double rlo, double rhi, const double *efile, const double *ffile,
                     double cut, int tablength, double *e_out, double *f_out, double *params_out)
{
    if (ninput <= 1 || rlo <= 0.0 || cut <= rlo || cut > rhi) return -1;
    double *r = (double *)malloc(sizeof(double) * (size_t)ninput * 3);
    double *e2 = r + ninput, *f2 = r + 2 * ninput;
    for (int i = 0; i < ninput; i++) {
        double rnew = rlo * rlo + (rhi * rhi - rlo * rlo) * i / (ninput - 1);
        r[i] = sqrt(rnew);
    }
    nr_spline(r, efile, ninput, -ffile[0], -ffile[ninput - 1], e2);
    const double fplo = (ffile[1] - ffile[0]) / (r[1] - r[0]);
    const double fphi = (ffile[ninput - 1] - ffile[ninput - 2]) / (r[ninput - 1] - r[ninput - 2]);
    nr_spline(r, ffile, ninput, fplo, fphi, f2);

    const int tlm1 = tablength - 1;
    const double innersq = rlo * rlo;
    const double delta = (cut * cut - innersq) / tlm1;
    for (int i = 0; i < tlm1; i++) {
        const double rsq = innersq + (i + 0.5) * delta;
        const double rr = sqrt(rsq);
        e_out[i] = nr_splint(r, efile, e2, ninput, rr);
        f_out[i] = nr_splint(r, ffile, f2, ninput, rr) / rr;
    }
    params_out[0] = innersq;
    params_out[1] = delta;
    params_out[2] = 1.0 / delta;
    free(r);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* System                                                                                 */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int kind;      /* 0 = pair soft prefactor A (all pairs), 1 = bond r0 of bond type `index` */
    int index;
    int nevery;
    double v0, v1; /* ramp(v0, v1) */
} esbo_adapt;

typedef struct esbo_system {
    int n, ntypes;
    double lo[3], hi[3];
    double *x, *v, *f;
    double *x_build; /* positions at the last neighbour build */
    int *type;
    unsigned char *frozen;
    int nbonds, *bond;   /* (btype, i, j) */
    int nangles, *angle; /* (atype, i, j, k); j is the vertex */
    int *spec_start, *spec; /* CSR: partners excluded from pair interactions (1-2,1-3,1-4) */
    /* integrator / fixes */
    double dt, t_period, t_target, wall_eps, wall_cut;
    int langevin_on; /* 1: drag + noise, 0: fix langevin absent, 2: drag only (zero-noise runs) */
    double bond_k[4], bond_r0[4], angle_k[4];
    /* pair */
    int pair_mode; /* 0 none, 1 soft, 2 table */
    double soft_a, soft_cut[ESBO_MAXT * ESBO_MAXT];
    int ntables, tlm1;
    double *tab_f, *tab_e, *tab_innersq, *tab_delta, *tab_invdelta, *tab_cut;
    int tabmap[ESBO_MAXT * ESBO_MAXT];
    /* adapt */
    int nadapt;
    esbo_adapt adapt[4];
    /* neighbour list (half, newton on) */
    double skin;
    int delay, ago;              /* neigh_modify delay N; steps since the last build */
    int *nl_start, *nl, nl_cap;
    int nl_valid;
    /* rng */
    uint32_t key[2];
    uint64_t eval_counter;
    long long step;
    /* energies of the last force evaluation: pair, bond, angle, wall */
    double energy[4];
    /* counters */
    long long n_evals, n_builds, n_pairs_in_cut, n_pairs_listed, n_dangerous;
    int status; /* 0 ok, 1 particle on/inside wall, 2 lost atom, 3 pair distance < table inner, 4 nonfinite */
} esbo_system;

esbo_system *esbo_create(int n, int ntypes)
{
    esbo_system *s = (esbo_system *)calloc(1, sizeof(esbo_system));
    s->n = n; s->ntypes = ntypes;
    s->x = (double *)calloc((size_t)n * 3, sizeof(double));
    s->v = (double *)calloc((size_t)n * 3, sizeof(double));
    s->f = (double *)calloc((size_t)n * 3, sizeof(double));
    s->x_build = (double *)calloc((size_t)n * 3, sizeof(double));
    s->type = (int *)calloc((size_t)n, sizeof(int));
    s->frozen = (unsigned char *)calloc((size_t)n, 1);
    s->spec_start = (int *)calloc((size_t)n + 1, sizeof(int));
    s->nl_start = (int *)calloc((size_t)n + 1, sizeof(int));
    s->dt = 0.005; s->t_period = 0.5; s->t_target = 1.0;
    s->wall_eps = 100.0; s->wall_cut = 3.0;
    s->langevin_on = 1;
    s->skin = 0.3;
    s->delay = 10;
    return s;
}

void esbo_destroy(esbo_system *s)
{
    if (!s) return;
    free(s->x); free(s->v); free(s->f); free(s->x_build); free(s->type); free(s->frozen);
    free(s->bond); free(s->angle); free(s->spec_start); free(s->spec);
    free(s->tab_f); free(s->tab_e); free(s->tab_innersq); free(s->tab_delta);
    free(s->tab_invdelta); free(s->tab_cut); free(s->nl_start); free(s->nl);
    free(s);
}

void esbo_set_box(esbo_system *s, const double *lo, const double *hi)
{
    for (int d = 0; d < 3; d++) { s->lo[d] = lo[d]; s->hi[d] = hi[d]; }
}
void esbo_set_atoms(esbo_system *s, const int *type, const double *x, const double *v)
{
    memcpy(s->type, type, sizeof(int) * (size_t)s->n);
    memcpy(s->x, x, sizeof(double) * (size_t)s->n * 3);
    if (v) memcpy(s->v, v, sizeof(double) * (size_t)s->n * 3);
    else memset(s->v, 0, sizeof(double) * (size_t)s->n * 3);
    s->nl_valid = 0;
}
void esbo_set_types(esbo_system *s, const int *type)
{
    memcpy(s->type, type, sizeof(int) * (size_t)s->n);
    s->nl_valid = 0;
}
void esbo_get_state(const esbo_system *s, double *x, double *v, double *f, int *type)
{
    if (x) memcpy(x, s->x, sizeof(double) * (size_t)s->n * 3);
    if (v) memcpy(v, s->v, sizeof(double) * (size_t)s->n * 3);
    if (f) memcpy(f, s->f, sizeof(double) * (size_t)s->n * 3);
    if (type) memcpy(type, s->type, sizeof(int) * (size_t)s->n);
}
void esbo_set_frozen(esbo_system *s, int nf, const int *idx)
{
    memset(s->frozen, 0, (size_t)s->n);
    for (int k = 0; k < nf; k++) s->frozen[idx[k]] = 1;
}

/* Bonds/angles + the special (exclusion) lists they imply.
 * [LAMMPS-ext] default `special_bonds lj 0.0 0.0 0.0` (no special_bonds command is issued;
 * VGM/run_single_GENES_STAGES.py:609 is commented out): atoms 1, 2 or 3 bonds apart do not
 * interact through the pair style. */
void esbo_set_topology(esbo_system *s, int nbonds, const int *bond, int nangles, const int *angle)
{
    free(s->bond); free(s->angle); free(s->spec);
    s->nbonds = nbonds; s->nangles = nangles;
    s->bond = (int *)malloc(sizeof(int) * 3 * (size_t)(nbonds > 0 ? nbonds : 1));
    s->angle = (int *)malloc(sizeof(int) * 4 * (size_t)(nangles > 0 ? nangles : 1));
    memcpy(s->bond, bond, sizeof(int) * 3 * (size_t)nbonds);
    memcpy(s->angle, angle, sizeof(int) * 4 * (size_t)nangles);
    const int n = s->n;
    /* adjacency (CSR) */
    int *deg = (int *)calloc((size_t)n + 1, sizeof(int));
    for (int b = 0; b < nbonds; b++) { deg[bond[3 * b + 1] + 1]++; deg[bond[3 * b + 2] + 1]++; }
    for (int i = 0; i < n; i++) deg[i + 1] += deg[i];
    int *adj = (int *)malloc(sizeof(int) * (size_t)(2 * nbonds + 1));
    int *fill = (int *)calloc((size_t)n, sizeof(int));
    for (int b = 0; b < nbonds; b++) {
        const int i = bond[3 * b + 1], j = bond[3 * b + 2];
        adj[deg[i] + fill[i]++] = j;
        adj[deg[j] + fill[j]++] = i;
    }
    /* BFS to depth 3 from every atom that has bonds */
    int cap = 16 * nbonds + 16, cnt = 0;
    s->spec = (int *)malloc(sizeof(int) * (size_t)cap);
    int *mark = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) mark[i] = -1;
    int *frontier = (int *)malloc(sizeof(int) * (size_t)n * 2);
    for (int i = 0; i < n; i++) {
        s->spec_start[i] = cnt;
        if (deg[i + 1] == deg[i]) continue;
        mark[i] = i;
        int *cur = frontier, *nxt = frontier + n, ncur = 1;
        cur[0] = i;
        for (int depth = 0; depth < 3; depth++) {
            int nn = 0;
            for (int q = 0; q < ncur; q++) {
                const int a = cur[q];
                for (int e = deg[a]; e < deg[a + 1]; e++) {
                    const int b = adj[e];
                    if (mark[b] == i) continue;
                    mark[b] = i;
                    nxt[nn++] = b;
                    if (cnt >= cap) { cap *= 2; s->spec = (int *)realloc(s->spec, sizeof(int) * (size_t)cap); }
                    s->spec[cnt++] = b;
                }
            }
            int *t = cur; cur = nxt; nxt = t; ncur = nn;
        }
    }
    s->spec_start[n] = cnt;
    free(deg); free(adj); free(fill); free(mark); free(frontier);
    s->nl_valid = 0;
}
int esbo_get_special(const esbo_system *s, int *start_out, int *spec_out)
{
    if (start_out) memcpy(start_out, s->spec_start, sizeof(int) * (size_t)(s->n + 1));
    if (spec_out) memcpy(spec_out, s->spec, sizeof(int) * (size_t)s->spec_start[s->n]);
    return s->spec_start[s->n];
}

void esbo_set_bond_coeff(esbo_system *s, int btype, double k, double r0) { s->bond_k[btype] = k; s->bond_r0[btype] = r0; }
void esbo_set_angle_coeff(esbo_system *s, int atype, double k) { s->angle_k[atype] = k; }
double esbo_get_bond_r0(const esbo_system *s, int btype) { return s->bond_r0[btype]; }
double esbo_get_soft_a(const esbo_system *s) { return s->soft_a; }
void esbo_set_fixes(esbo_system *s, double dt, double t_target, double t_period, int langevin_on,
                    double wall_eps, double wall_cut)
{
    s->dt = dt; s->t_target = t_target; s->t_period = t_period; s->langevin_on = langevin_on;
    s->wall_eps = wall_eps; s->wall_cut = wall_cut;
}
void esbo_set_seed(esbo_system *s, uint64_t seed, uint64_t eval_counter)
{
    s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
    s->eval_counter = eval_counter;
}
void esbo_set_skin(esbo_system *s, double skin) { s->skin = skin; s->nl_valid = 0; }
/* `neigh_modify every 1 delay N check yes` (run_single_shoebox.py:533, N = 10) */
void esbo_set_neigh_delay(esbo_system *s, int delay) { s->delay = delay < 0 ? 0 : delay; s->nl_valid = 0; }

/* pair_style soft rc_global; pair_coeff I J A rc (run_single_shoebox.py:589-625).
 * cut is an ESBO_MAXT x ESBO_MAXT symmetric matrix indexed by (type_i, type_j), types 1-based. */
void esbo_pair_soft(esbo_system *s, double a, const double *cut)
{
    s->pair_mode = 1; s->soft_a = a;
    memcpy(s->soft_cut, cut, sizeof(s->soft_cut));
    s->nl_valid = 0;
}
/* pair_style table lookup N with the effective (I,J) -> table map after replaying the
 * pair_coeff stream (later commands override earlier ones). */
void esbo_pair_table(esbo_system *s, int ntables, int tlm1, const double *f, const double *e,
                     const double *innersq, const double *delta, const double *cut, const int *tabmap)
{
    free(s->tab_f); free(s->tab_e); free(s->tab_innersq); free(s->tab_delta);
    free(s->tab_invdelta); free(s->tab_cut);
    s->pair_mode = 2; s->ntables = ntables; s->tlm1 = tlm1;
    const size_t sz = (size_t)ntables * (size_t)tlm1;
    s->tab_f = (double *)malloc(sizeof(double) * sz);
    s->tab_e = (double *)malloc(sizeof(double) * sz);
    memcpy(s->tab_f, f, sizeof(double) * sz);
    memcpy(s->tab_e, e, sizeof(double) * sz);
    s->tab_innersq = (double *)malloc(sizeof(double) * (size_t)ntables);
    s->tab_delta = (double *)malloc(sizeof(double) * (size_t)ntables);
    s->tab_invdelta = (double *)malloc(sizeof(double) * (size_t)ntables);
    s->tab_cut = (double *)malloc(sizeof(double) * (size_t)ntables);
    for (int t = 0; t < ntables; t++) {
        s->tab_innersq[t] = innersq[t]; s->tab_delta[t] = delta[t];
        s->tab_invdelta[t] = 1.0 / delta[t]; s->tab_cut[t] = cut[t];
    }
    memcpy(s->tabmap, tabmap, sizeof(s->tabmap));
    s->nl_valid = 0;
}
/* Only the (I,J)->table map changes (later pair_coeff commands, e.g. Ser5P switch-on). */
void esbo_pair_table_remap(esbo_system *s, const int *tabmap)
{
    memcpy(s->tabmap, tabmap, sizeof(s->tabmap));
    s->nl_valid = 0;
}

/* fix ID all adapt N pair soft a * * v_ramp  /  fix ID all adapt N bond harmonic r0 T v_ramp
 * (run_single_shoebox.py:627-639).  [LAMMPS-ext] default `reset no`: the last value sticks. */
void esbo_clear_adapt(esbo_system *s) { s->nadapt = 0; }
void esbo_add_adapt(esbo_system *s, int kind, int index, int nevery, double v0, double v1)
{
    esbo_adapt *a = &s->adapt[s->nadapt++];
    a->kind = kind; a->index = index; a->nevery = nevery; a->v0 = v0; a->v1 = v1;
}

static inline double pair_cut(const esbo_system *s, int ti, int tj)
{
    if (s->pair_mode == 1) return s->soft_cut[ti * ESBO_MAXT + tj];
    if (s->pair_mode == 2) return s->tab_cut[s->tabmap[ti * ESBO_MAXT + tj]];
    return 0.0;
}

/* ------------------------------------------------------------------------------------ */
/* Neighbour list: binned build, half list, per-type-pair cutoff + skin                   */
/* (`neighbor 0.3 multi`, run_single_shoebox.py:532; the list only has to be complete:    */
/* it does not change the forces.)                                                        */
/* ------------------------------------------------------------------------------------ */
static void build_neighbors(esbo_system *s)
{
    const int n = s->n;
    double cutmax_t[ESBO_MAXT] = {0};
    double cutmax = 0.0;
    for (int a = 1; a <= s->ntypes; a++)
        for (int b = 1; b <= s->ntypes; b++) {
            const double c = pair_cut(s, a, b);
            if (c > cutmax_t[a]) cutmax_t[a] = c;
            if (c > cutmax) cutmax = c;
        }
    const double bs = 1.0; /* bin edge */
    int nb[3];
    double blo[3];
    for (int d = 0; d < 3; d++) {
        blo[d] = s->lo[d] - 1.0;
        nb[d] = (int)ceil((s->hi[d] - s->lo[d] + 2.0) / bs);
        if (nb[d] < 1) nb[d] = 1;
    }
    const int nbins = nb[0] * nb[1] * nb[2];
    int *head = (int *)malloc(sizeof(int) * (size_t)nbins);
    int *next = (int *)malloc(sizeof(int) * (size_t)n);
    int *bin3 = (int *)malloc(sizeof(int) * (size_t)n * 3);
    for (int b = 0; b < nbins; b++) head[b] = -1;
    for (int i = n - 1; i >= 0; i--) {
        int c[3];
        for (int d = 0; d < 3; d++) {
            c[d] = (int)floor((s->x[3 * i + d] - blo[d]) / bs);
            if (c[d] < 0) c[d] = 0;
            if (c[d] >= nb[d]) c[d] = nb[d] - 1;
            bin3[3 * i + d] = c[d];
        }
        const int b = (c[2] * nb[1] + c[1]) * nb[0] + c[0];
        next[i] = head[b]; head[b] = i;
    }
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        s->nl_start[i] = cnt;
        const int ti = s->type[i];
        const double rmax = cutmax_t[ti] + s->skin;
        const int rng = (int)ceil(rmax / bs);
        const double xi = s->x[3 * i], yi = s->x[3 * i + 1], zi = s->x[3 * i + 2];
        for (int cz = bin3[3 * i + 2] - rng; cz <= bin3[3 * i + 2] + rng; cz++) {
            if (cz < 0 || cz >= nb[2]) continue;
            for (int cy = bin3[3 * i + 1] - rng; cy <= bin3[3 * i + 1] + rng; cy++) {
                if (cy < 0 || cy >= nb[1]) continue;
                for (int cx = bin3[3 * i] - rng; cx <= bin3[3 * i] + rng; cx++) {
                    if (cx < 0 || cx >= nb[0]) continue;
                    for (int j = head[(cz * nb[1] + cy) * nb[0] + cx]; j >= 0; j = next[j]) {
                        if (j <= i) continue;
                        const double c = pair_cut(s, ti, s->type[j]);
                        if (c <= 0.0) continue;
                        const double dx = xi - s->x[3 * j], dy = yi - s->x[3 * j + 1], dz = zi - s->x[3 * j + 2];
                        const double rsq = dx * dx + dy * dy + dz * dz;
                        const double cs = c + s->skin;
                        if (rsq >= cs * cs) continue;
                        int excluded = 0;
                        for (int e = s->spec_start[i]; e < s->spec_start[i + 1]; e++)
                            if (s->spec[e] == j) { excluded = 1; break; }
                        if (excluded) continue;
                        if (cnt >= s->nl_cap) {
                            s->nl_cap = s->nl_cap ? 2 * s->nl_cap : 65536;
                            s->nl = (int *)realloc(s->nl, sizeof(int) * (size_t)s->nl_cap);
                        }
                        s->nl[cnt++] = j;
                    }
                }
            }
        }
    }
    s->nl_start[n] = cnt;
    memcpy(s->x_build, s->x, sizeof(double) * (size_t)n * 3);
    s->nl_valid = 1;
    s->n_builds++;
    free(head); free(next); free(bin3);
}

/* [LAMMPS-ext] Neighbor::decide + check_distance with `every 1 delay N check yes`: called once per step
 * (ago++), a build happens when at least N steps have passed since the last one AND some atom has moved
 * further than skin/2 from where it was at that build.  A build that was already due at the first permitted
 * step is what LAMMPS reports as a "dangerous build": in the steps before it a pair may have come inside the
 * cutoff without being in the list -- the reference's settings accept that, and so does this restatement.
 * Every `run` rebuilds at set-up (nl_valid = 0). */
static int needs_rebuild(esbo_system *s)
{
    if (!s->nl_valid) { s->ago = 0; return 1; }
    s->ago++;
    if (s->ago < s->delay) return 0;
    const double trig = 0.25 * s->skin * s->skin;
    for (int i = 0; i < s->n; i++) {
        const double dx = s->x[3 * i] - s->x_build[3 * i];
        const double dy = s->x[3 * i + 1] - s->x_build[3 * i + 1];
        const double dz = s->x[3 * i + 2] - s->x_build[3 * i + 2];
        if (dx * dx + dy * dy + dz * dz > trig) {
            if (s->ago == (s->delay > 1 ? s->delay : 1)) s->n_dangerous++;
            s->ago = 0;
            return 1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Force evaluation.  with_post: apply the post_force fixes (langevin -> walls -> setforce) */
/* in the order the reference creates them (run_single_shoebox.py:643-671).               */
/* ------------------------------------------------------------------------------------ */
static void compute_forces(esbo_system *s, int with_post, int with_noise)
{
    const int n = s->n;
    double *x = s->x, *f = s->f;
    memset(f, 0, sizeof(double) * (size_t)n * 3);
    double e_pair = 0.0, e_bond = 0.0, e_angle = 0.0, e_wall = 0.0;

    if (s->pair_mode) {
        if (needs_rebuild(s)) build_neighbors(s);
        s->n_pairs_listed += s->nl_start[n];
        for (int i = 0; i < n; i++) {
            const int ti = s->type[i];
            const double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
            double fxi = 0.0, fyi = 0.0, fzi = 0.0;
            for (int q = s->nl_start[i]; q < s->nl_start[i + 1]; q++) {
                const int j = s->nl[q];
                const int tj = s->type[j];
                const double dx = xi - x[3 * j], dy = yi - x[3 * j + 1], dz = zi - x[3 * j + 2];
                const double rsq = dx * dx + dy * dy + dz * dz;
                double fpair;
                if (s->pair_mode == 1) {
                    /* [LAMMPS-ext] PairSoft::compute */
                    const double c = s->soft_cut[ti * ESBO_MAXT + tj];
                    if (rsq >= c * c) continue;
                    const double r = sqrt(rsq);
                    const double arg = M_PI * r / c;
                    fpair = (r > 0.0) ? s->soft_a * sin(arg) * M_PI / c / r : 0.0;
                    e_pair += s->soft_a * (1.0 + cos(arg));
                } else {
                    /* [LAMMPS-ext] PairTable::compute, tabstyle LOOKUP */
                    const int t = s->tabmap[ti * ESBO_MAXT + tj];
                    const double c = s->tab_cut[t];
                    if (rsq >= c * c) continue;
                    if (rsq < s->tab_innersq[t]) { s->status = 3; continue; }
                    int it = (int)((rsq - s->tab_innersq[t]) * s->tab_invdelta[t]);
                    if (it >= s->tlm1) it = s->tlm1 - 1; /* LAMMPS errors out; unreachable for rsq<cut^2 up to rounding */
                    fpair = s->tab_f[(size_t)t * s->tlm1 + it];
                    e_pair += s->tab_e[(size_t)t * s->tlm1 + it];
                }
                s->n_pairs_in_cut++;
                fxi += dx * fpair; fyi += dy * fpair; fzi += dz * fpair;
                f[3 * j] -= dx * fpair; f[3 * j + 1] -= dy * fpair; f[3 * j + 2] -= dz * fpair;
            }
            f[3 * i] += fxi; f[3 * i + 1] += fyi; f[3 * i + 2] += fzi;
        }
    }

    /* [LAMMPS-ext] BondHarmonic::compute: E = K (r-r0)^2 (bond_coeff at run_single_shoebox.py:632-633) */
    for (int b = 0; b < s->nbonds; b++) {
        const int bt = s->bond[3 * b], i = s->bond[3 * b + 1], j = s->bond[3 * b + 2];
        const double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        const double dr = r - s->bond_r0[bt];
        const double rk = s->bond_k[bt] * dr;
        const double fb = (r > 0.0) ? -2.0 * rk / r : 0.0;
        e_bond += rk * dr;
        f[3 * i] += dx * fb; f[3 * i + 1] += dy * fb; f[3 * i + 2] += dz * fb;
        f[3 * j] -= dx * fb; f[3 * j + 1] -= dy * fb; f[3 * j + 2] -= dz * fb;
    }

    /* [LAMMPS-ext] AngleCosine::compute: E = K (1 + cos theta) (angle_coeff at :583-584) */
    for (int a = 0; a < s->nangles; a++) {
        const int at = s->angle[4 * a], i1 = s->angle[4 * a + 1], i2 = s->angle[4 * a + 2], i3 = s->angle[4 * a + 3];
        const double d1x = x[3 * i1] - x[3 * i2], d1y = x[3 * i1 + 1] - x[3 * i2 + 1], d1z = x[3 * i1 + 2] - x[3 * i2 + 2];
        const double d2x = x[3 * i3] - x[3 * i2], d2y = x[3 * i3 + 1] - x[3 * i2 + 1], d2z = x[3 * i3 + 2] - x[3 * i2 + 2];
        const double rsq1 = d1x * d1x + d1y * d1y + d1z * d1z, rsq2 = d2x * d2x + d2y * d2y + d2z * d2z;
        const double r1 = sqrt(rsq1), r2 = sqrt(rsq2);
        double c = (d1x * d2x + d1y * d2y + d1z * d2z) / (r1 * r2);
        if (c > 1.0) c = 1.0;
        if (c < -1.0) c = -1.0;
        const double k = s->angle_k[at];
        e_angle += k * (1.0 + c);
        const double a11 = k * c / rsq1, a12 = -k / (r1 * r2), a22 = k * c / rsq2;
        const double f1x = a11 * d1x + a12 * d2x, f1y = a11 * d1y + a12 * d2y, f1z = a11 * d1z + a12 * d2z;
        const double f3x = a22 * d2x + a12 * d1x, f3y = a22 * d2y + a12 * d1y, f3z = a22 * d2z + a12 * d1z;
        f[3 * i1] += f1x; f[3 * i1 + 1] += f1y; f[3 * i1 + 2] += f1z;
        f[3 * i2] -= f1x + f3x; f[3 * i2 + 1] -= f1y + f3y; f[3 * i2 + 2] -= f1z + f3z;
        f[3 * i3] += f3x; f[3 * i3 + 1] += f3y; f[3 * i3 + 2] += f3z;
    }

    if (with_post) {
        /* [LAMMPS-ext] FixLangevin::post_force, default (uniform) noise:
         * gamma1 = -m/damp, gamma2 = sqrt(24 kB T m / (damp dt)); fix 2 all langevin 1 1 0.5 (:644) */
        if (s->langevin_on) {
            const double gamma1 = -1.0 / s->t_period;
            const double gamma2 = sqrt(24.0 / s->t_period / s->dt) * sqrt(s->t_target);
            const int noisy = (s->langevin_on == 1) && with_noise;
            for (int i = 0; i < n; i++) {
                double r3[3] = {0.0, 0.0, 0.0};
                if (noisy) {
                    const uint32_t ctr[4] = {(uint32_t)i, (uint32_t)s->eval_counter, (uint32_t)(s->eval_counter >> 32), 0u};
                    uint32_t o[4];
                    esbo_philox4x32_10(ctr, s->key, o);
                    for (int d = 0; d < 3; d++) r3[d] = u32_to_centered(o[d]);
                }
                for (int d = 0; d < 3; d++) f[3 * i + d] += gamma1 * s->v[3 * i + d] + gamma2 * r3[d];
            }
        }
        /* [LAMMPS-ext] FixWallHarmonic::wall_particle, 6 faces, EDGE (:647) */
        for (int i = 0; i < n; i++)
            for (int d = 0; d < 3; d++) {
                const double dlo = x[3 * i + d] - s->lo[d];
                const double dhi = s->hi[d] - x[3 * i + d];
                if (dlo < s->wall_cut) {
                    if (dlo <= 0.0) { if (!s->status) s->status = 1; }
                    else { const double dr = s->wall_cut - dlo; f[3 * i + d] += 2.0 * s->wall_eps * dr; e_wall += s->wall_eps * dr * dr; }
                }
                if (dhi < s->wall_cut) {
                    if (dhi <= 0.0) { if (!s->status) s->status = 1; }
                    else { const double dr = s->wall_cut - dhi; f[3 * i + d] -= 2.0 * s->wall_eps * dr; e_wall += s->wall_eps * dr * dr; }
                }
            }
        /* [LAMMPS-ext] FixSetForce: defined last, so it overwrites everything above (:671) */
        for (int i = 0; i < n; i++)
            if (s->frozen[i]) { f[3 * i] = 0.0; f[3 * i + 1] = 0.0; f[3 * i + 2] = 0.0; }
        s->eval_counter++;
    }
    s->energy[0] = e_pair; s->energy[1] = e_bond; s->energy[2] = e_angle; s->energy[3] = e_wall;
    s->n_evals++;
    for (int i = 0; i < 3 * n; i++)
        if (!isfinite(f[i])) { s->status = 4; break; }
}

/* Parity hook: forces at the current positions.
 * with_post = 0: pair+bond+angle only.  with_post = 1: also langevin/wall/setforce
 * (consumes one noise evaluation index when noise is on). */
int esbo_compute_forces(esbo_system *s, int with_post, int with_noise, double *f_out, double *e_out)
{
    s->nl_valid = 0;   /* the hook is exact at the current positions: complete list */
    compute_forces(s, with_post, with_noise);
    if (f_out) memcpy(f_out, s->f, sizeof(double) * (size_t)s->n * 3);
    if (e_out) memcpy(e_out, s->energy, sizeof(s->energy));
    return s->status;
}

static void apply_adapt(esbo_system *s, long long beginstep, long long endstep, int at_setup)
{
    for (int k = 0; k < s->nadapt; k++) {
        const esbo_adapt *a = &s->adapt[k];
        if (!at_setup) {
            if (a->nevery == 0) continue;
            if (s->step % a->nevery) continue;
        }
        /* [LAMMPS-ext] Variable ramp(x,y): delta = step - begin; if (delta != 0) delta /= end - begin;
         * value = x + delta*(y-x)  (so `run 0` evaluates to x instead of 0/0) */
        double delta = (double)(s->step - beginstep);
        if (delta != 0.0) delta /= (double)(endstep - beginstep);
        const double val = a->v0 + delta * (a->v1 - a->v0);
        if (a->kind == 0) s->soft_a = val;
        else s->bond_r0[a->index] = val;
    }
}

/* `run nsteps` or `run nsteps start S stop E` (start_step < 0: plain run).
 * [LAMMPS-ext] Run::command -> Verlet::setup (neighbour build + one force evaluation with a
 * fresh Langevin draw) -> nsteps x { nve initial; adapt pre_force; forces; post_force; nve final }.
 * Returns the status word (0 = ok). */
int esbo_run(esbo_system *s, int nsteps, long long start_step, long long stop_step)
{
    const long long first = s->step, last = s->step + nsteps;
    const long long rb = (start_step >= 0) ? start_step : first;
    const long long re = (start_step >= 0) ? stop_step : last;
    const int n = s->n;
    const double dtv = s->dt, dtf = 0.5 * s->dt;
    s->nl_valid = 0; /* setup always rebuilds */
    apply_adapt(s, rb, re, 1);
    compute_forces(s, 1, 1);
    for (int it = 0; it < nsteps && s->status == 0; it++) {
        s->step++;
        for (int i = 0; i < 3 * n; i++) {
            s->v[i] += dtf * s->f[i];
            s->x[i] += dtv * s->v[i];
        }
        /* [LAMMPS-ext] boundary f f f: an atom outside the box at re-neighbouring is lost */
        apply_adapt(s, rb, re, 0);
        compute_forces(s, 1, 1);
        for (int i = 0; i < 3 * n; i++) s->v[i] += dtf * s->f[i];
    }
    for (int i = 0; i < n && s->status == 0; i++)
        for (int d = 0; d < 3; d++)
            if (!(s->x[3 * i + d] > s->lo[d] && s->x[3 * i + d] < s->hi[d])) s->status = 2;
    return s->status;
}

long long esbo_get_step(const esbo_system *s) { return s->step; }
unsigned long long esbo_get_eval_counter(const esbo_system *s) { return s->eval_counter; }
int esbo_get_status(const esbo_system *s) { return s->status; }
void esbo_get_counters(const esbo_system *s, long long *out)
{
    out[0] = s->n_evals; out[1] = s->n_builds; out[2] = s->n_pairs_in_cut;
    out[3] = s->n_pairs_listed; out[4] = s->n_dangerous;
}
void esbo_reset_counters(esbo_system *s)
{
    s->n_evals = s->n_builds = s->n_pairs_in_cut = s->n_pairs_listed = s->n_dangerous = 0;
}
void esbo_get_energy(const esbo_system *s, double *e) { memcpy(e, s->energy, sizeof(s->energy)); }
