// esb_device.cuh -- device-side data layout and helpers shared by the kernels of libesb.so.
//
// One CTA owns one replica: positions (+type), velocities, forces and the positions of the last
// neighbour build live in shared memory for a whole chunk; the neighbour list (half list of 16-bit entries = partner
// index | table id << 11 with at most 32 tables, u32 entries otherwise; full list in the cluster variant) lives in
// per-bead arenas of a per-replica global pool.  See DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/esb.h"

#ifndef ESB_THREADS
#define ESB_THREADS 512
#endif
#ifndef ESB_PREFETCH
#define ESB_PREFETCH 2                  // neighbour-list words each lane keeps in flight (even; 4 entries with 16-bit lists); 2 is best with 16-warp CTAs
#endif
#ifndef ESB_MIN_CTAS
#define ESB_MIN_CTAS 2                  // resident CTAs per SM the step kernel is compiled for (2 x 16 warps: 64 registers)
#endif
#ifndef ESB_G
#if defined(ESB_CLUSTER) && ESB_CLUSTER
#define ESB_G 8                      // lanes cooperating on one bead in the pair loop: 8 with the full list of the cluster variant,
#else
#define ESB_G 4                      // 4 with the half list (half as many entries per bead: 8 beads per warp)
#endif
#endif
#define ESB_GROUPS (ESB_THREADS / ESB_G)
#define ESB_MAX_TERMS 8              // bonded terms per atom (linear chain: 2 bonds + 3 angles)
#define ESB_MAX_SPEC 8               // excluded partners per atom (linear chain: 6)
#define ESB_NL_MAXNB 2047            // neighbours per atom in the full list (16-bit count in the descriptor); every
#ifndef ESB_NL_STRIDE
#define ESB_NL_STRIDE 2048           // bead owns an arena of this many entries: a system of up to 2048 beads cannot
#endif
                                     // overflow it (LAMMPS: neigh_modify one 10000 page 100000, run_single_shoebox.py:533)
#ifndef ESB_BUILD_REVERSE
#define ESB_BUILD_REVERSE 0              // 1: hand out the build's batches last class first (paid with the full list of round 1; -1 % with the half list)
#endif
#define ESB_NCLASS 6                 // species classes for the row sort: chromatin-like, actin, RBP, RNP, Ser5P, gene body (induced/active)
#define ESB_N_COUNTERS 12             // per replica: evals, builds, pairs, listed, steps, total_active, cycles in atom/build/force/observe phases, dangerous builds, max |pair force| * 256
#define ESB_ACT_WORDS(count) (2 + 3 * (count))
#define ESB_NOLINK 0xffffu
#define ESB_N_SETUPS 9               // 0 = pair soft, 1..4 = table maps 0..3 (closed form + patch), 5..8 = same maps, array lookups

// One (KEY,cut) table, 48 bytes, read in the force loop as two float4 (+ two doubles on the rare
// fp64 path).  Bins [p0hi, plo) are evaluated with the LJsoft closed form at the bin midpoint; bins
// [0, p0hi) (the file's r grid is too coarse there for the spline to follow the closed form) and
// [plo, tlm1) (the spline rings across the force jump at rc, then the table is ~0) are read from
// the LOOKUP array.  Array-only tables have p0hi = plo = 0 ... i.e. an empty closed-form range.
struct __align__(16) DevTab {
    float cutsq;                // effective cutoff^2: beyond it every bin value is 0
    float invdelta_f;           // fp32 bin-index estimate (verified against the bin edges, see force_phase)
    float delta;
    float xoff_f;               // innersq * invdelta: the bin estimate is r2 * invdelta_f - xoff_f
    float K;                    // 24 eps_eff / s^6
    float is6;                  // 1 / s^6
    float c;                    // alpha (1-lam)^2
    unsigned int patch;         // p0hi | (plo - p0hi) << 16
    double innersq, invdelta;   // exact bin index: int((rsq - innersq) * invdelta)
};
static_assert(sizeof(DevTab) == 48, "DevTab layout is read with fixed offsets");

// Per pair-style set-up: everything indexed by ti*ESB_MAXT+tj.
struct PairSetup {
    float cutsq[ESB_MAXT * ESB_MAXT];       // accept test (soft: rc^2; table: effective cut^2)
    float cutsq_skin[ESB_MAXT * ESB_MAXT];  // list test (cut+skin)^2
    unsigned char tabid[ESB_MAXT * ESB_MAXT];
    float reach[ESB_MAXT * 8];              // [type][class]: largest cut+skin towards any type of that class (0: none)
};

struct Layout {
    int gene_start, promoter_length, n_enh, n_s5p;
};

struct ReplicaScalars {       // persistent per-replica state outside the bead arrays
    unsigned long long seed;
    unsigned long long eval;  // force-evaluation counter = Langevin noise index
    long long step;           // MD step
    double r0[3];             // current bond r0 (fix adapt leaves its last value behind)
    double soft_a;
    int threshold;            // ser5p_to_activate
    float p_activation;       // stored as the fp32 the device compares with (see observe)
    double p_activation_d;
    int gene_state, marked, delay, duration, total_active;
    int status;
    int ago;                  // steps since the last list build  } carried between the segments of a chunk
    int rebuild;              // a build is due at the next step  } (a chunk can be run as several work items)
    int curlist;              // entries of the current lists     }
};

struct KArgs {
    int n_atoms, n_pad, n_topo, n_replicas;
    float lo[3], hi[3];
    float dt, gamma1, gamma2, wall_eps, wall_cut, skin_half_sq;
    int langevin_mode;
    int neigh_delay;          // neigh_modify delay N: no list build sooner than N steps after the last one
    float bond_k[3], angle_k[3];
    double ramp_bond[3][2];   // ramp(v0,v1) of fix s2/s3 per bond type
    double ramp_soft[2];
    // cell grid of the list build: rows of edge row_edge = 1/row_inv over (y, z), every row cut into ncx cells of edge
    // 1/xc_inv along x; slots of the candidate sort = (class, z row, y row, x cell), x-minor
    int nry, nrz, nrows;
    float row_inv, row_edge;
    int ncx;
    float xc_inv;
    // device arrays
    float4 *pos4;             // [R][n_pad] x,y,z,type bits
    float4 *vel4;             // [R][n_pad]
    float4 *bref4;            // [R][n_pad] positions at the last list build, carried between the segments of a chunk
    ReplicaScalars *rs;       // [R]
    const unsigned char *frozen;     // [n_pad]
    uint2 *bterm;             // [ESB_MAX_TERMS][n_topo_pad]  (+ r * bterm_stride: per replica with actin dynamics)
    unsigned short *spec;     // [n_topo_pad][ESB_MAX_SPEC]   (+ r * spec_stride)
    int n_topo_pad;
    size_t bterm_stride, spec_stride;
    // synthetic microscopy (mic_out == nullptr: off): bin edges of np.histogramdd per axis, frames kept per replica
    double mic_edges[3][32];
    int mic_n[3], mic_frames;
    unsigned short *mic_out;  // [R][mic_frames][5][nx][ny][nz]
    // actin polymerisation dynamics (act.count == 0: off)
    esb_actin_params act;
    unsigned int *act_link;   // [R][act.count]  neighbour towards the plus end | neighbour towards the minus end << 16 (atom ids, 0xffff none)
    unsigned short *act_lists;       // [R][ESB_ACT_WORDS(count)]: n_free, n_fil, free[count], fil_plus[count], fil_minus[count]
    int *act_stats;           // [R][max_chunks][6]
    unsigned short *act_hist; // [R][max_chunks][49]
    const PairSetup *setups;  // [ESB_N_SETUPS]
    const DevTab *tabs;       // [n_tables] (the launch picks the closed-form or the array flavour)
    int n_tables;
    const float *tabval;      // patch / array values (F/r), fp32
    const float *tabe;        // energy arrays (hook)
    int tlm1;
    unsigned int *nl_pool;    // [R][n_pad][ESB_NL_STRIDE]  neighbour index | table id << 16; or two 16-bit entries per word:
    int *nl_levels;           // [R][64] half list: work items of the force loop per chunk level (running sums), kept for the segments of a chunk
    int pack16;               // 1: 16-bit list entries (neighbour index | table id << 11), possible with at most 32 tables
    uint2 *nl_qinfo;          // [R][2][n_pad] {list start, atom | count << 16}: [0] sorted by list length (force order), [1] slot order (build)
    unsigned char class_of[ESB_MAXT + 4];
    const esb_chunk_desc *descs;
    // observation
    Layout lay;
    const int *enh_idx, *s5p_idx;
    const double *cluster_r;  // [50]
    double *frames;           // [R][max_chunks][9]
    double *rawd;             // [R][max_chunks][2]  d_rp, d_rg in bead units (gene_stats.txt needs d_rg < 3 exactly)
    unsigned short *hist;     // [R][max_chunks][49]
    int max_chunks;
    unsigned long long *counters;  // [R][ESB_N_COUNTERS]
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// u32 -> uniform in [-0.5, 0.5]; identical to the oracle's u32_to_centered.
__device__ __forceinline__ float u32_centered(uint32_t u) { return (float)(int32_t)u * 0x1p-32f; }

// esb_kernels.cu is compiled three times into libesb.so: ESB_TAG 512 = 16-warp CTAs, two per SM (full batches);
// ESB_TAG 1024 = 32-warp CTAs, one per SM, which steps a replica twice as fast and is chosen when the batch has
// no more replicas than the device has SMs; ESB_TAG 2048 = the same CTA in thread-block clusters of 2/4/8 that step
// ONE replica together (distributed shared memory), for batches of at most half / a quarter / an eighth of the SMs
// (e.g. configs[1]: 64 actin seeds on one GPU).
#ifndef ESB_TAG
#define ESB_TAG 512
#endif
#define ESB_CAT2_(a, b) a##b
#define ESB_CAT2(a, b) ESB_CAT2_(a, b)
#define ESB_VARIANT(name) ESB_CAT2(name##_v, ESB_TAG)
size_t esb_smem_bytes_v512(int n_pad, int n_tables, int n_topo_pad);
size_t esb_smem_bytes_v1024(int n_pad, int n_tables, int n_topo_pad);
size_t esb_smem_bytes_v2048(int n_pad, int n_tables, int n_topo_pad);
cudaError_t esb_launch_run_v2048(const KArgs &k, int chunk_begin, int n_chunks, double delt, double sig_nm, size_t smem, cudaStream_t st,
                                 int *sched, int n_sm, int cluster_size);
cudaError_t esb_launch_run_v512(const KArgs &k, int chunk_begin, int n_chunks, double delt, double sig_nm, size_t smem, cudaStream_t st,
                                int *sched, int n_sm, int n_seg);
cudaError_t esb_launch_run_v1024(const KArgs &k, int chunk_begin, int n_chunks, double delt, double sig_nm, size_t smem, cudaStream_t st,
                                 int *sched, int n_sm, int n_seg);
cudaError_t esb_launch_forces(const KArgs &k, int pair_mode, int setup_id, float soft_a, int with_post, unsigned long long eval,
                              double *f_out, double *e_out, size_t smem, cudaStream_t st);
