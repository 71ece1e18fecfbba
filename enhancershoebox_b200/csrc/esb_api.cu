// esb_api.cu -- host side of libesb.so: the C ABI declared in include/esb.h.
//
// Owns all device memory of a replica batch, turns the reference-shaped inputs (data-file
// topology, pair_coeff tables and maps, per-chunk programme) into the device layout of
// esb_device.cuh and launches the kernels of esb_kernels.cu.  There is no CPU fallback: without a
// CUDA device every entry point that needs one fails with ESB_ENODEV.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "esb_device.cuh"

extern "C" __global__ void esb_pack_kernel(const double *x, const double *v, const int *types, const unsigned char *frozen,
                                           float4 *pos4, float4 *vel4, int n_atoms, int n_pad, int n_replicas, int replicate);
extern "C" __global__ void esb_unpack_kernel(const float4 *pos4, const float4 *vel4, double *x, double *v, int *types,
                                             int n_atoms, int n_pad, int n_replicas);
__global__ void esb_mark_frozen_kernel(float4 *pos4, const unsigned char *frozen, int n_pad, size_t total)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        float4 p = pos4[q];
        const int t = (__float_as_int(p.w) & 0xff) | (frozen[q % n_pad] ? 0x100 : 0);
        p.w = __int_as_float(t);
        pos4[q] = p;
    }
}

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ESB_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// temporary device buffer of a host-facing call: freed on every return path
template <class T>
struct Scratch {
    T *p = nullptr;
    ~Scratch() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) { return cudaMalloc((void **)&p, (count ? count : 1) * sizeof(T)); }
};

template <class T>
int dev_alloc(T **p, size_t count)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
    if (e != cudaSuccess) return fail(ESB_ENOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
    return ESB_OK;
}

template <class T>
int dev_upload(T **p, const std::vector<T> &v)
{
    int rc = dev_alloc(p, v.size());
    if (rc) return rc;
    if (!v.empty()) CK(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ESB_OK;
}

struct HostTable {
    std::vector<double> f, e;
    double innersq, delta, cut;
    esb_closed_form cf;
};

}  // namespace

struct esb_handle {
    int device = 0, R = 0, N = 0, n_pad = 0, n_types = 0;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    bool have_box = false, have_topo = false, have_soft = false, have_tables = false, have_layout = false, have_state = false;
    bool setups_dirty = true;
    double dt = 0.005, t_target = 1.0, t_damp = 0.5, wall_eps = 100.0, wall_cut = 3.0, skin = 0.3, cell_edge = 2.8, xcell_edge = 1.0;
    int neigh_delay = 10;
    int langevin_mode = 1;
    double bond_k[3] = {0, 90.0, 90.0}, bond_r0[3] = {0, 1.0, 0.3}, angle_k[3] = {0, 100.0 / 60.0, 33.0};
    double soft_cut[ESB_MAXT * ESB_MAXT];
    std::vector<HostTable> tables;
    int tlm1 = 0;
    int maps[4][ESB_MAXT * ESB_MAXT];
    bool have_map[4] = {false, false, false, false};
    std::vector<unsigned char> frozen;
    int n_topo = 0, n_topo_pad = 0;
    int n_chunks = 0;
    // device
    float4 *d_pos4 = nullptr, *d_vel4 = nullptr;
    ReplicaScalars *d_rs = nullptr;
    unsigned char *d_frozen = nullptr;
    uint2 *d_bterm = nullptr;
    unsigned short *d_spec = nullptr;
    PairSetup *d_setups = nullptr;
    DevTab *d_tabs_closed = nullptr, *d_tabs_array = nullptr;
    float *d_tabval = nullptr, *d_tabe = nullptr;
    float4 *d_bref4 = nullptr;
    unsigned int *d_nl_pool = nullptr;
    uint2 *d_nl_qinfo = nullptr;
    int *d_nl_levels = nullptr;
    esb_chunk_desc *d_descs = nullptr;
    int *d_enh = nullptr, *d_s5p = nullptr;
    double *d_cluster_r = nullptr;
    double *d_frames = nullptr, *d_rawd = nullptr;
    unsigned short *d_hist = nullptr;
    unsigned long long *d_counters = nullptr;
    int *d_sched = nullptr;
    // topology as given (kept for esb_set_actin) and the dynamic actin state
    std::vector<int> bonds, angles;
    esb_actin_params act = {};
    bool have_actin = false;
    unsigned int *d_act_link = nullptr;
    unsigned short *d_act_lists = nullptr;
    int *d_act_stats = nullptr;
    unsigned short *d_act_hist = nullptr;
    size_t bterm_stride = 0, spec_stride = 0;
    double mic_edges[3][32];
    int mic_n[3] = {0, 0, 0}, mic_frames = 0;
    unsigned short *d_mic = nullptr;
    int n_sm = 148;
    int n_seg = 0;                 // esb_set_segments
    std::vector<esb_chunk_desc> descs_host;      // the schedule, for validation at run time
    cudaStream_t last_stream = nullptr;          // stream of the last esb_run: the blocking getters wait for it
    bool run_pending = false;
    int force_wide = -1;           // esb_set_segments: CTA width of the step kernel (-1: by batch size)
    size_t smem_per_sm = 233472;
    Layout lay = {0, 0, 0, 0};
    std::vector<ReplicaScalars> rs_host;
    int nry = 1, nrz = 1, ncx = 1;
};

namespace {

// esb_run is asynchronous on the caller's stream; the getters below copy on the legacy default stream, which does not
// wait for a cudaStreamNonBlocking stream: wait for the last run explicitly.
int wait_for_run(esb_handle *h)
{
    if (h->run_pending) {
        CK(cudaStreamSynchronize(h->last_stream));
        h->run_pending = false;
    }
    return ESB_OK;
}

int check_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(ESB_ENODEV, "no CUDA device visible (%s); libesb has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(ESB_EINVAL, "device %d out of range (count %d)", device, count);
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return fail(ESB_ENODEV, "device %d is sm_%d%d; libesb is built for sm_100a only", device, p.major, p.minor);
    return ESB_OK;
}

// F(r)/r of the LJsoft closed form at r (ljsoft_potential.py:10-16), fp64.
double closed_f_over_r(const esb_closed_form &cf, double r)
{
    if (!(r < cf.rc)) return 0.0;
    const double rs = r / cf.s;
    const double x = cf.c + rs * rs * rs * rs * rs * rs;
    const double s6 = pow(cf.s, 6);
    return 24.0 * cf.eps_eff * r * r * r * r / s6 * (2.0 / (x * x * x) - 1.0 / (x * x));
}

// species classes of the row sort (run_single_shoebox.py:23-33): chromatin-like {1,2,9,10,11} 0, actin 3 -> 1,
// RBP 4 -> 2, RNP 5 -> 3, Ser5P 8 -> 4, gene body while induced/active {6,7} -> 5 (the only chromatin types with
// the long RNP cutoffs; kept apart so that they do not widen every chromatin search)
const unsigned char k_class_of[ESB_MAXT] = {0, 0, 0, 1, 2, 3, 5, 5, 4, 0, 0, 0};

int rebuild_setups(esb_handle *h)
{
    if (!h->have_box) return fail(ESB_ESTATE, "esb_set_box must be called first");
    // ---- cell grid of the list build: rows over (y, z) of edge cell_edge, every row cut into cells of edge xcell_edge
    //      along x.  The build scratch (the 3 force arrays, 12 n_pad bytes) holds the (class, row, x cell) table (2 B per
    //      slot), one u16 index array and the list lengths (4 B per bead): slots + 2 <= 3 n_pad - 16.  The build packs
    //      a bead's first row and row count in y and z into 8 bits each.
    const int cap = 3 * h->n_pad - 32;
    double c = h->cell_edge;
    for (;;) {
        h->nry = std::max(1, (int)ceil((h->hi[1] - h->lo[1]) / c));
        h->nrz = std::max(1, (int)ceil((h->hi[2] - h->lo[2]) / c));
        if ((long long)h->nry * h->nrz * ESB_NCLASS <= cap && h->nry <= 255 && h->nrz <= 255) break;
        c *= 1.05;
    }
    h->cell_edge = c;
    h->ncx = std::max(1, (int)ceil((h->hi[0] - h->lo[0]) / h->xcell_edge));
    h->ncx = std::max(1, std::min(h->ncx, cap / (h->nry * h->nrz * ESB_NCLASS)));

    const int nt = (int)h->tables.size();
    std::vector<DevTab> closed(nt), arr(nt);
    std::vector<double> eff_cut_closed(nt), eff_cut_arr(nt);
    for (int t = 0; t < nt; t++) {
        const HostTable &T = h->tables[t];
        DevTab d;
        memset(&d, 0, sizeof(d));
        d.innersq = T.innersq; d.invdelta = 1.0 / T.delta; d.delta = (float)T.delta; d.invdelta_f = (float)(1.0 / T.delta);
        d.xoff_f = (float)(T.innersq / T.delta);
        d.patch = 0u;                                  // empty closed-form range: every bin from the array
        d.cutsq = nextafterf((float)(T.cut * T.cut * (1.0 + 2e-7)), INFINITY);
        d.K = 0.0f; d.is6 = 0.0f; d.c = 1.0f;
        arr[t] = d;
        eff_cut_arr[t] = T.cut;
        eff_cut_closed[t] = T.cut;
        if (T.cf.valid && h->tlm1 < 65536) {
            // bins where the spline-built LOOKUP value departs from the closed form at the midpoint:
            // a run at the start of the table and a run around rc
            std::vector<char> bad(h->tlm1);
            for (int i = 0; i < h->tlm1; i++) {
                const double mid = T.innersq + (i + 0.5) * T.delta;
                const double fc = closed_f_over_r(T.cf, sqrt(mid));
                bad[i] = fabs(T.f[i] - fc) > 1e-7 * (1.0 + fabs(T.f[i]));
            }
            const int irc = (int)std::min((double)h->tlm1 - 1, std::max(0.0, (T.cf.rc * T.cf.rc - T.innersq) / T.delta));
            int last_bad = -1, plo = h->tlm1, phi = 0;
            const int stop = std::max(0, irc - 2000);
            for (int i = 0; i < stop; i++) if (bad[i]) last_bad = i;
            const int p0hi = last_bad + 1;
            for (int i = stop; i < h->tlm1; i++) if (bad[i]) { plo = std::min(plo, i); phi = std::max(phi, i + 1); }
            if (p0hi <= h->tlm1 / 2 && plo > p0hi) {
                // closed form on [p0hi, plo); everything from plo on comes from the array
                d.patch = (unsigned int)p0hi | ((unsigned int)(plo - p0hi) << 16);
                const double s6 = pow(T.cf.s, 6);
                d.K = (float)(24.0 * T.cf.eps_eff / s6);
                d.is6 = (float)(1.0 / s6);
                d.c = (float)T.cf.c;
                // beyond the last departing bin the table is the closed form's 0: nothing to evaluate there
                double eff = T.cut * T.cut;
                if (phi > 0 && T.cf.rc < T.cut) eff = std::min(eff, std::max(T.cf.rc * T.cf.rc, T.innersq + phi * T.delta));
                d.cutsq = nextafterf((float)(eff * (1.0 + 2e-7)), INFINITY);
                eff_cut_closed[t] = sqrt(eff);
            }
        }
        closed[t] = d;
    }
    int rc;
    if ((rc = dev_upload(&h->d_tabs_closed, closed))) return rc;
    if ((rc = dev_upload(&h->d_tabs_array, arr))) return rc;

    std::vector<PairSetup> setups(ESB_N_SETUPS);
    memset(setups.data(), 0, sizeof(PairSetup) * ESB_N_SETUPS);
    auto fill = [&](PairSetup &ps, auto cut_of, auto tab_of) {
        for (int a = 1; a <= h->n_types; a++) {
            for (int b = 1; b <= h->n_types; b++) {
                const double cu = cut_of(a, b);
                const int q = a * ESB_MAXT + b;
                ps.tabid[q] = (unsigned char)tab_of(a, b);
                ps.cutsq_skin[q] = (float)((cu + h->skin) * (cu + h->skin));
                float &reach = ps.reach[a * 8 + k_class_of[b]];
                if (cu > 0.0) reach = std::max(reach, (float)((cu + h->skin) * (1.0 + 1e-6) + 1e-6));
            }
        }
    };
    if (h->have_soft) {
        PairSetup &ps = setups[0];
        fill(ps, [&](int a, int b) { return h->soft_cut[a * ESB_MAXT + b]; }, [&](int, int) { return 0; });
        for (int q = 0; q < ESB_MAXT * ESB_MAXT; q++) ps.cutsq[q] = (float)(h->soft_cut[q] * h->soft_cut[q]);
    }
    for (int m = 0; m < 4; m++) {
        if (!h->have_map[m] || !h->have_tables) continue;
        for (int flavour = 0; flavour < 2; flavour++) {
            PairSetup &ps = setups[1 + m + 4 * flavour];
            const std::vector<double> &eff = flavour ? eff_cut_arr : eff_cut_closed;
            const std::vector<DevTab> &tb = flavour ? arr : closed;
            fill(ps, [&](int a, int b) { return eff[h->maps[m][a * ESB_MAXT + b]]; },
                 [&](int a, int b) { return h->maps[m][a * ESB_MAXT + b]; });
            for (int a = 1; a <= h->n_types; a++)
                for (int b = 1; b <= h->n_types; b++)
                    ps.cutsq[a * ESB_MAXT + b] = tb[h->maps[m][a * ESB_MAXT + b]].cutsq;
        }
    }
    if ((rc = dev_upload(&h->d_setups, setups))) return rc;
    h->setups_dirty = false;
    return ESB_OK;
}

int fill_kargs(esb_handle *h, KArgs &k, bool array_flavour)
{
    memset(&k, 0, sizeof(k));
    k.n_atoms = h->N; k.n_pad = h->n_pad; k.n_topo = h->n_topo; k.n_replicas = h->R;
    for (int d = 0; d < 3; d++) { k.lo[d] = (float)h->lo[d]; k.hi[d] = (float)h->hi[d]; }
    k.dt = (float)h->dt;
    k.gamma1 = (float)(-1.0 / h->t_damp);
    k.gamma2 = (float)(sqrt(24.0 / h->t_damp / h->dt) * sqrt(h->t_target));
    k.wall_eps = (float)h->wall_eps; k.wall_cut = (float)h->wall_cut;
    k.skin_half_sq = (float)(0.25 * h->skin * h->skin);
    k.langevin_mode = h->langevin_mode;
    k.neigh_delay = h->neigh_delay;
    for (int t = 0; t < 3; t++) { k.bond_k[t] = (float)h->bond_k[t]; k.angle_k[t] = (float)h->angle_k[t]; }
    k.ramp_bond[1][0] = 0.033; k.ramp_bond[1][1] = 1.0;   // variable chromatinr0 equal ramp(0.033,1.0)  (:635)
    k.ramp_bond[2][0] = 0.033; k.ramp_bond[2][1] = 0.3;   // variable actinr0 equal ramp(0.033,0.3)      (:638)
    k.ramp_soft[0] = 0.0; k.ramp_soft[1] = 60.0;          // variable prefactor equal ramp(0,60)         (:627)
    k.nry = h->nry; k.nrz = h->nrz; k.nrows = h->nry * h->nrz;
    k.row_inv = (float)(1.0 / h->cell_edge); k.row_edge = (float)h->cell_edge;
    k.ncx = h->ncx; k.xc_inv = (float)(h->ncx / (h->hi[0] - h->lo[0]));
    k.pos4 = h->d_pos4; k.vel4 = h->d_vel4; k.bref4 = h->d_bref4; k.rs = h->d_rs; k.frozen = h->d_frozen;
    k.bterm = h->d_bterm; k.spec = h->d_spec; k.n_topo_pad = h->n_topo_pad;
    k.bterm_stride = h->bterm_stride; k.spec_stride = h->spec_stride;
    if (h->d_mic) {
        memcpy(k.mic_edges, h->mic_edges, sizeof(k.mic_edges));
        for (int d = 0; d < 3; d++) k.mic_n[d] = h->mic_n[d];
        k.mic_frames = h->mic_frames; k.mic_out = h->d_mic;
    }
    if (h->have_actin) { k.act = h->act; k.act_link = h->d_act_link; k.act_lists = h->d_act_lists; k.act_stats = h->d_act_stats; k.act_hist = h->d_act_hist; }
    k.setups = h->d_setups;
    k.tabs = array_flavour ? h->d_tabs_array : h->d_tabs_closed;
    k.n_tables = (int)h->tables.size();
    k.tabval = h->d_tabval; k.tabe = h->d_tabe; k.tlm1 = h->tlm1;
    k.nl_pool = h->d_nl_pool; k.nl_qinfo = h->d_nl_qinfo; k.nl_levels = h->d_nl_levels;
    // 16-bit list entries need the table id in 5 bits and the bead index in 11 (always true: n_pad <= ESB_NL_STRIDE)
    k.pack16 = getenv("ESB_PACK16") ? atoi(getenv("ESB_PACK16")) != 0 && h->tables.size() <= 32 : h->tables.size() <= 32;
    memcpy(k.class_of, k_class_of, ESB_MAXT);
    k.descs = h->d_descs;
    k.lay = h->lay; k.enh_idx = h->d_enh; k.s5p_idx = h->d_s5p; k.cluster_r = h->d_cluster_r;
    k.frames = h->d_frames; k.rawd = h->d_rawd; k.hist = h->d_hist; k.max_chunks = h->n_chunks;
    k.counters = h->d_counters;
    return ESB_OK;
}

int push_rs(esb_handle *h)
{
    CK(cudaMemcpy(h->d_rs, h->rs_host.data(), sizeof(ReplicaScalars) * h->R, cudaMemcpyHostToDevice));
    return ESB_OK;
}
int pull_rs(esb_handle *h)
{
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    CK(cudaMemcpy(h->rs_host.data(), h->d_rs, sizeof(ReplicaScalars) * h->R, cudaMemcpyDeviceToHost));
    return ESB_OK;
}

}  // namespace

extern "C" {

const char *esb_last_error(void) { return g_err.c_str(); }
int esb_version(void) { return 100; }

int esb_create(int device, int n_replicas, int n_atoms, int n_types, esb_handle **out)
{
    if (!out) return fail(ESB_EINVAL, "out is NULL");
    *out = nullptr;
    if (n_replicas < 1 || n_atoms < 1 || n_atoms > ESB_NL_STRIDE || n_types < 1 || n_types >= ESB_MAXT)
        return fail(ESB_EINVAL, "bad sizes: replicas=%d atoms=%d (max %d) types=%d (max %d)", n_replicas, n_atoms, ESB_NL_STRIDE, n_types, ESB_MAXT - 1);
    int rc = check_device(device);
    if (rc) return rc;
    CK(cudaSetDevice(device));
    esb_handle *h = new esb_handle();
    { cudaDeviceProp p; if (cudaGetDeviceProperties(&p, device) == cudaSuccess) { h->n_sm = p.multiProcessorCount; h->smem_per_sm = p.sharedMemPerMultiprocessor; } }
    h->device = device; h->R = n_replicas; h->N = n_atoms; h->n_types = n_types;
    h->n_pad = (n_atoms + 31) & ~31;
    if (const char *s = getenv("ESB_SKIN")) h->skin = atof(s);
    if (const char *s = getenv("ESB_DELAY")) h->neigh_delay = atoi(s);
    if (const char *s = getenv("ESB_CELL")) h->cell_edge = atof(s);
    if (const char *s = getenv("ESB_XCELL")) h->xcell_edge = atof(s);
    memset(h->soft_cut, 0, sizeof(h->soft_cut));
    memset(h->maps, 0, sizeof(h->maps));
    h->frozen.assign(h->n_pad, 0);
    h->rs_host.assign(n_replicas, ReplicaScalars());
    for (auto &rs : h->rs_host) {
        memset(&rs, 0, sizeof(rs));
        rs.seed = 1; rs.r0[1] = 1.0; rs.r0[2] = 0.3; rs.threshold = 80; rs.p_activation_d = 0.03; rs.p_activation = 0.03f;
    }
    const size_t tot = (size_t)n_replicas * h->n_pad;
    if ((rc = dev_alloc(&h->d_pos4, tot)) || (rc = dev_alloc(&h->d_vel4, tot)) || (rc = dev_alloc(&h->d_bref4, tot)) || (rc = dev_alloc(&h->d_rs, (size_t)n_replicas)) ||
        (rc = dev_alloc(&h->d_frozen, (size_t)h->n_pad)) || (rc = dev_alloc(&h->d_nl_pool, (size_t)n_replicas * h->n_pad * ESB_NL_STRIDE + 1024)) ||
        (rc = dev_alloc(&h->d_nl_qinfo, 2 * tot)) || (rc = dev_alloc(&h->d_nl_levels, (size_t)n_replicas * 64)) || (rc = dev_alloc(&h->d_counters, (size_t)n_replicas * ESB_N_COUNTERS)) || (rc = dev_alloc(&h->d_sched, (size_t)n_replicas + 1))) {
        esb_destroy(h);
        return rc;
    }
    cudaMemset(h->d_pos4, 0, tot * sizeof(float4));
    cudaMemset(h->d_vel4, 0, tot * sizeof(float4));
    cudaMemset(h->d_frozen, 0, h->n_pad);
    cudaMemset(h->d_counters, 0, (size_t)n_replicas * ESB_N_COUNTERS * sizeof(unsigned long long));
    cudaMemset(h->d_nl_qinfo, 0, 2 * tot * sizeof(uint2));
    if ((rc = push_rs(h))) { esb_destroy(h); return rc; }
    // empty topology by default
    std::vector<uint2> bt((size_t)ESB_MAX_TERMS * 32, make_uint2(0, 0));
    std::vector<unsigned short> sp((size_t)32 * ESB_MAX_SPEC, 0xffff);
    h->n_topo = 0; h->n_topo_pad = 32;
    if ((rc = dev_upload(&h->d_bterm, bt)) || (rc = dev_upload(&h->d_spec, sp))) { esb_destroy(h); return rc; }
    *out = h;
    return ESB_OK;
}

int esb_destroy(esb_handle *h)
{
    if (!h) return ESB_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_pos4); cudaFree(h->d_vel4); cudaFree(h->d_rs); cudaFree(h->d_frozen); cudaFree(h->d_bterm);
    cudaFree(h->d_spec); cudaFree(h->d_setups); cudaFree(h->d_tabs_closed); cudaFree(h->d_tabs_array);
    cudaFree(h->d_tabval); cudaFree(h->d_tabe); cudaFree(h->d_nl_pool); cudaFree(h->d_nl_qinfo); cudaFree(h->d_nl_levels); cudaFree(h->d_descs);
    cudaFree(h->d_enh); cudaFree(h->d_s5p); cudaFree(h->d_cluster_r); cudaFree(h->d_frames); cudaFree(h->d_rawd); cudaFree(h->d_hist);
    cudaFree(h->d_counters); cudaFree(h->d_sched); cudaFree(h->d_bref4);
    cudaFree(h->d_mic);
    cudaFree(h->d_act_link); cudaFree(h->d_act_lists); cudaFree(h->d_act_stats); cudaFree(h->d_act_hist);
    delete h;
    return ESB_OK;
}

int esb_set_box(esb_handle *h, const double lo[3], const double hi[3])
{
    if (!h || !lo || !hi) return fail(ESB_EINVAL, "NULL argument");
    for (int d = 0; d < 3; d++) {
        if (!(hi[d] > lo[d])) return fail(ESB_EINVAL, "box dimension %d is empty", d);
        h->lo[d] = lo[d]; h->hi[d] = hi[d];
    }
    h->have_box = true; h->setups_dirty = true;
    return ESB_OK;
}

int esb_set_topology(esb_handle *h, int n_bonds, const int *bonds, int n_angles, const int *angles, int n_frozen, const int *frozen)
{
    if (!h || n_bonds < 0 || n_angles < 0 || n_frozen < 0) return fail(ESB_EINVAL, "bad argument");
    CK(cudaSetDevice(h->device));
    const int N = h->N;
    if (bonds != h->bonds.data()) h->bonds.assign(bonds, bonds + (size_t)3 * n_bonds);
    if (angles != h->angles.data()) h->angles.assign(angles, angles + (size_t)4 * n_angles);
    bonds = h->bonds.data(); angles = h->angles.data();
    int n_topo = h->have_actin ? h->act.first + h->act.count : 0;    // dynamic actin: every actin bead owns table rows
    for (int b = 0; b < n_bonds; b++)
        for (int q = 1; q <= 2; q++) {
            const int a = bonds[3 * b + q];
            if (a < 0 || a >= N) return fail(ESB_EINVAL, "bond %d references atom %d", b, a);
            n_topo = std::max(n_topo, a + 1);
        }
    for (int g = 0; g < n_angles; g++)
        for (int q = 1; q <= 3; q++) {
            const int a = angles[4 * g + q];
            if (a < 0 || a >= N) return fail(ESB_EINVAL, "angle %d references atom %d", g, a);
            n_topo = std::max(n_topo, a + 1);
        }
    const int ntp = std::max(32, (n_topo + 31) & ~31);
    std::vector<uint2> bt((size_t)ESB_MAX_TERMS * ntp, make_uint2(0, 0));
    std::vector<int> nterm(ntp, 0);
    auto add = [&](int atom, unsigned a, unsigned b, int kind, int type) -> bool {
        if (nterm[atom] >= ESB_MAX_TERMS) return false;
        bt[(size_t)nterm[atom] * ntp + atom] = make_uint2(a | (b << 16), (unsigned)kind | ((unsigned)type << 8));
        nterm[atom]++;
        return true;
    };
    for (int b = 0; b < n_bonds; b++) {
        const int t = bonds[3 * b], i = bonds[3 * b + 1], j = bonds[3 * b + 2];
        if (t < 1 || t > 2) return fail(ESB_EINVAL, "bond type %d (only 1..2)", t);
        if (!add(i, j, 0, 1, t) || !add(j, i, 0, 1, t)) return fail(ESB_EINVAL, "more than %d bonded terms on one atom", ESB_MAX_TERMS);
    }
    // a bead's row: its bonds, the angles it is the vertex of (the full-batch kernel evaluates an angle there, once),
    // then the angles it is an end of
    for (int pass = 0; pass < 2; pass++)
        for (int g = 0; g < n_angles; g++) {
            const int t = angles[4 * g], i = angles[4 * g + 1], j = angles[4 * g + 2], k = angles[4 * g + 3];
            if (t < 1 || t > 2) return fail(ESB_EINVAL, "angle type %d (only 1..2)", t);
            if (pass == 0 ? !add(j, i, k, 3, t) : (!add(i, j, k, 2, t) || !add(k, j, i, 2, t)))
                return fail(ESB_EINVAL, "more than %d bonded terms on one atom", ESB_MAX_TERMS);
        }
    // 1-2 / 1-3 / 1-4 exclusions: breadth-first to depth 3 over the bond graph
    std::vector<std::vector<int>> adj(ntp);
    for (int b = 0; b < n_bonds; b++) { adj[bonds[3 * b + 1]].push_back(bonds[3 * b + 2]); adj[bonds[3 * b + 2]].push_back(bonds[3 * b + 1]); }
    std::vector<unsigned short> sp((size_t)ntp * ESB_MAX_SPEC, 0xffff);
    std::vector<int> mark(ntp, -1);
    for (int i = 0; i < n_topo; i++) {
        if (adj[i].empty()) continue;
        std::vector<int> cur{i}, nxt;
        mark[i] = i;
        int cnt = 0;
        for (int depth = 0; depth < 3; depth++) {
            nxt.clear();
            for (int a : cur)
                for (int b : adj[a]) {
                    if (mark[b] == i) continue;
                    mark[b] = i;
                    nxt.push_back(b);
                    if (cnt >= ESB_MAX_SPEC) return fail(ESB_EINVAL, "more than %d excluded partners on atom %d", ESB_MAX_SPEC, i);
                    sp[(size_t)i * ESB_MAX_SPEC + cnt++] = (unsigned short)b;
                }
            cur.swap(nxt);
        }
    }
    if (frozen) {
        h->frozen.assign(h->n_pad, 0);
        for (int q = 0; q < n_frozen; q++) {
            if (frozen[q] < 0 || frozen[q] >= N) return fail(ESB_EINVAL, "frozen atom %d out of range", frozen[q]);
            h->frozen[frozen[q]] = 1;
        }
    }
    int rc;
    if (h->have_actin) {
        // bonds, angles and exclusions are per-replica state: R copies of the tables
        std::vector<uint2> btr((size_t)h->R * bt.size());
        std::vector<unsigned short> spr((size_t)h->R * sp.size());
        for (int r = 0; r < h->R; r++) {
            std::copy(bt.begin(), bt.end(), btr.begin() + (size_t)r * bt.size());
            std::copy(sp.begin(), sp.end(), spr.begin() + (size_t)r * sp.size());
        }
        h->bterm_stride = bt.size(); h->spec_stride = sp.size();
        if ((rc = dev_upload(&h->d_bterm, btr)) || (rc = dev_upload(&h->d_spec, spr))) return rc;
    } else {
        h->bterm_stride = h->spec_stride = 0;
        if ((rc = dev_upload(&h->d_bterm, bt)) || (rc = dev_upload(&h->d_spec, sp))) return rc;
    }
    if ((rc = dev_upload(&h->d_frozen, h->frozen))) return rc;
    h->n_topo = n_topo; h->n_topo_pad = ntp; h->have_topo = true;
    if (h->have_state) {
        const size_t tot = (size_t)h->R * h->n_pad;
        esb_mark_frozen_kernel<<<256, 256>>>(h->d_pos4, h->d_frozen, h->n_pad, tot);
        CK(cudaGetLastError());
    }
    return ESB_OK;
}

int esb_set_coeffs(esb_handle *h, const double bond_k[3], const double bond_r0[3], const double angle_k[3])
{
    if (!h || !bond_k || !bond_r0 || !angle_k) return fail(ESB_EINVAL, "NULL argument");
    CK(cudaSetDevice(h->device));
    for (int t = 0; t < 3; t++) { h->bond_k[t] = bond_k[t]; h->bond_r0[t] = bond_r0[t]; h->angle_k[t] = angle_k[t]; }
    int rc = pull_rs(h);
    if (rc) return rc;
    for (auto &rs : h->rs_host) { rs.r0[1] = bond_r0[1]; rs.r0[2] = bond_r0[2]; }
    return push_rs(h);
}

int esb_set_fixes(esb_handle *h, double dt, double t_target, double t_damp, int langevin_mode, double wall_eps, double wall_cut)
{
    if (!h || !(dt > 0) || !(t_damp > 0) || t_target < 0 || langevin_mode < 0 || langevin_mode > 2)
        return fail(ESB_EINVAL, "bad fix parameters");
    h->dt = dt; h->t_target = t_target; h->t_damp = t_damp; h->langevin_mode = langevin_mode;
    h->wall_eps = wall_eps; h->wall_cut = wall_cut;
    return ESB_OK;
}

int esb_set_actin(esb_handle *h, const esb_actin_params *p)
{
    if (!h || !p) return fail(ESB_EINVAL, "NULL argument");
    if (!h->have_topo) return fail(ESB_ESTATE, "esb_set_topology must be called first");
    if (p->first < 0 || p->count < 2 || p->first + p->count > h->N || p->count > 16000 || p->max_tries < 1)
        return fail(ESB_EINVAL, "bad actin range [%d, %d) of %d atoms", p->first, p->first + p->count, h->N);
    if ((size_t)p->count * 29 > (size_t)12 * h->n_pad) return fail(ESB_EINVAL, "too many actin beads for the event scratch");
    CK(cudaSetDevice(h->device));
    h->act = *p; h->have_actin = true;
    // per-replica tables (rows for every actin bead) from the stored topology
    int rc = esb_set_topology(h, (int)(h->bonds.size() / 3), h->bonds.data(), (int)(h->angles.size() / 4), h->angles.data(), 0, nullptr);
    if (rc) return rc;
    // initial filaments: the type-2 bonds among the actin beads; the lower atom id is towards the plus end
    const int first = p->first, count = p->count;
    std::vector<unsigned int> link(count, 0xffffffffu);
    for (size_t b = 0; b < h->bonds.size() / 3; b++) {
        int i = h->bonds[3 * b + 1], j = h->bonds[3 * b + 2];
        const bool ai = i >= first && i < first + count, aj = j >= first && j < first + count;
        if (!ai && !aj) continue;
        if (!(ai && aj) || h->bonds[3 * b] != 2) return fail(ESB_EINVAL, "bond %zu ties an actin bead to something that is not an actin filament", b);
        if (i > j) std::swap(i, j);
        if ((link[i - first] >> 16) != ESB_NOLINK || (link[j - first] & 0xffffu) != ESB_NOLINK)
            return fail(ESB_EINVAL, "actin bonds do not form linear filaments ordered by atom id");
        link[i - first] = (link[i - first] & 0x0000ffffu) | ((unsigned int)j << 16);
        link[j - first] = (link[j - first] & 0xffff0000u) | (unsigned int)i;
    }
    std::vector<unsigned short> lists(ESB_ACT_WORDS(count), 0);
    unsigned short *fr = lists.data() + 2, *fplus = fr + count, *fminus = fplus + count;
    int nfree = 0, nfil = 0;
    for (int k = 0; k < count; k++) {
        if (link[k] == 0xffffffffu) { fr[nfree++] = (unsigned short)(first + k); continue; }
        if ((link[k] & 0xffffu) != ESB_NOLINK) continue;               // not a plus end
        int b = first + k;
        fplus[nfil] = (unsigned short)b;
        while ((link[b - first] >> 16) != ESB_NOLINK) b = (int)(link[b - first] >> 16);
        fminus[nfil++] = (unsigned short)b;
    }
    lists[0] = (unsigned short)nfree; lists[1] = (unsigned short)nfil;
    std::vector<unsigned int> linkr((size_t)h->R * count);
    std::vector<unsigned short> listr((size_t)h->R * lists.size());
    for (int r = 0; r < h->R; r++) {
        std::copy(link.begin(), link.end(), linkr.begin() + (size_t)r * count);
        std::copy(lists.begin(), lists.end(), listr.begin() + (size_t)r * lists.size());
    }
    if ((rc = dev_upload(&h->d_act_link, linkr)) || (rc = dev_upload(&h->d_act_lists, listr))) return rc;
    if (h->n_chunks > 0) {
        if ((rc = dev_alloc(&h->d_act_stats, (size_t)h->R * h->n_chunks * 6)) || (rc = dev_alloc(&h->d_act_hist, (size_t)h->R * h->n_chunks * (ESB_N_SHELL - 1)))) return rc;
        CK(cudaMemset(h->d_act_stats, 0, (size_t)h->R * h->n_chunks * 6 * sizeof(int)));
        CK(cudaMemset(h->d_act_hist, 0, (size_t)h->R * h->n_chunks * (ESB_N_SHELL - 1) * sizeof(unsigned short)));
    }
    return ESB_OK;
}

int esb_get_actin(esb_handle *h, int *plus, int *minus, int *n_free, int *free_list, int *n_fil, int *fil_plus)
{
    if (!h || !h->have_actin) return fail(ESB_ESTATE, "esb_set_actin has not been called");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    const int count = h->act.count;
    std::vector<unsigned int> link((size_t)h->R * count);
    std::vector<unsigned short> lists((size_t)h->R * ESB_ACT_WORDS(count));
    CK(cudaMemcpy(link.data(), h->d_act_link, link.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lists.data(), h->d_act_lists, lists.size() * sizeof(unsigned short), cudaMemcpyDeviceToHost));
    for (int r = 0; r < h->R; r++) {
        const unsigned short *L = lists.data() + (size_t)r * ESB_ACT_WORDS(count);
        for (int k = 0; k < count; k++) {
            const unsigned int w = link[(size_t)r * count + k];
            if (plus) plus[(size_t)r * count + k] = (w & 0xffffu) == ESB_NOLINK ? -1 : (int)(w & 0xffffu);
            if (minus) minus[(size_t)r * count + k] = (w >> 16) == ESB_NOLINK ? -1 : (int)(w >> 16);
            if (free_list) free_list[(size_t)r * count + k] = k < L[0] ? (int)L[2 + k] : -1;
            if (fil_plus) fil_plus[(size_t)r * count + k] = k < L[1] ? (int)L[2 + count + k] : -1;
        }
        if (n_free) n_free[r] = L[0];
        if (n_fil) n_fil[r] = L[1];
    }
    return ESB_OK;
}

int esb_get_actin_frames(esb_handle *h, int chunk_begin, int n_chunks, int *stats, int *hist)
{
    if (!h || !stats) return fail(ESB_EINVAL, "NULL argument");
    if (!h->have_actin || !h->d_act_stats) return fail(ESB_ESTATE, "no actin records (esb_set_actin + esb_set_schedule first)");
    if (chunk_begin < 0 || n_chunks < 1 || chunk_begin + n_chunks > h->n_chunks) return fail(ESB_EINVAL, "chunk range outside the schedule");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    CK(cudaMemcpy2D(stats, (size_t)n_chunks * 6 * sizeof(int), h->d_act_stats + (size_t)chunk_begin * 6, (size_t)h->n_chunks * 6 * sizeof(int),
                    (size_t)n_chunks * 6 * sizeof(int), h->R, cudaMemcpyDeviceToHost));
    if (hist) {
        const int nb = ESB_N_SHELL - 1;
        std::vector<unsigned short> tmp((size_t)h->R * n_chunks * nb);
        CK(cudaMemcpy2D(tmp.data(), (size_t)n_chunks * nb * 2, h->d_act_hist + (size_t)chunk_begin * nb, (size_t)h->n_chunks * nb * 2,
                        (size_t)n_chunks * nb * 2, h->R, cudaMemcpyDeviceToHost));
        for (size_t q = 0; q < tmp.size(); q++) hist[q] = tmp[q];
    }
    return ESB_OK;
}

int esb_set_microscopy(esb_handle *h, int n_x_edges, const double *x_edges, int n_y_edges, const double *y_edges,
                       int n_z_edges, const double *z_edges, int n_frames)
{
    if (!h || !x_edges || !y_edges || !z_edges || n_frames < 1) return fail(ESB_EINVAL, "bad argument");
    const int ne[3] = {n_x_edges, n_y_edges, n_z_edges};
    const double *e[3] = {x_edges, y_edges, z_edges};
    for (int d = 0; d < 3; d++) {
        if (ne[d] < 2 || ne[d] > 32) return fail(ESB_EINVAL, "axis %d: %d bin edges (2..32)", d, ne[d]);
        for (int k = 1; k < ne[d]; k++) if (!(e[d][k] > e[d][k - 1])) return fail(ESB_EINVAL, "axis %d: edges must increase", d);
    }
    const size_t nvox = (size_t)(ne[0] - 1) * (ne[1] - 1) * (ne[2] - 1);
    if (nvox > (size_t)6 * h->n_pad) return fail(ESB_EINVAL, "%zu voxels do not fit the histogram scratch (%d)", nvox, 6 * h->n_pad);
    CK(cudaSetDevice(h->device));
    for (int d = 0; d < 3; d++) { h->mic_n[d] = ne[d]; memcpy(h->mic_edges[d], e[d], sizeof(double) * ne[d]); }
    h->mic_frames = n_frames;
    int rc;
    if ((rc = dev_alloc(&h->d_mic, (size_t)h->R * n_frames * 5 * nvox))) return rc;
    CK(cudaMemset(h->d_mic, 0, (size_t)h->R * n_frames * 5 * nvox * sizeof(unsigned short)));
    return ESB_OK;
}

int esb_get_microscopy(esb_handle *h, int frame_begin, int n_frames, unsigned short *out)
{
    if (!h || !out) return fail(ESB_EINVAL, "NULL argument");
    if (!h->d_mic) return fail(ESB_ESTATE, "esb_set_microscopy has not been called");
    if (frame_begin < 0 || n_frames < 1 || frame_begin + n_frames > h->mic_frames) return fail(ESB_EINVAL, "frame range outside the buffer");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    const size_t per = (size_t)5 * (h->mic_n[0] - 1) * (h->mic_n[1] - 1) * (h->mic_n[2] - 1) * sizeof(unsigned short);
    CK(cudaMemcpy2D(out, (size_t)n_frames * per, h->d_mic + (size_t)frame_begin * per / sizeof(unsigned short), (size_t)h->mic_frames * per,
                    (size_t)n_frames * per, h->R, cudaMemcpyDeviceToHost));
    return ESB_OK;
}

int esb_set_neighbor(esb_handle *h, double skin, int delay)
{
    if (!h || !(skin > 0) || delay < 0) return fail(ESB_EINVAL, "bad neighbour parameters (skin %g, delay %d)", skin, delay);
    h->skin = skin; h->neigh_delay = delay; h->setups_dirty = true;
    return ESB_OK;
}

int esb_set_soft(esb_handle *h, const double *cut)
{
    if (!h || !cut) return fail(ESB_EINVAL, "NULL argument");
    memcpy(h->soft_cut, cut, sizeof(h->soft_cut));
    h->have_soft = true; h->setups_dirty = true;
    return ESB_OK;
}

int esb_set_tables(esb_handle *h, int n_tables, int tlm1, const double *f, const double *e, const double *innersq,
                   const double *delta, const double *cut, const esb_closed_form *cf)
{
    if (!h || n_tables < 1 || n_tables > ESB_MAX_TABLES || tlm1 < 1 || !f || !e || !innersq || !delta || !cut)
        return fail(ESB_EINVAL, "bad table arguments (n_tables=%d, max %d)", n_tables, ESB_MAX_TABLES);
    CK(cudaSetDevice(h->device));
    h->tables.assign(n_tables, HostTable());
    h->tlm1 = tlm1;
    std::vector<float> fv((size_t)n_tables * tlm1), ev((size_t)n_tables * tlm1);
    for (int t = 0; t < n_tables; t++) {
        HostTable &T = h->tables[t];
        T.f.assign(f + (size_t)t * tlm1, f + (size_t)(t + 1) * tlm1);
        T.e.assign(e + (size_t)t * tlm1, e + (size_t)(t + 1) * tlm1);
        T.innersq = innersq[t]; T.delta = delta[t]; T.cut = cut[t];
        if (cf) T.cf = cf[t]; else { memset(&T.cf, 0, sizeof(T.cf)); }
        if (!(T.delta > 0) || !(T.cut > 0)) return fail(ESB_EINVAL, "table %d: bad delta/cut", t);
        for (int i = 0; i < tlm1; i++) { fv[(size_t)t * tlm1 + i] = (float)T.f[i]; ev[(size_t)t * tlm1 + i] = (float)T.e[i]; }
    }
    int rc;
    if ((rc = dev_upload(&h->d_tabval, fv)) || (rc = dev_upload(&h->d_tabe, ev))) return rc;
    h->have_tables = true; h->setups_dirty = true;
    return ESB_OK;
}

// Cubic-spline helpers of the table construction (Numerical-Recipes style second-derivative
// spline with given end slopes, as used by LAMMPS' PairTable).
static void spline_second_derivs(const std::vector<double> &x, const double *y, double yp1, double ypn, std::vector<double> &y2)
{
    const int n = (int)x.size();
    std::vector<double> u(n);
    y2.assign(n, 0.0);
    if (yp1 > 0.99e30) { y2[0] = 0.0; u[0] = 0.0; }
    else { y2[0] = -0.5; u[0] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1); }
    for (int i = 1; i < n - 1; i++) {
        const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
        const double p = sig * y2[i - 1] + 2.0;
        y2[i] = (sig - 1.0) / p;
        u[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
        u[i] = (6.0 * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
    }
    double qn = 0.0, un = 0.0;
    if (!(ypn > 0.99e30)) { qn = 0.5; un = (3.0 / (x[n - 1] - x[n - 2])) * (ypn - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2])); }
    y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.0);
    for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
}

static double spline_eval(const std::vector<double> &xa, const double *ya, const std::vector<double> &y2a, double x)
{
    int klo = 0, khi = (int)xa.size() - 1;
    while (khi - klo > 1) {
        const int k = (khi + klo) >> 1;
        if (xa[k] > x) khi = k; else klo = k;
    }
    const double hh = xa[khi] - xa[klo];
    const double a = (xa[khi] - x) / hh, b = (x - xa[klo]) / hh;
    return a * ya[klo] + b * ya[khi] + ((a * a * a - a) * y2a[klo] + (b * b * b - b) * y2a[khi]) * (hh * hh) / 6.0;
}

int esb_table_from_section(int n_input, double rlo, double rhi, const double *e_file, const double *f_file,
                           double cut, int tablength, double *e_out, double *f_out, double *params_out)
{
    if (n_input <= 1 || !e_file || !f_file || !e_out || !f_out || !params_out || tablength < 2)
        return fail(ESB_EINVAL, "Invalid pair table length");
    if (rlo <= 0.0) return fail(ESB_EINVAL, "Invalid pair table lower boundary");
    if (cut <= rlo || cut > rhi) return fail(ESB_EINVAL, "Pair table cutoff outside of table");
    std::vector<double> r(n_input), e2, f2;
    for (int i = 0; i < n_input; i++) r[i] = sqrt(rlo * rlo + (rhi * rhi - rlo * rlo) * i / (n_input - 1));
    spline_second_derivs(r, e_file, -f_file[0], -f_file[n_input - 1], e2);
    const double fplo = (f_file[1] - f_file[0]) / (r[1] - r[0]);
    const double fphi = (f_file[n_input - 1] - f_file[n_input - 2]) / (r[n_input - 1] - r[n_input - 2]);
    spline_second_derivs(r, f_file, fplo, fphi, f2);
    const int tlm1 = tablength - 1;
    const double innersq = rlo * rlo;
    const double delta = (cut * cut - innersq) / tlm1;
    for (int i = 0; i < tlm1; i++) {
        const double rr = sqrt(innersq + (i + 0.5) * delta);
        e_out[i] = spline_eval(r, e_file, e2, rr);
        f_out[i] = spline_eval(r, f_file, f2, rr) / rr;
    }
    params_out[0] = innersq; params_out[1] = delta; params_out[2] = 1.0 / delta;
    return ESB_OK;
}

int esb_set_table_map(esb_handle *h, int map_id, const int *map)
{
    if (!h || !map || map_id < 0 || map_id >= 4) return fail(ESB_EINVAL, "bad map_id %d", map_id);
    if (!h->have_tables) return fail(ESB_ESTATE, "esb_set_tables must precede esb_set_table_map");
    for (int a = 1; a <= h->n_types; a++)
        for (int b = 1; b <= h->n_types; b++) {
            const int t = map[a * ESB_MAXT + b];
            if (t < 0 || t >= (int)h->tables.size()) return fail(ESB_EINVAL, "map entry (%d,%d) = %d is not a table", a, b, t);
            if (t != map[b * ESB_MAXT + a]) return fail(ESB_EINVAL, "map is not symmetric at (%d,%d)", a, b);
        }
    memcpy(h->maps[map_id], map, sizeof(h->maps[map_id]));
    h->have_map[map_id] = true; h->setups_dirty = true;
    return ESB_OK;
}

int esb_set_layout(esb_handle *h, int gene_start, int promoter_length, int n_enhancer, const int *enhancer, int n_ser5p,
                   const int *ser5p, const double *cluster_r)
{
    if (!h || !enhancer || !ser5p || !cluster_r) return fail(ESB_EINVAL, "NULL argument");
    if (gene_start - promoter_length < 0 || gene_start + 5 > h->N || promoter_length < 1 || gene_start - promoter_length + 2 >= h->N)
        return fail(ESB_EINVAL, "gene_start=%d promoter_length=%d does not fit %d atoms", gene_start, promoter_length, h->N);
    CK(cudaSetDevice(h->device));
    std::vector<int> en(enhancer, enhancer + n_enhancer), s5(ser5p, ser5p + n_ser5p);
    for (int v : en) if (v < 0 || v >= h->N) return fail(ESB_EINVAL, "enhancer index %d out of range", v);
    for (int v : s5) if (v < 0 || v >= h->N) return fail(ESB_EINVAL, "ser5p index %d out of range", v);
    std::vector<double> cr(cluster_r, cluster_r + ESB_N_SHELL);
    int rc;
    if ((rc = dev_upload(&h->d_enh, en)) || (rc = dev_upload(&h->d_s5p, s5)) || (rc = dev_upload(&h->d_cluster_r, cr))) return rc;
    h->lay.gene_start = gene_start; h->lay.promoter_length = promoter_length; h->lay.n_enh = n_enhancer; h->lay.n_s5p = n_ser5p;
    h->have_layout = true;
    return ESB_OK;
}

int esb_set_replica_params(esb_handle *h, const uint64_t *seed, const int *threshold, const double *p_activation)
{
    if (!h) return fail(ESB_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    int rc = pull_rs(h);
    if (rc) return rc;
    for (int r = 0; r < h->R; r++) {
        ReplicaScalars &rs = h->rs_host[r];
        if (seed) rs.seed = seed[r];
        if (threshold) rs.threshold = threshold[r];
        if (p_activation) { rs.p_activation_d = p_activation[r]; rs.p_activation = (float)p_activation[r]; }
    }
    return push_rs(h);
}

int esb_upload_state(esb_handle *h, const double *x, const double *v, const int *types, int replicate)
{
    if (!h || !x || !types) return fail(ESB_EINVAL, "NULL argument");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    const size_t nsrc = (size_t)(replicate ? 1 : h->R) * h->N;
    Scratch<double> dx, dv;
    Scratch<int> dt;
    CK(dx.alloc(nsrc * 3));
    CK(dt.alloc(nsrc));
    CK(cudaMemcpy(dx.p, x, nsrc * 3 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dt.p, types, nsrc * sizeof(int), cudaMemcpyHostToDevice));
    if (v) {
        CK(dv.alloc(nsrc * 3));
        CK(cudaMemcpy(dv.p, v, nsrc * 3 * sizeof(double), cudaMemcpyHostToDevice));
    }
    esb_pack_kernel<<<512, 256>>>(dx.p, dv.p, dt.p, h->d_frozen, h->d_pos4, h->d_vel4, h->N, h->n_pad, h->R, replicate);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    h->have_state = true;
    return ESB_OK;
}

int esb_download_state(esb_handle *h, double *x, double *v, int *types)
{
    if (!h) return fail(ESB_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    const size_t n = (size_t)h->R * h->N;
    Scratch<double> dx, dv;
    Scratch<int> dt;
    if (x) CK(dx.alloc(n * 3));
    if (v) CK(dv.alloc(n * 3));
    if (types) CK(dt.alloc(n));
    esb_unpack_kernel<<<512, 256>>>(h->d_pos4, h->d_vel4, dx.p, dv.p, dt.p, h->N, h->n_pad, h->R);
    CK(cudaGetLastError());
    if (x) CK(cudaMemcpy(x, dx.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (v) CK(cudaMemcpy(v, dv.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (types) CK(cudaMemcpy(types, dt.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    return ESB_OK;
}

int esb_n_pad(const esb_handle *h) { return h ? h->n_pad : 0; }

int esb_upload_state_f32(esb_handle *h, const float *pos4, const float *vel4, void *stream)
{
    if (!h || !pos4) return fail(ESB_EINVAL, "NULL argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tot = (size_t)h->R * h->n_pad;
    CK(cudaMemcpyAsync(h->d_pos4, pos4, tot * sizeof(float4), cudaMemcpyHostToDevice, st));
    if (vel4) CK(cudaMemcpyAsync(h->d_vel4, vel4, tot * sizeof(float4), cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(h->d_vel4, 0, tot * sizeof(float4), st));
    esb_mark_frozen_kernel<<<256, 256, 0, st>>>(h->d_pos4, h->d_frozen, h->n_pad, tot);
    CK(cudaGetLastError());
    h->have_state = true;
    return ESB_OK;
}

int esb_download_state_f32(esb_handle *h, float *pos4, float *vel4, void *stream)
{
    if (!h) return fail(ESB_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tot = (size_t)h->R * h->n_pad;
    if (pos4) CK(cudaMemcpyAsync(pos4, h->d_pos4, tot * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (vel4) CK(cudaMemcpyAsync(vel4, h->d_vel4, tot * sizeof(float4), cudaMemcpyDeviceToHost, st));
    return ESB_OK;
}

int esb_set_schedule(esb_handle *h, int n_chunks, const esb_chunk_desc *descs)
{
    if (!h || n_chunks < 1 || !descs) return fail(ESB_EINVAL, "bad schedule");
    CK(cudaSetDevice(h->device));
    for (int c = 0; c < n_chunks; c++) {
        const esb_chunk_desc &d = descs[c];
        if (d.pair_mode < 1 || d.pair_mode > 2 || d.n_steps < 0 || d.map_id < 0 || d.map_id >= 4)
            return fail(ESB_EINVAL, "chunk %d: bad pair_mode/map_id/n_steps", c);
        if (d.ramp_on && d.ramp_stop == d.ramp_start) return fail(ESB_EINVAL, "chunk %d: empty ramp interval", c);
    }
    std::vector<esb_chunk_desc> v(descs, descs + n_chunks);
    int rc;
    if ((rc = dev_upload(&h->d_descs, v))) return rc;
    if (n_chunks != h->n_chunks || !h->d_frames) {
        if ((rc = dev_alloc(&h->d_frames, (size_t)h->R * n_chunks * ESB_FRAME_DOUBLES))) return rc;
        if ((rc = dev_alloc(&h->d_hist, (size_t)h->R * n_chunks * (ESB_N_SHELL - 1)))) return rc;
        if ((rc = dev_alloc(&h->d_rawd, (size_t)h->R * n_chunks * 2))) return rc;
        CK(cudaMemset(h->d_rawd, 0, (size_t)h->R * n_chunks * 2 * sizeof(double)));
        CK(cudaMemset(h->d_frames, 0, (size_t)h->R * n_chunks * ESB_FRAME_DOUBLES * sizeof(double)));
        CK(cudaMemset(h->d_hist, 0, (size_t)h->R * n_chunks * (ESB_N_SHELL - 1) * sizeof(unsigned short)));
    }
    if (h->have_actin && (n_chunks != h->n_chunks || !h->d_act_stats)) {
        if ((rc = dev_alloc(&h->d_act_stats, (size_t)h->R * n_chunks * 6)) || (rc = dev_alloc(&h->d_act_hist, (size_t)h->R * n_chunks * (ESB_N_SHELL - 1)))) return rc;
        CK(cudaMemset(h->d_act_stats, 0, (size_t)h->R * n_chunks * 6 * sizeof(int)));
        CK(cudaMemset(h->d_act_hist, 0, (size_t)h->R * n_chunks * (ESB_N_SHELL - 1) * sizeof(unsigned short)));
    }
    h->n_chunks = n_chunks;
    h->descs_host = v;             // validated against the pair set-ups at run time
    return ESB_OK;
}

static int prepare_launch(esb_handle *h, size_t *smem)
{
    if (!h->have_state) return fail(ESB_ESTATE, "no state uploaded");
    if (h->setups_dirty) {
        int rc = rebuild_setups(h);
        if (rc) return rc;
    }
    if ((size_t)12 * h->n_pad < (size_t)(ESB_NL_MAXNB + 1) * 4)
        return fail(ESB_EINVAL, "replica of %d atoms is too small: the list build keeps %d counters in the force arrays (>= %d beads)",
                    h->N, ESB_NL_MAXNB + 1, (ESB_NL_MAXNB + 1) / 3 + 1);
    *smem = esb_smem_bytes_v512(h->n_pad, (int)h->tables.size(), h->n_topo_pad);
    if (*smem > 227 * 1024) return fail(ESB_EINVAL, "replica of %d atoms needs %zu B of shared memory (max 232448)", h->N, *smem);
    return ESB_OK;
}

int esb_run(esb_handle *h, int chunk_begin, int n_chunks, void *stream)
{
    if (!h) return fail(ESB_EINVAL, "NULL handle");
    if (chunk_begin < 0 || n_chunks < 1 || chunk_begin + n_chunks > h->n_chunks)
        return fail(ESB_EINVAL, "chunks [%d,%d) outside the schedule of %d", chunk_begin, chunk_begin + n_chunks, h->n_chunks);
    CK(cudaSetDevice(h->device));
    if (!h->have_layout) return fail(ESB_ESTATE, "esb_set_layout must precede esb_run");
    size_t smem = 0;
    int rc = prepare_launch(h, &smem);
    if (rc) return rc;
    for (int c = chunk_begin; c < chunk_begin + n_chunks; c++) {      // a chunk whose pair set-up was never provided would run without pair forces
        const esb_chunk_desc &d = h->descs_host[c];
        if (d.pair_mode == 1 && !h->have_soft) return fail(ESB_ESTATE, "chunk %d runs pair_style soft but esb_set_soft was not called", c);
        if (d.pair_mode == 2 && (!h->have_tables || !h->have_map[d.map_id]))
            return fail(ESB_ESTATE, "chunk %d runs pair_style table with map %d but %s", c, d.map_id, h->have_tables ? "that map was not set (esb_set_table_map)" : "no tables were set (esb_set_tables)");
    }
    KArgs k;
    fill_kargs(h, k, false);
    // a batch that cannot fill the device with two CTAs per SM gets the 32-warp CTA: a replica then advances twice as fast
    const int n_seg = getenv("ESB_SEG") ? atoi(getenv("ESB_SEG")) : h->n_seg;      // 0: the launcher decides
    // (also a system too large for two 16-warp CTAs per SM -- more than ~1700 beads: one 32-warp CTA fills the SM better than one 16-warp CTA)
    const bool two_fit = 2 * (smem + 1024) <= h->smem_per_sm;
    const bool wide = getenv("ESB_WIDE") ? atoi(getenv("ESB_WIDE")) != 0 : h->force_wide >= 0 ? (h->force_wide != 0 || !two_fit) : (h->R <= h->n_sm || !two_fit);
    // even smaller batches: a thread-block cluster of 2, 4 or 8 CTAs (SMs) per replica
    int cluster = 1;
    if (getenv("ESB_CLUSTER_SIZE")) cluster = atoi(getenv("ESB_CLUSTER_SIZE"));
    else if (wide && h->force_wide < 0 && h->R <= h->n_sm) for (int c = 8; c >= 2; c >>= 1) if ((long long)h->R * c <= (long long)h->n_sm * 7 / 8) { cluster = c; break; }
    if (cluster == 2 || cluster == 4 || cluster == 8) {
        const size_t smem_c = esb_smem_bytes_v2048(h->n_pad, (int)h->tables.size(), h->n_topo_pad);
        CK(esb_launch_run_v2048(k, chunk_begin, n_chunks, 0.03 * 200, 60.0, smem_c, (cudaStream_t)stream, h->d_sched, h->n_sm, cluster));
    } else if (wide) {
        const size_t smem_w = esb_smem_bytes_v1024(h->n_pad, (int)h->tables.size(), h->n_topo_pad);
        CK(esb_launch_run_v1024(k, chunk_begin, n_chunks, 0.03 * 200, 60.0, smem_w, (cudaStream_t)stream, h->d_sched, h->n_sm, n_seg));
    } else {
        CK(esb_launch_run_v512(k, chunk_begin, n_chunks, 0.03 * 200, 60.0, smem, (cudaStream_t)stream, h->d_sched, h->n_sm, n_seg));
    }
    h->last_stream = (cudaStream_t)stream;
    h->run_pending = true;
    return ESB_OK;
}

int esb_set_segments(esb_handle *h, int n_seg, int co_scheduled)
{
    if (!h) return fail(ESB_EINVAL, "NULL handle");
    if (n_seg < 0 || n_seg > 64) return fail(ESB_EINVAL, "n_seg %d outside [0, 64]", n_seg);
    h->n_seg = n_seg;
    h->force_wide = co_scheduled ? 0 : -1;
    return ESB_OK;
}

int esb_sync(esb_handle *h)
{
    if (!h) return fail(ESB_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    return ESB_OK;
}

int esb_forces(esb_handle *h, int pair_mode, int map_id, double soft_a, int with_post, uint64_t eval, int use_arrays,
               double *f_out, double *e_out)
{
    if (!h || !f_out) return fail(ESB_EINVAL, "NULL argument");
    if (pair_mode < 0 || pair_mode > 2 || map_id < 0 || map_id >= 4) return fail(ESB_EINVAL, "bad pair_mode/map_id");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    size_t smem = 0;
    int rc = prepare_launch(h, &smem);
    if (rc) return rc;
    if (pair_mode == 2 && !h->have_map[map_id]) return fail(ESB_ESTATE, "table map %d not set", map_id);
    if (pair_mode == 1 && !h->have_soft) return fail(ESB_ESTATE, "soft cutoffs not set");
    KArgs k;
    fill_kargs(h, k, use_arrays != 0);
    Scratch<double> df, de;
    CK(df.alloc((size_t)h->R * h->N * 3));
    CK(de.alloc((size_t)h->R * 4));
    const int setup_id = pair_mode == 2 ? 1 + map_id + (use_arrays ? 4 : 0) : 0;
    CK(esb_launch_forces(k, pair_mode, setup_id, (float)soft_a, with_post, eval, df.p, de.p, smem, 0));
    CK(cudaMemcpy(f_out, df.p, (size_t)h->R * h->N * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (e_out) CK(cudaMemcpy(e_out, de.p, (size_t)h->R * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return ESB_OK;
}

int esb_get_frames(esb_handle *h, int chunk_begin, int n_chunks, double *out, int *hist)
{
    if (!h || !out) return fail(ESB_EINVAL, "NULL argument");
    if (chunk_begin < 0 || n_chunks < 1 || chunk_begin + n_chunks > h->n_chunks) return fail(ESB_EINVAL, "chunk range outside the schedule");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    CK(cudaMemcpy2D(out, (size_t)n_chunks * ESB_FRAME_DOUBLES * sizeof(double),
                    h->d_frames + (size_t)chunk_begin * ESB_FRAME_DOUBLES, (size_t)h->n_chunks * ESB_FRAME_DOUBLES * sizeof(double),
                    (size_t)n_chunks * ESB_FRAME_DOUBLES * sizeof(double), h->R, cudaMemcpyDeviceToHost));
    if (hist) {
        const int nb = ESB_N_SHELL - 1;
        std::vector<unsigned short> tmp((size_t)h->R * n_chunks * nb);
        CK(cudaMemcpy2D(tmp.data(), (size_t)n_chunks * nb * 2, h->d_hist + (size_t)chunk_begin * nb, (size_t)h->n_chunks * nb * 2,
                        (size_t)n_chunks * nb * 2, h->R, cudaMemcpyDeviceToHost));
        for (size_t q = 0; q < tmp.size(); q++) hist[q] = tmp[q];
    }
    return ESB_OK;
}

int esb_get_raw_distances(esb_handle *h, int chunk_begin, int n_chunks, double *out)
{
    if (!h || !out) return fail(ESB_EINVAL, "NULL argument");
    if (chunk_begin < 0 || n_chunks < 1 || chunk_begin + n_chunks > h->n_chunks) return fail(ESB_EINVAL, "chunk range outside the schedule");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    CK(cudaMemcpy2D(out, (size_t)n_chunks * 2 * sizeof(double), h->d_rawd + (size_t)chunk_begin * 2, (size_t)h->n_chunks * 2 * sizeof(double),
                    (size_t)n_chunks * 2 * sizeof(double), h->R, cudaMemcpyDeviceToHost));
    return ESB_OK;
}

int esb_get_frames_async(esb_handle *h, int chunk_begin, int n_chunks, double *out, void *stream)
{
    if (!h || !out) return fail(ESB_EINVAL, "NULL argument");
    if (chunk_begin < 0 || n_chunks < 1 || chunk_begin + n_chunks > h->n_chunks) return fail(ESB_EINVAL, "chunk range outside the schedule");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy2DAsync(out, (size_t)n_chunks * ESB_FRAME_DOUBLES * sizeof(double),
                         h->d_frames + (size_t)chunk_begin * ESB_FRAME_DOUBLES, (size_t)h->n_chunks * ESB_FRAME_DOUBLES * sizeof(double),
                         (size_t)n_chunks * ESB_FRAME_DOUBLES * sizeof(double), h->R, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return ESB_OK;
}

int esb_get_status(esb_handle *h, int *status)
{
    if (!h || !status) return fail(ESB_EINVAL, "NULL argument");
    CK(cudaSetDevice(h->device));
    int rc = pull_rs(h);
    if (rc) return rc;
    for (int r = 0; r < h->R; r++) status[r] = h->rs_host[r].status;
    return ESB_OK;
}

int esb_get_counters(esb_handle *h, int64_t *out)
{
    if (!h || !out) return fail(ESB_EINVAL, "NULL argument");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    CK(cudaMemcpy(out, h->d_counters, (size_t)h->R * ESB_N_COUNTERS * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return ESB_OK;
}

int esb_reset_counters(esb_handle *h)
{
    if (!h) return fail(ESB_EINVAL, "NULL handle");
    CK(cudaSetDevice(h->device));
    { const int rc_ = wait_for_run(h); if (rc_) return rc_; }
    CK(cudaMemset(h->d_counters, 0, (size_t)h->R * ESB_N_COUNTERS * sizeof(int64_t)));
    return ESB_OK;
}

int esb_get_bond_r0(esb_handle *h, double r0[3])
{
    if (!h || !r0) return fail(ESB_EINVAL, "NULL argument");
    CK(cudaSetDevice(h->device));
    int rc = pull_rs(h);
    if (rc) return rc;
    r0[0] = 0.0; r0[1] = h->rs_host[0].r0[1]; r0[2] = h->rs_host[0].r0[2];
    return ESB_OK;
}

int64_t esb_get_step(const esb_handle *h)
{
    if (!h) return -1;
    esb_handle *hh = const_cast<esb_handle *>(h);
    cudaSetDevice(hh->device);
    if (pull_rs(hh)) return -1;
    return hh->rs_host[0].step;
}

void *esb_device_ptr(esb_handle *h, int which)
{
    if (!h) return nullptr;
    switch (which) {
        case 0: return h->d_pos4;
        case 1: return h->d_vel4;
        case 2: return h->d_frames;
        case 3: return h->d_rs;
        case 4: return h->d_counters;
        default: return nullptr;
    }
}

}  // extern "C"
