// esb_aggregate.cu -- dist_genestages_grouped.py on the device.
//
// Reference: VisitorGeneMaps/dist_genestages_grouped.py:70-106 (5-run groups; the root copy is
// identical in this part).  Per run (geneTrack columns x[4]=d_rp, x[6]=s5p_promoter,
// x[8]=gene_state):
//     I = [k for k in 1..L-1 if int(state[k]) == 2 and int(state[k-1]) != 2]
//     J = [k for k in 1..L-1 if int(state[k]) == 2]
//     s5p_group.extend(s5p)                       active.append(100*len(J)/(L-150))
//     contact.append(100*sum(d_rp < 250)/L)       act_dist.extend(d_rp[k] for k in I)
// and on every 5th run one row = np.mean of each list.  Bit-exactness with numpy means
// reproducing np.add.reduce's pairwise summation (blocks of <=128 elements, 8 running sums,
// halves rounded down to a multiple of 8) and dividing by the element count at the end; an empty
// list gives NaN.  One CTA per group: per-run integer counts by block reduction, leaf blocks of the
// pairwise tree summed in parallel, the tree combined by one thread in numpy's order.
#include <cmath>
#include <cstdio>
#include <vector>

#include "esb_device.cuh"

namespace {

constexpr int AGG_THREADS = 256;
constexpr int AGG_MAX_LEAVES = 2048;
constexpr int PW_BLOCK = 128;

struct AggArgs {
    const double *frames;   // [n_runs][n_frames][9]
    double *scratch;        // [n_groups][group*n_frames]
    double *out;            // [n_groups][4]
    int n_frames, group, t_skip;
    double contact_dist;
};

// numpy's leaf: n < 8 sequential; 8 <= n <= 128 with eight running sums.
template <class Get>
__device__ double pw_leaf(Get get, long long off, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; i++) res = __dadd_rn(res, get(off + i));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = get(off + j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], get(off + i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __dadd_rn(res, get(off + i));
    return res;
}

// Leaves of pairwise_sum(a, n) in order.  Returns the leaf count; offsets[k]..offsets[k+1].
__device__ int pw_leaves(long long n, long long *offsets)
{
    // explicit stack of (offset, length); left child first
    long long st_off[64], st_len[64];
    int sp = 0, nl = 0;
    st_off[0] = 0; st_len[0] = n; sp = 1;
    while (sp) {
        sp--;
        const long long off = st_off[sp], len = st_len[sp];
        if (len <= PW_BLOCK) {
            if (nl < AGG_MAX_LEAVES) offsets[nl] = off;
            nl++;
        } else {
            long long n2 = len / 2;
            n2 -= n2 % 8;
            st_off[sp] = off + n2; st_len[sp] = len - n2; sp++;   // right (popped second)
            st_off[sp] = off; st_len[sp] = n2; sp++;              // left (popped first)
        }
    }
    if (nl <= AGG_MAX_LEAVES) offsets[nl] = n;
    return nl;
}

// Combine leaf sums in the order of the recursion: pw(n) = pw(left) + pw(right).
__device__ double pw_combine(long long n, const double *leaf_sum)
{
    // post-order evaluation with an explicit stack of pending frames
    struct Fr { long long len; int stage; double left; };
    Fr st[64];
    int sp = 0, next_leaf = 0;
    double ret = 0.0;
    st[sp++] = {n, 0, 0.0};
    while (sp) {
        Fr &f = st[sp - 1];
        if (f.len <= PW_BLOCK) { ret = leaf_sum[next_leaf++]; sp--; continue; }
        long long n2 = f.len / 2;
        n2 -= n2 % 8;
        if (f.stage == 0) { f.stage = 1; st[sp++] = {n2, 0, 0.0}; }
        else if (f.stage == 1) { f.left = ret; f.stage = 2; st[sp++] = {f.len - n2, 0, 0.0}; }
        else { ret = __dadd_rn(f.left, ret); sp--; }
    }
    return ret;
}

template <class Get>
__device__ double block_pairwise_mean(Get get, long long n, long long *s_off, double *s_sum, int *s_nl)
{
    // all threads call; result valid on thread 0
    if (n == 0) return nan("");
    if (threadIdx.x == 0) *s_nl = pw_leaves(n, s_off);
    __syncthreads();
    const int nl = *s_nl;
    if (nl > AGG_MAX_LEAVES) return nan("");
    for (int k = threadIdx.x; k < nl; k += AGG_THREADS) s_sum[k] = pw_leaf(get, s_off[k], (int)(s_off[k + 1] - s_off[k]));
    __syncthreads();
    double res = 0.0;
    if (threadIdx.x == 0) res = __ddiv_rn(pw_combine(n, s_sum), (double)n);
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(AGG_THREADS) esb_aggregate_kernel(const AggArgs a)
{
    __shared__ long long s_off[AGG_MAX_LEAVES + 1];
    __shared__ double s_sum[AGG_MAX_LEAVES];
    __shared__ int s_nl, s_cnt[2], s_scan[AGG_THREADS];
    __shared__ double s_pct[2][16];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int L = a.n_frames;
    const double *base = a.frames + (size_t)g * a.group * L * ESB_FRAME_DOUBLES;
    // per-run percentages
    for (int run = 0; run < a.group; run++) {
        if (tid < 2) s_cnt[tid] = 0;
        __syncthreads();
        int nj = 0, nc = 0;
        const double *f = base + (size_t)run * L * ESB_FRAME_DOUBLES;
        for (int k = tid; k < L; k += AGG_THREADS) {
            if (k >= 1 && (int)f[(size_t)k * 9 + 8] == 2) nj++;
            if (f[(size_t)k * 9 + 4] < a.contact_dist) nc++;
        }
        if (nj) atomicAdd(&s_cnt[0], nj);
        if (nc) atomicAdd(&s_cnt[1], nc);
        __syncthreads();
        if (tid == 0) {
            s_pct[0][run] = __ddiv_rn((double)(100LL * s_cnt[0]), (double)(L - a.t_skip));
            s_pct[1][run] = __ddiv_rn((double)(100LL * s_cnt[1]), (double)L);
        }
        __syncthreads();
    }
    // S5PInt: mean over the concatenated s5p_promoter columns
    const long long n_all = (long long)a.group * L;
    auto get_s5p = [&](long long k) { return base[(size_t)k * 9 + 6]; };
    const double s5p_mean = block_pairwise_mean(get_s5p, n_all, s_off, s_sum, &s_nl);
    // activation onsets, ordered compaction of d_rp[I] into scratch
    double *scr = a.scratch + (size_t)g * n_all;
    const long long seg = (n_all + AGG_THREADS - 1) / AGG_THREADS;
    const long long k0 = (long long)tid * seg, k1 = min(k0 + seg, n_all);
    int mine = 0;
    for (long long k = k0; k < k1; k++) {
        const int fr = (int)(k % L);
        if (fr >= 1 && (int)base[(size_t)k * 9 + 8] == 2 && (int)base[(size_t)(k - 1) * 9 + 8] != 2) mine++;
    }
    s_scan[tid] = mine;
    __syncthreads();
    if (tid == 0) { int acc = 0; for (int t = 0; t < AGG_THREADS; t++) { const int c = s_scan[t]; s_scan[t] = acc; acc += c; } s_cnt[0] = acc; }
    __syncthreads();
    int w = s_scan[tid];
    for (long long k = k0; k < k1; k++) {
        const int fr = (int)(k % L);
        if (fr >= 1 && (int)base[(size_t)k * 9 + 8] == 2 && (int)base[(size_t)(k - 1) * 9 + 8] != 2) scr[w++] = base[(size_t)k * 9 + 4];
    }
    __syncthreads();
    const long long n_on = s_cnt[0];
    auto get_scr = [&](long long k) { return scr[k]; };
    const double dist_mean = block_pairwise_mean(get_scr, n_on, s_off, s_sum, &s_nl);
    if (tid == 0) {
        auto small_mean = [&](const double *v, int n) {   // np.mean of a list shorter than 8 (sequential)
            double s = 0.0;
            if (n >= 8) {   // general case through the same pairwise rule
                double r[8];
                for (int j = 0; j < 8; j++) r[j] = v[j];
                int i;
                for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], v[i + j]);
                s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                for (; i < n; i++) s = __dadd_rn(s, v[i]);
            } else {
                for (int i = 0; i < n; i++) s = __dadd_rn(s, v[i]);
            }
            return __ddiv_rn(s, (double)n);
        };
        double *o = a.out + (size_t)g * 4;
        o[0] = s5p_mean;
        o[1] = small_mean(s_pct[0], a.group);
        o[2] = small_mean(s_pct[1], a.group);
        o[3] = dist_mean;
    }
}



// ------------------------------------------------------------------------------------------------
// Per-run statistics of VisitorGeneMaps/dist_genestages_grouped_5percentile_10xaveraging.py:78-89
// (also the rows of dist_genestages_all_5percentile.py): mean Ser5P count, S2PInt, Contact, mean
// activation distance (NaN when the gene never activates) and np.percentile(d[t_skip:], q) of d_rg
// and d_rp.  The percentile is numpy's default 'linear' method: the host passes the lower order
// statistic `prev` and the interpolation weight `gamma` computed with numpy's own expression; here
// the tail is bitonic-sorted in shared memory and numpy's _lerp is applied.
// ------------------------------------------------------------------------------------------------
struct RunArgs {
    const double *frames;   // [n_runs][n_frames][9]
    double *scratch;        // [n_runs][n_frames]
    double *out;            // [n_runs][6]
    int n_frames, t_skip, prev, npow2;
    double contact_dist, gamma;
};

__device__ double np_lerp(const double a, const double b, const double t)
{
    const double diff = __dsub_rn(b, a);
    if (t >= 0.5) return __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
    return __dadd_rn(a, __dmul_rn(diff, t));
}

__global__ void __launch_bounds__(AGG_THREADS) esb_aggregate_runs_kernel(const RunArgs a)
{
    extern __shared__ double s_sort[];
    __shared__ long long s_off[AGG_MAX_LEAVES + 1];
    __shared__ double s_sum[AGG_MAX_LEAVES];
    __shared__ int s_nl, s_cnt[2], s_scan[AGG_THREADS];
    const int run = blockIdx.x, tid = threadIdx.x, L = a.n_frames;
    const double *f = a.frames + (size_t)run * L * ESB_FRAME_DOUBLES;
    double *o = a.out + (size_t)run * 6;
    if (tid < 2) s_cnt[tid] = 0;
    __syncthreads();
    int nj = 0, nc = 0;
    for (int k = tid; k < L; k += AGG_THREADS) {
        if (k >= 1 && (int)f[(size_t)k * 9 + 8] == 2) nj++;
        if (f[(size_t)k * 9 + 4] < a.contact_dist) nc++;
    }
    if (nj) atomicAdd(&s_cnt[0], nj);
    if (nc) atomicAdd(&s_cnt[1], nc);
    __syncthreads();
    if (tid == 0) {
        o[1] = __ddiv_rn((double)(100LL * s_cnt[0]), (double)(L - a.t_skip));
        o[2] = __ddiv_rn((double)(100LL * s_cnt[1]), (double)L);
    }
    auto get_s5p = [&](long long k) { return f[(size_t)k * 9 + 6]; };
    const double s5p_mean = block_pairwise_mean(get_s5p, (long long)L, s_off, s_sum, &s_nl);
    // activation onsets in frame order
    double *scr = a.scratch + (size_t)run * L;
    const int seg = (L + AGG_THREADS - 1) / AGG_THREADS;
    const int k0 = tid * seg, k1 = min(k0 + seg, L);
    int mine = 0;
    for (int k = max(k0, 1); k < k1; k++)
        if ((int)f[(size_t)k * 9 + 8] == 2 && (int)f[(size_t)(k - 1) * 9 + 8] != 2) mine++;
    s_scan[tid] = mine;
    __syncthreads();
    if (tid == 0) { int acc = 0; for (int t = 0; t < AGG_THREADS; t++) { const int c = s_scan[t]; s_scan[t] = acc; acc += c; } s_cnt[0] = acc; }
    __syncthreads();
    int w = s_scan[tid];
    for (int k = max(k0, 1); k < k1; k++)
        if ((int)f[(size_t)k * 9 + 8] == 2 && (int)f[(size_t)(k - 1) * 9 + 8] != 2) scr[w++] = f[(size_t)k * 9 + 4];
    __syncthreads();
    auto get_scr = [&](long long k) { return scr[k]; };
    const double dist_mean = block_pairwise_mean(get_scr, (long long)s_cnt[0], s_off, s_sum, &s_nl);
    if (tid == 0) { o[0] = s5p_mean; o[3] = dist_mean; }
    // percentiles of columns 5 (d_rg) and 4 (d_rp) over frames t_skip..L-1
    const int n = L - a.t_skip;
    for (int which = 0; which < 2; which++) {
        const int col = which == 0 ? 5 : 4;
        for (int k = tid; k < a.npow2; k += AGG_THREADS) s_sort[k] = k < n ? f[(size_t)(k + a.t_skip) * 9 + col] : __longlong_as_double(0x7ff0000000000000LL);
        __syncthreads();
        for (int size = 2; size <= a.npow2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int k = tid; k < a.npow2 / 2; k += AGG_THREADS) {
                    const int lo = 2 * k - (k & (stride - 1));
                    const int hi = lo + stride;
                    const bool up = (lo & size) == 0;
                    const double x = s_sort[lo], y = s_sort[hi];
                    if ((x > y) == up) { s_sort[lo] = y; s_sort[hi] = x; }
                }
                __syncthreads();
            }
        if (tid == 0) o[4 + which] = np_lerp(s_sort[a.prev], s_sort[min(a.prev + 1, n - 1)], a.gamma);
        __syncthreads();
    }
}

}  // namespace

extern "C" int esb_aggregate_grouped(int device, const double *frames, int frames_on_device, int n_runs, int n_frames,
                                     int group, double contact_dist, int t_skip, double *out, void *stream)
{
    if (!frames || !out || n_runs < 1 || n_frames < 2 || group < 1 || group > 16 || n_frames <= t_skip) return ESB_EINVAL;
    const int n_groups = n_runs / group;
    if (n_groups == 0) return ESB_OK;
    if ((long long)group * n_frames > (long long)AGG_MAX_LEAVES * 64) return ESB_EINVAL;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return ESB_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return ESB_ECUDA;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nd = (size_t)n_groups * group * n_frames;
    double *d_frames = nullptr, *d_scr = nullptr, *d_out = nullptr;
    int rc = ESB_OK;
    do {
        if (!frames_on_device) {
            if (cudaMalloc(&d_frames, nd * 9 * sizeof(double)) != cudaSuccess) { rc = ESB_ENOMEM; break; }
            if (cudaMemcpyAsync(d_frames, frames, nd * 9 * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = ESB_ECUDA; break; }
        }
        if (cudaMalloc(&d_scr, nd * sizeof(double)) != cudaSuccess || cudaMalloc(&d_out, (size_t)n_groups * 4 * sizeof(double)) != cudaSuccess) { rc = ESB_ENOMEM; break; }
        AggArgs a;
        a.frames = frames_on_device ? frames : d_frames; a.scratch = d_scr; a.out = d_out;
        a.n_frames = n_frames; a.group = group; a.t_skip = t_skip; a.contact_dist = contact_dist;
        esb_aggregate_kernel<<<n_groups, AGG_THREADS, 0, st>>>(a);
        if (cudaGetLastError() != cudaSuccess) { rc = ESB_ECUDA; break; }
        if (cudaMemcpyAsync(out, d_out, (size_t)n_groups * 4 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = ESB_ECUDA; break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = ESB_ECUDA; break; }
    } while (0);
    cudaFree(d_frames); cudaFree(d_scr); cudaFree(d_out);
    return rc;
}

extern "C" int esb_aggregate_runs(int device, const double *frames, int frames_on_device, int n_runs, int n_frames,
                                  double contact_dist, int t_skip, int prev_index, double gamma, double *out, void *stream)
{
    if (!frames || !out || n_runs < 1 || n_frames < 2 || n_frames <= t_skip + 1 || prev_index < 0 || prev_index >= n_frames - t_skip) return ESB_EINVAL;
    if ((long long)n_frames > (long long)AGG_MAX_LEAVES * 64) return ESB_EINVAL;
    int npow2 = 1;
    while (npow2 < n_frames - t_skip) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * sizeof(double);
    if (smem > 160 * 1024) return ESB_EINVAL;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return ESB_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return ESB_ECUDA;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nd = (size_t)n_runs * n_frames;
    double *d_frames = nullptr, *d_scr = nullptr, *d_out = nullptr;
    int rc = ESB_OK;
    do {
        if (!frames_on_device) {
            if (cudaMalloc(&d_frames, nd * 9 * sizeof(double)) != cudaSuccess) { rc = ESB_ENOMEM; break; }
            if (cudaMemcpyAsync(d_frames, frames, nd * 9 * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = ESB_ECUDA; break; }
        }
        if (cudaMalloc(&d_scr, nd * sizeof(double)) != cudaSuccess || cudaMalloc(&d_out, (size_t)n_runs * 6 * sizeof(double)) != cudaSuccess) { rc = ESB_ENOMEM; break; }
        RunArgs a;
        a.frames = frames_on_device ? frames : d_frames; a.scratch = d_scr; a.out = d_out;
        a.n_frames = n_frames; a.t_skip = t_skip; a.prev = prev_index; a.npow2 = npow2; a.contact_dist = contact_dist; a.gamma = gamma;
        if (cudaFuncSetAttribute(esb_aggregate_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { rc = ESB_ECUDA; break; }
        esb_aggregate_runs_kernel<<<n_runs, AGG_THREADS, smem, st>>>(a);
        if (cudaGetLastError() != cudaSuccess) { rc = ESB_ECUDA; break; }
        if (cudaMemcpyAsync(out, d_out, (size_t)n_runs * 6 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = ESB_ECUDA; break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = ESB_ECUDA; break; }
    } while (0);
    cudaFree(d_frames); cudaFree(d_scr); cudaFree(d_out);
    return rc;
}
