// esb_kernels.cu -- sm_100a kernels of the EnhancerShoeBox hot path.
//
// esb_run_kernel      persistent; one CTA steps one replica at a time (the whole replica is resident in shared
//                     memory); work items = (replica, chunk[, segment]) from a ticket counter.  Per chunk:
//                     set-up force evaluation + n_steps x {integrate, (re)build list, forces}, then the per-chunk
//                     measurements / gene state machine / frame record, the actin events and the synthetic-
//                     microscopy voxel counts.  This file is compiled three times (ESB_TAG 512 / 1024: CTA size; 2048: clusters).
// esb_forces_kernel   parity hook: one force evaluation (+ energies) at the stored positions.
// esb_pack/unpack     double <-> float4 state conversion for the host-facing ABI.
//
// The kernels are templates on the bead capacity NP (n_pad) so that every shared-memory array
// offset is an immediate; the phases are functions that take the __grid_constant__ kernel arguments by reference:
// the force and the per-atom phase are inlined into the kernel, the list build, the observation and the actin events
// are __noinline__ (ESB_PHASE_INLINE).
// Reference semantics (what each phase stands in for) are cited at the device functions.
#include <type_traits>

#include "esb_device.cuh"

// ESB_CLUSTER = 1 (third compiled variant, ESB_TAG 2048): one replica is stepped by a thread-block CLUSTER of 2, 4 or 8
// CTAs on as many SMs.  Every CTA keeps the whole replica in its own shared memory; the two expensive loops (candidate
// scan of the list build, pair loop of the force phase) are split by quad, every bead is integrated by the CTA that
// owns its quad, and new positions / velocities are written straight into the shared memory of all CTAs of the cluster
// (distributed shared memory), two cluster barriers per step.  For batches far smaller than the SM count.
#ifndef ESB_CLUSTER
#define ESB_CLUSTER 0
#endif
#if ESB_CLUSTER
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#endif

#ifndef ESB_PHASE_INLINE
#define ESB_PHASE_INLINE 2               // 0 = every phase its own function, 1 = list build, force and per-atom phase inlined into the kernel, 2 = force and per-atom phase only (measured best: +3.5 % over 0)
#endif
#if ESB_PHASE_INLINE == 1
#define ESB_PHASE_B __forceinline__
#define ESB_PHASE_F __forceinline__
#elif ESB_PHASE_INLINE == 2
#define ESB_PHASE_B __noinline__
#define ESB_PHASE_F __forceinline__
#else
#define ESB_PHASE_B __noinline__
#define ESB_PHASE_F __noinline__
#endif

namespace {

#if ESB_CLUSTER
__device__ __forceinline__ int crank() { return (int)cg::this_cluster().block_rank(); }
__device__ __forceinline__ int csize() { return (int)cg::this_cluster().num_blocks(); }
__device__ __forceinline__ void csync() { cg::this_cluster().sync(); }
template <class T>
__device__ __forceinline__ T *peer(T *p, const int rank) { return cg::this_cluster().map_shared_rank(p, rank); }
#else
__device__ __forceinline__ int crank() { return 0; }
__device__ __forceinline__ int csize() { return 1; }
#endif

struct Smem {
    float4 *pos;                 // x, y, z, (type | frozen << 8) as int bits
    float *vx, *vy, *vz;
    float *fx, *fy, *fz;         // pair forces; the same bytes serve as scratch during a list build
    float *gx, *gy, *gz;         // [n_topo_pad] bonded forces of the beads that have bonded terms (own arrays: the bonded pass and
                                 // the pair loop run without a barrier between them); last in the layout: the only run-time offsets
    float *bx, *by, *bz;         // positions at the last neighbour build
    PairSetup *setup;
    DevTab *tab;
    int *misc;                   // see M_* below
    double *dmisc;
    // aliases into the force region, valid only inside build_lists()
    unsigned short *cend;        // [-1 .. nslots + 2]  running end offset of every slot (cend[-1] = 0)
    unsigned short *sidx;        // [n_pad]       atom indices in candidate order: sorted by (slot, atom index)
    int *cnt;                    // [n_pad]       list length of every bead (first the scratch of the sort)
    unsigned int *wtab;          // [ESB_THREADS] list build: the non-empty spans of a warp's round (own region, 32 words per warp)
    unsigned char *owner;        // [n_pad]       cluster variant: rank of the CTA that owns the bead (its force quad)
    float4 *pos_next;            // cluster variant: the other position buffer (written by the per-atom phase of all CTAs)
    unsigned short *mine;        // cluster variant: the beads this CTA owns (any order), misc[M_NMINE] of them
};

enum { M_QCTR = 0 /* next quad of the running build / force loop */, M_FAIL, M_PAIRS, M_CURLIST, M_FMAX /* largest |pair force| of the chunk, 2^-8 units */, M_SCAN /* one per warp */,
       M_LEVEL = M_SCAN + ESB_THREADS / 32 /* 64: work items of the force loop up to and including every chunk level (see build_lists) */,
       M_CNT_PROM = M_LEVEL + 64, M_CNT_GENE, M_NRNP, M_NRBP, M_HIST /* 49 */,
       M_CBOX = M_HIST + 52 /* list build: bounding box (6 floats) of every species class of at most 32 beads */,
       M_CROW = M_CBOX + 6 * ESB_NCLASS /* list build: occupied rows of every class, first y | last y << 8 | first z << 16 | last z << 24 */,
#if ESB_CLUSTER
       M_PARITY = M_CROW + ESB_NCLASS /* which of the two position buffers is current */, M_NMINE /* beads this CTA owns */, M_VOTE /* rebuild vote [parity][rank] */,
       M_CFAIL = M_VOTE + 16 /* failure bits [parity][rank] */, M_END = M_CFAIL + 16 };
#else
       M_END = M_CROW + ESB_NCLASS };
#endif
enum { D_RR = 0, D_PROBE = 3, D_GENE = 6,
       // per-chunk statistics, kept by thread 0 in shared memory (as 64-bit integers) instead of in registers of
       // every thread: the chunk loop runs at the 64-register cap and would spill them to L2-backed local memory
       L_EVALS = 10, L_BUILDS, L_STEPS, L_LISTED, L_DANGER, L_TATOM, L_TBUILD, L_TFORCE, L_TOBS, L_TMARK, D_END };

__host__ __device__ inline size_t smem_bytes_for(int n_pad, int n_tables, int n_topo_pad)
{
    size_t b = 0;
    b += (size_t)n_pad * 16;                 // pos
    b += (size_t)n_pad * 4 * 9;              // v, f (pair), xbuild
    b += (size_t)n_topo_pad * 4 * 3 + 16;    // f (bonded), behind everything else (16-byte aligned)
    b += (sizeof(PairSetup) + 15) & ~(size_t)15;
    b += ((size_t)n_tables * sizeof(DevTab) + 15) & ~(size_t)15;
    b += ((size_t)M_END * 4 + 15) & ~(size_t)15;
    b += (size_t)D_END * 8;
    b += (size_t)ESB_THREADS * 4;            // span table of the list build
#if ESB_CLUSTER
    b += ((size_t)n_pad + 15) & ~(size_t)15; // bead owners
    b += (size_t)n_pad * 16;                 // second position buffer (the CTAs of a cluster write the next positions there)
    b += (size_t)n_pad * 2;                  // the beads this CTA owns
#endif
    return (b + 15) & ~(size_t)15;
}

extern __shared__ __align__(16) unsigned char smem_raw[];

template <int NP>
__device__ __forceinline__ Smem carve(const KArgs &a, const bool first_buffer = false)
{
    const int np = NP ? NP : a.n_pad;   // compile-time bead capacity: every array offset is an immediate
    Smem S;
    unsigned char *base = smem_raw;
    size_t o = 0;
    S.pos = reinterpret_cast<float4 *>(base + o); o += (size_t)np * 16;
    S.vx = reinterpret_cast<float *>(base + o); o += (size_t)np * 4;
    S.vy = reinterpret_cast<float *>(base + o); o += (size_t)np * 4;
    S.vz = reinterpret_cast<float *>(base + o); o += (size_t)np * 4;
    S.fx = reinterpret_cast<float *>(base + o); o += (size_t)np * 4;
    S.fy = reinterpret_cast<float *>(base + o); o += (size_t)np * 4;
    S.fz = reinterpret_cast<float *>(base + o); o += (size_t)np * 4;
    S.bx = reinterpret_cast<float *>(base + o); o += (size_t)np * 4;
    S.by = reinterpret_cast<float *>(base + o); o += (size_t)np * 4;
    S.bz = reinterpret_cast<float *>(base + o); o += (size_t)np * 4;
    S.setup = reinterpret_cast<PairSetup *>(base + o); o += (sizeof(PairSetup) + 15) & ~(size_t)15;
    S.tab = reinterpret_cast<DevTab *>(base + o); o += ((size_t)a.n_tables * sizeof(DevTab) + 15) & ~(size_t)15;
    S.dmisc = reinterpret_cast<double *>(base + o); o += (size_t)D_END * 8;
    S.misc = reinterpret_cast<int *>(base + o); o += ((size_t)M_END * 4 + 15) & ~(size_t)15;
    S.wtab = reinterpret_cast<unsigned int *>(base + o); o += (size_t)ESB_THREADS * 4;
    S.owner = reinterpret_cast<unsigned char *>(base + o);
#if ESB_CLUSTER
    o += ((size_t)np + 15) & ~(size_t)15;
    {   // double-buffered positions: the current buffer is selected by a flag that only changes between phases
        float4 *alt = reinterpret_cast<float4 *>(base + o);
        S.pos_next = alt;
        if (!first_buffer && S.misc[M_PARITY]) { S.pos_next = S.pos; S.pos = alt; }
        S.mine = reinterpret_cast<unsigned short *>(base + o + (size_t)np * 16);
    }
#endif
#if ESB_CLUSTER
    o += (size_t)np * 16 + (size_t)np * 2;
#endif
    o = (o + 15) & ~(size_t)15;
    S.gx = reinterpret_cast<float *>(base + o);
    S.gy = S.gx + a.n_topo_pad;
    S.gz = S.gy + a.n_topo_pad;
    S.cend = reinterpret_cast<unsigned short *>(S.fx) + 2;
    S.sidx = S.cend + 6 + ((a.nrows * ESB_NCLASS * a.ncx + 2 + 7) & ~7);
    S.cnt = reinterpret_cast<int *>(S.sidx + np);
    return S;
}

// MUFU.RCP without the denormal wrapper that __fdividef(1, x) drags in (x is O(0.3 .. 1e3) here)
__device__ __forceinline__ float rcp_approx(const float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Shared memory viewed as float4 words: indexing this array (instead of dereferencing pointers
// carried in a struct) lets the compiler emit LDS [reg + immediate] in the hot loops.
extern __shared__ __align__(16) float4 smem4[];
// The hot loops address shared memory through a 32-bit base held in a register and ld.shared: the compiler's
// own addressing of smem4[] re-derives the CTA's shared window (S2UR CgaCtaId, ULEA ...) at every use.
__device__ __forceinline__ unsigned int smem_base()
{
    unsigned int b = (unsigned int)__cvta_generic_to_shared(smem_raw);
    asm volatile("mov.u32 %0, %0;" : "+r"(b));        // opaque: kept in a register, not rematerialised
    return b;
}
__device__ __forceinline__ float4 lds128(const unsigned int addr)
{
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// Pair forces are accumulated as 32-bit fixed-point integers (ESB_FSCALE counts per force unit): integer addition is
// associative, so a bead's force does not depend on the order in which its pairs are visited, and shared-memory integer
// atomics are native (ATOMS.ADD, ~3 cycles per warp instruction at random addresses -- profiles/r2_microbench_smem.log --
// whereas fp32 shared atomics are CAS loops on sm_100).  That is what makes Newton's third law affordable here: the
// HALF list (every unordered pair stored once) evaluates a pair once, keeps bead i's share in registers and scatters
// -f to bead j.  Scale 2^16: resolution 1.5e-5 (a bead's ~60 roundings add up to ~3e-5 rms on forces of order 10..100,
// 1e-6 of the Langevin noise amplitude), range +-32768.  Measured headroom (scripts/force_headroom.py, 296 box-11 runs x
// 400 chunks, profiles/r2_force_headroom.log): the largest pair-force component of a chunk is typically 150..350 and
// reached 1923 once (a bead whose type -- hence core size -- had just changed inside the Ser5P cluster); one LJsoft
// pair contributes at most ~365.  A bead whose sum leaves half the range (16384) flags the replica (ESB_ST_NONFINITE)
// in the per-atom phase; counter 11 reports the largest sum seen.
#define ESB_FSCALE 0x1p16f
// Bonded shares have their own accumulators (S.g*) at a coarser scale, 2^12 counts per force unit: resolution 2.4e-4 (five
// terms per bead: <= 6e-4, 1e-5 of the Langevin noise amplitude), range +-524288.  The actin filaments need the range: fix
// s3 re-ramps their bond r0 from 0.033 in every run (run_single_shoebox.py:638-639), bonds pass through lengths of ~1e-3
// and the cosine-angle force K / r reaches 1e4..1e5 for a step -- survivable in the reference's doubles and in fp32, but
// beyond the +-16384 at which a 2^16 accumulator is flagged (with the bonded shares in the pair accumulators 59 of 64
// configs[1] repeats were flagged by chunk 400 instead of the 20 that die of a wall crossing).
#define ESB_BSCALE 0x1p12f
// Work items of the half-list force loop: a quad (32 / ESB_G beads, consecutive in the order of decreasing list length) times a
// CHUNK of its lists, ESB_ITEM_ROUNDS entries per lane.  A bead whose list is longer than a chunk is finished by other
// warps: every item adds its part of bead i's force with the same integer atomics that reach bead j, so the items are
// independent and the phase no longer waits for the warp that drew the longest lists (those of the Ser5P cluster are
// ten times the median).  Level c = the c-th chunk of every quad that has one; the lengths are sorted, so level c is the
// first Q_c quads and item t is found from the running sums of Q_c.
#ifndef ESB_ITEM_ROUNDS
#define ESB_ITEM_ROUNDS 32
#endif
#define ESB_CHUNK (ESB_ITEM_ROUNDS * ESB_G)      // entries of one bead per item
#ifndef ESB_CUT32
#define ESB_CUT32 1
#endif
#ifndef ESB_ITEM_ORDER
#define ESB_ITEM_ORDER 1
#endif
#ifndef ESB_ITEM_FAST
#define ESB_ITEM_FAST 1
#endif
#ifndef ESB_BUILD_NB
#define ESB_BUILD_NB 5
#endif
#ifndef ESB_FGUARD
#define ESB_FGUARD 1
#endif
#ifndef ESB_HALF
#define ESB_HALF (!ESB_CLUSTER)      // the cluster variant deals the quads to several CTAs and keeps the full list (no remote scatter)
#endif
static_assert(!ESB_HALF || 64 * ESB_CHUNK > ESB_NL_MAXNB, "64 chunk levels must cover the longest list");
template <int NP>
__device__ __forceinline__ unsigned int force_off(const KArgs &a) { return 28u * (unsigned int)(NP ? NP : a.n_pad); }   // byte offset of S.fx
// f[j] += (vx, vy, vz) for the three SoA force arrays; addr = shared address of fx[j]
template <int NP>
__device__ __forceinline__ void red_add3(const KArgs &a, const unsigned int addr, const int vx, const int vy, const int vz)
{
    if (NP) {
        asm volatile("red.shared.add.s32 [%0], %1;\n\tred.shared.add.s32 [%0+%4], %2;\n\tred.shared.add.s32 [%0+%5], %3;"
                     :: "r"(addr), "r"(vx), "r"(vy), "r"(vz), "n"(4 * NP), "n"(8 * NP) : "memory");
    } else {
        const unsigned int st = 4u * (unsigned int)a.n_pad;
        asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr), "r"(vx) : "memory");
        asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr + st), "r"(vy) : "memory");
        asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr + 2u * st), "r"(vz) : "memory");
    }
}
// the same for arrays of a run-time stride (the bonded accumulators, [n_topo_pad] each)
[[maybe_unused]] __device__ __forceinline__ void red_add3s(const unsigned int addr, const unsigned int stride_bytes, const int vx, const int vy, const int vz)
{
    asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr), "r"(vx) : "memory");
    asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr + stride_bytes), "r"(vy) : "memory");
    asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr + 2u * stride_bytes), "r"(vz) : "memory");
}
template <int NP>
__device__ __forceinline__ int tab_off4(const KArgs &a)
{
    const int np = NP ? NP : a.n_pad;
    return (np * 52 + (int)((sizeof(PairSetup) + 15) & ~(size_t)15)) / 16;   // float4 index of S.tab
}

__device__ __forceinline__ int cell_coord(float x, float lo, float inv, int n)
{
    int c = (int)((x - lo) * inv);
    return min(max(c, 0), n - 1);
}

// slot of a bead in the candidate sort: (species class, z row, y row, x cell), class-major and x-minor: a row is one
// contiguous range of the sorted order, and so is every run of consecutive x cells of a row
__device__ __forceinline__ int slot_of(const KArgs &a, const float4 p)
{
    const int ry = cell_coord(p.y, a.lo[1], a.row_inv, a.nry);
    const int rz = cell_coord(p.z, a.lo[2], a.row_inv, a.nrz);
    return (a.class_of[__float_as_int(p.w) & 0xff] * a.nrows + rz * a.nry + ry) * a.ncx + cell_coord(p.x, a.lo[0], a.xc_inv, a.ncx);
}

// Counting sort of the beads into `nslots` slots (CTA-wide; called by all threads).  Afterwards S.cend[s] is
// the END offset of slot s (start = S.cend[s-1]; S.cend[-1] is a zero that belongs to the array) and `tmp` holds the
// bead indices grouped by slot, in arbitrary order inside a slot.  The counters are u16 packed in pairs for the
// shared-memory atomics.
template <class SlotFn>
__device__ __forceinline__ void slot_sort(const Smem &S, const KArgs &a, const int nslots, unsigned short *tmp, SlotFn slot_fn)
{
    const int tid = threadIdx.x, n = a.n_atoms;
    const int lane = tid & 31, warp = tid >> 5;
    unsigned int *cw = reinterpret_cast<unsigned int *>(S.cend);
    const int ncw = (nslots + 2 + 1) >> 1;
    for (int w = tid; w < ncw; w += ESB_THREADS) cw[w] = 0u;
    if (tid == 0) cw[-1] = 0u;                                       // S.cend[-1] (and the unused half word before it)
    __syncthreads();
    // count (slot s is stored at position s+1 so that, after the scan, position s holds start(s))
    for (int i = tid; i < n; i += ESB_THREADS) {
        const int slot = slot_fn(S.pos[i]) + 1;
        atomicAdd(&cw[slot >> 1], 1u << (16 * (slot & 1)));
    }
    __syncthreads();
    // inclusive scan over positions 0..nslots (position 0 is 0): afterwards cend[s] = start of slot s
    {
        const int total = nslots + 1;
        const int per = (total + ESB_THREADS - 1) / ESB_THREADS;
        const int b0 = min(tid * per, total), b1 = min(b0 + per, total);
        int sum = 0;
        for (int c = b0; c < b1; c++) sum += S.cend[c];
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) S.misc[M_SCAN + warp] = incl;
        __syncthreads();
        int base = incl - sum;
        for (int w = 0; w < warp; w++) base += S.misc[M_SCAN + w];
        for (int c = b0; c < b1; c++) { base += S.cend[c]; S.cend[c] = (unsigned short)base; }
    }
    __syncthreads();
    // scatter: cend[s] (= start of s) is the cursor of slot s; afterwards cend[s] = end of s
    for (int i = tid; i < n; i += ESB_THREADS) {
        const int slot = slot_fn(S.pos[i]);
        const unsigned int old = atomicAdd(&cw[slot >> 1], 1u << (16 * (slot & 1)));
        tmp[(old >> (16 * (slot & 1))) & 0xffffu] = (unsigned short)i;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Neighbour list build.  Stands in for LAMMPS' binned neighbour build (`neighbor 0.3 multi`,
// `neigh_modify every 1 delay 10 check yes`, run_single_shoebox.py:532-533): per-type-pair
// cutoff + skin, 1-2/1-3/1-4 partners excluded (default special_bonds); every entry carries the pair's table id.
//
// Beads are counting-sorted into slots = (species class, z row, y row, x cell) -- rows of edge 2.8 over (y, z), cut
// into cells of edge ~1 along x -- by atom index inside a slot ("candidate order").  The candidates of ONE bead towards
// class B are then, for every row within reach(type, B) in y and z, one contiguous span of that order: the cells of the
// row within sqrt(reach^2 - d^2) in x, d being the distance from the bead to the row's slab -- two table look-ups per
// span, no search, and the cutoff of that class pair (0.6 for RBP-RBP ... 3.0 for RNP-RNP) instead of the bead's
// largest.  A warp takes four consecutive beads of the candidate order at a time: lanes 0..23 work out the row ranges of
// the 4 x 6 (bead, class) combinations, then the lanes resolve the spans of all four beads (one span per lane and round),
// prefix-sum their lengths and run through the concatenation 32 candidates at a time, ONE (bead, candidate) pair per
// lane.  An accepted pair takes its place in the bead's arena from a shared-memory counter (atomic: the order of a list
// is immaterial, the force sums are integers).  Half list: bead i keeps candidate j only when j comes after i in
// candidate order, so classes before the bead's own and rows before its own are never scanned and every span is
// clipped to the positions behind the bead.  (Round 2 scanned the bounding box of a quad of beads over whole rows with
// binary searches in x: 8 times the distance tests.)  Afterwards the descriptors are sorted by list length: that is
// the order in which the force loop forms its quads.
// ------------------------------------------------------------------------------------------------
#if defined(ESB_BUILD_PROF) && ESB_BUILD_PROF      // experiments: the build's sort goes to the per-atom timer, thread 0's batches to the observe timer
#define ESB_BP_TICK(acc) { if (tid == 0) { long long *L_ = reinterpret_cast<long long *>(S.dmisc); const long long now_ = clock64(); L_[acc] += now_ - L_[L_TMARK]; L_[L_TMARK] = now_; } }
#else
#define ESB_BP_TICK(acc)
#endif
template <int NP, bool P16>
__device__ ESB_PHASE_B void build_lists(const KArgs &a, const int r)
{
    const Smem S = carve<NP>(a);
    const int tid = threadIdx.x, n = a.n_atoms;
    const int lane = tid & 31, warp = tid >> 5;
    const int nrows = a.nrows, nry = a.nry, nrz = a.nrz, ncx = a.ncx;
    const int nslots = nrows * ESB_NCLASS * ncx;
    if (tid == 0) { S.misc[M_CURLIST] = 0; S.misc[M_QCTR] = 0; }
    // ---- candidate order: beads grouped by (class, row, x cell), by atom index inside a cell ----
    {
        unsigned short *tmp = reinterpret_cast<unsigned short *>(S.cnt);
        auto cell_of = [&](const float4 p) { return slot_of(a, p); };
        slot_sort(S, a, nslots, tmp, cell_of);
        for (int i = tid; i < n; i += ESB_THREADS) {
            const float4 p = S.pos[i];
            S.bx[i] = p.x; S.by[i] = p.y; S.bz[i] = p.z;
            const int slot = slot_of(a, p);
            const int s0 = S.cend[slot - 1], s1 = S.cend[slot];
            int rank = 0;
            for (int q = s0; q < s1; q++) rank += (tmp[q] < (unsigned short)i);
            S.sidx[s0 + rank] = (unsigned short)i;
        }
        __syncthreads();
        for (int i = tid; i < n; i += ESB_THREADS) S.cnt[i] = 0;     // (the sort scratch becomes the list lengths)
        // a class of at most 32 beads (the gene body) is ONE span for everybody, offered only within reach of its bounding box
        // and nobody looks for a class outside the rows it occupies (the Ser5P cluster sits in a few of them)
        for (int B = warp; B < ESB_NCLASS; B += ESB_THREADS / 32) {
            const int s0 = S.cend[B * nrows * ncx - 1], pop = (int)S.cend[(B + 1) * nrows * ncx - 1] - s0;
            {
                int ya = 255, yb = 0, za = 255, zb = 0;
                for (int rw = lane; rw < nrows; rw += 32) {
                    if ((int)S.cend[(B * nrows + rw + 1) * ncx - 1] > (int)S.cend[(B * nrows + rw) * ncx - 1]) {
                        const int rz = rw / nry, ry = rw - rz * nry;
                        ya = min(ya, ry); yb = max(yb, ry); za = min(za, rz); zb = max(zb, rz);
                    }
                }
                ya = __reduce_min_sync(0xffffffffu, ya); yb = __reduce_max_sync(0xffffffffu, yb);
                za = __reduce_min_sync(0xffffffffu, za); zb = __reduce_max_sync(0xffffffffu, zb);
                if (lane == 0) S.misc[M_CROW + B] = ya | (yb << 8) | (za << 16) | (zb << 24);
            }
            if (pop > 0 && pop <= 32) {
                const float4 p = S.pos[S.sidx[s0 + min(lane, pop - 1)]];
                float x0 = p.x, x1 = p.x, y0 = p.y, y1 = p.y, z0 = p.z, z1 = p.z;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, d)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, d));
                    y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, d)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, d));
                    z0 = fminf(z0, __shfl_xor_sync(0xffffffffu, z0, d)); z1 = fmaxf(z1, __shfl_xor_sync(0xffffffffu, z1, d));
                }
                if (lane == 0) {
                    float *bb = reinterpret_cast<float *>(S.misc + M_CBOX + 6 * B);
                    bb[0] = x0; bb[1] = x1; bb[2] = y0; bb[3] = y1; bb[4] = z0; bb[5] = z1;
                }
            }
        }
        __syncthreads();
    }
    ESB_BP_TICK(L_TATOM)

    const unsigned int FULL = 0xffffffffu;
    const unsigned int lt_mask = (1u << lane) - 1u, le_mask = 0xffffffffu >> (31 - lane);
    // P16 (at most 32 tables): 16-bit entries (neighbour | table id << 11), two per word; half the list traffic
    constexpr unsigned int ESTRIDE = P16 ? ESB_NL_STRIDE / 2 : ESB_NL_STRIDE;      // arena of a bead in words
    unsigned int *pool = a.nl_pool + (size_t)r * a.n_pad * ESTRIDE;
    unsigned short *pool16 = reinterpret_cast<unsigned short *>(pool);
    uint2 *qbuild = a.nl_qinfo + ((size_t)r * 2 + 1) * a.n_pad;     // descriptors in build (candidate) order
    const float lox = a.lo[0], loy = a.lo[1], loz = a.lo[2], rinv = a.row_inv, redge = a.row_edge, xinv = a.xc_inv;
    const int n_topo = a.n_topo;
    const float *cs_row = S.setup->cutsq_skin;
    const unsigned char *tab_row = S.setup->tabid;
    unsigned int listed = 0;
    constexpr int NB = ESB_BUILD_NB;                                 // beads per batch: NB * ESB_NCLASS (bead, class) combinations <= 32 lanes
    static_assert(NB * ESB_NCLASS <= 32, "one lane per (bead, class) combination");
    const int nbatch = (n + NB - 1) / NB;
    // this lane's (bead, class) combination of a batch
    const int cb_bead = lane / ESB_NCLASS, cb_class = lane - cb_bead * ESB_NCLASS;
    int cb_pop = 0, cb_rows = 0;                                     // beads of class cb_class, the rows it occupies
    if (lane < NB * ESB_NCLASS) {
        cb_pop = (int)S.cend[(cb_class + 1) * nrows * ncx - 1] - (int)S.cend[cb_class * nrows * ncx - 1];
        cb_rows = S.misc[M_CROW + cb_class];
    }
    for (;;) {
        // batches are handed out dynamically: their cost varies by two orders of magnitude with the species
        int q = 0;
        if (lane == 0) q = atomicAdd(&S.misc[M_QCTR], 1);
        q = __shfl_sync(FULL, q, 0) * csize() + crank();         // (cluster variant: the batches are dealt to the CTAs of the cluster)
        if (q >= nbatch) break;
        if (ESB_BUILD_REVERSE) q = nbatch - 1 - q;               // last classes of the candidate order first (gene body, Ser5P, RNP: the long scans)
        // ---- the row ranges of the (bead, class) combinations; my_* are valid in lanes 0 .. NB * ESB_NCLASS - 1 ----
        int my_pack = 0, my_end = 0, my_cum = 0, my_bead = 0;     // first y row | rows in y << 8 | first z row << 16;  atom | candidate position << 11 | type << 22 | class << 26
        float my_reach = 0.0f;
        {
            int nsp = 0;
            const int sp_i = NB * q + cb_bead;                   // position of the bead in the candidate order
            if (lane < NB * ESB_NCLASS && sp_i < n) {
                const int i = S.sidx[sp_i];
                const float4 pi = S.pos[i];
                const int ti = __float_as_int(pi.w) & 0xff, cls = a.class_of[ti];
                const float c = S.setup->reach[ti * 8 + cb_class];
                my_bead = i | (sp_i << 11) | (ti << 22) | (cls << 26);
                if (c > 0.0f && cb_pop > 0 && cb_pop <= 32 && (!ESB_HALF || cb_class >= cls)) {
                    // small class: the whole class as one span (rows in y = 0 marks it), when the bead is within reach of its box
                    const float *bb = reinterpret_cast<const float *>(S.misc + M_CBOX + 6 * cb_class);
                    const float ex = fmaxf(fmaxf(bb[0] - pi.x, pi.x - bb[1]), 0.0f), ey = fmaxf(fmaxf(bb[2] - pi.y, pi.y - bb[3]), 0.0f),
                                ez = fmaxf(fmaxf(bb[4] - pi.z, pi.z - bb[5]), 0.0f);
                    if (ex * ex + ey * ey + ez * ez < c * c) { nsp = 1; my_reach = c; }
                } else if (c > 0.0f && cb_pop > 0 && (!ESB_HALF || cb_class >= cls)) {
                    const int y0 = max(cell_coord(pi.y - c, loy, rinv, nry), cb_rows & 0xff), y1 = min(cell_coord(pi.y + c, loy, rinv, nry), (cb_rows >> 8) & 0xff);
                    int z0 = max(cell_coord(pi.z - c, loz, rinv, nrz), (cb_rows >> 16) & 0xff);
                    const int z1 = min(cell_coord(pi.z + c, loz, rinv, nrz), (cb_rows >> 24) & 0xff);
                    // HALF list: in the bead's own class nothing before its own z row comes after it in candidate order
                    if (ESB_HALF && cb_class == cls) z0 = max(z0, cell_coord(pi.z, loz, rinv, nrz));
                    if (y1 >= y0 && z1 >= z0) {
                        my_pack = y0 | ((y1 - y0 + 1) << 8) | (z0 << 16);
                        my_reach = c;
                        nsp = (y1 - y0 + 1) * (z1 - z0 + 1);
                    }
                }
            }
            my_end = nsp;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(FULL, my_end, d);
                if (lane >= d) my_end += t;
            }
            my_cum = my_end - nsp;
        }
        const int n_spans = __shfl_sync(FULL, my_end, 31);
        for (int sp0 = 0; sp0 < n_spans; sp0 += 32) {
            // ---- this lane's span: (bead, class, row) -> the cells of the row within reach in x ----
            const int sp = sp0 + lane;
            int c = 0;                                           // combination of span sp: the first lane with my_end > sp
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) {
                const int t = __shfl_sync(FULL, my_end, c + st - 1);
                if (t <= sp) c += st;
            }
            const int pack = __shfl_sync(FULL, my_pack, c & 31), cum = __shfl_sync(FULL, my_cum, c & 31);
            const int bead = __shfl_sync(FULL, my_bead, c & 31);
            const float reach = __shfl_sync(FULL, my_reach, c & 31);
            int ka = 0, kb = 0;
            if (sp < n_spans) {
                const int B = c - (c / ESB_NCLASS) * ESB_NCLASS;
                const int li = sp - cum, ny = (pack >> 8) & 0xff;
                if (ny == 0) {                                   // small class: all of it
                    ka = S.cend[B * nrows * ncx - 1];
                    kb = S.cend[(B + 1) * nrows * ncx - 1];
                } else {
                    const int dzr = (int)(((float)li + 0.5f) * rcp_approx((float)ny));
                    const int cy = (pack & 0xff) + li - dzr * ny, cz = (pack >> 16) + dzr;
                    const float4 pi = S.pos[bead & 0x7ff];
                    // distance from the bead to the slab of the row (a little short: the row of a bead comes from a rounded product)
                    const float ylo = fmaf((float)cy, redge, loy), zlo = fmaf((float)cz, redge, loz);
                    const float dy = fmaxf(fmaxf(ylo - pi.y, pi.y - (ylo + redge)) - 1.0e-5f, 0.0f);
                    const float dz = fmaxf(fmaxf(zlo - pi.z, pi.z - (zlo + redge)) - 1.0e-5f, 0.0f);
                    const float rr = fmaf(reach, reach, -fmaf(dy, dy, dz * dz));
                    if (rr > 0.0f) {
                        const float rx = rr * rsqrtf(rr) * 1.00001f;            // (2 ulp of MUFU.RSQ are inside the factor)
                        const int x0 = cell_coord(pi.x - rx, lox, xinv, ncx), x1 = cell_coord(pi.x + rx, lox, xinv, ncx);
                        const int base = (B * nrows + cz * nry + cy) * ncx;
                        ka = S.cend[base + x0 - 1];
                        kb = S.cend[base + x1];
#if defined(ESB_BUILD_CHECK) && ESB_BUILD_CHECK          // self-check build (compute-sanitizer is not available on the pool): index ranges of the span
                        if (B < 0 || B >= ESB_NCLASS || cy < 0 || cy >= nry || cz < 0 || cz >= nrz || x0 < 0 || x1 >= ncx || x0 > x1 || base + x1 >= nslots || ka > kb || kb > n)
                            atomicOr(&S.misc[M_FAIL], ESB_ST_NEIGH);
#endif
                    }
                }
                // nothing at or before the bead itself can be listed (its own class only: later classes lie behind it anyway)
                if (ESB_HALF) ka = min(max(ka, ((bead >> 11) & 0x7ff) + 1), kb);
            }
            const int len = kb - ka;
            int incl = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            const int total = __shfl_sync(FULL, incl, 31);
            const int start = incl - len;                        // first candidate of this span in the concatenation
            // the non-empty spans, packed, go to the warp's table: (sorted position of candidate f) - f + 65536 | atom << 17 | type << 28
            // (a round has at most 32 n <= 65536 candidates)
            unsigned int *wt = S.wtab + 32 * warp;
            {
                const unsigned int ne = __ballot_sync(FULL, len > 0);
                __syncwarp();                                    // (the previous round's readers are done)
                if (len > 0) wt[__popc(ne & lt_mask)] = (unsigned int)(ka - start + 65536) | ((unsigned int)(bead & 0x7ff) << 17) | ((unsigned int)((bead >> 22) & 0xf) << 28);
                __syncwarp();
            }
            // ---- the concatenated spans, 32 (bead, candidate) pairs per iteration: candidate f belongs to the last span that
            //      starts at or before f = (spans started before the window) + (starts inside the window up to f) ----
            int before = 0;
            for (int f0 = 0; f0 < total; f0 += 32) {
                const int f = f0 + lane;
                const bool live = f < total;
                const unsigned int rel = (unsigned int)(start - f0);
                const unsigned int starts = __reduce_or_sync(FULL, (len > 0 && rel < 32u) ? 1u << rel : 0u);
                const unsigned int word = wt[live ? before + __popc(starts & le_mask) - 1 : 0];
                before += __popc(starts);
                const int k = f + (int)(word & 0x1ffffu) - 65536;
                const int bd = (int)(word >> 17);                // atom | type << 11
#if defined(ESB_BUILD_CHECK) && ESB_BUILD_CHECK
                if (live && (k < 0 || k >= n || (bd & 0x7ff) >= n || (ESB_HALF && k <= 0))) atomicOr(&S.misc[M_FAIL], ESB_ST_NEIGH);
#endif
                const int j = live ? S.sidx[k] : 0;
                const float4 pj = S.pos[j];
                const int i = bd & 0x7ff, ti = bd >> 11;
                const float4 pi = S.pos[i];
                const int tj = __float_as_int(pj.w) & 0xff;
                const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                const float r2 = dx * dx + dy * dy + dz * dz;
                bool ok = live && (r2 < cs_row[ti * ESB_MAXT + tj]) && (ESB_HALF || j != i);
                if (ok && j < n_topo && i < n_topo) {            // only then can j be an excluded 1-2/1-3/1-4 partner
                    const uint4 sp4 = __ldcg(reinterpret_cast<const uint4 *>(a.spec + (size_t)r * a.spec_stride + (size_t)i * ESB_MAX_SPEC));
                    const unsigned int jj = (unsigned int)j;
                    ok = !((sp4.x & 0xffffu) == jj || (sp4.x >> 16) == jj || (sp4.y & 0xffffu) == jj || (sp4.y >> 16) == jj ||
                           (sp4.z & 0xffffu) == jj || (sp4.z >> 16) == jj || (sp4.w & 0xffffu) == jj || (sp4.w >> 16) == jj);
                }
                if (ok) {
                    const int at = atomicAdd(&S.cnt[i], 1);
                    if (at < ESB_NL_MAXNB) {
                        const unsigned int tb = tab_row[ti * ESB_MAXT + tj];
                        if (P16) pool16[(unsigned int)i * ESB_NL_STRIDE + at] = (unsigned short)((unsigned int)j | (tb << 11));
                        else pool[(unsigned int)i * ESB_NL_STRIDE + at] = (unsigned int)j | (tb << 16);
                    }
                }
            }
        }
        __syncwarp();
        if (lane < NB) {
            const int idx = NB * q + lane;
            if (idx < n) {
                const int i = S.sidx[idx];
                int c = S.cnt[i];
                if (c > ESB_NL_MAXNB) { atomicOr(&S.misc[M_FAIL], ESB_ST_NEIGH); c = ESB_NL_MAXNB; }
                qbuild[idx] = make_uint2((unsigned int)i * ESTRIDE, (unsigned int)i | ((unsigned int)c << 16));
                listed += ESB_HALF ? 2 * c : c;                  // the counters report ordered pairs
            }
        }
    }
    if (listed) atomicAdd(&S.misc[M_CURLIST], (int)listed);
    ESB_BP_TICK(L_TOBS)
    __syncthreads();
#if ESB_CLUSTER
    __threadfence();
    csync();                                                     // the descriptors of all batches (global) are written
#endif
    // ---- force order: descriptors sorted by decreasing list length (counting sort on the length) ----
    if (crank() == 0) {
        uint2 *qforce = a.nl_qinfo + (size_t)r * 2 * a.n_pad;
        // one counter per possible length 0 .. hmax: a list cannot be longer than the system has beads (nor than an arena),
        // so the counters fit the scratch -- free again -- of a system of any size
        const int hmax = min(ESB_NL_MAXNB, n);
        int *hist = reinterpret_cast<int *>(S.fx);
        for (int k = tid; k < hmax + 1; k += ESB_THREADS) hist[k] = 0;
        __syncthreads();
        for (int k = tid; k < n; k += ESB_THREADS) atomicAdd(&hist[hmax - (int)(__ldcg(&qbuild[k]).y >> 16)], 1);
        __syncthreads();
        // exclusive scan of the counters: an equal share per thread
        constexpr int PER = (ESB_NL_MAXNB + ESB_THREADS) / ESB_THREADS;
        int v[PER], s = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) { v[k] = (PER * tid + k <= hmax) ? hist[PER * tid + k] : 0; s += v[k]; }
        int incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) S.misc[M_SCAN + warp] = incl;
        __syncthreads();
        int base = incl - s;
        for (int w = 0; w < warp; w++) base += S.misc[M_SCAN + w];
#pragma unroll
        for (int k = 0; k < PER; k++) { if (PER * tid + k <= hmax) hist[PER * tid + k] = base; base += v[k]; }
        __syncthreads();
#if ESB_HALF
        // hist[hmax - L] = beads with a list longer than L = the beads of chunk level L / ESB_CHUNK
        if (tid < 64) {
            const int L = tid * ESB_CHUNK;
            const int beads = L <= hmax ? hist[hmax - L] : 0;
            S.misc[M_LEVEL + tid] = (beads + 32 / ESB_G - 1) / (32 / ESB_G);       // quads of this level
        }
        __syncthreads();
        if (warp == 0) {
            int v0 = S.misc[M_LEVEL + lane], v1 = S.misc[M_LEVEL + 32 + lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t0 = __shfl_up_sync(FULL, v0, d), t1 = __shfl_up_sync(FULL, v1, d);
                if (lane >= d) { v0 += t0; v1 += t1; }
            }
            v1 += __shfl_sync(FULL, v0, 31);
            S.misc[M_LEVEL + lane] = v0; S.misc[M_LEVEL + 32 + lane] = v1;
            a.nl_levels[(size_t)r * 64 + lane] = v0; a.nl_levels[(size_t)r * 64 + 32 + lane] = v1;       // (for a later segment of the chunk)
        }
#endif
        for (int k = tid; k < n; k += ESB_THREADS) {
            const uint2 d = __ldcg(&qbuild[k]);
            const int at = atomicAdd(&hist[hmax - (int)(d.y >> 16)], 1);
            qforce[at] = d;
#if ESB_CLUSTER
            // the CTA that takes force quad at/4 owns the bead: bonded terms, pair loop, integration
            for (int rk = 0; rk < csize(); rk++) peer(S.owner, rk)[d.y & 0xffffu] = (unsigned char)((at / (32 / ESB_G)) % csize());
#endif
        }
    }
#if ESB_CLUSTER
    if (tid == 0) S.misc[M_NMINE] = 0;
#endif
    if (tid == 0) S.misc[M_QCTR] = 0;                            // the force loop hands out quads from 0
    __syncthreads();   // lists (global) and the scratch alias are settled before forces are written
    {   // the scratch becomes the (integer) pair-force accumulators again: they are added to, so they start at zero
        int *fi = reinterpret_cast<int *>(S.fx);
        for (int k = tid; k < 3 * a.n_pad; k += ESB_THREADS) fi[k] = 0;
    }
    __syncthreads();
#if ESB_CLUSTER
    __threadfence();
    csync();                                                     // force order (global) and owners (every CTA) are in place
    for (int i0 = warp * 32; i0 < n; i0 += ESB_THREADS) {        // dense list of the owned beads: full warps in the per-bead loops
        const int i = i0 + lane;
        const bool own = i < n && S.owner[i] == crank();
        const unsigned int m = __ballot_sync(FULL, own);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&S.misc[M_NMINE], __popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (own) S.mine[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
    }
    __syncthreads();
#endif
}

// ------------------------------------------------------------------------------------------------
// Bonded forces.  Like the pair forces they are accumulated in fixed point (round(f * ESB_BSCALE) per component and
// term): every term is evaluated ONCE -- a bond by its end of lower index, an angle by its vertex -- and the shares of
// the other beads are scattered with the same shared-memory integer atomics, into accumulators of their own (S.g*; the
// first round-2 kernel evaluated every term from every bead it touches: 5 evaluations per polymer bead instead of 2).  The cluster variant
// still does that (a bead's force lives with its owner CTA) but with the same two functions below, whose results are
// exact mirror images between the two ends of a term: every variant adds up the same integers.
// A bead's row holds its bonds, then the angle it is the vertex of, then the angles it is an end of.
// [LAMMPS-ext] BondHarmonic::compute  E = K (r - r0)^2       (bond_coeff, run_single_shoebox.py:632-633)
// [LAMMPS-ext] AngleCosine::compute   E = K (1 + cos theta)  (angle_coeff, run_single_shoebox.py:583-584)
// ------------------------------------------------------------------------------------------------
struct I3 { int x, y, z; };
// r^2 with a fixed operation order: the soft-phase force depends on its last bit, and every kernel variant must agree
__device__ __forceinline__ float dist_sq(const float dx, const float dy, const float dz) { return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))); }
// share of the bead at p of its bond with the bead at q (the other end gets the exact negative: dx -> -dx, r^2 unchanged).
// (reciprocal square roots, MUFU.RSQ, ~2 ulp, instead of sqrt and divisions; the bonded forces stay within 1e-6 of the fp64 oracle's)
__device__ __forceinline__ I3 bond_share(const float4 p, const float4 q, const float r0, const float K, float &e)
{
    const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
    const float r2 = dist_sq(dx, dy, dz);
    const float ri = rsqrtf(r2);
    const float dr = __fsub_rn(__fmul_rn(r2, ri), r0);
    const float rk = __fmul_rn(K, dr);
    e = __fmul_rn(rk, dr);
    const float fs = (r2 > 0.0f) ? __fmul_rn(__fmul_rn(-2.0f * ESB_BSCALE, rk), ri) : 0.0f;
    I3 o; o.x = __float2int_rn(__fmul_rn(dx, fs)); o.y = __float2int_rn(__fmul_rn(dy, fs)); o.z = __float2int_rn(__fmul_rn(dz, fs));
    return o;
}
// share of the END bead at p1 of the angle p1 - pv - p3 (vertex pv); the other end's share is the same function with the
// ends exchanged (every intermediate that both calls need -- cos theta, 1/(r1 r2) -- is symmetric in the exchange bit for
// bit), the vertex gets minus the sum of the two
__device__ __forceinline__ I3 angle_end_share(const float4 p1, const float4 pv, const float4 p3, const float K, float &c_out)
{
    const float d1x = __fsub_rn(p1.x, pv.x), d1y = __fsub_rn(p1.y, pv.y), d1z = __fsub_rn(p1.z, pv.z);
    const float d2x = __fsub_rn(p3.x, pv.x), d2y = __fsub_rn(p3.y, pv.y), d2z = __fsub_rn(p3.z, pv.z);
    const float ri1 = rsqrtf(dist_sq(d1x, d1y, d1z)), ri2 = rsqrtf(dist_sq(d2x, d2y, d2z)), ri12 = __fmul_rn(ri1, ri2);
    float c = __fmul_rn(__fmaf_rn(d1z, d2z, __fmaf_rn(d1y, d2y, __fmul_rn(d1x, d2x))), ri12);
    c = fminf(fmaxf(c, -1.0f), 1.0f);
    c_out = c;
    const float a11 = __fmul_rn(__fmul_rn(K * ESB_BSCALE, c), __fmul_rn(ri1, ri1)), a12 = -__fmul_rn(K * ESB_BSCALE, ri12);
    I3 o;
    o.x = __float2int_rn(__fmaf_rn(a11, d1x, __fmul_rn(a12, d2x)));
    o.y = __float2int_rn(__fmaf_rn(a11, d1y, __fmul_rn(a12, d2y)));
    o.z = __float2int_rn(__fmaf_rn(a11, d1z, __fmul_rn(a12, d2z)));
    return o;
}

template <int NP, bool ENERGY>
__device__ __forceinline__ void bonded_forces(const Smem &S, const KArgs &a, const int r, const float r0_1, const float r0_2,
                                              double &e_bond, double &e_angle)
{
    const int n_topo = a.n_topo, ntp = a.n_topo_pad;
    const uint2 *bterm = a.bterm + (size_t)r * a.bterm_stride;     // per replica when the actin topology is dynamic
#if ESB_HALF
    const unsigned int gbase = smem_base() + (unsigned int)(reinterpret_cast<unsigned char *>(S.gx) - smem_raw), gstride = 4u * (unsigned int)ntp;
#endif
#if ESB_CLUSTER
    for (int k = threadIdx.x; k < S.misc[M_NMINE]; k += ESB_THREADS) {   // the other beads belong to other CTAs of the cluster
        const int i = S.mine[k];
#else
    for (int i = threadIdx.x; i < n_topo; i += ESB_THREADS) {
#endif
        int fx = 0, fy = 0, fz = 0;
        if (i < n_topo) {
            const float4 pi = S.pos[i];
#pragma unroll 1
            for (int t = 0; t < ESB_MAX_TERMS; t++) {
                // (with actin dynamics another SM may have rewritten the row one chunk ago: bypass L1 then)
                const uint2 term = a.bterm_stride ? __ldcg(&bterm[(size_t)t * ntp + i]) : bterm[(size_t)t * ntp + i];
                const int kind = term.y & 0xff, type = (term.y >> 8) & 0xff;
                if (kind == 0) break;                                         // a bead's terms are stored without gaps
                const int ja = (int)(term.x & 0xffffu);
                const float4 pa = S.pos[ja];
                float e;
                if (kind == 1) {
#if ESB_HALF
                    if (ja < i) continue;                                     // the other end evaluates this bond
#endif
                    const I3 s = bond_share(pi, pa, type == 1 ? r0_1 : r0_2, a.bond_k[type], e);
                    fx += s.x; fy += s.y; fz += s.z;
#if ESB_HALF
                    red_add3s(gbase + 4u * (unsigned int)ja, gstride, -s.x, -s.y, -s.z);
                    if (ENERGY) e_bond += (double)e;
#else
                    if (ENERGY) e_bond += 0.5 * (double)e;                    // each bond is visited from both ends
#endif
                    continue;
                }
#if ESB_HALF
                if (kind == 2) break;                                         // angles are evaluated by their vertex (the end rows come last)
#endif
                const int jb = (int)(term.x >> 16);
                const float4 pb = S.pos[jb];
                const float k = a.angle_k[type];
                if (kind == 2) {
                    // this bead is an end (i1); pa = vertex (i2), pb = other end (i3)
                    const I3 s = angle_end_share(pi, pa, pb, k, e);
                    fx += s.x; fy += s.y; fz += s.z;
                } else {
                    // this bead is the vertex (i2); pa = i1, pb = i3
                    const I3 s1 = angle_end_share(pa, pi, pb, k, e), s3 = angle_end_share(pb, pi, pa, k, e);
                    fx -= s1.x + s3.x; fy -= s1.y + s3.y; fz -= s1.z + s3.z;
#if ESB_HALF
                    red_add3s(gbase + 4u * (unsigned int)ja, gstride, s1.x, s1.y, s1.z);
                    red_add3s(gbase + 4u * (unsigned int)jb, gstride, s3.x, s3.y, s3.z);
#endif
                    if (ENERGY) e_angle += (double)(k * (1.0f + e));
                }
            }
#if ESB_HALF
            red_add3s(gbase + 4u * (unsigned int)i, gstride, fx, fy, fz);
#else
            S.gx[i] = __int_as_float(fx); S.gy[i] = __int_as_float(fy); S.gz[i] = __int_as_float(fz);
#endif
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Force evaluation: bonded prologue, then the pair loop gathered per bead by a group of ESB_G lanes.
//
// MODE 1  [LAMMPS-ext] PairSoft::compute   E = A [1 + cos(pi r / rc)]   (run_single_shoebox.py:589-628)
// MODE 2  [LAMMPS-ext] PairTable::compute, tabstyle LOOKUP: itable = int((rsq - innersq) * invdelta),
//         fpair = f[itable] (run_single_shoebox.py:800-874).  The bin index comes from an fp32
//         estimate that is accepted only when it is provably on the same side of every bin edge as
//         the exact value, and is otherwise recomputed in fp64 from the fp32 coordinates with the
//         oracle's (unfused) operations: it is bit-exact.  The bin value is the LJsoft closed form at
//         the bin midpoint (ljsoft_potential.py:10-16) inside [p0hi, plo) and the LOOKUP array
//         entry elsewhere.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lds32f(const unsigned int addr)
{
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <int NP, int MODE, bool ENERGY, bool P16>
__device__ ESB_PHASE_F void force_phase(const KArgs &a, const int r, const float soft_a, const float r0_1, const float r0_2,
                                         double *energy /* [3] pair, bond, angle: per-thread partials (ENERGY only) */)
{
    const Smem S = carve<NP>(a);
    const unsigned int FULL = 0xffffffffu;
    const int tid = threadIdx.x, n = a.n_atoms;
    const int lane = tid & 31;
    const int gl = lane & (ESB_G - 1), grp = lane / ESB_G;
    double e_pair = 0.0, e_bond = 0.0, e_angle = 0.0;
    // no barrier between the two: the bonded pass writes S.g*, the pair loop S.f*; a warp whose beads have no
    // bonded terms starts on the quads while the others wait for their term rows.  (S.misc[M_QCTR] is 0 on entry.)
    bonded_forces<NP, ENERGY>(S, a, r, r0_1, r0_2, e_bond, e_angle);
    const unsigned int *pool = a.nl_pool + (size_t)r * a.n_pad * (P16 ? ESB_NL_STRIDE / 2 : ESB_NL_STRIDE);
    const uint2 *qinfo = a.nl_qinfo + (size_t)r * 2 * a.n_pad;
#if ESB_CLUSTER
    const unsigned int sb0 = smem_base(), tbase = sb0 + 16u * (unsigned int)tab_off4<NP>(a);
    const unsigned int sb = sb0 + (unsigned int)(reinterpret_cast<unsigned char *>(S.pos) - smem_raw);      // current position buffer
    const unsigned int sb0_ = sb0;
#else
    const unsigned int sb = smem_base(), tbase = sb + 16u * (unsigned int)tab_off4<NP>(a);
    const unsigned int sb0_ = sb;
#endif
    const float *tabval = a.tabval;
    const int tlm1 = a.tlm1;
    const unsigned int fbase = sb0_ + force_off<NP>(a);          // shared address of the force accumulators
    unsigned int npairs = 0;
    // quads in order of decreasing list length, handed out dynamically (longest first: the warps finish
    // together); the next quad's descriptor is fetched (L2) while the current one is processed
    constexpr int BPW = 32 / ESB_G;                               // beads per warp: a "quad" of the force loop
#if ESB_HALF
    // work items (quad, chunk level): see ESB_CHUNK.  Lane l keeps the items up to and including levels l and l + 32.
    const int lev_lo = S.misc[M_LEVEL + lane], lev_hi = S.misc[M_LEVEL + 32 + lane];
    const int nquads = __shfl_sync(FULL, lev_hi, 31);             // (= number of items)
    int e_base = 0, e_base_next = 0;
    // Order of the items: the chunks beyond the first of the long lists (levels >= 1), then level 0, whose quads come in
    // order of decreasing length -- the last items handed out are the near-empty lists of the dilute species, so the
    // warps reach the barrier behind the phase together.
    const int n_rest = nquads - __shfl_sync(FULL, lev_lo, 0);     // items of the levels >= 1
    auto next_quad = [&]() {                                      // returns the quad of the next item, its level in e_base_next
        int t = 0;
        if (lane == 0) t = atomicAdd(&S.misc[M_QCTR], 1);         // (drawing the ticket one item ahead of its use: -4 %, the live ticket costs spills)
        t = __shfl_sync(FULL, t, 0);
        if (t >= nquads) return 0x7fffffff;
        if (ESB_ITEM_ORDER) {
            if (ESB_ITEM_FAST && t >= n_rest) { e_base_next = 0; return t - n_rest; }     // level 0 (nearly every item): nothing to decode
            t = t < n_rest ? t + (nquads - n_rest) : t - n_rest;
        }
        // (lists have at most ESB_NL_MAXNB entries: with chunks of 64 and more the levels 32..63 do not exist)
        const int lev = __popc(__ballot_sync(FULL, t >= lev_lo)) + (32 * ESB_CHUNK > ESB_NL_MAXNB ? 0 : __popc(__ballot_sync(FULL, t >= lev_hi)));
        e_base_next = lev * ESB_CHUNK;
        return t - (lev ? S.misc[M_LEVEL + lev - 1] : 0);
    };
#else
    const int nquads = (n + BPW - 1) / BPW;
    constexpr int e_base = 0;
    auto next_quad = [&]() {
        int t = 0;
        if (lane == 0) t = atomicAdd(&S.misc[M_QCTR], 1);
        return __shfl_sync(FULL, t, 0) * csize() + crank();
    };
#endif
    int q = next_quad();
#if ESB_HALF
    e_base = e_base_next;
#endif
    uint2 qi = make_uint2(0u, 0u);
    if (q < nquads && BPW * q + grp < n) qi = __ldcg(&qinfo[BPW * q + grp]);
    while (q < nquads) {
        const int qn = next_quad();
        uint2 qi_next = make_uint2(0u, 0u);
        if (qn < nquads && BPW * qn + grp < n) qi_next = __ldcg(&qinfo[BPW * qn + grp]);
        const bool valid = BPW * q + grp < n;
        const int i = (int)(qi.y & 0xffffu);
        // this item's part of the list: entries [e_base, e_base + ESB_CHUNK) (half list) or all of it
        const int cnt = valid ? (ESB_HALF ? min(max((int)(qi.y >> 16) - e_base, 0), ESB_CHUNK) : (int)(qi.y >> 16)) : 0;
        const unsigned int *nl = pool + qi.x + (P16 ? e_base / 2 : e_base);
        const float4 pi = lds128(sb + 16u * (unsigned int)i);
        const int cnt_max = __reduce_max_sync(FULL, cnt);
        int fx = 0, fy = 0, fz = 0;                               // fixed point, ESB_FSCALE_* counts per force unit
        // The list lives in L2/HBM: ESB_PREFETCH loads per lane are kept in flight (entries e, e+G, ...)
        // so that their latency overlaps the arithmetic of the entries before them.
        // Loads past the end of a list are harmless (every bead owns a fixed arena and the pool is padded),
        // so the prefetches carry no predicate; only the use of an entry is guarded.
        const unsigned int *lp = nl + gl;
        auto pair_body = [&](const unsigned int ent, const int e) {
            if (e < cnt) {
                const float4 pj = lds128(sb + 16u * (ent & (P16 ? 0x7ffu : 0xffffu)));
                // half list: the difference points to bead j, whose share goes out as it is; bead i's is its exact negative
                const float dx = ESB_HALF ? pj.x - pi.x : pi.x - pj.x, dy = ESB_HALF ? pj.y - pi.y : pi.y - pj.y, dz = ESB_HALF ? pj.z - pi.z : pi.z - pj.z;
                const float r2 = dist_sq(dx, dy, dz);
                float fpair;
                if (MODE == 1) {
                    const int ti = __float_as_int(pi.w) & 0xff, tj = __float_as_int(pj.w) & 0xff;
                    const float csq = S.setup->cutsq[ti * ESB_MAXT + tj];
                    if (!(r2 < csq)) return;
                    const float irc = rsqrtf(csq);
                    const float rr = sqrtf(r2);
                    const float xr = rr * irc;                               // r / rc in [0, 1)
                    fpair = (rr > 0.0f) ? soft_a * sinpif(xr) * 3.14159265358979323846f * irc / rr : 0.0f;
                    if (ENERGY) e_pair += (ESB_HALF ? 1.0 : 0.5) * (double)(soft_a * (1.0f + cospif(xr)));
                    npairs += ESB_HALF ? 2 : 1;
                } else {
                    const unsigned int tb = P16 ? (ent >> 11) & 0x1fu : ent >> 16;   // table id of the pair (fixed between builds)
                    const unsigned int T4 = tbase + 48u * tb;                         // DevTab = 3 float4 words
#if ESB_CUT32
                    if (!(r2 < lds32f(T4))) return;                                   // cutsq (one word: a quarter of the wavefronts of the full row)
                    const float4 t0 = lds128(T4);                                     // cutsq, invdelta_f, delta, xoff_f
#else
                    const float4 t0 = lds128(T4);                                     // cutsq, invdelta_f, delta, xoff_f
                    if (!(r2 < t0.x)) return;
#endif
                    const float4 t1 = lds128(T4 + 16u);                               // K, is6, c, patch
                    // bin estimate u = w - X with w = r2 * invdelta, X = innersq * invdelta (t0.w).  |u - u_exact| <= 5.4e-7 w:
                    // r2 carries 3e-7 relative (differences 2^-24 each, squares and sums), the scale, its product, X and the
                    // subtraction 2^-24 of at most w each -- relative to w, NOT to u: the offset amplifies the error of r2
                    const float w = r2 * t0.y;
                    const float u = w - t0.w;
                    int it = (int)u;
                    const float frac = u - (float)it;
                    const float margin = fmaf(w, 5.5e-7f, 1.0e-6f);
                    if (!(frac > margin && frac < 1.0f - margin)) {
                        const float4 t2f = lds128(T4 + 32u);
                        const double2 t2 = make_double2(__hiloint2double(__float_as_int(t2f.y), __float_as_int(t2f.x)), __hiloint2double(__float_as_int(t2f.w), __float_as_int(t2f.z)));
                        const double innersq = t2.x, invdelta = t2.y;
                        const double ddx = __dsub_rn((double)pi.x, (double)pj.x), ddy = __dsub_rn((double)pi.y, (double)pj.y),
                                     ddz = __dsub_rn((double)pi.z, (double)pj.z);
                        const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz));
                        if (rsq < innersq) { atomicOr(&S.misc[M_FAIL], ESB_ST_INNER); return; }
                        it = min((int)__dmul_rn(__dsub_rn(rsq, innersq), invdelta), tlm1 - 1);
                    }
                    const unsigned int patch = __float_as_uint(t1.w);            // p0hi | (plo - p0hi) << 16
                    const bool closed = (unsigned int)(it - (int)(patch & 0xffffu)) < (patch >> 16);
                    // closed form at the bin midpoint innersq + (it + 1/2) delta = (it + 1/2 + X) delta (bins [p0hi, plo) lie below rc)
                    const float qm = (((float)it + 0.5f) + t0.w) * t0.z;
                    const float q2 = qm * qm;
                    const float x = fmaf(q2 * qm, t1.y, t1.z);
                    const float inv = rcp_approx(x);                                // <= 1 ulp: 3 ulp on inv^3, far inside the 1e-5 parity bound
                    fpair = t1.x * q2 * (2.0f - x) * (inv * inv * inv);
                    if (!closed) fpair = tabval[tb * (unsigned int)tlm1 + it];   // predicated load: patch bins are rare
                    if (ENERGY) e_pair += (ESB_HALF ? 1.0 : 0.5) * (double)a.tabe[tb * (unsigned int)tlm1 + it];
                    npairs += ESB_HALF ? 2 : 1;
                }
                // round(f * scale) is odd in f, so bead j's share is exactly the negative of bead i's: the sums are the
                // same integers whether a pair is visited once (half list + scatter) or from both sides (full list)
                const float fs = fpair * ESB_FSCALE;
                const int ix = __float2int_rn(dx * fs), iy = __float2int_rn(dy * fs), iz = __float2int_rn(dz * fs);
                if (ESB_HALF) {
                    fx -= ix; fy -= iy; fz -= iz;
                    red_add3<NP>(a, fbase + 4u * (ent & (P16 ? 0x7ffu : 0xffffu)), ix, iy, iz);
                } else { fx += ix; fy += iy; fz += iz; }
            }
        };
        constexpr int DEPTH = ESB_PREFETCH;                       // list words in flight per lane
        unsigned int w[DEPTH];
#pragma unroll
        for (int k = 0; k < DEPTH; k++) w[k] = __ldcg(lp + k * ESB_G);
        if (P16) {                                                // word wi of the list = its entries 2 wi and 2 wi + 1
            for (int e0 = 0; e0 < cnt_max; e0 += DEPTH * 2 * ESB_G) {
                lp += DEPTH * ESB_G;
#pragma unroll
                for (int k = 0; k < DEPTH; k++) {
                    if (k > 0 && e0 + k * 2 * ESB_G >= cnt_max) break;
                    const unsigned int c = w[k];
                    w[k] = __ldcg(lp + k * ESB_G);
                    const int e = e0 + 2 * (k * ESB_G + gl);
                    pair_body(c, e); pair_body(c >> 16, e + 1);
                }
            }
        } else {
            for (int e0 = 0; e0 < cnt_max; e0 += DEPTH * ESB_G) {
                lp += DEPTH * ESB_G;
#pragma unroll
                for (int k = 0; k < DEPTH; k += 2) {
                    if (k > 0 && e0 + k * ESB_G >= cnt_max) break;
                    const unsigned int c0 = w[k], c1 = w[k + 1];
                    w[k] = __ldcg(lp + k * ESB_G); w[k + 1] = __ldcg(lp + (k + 1) * ESB_G);
                    pair_body(c0, e0 + k * ESB_G + gl); pair_body(c1, e0 + (k + 1) * ESB_G + gl);
                }
            }
        }
#pragma unroll
        for (int d = ESB_G >> 1; d >= 1; d >>= 1) {                // (every lane issuing its own atomics instead: -1 %)
            fx += __shfl_xor_sync(FULL, fx, d);
            fy += __shfl_xor_sync(FULL, fy, d);
            fz += __shfl_xor_sync(FULL, fz, d);
        }
        if (valid && gl == 0) {
            if (ESB_HALF) red_add3<NP>(a, fbase + 4u * (unsigned int)i, fx, fy, fz);      // other beads' lists scatter into this one too
            else { S.fx[i] = __int_as_float(fx); S.fy[i] = __int_as_float(fy); S.fz[i] = __int_as_float(fz); }
        }
        q = qn; qi = qi_next;
#if ESB_HALF
        e_base = e_base_next;
#endif
    }
    if (npairs) atomicAdd(&S.misc[M_PAIRS], (int)npairs);
    if (ENERGY) { energy[0] += e_pair; energy[1] += e_bond; energy[2] += e_angle; }
}

// ------------------------------------------------------------------------------------------------
// post_force fixes of one atom, in the order the reference creates them
// (run_single_shoebox.py:643-671): langevin -> wall/harmonic x6 -> setforce on the anchors.
// [LAMMPS-ext] FixLangevin::post_force (uniform noise), FixWallHarmonic::wall_particle, FixSetForce.
// ------------------------------------------------------------------------------------------------
template <bool ENERGY>
__device__ __forceinline__ int post_force(const KArgs &a, int i, const float4 p, const float vx, const float vy, const float vz,
                                          float &fx, float &fy, float &fz, const uint32_t k0, const uint32_t k1,
                                          const unsigned long long eval, double &e_wall)
{
    int fail = 0;
    if (a.langevin_mode) {
        float nx = 0.0f, ny = 0.0f, nz = 0.0f;
        if (a.langevin_mode == 1) {
            uint32_t o[4];
            philox4x32_10((uint32_t)i, (uint32_t)eval, (uint32_t)(eval >> 32), 0u, k0, k1, o);
            nx = u32_centered(o[0]); ny = u32_centered(o[1]); nz = u32_centered(o[2]);
        }
        fx += a.gamma1 * vx + a.gamma2 * nx;
        fy += a.gamma1 * vy + a.gamma2 * ny;
        fz += a.gamma1 * vz + a.gamma2 * nz;
    }
#define ESB_WALL(X, LO, HI, F)                                                                          \
    {                                                                                                   \
        const float dlo = (X) - (LO), dhi = (HI) - (X);                                                 \
        if (dlo < a.wall_cut) {                                                                         \
            if (!(dlo > 0.0f)) fail = ESB_ST_WALL;                                                      \
            else { const float dr = a.wall_cut - dlo; (F) += 2.0f * a.wall_eps * dr; if (ENERGY) e_wall += (double)(a.wall_eps * dr * dr); } \
        }                                                                                               \
        if (dhi < a.wall_cut) {                                                                         \
            if (!(dhi > 0.0f)) fail = ESB_ST_WALL;                                                      \
            else { const float dr = a.wall_cut - dhi; (F) -= 2.0f * a.wall_eps * dr; if (ENERGY) e_wall += (double)(a.wall_eps * dr * dr); } \
        }                                                                                               \
        if (!(dlo == dlo)) fail = ESB_ST_NONFINITE;                                                     \
    }
    ESB_WALL(p.x, a.lo[0], a.hi[0], fx)
    ESB_WALL(p.y, a.lo[1], a.hi[1], fy)
    ESB_WALL(p.z, a.lo[2], a.hi[2], fz)
#undef ESB_WALL
    if (__float_as_int(p.w) & 0x100) { fx = 0.0f; fy = 0.0f; fz = 0.0f; }
    if (!(fx == fx) || !(fy == fy) || !(fz == fz)) fail |= ESB_ST_NONFINITE;
    return fail;
}

// ------------------------------------------------------------------------------------------------
// Per-atom update between two force evaluations.
// [LAMMPS-ext] FixNVE: final_integrate of step n (v += dt/2 f) and initial_integrate of step n+1
// (v += dt/2 f; x += dt v) are adjacent per-atom operations and are fused here
// (fix 1 all nve, timestep 0.005: run_single_shoebox.py:643, 680).
// Returns non-zero when this thread saw a bead move further than skin/2 since the last build;
// failures (wall crossing, non-finite values) are OR-ed into S.misc[M_FAIL].
// ------------------------------------------------------------------------------------------------
template <int NP>
__device__ ESB_PHASE_F int atom_phase(const KArgs &a, const bool do_final, const bool do_initial,
                                       const uint32_t k0, const uint32_t k1, const unsigned long long eval, const float inv_fscale)
{
    const Smem S = carve<NP>(a);
    int flags = 0;
    const float dtf = 0.5f * a.dt, dtv = a.dt;
    double dummy = 0.0;
    unsigned int fmax_seen = 0u;
#if ESB_CLUSTER
    for (int k = threadIdx.x; k < S.misc[M_NMINE]; k += ESB_THREADS) {   // integrated by its owner, which writes it into every CTA
        const int i = S.mine[k];
#else
    for (int i = threadIdx.x; i < a.n_atoms; i += ESB_THREADS) {
#endif
        float4 p = S.pos[i];
        float vx = S.vx[i], vy = S.vy[i], vz = S.vz[i];
        const bool topo = i < a.n_topo;
        // bonded + pair (fixed point -> fp32; the accumulators go back to zero for the next evaluation)
        const int ifx = __float_as_int(S.fx[i]), ify = __float_as_int(S.fy[i]), ifz = __float_as_int(S.fz[i]);
        if (ESB_HALF) { S.fx[i] = 0.0f; S.fy[i] = 0.0f; S.fz[i] = 0.0f; }
        float fx = (float)ifx * inv_fscale, fy = (float)ify * inv_fscale, fz = (float)ifz * inv_fscale;
        unsigned int gabs = 0u;
        if (topo) {                                                  // bonded shares: own accumulators, ESB_BSCALE
            const int igx = __float_as_int(S.gx[i]), igy = __float_as_int(S.gy[i]), igz = __float_as_int(S.gz[i]);
            if (ESB_HALF) { S.gx[i] = 0.0f; S.gy[i] = 0.0f; S.gz[i] = 0.0f; }
            fx += (float)igx * (1.0f / ESB_BSCALE); fy += (float)igy * (1.0f / ESB_BSCALE); fz += (float)igz * (1.0f / ESB_BSCALE);
            gabs = max(max((unsigned int)abs(igx), (unsigned int)abs(igy)), (unsigned int)abs(igz));
        }
        int fail = post_force<false>(a, i, p, vx, vy, vz, fx, fy, fz, k0, k1, eval, dummy);
        const unsigned int iabs = max(max((unsigned int)abs(ifx), (unsigned int)abs(ify)), (unsigned int)abs(ifz));
        if (ESB_FGUARD && max(iabs, gabs) > 0x40000000u) fail |= ESB_ST_NONFINITE;      // a force sum left half the fixed-point range
        fmax_seen = max(fmax_seen, iabs);
        if (fail) atomicOr(&S.misc[M_FAIL], fail);
        if (do_final) { vx += dtf * fx; vy += dtf * fy; vz += dtf * fz; }
        if (do_initial) {
            vx += dtf * fx; vy += dtf * fy; vz += dtf * fz;
            p.x += dtv * vx; p.y += dtv * vy; p.z += dtv * vz;
#if !ESB_CLUSTER
            S.pos[i] = p;
#endif
            const float ex = p.x - S.bx[i], ey = p.y - S.by[i], ez = p.z - S.bz[i];
            if (ex * ex + ey * ey + ez * ez > a.skin_half_sq) flags |= 1;
        }
#if ESB_CLUSTER
        if (do_initial)                                          // distributed shared memory: the other position buffer of every CTA
            for (int rk = 0; rk < csize(); rk++) peer(S.pos_next, rk)[i] = p;    // (nobody reads it during this step)
#endif
        S.vx[i] = vx; S.vy[i] = vy; S.vz[i] = vz;                // (cluster: only the owner needs the velocity until ownership changes)
    }
    {   // headroom diagnostic: largest pair-force component of the chunk in 2^-8 force units (one atomic per warp)
        fmax_seen = __reduce_max_sync(0xffffffffu, fmax_seen);
        const int v8 = (int)fminf((float)fmax_seen * inv_fscale * 256.0f, 2.0e9f);
        if ((threadIdx.x & 31) == 0 && v8 > S.misc[M_FMAX]) atomicMax(&S.misc[M_FMAX], v8);
    }
    return flags;
}

#if ESB_CLUSTER
// Velocities live with the bead's owner; they are handed to every CTA of the cluster before the ownership changes
// (a list build re-deals the beads) and before the state is stored.  The next cluster barrier publishes them.
template <int NP>
__device__ __noinline__ void publish_velocities(const KArgs &a)
{
    const Smem S = carve<NP>(a);
    for (int k = threadIdx.x; k < S.misc[M_NMINE]; k += ESB_THREADS) {
        const int i = S.mine[k];
        const float vx = S.vx[i], vy = S.vy[i], vz = S.vz[i];
        for (int rk = 0; rk < csize(); rk++)
            if (rk != crank()) { peer(S.vx, rk)[i] = vx; peer(S.vy, rk)[i] = vy; peer(S.vz, rk)[i] = vz; }
    }
}
#endif

// [LAMMPS-ext] Variable ramp(x,y): delta = step - start; if (delta != 0) delta /= stop - start; x + delta*(y-x),
// evaluated by fix adapt at set-up and every nevery steps (run_single_shoebox.py:627-639); `run 0` gives x.
__device__ __forceinline__ double ramp(const double v0, const double v1, long long step, long long rb, long long re)
{
    double delta = (double)(step - rb);
    if (delta != 0.0) delta /= (double)(re - rb);
    return v0 + delta * (v1 - v0);
}

// ------------------------------------------------------------------------------------------------
// What the drivers do between two `run` commands, on the device (run_single_shoebox.py:913-961,
// 1304-1400, 1457-1519 == VGM/run_single_GENES_STAGES.py:822-865, 871-978, 1033-1089).
// All arithmetic that decides a count or ends up in geneTrack.txt is fp64 with unfused
// multiply/add in numpy's / scipy's evaluation order, so that given the same coordinates the
// record is bit-identical to the reference expressions.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dist_sq3(const double ax, const double ay, const double az, const float4 q)
{
    const double dx = __dsub_rn(ax, (double)q.x), dy = __dsub_rn(ay, (double)q.y), dz = __dsub_rn(az, (double)q.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

template <int NP>
__device__ __noinline__ void observe(const KArgs &a, int r, const esb_chunk_desc &d, ReplicaScalars &rs, double delt, double sig_nm, int slot)
{
    const Smem S = carve<NP>(a);
    const int tid = threadIdx.x;
    const Layout L = a.lay;
    const int probe = L.gene_start - L.promoter_length + 2;
    for (int k = tid; k < 4 + ESB_N_SHELL; k += ESB_THREADS) S.misc[M_CNT_PROM + k] = 0;
    if (tid == 0) {
        // RRPosition = np.mean(atomPositions[regulatory_atoms], axis=0)   (:932)
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int k = 0; k < L.n_enh; k++) {
            const float4 q = S.pos[a.enh_idx[k]];
            sx = __dadd_rn(sx, (double)q.x); sy = __dadd_rn(sy, (double)q.y); sz = __dadd_rn(sz, (double)q.z);
        }
        const double ne = (double)L.n_enh;
        S.dmisc[D_RR] = __ddiv_rn(sx, ne); S.dmisc[D_RR + 1] = __ddiv_rn(sy, ne); S.dmisc[D_RR + 2] = __ddiv_rn(sz, ne);
        const float4 pp = S.pos[probe], pg = S.pos[L.gene_start];
        S.dmisc[D_PROBE] = pp.x; S.dmisc[D_PROBE + 1] = pp.y; S.dmisc[D_PROBE + 2] = pp.z;
        S.dmisc[D_GENE] = pg.x; S.dmisc[D_GENE + 1] = pg.y; S.dmisc[D_GENE + 2] = pg.z;
    }
    __syncthreads();
    {
        const double px = S.dmisc[D_PROBE], py = S.dmisc[D_PROBE + 1], pz = S.dmisc[D_PROBE + 2];
        const double gx = S.dmisc[D_GENE], gy = S.dmisc[D_GENE + 1], gz = S.dmisc[D_GENE + 2];
        int c_prom = 0, c_gene = 0;
        for (int k = tid; k < L.n_s5p; k += ESB_THREADS) {
            const float4 q = S.pos[a.s5p_idx[k]];
            // ds = cdist(PromoterPositions, ser5p_positions); IPs = ds < pol2release_radius   (:948-954)
            if (__dsqrt_rn(dist_sq3(px, py, pz, q)) < 3.0) c_prom++;
            if (__dsqrt_rn(dist_sq3(gx, gy, gz, q)) < 3.0) c_gene++;
            // min distance to any enhancer bead, binned on cluster_r (:1504-1519)
            const double qx = (double)q.x, qy = (double)q.y, qz = (double)q.z;
            double dmin = 1e300;
            for (int e = 0; e < L.n_enh; e++) dmin = fmin(dmin, dist_sq3(qx, qy, qz, S.pos[a.enh_idx[e]]));
            const double dm = __dsqrt_rn(dmin);
            int k0 = ESB_N_SHELL;
            for (int c = 0; c < ESB_N_SHELL; c++) if (dm < a.cluster_r[c]) { k0 = c; break; }
            if (k0 >= 1 && k0 < ESB_N_SHELL) atomicAdd(&S.misc[M_HIST + k0 - 1], 1);
        }
        if (c_prom) atomicAdd(&S.misc[M_CNT_PROM], c_prom);
        if (c_gene) atomicAdd(&S.misc[M_CNT_GENE], c_gene);
        int nrnp = 0, nrbp = 0;
        for (int i = tid; i < a.n_atoms; i += ESB_THREADS) {
            const int t = __float_as_int(S.pos[i].w) & 0xff;
            nrnp += (t == 5); nrbp += (t == 4);
        }
        if (nrnp) atomicAdd(&S.misc[M_NRNP], nrnp);
        if (nrbp) atomicAdd(&S.misc[M_NRBP], nrbp);
    }
    __syncthreads();
    if (tid == 0) {
        const int i = d.chunk_index;
        const int g = L.gene_start, p = L.promoter_length;
        const double rx = S.dmisc[D_RR], ry = S.dmisc[D_RR + 1], rz = S.dmisc[D_RR + 2];
        // d_rg (:934)
        const double d_rg = __dsqrt_rn(dist_sq3(rx, ry, rz, S.pos[g]));
        // d_rp: distance to the mean of the promoter beads g-1, g-2, ..., g-p (:946)
        double mx = 0.0, my = 0.0, mz = 0.0;
        for (int k = 0; k < p; k++) {
            const float4 q = S.pos[g - 1 - k];
            mx = __dadd_rn(mx, (double)q.x); my = __dadd_rn(my, (double)q.y); mz = __dadd_rn(mz, (double)q.z);
        }
        mx = __ddiv_rn(mx, (double)p); my = __ddiv_rn(my, (double)p); mz = __ddiv_rn(mz, (double)p);
        const double ex = __dsub_rn(rx, mx), ey = __dsub_rn(ry, my), ez = __dsub_rn(rz, mz);
        const double d_rp = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)));
        const int n_prom = S.misc[M_CNT_PROM], n_gene = S.misc[M_CNT_GENE];

        // --- gene state machine (:956-961, 1304-1400) ---
        auto set_types = [&](int first, int count, int t) {
            for (int k = 0; k < count; k++) {
                float4 q = S.pos[first + k];
                q.w = __int_as_float((__float_as_int(q.w) & ~0xff) | t);
                S.pos[first + k] = q;
            }
        };
        // which of the three processes run in this chunk is the host's decision (t_induction_on ==
        // t_activation_on == 150, :472-475; Flavopiridol stops activation at NRuns, :885-886)
        if (rs.gene_state == 2) { rs.duration++; rs.total_active++; }
        if (d.induce_on && rs.gene_state == 0) { set_types(g, 5, 6); rs.gene_state = 1; }
        if (d.activate_on && rs.gene_state == 1 && n_prom > rs.threshold) {
            uint32_t o[4];
            philox4x32_10((uint32_t)i, 0u, 0u, 1u, (uint32_t)rs.seed, (uint32_t)(rs.seed >> 32), o);
            const double u = (double)(o[0] >> 8) * 0x1p-24;
            if (!(u > rs.p_activation_d)) { rs.marked = 1; rs.delay = 0; rs.gene_state = 2; }
        }
        if (rs.marked) {
            if (rs.delay < 2) rs.delay++;
            else { rs.marked = 0; set_types(g, 5, 7); set_types(g - p, p, 11); }
        }
        if (d.inactivate_on && rs.gene_state == 2 && rs.duration >= 50) {
            rs.duration = 0; set_types(g, 5, 2); set_types(g - p, p, 10); rs.gene_state = 0;
        }
        // --- RBP -> RNP ramp (:1457-1470) ---
        if (d.expected_rnp >= 0) {
            int actual = S.misc[M_NRNP], nrbp = S.misc[M_NRBP];
            uint32_t m = 0;
            while (d.expected_rnp > actual && nrbp > 0) {
                uint32_t o[4];
                philox4x32_10((uint32_t)i, m, 0u, 2u, (uint32_t)rs.seed, (uint32_t)(rs.seed >> 32), o);
                int k = (int)(((unsigned long long)(o[0] >> 8) * (unsigned long long)nrbp) >> 24);
                for (int q = 0; q < a.n_atoms; q++) {
                    if ((__float_as_int(S.pos[q].w) & 0xff) == 4) {
                        if (k == 0) { set_types(q, 1, 5); break; }
                        k--;
                    }
                }
                actual++; nrbp--; m++;
            }
        }
        // --- geneTrack record (:1496) ---
        if (slot < a.max_chunks) {
            double *f = a.frames + ((size_t)r * a.max_chunks + slot) * ESB_FRAME_DOUBLES;
            f[0] = __dmul_rn((double)(i + 1), delt);
            f[1] = __dmul_rn(S.dmisc[D_PROBE], sig_nm);
            f[2] = __dmul_rn(S.dmisc[D_PROBE + 1], sig_nm);
            f[3] = __dmul_rn(S.dmisc[D_PROBE + 2], sig_nm);
            f[4] = __dmul_rn(d_rp, sig_nm);
            f[5] = __dmul_rn(d_rg, sig_nm);
            f[6] = (double)n_prom;
            f[7] = (double)n_gene;
            f[8] = (double)rs.gene_state;
            a.rawd[((size_t)r * a.max_chunks + slot) * 2] = d_rp;
            a.rawd[((size_t)r * a.max_chunks + slot) * 2 + 1] = d_rg;
        }
    }
    if (slot < a.max_chunks && a.hist)
        for (int k = tid; k < ESB_N_SHELL - 1; k += ESB_THREADS)
            a.hist[((size_t)r * a.max_chunks + slot) * (ESB_N_SHELL - 1) + k] = (unsigned short)S.misc[M_HIST + k];
    __syncthreads();
}



// ------------------------------------------------------------------------------------------------
// Synthetic microscopy (run_single_shoebox.py:152-213): the voxel counts np.histogramdd would give for the five
// channels, one channel at a time through a packed-u16 table in the (idle) force arrays.  Bin of a coordinate as in
// numpy: searchsorted(edges, x, side='right') - 1, a value on the last edge belongs to the last bin, anything outside
// [first edge, last edge] is dropped; comparisons in fp64 on the fp32 coordinates.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int hist_bin(const double x, const double *edges, const int n_edges)
{
    int c = 0;
    for (int k = 0; k < n_edges; k++) c += (edges[k] <= x);
    if (x == edges[n_edges - 1]) c -= 1;
    return (c >= 1 && c <= n_edges - 1) ? c - 1 : -1;
}

template <int NP>
__device__ __noinline__ void voxel_counts(const KArgs &a, const int r, const int frame)
{
    const Smem S = carve<NP>(a);
    const int tid = threadIdx.x;
    const int nx = a.mic_n[0] - 1, ny = a.mic_n[1] - 1, nz = a.mic_n[2] - 1, nvox = nx * ny * nz;
    unsigned int *cw = reinterpret_cast<unsigned int *>(S.fx);
    const unsigned int masks[5] = {1u << 8, (1u << 7) | (1u << 11), (1u << 1) | (1u << 2) | (1u << 6) | (1u << 7) | (1u << 9) | (1u << 10) | (1u << 11),
                                   (1u << 2) | (1u << 6) | (1u << 7) | (1u << 10) | (1u << 11), 1u << 3};
    unsigned short *out = a.mic_out + ((size_t)r * a.mic_frames + frame) * 5 * (size_t)nvox;
    for (int ch = 0; ch < 5; ch++) {
        for (int w = tid; w < (nvox + 1) / 2; w += ESB_THREADS) cw[w] = 0u;
        __syncthreads();
        for (int i = tid; i < a.n_atoms; i += ESB_THREADS) {
            const float4 p = S.pos[i];
            if (!((masks[ch] >> (__float_as_int(p.w) & 0xff)) & 1u)) continue;
            const int bx = hist_bin((double)p.x, a.mic_edges[0], a.mic_n[0]);
            const int by = hist_bin((double)p.y, a.mic_edges[1], a.mic_n[1]);
            const int bz = hist_bin((double)p.z, a.mic_edges[2], a.mic_n[2]);
            if ((bx | by | bz) < 0) continue;
            const int v = (bx * ny + by) * nz + bz;
            atomicAdd(&cw[v >> 1], 1u << (16 * (v & 1)));
        }
        __syncthreads();
        const unsigned short *c16 = reinterpret_cast<const unsigned short *>(cw);
        for (int v = tid; v < nvox; v += ESB_THREADS) out[(size_t)ch * nvox + v] = c16[v];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Actin events between two `run` commands (run_single_shoebox.py:969-1298): nucleation of two free
// monomers, polymerisation at the plus end (every chunk) and at the minus end (every 40th),
// depolymerisation at the minus end (every 5th) and at the plus end (i % 4000 == 0), each with the relocation of the moved monomer by
// rejection sampling (`set atom ID x y z`), followed by what `create_bonds single/bond 2`,
// `single/angle 2` and `delete_bonds ... remove special` do to the topology.  Filaments are linear, so the
// whole topology of the actin beads is two links per bead (neighbour towards the plus / minus end) from
// which bonds, angles and 1-2/1-3/1-4 exclusions are re-derived after the events.  The bookkeeping
// lists keep the reference's order (free_actin, filament_details).  Python's unseeded random /
// numpy.random are replaced by Philox draws keyed by (replica seed; chunk, draw index) -- the test-side
// restatement of these lines consumes the same stream in the same order; all decisions are taken in fp64 with
// numpy's evaluation order on the fp32 coordinates, so device and oracle agree event by event.
// ------------------------------------------------------------------------------------------------
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3_of(const float4 q) { D3 o; o.x = (double)q.x; o.y = (double)q.y; o.z = (double)q.z; return o; }
__device__ __forceinline__ double d3_dist(const D3 a, const D3 b)
{
    const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

template <int NP>
__device__ __noinline__ void actin_events(const KArgs &a, const int r, const esb_chunk_desc &d, const ReplicaScalars &rs, const int slot)
{
    const Smem S = carve<NP>(a);
    const int tid = threadIdx.x;
    const int first = a.act.first, count = a.act.count;
    // volatile: this replica's bookkeeping was last written by another SM (L1 is not coherent)
    volatile unsigned int *link = a.act_link + (size_t)r * count;            // indexed by atom - first
    volatile unsigned short *L = a.act_lists + (size_t)r * ESB_ACT_WORDS(count);
    volatile unsigned short *fr = L + 2, *fplus = fr + count, *fminus = fplus + count;
    // scratch in the (idle) force arrays: stale free positions, a snapshot of the free list, flags
    double *fpos = reinterpret_cast<double *>(S.fx);                         // [3 * count]
    unsigned short *snap = reinterpret_cast<unsigned short *>(fpos + 3 * count);      // [count]
    unsigned char *flag = reinterpret_cast<unsigned char *>(snap + count);            // [count]
    auto lp = [&](int atom) { return (int)(link[atom - first] & 0xffffu); };          // towards the plus end
    auto lm = [&](int atom) { return (int)(link[atom - first] >> 16); };              // towards the minus end
    auto set_lp = [&](int atom, int v) { link[atom - first] = (link[atom - first] & 0xffff0000u) | (unsigned int)v; };
    auto set_lm = [&](int atom, int v) { link[atom - first] = (link[atom - first] & 0x0000ffffu) | ((unsigned int)v << 16); };
    if ((d.actin_flags & 1) && tid == 0) {
        const int i = d.chunk_index;
        uint32_t m = 0;
        auto draw = [&](uint32_t o[4]) { philox4x32_10((uint32_t)i, m++, 0u, 3u, (uint32_t)rs.seed, (uint32_t)(rs.seed >> 32), o); };
        auto uni = [](uint32_t w) { return (double)(w >> 8) * 0x1p-24; };
        auto place = [&](int atom, const D3 q) {
            float4 p = S.pos[atom];
            p.x = (float)q.x; p.y = (float)q.y; p.z = (float)q.z;
            S.pos[atom] = p;
        };
        // new_loc = np.random.uniform(low=c-h, high=c+h, size=(1,3))
        auto sample = [&](const D3 c, const double h) {
            uint32_t o[4];
            draw(o);
            D3 q;
            const double lx = __dsub_rn(c.x, h), ly = __dsub_rn(c.y, h), lz = __dsub_rn(c.z, h);
            q.x = __dadd_rn(lx, __dmul_rn(__dsub_rn(__dadd_rn(c.x, h), lx), uni(o[0])));
            q.y = __dadd_rn(ly, __dmul_rn(__dsub_rn(__dadd_rn(c.y, h), ly), uni(o[1])));
            q.z = __dadd_rn(lz, __dmul_rn(__dsub_rn(__dadd_rn(c.z, h), lz), uni(o[2])));
            return q;
        };
        auto free_remove = [&](int &nfree, int atom) {
            int w = 0;
            for (int k = 0; k < nfree; k++) if (fr[k] != atom) fr[w++] = fr[k];
            nfree = w;
        };
        int nfree = L[0], nfil = L[1];
        const double sh_lo = a.act.shell_lo, sh_hi = a.act.shell_hi, half = a.act.shell_hi;
        // ---------------- nucleation (:973-1027) ----------------
        {
            for (int k = 0; k < nfree; k++) { const D3 q = d3_of(S.pos[fr[k]]); fpos[3 * k] = q.x; fpos[3 * k + 1] = q.y; fpos[3 * k + 2] = q.z; flag[k] = 0; }
            unsigned short *pc = snap;                                        // pairs: plus-end atom, minus-end atom
            int npairs = 0;
            auto in_pairs = [&](int atom) { for (int q = 0; q < 2 * npairs; q++) if (pc[q] == atom) return true; return false; };
            for (int j = 0; j < nfree; j++) {
                uint32_t o[4];
                draw(o);
                if (!(uni(o[0]) <= a.act.p_nucleation)) continue;
                const int curr = fr[j];
                if (in_pairs(curr) || nfree < 2) continue;                   // (the reference finds this out later; no draw in between)
                const D3 c = d3_of(S.pos[curr]);
                double dmin = 1e300;
                int kmin = -1;
                for (int k = 0; k < nfree; k++) {
                    if (k == j) continue;
                    D3 q; q.x = fpos[3 * k]; q.y = fpos[3 * k + 1]; q.z = fpos[3 * k + 2];
                    const double dd = d3_dist(c, q);
                    if (dd < dmin) { dmin = dd; kmin = k; }
                }
                if (!(dmin < a.act.monomer_radius)) continue;
                const int tba = fr[kmin];
                if (in_pairs(tba)) continue;
                pc[2 * npairs] = (unsigned short)curr; pc[2 * npairs + 1] = (unsigned short)tba; npairs++;
                int n_tries = 1;
                D3 q;
                for (;;) {
                    n_tries++;
                    q = sample(c, half);
                    const double dd = d3_dist(c, q);
                    if (dd > sh_lo && dd < sh_hi) break;
                    if (n_tries > a.act.max_tries) break;
                }
                if (n_tries > a.act.max_tries) continue;
                place(tba, q);
            }
            for (int q = 0; q < npairs; q++) {                               // create_bonds single/bond 2 curr to_be_added (:1029-1038)
                const int curr = pc[2 * q], tba = pc[2 * q + 1];
                set_lm(curr, tba); set_lp(tba, curr);
                free_remove(nfree, curr); free_remove(nfree, tba);
                fplus[nfil] = (unsigned short)curr; fminus[nfil] = (unsigned short)tba; nfil++;
            }
        }
        // ---------------- polymerisation (:1062-1135) ----------------
        auto grow = [&](const int j, const bool plus_end) {
            int curr = plus_end ? fplus[j] : fminus[j];
            int adj = plus_end ? lm(curr) : lp(curr);
            const D3 c0 = d3_of(S.pos[curr]);
            const int n0 = nfree;
            for (int k = 0; k < n0; k++) { snap[k] = fr[k]; flag[k] = d3_dist(c0, d3_of(S.pos[fr[k]])) < a.act.end_radius; }
            for (int k = 0; k < n0; k++) {
                if (!flag[k]) continue;
                const int tba = snap[k];
                const D3 c = d3_of(S.pos[curr]), ca = d3_of(S.pos[adj]);
                int n_tries = 0;
                D3 q;
                for (;;) {
                    n_tries++;
                    q = sample(c, half);
                    const double dd = d3_dist(c, q), da = d3_dist(ca, q);
                    if (dd > sh_lo && dd < sh_hi && da > a.act.adj_min) break;
                    if (n_tries > a.act.max_tries) break;
                }
                if (n_tries > a.act.max_tries) continue;
                place(tba, q);
                if (plus_end) { set_lp(curr, tba); set_lm(tba, curr); fplus[j] = (unsigned short)tba; }
                else { set_lm(curr, tba); set_lp(tba, curr); fminus[j] = (unsigned short)tba; }
                free_remove(nfree, tba);
                adj = curr; curr = tba;
            }
        };
        {
            const int nf0 = nfil;
            for (int j = 0; j < nf0; j++) {
                uint32_t o[4];
                draw(o);
                if (uni(o[0]) <= a.act.p_on) grow(j, true);
                if (d.actin_flags & 2) {
                    draw(o);
                    if (uni(o[0]) <= a.act.p_on) grow(j, false);
                }
            }
        }
        // ---------------- depolymerisation at the minus end (:1168-1203) ----------------
        if (d.actin_flags & 4) {
            for (int j = 0; j < nfil; j++) {
                const int curr = fminus[j], nxt = lp(curr);
                link[curr - first] = 0xffffffffu;
                set_lm(nxt, ESB_NOLINK);
                fr[nfree++] = (unsigned short)curr;
                fminus[j] = (unsigned short)nxt;
                const D3 c = d3_of(S.pos[curr]), cn = d3_of(S.pos[nxt]);
                int n_tries = 0;
                D3 q;
                for (;;) {
                    n_tries++;
                    q = sample(c, a.act.free_box);
                    if (d3_dist(cn, q) > a.act.free_min) break;
                    if (n_tries > a.act.max_tries) break;
                }
                if (!(n_tries > a.act.max_tries)) place(curr, q);
                if (lp(nxt) == (int)ESB_NOLINK) {                            // a filament of length 1 is no filament (:1196-1203)
                    fr[nfree++] = (unsigned short)nxt;
                    fplus[j] = ESB_NOLINK;
                }
            }
            int w = 0;
            for (int j = 0; j < nfil; j++) if (fplus[j] != ESB_NOLINK) { fplus[w] = fplus[j]; fminus[w] = fminus[j]; w++; }
            nfil = w;
        }
        // ---------------- depolymerisation at the plus end (:1220-1250; i % t_off_plus == 0) ----------------
        if (d.actin_flags & 8) {
            for (int j = 0; j < nfil; j++) {
                const int curr = fplus[j], nxt = lm(curr);
                link[curr - first] = 0xffffffffu;
                set_lp(nxt, ESB_NOLINK);
                fr[nfree++] = (unsigned short)curr;
                fplus[j] = (unsigned short)nxt;
                const D3 c = d3_of(S.pos[curr]), cn = d3_of(S.pos[nxt]);
                int n_tries = 0;
                D3 q;
                for (;;) {
                    n_tries++;
                    q = sample(c, a.act.free_box);
                    if (d3_dist(cn, q) > a.act.free_min) break;
                    if (n_tries > a.act.max_tries) break;
                }
                if (!(n_tries > a.act.max_tries)) place(curr, q);
                if (lm(nxt) == (int)ESB_NOLINK) {                            // a filament of length 1 is no filament (:1245-1248)
                    fr[nfree++] = (unsigned short)nxt;
                    fplus[j] = ESB_NOLINK;
                }
            }
            int w = 0;
            for (int j = 0; j < nfil; j++) if (fplus[j] != ESB_NOLINK) { fplus[w] = fplus[j]; fminus[w] = fminus[j]; w++; }
            nfil = w;
        }
        L[0] = (unsigned short)nfree; L[1] = (unsigned short)nfil;
    }
    __syncthreads();
    // ---- bonds, angles and exclusions of the actin beads from their links (bonds, vertex, ends; no gaps) ----
    if (d.actin_flags & 1) {
        uint2 *bt = a.bterm + (size_t)r * a.bterm_stride;
        unsigned short *sp = a.spec + (size_t)r * a.spec_stride;
        const int ntp = a.n_topo_pad;
        for (int k = tid; k < count; k += ESB_THREADS) {
            const int b = first + k;
            const unsigned int NO = ESB_NOLINK;
            const unsigned int p1 = lp(b), m1 = lm(b);
            const unsigned int p2 = p1 != NO ? lp(p1) : NO, m2 = m1 != NO ? lm(m1) : NO;
            const unsigned int p3 = p2 != NO ? lp(p2) : NO, m3 = m2 != NO ? lm(m2) : NO;
            int nt = 0;
            auto add = [&](unsigned int x, unsigned int y, int kind) { bt[(size_t)nt * ntp + b] = make_uint2(x | (y << 16), (unsigned int)kind | (2u << 8)); nt++; };
            if (p1 != NO) add(p1, 0u, 1);
            if (m1 != NO) add(m1, 0u, 1);
            if (p1 != NO && m1 != NO) add(p1, m1, 3);
            if (p2 != NO) add(p1, p2, 2);
            if (m2 != NO) add(m1, m2, 2);
            for (; nt < ESB_MAX_TERMS; nt++) bt[(size_t)nt * ntp + b] = make_uint2(0u, 0u);
            unsigned short *row = sp + (size_t)b * ESB_MAX_SPEC;
            row[0] = (unsigned short)p1; row[1] = (unsigned short)p2; row[2] = (unsigned short)p3;
            row[3] = (unsigned short)m1; row[4] = (unsigned short)m2; row[5] = (unsigned short)m3;
            row[6] = 0xffff; row[7] = 0xffff;
        }
    }
    // ---- actinTrack / actinAroundCluster records of this chunk (:1523-1544) ----
    if (slot < a.max_chunks) {
        for (int k = tid; k < ESB_N_SHELL; k += ESB_THREADS) S.misc[M_HIST + k] = 0;
        __syncthreads();
        const Layout Lay = a.lay;
        for (int k = tid; k < count; k += ESB_THREADS) {
            const float4 q = S.pos[first + k];
            const double qx = (double)q.x, qy = (double)q.y, qz = (double)q.z;
            double dmin = 1e300;
            for (int e = 0; e < Lay.n_enh; e++) dmin = fmin(dmin, dist_sq3(qx, qy, qz, S.pos[a.enh_idx[e]]));
            const double dm = __dsqrt_rn(dmin);
            int k0 = ESB_N_SHELL;
            for (int c = 0; c < ESB_N_SHELL; c++) if (dm < a.cluster_r[c]) { k0 = c; break; }
            if (k0 >= 1 && k0 < ESB_N_SHELL) atomicAdd(&S.misc[M_HIST + k0 - 1], 1);
        }
        __syncthreads();
        for (int k = tid; k < ESB_N_SHELL - 1; k += ESB_THREADS)
            a.act_hist[((size_t)r * a.max_chunks + slot) * (ESB_N_SHELL - 1) + k] = (unsigned short)S.misc[M_HIST + k];
        if (tid == 0) {
            const int nfree = L[0], nfil = L[1];
            int sum = 0, sum2 = 0, mn = 1 << 30, mx = 0;
            for (int j = 0; j < nfil; j++) {
                int len = 1;
                for (int b = fplus[j]; lm(b) != (int)ESB_NOLINK; b = lm(b)) len++;
                sum += len; sum2 += len * len; mn = min(mn, len); mx = max(mx, len);
            }
            int *o = a.act_stats + ((size_t)r * a.max_chunks + slot) * 6;
            o[0] = nfil; o[1] = sum; o[2] = sum2; o[3] = nfree; o[4] = nfil ? mn : 0; o[5] = mx;
        }
    }
    __syncthreads();
}

}  // namespace

// ================================================================================================
// One persistent CTA per SM slot.  Work items are (replica, chunk) pairs handed out in chunk-major
// order from a global ticket counter; item (r, c) may start as soon as (r, c-1) is finished, which
// its producer publishes in progress[r].  With more replicas than resident CTAs the replicas thus
// advance in a ring and no SM idles while another still has a queue (a plain grid of R CTAs would
// leave 148*3 - (R mod 148*3) slots empty for the whole last wave).  The grid never exceeds the
// number of co-resident CTAs, so a waiting CTA's producer is always running.
// (own namespace per compiled variant: the two instantiations must not be merged by the linker)
namespace ESB_VARIANT(esb_step) {
template <int NP, bool SEG, bool P16>
__global__ void __launch_bounds__(ESB_THREADS, ESB_MIN_CTAS)
esb_run_kernel(const __grid_constant__ KArgs a, const int chunk_begin, const int n_chunks, const int n_seg_arg, const double delt, const double sig_nm,
               int *ticket, int *progress)
{
    const int n_seg = SEG ? n_seg_arg : 1;          // SEG = false: the plain one-item-per-chunk kernel, compiled without the segment code
    const Smem S = carve<NP>(a, true);
    const int tid = threadIdx.x;
    const int np = NP ? NP : a.n_pad;
    for (int k = tid; k < a.n_tables * (int)(sizeof(DevTab) / 4); k += ESB_THREADS)
        reinterpret_cast<int *>(S.tab)[k] = reinterpret_cast<const int *>(a.tabs)[k];
    __shared__ int s_item;
    __shared__ esb_chunk_desc s_desc;
    // A chunk can be cut into n_seg consecutive segments, each its own work item (same arithmetic, finer scheduling
    // grain: a launch of few chunks over more replicas than CTA slots then packs the SMs better).
    const int n_items = a.n_replicas * n_chunks * n_seg;

    const int crk = crank();
    for (;;) {
        __syncthreads();
#if ESB_CLUSTER
        csync();                                                     // nobody of the cluster still reads the previous item
        if (crk == 0 && tid == 0) {                                  // the cluster takes ONE item: rank 0 draws it and waits for its turn
            const int it_ = atomicAdd(ticket, 1);
            if (it_ < n_items) {
                while (atomicAdd(&progress[it_ % a.n_replicas], 0) < it_ / a.n_replicas) __nanosleep(256);
                __threadfence();
            }
            for (int rk = 0; rk < csize(); rk++) *peer(&s_item, rk) = it_;
        }
        csync();
#else
        if (tid == 0) s_item = atomicAdd(ticket, 1);
        __syncthreads();
#endif
        const int item = s_item;
        if (item >= n_items) break;
        const int r = item % a.n_replicas, seq = item / a.n_replicas;      // seq = (local chunk, segment) of this replica, in order
        const int lc = SEG ? seq / n_seg : seq, seg = SEG ? seq - lc * n_seg : 0;
        const int c = chunk_begin + lc;
#if !ESB_CLUSTER
        if (tid == 0) {
            while (atomicAdd(&progress[r], 0) < seq) __nanosleep(256);
            __threadfence();
        }
        __syncthreads();
#endif
        ReplicaScalars rs;
        {
            const volatile int *src = reinterpret_cast<const volatile int *>(&a.rs[r]);
            int *dst = reinterpret_cast<int *>(&rs);
            for (int k = 0; k < (int)(sizeof(ReplicaScalars) / 4); k++) dst[k] = src[k];
        }
        const int n_steps_c = a.descs[c].n_steps;
        const int seg_len = (n_steps_c + n_seg) / n_seg;                     // iterations 0 .. n_steps in n_seg pieces
        const int it0 = SEG ? seg * seg_len : 0, it1 = SEG ? min(it0 + seg_len, n_steps_c + 1) : n_steps_c + 1;
        if (rs.status != 0 || (SEG && it0 >= it1)) {                         // a failed replica no longer advances
            __syncthreads();
            if (tid == 0 && crk == 0) { __threadfence(); atomicExch(&progress[r], seq + 1); }
            continue;
        }
        for (int k = tid; k < np; k += ESB_THREADS) {
            S.pos[k] = __ldcg(&a.pos4[(size_t)r * np + k]);                 // written by another SM one chunk ago: bypass L1
            const float4 v = __ldcg(&a.vel4[(size_t)r * np + k]);
            S.vx[k] = v.x; S.vy[k] = v.y; S.vz[k] = v.z;
            S.fx[k] = 0.0f; S.fy[k] = 0.0f; S.fz[k] = 0.0f;       // (a segment that continues a chunk starts without a build)
        }
        if (ESB_HALF) for (int k = tid; k < 3 * a.n_topo_pad; k += ESB_THREADS) S.gx[k] = 0.0f;     // bonded accumulators (gx, gy, gz are contiguous)
        for (int k = tid; k < M_END; k += ESB_THREADS)
            S.misc[k] = (SEG && it0 > 0 && k == M_CURLIST) ? rs.curlist
                        : (SEG && ESB_HALF && it0 > 0 && k >= M_LEVEL && k < M_LEVEL + 64) ? __ldcg(&a.nl_levels[(size_t)r * 64 + (k - M_LEVEL)]) : 0;
        const uint32_t k0 = (uint32_t)rs.seed, k1 = (uint32_t)(rs.seed >> 32);
        long long *const L = reinterpret_cast<long long *>(S.dmisc);          // L_* slots: touched by thread 0 only
        if (tid == 0) { for (int k = L_EVALS; k < L_TMARK; k++) L[k] = 0; L[L_TMARK] = clock64(); }
#define ESB_TICK(acc) { if (tid == 0) { const long long now_ = clock64(); L[acc] += now_ - L[L_TMARK]; L[L_TMARK] = now_; } }
        int failed = 0;
        {   // the chunk descriptor: one copy per CTA in shared memory (a private copy would sit in L2-backed local memory)
            const int *src = reinterpret_cast<const int *>(&a.descs[c]);
            int *dst = reinterpret_cast<int *>(&s_desc);
            for (int k = tid; k < (int)(sizeof(esb_chunk_desc) / 4); k += ESB_THREADS) dst[k] = src[k];
        }
        __syncthreads();
        const esb_chunk_desc &d = s_desc;
        {
            const int *src = reinterpret_cast<const int *>(a.setups + (d.pair_mode == 1 ? 0 : 1 + d.map_id));
            int *dst = reinterpret_cast<int *>(S.setup);
            for (int k = tid; k < (int)(sizeof(PairSetup) / 4); k += ESB_THREADS) dst[k] = src[k];
        }
        // ---- Verlet::setup: fix adapt, neighbour build, one force evaluation; then n_steps x
        //      { nve initial; fix adapt; [build]; forces; post_force; nve final } ----
        // the scalars the step loop advances, as plain values (rs itself is addressable, i.e. in local memory)
        long long step = rs.step;
        unsigned long long eval = rs.eval;
        double soft_a = rs.soft_a, r0_1 = rs.r0[1], r0_2 = rs.r0[2];
        const long long first = step - it0, last = first + d.n_steps;        // (it0 steps of this run are done already)
        const long long rb = d.ramp_on ? d.ramp_start : first;
        const long long re = d.ramp_on ? d.ramp_stop : last;
        if (!SEG || it0 == 0) {
            if (d.adapt_soft) soft_a = ramp(a.ramp_soft[0], a.ramp_soft[1], step, rb, re);
            if (d.adapt_bond1) r0_1 = ramp(a.ramp_bond[1][0], a.ramp_bond[1][1], step, rb, re);
            if (d.adapt_bond2) r0_2 = ramp(a.ramp_bond[2][0], a.ramp_bond[2][1], step, rb, re);
        } else {
            for (int k = tid; k < a.n_atoms; k += ESB_THREADS) {             // positions at the last build: the rebuild vote goes on
                const float4 b = __ldcg(&a.bref4[(size_t)r * np + k]);
                S.bx[k] = b.x; S.by[k] = b.y; S.bz[k] = b.z;
            }
        }
        __syncthreads();
        ESB_TICK(L_TOBS)
        // [LAMMPS-ext] Neighbor::decide with `neigh_modify every 1 delay 10 check yes` (run_single_shoebox.py:533):
        // every run builds at set-up; afterwards a build needs `delay` steps since the last one AND a bead that
        // moved further than skin/2.  A build already due at the first permitted step is a "dangerous build".
        int rebuild = (!SEG || it0 == 0) ? 1 : rs.rebuild, ago = (!SEG || it0 == 0) ? 0 : rs.ago;
        const int delay = max(a.neigh_delay, 1);
        const bool last_seg = SEG ? it1 == d.n_steps + 1 : true;
        for (int it = it0; it < it1; it++) {
            if (rebuild) { build_lists<NP, P16>(a, r); if (tid == 0) { L[L_BUILDS]++; if (ago == delay) L[L_DANGER]++; } ago = 0; ESB_TICK(L_TBUILD) }
            if (d.pair_mode == 1) force_phase<NP, 1, false, P16>(a, r, (float)soft_a, (float)r0_1, (float)r0_2, nullptr);
            else force_phase<NP, 2, false, P16>(a, r, 0.0f, (float)r0_1, (float)r0_2, nullptr);
            __syncthreads();
            if (tid == 0) S.misc[M_QCTR] = 0;                        // (visible to the next force loop through the per-atom phase's barrier)
            ESB_TICK(L_TFORCE)
            if (tid == 0) { L[L_EVALS]++; L[L_LISTED] += (unsigned int)S.misc[M_CURLIST]; }
            // post_force of this evaluation, final integrate of the step it closes (it > 0) and initial
            // integrate of the next one (it < n_steps).  __syncthreads_or only tells whether ANY thread
            // passed non-zero, so it carries the rebuild vote; failure bits travel through shared memory.
            rebuild = __syncthreads_or(atom_phase<NP>(a, it > 0, it < d.n_steps, k0, k1, eval, 1.0f / ESB_FSCALE));
#if ESB_CLUSTER
            {   // ONE cluster barrier per step: the next positions went into the other buffer of every CTA, votes and
                // failure bits into slots of this step's parity (a CTA is never more than one barrier ahead)
                const int par = it & 1;
                if (tid == 0)
                    for (int rk = 0; rk < csize(); rk++) { peer(S.misc, rk)[M_VOTE + 8 * par + crk] = rebuild; peer(S.misc, rk)[M_CFAIL + 8 * par + crk] = S.misc[M_FAIL]; }
                csync();
                rebuild = 0;
                int cf = 0;
                for (int rk = 0; rk < csize(); rk++) { rebuild |= S.misc[M_VOTE + 8 * par + rk]; cf |= S.misc[M_CFAIL + 8 * par + rk]; }
                __syncthreads();
                if (tid == 0) { S.misc[M_FAIL] = S.misc[M_FAIL] | cf; if (it < d.n_steps) S.misc[M_PARITY] = S.misc[M_PARITY] ^ 1; }
                __syncthreads();
            }
#endif
            ago++;
            rebuild = rebuild && (ago >= delay);
#if ESB_CLUSTER
            if (rebuild && !S.misc[M_FAIL] && it < d.n_steps) publish_velocities<NP>(a);   // the build's barriers come before the next per-atom phase
#endif
            eval++;
            ESB_TICK(L_TATOM)
            failed = S.misc[M_FAIL];
            if (failed || it == d.n_steps) break;
            step++;
            if (tid == 0) L[L_STEPS]++;
            if (d.adapt_soft) soft_a = ramp(a.ramp_soft[0], a.ramp_soft[1], step, rb, re);
            if (step % 10 == 0) {
                if (d.adapt_bond1) r0_1 = ramp(a.ramp_bond[1][0], a.ramp_bond[1][1], step, rb, re);
                if (d.adapt_bond2) r0_2 = ramp(a.ramp_bond[2][0], a.ramp_bond[2][1], step, rb, re);
            }
        }
        const unsigned long long n_pairs = (unsigned int)S.misc[M_PAIRS];
#if ESB_CLUSTER
        publish_velocities<NP>(a);
        csync();                                                             // rank 0 stores the state: it needs every bead's velocity
#endif
        rs.step = step; rs.eval = eval; rs.soft_a = soft_a; rs.r0[1] = r0_1; rs.r0[2] = r0_2;
        if (SEG) { rs.ago = ago; rs.rebuild = rebuild; rs.curlist = S.misc[M_CURLIST]; }
        if (crk == 0) {                                                      // (cluster variant: every CTA holds the same state; one of them reports)
        if (last_seg && !failed && d.observe) observe<NP>(a, r, d, rs, delt, sig_nm, c);
        if (last_seg && !failed && d.observe && a.act.count > 0) actin_events<NP>(a, r, d, rs, c);
        if (last_seg && !failed && d.microscopy > 0 && a.mic_out && d.microscopy <= a.mic_frames) voxel_counts<NP>(a, r, d.microscopy - 1);
        }
        ESB_TICK(L_TOBS)
        __syncthreads();
        failed |= S.misc[M_FAIL];
#if ESB_CLUSTER
        if (tid == 0) {                                                       // pair statistics: every CTA saw its own quads
            unsigned long long *cn = a.counters + (size_t)r * ESB_N_COUNTERS;
            atomicAdd(&cn[2], n_pairs); atomicAdd(&cn[3], (unsigned long long)L[L_LISTED]);
        }
        if (crk != 0) continue;
#endif
#if ESB_CLUSTER
        const float4 *pos_now = carve<NP>(a).pos;                            // whichever buffer is current
#else
        const float4 *pos_now = S.pos;
#endif
        for (int k = tid; k < np; k += ESB_THREADS) {
            a.pos4[(size_t)r * np + k] = pos_now[k];
            a.vel4[(size_t)r * np + k] = make_float4(S.vx[k], S.vy[k], S.vz[k], 0.0f);
            if (SEG && !last_seg) a.bref4[(size_t)r * np + k] = make_float4(S.bx[k], S.by[k], S.bz[k], 0.0f);
        }
        if (tid == 0) {
            rs.status |= failed;
            a.rs[r] = rs;
            unsigned long long *cn = a.counters + (size_t)r * ESB_N_COUNTERS;
            cn[6] += (unsigned long long)L[L_TATOM]; cn[7] += (unsigned long long)L[L_TBUILD];
            cn[8] += (unsigned long long)L[L_TFORCE]; cn[9] += (unsigned long long)L[L_TOBS];
            cn[0] += (unsigned long long)L[L_EVALS]; cn[1] += (unsigned long long)L[L_BUILDS];
#if !ESB_CLUSTER
            cn[2] += n_pairs;
            cn[3] += (unsigned long long)L[L_LISTED];
#endif
            cn[4] += (unsigned long long)L[L_STEPS];
            cn[5] = (unsigned long long)rs.total_active;
            cn[10] += (unsigned long long)L[L_DANGER];
            if ((unsigned long long)S.misc[M_FMAX] > cn[11]) cn[11] = (unsigned long long)S.misc[M_FMAX];
        }
        __syncthreads();                         // every thread's state stores are issued ...
        if (tid == 0) { __threadfence(); atomicExch(&progress[r], seq + 1); }  // ... and published
    }
}

}  // namespace esb_step_v*
using ESB_VARIANT(esb_step)::esb_run_kernel;

#if ESB_TAG == 512
// Parity hook: one force evaluation at the stored positions (esb_forces()).
template <int NP>
__global__ void __launch_bounds__(ESB_THREADS, 2)
esb_forces_kernel(const __grid_constant__ KArgs a, const int pair_mode, const int setup_id, const float soft_a, const int with_post,
                  const unsigned long long eval, double *f_out, double *e_out)
{
    const Smem S = carve<NP>(a);
    const int r = blockIdx.x, tid = threadIdx.x;
    const int np = NP ? NP : a.n_pad;
    const ReplicaScalars rs = a.rs[r];
    for (int k = tid; k < np; k += ESB_THREADS) {
        S.pos[k] = a.pos4[(size_t)r * np + k];
        const float4 v = a.vel4[(size_t)r * np + k];
        S.vx[k] = v.x; S.vy[k] = v.y; S.vz[k] = v.z;
    }
    for (int k = tid; k < a.n_tables * (int)(sizeof(DevTab) / 4); k += ESB_THREADS)
        reinterpret_cast<int *>(S.tab)[k] = reinterpret_cast<const int *>(a.tabs)[k];
    for (int k = tid; k < M_END; k += ESB_THREADS) S.misc[k] = 0;
    for (int k = tid; k < 3 * a.n_topo_pad; k += ESB_THREADS) S.gx[k] = 0.0f;
    {
        const int *src = reinterpret_cast<const int *>(a.setups + setup_id);
        int *dst = reinterpret_cast<int *>(S.setup);
        for (int k = tid; k < (int)(sizeof(PairSetup) / 4); k += ESB_THREADS) dst[k] = src[k];
    }
    __syncthreads();
    build_lists<NP, false>(a, r);
    double en[3] = {0.0, 0.0, 0.0};
    if (pair_mode == 1) force_phase<NP, 1, true, false>(a, r, soft_a, (float)rs.r0[1], (float)rs.r0[2], en);
    else if (pair_mode == 2) force_phase<NP, 2, true, false>(a, r, 0.0f, (float)rs.r0[1], (float)rs.r0[2], en);
    else {
        double eb = 0.0, ea = 0.0;
        for (int k = tid; k < np; k += ESB_THREADS) { S.fx[k] = 0.0f; S.fy[k] = 0.0f; S.fz[k] = 0.0f; }
        __syncthreads();
        bonded_forces<NP, true>(S, a, r, (float)rs.r0[1], (float)rs.r0[2], eb, ea); en[1] = eb; en[2] = ea;
    }
    __syncthreads();
    double e_wall = 0.0;
    for (int i = tid; i < a.n_atoms; i += ESB_THREADS) {
        const bool topo = i < a.n_topo;
        const float inv_fscale = 1.0f / ESB_FSCALE;
        float fx = (float)__float_as_int(S.fx[i]) * inv_fscale, fy = (float)__float_as_int(S.fy[i]) * inv_fscale, fz = (float)__float_as_int(S.fz[i]) * inv_fscale;
        if (topo) {
            fx += (float)__float_as_int(S.gx[i]) * (1.0f / ESB_BSCALE); fy += (float)__float_as_int(S.gy[i]) * (1.0f / ESB_BSCALE);
            fz += (float)__float_as_int(S.gz[i]) * (1.0f / ESB_BSCALE);
        }
        if (with_post) {
            const int fail = post_force<true>(a, i, S.pos[i], S.vx[i], S.vy[i], S.vz[i], fx, fy, fz, (uint32_t)rs.seed, (uint32_t)(rs.seed >> 32), eval, e_wall);
            if (fail) atomicOr(&S.misc[M_FAIL], fail);
        }
        double *o = f_out + ((size_t)r * a.n_atoms + i) * 3;
        o[0] = fx; o[1] = fy; o[2] = fz;
    }
    // block-reduce the four energies through shared memory (the force arrays are free now)
    __syncthreads();
    double *red = reinterpret_cast<double *>(S.fx);
    const double mine[4] = {en[0], en[1], en[2], e_wall};
    for (int q = 0; q < 4; q++) red[q * ESB_THREADS + tid] = mine[q];
    __syncthreads();
    if (tid < 4) {
        double s = 0.0;
        for (int k = 0; k < ESB_THREADS; k++) s += red[tid * ESB_THREADS + k];
        e_out[(size_t)r * 4 + tid] = s;
    }
    if (tid == 0 && S.misc[M_FAIL]) atomicOr(&a.rs[r].status, S.misc[M_FAIL]);
}

// double [R][N][3] (+types) <-> float4 state.  frozen beads carry bit 8 of the type word.
extern "C" __global__ void esb_pack_kernel(const double *x, const double *v, const int *types, const unsigned char *frozen,
                                           float4 *pos4, float4 *vel4, int n_atoms, int n_pad, int n_replicas, int replicate)
{
    const size_t total = (size_t)n_replicas * n_pad;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(q / n_pad), i = (int)(q % n_pad);
        float4 p = make_float4(0.f, 0.f, 0.f, __int_as_float(0)), w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n_atoms) {
            const size_t s = ((size_t)(replicate ? 0 : r) * n_atoms + i);
            const int t = types[s] | (frozen[i] ? 0x100 : 0);
            p = make_float4((float)x[3 * s], (float)x[3 * s + 1], (float)x[3 * s + 2], __int_as_float(t));
            if (v) w = make_float4((float)v[3 * s], (float)v[3 * s + 1], (float)v[3 * s + 2], 0.f);
        }
        pos4[q] = p; vel4[q] = w;
    }
}

extern "C" __global__ void esb_unpack_kernel(const float4 *pos4, const float4 *vel4, double *x, double *v, int *types,
                                             int n_atoms, int n_pad, int n_replicas)
{
    const size_t total = (size_t)n_replicas * n_atoms;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(q / n_atoms), i = (int)(q % n_atoms);
        const float4 p = pos4[(size_t)r * n_pad + i], w = vel4[(size_t)r * n_pad + i];
        if (x) { x[3 * q] = p.x; x[3 * q + 1] = p.y; x[3 * q + 2] = p.z; }
        if (v) { v[3 * q] = w.x; v[3 * q + 1] = w.y; v[3 * q + 2] = w.z; }
        if (types) types[q] = __float_as_int(p.w) & 0xff;
    }
}

#endif   // ESB_TAG == 512 (parity hook and the pack/unpack kernels exist once)

size_t ESB_VARIANT(esb_smem_bytes)(int n_pad, int n_tables, int n_topo_pad) { return smem_bytes_for(n_pad, n_tables, n_topo_pad); }

// Host launchers: the kernels are instantiated for the bead capacities of the reference's box sizes
// (box 9..12 without actin: 1050/1172/1290/1400 beads -> 1056/1184/1312/1408; box 11 with actin:
// 1593 -> 1600) so that all shared-memory array offsets are immediates; any other size runs the
// generic instantiation.
#define ESB_FOR_NP(NP_, ...)         \
    switch (NP_) {                   \
        case 1056: { constexpr int NPC = 1056; __VA_ARGS__; break; } \
        case 1184: { constexpr int NPC = 1184; __VA_ARGS__; break; } \
        case 1312: { constexpr int NPC = 1312; __VA_ARGS__; break; } \
        case 1408: { constexpr int NPC = 1408; __VA_ARGS__; break; } \
        case 1600: { constexpr int NPC = 1600; __VA_ARGS__; break; } \
        default: { constexpr int NPC = 0; __VA_ARGS__; break; }      \
    }

#if ESB_CLUSTER
// Cluster variant: n_seg carries the cluster size (2, 4 or 8 CTAs per replica); one cluster per co-resident replica.
cudaError_t ESB_VARIANT(esb_launch_run)(const KArgs &k, int chunk_begin, int n_chunks, double delt, double sig_nm, size_t smem, cudaStream_t st,
                                        int *sched /* [1 + n_replicas] device ints: ticket, progress[] */, int n_sm, int cluster_size)
{
    (void)n_sm;
    cudaError_t e = cudaSuccess;
    ESB_FOR_NP(k.n_pad, {
        auto kern = k.pack16 ? esb_run_kernel<NPC, false, true> : esb_run_kernel<NPC, false, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned int)cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3((unsigned int)(cluster_size * k.n_replicas)); cfg.blockDim = dim3(ESB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n_clusters = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg);
        if (e == cudaSuccess && n_clusters < 1) e = cudaErrorLaunchOutOfResources;
        if (e == cudaSuccess) e = cudaMemsetAsync(sched, 0, sizeof(int) * (size_t)(1 + k.n_replicas), st);
        if (e == cudaSuccess) {
            cfg.gridDim = dim3((unsigned int)(cluster_size * (k.n_replicas < n_clusters ? k.n_replicas : n_clusters)));   // all clusters co-resident
            e = cudaLaunchKernelEx(&cfg, kern, k, chunk_begin, n_chunks, 1, delt, sig_nm, sched, sched + 1);
        }
    })
    return e;
}
#else
cudaError_t ESB_VARIANT(esb_launch_run)(const KArgs &k, int chunk_begin, int n_chunks, double delt, double sig_nm, size_t smem, cudaStream_t st,
                                        int *sched /* [1 + n_replicas] device ints: ticket, progress[] */, int n_sm, int n_seg)
{
    cudaError_t e = cudaSuccess;
    ESB_FOR_NP(k.n_pad, {
        auto kern = k.pack16 ? esb_run_kernel<NPC, false, true> : esb_run_kernel<NPC, false, false>;
        auto kern_seg = k.pack16 ? esb_run_kernel<NPC, true, true> : esb_run_kernel<NPC, true, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ESB_THREADS, smem);
        if (e == cudaSuccess && per_sm < 1) e = cudaErrorLaunchOutOfResources;
        if (e == cudaSuccess) e = cudaMemsetAsync(sched, 0, sizeof(int) * (size_t)(1 + k.n_replicas), st);
        if (e == cudaSuccess) {
            const int grid = k.n_replicas < per_sm * n_sm ? k.n_replicas : per_sm * n_sm;   // all CTAs co-resident
            // few chunks over more replicas than CTA slots: cut the chunks into segments so that the last wave is short
            // (measured, 512 replicas on 296 slots: +12 % for 1 chunk, +6 % for 5, -2 % for 10: segments cost ~3 %)
            if (n_seg <= 0) {
                const long long items = (long long)k.n_replicas * n_chunks;
                n_seg = (k.n_replicas <= grid || k.n_replicas % grid == 0 || items >= 8LL * grid) ? 1 : (items < 3LL * grid ? 8 : 4);
            }
            if (n_seg > 1) {
                e = cudaFuncSetAttribute(kern_seg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e == cudaSuccess) kern_seg<<<grid, ESB_THREADS, smem, st>>>(k, chunk_begin, n_chunks, n_seg, delt, sig_nm, sched, sched + 1);
            } else {
                kern<<<grid, ESB_THREADS, smem, st>>>(k, chunk_begin, n_chunks, 1, delt, sig_nm, sched, sched + 1);
            }
            e = cudaGetLastError();
        }
    })
    return e;
}

#endif   // !ESB_CLUSTER

#if ESB_TAG == 512
cudaError_t esb_launch_forces(const KArgs &k, int pair_mode, int setup_id, float soft_a, int with_post, unsigned long long eval,
                              double *f_out, double *e_out, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaSuccess;
    ESB_FOR_NP(k.n_pad, {
        e = cudaFuncSetAttribute(esb_forces_kernel<NPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            esb_forces_kernel<NPC><<<k.n_replicas, ESB_THREADS, smem, st>>>(k, pair_mode, setup_id, soft_a, with_post, eval, f_out, e_out);
            e = cudaGetLastError();
        }
    })
    return e;
}
#endif
