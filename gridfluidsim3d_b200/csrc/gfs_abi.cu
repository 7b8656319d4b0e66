// gfs_abi.cu -- the C-ABI of include/gfs_b200.h: context, device-resident domain, operator launches.
//
// Host code is C++11; all device memory is owned by the context; every entry point reports through the
// reference's error convention (trailing int *err, 1 = success, 0 = fail, message buffer), see
// /root/reference/src/c_bindings/cbindings.cpp:11-19.  No CPU fallback exists: without a usable CUDA
// device gfs_create fails and nothing else can be called.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "gfs_kernels.cuh"
#include "gfs_p2g2.cuh"
#include "gfs_g2p2.cuh"
#include "gfs_sources.cuh"
#include "gfs_pressure.cuh"

namespace {

thread_local char g_error[4096] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

struct GfsError : std::runtime_error {
    explicit GfsError(const std::string &m) : std::runtime_error(m) {}
};

#define GFS_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw GfsError(b_);                                                                    \
        }                                                                                          \
    } while (0)

#define GFS_REQUIRE(cond, msg)                                                                     \
    do {                                                                                           \
        if (!(cond)) throw GfsError(std::string(msg) + " (" #cond ")");                            \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        GFS_CUDA(cudaMalloc((void **)&p, n * sizeof(T)));
        cap = n;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace

struct ProfSpan { int name; cudaEvent_t a, b; };

struct gfs_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;

    // ---- optional per-kernel CUDA-event timing (gfs_profile_*): one event pair per launch
    bool profiling = false;
    std::vector<std::string> prof_names;
    std::vector<ProfSpan> prof_spans;
    std::vector<cudaEvent_t> prof_pool;

    int prof_name_id(const char *nm) {
        for (size_t i = 0; i < prof_names.size(); i++) if (prof_names[i] == nm) return (int)i;
        prof_names.push_back(nm);
        return (int)prof_names.size() - 1;
    }
    cudaEvent_t prof_event() {
        if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    int prof_begin(const char *nm) {
        if (!profiling) return -1;
        ProfSpan sp; sp.name = prof_name_id(nm); sp.a = prof_event(); sp.b = prof_event();
        cudaEventRecord(sp.a, stream);
        prof_spans.push_back(sp);
        return (int)prof_spans.size() - 1;
    }
    void prof_end(int id) { if (id >= 0) cudaEventRecord(prof_spans[id].b, stream); }

    // ---- domain
    bool has_domain = false;
    gfs::Grid grid;
    size_t face_count[3] = {0, 0, 0};
    size_t cell_count = 0;
    uint32_t nkeys = 0;
    DevBuf<float> field[3][3];            // [slot][comp]
    DevBuf<uint8_t> material;
    DevBuf<float> val[3];                 // node grids after normalisation ("ugrid")
    DevBuf<uint8_t> setmask[3];
    DevBuf<unsigned long long> acc[3];    // fixed-point accumulators, 2 per node
    DevBuf<int32_t> cell_start;           // nkeys + 3: exclusive scan of the per-cell counts (+ out-of-grid bin, dead bin, end)
    DevBuf<uint32_t> counts;              // nkeys + 3
    gfs::Sources sources;

    // ---- particles (double-buffered SoA: x,y,z,vx,vy,vz) + original-index tags
    int64_t n = 0;                        // slots in use (includes `dead`)
    int64_t dead = 0;                     // slots whose particle migrated away in the fused G2P: key nkeys + 1, last in sorted order
    int cur = 0;
    bool sorted = false;
    DevBuf<float> soa[2][6];
    DevBuf<int32_t> tag[2];
    DevBuf<uint32_t> keys[2];
    DevBuf<uint32_t> rank;
    DevBuf<int32_t> perm[2];
    bool keys_ready = false;
    bool storage_sorted = false;          // the SoA arrays are (nearly) in cell order: gathers through `index` stay coalesced
    bool indexed = false;                 // sorted order exists only as `index` (sorted slot -> storage slot); see k_build_index
    DevBuf<int32_t> index;              // keys/rank/counts/vmax of the current buffer were produced by the G2P epilogue
    DevBuf<unsigned char> cub_tmp;
    DevBuf<int32_t> n_valid;              // 1 word
    DevBuf<unsigned int> vmax_bits;       // 1 word
    DevBuf<int8_t> ext_layer;             // gfs_extrapolate: layer index per cell
    // ---- pressure solve (gfs_pressure.cuh): dense vectors over the cells, wavefront tiles, reduction partials
    struct Pressure {
        DevBuf<double> vec[6];            // r, z, s, p, q, precon
        DevBuf<double> scal;              // partial[3 * kPressBlocks], sigma[2], resid[1]
        DevBuf<uint8_t> flags;
        DevBuf<int> state, order;
        DevBuf<unsigned long long> ticket;
        DevBuf<unsigned int> tile_done;
        DevBuf<float> pressure;
        DevBuf<long long> trace;          // option 13: per-tile timestamps of the last substitution sweeps (debugging)
        int trace_on = 0;
        int dims[3] = {0, 0, 0};          // grid the tile order was built for
        unsigned int epoch = 0;
        bool valid = false;               // `pressure` holds the result of a solve on the current domain
        int *host_state = nullptr;        // pinned {done, iterations, spare} + resid behind it
        double *host_resid = nullptr;
    } press;
    DevBuf<float4> coll_list;             // particles advected into a solid cell: {slot, p1}, resolved by k_resolve_collisions
    DevBuf<unsigned int> coll_count;
    // ---- CUDA graphs of the fused single-domain substep: one per buffer parity, replayed while nothing it baked in changes
    struct SubstepGraph {
        cudaGraphExec_t exec = nullptr;
        uint64_t epoch = 0;               // graph_epoch at capture
        int64_t n = -1, launches = 0;
        double dt = 0, ratio = 0;
        int order = 0, interp = 0;
        const void *scratch[3] = {nullptr, nullptr, nullptr};       // lazily sized buffers the capture saw
    } graphs[2];
    uint64_t graph_epoch = 1;             // bumped by every call that changes what a captured substep baked in
    int use_graphs = 1;                   // option 4
    int64_t graph_replays = 0;
    int cell_cap = 0;                     // option 5: at most this many particles per cell survive a sort / G2P (0 = no cap)
    int remove_in_solid = 0;              // option 6: particles found inside solid cells by the binning are removed
    int64_t removed = 0;                  // particles removed by the two rules since creation
    unsigned int *removal_host = nullptr; // pinned word: dead-bin count read back after a binning pass
    int resolve_collisions = 1;           // option 3: 1 = the reference's collision resolve, 0 = solid test only (keep p0)
    DevBuf<unsigned long long> counters;  // [0] in_solid, [1] fluid cells, [2] solid hits, [3] spare
    int64_t out_of_grid = 0;
    int p2g_arith = 0;
    // ---- peer-memory exchange (one comm block per side: 0 = down, 1 = up; written by the neighbour on that side)
    struct CommSide {
        unsigned char *block = nullptr;       // my block: [flags 256 B][layers buf 0][layers buf 1][arrivals 0][arrivals 1]
        unsigned char *peer = nullptr;        // the neighbour's block for the opposite side, IPC-mapped
        unsigned int seq_layers = 0, seq_particles = 0;
        bool peer_ipc = false;                // `peer` came from cudaIpcOpenMemHandle (closed by gfs_destroy)
    } comm[2];
    size_t comm_layer_bytes = 0;          // capacity of one layers buffer
    int64_t comm_particle_cap = 0;        // capacity (particles) of one arrivals buffer
    unsigned int *comm_host = nullptr;    // pinned: counts published by k_gather_counts
    struct CommPlan {                     // the merged C1+C2 exchange of one side, as gfs_comm_set_plan stored it
        int n_push = 0, n_pull = 0;
        int push_what[16], push_first[16], push_count[16], pull_what[16], pull_first[16], pull_count[16], pull_add[16];
        int64_t push_off[16], pull_off[16];
    } comm_plan[2];
    bool comm_fused = false;              // the pending migration was done by the G2P kernel (leavers are dead slots)
    int comm_rank = -1, comm_world = 0;   // all-ranks table (k_allmax)
    unsigned long long *world_table = nullptr;        // mine: [2 parities][16 ranks]
    unsigned long long *world_peer[16] = {};          // everyone's, IPC-mapped (mine included)
    bool world_peer_ipc[16] = {};
    unsigned int seq_world = 0;
    DevBuf<unsigned int> comm_error;
    long long comm_timeout_cycles = 8000000000ll;     // option 7: device-side wait limit (SM clocks; ~4 s)
    int64_t coll_cap_user = 0;                          // option 8: collision list capacity in particles (0 = n/16 + 4096)
    int own_k0 = 0, own_k1 = 0;           // cell layers this context owns (z-slab sharding); grid kernels run on them + 1 halo
    uint32_t key_lo = 0, key_hi = 0;      // keys of the bricks around the owned layers (+- 8 layers): the cell table, the scan, the
    uint32_t brick_lo = 0, brick_hi = 0;  // count resets and the brick kernels cover only them (everything, single domain)
    bool velocities_valid = true;         // false after gfs_advect_substep (positions only): P2G / G2P need a fresh upload
    int press_variant = 2;                // option 12 (3 = 2 with per-value waiting inside the steps; measured slower): substitution sweeps of the pressure solve: 0 = global memory + tile flags, 1 = staged in shared memory + tile flags, 2 = staged + data-flow (sentinel) synchronisation
    int fused_grid = 1;                   // option 11: 1 = k_finalize_assemble (no node grid / mask in HBM), 0 = k_p2g_finalize + k_assemble
    bool acc_dirty = false;               // the accumulators still hold the previous splat (fused grid pass): memset before the next
    int split_wait = 0;                   // option 10: device-side waits in a single-thread kernel of their own (slabs sharing a GPU)
    int allmax_early = 1;                 // option 9: post the max right after G2P (1, default) or exchange it where it is needed (0)
    bool allmax_redo = false;             // the particle set was replaced after the post: consume it, then exchange afresh
    bool allmax_posted = false;           // this rank's max |v| of the coming substep is already on its way (post after G2P)
    DevBuf<unsigned int> split_counters;  // kept, down, up
    int p2g_variant = 3;                  // 0 = global atomics only; brick tiles in shared memory: 1 = round-1 kernel, 2 = round-2
                                          // kernel, 3 = round-2 kernel with the lane transposition (default)
    int lazy_sort = 1;                    // fused substep: sort by index only (no physical scatter)
    int g2p_variant = 2;                  // 0 = global loads only; TMA-staged brick tiles: 1 = round-1 kernel, 2 = round-2 trilinear
                                          // kernel on the 16-wide tile (default), 3 = on the 20-wide tile; tricubic: k_g2p_brick<1>
    gfs::BrickMaps maps[2];               // [interp]: NEW u,v,w + SAVED u,v,w tensor maps
    gfs::BrickMaps maps_tri_wide;         // dense boxes 20 columns wide (k_g2p_tri<true>: bank-conflict free row pitch)
    DevBuf<unsigned int> src_occ, src_count;   // sources: sub-cell occupancy bitmap, emission / survivor counter
    DevBuf<uint8_t> src_removal;          // outflow: cells whose particles go
    DevBuf<unsigned int> slow_count;      // k_g2p_tri's list of particles left to k_g2p_slow (the list itself lives in perm[0])
    bool have_maps = false;
    size_t field_floats[3] = {0, 0, 0};   // padded element counts of the resident u,v,w arrays

    // ---- scratch for host-pointer operators
    DevBuf<float> h_pos, h_out, h_val, h_fld, h_wgt, h_field[3];
    DevBuf<float> aos_stage;              // the AoS records of the last gfs_set_particles, kept for the first sort (k_gather_sorted_aos)
    bool aos_valid = false;               // aos_stage still equals the SoA storage, slot for slot (nothing moved or edited since the upload)
    DevBuf<uint8_t> h_mat;
    DevBuf<int8_t> h_layer;
    DevBuf<unsigned long long> h_acc;

    void reserve_particles(int64_t m) {
        for (int b = 0; b < 2; b++) {
            for (int a = 0; a < 6; a++) soa[b][a].reserve((size_t)m);
            tag[b].reserve((size_t)m);
            keys[b].reserve((size_t)m);
            perm[b].reserve((size_t)m);
        }
        rank.reserve((size_t)m);
        index.reserve((size_t)m);
    }
};

namespace {

using gfs::Grid;

Grid make_grid(int I, int J, int K, double dx, int k0, int k1, bool padded = false) {
    Grid g;
    g.I = I; g.J = J; g.K = K; g.k0 = k0; g.k1 = k1;
    g.dx = dx; g.invdx = 1.0 / dx;
    g.xmax = dx * I; g.ymax = dx * J; g.zmax = dx * K;
    g.halfdx = 0.5 * dx;
    g.dxf = (float)dx; g.invdxf = (float)g.invdx; g.halfdxf = (float)g.halfdx;
    g.xmaxf = (float)g.xmax; g.ymaxf = (float)g.ymax; g.zmaxf = (float)g.zmax;
    int e = 0;
    g.pow2 = (std::frexp(dx, &e) == 0.5 && I < (1 << 20) && J < (1 << 20) && K < (1 << 20) && e > -100 && e < 100) ? 1 : 0;
    const int ni[3] = {I + 1, I, I};
    // resident fields: 4 zero floats in front of every row (gfs::kRowPad), rows padded to whole 32-byte groups
    for (int a = 0; a < 3; a++) g.pitch[a] = padded ? (ni[a] + gfs::kRowPad + 7) / 8 * 8 : ni[a];
    g.nbi = (I + 1 + gfs::kBrick - 1) / gfs::kBrick;
    g.nbj = (J + 1 + gfs::kBrick - 1) / gfs::kBrick;
    g.nbk = (k1 - k0 + 1 + gfs::kBrick - 1) / gfs::kBrick;
    return g;
}

gfs::SplatParams make_splat(double r, const unsigned int *vmax_bits) {
    gfs::SplatParams sp;
    sp.radius = r; sp.rsq = r * r;
    sp.c1 = (4.0 / 9.0) * (1.0 / (r * r * r * r * r * r));
    sp.c2 = (17.0 / 9.0) * (1.0 / (r * r * r * r));
    sp.c3 = (22.0 / 9.0) * (1.0 / (r * r));
    sp.c1f = (float)sp.c1; sp.c2f = (float)sp.c2; sp.c3f = (float)sp.c3;
    sp.inv_rsq = 1.0 / sp.rsq;
    sp.inv_rsqf = (float)sp.inv_rsq;
    sp.rsqf = (float)sp.rsq;                     // smallest float >= rsq: d2 < rsqf  <=>  (double)d2 < rsq
    if ((double)sp.rsqf < sp.rsq) sp.rsqf = std::nextafterf(sp.rsqf, INFINITY);
    sp.vmax_bits = vmax_bits;
    return sp;
}

gfs::RkCoef make_rk(double dt) {       // casts exactly where the reference casts (particleadvector.cpp:1045-1078)
    gfs::RkCoef c;
    c.dt = (float)dt;
    c.half_dt = (float)(0.5 * dt);
    c.three_quarter_dt = (float)(0.75 * dt);
    c.dt_over_6 = (float)(dt / 6.0f);
    c.dt_over_9 = (float)(dt / 9.0f);
    return c;
}

dim3 grid3(int ni, int nj, int nk, int bx = 128) { return dim3((unsigned)ceil_div(ni, bx), (unsigned)nj, (unsigned)nk); }

#define LAUNCH(ctx, kernel, gridDim, blockDim, ...)                                                \
    do {                                                                                           \
        int prof_id_ = (ctx)->prof_begin(#kernel);                                                 \
        kernel<<<(gridDim), (blockDim), 0, (ctx)->stream>>>(__VA_ARGS__);                          \
        (ctx)->prof_end(prof_id_);                                                                 \
        (ctx)->launches++;                                                                         \
        GFS_CUDA(cudaGetLastError());                                                              \
    } while (0)

void ensure_capacity(gfs_context *c, int64_t n);
void require_domain(gfs_context *c) { GFS_REQUIRE(c && c->has_domain, "gfs_domain_init has not been called"); }

gfs::FieldPtrs field_ptrs(gfs_context *c, int slot) {
    gfs::FieldPtrs f;
    for (int a = 0; a < 3; a++) f.c[a] = c->field[slot][a].p + gfs::kRowPad;          // element (0,0,0)
    return f;
}

// K0.  stable = true: LSD radix sort of (key, index) pairs (cub), particles keep their relative order inside a
// cell -- what the exact-arithmetic P2G needs to reproduce the reference's summation order.  stable = false:
// counting sort (cell histogram with atomic tickets, exclusive scan, scatter); when the previous G2P already
// binned the advected positions in its epilogue only the scan and the scatter remain.
// With a removal rule on, a binning pass (k_hist or a G2P epilogue) may have put particles into the dead bin: the host
// needs their number to keep its slot accounting.  One 4-byte read and a stream synchronisation -- only when a rule is on.
bool removal_on(const gfs_context *c) { return c->cell_cap > 0 || c->remove_in_solid; }

int64_t read_dead_bin(gfs_context *c) {
    if (!c->removal_host) GFS_CUDA(cudaHostAlloc((void **)&c->removal_host, 64, cudaHostAllocDefault));
    GFS_CUDA(cudaMemcpyAsync(c->removal_host, c->counts.p + c->nkeys + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    return (int64_t)*c->removal_host;
}

// brick layers around the owned cell layers, one brick layer (8 cells) of slack each side: where this rank's particles can
// be between two migrations
void set_key_range(gfs_context *c) {
    const Grid &g = c->grid;
    const bool whole = c->own_k0 <= g.k0 && c->own_k1 >= g.k1;
    int bk_lo = 0, bk_hi = g.nbk;
    if (!whole) {
        bk_lo = (c->own_k0 - g.k0 - gfs::kBrick) / gfs::kBrick;
        bk_hi = (c->own_k1 - g.k0 + gfs::kBrick - 1) / gfs::kBrick + 1;
        if (bk_lo < 0) bk_lo = 0;
        if (bk_hi > g.nbk) bk_hi = g.nbk;
    }
    const uint32_t per_layer = (uint32_t)g.nbi * (uint32_t)g.nbj;
    c->brick_lo = (uint32_t)bk_lo * per_layer; c->brick_hi = (uint32_t)bk_hi * per_layer;
    c->key_lo = c->brick_lo * gfs::kBrickCells; c->key_hi = c->brick_hi * gfs::kBrickCells;
}

gfs::KeyRange key_range(const gfs_context *c) { gfs::KeyRange kr; kr.lo = c->key_lo; kr.hi = c->key_hi; return kr; }
bool full_range(const gfs_context *c) { return c->key_lo == 0 && c->key_hi == c->nkeys; }

// zero the cell counters the next binning pass will tick: the key range of this rank plus the three tail bins
void reset_counts(gfs_context *c) {
    if (full_range(c)) {
        GFS_CUDA(cudaMemsetAsync(c->counts.p, 0, sizeof(uint32_t) * ((size_t)c->nkeys + 3), c->stream));
    } else {
        GFS_CUDA(cudaMemsetAsync(c->counts.p + c->key_lo, 0, sizeof(uint32_t) * (size_t)(c->key_hi - c->key_lo), c->stream));
        GFS_CUDA(cudaMemsetAsync(c->counts.p + c->nkeys, 0, sizeof(uint32_t) * 3, c->stream));
    }
}

void scan_counts(gfs_context *c) {
    if (c->cell_cap > 0)
        LAUNCH(c, gfs::k_clamp_counts, ceil_div((long long)c->nkeys, 256), 256, c->nkeys, (uint32_t)c->cell_cap, c->counts.p);
    // single domain: one scan over every cell and the tail bins; z-slab rank: only the keys of its own bricks (an eighth of
    // the table at 8 GPUs), the tail bins fixed up by k_scan_tail
    const bool full = full_range(c);
    const size_t first = full ? 0 : c->key_lo, nbins = full ? (size_t)c->nkeys + 3 : (size_t)(c->key_hi - c->key_lo);
    size_t tmp_bytes = 0;
    GFS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->counts.p + first, (uint32_t *)c->cell_start.p + first, (int)nbins, c->stream));
    c->cub_tmp.reserve(tmp_bytes);
    int prof_id = c->prof_begin("cub::DeviceScan::ExclusiveSum");
    if (nbins > 0)
        GFS_CUDA(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp_bytes, c->counts.p + first, (uint32_t *)c->cell_start.p + first, (int)nbins, c->stream));
    c->prof_end(prof_id);
    c->launches += 2;
    if (!full) LAUNCH(c, gfs::k_scan_tail, 1, 1, c->counts.p, c->cell_start.p, c->key_lo, c->key_hi, c->nkeys);
}

// Dead slots (particles the fused G2P handed to a neighbour GPU) only make sense to kernels that walk the sorted
// index; anything else first compacts them away with one physical counting-sort scatter (dead bin = tail).
void drop_dead(gfs_context *c) {
    if (c->dead == 0) return;
    GFS_REQUIRE(c->keys_ready || c->indexed, "internal: dead slots without their keys");
    if (c->keys_ready) scan_counts(c);          // otherwise cell_start is the scan these keys were indexed with
    const int src = c->cur, dst = 1 - c->cur;
    LAUNCH(c, gfs::k_scatter_sorted, ceil_div(c->n, 256), 256, c->n, c->keys[0].p, c->rank.p, c->cell_start.p,
           c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,
           c->tag[src].p,
           c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,
           c->tag[dst].p);
    c->cur = dst;
    c->n -= c->dead;
    c->dead = 0;
    c->indexed = false;
    c->keys_ready = false;
    c->storage_sorted = true;
    c->sorted = true;           // physically sorted now, and cell_start describes it
    c->aos_valid = false;
}

void do_sort(gfs_context *c, bool stable, bool lazy = false) {
    require_domain(c);
    // freshly uploaded particles are in arbitrary (the reference: shuffled) order: an index over them would turn every
    // particle read of the P2G and G2P kernels into a random 4-byte gather (measured 13 ms instead of 3.3 ms per kernel at
    // 96 M particles).  The first sort after an upload therefore moves the data; later ones only re-index it.
    if (lazy && !c->storage_sorted) lazy = false;
    if (c->dead > 0 && !(lazy && !stable && c->keys_ready)) { drop_dead(c); c->sorted = false; c->aos_valid = false; }
    const int64_t n = c->n;
    if (c->resolve_collisions) {
        const size_t want = c->coll_cap_user > 0 ? (size_t)c->coll_cap_user : (size_t)(n / 16 + 4096);
        if (want > c->coll_list.cap) c->coll_list.reserve(c->coll_cap_user > 0 ? want : want + (size_t)n / 64);
        c->coll_count.reserve(1);
    }
    const int src = c->cur, dst = 1 - c->cur;
    const int B = 256;
    if (!c->keys_ready) {
        reset_counts(c);
        // (a max |v| already posted to the other ranks stays valid: this pass re-derives the LOCAL maximum of the same
        // global particle set -- e.g. after gfs_get_particles compacted dead slots on this rank only -- and whether a post is
        // outstanding must never depend on rank-local state, or the ranks' sequence numbers part ways)
        GFS_CUDA(cudaMemsetAsync(c->vmax_bits.p, 0, sizeof(unsigned int), c->stream));
        if (n > 0)
            LAUNCH(c, gfs::k_hist, ceil_div(n, B), B, c->grid, c->nkeys, c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p,
                   c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p, n, c->keys[0].p, c->rank.p, c->perm[0].p,
                   c->counts.p, c->vmax_bits.p, c->remove_in_solid ? c->material.p : (const uint8_t *)nullptr, (uint32_t)c->cell_cap, key_range(c));
        if (removal_on(c) && n > 0) { c->dead = read_dead_bin(c); c->removed += c->dead; }
    }
    scan_counts(c);
    if (n > 0) {
        if (stable) {
            GFS_REQUIRE(!c->keys_ready, "internal: stable sort after a binning G2P");
            int bits = 1;
            while (bits < 32 && (1ull << bits) <= (unsigned long long)c->nkeys) bits++;      // keys are in [0, nkeys]
            size_t tmp_bytes = 0;
            GFS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->keys[0].p, c->keys[1].p, c->perm[0].p, c->perm[1].p,
                                                      (int)n, 0, bits, c->stream));
            c->cub_tmp.reserve(tmp_bytes);
            int prof_id = c->prof_begin("cub::DeviceRadixSort::SortPairs");
            GFS_CUDA(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp_bytes, c->keys[0].p, c->keys[1].p, c->perm[0].p,
                                                      c->perm[1].p, (int)n, 0, bits, c->stream));
            c->prof_end(prof_id);
            c->launches += 4;      // cub: histogram + onesweep passes (counted conservatively)
            LAUNCH(c, gfs::k_reorder, ceil_div(n, B), B, n, c->perm[1].p,
                   c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,
                   c->tag[src].p,
                   c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,
                   c->tag[dst].p);
        } else if (lazy) {
            LAUNCH(c, gfs::k_build_index, ceil_div(n, 4 * B), B, n, c->keys[0].p, c->rank.p, c->cell_start.p, c->index.p);
        } else if (c->aos_valid && src == 0) {
            // first sort after an upload: gather the sorted SoA arrays from the uploaded AoS records
            LAUNCH(c, gfs::k_build_index, ceil_div(n, 4 * B), B, n, c->keys[0].p, c->rank.p, c->cell_start.p, c->index.p);
            LAUNCH(c, gfs::k_gather_sorted_aos, ceil_div(n, B), B, n, c->index.p, reinterpret_cast<const float2 *>(c->aos_stage.p),
                   c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p, c->tag[dst].p);
        } else {
            LAUNCH(c, gfs::k_scatter_sorted, ceil_div(n, B), B, n, c->keys[0].p, c->rank.p, c->cell_start.p,
                   c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,
                   c->tag[src].p,
                   c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,
                   c->tag[dst].p);
        }
        if (!(lazy && !stable)) {
            c->cur = dst; c->storage_sorted = true;
            c->n -= c->dead; c->dead = 0;            // the dead bin is the tail of the physically sorted arrays
        }
    }
    c->indexed = lazy && !stable && n > 0;
    c->keys_ready = false;
    c->sorted = true;
    if (!(lazy && !stable)) c->aos_valid = false;          // the storage order changed
}

// layer range [lo,hi) of cells the grid kernels process: the owned slab plus one halo layer each side
void work_layers(gfs_context *c, int *lo, int *hi) {
    *lo = c->own_k0 > 0 ? c->own_k0 - 1 : 0;
    *hi = c->own_k1 < c->grid.K ? c->own_k1 + 1 : c->grid.K;
}

// P2G, first half: classification of the work layers + splat of the resident particles into the accumulators
void do_p2g_begin(gfs_context *c, int arith) {
    require_domain(c);
    GFS_REQUIRE(c->sorted, "gfs_p2g needs gfs_sort first");
    GFS_REQUIRE(c->velocities_valid, "the particle velocities are undefined after gfs_advect_substep: upload the particles again");
    const Grid &g = c->grid;
    const int b = c->cur;
    gfs::SplatParams sp = make_splat(g.dx, c->vmax_bits.p);
    int lo, hi;
    work_layers(c, &lo, &hi);
    GFS_CUDA(cudaMemsetAsync(c->counters.p, 0, 2 * sizeof(unsigned long long), c->stream));
    const long long plane = (long long)g.I * g.J;
    LAUNCH(c, gfs::k_classify, ceil_div(plane * (hi - lo), 256), 256, g, c->cell_start.p, c->material.p, c->counters.p,
           plane * lo, plane * (hi - lo), c->own_k0, c->own_k1);
    c->p2g_arith = arith;
    GFS_REQUIRE(!(arith == GFS_EXACT && c->indexed), "internal: exact P2G needs physically sorted particles");
    if (arith == GFS_EXACT) return;               // the exact gather does everything in do_p2g_end
    if (c->acc_dirty) {
        // the fused grid pass leaves the accumulators as they are (its threads read their neighbours'): clear the node
        // layers this rank splats into or receives partial sums for
        const int kl_ = g.k1 - g.k0;
        const int dims_[3][2] = {{g.I + 1, g.J}, {g.I, g.J + 1}, {g.I, g.J}};
        for (int comp = 0; comp < 3; comp++) {
            const int layers = kl_ + (comp == 2 ? 1 : 0);
            int l0 = lo - 1, l1 = hi + 1 + (comp == 2 ? 1 : 0);
            if (l0 < 0) l0 = 0;
            if (l1 > layers) l1 = layers;
            const size_t plane = (size_t)dims_[comp][0] * dims_[comp][1] * 2;
            GFS_CUDA(cudaMemsetAsync(c->acc[comp].p + plane * (size_t)l0, 0, plane * (size_t)(l1 - l0) * sizeof(unsigned long long), c->stream));
        }
        c->acc_dirty = false;
    }
    if (c->n > 0) {
        const bool pow2 = g.pow2 != 0;
        const int32_t *idx = c->indexed ? c->index.p : nullptr;
        if (c->p2g_variant == 0) {
            if (pow2)
                LAUNCH(c, gfs::k_p2g_scatter<2>, ceil_div(c->n, 256), 256, g, sp, c->cell_start.p + c->nkeys, idx,
                       c->soa[b][0].p, c->soa[b][1].p, c->soa[b][2].p, c->soa[b][3].p, c->soa[b][4].p, c->soa[b][5].p,
                       c->acc[0].p, c->acc[1].p, c->acc[2].p);
            else
                LAUNCH(c, gfs::k_p2g_scatter<0>, ceil_div(c->n, 256), 256, g, sp, c->cell_start.p + c->nkeys, idx,
                       c->soa[b][0].p, c->soa[b][1].p, c->soa[b][2].p, c->soa[b][3].p, c->soa[b][4].p, c->soa[b][5].p,
                       c->acc[0].p, c->acc[1].p, c->acc[2].p);
        } else {
            const int nbricks = (int)(c->brick_hi - c->brick_lo);          // the bricks around the owned layers (all, single domain)
            const size_t smem = 12 * gfs::kTileNodes * sizeof(uint32_t);
#define GFS_TILE_ARGS g, sp, c->cell_start.p, c->brick_lo, idx, c->soa[b][0].p, c->soa[b][1].p, c->soa[b][2].p, c->soa[b][3].p, c->soa[b][4].p, \
                      c->soa[b][5].p, c->acc[0].p, c->acc[1].p, c->acc[2].p
            if (pow2 && c->p2g_variant == 3) {
                const size_t smem3 = smem + 6 * gfs::kStageChunk * sizeof(float);
                static bool attr_set = false;
                if (!attr_set) {
                    GFS_CUDA(cudaFuncSetAttribute(gfs::k_p2g_tile2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
                    attr_set = true;
                }
                int prof_id_ = c->prof_begin("gfs::k_p2g_tile2<1>");
                gfs::k_p2g_tile2<true><<<nbricks, 512, smem3, c->stream>>>(GFS_TILE_ARGS);
                c->prof_end(prof_id_);
            } else if (pow2 && c->p2g_variant == 2) {
                int prof_id_ = c->prof_begin("gfs::k_p2g_tile2<0>");
                gfs::k_p2g_tile2<false><<<nbricks, 256, smem, c->stream>>>(GFS_TILE_ARGS);
                c->prof_end(prof_id_);
            } else {
                int prof_id_ = c->prof_begin(pow2 ? "gfs::k_p2g_tile<2>" : "gfs::k_p2g_tile<0>");
                if (pow2) gfs::k_p2g_tile<2><<<nbricks, 256, smem, c->stream>>>(GFS_TILE_ARGS);
                else gfs::k_p2g_tile<0><<<nbricks, 256, smem, c->stream>>>(GFS_TILE_ARGS);
                c->prof_end(prof_id_);
            }
#undef GFS_TILE_ARGS
            c->launches++;
            GFS_CUDA(cudaGetLastError());
        }
    }
}

// P2G, second half: normalisation + inflow override + face assembly on the work layers
void do_p2g_end(gfs_context *c) {
    require_domain(c);
    const Grid &g = c->grid;
    const int kl = g.k1 - g.k0;
    const int b = c->cur;
    gfs::SplatParams sp = make_splat(g.dx, c->vmax_bits.p);
    const int dims[3][3] = {{g.I + 1, g.J, kl}, {g.I, g.J + 1, kl}, {g.I, g.J, kl + 1}};
    int lo, hi;
    work_layers(c, &lo, &hi);
    const bool whole = lo == 0 && hi == g.K;
    if (c->p2g_arith == GFS_EXACT) {
        GFS_REQUIRE(whole, "exact-arithmetic P2G is single-domain only");
        for (int comp = 0; comp < 3; comp++)
            LAUNCH(c, gfs::k_p2g_gather<1>, grid3(dims[comp][0], dims[comp][1], dims[comp][2], 64), 64, g, comp, sp, c->sources,
                   c->cell_start.p, c->soa[b][0].p, c->soa[b][1].p, c->soa[b][2].p, c->soa[b][3 + comp].p,
                   c->val[comp].p, c->setmask[comp].p);
    }
    // node layers [lo, hi) for u,v and [lo, hi] for w.  Finalize covers them all (it also re-zeroes the accumulators);
    // assemble only the owned layers' faces -- its 26-neighbourhood reads one finalized layer beyond them.
    long long plane[3] = {(long long)dims[0][0] * dims[0][1], (long long)dims[1][0] * dims[1][1], (long long)dims[2][0] * dims[2][1]};
    if (c->p2g_arith != GFS_EXACT && c->fused_grid) {
        gfs::FusedArgs fu;
        const int a_lo = whole ? 0 : c->own_k0, a_hi = whole ? g.K : c->own_k1;
        for (int comp = 0; comp < 3; comp++) { fu.acc[comp] = c->acc[comp].p; fu.out[comp] = c->field[GFS_FIELD_P2G][comp].p + gfs::kRowPad; }
        fu.k_lo = a_lo; fu.k_hi = a_hi; fu.k_hi_w = a_hi + (a_hi == g.K ? 1 : 0);
        LAUNCH(c, gfs::k_finalize_assemble, dim3((unsigned)ceil_div((long long)(g.I + 1) * (g.J + 1), 256), (unsigned)(fu.k_hi_w - a_lo)), 256,
               g, sp, c->sources, c->material.p, fu);
        c->acc_dirty = true;
        return;
    }
    if (c->p2g_arith != GFS_EXACT) {
        gfs::FinalizeArgs fa;
        long long total = 0;
        for (int comp = 0; comp < 3; comp++) {
            fa.acc[comp] = c->acc[comp].p; fa.val[comp] = c->val[comp].p; fa.setmask[comp] = c->setmask[comp].p;
            fa.first[comp] = plane[comp] * lo;
            fa.count[comp] = plane[comp] * (hi - lo + (comp == 2 ? 1 : 0));
            total += fa.count[comp];
        }
        LAUNCH(c, gfs::k_p2g_finalize, ceil_div(total, 256), 256, g, sp, c->sources, fa);
    }
    gfs::AssembleArgs aa;
    const int a_lo = whole ? 0 : c->own_k0, a_hi = whole ? g.K : c->own_k1;
    for (int comp = 0; comp < 3; comp++) {
        aa.val[comp] = c->val[comp].p; aa.setmask[comp] = c->setmask[comp].p; aa.out[comp] = c->field[GFS_FIELD_P2G][comp].p + gfs::kRowPad;
    }
    // the w face layer own_k1 is the upper slab's lower face: it belongs to the upper slab, except the top of the domain
    aa.k_lo = a_lo; aa.k_hi = a_hi; aa.k_hi_w = a_hi + (a_hi == g.K ? 1 : 0);
    LAUNCH(c, gfs::k_assemble, dim3((unsigned)ceil_div((long long)(g.I + 1) * (g.J + 1), 256), (unsigned)(aa.k_hi_w - a_lo)), 256, g, c->material.p, aa);
}

void do_p2g(gfs_context *c, int arith) {
    do_p2g_begin(c, arith);
    do_p2g_end(c);
}

bool g2p_uses_bricks(gfs_context *c, int arith) {
    return arith != GFS_EXACT && c->grid.pow2 && c->sorted && c->have_maps && c->g2p_variant >= 1;
}

void do_g2p(gfs_context *c, double dt, double ratio, int order, int interp, int arith, bool bin_next, const gfs::Migrate *migrate = nullptr) {
    require_domain(c);
    GFS_REQUIRE(order >= 1 && order <= 4, "RK order must be 1..4");
    GFS_REQUIRE(interp == GFS_TRILINEAR || interp == GFS_TRICUBIC, "bad interpolation mode");
    GFS_REQUIRE(c->velocities_valid, "the particle velocities are undefined after gfs_advect_substep: upload the particles again");
    if (c->dead > 0 && !(g2p_uses_bricks(c, arith) && c->indexed)) drop_dead(c);
    GFS_REQUIRE(!migrate || (g2p_uses_bricks(c, arith) && bin_next), "internal: fused migration needs the brick kernel");
    if (c->n == 0) return;
    const int src = c->cur, dst = 1 - c->cur;
    gfs::RkCoef rk = make_rk(dt);
    float rp = (float)ratio, rf = (float)(1 - ratio);      // fluidsimulation.cpp:3126
    GFS_CUDA(cudaMemsetAsync(c->counters.p + 2, 0, sizeof(unsigned long long), c->stream));
    // fast arithmetic: bin the advected positions for the next counting sort in the kernel's epilogue
    uint32_t *keys_out = nullptr, *rank_out = nullptr, *counts = nullptr;
    if (bin_next) {
        reset_counts(c);
        GFS_CUDA(cudaMemsetAsync(c->vmax_bits.p, 0, sizeof(unsigned int), c->stream));
        keys_out = c->keys[0].p; rank_out = c->rank.p; counts = c->counts.p;
    }
#define GFS_G2P_ARGS c->grid, field_ptrs(c, GFS_FIELD_NEW), field_ptrs(c, GFS_FIELD_SAVED), c->material.p, interp, order, rk, rp, rf, c->n, \
               c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,            \
               c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,            \
               c->counters.p, c->nkeys, keys_out, rank_out, counts, c->vmax_bits.p, coll
    const bool brick = g2p_uses_bricks(c, arith);
    gfs::CollList coll;
    coll.list = nullptr; coll.count = nullptr; coll.cap = 0;
    coll.cell_cap = migrate ? 0u : (unsigned int)c->cell_cap;          // (the cap is a single-domain rule so far)
    if (c->resolve_collisions) {
        // colliders are rare (a few per thousand at CFL 0.5 next to walls); a full list falls back to "keep p0"
        // (sized by do_sort at the start of the step: no allocation -- an implicit device synchronisation -- here,
        // between the device-side waits of a sharded substep)
        if (c->coll_list.cap == 0) c->coll_list.reserve(c->coll_cap_user > 0 ? (size_t)c->coll_cap_user : (size_t)(c->n / 16 + 4096));
        const size_t cap = c->coll_list.cap;
        c->coll_count.reserve(1);
        GFS_CUDA(cudaMemsetAsync(c->coll_count.p, 0, sizeof(unsigned int), c->stream));
        coll.list = c->coll_list.p; coll.count = c->coll_count.p; coll.cap = (unsigned int)cap;
    }
    gfs::Migrate mg;
    if (migrate) mg = *migrate;
    else { mg.own_lo = (int)0x80000000; mg.own_hi = 0x7FFFFFFF; mg.out[0] = mg.out[1] = nullptr; mg.count = nullptr; mg.cap = 0; }
    if (brick) {
        const int nbricks_all = (int)(c->nkeys / gfs::kBrickCells);          // the out-of-grid bin's CTA index
        const int nbricks = (int)(c->brick_hi - c->brick_lo);               // bricks around the owned layers (all, single domain)
        const bool ranged = !full_range(c);
#define GFS_BRICK_ARGS c->grid, c->maps[interp], field_ptrs(c, GFS_FIELD_NEW), field_ptrs(c, GFS_FIELD_SAVED), c->material.p, c->cell_start.p, \
               (c->indexed ? c->index.p : nullptr), c->tag[src].p, c->tag[dst].p, order, rk, rp, rf, c->n,                                                                                                  \
               c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,            \
               c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,            \
               c->counters.p, c->nkeys, keys_out, rank_out, counts, c->vmax_bits.p, mg, coll
        if (interp == GFS_TRILINEAR && c->g2p_variant >= 2) {
            // round-2 trilinear kernel: bricks only; the out-of-grid bin goes through k_g2p_brick (one CTA), particles whose
            // RK stages leave the staged block through k_g2p_slow
            const bool skew = c->g2p_variant == 3;
            gfs::SlowList slow;
            c->slow_count.reserve(1);
            slow.list = c->perm[0].p; slow.count = c->slow_count.p;
            GFS_CUDA(cudaMemsetAsync(c->slow_count.p, 0, sizeof(unsigned int), c->stream));
#define GFS_TRI_ARGS c->grid, (skew ? c->maps_tri_wide : c->maps[0]), c->material.p, c->cell_start.p, c->brick_lo, (c->indexed ? c->index.p : nullptr), \
               c->tag[src].p, c->tag[dst].p, order, rk, rp, rf,                                                                         \
               c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,            \
               c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,            \
               c->counters.p, c->nkeys, keys_out, rank_out, counts, c->vmax_bits.p, mg, coll, slow
            int prof_id_ = c->prof_begin(skew ? "gfs::k_g2p_tri<1>" : "gfs::k_g2p_tri<0>");
            if (skew) {
                if (migrate) gfs::k_g2p_tri<true, true><<<nbricks, 256, gfs::TriTile<true>::kSmemBytes, c->stream>>>(GFS_TRI_ARGS);
                else gfs::k_g2p_tri<true, false><<<nbricks, 256, gfs::TriTile<true>::kSmemBytes, c->stream>>>(GFS_TRI_ARGS);
            } else {
                if (migrate) gfs::k_g2p_tri<false, true><<<nbricks, 256, gfs::TriTile<false>::kSmemBytes, c->stream>>>(GFS_TRI_ARGS);
                else gfs::k_g2p_tri<false, false><<<nbricks, 256, gfs::TriTile<false>::kSmemBytes, c->stream>>>(GFS_TRI_ARGS);
            }
            c->prof_end(prof_id_);
            GFS_CUDA(cudaGetLastError());
#undef GFS_TRI_ARGS
            int prof_id2_ = c->prof_begin("gfs::k_g2p_brick<0> (out-of-grid bin)");
            if (migrate) gfs::k_g2p_brick<0, true><<<1, 256, gfs::BrickTile<0>::kSmemBytes, c->stream>>>(GFS_BRICK_ARGS, (uint32_t)nbricks_all);
            else gfs::k_g2p_brick<0, false><<<1, 256, gfs::BrickTile<0>::kSmemBytes, c->stream>>>(GFS_BRICK_ARGS, (uint32_t)nbricks_all);
            c->prof_end(prof_id2_);
            GFS_CUDA(cudaGetLastError());
#define GFS_SLOW_ARGS c->grid, field_ptrs(c, GFS_FIELD_NEW), field_ptrs(c, GFS_FIELD_SAVED), c->material.p, (c->indexed ? c->index.p : nullptr), \
               c->tag[src].p, c->tag[dst].p, interp, order, rk, rp, rf,                                                                 \
               c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,            \
               c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,            \
               c->counters.p, c->nkeys, keys_out, rank_out, counts, c->vmax_bits.p, mg, coll, slow
            if (migrate) LAUNCH(c, gfs::k_g2p_slow<true>, 296, 128, GFS_SLOW_ARGS);
            else LAUNCH(c, gfs::k_g2p_slow<false>, 296, 128, GFS_SLOW_ARGS);
#undef GFS_SLOW_ARGS
            c->launches += 2;
        } else {
            // single domain: every brick + the out-of-grid bin's CTA in one launch; z-slab rank: its brick range, then that CTA alone
            int prof_id_ = c->prof_begin(interp == GFS_TRICUBIC ? "gfs::k_g2p_brick<1>" : "gfs::k_g2p_brick<0>");
            for (int part = 0; part < (ranged ? 2 : 1); part++) {
                const int nb = ranged ? (part == 0 ? nbricks : 1) : nbricks + 1;
                const uint32_t b0 = ranged ? (part == 0 ? c->brick_lo : (uint32_t)nbricks_all) : 0u;
                if (nb <= 0) continue;
                if (interp == GFS_TRICUBIC) {
                    if (migrate) gfs::k_g2p_brick<1, true><<<nb, 256, gfs::BrickTile<1>::kSmemBytes, c->stream>>>(GFS_BRICK_ARGS, b0);
                    else gfs::k_g2p_brick<1, false><<<nb, 256, gfs::BrickTile<1>::kSmemBytes, c->stream>>>(GFS_BRICK_ARGS, b0);
                } else {
                    if (migrate) gfs::k_g2p_brick<0, true><<<nb, 256, gfs::BrickTile<0>::kSmemBytes, c->stream>>>(GFS_BRICK_ARGS, b0);
                    else gfs::k_g2p_brick<0, false><<<nb, 256, gfs::BrickTile<0>::kSmemBytes, c->stream>>>(GFS_BRICK_ARGS, b0);
                }
                c->launches++;
                GFS_CUDA(cudaGetLastError());
            }
            c->prof_end(prof_id_);
        }
#undef GFS_BRICK_ARGS
    }
    else if (arith == GFS_EXACT) LAUNCH(c, gfs::k_g2p_advect<1>, ceil_div(c->n, 256), 256, GFS_G2P_ARGS);
    else if (c->grid.pow2) LAUNCH(c, gfs::k_g2p_advect<2>, ceil_div(c->n, 256), 256, GFS_G2P_ARGS);
    else LAUNCH(c, gfs::k_g2p_advect<0>, ceil_div(c->n, 256), 256, GFS_G2P_ARGS);
#undef GFS_G2P_ARGS
    if (coll.list)
        LAUNCH(c, gfs::k_resolve_collisions, 64, 128, c->grid, c->material.p, coll, c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p,
               c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p, c->nkeys, keys_out, rank_out, counts, mg, key_range(c));
    // tags travel with the slot: the per-particle kernels keep slot order (copy the tag array across buffers); the brick
    // kernel moves the tags itself (it may be reading through the lazy sort index)
    if (!brick)
        GFS_CUDA(cudaMemcpyAsync(c->tag[dst].p, c->tag[src].p, sizeof(int32_t) * (size_t)c->n, cudaMemcpyDeviceToDevice, c->stream));
    if (brick && c->indexed) { c->n -= c->dead; c->dead = 0; }      // the kernel walked the index: only live slots were written
    if (bin_next && c->cell_cap > 0 && !migrate) {                  // particles past the per-cell cap went to the dead bin
        const int64_t d = read_dead_bin(c);
        c->dead += d; c->removed += d;
    }
    c->indexed = false;
    c->cur = dst;
    c->sorted = false; c->aos_valid = false;          // positions moved: the cell table no longer describes them
    c->keys_ready = bin_next;
}

// upload three host face arrays of a full (non-slab) field into scratch and return device pointers
gfs::FieldPtrs upload_field(gfs_context *c, const float *u, const float *v, const float *w, int I, int J, int K) {
    const size_t cnt[3] = {(size_t)(I + 1) * J * K, (size_t)I * (J + 1) * K, (size_t)I * J * (K + 1)};
    const float *h[3] = {u, v, w};
    gfs::FieldPtrs f;
    for (int a = 0; a < 3; a++) {
        GFS_REQUIRE(h[a] != nullptr, "null field pointer");
        c->h_field[a].reserve(cnt[a]);
        GFS_CUDA(cudaMemcpyAsync(c->h_field[a].p, h[a], cnt[a] * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        f.c[a] = c->h_field[a].p;
    }
    return f;
}

}  // namespace

// ---- TMA tensor maps over the resident (row-padded) field arrays ----------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    GFS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    GFS_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
    fn = (EncodeTiledFn)p;
    return fn;
}

// dense 3-D boxes {x 16, y, z} over the storage columns (trilinear bricks)
void make_field_map_dense(CUtensorMap *map, float *storage, int pitch, int nj, int nkl, int by, int bz, int bx = 16) {
    cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)nj, (cuuint64_t)nkl};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * (cuuint64_t)nj * 4};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, storage, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[160];
        snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed with CUresult %d (rows %d, planes %d, pitch %d)", (int)r, nj, nkl, pitch);
        throw GfsError(b);
    }
}

// 4-D view of a resident field for the brick boxes: (x_lo 8 floats, z, x_hi, y), see gfs::BrickTile
void make_field_map(CUtensorMap *map, float *storage, int pitch, int nj, int nkl, int by, int bz) {
    cuuint64_t dims[4] = {8, (cuuint64_t)nkl, (cuuint64_t)(pitch / 8), (cuuint64_t)nj};      // outside: zero fill
    cuuint64_t strides[3] = {(cuuint64_t)pitch * (cuuint64_t)nj * 4, 32, (cuuint64_t)pitch * 4};
    cuuint32_t box[4] = {8, (cuuint32_t)bz, 2, (cuuint32_t)by};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, storage, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[160];
        snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed with CUresult %d (rows %d, planes %d, pitch %d)", (int)r, nj, nkl, pitch);
        throw GfsError(b);
    }
}

void make_brick_maps(gfs_context *c) {
    const Grid &g = c->grid;
    const int kl = g.k1 - g.k0;
    const int nj[3] = {g.J, g.J + 1, g.J}, nk[3] = {kl, kl, kl + 1};
    for (int a = 0; a < 3; a++) {
        if (gfs::BrickTile<0>::kSkew) {
            make_field_map(&c->maps[0].m[a], c->field[GFS_FIELD_NEW][a].p, g.pitch[a], nj[a], nk[a], gfs::BrickTile<0>::nY, gfs::BrickTile<0>::nZ);
            make_field_map(&c->maps[0].m[3 + a], c->field[GFS_FIELD_SAVED][a].p, g.pitch[a], nj[a], nk[a], gfs::BrickTile<0>::sY, gfs::BrickTile<0>::sZ);
        } else {
            make_field_map_dense(&c->maps[0].m[a], c->field[GFS_FIELD_NEW][a].p, g.pitch[a], nj[a], nk[a], gfs::BrickTile<0>::nY, gfs::BrickTile<0>::nZ);
            make_field_map_dense(&c->maps[0].m[3 + a], c->field[GFS_FIELD_SAVED][a].p, g.pitch[a], nj[a], nk[a], gfs::BrickTile<0>::sY, gfs::BrickTile<0>::sZ);
        }
        make_field_map_dense(&c->maps_tri_wide.m[a], c->field[GFS_FIELD_NEW][a].p, g.pitch[a], nj[a], nk[a], gfs::TriTile<true>::nY, gfs::TriTile<true>::nZ, gfs::TriTile<true>::kX);
        make_field_map_dense(&c->maps_tri_wide.m[3 + a], c->field[GFS_FIELD_SAVED][a].p, g.pitch[a], nj[a], nk[a], gfs::TriTile<true>::sY, gfs::TriTile<true>::sZ, gfs::TriTile<true>::kX);
        make_field_map(&c->maps[1].m[a], c->field[GFS_FIELD_NEW][a].p, g.pitch[a], nj[a], nk[a], gfs::BrickTile<1>::nY, gfs::BrickTile<1>::nZ);
        make_field_map(&c->maps[1].m[3 + a], c->field[GFS_FIELD_SAVED][a].p, g.pitch[a], nj[a], nk[a], gfs::BrickTile<1>::sY, gfs::BrickTile<1>::sZ);
    }
    GFS_CUDA(cudaFuncSetAttribute(gfs::k_g2p_brick<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gfs::BrickTile<0>::kSmemBytes));
    GFS_CUDA(cudaFuncSetAttribute(gfs::k_g2p_brick<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gfs::BrickTile<1>::kSmemBytes));
    GFS_CUDA(cudaFuncSetAttribute(gfs::k_g2p_brick<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gfs::BrickTile<0>::kSmemBytes));
    GFS_CUDA(cudaFuncSetAttribute(gfs::k_g2p_brick<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gfs::BrickTile<1>::kSmemBytes));
    GFS_CUDA(cudaFuncSetAttribute(gfs::k_g2p_tri<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gfs::TriTile<false>::kSmemBytes));
    GFS_CUDA(cudaFuncSetAttribute(gfs::k_g2p_tri<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gfs::TriTile<false>::kSmemBytes));
    GFS_CUDA(cudaFuncSetAttribute(gfs::k_g2p_tri<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gfs::TriTile<true>::kSmemBytes));
    GFS_CUDA(cudaFuncSetAttribute(gfs::k_g2p_tri<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gfs::TriTile<true>::kSmemBytes));
    static_assert(gfs::TriTile<false>::nY == gfs::BrickTile<0>::nY && gfs::TriTile<false>::sY == gfs::BrickTile<0>::sY &&
                  gfs::TriTile<false>::nOrg == gfs::BrickTile<0>::nOrg && gfs::TriTile<false>::sOrg == gfs::BrickTile<0>::sOrg &&
                  !gfs::BrickTile<0>::kSkew, "the dense trilinear tile of k_g2p_tri uses the tensor maps of k_g2p_brick<0>");
    c->have_maps = true;
}

#define GFS_BEGIN                                                                                  \
    if (err) *err = GFS_SUCCESS;                                                                   \
    try {
#define GFS_END(retval)                                                                            \
    } catch (const std::exception &ex) {                                                           \
        set_error("%s", ex.what());                                                                \
        if (err) *err = GFS_FAIL;                                                                  \
        return retval;                                                                             \
    }

extern "C" {

const char *gfs_get_error_message(void) { return g_error; }

gfs_context *gfs_create(int device, void *stream, int *err) {
    GFS_BEGIN
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw GfsError(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                       "); this library has no CPU fallback");
    GFS_REQUIRE(device >= 0 && device < count, "device index out of range");
    GFS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GFS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (!(prop.major == 10 && prop.minor == 0)) {      // the library carries sm_100a SASS only (no PTX): sm_103 / sm_120 parts cannot run it
        char b[256];
        snprintf(b, sizeof(b), "device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; kernels are built for sm_100a only",
                 device, prop.name, prop.major, prop.minor);
        throw GfsError(b);
    }
    gfs_context *c = new gfs_context();
    c->device = device;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { GFS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    c->n_valid.reserve(1);
    c->vmax_bits.reserve(1);
    c->counters.reserve(4);
    GFS_CUDA(cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(unsigned long long), c->stream));
    c->sources.n = 0;
    return c;
    GFS_END(nullptr)
}

void gfs_destroy(gfs_context *c, int *err) {
    GFS_BEGIN
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int s = 0; s < 3; s++) for (int a = 0; a < 3; a++) c->field[s][a].release();
    for (int a = 0; a < 3; a++) { c->val[a].release(); c->setmask[a].release(); c->acc[a].release(); c->h_field[a].release(); }
    c->material.release(); c->cell_start.release(); c->counts.release(); c->rank.release(); c->index.release();
    for (int b = 0; b < 2; b++) { for (int a = 0; a < 6; a++) c->soa[b][a].release(); c->tag[b].release(); c->keys[b].release(); c->perm[b].release(); }
    for (int a = 0; a < 6; a++) c->press.vec[a].release();
    c->press.scal.release(); c->press.flags.release(); c->press.state.release(); c->press.order.release();
    c->press.ticket.release(); c->press.tile_done.release(); c->press.pressure.release(); c->press.trace.release();
    if (c->press.host_state) cudaFreeHost(c->press.host_state);
    if (c->press.host_resid) cudaFreeHost(c->press.host_resid);
    c->split_counters.release(); c->comm_error.release(); c->ext_layer.release(); c->coll_list.release(); c->coll_count.release(); c->h_mat.release(); c->h_layer.release();
    for (int sd = 0; sd < 2; sd++) {
        if (c->comm[sd].peer && c->comm[sd].peer_ipc) cudaIpcCloseMemHandle(c->comm[sd].peer);
        if (c->comm[sd].block) cudaFree(c->comm[sd].block);
    }
    for (int r = 0; r < 16; r++) if (c->world_peer[r] && c->world_peer_ipc[r]) cudaIpcCloseMemHandle(c->world_peer[r]);
    for (size_t i = 0; i < c->prof_spans.size(); i++) { cudaEventDestroy(c->prof_spans[i].a); cudaEventDestroy(c->prof_spans[i].b); }
    for (size_t i = 0; i < c->prof_pool.size(); i++) cudaEventDestroy(c->prof_pool[i]);
    if (c->comm_host) cudaFreeHost(c->comm_host);
    if (c->removal_host) cudaFreeHost(c->removal_host);
    if (c->world_table) cudaFree(c->world_table);
    for (int gi = 0; gi < 2; gi++) if (c->graphs[gi].exec) cudaGraphExecDestroy(c->graphs[gi].exec);
    c->cub_tmp.release(); c->n_valid.release(); c->vmax_bits.release(); c->counters.release();
    c->aos_stage.release(); c->h_pos.release(); c->h_out.release(); c->h_val.release(); c->h_fld.release(); c->h_wgt.release(); c->h_acc.release();
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    GFS_END()
}

void gfs_device_info(gfs_context *c, char *buf, int buflen, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && buf && buflen > 0, "bad arguments");
    cudaDeviceProp prop;
    GFS_CUDA(cudaGetDeviceProperties(&prop, c->device));
    snprintf(buf, (size_t)buflen, "CUDA device %d: %s, sm_%d%d, %d SMs, %.1f GB, L2 %.0f MB, %d KB smem/SM",
             c->device, prop.name, prop.major, prop.minor, prop.multiProcessorCount,
             (double)prop.totalGlobalMem / 1e9, (double)prop.l2CacheSize / 1048576.0,
             (int)(prop.sharedMemPerMultiprocessor / 1024));
    GFS_END()
}

void gfs_sync(gfs_context *c, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

void gfs_get_stats(gfs_context *c, gfs_stats_t *out, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && out, "bad arguments");
    unsigned long long h[4] = {0, 0, 0, 0};
    int32_t nv = 0;
    GFS_CUDA(cudaMemcpyAsync(h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    if (c->has_domain && c->sorted)
        GFS_CUDA(cudaMemcpyAsync(&nv, c->cell_start.p + c->nkeys, sizeof(nv), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    out->num_particles = c->n - c->dead;
    out->out_of_grid = c->sorted ? c->n - c->dead - nv : 0;
    out->in_solid = (int64_t)h[0];
    out->fluid_cells = (int64_t)h[1];
    out->solid_hits = (int64_t)h[2];
    out->kernel_launches = c->launches;
    out->graph_replays = c->graph_replays;
    out->removed_particles = c->removed;
    out->collision_overflow = (int64_t)h[3];
    GFS_END()
}

void gfs_profile_enable(gfs_context *c, int on, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    c->profiling = on != 0;
    GFS_END()
}

int gfs_profile_read(gfs_context *c, char *names, double *total_ms, int64_t *counts, int cap, int reset, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && names && total_ms && counts && cap > 0, "bad arguments");
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    int nn = (int)c->prof_names.size() < cap ? (int)c->prof_names.size() : cap;
    for (int i = 0; i < nn; i++) {
        snprintf(names + 64 * i, 64, "%s", c->prof_names[i].c_str());
        total_ms[i] = 0.0; counts[i] = 0;
    }
    for (size_t i = 0; i < c->prof_spans.size(); i++) {
        const ProfSpan &sp = c->prof_spans[i];
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.name < nn) { total_ms[sp.name] += ms; counts[sp.name]++; }
    }
    if (reset) {
        for (size_t i = 0; i < c->prof_spans.size(); i++) { c->prof_pool.push_back(c->prof_spans[i].a); c->prof_pool.push_back(c->prof_spans[i].b); }
        c->prof_spans.clear();
    }
    return nn;
    GFS_END(-1)
}

/* ---- host-pointer operators ------------------------------------------------------------------ */

void gfs_sample(gfs_context *c, const float *pos, int64_t n, const float *u, const float *v, const float *w,
                int I, int J, int K, double dx, int interp, int arith, int validate, float *out, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && (n == 0 || (pos && out)), "bad arguments");
    GFS_REQUIRE(I > 0 && J > 0 && K > 0 && dx > 0, "bad grid");
    GFS_REQUIRE(interp == GFS_TRILINEAR || interp == GFS_TRICUBIC, "bad interpolation mode");
    if (n == 0) return;
    GFS_CUDA(cudaSetDevice(c->device));
    Grid g = make_grid(I, J, K, dx, 0, K);
    gfs::FieldPtrs f = upload_field(c, u, v, w, I, J, K);
    c->h_pos.reserve((size_t)n * 3); c->h_out.reserve((size_t)n * 3);
    GFS_CUDA(cudaMemcpyAsync(c->h_pos.p, pos, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
    if (arith == GFS_EXACT) LAUNCH(c, gfs::k_sample<1>, ceil_div(n, 128), 128, g, f, interp, validate, n, c->h_pos.p, c->h_out.p);
    else if (g.pow2) LAUNCH(c, gfs::k_sample<2>, ceil_div(n, 128), 128, g, f, interp, validate, n, c->h_pos.p, c->h_out.p);   // fp32-exact index math
    else LAUNCH(c, gfs::k_sample<0>, ceil_div(n, 128), 128, g, f, interp, validate, n, c->h_pos.p, c->h_out.p);
    GFS_CUDA(cudaMemcpyAsync(out, c->h_out.p, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

void gfs_advect(gfs_context *c, const float *pos, int64_t n, const float *u, const float *v, const float *w,
                int I, int J, int K, double dx, double dt, int order, int interp, int arith, float *out, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && (n == 0 || (pos && out)), "bad arguments");
    GFS_REQUIRE(I > 0 && J > 0 && K > 0 && dx > 0, "bad grid");
    GFS_REQUIRE(order >= 1 && order <= 4, "RK order must be 1..4");
    GFS_REQUIRE(interp == GFS_TRILINEAR || interp == GFS_TRICUBIC, "bad interpolation mode");
    if (n == 0) return;
    GFS_CUDA(cudaSetDevice(c->device));
    Grid g = make_grid(I, J, K, dx, 0, K);
    gfs::FieldPtrs f = upload_field(c, u, v, w, I, J, K);
    c->h_pos.reserve((size_t)n * 3); c->h_out.reserve((size_t)n * 3);
    GFS_CUDA(cudaMemcpyAsync(c->h_pos.p, pos, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
    gfs::RkCoef rk = make_rk(dt);
    if (arith == GFS_EXACT) LAUNCH(c, gfs::k_advect<1>, ceil_div(n, 128), 128, g, f, interp, order, rk, n, c->h_pos.p, c->h_out.p);
    else if (g.pow2) LAUNCH(c, gfs::k_advect<2>, ceil_div(n, 128), 128, g, f, interp, order, rk, n, c->h_pos.p, c->h_out.p);
    else LAUNCH(c, gfs::k_advect<0>, ceil_div(n, 128), 128, g, f, interp, order, rk, n, c->h_pos.p, c->h_out.p);
    GFS_CUDA(cudaMemcpyAsync(out, c->h_out.p, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

/* MACVelocityField::extrapolateVelocityField (src/macvelocityfield.cpp:786-798) on caller-owned host arrays, in place:
 * the body a maintainer would put into that method (u, v, w = getRawArrayU/V/W(), material = one byte per cell). */
void gfs_extrapolate_field(gfs_context *c, float *u, float *v, float *w, int I, int J, int K, const uint8_t *material,
                           int num_layers, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && u && v && w && material, "bad arguments");
    GFS_REQUIRE(I > 0 && J > 0 && K > 0, "bad grid");
    GFS_REQUIRE(num_layers >= 0 && num_layers <= 126, "layer count must be 0..126");
    GFS_CUDA(cudaSetDevice(c->device));
    Grid g = make_grid(I, J, K, 1.0, 0, K);                 // unpadded rows; dx plays no role in the extrapolation
    gfs::FieldPtrs fp = upload_field(c, u, v, w, I, J, K);
    gfs::FieldRW f;
    for (int a = 0; a < 3; a++) f.c[a] = const_cast<float *>(fp.c[a]);
    const size_t cells = (size_t)I * J * K;
    c->h_mat.reserve(cells);
    c->h_layer.reserve(cells);
    GFS_CUDA(cudaMemcpyAsync(c->h_mat.p, material, cells, cudaMemcpyHostToDevice, c->stream));
    const unsigned nodes = (unsigned)ceil_div((long long)(I + 1) * (J + 1), 256), percell = (unsigned)ceil_div((long long)I * J, 256);
    LAUNCH(c, gfs::k_extrapolate_reset, dim3(nodes, (unsigned)K + 1), 256, g, c->h_mat.p, c->h_layer.p, f);
    for (int L = 1; L <= num_layers; L++)
        LAUNCH(c, gfs::k_extrapolate_mark, dim3(percell, (unsigned)K), 256, g, c->h_mat.p, c->h_layer.p, L);
    for (int L = 1; L <= num_layers; L++)
        LAUNCH(c, gfs::k_extrapolate_faces, dim3(nodes, (unsigned)K + 1), 256, g, c->h_mat.p, c->h_layer.p, f, L);
    const size_t cnt[3] = {(size_t)(I + 1) * J * K, (size_t)I * (J + 1) * K, (size_t)I * J * (K + 1)};
    float *h[3] = {u, v, w};
    for (int a = 0; a < 3; a++)
        GFS_CUDA(cudaMemcpyAsync(h[a], f.c[a], cnt[a] * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

namespace {
void splat_points_host(gfs_context *c, const float *pos, const float *values, int64_t n, double radius,
                       const float *offset3, double dx, int ni, int nj, int nk, float *field, float *weight,
                       int accumulate, int arith, int use_threshold, float threshold) {
    GFS_REQUIRE(c && field && offset3 && (n == 0 || pos), "bad arguments");
    GFS_REQUIRE(ni > 0 && nj > 0 && nk > 0 && dx > 0 && radius > 0, "bad grid");
    GFS_CUDA(cudaSetDevice(c->device));
    const size_t count = (size_t)ni * nj * nk;
    c->h_acc.reserve(2 * count);
    c->h_fld.reserve(count); c->h_wgt.reserve(count);
    GFS_CUDA(cudaMemsetAsync(c->h_acc.p, 0, 2 * count * sizeof(unsigned long long), c->stream));
    // numerator scale from the largest |value| (order-independent, so the result stays order-independent)
    float vmax = values ? 0.0f : 1.0f;                    // values == NULL: every point carries 1 (ScalarField::addPoint)
    if (values) for (int64_t i = 0; i < n; i++) { float a = std::fabs(values[i]); if (a < 3.0e38f && a > vmax) vmax = a; }
    int vexp = 0;
    if (vmax > 0.0f) { int e; std::frexp(vmax, &e); vexp = e; }      // vmax < 2^e
    vexp = vexp < -24 ? -24 : (vexp > 40 ? 40 : vexp);
    gfs::SplatParams sp = make_splat(radius, nullptr);
    if (n > 0) {
        c->h_pos.reserve((size_t)n * 3); c->h_val.reserve((size_t)n);
        GFS_CUDA(cudaMemcpyAsync(c->h_pos.p, pos, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
        if (values) GFS_CUDA(cudaMemcpyAsync(c->h_val.p, values, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        if (arith == GFS_EXACT)
            LAUNCH(c, gfs::k_splat_points<1>, ceil_div(n, 128), 128, sp, vexp, dx, offset3[0], offset3[1], offset3[2], ni, nj, nk, n,
                   c->h_pos.p, values ? c->h_val.p : nullptr, c->h_acc.p);
        else
            LAUNCH(c, gfs::k_splat_points<0>, ceil_div(n, 128), 128, sp, vexp, dx, offset3[0], offset3[1], offset3[2], ni, nj, nk, n,
                   c->h_pos.p, values ? c->h_val.p : nullptr, c->h_acc.p);
    }
    if (accumulate) {
        GFS_CUDA(cudaMemcpyAsync(c->h_fld.p, field, count * 4, cudaMemcpyHostToDevice, c->stream));
        if (weight) GFS_CUDA(cudaMemcpyAsync(c->h_wgt.p, weight, count * 4, cudaMemcpyHostToDevice, c->stream));
    }
    LAUNCH(c, gfs::k_splat_points_store, ceil_div((int64_t)count, 256), 256, (int64_t)count, vexp, c->h_acc.p, c->h_fld.p,
           weight ? c->h_wgt.p : nullptr, accumulate, use_threshold, threshold);
    GFS_CUDA(cudaMemcpyAsync(field, c->h_fld.p, count * 4, cudaMemcpyDeviceToHost, c->stream));
    if (weight) GFS_CUDA(cudaMemcpyAsync(weight, c->h_wgt.p, count * 4, cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
}
}  // namespace

void gfs_add_point_values(gfs_context *c, const float *pos, const float *values, int64_t n, double radius,
                          const float *offset3, double dx, int ni, int nj, int nk, float *field, float *weight,
                          int accumulate, int arith, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(n == 0 || values, "null values");
    splat_points_host(c, pos, values, n, radius, offset3, dx, ni, nj, nk, field, weight, accumulate, arith, 0, 0.0f);
    GFS_END()
}

/* CLScalarField::addPoints (src/clscalarfield.cpp:62-145): every point carries the value 1.  use_threshold: the caller
 * set a max-scalar-field-value threshold (IsotropicParticleMesher, src/isotropicparticlemesher.cpp:334-359).  The
 * reference honours it three different ways -- ScalarField::addPoint skips a node whose RUNNING value exceeds it
 * (src/scalarfield.cpp:182-184, order dependent), the OpenCL path skips whole 8^3 chunks whose minimum already does
 * (src/clscalarfield.cpp:1002-1010), and the class's own CPU path ignores it (:1490-1505) -- all of which leave the
 * iso-surface level (0.5 < threshold 1.0) where it is.  Here: a node whose value BEFORE this batch already exceeds the
 * threshold receives nothing from the batch (the per-node form of the chunk rule; order independent). */
void gfs_add_points(gfs_context *c, const float *pos, int64_t n, double radius, const float *offset3, double dx,
                    int ni, int nj, int nk, float *field, int accumulate, int use_threshold, float threshold, int arith, int *err) {
    GFS_BEGIN
    splat_points_host(c, pos, nullptr, n, radius, offset3, dx, ni, nj, nk, field, nullptr, accumulate, arith,
                      (use_threshold && accumulate) ? 1 : 0, threshold);
    GFS_END()
}

/* ---- device-resident domain ------------------------------------------------------------------- */

void gfs_domain_init(gfs_context *c, int I, int J, int K, double dx, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_REQUIRE(I > 0 && J > 0 && K > 0 && dx > 0, "bad grid");
    GFS_CUDA(cudaSetDevice(c->device));
    if (c->has_domain && c->dead > 0) drop_dead(c);          // dead slots are only defined by the OLD domain's keys
    c->graph_epoch++;
    c->grid = make_grid(I, J, K, dx, 0, K, true);
    const Grid &g = c->grid;
    const int kl = g.k1 - g.k0;
    c->field_floats[0] = (size_t)g.pitch[0] * J * kl;
    c->field_floats[1] = (size_t)g.pitch[1] * (J + 1) * kl;
    c->field_floats[2] = (size_t)g.pitch[2] * J * (kl + 1);
    GFS_REQUIRE((uint64_t)g.nbi * g.nbj * g.nbk * gfs::kBrickCells < 0x7FFFFFFFull, "grid too large for 31-bit cell keys");
    c->nkeys = (uint32_t)((uint64_t)g.nbi * g.nbj * g.nbk * gfs::kBrickCells);
    c->face_count[0] = (size_t)(I + 1) * J * kl;
    c->face_count[1] = (size_t)I * (J + 1) * kl;
    c->face_count[2] = (size_t)I * J * (kl + 1);
    c->cell_count = (size_t)I * J * kl;
    for (int s = 0; s < 3; s++)
        for (int a = 0; a < 3; a++) {
            c->field[s][a].reserve(c->field_floats[a]);
            GFS_CUDA(cudaMemsetAsync(c->field[s][a].p, 0, c->field_floats[a] * sizeof(float), c->stream));
        }
    for (int a = 0; a < 3; a++) {
        c->val[a].reserve(c->face_count[a]);
        c->setmask[a].reserve(c->face_count[a]);
        c->acc[a].reserve(2 * c->face_count[a]);
        GFS_CUDA(cudaMemsetAsync(c->acc[a].p, 0, 2 * c->face_count[a] * sizeof(unsigned long long), c->stream));
    }
    c->material.reserve(c->cell_count);
    c->cell_start.reserve((size_t)c->nkeys + 3);
    c->counts.reserve((size_t)c->nkeys + 3);
    // a new domain invalidates every key, index and cell table of the old one
    c->keys_ready = false; c->indexed = false; c->storage_sorted = false; c->sorted = false; c->aos_valid = false;
    c->has_domain = true;
    c->press.valid = false;
    c->have_maps = false;
    c->own_k0 = 0; c->own_k1 = K;
    set_key_range(c);
    if (g.pow2) make_brick_maps(c);
    LAUNCH(c, gfs::k_border_solid, grid3(I, J, kl), 128, g, c->material.p);
    GFS_END()
}

void gfs_set_material(gfs_context *c, const uint8_t *material, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(material, "null pointer");
    if (c->remove_in_solid) {
        // option 6 is applied by the binning pass that reads the positions (k_hist): keys binned by a G2P epilogue
        // against the OLD material must not be reused, or particles inside newly added solids would survive
        if (c->dead > 0) drop_dead(c);
        c->keys_ready = false;
    }
    GFS_CUDA(cudaMemcpyAsync(c->material.p, material, c->cell_count, cudaMemcpyHostToDevice, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

/* FluidSimulation::_fluidCellIndices (src/fluidsimulation.cpp:2019-2039): the fluid cells of the resident material grid in
 * the reference's k, j, i scan order, compacted on the device.  cells receives at most `capacity` triples; *count the number
 * of fluid cells (call with capacity 0 to size the buffer).  Single domain. */
void gfs_get_fluid_cells(gfs_context *c, gfs_grid_index_t *cells, int64_t capacity, int64_t *count, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(count && capacity >= 0 && (capacity == 0 || cells), "bad arguments");
    const Grid &g = c->grid;
    GFS_REQUIRE(c->own_k0 == 0 && c->own_k1 == g.K, "gfs_get_fluid_cells is single-domain only");
    GFS_CUDA(cudaSetDevice(c->device));
    const long long n = (long long)c->cell_count;
    DevBuf<uint32_t> flags, offs;
    flags.reserve((size_t)n + 1); offs.reserve((size_t)n + 1);
    GFS_CUDA(cudaMemsetAsync(flags.p + n, 0, sizeof(uint32_t), c->stream));
    LAUNCH(c, gfs::k_fluid_flags, ceil_div(n, 256), 256, c->material.p, n, flags.p);
    size_t tmp = 0;
    GFS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, flags.p, offs.p, (int)(n + 1), c->stream));
    c->cub_tmp.reserve(tmp);
    GFS_CUDA(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, flags.p, offs.p, (int)(n + 1), c->stream));
    uint32_t total = 0;
    GFS_CUDA(cudaMemcpyAsync(&total, offs.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    *count = (int64_t)total;
    const int64_t m = std::min<int64_t>(capacity, (int64_t)total);
    if (m > 0) {
        DevBuf<int> out;
        out.reserve((size_t)m * 3);
        LAUNCH(c, gfs::k_fluid_cells, ceil_div(n, 256), 256, c->material.p, offs.p, n, g.I, g.J, (long long)m, out.p);
        GFS_CUDA(cudaMemcpyAsync(cells, out.p, (size_t)m * 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GFS_CUDA(cudaStreamSynchronize(c->stream));
        out.release();
    }
    flags.release(); offs.release();
    GFS_END()
}

void gfs_get_material(gfs_context *c, uint8_t *material, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(material, "null pointer");
    GFS_CUDA(cudaMemcpyAsync(material, c->material.p, c->cell_count, cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

void gfs_set_sources(gfs_context *c, const gfs_source_t *sources, int nsources, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_REQUIRE(nsources >= 0 && nsources <= 8, "at most 8 inflow sources");
    GFS_REQUIRE(nsources == 0 || sources, "null pointer");
    c->sources.n = nsources;
    for (int i = 0; i < nsources; i++) c->sources.s[i] = sources[i];
    c->graph_epoch++;
    GFS_END()
}

void gfs_set_particles(gfs_context *c, const gfs_marker_particle_t *particles, int64_t n, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && n >= 0 && (n == 0 || particles), "bad arguments");
    GFS_REQUIRE(n < 0x7FFFFFFFll, "particle count must fit int32");
    GFS_CUDA(cudaSetDevice(c->device));
    c->reserve_particles(n > 0 ? n : 1);
    c->n = n; c->dead = 0; c->cur = 0; c->sorted = false; c->aos_valid = false; c->keys_ready = false; c->indexed = false; c->storage_sorted = false;
    c->velocities_valid = true;
    if (c->allmax_posted) c->allmax_redo = true;
    c->graph_epoch++;
    if (n > 0) {
        c->aos_stage.reserve((size_t)n * 6);
        GFS_CUDA(cudaMemcpyAsync(c->aos_stage.p, particles, (size_t)n * 24, cudaMemcpyHostToDevice, c->stream));
        LAUNCH(c, gfs::k_aos_to_soa, ceil_div(n, 256), 256, n, c->aos_stage.p, c->soa[0][0].p, c->soa[0][1].p, c->soa[0][2].p,
               c->soa[0][3].p, c->soa[0][4].p, c->soa[0][5].p, c->tag[0].p);
    }
    c->aos_valid = n > 0;
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

int64_t gfs_num_particles(gfs_context *c, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    return c->n - c->dead;
    GFS_END(-1)
}

void gfs_get_particles(gfs_context *c, gfs_marker_particle_t *particles, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && (c->n == 0 || particles), "bad arguments");
    GFS_CUDA(cudaSetDevice(c->device));
    drop_dead(c);
    if (c->n == 0) return;
    const int b = c->cur;
    c->h_pos.reserve((size_t)c->n * 6);
    LAUNCH(c, gfs::k_soa_to_aos, ceil_div(c->n, 256), 256, c->n, c->soa[b][0].p, c->soa[b][1].p, c->soa[b][2].p,
           c->soa[b][3].p, c->soa[b][4].p, c->soa[b][5].p, c->h_pos.p);
    GFS_CUDA(cudaMemcpyAsync(particles, c->h_pos.p, (size_t)c->n * 24, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long overflow = 0;
    GFS_CUDA(cudaMemcpyAsync(&overflow, c->counters.p + 3, sizeof(overflow), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    if (overflow) {         // a silent parity break otherwise: those particles kept p0 instead of the reference's resolved position
        GFS_CUDA(cudaMemsetAsync(c->counters.p + 3, 0, sizeof(unsigned long long), c->stream));
        char b[256];
        snprintf(b, sizeof(b), "%llu particles advected into solid cells did not fit the collision list (capacity %zu) and kept their old "
                 "position; raise it with gfs_set_option(ctx, 8, capacity)", overflow, c->coll_list.cap);
        throw GfsError(b);
    }
    GFS_END()
}

void gfs_get_particle_order(gfs_context *c, int32_t *order, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && (c->n == 0 || order), "bad arguments");
    GFS_CUDA(cudaSetDevice(c->device));
    drop_dead(c);
    if (c->n == 0) return;
    GFS_CUDA(cudaMemcpyAsync(order, c->tag[c->cur].p, (size_t)c->n * 4, cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

void gfs_set_field(gfs_context *c, int slot, const float *u, const float *v, const float *w, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(slot >= 0 && slot < 3 && u && v && w, "bad arguments");
    const float *h[3] = {u, v, w};
    const int ni[3] = {c->grid.I + 1, c->grid.I, c->grid.I};
    // reference rows (ni floats) -> resident rows (pitch floats): one contiguous copy over PCIe into scratch, then a
    // strided copy on the device (a strided host-to-device copy ran at 18 GB/s against 55 GB/s for the contiguous one)
    for (int a = 0; a < 3; a++) {
        c->h_field[a].reserve(c->face_count[a]);
        GFS_CUDA(cudaMemcpyAsync(c->h_field[a].p, h[a], c->face_count[a] * 4, cudaMemcpyHostToDevice, c->stream));
        GFS_CUDA(cudaMemcpy2DAsync(c->field[slot][a].p + gfs::kRowPad, (size_t)c->grid.pitch[a] * 4, c->h_field[a].p, (size_t)ni[a] * 4, (size_t)ni[a] * 4,
                                   c->face_count[a] / (size_t)ni[a], cudaMemcpyDeviceToDevice, c->stream));
    }
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

void gfs_get_field(gfs_context *c, int slot, float *u, float *v, float *w, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(slot >= 0 && slot < 3 && u && v && w, "bad arguments");
    float *h[3] = {u, v, w};
    const int ni[3] = {c->grid.I + 1, c->grid.I, c->grid.I};
    for (int a = 0; a < 3; a++)
        GFS_CUDA(cudaMemcpy2DAsync(h[a], (size_t)ni[a] * 4, c->field[slot][a].p + gfs::kRowPad, (size_t)c->grid.pitch[a] * 4, (size_t)ni[a] * 4,
                                   c->face_count[a] / (size_t)ni[a], cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

/* Layer-range forms of gfs_set_field / gfs_get_field / gfs_get_material for z-slab ranks: u, v, w (material) are the
 * caller's WHOLE arrays in the reference's layout; only the cell layers [k_first, k_first + k_count) travel (the w array
 * has one more face layer: it is moved too).  A rank uploads its owned layers plus the halo and downloads what it owns. */
void gfs_set_field_layers(gfs_context *c, int slot, const float *u, const float *v, const float *w, int k_first, int k_count, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(slot >= 0 && slot < 3 && u && v && w, "bad arguments");
    const Grid &g = c->grid;
    GFS_REQUIRE(k_first >= 0 && k_count >= 0 && k_first + k_count <= g.K, "layer range out of bounds");
    const float *h[3] = {u, v, w};
    const int ni[3] = {g.I + 1, g.I, g.I}, nj[3] = {g.J, g.J + 1, g.J};
    for (int a = 0; a < 3; a++) {
        const size_t rows = (size_t)nj[a] * (size_t)(k_count + (a == 2 ? 1 : 0)), row0 = (size_t)nj[a] * (size_t)k_first;
        if (rows == 0) continue;
        c->h_field[a].reserve(c->face_count[a]);              // contiguous over PCIe, strided on the device (see gfs_set_field)
        GFS_CUDA(cudaMemcpyAsync(c->h_field[a].p, h[a] + row0 * ni[a], rows * ni[a] * 4, cudaMemcpyHostToDevice, c->stream));
        GFS_CUDA(cudaMemcpy2DAsync(c->field[slot][a].p + gfs::kRowPad + row0 * g.pitch[a], (size_t)g.pitch[a] * 4, c->h_field[a].p,
                                   (size_t)ni[a] * 4, (size_t)ni[a] * 4, rows, cudaMemcpyDeviceToDevice, c->stream));
    }
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

void gfs_get_field_layers(gfs_context *c, int slot, float *u, float *v, float *w, int k_first, int k_count, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(slot >= 0 && slot < 3 && u && v && w, "bad arguments");
    const Grid &g = c->grid;
    GFS_REQUIRE(k_first >= 0 && k_count >= 0 && k_first + k_count <= g.K, "layer range out of bounds");
    float *h[3] = {u, v, w};
    const int ni[3] = {g.I + 1, g.I, g.I}, nj[3] = {g.J, g.J + 1, g.J};
    for (int a = 0; a < 3; a++) {
        const size_t rows = (size_t)nj[a] * (size_t)(k_count + (a == 2 ? 1 : 0)), row0 = (size_t)nj[a] * (size_t)k_first;
        if (rows == 0) continue;
        GFS_CUDA(cudaMemcpy2DAsync(h[a] + row0 * ni[a], (size_t)ni[a] * 4, c->field[slot][a].p + gfs::kRowPad + row0 * g.pitch[a],
                                   (size_t)g.pitch[a] * 4, (size_t)ni[a] * 4, rows, cudaMemcpyDeviceToHost, c->stream));
    }
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

void gfs_get_material_layers(gfs_context *c, uint8_t *material, int k_first, int k_count, int *err) {
    GFS_BEGIN
    require_domain(c);
    const Grid &g = c->grid;
    GFS_REQUIRE(material && k_first >= 0 && k_count >= 0 && k_first + k_count <= g.K, "bad arguments");
    const size_t plane = (size_t)g.I * g.J;
    GFS_CUDA(cudaMemcpyAsync(material + plane * k_first, c->material.p + plane * k_first, plane * k_count, cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

/* MACVelocityField::extrapolateVelocityField(materialGrid, numLayers) (src/macvelocityfield.cpp:786-798) on a resident
 * field, with the resident material grid (what FluidSimulation runs on the saved field after P2G and on
 * the solved field before G2P, src/fluidsimulation.cpp:3067-3070, 3306-3307, 3334). */
void gfs_extrapolate(gfs_context *c, int slot, int num_layers, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(slot >= 0 && slot < 3, "bad field slot");
    GFS_REQUIRE(num_layers >= 0 && num_layers <= 126, "layer count must be 0..126");
    const Grid &g = c->grid;
    GFS_REQUIRE(c->own_k0 == 0 && c->own_k1 == g.K, "gfs_extrapolate is single-domain only (no slab exchange of the layer grid yet)");
    GFS_CUDA(cudaSetDevice(c->device));
    c->ext_layer.reserve(c->cell_count);
    gfs::FieldRW f;
    for (int a = 0; a < 3; a++) f.c[a] = c->field[slot][a].p + gfs::kRowPad;
    const unsigned nodes = (unsigned)ceil_div((long long)(g.I + 1) * (g.J + 1), 256), cells = (unsigned)ceil_div((long long)g.I * g.J, 256);
    LAUNCH(c, gfs::k_extrapolate_reset, dim3(nodes, (unsigned)g.K + 1), 256, g, c->material.p, c->ext_layer.p, f);
    for (int L = 1; L <= num_layers; L++)
        LAUNCH(c, gfs::k_extrapolate_mark, dim3(cells, (unsigned)g.K), 256, g, c->material.p, c->ext_layer.p, L);
    for (int L = 1; L <= num_layers; L++)
        LAUNCH(c, gfs::k_extrapolate_faces, dim3(nodes, (unsigned)g.K + 1), 256, g, c->material.p, c->ext_layer.p, f, L);
    GFS_END()
}

/* dst field slot := src field slot (e.g. "_savedVelocityField = _MACVelocity", src/fluidsimulation.cpp:3306) */
void gfs_copy_field(gfs_context *c, int dst_slot, int src_slot, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(dst_slot >= 0 && dst_slot < 3 && src_slot >= 0 && src_slot < 3, "bad field slot");
    GFS_CUDA(cudaSetDevice(c->device));
    if (dst_slot != src_slot)
        for (int a = 0; a < 3; a++)
            GFS_CUDA(cudaMemcpyAsync(c->field[dst_slot][a].p, c->field[src_slot][a].p, c->field_floats[a] * sizeof(float),
                                     cudaMemcpyDeviceToDevice, c->stream));
    GFS_END()
}

/* ---- stages 6-8 on the resident grid (kernels and the parity argument: gfs_pressure.cuh) -------- */

/* FluidSimulation::_applyConstantBodyForces (src/fluidsimulation.cpp:2765-2805): every face of `slot` bordering a fluid
 * cell gets (float)(force * dt) added; a component whose force is exactly zero is skipped, as the reference does. */
void gfs_apply_body_force(gfs_context *c, int slot, float fx, float fy, float fz, double dt, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(slot >= 0 && slot < 3, "bad field slot");
    const Grid &g = c->grid;
    GFS_REQUIRE(c->own_k0 == 0 && c->own_k1 == g.K, "gfs_apply_body_force is single-domain only");
    GFS_CUDA(cudaSetDevice(c->device));
    const float force[3] = {fx, fy, fz};
    gfs::Float3 add;
    int mask = 0;
    for (int a = 0; a < 3; a++) {
        add.v[a] = (float)(force[a] * dt);                     /* bodyForce.x * dt: float * double, narrowed by addU */
        if (std::fabs(force[a]) > 0.0) mask |= 1 << a;
    }
    if (mask) {
        gfs::FieldRW f;
        for (int a = 0; a < 3; a++) f.c[a] = c->field[slot][a].p + gfs::kRowPad;
        const unsigned nodes = (unsigned)ceil_div((long long)(g.I + 1) * (g.J + 1), 256);
        LAUNCH(c, gfs::k_body_force, dim3(nodes, (unsigned)g.K + 1), 256, g, c->material.p, f, add, mask);
        c->graph_epoch++;
    }
    GFS_END()
}

namespace {
gfs::PressSys pressure_system(gfs_context *c, const Grid &g, const uint8_t *material, double dt, double density, double tolerance) {
    gfs_context::Pressure &P = c->press;
    const size_t cells = (size_t)g.I * g.J * g.K;
    for (int a = 0; a < 6; a++) P.vec[a].reserve(cells);
    P.scal.reserve(3 * gfs::kPressBlocks + 4);
    P.flags.reserve(cells); P.state.reserve(4); P.ticket.reserve(4); P.pressure.reserve(cells);
    if (!P.host_state) GFS_CUDA(cudaMallocHost((void **)&P.host_state, 4 * sizeof(int)));
    if (!P.host_resid) GFS_CUDA(cudaMallocHost((void **)&P.host_resid, sizeof(double)));
    gfs::PressSys S;
    S.I = g.I; S.J = g.J; S.K = g.K;
    S.ntx = ceil_div(g.I, gfs::kTileX); S.nty = ceil_div(g.J, gfs::kTileY); S.ntz = ceil_div(g.K, gfs::kTileZ);
    S.ntiles = S.ntx * S.nty * S.ntz;
    if (P.dims[0] != g.I || P.dims[1] != g.J || P.dims[2] != g.K) {
        /* tiles in wavefront order: a tile's three predecessors always come earlier */
        std::vector<int> order((size_t)S.ntiles), start((size_t)(S.ntx + S.nty + S.ntz) + 1, 0);
        auto stage = [&](int t) { return t % S.ntx + (t / S.ntx) % S.nty + t / (S.ntx * S.nty); };
        for (int t = 0; t < S.ntiles; t++) start[(size_t)stage(t) + 1]++;
        for (size_t q = 1; q < start.size(); q++) start[q] += start[q - 1];
        for (int t = 0; t < S.ntiles; t++) order[(size_t)start[(size_t)stage(t)]++] = t;
        P.order.reserve((size_t)S.ntiles); P.tile_done.reserve((size_t)S.ntiles);
        GFS_CUDA(cudaMemcpyAsync(P.order.p, order.data(), (size_t)S.ntiles * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        GFS_CUDA(cudaStreamSynchronize(c->stream));                       /* `order` is a local */
        GFS_CUDA(cudaMemsetAsync(P.tile_done.p, 0, (size_t)S.ntiles * sizeof(unsigned int), c->stream));
        P.epoch = 0;
        P.dims[0] = g.I; P.dims[1] = g.J; P.dims[2] = g.K;
    }
    S.material = material; S.flags = P.flags.p;
    S.r = P.vec[0].p; S.z = P.vec[1].p; S.s = P.vec[2].p; S.p = P.vec[3].p; S.q = P.vec[4].p; S.precon = P.vec[5].p;
    S.partial = P.scal.p; S.sigma = P.scal.p + 3 * gfs::kPressBlocks; S.resid = P.scal.p + 3 * gfs::kPressBlocks + 2;
    S.state = P.state.p; S.ticket = P.ticket.p; S.tile_done = P.tile_done.p; S.order = P.order.p; S.pressure = P.pressure.p;
    S.cells = (long long)cells;
    S.scale = dt / (density * g.dx * g.dx);
    S.tol = tolerance;
    S.trace = nullptr;
    if (P.trace_on) { P.trace.reserve((size_t)S.ntiles * 8); GFS_CUDA(cudaMemsetAsync(P.trace.p, 0, (size_t)S.ntiles * 8 * sizeof(long long), c->stream)); S.trace = P.trace.p; }
    return S;
}

/* the solve itself, on any field / material the caller has on the device (g gives the rows' pitch) */
gfs::PressSys pressure_solve_device(gfs_context *c, const Grid &g, gfs::FieldPtrs f, const uint8_t *material, double dt, double density,
                                    double tolerance, int max_iterations, int *iterations, double *residual) {
    gfs_context::Pressure &P = c->press;
    gfs::PressSys S = pressure_system(c, g, material, dt, density, tolerance);
    const size_t cells = (size_t)S.cells;
    for (int a = 1; a < 6; a++) GFS_CUDA(cudaMemsetAsync(P.vec[a].p, 0, cells * sizeof(double), c->stream));
    GFS_CUDA(cudaMemsetAsync(P.state.p, 0, 4 * sizeof(int), c->stream));
    GFS_CUDA(cudaMemsetAsync(P.ticket.p, 0, 4 * sizeof(unsigned long long), c->stream));
    int sms = 148;
    GFS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const int sweep_blocks = 2 * sms;
    const int B = gfs::kPressBlocks, T = gfs::kPressThreads;

    const int sentinel = c->press_variant >= 2 ? 1 : 0;
    LAUNCH(c, gfs::k_press_setup, B, T, g, f, S, g.dx, sentinel);
    LAUNCH(c, gfs::k_press_check, 1, T, S, -1);
    const int subst_blocks = 5 * sms, subst_threads = gfs::kSubstWarps * 32;
    auto substitutions = [&]() {                              /* _applyPreconditioner: z = M^-1 r */
        if (c->press_variant == 0) {
            LAUNCH(c, gfs::k_press_sweep<1>, sweep_blocks, T, S, ++P.epoch);
            LAUNCH(c, gfs::k_press_sweep<2>, sweep_blocks, T, S, ++P.epoch);
        } else if (c->press_variant == 1) {
            LAUNCH(c, gfs::k_press_subst<false>, subst_blocks, subst_threads, S, ++P.epoch);
            LAUNCH(c, gfs::k_press_subst<true>, subst_blocks, subst_threads, S, ++P.epoch);
        } else if (c->press_variant == 2) {
            LAUNCH(c, (gfs::k_press_subst_df<false, false>), subst_blocks, subst_threads, S);
            LAUNCH(c, (gfs::k_press_subst_df<true, false>), subst_blocks, subst_threads, S);
        } else {
            LAUNCH(c, (gfs::k_press_subst_df<false, true>), subst_blocks, subst_threads, S);
            LAUNCH(c, (gfs::k_press_subst_df<true, true>), subst_blocks, subst_threads, S);
        }
    };
    LAUNCH(c, gfs::k_press_sweep<0>, sweep_blocks, T, S, ++P.epoch);
    substitutions();
    LAUNCH(c, gfs::k_press_dot_zr, B, T, S);
    LAUNCH(c, gfs::k_press_search, B, T, S, 0, 1);
    int it = 0;
    bool done = false;
    auto poll = [&]() {
        GFS_CUDA(cudaMemcpyAsync(P.host_state, P.state.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GFS_CUDA(cudaMemcpyAsync(P.host_resid, S.resid, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        GFS_CUDA(cudaStreamSynchronize(c->stream));
        done = P.host_state[0] != 0;
    };
    while (it < max_iterations && !done) {
        LAUNCH(c, gfs::k_press_apply_matrix, B, T, S);
        LAUNCH(c, gfs::k_press_update, B, T, S, it, sentinel);
        LAUNCH(c, gfs::k_press_check, 1, T, S, it);
        substitutions();
        LAUNCH(c, gfs::k_press_dot_zr, B, T, S);
        LAUNCH(c, gfs::k_press_search, B, T, S, it, 0);
        it++;
        if (it % 8 == 0) poll();
    }
    LAUNCH(c, gfs::k_press_finish, B, T, S);
    poll();
    if (iterations) *iterations = done ? P.host_state[1] : max_iterations;
    if (residual) *residual = *P.host_resid;
    return S;
}
}  // namespace

/* PressureSolver::solve (src/pressuresolver.cpp:116-139, 452-505) on the velocity field in `slot` and the resident material
 * grid, narrowed to the float grid of FluidSimulation::_updatePressureGrid (src/fluidsimulation.cpp:2870-2889).
 * tolerance / max_iterations: the reference's 1e-6 / 200 (src/pressuresolver.h:159-160), density its 20.0
 * (src/fluidsimulation.h:1154).  *iterations = the reference's iterationNumber when it returns (-1: the right-hand side
 * was already below the tolerance, pressure = 0; max_iterations: limit reached, the estimate so far is kept, as the
 * reference does).  *residual = the last max-norm of the residual.  The pressure stays on the device for
 * gfs_apply_pressure / gfs_get_pressure. */
void gfs_pressure_solve(gfs_context *c, int slot, double dt, double density, double tolerance, int max_iterations,
                        int *iterations, double *residual, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(slot >= 0 && slot < 3, "bad field slot");
    GFS_REQUIRE(dt > 0 && density > 0 && tolerance > 0 && max_iterations >= 0, "bad solver parameters");
    const Grid &g = c->grid;
    GFS_REQUIRE(c->own_k0 == 0 && c->own_k1 == g.K, "gfs_pressure_solve is single-domain only");
    GFS_CUDA(cudaSetDevice(c->device));
    c->press.valid = false;
    pressure_solve_device(c, g, field_ptrs(c, slot), c->material.p, dt, density, tolerance, max_iterations, iterations, residual);
    c->press.valid = true;
    GFS_END()
}

/* The same solve on caller-owned host arrays: the body of PressureSolver::solve for a simulator that keeps its grids on
 * the host (dropin/pressuresolver.cpp).  u, v, w: the reference's raw Array3d<float> storage; material: one byte per
 * cell; pressure: isize*jsize*ksize DOUBLES, i fastest, the solver's own precision (0 outside fluid cells) -- the caller
 * picks the fluid cells' entries in its own order. */
void gfs_pressure_solve_field(gfs_context *c, const float *u, const float *v, const float *w, int I, int J, int K, double dx,
                              const uint8_t *material, double dt, double density, double tolerance, int max_iterations,
                              double *pressure, int *iterations, double *residual, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && u && v && w && material && pressure, "bad arguments");
    GFS_REQUIRE(I > 0 && J > 0 && K > 0 && dx > 0, "bad grid");
    GFS_REQUIRE(dt > 0 && density > 0 && tolerance > 0 && max_iterations >= 0, "bad solver parameters");
    GFS_CUDA(cudaSetDevice(c->device));
    Grid g = make_grid(I, J, K, dx, 0, K);                  // unpadded rows
    gfs::FieldPtrs fp = upload_field(c, u, v, w, I, J, K);
    const size_t cells = (size_t)I * J * K;
    c->h_mat.reserve(cells);
    GFS_CUDA(cudaMemcpyAsync(c->h_mat.p, material, cells, cudaMemcpyHostToDevice, c->stream));
    c->press.valid = false;                                 // the resident float grid no longer belongs to the resident domain
    gfs::PressSys S = pressure_solve_device(c, g, fp, c->h_mat.p, dt, density, tolerance, max_iterations, iterations, residual);
    GFS_CUDA(cudaMemcpyAsync(pressure, S.p, cells * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

/* FluidSimulation::_applyPressureToVelocityField (src/fluidsimulation.cpp:2895-3061) with the pressure of the last
 * gfs_pressure_solve: dst_slot := projected src_slot (the two may be the same slot). */
void gfs_apply_pressure(gfs_context *c, int src_slot, int dst_slot, double dt, double density, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(src_slot >= 0 && src_slot < 3 && dst_slot >= 0 && dst_slot < 3, "bad field slot");
    GFS_REQUIRE(dt > 0 && density > 0, "bad parameters");
    GFS_REQUIRE(c->press.valid, "gfs_pressure_solve has not been called on this domain");
    const Grid &g = c->grid;
    GFS_CUDA(cudaSetDevice(c->device));
    gfs::FieldRW dst;
    for (int a = 0; a < 3; a++) dst.c[a] = c->field[dst_slot][a].p + gfs::kRowPad;
    const unsigned nodes = (unsigned)ceil_div((long long)(g.I + 1) * (g.J + 1), 256);
    LAUNCH(c, gfs::k_apply_pressure, dim3(nodes, (unsigned)g.K + 1), 256, g, c->material.p, field_ptrs(c, src_slot), dst,
           c->press.pressure.p, dt / (density * g.dx));
    c->graph_epoch++;
    GFS_END()
}

/* the float pressure grid of the last solve, one value per cell (i fastest), 0 outside fluid cells */
void gfs_get_pressure(gfs_context *c, float *pressure, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(pressure, "null output");
    GFS_REQUIRE(c->press.valid, "gfs_pressure_solve has not been called on this domain");
    GFS_CUDA(cudaSetDevice(c->device));
    GFS_CUDA(cudaMemcpyAsync(pressure, c->press.pressure.p, (size_t)c->grid.I * c->grid.J * c->grid.K * sizeof(float),
                             cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

void gfs_sort(gfs_context *c, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_CUDA(cudaSetDevice(c->device));
    do_sort(c, true);
    GFS_END()
}

void gfs_sort_unstable(gfs_context *c, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_CUDA(cudaSetDevice(c->device));
    do_sort(c, false);
    GFS_END()
}

/* Counting sort that leaves the particles where they are and materialises only the sorted index (P2G and G2P fetch
 * through it; G2P stores its results in sorted order, so the storage stays nearly sorted from step to step).  Falls
 * back to gfs_sort_unstable where the brick kernels do not apply. */
void gfs_sort_index(gfs_context *c, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_CUDA(cudaSetDevice(c->device));
    do_sort(c, false, /*lazy=*/c->has_domain && c->grid.pow2 && c->have_maps && c->g2p_variant >= 1 && c->p2g_variant >= 1 && c->lazy_sort);
    GFS_END()
}

void gfs_set_option(gfs_context *c, int option, int value, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    if (option == 0) { GFS_REQUIRE(value >= 0 && value <= 3, "p2g variant must be 0..3"); c->p2g_variant = value; }
    else if (option == 1) { GFS_REQUIRE(value >= 0 && value <= 3, "g2p variant must be 0..3"); c->g2p_variant = value; }
    else if (option == 2) { GFS_REQUIRE(value == 0 || value == 1, "lazy sort must be 0 or 1"); c->lazy_sort = value; }
    else if (option == 3) { GFS_REQUIRE(value == 0 || value == 1, "collision resolve must be 0 or 1"); c->resolve_collisions = value; }
    else if (option == 4) { GFS_REQUIRE(value == 0 || value == 1, "graph replay must be 0 or 1"); c->use_graphs = value; }
    else if (option == 5) { GFS_REQUIRE(value >= 0 && value < (1 << 20), "cell cap must be >= 0"); c->cell_cap = value; }
    else if (option == 6) { GFS_REQUIRE(value == 0 || value == 1, "solid-cell removal must be 0 or 1"); c->remove_in_solid = value; }
    else if (option == 7) { GFS_REQUIRE(value >= 1 && value <= 3600, "peer-exchange wait limit must be 1..3600 seconds"); c->comm_timeout_cycles = 2000000000ll * value; }
    else if (option == 13) { GFS_REQUIRE(value == 0 || value == 1, "pressure trace must be 0 or 1"); c->press.trace_on = value; }
    else if (option == 12) { GFS_REQUIRE(value >= 0 && value <= 3, "pressure sweep variant must be 0..3"); c->press_variant = value; }
    else if (option == 11) { GFS_REQUIRE(value == 0 || value == 1, "fused grid pass must be 0 or 1"); c->fused_grid = value; }
    else if (option == 10) { GFS_REQUIRE(value == 0 || value == 1, "split wait must be 0 or 1"); c->split_wait = value; }
    else if (option == 9) { GFS_REQUIRE(value == 0 || value == 1, "early all-ranks max must be 0 or 1"); c->allmax_early = value; }
    else if (option == 8) { GFS_REQUIRE(value >= 0, "collision list capacity must be >= 0"); c->coll_cap_user = value; c->coll_list.release(); }
    else throw GfsError("gfs_set_option: unknown option");
    c->graph_epoch++;
    GFS_END()
}

void gfs_p2g(gfs_context *c, int arith, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_CUDA(cudaSetDevice(c->device));
    do_p2g(c, arith);
    GFS_END()
}

void gfs_g2p_advect(gfs_context *c, double dt, double ratio, int order, int interp, int arith, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_CUDA(cudaSetDevice(c->device));
    do_g2p(c, dt, ratio, order, interp, arith, false);
    GFS_END()
}

namespace {
void substep_body(gfs_context *c, double dt, double ratio, int order, int interp, int arith) {
    // exact arithmetic needs the stable order; fast arithmetic is order-independent and uses the counting sort,
    // binned for the following substep by the G2P kernel's epilogue
    if (c->keys_ready && arith == GFS_EXACT) c->keys_ready = false;
    do_sort(c, arith == GFS_EXACT, /*lazy=*/arith != GFS_EXACT && c->grid.pow2 && c->have_maps && c->g2p_variant >= 1 && c->lazy_sort);
    do_p2g(c, arith);
    do_g2p(c, dt, ratio, order, interp, arith, arith != GFS_EXACT);
}

// steady state of the fused fast substep: everything the launch sequence depends on is either baked into the graph key
// or guarded by graph_epoch, and the host-side state transition of one substep is fixed (the buffer parity flips)
bool substep_graph_eligible(gfs_context *c, int arith) {
    return c->use_graphs && !c->profiling && arith != GFS_EXACT && c->has_domain && c->grid.pow2 && c->have_maps &&
           c->g2p_variant >= 1 && c->p2g_variant >= 1 && c->lazy_sort && c->storage_sorted && c->keys_ready && !c->sorted &&
           !c->indexed && c->dead == 0 && c->n > 0 && c->own_k0 == 0 && c->own_k1 == c->grid.K && !removal_on(c);
}
}  // namespace

void gfs_substep(gfs_context *c, double dt, double ratio, int order, int interp, int arith, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_REQUIRE(order >= 1 && order <= 4, "RK order must be 1..4");
    GFS_REQUIRE(interp == GFS_TRILINEAR || interp == GFS_TRICUBIC, "bad interpolation mode");
    GFS_CUDA(cudaSetDevice(c->device));
    if (!substep_graph_eligible(c, arith)) { substep_body(c, dt, ratio, order, interp, arith); return; }
    gfs_context::SubstepGraph &g = c->graphs[c->cur];
    const void *scratch[3] = {c->cub_tmp.p, c->coll_list.p, c->coll_count.p};
    const bool fresh = g.exec && g.epoch == c->graph_epoch && g.n == c->n && g.dt == dt && g.ratio == ratio && g.order == order &&
                       g.interp == interp && g.scratch[0] == scratch[0] && g.scratch[1] == scratch[1] && g.scratch[2] == scratch[2];
    if (fresh) {
        // replay; then the host-side state transition the captured calls made: sort (index) -> p2g -> g2p (parity flips)
        GFS_CUDA(cudaGraphLaunch(g.exec, c->stream));
        c->p2g_arith = arith;
        c->cur = 1 - c->cur;
        c->sorted = false; c->aos_valid = false; c->indexed = false; c->keys_ready = true;
        c->launches += g.launches;
        c->graph_replays++;
        return;
    }
    // scratch buffers are sized lazily by the first steps: capture only once they exist (no allocation inside a capture)
    if (!scratch[0] || (c->resolve_collisions && (!scratch[1] || !scratch[2]))) { substep_body(c, dt, ratio, order, interp, arith); return; }
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    const int64_t launches0 = c->launches;
    cudaGraph_t graph = nullptr;
    GFS_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    try {
        substep_body(c, dt, ratio, order, interp, arith);
    } catch (...) {
        cudaStreamEndCapture(c->stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
    }
    GFS_CUDA(cudaStreamEndCapture(c->stream, &graph));
    cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { g.exec = nullptr; throw GfsError(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
    g.epoch = c->graph_epoch; g.n = c->n; g.dt = dt; g.ratio = ratio; g.order = order; g.interp = interp;
    g.launches = c->launches - launches0;
    for (int i = 0; i < 3; i++) g.scratch[i] = scratch[i];
    GFS_CUDA(cudaGraphLaunch(g.exec, c->stream));          // the capture recorded the step, this performs it
    GFS_END()
}

/* Advection only on the resident particles (SURVEY 8d sub-metric "C5": ParticleAdvector::advectParticlesRK1..4,
 * src/particleadvector.cpp:209-399, on the device-resident, cell-sorted set): index sort + RK `order` through field slot
 * NEW with the trilinear brick kernel (TMA-staged tiles, positions only: 12 B read + 12 B written per particle), binned for
 * the next call by the kernel's epilogue.  No solid test (that is FluidSimulation's, :3198-3208), no velocity update: the
 * velocity arrays are left undefined.  dx must be a power of two. */
void gfs_advect_substep(gfs_context *c, double dt, int order, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(order >= 1 && order <= 4, "RK order must be 1..4");
    GFS_REQUIRE(c->grid.pow2 && c->have_maps, "gfs_advect_substep needs a power-of-two dx (brick kernels)");
    GFS_CUDA(cudaSetDevice(c->device));
    do_sort(c, false, /*lazy=*/c->lazy_sort != 0);
    if (c->n == 0) return;
    const int src = c->cur, dst = 1 - c->cur;
    gfs::RkCoef rk = make_rk(dt);
    reset_counts(c);
    gfs::CollList coll; coll.list = nullptr; coll.count = nullptr; coll.cap = 0; coll.cell_cap = 0;
    gfs::Migrate mg; mg.own_lo = (int)0x80000000; mg.own_hi = 0x7FFFFFFF; mg.out[0] = mg.out[1] = nullptr; mg.count = nullptr; mg.cap = 0;
    gfs::SlowList slow;
    c->slow_count.reserve(1);
    slow.list = c->perm[0].p; slow.count = c->slow_count.p;
    GFS_CUDA(cudaMemsetAsync(c->slow_count.p, 0, sizeof(unsigned int), c->stream));
    const int nbricks = (int)(c->brick_hi - c->brick_lo), nbricks_all = (int)(c->nkeys / gfs::kBrickCells);
    const int32_t *idx = c->indexed ? c->index.p : nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        GFS_CUDA(cudaFuncSetAttribute(gfs::k_g2p_tri<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gfs::TriTile<false>::kSmemBytes));
        attr_set = true;
    }
    int prof_id = c->prof_begin("gfs::k_g2p_tri<0> (advect only)");
    gfs::k_g2p_tri<false, false, true><<<nbricks, 256, gfs::TriTile<false>::kSmemBytes, c->stream>>>(
        c->grid, c->maps[0], (const uint8_t *)nullptr, c->cell_start.p, c->brick_lo, idx, c->tag[src].p, c->tag[dst].p, order, rk, 0.0f, 0.0f,
        c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,
        c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,
        c->counters.p, c->nkeys, c->keys[0].p, c->rank.p, c->counts.p, c->vmax_bits.p, mg, coll, slow);
    c->prof_end(prof_id);
    GFS_CUDA(cudaGetLastError());
    // the out-of-grid bin (global loads; its velocity output is as undefined as everyone's) and the stage-leavers
    gfs::k_g2p_brick<0, false><<<1, 256, gfs::BrickTile<0>::kSmemBytes, c->stream>>>(
        c->grid, c->maps[0], field_ptrs(c, GFS_FIELD_NEW), field_ptrs(c, GFS_FIELD_NEW), (const uint8_t *)nullptr, c->cell_start.p, idx,
        c->tag[src].p, c->tag[dst].p, order, rk, 0.0f, 0.0f, c->n,
        c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,
        c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,
        c->counters.p, c->nkeys, c->keys[0].p, c->rank.p, c->counts.p, c->vmax_bits.p, mg, coll, (uint32_t)nbricks_all);
    GFS_CUDA(cudaGetLastError());
    LAUNCH(c, (gfs::k_g2p_slow<false, true>), 296, 128, c->grid, field_ptrs(c, GFS_FIELD_NEW), field_ptrs(c, GFS_FIELD_NEW), (const uint8_t *)nullptr, idx,
           c->tag[src].p, c->tag[dst].p, GFS_TRILINEAR, order, rk, 0.0f, 0.0f,
           c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p,
           c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p,
           c->counters.p, c->nkeys, c->keys[0].p, c->rank.p, c->counts.p, c->vmax_bits.p, mg, coll, slow);
    c->launches += 2;
    if (c->indexed) { c->n -= c->dead; c->dead = 0; }
    c->indexed = false;
    c->cur = dst;
    c->sorted = false; c->aos_valid = false;
    c->keys_ready = true;
    c->velocities_valid = false;
    c->graph_epoch++;
    GFS_END()
}

void gfs_p2g_begin(gfs_context *c, int arith, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_CUDA(cudaSetDevice(c->device));
    do_p2g_begin(c, arith);
    GFS_END()
}

void gfs_p2g_end(gfs_context *c, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c, "null context");
    GFS_CUDA(cudaSetDevice(c->device));
    do_p2g_end(c);
    GFS_END()
}

void gfs_set_owned_layers(gfs_context *c, int k0, int k1, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(k0 >= 0 && k1 > k0 && k1 <= c->grid.K, "bad layer range");
    c->own_k0 = k0; c->own_k1 = k1;
    set_key_range(c);
    c->graph_epoch++;
    GFS_END()
}

namespace {
// (pointer, bytes per z-layer, number of z-layers) of a resident grid array; what: 0..2 NEW u,v,w; 3..5 SAVED; 6..8 P2G;
// 9 material; 10..12 accumulators of u,v,w (two 64-bit integers per node)
void layer_info(gfs_context *c, int what, unsigned char **base, size_t *bytes, int *layers) {
    const Grid &g = c->grid;
    const int kl = g.k1 - g.k0;
    const int nj[3] = {g.J, g.J + 1, g.J}, ni[3] = {g.I + 1, g.I, g.I};
    if (what >= 0 && what < 9) {
        int a = what % 3;
        *base = (unsigned char *)c->field[what / 3][a].p; *bytes = (size_t)g.pitch[a] * nj[a] * 4; *layers = kl + (a == 2);
    } else if (what == 9) {
        *base = (unsigned char *)c->material.p; *bytes = (size_t)g.I * g.J; *layers = kl;
    } else if (what >= 10 && what < 13) {
        int a = what - 10;
        *base = (unsigned char *)c->acc[a].p; *bytes = (size_t)ni[a] * nj[a] * 16; *layers = kl + (a == 2);
    } else throw GfsError("unknown grid array id");
}
}  // namespace

int64_t gfs_layer_bytes(gfs_context *c, int what, int *err) {
    GFS_BEGIN
    require_domain(c);
    unsigned char *base; size_t bytes; int layers;
    layer_info(c, what, &base, &bytes, &layers);
    return (int64_t)bytes;
    GFS_END(-1)
}

void gfs_pack_layers(gfs_context *c, int what, int k_first, int k_count, void *dst_device, int *err) {
    GFS_BEGIN
    require_domain(c);
    unsigned char *base; size_t bytes; int layers;
    layer_info(c, what, &base, &bytes, &layers);
    GFS_REQUIRE(dst_device && k_first >= 0 && k_count >= 0 && k_first + k_count <= layers, "layer range out of bounds");
    GFS_CUDA(cudaMemcpyAsync(dst_device, base + bytes * (size_t)k_first, bytes * (size_t)k_count, cudaMemcpyDeviceToDevice, c->stream));
    GFS_END()
}

void gfs_unpack_layers(gfs_context *c, int what, int k_first, int k_count, const void *src_device, int add, int *err) {
    GFS_BEGIN
    require_domain(c);
    unsigned char *base; size_t bytes; int layers;
    layer_info(c, what, &base, &bytes, &layers);
    GFS_REQUIRE(src_device && k_first >= 0 && k_count >= 0 && k_first + k_count <= layers, "layer range out of bounds");
    if (add) {
        GFS_REQUIRE(what >= 10 && what < 13, "only the integer accumulators can be added");
        long long count = (long long)(bytes * (size_t)k_count / 8);
        if (count > 0)
            LAUNCH(c, gfs::k_add_u64, ceil_div(count, 256), 256, count, (unsigned long long *)(base + bytes * (size_t)k_first),
                   (const unsigned long long *)src_device);
    } else {
        GFS_CUDA(cudaMemcpyAsync(base + bytes * (size_t)k_first, src_device, bytes * (size_t)k_count, cudaMemcpyDeviceToDevice, c->stream));
    }
    GFS_END()
}

// split kernel launch shared by the synchronous and the asynchronous entry points
static void launch_split(gfs_context *c, int k_lo, int k_hi, void *down_device, void *up_device, int64_t cap, unsigned int *counters) {
    GFS_CUDA(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned int), c->stream));
    const int src = c->cur, dst = 1 - c->cur;
    LAUNCH(c, gfs::k_split_by_layer, ceil_div(c->n, 256), 256, c->grid, c->n, k_lo, k_hi, (int)cap,
           c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p, c->tag[src].p,
           c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p, c->tag[dst].p,
           (float *)down_device, (float *)up_device, counters);
}

/* Batched form of gfs_pack_layers / gfs_unpack_layers: n <= 16 layer ranges to / from one caller-owned device buffer
 * at the given byte offsets, in ONE kernel launch.  direction 0 = pack (library -> buffer), 1 = unpack (buffer ->
 * library; add[i] != 0 adds 64-bit integers, accumulators only). */
void gfs_copy_layers_batch(gfs_context *c, int direction, int n, const int *what, const int *k_first, const int *k_count,
                           const int64_t *offsets, const int *add, void *buffer_device, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(n >= 0 && n <= 16 && (n == 0 || (what && k_first && k_count && offsets && buffer_device)), "bad arguments");
    GFS_REQUIRE(direction == 0 || direction == 1, "bad direction");
    GFS_CUDA(cudaSetDevice(c->device));
    gfs::CopyBatch cb;
    cb.n = 0;
    long long largest = 0;
    for (int i = 0; i < n; i++) {
        unsigned char *base; size_t bytes; int layers;
        layer_info(c, what[i], &base, &bytes, &layers);
        GFS_REQUIRE(k_first[i] >= 0 && k_count[i] >= 0 && k_first[i] + k_count[i] <= layers, "layer range out of bounds");
        if (k_count[i] == 0) continue;
        const bool adding = direction == 1 && add && add[i];
        GFS_REQUIRE(!adding || (what[i] >= 10 && what[i] < 13), "only the integer accumulators can be added");
        unsigned char *lib = base + bytes * (size_t)k_first[i], *buf = (unsigned char *)buffer_device + offsets[i];
        cb.src[cb.n] = direction == 0 ? lib : buf;
        cb.dst[cb.n] = direction == 0 ? buf : lib;
        cb.bytes[cb.n] = (long long)(bytes * (size_t)k_count[i]);
        cb.add[cb.n] = adding ? 1 : 0;
        if (cb.bytes[cb.n] > largest) largest = cb.bytes[cb.n];
        cb.n++;
    }
    if (cb.n == 0) return;
    int bx = (int)((largest / 16 + 255) / 256);
    if (bx < 1) bx = 1;
    if (bx > 592) bx = 592;                       // 4 waves of 148 SMs; the loops are grid-strided
    LAUNCH(c, gfs::k_copy_batch, dim3((unsigned)bx, (unsigned)cb.n), 256, cb);
    GFS_END()
}

static void launch_split(gfs_context *c, int k_lo, int k_hi, void *down_device, void *up_device, int64_t cap, unsigned int *counters);

/* ---- peer-memory exchange over CUDA IPC (same node; NVLink / NVSwitch) -------------------------------------------
 * Each context owns one "comm block" per side; the neighbour on that side maps it (gfs_comm_export / gfs_comm_connect)
 * and WRITES into it directly from its kernels: packed layers, migrating particles and the flag words that announce
 * them.  No NCCL call, no host round trip except the single count read of gfs_comm_migrate_finish. */
namespace {
constexpr size_t kCommFlagBytes = 256;       // uint32 words: [0] layers seq, [1] particles seq, [2] particle count
size_t comm_block_bytes(gfs_context *c) { return kCommFlagBytes + 2 * c->comm_layer_bytes + 2 * (size_t)c->comm_particle_cap * 24; }
unsigned char *comm_layers(gfs_context *c, unsigned char *block, unsigned int seq) { return block + kCommFlagBytes + (seq & 1) * c->comm_layer_bytes; }
unsigned char *comm_arrivals(gfs_context *c, unsigned char *block, unsigned int seq) {
    return block + kCommFlagBytes + 2 * c->comm_layer_bytes + (seq & 1) * (size_t)c->comm_particle_cap * 24;
}
void fill_batch(gfs_context *c, gfs::CopyBatch &cb, int direction, int n, const int *what, const int *k_first, const int *k_count,
                const int64_t *offsets, const int *add, unsigned char *buffer, long long *largest) {
    cb.n = 0;
    *largest = 0;
    for (int i = 0; i < n; i++) {
        unsigned char *base; size_t bytes; int layers;
        layer_info(c, what[i], &base, &bytes, &layers);
        GFS_REQUIRE(k_first[i] >= 0 && k_count[i] >= 0 && k_first[i] + k_count[i] <= layers, "layer range out of bounds");
        if (k_count[i] == 0) continue;
        const bool adding = direction == 1 && add && add[i];
        GFS_REQUIRE(!adding || (what[i] >= 10 && what[i] < 13), "only the integer accumulators can be added");
        GFS_REQUIRE((size_t)offsets[i] + bytes * (size_t)k_count[i] <= c->comm_layer_bytes, "comm buffer too small");
        unsigned char *lib = base + bytes * (size_t)k_first[i], *buf = buffer + offsets[i];
        cb.src[cb.n] = direction == 0 ? lib : buf;
        cb.dst[cb.n] = direction == 0 ? buf : lib;
        cb.bytes[cb.n] = (long long)(bytes * (size_t)k_count[i]);
        cb.add[cb.n] = adding ? 1 : 0;
        if (cb.bytes[cb.n] > *largest) *largest = cb.bytes[cb.n];
        cb.n++;
    }
}
int batch_blocks(long long largest) {
    long long bx = (largest / 16 + 255) / 256;
    return (int)(bx < 1 ? 1 : (bx > 592 ? 592 : bx));
}
}  // namespace

void gfs_comm_alloc(gfs_context *c, int64_t layer_bytes, int64_t particle_cap, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(layer_bytes >= 0 && particle_cap >= 0 && particle_cap < 0x7FFFFFFFll, "bad arguments");
    GFS_CUDA(cudaSetDevice(c->device));
    c->comm_layer_bytes = ((size_t)layer_bytes + 255) / 256 * 256;
    c->comm_particle_cap = particle_cap;
    for (int s = 0; s < 2; s++) {
        if (c->comm[s].block) GFS_CUDA(cudaFree(c->comm[s].block));
        GFS_CUDA(cudaMalloc((void **)&c->comm[s].block, comm_block_bytes(c)));
        GFS_CUDA(cudaMemset(c->comm[s].block, 0, kCommFlagBytes));
        c->comm[s].seq_layers = c->comm[s].seq_particles = 0;
        if (c->comm[s].peer && c->comm[s].peer_ipc) cudaIpcCloseMemHandle(c->comm[s].peer);
        c->comm[s].peer = nullptr; c->comm[s].peer_ipc = false;
    }
    if (!c->comm_host) GFS_CUDA(cudaHostAlloc((void **)&c->comm_host, 64, cudaHostAllocMapped));
    c->comm_error.reserve(1);
    GFS_CUDA(cudaMemset(c->comm_error.p, 0, sizeof(unsigned int)));
    c->split_counters.reserve(8);
    GFS_END()
}

void gfs_comm_export(gfs_context *c, int side, void *handle64, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && (side == 0 || side == 1) && handle64 && c->comm[side].block, "gfs_comm_alloc first");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    GFS_CUDA(cudaIpcGetMemHandle(&h, c->comm[side].block));
    memcpy(handle64, &h, 64);
    GFS_END()
}

/* `side`: where the neighbour sits; handle64: ITS block for the opposite side (the one I write into) */
void gfs_comm_connect(gfs_context *c, int side, const void *handle64, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && (side == 0 || side == 1) && handle64, "bad arguments");
    GFS_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    GFS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    if (c->comm[side].peer && c->comm[side].peer_ipc) cudaIpcCloseMemHandle(c->comm[side].peer);
    c->comm[side].peer = (unsigned char *)p;
    c->comm[side].peer_ipc = true;
    GFS_END()
}

/* in-process variant for several contexts of one process (tests): connect to another context's block directly */
void gfs_comm_connect_local(gfs_context *c, int side, gfs_context *neighbour, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && neighbour && (side == 0 || side == 1) && neighbour->comm[1 - side].block, "bad arguments");
    GFS_REQUIRE(neighbour->comm_layer_bytes == c->comm_layer_bytes && neighbour->comm_particle_cap == c->comm_particle_cap,
                "both ends must use the same comm sizes");
    c->comm[side].peer = neighbour->comm[1 - side].block;
    c->comm[side].peer_ipc = false;
    GFS_END()
}

/* pack the layer ranges straight into the neighbour's buffer and raise its flag */
void gfs_comm_push_layers(gfs_context *c, int side, int n, const int *what, const int *k_first, const int *k_count,
                          const int64_t *offsets, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE((side == 0 || side == 1) && c->comm[side].peer, "gfs_comm_connect first");
    GFS_REQUIRE(n >= 0 && n <= 16, "at most 16 ranges");
    GFS_CUDA(cudaSetDevice(c->device));
    const unsigned int seq = ++c->comm[side].seq_layers;
    gfs::CopyBatch cb;
    long long largest;
    fill_batch(c, cb, 0, n, what, k_first, k_count, offsets, nullptr, comm_layers(c, c->comm[side].peer, seq), &largest);
    if (cb.n > 0) LAUNCH(c, gfs::k_copy_batch, dim3((unsigned)batch_blocks(largest), (unsigned)cb.n), 256, cb);
    LAUNCH(c, gfs::k_signal, 1, 1, (volatile unsigned int *)c->comm[side].peer, seq, nullptr, nullptr, 0);
    GFS_END()
}

/* wait for the neighbour's layers (device-side spin on my flag) and unpack / add them */
void gfs_comm_pull_layers(gfs_context *c, int side, int n, const int *what, const int *k_first, const int *k_count,
                          const int64_t *offsets, const int *add, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE((side == 0 || side == 1) && c->comm[side].block && c->comm[side].peer, "gfs_comm_connect first");
    GFS_REQUIRE(n >= 1 && n <= 16, "1..16 ranges");
    GFS_CUDA(cudaSetDevice(c->device));
    const unsigned int seq = c->comm[side].seq_layers;          // the push of this exchange already advanced it
    gfs::CopyBatch cb;
    long long largest;
    fill_batch(c, cb, 1, n, what, k_first, k_count, offsets, add, comm_layers(c, c->comm[side].block, seq), &largest);
    GFS_REQUIRE(cb.n > 0, "nothing to unpack");
    if (c->split_wait) {
        LAUNCH(c, gfs::k_wait_flag, 1, 1, (const volatile unsigned int *)c->comm[side].block, seq, c->comm_error.p, c->comm_timeout_cycles);
        LAUNCH(c, gfs::k_copy_batch, dim3((unsigned)batch_blocks(largest), (unsigned)cb.n), 256, cb);
    } else {
        LAUNCH(c, gfs::k_copy_batch_wait, dim3((unsigned)batch_blocks(largest), (unsigned)cb.n), 256, cb,
               (const volatile unsigned int *)c->comm[side].block, seq, c->comm_error.p, c->comm_timeout_cycles);
    }
    GFS_END()
}

namespace {
// raise the neighbours' particle flags (count word [2], then flag word [1] of their block) and queue the kernel that
// waits for theirs and publishes all counts to pinned host memory
void comm_signal_and_gather(gfs_context *c, const int has[2], const unsigned int seq[2]) {
    for (int s = 0; s < 2; s++)
        if (has[s])
            LAUNCH(c, gfs::k_signal, 1, 1, (volatile unsigned int *)c->comm[s].peer + 1, seq[s],
                   (volatile unsigned int *)c->comm[s].peer + 2, c->split_counters.p + 1 + s, 1);
    LAUNCH(c, gfs::k_gather_counts, 1, 1,
           has[0] ? (const volatile unsigned int *)c->comm[0].block + 1 : nullptr,
           has[1] ? (const volatile unsigned int *)c->comm[1].block + 1 : nullptr, seq[0], c->split_counters.p,
           (const volatile unsigned int *)c->comm[0].block + 2, (const volatile unsigned int *)c->comm[1].block + 2,
           c->comm_host, c->comm_error.p, c->comm_timeout_cycles);
}

void comm_migrate_split(gfs_context *c, int has_down, int has_up) {
    drop_dead(c);
    const int k_lo = has_down ? c->own_k0 : (int)0x80000000, k_hi = has_up ? c->own_k1 : 0x7FFFFFFF;
    unsigned int seq[2];
    for (int s = 0; s < 2; s++) seq[s] = ++c->comm[s].seq_particles;
    if (c->n > 0) {
        launch_split(c, k_lo, k_hi, has_down ? comm_arrivals(c, c->comm[0].peer, seq[0]) : nullptr,
                     has_up ? comm_arrivals(c, c->comm[1].peer, seq[1]) : nullptr, c->comm_particle_cap, c->split_counters.p);
    } else {
        GFS_CUDA(cudaMemsetAsync(c->split_counters.p, 0, 4 * sizeof(unsigned int), c->stream));
    }
    const int has[2] = {has_down, has_up};
    comm_signal_and_gather(c, has, seq);
    c->comm_fused = false;
}
}  // namespace

/* migration, first half: split the resident particles; leavers are written straight into the neighbours' arrival
 * buffers, followed by their count and the flag.  has_down / has_up: whether a neighbour exists on that side. */
void gfs_comm_migrate_begin(gfs_context *c, int has_down, int has_up, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE((!has_down || c->comm[0].peer) && (!has_up || c->comm[1].peer), "gfs_comm_connect first");
    GFS_CUDA(cudaSetDevice(c->device));
    comm_migrate_split(c, has_down, has_up);
    GFS_END()
}

/* G2P + advection with the migration fused into the kernel: a particle that leaves the owned layers is stored by the
 * G2P kernel itself into the neighbour's arrival buffer and becomes a dead slot here; the stayers are binned for the
 * next sort in the same epilogue.  Where the brick kernel does not apply (exact arithmetic, dx not a power of two)
 * this is gfs_g2p_advect followed by gfs_comm_migrate_begin.  Finish with gfs_comm_migrate_finish. */
void gfs_comm_g2p_advect(gfs_context *c, double dt, double ratio, int order, int interp, int arith, int has_down, int has_up, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE((!has_down || c->comm[0].peer) && (!has_up || c->comm[1].peer), "gfs_comm_connect first");
    GFS_REQUIRE(!removal_on(c), "options 5 and 6 (per-cell cap, removal in solids) are single-domain rules: switch them off for sharded runs");
    GFS_CUDA(cudaSetDevice(c->device));
    const bool fused = g2p_uses_bricks(c, arith) && c->p2g_variant >= 1 && c->lazy_sort && c->n - c->dead > 0 && (has_down || has_up);
    if (!fused) {
        do_g2p(c, dt, ratio, order, interp, arith, false);
        comm_migrate_split(c, has_down, has_up);
        return;
    }
    unsigned int seq[2];
    for (int s = 0; s < 2; s++) seq[s] = ++c->comm[s].seq_particles;
    GFS_CUDA(cudaMemsetAsync(c->split_counters.p, 0, 4 * sizeof(unsigned int), c->stream));
    gfs::Migrate mg;
    mg.own_lo = has_down ? c->own_k0 : (int)0x80000000;
    mg.own_hi = has_up ? c->own_k1 : 0x7FFFFFFF;
    mg.out[0] = has_down ? (float *)comm_arrivals(c, c->comm[0].peer, seq[0]) : nullptr;
    mg.out[1] = has_up ? (float *)comm_arrivals(c, c->comm[1].peer, seq[1]) : nullptr;
    mg.count = c->split_counters.p + 1;
    mg.cap = (unsigned int)c->comm_particle_cap;
    do_g2p(c, dt, ratio, order, interp, arith, true, &mg);
    const int has[2] = {has_down, has_up};
    comm_signal_and_gather(c, has, seq);
    c->comm_fused = true;
    GFS_END()
}

/* migration, second half: the one host synchronisation of a sharded substep; appends the arrivals.
 * moved[0] = particles sent away, moved[1] = particles received. */
void gfs_comm_migrate_finish(gfs_context *c, int64_t *moved, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_CUDA(cudaSetDevice(c->device));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    const unsigned int *h = c->comm_host;
    if (h[5] != 0) {
        // the exchange of this substep is incomplete (missing partial sums, stale halos or a local-only fixed-point scale):
        // the sharded result is NOT the single-domain one.  Clear the word so that the caller can retry or shut down.
        const unsigned int why = h[5];
        cudaMemsetAsync(c->comm_error.p, 0, sizeof(unsigned int), c->stream);
        cudaStreamSynchronize(c->stream);
        throw GfsError(std::string("peer exchange timed out: ") +
                       ((why & 2u) ? "a neighbour's grid layers or the all-ranks scale did not arrive (C1/C2/C0)" : "") +
                       ((why & 3u) == 3u ? "; " : "") + ((why & 1u) ? "a neighbour's migrating particles did not arrive (C3)" : "") +
                       " -- raise the limit with gfs_set_option(ctx, 7, seconds)");
    }
    GFS_REQUIRE((int64_t)h[1] <= c->comm_particle_cap && (int64_t)h[2] <= c->comm_particle_cap &&
                (int64_t)h[3] <= c->comm_particle_cap && (int64_t)h[4] <= c->comm_particle_cap, "migration buffer too small");
    const unsigned int n_in[2] = {h[3], h[4]};
    if (c->comm_fused) {
        // the leavers stay behind as dead slots (binned last); arrivals go to the end of the arrays, binned like the rest
        c->dead = (int64_t)h[1] + h[2];
        const int64_t total = c->n + n_in[0] + n_in[1];
        GFS_REQUIRE(total < 0x7FFFFFFFll, "particle count must fit int32");
        ensure_capacity(c, total);
        const int b = c->cur;
        for (int s = 0; s < 2; s++) {
            if (n_in[s] == 0) continue;
            LAUNCH(c, gfs::k_append_bin, ceil_div((int64_t)n_in[s], 256), 256, c->grid, c->nkeys, (int64_t)n_in[s], c->n,
                   (const float *)comm_arrivals(c, c->comm[s].block, c->comm[s].seq_particles),
                   c->soa[b][0].p, c->soa[b][1].p, c->soa[b][2].p, c->soa[b][3].p, c->soa[b][4].p, c->soa[b][5].p, c->tag[b].p,
                   c->keys[0].p, c->rank.p, c->counts.p, key_range(c));
            c->n += n_in[s];
        }
    } else {
        if (c->n > 0) c->cur = 1 - c->cur;
        c->n = h[0]; c->sorted = false; c->aos_valid = false; c->keys_ready = false; c->indexed = false;
        const bool redo0 = c->allmax_redo;          // appending migrated particles is not a "new particle set" (their max |v| travelled with the sender's)
        for (int s = 0; s < 2; s++) {
            if (n_in[s] == 0) continue;
            int e2 = GFS_SUCCESS;
            gfs_append_particles_device(c, comm_arrivals(c, c->comm[s].block, c->comm[s].seq_particles), n_in[s], &e2);
            if (e2 != GFS_SUCCESS) throw GfsError(g_error);
        }
        c->allmax_redo = redo0;
    }
    if (moved) { moved[0] = (int64_t)h[1] + h[2]; moved[1] = (int64_t)n_in[0] + n_in[1]; }
    GFS_END()
}

/* The merged C1 + C2 exchange of one side, stored for gfs_comm_substep (same arguments as push_layers / pull_layers). */
void gfs_comm_set_plan(gfs_context *c, int side, int n_push, const int *push_what, const int *push_first, const int *push_count,
                       const int64_t *push_offsets, int n_pull, const int *pull_what, const int *pull_first, const int *pull_count,
                       const int64_t *pull_offsets, const int *pull_add, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && (side == 0 || side == 1) && n_push >= 0 && n_push <= 16 && n_pull >= 0 && n_pull <= 16, "bad arguments");
    gfs_context::CommPlan &p = c->comm_plan[side];
    p.n_push = n_push; p.n_pull = n_pull;
    for (int i = 0; i < n_push; i++) { p.push_what[i] = push_what[i]; p.push_first[i] = push_first[i]; p.push_count[i] = push_count[i]; p.push_off[i] = push_offsets[i]; }
    for (int i = 0; i < n_pull; i++) { p.pull_what[i] = pull_what[i]; p.pull_first[i] = pull_first[i]; p.pull_count[i] = pull_count[i]; p.pull_off[i] = pull_offsets[i]; p.pull_add[i] = pull_add[i]; }
    GFS_END()
}

/* One sharded substep of this rank as a single call: index sort, agreed scale, P2G with the layer exchange of the stored
 * plan, G2P + RK with fused migration, and the host synchronisation that appends the arrivals.  Every rank calls it
 * once per substep.  moved[0] = particles sent away, moved[1] = received. */
void gfs_comm_substep(gfs_context *c, double dt, double ratio, int order, int interp, int arith, int has_down, int has_up,
                      int64_t *moved, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_CUDA(cudaSetDevice(c->device));
    GFS_REQUIRE(!removal_on(c), "options 5 and 6 (per-cell cap, removal in solids) are single-domain rules: switch them off for sharded runs");
    const int has[2] = {has_down, has_up};
    int e2 = GFS_SUCCESS;
#define GFS_SUB(call) do { call; if (e2 != GFS_SUCCESS) throw GfsError(g_error); } while (0)
    if (arith == GFS_EXACT) GFS_SUB(gfs_sort_unstable(c, &e2)); else GFS_SUB(gfs_sort_index(c, &e2));
    if (c->world_table && c->comm_world > 1) GFS_SUB(gfs_comm_allmax_scale(c, &e2));
    do_p2g_begin(c, arith);
    for (int s = 0; s < 2; s++) {
        if (!has[s]) continue;
        const gfs_context::CommPlan &p = c->comm_plan[s];
        GFS_SUB(gfs_comm_push_layers(c, s, p.n_push, p.push_what, p.push_first, p.push_count, p.push_off, &e2));
    }
    for (int s = 0; s < 2; s++) {
        if (!has[s]) continue;
        const gfs_context::CommPlan &p = c->comm_plan[s];
        GFS_REQUIRE(p.n_pull > 0, "gfs_comm_set_plan first");
        GFS_SUB(gfs_comm_pull_layers(c, s, p.n_pull, p.pull_what, p.pull_first, p.pull_count, p.pull_off, p.pull_add, &e2));
    }
    do_p2g_end(c);
    GFS_SUB(gfs_comm_g2p_advect(c, dt, ratio, order, interp, arith, has_down, has_up, &e2));
    // this rank's max |v| for the NEXT substep is final now (the G2P epilogue took it over every particle it advected,
    // leavers included; arrivals were counted by their sender): start the all-ranks maximum a whole sort ahead of its use
    // Whether to post is decided by CONFIGURATION only (the same on every rank): a rank whose G2P did not go through the
    // binning brick kernel this step (it held no particles) posts 0 -- its arrivals were counted by their senders.
    const bool config_fused = arith != GFS_EXACT && c->grid.pow2 && c->have_maps && c->g2p_variant >= 1 && c->p2g_variant >= 1 && c->lazy_sort;
    if (c->world_table && c->comm_world > 1 && c->allmax_early && config_fused) {
        if (!c->keys_ready) GFS_CUDA(cudaMemsetAsync(c->vmax_bits.p, 0, sizeof(unsigned int), c->stream));
        GFS_SUB(gfs_comm_allmax_post(c, &e2));
    }
    GFS_SUB(gfs_comm_migrate_finish(c, moved, &e2));
#undef GFS_SUB
    GFS_END()
}

/* ---- all-ranks table: the fixed-point scale of the P2G accumulators comes from max |v| over ALL particles, so the
 * integer partial sums of different GPUs are commensurable (and equal to the single-GPU run's).  gfs_comm_world_alloc,
 * exchange the handles, gfs_comm_world_connect for every rank (own rank included), then gfs_comm_allmax_scale once per
 * substep between the sort and gfs_p2g_begin. */
void gfs_comm_world_alloc(gfs_context *c, int rank, int world, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && world >= 1 && world <= 16 && rank >= 0 && rank < world, "world size must be 1..16");
    GFS_CUDA(cudaSetDevice(c->device));
    if (!c->world_table) GFS_CUDA(cudaMalloc((void **)&c->world_table, 2 * 16 * sizeof(unsigned long long)));
    GFS_CUDA(cudaMemset(c->world_table, 0, 2 * 16 * sizeof(unsigned long long)));
    c->comm_rank = rank; c->comm_world = world; c->seq_world = 0;
    c->allmax_posted = false; c->allmax_redo = false;
    for (int r = 0; r < 16; r++) c->world_peer[r] = nullptr;
    c->world_peer[rank] = c->world_table;
    c->comm_error.reserve(1);
    GFS_CUDA(cudaMemset(c->comm_error.p, 0, sizeof(unsigned int)));
    GFS_END()
}

void gfs_comm_world_export(gfs_context *c, void *handle64, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && handle64 && c->world_table, "gfs_comm_world_alloc first");
    cudaIpcMemHandle_t h;
    GFS_CUDA(cudaIpcGetMemHandle(&h, c->world_table));
    memcpy(handle64, &h, 64);
    GFS_END()
}

void gfs_comm_world_connect(gfs_context *c, int rank, const void *handle64, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && c->world_table && rank >= 0 && rank < c->comm_world && handle64, "bad arguments");
    if (rank == c->comm_rank) return;
    GFS_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    GFS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->world_peer[rank] = (unsigned long long *)p;
    c->world_peer_ipc[rank] = true;
    GFS_END()
}

void gfs_comm_world_connect_local(gfs_context *c, int rank, gfs_context *other, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && other && c->world_table && other->world_table && rank >= 0 && rank < c->comm_world, "bad arguments");
    c->world_peer[rank] = other->world_table;
    GFS_END()
}

void gfs_comm_allmax_scale(gfs_context *c, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(c->world_table, "gfs_comm_world_alloc first");
    for (int r = 0; r < c->comm_world; r++) GFS_REQUIRE(c->world_peer[r], "gfs_comm_world_connect every rank first");
    GFS_CUDA(cudaSetDevice(c->device));
    gfs::AllMaxPeers peers;
    for (int r = 0; r < 16; r++) peers.table[r] = c->world_peer[r];
    // mode 0: the whole all-ranks maximum here; after gfs_comm_allmax_post only the wait is left.  A particle upload between
    // the post and here (a collective act: every rank of a sharded run replaces its particles or none does) consumes the
    // outstanding round and exchanges afresh.
    if (c->allmax_posted) {
        unsigned int *dst = c->vmax_bits.p;
        if (c->allmax_redo) { c->split_counters.reserve(8); dst = c->split_counters.p + 4; }        // discard into a scratch word
        LAUNCH(c, gfs::k_allmax, 1, 32, peers, c->comm_rank, c->comm_world, c->seq_world, dst, c->comm_error.p, c->comm_timeout_cycles, 2);
    }
    if (!c->allmax_posted || c->allmax_redo)
        LAUNCH(c, gfs::k_allmax, 1, 32, peers, c->comm_rank, c->comm_world, ++c->seq_world, c->vmax_bits.p, c->comm_error.p, c->comm_timeout_cycles, 0);
    c->allmax_posted = false; c->allmax_redo = false;
    GFS_END()
}

/* First half of gfs_comm_allmax_scale, issued as soon as this rank's max |v| is final (right after its G2P epilogue and
 * collision resolve): the value travels to every rank while they sort and scan; gfs_comm_allmax_scale then only waits. */
void gfs_comm_allmax_post(gfs_context *c, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(c->world_table, "gfs_comm_world_alloc first");
    for (int r = 0; r < c->comm_world; r++) GFS_REQUIRE(c->world_peer[r], "gfs_comm_world_connect every rank first");
    GFS_REQUIRE(!c->allmax_posted, "gfs_comm_allmax_post twice without gfs_comm_allmax_scale");
    GFS_CUDA(cudaSetDevice(c->device));
    gfs::AllMaxPeers peers;
    for (int r = 0; r < 16; r++) peers.table[r] = c->world_peer[r];
    LAUNCH(c, gfs::k_allmax, 1, 32, peers, c->comm_rank, c->comm_world, ++c->seq_world, c->vmax_bits.p, c->comm_error.p, c->comm_timeout_cycles, 1);
    c->allmax_posted = true;
    GFS_END()
}

void gfs_extract_particles(gfs_context *c, int k_lo, int k_hi, void *down_device, void *up_device, int64_t cap,
                           int64_t *n_down, int64_t *n_up, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(n_down && n_up && cap >= 0 && cap < 0x7FFFFFFFll && (cap == 0 || (down_device && up_device)), "bad arguments");
    GFS_CUDA(cudaSetDevice(c->device));
    *n_down = *n_up = 0;
    drop_dead(c);
    if (c->n == 0) return;
    c->split_counters.reserve(8);
    launch_split(c, k_lo, k_hi, down_device, up_device, cap, c->split_counters.p);
    unsigned int h[4];
    GFS_CUDA(cudaMemcpyAsync(h, c->split_counters.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_REQUIRE((int64_t)h[1] <= cap && (int64_t)h[2] <= cap, "migration buffer too small");
    c->n = h[0]; c->cur = 1 - c->cur; c->sorted = false; c->aos_valid = false; c->keys_ready = false;
    *n_down = h[1]; *n_up = h[2];
    GFS_END()
}

/* Asynchronous variant: the three counts (kept, down, up, + one spare word) are left in caller-owned DEVICE memory and
 * nothing is synchronised; the caller ships the counts to its neighbours, reads everything back in one go and then
 * calls gfs_extract_commit with the number of kept particles. */
void gfs_extract_particles_async(gfs_context *c, int k_lo, int k_hi, void *down_device, void *up_device, int64_t cap,
                                 void *counters_device, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(counters_device && cap >= 0 && cap < 0x7FFFFFFFll && (cap == 0 || (down_device && up_device)), "bad arguments");
    GFS_CUDA(cudaSetDevice(c->device));
    drop_dead(c);
    if (c->n == 0) { GFS_CUDA(cudaMemsetAsync(counters_device, 0, 4 * sizeof(unsigned int), c->stream)); return; }
    launch_split(c, k_lo, k_hi, down_device, up_device, cap, (unsigned int *)counters_device);
    GFS_END()
}

void gfs_extract_commit(gfs_context *c, int64_t n_kept, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(n_kept >= 0 && n_kept <= c->n, "bad kept count");
    if (c->n > 0) c->cur = 1 - c->cur;
    c->n = n_kept; c->sorted = false; c->aos_valid = false; c->keys_ready = false;
    GFS_END()
}

void gfs_append_particles_device(gfs_context *c, const void *aos_device, int64_t n, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && n >= 0 && (n == 0 || aos_device), "bad arguments");
    GFS_CUDA(cudaSetDevice(c->device));
    if (n == 0) return;
    drop_dead(c);
    const int64_t old = c->n;
    int e2 = GFS_SUCCESS;
    gfs_resize_particles(c, old + n, &e2);
    if (e2 != GFS_SUCCESS) throw GfsError(g_error);
    const int b = c->cur;
    LAUNCH(c, gfs::k_append_aos, ceil_div(n, 256), 256, n, old, (const float *)aos_device, c->soa[b][0].p, c->soa[b][1].p,
           c->soa[b][2].p, c->soa[b][3].p, c->soa[b][4].p, c->soa[b][5].p, c->tag[b].p);
    GFS_END()
}

void *gfs_device_ptr(gfs_context *c, int which, int *err) {
    GFS_BEGIN
    require_domain(c);
    if (which >= 0 && which < 9) return c->field[which / 3][which % 3].p + gfs::kRowPad;
    if (which == 9) return c->material.p;
    if (which >= 10 && which < 16) return c->soa[c->cur][which - 10].p;
    if (which == 16) return c->vmax_bits.p;
    if (which >= 40 && which < 46) return c->press.vec[which - 40].p;
    if (which == 46) return c->press.trace.p;       // pressure system: r, z, s, p, q, precon (doubles per cell)
    throw GfsError("gfs_device_ptr: unknown buffer id");
    GFS_END(nullptr)
}

/* Verification hook: order- and distribution-independent 64-bit hashes of the resident state.  out[0] material of the
 * owned cell layers, out[1..3] the P2G u, v, w faces of the owned layers (w: the top face layer with the last slab),
 * out[4] the particle set (sum over particles of a hash of position + velocity bits).  Every value is a sum mod 2^64 of
 * per-element hashes keyed by the element's GLOBAL index, so the hashes of a sharded run, added over the ranks, equal
 * the single-GPU run's -- bench.py prints them at every GPU count.  Synchronises. */
void gfs_state_hash(gfs_context *c, uint64_t *out5, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(out5, "null pointer");
    GFS_CUDA(cudaSetDevice(c->device));
    drop_dead(c);
    const Grid &g = c->grid;
    DevBuf<unsigned long long> d;
    d.reserve(5);
    GFS_CUDA(cudaMemsetAsync(d.p, 0, 5 * sizeof(unsigned long long), c->stream));
    const int k0 = c->own_k0, k1 = c->own_k1;
    const int nblk = 148 * 8;
    LAUNCH(c, gfs::k_hash_grid, nblk, 256, c->material.p + (size_t)g.I * g.J * (size_t)(k0 - g.k0), 1, (long long)g.I, (long long)g.I,
           (long long)g.J * k0, (long long)g.J * (k1 - k0), 0x6d6174ull, d.p);
    const int ni[3] = {g.I + 1, g.I, g.I}, nj[3] = {g.J, g.J + 1, g.J};
    for (int a = 0; a < 3; a++) {
        const int layers = (k1 - k0) + ((a == 2 && k1 == g.K) ? 1 : 0);
        const float *base = c->field[GFS_FIELD_P2G][a].p + gfs::kRowPad + (size_t)g.pitch[a] * nj[a] * (size_t)(k0 - g.k0);
        LAUNCH(c, gfs::k_hash_grid, nblk, 256, base, 4, (long long)ni[a], (long long)g.pitch[a], (long long)nj[a] * k0,
               (long long)nj[a] * layers, 0x750000ull + (unsigned long long)a, d.p + 1 + a);
    }
    if (c->n > 0) {
        const int b = c->cur;
        LAUNCH(c, gfs::k_hash_particles, nblk, 256, c->n, (const uint32_t *)c->soa[b][0].p, (const uint32_t *)c->soa[b][1].p,
               (const uint32_t *)c->soa[b][2].p, (const uint32_t *)c->soa[b][3].p, (const uint32_t *)c->soa[b][4].p,
               (const uint32_t *)c->soa[b][5].p, d.p + 4);
    }
    unsigned long long h[5];
    GFS_CUDA(cudaMemcpyAsync(h, d.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    d.release();
    for (int i = 0; i < 5; i++) out5[i] = (uint64_t)h[i];
    GFS_END()
}

namespace {
// the geometry FluidSimulation derives from a source before it touches a particle, with the reference's arithmetic:
// FluidSource::_getOverlappingCells bounds, getAABB, Grid3d::fitAABBtoGrid, Grid3d::getGridIndexBounds(AABB), GridIndexToPosition
gfs::Emitter make_emitter(const gfs_source_t &src, const Grid &g) {
    gfs::Emitter e;
    e.src = src;
    const double dx = g.dx, inv = 1.0 / dx;
    const int size[3] = {g.I, g.J, g.K};
    auto idx = [&](float p) { return (int)std::floor((double)p * inv); };                       // positionToGridIndex, grid3d.h:58-63
    auto gpos = [&](int i) { return (float)((double)(float)i * dx); };                          // GridIndexToPosition, grid3d.h:83-85
    float bp[3];
    double bw[3];
    if (src.kind == 0) {
        const double r = src.a;
        for (int a = 0; a < 3; a++) {                                                           // getGridIndexBounds(p, r, dx, ...), grid3d.h:350-371
            const int c = idx(src.p[a]);
            const float trans = src.p[a] - gpos(c);
            const int lo = c - (int)std::fmax(0.0, std::ceil((r - (double)trans) * inv));
            const int hi = c + (int)std::fmax(0.0, std::ceil((r - dx + (double)trans) * inv));
            e.smin[a] = (int)std::fmax((double)lo, 0.0);
            e.smax[a] = (int)std::fmin((double)hi, (double)(size[a] - 1));
            bp[a] = src.p[a] - (float)r;                                                        // SphericalFluidSource::getAABB
            bw[a] = 2.0 * r;
        }
    } else {
        const double ext[3] = {src.a, src.b, src.c};
        for (int a = 0; a < 3; a++) {                                                           // getGridIndexBounds(AABB), grid3d.h:426-439
            bp[a] = src.p[a]; bw[a] = ext[a];
            e.smin[a] = (int)std::fmax((double)idx(bp[a]), 0.0);
            e.smax[a] = (int)std::fmin((double)idx(bp[a] + (float)bw[a]), (double)(size[a] - 1));
        }
    }
    // fitAABBtoGrid (grid3d.h:485-501)
    float pmin[3], pmax[3];
    int gmin[3], gmax[3];
    for (int a = 0; a < 3; a++) { pmin[a] = bp[a]; pmax[a] = bp[a] + (float)bw[a]; gmin[a] = idx(pmin[a]); gmax[a] = idx(pmax[a]); }
    auto inrange = [&](const int *q) { return q[0] >= 0 && q[1] >= 0 && q[2] >= 0 && q[0] < size[0] && q[1] < size[1] && q[2] < size[2]; };
    if (!inrange(gmin)) for (int a = 0; a < 3; a++) pmin[a] = 0.0f;
    if (!inrange(gmax)) for (int a = 0; a < 3; a++) pmax[a] = (gpos(gmax[a]) + (float)dx) - 10e-9f;
    for (int a = 0; a < 3; a++) {                                                               // AABB(p1, p2), aabb.cpp:33-45
        const double lo = std::fmin((double)pmin[a], (double)pmax[a]), hi = std::fmax((double)pmin[a], (double)pmax[a]);
        e.bpos[a] = (float)lo; e.bext[a] = hi - lo;
        e.bmin[a] = (int)std::fmax((double)idx(e.bpos[a]), 0.0);
        e.bmax[a] = (int)std::fmin((double)idx(e.bpos[a] + (float)e.bext[a]), (double)(size[a] - 1));
        e.offset[a] = gpos(e.bmin[a]);
    }
    return e;
}
}  // namespace

/* FluidSimulation::_updateFluidSources for the ACTIVE INFLOW sources last given to gfs_set_sources
 * (src/fluidsimulation.cpp:1823-1879): air cells a source overlaps are seeded with 8 particles, and every empty half-dx
 * sub-cell of its fluid-or-air cells (occupancy taken from the particles inside its grid-fitted bounding box) gets one
 * particle; all with the source's velocity.  jitter = 0.25 * jitter factor * dx (src/fluidsimulation.cpp:1219-1221); the
 * reference draws it from rand(), here it is a hash of (seed, cell, sub-cell): the set of emitting sub-cells is the
 * reference's, positions agree to the jitter.  Uses the resident material grid (the previous classification) like the
 * reference; single domain.  *emitted = particles added. */
void gfs_emit_from_sources(gfs_context *c, double jitter, uint64_t seed, int64_t *emitted, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(jitter >= 0, "bad jitter");
    GFS_REQUIRE(c->own_k0 == 0 && c->own_k1 == c->grid.K, "gfs_emit_from_sources is single-domain only");
    GFS_CUDA(cudaSetDevice(c->device));
    if (emitted) *emitted = 0;
    if (c->sources.n == 0) return;
    drop_dead(c);
    const Grid &g = c->grid;
    std::vector<gfs::Emitter> em;
    int64_t bound = 0;
    size_t occ_words = 1;
    for (int s = 0; s < c->sources.n; s++) {
        gfs::Emitter e = make_emitter(c->sources.s[s], g);
        const int64_t w = e.bmax[0] - e.bmin[0] + 1, h = e.bmax[1] - e.bmin[1] + 1, d = e.bmax[2] - e.bmin[2] + 1;
        if (w <= 0 || h <= 0 || d <= 0) continue;
        bound += 16 * w * h * d;
        occ_words = std::max(occ_words, (size_t)((8 * w * h * d + 31) / 32));
        em.push_back(e);
    }
    if (em.empty()) return;
    GFS_REQUIRE(c->n + bound < 0x7FFFFFFFll, "particle count must fit int32");
    const int64_t n0 = c->n;
    ensure_capacity(c, n0 + bound);
    c->src_occ.reserve(occ_words);
    c->src_count.reserve(1);
    GFS_CUDA(cudaMemsetAsync(c->src_count.p, 0, sizeof(unsigned int), c->stream));
    const int b = c->cur;
    gfs::EmitOut o;
    o.x = c->soa[b][0].p; o.y = c->soa[b][1].p; o.z = c->soa[b][2].p; o.vx = c->soa[b][3].p; o.vy = c->soa[b][4].p; o.vz = c->soa[b][5].p;
    o.tag = c->tag[b].p; o.count = c->src_count.p; o.at = n0; o.cap = bound;
    for (size_t s = 0; s < em.size(); s++) {
        const gfs::Emitter &e = em[s];
        const long long cells = (long long)(e.bmax[0] - e.bmin[0] + 1) * (e.bmax[1] - e.bmin[1] + 1) * (e.bmax[2] - e.bmin[2] + 1);
        GFS_CUDA(cudaMemsetAsync(c->src_occ.p, 0, occ_words * sizeof(unsigned int), c->stream));
        if (n0 > 0)
            LAUNCH(c, gfs::k_source_mark, ceil_div(n0, 256), 256, g, e, (int64_t)0, n0, (const unsigned int *)nullptr, o.x, o.y, o.z, c->src_occ.p);
        if (s > 0)          // what the earlier sources of this call emitted counts too
            LAUNCH(c, gfs::k_source_mark, ceil_div(bound, 256), 256, g, e, n0, bound, (const unsigned int *)c->src_count.p, o.x, o.y, o.z, c->src_occ.p);
        LAUNCH(c, gfs::k_source_emit, ceil_div(cells, 128), 128, g, e, c->material.p, c->src_occ.p, o, (unsigned long long)seed + 0x9E3779B97F4A7C15ull * (s + 1), jitter);
    }
    unsigned int count = 0;
    GFS_CUDA(cudaMemcpyAsync(&count, c->src_count.p, sizeof(count), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_REQUIRE((int64_t)count <= bound, "internal: emission bound exceeded");
    c->n = n0 + count;
    if (count > 0) {
        c->sorted = false; c->aos_valid = false; c->keys_ready = false; c->indexed = false;
        if (c->allmax_posted) c->allmax_redo = true;
    }
    c->graph_epoch++;
    if (emitted) *emitted = count;
    GFS_END()
}

/* The outflow half of _updateFluidSources (:1853-1877): the particles in the FLUID cells (resident material grid) that the
 * given outflow sources overlap are removed (_removeMarkerParticlesFromCells, :1717-1730).  *removed = particles gone. */
void gfs_remove_in_sources(gfs_context *c, const gfs_source_t *outflow, int nsources, int64_t *removed, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(nsources >= 0 && (nsources == 0 || outflow), "bad arguments");
    GFS_REQUIRE(c->own_k0 == 0 && c->own_k1 == c->grid.K, "gfs_remove_in_sources is single-domain only");
    GFS_CUDA(cudaSetDevice(c->device));
    if (removed) *removed = 0;
    if (nsources == 0 || c->n - c->dead == 0) return;
    drop_dead(c);
    const Grid &g = c->grid;
    c->src_removal.reserve(c->cell_count);
    c->src_count.reserve(1);
    GFS_CUDA(cudaMemsetAsync(c->src_removal.p, 0, c->cell_count, c->stream));
    GFS_CUDA(cudaMemsetAsync(c->src_count.p, 0, sizeof(unsigned int), c->stream));
    for (int s = 0; s < nsources; s++) {
        const gfs::Emitter e = make_emitter(outflow[s], g);
        const long long cells = (long long)(e.smax[0] - e.smin[0] + 1) * (e.smax[1] - e.smin[1] + 1) * (e.smax[2] - e.smin[2] + 1);
        if (cells > 0) LAUNCH(c, gfs::k_outflow_cells, ceil_div(cells, 128), 128, g, e, c->material.p, c->src_removal.p);
    }
    const int src = c->cur, dst = 1 - c->cur;
    LAUNCH(c, gfs::k_remove_in_cells, ceil_div(c->n, 256), 256, g, c->n, c->src_removal.p,
           c->soa[src][0].p, c->soa[src][1].p, c->soa[src][2].p, c->soa[src][3].p, c->soa[src][4].p, c->soa[src][5].p, c->tag[src].p,
           c->soa[dst][0].p, c->soa[dst][1].p, c->soa[dst][2].p, c->soa[dst][3].p, c->soa[dst][4].p, c->soa[dst][5].p, c->tag[dst].p, c->src_count.p);
    unsigned int kept = 0;
    GFS_CUDA(cudaMemcpyAsync(&kept, c->src_count.p, sizeof(kept), cudaMemcpyDeviceToHost, c->stream));
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    if (removed) *removed = c->n - (int64_t)kept;
    c->removed += c->n - (int64_t)kept;
    c->n = kept; c->cur = dst;
    c->sorted = false; c->aos_valid = false; c->keys_ready = false; c->indexed = false;
    c->graph_epoch++;
    GFS_END()
}

/* Allocate, now, everything a substep would otherwise size lazily: particle arrays for `particle_capacity` slots (current
 * contents kept) and the scratch of the sort / G2P / exchange.  cudaMalloc / cudaFree synchronise the device, which a
 * sharded step cannot afford between its device-side waits when several slabs share one GPU -- and is wasted time anywhere. */
void gfs_reserve(gfs_context *c, int64_t particle_capacity, int *err) {
    GFS_BEGIN
    require_domain(c);
    GFS_REQUIRE(particle_capacity >= 0 && particle_capacity < 0x7FFFFFFFll, "bad capacity");
    GFS_CUDA(cudaSetDevice(c->device));
    const int64_t cap = particle_capacity > c->n ? particle_capacity : c->n;
    if ((size_t)cap > c->soa[0][0].cap) {
        const int64_t n0 = c->n;
        // ensure_capacity grows to n + n/8 + 1024 of its argument: ask for what is wanted, net of that slack
        ensure_capacity(c, cap);
        c->n = n0;
    }
    size_t tmp_bytes = 0;
    GFS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->counts.p, (uint32_t *)c->cell_start.p, (int)((size_t)c->nkeys + 3), c->stream));
    c->cub_tmp.reserve(tmp_bytes);
    if (c->resolve_collisions) {
        const size_t want = c->coll_cap_user > 0 ? (size_t)c->coll_cap_user : (size_t)(cap / 16 + 4096) + (size_t)cap / 64;
        if (want > c->coll_list.cap) c->coll_list.reserve(want);
        c->coll_count.reserve(1);
    }
    c->slow_count.reserve(1);
    c->split_counters.reserve(8);
    if (!c->removal_host) GFS_CUDA(cudaHostAlloc((void **)&c->removal_host, 64, cudaHostAllocDefault));
    // CUDA loads kernels lazily, at their first launch, and a load can synchronise the device like an allocation does:
    // touch every kernel a (sharded) substep launches, and run the scan once for cub's
    {
        cudaFuncAttributes fa;
#define GFS_TOUCH(k) GFS_CUDA(cudaFuncGetAttributes(&fa, k))
        GFS_TOUCH(gfs::k_hist); GFS_TOUCH(gfs::k_build_index); GFS_TOUCH(gfs::k_scatter_sorted); GFS_TOUCH(gfs::k_scan_tail);
        GFS_TOUCH(gfs::k_clamp_counts); GFS_TOUCH(gfs::k_classify); GFS_TOUCH(gfs::k_p2g_finalize); GFS_TOUCH(gfs::k_assemble);
        GFS_TOUCH(gfs::k_finalize_assemble);
        GFS_TOUCH(gfs::k_p2g_tile<0>); GFS_TOUCH(gfs::k_p2g_tile<2>); GFS_TOUCH(gfs::k_p2g_tile2<false>); GFS_TOUCH(gfs::k_p2g_tile2<true>);
        GFS_TOUCH(gfs::k_p2g_scatter<0>); GFS_TOUCH(gfs::k_p2g_scatter<2>);
        GFS_TOUCH((gfs::k_g2p_tri<false, false>)); GFS_TOUCH((gfs::k_g2p_tri<false, true>)); GFS_TOUCH((gfs::k_g2p_tri<true, false>)); GFS_TOUCH((gfs::k_g2p_tri<true, true>));
        GFS_TOUCH((gfs::k_g2p_brick<0, false>)); GFS_TOUCH((gfs::k_g2p_brick<0, true>)); GFS_TOUCH((gfs::k_g2p_brick<1, false>)); GFS_TOUCH((gfs::k_g2p_brick<1, true>));
        GFS_TOUCH(gfs::k_g2p_slow<false>); GFS_TOUCH(gfs::k_g2p_slow<true>); GFS_TOUCH(gfs::k_g2p_advect<0>); GFS_TOUCH(gfs::k_g2p_advect<2>);
        GFS_TOUCH(gfs::k_resolve_collisions); GFS_TOUCH(gfs::k_copy_batch); GFS_TOUCH(gfs::k_copy_batch_wait); GFS_TOUCH(gfs::k_wait_flag);
        GFS_TOUCH(gfs::k_signal); GFS_TOUCH(gfs::k_gather_counts); GFS_TOUCH(gfs::k_allmax); GFS_TOUCH(gfs::k_append_bin);
        GFS_TOUCH(gfs::k_append_aos); GFS_TOUCH(gfs::k_split_by_layer); GFS_TOUCH(gfs::k_add_u64); GFS_TOUCH(gfs::k_gather_sorted_aos);
#undef GFS_TOUCH
        GFS_CUDA(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp_bytes, c->counts.p, (uint32_t *)c->cell_start.p, (int)((size_t)c->nkeys + 3), c->stream));
        c->sorted = false;                         // (the cell table is scratch until the next sort; the particles are untouched)
    }
    GFS_CUDA(cudaStreamSynchronize(c->stream));
    GFS_END()
}

void gfs_resize_particles(gfs_context *c, int64_t n, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(c && n >= 0 && n < 0x7FFFFFFFll, "bad arguments");
    GFS_CUDA(cudaSetDevice(c->device));
    drop_dead(c);
    ensure_capacity(c, n);
    c->n = n;
    c->sorted = false; c->aos_valid = false;
    c->keys_ready = false;
    if (c->allmax_posted) c->allmax_redo = true;
    GFS_END()
}

extern "C++" {
namespace {
// grow the particle buffers to hold n slots, keeping the current contents (particles, tags and their binning)
void ensure_capacity(gfs_context *c, int64_t n) {
    if ((size_t)n > c->soa[0][0].cap) {
        c->graph_epoch++;
        // grow both buffer sets, keeping the current contents
        const int b = c->cur;
        int64_t keep = c->n < n ? c->n : n;
        size_t newcap = (size_t)n + (size_t)n / 8 + 1024;
        for (int a = 0; a < 6; a++) {
            DevBuf<float> nb; nb.reserve(newcap);
            if (keep > 0) GFS_CUDA(cudaMemcpyAsync(nb.p, c->soa[b][a].p, (size_t)keep * 4, cudaMemcpyDeviceToDevice, c->stream));
            GFS_CUDA(cudaStreamSynchronize(c->stream));
            c->soa[b][a].release(); c->soa[b][a] = nb;
            c->soa[1 - b][a].release(); c->soa[1 - b][a].reserve(newcap);
        }
        DevBuf<int32_t> nt; nt.reserve(newcap);
        if (keep > 0) GFS_CUDA(cudaMemcpyAsync(nt.p, c->tag[b].p, (size_t)keep * 4, cudaMemcpyDeviceToDevice, c->stream));
        GFS_CUDA(cudaStreamSynchronize(c->stream));
        c->tag[b].release(); c->tag[b] = nt;
        c->tag[1 - b].release(); c->tag[1 - b].reserve(newcap);
        DevBuf<uint32_t> nk, nr; nk.reserve(newcap); nr.reserve(newcap);
        if (keep > 0) {
            GFS_CUDA(cudaMemcpyAsync(nk.p, c->keys[0].p, (size_t)keep * 4, cudaMemcpyDeviceToDevice, c->stream));
            GFS_CUDA(cudaMemcpyAsync(nr.p, c->rank.p, (size_t)keep * 4, cudaMemcpyDeviceToDevice, c->stream));
        }
        GFS_CUDA(cudaStreamSynchronize(c->stream));
        c->keys[0].release(); c->keys[0] = nk;
        c->rank.release(); c->rank = nr;
        c->keys[1].release(); c->keys[1].reserve(newcap);
        for (int q = 0; q < 2; q++) { c->perm[q].release(); c->perm[q].reserve(newcap); }
        c->index.release(); c->index.reserve(newcap);
    }
}
}  // namespace
}  // extern "C++"

/* ---- z-slab helpers (host arithmetic only) ------------------------------------------------------ */

void gfs_slab_range(int ksize, int nranks, int rank, int *k0, int *k1, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(ksize > 0 && nranks > 0 && rank >= 0 && rank < nranks && k0 && k1, "bad arguments");
    *k0 = (int)((int64_t)ksize * rank / nranks);
    *k1 = (int)((int64_t)ksize * (rank + 1) / nranks);
    GFS_END()
}

int gfs_slab_owner(int ksize, int nranks, int k, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(ksize > 0 && nranks > 0, "bad arguments");
    if (k < 0) k = 0;
    if (k >= ksize) k = ksize - 1;
    // inverse of gfs_slab_range: the largest r with floor(ksize*r/nranks) <= k
    int r = (int)(((int64_t)(k + 1) * nranks - 1) / ksize);
    while (r > 0 && (int64_t)ksize * r / nranks > k) r--;
    while (r + 1 < nranks && (int64_t)ksize * (r + 1) / nranks <= k) r++;
    return r;
    GFS_END(-1)
}

int gfs_slab_halo_cells(int interp, double max_displacement, double dx, int *err) {
    GFS_BEGIN
    GFS_REQUIRE(dx > 0 && max_displacement >= 0, "bad arguments");
    int stencil = interp == GFS_TRICUBIC ? 2 : 1;
    return stencil + (int)std::ceil(max_displacement / dx);
    GFS_END(-1)
}

}  // extern "C"
