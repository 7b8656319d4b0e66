// gfs_mg.cu -- the native multi-GPU entry points (gfs_mg_create / scatter / substep / gather): one process,
// one context + one host thread per GPU, z-slab sharding, neighbour exchange through peer memory (the gfs_comm_* machinery:
// kernels of one GPU write layers, migrating particles and flags straight into the neighbour's HBM over NVLink).
//
// Everything here is host C++11 on top of the single-GPU C-ABI of this library -- what gridfluidsim3d_b200/slabs.py does
// from Python for the multi-process runs (exchange plans, particle-weighted cuts), so that a C++ host can shard without
// an interpreter, without NCCL and without IPC handles (the contexts share an address space: cudaDeviceEnablePeerAccess).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/gfs_b200.h"

namespace {
thread_local char g_mg_error[4096] = "";
void mg_set_error(const std::string &m) { snprintf(g_mg_error, sizeof(g_mg_error), "%s", m.c_str()); }
struct MgError { std::string what; };
#define MG_REQUIRE(cond, msg) do { if (!(cond)) throw MgError{std::string(msg) + " (" #cond ")"}; } while (0)
#define MG_CALL(call) do { int e_ = GFS_SUCCESS; call; if (e_ != GFS_SUCCESS) throw MgError{std::string(#call ": ") + gfs_get_error_message()}; } while (0)
}  // namespace

struct gfs_mg {
    int n = 0;
    std::vector<int> device;
    std::vector<gfs_context *> ctx;
    int I = 0, J = 0, K = 0, halo = 2;
    double dx = 0;
    std::vector<int> k0, k1;                 // owned cell layers per rank
    bool connected = false;
    std::vector<uint8_t> material;           // the domain's material grid as last set (uploaded to every rank)
    bool have_material = false;
};

namespace {

// run f(rank) on one host thread per rank; the first error wins
template <typename F>
void for_each_rank(gfs_mg *mg, F f) {
    std::vector<std::string> errors(mg->n);
    std::vector<std::thread> th;
    for (int r = 0; r < mg->n; r++)
        th.emplace_back([&, r]() {
            try { f(r); } catch (const MgError &e) { errors[r] = e.what.empty() ? "error" : e.what; }
        });
    for (auto &t : th) t.join();
    for (int r = 0; r < mg->n; r++)
        if (!errors[r].empty()) throw MgError{"rank " + std::to_string(r) + ": " + errors[r]};
}

// cuts that balance the PARTICLES over the slabs (slabs.slab_ranges_weighted): cut r goes after the layer where the running
// count reaches r/n of the total; every slab keeps at least min_layers layers
void weighted_cuts(const std::vector<double> &counts, int n, int min_layers, std::vector<int> &k0, std::vector<int> &k1) {
    const int K = (int)counts.size();
    double total = 0;
    for (double c : counts) total += c;
    std::vector<int> cuts(1, 0);
    if (n == 1 || total <= 0 || K < n * min_layers) {
        for (int r = 1; r < n; r++) cuts.push_back((int)((long long)K * r / n));
    } else {
        double run = 0;
        int k = 0;
        for (int r = 1; r < n; r++) {
            const double target = total * r / n;
            while (k < K && run + counts[k] <= target) run += counts[k++];
            if (k < K && (run + counts[k] - target) < (target - run)) run += counts[k++];
            k = std::max(k, cuts.back() + min_layers);
            k = std::min(k, K - (n - r) * min_layers);
            run = 0;
            for (int q = 0; q < k; q++) run += counts[q];
            cuts.push_back(k);
        }
    }
    cuts.push_back(K);
    k0.assign(cuts.begin(), cuts.end() - 1);
    k1.assign(cuts.begin() + 1, cuts.end());
}

struct Item { int what, sf, sc, rf, rc, add; };

// what travels to / from the neighbour on `side` (0 = down, 1 = up) in the merged C1 + C2 exchange of one substep:
// the two accumulator node layers either side of the cut (added as integers) + my classified boundary material layer,
// then `halo` layers of the NEW and SAVED fields (slabs.SlabDriver._items)
std::vector<Item> exchange_items(const gfs_mg *mg, int r, int side) {
    const int k0 = mg->k0[r], k1 = mg->k1[r], H = mg->halo, K = mg->K;
    std::vector<Item> it;
    const int cut = side == 0 ? k0 : k1;
    for (int w = 10; w < 13; w++) it.push_back({w, cut - 1, 2, cut - 1, 2, 1});
    if (side == 0) it.push_back({9, k0, 1, k0 - 1, 1, 0}); else it.push_back({9, k1 - 1, 1, k1, 1, 0});
    int s0, s1, r0, r1;
    if (side == 0) { s0 = k0; s1 = std::min(H, k1 - k0); r0 = std::max(k0 - H, 0); r1 = k0 - r0; }
    else { s0 = std::max(k1 - H, k0); s1 = k1 - s0; r0 = k1; r1 = std::min(k1 + H, K) - k1; }
    for (int w = 0; w < 6; w++) it.push_back({w, s0, s1, r0, r1, 0});
    return it;
}

void connect(gfs_mg *mg, int64_t particle_cap) {
    const int n = mg->n;
    // message sizes: the same on both ends of every link -> take the maximum over all ranks and sides
    int64_t layer_bytes = 256;
    for (int r = 0; r < n; r++)
        for (int side = 0; side < 2; side++) {
            if ((side == 0 && r == 0) || (side == 1 && r == n - 1)) continue;
            int64_t sb = 0, rb = 0;
            for (const Item &it : exchange_items(mg, r, side)) {
                int e = GFS_SUCCESS;
                const int64_t lb = gfs_layer_bytes(mg->ctx[r], it.what, &e);
                if (e != GFS_SUCCESS) throw MgError{gfs_get_error_message()};
                sb += lb * it.sc; rb += lb * it.rc;
            }
            layer_bytes = std::max(layer_bytes, std::max(sb, rb));
        }
    for (int r = 0; r < n; r++) {
        MG_CALL(gfs_set_owned_layers(mg->ctx[r], mg->k0[r], mg->k1[r], &e_));
        MG_CALL(gfs_comm_alloc(mg->ctx[r], layer_bytes, particle_cap, &e_));
        MG_CALL(gfs_comm_world_alloc(mg->ctx[r], r, n, &e_));
    }
    for (int r = 0; r < n; r++) {
        if (r > 0) MG_CALL(gfs_comm_connect_local(mg->ctx[r], 0, mg->ctx[r - 1], &e_));
        if (r + 1 < n) MG_CALL(gfs_comm_connect_local(mg->ctx[r], 1, mg->ctx[r + 1], &e_));
        for (int q = 0; q < n; q++)
            if (q != r) MG_CALL(gfs_comm_world_connect_local(mg->ctx[r], q, mg->ctx[q], &e_));
        for (int side = 0; side < 2; side++) {
            if ((side == 0 && r == 0) || (side == 1 && r == n - 1)) continue;
            const std::vector<Item> items = exchange_items(mg, r, side);
            std::vector<int> pw, pf, pc, qw, qf, qc, qa;
            std::vector<int64_t> po, qo;
            int64_t sb = 0, rb = 0;
            for (const Item &it : items) {
                int e = GFS_SUCCESS;
                const int64_t lb = gfs_layer_bytes(mg->ctx[r], it.what, &e);
                pw.push_back(it.what); pf.push_back(it.sf); pc.push_back(it.sc); po.push_back(sb);
                qw.push_back(it.what); qf.push_back(it.rf); qc.push_back(it.rc); qo.push_back(rb); qa.push_back(it.add);
                sb += lb * it.sc; rb += lb * it.rc;
            }
            MG_CALL(gfs_comm_set_plan(mg->ctx[r], side, (int)items.size(), pw.data(), pf.data(), pc.data(), po.data(),
                                      (int)items.size(), qw.data(), qf.data(), qc.data(), qo.data(), qa.data(), &e_));
        }
    }
    mg->connected = true;
}

}  // namespace

#define MG_BEGIN if (err) *err = GFS_SUCCESS; try {
#define MG_END(retval) } catch (const MgError &e) { mg_set_error(e.what); if (err) *err = GFS_FAIL; return retval; }

extern "C" {

const char *gfs_mg_get_error_message(void) { return g_mg_error; }

gfs_mg *gfs_mg_create(int ndev, const int *devices, int I, int J, int K, double dx, int halo_layers, int *err) {
    MG_BEGIN
    MG_REQUIRE(ndev >= 1 && ndev <= 16, "1..16 GPUs");
    MG_REQUIRE(I > 0 && J > 0 && K > 0 && dx > 0 && halo_layers >= 1, "bad domain");
    MG_REQUIRE(K >= ndev * std::max(halo_layers, 4), "too few cell layers for this many slabs");
    int count = 0;
    MG_REQUIRE(cudaGetDeviceCount(&count) == cudaSuccess && count > 0, "no CUDA device");
    gfs_mg *mg = new gfs_mg();
    mg->n = ndev; mg->I = I; mg->J = J; mg->K = K; mg->dx = dx; mg->halo = halo_layers;
    try {
        for (int r = 0; r < ndev; r++) {
            const int d = devices ? devices[r] : r % count;        // several ranks may share a device (tests on one GPU)
            MG_REQUIRE(d >= 0 && d < count, "device index out of range");
            mg->device.push_back(d);
        }
        for (int r = 0; r < ndev; r++)
            for (int q = 0; q < ndev; q++) {
                if (mg->device[r] == mg->device[q]) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, mg->device[r], mg->device[q]);
                MG_REQUIRE(can, "the GPUs of a gfs_mg group must have peer access to each other (NVLink / NVSwitch)");
                cudaSetDevice(mg->device[r]);
                cudaError_t pe = cudaDeviceEnablePeerAccess(mg->device[q], 0);
                if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else MG_REQUIRE(pe == cudaSuccess, "cudaDeviceEnablePeerAccess failed");
            }
        for (int r = 0; r < ndev; r++) {
            int e = GFS_SUCCESS;
            gfs_context *c = gfs_create(mg->device[r], nullptr, &e);
            if (e != GFS_SUCCESS) throw MgError{gfs_get_error_message()};
            mg->ctx.push_back(c);
            MG_CALL(gfs_domain_init(c, I, J, K, dx, &e_));
            bool shared = false;
            for (int q = 0; q < ndev; q++) shared = shared || (q != r && mg->device[q] == mg->device[r]);
            if (shared) MG_CALL(gfs_set_option(c, 10, 1, &e_));      // slabs sharing a GPU: waits must not occupy the SMs
        }
        // uniform slabs until the particles are known
        for (int r = 0; r < ndev; r++) { mg->k0.push_back((int)((long long)K * r / ndev)); mg->k1.push_back((int)((long long)K * (r + 1) / ndev)); }
    } catch (...) {
        for (gfs_context *c : mg->ctx) { int e; gfs_destroy(c, &e); }
        delete mg;
        throw;
    }
    return mg;
    MG_END(nullptr)
}

void gfs_mg_destroy(gfs_mg *mg, int *err) {
    MG_BEGIN
    if (!mg) return;
    for (gfs_context *c : mg->ctx) { int e; gfs_sync(c, &e); }
    for (gfs_context *c : mg->ctx) { int e; gfs_destroy(c, &e); }
    delete mg;
    MG_END()
}

int gfs_mg_num_devices(gfs_mg *mg) { return mg ? mg->n : 0; }
gfs_context *gfs_mg_context(gfs_mg *mg, int rank) { return (mg && rank >= 0 && rank < mg->n) ? mg->ctx[rank] : nullptr; }

void gfs_mg_get_slab(gfs_mg *mg, int rank, int *k0, int *k1, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg && rank >= 0 && rank < mg->n && k0 && k1, "bad arguments");
    *k0 = mg->k0[rank]; *k1 = mg->k1[rank];
    MG_END()
}

void gfs_mg_set_option(gfs_mg *mg, int option, int value, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg, "null group");
    for (gfs_context *c : mg->ctx) MG_CALL(gfs_set_option(c, option, value, &e_));
    MG_END()
}

void gfs_mg_set_material(gfs_mg *mg, const uint8_t *material, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg && material, "bad arguments");
    mg->material.assign(material, material + (size_t)mg->I * mg->J * mg->K);
    mg->have_material = true;
    for_each_rank(mg, [&](int r) { MG_CALL(gfs_set_material(mg->ctx[r], material, &e_)); });
    MG_END()
}

void gfs_mg_set_sources(gfs_mg *mg, const gfs_source_t *sources, int nsources, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg, "null group");
    for (gfs_context *c : mg->ctx) MG_CALL(gfs_set_sources(c, sources, nsources, &e_));
    MG_END()
}

/* u, v, w: the WHOLE arrays in the reference's layout; every rank takes its owned layers plus the halo */
void gfs_mg_set_field(gfs_mg *mg, int slot, const float *u, const float *v, const float *w, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg && u && v && w, "bad arguments");
    for_each_rank(mg, [&](int r) {
        const int lo = std::max(0, mg->k0[r] - mg->halo), hi = std::min(mg->K, mg->k1[r] + mg->halo);
        MG_CALL(gfs_set_field_layers(mg->ctx[r], slot, u, v, w, lo, hi - lo, &e_));
    });
    MG_END()
}

/* every rank contributes the layers it owns (w: the top face layer with the last slab) */
void gfs_mg_get_field(gfs_mg *mg, int slot, float *u, float *v, float *w, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg && u && v && w, "bad arguments");
    for_each_rank(mg, [&](int r) { MG_CALL(gfs_get_field_layers(mg->ctx[r], slot, u, v, w, mg->k0[r], mg->k1[r] - mg->k0[r], &e_)); });
    // gfs_get_field_layers moves one extra w face layer per rank: the upper rank's copy of a shared layer is the owner's;
    // ranks ran concurrently, so rewrite the shared layers in rank order
    for (int r = 1; r < mg->n; r++)      // w face layer k0[r] belongs to rank r: fetch exactly that layer again (zero u, v layers)
        MG_CALL(gfs_get_field_layers(mg->ctx[r], slot, u, v, w, mg->k0[r], 0, &e_));
    MG_END()
}

void gfs_mg_get_material(gfs_mg *mg, uint8_t *material, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg && material, "bad arguments");
    for_each_rank(mg, [&](int r) { MG_CALL(gfs_get_material_layers(mg->ctx[r], material, mg->k0[r], mg->k1[r] - mg->k0[r], &e_)); });
    MG_END()
}

/* Distribute n particles over the slabs: cuts are chosen so that every GPU gets about the same number of particles
 * (cell layer of a particle: floor(z / dx) in double, Grid3d::positionToGridIndex, src/grid3d.h:58-63; particles outside
 * the grid in z go to the nearest slab), then each rank uploads its share.  (Re)builds the exchange plans. */
void gfs_mg_scatter_particles(gfs_mg *mg, const gfs_marker_particle_t *particles, int64_t n, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg && n >= 0 && (n == 0 || particles), "bad arguments");
    const double inv = 1.0 / mg->dx;
    std::vector<double> counts(mg->K, 0.0);
    std::vector<int> layer((size_t)n);
    for (int64_t p = 0; p < n; p++) {
        double kz = std::floor((double)particles[p].position.z * inv);
        int k = kz != kz ? 0 : (kz < 0 ? 0 : (kz >= mg->K ? mg->K - 1 : (int)kz));
        layer[(size_t)p] = k;
        counts[k] += 1.0;
    }
    weighted_cuts(counts, mg->n, std::max(4, mg->halo), mg->k0, mg->k1);
    std::vector<int> owner(mg->K);
    for (int r = 0; r < mg->n; r++) for (int k = mg->k0[r]; k < mg->k1[r]; k++) owner[k] = r;
    std::vector<std::vector<gfs_marker_particle_t>> share(mg->n);
    for (int r = 0; r < mg->n; r++) {
        double c = 0;
        for (int k = mg->k0[r]; k < mg->k1[r]; k++) c += counts[k];
        share[r].reserve((size_t)c);
    }
    for (int64_t p = 0; p < n; p++) share[owner[layer[(size_t)p]]].push_back(particles[p]);
    int64_t cap = 4096;
    for (int r = 0; r < mg->n; r++) cap = std::max<int64_t>(cap, (int64_t)share[r].size() / 4);
    connect(mg, cap);
    for_each_rank(mg, [&](int r) {
        MG_CALL(gfs_set_particles(mg->ctx[r], share[r].empty() ? nullptr : share[r].data(), (int64_t)share[r].size(), &e_));
        // room for arrivals and every scratch buffer now: no allocation (a device-wide synchronisation) inside the substeps
        MG_CALL(gfs_reserve(mg->ctx[r], (int64_t)share[r].size() + (int64_t)share[r].size() / 4 + 2 * cap, &e_));
    });
    MG_END()
}

int64_t gfs_mg_num_particles(gfs_mg *mg, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg, "null group");
    int64_t total = 0;
    for (gfs_context *c : mg->ctx) { int e = GFS_SUCCESS; total += gfs_num_particles(c, &e); if (e != GFS_SUCCESS) throw MgError{gfs_get_error_message()}; }
    return total;
    MG_END(-1)
}

/* all particles of all ranks, rank after rank (caller-allocated: gfs_mg_num_particles entries) */
void gfs_mg_gather_particles(gfs_mg *mg, gfs_marker_particle_t *particles, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg && particles, "bad arguments");
    std::vector<int64_t> off(mg->n + 1, 0);
    for (int r = 0; r < mg->n; r++) { int e = GFS_SUCCESS; off[r + 1] = off[r] + gfs_num_particles(mg->ctx[r], &e); }
    for_each_rank(mg, [&](int r) {
        if (off[r + 1] > off[r]) MG_CALL(gfs_get_particles(mg->ctx[r], particles + off[r], &e_));
    });
    MG_END()
}

/* One substep on every GPU: gfs_comm_substep per rank, each on its own host thread (index sort, all-ranks scale, splat,
 * C1 + C2 exchange with both neighbours, finalize + assembly, G2P + RK with fused migration, arrivals appended).
 * moved2[0] / [1]: particles that changed GPU (sent / received, summed over ranks). */
void gfs_mg_substep(gfs_mg *mg, double dt, double ratio_picflip, int rk_order, int interp, int arith, int64_t *moved2, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg && mg->connected, "gfs_mg_scatter_particles first");
    std::vector<int64_t> moved(2 * mg->n, 0);
    if (mg->n == 1) {
        MG_CALL(gfs_substep(mg->ctx[0], dt, ratio_picflip, rk_order, interp, arith, &e_));
    } else {
        for_each_rank(mg, [&](int r) {
            MG_CALL(gfs_comm_substep(mg->ctx[r], dt, ratio_picflip, rk_order, interp, arith, r > 0, r + 1 < mg->n, &moved[2 * r], &e_));
        });
    }
    if (moved2) { moved2[0] = moved2[1] = 0; for (int r = 0; r < mg->n; r++) { moved2[0] += moved[2 * r]; moved2[1] += moved[2 * r + 1]; } }
    MG_END()
}

/* gfs_state_hash summed over the ranks (mod 2^64): equals the single-GPU hash of the same state */
void gfs_mg_state_hash(gfs_mg *mg, uint64_t *out5, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg && out5, "bad arguments");
    for (int i = 0; i < 5; i++) out5[i] = 0;
    for (gfs_context *c : mg->ctx) {
        uint64_t h[5];
        MG_CALL(gfs_state_hash(c, h, &e_));
        for (int i = 0; i < 5; i++) out5[i] += h[i];
    }
    MG_END()
}

void gfs_mg_sync(gfs_mg *mg, int *err) {
    MG_BEGIN
    MG_REQUIRE(mg, "null group");
    for (gfs_context *c : mg->ctx) MG_CALL(gfs_sync(c, &e_));
    MG_END()
}

}  // extern "C"
