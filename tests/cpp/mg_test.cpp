// mg_test.cpp -- a C++11 host shards the transfer path over several GPUs through the native group API (gfs_mg_*), no Python,
// no NCCL: the sharded run must leave the very state the single-GPU run leaves (order- and distribution-independent
// 64-bit state hashes: material, P2G u / v / w, particle set) after every substep, trilinear and tricubic.
//
//   mg_test [nslabs] [device ...]        devices default to 0,1,2,... modulo the GPU count; a GPU may be named twice
//                                        (several slabs on one GPU: how the single-GPU test box runs this)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gfs_b200.h"

static uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }
static float unit(uint32_t &s) { return (float)(lcg(s) >> 8) * (1.0f / 16777216.0f); }

#define CHECK(call) do { int err_ = GFS_SUCCESS; call; if (err_ != GFS_SUCCESS) { \
    fprintf(stderr, "FAILED %s: %s | %s\n", #call, gfs_get_error_message(), gfs_mg_get_error_message()); return 1; } } while (0)

int main(int argc, char **argv) {
    const int nslabs = argc > 1 ? atoi(argv[1]) : 2;
    std::vector<int> devices;
    for (int a = 2; a < argc; a++) devices.push_back(atoi(argv[a]));
    const int I = 32, J = 24, K = 48;
    const double dx = 0.25;
    // border solids (src/fluidsimulation.cpp:1191-1213) + a solid block inside
    std::vector<uint8_t> material((size_t)I * J * K, GFS_AIR);
    for (int k = 0; k < K; k++) for (int j = 0; j < J; j++) for (int i = 0; i < I; i++) {
        const bool border = i == 0 || j == 0 || k == 0 || i == I - 1 || j == J - 1 || k == K - 1;
        const bool block = i >= 20 && i < 24 && j >= 2 && j < 6 && k >= 30 && k < 34;
        if (border || block) material[(size_t)i + I * ((size_t)j + (size_t)J * k)] = GFS_SOLID;
    }
    // 8 jittered particles per cell of a pool + a blob, velocities from a smooth function
    std::vector<gfs_marker_particle_t> particles;
    uint32_t seed = 12345u;
    for (int k = 1; k < K - 1; k++) for (int j = 1; j < J - 1; j++) for (int i = 1; i < I - 1; i++) {
        if (material[(size_t)i + I * ((size_t)j + (size_t)J * k)] == GFS_SOLID) continue;
        const bool pool = j < 10, blob = (i - 12) * (i - 12) + (j - 16) * (j - 16) + (k - 24) * (k - 24) < 30;
        if (!pool && !blob) continue;
        for (int s = 0; s < 8; s++) {
            gfs_marker_particle_t p;
            p.position.x = (float)((i + 0.25 + 0.5 * (s & 1) + 0.05 * (unit(seed) - 0.5)) * dx);
            p.position.y = (float)((j + 0.25 + 0.5 * ((s >> 1) & 1) + 0.05 * (unit(seed) - 0.5)) * dx);
            p.position.z = (float)((k + 0.25 + 0.5 * ((s >> 2) & 1) + 0.05 * (unit(seed) - 0.5)) * dx);
            p.velocity.x = 0.7f * sinf(p.position.y); p.velocity.y = -0.4f * cosf(p.position.z); p.velocity.z = 0.9f + 0.3f * sinf(p.position.x);
            particles.push_back(p);
        }
    }
    const int64_t n = (int64_t)particles.size();
    // fields: smooth + a +z drift so that particles cross the slab cuts
    std::vector<float> fu((size_t)(I + 1) * J * K), fv((size_t)I * (J + 1) * K), fw((size_t)I * J * (K + 1));
    for (size_t q = 0; q < fu.size(); q++) fu[q] = 0.3f * sinf(0.01f * (float)q);
    for (size_t q = 0; q < fv.size(); q++) fv[q] = 0.2f * cosf(0.013f * (float)q);
    for (size_t q = 0; q < fw.size(); q++) fw[q] = 0.8f + 0.1f * sinf(0.007f * (float)q);
    std::vector<float> su(fu), sv(fv), sw(fw);
    for (auto &x : su) x *= 0.9f; for (auto &x : sv) x *= 0.9f; for (auto &x : sw) x *= 0.9f;
    const double dt = 0.5 * dx;
    int rc = 0;
    for (int interp = 0; interp < 2; interp++) {
        int err = GFS_SUCCESS;
        const int halo = gfs_slab_halo_cells(interp, 1.0 * dt, dx, &err);
        // ---- single GPU
        gfs_context *one = gfs_create(devices.empty() ? 0 : devices[0], NULL, &err);
        if (err != GFS_SUCCESS) { fprintf(stderr, "gfs_create: %s\n", gfs_get_error_message()); return 1; }
        CHECK(gfs_domain_init(one, I, J, K, dx, &err_));
        CHECK(gfs_set_material(one, material.data(), &err_));
        CHECK(gfs_set_particles(one, particles.data(), n, &err_));
        CHECK(gfs_set_field(one, GFS_FIELD_NEW, fu.data(), fv.data(), fw.data(), &err_));
        CHECK(gfs_set_field(one, GFS_FIELD_SAVED, su.data(), sv.data(), sw.data(), &err_));
        // ---- the group
        gfs_mg *mg = gfs_mg_create(nslabs, devices.empty() ? NULL : devices.data(), I, J, K, dx, halo, &err);
        if (err != GFS_SUCCESS) { fprintf(stderr, "gfs_mg_create: %s\n", gfs_mg_get_error_message()); return 1; }
        CHECK(gfs_mg_set_material(mg, material.data(), &err_));
        CHECK(gfs_mg_scatter_particles(mg, particles.data(), n, &err_));
        CHECK(gfs_mg_set_field(mg, GFS_FIELD_NEW, fu.data(), fv.data(), fw.data(), &err_));
        CHECK(gfs_mg_set_field(mg, GFS_FIELD_SAVED, su.data(), sv.data(), sw.data(), &err_));
        int64_t moved_total = 0;
        for (int step = 0; step < 4; step++) {
            CHECK(gfs_substep(one, dt, 0.05f, 4, interp, GFS_FAST, &err_));
            int64_t moved[2] = {0, 0};
            CHECK(gfs_mg_substep(mg, dt, 0.05f, 4, interp, GFS_FAST, moved, &err_));
            moved_total += moved[0];
            uint64_t h1[5], hn[5];
            CHECK(gfs_state_hash(one, h1, &err_));
            CHECK(gfs_mg_state_hash(mg, hn, &err_));
            bool same = true;
            for (int q = 0; q < 5; q++) same = same && h1[q] == hn[q];
            printf("interp %d step %d: %lld particles, %lld crossed a cut, hashes %s (%016llx %016llx %016llx %016llx %016llx)\n", interp, step,
                   (long long)gfs_mg_num_particles(mg, &err), (long long)moved[0], same ? "equal" : "DIFFER",
                   (unsigned long long)hn[0], (unsigned long long)hn[1], (unsigned long long)hn[2], (unsigned long long)hn[3], (unsigned long long)hn[4]);
            if (!same) rc = 1;
        }
        if (gfs_mg_num_particles(mg, &err) != n) { fprintf(stderr, "particle count changed\n"); rc = 1; }
        if (nslabs > 1 && moved_total == 0) { fprintf(stderr, "no particle crossed a cut: the test does not test migration\n"); rc = 1; }
        // gathered fields and material equal the single-GPU arrays bit for bit
        std::vector<float> a((size_t)(I + 1) * J * K), b((size_t)I * (J + 1) * K), c((size_t)I * J * (K + 1)), a1(a.size()), b1(b.size()), c1(c.size());
        std::vector<uint8_t> m(material.size()), m1(material.size());
        CHECK(gfs_mg_get_field(mg, GFS_FIELD_P2G, a.data(), b.data(), c.data(), &err_));
        CHECK(gfs_get_field(one, GFS_FIELD_P2G, a1.data(), b1.data(), c1.data(), &err_));
        CHECK(gfs_mg_get_material(mg, m.data(), &err_));
        CHECK(gfs_get_material(one, m1.data(), &err_));
        if (memcmp(a.data(), a1.data(), a.size() * 4) || memcmp(b.data(), b1.data(), b.size() * 4) || memcmp(c.data(), c1.data(), c.size() * 4) || m != m1) {
            fprintf(stderr, "gathered P2G fields / material differ from the single-GPU arrays\n"); rc = 1;
        }
        CHECK(gfs_mg_destroy(mg, &err_));
        CHECK(gfs_destroy(one, &err_));
    }
    printf(rc == 0 ? "MG_TEST_OK\n" : "MG_TEST_FAILED\n");
    return rc;
}
