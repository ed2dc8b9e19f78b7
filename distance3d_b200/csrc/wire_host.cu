// Host side of the compact wire format (include/d3d_b200.h d3d_unpack_colliders): packs a
// structure-of-arrays collider set that lives in HOST memory into the type-specific records
// that travel over PCIe.  Plain C++ on the host cores (std::thread), no CUDA calls: this is the
// staging step a caller runs before stream.GjkDistanceStream.submit.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/d3d_b200.h"
#include "d3d_common.cuh"

namespace {

inline int wire_doubles_host(int type) {
    static const int size_of[D3D_NUM_TYPES] = {4, 14, 16, 15, 14, 1, 13, 7, 11, 14};
    return (type >= 0 && type < D3D_NUM_TYPES) ? size_of[type] : -1;
}

void pack_range(const d3d_colliders *c, int64_t begin, int64_t end, uint8_t *wire_type, const int32_t *wire_off,
                double *wire) {
    for (int64_t i = begin; i < end; ++i) {
        const int t = c->type[i];
        const double *T = c->pose + 16 * i, *p = c->param + 3 * i;
        double *r = wire + wire_off[i];
        wire_type[i] = (uint8_t)t;
        int32_t range[2] = {c->vert_off[i], c->vert_len[i]};
        switch (t) {
        case D3D_SPHERE: r[0] = T[3]; r[1] = T[7]; r[2] = T[11]; r[3] = p[0]; break;
        case D3D_CAPSULE: case D3D_CYLINDER: case D3D_CONE:
            memcpy(r, T, 12 * sizeof(double)); r[12] = p[0]; r[13] = p[1]; break;
        case D3D_ELLIPSOID: memcpy(r, T, 12 * sizeof(double)); r[12] = p[0]; r[13] = p[1]; r[14] = p[2]; break;
        case D3D_BOX:
            memcpy(r, T, 12 * sizeof(double)); r[12] = p[0]; r[13] = p[1]; r[14] = p[2];
            memcpy(r + 15, range, 8);
            break;
        case D3D_HULL: memcpy(r, range, 8); break;
        case D3D_MESH: memcpy(r, T, 12 * sizeof(double)); memcpy(r + 12, range, 8); break;
        case D3D_DISK:
            r[0] = T[3]; r[1] = T[7]; r[2] = T[11]; r[3] = T[2]; r[4] = T[6]; r[5] = T[10]; r[6] = p[0];
            break;
        case D3D_ELLIPSE:
            r[0] = T[3]; r[1] = T[7]; r[2] = T[11]; r[3] = T[0]; r[4] = T[4]; r[5] = T[8];
            r[6] = T[1]; r[7] = T[5]; r[8] = T[9]; r[9] = p[0]; r[10] = p[1];
            break;
        }
    }
}

}  // namespace

extern "C" {

int64_t d3d_wire_size(const int32_t *type, int64_t n) {
    if (n < 0 || (n > 0 && !type)) { d3d_set_error("d3d_wire_size: null argument"); return -1; }
    // large sets: the host threads share the pass (8 M colliders took 15 ms on one core, a third of
    // the whole packing step)
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(std::thread::hardware_concurrency(), 32u),
                                                                 (n + (1 << 18) - 1) >> 18));
    std::vector<int64_t> total(nt, 0), bad(nt, -1);
    auto work = [&](int k) {
        int64_t acc = 0;
        for (int64_t i = n * k / nt; i < n * (k + 1) / nt; ++i) {
            int s = wire_doubles_host(type[i]);
            if (s < 0) { if (bad[k] < 0) bad[k] = i; s = 0; }
            acc += s;
        }
        total[k] = acc;
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back(work, k);
        for (auto &t : th) t.join();
    }
    int64_t sum = 0;
    for (int k = 0; k < nt; ++k) {
        if (bad[k] >= 0) {
            d3d_set_error("d3d_wire_size: unknown collider type %d at %lld", type[bad[k]], (long long)bad[k]);
            return -1;
        }
        sum += total[k];
    }
    return sum;
}

int d3d_pack_wire_host(const d3d_colliders *c, uint8_t *wire_type, int32_t *wire_off, double *wire,
                       int n_threads) {
    if (!c || c->n == 0) return 0;
    if (!wire_type || !wire_off || !wire || !c->type || !c->pose || !c->param || !c->vert_off || !c->vert_len)
        return d3d_set_error("d3d_pack_wire_host: null argument");
    const int64_t n = c->n;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, (n + 65535) / 65536));
    std::vector<int64_t> chunk_total(nt + 1, 0);
    auto bounds = [&](int k) { return n * k / nt; };
    {   // pass 1: offsets inside every chunk, chunk totals
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k)
            th.emplace_back([&, k] {
                int64_t acc = 0;
                for (int64_t i = bounds(k); i < bounds(k + 1); ++i) {
                    wire_off[i] = (int32_t)acc;
                    int s = wire_doubles_host(c->type[i]);
                    acc += s < 0 ? 0 : s;
                }
                chunk_total[k + 1] = acc;
            });
        for (auto &t : th) t.join();
    }
    for (int k = 0; k < nt; ++k) chunk_total[k + 1] += chunk_total[k];
    if (chunk_total[nt] > 0x7fffffff) return d3d_set_error("d3d_pack_wire_host: more than 2^31-1 doubles; split the batch");
    {   // pass 2: global offsets + the records
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k)
            th.emplace_back([&, k] {
                const int32_t base = (int32_t)chunk_total[k];
                for (int64_t i = bounds(k); i < bounds(k + 1); ++i) wire_off[i] += base;
                pack_range(c, bounds(k), bounds(k + 1), wire_type, wire_off, wire);
            });
        for (auto &t : th) t.join();
    }
    return 0;
}

}  // extern "C"
