// TEST-ONLY host harness for haslr_b200/csrc/coords_core.cuh (the per-edge core is __host__ __device__): the same four
// steps as k4_edge_coords, serially — key lists, sort, the two sweeps on bitmasks, one walk per member of both best sets.
// Never shipped.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../haslr_b200/csrc/coords_core.cuh"

using namespace hgpu;

extern "C" int coordshost_edge_coords(uint32_t n_edges, const uint8_t* edge_rev, const uint32_t* supp_off, const hgpu_edge_supp* supp,
                                      const hgpu_cl_elem* elems, const uint32_t* cl_read_off, const uint32_t* read_len,
                                      const uint8_t* hit_is_rev, const uint32_t* cg_off, const uint32_t* cg_ops,
                                      hgpu_edge_coord* out_edge, hgpu_supp_coord* out_supp) {
    CoordIn in{supp, elems, cl_read_off, read_len, hit_is_rev, cg_off, cg_ops};
    for (uint32_t e = 0; e < n_edges; ++e) {
        const uint32_t b = supp_off[e], n = supp_off[e + 1] - b;
        const uint32_t rev1 = edge_rev[e] & 1u, rev2 = (edge_rev[e] >> 1) & 1u;
        const hgpu_edge_supp* es = supp + b;
        std::vector<uint64_t> k[4];
        for (uint32_t i = 0; i < n; ++i) {
            const hgpu_cl_elem& h = k4_elem(in, es[i], true);
            const hgpu_cl_elem& t = k4_elem(in, es[i], false);
            k[0].push_back(k4_key(h.t_start, i)); k[1].push_back(k4_key(h.t_end, i));
            k[2].push_back(k4_key(t.t_start, i)); k[3].push_back(k4_key(t.t_end, i));
        }
        for (auto& v : k) std::sort(v.begin(), v.end());
        std::vector<uint32_t> m1((n + 31) / 32 + 1, 0), m2((n + 31) / 32 + 1, 0);
        uint32_t i1lo = 0, i1hi = 0, i2lo = 0, i2hi = 0;
        k4_best_interval(k[0].data(), k[1].data(), n, true, m1.data(), &i1lo, &i1hi);
        k4_best_interval(k[2].data(), k[3].data(), n, false, m2.data(), &i2lo, &i2hi);
        const uint32_t c1 = rev1 == 0 ? i1hi - 1 : i1lo, c2 = rev2 == 0 ? i2lo : i2hi - 1;
        uint32_t n_best = 0, n_cns = 0;
        for (uint32_t i = 0; i < n; ++i) {
            hgpu_supp_coord o;
            o.lr_start = -1; o.lr_end = -1; o.lr_strand = 0; o.in_best = 0;
            if (((m1[i >> 5] & m2[i >> 5]) >> (i & 31u)) & 1u) {
                k4_walk(in, es[i], rev1, rev2, c1, c2, &o);
                ++n_best;
                if (o.lr_start != -1 && o.lr_end != -1) ++n_cns;
            }
            out_supp[b + i] = o;
        }
        out_edge[e] = hgpu_edge_coord{i1lo, i1hi, i2lo, i2hi, c1, c2, n_best, n_cns};
    }
    return 0;
}
