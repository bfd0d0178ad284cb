// host_plan.cpp — see host_plan.hpp.  Host-only; no CUDA.
#include "host_plan.hpp"

#include <algorithm>
#include <array>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <unordered_map>

namespace mmdgpu {

namespace {

constexpr double kEpsD = 1e-7;  // mmd_math_const_eps (L/util/math.inl:24), a double
constexpr double kPiD = 3.141592653589793238462643383279502884;
constexpr uint32_t kMaxNodes = 1u << 20;
constexpr uint32_t kMaxBones = 8191;  // 13-bit ids in the packed vertex stream

mmdgpu_status fail(std::string& err, mmdgpu_status s, const std::string& msg) {
    err = msg;
    return s;
}

}  // namespace

// --------------------------------------------------------------------------------------------------
// Bezier::SetC / presample / interpolate, L/util/math_impl.inl:1393-1428, control points scaled as
// VmdReader does (L/reader/vmd_reader_impl.inl:29-37).  fp32 throughout; built with -ffp-contract=off.
bool bezier_table(const int8_t c[4], float table[32]) {
    const float r = 1.0f / 127.0f;
    const float c0x = (float(c[0]) * r) * 3.0f, c0y = (float(c[1]) * r) * 3.0f;
    const float c1x = (float(c[2]) * r) * 3.0f, c1y = (float(c[3]) * r) * 3.0f;
    if (c0x == c0y && c1x == c1y) return true;
    for (int i = 0; i < 32; ++i) {
        const float x = float(i) / 31.0f;
        float lo = 0.0f, hi = 1.0f, mid = 0.0f, rm, m;
        for (int it = 0; it < 32; ++it) {
            mid = (lo + hi) * 0.5f;
            rm = 1.0f - mid;
            m = mid * (rm * (rm * c0x + mid * c1x) + mid * mid);
            if (std::fabs(m - x) < float(kEpsD)) break;
            if (m > x) hi = mid; else lo = mid;
        }
        rm = 1.0f - mid;
        table[i] = mid * (rm * (rm * c0y + mid * c1y) + mid * mid);
    }
    return false;
}

// --------------------------------------------------------------------------------------------------
// Round 1's tile order, kept behind MMDGPU_TILE_PAIRING=0 for A/B runs (pair_tile_order below is the default): second step
// after the stable sort by (skinning type, morph entry count, PMX index).
//
// The skinning kernel scatters a group's 32 results into a shared-memory staging tile at 12 bytes x PMX index, so two
// lanes of one group whose indices are congruent mod 32 hit the same banks (3 is coprime to 32): a group costs as many
// wavefronts per store as its most frequent residue.  After the sort a group is an irregular subset of the tile and its
// residues collide like random numbers (3.2 wavefronts per store on the 1 M-vertex benchmark model, measured with ncu).
// Hill climbing over swaps of two vertices of the SAME skinning type between groups lowers
//     cost(group) = 6 x sum over its types (worst residue multiplicity) + 2 x morph rounds (largest entry count)
// (six stores per vertex and slot against one extra morph round for the group; lanes of different types run in
// different branches, so they never store together).  Types never mix more than the sort left them.
template <class TypeOf, class CountOf>
void refine_tile_order(std::vector<uint32_t>& order, const TypeOf& type_of, const CountOf& count_of) {
    constexpr uint32_t G = kTileGroups, kTypes = 5;
    // in units of 1/8 wavefront; the sum of squared multiplicities breaks the plateaus of the worst-multiplicity term
    constexpr int kWConflict = 48, kWRound = 16, kWSpread = 1;
    struct Bucket {               // one (group, type): residue histogram and how many residues sit at each multiplicity
        uint8_t hist[32];
        uint8_t at[33];
        uint8_t mx;
    };
    struct Group {
        Bucket b[kTypes];
        uint32_t top1, n_top1, top2;   // largest entry count, how many members have it, the next smaller count
    };
    std::vector<Group> gr(G);
    std::vector<uint8_t> ty(kTileVerts);
    std::vector<uint32_t> cnt(kTileVerts);
    for (uint32_t i = 0; i < kTileVerts; ++i) { ty[i] = uint8_t(std::min<uint32_t>(type_of(i), kTypes - 1)); cnt[i] = count_of(i); }
    auto rebuild = [&](uint32_t g) {
        Group& Gp = gr[g];
        std::memset(&Gp, 0, sizeof Gp);
        for (uint32_t t = 0; t < kTypes; ++t) Gp.b[t].at[0] = 32;
        for (uint32_t l = 0; l < 32; ++l) {
            const uint32_t v = order[g * 32 + l];
            Bucket& B = Gp.b[ty[v]];
            const uint8_t h = B.hist[v & 31u]++;
            B.at[h]--; B.at[h + 1]++;
            if (h + 1 > B.mx) B.mx = uint8_t(h + 1);
            const uint32_t c = cnt[v];
            if (c > Gp.top1) { Gp.top2 = Gp.top1; Gp.top1 = c; Gp.n_top1 = 1; }
            else if (c == Gp.top1) Gp.n_top1++;
            else if (c > Gp.top2) Gp.top2 = c;
        }
    };
    for (uint32_t g = 0; g < G; ++g) rebuild(g);
    // worst multiplicity of a bucket after one vertex with residue `out` leaves and one with residue `in` joins
    auto mx_after = [](const Bucket& B, uint32_t out, uint32_t in) -> int {
        if (out == in) return B.mx;
        const int ho = B.hist[out];
        int m = (ho == B.mx && B.at[B.mx] == 1) ? B.mx - 1 : B.mx;
        return std::max(m, int(B.hist[in]) + 1);
    };
    auto rounds_after = [](const Group& Gp, uint32_t out, uint32_t in) -> uint32_t {
        const uint32_t rest = (out == Gp.top1 && Gp.n_top1 == 1) ? Gp.top2 : Gp.top1;
        return std::max(rest, in);
    };
    // rank range of every type (the sort made them contiguous)
    uint32_t t_begin[kTypes + 1];
    {
        uint32_t r = 0;
        for (uint32_t t = 0; t < kTypes; ++t) {
            t_begin[t] = r;
            while (r < kTileVerts && ty[order[r]] == t) ++r;
        }
        t_begin[kTypes] = kTileVerts;
        if (r != kTileVerts) return;   // not type-sorted (cannot happen): leave the order alone
    }
    for (int pass = 0; pass < 12; ++pass) {
        bool improved = false;
        for (uint32_t ra = 0; ra < kTileVerts; ++ra) {
            const uint32_t a = order[ra], ga = ra / 32, t = ty[a];
            const Bucket& A = gr[ga].b[t];
            if (A.hist[a & 31u] <= 1) continue;                   // collides with nobody
            int best = 0;
            uint32_t best_rb = 0;
            for (uint32_t rb = t_begin[t]; rb < t_begin[t + 1]; ++rb) {
                const uint32_t gb = rb / 32;
                if (gb == ga) continue;
                const uint32_t b = order[rb];
                const Bucket& B = gr[gb].b[t];
                const uint32_t ra5 = a & 31u, rb5 = b & 31u;
                if (ra5 == rb5) continue;
                // change of the sum of squares: a bucket loses one at `out` and gains one at `in`
                const int spread = 2 * (int(A.hist[rb5]) - int(A.hist[ra5])) + 2 + 2 * (int(B.hist[ra5]) - int(B.hist[rb5])) + 2;
                const int d = kWConflict * ((mx_after(A, ra5, rb5) - A.mx) + (mx_after(B, rb5, ra5) - B.mx)) + kWSpread * spread +
                              kWRound * (int(rounds_after(gr[ga], cnt[a], cnt[b])) - int(gr[ga].top1) +
                                         int(rounds_after(gr[gb], cnt[b], cnt[a])) - int(gr[gb].top1));
                if (d < best) { best = d; best_rb = rb; }
            }
            if (best < 0) {
                std::swap(order[ra], order[best_rb]);
                rebuild(ga);
                rebuild(best_rb / 32);
                improved = true;
            }
        }
        if (!improved) break;
    }
    // inside a group: by (type, PMX index) again, so that the layout does not depend on the order of the swaps.  (Merely
    // sorting the lanes by bone ids changes nothing: an LDS.128 takes its 2-wavefront path only when EVERY aligned lane pair
    // of the warp reads one cell - profiles/r02_smem_patterns.txt - which is what pair_tile_order arranges.)
    for (uint32_t g = 0; g < G; ++g)
        std::sort(order.begin() + g * 32, order.begin() + g * 32 + 32, [&](uint32_t x, uint32_t y) {
            return ty[x] != ty[y] ? ty[x] < ty[y] : x < y;
        });
}

// --------------------------------------------------------------------------------------------------
// Tile order, pairing form (default).  profiles/r02_smem_patterns.txt: an LDS.128 whose 32 lanes read unrelated 16-byte
// cells costs 4 wavefronts, but 2 when EVERY aligned lane pair (2k, 2k+1) of the instruction reads one cell.  The skinning
// kernel reads bone k of a vertex with six such loads per slot pair, so a 32-lane group whose lane pairs agree on bone k
// gets those six loads at half price.  Vertices of a tile use few bones (the point of the tile-local palette), so most of
// them can be paired with a vertex of the same skinning type and the same leading bone ids:
//   1. per type, vertices are matched hierarchically: first those that agree on all bone ids, the leftovers of each such
//      class on one id fewer, ... down to "same first bone", then arbitrarily (level 0); partners are chosen close in morph
//      entry count.  At most one vertex per type stays single; singles fill the last lane pairs of the tile.
//   2. a type's pairs are ordered by (level class, morph entry count) and cut into groups of 16 pairs in sequence; the
//      class of a (group, type) bucket is the lowest level in it, i.e. the number of leading bone ids whose loads take the
//      2-wavefront path.
//   3. hill climbing as in refine_tile_order (staging-scatter conflicts, morph rounds), with two kinds of moves that keep
//      every bucket's class: two vertices of one type trade places if each still agrees with its new partner on the
//      bucket's class many ids; two pairs trade places if each meets the other bucket's class.
struct TileVert {
    uint8_t ty, keep;
    uint16_t id[4];
    uint32_t cnt;
};
inline uint32_t prefix_match(const TileVert& a, const TileVert& b) {
    if (a.ty != b.ty) return 0;
    uint32_t k = 0;
    while (k < a.keep && a.id[k] == b.id[k]) ++k;
    return k;
}

void pair_tile_order(std::vector<uint32_t>& order, const std::vector<TileVert>& tv) {
    constexpr uint32_t G = kTileGroups, kTypes = 5, N = kTileVerts, kUnitsPerGroup = 16;
    constexpr int kWConflict = 48, kWRound = 16, kWSpread = 1;   // 1/8 wavefront per slot, as in refine_tile_order
    struct Unit { uint32_t a, b; uint8_t ty, level, cls; uint32_t cnt; };
    auto by_cnt = [&](uint32_t x, uint32_t y) { return tv[x].cnt != tv[y].cnt ? tv[x].cnt < tv[y].cnt : x < y; };
    std::vector<Unit> units;
    units.reserve(N / 2);
    std::vector<uint32_t> singles, rem, nxt;
    for (uint32_t t = 0; t < kTypes; ++t) {
        rem.clear();
        for (uint32_t i = 0; i < N; ++i) if (tv[i].ty == t) rem.push_back(i);
        if (rem.empty()) continue;
        const uint32_t keep = tv[rem[0]].keep;
        std::sort(rem.begin(), rem.end(), [&](uint32_t x, uint32_t y) {
            for (uint32_t k = 0; k < keep; ++k) if (tv[x].id[k] != tv[y].id[k]) return tv[x].id[k] < tv[y].id[k];
            return x < y;
        });
        const size_t u0 = units.size();
        for (uint32_t L = keep; L >= 1; --L) {
            nxt.clear();
            for (size_t i = 0; i < rem.size();) {
                size_t j = i + 1;
                while (j < rem.size() && prefix_match(tv[rem[i]], tv[rem[j]]) >= L) ++j;
                std::sort(rem.begin() + i, rem.begin() + j, by_cnt);
                size_t k = i;
                for (; k + 1 < j; k += 2)
                    units.push_back(Unit{std::min(rem[k], rem[k + 1]), std::max(rem[k], rem[k + 1]), uint8_t(t), uint8_t(L), 0,
                                         std::max(tv[rem[k]].cnt, tv[rem[k + 1]].cnt)});
                if (k < j) nxt.push_back(rem[k]);   // one leftover per class, classes stay in lexicographic order
                i = j;
            }
            rem.swap(nxt);
        }
        std::sort(rem.begin(), rem.end(), by_cnt);
        size_t k = 0;
        for (; k + 1 < rem.size(); k += 2)
            units.push_back(Unit{std::min(rem[k], rem[k + 1]), std::max(rem[k], rem[k + 1]), uint8_t(t), 0, 0,
                                 std::max(tv[rem[k]].cnt, tv[rem[k + 1]].cnt)});
        if (k < rem.size()) singles.push_back(rem[k]);
        // level-major order, then the class of each (group, type) bucket, then (class, count) order inside the type's range
        auto unit_less = [](const Unit& x, const Unit& y) { return x.cnt != y.cnt ? x.cnt < y.cnt : x.a < y.a; };
        std::sort(units.begin() + u0, units.end(), [&](const Unit& x, const Unit& y) {
            return x.level != y.level ? x.level > y.level : unit_less(x, y);
        });
        for (size_t i = u0; i < units.size();) {
            const size_t g_end = std::min(units.size(), (i / kUnitsPerGroup + 1) * kUnitsPerGroup);
            uint8_t lo = 255;
            for (size_t j = i; j < g_end; ++j) lo = std::min(lo, units[j].level);
            for (size_t j = i; j < g_end; ++j) units[j].cls = lo;
            i = g_end;
        }
        std::sort(units.begin() + u0, units.end(), [&](const Unit& x, const Unit& y) {
            return x.cls != y.cls ? x.cls > y.cls : unit_less(x, y);
        });
    }
    for (size_t k = 0; k + 1 < singles.size(); k += 2)   // N is even, so the singles pair up; mixed types, level 0
        units.push_back(Unit{singles[k], singles[k + 1], tv[singles[k]].ty, 0, 0, std::max(tv[singles[k]].cnt, tv[singles[k + 1]].cnt)});
    for (size_t u = 0; u < units.size(); ++u) { order[2 * u] = units[u].a; order[2 * u + 1] = units[u].b; }

    // ---- state of the hill climbing: per (group, type) residue histogram, required class; per group the lanes' counts
    struct Bucket { uint8_t hist[32]; uint8_t at[33]; uint8_t mx; uint8_t req; };
    struct Group { Bucket b[kTypes]; uint32_t top; uint32_t sorted_cnt[32]; };   // the lanes' counts, descending
    std::vector<Group> gr(G);
    auto rebuild = [&](uint32_t g) {
        Group& Gp = gr[g];
        uint8_t req[kTypes];
        for (uint32_t t = 0; t < kTypes; ++t) req[t] = Gp.b[t].req;
        std::memset(&Gp, 0, sizeof Gp);
        for (uint32_t t = 0; t < kTypes; ++t) { Gp.b[t].at[0] = 32; Gp.b[t].req = req[t]; }
        for (uint32_t l = 0; l < 32; ++l) {
            const uint32_t v = order[g * 32 + l];
            Bucket& B = Gp.b[tv[v].ty];
            const uint8_t h = B.hist[v & 31u]++;
            B.at[h]--; B.at[h + 1]++;
            if (h + 1 > B.mx) B.mx = uint8_t(h + 1);
            Gp.sorted_cnt[l] = tv[v].cnt;
        }
        std::sort(Gp.sorted_cnt, Gp.sorted_cnt + 32, std::greater<uint32_t>());
        Gp.top = Gp.sorted_cnt[0];
    };
    for (uint32_t g = 0; g < G; ++g) {
        for (uint32_t t = 0; t < kTypes; ++t) gr[g].b[t].req = 255;
        for (uint32_t u = g * kUnitsPerGroup; u < (g + 1) * kUnitsPerGroup; ++u) {
            const uint32_t a = order[2 * u], b = order[2 * u + 1];
            const uint8_t lv = uint8_t(prefix_match(tv[a], tv[b]));
            gr[g].b[tv[a].ty].req = std::min(gr[g].b[tv[a].ty].req, lv);
            gr[g].b[tv[b].ty].req = std::min(gr[g].b[tv[b].ty].req, lv);
        }
        for (uint32_t t = 0; t < kTypes; ++t) if (gr[g].b[t].req == 255) gr[g].b[t].req = 0;
        rebuild(g);
    }
    // worst multiplicity / change of the sum of squares of a bucket after n <= 2 residues lose and n gain one member
    auto bucket_delta = [](const Bucket& B, const uint32_t* out, const uint32_t* in, int n, int& d_mx, int& d_sq) {
        uint32_t r[4]; int d[4]; int m = 0;
        auto add = [&](uint32_t res, int dv) {
            for (int i = 0; i < m; ++i) if (r[i] == res) { d[i] += dv; return; }
            r[m] = res; d[m] = dv; ++m;
        };
        for (int i = 0; i < n; ++i) { add(out[i], -1); add(in[i], +1); }
        int best = 0; d_sq = 0;
        for (int i = 0; i < m; ++i) {
            const int o = B.hist[r[i]], nw = o + d[i];
            best = std::max(best, nw);
            d_sq += nw * nw - o * o;
        }
        int res_mx = best;
        for (int h = B.mx; h > best; --h) {
            int c = B.at[h];
            for (int i = 0; i < m; ++i) if (d[i] != 0 && B.hist[r[i]] == h) --c;
            if (c > 0) { res_mx = h; break; }
        }
        d_mx = res_mx - int(B.mx);
    };
    // largest count of group g when members with counts o0, o1 (o1 = ~0u: only one leaves) are replaced by counts c0, c1
    auto top_after = [&](uint32_t g, uint32_t o0, uint32_t o1, uint32_t c0, uint32_t c1) -> uint32_t {
        const uint32_t* sc = gr[g].sorted_cnt;
        bool gone0 = false, gone1 = (o1 == ~0u);
        uint32_t i = 0;
        for (; i < 32; ++i) {
            if (!gone0 && sc[i] == o0) { gone0 = true; continue; }
            if (!gone1 && sc[i] == o1) { gone1 = true; continue; }
            break;
        }
        return std::max(std::max(c0, c1), i < 32 ? sc[i] : 0u);
    };
    std::vector<uint32_t> ranks_of_type[kTypes];
    for (uint32_t r = 0; r < N; ++r) ranks_of_type[tv[order[r]].ty].push_back(r);   // moves keep the type of every rank

    for (int pass = 0; pass < 10; ++pass) {
        bool improved = false;
        // (a) two vertices of one type trade places
        for (uint32_t ra = 0; ra < N; ++ra) {
            const uint32_t a = order[ra], ga = ra / 32, t = tv[a].ty;
            const Bucket& A = gr[ga].b[t];
            if (A.hist[a & 31u] <= 1) continue;                   // collides with nobody
            const uint32_t pa = order[ra ^ 1u];
            int best = 0; uint32_t best_rb = 0;
            for (uint32_t rb : ranks_of_type[t]) {
                const uint32_t gb = rb / 32;
                if (gb == ga) continue;
                const uint32_t b = order[rb], pb = order[rb ^ 1u];
                const Bucket& B = gr[gb].b[t];
                if (prefix_match(tv[b], tv[pa]) < A.req || prefix_match(tv[a], tv[pb]) < B.req) continue;
                const uint32_t ra5 = a & 31u, rb5 = b & 31u;
                int dmA = 0, dsA = 0, dmB = 0, dsB = 0;
                if (ra5 != rb5) { bucket_delta(A, &ra5, &rb5, 1, dmA, dsA); bucket_delta(B, &rb5, &ra5, 1, dmB, dsB); }
                int d = kWConflict * (dmA + dmB) + kWSpread * (dsA + dsB);
                if (tv[a].cnt != tv[b].cnt)
                    d += kWRound * (int(top_after(ga, tv[a].cnt, ~0u, tv[b].cnt, 0)) - int(gr[ga].top) + int(top_after(gb, tv[b].cnt, ~0u, tv[a].cnt, 0)) - int(gr[gb].top));
                if (d < best) { best = d; best_rb = rb; }
            }
            if (best < 0) {
                std::swap(order[ra], order[best_rb]);
                rebuild(ga); rebuild(best_rb / 32);
                improved = true;
            }
        }
        // (b) two pairs of one type trade places
        for (uint32_t ua = 0; ua < N / 2; ++ua) {
            const uint32_t a0 = order[2 * ua], a1 = order[2 * ua + 1], ga = ua / kUnitsPerGroup, t = tv[a0].ty;
            if (tv[a1].ty != t) continue;
            const Bucket& A = gr[ga].b[t];
            if (A.hist[a0 & 31u] <= 1 && A.hist[a1 & 31u] <= 1) continue;
            const uint32_t la = prefix_match(tv[a0], tv[a1]);
            int best = 0; uint32_t best_ub = 0;
            for (uint32_t ub = 0; ub < N / 2; ++ub) {
                const uint32_t gb = ub / kUnitsPerGroup;
                if (gb == ga) continue;
                const uint32_t b0 = order[2 * ub], b1 = order[2 * ub + 1];
                if (tv[b0].ty != t || tv[b1].ty != t) continue;
                const Bucket& B = gr[gb].b[t];
                if (la < B.req || prefix_match(tv[b0], tv[b1]) < A.req) continue;
                const uint32_t outA[2] = {a0 & 31u, a1 & 31u}, outB[2] = {b0 & 31u, b1 & 31u};
                int dmA, dsA, dmB, dsB;
                bucket_delta(A, outA, outB, 2, dmA, dsA);
                bucket_delta(B, outB, outA, 2, dmB, dsB);
                int d = kWConflict * (dmA + dmB) + kWSpread * (dsA + dsB);
                d += kWRound * (int(top_after(ga, tv[a0].cnt, tv[a1].cnt, tv[b0].cnt, tv[b1].cnt)) - int(gr[ga].top) +
                                int(top_after(gb, tv[b0].cnt, tv[b1].cnt, tv[a0].cnt, tv[a1].cnt)) - int(gr[gb].top));
                if (d < best) { best = d; best_ub = ub; }
            }
            if (best < 0) {
                std::swap(order[2 * ua], order[2 * best_ub]);
                std::swap(order[2 * ua + 1], order[2 * best_ub + 1]);
                rebuild(ga); rebuild(best_ub / kUnitsPerGroup);
                improved = true;
            }
        }
        if (!improved) break;
    }
    // canonical order inside a group: pairs by (type, smaller PMX index), the smaller index in the even lane
    for (uint32_t g = 0; g < G; ++g) {
        std::array<std::pair<uint32_t, uint32_t>, kUnitsPerGroup> us;
        for (uint32_t u = 0; u < kUnitsPerGroup; ++u) {
            const uint32_t x = order[g * 32 + 2 * u], y = order[g * 32 + 2 * u + 1];
            const bool swap = tv[x].ty != tv[y].ty ? tv[x].ty > tv[y].ty : x > y;
            us[u] = swap ? std::make_pair(y, x) : std::make_pair(x, y);
        }
        std::sort(us.begin(), us.end(), [&](const std::pair<uint32_t, uint32_t>& x, const std::pair<uint32_t, uint32_t>& y) {
            if (tv[x.first].ty != tv[y.first].ty) return tv[x.first].ty < tv[y.first].ty;
            if (tv[x.second].ty != tv[y.second].ty) return tv[x.second].ty < tv[y.second].ty;
            return x.first < y.first;
        });
        for (uint32_t u = 0; u < kUnitsPerGroup; ++u) { order[g * 32 + 2 * u] = us[u].first; order[g * 32 + 2 * u + 1] = us[u].second; }
    }
}

mmdgpu_status build_plan(const mmdgpu_model_desc& d, const mmdgpu_options* opt, Plan& p, std::string& err) {
    const uint32_t nv = d.n_vertices, nb = d.n_bones, nm = d.n_morphs;
    p = Plan();
    p.nv = nv; p.nb = nb; p.nm = nm;
    p.extensions = opt && opt->extensions != 0;

    if (nb == 0) return fail(err, MMDGPU_ERR_INVALID_ARG, "model has no bones");
    if (nb > kMaxBones) return fail(err, MMDGPU_ERR_UNSUPPORTED, "more than 8191 bones (13-bit packed bone ids)");
    if (nv && (!d.position || !d.normal || !d.skin_type || !d.bone_id || !d.weight))
        return fail(err, MMDGPU_ERR_INVALID_ARG, "vertex arrays missing");
    if (!d.bone_position || !d.bone_parent || !d.bone_transform_level || !d.bone_flags)
        return fail(err, MMDGPU_ERR_INVALID_ARG, "bone arrays missing");
    if (nm && (!d.morph_type || !d.morph_entry_begin || !d.morph_entry_count))
        return fail(err, MMDGPU_ERR_INVALID_ARG, "morph arrays missing");

    // ---------------------------------------------------------------- bones (Poser::Poser, poser_impl.inl:30-105)
    p.bones.resize(nb);
    std::vector<int32_t> ik_of_bone(nb, -1);
    for (uint32_t b = 0; b < nb; ++b) {
        BoneStatic& s = p.bones[b];
        std::memset(&s, 0, sizeof s);
        const uint16_t fl = d.bone_flags[b];
        const int32_t par = d.bone_parent[b];
        for (int k = 0; k < 3; ++k) s.position[k] = d.bone_position[3 * b + k];
        s.link_slot = s.morph_slot = -1;
        s.append_parent = -1;
        if (par >= 0 && uint32_t(par) < nb) {
            s.parent = par;
            s.flags |= kHasParent;
            for (int k = 0; k < 3; ++k) s.local_offset[k] = d.bone_position[3 * b + k] - d.bone_position[3 * par + k];
        } else {
            s.parent = -1;
            for (int k = 0; k < 3; ++k) s.local_offset[k] = d.bone_position[3 * b + k];
        }
        if (fl & (MMDGPU_BONE_APPEND_ROTATE | MMDGPU_BONE_APPEND_TRANSLATE)) {
            const int32_t ap = d.bone_append_parent ? d.bone_append_parent[b] : -1;
            if (ap >= 0 && uint32_t(ap) < nb) {  // dropped otherwise, poser_impl.inl:51-57
                s.append_parent = ap;
                s.append_ratio = d.bone_append_ratio ? d.bone_append_ratio[b] : 0.0f;
                if (fl & MMDGPU_BONE_APPEND_ROTATE) s.flags |= kAppendRot;
                if (fl & MMDGPU_BONE_APPEND_TRANSLATE) s.flags |= kAppendTrans;
            }
        }
        if (fl & MMDGPU_BONE_POST_PHYSICS) s.flags |= kPostPhysics;
        if (fl & MMDGPU_BONE_HAS_IK) s.flags |= kHasIk;
    }
    // IK descriptors.  Link records keep the descriptor's order so PLAN_IK_* line up with ik_link_*.
    p.links.resize(d.n_ik_links);
    std::vector<uint8_t> link_seen(d.n_ik_links, 0);
    for (uint32_t b = 0; b < nb; ++b) {
        if (!(p.bones[b].flags & kHasIk)) continue;
        if (!d.ik_target || !d.ik_iterations || !d.ik_angle_limit || !d.ik_link_begin || !d.ik_link_count)
            return fail(err, MMDGPU_ERR_INVALID_ARG, "IK arrays missing");
        IkDesc k{};
        k.bone = int32_t(b);
        k.target = d.ik_target[b];
        if (k.target < 0 || uint32_t(k.target) >= nb)
            return fail(err, MMDGPU_ERR_BAD_INDEX, "IK target out of range at bone " + std::to_string(b));
        const uint32_t it = uint32_t(d.ik_iterations[b]);
        k.iterations = int32_t(std::min<uint32_t>(it, 256u));
        k.angle_limit = d.ik_angle_limit[b];
        k.link_begin = int32_t(d.ik_link_begin[b]);
        k.link_count = int32_t(d.ik_link_count[b]);
        if (uint64_t(d.ik_link_begin[b]) + d.ik_link_count[b] > d.n_ik_links)
            return fail(err, MMDGPU_ERR_BAD_INDEX, "IK link range out of range at bone " + std::to_string(b));
        if (k.link_count && (!d.ik_link_bone || !d.ik_link_has_limit || !d.ik_link_lo || !d.ik_link_hi))
            return fail(err, MMDGPU_ERR_INVALID_ARG, "IK link arrays missing");
        for (int32_t j = 0; j < k.link_count; ++j) {
            const uint32_t l = uint32_t(k.link_begin + j);
            if (link_seen[l]) return fail(err, MMDGPU_ERR_INVALID_ARG, "IK link record shared by two IK bones");
            link_seen[l] = 1;
            IkLink& L = p.links[l];
            std::memset(&L, 0, sizeof L);
            L.bone = d.ik_link_bone[l];
            if (L.bone < 0 || uint32_t(L.bone) >= nb)
                return fail(err, MMDGPU_ERR_BAD_INDEX, "IK link bone out of range at bone " + std::to_string(b));
            L.limited = d.ik_link_has_limit[l] != 0;
            L.order = 0;  // ORDER_YZX default (poser_impl.inl:64)
            L.fix = 0;    // FIX_NONE
            if (L.limited) {
                for (int c = 0; c < 3; ++c) {
                    const float lo = d.ik_link_lo[3 * l + c], hi = d.ik_link_hi[3 * l + c];
                    L.lo[c] = std::min(lo, hi);
                    L.hi[c] = std::max(lo, hi);
                }
                // poser_impl.inl:78-82 — float limits compared as double against -(pi)*0.5f
                if (double(L.lo[0]) > -kPiD * 0.5f && double(L.hi[0]) < kPiD * 0.5f) L.order = 1;       // ZXY
                else if (double(L.lo[1]) > -kPiD * 0.5f && double(L.hi[1]) < kPiD * 0.5f) L.order = 2;  // XYZ
                // poser_impl.inl:83-91 — float abs (main.cpp's TU), double compare
                const bool zx = double(std::fabs(L.lo[0])) < kEpsD && double(std::fabs(L.hi[0])) < kEpsD;
                const bool zy = double(std::fabs(L.lo[1])) < kEpsD && double(std::fabs(L.hi[1])) < kEpsD;
                const bool zz = double(std::fabs(L.lo[2])) < kEpsD && double(std::fabs(L.hi[2])) < kEpsD;
                if (zx && zy && zz) L.fix = 4;
                else if (zy && zz) L.fix = 1;
                else if (zx && zz) L.fix = 2;
                else if (zx && zy) L.fix = 3;
            }
            p.bones[L.bone].flags |= kIsLink;
        }
        ik_of_bone[b] = int32_t(p.iks.size());
        p.iks.push_back(k);
    }
    // Nested solves: libmmd re-evaluates a solve's links and target with UpdateBoneTransform (poser_impl.inl:203-206, :303),
    // which re-enters the IK block when that bone has IK itself.  The device runs the same recursion to a fixed depth.
    // A cycle (an IK bone reached again from its own solve) makes libmmd recurse until the stack overflows: rejected.
    {
        const size_t nik = p.iks.size();
        if (nik > 0xFFFFu) return fail(err, MMDGPU_ERR_UNSUPPORTED, "more than 65535 IK bones");
        std::vector<std::vector<int32_t>> inner(nik);
        for (size_t i = 0; i < nik; ++i) {
            const IkDesc& k = p.iks[i];
            p.bones[size_t(k.bone)].flags |= uint32_t(i) << 16;      // bits 31:16 of an IK bone's flags: its IkDesc index
            for (int32_t j = 0; j < k.link_count; ++j) {
                const int32_t lb = p.links[size_t(k.link_begin + j)].bone;
                if (p.bones[size_t(lb)].flags & kHasIk) inner[i].push_back(ik_of_bone[size_t(lb)]);
            }
            if (p.bones[size_t(k.target)].flags & kHasIk) inner[i].push_back(ik_of_bone[size_t(k.target)]);
            if (!inner[i].empty()) p.ik_nested = true;
        }
        std::vector<int32_t> depth(nik, 0);      // 0 unvisited, -1 on the stack, > 0 levels of solves below and including this one
        struct Frame { int32_t ik; size_t next; };
        std::vector<Frame> stack;
        for (size_t root = 0; root < nik && p.ik_nested; ++root) {
            if (depth[root] != 0) continue;
            stack.push_back({int32_t(root), 0});
            depth[root] = -1;
            while (!stack.empty()) {
                Frame& f = stack.back();
                if (f.next < inner[size_t(f.ik)].size()) {
                    const int32_t c = inner[size_t(f.ik)][f.next++];
                    if (depth[size_t(c)] == -1)
                        return fail(err, MMDGPU_ERR_BAD_INDEX, "IK bones reach each other through their links / targets (libmmd would recurse forever)");
                    if (depth[size_t(c)] == 0) { depth[size_t(c)] = -1; stack.push_back({c, 0}); }
                    continue;
                }
                int32_t dmax = 0;
                for (int32_t c : inner[size_t(f.ik)]) dmax = std::max(dmax, depth[size_t(c)]);
                depth[size_t(f.ik)] = dmax + 1;
                if (dmax + 1 > kMaxIkDepth)
                    return fail(err, MMDGPU_ERR_UNSUPPORTED, "IK solves nested deeper than " + std::to_string(kMaxIkDepth) + " levels");
                stack.pop_back();
            }
        }
    }
    for (uint32_t b = 0; b < nb; ++b)
        if (p.bones[b].flags & kIsLink) {
            p.bones[b].link_slot = int32_t(p.link_bones.size());
            p.link_bones.push_back(int32_t(b));
        }

    // evaluation order: two lists, each std::sort'ed by (transform level, index) (poser_impl.inl:100-109, 500-510)
    for (uint32_t b = 0; b < nb; ++b) (p.bones[b].flags & kPostPhysics ? p.order_post : p.order_pre).push_back(int32_t(b));
    auto by_level = [&](int32_t a, int32_t b) {
        const size_t la = size_t(d.bone_transform_level[a]), lb = size_t(d.bone_transform_level[b]);
        if (la < lb) return true;
        if (la > lb) return false;
        return a < b;
    };
    std::sort(p.order_pre.begin(), p.order_pre.end(), by_level);
    std::sort(p.order_post.begin(), p.order_post.end(), by_level);

    // ---------------------------------------------------------------- program + wave schedule
    // Symbolic execution of PrePhysicsPosing / PostPhysicsPosing (poser_impl.inl:362-394): every op gets
    // the read / write sets of the state it touches; an op's wave is one past the last conflicting access.
    enum { V_TOT = 0, V_LOCAL = 1, V_IK = 2, V_PRE = 3, V_SKIN = 4, V_KINDS = 5 };
    auto var = [&](int kind, int32_t b) { return size_t(kind) * nb + size_t(b); };
    std::vector<int32_t> lastW(size_t(V_KINDS) * nb, -1), lastR(size_t(V_KINDS) * nb, -1);
    std::vector<uint8_t> need_reset(nb, 0);
    auto eval_sets = [&](int32_t b, std::vector<size_t>& R, std::vector<size_t>& W) {
        const BoneStatic& s = p.bones[b];
        if (s.flags & (kAppendRot | kAppendTrans)) R.push_back(var(V_TOT, s.append_parent));
        if (s.flags & kIsLink) { R.push_back(var(V_IK, b)); W.push_back(var(V_PRE, b)); }
        if (s.flags & kHasParent) R.push_back(var(V_LOCAL, s.parent));
        W.push_back(var(V_TOT, b));
        W.push_back(var(V_LOCAL, b));
    };
    int32_t phase_floor = 0, max_wave = -1;
    auto emit = [&](uint8_t kind, int32_t arg, std::vector<size_t>& R, std::vector<size_t>& W) {
        int32_t w = phase_floor;
        // reads that a write of the same op precedes do not count as "read before written"
        for (size_t v : R) {
            w = std::max(w, lastW[v] + 1);
            if (lastW[v] < 0 && (v / nb == V_TOT || v / nb == V_LOCAL)) need_reset[v % nb] = 1;
        }
        for (size_t v : W) w = std::max(w, std::max(lastW[v], lastR[v]) + 1);
        for (size_t v : R) lastR[v] = std::max(lastR[v], w);
        for (size_t v : W) lastW[v] = w;
        p.ops.push_back(Op{kind, arg});
        p.op_wave.push_back(w);
        max_wave = std::max(max_wave, w);
    };
    auto run_list = [&](const std::vector<int32_t>& list) {
        std::vector<size_t> R, W;
        for (int32_t b : list) {
            R.clear(); W.clear();
            eval_sets(b, R, W);
            emit(kOpEval, b, R, W);
            if (p.bones[b].flags & kHasIk) {
                R.clear(); W.clear();
                // Internal order of a solve (poser_impl.inl:199-206): reset ikR, re-evaluate links root-most first, then
                // the target - each followed by its own solve if that bone has IK.  State an inner step reads after an
                // earlier inner step wrote it is not an external read; the sets below are the external view.
                std::vector<size_t> wrote;
                auto ext_read = [&](size_t v) { if (std::find(wrote.begin(), wrote.end(), v) == wrote.end()) R.push_back(v); };
                std::function<void(int32_t)> solve_sets = [&](int32_t ki) {
                    const IkDesc& k = p.iks[size_t(ki)];
                    ext_read(var(V_LOCAL, k.bone));
                    auto add_eval = [&](int32_t x) {
                        std::vector<size_t> r2, w2;
                        eval_sets(x, r2, w2);
                        for (size_t v : r2) ext_read(v);
                        for (size_t v : w2) { W.push_back(v); wrote.push_back(v); }
                        if (p.bones[size_t(x)].flags & kHasIk) solve_sets(ik_of_bone[size_t(x)]);   // depth bounded above
                    };
                    for (int32_t j = 0; j < k.link_count; ++j) {
                        const size_t v = var(V_IK, p.links[k.link_begin + j].bone);
                        W.push_back(v); wrote.push_back(v);
                    }
                    for (int32_t j = k.link_count - 1; j >= 0; --j) add_eval(p.links[k.link_begin + j].bone);
                    add_eval(k.target);
                    for (int32_t j = 0; j < k.link_count; ++j) {
                        const BoneStatic& ls = p.bones[p.links[k.link_begin + j].bone];
                        if (ls.flags & kHasParent) ext_read(var(V_LOCAL, ls.parent));
                    }
                };
                solve_sets(ik_of_bone[b]);
                emit(kOpIk, ik_of_bone[b], R, W);
            }
        }
        for (int32_t b : list) {  // UpdateBoneSkinningMatrix (poser_impl.inl:320-326)
            R.clear(); W.clear();
            R.push_back(var(V_LOCAL, b));
            W.push_back(var(V_SKIN, b));
            emit(kOpSkin, b, R, W);
        }
    };
    run_list(p.order_pre);
    p.phase_split = max_wave + 1;
    phase_floor = p.phase_split;
    run_list(p.order_post);
    const int32_t n_waves = std::max(max_wave + 1, p.phase_split);
    p.wave_begin.assign(size_t(n_waves) + 1, 0);
    for (int32_t w : p.op_wave) p.wave_begin[size_t(w) + 1]++;
    for (int32_t w = 0; w < n_waves; ++w) p.wave_begin[w + 1] += p.wave_begin[w];
    p.wave_ops.resize(p.ops.size());
    {
        std::vector<int32_t> cur(p.wave_begin.begin(), p.wave_begin.end() - 1);
        for (size_t i = 0; i < p.ops.size(); ++i) p.wave_ops[size_t(cur[p.op_wave[i]]++)] = int32_t(i);
    }
    for (uint32_t b = 0; b < nb; ++b)
        if (need_reset[b]) p.reset_bones.push_back(int32_t(b));

    // ---------------------------------------------------------------- morph application slots
    // DFS expansion of UpdateMorphTransform's recursion (poser_impl.inl:328-360): node order == the order in
    // which libmmd applies morph data.
    for (uint32_t m = 0; m < nm; ++m) {
        const uint8_t t = d.morph_type[m];
        const uint64_t end = uint64_t(d.morph_entry_begin[m]) + d.morph_entry_count[m];
        uint32_t pool = 0xFFFFFFFFu;
        if (t == MMDGPU_MORPH_GROUP) pool = d.n_group_morph_entries;
        else if (t == MMDGPU_MORPH_VERTEX) pool = d.n_vertex_morph_entries;
        else if (t == MMDGPU_MORPH_BONE) pool = d.n_bone_morph_entries;
        else if (t >= MMDGPU_MORPH_UV && t <= MMDGPU_MORPH_EXT_UV4) pool = d.n_uv_morph_entries;
        else if (t == MMDGPU_MORPH_MATERIAL && d.material_morph_entries) pool = d.n_material_morph_entries;
        if (pool != 0xFFFFFFFFu && d.morph_entry_count[m] && end > pool)
            return fail(err, MMDGPU_ERR_BAD_INDEX, "morph entry range out of range at morph " + std::to_string(m));
    }
    {
        // Depth-first with an explicit stack (a crafted chain of hundreds of thousands of nested groups must not
        // overflow the host stack) and a nesting cap far above any real model: libmmd itself recurses once per level.
        constexpr int32_t kMaxGroupDepth = 64;
        std::vector<uint8_t> on_stack(nm, 0);
        struct Visit { uint32_t morph; int32_t node; uint32_t next_child; };
        std::vector<Visit> stack;
        auto enter_node = [&](uint32_t m, int32_t parent, float mult, int32_t depth) -> mmdgpu_status {
            if (p.node_morph.size() >= kMaxNodes) return fail(err, MMDGPU_ERR_UNSUPPORTED, "group morph expansion too large");
            if (depth > kMaxGroupDepth)
                return fail(err, MMDGPU_ERR_UNSUPPORTED, "group morphs nested deeper than " + std::to_string(kMaxGroupDepth));
            const int32_t me = int32_t(p.node_morph.size());
            p.node_morph.push_back(int32_t(m));
            p.node_parent.push_back(parent);
            p.node_mult.push_back(mult);
            p.node_depth.push_back(depth);
            if (d.morph_type[m] == MMDGPU_MORPH_GROUP) {
                on_stack[m] = 1;
                stack.push_back({m, me, 0u});
            }
            return MMDGPU_OK;
        };
        for (uint32_t m = 0; m < nm; ++m) {
            if (mmdgpu_status st = enter_node(m, -1, 1.0f, 0)) return st;
            while (!stack.empty()) {
                Visit& v = stack.back();
                if (v.next_child >= d.morph_entry_count[v.morph]) {
                    on_stack[v.morph] = 0;
                    stack.pop_back();
                    continue;
                }
                const mmdgpu_group_morph_entry& g = d.group_morph_entries[d.morph_entry_begin[v.morph] + v.next_child++];
                if (g.morph >= nm) return fail(err, MMDGPU_ERR_BAD_INDEX, "group morph child out of range");
                if (on_stack[g.morph]) return fail(err, MMDGPU_ERR_BAD_INDEX, "group morph cycle (libmmd would recurse forever)");
                const int32_t parent = v.node, depth = p.node_depth[size_t(v.node)] + 1;
                if (mmdgpu_status st = enter_node(g.morph, parent, g.rate, depth)) return st;  // may grow `stack`: v is dead
            }
        }
    }
    const size_t n_nodes = p.node_morph.size();
    {
        int32_t max_depth = -1;
        for (int32_t x : p.node_depth) max_depth = std::max(max_depth, x);
        p.depth_begin.assign(size_t(max_depth + 2), 0);
        for (int32_t x : p.node_depth) p.depth_begin[size_t(x) + 1]++;
        for (int32_t x = 0; x <= max_depth; ++x) p.depth_begin[x + 1] += p.depth_begin[x];
        p.nodes_by_depth.resize(n_nodes);
        std::vector<int32_t> cur(p.depth_begin.begin(), p.depth_begin.end() - 1);
        for (size_t i = 0; i < n_nodes; ++i) p.nodes_by_depth[size_t(cur[p.node_depth[i]]++)] = int32_t(i);
    }

    // per-vertex CSR: rows sorted by (application slot, entry order) because nodes are visited in order
    auto build_vertex_csr = [&](bool uv, std::vector<uint32_t>& row, std::vector<uint32_t>& node_of,
                                std::vector<float>& off) -> mmdgpu_status {
        row.assign(size_t(nv) + 1, 0);
        const int width = uv ? 4 : 3;
        auto is_mine = [&](uint8_t t) { return uv ? (t == MMDGPU_MORPH_UV) : (t == MMDGPU_MORPH_VERTEX); };
        for (size_t n = 0; n < n_nodes; ++n) {
            const uint32_t m = uint32_t(p.node_morph[n]);
            if (!is_mine(d.morph_type[m])) continue;
            for (uint32_t j = 0; j < d.morph_entry_count[m]; ++j) {
                const uint32_t v = uv ? d.uv_morph_entries[d.morph_entry_begin[m] + j].vertex
                                      : d.vertex_morph_entries[d.morph_entry_begin[m] + j].vertex;
                if (v >= nv) return fail(err, MMDGPU_ERR_BAD_INDEX, "morph vertex index out of range at morph " + std::to_string(m));
                {
                    // the device applies skipped morphs with rate 0 instead of branching; that is bit-identical to
                    // libmmd's skip only for finite offsets
                    const float* o = uv ? d.uv_morph_entries[d.morph_entry_begin[m] + j].offset
                                        : d.vertex_morph_entries[d.morph_entry_begin[m] + j].offset;
                    if (!std::isfinite(o[0]) || !std::isfinite(o[1]) || !std::isfinite(o[2]) || (uv && !std::isfinite(o[3])))
                        return fail(err, MMDGPU_ERR_INVALID_ARG, "non-finite morph offset at morph " + std::to_string(m));
                }
                row[size_t(v) + 1]++;
            }
        }
        for (uint32_t v = 0; v < nv; ++v) row[v + 1] += row[v];
        node_of.resize(row[nv]);
        off.resize(size_t(row[nv]) * width);
        std::vector<uint32_t> cur(row.begin(), row.end() - 1);
        for (size_t n = 0; n < n_nodes; ++n) {
            const uint32_t m = uint32_t(p.node_morph[n]);
            if (!is_mine(d.morph_type[m])) continue;
            for (uint32_t j = 0; j < d.morph_entry_count[m]; ++j) {
                const uint32_t e = d.morph_entry_begin[m] + j;
                const uint32_t v = uv ? d.uv_morph_entries[e].vertex : d.vertex_morph_entries[e].vertex;
                const float* o = uv ? d.uv_morph_entries[e].offset : d.vertex_morph_entries[e].offset;
                const uint32_t at = cur[v]++;
                node_of[at] = uint32_t(n);
                for (int k = 0; k < width; ++k) off[size_t(at) * width + k] = o[k];
            }
        }
        return MMDGPU_OK;
    };
    if (mmdgpu_status st = build_vertex_csr(false, p.csr_row, p.csr_node, p.csr_offset)) return st;
    if (p.extensions)
        if (mmdgpu_status st = build_vertex_csr(true, p.uv_row, p.uv_node, p.uv_offset)) return st;

    // bone morphs grouped by affected bone, application order inside a bone
    {
        std::vector<std::vector<BoneMorphEntry>> per_bone(nb);
        for (size_t n = 0; n < n_nodes; ++n) {
            const uint32_t m = uint32_t(p.node_morph[n]);
            if (d.morph_type[m] != MMDGPU_MORPH_BONE) continue;
            for (uint32_t j = 0; j < d.morph_entry_count[m]; ++j) {
                const mmdgpu_bone_morph_entry& e = d.bone_morph_entries[d.morph_entry_begin[m] + j];
                if (e.bone >= nb) return fail(err, MMDGPU_ERR_BAD_INDEX, "bone morph bone index out of range at morph " + std::to_string(m));
                BoneMorphEntry x;
                x.node = int32_t(n);
                std::memcpy(x.translation, e.translation, 12);
                std::memcpy(x.rotation, e.rotation, 16);
                per_bone[e.bone].push_back(x);
            }
        }
        p.bone_morph_row.push_back(0);
        for (uint32_t b = 0; b < nb; ++b) {
            if (per_bone[b].empty()) continue;
            p.bones[b].morph_slot = int32_t(p.morph_bones.size());
            p.morph_bones.push_back(int32_t(b));
            p.bone_morph_entries.insert(p.bone_morph_entries.end(), per_bone[b].begin(), per_bone[b].end());
            p.bone_morph_row.push_back(int32_t(p.bone_morph_entries.size()));
        }
    }

    // ---------------------------------------------------------------- chain-local images of the IK solves
    // (device design: hierarchy_flat_kernel runs a solve on a private copy of just the bones it touches)
    {
        p.ik_img_ok = !p.iks.empty() && !p.ik_nested;   // a nested solve touches bones outside its parent's image
        for (const IkDesc& k : p.iks) {
            std::vector<int32_t> bones;                 // image index -> global bone id
            std::vector<uint8_t> written;
            auto add = [&](int32_t b, bool wr) -> int32_t {
                for (size_t i = 0; i < bones.size(); ++i)
                    if (bones[i] == b) { written[i] |= uint8_t(wr); return int32_t(i); }
                bones.push_back(b); written.push_back(uint8_t(wr));
                return int32_t(bones.size() - 1);
            };
            std::vector<int32_t> evaluated;             // links and the target: eval_bone / set_local run on them
            for (int32_t j = 0; j < k.link_count; ++j) evaluated.push_back(p.links[size_t(k.link_begin + j)].bone);
            evaluated.push_back(k.target);
            for (int32_t b : evaluated) add(b, true);
            add(k.bone, false);
            for (int32_t b : evaluated) {               // what evaluating them reads
                const BoneStatic& s0 = p.bones[size_t(b)];
                if (s0.flags & kHasParent) add(s0.parent, false);
                if (s0.flags & (kAppendRot | kAppendTrans)) add(s0.append_parent, false);
            }
            std::vector<int32_t> lslots, mslots;
            auto slot_of = [](std::vector<int32_t>& v, int32_t x) -> int32_t {
                for (size_t i = 0; i < v.size(); ++i) if (v[i] == x) return int32_t(i);
                v.push_back(x);
                return int32_t(v.size() - 1);
            };
            IkImage I{};
            I.bones_begin = int32_t(p.ik_img_bones.size());
            I.n_bones = int32_t(bones.size());
            for (size_t i = 0; i < size_t(I.n_bones); ++i) {
                BoneStatic s1 = p.bones[size_t(bones[i])];
                const bool ev = written[i] != 0;
                // records of bones that are only read keep no references: they are never evaluated
                if (ev && (s1.flags & kHasParent)) s1.parent = add(s1.parent, false);
                else { s1.parent = -1; if (!ev) s1.flags &= ~kHasParent; }
                if (ev && (s1.flags & (kAppendRot | kAppendTrans))) s1.append_parent = add(s1.append_parent, false);
                else { s1.append_parent = -1; if (!ev) s1.flags &= ~(kAppendRot | kAppendTrans); }
                s1.link_slot = (ev && s1.link_slot >= 0) ? slot_of(lslots, s1.link_slot) : -1;
                if (!ev) s1.flags &= ~kIsLink;
                s1.morph_slot = (ev && s1.morph_slot >= 0) ? slot_of(mslots, s1.morph_slot) : -1;
                p.ik_img_static.push_back(s1);
            }
            if (int32_t(bones.size()) != I.n_bones) p.ik_img_ok = false;   // the translation must not grow the set
            p.ik_img_bones.insert(p.ik_img_bones.end(), bones.begin(), bones.begin() + I.n_bones);
            p.ik_img_written.insert(p.ik_img_written.end(), written.begin(), written.begin() + I.n_bones);
            I.lslots_begin = int32_t(p.ik_img_lslots.size()); I.n_lslots = int32_t(lslots.size());
            I.mslots_begin = int32_t(p.ik_img_mslots.size()); I.n_mslots = int32_t(mslots.size());
            p.ik_img_lslots.insert(p.ik_img_lslots.end(), lslots.begin(), lslots.end());
            p.ik_img_mslots.insert(p.ik_img_mslots.end(), mslots.begin(), mslots.end());
            I.region_f4 = 7 * I.n_bones + 2 * I.n_lslots + 2 * I.n_mslots;
            I.region_f4 |= 1;                           // odd: consecutive threads' regions start in different banks
            p.ik_img_max_region = std::max<uint32_t>(p.ik_img_max_region, uint32_t(I.region_f4));
            IkDesc kd = k;
            kd.bone = add(k.bone, false);
            kd.target = add(k.target, false);
            kd.link_begin = int32_t(p.ik_img_links.size());
            for (int32_t j = 0; j < k.link_count; ++j) {
                IkLink l = p.links[size_t(k.link_begin + j)];
                l.bone = add(l.bone, false);
                p.ik_img_links.push_back(l);
            }
            p.ik_img_desc.push_back(kd);
            p.ik_img.push_back(I);
        }
    }

    // material morphs grouped by affected material (extensions; libmmd leaves its material images untouched,
    // poser_impl.inl:355-358)
    p.n_materials = d.n_materials;
    if (p.extensions && d.n_materials && d.material_morph_entries) {
        if (d.n_materials > 65535u) return fail(err, MMDGPU_ERR_UNSUPPORTED, "more than 65535 materials");
        std::vector<std::vector<MaterialMorphEntry>> per_mat(d.n_materials);
        for (size_t n = 0; n < n_nodes; ++n) {
            const uint32_t m = uint32_t(p.node_morph[n]);
            if (d.morph_type[m] != MMDGPU_MORPH_MATERIAL) continue;
            for (uint32_t j = 0; j < d.morph_entry_count[m]; ++j) {
                const mmdgpu_material_morph_entry& e = d.material_morph_entries[d.morph_entry_begin[m] + j];
                if (e.method > MMDGPU_MATERIAL_ADD)
                    return fail(err, MMDGPU_ERR_INVALID_ARG, "unknown material morph method at morph " + std::to_string(m));
                MaterialMorphEntry x;
                x.node = int32_t(n);
                x.method = e.method;
                for (int k = 0; k < MMDGPU_MATERIAL_FIELDS; ++k) {
                    if (!std::isfinite(e.value[k]))
                        return fail(err, MMDGPU_ERR_INVALID_ARG, "non-finite material morph value at morph " + std::to_string(m));
                    x.value[k] = e.value[k];
                }
                if (e.material < 0 || uint32_t(e.material) >= d.n_materials) {
                    for (auto& v : per_mat) v.push_back(x);
                } else {
                    per_mat[size_t(e.material)].push_back(x);
                }
            }
        }
        p.material_morph_row.push_back(0);
        for (uint32_t i = 0; i < d.n_materials; ++i) {
            p.material_morph_entries.insert(p.material_morph_entries.end(), per_mat[i].begin(), per_mat[i].end());
            if (p.material_morph_entries.size() > (size_t(1) << 24))
                return fail(err, MMDGPU_ERR_UNSUPPORTED, "material morph expansion too large");
            p.material_morph_row.push_back(int32_t(p.material_morph_entries.size()));
        }
    }

    // ---------------------------------------------------------------- vertices
    p.norm_type.resize(nv);
    p.dev_type.resize(nv);
    p.bone_id.assign(size_t(nv) * 4, 0);
    p.weight.assign(size_t(nv) * 4, 0.0f);
    p.position.assign(d.position, d.position + size_t(nv) * 3);
    p.normal.assign(d.normal, d.normal + size_t(nv) * 3);
    if (d.uv) p.uv.assign(d.uv, d.uv + size_t(nv) * 2); else p.uv.assign(size_t(nv) * 2, 0.0f);
    bool any_sdef = false;
    for (uint32_t i = 0; i < nv; ++i) {
        uint8_t t = d.skin_type[i];
        if (t > MMDGPU_SKIN_QDEF) return fail(err, MMDGPU_ERR_INVALID_ARG, "unknown skinning type at vertex " + std::to_string(i));
        const int n_ids = (t == MMDGPU_SKIN_BDEF1) ? 1 : (t == MMDGPU_SKIN_BDEF2 || t == MMDGPU_SKIN_SDEF) ? 2 : 4;
        int32_t id[4] = {0, 0, 0, 0};
        float w[4] = {0, 0, 0, 0};
        for (int k = 0; k < n_ids; ++k) {
            id[k] = d.bone_id[size_t(i) * 4 + k];
            // libmmd reads bone_images_[id] unchecked (poser_impl.inl:412-432); reject instead
            if (id[k] < 0 || uint32_t(id[k]) >= nb)
                return fail(err, MMDGPU_ERR_BAD_INDEX, "bone id out of range at vertex " + std::to_string(i));
        }
        for (int k = 0; k < 4; ++k) w[k] = d.weight[size_t(i) * 4 + k];
        const bool was_qdef = (t == MMDGPU_SKIN_QDEF);
        if (was_qdef) t = MMDGPU_SKIN_BDEF4;  // libmmd cannot represent QDEF (L/model/model.inl:23-28)
        // Model::Normalize, L/model/model_impl.inl:406-452
        if (t == MMDGPU_SKIN_BDEF2) {
            if (w[0] == 0.0f) { id[0] = id[1]; t = MMDGPU_SKIN_BDEF1; }
            else if (w[0] == 1.0f) { t = MMDGPU_SKIN_BDEF1; }
        } else if (t == MMDGPU_SKIN_SDEF) {
            if (d.bone_parent[id[0]] != id[1] && d.bone_parent[id[1]] != id[0]) {
                if (w[0] == 0.0f) { id[0] = id[1]; t = MMDGPU_SKIN_BDEF1; }
                else if (w[0] == 1.0f) { t = MMDGPU_SKIN_BDEF1; }
                else t = MMDGPU_SKIN_BDEF2;
            }
        }
        p.norm_type[i] = t;
        // device type: fold the Lerp shortcuts (math_impl.inl:1246-1250) of Deform's BDEF2 / SDEF branch
        uint8_t dt;
        if (t == MMDGPU_SKIN_BDEF1) dt = kDevBdef1;
        else if (t == MMDGPU_SKIN_BDEF4) dt = (was_qdef && p.extensions) ? kDevQdef : kDevBdef4;
        else if (t == MMDGPU_SKIN_SDEF && p.extensions) { dt = kDevSdef; any_sdef = true; }
        else {
            if (w[0] < float(kEpsD)) { id[0] = id[1]; dt = kDevBdef1; }
            else if (w[0] > float(1.0 - kEpsD)) dt = kDevBdef1;
            else dt = kDevBdef2;
        }
        p.dev_type[i] = dt;
        const int keep = (dt == kDevBdef1) ? 1 : (dt == kDevBdef2 || dt == kDevSdef) ? 2 : 4;
        for (int k = 0; k < keep; ++k) p.bone_id[size_t(i) * 4 + k] = uint16_t(id[k]);
        if (dt == kDevBdef2 || dt == kDevSdef) p.weight[size_t(i) * 4] = w[0];
        else if (dt == kDevBdef4 || dt == kDevQdef) for (int k = 0; k < 4; ++k) p.weight[size_t(i) * 4 + k] = w[k];
        else p.weight[size_t(i) * 4] = 1.0f;
    }
    if (any_sdef) {
        auto grab = [&](const float* src, std::vector<float>& dst) {
            if (src) dst.assign(src, src + size_t(nv) * 3); else dst.assign(size_t(nv) * 3, 0.0f);
        };
        grab(d.sdef_c, p.sdef_c); grab(d.sdef_r0, p.sdef_r0); grab(d.sdef_r1, p.sdef_r1);
    }

    // ---------------------------------------------------------------- tiles (device vertex layout)
    // Vertices stay in PMX order in the OUTPUT (the renderer's index buffer is reused unchanged), but the static
    // streams of each kTileVerts-vertex tile are stored sorted by (skinning type, morph entry count, PMX index), so
    // that the 32 lanes of a warp step run one skinning branch and one morph loop trip count.  Sorted rank r
    // lands at storage position tile_position_of_rank(r).  Bone ids become indices into the tile's own list of
    // distinct bones (the CTA stages only those matrices).
    p.nv_pad = (nv + kTileVerts - 1) / kTileVerts * kTileVerts;
    p.n_tiles = p.nv_pad / kTileVerts;
    p.tile_orig.assign(p.nv_pad, 0);
    p.st_type.assign(p.nv_pad, kDevBdef1);
    p.st_local_id.assign(size_t(p.nv_pad) * 4, 0);
    p.st_weight.assign(size_t(p.nv_pad) * 4, 0.0f);
    p.tile_bone_begin.assign(size_t(p.n_tiles) + 1, 0);
    p.ell_base.assign(size_t(p.n_tiles) * kTileGroups, 0);
    p.ell_rounds.assign(size_t(p.n_tiles) * kTileGroups, 0);
    p.pad_node = uint32_t(n_nodes);
    {
        std::vector<int32_t> local_of(nb, -1);
        std::vector<uint16_t> rank_vertex(p.nv_pad);  // tile rank -> PMX index within the tile
        // tile orders, tiles in parallel (the hill climbing is the expensive part of the plan: ~1 ms per tile)
        static const bool pairing = [] { const char* e = std::getenv("MMDGPU_TILE_PAIRING"); return !(e && e[0] == '0'); }();
        auto order_tile = [&](uint32_t t, std::vector<uint32_t>& order, std::vector<TileVert>& tv) {
            const uint32_t v0 = t * kTileVerts;
            auto type_of = [&](uint32_t i) -> uint32_t { return (v0 + i < nv) ? p.dev_type[v0 + i] : uint32_t(kDevBdef1); };
            auto count_of = [&](uint32_t i) -> uint32_t { return (v0 + i < nv) ? p.csr_row[v0 + i + 1] - p.csr_row[v0 + i] : 0u; };
            if (pairing) {
                for (uint32_t i = 0; i < kTileVerts; ++i) {
                    TileVert& x = tv[i];
                    x.ty = uint8_t(std::min<uint32_t>(type_of(i), 4u));
                    x.keep = (x.ty == kDevBdef1) ? 1 : (x.ty == kDevBdef2 || x.ty == kDevSdef) ? 2 : 4;
                    x.cnt = count_of(i);
                    for (int k = 0; k < 4; ++k) x.id[k] = (v0 + i < nv && k < x.keep) ? p.bone_id[size_t(v0 + i) * 4 + k] : uint16_t(0xFFFF);
                }
                pair_tile_order(order, tv);
            } else {
                for (uint32_t i = 0; i < kTileVerts; ++i) order[i] = i;
                std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                    const uint32_t ta = type_of(a), tb = type_of(b);
                    if (ta != tb) return ta < tb;
                    return count_of(a) < count_of(b);
                });
                refine_tile_order(order, type_of, count_of);
            }
            for (uint32_t r = 0; r < kTileVerts; ++r) rank_vertex[size_t(v0) + r] = uint16_t(order[r]);
        };
        {
            const uint32_t n_thr = std::max(1u, std::min<uint32_t>({std::thread::hardware_concurrency(), 16u, p.n_tiles / 8u}));
            std::atomic<uint32_t> next{0};
            auto work = [&] {
                std::vector<uint32_t> order(kTileVerts);
                std::vector<TileVert> tv(kTileVerts);
                for (uint32_t t; (t = next.fetch_add(1)) < p.n_tiles;) order_tile(t, order, tv);
            };
            std::vector<std::thread> pool;
            for (uint32_t i = 1; i < n_thr; ++i) pool.emplace_back(work);
            work();
            for (auto& th : pool) th.join();
        }
        for (uint32_t t = 0; t < p.n_tiles; ++t) {
            const uint32_t v0 = t * kTileVerts;
            const uint16_t* order = &rank_vertex[size_t(v0)];
            // distinct bones of the tile, ascending
            std::vector<uint16_t> used;
            for (uint32_t i = 0; i < kTileVerts && v0 + i < nv; ++i) {
                const uint8_t dt = p.dev_type[v0 + i];
                const int keep = (dt == kDevBdef1) ? 1 : (dt == kDevBdef2 || dt == kDevSdef) ? 2 : 4;
                for (int k = 0; k < keep; ++k) {
                    const uint16_t b = p.bone_id[size_t(v0 + i) * 4 + k];
                    if (local_of[b] < 0) { local_of[b] = 0; used.push_back(b); }
                }
            }
            if (used.empty()) { used.push_back(0); local_of[0] = 0; }  // padded vertices read bone 0
            std::sort(used.begin(), used.end());
            for (size_t k = 0; k < used.size(); ++k) local_of[used[k]] = int32_t(k);
            p.tile_bone_begin[t] = uint32_t(p.tile_bones.size());
            p.tile_bones.insert(p.tile_bones.end(), used.begin(), used.end());
            p.max_tile_bones = std::max<uint32_t>(p.max_tile_bones, uint32_t(used.size()));
            for (uint32_t r = 0; r < kTileVerts; ++r) {
                const uint32_t i = order[r];
                const size_t pos = size_t(v0) + tile_position_of_rank(r);
                p.tile_orig[pos] = uint16_t(i);
                if (v0 + i < nv) {
                    p.st_type[pos] = p.dev_type[v0 + i];
                    for (int k = 0; k < 4; ++k) {
                        p.st_local_id[pos * 4 + k] = uint16_t(local_of[p.bone_id[size_t(v0 + i) * 4 + k]] < 0
                                                                   ? 0 : local_of[p.bone_id[size_t(v0 + i) * 4 + k]]);
                        p.st_weight[pos * 4 + k] = p.weight[size_t(v0 + i) * 4 + k];
                    }
                } else {
                    p.st_local_id[pos * 4] = uint16_t(local_of[0] < 0 ? 0 : local_of[0]);
                }
            }
            for (uint16_t b : used) local_of[b] = -1;
        }
        p.tile_bone_begin[p.n_tiles] = uint32_t(p.tile_bones.size());

        // sliced ELL of a per-vertex CSR: group g = rank / 32 holds ranks [32 g, 32 g + 32); rounds = the group's
        // largest row; entry (round k, lane l) at base + 32 k + l; padding points at the always-zero slot
        auto build_ell = [&](const std::vector<uint32_t>& row, const std::vector<uint32_t>& node, const std::vector<float>& off,
                             int width, std::vector<uint32_t>& base_of, std::vector<uint32_t>& rounds_of,
                             std::vector<uint32_t>& ell_node, std::vector<float>& ell_off) -> mmdgpu_status {
            base_of.assign(size_t(p.n_tiles) * kTileGroups, 0);
            rounds_of.assign(size_t(p.n_tiles) * kTileGroups, 0);
            ell_node.clear(); ell_off.clear();
            for (uint32_t t = 0; t < p.n_tiles; ++t) {
                const uint32_t v0 = t * kTileVerts;
                auto count_of = [&](uint32_t i) -> uint32_t { return (v0 + i < nv) ? row[v0 + i + 1] - row[v0 + i] : 0u; };
                for (uint32_t g = 0; g < kTileGroups; ++g) {
                    uint32_t rounds = 0;
                    for (uint32_t l = 0; l < 32; ++l) rounds = std::max(rounds, count_of(rank_vertex[size_t(v0) + g * 32 + l]));
                    const size_t base = ell_node.size();
                    if (base + size_t(rounds) * 32 > 0xFFFFFFFFull) return fail(err, MMDGPU_ERR_UNSUPPORTED, "morph entry table exceeds 2^32 entries");
                    base_of[size_t(t) * kTileGroups + g] = uint32_t(base);
                    rounds_of[size_t(t) * kTileGroups + g] = rounds;
                    ell_node.resize(base + size_t(rounds) * 32, p.pad_node);
                    ell_off.resize((base + size_t(rounds) * 32) * size_t(width), 0.0f);
                    for (uint32_t l = 0; l < 32; ++l) {
                        const uint32_t i = rank_vertex[size_t(v0) + g * 32 + l];
                        const uint32_t cnt = count_of(i);
                        for (uint32_t k = 0; k < cnt; ++k) {
                            const size_t e = size_t(row[v0 + i]) + k, at = base + size_t(k) * 32 + l;
                            ell_node[at] = node[e];
                            for (int c = 0; c < width; ++c) ell_off[at * size_t(width) + c] = off[e * size_t(width) + c];
                        }
                    }
                }
            }
            return MMDGPU_OK;
        };
        if (mmdgpu_status st = build_ell(p.csr_row, p.csr_node, p.csr_offset, 3, p.ell_base, p.ell_rounds, p.ell_node, p.ell_offset)) return st;
        if (p.extensions)
            if (mmdgpu_status st = build_ell(p.uv_row, p.uv_node, p.uv_offset, 4, p.uv_ell_base, p.uv_ell_rounds, p.uv_ell_node, p.uv_ell_offset)) return st;

        // spherical-deform parameters per storage position (extensions): C and the two blended centres
        // cr0 = (C + (C + R0 - rw)) / 2, cr1 = (C + (C + R1 - rw)) / 2 with rw = R0*w0 + R1*w1
        if (p.extensions && !p.sdef_c.empty()) {
            p.st_sdef.assign(size_t(p.nv_pad) * 12, 0.0f);
            for (uint32_t pos = 0; pos < p.nv_pad; ++pos) {
                const uint32_t src = (pos / kTileVerts) * kTileVerts + p.tile_orig[pos];
                if (src >= nv || p.dev_type[src] != kDevSdef) continue;
                const float w0 = p.weight[size_t(src) * 4], w1 = 1.0f - w0;
                float* o = &p.st_sdef[size_t(pos) * 12];
                for (int c = 0; c < 3; ++c) {
                    const float C = p.sdef_c[size_t(src) * 3 + c], R0 = p.sdef_r0[size_t(src) * 3 + c], R1 = p.sdef_r1[size_t(src) * 3 + c];
                    const float rw = R0 * w0 + R1 * w1;
                    o[c] = C;
                    o[4 + c] = (C + (C + R0 - rw)) * 0.5f;
                    o[8 + c] = (C + (C + R1 - rw)) * 0.5f;
                }
            }
        }
    }

    // introspection mirrors
    p.op_kind_u8.resize(p.ops.size());
    p.op_arg_i32.resize(p.ops.size());
    for (size_t i = 0; i < p.ops.size(); ++i) {
        p.op_kind_u8[i] = p.ops[i].kind;
        p.op_arg_i32[i] = (p.ops[i].kind == kOpIk) ? p.iks[p.ops[i].arg].bone : p.ops[i].arg;
    }
    p.ik_fix_u8.resize(p.links.size());
    p.ik_order_u8.resize(p.links.size());
    for (size_t i = 0; i < p.links.size(); ++i) { p.ik_fix_u8[i] = p.links[i].fix; p.ik_order_u8[i] = p.links[i].order; }
    p.phase_split_i32.assign(1, p.phase_split);
    return MMDGPU_OK;
}

// --------------------------------------------------------------------------------------------------
// mmd::Motion storage (std::map<frame, key>, L/motion/motion.inl:128-129): sorted by frame, a later record
// with the same frame replaces the earlier one.
mmdgpu_status build_anim(const mmdgpu_anim_desc& d, uint32_t nb, uint32_t nm, HostAnim& a, std::string& err) {
    a = HostAnim();
    a.nb = nb; a.nm = nm;
    a.bone_key_begin.assign(nb, 0); a.bone_key_count.assign(nb, 0); a.bone_tracked.assign(nb, 0);
    a.morph_key_begin.assign(nm, 0); a.morph_key_count.assign(nm, 0); a.morph_tracked.assign(nm, 0);
    std::unordered_map<uint32_t, uint32_t> table_of;
    auto sorted_unique = [](std::vector<uint32_t>& idx, auto frame_of) {
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return frame_of(x) < frame_of(y); });
        size_t o = 0;
        for (size_t i = 0; i < idx.size(); ++i) {
            if (i + 1 < idx.size() && frame_of(idx[i + 1]) == frame_of(idx[i])) continue;
            idx[o++] = idx[i];
        }
        idx.resize(o);
    };
    if (d.n_bone_tracks && (!d.bone_track_bone || !d.bone_track_key_begin || !d.bone_track_key_count))
        return fail(err, MMDGPU_ERR_INVALID_ARG, "bone track arrays missing");
    if (d.n_morph_tracks && (!d.morph_track_morph || !d.morph_track_key_begin || !d.morph_track_key_count))
        return fail(err, MMDGPU_ERR_INVALID_ARG, "morph track arrays missing");
    std::vector<uint32_t> idx;
    for (uint32_t t = 0; t < d.n_bone_tracks; ++t) {
        const int32_t b = d.bone_track_bone[t];
        if (b < 0 || uint32_t(b) >= nb) return fail(err, MMDGPU_ERR_BAD_INDEX, "bone track " + std::to_string(t) + " names a bone out of range");
        if (a.bone_tracked[b]) return fail(err, MMDGPU_ERR_INVALID_ARG, "bone " + std::to_string(b) + " appears in two tracks");
        const uint32_t k0 = d.bone_track_key_begin[t], kn = d.bone_track_key_count[t];
        if (uint64_t(k0) + kn > d.n_bone_keys) return fail(err, MMDGPU_ERR_BAD_INDEX, "bone key range out of range");
        if (kn && !d.bone_keys) return fail(err, MMDGPU_ERR_INVALID_ARG, "bone keys missing");
        idx.resize(kn);
        for (uint32_t i = 0; i < kn; ++i) idx[i] = k0 + i;
        sorted_unique(idx, [&](uint32_t i) { return d.bone_keys[i].frame; });
        a.bone_tracked[b] = 1;
        a.bone_key_begin[b] = uint32_t(a.key_frame.size());
        a.bone_key_count[b] = uint32_t(idx.size());
        for (uint32_t i : idx) {
            const mmdgpu_bone_key& k = d.bone_keys[i];
            a.length = std::max(a.length, k.frame);
            a.key_frame.push_back(k.frame);
            a.key_T.insert(a.key_T.end(), {k.translation[0], k.translation[1], k.translation[2], 0.0f});
            a.key_R.insert(a.key_R.end(), k.rotation, k.rotation + 4);
            for (int c = 0; c < 4; ++c) {
                uint32_t packed;
                std::memcpy(&packed, k.interp[c], 4);
                auto it = table_of.find(packed);
                uint32_t ti;
                if (it != table_of.end()) ti = it->second;
                else {
                    float tab[32];
                    if (bezier_table(k.interp[c], tab)) ti = 0xFFFFFFFFu;
                    else {
                        ti = uint32_t(a.tables.size() / 32);
                        a.tables.insert(a.tables.end(), tab, tab + 32);
                    }
                    table_of.emplace(packed, ti);
                }
                a.key_curve.push_back(ti);
            }
        }
    }
    for (uint32_t t = 0; t < d.n_morph_tracks; ++t) {
        const int32_t m = d.morph_track_morph[t];
        if (m < 0 || uint32_t(m) >= nm) return fail(err, MMDGPU_ERR_BAD_INDEX, "morph track " + std::to_string(t) + " names a morph out of range");
        if (a.morph_tracked[m]) return fail(err, MMDGPU_ERR_INVALID_ARG, "morph " + std::to_string(m) + " appears in two tracks");
        const uint32_t k0 = d.morph_track_key_begin[t], kn = d.morph_track_key_count[t];
        if (uint64_t(k0) + kn > d.n_morph_keys) return fail(err, MMDGPU_ERR_BAD_INDEX, "morph key range out of range");
        if (kn && !d.morph_keys) return fail(err, MMDGPU_ERR_INVALID_ARG, "morph keys missing");
        idx.resize(kn);
        for (uint32_t i = 0; i < kn; ++i) idx[i] = k0 + i;
        sorted_unique(idx, [&](uint32_t i) { return d.morph_keys[i].frame; });
        a.morph_tracked[m] = 1;
        a.morph_key_begin[m] = uint32_t(a.mkey_frame.size());
        a.morph_key_count[m] = uint32_t(idx.size());
        for (uint32_t i : idx) {
            a.length = std::max(a.length, d.morph_keys[i].frame);
            a.mkey_frame.push_back(d.morph_keys[i].frame);
            a.mkey_weight.push_back(d.morph_keys[i].weight);
        }
    }
    return MMDGPU_OK;
}

}  // namespace mmdgpu
