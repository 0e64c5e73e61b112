// Host-side column numbering for the K2+K3 kernel (no device work).
//
// Phase 2 of diffuse_score_kernel lets lane t of a warp score pairs 128*b + 4*t + u (u = 0..3):
// the 32 lanes read D[lig[128b+4t+u]] in one shared-memory instruction, then D[rec[..]].  Those 32
// reads are conflict-free when distinct columns fall into distinct banks (column mod 32).  The Xu
// column ids are ours to choose, so this routine colours the "appear in the same instruction" graph
// of the LR table with 32 colours (largest-degree-first greedy, smallest colour class first) and
// returns column = colour + 32 * (rank inside the colour class).  A vertex with no free colour takes
// the least-conflicting one: a failure only costs bank conflicts, never correctness.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

using namespace laris;

extern "C" int laris_lr_column_layout(const int32_t* lig, const int32_t* rec, int32_t P, int32_t G_u,
                                      int32_t* column_of, int32_t* n_columns) {
    LARIS_CHECK_ARG(lig && rec && column_of && n_columns && P >= 1 && G_u >= 1, "laris_lr_column_layout: bad arguments");
    for (int p = 0; p < P; ++p)
        LARIS_CHECK_ARG(lig[p] >= 0 && lig[p] < G_u && rec[p] >= 0 && rec[p] < G_u,
                        "laris_lr_column_layout: gene index out of range at pair %d", p);
    std::vector<std::vector<int32_t>> adj(G_u);
    std::vector<int32_t> grp;
    for (int side = 0; side < 2; ++side) {
        const int32_t* a = side ? rec : lig;
        for (int b0 = 0; b0 < P; b0 += 128)
            for (int u = 0; u < 4; ++u) {
                grp.clear();
                for (int t = 0; t < 32; ++t) {
                    const int p = b0 + 4 * t + u;
                    if (p < P) grp.push_back(a[p]);
                }
                std::sort(grp.begin(), grp.end());
                grp.erase(std::unique(grp.begin(), grp.end()), grp.end());
                for (size_t x = 0; x < grp.size(); ++x)
                    for (size_t y = 0; y < grp.size(); ++y)
                        if (x != y) adj[grp[x]].push_back(grp[y]);
            }
    }
    for (auto& v : adj) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    std::vector<int32_t> order(G_u);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return adj[x].size() > adj[y].size(); });
    const int cap = (G_u + 31) / 32 + 2;                 // keeps the column range within ~G_u + 96
    std::vector<int32_t> colour(G_u, -1), cls(32, 0);
    auto pick = [&](int32_t g, bool allow_full) {
        int hits[32] = {0};
        for (int32_t h : adj[g])
            if (colour[h] >= 0) ++hits[colour[h]];
        int best = -1;
        for (int pass = 0; pass < 3 && best < 0; ++pass)   // free+roomy, then least-conflicting roomy, then anything
            for (int c = 0; c < 32; ++c) {
                if (pass < 2 && cls[c] >= cap) continue;
                if (pass == 2 && !allow_full) continue;
                if (pass == 0 && hits[c]) continue;
                if (best < 0 || hits[c] < hits[best] || (hits[c] == hits[best] && cls[c] < cls[best])) best = c;
            }
        return best;
    };
    for (int32_t g : order) {
        const int c = pick(g, true);
        colour[g] = c;
        ++cls[c];
    }
    // repair sweeps: a vertex that still shares a colour with a neighbour moves to a better colour
    for (int sweep = 0; sweep < 8; ++sweep) {
        bool moved = false;
        for (int32_t g : order) {
            int cur = 0;
            for (int32_t h : adj[g]) cur += colour[h] == colour[g];
            if (!cur) continue;
            const int old = colour[g];
            colour[g] = -1;
            --cls[old];
            int c = pick(g, false);
            int now = 0;
            if (c >= 0) for (int32_t h : adj[g]) now += colour[h] == c;
            if (c < 0 || now >= cur) c = old; else moved = true;
            colour[g] = c;
            ++cls[c];
        }
        if (!moved) break;
    }
    std::fill(cls.begin(), cls.end(), 0);
    for (int32_t g = 0; g < G_u; ++g) column_of[g] = colour[g] + 32 * cls[colour[g]]++;
    *n_columns = 32 * *std::max_element(cls.begin(), cls.end());
    return LARIS_OK;
}
