// symbolic.hpp -- one-time host-side symbolic analysis shared by every sweep point.
//
// The reference solves each point with a dense pivoting LU inside third-party solvers
// (LinearSolve 2.31.2).  Here the sparsity pattern, a static pivot
// sequence, the fill pattern and the elimination / assembly schedules are computed once on
// the host; the GPU then refactors all instances with identical control flow.
//
//   1. pattern of J = G + alpha*C from the stamps of all devices
//   2. row matching: put a numerically safe entry on every diagonal (constant +-1 incidence
//      entries of voltage-defined branches first, then KCL self-conductances) -- Hungarian
//      assignment on a small class-cost matrix
//   3. Markowitz ordering of the (row-permuted) diagonal pivots with exact fill tracking
//   4. fill pattern of L+U, elimination schedule (pair updates), triangular-solve schedule
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <vector>

namespace cb {

struct PatternEntry {
    int row, col;
    int cls;  // 3 constant incidence, 2 self-conductance, 1 general, 0 unsafe (may be zero)
};

struct Symbolic {
    int N = 0;
    // permutations: elimination step k pivots on original (prow[k], pcol[k])
    std::vector<int> prow, pcol;      // [N]
    std::vector<int> row_to_step, col_to_step;  // inverse
    // LU storage: entries in step space (i = row step, j = col step); position = index into values
    int nnz_lu = 0;
    std::vector<int> lu_i, lu_j;      // [nnz_lu]
    std::vector<int> diag_pos;        // [N]
    // elimination schedule per pivot k
    std::vector<int> l_ptr, l_pos, l_row;   // L entries of column k (rows below), positions, row steps
    std::vector<int> u_ptr, u_pos, u_col;   // U entries of row k (cols right of diag)
    std::vector<int> pair_ptr, pair_dst;    // for pivot k: nL*nU destinations, index li*nU+uj
    // column-oriented U for the backward solve: entries (i<k, k)
    std::vector<int> uc_ptr, uc_pos, uc_row;
    int64_t flops = 0;                // 2 * sum nL*nU + divisions
    int nnz_a = 0;
    std::map<std::pair<int, int>, int> pos_of_orig;  // original (row, col) -> LU position
    std::string error;
};

// Hungarian algorithm (Kuhn-Munkres, O(n^3)) for a square cost matrix; returns row -> col.
inline bool hungarian(const std::vector<std::vector<double>>& cost, std::vector<int>& row_to_col, double big) {
    const int n = (int)cost.size();
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<double> u(n + 1, 0), v(n + 1, 0);
    std::vector<int> p(n + 1, 0), way(n + 1, 0);
    for (int i = 1; i <= n; i++) {
        p[0] = i;
        int j0 = 0;
        std::vector<double> minv(n + 1, INF);
        std::vector<char> used(n + 1, 0);
        do {
            used[j0] = 1;
            int i0 = p[j0], j1 = 0;
            double delta = INF;
            for (int j = 1; j <= n; j++) {
                if (used[j]) continue;
                double cur = cost[i0 - 1][j - 1] - u[i0] - v[j];
                if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                if (minv[j] < delta) { delta = minv[j]; j1 = j; }
            }
            for (int j = 0; j <= n; j++) {
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            }
            j0 = j1;
        } while (p[j0] != 0);
        do {
            int j1 = way[j0];
            p[j0] = p[j1];
            j0 = j1;
        } while (j0);
    }
    row_to_col.assign(n, -1);
    for (int j = 1; j <= n; j++) row_to_col[p[j] - 1] = j - 1;
    for (int i = 0; i < n; i++)
        if (row_to_col[i] < 0 || cost[i][row_to_col[i]] >= big) return false;
    return true;
}

inline bool analyze(int N, const std::vector<PatternEntry>& raw, Symbolic& S) {
    S = Symbolic();
    S.N = N;
    // merge duplicates keeping the best class
    std::map<std::pair<int, int>, int> cls;
    for (const auto& e : raw) {
        if (e.row < 0 || e.col < 0) continue;
        auto key = std::make_pair(e.row, e.col);
        auto it = cls.find(key);
        if (it == cls.end()) cls[key] = e.cls;
        else it->second = std::max(it->second, e.cls);
    }
    S.nnz_a = (int)cls.size();
    // ---- 2. matching
    const double BIG = 1e9;
    static const double class_cost[4] = {400.0, 20.0, 1.0, 0.0};
    std::vector<std::vector<double>> cost(N, std::vector<double>(N, BIG));
    for (const auto& kv : cls) {
        double c = class_cost[std::min(3, std::max(0, kv.second))];
        // off-diagonal class-2/1 entries are worse pivots than the same class on the diagonal
        if (kv.first.first != kv.first.second && kv.second < 3) c += 40.0;
        cost[kv.first.first][kv.first.second] = c;
    }
    std::vector<int> row_to_col;
    if (N > 0 && !hungarian(cost, row_to_col, BIG)) {
        S.error = "structurally singular MNA matrix (no perfect matching): floating node or source loop";
        return false;
    }
    // permuted structure: pivot p <-> (row r_p, col c_p = row_to_col[r_p]); work in "pivot ids" = row ids
    // column j belongs to pivot col_owner[j]
    std::vector<int> col_owner(N);
    for (int r = 0; r < N; r++) col_owner[row_to_col[r]] = r;
    // adjacency in pivot-id space: rows[p] = set of pivot-ids q such that entry (row p, col of q) exists
    std::vector<std::set<int>> rows(N), cols(N);
    for (const auto& kv : cls) {
        int p = kv.first.first, q = col_owner[kv.first.second];
        rows[p].insert(q);
        cols[q].insert(p);
    }
    // ---- 3. Markowitz ordering with fill
    std::vector<char> done(N, 0);
    std::vector<int> order;
    order.reserve(N);
    for (int step = 0; step < N; step++) {
        int best = -1;
        long bestcost = 0;
        int bestcls = -1;
        for (int p = 0; p < N; p++) {
            if (done[p]) continue;
            long r = (long)rows[p].size() - 1, c = (long)cols[p].size() - 1;
            long mk = r * c;
            int pc = cls.count({p, row_to_col[p]}) ? cls[{p, row_to_col[p]}] : 0;
            if (best < 0 || mk < bestcost || (mk == bestcost && pc > bestcls)) {
                best = p; bestcost = mk; bestcls = pc;
            }
        }
        const int k = best;
        done[k] = 1;
        order.push_back(k);
        // eliminate k: every remaining row i with entry in column k gets row k's remaining entries
        std::vector<int> li, uj;
        for (int i : cols[k]) if (!done[i]) li.push_back(i);
        for (int j : rows[k]) if (!done[j]) uj.push_back(j);
        for (int i : li)
            for (int j : uj) {
                rows[i].insert(j);
                cols[j].insert(i);
            }
        // remove k from the active structure (keep the sets for the final pattern pass below)
        for (int i : li) { /* L entry (i,k) stays recorded in cols[k] */ (void)i; }
        for (int j : uj) cols[j].erase(k);
        for (int i : li) rows[i].erase(k);
        // remember L/U structure of this pivot
        // (re-derived below from a clean symbolic pass to keep this loop simple)
    }
    // step numbering
    std::vector<int> step_of(N);
    for (int s = 0; s < N; s++) step_of[order[s]] = s;
    S.prow.resize(N); S.pcol.resize(N); S.row_to_step.resize(N); S.col_to_step.resize(N);
    for (int s = 0; s < N; s++) {
        S.prow[s] = order[s];
        S.pcol[s] = row_to_col[order[s]];
        S.row_to_step[S.prow[s]] = s;
        S.col_to_step[S.pcol[s]] = s;
    }
    // ---- 4. clean symbolic factorisation in step space
    std::vector<std::set<int>> R(N);  // R[i] = set of column steps in row step i
    for (const auto& kv : cls) R[S.row_to_step[kv.first.first]].insert(S.col_to_step[kv.first.second]);
    for (int i = 0; i < N; i++) R[i].insert(i);
    std::vector<std::vector<int>> colrows(N);  // rows below diag with entry in column k (built progressively)
    for (int k = 0; k < N; k++) {
        // rows i > k with (i,k) present
        std::vector<int> li;
        for (int i = k + 1; i < N; i++) if (R[i].count(k)) li.push_back(i);
        std::vector<int> uj;
        for (int j : R[k]) if (j > k) uj.push_back(j);
        for (int i : li) for (int j : uj) R[i].insert(j);
    }
    // positions: row-major over step rows
    std::vector<std::map<int, int>> pos(N);
    S.diag_pos.resize(N);
    for (int i = 0; i < N; i++)
        for (int j : R[i]) {
            pos[i][j] = S.nnz_lu++;
            S.lu_i.push_back(i);
            S.lu_j.push_back(j);
            if (i == j) S.diag_pos[i] = pos[i][j];
        }
    for (const auto& kv : cls)
        S.pos_of_orig[kv.first] = pos[S.row_to_step[kv.first.first]][S.col_to_step[kv.first.second]];
    S.l_ptr.assign(1, 0); S.u_ptr.assign(1, 0); S.pair_ptr.assign(1, 0); S.uc_ptr.assign(1, 0);
    for (int k = 0; k < N; k++) {
        std::vector<int> li, uj;
        for (int i = k + 1; i < N; i++) if (R[i].count(k)) li.push_back(i);
        for (int j : R[k]) if (j > k) uj.push_back(j);
        for (int i : li) { S.l_pos.push_back(pos[i][k]); S.l_row.push_back(i); }
        for (int j : uj) { S.u_pos.push_back(pos[k][j]); S.u_col.push_back(j); }
        for (int i : li) for (int j : uj) S.pair_dst.push_back(pos[i][j]);
        S.l_ptr.push_back((int)S.l_pos.size());
        S.u_ptr.push_back((int)S.u_pos.size());
        S.pair_ptr.push_back((int)S.pair_dst.size());
        S.flops += 2LL * (int64_t)li.size() * (int64_t)uj.size() + (int64_t)li.size();
    }
    for (int k = 0; k < N; k++) {
        for (int i = 0; i < k; i++)
            if (R[i].count(k)) { S.uc_pos.push_back(pos[i][k]); S.uc_row.push_back(i); }
        S.uc_ptr.push_back((int)S.uc_pos.size());
    }
    return true;
}

}  // namespace cb
