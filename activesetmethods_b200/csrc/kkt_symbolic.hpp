// Host-side symbolic analysis of the quasi-definite KKT matrix of the barrier engine (ipm.cuh)
//
//        M = [ -(Dx + d)    K'     ]      nodes 0..n-1   : columns of the LP (primal step)
//            [    K      (Ew + d)  ]      nodes n..n+m-1 : rows of the LP    (dual step)
//
// The sparsity pattern of the Jacobian is fixed for the life of a handle (the reference uploads j_str once,
// /root/reference/src/model.jl:10, and only pushes values per iteration, src/algorithms/subproblem.jl:438-457),
// so everything that depends on the pattern alone is done here once, on the host, in plain C++:
//   * a fill-reducing ordering (exact minimum degree on the explicit elimination graph),
//   * the pattern of L in  P M P' = L D L'  (any symmetric permutation of a quasi-definite matrix has such a
//     factorisation with D diagonal -- no numerical pivoting, hence a schedule that is identical for every LP of
//     a batch and every iteration),
//   * a level schedule (column j only needs columns of lower levels; levels grow strictly along the
//     elimination tree),
//   * for every entry of L and every pivot the products  W[i,k] * W[j,k] / d[k]  that update it ("terms";
//     W = L D), grouped by the LEVEL OF THE SOURCE COLUMN k: step l of the numerical factorisation applies the
//     updates of the columns of level l to all their targets (fan-out), one "chunk" = the terms of one target
//     in that step, summed in ascending k by one thread -- no atomics, bit-reproducible, and the critical path
//     of a step is the longest chunk (one term near the root of the tree, where a fan-in dot product would be
//     hundreds of terms long).  Nine chunks in ten are a single term: those are stored first in every step so
//     that the kernel can keep four of them in flight per warp,
//   * the same chunked fan-out lists for the forward substitution; the backward substitution needs no chunks
//     (the rows of a column are its ancestors, which sit on distinct levels),
//   * the launch plan (wide steps: one launch each; runs of narrow steps: one launch of a single block per 32
//     scenarios that walks them with __syncthreads()).
// No CUDA in this header: it is also compiled into the host-only self test (asm_kkt_selftest).
#pragma once
#include <stdint.h>
#include <algorithm>
#include <functional>
#include <queue>
#include <utility>
#include <vector>

namespace asmb {

struct KktTerm {
    int a, b, k, t;  // target t -= W[a] * W[b] * invd[k];  t < nnzL: entry of L, else pivot t - nnzL (see kLastBit)
};
struct KktFwdItem {
    int pos, src, k, dst;  // v[dst] -= W[pos] * v[src] * invd[k]   (src, dst: node ids; k: permuted column of src)
};
struct KktBwdItem {
    int pos, src, dst, k;  // v[dst] -= invd[k] * W[pos] * v[src]   (k: permuted column of dst)
};
struct KktRange {
    int begin, end;
};
struct KktLaunch {
    int l0, l1;  // steps [l0, l1)
    int fused;   // 1: one block walks the steps; 0: l1 == l0 + 1, grid over the items of the step
    int items;   // largest step of the launch
};
constexpr int kLastBit = 1 << 30;  // in KktTerm::t of the first term of a chunk: last update of this pivot -> invert it

struct KktSymbolic {
    int n = 0, m = 0, N = 0;
    int64_t nnzL = 0, nterms = 0;
    int n_levels = 0;
    std::vector<int> perm, inv;  // perm[k] = node eliminated k-th; inv[node] = k
    std::vector<int> Lp, Li;     // columns of L (permuted numbering), rows ascending
    std::vector<int> kmap;       // CSR entry of K -> entry of L
    std::vector<int> level;      // per permuted column
    // factorisation: terms grouped by step (= level of the source column).  Inside a step the chunks that consist of
    // a single term come first -- terms [fs_beg[l], fs_end[l]), one independent update each, which the kernel
    // software-pipelines four at a time -- followed by the multi-term chunks fmchunk[fmstep[l] .. fmstep[l+1]) =
    // (begin, end) ranges of terms with a common target, summed in ascending k
    std::vector<KktTerm> terms;
    std::vector<int> fs_beg, fs_end, fmstep;
    std::vector<KktRange> fmchunk;
    int64_t n_fchunks = 0, n_wchunks = 0;
    // distinct f64 operands a step has to read / targets it has to read and write, summed over the steps: the traffic a
    // batch whose working set exceeds L2 cannot avoid (nothing survives in cache from one level to the next)
    int64_t f_distinct_reads = 0, f_targets = 0, w_distinct_reads = 0, w_targets = 0, b_distinct_reads = 0, b_targets = 0;
    int longest_chunk = 0;
    // forward substitution, same layout
    std::vector<KktFwdItem> fwd;
    std::vector<int> ws_beg, ws_end, wmstep;
    std::vector<KktRange> wmchunk;
    // backward substitution: items of step l = entries whose row has level l (walked downwards)
    std::vector<KktBwdItem> bwd;
    std::vector<int> bstep;
    std::vector<KktLaunch> flaunch, wlaunch, blaunch;

    // K: m x n CSR pattern (0-based).  Returns 0, or -1 when an index is out of range / the term count overflows.
    int build(int n_, int m_, const int *row_ptr, const int *col_idx, int narrow = 64) {
        n = n_;
        m = m_;
        N = n + m;
        std::vector<std::vector<int>> adj(N);
        for (int i = 0; i < m; ++i)
            for (int q = row_ptr[i]; q < row_ptr[i + 1]; ++q) {
                const int j = col_idx[q];
                if (j < 0 || j >= n) return -1;
                adj[j].push_back(n + i);
                adj[n + i].push_back(j);
            }
        for (auto &a : adj) {
            std::sort(a.begin(), a.end());
            a.erase(std::unique(a.begin(), a.end()), a.end());
        }
        // ---- minimum degree, explicit fill, ties by node id (deterministic)
        typedef std::pair<int, int> DN;
        std::priority_queue<DN, std::vector<DN>, std::greater<DN>> heap;
        std::vector<char> gone(N, 0);
        for (int v = 0; v < N; ++v) heap.push(DN((int)adj[v].size(), v));
        perm.assign(N, -1);
        inv.assign(N, -1);
        std::vector<std::vector<int>> colpat(N);
        std::vector<int> merged;
        for (int k = 0; k < N; ++k) {
            int v = -1;
            while (true) {
                const DN top = heap.top();
                heap.pop();
                if (!gone[top.second] && (int)adj[top.second].size() == top.first) {
                    v = top.second;
                    break;
                }
            }
            gone[v] = 1;
            perm[k] = v;
            inv[v] = k;
            std::vector<int> S;
            S.swap(adj[v]);
            for (int u : S) {
                const std::vector<int> &a = adj[u];
                merged.clear();
                merged.reserve(a.size() + S.size());
                size_t ia = 0, is = 0;
                while (ia < a.size() || is < S.size()) {
                    int x;
                    if (is >= S.size() || (ia < a.size() && a[ia] < S[is]))
                        x = a[ia++];
                    else if (ia >= a.size() || S[is] < a[ia])
                        x = S[is++];
                    else {
                        x = a[ia++];
                        ++is;
                    }
                    if (x != u && x != v) merged.push_back(x);
                }
                if (merged.size() != a.size()) heap.push(DN((int)merged.size(), u));
                adj[u].assign(merged.begin(), merged.end());
            }
            colpat[v].swap(S);
        }
        adj.clear();
        adj.shrink_to_fit();
        // ---- pattern of L in the permuted numbering
        Lp.assign(N + 1, 0);
        for (int k = 0; k < N; ++k) Lp[k + 1] = Lp[k] + (int)colpat[perm[k]].size();
        nnzL = Lp[N];
        Li.resize(nnzL);
        for (int k = 0; k < N; ++k) {
            std::vector<int> &S = colpat[perm[k]];
            int *dst = Li.data() + Lp[k];
            for (size_t t = 0; t < S.size(); ++t) dst[t] = inv[S[t]];
            std::sort(dst, dst + S.size());
            std::vector<int>().swap(S);
        }
        colpat.clear();
        auto find_entry = [&](int col, int row) -> int {
            const int *b = Li.data() + Lp[col], *e = Li.data() + Lp[col + 1];
            const int *p = std::lower_bound(b, e, row);
            return (p != e && *p == row) ? (int)(p - Li.data()) : -1;
        };
        kmap.assign(row_ptr[m], -1);
        for (int i = 0; i < m; ++i)
            for (int q = row_ptr[i]; q < row_ptr[i + 1]; ++q) {
                const int a = inv[col_idx[q]], b = inv[n + i];
                kmap[q] = find_entry(std::min(a, b), std::max(a, b));
                if (kmap[q] < 0) return -1;
            }
        // ---- levels
        level.assign(N, 0);
        for (int k = 0; k < N; ++k)
            for (int p = Lp[k]; p < Lp[k + 1]; ++p) level[Li[p]] = std::max(level[Li[p]], level[k] + 1);
        n_levels = 0;
        for (int k = 0; k < N; ++k) n_levels = std::max(n_levels, level[k] + 1);
        // ---- terms, target-major first (column k updates entry (r_a, r_b) and pivot r_b for every pair of its rows
        //      r_a >= r_b), then a stable counting sort by the level of k: (level, target, k) order
        const int64_t n_targets = nnzL + N;
        {
            std::vector<int64_t> cnt(n_targets + 1, 0);
            std::vector<KktTerm> tm;
            for (int pass = 0; pass < 2; ++pass) {
                for (int k = 0; k < N; ++k) {
                    const int p0 = Lp[k], p1 = Lp[k + 1];
                    for (int pb = p0; pb < p1; ++pb) {
                        const int j = Li[pb];
                        const int64_t td = nnzL + j;
                        if (pass == 0)
                            ++cnt[td + 1];
                        else
                            tm[cnt[td]++] = KktTerm{pb, pb, k, (int)td};
                        int ptr = Lp[j];
                        for (int pa = pb + 1; pa < p1; ++pa) {
                            const int i = Li[pa];
                            while (Li[ptr] < i) ++ptr;  // (i, j) exists: fill of column k
                            if (pass == 0)
                                ++cnt[(int64_t)ptr + 1];
                            else
                                tm[cnt[ptr]++] = KktTerm{pa, pb, k, ptr};
                        }
                    }
                }
                if (pass == 0) {
                    for (int64_t t = 0; t < n_targets; ++t) cnt[t + 1] += cnt[t];
                    nterms = cnt[n_targets];
                    if (nterms > 0x3fffffffLL || n_targets >= kLastBit) return -1;
                    tm.resize(nterms);  // cnt[t] is now the running write position of target t
                }
            }
            std::vector<int64_t> lpos(n_levels + 1, 0);
            for (const KktTerm &u : tm) ++lpos[level[u.k] + 1];
            for (int l = 0; l < n_levels; ++l) lpos[l + 1] += lpos[l];
            terms.resize(nterms);
            for (const KktTerm &u : tm) terms[lpos[level[u.k]]++] = u;
        }
        n_fchunks = regroup(terms, [&](const KktTerm &u) { return level[u.k]; }, [](const KktTerm &u) { return u.t; },
                            fs_beg, fs_end, fmstep, fmchunk);
        {   // the last chunk of every pivot inverts it
            std::vector<int> last(N, -1);
            for (size_t q = 0; q < terms.size(); ++q)
                if (terms[q].t >= nnzL) last[terms[q].t - nnzL] = (int)q;   // terms of a chunk are contiguous: its first
            for (int l = 0; l < n_levels; ++l) {                             // term carries the flag
                for (int q = fs_beg[l]; q < fs_end[l]; ++q)
                    if (terms[q].t >= nnzL && last[terms[q].t - nnzL] == q) terms[q].t |= kLastBit;
                for (int c = fmstep[l]; c < fmstep[l + 1]; ++c) {
                    const int t = terms[fmchunk[c].begin].t;
                    if (t >= nnzL && last[t - nnzL] >= fmchunk[c].begin && last[t - nnzL] < fmchunk[c].end)
                        terms[fmchunk[c].begin].t |= kLastBit;
                }
            }
        }
        // ---- forward substitution: row-major items (target i, sources ascending), then by level of the source
        {
            std::vector<int> rptr(N + 1, 0);
            for (int64_t p = 0; p < nnzL; ++p) ++rptr[Li[p] + 1];
            for (int k = 0; k < N; ++k) rptr[k + 1] += rptr[k];
            std::vector<KktFwdItem> rl(nnzL);
            std::vector<int> pos(rptr.begin(), rptr.end() - 1);
            for (int k = 0; k < N; ++k)
                for (int p = Lp[k]; p < Lp[k + 1]; ++p) rl[pos[Li[p]]++] = KktFwdItem{p, perm[k], k, perm[Li[p]]};
            std::vector<int64_t> lpos(n_levels + 1, 0);
            for (const KktFwdItem &u : rl) ++lpos[level[u.k] + 1];
            for (int l = 0; l < n_levels; ++l) lpos[l + 1] += lpos[l];
            fwd.resize(nnzL);
            for (const KktFwdItem &u : rl) fwd[lpos[level[u.k]]++] = u;
        }
        n_wchunks = regroup(fwd, [&](const KktFwdItem &u) { return level[u.k]; }, [](const KktFwdItem &u) { return u.dst; },
                            ws_beg, ws_end, wmstep, wmchunk);
        // ---- backward substitution: entry (i, j) is applied when x_i is final, i.e. at the level of its row
        {
            bstep.assign(n_levels + 1, 0);
            for (int64_t p = 0; p < nnzL; ++p) ++bstep[level[Li[p]] + 1];
            for (int l = 0; l < n_levels; ++l) bstep[l + 1] += bstep[l];
            bwd.resize(nnzL);
            std::vector<int> pos(bstep.begin(), bstep.end() - 1);
            for (int k = 0; k < N; ++k)
                for (int p = Lp[k]; p < Lp[k + 1]; ++p)
                    bwd[pos[level[Li[p]]]++] = KktBwdItem{p, perm[Li[p]], perm[k], k};
        }
        auto work = [&](const std::vector<int> &sb, const std::vector<int> &se, const std::vector<int> &ms) {
            std::vector<int> w(n_levels + 1, 0);
            for (int l = 0; l < n_levels; ++l) w[l + 1] = w[l] + (se[l] - sb[l]) + (ms[l + 1] - ms[l]);
            return w;
        };
        {   // compulsory traffic per step (see f_distinct_reads)
            std::vector<int> stampW(nnzL + 1, -1), stampD(N, -1), stampV(N, -1), stampT(nnzL + N + 1, -1);
            size_t q = 0;
            for (int l = 0; l < n_levels; ++l)
                for (; q < terms.size() && level[terms[q].k] == l; ++q) {
                    const KktTerm &u = terms[q];
                    const int t = u.t & ~kLastBit;
                    if (stampW[u.a] != l) { stampW[u.a] = l; ++f_distinct_reads; }
                    if (stampW[u.b] != l) { stampW[u.b] = l; ++f_distinct_reads; }
                    if (stampD[u.k] != l) { stampD[u.k] = l; ++f_distinct_reads; }
                    if (stampT[t] != l) { stampT[t] = l; ++f_targets; }
                }
            std::fill(stampD.begin(), stampD.end(), -1);
            std::vector<int> stampDst(N, -1);
            q = 0;
            for (int l = 0; l < n_levels; ++l)
                for (; q < fwd.size() && level[fwd[q].k] == l; ++q) {
                    const KktFwdItem &u = fwd[q];
                    ++w_distinct_reads;                                   // W[pos]: every entry of L exactly once
                    if (stampV[u.src] != l) { stampV[u.src] = l; w_distinct_reads += 2; }   // v[src], 1/d[k]
                    if (stampDst[u.dst] != l) { stampDst[u.dst] = l; ++w_targets; }
                }
            std::fill(stampV.begin(), stampV.end(), -1);
            std::fill(stampDst.begin(), stampDst.end(), -1);
            for (int l = 0; l < n_levels; ++l)
                for (int p = bstep[l]; p < bstep[l + 1]; ++p) {
                    const KktBwdItem &u = bwd[p];
                    ++b_distinct_reads;
                    if (stampV[u.src] != l) { stampV[u.src] = l; ++b_distinct_reads; }
                    if (stampDst[u.dst] != l) { stampDst[u.dst] = l; b_distinct_reads += 1; ++b_targets; }   // 1/d[dst]
                }
        }
        plan(work(fs_beg, fs_end, fmstep), narrow, flaunch);
        plan(work(ws_beg, ws_end, wmstep), narrow, wlaunch);
        plan(bstep, narrow, blaunch);
        return 0;
    }

    // chunk = maximal run of items with the same (step, target).  Reorders the items of every step (singles first) and
    // fills the step tables; returns the number of chunks
    template <class T, class FL, class FT>
    int64_t regroup(std::vector<T> &items, FL lvl, FT tgt, std::vector<int> &sbeg, std::vector<int> &send,
                    std::vector<int> &mstep, std::vector<KktRange> &mchunk) {
        sbeg.assign(n_levels, 0);
        send.assign(n_levels, 0);
        mstep.assign(n_levels + 1, 0);
        mchunk.clear();
        std::vector<T> out;
        out.reserve(items.size());
        int64_t n_chunks = 0;
        size_t q = 0;
        for (int l = 0; l < n_levels; ++l) {
            const size_t q0 = q;
            while (q < items.size() && lvl(items[q]) == l) ++q;
            sbeg[l] = (int)out.size();
            // pass 1: singles
            for (size_t a = q0; a < q;) {
                size_t b = a + 1;
                while (b < q && tgt(items[b]) == tgt(items[a])) ++b;
                if (b - a == 1) out.push_back(items[a]);
                a = b;
            }
            send[l] = (int)out.size();
            // pass 2: multi-item chunks
            for (size_t a = q0; a < q;) {
                size_t b = a + 1;
                while (b < q && tgt(items[b]) == tgt(items[a])) ++b;
                ++n_chunks;
                if (b - a > 1) {
                    mchunk.push_back(KktRange{(int)out.size(), (int)(out.size() + (b - a))});
                    out.insert(out.end(), items.begin() + a, items.begin() + b);
                    longest_chunk = std::max(longest_chunk, (int)(b - a));
                }
                a = b;
            }
            mstep[l + 1] = (int)mchunk.size();
        }
        items.swap(out);
        return n_chunks;
    }

    void plan(const std::vector<int> &sptr, int narrow, std::vector<KktLaunch> &out) const {
        out.clear();
        int l = 0;
        while (l < n_levels) {
            const int items = sptr[l + 1] - sptr[l];
            if (items == 0) {
                ++l;
                continue;
            }
            if (items > narrow) {
                out.push_back(KktLaunch{l, l + 1, 0, items});
                ++l;
                continue;
            }
            int e = l, mx = 0;
            while (e < n_levels && sptr[e + 1] - sptr[e] <= narrow) {
                mx = std::max(mx, sptr[e + 1] - sptr[e]);
                ++e;
            }
            out.push_back(KktLaunch{l, e, e - l > 1 ? 1 : 0, std::max(mx, 1)});
            l = e;
        }
    }

    // ---- host reference of the device numerics (same lists, same order): used by the CPU self test ------------
    // W: nnzL values (in: assembled lower triangle, out: L D), diag: N assembled pivots (permuted; out: D),
    // invd: N out
    void factor_host(std::vector<double> &W, std::vector<double> &diag, std::vector<double> &invd) const {
        invd.resize(N);
        for (int j = 0; j < N; ++j) invd[j] = 1.0 / diag[j];
        auto apply = [&](int q0, int q1) {
            int t = terms[q0].t;
            const bool last = t & kLastBit;
            t &= ~kLastBit;
            double acc = 0.0;
            for (int q = q0; q < q1; ++q) acc += W[terms[q].a] * W[terms[q].b] * invd[terms[q].k];
            if (t >= nnzL) {
                diag[t - nnzL] -= acc;
                if (last) invd[t - nnzL] = 1.0 / diag[t - nnzL];
            } else {
                W[t] -= acc;
            }
        };
        for (int l = 0; l < n_levels; ++l) {
            for (int q = fs_beg[l]; q < fs_end[l]; ++q) apply(q, q + 1);
            for (int c = fmstep[l]; c < fmstep[l + 1]; ++c) apply(fmchunk[c].begin, fmchunk[c].end);
        }
    }
    // v (indexed by node id): right-hand side in, solution out
    void solve_host(const std::vector<double> &W, const std::vector<double> &invd, std::vector<double> &v) const {
        auto apply = [&](int q0, int q1) {
            double acc = 0.0;
            for (int q = q0; q < q1; ++q) acc += W[fwd[q].pos] * v[fwd[q].src] * invd[fwd[q].k];
            v[fwd[q0].dst] -= acc;
        };
        for (int l = 0; l < n_levels; ++l) {
            for (int q = ws_beg[l]; q < ws_end[l]; ++q) apply(q, q + 1);
            for (int c = wmstep[l]; c < wmstep[l + 1]; ++c) apply(wmchunk[c].begin, wmchunk[c].end);
        }
        for (int k = 0; k < N; ++k) v[perm[k]] *= invd[k];
        for (int l = n_levels - 1; l >= 0; --l)
            for (int q = bstep[l]; q < bstep[l + 1]; ++q) v[bwd[q].dst] -= invd[bwd[q].k] * W[bwd[q].pos] * v[bwd[q].src];
    }
};

}  // namespace asmb
