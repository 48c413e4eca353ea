// Host-side plan construction.  See plan.hpp for what each piece replaces in the reference.
#include "plan.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <unordered_map>

#include "../../include/tcc_b200.h"

namespace tcc {

static inline int popc(uint64_t x) { return __builtin_popcountll(x); }

int HostPlan::init(int n_qubits_, int n_elec_, int hcb_) {
    n_qubits = n_qubits_;
    n_elec = n_elec_;
    hcb = hcb_ ? 1 : 0;
    if (n_qubits <= 0 || n_elec < 0) {
        error = "n_qubits must be positive and n_elec non-negative";
        return TCC_ERR_BAD_ARGUMENT;
    }
    // ci_utils.py:19-20: 2**n_qubits must fit the unsigned index type
    if (n_qubits >= 64) {
        error = "Too many qubits: " + std::to_string(n_qubits) + ", try using complex128 datatype";
        return TCC_ERR_TOO_MANY_QUBITS;
    }
    if (!hcb && (n_qubits % 2)) {
        error = "n_qubits must be even for a fermionic (non-hcb) CI space";
        return TCC_ERR_BAD_ARGUMENT;
    }
    n = hcb ? n_qubits : n_qubits / 2;
    k = n_elec / 2;
    if (k > n) {
        error = "more electrons than orbitals";
        return TCC_ERR_BAD_ARGUMENT;
    }
    if (n > 31) {
        error = "Too many qubits: spin strings longer than 31 bits are not supported by the B200 engine";
        return TCC_ERR_TOO_MANY_QUBITS;
    }
    binom.assign(n + 2, std::vector<int64_t>(n + 2, 0));
    for (int i = 0; i <= n + 1; ++i) {
        binom[i][0] = 1;
        for (int j = 1; j <= i; ++j) binom[i][j] = binom[i - 1][j - 1] + (j <= i - 1 ? binom[i - 1][j] : 0);
    }
    nstr = binom[n][k];
    if (nstr >= (int64_t(1) << 31)) {
        error = "Too many qubits: more than 2^31 strings per spin";
        return TCC_ERR_TOO_MANY_QUBITS;
    }
    na = hcb ? 1 : nstr;
    nb = nstr;
    N = na * nb;
    bitpos.resize(n);
    std::iota(bitpos.begin(), bitpos.end(), 0);
    return 0;
}

uint32_t HostPlan::permute_bits(uint32_t s) const {
    if (identity_layout) return s;
    uint32_t t = 0;
    for (uint32_t x = s; x; x &= x - 1) t |= 1u << bitpos[__builtin_ctz(x)];
    return t;
}

int64_t HostPlan::colex_rank(uint32_t t) const {
    int64_t r = 0;
    int i = 1;
    while (t) {
        r += binom[__builtin_ctz(t)][i];
        ++i;
        t &= t - 1;
    }
    return r;
}

int64_t HostPlan::ref_rank_or_zero(uint64_t s) const {
    if (s >> n) return 0;
    if (popc(s) != k) return 0;
    return colex_rank(uint32_t(s));
}

static uint32_t zmask_of(uint32_t bits) {
    // bits strictly inside (i1,i2) and (i3,i4) of the sorted positions
    uint32_t z = 0;
    int pos[4], m = 0;
    for (uint32_t b = bits; b; b &= b - 1) pos[m++] = __builtin_ctz(b);
    for (int i = 0; i + 1 < m; i += 2)
        for (int p = pos[i] + 1; p < pos[i + 1]; ++p) z |= 1u << p;
    return z;
}

int HostPlan::find_or_build_hop(uint32_t cre, uint32_t ann) {
    uint64_t key = (uint64_t(cre) << 32) | ann;
    auto it = hop_index.find(key);
    if (it != hop_index.end()) return it->second;
    HopTable h;
    h.cre = cre;
    h.ann = ann;
    h.zmask = hcb ? 0u : zmask_of(cre | ann);
    for (int64_t i = 0; i < nstr; ++i) {
        uint32_t s = strings[i];
        if ((s & ann) != ann || (s & cre) != 0) continue;
        uint32_t t = s ^ (cre | ann);
        uint32_t par = popc(s & h.zmask) & 1;
        h.src.push_back(uint32_t(i));
        h.dst.push_back(uint32_t(rank(t)) | (par << 31));
    }
    hops.push_back(std::move(h));
    int id = int(hops.size()) - 1;
    hop_index[key] = id;
    return id;
}

int HostPlan::parse_op(const int32_t* idx, int len, int param, OpDesc* op) {
    auto fmt = [&]() {
        std::string s = "(";
        for (int i = 0; i < len; ++i) s += std::to_string(idx[i]) + (i + 1 < len ? ", " : "");
        return s + ")";
    };
    if (len != 2 && len != 4) {
        error = "Excitation " + fmt() + " not supported: only 2- and 4-index tuples";
        return TCC_ERR_BAD_EXCITATION;
    }
    uint64_t seen = 0;
    for (int i = 0; i < len; ++i) {
        if (idx[i] < 0 || idx[i] >= n_qubits) {
            error = "Excitation " + fmt() + " not supported: index out of range";
            return TCC_ERR_BAD_EXCITATION;
        }
        if (seen >> idx[i] & 1) {  // evolve_civector.py:92-93
            error = "Excitation " + fmt() + " not supported";
            return TCC_ERR_BAD_EXCITATION;
        }
        seen |= uint64_t(1) << idx[i];
    }
    *op = OpDesc();
    op->len = len;
    op->param = param;
    for (int i = 0; i < 4; ++i) op->idx[i] = i < len ? idx[i] : 0;
    uint64_t cre = 0, ann = 0;
    for (int i = 0; i < len / 2; ++i) cre |= uint64_t(1) << idx[i];
    for (int i = len / 2; i < len; ++i) ann |= uint64_t(1) << idx[i];
    uint64_t lo = (uint64_t(1) << n) - 1;
    op->cre_b = uint32_t(cre & lo);
    op->ann_b = uint32_t(ann & lo);
    op->cre_a = hcb ? 0u : uint32_t(cre >> n);
    op->ann_a = hcb ? 0u : uint32_t(ann >> n);
    if (popc(op->cre_a) != popc(op->ann_a) || popc(op->cre_b) != popc(op->ann_b)) {
        // the reference silently mis-computes these: partner falls outside the CI space
        error = "Excitation " + fmt() + " not supported: it does not conserve the electron number per spin";
        return TCC_ERR_BAD_EXCITATION;
    }
    if (hcb) {
        op->s0 = -1;  // fket_phase = -positive + negative, no fermionic sign (evolve_civector.py:97-101)
    } else if (len == 2) {
        op->s0 = -1;
    } else {
        int s_t = ((idx[2] < idx[3]) ^ (idx[1] < idx[0])) ? -1 : 1;
        op->s0 = -s_t;
    }
    bool has_a = (op->cre_a | op->ann_a) != 0, has_b = (op->cre_b | op->ann_b) != 0;
    op->kind = !has_a ? OP_BETA : (!has_b ? OP_ALPHA : OP_ALPHABETA);
    return 0;
}

void HostPlan::attach_hops(OpDesc* op) {
    op->hop_a = (op->cre_a | op->ann_a) ? find_or_build_hop(op->cre_a, op->ann_a) : -1;
    op->hop_b = (op->cre_b | op->ann_b) ? find_or_build_hop(op->cre_b, op->ann_b) : -1;
}

int64_t HostPlan::support_pairs(const OpDesc& op) const {
    int64_t la = op.hop_a < 0 ? na : int64_t(hops[op.hop_a].src.size());
    int64_t lb = op.hop_b < 0 ? nb : int64_t(hops[op.hop_b].src.size());
    return la * lb;
}

void HostPlan::expand_tables(const OpDesc& op, uint64_t* perm, int8_t* phase, int8_t* f2) const {
    // partner address for EVERY determinant, including the reference's "invalid string -> 0" rule;
    // everything here is in REFERENCE order
    uint64_t mask = 0;
    for (int i = 0; i < op.len; ++i) mask |= uint64_t(1) << op.idx[i];
    uint64_t lo = (uint64_t(1) << n) - 1;
    uint32_t mb = uint32_t(mask & lo), ma = hcb ? 0u : uint32_t(mask >> n);
    std::vector<int64_t> pa(na), pb(nb);
    for (int64_t a = 0; a < na; ++a) pa[a] = hcb ? 0 : ref_rank_or_zero(uint64_t(strings[ref2int[a]]) ^ ma);
    for (int64_t b = 0; b < nb; ++b) pb[b] = ref_rank_or_zero(uint64_t(strings[ref2int[b]]) ^ mb);
    for (int64_t a = 0; a < na; ++a)
        for (int64_t b = 0; b < nb; ++b) {
            perm[a * nb + b] = uint64_t(pa[a] * nstr + pb[b]);
            phase[a * nb + b] = 0;
            f2[a * nb + b] = 0;
        }
    const HopTable* ha = op.hop_a >= 0 ? &hops[op.hop_a] : nullptr;
    const HopTable* hb = op.hop_b >= 0 ? &hops[op.hop_b] : nullptr;
    int64_t la = ha ? int64_t(ha->src.size()) : na;
    int64_t lb = hb ? int64_t(hb->src.size()) : nb;
    auto to_ref = [&](int64_t x) { return int64_t(int2ref[x]); };
    for (int64_t i = 0; i < la; ++i) {
        int64_t a = ha ? to_ref(ha->src[i]) : i;
        int64_t a2 = ha ? to_ref(ha->dst[i] & 0x7fffffffu) : i;
        int sa = ha ? int(ha->dst[i] >> 31) : 0;
        for (int64_t j = 0; j < lb; ++j) {
            int64_t b = hb ? to_ref(hb->src[j]) : j;
            int64_t b2 = hb ? to_ref(hb->dst[j] & 0x7fffffffu) : j;
            int sb = hb ? int(hb->dst[j] >> 31) : 0;
            int p = (sa ^ sb) ? -op.s0 : op.s0;
            int64_t I = a * nb + b, J = a2 * nb + b2;
            phase[I] = int8_t(p);
            phase[J] = int8_t(-p);
            f2[I] = -1;
            f2[J] = -1;
        }
    }
}

// ---- batch schedule ----------------------------------------------------------------------------
int64_t HostPlan::max_class(int m) const {
    int lo = std::max(0, k - (n - m)), hi = std::min(m, k);
    int64_t best = 1;
    for (int e = lo; e <= hi; ++e) best = std::max(best, binom[m][e]);
    return best;
}

std::vector<std::vector<int>> HostPlan::group_ops(int tile_elems) const {
    if (sched_mode == 2) {  // both schedulers, the one with fewer batches (ties: first come, first served)
        HostPlan* self = const_cast<HostPlan*>(this);
        self->sched_mode = 0;
        std::vector<std::vector<int>> g0 = group_ops(tile_elems);
        self->sched_mode = 1;
        std::vector<std::vector<int>> g1 = group_ops(tile_elems);
        self->sched_mode = 2;
        return g1.size() < g0.size() ? g1 : g0;
    }
    // List scheduling.  A batch is described by its active sets (S_alpha, S_beta); given the sets, one pass over the
    // remaining excitations in order takes every excitation that lies inside them AND commutes with every excitation
    // skipped before it (disjoint spin-orbital sets => the generators commute), so the product of factors is unchanged.
    // sched_mode 0 grows the sets first come, first served: the union with each excitation met in order while the
    // tiles still fit on chip -- good for the generator order of the ansatz classes, where neighbours share orbitals.
    // sched_mode 1 chooses the sets: starting from the first remaining excitation (it must run now), the
    // union with that one of the next candidates is taken which lets the pass capture the most excitations, until no
    // candidate helps.  For amplitude-sorted lists (the reference's default UCCSD, uccsd.py:159-178), whose neighbours
    // are unrelated, this saves 7 % of the batches (492 instead of 530 for the screened H16 ansatz; 253 instead of
    // 258 for unscreened (16e,16o) UCCSD) -- the dependency chains of such a list are dense (two random opposite-spin
    // doubles share an orbital with probability 0.43), so few excitations are ever free to move -- and the batches it builds are
    // SLOWER per batch at (16e,16o) (forward 351 vs 341 ms, reverse 678 vs 664 ms for the whole sweep: other active sets,
    // other storage order), so sched_mode 0 stays the default.  sched_mode 2 runs both and keeps the shorter schedule.
    std::vector<std::vector<int>> groups;
    std::vector<int> remaining(ops.size());
    std::iota(remaining.begin(), remaining.end(), 0);
    const size_t window = 8192;
    const size_t max_ops = 256;
    const uint32_t all = n >= 32 ? 0xffffffffu : ((1u << n) - 1);
    int64_t side_cap = 64;
    while (side_cap * side_cap < 4 * int64_t(tile_elems)) ++side_cap;  // 2 * sqrt(tile_elems)
    if (hcb) side_cap = tile_elems;
    auto fits = [&](uint32_t nsa, uint32_t nsb) {
        const int64_t ca = max_class(hcb ? 0 : popc(nsa)), cb = max_class(popc(nsb));
        // keep tiles roughly square: a long thin tile has row segments too short to coalesce
        // sharded plans: log2(world) alpha orbitals must stay inactive, they carry the row layout of the batch
        const bool rows_ok = shard_world == 1 || popc(nsa) + shard_bits() <= n;
        return ca * cb <= tile_elems && ca <= side_cap && cb <= side_cap && rows_ok;
    };
    // number of excitations the in-order pass captures with FIXED active sets
    auto captured = [&](uint32_t sa, uint32_t sb) {
        uint32_t skip_a = 0, skip_b = 0;
        size_t cnt = 0;
        const size_t lim = std::min(remaining.size(), window);
        for (size_t i = 0; i < lim && cnt < max_ops; ++i) {
            const OpDesc& op = ops[remaining[i]];
            const uint32_t oa = op.cre_a | op.ann_a, ob = op.cre_b | op.ann_b;
            if (!((oa & skip_a) || (ob & skip_b)) && !(oa & ~sa) && !(ob & ~sb)) {
                ++cnt;
                continue;
            }
            skip_a |= oa;
            skip_b |= ob;
            if ((hcb || skip_a == all) && skip_b == all) break;
        }
        return cnt;
    };
    while (!remaining.empty()) {
        std::vector<int> batch, rest;
        uint32_t sa = 0, sb = 0, skip_a = 0, skip_b = 0;
        bool fixed_sets = false;
        if (sched_mode == 1) {
            const OpDesc& first = ops[remaining[0]];
            sa = first.cre_a | first.ann_a;
            sb = first.cre_b | first.ann_b;
            size_t best = captured(sa, sb);
            const size_t n_cand = 64;
            for (;;) {
                // candidates: the next excitations (in order) that are not inside the sets, not blocked by an excitation
                // outside them, and whose union still fits
                uint32_t ka = 0, kb = 0, best_a = 0, best_b = 0;
                size_t seen = 0, best_cnt = best;
                int best_grow = 1 << 30;
                const size_t lim = std::min(remaining.size(), window);
                for (size_t i = 0; i < lim && seen < n_cand; ++i) {
                    const OpDesc& op = ops[remaining[i]];
                    const uint32_t oa = op.cre_a | op.ann_a, ob = op.cre_b | op.ann_b;
                    const bool inside = !(oa & ~sa) && !(ob & ~sb);
                    const bool blocked = (oa & ka) || (ob & kb);
                    if (inside && !blocked) continue;
                    if (!blocked && fits(sa | oa, sb | ob)) {
                        ++seen;
                        const size_t c = captured(sa | oa, sb | ob);
                        const int grow = popc((sa | oa) ^ sa) + popc((sb | ob) ^ sb);
                        if (c > best_cnt || (c == best_cnt && c > best && grow < best_grow)) {
                            best_cnt = c;
                            best_grow = grow;
                            best_a = sa | oa;
                            best_b = sb | ob;
                        }
                    }
                    ka |= oa;
                    kb |= ob;
                    if ((hcb || ka == all) && kb == all) break;
                }
                if (best_cnt <= best) break;
                best = best_cnt;
                sa = best_a;
                sb = best_b;
            }
            fixed_sets = true;
        }
        bool closed = false;
        for (size_t i = 0; i < remaining.size(); ++i) {
            int j = remaining[i];
            if (closed || i >= window || batch.size() >= max_ops) {
                rest.push_back(j);
                continue;
            }
            const OpDesc& op = ops[j];
            uint32_t oa = op.cre_a | op.ann_a, ob = op.cre_b | op.ann_b;
            bool blocked = (oa & skip_a) || (ob & skip_b);
            if (!blocked) {
                uint32_t nsa = sa | oa, nsb = sb | ob;
                const bool ok = fixed_sets ? (nsa == sa && nsb == sb) : fits(nsa, nsb);
                if (ok || batch.empty()) {
                    batch.push_back(j);
                    sa = nsa;
                    sb = nsb;
                    continue;
                }
            }
            rest.push_back(j);
            skip_a |= oa;
            skip_b |= ob;
            if ((hcb || skip_a == all) && skip_b == all) closed = true;
        }
        groups.push_back(std::move(batch));
        remaining.swap(rest);
    }
    return groups;
}

void HostPlan::choose_layout(const std::vector<std::vector<int>>& groups, bool use_layout) {
    std::iota(bitpos.begin(), bitpos.end(), 0);
    identity_layout = true;
    if (!use_layout || groups.empty()) return;
    // orbitals that are active in many (large) batches go to the low bit positions: the strings of a
    // class then differ in the fastest-varying positions and sit next to each other in memory
    std::vector<double> weight(n, 0.0);
    for (const auto& g : groups) {
        uint32_t s = 0;
        for (int j : g) s |= ops[j].cre_a | ops[j].ann_a | ops[j].cre_b | ops[j].ann_b;
        for (int o = 0; o < n; ++o)
            if (s >> o & 1) weight[o] += double(g.size());
    }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return weight[x] > weight[y]; });
    for (int pos = 0; pos < n; ++pos) bitpos[order[pos]] = pos;
    for (int o = 0; o < n; ++o)
        if (bitpos[o] != o) identity_layout = false;
}

void HostPlan::build_strings() {
    strings.resize(nstr);
    ref2int.resize(nstr);
    int2ref.resize(nstr);
    std::vector<int> invpos(n);
    for (int o = 0; o < n; ++o) invpos[bitpos[o]] = o;
    uint64_t v = k ? (uint64_t(1) << k) - 1 : 0;
    for (int64_t i = 0; i < nstr; ++i) {
        uint32_t s = 0;
        for (uint64_t x = v; x; x &= x - 1) s |= 1u << invpos[__builtin_ctzll(x)];
        strings[i] = s;
        int64_t r = colex_rank(s);
        ref2int[r] = uint32_t(i);
        int2ref[i] = uint32_t(r);
        if (k) {
            uint64_t t = v | (v - 1);  // Gosper: next integer with the same popcount
            v = (t + 1) | (((~t & (t + 1)) - 1) >> (__builtin_ctzll(v) + 1));
        }
    }
}

void HostPlan::build_side(uint32_t S, int panel_target, SideBatch* side, bool single_string, bool pack_classes,
                          const RowLayout* rows) {
    side->S = S;
    side->m = popc(S);
    const int m = side->m;
    const int64_t count = single_string ? 1 : nstr;
    // classes in order of first appearance (= ascending first address), members in address order
    std::unordered_map<uint32_t, int> cls_of;
    std::vector<std::vector<uint32_t>> members;
    std::vector<uint32_t> cls_sigma;
    std::vector<int> cls_e;
    for (int64_t a = 0; a < count; ++a) {
        // sharded alpha side: only the rows of this rank, addressed by their local row
        if (rows && rows->owner[a] != shard_rank) continue;
        uint32_t s = single_string ? 0u : strings[a];
        uint32_t sig = s & ~S;
        auto it = cls_of.find(sig);
        int id;
        if (it == cls_of.end()) {
            id = int(members.size());
            cls_of.emplace(sig, id);
            members.emplace_back();
            cls_sigma.push_back(sig);
            cls_e.push_back(popc(s & S));
        } else {
            id = it->second;
        }
        members[id].push_back(rows ? rows->local_row[a] : uint32_t(a));
    }
    side->addr.clear();
    side->ord.clear();
    side->addr_sorted.clear();
    side->sigma.clear();
    side->panels.clear();
    side->max_panel = 0;
    side->max_classes = 0;
    // Classes per panel.  A side without active orbitals packs single-string classes up to the panel target.  On an
    // active side the small classes (few or many active electrons) are packed 2 or 4 to a panel (at most 16 sub-blocks
    // per tile, each with its own group of threads of the CTA), as long as the
    // panel stays within the largest class: the tile count drops by a third and a tile is never mostly idle threads.
    // The kernel applies the pair list of the class pair to each (alpha class, beta class) sub-block of the tile.
    std::vector<int> n_cls_e(m + 1, 0), g_of_e(m + 1, 1);
    for (size_t c = 0; c < members.size(); ++c) n_cls_e[cls_e[c]] += 1;
    int64_t cmax = 1;
    for (int e = 0; e <= m; ++e)
        if (n_cls_e[e]) cmax = std::max<int64_t>(cmax, binom[m][e]);
    side->ld_of_e.assign(m + 1, 0);
    for (int e = 0; e <= m; ++e) {
        if (!n_cls_e[e]) continue;
        int g = 1;
        if (m > 0 && pack_classes)
            while (g * 2 <= 4 && g * 2 <= n_cls_e[e] && int64_t(g) * 2 * binom[m][e] <= cmax) g *= 2;
        g_of_e[e] = g;
        side->ld_of_e[e] = int(g * binom[m][e]) | 1;
    }
    // largest panels first: tiles of similar size are then contiguous rectangles of the panel grid
    std::vector<int> e_order;
    for (int e = 0; e <= m; ++e) e_order.push_back(e);
    std::stable_sort(e_order.begin(), e_order.end(),
                     [&](int x, int y) { return g_of_e[x] * binom[m][x] > g_of_e[y] * binom[m][y]; });
    for (int e : e_order) {
        std::vector<int> ids;
        for (size_t c = 0; c < members.size(); ++c)
            if (cls_e[c] == e) ids.push_back(int(c));
        if (ids.empty()) continue;
        const int csize = int(members[ids[0]].size());
        const int per_panel = m > 0 ? g_of_e[e] : std::max(1, std::min(panel_target, 4096));
        for (size_t p0 = 0; p0 < ids.size(); p0 += per_panel) {
            size_t p1 = std::min(ids.size(), p0 + per_panel);
            PanelDesc pd;
            pd.off = uint32_t(side->addr.size());
            pd.class_off = uint32_t(side->sigma.size());
            pd.G = uint16_t(p1 - p0);
            pd.c = uint16_t(csize);
            pd.e = uint16_t(e);
            pd.ld = uint16_t(side->ld_of_e[e]);
            pd.Gn = uint16_t(m > 0 ? g_of_e[e] : 1);
            pd.pad = 0;
            for (size_t q = p0; q < p1; ++q) {
                side->sigma.push_back(cls_sigma[ids[q]]);
                for (uint32_t a : members[ids[q]]) side->addr.push_back(a);
            }
            const int rows = int(pd.G) * csize;
            std::vector<uint16_t> ord(rows);
            std::iota(ord.begin(), ord.end(), uint16_t(0));
            const uint32_t* base = side->addr.data() + pd.off;
            std::sort(ord.begin(), ord.end(), [&](uint16_t x, uint16_t y) { return base[x] < base[y]; });
            side->ord.insert(side->ord.end(), ord.begin(), ord.end());
            for (uint16_t x : ord) side->addr_sorted.push_back(base[x]);
            side->panels.push_back(pd);
            side->max_panel = std::max(side->max_panel, rows);
            side->max_classes = std::max(side->max_classes, int(pd.G));
        }
    }
}

static uint32_t compress_bits(uint32_t bits, const std::vector<int>& act) {
    uint32_t t = 0;
    for (size_t i = 0; i < act.size(); ++i)
        if (bits >> act[i] & 1) t |= 1u << i;
    return t;
}

void HostPlan::active_sets(const std::vector<int>& group, int tile_elems, uint32_t* sa_out, uint32_t* sb_out) const {
    uint32_t sa = 0, sb = 0;
    for (int j : group) {
        sa |= ops[j].cre_a | ops[j].ann_a;
        sb |= ops[j].cre_b | ops[j].ann_b;
    }
    // Pad the active sets with unused orbitals (lowest storage positions first) while the tiles still fit:
    // small active sets would split the space into a huge number of tiny classes.
    {
        std::vector<int> by_pos(n);
        std::iota(by_pos.begin(), by_pos.end(), 0);
        std::sort(by_pos.begin(), by_pos.end(), [&](int x, int y) { return bitpos[x] < bitpos[y]; });
        bool grew = true;
        while (grew) {
            grew = false;
            // grow the smaller side first so that tiles stay roughly square
            for (int pass = 0; pass < 2 && !grew; ++pass) {
                bool beta = (popc(sb) <= popc(sa)) ? pass == 0 : pass == 1;
                if (!beta && (hcb || sa == 0)) continue;  // a side without hops stays unexpanded
                if (beta && sb == 0) continue;
                uint32_t& s = beta ? sb : sa;
                // a sharded plan keeps log2(world) alpha orbitals out of every active set: they carry the row layout
                if (!beta && shard_world > 1 && popc(sa) + shard_bits() >= n) continue;
                for (int o : by_pos) {
                    if (s >> o & 1) continue;
                    uint32_t ns = s | (1u << o);
                    int64_t elems = beta ? max_class(hcb ? 0 : popc(sa)) * max_class(popc(ns))
                                         : max_class(popc(ns)) * max_class(popc(sb));
                    if (elems <= tile_elems) {
                        s = ns;
                        grew = true;
                    }
                    break;
                }
            }
        }
    }
    *sa_out = sa;
    *sb_out = sb;
}

// order of the pairs of one excitation: groups of `group` consecutive entries (one shared-memory wavefront of the
// access width in use) whose first elements fall into different banks, and likewise the second elements, as far as a
// greedy search in a small window finds them.  The pairs of an excitation are independent, so any order is valid.
static void bank_ordered(const std::vector<uint32_t>& raw, int group, std::vector<char>* taken, std::vector<uint32_t>* out) {
    const uint32_t gm = uint32_t(group - 1);
    taken->assign(raw.size(), 0);
    size_t head = 0, emitted = 0;
    while (emitted < raw.size()) {
        uint32_t m1 = 0, m2 = 0;
        for (int kk = 0; kk < group && emitted < raw.size(); ++kk) {
            while (head < raw.size() && (*taken)[head]) ++head;
            size_t pick = head;
            const size_t lim = std::min(raw.size(), head + 96);
            for (size_t c = head; c < lim; ++c) {
                if ((*taken)[c]) continue;
                const uint32_t b1 = 1u << (raw[c] & gm), b2 = 1u << ((raw[c] >> 15) & gm);
                if (!(m1 & b1) && !(m2 & b2)) {
                    pick = c;
                    break;
                }
            }
            (*taken)[pick] = 1;
            m1 |= 1u << (raw[pick] & gm);
            m2 |= 1u << ((raw[pick] >> 15) & gm);
            out->push_back(raw[pick]);
            ++emitted;
        }
    }
}

// Conflict-free groups by bipartite matching (lean streams).  A group of `group` consecutive stream entries is what one
// quarter-warp (16-byte elements, group = 8) or half-warp (8-byte elements, group = 16) sends to shared memory in one
// wavefront -- if the first elements of the group fall into `group` different banks and the second elements too; ONE
// clash costs the whole group a second wavefront, which is why the greedy order above still leaves 1.43 wavefronts per
// ideal one (host model = ncu, profiles/r02_reverse_kernel_breakdown.md).  Here the entries are bucketed by (bank of the
// first element, bank of the second element); a group is a perfect matching between the first-element banks and the
// second-element banks over the non-empty buckets (Kuhn's augmenting paths on a group x group graph, fullest buckets
// first so that the buckets drain evenly).  When no perfect matching is left, the remaining entries form maximum
// matchings padded with invalid entries while the stream has spare slots (`capacity` - raw.size(): the last round of an
// excitation is rarely full), and only then groups with clashes.  The stream is emitted in whole groups.
static void bank_matched(const std::vector<uint32_t>& raw, int group, size_t capacity, size_t last_round_at, std::vector<uint32_t>* out) {
    const uint32_t gm = uint32_t(group - 1);
    const size_t start = out->size();
    const int G = group;
    std::vector<std::vector<uint32_t>> bucket(size_t(G) * G);
    {
        // Which end of a pair is "first" is free: (i1, i2, sign) and (i2, i1, -sign) are the same rotation, bit for bit
        // (x' = c x + p y, y' = c y - p x with x <-> y, p -> -p).  Each pair is oriented so that the banks are used
        // evenly as first and as second element: the number of perfect groups is bounded by the rarest bank.
        std::vector<int> c1(G, 0), c2(G, 0);
        for (uint32_t e : raw) {
            const uint32_t i1 = e & 0x7fffu, i2 = (e >> 15) & 0x7fffu, sg = (e >> 30) & 1u;
            const int x = int(i1 & gm), y = int(i2 & gm);
            if (std::max(c1[y], c2[x]) < std::max(c1[x], c2[y])) {
                e = i2 | (i1 << 15) | ((sg ^ 1u) << 30);
                ++c1[y];
                ++c2[x];
            } else {
                ++c1[x];
                ++c2[y];
            }
            bucket[size_t(e & gm) * G + ((e >> 15) & gm)].push_back(e);
        }
    }
    size_t left = raw.size();
    size_t spare = capacity > raw.size() ? capacity - raw.size() : 0;
    std::vector<int> match2(G), order(G), cand(G);
    std::vector<char> seen(G);
    // Kuhn: try to give first-bank b1 a second-bank, fullest bucket first
    struct Kuhn {
        int G;
        std::vector<std::vector<uint32_t>>* bucket;
        std::vector<int>* match2;
        std::vector<char>* seen;
        bool go(int b1) {
            int idx[32];
            int n = 0;
            for (int b2 = 0; b2 < G; ++b2)
                if (!(*bucket)[size_t(b1) * G + b2].empty() && !(*seen)[b2]) idx[n++] = b2;
            std::sort(idx, idx + n, [&](int x, int y) { return (*bucket)[size_t(b1) * G + x].size() > (*bucket)[size_t(b1) * G + y].size(); });
            for (int i = 0; i < n; ++i) {
                const int b2 = idx[i];
                if ((*seen)[b2]) continue;
                (*seen)[b2] = 1;
                if ((*match2)[b2] < 0 || go((*match2)[b2])) {
                    (*match2)[b2] = b1;
                    return true;
                }
            }
            return false;
        }
    } kuhn{G, &bucket, &match2, &seen};
    while (left > 0) {
        // first-element banks with the fewest alternatives go first
        for (int b1 = 0; b1 < G; ++b1) {
            order[b1] = b1;
            int n = 0;
            for (int b2 = 0; b2 < G; ++b2) n += !bucket[size_t(b1) * G + b2].empty();
            cand[b1] = n;
        }
        std::sort(order.begin(), order.end(), [&](int x, int y) { return cand[x] < cand[y]; });
        std::fill(match2.begin(), match2.end(), -1);
        int matched = 0;
        for (int i = 0; i < G; ++i) {
            if (cand[order[i]] == 0) continue;
            std::fill(seen.begin(), seen.end(), 0);
            if (kuhn.go(order[i])) ++matched;
        }
        if (matched == 0) break;
        size_t pads = size_t(G - matched);
        for (int b2 = 0; b2 < G; ++b2)
            if (match2[b2] >= 0) {
                std::vector<uint32_t>& bk = bucket[size_t(match2[b2]) * G + b2];
                out->push_back(bk.back());
                bk.pop_back();
                --left;
            }
        if (pads == 0 || left == 0) continue;
        // an incomplete group: invalid entries may only stand in the LAST round of the stream (a lane with an entry in
        // round r has entries in all rounds before it -- the kernels rely on that) and only while slots are spare;
        // otherwise the group is completed with entries that clash
        const bool last_round = out->size() - start >= last_round_at;
        if (last_round && pads <= spare) {
            for (size_t i = 0; i < pads; ++i) out->push_back(0u);
            spare -= pads;
        } else {
            for (auto& bk : bucket) {
                while (pads > 0 && !bk.empty()) {
                    out->push_back(bk.back());
                    bk.pop_back();
                    --left;
                    --pads;
                }
                if (pads == 0) break;
            }
        }
    }
    if (left > 0) {  // what could not be grouped without clashes and without padding, in bucket order
        for (auto& bk : bucket)
            for (uint32_t e : bk) out->push_back(e);
    }
}

void HostPlan::build_batch(const std::vector<int>& group, int tile_elems, Batch* out, int lean_threads) {
    const uint32_t sa = out->sa, sb = out->sb;  // set by finalize (active_sets)
    int64_t ca = max_class(hcb ? 0 : popc(sa)), cb = max_class(popc(sb));
    int64_t pa = ca, pb = cb;
    // spend the remaining capacity: first make the contiguous (beta) direction at least 128 wide, then rows
    while (pa * pb * 2 <= tile_elems) {
        if (pb < 128 && pb * 2 <= nb)
            pb *= 2;
        else if (pa * 2 <= na)
            pa *= 2;
        else if (pb * 2 <= nb)
            pb *= 2;
        else
            break;
    }
    // class packing needs hops on both sides (a side without active orbitals takes the repeat path of the kernel)
    const bool pack = popc(sb) > 0 && (hcb || popc(sa) > 0) &&
                      (std::getenv("TCC_B200_PACK") ? std::atoi(std::getenv("TCC_B200_PACK")) != 0 : true);
    build_side(sa, int(pa), &out->A, hcb != 0, pack, shard_world > 1 ? &layouts[out->layout] : nullptr);
    build_side(sb, int(pb), &out->B, false, pack, nullptr);
    auto make_tables = [&](SideBatch& side) {
        const int m = side.m;
        std::vector<int> act;  // active orbitals sorted by storage bit position
        for (int o = 0; o < n; ++o)
            if (side.S >> o & 1) act.push_back(o);
        std::sort(act.begin(), act.end(), [&](int x, int y) { return bitpos[x] < bitpos[y]; });
        side.tab.clear();
        side.tab_index.assign(side.hops.size() * size_t(m + 1) * 2, 0);
        for (size_t h = 0; h < side.hops.size(); ++h) {
            const LocalHop& lh = side.hops[h];
            for (int e = 0; e <= m; ++e) {
                int32_t off = int32_t(side.tab.size()), cnt = 0;
                if (e >= popc(lh.ann) && e <= m) {
                    // all patterns of e bits among m compressed positions, ascending = class-local order
                    uint32_t v = e ? (1u << e) - 1 : 0;
                    int64_t total = binom[m][e];
                    for (int64_t r = 0; r < total; ++r) {
                        uint32_t bits = 0;
                        for (uint32_t x = v; x; x &= x - 1) bits |= 1u << act[__builtin_ctz(x)];
                        if ((bits & lh.ann) == lh.ann && (bits & lh.cre) == 0) {
                            uint32_t nb_ = bits ^ (lh.cre | lh.ann);
                            uint32_t dst = uint32_t(colex_rank(compress_bits(nb_, act)));
                            uint32_t sg = popc(bits & lh.zmask) & 1;
                            side.tab.push_back(uint32_t(r) | (dst << 15) | (sg << 30));
                            ++cnt;
                        }
                        if (e) {
                            uint32_t t = v | (v - 1);
                            v = (t + 1) | (((~t & (t + 1)) - 1) >> (__builtin_ctz(v) + 1));
                        }
                    }
                }
                side.tab_index[(h * (m + 1) + e) * 2 + 0] = off;
                side.tab_index[(h * (m + 1) + e) * 2 + 1] = cnt;
            }
        }
        if (side.tab.empty()) side.tab.push_back(0);
        if (side.tab_index.empty()) side.tab_index.assign(2, 0);
    };
    auto local_hop = [&](SideBatch& side, uint32_t cre, uint32_t ann) -> int {
        if (!(cre | ann)) return -1;
        for (size_t h = 0; h < side.hops.size(); ++h)
            if (side.hops[h].cre == cre && side.hops[h].ann == ann) return int(h);
        side.hops.push_back(LocalHop{cre, ann, hcb ? 0u : zmask_of(cre | ann)});
        return int(side.hops.size()) - 1;
    };
    out->A.hops.clear();
    out->B.hops.clear();
    out->ops.clear();
    for (int j : group) {
        const OpDesc& op = ops[j];
        BatchOp bo;
        bo.op_index = j;
        bo.h_a = local_hop(out->A, op.cre_a, op.ann_a);
        bo.h_b = local_hop(out->B, op.cre_b, op.ann_b);
        bo.zspec_a = bo.h_a >= 0 ? (out->A.hops[bo.h_a].zmask & ~sa) : 0u;
        bo.zspec_b = bo.h_b >= 0 ? (out->B.hops[bo.h_b].zmask & ~sb) : 0u;
        bo.s0 = op.s0;
        out->ops.push_back(bo);
    }
    make_tables(out->A);
    make_tables(out->B);

    // fully expanded pair lists per (excitation, e_alpha, e_beta)
    const int ma = out->A.m, mb = out->B.m;
    out->ld_fixed = out->B.max_panel | 1;
    out->plist.clear();
    out->plist_index.assign(out->ops.size() * size_t(ma + 1) * (mb + 1) * 2, 0);
    std::vector<char> has_ea(ma + 1, 0), has_eb(mb + 1, 0);
    for (const auto& pd : out->A.panels) has_ea[pd.e] = 1;
    for (const auto& pd : out->B.panels) has_eb[pd.e] = 1;
    struct Ent {
        uint32_t s, d, sg;
    };
    auto side_entries = [&](const SideBatch& side, int h, int e, std::vector<Ent>* v) {
        v->clear();
        if (side.m == 0) {  // repeated by the kernel
            v->push_back(Ent{0, 0, 0});
            return;
        }
        if (h < 0) {
            int64_t c = binom[side.m][e];
            for (int64_t i = 0; i < c; ++i) v->push_back(Ent{uint32_t(i), uint32_t(i), 0});
            return;
        }
        const int32_t off = side.tab_index[(size_t(h) * (side.m + 1) + e) * 2];
        const int32_t cnt = side.tab_index[(size_t(h) * (side.m + 1) + e) * 2 + 1];
        for (int32_t i = 0; i < cnt; ++i) {
            uint32_t t = side.tab[off + i];
            v->push_back(Ent{t & 0x7fffu, (t >> 15) & 0x7fffu, t >> 30});
        }
    };
    std::vector<Ent> ea_list, eb_list;
    std::vector<uint32_t> raw;
    std::vector<char> taken;
    // lean streams: only for batches with active orbitals on both sides (the repeat path keeps the plain lists)
    const bool lean = lean_threads > 0 && ma > 0 && mb > 0;
    out->lean_threads = lean ? lean_threads : 0;
    out->stream8.clear();
    out->stream16.clear();
    out->stream_index.assign(lean ? out->plist_index.size() : 0, 0);
    std::vector<int> ga_nom(ma + 1, 1), gb_nom(mb + 1, 1);
    for (const auto& pd : out->A.panels) ga_nom[pd.e] = std::max<int>(1, pd.Gn);
    for (const auto& pd : out->B.panels) gb_nom[pd.e] = std::max<int>(1, pd.Gn);
    const bool bank_order = std::getenv("TCC_B200_BANK_ORDER") ? std::atoi(std::getenv("TCC_B200_BANK_ORDER")) != 0 : true;
    const bool bank_match_fwd = bank_order && (std::getenv("TCC_B200_BANK_MATCH_FWD") ? std::atoi(std::getenv("TCC_B200_BANK_MATCH_FWD")) != 0 : true);
    const bool bank_match = bank_order && (std::getenv("TCC_B200_BANK_MATCH") ? std::atoi(std::getenv("TCC_B200_BANK_MATCH")) != 0 : true);
    for (size_t t = 0; t < out->ops.size(); ++t) {
        const BatchOp& bo = out->ops[t];
        for (int ea = 0; ea <= ma; ++ea) {
            if (!has_ea[ea]) continue;
            side_entries(out->A, bo.h_a, ea, &ea_list);
            for (int eb = 0; eb <= mb; ++eb) {
                if (!has_eb[eb]) continue;
                side_entries(out->B, bo.h_b, eb, &eb_list);
                const uint32_t ld = mb > 0 ? uint32_t(out->B.ld_of_e[eb]) : uint32_t(out->ld_fixed);
                const size_t slot = ((t * (ma + 1) + ea) * (mb + 1) + eb) * 2;
                out->plist_index[slot] = int32_t(out->plist.size());
                out->plist_index[slot + 1] = int32_t(ea_list.size() * eb_list.size());
                raw.clear();
                for (const Ent& a : ea_list)
                    for (const Ent& b : eb_list)
                        raw.push_back((a.s * ld + b.s) | ((a.d * ld + b.d) << 15) | ((a.sg ^ b.sg) << 30));
                if (bank_match_fwd && lean && raw.size() > 16) {
                    // conflict-free half-warp groups by matching, completed with clashing entries where no perfect
                    // matching is left (host model: 1.43 wavefronts per half-warp access instead of 1.77).  Padding the
                    // incomplete groups with skip entries up to the end of the last round would reach 1.07, but the
                    // guard and the 9 % more slots cost more than the clashes (measured: 339 ms vs 328 ms forward sweep)
                    bank_matched(raw, 16, raw.size(), 0, &out->plist);
                } else if (bank_order && raw.size() > 16)
                    bank_ordered(raw, 16, &taken, &out->plist);
                else
                    out->plist.insert(out->plist.end(), raw.begin(), raw.end());
                if (lean) {
                    // per-thread stream of the class pair: `slots` rounds of tps entries, padded with invalid entries
                    const int tps = std::max(1, lean_threads / (ga_nom[ea] * gb_nom[eb]));
                    const int slots = int((raw.size() + tps - 1) / tps);
                    out->stream_index[slot] = int32_t(out->stream8.size());
                    out->stream_index[slot + 1] = slots;
                    for (int width = 0; width < 2; ++width) {
                        std::vector<uint32_t>& dst = width == 0 ? out->stream8 : out->stream16;
                        const size_t before = dst.size();
                        const int grp = width == 0 ? 16 : 8;
                        if (bank_match && width == 1 && raw.size() > size_t(grp) && tps % grp == 0)  // the 16-byte streams of the reverse sweep
                            bank_matched(raw, grp, size_t(slots) * tps, size_t(slots - 1) * tps, &dst);
                        else if (bank_order && raw.size() > 8)
                            bank_ordered(raw, grp, &taken, &dst);
                        else
                            dst.insert(dst.end(), raw.begin(), raw.end());
                        for (size_t i = before; i < dst.size(); ++i)
                            if (dst[i]) dst[i] |= 0x80000000u;  // padding entries of bank_matched stay invalid
                        dst.resize(before + size_t(slots) * tps, 0u);
                    }
                }
            }
        }
    }
    if (out->plist.empty()) out->plist.push_back(0);
}

// Row layouts of a sharded plan.  A batch can run in a layout whose orbitals T are not active on its alpha side.
// The layout is kept while it fits; when a batch needs a change, the new T takes the admissible orbitals whose
// next alpha activity lies farthest ahead (the UCCSD generator order sweeps the occupied index slowly, so a
// whole sweep needs a handful of layouts).
void HostPlan::choose_row_layouts() {
    layouts.clear();
    segments.clear();
    int t = 0;
    while ((1 << t) < shard_world) ++t;
    const int nb_ = int(batches.size());
    auto make_layout = [&](uint32_t T) {
        RowLayout L;
        L.T = T;
        L.owner.assign(nstr, 0);
        L.local_row.assign(nstr, 0);
        L.rows.assign(shard_world, 0);
        std::vector<int> torb;
        for (int o = 0; o < n; ++o)
            if (T >> o & 1) torb.push_back(o);
        for (int64_t a = 0; a < nstr; ++a) {
            int pat = 0;
            for (size_t i = 0; i < torb.size(); ++i)
                if (strings[a] >> torb[i] & 1) pat |= 1 << i;
            L.owner[a] = uint8_t(pat);
            L.local_row[a] = uint32_t(L.rows[pat]++);
        }
        return L;
    };
    if (shard_world == 1 || hcb) {
        layouts.push_back(make_layout(0));
        for (auto& b : batches) b.layout = 0;
        if (nb_) segments.push_back(Segment{0, 0, nb_});
        max_local_rows = hcb ? 1 : nstr;
        return;
    }
    // next_use[b][o]: first batch >= b whose alpha active set holds orbital o
    std::vector<std::vector<int>> next_use(nb_ + 1, std::vector<int>(n, nb_));
    for (int b = nb_ - 1; b >= 0; --b)
        for (int o = 0; o < n; ++o) next_use[b][o] = (batches[b].sa >> o & 1) ? b : next_use[b + 1][o];
    std::map<uint32_t, int> layout_of;
    uint32_t cur = 0;
    bool have = false;
    for (int b = 0; b < nb_; ++b) {
        if (!have || (cur & batches[b].sa)) {
            std::vector<int> cand;
            for (int o = 0; o < n; ++o)
                if (!(batches[b].sa >> o & 1)) cand.push_back(o);
            std::stable_sort(cand.begin(), cand.end(), [&](int x, int y) { return next_use[b][x] > next_use[b][y]; });
            cur = 0;
            for (int i = 0; i < t && i < int(cand.size()); ++i) cur |= 1u << cand[i];
            have = true;
        }
        auto it = layout_of.find(cur);
        if (it == layout_of.end()) {
            it = layout_of.emplace(cur, int(layouts.size())).first;
            layouts.push_back(make_layout(cur));
        }
        batches[b].layout = it->second;
        if (segments.empty() || segments.back().layout != it->second)
            segments.push_back(Segment{it->second, b, b + 1});
        else
            segments.back().batch_hi = b + 1;
    }
    if (layouts.empty()) layouts.push_back(make_layout(t ? (1u << t) - 1 : 0));
    // layouts of the sharded sigma: the layout the forward sweep ends in, and two more with disjoint T
    sigma_layout[0] = segments.empty() ? 0 : segments.back().layout;
    uint32_t used = layouts[sigma_layout[0]].T;
    for (int i = 1; i < 3; ++i) {
        uint32_t T = 0;
        int cnt = 0;
        for (int o = n - 1; o >= 0 && cnt < t; --o)
            if (!(used >> o & 1)) {
                T |= 1u << o;
                ++cnt;
            }
        used |= T;
        auto it = layout_of.find(T);
        if (it == layout_of.end()) {
            it = layout_of.emplace(T, int(layouts.size())).first;
            layouts.push_back(make_layout(T));
        }
        sigma_layout[i] = it->second;
    }
    max_local_rows = 0;
    for (const auto& L : layouts) max_local_rows = std::max(max_local_rows, L.rows[shard_rank]);
}

void HostPlan::build_sigma_pass_links() {
    if (shard_world == 1 || links.empty()) return;
    for (int pass = 0; pass < 3; ++pass) {
        const RowLayout& L = layouts[sigma_layout[pass]];
        auto pass_of = [&](int p, int q) {
            for (int i = 0; i < 3; ++i) {
                const uint32_t T = layouts[sigma_layout[i]].T;
                if (!(T >> p & 1) && !(T >> q & 1)) return i;
            }
            return -1;  // cannot happen: a pair touches at most two of three disjoint sets
        };
        const int64_t rows = L.rows[shard_rank];
        // longest filtered list
        int longest = 0;
        for (int64_t a = 0; a < nstr; ++a) {
            if (L.owner[a] != shard_rank) continue;
            int c = 0;
            for (int l = 0; l < ml; ++l) {
                const Link& lk = links[size_t(a) * ml + l];
                if (lk.sgn != 0 && (lk.pq / n == lk.pq % n ? pass == 0 : pass_of(lk.pq / n, lk.pq % n) == pass)) ++c;
            }
            longest = std::max(longest, c);
        }
        const int aml = (longest + 7) / 8 * 8;
        sigma_aml[pass] = aml;
        sigma_links[pass].assign(size_t(rows) * aml, Link{0, 0, 0, 0});
        for (int64_t a = 0; a < nstr; ++a) {
            if (L.owner[a] != shard_rank) continue;
            Link* dst = sigma_links[pass].data() + size_t(L.local_row[a]) * aml;
            int c = 0;
            for (int l = 0; l < ml; ++l) {
                const Link& lk = links[size_t(a) * ml + l];
                if (lk.sgn == 0) continue;
                const int p = lk.pq / n, q = lk.pq % n;
                if (!(p == q ? pass == 0 : pass_of(p, q) == pass)) continue;
                // E_pq with p, q outside T keeps the occupation pattern on T: the target row is on this rank
                dst[c++] = Link{L.local_row[lk.addr], lk.pq, lk.sgn, 0};
            }
        }
    }
}

void HostPlan::finalize(int tile_elems, bool use_layout, int rank, int world, int lean_threads, bool use_dyn) {
    tile_elems = std::max(tile_elems, 64);
    shard_rank = rank;
    shard_world = world;
    if (world > 1 && shard_bits() + 2 > n) {
        error = "too many ranks for this active space: log2(world) + 2 alpha orbitals are needed";
        return;
    }
    std::vector<std::vector<int>> groups = group_ops(tile_elems);
    choose_layout(groups, use_layout);
    build_strings();
    for (auto& op : ops) attach_hops(&op);
    batches.clear();
    batches.resize(groups.size());
    for (size_t g = 0; g < groups.size(); ++g) active_sets(groups[g], tile_elems, &batches[g].sa, &batches[g].sb);
    choose_row_layouts();
    {
        // the batches are independent (build_batch reads the plan and writes its own Batch): a few host threads
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned nt = unsigned(std::min<size_t>(std::min(4u, hw), groups.size()));
        if (nt <= 1) {
            for (size_t g = 0; g < groups.size(); ++g) build_batch(groups[g], tile_elems, &batches[g], lean_threads);
        } else {
            std::atomic<size_t> next{0};
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nt; ++t)
                pool.emplace_back([&]() {
                    for (size_t g = next.fetch_add(1); g < groups.size(); g = next.fetch_add(1))
                        build_batch(groups[g], tile_elems, &batches[g], lean_threads);
                });
            for (auto& th : pool) th.join();
        }
    }
    if (use_dyn) build_dyn_layouts();
}

uint32_t HostPlan::rank_in_order(const std::vector<int>& order, uint32_t s) const {
    int64_t r = 0;
    int idx = 1;
    for (int pos = 0; pos < n; ++pos)
        if (s >> order[pos] & 1) {
            r += binom[pos][idx];
            ++idx;
        }
    return uint32_t(r);
}

// Per-batch column orders and the gather / scatter tables that go with them (plan.hpp, DYNAMIC COLUMN LAYOUT).
void HostPlan::build_dyn_layouts() {
    dyn_layout = false;
    const size_t nbt = batches.size();
    if (hcb || shard_world > 1 || nbt < 2 || nstr > 65535u * 65535u) return;
    std::vector<int> stat(n);  // orbital at static position p
    for (int o = 0; o < n; ++o) stat[bitpos[o]] = o;
    // order of a batch with active set S next to a batch with active set T: [S & T][S & ~T][spectators], each part in
    // static position order
    auto order_of = [&](uint32_t S, uint32_t T) {
        std::vector<int> o;
        for (int pass = 0; pass < 3; ++pass)
            for (int p = 0; p < n; ++p) {
                const int orb = stat[p];
                const bool in_s = S >> orb & 1, in_t = T >> orb & 1;
                if ((in_s ? (in_t ? 0 : 1) : 2) == pass) o.push_back(orb);
            }
        return o;
    };
    // forward: batch k reads L^F_k = order_of(S_k, S_{k-1}); writes L^F_{k+1}; the last batch writes the static order.
    // reverse: batch k reads L^R_k = order_of(S_k, S_{k+1}) (the last batch: static); writes L^R_{k-1}.
    std::vector<std::vector<int>> lf(nbt + 1), lr(nbt);
    for (size_t k = 0; k < nbt; ++k) {
        const uint32_t sk = batches[k].B.S;
        lf[k] = order_of(sk, k ? batches[k - 1].B.S : sk);
        lr[k] = k + 1 < nbt ? order_of(sk, batches[k + 1].B.S) : stat;
    }
    lf[nbt] = stat;
    auto tables = [&](const SideBatch& side, const std::vector<int>& order, ColTables* out) {
        out->addr_sorted.resize(side.addr.size());
        out->ord.resize(side.addr.size());
        std::vector<uint32_t> a;
        for (const PanelDesc& pd : side.panels) {
            const int cnt = int(pd.G) * int(pd.c);
            a.resize(cnt);
            for (int i = 0; i < cnt; ++i) a[i] = rank_in_order(order, strings[side.addr[pd.off + i]]);
            uint16_t* ord = out->ord.data() + pd.off;
            std::iota(ord, ord + cnt, uint16_t(0));
            std::sort(ord, ord + cnt, [&](uint16_t x, uint16_t y) { return a[x] < a[y]; });
            for (int i = 0; i < cnt; ++i) out->addr_sorted[pd.off + i] = a[ord[i]];
        }
    };
    for (size_t k = 0; k < nbt; ++k) {
        Batch& b = batches[k];
        if (b.B.addr.size() != size_t(nstr)) return;  // a beta side that does not list every column (never for whole-vector plans)
        tables(b.B, lf[k], &b.dyn_fwd_in);
        tables(b.B, lf[k + 1], &b.dyn_fwd_out);
        tables(b.B, lr[k], &b.dyn_rev_in);
        if (k > 0)
            tables(b.B, lr[k - 1], &b.dyn_rev_out);
        else
            b.dyn_rev_out = b.dyn_rev_in;  // nothing reads what the last batch of the reverse sweep writes
    }
    dyn_col0.resize(size_t(nstr));
    for (int64_t a = 0; a < nstr; ++a) dyn_col0[a] = rank_in_order(lf[0], strings[a]);
    dyn_layout = true;
}

void HostPlan::build_links() {
    int per = k * (n - k) + k;
    ml = (per + 7) / 8 * 8;
    if (ml == 0) ml = 8;
    links.assign(size_t(nstr) * ml, Link{0, 0, 0, 0});
    for (int64_t i = 0; i < nstr; ++i) {
        uint32_t s = strings[i];
        int c = 0;
        for (int q = 0; q < n; ++q) {
            if (!(s >> q & 1)) continue;
            for (int p = 0; p < n; ++p) {
                if (p == q) {
                    links[i * ml + c++] = Link{uint32_t(i), uint16_t(p * n + q), 1, 0};
                    continue;
                }
                if (s >> p & 1) continue;
                uint32_t t = s ^ (1u << p) ^ (1u << q);
                int lo = std::min(p, q), hi = std::max(p, q);
                uint32_t between = ((1u << hi) - 1) & ~((1u << (lo + 1)) - 1);
                int8_t sg = (popc(s & between) & 1) ? -1 : 1;
                links[i * ml + c++] = Link{uint32_t(rank(t)), uint16_t(p * n + q), sg, 0};
            }
        }
    }
}

void HostPlan::build_same_spin_ell(const std::vector<double>& h2e) {
    // H_ss[b', b] = sum_{pq,rs} h2e[pq,rs] <b'|E_pq E_rs|b>.  Walk from the OUTPUT string b':
    // <b'|E_pq E_rs|b> = <b|E_sr E_qp|b'>, i.e. first link (x,y)=(q,p) of b', then link (u,v)=(s,r) of the
    // intermediate; the coefficient is h2e[y,x,v,u].
    const int n2 = n * n;
    int per = k * (n - k) + k;
    int width = 1 + k * (n - k) + (k * (k - 1) / 2) * ((n - k) * (n - k - 1) / 2);
    ell_width = std::max(width, 1);
    ell_col.assign(size_t(nstr) * ell_width, 0);
    ell_val.assign(size_t(nstr) * ell_width, 0.0);
    unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (nstr < 2048) nthreads = 1;
    auto work = [&](int64_t lo, int64_t hi) {
        std::vector<double> acc(nstr, 0.0);
        std::vector<uint8_t> touched_flag(nstr, 0);
        std::vector<uint32_t> touched;
        for (int64_t i = lo; i < hi; ++i) {
            touched.clear();
            for (int l1 = 0; l1 < per; ++l1) {
                const Link& a = links[i * ml + l1];
                int x = a.pq / n, y = a.pq % n;
                const double* hrow = &h2e[size_t(y * n + x) * n2];
                for (int l2 = 0; l2 < per; ++l2) {
                    const Link& b = links[int64_t(a.addr) * ml + l2];
                    int u = b.pq / n, v = b.pq % n;
                    double val = hrow[v * n + u] * double(a.sgn * b.sgn);
                    if (!touched_flag[b.addr]) {
                        touched_flag[b.addr] = 1;
                        touched.push_back(b.addr);
                    }
                    acc[b.addr] += val;
                }
            }
            std::sort(touched.begin(), touched.end());
            int c = 0;
            for (uint32_t t : touched) {
                if (c < ell_width) {
                    ell_col[i * ell_width + c] = t;
                    ell_val[i * ell_width + c] = acc[t];
                    ++c;
                }
                acc[t] = 0.0;
                touched_flag[t] = 0;
            }
            for (; c < ell_width; ++c) {
                ell_col[i * ell_width + c] = uint32_t(i);
                ell_val[i * ell_width + c] = 0.0;
            }
        }
    };
    if (nthreads == 1) {
        work(0, nstr);
    } else {
        std::vector<std::thread> th;
        int64_t chunk = (nstr + nthreads - 1) / nthreads;
        for (unsigned t = 0; t < nthreads; ++t) {
            int64_t lo = t * chunk, hi = std::min<int64_t>(nstr, lo + chunk);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& t : th) t.join();
    }
}

std::vector<double> absorb_h1e(const double* int1e, const double* int2e, int n, int n_elec, double fac) {
    const size_t n2 = size_t(n) * n;
    std::vector<double> h2e(int2e, int2e + n2 * n2);
    std::vector<double> f1e(n2);
    for (int j = 0; j < n; ++j)
        for (int kk = 0; kk < n; ++kk) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += int2e[((size_t(j) * n + i) * n + i) * n + kk];
            f1e[j * n + kk] = (int1e[j * n + kk] - 0.5 * s) * (1.0 / (n_elec + 1e-100));
        }
    for (int kk = 0; kk < n; ++kk)
        for (int r = 0; r < n; ++r)
            for (int s = 0; s < n; ++s) {
                h2e[((size_t(kk) * n + kk) * n + r) * n + s] += f1e[r * n + s];
                h2e[((size_t(r) * n + s) * n + kk) * n + kk] += f1e[r * n + s];
            }
    for (auto& v : h2e) v *= fac;
    return h2e;
}

}  // namespace tcc
