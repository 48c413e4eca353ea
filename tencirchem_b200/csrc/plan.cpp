// Host-side plan construction.  See plan.hpp for what each piece replaces in the reference.
#include "plan.hpp"

#include <algorithm>
#include <cstring>
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
    strings.resize(nstr);
    if (k == 0) {
        strings[0] = 0;
    } else {
        uint64_t v = (uint64_t(1) << k) - 1;
        for (int64_t i = 0; i < nstr; ++i) {
            strings[i] = uint32_t(v);
            uint64_t t = v | (v - 1);  // Gosper: next integer with the same popcount
            v = (t + 1) | (((~t & (t + 1)) - 1) >> (__builtin_ctzll(v) + 1));
        }
    }
    na = hcb ? 1 : nstr;
    nb = nstr;
    N = na * nb;
    return 0;
}

int64_t HostPlan::rank(uint32_t s) const {
    int64_t r = 0;
    int i = 1;
    while (s) {
        int p = __builtin_ctz(s);
        r += binom[p][i];
        ++i;
        s &= s - 1;
    }
    return r;
}

int64_t HostPlan::rank_or_zero(uint64_t s) const {
    if (s >> n) return 0;
    if (popc(s) != k) return 0;
    return rank(uint32_t(s));
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

int HostPlan::make_op(const int32_t* idx, int len, int param, OpDesc* op) {
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
    op->len = len;
    op->param = param;
    for (int i = 0; i < 4; ++i) op->idx[i] = i < len ? idx[i] : 0;
    uint64_t cre = 0, ann = 0;
    for (int i = 0; i < len / 2; ++i) cre |= uint64_t(1) << idx[i];
    for (int i = len / 2; i < len; ++i) ann |= uint64_t(1) << idx[i];
    uint64_t lo = (uint64_t(1) << n) - 1;
    uint32_t cre_b = uint32_t(cre & lo), ann_b = uint32_t(ann & lo);
    uint32_t cre_a = hcb ? 0u : uint32_t(cre >> n), ann_a = hcb ? 0u : uint32_t(ann >> n);
    if (popc(cre_a) != popc(ann_a) || popc(cre_b) != popc(ann_b)) {
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
    op->hop_a = (cre_a | ann_a) ? find_or_build_hop(cre_a, ann_a) : -1;
    op->hop_b = (cre_b | ann_b) ? find_or_build_hop(cre_b, ann_b) : -1;
    op->kind = op->hop_a < 0 ? OP_BETA : (op->hop_b < 0 ? OP_ALPHA : OP_ALPHABETA);
    if (hcb && op->hop_a >= 0) {
        error = "internal: alpha hop in an hcb plan";
        return TCC_ERR_BAD_EXCITATION;
    }
    return 0;
}

int64_t HostPlan::support_pairs(const OpDesc& op) const {
    int64_t la = op.hop_a < 0 ? na : int64_t(hops[op.hop_a].src.size());
    int64_t lb = op.hop_b < 0 ? nb : int64_t(hops[op.hop_b].src.size());
    return la * lb;
}

void HostPlan::expand_tables(const OpDesc& op, uint64_t* perm, int8_t* phase, int8_t* f2) const {
    // partner address for EVERY determinant, including the reference's "invalid string -> 0" rule
    uint64_t mask = 0;
    for (int i = 0; i < op.len; ++i) mask |= uint64_t(1) << op.idx[i];
    uint64_t lo = (uint64_t(1) << n) - 1;
    uint32_t mb = uint32_t(mask & lo), ma = hcb ? 0u : uint32_t(mask >> n);
    std::vector<int64_t> pa(na), pb(nb);
    for (int64_t a = 0; a < na; ++a) pa[a] = hcb ? 0 : rank_or_zero(uint64_t(strings[a]) ^ ma);
    for (int64_t b = 0; b < nb; ++b) pb[b] = rank_or_zero(uint64_t(strings[b]) ^ mb);
    for (int64_t a = 0; a < na; ++a)
        for (int64_t b = 0; b < nb; ++b) {
            perm[a * nb + b] = uint64_t(pa[a] * nstr + pb[b]);
            phase[a * nb + b] = 0;
            f2[a * nb + b] = 0;
        }
    static const std::vector<uint32_t> empty;
    const HopTable* ha = op.hop_a >= 0 ? &hops[op.hop_a] : nullptr;
    const HopTable* hb = op.hop_b >= 0 ? &hops[op.hop_b] : nullptr;
    int64_t la = ha ? int64_t(ha->src.size()) : na;
    int64_t lb = hb ? int64_t(hb->src.size()) : nb;
    for (int64_t i = 0; i < la; ++i) {
        int64_t a = ha ? ha->src[i] : i;
        int64_t a2 = ha ? (ha->dst[i] & 0x7fffffffu) : i;
        int sa = ha ? int(ha->dst[i] >> 31) : 0;
        for (int64_t j = 0; j < lb; ++j) {
            int64_t b = hb ? hb->src[j] : j;
            int64_t b2 = hb ? (hb->dst[j] & 0x7fffffffu) : j;
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
