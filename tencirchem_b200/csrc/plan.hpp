// Host-side plan of the B200 civector engine: CI strings, per-spin hop tables, op descriptors,
// single-replacement link tables and the string-space Hamiltonian.  Plain C++ (no CUDA types)
// so the same code is used by the device engine and by the host-only introspection calls.
//
// What it replaces in the reference (paths relative to /root/reference):
//   tencirchem/static/ci_utils.py:16-49            CI strings and addresses
//   tencirchem/static/evolve_civector.py:24-106    partner index / support class / fermionic sign
// The reference stores [M, N] tables; everything factorises per spin string, so the
// plan keeps one compact list of (source address, partner address, sign) per DISTINCT spin hop.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace tcc {

struct HopTable {
    uint32_t cre = 0, ann = 0, zmask = 0;  // spin-local bit masks
    std::vector<uint32_t> src;             // strings with ann occupied & cre empty ("positive"), ascending address
    std::vector<uint32_t> dst;             // partner address | (parity(src_string & zmask) << 31)
};

enum OpKind : int { OP_BETA = 0, OP_ALPHA = 1, OP_ALPHABETA = 2 };

struct OpDesc {
    int kind = 0;
    int hop_a = -1, hop_b = -1;  // indices into HostPlan::hops (-1 = identity on that spin)
    int s0 = 1;                  // global sign of the JW string
    int param = 0;
    int len = 0;
    int idx[4] = {0, 0, 0, 0};
};

struct Link {       // E_pq |string> = sgn |addr>
    uint32_t addr;
    uint16_t pq;    // p * n + q
    int8_t sgn;     // +1 / -1, 0 for padding
    int8_t pad;
};

struct HostPlan {
    int n_qubits = 0, n_elec = 0, hcb = 0;
    int n = 0;         // bits per spin string (spatial orbitals)
    int k = 0;         // electrons (pairs for hcb) per string
    int64_t nstr = 0;  // strings per spin
    int64_t na = 0, nb = 0, N = 0;
    std::vector<uint32_t> strings;  // ascending integers with k bits among n
    std::vector<std::vector<int64_t>> binom;
    std::vector<HopTable> hops;
    std::map<uint64_t, int> hop_index;
    std::vector<OpDesc> ops;
    int n_params = 0;
    std::string error;

    int init(int n_qubits, int n_elec, int hcb);
    // address of a string with exactly k bits (combinatorial number system)
    int64_t rank(uint32_t s) const;
    // rank if popcount == k else 0 (the reference's zero-initialised lookup, ci_utils.py:27-28)
    int64_t rank_or_zero(uint64_t s) const;
    int find_or_build_hop(uint32_t cre, uint32_t ann);
    // returns 0 or a TCC_ERR_* code; fills `op`
    int make_op(const int32_t* idx, int len, int param, OpDesc* op);
    int64_t support_pairs(const OpDesc& op) const;
    // reference-format tables for one op (evolve_civector.py:91-106)
    void expand_tables(const OpDesc& op, uint64_t* perm, int8_t* phase, int8_t* f2) const;

    // sigma tables (non-hcb)
    int ml = 0;                // links per string padded to a multiple of 8
    std::vector<Link> links;   // [nstr, ml]
    void build_links();
    // string-space same-spin Hamiltonian sum h2e[pq,rs] E_pq E_rs in ELL form, rows = output string
    int ell_width = 0;
    std::vector<uint32_t> ell_col;  // [nstr, ell_width]
    std::vector<double> ell_val;
    void build_same_spin_ell(const std::vector<double>& h2e);
};

// pyscf direct_nosym.absorb_h1e with the factor 0.5 of hamiltonian.py:148
std::vector<double> absorb_h1e(const double* int1e, const double* int2e, int n, int n_elec, double fac);

}  // namespace tcc
