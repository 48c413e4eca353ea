// Host-side plan of the B200 civector engine: CI strings, per-spin hop tables, op descriptors, the
// batch schedule with its closed tiles, single-replacement link tables and the string-space
// Hamiltonian.  Plain C++ (no CUDA types) so the same code serves the device engine and the host-only
// introspection calls.
//
// What it replaces in the reference (paths relative to /root/reference):
//   tencirchem/static/ci_utils.py:16-49            CI strings and addresses
//   tencirchem/static/evolve_civector.py:24-106    partner index / support class / fermionic sign
// The reference stores [M, N] tables; everything factorises per spin string, so the
// plan keeps one compact list of (source address, partner address, sign) per DISTINCT spin hop.
//
// STORAGE ORDER.  Inside the engine the CI matrix is stored with the spin strings in a plan-specific
// order: the orbitals are renumbered by a bit permutation `bitpos` (most frequently active orbitals
// in the lowest positions) and strings are sorted as integers of the permuted bits.  That makes the
// closed tiles of a batch (below) nearly contiguous in memory.  `rank()` is the storage address,
// `ref_rank()` the reference address (ascending integers of the ORIGINAL bits, ci_utils.py:22);
// vectors crossing the C ABI are always in reference order.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace tcc {

struct HopTable {
    uint32_t cre = 0, ann = 0, zmask = 0;  // spin-local bit masks (original orbital numbering)
    std::vector<uint32_t> src;             // storage addresses of strings with ann occupied & cre empty, ascending
    std::vector<uint32_t> dst;             // partner storage address | (parity(src_string & zmask) << 31)
};

enum OpKind : int { OP_BETA = 0, OP_ALPHA = 1, OP_ALPHABETA = 2 };

struct OpDesc {
    int kind = 0;
    uint32_t cre_a = 0, ann_a = 0, cre_b = 0, ann_b = 0;
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

// ---- batch schedule --------------------------------------------------------------------------
// A batch is a run of excitations (in application order, after provably commuting swaps) whose
// orbitals lie inside the active sets (S_alpha, S_beta).  Strings split into classes by the
// occupation of the spectator orbitals; a class is closed under every hop of the batch, so a tile
// = (panel of alpha classes) x (panel of beta classes) can be rotated on-chip by the whole batch.
struct PanelDesc {
    uint32_t off;        // first entry of the panel in SideBatch::addr / ::ord
    uint32_t class_off;  // first class of the panel in SideBatch::sigma
    uint16_t G;          // classes in the panel
    uint16_t c;          // strings per class
    uint16_t e;          // active electrons per class
    uint16_t ld;         // leading dimension of a tile whose columns are this panel: (full G * c) | 1
    uint16_t Gn;         // nominal classes per panel of this electron count (G < Gn only in a last, partly filled panel)
    uint16_t pad;
};

struct LocalHop {
    uint32_t cre, ann, zmask;
};

struct SideBatch {
    uint32_t S = 0;  // active orbitals (original numbering)
    int m = 0;
    std::vector<uint32_t> addr;    // [nstr] storage addresses, panel-major, class-major, local order
    std::vector<uint16_t> ord;     // [nstr] per panel: local indices sorted by storage address
    std::vector<uint32_t> addr_sorted;  // [nstr] per panel: addr[ord[t]] (ascending inside a panel)
    std::vector<uint32_t> sigma;   // per class: spectator bits
    std::vector<PanelDesc> panels;
    std::vector<LocalHop> hops;
    std::vector<uint32_t> tab;         // entries src | dst << 15 | sign << 30 (class-local indices)
    std::vector<int32_t> tab_index;    // [n_hops][m + 1][2] = (offset, count)
    int max_panel = 0;                 // largest panel (strings)
    int max_classes = 0;               // largest number of classes in a panel
    std::vector<int> ld_of_e;          // [m + 1] tile leading dimension when this side supplies the columns
};

// Column tables of a batch under a per-batch column order (dynamic column layout, see HostPlan::build_dyn_layouts):
// per panel, like SideBatch::addr_sorted / ::ord, but with the beta strings addressed in another bit order
struct ColTables {
    std::vector<uint32_t> addr_sorted;  // [nstr] per panel: column addresses, ascending
    std::vector<uint16_t> ord;          // [nstr] per panel: the local index of each of them
};

struct BatchOp {
    int32_t op_index;
    int32_t h_a, h_b;  // local hop ids (-1 = identity)
    uint32_t zspec_a, zspec_b;
    int32_t s0;
};

struct Batch {
    SideBatch A, B;
    std::vector<BatchOp> ops;  // application order of the forward sweep
    // Fully expanded pair lists: for excitation t and a tile whose classes hold (e_a, e_b) active electrons,
    // entries i1 | i2 << 15 | sign << 30 with i = local_row * ld + local_col (ld = PanelDesc::ld of the beta panel),
    // relative to the (alpha class, beta class) sub-block of the tile.
    // A side without active orbitals is not expanded: the kernel repeats the list over its rows / columns.
    std::vector<uint32_t> plist;
    std::vector<int32_t> plist_index;  // [n_ops][m_a + 1][m_b + 1][2] = (offset, count)
    // Lean per-thread streams (batches with active orbitals on both sides): the pairs of excitation t in a tile of class
    // pair (e_a, e_b) laid out as `slots` rounds of `tps` entries, tps = threads / (Gn_a * Gn_b) = threads of one
    // (alpha class, beta class) sub-block; thread `lid` of the sub-block applies entry [slot * tps + lid] of every slot.
    // Entry = i1 | i2 << 15 | sign << 30 | valid << 31.  Two orderings of the same pairs: `stream8` groups 16
    // consecutive entries on distinct 8-byte banks (forward sweep and the two-vector reverse sweep), `stream16` groups
    // 8 consecutive entries on distinct 16-byte banks (reverse sweep on interleaved (ket, bra) elements).
    std::vector<uint32_t> stream8, stream16;
    std::vector<int32_t> stream_index;  // [n_ops][m_a + 1][m_b + 1][2] = (offset, slots)
    int lean_threads = 0;               // block size the streams were laid out for (0 = no streams)
    int ld_fixed = 0;                  // tile leading dimension when the beta side has no active orbitals
    uint32_t sa = 0, sb = 0;           // padded active sets
    int layout = 0;                    // row distribution the batch runs in (sharded plans)
    // dynamic column layout: the tables the forward / reverse sweep gathers from and scatters to (empty = static order)
    ColTables dyn_fwd_in, dyn_fwd_out, dyn_rev_in, dyn_rev_out;
};

// Row distribution of a sharded plan: alpha string a lives on rank owner[a] (the occupation pattern of the
// orbitals T, none of which is active on the alpha side of the batches that use the layout, so every alpha class
// lies on one rank) at local row local_row[a] (ascending storage address within the rank).
struct RowLayout {
    uint32_t T = 0;
    std::vector<uint8_t> owner;       // [nstr]
    std::vector<uint32_t> local_row;  // [nstr]
    std::vector<int64_t> rows;        // [world] rows per rank
};

struct Segment {  // run of consecutive batches with the same layout
    int layout, batch_lo, batch_hi;
};

struct HostPlan {
    int n_qubits = 0, n_elec = 0, hcb = 0;
    int n = 0;         // bits per spin string (spatial orbitals)
    int k = 0;         // electrons (pairs for hcb) per string
    int64_t nstr = 0;  // strings per spin
    int64_t na = 0, nb = 0, N = 0;
    std::vector<int> bitpos;        // orbital -> storage bit position
    std::vector<uint32_t> strings;  // storage order; values are ORIGINAL occupation bits
    std::vector<uint32_t> ref2int;  // reference address -> storage address
    std::vector<uint32_t> int2ref;
    bool identity_layout = true;
    std::vector<std::vector<int64_t>> binom;
    std::vector<HopTable> hops;
    std::map<uint64_t, int> hop_index;
    std::vector<OpDesc> ops;
    std::vector<Batch> batches;
    // alpha-row sharding (world == 1: one layout that owns everything)
    int shard_rank = 0, shard_world = 1;
    std::vector<RowLayout> layouts;
    std::vector<Segment> segments;
    // sharded sigma: three layouts with pairwise disjoint T; an alpha replacement E_pq is contracted in the first of
    // them whose T it does not touch (there it maps a row to a row of the same rank)
    int sigma_layout[3] = {0, 0, 0};
    int sigma_aml[3] = {0, 0, 0};
    std::vector<Link> sigma_links[3];  // [local rows of the pass layout][sigma_aml], addr = LOCAL target row
    void build_sigma_pass_links();
    int64_t max_local_rows = 0;  // over the layouts, for this rank
    int shard_bits() const {
        int t = 0;
        while ((1 << t) < shard_world) ++t;
        return t;
    }
    int n_params = 0;
    int sched_mode = 0;  // batch scheduler (group_ops): 0 = active sets grown first come first served (default), 1 = chosen
                         // for the most excitations, 2 = whichever gives fewer batches
    std::string error;
    // DYNAMIC COLUMN LAYOUT (whole-vector plans).  In the static storage order a beta class of a batch is contiguous
    // only when the active beta orbitals of the batch happen to sit in the lowest bit positions; over a UCCSD sweep the
    // mean run of consecutive columns in a tile is below 3 elements, and the tile gather / scatter is bound by the
    // number of distinct 128-byte lines per warp instruction, not by bytes (profiles/r02_reverse_kernel_breakdown.md).
    // With the dynamic layout every batch k READS the CI matrix with its columns in an order L_k whose lowest bit
    // positions are the active beta orbitals of the batch (every beta class is one contiguous block of columns) and
    // WRITES it in the order of the batch that follows in the sweep (out of place, two buffers); the orbitals shared
    // with that batch come first in its order, so a class is written in runs of C(|shared|, .) columns.  The forward
    // sweep starts from L^F_0 (dyn_col0: static column address -> address in L^F_0) and ends in the static order; the
    // reverse sweep starts from the static order.  Rows (alpha strings) keep the static order.
    bool dyn_layout = false;
    std::vector<uint32_t> dyn_col0;  // [nstr] static column address -> column address in the first forward layout
    void build_dyn_layouts();
    uint32_t rank_in_order(const std::vector<int>& order, uint32_t s) const;  // order[pos] = orbital at bit position pos

    int init(int n_qubits, int n_elec, int hcb);
    // validation + masks + sign; no tables (strings may not exist yet)
    int parse_op(const int32_t* idx, int len, int param, OpDesc* op);
    // groups ops into batches (orbital sets only), picks the storage order, builds strings, hop tables
    // and batch descriptors.  tile_elems = max determinants per on-chip tile.
    void finalize(int tile_elems, bool use_layout, int rank = 0, int world = 1, int lean_threads = 256, bool use_dyn = false);
    void attach_hops(OpDesc* op);

    uint32_t permute_bits(uint32_t s) const;
    int64_t colex_rank(uint32_t t) const;                  // rank of a k-subset given as bits
    int64_t rank(uint32_t s) const { return colex_rank(permute_bits(s)); }  // storage address
    int64_t ref_rank(uint32_t s) const { return colex_rank(s); }
    // reference address if popcount == k else 0 (the reference's zero-initialised lookup, ci_utils.py:27-28)
    int64_t ref_rank_or_zero(uint64_t s) const;
    int find_or_build_hop(uint32_t cre, uint32_t ann);
    int64_t support_pairs(const OpDesc& op) const;
    // reference-format tables for one op (evolve_civector.py:91-106), reference order
    void expand_tables(const OpDesc& op, uint64_t* perm, int8_t* phase, int8_t* f2) const;

    // sigma tables (non-hcb), storage order
    int ml = 0;                // links per string padded to a multiple of 8
    std::vector<Link> links;   // [nstr, ml]
    void build_links();
    int ell_width = 0;
    std::vector<uint32_t> ell_col;  // [nstr, ell_width]
    std::vector<double> ell_val;
    void build_same_spin_ell(const std::vector<double>& h2e);

   private:
    std::vector<std::vector<int>> group_ops(int tile_elems) const;
    void choose_layout(const std::vector<std::vector<int>>& groups, bool use_layout);
    void build_strings();
    void active_sets(const std::vector<int>& group, int tile_elems, uint32_t* sa, uint32_t* sb) const;
    void choose_row_layouts();
    void build_batch(const std::vector<int>& group, int tile_elems, Batch* out, int lean_threads);
    void build_side(uint32_t S, int panel_target, SideBatch* side, bool single_string, bool pack_classes,
                    const RowLayout* rows);
    int64_t max_class(int m) const;
};

// pyscf direct_nosym.absorb_h1e with the factor 0.5 of hamiltonian.py:148
std::vector<double> absorb_h1e(const double* int1e, const double* int2e, int n, int n_elec, double fac);

}  // namespace tcc
