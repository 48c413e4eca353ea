// C ABI of the B200 civector engine (include/tcc_b200.h): device plan, drivers for get_civector /
// energy_and_grad / apply_excitation / H.c, host-buffer conveniences.
//
// Reference call sites this replaces (paths relative to /root/reference):
//   tencirchem/static/evolve_civector.py:180-198  get_civector                  -> tcc_civector
//   tencirchem/static/evolve_civector.py:201-243  get_energy_and_grad_civector  -> tcc_energy_and_grad
//   tencirchem/static/evolve_civector.py:311-313  apply_excitation_civector     -> tcc_apply_excitation
//   tencirchem/static/hamiltonian.py:146-183      fci_func (H.c)                -> tcc_set_hamiltonian / tcc_apply_h
#include <cuda_runtime.h>

#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/tcc_b200.h"
#include "batch_kernels.cuh"
#include "kernels.cuh"
#include "plan.hpp"

using namespace tcc;

static thread_local std::string g_error;

static int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(TCC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
    } while (0)

struct DevHop {
    uint32_t* src = nullptr;
    uint32_t* dst = nullptr;
    int len = 0;
};

// Per-phase device timing: CUDA events are recorded only where the phase (slot) changes, so a run of
// thousands of identical launches costs two events.  Slots are listed in include/tcc_b200.h.
struct Profiler {
    bool on = false;
    int cur = -1;
    std::vector<cudaEvent_t> pool;
    std::vector<std::pair<int, int>> marks;  // (slot, event index)
    size_t used = 0;
    double ms[TCC_PROF_SLOTS] = {0};
    int64_t launches[TCC_PROF_SLOTS] = {0};
    int64_t bytes[TCC_PROF_SLOTS] = {0};
    cudaEvent_t next_event() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    void mark(int slot, cudaStream_t st) {
        if (!on || slot == cur) return;
        cudaEvent_t e = next_event();
        cudaEventRecord(e, st);
        marks.push_back({slot, int(used) - 1});
        cur = slot;
    }
    void count(int slot, int64_t nlaunch, int64_t nbytes) {
        if (!on) return;
        launches[slot] += nlaunch;
        bytes[slot] += nbytes;
    }
    // call after the stream has been synchronised
    void resolve() {
        if (!on) return;
        for (size_t i = 0; i + 1 < marks.size(); ++i) {
            float t = 0.f;
            if (marks[i].first >= 0 && cudaEventElapsedTime(&t, pool[marks[i].second], pool[marks[i + 1].second]) == cudaSuccess)
                ms[marks[i].first] += t;
        }
        marks.clear();
        used = 0;
        cur = -1;
    }
    void reset() {
        for (int i = 0; i < TCC_PROF_SLOTS; ++i) ms[i] = 0, launches[i] = 0, bytes[i] = 0;
        marks.clear();
        used = 0;
        cur = -1;
    }
    ~Profiler() {
        for (auto e : pool) cudaEventDestroy(e);
    }
};

struct BatchDevSet {
    BatchDev fwd, rev;
    size_t smem_fwd = 0, smem_rev = 0;
    int n_tiles = 0;
    int nops = 0;
    int max_rows = 0, max_cols = 0;
    std::vector<TileRect> rects;      // forward launches
    std::vector<TileRect> rects_rev;  // reverse launches
    std::vector<void*> allocs;
};

struct BatchTimer {
    bool on = false;
    std::vector<cudaEvent_t> ev;  // fwd: [0 .. nb], rev: [nb+1 .. 2nb+1]
    std::vector<float> fwd_ms, rev_ms;
    void ensure(size_t nb) {
        while (ev.size() < 2 * nb + 2) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev.push_back(e);
        }
        fwd_ms.assign(nb, 0.f);
        rev_ms.assign(nb, 0.f);
    }
    ~BatchTimer() {
        for (auto e : ev) cudaEventDestroy(e);
    }
};

struct tcc_plan {
    HostPlan hp;
    Profiler prof;
    BatchTimer btimer;
    int device = 0;
    int mode = 0;  // 0 = batched tiles, 1 = one launch per excitation
    int batch_threads = 256;      // forward batch kernel block size
    int batch_threads_rev = 256;  // reverse batch kernel block size
    std::vector<DevHop> dhops;
    std::vector<BatchDevSet> dbatches;
    double* ket = nullptr;
    double* bra = nullptr;
    double* ws1 = nullptr;
    double* ws2 = nullptr;
    double* kb = nullptr;   // interleaved (ket, bra) elements of the reverse sweep, [N][2]
    int lean = 1;           // lean per-thread streams for two-sided batches (TCC_B200_LEAN=0: the plain pair lists)
    int lean_fwd = 0;       // forward sweep on the lean kernel too (TCC_B200_LEAN_FWD=1); measured slower than batch_kernel<1>
    int rev_interleave = 1; // reverse sweep on interleaved (ket, bra) elements (TCC_B200_REV_IL=0: two vectors)
    int persist_rev = 1;    // reverse sweep on the persistent kernel (lean_persist_kernel<3>; TCC_B200_PERSIST=0: one tile per CTA,
                            // =2: also for launches with few tiles -- the default uses it from 4 tiles per resident CTA on)
    int persist_min_waves = 4;
    int persist_fwd = 0;    // forward sweep on lean_persist_kernel<1> (TCC_B200_PERSIST_FWD=1; needs the lean forward streams)
    unsigned* persist_ctl = nullptr;  // ticket counters of the persistent launches [2][16][2], self-resetting
    int sms = 148;
    int dyn = 0;            // dynamic column layout (plan.hpp, TCC_B200_DYN_LAYOUT=1; off by default: measured no faster and it takes
                            // three more vectors): out-of-place sweeps, every batch gathers contiguous columns
    double* dyn_f = nullptr;   // second buffer of the forward sweep [N]
    double* kb2 = nullptr;     // second buffer of the interleaved reverse sweep [2 N]
    uint32_t* d_dyn_init = nullptr;  // [nstr] column map reference order -> first forward layout (init_state)
    int64_t hf_index_dyn = 0;
    double* io1 = nullptr;  // reference-order staging for the host-buffer entry points
    double* io2 = nullptr;
    double* partials = nullptr;
    unsigned* ticket = nullptr;
    double* grad_ops = nullptr;  // [n_ops + 8] on the device
    double* bpart = nullptr;     // per-tile gradient partials of the batched reverse sweep
    double* pinned = nullptr;    // pinned host staging [n_ops + 8]
    double2* d_rot_fwd = nullptr;  // [n_ops] (cos, sin * s0)
    double2* d_rot_rev = nullptr;  // [n_ops] (cos, -sin * s0)
    double2* h_rot = nullptr;      // pinned [2 * n_ops]
    cudaEvent_t rot_event = nullptr;
    StreamFan fan;  // concurrent rectangle launches of a batch
    uint32_t* d_ref2int = nullptr;
    uint32_t* d_int2ref = nullptr;
    int64_t hf_index = 0;
    bool has_h = false;
    Link* d_links = nullptr;
    double* d_w = nullptr;
    double *rows_ws1 = nullptr, *rows_ws2 = nullptr;  // scratch of tcc_apply_h_rows
    Link* d_sigma_links[3] = {nullptr, nullptr, nullptr};  // sharded sigma: alpha links per pass (local rows)
    size_t rows_ws_cap = 0;
    uint16_t* d_kmap = nullptr;  // beta link (p,q) -> row of the D intermediate
    bool w_pair_symmetric = false;
    int n2p = 0;
    uint32_t* d_ell_col = nullptr;
    double* d_ell_val = nullptr;
    uint32_t* d_strings = nullptr;
    double* d_diag1 = nullptr;
    double* d_hop = nullptr;
    double* d_diag2 = nullptr;
    int64_t* d_binom = nullptr;
    int* d_bitpos = nullptr;
    int64_t launches = 0;
    int64_t dev_bytes = 0;
};

template <typename T>
static cudaError_t dev_alloc(tcc_plan* p, T** ptr, size_t count) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(ptr), count * sizeof(T) + 16);
    if (e == cudaSuccess) p->dev_bytes += int64_t(count * sizeof(T));
    return e;
}

template <typename T>
static cudaError_t dev_upload(tcc_plan* p, const std::vector<T>& h, const T** out, std::vector<void*>* owner) {
    T* d = nullptr;
    cudaError_t e = dev_alloc(p, &d, h.size() ? h.size() : 1);
    if (e != cudaSuccess) return e;
    owner->push_back(d);
    if (!h.empty()) e = cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    *out = d;
    return e;
}

static int upload_hop(tcc_plan* p, int id) {
    if (int(p->dhops.size()) <= id) p->dhops.resize(id + 1);
    DevHop& d = p->dhops[id];
    if (d.src) return 0;
    const HopTable& h = p->hp.hops[id];
    d.len = int(h.src.size());
    size_t cnt = h.src.size() ? h.src.size() : 1;
    CUDA_TRY(dev_alloc(p, &d.src, cnt));
    CUDA_TRY(dev_alloc(p, &d.dst, cnt));
    if (d.len) {
        CUDA_TRY(cudaMemcpy(d.src, h.src.data(), d.len * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d.dst, h.dst.data(), d.len * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    return 0;
}

static int upload_side(tcc_plan* p, const SideBatch& s, SideDev* d, std::vector<void*>* owner) {
    CUDA_TRY(dev_upload(p, s.addr, &d->addr, owner));
    CUDA_TRY(dev_upload(p, s.ord, &d->ord, owner));
    CUDA_TRY(dev_upload(p, s.addr_sorted, &d->addr_sorted, owner));
    CUDA_TRY(dev_upload(p, s.sigma, &d->sigma, owner));
    CUDA_TRY(dev_upload(p, s.panels, &d->panels, owner));
    CUDA_TRY(dev_upload(p, s.tab, &d->tab, owner));
    CUDA_TRY(dev_upload(p, s.tab_index, &d->tab_index, owner));
    d->m = s.m;
    d->n_panels = int(s.panels.size());
    return 0;
}

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

static int upload_batches(tcc_plan* p) {
    HostPlan& hp = p->hp;
    p->dbatches.resize(hp.batches.size());
    size_t max_part = 1;
    CUDA_TRY(batch_kernels_configure());  // before the occupancy queries of the persistent kernel below
    for (size_t b = 0; b < hp.batches.size(); ++b) {
        const Batch& hb = hp.batches[b];
        BatchDevSet& db = p->dbatches[b];
        if (int rc = upload_side(p, hb.A, &db.fwd.A, &db.allocs)) return rc;
        if (int rc = upload_side(p, hb.B, &db.fwd.B, &db.allocs)) return rc;
        CUDA_TRY(dev_upload(p, hb.plist, &db.fwd.plist, &db.allocs));
        CUDA_TRY(dev_upload(p, hb.plist_index, &db.fwd.plist_index, &db.allocs));
        db.fwd.ld_fixed = hb.ld_fixed;
        db.fwd.reversed = 0;
        db.fwd.stream8 = db.fwd.stream16 = nullptr;
        db.fwd.stream_index = nullptr;
        db.fwd.lean = 0;
        if (p->lean && hb.lean_threads > 0 && hb.lean_threads == p->batch_threads && hb.lean_threads == p->batch_threads_rev) {
            CUDA_TRY(dev_upload(p, hb.stream8, &db.fwd.stream8, &db.allocs));
            CUDA_TRY(dev_upload(p, hb.stream16, &db.fwd.stream16, &db.allocs));
            CUDA_TRY(dev_upload(p, hb.stream_index, &db.fwd.stream_index, &db.allocs));
            db.fwd.lean = 1;
        }
        db.fwd.B.dyn_in_sorted = db.fwd.B.dyn_out_sorted = nullptr;
        db.fwd.B.dyn_in_ord = db.fwd.B.dyn_out_ord = nullptr;
        db.fwd.A.dyn_in_sorted = db.fwd.A.dyn_out_sorted = nullptr;
        db.fwd.A.dyn_in_ord = db.fwd.A.dyn_out_ord = nullptr;
        db.rev = db.fwd;
        if (hp.dyn_layout) {
            CUDA_TRY(dev_upload(p, hb.dyn_fwd_in.addr_sorted, &db.fwd.B.dyn_in_sorted, &db.allocs));
            CUDA_TRY(dev_upload(p, hb.dyn_fwd_in.ord, &db.fwd.B.dyn_in_ord, &db.allocs));
            CUDA_TRY(dev_upload(p, hb.dyn_fwd_out.addr_sorted, &db.fwd.B.dyn_out_sorted, &db.allocs));
            CUDA_TRY(dev_upload(p, hb.dyn_fwd_out.ord, &db.fwd.B.dyn_out_ord, &db.allocs));
            CUDA_TRY(dev_upload(p, hb.dyn_rev_in.addr_sorted, &db.rev.B.dyn_in_sorted, &db.allocs));
            CUDA_TRY(dev_upload(p, hb.dyn_rev_in.ord, &db.rev.B.dyn_in_ord, &db.allocs));
            CUDA_TRY(dev_upload(p, hb.dyn_rev_out.addr_sorted, &db.rev.B.dyn_out_sorted, &db.allocs));
            CUDA_TRY(dev_upload(p, hb.dyn_rev_out.ord, &db.rev.B.dyn_out_ord, &db.allocs));
        }
        db.fwd.lean = db.fwd.lean && p->lean_fwd;
        db.rev.reversed = 1;
        std::vector<BatchOp> rev_ops(hb.ops.rbegin(), hb.ops.rend());
        CUDA_TRY(dev_upload(p, hb.ops, &db.fwd.ops, &db.allocs));
        CUDA_TRY(dev_upload(p, rev_ops, &db.rev.ops, &db.allocs));
        db.nops = db.fwd.nops = db.rev.nops = int(hb.ops.size());
        db.n_tiles = db.fwd.A.n_panels * db.fwd.B.n_panels;
        db.max_rows = hb.A.max_panel;
        db.max_cols = hb.B.max_panel;
        // split the panel grid into rectangles of similar tile size (panels are sorted by class size, largest first):
        // each rectangle is one launch whose shared-memory request, hence occupancy, fits its tiles
        auto side_ranges = [](const SideBatch& sb) {
            std::vector<std::array<int, 3>> out;  // lo, count, max panel size
            const int np_ = int(sb.panels.size());
            // nominal size of a panel: the full panel of its electron count (a last, partly filled panel counts as full)
            auto nominal = [&](int i) {
                return sb.m > 0 ? int(sb.panels[i].ld) : int(sb.panels[i].G) * int(sb.panels[i].c);
            };
            int lo = 0;
            while (lo < np_) {
                int big = nominal(lo);
                int hi = lo + 1;
                // a range ends when the panels get smaller than 85 % of its first (largest) one; at most 3 ranges
                while (hi < np_ && (int(out.size()) == 2 || 100 * nominal(hi) >= 85 * big)) {
                    big = std::max(big, nominal(hi));
                    ++hi;
                }
                out.push_back({lo, hi - lo, big});
                lo = hi;
            }
            return out;
        };
        // 0 (default): one launch per batch; 1: the reverse kernel, whose occupancy is bound by shared memory (ket and
        // bra tile), gets one launch per rectangle; 2: the forward kernel too.  Measured at (16e,16o): the per-excitation
        // cost of the reverse kernel drops by 25 % but the per-batch cost grows more (tiles that share DRAM sectors no
        // longer run together), see profiles/r01_batch_kernel_experiments.md.  3: reverse kernel, split along alpha only
        const int split = env_int("TCC_B200_TILE_RECTS", 0);
        const std::vector<std::array<int, 3>> one_a = {{0, int(hb.A.panels.size()), hb.A.max_panel}};
        const std::vector<std::array<int, 3>> one_b = {{0, int(hb.B.panels.size()), hb.B.max_panel}};
        db.smem_fwd = db.smem_rev = 0;
        auto make_rects = [&](bool do_split, std::vector<TileRect>* out) {
            const std::vector<std::array<int, 3>> ra = do_split ? side_ranges(hb.A) : one_a;
            const std::vector<std::array<int, 3>> rb = do_split && split != 3 ? side_ranges(hb.B) : one_b;
            out->clear();
            for (auto& xa : ra)
                for (auto& xb : rb) {
                    TileRect r;
                    r.pa_lo = xa[0];
                    r.pa_cnt = xa[1];
                    r.pb_lo = xb[0];
                    r.pb_cnt = xb[1];
                    // ld of a tile follows its beta panel except when the beta side has no active orbitals (fixed ld)
                    const int cols = hb.B.m > 0 ? xb[2] : hb.B.max_panel;
                    r.smem_fwd = db.fwd.lean ? lean_smem_bytes(xa[2], cols, db.nops, 1) : batch_smem_bytes(xa[2], cols, db.nops, 1);
                    r.smem_rev = db.rev.lean ? lean_smem_bytes(xa[2], cols, db.nops, 2) : batch_smem_bytes(xa[2], cols, db.nops, 2);
                    r.smem_il = db.rev.lean ? lean_smem_bytes(xa[2], cols, db.nops, 3) : 0;
                    r.max_rows = xa[2];
                    r.max_cols = cols;
                    if (db.fwd.lean && p->persist_fwd) persist_config(1, p->batch_threads, xa[2], cols, db.nops, &r.smem_p_fwd, &r.ctas_p_fwd);
                    if (db.rev.lean && p->persist_rev && p->rev_interleave)
                        persist_config(3, p->batch_threads_rev, xa[2], cols, db.nops, &r.smem_p_rev, &r.ctas_p_rev);
                    db.smem_fwd = std::max(db.smem_fwd, r.smem_fwd);
                    db.smem_rev = std::max(db.smem_rev, r.smem_rev);
                    if (r.pa_cnt > 0 && r.pb_cnt > 0) out->push_back(r);
                }
        };
        make_rects(split >= 2, &db.rects);
        make_rects(split >= 1, &db.rects_rev);
        for (const TileRect& r : db.rects_rev) db.smem_rev = std::max(db.smem_rev, r.smem_il);
        if (db.smem_rev > 232448)
            return fail(TCC_ERR_BAD_ARGUMENT, "batch tile does not fit in shared memory: lower TCC_B200_TILE_ELEMS");
        max_part = std::max(max_part, size_t(db.nops) * size_t(db.n_tiles));
    }
    CUDA_TRY(dev_alloc(p, &p->bpart, max_part));
    CUDA_TRY(batch_kernels_configure());
    return 0;
}

static SideList row_side(const tcc_plan* p, const OpDesc& op) {
    if (op.hop_a < 0) return SideList{nullptr, nullptr, int(p->hp.na)};
    const DevHop& d = p->dhops[op.hop_a];
    return SideList{d.src, d.dst, d.len};
}
static SideList col_side(const tcc_plan* p, const OpDesc& op) {
    if (op.hop_b < 0) return SideList{nullptr, nullptr, int(p->hp.nb)};
    const DevHop& d = p->dhops[op.hop_b];
    return SideList{d.src, d.dst, d.len};
}

static int ensure_vec(tcc_plan* p, double** v) {
    if (*v) return 0;
    CUDA_TRY(dev_alloc(p, v, size_t(p->hp.N)));
    return 0;
}

static int use_device(const tcc_plan* p) {
    if (p->device < 0)
        return fail(TCC_ERR_NO_DEVICE, "host-only plan (device = -1) supports introspection only: there is no CPU fallback");
    CUDA_TRY(cudaSetDevice(p->device));
    return 0;
}

// the whole-vector entry points of a plan that holds only a block of alpha rows make no sense
static int whole_vector_only(const tcc_plan* p) {
    if (p->hp.shard_world > 1)
        return fail(TCC_ERR_BAD_ARGUMENT, "sharded plan: use the tcc_shard_* entry points (the plan holds 1/world of the rows)");
    return 0;
}


// reference order -> storage order and back (vectors crossing the ABI are in reference order)
static void to_storage(tcc_plan* p, const double* ref, double* st_vec, cudaStream_t st) {
    const HostPlan& hp = p->hp;
    if (hp.identity_layout) {
        if (ref != st_vec) cudaMemcpyAsync(st_vec, ref, hp.N * sizeof(double), cudaMemcpyDeviceToDevice, st);
        return;
    }
    launch_permute(ref, st_vec, hp.na, hp.nb, hp.hcb ? nullptr : p->d_int2ref, p->d_int2ref, st);
    p->launches += 1;
}
static void from_storage(tcc_plan* p, const double* st_vec, double* ref, cudaStream_t st) {
    const HostPlan& hp = p->hp;
    if (hp.identity_layout) {
        if (ref != st_vec) cudaMemcpyAsync(ref, st_vec, hp.N * sizeof(double), cudaMemcpyDeviceToDevice, st);
        return;
    }
    launch_permute(st_vec, ref, hp.na, hp.nb, hp.hcb ? nullptr : p->d_ref2int, p->d_ref2int, st);
    p->launches += 1;
}

extern "C" {

const char* tcc_last_error(void) { return g_error.c_str(); }
const char* tcc_version(void) { return "tencirchem_b200 0.2 (sm_100a)"; }

int tcc_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(TCC_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return n;
}

static int plan_create_impl(int n_qubits, int n_elec, int hcb, int n_ops, const int32_t* ex_ops_flat, const int32_t* ex_op_len,
                            const int32_t* param_ids, int device, int rank, int world, tcc_plan** out) {
    if (!out) return fail(TCC_ERR_BAD_ARGUMENT, "out is null");
    *out = nullptr;
    if (world < 1 || (world & (world - 1)) || world > 64 || rank < 0 || rank >= world)
        return fail(TCC_ERR_BAD_ARGUMENT, "world size must be a power of two (<= 64) and 0 <= rank < world");
    if (world > 1 && hcb) return fail(TCC_ERR_BAD_ARGUMENT, "hcb plans have a single alpha row and are not sharded");
    if (n_ops < 0 || (n_ops > 0 && (!ex_ops_flat || !ex_op_len))) return fail(TCC_ERR_BAD_ARGUMENT, "bad excitation list");
    tcc_plan* p = new tcc_plan();
    int rc = p->hp.init(n_qubits, n_elec, hcb);
    if (rc) {
        g_error = p->hp.error;
        delete p;
        return rc;
    }
    int off = 0, max_param = -1;
    p->hp.ops.resize(n_ops);
    for (int j = 0; j < n_ops; ++j) {
        int pid = param_ids ? param_ids[j] : j;  // engine_ucc.py:43-44
        if (pid < 0) {
            delete p;
            return fail(TCC_ERR_BAD_ARGUMENT, "negative parameter id");
        }
        int len = ex_op_len[j];
        if (len != 2 && len != 4) {
            delete p;
            return fail(TCC_ERR_BAD_EXCITATION, "Excitation of length " + std::to_string(len) + " not supported");
        }
        rc = p->hp.parse_op(ex_ops_flat + off, len, pid, &p->hp.ops[j]);
        if (rc) {
            g_error = p->hp.error;
            delete p;
            return rc;
        }
        off += len;
        if (pid > max_param) max_param = pid;
    }
    p->hp.n_params = max_param + 1;
    p->mode = env_int("TCC_B200_PER_OP", 0) ? 1 : 0;
    p->batch_threads = std::min(256, std::max(32, env_int("TCC_B200_BATCH_THREADS", 256) / 32 * 32));
    p->batch_threads_rev = std::min(256, std::max(32, env_int("TCC_B200_BATCH_THREADS_REV", p->batch_threads) / 32 * 32));
    if (world > 1) p->mode = 0;  // sharded plans run the batched engine only
    p->lean = env_int("TCC_B200_LEAN", 1) != 0 && p->batch_threads == p->batch_threads_rev && (p->batch_threads & (p->batch_threads - 1)) == 0 &&
              p->batch_threads >= 64;
    p->rev_interleave = p->lean && env_int("TCC_B200_REV_IL", 1) != 0;
    p->persist_rev = env_int("TCC_B200_PERSIST", 1) != 0;
    p->persist_min_waves = env_int("TCC_B200_PERSIST", 1) == 2 ? 0 : 4;
    p->persist_fwd = p->lean && env_int("TCC_B200_PERSIST_FWD", 0) != 0;
    p->lean_fwd = p->lean && (env_int("TCC_B200_LEAN_FWD", 0) != 0 || p->persist_fwd);
    p->hp.error.clear();
    p->hp.sched_mode = env_int("TCC_B200_SCHED", 0);
    p->hp.finalize(env_int("TCC_B200_TILE_ELEMS", 6144), env_int("TCC_B200_LAYOUT", 1) != 0, rank, world, p->lean ? p->batch_threads : 0,
                   env_int("TCC_B200_DYN_LAYOUT", 0) != 0 && world == 1);
    if (!p->hp.error.empty()) {
        g_error = p->hp.error;
        delete p;
        return TCC_ERR_BAD_ARGUMENT;
    }
    {
        uint32_t hf = p->hp.k ? ((1u << p->hp.k) - 1) : 0u;
        int64_t r = p->hp.rank(hf);
        p->hf_index = (p->hp.hcb ? 0 : r * p->hp.nb) + r;
    }
    if (device == -1) {  // host-only plan: strings / tables introspection, no compute
        p->device = -1;
        *out = p;
        return 0;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        delete p;
        return fail(TCC_ERR_NO_DEVICE, "no CUDA device available: the B200 engine has no CPU fallback");
    }
    if (device < 0 || device >= ndev) {
        delete p;
        return fail(TCC_ERR_BAD_ARGUMENT, "device ordinal out of range");
    }
    p->device = device;
    auto bail = [&](int code) {
        tcc_plan_destroy(p);
        return code;
    };
    if (int rc_dev = use_device(p)) return bail(rc_dev);
    cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaMalloc(&p->persist_ctl, sizeof(unsigned) * 64) != cudaSuccess) return bail(TCC_ERR_CUDA);
    cudaMemset(p->persist_ctl, 0, sizeof(unsigned) * 64);
    for (size_t h = 0; h < p->hp.hops.size(); ++h)
        if (upload_hop(p, int(h))) return bail(TCC_ERR_CUDA);
    if (int rc2 = upload_batches(p)) return bail(rc2);
    if (cudaMalloc(&p->partials, sizeof(double) * reverse_partials_capacity()) != cudaSuccess ||
        cudaMalloc(&p->ticket, sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc(&p->grad_ops, sizeof(double) * (n_ops + 8)) != cudaSuccess ||
        cudaMalloc(&p->d_rot_fwd, sizeof(double2) * (n_ops + 1)) != cudaSuccess ||
        cudaMalloc(&p->d_rot_rev, sizeof(double2) * (n_ops + 1)) != cudaSuccess ||
        cudaMalloc(&p->d_ref2int, sizeof(uint32_t) * p->hp.nstr) != cudaSuccess ||
        cudaMalloc(&p->d_int2ref, sizeof(uint32_t) * p->hp.nstr) != cudaSuccess ||
        cudaMallocHost(&p->h_rot, sizeof(double2) * (2 * n_ops + 2)) != cudaSuccess ||
        cudaMallocHost(&p->pinned, sizeof(double) * (n_ops + 8)) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->rot_event, cudaEventDisableTiming) != cudaSuccess ||
        p->fan.create(env_int("TCC_B200_FAN_STREAMS", 8)) != cudaSuccess) {
        g_error = std::string("plan workspace allocation: ") + cudaGetErrorString(cudaGetLastError());
        return bail(TCC_ERR_CUDA);
    }
    cudaMemcpy(p->d_ref2int, p->hp.ref2int.data(), sizeof(uint32_t) * p->hp.nstr, cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_int2ref, p->hp.int2ref.data(), sizeof(uint32_t) * p->hp.nstr, cudaMemcpyHostToDevice);
    if (p->hp.dyn_layout) {
        // column c' of the first forward layout holds the string with static address inv[c'], reference address int2ref[inv[c']]
        const HostPlan& hp = p->hp;
        std::vector<uint32_t> init_map(size_t(hp.nstr));
        for (int64_t a = 0; a < hp.nstr; ++a) init_map[hp.dyn_col0[a]] = hp.int2ref[a];
        if (cudaMalloc(&p->d_dyn_init, sizeof(uint32_t) * hp.nstr) == cudaSuccess) {
            cudaMemcpy(p->d_dyn_init, init_map.data(), sizeof(uint32_t) * hp.nstr, cudaMemcpyHostToDevice);
            const int64_t r = p->hf_index / hp.nb;  // static address of the Hartree-Fock string (rows keep the static order)
            p->hf_index_dyn = r * hp.nb + int64_t(hp.dyn_col0[r]);
            p->dyn = 1;
        } else {
            cudaGetLastError();
        }
    }
    cudaMemset(p->ticket, 0, sizeof(unsigned));
    cudaMemset(p->grad_ops, 0, sizeof(double) * (n_ops + 8));
    cudaEventRecord(p->rot_event, nullptr);
    *out = p;
    return 0;
}

int tcc_plan_create(int n_qubits, int n_elec, int hcb, int n_ops, const int32_t* ex_ops_flat, const int32_t* ex_op_len,
                    const int32_t* param_ids, int device, tcc_plan** out) {
    return plan_create_impl(n_qubits, n_elec, hcb, n_ops, ex_ops_flat, ex_op_len, param_ids, device, 0, 1, out);
}

int tcc_plan_create_sharded(int n_qubits, int n_elec, int n_ops, const int32_t* ex_ops_flat, const int32_t* ex_op_len,
                            const int32_t* param_ids, int device, int rank, int world, tcc_plan** out) {
    return plan_create_impl(n_qubits, n_elec, 0, n_ops, ex_ops_flat, ex_op_len, param_ids, device, rank, world, out);
}

static void free_sets_ws(tcc_plan* p);

void tcc_plan_destroy(tcc_plan* p) {
    if (!p) return;
    if (p->device >= 0) {
        cudaSetDevice(p->device);
        free_sets_ws(p);
    }
    if (p->device < 0) {
        delete p;
        return;
    }
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (auto& d : p->dhops) {
        cudaFree(d.src);
        cudaFree(d.dst);
    }
    for (auto& b : p->dbatches)
        for (void* q : b.allocs) cudaFree(q);
    cudaFree(p->ket);
    cudaFree(p->bra);
    cudaFree(p->ws1);
    cudaFree(p->ws2);
    cudaFree(p->kb);
    cudaFree(p->kb2);
    cudaFree(p->persist_ctl);
    cudaFree(p->dyn_f);
    cudaFree(p->d_dyn_init);
    cudaFree(p->io1);
    cudaFree(p->io2);
    cudaFree(p->partials);
    cudaFree(p->ticket);
    cudaFree(p->grad_ops);
    cudaFree(p->bpart);
    cudaFree(p->d_rot_fwd);
    cudaFree(p->d_rot_rev);
    cudaFree(p->d_ref2int);
    cudaFree(p->d_int2ref);
    if (p->h_rot) cudaFreeHost(p->h_rot);
    if (p->pinned) cudaFreeHost(p->pinned);
    if (p->rot_event) cudaEventDestroy(p->rot_event);
    p->fan.destroy();
    cudaFree(p->d_links);
    cudaFree(p->d_w);
    cudaFree(p->rows_ws1);
    cudaFree(p->rows_ws2);
    for (int i = 0; i < 3; ++i) cudaFree(p->d_sigma_links[i]);
    cudaFree(p->d_kmap);
    cudaFree(p->d_ell_col);
    cudaFree(p->d_ell_val);
    cudaFree(p->d_strings);
    cudaFree(p->d_diag1);
    cudaFree(p->d_hop);
    cudaFree(p->d_diag2);
    cudaFree(p->d_binom);
    cudaFree(p->d_bitpos);
    delete p;
}

int64_t tcc_civector_size(const tcc_plan* p) { return p ? p->hp.N : 0; }
int64_t tcc_num_strings(const tcc_plan* p) { return p ? p->hp.nstr : 0; }
int tcc_num_ops(const tcc_plan* p) { return p ? int(p->hp.ops.size()) : 0; }
int tcc_num_params(const tcc_plan* p) { return p ? p->hp.n_params : 0; }
int tcc_num_batches(const tcc_plan* p) { return p ? int(p->hp.batches.size()) : 0; }
int64_t tcc_plan_device_bytes(const tcc_plan* p) { return p ? p->dev_bytes : 0; }
int64_t tcc_launch_count(const tcc_plan* p) { return p ? p->launches : 0; }

int64_t tcc_op_support_bytes(const tcc_plan* p, int j) {
    if (!p || j < 0 || j >= int(p->hp.ops.size())) return 0;
    return 32 * p->hp.support_pairs(p->hp.ops[j]);  // 2 elements per pair, each read and written once (8 B)
}

int tcc_batch_info(const tcc_plan* p, int b, int32_t* n_ops, int32_t* active_alpha, int32_t* active_beta, int32_t* n_tiles,
                   int32_t* max_tile_rows, int32_t* max_tile_cols, int32_t* op_indices /* [n_ops of the batch] or NULL */) {
    if (!p || b < 0 || b >= int(p->hp.batches.size())) return fail(TCC_ERR_BAD_ARGUMENT, "batch index out of range");
    const Batch& hb = p->hp.batches[b];
    if (n_ops) *n_ops = int32_t(hb.ops.size());
    if (active_alpha) *active_alpha = hb.A.m;
    if (active_beta) *active_beta = hb.B.m;
    if (n_tiles) *n_tiles = int32_t(hb.A.panels.size() * hb.B.panels.size());
    if (max_tile_rows) *max_tile_rows = hb.A.max_panel;
    if (max_tile_cols) *max_tile_cols = hb.B.max_panel;
    if (op_indices)
        for (size_t i = 0; i < hb.ops.size(); ++i) op_indices[i] = hb.ops[i].op_index;
    return 0;
}

int tcc_batch_pairs(const tcc_plan* p, int b, int pos, int e_alpha, int e_beta, int which, uint32_t* out_host, int capacity,
                    int32_t* n_out, int32_t* round_len) {
    if (!p || !n_out || b < 0 || b >= int(p->hp.batches.size())) return fail(TCC_ERR_BAD_ARGUMENT, "batch index out of range");
    const Batch& hb = p->hp.batches[b];
    const int ma = hb.A.m, mb = hb.B.m;
    if (pos < 0 || pos >= int(hb.ops.size()) || e_alpha < 0 || e_alpha > ma || e_beta < 0 || e_beta > mb || which < 0 || which > 2)
        return fail(TCC_ERR_BAD_ARGUMENT, "bad excitation position, electron count or list kind");
    const size_t slot = ((size_t(pos) * (ma + 1) + e_alpha) * (mb + 1) + e_beta) * 2;
    const uint32_t* src = nullptr;
    int64_t n = 0;
    if (round_len) *round_len = 0;
    if (which == 0) {
        if (slot + 1 >= hb.plist_index.size()) return fail(TCC_ERR_BAD_ARGUMENT, "no pair list");
        src = hb.plist.data() + hb.plist_index[slot];
        n = hb.plist_index[slot + 1];
    } else {
        if (!hb.lean_threads || slot + 1 >= hb.stream_index.size()) {
            *n_out = 0;  // a batch without lean streams (a side without active orbitals)
            return 0;
        }
        int ga = 1, gb = 1;
        for (const PanelDesc& pd : hb.A.panels)
            if (int(pd.e) == e_alpha) ga = std::max<int>(1, pd.Gn);
        for (const PanelDesc& pd : hb.B.panels)
            if (int(pd.e) == e_beta) gb = std::max<int>(1, pd.Gn);
        const int tps = std::max(1, hb.lean_threads / (ga * gb));
        src = (which == 1 ? hb.stream8.data() : hb.stream16.data()) + hb.stream_index[slot];
        n = int64_t(hb.stream_index[slot + 1]) * tps;
        if (round_len) *round_len = tps;
    }
    *n_out = int32_t(n);
    if (out_host) {
        if (capacity < n) return fail(TCC_ERR_BAD_ARGUMENT, "output buffer too small");
        for (int64_t i = 0; i < n; ++i) out_host[i] = src[i];
    }
    return 0;
}

int tcc_ci_strings_spin(const tcc_plan* p, uint64_t* out_host) {
    if (!p || !out_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    for (int64_t r = 0; r < p->hp.nstr; ++r) out_host[r] = p->hp.strings[p->hp.ref2int[r]];
    return 0;
}

int tcc_storage_order(const tcc_plan* p, uint32_t* ref2int_host) {
    if (!p || !ref2int_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    for (int64_t r = 0; r < p->hp.nstr; ++r) ref2int_host[r] = p->hp.ref2int[r];
    return 0;
}

int tcc_excitation_tables(const tcc_plan* p, int j, uint64_t* perm_host, int8_t* phase_host, int8_t* f2_host) {
    if (!p || !perm_host || !phase_host || !f2_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (j < 0 || j >= int(p->hp.ops.size())) return fail(TCC_ERR_BAD_ARGUMENT, "excitation index out of range");
    p->hp.expand_tables(p->hp.ops[j], perm_host, phase_host, f2_host);
    return 0;
}

// strings, binomials and the storage bit positions of an hcb plan on the device (sigma_hcb_kernel, hcb_rdm_kernel)
static int ensure_hcb_tables(tcc_plan* p) {
    if (p->d_strings) return 0;
    HostPlan& hp = p->hp;
    const int n = hp.n;
    std::vector<int64_t> bin(size_t(n + 2) * (n + 2));
    for (int i = 0; i < n + 2; ++i)
        for (int j = 0; j < n + 2; ++j) bin[size_t(i) * (n + 2) + j] = hp.binom[i][j];
    CUDA_TRY(dev_alloc(p, &p->d_strings, size_t(hp.nstr)));
    CUDA_TRY(dev_alloc(p, &p->d_binom, bin.size()));
    CUDA_TRY(dev_alloc(p, &p->d_bitpos, size_t(32)));
    CUDA_TRY(cudaMemcpy(p->d_bitpos, hp.bitpos.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p->d_strings, hp.strings.data(), hp.nstr * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p->d_binom, bin.data(), bin.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    return 0;
}

int tcc_set_hamiltonian(tcc_plan* p, const double* int1e, const double* int2e) {
    if (!p || !int1e || !int2e) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    HostPlan& hp = p->hp;
    const int n = hp.n;
    if (hp.hcb) {
        std::vector<double> d1(n), hop(size_t(n) * n, 0.0), d2(size_t(n) * n, 0.0);
        auto i2 = [&](int a, int b, int c, int d) { return int2e[((size_t(a) * n + b) * n + c) * n + d]; };
        for (int q = 0; q < n; ++q) {
            d1[q] = 2 * int1e[q * n + q] + i2(q, q, q, q);
            for (int r = 0; r < q; ++r) {
                hop[q * n + r] = i2(q, r, q, r);
                d2[q * n + r] = 4 * i2(q, q, r, r) - 2 * i2(q, r, q, r);
            }
        }
        if (int rc = ensure_hcb_tables(p)) return rc;
        if (!p->d_diag1) {
            CUDA_TRY(dev_alloc(p, &p->d_diag1, size_t(n)));
            CUDA_TRY(dev_alloc(p, &p->d_hop, size_t(n) * n));
            CUDA_TRY(dev_alloc(p, &p->d_diag2, size_t(n) * n));
        }
        CUDA_TRY(cudaMemcpy(p->d_diag1, d1.data(), n * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->d_hop, hop.data(), hop.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->d_diag2, d2.data(), d2.size() * sizeof(double), cudaMemcpyHostToDevice));
        p->has_h = true;
        return 0;
    }
    std::vector<double> h2e = absorb_h1e(int1e, int2e, n, hp.n_elec, 0.5);
    if (hp.links.empty()) hp.build_links();
    hp.build_same_spin_ell(h2e);
    const int n2 = n * n;
    // W[pq,rs] = h2e[pq,rs] + h2e[rs,pq].  When W[pq,rs] == W[pq,sr] (true for real-orbital integrals; CHECKED here,
    // the engine follows direct_nosym and assumes nothing) the D intermediate is folded onto pairs r <= s, which
    // halves the K dimension of the opposite-spin GEMM.
    std::vector<double> wfull(size_t(n2) * n2);
    double wmax = 0.0;
    for (int pq = 0; pq < n2; ++pq)
        for (int rs = 0; rs < n2; ++rs) {
            wfull[size_t(pq) * n2 + rs] = h2e[size_t(pq) * n2 + rs] + h2e[size_t(rs) * n2 + pq];
            wmax = std::max(wmax, std::fabs(wfull[size_t(pq) * n2 + rs]));
        }
    bool sym = env_int("TCC_B200_SIGMA_PAIR_SYM", 1) != 0;
    for (int pq = 0; pq < n2 && sym; ++pq)
        for (int r = 0; r < n && sym; ++r)
            for (int s2 = 0; s2 < r; ++s2)
                if (std::fabs(wfull[size_t(pq) * n2 + r * n + s2] - wfull[size_t(pq) * n2 + s2 * n + r]) > 1e-14 * std::max(wmax, 1e-300)) {
                    sym = false;
                    break;
                }
    p->w_pair_symmetric = sym;
    std::vector<uint16_t> kmap(n2);
    int kdim = 0;
    if (sym) {
        std::vector<int> pair_id(n2, 0);
        for (int r = 0; r < n; ++r)
            for (int s2 = r; s2 < n; ++s2) pair_id[r * n + s2] = pair_id[s2 * n + r] = kdim++;
        for (int pq = 0; pq < n2; ++pq) kmap[pq] = uint16_t(pair_id[pq]);
    } else {
        kdim = n2;
        for (int p_ = 0; p_ < n; ++p_)
            for (int q_ = 0; q_ < n; ++q_) kmap[p_ * n + q_] = uint16_t(q_ * n + p_);
    }
    p->n2p = (kdim + 3) / 4 * 4;
    std::vector<double> w(size_t(n2) * p->n2p, 0.0);
    for (int pq = 0; pq < n2; ++pq)
        for (int r = 0; r < n; ++r)
            for (int s2 = 0; s2 < n; ++s2) {
                // column of W that multiplies D row kmap[(s2, r)] ... D[(q,p)] collects the forward link E_pq
                if (sym) {
                    if (r <= s2) w[size_t(pq) * p->n2p + kmap[r * n + s2]] = wfull[size_t(pq) * n2 + r * n + s2];
                } else {
                    w[size_t(pq) * p->n2p + r * n + s2] = wfull[size_t(pq) * n2 + r * n + s2];
                }
            }
    cudaFree(p->d_w);
    cudaFree(p->d_ell_col);
    cudaFree(p->d_ell_val);
    p->d_w = nullptr;
    cudaFree(p->d_kmap);
    p->d_kmap = nullptr;
    p->d_ell_col = nullptr;
    p->d_ell_val = nullptr;
    if (!p->d_links) {
        CUDA_TRY(dev_alloc(p, &p->d_links, hp.links.size()));
        CUDA_TRY(cudaMemcpy(p->d_links, hp.links.data(), hp.links.size() * sizeof(Link), cudaMemcpyHostToDevice));
        if (hp.shard_world > 1) {
            // sharded sigma: per-pass alpha link tables of this rank's rows (local target rows)
            hp.build_sigma_pass_links();
            for (int i = 0; i < 3; ++i) {
                if (hp.sigma_links[i].empty()) continue;
                CUDA_TRY(dev_alloc(p, &p->d_sigma_links[i], hp.sigma_links[i].size()));
                CUDA_TRY(cudaMemcpy(p->d_sigma_links[i], hp.sigma_links[i].data(), hp.sigma_links[i].size() * sizeof(Link),
                                    cudaMemcpyHostToDevice));
                std::vector<Link>().swap(hp.sigma_links[i]);
            }
        }
    }
    CUDA_TRY(dev_alloc(p, &p->d_w, w.size()));
    CUDA_TRY(cudaMemcpy(p->d_w, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(dev_alloc(p, &p->d_kmap, kmap.size()));
    CUDA_TRY(cudaMemcpy(p->d_kmap, kmap.data(), kmap.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CUDA_TRY(dev_alloc(p, &p->d_ell_col, hp.ell_col.size()));
    CUDA_TRY(dev_alloc(p, &p->d_ell_val, hp.ell_val.size()));
    CUDA_TRY(cudaMemcpy(p->d_ell_col, hp.ell_col.data(), hp.ell_col.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p->d_ell_val, hp.ell_val.data(), hp.ell_val.size() * sizeof(double), cudaMemcpyHostToDevice));
    // host copies of the ELL table are no longer needed
    std::vector<uint32_t>().swap(hp.ell_col);
    std::vector<double>().swap(hp.ell_val);
    p->has_h = true;
    return 0;
}

// sigma = H c for n_sets vectors stored back to back (storage order); ws1 / ws2: scratch of the same size
static int apply_h_sets(tcc_plan* p, const double* c, double* sigma, double* ws1, double* ws2, int n_sets, cudaStream_t st) {
    HostPlan& hp = p->hp;
    if (!p->has_h) return fail(TCC_ERR_NO_HAMILTONIAN, "tcc_set_hamiltonian has not been called");
    if (hp.hcb) {
        launch_sigma_hcb(c, sigma, hp.nstr, p->d_strings, hp.n, hp.k, p->d_diag1, p->d_hop, p->d_diag2, p->d_binom, hp.n + 2,
                         p->d_bitpos, n_sets, st);
        p->launches += 1;
        return 0;
    }
    const int64_t vec = hp.N * 8 * n_sets;
    // alpha-alpha: sigma(a', :) = sum_a Hss[a', a] c(a, :)
    p->prof.mark(TCC_PROF_SIGMA_SS, st);
    launch_spmm_rows(c, sigma, hp.na, hp.nb, p->d_ell_col, p->d_ell_val, hp.ell_width, false, n_sets, st);
    // beta-beta on the transposed matrix, then added back transposed
    launch_transpose(c, ws1, hp.na, hp.nb, false, n_sets, st);
    launch_spmm_rows(ws1, ws2, hp.nb, hp.na, p->d_ell_col, p->d_ell_val, hp.ell_width, false, n_sets, st);
    launch_transpose(ws2, sigma, hp.nb, hp.na, true, n_sets, st);
    p->prof.count(TCC_PROF_SIGMA_SS, 4, 9 * vec);  // 2+2+2+3 vector passes
    // alpha-beta on the FP64 tensor path
    p->prof.mark(TCC_PROF_SIGMA_AB, st);
    launch_sigma_ab(c, sigma, hp.na, hp.nb, p->d_links, hp.ml, p->d_w, hp.n, p->n2p, p->d_kmap, n_sets, st);
    p->prof.count(TCC_PROF_SIGMA_AB, (hp.na + 65534) / 65535, 3 * vec);
    p->prof.mark(-1, st);
    p->launches += 4 + (hp.na + 65534) / 65535;
    return 0;
}

static int apply_h_impl(tcc_plan* p, const double* c, double* sigma, cudaStream_t st) {
    if (!p->hp.hcb) {
        if (int rc = ensure_vec(p, &p->ws1)) return rc;
        if (int rc = ensure_vec(p, &p->ws2)) return rc;
    }
    return apply_h_sets(p, c, sigma, p->ws1, p->ws2, 1, st);
}

// ---- sweeps on storage-order vectors ------------------------------------------------------------
static int prepare_rot(tcc_plan* p, const double* params, cudaStream_t st) {
    const HostPlan& hp = p->hp;
    const int m = int(hp.ops.size());
    if (m == 0) return 0;
    CUDA_TRY(cudaEventSynchronize(p->rot_event));  // the previous upload has left the pinned buffer
    for (int j = 0; j < m; ++j) {
        const double th = params[hp.ops[j].param];
        const double c = std::cos(th), s = std::sin(th) * hp.ops[j].s0;
        p->h_rot[j] = make_double2(c, s);
        p->h_rot[m + j] = make_double2(c, -s);
    }
    CUDA_TRY(cudaMemcpyAsync(p->d_rot_fwd, p->h_rot, sizeof(double2) * m, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(p->d_rot_rev, p->h_rot + m, sizeof(double2) * m, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(p->rot_event, st));
    return 0;
}

static void rotate_op(tcc_plan* p, int j, double theta, double* c, cudaStream_t st) {
    const OpDesc& op = p->hp.ops[j];
    p->prof.mark(TCC_PROF_FWD_BETA + op.kind, st);
    launch_rotate(c, p->hp.nb, row_side(p, op), col_side(p, op), std::cos(theta), std::sin(theta) * op.s0, st);
    p->prof.count(TCC_PROF_FWD_BETA + op.kind, 1, 32 * p->hp.support_pairs(op));
    p->launches += 1;
}

static void reverse_op(tcc_plan* p, int j, double theta, double* ket, double* bra, double* grad_out, cudaStream_t st) {
    const OpDesc& op = p->hp.ops[j];
    p->prof.mark(TCC_PROF_REV_BETA + op.kind, st);
    launch_reverse(ket, bra, p->hp.nb, row_side(p, op), col_side(p, op), std::cos(theta), -std::sin(theta) * op.s0,
                   double(op.s0), p->partials, p->ticket, grad_out, st);
    p->prof.count(TCC_PROF_REV_BETA + op.kind, 1, 64 * p->hp.support_pairs(op));
    p->launches += 1;
}

// exp(theta_M G_M) ... exp(theta_1 G_1) applied to a storage-order vector (evolve_civector.py:165-177)
// dyn_start: null = in place in the static order; otherwise the buffer (vec or p->dyn_f, chosen by the parity of the number
// of batches so that the last batch writes vec) that holds the state in the first forward layout -- every batch then
// reads one buffer and writes the other, the last one in the static order (dynamic column layout, plan.hpp)
static int forward_sweep(tcc_plan* p, const double* params, double* vec, cudaStream_t st, double* dyn_start = nullptr) {
    HostPlan& hp = p->hp;
    if (p->mode == 1) {
        for (size_t j = 0; j < hp.ops.size(); ++j) rotate_op(p, int(j), params[hp.ops[j].param], vec, st);
    } else {
        if (int rc = prepare_rot(p, params, st)) return rc;
        p->prof.mark(TCC_PROF_FWD_BATCH, st);
        const size_t nbt = p->dbatches.size();
        if (p->btimer.on) p->btimer.ensure(nbt);
        double* cur = dyn_start ? dyn_start : vec;
        double* other = dyn_start ? (dyn_start == vec ? p->dyn_f : vec) : nullptr;
        for (size_t b = 0; b < nbt; ++b) {
            BatchDevSet& db = p->dbatches[b];
            if (p->btimer.on) cudaEventRecord(p->btimer.ev[b], st);
            launch_batch_forward(cur, other, hp.nb, db.fwd, p->d_rot_fwd, db.rects, p->batch_threads, 1, SetStrides{0, 0, 0, 0}, st, &p->fan,
                                 p->persist_ctl, p->sms, p->persist_min_waves);
            if (other) std::swap(cur, other);
            p->prof.count(TCC_PROF_FWD_BATCH, 1, 16 * hp.N);
            p->launches += 1;
        }
        if (p->btimer.on) cudaEventRecord(p->btimer.ev[nbt], st);
    }
    p->prof.mark(-1, st);
    return 0;
}

// reverse sweep (evolve_civector.py:221-243): grad_ops[j] = <bra| G_j |ket> after undoing factors M..j
static int reverse_sweep(tcc_plan* p, const double* params, double* ket, double* bra, cudaStream_t st) {
    HostPlan& hp = p->hp;
    if (p->mode == 1) {
        for (int j = int(hp.ops.size()) - 1; j >= 0; --j) reverse_op(p, j, params[hp.ops[j].param], ket, bra, p->grad_ops + j, st);
    } else {
        // d_rot_rev was uploaded together with d_rot_fwd by the forward sweep of the same parameters
        p->prof.mark(TCC_PROF_REV_BATCH, st);
        const int nbt = int(p->dbatches.size());
        // ticket counters of the persistent launches: self-resetting, cleared here anyway so that an aborted launch of an
        // earlier evaluation cannot leak into this one
        cudaMemsetAsync(p->persist_ctl, 0, sizeof(unsigned) * 64, st);
        // Interleaved mode: (ket, bra) become one 16-byte element per determinant for the whole sweep (one pass to
        // build it; nothing is converted back -- only the gradients leave the sweep).  Batches without lean streams
        // (a side without active orbitals) work on the two vectors: the data is converted around them.
        bool use_il = p->rev_interleave != 0;
        if (use_il) {
            bool any = false;
            for (auto& db : p->dbatches) any = any || db.rev.lean;
            use_il = any;
        }
        if (use_il && !p->kb) {
            double* kb = nullptr;
            if (cudaMalloc(&kb, sizeof(double) * 2 * size_t(hp.N)) == cudaSuccess) {
                p->kb = kb;
                p->dev_bytes += int64_t(sizeof(double)) * 2 * hp.N;
            } else {
                cudaGetLastError();  // no memory for the interleaved copy: sweep on the two vectors
                use_il = false;
            }
        }
        // dynamic column layout: every batch reads one interleaved buffer (the first one in the static order) and writes the
        // other in the order of the batch that follows; needs lean streams in every batch
        bool dyn = use_il && p->dyn && nbt >= 2;
        if (dyn)
            for (auto& db : p->dbatches) dyn = dyn && db.rev.lean;
        if (dyn && !p->kb2) {
            double* kb2 = nullptr;
            if (cudaMalloc(&kb2, sizeof(double) * 2 * size_t(hp.N)) == cudaSuccess) {
                p->kb2 = kb2;
                p->dev_bytes += int64_t(sizeof(double)) * 2 * hp.N;
            } else {
                cudaGetLastError();
                dyn = false;
            }
        }
        double* kb_cur = p->kb;
        double* kb_other = dyn ? p->kb2 : nullptr;
        bool in_kb = false;
        for (int b = nbt - 1; b >= 0; --b) {
            BatchDevSet& db = p->dbatches[b];
            if (p->btimer.on) cudaEventRecord(p->btimer.ev[nbt + 1 + (nbt - 1 - b)], st);
            if (use_il && db.rev.lean) {
                if (!in_kb) {
                    launch_interleave(ket, bra, p->kb, hp.N, st);
                    p->launches += 1;
                    in_kb = true;
                }
                launch_batch_reverse_il(kb_cur, kb_other, hp.nb, db.rev, p->d_rot_rev, db.rects_rev, p->batch_threads_rev, p->bpart,
                                        p->grad_ops, 1, SetStrides{0, 0, 0, 0}, st, &p->fan, p->persist_ctl, p->sms, p->persist_min_waves);
                if (kb_other) std::swap(kb_cur, kb_other);
            } else {
                if (in_kb) {
                    launch_deinterleave(p->kb, ket, bra, hp.N, st);
                    p->launches += 1;
                    in_kb = false;
                }
                launch_batch_reverse(ket, bra, hp.nb, db.rev, p->d_rot_rev, db.rects_rev, p->batch_threads_rev, p->bpart, p->grad_ops, 1,
                                     SetStrides{0, 0, 0, 0}, st, &p->fan);
            }
            p->prof.count(TCC_PROF_REV_BATCH, 1, 32 * hp.N);  // the tiny gradient-reduce launch rides along
            p->launches += 2;
        }
        if (p->btimer.on) cudaEventRecord(p->btimer.ev[2 * nbt + 1], st);
    }
    p->prof.mark(-1, st);
    return 0;
}

// state preparation into a storage-order buffer
static int prepare_state(tcc_plan* p, const double* params, const double* init_ref_dev, double* vec, cudaStream_t st) {
    HostPlan& hp = p->hp;
    // dynamic column layout: the state is prepared in the first forward layout, in the buffer from which an alternating
    // sequence of batches ends in `vec`
    bool dyn = p->dyn && p->mode != 1 && p->dbatches.size() >= 2 && init_ref_dev != vec;
    if (dyn && !p->dyn_f) {
        double* buf = nullptr;
        if (cudaMalloc(&buf, sizeof(double) * size_t(hp.N)) == cudaSuccess) {
            p->dyn_f = buf;
            p->dev_bytes += int64_t(sizeof(double)) * hp.N;
        } else {
            cudaGetLastError();  // no memory for the second buffer: in place, static order
            p->dyn = 0;
            dyn = false;
        }
    }
    double* start = dyn ? (p->dbatches.size() % 2 == 0 ? vec : p->dyn_f) : vec;
    if (init_ref_dev) {
        if (dyn) {
            launch_permute(init_ref_dev, start, hp.na, hp.nb, p->d_int2ref, p->d_dyn_init, st);
            p->launches += 1;
        } else {
            to_storage(p, init_ref_dev, vec, st);
        }
    } else {
        launch_set_unit(start, hp.N, dyn ? p->hf_index_dyn : p->hf_index, 1, st);
        p->launches += 1;
    }
    return forward_sweep(p, params, vec, st, dyn ? start : nullptr);
}

int tcc_apply_h(tcc_plan* p, const double* c_dev, double* sigma_dev, void* stream) {
    if (!p || !c_dev || !sigma_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (c_dev == sigma_dev) return fail(TCC_ERR_BAD_ARGUMENT, "c and sigma must not alias");
    if (int rc_dev = use_device(p)) return rc_dev;
    cudaStream_t st = cudaStream_t(stream);
    if (p->hp.identity_layout) {
        if (int rc = apply_h_impl(p, c_dev, sigma_dev, st)) return rc;
    } else {
        if (int rc = ensure_vec(p, &p->ket)) return rc;
        if (int rc = ensure_vec(p, &p->bra)) return rc;
        to_storage(p, c_dev, p->ket, st);
        if (int rc = apply_h_impl(p, p->ket, p->bra, st)) return rc;
        from_storage(p, p->bra, sigma_dev, st);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int tcc_civector(tcc_plan* p, const double* params_host, const double* init_dev, double* out_dev, void* stream) {
    if (!p || !out_dev || (!params_host && p->hp.n_params > 0)) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    cudaStream_t st = cudaStream_t(stream);
    if (p->hp.identity_layout) {
        if (int rc = prepare_state(p, params_host, init_dev, out_dev, st)) return rc;
    } else {
        if (int rc = ensure_vec(p, &p->ket)) return rc;
        if (int rc = prepare_state(p, params_host, init_dev, p->ket, st)) return rc;
        from_storage(p, p->ket, out_dev, st);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int tcc_to_storage_order(tcc_plan* p, const double* ref_dev, double* storage_dev, void* stream) {
    if (!p || !ref_dev || !storage_dev || ref_dev == storage_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null or aliased argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    to_storage(p, ref_dev, storage_dev, cudaStream_t(stream));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int tcc_from_storage_order(tcc_plan* p, const double* storage_dev, double* ref_dev, void* stream) {
    if (!p || !ref_dev || !storage_dev || ref_dev == storage_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null or aliased argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    from_storage(p, storage_dev, ref_dev, cudaStream_t(stream));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int tcc_rotate(tcc_plan* p, int j, double theta, double* c_dev, void* stream) {
    if (!p || !c_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (j < 0 || j >= int(p->hp.ops.size())) return fail(TCC_ERR_BAD_ARGUMENT, "excitation index out of range");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    rotate_op(p, j, theta, c_dev, cudaStream_t(stream));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int tcc_reverse_step(tcc_plan* p, int j, double theta, double* ket_dev, double* bra_dev, double* grad_dev, void* stream) {
    if (!p || !ket_dev || !bra_dev || !grad_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (j < 0 || j >= int(p->hp.ops.size())) return fail(TCC_ERR_BAD_ARGUMENT, "excitation index out of range");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    reverse_op(p, j, theta, ket_dev, bra_dev, grad_dev, cudaStream_t(stream));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int energy_impl(tcc_plan* p, const double* params, const double* init_dev, double* energy, double* grad, cudaStream_t st) {
    HostPlan& hp = p->hp;
    if (!p->has_h) return fail(TCC_ERR_NO_HAMILTONIAN, "tcc_set_hamiltonian has not been called");
    if (int rc = ensure_vec(p, &p->ket)) return rc;
    if (int rc = ensure_vec(p, &p->bra)) return rc;
    if (int rc = prepare_state(p, params, init_dev, p->ket, st)) return rc;
    if (int rc = apply_h_impl(p, p->ket, p->bra, st)) return rc;
    const int m = int(hp.ops.size());
    p->prof.mark(TCC_PROF_DOT, st);
    launch_dot(p->bra, p->ket, hp.N, p->partials, p->ticket, p->grad_ops + m, st);
    p->prof.count(TCC_PROF_DOT, 1, 16 * hp.N);
    p->prof.mark(-1, st);
    p->launches += 1;
    if (grad)
        if (int rc = reverse_sweep(p, params, p->ket, p->bra, st)) return rc;
    int first = grad ? 0 : m;
    p->prof.mark(-1, st);
    CUDA_TRY(cudaMemcpyAsync(p->pinned + first, p->grad_ops + first, sizeof(double) * (m + 1 - first), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->prof.resolve();
    CUDA_TRY(cudaGetLastError());
    *energy = p->pinned[m];
    if (grad) {
        for (int i = 0; i < hp.n_params; ++i) grad[i] = 0.0;
        for (int j = 0; j < m; ++j) grad[hp.ops[j].param] += p->pinned[j];  // evolve_civector.py:213-216
        for (int i = 0; i < hp.n_params; ++i) grad[i] *= 2.0;                // :218
    }
    return 0;
}

int tcc_energy(tcc_plan* p, const double* params_host, const double* init_dev, double* energy_host, void* stream) {
    if (!p || !energy_host || (!params_host && p->hp.n_params > 0)) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    return energy_impl(p, params_host, init_dev, energy_host, nullptr, cudaStream_t(stream));
}

int tcc_energy_and_grad(tcc_plan* p, const double* params_host, const double* init_dev, double* energy_host,
                        double* grad_host, void* stream) {
    if (!p || !energy_host || !grad_host || (!params_host && p->hp.n_params > 0)) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    return energy_impl(p, params_host, init_dev, energy_host, grad_host, cudaStream_t(stream));
}

// ADAPT-VQE pool screening (docs/source/tutorial_jupyter/adapt_vqe.ipynb cell 9: for every pool operator
// 2 * bra . apply_excitation(ket, op) with bra = H ket): the excitation list of THIS plan is the operator pool, and
// <sigma| G_k |c> for all k is the reverse sweep with every angle zero -- the rotations are exact identities, so one
// batched launch set over (c, sigma) yields all pool gradients; nothing is applied in sequence and the order is
// irrelevant.  c_dev, sigma_dev: reference-order CI vectors on the device (not modified); out_host[k] = <sigma|G_k|c>.
int tcc_pool_gradients(tcc_plan* p, const double* c_dev, const double* sigma_dev, double* out_host, void* stream) {
    if (!p || !c_dev || !sigma_dev || !out_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    HostPlan& hp = p->hp;
    const int m = int(hp.ops.size());
    if (m == 0) return 0;
    cudaStream_t st = cudaStream_t(stream);
    if (int rc = ensure_vec(p, &p->ket)) return rc;
    if (int rc = ensure_vec(p, &p->bra)) return rc;
    to_storage(p, c_dev, p->ket, st);
    to_storage(p, sigma_dev, p->bra, st);
    std::vector<double> zeros(size_t(std::max(hp.n_params, 1)), 0.0);
    if (p->mode != 1)
        if (int rc = prepare_rot(p, zeros.data(), st)) return rc;
    if (int rc = reverse_sweep(p, zeros.data(), p->ket, p->bra, st)) return rc;
    CUDA_TRY(cudaMemcpyAsync(p->pinned, p->grad_ops, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->prof.resolve();
    CUDA_TRY(cudaGetLastError());
    for (int j = 0; j < m; ++j) out_host[j] = p->pinned[j];
    return 0;
}

// Batched-theta evaluation (API gap i): n_sets parameter vectors evaluated together, one
// grid dimension per set; the CI vectors of all sets of a chunk are resident at once.
struct SetsWorkspace {
    int cap = 0;
    double *ket = nullptr, *bra = nullptr, *ws1 = nullptr, *ws2 = nullptr, *grad = nullptr, *bpart = nullptr;
    double2 *rot_fwd = nullptr, *rot_rev = nullptr, *h_rot = nullptr;
    double* h_grad = nullptr;
};
static std::map<tcc_plan*, SetsWorkspace> g_sets_ws;

static void free_sets_ws(tcc_plan* p) {
    auto it = g_sets_ws.find(p);
    if (it == g_sets_ws.end()) return;
    SetsWorkspace& w = it->second;
    cudaFree(w.ket);
    cudaFree(w.bra);
    cudaFree(w.ws1);
    cudaFree(w.ws2);
    cudaFree(w.grad);
    cudaFree(w.bpart);
    cudaFree(w.rot_fwd);
    cudaFree(w.rot_rev);
    if (w.h_rot) cudaFreeHost(w.h_rot);
    if (w.h_grad) cudaFreeHost(w.h_grad);
    g_sets_ws.erase(it);
}

int tcc_energy_and_grad_batch(tcc_plan* p, int n_sets, const double* params_host, double* energies_host, double* grads_host,
                              void* stream) {
    if (!p || n_sets < 0 || !energies_host || (!params_host && p->hp.n_params > 0 && n_sets > 0))
        return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    HostPlan& hp = p->hp;
    if (!p->has_h) return fail(TCC_ERR_NO_HAMILTONIAN, "tcc_set_hamiltonian has not been called");
    if (n_sets == 0) return 0;
    cudaStream_t st = cudaStream_t(stream);
    const int m = int(hp.ops.size()), np_ = hp.n_params;
    size_t max_part = 1;
    for (auto& db : p->dbatches) max_part = std::max(max_part, size_t(db.nops) * size_t(db.n_tiles));
    // chunk size: all vectors of a chunk are resident (4 vectors per set) within the memory budget
    const double budget = double(env_int("TCC_B200_BATCH_MEM_MB", 24576)) * 1048576.0;
    const double per_set = 4.0 * double(hp.N) * 8 + double(max_part) * 8 + double(m + 1) * 8 * 5;
    int cap = int(std::max(1.0, std::min(double(std::min(n_sets, 8192)), budget / per_set)));
    SetsWorkspace& w = g_sets_ws[p];
    if (w.cap < cap) {
        free_sets_ws(p);
        SetsWorkspace& nw = g_sets_ws[p];
        const size_t vec = size_t(hp.N) * cap;
        if (cudaMalloc(&nw.ket, vec * 8) != cudaSuccess || cudaMalloc(&nw.bra, vec * 8) != cudaSuccess ||
            cudaMalloc(&nw.ws1, vec * 8) != cudaSuccess || cudaMalloc(&nw.ws2, vec * 8) != cudaSuccess ||
            cudaMalloc(&nw.grad, size_t(cap) * (m + 1) * 8) != cudaSuccess || cudaMalloc(&nw.bpart, max_part * cap * 8) != cudaSuccess ||
            cudaMalloc(&nw.rot_fwd, size_t(cap) * (m + 1) * sizeof(double2)) != cudaSuccess ||
            cudaMalloc(&nw.rot_rev, size_t(cap) * (m + 1) * sizeof(double2)) != cudaSuccess ||
            cudaMallocHost(&nw.h_rot, size_t(cap) * (m + 1) * 2 * sizeof(double2)) != cudaSuccess ||
            cudaMallocHost(&nw.h_grad, size_t(cap) * (m + 1) * 8) != cudaSuccess) {
            free_sets_ws(p);
            return fail(TCC_ERR_CUDA, std::string("batched workspace allocation: ") + cudaGetErrorString(cudaGetLastError()));
        }
        nw.cap = cap;
    }
    SetsWorkspace& ws = g_sets_ws[p];
    cap = ws.cap;
    for (int s0 = 0; s0 < n_sets; s0 += cap) {
        const int ns = std::min(cap, n_sets - s0);
        double2* hf = ws.h_rot;
        double2* hr = ws.h_rot + size_t(cap) * (m + 1);
        for (int s = 0; s < ns; ++s) {
            const double* prm = params_host + size_t(s0 + s) * np_;
            for (int j = 0; j < m; ++j) {
                const double th = prm[hp.ops[j].param];
                const double c = std::cos(th), sn = std::sin(th) * hp.ops[j].s0;
                hf[size_t(s) * m + j] = make_double2(c, sn);
                hr[size_t(s) * m + j] = make_double2(c, -sn);
            }
        }
        if (m > 0) {
            CUDA_TRY(cudaMemcpyAsync(ws.rot_fwd, hf, sizeof(double2) * size_t(ns) * m, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(ws.rot_rev, hr, sizeof(double2) * size_t(ns) * m, cudaMemcpyHostToDevice, st));
        }
        launch_set_unit(ws.ket, hp.N, p->hf_index, ns, st);
        p->launches += 1;
        for (auto& db : p->dbatches) {
            launch_batch_forward(ws.ket, nullptr, hp.nb, db.fwd, ws.rot_fwd, db.rects, p->batch_threads, ns,
                                 SetStrides{hp.N, m, 0, 0}, st, &p->fan);
            p->launches += 1;
        }
        if (int rc = apply_h_sets(p, ws.ket, ws.bra, ws.ws1, ws.ws2, ns, st)) return rc;
        launch_dot_sets(ws.bra, ws.ket, hp.N, ns, ws.grad + m, m + 1, st);
        p->launches += 1;
        if (grads_host) {
            for (int b = int(p->dbatches.size()) - 1; b >= 0; --b) {
                BatchDevSet& db = p->dbatches[b];
                launch_batch_reverse(ws.ket, ws.bra, hp.nb, db.rev, ws.rot_rev, db.rects_rev, p->batch_threads_rev, ws.bpart, ws.grad, ns,
                                     SetStrides{hp.N, m, int64_t(max_part), m + 1}, st, &p->fan);
                p->launches += 2;
            }
        }
        CUDA_TRY(cudaMemcpyAsync(ws.h_grad, ws.grad, sizeof(double) * size_t(ns) * (m + 1), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
        for (int s = 0; s < ns; ++s) {
            const double* g = ws.h_grad + size_t(s) * (m + 1);
            energies_host[s0 + s] = g[m];
            if (grads_host) {
                double* out = grads_host + size_t(s0 + s) * np_;
                for (int i = 0; i < np_; ++i) out[i] = 0.0;
                for (int j = 0; j < m; ++j) out[hp.ops[j].param] += g[j];
                for (int i = 0; i < np_; ++i) out[i] *= 2.0;
            }
        }
    }
    return 0;
}

int tcc_apply_excitation(tcc_plan* p, const double* c_dev, const int32_t* ex_op, int ex_op_len, double* out_dev, void* stream) {
    if (!p || !c_dev || !ex_op || !out_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (c_dev == out_dev) return fail(TCC_ERR_BAD_ARGUMENT, "input and output must not alias");
    if (int rc_dev = use_device(p)) return rc_dev;
    OpDesc op;
    int rc = p->hp.parse_op(ex_op, ex_op_len, 0, &op);
    if (rc) return fail(rc, p->hp.error);
    p->hp.attach_hops(&op);
    if (op.hop_a >= 0)
        if (int r2 = upload_hop(p, op.hop_a)) return r2;
    if (op.hop_b >= 0)
        if (int r2 = upload_hop(p, op.hop_b)) return r2;
    cudaStream_t st = cudaStream_t(stream);
    const double* src = c_dev;
    double* dst = out_dev;
    if (!p->hp.identity_layout) {
        if (int r2 = ensure_vec(p, &p->ket)) return r2;
        if (int r2 = ensure_vec(p, &p->bra)) return r2;
        to_storage(p, c_dev, p->ket, st);
        src = p->ket;
        dst = p->bra;
    }
    CUDA_TRY(cudaMemsetAsync(dst, 0, p->hp.N * sizeof(double), st));
    launch_apply_g(src, dst, p->hp.nb, row_side(p, op), col_side(p, op), double(op.s0), st);
    p->launches += 1;
    if (!p->hp.identity_layout) from_storage(p, p->bra, out_dev, st);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- reduced density matrices (ucc.py:755-852 -> pyscf direct_spin1.make_rdm1 / make_rdm12) ----------------------
// rdm1[p,q] = <q+ p> and rdm2[p,q,r,s] = <p+ r+ s q>, spin-traced, of a reference-order CI vector on the device.
int tcc_make_rdm12(tcc_plan* p, const double* c_dev, double* rdm1_host, double* rdm2_host, void* stream) {
    if (!p || !c_dev || (!rdm1_host && !rdm2_host)) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    HostPlan& hp = p->hp;
    if (hp.hcb) return fail(TCC_ERR_BAD_ARGUMENT, "RDMs are defined for the fermionic (non-hcb) CI vector only (ucc.py:786)");
    cudaStream_t st = cudaStream_t(stream);
    if (hp.links.empty()) hp.build_links();
    if (!p->d_links) {
        CUDA_TRY(dev_alloc(p, &p->d_links, hp.links.size()));
        CUDA_TRY(cudaMemcpy(p->d_links, hp.links.data(), hp.links.size() * sizeof(Link), cudaMemcpyHostToDevice));
    }
    const double* src = c_dev;
    if (!hp.identity_layout) {
        if (int rc = ensure_vec(p, &p->ket)) return rc;
        to_storage(p, c_dev, p->ket, st);
        src = p->ket;
    }
    const int n = hp.n, n2 = n * n;
    double* d_gram = nullptr;
    CUDA_TRY(cudaMalloc(&d_gram, sizeof(double) * size_t(n2) * n2));
    CUDA_TRY(cudaMemsetAsync(d_gram, 0, sizeof(double) * size_t(n2) * n2, st));
    const int n_blocks = (n2 + 63) / 64, pairs = n_blocks * (n_blocks + 1) / 2;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, p->device));
    const int slices = std::max(1, 4 * prop.multiProcessorCount / pairs);
    launch_rdm_gram(src, hp.na, hp.nb, p->d_links, hp.ml, n, d_gram, slices, st);
    p->launches += 1;
    std::vector<double> gram(size_t(n2) * n2);
    cudaError_t ce = cudaMemcpyAsync(gram.data(), d_gram, sizeof(double) * gram.size(), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(d_gram);
    if (ce != cudaSuccess) return fail(TCC_ERR_CUDA, cudaGetErrorString(ce));
    CUDA_TRY(cudaGetLastError());
    // the kernel fills the 64 x 64 blocks I <= J: mirror the rest (the Gram matrix is symmetric)
    for (int i = 0; i < n2; ++i)
        for (int j = 0; j < i; ++j)
            if ((i >> 6) > (j >> 6)) gram[size_t(i) * n2 + j] = gram[size_t(j) * n2 + i];
    // <E_ps> = sum_q <E_ps E_qq> / n_elec = sum_q Gram[(s,p), (q,q)] / n_elec
    std::vector<double> e1(n2, 0.0);
    for (int pp = 0; pp < n; ++pp)
        for (int ss = 0; ss < n; ++ss) {
            double acc = 0.0;
            for (int q = 0; q < n; ++q) acc += gram[size_t(ss * n + pp) * n2 + q * n + q];
            e1[pp * n + ss] = hp.n_elec > 0 ? acc / double(hp.n_elec) : 0.0;
        }
    if (rdm1_host)
        for (int pp = 0; pp < n; ++pp)
            for (int q = 0; q < n; ++q) rdm1_host[pp * n + q] = e1[q * n + pp];  // dm1[p,q] = <q+ p>
    if (rdm2_host)
        for (int pp = 0; pp < n; ++pp)
            for (int q = 0; q < n; ++q)
                for (int r = 0; r < n; ++r)
                    for (int ss = 0; ss < n; ++ss)
                        rdm2_host[((size_t(pp) * n + q) * n + r) * n + ss] =
                            gram[size_t(q * n + pp) * n2 + r * n + ss] - (q == r ? e1[pp * n + ss] : 0.0);
    return 0;
}

// pUCCD density-matrix elements on the hard-core-boson CI vector (tencirchem/static/puccd.py:148-203): occ[p] =
// <n_p>, hop[p,q] = <b+_p b_q> and both[p,q] = <n_p n_q> for q < p (the other triangle is mirrored).  c_dev: the hcb CI
// vector in reference order on the device.
int tcc_make_rdm12_hcb(tcc_plan* p, const double* c_dev, double* occ_host, double* hop_host, double* both_host, void* stream) {
    if (!p || !c_dev || !occ_host || !hop_host || !both_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    HostPlan& hp = p->hp;
    if (!hp.hcb) return fail(TCC_ERR_BAD_ARGUMENT, "tcc_make_rdm12_hcb is for hard-core-boson plans (use tcc_make_rdm12)");
    cudaStream_t st = cudaStream_t(stream);
    if (int rc = ensure_hcb_tables(p)) return rc;
    const double* src = c_dev;
    if (!hp.identity_layout) {
        if (int rc = ensure_vec(p, &p->ket)) return rc;
        to_storage(p, c_dev, p->ket, st);
        src = p->ket;
    }
    const int n = hp.n;
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * 2 * size_t(n) * n));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(double) * 2 * size_t(n) * n, st));
    launch_hcb_rdm(src, hp.nstr, p->d_strings, n, p->d_binom, n + 2, p->d_bitpos, d_out, d_out + size_t(n) * n, st);
    p->launches += 1;
    std::vector<double> out(2 * size_t(n) * n);
    cudaError_t ce = cudaMemcpyAsync(out.data(), d_out, sizeof(double) * out.size(), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(d_out);
    if (ce != cudaSuccess) return fail(TCC_ERR_CUDA, cudaGetErrorString(ce));
    CUDA_TRY(cudaGetLastError());
    for (int pp = 0; pp < n; ++pp)
        for (int q = 0; q < n; ++q) {
            const int hi = std::max(pp, q), lo = std::min(pp, q);
            hop_host[pp * n + q] = pp == q ? 0.0 : out[size_t(hi) * n + lo];
            both_host[pp * n + q] = out[size_t(n) * n + size_t(hi) * n + lo];
        }
    for (int pp = 0; pp < n; ++pp) occ_host[pp] = out[size_t(n) * n + size_t(pp) * n + pp];
    return 0;
}

int tcc_profile_enable(tcc_plan* p, int on) {
    if (!p) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    p->prof.on = on != 0;
    p->prof.reset();
    return 0;
}

// ---- alpha-row sharded plans --------------------------------------------------------------------
// Every rank builds the same schedule; its alpha side holds only the rows the rank owns in the layout of each
// batch, addressed by local row, so the batch kernels run unchanged on the local [rows, nb] matrix.  Between
// segments (runs of batches with one layout) the caller moves rows between ranks (all-to-all).
int tcc_shard_info(const tcc_plan* p, int32_t* rank, int32_t* world, int32_t* n_layouts, int32_t* n_segments,
                   int64_t* max_local_rows) {
    if (!p) return fail(TCC_ERR_BAD_ARGUMENT, "null plan");
    if (rank) *rank = p->hp.shard_rank;
    if (world) *world = p->hp.shard_world;
    if (n_layouts) *n_layouts = int32_t(p->hp.layouts.size());
    if (n_segments) *n_segments = int32_t(p->hp.segments.size());
    if (max_local_rows) *max_local_rows = p->hp.max_local_rows;
    return 0;
}

int tcc_shard_segment(const tcc_plan* p, int seg, int32_t* layout, int32_t* batch_lo, int32_t* batch_hi) {
    if (!p || seg < 0 || seg >= int(p->hp.segments.size())) return fail(TCC_ERR_BAD_ARGUMENT, "segment index out of range");
    const Segment& s = p->hp.segments[seg];
    if (layout) *layout = s.layout;
    if (batch_lo) *batch_lo = s.batch_lo;
    if (batch_hi) *batch_hi = s.batch_hi;
    return 0;
}

int tcc_shard_layout(const tcc_plan* p, int layout, int32_t* owner_host /* [num_strings], storage order */,
                     uint32_t* orbital_mask) {
    if (!p || layout < 0 || layout >= int(p->hp.layouts.size())) return fail(TCC_ERR_BAD_ARGUMENT, "layout index out of range");
    const RowLayout& L = p->hp.layouts[layout];
    if (owner_host)
        for (int64_t a = 0; a < p->hp.nstr; ++a) owner_host[a] = L.owner[a];
    if (orbital_mask) *orbital_mask = L.T;
    return 0;
}

int tcc_shard_set_params(tcc_plan* p, const double* params_host, void* stream) {
    if (!p || (!params_host && p->hp.n_params > 0)) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc = prepare_rot(p, params_host, cudaStream_t(stream))) return rc;
    CUDA_TRY(cudaMemsetAsync(p->grad_ops, 0, sizeof(double) * (p->hp.ops.size() + 8), cudaStream_t(stream)));
    return 0;
}

int tcc_shard_forward(tcc_plan* p, int seg, double* vec_local_dev, void* stream) {
    if (!p || !vec_local_dev || seg < 0 || seg >= int(p->hp.segments.size())) return fail(TCC_ERR_BAD_ARGUMENT, "bad argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    cudaStream_t st = cudaStream_t(stream);
    const Segment& s = p->hp.segments[seg];
    p->prof.mark(TCC_PROF_FWD_BATCH, st);
    for (int b = s.batch_lo; b < s.batch_hi; ++b) {
        BatchDevSet& db = p->dbatches[b];
        launch_batch_forward(vec_local_dev, nullptr, p->hp.nb, db.fwd, p->d_rot_fwd, db.rects, p->batch_threads, 1, SetStrides{0, 0, 0, 0}, st,
                             &p->fan, p->persist_ctl, p->sms, p->persist_min_waves);
        p->prof.count(TCC_PROF_FWD_BATCH, 1, 16 * int64_t(p->hp.layouts[s.layout].rows[p->hp.shard_rank]) * p->hp.nb);
        p->launches += 1;
    }
    p->prof.mark(-1, st);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// row_stride: distance in doubles between consecutive local rows of ket (and of bra); nb for separate [rows, nb]
// matrices, 2 nb when ket and bra rows are interleaved in one [rows, 2, nb] block (one row move serves both)
int tcc_shard_reverse_strided(tcc_plan* p, int seg, double* ket_local_dev, double* bra_local_dev, int64_t row_stride, void* stream) {
    if (!p || !ket_local_dev || !bra_local_dev || seg < 0 || seg >= int(p->hp.segments.size()))
        return fail(TCC_ERR_BAD_ARGUMENT, "bad argument");
    if (row_stride < p->hp.nb) return fail(TCC_ERR_BAD_ARGUMENT, "row stride smaller than a row");
    if (int rc_dev = use_device(p)) return rc_dev;
    cudaStream_t st = cudaStream_t(stream);
    const Segment& s = p->hp.segments[seg];
    p->prof.mark(TCC_PROF_REV_BATCH, st);
    for (int b = s.batch_hi - 1; b >= s.batch_lo; --b) {
        BatchDevSet& db = p->dbatches[b];
        launch_batch_reverse(ket_local_dev, bra_local_dev, row_stride, db.rev, p->d_rot_rev, db.rects_rev, p->batch_threads_rev, p->bpart,
                             p->grad_ops, 1, SetStrides{0, 0, 0, 0}, st, &p->fan);
        p->prof.count(TCC_PROF_REV_BATCH, 1, 32 * int64_t(p->hp.layouts[s.layout].rows[p->hp.shard_rank]) * p->hp.nb);
        p->launches += 2;
    }
    p->prof.mark(-1, st);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// reverse sweep of a segment on ONE block of interleaved elements: kb_local = [local rows][nb][2] with (ket, bra) of a
// determinant next to each other (16-byte elements: one row move, one gather per element, one shared-memory access per
// pair end).  Returns TCC_ERR_BAD_ARGUMENT when a batch of the segment has no lean streams (use the two-vector call).
int tcc_shard_reverse_interleaved(tcc_plan* p, int seg, double* kb_local_dev, void* stream) {
    if (!p || !kb_local_dev || seg < 0 || seg >= int(p->hp.segments.size())) return fail(TCC_ERR_BAD_ARGUMENT, "bad argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    cudaStream_t st = cudaStream_t(stream);
    const Segment& s = p->hp.segments[seg];
    for (int b = s.batch_lo; b < s.batch_hi; ++b)
        if (!p->dbatches[b].rev.lean) return fail(TCC_ERR_BAD_ARGUMENT, "segment holds a batch without lean streams");
    cudaMemsetAsync(p->persist_ctl, 0, sizeof(unsigned) * 64, st);
    p->prof.mark(TCC_PROF_REV_BATCH, st);
    for (int b = s.batch_hi - 1; b >= s.batch_lo; --b) {
        BatchDevSet& db = p->dbatches[b];
        launch_batch_reverse_il(kb_local_dev, nullptr, p->hp.nb, db.rev, p->d_rot_rev, db.rects_rev, p->batch_threads_rev, p->bpart,
                                p->grad_ops, 1, SetStrides{0, 0, 0, 0}, st, &p->fan, p->persist_ctl, p->sms, p->persist_min_waves);
        p->prof.count(TCC_PROF_REV_BATCH, 1, 32 * int64_t(p->hp.layouts[s.layout].rows[p->hp.shard_rank]) * p->hp.nb);
        p->launches += 2;
    }
    p->prof.mark(-1, st);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// 1 when every batch of the plan can run the interleaved reverse sweep (tcc_shard_reverse_interleaved)
int tcc_shard_can_interleave(const tcc_plan* p) {
    if (!p || !p->rev_interleave || p->device < 0) return 0;
    for (const auto& db : p->dbatches)
        if (!db.rev.lean) return 0;
    return 1;
}

int tcc_shard_reverse(tcc_plan* p, int seg, double* ket_local_dev, double* bra_local_dev, void* stream) {
    return tcc_shard_reverse_strided(p, seg, ket_local_dev, bra_local_dev, p ? p->hp.nb : 0, stream);
}

// this rank's share of the gradient (sum over its tiles), by parameter, with the factor 2 of evolve_civector.py:218;
// the caller adds the shares of all ranks
int tcc_shard_gradient(tcc_plan* p, double* grad_host, void* stream) {
    if (!p || !grad_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    cudaStream_t st = cudaStream_t(stream);
    const HostPlan& hp = p->hp;
    const int m = int(hp.ops.size());
    CUDA_TRY(cudaMemcpyAsync(p->pinned, p->grad_ops, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->prof.resolve();
    for (int i = 0; i < hp.n_params; ++i) grad_host[i] = 0.0;
    for (int j = 0; j < m; ++j) grad_host[hp.ops[j].param] += p->pinned[j];
    for (int i = 0; i < hp.n_params; ++i) grad_host[i] *= 2.0;
    return 0;
}

// Contributions to sigma = H c of the storage rows [row_lo, row_lo + row_cnt): the same-spin terms OF these rows and
// the opposite-spin terms scattered FROM these rows, ADDED into sigma_full_dev (the caller zeroes it).  c_full_dev and
// sigma_full_dev are full [na, nb] storage-order matrices; the sum over a partition of the rows is H c.
int tcc_apply_h_rows(tcc_plan* p, const double* c_full_dev, double* sigma_full_dev, int64_t row_lo, int64_t row_cnt, void* stream) {
    if (!p || !c_full_dev || !sigma_full_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    HostPlan& hp = p->hp;
    if (hp.hcb) return fail(TCC_ERR_BAD_ARGUMENT, "hcb plans have a single row");
    if (!p->has_h) return fail(TCC_ERR_NO_HAMILTONIAN, "tcc_set_hamiltonian has not been called");
    if (row_lo < 0 || row_cnt < 0 || row_lo + row_cnt > hp.na) return fail(TCC_ERR_BAD_ARGUMENT, "row range out of bounds");
    if (row_cnt == 0) return 0;
    cudaStream_t st = cudaStream_t(stream);
    const size_t need = size_t(row_cnt) * size_t(hp.nb);
    if (p->rows_ws_cap < need) {
        cudaFree(p->rows_ws1);
        cudaFree(p->rows_ws2);
        p->rows_ws1 = p->rows_ws2 = nullptr;
        p->rows_ws_cap = 0;
        CUDA_TRY(dev_alloc(p, &p->rows_ws1, need));
        CUDA_TRY(dev_alloc(p, &p->rows_ws2, need));
        p->rows_ws_cap = need;
    }
    const double* c_rows = c_full_dev + row_lo * hp.nb;
    double* s_rows = sigma_full_dev + row_lo * hp.nb;
    p->prof.mark(TCC_PROF_SIGMA_SS, st);
    // alpha-alpha: output rows of the range gather from all rows of c
    launch_spmm_rows(c_full_dev, s_rows, row_cnt, hp.nb, p->d_ell_col + row_lo * hp.ell_width, p->d_ell_val + row_lo * hp.ell_width,
                     hp.ell_width, true, 1, st);
    // beta-beta acts inside each row: transpose the row block, contract, transpose back
    launch_transpose(c_rows, p->rows_ws1, row_cnt, hp.nb, false, 1, st);
    launch_spmm_rows(p->rows_ws1, p->rows_ws2, hp.nb, row_cnt, p->d_ell_col, p->d_ell_val, hp.ell_width, false, 1, st);
    launch_transpose(p->rows_ws2, s_rows, hp.nb, row_cnt, true, 1, st);
    p->prof.count(TCC_PROF_SIGMA_SS, 4, 9 * 8 * int64_t(need));
    p->prof.mark(TCC_PROF_SIGMA_AB, st);
    launch_sigma_ab(c_full_dev, sigma_full_dev, hp.na, hp.nb, p->d_links, hp.ml, p->d_w, hp.n, p->n2p, p->d_kmap, 1, st, row_lo, row_cnt);
    p->prof.count(TCC_PROF_SIGMA_AB, (row_cnt + 65534) / 65535, 3 * 8 * int64_t(need));
    p->prof.mark(-1, st);
    p->launches += 4 + (row_cnt + 65534) / 65535;
    CUDA_TRY(cudaGetLastError());
    return 0;
}


// ---- sharded sigma = H c: the vector stays distributed ---------------------------------------------
// beta-beta terms act inside a row: any layout.  alpha-beta terms: three passes, pass i in layout sigma_layout[i]
// with the alpha replacements E_pq that do not touch that layout's orbitals T (they keep a row on its rank); three
// layouts with disjoint T cover every pair.  alpha-alpha terms: on a column block of ALL rows (the caller transposes).
int tcc_shard_sigma_layouts(const tcc_plan* p, int32_t* layouts3) {
    if (!p || !layouts3) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    for (int i = 0; i < 3; ++i) layouts3[i] = p->hp.sigma_layout[i];
    return 0;
}

static int ensure_rows_ws(tcc_plan* p, size_t need) {
    if (p->rows_ws_cap >= need) return 0;
    cudaFree(p->rows_ws1);
    cudaFree(p->rows_ws2);
    p->rows_ws1 = p->rows_ws2 = nullptr;
    p->rows_ws_cap = 0;
    CUDA_TRY(dev_alloc(p, &p->rows_ws1, need));
    CUDA_TRY(dev_alloc(p, &p->rows_ws2, need));
    p->rows_ws_cap = need;
    return 0;
}

// sigma_local += (beta-beta part of H) c_local for `rows` local rows of any layout
int tcc_shard_sigma_bb(tcc_plan* p, const double* c_local_dev, double* sigma_local_dev, int64_t rows, void* stream) {
    if (!p || !c_local_dev || !sigma_local_dev || rows < 0) return fail(TCC_ERR_BAD_ARGUMENT, "bad argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (!p->has_h) return fail(TCC_ERR_NO_HAMILTONIAN, "tcc_set_hamiltonian has not been called");
    HostPlan& hp = p->hp;
    if (rows == 0) return 0;
    cudaStream_t st = cudaStream_t(stream);
    // row chunks: the two transposed scratch blocks stay below ~2 GB each whatever the shard size
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(rows, int64_t(env_int("TCC_B200_BB_CHUNK_MB", 2048)) * (1 << 17) / hp.nb));
    if (int rc = ensure_rows_ws(p, size_t(chunk) * size_t(hp.nb))) return rc;
    p->prof.mark(TCC_PROF_SIGMA_SS, st);
    for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
        const int64_t nr = std::min(chunk, rows - r0);
        launch_transpose(c_local_dev + r0 * hp.nb, p->rows_ws1, nr, hp.nb, false, 1, st);
        launch_spmm_rows(p->rows_ws1, p->rows_ws2, hp.nb, nr, p->d_ell_col, p->d_ell_val, hp.ell_width, false, 1, st);
        launch_transpose(p->rows_ws2, sigma_local_dev + r0 * hp.nb, hp.nb, nr, true, 1, st);
        p->launches += 3;
    }
    p->prof.count(TCC_PROF_SIGMA_SS, 3 * ((rows + chunk - 1) / chunk), 7 * 8 * rows * hp.nb);
    p->prof.mark(-1, st);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// sigma_cols = (alpha-alpha part of H) c_cols on a block of `ncols` columns of ALL alpha rows ([na, ncols], storage rows)
int tcc_shard_sigma_aa(tcc_plan* p, const double* c_cols_dev, double* sigma_cols_dev, int64_t ncols, void* stream) {
    if (!p || !c_cols_dev || !sigma_cols_dev || ncols < 0) return fail(TCC_ERR_BAD_ARGUMENT, "bad argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (!p->has_h) return fail(TCC_ERR_NO_HAMILTONIAN, "tcc_set_hamiltonian has not been called");
    HostPlan& hp = p->hp;
    if (ncols == 0) return 0;
    cudaStream_t st = cudaStream_t(stream);
    p->prof.mark(TCC_PROF_SIGMA_SS, st);
    launch_spmm_rows(c_cols_dev, sigma_cols_dev, hp.na, ncols, p->d_ell_col, p->d_ell_val, hp.ell_width, false, 1, st);
    p->prof.count(TCC_PROF_SIGMA_SS, 1, 2 * 8 * hp.na * ncols);
    p->prof.mark(-1, st);
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// sigma_local += pass `pass` of the alpha-beta part; c_local / sigma_local: the rank's rows in layout sigma_layout[pass]
int tcc_shard_sigma_ab(tcc_plan* p, int pass, const double* c_local_dev, double* sigma_local_dev, void* stream) {
    if (!p || pass < 0 || pass > 2 || !c_local_dev || !sigma_local_dev) return fail(TCC_ERR_BAD_ARGUMENT, "bad argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (!p->has_h) return fail(TCC_ERR_NO_HAMILTONIAN, "tcc_set_hamiltonian has not been called");
    HostPlan& hp = p->hp;
    if (hp.shard_world == 1) return fail(TCC_ERR_BAD_ARGUMENT, "not a sharded plan");
    const int64_t rows = hp.layouts[hp.sigma_layout[pass]].rows[hp.shard_rank];
    if (rows == 0 || hp.sigma_aml[pass] == 0) return 0;
    cudaStream_t st = cudaStream_t(stream);
    p->prof.mark(TCC_PROF_SIGMA_AB, st);
    launch_sigma_ab(c_local_dev, sigma_local_dev, rows, hp.nb, p->d_links, hp.ml, p->d_w, hp.n, p->n2p, p->d_kmap, 1, st, 0, rows,
                    p->d_sigma_links[pass], hp.sigma_aml[pass]);
    p->prof.count(TCC_PROF_SIGMA_AB, (rows + 65534) / 65535, 3 * 8 * rows * hp.nb);
    p->prof.mark(-1, st);
    p->launches += (rows + 65534) / 65535;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Row move over peer memory (no plan needed): see push_rows_kernel.  All pointers are device pointers of `device`;
// peer_base_dev[q] = base address of rank q's destination block as mapped into this process.
int tcc_push_rows(int device, const double* src_dev, int64_t row_doubles, int64_t n_rows, const int64_t* src_row_dev,
                  const int64_t* dst_row_dev, const int32_t* dst_rank_dev, const uint64_t* peer_base_dev, int64_t dst_offset_doubles,
                  void* stream) {
    if (n_rows < 0 || row_doubles < 0 || dst_offset_doubles < 0) return fail(TCC_ERR_BAD_ARGUMENT, "negative size");
    if (n_rows == 0 || row_doubles == 0) return 0;
    if (!src_dev || !src_row_dev || !dst_row_dev || !dst_rank_dev || !peer_base_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(device));
    launch_push_rows(src_dev, row_doubles, n_rows, src_row_dev, dst_row_dev, dst_rank_dev, peer_base_dev, dst_offset_doubles,
                     cudaStream_t(stream));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int tcc_profile_batches(tcc_plan* p, int on, float* fwd_ms, float* rev_ms) {
    if (!p) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (fwd_ms && rev_ms && p->btimer.on) {
        // timings of the LAST energy_and_grad call (the stream was synchronised by it)
        const size_t nbt = p->dbatches.size();
        if (p->btimer.ev.size() >= 2 * nbt + 2) {
            for (size_t b = 0; b < nbt; ++b) {
                float t = 0.f;
                cudaEventElapsedTime(&t, p->btimer.ev[b], p->btimer.ev[b + 1]);
                fwd_ms[b] = t;
                t = 0.f;
                cudaEventElapsedTime(&t, p->btimer.ev[nbt + 1 + (nbt - 1 - b)], p->btimer.ev[nbt + 2 + (nbt - 1 - b)]);
                rev_ms[b] = t;
            }
        }
        cudaGetLastError();
    }
    p->btimer.on = on != 0;
    return 0;
}

int tcc_profile_read(tcc_plan* p, double* ms, int64_t* launches, int64_t* bytes) {
    if (!p || !ms || !launches || !bytes) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    for (int i = 0; i < TCC_PROF_SLOTS; ++i) {
        ms[i] = p->prof.ms[i];
        launches[i] = p->prof.launches[i];
        bytes[i] = p->prof.bytes[i];
    }
    return 0;
}

// ---- host-buffer conveniences ---------------------------------------------------------------
int tcc_civector_host(tcc_plan* p, const double* params_host, const double* init_host, double* out_host) {
    if (!p || !out_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    if (int rc = ensure_vec(p, &p->io1)) return rc;
    if (int rc = ensure_vec(p, &p->io2)) return rc;
    const size_t bytes = p->hp.N * sizeof(double);
    if (init_host) CUDA_TRY(cudaMemcpy(p->io1, init_host, bytes, cudaMemcpyHostToDevice));
    int rc = tcc_civector(p, params_host, init_host ? p->io1 : nullptr, p->io2, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(out_host, p->io2, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int tcc_apply_h_host(tcc_plan* p, const double* c_host, double* sigma_host) {
    if (!p || !c_host || !sigma_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc = ensure_vec(p, &p->io1)) return rc;
    if (int rc = ensure_vec(p, &p->io2)) return rc;
    const size_t bytes = p->hp.N * sizeof(double);
    CUDA_TRY(cudaMemcpy(p->io1, c_host, bytes, cudaMemcpyHostToDevice));
    int rc = tcc_apply_h(p, p->io1, p->io2, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(sigma_host, p->io2, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int tcc_energy_and_grad_host(tcc_plan* p, const double* params_host, const double* init_host, double* energy_host,
                             double* grad_host) {
    if (!p || !energy_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc_sh = whole_vector_only(p)) return rc_sh;
    const double* init_dev = nullptr;
    if (init_host) {
        if (int rc = ensure_vec(p, &p->io1)) return rc;
        CUDA_TRY(cudaMemcpy(p->io1, init_host, p->hp.N * sizeof(double), cudaMemcpyHostToDevice));
        init_dev = p->io1;
    }
    if (grad_host) return tcc_energy_and_grad(p, params_host, init_dev, energy_host, grad_host, nullptr);
    return tcc_energy(p, params_host, init_dev, energy_host, nullptr);
}

int tcc_apply_excitation_host(tcc_plan* p, const double* c_host, const int32_t* ex_op, int ex_op_len, double* out_host) {
    if (!p || !c_host || !out_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc = ensure_vec(p, &p->io1)) return rc;
    if (int rc = ensure_vec(p, &p->io2)) return rc;
    const size_t bytes = p->hp.N * sizeof(double);
    CUDA_TRY(cudaMemcpy(p->io1, c_host, bytes, cudaMemcpyHostToDevice));
    int rc = tcc_apply_excitation(p, p->io1, ex_op, ex_op_len, p->io2, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(out_host, p->io2, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
