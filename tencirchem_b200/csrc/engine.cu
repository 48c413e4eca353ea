// C ABI of the B200 civector engine (include/tcc_b200.h): device plan, drivers for get_civector /
// energy_and_grad / apply_excitation / H.c, host-buffer conveniences.
//
// Reference call sites this replaces (paths relative to /root/reference):
//   tencirchem/static/evolve_civector.py:180-198  get_civector                  -> tcc_civector
//   tencirchem/static/evolve_civector.py:201-243  get_energy_and_grad_civector  -> tcc_energy_and_grad
//   tencirchem/static/evolve_civector.py:311-313  apply_excitation_civector     -> tcc_apply_excitation
//   tencirchem/static/hamiltonian.py:146-183      fci_func (H.c)                -> tcc_set_hamiltonian / tcc_apply_h
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tcc_b200.h"
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

struct tcc_plan {
    HostPlan hp;
    Profiler prof;
    int device = 0;
    std::vector<DevHop> dhops;
    double* ket = nullptr;
    double* bra = nullptr;
    double* ws1 = nullptr;
    double* ws2 = nullptr;
    double* partials = nullptr;
    unsigned* ticket = nullptr;
    double* grad_ops = nullptr;  // [n_ops] on the device
    double* scalars = nullptr;   // small device scratch
    double* pinned = nullptr;    // pinned host staging [n_ops + 8]
    bool has_h = false;
    Link* d_links = nullptr;
    double* d_w = nullptr;
    int n2p = 0;
    uint32_t* d_ell_col = nullptr;
    double* d_ell_val = nullptr;
    uint32_t* d_strings = nullptr;
    double* d_diag1 = nullptr;
    double* d_hop = nullptr;
    double* d_diag2 = nullptr;
    int64_t* d_binom = nullptr;
    int64_t launches = 0;
    int64_t dev_bytes = 0;
};

template <typename T>
static cudaError_t dev_alloc(tcc_plan* p, T** ptr, size_t count) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(ptr), count * sizeof(T) + 16);
    if (e == cudaSuccess) p->dev_bytes += int64_t(count * sizeof(T));
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

extern "C" {

const char* tcc_last_error(void) { return g_error.c_str(); }
const char* tcc_version(void) { return "tencirchem_b200 0.1 (sm_100a)"; }

int tcc_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(TCC_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return n;
}

int tcc_plan_create(int n_qubits, int n_elec, int hcb, int n_ops, const int32_t* ex_ops_flat, const int32_t* ex_op_len,
                    const int32_t* param_ids, int device, tcc_plan** out) {
    if (!out) return fail(TCC_ERR_BAD_ARGUMENT, "out is null");
    *out = nullptr;
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
        rc = p->hp.make_op(ex_ops_flat + off, len, pid, &p->hp.ops[j]);
        if (rc) {
            g_error = p->hp.error;
            delete p;
            return rc;
        }
        off += len;
        if (pid > max_param) max_param = pid;
    }
    p->hp.n_params = max_param + 1;
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
    for (size_t h = 0; h < p->hp.hops.size(); ++h)
        if (upload_hop(p, int(h))) return bail(TCC_ERR_CUDA);
    if (cudaMalloc(&p->partials, sizeof(double) * reverse_partials_capacity()) != cudaSuccess ||
        cudaMalloc(&p->ticket, sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc(&p->grad_ops, sizeof(double) * (n_ops + 8)) != cudaSuccess ||
        cudaMalloc(&p->scalars, sizeof(double) * 8) != cudaSuccess ||
        cudaMallocHost(&p->pinned, sizeof(double) * (n_ops + 8)) != cudaSuccess) {
        g_error = std::string("plan workspace allocation: ") + cudaGetErrorString(cudaGetLastError());
        return bail(TCC_ERR_CUDA);
    }
    cudaMemset(p->ticket, 0, sizeof(unsigned));
    cudaMemset(p->grad_ops, 0, sizeof(double) * (n_ops + 8));
    *out = p;
    return 0;
}

void tcc_plan_destroy(tcc_plan* p) {
    if (!p) return;
    if (p->device < 0) {
        delete p;
        return;
    }
    cudaSetDevice(p->device);
    for (auto& d : p->dhops) {
        cudaFree(d.src);
        cudaFree(d.dst);
    }
    cudaFree(p->ket);
    cudaFree(p->bra);
    cudaFree(p->ws1);
    cudaFree(p->ws2);
    cudaFree(p->partials);
    cudaFree(p->ticket);
    cudaFree(p->grad_ops);
    cudaFree(p->scalars);
    if (p->pinned) cudaFreeHost(p->pinned);
    cudaFree(p->d_links);
    cudaFree(p->d_w);
    cudaFree(p->d_ell_col);
    cudaFree(p->d_ell_val);
    cudaFree(p->d_strings);
    cudaFree(p->d_diag1);
    cudaFree(p->d_hop);
    cudaFree(p->d_diag2);
    cudaFree(p->d_binom);
    delete p;
}

int64_t tcc_civector_size(const tcc_plan* p) { return p ? p->hp.N : 0; }
int64_t tcc_num_strings(const tcc_plan* p) { return p ? p->hp.nstr : 0; }
int tcc_num_ops(const tcc_plan* p) { return p ? int(p->hp.ops.size()) : 0; }
int tcc_num_params(const tcc_plan* p) { return p ? p->hp.n_params : 0; }
int64_t tcc_plan_device_bytes(const tcc_plan* p) { return p ? p->dev_bytes : 0; }
int64_t tcc_launch_count(const tcc_plan* p) { return p ? p->launches : 0; }

int64_t tcc_op_support_bytes(const tcc_plan* p, int j) {
    if (!p || j < 0 || j >= int(p->hp.ops.size())) return 0;
    return 32 * p->hp.support_pairs(p->hp.ops[j]);  // 2 elements per pair, each read and written once (8 B)
}

int tcc_ci_strings_spin(const tcc_plan* p, uint64_t* out_host) {
    if (!p || !out_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    for (int64_t i = 0; i < p->hp.nstr; ++i) out_host[i] = p->hp.strings[i];
    return 0;
}

int tcc_excitation_tables(const tcc_plan* p, int j, uint64_t* perm_host, int8_t* phase_host, int8_t* f2_host) {
    if (!p || !perm_host || !phase_host || !f2_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (j < 0 || j >= int(p->hp.ops.size())) return fail(TCC_ERR_BAD_ARGUMENT, "excitation index out of range");
    p->hp.expand_tables(p->hp.ops[j], perm_host, phase_host, f2_host);
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
        std::vector<int64_t> bin(size_t(n + 2) * (n + 2));
        for (int i = 0; i < n + 2; ++i)
            for (int j = 0; j < n + 2; ++j) bin[size_t(i) * (n + 2) + j] = hp.binom[i][j];
        if (!p->d_strings) {
            CUDA_TRY(dev_alloc(p, &p->d_strings, size_t(hp.nstr)));
            CUDA_TRY(dev_alloc(p, &p->d_diag1, size_t(n)));
            CUDA_TRY(dev_alloc(p, &p->d_hop, size_t(n) * n));
            CUDA_TRY(dev_alloc(p, &p->d_diag2, size_t(n) * n));
            CUDA_TRY(dev_alloc(p, &p->d_binom, bin.size()));
        }
        CUDA_TRY(cudaMemcpy(p->d_strings, hp.strings.data(), hp.nstr * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->d_diag1, d1.data(), n * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->d_hop, hop.data(), hop.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->d_diag2, d2.data(), d2.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->d_binom, bin.data(), bin.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        p->has_h = true;
        return 0;
    }
    std::vector<double> h2e = absorb_h1e(int1e, int2e, n, hp.n_elec, 0.5);
    if (hp.links.empty()) hp.build_links();
    hp.build_same_spin_ell(h2e);
    const int n2 = n * n;
    p->n2p = (n2 + 3) / 4 * 4;
    std::vector<double> w(size_t(n2) * p->n2p, 0.0);
    for (int pq = 0; pq < n2; ++pq)
        for (int rs = 0; rs < n2; ++rs) w[size_t(pq) * p->n2p + rs] = h2e[size_t(pq) * n2 + rs] + h2e[size_t(rs) * n2 + pq];
    cudaFree(p->d_w);
    cudaFree(p->d_ell_col);
    cudaFree(p->d_ell_val);
    p->d_w = nullptr;
    p->d_ell_col = nullptr;
    p->d_ell_val = nullptr;
    if (!p->d_links) {
        CUDA_TRY(dev_alloc(p, &p->d_links, hp.links.size()));
        CUDA_TRY(cudaMemcpy(p->d_links, hp.links.data(), hp.links.size() * sizeof(Link), cudaMemcpyHostToDevice));
    }
    CUDA_TRY(dev_alloc(p, &p->d_w, w.size()));
    CUDA_TRY(cudaMemcpy(p->d_w, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
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

static int apply_h_impl(tcc_plan* p, const double* c, double* sigma, cudaStream_t st) {
    HostPlan& hp = p->hp;
    if (!p->has_h) return fail(TCC_ERR_NO_HAMILTONIAN, "tcc_set_hamiltonian has not been called");
    if (hp.hcb) {
        launch_sigma_hcb(c, sigma, hp.nstr, p->d_strings, hp.n, hp.k, p->d_diag1, p->d_hop, p->d_diag2, p->d_binom, hp.n + 2, st);
        p->launches += 1;
        return 0;
    }
    if (int rc = ensure_vec(p, &p->ws1)) return rc;
    if (int rc = ensure_vec(p, &p->ws2)) return rc;
    const int64_t vec = hp.N * 8;
    // alpha-alpha: sigma(a', :) = sum_a Hss[a', a] c(a, :)
    p->prof.mark(TCC_PROF_SIGMA_SS, st);
    launch_spmm_rows(c, sigma, hp.na, hp.nb, p->d_ell_col, p->d_ell_val, hp.ell_width, false, st);
    // beta-beta on the transposed matrix, then added back transposed
    launch_transpose(c, p->ws1, hp.na, hp.nb, false, st);
    launch_spmm_rows(p->ws1, p->ws2, hp.nb, hp.na, p->d_ell_col, p->d_ell_val, hp.ell_width, false, st);
    launch_transpose(p->ws2, sigma, hp.nb, hp.na, true, st);
    p->prof.count(TCC_PROF_SIGMA_SS, 4, 9 * vec);  // 2+2+2+3 vector passes
    // alpha-beta on the FP64 tensor path
    p->prof.mark(TCC_PROF_SIGMA_AB, st);
    launch_sigma_ab(c, sigma, hp.na, hp.nb, p->d_links, hp.ml, p->d_w, hp.n, p->n2p, st);
    p->prof.count(TCC_PROF_SIGMA_AB, (hp.na + 65534) / 65535, 3 * vec);
    p->prof.mark(-1, st);
    p->launches += 4 + (hp.na + 65534) / 65535;
    return 0;
}

int tcc_apply_h(tcc_plan* p, const double* c_dev, double* sigma_dev, void* stream) {
    if (!p || !c_dev || !sigma_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (c_dev == sigma_dev) return fail(TCC_ERR_BAD_ARGUMENT, "c and sigma must not alias");
    if (int rc_dev = use_device(p)) return rc_dev;
    int rc = apply_h_impl(p, c_dev, sigma_dev, cudaStream_t(stream));
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static void rotate_op(tcc_plan* p, int j, double theta, double* c, cudaStream_t st) {
    const OpDesc& op = p->hp.ops[j];
    p->prof.mark(TCC_PROF_FWD_BETA + op.kind, st);
    launch_rotate(c, p->hp.nb, row_side(p, op), col_side(p, op), std::cos(theta), std::sin(theta) * op.s0, st);
    p->prof.count(TCC_PROF_FWD_BETA + op.kind, 1, 32 * p->hp.support_pairs(op));
    p->launches += 1;
}

static int civector_impl(tcc_plan* p, const double* params, const double* init_dev, double* out, cudaStream_t st) {
    HostPlan& hp = p->hp;
    if (init_dev) {
        if (init_dev != out) CUDA_TRY(cudaMemcpyAsync(out, init_dev, hp.N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    } else {
        launch_set_unit(out, hp.N, st);
        p->launches += 1;
    }
    for (size_t j = 0; j < hp.ops.size(); ++j) rotate_op(p, int(j), params[hp.ops[j].param], out, st);
    p->prof.mark(-1, st);
    return 0;
}

int tcc_civector(tcc_plan* p, const double* params_host, const double* init_dev, double* out_dev, void* stream) {
    if (!p || !out_dev || (!params_host && p->hp.n_params > 0)) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    int rc = civector_impl(p, params_host, init_dev, out_dev, cudaStream_t(stream));
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int tcc_rotate(tcc_plan* p, int j, double theta, double* c_dev, void* stream) {
    if (!p || !c_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (j < 0 || j >= int(p->hp.ops.size())) return fail(TCC_ERR_BAD_ARGUMENT, "excitation index out of range");
    if (int rc_dev = use_device(p)) return rc_dev;
    rotate_op(p, j, theta, c_dev, cudaStream_t(stream));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static void reverse_op(tcc_plan* p, int j, double theta, double* ket, double* bra, double* grad_out, cudaStream_t st) {
    const OpDesc& op = p->hp.ops[j];
    p->prof.mark(TCC_PROF_REV_BETA + op.kind, st);
    launch_reverse(ket, bra, p->hp.nb, row_side(p, op), col_side(p, op), std::cos(theta), -std::sin(theta) * op.s0,
                   double(op.s0), p->partials, p->ticket, grad_out, st);
    p->prof.count(TCC_PROF_REV_BETA + op.kind, 1, 64 * p->hp.support_pairs(op));
    p->launches += 1;
}

int tcc_reverse_step(tcc_plan* p, int j, double theta, double* ket_dev, double* bra_dev, double* grad_dev, void* stream) {
    if (!p || !ket_dev || !bra_dev || !grad_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (j < 0 || j >= int(p->hp.ops.size())) return fail(TCC_ERR_BAD_ARGUMENT, "excitation index out of range");
    if (int rc_dev = use_device(p)) return rc_dev;
    reverse_op(p, j, theta, ket_dev, bra_dev, grad_dev, cudaStream_t(stream));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int energy_impl(tcc_plan* p, const double* params, const double* init_dev, double* energy, double* grad, cudaStream_t st) {
    HostPlan& hp = p->hp;
    if (!p->has_h) return fail(TCC_ERR_NO_HAMILTONIAN, "tcc_set_hamiltonian has not been called");
    if (int rc = ensure_vec(p, &p->ket)) return rc;
    if (int rc = ensure_vec(p, &p->bra)) return rc;
    if (int rc = civector_impl(p, params, init_dev, p->ket, st)) return rc;
    if (int rc = apply_h_impl(p, p->ket, p->bra, st)) return rc;
    const int m = int(hp.ops.size());
    p->prof.mark(TCC_PROF_DOT, st);
    launch_dot(p->bra, p->ket, hp.N, p->partials, p->ticket, p->grad_ops + m, st);
    p->prof.count(TCC_PROF_DOT, 1, 16 * hp.N);
    p->prof.mark(-1, st);
    p->launches += 1;
    if (grad) {
        for (int j = m - 1; j >= 0; --j) reverse_op(p, j, params[hp.ops[j].param], p->ket, p->bra, p->grad_ops + j, st);
    }
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
    return energy_impl(p, params_host, init_dev, energy_host, nullptr, cudaStream_t(stream));
}

int tcc_energy_and_grad(tcc_plan* p, const double* params_host, const double* init_dev, double* energy_host,
                        double* grad_host, void* stream) {
    if (!p || !energy_host || !grad_host || (!params_host && p->hp.n_params > 0)) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    return energy_impl(p, params_host, init_dev, energy_host, grad_host, cudaStream_t(stream));
}

int tcc_apply_excitation(tcc_plan* p, const double* c_dev, const int32_t* ex_op, int ex_op_len, double* out_dev, void* stream) {
    if (!p || !c_dev || !ex_op || !out_dev) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (c_dev == out_dev) return fail(TCC_ERR_BAD_ARGUMENT, "input and output must not alias");
    if (int rc_dev = use_device(p)) return rc_dev;
    OpDesc op;
    int rc = p->hp.make_op(ex_op, ex_op_len, 0, &op);
    if (rc) return fail(rc, p->hp.error);
    if (op.hop_a >= 0)
        if (int r2 = upload_hop(p, op.hop_a)) return r2;
    if (op.hop_b >= 0)
        if (int r2 = upload_hop(p, op.hop_b)) return r2;
    cudaStream_t st = cudaStream_t(stream);
    CUDA_TRY(cudaMemsetAsync(out_dev, 0, p->hp.N * sizeof(double), st));
    launch_apply_g(c_dev, out_dev, p->hp.nb, row_side(p, op), col_side(p, op), double(op.s0), st);
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int tcc_profile_enable(tcc_plan* p, int on) {
    if (!p) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    p->prof.on = on != 0;
    p->prof.reset();
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
    if (int rc = ensure_vec(p, &p->ket)) return rc;
    const size_t bytes = p->hp.N * sizeof(double);
    if (init_host) CUDA_TRY(cudaMemcpy(p->ket, init_host, bytes, cudaMemcpyHostToDevice));
    int rc = tcc_civector(p, params_host, init_host ? p->ket : nullptr, p->ket, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(out_host, p->ket, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int tcc_apply_h_host(tcc_plan* p, const double* c_host, double* sigma_host) {
    if (!p || !c_host || !sigma_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc = ensure_vec(p, &p->ket)) return rc;
    if (int rc = ensure_vec(p, &p->bra)) return rc;
    const size_t bytes = p->hp.N * sizeof(double);
    CUDA_TRY(cudaMemcpy(p->ket, c_host, bytes, cudaMemcpyHostToDevice));
    int rc = tcc_apply_h(p, p->ket, p->bra, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(sigma_host, p->bra, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int tcc_energy_and_grad_host(tcc_plan* p, const double* params_host, const double* init_host, double* energy_host,
                             double* grad_host) {
    if (!p || !energy_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    const double* init_dev = nullptr;
    if (init_host) {
        if (int rc = ensure_vec(p, &p->ket)) return rc;
        CUDA_TRY(cudaMemcpy(p->ket, init_host, p->hp.N * sizeof(double), cudaMemcpyHostToDevice));
        init_dev = p->ket;
    }
    if (grad_host) return tcc_energy_and_grad(p, params_host, init_dev, energy_host, grad_host, nullptr);
    return tcc_energy(p, params_host, init_dev, energy_host, nullptr);
}

int tcc_apply_excitation_host(tcc_plan* p, const double* c_host, const int32_t* ex_op, int ex_op_len, double* out_host) {
    if (!p || !c_host || !out_host) return fail(TCC_ERR_BAD_ARGUMENT, "null argument");
    if (int rc_dev = use_device(p)) return rc_dev;
    if (int rc = ensure_vec(p, &p->ket)) return rc;
    if (int rc = ensure_vec(p, &p->bra)) return rc;
    const size_t bytes = p->hp.N * sizeof(double);
    CUDA_TRY(cudaMemcpy(p->ket, c_host, bytes, cudaMemcpyHostToDevice));
    int rc = tcc_apply_excitation(p, p->ket, ex_op, ex_op_len, p->bra, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(out_host, p->bra, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
