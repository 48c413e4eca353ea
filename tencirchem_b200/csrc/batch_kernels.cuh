// Device-side descriptors and launchers of the batched excitation kernels (batch_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "plan.hpp"

namespace tcc {

struct SideDev {
    const uint32_t* addr;
    const uint16_t* ord;
    const uint32_t* addr_sorted;
    const uint32_t* sigma;
    const PanelDesc* panels;
    const uint32_t* tab;
    const int32_t* tab_index;
    int m;
    int n_panels;
    // dynamic column layout (beta side, null without it): gather / scatter tables of this sweep direction
    const uint32_t* dyn_in_sorted;
    const uint16_t* dyn_in_ord;
    const uint32_t* dyn_out_sorted;
    const uint16_t* dyn_out_ord;
};

struct BatchDev {
    SideDev A, B;
    const BatchOp* ops;  // application order for this direction
    const uint32_t* plist;
    const int32_t* plist_index;  // indexed by the FORWARD position of the op in the batch
    const uint32_t* stream8;     // lean per-thread streams (Batch::stream8 / stream16), null without `lean`
    const uint32_t* stream16;
    const int32_t* stream_index;
    int lean;                    // the batch runs on lean_batch_kernel
    int nops;
    int ld_fixed;
    int reversed;  // ops[] is the forward list reversed
};

// element strides between consecutive parameter sets of a batched-theta evaluation (all 0 for a single set)
struct SetStrides {
    int64_t vec;   // CI vectors
    int64_t rot;   // (cos, sin) records
    int64_t part;  // per-tile gradient partials
    int64_t grad;  // per-excitation gradients
};

// a rectangle of the (alpha panel) x (beta panel) grid handled by one launch
struct TileRect {
    int pa_lo, pa_cnt, pb_lo, pb_cnt;
    size_t smem_fwd, smem_rev;
    size_t smem_il;   // reverse sweep on interleaved (ket, bra) elements (lean_batch_kernel<3>)
    int max_rows, max_cols;
    // persistent kernel (lean_persist_kernel): shared memory and resident CTAs per SM; 0 CTAs = not usable for this rectangle
    size_t smem_p_fwd = 0, smem_p_rev = 0;
    int ctas_p_fwd = 0, ctas_p_rev = 0;
};

// Auxiliary streams for the rectangle launches of one batch: the rectangles are independent sets of tiles, so they
// run concurrently (fork from / join into the caller's stream with events) and fill each other's tails.
struct StreamFan {
    std::vector<cudaStream_t> aux;
    std::vector<cudaEvent_t> done;
    cudaEvent_t start = nullptr;
    cudaError_t create(int n_aux);
    void destroy();
};

cudaError_t batch_kernels_configure();
size_t batch_smem_bytes(int max_rows, int max_cols, int nops, int nv);
size_t lean_smem_bytes(int max_rows, int max_cols, int nops, int nv);
size_t lean_persist_smem_bytes(int max_rows, int max_cols, int nops, int nv);
void persist_config(int nv, int threads, int max_rows, int max_cols, int nops, size_t* smem, int* ctas_per_sm);
void launch_batch_reverse_il(double* kb, double* kb_out, int64_t nb, const BatchDev& bd, const double2* rot, const std::vector<TileRect>& rects,
                             int threads, double* partials, double* grad_ops, int n_sets, SetStrides ss, cudaStream_t st, StreamFan* fan,
                             unsigned* ctl = nullptr, int sms = 148, int min_waves = 4);
void launch_interleave(const double* a, const double* b, double* out, int64_t n, cudaStream_t st);
void launch_deinterleave(const double* in, double* a, double* b, int64_t n, cudaStream_t st);
void launch_batch_forward(double* v, double* v_out, int64_t nb, const BatchDev& bd, const double2* rot, const std::vector<TileRect>& rects,
                          int threads, int n_sets, SetStrides ss, cudaStream_t st, StreamFan* fan, unsigned* ctl = nullptr, int sms = 148, int min_waves = 4);
void launch_batch_reverse(double* ket, double* bra, int64_t nb, const BatchDev& bd, const double2* rot,
                          const std::vector<TileRect>& rects, int threads, double* partials, double* grad_ops, int n_sets,
                          SetStrides ss, cudaStream_t st, StreamFan* fan);
void launch_permute(const double* in, double* out, int64_t na, int64_t nb, const uint32_t* map_a, const uint32_t* map_b,
                    cudaStream_t st);

}  // namespace tcc
