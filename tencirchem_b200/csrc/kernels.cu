// sm_100a kernels of the civector UCC engine.
//
// Layout: the CI vector is a row-major [na, nb] fp64 matrix, alpha string = row, beta string = column
// (tencirchem/static/ci_utils.py:22-25).  An excitation is a set of independent 2x2 rotations between
// determinant pairs I = (a, b) <-> J = (a', b') with (a -> a') from the alpha hop list and (b -> b') from
// the beta hop list; the sign of the matrix element G_IJ is s0 * sign_a * sign_b.
#include "kernels.cuh"

#include <algorithm>
#include <cstdio>

#include "plan.hpp"

namespace tcc {

// ------------------------------------------------------------------------------------------------
// pair rotations
// ------------------------------------------------------------------------------------------------
constexpr int ROT_THREADS = 256;
constexpr int ROT_ROWS = 4;
constexpr int MAX_PARTIALS = 32768;

int reverse_partials_capacity() { return MAX_PARTIALS; }

struct PairCol {
    uint32_t b, b2, sb;
};

template <bool COL_ID>
__device__ __forceinline__ PairCol load_col(const uint32_t* __restrict__ csrc, const uint32_t* __restrict__ cdst, int t) {
    PairCol pc;
    if (COL_ID) {
        pc.b = pc.b2 = uint32_t(t);
        pc.sb = 0;
    } else {
        pc.b = __ldg(csrc + t);
        uint32_t d = __ldg(cdst + t);
        pc.b2 = d & 0x7fffffffu;
        pc.sb = d >> 31;
    }
    return pc;
}

template <bool ROW_ID>
__device__ __forceinline__ void load_row(const uint32_t* __restrict__ rsrc, const uint32_t* __restrict__ rdst, int r,
                                         uint32_t& a, uint32_t& a2, uint32_t& sa) {
    if (ROW_ID) {
        a = a2 = uint32_t(r);
        sa = 0;
    } else {
        a = __ldg(rsrc + r);
        uint32_t d = __ldg(rdst + r);
        a2 = d & 0x7fffffffu;
        sa = d >> 31;
    }
}

template <bool ROW_ID, bool COL_ID>
__global__ void __launch_bounds__(ROT_THREADS)
rotate_kernel(double* __restrict__ c, int64_t nb, const uint32_t* __restrict__ rsrc, const uint32_t* __restrict__ rdst,
              int la, const uint32_t* __restrict__ csrc, const uint32_t* __restrict__ cdst, int lb, double cs, double sn) {
    int t = blockIdx.x * ROT_THREADS + threadIdx.x;
    if (t >= lb) return;
    PairCol pc = load_col<COL_ID>(csrc, cdst, t);
    for (int r0 = blockIdx.y * ROT_ROWS; r0 < la; r0 += gridDim.y * ROT_ROWS) {
        double x[ROT_ROWS], y[ROT_ROWS], p[ROT_ROWS];
        int64_t I[ROT_ROWS], J[ROT_ROWS];
#pragma unroll
        for (int r = 0; r < ROT_ROWS; ++r) {
            if (r0 + r < la) {
                uint32_t a, a2, sa;
                load_row<ROW_ID>(rsrc, rdst, r0 + r, a, a2, sa);
                I[r] = int64_t(a) * nb + pc.b;
                J[r] = int64_t(a2) * nb + pc.b2;
                x[r] = c[I[r]];
                y[r] = c[J[r]];
                p[r] = (sa ^ pc.sb) ? -sn : sn;
            }
        }
#pragma unroll
        for (int r = 0; r < ROT_ROWS; ++r) {
            if (r0 + r < la) {
                c[I[r]] = cs * x[r] + p[r] * y[r];
                c[J[r]] = cs * y[r] - p[r] * x[r];
            }
        }
    }
}

static dim3 pair_grid(int la, int lb) {
    unsigned gx = (lb + ROT_THREADS - 1) / ROT_THREADS;
    unsigned gy = (la + ROT_ROWS - 1) / ROT_ROWS;
    unsigned cap = MAX_PARTIALS / (gx ? gx : 1);
    if (cap < 1) cap = 1;
    if (gy > cap) gy = cap;
    if (gy > 65535) gy = 65535;
    return dim3(gx, gy, 1);
}

void launch_rotate(double* c, int64_t nb, SideList rows, SideList cols, double cs, double sn_s0, cudaStream_t st) {
    if (rows.len <= 0 || cols.len <= 0) return;
    dim3 grid = pair_grid(rows.len, cols.len);
    bool rid = rows.src == nullptr, cid = cols.src == nullptr;
    if (rid && !cid)
        rotate_kernel<true, false><<<grid, ROT_THREADS, 0, st>>>(c, nb, rows.src, rows.dst, rows.len, cols.src, cols.dst, cols.len, cs, sn_s0);
    else if (!rid && cid)
        rotate_kernel<false, true><<<grid, ROT_THREADS, 0, st>>>(c, nb, rows.src, rows.dst, rows.len, cols.src, cols.dst, cols.len, cs, sn_s0);
    else
        rotate_kernel<false, false><<<grid, ROT_THREADS, 0, st>>>(c, nb, rows.src, rows.dst, rows.len, cols.src, cols.dst, cols.len, cs, sn_s0);
}

// block-wide sum, result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) warp_part[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? warp_part[threadIdx.x] : 0.0;
    if (w == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    return v;
}

// Deterministic grid reduction: every block stores its partial; the block that takes the last ticket
// sums all partials in a fixed order and writes the result.
__device__ __forceinline__ void grid_sum_finish(double block_val, double* partials, unsigned* ticket, double* out) {
    __shared__ bool is_last;
    unsigned nblocks = gridDim.x * gridDim.y;
    unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        partials[bid] = block_val;
        __threadfence();
        unsigned tk = atomicAdd(ticket, 1u);
        is_last = (tk == nblocks - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double v = 0.0;
        for (unsigned i = threadIdx.x; i < nblocks; i += blockDim.x) v += __ldcg(partials + i);
        v = block_sum(v);
        if (threadIdx.x == 0) {
            *out = v;
            *ticket = 0;
        }
    }
}

template <bool ROW_ID, bool COL_ID>
__global__ void __launch_bounds__(ROT_THREADS)
reverse_kernel(double* __restrict__ ket, double* __restrict__ bra, int64_t nb, const uint32_t* __restrict__ rsrc,
               const uint32_t* __restrict__ rdst, int la, const uint32_t* __restrict__ csrc,
               const uint32_t* __restrict__ cdst, int lb, double cs, double sn, double s0, double* partials,
               unsigned* ticket, double* out) {
    int t = blockIdx.x * ROT_THREADS + threadIdx.x;
    double acc = 0.0;
    if (t < lb) {
        PairCol pc = load_col<COL_ID>(csrc, cdst, t);
        for (int r0 = blockIdx.y * ROT_ROWS; r0 < la; r0 += gridDim.y * ROT_ROWS) {
            double kx[ROT_ROWS], ky[ROT_ROWS], bx[ROT_ROWS], by[ROT_ROWS];
            int64_t I[ROT_ROWS], J[ROT_ROWS];
            bool neg[ROT_ROWS];
#pragma unroll
            for (int r = 0; r < ROT_ROWS; ++r) {
                if (r0 + r < la) {
                    uint32_t a, a2, sa;
                    load_row<ROW_ID>(rsrc, rdst, r0 + r, a, a2, sa);
                    I[r] = int64_t(a) * nb + pc.b;
                    J[r] = int64_t(a2) * nb + pc.b2;
                    kx[r] = ket[I[r]];
                    ky[r] = ket[J[r]];
                    bx[r] = bra[I[r]];
                    by[r] = bra[J[r]];
                    neg[r] = (sa ^ pc.sb) != 0;
                }
            }
#pragma unroll
            for (int r = 0; r < ROT_ROWS; ++r) {
                if (r0 + r < la) {
                    double p = neg[r] ? -sn : sn;
                    double nkx = cs * kx[r] + p * ky[r], nky = cs * ky[r] - p * kx[r];
                    double nbx = cs * bx[r] + p * by[r], nby = cs * by[r] - p * bx[r];
                    ket[I[r]] = nkx;
                    ket[J[r]] = nky;
                    bra[I[r]] = nbx;
                    bra[J[r]] = nby;
                    double g = nbx * nky - nby * nkx;  // bra_I (G ket)_I + bra_J (G ket)_J without the sign
                    acc += neg[r] ? -g : g;
                }
            }
        }
    }
    double v = block_sum(acc * s0);
    grid_sum_finish(v, partials, ticket, out);
}

void launch_reverse(double* ket, double* bra, int64_t nb, SideList rows, SideList cols, double cs, double sn_s0,
                    double s0, double* partials, unsigned* ticket, double* grad_out, cudaStream_t st) {
    if (rows.len <= 0 || cols.len <= 0) {  // empty support: nothing to rotate, the gradient term is zero
        cudaMemsetAsync(grad_out, 0, sizeof(double), st);
        return;
    }
    dim3 grid = pair_grid(rows.len, cols.len);
    bool rid = rows.src == nullptr, cid = cols.src == nullptr;
    if (rid && !cid)
        reverse_kernel<true, false><<<grid, ROT_THREADS, 0, st>>>(ket, bra, nb, rows.src, rows.dst, rows.len, cols.src, cols.dst, cols.len, cs, sn_s0, s0, partials, ticket, grad_out);
    else if (!rid && cid)
        reverse_kernel<false, true><<<grid, ROT_THREADS, 0, st>>>(ket, bra, nb, rows.src, rows.dst, rows.len, cols.src, cols.dst, cols.len, cs, sn_s0, s0, partials, ticket, grad_out);
    else
        reverse_kernel<false, false><<<grid, ROT_THREADS, 0, st>>>(ket, bra, nb, rows.src, rows.dst, rows.len, cols.src, cols.dst, cols.len, cs, sn_s0, s0, partials, ticket, grad_out);
}

template <bool ROW_ID, bool COL_ID>
__global__ void __launch_bounds__(ROT_THREADS)
apply_g_kernel(const double* __restrict__ c, double* __restrict__ out, int64_t nb, const uint32_t* __restrict__ rsrc,
               const uint32_t* __restrict__ rdst, int la, const uint32_t* __restrict__ csrc,
               const uint32_t* __restrict__ cdst, int lb, double s0) {
    int t = blockIdx.x * ROT_THREADS + threadIdx.x;
    if (t >= lb) return;
    PairCol pc = load_col<COL_ID>(csrc, cdst, t);
    for (int r = blockIdx.y; r < la; r += gridDim.y) {
        uint32_t a, a2, sa;
        load_row<ROW_ID>(rsrc, rdst, r, a, a2, sa);
        int64_t I = int64_t(a) * nb + pc.b, J = int64_t(a2) * nb + pc.b2;
        double p = (sa ^ pc.sb) ? -s0 : s0;
        out[I] = p * c[J];
        out[J] = -p * c[I];
    }
}

void launch_apply_g(const double* c, double* out, int64_t nb, SideList rows, SideList cols, double s0, cudaStream_t st) {
    if (rows.len <= 0 || cols.len <= 0) return;
    dim3 grid((cols.len + ROT_THREADS - 1) / ROT_THREADS, rows.len > 65535 ? 65535 : rows.len, 1);
    bool rid = rows.src == nullptr, cid = cols.src == nullptr;
    if (rid && !cid)
        apply_g_kernel<true, false><<<grid, ROT_THREADS, 0, st>>>(c, out, nb, rows.src, rows.dst, rows.len, cols.src, cols.dst, cols.len, s0);
    else if (!rid && cid)
        apply_g_kernel<false, true><<<grid, ROT_THREADS, 0, st>>>(c, out, nb, rows.src, rows.dst, rows.len, cols.src, cols.dst, cols.len, s0);
    else
        apply_g_kernel<false, false><<<grid, ROT_THREADS, 0, st>>>(c, out, nb, rows.src, rows.dst, rows.len, cols.src, cols.dst, cols.len, s0);
}

// ------------------------------------------------------------------------------------------------
// vector utilities
// ------------------------------------------------------------------------------------------------
__global__ void set_unit_kernel(double* c, int64_t n, int64_t one_at) {
    c += int64_t(blockIdx.y) * n;  // one vector per parameter set
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) c[i] = (i == one_at) ? 1.0 : 0.0;
}

void launch_set_unit(double* c, int64_t n, int64_t one_at, int n_sets, cudaStream_t st) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    set_unit_kernel<<<dim3(unsigned(blocks), unsigned(n_sets), 1), 256, 0, st>>>(c, n, one_at);
}

// one CTA per parameter set: out[set * out_stride] = <x_set | y_set>  (vectors of the small-N regime)
__global__ void __launch_bounds__(256) dot_sets_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                                                       double* __restrict__ out, int64_t out_stride) {
    const double* xs = x + int64_t(blockIdx.x) * n;
    const double* ys = y + int64_t(blockIdx.x) * n;
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) a += xs[i] * ys[i];
    a = block_sum(a);
    if (threadIdx.x == 0) out[int64_t(blockIdx.x) * out_stride] = a;
}

void launch_dot_sets(const double* x, const double* y, int64_t n, int n_sets, double* out, int64_t out_stride, cudaStream_t st) {
    dot_sets_kernel<<<unsigned(n_sets), 256, 0, st>>>(x, y, n, out, out_stride);
}

__global__ void __launch_bounds__(256) dot_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                                                  double* partials, unsigned* ticket, double* out) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t stride = int64_t(gridDim.x) * blockDim.x;
    double a0 = 0.0, a1 = 0.0;
    for (; i + stride < n; i += 2 * stride) {
        a0 += x[i] * y[i];
        a1 += x[i + stride] * y[i + stride];
    }
    if (i < n) a0 += x[i] * y[i];
    double v = block_sum(a0 + a1);
    grid_sum_finish(v, partials, ticket, out);
}

void launch_dot(const double* x, const double* y, int64_t n, double* partials, unsigned* ticket, double* out, cudaStream_t st) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    dot_kernel<<<unsigned(blocks), 256, 0, st>>>(x, y, n, partials, ticket, out);
}

template <bool ACC>
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t rows, int64_t cols) {
    __shared__ double tile[32][33];
    in += int64_t(blockIdx.z) * rows * cols;
    out += int64_t(blockIdx.z) * rows * cols;
    int64_t c0 = int64_t(blockIdx.x) * 32, r0 = int64_t(blockIdx.y) * 32;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        int64_t r = r0 + ty + j, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + j][tx] = in[r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        int64_t c = c0 + ty + j, r = r0 + tx;  // out is [cols, rows]
        if (r < rows && c < cols) {
            double v = tile[tx][ty + j];
            if (ACC)
                out[c * rows + r] += v;
            else
                out[c * rows + r] = v;
        }
    }
}

void launch_transpose(const double* in, double* out, int64_t rows, int64_t cols, bool accumulate, int n_sets, cudaStream_t st) {
    dim3 grid(unsigned((cols + 31) / 32), unsigned((rows + 31) / 32), unsigned(n_sets));
    if (accumulate)
        transpose_kernel<true><<<grid, 256, 0, st>>>(in, out, rows, cols);
    else
        transpose_kernel<false><<<grid, 256, 0, st>>>(in, out, rows, cols);
}

// ------------------------------------------------------------------------------------------------
// same-spin sigma: out[r, :] (+)= sum_j val[r, j] * in[col[r, j], :]
// grid.x = output row (fast index), grid.y = 512-column slab: all rows of one slab run together so
// the slab (nrows x 4 KB) stays L2-resident while every output row gathers its source rows from it.
// ------------------------------------------------------------------------------------------------
constexpr int SPMM_CHUNK = 256;

template <bool ACC>
__global__ void __launch_bounds__(256) spmm_rows_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t nrows,
                                                        int64_t ncols, const uint32_t* __restrict__ ell_col,
                                                        const double* __restrict__ ell_val, int width) {
    __shared__ uint32_t s_col[SPMM_CHUNK];
    __shared__ double s_val[SPMM_CHUNK];
    in += int64_t(blockIdx.z) * nrows * ncols;
    out += int64_t(blockIdx.z) * nrows * ncols;
    int64_t r = blockIdx.x;
    int64_t c0 = int64_t(blockIdx.y) * 512 + threadIdx.x, c1 = c0 + 256;
    bool ok0 = c0 < ncols, ok1 = c1 < ncols;
    double acc0 = 0.0, acc1 = 0.0;
    for (int j0 = 0; j0 < width; j0 += SPMM_CHUNK) {
        int m = min(SPMM_CHUNK, width - j0);
        __syncthreads();
        if (threadIdx.x < m) {
            s_col[threadIdx.x] = ell_col[r * width + j0 + threadIdx.x];
            s_val[threadIdx.x] = ell_val[r * width + j0 + threadIdx.x];
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < m; ++j) {
            const double* src = in + int64_t(s_col[j]) * ncols;
            double v = s_val[j];
            if (ok0) acc0 += v * src[c0];
            if (ok1) acc1 += v * src[c1];
        }
    }
    if (ok0) out[r * ncols + c0] = ACC ? out[r * ncols + c0] + acc0 : acc0;
    if (ok1) out[r * ncols + c1] = ACC ? out[r * ncols + c1] + acc1 : acc1;
}

void launch_spmm_rows(const double* in, double* out, int64_t nrows, int64_t ncols, const uint32_t* ell_col,
                      const double* ell_val, int width, bool accumulate, int n_sets, cudaStream_t st) {
    dim3 grid(unsigned(nrows), unsigned((ncols + 511) / 512), unsigned(n_sets));
    if (accumulate)
        spmm_rows_kernel<true><<<grid, 256, 0, st>>>(in, out, nrows, ncols, ell_col, ell_val, width);
    else
        spmm_rows_kernel<false><<<grid, 256, 0, st>>>(in, out, nrows, ncols, ell_col, ell_val, width);
}

// ------------------------------------------------------------------------------------------------
// opposite-spin sigma on the FP64 tensor path (DMMA m8n8k4)
//   CTA = (source row a, tile of T output columns b').
//   1. D[rs, b'] = sum_b <b'|E_rs|b> c(a, b)        -- gather through the beta link table into smem
//   2. G[l, b']  = sum_rs W[pq_l, rs] D[rs, b']     -- [ml x n2] . [n2 x T] on mma.sync f64, A rows picked by
//                                                       the alpha links l of string a (E_pq|a> = sgn_l |a'_l>)
//   3. sigma(a'_l, b') += sgn_l G[l, b']            -- coalesced fp64 reductions into the target rows
// ------------------------------------------------------------------------------------------------
#ifndef SAB_T_COLS
#define SAB_T_COLS 32
#endif
constexpr int SAB_T = SAB_T_COLS;
constexpr int SAB_LD = SAB_T + 4;  // LD = 4 (mod 16): the B-fragment loads of a half-warp hit 16 distinct 8-byte banks

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void sigma_ab_kernel(const double* __restrict__ c, double* __restrict__ sigma, int64_t nb,
                                const Link* __restrict__ links, int ml, const double* __restrict__ w, int n, int n2p,
                                int64_t row0, int64_t set_stride, const uint16_t* __restrict__ kmap,
                                const Link* __restrict__ alinks, int aml) {
    // links / ml: single replacements of the beta strings (columns, always the full table); alinks / aml: those of the
    // alpha strings = rows of c and sigma (the same table, or the filtered table with local rows of a sharded plan)
    // n2p = K of the GEMM: n^2 (padded) in general, n(n+1)/2 (padded) when W[pq,rs] == W[pq,sr] was verified at
    // set-up; kmap[p*n+q] = row of D that a forward beta link E_pq feeds
    extern __shared__ double dsm[];  // [n2p][SAB_LD]
    const int64_t a = row0 + blockIdx.y;
    const int64_t b0 = int64_t(blockIdx.x) * SAB_T;
    c += int64_t(blockIdx.z) * set_stride;
    sigma += int64_t(blockIdx.z) * set_stride;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    for (int i = tid; i < n2p * SAB_LD; i += nthreads) dsm[i] = 0.0;
    __syncthreads();
    const double* crow = c + a * nb;
    for (int i = tid; i < SAB_T * ml; i += nthreads) {
        int bc = i / ml, l = i - bc * ml;
        int64_t bp = b0 + bc;
        if (bp < nb) {
            Link lk = links[bp * ml + l];
            if (lk.sgn != 0) {
                // forward link E_pq|b'> = sgn|b>  <=>  <b'|E_qp|b> = sgn : contributes to D[(q,p), b']
                dsm[int(__ldg(kmap + lk.pq)) * SAB_LD + bc] = double(lk.sgn) * __ldg(crow + lk.addr);
            }
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
    const int mt_count = aml >> 3;
    // unit of work of a warp: two 8-row tiles of G x 16 columns; the A fragments (rows of W) serve two column tiles and
    // the B fragments (columns of D) two row tiles: one load per DMMA instead of 1.5
    const int mp_count = (mt_count + 1) >> 1;
    const int units = mp_count * (SAB_T / 16);
    const int g = lane >> 2, q4 = lane & 3;
    for (int u = warp; u < units; u += nwarps) {
        const int mp = u / (SAB_T / 16), nh = u - mp * (SAB_T / 16);
        const int mt0 = mp * 2, mt1 = mp * 2 + 1;
        const bool two = mt1 < mt_count;
        const Link rl0 = alinks[a * aml + mt0 * 8 + g];
        Link rl1 = rl0;
        if (two) rl1 = alinks[a * aml + mt1 * 8 + g];
        const double* wrow0 = w + int64_t(rl0.pq) * n2p + q4;
        const double* wrow1 = w + int64_t(rl1.pq) * n2p + q4;
        const double* dcol = dsm + q4 * SAB_LD + nh * 16 + g;
        double acc0[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, acc1[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        if (two) {
#pragma unroll 4
            for (int k0 = 0; k0 < n2p; k0 += 4) {
                const double av0 = __ldg(wrow0 + k0), av1 = __ldg(wrow1 + k0);
                const double bv0 = dcol[k0 * SAB_LD];
                const double bv1 = dcol[k0 * SAB_LD + 8];
                dmma_m8n8k4(acc0[0][0], acc0[0][1], av0, bv0);
                dmma_m8n8k4(acc0[1][0], acc0[1][1], av0, bv1);
                dmma_m8n8k4(acc1[0][0], acc1[0][1], av1, bv0);
                dmma_m8n8k4(acc1[1][0], acc1[1][1], av1, bv1);
            }
        } else {
#pragma unroll 4
            for (int k0 = 0; k0 < n2p; k0 += 4) {
                const double av0 = __ldg(wrow0 + k0);
                const double bv0 = dcol[k0 * SAB_LD];
                const double bv1 = dcol[k0 * SAB_LD + 8];
                dmma_m8n8k4(acc0[0][0], acc0[0][1], av0, bv0);
                dmma_m8n8k4(acc0[1][0], acc0[1][1], av0, bv1);
            }
        }
        auto scatter = [&](const Link& rl, double (&acc)[2][2]) {
            if (rl.sgn == 0) return;
            const double sg = double(rl.sgn);
            double* srow = sigma + int64_t(rl.addr) * nb;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int64_t col = b0 + nh * 16 + nt * 8 + q4 * 2;
                if (col < nb) atomicAdd(srow + col, sg * acc[nt][0]);
                if (col + 1 < nb) atomicAdd(srow + col + 1, sg * acc[nt][1]);
            }
        };
        scatter(rl0, acc0);
        if (two) scatter(rl1, acc1);
    }
}

void launch_sigma_ab(const double* c, double* sigma, int64_t na, int64_t nb, const Link* links, int ml, const double* w,
                     int n, int n2p, const uint16_t* kmap, int n_sets, cudaStream_t st, int64_t row_lo, int64_t row_cnt,
                     const Link* alinks, int aml) {
    if (!alinks) {
        alinks = links;
        aml = ml;
    }
    if (aml <= 0) return;
    int units = ((aml / 8 + 1) / 2) * (SAB_T / 16);
    int best_nw = 4;
    double best_eff = 0.0;
    for (int nw = 4; nw <= 16; ++nw) {
        int rounds = (units + nw - 1) / nw;
        double eff = double(units) / double(rounds * nw);
        if (eff >= best_eff - 1e-12) {
            best_eff = eff;
            best_nw = nw;
        }
    }
    size_t smem = size_t(n2p) * SAB_LD * sizeof(double);
    // per device and cheap: no cached state (plans of several devices may live in one process)
    if (smem > 48 * 1024) cudaFuncSetAttribute(sigma_ab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    // source rows [row_lo, row_lo + row_cnt) (default: all); their contributions go to every linked row of sigma
    const int64_t r_begin = row_cnt < 0 ? 0 : row_lo, gy_total = row_cnt < 0 ? na : row_lo + row_cnt;
    unsigned gx = unsigned((nb + SAB_T - 1) / SAB_T);
    // grid.y is limited to 65535: launch in row batches
    for (int64_t r0 = r_begin; r0 < gy_total; r0 += 65535) {
        unsigned gy = unsigned(gy_total - r0 < 65535 ? gy_total - r0 : 65535);
        sigma_ab_kernel<<<dim3(gx, gy, unsigned(n_sets)), best_nw * 32, smem, st>>>(c, sigma, nb, links, ml, w, n, n2p, r0, na * nb, kmap, alinks, aml);
    }
}

// ------------------------------------------------------------------------------------------------
// hard-core-boson sigma (tencirchem/static/hamiltonian.py:158-183)
// ------------------------------------------------------------------------------------------------
__global__ void sigma_hcb_kernel(const double* __restrict__ c, double* __restrict__ sigma, int64_t nstr,
                                 const uint32_t* __restrict__ strings, int n, const double* __restrict__ diag1,
                                 const double* __restrict__ hop, const double* __restrict__ diag2,
                                 const int64_t* __restrict__ binom, int binom_ld, const int* __restrict__ bitpos) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nstr) return;
    c += int64_t(blockIdx.y) * nstr;
    sigma += int64_t(blockIdx.y) * nstr;
    uint32_t s = strings[i];
    double ci = c[i], acc = 0.0;
    for (int p = 0; p < n; ++p) {
        bool op = (s >> p) & 1u;
        if (op) acc += ci * diag1[p];
        for (int q = 0; q < p; ++q) {
            bool oq = (s >> q) & 1u;
            if (op != oq) {
                uint32_t t = s ^ (1u << p) ^ (1u << q);
                // storage address = rank of the string with its bits moved to their storage positions
                uint32_t tp = 0;
                for (uint32_t x = t; x; x &= x - 1) tp |= 1u << bitpos[__ffs(x) - 1];
                int64_t r = 0;
                int idx = 1;
                for (uint32_t x = tp; x; x &= x - 1) {
                    r += binom[(__ffs(x) - 1) * binom_ld + idx];
                    ++idx;
                }
                acc += c[r] * hop[p * n + q];
            } else if (op && oq) {
                acc += ci * diag2[p * n + q];
            }
        }
    }
    sigma[i] = acc;
}

void launch_sigma_hcb(const double* c, double* sigma, int64_t nstr, const uint32_t* strings, int n, int k,
                      const double* diag1, const double* hop, const double* diag2, const int64_t* binom, int binom_ld,
                      const int* bitpos, int n_sets, cudaStream_t st) {
    (void)k;
    unsigned blocks = unsigned((nstr + 127) / 128);
    sigma_hcb_kernel<<<dim3(blocks, unsigned(n_sets), 1), 128, 0, st>>>(c, sigma, nstr, strings, n, diag1, hop, diag2, binom, binom_ld, bitpos);
}


// ------------------------------------------------------------------------------------------------
// Row move of the alpha-row sharded engine over peer memory: every local row goes straight to its new owner.
//   dst(rank q)[dst_row[i], :] = src[src_row[i], :]   for the i-th moved row, q = dst_rank[i]
// peer_base[q] is the base address of rank q's destination block as mapped INTO THIS process (NVLink peer mapping of a
// symmetric allocation; for q == this rank the local address), so the copy is one pass of 16-byte stores over NVLink:
// no pack buffer, no collective, no unpack pass.  grid.x = moved rows x pieces of a row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) push_rows_kernel(const double* __restrict__ src, int64_t row_doubles, int pieces,
                                                        const int64_t* __restrict__ src_row, const int64_t* __restrict__ dst_row,
                                                        const int32_t* __restrict__ dst_rank, const uint64_t* __restrict__ peer_base,
                                                        int64_t dst_offset) {
    const int64_t i = blockIdx.x / pieces;
    const int piece = int(blockIdx.x - i * pieces);
    const double* s = src + src_row[i] * row_doubles;
    double* d = reinterpret_cast<double*>(peer_base[dst_rank[i]]) + dst_offset + dst_row[i] * row_doubles;
    // piece boundaries in units of 2 doubles so that 16-byte accesses stay aligned whenever the rows are
    const int64_t pairs = row_doubles >> 1;
    const int64_t lo = (pairs * piece) / pieces, hi = (pairs * (piece + 1)) / pieces;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
        const double2* s2 = reinterpret_cast<const double2*>(s);
        double2* d2 = reinterpret_cast<double2*>(d);
        for (int64_t k = lo + threadIdx.x; k < hi; k += 256) d2[k] = s2[k];
    } else {
        for (int64_t k = 2 * lo + threadIdx.x; k < 2 * hi; k += 256) d[k] = s[k];
    }
    if ((row_doubles & 1) && piece == pieces - 1 && threadIdx.x == 0) d[row_doubles - 1] = s[row_doubles - 1];
}

void launch_push_rows(const double* src, int64_t row_doubles, int64_t n_rows, const int64_t* src_row, const int64_t* dst_row,
                      const int32_t* dst_rank, const uint64_t* peer_base, int64_t dst_offset, cudaStream_t st) {
    if (n_rows <= 0 || row_doubles <= 0) return;
    // enough CTAs to fill the machine even for few long rows; a piece is at least 4 KB
    int pieces = int(std::max<int64_t>(1, std::min<int64_t>((row_doubles * 8) / 4096, (148 * 16 + n_rows - 1) / n_rows)));
    for (int64_t r0 = 0; r0 < n_rows; r0 += (int64_t(1) << 20)) {
        const int64_t nr = std::min<int64_t>(n_rows - r0, int64_t(1) << 20);
        push_rows_kernel<<<unsigned(nr * pieces), 256, 0, st>>>(src, row_doubles, pieces, src_row + r0, dst_row + r0, dst_rank + r0,
                                                              peer_base, dst_offset);
    }
}

}  // namespace tcc
