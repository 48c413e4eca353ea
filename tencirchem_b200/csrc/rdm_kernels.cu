// Spin-traced reduced density matrices of a CI vector on the GPU (tencirchem/static/ucc.py:755-852, which hands the
// vector to PySCF's direct_spin1.make_rdm1 / make_rdm12 on the host).
//
// With D[pq, K] = <K| E_pq |c> (E_pq = spin-summed single replacement, the intermediate of the direct-CI sigma),
//   <c| E_pq E_rs |c> = sum_K D[qp, K] D[rs, K] = Gram[qp, rs],
// and PySCF's dm2[p,q,r,s] = <p+ r+ s q> = Gram[qp, rs] - delta_qr <E_ps>,  <E_ps> = sum_q Gram[sp, qq] / n_elec.
// The Gram matrix of D ([n^2, n^2], 2 n^4 N flops) is the whole cost.  D is never stored: a CTA owns one 64 x 64 block
// of the Gram matrix, loops over (alpha row, 32-column tile) units of the CI matrix, rebuilds the two 64-row slices of
// D it needs in shared memory through the single-replacement link table, and accumulates D_I D_J^T on the FP64 tensor
// path (mma.sync m8n8k4) in registers; one atomicAdd per block element at the end.
#include "kernels.cuh"
#include "plan.hpp"

#include <algorithm>

namespace tcc {

constexpr int RDM_T = 32;        // columns per unit = K of one accumulation step
constexpr int RDM_LD = RDM_T + 4;  // 4 (mod 16): conflict-free fragment loads, as in sigma_ab_kernel
constexpr int RDM_B = 64;        // block of D rows

__device__ __forceinline__ void rdm_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// grid.x = persistent slices over the units, grid.y = block pair (I <= J); 256 threads
__global__ void __launch_bounds__(256) rdm_gram_kernel(const double* __restrict__ c, int64_t na, int64_t nb,
                                                       const Link* __restrict__ links, int ml, int n, int n2,
                                                       double* __restrict__ gram, int gram_ld, int n_blocks) {
    __shared__ double dI[RDM_B * RDM_LD];
    __shared__ double dJ[RDM_B * RDM_LD];
    // block pair from the linear index: (I, J) with I <= J
    int bi = 0, rem = blockIdx.y;
    while (rem >= n_blocks - bi) {
        rem -= n_blocks - bi;
        ++bi;
    }
    const int bj = bi + rem;
    const bool diag = bi == bj;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q4 = lane & 3;
    const int64_t col_tiles = (nb + RDM_T - 1) / RDM_T;
    const int64_t units = na * col_tiles;
    double acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;

    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int64_t a = u / col_tiles;
        const int64_t b0 = (u - a * col_tiles) * RDM_T;
        for (int i = tid; i < RDM_B * RDM_LD; i += 256) {
            dI[i] = 0.0;
            dJ[i] = 0.0;
        }
        __syncthreads();
        // beta replacements: D[(q,p), (a,b)] = sgn * c(a, b2) for every forward link E_pq |b> = sgn |b2>
        const double* crow = c + a * nb;
        for (int i = tid; i < RDM_T * ml; i += 256) {
            const int bc = i / ml, l = i - bc * ml;
            const int64_t b = b0 + bc;
            if (b < nb) {
                const Link lk = links[b * ml + l];
                if (lk.sgn != 0) {
                    const int p = lk.pq / n, q = lk.pq - p * n;
                    const int idx = q * n + p, blk = idx >> 6;
                    const double v = double(lk.sgn) * __ldg(crow + lk.addr);
                    if (blk == bi) dI[(idx & 63) * RDM_LD + bc] = v;
                    if (blk == bj) dJ[(idx & 63) * RDM_LD + bc] = v;
                }
            }
        }
        __syncthreads();
        // alpha replacements: D[(q,p), (a,b)] += sgn * c(a2, b) for every forward link E_pq |a> = sgn |a2>;
        // one warp per link, lanes over the 32 columns (a (row, idx) pair has at most one alpha link: no races)
        for (int l = warp; l < ml; l += 8) {
            const Link lk = links[a * ml + l];
            if (lk.sgn == 0) continue;
            const int p = lk.pq / n, q = lk.pq - p * n;
            const int idx = q * n + p, blk = idx >> 6;
            if (blk != bi && blk != bj) continue;
            const int64_t b = b0 + lane;
            const double v = b < nb ? double(lk.sgn) * __ldg(c + int64_t(lk.addr) * nb + b) : 0.0;
            if (blk == bi) dI[(idx & 63) * RDM_LD + lane] += v;
            if (blk == bj) dJ[(idx & 63) * RDM_LD + lane] += v;
        }
        __syncthreads();
        // Gram block += D_I D_J^T : warp w owns row tile w (8 rows of I) x all 8 column tiles of J
        const double* arow = dI + (warp * 8 + g) * RDM_LD + q4;
        const double* bbase = (diag ? dI : dJ) + g * RDM_LD + q4;
#pragma unroll
        for (int k0 = 0; k0 < RDM_T; k0 += 4) {
            const double av = arow[k0];
#pragma unroll
            for (int j = 0; j < 8; ++j) rdm_dmma(acc[j][0], acc[j][1], av, bbase[j * 8 * RDM_LD + k0]);
        }
        __syncthreads();
    }
    // accumulator layout of m8n8k4: row g, columns 2*q4, 2*q4 + 1
    const int row = bi * RDM_B + warp * 8 + g;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = bj * RDM_B + j * 8 + q4 * 2;
        if (row < n2) {
            if (col < n2) atomicAdd(gram + int64_t(row) * gram_ld + col, acc[j][0]);
            if (col + 1 < n2) atomicAdd(gram + int64_t(row) * gram_ld + col + 1, acc[j][1]);
        }
    }
}

void launch_rdm_gram(const double* c, int64_t na, int64_t nb, const Link* links, int ml, int n, double* gram, int slices,
                     cudaStream_t st) {
    const int n2 = n * n;
    const int n_blocks = (n2 + RDM_B - 1) / RDM_B;
    const int pairs = n_blocks * (n_blocks + 1) / 2;
    const int64_t units = na * ((nb + RDM_T - 1) / RDM_T);
    const unsigned gx = unsigned(std::max<int64_t>(1, std::min<int64_t>(slices, units)));
    rdm_gram_kernel<<<dim3(gx, unsigned(pairs), 1), 256, 0, st>>>(c, na, nb, links, ml, n, n2, gram, n2, n_blocks);
}


// ------------------------------------------------------------------------------------------------
// Hard-core-boson (pUCCD) density-matrix elements, tencirchem/static/puccd.py:148-203.  For the pair (p, q) of block
// blockIdx.x = p * n + q, q <= p, over the strings I of the hcb CI vector (storage order):
//   p == q:  occ[p]     = sum_I c_I^2 [p in I]
//   p != q:  hop[p, q]  = sum_I c_I c_{I xor pq} [q in I, p not in I]     (the pair hops from q to p)
//            both[p, q] = sum_I c_I^2 [p in I and q in I]
// One block per pair, fixed-order block reduction (deterministic).  The vectors are C(n, k) long: tiny.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hcb_rdm_kernel(const double* __restrict__ c, int64_t nstr, const uint32_t* __restrict__ strings,
                                                      int n, const int64_t* __restrict__ binom, int binom_ld,
                                                      const int* __restrict__ bitpos, double* __restrict__ hop,
                                                      double* __restrict__ both) {
    __shared__ double red[2][8];
    const int p = blockIdx.x / n, q = blockIdx.x - p * n;
    if (q > p) return;
    const uint32_t mp = 1u << p, mq = 1u << q;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = threadIdx.x; i < nstr; i += 256) {
        const uint32_t s = strings[i];
        const double ci = c[i];
        if (p == q) {
            if (s & mq) a1 += ci * ci;
            continue;
        }
        if ((s & mq) && (s & mp)) a1 += ci * ci;
        if ((s & mq) && !(s & mp)) {
            const uint32_t t = s ^ mp ^ mq;
            uint32_t tp = 0;
            for (uint32_t x = t; x; x &= x - 1) tp |= 1u << bitpos[__ffs(x) - 1];
            int64_t r = 0;
            int idx = 1;
            for (uint32_t x = tp; x; x &= x - 1) {
                r += binom[(__ffs(x) - 1) * binom_ld + idx];
                ++idx;
            }
            a0 += ci * c[r];
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, o);
        a1 += __shfl_down_sync(0xffffffffu, a1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = a0;
        red[1][threadIdx.x >> 5] = a1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < 8; ++w) {
            t0 += red[0][w];
            t1 += red[1][w];
        }
        hop[p * n + q] = t0;
        both[p * n + q] = t1;  // p == q: occ[p]
    }
}

void launch_hcb_rdm(const double* c, int64_t nstr, const uint32_t* strings, int n, const int64_t* binom, int binom_ld,
                    const int* bitpos, double* hop, double* both, cudaStream_t st) {
    hcb_rdm_kernel<<<unsigned(n * n), 256, 0, st>>>(c, nstr, strings, n, binom, binom_ld, bitpos, hop, both);
}

}  // namespace tcc
