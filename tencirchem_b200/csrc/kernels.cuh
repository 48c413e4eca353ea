// sm_100a kernels of the civector engine (declarations of the host launchers).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace tcc {

struct Link;

// One spin side of an excitation as the kernels see it: a list of (src, dst|sign<<31) string
// addresses, or the identity (src == nullptr) over `len` strings.
struct SideList {
    const uint32_t* src;
    const uint32_t* dst;
    int len;
};

// c'_I = cos c_I + sin p c_J ; c'_J = cos c_J - sin p c_I over all pairs (I=(a,b), J=(a',b')) of the op
void launch_rotate(double* c, int64_t nb, SideList rows, SideList cols, double cs, double sn_s0,
                   cudaStream_t st);
// same on ket and bra with -theta, plus *grad_out = sum p0 (bra_I ket_J - bra_J ket_I) (deterministic)
void launch_reverse(double* ket, double* bra, int64_t nb, SideList rows, SideList cols, double cs,
                    double sn_s0, double s0, double* partials, unsigned* ticket, double* grad_out,
                    cudaStream_t st);
// out = G c (out pre-zeroed)
void launch_apply_g(const double* c, double* out, int64_t nb, SideList rows, SideList cols, double s0,
                    cudaStream_t st);
int reverse_partials_capacity();

void launch_set_unit(double* c, int64_t n, int64_t one_at, int n_sets, cudaStream_t st);
void launch_dot_sets(const double* x, const double* y, int64_t n, int n_sets, double* out, int64_t out_stride, cudaStream_t st);
// deterministic dot product: out[0] = sum x*y
void launch_dot(const double* x, const double* y, int64_t n, double* partials, unsigned* ticket,
                double* out, cudaStream_t st);
void launch_transpose(const double* in, double* out, int64_t rows, int64_t cols, bool accumulate, int n_sets,
                      cudaStream_t st);
// out[r, :] (+)= sum_j val[r, j] * in[col[r, j], :]   (ELL rows, dense row vectors of length ncols)
void launch_spmm_rows(const double* in, double* out, int64_t nrows, int64_t ncols, const uint32_t* ell_col,
                      const double* ell_val, int width, bool accumulate, int n_sets, cudaStream_t st);
// sigma(a', b') += sum_{pq,rs} W[pq,rs] <a'|E_pq|a><b'|E_rs|b> c(a,b)  (opposite-spin part, DMMA)
void launch_sigma_ab(const double* c, double* sigma, int64_t na, int64_t nb, const Link* links, int ml,
                     const double* w, int n, int n2p, const uint16_t* kmap, int n_sets, cudaStream_t st, int64_t row_lo = 0,
                     int64_t row_cnt = -1, const Link* alinks = nullptr, int aml = 0);
// hard-core-boson sigma (hamiltonian.py:158-183)
void launch_sigma_hcb(const double* c, double* sigma, int64_t nstr, const uint32_t* strings, int n, int k,
                      const double* diag1, const double* hop, const double* diag2, const int64_t* binom,
                      int binom_ld, const int* bitpos, int n_sets, cudaStream_t st);

// row move over peer memory: dst(rank dst_rank[i])[dst_row[i], :] = src[src_row[i], :] (kernels.cu)
void launch_push_rows(const double* src, int64_t row_doubles, int64_t n_rows, const int64_t* src_row, const int64_t* dst_row,
                      const int32_t* dst_rank, const uint64_t* peer_base, int64_t dst_offset, cudaStream_t st);

// Gram matrix of the single-replacement intermediate D (rdm_kernels.cu); gram [n^2, n^2] must be zeroed; only the
// blocks I <= J of the 64 x 64 block grid are written
void launch_rdm_gram(const double* c, int64_t na, int64_t nb, const Link* links, int ml, int n, double* gram, int slices,
                     cudaStream_t st);

// hard-core-boson density-matrix elements (puccd.py:148-203): hop[p,q], both[p,q] for q <= p, both[p,p] = occupation
void launch_hcb_rdm(const double* c, int64_t nstr, const uint32_t* strings, int n, const int64_t* binom, int binom_ld,
                    const int* bitpos, double* hop, double* both, cudaStream_t st);

}  // namespace tcc
