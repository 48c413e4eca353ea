// Batched excitation kernels: one launch applies a whole batch of UCC factors to every closed tile of the
// CI matrix, so the vector crosses HBM once per batch instead of once per excitation
// (tencirchem/static/evolve_civector.py:165-177 forward sweep, :221-243 reverse sweep).
//
// CTA = one tile = (panel of alpha classes) x (panel of beta classes), see plan.hpp.  The tile is gathered
// into shared memory (rows by warp, columns in ascending address order so the global accesses coalesce),
// all excitations of the batch are applied in order with a barrier between them, and the tile is written
// back.  The reverse-sweep variant holds the ket and the bra tile and accumulates <bra|G_j|ket> per
// excitation; per-tile partial sums go to a [n_ops, n_tiles] buffer that a second kernel reduces in a fixed
// order (deterministic gradients).
#include "batch_kernels.cuh"

namespace tcc {

#ifdef TCC_DEBUG_SKIP_MEM  // diagnostic build: no tile gather / scatter (the excitations run on whatever shared memory holds)
#define TCC_MEM_ROWS(n) 0
#else
#define TCC_MEM_ROWS(n) (n)
#endif

constexpr int BK_THREADS = 256;  // upper bound of the launch; the actual block size is a launch parameter
constexpr int BK_WARPS = BK_THREADS / 32;
constexpr int BK_DEPTH = 4;  // excitations per prefetch group
constexpr int BK_COLCH = 4;  // tile columns per lane kept in registers (tiles up to 128 columns)

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned s = unsigned(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src) : "memory");
}

// shared-memory accesses by 32-bit shared-window address: no generic-address arithmetic in the excitation loop
__device__ __forceinline__ double lds64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v) : "memory"); }

// per-op record resolved for the tile's (e_alpha, e_beta) at kernel start
struct SmemOp {
    int off, cnt;    // slice of the expanded pair list
    double cs, sn;   // cos, sin * s0
    uint32_t smask;  // spectator sign of the excitation per class of the tile: bit g = alpha class g, bit 16 + g = beta class g
    uint32_t pad;
};

// One rotation pass over the expanded pair list of an excitation inside the tile.  REP: the tile has a side
// without active orbitals whose `n_rep` rows (or columns) all get the same pair list, `stride` apart.
template <int NV, bool REP>
__device__ __forceinline__ double rotate_pairs(double* __restrict__ t0, double* __restrict__ t1, const SmemOp& op,
                                               const uint32_t* __restrict__ plist, int n_rep, int stride, int tid, int nthreads) {
    double acc = 0.0;
    const int cnt = op.cnt;
    const int total = REP ? cnt * n_rep : cnt;
    if (tid >= total) return acc;
    // the repeat path has one class on its active side: spectator sign = bit 0 of either half of the mask
    const bool flip = ((op.smask ^ (op.smask >> 16)) & 1u) != 0;
    const double cs = op.cs, sn = flip ? -op.sn : op.sn;
    const uint32_t* lp = plist + op.off;
    int k = tid, base = 0, q = 0, r = 0;
    if (REP) {
        q = nthreads / cnt;
        r = nthreads - q * cnt;
        const int rep = tid / cnt;
        k = tid - rep * cnt;
        base = rep * stride;
        q *= stride;
    }
    for (int idx = tid; idx < total; idx += nthreads) {
        const uint32_t e = __ldg(lp + k);
        const int i1 = base + int(e & 0x7fffu), i2 = base + int((e >> 15) & 0x7fffu);
        const bool neg = (e >> 30) & 1u;
        const double p = neg ? -sn : sn;
        const double x = t0[i1], y = t0[i2];
        const double nx = cs * x + p * y, ny = cs * y - p * x;
        t0[i1] = nx;
        t0[i2] = ny;
        if (NV == 2) {
            const double bx = t1[i1], by = t1[i2];
            const double nbx = cs * bx + p * by, nby = cs * by - p * bx;
            t1[i1] = nbx;
            t1[i2] = nby;
            // t0 = ket, t1 = bra: bra_I (G ket)_I + bra_J (G ket)_J = +-(bra_I ket_J - bra_J ket_I)
            const double g = nbx * ny - nby * nx;
            acc += neg ? -g : g;
        }
        if (REP) {
            k += r;
            base += q;
            if (k >= cnt) {
                k -= cnt;
                base += stride;
            }
        } else {
            k += nthreads;
        }
    }
    return flip ? -acc : acc;
}

template <int NV>
__global__ void __launch_bounds__(BK_THREADS)
batch_kernel(double* __restrict__ v0, double* __restrict__ v1, double* __restrict__ vout, int64_t nb, BatchDev bd,
             const double2* __restrict__ rot, double* __restrict__ partials, int n_tiles, SetStrides ss, TileRect rect) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // vout != null (NV == 1 only): dynamic column layout -- the tile is gathered from v0 with the columns in this batch's
    // order (B.dyn_in_*) and scattered to vout in the order of the next batch of the sweep (B.dyn_out_*)
    const bool dyn = NV == 1 && vout != nullptr;
    // blockIdx.y = parameter set (batched-theta evaluation): own vectors, angles and gradient partials
    v0 += int64_t(blockIdx.y) * ss.vec;
    if (NV == 2) v1 += int64_t(blockIdx.y) * ss.vec;
    rot += int64_t(blockIdx.y) * ss.rot;
    if (NV == 2) partials += int64_t(blockIdx.y) * ss.part;
    // this launch covers the rectangle [pa_lo, pa_lo + ..) x [pb_lo, pb_lo + pb_cnt) of the panel grid (tiles of
    // similar size share a launch so that the shared-memory request, hence the occupancy, fits them)
    const int pa_loc = blockIdx.x / rect.pb_cnt;
    const int pa_id = rect.pa_lo + pa_loc;
    const int pb_id = rect.pb_lo + (blockIdx.x - pa_loc * rect.pb_cnt);
    const int tile_id = pa_id * bd.B.n_panels + pb_id;
    const PanelDesc pa = bd.A.panels[pa_id];
    const PanelDesc pb = bd.B.panels[pb_id];
    // an active side has one class per panel (G == 1); a side without active orbitals packs G single-string
    // classes, none of which is ever paired
    const int nR = int(pa.G) * int(pa.c), nC = int(pb.G) * int(pb.c);
    const int ld = bd.B.m > 0 ? int(pb.ld) : bd.ld_fixed;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    double* t0 = reinterpret_cast<double*>(smem_raw);
    double* t1 = t0 + (NV == 2 ? size_t(nR) * ld : 0);
    SmemOp* sops = reinterpret_cast<SmemOp*>(t1 + size_t(nR) * ld);
    double* wpart = reinterpret_cast<double*>(sops + bd.nops);  // [nops][nwarps] (NV == 2)
    uint32_t* row_addr = reinterpret_cast<uint32_t*>(wpart + (NV == 2 ? size_t(bd.nops) * nwarps : 0));  // [nR]
    uint32_t* ocol = row_addr + nR;                              // [nC] scatter column addresses (dynamic layout)
    uint16_t* olc = reinterpret_cast<uint16_t*>(ocol + nC);      // [nC] their local columns

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t* addr_a = bd.A.addr + pa.off;
    const uint32_t* cols_sorted = (dyn ? bd.B.dyn_in_sorted : bd.B.addr_sorted) + pb.off;
    const uint16_t* ord_b = (dyn ? bd.B.dyn_in_ord : bd.B.ord) + pb.off;
    for (int i = tid; i < nR; i += nthreads) row_addr[i] = __ldg(addr_a + i);
    if (dyn)
        for (int i = tid; i < nC; i += nthreads) {
            ocol[i] = __ldg(bd.B.dyn_out_sorted + pb.off + i);
            olc[i] = __ldg(bd.B.dyn_out_ord + pb.off + i);
        }
    // each lane owns up to BK_COLCH columns of the tile (ascending global address => coalesced accesses)
    int64_t colg[BK_COLCH];
    int lcs[BK_COLCH];
    const bool reg_cols = nC <= 32 * BK_COLCH;
#pragma unroll
    for (int j = 0; j < BK_COLCH; ++j) {
        const int t = lane + 32 * j;
        colg[j] = 0;
        lcs[j] = -1;
        if (reg_cols && t < nC) {
            colg[j] = __ldg(cols_sorted + t);
            lcs[j] = __ldg(ord_b + t);
        }
    }

    // sub-blocks: an active side may pack G classes into the panel; each (alpha class, beta class) pair of the tile is
    // rotated by its own group of `tps` threads with the pair list of the class pair
    const int Ga = bd.A.m > 0 ? int(pa.G) : 1, Gb = bd.B.m > 0 ? int(pb.G) : 1;
    const int nsub = Ga * Gb;
    const int tps = nthreads / nsub;
    const int sub = tid / tps;
    const int ga = sub / Gb, gb = sub - ga * Gb;
    // threads beyond the last sub-block never have a pair
    const int lid = sub < nsub ? tid - sub * tps : 0x3fffffff;
    const int wlid = __reduce_min_sync(0xffffffffu, lid);
    const int sbase = sub < nsub ? ga * int(pa.c) * ld + gb * int(pb.c) : 0;
    const int sshift_a = ga, sshift_b = 16 + gb;

    auto setup_ops = [&]() {
        for (int t = tid; t < bd.nops; t += nthreads) {
            const BatchOp op = bd.ops[t];
            const int fwd_pos = bd.reversed ? bd.nops - 1 - t : t;
            const int* pi = bd.plist_index + ((size_t(fwd_pos) * (bd.A.m + 1) + int(pa.e)) * (bd.B.m + 1) + int(pb.e)) * 2;
            SmemOp so;
            so.off = pi[0];
            so.cnt = pi[1];
            uint32_t smask = 0;
            if (op.h_a >= 0)
                for (int g = 0; g < Ga; ++g) smask |= (uint32_t(__popc(bd.A.sigma[pa.class_off + g] & op.zspec_a)) & 1u) << g;
            if (op.h_b >= 0)
                for (int g = 0; g < Gb; ++g) smask |= (uint32_t(__popc(bd.B.sigma[pb.class_off + g] & op.zspec_b)) & 1u) << (16 + g);
            const double2 cs = __ldg(rot + op.op_index);
            so.cs = cs.x;
            so.sn = cs.y;
            so.smask = smask;
            so.pad = 0;
            sops[t] = so;
        }
    };
    __syncthreads();  // row_addr
    // ---- gather the tile
    if (reg_cols) {
        // asynchronous global->shared copies (LDGSTS): every element of the tile is in flight before the first
        // wait, so the gather costs about one memory latency instead of one per row
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t ra = int64_t(row_addr[lr]) * nb;
#pragma unroll
            for (int j = 0; j < BK_COLCH; ++j)
                if (lcs[j] >= 0) {
                    cp_async8(t0 + lr * ld + lcs[j], v0 + ra + colg[j]);
                    if (NV == 2) cp_async8(t1 + lr * ld + lcs[j], v1 + ra + colg[j]);
                }
        }
        setup_ops();  // the dependent global loads of the op records overlap the gather
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else {
        setup_ops();
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t row = int64_t(row_addr[lr]) * nb;
            for (int t = lane; t < nC; t += 32) {
                const int lc = __ldg(ord_b + t);
                const int64_t g = row + __ldg(cols_sorted + t);
                t0[lr * ld + lc] = v0[g];
                if (NV == 2) t1[lr * ld + lc] = v1[g];
            }
        }
    }
    __syncthreads();

    const bool rep_a = bd.A.m == 0 && nR > 1, rep_b = bd.B.m == 0 && nC > 1;
    if (rep_a || rep_b) {
        // a side without active orbitals: the same pair list for each of its rows / columns (rare: one-sided
        // batches); not software-pipelined
        for (int t = 0; t < bd.nops; ++t) {
            const SmemOp& op = sops[t];
            double acc = rep_a ? rotate_pairs<NV, true>(t0, t1, op, bd.plist, nR, ld, tid, nthreads)
                               : rotate_pairs<NV, true>(t0, t1, op, bd.plist, nC, 1, tid, nthreads);
            if (NV == 2) {
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
                if (lane == 0) wpart[t * nwarps + warp] = acc;
            }
            __syncthreads();
        }
    } else {
        // Software pipeline: the pair-list entries do not depend on the data, so the entries of the NEXT group of
        // BK_DEPTH excitations are fetched (L2 latency) while the current group is applied; only shared-memory and
        // FP64 latency stays on the barrier-to-barrier critical path.
        const uint32_t* __restrict__ plist = bd.plist;
        uint32_t cur0[BK_DEPTH], cur1[BK_DEPTH], nxt0[BK_DEPTH], nxt1[BK_DEPTH];
        double accs[BK_DEPTH];
        const uint32_t* __restrict__ lp0 = plist + (lid < 0x3fffffff ? lid : 0);
        const uint32_t* __restrict__ lp1 = lp0 + tps;
        double* __restrict__ s0 = t0 + sbase;
        double* __restrict__ s1 = t1 + sbase;
        auto fetch = [&](int tbase, uint32_t* e0, uint32_t* e1) {
#pragma unroll
            for (int j = 0; j < BK_DEPTH; ++j) {
                const int t = tbase + j;
                e0[j] = 0;
                e1[j] = 0;
                if (t < bd.nops) {
                    const int cnt = sops[t].cnt;
                    if (wlid < cnt) {  // warp-uniform: warps without pairs skip the address arithmetic too
                        const int off = sops[t].off;
                        if (lid < cnt) e0[j] = __ldg(lp0 + off);
                        if (lid + tps < cnt) e1[j] = __ldg(lp1 + off);
                    }
                }
            }
        };
        const uint32_t sh0 = uint32_t(__cvta_generic_to_shared(s0));
        const uint32_t sh1 = uint32_t(__cvta_generic_to_shared(s1));
        // sflip: the sub-block's spectator sign of the excitation as an FP64 sign bit; together with the pair's own
        // sign (bit 30 of the entry) it is XORed into the high word of the sine and of the gradient term: no FP64
        // negate / select
        auto rotate_one = [&](uint32_t e, double cs, double sn, uint32_t sflip, double& acc) {
            const uint32_t o1 = (e & 0x7fffu) << 3, o2 = (e >> 12) & 0x3fff8u;
            const uint32_t flip = ((e << 1) & 0x80000000u) ^ sflip;
            const double p = __hiloint2double(__double2hiint(sn) ^ int(flip), __double2loint(sn));
            const double x = lds64(sh0 + o1), y = lds64(sh0 + o2);
            if (NV == 2) {
                const double bx = lds64(sh1 + o1), by = lds64(sh1 + o2);
                const double nbx = cs * bx + p * by, nby = cs * by - p * bx;
                // <bra|G|ket> of the pair, invariant under the common rotation: taken from the values before it
                const double g = bx * y - by * x;
                acc += __hiloint2double(__double2hiint(g) ^ int(flip), __double2loint(g));
                sts64(sh1 + o1, nbx);
                sts64(sh1 + o2, nby);
            }
            const double nx = cs * x + p * y, ny = cs * y - p * x;
            sts64(sh0 + o1, nx);
            sts64(sh0 + o2, ny);
        };
        fetch(0, cur0, cur1);
#ifdef TCC_DEBUG_SKIP_OPS  // diagnostic build: gather / scatter and per-tile set-up only
        const int nops_loop = 0;
#else
        const int nops_loop = bd.nops;
#endif
        for (int tb = 0; tb < nops_loop; tb += BK_DEPTH) {
            fetch(tb + BK_DEPTH, nxt0, nxt1);
#pragma unroll
            for (int j = 0; j < BK_DEPTH; ++j) {
                const int t = tb + j;
                accs[j] = 0.0;
                if (t < bd.nops) {
                    const int cnt = sops[t].cnt;
                    if (lid < cnt) {  // warps without a pair for this excitation go straight to the barrier
                        const double cs = sops[t].cs, sn = sops[t].sn;
                        const uint32_t sm = sops[t].smask;
                        const uint32_t sgn = ((sm >> sshift_a) ^ (sm >> sshift_b)) << 31;
                        double acc = 0.0;
                        rotate_one(cur0[j], cs, sn, sgn, acc);
                        if (lid + tps < cnt) {
                            rotate_one(cur1[j], cs, sn, sgn, acc);
                            const int off = sops[t].off;
                            for (int idx = lid + 2 * tps; idx < cnt; idx += tps) rotate_one(__ldg(plist + off + idx), cs, sn, sgn, acc);
                        }
                        accs[j] = acc;
                    }
                    __syncthreads();
                }
            }
            if (NV == 2) {
                // the BK_DEPTH warp reductions of the group run interleaved, off the barrier-to-barrier path
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int j = 0; j < BK_DEPTH; ++j) accs[j] += __shfl_down_sync(0xffffffffu, accs[j], o);
                }
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < BK_DEPTH; ++j)
                        if (tb + j < bd.nops) wpart[(tb + j) * nwarps + warp] = accs[j];
                }
            }
#pragma unroll
            for (int j = 0; j < BK_DEPTH; ++j) {
                cur0[j] = nxt0[j];
                cur1[j] = nxt1[j];
            }
        }
    }

    // ---- scatter the tile back
    if (dyn) {
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            double* dst = vout + int64_t(blockIdx.y) * ss.vec + int64_t(row_addr[lr]) * nb;
            for (int t = lane; t < nC; t += 32) dst[ocol[t]] = t0[lr * ld + olc[t]];
        }
    } else if (reg_cols) {
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            double* dst0 = v0 + int64_t(row_addr[lr]) * nb;
            double* dst1 = NV == 2 ? v1 + int64_t(row_addr[lr]) * nb : nullptr;
#pragma unroll
            for (int j = 0; j < BK_COLCH; ++j)
                if (lcs[j] >= 0) {
                    dst0[colg[j]] = t0[lr * ld + lcs[j]];
                    if (NV == 2) dst1[colg[j]] = t1[lr * ld + lcs[j]];
                }
        }
    } else {
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t row = int64_t(row_addr[lr]) * nb;
            for (int t = lane; t < nC; t += 32) {
                const int lc = __ldg(ord_b + t);
                const int64_t g = row + __ldg(cols_sorted + t);
                v0[g] = t0[lr * ld + lc];
                if (NV == 2) v1[g] = t1[lr * ld + lc];
            }
        }
    }
    if (NV == 2) {
        __syncthreads();
        for (int t = tid; t < bd.nops; t += nthreads) {
            double s = 0.0;
            for (int w = 0; w < nwarps; ++w) s += wpart[t * nwarps + w];
            partials[int64_t(t) * n_tiles + tile_id] = s;
        }
    }
}

// ================================================================================================================
// Lean batch kernels (batches with active orbitals on both sides).  Same tiles, same arithmetic and the same order of
// excitations as batch_kernel above; what changes is the work per excitation OUTSIDE the rotation itself:
//   * the pairs come from a per-thread stream (Batch::stream8 / stream16): thread `lid` of a sub-block applies entry
//     [slot * tps + lid] of each of the excitation's `slots` rounds -- no per-thread count tests, no offset arithmetic;
//   * the excitation record in shared memory is 16 bytes (cos, sin with the tile's spectator sign folded in) read by
//     one broadcast 128-bit load, plus 8 bytes (stream offset, slots) read only by the prefetch;
//   * NV == 3: ket and bra are INTERLEAVED element by element ((ket, bra) = one 16-byte element) in HBM and in shared
//     memory: one 128-bit gather / scatter per element and one 128-bit shared-memory access per pair end instead of
//     two 64-bit ones; NV == 2 keeps them in two vectors (needed where no memory is left for an interleaved copy);
//   * the four gradient partials of a prefetch group are reduced over the warp by a transposed butterfly (12 shuffles
//     instead of 40).
// ================================================================================================================
constexpr int LK_DEPTH = 4;
constexpr uint32_t LK_VALID = 0x80000000u;
#ifndef LK_FWD_MINB
#define LK_FWD_MINB 4  // resident CTAs per SM the forward kernel is compiled for (register cap 65536 / (256 * MINB))
#endif

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = unsigned(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ uint2 lds64u(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double flip_sign(double v, uint32_t flip) {
    return __hiloint2double(__double2hiint(v) ^ int(flip), __double2loint(v));
}

template <int NV>
__global__ void __launch_bounds__(BK_THREADS, NV == 1 ? LK_FWD_MINB : 2)
lean_batch_kernel(double* __restrict__ v0, double* __restrict__ v1, double* __restrict__ vout, int64_t nb, BatchDev bd,
                  const double2* __restrict__ rot, double* __restrict__ partials, int n_tiles, SetStrides ss, TileRect rect) {
    // vout != null (NV == 1 or 3): dynamic column layout, see batch_kernel
    const bool dyn = NV != 2 && vout != nullptr;
    constexpr int NVEC = NV == 1 ? 1 : 2;        // vectors rotated
    constexpr int ELW = NV == 3 ? 16 : 8;        // bytes of a tile element in shared memory
    constexpr int ESH = NV == 3 ? 4 : 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    v0 += int64_t(blockIdx.y) * ss.vec;
    if (NV == 2) v1 += int64_t(blockIdx.y) * ss.vec;
    rot += int64_t(blockIdx.y) * ss.rot;
    if (NVEC == 2) partials += int64_t(blockIdx.y) * ss.part;
    const int pa_loc = blockIdx.x / rect.pb_cnt;
    const int pa_id = rect.pa_lo + pa_loc;
    const int pb_id = rect.pb_lo + (blockIdx.x - pa_loc * rect.pb_cnt);
    const int tile_id = pa_id * bd.B.n_panels + pb_id;
    const PanelDesc pa = bd.A.panels[pa_id];
    const PanelDesc pb = bd.B.panels[pb_id];
    const int nR = int(pa.G) * int(pa.c), nC = int(pb.G) * int(pb.c);
    const int ld = int(pb.ld);
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int nops = bd.nops;
    // shared memory: tile(s) | records [nops] (cos, sin) | warp partials [nops][nwarps] | meta [nops] (offset, slots) |
    // spectator masks [nops] | row addresses [nR]
    const size_t tile_bytes = (size_t(nR) * ld * ELW + 15) & ~size_t(15);
    unsigned char* t0 = smem_raw;
    unsigned char* t1 = t0 + (NV == 2 ? tile_bytes : 0);
    double2* recs = reinterpret_cast<double2*>(t1 + tile_bytes);
    double* wpart = reinterpret_cast<double*>(recs + nops);
    uint2* metas = reinterpret_cast<uint2*>(wpart + (NVEC == 2 ? size_t(nops) * nwarps : 0));
    uint32_t* smasks = reinterpret_cast<uint32_t*>(metas + nops);
    uint32_t* row_addr = smasks + nops;
    uint32_t* ocol = row_addr + nR;                          // [nC] scatter column addresses (dynamic layout)
    uint16_t* olc = reinterpret_cast<uint16_t*>(ocol + nC);  // [nC] their local columns

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t* addr_a = bd.A.addr + pa.off;
    const uint32_t* cols_sorted = (dyn ? bd.B.dyn_in_sorted : bd.B.addr_sorted) + pb.off;
    const uint16_t* ord_b = (dyn ? bd.B.dyn_in_ord : bd.B.ord) + pb.off;
    for (int i = tid; i < nR; i += nthreads) row_addr[i] = __ldg(addr_a + i);
    if (dyn)
        for (int i = tid; i < nC; i += nthreads) {
            ocol[i] = __ldg(bd.B.dyn_out_sorted + pb.off + i);
            olc[i] = __ldg(bd.B.dyn_out_ord + pb.off + i);
        }
    int64_t colg[BK_COLCH];
    int lcs[BK_COLCH];
    const bool reg_cols = nC <= 32 * BK_COLCH;
#pragma unroll
    for (int j = 0; j < BK_COLCH; ++j) {
        const int t = lane + 32 * j;
        colg[j] = 0;
        lcs[j] = -1;
        if (reg_cols && t < nC) {
            colg[j] = __ldg(cols_sorted + t);
            lcs[j] = __ldg(ord_b + t);
        }
    }
    // sub-blocks on the NOMINAL class grid of the panel pair (the streams are laid out for it); a last, partly filled
    // panel leaves the threads of its missing classes idle
    const int Gan = max(1, int(pa.Gn)), Gbn = max(1, int(pb.Gn));
    const int nsub = Gan * Gbn;
    const int tps = nthreads / nsub;
    const int sub = tid / tps;
    const int ga = sub / Gbn, gb = sub - ga * Gbn;
    const bool active = sub < nsub && ga < int(pa.G) && gb < int(pb.G);
    const int lid = tid - sub * tps;
    const int sbase = active ? ga * int(pa.c) * ld + gb * int(pb.c) : 0;
    const int sshift_a = ga, sshift_b = 16 + gb;
    const bool multi = nsub > 1;

    auto setup_ops = [&]() {
        for (int t = tid; t < nops; t += nthreads) {
            const BatchOp op = bd.ops[t];
            const int fwd_pos = bd.reversed ? nops - 1 - t : t;
            const int* si = bd.stream_index + ((size_t(fwd_pos) * (bd.A.m + 1) + int(pa.e)) * (bd.B.m + 1) + int(pb.e)) * 2;
            metas[t] = make_uint2(uint32_t(si[0]), uint32_t(si[1]));
            uint32_t smask = 0;
            if (op.h_a >= 0)
                for (int g = 0; g < int(pa.G); ++g) smask |= (uint32_t(__popc(bd.A.sigma[pa.class_off + g] & op.zspec_a)) & 1u) << g;
            if (op.h_b >= 0)
                for (int g = 0; g < int(pb.G); ++g) smask |= (uint32_t(__popc(bd.B.sigma[pb.class_off + g] & op.zspec_b)) & 1u) << (16 + g);
            const double2 cs = __ldg(rot + op.op_index);
            // one class per side: the spectator sign of the tile is folded into the sine
            const bool fold = !multi && (((smask ^ (smask >> 16)) & 1u) != 0);
            recs[t] = make_double2(cs.x, fold ? -cs.y : cs.y);
            smasks[t] = smask;
        }
    };
    __syncthreads();  // row_addr
    // ---- gather the tile (asynchronous copies: every element is in flight before the first wait)
    if (reg_cols) {
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t ra = int64_t(row_addr[lr]) * nb;
#pragma unroll
            for (int j = 0; j < BK_COLCH; ++j)
                if (lcs[j] >= 0) {
                    if (NV == 3) {
                        cp_async16(t0 + (size_t(lr) * ld + lcs[j]) * 16, v0 + (ra + colg[j]) * 2);
                    } else {
                        cp_async8(reinterpret_cast<double*>(t0) + lr * ld + lcs[j], v0 + ra + colg[j]);
                        if (NV == 2) cp_async8(reinterpret_cast<double*>(t1) + lr * ld + lcs[j], v1 + ra + colg[j]);
                    }
                }
        }
        setup_ops();
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else {
        setup_ops();
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t row = int64_t(row_addr[lr]) * nb;
            for (int t = lane; t < nC; t += 32) {
                const int lc = __ldg(ord_b + t);
                const int64_t g = row + __ldg(cols_sorted + t);
                if (NV == 3) {
                    reinterpret_cast<double2*>(t0)[lr * ld + lc] = reinterpret_cast<const double2*>(v0)[g];
                } else {
                    reinterpret_cast<double*>(t0)[lr * ld + lc] = v0[g];
                    if (NV == 2) reinterpret_cast<double*>(t1)[lr * ld + lc] = v1[g];
                }
            }
        }
    }
    __syncthreads();

    // ---- the excitations of the batch, in order
    {
        const uint32_t* __restrict__ sp = (NV == 3 ? bd.stream16 : bd.stream8) + lid;
        const uint32_t sh0 = uint32_t(__cvta_generic_to_shared(t0)) + (uint32_t(sbase) << ESH);
        const uint32_t sh1 = uint32_t(__cvta_generic_to_shared(t1)) + (uint32_t(sbase) << ESH);
        const uint32_t rec_s = uint32_t(__cvta_generic_to_shared(recs));
        const uint32_t meta_s = uint32_t(__cvta_generic_to_shared(metas));
        const uint32_t mask_s = uint32_t(__cvta_generic_to_shared(smasks));
        uint32_t cur0[LK_DEPTH], cur1[LK_DEPTH], nxt0[LK_DEPTH], nxt1[LK_DEPTH];
        double accs[LK_DEPTH];
        auto fetch = [&](int tbase, uint32_t* e0, uint32_t* e1) {
#pragma unroll
            for (int j = 0; j < LK_DEPTH; ++j) {
                const int t = tbase + j;
                e0[j] = 0;
                e1[j] = 0;
                if (t < nops && active) {
                    const uint2 m = lds64u(meta_s + t * 8);
                    if (m.y > 0) e0[j] = __ldg(sp + m.x);
                    if (m.y > 1) e1[j] = __ldg(sp + m.x + tps);
                }
            }
        };
        // flip = sign of the pair (bit 30 of the entry) x spectator sign of the sub-block (packed tiles only), as an FP64
        // sign bit XORed into the sine and into the gradient term: no FP64 negate / select
        auto pair = [&](uint32_t e, double cs, double sn, uint32_t sflip, double& acc) {
            const uint32_t flip = ((e << 1) & 0x80000000u) ^ sflip;
            const double p = flip_sign(sn, flip);
            if (NV == 3) {
                const uint32_t o1 = (e & 0x7fffu) << 4, o2 = (e >> 11) & 0x7fff0u;
                const double2 a = lds128(sh0 + o1), b = lds128(sh0 + o2);  // (ket, bra) of either end
                // <bra|G|ket> of the pair is invariant under the common rotation: taken from the values before it
                acc = fma(flip_sign(a.y, flip), b.x, acc);
                acc = fma(-flip_sign(b.y, flip), a.x, acc);
                sts128(sh0 + o1, cs * a.x + p * b.x, cs * a.y + p * b.y);
                sts128(sh0 + o2, cs * b.x - p * a.x, cs * b.y - p * a.y);
            } else {
                const uint32_t o1 = (e & 0x7fffu) << 3, o2 = (e >> 12) & 0x3fff8u;
                const double x = lds64(sh0 + o1), y = lds64(sh0 + o2);
                if (NV == 2) {
                    const double bx = lds64(sh1 + o1), by = lds64(sh1 + o2);
                    acc = fma(flip_sign(bx, flip), y, acc);
                    acc = fma(-flip_sign(by, flip), x, acc);
                    sts64(sh1 + o1, cs * bx + p * by);
                    sts64(sh1 + o2, cs * by - p * bx);
                }
                sts64(sh0 + o1, cs * x + p * y);
                sts64(sh0 + o2, cs * y - p * x);
            }
        };
        fetch(0, cur0, cur1);
#ifdef TCC_DEBUG_SKIP_OPS  // diagnostic build: gather / scatter and per-tile set-up only
        const int nops_loop = 0;
#else
        const int nops_loop = nops;
#endif
        for (int tb = 0; tb < nops_loop; tb += LK_DEPTH) {
            fetch(tb + LK_DEPTH, nxt0, nxt1);
#pragma unroll
            for (int j = 0; j < LK_DEPTH; ++j) {
                const int t = tb + j;
                accs[j] = 0.0;
                if (t < nops) {
                    const uint32_t e0 = cur0[j], e1 = cur1[j];
                    if (e0 & LK_VALID) {  // an entry in round 1 implies one in round 0
                        const double2 r = lds128(rec_s + t * 16);
                        const double sn = r.y;
                        uint32_t sflip = 0;
                        if (multi) {
                            const uint32_t sm = lds32u(mask_s + t * 4);
                            sflip = ((sm >> sshift_a) ^ (sm >> sshift_b)) << 31;
                        }
                        double acc = 0.0;
                        pair(e0, r.x, sn, sflip, acc);
                        if (e1 & LK_VALID) {
                            pair(e1, r.x, sn, sflip, acc);
                            const uint2 m = lds64u(meta_s + t * 8);
                            for (uint32_t s_ = 2; s_ < m.y; ++s_) {
                                const uint32_t e = __ldg(sp + m.x + s_ * tps);
                                if (e & LK_VALID) pair(e, r.x, sn, sflip, acc);
                            }
                        }
                        accs[j] = acc;
                    }
                    __syncthreads();
                }
            }
            if (NVEC == 2) {
                // transposed butterfly: after two exchange steps every lane holds ONE partial of ONE excitation of the
                // group (excitation 2 * bit4(lane) + bit3(lane)); three plain steps finish the sum
                static_assert(LK_DEPTH == 4, "the reduction below is written for groups of four");
                const bool hi16 = lane & 16, hi8 = lane & 8;
                const double s0 = hi16 ? accs[0] : accs[2], s1 = hi16 ? accs[1] : accs[3];
                double k0 = (hi16 ? accs[2] : accs[0]) + __shfl_xor_sync(0xffffffffu, s0, 16);
                double k1 = (hi16 ? accs[3] : accs[1]) + __shfl_xor_sync(0xffffffffu, s1, 16);
                double k = (hi8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
                k += __shfl_xor_sync(0xffffffffu, k, 4);
                k += __shfl_xor_sync(0xffffffffu, k, 2);
                k += __shfl_xor_sync(0xffffffffu, k, 1);
                if ((lane & 7) == 0) {
                    const int j = ((lane >> 4) << 1) | ((lane >> 3) & 1);
                    if (tb + j < nops) wpart[(tb + j) * nwarps + warp] = k;
                }
            }
#pragma unroll
            for (int j = 0; j < LK_DEPTH; ++j) {
                cur0[j] = nxt0[j];
                cur1[j] = nxt1[j];
            }
        }
    }

    // ---- scatter the tile back
    if (dyn) {
        double* outv = vout + int64_t(blockIdx.y) * ss.vec;
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t ra = int64_t(row_addr[lr]) * nb;
            for (int t = lane; t < nC; t += 32) {
                if (NV == 3)
                    reinterpret_cast<double2*>(outv)[ra + ocol[t]] = reinterpret_cast<const double2*>(t0)[lr * ld + olc[t]];
                else
                    outv[ra + ocol[t]] = reinterpret_cast<const double*>(t0)[lr * ld + olc[t]];
            }
        }
    } else if (reg_cols) {
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t ra = int64_t(row_addr[lr]) * nb;
#pragma unroll
            for (int j = 0; j < BK_COLCH; ++j)
                if (lcs[j] >= 0) {
                    if (NV == 3) {
                        reinterpret_cast<double2*>(v0)[ra + colg[j]] = reinterpret_cast<const double2*>(t0)[lr * ld + lcs[j]];
                    } else {
                        v0[ra + colg[j]] = reinterpret_cast<const double*>(t0)[lr * ld + lcs[j]];
                        if (NV == 2) v1[ra + colg[j]] = reinterpret_cast<const double*>(t1)[lr * ld + lcs[j]];
                    }
                }
        }
    } else {
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t row = int64_t(row_addr[lr]) * nb;
            for (int t = lane; t < nC; t += 32) {
                const int lc = __ldg(ord_b + t);
                const int64_t g = row + __ldg(cols_sorted + t);
                if (NV == 3) {
                    reinterpret_cast<double2*>(v0)[g] = reinterpret_cast<const double2*>(t0)[lr * ld + lc];
                } else {
                    v0[g] = reinterpret_cast<const double*>(t0)[lr * ld + lc];
                    if (NV == 2) v1[g] = reinterpret_cast<const double*>(t1)[lr * ld + lc];
                }
            }
        }
    }
    if (NVEC == 2) {
        __syncthreads();
        for (int t = tid; t < nops; t += nthreads) {
            double s = 0.0;
            for (int w = 0; w < nwarps; ++w) s += wpart[t * nwarps + w];
            // one class per side: the tile's spectator sign was folded into the sine; the gradient term takes it here
            if (!multi && (((smasks[t] ^ (smasks[t] >> 16)) & 1u) != 0)) s = -s;
            partials[int64_t(t) * n_tiles + tile_id] = s;
        }
    }
}

// ================================================================================================================
// Persistent lean kernel.  Same tiles, streams and arithmetic as lean_batch_kernel; what changes is the life of a CTA.
// ncu on the one-tile-per-CTA kernels (profiles/r02_batch_kernels_16o_ncu.txt) shows that a tile spends more time
// waiting than moving data: the CTA start-up is a chain of dependent L2 round trips (panel descriptors -> row / column
// addresses -> tile gather; excitation records -> angles; stream offsets -> first stream entries), none of which can be
// hidden with two CTAs per SM, and every tile pays the CTA launch itself.  Here a CTA stays resident and takes tiles
// from a ticket counter; while the gather of tile k is in flight it resolves EVERYTHING tile k+1 needs except the
// data (descriptors, addresses, records, masks -- double-buffered in a few KB of shared memory), and the ticket of
// tile k+2.  Per tile the critical path is: gather -> excitations -> scatter.
// NV == 1: forward sweep; NV == 3: reverse sweep on interleaved (ket, bra) elements.  One parameter set, static column
// order, tiles of at most 32 * BK_COLCH columns (the launchers fall back to lean_batch_kernel otherwise).
// ================================================================================================================
struct PersistTile {  // what the op loop of a tile needs, resolved one tile ahead
    PanelDesc pa, pb;
    int tile_id;  // index into the per-tile gradient partials; -1 = no tile
    int pad[3];
};

template <int NV>
__global__ void __launch_bounds__(BK_THREADS, NV == 1 ? LK_FWD_MINB : 2)
lean_persist_kernel(double* __restrict__ v0, int64_t nb, BatchDev bd, const double2* __restrict__ rot, double* __restrict__ partials,
                    int n_tiles, TileRect rect, unsigned* __restrict__ ctl, int max_rows, int max_cols) {
    constexpr int NVEC = NV == 1 ? 1 : 2;
    constexpr int ELW = NV == 3 ? 16 : 8;
    constexpr int ESH = NV == 3 ? 4 : 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int nops = bd.nops;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rect_tiles = rect.pa_cnt * rect.pb_cnt;
    // shared memory: tile | records [2][nops] | warp partials [nops][nwarps] | meta [2][nops] | masks [2][nops] |
    // row addresses [2][max_rows] | column addresses [2][max_cols] | local columns [2][max_cols] | tile records [2] | tickets [2]
    const size_t tile_bytes = (size_t(max_rows) * (size_t(max_cols) | 1) * ELW + 15) & ~size_t(15);
    unsigned char* t0 = smem_raw;
    double2* recs = reinterpret_cast<double2*>(t0 + tile_bytes);
    double* wpart = reinterpret_cast<double*>(recs + 2 * nops);
    uint2* metas = reinterpret_cast<uint2*>(wpart + (NVEC == 2 ? size_t(nops) * nwarps : 0));
    uint32_t* smasks = reinterpret_cast<uint32_t*>(metas + 2 * nops);
    uint32_t* row_addr = smasks + 2 * nops;
    uint32_t* col_g = row_addr + 2 * max_rows;
    PersistTile* tiles = reinterpret_cast<PersistTile*>((reinterpret_cast<uintptr_t>(col_g + 2 * max_cols) + 15) & ~uintptr_t(15));
    int* tickets = reinterpret_cast<int*>(tiles + 2);
    uint16_t* col_l = reinterpret_cast<uint16_t*>(tickets + 4);

    // resolves tile `lin` of the rectangle into buffer `buf`: two rounds of independent global loads
    auto resolve = [&](int lin, int buf) {
        if (lin >= rect_tiles) {
            if (tid == 0) tiles[buf].tile_id = -1;
            return;
        }
        const int pa_loc = lin / rect.pb_cnt;
        const int pa_id = rect.pa_lo + pa_loc;
        const int pb_id = rect.pb_lo + (lin - pa_loc * rect.pb_cnt);
        const PanelDesc pa = bd.A.panels[pa_id];
        const PanelDesc pb = bd.B.panels[pb_id];
        BatchOp op0;
        if (tid < nops) op0 = bd.ops[tid];  // in flight together with the descriptors
        const int nR = int(pa.G) * int(pa.c), nC = int(pb.G) * int(pb.c);
        for (int i = tid; i < nR; i += nthreads) row_addr[buf * max_rows + i] = __ldg(bd.A.addr + pa.off + i);
        for (int i = tid; i < nC; i += nthreads) {
            col_g[buf * max_cols + i] = __ldg(bd.B.addr_sorted + pb.off + i);
            col_l[buf * max_cols + i] = __ldg(bd.B.ord + pb.off + i);
        }
        const bool multi = max(1, int(pa.Gn)) * max(1, int(pb.Gn)) > 1;
        for (int t = tid; t < nops; t += nthreads) {
            const BatchOp op = t == tid ? op0 : bd.ops[t];
            const int fwd_pos = bd.reversed ? nops - 1 - t : t;
            const int* si = bd.stream_index + ((size_t(fwd_pos) * (bd.A.m + 1) + int(pa.e)) * (bd.B.m + 1) + int(pb.e)) * 2;
            const int si0 = __ldg(si), si1 = __ldg(si + 1);
            const double2 cs = __ldg(rot + op.op_index);
            uint32_t smask = 0;
            if (op.h_a >= 0)
                for (int g = 0; g < int(pa.G); ++g) smask |= (uint32_t(__popc(__ldg(bd.A.sigma + pa.class_off + g) & op.zspec_a)) & 1u) << g;
            if (op.h_b >= 0)
                for (int g = 0; g < int(pb.G); ++g) smask |= (uint32_t(__popc(__ldg(bd.B.sigma + pb.class_off + g) & op.zspec_b)) & 1u) << (16 + g);
            const bool fold = !multi && (((smask ^ (smask >> 16)) & 1u) != 0);
            recs[buf * nops + t] = make_double2(cs.x, fold ? -cs.y : cs.y);
            metas[buf * nops + t] = make_uint2(uint32_t(si0), uint32_t(si1));
            smasks[buf * nops + t] = smask;
        }
        if (tid == 0) {
            tiles[buf].pa = pa;
            tiles[buf].pb = pb;
            tiles[buf].tile_id = pa_id * bd.B.n_panels + pb_id;
        }
    };

    if (tid == 0) {
        tickets[0] = int(atomicAdd(ctl, 1u));
        tickets[1] = int(atomicAdd(ctl, 1u));
    }
    __syncthreads();
    int lin_next = tickets[1];
    resolve(tickets[0], 0);
    __syncthreads();

    const uint32_t* __restrict__ stream = NV == 3 ? bd.stream16 : bd.stream8;
    const uint32_t t0_s = uint32_t(__cvta_generic_to_shared(t0));
    for (int cur = 0;; cur ^= 1) {
        const int tile_id = tiles[cur].tile_id;
        if (tile_id < 0) break;
        const PanelDesc pa = tiles[cur].pa, pb = tiles[cur].pb;
        const int nR = int(pa.G) * int(pa.c), nC = int(pb.G) * int(pb.c);
        const int ld = int(pb.ld);
        const int Gan = max(1, int(pa.Gn)), Gbn = max(1, int(pb.Gn));
        const int nsub = Gan * Gbn;
        const int tps = nthreads / nsub;
        const int sub = tid / tps;
        const int ga = sub / Gbn, gb = sub - ga * Gbn;
        const bool active = sub < nsub && ga < int(pa.G) && gb < int(pb.G);
        const int lid = tid - sub * tps;
        const int sbase = active ? ga * int(pa.c) * ld + gb * int(pb.c) : 0;
        const int sshift_a = ga, sshift_b = 16 + gb;
        const bool multi = nsub > 1;
        const uint32_t* ra_s = row_addr + cur * max_rows;

        // ---- gather (asynchronous copies); the lane's columns stay in registers for the scatter
        int64_t colg[BK_COLCH];
        int lcs[BK_COLCH];
#pragma unroll
        for (int j = 0; j < BK_COLCH; ++j) {
            const int t = lane + 32 * j;
            colg[j] = 0;
            lcs[j] = -1;
            if (t < nC) {
                colg[j] = col_g[cur * max_cols + t];
                lcs[j] = col_l[cur * max_cols + t];
            }
        }
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t ra = int64_t(ra_s[lr]) * nb;
#pragma unroll
            for (int j = 0; j < BK_COLCH; ++j)
                if (lcs[j] >= 0) {
                    if (NV == 3)
                        cp_async16(t0 + (size_t(lr) * ld + lcs[j]) * 16, v0 + (ra + colg[j]) * 2);
                    else
                        cp_async8(reinterpret_cast<double*>(t0) + lr * ld + lcs[j], v0 + ra + colg[j]);
                }
        }
        // ---- while the gather is in flight: first stream entries of this tile, everything of the next tile, next ticket
        const uint32_t* __restrict__ sp = stream + lid;
        const uint32_t sh0 = t0_s + (uint32_t(sbase) << ESH);
        const uint32_t rec_s = uint32_t(__cvta_generic_to_shared(recs + cur * nops));
        const uint32_t meta_s = uint32_t(__cvta_generic_to_shared(metas + cur * nops));
        const uint32_t mask_s = uint32_t(__cvta_generic_to_shared(smasks + cur * nops));
        uint32_t cur0[LK_DEPTH], cur1[LK_DEPTH], nxt0[LK_DEPTH], nxt1[LK_DEPTH];
        double accs[LK_DEPTH];
        auto fetch = [&](int tbase, uint32_t* e0, uint32_t* e1) {
#pragma unroll
            for (int j = 0; j < LK_DEPTH; ++j) {
                const int t = tbase + j;
                e0[j] = 0;
                e1[j] = 0;
                if (t < nops && active) {
                    const uint2 m = lds64u(meta_s + t * 8);
                    if (m.y > 0) e0[j] = __ldg(sp + m.x);
                    if (m.y > 1) e1[j] = __ldg(sp + m.x + tps);
                }
            }
        };
        auto pair = [&](uint32_t e, double cs, double sn, uint32_t sflip, double& acc) {
            const uint32_t flip = ((e << 1) & 0x80000000u) ^ sflip;
            const double p = flip_sign(sn, flip);
            if (NV == 3) {
                const uint32_t o1 = (e & 0x7fffu) << 4, o2 = (e >> 11) & 0x7fff0u;
                const double2 a = lds128(sh0 + o1), b = lds128(sh0 + o2);
                acc = fma(flip_sign(a.y, flip), b.x, acc);
                acc = fma(-flip_sign(b.y, flip), a.x, acc);
                sts128(sh0 + o1, cs * a.x + p * b.x, cs * a.y + p * b.y);
                sts128(sh0 + o2, cs * b.x - p * a.x, cs * b.y - p * a.y);
            } else {
                const uint32_t o1 = (e & 0x7fffu) << 3, o2 = (e >> 12) & 0x3fff8u;
                const double x = lds64(sh0 + o1), y = lds64(sh0 + o2);
                sts64(sh0 + o1, cs * x + p * y);
                sts64(sh0 + o2, cs * y - p * x);
            }
        };
        fetch(0, cur0, cur1);
        int lin_after = 0x7fffffff;
        if (tid == 0 && lin_next < rect_tiles) lin_after = int(atomicAdd(ctl, 1u));
        resolve(lin_next, cur ^ 1);
        if (tid == 0) tickets[2 + cur] = lin_after;
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();
        lin_next = tickets[2 + cur];

        // ---- the excitations of the batch, in order
#ifdef TCC_DEBUG_SKIP_OPS
        const int nops_loop = 0;
#else
        const int nops_loop = nops;
#endif
        for (int tb = 0; tb < nops_loop; tb += LK_DEPTH) {
            fetch(tb + LK_DEPTH, nxt0, nxt1);
#pragma unroll
            for (int j = 0; j < LK_DEPTH; ++j) {
                const int t = tb + j;
                accs[j] = 0.0;
                if (t < nops) {
                    const uint32_t e0 = cur0[j], e1 = cur1[j];
                    if (e0 & LK_VALID) {
                        const double2 r = lds128(rec_s + t * 16);
                        const double sn = r.y;
                        uint32_t sflip = 0;
                        if (multi) {
                            const uint32_t sm = lds32u(mask_s + t * 4);
                            sflip = ((sm >> sshift_a) ^ (sm >> sshift_b)) << 31;
                        }
                        double acc = 0.0;
                        pair(e0, r.x, sn, sflip, acc);
                        if (e1 & LK_VALID) {
                            pair(e1, r.x, sn, sflip, acc);
                            const uint2 m = lds64u(meta_s + t * 8);
                            for (uint32_t s_ = 2; s_ < m.y; ++s_) {
                                const uint32_t e = __ldg(sp + m.x + s_ * tps);
                                if (e & LK_VALID) pair(e, r.x, sn, sflip, acc);
                            }
                        }
                        accs[j] = acc;
                    }
                    __syncthreads();
                }
            }
            if (NVEC == 2) {
                static_assert(LK_DEPTH == 4, "the reduction below is written for groups of four");
                const bool hi16 = lane & 16, hi8 = lane & 8;
                const double s0 = hi16 ? accs[0] : accs[2], s1 = hi16 ? accs[1] : accs[3];
                double k0 = (hi16 ? accs[2] : accs[0]) + __shfl_xor_sync(0xffffffffu, s0, 16);
                double k1 = (hi16 ? accs[3] : accs[1]) + __shfl_xor_sync(0xffffffffu, s1, 16);
                double k = (hi8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
                k += __shfl_xor_sync(0xffffffffu, k, 4);
                k += __shfl_xor_sync(0xffffffffu, k, 2);
                k += __shfl_xor_sync(0xffffffffu, k, 1);
                if ((lane & 7) == 0) {
                    const int j = ((lane >> 4) << 1) | ((lane >> 3) & 1);
                    if (tb + j < nops) wpart[(tb + j) * nwarps + warp] = k;
                }
            }
#pragma unroll
            for (int j = 0; j < LK_DEPTH; ++j) {
                cur0[j] = nxt0[j];
                cur1[j] = nxt1[j];
            }
        }

        // ---- scatter the tile back
        for (int lr = warp; lr < TCC_MEM_ROWS(nR); lr += nwarps) {
            const int64_t ra = int64_t(ra_s[lr]) * nb;
#pragma unroll
            for (int j = 0; j < BK_COLCH; ++j)
                if (lcs[j] >= 0) {
                    if (NV == 3)
                        reinterpret_cast<double2*>(v0)[ra + colg[j]] = reinterpret_cast<const double2*>(t0)[lr * ld + lcs[j]];
                    else
                        v0[ra + colg[j]] = reinterpret_cast<const double*>(t0)[lr * ld + lcs[j]];
                }
        }
        __syncthreads();  // the tile buffer is free; the warp partials are complete
        if (NVEC == 2) {
            for (int t = tid; t < nops; t += nthreads) {
                double s = 0.0;
                for (int w = 0; w < nwarps; ++w) s += wpart[t * nwarps + w];
                const uint32_t sm = smasks[cur * nops + t];
                if (!multi && (((sm ^ (sm >> 16)) & 1u) != 0)) s = -s;
                partials[int64_t(t) * n_tiles + tile_id] = s;
            }
            __syncthreads();  // before the next tile overwrites the partials or this tile's masks
        }
    }
    // the last CTA to leave resets the ticket counter for the next launch that uses this slot
    if (tid == 0) {
        __threadfence();
        const unsigned d = atomicAdd(ctl + 1, 1u);
        if (d == gridDim.x - 1) {
            ctl[0] = 0;
            ctl[1] = 0;
            __threadfence();
        }
    }
}

size_t lean_persist_smem_bytes(int max_rows, int max_cols, int nops, int nv) {
    const size_t elw = nv == 3 ? 16 : 8;
    size_t bytes = (size_t(max_rows) * (size_t(max_cols) | 1) * elw + 15) & ~size_t(15);
    bytes += 2 * size_t(nops) * 16;                                    // records
    if (nv >= 2) bytes += size_t(nops) * BK_WARPS * sizeof(double);   // warp partials
    bytes += 2 * size_t(nops) * 8 + 2 * size_t(nops) * 4;              // meta, masks
    bytes += 2 * size_t(max_rows) * 4 + 2 * size_t(max_cols) * 4;      // row / column addresses
    bytes += 16 + 2 * sizeof(PersistTile) + 16;                        // alignment, tile records, tickets
    bytes += 2 * size_t(max_cols) * 2;                                 // local columns
    return bytes + 16;
}

size_t lean_smem_bytes(int max_rows, int max_cols, int nops, int nv) {
    const size_t ld = size_t(max_cols) | 1;
    const size_t elw = nv == 3 ? 16 : 8;
    size_t bytes = ((size_t(max_rows) * ld * elw + 15) & ~size_t(15)) * (nv == 2 ? 2 : 1);
    bytes += size_t(nops) * 16;                                // records
    if (nv >= 2) bytes += size_t(nops) * BK_WARPS * sizeof(double);  // warp partials
    bytes += size_t(nops) * 8 + size_t(nops) * 4;                 // meta, spectator masks
    bytes += size_t(max_rows) * sizeof(uint32_t);
    bytes += size_t(max_cols) * 6 + 8;  // scatter tables of the dynamic column layout
    return bytes + 16;
}

// (ket, bra) <-> interleaved [N][2]
__global__ void __launch_bounds__(256) interleave_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                         double2* __restrict__ out, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256) out[i] = make_double2(a[i], b[i]);
}
__global__ void __launch_bounds__(256) deinterleave_kernel(const double2* __restrict__ in, double* __restrict__ a,
                                                           double* __restrict__ b, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256) {
        const double2 v = in[i];
        a[i] = v.x;
        b[i] = v.y;
    }
}
void launch_interleave(const double* a, const double* b, double* out, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    const unsigned blocks = unsigned(std::min<int64_t>((n + 255) / 256, 148 * 16));
    interleave_kernel<<<blocks, 256, 0, st>>>(a, b, reinterpret_cast<double2*>(out), n);
}
void launch_deinterleave(const double* in, double* a, double* b, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    const unsigned blocks = unsigned(std::min<int64_t>((n + 255) / 256, 148 * 16));
    deinterleave_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const double2*>(in), a, b, n);
}

// grad_ops[op_index] = s0 * sum over tiles of partials[t, :], fixed order
__global__ void __launch_bounds__(256) batch_grad_reduce_kernel(const double* __restrict__ partials, int n_tiles,
                                                                const BatchOp* __restrict__ ops, double* __restrict__ grad_ops,
                                                                SetStrides ss) {
    __shared__ double wsum[8];
    partials += int64_t(blockIdx.y) * ss.part;
    grad_ops += int64_t(blockIdx.y) * ss.grad;
    const int t = blockIdx.x;
    const double* row = partials + int64_t(t) * n_tiles;
    double s = 0.0;
    for (int i = threadIdx.x; i < n_tiles; i += 256) s += row[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += wsum[w];
        grad_ops[ops[t].op_index] = tot * double(ops[t].s0);
    }
}

size_t batch_smem_bytes(int max_rows, int max_cols, int nops, int nv) {
    size_t ld = size_t(max_cols) | 1;
    size_t bytes = size_t(nv) * max_rows * ld * sizeof(double);
    bytes += size_t(nops) * sizeof(SmemOp);
    if (nv == 2) bytes += size_t(nops) * BK_WARPS * sizeof(double);
    bytes += size_t(max_rows) * sizeof(uint32_t);
    bytes += size_t(max_cols) * 6 + 8;  // scatter tables of the dynamic column layout
    return bytes + 16;
}

// shared memory and resident CTAs per SM of the persistent kernel for tiles up to max_rows x max_cols (0 CTAs: not usable)
void persist_config(int nv, int threads, int max_rows, int max_cols, int nops, size_t* smem, int* ctas_per_sm) {
    *smem = lean_persist_smem_bytes(max_rows, max_cols, nops, nv);
    *ctas_per_sm = 0;
    if (max_cols > 32 * BK_COLCH || *smem > 232448) return;
    int n = 0;
    cudaError_t e = nv == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lean_persist_kernel<1>, threads, *smem)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lean_persist_kernel<3>, threads, *smem);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    *ctas_per_sm = n;
}

cudaError_t batch_kernels_configure() {
    cudaError_t e = cudaFuncSetAttribute(batch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lean_batch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lean_batch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lean_batch_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lean_persist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(lean_persist_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(batch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
}

cudaError_t StreamFan::create(int n_aux) {
    cudaError_t e = cudaEventCreateWithFlags(&start, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    for (int i = 0; i < n_aux; ++i) {
        cudaStream_t s;
        cudaEvent_t ev;
        if ((e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) return e;
        aux.push_back(s);
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        done.push_back(ev);
    }
    return cudaSuccess;
}

void StreamFan::destroy() {
    for (auto s : aux) cudaStreamDestroy(s);
    for (auto e : done) cudaEventDestroy(e);
    if (start) cudaEventDestroy(start);
    aux.clear();
    done.clear();
    start = nullptr;
}

// runs launch(rect, stream) for every rectangle: the first on `st`, the others on the auxiliary streams between a
// fork and a join event (serially on `st` without a fan)
template <class F>
static void fan_out(const std::vector<TileRect>& rects, cudaStream_t st, StreamFan* fan, F launch) {
    const size_t n = rects.size();
    if (!fan || n < 2 || fan->aux.empty()) {
        for (const TileRect& r : rects) launch(r, st);
        return;
    }
    const size_t n_aux = fan->aux.size();
    cudaEventRecord(fan->start, st);
    std::vector<char> used(n_aux, 0);
    for (size_t i = 0; i < n; ++i) {
        if (i == 0) {
            launch(rects[i], st);
            continue;
        }
        const size_t k = (i - 1) % n_aux;
        if (!used[k]) cudaStreamWaitEvent(fan->aux[k], fan->start, 0);
        used[k] = 1;
        launch(rects[i], fan->aux[k]);
    }
    for (size_t k = 0; k < n_aux; ++k)
        if (used[k]) {
            cudaEventRecord(fan->done[k], fan->aux[k]);
            cudaStreamWaitEvent(st, fan->done[k], 0);
        }
}

// v_out: null = in place (static column order); otherwise out of place with the dynamic column layout (BatchDev::B.dyn_*)
void launch_batch_forward(double* v, double* v_out, int64_t nb, const BatchDev& bd, const double2* rot, const std::vector<TileRect>& rects,
                          int threads, int n_sets, SetStrides ss, cudaStream_t st, StreamFan* fan, unsigned* ctl, int sms, int min_waves) {
    fan_out(rects, st, fan, [&](const TileRect& r, cudaStream_t s) {
        dim3 grid(unsigned(r.pa_cnt) * unsigned(r.pb_cnt), unsigned(n_sets), 1);
        if (grid.x == 0) return;  // a rank of a sharded plan may own no class of this batch
        const size_t ri = size_t(&r - rects.data());
        if (bd.lean && ctl && !v_out && n_sets == 1 && r.ctas_p_fwd > 0 && ri < 16 && grid.x >= unsigned(min_waves) * unsigned(r.ctas_p_fwd) * unsigned(sms)) {
            const unsigned ctas = std::min(grid.x, unsigned(r.ctas_p_fwd) * unsigned(sms));
            lean_persist_kernel<1><<<ctas, threads, r.smem_p_fwd, s>>>(v, nb, bd, rot, nullptr, 0, r, ctl + 2 * ri, r.max_rows, r.max_cols);
            return;
        }
        if (bd.lean)
            lean_batch_kernel<1><<<grid, threads, r.smem_fwd, s>>>(v, nullptr, v_out, nb, bd, rot, nullptr, 0, ss, r);
        else
            batch_kernel<1><<<grid, threads, r.smem_fwd, s>>>(v, nullptr, v_out, nb, bd, rot, nullptr, 0, ss, r);
    });
}

void launch_batch_reverse(double* ket, double* bra, int64_t nb, const BatchDev& bd, const double2* rot,
                          const std::vector<TileRect>& rects, int threads, double* partials, double* grad_ops, int n_sets,
                          SetStrides ss, cudaStream_t st, StreamFan* fan) {
    const int n_tiles = bd.B.n_panels * bd.A.n_panels;
    fan_out(rects, st, fan, [&](const TileRect& r, cudaStream_t s) {
        dim3 grid(unsigned(r.pa_cnt) * unsigned(r.pb_cnt), unsigned(n_sets), 1);
        if (grid.x == 0) return;
        if (bd.lean)
            lean_batch_kernel<2><<<grid, threads, r.smem_rev, s>>>(ket, bra, nullptr, nb, bd, rot, partials, n_tiles, ss, r);
        else
            batch_kernel<2><<<grid, threads, r.smem_rev, s>>>(ket, bra, nullptr, nb, bd, rot, partials, n_tiles, ss, r);
    });
    batch_grad_reduce_kernel<<<dim3(bd.nops, unsigned(n_sets), 1), 256, 0, st>>>(partials, n_tiles, bd.ops, grad_ops, ss);
}

// reverse sweep of one batch on interleaved (ket, bra) elements: kb = [rows][nb][2].  kb_out: null = in place in the
// static column order; otherwise the batch reads kb with its columns in the batch's own order and writes kb_out in the
// order of the next batch of the sweep (BatchDev::B.dyn_*, dynamic column layout)
void launch_batch_reverse_il(double* kb, double* kb_out, int64_t nb, const BatchDev& bd, const double2* rot, const std::vector<TileRect>& rects,
                             int threads, double* partials, double* grad_ops, int n_sets, SetStrides ss, cudaStream_t st, StreamFan* fan,
                             unsigned* ctl, int sms, int min_waves) {
    const int n_tiles = bd.B.n_panels * bd.A.n_panels;
    fan_out(rects, st, fan, [&](const TileRect& r, cudaStream_t s) {
        const unsigned tiles = unsigned(r.pa_cnt) * unsigned(r.pb_cnt);
        if (tiles == 0) return;
        const size_t ri = size_t(&r - rects.data());
        // persistent kernel only when every resident CTA gets several tiles: with about as many tiles as CTAs the ticket
        // look-ahead (two tickets per CTA at start) would leave half the SMs idle -- (12e,12o): 2.46 vs 1.37 ms per sweep
        if (ctl && !kb_out && n_sets == 1 && r.ctas_p_rev > 0 && ri < 16 && tiles >= unsigned(min_waves) * unsigned(r.ctas_p_rev) * unsigned(sms)) {
            const unsigned ctas = std::min(tiles, unsigned(r.ctas_p_rev) * unsigned(sms));
            lean_persist_kernel<3><<<ctas, threads, r.smem_p_rev, s>>>(kb, nb, bd, rot, partials, n_tiles, r, ctl + 32 + 2 * ri, r.max_rows, r.max_cols);
            return;
        }
        lean_batch_kernel<3><<<dim3(tiles, unsigned(n_sets), 1), threads, r.smem_il, s>>>(kb, nullptr, kb_out, nb, bd, rot, partials, n_tiles, ss, r);
    });
    batch_grad_reduce_kernel<<<dim3(bd.nops, unsigned(n_sets), 1), 256, 0, st>>>(partials, n_tiles, bd.ops, grad_ops, ss);
}

// out[a, b] = in[map_a[a], map_b[b]]  (storage order <-> reference order)
__global__ void __launch_bounds__(256) permute_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t row0, int64_t nb,
                                                      const uint32_t* __restrict__ map_a, const uint32_t* __restrict__ map_b) {
    // `out` and `map_a` are already offset to row0 (row chunk of the launch); an identity row map reads row0 + a
    const int64_t a = blockIdx.y;
    const int64_t src_row = (map_a ? int64_t(map_a[a]) : row0 + a) * nb;
    for (int64_t b = int64_t(blockIdx.x) * 256 + threadIdx.x; b < nb; b += int64_t(gridDim.x) * 256)
        out[a * nb + b] = in[src_row + map_b[b]];
}

void launch_permute(const double* in, double* out, int64_t na, int64_t nb, const uint32_t* map_a, const uint32_t* map_b,
                    cudaStream_t st) {
    unsigned gx = unsigned(std::min<int64_t>((nb + 255) / 256, 64));
    for (int64_t r0 = 0; r0 < na; r0 += 65535) {
        unsigned gy = unsigned(std::min<int64_t>(na - r0, 65535));
        permute_kernel<<<dim3(gx, gy, 1), 256, 0, st>>>(in, out + r0 * nb, r0, nb, map_a ? map_a + r0 : nullptr, map_b);
    }
}

}  // namespace tcc
