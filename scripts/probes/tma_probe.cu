// Probe (diagnostic, not part of the library): achievable HBM throughput of a tile gather / scatter made of many small
// 1-D TMA bulk copies (cp.async.bulk), in the access pattern of the batched reverse sweep: a tile = ROWS rows x NRUN runs
// of RUN consecutive 16-byte elements; rows of a tile are `panels` rows apart, runs nb / NRUN elements apart.
// One CTA per SM (or more), a single producer warp, NBUF tile buffers: load tile k+1 while tile k is stored back.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/probes/tma_probe scripts/probes/tma_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

struct Geo {
    int rows, nrun, run, panels_a, panels_b;  // tile = rows x (nrun * run) elements
    int64_t nb;                                // elements per matrix row
    int nbuf, modify;
    int csync;     // 1: cluster barrier between the gather and the scatter (launched as clusters of adjacent column panels)
    int pf_dist;   // > 0: the CTA of tile (pa, pb) prefetches into L2 slice pb of the rows of row panel pa + pf_dist (contiguous pieces)
    int order;     // 0: consecutive CTAs take consecutive column panels, 1: consecutive row panels
    int contig;    // 1: a tile is one contiguous block of rows * nrun * run elements (plain streaming copy in tiles)
    int rowgroup;  // rows of a tile come in groups of `rowgroup` consecutive matrix rows (1: every row far from the others)
};

// mode 0: TMA bulk loads + bulk stores by one producer warp.
__global__ void __launch_bounds__(32) tma_copy_kernel(double2* v, Geo g) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[4];
    const int lane = threadIdx.x;
    const int ld = g.nrun * g.run + 1;
    const size_t buf_bytes = (size_t(g.rows) * ld * 16 + 127) & ~size_t(127);
    if (lane == 0)
        for (int i = 0; i < g.nbuf; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    __syncwarp();
    const int n_tiles = g.panels_a * g.panels_b;
    const int per_tile = g.rows * g.nrun;
    const uint32_t tile_bytes = uint32_t(per_tile) * g.run * 16;
    auto issue = [&](int tile, int buf, bool load) {
        const int pa = g.order ? tile % g.panels_a : tile / g.panels_b, pb = g.order ? tile / g.panels_a : tile % g.panels_b;
        unsigned char* base = smem + buf * buf_bytes;
        if (load) {
            if (lane == 0) mbar_expect_tx(&bars[buf], tile_bytes);
            __syncwarp();
        }
        for (int i = lane; i < per_tile; i += 32) {
            const int r = i / g.nrun, q = i - r * g.nrun;
            const int64_t row = int64_t(r / g.rowgroup) * g.panels_a * g.rowgroup + int64_t(pa) * g.rowgroup + r % g.rowgroup;
            const int64_t col = int64_t(q) * ((g.nb - 0) / g.nrun) + int64_t(pb) * g.run;
            double2* gp = g.contig ? v + (int64_t(tile) * g.rows + r) * (g.nrun * g.run) + q * g.run : v + row * g.nb + col;
            unsigned char* sp = base + (size_t(r) * ld + q * g.run) * 16;
            if (load)
                bulk_load(sp, gp, g.run * 16, &bars[buf]);
            else
                bulk_store(gp, sp, g.run * 16);
        }
        if (!load) bulk_commit();
    };
    uint32_t phase[4] = {0, 0, 0, 0};
    int k = 0;
    int tile = blockIdx.x;
    if (tile >= n_tiles) return;
    issue(tile, 0, true);
    for (;; ++k) {
        const int buf = k % g.nbuf;
        const int next = tile + gridDim.x;
        if (next < n_tiles) {
            if (g.nbuf > 1) {
                bulk_wait_read0();  // the buffer the next tile lands in has been read out by its stores
                issue(next, (k + 1) % g.nbuf, true);
            }
        }
        mbar_wait(&bars[buf], phase[buf]);
        phase[buf] ^= 1;
        if (g.modify) {
            double2* t = reinterpret_cast<double2*>(smem + buf * buf_bytes);
            for (int i = lane; i < g.rows * ld; i += 32) t[i].x += 1.0;
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncwarp();
        }
        issue(tile, buf, false);
        if (next >= n_tiles) break;
        if (g.nbuf == 1) {
            bulk_wait_read0();
            issue(next, 0, true);
        }
        tile = next;
    }
    bulk_wait0();
}

// mode 1: the current method -- 16-byte cp.async gather by all threads, LDS + STG scatter (256 threads, one tile per CTA).
// Address arithmetic as lean as in the library kernels: column offsets of a lane in registers, one row base per row.
template <int NJ>
__global__ void __launch_bounds__(256) ldgsts_copy_kernel(double2* v, Geo g) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int nC = g.nrun * g.run;
    const int ld = nC + 1;
    const int tile = blockIdx.x;
    const int pa = g.order ? tile % g.panels_a : tile / g.panels_b, pb = g.order ? tile / g.panels_a : tile % g.panels_b;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* t = reinterpret_cast<double2*>(smem);
    int64_t colofs[NJ];
    int lc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int c = lane + 32 * j;
        lc[j] = c < nC ? c : -1;
        const int q = c / g.run;
        colofs[j] = g.contig ? c : int64_t(q) * (g.nb / g.nrun) + int64_t(pb) * g.run + (c - q * g.run);
    }
    const int64_t row0 = g.contig ? int64_t(tile) * g.rows : (g.rowgroup == 1 ? int64_t(pa) : int64_t(pa) * g.rows);
    const int64_t rstep = g.contig ? 1 : (g.rowgroup == 1 ? g.panels_a : 1);
    const int64_t rowlen = g.contig ? nC : g.nb;
    if (g.pf_dist > 0 && !g.contig && pa + g.pf_dist < g.panels_a && warp == 0) {
        // rows of panel pa + pf_dist as one stream; this CTA prefetches slice pb of panels_b in 4 KB pieces
        const int64_t total = int64_t(g.rows) * g.nb;
        const int64_t lo = total * pb / g.panels_b, hi = total * (pb + 1) / g.panels_b;
        const int64_t prow0 = g.rowgroup == 1 ? int64_t(pa + g.pf_dist) : int64_t(pa + g.pf_dist) * g.rows;
        for (int64_t e = lo + lane * 256; e < hi; e += 32 * 256) {
            int64_t end = e + 256 < hi ? e + 256 : hi;
            const int64_t r = e / g.nb;
            const int64_t row_end = (r + 1) * g.nb;
            if (end > row_end) {
                if (r + 1 < g.rows)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(v + (prow0 + (r + 1) * rstep) * g.nb), "r"(uint32_t(end - row_end) * 16) : "memory");
                end = row_end;
            }
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(v + (prow0 + r * rstep) * g.nb + (e - r * g.nb)), "r"(uint32_t(end - e) * 16) : "memory");
        }
    }
    if (g.modify != 3)
        for (int r = warp; r < g.rows; r += 8) {
            const double2* rp = v + (row0 + r * rstep) * rowlen;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (lc[j] >= 0) {
                    const unsigned s = smem_u32(t + r * ld + lc[j]);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(rp + colofs[j]) : "memory");
                }
        }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
    if (g.csync) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
    if (g.modify == 2) {
        if (t[threadIdx.x].x == 123.456) v[0].x = 1.0;  // keep the loads alive
        return;
    }
    for (int r = warp; r < g.rows; r += 8) {
        double2* rp = v + (row0 + r * rstep) * rowlen;
        double2 x[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if (lc[j] >= 0) {
                x[j] = g.modify == 3 ? make_double2(1.0, 2.0) : t[r * ld + lc[j]];
                if (g.modify == 1) x[j].x += 1.0;
            }
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if (lc[j] >= 0) rp[colofs[j]] = x[j];
    }
}

template <int NJ>
static void launch_ldgsts(double2* v, Geo g, size_t buf_bytes, int cluster) {
    CK(cudaFuncSetAttribute(ldgsts_copy_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(buf_bytes)));
    if (cluster > 1) {
        CK(cudaFuncSetAttribute(ldgsts_copy_kernel<NJ>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(unsigned(g.panels_a * g.panels_b) / cluster * cluster);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = buf_bytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, ldgsts_copy_kernel<NJ>, v, g));
    } else
        ldgsts_copy_kernel<NJ><<<g.panels_a * g.panels_b, 256, buf_bytes>>>(v, g);
}

int main(int argc, char** argv) {
    Geo g;
    g.rows = argc > 9 ? atoi(argv[9]) : 70;
    g.order = argc > 10 ? atoi(argv[10]) : 0;
    const int pad = argc > 11 ? atoi(argv[11]) : 0;
    g.pf_dist = argc > 12 ? atoi(argv[12]) : 0;
    g.nrun = argc > 1 ? atoi(argv[1]) : 5;
    g.run = argc > 2 ? atoi(argv[2]) : 14;
    const int ctas_per_sm = argc > 3 ? atoi(argv[3]) : 1;
    g.nbuf = argc > 4 ? atoi(argv[4]) : 2;
    const int mode = argc > 5 ? atoi(argv[5]) : 0;
    g.modify = argc > 6 ? atoi(argv[6]) : 0;
    g.rowgroup = argc > 7 ? atoi(argv[7]) : 1;
    g.contig = argc > 8 ? atoi(argv[8]) : 0;
    g.panels_a = 12880 / g.rows;
    g.panels_b = 12880 / (g.nrun * g.run);
    const int64_t total_elems = argc > 16 ? atoll(argv[16]) : 165894400ll;
    const int reps = argc > 17 ? atoi(argv[17]) : 1;
    const int64_t nb_override = argc > 13 ? atoll(argv[13]) : 0;
    const int cluster = argc > 14 ? atoi(argv[14]) : 1;
    g.csync = argc > 15 ? atoi(argv[15]) : 0;
    if (nb_override > 0) {  // matrix of about the same size with rows of nb_override elements
        g.panels_b = int(nb_override / (g.nrun * g.run));
        g.panels_a = int(total_elems / nb_override / g.rows);
    }
    const int64_t na = int64_t(g.rows) * g.panels_a;
    g.nb = nb_override > 0 ? nb_override : int64_t(g.nrun) * g.run * g.panels_b + pad;
    const int64_t N = na * g.nb;
    double2* v;
    CK(cudaMalloc(&v, N * sizeof(double2)));
    CK(cudaMemset(v, 0, N * sizeof(double2)));
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int ld = g.nrun * g.run + 1;
    const size_t buf_bytes = (size_t(g.rows) * ld * 16 + 127) & ~size_t(127);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
        CK(cudaEventRecord(e0));
        for (int rep = 0; rep < reps; ++rep)
        if (mode == 0) {
            const size_t sm = buf_bytes * g.nbuf;
            CK(cudaFuncSetAttribute(tma_copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
            tma_copy_kernel<<<sms * ctas_per_sm, 32, sm>>>(v, g);
        } else {
            const int nC = g.nrun * g.run;
            if (nC <= 96) launch_ldgsts<3>(v, g, buf_bytes, cluster);
            else if (nC <= 160) launch_ldgsts<5>(v, g, buf_bytes, cluster);
            else if (nC <= 288) launch_ldgsts<9>(v, g, buf_bytes, cluster);
            else launch_ldgsts<18>(v, g, buf_bytes, cluster);
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        if (it > 0 && ms < best) best = ms;
    }
    if (g.modify == 1) {  // every element was incremented once per launch
        std::vector<double2> h(1000);
        CK(cudaMemcpy(h.data(), v + N / 2, sizeof(double2) * h.size(), cudaMemcpyDeviceToHost));
        for (auto& x : h)
            if (x.x != 6.0) { printf("MISMATCH %f\n", x.x); break; }
    }
    printf("cluster %d csync %d pf_dist %d rows %d order %d nb %lld contig %d rowgroup %d mode %d nrun %d run %d (%d B) ctas/sm %d nbuf %d modify %d: %.3f ms, %.0f GB/s (read + write), N = %lld\n", cluster, g.csync, g.pf_dist, g.rows, g.order, (long long)g.nb, g.contig, g.rowgroup, mode, g.nrun, g.run,
           g.run * 16, ctas_per_sm, g.nbuf, g.modify, best, (g.modify >= 2 ? 1.0 : 2.0) * 16.0 * double(N) / best * 1e-6, (long long)N);
    return 0;
}
