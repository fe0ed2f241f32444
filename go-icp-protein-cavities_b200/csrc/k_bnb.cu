// Branch-and-bound bound evaluation kernels (north_star (b)).
//
//  inner_bnb_kernel   one CTA runs one whole GoICP::InnerBnB call (jly_goicp.cpp:286-579): the rotated cloud, weights
//                     and rotation-uncertainty radii are staged in shared memory once, then each pop of the
//                     translation queue evaluates its 8 child cubes x Nd points (warp = child cube, lane = point:
//                     translate, voxel index in FP64, DT gather, x weight, - radius, clamp) plus the 27 lattice corners
//                     of the fork's incompatibility / c-FPFH terms.  Persistent CTAs pull calls from an atomic counter,
//                     so one launch evaluates every (rotation cube, level) request of every pair of a wave.
//                     EXACT=true reproduces the reference's sequential float sums bit for bit (16 independent
//                     FADD chains on 16 lanes), EXACT=false uses warp-shuffle tree sums.
//  eval_bounds_kernel flat wave: one warp per (rotation cube, translation sub-cube), leaf-level (ub, lb) only.
//
// The translation priority queue lives in global memory (one slab per CTA) and follows libstdc++'s
// push_heap/pop_heap step for step so that ties between equal (lb, w) keys pop in the reference's order.
#include "icp_device.cuh"
#include "launch.h"

namespace {

constexpr int BNB_MAX_THREADS = 512;

// ---- 1-D TMA (cp.async.bulk) global -> shared with mbarrier completion: the S<=~26 DT volume of a call is staged in shared
//      memory by the copy engine while the CTA rotates the cloud ----------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra.uni WAIT_DONE;\n bra.uni WAIT_LOOP;\n WAIT_DONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// TRANSNODE operator< (jly_goicp.h:79-86)
__device__ __forceinline__ bool node_less(const HeapEnt& a, const HeapEnt& b) {
    if (a.lb != b.lb) return a.lb > b.lb;
    return a.w < b.w;
}
// The translation queue: entries [0, HEAP_SMEM) live in shared memory (the levels every pop walks), the rest in the
// CTA's global slab.  Only thread 0 touches it.
constexpr int HEAP_SMEM = 128;
// corner-memo hash: the coordinates are dyadic floats (long runs of trailing zero bits), so mix with rotations and take the
// HIGH bits of a multiplicative hash
__device__ __forceinline__ unsigned memo_hash(unsigned kx, unsigned ky, unsigned kz) {
    unsigned h = kx * 0x9E3779B1u;
    h = __funnelshift_l(h, h, 13) ^ (ky * 0x85EBCA77u);
    h = __funnelshift_l(h, h, 11) ^ (kz * 0xC2B2AE3Du);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15;
    return h * 0x846CA68Bu;
}
__device__ __forceinline__ unsigned memo_slot(unsigned h, int shift) { return h >> shift; }
struct Heap {
    float4* s;       // shared: 2 x float4 per entry
    HeapEnt* g;      // global slab
    __device__ __forceinline__ void store(int i, const HeapEnt& e) const {
        float4* p = (i < HEAP_SMEM) ? s + 2 * i : reinterpret_cast<float4*>(g + i);
        p[0] = make_float4(e.lb, e.w, e.x, e.y);
        p[1] = make_float4(e.z, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ HeapEnt load(int i) const {
        const float4* p = (i < HEAP_SMEM) ? s + 2 * i : reinterpret_cast<const float4*>(g + i);
        const float4 a = p[0], b = p[1];
        HeapEnt e; e.lb = a.x; e.w = a.y; e.x = a.z; e.y = a.w; e.z = b.x; e.pad0 = e.pad1 = e.pad2 = 0.f;
        return e;
    }
};
// std::push_heap (__push_heap) on h[0..n) + val
__device__ __forceinline__ void heap_push(const Heap& h, int& n, const HeapEnt& val) {
    int hole = n++;
    int parent = (hole - 1) / 2;
    while (hole > 0) {
        const HeapEnt pe = h.load(parent);
        if (!node_less(pe, val)) break;
        h.store(hole, pe);
        hole = parent; parent = (hole - 1) / 2;
    }
    h.store(hole, val);
}
// std::pop_heap (__adjust_heap to the bottom, then __push_heap of the former last element)
__device__ __forceinline__ HeapEnt heap_pop(const Heap& h, int& n) {
    const HeapEnt top = h.load(0);
    const int len = --n;
    if (len > 0) {
        const HeapEnt val = h.load(len);
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            HeapEnt c1 = h.load(child); const HeapEnt c0 = h.load(child - 1);
            if (node_less(c1, c0)) { child--; c1 = c0; }
            h.store(hole, c1); hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            h.store(hole, h.load(child - 1)); hole = child - 1;
        }
        int parent = (hole - 1) / 2;
        while (hole > 0) {
            const HeapEnt pe = h.load(parent);
            if (!node_less(pe, val)) break;
            h.store(hole, pe);
            hole = parent; parent = (hole - 1) / 2;
        }
        h.store(hole, val);
    }
    return top;
}

struct BnbShared {
    float ub[8], lb[8];
    int cnt[27];
    float cf[27];
    float X[3], Y[3], Z[3];
    float wc, mtd;
    float optErrorT;
    int running, prob, heapN, status;
    int pops, subcubes, improved;
    float best[4];
    int missList[27];
    int nmiss, workCtr;
    unsigned gen;
    long long t0; int missTot;
#ifdef GOICP_PHASE_TIMING
    long long tp[6]; long long tmark;
#endif
};

// Per pop of the translation queue:
//   phase A (all warps)  flat list of (child cube | lattice corner) x 32-point chunks, dealt round-robin to the warps;
//   phase B (warp 0)     the 16 sequential (ub, lb) sums [EXACT] or the fixed-order combine of per-chunk partial sums;
//                        warp 1 does the 27 c-FPFH corner sums meanwhile; trimmed sums use warps 0..7;
//   phase C (warp 0)     per-child corner min/max on 8 lanes, then lane 0: decisions, pushes and the next pop.
// PERSIST=false: the calls are probs[0..nprob), claimed through counter[0] (one launch per wave).
// PERSIST=true : the kernel stays resident for a whole batch and serves the host's request ring `q` (see QueueDev).
// GS=true (needs SMEM): the call's DT volume (float distances + one colour-mask byte per voxel) is staged in shared memory by
// TMA, so the per-point gathers are LDS instead of L1/L2 sector gathers (S^3 * 5 bytes at dynamic-smem offset gridOff).
template <bool EXACT, bool PERSIST, bool SMEM, bool GS>
__global__ void __launch_bounds__(BNB_MAX_THREADS, 2)
inner_bnb_kernel(const PairDev* __restrict__ pairs, const InnerProb* probs, InnerOut* outs,
                 int nprob, int* __restrict__ counter, HeapEnt* __restrict__ heaps, int heapCap,
                 float* gscratch, size_t gstride, int NdP, int NdQ, int useSmem, QueueDev q,
                 uint4* memoAll, int memoCap, unsigned* genCounter, int gridOff, int S3p) {
    unsigned long long* dstat = reinterpret_cast<unsigned long long*>(genCounter) + 1;   // [0] busy cycles [1] pops [2] corner misses [3] calls [4] poll cycles   // gscratch is exchanged between threads: no __restrict__
    extern __shared__ float4 dyn_smem4[];
    __shared__ BnbShared sh;
    __shared__ float4 sheap[2 * HEAP_SMEM];
    __shared__ InnerProb s_pr;
    __shared__ InnerOut s_out;
    __shared__ unsigned long long s_gbar;   // mbarrier of the DT staging copies
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarps = blockDim.x >> 5;
    float* base = SMEM ? reinterpret_cast<float*>(dyn_smem4) : gscratch + (size_t)blockIdx.x * gstride;   // SMEM: address space known -> LDS/STS
    float* tx = base; float* ty = tx + NdP; float* tz = ty + NdP; float* wgt = tz + NdP; float* mrd = wgt + NdP;
    uint8_t* dprop_s = reinterpret_cast<uint8_t*>(mrd + NdP);   // [NdP] colour index of each data point
    float* part = mrd + NdP + (NdP >> 2);   // [8][nchunks][2] + [27][nchunks] per-chunk partial sums (tree-sum mode)
    float* md = part + 43 * (NdP >> 5);   // EXACT: [16][NdQ] sum terms (ub, lb per child); trimmed: [8][NdQ] residuals
    float* fp = md + (EXACT ? 16 : 8) * NdQ;   // [27][NdQ]  (EXACT with the c-FPFH term)
    Heap heap; heap.s = sheap; heap.g = heaps + (size_t)blockIdx.x * heapCap;
    uint4* memo = memoAll + 2 * (size_t)blockIdx.x * memoCap;
    const int memoShift = 32 - (31 - __clz(memoCap));   // this CTA's corner memo: direct-mapped, 32 B entries, tagged with the call's generation
    float* sdist = reinterpret_cast<float*>(dyn_smem4) + gridOff;                     // GS: [S3p] DT distances
    uint8_t* svm = reinterpret_cast<uint8_t*>(sdist + S3p);                          // GS: [S3p] colour mask of the voxel's closest cell
    float* icpTile;   // model tile of an ICP request: aliases the staging arrays (the dynamic region holds >= 3*NN_TILE floats)
    if constexpr (SMEM) icpTile = reinterpret_cast<float*>(dyn_smem4); else { __shared__ float s_tile[3 * NN_TILE]; icpTile = s_tile; }
    unsigned gphase = 0; int gpair = -1;                                              // mbarrier parity; pair whose volume is staged
    if (GS) { if (tid == 0) mbar_init(&s_gbar, 1); __syncthreads(); }

    for (;;) {
        __syncthreads();
        if (!PERSIST) {
            if (tid == 0) sh.prob = atomicAdd(counter, 1);
            __syncthreads();
            if (sh.prob >= nprob) {   // the last CTA to leave re-arms the counters, so the host never has to memset them
                if (tid == 0) { __threadfence(); if (atomicAdd(counter + 1, 1) == (int)gridDim.x - 1) { counter[0] = 0; counter[1] = 0; } }
                return;
            }
            if (tid < (int)(sizeof(InnerProb) / 4)) reinterpret_cast<int*>(&s_pr)[tid] = reinterpret_cast<const volatile int*>(probs + sh.prob)[tid];
        } else if (warp == 0) {
            // claim the next ring index; the 16 lanes read the 64-byte cell with one load until both lap tags are there
            unsigned i = 0;
            const long long tp0 = clock64();
            if (lane == 0) i = atomicAdd(q.claim, 1u);
            i = __shfl_sync(GOICP_FULL, i, 0);
            const unsigned want = (i >> q.cellShift) + 1u;
            const volatile unsigned* cell = reinterpret_cast<const volatile unsigned*>(q.cells + (i & q.cellMask));
            unsigned w = 0, backoff = 64; unsigned long long t0 = 0, now; int notReady = 0; bool dead = false;
            for (;;) {
                if (lane < 16) w = cell[lane];
                const unsigned a = __shfl_sync(GOICP_FULL, w, 0), b = __shfl_sync(GOICP_FULL, w, 15);
                if (a == want && b == want) break;
                notReady = 1;
                __nanosleep(backoff); if (backoff < 16384) backoff <<= 1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 60000000000ull) { dead = true; break; }   // safety net: 60 s without work -> leave
            }
            const unsigned slot = __shfl_sync(GOICP_FULL, w, 1);
            if (lane >= 2 && lane < 14) reinterpret_cast<unsigned*>(&s_pr)[lane - 2] = w;
            if (lane == 0) {
                sh.prob = (dead || slot == 0xFFFFFFFFu) ? -1 : (int)slot;
                atomicAdd(dstat + 4, (unsigned long long)(clock64() - tp0)); atomicAdd(dstat + 5, (unsigned long long)notReady);
            }
        }
        __syncthreads();
        const int p = sh.prob;
        if (PERSIST && p < 0) return;
        const InnerProb pr = s_pr;
        if (PERSIST && pr.level == GOICP_REQ_ICP) {   // an ICP / scoring request (GoICP::ICP): state pointer packed into R[0..1]
            IcpState* gst = reinterpret_cast<IcpState*>(((unsigned long long)__float_as_uint(pr.R[1]) << 32) | (unsigned long long)__float_as_uint(pr.R[0]));
            icp_fused_body(pairs, gst, icpTile);
            __syncthreads();
            if (warp == 0) {
                if (lane < 16) reinterpret_cast<unsigned*>(q.outs + p)[lane] = (lane == 0 || lane == 15) ? 1u : 0u;
            }
            continue;
        }
        const PairDev& P = pairs[pr.pair];
        const GridDev& g = P.g;
        const bool gload = GS && gpair != pr.pair;   // (the __syncthreads at the top of the loop ordered the last reads of the old volume)
        if (gload && tid == 0) {
            mbar_expect_tx(&s_gbar, (unsigned)S3p * 5u);
            tma_bulk_g2s(sdist, g.dist, (unsigned)S3p * 4u, &s_gbar);
            tma_bulk_g2s(svm, g.vmask8, (unsigned)S3p, &s_gbar);
        }
        const int Nd = P.Nd;
        const int nchunks = (Nd + 31) >> 5;
        const float* __restrict__ dist = g.dist;
        const bool corners = P.use_reg || P.use_fpfh;
        const bool doTrim = P.doTrim != 0;
        const bool useMd = EXACT || doTrim;
        const int ncp1 = g.ncells + 1;
        const int norm = P.norm;
        const int G = min(nchunks, 8), ngroups = (nchunks + G - 1) / G;   // chunks per work item, items per row
        // per-problem constants in registers (the PairDev lives in global memory)
        const int S = g.S;
        const double gx0 = g.xMin, gy0 = g.yMin, gz0 = g.zMin, gscale = g.scale;
        const int* __restrict__ vcell = g.vcell;
        const uint32_t* __restrict__ cmask = g.cmask;
        const uint32_t* __restrict__ vmask = g.vmask;
        const float* __restrict__ fpfhD = P.fpfhD;
        const bool use_reg = P.use_reg != 0, use_fpfh = P.use_fpfh != 0;
        const VoxFast vf = vox_fast_of(g);

        // ---- stage the rotated cloud (jly_goicp.cpp:750-756), weights and rotation radii ---------------------
        for (int i = tid; i < Nd; i += nthreads) {
            const float x = P.dx[i], y = P.dy[i], z = P.dz[i];
            tx[i] = pr.R[0] * x + pr.R[1] * y + pr.R[2] * z;
            ty[i] = pr.R[3] * x + pr.R[4] * y + pr.R[5] * z;
            tz[i] = pr.R[6] * x + pr.R[7] * y + pr.R[8] * z;
            wgt[i] = P.weights[i];
            mrd[i] = pr.level >= 0 ? P.maxRotDis[(size_t)pr.level * Nd + i] : 0.f;   // d - 0 == d
            dprop_s[i] = P.dprop[i];
        }
        if (tid < 27) sh.cnt[tid] = 0;
        if (tid == 0) {
            sh.heapN = 0; sh.status = 0; sh.pops = 1; sh.subcubes = 0; sh.improved = 0;
            sh.gen = atomicAdd(genCounter, 1u) + 1u; sh.nmiss = 0; sh.workCtr = 0; sh.t0 = clock64(); sh.missTot = 0;
#ifdef GOICP_PHASE_TIMING
            for (int k = 0; k < 6; k++) sh.tp[k] = 0; sh.tmark = clock64();
#endif
            sh.optErrorT = pr.optError;                                              // :297
            sh.best[0] = sh.best[1] = sh.best[2] = sh.best[3] = 0.f;
            // the first pop is always the initial node (:300,:314) with lb = 0
            if (pr.optError - 0.f < P.SSEThresh) sh.running = 0;                     // :317
            else {
                sh.running = 1;
                const float wc = P.tWidth / 2;
                sh.wc = wc; sh.mtd = (float)(GOICP_SQRT3 / 2.0 * wc);
                sh.X[0] = P.tMinX; sh.X[1] = P.tMinX + wc; sh.X[2] = sh.X[1] + wc;
                sh.Y[0] = P.tMinY; sh.Y[1] = P.tMinY + wc; sh.Y[2] = sh.Y[1] + wc;
                sh.Z[0] = P.tMinZ; sh.Z[1] = P.tMinZ + wc; sh.Z[2] = sh.Z[1] + wc;
            }
        }

        if (gload) { mbar_wait(&s_gbar, gphase); gphase ^= 1u; gpair = pr.pair; }
        for (;;) {
            __syncthreads();                                                         // (1) the popped node is visible
#ifdef GOICP_PHASE_TIMING
            if (tid == 0) { const long long n_ = clock64(); sh.tp[sh.pops == 1 && sh.subcubes == 0 ? 0 : 3] += n_ - sh.tmark; sh.tmark = n_; }
#endif
            if (!sh.running) break;
            const float wc = sh.wc, mtd = sh.mtd;
            const float half = wc / 2;

            // ---- phase A1: the cube.point bound evals (:343-382) on all warps; the last warp first looks the 27 lattice
            //      corners up in the call's corner memo (the reference memoises corner terms per InnerBnB call too, :304-305) ---
            if (corners && warp == nwarps - 1) {
                bool miss = false;
                if (lane < 27) {
                    const int cz_ = lane / 9, cy_ = (lane - 9 * cz_) / 3, cx_ = lane - 9 * cz_ - 3 * cy_;
                    const unsigned kx = __float_as_uint(sh.X[cx_]), ky = __float_as_uint(sh.Y[cy_]), kz = __float_as_uint(sh.Z[cz_]);
                    const unsigned hsh = memo_hash(kx, ky, kz);
                    const uint4* e = memo + 2 * (size_t)memo_slot(hsh, memoShift);
                    const uint4 e0 = e[0], e1 = e[1];
                    if (e0.x == kx && e0.y == ky && e0.z == kz && e0.w == sh.gen) { sh.cnt[lane] = (int)e1.x; sh.cf[lane] = __uint_as_float(e1.y); }
                    else { miss = true; sh.cnt[lane] = 0; }
                }
                const unsigned mm = __ballot_sync(GOICP_FULL, miss);
                if (miss) sh.missList[__popc(mm & ((1u << lane) - 1u))] = lane;
                if (lane == 0) { sh.nmiss = __popc(mm); sh.workCtr = 0; sh.missTot += __popc(mm); }
            }
            // item = (child cube, group of G 32-point chunks): the per-item overhead is paid once per G points of a lane
            for (int it = warp; it < 8 * ngroups; it += nwarps) {
                const int c = it / ngroups, gi = it - c * ngroups;
                const float transX = sh.X[c & 1] + half, transY = sh.Y[(c >> 1) & 1] + half, transZ = sh.Z[(c >> 2) & 1] + half;   // :331-333
                const int iEnd = min(Nd, (gi + 1) * G * 32);
                const float Cx = vox_fast_c(vf, transX, gx0, gscale), Cy = vox_fast_c(vf, transY, gy0, gscale), Cz = vox_fast_c(vf, transZ, gz0, gscale);
                float su = 0.f, sl = 0.f;
                for (int i = gi * G * 32 + lane; i < iEnd; i += 32) {
                    const float px = tx[i], py = ty[i], pz = tz[i];
                    const int vox = vox_fast(vf, S, px, py, pz, Cx, Cy, Cz);
                    float dv;
                    if (vox >= 0) dv = GS ? sdist[vox] : __ldg(dist + vox);
                    else dv = dt_distance_v<!GS>(S, gx0, gy0, gz0, gscale, GS ? sdist : dist, px + transX, py + transY, pz + transZ);
                    float d = wgt[i] * dv;
                    d = d - mrd[i];
                    if (d < 0.f) d = 0.f;
                    if (EXACT && !doTrim) {   // the two sum terms of this point (:393-415), summed in index order by the chain lanes
                        const float dis = fmaxf(d - mtd, 0.f);
                        md[(2 * c) * NdQ + i] = (norm == 2) ? d * d : d;
                        md[(2 * c + 1) * NdQ + i] = (norm == 2) ? dis * dis : dis;
                    } else if (useMd) md[c * NdQ + i] = d;
                    else {
                        su = su + ((norm == 2) ? d * d : d);
                        const float dis = d - mtd;
                        if (dis > 0.f) sl = sl + ((norm == 2) ? dis * dis : dis);
                    }
                }
                if (!useMd) {
                    su = warp_sum(su); sl = warp_sum(sl);
                    if (lane == 0) { part[2 * it] = su; part[2 * it + 1] = sl; }
                }
            }
            __syncthreads();                                                         // (2)
#ifdef GOICP_PHASE_TIMING
            if (tid == 0) { const long long n_ = clock64(); sh.tp[1] += n_ - sh.tmark; sh.tmark = n_; }
#endif
            // ---- phase A2: warp 0 sums the residuals while the other warps evaluate the corners the memo missed -----------
            if (P.doTrim) {   // radix select replaces intro_select (:384-390); one warp per child
                for (int c = warp; c < 8; c += nwarps) {
                    float su, sl;
                    warp_trimmed_sums(md + c * NdQ, Nd, P.inlierNum, lane, norm, mtd, &su, &sl);
                    if (lane == 0) { sh.ub[c] = su; sh.lb[c] = sl; }
                }
            } else if (warp == 0 && lane < 16) {
                const int c = lane >> 1;
                float acc = 0.f;
                if (EXACT) {   // sequential float sums in index order (:393-415): lane = (child, ub|lb); 16 independent chains
                    const float* m = md + lane * NdQ;
                    const int n = P.inlierNum;
#pragma unroll 8
                    for (int i = 0; i < n; ++i) acc = acc + m[i];
                } else {
                    const float* q = part + 2 * c * ngroups + (lane & 1);
                    for (int k = 0; k < ngroups; ++k) acc = acc + q[2 * k];
                }
                if (lane & 1) sh.lb[c] = acc; else sh.ub[c] = acc;
#ifdef GOICP_PHASE_TIMING
                if (tid == 0) sh.tp[4] += clock64() - sh.tmark;
#endif
            }
            if (corners) {   // corner terms (:431-550, checkCompatibilities :919, sumFPFH :1689): (missed corner, 32-point chunk) items
                const int nItems = sh.nmiss * ngroups;
                for (;;) {
                    int it = 0;
                    if (lane == 0) it = atomicAdd(&sh.workCtr, 1);
                    it = __shfl_sync(GOICP_FULL, it, 0);
                    if (it >= nItems) break;
                    const int m = it / ngroups, gi = it - m * ngroups;
                    const int c = sh.missList[m];
                    const int cz_ = c / 9, cy_ = (c - 9 * cz_) / 3, cx_ = c - 9 * cz_ - 3 * cy_;
                    const float cx = sh.X[cx_], cy = sh.Y[cy_], cz = sh.Z[cz_];
                    const int iEnd = min(Nd, (gi + 1) * G * 32);
                    const float Cx = vox_fast_c(vf, cx, gx0, gscale), Cy = vox_fast_c(vf, cy, gy0, gscale), Cz = vox_fast_c(vf, cz, gz0, gscale);
                    int bad = 0; float fs = 0.f;
                    for (int i = gi * G * 32 + lane; i < iEnd; i += 32) {
                        const float px = tx[i], py = ty[i], pz = tz[i];
                        int vox = vox_fast(vf, S, px, py, pz, Cx, Cy, Cz);
                        if (vox < 0) vox = clamp_vox_v(S, gx0, gy0, gz0, gscale, px + cx, py + cy, pz + cz);
                        if (!use_fpfh) bad += ((((GS ? (unsigned)svm[vox] : __ldg(vmask + vox))) >> dprop_s[i]) & 1u) ? 0 : 1;   // per-voxel mask of the closest cell: one gather
                        else {
                            const int cell = __ldg(vcell + vox);
                            if (use_reg) bad += ((__ldg(cmask + cell) >> dprop_s[i]) & 1u) ? 0 : 1;
                            const float fv = __ldg(fpfhD + (size_t)i * ncp1 + cell);
                            if (EXACT) fp[m * NdQ + i] = fv; else fs = fs + fv;
                        }
                    }
                    if (use_reg) { bad = warp_sum_i(bad); if (lane == 0 && bad) atomicAdd(&sh.cnt[c], bad); }
                    if (!EXACT && use_fpfh) { fs = warp_sum(fs); if (lane == 0) part[16 * ngroups + it] = fs; }
                }
            }
            __syncthreads();                                                         // (3)
#ifdef GOICP_PHASE_TIMING
            if (tid == 0) { const long long n_ = clock64(); sh.tp[2] += n_ - sh.tmark; sh.tmark = n_; }
#endif
            if (warp == 1 && use_fpfh) {   // c-FPFH sums of the missed corners, one chain per lane
                if (lane < sh.nmiss) {
                    float s_ = 0.f;
                    if (EXACT) { const float* f = fp + lane * NdQ; for (int i = 0; i < Nd; ++i) s_ = s_ + f[i]; }   // sumFPFH :1692-1695
                    else { const float* f = part + 16 * ngroups + lane * ngroups; for (int k = 0; k < ngroups; ++k) s_ = s_ + f[k]; }
                    sh.cf[sh.missList[lane]] = (float)(int)(s_ / (float)Nd);          // :1696, int truncation :468,:495 (H7)
                }
                __syncwarp();
                asm volatile("bar.sync 1, 64;" ::: "memory");
            }
            if (warp != 0) continue;
            __syncwarp();
            if (use_fpfh) asm volatile("bar.sync 1, 64;" ::: "memory");
            __syncwarp();
            if (corners && lane < sh.nmiss) {   // remember the freshly computed corners
                const int c = sh.missList[lane];
                const int cz_ = c / 9, cy_ = (c - 9 * cz_) / 3, cx_ = c - 9 * cz_ - 3 * cy_;
                const unsigned kx = __float_as_uint(sh.X[cx_]), ky = __float_as_uint(sh.Y[cy_]), kz = __float_as_uint(sh.Z[cz_]);
                const unsigned hsh = memo_hash(kx, ky, kz);
                uint4* e = memo + 2 * (size_t)memo_slot(hsh, memoShift);
                e[0] = make_uint4(kx, ky, kz, sh.gen);
                e[1] = make_uint4((unsigned)sh.cnt[c], __float_as_uint(sh.cf[c]), 0u, 0u);
            }
            // ---- phase C: corner min/max per child on 8 lanes (:431-550), then lane 0 alone -------------------------
            if (corners && lane < 8) {
                const int j = lane, jx = j & 1, jy = (j >> 1) & 1, jz = (j >> 2) & 1;
                float ub = sh.ub[j], lb = sh.lb[j];
                int minI = 0, maxI = 0; float minF = 0.f, maxF = 0.f;
                for (int k = 0; k < 8; ++k) {
                    const int c = (jx + (k & 1)) + 3 * (jy + ((k >> 1) & 1)) + 9 * (jz + ((k >> 2) & 1));
                    if (P.use_fpfh) { const float f = sh.cf[c]; if (k == 0) { minF = maxF = f; } else { if (f > maxF) maxF = f; if (f < minF) minF = f; } }
                    if (P.use_reg) { const int n = sh.cnt[c]; if (k == 0) { minI = maxI = n; } else { if (n > maxI) maxI = n; if (n < minI) minI = n; } }
                }
                if (P.use_reg) { ub = ub + P.reg * (float)(maxI * maxI); lb = lb + P.reg * (float)(minI * minI); }      // :536-538
                if (P.use_fpfh) { ub = ub + P.regF * (maxF * maxF); lb = lb + P.regF * (minF * minF); }                  // :546-549
                sh.ub[j] = ub; sh.lb[j] = lb;
            }
            __syncwarp();
            if (lane == 0) {   // decisions and pushes in child order (:417-575), then the next pop (:314-320)
                float optErrorT = sh.optErrorT;
                int heapN = sh.heapN;
                for (int j = 0; j < 8; ++j) {
                    const float ub = sh.ub[j], lb = sh.lb[j];
                    const float nx = sh.X[j & 1], ny = sh.Y[(j >> 1) & 1], nz = sh.Z[(j >> 2) & 1];
                    if (ub < optErrorT) {                                                                                         // :554-566
                        optErrorT = ub; sh.improved = 1;
                        sh.best[0] = nx; sh.best[1] = ny; sh.best[2] = nz; sh.best[3] = wc;
                    }
                    if (lb >= optErrorT) continue;                                                                                // :568-572
                    if (heapN >= heapCap) { sh.status = 4; break; }
                    HeapEnt e; e.lb = lb; e.w = wc; e.x = nx; e.y = ny; e.z = nz; e.pad0 = e.pad1 = e.pad2 = 0.f;
                    heap_push(heap, heapN, e);
                }
                sh.subcubes += 8;
                sh.optErrorT = optErrorT;
                if (heapN == 0 || sh.status != 0) sh.running = 0;
                else {
                    const HeapEnt par = heap_pop(heap, heapN);
                    sh.pops++;
                    if (optErrorT - par.lb < P.SSEThresh) sh.running = 0;            // :317
                    else {
                        const float w2 = par.w / 2;                                  // :322
                        sh.wc = w2;
                        sh.mtd = (float)(GOICP_SQRT3 / 2.0 * w2);                    // :323
                        sh.X[0] = par.x; sh.X[1] = par.x + w2; sh.X[2] = sh.X[1] + w2;   // child / corner lattice
                        sh.Y[0] = par.y; sh.Y[1] = par.y + w2; sh.Y[2] = sh.Y[1] + w2;
                        sh.Z[0] = par.z; sh.Z[1] = par.z + w2; sh.Z[2] = sh.Z[1] + w2;
                    }
                }
                sh.heapN = heapN;
            }
        }
        if (tid == 0) {
            atomicAdd(dstat + 0, (unsigned long long)(clock64() - sh.t0)); atomicAdd(dstat + 1, (unsigned long long)sh.pops);
            atomicAdd(dstat + 2, (unsigned long long)sh.missTot); atomicAdd(dstat + 3, 1ull);
#ifdef GOICP_PHASE_TIMING
            for (int k = 0; k < 5; k++) atomicAdd(dstat + 8 + k, (unsigned long long)sh.tp[k]);
#endif
            InnerOut o;
            o.err = sh.optErrorT; o.node[0] = sh.best[0]; o.node[1] = sh.best[1]; o.node[2] = sh.best[2]; o.node[3] = sh.best[3];
            o.improved = sh.improved; o.pops = sh.pops; o.subcubes = sh.subcubes; o.status = sh.status;
            o.seq0 = o.seq1 = 1u; for (int k = 0; k < 5; k++) o.pad[k] = 0;
            s_out = o;
        }
        if (warp == 0) {   // the record leaves the SM as ONE coalesced 64-byte store (it may live in mapped host memory)
            __syncwarp();
            InnerOut* dst = PERSIST ? q.outs + p : outs + p;
            if (lane < 16) reinterpret_cast<unsigned*>(dst)[lane] = reinterpret_cast<const unsigned*>(&s_out)[lane];
        }
    }
}

// Flat wave: every (rotation cube, translation sub-cube) of a frontier in one launch, one warp per sub-cube.
// Leaf-level bounds (pure functions) incl. the corner terms; tree sums.
__global__ void __launch_bounds__(256)
eval_bounds_kernel(const PairDev* __restrict__ pairs, int pair, const float* __restrict__ Rs, const int* __restrict__ levels,
                   const WaveCube* __restrict__ cubes, int nt, float* __restrict__ ub_out, float* __restrict__ lb_out,
                   int* __restrict__ incomp_mm, int* __restrict__ fpfh_mm, float* __restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const PairDev& P = pairs[pair];
    const GridDev& g = P.g;
    const int Nd = P.Nd;
    const int ncp1 = g.ncells + 1;
    const float* __restrict__ dist = g.dist;
    float* md = scratch + (size_t)wid * Nd;   // per-warp residual buffer (trimmed case only)
    for (int k = wid; k < nt; k += nw) {
        const WaveCube cb = cubes[k];
        const float* R = Rs + 9 * cb.rot;
        const int level = levels[cb.rot];
        const float r0 = R[0], r1 = R[1], r2 = R[2], r3 = R[3], r4 = R[4], r5 = R[5], r6 = R[6], r7 = R[7], r8 = R[8];
        const float half = cb.w / 2;
        const float transX = cb.x + half, transY = cb.y + half, transZ = cb.z + half;
        const float mtd = (float)(GOICP_SQRT3 / 2.0 * cb.w);
        float su = 0.f, sl = 0.f;
        int bad[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float fs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = lane; i < Nd; i += 32) {
            const float x = P.dx[i], y = P.dy[i], z = P.dz[i];
            const float px = r0 * x + r1 * y + r2 * z, py = r3 * x + r4 * y + r5 * z, pz = r6 * x + r7 * y + r8 * z;
            float d = P.weights[i] * dt_distance(g, dist, px + transX, py + transY, pz + transZ);
            if (level >= 0) d = d - P.maxRotDis[(size_t)level * Nd + i];
            if (d < 0.f) d = 0.f;
            if (P.doTrim) md[i] = d;
            else {
                su += (P.norm == 2) ? d * d : d;
                const float dis = d - mtd;
                if (dis > 0.f) sl += (P.norm == 2) ? dis * dis : dis;
            }
            if (P.use_reg || P.use_fpfh) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float cx = cb.x + (float)(c & 1) * cb.w, cy = cb.y + (float)((c >> 1) & 1) * cb.w, cz = cb.z + (float)((c >> 2) & 1) * cb.w;
                    const int cell = clamp_cell(g, px + cx, py + cy, pz + cz);
                    if (P.use_reg) bad[c] += ((__ldg(g.cmask + cell) >> P.dprop[i]) & 1u) ? 0 : 1;
                    if (P.use_fpfh) fs[c] += __ldg(P.fpfhD + (size_t)i * ncp1 + cell);
                }
            }
        }
        if (P.doTrim) { __syncwarp(); warp_trimmed_sums(md, Nd, P.inlierNum, lane, P.norm, mtd, &su, &sl); __syncwarp(); }
        else { su = warp_sum(su); sl = warp_sum(sl); }
        int minI = 0, maxI = 0; float minF = 0.f, maxF = 0.f;
        if (P.use_reg || P.use_fpfh) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (P.use_fpfh) { const float f = (float)(int)(warp_sum(fs[c]) / (float)Nd); if (c == 0) { minF = maxF = f; } else { maxF = fmaxf(maxF, f); minF = fminf(minF, f); } }
                if (P.use_reg) { const int n = warp_sum_i(bad[c]); if (c == 0) { minI = maxI = n; } else { maxI = max(maxI, n); minI = min(minI, n); } }
            }
            if (P.use_reg) { su = su + P.reg * (float)(maxI * maxI); sl = sl + P.reg * (float)(minI * minI); }
            if (P.use_fpfh) { su = su + P.regF * (maxF * maxF); sl = sl + P.regF * (minF * minF); }
        }
        if (lane == 0) {
            ub_out[k] = su; lb_out[k] = sl;
            if (incomp_mm) { incomp_mm[2 * k] = minI; incomp_mm[2 * k + 1] = maxI; }
            if (fpfh_mm) { fpfh_mm[2 * k] = (int)minF; fpfh_mm[2 * k + 1] = (int)maxF; }
        }
    }
}

}  // namespace

// ---- launchers ---------------------------------------------------------------------------------------------
size_t goicp_bnb_smem_floats(int NdP, int NdQ, bool exact, bool needMd, bool needFp) {
    const size_t n = (size_t)5 * NdP + (size_t)(NdP >> 2) + (size_t)43 * (NdP >> 5) + (needMd ? (size_t)(exact ? 16 : 8) * NdQ : 0) + (needFp ? (size_t)27 * NdQ : 0);
    return n > 3 * 512 ? n : 3 * 512;   // an ICP request tiles the model cloud through the same region (icp_device.cuh NN_TILE)
}

static int g_bnb_attr_set[16] = {0};
typedef void (*bnb_kernel_t)(const PairDev*, const InnerProb*, InnerOut*, int, int*, HeapEnt*, int, float*, size_t, int, int, int, QueueDev, uint4*, int, unsigned*, int, int);
static bnb_kernel_t bnb_kernel(int exact, int persist, int smem = 1) {
    if (smem == 2) {   // staging arrays AND the DT volume in shared memory
        if (persist) return exact ? inner_bnb_kernel<true, true, true, true> : inner_bnb_kernel<false, true, true, true>;
        return exact ? inner_bnb_kernel<true, false, true, true> : inner_bnb_kernel<false, false, true, true>;
    }
    if (smem) {
        if (persist) return exact ? inner_bnb_kernel<true, true, true, false> : inner_bnb_kernel<false, true, true, false>;
        return exact ? inner_bnb_kernel<true, false, true, false> : inner_bnb_kernel<false, false, true, false>;
    }
    if (persist) return exact ? inner_bnb_kernel<true, true, false, false> : inner_bnb_kernel<false, true, false, false>;
    return exact ? inner_bnb_kernel<true, false, false, false> : inner_bnb_kernel<false, false, false, false>;
}
static cudaError_t bnb_attr(int exact, int persist, int smem = 1) {
    const int k = (exact ? 1 : 0) + (persist ? 2 : 0) + 4 * smem;
    if (g_bnb_attr_set[k]) return cudaSuccess;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, bnb_kernel(exact, persist, smem));
    if (e != cudaSuccess) return e;
    const int maxDyn = 227 * 1024 - (int)fa.sharedSizeBytes - 1024;   // static (queue top, ICP tiles) + dynamic <= 227 KB per CTA
    e = cudaFuncSetAttribute(bnb_kernel(exact, persist, smem), cudaFuncAttributeMaxDynamicSharedMemorySize, maxDyn);
    if (e == cudaSuccess) g_bnb_attr_set[k] = 1;
    return e;
}

cudaError_t goicp_launch_inner_bnb(const PairDev* pairs, const InnerProb* probs, InnerOut* outs, int nprob, int* counter,
                                   HeapEnt* heaps, int heapCap, int maxCtas, float* gscratch, size_t gstride,
                                   int NdP, int NdQ, size_t smemBytes, int useSmem, int gridOff, int S3p, int exact, int threads, void* memo, int memoCap, unsigned* genCounter, cudaStream_t st, int* ctasLaunched) {
    if (nprob <= 0) { if (ctasLaunched) *ctasLaunched = 0; return cudaSuccess; }
    const size_t smem = useSmem ? smemBytes : 0;
    cudaError_t e = bnb_attr(exact, 0, useSmem);
    if (e != cudaSuccess) return e;
    int grid = nprob < maxCtas ? nprob : maxCtas;
    if (ctasLaunched) *ctasLaunched = grid;
    QueueDev q{};
    bnb_kernel(exact, 0, useSmem)<<<grid, threads, smem, st>>>(pairs, probs, outs, nprob, counter, heaps, heapCap, gscratch, gstride, NdP, NdQ, useSmem, q, (uint4*)memo, memoCap, genCounter, gridOff, S3p);
    return cudaGetLastError();
}

// the resident kernel of a batch: `ctas` CTAs serve the request ring until each has seen a shut-down marker
cudaError_t goicp_launch_inner_bnb_persistent(const PairDev* pairs, const QueueDev& q, HeapEnt* heaps, int heapCap, int ctas, float* gscratch, size_t gstride,
                                              int NdP, int NdQ, size_t smemBytes, int useSmem, int gridOff, int S3p, int exact, int threads, void* memo, int memoCap, unsigned* genCounter, cudaStream_t st) {
    const size_t smem = useSmem ? smemBytes : 0;
    cudaError_t e = bnb_attr(exact, 1, useSmem);
    if (e != cudaSuccess) return e;
    bnb_kernel(exact, 1, useSmem)<<<ctas, threads, smem, st>>>(pairs, nullptr, nullptr, 0, nullptr, heaps, heapCap, gscratch, gstride, NdP, NdQ, useSmem, q, (uint4*)memo, memoCap, genCounter, gridOff, S3p);
    return cudaGetLastError();
}

int goicp_inner_bnb_persistent_occupancy(size_t smemBytes, int exact, int threads, int useSmem) {
    if (bnb_attr(exact, 1, useSmem) != cudaSuccess) return 1;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bnb_kernel(exact, 1, useSmem), threads, useSmem ? smemBytes : 0) != cudaSuccess || n < 1) n = 1;
    return n;
}
int goicp_inner_bnb_occupancy(size_t smemBytes, int exact, int threads, int useSmem) {
    if (bnb_attr(exact, 0, useSmem) != cudaSuccess) return 1;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bnb_kernel(exact, 0, useSmem), threads, useSmem ? smemBytes : 0) != cudaSuccess || n < 1) n = 1;
    return n;
}

cudaError_t goicp_launch_eval_bounds(const PairDev* pairs, int pair, const float* Rs, const int* levels, const WaveCube* cubes,
                                     int nt, float* ub, float* lb, int* incomp_mm, int* fpfh_mm, float* scratch, int nwarps,
                                     cudaStream_t st) {
    if (nt <= 0) return cudaSuccess;
    const int blocks = (nwarps + 7) / 8;
    eval_bounds_kernel<<<blocks, 256, 0, st>>>(pairs, pair, Rs, levels, cubes, nt, ub, lb, incomp_mm, fpfh_mm, scratch);
    return cudaGetLastError();
}

// forces the (lazily loaded) kernels of this file into the context: a first launch while a resident kernel is spinning
// would otherwise wait for that kernel (CUDA lazy module loading)
cudaError_t goicp_preload_bnb() {
    cudaFuncAttributes a; cudaError_t e;
    if ((e = cudaFuncGetAttributes(&a, inner_bnb_kernel<true, true, true, true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, inner_bnb_kernel<false, true, true, true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, inner_bnb_kernel<true, false, true, true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, inner_bnb_kernel<false, false, true, true>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, inner_bnb_kernel<true, true, true, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, inner_bnb_kernel<true, true, false, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, inner_bnb_kernel<false, true, true, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, inner_bnb_kernel<true, false, true, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, inner_bnb_kernel<false, false, true, false>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, eval_bounds_kernel)) != cudaSuccess) return e;
    return cudaSuccess;
}
