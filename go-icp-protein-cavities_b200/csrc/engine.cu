// Host engine + C ABI of libgoicp_b200.so (include/goicp_b200.h).
//
// The rotation branch-and-bound of GoICP::OuterBnB (jly_goicp.cpp:582-876) stays on the host as a priority queue that
// follows the reference's pop order exactly; what it needs from the device are InnerBnB results (one CTA per call,
// k_bnb.cu), ICP refinements and pose scores (k_icp.cu).  To fill the GPU, every wave also evaluates SPECULATIVELY the
// InnerBnB calls the reference would make next if the incumbent error does not change (the rest of the current
// parent's children and the next `spec_width` queue nodes in pop order); a speculative result is consumed only when
// the reference's own order reaches that call with the same entry error, so results, traces and node counters are
// those of the sequential algorithm.  Pairs of a batch advance in lock-step waves that share each launch.
//
// There is no CPU fallback anywhere in this file: every numeric result comes from a kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <cmath>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <thread>
#include <map>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/goicp_b200.h"
#include "goicp_dev.h"
#include "launch.h"
#include "search_dev.h"

namespace {

using clk = std::chrono::steady_clock;
static double secs_since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

static std::string g_create_error;

// colour codes of the `properties` enum (transformation.hpp:36) that are keys of the identity compatibility map
// (jly_goicp.cpp:66-73); C = 1 is not a key.
static const int KNOWN_PROPS[8] = {8204959, 30894, 15219528, 15231913, 4646984, 16741671, 7566712, 0};
static bool known_prop(int p) { for (int k = 0; k < 8; k++) if (KNOWN_PROPS[k] == p) return true; return false; }

#define ROUND_HOST(x) ((int)((x) + 0.5))   // jly_3ddt.cpp:30

static std::atomic<bool> g_no_device_alloc(false);   // set while a persistent kernel is resident: cudaMalloc/cudaFree would dead-lock on it
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (g_no_device_alloc.load()) return cudaErrorMemoryAllocation;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (g_no_device_alloc.load()) return cudaErrorMemoryAllocation;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes * 2 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// ROTNODE (jly_goicp.h:59-73) + a unique id for the speculation cache
// pinned host memory mapped into the device address space: kernels read requests / write results directly over the bus,
// so a wave is one launch + one wait (no copies, no memset)
struct MapBuf {
    void* h = nullptr; void* d = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (g_no_device_alloc.load()) return cudaErrorMemoryAllocation;
        if (h) cudaFreeHost(h);
        h = d = nullptr; cap = 0;
        size_t want = bytes * 2 + 4096;
        cudaError_t e = cudaHostAlloc(&h, want, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e != cudaSuccess) return e;
        if ((e = cudaHostGetDevicePointer(&d, h, 0)) != cudaSuccess) return e;
        cap = want;
        return cudaSuccess;
    }
    void release() { if (h) cudaFreeHost(h); h = d = nullptr; cap = 0; }
};

struct RNode { float a, b, c, w, ub, lb; int l; int id; };
static inline bool rnode_less(const RNode& n1, const RNode& n2) {   // operator< :64-71
    if (n1.lb != n2.lb) return n1.lb > n2.lb;
    return n1.w < n2.w;
}
// std::priority_queue<ROTNODE> as libstdc++ implements it; spelled out so that equal keys pop in the reference's order
// independently of the standard library this file is compiled against.
static void rheap_push(std::vector<RNode>& h, const RNode& val) {
    h.push_back(val);
    int hole = (int)h.size() - 1, parent = (hole - 1) / 2;
    while (hole > 0 && rnode_less(h[parent], val)) { h[hole] = h[parent]; hole = parent; parent = (hole - 1) / 2; }
    h[hole] = val;
}
static RNode rheap_pop(std::vector<RNode>& h) {
    RNode top = h[0];
    const int len = (int)h.size() - 1;
    if (len > 0) {
        RNode val = h[len];
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (rnode_less(h[child], h[child - 1])) child--;
            h[hole] = h[child]; hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); h[hole] = h[child - 1]; hole = child - 1; }
        int parent = (hole - 1) / 2;
        while (hole > 0 && rnode_less(h[parent], val)) { h[hole] = h[parent]; hole = parent; parent = (hole - 1) / 2; }
        h[hole] = val;
    }
    h.pop_back();
    return top;
}

struct CallRes { float entryOpt; float err; float tn[4]; int pops, subcubes; };

enum Phase { PH_START, PH_WAIT_INIT, PH_POP, PH_CHILD_UB, PH_WAIT_ICP, PH_CHILD_LB, PH_DONE };

struct Problem {
    // ---- inputs (host copies) ----
    int Nm = 0, NdAll = 0, Nd = 0, ncolours = 1;
    std::vector<float> mxyz, dxyz;       // AoS as given
    std::vector<int> mc, dc;
    std::vector<float> mf, df;           // N x 41 or empty
    // ---- grid (host-derived) ----
    goicp_dt_info info{};
    std::vector<int> cell_vox, cell_start, cell_pts, cell_colour;
    std::vector<uint32_t> cmask;
    std::vector<uint8_t> dprop, mprop, dknown;
    bool prepared = false, dt_built = false, initialized = false;
    // ---- device ----
    PairDev dev{};
    size_t inBytes = 0, workBytes = 0, inOff = 0, workOff = 0;
    // ---- search state ----
    Phase phase = PH_START;
    std::vector<RNode> q;
    float optError = 0; double optR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, optT[3] = {0, 0, 0}; int optComp = 0;
    long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::string trace;
    RNode par{}, child{}; int j = 0; float R[9]; float ubChild = 0; float lastLb = 0;
    std::unordered_map<unsigned long long, CallRes> cache;
    int nextId = 1, quiet = 0;
    int status = 0;
    double t_dt = 0, t_reg = 0;
    // persistent-queue mode: requests in flight
    struct PendReq { int slot; unsigned long long key; float entryOpt; bool both; };   // both: key is the upper-bound call, key | 1 the lower-bound call of the same request
    std::vector<PendReq> pend;
    std::unordered_set<unsigned long long> inflight;
    bool icpQueued = false;
    bool dirty = false;       // resident scheduler: queued for the worker's next pass
    int icpSlot[2] = {-1, -1};
    // ICP exchange
    bool icpPending = false;
    float icpErr = 0; double icpR[9], icpT[3]; int icpIncomp = 0, compatPose = 0; float initErr = 0;
};

static void tracef(std::string& s, const char* fmt, ...) {
    char buf[256]; va_list ap; va_start(ap, fmt); int n = vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (n > 0) s.append(buf, std::min(n, (int)sizeof buf - 1));
}

// host-side parallel loop (pre-processing of a batch: voxelisation, cell lists, staging)
template <class F> static void parallel_for(int n, F fn) {
    const int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), (unsigned)std::max(1, n / 16));
    if (nt <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&]() { for (;;) { const int i = next.fetch_add(1); if (i >= n) break; fn(i); } });
    for (auto& t : th) t.join();
}

}  // namespace

// Everything one stream of waves needs: a worker thread of a batch owns one, the handle's own stream has `main`.
struct WaveCtx {
    cudaStream_t stream = nullptr; bool ownStream = false;
    DevBuf dCounter, dHeaps, dBnbScratch, dIcp, dMemo;
    PinBuf hIcp;
    MapBuf mProbs, mOuts, mIcp;
    bool counterReady = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evDone = nullptr;
    float ms[5] = {0, 0, 0, 0, 0}; long long launches[5] = {0, 0, 0, 0, 0};
    long long waves = 0, callsLaunched = 0, callsUsed = 0;
    double tLogic = 0, tInnerEnq = 0, tInnerWait = 0, tIcp = 0;   // host seconds
    int heapCap = 1 << 14;
    int ctaCap = 0;   // 0: numSM x occupancy
    goicp_status init(bool own, cudaStream_t st) {
        ownStream = own; stream = st;
        if (own && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return GOICP_ERR_CUDA;
        if (cudaEventCreate(&ev0) != cudaSuccess || cudaEventCreate(&ev1) != cudaSuccess ||
            cudaEventCreateWithFlags(&evDone, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) return GOICP_ERR_CUDA;
        return GOICP_OK;
    }
    // waits without spinning a host core (worker threads outnumber cores)
    cudaError_t sync() {
        static const int mode = [] { const char* e = getenv("GOICP_SYNC"); return e ? atoi(e) : 0; }();   // 0 blocking event, 1 stream sync (spin), 2 query + yield
        if (mode == 1) return cudaStreamSynchronize(stream);
        cudaError_t e = cudaEventRecord(evDone, stream); if (e != cudaSuccess) return e;
        if (mode == 2) { while ((e = cudaEventQuery(evDone)) == cudaErrorNotReady) std::this_thread::yield(); return e; }
        return cudaEventSynchronize(evDone);
    }
    void release() {
        DevBuf* bufs[] = {&dCounter, &dHeaps, &dBnbScratch, &dIcp, &dMemo};
        for (DevBuf* b : bufs) b->release();
        hIcp.release(); mProbs.release(); mOuts.release(); mIcp.release();
        if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); if (evDone) cudaEventDestroy(evDone);
        ev0 = ev1 = evDone = nullptr;
        if (ownStream && stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};

struct goicp_handle_s {
    int device = 0; cudaStream_t stream = nullptr; bool ownStream = false; int numSM = 148;
    goicp_params params; bool haveParams = false;
    int exact_sums = 1, spec_width = 32, use_dt_replay = 1;
    int groups = 0, slots = 0;   // 0 = auto
    int residentCtas = 0; bool tail_spec = true; int tail_spec_mult = 1, tail_thr = 1;   // CTAs of the running resident kernel (0: none)
    int batch_spec_width = 4;    // speculation width inside a batch (pairs already fill the GPU)
    bool dtUploaded = false;   // the DT came from goicp_dt_upload (test hook): no 16-bit distance codes
    bool merge_calls = true;   // resident scheduler: one request per rotation cube carries its upper- and lower-bound InnerBnB calls
    int bnb_threads = goicp_bnb_default_threads(); bool bnb_threads_set = false;   // threads per InnerBnB CTA (64..512): more threads = shorter pops, fewer resident calls   // batch: worker streams (0 = auto) and pairs advanced in lock-step per stream
    std::vector<Problem> probs;
    DevBuf arenaIn, arenaWork, dPairs, dTmp, dTmp2, dTmp3, dSepBits, dSepNx, dSepNxy;
    PinBuf hStage, hPairs;
    WaveCtx main;
    MapBuf qOuts, qOrder, qIcp, qDone; DevBuf qClaim, qHeaps, qScratch, qMemo, dGen;   // persistent-queue mode (batches)
    DevBuf sCtl, sHdrs, sSlots, sStates, sRq, sIcp, sOuts; PinBuf hOuts;   // device-resident search (k_search.cu)
    int resident_search = 1;     // 1: OuterBnB runs on the device (k_search.cu); 0: host state machine + request ring (the round-1 scheduler)
    int spec_groups = SR_NGROUP - 4;   // device-resident search: most rotation-queue nodes with speculative calls per owner CTA (0: never speculate)
    int shardRank = 0, shardN = 1; goicp_allgather_fn allgather = nullptr; void* allgatherUser = nullptr;   // frontier sharding
    std::vector<InnerOut> xSend, xRecv;
    std::atomic<int> outstanding{0};   // requests published and not yet harvested (persistent scheduler)
    std::atomic<int> activePairs{0};   // pairs currently being searched (persistent scheduler): few left -> speculate wider
    int persistent_single = 1;   // single registrations of small clouds also go through the resident kernel (no launches per wave)
    int persistent = 1;          // batches: 1 = resident kernel + request ring, 0 = one launch per wave
    std::vector<std::unique_ptr<WaveCtx>> workers;
    std::mutex errMutex;
    std::string err, trace;
    float ms[5] = {0, 0, 0, 0, 0}; long long launches[5] = {0, 0, 0, 0, 0};
    double stats[16] = {0};   // see goicp_get_stats
};

namespace {

typedef goicp_handle_s Eng;

static goicp_status fail(Eng* h, goicp_status s, const char* fmt, ...) {
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) { std::lock_guard<std::mutex> lk(h->errMutex); h->err = buf; } else g_create_error = buf;
    return s;
}
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(h, GOICP_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

static size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

struct EvTimer {   // CUDA-event time of a kernel group on a wave context's stream
    WaveCtx& c; int slot;
    EvTimer(WaveCtx& c_, int slot_) : c(c_), slot(slot_) { cudaEventRecord(c.ev0, c.stream); }
    void stop(int nlaunch) { cudaEventRecord(c.ev1, c.stream); cudaEventSynchronize(c.ev1); float ms = 0; cudaEventElapsedTime(&ms, c.ev0, c.ev1); c.ms[slot] += ms; c.launches[slot] += nlaunch; }
};

// ---- host preprocessing of one pair: bbox/scale (jly_3ddt.cpp:899-931), seeding (:976-995), cell lists and colour
//      masks (assignCellColor jly_goicp.cpp:951-969, checkProperty :1068-1092) -----------------------------------------
static goicp_status prepare_problem(Eng* h, Problem& P) {
    const goicp_params& p = h->params;
    const int S = p.distTransSize, num = P.Nm;
    if (S < 2 || S > 1024) return fail(h, GOICP_ERR_UNSUPPORTED, "distTransSize %d outside [2,1024]", S);
    if (num < 1 || P.NdAll < 1) return fail(h, GOICP_ERR_ARG, "empty cloud (Nm=%d Nd=%d)", num, P.NdAll);
    const float* m = P.mxyz.data();
    double xMin = m[0], xMax = m[0], yMin = m[1], yMax = m[1], zMin = m[2], zMax = m[2];
    for (int i = 1; i < num; i++) {
        double x = m[3 * i], y = m[3 * i + 1], z = m[3 * i + 2];
        if (xMin > x) xMin = x; if (xMax < x) xMax = x;
        if (yMin > y) yMin = y; if (yMax < y) yMax = y;
        if (zMin > z) zMin = z; if (zMax < z) zMax = z;
    }
    const double ef = p.distTransExpandFactor;
    const double xC = (xMin + xMax) / 2, yC = (yMin + yMax) / 2, zC = (zMin + zMax) / 2;
    xMin = xC - ef * (xMax - xC); xMax = xC + ef * (xMax - xC);
    yMin = yC - ef * (yMax - yC); yMax = yC + ef * (yMax - yC);
    zMin = zC - ef * (zMax - zC); zMax = zC + ef * (zMax - zC);
    double mx = xMax - xMin > yMax - yMin ? xMax - xMin : yMax - yMin;
    mx = mx > zMax - zMin ? mx : zMax - zMin;
    xMin = xC - mx / 2; xMax = xC + mx / 2; yMin = yC - mx / 2; yMax = yC + mx / 2; zMin = zC - mx / 2; zMax = zC + mx / 2;
    P.info.xMin = xMin; P.info.xMax = xMax; P.info.yMin = yMin; P.info.yMax = yMax; P.info.zMin = zMin; P.info.zMax = zMax;
    P.info.scale = S / mx; P.info.size = S;
    const double scale = P.info.scale;
    if (!(mx > 0)) return fail(h, GOICP_ERR_ARG, "degenerate model bounding box");

    // colour dictionary (<= 32 distinct colours over both clouds)
    std::map<int, int> dict;
    for (int i = 0; i < num; i++) dict.emplace(P.mc.empty() ? 0 : P.mc[i], 0);
    for (int i = 0; i < P.NdAll; i++) dict.emplace(P.dc.empty() ? 0 : P.dc[i], 0);
    if (dict.size() > 32) return fail(h, GOICP_ERR_UNSUPPORTED, "more than 32 distinct colour codes in one pair (%zu)", dict.size());
    { int k = 0; for (auto& kv : dict) kv.second = k++; }
    P.ncolours = (int)dict.size();
    P.mprop.resize(num); P.dprop.resize(P.NdAll); P.dknown.resize(P.NdAll);
    for (int i = 0; i < num; i++) P.mprop[i] = (uint8_t)dict[P.mc.empty() ? 0 : P.mc[i]];
    for (int i = 0; i < P.NdAll; i++) { int c = P.dc.empty() ? 0 : P.dc[i]; P.dprop[i] = (uint8_t)dict[c]; P.dknown[i] = known_prop(c) ? 1 : 0; }

    // seeding: voxel of every model point, cells = occupied voxels in ascending voxel order, points in index order
    std::vector<std::pair<int, int>> vp; vp.reserve(num);
    for (int i = 0; i < num; i++) {
        int x = ROUND_HOST(((double)m[3 * i] - xMin) * scale), y = ROUND_HOST(((double)m[3 * i + 1] - yMin) * scale), z = ROUND_HOST(((double)m[3 * i + 2] - zMin) * scale);
        if (x < 0 || x >= S || y < 0 || y >= S || z < 0 || z >= S) continue;   // :989 (only reachable for expandFactor <= 1)
        vp.emplace_back((z * S + y) * S + x, i);
    }
    std::sort(vp.begin(), vp.end());
    P.cell_vox.clear(); P.cell_start.clear(); P.cell_pts.clear(); P.cell_colour.clear(); P.cmask.clear();
    for (size_t k = 0; k < vp.size(); k++) {
        if (k == 0 || vp[k].first != vp[k - 1].first) { P.cell_vox.push_back(vp[k].first); P.cell_start.push_back((int)k); }
        P.cell_pts.push_back(vp[k].second);
    }
    P.cell_start.push_back((int)vp.size());
    const int nc = (int)P.cell_vox.size();
    P.info.ncells = nc;
    P.cell_colour.resize(nc); P.cmask.assign(nc + 1, 0u);
    for (int c = 0; c < nc; c++) {
        const int b = P.cell_start[c], e = P.cell_start[c + 1];
        auto col = [&](int k) { return P.mc.empty() ? 0 : P.mc[P.cell_pts[k]]; };
        int prop = col(b); bool mixed = false; uint32_t orbits = 0;
        for (int k = b; k < e; k++) { if (col(k) != prop) mixed = true; orbits |= 1u << dict[col(k)]; }
        P.cell_colour[c] = mixed ? -1 : prop;
        P.cmask[c] = mixed ? orbits : (known_prop(prop) ? (1u << dict[prop]) : 0u);   // checkProperty :1068-1092
    }
    P.prepared = true; P.dt_built = false; P.initialized = false;
    return GOICP_OK;
}

static bool need_corner_terms(const goicp_params& p) { return p.regularization > 0 || p.regularizationNeighbors > 0 || (p.regularizationFPFH > 0 && p.cfpfh != 0); }

// ---- device layout + upload of all problems ----------------------------------------------------------------------
static goicp_status upload_problems(Eng* h) {
    const goicp_params& p = h->params;
    const int S = p.distTransSize; const size_t S3 = (size_t)S * S * S;
    const bool useF = p.cfpfh != 0;
    const bool wantVcell = true;
    size_t inTot = 0, workTot = 0;
    for (auto& P : h->probs) {
        if (useF && (P.mf.empty() || P.df.empty())) return fail(h, GOICP_ERR_ARG, "cfpfh=%d needs c-FPFH descriptors for both clouds", p.cfpfh);
        const int nc = P.info.ncells;
        size_t in = 0;
        in += 3 * al256(sizeof(float) * P.NdAll) + 3 * al256(sizeof(float) * P.Nm);
        in += 2 * al256(P.NdAll) + al256(P.Nm);
        if (useF) in += al256(sizeof(float) * 41 * (size_t)P.NdAll) + al256(sizeof(float) * 41 * (size_t)P.Nm);
        in += al256(sizeof(int) * std::max(nc, 1)) + al256(sizeof(uint32_t) * (nc + 1)) + al256(sizeof(int) * (nc + 1)) + al256(sizeof(int) * P.Nm);
        P.inBytes = in; P.inOff = inTot; inTot += in;
        size_t w = 0;
        w += al256(sizeof(float) * S3) + al256(sizeof(int) * S3) + (wantVcell ? 2 * al256(sizeof(int) * S3) + al256(S3 + 16) : 0) + al256(sizeof(double) * GOICP_OVN)
             + (S <= 32 ? al256(2 * (S3 + 16)) + al256(sizeof(float) * (3 * (size_t)(S - 1) * (S - 1) + 6)) : 0);
        w += 2 * al256(sizeof(float) * P.NdAll) + al256(sizeof(float) * GOICP_MAXROTLEVEL * (size_t)P.NdAll);
        if (useF && p.regularizationFPFH > 0) w += al256(sizeof(float) * (size_t)P.NdAll * (nc + 1));
        if (p.regularizationNeighbors > 0) w += al256(sizeof(int) * P.NdAll) + al256(sizeof(int) * P.Nm);
        w += al256(sizeof(unsigned long long) * P.NdAll) + al256(sizeof(int) * P.NdAll) + al256(sizeof(float) * 8 * (size_t)std::max(P.NdAll, 1));
        if (!(p.trimFraction < 0.001)) w += al256(sizeof(unsigned long long) * 2048);
        P.workBytes = w; P.workOff = workTot; workTot += w;
    }
    CU(h->arenaIn.ensure(inTot));
    CU(h->arenaWork.ensure(workTot));
    CU(h->hStage.ensure(inTot));
    char* stage = h->hStage.as<char>();
    char* dIn = h->arenaIn.as<char>(); char* dWork = h->arenaWork.as<char>();
    parallel_for((int)h->probs.size(), [&](int pi) {
        Problem& P = h->probs[pi];
        const int nc = P.info.ncells;
        size_t o = P.inOff;
        PairDev& D = P.dev;
        memset(&D, 0, sizeof D);
        auto putf = [&](float*& dptr, size_t count, auto fill) { dptr = reinterpret_cast<float*>(dIn + o); fill(reinterpret_cast<float*>(stage + o)); o += al256(sizeof(float) * count); };
        putf(D.dx, P.NdAll, [&](float* s) { for (int i = 0; i < P.NdAll; i++) s[i] = P.dxyz[3 * i]; });
        putf(D.dy, P.NdAll, [&](float* s) { for (int i = 0; i < P.NdAll; i++) s[i] = P.dxyz[3 * i + 1]; });
        putf(D.dz, P.NdAll, [&](float* s) { for (int i = 0; i < P.NdAll; i++) s[i] = P.dxyz[3 * i + 2]; });
        putf(D.mx, P.Nm, [&](float* s) { for (int i = 0; i < P.Nm; i++) s[i] = P.mxyz[3 * i]; });
        putf(D.my, P.Nm, [&](float* s) { for (int i = 0; i < P.Nm; i++) s[i] = P.mxyz[3 * i + 1]; });
        putf(D.mz, P.Nm, [&](float* s) { for (int i = 0; i < P.Nm; i++) s[i] = P.mxyz[3 * i + 2]; });
        D.dprop = reinterpret_cast<uint8_t*>(dIn + o); memcpy(stage + o, P.dprop.data(), P.NdAll); o += al256(P.NdAll);
        D.dknown = reinterpret_cast<uint8_t*>(dIn + o); memcpy(stage + o, P.dknown.data(), P.NdAll); o += al256(P.NdAll);
        D.mprop = reinterpret_cast<uint8_t*>(dIn + o); memcpy(stage + o, P.mprop.data(), P.Nm); o += al256(P.Nm);
        if (useF) {
            D.dfpfh = reinterpret_cast<float*>(dIn + o); memcpy(stage + o, P.df.data(), sizeof(float) * 41 * (size_t)P.NdAll); o += al256(sizeof(float) * 41 * (size_t)P.NdAll);
            D.mfpfh = reinterpret_cast<float*>(dIn + o); memcpy(stage + o, P.mf.data(), sizeof(float) * 41 * (size_t)P.Nm); o += al256(sizeof(float) * 41 * (size_t)P.Nm);
        }
        D.g.cell_vox = reinterpret_cast<int*>(dIn + o); memcpy(stage + o, P.cell_vox.data(), sizeof(int) * nc); o += al256(sizeof(int) * std::max(nc, 1));
        D.g.cmask = reinterpret_cast<uint32_t*>(dIn + o); memcpy(stage + o, P.cmask.data(), sizeof(uint32_t) * (nc + 1)); o += al256(sizeof(uint32_t) * (nc + 1));
        D.cell_start = reinterpret_cast<int*>(dIn + o); memcpy(stage + o, P.cell_start.data(), sizeof(int) * (nc + 1)); o += al256(sizeof(int) * (nc + 1));
        D.cell_pts = reinterpret_cast<int*>(dIn + o); memcpy(stage + o, P.cell_pts.data(), sizeof(int) * P.cell_pts.size()); o += al256(sizeof(int) * P.Nm);
        size_t w = P.workOff;
        D.g.dist = reinterpret_cast<float*>(dWork + w); w += al256(sizeof(float) * S3);
        D.g.vnear = reinterpret_cast<int*>(dWork + w); w += al256(sizeof(int) * S3);
        if (wantVcell) { D.g.vcell = reinterpret_cast<int*>(dWork + w); w += al256(sizeof(int) * S3); D.g.vmask = reinterpret_cast<uint32_t*>(dWork + w); w += al256(sizeof(int) * S3); D.g.vmask8 = reinterpret_cast<uint8_t*>(dWork + w); w += al256(S3 + 16); }
        D.g.ovl = reinterpret_cast<double*>(dWork + w); w += al256(sizeof(double) * GOICP_OVN);
        if (S <= 32) {
            D.g.dcode = reinterpret_cast<uint16_t*>(dWork + w); w += al256(2 * (S3 + 16));
            D.g.dlut = reinterpret_cast<float*>(dWork + w); w += al256(sizeof(float) * (3 * (size_t)(S - 1) * (S - 1) + 6));
            D.g.nlut = 3 * (S - 1) * (S - 1) + 2;
        }
        D.normData = reinterpret_cast<float*>(dWork + w); w += al256(sizeof(float) * P.NdAll);
        D.weights = reinterpret_cast<float*>(dWork + w); w += al256(sizeof(float) * P.NdAll);
        D.maxRotDis = reinterpret_cast<float*>(dWork + w); w += al256(sizeof(float) * GOICP_MAXROTLEVEL * (size_t)P.NdAll);
        if (useF && p.regularizationFPFH > 0) { D.fpfhD = reinterpret_cast<float*>(dWork + w); w += al256(sizeof(float) * (size_t)P.NdAll * (nc + 1)); }
        if (p.regularizationNeighbors > 0) { D.nbD = reinterpret_cast<int*>(dWork + w); w += al256(sizeof(int) * P.NdAll); D.nbM = reinterpret_cast<int*>(dWork + w); w += al256(sizeof(int) * P.Nm); }
        D.nn = reinterpret_cast<unsigned long long*>(dWork + w); w += al256(sizeof(unsigned long long) * P.NdAll);
        D.order = reinterpret_cast<int*>(dWork + w); w += al256(sizeof(int) * P.NdAll);
        D.scratch = reinterpret_cast<float*>(dWork + w); w += al256(sizeof(float) * 8 * (size_t)std::max(P.NdAll, 1));
        if (!(p.trimFraction < 0.001)) { D.sortKeys = reinterpret_cast<unsigned long long*>(dWork + w); w += al256(sizeof(unsigned long long) * 2048); }
        D.g.S = S; D.g.ncells = nc; D.g.xMin = P.info.xMin; D.g.yMin = P.info.yMin; D.g.zMin = P.info.zMin; D.g.scale = P.info.scale;
        D.Nm = P.Nm; D.Nd = P.Nd; D.NdAll = P.NdAll;
    });
    CU(cudaMemcpyAsync(dIn, stage, inTot, cudaMemcpyHostToDevice, h->stream));
    return GOICP_OK;
}

// fills the parameter-derived fields of every PairDev (GoICP::Initialize :180-267 scalars) and uploads the array
static goicp_status upload_pairdevs(Eng* h) {
    const goicp_params& p = h->params;
    const bool doTrim = !(p.trimFraction < 0.001);   // GoICP() :54 sets true; readConfig clears it (jly_main.cpp:259)
    for (auto& P : h->probs) {
        PairDev& D = P.dev;
        D.Nd = P.Nd;
        D.doTrim = doTrim ? 1 : 0;
        D.inlierNum = doTrim ? (int)(P.Nd * (1 - p.trimFraction)) : P.Nd;           // :244-252
        D.norm = p.norm; D.cfpfh = p.cfpfh; D.ponderation = p.ponderation;
        D.fpfh_b = 0; D.fpfh_e = 0;
        if (p.cfpfh == 1) D.fpfh_e = 41; else if (p.cfpfh == 2) D.fpfh_e = 33; else if (p.cfpfh == 3) { D.fpfh_b = 33; D.fpfh_e = 41; }
        D.use_reg = p.regularization > 0 ? 1 : 0;
        D.use_fpfh = (p.regularizationFPFH > 0 && p.cfpfh != 0) ? 1 : 0;
        D.use_nb = p.regularizationNeighbors > 0 ? 1 : 0;
        D.reg = p.regularization; D.regF = p.regularizationFPFH; D.regN = p.regularizationNeighbors;
        D.MSEThresh = p.MSEThresh; D.trimFraction = p.trimFraction;
        D.SSEThresh = p.MSEThresh * D.inlierNum;                                     // :266
        D.tMinX = p.transMinX; D.tMinY = p.transMinY; D.tMinZ = p.transMinZ; D.tWidth = p.transWidth;
        {   // FP32 fast path of the voxel index (goicp_dev.h: GridDev.vf*): fraction bits kept and the ambiguity zone around
            // each rounding boundary, from a bound on the float evaluation error of (v - min) * scale + 0.5 inside the grid
            GridDev& g = D.g;
            const int S = g.S;
            int lg = 0; while ((1 << lg) < S + GOICP_OVLIM + 2) lg++;
            const int f = std::min(16, 22 - lg);
            const double ulp = std::ldexp(1.0, -f);
            const double ext = (double)(S + GOICP_OVLIM + 1) / g.scale, lo = (double)(GOICP_OVLIM + 1) / g.scale;
            const double A = std::max({std::fabs(g.xMin - lo), std::fabs(g.xMin + ext), std::fabs(g.yMin - lo), std::fabs(g.yMin + ext), std::fabs(g.zMin - lo), std::fabs(g.zMin + ext)});
            const double T = std::max({std::fabs((double)p.transMinX), std::fabs((double)p.transMinX + p.transWidth), std::fabs((double)p.transMinY), std::fabs((double)p.transMinY + p.transWidth),
                                       std::fabs((double)p.transMinZ), std::fabs((double)p.transMinZ + p.transWidth)});
            // |v| <= A for a voxel at most GOICP_OVLIM outside the grid, |p| = |v - trans| <= A + T; terms: float rounding of v = p + trans, of the scale, of C and of the fma
            const double err = g.scale * std::ldexp(1.0, -24) * (2 * A + T) * 1.25 + ulp + 1e-9;
            const int E = (int)std::ceil(err / ulp) + 1;
            g.vfScale = (float)g.scale; g.vfShift = f; g.vfMask = (1u << f) - 1u;
            const float magic = (float)std::ldexp(1.5, 23 - f);
            unsigned mb; memcpy(&mb, &magic, 4);
            g.vfBias = mb >> f;
            g.vfMagic = (double)magic + 0.5 + E * ulp;
            g.vfZone = (f >= 8 && (2 * E + 1) * 64 < (1 << f) && std::isfinite(err)) ? (unsigned)(2 * E + 1) : 0xFFFFFFFFu;
            if (getenv("GOICP_NO_VOXFAST")) g.vfZone = 0xFFFFFFFFu;
        }
        for (int l = 0; l < GOICP_MAXROTLEVEL; l++) {                                // :195-204, host libm as the reference
            float sigma = (float)(p.rotWidth / pow(2.0, l) / 2.0);
            float maxAngle = (float)(GOICP_SQRT3 * sigma);
            if (maxAngle > GOICP_PI) maxAngle = (float)GOICP_PI;
            D.s2[l] = 2 * sinf(maxAngle / 2);
        }
    }
    const size_t n = h->probs.size();
    CU(h->dPairs.ensure(sizeof(PairDev) * n));
    CU(h->hPairs.ensure(sizeof(PairDev) * n));
    PairDev* st = h->hPairs.as<PairDev>();
    for (size_t i = 0; i < n; i++) st[i] = h->probs[i].dev;
    CU(cudaMemcpyAsync(h->dPairs.p, st, sizeof(PairDev) * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

static goicp_status build_dt_all(Eng* h, bool replay) {
    const int S = h->params.distTransSize;
    auto t0 = clk::now();
    goicp_status s = upload_pairdevs(h);
    if (s) return s;
    EvTimer tm(h->main, 0);
    int nl = 0;
    if (replay) {
        if (S > 32) return fail(h, GOICP_ERR_UNSUPPORTED, "the 8SED replay builder supports distTransSize <= 32 (got %d)", S);
        CU(goicp_launch_dt_replay(h->dPairs.as<PairDev>(), 0, (int)h->probs.size(), S, h->stream)); nl = 1;
    } else {
        const int SW = (S + 31) / 32; const size_t S3 = (size_t)S * S * S;
        CU(h->dSepBits.ensure((size_t)S * S * SW * sizeof(unsigned)));
        CU(h->dSepNx.ensure(S3 * sizeof(unsigned short)));
        CU(h->dSepNxy.ensure(S3 * sizeof(unsigned)));
        for (auto& P : h->probs) { CU(goicp_launch_dt_separable(P.dev.g, h->dSepBits.as<unsigned>(), h->dSepNx.as<unsigned short>(), h->dSepNxy.as<unsigned>(), h->numSM, h->stream)); nl += 4; }
    }
    tm.stop(nl);
    CU(cudaGetLastError());
    const double dt = secs_since(t0) / std::max<size_t>(1, h->probs.size());
    for (auto& P : h->probs) { P.dt_built = true; P.t_dt = dt; }
    return GOICP_OK;
}

static goicp_status initialize_all(Eng* h) {
    const goicp_params& p = h->params;
    for (auto& P : h->probs) {
        if (!P.dt_built) return fail(h, GOICP_ERR_ARG, "initialize before build_dt");
        if (p.ponderation == 1 && P.Nd < 20) return fail(h, GOICP_ERR_UNSUPPORTED, "ponderation=1 needs Nd >= 20 (neighborsWeights never terminates below; Nd=%d)", P.Nd);
        if (p.norm != 1 && p.norm != 2) return fail(h, GOICP_ERR_UNSUPPORTED, "norm must be 1 or 2");
    }
    goicp_status s = upload_pairdevs(h);
    if (s) return s;
    EvTimer tm(h->main, 1);
    int nl = 1;
    CU(goicp_launch_initialize(h->dPairs.as<PairDev>(), 0, (int)h->probs.size(), h->stream));
    if (p.regularizationFPFH > 0 && p.cfpfh != 0) { CU(goicp_launch_fpfh_table(h->dPairs.as<PairDev>(), 0, (int)h->probs.size(), 32, h->stream)); nl++; }
    if (p.regularizationNeighbors > 0) {   // assignNeighbors (BuildDT :94); both clouds, every source point
        int maxN = 1; for (auto& P : h->probs) maxN = std::max(maxN, P.NdAll + P.Nm);
        CU(goicp_launch_assign_neighbors(h->dPairs.as<PairDev>(), 0, (int)h->probs.size(), std::min(64, (maxN + 255) / 256), h->stream)); nl++;
    }
    tm.stop(nl);
    CU(cudaGetLastError());
    for (auto& P : h->probs) P.initialized = true;
    return GOICP_OK;
}

// ---- one launch of InnerBnB calls (handles heap overflow by re-running the overflowed calls with larger heaps) -------
struct BnbCfg { int NdP, NdQ; size_t smemFloats, smemBytes; int useSmem, perSM, threads, gridOff, S3p, ct; };
static BnbCfg bnb_config(Eng* h) {
    int maxNd = 1, maxCol = 1; bool anyTrim = false, anyF = false, anyNb = false;
    for (auto& P : h->probs) { maxNd = std::max(maxNd, P.Nd); anyTrim |= P.dev.doTrim != 0; anyF |= P.dev.use_fpfh != 0; anyNb |= P.dev.use_nb != 0; maxCol = std::max(maxCol, P.ncolours); }
    BnbCfg c;
    c.ct = (anyF || anyNb) ? 1 : 0;
    c.NdP = (maxNd + 31) & ~31; c.NdQ = c.NdP + 4;   // row stride = 4 mod 32: the chain lanes' float4 reads of the 8 rows hit 8 different bank quads
    const bool needMd = h->exact_sums || anyTrim, needFp = h->exact_sums && anyF;
    c.smemFloats = goicp_bnb_smem_floats(c.NdP, c.NdQ, h->exact_sums != 0, needMd, needFp);
    c.smemBytes = c.smemFloats * sizeof(float);
    c.useSmem = c.smemBytes <= 180 * 1024;   // + ~32 KB static in the resident kernel
    c.gridOff = 0; c.S3p = 0;
    // small volumes (cavity grids, 20^3): distances + one colour-mask byte per voxel are staged in shared memory per call
    const int S = h->params.distTransSize; const size_t S3 = (size_t)S * S * S;
    const size_t S3p = (S3 + 15) & ~(size_t)15;
    const size_t nlutP = ((size_t)3 * (S - 1) * (S - 1) + 2 + 3) & ~(size_t)3;
    if (c.useSmem && S <= 32 && maxCol <= 8 && !h->dtUploaded && !getenv("GOICP_NO_GRID_SMEM") && ((c.smemFloats + 3) & ~(size_t)3) * 4 + nlutP * 4 + S3p * 3 <= 100 * 1024) {
        c.gridOff = (int)((c.smemFloats + 3) & ~(size_t)3); c.S3p = (int)S3p;
        c.smemBytes = (size_t)c.gridOff * 4 + nlutP * 4 + S3p * 3;   // distance table + 16-bit distance codes + colour-mask bytes
        c.useSmem = 2;
    }
    // batches on shared-memory volumes: 192-thread CTAs, four per SM (measured +5 % over 256 x 3; a single registration keeps
    // the shorter pops of 256-thread CTAs)
    c.threads = (h->bnb_threads_set || !(c.useSmem == 2 && h->probs.size() > 1)) ? h->bnb_threads : std::min(192, h->bnb_threads);
    c.perSM = goicp_inner_bnb_occupancy(c.smemBytes, h->exact_sums, c.threads, c.useSmem, c.ct);
    return c;
}
static goicp_status run_inner_local(Eng* h, WaveCtx& c, const BnbCfg& cfg, std::vector<InnerProb>& reqs, std::vector<InnerOut>& outs);
// One wave of InnerBnB calls.  With frontier sharding the calls are dealt round-robin to the ranks, evaluated locally and
// exchanged with one all-gather, so every rank continues with the complete, identical result set.
static goicp_status run_inner(Eng* h, WaveCtx& c, const BnbCfg& cfg, std::vector<InnerProb>& reqs, std::vector<InnerOut>& outs) {
    const int n = (int)reqs.size();
    if (h->shardN <= 1 || !h->allgather) return run_inner_local(h, c, cfg, reqs, outs);
    outs.resize(n);
    const int N = h->shardN, r = h->shardRank, per = (n + N - 1) / N;
    std::vector<InnerProb> mine; std::vector<InnerOut> mineOut;
    for (int k = r; k < n; k += N) mine.push_back(reqs[k]);
    goicp_status s = run_inner_local(h, c, cfg, mine, mineOut);
    if (s) return s;
    h->xSend.assign(std::max(per, 1), InnerOut{}); h->xRecv.assign((size_t)std::max(per, 1) * N, InnerOut{});
    for (size_t k = 0; k < mineOut.size(); k++) h->xSend[k] = mineOut[k];
    if (h->allgather(h->xSend.data(), h->xRecv.data(), (int64_t)sizeof(InnerOut) * std::max(per, 1), h->allgatherUser) != 0)
        return fail(h, GOICP_ERR_ARG, "frontier sharding: the all-gather callback failed");
    for (int k = 0; k < n; k++) outs[k] = h->xRecv[(size_t)(k % N) * std::max(per, 1) + k / N];
    return GOICP_OK;
}
static goicp_status run_inner_local(Eng* h, WaveCtx& c, const BnbCfg& cfg, std::vector<InnerProb>& reqs, std::vector<InnerOut>& outs) {
    const int n = (int)reqs.size();
    outs.resize(n);
    if (n == 0) return GOICP_OK;
    int maxCtas = h->numSM * cfg.perSM;
    if (c.ctaCap > 0) maxCtas = std::min(maxCtas, c.ctaCap);
    int heapCap = c.heapCap;
    CU(c.mProbs.ensure(sizeof(InnerProb) * (size_t)n));
    CU(c.mOuts.ensure(sizeof(InnerOut) * (size_t)n));
    if (!c.counterReady) { CU(c.dCounter.ensure(2 * sizeof(int))); CU(cudaMemsetAsync(c.dCounter.p, 0, 2 * sizeof(int), c.stream)); c.counterReady = true; }
    std::vector<int> todo(n); for (int i = 0; i < n; i++) todo[i] = i;
    for (int attempt = 0; attempt < 12 && !todo.empty(); attempt++) {
        const int m = (int)todo.size();
        InnerProb* hp = reinterpret_cast<InnerProb*>(c.mProbs.h);
        for (int i = 0; i < m; i++) hp[i] = reqs[todo[i]];
        const int ctas = std::min(m, maxCtas);
        CU(c.dHeaps.ensure(sizeof(HeapEnt) * (size_t)ctas * heapCap));
        if (!cfg.useSmem) CU(c.dBnbScratch.ensure(sizeof(float) * cfg.smemFloats * (size_t)ctas));
        const int memoCap = 4096;
        if (c.dMemo.cap < (size_t)32 * memoCap * ctas) { CU(c.dMemo.ensure((size_t)32 * memoCap * ctas)); CU(cudaMemsetAsync(c.dMemo.p, 0, c.dMemo.cap, c.stream)); }
        auto tq = clk::now();
        cudaEventRecord(c.ev0, c.stream);
        int launched = 0;
        CU(goicp_launch_inner_bnb(h->dPairs.as<PairDev>(), reinterpret_cast<const InnerProb*>(c.mProbs.d), reinterpret_cast<InnerOut*>(c.mOuts.d), m, c.dCounter.as<int>(),
                                  c.dHeaps.as<HeapEnt>(), heapCap, ctas, c.dBnbScratch.as<float>(), cfg.smemFloats, cfg.NdP, cfg.NdQ, cfg.smemBytes, cfg.useSmem, cfg.gridOff, cfg.S3p,
                                  h->exact_sums, cfg.ct, cfg.threads, c.dMemo.p, memoCap, h->dGen.as<unsigned>(), c.stream, &launched));
        cudaEventRecord(c.ev1, c.stream);
        c.tInnerEnq += secs_since(tq); tq = clk::now();
        CU(c.sync());
        c.tInnerWait += secs_since(tq);
        float ms = 0; cudaEventElapsedTime(&ms, c.ev0, c.ev1); c.ms[2] += ms; c.launches[2] += 1;
        const InnerOut* ho = reinterpret_cast<const InnerOut*>(c.mOuts.h);
        std::vector<int> again;
        for (int i = 0; i < m; i++) { if (ho[i].status == 4) again.push_back(todo[i]); else outs[todo[i]] = ho[i]; }
        todo.swap(again);
        if (!todo.empty()) { heapCap *= 4; const size_t fit = ((size_t)4 << 30) / sizeof(HeapEnt) / (size_t)heapCap; maxCtas = (int)std::max<size_t>(1, std::min<size_t>((size_t)maxCtas, fit)); }
    }
    if (!todo.empty()) return fail(h, GOICP_ERR_OVERFLOW, "translation queue exceeded %d entries", heapCap);
    c.callsLaunched += n;
    return GOICP_OK;
}

// ---- ICP / scoring pipeline for a set of states ----------------------------------------------------------------------
static goicp_status run_icp(Eng* h, WaveCtx& c, std::vector<IcpState>& states) {
    const int n = (int)states.size();
    if (n == 0) return GOICP_OK;
    int maxNd = 1, maxNm = 1; bool anyIcp = false; bool small = true;
    for (auto& s : states) {
        const Problem& P = h->probs[s.pair]; maxNd = std::max(maxNd, P.Nd); maxNm = std::max(maxNm, P.Nm); anyIcp |= s.mode == 0;
        if ((size_t)P.Nd * P.Nm > ((size_t)1 << 21) || (P.dev.doTrim && P.Nd > 2048)) small = false;
    }
    { static const char* env = getenv("GOICP_ICP_FUSED"); if (env && env[0] == '0') small = false; }   // debugging aid
    if (small) {   // whole ICP (begin, every iteration, re-score) in one launch, one CTA per request; states in mapped host memory
        CU(c.mIcp.ensure(sizeof(IcpState) * n));
        IcpState* ms_ = reinterpret_cast<IcpState*>(c.mIcp.h);
        for (int i = 0; i < n; i++) ms_[i] = states[i];
        cudaEventRecord(c.ev0, c.stream);
        CU(goicp_launch_icp_fused(h->dPairs.as<PairDev>(), reinterpret_cast<IcpState*>(c.mIcp.d), n, c.stream));
        cudaEventRecord(c.ev1, c.stream);
        CU(c.sync());
        float ms = 0; cudaEventElapsedTime(&ms, c.ev0, c.ev1); c.ms[3] += ms; c.launches[3] += 1;
        for (int i = 0; i < n; i++) states[i] = ms_[i];
        for (int i = 0; i < n; i++) if (states[i].status != 0) return fail(h, GOICP_ERR_UNSUPPORTED, "ICP with trimming supports Nd <= 2048");
        return GOICP_OK;
    }
    CU(c.dIcp.ensure(sizeof(IcpState) * n));
    CU(c.hIcp.ensure(sizeof(IcpState) * n));
    IcpState* hs = c.hIcp.as<IcpState>();
    for (int i = 0; i < n; i++) hs[i] = states[i];
    cudaEventRecord(c.ev0, c.stream);
    int nl = 0;
    CU(cudaMemcpyAsync(c.dIcp.p, hs, sizeof(IcpState) * n, cudaMemcpyHostToDevice, c.stream));
    {
        CU(goicp_launch_icp_begin(h->dPairs.as<PairDev>(), c.dIcp.as<IcpState>(), n, c.stream)); nl++;
        if (anyIcp) {
            int burst = 4;
            for (int it = 0; it < 10000;) {
                for (int b = 0; b < burst; b++) { CU(goicp_launch_icp_iter(h->dPairs.as<PairDev>(), c.dIcp.as<IcpState>(), n, maxNd, maxNm, h->numSM, c.stream)); nl += 2; }
                it += burst;
                CU(cudaMemcpyAsync(hs, c.dIcp.p, sizeof(IcpState) * n, cudaMemcpyDeviceToHost, c.stream));
                CU(c.sync());
                bool all = true;
                for (int i = 0; i < n; i++) if (hs[i].mode == 0 && !hs[i].done) all = false;
                if (all) break;
                if (burst < 16) burst *= 2;
            }
        }
        CU(goicp_launch_icp_score(h->dPairs.as<PairDev>(), c.dIcp.as<IcpState>(), n, c.stream)); nl++;
    }
    cudaEventRecord(c.ev1, c.stream);
    CU(cudaMemcpyAsync(hs, c.dIcp.p, sizeof(IcpState) * n, cudaMemcpyDeviceToHost, c.stream));
    CU(c.sync());
    float ms = 0; cudaEventElapsedTime(&ms, c.ev0, c.ev1); c.ms[3] += ms; c.launches[3] += nl;
    for (int i = 0; i < n; i++) states[i] = hs[i];
    for (int i = 0; i < n; i++) if (states[i].status != 0) return fail(h, GOICP_ERR_UNSUPPORTED, "ICP with trimming supports Nd <= 2048");
    return GOICP_OK;
}

static IcpState make_icp_state(int pair, int mode, const double* R, const double* t) {
    IcpState s; memset(&s, 0, sizeof s);
    s.pair = pair; s.mode = mode;
    for (int k = 0; k < 9; k++) s.R[k] = R ? R[k] : (k % 4 == 0 ? 1.0 : 0.0);
    for (int k = 0; k < 3; k++) s.t[k] = t ? t[k] : 0.0;
    s.err = -1.f;
    return s;
}

// ---- rotation of a child cube (jly_goicp.cpp:716-747): false if the cube lies outside the pi-ball --------------------
static bool child_rotation(const RNode& nr, float* R) {
    float v1 = nr.a + nr.w / 2, v2 = nr.b + nr.w / 2, v3 = nr.c + nr.w / 2;
    if ((double)sqrtf(v1 * v1 + v2 * v2 + v3 * v3) - GOICP_SQRT3 * nr.w / 2 > GOICP_PI) return false;   // :723
    float t = sqrtf(v1 * v1 + v2 * v2 + v3 * v3);                                                    // :729
    if (t > 0) {
        v1 /= t; v2 /= t; v3 /= t;
        float ct = cosf(t), ct2 = 1 - ct, st = sinf(t);
        float tmp121 = v1 * v2 * ct2, tmp122 = v3 * st, tmp131 = v1 * v3 * ct2, tmp132 = v2 * st, tmp231 = v2 * v3 * ct2, tmp232 = v1 * st;
        R[0] = ct + v1 * v1 * ct2; R[1] = tmp121 - tmp122; R[2] = tmp131 + tmp132;
        R[3] = tmp121 + tmp122; R[4] = ct + v2 * v2 * ct2; R[5] = tmp231 - tmp232;
        R[6] = tmp131 - tmp132; R[7] = tmp231 + tmp232; R[8] = ct + v3 * v3 * ct2;
    } else {
        for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.f : 0.f;   // :759-762 copies the cloud unrotated
    }
    return true;
}
static inline RNode child_of(const RNode& par, int j) {
    RNode nr{}; nr.w = par.w / 2; nr.l = par.l + 1;
    nr.a = par.a + (j & 1) * nr.w; nr.b = par.b + ((j >> 1) & 1) * nr.w; nr.c = par.c + ((j >> 2) & 1) * nr.w;   // :710-712
    return nr;
}
static inline unsigned long long call_key(int nodeId, int j, int kind) { return ((unsigned long long)(unsigned)nodeId << 4) | (unsigned)(j << 1) | (unsigned)kind; }

struct ReqTag { int prob; unsigned long long key; float entryOpt; bool both; };

// Advance one problem's OuterBnB as far as cached results allow; on return P.phase tells what it waits for.
static void advance(Eng* h, int pi) {
    Problem& P = h->probs[pi];
    const goicp_params& p = h->params;
    const float SSE = P.dev.SSEThresh;
    for (;;) {
        switch (P.phase) {
        case PH_START: case PH_WAIT_INIT: case PH_WAIT_ICP: case PH_DONE: return;
        case PH_POP: {
            if (P.q.empty()) { tracef(P.trace, "Rotation Queue Empty\nError*: %g, LB: %g\n", P.optError, P.lastLb); P.phase = PH_DONE; return; }   // :670-677
            P.par = rheap_pop(P.q); P.cnt[3]++;
            if ((P.optError - P.par.lb) <= SSE) {                                                      // :685
                tracef(P.trace, "Threshold reached\nError*: %g, LB: %g, epsilon: %g\n", P.optError, P.par.lb, SSE);
                P.phase = PH_DONE; return;
            }
            P.j = 0; P.phase = PH_CHILD_UB;
            break;
        }
        case PH_CHILD_UB: {
            if (P.j >= 8) { P.phase = PH_POP; break; }
            P.child = child_of(P.par, P.j);
            if (!child_rotation(P.child, P.R)) { P.j++; break; }
            auto it = P.cache.find(call_key(P.par.id, P.j, 0));
            if (it == P.cache.end() || it->second.entryOpt != P.optError) return;   // blocked
            const CallRes r = it->second; P.cache.erase(it);
            P.cnt[4]++; P.cnt[0]++; P.cnt[1] += r.pops; P.cnt[2] += r.subcubes;
            P.ubChild = r.err;
            if (r.err < P.optError) {   // :771-790
                P.optError = r.err;
                for (int k = 0; k < 9; k++) P.optR[k] = P.R[k];
                P.optT[0] = r.tn[0] + r.tn[3] / 2; P.optT[1] = r.tn[1] + r.tn[3] / 2; P.optT[2] = r.tn[2] + r.tn[3] / 2;   // float expr -> double
                P.cache.clear(); P.inflight.clear(); P.quiet = 0;
                P.phase = PH_WAIT_ICP; P.icpPending = true;
                return;
            }
            P.phase = PH_CHILD_LB;
            break;
        }
        case PH_CHILD_LB: {
            auto it = P.cache.find(call_key(P.par.id, P.j, 1));
            if (it == P.cache.end() || it->second.entryOpt != P.optError) return;   // blocked
            const CallRes r = it->second; P.cache.erase(it);
            P.cnt[0]++; P.cnt[1] += r.pops; P.cnt[2] += r.subcubes;
            P.lastLb = r.err;
            if (!(r.err >= P.optError)) {   // :863-871
                RNode nr = P.child; nr.ub = P.ubChild; nr.lb = r.err; nr.id = P.nextId++;
                rheap_push(P.q, nr);
            }
            P.j++; P.phase = PH_CHILD_UB;
            break;
        }
        }
    }
    (void)p;
}

// result of the post-improvement ICP (jly_goicp.cpp:791-854)
static void finish_improvement(Eng* h, int pi) {
    Problem& P = h->probs[pi];
    P.optComp = P.compatPose;                                                                          // :791
    tracef(P.trace, "Error*: %g (BNB)\n", P.optError);
    P.cnt[5]++;
    if (P.icpErr < P.optError) {                                                                       // :813-840
        P.optError = P.icpErr;
        memcpy(P.optR, P.icpR, sizeof P.optR); memcpy(P.optT, P.icpT, sizeof P.optT);
        P.optComp = P.icpIncomp;
        tracef(P.trace, "Error*: %g (ICP)\n", P.icpErr);
    }
    std::vector<RNode> qn;                                                                             // :843-853
    while (!P.q.empty()) { RNode n = rheap_pop(P.q); if (n.lb < P.optError) rheap_push(qn, n); else break; }
    P.q.swap(qn);
    P.cache.clear(); P.inflight.clear();
    P.phase = PH_CHILD_LB;
}

// Requests of one blocked problem: the blocking call first, then speculation in the reference's expected order.
static void gather_requests(Eng* h, int pi, std::vector<InnerProb>& reqs, std::vector<ReqTag>& tags, bool merge = false) {
    Problem& P = h->probs[pi];
    const float SSE = P.dev.SSEThresh;
    std::unordered_set<unsigned long long> seen;
    auto want = [&](const RNode& par, int j, int kind, const RNode& ch, const float* R) {
        const unsigned long long key = call_key(par.id, j, kind);
        auto it = P.cache.find(key);
        if (it != P.cache.end() && it->second.entryOpt == P.optError) return;
        if (P.inflight.count(key)) return;
        if (!seen.insert(key).second) return;
        // Q2: the reference indexes maxRotDis[level] without a bound check (undefined beyond level 19); we clamp.
        const int lbLevel = std::min(ch.l, GOICP_MAXROTLEVEL - 1);
        // resident-kernel scheduler: the lower-bound call of a cube rides on the request of its upper-bound call (one CTA runs
        // both, sharing the staged cloud and the corner memo; the second is skipped if the first improves the incumbent)
        if (merge && kind == 1 && !tags.empty() && tags.back().prob == pi && tags.back().key == call_key(par.id, j, 0) && !tags.back().both) {
            reqs.back().level = GOICP_REQ_BOTH + lbLevel; tags.back().both = true;
            return;
        }
        InnerProb ip; ip.pair = pi; ip.level = kind ? lbLevel : -1; ip.optError = P.optError;
        memcpy(ip.R, R, sizeof ip.R);
        reqs.push_back(ip); tags.push_back(ReqTag{pi, key, P.optError, false});
    };
    // a call that is cached under the current incumbent or already in flight needs no request (and no rotation matrix)
    auto known = [&](const RNode& par, int j, int kind) {
        const unsigned long long key = call_key(par.id, j, kind);
        auto it = P.cache.find(key);
        return (it != P.cache.end() && it->second.entryOpt == P.optError) || P.inflight.count(key) != 0;
    };
    // current parent, from the blocking call on
    for (int j = P.j; j < 8; j++) {
        const bool skipUb = j == P.j && P.phase == PH_CHILD_LB;
        if ((skipUb || known(P.par, j, 0)) && known(P.par, j, 1)) continue;
        RNode ch = child_of(P.par, j); float R[9];
        if (!child_rotation(ch, R)) continue;
        if (!skipUb) want(P.par, j, 0, ch, R);
        want(P.par, j, 1, ch, R);
    }
    // the next queue nodes in pop order; width grows while the incumbent stays unchanged
    // inside a batch the pairs themselves fill the GPU; in its tail (fewer requests in flight than the resident kernel has
    // CTAs) the remaining deep pairs speculate as widely as a single registration does
    const bool tail = h->tail_spec && h->residentCtas > 0 && h->outstanding.load(std::memory_order_relaxed) < h->tail_thr * h->residentCtas;
    const int specw = (h->probs.size() > 1 && !tail) ? std::min(h->spec_width, h->batch_spec_width) : (tail ? h->tail_spec_mult * h->spec_width : h->spec_width);
    int width = std::min(specw, P.quiet < 30 ? (1 << std::min(P.quiet, 20)) - 1 : specw);
    if (width > 0 && !P.q.empty()) {
        // the `width` best nodes of the rotation queue: P.q is a binary heap, so they are reached from the root through a
        // frontier of candidate positions (no copy, no sort of the whole queue)
        const int n = (int)P.q.size(), k = std::min(width, n);
        int cand[80]; int nc = 0; cand[nc++] = 0;
        for (int i = 0; i < k && nc > 0; i++) {
            int b = 0;
            for (int c = 1; c < nc; c++) if (rnode_less(P.q[cand[b]], P.q[cand[c]])) b = c;
            const int pos = cand[b]; cand[b] = cand[--nc];
            if (2 * pos + 1 < n && nc < 78) cand[nc++] = 2 * pos + 1;
            if (2 * pos + 2 < n && nc < 78) cand[nc++] = 2 * pos + 2;
            const RNode& nd = P.q[pos];
            if ((P.optError - nd.lb) <= SSE) break;
            for (int j = 0; j < 8; j++) {
                if (known(nd, j, 0) && known(nd, j, 1)) continue;
                RNode ch = child_of(nd, j); float R[9];
                if (!child_rotation(ch, R)) continue;
                want(nd, j, 0, ch, R); want(nd, j, 1, ch, R);
            }
        }
    }
    P.quiet++;
}

// results of an ICP / scoring request -> the problem's exchange fields
static void absorb_icp(Problem& P, const IcpState& st) {
    if (st.mode == 1) P.initErr = st.error;
    else if (st.mode == 2) P.compatPose = st.compat_pose;
    else { P.icpErr = st.error; memcpy(P.icpR, st.R, sizeof P.icpR); memcpy(P.icpT, st.t, sizeof P.icpT); P.icpIncomp = st.incomp; }
}
// continue a problem whose ICP results have arrived: start of OuterBnB (:601-664) or post-improvement (:791-854)
static void after_icp(Eng* h, int i) {
    const goicp_params& p = h->params;
    Problem& P = h->probs[i];
    if (P.phase == PH_WAIT_INIT) {
        float optError = P.initErr;
        if (p.regularization > 0) optError += p.regularization * (P.Nd * P.Nd);                       // :623
        if (p.regularizationFPFH > 0) optError += p.regularizationFPFH * (100 * 8 * 100 * 8);          // :624
        if (p.regularizationNeighbors > 0) optError += p.regularizationNeighbors * (P.Nd * 6 * P.Nd * 6);
        P.optError = optError;
        tracef(P.trace, "Error*: %g (Init)\n", P.optError);
        P.cnt[5]++;
        if (P.icpErr < P.optError) {                                                                   // :636-661
            P.optError = P.icpErr; memcpy(P.optR, P.icpR, sizeof P.optR); memcpy(P.optT, P.icpT, sizeof P.optT);
            P.optComp = P.icpIncomp;
            tracef(P.trace, "Error*: %g (ICP)\n", P.icpErr);
        }
        RNode root{}; root.a = p.rotMinX; root.b = p.rotMinY; root.c = p.rotMinZ; root.w = p.rotWidth; root.l = 0; root.lb = 0; root.id = 0;
        rheap_push(P.q, root);
        P.phase = PH_POP;
    } else if (P.phase == PH_WAIT_ICP && P.icpPending) {
        P.icpPending = false;
        finish_improvement(h, i);
    }
}

// Start-of-search state of one problem (GoICP::Initialize :240-241 resets optR/optT)
static void reset_search(Problem& P) {
    P.phase = PH_START; P.q.clear(); P.cache.clear(); P.trace.clear(); memset(P.cnt, 0, sizeof P.cnt);
    P.nextId = 1; P.quiet = 0; P.optComp = 0; P.lastLb = 0; P.status = 0; P.icpPending = false; P.icpQueued = false;
    P.pend.clear(); P.inflight.clear();
    for (int k = 0; k < 9; k++) P.optR[k] = (k % 4 == 0);
    P.optT[0] = P.optT[1] = P.optT[2] = 0;
}

// One stream of lock-step waves over up to `slots` problems at a time; finished problems are replaced from the shared
// counter `next` (so a deep pair never stalls more than its own stream).  GoICP::OuterBnB (jly_goicp.cpp:582) per problem.
static goicp_status register_group(Eng* h, WaveCtx& c, const BnbCfg& cfg, std::atomic<int>& next, int slots, const std::vector<int>* subset = nullptr) {
    const goicp_params& p = h->params;
    const int np = subset ? (int)subset->size() : (int)h->probs.size();
    std::vector<int> active;
    std::vector<InnerProb> reqs; std::vector<ReqTag> tags; std::vector<InnerOut> outs; std::vector<IcpState> icps; std::vector<int> icpOwner;
    goicp_status s;
    auto tl = clk::now();
    for (;;) {
        while ((int)active.size() < slots) { const int k = next.fetch_add(1); if (k >= np) break; const int i = subset ? (*subset)[k] : k; reset_search(h->probs[i]); active.push_back(i); }
        if (active.empty()) break;
        reqs.clear(); tags.clear(); icps.clear(); icpOwner.clear();
        for (int i : active) {
            Problem& P = h->probs[i];
            if (P.phase == PH_START) {   // initial error (:601-627) and ICP from the identity (:634)
                icps.push_back(make_icp_state(i, 1, nullptr, nullptr)); icpOwner.push_back(i);
                icps.push_back(make_icp_state(i, 0, P.optR, P.optT)); icpOwner.push_back(i);
                P.phase = PH_WAIT_INIT; continue;
            }
            advance(h, i);
            if (P.phase == PH_DONE) continue;
            if (P.phase == PH_WAIT_ICP) {   // updateCompatibilities (:791) + ICP(R,t) (:810) at the new incumbent
                icps.push_back(make_icp_state(i, 2, P.optR, P.optT)); icpOwner.push_back(i);
                icps.push_back(make_icp_state(i, 0, P.optR, P.optT)); icpOwner.push_back(i);
            } else gather_requests(h, i, reqs, tags);
        }
        active.erase(std::remove_if(active.begin(), active.end(), [&](int i) { return h->probs[i].phase == PH_DONE; }), active.end());
        if (reqs.empty() && icps.empty()) continue;
        c.waves++;
        c.tLogic += secs_since(tl);
        if ((s = run_inner(h, c, cfg, reqs, outs))) return s;
        { auto ti = clk::now(); if ((s = run_icp(h, c, icps))) return s; c.tIcp += secs_since(ti); }
        tl = clk::now();
        for (size_t k = 0; k < tags.size(); k++) {
            Problem& P = h->probs[tags[k].prob];
            CallRes r; r.entryOpt = tags[k].entryOpt; r.err = outs[k].err; memcpy(r.tn, outs[k].node, sizeof r.tn); r.pops = outs[k].pops; r.subcubes = outs[k].subcubes;
            P.cache[tags[k].key] = r;
        }
        for (size_t k = 0; k < icps.size(); k++) absorb_icp(h->probs[icpOwner[k]], icps[k]);
        for (int i : active) after_icp(h, i);
    }
    return GOICP_OK;
}


// ---- persistent-queue scheduler (batches) -----------------------------------------------------------------------------
// One resident inner_bnb_kernel<.., PERSIST> serves a request ring in mapped host memory for the whole batch; host worker
// threads advance their pairs INDEPENDENTLY (no lock-step): a pair publishes the InnerBnB calls it needs (+ speculation),
// keeps going as soon as the call its OuterBnB order waits for has completed, and sends ICP requests through its thread's
// side stream.  A deep pair therefore delays nobody else, and the GPU always holds a mix of calls of hundreds of pairs.
struct PQ {
    QueueCell* cells; InnerOut* outs; unsigned cellMask, cellShift;
    volatile unsigned* doneRing = nullptr; unsigned doneCap = 0;   // completion hints, one ring per worker (QueueDev)
    std::atomic<unsigned> reserve{0};
    // request words first, the two lap tags last (x86 stores are observed in program order; a 32-byte half of the cell that
    // shows its tag therefore shows its request words too)
    void publish(unsigned slot, const InnerProb* pr) {
        const unsigned idx = reserve.fetch_add(1);
        QueueCell* c = cells + (idx & cellMask);
        const unsigned tag = (idx >> cellShift) + 1u;
        c->slot = slot;
        if (pr) c->pr = *pr;
        std::atomic_thread_fence(std::memory_order_release);
        *reinterpret_cast<volatile unsigned*>(&c->seqA) = tag;
        *reinterpret_cast<volatile unsigned*>(&c->seqB) = tag;
    }
};
static inline bool out_ready(const InnerOut& o) {
    return *reinterpret_cast<const volatile unsigned*>(&o.seq0) == 1u && *reinterpret_cast<const volatile unsigned*>(&o.seq1) == 1u;
}
static inline void out_arm(InnerOut& o) { *reinterpret_cast<volatile unsigned*>(&o.seq0) = 0u; *reinterpret_cast<volatile unsigned*>(&o.seq1) = 0u; }

static goicp_status persistent_worker(Eng* h, WaveCtx& c, PQ& pq, std::atomic<int>& next, int slots, int slotLo, int slotHi, const BnbCfg& cfg, int worker) {
    const int np = (int)h->probs.size();
    std::vector<int> slotPair(slotHi - slotLo, -1);   // pair that published the request in each of this worker's result slots
    std::vector<int> dirtyList, todo;                 // pairs with news: a completed request, just admitted, or waiting for a free slot
    volatile unsigned* ring = pq.doneRing + (size_t)worker * pq.doneCap; unsigned ringHead = 0;
    auto mark = [&](int i) { Problem& Q = h->probs[i]; if (!Q.dirty) { Q.dirty = true; dirtyList.push_back(i); } };
    std::vector<int> freeSlots; freeSlots.reserve(slotHi - slotLo);
    for (int sidx = slotHi - 1; sidx >= slotLo; --sidx) freeSlots.push_back(sidx);
    std::vector<int> active;
    std::vector<Problem::PendReq> zombies;
    std::vector<InnerProb> reqs; std::vector<ReqTag> tags;
    IcpState* icpHost = reinterpret_cast<IcpState*>(h->qIcp.h);
    IcpState* icpDev = reinterpret_cast<IcpState*>(h->qIcp.d);
    Problem dummy;
    auto lastProgress = clk::now();
    const auto tStart = clk::now(); double nextSample = 0; const bool dbgTimeline = getenv("GOICP_TIMELINE") != nullptr;
    auto harvest = [&](Problem& P, std::vector<Problem::PendReq>& pend, bool live) {
        bool any = false;
        for (size_t k = 0; k < pend.size();) {
            const InnerOut& o = pq.outs[pend[k].slot];
            if (!out_ready(o)) { ++k; continue; }
            std::atomic_thread_fence(std::memory_order_acquire);
            if (o.status == 4 && live) P.status = GOICP_ERR_OVERFLOW;   // queue outgrew the CTA's slab: the pair is re-run by the wave scheduler
            if (live && pend[k].entryOpt == P.optError) {
                CallRes r; r.entryOpt = pend[k].entryOpt; r.err = o.err; memcpy(r.tn, (const void*)o.node, sizeof r.tn); r.pops = o.pops; r.subcubes = o.subcubes;
                P.cache[pend[k].key] = r;
                P.inflight.erase(pend[k].key);
                if (pend[k].both) {
                    if (o.ran2) { CallRes r2; r2.entryOpt = pend[k].entryOpt; r2.err = o.err2; memset(r2.tn, 0, sizeof r2.tn); r2.pops = o.pops2; r2.subcubes = o.subcubes2; P.cache[pend[k].key | 1ull] = r2; }
                    P.inflight.erase(pend[k].key | 1ull);
                }
            }
            freeSlots.push_back(pend[k].slot);
            pend[k] = pend.back(); pend.pop_back();
            h->outstanding.fetch_sub(1, std::memory_order_relaxed);
            any = true;
        }
        return any;
    };
    // GoICP::ICP requests of pair i: (initial error | updateCompatibilities) + ICP, two CTAs of the resident kernel
    auto send_icp = [&](int i) -> bool {
        Problem& P = h->probs[i];
        if (freeSlots.size() < 2) return false;
        const bool init = P.phase == PH_WAIT_INIT;
        icpHost[2 * i] = make_icp_state(i, init ? 1 : 2, init ? nullptr : P.optR, init ? nullptr : P.optT);
        icpHost[2 * i + 1] = make_icp_state(i, 0, P.optR, P.optT);
        for (int k = 0; k < 2; k++) {
            const int slot = freeSlots.back(); freeSlots.pop_back();
            InnerProb ip; memset(&ip, 0, sizeof ip);
            ip.pair = i; ip.level = GOICP_REQ_ICP; ip.optError = 0.f;
            const unsigned long long ptr = (unsigned long long)(uintptr_t)(icpDev + 2 * i + k);
            const unsigned lo = (unsigned)(ptr & 0xFFFFFFFFull), hi = (unsigned)(ptr >> 32);
            memcpy(&ip.R[0], &lo, 4); memcpy(&ip.R[1], &hi, 4);
            out_arm(pq.outs[slot]);
            P.icpSlot[k] = slot; slotPair[slot - slotLo] = i;
            pq.publish((unsigned)slot, &ip);
        }
        P.icpQueued = true;
        c.launches[3] += 2;
        return true;
    };
    for (;;) {
        while ((int)active.size() < slots) { const int i = next.fetch_add(1); if (i >= np) break; reset_search(h->probs[i]); h->probs[i].dirty = false; active.push_back(i); h->activePairs.fetch_add(1); mark(i); }
        if (active.empty() && zombies.empty()) break;
        bool progressed = false;
        c.callsUsed++;   // loop iterations
        if (slotLo == 0 && dbgTimeline) { const double tnow = secs_since(tStart); if (tnow >= nextSample) { fprintf(stderr, "[timeline] t=%.3f active_pairs=%d outstanding=%d\n", tnow, h->activePairs.load(), h->outstanding.load()); nextSample += 0.05; } }
        auto tIter = clk::now();
        // ---- completion hints of the resident kernel: which pairs have news ----
        for (unsigned v; (v = ring[ringHead & (pq.doneCap - 1u)]) != 0u; ringHead++) {
            ring[ringHead & (pq.doneCap - 1u)] = 0u;
            // the hint may overtake its record on the bus (two stores, no fence): wait for the record's own flags
            for (long spin = 0; !out_ready(pq.outs[v - 1u]) && spin < 100000000L; spin++) __builtin_ia32_pause();
            const int owner = slotPair[(int)(v - 1u) - slotLo];
            if (owner >= 0 && h->probs[owner].phase != PH_DONE) mark(owner);
            progressed = true;
        }
        // ---- every pair with news: harvest finished calls, advance, publish what it needs next ----
        todo.swap(dirtyList); dirtyList.clear();
        for (int i : todo) h->probs[i].dirty = false;
        for (int i : todo) {
            Problem& P = h->probs[i];
            if (P.phase == PH_DONE) continue;
            if (harvest(P, P.pend, true)) progressed = true;
            if (P.status == GOICP_ERR_OVERFLOW) {   // abandon the pair here; register_persistent re-runs it with growing queues
                for (auto& r : P.pend) zombies.push_back(r);
                P.pend.clear(); P.inflight.clear(); P.cache.clear();
                if (P.icpQueued) { zombies.push_back(Problem::PendReq{P.icpSlot[0], 0ull, 0.f, false}); zombies.push_back(Problem::PendReq{P.icpSlot[1], 0ull, 0.f, false}); P.icpQueued = false; }
                P.phase = PH_DONE; h->activePairs.fetch_sub(1); progressed = true;
                continue;
            }
            if (P.phase == PH_START) { P.phase = PH_WAIT_INIT; P.icpQueued = false; }
            if (P.phase == PH_WAIT_INIT || P.phase == PH_WAIT_ICP) {
                if (!P.icpQueued) { if (send_icp(i)) progressed = true; else mark(i); continue; }
                const bool d0 = out_ready(pq.outs[P.icpSlot[0]]), d1 = out_ready(pq.outs[P.icpSlot[1]]);
                if (!(d0 && d1)) continue;
                std::atomic_thread_fence(std::memory_order_acquire);
                freeSlots.push_back(P.icpSlot[0]); freeSlots.push_back(P.icpSlot[1]);
                if (icpHost[2 * i].status != 0 || icpHost[2 * i + 1].status != 0) return fail(h, GOICP_ERR_UNSUPPORTED, "ICP with trimming supports Nd <= 2048");
                absorb_icp(P, icpHost[2 * i]); absorb_icp(P, icpHost[2 * i + 1]);
                P.icpQueued = false;
                after_icp(h, i);
                progressed = true;
            }
            const Phase before = P.phase; const int jb = P.j; const int idb = P.par.id;
            advance(h, i);
            if (P.phase != before || P.j != jb || P.par.id != idb) progressed = true;
            if (P.phase == PH_DONE) { for (auto& r : P.pend) zombies.push_back(r); P.pend.clear(); P.inflight.clear(); h->activePairs.fetch_sub(1); continue; }
            if (P.phase == PH_WAIT_ICP) { P.icpQueued = false; if (send_icp(i)) progressed = true; else mark(i); continue; }
            // blocked on an InnerBnB result: is it already on its way?
            const unsigned long long need = call_key(P.par.id, P.j, P.phase == PH_CHILD_LB ? 1 : 0);
            if (P.inflight.count(need)) continue;
            reqs.clear(); tags.clear();
            auto tg = clk::now();
            gather_requests(h, i, reqs, tags, h->merge_calls);
            c.tLogic += secs_since(tg); tg = clk::now();
            for (size_t k = 0; k < reqs.size(); k++) {
                if (freeSlots.empty() || h->outstanding.load(std::memory_order_relaxed) > (int)(pq.cellMask >> 1)) break;   // out of slots / ring half full: the rest is regathered later
                const int slot = freeSlots.back(); freeSlots.pop_back();
                out_arm(pq.outs[slot]); slotPair[slot - slotLo] = i;
                P.pend.push_back(Problem::PendReq{slot, tags[k].key, tags[k].entryOpt, tags[k].both});
                P.inflight.insert(tags[k].key);
                if (tags[k].both) { P.inflight.insert(tags[k].key | 1ull); c.callsLaunched++; }
                pq.publish((unsigned)slot, &reqs[k]);
                h->outstanding.fetch_add(1, std::memory_order_relaxed);
                c.callsLaunched++;
            }
            c.tInnerEnq += secs_since(tg);
            if (!reqs.empty()) { progressed = true; c.waves++; }
            if (!P.inflight.count(need)) mark(i);   // out of slots: the blocking call itself is still unpublished
        }
        active.erase(std::remove_if(active.begin(), active.end(), [&](int i) { return h->probs[i].phase == PH_DONE; }), active.end());
        // ---- results of calls whose pair has already finished (speculation): just recycle the slots ----
        if (!zombies.empty()) { if (harvest(dummy, zombies, false)) progressed = true; }
        c.tIcp += secs_since(tIter);   // busy time of the loop body
        if (progressed) lastProgress = clk::now();
        else {
            auto ts0 = clk::now();
            static const bool dbg = getenv("GOICP_DEBUG") != nullptr;
            if (dbg && secs_since(lastProgress) > 3.0) {
                std::string m;
                for (int i : active) { Problem& P = h->probs[i]; char b[160]; snprintf(b, sizeof b, " [pair %d ph %d j %d pend %zu infl %zu icpq %d]", i, (int)P.phase, P.j, P.pend.size(), P.inflight.size(), (int)P.icpQueued); m += b; }
                fprintf(stderr, "main stream: %s; worker slots %d..%d: active %zu zombies %zu reserve %u free %zu%s\n", cudaGetErrorString(cudaStreamQuery(h->stream)), slotLo, slotHi, active.size(), zombies.size(), pq.reserve.load(), freeSlots.size(), m.c_str());
                return fail(h, GOICP_ERR_CUDA, "persistent scheduler: debug stop");
            }
            if (secs_since(lastProgress) > 45.0) return fail(h, GOICP_ERR_CUDA, "persistent scheduler: no progress for 45 s (device stalled?)");
            bool news = false;   // spin briefly on the completion ring before giving the core away
            for (int spin = 0; spin < 200 && !news; spin++) { news = ring[ringHead & (pq.doneCap - 1u)] != 0u; if (!news) __builtin_ia32_pause(); }
            if (!news) { struct timespec ts = {0, 5000}; nanosleep(&ts, nullptr); }
            c.tInnerWait += secs_since(ts0);
        }
    }
    (void)cfg; (void)c;
    return GOICP_OK;
}

static goicp_status register_persistent(Eng* h, const BnbCfg& cfg, int groups, int slots) {
    const int NSLOT = 1 << 17, ORDER = 1 << 18;   // result slots; ring cells (> requests ever in flight)
    const int np = (int)h->probs.size();
    // everything is allocated BEFORE the resident kernel starts (cudaMalloc / cudaFree would wait for it forever)
    CU(h->qOuts.ensure(sizeof(InnerOut) * (size_t)NSLOT));
    CU(h->qOrder.ensure(sizeof(QueueCell) * (size_t)ORDER));
    CU(h->qClaim.ensure(sizeof(unsigned) * 64));   // [0] ring claim counter, [1 + w] completion-ring tail of worker w
    CU(h->qIcp.ensure(sizeof(IcpState) * 2 * (size_t)np));
    CU(h->qDone.ensure(sizeof(unsigned) * (size_t)2 * NSLOT));   // per-worker completion rings: groups x (power of two >= NSLOT / groups)
    int perSM = goicp_inner_bnb_persistent_occupancy(cfg.smemBytes, h->exact_sums, cfg.threads, cfg.useSmem, cfg.ct);
    { const char* e = getenv("GOICP_CTAS_PER_SM"); if (e && atoi(e) >= 1) perSM = std::min(perSM, atoi(e)); }
    const int ctas = h->numSM * perSM;   // the resident kernel owns the GPU for the batch: InnerBnB and ICP requests both run on its CTAs
    h->residentCtas = ctas; { const char* e = getenv("GOICP_TAIL_SPEC"); if (e) h->tail_spec = atoi(e) != 0; }
    int heapCap = 1 << 15;
    { const char* e = getenv("GOICP_HEAPCAP"); if (e && atoi(e) >= 129) heapCap = atoi(e); }   // test hook: force the overflow pool
    CU(h->qHeaps.ensure(sizeof(HeapEnt) * (size_t)ctas * heapCap));
    if (!cfg.useSmem) CU(h->qScratch.ensure(sizeof(float) * cfg.smemFloats * (size_t)ctas));
    const int memoCap = 8192;
    if (h->qMemo.cap < (size_t)32 * memoCap * ctas) { CU(h->qMemo.ensure((size_t)32 * memoCap * ctas)); CU(cudaMemsetAsync(h->qMemo.p, 0, h->qMemo.cap, h->stream)); }
    while ((int)h->workers.size() < groups) {
        std::unique_ptr<WaveCtx> w(new WaveCtx());
        if (w->init(true, nullptr) != GOICP_OK) return fail(h, GOICP_ERR_CUDA, "worker stream creation failed");
        h->workers.push_back(std::move(w));
    }
    for (int g = 0; g < groups; g++) {
        WaveCtx* w = h->workers[g].get();
        memset(w->ms, 0, sizeof w->ms); memset(w->launches, 0, sizeof w->launches); w->waves = w->callsLaunched = 0;
        w->tLogic = w->tInnerEnq = w->tInnerWait = w->tIcp = 0; w->callsUsed = 0;
    }
    h->main.callsUsed = 0;
    memset(h->qOrder.h, 0, sizeof(QueueCell) * (size_t)ORDER);
    CU(cudaMemsetAsync(h->qClaim.p, 0, sizeof(unsigned) * 64, h->stream));
    const int per = NSLOT / groups;
    unsigned doneCap = 1; while ((int)doneCap < per) doneCap <<= 1;
    memset(h->qDone.h, 0, sizeof(unsigned) * (size_t)doneCap * groups);
    PQ pq; pq.doneRing = reinterpret_cast<volatile unsigned*>(h->qDone.h); pq.doneCap = doneCap;
    pq.cells = reinterpret_cast<QueueCell*>(h->qOrder.h); pq.outs = reinterpret_cast<InnerOut*>(h->qOuts.h);
    pq.cellMask = ORDER - 1; pq.cellShift = 18;
    QueueDev qd; qd.cells = reinterpret_cast<const QueueCell*>(h->qOrder.d); qd.outs = reinterpret_cast<InnerOut*>(h->qOuts.d);
    qd.cellMask = ORDER - 1; qd.cellShift = 18; qd.claim = h->qClaim.as<unsigned>();
    qd.doneTail = h->qClaim.as<unsigned>() + 1; qd.doneRing = reinterpret_cast<unsigned*>(h->qDone.d); qd.doneCap = doneCap; qd.slotsPerWorker = (unsigned)per;
    cudaEventRecord(h->main.ev0, h->stream);
    CU(goicp_launch_inner_bnb_persistent(h->dPairs.as<PairDev>(), qd, h->qHeaps.as<HeapEnt>(), heapCap, ctas, h->qScratch.as<float>(), cfg.smemFloats,
                                         cfg.NdP, cfg.NdQ, cfg.smemBytes, cfg.useSmem, cfg.gridOff, cfg.S3p, h->exact_sums, cfg.ct, cfg.threads, h->qMemo.p, memoCap, h->dGen.as<unsigned>(), h->stream));
    cudaEventRecord(h->main.ev1, h->stream);
    g_no_device_alloc.store(true);
    std::atomic<int> next(0);
    h->activePairs.store(0); h->outstanding.store(0); h->stats[14] = 0;
    std::vector<goicp_status> st(groups, GOICP_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < groups; g++) {
        WaveCtx* w = h->workers[g].get();
        th.emplace_back([h, w, &pq, &next, &st, g, slots, per, &cfg]() { cudaSetDevice(h->device); st[g] = persistent_worker(h, *w, pq, next, slots, g * per, (g + 1) * per, cfg, g); });
    }
    for (auto& t : th) t.join();
    for (int k = 0; k < ctas; k++) pq.publish(0xFFFFFFFFu, nullptr);   // one shut-down marker per CTA
    cudaError_t e = cudaStreamSynchronize(h->stream);
    g_no_device_alloc.store(false); h->residentCtas = 0;
    if (e != cudaSuccess) return fail(h, GOICP_ERR_CUDA, "resident inner_bnb kernel: %s", cudaGetErrorString(e));
    float ms = 0; cudaEventElapsedTime(&ms, h->main.ev0, h->main.ev1); h->main.ms[2] += ms; h->main.launches[2] += 1;
    for (int g = 0; g < groups; g++) if (st[g]) return st[g];
    {   // pairs whose translation queue outgrew the resident kernel's per-CTA slab: re-run them with the wave scheduler,
        // which grows the queue slabs on demand (the resident kernel cannot: no device allocation while it runs)
        std::vector<int> redo;
        for (int i = 0; i < np; i++) if (h->probs[i].status == GOICP_ERR_OVERFLOW) redo.push_back(i);
        if (!redo.empty()) {
            std::atomic<int> nx(0);
            h->main.ctaCap = 0;
            goicp_status s2 = register_group(h, h->main, cfg, nx, std::min<int>(64, (int)redo.size()), &redo);
            if (s2) return s2;
            h->stats[14] = (double)redo.size();
        }
    }
    for (int g = 0; g < groups; g++) {
        WaveCtx* w = h->workers[g].get();
        h->main.ms[3] += w->ms[3]; h->main.launches[3] += w->launches[3];
        h->main.waves += w->waves; h->main.callsLaunched += w->callsLaunched; h->main.callsUsed += w->callsUsed;
        h->main.tLogic += w->tLogic; h->main.tInnerEnq += w->tInnerEnq; h->main.tInnerWait += w->tInnerWait; h->main.tIcp += w->tIcp;
    }
    { unsigned long long st8[20]; cudaMemcpy(st8, h->dGen.as<char>() + 8, sizeof st8, cudaMemcpyDeviceToHost); cudaMemset(h->dGen.as<char>() + 8, 0, 160);
      if (getenv("GOICP_DEBUG") && st8[8]) fprintf(stderr, "[phases] cycles per pop: stage(per call) %.0f  A1 %.0f  A2 %.0f (chain on warp 0: %.0f)  C %.0f\n", (double)st8[8] / std::max<double>(1, st8[3]), (double)st8[9] / std::max<double>(1, st8[1]), (double)st8[10] / std::max<double>(1, st8[1]), (double)st8[12] / std::max<double>(1, st8[1]), (double)st8[11] / std::max<double>(1, st8[1]));
      if (getenv("GOICP_DEBUG") && st8[8]) fprintf(stderr, "[detail] A1 items with a fix-up per pop %.2f; corner items per pop %.2f of which with a fix-up %.2f; C: decisions %.0f  +pushes %.0f  +pop %.0f cycles (cumulative from barrier 3)\n", (double)st8[14] / std::max<double>(1, st8[1]), (double)(st8[15] & 0xFFFFFFFFull) / std::max<double>(1, st8[1]), (double)(st8[15] >> 32) / std::max<double>(1, st8[1]), (double)st8[16] / std::max<double>(1, st8[1]), (double)st8[17] / std::max<double>(1, st8[1]), (double)st8[18] / std::max<double>(1, st8[1]));
      if (getenv("GOICP_DEBUG") && st8[8]) fprintf(stderr, "[queue] pops with more than 1024 entries queued: %.1f %%; mean queue length at pop %.0f\n", 100.0 * (double)(st8[13] >> 38) / std::max<double>(1, st8[1]), 16.0 * (double)(st8[13] & ((1ull << 38) - 1)) / std::max<double>(1, st8[1]));
      h->stats[8] = (double)st8[3]; h->stats[9] = (double)st8[1]; h->stats[10] = (double)st8[0]; h->stats[11] = (double)st8[2]; h->stats[12] = (double)st8[4]; h->stats[13] = ctas;
      h->stats[5] = h->main.tLogic; h->stats[6] = h->main.tInnerEnq; h->stats[7] = h->main.tInnerWait;
      if (getenv("GOICP_DEBUG")) {
          fprintf(stderr, "[device] calls %llu pops %llu busy-cycles/pop %.0f corner-misses/pop %.2f busy-cycles %.4g poll-cycles %.4g (ctas %d)\n", st8[3], st8[1], (double)st8[0] / std::max<double>(1, st8[1]), (double)st8[2] / std::max<double>(1, st8[1]), (double)st8[0], (double)st8[4], ctas);
          fprintf(stderr, "[device] claims that found their cell empty: %llu of %llu; cycles in cell hand-back + system fence per call: %.0f; poll cycles per call %.0f\n", st8[5], st8[3], (double)st8[6] / std::max<double>(1, st8[3]), (double)st8[4] / std::max<double>(1, st8[3]));
          fprintf(stderr, "[persistent] loops %lld gather %.3fs publish %.3fs idle-sleep %.3fs loop-busy %.3fs (summed over %d workers)\n", h->main.callsUsed, h->main.tLogic, h->main.tInnerEnq, h->main.tInnerWait, h->main.tIcp, groups);
      } }
    return GOICP_OK;
}

// ---- device-resident search (k_search.cu): the whole batch in one launch, results read back once --------------------------------
static bool host_libm_uses_fma() {   // glibc's ifunc rule for sinf / cosf on x86-64 (sysdeps/x86_64/fpu/multiarch/ifunc-fma.h)
#if defined(__x86_64__)
    __builtin_cpu_init();
    return __builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2");
#else
    return true;
#endif
}
static goicp_status register_resident(Eng* h, const BnbCfg& cfg) {
    const goicp_params& p = h->params;
    const int np = (int)h->probs.size();
    int perSM = goicp_search_occupancy(cfg.smemBytes, h->exact_sums, cfg.threads, cfg.useSmem, cfg.ct);
    { const char* e = getenv("GOICP_CTAS_PER_SM"); if (e && atoi(e) >= 1) perSM = std::min(perSM, atoi(e)); }
    int ctas = h->numSM * perSM;
    { const char* e = getenv("GOICP_CTAS"); if (e && atoi(e) >= 1) ctas = std::min(ctas, atoi(e)); }
    int heapCap = 1 << 15;
    { const char* e = getenv("GOICP_HEAPCAP"); if (e && atoi(e) >= 129) heapCap = atoi(e); }   // test hook: force the overflow re-run
    const int rqCap = 1 << 13;
    const int memoCap = 8192;
    CU(h->sCtl.ensure(sizeof(SearchCtl)));
    CU(h->sHdrs.ensure(goicp_search_hdr_bytes() * (size_t)ctas));
    CU(h->sSlots.ensure(goicp_search_slot_bytes() * (size_t)ctas * SR_NSLOT));
    CU(h->sStates.ensure(sizeof(unsigned) * (size_t)ctas * SR_NSLOT));
    CU(h->sRq.ensure(goicp_search_rnode_bytes() * (size_t)ctas * 2 * rqCap));
    CU(h->sIcp.ensure(sizeof(IcpState) * 2 * (size_t)ctas));
    CU(h->sOuts.ensure(sizeof(PairOut) * (size_t)np));
    CU(h->hOuts.ensure(sizeof(PairOut) * (size_t)np));
    CU(h->qHeaps.ensure(sizeof(HeapEnt) * (size_t)ctas * heapCap));
    if (!cfg.useSmem) CU(h->qScratch.ensure(sizeof(float) * cfg.smemFloats * (size_t)ctas));
    if (h->qMemo.cap < (size_t)32 * memoCap * ctas) { CU(h->qMemo.ensure((size_t)32 * memoCap * ctas)); CU(cudaMemsetAsync(h->qMemo.p, 0, h->qMemo.cap, h->stream)); }
    CU(cudaMemsetAsync(h->sCtl.p, 0, sizeof(SearchCtl), h->stream));
    CU(cudaMemsetAsync(h->sHdrs.p, 0, goicp_search_hdr_bytes() * (size_t)ctas, h->stream));
    CU(cudaMemsetAsync(h->sStates.p, 0, sizeof(unsigned) * (size_t)ctas * SR_NSLOT, h->stream));
    SearchArgs A{};
    A.pairs = h->dPairs.as<PairDev>(); A.npairs = np; A.nCtas = ctas;
    A.rotMinX = p.rotMinX; A.rotMinY = p.rotMinY; A.rotMinZ = p.rotMinZ; A.rotWidth = p.rotWidth;
    A.fma = host_libm_uses_fma() ? 1 : 0;
    { const char* e = getenv("GOICP_LIBM_FMA"); if (e) A.fma = atoi(e) != 0; }
    A.specMax = std::max(0, std::min(h->spec_groups, SR_NGROUP - 4));
    { const char* e = getenv("GOICP_SPEC_GROUPS"); if (e) A.specMax = std::max(0, std::min(atoi(e), SR_NGROUP - 4)); }
    A.quietRamp = 1;
    { const char* e = getenv("GOICP_QUIET_RAMP"); if (e) A.quietRamp = atoi(e) != 0; }
    A.managerRatio = 8;
    { const char* e = getenv("GOICP_MANAGER_RATIO"); if (e && atoi(e) >= 0) A.managerRatio = atoi(e); }
    A.deepCalls = 2048;
    { const char* e = getenv("GOICP_DEEP_CALLS"); if (e && atoi(e) >= 1) A.deepCalls = atoi(e); }
    A.ctl = h->sCtl.as<SearchCtl>(); A.hdrs = h->sHdrs.as<OwnerHdr>(); A.slots = h->sSlots.as<SearchSlot>(); A.states = h->sStates.as<unsigned>(); A.rq = h->sRq.p; A.rqCap = rqCap;
    A.icp = h->sIcp.as<IcpState>(); A.outs = h->sOuts.as<PairOut>();
    A.heaps = h->qHeaps.as<HeapEnt>(); A.heapCap = heapCap; A.gscratch = h->qScratch.as<float>(); A.gstride = cfg.smemFloats; A.NdP = cfg.NdP; A.NdQ = cfg.NdQ; A.useSmem = cfg.useSmem;
    A.memo = reinterpret_cast<uint4*>(h->qMemo.p); A.memoCap = memoCap; A.genCounter = h->dGen.as<unsigned>(); A.gridOff = cfg.gridOff; A.S3p = cfg.S3p;
    cudaEventRecord(h->main.ev0, h->stream);
    CU(goicp_launch_search(A, ctas, cfg.threads, cfg.smemBytes, h->exact_sums, cfg.ct, h->stream));
    cudaEventRecord(h->main.ev1, h->stream);
    CU(cudaMemcpyAsync(h->hOuts.p, h->sOuts.p, sizeof(PairOut) * (size_t)np, cudaMemcpyDeviceToHost, h->stream));
    {
        cudaError_t e = h->main.sync();
        if (e != cudaSuccess) return fail(h, GOICP_ERR_CUDA, "device-resident search kernel: %s", cudaGetErrorString(e));
    }
    float ms = 0; cudaEventElapsedTime(&ms, h->main.ev0, h->main.ev1); h->main.ms[2] += ms; h->main.launches[2] += 1;
    const PairOut* outs = h->hOuts.as<PairOut>();
    std::vector<int> redo;
    long long icpCalls = 0;
    for (int i = 0; i < np; i++) {
        Problem& P = h->probs[i]; const PairOut& o = outs[i];
        reset_search(P);
        if (o.status == GOICP_SR_UNSUPPORTED) return fail(h, GOICP_ERR_UNSUPPORTED, "ICP with trimming supports Nd <= 2048");
        if (o.status != 0) { redo.push_back(i); continue; }
        memcpy(P.optR, o.R, sizeof P.optR); memcpy(P.optT, o.t, sizeof P.optT);
        P.optError = o.optError; P.optComp = o.optComp;
        for (int k = 0; k < 6; k++) P.cnt[k] = o.cnt[k];
        icpCalls += o.cnt[5];
        static const char* kinds[3] = {"Init", "ICP", "BNB"};
        for (int k = 0; k < o.nEvents; k++) tracef(P.trace, "Error*: %g (%s)\n", o.ev[k].v, kinds[o.ev[k].kind % 3]);
        if (o.endKind == 1) tracef(P.trace, "Rotation Queue Empty\nError*: %g, LB: %g\n", P.optError, o.endLb);
        else tracef(P.trace, "Threshold reached\nError*: %g, LB: %g, epsilon: %g\n", P.optError, o.endLb, P.dev.SSEThresh);
        P.phase = PH_DONE;
    }
    h->stats[14] = (double)redo.size();
    if (!redo.empty()) {   // a queue outgrew its per-CTA slab: re-run those pairs with the wave scheduler, which grows the slabs on demand
        std::atomic<int> nx(0);
        h->main.ctaCap = 0;
        goicp_status s2 = register_group(h, h->main, cfg, nx, std::min<int>(64, (int)redo.size()), &redo);
        if (s2) return s2;
    }
    { unsigned long long st8[20]; cudaMemcpy(st8, h->dGen.as<char>() + 8, sizeof st8, cudaMemcpyDeviceToHost); cudaMemset(h->dGen.as<char>() + 8, 0, 160);
      h->stats[8] = (double)st8[3]; h->stats[9] = (double)st8[1]; h->stats[10] = (double)st8[0]; h->stats[11] = (double)st8[2]; h->stats[12] = (double)st8[4]; h->stats[13] = ctas;
      h->stats[15] = (double)st8[7];
      h->stats[5] = h->stats[6] = h->stats[7] = 0;
      h->main.callsLaunched += (long long)st8[3];
      if (getenv("GOICP_DEBUG")) {
          SearchCtl ctl; cudaMemcpy(&ctl, h->sCtl.p, sizeof ctl, cudaMemcpyDeviceToHost);
          fprintf(stderr, "[search] CTA cycles: OuterBnB state machine %.3g, publishing %.3g, help scan %.3g, idle (no pair) %.3g, owner waiting %.3g, ICP(2nd) %.3g; helper calls %llu (abandoned %llu); pair counter ran out at %.1f..%.1f ms\n",
                  (double)ctl.dbg[0], (double)ctl.dbg[1], (double)ctl.dbg[2], (double)ctl.dbg[3], (double)ctl.dbg[4], (double)ctl.dbg[7], ctl.dbg[5], ctl.dbg[6], ctl.dbg[9] * 1e-6, ctl.dbg[8] * 1e-6);
          fprintf(stderr, "[search] CTA cycles: queue pruning after improvements %.3g, rotation-queue pops %.3g\n", (double)ctl.dbg[10], (double)ctl.dbg[11]);
          { std::vector<int> idx(np); for (int i = 0; i < np; i++) idx[i] = i; std::sort(idx.begin(), idx.end(), [&](int x, int y) { return outs[x].tEndMs > outs[y].tEndMs; });
            for (int k = 0; k < std::min(np, 12); k++) { const PairOut& o = outs[idx[k]]; fprintf(stderr, "[search] late pair %d: claimed %.1f ms, finished %.1f ms, %lld calls, %lld rotation pops, %d events\n", idx[k], o.tStartMs, o.tEndMs, o.cnt[0], o.cnt[3], o.nEvents); }
            std::sort(idx.begin(), idx.end(), [&](int x, int y) { return outs[x].cnt[0] > outs[y].cnt[0]; });
            for (int k = 0; k < std::min(np, 12); k++) { const PairOut& o = outs[idx[k]]; fprintf(stderr, "[search] deep pair %d: claimed %.1f ms, finished %.1f ms, %lld calls, %lld rotation pops, %d events\n", idx[k], o.tStartMs, o.tEndMs, o.cnt[0], o.cnt[3], o.nEvents); } }
          std::string a = "[search] pairs finished per 4 ms:", b = "[search] helper calls per 4 ms:  ";
          int last = 0; for (int k = 0; k < 256; k++) if (ctl.finishHist[k] || ctl.helpHist[k]) last = k;
          for (int k = 0; k <= last; k++) { char t[32]; snprintf(t, sizeof t, " %d", ctl.finishHist[k]); a += t; snprintf(t, sizeof t, " %d", ctl.helpHist[k]); b += t; }
          fprintf(stderr, "%s\n%s\n", a.c_str(), b.c_str());
          fprintf(stderr, "[search] ctas %d calls %llu pops %llu busy-cycles/pop %.0f corner-misses/pop %.2f; CTA cycles: total %.4g in calls %.4g scheduling+idle %.4g; icp requests %llu\n", ctas, st8[3], st8[1],
                  (double)st8[0] / std::max<double>(1, st8[1]), (double)st8[2] / std::max<double>(1, st8[1]), (double)st8[7], (double)st8[0], (double)st8[4], st8[6]);
          if (st8[8]) fprintf(stderr, "[phases] cycles per pop: stage(per call) %.0f  A1 %.0f  A2 %.0f (chain on warp 0: %.0f)  C %.0f\n", (double)st8[8] / std::max<double>(1, st8[3]), (double)st8[9] / std::max<double>(1, st8[1]), (double)st8[10] / std::max<double>(1, st8[1]), (double)st8[12] / std::max<double>(1, st8[1]), (double)st8[11] / std::max<double>(1, st8[1]));
      } }
    (void)icpCalls;
    return GOICP_OK;
}

// GoICP::Register (jly_goicp.cpp:878) for every problem of the handle.  One problem: waves on the handle's stream.
// A batch: `groups` worker threads, each with its own stream, pull pairs from a shared counter.
static goicp_status register_all(Eng* h) {
    auto t0 = clk::now();
    goicp_status s;
    if ((s = initialize_all(h))) return s;
    const int np = (int)h->probs.size();
    const BnbCfg cfg = bnb_config(h);
    std::atomic<int> next(0);
    int groups = 1, slots = 1;
    if (np > 1) {
        unsigned cores = std::max(1u, std::thread::hardware_concurrency());
        { const char* e = getenv("LOCAL_WORLD_SIZE"); const int lws = e ? atoi(e) : 1; if (lws > 1) cores = std::max(2u, cores / (unsigned)lws); }   // one process per GPU (torchrun): share the host cores
        groups = h->groups > 0 ? h->groups : (int)std::min<unsigned>(32u, std::max(cores >= 4u ? 4u : 2u, cores));
        slots = h->slots > 0 ? h->slots : std::min(128, std::max(8, (np + groups - 1) / groups));
        groups = std::min(groups, (np + slots - 1) / slots);
    }
    { const char* e = getenv("GOICP_PERSISTENT"); if (e) { h->persistent = atoi(e) != 0; h->persistent_single = atoi(e) == 1; } }   // 0 off, 1 on, 2 batches only
    bool allSmall = true;
    for (auto& P : h->probs) if ((size_t)P.Nd * P.Nm > ((size_t)1 << 21) || (P.dev.doTrim && P.Nd > 2048)) allSmall = false;
    { const char* e = getenv("GOICP_RESIDENT_SEARCH"); if (e) h->resident_search = atoi(e) != 0; }
    if (h->persistent && allSmall && h->shardN <= 1 && (np > 1 || h->persistent_single) && h->resident_search) {
        groups = 0;
        if ((s = register_resident(h, cfg))) return s;
    } else if (h->persistent && allSmall && h->shardN <= 1 && (np > 1 || h->persistent_single)) {
        if (h->slots <= 0) slots = std::min(512, std::max(8, (np + groups - 1) / groups));
        if ((s = register_persistent(h, cfg, groups, slots))) return s;
    } else if (groups <= 1) {
        h->main.ctaCap = 0;
        if ((s = register_group(h, h->main, cfg, next, slots))) return s;
    } else {
        while ((int)h->workers.size() < groups) {
            std::unique_ptr<WaveCtx> w(new WaveCtx());
            if (w->init(true, nullptr) != GOICP_OK) return fail(h, GOICP_ERR_CUDA, "worker stream creation failed");
            h->workers.push_back(std::move(w));
        }
        CU(cudaStreamSynchronize(h->stream));   // inputs / DT / Initialize were enqueued on the handle's stream
        std::vector<goicp_status> st(groups, GOICP_OK);
        std::vector<std::thread> th;
        for (int g = 0; g < groups; g++) {
            WaveCtx* w = h->workers[g].get();
            w->ctaCap = std::max(64, 2 * h->numSM * cfg.perSM / groups);
            memset(w->ms, 0, sizeof w->ms); memset(w->launches, 0, sizeof w->launches); w->waves = w->callsLaunched = 0;
            w->tLogic = w->tInnerEnq = w->tInnerWait = w->tIcp = 0;
            th.emplace_back([h, w, &cfg, &next, &st, g, slots]() { cudaSetDevice(h->device); st[g] = register_group(h, *w, cfg, next, slots); });
        }
        for (auto& t : th) t.join();
        for (int g = 0; g < groups; g++) if (st[g]) return st[g];
        for (int g = 0; g < groups; g++) {
            WaveCtx* w = h->workers[g].get();
            for (int k = 0; k < 5; k++) { h->main.ms[k] += w->ms[k]; h->main.launches[k] += w->launches[k]; }
            h->main.waves += w->waves; h->main.callsLaunched += w->callsLaunched;
            h->main.tLogic += w->tLogic; h->main.tInnerEnq += w->tInnerEnq; h->main.tInnerWait += w->tInnerWait; h->main.tIcp += w->tIcp;
        }
    }
    const double dt = secs_since(t0);
    for (auto& P : h->probs) P.t_reg = dt / std::max(1, np);
    long long used = 0; for (auto& P : h->probs) used += P.cnt[0];
    h->stats[0] = (double)h->main.waves; h->stats[1] = (double)h->main.callsLaunched; h->stats[2] = (double)used; h->stats[3] = groups; h->stats[4] = dt;
    if (!(h->persistent && allSmall && h->shardN <= 1 && (np > 1 || h->persistent_single))) { h->stats[5] = h->main.tLogic; h->stats[6] = h->main.tInnerEnq; h->stats[7] = h->main.tInnerWait; for (int k = 8; k < 16; k++) h->stats[k] = 0; }
    return GOICP_OK;
}

static void fill_result(Eng* h, const Problem& P, goicp_result* out) {
    memset(out, 0, sizeof *out);
    memcpy(out->R, P.optR, sizeof out->R); memcpy(out->t, P.optT, sizeof out->t);
    out->optError = P.optError; out->optComp = P.optComp;
    for (int k = 0; k < 8; k++) out->counters[k] = P.cnt[k];
    out->counters[6] = h->main.launches[0] + h->main.launches[1] + h->main.launches[2] + h->main.launches[3] + h->main.launches[4];
    out->counters[7] = (long long)h->stats[1] - (long long)h->stats[2];
    out->seconds_dt = P.t_dt; out->seconds_register = P.t_reg;
    out->gpu_ms_dt = h->main.ms[0]; out->gpu_ms_bnb = h->main.ms[2]; out->gpu_ms_icp = h->main.ms[3];
    out->status = P.status;
}

static goicp_status ensure_single(Eng* h) {
    if (h->probs.size() != 1) return fail(h, GOICP_ERR_ARG, "no single registration problem set (call goicp_set_model/goicp_set_data)");
    return GOICP_OK;
}
static goicp_status prepare_all(Eng* h) {
    if (!h->haveParams) return fail(h, GOICP_ERR_ARG, "goicp_set_params not called");
    goicp_status s;
    std::atomic<int> bad(0);
    parallel_for((int)h->probs.size(), [&](int i) { goicp_status r = prepare_problem(h, h->probs[i]); if (r) bad.store((int)r); });
    if (bad.load()) return (goicp_status)bad.load();
    if ((s = upload_problems(h))) return s;
    return GOICP_OK;
}
static void set_cloud(std::vector<float>& xyz, std::vector<int>& c, std::vector<float>& f, const float* pxyz, const int32_t* pc, const float* pf, int n) {
    xyz.assign(pxyz, pxyz + 3 * (size_t)n);
    if (pc) c.assign(pc, pc + n); else c.assign(n, 0);
    if (pf) f.assign(pf, pf + 41 * (size_t)n); else f.clear();
}

}  // namespace

// =====================================================================================================================
extern "C" {

const char* goicp_version(void) { return "goicp-b200 0.1 (sm_100a)"; }
const char* goicp_last_error(goicp_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void goicp_params_default(goicp_params* p) {   // shipped config.txt:4-53
    p->MSEThresh = 0.01f;
    p->rotMinX = p->rotMinY = p->rotMinZ = -3.1416f; p->rotWidth = 6.2832f;
    p->transMinX = p->transMinY = p->transMinZ = -0.5f; p->transWidth = 1.0f;
    p->trimFraction = 0.0f;
    p->regularization = 0.0005f; p->regularizationNeighbors = 0.0f; p->regularizationFPFH = 0.0f;
    p->cfpfh = 0; p->norm = 2; p->ponderation = 1;
    p->distTransSize = 20; p->distTransExpandFactor = 2.0;
}

goicp_status goicp_create(goicp_handle* out, int device, void* stream_or_null) {
    if (!out) return GOICP_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) return fail(nullptr, GOICP_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(nullptr, GOICP_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, GOICP_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, GOICP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10) return fail(nullptr, GOICP_ERR_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    Eng* h = new Eng();
    h->device = device; h->numSM = prop.multiProcessorCount;
    if (stream_or_null) { h->stream = (cudaStream_t)stream_or_null; h->ownStream = false; }
    else { if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete h; return fail(nullptr, GOICP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); } h->ownStream = true; }
    if ((e = goicp_preload_bnb()) != cudaSuccess || (e = goicp_preload_dt()) != cudaSuccess || (e = goicp_preload_icp()) != cudaSuccess || (e = goicp_preload_misc()) != cudaSuccess || (e = goicp_preload_search()) != cudaSuccess) {
        delete h; return fail(nullptr, GOICP_ERR_CUDA, "kernel preload failed: %s (library built for sm_100a only)", cudaGetErrorString(e));
    }
    if (h->dGen.ensure(256) != cudaSuccess || cudaMemset(h->dGen.p, 0, 256) != cudaSuccess) { delete h; return fail(nullptr, GOICP_ERR_CUDA, "device allocation failed"); }
    if (h->main.init(false, h->stream) != GOICP_OK) { delete h; return fail(nullptr, GOICP_ERR_CUDA, "event creation failed"); }
    goicp_params_default(&h->params); h->haveParams = true;
    *out = h;
    return GOICP_OK;
}

void goicp_destroy(goicp_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf* bufs[] = {&h->arenaIn, &h->arenaWork, &h->dPairs, &h->dTmp, &h->dTmp2, &h->dTmp3, &h->dSepBits, &h->dSepNx, &h->dSepNxy, &h->sCtl, &h->sHdrs, &h->sSlots, &h->sStates, &h->sRq, &h->sIcp, &h->sOuts};
    for (DevBuf* b : bufs) b->release();
    h->hStage.release(); h->hPairs.release(); h->hOuts.release(); h->qOuts.release(); h->qOrder.release(); h->qIcp.release(); h->qDone.release(); h->qClaim.release(); h->qHeaps.release(); h->qScratch.release(); h->qMemo.release(); h->dGen.release();
    h->main.release();
    for (auto& w : h->workers) w->release();
    if (h->ownStream) cudaStreamDestroy(h->stream);
    delete h;
}

goicp_status goicp_set_model(goicp_handle h, const float* xyz, const int32_t* c, const float* fpfh41, int32_t Nm) {
    if (!h || !xyz || Nm < 1) return h ? fail(h, GOICP_ERR_ARG, "set_model: bad arguments") : GOICP_ERR_ARG;
    if (h->probs.size() != 1) { h->probs.clear(); h->probs.resize(1); }
    Problem& P = h->probs[0];
    set_cloud(P.mxyz, P.mc, P.mf, xyz, c, fpfh41, Nm); P.Nm = Nm;
    P.prepared = P.dt_built = P.initialized = false;
    return GOICP_OK;
}
goicp_status goicp_set_data(goicp_handle h, const float* xyz, const int32_t* c, const float* fpfh41, int32_t Nd) {
    if (!h || !xyz || Nd < 1) return h ? fail(h, GOICP_ERR_ARG, "set_data: bad arguments") : GOICP_ERR_ARG;
    if (h->probs.size() != 1) { h->probs.clear(); h->probs.resize(1); }
    Problem& P = h->probs[0];
    set_cloud(P.dxyz, P.dc, P.df, xyz, c, fpfh41, Nd); P.NdAll = Nd; P.Nd = Nd;
    P.prepared = P.dt_built = P.initialized = false;
    return GOICP_OK;
}
goicp_status goicp_set_params(goicp_handle h, const goicp_params* p) {
    if (!h || !p) return GOICP_ERR_ARG;
    const bool gridChanged = !h->haveParams || p->distTransSize != h->params.distTransSize || p->distTransExpandFactor != h->params.distTransExpandFactor ||
                             (p->trimFraction < 0.001) != (h->params.trimFraction < 0.001) ||
                             p->cfpfh != h->params.cfpfh || p->regularizationFPFH != h->params.regularizationFPFH ||
                             (p->regularizationNeighbors > 0) != (h->params.regularizationNeighbors > 0);   // every key the arena layout (upload_problems) depends on
    h->params = *p; h->haveParams = true;
    for (auto& P : h->probs) { P.initialized = false; if (gridChanged) P.prepared = P.dt_built = false; }
    return GOICP_OK;
}
goicp_status goicp_set_options(goicp_handle h, int32_t exact_sums, int32_t spec_width, int32_t use_dt_replay) {
    if (!h) return GOICP_ERR_ARG;
    if (exact_sums >= 0) h->exact_sums = exact_sums ? 1 : 0;
    if (spec_width >= 0) h->spec_width = spec_width;
    if (use_dt_replay >= 0) h->use_dt_replay = use_dt_replay ? 1 : 0;
    return GOICP_OK;
}

static goicp_status build_dt_impl(goicp_handle h, goicp_dt_info* out, bool replay) {
    goicp_status s;
    if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0];
    if (P.Nm < 1 || P.NdAll < 1) return fail(h, GOICP_ERR_ARG, "build_dt: model and data clouds must be set first");
    cudaSetDevice(h->device);
    if ((s = prepare_all(h))) return s;
    if ((s = build_dt_all(h, replay))) return s;
    h->dtUploaded = false;
    if (out) *out = P.info;
    return GOICP_OK;
}
goicp_status goicp_build_dt(goicp_handle h, goicp_dt_info* out) {
    if (!h) return GOICP_ERR_ARG;
    return build_dt_impl(h, out, h->use_dt_replay && h->params.distTransSize <= 32);
}
goicp_status goicp_build_dt_replay(goicp_handle h, goicp_dt_info* out) { if (!h) return GOICP_ERR_ARG; return build_dt_impl(h, out, true); }

goicp_status goicp_dt_upload(goicp_handle h, const float* dist, const int32_t* nearest_xyz) {
    if (!h) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0];
    if (!P.dt_built) return fail(h, GOICP_ERR_ARG, "dt_upload before build_dt");
    cudaSetDevice(h->device);
    const int S = P.info.size; const size_t S3 = (size_t)S * S * S;
    if (dist) { CU(cudaMemcpyAsync(P.dev.g.dist, dist, sizeof(float) * S3, cudaMemcpyHostToDevice, h->stream)); h->dtUploaded = true; }
    if (nearest_xyz) {
        std::vector<int> vn(S3);
        for (size_t i = 0; i < S3; i++) vn[i] = (nearest_xyz[3 * i + 2] * S + nearest_xyz[3 * i + 1]) * S + nearest_xyz[3 * i];
        CU(cudaMemcpyAsync(P.dev.g.vnear, vn.data(), sizeof(int) * S3, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        CU(goicp_launch_dt_vcell(P.dev.g, h->numSM, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    P.initialized = false;
    return GOICP_OK;
}
goicp_status goicp_dt_download(goicp_handle h, float* dist, int32_t* nearest, int32_t* cellc) {
    if (!h) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0];
    if (!P.dt_built) return fail(h, GOICP_ERR_ARG, "dt_download before build_dt");
    cudaSetDevice(h->device);
    const int S = P.info.size; const size_t S3 = (size_t)S * S * S;
    if (dist) CU(cudaMemcpyAsync(dist, P.dev.g.dist, sizeof(float) * S3, cudaMemcpyDeviceToHost, h->stream));
    std::vector<int> vn;
    if (nearest) { vn.resize(S3); CU(cudaMemcpyAsync(vn.data(), P.dev.g.vnear, sizeof(int) * S3, cudaMemcpyDeviceToHost, h->stream)); }
    CU(cudaStreamSynchronize(h->stream));
    if (nearest) for (size_t i = 0; i < S3; i++) { int v = vn[i]; nearest[3 * i] = v % S; nearest[3 * i + 1] = (v / S) % S; nearest[3 * i + 2] = v / (S * S); }
    if (cellc) {
        for (size_t i = 0; i < S3; i++) cellc[i] = -2;
        for (int c = 0; c < P.info.ncells; c++) cellc[P.cell_vox[c]] = P.cell_colour[c];
    }
    return GOICP_OK;
}
goicp_status goicp_dt_distance(goicp_handle h, const double* xyz, int32_t n, float* dist, int32_t* cell) {
    if (!h || !xyz || !dist || n < 0) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    if (!h->probs[0].dt_built) return fail(h, GOICP_ERR_ARG, "dt_distance before build_dt");
    if (n == 0) return GOICP_OK;
    cudaSetDevice(h->device);
    CU(h->dTmp.ensure(sizeof(double) * 3 * (size_t)n)); CU(h->dTmp2.ensure(sizeof(float) * (size_t)n)); CU(h->dTmp3.ensure(sizeof(int) * 3 * (size_t)n));
    CU(cudaMemcpyAsync(h->dTmp.p, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(goicp_launch_dt_distance(h->dPairs.as<PairDev>(), 0, h->dTmp.as<double>(), n, h->dTmp2.as<float>(), h->dTmp3.as<int>(), h->stream));
    CU(cudaMemcpyAsync(dist, h->dTmp2.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    if (cell) CU(cudaMemcpyAsync(cell, h->dTmp3.p, sizeof(int) * 3 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}
goicp_status goicp_set_nd(goicp_handle h, int32_t nd) {
    if (!h) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0];
    if (nd < 1 || nd > P.NdAll) return fail(h, GOICP_ERR_ARG, "set_nd: %d outside [1,%d]", nd, P.NdAll);
    P.Nd = nd; P.initialized = false;
    return GOICP_OK;
}
goicp_status goicp_initialize(goicp_handle h) {
    if (!h) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    cudaSetDevice(h->device);
    return initialize_all(h);
}
goicp_status goicp_get_weights(goicp_handle h, float* w) {
    if (!h || !w) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0]; if (!P.initialized) return fail(h, GOICP_ERR_ARG, "not initialized");
    CU(cudaMemcpyAsync(w, P.dev.weights, sizeof(float) * P.Nd, cudaMemcpyDeviceToHost, h->stream)); CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}
goicp_status goicp_get_maxrotdis(goicp_handle h, float* out) {
    if (!h || !out) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0]; if (!P.initialized) return fail(h, GOICP_ERR_ARG, "not initialized");
    CU(cudaMemcpyAsync(out, P.dev.maxRotDis, sizeof(float) * GOICP_MAXROTLEVEL * P.Nd, cudaMemcpyDeviceToHost, h->stream)); CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}
goicp_status goicp_get_thresholds(goicp_handle h, float* sse, int32_t* inlier) {
    if (!h) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0]; if (!P.initialized) return fail(h, GOICP_ERR_ARG, "not initialized");
    if (sse) *sse = P.dev.SSEThresh; if (inlier) *inlier = P.dev.inlierNum;
    return GOICP_OK;
}

goicp_status goicp_eval_bounds(goicp_handle h, const float* R, const int32_t* level, int32_t nr, const float* tcube, const int32_t* rot_of,
                               int32_t nt, float* ub, float* lb, int32_t* incomp_minmax, int32_t* fpfh_minmax) {
    if (!h || !R || !level || !tcube || !ub || !lb || nr < 1 || nt < 0) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0]; if (!P.initialized) return fail(h, GOICP_ERR_ARG, "eval_bounds before initialize");
    if (nt == 0) return GOICP_OK;
    cudaSetDevice(h->device);
    std::vector<WaveCube> cubes(nt);
    for (int k = 0; k < nt; k++) { cubes[k].x = tcube[4 * k]; cubes[k].y = tcube[4 * k + 1]; cubes[k].z = tcube[4 * k + 2]; cubes[k].w = tcube[4 * k + 3]; cubes[k].rot = rot_of ? rot_of[k] : 0;
        if (cubes[k].rot < 0 || cubes[k].rot >= nr) return fail(h, GOICP_ERR_ARG, "rot_of[%d]=%d out of range", k, cubes[k].rot); }
    const int nwarps = std::min(nt, h->numSM * 8 * 8);
    const size_t bR = al256(sizeof(float) * 9 * nr), bL = al256(sizeof(int) * nr), bC = al256(sizeof(WaveCube) * nt), bF = al256(sizeof(float) * nt), bI = al256(sizeof(int) * 2 * nt);
    CU(h->dTmp.ensure(bR + bL + bC + 2 * bF + 2 * bI));
    CU(h->dTmp2.ensure(sizeof(float) * (size_t)nwarps * P.Nd));
    char* d = h->dTmp.as<char>();
    float* dR = (float*)d; int* dL = (int*)(d + bR); WaveCube* dC = (WaveCube*)(d + bR + bL); float* dU = (float*)(d + bR + bL + bC); float* dLb = (float*)(d + bR + bL + bC + bF);
    int* dI = (int*)(d + bR + bL + bC + 2 * bF); int* dFm = (int*)(d + bR + bL + bC + 2 * bF + bI);
    CU(cudaMemcpyAsync(dR, R, sizeof(float) * 9 * nr, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dL, level, sizeof(int) * nr, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dC, cubes.data(), sizeof(WaveCube) * nt, cudaMemcpyHostToDevice, h->stream));
    { EvTimer tm(h->main, 2);
      CU(goicp_launch_eval_bounds(h->dPairs.as<PairDev>(), 0, dR, dL, dC, nt, dU, dLb, dI, dFm, h->dTmp2.as<float>(), nwarps, h->stream));
      tm.stop(1); }
    CU(cudaMemcpyAsync(ub, dU, sizeof(float) * nt, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(lb, dLb, sizeof(float) * nt, cudaMemcpyDeviceToHost, h->stream));
    if (incomp_minmax) CU(cudaMemcpyAsync(incomp_minmax, dI, sizeof(int) * 2 * nt, cudaMemcpyDeviceToHost, h->stream));
    if (fpfh_minmax) CU(cudaMemcpyAsync(fpfh_minmax, dFm, sizeof(int) * 2 * nt, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

goicp_status goicp_eval_inclusion(goicp_handle h, const float* R, int32_t level, const float* tcube, int32_t nt, uint8_t* mask, float* resid) {
    if (!h || !R || !tcube || !mask || nt < 0) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0]; if (!P.initialized) return fail(h, GOICP_ERR_ARG, "eval_inclusion before initialize");
    if (level >= GOICP_MAXROTLEVEL) return fail(h, GOICP_ERR_ARG, "level %d >= MAXROTLEVEL", level);
    if (nt == 0) return GOICP_OK;
    cudaSetDevice(h->device);
    std::vector<WaveCube> cubes(nt);
    for (int k = 0; k < nt; k++) { cubes[k].x = tcube[4 * k]; cubes[k].y = tcube[4 * k + 1]; cubes[k].z = tcube[4 * k + 2]; cubes[k].w = tcube[4 * k + 3]; cubes[k].rot = 0; }
    const size_t bR = al256(sizeof(float) * 9), bC = al256(sizeof(WaveCube) * nt), bF = al256(sizeof(float) * (size_t)nt * P.Nd), bM = al256((size_t)nt * P.Nd);
    CU(h->dTmp.ensure(bR + bC + bF + bM));
    char* d = h->dTmp.as<char>();
    float* dR = (float*)d; WaveCube* dC = (WaveCube*)(d + bR); float* dF = (float*)(d + bR + bC); uint8_t* dM = (uint8_t*)(d + bR + bC + bF);
    CU(cudaMemcpyAsync(dR, R, sizeof(float) * 9, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dC, cubes.data(), sizeof(WaveCube) * nt, cudaMemcpyHostToDevice, h->stream));
    CU(goicp_launch_eval_inclusion(h->dPairs.as<PairDev>(), 0, dR, level, dC, nt, dF, dM, std::min(nt, h->numSM * 64), h->stream)); h->main.launches[2]++;
    CU(cudaMemcpyAsync(mask, dM, (size_t)nt * P.Nd, cudaMemcpyDeviceToHost, h->stream));
    if (resid) CU(cudaMemcpyAsync(resid, dF, sizeof(float) * (size_t)nt * P.Nd, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

goicp_status goicp_inner_bnb(goicp_handle h, const float* R, const int32_t* level, const float* opt_error, int32_t n, float* err,
                             float* tnode, int64_t* pops_subcubes) {
    if (!h || !R || !level || !opt_error || !err || n < 0) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0]; if (!P.initialized) return fail(h, GOICP_ERR_ARG, "inner_bnb before initialize");
    cudaSetDevice(h->device);
    std::vector<InnerProb> reqs(n); std::vector<InnerOut> outs;
    for (int k = 0; k < n; k++) {
        if (level[k] >= GOICP_MAXROTLEVEL) return fail(h, GOICP_ERR_ARG, "level %d >= MAXROTLEVEL", level[k]);
        reqs[k].pair = 0; reqs[k].level = level[k]; reqs[k].optError = opt_error[k]; memcpy(reqs[k].R, R + 9 * k, sizeof(float) * 9);
    }
    const BnbCfg cfg = bnb_config(h);
    h->main.ctaCap = 0;
    if ((s = run_inner(h, h->main, cfg, reqs, outs))) return s;
    for (int k = 0; k < n; k++) {
        err[k] = outs[k].err;
        if (tnode) memcpy(tnode + 4 * k, outs[k].node, sizeof(float) * 4);
        if (pops_subcubes) { pops_subcubes[2 * k] = outs[k].pops; pops_subcubes[2 * k + 1] = outs[k].subcubes; }
    }
    return GOICP_OK;
}

goicp_status goicp_icp(goicp_handle h, double* R, double* t, float* err, int32_t* corr) {
    if (!h || !R || !t || !err) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    Problem& P = h->probs[0]; if (!P.initialized) return fail(h, GOICP_ERR_ARG, "icp before initialize");
    cudaSetDevice(h->device);
    std::vector<IcpState> st{make_icp_state(0, 0, R, t)};
    if ((s = run_icp(h, h->main, st))) return s;
    memcpy(R, st[0].R, sizeof(double) * 9); memcpy(t, st[0].t, sizeof(double) * 3); *err = st[0].error;
    if (corr) {
        std::vector<unsigned long long> nn(P.Nd);
        CU(cudaMemcpyAsync(nn.data(), P.dev.nn, sizeof(unsigned long long) * P.Nd, cudaMemcpyDeviceToHost, h->stream)); CU(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < P.Nd; i++) corr[i] = (int)(nn[i] & 0xFFFFFFFFu);
    }
    return GOICP_OK;
}

goicp_status goicp_register(goicp_handle h, goicp_result* out) {
    if (!h || !out) return GOICP_ERR_ARG;
    goicp_status s; if ((s = ensure_single(h))) return s;
    cudaSetDevice(h->device);
    Problem& P = h->probs[0];
    memset(h->main.ms, 0, sizeof h->main.ms); memset(h->main.launches, 0, sizeof h->main.launches); h->main.waves = h->main.callsLaunched = 0; h->main.tLogic = h->main.tInnerEnq = h->main.tInnerWait = h->main.tIcp = 0;
    if (!P.dt_built) { const int nd = P.Nd; if ((s = goicp_build_dt(h, nullptr))) return s; P.Nd = nd; }
    if ((s = register_all(h))) return s;
    h->trace = P.trace;
    fill_result(h, P, out);
    return GOICP_OK;
}
const char* goicp_last_trace(goicp_handle h) { return h ? h->trace.c_str() : ""; }

// ---- batch -----------------------------------------------------------------------------------------------------------
goicp_status goicp_batch_upload(goicp_handle h, const goicp_params* p, int32_t npairs, const goicp_pair_desc* pairs) {
    if (!h || !p || npairs < 1 || !pairs) return GOICP_ERR_ARG;
    cudaSetDevice(h->device);
    h->params = *p; h->haveParams = true;
    h->probs.clear(); h->probs.resize(npairs);
    for (int i = 0; i < npairs; i++) {
        const goicp_pair_desc& d = pairs[i];
        if (!d.model_xyz || !d.data_xyz || d.Nm < 1 || d.NdAll < 1) return fail(h, GOICP_ERR_ARG, "pair %d: empty cloud", i);
    }
    const bool wantF = p->cfpfh != 0;   // descriptors are only copied when the configuration uses them
    parallel_for(npairs, [&](int i) {
        const goicp_pair_desc& d = pairs[i]; Problem& P = h->probs[i];
        set_cloud(P.mxyz, P.mc, P.mf, d.model_xyz, d.model_c, wantF ? d.model_fpfh : nullptr, d.Nm); P.Nm = d.Nm;
        set_cloud(P.dxyz, P.dc, P.df, d.data_xyz, d.data_c, wantF ? d.data_fpfh : nullptr, d.NdAll); P.NdAll = d.NdAll;
        P.Nd = (d.Nd > 0 && d.Nd <= d.NdAll) ? d.Nd : d.NdAll;
    });
    goicp_status s;
    if ((s = prepare_all(h))) return s;
    CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}
goicp_status goicp_batch_run(goicp_handle h, goicp_result* results) {
    if (!h || !results) return GOICP_ERR_ARG;
    if (h->probs.empty()) return fail(h, GOICP_ERR_ARG, "batch_run before batch_upload");
    cudaSetDevice(h->device);
    memset(h->main.ms, 0, sizeof h->main.ms); memset(h->main.launches, 0, sizeof h->main.launches); h->main.waves = h->main.callsLaunched = 0; h->main.tLogic = h->main.tInnerEnq = h->main.tInnerWait = h->main.tIcp = 0;
    goicp_status s;
    if ((s = build_dt_all(h, h->use_dt_replay && h->params.distTransSize <= 32))) return s;
    if ((s = register_all(h))) return s;
    for (size_t i = 0; i < h->probs.size(); i++) fill_result(h, h->probs[i], results + i);
    h->trace = h->probs[0].trace;
    return GOICP_OK;
}
goicp_status goicp_register_batch(goicp_handle h, const goicp_params* p, int32_t npairs, const goicp_pair_desc* pairs, goicp_result* results) {
    goicp_status s;
    if ((s = goicp_batch_upload(h, p, npairs, pairs))) return s;
    return goicp_batch_run(h, results);
}
goicp_status goicp_set_frontier_sharding(goicp_handle h, int32_t rank, int32_t nranks, goicp_allgather_fn allgather, void* user) {
    if (!h || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !allgather)) return h ? fail(h, GOICP_ERR_ARG, "set_frontier_sharding: bad arguments") : GOICP_ERR_ARG;
    h->shardRank = rank; h->shardN = nranks; h->allgather = allgather; h->allgatherUser = user;
    return GOICP_OK;
}
goicp_status goicp_test_exchange(goicp_handle h, const void* send, void* recv, int64_t bytes_per_rank) {
    if (!h || !send || !recv || bytes_per_rank < 1 || !h->allgather) return GOICP_ERR_ARG;
    return h->allgather(send, recv, bytes_per_rank, h->allgatherUser) == 0 ? GOICP_OK : fail(h, GOICP_ERR_ARG, "all-gather callback failed");
}
goicp_status goicp_set_batch_options(goicp_handle h, int32_t groups, int32_t slots) {
    if (!h) return GOICP_ERR_ARG;
    if (groups >= 0) h->groups = groups;
    if (slots >= 0) h->slots = slots;
    { const char* e = getenv("GOICP_TAIL_MULT"); if (e && atoi(e) >= 1) h->tail_spec_mult = std::min(2, atoi(e)); }
    { const char* e = getenv("GOICP_TAIL_THR"); if (e && atoi(e) >= 1) h->tail_thr = atoi(e); }
    { const char* e = getenv("GOICP_BATCH_SPEC"); if (e && atoi(e) >= 0) h->batch_spec_width = atoi(e); }
    { const char* e = getenv("GOICP_MERGE_CALLS"); if (e) h->merge_calls = atoi(e) != 0; }
    { const char* e = getenv("GOICP_BNB_THREADS"); if (e) { int t = atoi(e); if (t >= 64 && t <= goicp_bnb_default_threads() && t % 32 == 0) { h->bnb_threads = t; h->bnb_threads_set = true; } } }
    return GOICP_OK;
}
goicp_status goicp_get_stats(goicp_handle h, double* out16) {
    if (!h || !out16) return GOICP_ERR_ARG;
    for (int k = 0; k < 16; k++) out16[k] = h->stats[k];
    return GOICP_OK;
}
goicp_status goicp_get_timings(goicp_handle h, float* ms5, int64_t* launches5) {
    if (!h) return GOICP_ERR_ARG;
    for (int k = 0; k < 5; k++) { if (ms5) ms5[k] = h->main.ms[k]; if (launches5) launches5[k] = h->main.launches[k]; }
    return GOICP_OK;
}

// ---- Transformation ----------------------------------------------------------------------------------------------------
goicp_status goicp_normalize_cloud(goicp_handle h, double* xyz, int32_t n, double* mean3, double* max_norm) {
    if (!h || !xyz || n < 1) return GOICP_ERR_ARG;
    cudaSetDevice(h->device);
    CU(h->dTmp.ensure(sizeof(double) * 3 * (size_t)n)); CU(h->dTmp2.ensure(sizeof(double) * 4));
    CU(cudaMemcpyAsync(h->dTmp.p, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(goicp_launch_normalize(h->dTmp.as<double>(), n, h->dTmp2.as<double>(), h->stream)); h->main.launches[4]++;
    double out4[4];
    CU(cudaMemcpyAsync(xyz, h->dTmp.p, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(out4, h->dTmp2.p, sizeof out4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (mean3) { mean3[0] = out4[0]; mean3[1] = out4[1]; mean3[2] = out4[2]; }
    if (max_norm) *max_norm = out4[3];
    return GOICP_OK;
}
goicp_status goicp_scale_cloud(goicp_handle h, double* xyz, int32_t n, double scale) {
    if (!h || !xyz || n < 1) return GOICP_ERR_ARG;
    cudaSetDevice(h->device);
    CU(h->dTmp.ensure(sizeof(double) * 3 * (size_t)n));
    CU(cudaMemcpyAsync(h->dTmp.p, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(goicp_launch_scale(h->dTmp.as<double>(), n, scale, h->stream)); h->main.launches[4]++;
    CU(cudaMemcpyAsync(xyz, h->dTmp.p, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}
goicp_status goicp_apply_rigid(goicp_handle h, const double* xyz, int32_t n, const double* R, const double* t, double* out) {
    if (!h || !xyz || !R || !t || !out || n < 1) return GOICP_ERR_ARG;
    cudaSetDevice(h->device);
    CU(h->dTmp.ensure(sizeof(double) * 3 * (size_t)n)); CU(h->dTmp2.ensure(sizeof(double) * 3 * (size_t)n)); CU(h->dTmp3.ensure(sizeof(double) * 12));
    double Rt[12]; memcpy(Rt, R, sizeof(double) * 9); memcpy(Rt + 9, t, sizeof(double) * 3);
    CU(cudaMemcpyAsync(h->dTmp.p, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dTmp3.p, Rt, sizeof Rt, cudaMemcpyHostToDevice, h->stream));
    CU(goicp_launch_apply_rigid(h->dTmp.as<double>(), n, h->dTmp3.as<double>(), h->dTmp2.as<double>(), h->stream)); h->main.launches[4]++;
    CU(cudaMemcpyAsync(out, h->dTmp2.p, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}
goicp_status goicp_rescale_translation(goicp_handle h, double scale, const double* meanT, const double* meanS, const double* R, const double* t, double* out3) {
    if (!h || !meanT || !meanS || !R || !t || !out3) return GOICP_ERR_ARG;
    cudaSetDevice(h->device);
    double in[19]; in[0] = scale; memcpy(in + 1, meanT, 24); memcpy(in + 4, meanS, 24); memcpy(in + 7, R, 72); memcpy(in + 16, t, 24);
    CU(h->dTmp.ensure(sizeof(double) * 24));
    CU(cudaMemcpyAsync(h->dTmp.p, in, sizeof in, cudaMemcpyHostToDevice, h->stream));
    CU(goicp_launch_rescale(h->dTmp.as<double>(), h->dTmp.as<double>() + 19, h->stream)); h->main.launches[4]++;
    CU(cudaMemcpyAsync(out3, h->dTmp.as<double>() + 19, sizeof(double) * 3, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}
goicp_status goicp_rmsd(goicp_handle h, const double* a, const double* b, int32_t n, float* rmsd) {
    if (!h || !a || !b || !rmsd || n < 1) return GOICP_ERR_ARG;
    cudaSetDevice(h->device);
    CU(h->dTmp.ensure(sizeof(double) * 3 * (size_t)n)); CU(h->dTmp2.ensure(sizeof(double) * 3 * (size_t)n)); CU(h->dTmp3.ensure(sizeof(double) * (size_t)n + 64));
    CU(cudaMemcpyAsync(h->dTmp.p, a, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dTmp2.p, b, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    float* dOut = reinterpret_cast<float*>(h->dTmp3.as<char>() + sizeof(double) * (size_t)n);
    CU(goicp_launch_rmsd(h->dTmp.as<double>(), h->dTmp2.as<double>(), n, h->dTmp3.as<double>(), dOut, h->stream)); h->main.launches[4]++;
    CU(cudaMemcpyAsync(rmsd, dOut, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return GOICP_OK;
}

}  // extern "C"
