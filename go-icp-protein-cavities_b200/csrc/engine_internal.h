// Internals of the host engine shared by its translation units (engine.cu: the C ABI; engine_setup.cu: per-pair preprocessing,
// device layout, DT build, Initialize; engine_search.cu: the two schedulers of GoICP::Register).
#pragma once
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


using clk = std::chrono::steady_clock;
inline double secs_since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

extern std::string g_create_error;

// colour codes of the `properties` enum (transformation.hpp:36) that are keys of the identity compatibility map
// (jly_goicp.cpp:66-73); C = 1 is not a key.
static const int KNOWN_PROPS[8] = {8204959, 30894, 15219528, 15231913, 4646984, 16741671, 7566712, 0};
inline bool known_prop(int p) { for (int k = 0; k < 8; k++) if (KNOWN_PROPS[k] == p) return true; return false; }

#define ROUND_HOST(x) ((int)((x) + 0.5))   // jly_3ddt.cpp:30

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
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
inline bool rnode_less(const RNode& n1, const RNode& n2) {   // operator< :64-71
    if (n1.lb != n2.lb) return n1.lb > n2.lb;
    return n1.w < n2.w;
}
// std::priority_queue<ROTNODE> as libstdc++ implements it; spelled out so that equal keys pop in the reference's order
// independently of the standard library this file is compiled against.
inline void rheap_push(std::vector<RNode>& h, const RNode& val) {
    h.push_back(val);
    int hole = (int)h.size() - 1, parent = (hole - 1) / 2;
    while (hole > 0 && rnode_less(h[parent], val)) { h[hole] = h[parent]; hole = parent; parent = (hole - 1) / 2; }
    h[hole] = val;
}
inline RNode rheap_pop(std::vector<RNode>& h) {
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
    // ICP exchange
    bool icpPending = false;
    float icpErr = 0; double icpR[9], icpT[3]; int icpIncomp = 0, compatPose = 0; float initErr = 0;
};

inline void tracef(std::string& s, const char* fmt, ...) {
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
    int batch_spec_width = 4;    // wave scheduler: speculation width inside a batch (pairs already fill the GPU)
    bool dtUploaded = false;   // the DT came from goicp_dt_upload (test hook): no 16-bit distance codes
    int bnb_threads = goicp_bnb_default_threads(); bool bnb_threads_set = false;   // threads per InnerBnB CTA (64..512): more threads = shorter pops, fewer resident calls   // batch: worker streams (0 = auto) and pairs advanced in lock-step per stream
    std::vector<Problem> probs;
    DevBuf arenaIn, arenaWork, dPairs, dTmp, dTmp2, dTmp3, dSepBits, dSepNx, dSepNxy, dSepCid;
    PinBuf hStage, hPairs;
    WaveCtx main;
    DevBuf qHeaps, qScratch, qMemo, dGen;   // per-CTA slabs of the device-resident search: translation queues, staging arrays (large Nd), corner memo; counters
    DevBuf sCtl, sHdrs, sSlots, sStates, sRq, sIcp, sOuts; PinBuf hOuts;   // device-resident search (k_search.cu)
    int spec_groups = SR_NGROUP - 4;   // device-resident search: most rotation-queue nodes with speculative calls per owner CTA (0: never speculate)
    int shardRank = 0, shardN = 1; goicp_allgather_fn allgather = nullptr; void* allgatherUser = nullptr;   // frontier sharding
    std::vector<InnerOut> xSend, xRecv;
    int relaxed = 0, wave_nodes = 64;   // single registrations: 1 = relaxed-order wave search (goicp_set_search_mode), wave_nodes rotation nodes per wave
    int resident = 1;            // 1: device-resident search (k_search.cu) whenever the clouds allow it; 0: wave scheduler (one launch per wave, host-side OuterBnB)
    std::vector<std::unique_ptr<WaveCtx>> workers;
    std::mutex errMutex;
    std::string err, trace;
    float ms[5] = {0, 0, 0, 0, 0}; long long launches[5] = {0, 0, 0, 0, 0};
    double stats[16] = {0};   // see goicp_get_stats
};


typedef goicp_handle_s Eng;

inline goicp_status fail(Eng* h, goicp_status s, const char* fmt, ...) {
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) { std::lock_guard<std::mutex> lk(h->errMutex); h->err = buf; } else g_create_error = buf;
    return s;
}
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(h, GOICP_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

struct EvTimer {   // CUDA-event time of a kernel group on a wave context's stream
    WaveCtx& c; int slot;
    EvTimer(WaveCtx& c_, int slot_) : c(c_), slot(slot_) { cudaEventRecord(c.ev0, c.stream); }
    void stop(int nlaunch) { cudaEventRecord(c.ev1, c.stream); cudaEventSynchronize(c.ev1); float ms = 0; cudaEventElapsedTime(&ms, c.ev0, c.ev1); c.ms[slot] += ms; c.launches[slot] += nlaunch; }
};

// keys of the trimmed-ICP bitonic sort (icp_device.cuh): next power of two >= n, at least 32
inline size_t sort_cap(int n) { size_t c = 32; while (c < (size_t)n) c <<= 1; return c; }
struct BnbCfg { int NdP, NdQ; size_t smemFloats, smemBytes; int useSmem, perSM, threads, gridOff, S3p, ct; };

// engine_setup.cu
goicp_status prepare_problem(Eng* h, Problem& P);
goicp_status upload_problems(Eng* h);
goicp_status upload_pairdevs(Eng* h);
goicp_status build_dt_all(Eng* h, bool replay);
goicp_status initialize_all(Eng* h);
BnbCfg bnb_config(Eng* h);
goicp_status prepare_all(Eng* h);
void set_cloud(std::vector<float>& xyz, std::vector<int>& c, std::vector<float>& f, const float* pxyz, const int32_t* pc, const float* pf, int n);
// engine_search.cu
goicp_status run_inner(Eng* h, WaveCtx& c, const BnbCfg& cfg, std::vector<InnerProb>& reqs, std::vector<InnerOut>& outs);
goicp_status run_icp(Eng* h, WaveCtx& c, std::vector<IcpState>& states);
IcpState make_icp_state(int pair, int mode, const double* R, const double* t);
goicp_status register_all(Eng* h);
void fill_result(Eng* h, const Problem& P, goicp_result* out);
