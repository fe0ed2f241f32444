// Host engine, part 2: GoICP::Register.  Two schedulers:
//  * device-resident (k_search.cu): the whole batch -- OuterBnB, every InnerBnB call, every ICP -- in one cooperative launch;
//    the host uploads, launches, and reads one result record per pair;
//  * wave scheduler: GoICP::OuterBnB (jly_goicp.cpp:582-876) as a host priority queue that follows the reference's pop order
//    exactly, InnerBnB calls launched in waves (one CTA per call, k_bnb.cu) with the calls the reference would make next
//    evaluated speculatively; used for large clouds (multi-kernel ICP), frontier sharding and re-runs with growing queue slabs.
#include "engine_internal.h"

static goicp_status run_inner_local(Eng* h, WaveCtx& c, const BnbCfg& cfg, std::vector<InnerProb>& reqs, std::vector<InnerOut>& outs);

// ---- one launch of InnerBnB calls (handles heap overflow by re-running the overflowed calls with larger heaps) -------
// One wave of InnerBnB calls.  With frontier sharding the calls are dealt round-robin to the ranks, evaluated locally and
// exchanged with one all-gather, so every rank continues with the complete, identical result set.
goicp_status run_inner(Eng* h, WaveCtx& c, const BnbCfg& cfg, std::vector<InnerProb>& reqs, std::vector<InnerOut>& outs) {
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
        for (int i = 0; i < m; i++) {
            if (ho[i].status == 4) again.push_back(todo[i]);   // the translation queue outgrew its slab: re-run with a larger one
            else if (ho[i].status != 0) return fail(h, GOICP_ERR_CUDA, "InnerBnB call %d of pair %d ended with internal status %d", todo[i], reqs[todo[i]].pair, ho[i].status);
            else outs[todo[i]] = ho[i];
        }
        todo.swap(again);
        if (!todo.empty()) { heapCap *= 4; const size_t fit = ((size_t)4 << 30) / sizeof(HeapEnt) / (size_t)heapCap; maxCtas = (int)std::max<size_t>(1, std::min<size_t>((size_t)maxCtas, fit)); }
    }
    if (!todo.empty()) return fail(h, GOICP_ERR_OVERFLOW, "translation queue exceeded %d entries", heapCap);
    c.callsLaunched += n;
    return GOICP_OK;
}

// ---- ICP / scoring pipeline for a set of states ----------------------------------------------------------------------
goicp_status run_icp(Eng* h, WaveCtx& c, std::vector<IcpState>& states) {
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
        for (int i = 0; i < n; i++) if (states[i].status != 0) return fail(h, GOICP_ERR_ARG, "trimmed ICP without its sort workspace (trimFraction changed after the clouds were set?)");
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
    for (int i = 0; i < n; i++) if (states[i].status != 0) return fail(h, GOICP_ERR_ARG, "trimmed ICP without its sort workspace (trimFraction changed after the clouds were set?)");
    return GOICP_OK;
}

IcpState make_icp_state(int pair, int mode, const double* R, const double* t) {
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
                P.cache.clear(); P.quiet = 0;
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
    P.cache.clear();
    P.phase = PH_CHILD_LB;
}

// Requests of one blocked problem: the blocking call first, then speculation in the reference's expected order.
static void gather_requests(Eng* h, int pi, std::vector<InnerProb>& reqs, std::vector<ReqTag>& tags) {
    Problem& P = h->probs[pi];
    const float SSE = P.dev.SSEThresh;
    std::unordered_set<unsigned long long> seen;
    auto want = [&](const RNode& par, int j, int kind, const RNode& ch, const float* R) {
        const unsigned long long key = call_key(par.id, j, kind);
        auto it = P.cache.find(key);
        if (it != P.cache.end() && it->second.entryOpt == P.optError) return;
        if (!seen.insert(key).second) return;
        // Q2: the reference indexes maxRotDis[level] without a bound check (undefined beyond level 19); we clamp.
        const int lbLevel = std::min(ch.l, GOICP_MAXROTLEVEL - 1);
        InnerProb ip; ip.pair = pi; ip.level = kind ? lbLevel : -1; ip.optError = P.optError;
        memcpy(ip.R, R, sizeof ip.R);
        reqs.push_back(ip); tags.push_back(ReqTag{pi, key, P.optError, false});
    };
    // a call that is cached under the current incumbent or already in flight needs no request (and no rotation matrix)
    auto known = [&](const RNode& par, int j, int kind) {
        const unsigned long long key = call_key(par.id, j, kind);
        auto it = P.cache.find(key);
        return it != P.cache.end() && it->second.entryOpt == P.optError;
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
    // inside a batch the pairs themselves fill the GPU
    const int specw = h->probs.size() > 1 ? std::min(h->spec_width, h->batch_spec_width) : h->spec_width;
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
    P.nextId = 1; P.quiet = 0; P.optComp = 0; P.lastLb = 0; P.status = 0; P.icpPending = false;
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


// ---- relaxed-order wave search of ONE registration (north_star (b): "the host priority queue expands frontier waves") ----
// The exact-order schedulers reproduce the reference's visitation order, so a single deep registration advances one rotation node at
// a time along its dependency chain.  Here every wave pops the `wave_nodes` best rotation nodes at once, evaluates ALL their children
// (upper- and lower-bound InnerBnB call of each cube as one request) in one launch under the incumbent error of the wave's start,
// and then applies the results together: the best improving upper bound becomes the incumbent (followed by updateCompatibilities +
// ICP as in jly_goicp.cpp:791-840 and the queue pruning of :843-853), children whose lower bound stays below it are pushed.  Node
// visitation order differs from the reference; the certificate is the reference's own: the search ends when the queue is empty or
// optError - (smallest lower bound in the queue) <= SSEThresh (:685).  With frontier sharding (goicp_set_frontier_sharding) the
// requests of a wave are dealt to the ranks and their results exchanged once per wave (run_inner), so every rank applies the same
// results and holds the same queue and incumbent: the minimum over the wave's upper bounds is the per-wave "allreduce(min)".
static goicp_status register_relaxed(Eng* h, WaveCtx& c, const BnbCfg& cfg, int pi) {
    const goicp_params& p = h->params;
    Problem& P = h->probs[pi];
    reset_search(P);
    goicp_status s;
    std::vector<IcpState> icps;
    icps.push_back(make_icp_state(pi, 1, nullptr, nullptr)); icps.push_back(make_icp_state(pi, 0, P.optR, P.optT));
    P.phase = PH_WAIT_INIT;
    if ((s = run_icp(h, c, icps))) return s;
    absorb_icp(P, icps[0]); absorb_icp(P, icps[1]);
    after_icp(h, pi);   // initial error, first ICP, root node pushed (:601-664)
    const float SSE = P.dev.SSEThresh;
    struct Item { RNode node; float R[9]; bool lbOnly; float ub; };
    std::vector<Item> items, carry;   // carry: cubes whose upper bound improved the incumbent of their wave: their lower-bound call is still due (:856-861)
    std::vector<InnerProb> reqs; std::vector<InnerOut> outs;
    const int W = std::max(1, h->wave_nodes);
    for (;;) {
        items.clear(); items.swap(carry);
        int popped = 0;
        while (!P.q.empty() && popped < W && (P.optError - P.q.front().lb) > SSE) {
            const RNode par = rheap_pop(P.q); P.cnt[3]++; popped++;
            for (int j = 0; j < 8; j++) {
                Item it; it.node = child_of(par, j); it.lbOnly = false; it.ub = 0.f;
                if (!child_rotation(it.node, it.R)) continue;
                items.push_back(it);
            }
        }
        if (items.empty()) {
            if (P.q.empty()) tracef(P.trace, "Rotation Queue Empty\nError*: %g, LB: %g\n", P.optError, P.lastLb);
            else { tracef(P.trace, "Threshold reached\nError*: %g, LB: %g, epsilon: %g\n", P.optError, P.q.front().lb, SSE); P.cnt[3]++; }
            break;
        }
        reqs.resize(items.size());
        for (size_t k = 0; k < items.size(); k++) {
            const int lbLevel = std::min(items[k].node.l, GOICP_MAXROTLEVEL - 1);
            reqs[k].pair = pi; reqs[k].level = items[k].lbOnly ? lbLevel : GOICP_REQ_BOTH + lbLevel; reqs[k].optError = P.optError;
            memcpy(reqs[k].R, items[k].R, sizeof reqs[k].R);
        }
        c.waves++;
        if ((s = run_inner(h, c, cfg, reqs, outs))) return s;
        int best = -1;
        for (size_t k = 0; k < items.size(); k++) {
            const InnerOut& o = outs[k];
            P.cnt[0]++; P.cnt[1] += o.pops; P.cnt[2] += o.subcubes;
            if (items[k].lbOnly) continue;
            P.cnt[4]++;
            items[k].ub = o.err;
            if (o.err < P.optError && (best < 0 || o.err < outs[best].err)) best = (int)k;
        }
        const float entryOpt = P.optError;
        if (best >= 0) {   // :771-840 for the wave's best cube
            const InnerOut& o = outs[best];
            P.optError = o.err;
            for (int k = 0; k < 9; k++) P.optR[k] = items[best].R[k];
            P.optT[0] = o.node[0] + o.node[3] / 2; P.optT[1] = o.node[1] + o.node[3] / 2; P.optT[2] = o.node[2] + o.node[3] / 2;
            icps.clear();
            icps.push_back(make_icp_state(pi, 2, P.optR, P.optT)); icps.push_back(make_icp_state(pi, 0, P.optR, P.optT));
            if ((s = run_icp(h, c, icps))) return s;
            absorb_icp(P, icps[0]); absorb_icp(P, icps[1]);
            P.optComp = P.compatPose;
            tracef(P.trace, "Error*: %g (BNB)\n", P.optError);
            P.cnt[5]++;
            if (P.icpErr < P.optError) {
                P.optError = P.icpErr; memcpy(P.optR, P.icpR, sizeof P.optR); memcpy(P.optT, P.icpT, sizeof P.optT); P.optComp = P.icpIncomp;
                tracef(P.trace, "Error*: %g (ICP)\n", P.icpErr);
            }
        }
        for (size_t k = 0; k < items.size(); k++) {   // lower bounds (:856-871)
            const InnerOut& o = outs[k];
            float lb;
            if (items[k].lbOnly) lb = o.err;
            else if (o.ran2) { P.cnt[0]++; P.cnt[1] += o.pops2; P.cnt[2] += o.subcubes2; lb = o.err2; }
            else { Item it = items[k]; it.lbOnly = true; carry.push_back(it); continue; }   // its upper bound improved on the wave's incumbent: lower bound in the next wave
            P.lastLb = lb;
            if (!(lb >= P.optError)) { RNode nr = items[k].node; nr.ub = items[k].ub; nr.lb = lb; nr.id = P.nextId++; rheap_push(P.q, nr); }
        }
        if (best >= 0) {   // :843-853 (order-independent form: drop every node whose lower bound reaches the incumbent)
            std::vector<RNode> qn;
            for (const RNode& n : P.q) if (n.lb < P.optError) rheap_push(qn, n);
            P.q.swap(qn);
        }
        (void)entryOpt;
    }
    P.phase = PH_DONE;
    (void)p;
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
    // per-CTA slabs: translation queue (32 B per entry) and rotation queue (2 x 32 B per node); a pair that outgrows either is re-run by
    // the wave scheduler with growing slabs -- slow, so the slabs are generous (4 MB + 1 MB per CTA)
    int heapCap = 1 << 17;
    { const char* e = getenv("GOICP_HEAPCAP"); if (e && atoi(e) >= 129) heapCap = atoi(e); }   // test hook: force the overflow re-run
    const int rqCap = 1 << 14;
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
    A.walkEvery = np > 1 ? 16 : 8;   // measured: batch step 500 / 496 / 488 / 477 ms and pair 2 alone 86 / 71 / 66 / 71 ms at 1 / 4 / 8 / 16
    { const char* e = getenv("GOICP_WALK_EVERY"); if (e) A.walkEvery = std::max(1, atoi(e)); }
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
        if (o.status == GOICP_SR_UNSUPPORTED) return fail(h, GOICP_ERR_ARG, "trimmed ICP without its sort workspace (trimFraction changed after the clouds were set?)");
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
          fprintf(stderr, "[search] CTA cycles: queue pruning after improvements %.3g, rotation-queue pops %.3g; rotation nodes popped with calls made ahead %llu / without %llu; children whose result was there when the search reached them %llu / waited for %llu\n", (double)ctl.dbg[10], (double)ctl.dbg[11], ctl.dbg[12], ctl.dbg[13], ctl.dbg[14], ctl.dbg[15]);
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
goicp_status register_all(Eng* h) {
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
    { const char* e = getenv("GOICP_PERSISTENT"); if (e) h->resident = atoi(e) != 0; }   // 0: wave scheduler (one launch per wave, OuterBnB on the host)
    bool allSmall = true;
    for (auto& P : h->probs) if ((size_t)P.Nd * P.Nm > ((size_t)1 << 21) || (P.dev.doTrim && P.Nd > 2048)) allSmall = false;
    const bool resident = h->resident && allSmall && h->shardN <= 1 && !(h->relaxed && np == 1);
    if (h->relaxed && np == 1) {
        h->main.ctaCap = 0;
        if ((s = register_relaxed(h, h->main, cfg, 0))) return s;
    } else if (resident) {
        groups = 0;
        if ((s = register_resident(h, cfg))) return s;
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
    if (!resident) { h->stats[5] = h->main.tLogic; h->stats[6] = h->main.tInnerEnq; h->stats[7] = h->main.tInnerWait; for (int k = 8; k < 16; k++) h->stats[k] = 0; }
    return GOICP_OK;
}

void fill_result(Eng* h, const Problem& P, goicp_result* out) {
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

