// Device-resident registration: GoICP::Register (jly_goicp.cpp:878) = OuterBnB (:582-876) + every InnerBnB call (:286-579) + every
// ICP refinement (:102-178) of a whole BATCH of pairs in ONE kernel launch, with no host round trip.
//
//  * The kernel is launched with (SMs x occupancy) CTAs.  A CTA claims a pair from a device counter and becomes its OWNER: warp 0
//    runs the rotation branch-and-bound as a small state machine (rotation queue = a binary heap in the CTA's global slab that
//    follows libstdc++'s push_heap / pop_heap step for step, so equal (lb, w) keys pop in the reference's order; rotation matrices
//    from libm_exact.h = the host libm's own operation sequence), and the whole CTA executes what that needs: InnerBnB calls
//    (bnb_device.cuh:inner_call -- the upper- and lower-bound call of one rotation cube as one request) and GoICP::ICP calls
//    (icp_device.cuh).  While unclaimed pairs remain, every CTA works on its own pair strictly in the reference's order: nothing
//    speculative is ever computed and nothing is polled.
//  * When the pair counter runs out, CTAs without a pair become HELPERS.  An owner then publishes, in the order the reference
//    would make them if the incumbent error does not change, the calls of its current node's remaining children and of the next
//    rotation-queue nodes into its slot array in global memory; helpers scan the owners' slot arrays, claim a slot with a
//    compare-and-swap, run the call and store the result in the slot.  The owner consumes a result only when the reference's order
//    reaches that call (with the same incumbent), so the search stays semantically sequential: optimum, node counters and the
//    improvement trace are those of the reference.  An improvement bumps the owner's generation word; speculative calls made
//    under the old incumbent see it at their next queue pop and stop.
//  * Results (R, t, optError, optComp, counters, improvement events) are written to one record per pair; the host reads the
//    array back once after the kernel.
#include "bnb_device.cuh"
#include "libm_exact.h"
#include "search_dev.h"

namespace {

enum { SL_FREE = 0, SL_QUEUED = 1, SL_RUNNING = 2, SL_DONE = 3, SL_SKIP = 4 };
enum { OW_NONE = 0, OW_START, OW_AFTER_INIT, OW_POP, OW_CHILD, OW_AFTER_IMPROVE };
enum { ACT_NONE = 0, ACT_CALL, ACT_ICP, ACT_EXIT };
enum { RQ_ACTION = 0, RQ_SPAWN, RQ_FINDWORK };

struct alignas(16) RNodeD { float lb; int l; int group; float ub; float a, b, c, w; };   // ROTNODE (jly_goicp.h:59-73) + the slot group that holds its children's calls

struct OwnerSh {
    int pair, phase;
    float optError, SSE;
    double optR[9], optT[3];
    int optComp;
    long long cnt[6];
    RNodeD par; int j, needLbOnly; float ubChild, lastLb;
    int heapN, nEvents, status;
    int grpUse[SR_NGROUP];       // 0 free, 1 current node / queue node, 2 draining (calls of an abandoned incumbent may still run)
    int specGroups;              // groups attached to queue nodes
    int action, actOwner, actSlot, actCancel;
    int noMorePairs, lastSpawnJ;
    int cand[64]; int ncand;
};

__device__ __forceinline__ unsigned ld_vol(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
__device__ __forceinline__ int ld_voli(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_vol(unsigned* p, unsigned v) { *reinterpret_cast<volatile unsigned*>(p) = v; }

__device__ __forceinline__ bool rnode_less(const RNodeD& n1, const RNodeD& n2) {   // ROTNODE operator< (jly_goicp.h:64-71)
    if (n1.lb != n2.lb) return n1.lb > n2.lb;
    return n1.w < n2.w;
}
__device__ __forceinline__ RNodeD ld_node(const RNodeD* p) { const float4* q = reinterpret_cast<const float4*>(p); const float4 a = q[0], b = q[1]; RNodeD n; n.lb = a.x; n.l = __float_as_int(a.y); n.group = __float_as_int(a.z); n.ub = a.w; n.a = b.x; n.b = b.y; n.c = b.z; n.w = b.w; return n; }
__device__ __forceinline__ void st_node(RNodeD* p, const RNodeD& n) { float4* q = reinterpret_cast<float4*>(p); q[0] = make_float4(n.lb, __int_as_float(n.l), __int_as_float(n.group), n.ub); q[1] = make_float4(n.a, n.b, n.c, n.w); }
// std::priority_queue<ROTNODE>::push / pop as libstdc++ implements them (__push_heap, __adjust_heap); one lane
__device__ void rq_push(RNodeD* h, int& n, const RNodeD& val) {
    int hole = n++;
    while (hole > 0) {
        const int parent = (hole - 1) / 2;
        const RNodeD pe = ld_node(h + parent);
        if (!rnode_less(pe, val)) break;
        st_node(h + hole, pe); hole = parent;
    }
    st_node(h + hole, val);
}
__device__ RNodeD rq_pop(RNodeD* h, int& n) {
    const RNodeD top = ld_node(h);
    const int len = --n;
    if (len > 0) {
        const RNodeD val = ld_node(h + len);
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            RNodeD c1 = ld_node(h + child); const RNodeD c0 = ld_node(h + child - 1);
            if (rnode_less(c1, c0)) { child--; c1 = c0; }
            st_node(h + hole, c1); hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); st_node(h + hole, ld_node(h + child - 1)); hole = child - 1; }
        while (hole > 0) {
            const int parent = (hole - 1) / 2;
            const RNodeD pe = ld_node(h + parent);
            if (!rnode_less(pe, val)) break;
            st_node(h + hole, pe); hole = parent;
        }
        st_node(h + hole, val);
    }
    return top;
}

// child cube j of a rotation node (jly_goicp.cpp:707-712) and its rotation matrix (:716-747); false: the cube lies outside the pi-ball (:723)
__device__ __forceinline__ RNodeD child_of(const RNodeD& par, int j) {
    RNodeD nr; nr.w = par.w / 2; nr.l = par.l + 1; nr.group = -1; nr.ub = 0.f; nr.lb = 0.f;
    nr.a = par.a + (float)(j & 1) * nr.w; nr.b = par.b + (float)((j >> 1) & 1) * nr.w; nr.c = par.c + (float)((j >> 2) & 1) * nr.w;
    return nr;
}
__device__ bool child_rotation(bool fma, const RNodeD& nr, float* R) {
    float v1 = nr.a + nr.w / 2, v2 = nr.b + nr.w / 2, v3 = nr.c + nr.w / 2;
    const float t = __fsqrt_rn(v1 * v1 + v2 * v2 + v3 * v3);
    if ((double)t - GOICP_SQRT3 * (double)nr.w / 2 > GOICP_PI) return false;
    if (t > 0) {
        v1 = __fdiv_rn(v1, t); v2 = __fdiv_rn(v2, t); v3 = __fdiv_rn(v3, t);
        const float ct = libm_exact::cosf_glibc(fma, t), ct2 = 1 - ct, st = libm_exact::sinf_glibc(fma, t);
        const float tmp121 = v1 * v2 * ct2, tmp122 = v3 * st, tmp131 = v1 * v3 * ct2, tmp132 = v2 * st, tmp231 = v2 * v3 * ct2, tmp232 = v1 * st;
        R[0] = ct + v1 * v1 * ct2; R[1] = tmp121 - tmp122; R[2] = tmp131 + tmp132;
        R[3] = tmp121 + tmp122; R[4] = ct + v2 * v2 * ct2; R[5] = tmp231 - tmp232;
        R[6] = tmp131 - tmp132; R[7] = tmp231 + tmp232; R[8] = ct + v3 * v3 * ct2;
    } else {
        for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.f : 0.f;
    }
    return true;
}

struct Cta {   // what every routine of this file needs about the CTA's place in the world
    const SearchArgs& A;
    OwnerSh& os;
    SearchSlot* slots;       // this CTA's own slots
    OwnerHdr* hdr;           // this CTA's header
    RNodeD* rq;              // this CTA's rotation queue
    IcpState* icp;           // this CTA's two ICP states
    int me, lane;
};

__device__ __forceinline__ void add_event(Cta& c, int kind, float v) {
    OwnerSh& os = c.os;
    if (os.nEvents < SR_MAXEV) { PairOut& o = c.A.outs[os.pair]; o.ev[os.nEvents].kind = kind; o.ev[os.nEvents].v = v; }
    else os.status = GOICP_SR_OVERFLOW;
    os.nEvents++;
}

// a request for child `ch` (rotation R) of the current incumbent into slot s; state is set by the caller
__device__ __forceinline__ void fill_request(Cta& c, SearchSlot* s, const RNodeD& ch, const float* R, bool lbOnly, unsigned prio) {
    const int lbLevel = min(ch.l, GOICP_MAXROTLEVEL - 1);   // Q2: the reference indexes maxRotDis[level] without a bound check; clamped
    s->pr.pair = c.os.pair; s->pr.level = lbOnly ? lbLevel : GOICP_REQ_BOTH + lbLevel; s->pr.optError = c.os.optError;
#pragma unroll
    for (int k = 0; k < 9; k++) s->pr.R[k] = R[k];
    s->prio = prio; s->gen = ld_vol(&c.hdr->gen);
}

// the incumbent changed (or the pair ended): every speculative call is void.  Unclaimed ones are withdrawn, running ones are left to
// finish (they notice the new generation at their next pop); their groups drain.  Called by all lanes of warp 0.
__device__ void invalidate_spec(Cta& c) {
    OwnerSh& os = c.os;
    if (c.lane == 0) { st_vol(&c.hdr->gen, ld_vol(&c.hdr->gen) + 1u); __threadfence(); }
    __syncwarp();
    for (int s = c.lane; s < SR_NSLOT; s += 32) {
        if (os.grpUse[s >> 3] == 0) continue;
        SearchSlot* sl = c.slots + s;
        const unsigned st = ld_vol(&sl->state);
        if (st == SL_QUEUED) { if (atomicCAS(&sl->state, (unsigned)SL_QUEUED, (unsigned)SL_FREE) == SL_QUEUED) atomicSub(&c.hdr->nQueued, 1u); }
        else if (st == SL_DONE || st == SL_SKIP) st_vol(&sl->state, SL_FREE);
    }
    __syncwarp();
    if (c.lane == 0) { for (int g = 0; g < SR_NGROUP; g++) if (os.grpUse[g] != 0) os.grpUse[g] = 2; os.specGroups = 0; }
    __syncwarp();
}
// a free slot group (all 8 slots FREE), or -1.  Draining groups whose calls have all finished are recycled.  Lane 0.
__device__ int alloc_group(Cta& c, int reserve) {
    OwnerSh& os = c.os;
    if (reserve > 0) { int used = 0; for (int g = 0; g < SR_NGROUP; g++) used += os.grpUse[g] != 0; if (used + reserve >= SR_NGROUP) return -1; }   // speculation leaves groups for the node being expanded
    for (int g = 0; g < SR_NGROUP; g++) {
        if (os.grpUse[g] == 2) {
            bool busy = false;
            for (int k = 0; k < 8; k++) {
                SearchSlot* sl = c.slots + 8 * g + k;
                unsigned st = ld_vol(&sl->state);
                if (st == SL_QUEUED) { if (atomicCAS(&sl->state, (unsigned)SL_QUEUED, (unsigned)SL_FREE) == SL_QUEUED) { atomicSub(&c.hdr->nQueued, 1u); st = SL_FREE; } else st = SL_RUNNING; }
                if (st == SL_RUNNING) busy = true; else if (st != SL_FREE) st_vol(&sl->state, SL_FREE);
            }
            if (!busy) os.grpUse[g] = 0;
        }
        if (os.grpUse[g] == 0) { os.grpUse[g] = 1; return g; }
    }
    return -1;
}

__device__ void finish_pair(Cta& c, int endKind) {   // lane 0
    OwnerSh& os = c.os;
    PairOut& o = c.A.outs[os.pair];
    for (int k = 0; k < 9; k++) o.R[k] = os.optR[k];
    for (int k = 0; k < 3; k++) o.t[k] = os.optT[k];
    o.optError = os.optError; o.optComp = os.optComp;
    for (int k = 0; k < 6; k++) o.cnt[k] = os.cnt[k];
    o.status = os.status; o.endKind = endKind; o.endLb = endKind == 2 ? os.par.lb : os.lastLb; o.nEvents = min(os.nEvents, SR_MAXEV);
}

// ---- the OuterBnB state machine, lane 0 of the owner's warp 0.  Returns what it needs next. --------------------------------------
__device__ __noinline__ int owner_serial(Cta& c) {
    OwnerSh& os = c.os;
    const SearchArgs& A = c.A;
    for (;;) {
        switch (os.phase) {
        case OW_START: {   // initial error (:601-627) and ICP from the identity (:634)
            const PairDev& P = A.pairs[os.pair];
            os.SSE = P.SSEThresh; os.optComp = 0; os.status = 0; os.nEvents = 0; os.heapN = 0; os.lastLb = 0.f; os.needLbOnly = 0;
            for (int k = 0; k < 6; k++) os.cnt[k] = 0;
            for (int k = 0; k < 9; k++) os.optR[k] = (k % 4 == 0) ? 1.0 : 0.0;
            os.optT[0] = os.optT[1] = os.optT[2] = 0.0;
            for (int q = 0; q < 2; q++) {
                IcpState& s = c.icp[q];
                s.pair = os.pair; s.mode = q == 0 ? 1 : 0;
                for (int k = 0; k < 9; k++) s.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
                for (int k = 0; k < 3; k++) { s.t[k] = 0.0; s.mu_m[k] = 0.0; s.mu_d[k] = 0.0; }
                s.err = -1.f; s.iter = 0; s.done = 0; s.status = 0; s.error = 0.f; s.incomp = 0; s.fpfh = 0.f; s.compat_pose = 0;
            }
            __threadfence();
            os.phase = OW_AFTER_INIT; os.action = ACT_ICP;
            return RQ_ACTION;
        }
        case OW_AFTER_INIT: {   // :623-661
            const PairDev& P = A.pairs[os.pair];
            const volatile IcpState* s0 = c.icp; const volatile IcpState* s1 = c.icp + 1;
            if (s0->status != 0 || s1->status != 0) { os.status = GOICP_SR_UNSUPPORTED; finish_pair(c, 0); os.phase = OW_NONE; break; }
            float optError = s0->error;
            if (P.reg > 0) optError += P.reg * (float)(P.Nd * P.Nd);                       // :623
            if (P.regF > 0) optError += P.regF * (float)(100 * 8 * 100 * 8);               // :624
            if (P.regN > 0) optError += P.regN * (float)(P.Nd * 6 * P.Nd * 6);             // :625
            os.optError = optError;
            add_event(c, 0, optError);
            os.cnt[5]++;
            const float icpErr = s1->error;
            if (icpErr < os.optError) {                                                    // :636-661
                os.optError = icpErr;
                for (int k = 0; k < 9; k++) os.optR[k] = s1->R[k];
                for (int k = 0; k < 3; k++) os.optT[k] = s1->t[k];
                os.optComp = s1->incomp;
                add_event(c, 1, icpErr);
            }
            RNodeD root; root.a = A.rotMinX; root.b = A.rotMinY; root.c = A.rotMinZ; root.w = A.rotWidth; root.l = 0; root.lb = 0.f; root.ub = 0.f; root.group = -1;
            rq_push(c.rq, os.heapN, root);
            os.phase = OW_POP;
            break;
        }
        case OW_POP: {
            if (os.status != 0) { finish_pair(c, 0); os.phase = OW_NONE; break; }
            if (os.heapN == 0) { finish_pair(c, 1); os.phase = OW_NONE; break; }                   // :670-677
            os.par = rq_pop(c.rq, os.heapN); os.cnt[3]++;
            if (os.par.group >= 0) os.specGroups--;
            if ((os.optError - os.par.lb) <= os.SSE) { finish_pair(c, 2); os.phase = OW_NONE; break; }   // :685
            if (os.par.group < 0) {
                int g;
                while ((g = alloc_group(c, 0)) < 0) __nanosleep(200);   // every group still drains calls of an abandoned incumbent (they stop at their next pop)
                os.par.group = g;
            }
            os.j = 0; os.needLbOnly = 0; os.phase = OW_CHILD; os.lastSpawnJ = 0;
            return RQ_SPAWN;
        }
        case OW_CHILD: {
            if (os.j >= 8) {
                for (int k = 0; k < 8; k++) st_vol(&c.slots[8 * os.par.group + k].state, SL_FREE);
                os.grpUse[os.par.group] = 0;
                os.phase = OW_POP; break;
            }
            SearchSlot* sl = c.slots + 8 * os.par.group + os.j;
            unsigned st = ld_vol(&sl->state);
            if (st == SL_SKIP) { os.j++; break; }
            if (st == SL_FREE) {
                if (A.specMax > 0 && os.lastSpawnJ != os.j && ld_voli(&A.ctl->nextPair) >= A.npairs) { os.lastSpawnJ = os.j; return RQ_SPAWN; }   // helpers may have appeared since the last pop
                const RNodeD ch = child_of(os.par, os.j);
                float R[9];
                if (!child_rotation(A.fma != 0, ch, R)) { os.j++; break; }
                fill_request(c, sl, ch, R, os.needLbOnly != 0, 0u);
                st_vol(&sl->state, SL_RUNNING);
                os.action = ACT_CALL; os.actOwner = c.me; os.actSlot = 8 * os.par.group + os.j; os.actCancel = 0;
                return RQ_ACTION;
            }
            if (st == SL_QUEUED) {
                if (atomicCAS(&sl->state, (unsigned)SL_QUEUED, (unsigned)SL_RUNNING) == SL_QUEUED) {
                    atomicSub(&c.hdr->nQueued, 1u);
                    __threadfence();
                    os.action = ACT_CALL; os.actOwner = c.me; os.actSlot = 8 * os.par.group + os.j; os.actCancel = 0;
                    return RQ_ACTION;
                }
                break;   // a helper took it this instant
            }
            if (st == SL_RUNNING) return RQ_FINDWORK;   // a helper is on it: do something useful meanwhile
            // SL_DONE: consume in the reference's order
            __threadfence();
            const volatile InnerOut* o = &sl->out;
            const int ost = o->status;
            if (ost == 7) { st_vol(&sl->state, SL_FREE); break; }   // abandoned under an older generation (cannot be the call we wait for, but harmless): redo
            if (ost != 0) { os.status = GOICP_SR_OVERFLOW; finish_pair(c, 0); os.phase = OW_NONE; break; }
            if (*reinterpret_cast<const volatile float*>(&sl->pr.optError) != os.optError) { st_vol(&sl->state, SL_FREE); break; }     // made under another incumbent: redo (defensive; invalidation withdraws these)
            if (!os.needLbOnly) {
                os.cnt[4]++; os.cnt[0]++; os.cnt[1] += o->pops; os.cnt[2] += o->subcubes;   // :768
                os.ubChild = o->err;
                if (o->err < os.optError) {                                                  // :771-790
                    os.optError = o->err;
                    for (int k = 0; k < 9; k++) os.optR[k] = (double)*reinterpret_cast<const volatile float*>(&sl->pr.R[k]);
                    const float n0 = o->node[0], n1 = o->node[1], n2 = o->node[2], nw = o->node[3];
                    os.optT[0] = (double)(n0 + nw / 2); os.optT[1] = (double)(n1 + nw / 2); os.optT[2] = (double)(n2 + nw / 2);
                    for (int q = 0; q < 2; q++) {   // updateCompatibilities (:791) + ICP(R, t) (:810) at the new incumbent
                        IcpState& s = c.icp[q];
                        s.pair = os.pair; s.mode = q == 0 ? 2 : 0;
                        for (int k = 0; k < 9; k++) s.R[k] = os.optR[k];
                        for (int k = 0; k < 3; k++) { s.t[k] = os.optT[k]; s.mu_m[k] = 0.0; s.mu_d[k] = 0.0; }
                        s.err = -1.f; s.iter = 0; s.done = 0; s.status = 0; s.error = 0.f; s.incomp = 0; s.fpfh = 0.f; s.compat_pose = 0;
                    }
                    __threadfence();
                    st_vol(&sl->state, SL_FREE);
                    os.needLbOnly = 1; os.phase = OW_AFTER_IMPROVE; os.action = ACT_ICP;
                    return RQ_ACTION;
                }
                if (!o->ran2) { st_vol(&sl->state, SL_FREE); os.needLbOnly = 1; break; }     // (cannot happen: the second call is skipped only on improvement)
                os.cnt[0]++; os.cnt[1] += o->pops2; os.cnt[2] += o->subcubes2;                // :861
                os.lastLb = o->err2;
                if (!(o->err2 >= os.optError)) {                                             // :863-871
                    RNodeD nr = child_of(os.par, os.j); nr.ub = os.ubChild; nr.lb = o->err2;
                    if (os.heapN >= A.rqCap) { os.status = GOICP_SR_OVERFLOW; finish_pair(c, 0); os.phase = OW_NONE; break; }
                    rq_push(c.rq, os.heapN, nr);
                }
            } else {
                os.cnt[0]++; os.cnt[1] += o->pops; os.cnt[2] += o->subcubes;
                os.lastLb = o->err;
                if (!(o->err >= os.optError)) {
                    RNodeD nr = child_of(os.par, os.j); nr.ub = os.ubChild; nr.lb = o->err;
                    if (os.heapN >= A.rqCap) { os.status = GOICP_SR_OVERFLOW; finish_pair(c, 0); os.phase = OW_NONE; break; }
                    rq_push(c.rq, os.heapN, nr);
                }
                os.needLbOnly = 0;
            }
            st_vol(&sl->state, SL_FREE);
            os.j++;
            break;
        }
        case OW_AFTER_IMPROVE: {   // :791-854 (the speculative calls were invalidated by the caller)
            const volatile IcpState* s0 = c.icp; const volatile IcpState* s1 = c.icp + 1;
            if (s0->status != 0 || s1->status != 0) { os.status = GOICP_SR_UNSUPPORTED; finish_pair(c, 0); os.phase = OW_NONE; break; }
            os.optComp = s0->compat_pose;                                                   // :791
            add_event(c, 2, os.optError);
            os.cnt[5]++;
            const float icpErr = s1->error;
            if (icpErr < os.optError) {                                                     // :813-840
                os.optError = icpErr;
                for (int k = 0; k < 9; k++) os.optR[k] = s1->R[k];
                for (int k = 0; k < 3; k++) os.optT[k] = s1->t[k];
                os.optComp = s1->incomp;
                add_event(c, 1, icpErr);
            }
            // :843-853: pop in order into a new queue until the first node with lb >= optError.  Pushing in pop order never sifts
            // (no parent is "less" than a later key), so the new heap array is the popped prefix itself.
            {
                RNodeD* tmp = c.rq + A.rqCap;   // second half of the slab
                int n = os.heapN, m = 0;
                while (n > 0) { RNodeD nd = rq_pop(c.rq, n); if (nd.lb < os.optError) { nd.group = -1; st_node(tmp + m, nd); m++; } else break; }
                for (int k = 0; k < m; k++) st_node(c.rq + k, ld_node(tmp + k));
                os.heapN = m;
            }
            { int g; while ((g = alloc_group(c, 0)) < 0) __nanosleep(200); os.par.group = g; }   // the old group drains (calls of the previous incumbent)
            os.lastSpawnJ = os.j;
            os.phase = OW_CHILD;
            return RQ_SPAWN;
        }
        default: return RQ_ACTION;   // OW_NONE: handled by the scheduler
        }
        if (os.phase == OW_NONE) { os.action = ACT_NONE; return RQ_ACTION; }
    }
}

// How many slot groups the owner should keep attached to queue nodes: none while unclaimed pairs remain (idle CTAs take pairs),
// afterwards the helpers are shared among the owners.  All lanes (warp-uniform result).
__device__ __forceinline__ int spec_target(Cta& c) {
    const SearchArgs& A = c.A;
    int t = 0;
    if (c.lane == 0 && A.specMax > 0 && ld_voli(&A.ctl->nextPair) >= A.npairs) {
        const int owners = max(1, ld_voli(&A.ctl->owners));
        const int helpers = max(0, A.nCtas - owners);
        const int tasks = (helpers + owners - 1) / owners;     // calls in flight per owner that keep every helper busy
        t = min(A.specMax, (tasks + 7) / 8 + (tasks > 0 ? 1 : 0));
    }
    return __shfl_sync(GOICP_FULL, t, 0);
}

// Publish speculative calls: the remaining children of the current node, then the children of the next queue nodes in pop order.
// All lanes of warp 0.
__device__ __noinline__ void spawn_spec(Cta& c) {
    OwnerSh& os = c.os;
    const SearchArgs& A = c.A;
    const int target = spec_target(c);
    if (target <= 0 || os.phase != OW_CHILD) return;
    const int lane = c.lane;
    int published = 0;
    // children after the current one (the current one is the owner's own next call)
    if (lane < 8 && lane > os.j) {
        SearchSlot* sl = c.slots + 8 * os.par.group + lane;
        if (ld_vol(&sl->state) == SL_FREE) {
            const RNodeD ch = child_of(os.par, lane);
            float R[9];
            if (!child_rotation(A.fma != 0, ch, R)) st_vol(&sl->state, SL_SKIP);
            else { fill_request(c, sl, ch, R, false, 0u); __threadfence(); st_vol(&sl->state, SL_QUEUED); published++; }
        }
    }
    // the next nodes in pop order: the queue is a binary heap, so they are reached from the root through a frontier of candidate
    // positions; candidate keys are compared by all lanes at once
    if (lane == 0) { os.ncand = 0; if (os.heapN > 0) { os.cand[0] = 0; os.ncand = 1; } }
    __syncwarp();
    int withGroup = 0;
    for (int it = 0; it < A.specMax + 8 && withGroup < target; it++) {
        const int nc = os.ncand;
        if (nc == 0) break;
        unsigned long long best = ~0ull;
        for (int k = lane; k < nc; k += 32) {
            const RNodeD nd = ld_node(c.rq + os.cand[k]);
            const unsigned long long key = ((unsigned long long)__float_as_uint(nd.lb) << 32) | ((unsigned long long)(unsigned)nd.l << 8) | (unsigned)k;
            best = key < best ? key : best;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(GOICP_FULL, best, o); best = v < best ? v : best; }
        const int k = (int)(best & 0xFFu);
        const int pos = os.cand[k];
        __syncwarp();
        if (lane == 0) {
            os.cand[k] = os.cand[nc - 1]; int n2 = nc - 1;
            if (2 * pos + 1 < os.heapN && n2 < 62) os.cand[n2++] = 2 * pos + 1;
            if (2 * pos + 2 < os.heapN && n2 < 62) os.cand[n2++] = 2 * pos + 2;
            os.ncand = n2;
        }
        __syncwarp();
        RNodeD nd = ld_node(c.rq + pos);
        if ((os.optError - nd.lb) <= os.SSE) break;   // the search ends when this node is popped (:685)
        if (nd.group < 0) {
            int g = -1;
            if (lane == 0) g = alloc_group(c, 2);
            g = __shfl_sync(GOICP_FULL, g, 0);
            if (g < 0) break;
            nd.group = g;
            if (lane == 0) { st_node(c.rq + pos, nd); os.specGroups++; }
            if (lane < 8) {
                SearchSlot* sl = c.slots + 8 * g + lane;
                const RNodeD ch = child_of(nd, lane);
                float R[9];
                if (!child_rotation(A.fma != 0, ch, R)) st_vol(&sl->state, SL_SKIP);
                else { fill_request(c, sl, ch, R, false, __float_as_uint(nd.lb) | 1u); __threadfence(); st_vol(&sl->state, SL_QUEUED); published++; }
            }
        }
        withGroup++;
    }
    published = __reduce_add_sync(GOICP_FULL, published);
    if (lane == 0 && published) atomicAdd(&c.hdr->nQueued, (unsigned)published);
    __syncwarp();
}

// the unclaimed call with the smallest priority in owner `o`'s slots; claims it.  All lanes; returns the slot or -1.
__device__ int claim_from(Cta& c, int o) {
    SearchSlot* sl = c.A.slots + (size_t)o * SR_NSLOT;
    for (int tries = 0; tries < 4; tries++) {
        unsigned long long best = ~0ull;
        for (int s = c.lane; s < SR_NSLOT; s += 32) {
            if (ld_vol(&sl[s].state) == SL_QUEUED) { const unsigned long long key = ((unsigned long long)ld_vol(&sl[s].prio) << 16) | (unsigned)s; best = key < best ? key : best; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { const unsigned long long v = __shfl_xor_sync(GOICP_FULL, best, off); best = v < best ? v : best; }
        if (best == ~0ull) return -1;
        const int s = (int)(best & 0xFFFFu);
        int ok = 0;
        if (c.lane == 0) { ok = atomicCAS(&sl[s].state, (unsigned)SL_QUEUED, (unsigned)SL_RUNNING) == SL_QUEUED; if (ok) { atomicSub(&c.A.hdrs[o].nQueued, 1u); __threadfence(); } }
        ok = __shfl_sync(GOICP_FULL, ok, 0);
        if (ok) return s;
    }
    return -1;
}

// warp 0: decide what the CTA does next
__device__ __noinline__ void schedule(Cta& c) {
    OwnerSh& os = c.os;
    const SearchArgs& A = c.A;
    const int lane = c.lane;
    for (;;) {
        // ---- no pair: claim the next one ----
        if (os.pair < 0 && !os.noMorePairs) {
            int i = 0;
            if (lane == 0) i = atomicAdd(&A.ctl->nextPair, 1);
            i = __shfl_sync(GOICP_FULL, i, 0);
            if (lane == 0) {
                if (i < A.npairs) { os.pair = i; os.phase = OW_START; st_vol(reinterpret_cast<unsigned*>(&c.hdr->pair), (unsigned)i); atomicAdd(&A.ctl->owners, 1); }
                else os.noMorePairs = 1;
            }
            __syncwarp();
        }
        // ---- own pair: advance the search ----
        if (os.pair >= 0) {
            int rq = 0;
            if (lane == 0) rq = owner_serial(c);
            rq = __shfl_sync(GOICP_FULL, rq, 0);
            __syncwarp();
            if (rq == RQ_SPAWN) { spawn_spec(c); continue; }
            if (rq == RQ_ACTION) {
                if (os.phase == OW_AFTER_IMPROVE && os.action == ACT_ICP) invalidate_spec(c);   // the incumbent just improved
                if (os.phase == OW_NONE) {   // the pair is finished
                    invalidate_spec(c);
                    if (lane == 0) { os.pair = -1; st_vol(reinterpret_cast<unsigned*>(&c.hdr->pair), 0xFFFFFFFFu); atomicSub(&A.ctl->owners, 1); __threadfence(); atomicAdd(&A.ctl->pairsDone, 1); }
                    __syncwarp();
                    continue;
                }
                if (os.action != ACT_NONE) return;
                continue;
            }
            // RQ_FINDWORK: the call the search waits for runs on a helper; take the next unclaimed call of this pair instead
            const int s = claim_from(c, c.me);
            if (s >= 0) { if (lane == 0) { os.action = ACT_CALL; os.actOwner = c.me; os.actSlot = s; os.actCancel = 1; } __syncwarp(); return; }
        }
        // ---- nothing of its own to run: help another owner ----
        {
            int found = -1;
            const int start = (c.me * 7 + 1) % A.nCtas;
            for (int base = 0; base < A.nCtas && found < 0; base += 32) {
                const int o = (start + base + lane) % A.nCtas;
                const bool has = (base + lane) < A.nCtas && o != c.me && (int)ld_vol(&A.hdrs[o].nQueued) > 0;
                const unsigned m = __ballot_sync(GOICP_FULL, has);
                if (m) found = __shfl_sync(GOICP_FULL, o, __ffs(m) - 1);
            }
            if (found >= 0) {
                const int s = claim_from(c, found);
                if (s >= 0) {
                    if (lane == 0) { os.action = ACT_CALL; os.actOwner = found; os.actSlot = s; os.actCancel = 1; }
                    __syncwarp();
                    return;
                }
                continue;
            }
        }
        // ---- idle ----
        if (os.pair < 0) {
            int done = 0;
            if (lane == 0) done = ld_voli(&A.ctl->pairsDone) >= A.npairs;
            done = __shfl_sync(GOICP_FULL, done, 0);
            if (done) { if (lane == 0) os.action = ACT_EXIT; __syncwarp(); return; }
        }
        __nanosleep(os.pair >= 0 ? 200 : 1000);
        __syncwarp();
    }
}

template <bool EXACT, bool SMEM, bool GS, bool CT>
__global__ void __launch_bounds__(BNB_MAX_THREADS, GOICP_BNB_MIN_CTAS)
search_kernel(const SearchArgs A) {
    extern __shared__ float4 dyn_smem4[];
    __shared__ InnerProb s_pr;
    __shared__ InnerOut s_out;
    __shared__ unsigned long long s_gbar;
    __shared__ OwnerSh os;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* icpTile;
    if constexpr (SMEM) icpTile = reinterpret_cast<float*>(dyn_smem4); else { __shared__ float s_tile[3 * NN_TILE]; icpTile = s_tile; }
    CallCtx cx;
    unsigned long long* dstat = reinterpret_cast<unsigned long long*>(A.genCounter) + 1;
    if (tid == 0) {
        os.pair = -1; os.phase = OW_NONE; os.noMorePairs = 0; os.lastSpawnJ = -1; os.specGroups = 0; os.action = ACT_NONE;
        for (int g = 0; g < SR_NGROUP; g++) os.grpUse[g] = 0;
        if (GS) mbar_init(&s_gbar, 1);
    }
    __syncthreads();
    Cta c{A, os, A.slots + (size_t)blockIdx.x * SR_NSLOT, A.hdrs + blockIdx.x, reinterpret_cast<RNodeD*>(A.rq) + (size_t)blockIdx.x * 2 * A.rqCap, A.icp + 2 * (size_t)blockIdx.x, (int)blockIdx.x, lane};
    const long long tStart = clock64();
    long long tIdle = 0;

    for (;;) {
        __syncthreads();
        if (warp == 0) { const long long t0 = clock64(); schedule(c); if (lane == 0) tIdle += clock64() - t0; }
        __syncthreads();
        const int act = os.action;
        if (act == ACT_EXIT) break;
        if (act == ACT_ICP) {
            icp_fused_body(A.pairs, c.icp, icpTile);
            __syncthreads();
            icp_fused_body(A.pairs, c.icp + 1, icpTile);
            if (tid == 0) atomicAdd(dstat + 6, 2ull);
            continue;
        }
        // ACT_CALL: one request out of slot (actOwner, actSlot)
        SearchSlot* sl = A.slots + (size_t)os.actOwner * SR_NSLOT + os.actSlot;
        if (tid < (int)(sizeof(InnerProb) / 4)) reinterpret_cast<unsigned*>(&s_pr)[tid] = reinterpret_cast<const volatile unsigned*>(&sl->pr)[tid];
        cx.cancelWord = os.actCancel ? &A.hdrs[os.actOwner].gen : nullptr;
        cx.cancelGen = os.actCancel ? ld_vol(&sl->gen) : 0u;
        inner_call<EXACT, SMEM, GS, CT, true>(A.pairs, s_pr, s_out, s_gbar, cx, A.heaps, A.heapCap, A.gscratch, A.gstride, A.NdP, A.NdQ, A.useSmem, A.memo, A.memoCap, A.genCounter, A.gridOff, A.S3p);
        if (warp == 0) {
            __syncwarp();
            if (lane < 16) reinterpret_cast<volatile unsigned*>(&sl->out)[lane] = reinterpret_cast<const unsigned*>(&s_out)[lane];
            __threadfence();
            __syncwarp();
            if (lane == 0) st_vol(&sl->state, SL_DONE);
        }
    }
    if (tid == 0) { atomicAdd(dstat + 4, (unsigned long long)tIdle); atomicAdd(dstat + 7, (unsigned long long)(clock64() - tStart)); }
}

typedef void (*search_kernel_t)(const SearchArgs);
template <bool SMEM, bool GS>
search_kernel_t search_sel(int exact, int ct) {
    if (exact) return ct ? search_kernel<true, SMEM, GS, true> : search_kernel<true, SMEM, GS, false>;
    return ct ? search_kernel<false, SMEM, GS, true> : search_kernel<false, SMEM, GS, false>;
}
search_kernel_t search_fn(int exact, int smem, int ct) {
    if (smem == 2) return search_sel<true, true>(exact, ct);
    if (smem == 1) return search_sel<true, false>(exact, ct);
    return search_sel<false, false>(exact, ct);
}
int g_search_attr[16] = {0};
cudaError_t search_attr(int exact, int smem, int ct) {
    const int k = (exact ? 1 : 0) + 2 * smem + 8 * (ct ? 1 : 0);
    if (g_search_attr[k]) return cudaSuccess;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, search_fn(exact, smem, ct));
    if (e != cudaSuccess) return e;
    const int maxDyn = 227 * 1024 - (int)fa.sharedSizeBytes - 1024;
    e = cudaFuncSetAttribute(search_fn(exact, smem, ct), cudaFuncAttributeMaxDynamicSharedMemorySize, maxDyn);
    if (e == cudaSuccess) g_search_attr[k] = 1;
    return e;
}

}  // namespace

size_t goicp_search_slot_bytes() { return sizeof(SearchSlot); }
size_t goicp_search_hdr_bytes() { return sizeof(OwnerHdr); }
size_t goicp_search_rnode_bytes() { return sizeof(RNodeD); }

int goicp_search_occupancy(size_t smemBytes, int exact, int threads, int useSmem, int ct) {
    if (search_attr(exact, useSmem, ct) != cudaSuccess) return 1;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, search_fn(exact, useSmem, ct), threads, useSmem ? smemBytes : 0) != cudaSuccess || n < 1) n = 1;
    return n;
}
cudaError_t goicp_launch_search(const SearchArgs& A, int ctas, int threads, size_t smemBytes, int exact, int ct, cudaStream_t st) {
    cudaError_t e = search_attr(exact, A.useSmem & 3, ct);
    if (e != cudaSuccess) return e;
    search_fn(exact, A.useSmem & 3, ct)<<<ctas, threads, (A.useSmem & 3) ? smemBytes : 0, st>>>(A);
    return cudaGetLastError();
}
cudaError_t goicp_preload_search() {
    cudaFuncAttributes a; cudaError_t e;
    for (int exact = 0; exact < 2; exact++) for (int smem = 0; smem < 3; smem++) for (int ct = 0; ct < 2; ct++)
        if ((e = cudaFuncGetAttributes(&a, search_fn(exact, smem, ct))) != cudaSuccess) return e;
    return cudaSuccess;
}
