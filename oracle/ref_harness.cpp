// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// C-ABI harness around the UNMODIFIED reference sources (compiled where they lie under
// /root/reference by oracle/Makefile; nothing from the reference is copied into this repo).
// It lets the python tests and bench.py's cpu_baseline / --impl reference arm drive the reference's
// own GoICP / DT3D / ICP3D / Transformation classes on in-memory arrays, so that the C restatement
// (oracle/goicp_oracle.c) and the CUDA product can be pinned against the real thing.
//
// Reference entry points exercised (file:line under /root/reference):
//   GoICP::BuildDT jly_goicp.cpp:79        DT3D::Build jly_3ddt.cpp:897   DT3D::Distance jly_3ddt.cpp:1139
//   GoICP::Initialize jly_goicp.cpp:180    GoICP::InnerBnB jly_goicp.cpp:286
//   GoICP::OuterBnB jly_goicp.cpp:582      GoICP::ICP jly_goicp.cpp:102   GoICP::Register jly_goicp.cpp:878
//   Transformation::normalizeMolCloud transformation.cpp:311 ... computeRMSD transformation.cpp:453
//
// Every std header the reference pulls in is included BEFORE "#define private public" so that only the
// reference's own classes are opened up.
#include <queue>
#include <map>
#include <vector>
#include <iostream>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <string>
#include <random>
#include <iterator>
#include <limits>
#include <stdexcept>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <thread>
#define private public
#include "jly_goicp.h"
#undef private
#include "jly_sorting.hpp"   // intro_select (template header; jly_goicp.cpp:41 includes it the same way)
#include "ref_counters.h"

long long goicp_ref_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};

// The reference mallocs arrays of POINT3D (which holds a std::vector) and later `delete`s them
// (jly_goicp.cpp:210-211,271-272): the destructor of element 0 then runs on uninitialised bytes.
// In its own short-lived process the heap happens to be zero there; inside a long-lived python process it is
// not ("free(): invalid size").  Without touching the sources, the link step (-Wl,--wrap=malloc, oracle/Makefile)
// routes every malloc made by the objects of THIS library to a zeroing allocator, which is the state the
// stand-alone binary observes.
extern "C" void* __wrap_malloc(size_t n) { return calloc(1, n); }

extern "C" {

// Same field order as include/goicp_b200.h:goicp_params (kept in sync by tests/test_abi.py).
struct ref_params {
    float MSEThresh;
    float rotMinX, rotMinY, rotMinZ, rotWidth;
    float transMinX, transMinY, transMinZ, transWidth;
    float trimFraction;
    float regularization, regularizationNeighbors, regularizationFPFH;
    int cfpfh, norm, ponderation;
    int distTransSize;
    double distTransExpandFactor;
};

struct ref_result {
    double R[9];
    double t[3];
    float optError;
    int optComp;
    long long counters[8];  // inner calls, trans pops, trans subcubes, rot pops, rot cubes, icp calls, 0, 0
    double seconds_dt, seconds_register;
};

struct ref_handle {
    GoICP* g;
    int Nm, Nd;
    bool dt_built, initialized;
    std::string trace;
};

static POINT3D* make_cloud(const float* xyz, const int* c, const float* fpfh, int n) {
    POINT3D* p = (POINT3D*)calloc(n, sizeof(POINT3D));
    for (int i = 0; i < n; i++) {
        new (&p[i]) POINT3D();
        p[i].x = xyz[3 * i]; p[i].y = xyz[3 * i + 1]; p[i].z = xyz[3 * i + 2];
        p[i].c = c ? c[i] : 0;
        p[i].neighbors = 0; p[i].density = 0;
        if (fpfh) p[i].cfpfh.assign(fpfh + 41 * (size_t)i, fpfh + 41 * (size_t)(i + 1));
        else p[i].cfpfh.assign(41, 0.0f);
    }
    return p;
}

// Captures (and silences) the reference's std::cout chatter for the duration of a call.
struct CoutCapture {
    std::ostringstream ss; std::streambuf* old;
    CoutCapture() { old = std::cout.rdbuf(ss.rdbuf()); }
    ~CoutCapture() { std::cout.rdbuf(old); }
};

void* ref_create(const float* mxyz, const int* mc, const float* mfpfh, int Nm,
                 const float* dxyz, const int* dc, const float* dfpfh, int Nd, const ref_params* p) {
    ref_handle* h = new ref_handle();
    GoICP* g = new GoICP();
    h->g = g; h->Nm = Nm; h->Nd = Nd; h->dt_built = false; h->initialized = false;
    g->MSEThresh = p->MSEThresh;
    g->initNodeRot.a = p->rotMinX; g->initNodeRot.b = p->rotMinY; g->initNodeRot.c = p->rotMinZ; g->initNodeRot.w = p->rotWidth;
    g->initNodeTrans.x = p->transMinX; g->initNodeTrans.y = p->transMinY; g->initNodeTrans.z = p->transMinZ; g->initNodeTrans.w = p->transWidth;
    g->trimFraction = p->trimFraction;
    if (g->trimFraction < 0.001) g->doTrim = false;   // jly_main.cpp:259
    g->regularization = p->regularization;
    g->regularizationNeighbors = p->regularizationNeighbors;
    g->regularizationFPFH = p->regularizationFPFH;
    g->cfpfh = p->cfpfh; g->norm = p->norm; g->ponderation = p->ponderation;
    g->dt.SIZE = p->distTransSize; g->dt.expandFactor = p->distTransExpandFactor;
    g->optComp = 0;
    g->pModel = make_cloud(mxyz, mc, mfpfh, Nm); g->Nm = Nm;
    g->pData = make_cloud(dxyz, dc, dfpfh, Nd); g->Nd = Nd;
    return h;
}

void ref_destroy(void* hv) {
    ref_handle* h = (ref_handle*)hv;
    // The reference never frees cellPoints/emptyCells; release what we can reach.
    GoICP* g = h->g;
    if (h->dt_built) {
        int S = g->dt.SIZE;
        for (int i = 0; i < S; i++) {
            for (int j = 0; j < S; j++) { delete[] g->dt.emptyCells[i][j]; delete[] g->dt.cellPoints[i][j]; }
            delete[] g->dt.emptyCells[i]; delete[] g->dt.cellPoints[i];
        }
        delete[] g->dt.emptyCells; delete[] g->dt.cellPoints;
    }
    for (int i = 0; i < h->Nm; i++) g->pModel[i].~POINT3D();
    for (int i = 0; i < h->Nd; i++) g->pData[i].~POINT3D();
    free(g->pModel); free(g->pData);
    delete g;
    delete h;
}

void ref_set_nd(void* hv, int nd) { ((ref_handle*)hv)->g->Nd = nd; }   // jly_main.cpp:114-117

double ref_build_dt(void* hv) {
    ref_handle* h = (ref_handle*)hv; CoutCapture cap;
    clock_t t0 = clock();
    h->g->BuildDT();
    h->dt_built = true;
    return (double)(clock() - t0) / CLOCKS_PER_SEC;
}

// out[0..7] = xMin,xMax,yMin,yMax,zMin,zMax,scale,SIZE
void ref_dt_info(void* hv, double* out) {
    DT3D& d = ((ref_handle*)hv)->g->dt;
    out[0] = d.xMin; out[1] = d.xMax; out[2] = d.yMin; out[3] = d.yMax; out[4] = d.zMin; out[5] = d.zMax;
    out[6] = d.scale; out[7] = d.SIZE;
}

// dist[S^3] (z-major: ((z*S)+y)*S+x), off[S^3*3] = (v,h,d) offsets, nearest[S^3*3] = emptyCells (cx,cy,cz),
// cellc[S^3] = cellPoints.c (after assignCellColor). Any pointer may be NULL.
void ref_dt_download(void* hv, float* dist, short* off, int* nearest, int* cellc) {
    DT3D& d = ((ref_handle*)hv)->g->dt; int S = d.SIZE;
    for (int z = 0; z < S; z++) for (int y = 0; y < S; y++) for (int x = 0; x < S; x++) {
        size_t i = ((size_t)z * S + y) * S + x;
        if (dist) dist[i] = d.A.data[z][y][x].distance;
        if (off) { off[3 * i] = d.A.data[z][y][x].v; off[3 * i + 1] = d.A.data[z][y][x].h; off[3 * i + 2] = d.A.data[z][y][x].d; }
        if (nearest) { nearest[3 * i] = d.emptyCells[z][y][x].cx; nearest[3 * i + 1] = d.emptyCells[z][y][x].cy; nearest[3 * i + 2] = d.emptyCells[z][y][x].cz; }
        if (cellc) cellc[i] = d.cellPoints[z][y][x].c;
    }
}

void ref_dt_distance(void* hv, const double* xyz, int n, float* out, int* cell) {
    DT3D& d = ((ref_handle*)hv)->g->dt;
    for (int i = 0; i < n; i++) {
        int cx, cy, cz;
        out[i] = d.Distance(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], cx, cy, cz);
        if (cell) { cell[3 * i] = cx; cell[3 * i + 1] = cy; cell[3 * i + 2] = cz; }
    }
}

void ref_initialize(void* hv) {
    ref_handle* h = (ref_handle*)hv; CoutCapture cap;
    h->g->Initialize(); h->initialized = true;
}

void ref_get_weights(void* hv, float* w) { GoICP* g = ((ref_handle*)hv)->g; memcpy(w, g->weights, sizeof(float) * g->Nd); }
void ref_get_maxrotdis(void* hv, float* out /*20*Nd*/) {
    GoICP* g = ((ref_handle*)hv)->g;
    for (int l = 0; l < MAXROTLEVEL; l++) memcpy(out + (size_t)l * g->Nd, g->maxRotDis[l], sizeof(float) * g->Nd);
}
float ref_get_ssethresh(void* hv) { return ((ref_handle*)hv)->g->SSEThresh; }
int ref_get_inliernum(void* hv) { return ((ref_handle*)hv)->g->inlierNum; }

// One InnerBnB call exactly as OuterBnB makes it (jly_goicp.cpp:750-768,861): R == NULL means the
// memcpy (t==0) branch; level < 0 means "upper bound" (maxRotDisL == NULL).
float ref_inner_bnb(void* hv, const float* R, int level, float optError, float* tnode /*x,y,z,w or NULL*/) {
    ref_handle* h = (ref_handle*)hv; GoICP* g = h->g; CoutCapture cap;
    for (int i = 0; i < g->Nd; i++) {
        POINT3D& p = g->pData[i];
        if (R) {
            g->pDataTemp[i].x = R[0] * p.x + R[1] * p.y + R[2] * p.z;
            g->pDataTemp[i].y = R[3] * p.x + R[4] * p.y + R[5] * p.z;
            g->pDataTemp[i].z = R[6] * p.x + R[7] * p.y + R[8] * p.z;
        } else { g->pDataTemp[i].x = p.x; g->pDataTemp[i].y = p.y; g->pDataTemp[i].z = p.z; }
    }
    g->optError = optError;
    TRANSNODE tn; tn.x = tn.y = tn.z = tn.w = 0; tn.ub = tn.lb = 0;
    float e = g->InnerBnB(level >= 0 ? g->maxRotDis[level] : NULL, tnode ? &tn : NULL);
    if (tnode) { tnode[0] = tn.x; tnode[1] = tn.y; tnode[2] = tn.z; tnode[3] = tn.w; }
    return e;
}

// The residual row of one child translation cube as InnerBnB's loop fills minDis (jly_goicp.cpp:331-382), then the REFERENCE's
// own intro_select (jly_sorting.hpp:229, called as at jly_goicp.cpp:389).  firstk[n x inlierNum] receives the values the
// reference's trimmed sums run over (the first inlierNum entries after the partial partition), resid[n x Nd] the row before it.
void ref_eval_inclusion(void* hv, const float* R, int level, const float* tcube, int n, float* firstk, float* resid) {
    ref_handle* h = (ref_handle*)hv; GoICP* g = h->g; CoutCapture cap;
    const int Nd = g->Nd;
    for (int i = 0; i < Nd; i++) {
        POINT3D& p = g->pData[i];
        g->pDataTemp[i].x = R[0] * p.x + R[1] * p.y + R[2] * p.z;
        g->pDataTemp[i].y = R[3] * p.x + R[4] * p.y + R[5] * p.z;
        g->pDataTemp[i].z = R[6] * p.x + R[7] * p.y + R[8] * p.z;
    }
    float* maxRotDisL = level >= 0 ? g->maxRotDis[level] : NULL;
    for (int q = 0; q < n; q++) {
        const float w = tcube[4 * q + 3];
        float transX = tcube[4 * q] + w / 2, transY = tcube[4 * q + 1] + w / 2, transZ = tcube[4 * q + 2] + w / 2;
        int cx, cy, cz;
        for (int i = 0; i < Nd; i++) {
            g->minDis[i] = g->weights[i] * g->dt.Distance(g->pDataTemp[i].x + transX, g->pDataTemp[i].y + transY, g->pDataTemp[i].z + transZ, cx, cy, cz);
            if (maxRotDisL) g->minDis[i] -= maxRotDisL[i];
            if (g->minDis[i] < 0) g->minDis[i] = 0;
            if (resid) resid[(size_t)q * Nd + i] = g->minDis[i];
        }
        if (g->doTrim) intro_select(g->minDis, 0, Nd - 1, g->inlierNum - 1);
        memcpy(firstk + (size_t)q * g->inlierNum, g->minDis, sizeof(float) * g->inlierNum);
    }
}

// GoICP::ICP from a given pose; R,t in/out (row-major doubles), corr[Nd] = id_model per data point.
float ref_icp(void* hv, double* R, double* t, int* corr) {
    ref_handle* h = (ref_handle*)hv; GoICP* g = h->g; CoutCapture cap;
    Matrix Rm(3, 3, R), tm(3, 1, t);
    float e = g->ICP(Rm, tm);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[3 * i + j] = Rm.val[i][j]; t[i] = tm.val[i][0]; }
    if (corr) for (int i = 0; i < g->Nd; i++) corr[g->icp3d.points[i].id_data] = g->icp3d.points[i].id_model;
    free(g->icp3d.points);  // the reference leaks this every Run
    return e;
}

// BuildDT (unless already built) + Register.  trace receives the captured stdout (Error* lines etc.).
void ref_register(void* hv, int nd_downsampled, ref_result* out, char* trace, int trace_cap) {
    ref_handle* h = (ref_handle*)hv; GoICP* g = h->g; CoutCapture cap;
    for (int i = 0; i < 8; i++) goicp_ref_cnt[i] = 0;
    struct timespec a, b, c;
    clock_gettime(CLOCK_MONOTONIC, &a);
    if (!h->dt_built) { g->BuildDT(); h->dt_built = true; }
    if (nd_downsampled > 0) g->Nd = nd_downsampled;
    clock_gettime(CLOCK_MONOTONIC, &b);
    g->Register();
    clock_gettime(CLOCK_MONOTONIC, &c);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) out->R[3 * i + j] = g->optR.val[i][j]; out->t[i] = g->optT.val[i][0]; }
    out->optError = g->optError; out->optComp = g->optComp;
    for (int i = 0; i < 8; i++) out->counters[i] = goicp_ref_cnt[i];
    out->seconds_dt = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
    out->seconds_register = (c.tv_sec - b.tv_sec) + 1e-9 * (c.tv_nsec - b.tv_nsec);
    if (trace && trace_cap > 0) {
        std::string s = cap.ss.str();
        size_t n = s.size() < (size_t)trace_cap - 1 ? s.size() : (size_t)trace_cap - 1;
        memcpy(trace, s.data(), n); trace[n] = 0;
    }
}

// ---- Transformation (transformation.cpp) on in-memory clouds ----
// normalizeMolCloud :311 -- centres xyz in place, returns max norm, writes the mean.
double ref_normalize(double* xyz, int n, double* mean) {
    Transformation t; std::vector<point4D> cl(n);
    for (int i = 0; i < n; i++) { cl[i].x = xyz[3 * i]; cl[i].y = xyz[3 * i + 1]; cl[i].z = xyz[3 * i + 2]; cl[i].c = 0; }
    double s = t.normalizeMolCloud(cl, mean[0], mean[1], mean[2]);
    for (int i = 0; i < n; i++) { xyz[3 * i] = cl[i].x; xyz[3 * i + 1] = cl[i].y; xyz[3 * i + 2] = cl[i].z; }
    return s;
}
// scaleCloud :355
void ref_scale(double* xyz, int n, double scale) {
    Transformation t; std::vector<point4D> cl(n);
    for (int i = 0; i < n; i++) { cl[i].x = xyz[3 * i]; cl[i].y = xyz[3 * i + 1]; cl[i].z = xyz[3 * i + 2]; }
    t.scaleCloud(cl, scale);
    for (int i = 0; i < n; i++) { xyz[3 * i] = cl[i].x; xyz[3 * i + 1] = cl[i].y; xyz[3 * i + 2] = cl[i].z; }
}
// readMolFile :282 -- returns count; fills up to cap points (xyz doubles, colour codes).
int ref_read_mol2(const char* path, double* xyz, int* c, int cap) {
    Transformation t; std::ifstream f(path); if (!f.is_open()) return -1;
    std::vector<point4D> cl = t.readMolFile(f);
    int n = (int)cl.size();
    for (int i = 0; i < n && i < cap; i++) { xyz[3 * i] = cl[i].x; xyz[3 * i + 1] = cl[i].y; xyz[3 * i + 2] = cl[i].z; c[i] = cl[i].c; }
    return n;
}
// writeNormalizedMolCloudFile :340 (6 significant digits text round trip)
int ref_write_xyz(const char* path, const double* xyz, const int* c, int n) {
    Transformation t; std::vector<point4D> cl(n);
    for (int i = 0; i < n; i++) { cl[i].x = xyz[3 * i]; cl[i].y = xyz[3 * i + 1]; cl[i].z = xyz[3 * i + 2]; cl[i].c = c[i]; }
    std::ofstream f(path); if (!f.is_open()) return -1;
    t.writeNormalizedMolCloudFile(f, cl);
    return 0;
}
// rescaleCloud :403 -- writes the "_rescaled" file
int ref_rescale(const char* path, double scale, const double* meanT, const double* meanS, const double* R, const double* tr,
                double time, double error) {
    Transformation t; std::vector<point4D> cl; std::ofstream f(path); if (!f.is_open()) return -1;
    double rot[3][3]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rot[i][j] = R[3 * i + j];
    t.rescaleCloud(f, cl, scale, meanT[0], meanT[1], meanT[2], time, error, meanS[0], meanS[1], meanS[2], rot, tr[0], tr[1], tr[2]);
    return 0;
}
// computeRMSD :453
float ref_rmsd(const char* aligned, const char* rot) {
    Transformation t; std::ifstream a(aligned), r(rot); if (!a.is_open() || !r.is_open()) return -1;
    return t.computeRMSD(a, r);
}
// applyTransformationProtein :469 reads "cavitiesR/similar<pair>.txt" relative to cwd.
int ref_apply_protein(const char* out_path, const char* protein_path, int pair) {
    Transformation t; std::ofstream o(out_path); if (!o.is_open()) return -1;
    t.applyTransformationProtein(o, protein_path, pair);
    return 0;
}

// Multi-threaded sweep for bench.py --impl reference: each worker registers whole pairs (the reference is
// single-threaded per pair; pair-level parallelism is the only kind it admits).  NOT thread-safe
// inside the reference (file-static scratch in matrix.cpp, global cout), so workers are PROCESSES: see
// oracle/ref_pool.py.  This symbol only reports how the library was built.
const char* ref_build_info() { return "reference sources compiled in place; DT min-init fix + counters applied to a temp copy at build time"; }

}  // extern "C"
