/* TEST INFRASTRUCTURE ONLY -- see goicp_oracle.h.  Plain-C restatement of the reference hot path.
 * Compile with -ffp-contract=off on x86-64 (no FMA), as the reference build does.
 * All citations are file:line under /root/reference. */
#define _POSIX_C_SOURCE 200809L
#include "goicp_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_PI 3.1415926536   /* jly_goicp.h:44 */
#define ORC_SQRT3 1.732050808 /* jly_goicp.h:45 */
#define ORC_MAXROTLEVEL 20    /* jly_goicp.h:95 */
#define ORC_INFTY 32767       /* jly_3ddt.h:21 */
#define ORC_ROUND(x) ((int)((x) + 0.5)) /* jly_3ddt.cpp:30, jly_goicp.h:96: truncation toward zero */

/* colour codes of the `properties` enum, transformation.hpp:36; the identity compatibility map of
 * jly_goicp.cpp:66-73 has exactly these eight keys (C=1 is NOT a key). */
static const int ORC_KNOWN_PROPS[8] = {8204959, 30894, 15219528, 15231913, 4646984, 16741671, 7566712, 0};

typedef struct { int16_t v, h, d; float dist; } vox_t; /* DEucl3D, jly_3ddt.h:31 */
typedef struct { float a, b, c, w, ub, lb; int l; } node_t; /* ROTNODE / TRANSNODE, jly_goicp.h:59-87 */

typedef struct { node_t* v; int n, cap; } heap_t;

struct orc_ctx {
    orc_params p;
    int Nm, Nd, NdAll;
    float *mx, *my, *mz; int* mc; float* mf; /* model (target) */
    float *dx, *dy, *dz; int* dc; float* df; /* data (source) */
    int doTrim;
    /* DT */
    int S; double scale, xMin, xMax, yMin, yMax, zMin, zMax;
    vox_t* A;        /* S^3, index (z*S+y)*S+x */
    int* nearest;    /* S^3*3 : emptyCells cx,cy,cz */
    int* cellc;      /* S^3 : CELL.c  (-2 empty, -1 mixed, else uniform colour) */
    int *cell_start, *cell_pts; /* CSR of cellPoints[].points in insertion (index) order */
    int *nbD, *nbM;             /* POINT3D.neighbors of the data / model points (assignNeighbors, jly_goicp.cpp:1213-1248) */
    int dt_built;
    /* Initialize */
    float *normData, *minDis, *maxRotDis, *weights;
    float *tx, *ty, *tz;     /* pDataTemp */
    int inlierNum; float SSEThresh;
    int initialized;
    /* state */
    float optError; double optR[9], optT[3]; int optComp;
    int *icp_model, *opt_model; /* correspondences: id_model per data index (points / optPoints) */
    double icp_mu_m[3], icp_mu_d[3];
    long long cnt[8];
    /* corner memo (storedCompatibilities / storedFPFH, jly_goicp.cpp:304-305) */
    struct memo_e { uint32_t kx, ky, kz; int gen; int comp; float fpfh; int has; }* memo; int memo_cap, memo_gen, memo_fill;
    char* trace; int trace_len, trace_cap;
};

/* ------------------------------------------------------------------------------------------------
 * std::priority_queue<NODE> as libstdc++ implements it (push_heap / pop_heap), so that ties between
 * nodes with equal (lb,w) pop in the reference's order.  operator< : jly_goicp.h:64-71,79-86. */
static inline int node_less(const node_t* n1, const node_t* n2) {
    if (n1->lb != n2->lb) return n1->lb > n2->lb;
    return n1->w < n2->w;
}
static void heap_push(heap_t* h, node_t val) {
    if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 64; h->v = (node_t*)realloc(h->v, sizeof(node_t) * h->cap); }
    int hole = h->n++;
    int parent = (hole - 1) / 2;
    while (hole > 0 && node_less(&h->v[parent], &val)) { h->v[hole] = h->v[parent]; hole = parent; parent = (hole - 1) / 2; }
    h->v[hole] = val;
}
static node_t heap_pop(heap_t* h) {
    node_t top = h->v[0];
    int len = --h->n; /* length of the heap that remains; value to re-insert is old v[len] */
    if (len > 0) {
        node_t val = h->v[len];
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (node_less(&h->v[child], &h->v[child - 1])) child--;
            h->v[hole] = h->v[child]; hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); h->v[hole] = h->v[child - 1]; hole = child - 1; }
        int parent = (hole - 1) / 2;
        while (hole > 0 && node_less(&h->v[parent], &val)) { h->v[hole] = h->v[parent]; hole = parent; parent = (hole - 1) / 2; }
        h->v[hole] = val;
    }
    return top;
}

static void tracef(orc_ctx* c, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
#include <stdarg.h>
static void tracef(orc_ctx* c, const char* fmt, ...) {
    char buf[256]; va_list ap; va_start(ap, fmt); int n = vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (n < 0) return;
    if (c->trace_len + n + 1 > c->trace_cap) { c->trace_cap = 2 * (c->trace_cap + n + 64); c->trace = (char*)realloc(c->trace, c->trace_cap); }
    memcpy(c->trace + c->trace_len, buf, n + 1); c->trace_len += n;
}

/* ------------------------------------------------------------------------------------------------ */
orc_ctx* orc_create(const float* mxyz, const int* mc, const float* mfpfh, int Nm,
                    const float* dxyz, const int* dc, const float* dfpfh, int Nd, const orc_params* p) {
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    c->p = *p; c->Nm = Nm; c->Nd = Nd; c->NdAll = Nd;
    c->doTrim = !(p->trimFraction < 0.001); /* GoICP() sets true :54; readConfig clears it jly_main.cpp:259 */
    c->mx = malloc(sizeof(float) * Nm); c->my = malloc(sizeof(float) * Nm); c->mz = malloc(sizeof(float) * Nm);
    c->mc = calloc(Nm, sizeof(int)); c->mf = calloc((size_t)Nm * 41, sizeof(float));
    c->dx = malloc(sizeof(float) * Nd); c->dy = malloc(sizeof(float) * Nd); c->dz = malloc(sizeof(float) * Nd);
    c->dc = calloc(Nd, sizeof(int)); c->df = calloc((size_t)Nd * 41, sizeof(float));
    for (int i = 0; i < Nm; i++) { c->mx[i] = mxyz[3 * i]; c->my[i] = mxyz[3 * i + 1]; c->mz[i] = mxyz[3 * i + 2]; if (mc) c->mc[i] = mc[i]; }
    for (int i = 0; i < Nd; i++) { c->dx[i] = dxyz[3 * i]; c->dy[i] = dxyz[3 * i + 1]; c->dz[i] = dxyz[3 * i + 2]; if (dc) c->dc[i] = dc[i]; }
    if (mfpfh) memcpy(c->mf, mfpfh, sizeof(float) * 41 * (size_t)Nm);
    if (dfpfh) memcpy(c->df, dfpfh, sizeof(float) * 41 * (size_t)Nd);
    c->S = p->distTransSize;
    for (int i = 0; i < 9; i++) c->optR[i] = (i % 4 == 0);
    return c;
}
static void free_init(orc_ctx* c) {
    free(c->normData); free(c->minDis); free(c->maxRotDis); free(c->weights); free(c->tx); free(c->ty); free(c->tz);
    free(c->icp_model); free(c->opt_model);
    c->normData = c->minDis = c->maxRotDis = c->weights = c->tx = c->ty = c->tz = NULL; c->icp_model = c->opt_model = NULL;
    c->initialized = 0;
}
void orc_destroy(orc_ctx* c) {
    free_init(c);
    free(c->mx); free(c->my); free(c->mz); free(c->mc); free(c->mf);
    free(c->dx); free(c->dy); free(c->dz); free(c->dc); free(c->df);
    free(c->A); free(c->nearest); free(c->cellc); free(c->cell_start); free(c->cell_pts); free(c->memo); free(c->trace); free(c->nbD); free(c->nbM);
    free(c);
}
void orc_set_nd(orc_ctx* c, int nd) { c->Nd = nd; }

/* ================================================================================================
 * DT3D::Build -- jly_3ddt.cpp:897-1137.  The EDT is the reference's sequential 8SED-style vector
 * propagation (DEuclidean :716-750 with the six mask functions :57-712), restated table-driven.
 * Each mask entry: bounds condition bits, source offset, (v,h,d) increments.  The two copy-paste slips
 * (MINforwardDE3 mask[1] :88-93 and MINbackwardDE1 mask[10] :459-464 read [z+1][y][x]) are kept.
 * `min` starts as (infty,infty,infty,infty): the 6-line initialisation fix of SURVEY.md section 0.4. */
enum { ZM = 1, ZP = 2, YM = 4, YP = 8, XM = 16, XP = 32 };
typedef struct { int cond; int dz, dy, dx; int iv, ih, id; } mask_e;

static const mask_e M_FWD1[14] = { /* MINforwardDE1 :508-712 */
    {ZM | YM | XM, -1, -1, -1, 1, 1, 1}, {YM | XM, 0, -1, -1, 1, 1, 0}, {ZP | YM | XM, 1, -1, -1, 1, 1, 1},
    {ZM | XM, -1, 0, -1, 1, 0, 1},       {XM, 0, 0, -1, 1, 0, 0},       {XM | ZP, 1, 0, -1, 1, 0, 1},
    {XM | ZM | YP, -1, 1, -1, 1, 1, 1},  {XM | YP, 0, 1, -1, 1, 1, 0},  {XM | YP | ZP, 1, 1, -1, 1, 1, 1},
    {ZM | YM, -1, -1, 0, 0, 1, 1},       {YM, 0, -1, 0, 0, 1, 0},       {ZP | YM, 1, -1, 0, 0, 1, 1},
    {0, 0, 0, 0, 0, 0, 0},               {ZM, -1, 0, 0, 0, 0, 1}};
static const mask_e M_BWD1[14] = { /* MINbackwardDE1 :301-506 */
    {ZM | YM | XP, -1, -1, 1, 1, 1, 1},  {YM | XP, 0, -1, 1, 1, 1, 0},  {ZP | YM | XP, 1, -1, 1, 1, 1, 1},
    {ZM | XP, -1, 0, 1, 1, 0, 1},        {XP, 0, 0, 1, 1, 0, 0},        {XP | ZP, 1, 0, 1, 1, 0, 1},
    {XP | ZM | YP, -1, 1, 1, 1, 1, 1},   {XP | YP, 0, 1, 1, 1, 1, 0},   {XP | YP | ZP, 1, 1, 1, 1, 1, 1},
    {ZP, 1, 0, 0, 0, 0, 1},              {YP | ZP, 1, 0, 0, 0, 1, 1} /* sic: reads [z+1][y][x] */,
    {YP, 0, 1, 0, 0, 1, 0},              {0, 0, 0, 0, 0, 0, 0},         {ZM | YP, -1, 1, 0, 0, 1, 1}};
static const mask_e M_FWD2[2] = {{ZP, 1, 0, 0, 0, 0, 1}, {0, 0, 0, 0, 0, 0, 0}};   /* MINforwardDE2 :260-299 */
static const mask_e M_FWD4[2] = {{ZM, -1, 0, 0, 0, 0, 1}, {0, 0, 0, 0, 0, 0, 0}};  /* MINforwardDE4 :139-175 */
static const mask_e M_FWD3[5] = { /* MINforwardDE3 :57-136 */
    {ZP, 1, 0, 0, 0, 0, 1}, {YP | ZP, 1, 0, 0, 0, 1, 1} /* sic */, {YP, 0, 1, 0, 0, 1, 0}, {0, 0, 0, 0, 0, 0, 0}, {ZM | YP, -1, 1, 0, 0, 1, 1}};
static const mask_e M_BWD3[5] = { /* MINbackwardDE3 :177-258 */
    {ZM | YM, -1, -1, 0, 0, 1, 1}, {YM, 0, -1, 0, 0, 1, 0}, {ZP | YM, 1, -1, 0, 0, 1, 1}, {0, 0, 0, 0, 0, 0, 0}, {ZM, -1, 0, 0, 0, 0, 1}};

static inline vox_t apply_mask(const vox_t* A, int S, int z, int y, int x, const mask_e* m, int n) {
    int have = (z > 0 ? ZM : 0) | (z < S - 1 ? ZP : 0) | (y > 0 ? YM : 0) | (y < S - 1 ? YP : 0) | (x > 0 ? XM : 0) | (x < S - 1 ? XP : 0);
    vox_t min; min.v = min.h = min.d = ORC_INFTY; min.dist = ORC_INFTY;
    for (int k = 0; k < n; k++) {
        if ((m[k].cond & have) != m[k].cond) continue; /* else-branch: mask = infty, never < min */
        const vox_t* s = &A[((size_t)(z + m[k].dz) * S + (y + m[k].dy)) * S + (x + m[k].dx)];
        vox_t cnd;
        cnd.v = (int16_t)(s->v + m[k].iv); cnd.h = (int16_t)(s->h + m[k].ih); cnd.d = (int16_t)(s->d + m[k].id);
        /* int arithmetic as compiled (two's-complement wrap for the all-infty case -> sqrt(negative) = NaN) */
        int32_t q = (int32_t)((uint32_t)(cnd.v * cnd.v) + (uint32_t)(cnd.h * cnd.h) + (uint32_t)(cnd.d * cnd.d));
        cnd.dist = (float)sqrt((double)q); /* sqrt1, :28 */
        if (cnd.dist < min.dist) min = cnd;
    }
    return min;
}

static void deuclidean(vox_t* A, int S) { /* DEuclidean :716-750 */
#define AT(z, y, x) A[((size_t)(z) * S + (y)) * S + (x)]
    for (int x = 0; x < S; x++) {
        for (int y = 0; y < S; y++) {
            for (int z = 0; z < S; z++) AT(z, y, x) = apply_mask(A, S, z, y, x, M_FWD1, 14);
            for (int z = S - 1; z > -1; z--) AT(z, y, x) = apply_mask(A, S, z, y, x, M_FWD2, 2);
        }
        for (int y = S - 1; y > -1; y--) {
            for (int z = S - 1; z > -1; z--) AT(z, y, x) = apply_mask(A, S, z, y, x, M_FWD3, 5);
            for (int z = 0; z < S; z++) AT(z, y, x) = apply_mask(A, S, z, y, x, M_FWD4, 2);
        }
    }
    for (int x = S - 1; x > -1; x--) {
        for (int y = S - 1; y > -1; y--) {
            for (int z = S - 1; z > -1; z--) AT(z, y, x) = apply_mask(A, S, z, y, x, M_BWD1, 14);
            for (int z = 0; z < S; z++) AT(z, y, x) = apply_mask(A, S, z, y, x, M_FWD4, 2);
        }
        for (int y = 0; y < S; y++) {
            for (int z = 0; z < S; z++) AT(z, y, x) = apply_mask(A, S, z, y, x, M_BWD3, 5);
            for (int z = S - 1; z > -1; z--) AT(z, y, x) = apply_mask(A, S, z, y, x, M_FWD2, 2);
        }
    }
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

double orc_build_dt(orc_ctx* c) {
    double t0 = now_s();
    int S = c->S, num = c->Nm; size_t S3 = (size_t)S * S * S;
    /* GoICP::BuildDT :79-90 widens the float coordinates to double */
    double xMin = c->mx[0], xMax = c->mx[0], yMin = c->my[0], yMax = c->my[0], zMin = c->mz[0], zMax = c->mz[0];
    for (int i = 1; i < num; i++) { /* :899-910 */
        double x = c->mx[i], y = c->my[i], z = c->mz[i];
        if (xMin > x) xMin = x; if (xMax < x) xMax = x;
        if (yMin > y) yMin = y; if (yMax < y) yMax = y;
        if (zMin > z) zMin = z; if (zMax < z) zMax = z;
    }
    double ef = c->p.distTransExpandFactor;
    double xC = (xMin + xMax) / 2, yC = (yMin + yMax) / 2, zC = (zMin + zMax) / 2; /* :912-914 */
    xMin = xC - ef * (xMax - xC); xMax = xC + ef * (xMax - xC); /* :915-920 (xMin uses the OLD xMax) */
    yMin = yC - ef * (yMax - yC); yMax = yC + ef * (yMax - yC);
    zMin = zC - ef * (zMax - zC); zMax = zC + ef * (zMax - zC);
    double max = xMax - xMin > yMax - yMin ? xMax - xMin : yMax - yMin; /* :921-922 */
    max = max > zMax - zMin ? max : zMax - zMin;
    xMin = xC - max / 2; xMax = xC + max / 2; yMin = yC - max / 2; yMax = yC + max / 2; zMin = zC - max / 2; zMax = zC + max / 2;
    c->xMin = xMin; c->xMax = xMax; c->yMin = yMin; c->yMax = yMax; c->zMin = zMin; c->zMax = zMax;
    c->scale = S / max; /* :931 */

    free(c->A); free(c->nearest); free(c->cellc); free(c->cell_start); free(c->cell_pts);
    c->A = (vox_t*)malloc(sizeof(vox_t) * S3);
    c->nearest = (int*)malloc(sizeof(int) * 3 * S3);
    c->cellc = (int*)malloc(sizeof(int) * S3);
    c->cell_start = (int*)calloc(S3 + 1, sizeof(int));
    c->cell_pts = (int*)malloc(sizeof(int) * (num > 0 ? num : 1));
    for (size_t i = 0; i < S3; i++) { c->A[i].v = c->A[i].h = c->A[i].d = ORC_INFTY; c->A[i].dist = ORC_INFTY; c->cellc[i] = -2; } /* :946,:962-974 */
    int* vox_of = (int*)malloc(sizeof(int) * (num > 0 ? num : 1));
    for (int i = 0; i < num; i++) { /* :976-995.  (Q3: the reference indexes cellPoints before its bounds check;
                                       with expandFactor>1 every model point is inside, we require it.) */
        int x = ORC_ROUND(((double)c->mx[i] - xMin) * c->scale), y = ORC_ROUND(((double)c->my[i] - yMin) * c->scale), z = ORC_ROUND(((double)c->mz[i] - zMin) * c->scale);
        if (x < 0 || x >= S || y < 0 || y >= S || z < 0 || z >= S) { vox_of[i] = -1; continue; }
        size_t v = ((size_t)z * S + y) * S + x; vox_of[i] = (int)v;
        c->cellc[v] = -1; c->cell_start[v + 1]++;
        c->A[v].dist = 0; c->A[v].h = c->A[v].v = c->A[v].d = 0;
    }
    for (size_t i = 0; i < S3; i++) c->cell_start[i + 1] += c->cell_start[i];
    { int* fill = (int*)calloc(S3, sizeof(int));
      for (int i = 0; i < num; i++) if (vox_of[i] >= 0) { int v = vox_of[i]; c->cell_pts[c->cell_start[v] + fill[v]++] = i; } /* push_back order = index order */
      free(fill); }
    free(vox_of);

    deuclidean(c->A, S); /* :997 */

    vox_t* A = c->A;
#define DZ(zz, yy, xx) (A[((size_t)(zz) * S + (yy)) * S + (xx)].dist == 0)
    for (int z = 0; z < S; z++) for (int y = 0; y < S; y++) for (int x = 0; x < S; x++) { /* :999-1136 */
        vox_t* a = &A[((size_t)z * S + y) * S + x];
        a->dist = (float)((double)a->dist / c->scale); /* :1004 float/double -> double -> float */
        if (a->dist < 0) a->dist = 0;
        int xD = a->v, yD = a->h, zD = a->d;
        int cx = x, cy = y, cz = z;
        if (a->dist != 0) {
            /* Candidate sign combinations in the reference's order; zero offsets collapse the list.
             * NOTE the reference reads inDE[..].distance == 0 on voxels that the SAME raster loop may already
             * have divided by scale; 0/scale == 0 so the test is unaffected. */
            static const int SGN3[8][3] = {{1, 1, 1}, {-1, 1, 1}, {1, -1, 1}, {1, 1, -1}, {-1, -1, 1}, {-1, 1, -1}, {1, -1, -1}, {-1, -1, -1}}; /* :1092-1131 (x,y,z signs) */
            static const int SGN_YZ[4][3] = {{0, 1, 1}, {0, 1, -1}, {0, -1, 1}, {0, -1, -1}};   /* xD==0 :1030-1047 */
            static const int SGN_XZ[4][3] = {{1, 0, 1}, {1, 0, -1}, {-1, 0, 1}, {-1, 0, -1}};   /* yD==0 :1055-1072 */
            static const int SGN_XY[4][3] = {{1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}};   /* zD==0 :1074-1091 */
            static const int SGN_Z[2][3] = {{0, 0, 1}, {0, 0, -1}}, SGN_Y[2][3] = {{0, 1, 0}, {0, -1, 0}}, SGN_X[2][3] = {{1, 0, 0}, {-1, 0, 0}};
            const int (*tab)[3]; int nt;
            if (xD == 0) { if (yD == 0) { tab = SGN_Z; nt = 2; } else if (zD == 0) { tab = SGN_Y; nt = 2; } else { tab = SGN_YZ; nt = 4; } }
            else if (yD == 0) { if (zD == 0) { tab = SGN_X; nt = 2; } else { tab = SGN_XZ; nt = 4; } }
            else if (zD == 0) { tab = SGN_XY; nt = 4; }
            else { tab = SGN3; nt = 8; }
            for (int k = 0; k < nt; k++) {
                int nx = x + tab[k][0] * xD, ny = y + tab[k][1] * yD, nz = z + tab[k][2] * zD;
                if (nx < 0 || nx >= S || ny < 0 || ny >= S || nz < 0 || nz >= S) continue;
                if (DZ(nz, ny, nx)) { cx = nx; cy = ny; cz = nz; break; }
            }
        }
        size_t i = ((size_t)z * S + y) * S + x;
        c->nearest[3 * i] = cx; c->nearest[3 * i + 1] = cy; c->nearest[3 * i + 2] = cz;
    }
    /* assignCellColor jly_goicp.cpp:951-969 */
    for (size_t v = 0; v < S3; v++) if (c->cellc[v] == -1) {
        int b = c->cell_start[v], e = c->cell_start[v + 1];
        int prop = c->mc[c->cell_pts[b]]; c->cellc[v] = prop;
        for (int k = b + 1; k < e; k++) if (c->mc[c->cell_pts[k]] != prop) { c->cellc[v] = -1; break; }
    }
    if (c->p.regularizationNeighbors > 0) { /* assignNeighbors jly_goicp.cpp:1213-1248 (BuildDT :94): Nd is still ALL source points here */
        free(c->nbD); free(c->nbM);
        c->nbD = (int*)calloc(c->NdAll, sizeof(int)); c->nbM = (int*)calloc(c->Nm, sizeof(int));
        const double thr = (double)sqrtf(0.050f); /* isNeighbor :1097-1103: sqrt(float radius) is the float overload */
        for (int i = 0; i < c->NdAll; i++) { int n = 0; for (int j = 0; j < c->NdAll; j++) { if (j == i) continue;
            double d = sqrt(pow((double)(c->dx[j] - c->dx[i]), 2) + pow((double)(c->dy[j] - c->dy[i]), 2) + pow((double)(c->dz[j] - c->dz[i]), 2)); if (d < thr) n++; } c->nbD[i] = n; }
        for (int i = 0; i < c->Nm; i++) { int n = 0; for (int j = 0; j < c->Nm; j++) { if (j == i) continue;
            double d = sqrt(pow((double)(c->mx[j] - c->mx[i]), 2) + pow((double)(c->my[j] - c->my[i]), 2) + pow((double)(c->mz[j] - c->mz[i]), 2)); if (d < thr) n++; } c->nbM[i] = n; }
    }
    c->dt_built = 1;
    return now_s() - t0;
}

void orc_dt_info(orc_ctx* c, double* o) { o[0] = c->xMin; o[1] = c->xMax; o[2] = c->yMin; o[3] = c->yMax; o[4] = c->zMin; o[5] = c->zMax; o[6] = c->scale; o[7] = c->S; }
void orc_dt_download(orc_ctx* c, float* dist, short* off, int* nearest, int* cellc) {
    size_t S3 = (size_t)c->S * c->S * c->S;
    for (size_t i = 0; i < S3; i++) {
        if (dist) dist[i] = c->A[i].dist;
        if (off) { off[3 * i] = c->A[i].v; off[3 * i + 1] = c->A[i].h; off[3 * i + 2] = c->A[i].d; }
    }
    if (nearest) memcpy(nearest, c->nearest, sizeof(int) * 3 * S3);
    if (cellc) memcpy(cellc, c->cellc, sizeof(int) * S3);
}
void orc_dt_upload(orc_ctx* c, const float* dist, const int* nearest) {
    size_t S3 = (size_t)c->S * c->S * c->S;
    if (dist) for (size_t i = 0; i < S3; i++) c->A[i].dist = dist[i];
    if (nearest) memcpy(c->nearest, nearest, sizeof(int) * 3 * S3);
}

/* DT3D::Distance jly_3ddt.cpp:1139-1191 */
static inline float dt_distance(const orc_ctx* c, double _x, double _y, double _z, int* ox, int* oy, int* oz) {
    int S = c->S;
    int x = ORC_ROUND((_x - c->xMin) * c->scale), y = ORC_ROUND((_y - c->yMin) * c->scale), z = ORC_ROUND((_z - c->zMin) * c->scale);
    if (ox) { *ox = x; *oy = y; *oz = z; }
    if (x > -1 && x < S && y > -1 && y < S && z > -1 && z < S) return c->A[((size_t)z * S + y) * S + x].dist;
    float a = 0, b = 0, cc = 0;
    if (x < 0) { a = x; x = 0; } else if (x >= S) { a = x - S + 1; x = S - 1; }
    if (y < 0) { b = y; y = 0; } else if (y >= S) { b = y - S + 1; y = S - 1; }
    if (z < 0) { cc = z; z = 0; } else if (z >= S) { cc = z - S + 1; z = S - 1; }
    /* sqrt(float) is the float overload in the reference's C++ (<math.h> via libstdc++) */
    return (float)((double)sqrtf(a * a + b * b + cc * cc) / c->scale + (double)c->A[((size_t)z * S + y) * S + x].dist);
}
void orc_dt_distance(orc_ctx* c, const double* xyz, int n, float* out, int* cell) {
    for (int i = 0; i < n; i++) {
        int x, y, z; out[i] = dt_distance(c, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &x, &y, &z);
        if (cell) { cell[3 * i] = x; cell[3 * i + 1] = y; cell[3 * i + 2] = z; }
    }
}

/* ================================================================================================
 * Fork error terms */
static int known_prop(int p) { for (int k = 0; k < 8; k++) if (ORC_KNOWN_PROPS[k] == p) return 1; return 0; }

/* clamp-into-grid voxel of a float position: checkCompatibility :976-984 (float args widened to double) */
static inline size_t clamp_cell(const orc_ctx* c, float x, float y, float z) {
    int S = c->S;
    int rx = ORC_ROUND(((double)x - c->xMin) * c->scale), ry = ORC_ROUND(((double)y - c->yMin) * c->scale), rz = ORC_ROUND(((double)z - c->zMin) * c->scale);
    if (rx < 0) rx = 0; if (rx >= S) rx = S - 1;
    if (ry < 0) ry = 0; if (ry >= S) ry = S - 1;
    if (rz < 0) rz = 0; if (rz >= S) rz = S - 1;
    size_t v = ((size_t)rz * S + ry) * S + rx;
    return ((size_t)c->nearest[3 * v + 2] * S + c->nearest[3 * v + 1]) * S + c->nearest[3 * v];
}
/* checkCompatibility :974-1041 + checkProperty :1068-1092 */
static int check_compat(const orc_ctx* c, int i, float x, float y, float z) {
    size_t cell = clamp_cell(c, x, y, z);
    int source = c->dc[i], target = c->cellc[cell];
    if (target == -1) { /* mixed cell: any member with the source's colour */
        for (int k = c->cell_start[cell]; k < c->cell_start[cell + 1]; k++) if (c->mc[c->cell_pts[k]] == source) return 1;
        return 0;
    }
    /* compatibilities[source] = {source} for the 8 known colours, empty otherwise (:66-73) */
    return known_prop(source) && target == source;
}
/* checkCompatibilities :919-928 */
static int check_compats(const orc_ctx* c, float x, float y, float z) {
    int n = 0;
    for (int i = 0; i < c->Nd; i++) if (!check_compat(c, i, c->tx[i] + x, c->ty[i] + y, c->tz[i] + z)) n++;
    return n;
}
static inline void fpfh_range(int mode, int* b, int* e) { *b = 0; *e = 0; if (mode == 1) { *e = 41; } else if (mode == 2) { *e = 33; } else if (mode == 3) { *b = 33; *e = 41; } }
/* computeFPFHDifference(false,...) :1643-1681 */
static float fpfh_diff_bnb(const orc_ctx* c, int point, float x, float y, float z) {
    size_t cell = clamp_cell(c, x, y, z);
    float minD = 1000000000;
    int b, e; fpfh_range(c->p.cfpfh, &b, &e);
    for (int k = c->cell_start[cell]; k < c->cell_start[cell + 1]; k++) {
        int p = c->cell_pts[k]; float diff = 0;
        for (int i = b; i < e; i++) diff += fabsf(c->df[(size_t)point * 41 + i] - c->mf[(size_t)p * 41 + i]);
        if (diff < minD) minD = diff;
    }
    return minD;
}
/* sumFPFH :1689-1697 */
static float sum_fpfh(const orc_ctx* c, float x, float y, float z) {
    float sum = 0;
    for (int i = 0; i < c->Nd; i++) sum += fpfh_diff_bnb(c, i, c->tx[i] + x, c->ty[i] + y, c->tz[i] + z);
    return sum / c->Nd;
}
/* nearestNeighbor :1200-1211 + compareNeighbors(false,...) :1250-1288: per point the closest model point INSIDE the closest
 * occupied cell (index 0 if that cell is empty), |neighbour-count difference| summed */
static int compare_neighbors_bnb(const orc_ctx* c, float x, float y, float z) {
    int sum = 0;
    for (int i = 0; i < c->Nd; i++) {
        const float ax = c->tx[i] + x, ay = c->ty[i] + y, az = c->tz[i] + z;
        const size_t cell = clamp_cell(c, ax, ay, az);
        double minD = 100; int ind = 0;
        for (int k = c->cell_start[cell]; k < c->cell_start[cell + 1]; k++) {
            const int p = c->cell_pts[k];
            const double d = sqrt(pow((double)(ax - c->mx[p]), 2) + pow((double)(ay - c->my[p]), 2) + pow((double)(az - c->mz[p]), 2));
            if (d < minD) { ind = p; minD = d; }
        }
        sum += abs(c->nbD[i] - c->nbM[ind]);
    }
    return sum;
}
/* countCompatibilities :890-914 over ICP correspondences */
static int count_compat_corr(const orc_ctx* c, const int* model_of) {
    int n = 0;
    for (int i = 0; i < c->Nd; i++) { int s = c->dc[i], t = c->mc[model_of[i]]; if (!(known_prop(s) && s == t)) n++; }
    return n;
}

/* memo of corner values within one InnerBnB call (:304-305; pure function of the float key) */
static struct memo_e* memo_probe(orc_ctx* c, uint32_t kx, uint32_t ky, uint32_t kz, int* fresh) {
    uint32_t h = (kx * 2654435761u) ^ (ky * 40503u + 0x9e3779b9u) ^ (kz * 2246822519u);
    for (int probe = 0;; probe++) {
        struct memo_e* e = &c->memo[(h + probe) & (c->memo_cap - 1)];
        if (e->gen != c->memo_gen) { e->gen = c->memo_gen; e->kx = kx; e->ky = ky; e->kz = kz; e->has = 0; *fresh = 1; return e; }
        if (e->kx == kx && e->ky == ky && e->kz == kz) { *fresh = 0; return e; }
    }
}
static struct memo_e* memo_get(orc_ctx* c, float x, float y, float z) {
    uint32_t kx, ky, kz; memcpy(&kx, &x, 4); memcpy(&ky, &y, 4); memcpy(&kz, &z, 4);
    /* -0.0f == 0.0f in the reference's float compare; normalise */
    if (kx == 0x80000000u) { kx = 0; }
    if (ky == 0x80000000u) { ky = 0; }
    if (kz == 0x80000000u) { kz = 0; }
    if (c->memo_fill * 2 > c->memo_cap) { /* grow, re-inserting this call's entries */
        struct memo_e* old = c->memo; int ocap = c->memo_cap, gen = c->memo_gen;
        c->memo_cap = 2 * ocap; c->memo = calloc(c->memo_cap, sizeof(*c->memo)); c->memo_gen = 1;
        for (int i = 0; i < ocap; i++) if (old[i].gen == gen) { int f; struct memo_e* e = memo_probe(c, old[i].kx, old[i].ky, old[i].kz, &f); e->comp = old[i].comp; e->fpfh = old[i].fpfh; e->has = old[i].has; }
        free(old);
    }
    int fresh; struct memo_e* e = memo_probe(c, kx, ky, kz, &fresh);
    c->memo_fill += fresh;
    return e;
}

/* ================================================================================================
 * GoICP::Initialize jly_goicp.cpp:180-267 */
static void neighbors_weights(orc_ctx* c) { /* :1453-1498, isNeighbor :1097-1103 */
    int Nd = c->Nd, maxN = 0, minN = 100;
    int* nb = (int*)calloc(Nd, sizeof(int));
    float distance = 0.035f;
    while (maxN < 19) {
        double thr = (double)sqrtf(distance);
        for (int i = 0; i < Nd; i++) {
            int count = 0;
            for (int j = 0; j < Nd; j++) {
                if (j == i) continue;
                double d = sqrt(pow((double)(c->dx[j] - c->dx[i]), 2) + pow((double)(c->dy[j] - c->dy[i]), 2) + pow((double)(c->dz[j] - c->dz[i]), 2));
                if (d < thr) count++;
            }
            nb[i] = count;
            if (count > maxN) maxN = count;
            if (count < minN) minN = count;
        }
        distance = (float)(distance + 0.001);
    }
    if (minN == 0) minN = 1;
    for (int i = 0; i < Nd; i++) {
        if (nb[i] == 0) nb[i] = 1;
        float f = ((float)minN / (float)nb[i]) * 2;
        c->weights[i] += f;
    }
    free(nb);
}

void orc_initialize(orc_ctx* c) {
    free_init(c);
    int Nd = c->Nd;
    c->normData = malloc(sizeof(float) * Nd); c->minDis = malloc(sizeof(float) * Nd);
    c->maxRotDis = malloc(sizeof(float) * Nd * ORC_MAXROTLEVEL); c->weights = malloc(sizeof(float) * Nd);
    c->tx = malloc(sizeof(float) * Nd); c->ty = malloc(sizeof(float) * Nd); c->tz = malloc(sizeof(float) * Nd);
    c->icp_model = calloc(Nd, sizeof(int)); c->opt_model = calloc(Nd, sizeof(int));
    for (int i = 0; i < Nd; i++) c->normData[i] = sqrtf(c->dx[i] * c->dx[i] + c->dy[i] * c->dy[i] + c->dz[i] * c->dz[i]); /* :191 */
    for (int l = 0; l < ORC_MAXROTLEVEL; l++) { /* :195-206 */
        float sigma = (float)(c->p.rotWidth / pow(2.0, l) / 2.0);
        float maxAngle = (float)(ORC_SQRT3 * sigma);
        if (maxAngle > ORC_PI) maxAngle = (float)ORC_PI;
        float s2 = 2 * sinf(maxAngle / 2);
        for (int j = 0; j < Nd; j++) c->maxRotDis[(size_t)l * Nd + j] = s2 * c->normData[j];
    }
    for (int i = 0; i < 9; i++) c->optR[i] = (i % 4 == 0);
    c->optT[0] = c->optT[1] = c->optT[2] = 0; /* :240-241 */
    c->inlierNum = c->doTrim ? (int)(Nd * (1 - c->p.trimFraction)) : Nd; /* :244-252 */
    for (int i = 0; i < Nd; i++) c->weights[i] = 1;
    if (c->p.ponderation == 1) neighbors_weights(c); /* :262 */
    c->SSEThresh = c->p.MSEThresh * c->inlierNum; /* :266 */
    c->icp_mu_m[0] = c->icp_mu_m[1] = c->icp_mu_m[2] = 0; c->icp_mu_d[0] = c->icp_mu_d[1] = c->icp_mu_d[2] = 0;
    if (!c->memo) { c->memo_cap = 1 << 16; c->memo = calloc(c->memo_cap, sizeof(*c->memo)); c->memo_gen = 0; }
    c->initialized = 1;
}
void orc_get_weights(orc_ctx* c, float* w) { memcpy(w, c->weights, sizeof(float) * c->Nd); }
void orc_get_maxrotdis(orc_ctx* c, float* o) { memcpy(o, c->maxRotDis, sizeof(float) * c->Nd * ORC_MAXROTLEVEL); }
float orc_get_ssethresh(orc_ctx* c) { return c->SSEThresh; }
int orc_get_inliernum(orc_ctx* c) { return c->inlierNum; }

/* ================================================================================================
 * intro_select jly_sorting.hpp:229-313 -- the oracle only needs its CONTRACT (the inlierNum smallest values
 * first); which permutation it leaves is irrelevant to the inclusion set, and the float sum over them is
 * compared at tolerance.  We therefore sort ascending (a valid outcome of the contract). */
static int cmp_float(const void* a, const void* b) { float x = *(const float*)a, y = *(const float*)b; return (x > y) - (x < y); }
static void select_smallest(float* d, int n) { qsort(d, n, sizeof(float), cmp_float); }

/* leaf evaluation of one child translation cube: jly_goicp.cpp:331-550 */
typedef struct { float ub, lb; int minIncomp, maxIncomp; float minFPFH, maxFPFH; } leaf_t;

static int corner_comp(orc_ctx* c, float x, float y, float z) {
    struct memo_e* e = memo_get(c, x, y, z);
    if (!(e->has & 1)) { e->comp = check_compats(c, x, y, z); e->has |= 1; }
    return e->comp;
}
static float corner_fpfh(orc_ctx* c, float x, float y, float z) {
    struct memo_e* e = memo_get(c, x, y, z);
    if (!(e->has & 2)) { e->fpfh = sum_fpfh(c, x, y, z); e->has |= 2; }
    return e->fpfh;
}

static leaf_t eval_leaf(orc_ctx* c, const float* maxRotDisL, float nx, float ny, float nz, float nw) {
    leaf_t r; memset(&r, 0, sizeof r);
    int Nd = c->Nd, norm = c->p.norm;
    float transX = nx + nw / 2, transY = ny + nw / 2, transZ = nz + nw / 2; /* :331-333 */
    float maxTransDis = (float)(ORC_SQRT3 / 2.0 * nw);                        /* :323 */
    for (int i = 0; i < Nd; i++) { /* :343-382 -- one cube.point bound eval */
        float d = c->weights[i] * dt_distance(c, (double)(c->tx[i] + transX), (double)(c->ty[i] + transY), (double)(c->tz[i] + transZ), 0, 0, 0);
        if (maxRotDisL) d -= maxRotDisL[i];
        if (d < 0) d = 0;
        c->minDis[i] = d;
    }
    c->cnt[2]++;
    if (c->doTrim) select_smallest(c->minDis, Nd); /* :384-390 */
    float ub = 0, lb = 0;
    for (int i = 0; i < c->inlierNum; i++) { if (norm == 2) ub += c->minDis[i] * c->minDis[i]; if (norm == 1) ub += c->minDis[i]; } /* :393-401 */
    for (int i = 0; i < c->inlierNum; i++) { /* :403-415 */
        float dis = c->minDis[i] - maxTransDis;
        if (dis > 0) { if (norm == 2) lb += dis * dis; if (norm == 1) lb += dis; }
    }
    float reg = c->p.regularization, regF = c->p.regularizationFPFH;
    if (reg > 0 || c->p.regularizationNeighbors > 0 || (regF > 0 && c->p.cfpfh != 0)) { /* :436 */
        int minI = 0, maxI = 0, minN = 0, maxN = 0; float minF = 0, maxF = 0; /* floats holding int-truncated values (H7) */
        const float regN = c->p.regularizationNeighbors;
        for (int k = 0; k < 8; k++) { /* :437-439,:486-488 */
            float xI = nx + (k & 1) * nw, yI = ny + (k >> 1 & 1) * nw, zI = nz + (k >> 2 & 1) * nw;
            if (regF > 0) { int f = (int)corner_fpfh(c, xI, yI, zI); if (k == 0) { minF = maxF = (float)f; } else { if (f > maxF) maxF = (float)f; if (f < minF) minF = (float)f; } }
            if (reg > 0) { int n = corner_comp(c, xI, yI, zI); if (k == 0) { minI = maxI = n; } else { if (n > maxI) maxI = n; if (n < minI) minI = n; } }
            if (regN > 0) { int n = compare_neighbors_bnb(c, xI, yI, zI); if (k == 0) { minN = maxN = n; } else { if (n > maxN) maxN = n; if (n < minN) minN = n; } } /* :462-466,:489-493, not memoised */
        }
        if (reg > 0) { ub += reg * (maxI * maxI); lb += reg * (minI * minI); }       /* :536-538 */
        if (regN > 0) { ub += regN * (maxN * maxN); lb += regN * (minN * minN); }   /* :542-545 */
        if (regF > 0) { ub += regF * (maxF * maxF); lb += regF * (minF * minF); }    /* :546-549 */
        r.minIncomp = minI; r.maxIncomp = maxI; r.minFPFH = minF; r.maxFPFH = maxF;
    }
    r.ub = ub; r.lb = lb;
    return r;
}

static void rotate_data(orc_ctx* c, const float* R) { /* jly_goicp.cpp:750-762 */
    for (int i = 0; i < c->Nd; i++) {
        if (R) {
            c->tx[i] = R[0] * c->dx[i] + R[1] * c->dy[i] + R[2] * c->dz[i];
            c->ty[i] = R[3] * c->dx[i] + R[4] * c->dy[i] + R[5] * c->dz[i];
            c->tz[i] = R[6] * c->dx[i] + R[7] * c->dy[i] + R[8] * c->dz[i];
        } else { c->tx[i] = c->dx[i]; c->ty[i] = c->dy[i]; c->tz[i] = c->dz[i]; }
    }
}

/* GoICP::InnerBnB jly_goicp.cpp:286-579 on the current pDataTemp */
static float inner_bnb(orc_ctx* c, const float* maxRotDisL, node_t* nodeTransOut) {
    heap_t q = {0, 0, 0};
    float optErrorT = c->optError; /* :297 */
    c->cnt[0]++;
    node_t init; memset(&init, 0, sizeof init);
    init.a = c->p.transMinX; init.b = c->p.transMinY; init.c = c->p.transMinZ; init.w = c->p.transWidth; init.lb = 0;
    heap_push(&q, init);
    c->memo_gen++; c->memo_fill = 0;
    while (q.n) {
        node_t parent = heap_pop(&q); c->cnt[1]++;
        if (optErrorT - parent.lb < c->SSEThresh) break; /* :317 */
        node_t nt; memset(&nt, 0, sizeof nt);
        nt.w = parent.w / 2;
        for (int j = 0; j < 8; j++) {
            nt.a = parent.a + (j & 1) * nt.w; nt.b = parent.b + (j >> 1 & 1) * nt.w; nt.c = parent.c + (j >> 2 & 1) * nt.w; /* :327-329 */
            leaf_t r = eval_leaf(c, maxRotDisL, nt.a, nt.b, nt.c, nt.w);
            if (r.ub < optErrorT) { optErrorT = r.ub; if (nodeTransOut) *nodeTransOut = nt; } /* :554-566 */
            if (r.lb >= optErrorT) continue;                                                /* :568-572 */
            nt.ub = r.ub; nt.lb = r.lb; heap_push(&q, nt);
        }
    }
    free(q.v);
    return optErrorT;
}

float orc_inner_bnb(orc_ctx* c, const float* R, int level, float optError, float* tnode) {
    rotate_data(c, R);
    c->optError = optError;
    node_t out; memset(&out, 0, sizeof out);
    float e = inner_bnb(c, level >= 0 ? c->maxRotDis + (size_t)level * c->Nd : NULL, tnode ? &out : NULL);
    if (tnode) { tnode[0] = out.a; tnode[1] = out.b; tnode[2] = out.c; tnode[3] = out.w; }
    return e;
}

void orc_eval_leaf(orc_ctx* c, const float* R, int level, const float* tcube, int n, float* ub, float* lb, int* incomp, int* fpfh) {
    rotate_data(c, R);
    c->memo_gen++; c->memo_fill = 0;
    for (int k = 0; k < n; k++) {
        leaf_t r = eval_leaf(c, level >= 0 ? c->maxRotDis + (size_t)level * c->Nd : NULL, tcube[4 * k], tcube[4 * k + 1], tcube[4 * k + 2], tcube[4 * k + 3]);
        ub[k] = r.ub; lb[k] = r.lb;
        if (incomp) { incomp[2 * k] = r.minIncomp; incomp[2 * k + 1] = r.maxIncomp; }
        if (fpfh) { fpfh[2 * k] = (int)r.minFPFH; fpfh[2 * k + 1] = (int)r.maxFPFH; }
    }
}

/* Per-cube point-inclusion set of the trimmed error (jly_goicp.cpp:384-390: intro_select leaves the inlierNum smallest
 * residuals first).  resid[k*Nd+i] = the clamped residual of point i under cube k (:343-382); mask[k*Nd+i] = 1 when point
 * i is among the inlierNum smallest.  The set is unique up to ties AT the inlierNum-th smallest value (the reference's
 * choice among equal values depends on intro_select's permutation and does not change the sum); the canonical form here
 * takes tied points in ascending index order.  Without trimming every point is included. */
void orc_eval_inclusion(orc_ctx* c, const float* R, int level, const float* tcube, int n, unsigned char* mask, float* resid) {
    rotate_data(c, R);
    const int Nd = c->Nd, k = c->inlierNum;
    const float* maxRotDisL = level >= 0 ? c->maxRotDis + (size_t)level * Nd : NULL;
    float* tmp = malloc(sizeof(float) * Nd);
    for (int q = 0; q < n; q++) {
        float nw = tcube[4 * q + 3];
        float transX = tcube[4 * q] + nw / 2, transY = tcube[4 * q + 1] + nw / 2, transZ = tcube[4 * q + 2] + nw / 2;
        for (int i = 0; i < Nd; i++) {
            float d = c->weights[i] * dt_distance(c, (double)(c->tx[i] + transX), (double)(c->ty[i] + transY), (double)(c->tz[i] + transZ), 0, 0, 0);
            if (maxRotDisL) d -= maxRotDisL[i];
            if (d < 0) d = 0;
            tmp[i] = d;
            if (resid) resid[(size_t)q * Nd + i] = d;
        }
        if (!c->doTrim || k >= Nd) { memset(mask + (size_t)q * Nd, 1, Nd); continue; }
        float* srt = c->minDis; memcpy(srt, tmp, sizeof(float) * Nd);
        qsort(srt, Nd, sizeof(float), cmp_float);
        const float T = srt[k - 1];
        int below = 0; for (int i = 0; i < Nd; i++) below += tmp[i] < T;
        int need_eq = k - below;
        for (int i = 0; i < Nd; i++) {
            unsigned char in = 0;
            if (tmp[i] < T) in = 1; else if (tmp[i] == T && need_eq > 0) { in = 1; need_eq--; }
            mask[(size_t)q * Nd + i] = in;
        }
    }
    free(tmp);
}

/* ================================================================================================
 * ICP: ICP3D<float>::Run jly_icp3d.hpp:197-311 (exact NN by exhaustive search instead of nanoflann; same
 * float distance expression as L2_Simple_Adaptor nanoflann.hpp, first-found wins ties), then the DT re-score
 * of GoICP::ICP jly_goicp.cpp:102-178.  SVD: one-sided Jacobi in double (Matrix::svd matrix.cpp:582 is a
 * Golub-Kahan SVD; only V*diag(1,1,det)*U^T is consumed, which is unique for a full-rank H). */
static void svd3(const double H[9], double U[9], double W[3], double V[9]) {
    double A[9]; memcpy(A, H, sizeof A);
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
            double al = 0, be = 0, ga = 0;
            for (int k = 0; k < 3; k++) { al += A[3 * k + p] * A[3 * k + p]; be += A[3 * k + q] * A[3 * k + q]; ga += A[3 * k + p] * A[3 * k + q]; }
            if (ga == 0) continue;
            if (fabs(ga) > off * 0 + 1e-300) { double r = fabs(ga) / sqrt(al * be + 1e-300); if (r > off) off = r; }
            double zeta = (be - al) / (2 * ga);
            double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
            double cs = 1 / sqrt(1 + t * t), sn = cs * t;
            for (int k = 0; k < 3; k++) {
                double ap = A[3 * k + p], aq = A[3 * k + q]; A[3 * k + p] = cs * ap - sn * aq; A[3 * k + q] = sn * ap + cs * aq;
                double vp = V[3 * k + p], vq = V[3 * k + q]; V[3 * k + p] = cs * vp - sn * vq; V[3 * k + q] = sn * vp + cs * vq;
            }
        }
        if (off < 1e-15) break;
    }
    for (int j = 0; j < 3; j++) {
        double n = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
        W[j] = n;
        for (int k = 0; k < 3; k++) U[3 * k + j] = n > 0 ? A[3 * k + j] / n : 0;
    }
    /* order the singular triplets by descending singular value, as Matrix::svd does (matrix.cpp:770-800): the float `det`
     * of jly_icp3d.hpp:291-297 then scales the same (smallest) direction as in the reference */
    for (int a = 0; a < 2; a++) for (int b = a + 1; b < 3; b++) if (W[b] > W[a]) {
        double tw = W[a]; W[a] = W[b]; W[b] = tw;
        for (int k = 0; k < 3; k++) { double tu = U[3 * k + a]; U[3 * k + a] = U[3 * k + b]; U[3 * k + b] = tu; double tv = V[3 * k + a]; V[3 * k + a] = V[3 * k + b]; V[3 * k + b] = tv; }
    }
    /* rank-deficient column: complete U to an orthonormal basis (cross product of the other two) */
    for (int j = 0; j < 3; j++) if (W[j] <= 1e-300) {
        int a = (j + 1) % 3, b = (j + 2) % 3;
        U[j] = U[3 + a] * U[6 + b] - U[6 + a] * U[3 + b];
        U[3 + j] = U[6 + a] * U[b] - U[a] * U[6 + b];
        U[6 + j] = U[a] * U[3 + b] - U[3 + a] * U[b];
    }
}

typedef struct { double dis; int id_data, id_model; } pref_t;
static int cmp_pref(const void* a, const void* b) { return ((const pref_t*)a)->dis > ((const pref_t*)b)->dis ? 1 : -1; } /* jly_icp3d.hpp:173 */

static float icp_run(orc_ctx* c, double* R, double* t, pref_t* points) {
    int n = c->Nd, Nm = c->Nm;
    int num = c->doTrim ? (int)(n * (1 - c->p.trimFraction)) : n; /* :206-213 (trim_fraction is float) */
    float err_diff = c->p.MSEThresh / 10000;                       /* jly_goicp.cpp:232 */
    double* p_m = malloc(sizeof(double) * 3 * num); double* p_d = malloc(sizeof(double) * 3 * num);
    double mu_m[3] = {0, 0, 0}, mu_d[3] = {0, 0, 0}; /* Matrix mu_m(1,3): constructed once per Run, never reset (Q4) */
    float err = -1, err_new = 0;
    for (size_t iter = 0; iter < 10000; iter++) {
        float r00 = (float)R[0], r01 = (float)R[1], r02 = (float)R[2], r10 = (float)R[3], r11 = (float)R[4], r12 = (float)R[5], r20 = (float)R[6], r21 = (float)R[7], r22 = (float)R[8];
        float t0 = (float)t[0], t1 = (float)t[1], t2 = (float)t[2];
        err_new = 0;
        for (int i = 0; i < n; i++) { /* :234-250 */
            float q0 = r00 * c->dx[i] + r01 * c->dy[i] + r02 * c->dz[i] + t0;
            float q1 = r10 * c->dx[i] + r11 * c->dy[i] + r12 * c->dz[i] + t1;
            float q2 = r20 * c->dx[i] + r21 * c->dy[i] + r22 * c->dz[i] + t2;
            float best = INFINITY; int bi = 0;
            for (int m = 0; m < Nm; m++) {
                float d0 = q0 - c->mx[m], d1 = q1 - c->my[m], d2 = q2 - c->mz[m];
                float d = d0 * d0; d += d1 * d1; d += d2 * d2;
                if (d < best) { best = d; bi = m; }
            }
            points[i].dis = best; points[i].id_data = i; points[i].id_model = bi;
        }
        if (c->doTrim) qsort(points, n, sizeof(pref_t), cmp_pref); /* :252-255 */
        for (int i = 0; i < num; i++) { /* :257-271 */
            int m = points[i].id_model, id = points[i].id_data;
            p_m[3 * i] = c->mx[m]; mu_m[0] += p_m[3 * i]; p_m[3 * i + 1] = c->my[m]; mu_m[1] += p_m[3 * i + 1]; p_m[3 * i + 2] = c->mz[m]; mu_m[2] += p_m[3 * i + 2];
            p_d[3 * i] = r00 * c->dx[id] + r01 * c->dy[id] + r02 * c->dz[id] + t0; mu_d[0] += p_d[3 * i];
            p_d[3 * i + 1] = r10 * c->dx[id] + r11 * c->dy[id] + r12 * c->dz[id] + t1; mu_d[1] += p_d[3 * i + 1];
            p_d[3 * i + 2] = r20 * c->dx[id] + r21 * c->dy[id] + r22 * c->dz[id] + t2; mu_d[2] += p_d[3 * i + 2];
            err_new = (float)((double)err_new + points[i].dis);
        }
        if (err > 0 && err - err_new < err_diff * num) break; /* :273 */
        err = err_new;
        for (int k = 0; k < 3; k++) { mu_m[k] = mu_m[k] / (float)n; mu_d[k] = mu_d[k] / (float)n; } /* :278-279: /n not /num */
        double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; /* H = ~q_t * q_m :284, Matrix operator* accumulates over k in order */
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double s = 0; for (int k = 0; k < num; k++) s += (p_d[3 * k + a] - mu_d[a]) * (p_m[3 * k + b] - mu_m[b]); H[3 * a + b] = s; }
        double U[9], W[3], V[9]; svd3(H, U, W, V);
        double R_[9];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double s = 0; for (int k = 0; k < 3; k++) s += V[3 * a + k] * U[3 * b + k]; R_[3 * a + b] = s; } /* V*~U :287 */
        float da = (float)(R_[0] * (R_[4] * R_[8] - R_[5] * R_[7])), db = (float)(-R_[1] * (R_[3] * R_[8] - R_[5] * R_[6])), dc = (float)(R_[2] * (R_[3] * R_[7] - R_[4] * R_[6]));
        float det = da + db + dc; /* T = float :291-297 */
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double s = 0; for (int k = 0; k < 3; k++) s += V[3 * a + k] * (k == 2 ? (double)det : 1.0) * U[3 * b + k]; R_[3 * a + b] = s; } /* :299-302 */
        double t_[3]; for (int a = 0; a < 3; a++) t_[a] = mu_m[a] - (R_[3 * a] * mu_d[0] + R_[3 * a + 1] * mu_d[1] + R_[3 * a + 2] * mu_d[2]); /* :304 */
        double Rn[9], tn[3];
        for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) { double s = 0; for (int k = 0; k < 3; k++) s += R_[3 * a + k] * R[3 * k + b]; Rn[3 * a + b] = s; }
                                      tn[a] = R_[3 * a] * t[0] + R_[3 * a + 1] * t[1] + R_[3 * a + 2] * t[2] + t_[a]; } /* :307-308 */
        memcpy(R, Rn, sizeof Rn); memcpy(t, tn, sizeof tn);
    }
    free(p_m); free(p_d);
    return err_new;
}

/* GoICP::ICP jly_goicp.cpp:102-178.  geom/incomp/fpfh decomposition is diagnostic only. */
static float goicp_icp(orc_ctx* c, double* R, double* t) {
    int Nd = c->Nd, norm = c->p.norm;
    pref_t* points = malloc(sizeof(pref_t) * Nd);
    c->cnt[5]++;
    icp_run(c, R, t, points);
    float error = 0, fpfh = 0;
    int b, e; fpfh_range(c->p.cfpfh, &b, &e);
    for (int i = 0; i < Nd; i++) {
        float x = (float)(R[0] * c->dx[i] + R[1] * c->dy[i] + R[2] * c->dz[i] + t[0]); /* double expr -> float store :120-122 */
        float y = (float)(R[3] * c->dx[i] + R[4] * c->dy[i] + R[5] * c->dz[i] + t[1]);
        float z = (float)(R[6] * c->dx[i] + R[7] * c->dy[i] + R[8] * c->dz[i] + t[2]);
        if (!c->doTrim) {
            float dis = c->weights[i] * dt_distance(c, x, y, z, 0, 0, 0);
            if (norm == 2) error += dis * dis; if (norm == 1) error += dis;
        } else c->minDis[i] = dt_distance(c, x, y, z, 0, 0, 0); /* :135 -- no weight when trimming */
        if (c->p.cfpfh != 0) { /* computeFPFHDifference(true, i) :1625-1641 uses points[i] (sorted order if trimmed) */
            float d = 0; int idd = points[i].id_data, idm = points[i].id_model;
            for (int k = b; k < e; k++) d += fabsf(c->df[(size_t)idd * 41 + k] - c->mf[(size_t)idm * 41 + k]);
            fpfh += d;
        }
    }
    fpfh = fpfh / Nd; /* :147 */
    /* correspondences per DATA index (countCompatibilities :890-914 sums over all pairs, order-free) */
    for (int i = 0; i < Nd; i++) c->icp_model[points[i].id_data] = points[i].id_model;
    if (c->p.regularizationNeighbors > 0) { /* :149-153 compareNeighbors(true): over the correspondences */
        int nb = 0;
        for (int i = 0; i < Nd; i++) nb += abs(c->nbD[points[i].id_data] - c->nbM[points[i].id_model]);
        error += c->p.regularizationNeighbors * (nb * nb);
    }
    if (c->p.regularization > 0) { /* :154-159 countCompatibilities(true) */
        int incomp = 0;
        for (int i = 0; i < Nd; i++) { int s = c->dc[points[i].id_data], tt = c->mc[points[i].id_model]; if (!(known_prop(s) && s == tt)) incomp++; }
        error += c->p.regularization * (incomp * incomp);
    }
    if (c->p.regularizationFPFH > 0) error += c->p.regularizationFPFH * (fpfh * fpfh); /* :160-163 */
    if (c->doTrim) { /* :166-175 */
        select_smallest(c->minDis, Nd);
        for (int i = 0; i < c->inlierNum; i++) error += c->minDis[i] * c->minDis[i];
    }
    free(points);
    return error;
}

float orc_icp(orc_ctx* c, double* R, double* t, int* corr) {
    /* correspondences returned per DATA index */
    int Nd = c->Nd;
    pref_t* points = malloc(sizeof(pref_t) * Nd);
    double R2[9], t2[3]; memcpy(R2, R, sizeof R2); memcpy(t2, t, sizeof t2);
    icp_run(c, R2, t2, points);
    if (corr) for (int i = 0; i < Nd; i++) corr[points[i].id_data] = points[i].id_model;
    free(points);
    long long save = c->cnt[5];
    float e = goicp_icp(c, R, t);
    c->cnt[5] = save + 1;
    return e;
}

/* updateCompatibilities jly_goicp.cpp:933-946 : pose in double (Matrix), coordinates stored to float */
static int update_compat(orc_ctx* c) {
    int n = 0;
    for (int i = 0; i < c->Nd; i++) {
        float x = (float)(c->optR[0] * c->dx[i] + c->optR[1] * c->dy[i] + c->optR[2] * c->dz[i] + c->optT[0]);
        float y = (float)(c->optR[3] * c->dx[i] + c->optR[4] * c->dy[i] + c->optR[5] * c->dz[i] + c->optT[1]);
        float z = (float)(c->optR[6] * c->dx[i] + c->optR[7] * c->dy[i] + c->optR[8] * c->dz[i] + c->optT[2]);
        if (!check_compat(c, i, x, y, z)) n++;
    }
    return n;
}

/* GoICP::OuterBnB jly_goicp.cpp:582-876 */
static float outer_bnb(orc_ctx* c) {
    int Nd = c->Nd, norm = c->p.norm;
    heap_t q = {0, 0, 0};
    c->optError = 0;
    for (int i = 0; i < Nd; i++) c->minDis[i] = c->weights[i] * dt_distance(c, c->dx[i], c->dy[i], c->dz[i], 0, 0, 0); /* :602-605 */
    if (c->doTrim) select_smallest(c->minDis, Nd);
    for (int i = 0; i < c->inlierNum; i++) { if (norm == 2) c->optError += c->minDis[i] * c->minDis[i]; if (norm == 1) c->optError += c->minDis[i]; }
    if (c->p.regularization > 0) c->optError += c->p.regularization * (Nd * Nd);                 /* :623 */
    if (c->p.regularizationFPFH > 0) c->optError += c->p.regularizationFPFH * (100 * 8 * 100 * 8); /* :624 */
    if (c->p.regularizationNeighbors > 0) c->optError += c->p.regularizationNeighbors * (Nd * 6 * Nd * 6);
    tracef(c, "Error*: %g (Init)\n", c->optError);

    double R_icp[9], t_icp[3]; memcpy(R_icp, c->optR, sizeof R_icp); memcpy(t_icp, c->optT, sizeof t_icp);
    float error = goicp_icp(c, R_icp, t_icp); /* :634 */
    memcpy(c->opt_model, c->icp_model, sizeof(int) * Nd); /* :635 optPoints = points (unconditional) */
    if (error < c->optError) { /* :636-661 */
        c->optError = error; memcpy(c->optR, R_icp, sizeof R_icp); memcpy(c->optT, t_icp, sizeof t_icp);
        c->optComp = count_compat_corr(c, c->opt_model); /* :650 -- untrimmed: points[i].id_data == i */
        tracef(c, "Error*: %g (ICP)\n", error);
    }
    node_t init; memset(&init, 0, sizeof init);
    init.a = c->p.rotMinX; init.b = c->p.rotMinY; init.c = c->p.rotMinZ; init.w = c->p.rotWidth; init.l = 0; init.lb = 0;
    heap_push(&q, init);
    float lb = 0;
    while (1) {
        if (q.n == 0) { tracef(c, "Rotation Queue Empty\nError*: %g, LB: %g\n", c->optError, lb); break; } /* :670-677 */
        node_t parent = heap_pop(&q); c->cnt[3]++;
        if ((c->optError - parent.lb) <= c->SSEThresh) { tracef(c, "Threshold reached\nError*: %g, LB: %g, epsilon: %g\n", c->optError, parent.lb, c->SSEThresh); break; } /* :685 */
        node_t nr; memset(&nr, 0, sizeof nr);
        nr.w = parent.w / 2; nr.l = parent.l + 1;
        for (int j = 0; j < 8; j++) {
            nr.a = parent.a + (j & 1) * nr.w; nr.b = parent.b + (j >> 1 & 1) * nr.w; nr.c = parent.c + (j >> 2 & 1) * nr.w; /* :710-712 */
            float v1 = nr.a + nr.w / 2, v2 = nr.b + nr.w / 2, v3 = nr.c + nr.w / 2;
            if ((double)sqrtf(v1 * v1 + v2 * v2 + v3 * v3) - ORC_SQRT3 * nr.w / 2 > ORC_PI) continue; /* :723 */
            float t = sqrtf(v1 * v1 + v2 * v2 + v3 * v3); /* :729 */
            float R[9];
            if (t > 0) { /* :730-756 */
                v1 /= t; v2 /= t; v3 /= t;
                float ct = cosf(t), ct2 = 1 - ct, st = sinf(t);
                float tmp121 = v1 * v2 * ct2, tmp122 = v3 * st, tmp131 = v1 * v3 * ct2, tmp132 = v2 * st, tmp231 = v2 * v3 * ct2, tmp232 = v1 * st;
                R[0] = ct + v1 * v1 * ct2; R[1] = tmp121 - tmp122; R[2] = tmp131 + tmp132;
                R[3] = tmp121 + tmp122; R[4] = ct + v2 * v2 * ct2; R[5] = tmp231 - tmp232;
                R[6] = tmp131 - tmp132; R[7] = tmp231 + tmp232; R[8] = ct + v3 * v3 * ct2;
                rotate_data(c, R);
            } else {
                /* :759-762 memcpy branch; R11.. keep their previous (stale) values in the reference. We use identity. */
                for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0);
                rotate_data(c, NULL);
            }
            c->cnt[4]++;
            node_t ntrans; memset(&ntrans, 0, sizeof ntrans);
            float ub = inner_bnb(c, NULL, &ntrans); /* :768 */
            if (ub < c->optError) { /* :771-854 */
                c->optError = ub;
                for (int k = 0; k < 9; k++) c->optR[k] = R[k];
                c->optT[0] = ntrans.a + ntrans.w / 2; c->optT[1] = ntrans.b + ntrans.w / 2; c->optT[2] = ntrans.c + ntrans.w / 2; /* float expr -> double */
                c->optComp = update_compat(c); /* :791 */
                tracef(c, "Error*: %g (BNB)\n", c->optError);
                memcpy(R_icp, c->optR, sizeof R_icp); memcpy(t_icp, c->optT, sizeof t_icp);
                error = goicp_icp(c, R_icp, t_icp); /* :810 */
                if (error < c->optError) { /* :813-840 */
                    c->optError = error; memcpy(c->optR, R_icp, sizeof R_icp); memcpy(c->optT, t_icp, sizeof t_icp);
                    memcpy(c->opt_model, c->icp_model, sizeof(int) * Nd);
                    c->optComp = count_compat_corr(c, c->opt_model);
                    tracef(c, "Error*: %g (ICP)\n", error);
                }
                heap_t qn = {0, 0, 0}; /* :843-853 */
                while (q.n) { node_t n = heap_pop(&q); if (n.lb < c->optError) heap_push(&qn, n); else break; }
                free(q.v); q = qn;
            }
            lb = inner_bnb(c, c->maxRotDis + (size_t)nr.l * Nd, NULL); /* :861 (Q2: no bound check on level) */
            if (lb >= c->optError) continue;
            nr.ub = ub; nr.lb = lb; heap_push(&q, nr);
        }
    }
    free(q.v);
    return c->optError;
}

void orc_register(orc_ctx* c, int nd_downsampled, orc_result* out, char* trace, int trace_cap) {
    memset(c->cnt, 0, sizeof c->cnt); c->trace_len = 0; if (c->trace) c->trace[0] = 0;
    double t0 = now_s();
    if (!c->dt_built) orc_build_dt(c);
    if (nd_downsampled > 0) c->Nd = nd_downsampled;
    double t1 = now_s();
    orc_initialize(c);
    outer_bnb(c);
    double t2 = now_s();
    memcpy(out->R, c->optR, sizeof out->R); memcpy(out->t, c->optT, sizeof out->t);
    out->optError = c->optError; out->optComp = c->optComp;
    memcpy(out->counters, c->cnt, sizeof c->cnt);
    out->seconds_dt = t1 - t0; out->seconds_register = t2 - t1;
    if (trace && trace_cap > 0) { int n = c->trace_len < trace_cap - 1 ? c->trace_len : trace_cap - 1; if (n > 0) memcpy(trace, c->trace, n); trace[n] = 0; }
}

/* ================================================================================================
 * Transformation (transformation.cpp) */
double orc_normalize(double* xyz, int n, double* mean) { /* :311-335 */
    double xm = 0, ym = 0, zm = 0;
    for (int i = 0; i < n; i++) { xm += xyz[3 * i]; ym += xyz[3 * i + 1]; zm += xyz[3 * i + 2]; }
    xm /= (unsigned long)n; ym /= (unsigned long)n; zm /= (unsigned long)n;
    double maxNorm = 0;
    for (int i = 0; i < n; i++) {
        xyz[3 * i] -= xm; xyz[3 * i + 1] -= ym; xyz[3 * i + 2] -= zm;
        double norm = sqrt(pow(xyz[3 * i], 2) + pow(xyz[3 * i + 1], 2) + pow(xyz[3 * i + 2], 2));
        if (norm > maxNorm) maxNorm = norm;
    }
    mean[0] = xm; mean[1] = ym; mean[2] = zm;
    return maxNorm;
}
void orc_scale(double* xyz, int n, double scale) { for (int i = 0; i < 3 * n; i++) xyz[i] /= scale; } /* :355-361 */
/* writeNormalizedMolCloudFile :340-349 prints with the default ostream precision (6 significant digits, %g);
 * loadPointCloud jly_main.cpp:307 parses the text back into float. */
void orc_round6(const double* xyz, int n, float* out) {
    char buf[64];
    for (int i = 0; i < 3 * n; i++) { snprintf(buf, sizeof buf, "%g", xyz[i]); out[i] = strtof(buf, NULL); }
}
void orc_rescale_translation(double scale, const double* mT, const double* mS, const double* R, const double* t, double* o) { /* :410-412 */
    for (int a = 0; a < 3; a++) o[a] = -(R[3 * a] * mS[0] + R[3 * a + 1] * mS[1] + R[3 * a + 2] * mS[2]) + (scale * t[a]) + mT[a];
}
void orc_apply_rigid(const double* xyz, int n, const double* R, const double* t, double* out) { /* :485-497 */
    for (int i = 0; i < n; i++) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        for (int a = 0; a < 3; a++) { double v = R[3 * a] * x + R[3 * a + 1] * y + R[3 * a + 2] * z; v += t[a]; out[3 * i + a] = v; }
    }
}
float orc_rmsd(const double* a, const double* b, int n) { /* :453-464: float accumulator, double terms */
    float rmsd = 0;
    for (int i = 0; i < n; i++) rmsd = (float)((double)rmsd + (pow(a[3 * i] - b[3 * i], 2) + pow(a[3 * i + 1] - b[3 * i + 1], 2) + pow(a[3 * i + 2] - b[3 * i + 2], 2)));
    return sqrtf(rmsd / (unsigned long)n); /* float / size_t -> float; sqrt(float) float overload */
}
