/* TEST INFRASTRUCTURE ONLY -- the CPU oracle: a plain-C restatement of the reference's Go-ICP hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (go-icp-protein-cavities_b200/) never links, imports or calls anything under oracle/.
 *
 * Parity is PINNED: tests/test_oracle_*.py check this restatement against (a) the reference's shipped golden
 * files (output/similar1.txt, demo/output.txt ...; fixtures in tests/golden/) and (b) the reference itself
 * compiled in place (oracle/_ref/libgoicp_ref.so) function by function: DT values + offsets + index map,
 * Distance, weights, maxRotDis, every InnerBnB call, ICP, full Register with traces and node counters.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef GOICP_ORACLE_H
#define GOICP_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* config.txt keys as read by readConfig (jly_main.cpp:231-270); same layout as goicp_params. */
typedef struct orc_params {
    float MSEThresh;
    float rotMinX, rotMinY, rotMinZ, rotWidth;
    float transMinX, transMinY, transMinZ, transWidth;
    float trimFraction;
    float regularization, regularizationNeighbors, regularizationFPFH;
    int cfpfh, norm, ponderation;
    int distTransSize;
    double distTransExpandFactor;
} orc_params;

typedef struct orc_result {
    double R[9];
    double t[3];
    float optError;
    int optComp;
    long long counters[8]; /* inner calls, trans pops, trans subcubes, rot pops, rot cubes, icp calls */
    double seconds_dt, seconds_register;
} orc_result;

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(const float* mxyz, const int* mc, const float* mfpfh, int Nm,
                    const float* dxyz, const int* dc, const float* dfpfh, int Nd, const orc_params* p);
void orc_destroy(orc_ctx*);
void orc_set_nd(orc_ctx*, int nd);

/* DT3D::Build (jly_3ddt.cpp:897) + assignCellColor (jly_goicp.cpp:951) */
double orc_build_dt(orc_ctx*);
void orc_dt_info(orc_ctx*, double* out8);
void orc_dt_download(orc_ctx*, float* dist, short* off, int* nearest, int* cellc);
/* overwrite the grid (used to feed the search with a DT produced elsewhere, e.g. by the CUDA build) */
void orc_dt_upload(orc_ctx*, const float* dist, const int* nearest);
/* DT3D::Distance (jly_3ddt.cpp:1139) */
void orc_dt_distance(orc_ctx*, const double* xyz, int n, float* out, int* cell);

/* GoICP::Initialize (jly_goicp.cpp:180) */
void orc_initialize(orc_ctx*);
void orc_get_weights(orc_ctx*, float* w);
void orc_get_maxrotdis(orc_ctx*, float* out);
float orc_get_ssethresh(orc_ctx*);
int orc_get_inliernum(orc_ctx*);

/* GoICP::InnerBnB (jly_goicp.cpp:286) called as OuterBnB does (:750-768,:861) */
float orc_inner_bnb(orc_ctx*, const float* R, int level, float optError, float* tnode);
/* pure leaf-level bounds of n CHILD translation cubes (x,y,z,w) under rotation R (jly_goicp.cpp:331-550):
 * ub/lb INCLUDING the corner terms; incomp[2n]/fpfh[2n] = (min,max) over the 8 corners (0 when disabled). */
void orc_eval_leaf(orc_ctx*, const float* R, int level, const float* tcube, int n,
                   float* ub, float* lb, int* incomp, int* fpfh);
/* trimmed-error inclusion set per cube (intro_select's contract, jly_goicp.cpp:384-390): mask[n x Nd] (1 = among the inlierNum
 * smallest residuals, ties at the threshold value in ascending index order), resid[n x Nd] (may be NULL) */
void orc_eval_inclusion(orc_ctx*, const float* R, int level, const float* tcube, int n, unsigned char* mask, float* resid);
/* GoICP::ICP (jly_goicp.cpp:102) -> ICP3D::Run (jly_icp3d.hpp:197) */
float orc_icp(orc_ctx*, double* R, double* t, int* corr);
/* GoICP::Register (jly_goicp.cpp:878) incl. BuildDT when not built. trace = the reference's "Error*:" lines. */
void orc_register(orc_ctx*, int nd_downsampled, orc_result* out, char* trace, int trace_cap);

/* Transformation (transformation.cpp) */
double orc_normalize(double* xyz, int n, double* mean);               /* :311 */
void orc_scale(double* xyz, int n, double scale);                     /* :355 */
void orc_round6(const double* xyz, int n, float* out);                /* :340 + jly_main.cpp:272 text round trip */
void orc_rescale_translation(double scale, const double* meanT, const double* meanS, const double* R,
                             const double* t, double* out3);          /* :403-412 */
void orc_apply_rigid(const double* xyz, int n, const double* R, const double* t, double* out); /* :485-497 */
float orc_rmsd(const double* a, const double* b, int n);              /* :453-464 */

#ifdef __cplusplus
}
#endif
#endif
