// goicp_dropin.hpp -- the reference's C++ class surface (jly_goicp.h:112-218, jly_3ddt.h:123-139,
// transformation.hpp:38-68) as thin, header-only wrappers over the C ABI of libgoicp_b200.so (goicp_b200.h).
//
// A caller written against the reference (jly_main.cpp:61-156 or the demo harness of READMEGo-ICP.md) keeps its code:
//     GoICP goicp;  goicp.pModel = ...; goicp.Nm = ...; goicp.pData = ...; goicp.Nd = ...;
//     goicp.MSEThresh = ...; goicp.initNodeRot.a = ...; goicp.dt.SIZE = ...;  goicp.BuildDT();  goicp.Register();
//     goicp.optR / goicp.optT / goicp.optError / goicp.optComp
// Differences from the reference, all on purpose:
//   * POINT3D keeps the reference's fields but `cfpfh` stays a std::vector<float> of 41 bins (jly_goicp.h:47-56); arrays
//     are owned by the caller exactly as there.
//   * optR / optT are plain row-major double arrays (the reference uses A. Geiger's Matrix, `val[i][j]`); the
//     accessor optR_at(i,j) mirrors optR.val[i][j].
//   * every numeric step runs on the GPU; there is no CPU fallback: construction throws std::runtime_error when
//     no CUDA device is usable (the reference exits on unopenable files, this is the analogous hard failure).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "goicp_b200.h"

#define PI 3.1415926536     // jly_goicp.h:44
#define SQRT3 1.732050808   // jly_goicp.h:45

typedef struct _POINT3D {   // jly_goicp.h:47-56
    float x, y, z;
    int c;
    int neighbors;
    float density;
    std::vector<float> cfpfh;
} POINT3D;

typedef struct _ROTNODE { float a, b, c, w; float ub, lb; int l; } ROTNODE;   // jly_goicp.h:59-73
typedef struct _TRANSNODE { float x, y, z, w; float ub, lb; } TRANSNODE;      // jly_goicp.h:75-87

namespace goicp_b200_detail {
inline void check(goicp_handle h, goicp_status s, const char* what) {
    if (s != GOICP_OK) throw std::runtime_error(std::string(what) + ": " + goicp_last_error(h));
}
}  // namespace goicp_b200_detail

// DT3D (jly_3ddt.h:123-139): public geometry + Build/Distance.  Owned by a GoICP object (goicp.dt) or stand-alone.
class DT3D {
public:
    int SIZE = 300;                 // jly_3ddt.cpp:893
    double scale = 0, expandFactor = 2.0;
    double xMin = 0, xMax = 0, yMin = 0, yMax = 0, zMin = 0, zMax = 0;

    DT3D() {}
    ~DT3D() { if (own_) goicp_destroy(h_); }
    // DT3D::Build (jly_3ddt.cpp:897)
    void Build(double* x, double* y, double* z, int num) {
        ensure();
        std::vector<float> xyz(3 * (size_t)num);
        for (int i = 0; i < num; i++) { xyz[3 * i] = (float)x[i]; xyz[3 * i + 1] = (float)y[i]; xyz[3 * i + 2] = (float)z[i]; }
        goicp_params p; goicp_params_default(&p); p.distTransSize = SIZE; p.distTransExpandFactor = expandFactor;
        goicp_b200_detail::check(h_, goicp_set_params(h_, &p), "set_params");
        goicp_b200_detail::check(h_, goicp_set_model(h_, xyz.data(), nullptr, nullptr, num), "set_model");
        goicp_b200_detail::check(h_, goicp_set_data(h_, xyz.data(), nullptr, nullptr, num > 1 ? 1 : num), "set_data");
        goicp_dt_info info;
        goicp_b200_detail::check(h_, goicp_build_dt(h_, &info), "build_dt");
        take(info);
    }
    // DT3D::Distance (jly_3ddt.cpp:1139); cx,cy,cz receive the unclamped voxel
    float Distance(double x, double y, double z, int& cx, int& cy, int& cz) {
        double q[3] = {x, y, z}; float d = 0; int32_t c[3];
        goicp_b200_detail::check(h_, goicp_dt_distance(h_, q, 1, &d, c), "dt_distance");
        cx = c[0]; cy = c[1]; cz = c[2];
        return d;
    }
    void attach(goicp_handle h) { h_ = h; own_ = false; }
    void take(const goicp_dt_info& i) { xMin = i.xMin; xMax = i.xMax; yMin = i.yMin; yMax = i.yMax; zMin = i.zMin; zMax = i.zMax; scale = i.scale; }

private:
    void ensure() { if (!h_) { goicp_b200_detail::check(nullptr, goicp_create(&h_, 0, nullptr), "goicp_create"); own_ = true; } }
    goicp_handle h_ = nullptr;
    bool own_ = false;
};

// GoICP (jly_goicp.h:112-218)
class GoICP {
public:
    int Nm = 0, Nd = 0;
    POINT3D *pModel = nullptr, *pData = nullptr;
    ROTNODE initNodeRot{}, optNodeRot{};
    TRANSNODE initNodeTrans{}, optNodeTrans{};
    DT3D dt;
    float MSEThresh = 0.001f, SSEThresh = 0, icpThresh = 0;
    float optError = 0;
    double optR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, optT[3] = {0, 0, 0};
    int optComp = 0;
    float trimFraction = 0;
    int inlierNum = 0;
    bool doTrim = true;                       // GoICP() :54
    float regularization = 0, regularizationNeighbors = 0, regularizationFPFH = 0;
    int norm = 2, ponderation = 0, cfpfh = 0;
    long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // node counters of the last Register (goicp_result)

    GoICP(int device = 0) {
        goicp_b200_detail::check(nullptr, goicp_create(&h_, device, nullptr), "goicp_create");
        initNodeRot.a = initNodeRot.b = initNodeRot.c = (float)-PI; initNodeRot.w = (float)(2 * PI); initNodeRot.l = 0;   // :47-51
        initNodeRot.lb = 0; initNodeTrans.lb = 0;
        dt.SIZE = 300; dt.expandFactor = 2.0;
        dt.attach(h_);
    }
    ~GoICP() { goicp_destroy(h_); }
    GoICP(const GoICP&) = delete;
    GoICP& operator=(const GoICP&) = delete;

    double optR_at(int i, int j) const { return optR[3 * i + j]; }
    double optT_at(int i) const { return optT[i]; }

    // GoICP::BuildDT (jly_goicp.cpp:79)
    void BuildDT() {
        push_params();
        push_cloud(pModel, Nm, true);
        push_cloud(pData, Nd, false);
        goicp_dt_info info;
        goicp_b200_detail::check(h_, goicp_build_dt(h_, &info), "build_dt");
        dt.take(info);
        ndAtBuild_ = Nd;
    }
    // GoICP::Register (jly_goicp.cpp:878) = Initialize + OuterBnB + Clear
    float Register() {
        push_params();
        if (Nd != ndAtBuild_) goicp_b200_detail::check(h_, goicp_set_nd(h_, Nd), "set_nd");   // `goicp.Nd = NdDownsampled` (jly_main.cpp:114)
        goicp_result r;
        goicp_b200_detail::check(h_, goicp_register(h_, &r), "register");
        for (int k = 0; k < 9; k++) optR[k] = r.R[k];
        for (int k = 0; k < 3; k++) optT[k] = r.t[k];
        optError = r.optError; optComp = r.optComp;
        for (int k = 0; k < 8; k++) counters[k] = r.counters[k];
        goicp_get_thresholds(h_, &SSEThresh, &inlierNum);
        return optError;
    }
    // the improvement trace OuterBnB prints (jly_goicp.cpp:627-839)
    const char* Trace() const { return goicp_last_trace(h_); }
    goicp_handle handle() const { return h_; }

private:
    void push_params() {
        goicp_params p;
        p.MSEThresh = MSEThresh;
        p.rotMinX = initNodeRot.a; p.rotMinY = initNodeRot.b; p.rotMinZ = initNodeRot.c; p.rotWidth = initNodeRot.w;
        p.transMinX = initNodeTrans.x; p.transMinY = initNodeTrans.y; p.transMinZ = initNodeTrans.z; p.transWidth = initNodeTrans.w;
        p.trimFraction = doTrim ? trimFraction : 0.0f;   // readConfig: trimFraction < 0.001 <=> doTrim = false (jly_main.cpp:259)
        p.regularization = regularization; p.regularizationNeighbors = regularizationNeighbors; p.regularizationFPFH = regularizationFPFH;
        p.cfpfh = cfpfh; p.norm = norm; p.ponderation = ponderation;
        p.distTransSize = dt.SIZE; p.distTransExpandFactor = dt.expandFactor;
        goicp_b200_detail::check(h_, goicp_set_params(h_, &p), "set_params");
    }
    void push_cloud(const POINT3D* pts, int n, bool model) {
        std::vector<float> xyz(3 * (size_t)n), f;
        std::vector<int32_t> c(n);
        bool haveF = n > 0 && pts[0].cfpfh.size() == 41;
        if (haveF) f.resize(41 * (size_t)n);
        for (int i = 0; i < n; i++) {
            xyz[3 * i] = pts[i].x; xyz[3 * i + 1] = pts[i].y; xyz[3 * i + 2] = pts[i].z; c[i] = pts[i].c;
            if (haveF) for (int k = 0; k < 41; k++) f[41 * (size_t)i + k] = pts[i].cfpfh.size() == 41 ? pts[i].cfpfh[k] : 0.f;
        }
        goicp_status s = model ? goicp_set_model(h_, xyz.data(), c.data(), haveF ? f.data() : nullptr, n)
                               : goicp_set_data(h_, xyz.data(), c.data(), haveF ? f.data() : nullptr, n);
        goicp_b200_detail::check(h_, s, model ? "set_model" : "set_data");
    }
    goicp_handle h_ = nullptr;
    int ndAtBuild_ = -1;
};

// Transformation (transformation.hpp:38-68): the numeric members; file parsing/writing stays with the caller.
struct point4D { double x, y, z; int c; };   // transformation.hpp:22-34
class Transformation {
public:
    Transformation(int device = 0) { goicp_b200_detail::check(nullptr, goicp_create(&h_, device, nullptr), "goicp_create"); }
    ~Transformation() { goicp_destroy(h_); }
    Transformation(const Transformation&) = delete;
    Transformation& operator=(const Transformation&) = delete;
    // normalizeMolCloud (transformation.cpp:311): centres the cloud in place, fills mean, returns the max norm
    double normalizeMolCloud(std::vector<point4D>& cloud, double& xm, double& ym, double& zm) {
        std::vector<double> a = flat(cloud); double mean[3], mx = 0;
        goicp_b200_detail::check(h_, goicp_normalize_cloud(h_, a.data(), (int)cloud.size(), mean, &mx), "normalize");
        unflat(a, cloud); xm = mean[0]; ym = mean[1]; zm = mean[2];
        return mx;
    }
    // scaleCloud (:355)
    void scaleCloud(std::vector<point4D>& cloud, double scale) {
        std::vector<double> a = flat(cloud);
        goicp_b200_detail::check(h_, goicp_scale_cloud(h_, a.data(), (int)cloud.size(), scale), "scale");
        unflat(a, cloud);
    }
    // rescaleCloud (:403-412): the rescaled translation written to <output>_rescaled.txt
    void rescaleTranslation(double scale, const double meanT[3], const double meanS[3], const double R[9], const double t[3], double out[3]) {
        goicp_b200_detail::check(h_, goicp_rescale_translation(h_, scale, meanT, meanS, R, t, out), "rescale");
    }
    // applyTransformationProtein (:485-497) on in-memory atoms
    std::vector<point4D> applyTransformation(const std::vector<point4D>& pts, const double R[9], const double t[3]) {
        std::vector<double> a = flat(pts), o(a.size());
        goicp_b200_detail::check(h_, goicp_apply_rigid(h_, a.data(), (int)pts.size(), R, t, o.data()), "apply_rigid");
        std::vector<point4D> out(pts); unflat(o, out);
        return out;
    }
    // computeRMSD (:453-464) over already selected backbone atoms
    float computeRMSD(const std::vector<point4D>& aligned, const std::vector<point4D>& transformed) {
        std::vector<double> a = flat(aligned), b = flat(transformed); float r = 0;
        goicp_b200_detail::check(h_, goicp_rmsd(h_, a.data(), b.data(), (int)aligned.size(), &r), "rmsd");
        return r;
    }

private:
    static std::vector<double> flat(const std::vector<point4D>& c) {
        std::vector<double> a(3 * c.size());
        for (size_t i = 0; i < c.size(); i++) { a[3 * i] = c[i].x; a[3 * i + 1] = c[i].y; a[3 * i + 2] = c[i].z; }
        return a;
    }
    static void unflat(const std::vector<double>& a, std::vector<point4D>& c) {
        for (size_t i = 0; i < c.size(); i++) { c[i].x = a[3 * i]; c[i].y = a[3 * i + 1]; c[i].z = a[3 * i + 2]; }
    }
    goicp_handle h_ = nullptr;
};
