// goicp_dropin.hpp -- the reference's C++ class surface (jly_goicp.h:112-218, jly_3ddt.h:123-139, matrix.h:45-133 as far as the
// pipeline uses it, transformation.hpp:38-68, ConfigMap.hpp:13-53) as thin, header-only wrappers over the C ABI of libgoicp_b200.so
// (goicp_b200.h).  include/compat/ holds one-line headers with the reference's own file names (jly_goicp.h, jly_3ddt.h, matrix.h,
// ConfigMap.hpp, transformation.hpp) that include this file, so the reference's UNMODIFIED jly_main.cpp compiles against
// `-Iinclude/compat -Iinclude` and links against libgoicp_b200.so alone (examples/Makefile: GoICP_ref_main).
//
//     GoICP goicp;  goicp.pModel = ...; goicp.Nm = ...; goicp.pData = ...; goicp.Nd = ...;
//     goicp.MSEThresh = ...; goicp.initNodeRot.a = ...; goicp.dt.SIZE = ...;  goicp.BuildDT();  goicp.Register();
//     goicp.optR.val[i][j] / goicp.optT.val[i][0] / goicp.optError / goicp.optComp;   cout << goicp.optR;
//
// What differs from the reference, on purpose:
//   * every numeric step runs on the GPU and there is no CPU fallback: constructing GoICP / Transformation / DT3D::Build throws
//     std::runtime_error when no CUDA device is usable (the reference exits on unopenable files; this is the analogous hard failure);
//   * GoICP::InnerBnB takes the rotation of the cube through the public member R_cur (9 floats, row-major) instead of reading the
//     pre-rotated scratch cloud pDataTemp that the reference's OuterBnB fills in place (jly_goicp.cpp:750-756);
//   * DT3D::emptyCells / cellPoints are filled on first use (a download of the index map), not at Build time.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "goicp_b200.h"

#define PI 3.1415926536     // jly_goicp.h:44
#define SQRT3 1.732050808   // jly_goicp.h:45
#define MAXROTLEVEL 20      // jly_goicp.h:95

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

// The slice of A. Geiger's Matrix (matrix.h:45-133) the pipeline touches: row pointers `val[i][j]`, sizes m x n, construction from
// a row-major array, copies, and the "%12.7f " stream format of matrix.cpp:812-827 that the output files are written with.
class Matrix {
public:
    double** val = nullptr;
    int32_t m = 0, n = 0;
    Matrix() {}
    Matrix(int32_t rows, int32_t cols) { allocate(rows, cols); }
    Matrix(int32_t rows, int32_t cols, const double* rowMajor) { allocate(rows, cols); for (int32_t i = 0; i < m; i++) for (int32_t j = 0; j < n; j++) val[i][j] = rowMajor[i * n + j]; }
    Matrix(const Matrix& o) { allocate(o.m, o.n); for (int32_t i = 0; i < m; i++) for (int32_t j = 0; j < n; j++) val[i][j] = o.val[i][j]; }
    Matrix& operator=(const Matrix& o) {
        if (this != &o) { release(); allocate(o.m, o.n); for (int32_t i = 0; i < m; i++) for (int32_t j = 0; j < n; j++) val[i][j] = o.val[i][j]; }
        return *this;
    }
    ~Matrix() { release(); }
    static Matrix eye(int32_t k) { Matrix r(k, k); for (int32_t i = 0; i < k; i++) r.val[i][i] = 1; return r; }
    void getData(double* rowMajor) const { for (int32_t i = 0; i < m; i++) for (int32_t j = 0; j < n; j++) rowMajor[i * n + j] = val[i][j]; }

private:
    void allocate(int32_t rows, int32_t cols) {
        m = rows < 0 ? -rows : rows; n = cols < 0 ? -cols : cols;
        if (m == 0 || n == 0) { val = nullptr; return; }
        val = (double**)malloc(sizeof(double*) * m);
        val[0] = (double*)calloc((size_t)m * n, sizeof(double));
        for (int32_t i = 1; i < m; i++) val[i] = val[i - 1] + n;
    }
    void release() { if (val) { free(val[0]); free(val); val = nullptr; } m = n = 0; }
};
inline std::ostream& operator<<(std::ostream& out, const Matrix& M) {
    if (M.m == 0 || M.n == 0) return out << "[empty matrix]";
    char buffer[64];
    for (int32_t i = 0; i < M.m; i++) {
        for (int32_t j = 0; j < M.n; j++) { snprintf(buffer, sizeof buffer, "%12.7f ", M.val[i][j]); out << buffer; }
        if (i < M.m - 1) out << std::endl;
    }
    return out;
}

struct EMPTYCELL { int x, y, z; };                         // jly_3ddt.h:117-120: the closest occupied cell of a voxel
struct CELL { int c = -2; std::vector<int> points; };      // jly_3ddt.h:112-115: model points of a cell; c = -2 empty, -1 mixed colours, else the colour

// DT3D (jly_3ddt.h:123-139): public geometry + Build / Distance.  Owned by a GoICP object (goicp.dt) or stand-alone.
class DT3D {
    template <class T> struct Row { T* p; T& operator[](int x) const { return p[x]; } };
    template <class T> struct Plane { T* p; int S; Row<T> operator[](int y) const { return Row<T>{p + (size_t)y * S}; } };
    template <class T> struct Vol {   // indexed [z][y][x] like the reference's T*** arrays; filled on first use
        DT3D* dt = nullptr; std::vector<T> v;
        Plane<T> operator[](int z) { dt->load_cells(); return Plane<T>{v.data() + (size_t)z * dt->SIZE * dt->SIZE, dt->SIZE}; }
    };

public:
    int SIZE = 300;                 // jly_3ddt.cpp:893
    double scale = 0, expandFactor = 2.0;
    double xMin = 0, xMax = 0, yMin = 0, yMax = 0, zMin = 0, zMax = 0;
    Vol<EMPTYCELL> emptyCells;      // emptyCells[z][y][x] (jly_3ddt.cpp:999-1136)
    Vol<CELL> cellPoints;           // cellPoints[z][y][x] (jly_3ddt.cpp:984-985, assignCellColor jly_goicp.cpp:951-969)

    DT3D() { emptyCells.dt = this; cellPoints.dt = this; }
    ~DT3D() { if (own_) goicp_destroy(h_); }
    DT3D(const DT3D&) = delete;
    DT3D& operator=(const DT3D&) = delete;
    // DT3D::Build (jly_3ddt.cpp:897)
    void Build(double* x, double* y, double* z, int num) {
        if (!h_) { goicp_b200_detail::check(nullptr, goicp_create(&h_, 0, nullptr), "goicp_create"); own_ = true; }
        std::vector<float> xyz(3 * (size_t)num);
        for (int i = 0; i < num; i++) { xyz[3 * i] = (float)x[i]; xyz[3 * i + 1] = (float)y[i]; xyz[3 * i + 2] = (float)z[i]; }
        goicp_params p; goicp_params_default(&p); p.distTransSize = SIZE; p.distTransExpandFactor = expandFactor;
        goicp_b200_detail::check(h_, goicp_set_params(h_, &p), "set_params");
        goicp_b200_detail::check(h_, goicp_set_model(h_, xyz.data(), nullptr, nullptr, num), "set_model");
        goicp_b200_detail::check(h_, goicp_set_data(h_, xyz.data(), nullptr, nullptr, num > 1 ? 1 : num), "set_data");
        goicp_dt_info info;
        goicp_b200_detail::check(h_, goicp_build_dt(h_, &info), "build_dt");
        take(info, xyz.data(), num);
    }
    // DT3D::Distance (jly_3ddt.cpp:1139); cx,cy,cz receive the unclamped voxel
    float Distance(double x, double y, double z, int& cx, int& cy, int& cz) {
        double q[3] = {x, y, z}; float d = 0; int32_t c[3];
        goicp_b200_detail::check(h_, goicp_dt_distance(h_, q, 1, &d, c), "dt_distance");
        cx = c[0]; cy = c[1]; cz = c[2];
        return d;
    }
    void attach(goicp_handle h) { h_ = h; own_ = false; }
    void take(const goicp_dt_info& i, const float* modelXyz, int num) {
        xMin = i.xMin; xMax = i.xMax; yMin = i.yMin; yMax = i.yMax; zMin = i.zMin; zMax = i.zMax; scale = i.scale;
        model_.assign(modelXyz, modelXyz + 3 * (size_t)num); cellsLoaded_ = false;
    }
    // the index map and the per-cell point lists, as the reference keeps them after Build
    void load_cells() {
        if (cellsLoaded_) return;
        const size_t S3 = (size_t)SIZE * SIZE * SIZE;
        std::vector<int32_t> nearest(3 * S3), cellc(S3);
        goicp_b200_detail::check(h_, goicp_dt_download(h_, nullptr, nearest.data(), cellc.data()), "dt_download");
        emptyCells.v.resize(S3); cellPoints.v.assign(S3, CELL());
        for (size_t i = 0; i < S3; i++) { emptyCells.v[i] = EMPTYCELL{nearest[3 * i], nearest[3 * i + 1], nearest[3 * i + 2]}; cellPoints.v[i].c = cellc[i]; }
        for (size_t k = 0; k < model_.size() / 3; k++) {   // ROUND((p - min) * scale), jly_3ddt.cpp:30,976-995
            const int x = (int)(((double)model_[3 * k] - xMin) * scale + 0.5), y = (int)(((double)model_[3 * k + 1] - yMin) * scale + 0.5), z = (int)(((double)model_[3 * k + 2] - zMin) * scale + 0.5);
            if (x < 0 || x >= SIZE || y < 0 || y >= SIZE || z < 0 || z >= SIZE) continue;
            cellPoints.v[((size_t)z * SIZE + y) * SIZE + x].points.push_back((int)k);
        }
        cellsLoaded_ = true;
    }

private:
    goicp_handle h_ = nullptr;
    bool own_ = false, cellsLoaded_ = false;
    std::vector<float> model_;
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
    Matrix optR = Matrix::eye(3), optT = Matrix(3, 1);
    int optComp = 0;
    float trimFraction = 0;
    int inlierNum = 0;
    bool doTrim = true;                       // GoICP() :54
    float regularization = 0, regularizationNeighbors = 0, regularizationFPFH = 0;
    int norm = 2, ponderation = 0, cfpfh = 0;
    float** maxRotDis = nullptr;              // [MAXROTLEVEL][Nd] after Initialize (jly_goicp.cpp:194-206)
    float* weights = nullptr;                 // [Nd] after Initialize (:262)
    float R_cur[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // rotation the next InnerBnB call is made under (see the header comment)
    long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // node counters of the last Register / OuterBnB (goicp_result)

    GoICP(int device = 0) {
        goicp_b200_detail::check(nullptr, goicp_create(&h_, device, nullptr), "goicp_create");
        initNodeRot.a = initNodeRot.b = initNodeRot.c = (float)-PI; initNodeRot.w = (float)(2 * PI); initNodeRot.l = 0;   // :47-51
        initNodeRot.lb = 0; initNodeTrans.lb = 0;
        dt.SIZE = 300; dt.expandFactor = 2.0;
        dt.attach(h_);
    }
    ~GoICP() { Clear(); goicp_destroy(h_); }
    GoICP(const GoICP&) = delete;
    GoICP& operator=(const GoICP&) = delete;

    double optR_at(int i, int j) const { return optR.val[i][j]; }
    double optT_at(int i) const { return optT.val[i][0]; }

    // GoICP::BuildDT (jly_goicp.cpp:79)
    void BuildDT() {
        push_params();
        std::vector<float> mxyz = push_cloud(pModel, Nm, true);
        push_cloud(pData, Nd, false);
        goicp_dt_info info;
        goicp_b200_detail::check(h_, goicp_build_dt(h_, &info), "build_dt");
        dt.take(info, mxyz.data(), Nm);
        ndAtBuild_ = Nd;
    }
    // GoICP::Register (jly_goicp.cpp:878) = Initialize + OuterBnB + Clear
    float Register() {
        push_params();
        sync_nd();
        goicp_result r;
        goicp_b200_detail::check(h_, goicp_register(h_, &r), "register");
        take_result(r);
        return optError;
    }
    // GoICP::Initialize (:180-267): norms, maxRotDis[20][Nd], weights, inlierNum, SSEThresh (computed on the device, mirrored here)
    void Initialize() {
        push_params();
        sync_nd();
        goicp_b200_detail::check(h_, goicp_initialize(h_), "initialize");
        goicp_b200_detail::check(h_, goicp_get_thresholds(h_, &SSEThresh, &inlierNum), "get_thresholds");
        Clear();
        rotBuf_.resize((size_t)MAXROTLEVEL * Nd); wBuf_.resize(Nd);
        goicp_b200_detail::check(h_, goicp_get_maxrotdis(h_, rotBuf_.data()), "get_maxrotdis");
        goicp_b200_detail::check(h_, goicp_get_weights(h_, wBuf_.data()), "get_weights");
        maxRotDis = new float*[MAXROTLEVEL];
        for (int l = 0; l < MAXROTLEVEL; l++) maxRotDis[l] = rotBuf_.data() + (size_t)l * Nd;
        weights = wBuf_.data();
        optError = 0; optR = Matrix::eye(3); optT = Matrix(3, 1);   // :240-241
    }
    // GoICP::OuterBnB (:582-876) on the initialised problem
    float OuterBnB() {
        goicp_result r;
        goicp_b200_detail::check(h_, goicp_outer_bnb(h_, &r), "outer_bnb");
        take_result(r);
        return optError;
    }
    // GoICP::InnerBnB (:286-579): translation search under rotation R_cur with the incumbent `optError`; maxRotDisL = one of the
    // rows maxRotDis[level] (lower bound) or NULL (upper bound); the best translation node is returned through nodeTransOut.
    float InnerBnB(float* maxRotDisL, TRANSNODE* nodeTransOut) {
        int32_t level = -1;
        if (maxRotDisL) { for (int l = 0; l < MAXROTLEVEL; l++) if (maxRotDis && maxRotDisL == maxRotDis[l]) level = l; if (level < 0) throw std::runtime_error("InnerBnB: maxRotDisL is not a row of maxRotDis"); }
        float err = 0, tn[4] = {0, 0, 0, 0};
        goicp_b200_detail::check(h_, goicp_inner_bnb(h_, R_cur, &level, &optError, 1, &err, tn, nullptr), "inner_bnb");
        if (nodeTransOut && err < optError) { nodeTransOut->x = tn[0]; nodeTransOut->y = tn[1]; nodeTransOut->z = tn[2]; nodeTransOut->w = tn[3]; }
        return err;
    }
    // GoICP::ICP (:102-178): refine (R_icp, t_icp) in place, return the DT-scored error of the refined pose
    float ICP(Matrix& R_icp, Matrix& t_icp) {
        double R[9], t[3]; float err = 0;
        R_icp.getData(R); t_icp.getData(t);
        goicp_b200_detail::check(h_, goicp_icp(h_, R, t, &err, nullptr), "icp");
        R_icp = Matrix(3, 3, R); t_icp = Matrix(3, 1, t);
        return err;
    }
    // GoICP::Clear (:269-283): the host mirrors of Initialize (device memory belongs to the handle)
    void Clear() { delete[] maxRotDis; maxRotDis = nullptr; weights = nullptr; rotBuf_.clear(); wBuf_.clear(); }
    // the improvement trace OuterBnB prints (jly_goicp.cpp:627-839)
    const char* Trace() const { return goicp_last_trace(h_); }
    goicp_handle handle() const { return h_; }

private:
    void sync_nd() { if (Nd != ndAtBuild_) { goicp_b200_detail::check(h_, goicp_set_nd(h_, Nd), "set_nd"); ndAtBuild_ = Nd; } }   // `goicp.Nd = NdDownsampled` (jly_main.cpp:114)
    void take_result(const goicp_result& r) {
        optR = Matrix(3, 3, r.R); optT = Matrix(3, 1, r.t);
        optError = r.optError; optComp = r.optComp;
        for (int k = 0; k < 8; k++) counters[k] = r.counters[k];
        goicp_get_thresholds(h_, &SSEThresh, &inlierNum);
    }
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
    std::vector<float> push_cloud(const POINT3D* pts, int n, bool model) {
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
        return xyz;
    }
    goicp_handle h_ = nullptr;
    int ndAtBuild_ = -1;
    std::vector<float> rotBuf_, wBuf_;
};

// ConfigMap (ConfigMap.hpp:13-53, ConfigMap.cpp:3-151): key=value lines, '#' comments, CR stripped, tokens split on " =;", lines
// without exactly two tokens ignored, a missing key reads as "" (atof/atoi give 0); exit(-2) when the file cannot be opened.
class ConfigMap {
public:
    ConfigMap() {}
    explicit ConfigMap(const char* config_file) {
        std::ifstream in(config_file);
        if (!in.is_open()) { std::cout << "Unable to open config file '" << config_file << "'" << std::endl; exit(-2); }
        std::string line;
        while (std::getline(in, line)) addLine(line);
    }
    void addLine(std::string line) {
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        std::vector<std::string> tok; std::string cur;
        for (size_t i = 0; i < line.size(); i++) {
            const char ch = line[i];
            if (ch == ' ' || ch == '=' || ch == ';') { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } } else cur += ch;
        }
        if (!cur.empty()) tok.push_back(cur);
        if (tok.size() == 2) addPair(tok[0], tok[1]);
    }
    void addPair(std::string key, std::string value) { mappings[key] = value; }
    const char* get(const char* key) { std::map<std::string, std::string>::const_iterator it = mappings.find(key); return it == mappings.end() ? "" : it->second.c_str(); }
    int getI(const char* key) { return atoi(get(key)); }
    double getF(const char* key) { return atof(get(key)); }
    void print() { for (std::map<std::string, std::string>::const_iterator it = mappings.begin(); it != mappings.end(); ++it) std::cout << "(" << it->first << ")->(" << it->second << ")" << std::endl; }

private:
    std::map<std::string, std::string> mappings;
};

// Transformation (transformation.hpp:38-68).  The arithmetic (centroid / max norm, scaling, the rescaled translation, the rigid
// transform of the protein atoms, the RMSD) runs on the GPU through the C ABI; the text formats are read and written here with
// the reference's conventions (default ostream precision = 6 significant digits, to_string = 6 decimals, the getline / >> pairing
// of the mol2 reader that drops the first record it cannot parse).
struct point { double x, y, z; };                                                    // transformation.hpp:22-24
struct point4D { double x, y, z; int c; bool operator<(const point4D& a) const { return y < a.y; } };   // transformation.hpp:26-34
class Transformation {
public:
    Transformation(int device = 0) { goicp_b200_detail::check(nullptr, goicp_create(&h_, device, nullptr), "goicp_create"); }
    ~Transformation() { goicp_destroy(h_); }
    Transformation(const Transformation&) = delete;
    Transformation& operator=(const Transformation&) = delete;

    // colour code of an atom name (the `properties` enum, transformation.hpp:36); unknown names are OG (transformation.cpp:46)
    static int colour_of(const std::string& type) {
        static const char* names[9] = {"OG", "N", "O", "NZ", "CZ", "CA", "DU", "OD1", "C"};
        static const int codes[9] = {8204959, 30894, 15219528, 15231913, 4646984, 16741671, 7566712, 0, 1};
        for (int k = 0; k < 9; k++) if (type == names[k]) return codes[k];
        return codes[0];
    }
    // readMolFile (transformation.cpp:282-306): from the @<TRIPOS>ATOM line on, each getline is followed by a formatted read of
    // the NEXT record's (id, name, x, y, z); the last (failed) read is dropped
    std::vector<point4D> readMolFile(std::ifstream& file) {
        std::vector<point4D> cloud;
        if (!file.is_open()) return cloud;
        std::string line; bool inAtoms = false;
        while (std::getline(file, line)) {
            if (line.find("@<TRIPOS>ATOM") != std::string::npos) inAtoms = true;
            if (!inAtoms) continue;
            std::string id, name; point4D p{};
            file >> id >> name >> p.x >> p.y >> p.z;
            p.c = colour_of(name);
            cloud.push_back(p);
        }
        if (!cloud.empty()) cloud.pop_back();
        return cloud;
    }
    // normalizeMolCloud (:311-335): centres the cloud in place, fills the mean, returns the max norm
    double normalizeMolCloud(std::vector<point4D>& cloud, double& xm, double& ym, double& zm) {
        std::vector<double> a = flat(cloud); double mean[3], mx = 0;
        goicp_b200_detail::check(h_, goicp_normalize_cloud(h_, a.data(), (int)cloud.size(), mean, &mx), "normalize");
        unflat(a, cloud); xm = mean[0]; ym = mean[1]; zm = mean[2];
        return mx;
    }
    // scaleCloud (:355-361)
    void scaleCloud(std::vector<point4D>& cloud, double scale) {
        std::vector<double> a = flat(cloud);
        goicp_b200_detail::check(h_, goicp_scale_cloud(h_, a.data(), (int)cloud.size(), scale), "scale");
        unflat(a, cloud);
    }
    // writeNormalizedMolCloudFile (:340-350): "N" then "x y z c" per point, 6 significant digits
    void writeNormalizedMolCloudFile(std::ofstream& file, std::vector<point4D>& cloud) {
        if (file.is_open()) {
            file << cloud.size() << std::endl;
            for (size_t i = 0; i < cloud.size(); i++) { std::ostringstream row; row << cloud[i].x << " " << cloud[i].y << " " << cloud[i].z << " " << cloud[i].c << std::endl; file << row.str(); }
        }
        file.close();
    }
    // rescaleCloud (:403-417): <output>_rescaled.txt with t' = -R*meanSource + scale*t + meanTarget
    void rescaleCloud(std::ofstream& file, std::vector<point4D>& cloud, double scale, double xMeanTemp, double yMeanTemp, double zMeanTemp, double time, double error,
                      double xMean, double yMean, double zMean, double rot[3][3], double xTrans, double yTrans, double zTrans) {
        (void)cloud;
        if (file.is_open()) {
            const double meanT[3] = {xMeanTemp, yMeanTemp, zMeanTemp}, meanS[3] = {xMean, yMean, zMean}, t[3] = {xTrans, yTrans, zTrans};
            double R[9], out[3];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = rot[i][j];
            rescaleTranslation(scale, meanT, meanS, R, t, out);
            file << "Time: " << time << std::endl;
            file << "Rotation Matrix:" << std::endl;
            for (int i = 0; i < 3; i++) file << "   " << rot[i][0] << "   " << rot[i][1] << "   " << rot[i][2] << std::endl;
            file << "Translation Vector:" << std::endl << "   " << out[0] << std::endl << "   " << out[1] << std::endl << "   " << out[2] << std::endl;
            file << "Error: " << error << std::endl;
        }
        file.close();
    }
    void rescaleTranslation(double scale, const double meanT[3], const double meanS[3], const double R[9], const double t[3], double out[3]) {
        goicp_b200_detail::check(h_, goicp_rescale_translation(h_, scale, meanT, meanS, R, t, out), "rescale");
    }
    // readOutput (:120-139): Time / Rotation Matrix / Translation Vector / Error of a result file
    void readOutput(std::ifstream& file, double rot[3][3], double tra[3][1], double& time, double& error) {
        if (!file.is_open()) return;
        std::string s;
        file >> s >> time >> s >> s;
        for (int i = 0; i < 3; i++) file >> rot[i][0] >> rot[i][1] >> rot[i][2];
        file >> s >> s;
        for (int i = 0; i < 3; i++) file >> tra[i][0];
        file >> s >> error >> s >> s;
    }
    // getAtomBlock (:423-448): the backbone atoms (C, CA, N, O) of a mol2 file, read with the same getline / >> pairing
    std::vector<point4D> getAtomBlock(std::ifstream& file) {
        std::vector<point4D> cloud;
        if (!file.is_open()) return cloud;
        std::string line; bool inAtoms = false;
        while (std::getline(file, line)) {
            if (line == "@<TRIPOS>ATOM") inAtoms = true;
            if (!inAtoms) continue;
            std::string id, name; point4D p{};
            file >> id >> name >> p.x >> p.y >> p.z;
            p.c = colour_of(name);
            if (p.c == 1 || p.c == 16741671 || p.c == 30894 || p.c == 15219528) cloud.push_back(p);
        }
        return cloud;
    }
    // computeRMSD (:453-464) between the backbone atoms of two mol2 files
    float computeRMSD(std::ifstream& alignedFile, std::ifstream& rotFile) {
        std::vector<point4D> a = getAtomBlock(alignedFile), b = getAtomBlock(rotFile);
        if (a.empty() || b.size() < a.size()) return 0.f;
        b.resize(a.size());
        return computeRMSD(a, b);
    }
    float computeRMSD(const std::vector<point4D>& aligned, const std::vector<point4D>& transformed) {
        std::vector<double> a = flat(aligned), b = flat(transformed); float r = 0;
        goicp_b200_detail::check(h_, goicp_rmsd(h_, a.data(), b.data(), (int)aligned.size(), &r), "rmsd");
        return r;
    }
    // applyTransformationProtein (:469-539): the rescaled transform of cavitiesR/similar<pair>.txt applied to every atom of a protein
    // mol2 file; the atom records are re-emitted tab-separated with to_string coordinates, everything else is copied
    void applyTransformationProtein(std::ofstream& output, std::string protein, int pair) {
        std::ifstream resultFile("cavitiesR/similar" + std::to_string(pair) + ".txt");
        double time = 0, error = 0, rot[3][3] = {}, tra[3][1] = {};
        readOutput(resultFile, rot, tra, time, error);
        std::ifstream atomsIn(protein);
        std::vector<point4D> atoms = readMolFile(atomsIn);
        double R[9], t[3] = {tra[0][0], tra[1][0], tra[2][0]};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = rot[i][j];
        std::vector<point4D> moved = atoms.empty() ? atoms : applyTransformation(atoms, R, t);
        std::ifstream in(protein);
        std::string line; size_t count = 0; bool inAtoms = false, afterAtoms = false;
        if (in.is_open()) {
            while (std::getline(in, line)) {
                if (line == "@<TRIPOS>ATOM") { output << line << std::endl; inAtoms = true; }
                if (!inAtoms) { output << line << std::endl; continue; }
                if (afterAtoms) { output << line << std::endl; continue; }
                std::string f[9];
                in >> f[0];
                if (f[0] == "@<TRIPOS>BOND") { output << f[0] << std::endl; afterAtoms = true; continue; }
                for (int k = 1; k < 9; k++) in >> f[k];
                if (count < moved.size()) { f[2] = std::to_string(moved[count].x); f[3] = std::to_string(moved[count].y); f[4] = std::to_string(moved[count].z); }
                else { f[2] = std::to_string(0.0); f[3] = f[2]; f[4] = f[2]; }
                output << f[0]; for (int k = 1; k < 9; k++) output << "\t" << f[k]; output << std::endl;
                count++;
            }
        }
        in.close();
        output.close();
    }
    // the rigid transform on in-memory atoms (:485-497)
    std::vector<point4D> applyTransformation(const std::vector<point4D>& pts, const double R[9], const double t[3]) {
        std::vector<double> a = flat(pts), o(a.size());
        goicp_b200_detail::check(h_, goicp_apply_rigid(h_, a.data(), (int)pts.size(), R, t, o.data()), "apply_rigid");
        std::vector<point4D> out(pts); unflat(o, out);
        return out;
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
