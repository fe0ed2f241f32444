// goicp_demo -- a caller written against the reference's C++ surface (GoICP / POINT3D / config.txt keys), built on
// include/goicp_dropin.hpp.  It is the upstream Go-ICP demo main (READMEGo-ICP.md:20-51; the fork's jly_main.cpp:54-179
// without the mol2 pre/post-processing):
//     goicp_demo <model.txt> <data.txt> <NdDownsampled> <config.txt> <output.txt>
// Point files: N, then N lines "x y z" (READMEGo-ICP.md:47-50).  config.txt: key=value lines, '#' comments
// (ConfigMap.cpp:3-151 semantics: tokens split on " =;", lines without exactly two tokens are ignored, missing
// keys read as 0).  Output: time, R (3 rows), t (3 rows) as the upstream demo writes them (demo/output.txt).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

#include "goicp_dropin.hpp"

static std::map<std::string, std::string> read_config(const char* path) {
    std::ifstream in(path);
    if (!in.is_open()) { std::cerr << "Unable to open config file " << path << std::endl; exit(-2); }   // ConfigMap.cpp:12
    std::map<std::string, std::string> m;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        size_t h = line.find('#');
        if (h != std::string::npos) line = line.substr(0, h);
        std::vector<std::string> tok; std::string cur;
        for (char ch : line) { if (ch == ' ' || ch == '=' || ch == ';' || ch == '\t') { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } } else cur += ch; }
        if (!cur.empty()) tok.push_back(cur);
        if (tok.size() == 2) m[tok[0]] = tok[1];
    }
    return m;
}
static double getF(const std::map<std::string, std::string>& m, const char* k) { auto it = m.find(k); return it == m.end() ? 0.0 : atof(it->second.c_str()); }
static int getI(const std::map<std::string, std::string>& m, const char* k) { auto it = m.find(k); return it == m.end() ? 0 : atoi(it->second.c_str()); }

static int load_cloud(const char* path, int& N, POINT3D** p) {
    std::ifstream in(path);
    if (!in.is_open()) { std::cerr << "Unable to open point file " << path << std::endl; exit(-1); }   // jly_main.cpp:287
    in >> N;
    *p = new POINT3D[N];
    for (int i = 0; i < N; i++) { in >> (*p)[i].x >> (*p)[i].y >> (*p)[i].z; (*p)[i].c = 0; (*p)[i].neighbors = 0; (*p)[i].density = 0; }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 6) { std::cerr << "usage: goicp_demo <model> <data> <NdDownsampled> <config> <output>" << std::endl; return 1; }
    GoICP goicp;
    auto cfg = read_config(argv[4]);   // readConfig, jly_main.cpp:231-270
    goicp.MSEThresh = (float)getF(cfg, "MSEThresh");
    goicp.initNodeRot.a = (float)getF(cfg, "rotMinX"); goicp.initNodeRot.b = (float)getF(cfg, "rotMinY"); goicp.initNodeRot.c = (float)getF(cfg, "rotMinZ");
    goicp.initNodeRot.w = (float)getF(cfg, "rotWidth");
    goicp.initNodeTrans.x = (float)getF(cfg, "transMinX"); goicp.initNodeTrans.y = (float)getF(cfg, "transMinY"); goicp.initNodeTrans.z = (float)getF(cfg, "transMinZ");
    goicp.initNodeTrans.w = (float)getF(cfg, "transWidth");
    goicp.trimFraction = (float)getF(cfg, "trimFraction");
    if (goicp.trimFraction < 0.001) goicp.doTrim = false;
    goicp.regularization = (float)getF(cfg, "regularization"); goicp.regularizationNeighbors = (float)getF(cfg, "regularizationNeighbors");
    goicp.regularizationFPFH = (float)getF(cfg, "regularizationFPFH");
    goicp.cfpfh = getI(cfg, "cfpfh"); goicp.norm = getI(cfg, "norm") ? getI(cfg, "norm") : 2; goicp.ponderation = getI(cfg, "ponderation");
    goicp.dt.SIZE = getI(cfg, "distTransSize"); goicp.dt.expandFactor = getF(cfg, "distTransExpandFactor");

    load_cloud(argv[1], goicp.Nm, &goicp.pModel);
    load_cloud(argv[2], goicp.Nd, &goicp.pData);
    const int NdDownsampled = atoi(argv[3]);

    auto t0 = std::chrono::steady_clock::now();
    goicp.BuildDT();
    auto t1 = std::chrono::steady_clock::now();
    if (NdDownsampled > 0 && NdDownsampled < goicp.Nd) goicp.Nd = NdDownsampled;   // jly_main.cpp:114-117
    goicp.Register();
    auto t2 = std::chrono::steady_clock::now();
    const double dtBuild = std::chrono::duration<double>(t1 - t0).count(), dtReg = std::chrono::duration<double>(t2 - t1).count();
    std::cout << goicp.Trace();
    std::cout << "Optimal Rotation Matrix:" << std::endl;
    for (int i = 0; i < 3; i++) printf("%12.7f %12.7f %12.7f\n", goicp.optR_at(i, 0), goicp.optR_at(i, 1), goicp.optR_at(i, 2));   // Matrix operator<< %12.7f (matrix.cpp:812)
    std::cout << "Optimal Translation Vector:" << std::endl;
    for (int i = 0; i < 3; i++) printf("%12.7f\n", goicp.optT_at(i));
    printf("DT build %.4f s, Register %.4f s, optError %.9g, Compatibilities %d\n", dtBuild, dtReg, goicp.optError, goicp.Nd - goicp.optComp);

    FILE* f = fopen(argv[5], "w");
    if (f) {
        fprintf(f, "%g\n", dtReg);
        for (int i = 0; i < 3; i++) fprintf(f, "%12.7f %12.7f %12.7f\n", goicp.optR_at(i, 0), goicp.optR_at(i, 1), goicp.optR_at(i, 2));
        for (int i = 0; i < 3; i++) fprintf(f, "%12.7f\n", goicp.optT_at(i));
        fclose(f);
    }
    delete[] goicp.pModel; delete[] goicp.pData;
    return 0;
}
