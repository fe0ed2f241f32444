// File formats of the reference pipeline shared by the two command lines (GoICP_b200, GoICP_b200_sweep): mol2 atoms, the
// normalised xyzc text with its 6-significant-digit round trip, cfpfh descriptors, config.txt, and the two
// per-pair result files.  Host-side parsing / formatting only.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "goicp_dropin.hpp"

using std::string;

// colour codes of the `properties` enum (transformation.hpp:36); unknown atom names map to OG (transformation.cpp:46)
static int atom_colour(const string& name) {
    static const std::map<string, int> table = {{"OG", 8204959}, {"N", 30894}, {"O", 15219528}, {"NZ", 15231913}, {"CZ", 4646984},
                                                {"CA", 16741671}, {"DU", 7566712}, {"OD1", 0}, {"C", 1}};
    auto it = table.find(name);
    return it == table.end() ? 8204959 : it->second;
}

// The @<TRIPOS>ATOM block of a mol2 file: id, name, x, y, z per atom (transformation.cpp:282-306 keeps every atom
// whose five leading fields parse and drops the first record that does not).
static std::vector<point4D> read_mol2_atoms(const string& path) {
    std::vector<point4D> cloud;
    std::ifstream in(path);
    string line;
    while (std::getline(in, line))
        if (line.find("@<TRIPOS>ATOM") != string::npos) break;
    string id, name;
    point4D p{};
    while (in >> id >> name >> p.x >> p.y >> p.z) {
        p.c = atom_colour(name);
        cloud.push_back(p);
        std::getline(in, line);   // rest of the atom record
    }
    return cloud;
}

// writeNormalizedMolCloudFile (transformation.cpp:340-350): default ostream formatting = 6 significant digits
static void write_xyzc(const string& path, const std::vector<point4D>& cloud) {
    std::ofstream out(path);
    out << cloud.size() << std::endl;
    for (const point4D& p : cloud) out << p.x << " " << p.y << " " << p.z << " " << p.c << std::endl;
}

// ConfigMap (ConfigMap.cpp:3-151): key=value, '#' comments, tokens split on " =;", lines without exactly two tokens ignored
static std::map<string, string> read_config(const string& path) {
    std::ifstream in(path);
    if (!in.is_open()) { std::cout << "Unable to open config file '" << path << "'" << std::endl; exit(-2); }
    std::map<string, string> m;
    string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        const size_t h = line.find('#');
        if (h != string::npos) line = line.substr(0, h);
        std::vector<string> tok; string cur;
        for (char ch : line) { if (ch == ' ' || ch == '=' || ch == ';') { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } } else cur += ch; }
        if (!cur.empty()) tok.push_back(cur);
        if (tok.size() == 2) m[tok[0]] = tok[1];
    }
    return m;
}
static double cfgF(const std::map<string, string>& m, const char* k) { auto it = m.find(k); return it == m.end() ? 0.0 : atof(it->second.c_str()); }
static int cfgI(const std::map<string, string>& m, const char* k) { auto it = m.find(k); return it == m.end() ? 0 : atoi(it->second.c_str()); }

// loadPointCloud (jly_main.cpp:272-314): the normalised xyzc text + cfpfh/<id>.cfpfh (41 bins per point)
static void load_cloud(const string& fname, int& N, POINT3D** p) {
    const string fp = "cfpfh/" + fname.substr(fname.find("/") + 1, fname.find_last_of("_") - fname.find("/") - 1) + ".cfpfh";
    std::ifstream in(fname), fin(fp);
    if (!in.is_open()) { std::cout << "Unable to open point file '" << fname << "'" << std::endl; exit(-1); }
    in >> N;
    *p = new POINT3D[N];
    if (!fin.is_open()) { std::cout << "Unable to open fpfh file '" << fp << "'" << std::endl; exit(-1); }
    for (int i = 0; i < N; i++) {
        POINT3D& q = (*p)[i];
        in >> q.x >> q.y >> q.z >> q.c;
        q.neighbors = 0; q.density = 0;
        q.cfpfh.resize(41);
        for (int j = 0; j < 41; j++) { float b = 0; fin >> b; q.cfpfh[j] = b; }
    }
}

static string stem_between(const string& s) { return s.substr(s.find("/") + 1, s.find(".") - s.find("/") - 1); }   // jly_main.cpp:68-69

// Matrix operator<< (matrix.cpp:812-827): "%12.7f " per entry
static string matrix_rows(const double* v, int rows, int cols) {
    std::ostringstream o; char b[64];
    for (int i = 0; i < rows; i++) { for (int j = 0; j < cols; j++) { snprintf(b, sizeof b, "%12.7f ", v[i * cols + j]); o << b; } if (i < rows - 1) o << std::endl; }
    return o.str();
}
// <output> (jly_main.cpp:131-141)
static void write_output_file(const string& path, double time, const double* R, const double* t, float optError, int compat) {
    std::ofstream ofile(path.c_str());
    ofile << "Time: " << time << std::endl;
    ofile << "Rotation Matrix: " << std::endl << matrix_rows(R, 3, 3) << std::endl;
    ofile << "Translation Vector: " << std::endl << matrix_rows(t, 3, 1) << std::endl;
    ofile << "Error: " << optError << std::endl;
    ofile << "Compatibilities: " << compat << std::endl;
}
// <output-stem>_rescaled.txt (transformation.cpp:403-417)
static void write_rescaled_file(const string& outputF, double time, const double* R, const double* tr, float optError) {
    std::ofstream rf(outputF.substr(0, outputF.find(".")) + "_rescaled.txt");
    rf << "Time: " << time << std::endl;
    rf << "Rotation Matrix:" << std::endl << "   " << R[0] << "   " << R[1] << "   " << R[2] << std::endl;
    rf << "   " << R[3] << "   " << R[4] << "   " << R[5] << std::endl;
    rf << "   " << R[6] << "   " << R[7] << "   " << R[8] << std::endl;
    rf << "Translation Vector:" << std::endl << "   " << tr[0] << std::endl << "   " << tr[1] << std::endl << "   " << tr[2] << std::endl;
    rf << "Error: " << optError << std::endl;
}
// every config.txt key (readConfig jly_main.cpp:231-270) into the C-ABI parameter block
static goicp_params params_from_config(const std::map<string, string>& cfg) {
    goicp_params p; goicp_params_default(&p);
    p.MSEThresh = (float)cfgF(cfg, "MSEThresh");
    p.rotMinX = (float)cfgF(cfg, "rotMinX"); p.rotMinY = (float)cfgF(cfg, "rotMinY"); p.rotMinZ = (float)cfgF(cfg, "rotMinZ"); p.rotWidth = (float)cfgF(cfg, "rotWidth");
    p.transMinX = (float)cfgF(cfg, "transMinX"); p.transMinY = (float)cfgF(cfg, "transMinY"); p.transMinZ = (float)cfgF(cfg, "transMinZ"); p.transWidth = (float)cfgF(cfg, "transWidth");
    p.trimFraction = (float)cfgF(cfg, "trimFraction");
    p.regularization = (float)cfgF(cfg, "regularization"); p.regularizationNeighbors = (float)cfgF(cfg, "regularizationNeighbors"); p.regularizationFPFH = (float)cfgF(cfg, "regularizationFPFH");
    p.cfpfh = cfgI(cfg, "cfpfh"); p.norm = cfgI(cfg, "norm"); p.ponderation = cfgI(cfg, "ponderation");
    p.distTransSize = cfgI(cfg, "distTransSize"); p.distTransExpandFactor = cfgF(cfg, "distTransExpandFactor");
    return p;
}
